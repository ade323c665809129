#!/bin/bash
# registers / stack / shared memory per kernel of an object file: scripts/kernel_resources.sh rfrt_trace.o [pattern]
cuobjdump --dump-resource-usage "$1" 2>/dev/null | paste - - | grep "Function" | sed 's/Function \(.*\):.*REG:\([0-9]*\) STACK:\([0-9]*\) SHARED:\([0-9]*\).*/\2 regs \3 stack \4 smem \1/' | while read r a st b sm c name; do echo "$r regs $st stack $(echo $name | c++filt | sed 's/rfrt::(anonymous namespace):://; s/(.*//')"; done | grep "${2:-.}"

// rfrt_small.cu — host-side tables of the small-scene (<= 64 triangles) candidate filter.
//
// The lockstep sweep of rfrt_trace.cuh finds, for every ray segment, a SUPERSET of the triangles the exact
// watertight test (kernel.py:82 -> intersect_ray_tri_woop) can accept, with far fewer instructions than the exact
// test itself; the exact test then runs on the 1-3 survivors only.  The filter works per supporting plane:
//   plane   : unit normal n, offset d  (n.x = d on the plane) -> ray parameter t and hit point h = p + t*dir
//   triangle: three in-plane signed edge distances e_i(h) = m_i.h + c_i (metres, positive inside)
// A triangle stays a candidate unless min_i e_i(h) < -tol or its plane lies behind the origin by more than the
// tolerance, where tol = 2^-16 * (scene extent + |p|_1) / sin(angle between ray and plane) bounds both the
// rounding of the exact test (a few 2^-24 * extent perpendicular to the ray) and of this filter.
// Two coplanar triangles share one plane entry (a rectangular wall = one record).
#include <cmath>
#include <cstring>
#include <vector>

#include "rfrt_internal.h"

namespace rfrt {

namespace {

struct V3 {
    double x, y, z;
};
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double len(V3 a) { return std::sqrt(dot(a, a)); }

} // namespace

// recs     : [n_pairs*28]  per pair of coplanar triangles: (n.xyz, d) + 2 x (m0.xyz, c0, m1.xyz, c1, m2.xyz, c2)
// slot_tri : [n_pairs*2]   slot -> original triangle index (slot 2k, 2k+1 = the triangles of pair k)
// Triangles are paired inside their supporting plane (index order); the odd one out is paired with itself (a
// duplicate candidate is harmless).  A degenerate triangle (no usable normal) gets n = 0: every lane then computes
// NaN and keeps it as a candidate, so the exact test alone decides (it never accepts a zero-area triangle, but a
// sliver might be hit).  Returns RFRT_ERR_INVALID when the scene needs more than RFRT_SMALL_MAX_TRIS slots.
int small_scene_tables(const float *soup, int n_tris, float *recs, int32_t *slot_tri, int32_t *n_pairs, float *extent)
{
    if (n_tris < 0 || n_tris > RFRT_SMALL_MAX_TRIS) return RFRT_ERR_INVALID;
    struct Pl {
        V3 n;
        double d;
        std::vector<int> tris;
        bool degenerate;
    };
    std::vector<Pl> pls;
    double ext = 0.0;
    for (int i = 0; i < 9 * n_tris; ++i) ext = std::fmax(ext, std::fabs((double)soup[i]));
    for (int f = 0; f < n_tris; ++f) {
        const float *v = soup + 9 * f;
        V3 a{v[0], v[1], v[2]}, b{v[3], v[4], v[5]}, c{v[6], v[7], v[8]};
        V3 n = cross(sub(b, a), sub(c, a));
        double l = len(n);
        double emax = std::fmax(len(sub(b, a)), std::fmax(len(sub(c, a)), len(sub(c, b))));
        // height of the triangle over its longest edge, relative to that edge: below 1e-5 the fp32 plane is not
        // trustworthy to the filter's tolerance
        bool degenerate = !(l > 1.0e-5 * emax * emax) || !std::isfinite(l);
        Pl p;
        p.degenerate = degenerate;
        if (degenerate) {
            p.n = {0, 0, 0}; p.d = 0;
            p.tris.push_back(f);
            pls.push_back(p);
            continue;
        }
        n = {n.x / l, n.y / l, n.z / l};
        // canonical sign: first component of magnitude > 1e-9 is positive
        double lead = std::fabs(n.x) > 1e-9 ? n.x : (std::fabs(n.y) > 1e-9 ? n.y : n.z);
        if (lead < 0) n = {-n.x, -n.y, -n.z};
        double d = dot(n, a);
        int found = -1;
        for (size_t k = 0; k < pls.size(); ++k) {
            if (pls[k].degenerate) continue;
            const Pl &q = pls[k];
            if (std::fabs(q.n.x - n.x) < 1e-9 && std::fabs(q.n.y - n.y) < 1e-9 && std::fabs(q.n.z - n.z) < 1e-9 &&
                std::fabs(q.d - d) <= 1e-9 * (1.0 + ext) &&
                // all three vertices on q's plane (guards the grouping against a tilted near-parallel neighbour)
                std::fabs(dot(q.n, b) - q.d) <= 1e-9 * (1.0 + ext) && std::fabs(dot(q.n, c) - q.d) <= 1e-9 * (1.0 + ext)) {
                found = (int)k;
                break;
            }
        }
        if (found >= 0) pls[found].tris.push_back(f);
        else { p.n = n; p.d = d; p.tris.push_back(f); pls.push_back(p); }
    }
    int need = 0;
    for (const Pl &p : pls) need += ((int)p.tris.size() + 1) / 2;
    if (2 * need > RFRT_SMALL_MAX_TRIS) return RFRT_ERR_INVALID;
    int pair = 0;
    for (const Pl &p : pls) {
        for (size_t j = 0; j < p.tris.size(); j += 2) {
            float *R = recs + 28 * pair;
            R[0] = (float)p.n.x; R[1] = (float)p.n.y; R[2] = (float)p.n.z; R[3] = (float)p.d;
            for (int h = 0; h < 2; ++h) {
                const int f = p.tris[j + h < p.tris.size() ? j + h : j];
                const float *v = soup + 9 * f;
                V3 vv[3] = {{v[0], v[1], v[2]}, {v[3], v[4], v[5]}, {v[6], v[7], v[8]}};
                float *E = R + 4 + 12 * h;
                for (int i = 0; i < 3; ++i) {
                    V3 m{0, 0, 0};
                    double c = 0.0;
                    if (!p.degenerate) {
                        V3 e = sub(vv[(i + 1) % 3], vv[i]);
                        m = cross(p.n, e);
                        double ml = len(m);
                        m = {m.x / ml, m.y / ml, m.z / ml};
                        if (dot(m, sub(vv[(i + 2) % 3], vv[i])) < 0) m = {-m.x, -m.y, -m.z}; // positive towards the third vertex
                        c = -dot(m, vv[i]);
                    }
                    E[4 * i] = (float)m.x; E[4 * i + 1] = (float)m.y; E[4 * i + 2] = (float)m.z; E[4 * i + 3] = (float)c;
                }
                slot_tri[2 * pair + h] = f;
            }
            ++pair;
        }
    }
    *n_pairs = pair;
    *extent = (float)ext;
    return RFRT_OK;
}

} // namespace rfrt

extern "C" int rfrt_small_scene_tables(const float *h_soup, int32_t n_triangles, float *h_recs, int32_t *h_slot_tri,
                                       int32_t *n_pairs, float *extent)
{
    if (!h_soup || !h_recs || !h_slot_tri || !n_pairs || !extent) {
        rfrt::set_error("rfrt_small_scene_tables: null argument");
        return RFRT_ERR_INVALID;
    }
    int rc = rfrt::small_scene_tables(h_soup, n_triangles, h_recs, h_slot_tri, n_pairs, extent);
    if (rc) rfrt::set_error("rfrt_small_scene_tables: the scene does not fit 64 filter slots");
    return rc;
}

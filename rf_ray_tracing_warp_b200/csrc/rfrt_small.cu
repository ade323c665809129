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
inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 mul(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double clamp01(double x) { return x < 0 ? 0 : (x > 1 ? 1 : x); }

// distance point - triangle (closest-point regions, Ericson 5.1.5)
double dist_point_tri(V3 p, V3 a, V3 b, V3 c)
{
    V3 ab = sub(b, a), ac = sub(c, a), ap = sub(p, a);
    double d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) return len(ap);
    V3 bp = sub(p, b);
    double d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0 && d4 <= d3) return len(bp);
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) return len(sub(p, add(a, mul(ab, d1 / (d1 - d3)))));
    V3 cp = sub(p, c);
    double d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) return len(cp);
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) return len(sub(p, add(a, mul(ac, d2 / (d2 - d6)))));
    double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0)
        return len(sub(p, add(b, mul(sub(c, b), (d4 - d3) / ((d4 - d3) + (d5 - d6))))));
    double denom = 1.0 / (va + vb + vc);
    return len(sub(p, add(a, add(mul(ab, vb * denom), mul(ac, vc * denom)))));
}

// distance segment - segment (Ericson 5.1.9)
double dist_seg_seg(V3 p1, V3 q1, V3 p2, V3 q2)
{
    V3 d1 = sub(q1, p1), d2 = sub(q2, p2), r = sub(p1, p2);
    double a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), s, t;
    const double EPS = 1e-300;
    if (a <= EPS && e <= EPS) return len(r);
    if (a <= EPS) { s = 0; t = clamp01(f / e); }
    else {
        double c = dot(d1, r);
        if (e <= EPS) { t = 0; s = clamp01(-c / a); }
        else {
            double b = dot(d1, d2), denom = a * e - b * b;
            s = denom > 0 ? clamp01((b * f - c * e) / denom) : 0;
            t = (b * s + f) / e;
            if (t < 0) { t = 0; s = clamp01(-c / a); }
            else if (t > 1) { t = 1; s = clamp01((b - c) / a); }
        }
    }
    return len(sub(add(p1, mul(d1, s)), add(p2, mul(d2, t))));
}

// distance segment - triangle: 0 when the segment pierces the triangle
double dist_seg_tri(V3 p, V3 q, V3 a, V3 b, V3 c)
{
    V3 n = cross(sub(b, a), sub(c, a));
    double sp = dot(n, sub(p, a)), sq = dot(n, sub(q, a));
    if ((sp <= 0 && sq >= 0) || (sp >= 0 && sq <= 0)) {
        double den = sp - sq;
        V3 x = den != 0 ? add(p, mul(sub(q, p), sp / den)) : p;
        if (dist_point_tri(x, a, b, c) <= 1e-12 * (1.0 + len(x))) return 0.0;
    }
    double d = std::fmin(dist_point_tri(p, a, b, c), dist_point_tri(q, a, b, c));
    d = std::fmin(d, dist_seg_seg(p, q, a, b));
    d = std::fmin(d, dist_seg_seg(p, q, b, c));
    return std::fmin(d, dist_seg_seg(p, q, c, a));
}

double dist_tri_tri(const V3 *A, const V3 *B)
{
    double d = 1e300;
    for (int i = 0; i < 3; ++i) {
        d = std::fmin(d, dist_seg_tri(A[i], A[(i + 1) % 3], B[0], B[1], B[2]));
        d = std::fmin(d, dist_seg_tri(B[i], B[(i + 1) % 3], A[0], A[1], A[2]));
    }
    return d;
}

} // namespace

// Neighbour pair masks of the self-re-hit shortcut, 4 words per triangle f (bit k = pair k holds a triangle g != f that
// comes within `reach` of f):
//   nbr[4f+0] interior neighbours: g also comes within reach of f ERODED by 2*reach (T-junctions, overlaps, slivers)
//   nbr[4f+1] boundary neighbours: g only comes near the rim of f — irrelevant for a point of f that keeps a
//             clearance of 2*reach (+ tolerance) from f's three edges
//   nbr[4f+2], nbr[4f+3]: the same restricted to g < f (all that matters when f is hit at t == 0: ties go to the
//             lowest index)
// tri_slot[f] : a filter slot of triangle f (its edge functions give the clearance)
void small_scene_neighbours(const float *soup, int n_tris, const int32_t *slot_tri, int n_pairs, double reach, uint32_t *nbr,
                            int32_t *tri_slot)
{
    for (int f = 0; f < n_tris; ++f) {
        const float *v = soup + 9 * f;
        V3 A[3] = {{v[0], v[1], v[2]}, {v[3], v[4], v[5]}, {v[6], v[7], v[8]}};
        // f eroded by 2*reach: homothety about the incentre (empty when the inradius is too small)
        V3 E[3];
        bool eroded = false;
        {
            double la = len(sub(A[2], A[1])), lb = len(sub(A[0], A[2])), lc = len(sub(A[1], A[0]));
            double per = la + lb + lc;
            double area2 = len(cross(sub(A[1], A[0]), sub(A[2], A[0])));
            if (per > 0 && std::isfinite(per)) {
                double rin = area2 / per; // inradius = 2*area / perimeter
                if (rin > 2.02 * reach) {
                    V3 I = mul(add(add(mul(A[0], la), mul(A[1], lb)), mul(A[2], lc)), 1.0 / per);
                    double sc = 1.0 - 2.0 * reach / rin;
                    for (int i = 0; i < 3; ++i) E[i] = add(I, mul(sub(A[i], I), sc));
                    eroded = true;
                }
            }
        }
        uint32_t m_in = 0, m_bd = 0, l_in = 0, l_bd = 0;
        for (int k = 0; k < n_pairs; ++k)
            for (int h = 0; h < 2; ++h) {
                const int g = slot_tri[2 * k + h];
                if (g == f) { if (tri_slot) tri_slot[f] = 2 * k + h; continue; }
                const float *w = soup + 9 * g;
                V3 B[3] = {{w[0], w[1], w[2]}, {w[3], w[4], w[5]}, {w[6], w[7], w[8]}};
                if (dist_tri_tri(A, B) > reach) continue; // (NaN -> neighbour)
                const bool interior = !eroded || !(dist_tri_tri(E, B) > reach);
                if (interior) { m_in |= 1u << k; if (g < f) l_in |= 1u << k; }
                else { m_bd |= 1u << k; if (g < f) l_bd |= 1u << k; }
            }
        m_bd &= ~m_in; l_bd &= ~l_in; // a pair with one interior neighbour is always visited
        nbr[4 * f] = m_in; nbr[4 * f + 1] = m_bd; nbr[4 * f + 2] = l_in; nbr[4 * f + 3] = l_bd;
    }
}

// recs     : [n_pairs*28]  per pair of coplanar triangles: (n.xyz, d) + 2 x (m0.xyz, c0, m1.xyz, c1, m2.xyz, c2)
// slot_tri : [n_pairs*2]   slot -> original triangle index (slot 2k, 2k+1 = the triangles of pair k)
// Triangles are paired inside their supporting plane (index order); the odd one out is paired with itself (a
// duplicate candidate is harmless).  A degenerate triangle (no usable normal) gets n = 0: every lane then computes
// NaN and keeps it as a candidate, so the exact test alone decides (it never accepts a zero-area triangle, but a
// sliver might be hit).  Returns RFRT_ERR_INVALID when the scene needs more than RFRT_SMALL_MAX_TRIS slots.
// class_begin[5]: the pairs are ordered by the class of their plane — general (incl. degenerate) first, then planes
// normal to x, y, z — and class c occupies pairs [class_begin[c], class_begin[c+1]).  The device sweeps axis-aligned
// planes with the zero terms of the general formulas left out (same values, two thirds of the instructions).
int small_scene_tables(const float *soup, int n_tris, float *recs, int32_t *slot_tri, int32_t *n_pairs, float *extent,
                       int32_t *class_begin)
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
    auto plane_class = [](const Pl &p) {
        if (p.degenerate) return 0;
        const bool zx = p.n.x == 0.0, zy = p.n.y == 0.0, zz = p.n.z == 0.0;
        if (zy && zz && p.n.x == 1.0) return 1;
        if (zx && zz && p.n.y == 1.0) return 2;
        if (zx && zy && p.n.z == 1.0) return 3;
        return 0;
    };
    int pair = 0;
    for (int cls = 0; cls < 4; ++cls) {
      if (class_begin) class_begin[cls] = pair;
      for (const Pl &p : pls) {
        if (plane_class(p) != cls) continue;
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
    }
    if (class_begin) class_begin[4] = pair;
    *n_pairs = pair;
    *extent = (float)ext;
    return RFRT_OK;
}

// Receiver face records in UNIT space (one per face of the unit shape, shared by all receivers of a set):
// recs[16*f] = plane (n.xyz, d) + three in-plane edge functions (m.xyz, c) — the candidate filter of the receiver query
// in the replay kernel (rx_query_sweep).  A degenerate face gets n = 0 (always a candidate).
void unit_face_records(const double *unit_v, const int32_t *faces, int n_faces, float *recs)
{
    for (int f = 0; f < n_faces; ++f) {
        V3 vv[3];
        for (int c = 0; c < 3; ++c) vv[c] = {unit_v[3 * faces[3 * f + c]], unit_v[3 * faces[3 * f + c] + 1], unit_v[3 * faces[3 * f + c] + 2]};
        float *R = recs + 16 * f;
        V3 n = cross(sub(vv[1], vv[0]), sub(vv[2], vv[0]));
        double l = len(n);
        double emax = std::fmax(len(sub(vv[1], vv[0])), std::fmax(len(sub(vv[2], vv[0])), len(sub(vv[2], vv[1]))));
        const bool degenerate = !(l > 1.0e-5 * emax * emax) || !std::isfinite(l);
        if (degenerate) {
            for (int i = 0; i < 16; ++i) R[i] = 0.0f;
            continue;
        }
        n = {n.x / l, n.y / l, n.z / l};
        R[0] = (float)n.x; R[1] = (float)n.y; R[2] = (float)n.z; R[3] = (float)dot(n, vv[0]);
        for (int i = 0; i < 3; ++i) {
            V3 e = sub(vv[(i + 1) % 3], vv[i]);
            V3 m = cross(n, e);
            double ml = len(m);
            m = {m.x / ml, m.y / ml, m.z / ml};
            if (dot(m, sub(vv[(i + 2) % 3], vv[i])) < 0) m = {-m.x, -m.y, -m.z};
            R[4 + 4 * i] = (float)m.x; R[5 + 4 * i] = (float)m.y; R[6 + 4 * i] = (float)m.z; R[7 + 4 * i] = (float)(-dot(m, vv[i]));
        }
    }
}

} // namespace rfrt

extern "C" int rfrt_small_scene_tables(const float *h_soup, int32_t n_triangles, float *h_recs, int32_t *h_slot_tri,
                                       int32_t *n_pairs, float *extent, uint32_t *h_nbr, int32_t *h_class_begin)
{
    if (!h_soup || !h_recs || !h_slot_tri || !n_pairs || !extent) {
        rfrt::set_error("rfrt_small_scene_tables: null argument");
        return RFRT_ERR_INVALID;
    }
    int rc = rfrt::small_scene_tables(h_soup, n_triangles, h_recs, h_slot_tri, n_pairs, extent, h_class_begin);
    if (rc) { rfrt::set_error("rfrt_small_scene_tables: the scene does not fit 64 filter slots"); return rc; }
    if (h_nbr)
        rfrt::small_scene_neighbours(h_soup, n_triangles, h_slot_tri, *n_pairs, rfrt::SMALL_REACH_REL * (double)*extent, h_nbr, nullptr);
    return rc;
}

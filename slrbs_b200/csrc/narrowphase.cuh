// narrowphase.cuh -- the reference's three narrow-phase routines and its pair dispatch chain
// (CollisionDetect.cpp:41-73,102-274), written against the staged body proxies. Bit-exact with the
// reference arithmetic (vecmath.cuh associations, -fmad=false).
#pragma once
#include "layout.cuh"
#include "vecmath.cuh"
#include "sdf.cuh"

namespace slrbs {

struct PairOut {
    int b0, b1, cnt;        // role-ordered body indices (sphere/box, cylinder/plane), contacts of this pair
    V3 n;
    V3 p[4];
    float phi[4];
    __device__ __forceinline__ void add(const V3& pp, float f) { p[cnt] = pp; phi[cnt] = f; ++cnt; }
};

__device__ __forceinline__ float plane_dist(const V3& p, const V3& plane_p, const V3& plane_n)
{
    return dot(p - plane_p, plane_n);                            // CollisionDetect.cpp:12-17
}

// CollisionDetect.cpp:102-123
__device__ __forceinline__ void sphere_sphere(PairOut& o, int i0, int i1, const float4& c0, const float4& c1)
{
    const V3 x0 = mk3(c0), x1 = mk3(c1);
    const V3 vec = x0 - x1;
    const float rsum = (c0.w + c1.w);
    const float dist = norm(vec);
    if (dist < (rsum + 1e-3f)) {
        o.b0 = i0; o.b1 = i1;
        o.n = vec / dist;
        o.add(0.5f * ((x0 - c0.w * o.n) + (x1 + c1.w * o.n)), fminf(0.0f, dist - rsum));
    }
}

// CollisionDetect.cpp:125-150 (body0 = sphere, body1 = box)
static __device__ __noinline__ void sphere_box(PairOut& o, const float4* __restrict__ quat, const float4* __restrict__ geomA, size_t bi,
                                       int is, int ib, const float4& cs, const float4& cb)
{
    const V3 xs = mk3(cs), xb = mk3(cb);
    const float r = cs.w;
    const Q4 qb = mkq(__ldg(&quat[bi]));
    const float4 dim = __ldg(&geomA[bi]);
    const V3 cl = rotate(qinverse(qb), xs - xb);
    V3 q;
    q.x = cl.x; if (q.x < (-dim.x / 2.0f)) q.x = -dim.x / 2.0f; else if (q.x > (dim.x / 2.0f)) q.x = dim.x / 2.0f;
    q.y = cl.y; if (q.y < (-dim.y / 2.0f)) q.y = -dim.y / 2.0f; else if (q.y > (dim.y / 2.0f)) q.y = dim.y / 2.0f;
    q.z = cl.z; if (q.z < (-dim.z / 2.0f)) q.z = -dim.z / 2.0f; else if (q.z > (dim.z / 2.0f)) q.z = dim.z / 2.0f;
    const V3 dx = cl - q;
    const float dist = norm(dx);
    if (dist < (r + 1e-3f)) {
        o.b0 = is; o.b1 = ib;
        // dx has two exactly-zero components at a face contact: a zero numerator sends the compiler's division through its
        // out-of-range subroutine, twice per wall contact
        o.n = rotate(qb, div_by_length(dx, dist));
        o.add(rotate(qb, q) + xb, fminf(0.0f, dist - r));
    }
}

// CollisionDetect.cpp:152-274 (body0 = cylinder, body1 = plane)
static __device__ __noinline__ void cylinder_plane(PairOut& o, const float4* __restrict__ quat, const float4* __restrict__ geomA,
                                           const float4* __restrict__ geomB, size_t ci, size_t pi, int ic, int ip,
                                           const float4& cc, const float4& cpl)
{
    const V3 xc = mk3(cc), xp = mk3(cpl);
    const Q4 qc = mkq(__ldg(&quat[ci]));
    const float4 cg = __ldg(&geomA[ci]);
    const float h = cg.x, r = cg.y;
    const Q4 qp = mkq(__ldg(&quat[pi]));
    const V3 cyldir = rotate(qc, mk3(0.f, 1.f, 0.f));
    const V3 planen = rotate(qp, mk3(__ldg(&geomB[pi])));
    const V3 planep = rotate(qp, mk3(__ldg(&geomA[pi]))) + xp;
    const float dp = dot(cyldir, planen);
    o.b0 = ic; o.b1 = ip; o.n = planen;

    if (fabsf(dp) > 0.995f) {                                    // :164 cap face on the plane
        const V3 w = (dp < 0.0f) ? cyldir : -cyldir;
        V3 u = (fabsf(dot(w, mk3(1.f, 0.f, 0.f))) > 0.01f) ? cross(w, mk3(1.f, 0.f, 0.f)) : cross(w, mk3(0.f, 0.f, 1.f));
        u = normalized(u);
        const V3 vv = normalized(cross(w, u));
        const V3 a = xc + (0.5f * h) * w;
        if (plane_dist(a, planep, planen) < 1e-3f) {
            const V3 pA = a + r * u, pB = a - r * u, pC = a + r * vv, pD = a - r * vv;
            const float fA = plane_dist(pA, planep, planen), fB = plane_dist(pB, planep, planen);
            const float fC = plane_dist(pC, planep, planen), fD = plane_dist(pD, planep, planen);
            if (fA < 1e-3f) o.add(pA, fminf(0.0f, fA));
            if (fB < 1e-3f) o.add(pB, fminf(0.0f, fB));
            if (fC < 1e-3f) o.add(pC, fminf(0.0f, fC));
            if (fD < 1e-3f) o.add(pD, fminf(0.0f, fD));
        }
    } else if (fabsf(dp) < 1e-2f) {                              // :215 lying on its side
        V3 u = cross(cross(planen, cyldir), cyldir);
        if (dot(u, planen) > 0.0f) u = -u;
        u = normalized(u);
        const float dist = plane_dist(xc, planep, planen) - r;
        if (dist < 0.01f) {                                      // :227
            const V3 hw = (0.5f * h) * cyldir, ru = r * u;
            const V3 pA = (xc + hw) + ru, pB = (xc - hw) + ru;
            const float fA = plane_dist(pA, planep, planen), fB = plane_dist(pB, planep, planen);
            if (fA < 1e-3f) o.add(pA, fminf(0.0f, fA));
            if (fB < 1e-3f) o.add(pB, fminf(0.0f, fB));
        }
    } else {                                                     // :241 single lowest rim point
        V3 w = (dp < 0.0f) ? cyldir : -cyldir;
        w = normalized(w);
        V3 u = normalized(cross(cross(planen, w), w));
        if (dot(u, planen) > 0.0f) u = -u;
        u = normalized(u);
        const V3 a = (xc + (0.5f * h) * w) + r * u;
        const float dist = plane_dist(a, planep, planen);
        if (dist < 1e-3f) o.add(a, fminf(0.0f, dist));
    }
}

// Conservative cull for sphere vs box: the sphere centre must lie inside the box's world AABB grown by the
// radius and the 1e-3 contact margin for the exact test (CollisionDetect.cpp:142) to succeed; the extra
// slack (1e-3 absolute + a few ulps of the coordinates) covers the rounding of the AABB itself.
__device__ __forceinline__ bool sphere_box_may_touch(const float4& cs, const float4& cb, const float4& ext)
{
    const float r = cs.w + 2e-3f;
    const float dx = fabsf(cs.x - cb.x), dy = fabsf(cs.y - cb.y), dz = fabsf(cs.z - cb.z);
    return dx <= ext.x + r + 4e-6f * (fabsf(cs.x) + fabsf(cb.x) + ext.x) &&
           dy <= ext.y + r + 4e-6f * (fabsf(cs.y) + fabsf(cb.y) + ext.y) &&
           dz <= ext.z + r + 4e-6f * (fabsf(cs.z) + fabsf(cb.z) + ext.z);
}

// ------------------------------------------------------------------------------------
// EXTENSION pairs: sphere-plane, box-plane, box-box. The reference ignores these pairs
// (CollisionDetect.cpp:45-73), so there is no reference behaviour to match; they are off unless
// slrbs_set_extensions(ctx, 1) and are appended AFTER the reference's chain, which is unchanged. Contact
// convention as everywhere: normal from body1 to body0, at most 4 contacts per pair.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float proj_radius(const V3* ax, const float* h, const V3& L)
{
    return sum3(h[0] * fabsf(dot(ax[0], L)), h[1] * fabsf(dot(ax[1], L)), h[2] * fabsf(dot(ax[2], L)));
}

// sphere (body0) against the half space of a plane (body1)
__device__ __forceinline__ void ext_sphere_plane(PairOut& o, const float4* __restrict__ quat, const float4* __restrict__ geomA,
                                                 const float4* __restrict__ geomB, size_t pi, int is, int ip, const float4& cs, const float4& cpl)
{
    const Q4 qp = mkq(__ldg(&quat[pi]));
    const V3 planen = rotate(qp, mk3(__ldg(&geomB[pi])));
    const V3 planep = rotate(qp, mk3(__ldg(&geomA[pi]))) + mk3(cpl);
    const V3 xs = mk3(cs);
    const float dist = plane_dist(xs, planep, planen) - cs.w;
    if (dist < 1e-3f) { o.b0 = is; o.b1 = ip; o.n = planen; o.add(xs - cs.w * planen, fminf(0.0f, dist)); }
}

// box (body0) against a plane (body1): corners below the plane, at most 4, corner order bit0 = +x, bit1 = +y, bit2 = +z
__device__ __forceinline__ void ext_box_plane(PairOut& o, const float4* __restrict__ quat, const float4* __restrict__ geomA,
                                              const float4* __restrict__ geomB, size_t bi, size_t pi, int ib, int ip, const float4& cb,
                                              const float4& cpl)
{
    const Q4 qp = mkq(__ldg(&quat[pi])), qb = mkq(__ldg(&quat[bi]));
    const V3 planen = rotate(qp, mk3(__ldg(&geomB[pi])));
    const V3 planep = rotate(qp, mk3(__ldg(&geomA[pi]))) + mk3(cpl);
    const float4 dim = __ldg(&geomA[bi]);
    const V3 h = mk3(0.5f * dim.x, 0.5f * dim.y, 0.5f * dim.z);
    o.b0 = ib; o.b1 = ip; o.n = planen;
    for (int k = 0; k < 8 && o.cnt < 4; ++k) {
        const V3 local = mk3((k & 1) ? h.x : -h.x, (k & 2) ? h.y : -h.y, (k & 4) ? h.z : -h.z);
        const V3 c = rotate(qb, local) + mk3(cb);
        const float d = plane_dist(c, planep, planen);
        if (d < 1e-3f) o.add(c, fminf(0.0f, d));
    }
}

// box (body0 = i) against box (body1 = j): separating-axis test over 15 axes, least overlap gives the normal (faces
// preferred); face contact = incident face clipped against the reference face's side planes, corners at or below the
// reference face kept (at most 4: the extremes along the face diagonals); edge contact = midpoint of the closest points
__device__ __forceinline__ void ext_box_box(PairOut& o, const float4* __restrict__ quat, const float4* __restrict__ geomA, size_t ai, size_t bi,
                                            int i0, int i1, const float4& c0, const float4& c1)
{
    const M3 Ra = rotation_matrix(mkq(__ldg(&quat[ai]))), Rb = rotation_matrix(mkq(__ldg(&quat[bi])));
    const float4 da = __ldg(&geomA[ai]), db4 = __ldg(&geomA[bi]);
    V3 ax[2][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ax[0][k] = mk3(Ra.m[0][k], Ra.m[1][k], Ra.m[2][k]); ax[1][k] = mk3(Rb.m[0][k], Rb.m[1][k], Rb.m[2][k]); }
    const float hh[2][3] = { { 0.5f * da.x, 0.5f * da.y, 0.5f * da.z }, { 0.5f * db4.x, 0.5f * db4.y, 0.5f * db4.z } };
    const V3 xc[2] = { mk3(c0), mk3(c1) };
    const V3 t = xc[1] - xc[0];
    float best = 3.0e38f; int best_idx = -1; V3 bestL = mk3(0.f, 0.f, 0.f);
    for (int idx = 0; idx < 15; ++idx) {
        V3 L;
        if (idx < 3) L = ax[0][idx];
        else if (idx < 6) L = ax[1][idx - 3];
        else {
            const V3 c = cross(ax[0][(idx - 6) / 3], ax[1][(idx - 6) % 3]);
            const float l2 = sqnorm(c);
            if (l2 < 1e-6f) continue;
            L = c / sqrtf(l2);
        }
        const float overlap = (proj_radius(ax[0], hh[0], L) + proj_radius(ax[1], hh[1], L)) - fabsf(dot(t, L));
        if (overlap < -1e-3f) return;
        if (idx < 6) { if (overlap < best) { best = overlap; best_idx = idx; bestL = L; } }
        else if (overlap < best - 0.05f * fabsf(best) - 1e-4f) { best = overlap; best_idx = idx; bestL = L; }
    }
    if (best_idx < 0) return;
    o.b0 = i0; o.b1 = i1;
    if (best_idx >= 6) {
        const int ea = (best_idx - 6) / 3, eb = (best_idx - 6) % 3;
        const V3 nab = (dot(t, bestL) > 0.0f) ? bestL : -bestL;
        V3 pa = xc[0], pb = xc[1];
        for (int k = 0; k < 3; ++k) {
            if (k != ea) pa = pa + ((dot(ax[0][k], nab) > 0.0f) ? hh[0][k] : -hh[0][k]) * ax[0][k];
            if (k != eb) pb = pb + ((dot(ax[1][k], nab) < 0.0f) ? hh[1][k] : -hh[1][k]) * ax[1][k];
        }
        const V3 ua = ax[0][ea], ub = ax[1][eb];
        const V3 d = pb - pa;
        const float a = dot(ua, ub), dda = dot(d, ua), ddb = dot(d, ub);
        const float den = 1.0f - a * a;
        const float sa = (dda - a * ddb) / den, sb = (a * dda - ddb) / den;
        const V3 ca = pa + sa * ua, cb = pb + sb * ub;
        o.n = -nab;
        o.add(0.5f * (ca + cb), fminf(0.0f, -best));
        return;
    }
    const int r = best_idx < 3 ? 0 : 1, c = 1 - r, k = best_idx < 3 ? best_idx : best_idx - 3;
    const V3 rc = xc[c] - xc[r];
    const V3 nr = (dot(rc, ax[r][k]) > 0.0f) ? ax[r][k] : -ax[r][k];
    int m = 0; float dmax = -1.0f;
    for (int q = 0; q < 3; ++q) { const float dq = fabsf(dot(ax[c][q], nr)); if (dq > dmax) { dmax = dq; m = q; } }
    const float sgn = (dot(ax[c][m], nr) > 0.0f) ? -1.0f : 1.0f;
    const int u = (m + 1) % 3, w = (m + 2) % 3;
    const V3 fc = xc[c] + (sgn * hh[c][m]) * ax[c][m];
    const V3 eu = hh[c][u] * ax[c][u], ew = hh[c][w] * ax[c][w];
    V3 poly[16], tmp[16];
    int np = 4;
    poly[0] = (fc + eu) + ew; poly[1] = (fc - eu) + ew; poly[2] = (fc - eu) - ew; poly[3] = (fc + eu) - ew;
    const int ru = (k + 1) % 3, rw = (k + 2) % 3;
    for (int side = 0; side < 4 && np > 0; ++side) {
        const int axn = side < 2 ? ru : rw;
        const float sg = (side & 1) ? -1.0f : 1.0f;
        const V3 pn = sg * ax[r][axn];
        const float lim = hh[r][axn];
        int nn = 0;
        for (int e = 0; e < np; ++e) {
            const V3 p0 = poly[e], p1 = poly[(e + 1) % np];
            const float d0 = dot(p0 - xc[r], pn) - lim, d1 = dot(p1 - xc[r], pn) - lim;
            if (d0 <= 0.0f) tmp[nn++] = p0;
            if ((d0 <= 0.0f) != (d1 <= 0.0f)) {
                const float tt = d0 / (d0 - d1);
                tmp[nn++] = p0 + tt * (p1 - p0);
            }
        }
        np = nn;
        for (int e = 0; e < np; ++e) poly[e] = tmp[e];
    }
    V3 pts[16]; float dep[16], ca_[16], cb_[16];
    int nk = 0;
    for (int e = 0; e < np; ++e) {
        const V3 rel = poly[e] - xc[r];
        const float d = dot(rel, nr) - hh[r][k];
        if (d < 1e-3f) { pts[nk] = poly[e]; dep[nk] = d; ca_[nk] = dot(rel, ax[r][ru]); cb_[nk] = dot(rel, ax[r][rw]); ++nk; }
    }
    o.n = (r == 0) ? -nr : nr;
    if (nk <= 4) {
        for (int e = 0; e < nk; ++e) o.add(pts[e], fminf(0.0f, dep[e]));
        return;
    }
    int pick[4]; int npick = 0;
    for (int dgn = 0; dgn < 4; ++dgn) {
        const float sa = (dgn == 0 || dgn == 3) ? 1.0f : -1.0f, sb = (dgn < 2) ? 1.0f : -1.0f;
        int arg = 0; float bestv = -3.0e38f;
        for (int e = 0; e < nk; ++e) { const float val = sa * ca_[e] + sb * cb_[e]; if (val > bestv) { bestv = val; arg = e; } }
        bool dup = false;
        for (int q = 0; q < npick; ++q) dup |= (pick[q] == arg);
        if (!dup) pick[npick++] = arg;
    }
    for (int q = 0; q < npick; ++q) o.add(pts[pick[q]], fminf(0.0f, dep[pick[q]]));
}

// extension dispatch, reached only when the reference's chain matched nothing and the extension is enabled; one
// out-of-line function taking plain pointers, so that the detection kernels' own register budget is untouched
// defer: a mesh pair whose samples a whole warp will walk (test_pair_warp) is only marked, cnt = -1
static __device__ __noinline__ void ext_pair(PairOut& o, const SdfShape* shapes, const float4* __restrict__ quat, const float4* __restrict__ geomA,
                                      const float4* __restrict__ geomB, size_t ii, size_t jj, int i, int j, float4 ci, int ti, float4 cj, int tj,
                                      bool defer)
{
    // bounding spheres first: none of the routines below can produce a contact for bodies further apart than this
    // (their own tolerances are 1e-3), and it keeps far pairs out of the box-box SAT and the deferred mesh walks
    {
        float bi = -1.f, bj = -1.f;
        if (ti == GEOM_SPHERE) bi = ci.w;
        else if (ti == GEOM_BOX) { const float4 d = __ldg(&geomA[ii]); bi = norm(mk3(0.5f * d.x, 0.5f * d.y, 0.5f * d.z)); }
        else if (ti == GEOM_SDF) bi = sdf_shape_of(shapes, geomA, ii).radius;
        else if (ti == GEOM_CYLINDER) bi = cyl_bound(__ldg(&geomA[ii]));
        if (tj == GEOM_SPHERE) bj = cj.w;
        else if (tj == GEOM_BOX) { const float4 d = __ldg(&geomA[jj]); bj = norm(mk3(0.5f * d.x, 0.5f * d.y, 0.5f * d.z)); }
        else if (tj == GEOM_SDF) bj = sdf_shape_of(shapes, geomA, jj).radius;
        else if (tj == GEOM_CYLINDER) bj = cyl_bound(__ldg(&geomA[jj]));
        if (bi >= 0.f && bj >= 0.f) {
            const float lim = bi + bj + 0.01f;
            if (sqnorm(mk3(ci) - mk3(cj)) > lim * lim) return;
        }
    }
    if (defer && ((ti == GEOM_SDF && (tj == GEOM_SDF || tj == GEOM_PLANE || tj == GEOM_BOX)) || (tj == GEOM_SDF && (ti == GEOM_PLANE || ti == GEOM_BOX)))) { o.cnt = -1; return; }
    if (ti == GEOM_SPHERE && tj == GEOM_PLANE)      ext_sphere_plane(o, quat, geomA, geomB, jj, i, j, ci, cj);
    else if (tj == GEOM_SPHERE && ti == GEOM_PLANE) ext_sphere_plane(o, quat, geomA, geomB, ii, j, i, cj, ci);
    else if (ti == GEOM_BOX && tj == GEOM_PLANE)    ext_box_plane(o, quat, geomA, geomB, ii, jj, i, j, ci, cj);
    else if (tj == GEOM_BOX && ti == GEOM_PLANE)    ext_box_plane(o, quat, geomA, geomB, jj, ii, j, i, cj, ci);
    else if (ti == GEOM_BOX && tj == GEOM_BOX)      ext_box_box(o, quat, geomA, ii, jj, i, j, ci, cj);
    else if (ti == GEOM_SDF && tj == GEOM_PLANE)    ext_sdf_plane(o, shapes, quat, geomA, geomB, ii, jj, i, j, ci, cj);
    else if (tj == GEOM_SDF && ti == GEOM_PLANE)    ext_sdf_plane(o, shapes, quat, geomA, geomB, jj, ii, j, i, cj, ci);
    else if (ti == GEOM_SPHERE && tj == GEOM_SDF)   ext_sphere_sdf(o, shapes, quat, geomA, jj, i, j, ci, cj);
    else if (tj == GEOM_SPHERE && ti == GEOM_SDF)   ext_sphere_sdf(o, shapes, quat, geomA, ii, j, i, cj, ci);
    else if (ti == GEOM_SDF && tj == GEOM_SDF)      ext_sdf_sdf(o, shapes, quat, geomA, ii, jj, i, j, ci, cj);
    else if (ti == GEOM_SDF && tj == GEOM_BOX)      ext_sdf_box(o, shapes, quat, geomA, ii, jj, i, j, ci, cj);
    else if (tj == GEOM_SDF && ti == GEOM_BOX)      ext_sdf_box(o, shapes, quat, geomA, jj, ii, j, i, cj, ci);
    else if (ti == GEOM_SPHERE && tj == GEOM_CYLINDER)   ext_sphere_cyl(o, quat, geomA, jj, i, j, ci, cj);
    else if (tj == GEOM_SPHERE && ti == GEOM_CYLINDER)   ext_sphere_cyl(o, quat, geomA, ii, j, i, cj, ci);
    else if (ti == GEOM_CYLINDER && tj == GEOM_BOX)      ext_cyl_box(o, quat, geomA, ii, jj, i, j, ci, cj);
    else if (tj == GEOM_CYLINDER && ti == GEOM_BOX)      ext_cyl_box(o, quat, geomA, jj, ii, j, i, cj, ci);
    else if (ti == GEOM_CYLINDER && tj == GEOM_CYLINDER) ext_cyl_cyl(o, quat, geomA, ii, jj, i, j, ci, cj);
    else if (ti == GEOM_SDF && tj == GEOM_CYLINDER)      ext_sdf_cyl(o, shapes, quat, geomA, ii, jj, i, j, ci, cj);
    else if (tj == GEOM_SDF && ti == GEOM_CYLINDER)      ext_sdf_cyl(o, shapes, quat, geomA, jj, ii, j, i, cj, ci);
}

// the reference's dispatch chain for the ordered pair (i < j), CollisionDetect.cpp:41-73
template <bool EXT, bool DEFER = false>
__device__ __forceinline__ void test_pair(PairOut& o, const View& v, int scene, int i, int j,
                                          const float4& ci, int fi, const float4& cj, int fj)
{
    o.cnt = 0;
    if ((fi & fj) & FLAG_FIXED) return;
    const int ti = fi & FLAG_TYPE_MASK, tj = fj & FLAG_TYPE_MASK;
    if (ti == GEOM_SPHERE && tj == GEOM_SPHERE)          sphere_sphere(o, i, j, ci, cj);
    else if (ti == GEOM_SPHERE && tj == GEOM_BOX) {
        const size_t bj = at(v, j, scene);
        if (sphere_box_may_touch(ci, cj, __ldg(&v.bext[bj]))) sphere_box(o, v.quat, v.geomA, bj, i, j, ci, cj);
    } else if (tj == GEOM_SPHERE && ti == GEOM_BOX) {
        const size_t bi = at(v, i, scene);
        if (sphere_box_may_touch(cj, ci, __ldg(&v.bext[bi]))) sphere_box(o, v.quat, v.geomA, bi, j, i, cj, ci);
    }
    else if (ti == GEOM_CYLINDER && tj == GEOM_PLANE)    cylinder_plane(o, v.quat, v.geomA, v.geomB, at(v, i, scene), at(v, j, scene), i, j, ci, cj);
    else if (tj == GEOM_CYLINDER && ti == GEOM_PLANE)    cylinder_plane(o, v.quat, v.geomA, v.geomB, at(v, j, scene), at(v, i, scene), j, i, cj, ci);
    else if (EXT)                                        ext_pair(o, v.sdf_shapes, v.quat, v.geomA, v.geomB, at(v, i, scene), at(v, j, scene), i, j, ci, ti, cj, tj, DEFER);
}

// a mesh pair (i, j) walked by the whole warp; every lane gets the result
static __device__ __noinline__ void coop_pair(PairOut& o, const SdfShape* shapes, const float4* __restrict__ quat, const float4* __restrict__ geomA,
                                       const float4* __restrict__ geomB, const float4* __restrict__ cproxy, const int* __restrict__ flags,
                                       size_t ii, size_t jj, int i, int j, int lane)
{
    const int ti = __ldg(&flags[ii]) & FLAG_TYPE_MASK, tj = __ldg(&flags[jj]) & FLAG_TYPE_MASK;
    const float4 ci = __ldg(&cproxy[ii]), cj = __ldg(&cproxy[jj]);
    if (ti == GEOM_SDF && tj == GEOM_PLANE)      coop_sdf_plane(o, shapes, quat, geomA, geomB, ii, jj, i, j, ci, cj, lane);
    else if (tj == GEOM_SDF && ti == GEOM_PLANE) coop_sdf_plane(o, shapes, quat, geomA, geomB, jj, ii, j, i, cj, ci, lane);
    else if (ti == GEOM_SDF && tj == GEOM_BOX)   coop_sdf_box(o, shapes, quat, geomA, ii, jj, i, j, ci, cj, lane);
    else if (tj == GEOM_SDF && ti == GEOM_BOX)   coop_sdf_box(o, shapes, quat, geomA, jj, ii, j, i, cj, ci, lane);
    else                                         coop_sdf_sdf(o, shapes, quat, geomA, ii, jj, i, j, ci, cj, lane);
}

// Warp-collective form of test_pair: every lane passes its own pair (i, j), j = -1 for none. Pairs are
// tested lane-parallel as usual, except mesh pairs (extension), which are deferred and then walked one after the other by
// all 32 lanes together.
template <bool EXT>
__device__ __forceinline__ void test_pair_warp(PairOut& o, const View& v, int scene, int i, int j, const float4& ci, int fi,
                                               const float4& cj, int fj, int lane)
{
    o.cnt = 0;
    if (j >= 0) test_pair<EXT, true>(o, v, scene, i, j, ci, fi, cj, fj);
    if (EXT) {
        unsigned def = __ballot_sync(0xffffffffu, o.cnt < 0);
        while (def) {
            const int src = __ffs(def) - 1;
            def &= def - 1;
            const int pi = __shfl_sync(0xffffffffu, i, src), pj = __shfl_sync(0xffffffffu, j, src);
            PairOut r;
            r.cnt = 0;
            coop_pair(r, v.sdf_shapes, v.quat, v.geomA, v.geomB, v.cproxy, v.flags, at(v, pi, scene), at(v, pj, scene), pi, pj, lane);
            if (lane == src) o = r;
        }
    }
}

__device__ __forceinline__ void write_contacts(const View& v, int scene, int base, const PairOut& o)
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < o.cnt) {
            if (base + k < v.NC) {
                const size_t c = atc(v, base + k, scene);
                v.cids[c] = make_int2(o.b0, o.b1);
                v.cp[c] = mk4(o.p[k], o.phi[k]);
                v.cn[c] = mk4(o.n, 0.f);
            } else {
                atomicOr(v.overflow, 1);
            }
        }
    }
}


}  // namespace slrbs

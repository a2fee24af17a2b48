// narrowphase.cuh -- the reference's three narrow-phase routines and its pair dispatch chain
// (CollisionDetect.cpp:41-73,102-274), written against the staged body proxies. Bit-exact with the
// reference arithmetic (vecmath.cuh associations, -fmad=false).
#pragma once
#include "layout.cuh"
#include "vecmath.cuh"

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
__device__ __noinline__ void sphere_box(PairOut& o, const float4* __restrict__ quat, const float4* __restrict__ geomA, size_t bi,
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
        o.n = rotate(qb, dx / dist);
        o.add(rotate(qb, q) + xb, fminf(0.0f, dist - r));
    }
}

// CollisionDetect.cpp:152-274 (body0 = cylinder, body1 = plane)
__device__ __noinline__ void cylinder_plane(PairOut& o, const float4* __restrict__ quat, const float4* __restrict__ geomA,
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

// the reference's dispatch chain for the ordered pair (i < j), CollisionDetect.cpp:41-73
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
}

__device__ __forceinline__ void write_contacts(const View& v, int scene, int base, const PairOut& o)
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < o.cnt) {
            if (base + k < v.NC) {
                const size_t c = at(v, base + k, scene);
                v.cids[c] = make_int2(o.b0, o.b1);
                v.cp[c] = mk4(o.p[k], o.phi[k]);
                v.cn[c] = mk4(o.n, 0.f);
            } else {
                *v.overflow = 1;
            }
        }
    }
}


}  // namespace slrbs

// sdf.cuh -- EXTENSION: triangle meshes as signed-distance grids. The reference has no SDF code;
// these routines follow oracle/slrbs_oracle.c (sdf_sample, ext_sdf_plane, ext_sphere_sdf, ext_sdf_sdf) operation for
// operation (-fmad=false), so contact lists agree with the restatement bit for bit on the same grid.
// A shape = grid (node (i,j,k) at origin + cell*(i,j,k), value grid[(k*ny+j)*nx+i], negative inside) + the mesh vertices
// as sample points, both in the body frame. A pair yields at most its four deepest sample points, deepest first, with
// the normal of the deepest one (the pair record carries a single normal).
#pragma once
#include "layout.cuh"
#include "vecmath.cuh"

namespace slrbs {

__device__ __forceinline__ float lerpf(float a, float b, float t) { return a + t * (b - a); }

// shape of a GEOM_SDF body (geomA.x = shape id); an id outside the table or a slot never filled gives an empty shape
__device__ __forceinline__ SdfShape sdf_shape_of(const SdfShape* shapes, const float4* __restrict__ geomA, size_t bi)
{
    const int id = (int)__ldg(&geomA[bi]).x;
    if (id < 0 || id >= MAX_SDF_SHAPES) { SdfShape e; e.grid = nullptr; e.verts = nullptr; e.nx = e.ny = e.nz = 2; e.nv = 0; e.ox = e.oy = e.oz = 0.f; e.cell = 1.f; e.radius = 0.f; return e; }
    return shapes[id];
}

__device__ __forceinline__ void sdf_sample(const SdfShape& S, const V3& pl, float& phi, V3& grad)
{
    float fx = (pl.x - S.ox) / S.cell, fy = (pl.y - S.oy) / S.cell, fz = (pl.z - S.oz) / S.cell;
    const float cx = fminf(fmaxf(fx, 0.f), (float)(S.nx - 1)), cy = fminf(fmaxf(fy, 0.f), (float)(S.ny - 1)),
                cz = fminf(fmaxf(fz, 0.f), (float)(S.nz - 1));
    const float outside = S.cell * norm(mk3(fx - cx, fy - cy, fz - cz));
    fx = cx; fy = cy; fz = cz;
    int i = (int)fx, j = (int)fy, k = (int)fz;
    if (i > S.nx - 2) i = S.nx - 2;
    if (j > S.ny - 2) j = S.ny - 2;
    if (k > S.nz - 2) k = S.nz - 2;
    const float tx = fx - (float)i, ty = fy - (float)j, tz = fz - (float)k;
    const float* g = S.grid + ((size_t)k * S.ny + j) * S.nx + i;
    const size_t sy = (size_t)S.nx, sz = (size_t)S.nx * S.ny;
    const float c000 = __ldg(g), c100 = __ldg(g + 1), c010 = __ldg(g + sy), c110 = __ldg(g + sy + 1);
    const float c001 = __ldg(g + sz), c101 = __ldg(g + sz + 1), c011 = __ldg(g + sz + sy), c111 = __ldg(g + sz + sy + 1);
    phi = lerpf(lerpf(lerpf(c000, c100, tx), lerpf(c010, c110, tx), ty), lerpf(lerpf(c001, c101, tx), lerpf(c011, c111, tx), ty), tz) + outside;
    V3 gr;
    gr.x = lerpf(lerpf(c100 - c000, c110 - c010, ty), lerpf(c101 - c001, c111 - c011, ty), tz);
    gr.y = lerpf(lerpf(c010 - c000, c110 - c100, tx), lerpf(c011 - c001, c111 - c101, tx), tz);
    gr.z = lerpf(lerpf(c001 - c000, c101 - c100, tx), lerpf(c011 - c010, c111 - c110, tx), ty);
    const float n2 = sqnorm(gr);
    grad = n2 > 0.f ? gr / sqrtf(n2) : mk3(0.f, 1.f, 0.f);
}

struct SdfHits { int cnt; float phi[4]; V3 p[4], n[4]; };
__device__ __forceinline__ void sdf_hit(SdfHits& H, float phi, const V3& p, const V3& n)
{
    int at = H.cnt;
    while (at > 0 && phi < H.phi[at - 1]) --at;
    if (at >= 4) return;
    const int last = H.cnt < 4 ? H.cnt : 3;
    for (int k = last; k > at; --k) { H.phi[k] = H.phi[k - 1]; H.p[k] = H.p[k - 1]; H.n[k] = H.n[k - 1]; }
    H.phi[at] = phi; H.p[at] = p; H.n[at] = n;
    if (H.cnt < 4) ++H.cnt;
}
template <class Out>
__device__ __forceinline__ void sdf_emit(Out& o, int b0, int b1, const SdfHits& H)
{
    if (H.cnt == 0) return;
    o.b0 = b0; o.b1 = b1; o.n = H.n[0];
    for (int k = 0; k < H.cnt; ++k) o.add(H.p[k], fminf(0.0f, H.phi[k]));
}
__device__ __forceinline__ V3 sdf_vertex(const SdfShape& S, int k) { return mk3(__ldg(S.verts + 3 * k), __ldg(S.verts + 3 * k + 1), __ldg(S.verts + 3 * k + 2)); }

// mesh (body0): vertices under the plane (body1)
template <class Out>
__device__ void ext_sdf_plane(Out& o, const SdfShape* shapes, const float4* __restrict__ quat, const float4* __restrict__ geomA,
                              const float4* __restrict__ geomB, size_t mi, size_t pi, int im, int ip, const float4& cm, const float4& cpl)
{
    const SdfShape S = sdf_shape_of(shapes, geomA, mi);
    if (!S.grid) return;
    const Q4 qp = mkq(__ldg(&quat[pi]));
    const V3 planen = rotate(qp, mk3(__ldg(&geomB[pi])));
    const V3 planep = rotate(qp, mk3(__ldg(&geomA[pi]))) + mk3(cpl);
    const V3 xm = mk3(cm);
    if (dot(xm - planep, planen) > S.radius + 1e-3f) return;
    const Q4 qm = mkq(__ldg(&quat[mi]));
    SdfHits H; H.cnt = 0;
    for (int k = 0; k < S.nv; ++k) {
        const V3 pw = rotate(qm, sdf_vertex(S, k)) + xm;
        const float d = dot(pw - planep, planen);
        if (d < 1e-3f) sdf_hit(H, d, pw, planen);
    }
    sdf_emit(o, im, ip, H);
}

// sphere (body0): its centre in the grid of the mesh (body1)
template <class Out>
__device__ void ext_sphere_sdf(Out& o, const SdfShape* shapes, const float4* __restrict__ quat, const float4* __restrict__ geomA,
                               size_t mi, int is, int im, const float4& cs, const float4& cm)
{
    const SdfShape S = sdf_shape_of(shapes, geomA, mi);
    if (!S.grid) return;
    const V3 xs = mk3(cs), rel = xs - mk3(cm);
    if (norm(rel) > S.radius + cs.w + 1e-3f) return;
    const Q4 qm = mkq(__ldg(&quat[mi]));
    float val; V3 g;
    sdf_sample(S, rotate(qinverse(qm), rel), val, g);
    const float d = val - cs.w;
    if (d < 1e-3f) {
        const V3 n = rotate(qm, g);
        o.b0 = is; o.b1 = im; o.n = n;
        o.add(xs - cs.w * n, fminf(0.0f, d));
    }
}

__device__ __forceinline__ void sdf_verts_in(const SdfShape& Sf, const Q4& qf, const V3& xf, const SdfShape& Si, const Q4& qi, const V3& xi,
                                             float flip, SdfHits& H)
{
    const Q4 qinv = qinverse(qi);
    for (int k = 0; k < Sf.nv; ++k) {
        const V3 pw = rotate(qf, sdf_vertex(Sf, k)) + xf;
        float val; V3 g;
        sdf_sample(Si, rotate(qinv, pw - xi), val, g);
        if (val < 1e-3f) sdf_hit(H, val, pw, flip * rotate(qi, g));
    }
}

// mesh i (body0) against mesh j (body1): body0's vertices in body1's grid, then body1's in body0's
template <class Out>
__device__ void ext_sdf_sdf(Out& o, const SdfShape* shapes, const float4* __restrict__ quat, const float4* __restrict__ geomA,
                            size_t ai, size_t bi, int ia, int ib, const float4& ca, const float4& cb)
{
    const SdfShape Sa = sdf_shape_of(shapes, geomA, ai), Sb = sdf_shape_of(shapes, geomA, bi);
    if (!Sa.grid || !Sb.grid) return;
    const V3 xa = mk3(ca), xb = mk3(cb);
    if (norm(xa - xb) > Sa.radius + Sb.radius + 1e-3f) return;
    const Q4 qa = mkq(__ldg(&quat[ai])), qb = mkq(__ldg(&quat[bi]));
    SdfHits H; H.cnt = 0;
    sdf_verts_in(Sa, qa, xa, Sb, qb, xb, 1.0f, H);
    sdf_verts_in(Sb, qb, xb, Sa, qa, xa, -1.0f, H);
    sdf_emit(o, ia, ib, H);
}

// signed distance and outward unit normal of a box of half extents h at the box-frame point c
__device__ __forceinline__ float box_sdf(const V3& c, const V3& h, V3& n)
{
    const V3 q = mk3(fminf(fmaxf(c.x, -h.x), h.x), fminf(fmaxf(c.y, -h.y), h.y), fminf(fmaxf(c.z, -h.z), h.z));
    const V3 dx = c - q;
    const float dist = norm(dx);
    if (dist > 0.f) { n = dx / dist; return dist; }
    const float ex = h.x - fabsf(c.x), ey = h.y - fabsf(c.y), ez = h.z - fabsf(c.z);
    if (ex <= ey && ex <= ez) { n = mk3(c.x < 0.f ? -1.f : 1.f, 0.f, 0.f); return -ex; }
    if (ey <= ez)             { n = mk3(0.f, c.y < 0.f ? -1.f : 1.f, 0.f); return -ey; }
    n = mk3(0.f, 0.f, c.z < 0.f ? -1.f : 1.f); return -ez;
}
__device__ __forceinline__ V3 box_corner(const V3& h, int k) { return mk3((k & 1) ? h.x : -h.x, (k & 2) ? h.y : -h.y, (k & 4) ? h.z : -h.z); }

// mesh (body0) against box (body1): the mesh vertices against the box, then the box corners in the mesh's grid
template <class Out>
__device__ void ext_sdf_box(Out& o, const SdfShape* shapes, const float4* __restrict__ quat, const float4* __restrict__ geomA,
                            size_t mi, size_t bi, int im, int ib, const float4& cm, const float4& cb)
{
    const SdfShape S = sdf_shape_of(shapes, geomA, mi);
    if (!S.grid) return;
    const float4 dim = __ldg(&geomA[bi]);
    const V3 h = mk3(0.5f * dim.x, 0.5f * dim.y, 0.5f * dim.z);
    const V3 xm = mk3(cm), xb = mk3(cb);
    if (norm(xm - xb) > S.radius + norm(h) + 1e-3f) return;
    const Q4 qm = mkq(__ldg(&quat[mi])), qb = mkq(__ldg(&quat[bi]));
    const Q4 qbinv = qinverse(qb), qminv = qinverse(qm);
    SdfHits H; H.cnt = 0;
    for (int k = 0; k < S.nv; ++k) {
        const V3 pw = rotate(qm, sdf_vertex(S, k)) + xm;
        V3 nl;
        const float d = box_sdf(rotate(qbinv, pw - xb), h, nl);
        if (d < 1e-3f) sdf_hit(H, d, pw, rotate(qb, nl));
    }
    for (int k = 0; k < 8; ++k) {
        const V3 pw = rotate(qb, box_corner(h, k)) + xb;
        float val; V3 g;
        sdf_sample(S, rotate(qminv, pw - xm), val, g);
        if (val < 1e-3f) sdf_hit(H, val, pw, -rotate(qm, g));
    }
    sdf_emit(o, im, ib, H);
}

// ---- EXTENSION: cylinder pairs the reference ignores (sphere-cylinder, cylinder-box, cylinder-cylinder, mesh-cylinder;
// CollisionDetect.cpp:45-73 only knows cylinder-plane). Follows oracle/slrbs_oracle.c (cyl_sdf, cyl_sample, ext_sphere_cyl,
// ext_cyl_cyl, ext_cyl_box, ext_sdf_cyl) operation for operation: a finite cylinder (axis = local y) is an analytic signed
// distance plus CYL_SAMPLES surface points (three rings of 16 and the cap centres); sample points of one shape are tested
// in the distance field of the other, the four deepest kept.
enum { CYL_RING = 16, CYL_SAMPLES = 50 };
__device__ __forceinline__ V3 cyl_sample(int k, float h2, float r)
{
    // cos / sin of 2 pi a / 16 as the same literals the restatement uses (libm and CUDA need not round alike)
    const float C[CYL_RING] = { 1.0f, 0.9238795f, 0.70710677f, 0.38268343f, 0.0f, -0.38268343f, -0.70710677f, -0.9238795f,
                                -1.0f, -0.9238795f, -0.70710677f, -0.38268343f, 0.0f, 0.38268343f, 0.70710677f, 0.9238795f };
    if (k >= 3 * CYL_RING) return mk3(0.f, k == 3 * CYL_RING ? -h2 : h2, 0.f);
    const int ring = k / CYL_RING, a = k - ring * CYL_RING;
    return mk3(r * C[a], ring == 0 ? -h2 : (ring == 1 ? 0.f : h2), r * C[(a + 12) & 15]);       // sin a = cos (a - 4) = C[a + 12]
}
__device__ __forceinline__ float cyl_sdf(const V3& c, float h2, float r, V3& n)
{
    const float rho = sqrtf(c.x * c.x + c.z * c.z);
    const V3 radial = rho > 0.f ? mk3(c.x / rho, 0.f, c.z / rho) : mk3(1.f, 0.f, 0.f);
    const float dr = rho - r, dy = fabsf(c.y) - h2, sy = c.y < 0.f ? -1.f : 1.f;
    if (dr <= 0.f && dy <= 0.f) {
        if (dr > dy) { n = radial; return dr; }
        n = mk3(0.f, sy, 0.f); return dy;
    }
    if (dr > 0.f && dy > 0.f) {
        const float d = sqrtf(dr * dr + dy * dy);
        n = mk3((dr * radial.x) / d, (dy * sy) / d, (dr * radial.z) / d);
        return d;
    }
    if (dr > 0.f) { n = radial; return dr; }
    n = mk3(0.f, sy, 0.f); return dy;
}
__device__ __forceinline__ float cyl_bound(const float4& hr) { const float h2 = 0.5f * hr.x; return sqrtf(h2 * h2 + hr.y * hr.y); }

// sphere (body0) against cylinder (body1)
template <class Out>
__device__ void ext_sphere_cyl(Out& o, const float4* __restrict__ quat, const float4* __restrict__ geomA, size_t ci_, int is, int ic,
                               const float4& cs, const float4& cc)
{
    const float4 hr = __ldg(&geomA[ci_]);
    const V3 xs = mk3(cs), rel = xs - mk3(cc);
    if (norm(rel) > cs.w + cyl_bound(hr) + 1e-3f) return;
    const Q4 qc = mkq(__ldg(&quat[ci_]));
    V3 nl;
    const float d = cyl_sdf(rotate(qinverse(qc), rel), 0.5f * hr.x, hr.y, nl) - cs.w;
    if (d < 1e-3f) {
        const V3 n = rotate(qc, nl);
        o.b0 = is; o.b1 = ic; o.n = n;
        o.add(xs - cs.w * n, fminf(0.0f, d));
    }
}

__device__ __forceinline__ void cyl_samples_in_cyl(const float4& hf, const Q4& qf, const V3& xf, const float4& hi, const Q4& qi, const V3& xi,
                                                   float flip, SdfHits& H)
{
    const Q4 qinv = qinverse(qi);
    for (int k = 0; k < CYL_SAMPLES; ++k) {
        const V3 pw = rotate(qf, cyl_sample(k, 0.5f * hf.x, hf.y)) + xf;
        V3 nl;
        const float d = cyl_sdf(rotate(qinv, pw - xi), 0.5f * hi.x, hi.y, nl);
        if (d < 1e-3f) sdf_hit(H, d, pw, flip * rotate(qi, nl));
    }
}

// cylinder (body0) against cylinder (body1)
template <class Out>
__device__ void ext_cyl_cyl(Out& o, const float4* __restrict__ quat, const float4* __restrict__ geomA, size_t ai, size_t bi, int ia, int ib,
                            const float4& ca, const float4& cb)
{
    const float4 ha = __ldg(&geomA[ai]), hb = __ldg(&geomA[bi]);
    const V3 xa = mk3(ca), xb = mk3(cb);
    if (norm(xa - xb) > cyl_bound(ha) + cyl_bound(hb) + 1e-3f) return;
    const Q4 qa = mkq(__ldg(&quat[ai])), qb = mkq(__ldg(&quat[bi]));
    SdfHits H; H.cnt = 0;
    cyl_samples_in_cyl(ha, qa, xa, hb, qb, xb, 1.0f, H);
    cyl_samples_in_cyl(hb, qb, xb, ha, qa, xa, -1.0f, H);
    sdf_emit(o, ia, ib, H);
}

// cylinder (body0) against box (body1)
template <class Out>
__device__ void ext_cyl_box(Out& o, const float4* __restrict__ quat, const float4* __restrict__ geomA, size_t ci_, size_t bi, int ic, int ib,
                            const float4& cc, const float4& cb)
{
    const float4 hr = __ldg(&geomA[ci_]), dim = __ldg(&geomA[bi]);
    const V3 h = mk3(0.5f * dim.x, 0.5f * dim.y, 0.5f * dim.z);
    const V3 xc = mk3(cc), xb = mk3(cb);
    if (norm(xc - xb) > cyl_bound(hr) + norm(h) + 1e-3f) return;
    const Q4 qc = mkq(__ldg(&quat[ci_])), qb = mkq(__ldg(&quat[bi]));
    const Q4 qbinv = qinverse(qb), qcinv = qinverse(qc);
    SdfHits H; H.cnt = 0;
    for (int k = 0; k < CYL_SAMPLES; ++k) {
        const V3 pw = rotate(qc, cyl_sample(k, 0.5f * hr.x, hr.y)) + xc;
        V3 nl;
        const float d = box_sdf(rotate(qbinv, pw - xb), h, nl);
        if (d < 1e-3f) sdf_hit(H, d, pw, rotate(qb, nl));
    }
    for (int k = 0; k < 8; ++k) {
        const V3 pw = rotate(qb, box_corner(h, k)) + xb;
        V3 nl;
        const float d = cyl_sdf(rotate(qcinv, pw - xc), 0.5f * hr.x, hr.y, nl);
        if (d < 1e-3f) sdf_hit(H, d, pw, -rotate(qc, nl));
    }
    sdf_emit(o, ic, ib, H);
}

// mesh (body0) against cylinder (body1), one thread (the warp-cooperative walk is kept for the mesh pairs with many samples on
// BOTH sides; a cylinder adds 50)
template <class Out>
__device__ void ext_sdf_cyl(Out& o, const SdfShape* shapes, const float4* __restrict__ quat, const float4* __restrict__ geomA,
                            size_t mi, size_t ci_, int im, int ic, const float4& cm, const float4& cc)
{
    const SdfShape S = sdf_shape_of(shapes, geomA, mi);
    if (!S.grid) return;
    const float4 hr = __ldg(&geomA[ci_]);
    const V3 xm = mk3(cm), xc = mk3(cc);
    if (norm(xm - xc) > S.radius + cyl_bound(hr) + 1e-3f) return;
    const Q4 qm = mkq(__ldg(&quat[mi])), qc = mkq(__ldg(&quat[ci_]));
    const Q4 qcinv = qinverse(qc), qminv = qinverse(qm);
    SdfHits H; H.cnt = 0;
    for (int k = 0; k < S.nv; ++k) {
        const V3 pw = rotate(qm, sdf_vertex(S, k)) + xm;
        V3 nl;
        const float d = cyl_sdf(rotate(qcinv, pw - xc), 0.5f * hr.x, hr.y, nl);
        if (d < 1e-3f) sdf_hit(H, d, pw, rotate(qc, nl));
    }
    for (int k = 0; k < CYL_SAMPLES; ++k) {
        const V3 pw = rotate(qc, cyl_sample(k, 0.5f * hr.x, hr.y)) + xc;
        float val; V3 g;
        sdf_sample(S, rotate(qminv, pw - xm), val, g);
        if (val < 1e-3f) sdf_hit(H, val, pw, -rotate(qm, g));
    }
    sdf_emit(o, im, ic, H);
}

// ---- warp-cooperative versions: all 32 lanes call with the same arguments, the samples are dealt to the lanes, and the
// four deepest of the pair are merged in the order (phi, sample index) -- the order in which the one-thread routines
// above keep them (stable insertion), so the result is the same bit for bit. Every lane returns the full result.
struct SdfHitsIdx { int cnt; float phi[4]; int idx[4]; V3 p[4], n[4]; };
__device__ __forceinline__ void sdf_hit_idx(SdfHitsIdx& H, float phi, int idx, const V3& p, const V3& n)
{
    int at = H.cnt;
    while (at > 0 && phi < H.phi[at - 1]) --at;
    if (at >= 4) return;
    const int last = H.cnt < 4 ? H.cnt : 3;
    for (int k = last; k > at; --k) { H.phi[k] = H.phi[k - 1]; H.idx[k] = H.idx[k - 1]; H.p[k] = H.p[k - 1]; H.n[k] = H.n[k - 1]; }
    H.phi[at] = phi; H.idx[at] = idx; H.p[at] = p; H.n[at] = n;
    if (H.cnt < 4) ++H.cnt;
}
template <class Out>
__device__ __forceinline__ void sdf_merge_emit(Out& o, int b0, int b1, const SdfHitsIdx& H, int lane)
{
    const unsigned full = 0xffffffffu;
    int head = 0;
    for (int r = 0; r < 4; ++r) {
        const bool have = head < H.cnt;
        const float ph = have ? H.phi[head] : 3.0e38f;
        const int id = have ? H.idx[head] : 0x7fffffff;
        float mn = ph;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mn = fminf(mn, __shfl_xor_sync(full, mn, d));
        if (mn >= 3.0e38f) break;                                   // warp-uniform
        int mi = (ph == mn) ? id : 0x7fffffff;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mi = min(mi, __shfl_xor_sync(full, mi, d));
        const int win = __ffs(__ballot_sync(full, have && ph == mn && id == mi)) - 1;
        V3 p = mk3(0.f, 0.f, 0.f), n = p;
        if (have) { p = H.p[head]; n = H.n[head]; }
        p = mk3(__shfl_sync(full, p.x, win), __shfl_sync(full, p.y, win), __shfl_sync(full, p.z, win));
        n = mk3(__shfl_sync(full, n.x, win), __shfl_sync(full, n.y, win), __shfl_sync(full, n.z, win));
        if (r == 0) { o.b0 = b0; o.b1 = b1; o.n = n; }
        o.add(p, fminf(0.0f, mn));
        if (lane == win) ++head;
    }
}

template <class Out>
__device__ void coop_sdf_plane(Out& o, const SdfShape* shapes, const float4* __restrict__ quat, const float4* __restrict__ geomA,
                               const float4* __restrict__ geomB, size_t mi, size_t pi, int im, int ip, const float4& cm, const float4& cpl, int lane)
{
    const SdfShape S = sdf_shape_of(shapes, geomA, mi);
    if (!S.grid) return;
    const Q4 qp = mkq(__ldg(&quat[pi]));
    const V3 planen = rotate(qp, mk3(__ldg(&geomB[pi])));
    const V3 planep = rotate(qp, mk3(__ldg(&geomA[pi]))) + mk3(cpl);
    const V3 xm = mk3(cm);
    if (dot(xm - planep, planen) > S.radius + 1e-3f) return;
    const Q4 qm = mkq(__ldg(&quat[mi]));
    SdfHitsIdx H; H.cnt = 0;
    for (int k = lane; k < S.nv; k += 32) {
        const V3 pw = rotate(qm, sdf_vertex(S, k)) + xm;
        const float d = dot(pw - planep, planen);
        if (d < 1e-3f) sdf_hit_idx(H, d, k, pw, planen);
    }
    sdf_merge_emit(o, im, ip, H, lane);
}

__device__ __forceinline__ void coop_verts_in(const SdfShape& Sf, const Q4& qf, const V3& xf, const SdfShape& Si, const Q4& qi, const V3& xi,
                                              float flip, int idx0, SdfHitsIdx& H, int lane)
{
    const Q4 qinv = qinverse(qi);
    for (int k = lane; k < Sf.nv; k += 32) {
        const V3 pw = rotate(qf, sdf_vertex(Sf, k)) + xf;
        float val; V3 g;
        sdf_sample(Si, rotate(qinv, pw - xi), val, g);
        if (val < 1e-3f) sdf_hit_idx(H, val, idx0 + k, pw, flip * rotate(qi, g));
    }
}

template <class Out>
__device__ void coop_sdf_sdf(Out& o, const SdfShape* shapes, const float4* __restrict__ quat, const float4* __restrict__ geomA,
                             size_t ai, size_t bi, int ia, int ib, const float4& ca, const float4& cb, int lane)
{
    const SdfShape Sa = sdf_shape_of(shapes, geomA, ai), Sb = sdf_shape_of(shapes, geomA, bi);
    if (!Sa.grid || !Sb.grid) return;
    const V3 xa = mk3(ca), xb = mk3(cb);
    if (norm(xa - xb) > Sa.radius + Sb.radius + 1e-3f) return;
    const Q4 qa = mkq(__ldg(&quat[ai])), qb = mkq(__ldg(&quat[bi]));
    SdfHitsIdx H; H.cnt = 0;
    coop_verts_in(Sa, qa, xa, Sb, qb, xb, 1.0f, 0, H, lane);
    coop_verts_in(Sb, qb, xb, Sa, qa, xa, -1.0f, Sa.nv, H, lane);
    sdf_merge_emit(o, ia, ib, H, lane);
}

template <class Out>
__device__ void coop_sdf_box(Out& o, const SdfShape* shapes, const float4* __restrict__ quat, const float4* __restrict__ geomA,
                             size_t mi, size_t bi, int im, int ib, const float4& cm, const float4& cb, int lane)
{
    const SdfShape S = sdf_shape_of(shapes, geomA, mi);
    if (!S.grid) return;
    const float4 dim = __ldg(&geomA[bi]);
    const V3 h = mk3(0.5f * dim.x, 0.5f * dim.y, 0.5f * dim.z);
    const V3 xm = mk3(cm), xb = mk3(cb);
    if (norm(xm - xb) > S.radius + norm(h) + 1e-3f) return;
    const Q4 qm = mkq(__ldg(&quat[mi])), qb = mkq(__ldg(&quat[bi]));
    const Q4 qbinv = qinverse(qb), qminv = qinverse(qm);
    SdfHitsIdx H; H.cnt = 0;
    for (int k = lane; k < S.nv; k += 32) {
        const V3 pw = rotate(qm, sdf_vertex(S, k)) + xm;
        V3 nl;
        const float d = box_sdf(rotate(qbinv, pw - xb), h, nl);
        if (d < 1e-3f) sdf_hit_idx(H, d, k, pw, rotate(qb, nl));
    }
    if (lane < 8) {
        const V3 pw = rotate(qb, box_corner(h, lane)) + xb;
        float val; V3 g;
        sdf_sample(S, rotate(qminv, pw - xm), val, g);
        if (val < 1e-3f) sdf_hit_idx(H, val, S.nv + lane, pw, -rotate(qm, g));
    }
    sdf_merge_emit(o, im, ib, H, lane);
}

}  // namespace slrbs

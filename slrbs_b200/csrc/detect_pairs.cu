// detect_pairs.cu -- pair-parallel collision detection for scenes with the extension pairs (slrbs_set_extensions) and
// >= 128 bodies: boxes, meshes and planes mixed with spheres. The row-parallel kernels of detect.cu give a scene one CTA
// and bin only spheres, so in such a scene every non-sphere body was walked against every later body on one SM.
//   k_pair_candidates   thread per body i, all CTAs of the GPU: partners j > i whose bounding spheres come within 1 cm
//                       (bodies without a finite bound -- planes, cylinders -- pair with everything) are appended to the
//                       scene's candidate list in any order
//   k_pair_narrow       lane per candidate: the same test_pair dispatch as the other kernels; mesh pairs are walked by
//                       the whole warp (test_pair_warp); a pair that touches appends its result to the scene's hit list
//   k_pair_order        CTA per scene: the hit list is sorted by (i, j) (bitonic, shared memory), a block scan
//                       of their contact counts gives the offsets, and the contacts are written in the reference's
//                       lexicographic order (CollisionDetect.cpp:32-34) -- the same list, bit for bit, as the other kernels
#include "kernels.cuh"
#include "narrowphase.cuh"

namespace slrbs {

enum { PAIR_REC = 6 };            // float4 per touching pair: (b0 b1 cnt key) | n phi0 | p0 phi1 | p1 phi2 | p2 phi3 | p3 -

// half extents of a world-axis-aligned box around the body (x < 0: no finite bound) | bounding radius (0: use the box only)
__device__ __forceinline__ float4 body_bound(const View& v, size_t bi, int type, float sphere_r)
{
    if (type == GEOM_SPHERE) return make_float4(sphere_r, sphere_r, sphere_r, sphere_r);
    if (type == GEOM_BOX) { const float4 e = __ldg(&v.bext[bi]); return make_float4(e.x, e.y, e.z, 0.f); }      // k_prepare: |R| dim/2
    if (type == GEOM_SDF) { const float r = sdf_shape_of(v.sdf_shapes, v.geomA, bi).radius; return make_float4(r, r, r, r); }
    return make_float4(-1.f, -1.f, -1.f, 0.f);
}

__global__ void __launch_bounds__(128) k_pair_candidates(View v, int jchunk)
{
    const int scene = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = v.nbodies[scene];
    const bool live = i < n;
    float4 ci = make_float4(0.f, 0.f, 0.f, 0.f), bi = make_float4(-1.f, -1.f, -1.f, 0.f);
    int fi = 0;
    if (live) {
        const size_t ii = at(v, i, scene);
        ci = __ldg(&v.cproxy[ii]); fi = __ldg(&v.flags[ii]);
        bi = body_bound(v, ii, fi & FLAG_TYPE_MASK, ci.w);
    }
    // every lane visits the same j (broadcast loads); a pair is taken by the lane with i < j. The bounds are conservative
    // (1 cm of slack against the routines' own 1e-3 tolerances), so no pair that can touch is lost.
    // blockIdx.z deals the partner range to several CTAs when the batch alone would not fill the GPU
    const int j0 = max((int)(blockIdx.x * blockDim.x + 1), (int)blockIdx.z * jchunk);     // no lane of this CTA has a partner below this
    const int j1 = min(n, ((int)blockIdx.z + 1) * jchunk);
    for (int j = j0; j < j1; ++j) {
        const size_t jj = at(v, j, scene);
        const float4 cj = __ldg(&v.cproxy[jj]);
        const int fj = __ldg(&v.flags[jj]);
        const float4 bj = body_bound(v, jj, fj & FLAG_TYPE_MASK, cj.w);
        if (!live || j <= i || ((fi & fj) & FLAG_FIXED)) continue;
        if (!(bi.x >= 0.f || bi.x < 0.f) || !(bj.x >= 0.f || bj.x < 0.f)) continue;      // NaN bound (a box with a NaN pose): see below
        if (bi.x >= 0.f && bj.x >= 0.f) {
            const V3 d = mk3(ci) - mk3(cj);
            // written so that a NaN position (the reference's sphere-centre-inside-a-box quirk) fails the test: every
            // routine yields nothing for such a body, and it must not flood the candidate list
            if (!(fabsf(d.x) <= bi.x + bj.x + 0.01f && fabsf(d.y) <= bi.y + bj.y + 0.01f && fabsf(d.z) <= bi.z + bj.z + 0.01f)) continue;
            if (bi.w > 0.f && bj.w > 0.f) { const float lim = bi.w + bj.w + 0.01f; if (!(sqnorm(d) <= lim * lim)) continue; }
        }
        const int slot = atomicAdd(&v.pc_count[scene], 1);
        if (slot < v.pc_cap) v.pc_pairs[(size_t)scene * v.pc_cap + slot] = make_int2(i, j);
    }
}

template <bool EXT>
__global__ void __launch_bounds__(128) k_pair_narrow(View v)
{
    const int scene = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const int total = min(v.pc_count[scene], v.pc_cap);
    if ((k & ~31) >= total) return;                               // whole warp out of range
    const bool live = k < total;
    int i = 0, j = -1;
    float4 ci = make_float4(0.f, 0.f, 0.f, 0.f), cj = ci;
    int fi = 0, fj = 0;
    if (live) {
        const int2 pr = v.pc_pairs[(size_t)scene * v.pc_cap + k];
        i = pr.x; j = pr.y;
        const size_t ii = at(v, i, scene), jj = at(v, j, scene);
        ci = __ldg(&v.cproxy[ii]); fi = __ldg(&v.flags[ii]);
        cj = __ldg(&v.cproxy[jj]); fj = __ldg(&v.flags[jj]);
    }
    PairOut o;
    test_pair_warp<EXT>(o, v, scene, i, j, ci, fi, cj, fj, lane);
    if (!live || o.cnt <= 0) return;
    // pairs that touch claim a slot of the scene's hit list (any order; k_pair_order sorts them by (i, j))
    const int slot = atomicAdd(&v.pc_hits[scene], 1);
    if (slot >= v.NC) return;                                     // more touching pairs than contact slots: flagged in k_pair_order
    float4* rec = v.pc_out + ((size_t)scene * v.NC + slot) * PAIR_REC;
    rec[0] = make_float4(__int_as_float(o.b0), __int_as_float(o.b1), __int_as_float(o.cnt), __int_as_float((i << 16) | j));
    rec[1] = mk4(o.n, o.phi[0]);
    rec[2] = mk4(o.p[0], o.cnt > 1 ? o.phi[1] : 0.f);
    if (o.cnt > 1) rec[3] = mk4(o.p[1], o.cnt > 2 ? o.phi[2] : 0.f);
    if (o.cnt > 2) rec[4] = mk4(o.p[2], o.cnt > 3 ? o.phi[3] : 0.f);
    if (o.cnt > 3) rec[5] = mk4(o.p[3], 0.f);
}

enum { ORDER_THREADS = 512 };

// WIDE = false: keys and record indices sorted together in shared memory (8 B per touching pair, <= 16384 pairs).
// WIDE = true : scenes with up to 32768 touching pairs (one large scene of ~10^4 bodies): only the keys fit (4 B each); after
//               the sort every record finds its rank by binary search (keys are unique) and the sorted order of the
//               record indices goes to a global scratch list (the scene's cslot row, free until the rows are ordered).
template <bool WIDE>
__global__ void __launch_bounds__(ORDER_THREADS) k_pair_order(View v, int sort_cap)
{
    extern __shared__ unsigned sm_u[];
    unsigned* key = sm_u;                      // [sort_cap] (i << 16) | j
    unsigned* val = WIDE ? reinterpret_cast<unsigned*>(v.cslot + (size_t)blockIdx.x * v.NC) : sm_u + sort_cap;   // [sort_cap] record index
    __shared__ int s_warp[ORDER_THREADS / 32];
    const int scene = blockIdx.x, tid = threadIdx.x, T = ORDER_THREADS, lane = tid & 31, warp = tid >> 5;
    if (tid == 0 && v.pc_count[scene] > v.pc_cap) atomicOr(v.overflow, 4);
    const float4* out = v.pc_out + (size_t)scene * v.NC * PAIR_REC;
    int m = v.pc_hits[scene];
    const int lim = sort_cap < v.NC ? sort_cap : v.NC;
    if (m > lim) { if (tid == 0) atomicOr(v.overflow, 4); m = lim; }
    for (int k = tid; k < m; k += T) { key[k] = (unsigned)__float_as_int(out[(size_t)k * PAIR_REC].w); if (!WIDE) val[k] = (unsigned)k; }
    int P = 1;
    while (P < m) P <<= 1;
    for (int k = m + tid; k < P; k += T) { key[k] = 0xffffffffu; if (!WIDE) val[k] = 0u; }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (P >> 1); t += T) {
                const int lo = ((t / stride) * (stride << 1)) + (t % stride), hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned a = key[lo], b = key[hi];
                if ((a > b) == up) {
                    key[lo] = b; key[hi] = a;
                    if (!WIDE) { const unsigned x = val[lo]; val[lo] = val[hi]; val[hi] = x; }
                }
            }
            __syncthreads();
        }
    }
    if (WIDE) {
        for (int k = tid; k < m; k += T) {
            const unsigned mine = (unsigned)__float_as_int(out[(size_t)k * PAIR_REC].w);
            int lo = 0, hi = m - 1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (key[mid] < mine) lo = mid + 1; else hi = mid; }
            val[lo] = (unsigned)k;
        }
        __syncthreads();
    }
    // exclusive scan of the contact counts in sorted order
    const int per = (m + T - 1) / T;
    const int r0 = min(tid * per, m), r1 = min(r0 + per, m);
    int sum = 0;
    for (int r = r0; r < r1; ++r) sum += __float_as_int(out[(size_t)val[r] * PAIR_REC].z);
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    int run = base + incl - sum;
    for (int r = r0; r < r1; ++r) {
        const float4* rec = out + (size_t)val[r] * PAIR_REC;
        const float4 h = rec[0], nn = rec[1];
        PairOut o;
        o.b0 = __float_as_int(h.x); o.b1 = __float_as_int(h.y); o.cnt = __float_as_int(h.z);
        o.n = mk3(nn);
        o.phi[0] = nn.w;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (q < o.cnt) { const float4 pq = rec[2 + q]; o.p[q] = mk3(pq); if (q + 1 < 4) o.phi[q + 1] = pq.w; }
        }
        write_contacts(v, scene, run, o);
        run += o.cnt;
    }
    if (tid == T - 1) v.ncontacts[scene] = run < v.NC ? run : v.NC;
}

// touching pairs <= contacts; 16384 x 8 B = 128 KB of keys + indices, or (contexts with cslot, i.e. level / colour mode)
// 32768 x 4 B of keys with the indices in global memory
static bool pair_sort_wide(const View& v) { return v.NC > 16384 && v.cslot != nullptr; }
int pair_sort_capacity(const View& v)
{
    const int top = pair_sort_wide(v) ? 32768 : 16384;
    int cap = 1024;
    while (cap < v.NC && cap < top) cap <<= 1;
    return cap;
}

void init_pair_kernels()
{
    cudaFuncSetAttribute(k_pair_order<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
    cudaFuncSetAttribute(k_pair_order<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4);
}

void launch_detect_pairs(const View& v, cudaStream_t s)
{
    const dim3 gb((unsigned)((v.NB + 127) / 128), (unsigned)v.B), gc((unsigned)((v.pc_cap + 127) / 128), (unsigned)v.B);
    // about eight CTAs per SM in total
    const int rows = (v.NB + 127) / 128;
    int split = (int)((148 * 8 + (long)rows * v.B - 1) / ((long)rows * v.B));
    split = split < 1 ? 1 : (split > 32 ? 32 : split);
    const int jchunk = (v.NB + split - 1) / split;
    k_pair_candidates<<<dim3(gb.x, gb.y, (unsigned)((v.NB + jchunk - 1) / jchunk)), 128, 0, s>>>(v, jchunk);
    if (v.ext_pairs) k_pair_narrow<true><<<gc, 128, 0, s>>>(v);
    else k_pair_narrow<false><<<gc, 128, 0, s>>>(v);
    const int sc = pair_sort_capacity(v);
    if (pair_sort_wide(v)) k_pair_order<true><<<(unsigned)v.B, ORDER_THREADS, (size_t)sc * 4, s>>>(v, sc);
    else k_pair_order<false><<<(unsigned)v.B, ORDER_THREADS, (size_t)sc * 8, s>>>(v, sc);
}

}  // namespace slrbs

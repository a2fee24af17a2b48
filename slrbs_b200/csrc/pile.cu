// pile.cu -- device-side re-partition of ONE large sphere pile that is stepped as a batch of overlapping block
// scenes (BASELINE.json config 5; no reference counterpart: the reference is one O(n^2) system).
//
// Space is a fixed grid of cubic blocks; a rank owns the blocks of a slab of block-x indices, one scene per
// block. A rebuild takes the dense table of all bodies' states (13 floats per global body id, owners' values,
// all-reduced over the ranks by the caller) and, entirely on the device,
//   k_pile_classify  puts every body into its own block and, as a halo copy, into each of the up to 7
//                    neighbouring blocks whose face / edge / corner it is within halo_w of (atomic slot claim),
//   k_pile_sort      orders every scene's bodies by global id (one CTA per scene, bitonic sort in shared
//                    memory) so that the scene tables, and with them the solver's row order, do not depend
//                    on the order the atomics happened to resolve in,
//   k_pile_fill      writes the scene tables (state, mass, sphere inertia, the fixed container boxes behind the
//                    dynamic bodies) and the owner's address of every owned body,
//   k_pile_halo_map  lists, for every halo copy, where its owner lives: (source, destination) address pairs for
//                    owners on this rank, (destination, global id) requests for the left / right neighbour rank.
// k_pile_export scatters the owners' states back into the table before the next rebuild.
#include "kernels.cuh"

#include <climits>

namespace slrbs {

__device__ __forceinline__ int pile_cell(float rel, float block, int nc)
{
    int c = (int)floorf(rel / block);
    c = c < 0 ? 0 : c;
    return c >= nc ? nc - 1 : c;
}

// claims one slot of a list for every thread with `pred`, one atomic per warp; returns the slot (or -1)
__device__ __forceinline__ int warp_claim(int* ctr, bool pred)
{
    const unsigned act = __activemask();
    const unsigned m = __ballot_sync(act, pred);
    if (!pred) return -1;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(ctr, __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(256) k_pile_export(View v, PileGrid g, const int* __restrict__ count, const int* __restrict__ keys,
                                                     float* __restrict__ table, unsigned* __restrict__ vmax2)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float s2 = 0.f;
    if (tid < (size_t)g.nbdyn * v.B) {
        const int scene = (int)(tid % v.B), slot = (int)(tid / v.B);
        const int n = min(count[scene], g.nbdyn);
        if (slot < n) {
            const int key = keys[(size_t)scene * g.nbdyn + slot];
            if (!(key & 1)) {                                    // owned
                const size_t a = at(v, slot, scene);
                const float4 p = v.pos[a], q = v.quat[a], l = v.vel[a], w = v.omg[a];
                float* o = table + (size_t)13 * (key >> 1);
                o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = q.x; o[4] = q.y; o[5] = q.z; o[6] = q.w;
                o[7] = l.x; o[8] = l.y; o[9] = l.z; o[10] = w.x; o[11] = w.y; o[12] = w.z;
                s2 = l.x * l.x + l.y * l.y + l.z * l.z;
                if (!(s2 == s2)) s2 = 3.0e38f;                   // NaN speed: report "very fast"
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s2 = fmaxf(s2, __shfl_xor_sync(0xffffffffu, s2, d));
    if ((threadIdx.x & 31) == 0 && s2 > 0.f) atomicMax(vmax2, __float_as_uint(s2));   // s2 >= 0: integer order = float order
}

__global__ void __launch_bounds__(256) k_pile_classify(PileGrid g, int B, const float* __restrict__ table, int* __restrict__ count,
                                                       int* __restrict__ keys, int* __restrict__ info)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= g.n_global) return;
    const float* s = table + (size_t)13 * gid;
    const float rel[3] = { s[0] - g.ox, s[1] - g.oy, s[2] - g.oz };
    const int nc[3] = { g.ncx, g.ncy, g.ncz };
    int cell[3];
    bool lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        cell[a] = pile_cell(rel[a], g.block, nc[a]);
        const float in = rel[a] - (float)cell[a] * g.block;      // position inside the own cell
        lo[a] = in <= g.halo_w;
        hi[a] = (g.block - in) <= g.halo_w;
    }
    warp_claim(&info[6], cell[0] >= g.cx0 && cell[0] < g.cx1);   // bodies this rank owns
    for (int dx = -1; dx <= 1; ++dx) {
        const int cx = cell[0] + dx;
        if (cx < g.cx0 || cx >= g.cx1 || (dx < 0 && !lo[0]) || (dx > 0 && !hi[0])) continue;
        for (int dy = -1; dy <= 1; ++dy) {
            const int cy = cell[1] + dy;
            if (cy < 0 || cy >= g.ncy || (dy < 0 && !lo[1]) || (dy > 0 && !hi[1])) continue;
            for (int dz = -1; dz <= 1; ++dz) {
                const int cz = cell[2] + dz;
                if (cz < 0 || cz >= g.ncz || (dz < 0 && !lo[2]) || (dz > 0 && !hi[2])) continue;
                const int scene = ((cx - g.cx0) * g.ncy + cy) * g.ncz + cz;
                const int slot = atomicAdd(&count[scene], 1);
                if (slot < g.nbdyn) keys[(size_t)scene * g.nbdyn + slot] = gid * 2 + ((dx | dy | dz) != 0 ? 1 : 0);
                else info[0] = 1;                                // capacity overflow: the caller grows the context
            }
        }
    }
    (void)B;
}

// one CTA per scene: bitonic sort of the scene's keys (global id * 2 + halo flag; ids are unique within a scene)
__global__ void __launch_bounds__(512) k_pile_sort(PileGrid g, const int* __restrict__ count, int* __restrict__ keys, int npow2,
                                                   int* __restrict__ info)
{
    extern __shared__ int sk[];
    const int scene = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int n = min(count[scene], g.nbdyn);
    int* k = keys + (size_t)scene * g.nbdyn;
    int m = 1;
    while (m < n) m <<= 1;                                       // sort only the occupied power of two
    (void)npow2;
    for (int i = tid; i < m; i += T) sk[i] = i < n ? k[i] : INT_MAX;
    __syncthreads();
    for (int size = 2; size <= m; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (m >> 1); i += T) {
                const int lo = 2 * i - (i & (stride - 1));       // index of the lower element of the pair
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const int a = sk[lo], b = sk[hi];
                if ((a > b) == up) { sk[lo] = b; sk[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < n; i += T) k[i] = sk[i];
    if (tid == 0) {
        atomicMax(&info[1], count[scene]);                        // unclipped: tells the caller the capacity it needs
        atomicAdd(&info[5], n);
    }
}

__global__ void __launch_bounds__(256) k_pile_fill(View v, PileGrid g, const int* __restrict__ count, const int* __restrict__ keys,
                                                   const float* __restrict__ table, const float* __restrict__ radius,
                                                   const float* __restrict__ mass, const float* __restrict__ fixed10,
                                                   int* __restrict__ loc, int* __restrict__ info)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)v.NB * v.B) return;
    const int scene = (int)(tid % v.B), slot = (int)(tid / v.B);
    const int n = min(count[scene], g.nbdyn);
    if (slot == 0) v.nbodies[scene] = n + g.nfixed;
    if (slot >= n + g.nfixed) return;
    const size_t a = at(v, slot, scene);
    float4 px, q, l = make_float4(0.f, 0.f, 0.f, 0.f), w = make_float4(0.f, 0.f, 0.f, 0.f), ga;
    float diag[3];
    int flags;
    if (slot < n) {
        const int key = keys[(size_t)scene * g.nbdyn + slot];
        const int gid = key >> 1;
        const float* s = table + (size_t)13 * gid;
        const float r = radius[gid], m = mass[gid];
        px = make_float4(s[0], s[1], s[2], m);
        q = make_float4(s[3], s[4], s[5], s[6]);
        l = make_float4(s[7], s[8], s[9], 0.f);
        w = make_float4(s[10], s[11], s[12], 0.f);
        ga = make_float4(r, 0.f, 0.f, 0.f);
        const float inertia = (0.4f * m) * (r * r);              // Sphere::computeInertia, Geometry.h:38-43
        diag[0] = diag[1] = diag[2] = inertia;
        flags = GEOM_SPHERE;
        if (!(key & 1)) loc[gid] = scene * v.NB + slot;
    } else {
        const float* f = fixed10 + 10 * (slot - n);              // container box: x3 q4 dim3, mass 1, fixed
        px = make_float4(f[0], f[1], f[2], 1.0f);
        q = make_float4(f[3], f[4], f[5], f[6]);
        ga = make_float4(f[7], f[8], f[9], 0.f);
        diag[0] = (f[8] * f[8] + f[9] * f[9]) / 12.0f;           // Box::computeInertia, Geometry.h:63-70 (mass 1)
        diag[1] = (f[7] * f[7] + f[9] * f[9]) / 12.0f;
        diag[2] = (f[7] * f[7] + f[8] * f[8]) / 12.0f;
        flags = GEOM_BOX | FLAG_FIXED;
    }
    v.pos[a] = px; v.quat[a] = q; v.vel[a] = l; v.omg[a] = w;
    v.flags[a] = flags;
    v.geomA[a] = ga; v.geomB[a] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const bool d = (k == 0 || k == 4 || k == 8);
        v.ibody[at_b9(v, k, slot, scene)] = d ? diag[k / 4] : 0.f;
        v.ibodyinv[at_b9(v, k, slot, scene)] = d ? 1.0f / diag[k / 4] : 0.f;
    }
}

__global__ void __launch_bounds__(256) k_pile_halo_map(View v, PileGrid g, const int* __restrict__ count, const int* __restrict__ keys,
                                                       const float* __restrict__ table, const int* __restrict__ loc,
                                                       int* __restrict__ info, int* __restrict__ hsrc, int* __restrict__ hdst,
                                                       int* __restrict__ ldst, int* __restrict__ lgid, int* __restrict__ rdst,
                                                       int* __restrict__ rgid)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)g.nbdyn * v.B) return;
    const int scene = (int)(tid / g.nbdyn), slot = (int)(tid % g.nbdyn);
    if (slot >= min(count[scene], g.nbdyn)) return;
    const int key = keys[(size_t)scene * g.nbdyn + slot];
    if (!(key & 1)) return;
    const int gid = key >> 1;
    const int cx = pile_cell(table[(size_t)13 * gid] - g.ox, g.block, g.ncx);      // the owner's block-x decides its rank
    const int flat = scene * v.NB + slot;
    int i;
    if ((i = warp_claim(&info[2], cx >= g.cx0 && cx < g.cx1)) >= 0) { hsrc[i] = loc[gid]; hdst[i] = flat; }
    if ((i = warp_claim(&info[3], cx < g.cx0)) >= 0) { ldst[i] = flat; lgid[i] = gid; }
    if ((i = warp_claim(&info[4], cx >= g.cx1)) >= 0) { rdst[i] = flat; rgid[i] = gid; }
}

// a neighbour rank asked for the bodies `gids` (in its own order): where do their owned copies live here?
__global__ void k_pile_lookup(const int* __restrict__ gids, int n, int n_global, const int* __restrict__ loc, int* __restrict__ out,
                              int* __restrict__ info)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = gids[i];
    const int a = (g >= 0 && g < n_global) ? loc[g] : -1;
    out[i] = a < 0 ? 0 : a;
    if (a < 0) info[7] = 1;                                      // asked for a body this rank does not own
}

static inline unsigned blocks_for(size_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

void launch_pile_export(const View& v, const PileGrid& g, const int* count, const int* keys, float* table, unsigned* vmax2, cudaStream_t s)
{ k_pile_export<<<blocks_for((size_t)g.nbdyn * v.B, 256), 256, 0, s>>>(v, g, count, keys, table, vmax2); }

size_t pile_sort_smem(int nbdyn) { int m = 1; while (m < nbdyn) m <<= 1; return (size_t)m * sizeof(int); }

void init_pile_kernels() { cudaFuncSetAttribute(k_pile_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024); }

void launch_pile_rebuild(const View& v, const PileGrid& g, const float* table, const float* radius, const float* mass, const float* fixed10,
                         int* count, int* keys, int* loc, int* info, int* hsrc, int* hdst, int* ldst, int* lgid, int* rdst, int* rgid,
                         cudaStream_t s)
{
    cudaMemsetAsync(count, 0, sizeof(int) * (size_t)v.B, s);
    cudaMemsetAsync(info, 0, sizeof(int) * 8, s);
    cudaMemsetAsync(loc, 0xFF, sizeof(int) * (size_t)g.n_global, s);
    k_pile_classify<<<blocks_for((size_t)g.n_global, 256), 256, 0, s>>>(g, v.B, table, count, keys, info);
    int m = 1; while (m < g.nbdyn) m <<= 1;
    k_pile_sort<<<(unsigned)v.B, 512, pile_sort_smem(g.nbdyn), s>>>(g, count, keys, m, info);
    k_pile_fill<<<blocks_for((size_t)v.NB * v.B, 256), 256, 0, s>>>(v, g, count, keys, table, radius, mass, fixed10, loc, info);
    k_pile_halo_map<<<blocks_for((size_t)g.nbdyn * v.B, 256), 256, 0, s>>>(v, g, count, keys, table, loc, info, hsrc, hdst, ldst, lgid, rdst, rgid);
}

void launch_pile_lookup(const int* gids, int n, int n_global, const int* loc, int* out, int* info, cudaStream_t s)
{ if (n > 0) k_pile_lookup<<<blocks_for((size_t)n, 256), 256, 0, s>>>(gids, n, n_global, loc, out, info); }

}  // namespace slrbs

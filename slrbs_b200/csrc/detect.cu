// detect.cu -- collision detection (CollisionDetect::detectCollisions, CollisionDetect.cpp:27-76).
//
// k_detect_warp<GRID>  one WARP per scene. The scene's body proxies are staged in shared memory. For
//   each body i in index order the lanes test up to 32 partners j > i at a time, in ascending j, and
//   contacts are appended in the reference's lexicographic (i, j) order by a warp-ballot prefix sum
//   (up to 4 per pair in the fixed A,B,C,D order).
//   GRID = false: partners are all j > i (the reference's O(n^2) loop, CollisionDetect.cpp:32-34).
//   GRID = true : broad phase. Spheres are binned into a hashed uniform grid of cell size
//     2 r_max + 2e-3; the partners of sphere i are the spheres of its 27 neighbouring cells plus every
//     non-sphere body, restricted to j > i and sorted ascending. A contact needs a centre distance
//     below r_i + r_j + 1e-3 <= cell size - 1e-3, so the candidate set is a superset of the pairs that
//     can touch (hash collisions only add candidates): the emitted contact list is identical.
// k_detect_rows       one CTA per scene, one THREAD per sphere row (scenes of >= 128 bodies; the default for them).
// k_detect_cta<GRID>  one CTA per scene, one warp per row (kept for scenes too large for k_detect_rows' shared memory).
// k_detect_serial      one thread per scene, for scenes of a few bodies (a car has 15 pairs).
// The kernels themselves are in detect_kernels.inl, compiled twice: namespace dref (the reference's pair dispatch) and
// namespace dext (plus the extension pairs, slrbs_set_extensions); this file holds the shared definitions and launchers.
#include "kernels.cuh"
#include <cstdlib>
#include "narrowphase.cuh"

namespace slrbs {

enum { CAND_CAP = 128 };          // candidate partners of one row in the grid path; longer rows fall back to all j > i
enum { HIT_CAP = 32, HIT_WORDS = 24 };   // contacts pairs of one grid row buffered for ordering; more -> all j > i walk

struct DetectSmem {               // per-warp shared-memory arrays
    float4* sp;                   // [NB] proxies x y z | sphere radius
    int* sf;                      // [NB] flags
    int* cell_start;              // [H+1] grid: bucket offsets (after the fill: bucket ends)
    int* cell_items;              // [NB] grid: sphere indices grouped by bucket
    int* special;                 // [NB] grid: non-sphere body indices, ascending
    int* counter;                 // [2] grid: number of candidates gathered for the current row
    int* cand;                    // [CAND_CAP] grid: candidate partners of the current row (unordered)
    int* order;                   // [HIT_CAP] grid: hit slots in ascending-j order
    int* hits;                    // [HIT_CAP][HIT_WORDS] grid: j, b0, b1, cnt, n, p[4], phi[4] of the pairs that touch
};

__host__ __device__ inline size_t detect_smem_ints(int NB, bool grid, int H)
{
    size_t n = (size_t)NB;                                     // sf
    if (grid) n += (size_t)H + 1 + NB + NB + 2 + CAND_CAP + HIT_CAP + HIT_CAP * HIT_WORDS;
    return n;
}

__device__ __forceinline__ int cell_hash(int ix, int iy, int iz, int Hmask)
{
    return (int)(((unsigned)ix * 73856093u) ^ ((unsigned)iy * 19349663u) ^ ((unsigned)iz * 83492791u)) & Hmask;
}

enum { DETECT_CTA_THREADS = 512 };

__host__ __device__ inline size_t detect_cta_smem_ints(int NB, bool grid, int H, int warps)
{
    size_t n = (size_t)NB + NB + 2;                            // sf, rowcnt
    if (grid) n += (size_t)H + 1 + NB + NB + (size_t)(2 + CAND_CAP + HIT_CAP + HIT_CAP * HIT_WORDS) * warps;
    return n;
}

#ifndef SLRBS_ROW_HITS
#define SLRBS_ROW_HITS 12     // touching partners j > i kept per sphere row (fuller rows take the warp-per-row walk)
#endif
enum { ROW_HITS = SLRBS_ROW_HITS, ROW_DENSE = 0xFF };

__host__ __device__ inline size_t detect_rows_smem_bytes(int NB, int H)
{
    // sp | sf, rowoff[NB+1], cell_start[H+2], cell_items, special, dense_list | hits u16[NB][ROW_HITS] | nh u8[NB]
    return (size_t)NB * 16 + ((size_t)NB * 5 + H + 8) * 4 + (size_t)NB * ROW_HITS * 2 + (((size_t)NB + 15) & ~(size_t)15);
}

#define DETECT_EXT false
namespace dref {
#include "detect_kernels.inl"
}
#undef DETECT_EXT
#define DETECT_EXT true
namespace dext {
#include "detect_kernels.inl"
}
#undef DETECT_EXT
// picks the instantiation with or without the extension pairs
#define DETECT_KERNEL(v, name) ((v).ext_pairs ? dext::name : dref::name)

bool detect_warp_fits(int NB)
{
    const bool grid = NB >= 512;
    const size_t warp_bytes = (size_t)NB * 16 + detect_smem_ints(NB, grid, 4096) * 4;
    const size_t cta_bytes = (size_t)NB * 16 + detect_cta_smem_ints(NB, grid, 4096, DETECT_CTA_THREADS / 32) * 4;
    return warp_bytes <= 200 * 1024 || cta_bytes <= 200 * 1024;
}

// per-device kernel attributes (called from slrbs_create with the device current)
void init_detect_kernels()
{
#define DETECT_ATTR(ns)                                                                                          \
    cudaFuncSetAttribute(ns::k_detect_warp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);    \
    cudaFuncSetAttribute(ns::k_detect_warp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);     \
    cudaFuncSetAttribute(ns::k_detect_cta<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);     \
    cudaFuncSetAttribute(ns::k_detect_cta<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);      \
    cudaFuncSetAttribute(ns::k_detect_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    DETECT_ATTR(dref)
    DETECT_ATTR(dext)
#undef DETECT_ATTR
}

// Buckets of the hashed grid: more buckets per body mean fewer colliding cells per bucket and so fewer spurious
// candidates (1005-body scenes: 1.00 ms with one bucket per body, 0.87 ms with four), as long as the larger table does
// not cost a resident CTA of k_detect_rows. SLRBS_DETECT_HASH forces the factor.
static int grid_buckets(int NB)
{
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("SLRBS_DETECT_HASH"); forced = e ? atoi(e) : 0; }
    auto buckets = [NB](int f) { int H = 64; while (H < NB * f && H < 8192) H <<= 1; return H; };
    if (forced > 0) return buckets(forced);
    // resident CTAs per SM: by shared memory, and by the 64 registers x threads of the kernel
    const int threads = NB >= 512 ? 512 : (NB >= 256 ? 256 : 128);
    auto ctas = [NB, threads](int H) { return min((int)((227 * 1024) / (detect_rows_smem_bytes(NB, H) + 1024)), 65536 / (64 * threads)); };
    int f = 4;
    while (f > 1 && ctas(buckets(f)) < ctas(buckets(1))) f >>= 1;
    return buckets(f);
}

// SLRBS_DETECT_ROWS=0 disables the thread-per-row kernel (A/B against the warp-per-row kernels)
static bool use_rows_kernel(const View& v, bool grid, int H)
{
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("SLRBS_DETECT_ROWS"); enabled = e ? atoi(e) : 1; }
    return enabled && grid && v.NB >= 128 && v.NB <= 65535 && detect_rows_smem_bytes(v.NB, H) <= 200 * 1024;
}

// name of the kernel launch_detect() picks (same decision tree), for slrbs_describe
const char* detect_variant(const View& v, int broadphase)
{
    if (v.pc_pairs) return "k_pair_candidates + k_pair_narrow + k_pair_order";
    const bool grid = broadphase > 0 || (broadphase < 0 && v.NB >= 128);
    const int H = grid_buckets(v.NB);
    if (use_rows_kernel(v, grid, H)) return "k_detect_rows";
    const size_t cta_smem = (size_t)v.NB * 16 + detect_cta_smem_ints(v.NB, grid, H, DETECT_CTA_THREADS / 32) * 4;
    if (((v.NB >= 512 && v.B <= 4096) || (v.NB >= 128 && v.B <= 512)) && cta_smem <= 200 * 1024)
        return grid ? "k_detect_cta<grid>" : "k_detect_cta<all-pairs>";
    const size_t per_warp = (size_t)v.NB * 16 + detect_smem_ints(v.NB, grid, H) * 4;
    int warps = 8;
    while (warps > 1 && per_warp * warps > 96 * 1024) warps >>= 1;
    if ((v.NB < 24 && !v.level_mode && !v.n_sdf) || per_warp * warps > 200 * 1024) return "k_detect_serial";
    return grid ? "k_detect_warp<grid>" : "k_detect_warp<all-pairs>";
}

void launch_detect(const View& v, int broadphase, cudaStream_t s)
{
    // broadphase: 0 = off (all pairs), 1 = hashed grid, -1 = auto (grid for scenes of >= 128 bodies)
    if (v.pc_pairs) { launch_detect_pairs(v, s); return; }       // mixed scenes of >= 128 bodies, scenes too large for one CTA (detect_pairs.cu)
    const bool grid = broadphase > 0 || (broadphase < 0 && v.NB >= 128);
    const int H = grid_buckets(v.NB);
    // scenes of >= 128 bodies: one CTA per scene, one thread per sphere row
    if (use_rows_kernel(v, grid, H)) {
        static int forced = -1;
        if (forced < 0) { const char* e = getenv("SLRBS_DETECT_THREADS"); forced = e ? atoi(e) : 0; }
        const int threads = forced ? forced : (v.NB >= 512 ? 512 : (v.NB >= 256 ? 256 : 128));
        DETECT_KERNEL(v, k_detect_rows)<<<(unsigned)v.B, threads, detect_rows_smem_bytes(v.NB, H), s>>>(v, H);
        return;
    }
    // big scenes in a batch too small to fill the GPU with one warp per scene: one CTA (16 warps) per scene
    const size_t cta_smem = (size_t)v.NB * 16 + detect_cta_smem_ints(v.NB, grid, H, DETECT_CTA_THREADS / 32) * 4;
    if (((v.NB >= 512 && v.B <= 4096) || (v.NB >= 128 && v.B <= 512)) && cta_smem <= 200 * 1024) {
        if (grid) DETECT_KERNEL(v, k_detect_cta<true>)<<<(unsigned)v.B, DETECT_CTA_THREADS, cta_smem, s>>>(v, H);
        else DETECT_KERNEL(v, k_detect_cta<false>)<<<(unsigned)v.B, DETECT_CTA_THREADS, cta_smem, s>>>(v, H);
        return;
    }
    const size_t per_warp = (size_t)v.NB * 16 + detect_smem_ints(v.NB, grid, H) * 4;
    int warps = 8;
    while (warps > 1 && per_warp * warps > 96 * 1024) warps >>= 1;
    const size_t smem = per_warp * warps;
    // tiny scenes (car: 15 pairs) leave a warp idle: one thread per scene is faster there (not with mesh bodies, whose
    // pairs the warp walks together); huge ones do not fit
    if ((v.NB < 24 && !v.level_mode && !v.n_sdf) || smem > 200 * 1024) { DETECT_KERNEL(v, k_detect_serial)<<<(unsigned)((v.B + 127) / 128), 128, 0, s>>>(v); return; }
    const unsigned blocks = (unsigned)((v.B + warps - 1) / warps);
    if (grid) DETECT_KERNEL(v, k_detect_warp<true>)<<<blocks, warps * 32, smem, s>>>(v, warps, H);
    else DETECT_KERNEL(v, k_detect_warp<false>)<<<blocks, warps * 32, smem, s>>>(v, warps, H);
}

}  // namespace slrbs

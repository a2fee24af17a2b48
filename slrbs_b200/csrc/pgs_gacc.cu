// pgs_gacc.cu -- level-scheduled Box-PGS for contact-only scenes, one warp per scene, body accumulators in GLOBAL memory.
//
// The scene-resident kernels keep a scene's accumulators (24 B x bodies) in shared memory, which caps an SM at 9
// thousand-body scenes = 9 warps, and a level step is a ~2500-cycle dependent chain (row loads -> accumulator gather ->
// solve -> scatter -> barrier): nine warps cannot cover each other's waits, the sweep runs at 0.6 of the HBM peak with two
// thirds of the issue slots empty (profiles/r01_k_pgs_lv_mb1k_ncu_details.txt, r02_k_pgs_pipe_*: prefetching the next
// chunk's rows into registers does not help, ptxas drains the load scoreboard at the first branch). What the chain needs
// is more warps, and the only thing in the way is where the accumulators live. Here they are rows of a global table
// ([scene][body][2 x float4], 32 KB per thousand-body scene): private to their warp, re-read within a few levels, so they
// sit in L1 / L2 (a wave of 18 scenes per SM holds 43 MB of them per die against 63 MB of L2; the row stream passes
// through with an evict-first policy and without allocating in L1), and the kernel needs no shared memory at all:
// 18 warps per SM at 96 registers, twice the residency, a whole 2664-scene batch in ONE wave instead of two.
// The gather costs an L1/L2 round trip per level step instead of a shared-memory access; with twice the warps in flight the
// sweep moves from latency bound towards the HBM roof.
//
// Same schedule, same per-row functions as k_pgs_lv / k_pgs_pipe / k_pgs (pgs_math.cuh): impulses, constraint forces and
// trajectories are bit-identical (tests/test_gpu_level_mode.py).
#include "kernels.cuh"
#include "pgs_level.cuh"

#include <cstdlib>

namespace slrbs {

namespace {

__device__ __forceinline__ float4 ldg_stream(const float4* p, unsigned long long policy)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;\n"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(policy));
    return r;
}
// accumulators of one body: linear | angular (two float4, one 32 B sector)
__device__ __forceinline__ void ld_acc(const float4* w, int body, V3& l, V3& a) { l = mk3(w[2 * body]); a = mk3(w[2 * body + 1]); }
__device__ __forceinline__ void st_acc(float4* w, int body, const V3& l, const V3& a) { w[2 * body] = mk4(l, 0.f); w[2 * body + 1] = mk4(a, 0.f); }

}  // namespace

template <bool STREAM, bool COMPACT>
__global__ void __launch_bounds__(32, 18) k_pgs_gacc(View v, float dt, int iters, float mu)
{
    constexpr int NF = COMPACT ? CROW_COMPACT_FIELDS : CROW_LAMBDA;
    unsigned long long policy = 0;
    if (STREAM) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(policy));
    const int scene = blockIdx.x, lane = threadIdx.x;
    const int nb = v.nbodies[scene], nc = v.ncontacts[scene];
    const int ncl = nc > 0 ? v.nclev[scene] : 0;                  // no rows, no levels (stale tables are never walked)
    float4* rows = v.crow + (size_t)scene * CROW_FIELDS * (size_t)v.NC;
    asm volatile("" : "+l"(rows));
    float4* w = v.wacc + (size_t)scene * v.NB * 2;
    asm volatile("" : "+l"(w));
    const unsigned uNC = (unsigned)v.NC;
    const int* lev_end = v.clev_end + scene;
    const size_t B = (size_t)v.B;
    const bool l2pf = v.B >= 256;

    for (int i = lane; i < 2 * nb; i += 32) w[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    for (int it = 0; it < iters; ++it) {                          // SolverBoxPGS.cpp:217-248
        for (int lev = 1, p0 = 0; lev <= ncl; ++lev) {
            const int p1 = lev_end[(size_t)lev * B];
            if (l2pf) {
                // pull the NEXT level's rows (or the first level of the next sweep) towards L2 while this one is solved
                int q0 = p1, q1 = p1;
                if (lev < ncl) q1 = lev_end[(size_t)(lev + 1) * B];
                else if (it + 1 < iters) { q0 = 0; q1 = lev_end[B]; }
                if (q1 > q0) prefetch_rows_l2(rows, uNC, q0, q1, NF, lane, 32);
            }
            for (int p = p0 + lane; p < p1; p += 32) {
                float4 S[NF];
                const unsigned q = (unsigned)p;
#pragma unroll
                for (int f = 0; f < NF; ++f) S[f] = STREAM ? ldg_stream(rows + (q + (unsigned)f * uNC), policy) : __ldg(rows + (q + (unsigned)f * uNC));
                float4* const lp = rows + (q + (unsigned)CROW_LAMBDA * uNC);
                const float4 lam = *lp;
                float4 X[CROW_LAMBDA];
                if (COMPACT) crow_expand(S, X);
                const float4* F = COMPACT ? X : S;
                const ContactIds c = contact_ids(F);
                V3 w0l = mk3(0.f, 0.f, 0.f), w0a = w0l, w1l = w0l, w1a = w0l;
                if (c.dyn0) ld_acc(w, c.id0, w0l, w0a);
                if (c.dyn1) ld_acc(w, c.id1, w1l, w1a);
                const float4 out = contact_solve(F, lam, mu, c.dyn0, c.dyn1, w0l, w0a, w1l, w1a);
                if (c.dyn0) st_acc(w, c.id0, w0l, w0a);
                if (c.dyn1) st_acc(w, c.id1, w1l, w1a);
                *lp = out;
            }
            p0 = p1;
            __syncwarp();                                         // orders this level's accumulator stores before the next level's loads
        }
    }

    // constraint forces (RigidBodySystem.cpp:202-220) in the same per-body order: w becomes fc / tauc
    for (int i = lane; i < 2 * nb; i += 32) w[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    for (int lev = 1, p0 = 0; lev <= ncl; ++lev) {
        const int p1 = lev_end[(size_t)lev * B];
        for (int p = p0 + lane; p < p1; p += 32) {
            float4 F[CROW_LAMBDA];
            if (COMPACT) {
                float4 C[CROW_COMPACT_FIELDS];
#pragma unroll
                for (int f = 0; f < 5; ++f) C[f] = __ldg(rows + ((unsigned)p + (unsigned)f * uNC));
#pragma unroll
                for (int f = 5; f < CROW_COMPACT_FIELDS; ++f) C[f] = make_float4(0.f, 0.f, 0.f, 0.f);     // contact_force reads n t b a0 a1 only
                crow_expand(C, F);
            } else {
#pragma unroll
                for (int f = 0; f < 9; ++f) F[f] = __ldg(rows + ((unsigned)p + (unsigned)f * uNC));
            }
            const float4 lam = rows[(unsigned)p + (unsigned)CROW_LAMBDA * uNC];
            const ContactIds c = contact_ids(F);
            V3 fl, fa, wl, wa;
            if (c.dyn0) { contact_force(F, lam, dt, 0, fl, fa); ld_acc(w, c.id0, wl, wa); st_acc(w, c.id0, wl + fl, wa + fa); }
            if (c.dyn1) { contact_force(F, lam, dt, 1, fl, fa); ld_acc(w, c.id1, wl, wa); st_acc(w, c.id1, wl + fl, wa + fa); }
        }
        p0 = p1;
        __syncwarp();
    }
    for (int i = lane; i < nb; i += 32) {
        V3 fl, fa;
        ld_acc(w, i, fl, fa);
        v.wl[at(v, i, scene)] = mk4(fl, 0.f);
        v.wa[at(v, i, scene)] = mk4(fa, 0.f);
    }
    if (v.profiling && lane == 0) atomicAdd(&v.visit_counters[0], (unsigned long long)nc * iters);
}

// ------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------
void init_gacc_kernels()
{
    // no shared memory: leave the whole 256 KB to L1 (the accumulators are re-read within a few levels)
    cudaFuncSetAttribute(k_pgs_gacc<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    cudaFuncSetAttribute(k_pgs_gacc<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    cudaFuncSetAttribute(k_pgs_gacc<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    cudaFuncSetAttribute(k_pgs_gacc<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
}

// Contact-only contexts (group 32); SLRBS_PGS_GACC=1 selects it (read when a context is created).
bool level_gacc_wanted(const View& v, int group)
{
    if (group != 32 || v.NJ != 0) return false;
    // opt-in: measured slower than the shared-memory kernels on marble_box_1k x 2664 (5.9 vs 4.7 ms, profiles/README.md)
    const char* e = getenv("SLRBS_PGS_GACC");
    return e && atoi(e) != 0;
}

void launch_pgs_gacc(const View& v, float dt, int iters, float mu, cudaStream_t s)
{
    static int stream = -1;                                      // SLRBS_PIPE_STREAM=0: plain row loads (A/B switch)
    if (stream < 0) { const char* e = getenv("SLRBS_PIPE_STREAM"); stream = e ? atoi(e) : 1; }
    const bool st = stream && v.B >= 256;
    if (st) { if (v.crow_compact) k_pgs_gacc<true, true><<<(unsigned)v.B, 32, 0, s>>>(v, dt, iters, mu); else k_pgs_gacc<true, false><<<(unsigned)v.B, 32, 0, s>>>(v, dt, iters, mu); }
    else    { if (v.crow_compact) k_pgs_gacc<false, true><<<(unsigned)v.B, 32, 0, s>>>(v, dt, iters, mu); else k_pgs_gacc<false, false><<<(unsigned)v.B, 32, 0, s>>>(v, dt, iters, mu); }
}

}  // namespace slrbs

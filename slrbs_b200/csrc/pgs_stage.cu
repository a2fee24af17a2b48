// pgs_stage.cu -- level-scheduled Box-PGS for contact-only scenes, one warp per scene, the NEXT chunk's row records
// copied global -> shared memory asynchronously (cp.async / LDGSTS) while the current chunk is solved.
//
// A level step of the scene-resident solver is a dependent chain: row loads (~900 cycles even from L2) -> accumulator
// gather -> solve -> scatter -> barrier, ~2500 cycles, and the 24 KB of accumulators per thousand-body scene cap an SM at
// 9 such warps. Taking the load wait off the chain needs the next chunk's rows in flight during the solve:
//   * into registers (k_pgs_pipe): issued early but not overlapped -- ptxas drains the load scoreboard at the first
//     branch of the solve (profiles/README.md);
//   * into more warps (k_pgs_gacc, accumulators in L1/L2, 18 warps per SM): the gather then costs an L2 round trip per
//     step and the SM ends up slower than with 9 shared-memory warps.
// cp.async has no destination register and is tracked by commit / wait groups, not by the scoreboard, so the copy really
// runs under the solve. It needs a landing buffer: one chunk of 32 rows x 14 fields of the COMPACT record (13 read-only
// fields + lambda, 7 KB) per scene, which with the accumulators (24 KB) leaves room for 7 scenes per SM instead of 9. The
// buffer is single: a lane copies its row into registers, the warp barriers, the next chunk's copy is issued into the same
// buffer, then the solve runs.
//
// Same schedule, same per-row functions as the other solver kernels: bit-identical impulses and trajectories
// (tests/test_gpu_level_mode.py).
#include "kernels.cuh"
#include "pgs_level.cuh"

#include <cstdlib>

namespace slrbs {

namespace {

enum { STAGE_FIELDS = CROW_COMPACT_FIELDS + 1, STAGE_BYTES = STAGE_FIELDS * 32 * 16 };     // 12 read-only fields + rr.. + lambda

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

struct Chunk { int it, lev, base, p1; bool ok; };

// level ends, 32 at a time in a register per lane and handed out by shuffle: a dependent global load in front of every
// level's copy would put an L2 round trip back on the chain
struct LevelEnds {
    const int* lev_end; size_t B; int ncl, blk, ends;
    __device__ __forceinline__ void init(const int* p, size_t b, int n) { lev_end = p; B = b; ncl = n; blk = -1; ends = 0; }
    __device__ __forceinline__ int end_of(int L, int lane)     // end position of level L (1-based)
    {
        const int b = (L - 1) >> 5;
        if (b != blk) {                                           // warp-uniform
            const int l = (b << 5) + 1 + lane;
            ends = l <= ncl ? lev_end[(size_t)l * B] : 0;
            blk = b;
        }
        return __shfl_sync(0xffffffffu, ends, (L - 1) & 31);
    }
};

}  // namespace

#define ld_w AccSh<true>::ld
#define st_w AccSh<true>::st

// accumulators: 24 B per body (three float2); stage: [field][lane] float4 behind them
__global__ void __launch_bounds__(32, 7) k_pgs_stage(View v, float dt, int iters, float mu)
{
    extern __shared__ float4 smem_st4[];
    const int scene = blockIdx.x, lane = threadIdx.x;
    float4* stage = smem_st4;                                     // [STAGE_FIELDS][32]
    float* w = reinterpret_cast<float*>(smem_st4 + STAGE_FIELDS * 32);
    unsigned wsh = (unsigned)__cvta_generic_to_shared(w);        // accumulators in the shared state space (pgs_level.cuh: AccSh)
    asm volatile("" : "+r"(wsh));
    const int nb = v.nbodies[scene], nc = v.ncontacts[scene];
    const int ncl = nc > 0 ? v.nclev[scene] : 0;                  // no rows, no levels (stale tables are never walked)
    float4* rows = v.crow + (size_t)scene * CROW_FIELDS * (size_t)v.NC;
    asm volatile("" : "+l"(rows));
    const unsigned uNC = (unsigned)v.NC;
    LevelEnds le;
    le.init(v.clev_end + scene, (size_t)v.B, ncl);
    const bool l2pf = v.B >= 256;

    for (int i = lane; i < 6 * nb; i += 32) w[i] = 0.f;
    __syncwarp();

    // next chunk; entering a level, pull the level after it towards L2 (one prefetch per 128 B line of every stored field)
    auto advance = [&](Chunk c) -> Chunk {
        c.base += 32;
        if (c.base >= c.p1) {
            if (c.lev < ncl) { c.lev += 1; c.base = c.p1; }
            else { c.it += 1; c.lev = 1; c.base = 0; if (c.it >= iters) { c.ok = false; return c; } }
            c.p1 = le.end_of(c.lev, lane);
            if (l2pf) {
                int q0 = c.p1, q1 = c.p1;
                if (c.lev < ncl) q1 = le.end_of(c.lev + 1, lane);
                else if (c.it + 1 < iters) { q0 = 0; q1 = le.end_of(1, lane); }
                if (q1 > q0) prefetch_rows_l2(rows, uNC, q0, q1, CROW_COMPACT_FIELDS, lane, 32);
            }
        }
        return c;
    };
    // asynchronous copy of a chunk's records into the stage: lane l brings row base + l, field by field
    auto issue = [&](const Chunk& c) {
        if (c.ok && c.base + lane < c.p1) {
            const unsigned q = (unsigned)(c.base + lane);
#pragma unroll
            for (int f = 0; f < CROW_COMPACT_FIELDS; ++f) cp_async16(stage + f * 32 + lane, rows + (q + (unsigned)f * uNC));
            cp_async16(stage + CROW_COMPACT_FIELDS * 32 + lane, rows + (q + (unsigned)CROW_LAMBDA * uNC));
        }
        cp_async_commit();
    };

    if (ncl > 0 && iters > 0) {                                   // SolverBoxPGS.cpp:217-248
        const int first_end = le.end_of(1, lane);
        // a sweep of one chunk would fetch the impulses it is about to write: such scenes wait for their copy before solving
        const bool single = ncl == 1 && first_end <= 32;
        Chunk cur; cur.it = 0; cur.lev = 1; cur.base = 0; cur.p1 = first_end; cur.ok = true;
        if (l2pf && ncl > 1) {
            const int q0 = first_end, q1 = le.end_of(2, lane);
            if (q1 > q0) prefetch_rows_l2(rows, uNC, q0, q1, CROW_COMPACT_FIELDS, lane, 32);
        }
        issue(cur);
        while (cur.ok) {
            const bool live = cur.base + lane < cur.p1;
            cp_async_wait_all();
            __syncwarp();
            float4 S[CROW_COMPACT_FIELDS], lam = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
#pragma unroll
                for (int f = 0; f < CROW_COMPACT_FIELDS; ++f) S[f] = stage[f * 32 + lane];
                lam = stage[CROW_COMPACT_FIELDS * 32 + lane];
            }
            __syncwarp();                                         // every lane has its row: the stage is free again
            const Chunk nxt = advance(cur);
            if (!single) issue(nxt);
            if (live) {
                float4 F[CROW_LAMBDA];
                crow_expand(S, F);
                const ContactIds ci = contact_ids(F);
                V3 w0l = mk3(0.f, 0.f, 0.f), w0a = w0l, w1l = w0l, w1a = w0l;
                if (ci.dyn0) ld_w(wsh, ci.id0, w0l, w0a);
                if (ci.dyn1) ld_w(wsh, ci.id1, w1l, w1a);
                const float4 out = contact_solve(F, lam, mu, ci.dyn0, ci.dyn1, w0l, w0a, w1l, w1a);
                if (ci.dyn0) st_w(wsh, ci.id0, w0l, w0a);
                if (ci.dyn1) st_w(wsh, ci.id1, w1l, w1a);
                rows[(unsigned)(cur.base + lane) + (unsigned)CROW_LAMBDA * uNC] = out;
            }
            __syncwarp();
            if (single) issue(nxt);                               // after this chunk's impulses are written
            cur = nxt;
        }
        cp_async_wait_all();
    }

    // constraint forces (RigidBodySystem.cpp:202-220) in the same per-body order: w becomes fc / tauc
    for (int i = lane; i < 6 * nb; i += 32) w[i] = 0.f;
    __syncwarp();
    for (int lev = 1, p0 = 0; lev <= ncl; ++lev) {
        const int p1 = le.end_of(lev, lane);
        for (int p = p0 + lane; p < p1; p += 32) {
            float4 C[CROW_COMPACT_FIELDS], F[CROW_LAMBDA];
#pragma unroll
            for (int f = 0; f < 5; ++f) C[f] = __ldg(rows + ((unsigned)p + (unsigned)f * uNC));
#pragma unroll
            for (int f = 5; f < CROW_COMPACT_FIELDS; ++f) C[f] = make_float4(0.f, 0.f, 0.f, 0.f);     // contact_force reads n t b a0 a1 only
            crow_expand(C, F);
            const float4 lam = rows[(unsigned)p + (unsigned)CROW_LAMBDA * uNC];
            const ContactIds c = contact_ids(F);
            V3 fl, fa, wl, wa;
            if (c.dyn0) { contact_force(F, lam, dt, 0, fl, fa); ld_w(wsh, c.id0, wl, wa); st_w(wsh, c.id0, wl + fl, wa + fa); }
            if (c.dyn1) { contact_force(F, lam, dt, 1, fl, fa); ld_w(wsh, c.id1, wl, wa); st_w(wsh, c.id1, wl + fl, wa + fa); }
        }
        p0 = p1;
        __syncwarp();
    }
    for (int i = lane; i < nb; i += 32) {
        V3 fl, fa;
        ld_w(wsh, i, fl, fa);
        v.wl[at(v, i, scene)] = mk4(fl, 0.f);
        v.wa[at(v, i, scene)] = mk4(fa, 0.f);
    }
    if (v.profiling && lane == 0) atomicAdd(&v.visit_counters[0], (unsigned long long)nc * iters);
}

// ------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------
static size_t stage_smem(int NB) { return (size_t)STAGE_BYTES + (((size_t)NB * 24 + 15) & ~(size_t)15); }

void init_stage_kernels()
{
    cudaFuncSetAttribute(k_pgs_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
}

// Opt-in (SLRBS_PGS_STAGE=1, read when a context is created) for contact-only contexts whose scenes get one warp each
// (group 32). History: the landing buffer costs residency (7 instead of 9 thousand-body scenes per SM) and bought a level step
// of ~2250 instead of ~3000 cycles against the FIRST register-prefetch kernel, which did not overlap its loads; for batches of
// up to one wave this kernel was then the default. k_pgs_pipe as it is now overlaps its loads at 9 scenes per SM and is
// faster everywhere it was measured (2664 scenes: 3.44 against 4.70 ms; the 5832 block scenes of a 10^6-sphere pile: 3.71
// against 3.97 ms; the 666-scene chunk contexts of the host-driven loop: 4.18 against 4.07 10^8 body-steps/s end to end).
bool level_stage_wanted(const View& v, int group)
{
    if (group != 32 || v.NJ != 0 || stage_smem(v.NB) > 200 * 1024) return false;
    const char* e = getenv("SLRBS_PGS_STAGE");
    return e && atoi(e) != 0;
}

void launch_pgs_stage(const View& v, float dt, int iters, float mu, cudaStream_t s)
{
    k_pgs_stage<<<(unsigned)v.B, 32, stage_smem(v.NB), s>>>(v, dt, iters, mu);
}

}  // namespace slrbs

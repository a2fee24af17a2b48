// pgs_pipe.cu -- level-scheduled Box-PGS for contact-only scenes, one warp per scene, software-pipelined: while the
// warp solves one chunk of rows (<= 32 rows of one dependency level, one per lane) the row records of the NEXT chunk
// are already in flight into a second set of registers.
//
// Why: k_pgs_lv on a thousand-sphere scene walks ~1400 chunk steps per solve (130 levels x 10 sweeps + the force
// scatter), each a strict sequence  16 row loads -> accumulator gather (shared memory) -> ~250 flop chain -> scatter ->
// warp barrier.  ncu (profiles/r01_k_pgs_lv_mb1k_ncu_details.txt): ~2500 cycles per step, ~900 of them waiting for the
// row loads even with the level prefetched into L2; nine such warps per SM (the scene's 24 KB of accumulators are what
// limits residency) cannot cover each other's waits: 0.49 eligible warps per scheduler, DRAM at 0.60 of peak.  Residency
// being capped by shared memory, the register file is half empty (9 warps x 160 registers = 45 KB of 256 KB): a second
// row buffer (64 registers) costs no occupancy and takes the load wait off the chain.  Two lanes per row on the same
// shared memory was the other candidate (pgs_split.cu): measured slower, the chain shortens by less than the extra
// barrier costs.
//
// Same schedule, same per-row functions (pgs_math.cuh) as k_pgs_lv and k_pgs: impulses, constraint forces and
// trajectories are bit-identical (tests/test_gpu_level_mode.py).  Row order inside a level is free (rows of a level
// share no dynamic body), levels are processed in order, which is the reference's sequential sweep
// (SolverBoxPGS.cpp:217-248) regrouped.
#include "kernels.cuh"
#include "pgs_level.cuh"

#include <cstdlib>

namespace slrbs {

#define ld_w AccSh<W24>::ld
#define st_w AccSh<W24>::st

namespace {

// chunk = up to 32 consecutive positions [base, min(base + 32, p1)) of level `lev` in sweep `it`
struct Chunk { int it, lev, base, p1; bool ok; };

// level ends, 32 at a time in a register per lane (one strided load per 32 levels instead of one dependent load in
// front of every level's prefetch)
struct LevelEnds {
    const int* lev_end; size_t B; int ncl, blk, ends;
    __device__ __forceinline__ void init(const int* p, size_t b, int n) { lev_end = p; B = b; ncl = n; blk = -1; ends = 0; }
    // end position of level L (1-based; L = 0 gives 0)
    __device__ __forceinline__ int end_of(int L, int lane)
    {
        if (L <= 0) return 0;
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

// STREAM: the row records are read once per sweep and the whole stream is far larger than L2 -- demand loads then carry an
// L2 evict-first policy (and do not allocate in L1), so that a line leaves L2 as soon as it has been consumed and the lines the
// bulk prefetch brought in for the next level are not the ones pushed out.
__device__ __forceinline__ float4 ldg_stream(const float4* p, unsigned long long policy)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;\n"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(policy));
    return r;
}

template <bool W24, bool STREAM, bool COMPACT>
__global__ void __launch_bounds__(32, 9) k_pgs_pipe(View v, float dt, int iters, float mu)
{
    constexpr int NF = COMPACT ? CROW_COMPACT_FIELDS : CROW_LAMBDA;      // read-only fields of a stored record
    unsigned long long policy = 0;
    if (STREAM) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(policy));
    extern __shared__ float4 smem_pp4[];
    float* w = reinterpret_cast<float*>(smem_pp4);
    unsigned wsh = (unsigned)__cvta_generic_to_shared(w);        // accumulators in the shared state space (pgs_level.cuh: AccSh)
    asm volatile("" : "+r"(wsh));
    const int scene = blockIdx.x, lane = threadIdx.x;
    const int nb = v.nbodies[scene], nc = v.ncontacts[scene];
    const int ncl = nc > 0 ? v.nclev[scene] : 0;                  // no rows, no levels (stale tables are never walked)
    float4* rows = v.crow + (size_t)scene * CROW_FIELDS * (size_t)v.NC;
    asm volatile("" : "+l"(rows));                                // one base register pair; a load's address is one IMAD.WIDE.U32 off it
    const unsigned uNC = (unsigned)v.NC;
    LevelEnds le;
    le.init(v.clev_end + scene, (size_t)v.B, ncl);
    const bool l2pf = v.B >= 256;                                 // the row stream exceeds L2 only in batches

    for (int i = lane; i < Acc<W24>::FLOATS * nb; i += 32) w[i] = 0.f;
    __syncwarp();

    // advance to the chunk after c; entering a level, pull the level after it towards L2 (one bulk prefetch per field, the
    // rows of a level being contiguous per field), so that the register prefetch below finds its rows in L2
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
                if (lane < CROW_FIELDS && (!COMPACT || lane < CROW_COMPACT_FIELDS || lane == CROW_LAMBDA) && q1 > q0)
                    prefetch_l2_bulk(rows + (size_t)lane * uNC + q0, (unsigned)(q1 - q0) * 16u);
            }
        }
        return c;
    };

#define PIPE_ISSUE(F, LAM, c)                                                                                          \
    if ((c).ok && (c).base + lane < (c).p1) {                                                                          \
        const unsigned q_ = (unsigned)((c).base + lane);                                                               \
        _Pragma("unroll") for (int f = 0; f < NF; ++f)                                                                  \
            F[f] = STREAM ? ldg_stream(rows + (q_ + (unsigned)f * uNC), policy) : __ldg(rows + (q_ + (unsigned)f * uNC)); \
        LAM = rows[q_ + (unsigned)CROW_LAMBDA * uNC];                                                                  \
    }
#define PIPE_SOLVE(F, LAM, c)                                                                                          \
    if ((c).base + lane < (c).p1) {                                                                                    \
        float4 X[CROW_LAMBDA];                                                                                         \
        if (COMPACT) crow_expand(F, X);                                                                                \
        const float4* E = COMPACT ? X : F;                                                                             \
        const ContactIds ci = contact_ids(E);                                                                          \
        V3 w0l = mk3(0.f, 0.f, 0.f), w0a = w0l, w1l = w0l, w1a = w0l;                                                  \
        if (ci.dyn0) ld_w(wsh, ci.id0, w0l, w0a);                                                                        \
        if (ci.dyn1) ld_w(wsh, ci.id1, w1l, w1a);                                                                        \
        const float4 out = contact_solve(E, LAM, mu, ci.dyn0, ci.dyn1, w0l, w0a, w1l, w1a);                            \
        if (ci.dyn0) st_w(wsh, ci.id0, w0l, w0a);                                                                        \
        if (ci.dyn1) st_w(wsh, ci.id1, w1l, w1a);                                                                        \
        rows[(unsigned)((c).base + lane) + (unsigned)CROW_LAMBDA * uNC] = out;                                         \
    }                                                                                                                  \
    __syncwarp();

    if (ncl > 0 && iters > 0) {                                   // SolverBoxPGS.cpp:217-248
        // a sweep of a single chunk would prefetch the impulses it is about to write: such scenes (<= 32 rows in one level)
        // run unpipelined
        const int first_end = le.end_of(1, lane);
        const bool single = ncl == 1 && first_end <= 32;
        float4 FA[NF], FB[NF], lamA, lamB;
        Chunk cur; cur.it = 0; cur.lev = 1; cur.base = 0; cur.p1 = first_end; cur.ok = true;
        if (single) {
            for (int it = 0; it < iters; ++it) {
                PIPE_ISSUE(FA, lamA, cur)
                PIPE_SOLVE(FA, lamA, cur)
            }
        } else {
            if (l2pf && ncl > 1) {
                const int q0 = first_end, q1 = le.end_of(2, lane);
                if (lane < CROW_FIELDS && (!COMPACT || lane < CROW_COMPACT_FIELDS || lane == CROW_LAMBDA) && q1 > q0)
                    prefetch_l2_bulk(rows + (size_t)lane * uNC + q0, (unsigned)(q1 - q0) * 16u);
            }
#ifndef SLRBS_PIPE_TWO_BUFFERS
            // ONE landing buffer FB, copied into FA before the next chunk's loads are issued into it. The copy is the point: ptxas
            // puts every row load on one scoreboard (a counter), so "wait for chunk k's rows" also waits for chunk k+1's loads
            // if those have been issued already. With two alternating buffers (the #else branch, round-2's first version) the
            // solve of chunk k read registers a load had written, and its first instruction carried that wait -- the loads
            // were early but nothing ran under them (profiles/README.md). Here the wait sits on the copy, BEFORE the next
            // loads are issued, and the solve reads only registers written by MOVs: nothing in it touches the load scoreboard
            // (checked in the SASS control bits), so the next chunk's 13 LDG.128 are in flight for the whole solve.
            PIPE_ISSUE(FB, lamB, cur)
            while (true) {
                _Pragma("unroll") for (int f = 0; f < NF; ++f) FA[f] = FB[f];
                lamA = lamB;
                Chunk nxt = advance(cur);
                PIPE_ISSUE(FB, lamB, nxt)
                PIPE_SOLVE(FA, lamA, cur)
                if (!nxt.ok) break;
                cur = nxt;
            }
#else
            PIPE_ISSUE(FA, lamA, cur)
            while (true) {
                Chunk nxt = advance(cur);
                PIPE_ISSUE(FB, lamB, nxt)
                PIPE_SOLVE(FA, lamA, cur)
                if (!nxt.ok) break;
                cur = advance(nxt);
                PIPE_ISSUE(FA, lamA, cur)
                PIPE_SOLVE(FB, lamB, nxt)
                if (!cur.ok) break;
            }
#endif
        }
    }
#undef PIPE_ISSUE
#undef PIPE_SOLVE

    // constraint forces (RigidBodySystem.cpp:202-220) in the same per-body order: w becomes fc / tauc
    for (int i = lane; i < Acc<W24>::FLOATS * nb; i += 32) w[i] = 0.f;
    __syncwarp();
    for (int lev = 1, p0 = 0; lev <= ncl; ++lev) {
        const int p1 = le.end_of(lev, lane);
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
// padded 32 B accumulators unless they would keep the SM below 9 scenes
static bool pp_pack24(int NB) { return ((size_t)NB * 32 + 1024) * 9 > 227u * 1024u; }

template <bool W24, bool STREAM, bool COMPACT>
static void pipe_variant(const View& v, float dt, int iters, float mu, cudaStream_t s, bool launch)
{
    if (launch) k_pgs_pipe<W24, STREAM, COMPACT><<<(unsigned)v.B, 32, (size_t)v.NB * (W24 ? 24 : 32), s>>>(v, dt, iters, mu);
    else cudaFuncSetAttribute(k_pgs_pipe<W24, STREAM, COMPACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
}
static void pipe_dispatch(const View& v, float dt, int iters, float mu, cudaStream_t s, bool launch, bool w24, bool stream, bool compact)
{
#define PV(a, b, c) pipe_variant<a, b, c>(v, dt, iters, mu, s, launch)
    if (w24) { if (stream) { if (compact) PV(true, true, true); else PV(true, true, false); } else { if (compact) PV(true, false, true); else PV(true, false, false); } }
    else     { if (stream) { if (compact) PV(false, true, true); else PV(false, true, false); } else { if (compact) PV(false, false, true); else PV(false, false, false); } }
#undef PV
}

void init_pipe_kernels()
{
    View v{};
    for (int k = 0; k < 8; ++k) pipe_dispatch(v, 0.f, 0, 0.f, nullptr, false, k & 1, k & 2, k & 4);
}

// contexts without joint capacity whose scenes get one warp each (group 32)
bool level_pipe_wanted(const View& v, int group)
{
    const char* e = getenv("SLRBS_PGS_PIPE");                    // read when a context is created; 0 keeps k_pgs_lv (A/B)
    const int on = e ? atoi(e) : 1;
    return on && group == 32 && v.NJ == 0 && (size_t)v.NB * 24 <= 200 * 1024;
}
// 13-field row records for those contexts (SLRBS_ROWS_COMPACT, read when a context is created)
bool pipe_rows_compact()
{
    const char* e = getenv("SLRBS_ROWS_COMPACT");
    return e ? atoi(e) != 0 : true;       // measured: 4.75 -> 4.51 ms per launch on marble_box_1k x 2664
}

void launch_pgs_pipe(const View& v, float dt, int iters, float mu, cudaStream_t s)
{
    static int stream = -1;                                      // SLRBS_PIPE_STREAM=1: evict-first row loads (A/B switch)
    if (stream < 0) { const char* e = getenv("SLRBS_PIPE_STREAM"); stream = e ? atoi(e) : 0; }
    pipe_dispatch(v, dt, iters, mu, s, true, pp_pack24(v.NB), stream && v.B >= 256, v.crow_compact != 0);
}

}  // namespace slrbs

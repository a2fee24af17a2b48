// pgs_pipe.cu -- level-scheduled Box-PGS for contact-only scenes, one warp per scene, software-pipelined: while the
// warp solves one chunk of rows (<= 32 rows of one dependency level, one per lane) the row records of the NEXT chunk
// are in flight into landing registers and a window of rows further ahead is being pulled into L2.
//
// Why: a thousand-sphere scene is ~1400 chunk steps per solve (130 levels x 10 sweeps + the force pass), each a strict
// sequence  13 row loads -> accumulator gather (shared memory) -> ~220 instruction chain -> scatter -> warp barrier, and
// only nine such warps fit an SM (the scene's 24 KB of accumulators are what limits residency): they cannot cover each
// other's load waits (round 1: 0.49 eligible warps per scheduler, DRAM at 0.60 of peak). Residency being capped by shared
// memory, the register file is half empty: a landing buffer costs no occupancy and takes the load wait off the chain.
// What it took to make the loads really overlap (SASS-level findings, profiles/README.md): a copy of the landing
// registers in front of the next loads, divisions without the slow-path CALL, a chunk table instead of a walk over the
// level ends, and per-lane L2 prefetches instead of cp.async.bulk.prefetch (which the compiler serialises over the lanes).
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

// One entry of a scene's chunk table: up to 32 consecutive positions of ONE level, `base | count << 16` (positions are below
// 65535 in level mode, slrbs_create). A sweep is the table walked front to back; every sweep walks the same table.
__device__ __forceinline__ int ch_base(unsigned e) { return (int)(e & 0xffffu); }
__device__ __forceinline__ int ch_count(unsigned e) { return (int)((e >> 16) & 0x3fu); }

// cursor over (sweep, chunk) for `sweeps` passes over a table of `nch` entries. The entry after the one handed out is
// already being loaded (one uniform 4 B load per chunk, issued a whole chunk before it is needed), so next() costs neither a
// shuffle nor a branch on the chain.
struct ChunkCursor {
    const unsigned* tab; int nch, sweeps, k, it; unsigned ahead;
    __device__ __forceinline__ void init(const unsigned* t, int n, int sw)
    {
        tab = t; nch = n; sweeps = sw; k = 0; it = 0;
        ahead = (sw > 0 && n > 0) ? t[0] : 0u;                    // the table was written by this warp: plain loads
    }
    // next entry, 0 when all sweeps have been handed out
    __device__ __forceinline__ unsigned next()
    {
        const unsigned e = ahead;
        if (++k == nch) { k = 0; ++it; }
        ahead = it < sweeps ? tab[k] : 0u;
        return e;
    }
};

}  // namespace

// STREAM: the row records are read once per sweep and the whole stream is far larger than L2 -- demand loads then carry an
// L2 evict-first policy (and do not allocate in L1), so that a line leaves L2 as soon as it has been consumed and the lines the
// L2 window brought in for the chunks ahead are not the ones pushed out. (SLRBS_PIPE_STREAM=1; measured neutral, off by default.)
__device__ __forceinline__ float4 ldg_stream(const float4* p, unsigned long long policy)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;\n"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(policy));
    return r;
}

__device__ __forceinline__ float copy_f32(float x)
{
    float y;
    asm("add.f32 %0, %1, 0f80000000;" : "=f"(y) : "f"(x));
    return y;
}

template <bool W24, bool STREAM, bool COMPACT>
__global__ void __launch_bounds__(32, 9) k_pgs_pipe(View v, float dt, int iters, float mu, int l2mode)
{
    constexpr int NF = COMPACT ? CROW_COMPACT_FIELDS : CROW_LAMBDA;      // read-only fields of a stored record
    constexpr int NFORCE = COMPACT ? 5 : 9;                       // the fields contact_force needs: n t b rr0 rr1 | n t b a0* a1*
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

    for (int i = lane; i < Acc<W24>::FLOATS * nb; i += 32) w[i] = 0.f;

    // ---- chunk table of the scene: levels in order, each cut into chunks of <= 32 rows ------------------------------
    // (the sweep loop used to find its next chunk by walking level ends: 140 instructions per chunk and a third of the
    // kernel's stall samples, profiles/r02_k_pgs_pipe_copy_mb1k_*; with the table it is one uniform load, issued a chunk ahead)
    unsigned* tab = v.cchunk + (size_t)scene * v.cchunk_cap;
    int nch = 0;
    for (int l0 = 1, carry_end = 0; l0 <= ncl; l0 += 32) {        // 32 levels per trip, one per lane
        const int l = l0 + lane;
        const int e1 = l <= ncl ? v.clev_end[at(v, l, scene)] : 0;
        int e0 = __shfl_up_sync(0xffffffffu, e1, 1);
        if (lane == 0) e0 = carry_end;
        carry_end = __shfl_sync(0xffffffffu, e1, 31);
        const int n = l <= ncl ? (e1 - e0 + 31) >> 5 : 0;         // chunks of this level
        int incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const int off = nch + incl - n;
        for (int k = 0; k < n; ++k) tab[off + k] = (unsigned)(e0 + 32 * k) | ((unsigned)min(32, e1 - e0 - 32 * k) << 16);
        nch += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();

    // A pipelined loop over iters sweeps (SolverBoxPGS.cpp:217-248), then one over the pass that forms the constraint forces
    // from the final impulses (RigidBodySystem.cpp:202-220; same table, same per-body order: the accumulators are cleared and
    // become fc / tauc). While chunk g is worked on, the records of chunk g + 1 are in flight into the landing registers FB
    // and a window of rows 64 .. 96 positions further on is being pulled towards L2.
    // ISF (a literal): the pass that forms the constraint forces -- 5 (9) fields + the impulse, no impulse written back
#define PIPE_ISSUE(F, LAM, e, ISF)                                                                                     \
    if (lane < ch_count(e)) {                                                                                          \
        const unsigned q_ = (unsigned)(ch_base(e) + lane);                                                             \
        if (ISF) {                                                                                                     \
            _Pragma("unroll") for (int f = 0; f < NFORCE; ++f) F[f] = __ldg(rows + (q_ + (unsigned)f * uNC));          \
        } else {                                                                                                       \
            _Pragma("unroll") for (int f = 0; f < NF; ++f)                                                              \
                F[f] = STREAM ? ldg_stream(rows + (q_ + (unsigned)f * uNC), policy) : __ldg(rows + (q_ + (unsigned)f * uNC)); \
        }                                                                                                              \
        LAM = rows[q_ + (unsigned)CROW_LAMBDA * uNC];                                                                  \
    }
    // Pulling rows towards L2 ahead of the register loads (batches only: one scene's rows fit L2; l2mode 0 = off): a window of
    // 32 positions x every stored field, kept 64 .. 96 positions ahead of the chunk whose loads are being issued; positions
    // only grow inside a sweep, so the window just slides (two prefetch instructions per lane and window: 13 or 16 fields x four
    // 128 B lines). Round 2's first version issued one cp.async.bulk.prefetch per field of the chunk after next: that
    // instruction takes its address from a uniform register, so the compiler serialises the 13 lanes' prefetches in a loop --
    // 100 instructions per chunk and 36 % of the kernel's stall samples (profiles/r02_k_pgs_pipe_table_mb1k_*).
    int psw = 0, ppos = 0;                                        // next window: sweep, first position (multiple of 32)
    auto l2_window = [&](int isw, int ibase) {
        const int ncr = (nc + 31) & ~31;
        if (psw > iters || (psw - isw) * ncr + ppos - ibase >= 96) return;
#pragma unroll
        for (int j = 0; j < ((NF + 1) * 4 + 31) / 32; ++j) {
            const int idx = lane + 32 * j, fi = idx >> 2, row = ppos + (idx & 3) * 8;
            if (fi <= NF && row < nc) prefetch_l2(rows + ((unsigned)(fi < NF ? fi : CROW_LAMBDA) * uNC + (unsigned)row));
        }
        ppos += 32;
        if (ppos >= nc) { ppos = 0; ++psw; }
    };
#define PIPE_SOLVE(F, LAM, e, ISF)                                                                                     \
    if (lane < ch_count(e)) {                                                                                          \
        float4 X[CROW_LAMBDA];                                                                                         \
        if (COMPACT) crow_expand(F, X);                                                                                \
        const float4* E = COMPACT ? X : F;                                                                             \
        const ContactIds ci = contact_ids(E);                                                                          \
        if (ISF) {                                                                                                     \
            V3 fl, fa, wl, wa;                                                                                         \
            if (ci.dyn0) { contact_force(E, LAM, dt, 0, fl, fa); ld_w(wsh, ci.id0, wl, wa); st_w(wsh, ci.id0, wl + fl, wa + fa); } \
            if (ci.dyn1) { contact_force(E, LAM, dt, 1, fl, fa); ld_w(wsh, ci.id1, wl, wa); st_w(wsh, ci.id1, wl + fl, wa + fa); } \
        } else {                                                                                                       \
            V3 w0l = mk3(0.f, 0.f, 0.f), w0a = w0l, w1l = w0l, w1a = w0l;                                              \
            if (ci.dyn0) ld_w(wsh, ci.id0, w0l, w0a);                                                                    \
            if (ci.dyn1) ld_w(wsh, ci.id1, w1l, w1a);                                                                    \
            const float4 out = contact_solve(E, LAM, mu, ci.dyn0, ci.dyn1, w0l, w0a, w1l, w1a);                        \
            if (ci.dyn0) st_w(wsh, ci.id0, w0l, w0a);                                                                    \
            if (ci.dyn1) st_w(wsh, ci.id1, w1l, w1a);                                                                    \
            rows[(unsigned)(ch_base(e) + lane) + (unsigned)CROW_LAMBDA * uNC] = out;                                   \
        }                                                                                                              \
    }                                                                                                                  \
    __syncwarp();
#define PIPE_CLEAR_FOR_FORCES()                                                                                        \
    { for (int i = lane; i < Acc<W24>::FLOATS * nb; i += 32) w[i] = 0.f; __syncwarp(); }
    // ONE landing buffer FB, copied into FA before the next chunk's loads are issued into it. The copy is the point: ptxas puts
    // every row load on one scoreboard (a counter), so "wait for chunk k's rows" also waits for chunk k+1's loads if those
    // have been issued already. With two alternating buffers (round 2's first version) the solve of chunk k read registers a
    // load had written, and its first instruction carried that wait -- the loads were early but nothing ran under them
    // (profiles/README.md). Here the wait sits on the copy, BEFORE the next loads are issued, and the solve reads only
    // registers written by the copy: nothing in it touches the load scoreboard (checked in the SASS control bits,
    // tools/sass_ctrl.py), so the next chunk's LDG.128 are in flight for the whole solve. The copy goes through the FP32 pipe
    // (one warp instruction per cycle; a MOV issues every second cycle): x + (-0) is x for every float, -0 and denormals
    // included; field 2's w holds the packed body ids (any bit pattern) and is moved.
#define PIPE_LOOP(ISF, NSWEEPS, SW0)                                                                                   \
    {                                                                                                                  \
        cc.init(tab, nch, NSWEEPS);                                                                                    \
        unsigned eI = cc.next();                              /* chunk whose loads are issued next */              \
        int isw = SW0;                                            /* its sweep, for the L2 window */                   \
        if (l2mode && !(ISF)) { l2_window(0, 0); l2_window(0, 0); l2_window(0, 0); }                                   \
        PIPE_ISSUE(FB, lamB, eI, ISF)                                                                                  \
        while (eI != 0u) {                                                                                             \
            const unsigned eS = eI;                                                                                    \
            _Pragma("unroll") for (int f = 0; f < ((ISF) ? NFORCE : NF); ++f) {                                         \
                FA[f].x = copy_f32(FB[f].x); FA[f].y = copy_f32(FB[f].y); FA[f].z = copy_f32(FB[f].z);                 \
                FA[f].w = f == 2 ? FB[f].w : copy_f32(FB[f].w);                                                        \
            }                                                                                                          \
            lamA = lamB;                                                                                               \
            eI = cc.next();                                                                                        \
            if (l2mode && eI != 0u) {                                                                                  \
                if (ch_base(eI) < ch_base(eS)) ++isw;             /* positions only grow inside a sweep */             \
                l2_window(isw, ch_base(eI));                      /* one window per chunk keeps up: a chunk is <= 32 positions */ \
            }                                                                                                          \
            PIPE_ISSUE(FB, lamB, eI, ISF)                         /* no lane is below the count of entry 0 */          \
            PIPE_SOLVE(FA, lamA, eS, ISF)                                                                              \
        }                                                                                                              \
    }

    if (nch > 0) {
        ChunkCursor cc;
        float4 FA[NF], FB[NF], lamA, lamB;
        if (nch == 1) {
            // a table of one chunk would prefetch the impulses it is about to write: such scenes (<= 32 rows in one level)
            // run unpipelined
            cc.init(tab, nch, iters + 1);
            for (int it = 0; it < iters; ++it) {
                const unsigned e = cc.next();
                PIPE_ISSUE(FA, lamA, e, false)
                PIPE_SOLVE(FA, lamA, e, false)
            }
            const unsigned e = cc.next();
            PIPE_CLEAR_FOR_FORCES()
            PIPE_ISSUE(FA, lamA, e, true)
            PIPE_SOLVE(FA, lamA, e, true)
        } else {
            // two loops over the same table: the sweeps, then (accumulators cleared) the pass that forms the constraint forces --
            // kept apart so that neither carries the other's branches (a predicate costs a branch ~18 cycles on this chain);
            // the one load wait exposed between them is a microsecond per scene
            PIPE_LOOP(false, iters, 0)
            PIPE_CLEAR_FOR_FORCES()
            PIPE_LOOP(true, 1, iters)
        }
    } else {
        __syncwarp();
    }
#undef PIPE_ISSUE
#undef PIPE_SOLVE
#undef PIPE_CLEAR_FOR_FORCES
#undef PIPE_LOOP

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
static void pipe_variant(const View& v, float dt, int iters, float mu, cudaStream_t s, bool launch, int l2mode)
{
    if (launch) k_pgs_pipe<W24, STREAM, COMPACT><<<(unsigned)v.B, 32, (size_t)v.NB * (W24 ? 24 : 32), s>>>(v, dt, iters, mu, l2mode);
    else cudaFuncSetAttribute(k_pgs_pipe<W24, STREAM, COMPACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
}
static void pipe_dispatch(const View& v, float dt, int iters, float mu, cudaStream_t s, bool launch, bool w24, bool stream, bool compact, int l2mode)
{
#define PV(a, b, c) pipe_variant<a, b, c>(v, dt, iters, mu, s, launch, l2mode)
    if (w24) { if (stream) { if (compact) PV(true, true, true); else PV(true, true, false); } else { if (compact) PV(true, false, true); else PV(true, false, false); } }
    else     { if (stream) { if (compact) PV(false, true, true); else PV(false, true, false); } else { if (compact) PV(false, false, true); else PV(false, false, false); } }
#undef PV
}

void init_pipe_kernels()
{
    View v{};
    for (int k = 0; k < 8; ++k) pipe_dispatch(v, 0.f, 0, 0.f, nullptr, false, k & 1, k & 2, k & 4, 0);
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
    static int l2 = -1;                                          // SLRBS_PIPE_L2 = 0: no L2 prefetch (A/B switch)
    if (l2 < 0) { const char* e = getenv("SLRBS_PIPE_L2"); l2 = e ? (atoi(e) != 0) : 1; }
    // the row stream exceeds L2 only in batches
    pipe_dispatch(v, dt, iters, mu, s, true, pp_pack24(v.NB), stream && v.B >= 256, v.crow_compact != 0, v.B >= 256 ? l2 : 0);
}

}  // namespace slrbs

// pgs_split.cu -- level-scheduled Box-PGS with TWO lanes per contact row, one per body side.
//
// Same schedule and the same arithmetic as k_pgs_lv (pgs_level.cu): rows of one dependency level in parallel, levels
// in order, which reproduces the reference's sequential sweep (SolverBoxPGS.cpp:217-248) bit for bit. What changes is
// the mapping. A thousand-sphere scene has ~2500 contact rows in ~130 level steps, i.e. ~19 rows per step: one lane per
// row leaves a warp 40 % empty, the scene's accumulators (24 KB of shared memory) cap the SM at 9 such warps, and each
// of them walks a ~400-instruction dependent chain per step (ncu, profiles/r01_k_pgs_lv_mb1k_ncu_details.txt: 2.2
// warps per scheduler, one instruction issued every 6 cycles per warp, 34 % of it waiting on the row loads).
// Here the even lane of a pair owns body0's side of the row and the odd lane body1's:
//   side s loads  n|im0  t|im1  b|ids  a_s{n,t,b}|.  m_s{n,t,b}|.  lambda      (10 of the 17 float4 of a record)
//   c_s[k] = JMinv_s,k . (w_s - J_s^T lambda)                                  its body's term of x = b - sum(...)
//   odd -> even: c_1[0..2], A00 A11 A22, A12 A20 A21  (9 shuffles);  even lane: solveContact;  even -> odd: y[0..2]
//   w_s += J_s^T (y - lambda)                                                  each lane its own body
// x0 = (rhs0 - c_0[0]) - c_1[0] is exactly the order contact_solve uses, so the impulses stay bit-identical
// (tests/test_gpu_level_mode.py). A scene gets 64 threads (32 rows per step) on the same 24 KB: 18 warps per SM with
// chains half as long. Joints (few, and absent from the contact-dense scenes this kernel is chosen for) are solved by
// the even lanes with the one-lane-per-row code.
#include "kernels.cuh"
#include "pgs_level.cuh"

#include <cstdlib>

namespace slrbs {

#define ld_w Acc<W24>::ld
#define st_w Acc<W24>::st

// ROWS = contact rows per level step (threads = 2 * ROWS); MINB = CTAs per SM the register allocation must allow
template <int ROWS, int MINB, bool W24>
__global__ void __launch_bounds__(2 * ROWS, MINB) k_pgs_sp(View v, float dt, int iters, float mu)
{
    extern __shared__ float4 smem_sp4[];
    float* w = reinterpret_cast<float*>(smem_sp4);
    const int scene = blockIdx.x;
    const int tid = threadIdx.x, r = tid >> 1, side = tid & 1;
    const float sgn = side ? -1.f : 1.f;
    const int nb = v.nbodies[scene], nj = v.njoints[scene], nc = v.ncontacts[scene];
    const int njl = nj > 0 ? v.njlev[scene] : 0, ncl = nc > 0 ? v.nclev[scene] : 0;
    const size_t NC = (size_t)v.NC;
    float4* rows = v.crow + (size_t)scene * CROW_FIELDS * NC;
    asm volatile("" : "+l"(rows));                                // keep the scene's base in one register pair (else every load re-adds it)
    const int* const lev_end = v.clev_end + scene;                // clev_end[at(v, lev, scene)] = lev_end[lev * B]
    const size_t B = (size_t)v.B;
    // 32-bit record offsets (16 fields x NC rows of a scene stay far below 2^31 float4): one IMAD.WIDE per load
    // unsigned 32-bit record offsets (16 fields x NC rows of a scene stay far below 2^32 float4): the address of a load is
    // one IMAD.WIDE.U32 off the scene's base
    const unsigned uNC = (unsigned)v.NC, oA = (3u + 3u * side) * uNC, oM = (9u + 3u * side) * uNC, oL = (unsigned)CROW_LAMBDA * uNC;

    for (int i = tid; i < Acc<W24>::FLOATS * nb; i += 2 * ROWS) w[i] = 0.f;
    __syncthreads();
    // joints keep last step's lambda (Hinge.cpp:33-63 never zeroes it): seed w with it, level by level
    for (int lev = 1, p0 = 0; lev <= njl; ++lev) {
        const int p1 = v.jlev_end[at(v, lev, scene)];
        if (side == 0)
            for (int p = p0 + r; p < p1; p += ROWS) {
                const JointRow j = load_joint_head(v, p, scene);
                const int slot = __float_as_int(__ldg(&v.jrow[at_jrow(v, 16, p, scene)]).w);
                const size_t ji = at(v, slot, scene);
                const float4 lam03 = v.jlam[ji];
                const float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, v.jlam4[ji] };
                V3 wl, wa;
                if (j.id0 >= 0) { ld_w(w, j.id0, wl, wa); joint_seed(j, l, 0, wl, wa); st_w(w, j.id0, wl, wa); }
                if (j.id1 >= 0) { const int b = j.id1 & 0x7fffffff; ld_w(w, b, wl, wa); joint_seed(j, l, 1, wl, wa); st_w(w, b, wl, wa); }
            }
        p0 = p1;
        __syncthreads();
    }

    for (int it = 0; it < iters; ++it) {                          // SolverBoxPGS.cpp:217-248
        for (int lev = 1, p0 = 0; lev <= njl; ++lev) {
            const int p1 = v.jlev_end[at(v, lev, scene)];
            if (side == 0)
                for (int p = p0 + r; p < p1; p += ROWS) {
                    JointRec rec;
                    load_joint_rec(v, p, scene, rec);
                    const size_t ji = at(v, joint_slot(rec), scene);
                    const float4 lam03 = v.jlam[ji];
                    float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, v.jlam4[ji] };
                    const bool dyn0 = rec.h.id0 >= 0, dyn1 = rec.h.id1 >= 0;
                    const int b0 = rec.h.id0 & 0x7fffffff, b1 = rec.h.id1 & 0x7fffffff;
                    V3 w0l = mk3(0.f, 0.f, 0.f), w0a = w0l, w1l = w0l, w1a = w0l;
                    if (dyn0) ld_w(w, b0, w0l, w0a);
                    if (dyn1) ld_w(w, b1, w1l, w1a);
                    joint_solve(rec, l, dyn0, dyn1, w0l, w0a, w1l, w1a);
                    if (dyn0) st_w(w, b0, w0l, w0a);
                    if (dyn1) st_w(w, b1, w1l, w1a);
                    v.jlam[ji] = make_float4(l[0], l[1], l[2], l[3]);
                    v.jlam4[ji] = l[4];
                }
            p0 = p1;
            __syncthreads();
        }
        for (int lev = 1, p0 = 0; lev <= ncl; ++lev) {
            const int p1 = lev_end[(size_t)lev * B];
            if (v.B >= 256) {
                // pull the NEXT level's rows (or the first level of the next sweep) towards L2 while this one is solved: the
                // row stream is far larger than L2, and a level would otherwise start with an HBM round trip. The rows of a
                // level are contiguous per field: one prefetch per 128 B line, dealt over the lanes.
                int q0 = p1, q1 = p1;
                if (lev < ncl) q1 = lev_end[(size_t)(lev + 1) * B];
                else if (it + 1 < iters) { q0 = 0; q1 = lev_end[B]; }
                if (q1 > q0) prefetch_rows_l2(rows, (unsigned)NC, q0, q1, CROW_LAMBDA, tid, (int)blockDim.x);
            }
            for (int base = p0; base < p1; base += ROWS) {        // every lane runs every trip: the shuffles below are warp-wide
                const int p = base + r;
                const bool live = p < p1;
                // lanes past the end of the level re-read its first row (same cache lines, no extra traffic) and drop the result
                const unsigned q = (unsigned)(live ? p : p0);
                const float4 Fn = __ldg(rows + q), Ft = __ldg(rows + (q + uNC)), Fb = __ldg(rows + (q + 2u * uNC));
                const float4 An = __ldg(rows + (q + oA)), At = __ldg(rows + (q + oA + uNC)), Ab = __ldg(rows + (q + oA + 2u * uNC));
                const float4 Mn = __ldg(rows + (q + oM)), Mt = __ldg(rows + (q + oM + uNC)), Mb = __ldg(rows + (q + oM + 2u * uNC));
                float4* const lp = rows + (q + oL);
                const float4 lam = *lp;
                const int ids = __float_as_int(Fb.w);
                const bool dyn0 = !(ids & 0x8000), dyn1 = !(ids & 0x80000000);
                const bool dyn = live && (side ? dyn1 : dyn0);
                const int body = (side ? (ids >> 16) : ids) & 0x7fff;
                const V3 n = mk3(Fn), t = mk3(Ft), b = mk3(Fb);
                const V3 an = mk3(An), at_ = mk3(At), ab = mk3(Ab);
                const float l0 = lam.x, l1 = lam.y, l2 = lam.z;
                const V3 dl = lin3(n, l0, t, l1, b, l2);         // J0_lin^T lambda ( = -J1_lin^T lambda)
                V3 wl = mk3(0.f, 0.f, 0.f), wa = wl;
                float c0 = 0.f, c1 = 0.f, c2 = 0.f;
                if (dyn) {
                    ld_w(w, body, wl, wa);
                    const V3 tl = wl - sgn * dl;                  // side 0: w0l - dl, side 1: w1l + dl
                    const V3 ta = wa - lin3(an, l0, at_, l1, ab, l2);
                    const float ims = side ? -Ft.w : Fn.w;        // im0 | -im1
                    c0 = fmaf(ims, dotf(n, tl), dotf(mk3(Mn), ta));
                    c1 = fmaf(ims, dotf(t, tl), dotf(mk3(Mt), ta));
                    c2 = fmaf(ims, dotf(b, tl), dotf(mk3(Mb), ta));
                }
                // odd lane -> even lane: its body's terms, the diagonal and the A12 A20 A21 entries its fields carry
                const float o0 = __shfl_xor_sync(0xffffffffu, c0, 1), o1 = __shfl_xor_sync(0xffffffffu, c1, 1), o2 = __shfl_xor_sync(0xffffffffu, c2, 1);
                const float oA0 = __shfl_xor_sync(0xffffffffu, An.w, 1), oA1 = __shfl_xor_sync(0xffffffffu, At.w, 1), oA2 = __shfl_xor_sync(0xffffffffu, Ab.w, 1);
                const float oM0 = __shfl_xor_sync(0xffffffffu, Mn.w, 1), oM1 = __shfl_xor_sync(0xffffffffu, Mt.w, 1), oM2 = __shfl_xor_sync(0xffffffffu, Mb.w, 1);
                float y0 = 0.f, y1 = 0.f, y2 = 0.f;
                if (live && side == 0) {
                    float x0 = An.w, x1 = At.w, x2 = Ab.w;        // rhs
                    if (dyn0) { x0 -= c0; x1 -= c1; x2 -= c2; }
                    if (dyn1) { x0 -= o0; x1 -= o1; x2 -= o2; }
                    contact_project(x0, x1, x2, l1, l2, oA0, oA1, oA2, Mn.w, Mt.w, Mb.w, oM0, oM1, oM2, mu, y0, y1, y2);
                    *lp = make_float4(y0, y1, y2, 0.f);
                }
                y0 = __shfl_sync(0xffffffffu, y0, (tid & 30));
                y1 = __shfl_sync(0xffffffffu, y1, (tid & 30));
                y2 = __shfl_sync(0xffffffffu, y2, (tid & 30));
                if (dyn) {
                    const float d0 = y0 - l0, d1 = y1 - l1, d2 = y2 - l2;
                    const V3 ddl = lin3(n, d0, t, d1, b, d2);
                    wl = wl + sgn * ddl;                          // side 0: w0l + ddl, side 1: w1l - ddl
                    wa = wa + lin3(an, d0, at_, d1, ab, d2);
                    st_w(w, body, wl, wa);
                }
            }
            p0 = p1;
            __syncthreads();
        }
    }

    // constraint forces (RigidBodySystem.cpp:185-220) in the same per-body order: w becomes fc / tauc
    for (int i = tid; i < Acc<W24>::FLOATS * nb; i += 2 * ROWS) w[i] = 0.f;
    __syncthreads();
    for (int lev = 1, p0 = 0; lev <= njl; ++lev) {
        const int p1 = v.jlev_end[at(v, lev, scene)];
        if (side == 0)
            for (int p = p0 + r; p < p1; p += ROWS) {
                const JointRow j = load_joint_head(v, p, scene);
                const float4 tail = __ldg(&v.jrow[at_jrow(v, 16, p, scene)]);
                const size_t ji = at(v, __float_as_int(tail.w), scene);
                const float4 lam03 = v.jlam[ji];
                const float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, v.jlam4[ji] };
                float f0[6], f1[6];
                joint_force(j, __float_as_int(tail.z), l, dt, f0, f1);
                V3 wl, wa;
                if (j.id0 >= 0) { ld_w(w, j.id0, wl, wa); st_w(w, j.id0, wl + mk3(f0[0], f0[1], f0[2]), wa + mk3(f0[3], f0[4], f0[5])); }
                if (j.id1 >= 0) { const int b = j.id1 & 0x7fffffff; ld_w(w, b, wl, wa); st_w(w, b, wl + mk3(f1[0], f1[1], f1[2]), wa + mk3(f1[3], f1[4], f1[5])); }
            }
        p0 = p1;
        __syncthreads();
    }
    // the reference also adds joint forces to FIXED bodies (:189-192, harmless: they do not integrate).
    // Levels do not separate rows that share a fixed body, so one lane adds those in joint order.
    if (tid == 0) {
        for (int s = 0; s < nj; ++s) {
            const int4 def = v.jdef[at(v, s, scene)];
            const bool fix0 = v.flags[at(v, def.y, scene)] & FLAG_FIXED, fix1 = v.flags[at(v, def.z, scene)] & FLAG_FIXED;
            if (!fix0 && !fix1) continue;
            const int p = v.jpos[at(v, s, scene)];
            const JointRow j = load_joint_head(v, p, scene);
            const size_t ji = at(v, s, scene);
            const float4 lam03 = v.jlam[ji];
            const float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, v.jlam4[ji] };
            float f0[6], f1[6];
            joint_force(j, def.w, l, dt, f0, f1);
            V3 wl, wa;
            if (fix0) { ld_w(w, def.y, wl, wa); st_w(w, def.y, wl + mk3(f0[0], f0[1], f0[2]), wa + mk3(f0[3], f0[4], f0[5])); }
            if (fix1) { ld_w(w, def.z, wl, wa); st_w(w, def.z, wl + mk3(f1[0], f1[1], f1[2]), wa + mk3(f1[3], f1[4], f1[5])); }
        }
    }
    __syncthreads();
    for (int lev = 1, p0 = 0; lev <= ncl; ++lev) {
        const int p1 = lev_end[(size_t)lev * B];
        for (int p = p0 + r; p < p1; p += ROWS) {
            const float4* R = rows + p;
            float4 F[9];                                          // contact_force reads n t b and the a-fields of its side
            F[0] = __ldg(R); F[1] = __ldg(R + NC); F[2] = __ldg(R + 2 * NC);
            const float4* Ra = R + (size_t)(3 + 3 * side) * NC;
            const float4 A0 = __ldg(Ra), A1 = __ldg(Ra + NC), A2 = __ldg(Ra + 2 * NC);
            F[3] = A0; F[4] = A1; F[5] = A2; F[6] = A0; F[7] = A1; F[8] = A2;
            const float4 lam = rows[(size_t)CROW_LAMBDA * NC + p];
            const int ids = __float_as_int(F[2].w);
            const bool dyn = side ? !(ids & 0x80000000) : !(ids & 0x8000);
            const int body = (side ? (ids >> 16) : ids) & 0x7fff;
            if (dyn) {
                V3 fl, fa, wl, wa;
                contact_force(F, lam, dt, side, fl, fa);
                ld_w(w, body, wl, wa);
                st_w(w, body, wl + fl, wa + fa);
            }
        }
        p0 = p1;
        __syncthreads();
    }
    for (int i = tid; i < nb; i += 2 * ROWS) {
        V3 fl, fa;
        ld_w(w, i, fl, fa);
        v.wl[at(v, i, scene)] = mk4(fl, 0.f);
        v.wa[at(v, i, scene)] = mk4(fa, 0.f);
    }
    if (v.profiling && tid == 0) {
        atomicAdd(&v.visit_counters[0], (unsigned long long)nc * iters);
        atomicAdd(&v.visit_counters[1], (unsigned long long)nj * iters);
    }
}

// ------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------
// padded 32 B accumulators unless they would keep the SM below 8 scenes
static bool sp_pack24(int NB) { return ((size_t)NB * 32 + 1024) * 8 > 227u * 1024u; }
size_t split_smem_bytes(int NB) { return (size_t)NB * (sp_pack24(NB) ? 24 : 32); }

template <int ROWS, int MINB>
static void launch_sp(const View& v, float dt, int iters, float mu, cudaStream_t s)
{
    const size_t smem = split_smem_bytes(v.NB);
    if (sp_pack24(v.NB)) k_pgs_sp<ROWS, MINB, true><<<(unsigned)v.B, 2 * ROWS, smem, s>>>(v, dt, iters, mu);
    else k_pgs_sp<ROWS, MINB, false><<<(unsigned)v.B, 2 * ROWS, smem, s>>>(v, dt, iters, mu);
}

void init_split_kernels()
{
    cudaFuncSetAttribute(k_pgs_sp<32, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_pgs_sp<32, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_pgs_sp<64, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_pgs_sp<64, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
}

// rows per level step: 32 (64 threads per scene) for batches, 64 for a few wide scenes
void launch_pgs_split(const View& v, float dt, int iters, float mu, int rows, cudaStream_t s)
{
    if (rows >= 64) launch_sp<64, 4>(v, dt, iters, mu, s);
    else launch_sp<32, 8>(v, dt, iters, mu, s);
}

}  // namespace slrbs

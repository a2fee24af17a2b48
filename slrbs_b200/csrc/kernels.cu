// kernels.cu -- hand-written sm_100a kernels of the per-step rigid-body pipeline.
//
// Compiled with -fmad=false: the narrow phase, the Jacobian/RHS assembly, the force scatter
// and the integrator reproduce the reference's fp32 arithmetic operation by operation (see
// vecmath.cuh); only the PGS sweep uses (explicit) fused multiply-adds.
//
// Kernel            replaces (reference file:line)                                   mapping
// k_prepare         RigidBodySystem.cpp:58-65, RigidBody.cpp:78-89                    thread / (body, scene)
// k_detect_warp     CollisionDetect.cpp:27-76,102-274, Contact.cpp:11-26              warp / scene (ballot-ordered emit)
// k_contact_rows    Contact.cpp:33-94, SolverBoxPGS.cpp:27-44,165-189,206-211         thread / (contact, scene)
// k_joint_rows      Hinge.cpp:33-63, Spherical.cpp:34-52, SolverBoxPGS.cpp:134-163    thread / (joint, scene)
// k_pgs             SolverBoxPGS.cpp:217-248, RigidBodySystem.cpp:185-220             thread / scene (sequential rows)
// k_integrate       RigidBodySystem.cpp:99-120                                        thread / (body, scene)
#include "kernels.cuh"
#include "pgs_math.cuh"

#include <cstdlib>
#include "vecmath.cuh"

namespace slrbs {

// ------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ M3 load_m3(const float* base, const View& v, int slot, int scene)
{
    M3 A;
#pragma unroll
    for (int k = 0; k < 9; ++k) A.m[k / 3][k % 3] = base[at_b9(v, k, slot, scene)];
    return A;
}
__device__ __forceinline__ void store_m3(float* base, const View& v, int slot, int scene, const M3& A)
{
#pragma unroll
    for (int k = 0; k < 9; ++k) base[at_b9(v, k, slot, scene)] = A.m[k / 3][k % 3];
}

// ------------------------------------------------------------------------------------
// K1  force reset + world inertia
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_prepare(View v)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)v.NB * v.B) return;
    const int scene = (int)(tid % v.B), slot = (int)(tid / v.B);
    if (tid == 0 && v.nclev_max) *v.nclev_max = 0;
    if (slot == 0) {                                             // CollisionDetect::clear, RigidBodySystem.cpp:75
        v.ncontacts[scene] = 0;
        if (v.nclev) v.nclev[scene] = 0;                         // the level kernel walks levels, not counts: no stale rows when detection is skipped
        if (v.pc_count) { v.pc_count[scene] = 0; v.pc_hits[scene] = 0; }
    }
    if (slot >= v.nbodies[scene]) return;
    const size_t i = at(v, slot, scene);
    const float4 px = v.pos[i];
    const float mass = px.w;
    // b->f = b->mass * Vector3f(0, -9.81, 0); tau = 0  (RigidBodySystem.cpp:60-61)
    V3 f = mass * mk3(0.f, -9.81f, 0.f);
    V3 tau = mk3(0.f, 0.f, 0.f);
    const int ext = v.has_ext ? v.ext_mode[scene] : 0;           // per scene: a sub-range call must not change the other scenes
    if (ext == 1) {                                              // pre-step callback contribution (:70-73)
        f = f + mk3(v.extf[i]);
        tau = tau + mk3(v.extt[i]);
    } else if (ext == 2) {                                       // callback result given absolutely
        f = mk3(v.extf[i]);
        tau = mk3(v.extt[i]);
    }
    v.force[i] = mk4(f, 0.f);
    v.torque[i] = mk4(tau, 0.f);
    const int fl = v.flags[i];
    const float4 ga = v.geomA[i];
    v.cproxy[i] = make_float4(px.x, px.y, px.z, ((fl & FLAG_TYPE_MASK) == GEOM_SPHERE) ? ga.x : 0.f);
    if ((fl & FLAG_TYPE_MASK) == GEOM_BOX) {
        // half extents of the box's world AABB: |R| * dim/2 (only used to skip sphere-box tests that cannot touch)
        const M3 R = rotation_matrix(mkq(v.quat[i]));
        const V3 hd = mk3(0.5f * ga.x, 0.5f * ga.y, 0.5f * ga.z);
        v.bext[i] = make_float4(fabsf(R.m[0][0]) * hd.x + fabsf(R.m[0][1]) * hd.y + fabsf(R.m[0][2]) * hd.z,
                                fabsf(R.m[1][0]) * hd.x + fabsf(R.m[1][1]) * hd.y + fabsf(R.m[1][2]) * hd.z,
                                fabsf(R.m[2][0]) * hd.x + fabsf(R.m[2][1]) * hd.y + fabsf(R.m[2][2]) * hd.z, 0.f);
    }
    // RigidBody::updateInertiaMatrix (RigidBody.cpp:78-89)
    M3 Iinv;
    if (!(fl & FLAG_FIXED)) {
        const Q4 q = mkq(v.quat[i]);
        const M3 R = rotation_matrix(q);
        const M3 Rinv = rotation_matrix(qinverse(q));
        const M3 Ib = load_m3(v.ibody, v, slot, scene);
        const M3 Ibi = load_m3(v.ibodyinv, v, slot, scene);
        store_m3(v.Iw, v, slot, scene, matmul(matmul(R, Ib), Rinv));
        Iinv = matmul(matmul(R, Ibi), Rinv);
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) Iinv.m[k / 3][k % 3] = 0.f;
    }
    store_m3(v.Iinvw, v, slot, scene, Iinv);
    // everything the row builders gather per body, as one 128 B line
    float4* rec = v.brec + at_brec(v, slot, scene);
    const float4 vl = v.vel[i], om = v.omg[i];
    rec[0] = px;
    rec[1] = make_float4(vl.x, vl.y, vl.z, (fl & FLAG_FIXED) ? 1.f : 0.f);
    rec[2] = make_float4(om.x, om.y, om.z, 0.f);
    rec[3] = mk4(f, 0.f);
    rec[4] = mk4(tau, 0.f);
    rec[5] = make_float4(Iinv.m[0][0], Iinv.m[0][1], Iinv.m[0][2], 0.f);
    rec[6] = make_float4(Iinv.m[1][0], Iinv.m[1][1], Iinv.m[1][2], 0.f);
    rec[7] = make_float4(Iinv.m[2][0], Iinv.m[2][1], Iinv.m[2][2], 0.f);
}

// ------------------------------------------------------------------------------------
// K4  contact frame, Jacobian blocks, J M^-1, 3x3 diagonal block, RHS
// ------------------------------------------------------------------------------------
#ifndef CONTACT_ROWS_MIN_BLOCKS
#define CONTACT_ROWS_MIN_BLOCKS 5
#endif
struct BodyDyn { V3 x, xdot, omega, f, tau; float mass; M3 Iinv; bool fixed; };

__device__ __forceinline__ BodyDyn load_body_dyn(const View& v, int slot, int scene)
{
    BodyDyn b;
    const float4* __restrict__ rec = v.brec + at_brec(v, slot, scene);
    const float4 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3], r4 = rec[4], r5 = rec[5], r6 = rec[6], r7 = rec[7];
    b.x = mk3(r0); b.mass = r0.w;
    b.xdot = mk3(r1); b.fixed = r1.w != 0.f;
    b.omega = mk3(r2);
    b.f = mk3(r3); b.tau = mk3(r4);
    b.Iinv.m[0][0] = r5.x; b.Iinv.m[0][1] = r5.y; b.Iinv.m[0][2] = r5.z;
    b.Iinv.m[1][0] = r6.x; b.Iinv.m[1][1] = r6.y; b.Iinv.m[1][2] = r6.z;
    b.Iinv.m[2][0] = r7.x; b.Iinv.m[2][1] = r7.y; b.Iinv.m[2][2] = r7.z;
    return b;
}

// multAndSub (SolverBoxPGS.cpp:13-21) on one row: b -= (a * G[k]) * x[k], k = 0..5
__device__ __forceinline__ float mult_and_sub_row(float b, const float* G, const V3& x, const V3& y, float a)
{
    b -= (a * G[0]) * x.x; b -= (a * G[1]) * x.y; b -= (a * G[2]) * x.z;
    b -= (a * G[3]) * y.x; b -= (a * G[4]) * y.y; b -= (a * G[5]) * y.z;
    return b;
}

// J M^-1 row: linear (1/mass) * J_lin, angular J_ang * Iinv   (Contact.cpp:90-93, Hinge.cpp:59-62)
__device__ __forceinline__ void jminv_row(const float* J, float* JM, const BodyDyn& b)
{
    const float im = 1.0f / b.mass;
    JM[0] = im * J[0]; JM[1] = im * J[1]; JM[2] = im * J[2];
    const V3 a = vecmat(mk3(J[3], J[4], J[5]), b.Iinv);
    JM[3] = a.x; JM[4] = a.y; JM[5] = a.z;
}

__device__ __forceinline__ float dot6_seq(const float* a, const float* b)
{
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) acc += a[k] * b[k];
    return acc;
}

// normalized() with the division spelled out (pgs_math.cuh: div_rn_nocall, same bits): a tangent always has a component
// that is exactly zero (t = n x e_x), and a zero numerator sends the compiler's division through its out-of-range
// subroutine -- several calls per contact in this kernel (profiles/r02_k_contact_rows_lv_mb1k_ncu_details.txt)
__device__ __forceinline__ V3 normalized_nocall(const V3& a)
{
    const float z = sqnorm(a);
    if (!(z > 0.0f)) return a;
    const float d = sqrtf(z);
    return mk3(div_rn_nocall(a.x, d), div_rn_nocall(a.y, d), div_rn_nocall(a.z, d));
}

// one contact: frame, Jacobian blocks, J M^-1, diagonal block, RHS -> its row record at `pos` of the scene's stream
__device__ __forceinline__ void contact_row(const View& v, float h, int scene, int slot, int pos)
{
    const size_t c = atc(v, slot, scene);
    const int2 ids = v.cids[c];
    const float4 pp = v.cp[c];
    const V3 p = mk3(pp), n = mk3(v.cn[c]);
    const float pene = pp.w;
    const BodyDyn b0 = load_body_dyn(v, ids.x, scene);
    const BodyDyn b1 = load_body_dyn(v, ids.y, scene);

    // Contact::computeContactFrame (Contact.cpp:33-53)
    V3 t = cross(n, mk3(1.f, 0.f, 0.f));
    if (norm(t) < 1e-5f) t = cross(n, mk3(0.f, 1.f, 0.f));
    t = normalized_nocall(t);
    const V3 bb = normalized_nocall(cross(n, t));

    // Contact::computeJacobian (Contact.cpp:55-94)
    const V3 rr0 = p - b0.x, rr1 = p - b1.x;
    float J0[3][6], J1[3][6], M0[3][6], M1[3][6];
    const V3 d[3] = { n, t, bb };
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        set_row(J0[k], d[k], cross(rr0, d[k]));
        set_row(J1[k], -d[k], -cross(rr1, d[k]));
        jminv_row(J0[k], M0[k], b0);
        jminv_row(J1[k], M1[k], b1);
    }
    // diagonal block (SolverBoxPGS.cpp:165-189) and RHS (buildRHS :27-44)
    float A[3][3], rhs[3];
    const float hinv = 1.0f / h;
    const float phi[3] = { pene, 0.f, 0.f };
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            float a = 0.f;
            if (!b0.fixed) a += dot6_seq(M0[r], J0[cc]);
            if (!b1.fixed) a += dot6_seq(M1[r], J1[cc]);
            A[r][cc] = a;
        }
        float b = (-hinv * 0.3f) * phi[r];
        if (!b0.fixed) { b = mult_and_sub_row(b, M0[r], b0.f, b0.tau, h); b = mult_and_sub_row(b, J0[r], b0.xdot, b0.omega, 1.0f); }
        if (!b1.fixed) { b = mult_and_sub_row(b, M1[r], b1.f, b1.tau, h); b = mult_and_sub_row(b, J1[r], b1.xdot, b1.omega, 1.0f); }
        rhs[r] = b;
    }
    const int packed = (ids.x | (b0.fixed ? 0x8000 : 0)) | ((ids.y | (b1.fixed ? 0x8000 : 0)) << 16);
    float4* R = v.crow;
#define CROW(f) R[at_crow(v, f, pos, scene)]
    if (v.crow_compact) {                                       // 13 fields (pgs_math.cuh: crow_expand)
        CROW(0)  = mk4(n, 1.0f / b0.mass);
        CROW(1)  = mk4(t, 1.0f / b1.mass);
        CROW(2)  = mk4(bb, __int_as_float(packed));
        CROW(3)  = mk4(rr0, rhs[0]);
        CROW(4)  = mk4(rr1, rhs[1]);
        CROW(5)  = make_float4(M0[0][3], M0[0][4], M0[0][5], rhs[2]);
        CROW(6)  = make_float4(M0[1][3], M0[1][4], M0[1][5], A[0][0]);
        CROW(7)  = make_float4(M0[2][3], M0[2][4], M0[2][5], A[1][1]);
        CROW(8)  = make_float4(M1[0][3], M1[0][4], M1[0][5], A[2][2]);
        CROW(9)  = make_float4(M1[1][3], M1[1][4], M1[1][5], A[0][1]);
        CROW(10) = make_float4(M1[2][3], M1[2][4], M1[2][5], A[0][2]);
        CROW(11) = make_float4(A[1][0], A[1][2], A[2][0], A[2][1]);
        CROW(CROW_LAMBDA) = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    CROW(0)  = mk4(n, 1.0f / b0.mass);
    CROW(1)  = mk4(t, 1.0f / b1.mass);
    CROW(2)  = mk4(bb, __int_as_float(packed));
    CROW(3)  = make_float4(J0[0][3], J0[0][4], J0[0][5], rhs[0]);
    CROW(4)  = make_float4(J0[1][3], J0[1][4], J0[1][5], rhs[1]);
    CROW(5)  = make_float4(J0[2][3], J0[2][4], J0[2][5], rhs[2]);
    CROW(6)  = make_float4(J1[0][3], J1[0][4], J1[0][5], A[0][0]);
    CROW(7)  = make_float4(J1[1][3], J1[1][4], J1[1][5], A[1][1]);
    CROW(8)  = make_float4(J1[2][3], J1[2][4], J1[2][5], A[2][2]);
    CROW(9)  = make_float4(M0[0][3], M0[0][4], M0[0][5], A[0][1]);
    CROW(10) = make_float4(M0[1][3], M0[1][4], M0[1][5], A[0][2]);
    CROW(11) = make_float4(M0[2][3], M0[2][4], M0[2][5], A[1][0]);
    CROW(12) = make_float4(M1[0][3], M1[0][4], M1[0][5], A[1][2]);
    CROW(13) = make_float4(M1[1][3], M1[1][4], M1[1][5], A[2][0]);
    CROW(14) = make_float4(M1[2][3], M1[2][4], M1[2][5], A[2][1]);
    CROW(CROW_LAMBDA) = make_float4(0.f, 0.f, 0.f, 0.f);       // lambda = 0 (Contact.cpp:66, SolverBoxPGS.cpp:210)
#undef CROW
}

// thread / (contact slot, scene), scenes fastest: the warp-blocked tiles of mode 0 are written as full 512 B segments
__global__ void __launch_bounds__(128, CONTACT_ROWS_MIN_BLOCKS) k_contact_rows(View v, float h)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)v.NC * v.B) return;
    const int scene = (int)(tid % v.B), slot = (int)(tid / v.B);
    if (slot >= v.ncontacts[scene]) return;
    contact_row(v, h, scene, slot, contact_pos(v, slot, scene));
}

// level mode: thread / (position in the scene's level-ordered stream), positions fastest, so that the lanes of a warp
// write 32 consecutive records of one scene (the [scene][field][pos] layout makes that one 512 B segment per field;
// with scenes fastest every 16 B field of every row was a store of its own, 1 MB apart from its neighbours')
__global__ void __launch_bounds__(128, CONTACT_ROWS_MIN_BLOCKS) k_contact_rows_lv(View v, float h)
{
    const int scene = blockIdx.y, pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= v.ncontacts[scene]) return;
    contact_row(v, h, scene, v.cslot[(size_t)scene * v.NC + pos], pos);
}

// ------------------------------------------------------------------------------------
// K5  joint Jacobians, J M^-1, dim x dim diagonal block + LDL^T, RHS
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_joint_rows(View v, float h)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)v.NJ * v.B) return;
    const int scene = (int)(tid % v.B), slot = (int)(tid / v.B);
    if (slot >= v.njoints[scene]) return;
    const size_t j = at(v, slot, scene);
    const int4 def = v.jdef[j];                                  // type, body0, body1, dim
    const int dim = def.w;
    const BodyDyn b0 = load_body_dyn(v, def.y, scene);
    const BodyDyn b1 = load_body_dyn(v, def.z, scene);
    const Q4 qb0 = mkq(v.quat[at(v, def.y, scene)]);
    const Q4 qb1 = mkq(v.quat[at(v, def.z, scene)]);

    // Hinge::computeJacobian (Hinge.cpp:33-63) / Spherical::computeJacobian (Spherical.cpp:34-52)
    const V3 rr0 = rotate(qb0, mk3(v.jr0[j]));
    const V3 rr1 = rotate(qb1, mk3(v.jr1[j]));
    float phi[5] = { 0.f, 0.f, 0.f, 0.f, 0.f };
    const V3 e = ((b0.x + rr0) - b1.x) - rr1;
    phi[0] = e.x; phi[1] = e.y; phi[2] = e.z;
    float J0[5][6], J1[5][6], M0[5][6], M1[5][6];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int k = 0; k < 6; ++k) { J0[r][k] = 0.f; J1[r][k] = 0.f; }
    J0[0][0] = J0[1][1] = J0[2][2] = 1.f;
    // J1.block(0,0,3,3) = -sEye (Hinge.cpp:54): the off-diagonals are -0, kept so exported rows match to the bit
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) J1[r][k] = (r == k) ? -1.f : -0.f;
    set_hat(J0, -rr0);
    set_hat(J1, rr1);
    V3 nxu = mk3(0.f, 0.f, 0.f), nxv = mk3(0.f, 0.f, 0.f);
    if (dim == 5) {
        const V3 nn = rotate(qb0, rotate(mkq(v.jq0[j]), mk3(1.f, 0.f, 0.f)));
        const Q4 q1 = mkq(v.jq1[j]);
        const V3 uu = rotate(qb1, rotate(q1, mk3(0.f, 1.f, 0.f)));
        const V3 vv = rotate(qb1, rotate(q1, mk3(0.f, 0.f, 1.f)));
        nxu = cross(nn, uu); nxv = cross(nn, vv);
        phi[3] = dot(nn, uu); phi[4] = dot(nn, vv);
        set_row(J0[3], mk3(0.f, 0.f, 0.f), nxu); set_row(J0[4], mk3(0.f, 0.f, 0.f), nxv);
        set_row(J1[3], mk3(0.f, 0.f, 0.f), -nxu); set_row(J1[4], mk3(0.f, 0.f, 0.f), -nxv);
    }
#pragma unroll
    for (int r = 0; r < 5; ++r) { jminv_row(J0[r], M0[r], b0); jminv_row(J1[r], M1[r], b1); }

    // A = eps*I + sum_{non-fixed} JMinv J^T  (SolverBoxPGS.cpp:140-159); RHS (:27-44)
    float A[5][5], rhs[5];
    const float hinv = 1.0f / h;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            float a = (r == c) ? 1e-5f : 0.f;
            if (!b0.fixed) a += dot6_seq(M0[r], J0[c]);
            if (!b1.fixed) a += dot6_seq(M1[r], J1[c]);
            A[r][c] = a;
        }
        float b = (-hinv * 0.3f) * phi[r];
        if (!b0.fixed) { b = mult_and_sub_row(b, M0[r], b0.f, b0.tau, h); b = mult_and_sub_row(b, J0[r], b0.xdot, b0.omega, 1.0f); }
        if (!b1.fixed) { b = mult_and_sub_row(b, M1[r], b1.f, b1.tau, h); b = mult_and_sub_row(b, J1[r], b1.xdot, b1.omega, 1.0f); }
        rhs[r] = b;
    }
    if (dim < 5) {   // spherical: rows 3,4 are inert (lambda stays 0)
        for (int r = 3; r < 5; ++r) { for (int c = 0; c < 5; ++c) { A[r][c] = 0.f; A[c][r] = 0.f; } A[r][r] = 1.f; rhs[r] = 0.f; }
    }
    // LDL^T of the SPD block (Eigen::LDLT in the reference, SolverBoxPGS.cpp:161; the block is
    // J M^-1 J^T + 1e-5 I so no pivoting is needed for accuracy)
    float L[5][5], D[5], Dinv[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        float dsum = A[c][c];
#pragma unroll
        for (int k = 0; k < 5; ++k) if (k < c) dsum -= (L[c][k] * L[c][k]) * D[k];
        D[c] = dsum;
        Dinv[c] = 1.0f / dsum;
#pragma unroll
        for (int r = 0; r < 5; ++r) if (r > c) {
            float s = A[r][c];
#pragma unroll
            for (int k = 0; k < 5; ++k) if (k < c) s -= (L[r][k] * L[c][k]) * D[k];
            L[r][c] = s / dsum;
        }
    }
    const int id0 = def.y | (b0.fixed ? BODY_FIXED_BIT : 0);
    const int id1 = def.z | (b1.fixed ? BODY_FIXED_BIT : 0);
    float4* R = v.jrow;
    const int pos = joint_pos(v, slot, scene);
#define JROW(f) R[at_jrow(v, f, pos, scene)]
    JROW(0) = mk4(rr0, 1.0f / b0.mass);
    JROW(1) = mk4(rr1, 1.0f / b1.mass);
    JROW(2) = mk4(nxu, __int_as_float(id0));
    JROW(3) = mk4(nxv, __int_as_float(id1));
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        JROW(4 + r) = make_float4(M0[r][3], M0[r][4], M0[r][5], rhs[r]);
        JROW(9 + r) = make_float4(M1[r][3], M1[r][4], M1[r][5], Dinv[r]);
    }
    JROW(14) = make_float4(L[1][0], L[2][0], L[2][1], L[3][0]);
    JROW(15) = make_float4(L[3][1], L[3][2], L[4][0], L[4][1]);
    JROW(16) = make_float4(L[4][2], L[4][3], __int_as_float(dim), __int_as_float(slot));
#undef JROW
}

// ------------------------------------------------------------------------------------
// K6  Box-PGS sweep, thread-per-scene variant: rows in the reference's sequential order
// (row arithmetic in pgs_math.cuh)
// ------------------------------------------------------------------------------------
// One-entry write-back register cache of a body accumulator. Rows are visited in (i, j) order, so
// body0 of consecutive rows is almost always the same body (all contacts of body i, all hinges of a
// chassis): it stays in registers instead of making a dependent HBM round trip per row.
// L2 eviction-priority policies: the row stream is read once per sweep and must not push the body
// accumulators (re-read every few rows) out of the 126 MB L2
__device__ __forceinline__ unsigned long long l2_policy_evict_first()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ld_hint(const float4* p, unsigned long long pol)
{
    float4 r;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;\n"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol) : "memory");
    return r;
}
__device__ __forceinline__ void st_hint(float4* p, const float4& v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;\n"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
struct WCache {
    int id;          // cached body index, -1 = empty
    V3 l, a;
};
__device__ __forceinline__ void wc_flush(const View& v, int scene, WCache& c)
{
    if (c.id >= 0) {
        const size_t i = at(v, c.id, scene);
        const unsigned long long keep = l2_policy_evict_last();
        st_hint(&v.wl[i], mk4(c.l, 0.f), keep);
        st_hint(&v.wa[i], mk4(c.a, 0.f), keep);
    }
}
__device__ __forceinline__ void wc_acquire(const View& v, int scene, WCache& c, int id)
{
    if (c.id != id) {
        wc_flush(v, scene, c);
        const size_t i = at(v, id, scene);
        const unsigned long long keep = l2_policy_evict_last();
        c.l = mk3(ld_hint(&v.wl[i], keep)); c.a = mk3(ld_hint(&v.wa[i], keep));
        c.id = id;
    }
}

// cp.async (LDGSTS) 16-byte global -> shared copy and group bookkeeping
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16_hint(void* smem_dst, const void* gmem_src, unsigned long long pol)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// fetch the accumulators of a row's two bodies (body0 through the register cache), solve, write back
template <class Solve>
__device__ __forceinline__ void with_bodies(const View& v, int scene, WCache& wc, int id0, int id1, bool dyn0, bool dyn1, Solve solve)
{
    size_t w1i = 0;
    V3 w1l = mk3(0.f, 0.f, 0.f), w1a = w1l;
    const unsigned long long keep = l2_policy_evict_last();
    const bool hit1 = dyn1 && (id1 == wc.id) && !(dyn0 && id0 != wc.id);
    if (dyn0) wc_acquire(v, scene, wc, id0);
    if (dyn1) {
        if (hit1) { w1l = wc.l; w1a = wc.a; }
        else { w1i = at(v, id1, scene); w1l = mk3(ld_hint(&v.wl[w1i], keep)); w1a = mk3(ld_hint(&v.wa[w1i], keep)); }
    }
    if (hit1) {                                  // body0 is fixed here, the cache line is body1's
        V3 dl = mk3(0.f, 0.f, 0.f), da = dl;
        solve(dl, da, w1l, w1a);
        wc.l = w1l; wc.a = w1a;
    } else {
        solve(wc.l, wc.a, w1l, w1a);
        if (dyn1) { st_hint(&v.wl[w1i], mk4(w1l, 0.f), keep); st_hint(&v.wa[w1i], mk4(w1a, 0.f), keep); }
    }
}

__device__ __forceinline__ void contact_visit(const View& v, int slot, int scene, float mu, const float4* F, const float4& lam, WCache& wc)
{
    const ContactIds c = contact_ids(F);
    float4 out;
    with_bodies(v, scene, wc, c.id0, c.id1, c.dyn0, c.dyn1, [&](V3& w0l, V3& w0a, V3& w1l, V3& w1a) {
        out = contact_solve(F, lam, mu, c.dyn0, c.dyn1, w0l, w0a, w1l, w1a);
    });
    v.crow[at_crow(v, CROW_LAMBDA, slot, scene)] = out;
}

__device__ __forceinline__ void joint_visit(const View& v, int slot, int scene, WCache& wc)
{
    JointRec r;
    load_joint_rec(v, slot, scene, r);
    const size_t ji = at(v, slot, scene);
    const float4 lam03 = v.jlam[ji];
    float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, v.jlam4[ji] };
    const bool dyn0 = r.h.id0 >= 0, dyn1 = r.h.id1 >= 0;
    with_bodies(v, scene, wc, r.h.id0 & 0x7fffffff, r.h.id1 & 0x7fffffff, dyn0, dyn1, [&](V3& w0l, V3& w0a, V3& w1l, V3& w1a) {
        joint_solve(r, l, dyn0, dyn1, w0l, w0a, w1l, w1a);
    });
    v.jlam[ji] = make_float4(l[0], l[1], l[2], l[3]);
    v.jlam4[ji] = l[4];
}

__device__ __forceinline__ void add_force(const View& v, int body, int scene, const V3& fl, const V3& fa)
{
    const size_t i = at(v, body, scene);
    v.wl[i] = mk4(mk3(v.wl[i]) + fl, 0.f);
    v.wa[i] = mk4(mk3(v.wa[i]) + fa, 0.f);
}

#ifndef SLRBS_PGS_MIN_CTAS
#define SLRBS_PGS_MIN_CTAS 12
#endif
enum { PGS_THREADS = 32, PGS_STAGES = 2, PGS_MIN_CTAS = SLRBS_PGS_MIN_CTAS };

__global__ void __launch_bounds__(PGS_THREADS, PGS_MIN_CTAS) k_pgs(View v, float dt, int iters, float mu)
{
    // thread-private staging ring for the contact row stream: [stage][field][thread] float4
    __shared__ float4 ring[PGS_STAGES][CROW_FIELDS][PGS_THREADS];
    const int tid = threadIdx.x;
    const int scene = blockIdx.x * PGS_THREADS + tid;
    if (scene >= v.B) return;
    const int nb = v.nbodies[scene], nj = v.njoints[scene], nc = v.ncontacts[scene];
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < nb; ++i) { v.wl[at(v, i, scene)] = z4; v.wa[at(v, i, scene)] = z4; }
    // joints keep last step's lambda (Hinge.cpp:33-63 never zeroes it): seed w with it
    for (int s = 0; s < nj; ++s) {
        const JointRow j = load_joint_head(v, s, scene);
        const size_t ji = at(v, s, scene);
        const float4 lam03 = v.jlam[ji];
        const float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, v.jlam4[ji] };
        if (j.id0 >= 0) {
            const size_t wi = at(v, j.id0, scene);
            V3 wl = mk3(v.wl[wi]), wa = mk3(v.wa[wi]);
            joint_seed(j, l, 0, wl, wa);
            v.wl[wi] = mk4(wl, 0.f); v.wa[wi] = mk4(wa, 0.f);
        }
        if (j.id1 >= 0) {
            const size_t wi = at(v, j.id1 & 0x7fffffff, scene);
            V3 wl = mk3(v.wl[wi]), wa = mk3(v.wa[wi]);
            joint_seed(j, l, 1, wl, wa);
            v.wl[wi] = mk4(wl, 0.f); v.wa[wi] = mk4(wa, 0.f);
        }
    }
    WCache wc; wc.id = -1; wc.l = mk3(0.f, 0.f, 0.f); wc.a = wc.l;

    // The contact rows are static during the solve: stream them (and each row's lambda, which was
    // last written a whole sweep earlier) through shared memory with cp.async, PGS_STAGES - 1 rows
    // ahead of the row being solved, across sweep boundaries. With fewer rows than stages the
    // prefetched lambda could be stale, so it is then re-read directly.
    const int total = nc * iters;
    const bool lam_direct = nc < PGS_STAGES + 1;
    const unsigned long long pol_stream = l2_policy_evict_first();
    int fetch_row = 0, fetched = 0;
    auto fetch = [&]() {
        if (fetched < total) {
            const int st = fetched % PGS_STAGES;
            const float4* src = &v.crow[at_crow(v, 0, fetch_row, scene)];       // one contiguous 8 KB tile per warp
#pragma unroll
            for (int f = 0; f < CROW_FIELDS; ++f) cp_async16_hint(&ring[st][f][tid], src + f * 32, pol_stream);
            ++fetched;
            if (++fetch_row == nc) fetch_row = 0;
        }
        cp_async_commit();
    };
#pragma unroll
    for (int k = 0; k < PGS_STAGES - 1; ++k) fetch();
    int g = 0;
    for (int it = 0; it < iters; ++it) {                          // SolverBoxPGS.cpp:217-248
        for (int s = 0; s < nj; ++s) joint_visit(v, s, scene, wc);
        for (int s = 0; s < nc; ++s, ++g) {
            fetch();
            cp_async_wait<PGS_STAGES - 1>();
            const int st = g % PGS_STAGES;
            float4 F[CROW_FIELDS];
#pragma unroll
            for (int f = 0; f < CROW_FIELDS; ++f) F[f] = ring[st][f][tid];
            float4 lam = F[CROW_LAMBDA];
            if (lam_direct) lam = v.crow[at_crow(v, CROW_LAMBDA, s, scene)];
            contact_visit(v, s, scene, mu, F, lam, wc);
        }
    }
    cp_async_wait<0>();
    // constraint forces (RigidBodySystem.cpp:185-220): wl/wa become fc/tauc
    for (int i = 0; i < nb; ++i) { v.wl[at(v, i, scene)] = z4; v.wa[at(v, i, scene)] = z4; }
    for (int s = 0; s < nj; ++s) {
        const JointRow j = load_joint_head(v, s, scene);
        const size_t ji = at(v, s, scene);
        const float4 lam03 = v.jlam[ji];
        const float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, v.jlam4[ji] };
        const int dim = __float_as_int(__ldg(&v.jrow[at_jrow(v, 16, s, scene)]).z);
        float f0[6], f1[6];
        joint_force(j, dim, l, dt, f0, f1);
        add_force(v, j.id0 & 0x7fffffff, scene, mk3(f0[0], f0[1], f0[2]), mk3(f0[3], f0[4], f0[5]));   // unconditional (:189-192)
        add_force(v, j.id1 & 0x7fffffff, scene, mk3(f1[0], f1[1], f1[2]), mk3(f1[3], f1[4], f1[5]));
    }
    for (int s = 0; s < nc; ++s) {
        float4 F[9];
#pragma unroll
        for (int f = 0; f < 9; ++f) F[f] = __ldg(&v.crow[at_crow(v, f, s, scene)]);
        const float4 lam = v.crow[at_crow(v, CROW_LAMBDA, s, scene)];
        const ContactIds c = contact_ids(F);
        V3 fl, fa;
        if (c.dyn0) { contact_force(F, lam, dt, 0, fl, fa); add_force(v, c.id0, scene, fl, fa); }
        if (c.dyn1) { contact_force(F, lam, dt, 1, fl, fa); add_force(v, c.id1, scene, fl, fa); }
    }
    if (v.profiling) {
        atomicAdd(&v.visit_counters[0], (unsigned long long)nc * iters);
        atomicAdd(&v.visit_counters[1], (unsigned long long)nj * iters);
    }
}

// ------------------------------------------------------------------------------------
// K6 (TMA variant): the same sweep, but the warp's 8 KB row tile is fetched by ONE bulk asynchronous
// copy (cp.async.bulk global -> shared, completion on an mbarrier) issued by an elected lane instead of
// 16 x 32 per-lane cp.async: the tile layout [field][lane] in HBM is exactly the ring stage layout.
// The loop is warp-uniform (trip count = the largest contact count of the 32 scenes; lanes past their
// own count idle), lambda written through the generic proxy is fenced before the async proxy re-reads it.
// MEASURED (marble box x 32768): 13.8 ms per launch against 10.7 ms for the per-lane cp.async ring
// (11.8 ms without the per-visit proxy fence): the warp-wide barrier per row and the cross-proxy fence
// cost more than the saved LDGSTS issue slots, so this variant is opt-in (SLRBS_PGS_TMA=1) only.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

enum { TMA_STAGES = 3, TILE_BYTES = CROW_FIELDS * 32 * 16 };

__global__ void __launch_bounds__(32, 9) k_pgs_tma(View v, float dt, int iters, float mu)
{
    __shared__ __align__(128) float4 ring[TMA_STAGES][CROW_FIELDS][32];
    __shared__ __align__(8) unsigned long long bars[TMA_STAGES];
    const int lane = threadIdx.x;
    const int scene = blockIdx.x * 32 + lane;
    const bool live = scene < v.B;
    const int sc = live ? scene : v.B - 1;                       // dead lanes shadow the last scene, read-only
    const int nb = live ? v.nbodies[sc] : 0, nj = live ? v.njoints[sc] : 0, nc = live ? v.ncontacts[sc] : 0;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < TMA_STAGES; ++k) mbar_init(&bars[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < nb; ++i) { v.wl[at(v, i, sc)] = z4; v.wa[at(v, i, sc)] = z4; }
    for (int s = 0; s < nj; ++s) {                                // seed w with last step's joint impulses
        const JointRow j = load_joint_head(v, s, sc);
        const size_t ji = at(v, s, sc);
        const float4 lam03 = v.jlam[ji];
        const float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, v.jlam4[ji] };
        if (j.id0 >= 0) {
            const size_t wi = at(v, j.id0, sc);
            V3 wl = mk3(v.wl[wi]), wa = mk3(v.wa[wi]);
            joint_seed(j, l, 0, wl, wa);
            v.wl[wi] = mk4(wl, 0.f); v.wa[wi] = mk4(wa, 0.f);
        }
        if (j.id1 >= 0) {
            const size_t wi = at(v, j.id1 & 0x7fffffff, sc);
            V3 wl = mk3(v.wl[wi]), wa = mk3(v.wa[wi]);
            joint_seed(j, l, 1, wl, wa);
            v.wl[wi] = mk4(wl, 0.f); v.wa[wi] = mk4(wa, 0.f);
        }
    }
    WCache wc; wc.id = -1; wc.l = mk3(0.f, 0.f, 0.f); wc.a = wc.l;

    int ncmax = nc, njmax = nj;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        ncmax = max(ncmax, __shfl_xor_sync(0xffffffffu, ncmax, d));
        njmax = max(njmax, __shfl_xor_sync(0xffffffffu, njmax, d));
    }
    const int total = ncmax * iters;
    const bool lam_direct = ncmax < TMA_STAGES + 1;
    const float4* tiles = &v.crow[at_crow(v, 0, 0, blockIdx.x * 32)];       // tile of row r: tiles + r * CROW_FIELDS * 32
    auto fetch = [&](int k) {                                     // warp-uniform
        if (k < total && lane == 0) {
            const int st = k % TMA_STAGES, row = k % ncmax;
            fence_proxy_async();
            mbar_expect_tx(&bars[st], TILE_BYTES);
            bulk_g2s(&ring[st][0][0], tiles + (size_t)row * CROW_FIELDS * 32, TILE_BYTES, &bars[st]);
        }
    };
    for (int k = 0; k < TMA_STAGES - 1; ++k) fetch(k);
    int g = 0;
    for (int it = 0; it < iters; ++it) {                          // SolverBoxPGS.cpp:217-248
        for (int s = 0; s < njmax; ++s) if (s < nj) joint_visit(v, s, sc, wc);
        for (int s = 0; s < ncmax; ++s, ++g) {
            __syncwarp();                                         // everyone is done reading the stage about to be refilled
            fetch(g + TMA_STAGES - 1);
            const int st = g % TMA_STAGES;
            mbar_wait(&bars[st], (unsigned)(g / TMA_STAGES) & 1u);
            if (s < nc) {
                float4 F[CROW_FIELDS];
#pragma unroll
                for (int f = 0; f < CROW_FIELDS; ++f) F[f] = ring[st][f][lane];
                float4 lam = F[CROW_LAMBDA];
                if (lam_direct) lam = v.crow[at_crow(v, CROW_LAMBDA, s, sc)];
                contact_visit(v, s, sc, mu, F, lam, wc);
                fence_proxy_async();                              // lambda must be visible to the next bulk read of this tile
            }
        }
    }
    // constraint forces (RigidBodySystem.cpp:185-220): wl/wa become fc/tauc
    for (int i = 0; i < nb; ++i) { v.wl[at(v, i, sc)] = z4; v.wa[at(v, i, sc)] = z4; }
    for (int s = 0; s < nj; ++s) {
        const JointRow j = load_joint_head(v, s, sc);
        const size_t ji = at(v, s, sc);
        const float4 lam03 = v.jlam[ji];
        const float l[5] = { lam03.x, lam03.y, lam03.z, lam03.w, v.jlam4[ji] };
        const int dim = __float_as_int(__ldg(&v.jrow[at_jrow(v, 16, s, sc)]).z);
        float f0[6], f1[6];
        joint_force(j, dim, l, dt, f0, f1);
        add_force(v, j.id0 & 0x7fffffff, sc, mk3(f0[0], f0[1], f0[2]), mk3(f0[3], f0[4], f0[5]));
        add_force(v, j.id1 & 0x7fffffff, sc, mk3(f1[0], f1[1], f1[2]), mk3(f1[3], f1[4], f1[5]));
    }
    for (int s = 0; s < nc; ++s) {
        float4 F[9];
#pragma unroll
        for (int f = 0; f < 9; ++f) F[f] = __ldg(&v.crow[at_crow(v, f, s, sc)]);
        const float4 lam = v.crow[at_crow(v, CROW_LAMBDA, s, sc)];
        const ContactIds c = contact_ids(F);
        V3 fl, fa;
        if (c.dyn0) { contact_force(F, lam, dt, 0, fl, fa); add_force(v, c.id0, sc, fl, fa); }
        if (c.dyn1) { contact_force(F, lam, dt, 1, fl, fa); add_force(v, c.id1, sc, fl, fa); }
    }
    if (v.profiling && live) {
        atomicAdd(&v.visit_counters[0], (unsigned long long)nc * iters);
        atomicAdd(&v.visit_counters[1], (unsigned long long)nj * iters);
    }
}

// ------------------------------------------------------------------------------------
// K6c  joint-only Krylov solvers (solverId 1 = CG, 2 = CR): SolverConjGradient.cpp:109-162,
// SolverConjResidual.cpp:100-158. One thread per scene; contacts are ignored (lambda stays 0).
// A x is evaluated body-wise: u_body = sum_{joints on body} J^T x, then (A x)_j = J0Minv_j u_b0 +
// J1Minv_j u_b1 over the non-fixed bodies, which is the reference's self + neighbour sum
// (SolverConjGradient.cpp:49-100) regrouped. The reference's quirks are kept: CG updates
// p = r - beta p (:150), CR adds 1e-9 x to A x (SolverConjResidual.cpp:71,78), both stop at 1e-12.
// ------------------------------------------------------------------------------------
enum { KV_X = 0, KV_R = 1, KV_P = 2, KV_B = 3, KV_AP = 4, KV_AR = 5, KV_AX = 6 };

__device__ __forceinline__ void kry_load(const View& v, int vec, int s, int scene, float* x)
{
#pragma unroll
    for (int k = 0; k < 5; ++k) x[k] = v.kry[at_kry(v, vec, k, s, scene)];
}
__device__ __forceinline__ void kry_store(const View& v, int vec, int s, int scene, const float* x)
{
#pragma unroll
    for (int k = 0; k < 5; ++k) v.kry[at_kry(v, vec, k, s, scene)] = x[k];
}

// out = A * in   (vectors live in v.kry); wl/wa are used as the per-body u accumulators
__device__ void kry_apply(const View& v, int scene, int nb, int nj, int vin, int vout, float eps)
{
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < nb; ++i) { v.wl[at(v, i, scene)] = z4; v.wa[at(v, i, scene)] = z4; }
    for (int s = 0; s < nj; ++s) {
        const JointRow j = load_joint_head(v, joint_pos(v, s, scene), scene);
        float x[5]; kry_load(v, vin, s, scene, x);
        const V3 ll = mk3(x[0], x[1], x[2]);
        if (j.id0 >= 0) {
            const size_t wi = at(v, j.id0, scene);
            v.wl[wi] = mk4(mk3(v.wl[wi]) + ll, 0.f);
            v.wa[wi] = mk4(mk3(v.wa[wi]) + joint_jt_ang(j.rr0, j.nxu, j.nxv, x), 0.f);
        }
        if (j.id1 >= 0) {
            const size_t wi = at(v, j.id1 & 0x7fffffff, scene);
            v.wl[wi] = mk4(mk3(v.wl[wi]) - ll, 0.f);
            v.wa[wi] = mk4(mk3(v.wa[wi]) - joint_jt_ang(j.rr1, j.nxu, j.nxv, x), 0.f);
        }
    }
    for (int s = 0; s < nj; ++s) {
        const int jp = joint_pos(v, s, scene);
        const JointRow j = load_joint_head(v, jp, scene);
        const int dim = __float_as_int(__ldg(&v.jrow[at_jrow(v, 16, jp, scene)]).z);
        float x[5]; kry_load(v, vin, s, scene, x);
        float y[5];
#pragma unroll
        for (int r = 0; r < 5; ++r) y[r] = eps * x[r];
        if (j.id0 >= 0) {
            const size_t wi = at(v, j.id0, scene);
            const V3 ul = mk3(v.wl[wi]), ua = mk3(v.wa[wi]);
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                const V3 m = mk3(__ldg(&v.jrow[at_jrow(v, 4 + r, jp, scene)]));
                y[r] += (r < 3 ? j.im0 * (r == 0 ? ul.x : (r == 1 ? ul.y : ul.z)) : 0.f) + dotf(m, ua);
            }
        }
        if (j.id1 >= 0) {
            const size_t wi = at(v, j.id1 & 0x7fffffff, scene);
            const V3 ul = mk3(v.wl[wi]), ua = mk3(v.wa[wi]);
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                const V3 m = mk3(__ldg(&v.jrow[at_jrow(v, 9 + r, jp, scene)]));
                y[r] += (r < 3 ? -j.im1 * (r == 0 ? ul.x : (r == 1 ? ul.y : ul.z)) : 0.f) + dotf(m, ua);
            }
        }
#pragma unroll
        for (int r = 0; r < 5; ++r) if (r >= dim) y[r] = 0.f;
        kry_store(v, vout, s, scene, y);
    }
}

__device__ float kry_dot(const View& v, int scene, int nj, int va, int vb)
{
    float acc = 0.f;
    for (int s = 0; s < nj; ++s) {
        float a[5], b[5]; kry_load(v, va, s, scene, a); kry_load(v, vb, s, scene, b);
#pragma unroll
        for (int k = 0; k < 5; ++k) acc += a[k] * b[k];
    }
    return acc;
}

template <bool CR>
__global__ void __launch_bounds__(64) k_krylov(View v, int iters)
{
    const int scene = blockIdx.x * blockDim.x + threadIdx.x;
    if (scene >= v.B) return;
    const int nb = v.nbodies[scene], nj = v.njoints[scene];
    if (nj == 0) return;
    const float eps = CR ? 1e-9f : 0.f;
    // x = 0, b = rhs from the row records, r = b - A x = b, p = r
    for (int s = 0; s < nj; ++s) {
        float b[5], z[5] = { 0.f, 0.f, 0.f, 0.f, 0.f };
#pragma unroll
        for (int r = 0; r < 5; ++r) b[r] = __ldg(&v.jrow[at_jrow(v, 4 + r, joint_pos(v, s, scene), scene)]).w;
        kry_store(v, KV_X, s, scene, z); kry_store(v, KV_B, s, scene, b);
        kry_store(v, KV_R, s, scene, b); kry_store(v, KV_P, s, scene, b);
    }
    kry_apply(v, scene, nb, nj, KV_P, KV_AP, eps);
    if (!CR) {
        float rTr = kry_dot(v, scene, nj, KV_R, KV_R);
        float pAp = kry_dot(v, scene, nj, KV_P, KV_AP);
        for (int it = 0; it < iters; ++it) {
            if (rTr < 1e-12f) break;
            if (pAp < 1e-12f) break;
            const float alpha = rTr / pAp;
            for (int s = 0; s < nj; ++s) {
                float x[5], r[5], p[5], ap[5];
                kry_load(v, KV_X, s, scene, x); kry_load(v, KV_R, s, scene, r); kry_load(v, KV_P, s, scene, p); kry_load(v, KV_AP, s, scene, ap);
#pragma unroll
                for (int k = 0; k < 5; ++k) { x[k] += alpha * p[k]; r[k] -= alpha * ap[k]; }
                kry_store(v, KV_X, s, scene, x); kry_store(v, KV_R, s, scene, r);
            }
            const float rTr_next = kry_dot(v, scene, nj, KV_R, KV_R);
            const float beta = rTr_next / rTr;
            for (int s = 0; s < nj; ++s) {
                float r[5], p[5];
                kry_load(v, KV_R, s, scene, r); kry_load(v, KV_P, s, scene, p);
#pragma unroll
                for (int k = 0; k < 5; ++k) p[k] = r[k] - beta * p[k];          // sic (SolverConjGradient.cpp:150)
                kry_store(v, KV_P, s, scene, p);
            }
            kry_apply(v, scene, nb, nj, KV_P, KV_AP, eps);
            pAp = kry_dot(v, scene, nj, KV_P, KV_AP);
            rTr = rTr_next;
        }
    } else {
        kry_apply(v, scene, nb, nj, KV_R, KV_AR, eps);
        float rAr = kry_dot(v, scene, nj, KV_R, KV_AR);
        float pATAp = kry_dot(v, scene, nj, KV_AP, KV_AP);
        for (int it = 0; it < iters; ++it) {
            if (rAr < 1e-12f) break;
            if (pATAp < 1e-12f) break;
            const float alpha = rAr / pATAp;
            for (int s = 0; s < nj; ++s) {
                float x[5], r[5], p[5], ap[5];
                kry_load(v, KV_X, s, scene, x); kry_load(v, KV_R, s, scene, r); kry_load(v, KV_P, s, scene, p); kry_load(v, KV_AP, s, scene, ap);
#pragma unroll
                for (int k = 0; k < 5; ++k) { x[k] += alpha * p[k]; r[k] -= alpha * ap[k]; }
                kry_store(v, KV_X, s, scene, x); kry_store(v, KV_R, s, scene, r);
            }
            kry_apply(v, scene, nb, nj, KV_R, KV_AR, eps);
            const float rAr_next = kry_dot(v, scene, nj, KV_R, KV_AR);
            const float beta = rAr_next / rAr;
            for (int s = 0; s < nj; ++s) {
                float r[5], p[5], ap[5], ar[5];
                kry_load(v, KV_R, s, scene, r); kry_load(v, KV_P, s, scene, p); kry_load(v, KV_AP, s, scene, ap); kry_load(v, KV_AR, s, scene, ar);
#pragma unroll
                for (int k = 0; k < 5; ++k) { p[k] = r[k] + beta * p[k]; ap[k] = ar[k] + beta * ap[k]; }
                kry_store(v, KV_P, s, scene, p); kry_store(v, KV_AP, s, scene, ap);
            }
            rAr = rAr_next;
            pATAp = kry_dot(v, scene, nj, KV_AP, KV_AP);
        }
    }
    for (int s = 0; s < nj; ++s) {                               // j->lambda = x.segment(j->idx, j->dim)
        float x[5]; kry_load(v, KV_X, s, scene, x);
        const size_t ji = at(v, s, scene);
        v.jlam[ji] = make_float4(x[0], x[1], x[2], x[3]);
        v.jlam4[ji] = x[4];
    }
}

// ------------------------------------------------------------------------------------
// K7  symplectic Euler + quaternion update (RigidBodySystem.cpp:99-120)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_integrate(View v, float dt)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)v.NB * v.B) return;
    const int scene = (int)(tid % v.B), slot = (int)(tid / v.B);
    if (slot >= v.nbodies[scene]) return;
    const size_t i = at(v, slot, scene);
    const int fl = v.flags[i];
    float4 vv = v.vel[i], ww = v.omg[i];
    if (!(fl & FLAG_FIXED)) {
        float4 px = v.pos[i];
        const float mass = px.w;
        const V3 f = mk3(v.force[i]), tau = mk3(v.torque[i]);
        const V3 fc = mk3(v.wl[i]), tauc = mk3(v.wa[i]);
        V3 xdot = mk3(vv), omega = mk3(ww), x = mk3(px);
        Q4 q = mkq(v.quat[i]);
        const M3 I = load_m3(v.Iw, v, slot, scene);
        const M3 Iinv = load_m3(v.Iinvw, v, slot, scene);
        xdot = xdot + (dt * (1.0f / mass)) * (f + fc);
        omega = omega + dt * matvec(Iinv, (tau + tauc) - cross(omega, matvec(I, omega)));
        x = x + dt * xdot;
        const float hdt = 0.5f * dt;
        Q4 wq; wq.w = hdt * 0.0f; wq.x = hdt * omega.x; wq.y = hdt * omega.y; wq.z = hdt * omega.z;
        const Q4 dq = qmul(wq, q);
        Q4 qn; qn.x = q.x + dq.x; qn.y = q.y + dq.y; qn.z = q.z + dq.z; qn.w = q.w + dq.w;
        qn = qnormalized(qn);
        v.pos[i] = make_float4(x.x, x.y, x.z, mass);
        v.quat[i] = make_float4(qn.x, qn.y, qn.z, qn.w);
        v.vel[i] = make_float4(xdot.x, xdot.y, xdot.z, vv.w);
        v.omg[i] = make_float4(omega.x, omega.y, omega.z, ww.w);
    } else {
        v.vel[i] = make_float4(0.f, 0.f, 0.f, vv.w);
        v.omg[i] = make_float4(0.f, 0.f, 0.f, ww.w);
    }
}

// ------------------------------------------------------------------------------------
// host <-> device repacking (scene-major host arrays <-> slot-major device arrays)
// ------------------------------------------------------------------------------------
__global__ void k_upload_bodies(View v, int scene0, int ns, int n, const int* counts, BodyStage st)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)ns * n) return;
    const int ls = (int)(tid % ns), slot = (int)(tid / ns);
    const int scene = scene0 + ls;
    if (slot == 0) v.nbodies[scene] = counts ? counts[ls] : n;
    const size_t h = (size_t)ls * n + slot;
    const size_t i = at(v, slot, scene);
    float4 px = v.pos[i];
    if (st.x) { px.x = st.x[3 * h]; px.y = st.x[3 * h + 1]; px.z = st.x[3 * h + 2]; }
    if (st.mass) px.w = st.mass[h];
    v.pos[i] = px;
    if (st.q) v.quat[i] = make_float4(st.q[4 * h], st.q[4 * h + 1], st.q[4 * h + 2], st.q[4 * h + 3]);
    if (st.v) v.vel[i] = make_float4(st.v[3 * h], st.v[3 * h + 1], st.v[3 * h + 2], 0.f);
    if (st.w) v.omg[i] = make_float4(st.w[3 * h], st.w[3 * h + 1], st.w[3 * h + 2], 0.f);
    if (st.geom_type) v.flags[i] = (st.geom_type[h] & FLAG_TYPE_MASK) | ((st.fixed && st.fixed[h]) ? FLAG_FIXED : 0);
    if (st.geom) {
        const float* g = st.geom + 6 * h;
        v.geomA[i] = make_float4(g[0], g[1], g[2], 0.f);
        v.geomB[i] = make_float4(g[3], g[4], g[5], 0.f);
    }
    if (st.ibody) for (int k = 0; k < 9; ++k) v.ibody[at_b9(v, k, slot, scene)] = st.ibody[9 * h + k];
    if (st.ibodyinv) for (int k = 0; k < 9; ++k) v.ibodyinv[at_b9(v, k, slot, scene)] = st.ibodyinv[9 * h + k];
}

__global__ void k_download_state(View v, int scene0, int ns, int n, BodyStage st)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)ns * n) return;
    const int ls = (int)(tid % ns), slot = (int)(tid / ns);
    const int scene = scene0 + ls;
    const size_t h = (size_t)ls * n + slot;
    const size_t i = at(v, slot, scene);
    float* x = (float*)st.x; float* q = (float*)st.q; float* vv = (float*)st.v; float* w = (float*)st.w;
    if (x) { const float4 a = v.pos[i]; x[3 * h] = a.x; x[3 * h + 1] = a.y; x[3 * h + 2] = a.z; }
    if (q) { const float4 a = v.quat[i]; q[4 * h] = a.x; q[4 * h + 1] = a.y; q[4 * h + 2] = a.z; q[4 * h + 3] = a.w; }
    if (vv) { const float4 a = v.vel[i]; vv[3 * h] = a.x; vv[3 * h + 1] = a.y; vv[3 * h + 2] = a.z; }
    if (w) { const float4 a = v.omg[i]; w[3 * h] = a.x; w[3 * h + 1] = a.y; w[3 * h + 2] = a.z; }
}

__global__ void k_upload_joints(View v, int scene0, int ns, int n, const int* counts, JointStage st)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)ns * n) return;
    const int ls = (int)(tid % ns), slot = (int)(tid / ns);
    const int scene = scene0 + ls;
    if (slot == 0) v.njoints[scene] = counts ? counts[ls] : n;
    const size_t h = (size_t)ls * n + slot;
    const size_t j = at(v, slot, scene);
    const int type = st.type[h];
    v.jdef[j] = make_int4(type, st.body0[h], st.body1[h], type == 2 ? 5 : 3);
    v.jr0[j] = make_float4(st.r0[3 * h], st.r0[3 * h + 1], st.r0[3 * h + 2], 0.f);
    v.jr1[j] = make_float4(st.r1[3 * h], st.r1[3 * h + 1], st.r1[3 * h + 2], 0.f);
    v.jq0[j] = st.q0 ? make_float4(st.q0[4 * h], st.q0[4 * h + 1], st.q0[4 * h + 2], st.q0[4 * h + 3]) : make_float4(0.f, 0.f, 0.f, 1.f);
    v.jq1[j] = st.q1 ? make_float4(st.q1[4 * h], st.q1[4 * h + 1], st.q1[4 * h + 2], st.q1[4 * h + 3]) : make_float4(0.f, 0.f, 0.f, 1.f);
    v.jlam[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    v.jlam4[j] = 0.f;
}

__global__ void k_clear_joint_lambda(View v, int scene0, int ns)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)ns * v.NJ) return;
    const size_t j = at(v, (int)(tid / ns), scene0 + (int)(tid % ns));
    v.jlam[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    v.jlam4[j] = 0.f;
}

__global__ void k_joint_lambda(View v, int scene0, int ns, int n, float* lam5, int upload)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)ns * n) return;
    const int ls = (int)(tid % ns), slot = (int)(tid / ns);
    const size_t h = (size_t)ls * n + slot;
    const size_t j = at(v, slot, scene0 + ls);
    if (upload) {
        v.jlam[j] = make_float4(lam5[5 * h], lam5[5 * h + 1], lam5[5 * h + 2], lam5[5 * h + 3]);
        v.jlam4[j] = lam5[5 * h + 4];
    } else {
        const float4 a = v.jlam[j];
        lam5[5 * h] = a.x; lam5[5 * h + 1] = a.y; lam5[5 * h + 2] = a.z; lam5[5 * h + 3] = a.w; lam5[5 * h + 4] = v.jlam4[j];
    }
}

__global__ void k_set_ext(View v, int scene0, int ns, int n, const float* f3, const float* t3, int mode)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)ns * n) return;
    const int ls = (int)(tid % ns), slot = (int)(tid / ns);
    if (slot == 0) v.ext_mode[scene0 + ls] = mode;
    const size_t h = (size_t)ls * n + slot;
    const size_t i = at(v, slot, scene0 + ls);
    v.extf[i] = f3 ? make_float4(f3[3 * h], f3[3 * h + 1], f3[3 * h + 2], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    v.extt[i] = t3 ? make_float4(t3[3 * h], t3[3 * h + 1], t3[3 * h + 2], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void k_snapshot(View v, int scene0, int ns, int restore)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)ns * v.NB) return;
    const int ls = (int)(tid % ns), slot = (int)(tid / ns);
    const int scene = scene0 + ls;
    if (restore && slot == 0) { v.ncontacts[scene] = 0; if (v.nclev) v.nclev[scene] = 0; }   // RigidBodyState.cpp:71
    if (slot >= v.nbodies[scene]) return;
    const size_t i = at(v, slot, scene);
    if (restore) {
        const float4 sp = v.s_pos[i];
        const float4 p = v.pos[i];
        v.pos[i] = make_float4(sp.x, sp.y, sp.z, p.w);
        v.quat[i] = v.s_quat[i]; v.vel[i] = v.s_vel[i]; v.omg[i] = v.s_omg[i];
        v.force[i] = make_float4(0.f, 0.f, 0.f, 0.f);            // RigidBodyState.cpp:35-36
        v.wl[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        v.s_pos[i] = v.pos[i]; v.s_quat[i] = v.quat[i]; v.s_vel[i] = v.vel[i]; v.s_omg[i] = v.omg[i];
    }
}

// halo refresh for a scene cut into overlapping blocks: indexed state copies (flat index = scene * NB + slot)
__device__ __forceinline__ size_t flat_to_at(const View& v, int flat) { return at(v, flat % v.NB, flat / v.NB); }

__global__ void k_halo_copy(View v, int n, const int* __restrict__ src, const int* __restrict__ dst)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const size_t a = flat_to_at(v, src[k]), b = flat_to_at(v, dst[k]);
    const float4 p = v.pos[a], pd = v.pos[b];
    v.pos[b] = make_float4(p.x, p.y, p.z, pd.w);                 // keep the copy's own mass field
    v.quat[b] = v.quat[a]; v.vel[b] = v.vel[a]; v.omg[b] = v.omg[a];
}

__global__ void k_halo_pack(View v, int n, const int* __restrict__ idx, float* __restrict__ buf)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const size_t a = flat_to_at(v, idx[k]);
    const float4 p = v.pos[a], q = v.quat[a], l = v.vel[a], w = v.omg[a];
    float* o = buf + (size_t)13 * k;
    o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = q.x; o[4] = q.y; o[5] = q.z; o[6] = q.w;
    o[7] = l.x; o[8] = l.y; o[9] = l.z; o[10] = w.x; o[11] = w.y; o[12] = w.z;
}

__global__ void k_halo_unpack(View v, int n, const int* __restrict__ idx, const float* __restrict__ buf)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const size_t b = flat_to_at(v, idx[k]);
    const float* o = buf + (size_t)13 * k;
    v.pos[b] = make_float4(o[0], o[1], o[2], v.pos[b].w);
    v.quat[b] = make_float4(o[3], o[4], o[5], o[6]);
    v.vel[b] = make_float4(o[7], o[8], o[9], 0.f);
    v.omg[b] = make_float4(o[10], o[11], o[12], 0.f);
}

// dense exports of one scene for parity tests / getContacts()
__global__ void k_export_contacts(View v, int scene, int n, int* body0, int* body1, float* p3, float* n3,
                                  float* t3, float* b3, float* phi, float* lam3, float* J0, float* J1,
                                  float* M0, float* M1, float* A9, float* rhs3)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const size_t ci = atc(v, c, scene);
    const int2 ids = v.cids[ci];
    if (body0) body0[c] = ids.x;
    if (body1) body1[c] = ids.y;
    const float4 pp = v.cp[ci], nn = v.cn[ci];
    if (p3) { p3[3 * c] = pp.x; p3[3 * c + 1] = pp.y; p3[3 * c + 2] = pp.z; }
    if (n3) { n3[3 * c] = nn.x; n3[3 * c + 1] = nn.y; n3[3 * c + 2] = nn.z; }
    if (phi) phi[c] = pp.w;
    float4 F[CROW_FIELDS];
    const int cp_ = contact_pos(v, c, scene);
    load_crow_fields(v, cp_, scene, F);
    F[CROW_LAMBDA] = v.crow[at_crow(v, CROW_LAMBDA, cp_, scene)];
    if (t3) { t3[3 * c] = F[1].x; t3[3 * c + 1] = F[1].y; t3[3 * c + 2] = F[1].z; }
    if (b3) { b3[3 * c] = F[2].x; b3[3 * c + 1] = F[2].y; b3[3 * c + 2] = F[2].z; }
    if (lam3) { const float4 l = F[CROW_LAMBDA]; lam3[3 * c] = l.x; lam3[3 * c + 1] = l.y; lam3[3 * c + 2] = l.z; }
    const float im0 = F[0].w, im1 = F[1].w;
    for (int r = 0; r < 3; ++r) {
        const float4 d = F[r], a0 = F[3 + r], a1 = F[6 + r], m0 = F[9 + r], m1 = F[12 + r];
        const int o = 18 * c + 6 * r;
        if (J0) { J0[o] = d.x; J0[o + 1] = d.y; J0[o + 2] = d.z; J0[o + 3] = a0.x; J0[o + 4] = a0.y; J0[o + 5] = a0.z; }
        if (J1) { J1[o] = -d.x; J1[o + 1] = -d.y; J1[o + 2] = -d.z; J1[o + 3] = a1.x; J1[o + 4] = a1.y; J1[o + 5] = a1.z; }
        if (M0) { M0[o] = im0 * d.x; M0[o + 1] = im0 * d.y; M0[o + 2] = im0 * d.z; M0[o + 3] = m0.x; M0[o + 4] = m0.y; M0[o + 5] = m0.z; }
        if (M1) { M1[o] = im1 * -d.x; M1[o + 1] = im1 * -d.y; M1[o + 2] = im1 * -d.z; M1[o + 3] = m1.x; M1[o + 4] = m1.y; M1[o + 5] = m1.z; }
    }
    if (A9) {
        float* A = A9 + 9 * c;
        A[0] = F[6].w; A[1] = F[9].w; A[2] = F[10].w; A[3] = F[11].w; A[4] = F[7].w; A[5] = F[12].w;
        A[6] = F[13].w; A[7] = F[14].w; A[8] = F[8].w;
    }
    if (rhs3) { rhs3[3 * c] = F[3].w; rhs3[3 * c + 1] = F[4].w; rhs3[3 * c + 2] = F[5].w; }
}

// Contact::lambda written by a host-side Solver subclass (the Solver seam, include/solvers/Solver.h:9-43)
__global__ void k_import_contact_lambda(View v, int scene, int n, const float* lam3)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    v.crow[at_crow(v, CROW_LAMBDA, contact_pos(v, c, scene), scene)] = make_float4(lam3[3 * c], lam3[3 * c + 1], lam3[3 * c + 2], 0.f);
}

__global__ void k_export_joints(View v, int scene, int n, float* J0, float* J1, float* M0, float* M1, float* rhs5)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    float4 G[JROW_FIELDS];
    const int jp_ = joint_pos(v, s, scene);
    for (int f = 0; f < JROW_FIELDS; ++f) G[f] = v.jrow[at_jrow(v, f, jp_, scene)];
    const V3 rr0 = mk3(G[0]), rr1 = mk3(G[1]), nxu = mk3(G[2]), nxv = mk3(G[3]);
    const float im0 = G[0].w, im1 = G[1].w;
    const int dim = __float_as_int(G[16].z);
    float A0[5][6], A1[5][6];
    for (int r = 0; r < 5; ++r) for (int k = 0; k < 6; ++k) { A0[r][k] = 0.f; A1[r][k] = 0.f; }
    A0[0][0] = A0[1][1] = A0[2][2] = 1.f;
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) A1[r][k] = (r == k) ? -1.f : -0.f;   // -sEye
    set_hat(A0, -rr0); set_hat(A1, rr1);
    if (dim == 5) {
        set_row(A0[3], mk3(0.f, 0.f, 0.f), nxu); set_row(A0[4], mk3(0.f, 0.f, 0.f), nxv);
        set_row(A1[3], mk3(0.f, 0.f, 0.f), -nxu); set_row(A1[4], mk3(0.f, 0.f, 0.f), -nxv);
    }
    for (int r = 0; r < 5; ++r) {
        const bool live = r < dim;
        for (int k = 0; k < 6; ++k) {
            const int o = 30 * s + 6 * r + k;
            if (J0) J0[o] = live ? A0[r][k] : 0.f;
            if (J1) J1[o] = live ? A1[r][k] : 0.f;
            if (M0) M0[o] = !live ? 0.f : (k < 3 ? im0 * A0[r][k] : (k == 3 ? G[4 + r].x : (k == 4 ? G[4 + r].y : G[4 + r].z)));
            if (M1) M1[o] = !live ? 0.f : (k < 3 ? im1 * A1[r][k] : (k == 3 ? G[9 + r].x : (k == 4 ? G[9 + r].y : G[9 + r].z)));
        }
        if (rhs5) rhs5[5 * s + r] = live ? G[4 + r].w : 0.f;
    }
}

__global__ void k_export_body_dyn(View v, int scene, int n, float* f3, float* tau3, float* fc3, float* tauc3, float* I9, float* Iinv9)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const size_t i = at(v, b, scene);
    const float4 f = v.force[i], t = v.torque[i], fc = v.wl[i], tc = v.wa[i];
    if (f3) { f3[3 * b] = f.x; f3[3 * b + 1] = f.y; f3[3 * b + 2] = f.z; }
    if (tau3) { tau3[3 * b] = t.x; tau3[3 * b + 1] = t.y; tau3[3 * b + 2] = t.z; }
    if (fc3) { fc3[3 * b] = fc.x; fc3[3 * b + 1] = fc.y; fc3[3 * b + 2] = fc.z; }
    if (tauc3) { tauc3[3 * b] = tc.x; tauc3[3 * b + 1] = tc.y; tauc3[3 * b + 2] = tc.z; }
    for (int k = 0; k < 9; ++k) {
        if (I9) I9[9 * b + k] = v.Iw[at_b9(v, k, b, scene)];
        if (Iinv9) Iinv9[9 * b + k] = v.Iinvw[at_b9(v, k, b, scene)];
    }
}

// ------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------
static inline unsigned blocks_for(size_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

void launch_prepare(const View& v, cudaStream_t s) { k_prepare<<<blocks_for((size_t)v.NB * v.B, 256), 256, 0, s>>>(v); }
void launch_contact_rows(const View& v, float h, cudaStream_t s)
{
    if (v.level_mode && v.B <= 65535) k_contact_rows_lv<<<dim3(blocks_for((size_t)v.NC, 128), (unsigned)v.B), 128, 0, s>>>(v, h);
    else k_contact_rows<<<blocks_for((size_t)v.NC * v.B, 128), 128, 0, s>>>(v, h);
}
void launch_joint_rows(const View& v, float h, cudaStream_t s) { if (v.NJ > 0) k_joint_rows<<<blocks_for((size_t)v.NJ * v.B, 128), 128, 0, s>>>(v, h); }
void launch_pgs(const View& v, float dt, int iters, float mu, cudaStream_t s)
{
    static int use_tma = -1;
    if (use_tma < 0) { const char* e = getenv("SLRBS_PGS_TMA"); use_tma = e ? atoi(e) : 0; }
    if (use_tma) k_pgs_tma<<<blocks_for(v.B, 32), 32, 0, s>>>(v, dt, iters, mu);
    else k_pgs<<<blocks_for(v.B, PGS_THREADS), PGS_THREADS, 0, s>>>(v, dt, iters, mu);
}
void launch_krylov(const View& v, int solver_id, int iters, cudaStream_t s)
{
    if (v.NJ == 0) return;
    if (solver_id == 2) k_krylov<true><<<blocks_for(v.B, 64), 64, 0, s>>>(v, iters);
    else k_krylov<false><<<blocks_for(v.B, 64), 64, 0, s>>>(v, iters);
}
void launch_integrate(const View& v, float dt, cudaStream_t s) { k_integrate<<<blocks_for((size_t)v.NB * v.B, 256), 256, 0, s>>>(v, dt); }

void launch_upload_bodies(const View& v, int scene0, int ns, int n, const int* counts, const BodyStage& st, cudaStream_t s)
{ k_upload_bodies<<<blocks_for((size_t)ns * n, 256), 256, 0, s>>>(v, scene0, ns, n, counts, st); }
void launch_download_state(const View& v, int scene0, int ns, int n, const BodyStage& st, cudaStream_t s)
{ k_download_state<<<blocks_for((size_t)ns * n, 256), 256, 0, s>>>(v, scene0, ns, n, st); }
void launch_upload_joints(const View& v, int scene0, int ns, int n, const int* counts, const JointStage& st, cudaStream_t s)
{ k_upload_joints<<<blocks_for((size_t)ns * n, 256), 256, 0, s>>>(v, scene0, ns, n, counts, st); }
void launch_clear_joint_lambda(const View& v, int scene0, int ns, cudaStream_t s)
{ k_clear_joint_lambda<<<blocks_for((size_t)ns * v.NJ, 256), 256, 0, s>>>(v, scene0, ns); }
void launch_joint_lambda(const View& v, int scene0, int ns, int n, float* lam5, int upload, cudaStream_t s)
{ k_joint_lambda<<<blocks_for((size_t)ns * n, 256), 256, 0, s>>>(v, scene0, ns, n, lam5, upload); }
void launch_set_ext(const View& v, int scene0, int ns, int n, const float* f3, const float* t3, int mode, cudaStream_t s)
{ k_set_ext<<<blocks_for((size_t)ns * n, 256), 256, 0, s>>>(v, scene0, ns, n, f3, t3, mode); }
void launch_snapshot(const View& v, int scene0, int ns, int restore, cudaStream_t s)
{ k_snapshot<<<blocks_for((size_t)ns * v.NB, 256), 256, 0, s>>>(v, scene0, ns, restore); }
void launch_halo_copy(const View& v, int n, const int* src, const int* dst, cudaStream_t s)
{ if (n > 0) k_halo_copy<<<blocks_for(n, 256), 256, 0, s>>>(v, n, src, dst); }
void launch_halo_pack(const View& v, int n, const int* idx, float* buf, cudaStream_t s)
{ if (n > 0) k_halo_pack<<<blocks_for(n, 256), 256, 0, s>>>(v, n, idx, buf); }
void launch_halo_unpack(const View& v, int n, const int* idx, const float* buf, cudaStream_t s)
{ if (n > 0) k_halo_unpack<<<blocks_for(n, 256), 256, 0, s>>>(v, n, idx, buf); }
void launch_export_contacts(const View& v, int scene, int n, int* body0, int* body1, float* p3, float* n3, float* t3,
                            float* b3, float* phi, float* lam3, float* J0, float* J1, float* M0, float* M1,
                            float* A9, float* rhs3, cudaStream_t s)
{ if (n > 0) k_export_contacts<<<blocks_for(n, 128), 128, 0, s>>>(v, scene, n, body0, body1, p3, n3, t3, b3, phi, lam3, J0, J1, M0, M1, A9, rhs3); }
void launch_import_contact_lambda(const View& v, int scene, int n, const float* lam3, cudaStream_t s)
{ if (n > 0) k_import_contact_lambda<<<blocks_for(n, 128), 128, 0, s>>>(v, scene, n, lam3); }
void launch_export_joints(const View& v, int scene, int n, float* J0, float* J1, float* M0, float* M1, float* rhs5, cudaStream_t s)
{ if (n > 0) k_export_joints<<<blocks_for(n, 128), 128, 0, s>>>(v, scene, n, J0, J1, M0, M1, rhs5); }
void launch_export_body_dyn(const View& v, int scene, int n, float* f3, float* tau3, float* fc3, float* tauc3, float* I9, float* Iinv9, cudaStream_t s)
{ if (n > 0) k_export_body_dyn<<<blocks_for(n, 128), 128, 0, s>>>(v, scene, n, f3, tau3, fc3, tauc3, I9, Iinv9); }

}  // namespace slrbs

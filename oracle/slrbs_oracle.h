/*
 * slrbs_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C, single-threaded, fp32 restatement of the per-step rigid-body
 * pipeline of sheldona/slrbs (RigidBodySystem::step and everything it calls).
 * Every function in slrbs_oracle.c cites the reference file:line it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (slrbs_b200/csrc, slrbs_b200/host) never links or calls it.
 *
 * PARITY STATUS: pinned against the reference's own code, with one stated gap.
 * The reference ships no tests, golden vectors or fixtures and its
 * build system cannot run here (Eigen 3.4 + Polyscope come from network FetchContent,
 * CMakeLists.txt:15-33).  Instead, oracle/Makefile (target `ref`) compiles the reference's
 * UNMODIFIED step-path sources, where they lie under /root/reference, against an Eigen *API
 * shim* (oracle/ref_shim/Eigen/Dense) into oracle/_ref/libslrbs_ref.so, and
 * tests/test_ref_vs_oracle.py shows this restatement equal to that build BIT FOR BIT: states,
 * contact lists (pairs, order, p, n, t, b, pene, lambda), Jacobian rows and joint impulses, on
 * all five Scenarios.h scenes, two synthetic scenes, random piles, Box-PGS / CG / CR,
 * save/restore and the pre-step force hook.  tests/golden/ref_*.npz are outputs of that build
 * (tests/golden/make_golden.py) and travel to the GPU box.
 * THE GAP: Eigen itself is absent, so the summation orders INSIDE Eigen calls (listed below)
 * are shared assumptions of the shim and of this file, not observations of Eigen 3.4.  They
 * decide last-bit rounding only.  The known answers of SURVEY.md Appendix B
 * (tests/test_oracle_known_answers.py) and analytic invariants remain as independent checks.
 *
 * EXTENSIONS without a reference counterpart -- PARITY UNPINNED for these, by construction: the contact pairs behind
 * orc_set_extensions (sphere-plane, box-plane, box-box) and solver_id 3 (block principal pivoting, solve_bpp).  The
 * restatement DEFINES them; the device is compared with it, and physics is checked on residuals and resting scenes.
 *
 * Third-party arithmetic restated: Eigen, tag 3.4 (CMakeLists.txt:24-33, not
 * vendored).  Operation orders chosen (documented in slrbs_oracle.c "eig_*"):
 *   - fixed-size 3-term reductions (dot, squaredNorm, 3x3 products) use Eigen's
 *     unrolled binary tree  a0 + (a1 + a2)          (Core/Redux.h, redux_novec_unroller)
 *   - 4-term reductions on quaternion coeffs use the SSE2 packet order
 *     (a0 + a2) + (a1 + a3)                         (arch/SSE/PacketMath.h predux)
 *   - quaternion product uses the SSE2 specialisation's add order
 *     (Geometry/arch/Geometry_SIMD.h)
 *   - dynamic-size sums (JBlock products) are left-to-right.
 * These affect only last-bit rounding; compile with -O2 -ffp-contract=off.
 *
 * Conventions: quaternions cross this API as (x,y,z,w) = Eigen coeffs() order.
 * Geometry types follow include/collision/Geometry.h:7:
 *   0 = sphere {radius}, 1 = box {dimx,dimy,dimz}, 2 = plane {px,py,pz,nx,ny,nz},
 *   3 = cylinder {height, radius}.
 */
#ifndef SLRBS_ORACLE_H
#define SLRBS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_system orc_system;

enum { ORC_SPHERE = 0, ORC_BOX = 1, ORC_PLANE = 2, ORC_CYLINDER = 3,
       ORC_SDF = 4 /* EXTENSION: mesh as signed-distance grid, geom6 = {shape id} */ };
enum { ORC_CONTACT = 0, ORC_SPHERICAL = 1, ORC_HINGE = 2 };

orc_system* orc_create(void);
void        orc_destroy(orc_system* s);
void        orc_clear(orc_system* s);

/* body / joint setup (RigidBody ctor + direct field writes, RigidBodySystem::addBody/addJoint) */
int  orc_add_body(orc_system* s, float mass, int geom_type, const float* geom6, int fixed,
                  const float* x3, const float* q_xyzw, const float* v3, const float* w3);
int  orc_add_hinge(orc_system* s, int b0, int b1, const float* r0, const float* q0_xyzw,
                   const float* r1, const float* q1_xyzw);
int  orc_add_spherical(orc_system* s, int b0, int b1, const float* r0, const float* r1);
void orc_set_params(orc_system* s, float mu, int solver_id, int solver_iter, int collisions);

/* whole step and its stages (RigidBodySystem::step, src/rigidbody/RigidBodySystem.cpp:52-121) */
void orc_step(orc_system* s, float dt);
void orc_stage_prepare(orc_system* s);            /* :58-65 + computeInertias :68 + clear :75 */
void orc_stage_detect(orc_system* s);             /* detectCollisions */
void orc_stage_jacobians(orc_system* s);          /* computeContactJacobians + joint computeJacobian + zero fc/tauc */
void orc_stage_solve(orc_system* s, float dt);    /* s_solvers[solverId]->solve */
void orc_stage_scatter(orc_system* s, float dt);  /* calcConstraintForces tail :185-220 */
void orc_stage_integrate(orc_system* s, float dt);/* :99-120 */
void orc_add_force_at_pos(orc_system* s, int body, const float* pos3, const float* force3);
/* external force injected after prepare (what a pre-step callback would add) */
void orc_add_force(orc_system* s, int body, const float* f3, const float* tau3);

/* sizes */
int orc_num_bodies(const orc_system* s);
int orc_num_joints(const orc_system* s);
int orc_num_contacts(const orc_system* s);
int orc_num_joint_rows(const orc_system* s);
long long orc_pairs_tested(const orc_system* s);  /* narrow-phase dispatch attempts in the last detect */

/* state i/o; any pointer may be NULL */
void orc_get_state(const orc_system* s, float* x3, float* q_xyzw, float* v3, float* w3);
void orc_set_state(orc_system* s, const float* x3, const float* q_xyzw, const float* v3, const float* w3);
void orc_get_body_params(const orc_system* s, float* mass, float* ibody9, float* ibodyinv9,
                         int* fixed, int* geom_type, float* geom6);
void orc_get_body_dyn(const orc_system* s, float* f3, float* tau3, float* fc3, float* tauc3,
                      float* I9, float* Iinv9);
void orc_get_contacts(const orc_system* s, int* body0, int* body1, float* p3, float* n3,
                      float* t3, float* b3, float* phi, float* lambda3);
void orc_get_contact_rows(const orc_system* s, float* J0_18, float* J1_18, float* J0Minv_18, float* J1Minv_18);
void orc_get_joints(const orc_system* s, int* type, int* body0, int* body1, float* r0, float* q0_xyzw,
                    float* r1, float* q1_xyzw);
/* rows are padded to 5x6 = 30 floats per joint (unused rows zero), phi/lambda to 5 */
void orc_get_joint_rows(const orc_system* s, float* J0_30, float* J1_30, float* J0Minv_30, float* J1Minv_30,
                        float* phi5, float* lambda5);
void orc_set_joint_lambda(orc_system* s, const float* lambda5);
/* EXTENSION (no reference counterpart): solver_id 3 = block principal pivoting over joints and contacts;
 * residual of the box LCP it poses for the impulses currently stored (after orc_stage_solve) */
float orc_blcp_residual(orc_system* s, float h);
/* solver diagonal blocks and RHS as built before the PGS loop (SolverBoxPGS.cpp:134-211); needs jacobians */
void orc_get_solver_blocks(orc_system* s, float h, float* Acontact9, float* rhs_contact3, float* Ajoint25, float* rhs_joint5);

/* EXTENSION (no reference counterpart): triangle meshes as signed-distance grids. A shape = grid
 * (node (i,j,k) at origin + cell*(i,j,k), value grid[(k*ny+j)*nx+i], negative inside) + its vertices as sample points,
 * body frame. Contacts (with orc_set_extensions): mesh-plane, sphere-mesh, mesh-mesh, mesh-box; at most the 4 deepest points of a
 * pair, one normal per pair. orc_build_sdf: grid of a closed triangle mesh. orc_load_obj: 'v' / 'f' lines as
 * src/util/OBJLoader.cpp:34-122 reads them. */
int  orc_add_sdf_shape(orc_system* s, int nx, int ny, int nz, const float* origin3, float cell, const float* grid,
                       int nv, const float* verts);
void orc_build_sdf(int nv, const float* verts, int nt, const int* tris, int nx, int ny, int nz, const float* origin3,
                   float cell, float* out);
int  orc_load_obj(const char* path, float* verts, int max_v, int* tris, int max_t, int* nv_out, int* nt_out);

/* RigidBodySystemState save/restore (src/rigidbody/RigidBodyState.cpp:18-72) */
void orc_save_state(orc_system* s);
void orc_restore_state(orc_system* s);

/* Scenarios.h scene builders (include/rigidbody/Scenarios.h) */
void orc_scene_marble_box(orc_system* s);        /* :18-80   */
void orc_scene_sphere_on_box(orc_system* s);     /* :84-108  */
void orc_scene_swinging_boxes(orc_system* s);    /* :112-148 */
void orc_scene_cylinder_on_plane(orc_system* s); /* :153-174 */
void orc_scene_car(orc_system* s);               /* :178-247 */
/* synthetic scenes of SURVEY.md 8(d) built from the same reference primitives */
void orc_scene_sphere_stack(orc_system* s, int k);                 /* C1p */
void orc_scene_marble_box_n(orc_system* s, int nx, int nz, int ny);/* MB1k family */

/* EXTENSIONS without a reference counterpart: flags bit 0 = sphere-plane, box-plane and box-box
 * contacts, appended after the reference's dispatch chain (off by default) */
void orc_set_extensions(orc_system* s, int flags);
void orc_scene_box_stack(orc_system* s, int k);  /* k unit boxes on a Plane (BASELINE config 0); enables the extension */

/* timing helper for the CPU baseline: runs n steps, returns seconds (CLOCK_MONOTONIC) */
double orc_time_steps(orc_system* s, float dt, int n);

#ifdef __cplusplus
}
#endif
#endif

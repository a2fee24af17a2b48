/*
 * slrbs_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 * See slrbs_oracle.h for status ("parity unpinned by reference tests"), conventions
 * and the Eigen 3.4 operation orders restated here.  All citations are relative
 * to the reference tree (sheldona/slrbs).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 */
#include "slrbs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ */
/* Eigen 3.4 arithmetic restated                */
/* ------------------------------------------------------------------ */
typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } quat;     /* Eigen coeffs() order */
typedef struct { float m[3][3]; } m3;          /* m[row][col]          */

static inline v3 V3(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 v3_add(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_neg(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline v3 v3_scale(float s, v3 a) { return V3(s * a.x, s * a.y, s * a.z); }
static inline v3 v3_div(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
/* fixed-size 3-term reduction: Eigen redux_novec_unroller tree a0 + (a1 + a2) */
static inline float sum3(float a0, float a1, float a2) { return a0 + (a1 + a2); }
static inline float v3_dot(v3 a, v3 b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline float v3_sqnorm(v3 a) { return sum3(a.x * a.x, a.y * a.y, a.z * a.z); }
static inline float v3_norm(v3 a) { return sqrtf(v3_sqnorm(a)); }
/* MatrixBase::cross (Geometry/OrthoMethods.h) */
static inline v3 v3_cross(v3 a, v3 b)
{
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* MatrixBase::normalize: z = squaredNorm(); if (z > 0) *this /= sqrt(z) */
static inline v3 v3_normalized(v3 a)
{
    float z = v3_sqnorm(a);
    if (z > 0.0f) return v3_div(a, sqrtf(z));
    return a;
}
static inline float get3(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

/* 4-term reduction on quaternion coeffs: SSE2 predux (a0 + a2) + (a1 + a3) */
static inline float q_sqnorm(quat q) { return (q.x * q.x + q.z * q.z) + (q.y * q.y + q.w * q.w); }
/* QuaternionBase::_transformVector */
static inline v3 q_rot(quat q, v3 v)
{
    v3 qv = V3(q.x, q.y, q.z);
    v3 uv = v3_cross(qv, v);
    uv = v3_add(uv, uv);
    return v3_add(v3_add(v, v3_scale(q.w, uv)), v3_cross(qv, uv));
}
/* QuaternionBase::inverse: conjugate / squaredNorm (true divides), zero if n2 == 0 */
static inline quat q_inverse(quat q)
{
    float n2 = q_sqnorm(q);
    quat r;
    if (n2 > 0.0f) { r.x = -q.x / n2; r.y = -q.y / n2; r.z = -q.z / n2; r.w = q.w / n2; }
    else { r.x = r.y = r.z = r.w = 0.0f; }
    return r;
}
/* quaternion product, SSE2 specialisation add order (Geometry_SIMD.h):
 * res = (a * b.wwww - a.zxyx * b.yzxx) + (+-)(a.yzxz * b.zxyz + a.wwwy * b.xyzy), sign - on w */
static inline quat q_mul(quat a, quat b)
{
    quat r;
    r.x = (a.x * b.w - a.z * b.y) + (a.y * b.z + a.w * b.x);
    r.y = (a.y * b.w - a.x * b.z) + (a.z * b.x + a.w * b.y);
    r.z = (a.z * b.w - a.y * b.x) + (a.x * b.y + a.w * b.z);
    r.w = (a.w * b.w - a.x * b.x) + (-(a.z * b.z + a.y * b.y));
    return r;
}
/* QuaternionBase::normalize -> coeffs().normalize() */
static inline quat q_normalized(quat q)
{
    float z = q_sqnorm(q);
    if (z > 0.0f) {
        float s = sqrtf(z);
        q.x /= s; q.y /= s; q.z /= s; q.w /= s;
    }
    return q;
}
/* QuaternionBase::toRotationMatrix */
static inline m3 q_to_matrix(quat q)
{
    m3 R;
    const float tx = 2.0f * q.x, ty = 2.0f * q.y, tz = 2.0f * q.z;
    const float twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const float txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const float tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R.m[0][0] = 1.0f - (tyy + tzz); R.m[0][1] = txy - twz;          R.m[0][2] = txz + twy;
    R.m[1][0] = txy + twz;          R.m[1][1] = 1.0f - (txx + tzz); R.m[1][2] = tyz - twx;
    R.m[2][0] = txz - twy;          R.m[2][1] = tyz + twx;          R.m[2][2] = 1.0f - (txx + tyy);
    return R;
}
/* AngleAxis -> Quaternion (Geometry/Quaternion.h operator=(AngleAxis)) */
static inline quat q_from_angle_axis(float angle, v3 axis)
{
    const float ha = 0.5f * angle;
    const float s = sinf(ha);
    quat q; q.w = cosf(ha); q.x = s * axis.x; q.y = s * axis.y; q.z = s * axis.z;
    return q;
}
static inline quat q_identity(void) { quat q = { 0.f, 0.f, 0.f, 1.f }; return q; }

static inline m3 m3_zero(void) { m3 r; memset(&r, 0, sizeof r); return r; }
static inline m3 m3_mul(const m3* A, const m3* B)
{
    m3 C;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C.m[i][j] = sum3(A->m[i][0] * B->m[0][j], A->m[i][1] * B->m[1][j], A->m[i][2] * B->m[2][j]);
    return C;
}
static inline v3 m3_mulv(const m3* A, v3 v)
{
    return V3(sum3(A->m[0][0] * v.x, A->m[0][1] * v.y, A->m[0][2] * v.z),
              sum3(A->m[1][0] * v.x, A->m[1][1] * v.y, A->m[1][2] * v.z),
              sum3(A->m[2][0] * v.x, A->m[2][1] * v.y, A->m[2][2] * v.z));
}
/* Matrix3f::inverse (LU/InverseImpl.h compute_inverse_size3_helper: cofactors * (1/det)) */
static inline float cof3(const m3* A, int i, int j)
{
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return A->m[i1][j1] * A->m[i2][j2] - A->m[i1][j2] * A->m[i2][j1];
}
static inline m3 m3_inverse(const m3* A)
{
    m3 R;
    const float c00 = cof3(A, 0, 0), c10 = cof3(A, 1, 0), c20 = cof3(A, 2, 0);
    const float det = sum3(c00 * A->m[0][0], c10 * A->m[1][0], c20 * A->m[2][0]);
    const float invdet = 1.0f / det;
    /* result(r,c) = cofactor(c,r) * invdet */
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            R.m[r][c] = cof3(A, c, r) * invdet;
    return R;
}

/* ------------------------------------------------------------------ */
/* data model (RigidBody.h:38-57, Joint.h:24-39, Contact.h:20-42)      */
/* ------------------------------------------------------------------ */
typedef struct { int kind; int idx; } cref;   /* kind 0 = contact, 1 = joint */

typedef struct {
    int   fixed;
    float mass;
    m3    I, Iinv, Ibody, IbodyInv;
    v3    x; quat q;
    v3    xdot, omega, f, tau, fc, tauc;
    int   geom_type;
    float radius;          /* sphere, cylinder */
    v3    dim;             /* box */
    float height;          /* cylinder */
    v3    plane_p, plane_n;
    int   shape;           /* ORC_SDF: index into sys->shapes (extension) */
    int  *contacts; int n_contacts, cap_contacts;   /* indices into sys->contacts, detection order */
    int  *joints;   int n_joints,   cap_joints;     /* indices into sys->joints, insertion order   */
    /* RigidBodyState snapshot */
    v3 sx, sxdot, somega; quat sq;
} body_t;

typedef struct {
    int   type;            /* ORC_CONTACT / ORC_SPHERICAL / ORC_HINGE */
    int   body0, body1;
    int   dim;
    float J0[5][6], J1[5][6], J0Minv[5][6], J1Minv[5][6];
    float phi[5], lambda[5];
    unsigned idx;
    v3    r0, r1; quat q0, q1;
    v3    p, n, t, b; float pene;    /* contact only */
} con_t;

/* EXTENSION: a triangle mesh as a signed-distance grid (negative inside) plus its vertices as sample points, both in
 * the body frame; node (i, j, k) sits at origin + cell * (i, j, k), value grid[(k * ny + j) * nx + i] */
typedef struct {
    int   nx, ny, nz, nv;
    v3    origin; float cell, radius;     /* radius: largest vertex norm (bounding sphere about the body origin) */
    v3    bbmin, bbmax;
    float *grid, *verts;
} sdf_shape;

struct orc_system {
    sdf_shape *shapes; int n_shapes, cap_shapes;
    body_t *bodies;   int n_bodies,   cap_bodies;
    con_t  *joints;   int n_joints,   cap_joints;
    con_t  *contacts; int n_contacts, cap_contacts;
    float mu;               /* Contact::mu, Contact.cpp:4 */
    int   solverId;         /* RigidBodySystem.cpp:23 */
    int   solverIter;       /* RigidBodySystem.cpp:22 */
    int   collisions;       /* m_collisionsEnabled, :26 */
    long long pairs_tested;
    int   ext_pairs;        /* extension: sphere-plane, box-plane, box-box contacts (no reference counterpart) */
};

static void push_int(int **arr, int *n, int *cap, int v)
{
    if (*n == *cap) { *cap = *cap ? *cap * 2 : 8; *arr = (int*)realloc(*arr, (size_t)*cap * sizeof(int)); }
    (*arr)[(*n)++] = v;
}

orc_system* orc_create(void)
{
    orc_system* s = (orc_system*)calloc(1, sizeof *s);
    s->mu = 0.8f; s->solverId = 0; s->solverIter = 10; s->collisions = 1;
    return s;
}

/* RigidBodySystem::clear, RigidBodySystem.cpp:123-146 */
void orc_clear(orc_system* s)
{
    for (int i = 0; i < s->n_bodies; ++i) { free(s->bodies[i].contacts); free(s->bodies[i].joints); }
    s->n_bodies = s->n_joints = s->n_contacts = 0;
}

void orc_destroy(orc_system* s)
{
    if (!s) return;
    orc_clear(s);
    for (int i = 0; i < s->n_shapes; ++i) { free(s->shapes[i].grid); free(s->shapes[i].verts); }
    free(s->shapes);
    free(s->bodies); free(s->joints); free(s->contacts);
    free(s);
}

void orc_set_params(orc_system* s, float mu, int solver_id, int solver_iter, int collisions)
{
    s->mu = mu; s->solverId = solver_id; s->solverIter = solver_iter; s->collisions = collisions;
}

/* Geometry::computeInertia, include/collision/Geometry.h:38-43,63-70,91-102,120-124 */
static m3 compute_inertia(const body_t* b, float mass)
{
    m3 I = m3_zero();
    switch (b->geom_type) {
    case ORC_SPHERE:
        I.m[0][0] = I.m[1][1] = I.m[2][2] = (2.0f / 5.0f) * mass * b->radius * b->radius;
        break;
    case ORC_BOX:
    case ORC_SDF:
        I.m[0][0] = (1.0f / 12.0f) * mass * (b->dim.y * b->dim.y + b->dim.z * b->dim.z);
        I.m[1][1] = (1.0f / 12.0f) * mass * (b->dim.x * b->dim.x + b->dim.z * b->dim.z);
        I.m[2][2] = (1.0f / 12.0f) * mass * (b->dim.x * b->dim.x + b->dim.y * b->dim.y);
        break;
    case ORC_CYLINDER: {
        const float s12 = 1.0f / 12.0f;
        const float h2 = b->height * b->height;
        const float r2 = b->radius * b->radius;
        const float r2_h2_3 = (3.0f * r2 + h2);
        I.m[0][0] = s12 * mass * r2_h2_3;
        I.m[1][1] = 0.5f * mass * r2;
        I.m[2][2] = s12 * mass * r2_h2_3;
        break; }
    default: /* plane: zero inertia, Geometry.h:120-124 */
        break;
    }
    return I;
}

/* RigidBody::RigidBody, src/rigidbody/RigidBody.cpp:9-45 (mesh registration omitted: render-only) */
int orc_add_body(orc_system* s, float mass, int geom_type, const float* g, int fixed,
                 const float* x3, const float* q, const float* v, const float* w)
{
    if (s->n_bodies == s->cap_bodies) {
        s->cap_bodies = s->cap_bodies ? s->cap_bodies * 2 : 64;
        s->bodies = (body_t*)realloc(s->bodies, (size_t)s->cap_bodies * sizeof(body_t));
    }
    body_t* b = &s->bodies[s->n_bodies];
    memset(b, 0, sizeof *b);
    b->fixed = fixed; b->mass = mass;
    b->geom_type = geom_type;
    switch (geom_type) {
    case ORC_SPHERE:   b->radius = g[0]; break;
    case ORC_BOX:      b->dim = V3(g[0], g[1], g[2]); break;
    case ORC_PLANE:    b->plane_p = V3(g[0], g[1], g[2]); b->plane_n = V3(g[3], g[4], g[5]); break;
    case ORC_CYLINDER: b->height = g[0]; b->radius = g[1]; break;
    case ORC_SDF:      /* extension: inertia of the solid box around the mesh vertices */
        b->shape = (int)g[0];
        b->dim = v3_sub(s->shapes[b->shape].bbmax, s->shapes[b->shape].bbmin);
        break;
    }
    b->x = x3 ? V3(x3[0], x3[1], x3[2]) : V3(0, 0, 0);
    if (q) { b->q.x = q[0]; b->q.y = q[1]; b->q.z = q[2]; b->q.w = q[3]; } else b->q = q_identity();
    b->xdot  = v ? V3(v[0], v[1], v[2]) : V3(0, 0, 0);
    b->omega = w ? V3(w[0], w[1], w[2]) : V3(0, 0, 0);
    b->Iinv = m3_zero();               /* RigidBody.cpp:18 (I is left uninitialised there; zero here) */
    b->I = m3_zero();
    b->Ibody = compute_inertia(b, mass);   /* :27 */
    b->IbodyInv = m3_inverse(&b->Ibody);   /* :28 */
    return s->n_bodies++;
}

static con_t* new_joint(orc_system* s)
{
    if (s->n_joints == s->cap_joints) {
        s->cap_joints = s->cap_joints ? s->cap_joints * 2 : 16;
        s->joints = (con_t*)realloc(s->joints, (size_t)s->cap_joints * sizeof(con_t));
    }
    con_t* j = &s->joints[s->n_joints];
    memset(j, 0, sizeof *j);
    return j;
}

/* RigidBodySystem::addJoint, RigidBodySystem.cpp:44-50 */
static int register_joint(orc_system* s)
{
    const int id = s->n_joints++;
    con_t* j = &s->joints[id];
    body_t* b0 = &s->bodies[j->body0];
    body_t* b1 = &s->bodies[j->body1];
    push_int(&b0->joints, &b0->n_joints, &b0->cap_joints, id);
    push_int(&b1->joints, &b1->n_joints, &b1->cap_joints, id);
    return id;
}

/* Hinge::Hinge, src/joint/Hinge.cpp:21-31 */
int orc_add_hinge(orc_system* s, int b0, int b1, const float* r0, const float* q0,
                  const float* r1, const float* q1)
{
    con_t* j = new_joint(s);
    j->type = ORC_HINGE; j->dim = 5; j->body0 = b0; j->body1 = b1;
    j->r0 = V3(r0[0], r0[1], r0[2]); j->r1 = V3(r1[0], r1[1], r1[2]);
    j->q0.x = q0[0]; j->q0.y = q0[1]; j->q0.z = q0[2]; j->q0.w = q0[3];
    j->q1.x = q1[0]; j->q1.y = q1[1]; j->q1.z = q1[2]; j->q1.w = q1[3];
    return register_joint(s);
}

/* Spherical::Spherical, src/joint/Spherical.cpp:22-32 */
int orc_add_spherical(orc_system* s, int b0, int b1, const float* r0, const float* r1)
{
    con_t* j = new_joint(s);
    j->type = ORC_SPHERICAL; j->dim = 3; j->body0 = b0; j->body1 = b1;
    j->r0 = V3(r0[0], r0[1], r0[2]); j->r1 = V3(r1[0], r1[1], r1[2]);
    j->q0 = q_identity(); j->q1 = q_identity();
    return register_joint(s);
}

/* ------------------------------------------------------------------ */
/* S1/S2/S4: reset, inertia update, contact clear                      */
/* ------------------------------------------------------------------ */
/* RigidBody::updateInertiaMatrix, src/rigidbody/RigidBody.cpp:78-89
 * I = q * Ibody * q.inverse()  ==  (R(q) * Ibody) * R(q^-1) */
static void update_inertia(body_t* b)
{
    if (!b->fixed) {
        const m3 R = q_to_matrix(b->q);
        const m3 Rinv = q_to_matrix(q_inverse(b->q));
        m3 T = m3_mul(&R, &b->Ibody);
        b->I = m3_mul(&T, &Rinv);
        T = m3_mul(&R, &b->IbodyInv);
        b->Iinv = m3_mul(&T, &Rinv);
    } else {
        b->Iinv = m3_zero();
    }
}

/* RigidBodySystem.cpp:58-65, :68 (computeInertias :148-154), :75 (CollisionDetect::clear, CollisionDetect.cpp:87-100) */
void orc_stage_prepare(orc_system* s)
{
    for (int i = 0; i < s->n_bodies; ++i) {
        body_t* b = &s->bodies[i];
        b->f = v3_scale(b->mass, V3(0.f, -9.81f, 0.f));
        b->tau = V3(0, 0, 0); b->fc = V3(0, 0, 0); b->tauc = V3(0, 0, 0);
        b->n_contacts = 0;
    }
    for (int i = 0; i < s->n_bodies; ++i) update_inertia(&s->bodies[i]);
    s->n_contacts = 0;
    for (int i = 0; i < s->n_bodies; ++i) s->bodies[i].n_contacts = 0;
}

/* RigidBody::addForceAtPos, src/rigidbody/RigidBody.cpp:91-96 */
void orc_add_force_at_pos(orc_system* s, int body, const float* pos, const float* force)
{
    body_t* b = &s->bodies[body];
    const v3 F = V3(force[0], force[1], force[2]);
    const v3 r = v3_sub(V3(pos[0], pos[1], pos[2]), b->x);
    b->f = v3_add(b->f, F);
    b->tau = v3_add(b->tau, v3_cross(r, F));
}

void orc_add_force(orc_system* s, int body, const float* f3, const float* tau3)
{
    body_t* b = &s->bodies[body];
    if (f3)   b->f   = v3_add(b->f,   V3(f3[0], f3[1], f3[2]));
    if (tau3) b->tau = v3_add(b->tau, V3(tau3[0], tau3[1], tau3[2]));
}

/* ------------------------------------------------------------------ */
/* S5-S9: collision detection                                          */
/* ------------------------------------------------------------------ */
/* Contact::Contact, src/contact/Contact.cpp:11-26 */
static void new_contact(orc_system* s, int body0, int body1, v3 p, v3 n, float pene)
{
    if (s->n_contacts == s->cap_contacts) {
        s->cap_contacts = s->cap_contacts ? s->cap_contacts * 2 : 256;
        s->contacts = (con_t*)realloc(s->contacts, (size_t)s->cap_contacts * sizeof(con_t));
    }
    const int id = s->n_contacts++;
    con_t* c = &s->contacts[id];
    memset(c, 0, sizeof *c);
    c->type = ORC_CONTACT; c->dim = 3;
    c->body0 = body0; c->body1 = body1;
    c->p = p; c->n = n; c->pene = pene;
    c->phi[0] = pene;
    c->q0 = q_identity(); c->q1 = q_identity();
    body_t* b0 = &s->bodies[body0];
    body_t* b1 = &s->bodies[body1];
    push_int(&b0->contacts, &b0->n_contacts, &b0->cap_contacts, id);   /* :24 */
    push_int(&b1->contacts, &b1->n_contacts, &b1->cap_contacts, id);   /* :25 */
}

/* distancePointPlane, CollisionDetect.cpp:12-17 */
static inline float distance_point_plane(v3 p, v3 plane_p, v3 plane_n)
{
    return v3_dot(v3_sub(p, plane_p), plane_n);
}

/* collisionDetectSphereSphere, CollisionDetect.cpp:102-123 */
static void detect_sphere_sphere(orc_system* s, int i0, int i1)
{
    const body_t* body0 = &s->bodies[i0];
    const body_t* body1 = &s->bodies[i1];
    const v3 vec = v3_sub(body0->x, body1->x);
    const float rsum = (body0->radius + body1->radius);
    const float dist = v3_norm(vec);
    if (dist < (rsum + 1e-3f)) {
        const v3 n = v3_div(vec, dist);
        const v3 p = v3_scale(0.5f, v3_add(v3_sub(body0->x, v3_scale(body0->radius, n)),
                                           v3_add(body1->x, v3_scale(body1->radius, n))));
        const float phi = fminf(0.0f, dist - rsum);
        new_contact(s, i0, i1, p, n, phi);
    }
}

/* collisionDetectSphereBox, CollisionDetect.cpp:125-150 (body0 = sphere, body1 = box) */
static void detect_sphere_box(orc_system* s, int i0, int i1)
{
    const body_t* body0 = &s->bodies[i0];
    const body_t* body1 = &s->bodies[i1];
    const v3 clocal = q_rot(q_inverse(body1->q), v3_sub(body0->x, body1->x));
    float c[3] = { clocal.x, clocal.y, clocal.z };
    float d[3] = { body1->dim.x, body1->dim.y, body1->dim.z };
    float q[3] = { 0.f, 0.f, 0.f };
    for (int i = 0; i < 3; ++i) {
        q[i] = c[i];
        if (q[i] < (-d[i] / 2.0f)) q[i] = -d[i] / 2.0f;
        else if (q[i] > (d[i] / 2.0f)) q[i] = (d[i] / 2.0f);
    }
    const v3 qv = V3(q[0], q[1], q[2]);
    const v3 dx = v3_sub(clocal, qv);
    const float dist = v3_norm(dx);
    if (dist < (body0->radius + 1e-3f)) {
        const v3 n = q_rot(body1->q, v3_div(dx, dist));
        const v3 p = v3_add(q_rot(body1->q, qv), body1->x);
        const float phi = fminf(0.0f, dist - body0->radius);
        new_contact(s, i0, i1, p, n, phi);
    }
}

/* collisionDetectCylinderPlane, CollisionDetect.cpp:152-274 (body0 = cylinder, body1 = plane) */
static void detect_cylinder_plane(orc_system* s, int i0, int i1)
{
    const body_t* body0 = &s->bodies[i0];
    const body_t* body1 = &s->bodies[i1];
    const float cyl_h = body0->height, cyl_r = body0->radius;

    const v3 cyldir = q_rot(body0->q, V3(0, 1, 0));                          /* :158 */
    const v3 planen = q_rot(body1->q, body1->plane_n);                       /* :159 */
    const v3 planep = v3_add(q_rot(body1->q, body1->plane_p), body1->x);     /* :160 */
    const float dp = v3_dot(cyldir, planen);

    if (fabsf(dp) > 0.995f) {                                                /* :164 cap face */
        v3 w = (dp < 0.0f) ? cyldir : v3_neg(cyldir);
        v3 u, v;
        if (fabsf(v3_dot(w, V3(1, 0, 0))) > 0.01f) u = v3_cross(w, V3(1, 0, 0));
        else                                        u = v3_cross(w, V3(0, 0, 1));
        u = v3_normalized(u);
        v = v3_cross(w, u);
        v = v3_normalized(v);
        const v3 a = v3_add(body0->x, v3_scale(0.5f * cyl_h, w));            /* :189 (0.5f*h)*w */
        const float dist = distance_point_plane(a, planep, planen);
        if (dist < 1e-3f) {
            const v3 n = planen;
            const v3 pA = v3_add(a, v3_scale(cyl_r, u));
            const v3 pB = v3_sub(a, v3_scale(cyl_r, u));
            const v3 pC = v3_add(a, v3_scale(cyl_r, v));
            const v3 pD = v3_sub(a, v3_scale(cyl_r, v));
            const float phiA = distance_point_plane(pA, planep, planen);
            const float phiB = distance_point_plane(pB, planep, planen);
            const float phiC = distance_point_plane(pC, planep, planen);
            const float phiD = distance_point_plane(pD, planep, planen);
            if (phiA < 1e-3f) new_contact(s, i0, i1, pA, n, fminf(0.0f, phiA));
            if (phiB < 1e-3f) new_contact(s, i0, i1, pB, n, fminf(0.0f, phiB));
            if (phiC < 1e-3f) new_contact(s, i0, i1, pC, n, fminf(0.0f, phiC));
            if (phiD < 1e-3f) new_contact(s, i0, i1, pD, n, fminf(0.0f, phiD));
        }
    } else if (fabsf(dp) < 1e-2f) {                                          /* :215 side */
        const v3 w = cyldir;
        v3 u = v3_cross(v3_cross(planen, w), w);
        if (v3_dot(u, planen) > 0.0f) u = v3_neg(u);
        u = v3_normalized(u);
        const v3 cylpos = body0->x;
        const float dist = distance_point_plane(cylpos, planep, planen) - cyl_r;
        if (dist < 0.01f) {                                                   /* :227 */
            const v3 n = planen;
            /* cylpos + 0.5f*h*cyldir + r*u : ((cylpos + (0.5f*h)*cyldir) + r*u) */
            const v3 hw = v3_scale(0.5f * cyl_h, cyldir);
            const v3 ru = v3_scale(cyl_r, u);
            const v3 pA = v3_add(v3_add(cylpos, hw), ru);
            const v3 pB = v3_add(v3_sub(cylpos, hw), ru);
            const float phiA = distance_point_plane(pA, planep, planen);
            const float phiB = distance_point_plane(pB, planep, planen);
            if (phiA < 1e-3f) new_contact(s, i0, i1, pA, n, fminf(0.0f, phiA));
            if (phiB < 1e-3f) new_contact(s, i0, i1, pB, n, fminf(0.0f, phiB));
        }
    } else {                                                                  /* :241 general */
        v3 w = (dp < 0.0f) ? cyldir : v3_neg(cyldir);
        w = v3_normalized(w);
        v3 u = v3_cross(v3_cross(planen, w), w);
        u = v3_normalized(u);
        if (v3_dot(u, planen) > 0.0f) u = v3_neg(u);
        u = v3_normalized(u);
        const v3 a = v3_add(v3_add(body0->x, v3_scale(0.5f * cyl_h, w)), v3_scale(cyl_r, u));
        const float dist = distance_point_plane(a, planep, planen);
        if (dist < 1e-3f) {
            new_contact(s, i0, i1, a, planen, fminf(0.0f, dist));
        }
    }
}

/* ------------------------------------------------------------------ */
/* EXTENSION pairs: sphere-plane, box-plane, box-box. */
/* The reference silently ignores these pairs (CollisionDetect.cpp:45-73): */
/* there is NO reference behaviour to match ("parity unpinned - no         */
/* reference path"). Off by default; orc_set_extensions(s, 1) appends them */
/* AFTER the reference's five cases, so existing results do not change.    */
/* This file only restates the device routines so that the contact lists   */
/* can be compared bit for bit; physics is validated on penetration.       */
/* ------------------------------------------------------------------ */
void orc_set_extensions(orc_system* s, int flags) { s->ext_pairs = flags & 1; }

/* sphere (body0) against the half space of a plane (body1) */
static void ext_sphere_plane(orc_system* s, int i0, int i1)
{
    const body_t* sp = &s->bodies[i0];
    const body_t* pl = &s->bodies[i1];
    const v3 planen = q_rot(pl->q, pl->plane_n);
    const v3 planep = v3_add(q_rot(pl->q, pl->plane_p), pl->x);
    const float dist = distance_point_plane(sp->x, planep, planen) - sp->radius;
    if (dist < 1e-3f)
        new_contact(s, i0, i1, v3_sub(sp->x, v3_scale(sp->radius, planen)), planen, fminf(0.0f, dist));
}

/* box (body0) against a plane (body1): its corners below the plane, at most 4, in corner order
 * (bit 0 of the corner index = +x, bit 1 = +y, bit 2 = +z) */
static void ext_box_plane(orc_system* s, int i0, int i1)
{
    const body_t* bx = &s->bodies[i0];
    const body_t* pl = &s->bodies[i1];
    const v3 planen = q_rot(pl->q, pl->plane_n);
    const v3 planep = v3_add(q_rot(pl->q, pl->plane_p), pl->x);
    const v3 h = V3(0.5f * bx->dim.x, 0.5f * bx->dim.y, 0.5f * bx->dim.z);
    int cnt = 0;
    for (int k = 0; k < 8 && cnt < 4; ++k) {
        const v3 local = V3((k & 1) ? h.x : -h.x, (k & 2) ? h.y : -h.y, (k & 4) ? h.z : -h.z);
        const v3 c = v3_add(q_rot(bx->q, local), bx->x);
        const float d = distance_point_plane(c, planep, planen);
        if (d < 1e-3f) { new_contact(s, i0, i1, c, planen, fminf(0.0f, d)); ++cnt; }
    }
}

static inline v3 m3_col(const m3* R, int k) { return V3(R->m[0][k], R->m[1][k], R->m[2][k]); }
static inline float proj_radius(const v3 ax[3], const float h[3], v3 L)
{
    return sum3(h[0] * fabsf(v3_dot(ax[0], L)), h[1] * fabsf(v3_dot(ax[1], L)), h[2] * fabsf(v3_dot(ax[2], L)));
}

/* box (body0 = i) against box (body1 = j): separating-axis test over the 15 axes (3 + 3 face normals, 9 edge
 * cross products); the axis of least overlap gives the normal (faces preferred). Face contact: the incident face
 * is clipped against the side planes of the reference face and the clipped corners at or below the reference face
 * become contacts (at most 4: the extremes along the face diagonals). Edge contact: the midpoint of the closest
 * points of the two edges. The normal points from body1 to body0 like every other contact. */
static void ext_box_box(orc_system* s, int i0, int i1)
{
    const body_t* A = &s->bodies[i0];
    const body_t* B = &s->bodies[i1];
    const m3 Ra = q_to_matrix(A->q), Rb = q_to_matrix(B->q);
    const v3 ax[2][3] = { { m3_col(&Ra, 0), m3_col(&Ra, 1), m3_col(&Ra, 2) }, { m3_col(&Rb, 0), m3_col(&Rb, 1), m3_col(&Rb, 2) } };
    const float hh[2][3] = { { 0.5f * A->dim.x, 0.5f * A->dim.y, 0.5f * A->dim.z }, { 0.5f * B->dim.x, 0.5f * B->dim.y, 0.5f * B->dim.z } };
    const v3 xc[2] = { A->x, B->x };
    const v3 t = v3_sub(B->x, A->x);
    float best = 3.0e38f; int best_idx = -1; v3 bestL = V3(0, 0, 0);
    for (int idx = 0; idx < 15; ++idx) {
        v3 L;
        if (idx < 3) L = ax[0][idx];
        else if (idx < 6) L = ax[1][idx - 3];
        else {
            const v3 c = v3_cross(ax[0][(idx - 6) / 3], ax[1][(idx - 6) % 3]);
            const float l2 = v3_sqnorm(c);
            if (l2 < 1e-6f) continue;                         /* (nearly) parallel edges: covered by the face axes */
            L = v3_div(c, sqrtf(l2));
        }
        const float overlap = (proj_radius(ax[0], hh[0], L) + proj_radius(ax[1], hh[1], L)) - fabsf(v3_dot(t, L));
        if (overlap < -1e-3f) return;                         /* separated by more than the contact margin */
        if (idx < 6) { if (overlap < best) { best = overlap; best_idx = idx; bestL = L; } }
        else if (overlap < best - 0.05f * fabsf(best) - 1e-4f) { best = overlap; best_idx = idx; bestL = L; }
    }
    if (best_idx < 0) return;
    if (best_idx >= 6) {
        /* edge - edge */
        const int ea = (best_idx - 6) / 3, eb = (best_idx - 6) % 3;
        const v3 nab = (v3_dot(t, bestL) > 0.0f) ? bestL : v3_neg(bestL);        /* from A towards B */
        v3 pa = A->x, pb = B->x;
        for (int k = 0; k < 3; ++k) {
            if (k != ea) pa = v3_add(pa, v3_scale((v3_dot(ax[0][k], nab) > 0.0f) ? hh[0][k] : -hh[0][k], ax[0][k]));
            if (k != eb) pb = v3_add(pb, v3_scale((v3_dot(ax[1][k], nab) < 0.0f) ? hh[1][k] : -hh[1][k], ax[1][k]));
        }
        const v3 ua = ax[0][ea], ub = ax[1][eb];
        const v3 d = v3_sub(pb, pa);
        const float a = v3_dot(ua, ub), da = v3_dot(d, ua), db = v3_dot(d, ub);
        const float den = 1.0f - a * a;
        const float sa = (da - a * db) / den, sb = (a * da - db) / den;
        const v3 ca = v3_add(pa, v3_scale(sa, ua)), cb = v3_add(pb, v3_scale(sb, ub));
        new_contact(s, i0, i1, v3_scale(0.5f, v3_add(ca, cb)), v3_neg(nab), fminf(0.0f, -best));
        return;
    }
    /* face contact: reference box r (0 = A, 1 = B), incident box c */
    const int r = best_idx < 3 ? 0 : 1, c = 1 - r, k = best_idx < 3 ? best_idx : best_idx - 3;
    const v3 rc = v3_sub(xc[c], xc[r]);
    const v3 nr = (v3_dot(rc, ax[r][k]) > 0.0f) ? ax[r][k] : v3_neg(ax[r][k]);   /* outward normal of the reference face */
    int m = 0; float dmax = -1.0f;
    for (int q = 0; q < 3; ++q) { const float dq = fabsf(v3_dot(ax[c][q], nr)); if (dq > dmax) { dmax = dq; m = q; } }
    const float sgn = (v3_dot(ax[c][m], nr) > 0.0f) ? -1.0f : 1.0f;               /* incident face looks against nr */
    const int u = (m + 1) % 3, w = (m + 2) % 3;
    const v3 fc = v3_add(xc[c], v3_scale(sgn * hh[c][m], ax[c][m]));
    const v3 eu = v3_scale(hh[c][u], ax[c][u]), ew = v3_scale(hh[c][w], ax[c][w]);
    v3 poly[16], tmp[16];
    int np = 4;
    poly[0] = v3_add(v3_add(fc, eu), ew); poly[1] = v3_add(v3_sub(fc, eu), ew);
    poly[2] = v3_sub(v3_sub(fc, eu), ew); poly[3] = v3_sub(v3_add(fc, eu), ew);
    const int ru = (k + 1) % 3, rw = (k + 2) % 3;
    for (int side = 0; side < 4 && np > 0; ++side) {         /* Sutherland-Hodgman against the 4 side planes */
        const int axn = side < 2 ? ru : rw;
        const float sg = (side & 1) ? -1.0f : 1.0f;
        const v3 pn = v3_scale(sg, ax[r][axn]);
        const float lim = hh[r][axn];
        int nn = 0;
        for (int e = 0; e < np; ++e) {
            const v3 p0 = poly[e], p1 = poly[(e + 1) % np];
            const float d0 = v3_dot(v3_sub(p0, xc[r]), pn) - lim, d1 = v3_dot(v3_sub(p1, xc[r]), pn) - lim;
            if (d0 <= 0.0f) tmp[nn++] = p0;
            if ((d0 <= 0.0f) != (d1 <= 0.0f)) {
                const float tt = d0 / (d0 - d1);
                tmp[nn++] = v3_add(p0, v3_scale(tt, v3_sub(p1, p0)));
            }
        }
        np = nn;
        for (int e = 0; e < np; ++e) poly[e] = tmp[e];
    }
    /* keep the corners at or below the reference face (+ margin) */
    v3 pts[16]; float dep[16], ca_[16], cb_[16];
    int nk = 0;
    for (int e = 0; e < np; ++e) {
        const v3 rel = v3_sub(poly[e], xc[r]);
        const float d = v3_dot(rel, nr) - hh[r][k];
        if (d < 1e-3f) { pts[nk] = poly[e]; dep[nk] = d; ca_[nk] = v3_dot(rel, ax[r][ru]); cb_[nk] = v3_dot(rel, ax[r][rw]); ++nk; }
    }
    const v3 n = (r == 0) ? v3_neg(nr) : nr;                  /* from body1 (B) towards body0 (A) */
    if (nk <= 4) {
        for (int e = 0; e < nk; ++e) new_contact(s, i0, i1, pts[e], n, fminf(0.0f, dep[e]));
        return;
    }
    int pick[4]; int npick = 0;
    for (int dgn = 0; dgn < 4; ++dgn) {                       /* extremes along (+a+b), (-a+b), (-a-b), (+a-b) */
        const float sa = (dgn == 0 || dgn == 3) ? 1.0f : -1.0f, sb = (dgn < 2) ? 1.0f : -1.0f;
        int arg = 0; float bestv = -3.0e38f;
        for (int e = 0; e < nk; ++e) { const float val = sa * ca_[e] + sb * cb_[e]; if (val > bestv) { bestv = val; arg = e; } }
        int dup = 0;
        for (int q = 0; q < npick; ++q) dup |= (pick[q] == arg);
        if (!dup) pick[npick++] = arg;
    }
    for (int q = 0; q < npick; ++q) new_contact(s, i0, i1, pts[pick[q]], n, fminf(0.0f, dep[pick[q]]));
}

/* CollisionDetect::detectCollisions, CollisionDetect.cpp:27-76 */
/* ------------------------------------------------------------------ */
/* EXTENSION: triangle meshes as signed-distance grids. */
/* No reference counterpart (the reference has no SDF code): parity    */
/* unpinned; csrc/sdf.cuh follows these routines operation for         */
/* operation so contact lists can be compared bit for bit.             */
/* ------------------------------------------------------------------ */
int orc_add_sdf_shape(orc_system* s, int nx, int ny, int nz, const float* origin3, float cell, const float* grid,
                      int nv, const float* verts)
{
    if (s->n_shapes == s->cap_shapes) {
        s->cap_shapes = s->cap_shapes ? s->cap_shapes * 2 : 4;
        s->shapes = (sdf_shape*)realloc(s->shapes, (size_t)s->cap_shapes * sizeof(sdf_shape));
    }
    sdf_shape* S = &s->shapes[s->n_shapes];
    S->nx = nx; S->ny = ny; S->nz = nz; S->nv = nv;
    S->origin = V3(origin3[0], origin3[1], origin3[2]); S->cell = cell;
    S->grid = (float*)malloc((size_t)nx * ny * nz * sizeof(float));
    memcpy(S->grid, grid, (size_t)nx * ny * nz * sizeof(float));
    S->verts = (float*)malloc((size_t)nv * 3 * sizeof(float));
    memcpy(S->verts, verts, (size_t)nv * 3 * sizeof(float));
    S->radius = 0.f;
    S->bbmin = V3(1e30f, 1e30f, 1e30f); S->bbmax = V3(-1e30f, -1e30f, -1e30f);
    for (int i = 0; i < nv; ++i) {
        const v3 p = V3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]);
        const float r = v3_norm(p);
        if (r > S->radius) S->radius = r;
        S->bbmin = V3(fminf(S->bbmin.x, p.x), fminf(S->bbmin.y, p.y), fminf(S->bbmin.z, p.z));
        S->bbmax = V3(fmaxf(S->bbmax.x, p.x), fmaxf(S->bbmax.y, p.y), fmaxf(S->bbmax.z, p.z));
    }
    return s->n_shapes++;
}

/* closest point of triangle abc to p (Voronoi-region walk) */
static v3 closest_on_triangle(v3 p, v3 a, v3 b, v3 c)
{
    const v3 ab = v3_sub(b, a), ac = v3_sub(c, a), ap = v3_sub(p, a);
    const float d1 = v3_dot(ab, ap), d2 = v3_dot(ac, ap);
    if (d1 <= 0.f && d2 <= 0.f) return a;
    const v3 bp = v3_sub(p, b);
    const float d3 = v3_dot(ab, bp), d4 = v3_dot(ac, bp);
    if (d3 >= 0.f && d4 <= d3) return b;
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) return v3_add(a, v3_scale(d1 / (d1 - d3), ab));
    const v3 cp = v3_sub(p, c);
    const float d5 = v3_dot(ab, cp), d6 = v3_dot(ac, cp);
    if (d6 >= 0.f && d5 <= d6) return c;
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) return v3_add(a, v3_scale(d2 / (d2 - d6), ac));
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f)
        return v3_add(b, v3_scale((d4 - d3) / ((d4 - d3) + (d5 - d6)), v3_sub(c, b)));
    const float denom = 1.0f / (va + vb + vc);
    return v3_add(a, v3_add(v3_scale(vb * denom, ab), v3_scale(vc * denom, ac)));
}

/* does the ray p + t d, t > 0, cross triangle abc (Moller-Trumbore) */
static int ray_hits_triangle(v3 p, v3 d, v3 a, v3 b, v3 c)
{
    const v3 e1 = v3_sub(b, a), e2 = v3_sub(c, a);
    const v3 h = v3_cross(d, e2);
    const float det = v3_dot(e1, h);
    if (fabsf(det) < 1e-12f) return 0;
    const float f = 1.0f / det;
    const v3 sv = v3_sub(p, a);
    const float u = f * v3_dot(sv, h);
    if (u < 0.f || u > 1.f) return 0;
    const v3 q = v3_cross(sv, e1);
    const float w = f * v3_dot(d, q);
    if (w < 0.f || u + w > 1.f) return 0;
    return f * v3_dot(e2, q) > 0.f;
}

/* signed distance of every grid node to a closed triangle mesh: distance to the closest triangle, negative when a ray in
 * a fixed skew direction crosses the surface an odd number of times */
void orc_build_sdf(int nv, const float* verts, int nt, const int* tris, int nx, int ny, int nz, const float* origin3,
                   float cell, float* out)
{
    (void)nv;
    const v3 dir = V3(0.9f, 0.3137f, 0.2918f);
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
        const v3 p = V3(origin3[0] + cell * (float)i, origin3[1] + cell * (float)j, origin3[2] + cell * (float)k);
        float best = 1e30f;
        int crossings = 0;
        for (int t = 0; t < nt; ++t) {
            const float* pa = verts + 3 * tris[3 * t]; const float* pb = verts + 3 * tris[3 * t + 1]; const float* pc = verts + 3 * tris[3 * t + 2];
            const v3 a = V3(pa[0], pa[1], pa[2]), b = V3(pb[0], pb[1], pb[2]), c = V3(pc[0], pc[1], pc[2]);
            const float d2 = v3_sqnorm(v3_sub(p, closest_on_triangle(p, a, b, c)));
            if (d2 < best) best = d2;
            crossings += ray_hits_triangle(p, dir, a, b, c);
        }
        const float d = sqrtf(best);
        out[((size_t)k * ny + j) * nx + i] = (crossings & 1) ? -d : d;
    }
}

/* 'v x y z' and 'f a/b/c ...' lines as OBJLoader::load reads them (src/util/OBJLoader.cpp:34-122): first index of each
 * face token, 1-based, faces are triangles. Returns 0 on success. */
int orc_load_obj(const char* path, float* verts, int max_v, int* tris, int max_t, int* nv_out, int* nt_out)
{
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    char line[512];
    int nv = 0, nt = 0;
    while (fgets(line, sizeof line, f)) {
        if (line[0] == 'v' && line[1] == ' ') {
            float x, y, z;
            if (sscanf(line + 2, "%f %f %f", &x, &y, &z) == 3 && nv < max_v) { verts[3 * nv] = x; verts[3 * nv + 1] = y; verts[3 * nv + 2] = z; ++nv; }
        } else if (line[0] == 'f' && line[1] == ' ') {
            int idx[3], k = 0;
            for (char* tok = strtok(line + 2, " \t\r\n"); tok && k < 3; tok = strtok(NULL, " \t\r\n")) idx[k++] = atoi(tok);
            if (k == 3 && nt < max_t) { tris[3 * nt] = idx[0] - 1; tris[3 * nt + 1] = idx[1] - 1; tris[3 * nt + 2] = idx[2] - 1; ++nt; }
        }
    }
    fclose(f);
    *nv_out = nv; *nt_out = nt;
    return 0;
}

static inline float lerpf(float a, float b, float t) { return a + t * (b - a); }

/* trilinear value and unit gradient of the grid at the body-frame point pl. Outside the grid the point is clamped onto
 * the grid's box and its distance to the box is added (the surface is at least a margin inside the box) */
static void sdf_sample(const sdf_shape* S, v3 pl, float* phi, v3* grad)
{
    float fx = (pl.x - S->origin.x) / S->cell, fy = (pl.y - S->origin.y) / S->cell, fz = (pl.z - S->origin.z) / S->cell;
    const float cx = fminf(fmaxf(fx, 0.f), (float)(S->nx - 1)), cy = fminf(fmaxf(fy, 0.f), (float)(S->ny - 1)),
                cz = fminf(fmaxf(fz, 0.f), (float)(S->nz - 1));
    const float outside = S->cell * v3_norm(V3(fx - cx, fy - cy, fz - cz));
    fx = cx; fy = cy; fz = cz;
    int i = (int)fx, j = (int)fy, k = (int)fz;
    if (i > S->nx - 2) i = S->nx - 2;
    if (j > S->ny - 2) j = S->ny - 2;
    if (k > S->nz - 2) k = S->nz - 2;
    const float tx = fx - (float)i, ty = fy - (float)j, tz = fz - (float)k;
    const float* g = S->grid + ((size_t)k * S->ny + j) * S->nx + i;
    const size_t sy = (size_t)S->nx, sz = (size_t)S->nx * S->ny;
    const float c000 = g[0], c100 = g[1], c010 = g[sy], c110 = g[sy + 1];
    const float c001 = g[sz], c101 = g[sz + 1], c011 = g[sz + sy], c111 = g[sz + sy + 1];
    *phi = lerpf(lerpf(lerpf(c000, c100, tx), lerpf(c010, c110, tx), ty), lerpf(lerpf(c001, c101, tx), lerpf(c011, c111, tx), ty), tz) + outside;
    v3 gr;
    gr.x = lerpf(lerpf(c100 - c000, c110 - c010, ty), lerpf(c101 - c001, c111 - c011, ty), tz);
    gr.y = lerpf(lerpf(c010 - c000, c110 - c100, tx), lerpf(c011 - c001, c111 - c101, tx), tz);
    gr.z = lerpf(lerpf(c001 - c000, c101 - c100, tx), lerpf(c011 - c010, c111 - c110, tx), ty);
    const float n2 = v3_sqnorm(gr);
    *grad = n2 > 0.f ? v3_div(gr, sqrtf(n2)) : V3(0.f, 1.f, 0.f);
}

/* the (at most) four deepest sample points of a pair, deepest first, ties in sampling order */
typedef struct { int cnt; float phi[4]; v3 p[4], n[4]; } sdf_hits;
static void sdf_hit(sdf_hits* H, float phi, v3 p, v3 n)
{
    int at = H->cnt;
    while (at > 0 && phi < H->phi[at - 1]) --at;
    if (at >= 4) return;
    const int last = H->cnt < 4 ? H->cnt : 3;
    for (int k = last; k > at; --k) { H->phi[k] = H->phi[k - 1]; H->p[k] = H->p[k - 1]; H->n[k] = H->n[k - 1]; }
    H->phi[at] = phi; H->p[at] = p; H->n[at] = n;
    if (H->cnt < 4) ++H->cnt;
}
static void sdf_emit(orc_system* s, int b0, int b1, const sdf_hits* H)
{
    /* one normal per pair, that of the deepest point (the device's pair record carries a single normal) */
    for (int k = 0; k < H->cnt; ++k) new_contact(s, b0, b1, H->p[k], H->n[0], fminf(0.0f, H->phi[k]));
}

/* mesh (body0): vertices under the plane (body1) */
static void ext_sdf_plane(orc_system* s, int i0, int i1)
{
    const body_t* m = &s->bodies[i0];
    const body_t* pl = &s->bodies[i1];
    const sdf_shape* S = &s->shapes[m->shape];
    const v3 planen = q_rot(pl->q, pl->plane_n);
    const v3 planep = v3_add(q_rot(pl->q, pl->plane_p), pl->x);
    if (distance_point_plane(m->x, planep, planen) > S->radius + 1e-3f) return;
    sdf_hits H; H.cnt = 0;
    for (int k = 0; k < S->nv; ++k) {
        const v3 pw = v3_add(q_rot(m->q, V3(S->verts[3 * k], S->verts[3 * k + 1], S->verts[3 * k + 2])), m->x);
        const float d = distance_point_plane(pw, planep, planen);
        if (d < 1e-3f) sdf_hit(&H, d, pw, planen);
    }
    sdf_emit(s, i0, i1, &H);
}

/* sphere (body0): its centre in the grid of the mesh (body1) */
static void ext_sphere_sdf(orc_system* s, int i0, int i1)
{
    const body_t* sp = &s->bodies[i0];
    const body_t* m = &s->bodies[i1];
    const sdf_shape* S = &s->shapes[m->shape];
    const v3 rel = v3_sub(sp->x, m->x);
    if (v3_norm(rel) > S->radius + sp->radius + 1e-3f) return;
    float val; v3 g;
    sdf_sample(S, q_rot(q_inverse(m->q), rel), &val, &g);
    const float d = val - sp->radius;
    if (d < 1e-3f) {
        const v3 n = q_rot(m->q, g);
        new_contact(s, i0, i1, v3_sub(sp->x, v3_scale(sp->radius, n)), n, fminf(0.0f, d));
    }
}

/* vertices of mesh `from` sampled in the grid of mesh `in`; flip = -1 when the normal must point the other way */
static void sdf_verts_in(const orc_system* s, const body_t* from, const body_t* in, float flip, sdf_hits* H)
{
    const sdf_shape* Sf = &s->shapes[from->shape];
    const sdf_shape* Si = &s->shapes[in->shape];
    const quat qinv = q_inverse(in->q);
    for (int k = 0; k < Sf->nv; ++k) {
        const v3 pw = v3_add(q_rot(from->q, V3(Sf->verts[3 * k], Sf->verts[3 * k + 1], Sf->verts[3 * k + 2])), from->x);
        float val; v3 g;
        sdf_sample(Si, q_rot(qinv, v3_sub(pw, in->x)), &val, &g);
        if (val < 1e-3f) sdf_hit(H, val, pw, v3_scale(flip, q_rot(in->q, g)));
    }
}

/* signed distance and outward unit normal of a box of half extents h at the box-frame point c */
static float box_sdf(v3 c, v3 h, v3* n)
{
    const v3 q = V3(fminf(fmaxf(c.x, -h.x), h.x), fminf(fmaxf(c.y, -h.y), h.y), fminf(fmaxf(c.z, -h.z), h.z));
    const v3 dx = v3_sub(c, q);
    const float dist = v3_norm(dx);
    if (dist > 0.f) { *n = v3_div(dx, dist); return dist; }
    const float ex = h.x - fabsf(c.x), ey = h.y - fabsf(c.y), ez = h.z - fabsf(c.z);      /* inside: nearest face */
    if (ex <= ey && ex <= ez) { *n = V3(c.x < 0.f ? -1.f : 1.f, 0.f, 0.f); return -ex; }
    if (ey <= ez)             { *n = V3(0.f, c.y < 0.f ? -1.f : 1.f, 0.f); return -ey; }
    *n = V3(0.f, 0.f, c.z < 0.f ? -1.f : 1.f); return -ez;
}

/* mesh (body0) against box (body1): the mesh vertices against the box, then the box corners (bit 0 = +x, bit 1 = +y,
 * bit 2 = +z) in the mesh's grid */
static void ext_sdf_box(orc_system* s, int i0, int i1)
{
    const body_t* m = &s->bodies[i0];
    const body_t* bx = &s->bodies[i1];
    const sdf_shape* S = &s->shapes[m->shape];
    const v3 h = V3(0.5f * bx->dim.x, 0.5f * bx->dim.y, 0.5f * bx->dim.z);
    if (v3_norm(v3_sub(m->x, bx->x)) > S->radius + v3_norm(h) + 1e-3f) return;
    const quat qbinv = q_inverse(bx->q), qminv = q_inverse(m->q);
    sdf_hits H; H.cnt = 0;
    for (int k = 0; k < S->nv; ++k) {
        const v3 pw = v3_add(q_rot(m->q, V3(S->verts[3 * k], S->verts[3 * k + 1], S->verts[3 * k + 2])), m->x);
        v3 nl;
        const float d = box_sdf(q_rot(qbinv, v3_sub(pw, bx->x)), h, &nl);
        if (d < 1e-3f) sdf_hit(&H, d, pw, q_rot(bx->q, nl));
    }
    for (int k = 0; k < 8; ++k) {
        const v3 pw = v3_add(q_rot(bx->q, V3((k & 1) ? h.x : -h.x, (k & 2) ? h.y : -h.y, (k & 4) ? h.z : -h.z)), bx->x);
        float val; v3 g;
        sdf_sample(S, q_rot(qminv, v3_sub(pw, m->x)), &val, &g);
        if (val < 1e-3f) sdf_hit(&H, val, pw, v3_neg(q_rot(m->q, g)));
    }
    sdf_emit(s, i0, i1, &H);
}

/* mesh i0 (body0) against mesh i1 (body1): body0's vertices in body1's grid, then body1's in body0's */
static void ext_sdf_sdf(orc_system* s, int i0, int i1)
{
    const body_t* a = &s->bodies[i0];
    const body_t* b = &s->bodies[i1];
    if (v3_norm(v3_sub(a->x, b->x)) > s->shapes[a->shape].radius + s->shapes[b->shape].radius + 1e-3f) return;
    sdf_hits H; H.cnt = 0;
    sdf_verts_in(s, a, b, 1.0f, &H);       /* gradient of b's grid points out of b, towards body0 */
    sdf_verts_in(s, b, a, -1.0f, &H);      /* gradient of a's grid points out of a: flip */
    sdf_emit(s, i0, i1, &H);
}

/* ---- EXTENSION: cylinder pairs the reference ignores (sphere-cylinder, cylinder-box, cylinder-cylinder, mesh-cylinder;
 * CollisionDetect.cpp:45-73 only knows cylinder-plane). No reference behaviour exists: a finite cylinder (axis = local y,
 * Geometry.h:79-106) is an analytic signed-distance function plus CYL_SAMPLES surface points -- three rings of 16 (both
 * rims and mid height) and the two cap centres -- and the pairs reuse the mesh machinery: sample points of one shape in
 * the distance field of the other, at most the four deepest per pair, the normal of the deepest. */
enum { CYL_RING = 16, CYL_SAMPLES = 50 };
static const float CYL_COS[CYL_RING] = { 1.0f, 0.9238795f, 0.70710677f, 0.38268343f, 0.0f, -0.38268343f, -0.70710677f, -0.9238795f,
                                         -1.0f, -0.9238795f, -0.70710677f, -0.38268343f, 0.0f, 0.38268343f, 0.70710677f, 0.9238795f };
static const float CYL_SIN[CYL_RING] = { 0.0f, 0.38268343f, 0.70710677f, 0.9238795f, 1.0f, 0.9238795f, 0.70710677f, 0.38268343f,
                                         0.0f, -0.38268343f, -0.70710677f, -0.9238795f, -1.0f, -0.9238795f, -0.70710677f, -0.38268343f };
static v3 cyl_sample(int k, float h2, float r)
{
    if (k >= 3 * CYL_RING) return V3(0.f, k == 3 * CYL_RING ? -h2 : h2, 0.f);
    const int ring = k / CYL_RING, a = k - ring * CYL_RING;
    return V3(r * CYL_COS[a], ring == 0 ? -h2 : (ring == 1 ? 0.f : h2), r * CYL_SIN[a]);
}
/* signed distance and outward unit normal of the cylinder (half height h2, radius r) at the cylinder-frame point c */
static float cyl_sdf(v3 c, float h2, float r, v3* n)
{
    const float rho = sqrtf(c.x * c.x + c.z * c.z);
    const v3 radial = rho > 0.f ? V3(c.x / rho, 0.f, c.z / rho) : V3(1.f, 0.f, 0.f);
    const float dr = rho - r, dy = fabsf(c.y) - h2, sy = c.y < 0.f ? -1.f : 1.f;
    if (dr <= 0.f && dy <= 0.f) {                         /* inside: nearest of the side and the caps */
        if (dr > dy) { *n = radial; return dr; }
        *n = V3(0.f, sy, 0.f); return dy;
    }
    if (dr > 0.f && dy > 0.f) {                           /* nearest point on a rim */
        const float d = sqrtf(dr * dr + dy * dy);
        *n = V3((dr * radial.x) / d, (dy * sy) / d, (dr * radial.z) / d);
        return d;
    }
    if (dr > 0.f) { *n = radial; return dr; }
    *n = V3(0.f, sy, 0.f); return dy;
}
static inline float cyl_bound(const body_t* c) { const float h2 = 0.5f * c->height; return sqrtf(h2 * h2 + c->radius * c->radius); }

/* sphere (body0) against cylinder (body1) */
static void ext_sphere_cyl(orc_system* s, int i0, int i1)
{
    const body_t* sp = &s->bodies[i0];
    const body_t* cy = &s->bodies[i1];
    const v3 rel = v3_sub(sp->x, cy->x);
    if (v3_norm(rel) > sp->radius + cyl_bound(cy) + 1e-3f) return;
    v3 nl;
    const float d = cyl_sdf(q_rot(q_inverse(cy->q), rel), 0.5f * cy->height, cy->radius, &nl) - sp->radius;
    if (d < 1e-3f) {
        const v3 n = q_rot(cy->q, nl);
        new_contact(s, i0, i1, v3_sub(sp->x, v3_scale(sp->radius, n)), n, fminf(0.0f, d));
    }
}

/* surface points of cylinder `from` in the distance field of cylinder `in`; flip as in sdf_verts_in */
static void cyl_samples_in_cyl(const body_t* from, const body_t* in, float flip, sdf_hits* H)
{
    const quat qinv = q_inverse(in->q);
    for (int k = 0; k < CYL_SAMPLES; ++k) {
        const v3 pw = v3_add(q_rot(from->q, cyl_sample(k, 0.5f * from->height, from->radius)), from->x);
        v3 nl;
        const float d = cyl_sdf(q_rot(qinv, v3_sub(pw, in->x)), 0.5f * in->height, in->radius, &nl);
        if (d < 1e-3f) sdf_hit(H, d, pw, v3_scale(flip, q_rot(in->q, nl)));
    }
}

/* cylinder (body0) against cylinder (body1) */
static void ext_cyl_cyl(orc_system* s, int i0, int i1)
{
    const body_t* a = &s->bodies[i0];
    const body_t* b = &s->bodies[i1];
    if (v3_norm(v3_sub(a->x, b->x)) > cyl_bound(a) + cyl_bound(b) + 1e-3f) return;
    sdf_hits H; H.cnt = 0;
    cyl_samples_in_cyl(a, b, 1.0f, &H);
    cyl_samples_in_cyl(b, a, -1.0f, &H);
    sdf_emit(s, i0, i1, &H);
}

/* cylinder (body0) against box (body1): the cylinder's surface points against the box, then the box corners in the cylinder */
static void ext_cyl_box(orc_system* s, int i0, int i1)
{
    const body_t* cy = &s->bodies[i0];
    const body_t* bx = &s->bodies[i1];
    const v3 h = V3(0.5f * bx->dim.x, 0.5f * bx->dim.y, 0.5f * bx->dim.z);
    if (v3_norm(v3_sub(cy->x, bx->x)) > cyl_bound(cy) + v3_norm(h) + 1e-3f) return;
    const quat qbinv = q_inverse(bx->q), qcinv = q_inverse(cy->q);
    sdf_hits H; H.cnt = 0;
    for (int k = 0; k < CYL_SAMPLES; ++k) {
        const v3 pw = v3_add(q_rot(cy->q, cyl_sample(k, 0.5f * cy->height, cy->radius)), cy->x);
        v3 nl;
        const float d = box_sdf(q_rot(qbinv, v3_sub(pw, bx->x)), h, &nl);
        if (d < 1e-3f) sdf_hit(&H, d, pw, q_rot(bx->q, nl));
    }
    for (int k = 0; k < 8; ++k) {
        const v3 pw = v3_add(q_rot(bx->q, V3((k & 1) ? h.x : -h.x, (k & 2) ? h.y : -h.y, (k & 4) ? h.z : -h.z)), bx->x);
        v3 nl;
        const float d = cyl_sdf(q_rot(qcinv, v3_sub(pw, cy->x)), 0.5f * cy->height, cy->radius, &nl);
        if (d < 1e-3f) sdf_hit(&H, d, pw, v3_neg(q_rot(cy->q, nl)));
    }
    sdf_emit(s, i0, i1, &H);
}

/* mesh (body0) against cylinder (body1): the mesh vertices against the cylinder, then the cylinder's surface points in the grid */
static void ext_sdf_cyl(orc_system* s, int i0, int i1)
{
    const body_t* m = &s->bodies[i0];
    const body_t* cy = &s->bodies[i1];
    const sdf_shape* S = &s->shapes[m->shape];
    if (v3_norm(v3_sub(m->x, cy->x)) > S->radius + cyl_bound(cy) + 1e-3f) return;
    const quat qcinv = q_inverse(cy->q), qminv = q_inverse(m->q);
    sdf_hits H; H.cnt = 0;
    for (int k = 0; k < S->nv; ++k) {
        const v3 pw = v3_add(q_rot(m->q, V3(S->verts[3 * k], S->verts[3 * k + 1], S->verts[3 * k + 2])), m->x);
        v3 nl;
        const float d = cyl_sdf(q_rot(qcinv, v3_sub(pw, cy->x)), 0.5f * cy->height, cy->radius, &nl);
        if (d < 1e-3f) sdf_hit(&H, d, pw, q_rot(cy->q, nl));
    }
    for (int k = 0; k < CYL_SAMPLES; ++k) {
        const v3 pw = v3_add(q_rot(cy->q, cyl_sample(k, 0.5f * cy->height, cy->radius)), cy->x);
        float val; v3 g;
        sdf_sample(S, q_rot(qminv, v3_sub(pw, m->x)), &val, &g);
        if (val < 1e-3f) sdf_hit(&H, val, pw, v3_neg(q_rot(m->q, g)));
    }
    sdf_emit(s, i0, i1, &H);
}

void orc_stage_detect(orc_system* s)
{
    s->pairs_tested = 0;
    if (!s->collisions) return;                         /* RigidBodySystem.cpp:77 */
    const int n = s->n_bodies;
    for (int i = 0; i < n; ++i) {
        for (int j = i + 1; j < n; ++j) {
            const body_t* body0 = &s->bodies[i];
            const body_t* body1 = &s->bodies[j];
            if (body0->fixed && body1->fixed) continue;
            s->pairs_tested++;
            const int t0 = body0->geom_type, t1 = body1->geom_type;
            if (t0 == ORC_SPHERE && t1 == ORC_SPHERE)        detect_sphere_sphere(s, i, j);
            else if (t0 == ORC_SPHERE && t1 == ORC_BOX)      detect_sphere_box(s, i, j);
            else if (t1 == ORC_SPHERE && t0 == ORC_BOX)      detect_sphere_box(s, j, i);
            else if (t0 == ORC_CYLINDER && t1 == ORC_PLANE)  detect_cylinder_plane(s, i, j);
            else if (t1 == ORC_CYLINDER && t0 == ORC_PLANE)  detect_cylinder_plane(s, j, i);
            else if (s->ext_pairs) {                          /* extension pairs, after the reference's chain */
                if (t0 == ORC_SPHERE && t1 == ORC_PLANE)      ext_sphere_plane(s, i, j);
                else if (t1 == ORC_SPHERE && t0 == ORC_PLANE) ext_sphere_plane(s, j, i);
                else if (t0 == ORC_BOX && t1 == ORC_PLANE)    ext_box_plane(s, i, j);
                else if (t1 == ORC_BOX && t0 == ORC_PLANE)    ext_box_plane(s, j, i);
                else if (t0 == ORC_BOX && t1 == ORC_BOX)      ext_box_box(s, i, j);
                else if (t0 == ORC_SDF && t1 == ORC_PLANE)    ext_sdf_plane(s, i, j);
                else if (t1 == ORC_SDF && t0 == ORC_PLANE)    ext_sdf_plane(s, j, i);
                else if (t0 == ORC_SPHERE && t1 == ORC_SDF)   ext_sphere_sdf(s, i, j);
                else if (t1 == ORC_SPHERE && t0 == ORC_SDF)   ext_sphere_sdf(s, j, i);
                else if (t0 == ORC_SDF && t1 == ORC_SDF)      ext_sdf_sdf(s, i, j);
                else if (t0 == ORC_SDF && t1 == ORC_BOX)      ext_sdf_box(s, i, j);
                else if (t1 == ORC_SDF && t0 == ORC_BOX)      ext_sdf_box(s, j, i);
                else if (t0 == ORC_SPHERE && t1 == ORC_CYLINDER)   ext_sphere_cyl(s, i, j);
                else if (t1 == ORC_SPHERE && t0 == ORC_CYLINDER)   ext_sphere_cyl(s, j, i);
                else if (t0 == ORC_CYLINDER && t1 == ORC_BOX)      ext_cyl_box(s, i, j);
                else if (t1 == ORC_CYLINDER && t0 == ORC_BOX)      ext_cyl_box(s, j, i);
                else if (t0 == ORC_CYLINDER && t1 == ORC_CYLINDER) ext_cyl_cyl(s, i, j);
                else if (t0 == ORC_SDF && t1 == ORC_CYLINDER)      ext_sdf_cyl(s, i, j);
                else if (t1 == ORC_SDF && t0 == ORC_CYLINDER)      ext_sdf_cyl(s, j, i);
            }
        }
    }
}

/* ------------------------------------------------------------------ */
/* S10-S13: frames and Jacobians                                       */
/* ------------------------------------------------------------------ */
/* Contact::computeContactFrame, Contact.cpp:33-53 */
static void contact_frame(con_t* c)
{
    c->t = v3_cross(c->n, V3(1, 0, 0));
    if (v3_norm(c->t) < 1e-5f) c->t = v3_cross(c->n, V3(0, 1, 0));
    c->t = v3_normalized(c->t);
    c->b = v3_cross(c->n, c->t);
    c->b = v3_normalized(c->b);
}

static inline void set_row(float* row, v3 lin, v3 ang)
{
    row[0] = lin.x; row[1] = lin.y; row[2] = lin.z; row[3] = ang.x; row[4] = ang.y; row[5] = ang.z;
}

/* J*Minv blocks: linear = (1/mass) * J_lin ; angular = J_ang * Iinv (Contact.cpp:90-93, Hinge.cpp:59-62) */
static void compute_JMinv(int dim, float J[5][6], float JM[5][6], const body_t* b)
{
    const float im = 1.0f / b->mass;
    for (int r = 0; r < dim; ++r) {
        for (int c = 0; c < 3; ++c) JM[r][c] = im * J[r][c];
        for (int c = 0; c < 3; ++c)
            JM[r][3 + c] = sum3(J[r][3] * b->Iinv.m[0][c], J[r][4] * b->Iinv.m[1][c], J[r][5] * b->Iinv.m[2][c]);
    }
}

/* Contact::computeJacobian, Contact.cpp:55-94 */
static void contact_jacobian(orc_system* s, con_t* c)
{
    const body_t* body0 = &s->bodies[c->body0];
    const body_t* body1 = &s->bodies[c->body1];
    const v3 rr0 = v3_sub(c->p, body0->x);
    const v3 rr1 = v3_sub(c->p, body1->x);
    memset(c->J0, 0, sizeof c->J0); memset(c->J1, 0, sizeof c->J1);
    memset(c->J0Minv, 0, sizeof c->J0Minv); memset(c->J1Minv, 0, sizeof c->J1Minv);
    memset(c->lambda, 0, sizeof c->lambda); memset(c->phi, 0, sizeof c->phi);
    c->phi[0] = c->pene;
    const v3 d[3] = { c->n, c->t, c->b };
    for (int k = 0; k < 3; ++k) {
        set_row(c->J0[k], d[k], v3_cross(rr0, d[k]));
        set_row(c->J1[k], v3_neg(d[k]), v3_neg(v3_cross(rr1, d[k])));
    }
    compute_JMinv(3, c->J0, c->J0Minv, body0);
    compute_JMinv(3, c->J1, c->J1Minv, body1);
}

/* hat(), Hinge.cpp:6-13 */
static void set_hat(float J[5][6], v3 v)
{
    J[0][3] = 0.f;   J[0][4] = -v.z; J[0][5] = v.y;
    J[1][3] = v.z;   J[1][4] = 0.f;  J[1][5] = -v.x;
    J[2][3] = -v.y;  J[2][4] = v.x;  J[2][5] = 0.f;
}

/* Hinge::computeJacobian, Hinge.cpp:33-63 ; Spherical::computeJacobian, Spherical.cpp:34-52 */
static void joint_jacobian(orc_system* s, con_t* j)
{
    const body_t* body0 = &s->bodies[j->body0];
    const body_t* body1 = &s->bodies[j->body1];
    const v3 rr0 = q_rot(body0->q, j->r0);
    const v3 rr1 = q_rot(body1->q, j->r1);
    /* phi(0:3) = x0 + rr0 - x1 - rr1 (left to right) */
    const v3 e = v3_sub(v3_sub(v3_add(body0->x, rr0), body1->x), rr1);
    j->phi[0] = e.x; j->phi[1] = e.y; j->phi[2] = e.z;
    /* Rows are rewritten in place each step; entries never written stay 0 from construction. */
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
        j->J0[r][c] = (r == c) ? 1.f : 0.f;
        j->J1[r][c] = (r == c) ? -1.f : -0.f;     /* J1.block(0,0,3,3) = -sEye: off-diagonals are -0 */
    }
    set_hat(j->J0, v3_neg(rr0));
    set_hat(j->J1, rr1);
    if (j->type == ORC_HINGE) {
        const v3 nn = q_rot(body0->q, q_rot(j->q0, V3(1, 0, 0)));
        const v3 uu = q_rot(body1->q, q_rot(j->q1, V3(0, 1, 0)));
        const v3 vv = q_rot(body1->q, q_rot(j->q1, V3(0, 0, 1)));
        const v3 nxu = v3_cross(nn, uu);
        const v3 nxv = v3_cross(nn, vv);
        j->phi[3] = v3_dot(nn, uu);
        j->phi[4] = v3_dot(nn, vv);
        set_row(j->J0[3], V3(0, 0, 0), nxu);
        set_row(j->J0[4], V3(0, 0, 0), nxv);
        set_row(j->J1[3], V3(0, 0, 0), v3_neg(nxu));
        set_row(j->J1[4], V3(0, 0, 0), v3_neg(nxv));
    }
    compute_JMinv(j->dim, j->J0, j->J0Minv, body0);
    compute_JMinv(j->dim, j->J1, j->J1Minv, body1);
}

/* CollisionDetect::computeContactJacobians (CollisionDetect.cpp:78-85), joints (RigidBodySystem.cpp:83-86),
 * zero fc/tauc (:88-92) */
void orc_stage_jacobians(orc_system* s)
{
    if (s->collisions) {
        for (int i = 0; i < s->n_contacts; ++i) {
            contact_frame(&s->contacts[i]);
            contact_jacobian(s, &s->contacts[i]);
        }
    }
    for (int i = 0; i < s->n_joints; ++i) joint_jacobian(s, &s->joints[i]);
    for (int i = 0; i < s->n_bodies; ++i) { s->bodies[i].fc = V3(0, 0, 0); s->bodies[i].tauc = V3(0, 0, 0); }
}

/* ------------------------------------------------------------------ */
/* S14-S18: Box-PGS (src/solvers/SolverBoxPGS.cpp)                     */
/* ------------------------------------------------------------------ */
/* multAndSub, SolverBoxPGS.cpp:13-21: b -= (a * G.col(k)) * x(k) */
static void mult_and_sub(int dim, float G[5][6], v3 x, v3 y, float a, float* b)
{
    const float xs[6] = { x.x, x.y, x.z, y.x, y.y, y.z };
    for (int k = 0; k < 6; ++k)
        for (int r = 0; r < dim; ++r)
            b[r] -= (a * G[r][k]) * xs[k];
}

/* buildRHS, SolverBoxPGS.cpp:27-44 */
static void build_rhs(const orc_system* s, con_t* j, float h, float* b)
{
    const float hinv = 1.0f / h;
    const float gamma = 0.3f;
    const body_t* body0 = &s->bodies[j->body0];
    const body_t* body1 = &s->bodies[j->body1];
    for (int r = 0; r < j->dim; ++r) b[r] = (-hinv * gamma) * j->phi[r];
    if (!body0->fixed) {
        mult_and_sub(j->dim, j->J0Minv, body0->f, body0->tau, h, b);
        mult_and_sub(j->dim, j->J0, body0->xdot, body0->omega, 1.0f, b);
    }
    if (!body1->fixed) {
        mult_and_sub(j->dim, j->J1Minv, body1->f, body1->tau, h, b);
        mult_and_sub(j->dim, j->J1, body1->xdot, body1->omega, 1.0f, b);
    }
}

/* b -= JMinv * (Jother^T * lambda_other) */
static void sub_coupling(int dim, float JMinv[5][6], const con_t* o, float Jo[5][6], float* b)
{
    float t[6];
    for (int c = 0; c < 6; ++c) {
        float acc = 0.f;
        for (int k = 0; k < o->dim; ++k) acc += Jo[k][c] * o->lambda[k];
        t[c] = acc;
    }
    for (int r = 0; r < dim; ++r) {
        float acc = 0.f;
        for (int c = 0; c < 6; ++c) acc += JMinv[r][c] * t[c];
        b[r] -= acc;
    }
}

/* accumulateCoupledContactsAndJoints, SolverBoxPGS.cpp:49-82 */
static void accumulate_coupled(orc_system* s, int self_kind, int self_idx, int dim,
                               float JMinv[5][6], int body, float* b)
{
    body_t* bd = &s->bodies[body];
    if (bd->fixed) return;
    for (int k = 0; k < bd->n_contacts; ++k) {
        const int ci = bd->contacts[k];
        if (self_kind == 0 && ci == self_idx) continue;
        con_t* cc = &s->contacts[ci];
        if (body == cc->body0) sub_coupling(dim, JMinv, cc, cc->J0, b);
        else                   sub_coupling(dim, JMinv, cc, cc->J1, b);
    }
    for (int k = 0; k < bd->n_joints; ++k) {
        const int ji = bd->joints[k];
        if (self_kind == 1 && ji == self_idx) continue;
        con_t* jj = &s->joints[ji];
        if (body == jj->body0) sub_coupling(dim, JMinv, jj, jj->J0, b);
        else                   sub_coupling(dim, JMinv, jj, jj->J1, b);
    }
}

/* solveContact, SolverBoxPGS.cpp:94-113 */
static void solve_contact(const float A[3][3], const float* b, float* x, float mu)
{
    x[0] = (b[0] - A[0][1] * x[1] - A[0][2] * x[2]) / (A[0][0] + 1e-3f);
    if (x[0] < 0.0f) x[0] = 0.0f;
    const float lowerx = -mu * x[0];
    const float upperx = mu * x[0];
    x[1] = (b[1] - A[1][0] * x[0] - A[1][2] * x[2]) / A[1][1];
    if (x[1] < lowerx) x[1] = lowerx; else if (x[1] > upperx) x[1] = upperx;
    x[2] = (b[2] - A[2][0] * x[0] - A[2][1] * x[1]) / A[2][2];
    if (x[2] < lowerx) x[2] = lowerx; else if (x[2] > upperx) x[2] = upperx;
}

/* Eigen::LDLT<MatrixXf> restated (Cholesky/LDLT.h): symmetric diagonal pivoting on the
 * largest |diagonal| entry, in-place lower factorisation; solve = P^T L^-T D^-1 L^-1 P b
 * with Eigen's pseudo-inverse tolerance on D. */
typedef struct { int n; float m[5][5]; int tr[5]; } ldlt_t;

static void ldlt_factor(ldlt_t* F, int n, float A[5][5])
{
    F->n = n;
    memcpy(F->m, A, sizeof F->m);
    float (*M)[5] = F->m;
    for (int k = 0; k < n; ++k) {
        int piv = k; float big = fabsf(M[k][k]);
        for (int i = k + 1; i < n; ++i) if (fabsf(M[i][i]) > big) { big = fabsf(M[i][i]); piv = i; }
        F->tr[k] = piv;
        if (piv != k) {
            /* symmetric swap of rows/cols k and piv in the lower triangle */
            for (int c = 0; c < k; ++c) { float t = M[k][c]; M[k][c] = M[piv][c]; M[piv][c] = t; }
            for (int r = piv + 1; r < n; ++r) { float t = M[r][k]; M[r][k] = M[r][piv]; M[r][piv] = t; }
            { float t = M[k][k]; M[k][k] = M[piv][piv]; M[piv][piv] = t; }
            for (int i = k + 1; i < piv; ++i) { float t = M[i][k]; M[i][k] = M[piv][i]; M[piv][i] = t; }
        }
        /* A10 = M[k][0:k], A20 = M[k+1:n][0:k], A21 = M[k+1:n][k] */
        if (k > 0) {
            float temp[5];
            for (int c = 0; c < k; ++c) temp[c] = M[c][c] * M[k][c];
            float acc = 0.f;
            for (int c = 0; c < k; ++c) acc += M[k][c] * temp[c];
            M[k][k] -= acc;
            for (int r = k + 1; r < n; ++r) {
                float a2 = 0.f;
                for (int c = 0; c < k; ++c) a2 += M[r][c] * temp[c];
                M[r][k] -= a2;
            }
        }
        const float piv_val = M[k][k];
        if (fabsf(piv_val) > 0.0f)
            for (int r = k + 1; r < n; ++r) M[r][k] /= piv_val;
    }
}

static void ldlt_solve(const ldlt_t* F, const float* b, float* x)
{
    const int n = F->n;
    float y[5];
    for (int i = 0; i < n; ++i) y[i] = b[i];
    for (int k = 0; k < n; ++k) { const int p = F->tr[k]; if (p != k) { float t = y[k]; y[k] = y[p]; y[p] = t; } }
    for (int i = 0; i < n; ++i) { float acc = y[i]; for (int c = 0; c < i; ++c) acc -= F->m[i][c] * y[c]; y[i] = acc; }
    const float tol = 1.17549435e-38f;   /* (std::numeric_limits<float>::min)() */
    for (int i = 0; i < n; ++i) { if (fabsf(F->m[i][i]) > tol) y[i] /= F->m[i][i]; else y[i] = 0.f; }
    for (int i = n - 1; i >= 0; --i) { float acc = y[i]; for (int r = i + 1; r < n; ++r) acc -= F->m[r][i] * y[r]; y[i] = acc; }
    for (int k = n - 1; k >= 0; --k) { const int p = F->tr[k]; if (p != k) { float t = y[k]; y[k] = y[p]; y[p] = t; } }
    for (int i = 0; i < n; ++i) x[i] = y[i];
}

/* A += JMinv * J^T  (dim x dim) */
static void add_JMinvJT(int dim, float JM[5][6], float J[5][6], float A[5][5])
{
    for (int r = 0; r < dim; ++r)
        for (int c = 0; c < dim; ++c) {
            float acc = 0.f;
            for (int k = 0; k < 6; ++k) acc += JM[r][k] * J[c][k];
            A[r][c] += acc;
        }
}

/* SolverBoxPGS::solve, SolverBoxPGS.cpp:126-250 */
static void solve_box_pgs(orc_system* s, float h)
{
    const int numContacts = s->n_contacts;
    const int numJoints = s->n_joints;
    const int N = numJoints + numContacts;
    ldlt_t* LLT = numJoints ? (ldlt_t*)malloc((size_t)numJoints * sizeof(ldlt_t)) : NULL;
    float (*Acc)[3][3] = numContacts ? (float (*)[3][3])malloc((size_t)numContacts * sizeof(float[3][3])) : NULL;

    for (int i = 0; i < numJoints; ++i) {                                   /* :134-163 */
        con_t* j = &s->joints[i];
        float A[5][5]; memset(A, 0, sizeof A);
        for (int d = 0; d < j->dim; ++d) A[d][d] = 1e-5f;
        if (!s->bodies[j->body0].fixed) add_JMinvJT(j->dim, j->J0Minv, j->J0, A);
        if (!s->bodies[j->body1].fixed) add_JMinvJT(j->dim, j->J1Minv, j->J1, A);
        ldlt_factor(&LLT[i], j->dim, A);
    }
    for (int i = 0; i < numContacts; ++i) {                                 /* :165-189 */
        con_t* c = &s->contacts[i];
        float A[5][5]; memset(A, 0, sizeof A);
        if (!s->bodies[c->body0].fixed) add_JMinvJT(3, c->J0Minv, c->J0, A);
        if (!s->bodies[c->body1].fixed) add_JMinvJT(3, c->J1Minv, c->J1, A);
        for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) Acc[i][r][cc] = A[r][cc];
    }
    if (N > 0) {
        float (*b)[5] = (float (*)[5])calloc((size_t)N, sizeof(float[5]));
        for (int i = 0; i < numJoints; ++i) build_rhs(s, &s->joints[i], h, b[i]);          /* :200-204 */
        for (int i = 0; i < numContacts; ++i) {                                              /* :206-211 */
            build_rhs(s, &s->contacts[i], h, b[i + numJoints]);
            s->contacts[i].lambda[0] = s->contacts[i].lambda[1] = s->contacts[i].lambda[2] = 0.f;
        }
        for (int iter = 0; iter < s->solverIter; ++iter) {                                   /* :217-248 */
            for (int i = 0; i < numJoints; ++i) {
                con_t* j = &s->joints[i];
                float x[5];
                for (int r = 0; r < j->dim; ++r) x[r] = b[i][r];
                accumulate_coupled(s, 1, i, j->dim, j->J0Minv, j->body0, x);
                accumulate_coupled(s, 1, i, j->dim, j->J1Minv, j->body1, x);
                ldlt_solve(&LLT[i], x, j->lambda);
            }
            for (int i = 0; i < numContacts; ++i) {
                con_t* c = &s->contacts[i];
                float x[5];
                for (int r = 0; r < 3; ++r) x[r] = b[i + numJoints][r];
                accumulate_coupled(s, 0, i, 3, c->J0Minv, c->body0, x);
                accumulate_coupled(s, 0, i, 3, c->J1Minv, c->body1, x);
                solve_contact(Acc[i], x, c->lambda, s->mu);
            }
        }
        free(b);
    }
    free(LLT); free(Acc);
}

/* ------------------------------------------------------------------ */
/* S23: CG / CR joint-only solvers                                     */
/* ------------------------------------------------------------------ */
static int assign_joint_idx(orc_system* s)
{
    unsigned idx = 0;
    for (int i = 0; i < s->n_joints; ++i) { s->joints[i].idx = idx; idx += (unsigned)s->joints[i].dim; }
    return (int)idx;
}

/* buildRHS, SolverConjGradient.cpp:26-47 / SolverConjResidual.cpp:25-46 */
static void krylov_build_rhs(orc_system* s, float h, float* b)
{
    for (int i = 0; i < s->n_joints; ++i) build_rhs(s, &s->joints[i], h, b + s->joints[i].idx);
}

/* Ax.segment(j) += JMinv * (Jo^T * x.segment(o)) */
static void add_coupling(con_t* j, float JMinv[5][6], con_t* o, float Jo[5][6], const float* x, float* Ax)
{
    float t[6];
    for (int c = 0; c < 6; ++c) {
        float acc = 0.f;
        for (int k = 0; k < o->dim; ++k) acc += Jo[k][c] * x[o->idx + k];
        t[c] = acc;
    }
    for (int r = 0; r < j->dim; ++r) {
        float acc = 0.f;
        for (int c = 0; c < 6; ++c) acc += JMinv[r][c] * t[c];
        Ax[j->idx + r] += acc;
    }
}

/* computeAx, SolverConjGradient.cpp:49-100 (eps = 0) and SolverConjResidual.cpp:68-92 (eps = 1e-9) */
static void krylov_compute_Ax(orc_system* s, const float* x, float* Ax, int n, float eps, int use_eps)
{
    for (int i = 0; i < n; ++i) Ax[i] = 0.f;
    for (int i = 0; i < s->n_joints; ++i) {
        con_t* j = &s->joints[i];
        body_t* body0 = &s->bodies[j->body0];
        body_t* body1 = &s->bodies[j->body1];
        if (use_eps) for (int r = 0; r < j->dim; ++r) Ax[j->idx + r] += eps * x[j->idx + r];
        if (!body0->fixed) {
            add_coupling(j, j->J0Minv, j, j->J0, x, Ax);
            for (int k = 0; k < body0->n_joints; ++k) {
                const int ji = body0->joints[k];
                if (ji == i) continue;
                con_t* jj = &s->joints[ji];
                add_coupling(j, j->J0Minv, jj, (j->body0 == jj->body0) ? jj->J0 : jj->J1, x, Ax);
            }
        }
        if (!body1->fixed) {
            add_coupling(j, j->J1Minv, j, j->J1, x, Ax);
            for (int k = 0; k < body1->n_joints; ++k) {
                const int ji = body1->joints[k];
                if (ji == i) continue;
                con_t* jj = &s->joints[ji];
                add_coupling(j, j->J1Minv, jj, (j->body1 == jj->body0) ? jj->J0 : jj->J1, x, Ax);
            }
        }
    }
}

static float dotn(const float* a, const float* b, int n) { float acc = 0.f; for (int i = 0; i < n; ++i) acc += a[i] * b[i]; return acc; }

/* SolverConjGradient::solve, SolverConjGradient.cpp:109-162 (keeps p = r - beta*p, :150) */
static void solve_conj_gradient(orc_system* s, float h)
{
    const int n = assign_joint_idx(s);
    if (n == 0) return;
    float* buf = (float*)calloc((size_t)n * 6, sizeof(float));
    float *x = buf, *r = buf + n, *p = buf + 2 * n, *b = buf + 3 * n, *Ax = buf + 4 * n, *Ap = buf + 5 * n;
    krylov_build_rhs(s, h, b);
    krylov_compute_Ax(s, x, Ax, n, 0.f, 0);
    for (int i = 0; i < n; ++i) { r[i] = b[i] - Ax[i]; p[i] = r[i]; }
    float rTr = dotn(r, r, n);
    krylov_compute_Ax(s, p, Ap, n, 0.f, 0);
    float pAp = dotn(p, Ap, n);
    for (int it = 0; it < s->solverIter; ++it) {
        if (rTr < 1e-12f) break;
        if (pAp < 1e-12f) break;
        const float alpha = rTr / pAp;
        for (int i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; }
        const float rTr_next = dotn(r, r, n);
        const float beta = rTr_next / rTr;
        for (int i = 0; i < n; ++i) p[i] = r[i] - beta * p[i];
        krylov_compute_Ax(s, p, Ap, n, 0.f, 0);
        pAp = dotn(p, Ap, n);
        rTr = rTr_next;
    }
    for (int i = 0; i < s->n_joints; ++i)
        for (int k = 0; k < s->joints[i].dim; ++k) s->joints[i].lambda[k] = x[s->joints[i].idx + k];
    free(buf);
}

/* SolverConjResidual::solve, SolverConjResidual.cpp:100-158 */
static void solve_conj_residual(orc_system* s, float h)
{
    const int n = assign_joint_idx(s);
    if (n == 0) return;
    float* buf = (float*)calloc((size_t)n * 7, sizeof(float));
    float *x = buf, *r = buf + n, *p = buf + 2 * n, *b = buf + 3 * n, *Ax = buf + 4 * n, *Ap = buf + 5 * n, *Ar = buf + 6 * n;
    const float eps = 1e-9f;
    krylov_build_rhs(s, h, b);
    krylov_compute_Ax(s, x, Ax, n, eps, 1);
    for (int i = 0; i < n; ++i) { r[i] = b[i] - Ax[i]; p[i] = r[i]; }
    krylov_compute_Ax(s, p, Ap, n, eps, 1);
    krylov_compute_Ax(s, r, Ar, n, eps, 1);
    float rAr = dotn(r, Ar, n);
    float pATAp = dotn(Ap, Ap, n);
    for (int it = 0; it < s->solverIter; ++it) {
        if (rAr < 1e-12f) break;
        if (pATAp < 1e-12f) break;
        const float alpha = rAr / pATAp;
        for (int i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; }
        krylov_compute_Ax(s, r, Ar, n, eps, 1);
        const float rAr_next = dotn(r, Ar, n);
        const float beta = rAr_next / rAr;
        for (int i = 0; i < n; ++i) { p[i] = r[i] + beta * p[i]; Ap[i] = Ar[i] + beta * Ap[i]; }
        rAr = rAr_next;
        pATAp = dotn(Ap, Ap, n);
    }
    for (int i = 0; i < s->n_joints; ++i)
        for (int k = 0; k < s->joints[i].dim; ++k) s->joints[i].lambda[k] = x[s->joints[i].idx + k];
    free(buf);
}


/* The solver's per-constraint diagonal blocks and right-hand sides as SolverBoxPGS::solve builds
 * them before its main loop (SolverBoxPGS.cpp:134-211); exported for per-stage parity tests. */
void orc_get_solver_blocks(orc_system* s, float h, float* Acontact9, float* rhs_contact3, float* Ajoint25, float* rhs_joint5)
{
    for (int i = 0; i < s->n_contacts; ++i) {
        con_t* c = &s->contacts[i];
        float A[5][5]; memset(A, 0, sizeof A);
        if (!s->bodies[c->body0].fixed) add_JMinvJT(3, c->J0Minv, c->J0, A);
        if (!s->bodies[c->body1].fixed) add_JMinvJT(3, c->J1Minv, c->J1, A);
        float b[5] = { 0, 0, 0, 0, 0 };
        build_rhs(s, c, h, b);
        for (int r = 0; r < 3; ++r) {
            for (int k = 0; k < 3; ++k) if (Acontact9) Acontact9[9 * i + 3 * r + k] = A[r][k];
            if (rhs_contact3) rhs_contact3[3 * i + r] = b[r];
        }
    }
    for (int i = 0; i < s->n_joints; ++i) {
        con_t* j = &s->joints[i];
        float A[5][5]; memset(A, 0, sizeof A);
        for (int d = 0; d < j->dim; ++d) A[d][d] = 1e-5f;
        if (!s->bodies[j->body0].fixed) add_JMinvJT(j->dim, j->J0Minv, j->J0, A);
        if (!s->bodies[j->body1].fixed) add_JMinvJT(j->dim, j->J1Minv, j->J1, A);
        float b[5] = { 0, 0, 0, 0, 0 };
        build_rhs(s, j, h, b);
        for (int r = 0; r < 5; ++r) {
            for (int k = 0; k < 5; ++k) if (Ajoint25) Ajoint25[25 * i + 5 * r + k] = A[r][k];
            if (rhs_joint5) rhs_joint5[5 * i + r] = r < j->dim ? b[r] : 0.f;
        }
    }
}

/* ------------------------------------------------------------------ */
/* EXTENSION: block principal pivoting, solverId 3.    */
/* No reference counterpart: parity unpinned. The CUDA kernel k_bpp    */
/* follows this function operation for operation.                      */
/*                                                                     */
/* Unknowns: the scalar rows of all joints (insertion order, bilateral)*/
/* then of all contacts (detection order; normal, two friction rows).  */
/* A = J Minv J^T over non-fixed bodies + eps I (1e-5 on joint rows as */
/* in SolverBoxPGS.cpp:143, 1e-3 on contact rows), b = buildRHS.       */
/* Solve the box LCP  w = A lambda - b,  lo <= lambda <= hi,           */
/*   lambda = lo => w >= 0, lambda = hi => w <= 0, else w = 0          */
/* with lo/hi = 0/inf (normal), -/+ inf (joints), -/+ mu lambda_n      */
/* (friction) by Judice-Pires block pivoting: each step solves the     */
/* free rows exactly (dense Cholesky), then every infeasible row       */
/* changes set; after 3 steps without a new minimum of infeasible rows */
/* only the highest infeasible row is swapped. The friction bounds are */
/* held constant while the sets settle (first pass: 0, frictionless),  */
/* then refreshed from the pass's normal impulses, until they move by  */
/* less than 1e-5 (1 + bound). At most 5 * solverIter pivot steps over */
/* all passes; the result is projected onto the bounds.                */
/* ------------------------------------------------------------------ */
enum { BPP_FREE = 0, BPP_LOWER = 1, BPP_UPPER = 2 };
enum { BPP_BILATERAL = 0, BPP_NORMAL = 1, BPP_FRICTION = 2 };
#define BPP_TOL 1e-6f

typedef struct {
    float J0[6], J1[6], M0[6], M1[6];
    int   b0, b1;                /* body index, -1 when fixed */
    int   kind, parent;          /* parent: the contact's normal row */
    float rhs, eps;
} bpp_row;

static int bpp_rows(orc_system* s, float h, bpp_row** out)
{
    int n = 3 * s->n_contacts;
    for (int i = 0; i < s->n_joints; ++i) n += s->joints[i].dim;
    bpp_row* R = (bpp_row*)calloc((size_t)(n ? n : 1), sizeof(bpp_row));
    int k = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const int cnt = pass == 0 ? s->n_joints : s->n_contacts;
        for (int i = 0; i < cnt; ++i) {
            con_t* c = pass == 0 ? &s->joints[i] : &s->contacts[i];
            float b[5] = { 0, 0, 0, 0, 0 };
            build_rhs(s, c, h, b);
            const int first = k;
            for (int r = 0; r < c->dim; ++r, ++k) {
                memcpy(R[k].J0, c->J0[r], sizeof R[k].J0); memcpy(R[k].J1, c->J1[r], sizeof R[k].J1);
                memcpy(R[k].M0, c->J0Minv[r], sizeof R[k].M0); memcpy(R[k].M1, c->J1Minv[r], sizeof R[k].M1);
                R[k].b0 = s->bodies[c->body0].fixed ? -1 : c->body0;
                R[k].b1 = s->bodies[c->body1].fixed ? -1 : c->body1;
                R[k].kind = pass == 0 ? BPP_BILATERAL : (r == 0 ? BPP_NORMAL : BPP_FRICTION);
                R[k].parent = first;
                R[k].rhs = b[r];
                R[k].eps = pass == 0 ? 1e-5f : 1e-3f;
            }
        }
    }
    *out = R;
    return n;
}

static inline float dot6(const float* a, const float* b)
{
    float acc = 0.f;
    for (int k = 0; k < 6; ++k) acc += a[k] * b[k];
    return acc;
}

static float bpp_entry(const bpp_row* R, int i, int j)
{
    const bpp_row* a = &R[i]; const bpp_row* c = &R[j];
    float acc = 0.f;
    if (a->b0 >= 0) {
        if (a->b0 == c->b0) acc += dot6(a->M0, c->J0);
        if (a->b0 == c->b1) acc += dot6(a->M0, c->J1);
    }
    if (a->b1 >= 0) {
        if (a->b1 == c->b0) acc += dot6(a->M1, c->J0);
        if (a->b1 == c->b1) acc += dot6(a->M1, c->J1);
    }
    if (i == j) acc += a->eps;
    return acc;
}

static inline float bpp_hi(const bpp_row* R, const float* lam, int i, float mu)
{
    const float ln = lam[R[i].parent];
    return mu * (ln > 0.f ? ln : 0.f);
}

static void bpp_store(orc_system* s, const float* lam)
{
    int k = 0;
    for (int i = 0; i < s->n_joints; ++i) for (int r = 0; r < s->joints[i].dim; ++r) s->joints[i].lambda[r] = lam[k++];
    for (int i = 0; i < s->n_contacts; ++i) for (int r = 0; r < 3; ++r) s->contacts[i].lambda[r] = lam[k++];
}

static void solve_bpp(orc_system* s, float h)
{
    bpp_row* R;
    const int n = bpp_rows(s, h, &R);
    for (int i = 0; i < s->n_contacts; ++i) s->contacts[i].lambda[0] = s->contacts[i].lambda[1] = s->contacts[i].lambda[2] = 0.f;
    const int maxit = 5 * s->solverIter;
    if (n == 0 || maxit <= 0) { free(R); return; }
    float* lam = (float*)calloc((size_t)n * 4, sizeof(float));
    float *w = lam + n, *y = lam + 2 * n, *hi = lam + 3 * n;
    int* state = (int*)calloc((size_t)n * 3, sizeof(int));
    int *F = state + n, *move = state + 2 * n;
    float* M = (float*)calloc((size_t)n * (n + 1) / 2, sizeof(float));
#define TRI(i, j) ((size_t)(i) * ((i) + 1) / 2 + (j))
    for (int i = 0; i < n; ++i) state[i] = R[i].kind == BPP_BILATERAL ? BPP_FREE : BPP_LOWER;
    int best = n + 1, budget = 3, refresh = 1;
    for (int it = 0; it < maxit; ++it) {
        if (refresh) {
            /* friction bounds from the normal impulses of the last converged pass (0 on the first): the pass
             * below is a box LCP with constant bounds; a friction row of an open contact is pinned at 0 */
            for (int i = 0; i < n; ++i) {
                hi[i] = R[i].kind == BPP_FRICTION ? bpp_hi(R, lam, i, s->mu) : 0.f;
                if (R[i].kind == BPP_FRICTION && hi[i] <= 0.f) state[i] = BPP_LOWER;
            }
            refresh = 0; best = n + 1; budget = 3;
        }
        int nf = 0;
        for (int i = 0; i < n; ++i) {
            if (state[i] == BPP_FREE) F[nf++] = i;
            else lam[i] = R[i].kind == BPP_FRICTION ? (state[i] == BPP_LOWER ? -hi[i] : hi[i]) : 0.f;
        }
        /* free rows: A_FF lam_F = b_F - A_FB lam_B */
        for (int a = 0; a < nf; ++a) {
            for (int c = 0; c <= a; ++c) M[TRI(a, c)] = bpp_entry(R, F[a], F[c]);
            float acc = R[F[a]].rhs;
            for (int j = 0; j < n; ++j) if (state[j] != BPP_FREE && lam[j] != 0.f) acc -= bpp_entry(R, F[a], j) * lam[j];
            y[a] = acc;
        }
        for (int k = 0; k < nf; ++k) {
            float d = M[TRI(k, k)];
            d = sqrtf(d > 1e-12f ? d : 1e-12f);
            M[TRI(k, k)] = d;
            for (int i = k + 1; i < nf; ++i) M[TRI(i, k)] /= d;
            for (int i = k + 1; i < nf; ++i)
                for (int j = k + 1; j <= i; ++j) M[TRI(i, j)] -= M[TRI(i, k)] * M[TRI(j, k)];
        }
        for (int k = 0; k < nf; ++k) {
            y[k] /= M[TRI(k, k)];
            for (int i = k + 1; i < nf; ++i) y[i] -= M[TRI(i, k)] * y[k];
        }
        for (int k = nf - 1; k >= 0; --k) {
            y[k] /= M[TRI(k, k)];
            for (int i = 0; i < k; ++i) y[i] -= M[TRI(k, i)] * y[k];
        }
        for (int a = 0; a < nf; ++a) lam[F[a]] = y[a];
        /* residual velocities of the bound rows and the rows that must change set */
        int ninf = 0, last = -1;
        for (int i = 0; i < n; ++i) {
            move[i] = -1;
            if (R[i].kind == BPP_BILATERAL) continue;
            if (R[i].kind == BPP_FRICTION && hi[i] <= 0.f) continue;
            if (state[i] == BPP_FREE) {
                if (R[i].kind == BPP_NORMAL) { if (lam[i] < -BPP_TOL) move[i] = BPP_LOWER; }
                else if (lam[i] < -hi[i] - BPP_TOL) move[i] = BPP_LOWER;
                else if (lam[i] > hi[i] + BPP_TOL) move[i] = BPP_UPPER;
            } else {
                float acc = -R[i].rhs;
                for (int j = 0; j < n; ++j) if (lam[j] != 0.f) acc += bpp_entry(R, i, j) * lam[j];
                w[i] = acc;
                if (state[i] == BPP_LOWER && acc < -BPP_TOL) move[i] = BPP_FREE;
                if (state[i] == BPP_UPPER && acc > BPP_TOL) move[i] = BPP_FREE;
            }
            if (move[i] >= 0) { ++ninf; last = i; }
        }
        if (ninf == 0) {
            float shift = 0.f;
            for (int i = 0; i < n; ++i) if (R[i].kind == BPP_FRICTION) {
                const float nh = bpp_hi(R, lam, i, s->mu);
                const float ds = fabsf(nh - hi[i]) / (1.0f + nh);
                if (ds > shift) shift = ds;
            }
            if (shift <= 1e-5f) break;         /* sets and friction bounds are both stationary */
            refresh = 1;
            continue;
        }
        int block = 1;
        if (ninf < best) { best = ninf; budget = 3; }
        else if (budget > 0) --budget;
        else block = 0;
        if (block) { for (int i = 0; i < n; ++i) if (move[i] >= 0) state[i] = move[i]; }
        else state[last] = move[last];
    }
    /* admissible impulses whatever happened: project onto the bounds */
    for (int i = 0; i < n; ++i) if (R[i].kind == BPP_NORMAL && lam[i] < 0.f) lam[i] = 0.f;
    for (int i = 0; i < n; ++i) if (R[i].kind == BPP_FRICTION) {
        const float nh = bpp_hi(R, lam, i, s->mu);
        if (lam[i] < -nh) lam[i] = -nh; else if (lam[i] > nh) lam[i] = nh;
    }
#undef TRI
    bpp_store(s, lam);
    free(M); free(state); free(lam); free(R);
}

/* natural-map residual of the box LCP solve_bpp poses, for the impulses currently stored:
 * max_i | lambda_i - clamp(lambda_i - w_i, lo_i, hi_i) |  (0 at a solution) */
float orc_blcp_residual(orc_system* s, float h)
{
    bpp_row* R;
    const int n = bpp_rows(s, h, &R);
    float* lam = (float*)calloc((size_t)(n ? n : 1), sizeof(float));
    int k = 0;
    for (int i = 0; i < s->n_joints; ++i) for (int r = 0; r < s->joints[i].dim; ++r) lam[k++] = s->joints[i].lambda[r];
    for (int i = 0; i < s->n_contacts; ++i) for (int r = 0; r < 3; ++r) lam[k++] = s->contacts[i].lambda[r];
    float worst = 0.f;
    for (int i = 0; i < n; ++i) {
        float acc = -R[i].rhs;
        for (int j = 0; j < n; ++j) acc += bpp_entry(R, i, j) * lam[j];
        float res;
        if (R[i].kind == BPP_BILATERAL) res = fabsf(acc);
        else {
            const float hi = R[i].kind == BPP_NORMAL ? INFINITY : bpp_hi(R, lam, i, s->mu);
            const float lo = R[i].kind == BPP_NORMAL ? 0.f : -hi;
            float z = lam[i] - acc;
            if (z < lo) z = lo; else if (z > hi) z = hi;
            res = fabsf(lam[i] - z);
        }
        if (res > worst) worst = res;
    }
    free(lam); free(R);
    return worst;
}

/* calcConstraintForces head, RigidBodySystem.cpp:181-182 */
void orc_stage_solve(orc_system* s, float dt)
{
    switch (s->solverId) {
    case 1:  solve_conj_gradient(s, dt); break;
    case 2:  solve_conj_residual(s, dt); break;
    case 3:  solve_bpp(s, dt); break;                 /* extension, no reference counterpart */
    default: solve_box_pgs(s, dt); break;
    }
}

/* ------------------------------------------------------------------ */
/* S19: constraint-force scatter, RigidBodySystem.cpp:185-220          */
/* ------------------------------------------------------------------ */
static void jt_lambda_over_dt(const con_t* c, float J[5][6], float dt, float* f6)
{
    for (int k = 0; k < 6; ++k) {
        float acc = 0.f;
        for (int r = 0; r < c->dim; ++r) acc += J[r][k] * c->lambda[r];
        f6[k] = acc / dt;
    }
}

void orc_stage_scatter(orc_system* s, float dt)
{
    for (int i = 0; i < s->n_joints; ++i) {
        con_t* j = &s->joints[i];
        float f0[6], f1[6];
        jt_lambda_over_dt(j, j->J0, dt, f0);
        jt_lambda_over_dt(j, j->J1, dt, f1);
        body_t* b0 = &s->bodies[j->body0];
        body_t* b1 = &s->bodies[j->body1];
        b0->fc = v3_add(b0->fc, V3(f0[0], f0[1], f0[2])); b0->tauc = v3_add(b0->tauc, V3(f0[3], f0[4], f0[5]));
        b1->fc = v3_add(b1->fc, V3(f1[0], f1[1], f1[2])); b1->tauc = v3_add(b1->tauc, V3(f1[3], f1[4], f1[5]));
    }
    for (int i = 0; i < s->n_contacts; ++i) {
        con_t* c = &s->contacts[i];
        float f0[6], f1[6];
        jt_lambda_over_dt(c, c->J0, dt, f0);
        jt_lambda_over_dt(c, c->J1, dt, f1);
        body_t* b0 = &s->bodies[c->body0];
        body_t* b1 = &s->bodies[c->body1];
        if (!b0->fixed) { b0->fc = v3_add(b0->fc, V3(f0[0], f0[1], f0[2])); b0->tauc = v3_add(b0->tauc, V3(f0[3], f0[4], f0[5])); }
        if (!b1->fixed) { b1->fc = v3_add(b1->fc, V3(f1[0], f1[1], f1[2])); b1->tauc = v3_add(b1->tauc, V3(f1[3], f1[4], f1[5])); }
    }
}

/* ------------------------------------------------------------------ */
/* S20: integration, RigidBodySystem.cpp:99-120                        */
/* ------------------------------------------------------------------ */
void orc_stage_integrate(orc_system* s, float dt)
{
    for (int i = 0; i < s->n_bodies; ++i) {
        body_t* b = &s->bodies[i];
        if (!b->fixed) {
            b->xdot = v3_add(b->xdot, v3_scale(dt * (1.0f / b->mass), v3_add(b->f, b->fc)));
            const v3 Iw = m3_mulv(&b->I, b->omega);
            const v3 rhs = v3_sub(v3_add(b->tau, b->tauc), v3_cross(b->omega, Iw));
            /* (dt * Iinv) * v evaluated as dt * (Iinv * v)   */
            b->omega = v3_add(b->omega, v3_scale(dt, m3_mulv(&b->Iinv, rhs)));
            b->x = v3_add(b->x, v3_scale(dt, b->xdot));
            /* q = q + (0.5*dt) * Quaternionf(0, omega) * q ; helpers RigidBodySystem.h:78-88 */
            const float hdt = 0.5f * dt;
            quat wq; wq.w = hdt * 0.0f; wq.x = hdt * b->omega.x; wq.y = hdt * b->omega.y; wq.z = hdt * b->omega.z;
            const quat dq = q_mul(wq, b->q);
            quat qn; qn.w = b->q.w + dq.w; qn.x = b->q.x + dq.x; qn.y = b->q.y + dq.y; qn.z = b->q.z + dq.z;
            b->q = q_normalized(qn);
        } else {
            b->xdot = V3(0, 0, 0);
            b->omega = V3(0, 0, 0);
        }
    }
}

/* RigidBodySystem::step, RigidBodySystem.cpp:52-121 (pre-step callback :70-73 is the caller's
 * business: run orc_stage_prepare, mutate, then the remaining stages) */
void orc_step(orc_system* s, float dt)
{
    orc_stage_prepare(s);
    orc_stage_detect(s);
    orc_stage_jacobians(s);
    orc_stage_solve(s, dt);
    orc_stage_scatter(s, dt);
    orc_stage_integrate(s, dt);
}

double orc_time_steps(orc_system* s, float dt, int n)
{
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < n; ++i) orc_step(s, dt);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------ */
/* accessors                                                           */
/* ------------------------------------------------------------------ */
int orc_num_bodies(const orc_system* s) { return s->n_bodies; }
int orc_num_joints(const orc_system* s) { return s->n_joints; }
int orc_num_contacts(const orc_system* s) { return s->n_contacts; }
int orc_num_joint_rows(const orc_system* s) { int n = 0; for (int i = 0; i < s->n_joints; ++i) n += s->joints[i].dim; return n; }
long long orc_pairs_tested(const orc_system* s) { return s->pairs_tested; }

static void put3(float* dst, int i, v3 v) { if (dst) { dst[3 * i] = v.x; dst[3 * i + 1] = v.y; dst[3 * i + 2] = v.z; } }
static void putq(float* dst, int i, quat q) { if (dst) { dst[4 * i] = q.x; dst[4 * i + 1] = q.y; dst[4 * i + 2] = q.z; dst[4 * i + 3] = q.w; } }
static void putm(float* dst, int i, const m3* m) { if (dst) memcpy(dst + 9 * i, m->m, 9 * sizeof(float)); }

void orc_get_state(const orc_system* s, float* x3, float* q, float* v3_, float* w3)
{
    for (int i = 0; i < s->n_bodies; ++i) {
        const body_t* b = &s->bodies[i];
        put3(x3, i, b->x); putq(q, i, b->q); put3(v3_, i, b->xdot); put3(w3, i, b->omega);
    }
}

void orc_set_state(orc_system* s, const float* x3, const float* q, const float* v, const float* w)
{
    for (int i = 0; i < s->n_bodies; ++i) {
        body_t* b = &s->bodies[i];
        if (x3) b->x = V3(x3[3 * i], x3[3 * i + 1], x3[3 * i + 2]);
        if (q) { b->q.x = q[4 * i]; b->q.y = q[4 * i + 1]; b->q.z = q[4 * i + 2]; b->q.w = q[4 * i + 3]; }
        if (v) b->xdot = V3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        if (w) b->omega = V3(w[3 * i], w[3 * i + 1], w[3 * i + 2]);
    }
}

void orc_get_body_params(const orc_system* s, float* mass, float* ibody9, float* ibodyinv9,
                         int* fixed, int* geom_type, float* g)
{
    for (int i = 0; i < s->n_bodies; ++i) {
        const body_t* b = &s->bodies[i];
        if (mass) mass[i] = b->mass;
        putm(ibody9, i, &b->Ibody); putm(ibodyinv9, i, &b->IbodyInv);
        if (fixed) fixed[i] = b->fixed;
        if (geom_type) geom_type[i] = b->geom_type;
        if (g) {
            float* o = g + 6 * i;
            for (int k = 0; k < 6; ++k) o[k] = 0.f;
            switch (b->geom_type) {
            case ORC_SPHERE:   o[0] = b->radius; break;
            case ORC_BOX:      o[0] = b->dim.x; o[1] = b->dim.y; o[2] = b->dim.z; break;
            case ORC_PLANE:    o[0] = b->plane_p.x; o[1] = b->plane_p.y; o[2] = b->plane_p.z;
                               o[3] = b->plane_n.x; o[4] = b->plane_n.y; o[5] = b->plane_n.z; break;
            case ORC_CYLINDER: o[0] = b->height; o[1] = b->radius; break;
            case ORC_SDF:      o[0] = (float)b->shape; break;
            }
        }
    }
}

void orc_get_body_dyn(const orc_system* s, float* f3, float* tau3, float* fc3, float* tauc3, float* I9, float* Iinv9)
{
    for (int i = 0; i < s->n_bodies; ++i) {
        const body_t* b = &s->bodies[i];
        put3(f3, i, b->f); put3(tau3, i, b->tau); put3(fc3, i, b->fc); put3(tauc3, i, b->tauc);
        putm(I9, i, &b->I); putm(Iinv9, i, &b->Iinv);
    }
}

void orc_get_contacts(const orc_system* s, int* body0, int* body1, float* p3, float* n3,
                      float* t3, float* b3, float* phi, float* lambda3)
{
    for (int i = 0; i < s->n_contacts; ++i) {
        const con_t* c = &s->contacts[i];
        if (body0) body0[i] = c->body0;
        if (body1) body1[i] = c->body1;
        put3(p3, i, c->p); put3(n3, i, c->n); put3(t3, i, c->t); put3(b3, i, c->b);
        if (phi) phi[i] = c->pene;
        if (lambda3) { lambda3[3 * i] = c->lambda[0]; lambda3[3 * i + 1] = c->lambda[1]; lambda3[3 * i + 2] = c->lambda[2]; }
    }
}

void orc_get_contact_rows(const orc_system* s, float* J0, float* J1, float* J0M, float* J1M)
{
    for (int i = 0; i < s->n_contacts; ++i) {
        const con_t* c = &s->contacts[i];
        for (int r = 0; r < 3; ++r) for (int k = 0; k < 6; ++k) {
            const int o = 18 * i + 6 * r + k;
            if (J0) J0[o] = c->J0[r][k];
            if (J1) J1[o] = c->J1[r][k];
            if (J0M) J0M[o] = c->J0Minv[r][k];
            if (J1M) J1M[o] = c->J1Minv[r][k];
        }
    }
}

void orc_get_joints(const orc_system* s, int* type, int* body0, int* body1, float* r0, float* q0,
                    float* r1, float* q1)
{
    for (int i = 0; i < s->n_joints; ++i) {
        const con_t* j = &s->joints[i];
        if (type) type[i] = j->type;
        if (body0) body0[i] = j->body0;
        if (body1) body1[i] = j->body1;
        put3(r0, i, j->r0); put3(r1, i, j->r1); putq(q0, i, j->q0); putq(q1, i, j->q1);
    }
}

void orc_get_joint_rows(const orc_system* s, float* J0, float* J1, float* J0M, float* J1M, float* phi5, float* lambda5)
{
    for (int i = 0; i < s->n_joints; ++i) {
        const con_t* j = &s->joints[i];
        for (int r = 0; r < 5; ++r) {
            for (int k = 0; k < 6; ++k) {
                const int o = 30 * i + 6 * r + k;
                const int live = r < j->dim;
                if (J0) J0[o] = live ? j->J0[r][k] : 0.f;
                if (J1) J1[o] = live ? j->J1[r][k] : 0.f;
                if (J0M) J0M[o] = live ? j->J0Minv[r][k] : 0.f;
                if (J1M) J1M[o] = live ? j->J1Minv[r][k] : 0.f;
            }
            if (phi5) phi5[5 * i + r] = r < j->dim ? j->phi[r] : 0.f;
            if (lambda5) lambda5[5 * i + r] = r < j->dim ? j->lambda[r] : 0.f;
        }
    }
}

void orc_set_joint_lambda(orc_system* s, const float* lambda5)
{
    for (int i = 0; i < s->n_joints; ++i)
        for (int r = 0; r < s->joints[i].dim; ++r) s->joints[i].lambda[r] = lambda5[5 * i + r];
}

/* RigidBodySystemState::save / restore, RigidBodyState.cpp:18-72 (joint lambda is NOT reset) */
void orc_save_state(orc_system* s)
{
    for (int i = 0; i < s->n_bodies; ++i) {
        body_t* b = &s->bodies[i];
        b->sx = b->x; b->sxdot = b->xdot; b->sq = b->q; b->somega = b->omega;
    }
}

void orc_restore_state(orc_system* s)
{
    for (int i = 0; i < s->n_bodies; ++i) {
        body_t* b = &s->bodies[i];
        b->x = b->sx; b->xdot = b->sxdot; b->q = b->sq; b->omega = b->somega;
        b->f = V3(0, 0, 0); b->fc = V3(0, 0, 0);
        b->n_contacts = 0;
    }
    s->n_contacts = 0;
}

/* ------------------------------------------------------------------ */
/* Scenarios.h scene data (include/rigidbody/Scenarios.h)              */
/* ------------------------------------------------------------------ */
static int add_simple(orc_system* s, float mass, int type, float g0, float g1, float g2, float g3, float g4, float g5)
{
    const float g[6] = { g0, g1, g2, g3, g4, g5 };
    return orc_add_body(s, mass, type, g, 0, NULL, NULL, NULL, NULL);
}

/* createMarbleBox, Scenarios.h:18-80 */
void orc_scene_marble_box(orc_system* s)
{
    orc_clear(s);
    const float radius = 0.5f;
    for (int i = 0; i < 9; ++i) {
        for (int j = 0; j < 9; ++j) {
            int b1 = add_simple(s, 1.0f, ORC_SPHERE, radius, 0, 0, 0, 0, 0);
            s->bodies[b1].x.x = -4.0f + (float)i * 1.0f;
            s->bodies[b1].x.z = -4.0f + (float)j * 1.0f;
            s->bodies[b1].x.y = 2.0f;
            int b2 = add_simple(s, 1.0f, ORC_SPHERE, radius, 0, 0, 0, 0, 0);
            s->bodies[b2].x.x = -4.0f + (float)i * 1.0f;
            s->bodies[b2].x.z = -4.0f + (float)j * 1.0f;
            s->bodies[b2].x.y = 3.0f;
        }
    }
    const int s0 = add_simple(s, 1.0f, ORC_BOX, 0.4f, 4.0f, 10.0f, 0, 0, 0);
    const int s1 = add_simple(s, 1.0f, ORC_BOX, 0.4f, 4.0f, 10.0f, 0, 0, 0);
    const int s2 = add_simple(s, 1.0f, ORC_BOX, 0.4f, 4.0f, 10.0f, 0, 0, 0);
    const int s3 = add_simple(s, 1.0f, ORC_BOX, 0.4f, 4.0f, 10.0f, 0, 0, 0);
    const int s4 = add_simple(s, 1.0f, ORC_BOX, 10.0f, 0.4f, 10.0f, 0, 0, 0);
    body_t* B = s->bodies;
    B[s0].fixed = B[s1].fixed = B[s2].fixed = B[s3].fixed = B[s4].fixed = 1;
    B[s0].x = V3(4.75f, 2.0f, 0.0f);
    B[s1].x = V3(-4.75f, 2.0f, 0.0f);
    B[s2].x = V3(0.0f, 2.0f, 4.75f);
    B[s2].q = q_from_angle_axis(1.57f, V3(0, 1, 0));
    B[s3].x = V3(0.0f, 2.0f, -4.75f);
    B[s3].q = q_from_angle_axis(1.57f, V3(0, 1, 0));
    B[s4].x = V3(0.0f, 0.0f, 0.0f);
}

/* createSphereOnBox, Scenarios.h:84-108 */
void orc_scene_sphere_on_box(orc_system* s)
{
    orc_clear(s);
    const int sp = add_simple(s, 1.0f, ORC_SPHERE, 0.5f, 0, 0, 0, 0, 0);
    s->bodies[sp].x.y = 4.0f;
    s->bodies[sp].omega = V3(10.0f, 0.0f, 0.0f);
    const int bx = add_simple(s, 1.0f, ORC_BOX, 10.0f, 0.4f, 10.0f, 0, 0, 0);
    s->bodies[bx].fixed = 1;
}

/* createSwingingBoxes, Scenarios.h:112-148 */
void orc_scene_swinging_boxes(orc_system* s)
{
    orc_clear(s);
    const int N = 20;
    const int top = add_simple(s, 1.0f, ORC_BOX, 1.0f, 1.0f, 1.0f, 0, 0, 0);
    s->bodies[top].x = V3(0.0f, 1.5f * (float)N, 0.0f);
    s->bodies[top].fixed = 1;
    const v3 dx = V3(0.0f, 1.5f, 0.0f);
    int parent = top;
    const float qi[4] = { 0.f, 0.f, 0.f, 1.f };
    for (int i = 0; i < N - 1; ++i) {
        const float mass = (i == (N - 2)) ? 100.0f : 1.0f;
        const int next = add_simple(s, mass, ORC_BOX, 1.0f, 1.0f, 1.0f, 0, 0, 0);
        s->bodies[next].x = v3_sub(s->bodies[parent].x, dx);
        const v3 a0 = v3_scale(-0.5f, dx), a1 = v3_scale(0.5f, dx);
        const float r0[3] = { a0.x, a0.y, a0.z }, r1[3] = { a1.x, a1.y, a1.z };
        orc_add_hinge(s, parent, next, r0, qi, r1, qi);
        parent = next;
    }
    s->bodies[parent].xdot = V3(0.0f, 0.0f, 10.0f);
}

/* createCylinderOnPlane, Scenarios.h:153-174 */
void orc_scene_cylinder_on_plane(orc_system* s)
{
    orc_clear(s);
    const int cyl = add_simple(s, 1.0f, ORC_CYLINDER, 2.0f, 1.0f, 0, 0, 0, 0);
    s->bodies[cyl].x = V3(0.0f, 2.0f, 0.0f);
    s->bodies[cyl].q = q_from_angle_axis(0.57f, V3(0.0f, 0.0f, 1.0f));
    const int pl = add_simple(s, 1.0f, ORC_PLANE, 0, 0, 0, 0, 1.0f, 0);
    s->bodies[pl].x = V3(0, 0, 0);
    s->bodies[pl].fixed = 1;
}

/* createCarScene, Scenarios.h:178-247 */
void orc_scene_car(orc_system* s)
{
    orc_clear(s);
    const int chassis = add_simple(s, 5.0f, ORC_BOX, 2.0f, 0.5f, 3.0f, 0, 0, 0);
    int wheel[4];
    for (int k = 0; k < 4; ++k) wheel[k] = add_simple(s, 1.0f, ORC_CYLINDER, 0.2f, 0.5f, 0, 0, 0, 0);
    body_t* B = s->bodies;
    B[chassis].x = V3(0.0f, 0.5f, 0.0f);
    B[chassis].xdot = V3(0.0f, 0.0f, -10.0f);
    const float wx[4] = { 1.0f, -1.0f, 1.0f, -1.0f };
    const float wz[4] = { 1.5f, 1.5f, -1.5f, -1.5f };
    for (int k = 0; k < 4; ++k) {
        B[wheel[k]].x = V3(wx[k], 0.5f, wz[k]);
        B[wheel[k]].q = q_from_angle_axis(1.57079f, V3(0.0f, 0.0f, 1.0f));
    }
    const int plane = add_simple(s, 1.0f, ORC_PLANE, 0, 0, 0, 0, 1.0f, 0);
    s->bodies[plane].x = V3(0, 0, 0);
    s->bodies[plane].fixed = 1;
    const quat hq = q_from_angle_axis(1.57f, V3(0, 0, 1));
    const float q0[4] = { 0.f, 0.f, 0.f, 1.f };
    const float q1[4] = { hq.x, hq.y, hq.z, hq.w };
    const float r1[3] = { 0.f, 0.f, 0.f };
    for (int k = 0; k < 4; ++k) {
        const float r0[3] = { wx[k], 0.0f, wz[k] };
        orc_add_hinge(s, chassis, wheel[k], r0, q0, r1, q1);
    }
}

/* K spheres (r = 0.5, m = 1) stacked at y = 0.7 + k on the fixed
 * 10 x 0.4 x 10 box of createSphereOnBox (Scenarios.h:99-101). Spheres first, then the box. */
void orc_scene_sphere_stack(orc_system* s, int k)
{
    orc_clear(s);
    for (int i = 0; i < k; ++i) {
        const int sp = add_simple(s, 1.0f, ORC_SPHERE, 0.5f, 0, 0, 0, 0, 0);
        s->bodies[sp].x = V3(0.0f, 0.7f + (float)i * 1.0f, 0.0f);
    }
    const int bx = add_simple(s, 1.0f, ORC_BOX, 10.0f, 0.4f, 10.0f, 0, 0, 0);
    s->bodies[bx].fixed = 1;
}

/* SURVEY.md 8(d) MB1k family: createMarbleBox's construction scaled to nx x nz x ny spheres
 * (r = 0.5, spacing 1, lowest layer y = 2), container = 4 fixed side boxes + floor sized to fit.
 * Body order: for i, for j, for k (layers interleaved like the reference's lower/upper pair). */
void orc_scene_marble_box_n(orc_system* s, int nx, int nz, int ny)
{
    orc_clear(s);
    const float radius = 0.5f;
    const float x0 = -0.5f * (float)(nx - 1), z0 = -0.5f * (float)(nz - 1);
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < nz; ++j)
            for (int k = 0; k < ny; ++k) {
                const int b = add_simple(s, 1.0f, ORC_SPHERE, radius, 0, 0, 0, 0, 0);
                s->bodies[b].x = V3(x0 + (float)i * 1.0f, 2.0f + (float)k * 1.0f, z0 + (float)j * 1.0f);
            }
    const float ex = (float)nx + 1.0f, ez = (float)nz + 1.0f;     /* inner clearance like 10 for 9 spheres */
    const float hy = (float)ny + 2.0f;
    const float wallx = 0.5f * ex - 0.25f, wallz = 0.5f * ez - 0.25f;
    const int s0 = add_simple(s, 1.0f, ORC_BOX, 0.4f, hy, ez, 0, 0, 0);
    const int s1 = add_simple(s, 1.0f, ORC_BOX, 0.4f, hy, ez, 0, 0, 0);
    const int s2 = add_simple(s, 1.0f, ORC_BOX, 0.4f, hy, ex, 0, 0, 0);
    const int s3 = add_simple(s, 1.0f, ORC_BOX, 0.4f, hy, ex, 0, 0, 0);
    const int s4 = add_simple(s, 1.0f, ORC_BOX, ex, 0.4f, ez, 0, 0, 0);
    body_t* B = s->bodies;
    B[s0].fixed = B[s1].fixed = B[s2].fixed = B[s3].fixed = B[s4].fixed = 1;
    B[s0].x = V3(wallx, 0.5f * hy, 0.0f);
    B[s1].x = V3(-wallx, 0.5f * hy, 0.0f);
    B[s2].x = V3(0.0f, 0.5f * hy, wallz);
    B[s2].q = q_from_angle_axis(1.57f, V3(0, 1, 0));
    B[s3].x = V3(0.0f, 0.5f * hy, -wallz);
    B[s3].q = q_from_angle_axis(1.57f, V3(0, 1, 0));
    B[s4].x = V3(0.0f, 0.0f, 0.0f);
}

/* EXTENSION scene (BASELINE.json config 0 "box-stack on plane"; no such scene in Scenarios.h): k unit boxes
 * (mass 1) stacked on a fixed Plane at y = 0, centres at y = 0.5 + i * 1.0, alternately turned by 0.3 rad about y
 * so that the face-face clipping is exercised. Needs orc_set_extensions(s, 1). Bodies: boxes bottom-up, then the plane. */
void orc_scene_box_stack(orc_system* s, int k)
{
    orc_clear(s);
    for (int i = 0; i < k; ++i) {
        const int b = add_simple(s, 1.0f, ORC_BOX, 1.0f, 1.0f, 1.0f, 0, 0, 0);
        s->bodies[b].x = V3(0.0f, 0.5f + (float)i * 1.0f, 0.0f);
        if (i & 1) s->bodies[b].q = q_from_angle_axis(0.3f, V3(0, 1, 0));
    }
    const int pl = add_simple(s, 1.0f, ORC_PLANE, 0, 0, 0, 0, 1.0f, 0);
    s->bodies[pl].fixed = 1;
    s->ext_pairs = 1;
}

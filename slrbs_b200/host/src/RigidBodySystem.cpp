// RigidBodySystem.cpp -- host mirror of src/rigidbody/RigidBodySystem.cpp. step() keeps the
// reference's order (:52-121) but every stage runs on the device through include/slrbs_b200.h:
//   force reset + inertias -> [pre-step callback on the host] -> detect -> Jacobians -> solve
//   (Solver seam) -> constraint forces -> integrate -> refresh of the public host fields.
#include "rigidbody/RigidBodySystem.h"

#include "collision/CollisionDetect.h"
#include "contact/Contact.h"
#include "joint/Joint.h"
#include "rigidbody/RigidBody.h"
#include "solvers/SolverBoxPGS.h"
#include "solvers/SolverConjGradient.h"
#include "solvers/SolverConjResidual.h"
#include "solvers/SolverBPP.h"
#include "collision/SDFMesh.h"

#include "slrbs_b200.h"

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>

// body pointer -> index without the O(bodies) scan per joint (packParams runs every step)
int RigidBodySystem::bodyIndexOf(const RigidBody* b) const
{
    if (m_bodyIndexDirty) {
        m_bodyIndex.clear();
        for (size_t i = 0; i < m_bodies.size(); ++i) m_bodyIndex[m_bodies[i]] = (int)i;
        m_bodyIndexDirty = false;
    }
    const auto it = m_bodyIndex.find(b);
    return it == m_bodyIndex.end() ? -1 : it->second;
}

void RigidBodySystem::setSolver(int id, Solver* solver)
{
    if (id < 0 || id > 3) { delete solver; throw std::runtime_error("RigidBodySystem::setSolver: id must be 0..3"); }
    m_solvers[id].reset(solver);
}

RigidBodySystem::RigidBodySystem() : solverIter(10), solverId(0), m_collisionsEnabled(true), m_preStepFunc(nullptr), m_resetFunc(nullptr)
{
    m_collisionDetect = std::make_unique<CollisionDetect>(this);
    m_solvers[0] = std::make_unique<SolverBoxPGS>(this);
    m_solvers[1] = std::make_unique<SolverConjGradient>(this);
    m_solvers[2] = std::make_unique<SolverConjResidual>(this);
    m_solvers[3] = std::make_unique<SolverBPP>(this);          // extension
}

RigidBodySystem::~RigidBodySystem()
{
    clear();
    if (m_ctx) slrbs_destroy(m_ctx);
}

void RigidBodySystem::check(int rc, const char* what) const
{
    if (rc < 0) throw std::runtime_error(std::string(what) + ": " + slrbs_last_error());
}

void RigidBodySystem::addBody(RigidBody* _b) { m_bodies.push_back(_b); m_topologyDirty = true; m_bodyIndexDirty = true; }

void RigidBodySystem::addJoint(Joint* _j)
{
    m_joints.push_back(_j);
    if (_j->body0) _j->body0->joints.push_back(_j);
    if (_j->body1) _j->body1->joints.push_back(_j);
    m_topologyDirty = true;
}

void RigidBodySystem::clear()
{
    if (m_resetFunc) m_resetFunc();
    m_collisionDetect->clear();
    for (Joint* j : m_joints) delete j;
    m_joints.clear();
    for (RigidBody* b : m_bodies) delete b;
    m_bodies.clear();
    m_topologyDirty = true;
    m_bodyIndexDirty = true;
}

const std::vector<Contact*>& RigidBodySystem::getContacts() const { return m_collisionDetect->getContacts(); }
std::vector<Contact*>& RigidBodySystem::getContacts() { return m_collisionDetect->getContacts(); }

void RigidBodySystem::computeInertias() { for (RigidBody* b : m_bodies) b->updateInertiaMatrix(); }

// ---- packing of the public host fields into the C ABI's dense arrays ------------------------------
void RigidBodySystem::packParams(std::vector<float>& out) const
{
    out.clear();
    for (const RigidBody* b : m_bodies) {
        float g[6] = { 0, 0, 0, 0, 0, 0 };
        b->geometry->packParams(g);
        out.push_back(b->mass); out.push_back(b->fixed ? 1.f : 0.f); out.push_back((float)b->geometry->getType());
        out.insert(out.end(), g, g + 6);
        out.insert(out.end(), b->Ibody.data(), b->Ibody.data() + 9);
        out.insert(out.end(), b->IbodyInv.data(), b->IbodyInv.data() + 9);
    }
    for (const Joint* j : m_joints) {
        out.push_back((float)j->getType()); out.push_back((float)bodyIndexOf(j->body0)); out.push_back((float)bodyIndexOf(j->body1));
        for (int k = 0; k < 3; ++k) out.push_back(j->r0[k]);
        for (int k = 0; k < 3; ++k) out.push_back(j->r1[k]);
        out.push_back(j->q0.x()); out.push_back(j->q0.y()); out.push_back(j->q0.z()); out.push_back(j->q0.w());
        out.push_back(j->q1.x()); out.push_back(j->q1.y()); out.push_back(j->q1.z()); out.push_back(j->q1.w());
    }
}

void RigidBodySystem::packState(std::vector<float>& out) const
{
    const size_t n = m_bodies.size();
    out.resize(13 * n);
    float *x = out.data(), *q = x + 3 * n, *v = q + 4 * n, *w = v + 3 * n;
    for (size_t i = 0; i < n; ++i) {
        const RigidBody* b = m_bodies[i];
        for (int k = 0; k < 3; ++k) { x[3 * i + k] = b->x[k]; v[3 * i + k] = b->xdot[k]; w[3 * i + k] = b->omega[k]; }
        q[4 * i] = b->q.x(); q[4 * i + 1] = b->q.y(); q[4 * i + 2] = b->q.z(); q[4 * i + 3] = b->q.w();
    }
}

void RigidBodySystem::packLambda(std::vector<float>& out) const
{
    out.assign(5 * m_joints.size(), 0.f);
    for (size_t i = 0; i < m_joints.size(); ++i)
        for (unsigned k = 0; k < m_joints[i]->dim && k < 5; ++k) out[5 * i + k] = m_joints[i]->lambda(k);
}

static bool sameBits(const std::vector<float>& a, const std::vector<float>& b)
{
    return a.size() == b.size() && (a.empty() || std::memcmp(a.data(), b.data(), a.size() * sizeof(float)) == 0);
}

// extension: distinct SDF shapes of the bodies, in order of first appearance = their slots in the device shape table
static std::vector<const SDFShape*> collectShapes(const std::vector<RigidBody*>& bodies)
{
    std::vector<const SDFShape*> shapes;
    for (const RigidBody* b : bodies) {
        if (b->geometry->getType() != kSDFMesh) continue;
        const SDFShape* sh = static_cast<const SDFMesh*>(b->geometry.get())->shape.get();
        auto it = std::find(shapes.begin(), shapes.end(), sh);
        sh->shapeId = (int)(it - shapes.begin());
        if (it == shapes.end()) shapes.push_back(sh);
    }
    return shapes;
}

void RigidBodySystem::uploadAll()
{
    const int n = (int)m_bodies.size(), nj = (int)m_joints.size();
    for (const SDFShape* sh : collectShapes(m_bodies)) {
        const int id = slrbs_sdf_add_shape(m_ctx, sh->nx, sh->ny, sh->nz, sh->origin, sh->cell, sh->grid.data(), (int)sh->verts.size() / 3, sh->verts.data());
        check(id, "slrbs_sdf_add_shape");
        if (id != sh->shapeId) throw std::runtime_error("RigidBodySystem: SDF shape table out of step");
    }
    std::vector<float> st; packState(st);
    std::vector<float> mass(n), ib(9 * n), ibi(9 * n), geom(6 * n, 0.f);
    std::vector<int32_t> fixed(n), type(n);
    for (int i = 0; i < n; ++i) {
        const RigidBody* b = m_bodies[i];
        mass[i] = b->mass; fixed[i] = b->fixed ? 1 : 0; type[i] = (int)b->geometry->getType();
        b->geometry->packParams(&geom[6 * i]);
        std::memcpy(&ib[9 * i], b->Ibody.data(), 9 * sizeof(float));
        std::memcpy(&ibi[9 * i], b->IbodyInv.data(), 9 * sizeof(float));
    }
    const float *x = st.data(), *q = x + 3 * n, *v = q + 4 * n, *w = v + 3 * n;
    check(slrbs_upload_bodies(m_ctx, 0, 1, n, nullptr, x, q, v, w, mass.data(), ib.data(), ibi.data(), fixed.data(), type.data(), geom.data()),
          "slrbs_upload_bodies");
    if (nj > 0) {
        std::vector<int32_t> jt(nj), b0(nj), b1(nj);
        std::vector<float> r0(3 * nj), r1(3 * nj), q0(4 * nj), q1(4 * nj);
        for (int i = 0; i < nj; ++i) {
            const Joint* j = m_joints[i];
            jt[i] = (int)j->getType(); b0[i] = bodyIndexOf(j->body0); b1[i] = bodyIndexOf(j->body1);
            for (int k = 0; k < 3; ++k) { r0[3 * i + k] = j->r0[k]; r1[3 * i + k] = j->r1[k]; }
            q0[4 * i] = j->q0.x(); q0[4 * i + 1] = j->q0.y(); q0[4 * i + 2] = j->q0.z(); q0[4 * i + 3] = j->q0.w();
            q1[4 * i] = j->q1.x(); q1[4 * i + 1] = j->q1.y(); q1[4 * i + 2] = j->q1.z(); q1[4 * i + 3] = j->q1.w();
        }
        check(slrbs_upload_joints(m_ctx, 0, 1, nj, nullptr, jt.data(), b0.data(), b1.data(), r0.data(), q0.data(), r1.data(), q1.data()),
              "slrbs_upload_joints");
        std::vector<float> lam; packLambda(lam);
        check(slrbs_upload_joint_lambda(m_ctx, 0, 1, nj, lam.data()), "slrbs_upload_joint_lambda");
        m_lambdaSnapshot = lam;
    }
    m_stateSnapshot = st;
    packParams(m_paramSnapshot);
}

void RigidBodySystem::ensureContext()
{
    const int n = (int)m_bodies.size(), nj = (int)m_joints.size();
    if (n == 0) return;
    collectShapes(m_bodies);
    std::vector<float> params; packParams(params);
    if (!m_topologyDirty && m_ctx && sameBits(params, m_paramSnapshot)) return;
    if (m_contactCapacity == 0) m_contactCapacity = std::max(64, 8 * n);
    if (m_ctx) { slrbs_destroy(m_ctx); m_ctx = nullptr; }
    check(slrbs_create(&m_ctx, m_device, 1, n, nj, m_contactCapacity), "slrbs_create");
    uploadAll();
    m_extActive = false;
    m_topologyDirty = false;
}

void RigidBodySystem::uploadStateIfChanged()
{
    const int n = (int)m_bodies.size(), nj = (int)m_joints.size();
    std::vector<float> st; packState(st);
    if (!sameBits(st, m_stateSnapshot)) {
        const float *x = st.data(), *q = x + 3 * n, *v = q + 4 * n, *w = v + 3 * n;
        check(slrbs_upload_state(m_ctx, 0, 1, n, x, q, v, w), "slrbs_upload_state");
        m_stateSnapshot = st;
    }
    if (nj > 0) {
        std::vector<float> lam; packLambda(lam);
        if (!sameBits(lam, m_lambdaSnapshot)) {
            check(slrbs_upload_joint_lambda(m_ctx, 0, 1, nj, lam.data()), "slrbs_upload_joint_lambda");
            m_lambdaSnapshot = lam;
        }
    }
}

// RigidBodySystem.cpp:58-65 on the host fields (what a pre-step callback expects to see)
void RigidBodySystem::hostForceReset()
{
    for (RigidBody* b : m_bodies) {
        b->f = b->mass * Eigen::Vector3f(0.f, -9.81f, 0.f);
        b->tau.setZero(); b->fc.setZero(); b->tauc.setZero();
        b->contacts.clear();
    }
}

void RigidBodySystem::deviceDetect()
{
    check(slrbs_set_params(m_ctx, Contact::mu, solverId, solverIter, m_collisionsEnabled ? 1 : 0), "slrbs_set_params");
    check(slrbs_stage_detect(m_ctx), "slrbs_stage_detect");
}
void RigidBodySystem::deviceContactJacobians(float h) { check(slrbs_stage_jacobians(m_ctx, h), "slrbs_stage_jacobians"); }
void RigidBodySystem::deviceSolve(int id, int iters, float h)
{
    check(slrbs_set_params(m_ctx, Contact::mu, id, iters, m_collisionsEnabled ? 1 : 0), "slrbs_set_params");
    check(slrbs_stage_solve(m_ctx, h), "slrbs_stage_solve");
}

void RigidBodySystem::mirrorContacts()
{
    m_collisionDetect->clear();
    int32_t count = 0;
    check(slrbs_download_contact_counts(m_ctx, 0, 1, &count), "slrbs_download_contact_counts");
    if (count <= 0) return;
    std::vector<int32_t> b0(count), b1(count);
    std::vector<float> p(3 * count), n(3 * count), t(3 * count), b(3 * count), phi(count), lam(3 * count);
    check(slrbs_download_contacts(m_ctx, 0, count, b0.data(), b1.data(), p.data(), n.data(), t.data(), b.data(), phi.data(), lam.data()),
          "slrbs_download_contacts");
    std::vector<float> J0, J1, M0, M1;
    if (m_rowReadback) {
        J0.resize(18 * count); J1.resize(18 * count); M0.resize(18 * count); M1.resize(18 * count);
        check(slrbs_download_contact_rows(m_ctx, 0, count, J0.data(), J1.data(), M0.data(), M1.data(), nullptr, nullptr),
              "slrbs_download_contact_rows");
    }
    auto& out = m_collisionDetect->getContacts();
    out.reserve(count);
    for (int i = 0; i < count; ++i) {
        Contact* c = new Contact(m_bodies[b0[i]], m_bodies[b1[i]], Eigen::Vector3f(p[3 * i], p[3 * i + 1], p[3 * i + 2]),
                                 Eigen::Vector3f(n[3 * i], n[3 * i + 1], n[3 * i + 2]), phi[i]);
        c->t = Eigen::Vector3f(t[3 * i], t[3 * i + 1], t[3 * i + 2]);
        c->b = Eigen::Vector3f(b[3 * i], b[3 * i + 1], b[3 * i + 2]);
        for (int k = 0; k < 3; ++k) c->lambda(k) = lam[3 * i + k];
        if (m_rowReadback)
            for (int r = 0; r < 3; ++r) for (int k = 0; k < 6; ++k) {
                c->J0(r, k) = J0[18 * i + 6 * r + k]; c->J1(r, k) = J1[18 * i + 6 * r + k];
                c->J0Minv(r, k) = M0[18 * i + 6 * r + k]; c->J1Minv(r, k) = M1[18 * i + 6 * r + k];
            }
        out.push_back(c);
    }
}

void RigidBodySystem::readBack(bool contacts)
{
    const int n = (int)m_bodies.size(), nj = (int)m_joints.size();
    std::vector<float> st(13 * n);
    float *x = st.data(), *q = x + 3 * n, *v = q + 4 * n, *w = v + 3 * n;
    check(slrbs_download_state(m_ctx, 0, 1, n, x, q, v, w), "slrbs_download_state");
    std::vector<float> f(3 * n), tau(3 * n), fc(3 * n), tauc(3 * n), I(9 * n), Ii(9 * n);
    check(slrbs_download_body_dyn(m_ctx, 0, n, f.data(), tau.data(), fc.data(), tauc.data(), I.data(), Ii.data()), "slrbs_download_body_dyn");
    for (int i = 0; i < n; ++i) {
        RigidBody* b = m_bodies[i];
        b->x = Eigen::Vector3f(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
        b->q = Eigen::Quaternionf(q[4 * i + 3], q[4 * i], q[4 * i + 1], q[4 * i + 2]);
        b->xdot = Eigen::Vector3f(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        b->omega = Eigen::Vector3f(w[3 * i], w[3 * i + 1], w[3 * i + 2]);
        b->f = Eigen::Vector3f(f[3 * i], f[3 * i + 1], f[3 * i + 2]);
        b->tau = Eigen::Vector3f(tau[3 * i], tau[3 * i + 1], tau[3 * i + 2]);
        b->fc = Eigen::Vector3f(fc[3 * i], fc[3 * i + 1], fc[3 * i + 2]);
        b->tauc = Eigen::Vector3f(tauc[3 * i], tauc[3 * i + 1], tauc[3 * i + 2]);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { b->Iinv(r, c) = Ii[9 * i + 3 * r + c]; if (!b->fixed) b->I(r, c) = I[9 * i + 3 * r + c]; }
    }
    m_stateSnapshot = st;
    if (nj > 0) {
        std::vector<float> lam(5 * nj);
        check(slrbs_download_joint_lambda(m_ctx, 0, 1, nj, lam.data()), "slrbs_download_joint_lambda");
        for (int i = 0; i < nj; ++i) for (unsigned k = 0; k < m_joints[i]->dim; ++k) m_joints[i]->lambda(k) = lam[5 * i + k];
        m_lambdaSnapshot = lam;
        if (m_rowReadback) {
            std::vector<float> J0(30 * nj), J1(30 * nj), M0(30 * nj), M1(30 * nj);
            check(slrbs_download_joint_rows(m_ctx, 0, nj, J0.data(), J1.data(), M0.data(), M1.data(), nullptr), "slrbs_download_joint_rows");
            for (int i = 0; i < nj; ++i) for (unsigned r = 0; r < m_joints[i]->dim; ++r) for (int k = 0; k < 6; ++k) {
                Joint* j = m_joints[i];
                j->J0(r, k) = J0[30 * i + 6 * r + k]; j->J1(r, k) = J1[30 * i + 6 * r + k];
                j->J0Minv(r, k) = M0[30 * i + 6 * r + k]; j->J1Minv(r, k) = M1[30 * i + 6 * r + k];
            }
        }
    }
    if (contacts) mirrorContacts();
    else m_collisionDetect->clear();
}

void RigidBodySystem::step(float dt)
{
    if (m_bodies.empty()) return;
    m_lastDt = dt;
    std::vector<float> cbF, cbTau;
    if (m_preStepFunc) {
        // slow path: the callback sees the reset forces and fresh inertias on the host
        // (RigidBodySystem.cpp:58-73) and may change any field; the f / tau it leaves are handed
        // to the device verbatim (mode 1)
        hostForceReset();
        computeInertias();
        m_preStepFunc(m_bodies);
        const size_t n = m_bodies.size();
        cbF.resize(3 * n); cbTau.resize(3 * n);
        for (size_t i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) { cbF[3 * i + k] = m_bodies[i]->f[k]; cbTau[3 * i + k] = m_bodies[i]->tau[k]; }
    }
    for (int attempt = 0; attempt < 8; ++attempt) {
        ensureContext();
        uploadStateIfChanged();
        const int n = (int)m_bodies.size();
        if (m_preStepFunc) {
            check(slrbs_set_external_forces(m_ctx, 0, 1, n, cbF.data(), cbTau.data(), 1), "slrbs_set_external_forces");
            m_extActive = true;
        } else if (m_extActive) {
            check(slrbs_set_external_forces(m_ctx, 0, 1, n, nullptr, nullptr, 0), "slrbs_set_external_forces");
            m_extActive = false;
        }

        check(slrbs_set_params(m_ctx, Contact::mu, solverId, solverIter, m_collisionsEnabled ? 1 : 0), "slrbs_set_params");
        check(slrbs_set_extensions(m_ctx, m_extContacts ? 1 : 0), "slrbs_set_extensions");
        check(slrbs_stage_prepare(m_ctx), "slrbs_stage_prepare");
        if (m_collisionsEnabled) {
            m_collisionDetect->detectCollisions();
            m_collisionDetect->computeContactJacobians();      // also assembles the joint rows
        } else {
            deviceContactJacobians(dt);
        }
        // calcConstraintForces (RigidBodySystem.cpp:177-183): the selected Solver object runs the solve
        Solver* solver = m_solvers[std::min(std::max(solverId, 0), 3)].get();
        solver->setMaxIter(solverIter);
        if (solver->deviceSolverId() >= 0) solver->solve(dt);
        else if (!hostSolve(solver, dt)) { m_contactCapacity *= 2; m_topologyDirty = true; continue; }
        check(slrbs_stage_integrate(m_ctx, dt), "slrbs_stage_integrate");

        const int rc = slrbs_sync(m_ctx);
        if (rc == SLRBS_E_CAPACITY) {
            // more contacts than the context was sized for: grow and redo the step from the
            // (still unchanged) host state
            m_contactCapacity *= 2;
            m_topologyDirty = true;
            continue;
        }
        check(rc, "slrbs_sync");
        readBack(m_contactReadback);
        return;
    }
    throw std::runtime_error(std::string("RigidBodySystem::step: capacity kept overflowing: ") + slrbs_last_error());
}

// A Solver subclass without a device kernel (deviceSolverId() < 0, i.e. one written against the reference's
// include/solvers/Solver.h): give it what SolverBoxPGS::solve reads in the reference -- every body's f, tau, xdot, omega,
// Iinv and contact / joint lists, every constraint's J0, J1, J0Minv, J1Minv, phi, lambda (SolverBoxPGS.cpp:27-82) -- run
// solve(h) on the host, then hand the lambdas back: zero device sweeps leave them as uploaded and scatter the constraint
// forces (RigidBodySystem.cpp:185-220) ahead of the integration.
bool RigidBodySystem::hostSolve(Solver* solver, float dt)
{
    const int n = (int)m_bodies.size(), nj = (int)m_joints.size();
    const int rc0 = slrbs_sync(m_ctx);
    if (rc0 == SLRBS_E_CAPACITY) return false;                   // more contacts than the context holds: the caller grows it and redoes the step
    check(rc0, "slrbs_sync");
    std::vector<float> f(3 * n), tau(3 * n), I(9 * n), Ii(9 * n);
    check(slrbs_download_body_dyn(m_ctx, 0, n, f.data(), tau.data(), nullptr, nullptr, I.data(), Ii.data()), "slrbs_download_body_dyn");
    for (int i = 0; i < n; ++i) {
        RigidBody* b = m_bodies[i];
        b->f = Eigen::Vector3f(f[3 * i], f[3 * i + 1], f[3 * i + 2]);
        b->tau = Eigen::Vector3f(tau[3 * i], tau[3 * i + 1], tau[3 * i + 2]);
        b->fc.setZero(); b->tauc.setZero();
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { b->Iinv(r, c) = Ii[9 * i + 3 * r + c]; if (!b->fixed) b->I(r, c) = I[9 * i + 3 * r + c]; }
    }
    const bool rows = m_rowReadback;
    m_rowReadback = true;                                        // J blocks exactly as the device assembled them
    mirrorContacts();
    m_rowReadback = rows;
    if (nj > 0) {
        std::vector<float> J0(30 * nj), J1(30 * nj), M0(30 * nj), M1(30 * nj);
        check(slrbs_download_joint_rows(m_ctx, 0, nj, J0.data(), J1.data(), M0.data(), M1.data(), nullptr), "slrbs_download_joint_rows");
        for (int i = 0; i < nj; ++i) {
            Joint* j = m_joints[i];
            j->computeJacobian();                                // phi (and the blocks) from the host fields ...
            for (unsigned r = 0; r < j->dim; ++r) for (int k = 0; k < 6; ++k) {   // ... the blocks then taken from the device
                j->J0(r, k) = J0[30 * i + 6 * r + k]; j->J1(r, k) = J1[30 * i + 6 * r + k];
                j->J0Minv(r, k) = M0[30 * i + 6 * r + k]; j->J1Minv(r, k) = M1[30 * i + 6 * r + k];
            }
        }
    }
    solver->solve(dt);
    auto& contacts = m_collisionDetect->getContacts();
    if (!contacts.empty()) {
        std::vector<float> lam(3 * contacts.size());
        for (size_t i = 0; i < contacts.size(); ++i) for (int k = 0; k < 3; ++k) lam[3 * i + k] = contacts[i]->lambda(k);
        check(slrbs_upload_contact_lambda(m_ctx, 0, (int)contacts.size(), lam.data()), "slrbs_upload_contact_lambda");
    }
    if (nj > 0) {
        std::vector<float> lam; packLambda(lam);
        check(slrbs_upload_joint_lambda(m_ctx, 0, 1, nj, lam.data()), "slrbs_upload_joint_lambda");
        m_lambdaSnapshot = lam;
    }
    deviceSolve(0, 0, dt);                                       // zero sweeps: force scatter of the uploaded impulses
    return true;
}

// ---- CollisionDetect / Solver seams ---------------------------------------------------------------
void CollisionDetect::detectCollisions() { m_rigidBodySystem->deviceDetect(); }
void CollisionDetect::computeContactJacobians() { m_rigidBodySystem->deviceContactJacobians(m_rigidBodySystem->lastDt()); }
void CollisionDetect::clear()
{
    for (Contact* c : m_contacts) delete c;
    m_contacts.clear();
    for (RigidBody* b : m_rigidBodySystem->getBodies()) b->contacts.clear();
}

void SolverBoxPGS::solve(float h) { m_rigidBodySystem->deviceSolve(0, m_maxIter, h); }
void SolverConjGradient::solve(float h) { m_rigidBodySystem->deviceSolve(1, m_maxIter, h); }
void SolverConjResidual::solve(float h) { m_rigidBodySystem->deviceSolve(2, m_maxIter, h); }
void SolverBPP::solve(float h) { m_rigidBodySystem->deviceSolve(3, m_maxIter, h); }

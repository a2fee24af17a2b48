// RigidBodySystemBatch.cpp -- N copies of one RigidBodySystem in one device context (include/slrbs_b200.h);
// see rigidbody/RigidBodySystemBatch.h.
#include "rigidbody/RigidBodySystemBatch.h"

#include "collision/SDFMesh.h"
#include "contact/Contact.h"
#include "joint/Joint.h"
#include "rigidbody/RigidBody.h"

#include "slrbs_b200.h"

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>

void RigidBodySystemBatch::check(int rc, const char* what) const
{
    if (rc < 0) throw std::runtime_error(std::string(what) + ": " + slrbs_last_error());
}

RigidBodySystemBatch::RigidBodySystemBatch(const RigidBodySystem& proto, int nScenes, int device, int contactsPerScene)
    : solverIter(proto.solverIter), solverId(proto.solverId), mu(Contact::mu), m_scenes(nScenes), m_device(device)
{
    const auto& bodies = proto.getBodies();
    const auto& joints = proto.getJoints();
    m_nb = (int)bodies.size(); m_nj = (int)joints.size();
    if (nScenes <= 0 || m_nb == 0) throw std::runtime_error("RigidBodySystemBatch: needs at least one scene and one body");
    const int n = m_nb, nj = m_nj;
    const int cap = contactsPerScene > 0 ? contactsPerScene : std::max(64, 8 * n);
    check(slrbs_create(&m_ctx, device, nScenes, n, nj, cap), "slrbs_create");
    try {
        // mesh shapes first, in order of first appearance (their slots in the device shape table)
        std::vector<const SDFShape*> shapes;
        for (const RigidBody* b : bodies) {
            if (b->geometry->getType() != kSDFMesh) continue;
            const SDFShape* sh = static_cast<const SDFMesh*>(b->geometry.get())->shape.get();
            auto it = std::find(shapes.begin(), shapes.end(), sh);
            sh->shapeId = (int)(it - shapes.begin());
            if (it == shapes.end()) {
                shapes.push_back(sh);
                check(slrbs_sdf_add_shape(m_ctx, sh->nx, sh->ny, sh->nz, sh->origin, sh->cell, sh->grid.data(), (int)sh->verts.size() / 3, sh->verts.data()),
                      "slrbs_sdf_add_shape");
            }
        }
        // one scene's arrays, replicated nScenes times (the C ABI takes dense [scene][body][k] arrays)
        std::vector<float> x(3 * n), q(4 * n), v(3 * n), w(3 * n), mass(n), ib(9 * n), ibi(9 * n), geom(6 * n, 0.f);
        std::vector<int32_t> fixed(n), type(n);
        std::unordered_map<const RigidBody*, int> index;
        for (int i = 0; i < n; ++i) {
            const RigidBody* b = bodies[i];
            index[b] = i;
            for (int k = 0; k < 3; ++k) { x[3 * i + k] = b->x[k]; v[3 * i + k] = b->xdot[k]; w[3 * i + k] = b->omega[k]; }
            q[4 * i] = b->q.x(); q[4 * i + 1] = b->q.y(); q[4 * i + 2] = b->q.z(); q[4 * i + 3] = b->q.w();
            mass[i] = b->mass; fixed[i] = b->fixed ? 1 : 0; type[i] = (int)b->geometry->getType();
            b->geometry->packParams(&geom[6 * i]);
            std::memcpy(&ib[9 * i], b->Ibody.data(), 9 * sizeof(float));
            std::memcpy(&ibi[9 * i], b->IbodyInv.data(), 9 * sizeof(float));
        }
        auto rep = [nScenes](const auto& one) {
            typename std::decay<decltype(one)>::type all(one.size() * (size_t)nScenes);
            for (int s = 0; s < nScenes; ++s) std::copy(one.begin(), one.end(), all.begin() + (size_t)s * one.size());
            return all;
        };
        check(slrbs_upload_bodies(m_ctx, 0, nScenes, n, nullptr, rep(x).data(), rep(q).data(), rep(v).data(), rep(w).data(), rep(mass).data(),
                                  rep(ib).data(), rep(ibi).data(), rep(fixed).data(), rep(type).data(), rep(geom).data()),
              "slrbs_upload_bodies");
        if (nj > 0) {
            std::vector<int32_t> jt(nj), b0(nj), b1(nj);
            std::vector<float> r0(3 * nj), r1(3 * nj), q0(4 * nj), q1(4 * nj), lam(5 * nj, 0.f);
            for (int i = 0; i < nj; ++i) {
                const Joint* j = joints[i];
                jt[i] = (int)j->getType(); b0[i] = index.at(j->body0); b1[i] = index.at(j->body1);
                for (int k = 0; k < 3; ++k) { r0[3 * i + k] = j->r0[k]; r1[3 * i + k] = j->r1[k]; }
                q0[4 * i] = j->q0.x(); q0[4 * i + 1] = j->q0.y(); q0[4 * i + 2] = j->q0.z(); q0[4 * i + 3] = j->q0.w();
                q1[4 * i] = j->q1.x(); q1[4 * i + 1] = j->q1.y(); q1[4 * i + 2] = j->q1.z(); q1[4 * i + 3] = j->q1.w();
                for (unsigned k = 0; k < j->dim && k < 5; ++k) lam[5 * i + k] = j->lambda(k);
            }
            check(slrbs_upload_joints(m_ctx, 0, nScenes, nj, nullptr, rep(jt).data(), rep(b0).data(), rep(b1).data(), rep(r0).data(), rep(q0).data(),
                                      rep(r1).data(), rep(q1).data()), "slrbs_upload_joints");
            check(slrbs_upload_joint_lambda(m_ctx, 0, nScenes, nj, rep(lam).data()), "slrbs_upload_joint_lambda");
        }
    } catch (...) {
        slrbs_destroy(m_ctx);
        m_ctx = nullptr;
        throw;
    }
}

RigidBodySystemBatch::~RigidBodySystemBatch() { if (m_ctx) slrbs_destroy(m_ctx); }

void RigidBodySystemBatch::pushParams()
{
    check(slrbs_set_params(m_ctx, mu, solverId, solverIter, m_collisions ? 1 : 0), "slrbs_set_params");
    check(slrbs_set_extensions(m_ctx, m_ext ? 1 : 0), "slrbs_set_extensions");
}

void RigidBodySystemBatch::step(float dt, int nSteps)
{
    pushParams();
    check(slrbs_step(m_ctx, dt, nSteps), "slrbs_step");
}

void RigidBodySystemBatch::sync() { check(slrbs_sync(m_ctx), "slrbs_sync"); }

void RigidBodySystemBatch::setState(int scene0, int ns, const float* x3, const float* q4, const float* v3, const float* w3)
{
    check(slrbs_upload_state(m_ctx, scene0, ns, m_nb, x3, q4, v3, w3), "slrbs_upload_state");
}

void RigidBodySystemBatch::getState(int scene0, int ns, float* x3, float* q4, float* v3, float* w3)
{
    check(slrbs_download_state(m_ctx, scene0, ns, m_nb, x3, q4, v3, w3), "slrbs_download_state");
}

void RigidBodySystemBatch::setExternalForces(int scene0, int ns, const float* f3, const float* tau3, bool absolute)
{
    check(slrbs_set_external_forces(m_ctx, scene0, ns, m_nb, f3, tau3, absolute ? 1 : 0), "slrbs_set_external_forces");
}

void RigidBodySystemBatch::clearExternalForces(int scene0, int ns)
{
    check(slrbs_set_external_forces(m_ctx, scene0, ns, m_nb, nullptr, nullptr, 0), "slrbs_set_external_forces");
}

void RigidBodySystemBatch::saveState() { check(slrbs_save_state(m_ctx), "slrbs_save_state"); }
void RigidBodySystemBatch::resetScenes(int scene0, int ns) { check(slrbs_restore_state(m_ctx, scene0, ns), "slrbs_restore_state"); }
void RigidBodySystemBatch::getContactCounts(int scene0, int ns, int* counts)
{
    check(slrbs_download_contact_counts(m_ctx, scene0, ns, counts), "slrbs_download_contact_counts");
}

void RigidBodySystemBatch::readScene(int scene, RigidBodySystem& into)
{
    auto& bodies = into.getBodies();
    auto& joints = into.getJoints();
    if ((int)bodies.size() != m_nb || (int)joints.size() != m_nj) throw std::runtime_error("RigidBodySystemBatch::readScene: the target system has other bodies / joints");
    const int n = m_nb;
    std::vector<float> x(3 * n), q(4 * n), v(3 * n), w(3 * n);
    getState(scene, 1, x.data(), q.data(), v.data(), w.data());
    for (int i = 0; i < n; ++i) {
        RigidBody* b = bodies[i];
        b->x = Eigen::Vector3f(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
        b->q = Eigen::Quaternionf(q[4 * i + 3], q[4 * i], q[4 * i + 1], q[4 * i + 2]);
        b->xdot = Eigen::Vector3f(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        b->omega = Eigen::Vector3f(w[3 * i], w[3 * i + 1], w[3 * i + 2]);
    }
    if (m_nj > 0) {
        std::vector<float> lam(5 * m_nj);
        check(slrbs_download_joint_lambda(m_ctx, scene, 1, m_nj, lam.data()), "slrbs_download_joint_lambda");
        for (int i = 0; i < m_nj; ++i) for (unsigned k = 0; k < joints[i]->dim && k < 5; ++k) joints[i]->lambda(k) = lam[5 * i + k];
    }
}

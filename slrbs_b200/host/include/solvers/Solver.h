// solvers/Solver.h -- the solver plug-in seam (include/solvers/Solver.h:9-43). The three solver
// classes keep their names; solve(h) runs the selected device kernel on the system's context.
// A subclass written against the reference's header (only solve(h) overridden) compiles unchanged:
// deviceSolverId() then stays -1 and RigidBodySystem::step runs that solver on the HOST -- contacts,
// Jacobian blocks, phi and the bodies' f / tau / Iinv are mirrored into the public fields before
// solve(h), and the lambdas it leaves are sent back for the force scatter and integration
// (RigidBodySystem::setSolver installs it).
#pragma once
#include "util/Types.h"

class RigidBodySystem;

class Solver {
public:
    Solver(RigidBodySystem* _rigidBodySystem) : m_rigidBodySystem(_rigidBodySystem), m_maxIter(20) {}
    virtual ~Solver() {}
    void setMaxIter(int _maxIter) { m_maxIter = _maxIter; }
    int getMaterIter() const { return m_maxIter; }      // sic: the reference's spelling
    virtual void solve(float h) = 0;
    // >= 0: the device kernel that implements this solver (0 Box-PGS, 1 CG, 2 CR, 3 BPP); -1: solve(h) runs on the host
    virtual int deviceSolverId() const { return -1; }
protected:
    RigidBodySystem* m_rigidBodySystem;
    int m_maxIter;
};

// solvers/Solver.h -- the solver plug-in seam (include/solvers/Solver.h:9-43). The three solver
// classes keep their names; solve(h) runs the selected device kernel on the system's context.
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
    virtual int deviceSolverId() const = 0;
protected:
    RigidBodySystem* m_rigidBodySystem;
    int m_maxIter;
};

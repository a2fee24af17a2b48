// solvers/SolverConjGradient.h -- include/solvers/SolverConjGradient.h of the reference; device solver id 1.
#pragma once
#include "solvers/Solver.h"

class SolverConjGradient : public Solver {
public:
    SolverConjGradient(RigidBodySystem* _rigidBodySystem) : Solver(_rigidBodySystem) {}
    void solve(float h) override;
    int deviceSolverId() const override { return 1; }
};

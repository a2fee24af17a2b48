// solvers/SolverConjResidual.h -- include/solvers/SolverConjResidual.h of the reference; device solver id 2.
#pragma once
#include "solvers/Solver.h"

class SolverConjResidual : public Solver {
public:
    SolverConjResidual(RigidBodySystem* _rigidBodySystem) : Solver(_rigidBodySystem) {}
    void solve(float h) override;
    int deviceSolverId() const override { return 2; }
};

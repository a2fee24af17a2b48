// solvers/SolverBoxPGS.h -- include/solvers/SolverBoxPGS.h of the reference; device solver id 0.
#pragma once
#include "solvers/Solver.h"

class SolverBoxPGS : public Solver {
public:
    SolverBoxPGS(RigidBodySystem* _rigidBodySystem) : Solver(_rigidBodySystem) {}
    void solve(float h) override;
    int deviceSolverId() const override { return 0; }
};

// solvers/SolverBPP.h -- EXTENSION, no reference counterpart: block principal pivoting over
// joints and contacts, device solver id 3 (csrc/bpp.cu; algorithm defined by oracle/slrbs_oracle.c solve_bpp).
// Selected like the reference's solvers: RigidBodySystem::solverId = 3; solverIter bounds the pivot steps (5 per
// iteration). Dense per scene: at most 288 scalar constraint rows.
#pragma once
#include "solvers/Solver.h"

class SolverBPP : public Solver {
public:
    SolverBPP(RigidBodySystem* _rigidBodySystem) : Solver(_rigidBodySystem) {}
    void solve(float h) override;
    int deviceSolverId() const override { return 3; }
};

// util/Types.h -- mirrors include/util/Types.h:12-16 of the reference: the 6-wide constraint row
// layout [linear xyz | angular xyz] per body per row.
#pragma once
#include <Eigen/Dense>

typedef Eigen::RowMatrixX6f JBlock;

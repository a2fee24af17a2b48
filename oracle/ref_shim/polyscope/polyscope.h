#pragma once
#include "polyscope/surface_mesh.h"

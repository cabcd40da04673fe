// Internal declarations shared by the 3DGS kernels.
#pragma once
#include "common.cuh"
#include "../../include/starst3r_b200.h"

int gs_tile_bits(int n_tiles);

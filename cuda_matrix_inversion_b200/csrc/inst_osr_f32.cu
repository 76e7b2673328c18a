#define INVGPU_TILE_DEFINE
#include "tile_launch.cuh"
#include "tile_configs.h"
INVGPU_OSR_F32(INVGPU_ONESWEEP_ROLLED_INSTANTIATE)

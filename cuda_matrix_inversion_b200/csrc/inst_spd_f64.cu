#define INVGPU_TILE_DEFINE
#include "tile_launch.cuh"
#include "tile_configs.h"
INVGPU_TILE_SPD_F64_INV(INVGPU_TILE_INSTANTIATE)

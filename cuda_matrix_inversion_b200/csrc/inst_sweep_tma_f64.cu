#define INVGPU_TILE_DEFINE
#include "tile_launch.cuh"
#include "tile_configs.h"
INVGPU_SWEEP_TMA_F64(INVGPU_SWEEP_TMA_INSTANTIATE)

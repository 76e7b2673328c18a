// fast_tiers.cuh -- dispatch into the register-tiled tiers (warp tier n <= 32, CTA tier
// n in {64,128}); returns INVGPU_NO_FAST_PATH when a shape has no specialised kernel, in which
// case the any-n shared-memory kernels of generic_smem.cuh serve it.
#pragma once

#include "engine.cuh"

#define INVGPU_NO_FAST_PATH (-1000)

namespace invgpu {

template <typename T, typename IO, int STAGES>
static int fast_spd(IO, int, i64, int *, cudaStream_t, DeviceState *) { return INVGPU_NO_FAST_PATH; }

template <typename T, typename IO>
static int fast_general(IO, int, i64, int *, cudaStream_t, DeviceState *) { return INVGPU_NO_FAST_PATH; }

template <typename T>
static int fast_gp(GpIO<T>, int, i64, int *, cudaStream_t, DeviceState *) { return INVGPU_NO_FAST_PATH; }

static const char *fast_tier_name(int, int, int) { return "generic"; }

}  // namespace invgpu

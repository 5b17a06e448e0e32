// Activation / dropout helpers shared by the layer kernels (mlp.cu, layer_ops.cu).
#pragma once
#include "common.cuh"

namespace fr {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_SIGMOID = 3, ACT_TANH = 4 };

// counter-based dropout mask: keep-scale of element idx of layer `layer` (1/(1-p) or 0); p == 0 -> 1
__device__ __forceinline__ float drop_scale(unsigned long long seed, uint32_t layer, uint32_t idx, float p) {
  if (p <= 0.f) return 1.f;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)layer << 32 | idx);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = (float)(z >> 40) * (1.0f / 16777216.0f);
  return u < p ? 0.f : 1.f / (1.f - p);
}

__device__ __forceinline__ float act_fwd(float x, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(x, 0.f);
    case ACT_LEAKY: return x > 0.f ? x : 0.01f * x;
    case ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case ACT_TANH: return tanhf(x);
    default: return x;
  }
}
// derivative expressed through the activation OUTPUT y (what the forward pass keeps)
__device__ __forceinline__ float act_bwd(float y, int act) {
  switch (act) {
    case ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case ACT_LEAKY: return y > 0.f ? 1.f : 0.01f;
    case ACT_SIGMOID: return y * (1.f - y);
    case ACT_TANH: return 1.f - y * y;
    default: return 1.f;
  }
}

}  // namespace fr

// Epilogue math shared by the tcgen05 GEMM kernels.
#pragma once
#include <cuda_fp16.h>

namespace mdm {
// Exact-erf GELU (nn.GELU() default, unet.py:270) with erf from Abramowitz-Stegun 7.1.26
// (|abs error| <= 1.5e-7, far below the fp16 rounding of the stored result); one ex2 + one rcp.
__device__ __forceinline__ float gelu_erf(float v) {
  const float x = fabsf(v) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.0f - poly * t * __expf(-x * x);
  const float erf_v = copysignf(erf_abs, v);
  return 0.5f * v * (1.0f + erf_v);
}

// d/dv of the exact-erf GELU, same erf approximation as gelu_erf
__device__ __forceinline__ float gelu_grad(float v) {
  const float x = fabsf(v) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float ex = __expf(-x * x);
  const float erf_v = copysignf(1.0f - poly * t * ex, v);
  return 0.5f * (1.0f + erf_v) + v * 0.39894228040143267794f * ex;  // cdf + v * pdf
}


}  // namespace mdm

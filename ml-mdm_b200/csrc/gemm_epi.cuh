// Epilogue math shared by the tcgen05 GEMM kernels.
#pragma once
#include <cuda_fp16.h>

namespace mdm {
// SFU primitives without the range/denormal guard code of __fdividef / __expf: the arguments here are
// bounded (1 <= 1 + p x, -x^2 <= 0) and a result below 2^-126 may flush to zero.
__device__ __forceinline__ float sfu_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sfu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// erf(|v| / sqrt 2) by Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7, far below the fp16 rounding of
// the stored result) and exp(-v^2 / 2); one rcp + one ex2.
__device__ __forceinline__ void erf_abs_and_gauss(float v, float& erf_abs, float& gauss) {
  const float x = fabsf(v) * 0.70710678118654752440f;
  const float t = sfu_rcp(fmaf(0.3275911f, x, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  gauss = sfu_ex2(x * -1.4426950408889634f * x);
  erf_abs = fmaf(-poly * t, gauss, 1.0f);
}

// Exact-erf GELU (nn.GELU() default, unet.py:270): 0.5 v (1 + erf(v / sqrt 2))
__device__ __forceinline__ float gelu_erf(float v) {
  float e, g;
  erf_abs_and_gauss(v, e, g);
  return 0.5f * fmaf(fabsf(v), e, v);  // v * sign(v) erf_abs = |v| erf_abs
}

// d/dv of the exact-erf GELU, same erf approximation as gelu_erf: cdf + v * pdf
__device__ __forceinline__ float gelu_grad(float v) {
  float e, g;
  erf_abs_and_gauss(v, e, g);
  const float cdf = fmaf(0.5f, copysignf(e, v), 0.5f);
  return fmaf(v * 0.39894228040143267794f, g, cdf);
}

}  // namespace mdm

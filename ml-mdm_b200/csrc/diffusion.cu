// Diffusion algebra of the path as single-pass kernels over NCHW fp32 images with PER-SAMPLE
// gamma scalars (the reference materialises gamma as full (B,C,H,W) maps, samplers.py:196-199).
//   q-sample            samplers.py:244-246 (+ NestedSampler.get_xt :625-637)
//   targets / loss      samplers.py:266-279,347-390 ; diffusion.py:123-136,160-168,367-386
//   reverse step        samplers.py:281-345 (DDPM posterior mean or DDIM(eta)), clip :500-508
//   CFG combine         samplers.py:449-455
//   avg_pool pyramid    diffusion.py:346
#include <math.h>

#include "engine.cuh"
#include "mdm_b200.h"

namespace mdm {
namespace {

enum { PT_DDPM = 3, PT_DDIM = 4, PT_V = 5 };  // values of samplers.PredictionType

__device__ __forceinline__ float x0_from_pred(int ptype, float xt, float pred, float a, float c) {
  // a = sqrt(g), c = sqrt(1-g)   (samplers.py:359-367)
  if (ptype == PT_V) return xt * a - pred * c;
  return (xt - pred * c) / a;
}
__device__ __forceinline__ float pred_from_x0(int ltype, float xt, float x0, float a, float c) {
  // samplers.py:381-389
  if (ltype == PT_V) return (a * xt - x0) / c;
  return (xt - x0 * a) / c;
}

__global__ void q_sample_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                const long long* __restrict__ t, const float* __restrict__ gammas, int t_off,
                                float image_div, float* __restrict__ xt, long long per, long long total) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long gs = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += gs) {
    const int b = static_cast<int>(i / per);
    const float g = gammas[t[b] + t_off];
    float xi = x[i];
    if (image_div != 1.0f) xi = xi / image_div;
    // no FMA contraction: bit-identical to the reference's separate fp32 mul/add (samplers.py:245)
    xt[i] = __fadd_rn(__fmul_rn(sqrtf(g), xi), __fmul_rn(sqrtf(1.0f - g), eps[i]));
  }
}

// loss[b] = mean_chw (pred_loss - target)^2 * weight ; optionally writes pred_loss / target.
// One block per (sample, chunk); partial sums by atomics into loss (zeroed by the launcher).
__global__ void __launch_bounds__(256)
loss_fwd_kernel(const float* __restrict__ model_out, const float* __restrict__ xt, const float* __restrict__ x,
                const float* __restrict__ eps, const long long* __restrict__ t, const float* __restrict__ gammas,
                int ptype, int ltype, float image_div, float weight, float* __restrict__ loss,
                float* __restrict__ pred_out, float* __restrict__ tgt_out, long long per) {
  const int b = blockIdx.y;
  const float g = gammas[t[b] + 1];
  const float a = sqrtf(g), c = sqrtf(1.0f - g);
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < per;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long k = static_cast<long long>(b) * per + i;
    const float v = model_out[k], xti = xt[k];
    float xi = x[k];
    if (image_div != 1.0f) xi = xi / image_div;
    const float tgt = (ltype == PT_V) ? (a * eps[k] - c * xi) : eps[k];
    float p = v;
    if (ltype != ptype) p = pred_from_x0(ltype, xti, x0_from_pred(ptype, xti, v, a, c), a, c);
    if (pred_out != nullptr) pred_out[k] = p;
    if (tgt_out != nullptr) tgt_out[k] = tgt;
    const float d = p - tgt;
    acc += d * d;
  }
  __shared__ float sh[8];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += sh[w];
    atomicAdd(&loss[b], weight * s / static_cast<float>(per));
  }
}

// d model_out = dloss[b] * weight * 2 (p - tgt) / per * dp/dv
__global__ void loss_bwd_kernel(const float* __restrict__ model_out, const float* __restrict__ xt,
                                const float* __restrict__ x, const float* __restrict__ eps,
                                const long long* __restrict__ t, const float* __restrict__ gammas, int ptype,
                                int ltype, float image_div, float weight, const float* __restrict__ dloss,
                                float* __restrict__ dout, long long per, long long total) {
  long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long gs = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; k < total; k += gs) {
    const int b = static_cast<int>(k / per);
    const float g = gammas[t[b] + 1];
    const float a = sqrtf(g), c = sqrtf(1.0f - g);
    const float v = model_out[k], xti = xt[k];
    float xi = x[k];
    if (image_div != 1.0f) xi = xi / image_div;
    const float tgt = (ltype == PT_V) ? (a * eps[k] - c * xi) : eps[k];
    float p = v, dpdv = 1.0f;
    if (ltype != ptype) {
      p = pred_from_x0(ltype, xti, x0_from_pred(ptype, xti, v, a, c), a, c);
      const float dx0dv = (ptype == PT_V) ? -c : (-c / a);
      const float dpdx0 = (ltype == PT_V) ? (-1.0f / c) : (-a / c);
      dpdv = dpdx0 * dx0dv;
    }
    dout[k] = dloss[b] * weight * 2.0f * (p - tgt) / static_cast<float>(per) * dpdv;
  }
}

// One reverse step for one resolution level (samplers.py:281-345).
//   mode 0: DDPM posterior mean (ddim_eta is None); mode 1: DDIM with eta (>= 0)
__global__ void sampler_step_kernel(const float* __restrict__ xt, const float* __restrict__ pred,
                                    const float* __restrict__ noise, const float* __restrict__ gammas, int t_idx,
                                    int s_idx, int ptype, int clip, float image_scale, int mode, float eta,
                                    int need_noise, float* __restrict__ x0_out, float* __restrict__ xs_out,
                                    long long total) {
  const float g = gammas[t_idx], gl = gammas[s_idx];
  const float alpha = g / gl;
  const float beta = 1.0f - alpha;
  float beta_tilde = beta * (1.0f - gl) / (1.0f - g);
  const float a = sqrtf(g), c = sqrtf(1.0f - g);
  long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long gs = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; k < total; k += gs) {
    const float x = xt[k];
    float x0 = x0_from_pred(ptype, x, pred[k], a, c);
    if (clip) x0 = fminf(fmaxf(x0 * image_scale, -1.0f), 1.0f) / image_scale;
    float xs;
    float bt = beta_tilde;
    int nn = need_noise;
    if (mode == 0) {
      xs = x0 * beta * sqrtf(gl) / (1.0f - g) + x * sqrtf(alpha) * (1.0f - gl) / (1.0f - g);
    } else {
      const float e = (x - x0 * a) / c;
      if (eta > 0.f) {
        bt = (eta * eta) * beta_tilde;
        xs = x0 * sqrtf(gl) + e * sqrtf(1.0f - gl - bt);
      } else {
        nn = 0;
        xs = x0 * sqrtf(gl) + e * sqrtf(1.0f - gl);
      }
    }
    if (nn) xs = xs + sqrtf(bt) * noise[k];
    if (x0_out != nullptr) x0_out[k] = x0;
    xs_out[k] = xs;
  }
}

__global__ void cfg_combine_kernel(const float* __restrict__ uncond, const float* __restrict__ cond, float w,
                                   float* __restrict__ out, long long n) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long gs = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n; i += gs) out[i] = uncond[i] + w * (cond[i] - uncond[i]);
}

__global__ void avg_pool_kernel(const float* __restrict__ x, float* __restrict__ y, int planes, int H, int W,
                                int r) {
  const int Ho = H / r, Wo = W / r;
  const long long total = static_cast<long long>(planes) * Ho * Wo;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long gs = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += gs) {
    const int wo = static_cast<int>(i % Wo);
    const long long tt = i / Wo;
    const int ho = static_cast<int>(tt % Ho);
    const long long pl = tt / Ho;
    const float* src = x + (pl * H + static_cast<long long>(ho) * r) * W + static_cast<long long>(wo) * r;
    float s = 0.f;
    for (int a = 0; a < r; ++a)
      for (int b = 0; b < r; ++b) s += src[static_cast<long long>(a) * W + b];
    y[i] = s / static_cast<float>(r * r);
  }
}

__global__ void clip_scale_kernel(const float* __restrict__ x, float scale, int clip, float* __restrict__ y,
                                  long long n) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long gs = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n; i += gs) {
    float v = x[i] * scale;
    if (clip) v = fminf(fmaxf(v, -1.0f), 1.0f);
    y[i] = v;
  }
}

inline int grid_for(long long n) {
  long long g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace
}  // namespace mdm

#define MDM_TRY(...)                     \
  try {                                  \
    __VA_ARGS__;                         \
    return 0;                            \
  } catch (const std::exception& e) {    \
    mdm::set_error("%s", e.what());      \
    return -1;                           \
  }

extern "C" {

int mdm_q_sample(const float* x, const float* eps, const int64_t* t, const float* gammas, int t_offset,
                 float image_div, float* x_t, int batch, int64_t per_sample, mdm_stream_t stream) {
  MDM_TRY({
    const long long total = static_cast<long long>(batch) * per_sample;
    mdm::q_sample_kernel<<<mdm::grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, eps, reinterpret_cast<const long long*>(t), gammas, t_offset, image_div, x_t, per_sample, total);
    ++mdm::g_launch_count;
    MDM_CUDA(cudaGetLastError());
  })
}

int mdm_loss_fwd(const float* model_out, const float* x_t, const float* x, const float* eps, const int64_t* t,
                 const float* gammas, int prediction_type, int loss_type, float image_div, float weight,
                 float* loss, float* pred_out, float* tgt_out, int batch, int64_t per_sample, mdm_stream_t stream) {
  MDM_TRY({
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int chunks = static_cast<int>(std::min<long long>((per_sample + 1023) / 1024, 64));
    if (chunks < 1) chunks = 1;
    dim3 grid(chunks, batch);
    mdm::loss_fwd_kernel<<<grid, 256, 0, st>>>(model_out, x_t, x, eps, reinterpret_cast<const long long*>(t), gammas,
                                               prediction_type, loss_type, image_div, weight, loss, pred_out, tgt_out,
                                               per_sample);
    ++mdm::g_launch_count;
    MDM_CUDA(cudaGetLastError());
  })
}

int mdm_loss_bwd(const float* model_out, const float* x_t, const float* x, const float* eps, const int64_t* t,
                 const float* gammas, int prediction_type, int loss_type, float image_div, float weight,
                 const float* dloss, float* dmodel_out, int batch, int64_t per_sample, mdm_stream_t stream) {
  MDM_TRY({
    const long long total = static_cast<long long>(batch) * per_sample;
    mdm::loss_bwd_kernel<<<mdm::grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        model_out, x_t, x, eps, reinterpret_cast<const long long*>(t), gammas, prediction_type, loss_type, image_div,
        weight, dloss, dmodel_out, per_sample, total);
    ++mdm::g_launch_count;
    MDM_CUDA(cudaGetLastError());
  })
}

int mdm_sampler_step(const float* x_t, const float* pred, const float* noise, const float* gammas, int t_index,
                     int s_index, int prediction_type, int clip, float image_scale, int use_ddim, float ddim_eta,
                     int need_noise, float* x0_out, float* x_s_out, int64_t numel, mdm_stream_t stream) {
  MDM_TRY({
    if (need_noise && !(use_ddim && ddim_eta <= 0.f) && noise == nullptr)
      throw mdm::MdmFail("mdm_sampler_step: noise tensor required for a stochastic step");
    mdm::sampler_step_kernel<<<mdm::grid_for(numel), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x_t, pred, noise, gammas, t_index, s_index, prediction_type, clip, image_scale, use_ddim ? 1 : 0, ddim_eta,
        need_noise, x0_out, x_s_out, numel);
    ++mdm::g_launch_count;
    MDM_CUDA(cudaGetLastError());
  })
}

int mdm_cfg_combine(const float* uncond, const float* cond, float guidance_scale, float* out, int64_t numel,
                    mdm_stream_t stream) {
  MDM_TRY({
    mdm::cfg_combine_kernel<<<mdm::grid_for(numel), 256, 0, static_cast<cudaStream_t>(stream)>>>(uncond, cond,
                                                                                                guidance_scale, out,
                                                                                                numel);
    ++mdm::g_launch_count;
    MDM_CUDA(cudaGetLastError());
  })
}

int mdm_avg_pool(const float* x, float* y, int planes, int H, int W, int ratio, mdm_stream_t stream) {
  MDM_TRY({
    if (ratio < 1 || H % ratio != 0 || W % ratio != 0) throw mdm::MdmFail("mdm_avg_pool: size not divisible");
    const long long total = static_cast<long long>(planes) * (H / ratio) * (W / ratio);
    mdm::avg_pool_kernel<<<mdm::grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, planes, H, W,
                                                                                             ratio);
    ++mdm::g_launch_count;
    MDM_CUDA(cudaGetLastError());
  })
}

int mdm_clip_scale(const float* x, float scale, int clip, float* y, int64_t numel, mdm_stream_t stream) {
  MDM_TRY({
    mdm::clip_scale_kernel<<<mdm::grid_for(numel), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, scale, clip, y,
                                                                                               numel);
    ++mdm::g_launch_count;
    MDM_CUDA(cudaGetLastError());
  })
}

}  // extern "C"

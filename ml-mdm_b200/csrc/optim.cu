// The step on the far side of the denoising path: what the reference does after
// loss.backward() in five separate full-parameter passes --
//   nn.utils.clip_grad_norm_(model.parameters(), clip)        ml_mdm/trainer.py:78-80
//   optimizer.step()   (torch.optim.Adam / AdamW, eps 1e-8)   ml_mdm/trainer.py:81, clis/train_parallel.py:122-134
//   ema_model.update(vision_model)                            ml_mdm/trainer.py:82-85, models/model_ema.py:25-34
//   optimizer.zero_grad()                                     ml_mdm/trainer.py:92-93
// -- as ONE sweep over the flat gradient arena: 20 bytes read + 20 bytes written per parameter (g, p, m, v, ema),
// HBM-bound. The global gradient norm is a separate deterministic two-stage reduction whose result stays on the
// device; the sweep reads it, so the host never synchronises.
#include <cuda_runtime.h>
#include <stdint.h>

#include "engine.cuh"
#include "mdm_b200.h"

static_assert(sizeof(mdm_opt_chunk) == 48, "mdm_opt_chunk layout is part of the ABI (optim.py mirrors it)");
static_assert(sizeof(mdm_adam_cfg) == 72, "mdm_adam_cfg layout is part of the ABI (optim.py mirrors it)");

namespace mdm {

namespace {

constexpr int SQ_THREADS = 256;
constexpr int SQ_BLOCKS = 148 * 8;

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  if (w == 0) {
    t = l < (blockDim.x >> 5) ? sh[l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;  // valid in warp 0
}

// stage 1: fixed assignment of elements to CTAs and a fixed reduction tree -> run-to-run deterministic
__global__ void __launch_bounds__(SQ_THREADS) sqnorm_partial_kernel(const float* __restrict__ g, long long n,
                                                                    double* __restrict__ partials) {
  __shared__ double sh[32];
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  double acc = 0.0;
  for (long long i = static_cast<long long>(blockIdx.x) * SQ_THREADS + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * SQ_THREADS) {
    const float4 v = __ldg(g4 + i);
    acc += static_cast<double>(v.x) * v.x + static_cast<double>(v.y) * v.y + static_cast<double>(v.z) * v.z +
           static_cast<double>(v.w) * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 << 2; i < n; ++i) acc += static_cast<double>(g[i]) * g[i];
  const double t = block_sum_d(acc, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// stage 2: norm = sqrt(sum) * grad_scale  (the gradients are multiplied by grad_scale before clipping)
__global__ void __launch_bounds__(SQ_THREADS) sqnorm_final_kernel(const double* __restrict__ partials, int nparts,
                                                                  float grad_scale, float* __restrict__ out_norm) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nparts; i += SQ_THREADS) acc += partials[i];
  const double t = block_sum_d(acc, sh);
  if (threadIdx.x == 0) out_norm[0] = static_cast<float>(sqrt(t)) * grad_scale;
}

struct SweepScalars {
  float grad_scale, max_norm;  // max_norm <= 0: no clipping
  float one_m_beta1, beta2, one_m_beta2, eps;
  float weight_decay, decay_mul;   // Adam L2 coefficient; AdamW factor 1 - lr * wd
  float step_size, bc2_sqrt;       // lr / (1 - beta1^t), sqrt(1 - beta2^t): computed in double on the host like torch
  float ema_decay, one_m_ema;      // effective decay of this update (0 during EMA warm-up) and 1 - decay
  int adamw, zero_grad;
};

// One element, op for op as torch applies it (torch/optim/adam.py::_single_tensor_adam, no amsgrad / maximize):
// every Python scalar enters as the float torch would cast it to.
__device__ __forceinline__ void adam_elem(float& p, float& g, float& m, float& v, float* ema, float clip,
                                          const SweepScalars& s) {
  float gr = __fmul_rn(g, clip);
  if (s.weight_decay != 0.f) {
    if (s.adamw) p = __fmul_rn(p, s.decay_mul);                   // param.mul_(1 - lr * wd)
    else gr = fmaf(p, s.weight_decay, gr);                       // grad.add(param, alpha = wd)
  }
  m = fmaf(s.one_m_beta1, __fsub_rn(gr, m), m);                   // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(__fmul_rn(s.one_m_beta2, gr), gr, __fmul_rn(v, s.beta2));  // mul_(beta2).addcmul_(g, g, 1 - beta2)
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), s.bc2_sqrt), s.eps);
  p = fmaf(-s.step_size, __fdiv_rn(m, denom), p);                 // param.addcdiv_(exp_avg, denom, value=-step)
  if (ema != nullptr) *ema = fmaf(p, s.one_m_ema, __fmul_rn(*ema, s.ema_decay));  // mul_(d).add_(p, alpha=1-d)
  g = s.zero_grad ? 0.f : gr;
}

__global__ void __launch_bounds__(256) adam_ema_sweep_kernel(const mdm_opt_chunk* __restrict__ chunks,
                                                             const float* __restrict__ norm, SweepScalars s) {
  const mdm_opt_chunk c = chunks[blockIdx.x];
  float clip = s.grad_scale;
  if (s.max_norm > 0.f) {
    // clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1 (torch/nn/utils/clip_grad.py)
    const float coef = s.max_norm / (__ldg(norm) + 1e-6f);
    // torch.clamp(coef, max=1.0) propagates NaN (a NaN norm poisons every gradient, as in the reference); fminf would drop it
    clip = __fmul_rn(s.grad_scale, coef < 1.0f ? coef : (coef != coef ? coef : 1.0f));
  }
  const long long n = c.n;
  const bool vec = ((reinterpret_cast<uintptr_t>(c.p) | reinterpret_cast<uintptr_t>(c.g) |
                     reinterpret_cast<uintptr_t>(c.m) | reinterpret_cast<uintptr_t>(c.v) |
                     reinterpret_cast<uintptr_t>(c.ema)) & 15) == 0;
  const long long n4 = vec ? (n >> 2) : 0;
  for (long long i = threadIdx.x; i < n4; i += 256) {
    float4 p = reinterpret_cast<float4*>(c.p)[i];
    float4 g = reinterpret_cast<float4*>(c.g)[i];
    float4 m = reinterpret_cast<float4*>(c.m)[i];
    float4 v = reinterpret_cast<float4*>(c.v)[i];
    float4 e = make_float4(0, 0, 0, 0);
    if (c.ema != nullptr) e = reinterpret_cast<float4*>(c.ema)[i];
    float* ex = c.ema != nullptr ? &e.x : nullptr;
    adam_elem(p.x, g.x, m.x, v.x, ex, clip, s);
    adam_elem(p.y, g.y, m.y, v.y, ex ? &e.y : nullptr, clip, s);
    adam_elem(p.z, g.z, m.z, v.z, ex ? &e.z : nullptr, clip, s);
    adam_elem(p.w, g.w, m.w, v.w, ex ? &e.w : nullptr, clip, s);
    reinterpret_cast<float4*>(c.p)[i] = p;
    reinterpret_cast<float4*>(c.g)[i] = g;
    reinterpret_cast<float4*>(c.m)[i] = m;
    reinterpret_cast<float4*>(c.v)[i] = v;
    if (c.ema != nullptr) reinterpret_cast<float4*>(c.ema)[i] = e;
  }
  for (long long i = (n4 << 2) + threadIdx.x; i < n; i += 256)
    adam_elem(c.p[i], c.g[i], c.m[i], c.v[i], c.ema != nullptr ? c.ema + i : nullptr, clip, s);
}

}  // namespace
}  // namespace mdm

#define MDM_TRY(...)                  \
  try {                               \
    __VA_ARGS__;                      \
    return 0;                         \
  } catch (const std::exception& e) { \
    mdm::set_error("%s", e.what());   \
    return -1;                        \
  }

extern "C" {

int mdm_grad_norm(const float* grads, int64_t n, float grad_scale, double* scratch, int32_t scratch_elems,
                  float* out_norm, mdm_stream_t stream) {
  MDM_TRY({
    MDM_CHECK(grads != nullptr && out_norm != nullptr && scratch != nullptr, "null pointer");
    MDM_CHECK((reinterpret_cast<uintptr_t>(grads) & 15) == 0, "gradient arena must be 16-byte aligned");
    MDM_CHECK(scratch_elems >= MDM_GRAD_NORM_SCRATCH, "scratch too small (MDM_GRAD_NORM_SCRATCH doubles)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    mdm::sqnorm_partial_kernel<<<mdm::SQ_BLOCKS, mdm::SQ_THREADS, 0, st>>>(grads, n, scratch);
    mdm::sqnorm_final_kernel<<<1, mdm::SQ_THREADS, 0, st>>>(scratch, mdm::SQ_BLOCKS, grad_scale, out_norm);
    mdm::g_launch_count += 2;
    MDM_CUDA(cudaGetLastError());
  })
}

int mdm_adam_ema_sweep(const mdm_opt_chunk* chunks_dev, int32_t nchunks, const mdm_adam_cfg* cfg,
                       const float* norm_dev, mdm_stream_t stream) {
  MDM_TRY({
    MDM_CHECK(chunks_dev != nullptr && cfg != nullptr, "null pointer");
    MDM_CHECK(cfg->step >= 1, "step counts from 1");
    MDM_CHECK(cfg->max_norm <= 0.f || norm_dev != nullptr, "clipping needs the device norm (mdm_grad_norm)");
    if (nchunks <= 0) return 0;
    mdm::SweepScalars s;
    const double bc1 = 1.0 - pow(cfg->beta1, static_cast<double>(cfg->step));
    const double bc2 = 1.0 - pow(cfg->beta2, static_cast<double>(cfg->step));
    s.grad_scale = cfg->grad_scale;
    s.max_norm = cfg->max_norm;
    s.one_m_beta1 = static_cast<float>(1.0 - cfg->beta1);
    s.beta2 = static_cast<float>(cfg->beta2);
    s.one_m_beta2 = static_cast<float>(1.0 - cfg->beta2);
    s.eps = static_cast<float>(cfg->eps);
    s.weight_decay = static_cast<float>(cfg->weight_decay);
    s.decay_mul = static_cast<float>(1.0 - cfg->lr * cfg->weight_decay);
    s.step_size = static_cast<float>(cfg->lr / bc1);
    s.bc2_sqrt = static_cast<float>(sqrt(bc2));
    s.ema_decay = static_cast<float>(cfg->ema_decay);
    s.one_m_ema = static_cast<float>(1.0 - cfg->ema_decay);
    s.adamw = cfg->adamw;
    s.zero_grad = cfg->zero_grad;
    mdm::adam_ema_sweep_kernel<<<nchunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(chunks_dev, norm_dev, s);
    ++mdm::g_launch_count;
    MDM_CUDA(cudaGetLastError());
  })
}
}

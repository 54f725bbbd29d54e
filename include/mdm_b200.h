/*
 * mdm_b200 -- C ABI of the Blackwell-native Matryoshka denoising path.
 *
 * The reference (apple/ml-mdm) is pure Python/PyTorch and has no FFI of its own; the entry points
 * below are what a binding for its hot path would call.  Each one names the reference interface it
 * stands behind (paths relative to ml-mdm-matryoshka/ml_mdm/):
 *
 *   mdm_net_*            models/unet.py:579-987 (UNet), models/nested_unet.py:96-230 (NestedUNet):
 *                        construction from UNetConfig / NestedUNetConfig, forward(), and the autograd
 *                        backward that `loss.backward()` (trainer.py:46,75) runs through it.
 *   mdm_gammas_*         samplers.py:126-170,201-231,255-264 (noise schedules, shifted schedule)
 *   mdm_set_timesteps    samplers.py:601-609
 *   mdm_q_sample*, mdm_loss*   samplers.py:233-279, diffusion.py:123-168,315-387
 *   mdm_sampler_step     samplers.py:281-345 (get_prediction_xt_last) as used by :392-433,:655-713
 *   mdm_op_*             single fused operators, exported so parity tests can pin each kernel
 *
 * Conventions: every function returns 0 on success and a negative code on failure; the message is
 * available from mdm_last_error().  Pointers are raw device pointers unless named host_*.  Nothing
 * is allocated on the hot path after the first call with a given shape.  All work is enqueued on
 * the cudaStream_t passed by the caller (pass torch.cuda.current_stream().cuda_stream).
 * There is no CPU fallback: without an sm_100a device every compute entry point fails.
 */
#ifndef MDM_B200_H_
#define MDM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mdm_stream_t; /* cudaStream_t */

const char* mdm_last_error(void);
int mdm_version(void);
/* Number of CUDA kernels this library has launched since load (bench.py's gpu_launches). */
unsigned long long mdm_launch_count(void);

/* ---------------------------------------------------------------- low-level tcgen05 GEMM engine */

/* 4-D TMA view of an fp16 operand; dims[0] is contiguous.
 * rows mode: (inner, rows, z1, z2); patch mode: (channels, W, H, image). Strides in elements. */
typedef struct mdm_tmap_spec {
  const void* ptr;
  uint64_t dims[4];
  uint64_t strides[4];
  uint32_t box[4];
} mdm_tmap_spec;

enum { MDM_GEMM_PLAIN = 0, MDM_GEMM_CONV = 1, MDM_GEMM_CONV_WGRAD = 2 };
enum { MDM_ACT_NONE = 0, MDM_ACT_GELU = 1 };

typedef struct mdm_gemm_params {
  int32_t kind;
  int32_t M, N, K;
  int32_t block_n;
  int32_t nz1, nz2, nsplit;
  int32_t a_z1_off, b_z1_off, a_use_z, b_use_z;
  int32_t H, W, PW, PH, tiles_w, tiles_h, nimg;
  int32_t taps, flip, kblocks_c;
  int32_t num_kblocks;
  int32_t num_stages; /* filled by the launcher */
  float alpha;
  const float* alpha_dev;
  const float* bias;
  const float* residual;
  float* out_f32;
  void* out_f16;     /* __half* */
  void* out_act_f16; /* __half*: act(v); out_f16 then receives the pre-activation */
  int64_t ldc, c_z1_stride, c_z2_stride;
  int32_t act;
  int32_t atomic;
} mdm_gemm_params;

int mdm_gemm_raw(const mdm_tmap_spec* A, const mdm_tmap_spec* B, int a_mn, int b_mn,
                 const mdm_gemm_params* p, mdm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MDM_B200_H_ */

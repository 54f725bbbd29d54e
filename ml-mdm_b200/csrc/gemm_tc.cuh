// tcgen05 GEMM engine of the Matryoshka denoising path (sm_100a only).
//
// One kernel template covers every dense contraction of the nested U-Net forward and backward:
//   * Linear / 1x1 conv forward, dgrad, wgrad            (reference: nn.Linear / nn.Conv2d 1x1 calls,
//     ml_mdm/models/unet.py:206,219,260-271,605-626,763)
//   * 3x3 conv forward / dgrad as an implicit GEMM over NHWC pixel patches, the nine taps being nine
//     shifted TMA boxes with hardware zero fill for the padding (unet.py:199,210,515,525,632,751;
//     nested_unet.py:110,121), and 3x3 wgrad with pixels as the contraction dimension
//   * attention QK^T / PV and their backward contractions (unet.py:276-294)
//
// Operands are fp16 in HBM, staged by TMA into 128B-swizzled shared memory, multiplied by
// tcgen05.mma (M=128, N<=256, K=16 per instruction) with fp32 accumulators in TMEM, and drained by
// tcgen05.ld into a fused epilogue (alpha, bias, residual add, GELU, fp32/fp16 stores, split-K atomics).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>
#include <vector>

#include "mdm_b200.h"

namespace mdm {

enum GemmKind : int {
  GEMM_PLAIN = 0,       // A, B are (batched) matrices
  GEMM_CONV = 1,        // A = NHWC activation walked as pixel patches with tap shifts; B = packed weights
  GEMM_CONV_WGRAD = 2,  // A = dY patches, B = shifted X patches, contraction over pixels
};

enum GemmAct : int { ACT_NONE = 0, ACT_GELU = 1 };

using TmapSpec = mdm_tmap_spec;      // see include/mdm_b200.h
using GemmParams = mdm_gemm_params;  // see include/mdm_b200.h

// Host launcher. a_mn / b_mn select MN-major (transposed) operands. Returns cudaError_t as int.
int launch_gemm(const TmapSpec& A, const TmapSpec& B, int a_mn, int b_mn, const GemmParams& p,
                cudaStream_t stream);

int launch_gemm_persistent(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmO,
                           const CUtensorMap& tmBpart, const CUtensorMap& tmOp, int cluster, int a_mn, int b_mn,
                           GemmParams p, int m_tiles, int n_tiles, cudaStream_t stream);

// Counts kernel launches issued by this library (bench.py reports it as gpu_launches).
extern unsigned long long g_launch_count;
extern int g_sm_reserve;  // SMs the persistent kernels leave free (mdm_set_sm_reserve)
extern bool g_profile;
extern std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_profile_events;
extern std::vector<mdm_gemm_params> g_profile_params;
extern std::vector<int> g_profile_majors;

}  // namespace mdm

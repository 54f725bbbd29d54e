// tcgen05 / TMA / TMEM GEMM engine (see gemm_tc.cuh for the role of this kernel in the path).
#include "gemm_tc.cuh"

#include <stdio.h>

#include <utility>
#include <vector>

#include "ptx.cuh"

namespace mdm {

unsigned long long g_launch_count = 0;

// Optional per-launch timing of this kernel (bench.py's roofline leg): CUDA events on the launching
// stream around every launch while enabled.
bool g_profile = false;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_profile_events;

using namespace ptx;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // fp16 elements (K-major) or rows (MN-major) per pipeline stage
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int SLAB_BYTES = 64 * 64 * 2;  // one 64(k) x 64(mn) MN-major slab
constexpr int MAX_STAGES = 6;
constexpr int SMEM_BUDGET = 99 * 1024;  // two CTAs per SM: one runs its epilogue under the other's mainloop

__device__ __forceinline__ float gelu_erf(float v) {
  return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(128)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int nb_alloc = B_MN ? ((p.block_n + 63) / 64) * 64 : p.block_n;
  const int b_stage_bytes = nb_alloc * 128;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  const int nstages = p.num_stages;

  // ---- block coordinates
  int z = blockIdx.z;
  const int split = z % p.nsplit;
  z /= p.nsplit;
  const int z1 = z % p.nz1;
  const int z2 = z / p.nz1;
  const int per = (p.num_kblocks + p.nsplit - 1) / p.nsplit;
  const int kb_begin = split * per;
  const int kb_end = min(p.num_kblocks, kb_begin + per);
  const int nkb = kb_end - kb_begin;

  const int m_tile = blockIdx.x;
  const int m0 = m_tile * BLOCK_M;
  const int n0 = blockIdx.y * p.block_n;
  int img = 0, th = 0, tw = 0;
  if (p.kind == GEMM_CONV) {
    tw = m_tile % p.tiles_w;
    const int t = m_tile / p.tiles_w;
    th = t % p.tiles_h;
    img = t / p.tiles_h;
  }

  // ---- one-time setup
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(p.block_n)) tmem_cols <<= 1;
  if (warp == 1) tmem_alloc(&tmem_base_smem, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0 && lane == 0 && nkb > 0) {
    // =========================== TMA producer ===========================
    int stage = 0;
    uint32_t phase = 0;
    const int az1 = p.a_use_z ? z1 + p.a_z1_off : 0;
    const int az2 = p.a_use_z ? z2 : 0;
    const int bz1 = p.b_use_z ? z1 + p.b_z1_off : 0;
    const int bz2 = p.b_use_z ? z2 : 0;
    const int nslab_b = nb_alloc / 64;
    for (int i = 0; i < nkb; ++i) {
      const int kb = kb_begin + i;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* sA = smem + stage * stage_bytes;
      uint8_t* sB = sA + A_STAGE_BYTES;
      uint64_t* bar = &full_bar[stage];
      mbar_expect_tx(bar, static_cast<uint32_t>(stage_bytes));
      if (p.kind == GEMM_PLAIN) {
        if (!A_MN) {
          tma_load_4d(sA, &tmA, bar, kb * BLOCK_K, m0, az1, az2);
        } else {
          tma_load_4d(sA, &tmA, bar, m0, kb * BLOCK_K, az1, az2);
          tma_load_4d(sA + SLAB_BYTES, &tmA, bar, m0 + 64, kb * BLOCK_K, az1, az2);
        }
        if (!B_MN) {
          tma_load_4d(sB, &tmB, bar, kb * BLOCK_K, n0, bz1, bz2);
        } else {
          for (int s = 0; s < nslab_b; ++s)
            tma_load_4d(sB + s * SLAB_BYTES, &tmB, bar, n0 + 64 * s, kb * BLOCK_K, bz1, bz2);
        }
      } else if (p.kind == GEMM_CONV) {
        const int tap = kb / p.kblocks_c;
        const int cb = kb - tap * p.kblocks_c;
        const int kh = (p.taps == 9) ? tap / 3 : 1;
        const int kw = (p.taps == 9) ? tap % 3 : 1;
        tma_load_4d(sA, &tmA, bar, cb * BLOCK_K, tw * p.PW + kw - 1, th * p.PH + kh - 1, img);
        if (!B_MN) {
          tma_load_4d(sB, &tmB, bar, cb * BLOCK_K, n0, tap, 0);
        } else {
          const int wt = p.flip ? (p.taps - 1 - tap) : tap;
          for (int s = 0; s < nslab_b; ++s)
            tma_load_4d(sB + s * SLAB_BYTES, &tmB, bar, n0 + 64 * s, cb * BLOCK_K, wt, 0);
        }
      } else {  // GEMM_CONV_WGRAD: k block = one patch of 64 pixels
        const int ptw = kb % p.tiles_w;
        const int t = kb / p.tiles_w;
        const int pth = t % p.tiles_h;
        const int pimg = t / p.tiles_h;
        const int kh = (p.taps == 9) ? z1 / 3 : 1;
        const int kw = (p.taps == 9) ? z1 % 3 : 1;
        tma_load_4d(sA, &tmA, bar, m0, ptw * p.PW, pth * p.PH, pimg);
        tma_load_4d(sA + SLAB_BYTES, &tmA, bar, m0 + 64, ptw * p.PW, pth * p.PH, pimg);
        for (int s = 0; s < nslab_b; ++s)
          tma_load_4d(sB + s * SLAB_BYTES, &tmB, bar, n0 + 64 * s, ptw * p.PW + kw - 1,
                      pth * p.PH + kh - 1, pimg);
      }
      if (++stage == nstages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1 && lane == 0 && nkb > 0) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc_f16(BLOCK_M, p.block_n, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t a_base = smem_u32(smem + stage * stage_bytes);
      const uint32_t b_base = a_base + A_STAGE_BYTES;
#pragma unroll
      for (int k = 0; k < BLOCK_K / 16; ++k) {
        const uint64_t adesc = A_MN ? make_smem_desc_sw128(a_base + k * 2048, SLAB_BYTES, 1024)
                                    : make_smem_desc_sw128(a_base + k * 32, 16, 1024);
        const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_base + k * 2048, SLAB_BYTES, 1024)
                                    : make_smem_desc_sw128(b_base + k * 32, 16, 1024);
        umma_f16(tmem_base, adesc, bdesc, idesc, (i > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
      if (++stage == nstages) {
        stage = 0;
        phase ^= 1;
      }
    }
    umma_commit(&accum_bar);
  }
  __syncwarp();

  // =========================== epilogue (all four warps) ===========================
  if (nkb > 0) {
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
  }
  const int r = threadIdx.x;  // TMEM lane == tile row
  bool valid;
  long long row_off;
  if (p.kind == GEMM_CONV) {
    const int ph = r / p.PW;
    const int pw = r - ph * p.PW;
    const int h = th * p.PH + ph;
    const int w = tw * p.PW + pw;
    valid = (h < p.H) && (w < p.W);
    row_off = (static_cast<long long>(img * p.H + h) * p.W + w) * p.ldc;
  } else {
    const int row = m0 + r;
    valid = row < p.M;
    row_off = static_cast<long long>(row) * p.ldc + static_cast<long long>(z1) * p.c_z1_stride +
              static_cast<long long>(z2) * p.c_z2_stride;
  }
  float alpha = p.alpha;
  if (p.alpha_dev != nullptr) alpha *= __ldg(p.alpha_dev);
  const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

  for (int c = 0; c < p.block_n; c += 16) {
    float v[16];
    if (nkb > 0) {
      tmem_ld16(taddr_row + static_cast<uint32_t>(c), v);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
    }
    const int col0 = n0 + c;
    if (!valid || col0 >= p.N) continue;
    const long long off0 = row_off + col0;
    const bool full = (col0 + 16 <= p.N);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float x = v[j] * alpha;
      if (p.bias != nullptr && (full || col0 + j < p.N)) x += __ldg(p.bias + col0 + j);
      v[j] = x;
    }
    if (p.residual != nullptr) {
      if (full && ((off0 & 3) == 0)) {
        const float4* rp = reinterpret_cast<const float4*>(p.residual + off0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 t = __ldg(rp + q);
          v[4 * q + 0] += t.x;
          v[4 * q + 1] += t.y;
          v[4 * q + 2] += t.z;
          v[4 * q + 3] += t.w;
        }
      } else {
        for (int j = 0; j < 16; ++j)
          if (col0 + j < p.N) v[j] += __ldg(p.residual + off0 + j);
      }
    }
    if (p.atomic) {
      for (int j = 0; j < 16; ++j)
        if (col0 + j < p.N) atomicAdd(p.out_f32 + off0 + j, v[j]);
      continue;
    }
    if (p.out_f32 != nullptr) {
      if (full && ((off0 & 3) == 0)) {
        float4* op = reinterpret_cast<float4*>(p.out_f32 + off0);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          op[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      } else {
        for (int j = 0; j < 16; ++j)
          if (col0 + j < p.N) p.out_f32[off0 + j] = v[j];
      }
    }
    if (p.out_f16 != nullptr) {
      if (full && ((off0 & 7) == 0)) {
        __half2 h[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) h[q] = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
        uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out_f16) + off0);
        op[0] = *reinterpret_cast<uint4*>(&h[0]);
        op[1] = *reinterpret_cast<uint4*>(&h[4]);
      } else {
        for (int j = 0; j < 16; ++j)
          if (col0 + j < p.N) reinterpret_cast<__half*>(p.out_f16)[off0 + j] = __float2half_rn(v[j]);
      }
    }
    if (p.out_act_f16 != nullptr) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = (p.act == ACT_GELU) ? gelu_erf(v[j]) : v[j];
      if (full && ((off0 & 7) == 0)) {
        __half2 h[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) h[q] = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
        uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out_act_f16) + off0);
        op[0] = *reinterpret_cast<uint4*>(&h[0]);
        op[1] = *reinterpret_cast<uint4*>(&h[4]);
      } else {
        for (int j = 0; j < 16; ++j)
          if (col0 + j < p.N) reinterpret_cast<__half*>(p.out_act_f16)[off0 + j] = __float2half_rn(v[j]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_tmap(CUtensorMap* out, const TmapSpec& s) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return -1;
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) {
    gdim[i] = s.dims[i];
    box[i] = s.box[i];
  }
  for (int i = 0; i < 3; ++i) gstr[i] = s.strides[i + 1] * 2;  // bytes (fp16)
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(s.ptr), gdim, gstr, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr,
            "[mdm_b200] cuTensorMapEncodeTiled failed (%d): ptr=%p dims=(%llu,%llu,%llu,%llu) "
            "strides=(%llu,%llu,%llu,%llu) box=(%u,%u,%u,%u)\n",
            static_cast<int>(r), s.ptr, (unsigned long long)s.dims[0], (unsigned long long)s.dims[1],
            (unsigned long long)s.dims[2], (unsigned long long)s.dims[3],
            (unsigned long long)s.strides[0], (unsigned long long)s.strides[1],
            (unsigned long long)s.strides[2], (unsigned long long)s.strides[3], s.box[0], s.box[1],
            s.box[2], s.box[3]);
    return -2;
  }
  return 0;
}

template <bool A_MN, bool B_MN>
int launch_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, dim3 grid,
                size_t smem, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<A_MN, B_MN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_profile) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, stream);
  }
  gemm_tc_kernel<A_MN, B_MN><<<grid, 128, smem, stream>>>(tmA, tmB, p);
  if (g_profile) {
    cudaEventRecord(e1, stream);
    g_profile_events.emplace_back(e0, e1);
  }
  ++g_launch_count;
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

int launch_gemm(const TmapSpec& A, const TmapSpec& B, int a_mn, int b_mn, const GemmParams& pin,
                cudaStream_t stream) {
  GemmParams p = pin;
  if (p.block_n < 16 || p.block_n > 256 || (p.block_n % 16) != 0) return -10;
  if (p.nz1 < 1) p.nz1 = 1;
  if (p.nz2 < 1) p.nz2 = 1;
  if (p.nsplit < 1) p.nsplit = 1;
  if (p.num_kblocks < 1) return -11;
  if (p.nsplit > p.num_kblocks) p.nsplit = p.num_kblocks;
  // every split must own at least one k block
  while (p.nsplit > 1 && (p.nsplit - 1) * ((p.num_kblocks + p.nsplit - 1) / p.nsplit) >= p.num_kblocks)
    --p.nsplit;
  if (p.atomic == 0 && p.nsplit != 1) return -12;

  alignas(64) CUtensorMap tmA, tmB;
  if (encode_tmap(&tmA, A) != 0) return -20;
  if (encode_tmap(&tmB, B) != 0) return -21;

  const int nb_alloc = b_mn ? ((p.block_n + 63) / 64) * 64 : p.block_n;
  const int stage_bytes = A_STAGE_BYTES + nb_alloc * 128;
  int stages = (SMEM_BUDGET - 1024) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  const int per = (p.num_kblocks + p.nsplit - 1) / p.nsplit;
  if (stages > per) stages = per;
  if (stages < 1) return -13;
  p.num_stages = stages;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + 1024;

  int m_tiles;
  if (p.kind == GEMM_CONV) {
    m_tiles = p.nimg * p.tiles_h * p.tiles_w;
  } else {
    m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  }
  const int n_tiles = (p.N + p.block_n - 1) / p.block_n;
  dim3 grid(m_tiles, n_tiles, p.nz1 * p.nz2 * p.nsplit);
  if (grid.y > 65535 || grid.z > 65535) return -14;

  if (!a_mn && !b_mn) return launch_impl<false, false>(tmA, tmB, p, grid, smem, stream);
  if (!a_mn && b_mn) return launch_impl<false, true>(tmA, tmB, p, grid, smem, stream);
  if (a_mn && b_mn) return launch_impl<true, true>(tmA, tmB, p, grid, smem, stream);
  return launch_impl<true, false>(tmA, tmB, p, grid, smem, stream);
}

}  // namespace mdm

// tcgen05 / TMA / TMEM GEMM engine (see gemm_tc.cuh for the role of this kernel in the path).
#include "gemm_tc.cuh"

#include <stdio.h>
#include <stdlib.h>

#include <utility>
#include <vector>

#include "gemm_epi.cuh"
#include "ptx.cuh"

namespace mdm {

unsigned long long g_launch_count = 0;

// Optional per-launch timing of this kernel (bench.py's roofline leg): CUDA events on the launching
// stream around every launch while enabled.
bool g_profile = false;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_profile_events;
std::vector<GemmParams> g_profile_params;  // parallel to g_profile_events
std::vector<int> g_profile_majors;

using namespace ptx;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // fp16 elements (K-major) or rows (MN-major) per pipeline stage
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int SLAB_BYTES = 64 * 64 * 2;  // one 64(k) x 64(mn) MN-major slab
constexpr int MAX_STAGES = 6;
constexpr int SMEM_BUDGET = 99 * 1024;  // two CTAs per SM: one runs its epilogue under the other's mainloop

constexpr int NUM_THREADS = 256;  // warps 0-3: TMA / MMA / epilogue, warps 4-7: epilogue only

// PAIR is a template parameter because a kernel may not mix cta_group::1 and cta_group::2 tcgen05 instructions (a kernel
// holding any cta_group::2 instruction only launches with an even cluster size).
template <bool A_MN, bool B_MN, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO32, const __grid_constant__ CUtensorMap tmO16,
               const __grid_constant__ CUtensorMap tmOact, const __grid_constant__ CUtensorMap tmBpart,
               const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_bias[256];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // CTA pair (cta_group::2): this CTA stages its own 128 rows of A and its half of the B tile
  constexpr bool pair = PAIR;
  const int nb_local = pair ? p.block_n / 2 : p.block_n;
  const int nb_alloc = B_MN ? ((nb_local + 63) / 64) * 64 : nb_local;
  // weight gradients of narrow layers: kf x 64 pixel rows per stage, and a single A slab when M <= 64
  const int kf = (p.kind == GEMM_CONV_WGRAD && p.kfactor > 1) ? p.kfactor : 1;
  const int slab_bytes = SLAB_BYTES * kf;
  const int a_slabs = (kf > 1 && p.M <= 64) ? 1 : 2;
  const int a_stage_bytes = kf > 1 ? a_slabs * slab_bytes : A_STAGE_BYTES;
  const int b_stage_bytes = nb_alloc * 128 * kf;
  const int stage_bytes = a_stage_bytes + b_stage_bytes;
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

  // Cluster of cs CTAs along the M tiles: same B tile for all of them, each fetches 1/cs of it and multicasts.
  // (TMA moves ~one <= 128-byte row per 3 clocks per SM whatever the row holds; a 128 x 256 tile needs 128 A + 256 B
  // rows per k block against 512 clocks of MMA, so the loads, not the tensor pipe, set the pace without this.)
  const int cs = p.cluster > 1 ? p.cluster : 1;  // (multicast clusters and CTA pairs are mutually exclusive)
  const uint32_t crank = (cs > 1 || pair) ? cluster_ctarank() : 0u;
  const uint16_t cmask = static_cast<uint16_t>((1u << cs) - 1u);
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
      mbar_init(&empty_bar[s], static_cast<uint32_t>(cs));  // one arrival per consumer CTA of the cluster
    }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(p.block_n)) tmem_cols <<= 1;
  if (warp == 1) {
    if constexpr (pair) tmem_alloc_2sm(&tmem_base_smem, tmem_cols);
    else tmem_alloc(&tmem_base_smem, tmem_cols);
  }
  for (int i = threadIdx.x; i < p.block_n; i += blockDim.x)
    s_bias[i] = (p.bias != nullptr && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  if (cs > 1 || pair) cluster_sync_all();  // peers' barriers exist before anything is multicast to them
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
      uint8_t* sB = sA + a_stage_bytes;
      uint64_t* bar = &full_bar[stage];
      if constexpr (pair) {
        // Both CTAs' boxes complete on the LEADER's barrier (armed by the leader alone with both CTAs' bytes): its MMA
        // reads the two shared memories. This CTA's slot is freed by the leader's multicast commit on empty_bar.
        const uint32_t lbar = mapa_u32(smem_u32(bar), 0);
        if (crank == 0) mbar_expect_tx(bar, static_cast<uint32_t>(2 * stage_bytes));
        const int nh = nb_local;               // B columns of this CTA
        const int nb0 = n0 + static_cast<int>(crank) * nh;
        const int nslab = nb_alloc / 64;
        if (p.kind == GEMM_PLAIN) {
          if (!A_MN) {
            tma_load_4d_2sm(sA, &tmA, lbar, kb * BLOCK_K, m0, az1, az2);
          } else {
            tma_load_4d_2sm(sA, &tmA, lbar, m0, kb * BLOCK_K, az1, az2);
            tma_load_4d_2sm(sA + SLAB_BYTES, &tmA, lbar, m0 + 64, kb * BLOCK_K, az1, az2);
          }
          if (!B_MN) {
            tma_load_4d_2sm(sB, &tmBpart, lbar, kb * BLOCK_K, nb0, bz1, bz2);
          } else {
            for (int s = 0; s < nslab; ++s)
              tma_load_4d_2sm(sB + s * SLAB_BYTES, &tmB, lbar, nb0 + 64 * s, kb * BLOCK_K, bz1, bz2);
          }
        } else if (p.kind == GEMM_CONV) {
          const int tap = kb / p.kblocks_c;
          const int cb = kb - tap * p.kblocks_c;
          const int kh = (p.taps == 9) ? tap / 3 : 1;
          const int kw = (p.taps == 9) ? tap % 3 : 1;
          tma_load_4d_2sm(sA, &tmA, lbar, cb * BLOCK_K, tw * p.PW + kw - 1, th * p.PH + kh - 1, img);
          const int wt = p.flip ? (p.taps - 1 - tap) : tap;
          if (!B_MN) {
            tma_load_4d_2sm(sB, &tmBpart, lbar, cb * BLOCK_K, nb0, tap, 0);
          } else {
            for (int s = 0; s < nslab; ++s)
              tma_load_4d_2sm(sB + s * SLAB_BYTES, &tmB, lbar, nb0 + 64 * s, cb * BLOCK_K, wt, 0);
          }
        } else {  // GEMM_CONV_WGRAD
          const int ptw = kb % p.tiles_w;
          const int t = kb / p.tiles_w;
          const int pth = t % p.tiles_h;
          const int pimg = t / p.tiles_h;
          const int kh = (p.taps == 9) ? z1 / 3 : 1;
          const int kw = (p.taps == 9) ? z1 % 3 : 1;
          tma_load_4d_2sm(sA, &tmA, lbar, m0, ptw * p.PW, pth * p.PH, pimg);
          if (a_slabs == 2) tma_load_4d_2sm(sA + slab_bytes, &tmA, lbar, m0 + 64, ptw * p.PW, pth * p.PH, pimg);
          for (int s = 0; s < nslab; ++s)
            tma_load_4d_2sm(sB + s * slab_bytes, &tmB, lbar, nb0 + 64 * s, ptw * p.PW + kw - 1, pth * p.PH + kh - 1, pimg);
        }
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1;
        }
        continue;
      }
      mbar_expect_tx(bar, static_cast<uint32_t>(stage_bytes));
      if (p.kind == GEMM_PLAIN) {
        if (!A_MN) {
          tma_load_4d(sA, &tmA, bar, kb * BLOCK_K, m0, az1, az2);
        } else {
          tma_load_4d(sA, &tmA, bar, m0, kb * BLOCK_K, az1, az2);
          tma_load_4d(sA + SLAB_BYTES, &tmA, bar, m0 + 64, kb * BLOCK_K, az1, az2);
        }
        if (cs > 1) {
          if (!B_MN) {
            const int rows = p.block_n / cs;
            tma_load_4d_mc(sB + crank * rows * 128, &tmBpart, bar, kb * BLOCK_K, n0 + crank * rows, bz1, bz2, cmask);
          } else {
            for (int s = crank; s < nslab_b; s += cs)
              tma_load_4d_mc(sB + s * SLAB_BYTES, &tmB, bar, n0 + 64 * s, kb * BLOCK_K, bz1, bz2, cmask);
          }
        } else if (!B_MN) {
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
        const int wt = p.flip ? (p.taps - 1 - tap) : tap;
        if (cs > 1) {
          if (!B_MN) {
            const int rows = p.block_n / cs;
            tma_load_4d_mc(sB + crank * rows * 128, &tmBpart, bar, cb * BLOCK_K, n0 + crank * rows, tap, 0, cmask);
          } else {
            for (int s = crank; s < nslab_b; s += cs)
              tma_load_4d_mc(sB + s * SLAB_BYTES, &tmB, bar, n0 + 64 * s, cb * BLOCK_K, wt, 0, cmask);
          }
        } else if (!B_MN) {
          tma_load_4d(sB, &tmB, bar, cb * BLOCK_K, n0, tap, 0);
        } else {
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
        if (a_slabs == 2) tma_load_4d(sA + slab_bytes, &tmA, bar, m0 + 64, ptw * p.PW, pth * p.PH, pimg);
        if (cs > 1) {
          for (int s = crank; s < nslab_b; s += cs)
            tma_load_4d_mc(sB + s * slab_bytes, &tmB, bar, n0 + 64 * s, ptw * p.PW + kw - 1, pth * p.PH + kh - 1, pimg,
                           cmask);
        } else {
          for (int s = 0; s < nslab_b; ++s)
            tma_load_4d(sB + s * slab_bytes, &tmB, bar, n0 + 64 * s, ptw * p.PW + kw - 1,
                        pth * p.PH + kh - 1, pimg);
        }
      }
      if (++stage == nstages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1 && lane == 0 && nkb > 0 && !(pair && crank != 0)) {
    // =========================== MMA issuer (the leader CTA of a pair issues for both) ===========================
    const uint32_t idesc = make_idesc_f16(pair ? 2 * BLOCK_M : BLOCK_M, p.block_n, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t a_base = smem_u32(smem + stage * stage_bytes);
      const uint32_t b_base = a_base + a_stage_bytes;
      if (kf == 1) {
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {
          const uint64_t adesc = A_MN ? make_smem_desc_sw128(a_base + k * 2048, SLAB_BYTES, 1024)
                                      : make_smem_desc_sw128(a_base + k * 32, 16, 1024);
          const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_base + k * 2048, SLAB_BYTES, 1024)
                                      : make_smem_desc_sw128(b_base + k * 32, 16, 1024);
          if constexpr (pair) umma_f16_2sm(tmem_base, adesc, bdesc, idesc, (i > 0 || k > 0) ? 1u : 0u);
          else umma_f16(tmem_base, adesc, bdesc, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
      } else {
        // tall MN-major slabs (kf * 64 pixel rows x 64 columns). With one A slab (M <= 64) the second 64-column
        // block of A re-reads the first (leading-dimension offset 0): accumulator rows 64..127 are never stored.
        const uint32_t a_lbo = a_slabs == 2 ? static_cast<uint32_t>(slab_bytes) : 0u;
        for (int k = 0; k < 4 * kf; ++k) {
          const uint64_t adesc = make_smem_desc_sw128(a_base + k * 2048, a_lbo, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(b_base + k * 2048, static_cast<uint32_t>(slab_bytes), 1024);
          if constexpr (pair) umma_f16_2sm(tmem_base, adesc, bdesc, idesc, (i > 0 || k > 0) ? 1u : 0u);
          else umma_f16(tmem_base, adesc, bdesc, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
      }
      // frees the smem slot once these MMAs retire -- in every CTA of the cluster, whose producers all write into it
      if constexpr (pair) umma_commit_2sm(&empty_bar[stage]);
      else if (cs > 1) umma_commit_mc(&empty_bar[stage], cmask);
      else umma_commit(&empty_bar[stage]);
      if (++stage == nstages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if constexpr (pair) umma_commit_2sm(&accum_bar);
    else umma_commit(&accum_bar);
  }
  __syncwarp();

  // =========================== epilogue (all four warps) ===========================
  if (nkb > 0) {
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
  }
  const int r = threadIdx.x & 127;  // TMEM lane == tile row; warps w and w+4 share a lane quarter
  const int half_id = threadIdx.x >> 7;  // which half of the columns this thread drains
  bool valid;
  long long row_off;
  if (p.kind == GEMM_CONV) {
    const int ph = r / p.PW;
    const int pw = r - ph * p.PW;
    const int h = th * p.PH + ph;
    const int w = tw * p.PW + pw;
    valid = (h < p.H) && (w < p.W) && (img < p.nimg);  // (the grid is rounded up to whole clusters)
    row_off = (static_cast<long long>(img * p.H + h) * p.W + w) * p.ldc;
  } else {
    const int row = m0 + r;
    valid = row < p.M;
    row_off = static_cast<long long>(row) * p.ldc + static_cast<long long>(z1) * p.c_z1_stride +
              static_cast<long long>(z2) * p.c_z2_stride;
  }
  float alpha = p.alpha;
  if (p.alpha_dev != nullptr) alpha *= __ldg(p.alpha_dev);
  const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);

  if (p.epi_tma) {
    // ---- staged epilogue: registers -> 128B-swizzled smem tiles (the freed pipeline stages) -> TMA
    // stores. Rows/columns outside the tensor are clipped by the TMA unit.
    uint8_t* st32 = smem;  // 2 x [128][32] fp32
    uint8_t* st16 = smem + (p.out_f32 != nullptr ? 32768 : 0);
    uint8_t* stact = st16 + (p.out_f16 != nullptr ? 16384 : 0);
    const uint32_t swz = static_cast<uint32_t>(r & 7);
    const int ngroups = (p.block_n + 63) / 64;
    int oc1, oc2, oc3;  // output coordinates of the tile origin beyond the column
    if (p.kind == GEMM_CONV) {
      oc1 = tw * p.PW;
      oc2 = th * p.PH;
      oc3 = img;
    } else {
      oc1 = m0;
      oc2 = z1;
      oc3 = z2;
    }
    for (int g = 0; g < ngroups; ++g) {
      if (g > 0) {
        if (threadIdx.x == 0) bulk_wait_read();
        __syncthreads();
      }
#pragma unroll 1
      for (int ci = 0; ci < 2; ++ci) {
        const int cc = half_id * 2 + ci;
        const int c = g * 64 + cc * 16;
        if (c >= p.block_n) break;  // warp-uniform
        float v[16];
        tmem_ld16(taddr_row + static_cast<uint32_t>(c), v);
        const int col0 = n0 + c;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = v[j] * alpha + s_bias[c + j];
        if (p.residual != nullptr && valid && col0 < p.N) {
          const long long off0 = row_off + col0;
          if (col0 + 16 <= p.N && ((off0 & 3) == 0)) {
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
        if (p.gelu_grad_src != nullptr && valid && col0 < p.N) {
          const __half* gp = reinterpret_cast<const __half*>(p.gelu_grad_src) + row_off + col0;
          if (col0 + 16 <= p.N && (((row_off + col0) & 7) == 0)) {
            const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(gp));
            const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(gp) + 1);
            const __half2* h0 = reinterpret_cast<const __half2*>(&u0);
            const __half2* h1 = reinterpret_cast<const __half2*>(&u1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 a = __half22float2(h0[q]), b2 = __half22float2(h1[q]);
              v[2 * q] *= gelu_grad(a.x);
              v[2 * q + 1] *= gelu_grad(a.y);
              v[8 + 2 * q] *= gelu_grad(b2.x);
              v[8 + 2 * q + 1] *= gelu_grad(b2.y);
            }
          } else {
            for (int j = 0; j < 16; ++j)
              if (col0 + j < p.N) v[j] *= gelu_grad(__half2float(gp[j]));
          }
        }
        if (p.out_f32 != nullptr) {
          uint8_t* base = st32 + (cc >> 1) * 16384 + r * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t chunk = static_cast<uint32_t>((cc & 1) * 4 + q) ^ swz;
            *reinterpret_cast<float4*>(base + chunk * 16) =
                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
        if (p.out_f16 != nullptr) {
          uint8_t* base = st16 + r * 128;
          __half2 h[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) h[q] = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
          *reinterpret_cast<uint4*>(base + ((static_cast<uint32_t>(cc * 2) ^ swz) * 16)) = *reinterpret_cast<uint4*>(&h[0]);
          *reinterpret_cast<uint4*>(base + ((static_cast<uint32_t>(cc * 2 + 1) ^ swz) * 16)) = *reinterpret_cast<uint4*>(&h[4]);
        }
        if (p.out_act_f16 != nullptr) {
          uint8_t* base = stact + r * 128;
          __half2 h[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float a0 = (p.act == ACT_GELU) ? gelu_erf(v[2 * q]) : v[2 * q];
            const float a1 = (p.act == ACT_GELU) ? gelu_erf(v[2 * q + 1]) : v[2 * q + 1];
            h[q] = __floats2half2_rn(a0, a1);
          }
          *reinterpret_cast<uint4*>(base + ((static_cast<uint32_t>(cc * 2) ^ swz) * 16)) = *reinterpret_cast<uint4*>(&h[0]);
          *reinterpret_cast<uint4*>(base + ((static_cast<uint32_t>(cc * 2 + 1) ^ swz) * 16)) = *reinterpret_cast<uint4*>(&h[4]);
        }
      }
      fence_proxy_async();
      __syncthreads();
      if (threadIdx.x == 0) {
        const int colg = n0 + g * 64;
        if (colg < p.N) {
          if (p.out_f32 != nullptr) {
            if (p.atomic) {  // split-K / accumulate: reduce-add in the TMA unit
              tma_reduce_add_4d(&tmO32, st32, colg, oc1, oc2, oc3);
              if (colg + 32 < p.N && g * 64 + 32 < p.block_n) tma_reduce_add_4d(&tmO32, st32 + 16384, colg + 32, oc1, oc2, oc3);
            } else {
              tma_store_4d(&tmO32, st32, colg, oc1, oc2, oc3);
              if (colg + 32 < p.N && g * 64 + 32 < p.block_n) tma_store_4d(&tmO32, st32 + 16384, colg + 32, oc1, oc2, oc3);
            }
          }
          if (p.out_f16 != nullptr) tma_store_4d(&tmO16, st16, colg, oc1, oc2, oc3);
          if (p.out_act_f16 != nullptr) tma_store_4d(&tmOact, stact, colg, oc1, oc2, oc3);
        }
        bulk_commit();
      }
    }
    if (threadIdx.x == 0) bulk_wait_all();
    tc_fence_before();
    __syncthreads();
    if (cs > 1 || pair) cluster_sync_all();  // no peer may still arrive on this CTA's barriers after it is gone
    if (warp == 1) {
      if constexpr (pair) tmem_dealloc_2sm(tmem_base, tmem_cols);
      else tmem_dealloc(tmem_base, tmem_cols);
    }
    return;
  }

  for (int c = 0; c < p.block_n && half_id == 0; c += 16) {
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
      v[j] = x + s_bias[c + j];
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
    if (p.gelu_grad_src != nullptr) {
      const __half* gp = reinterpret_cast<const __half*>(p.gelu_grad_src) + off0;
      for (int j = 0; j < 16; ++j)
        if (col0 + j < p.N) v[j] *= gelu_grad(__half2float(gp[j]));
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
  if (cs > 1 || pair) cluster_sync_all();  // no peer may still arrive on this CTA's barriers after it is gone
  if (warp == 1) {
    if constexpr (pair) tmem_dealloc_2sm(tmem_base, tmem_cols);
    else tmem_dealloc(tmem_base, tmem_cols);
  }
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

int encode_tmap(CUtensorMap* out, const TmapSpec& s, bool f32 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return -1;
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) {
    gdim[i] = s.dims[i];
    box[i] = s.box[i];
  }
  for (int i = 0; i < 3; ++i) gstr[i] = s.strides[i + 1] * (f32 ? 4 : 2);  // bytes
  CUresult r = fn(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(s.ptr), gdim, gstr, box,
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

template <bool A_MN, bool B_MN, bool PAIR>
int launch_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmO, const CUtensorMap& tmBpart,
                const GemmParams& p, dim3 grid, size_t smem, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<A_MN, B_MN, PAIR>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_profile) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, stream);
  }
  if (p.cluster > 1 || p.pair) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.pair ? 2u : static_cast<unsigned>(p.cluster);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<A_MN, B_MN, PAIR>, tmA, tmB, tmO[0], tmO[1], tmO[2], tmBpart, p);
    if (le != cudaSuccess) return static_cast<int>(le);
  } else {
    gemm_tc_kernel<A_MN, B_MN, PAIR><<<grid, NUM_THREADS, smem, stream>>>(tmA, tmB, tmO[0], tmO[1], tmO[2], tmBpart, p);
  }
  if (g_profile) {
    cudaEventRecord(e1, stream);
    g_profile_events.emplace_back(e0, e1);
    g_profile_params.push_back(p);
    g_profile_majors.push_back((A_MN ? 2 : 0) | (B_MN ? 1 : 0));
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

  if (p.kind != GEMM_CONV_WGRAD || p.kfactor < 1) p.kfactor = 1;
  const int kf = p.kfactor;
  if (kf > 1 && (!a_mn || !b_mn || p.PW * p.PH != 64 * kf || kf > 4)) return -15;
  int m_tiles;
  if (p.kind == GEMM_CONV) {
    m_tiles = p.nimg * p.tiles_h * p.tiles_w;
  } else {
    m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  }
  const int n_tiles = (p.N + p.block_n - 1) / p.block_n;

  // ---- which kernel form. Persistent warp-specialised variant (gemm_persistent.cu) for the non-split, TMA-store
  // launches. Measured (B200): it wins where the epilogue matters (K <= ~3000: 16384x3072x768 runs at 1048 vs 824
  // TFLOP/s) and loses on long-K tiles, where two co-resident CTAs hide latency better (8192^3: 1196 vs 1353), and
  // when the tile count leaves a mostly empty last round.
  static const bool persistent = getenv("MDM_GEMM_NO_PERSISTENT") == nullptr;
  static const int persist_min_n = getenv("MDM_PERSIST_MIN_N") ? atoi(getenv("MDM_PERSIST_MIN_N")) : 96;
  const long long ntiles = static_cast<long long>(m_tiles) * n_tiles * p.nz1 * p.nz2;
  const bool rounds_ok = ntiles <= 148 || ntiles >= 296;
  // With CTA pairs in the one-tile form the crossover moved down for the linears: K >= 2304 is faster there (fc1 data
  // gradient 16384x768x3072 2.67 -> 2.17 ms per step, qkv data gradient 2.07 -> 1.78, fc2 forward 3.02 -> 2.64), K = 1536
  // with a GELU' epilogue is not (1.76 -> 2.26), and the 36-k-block 3x3 convs stay (2.21 vs 2.34):
  // profiles/r02_persist_kb_sweep.txt.
  static const int persist_max_kb = getenv("MDM_PERSIST_MAX_KBLOCKS") ? atoi(getenv("MDM_PERSIST_MAX_KBLOCKS")) : 0;
  const int max_kb = persist_max_kb > 0 ? persist_max_kb : (p.kind == GEMM_PLAIN ? 35 : 48);
  const bool persist_shape = persistent && p.block_n >= persist_min_n && p.nsplit == 1 && p.kind != GEMM_CONV_WGRAD &&
                             p.num_kblocks <= max_kb && rounds_ok;  // (and a TMA-store epilogue, checked below)

  // ---- CTA pairs (cta_group::2) for the MMA-bound launches of the one-tile form: wide tiles, long contraction.
  // Each CTA stages 128 rows of A and HALF of the B tile; the leader issues M = 256 instructions for both.
  // Measured (profiles/r02_pair_sweep.txt, one-tile form, pairs off -> on): 8192^3 1287 -> 1399 TFLOP/s, 8192x768x6912
  // 1016 -> 1199, 3x3 conv 768->768 @16x16 968 -> 1117, 512->512 @32x32 1135 -> 1265 (cuBLAS bf16 on the same box:
  // 1340-1556). MDM_GEMM_PAIR=0 turns them off.
  static const int pair_env = getenv("MDM_GEMM_PAIR") ? atoi(getenv("MDM_GEMM_PAIR")) : 1;
  p.pair = (pair_env != 0 && m_tiles >= 2 && p.block_n >= 128 && (p.block_n % 32) == 0 && !(kf > 1 && p.M <= 64)) ? 1 : 0;

  auto stage_bytes_of = [&](bool pr) {
    const int nl = pr ? p.block_n / 2 : p.block_n;
    const int na = b_mn ? ((nl + 63) / 64) * 64 : nl;
    return kf > 1 ? ((p.M <= 64 ? 1 : 2) + na / 64) * SLAB_BYTES * kf : A_STAGE_BYTES + na * 128;
  };
  const int nb_alloc = b_mn ? ((p.block_n + 63) / 64) * 64 : p.block_n;  // unpaired
  int stage_bytes = stage_bytes_of(p.pair != 0);
  static const int budget_kb = getenv("MDM_SMEM_BUDGET_KB") ? atoi(getenv("MDM_SMEM_BUDGET_KB")) : 0;  // dev knob
  // Narrow tiles (N <= 64: the 32/64-channel levels of the 256- and 1024-px nests) do so little work per tile that
  // the fixed per-tile latencies dominate; three co-resident CTAs of the one-tile-per-CTA form hide them better than
  // the persistent form or two fat CTAs. Measured (profiles/r02_narrow_tile_knobs.txt): 3x3 convs with N <= 64 of a
  // cc12m_256x256 step 11.0 ms persistent -> 9.7 ms one-tile form -> 7.0 ms with a 64 KB budget (8.4 ms at 44 KB);
  // weight gradients prefer the deep pipeline and keep the full budget.
  static const int narrow_kb = getenv("MDM_SMEM_NARROW_KB") ? atoi(getenv("MDM_SMEM_NARROW_KB")) : 64;
  int budget = budget_kb > 0 ? budget_kb * 1024 : SMEM_BUDGET;
  if (narrow_kb > 0 && p.block_n <= 64 && p.kind != GEMM_CONV_WGRAD) budget = narrow_kb * 1024;
  if (kf > 1) budget = 200 * 1024;  // tall stages: one CTA per SM, three 64 KB stages
  int stages = (budget - 1024) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  const int per = (p.num_kblocks + p.nsplit - 1) / p.nsplit;
  if (stages > per) stages = per;
  if (stages < 1) return -13;
  p.num_stages = stages;
  size_t smem = static_cast<size_t>(stages) * stage_bytes + 1024;

  dim3 grid(m_tiles, n_tiles, p.nz1 * p.nz2 * p.nsplit);
  if (grid.y > 65535 || grid.z > 65535) return -14;

  // ---- staged (TMA-store) epilogue when the output geometry allows it
  alignas(64) CUtensorMap tmO[3];
  alignas(64) CUtensorMap tmOp = tmA;  // epilogue operand (residual / GELU' source) of the persistent form
  tmO[0] = tmO[1] = tmO[2] = tmA;
  p.epi_tma = 0;
  p.epi_op = 0;
  {
    auto ok = [&](const void* ptr, int esz) {
      if (ptr == nullptr) return true;
      if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return false;
      if ((p.ldc * esz) % 16 != 0) return false;
      if (p.kind != GEMM_CONV && p.nz1 > 1 && (p.c_z1_stride * esz) % 16 != 0) return false;
      if (p.kind != GEMM_CONV && p.nz2 > 1 && (p.c_z2_stride * esz) % 16 != 0) return false;
      return true;
    };
    const bool any_out = p.out_f32 != nullptr || p.out_f16 != nullptr || p.out_act_f16 != nullptr;
    const bool shape_ok = (p.block_n % 64 == 0) || (n_tiles == 1);
    const bool atomic_ok = !p.atomic || (p.out_f16 == nullptr && p.out_act_f16 == nullptr && p.gelu_grad_src == nullptr);
    if (atomic_ok && any_out && shape_ok && ok(p.out_f32, 4) && ok(p.out_f16, 2) && ok(p.out_act_f16, 2) &&
        getenv("MDM_NO_TMA_EPILOGUE") == nullptr) {
      auto mk = [&](void* ptr, int esz, uint32_t box0) {
        TmapSpec o;
        o.ptr = ptr;
        if (p.kind == GEMM_CONV) {
          o.dims[0] = p.N; o.dims[1] = p.W; o.dims[2] = p.H; o.dims[3] = p.nimg;
          o.strides[0] = 1; o.strides[1] = p.ldc; o.strides[2] = static_cast<uint64_t>(p.W) * p.ldc;
          o.strides[3] = static_cast<uint64_t>(p.H) * p.W * p.ldc;
          o.box[0] = box0; o.box[1] = p.PW; o.box[2] = p.PH; o.box[3] = 1;
        } else {
          const uint64_t dflt = static_cast<uint64_t>(p.ldc) * p.M;
          o.dims[0] = p.N; o.dims[1] = p.M; o.dims[2] = p.nz1; o.dims[3] = p.nz2;
          o.strides[0] = 1; o.strides[1] = p.ldc;
          o.strides[2] = p.nz1 > 1 ? static_cast<uint64_t>(p.c_z1_stride) : dflt;
          o.strides[3] = p.nz2 > 1 ? static_cast<uint64_t>(p.c_z2_stride) : dflt;
          o.box[0] = box0; o.box[1] = BLOCK_M; o.box[2] = 1; o.box[3] = 1;
        }
        (void)esz;
        return o;
      };
      bool good = true;
      if (p.out_f32 != nullptr) good = good && encode_tmap(&tmO[0], mk(p.out_f32, 4, 32), true) == 0;
      if (p.out_f16 != nullptr) good = good && encode_tmap(&tmO[1], mk(p.out_f16, 2, 64), false) == 0;
      if (p.out_act_f16 != nullptr) good = good && encode_tmap(&tmO[2], mk(p.out_act_f16, 2, 64), false) == 0;
      if (good) {
        // the persistent form can take its epilogue operand by TMA into the staging tile (gemm_persistent.cu)
        static const bool op_tma = getenv("MDM_EPI_OP_TMA") == nullptr || atoi(getenv("MDM_EPI_OP_TMA")) != 0;
        if (op_tma && persist_shape && !p.atomic) {
          if (p.residual != nullptr && p.out_f32 != nullptr && p.gelu_grad_src == nullptr && ok(p.residual, 4)) {
            if (encode_tmap(&tmOp, mk(const_cast<float*>(p.residual), 4, 32), true) == 0) p.epi_op = 1;
          } else if (p.gelu_grad_src != nullptr && p.out_f16 != nullptr && p.residual == nullptr && ok(p.gelu_grad_src, 2)) {
            if (encode_tmap(&tmOp, mk(const_cast<void*>(p.gelu_grad_src), 2, 64), false) == 0) p.epi_op = 2;
          }
        }
        p.epi_tma = 1;
        const size_t need = (p.out_f32 ? 32768 : 0) + (p.out_f16 ? 16384 : 0) + (p.out_act_f16 ? 16384 : 0) + 1024;
        if (smem < need) smem = need;
      }
    }
  }

  // ---- B multicast over a cluster of CTAs that share the n tile (adjacent m tiles)
  // Off by default: measured (profiles/r02_multicast.txt) it changes nothing or costs 3-5 % -- the multicast removes
  // L2 -> SM requests, but every SM still writes the whole B tile into its shared memory, and shared-memory traffic
  // (TMA writes + tcgen05 operand reads, ~190 B/clk against 128 B/clk for a 128 x 256 tile) is what holds the
  // one-CTA MMA at ~1.1 PFLOP/s; only cta_group::2 (half of B per SM) removes that.
  static const int cluster_env = getenv("MDM_GEMM_CLUSTER") ? atoi(getenv("MDM_GEMM_CLUSTER")) : 1;
  alignas(64) CUtensorMap tmBpart = tmB;
  int cs = cluster_env;
  {
    if (cs != 2 && cs != 4) cs = 1;
    const int nslab = nb_alloc / 64;
    while (cs > 1 && !(m_tiles >= 2 * cs && p.block_n >= 128 &&
                       (b_mn ? (nslab % cs == 0) : ((p.block_n / cs) % 8 == 0))))
      cs >>= 1;
    if (cs > 1 && !b_mn) {  // K-major B: each CTA fetches block_n / cs rows of the tile
      TmapSpec part = B;
      part.box[1] = static_cast<uint32_t>(p.block_n / cs);
      if (encode_tmap(&tmBpart, part) != 0) cs = 1;
    }
  }
  p.cluster = 1;

  p.epi_op = (persist_shape && p.epi_tma) ? p.epi_op : 0;
  if (persist_shape && p.epi_tma) {
    p.pair = 0;
    return launch_gemm_persistent(tmA, tmB, tmO, tmBpart, tmOp, cs, a_mn, b_mn, p, m_tiles, n_tiles, stream);
  }

  if (p.pair) {
    cs = 1;
    if (!b_mn) {  // K-major B: this CTA's box is its half of the tile's rows
      TmapSpec part = B;
      part.box[1] = static_cast<uint32_t>(p.block_n / 2);
      if (encode_tmap(&tmBpart, part) != 0) return -22;
    }
    grid.x = static_cast<unsigned>((m_tiles + 1) / 2 * 2);
  }
  if (cs > 1) {
    p.cluster = cs;
    grid.x = static_cast<unsigned>((m_tiles + cs - 1) / cs * cs);
  }
  if (p.pair) {
    if (!a_mn && !b_mn) return launch_impl<false, false, true>(tmA, tmB, tmO, tmBpart, p, grid, smem, stream);
    if (!a_mn && b_mn) return launch_impl<false, true, true>(tmA, tmB, tmO, tmBpart, p, grid, smem, stream);
    if (a_mn && b_mn) return launch_impl<true, true, true>(tmA, tmB, tmO, tmBpart, p, grid, smem, stream);
    return launch_impl<true, false, true>(tmA, tmB, tmO, tmBpart, p, grid, smem, stream);
  }
  if (!a_mn && !b_mn) return launch_impl<false, false, false>(tmA, tmB, tmO, tmBpart, p, grid, smem, stream);
  if (!a_mn && b_mn) return launch_impl<false, true, false>(tmA, tmB, tmO, tmBpart, p, grid, smem, stream);
  if (a_mn && b_mn) return launch_impl<true, true, false>(tmA, tmB, tmO, tmBpart, p, grid, smem, stream);
  return launch_impl<true, false, false>(tmA, tmB, tmO, tmBpart, p, grid, smem, stream);
}

}  // namespace mdm

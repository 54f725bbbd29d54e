// Persistent, warp-specialised variant of the tcgen05 GEMM engine (same operand modes and fused
// epilogue as gemm_tc.cu, selected by launch_gemm for the non-split, TMA-store-eligible launches).
//
//   one CTA per SM, looping over output tiles (m fastest, so concurrently running CTAs share B tiles in L2)
//   warp 0      : TMA producer (one lane)          -- smem ring of 64-deep K stages, full/empty mbarriers
//   warp 1      : tcgen05.mma issuer (one lane)    -- two 256-column fp32 accumulators in TMEM
//   warps 2..17 : epilogue (512 threads)           -- drain accumulator t while the MMA warp fills t+1;
//                 warp w owns TMEM lanes 32 (w % 4).. and the 16-column quarter (w - 2) / 4 of each 64-column group
// The epilogue stages 64-column groups through 128B-swizzled shared memory and writes them with TMA stores.
#include <stdio.h>
#include <stdlib.h>

#include "gemm_epi.cuh"
#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace mdm {
using namespace ptx;

int g_sm_reserve = 0;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int SLAB_BYTES = 64 * 64 * 2;
constexpr int MAX_STAGES = 8;
constexpr int P_THREADS = 576;      // 18 warps: producer, MMA issuer, 16 epilogue warps

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

struct TileCoord {
  int m_tile, n_tile, z1, z2;
};

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(P_THREADS, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const __grid_constant__ CUtensorMap tmO32, const __grid_constant__ CUtensorMap tmO16,
                          const __grid_constant__ CUtensorMap tmOact, const __grid_constant__ CUtensorMap tmBpart,
                          const __grid_constant__ CUtensorMap tmOp, const GemmParams p, int m_tiles, int n_tiles,
                          int num_tiles, int staging_bytes) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full[2];
  __shared__ __align__(8) uint64_t tmem_empty[2];
  __shared__ __align__(8) uint64_t op_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_bias[256];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  // epilogue staging first (one buffer of staging_bytes, or two when the epilogue operand comes in by TMA -- see the
  // epilogue warps), pipeline stages after it
  const int epi_op = p.epi_op;  // 0 none, 1 residual (fp32, lands in the fp32 staging tile), 2 GELU' source (fp16 tile)
  uint8_t* staging = smem;
  uint8_t* stages = smem + (epi_op ? 2 : 1) * staging_bytes;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nb_alloc = B_MN ? ((p.block_n + 63) / 64) * 64 : p.block_n;
  const int stage_bytes = A_STAGE_BYTES + nb_alloc * 128;
  const int nstages = p.num_stages;
  const int nkb = p.num_kblocks;
  // Cluster of cs CTAs: consecutive CTAs walk consecutive m tiles of the same n tile (m_tiles is padded to a multiple
  // of cs by the launcher), so they share the B tile: each fetches 1/cs of it and multicasts (see gemm_tc.cu).
  const int cs = p.cluster > 1 ? p.cluster : 1;
  const uint32_t crank = cs > 1 ? cluster_ctarank() : 0u;
  const uint16_t cmask = static_cast<uint16_t>((1u << cs) - 1u);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], static_cast<uint32_t>(cs));
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 16);  // one arrival per epilogue warp
      mbar_init(&op_bar[a], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  auto decode = [&](int tile) {
    TileCoord c;
    c.m_tile = tile % m_tiles;
    int t = tile / m_tiles;
    c.n_tile = t % n_tiles;
    t /= n_tiles;
    c.z1 = t % p.nz1;
    c.z2 = t / p.nz1;
    return c;
  };

  if (warp == 0) {
    if (lane == 0) {
      // =========================== TMA producer ===========================
      int stage = 0;
      uint32_t phase = 0;
      const int nslab_b = nb_alloc / 64;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord c = decode(tile);
        const int m0 = c.m_tile * BLOCK_M, n0 = c.n_tile * p.block_n;
        const int az1 = p.a_use_z ? c.z1 + p.a_z1_off : 0, az2 = p.a_use_z ? c.z2 : 0;
        const int bz1 = p.b_use_z ? c.z1 + p.b_z1_off : 0, bz2 = p.b_use_z ? c.z2 : 0;
        int img = 0, th = 0, tw = 0;
        if (p.kind == GEMM_CONV) {
          tw = c.m_tile % p.tiles_w;
          const int t = c.m_tile / p.tiles_w;
          th = t % p.tiles_h;
          img = t / p.tiles_h;
        }
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = stages + stage * stage_bytes;
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
          } else {  // GEMM_CONV
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
          }
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // =========================== MMA issuer ===========================
      const uint32_t idesc = make_idesc_f16(BLOCK_M, p.block_n, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = static_cast<uint32_t>(it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + static_cast<uint32_t>(acc * 256);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(stages + stage * stage_bytes);
          const uint32_t b_base = a_base + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            const uint64_t adesc = A_MN ? make_smem_desc_sw128(a_base + k * 2048, SLAB_BYTES, 1024)
                                        : make_smem_desc_sw128(a_base + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_base + k * 2048, SLAB_BYTES, 1024)
                                        : make_smem_desc_sw128(b_base + k * 32, 16, 1024);
            umma_f16(tacc, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          if (cs > 1) umma_commit_mc(&empty_bar[stage], cmask);
          else umma_commit(&empty_bar[stage]);
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // =========================== epilogue warps ===========================
    const int et = threadIdx.x - 64;                 // 0..511
    const int r = ((warp & 3) << 5) + lane;          // TMEM lane == tile row (warp w may touch lanes 32*(w%4)..)
    const int quarter = (warp - 2) >> 2;             // 16-column chunk of each 64-column group
    const uint32_t swz = static_cast<uint32_t>(r & 7);
    uint8_t* st32 = staging;
    uint8_t* st16 = staging + (p.out_f32 != nullptr ? 32768 : 0);
    uint8_t* stact = st16 + (p.out_f16 != nullptr ? 16384 : 0);
    float alpha = p.alpha;
    if (p.alpha_dev != nullptr) alpha *= __ldg(p.alpha_dev);
    const int ngroups = (p.block_n + 63) / 64;
    int it = 0;
    if (epi_op != 0) {
      // ---- epilogue whose fp32 residual / fp16 GELU' source tile comes in by TMA
      // A thread owns a ROW of the tile, so reading the operand straight from global memory makes every warp request
      // touch 32 different lines (measured: + 23 us on a 16384 x 768 x 768 launch for a 50 MB residual, + 60 us on
      // 16384 x 3072 x 768 for the 100 MB GELU' source). Instead thread et == 0 TMA-loads the operand tile of the NEXT
      // 64-column group into the staging tile that group's results will be written to (same box geometry, same 128B
      // swizzle); each thread then reads, combines and overwrites its own 64 bytes in place, and the tile leaves by TMA
      // store as before. Two staging buffers alternate, so a load is one whole group ahead of its use and the store of
      // group G overlaps the work of group G + 1 (one CTA-wide barrier per group instead of two).
      auto origin = [&](int tile, int& on0, int& o1, int& o2, int& o3) {
        const TileCoord tc = decode(tile);
        on0 = tc.n_tile * p.block_n;
        if (p.kind == GEMM_CONV) {
          const int ttw = tc.m_tile % p.tiles_w;
          const int t = tc.m_tile / p.tiles_w;
          o1 = ttw * p.PW; o2 = (t % p.tiles_h) * p.PH; o3 = t / p.tiles_h;
        } else {
          o1 = tc.m_tile * BLOCK_M; o2 = tc.z1; o3 = tc.z2;
        }
      };
      auto load_op = [&](int tile, int g, int b) {  // thread et == 0
        int on0, o1, o2, o3;
        origin(tile, on0, o1, o2, o3);
        const int colg = on0 + g * 64;
        if (colg >= p.N) return;
        uint8_t* buf = staging + b * staging_bytes;
        if (epi_op == 1) {
          const bool two = colg + 32 < p.N && g * 64 + 32 < p.block_n;
          mbar_expect_tx(&op_bar[b], two ? 32768u : 16384u);
          tma_load_4d(buf, &tmOp, &op_bar[b], colg, o1, o2, o3);
          if (two) tma_load_4d(buf + 16384, &tmOp, &op_bar[b], colg + 32, o1, o2, o3);
        } else {
          mbar_expect_tx(&op_bar[b], 16384u);
          tma_load_4d(buf + (p.out_f32 != nullptr ? 32768 : 0), &tmOp, &op_bar[b], colg, o1, o2, o3);
        }
      };
      int G = 0;                    // running group index of this CTA; staging buffer G & 1
      uint32_t uses0 = 0, uses1 = 0;  // operand loads waited for so far, per buffer (parity of the next wait)
      if (et == 0 && static_cast<int>(blockIdx.x) < num_tiles) load_op(blockIdx.x, 0, 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = static_cast<uint32_t>(it >> 1) & 1;
        int n0, oc1, oc2, oc3;
        origin(tile, n0, oc1, oc2, oc3);
        // bias tile (the previous tile's readers are past its last group barrier)
        epi_bar_sync();
        if (et < 256) s_bias[et] = (p.bias != nullptr && et < p.block_n && n0 + et < p.N) ? __ldg(p.bias + n0 + et) : 0.f;
        epi_bar_sync();
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr_row = tmem_base + static_cast<uint32_t>(acc * 256) + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        for (int g = 0; g < ngroups; ++g, ++G) {
          const int b = G & 1;
          uint8_t* buf = staging + b * staging_bytes;
          uint8_t* b32 = buf;
          uint8_t* b16 = buf + (p.out_f32 != nullptr ? 32768 : 0);
          uint8_t* bact = b16 + (p.out_f16 != nullptr ? 16384 : 0);
          const int cidx = g * 64 + quarter * 16;
          const bool live = cidx < p.block_n;  // warp-uniform
          const int colg = n0 + g * 64;
          const bool has_op = colg < p.N;       // CTA-uniform
          float v[16];
          if (live) {
            tmem_ld16(taddr_row + static_cast<uint32_t>(cidx), v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 bq = *reinterpret_cast<const float4*>(&s_bias[cidx + 4 * q]);
              v[4 * q + 0] = fmaf(v[4 * q + 0], alpha, bq.x);
              v[4 * q + 1] = fmaf(v[4 * q + 1], alpha, bq.y);
              v[4 * q + 2] = fmaf(v[4 * q + 2], alpha, bq.z);
              v[4 * q + 3] = fmaf(v[4 * q + 3], alpha, bq.w);
            }
          }
          if (g == ngroups - 1) {
            // last TMEM read of this tile is done: hand the accumulator back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          if (et == 0) {
            bulk_wait_read();  // the store of group G - 1 has left the other buffer
            int nt = tile, ng = g + 1;
            if (ng == ngroups) {
              nt = tile + gridDim.x;
              ng = 0;
            }
            if (nt < num_tiles) load_op(nt, ng, b ^ 1);
          }
          if (has_op) {
            if (b == 0) mbar_wait(&op_bar[0], uses0++ & 1u);
            else mbar_wait(&op_bar[1], uses1++ & 1u);
          }
          if (live) {
            const int cc = quarter;
            if (epi_op == 1) {
              uint8_t* base = b32 + (cc >> 1) * 16384 + r * 128;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float4* slot = reinterpret_cast<float4*>(base + ((static_cast<uint32_t>((cc & 1) * 4 + q) ^ swz) * 16));
                if (has_op) {
                  const float4 t = *slot;
                  v[4 * q + 0] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
                }
                *slot = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
              }
            } else {
              uint8_t* base = b16 + r * 128;
              uint4* s0 = reinterpret_cast<uint4*>(base + ((static_cast<uint32_t>(cc * 2) ^ swz) * 16));
              uint4* s1 = reinterpret_cast<uint4*>(base + ((static_cast<uint32_t>(cc * 2 + 1) ^ swz) * 16));
              if (has_op) {
                const uint4 u0 = *s0, u1 = *s1;
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
              }
              __half2 h16[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) h16[q] = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
              *s0 = *reinterpret_cast<uint4*>(&h16[0]);
              *s1 = *reinterpret_cast<uint4*>(&h16[4]);
              if (p.out_f32 != nullptr) {
                uint8_t* base32 = b32 + (cc >> 1) * 16384 + r * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  *reinterpret_cast<float4*>(base32 + ((static_cast<uint32_t>((cc & 1) * 4 + q) ^ swz) * 16)) =
                      make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
              }
            }
            if (epi_op == 1 && p.out_f16 != nullptr) {
              __half2 h16[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) h16[q] = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
              uint8_t* base = b16 + r * 128;
              *reinterpret_cast<uint4*>(base + ((static_cast<uint32_t>(cc * 2) ^ swz) * 16)) = *reinterpret_cast<uint4*>(&h16[0]);
              *reinterpret_cast<uint4*>(base + ((static_cast<uint32_t>(cc * 2 + 1) ^ swz) * 16)) = *reinterpret_cast<uint4*>(&h16[4]);
            }
            if (p.out_act_f16 != nullptr) {
              __half2 hact[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float a0 = (p.act == ACT_GELU) ? gelu_erf(v[2 * q]) : v[2 * q];
                const float a1 = (p.act == ACT_GELU) ? gelu_erf(v[2 * q + 1]) : v[2 * q + 1];
                hact[q] = __floats2half2_rn(a0, a1);
              }
              uint8_t* base = bact + r * 128;
              *reinterpret_cast<uint4*>(base + ((static_cast<uint32_t>(cc * 2) ^ swz) * 16)) = *reinterpret_cast<uint4*>(&hact[0]);
              *reinterpret_cast<uint4*>(base + ((static_cast<uint32_t>(cc * 2 + 1) ^ swz) * 16)) = *reinterpret_cast<uint4*>(&hact[4]);
            }
          }
          fence_proxy_async();
          epi_bar_sync();
          if (et == 0) {
            if (colg < p.N) {
              if (p.out_f32 != nullptr) {
                tma_store_4d(&tmO32, b32, colg, oc1, oc2, oc3);
                if (colg + 32 < p.N && g * 64 + 32 < p.block_n) tma_store_4d(&tmO32, b32 + 16384, colg + 32, oc1, oc2, oc3);
              }
              if (p.out_f16 != nullptr) tma_store_4d(&tmO16, b16, colg, oc1, oc2, oc3);
              if (p.out_act_f16 != nullptr) tma_store_4d(&tmOact, bact, colg, oc1, oc2, oc3);
            }
            bulk_commit();
          }
        }
      }
      if (et == 0) bulk_wait_all();
    } else {
    bool stores_pending = false;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const TileCoord c = decode(tile);
      const int acc = it & 1;
      const uint32_t acc_phase = static_cast<uint32_t>(it >> 1) & 1;
      const int m0 = c.m_tile * BLOCK_M, n0 = c.n_tile * p.block_n;
      int img = 0, th = 0, tw = 0;
      bool valid;
      long long row_off;
      int oc1, oc2, oc3;
      if (p.kind == GEMM_CONV) {
        tw = c.m_tile % p.tiles_w;
        const int t = c.m_tile / p.tiles_w;
        th = t % p.tiles_h;
        img = t / p.tiles_h;
        const int ph = r / p.PW, pw = r - ph * p.PW;
        const int h = th * p.PH + ph, w = tw * p.PW + pw;
        valid = (h < p.H) && (w < p.W) && (img < p.nimg);  // (m_tiles is padded to whole clusters)
        row_off = (static_cast<long long>(img * p.H + h) * p.W + w) * p.ldc;
        oc1 = tw * p.PW; oc2 = th * p.PH; oc3 = img;
      } else {
        const int row = m0 + r;
        valid = row < p.M;
        row_off = static_cast<long long>(row) * p.ldc + static_cast<long long>(c.z1) * p.c_z1_stride +
                  static_cast<long long>(c.z2) * p.c_z2_stride;
        oc1 = m0; oc2 = c.z1; oc3 = c.z2;
      }
      // GELU' operand of the first column group: requested now, consumed after the accumulator wait
      uint4 nsrc[2];
      bool nsrc_ok = false;
      auto request = [&](int g) {
        const int cidx = g * 64 + quarter * 16;
        const int col0 = n0 + cidx;
        const long long off0 = row_off + col0;
        const bool inb = cidx < p.block_n && valid && col0 + 16 <= p.N;
        nsrc_ok = p.gelu_grad_src != nullptr && inb && ((off0 & 7) == 0);
        if (nsrc_ok) {
          const uint4* gp = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(p.gelu_grad_src) + off0);
          nsrc[0] = __ldg(gp);
          nsrc[1] = __ldg(gp + 1);
        }
        if (p.residual != nullptr && inb) prefetch_l2(p.residual + off0);
      };
      request(0);
      // bias tile (previous tile's readers are past the group barriers below)
      epi_bar_sync();
      if (et < 256) s_bias[et] = (p.bias != nullptr && et < p.block_n && n0 + et < p.N) ? __ldg(p.bias + n0 + et) : 0.f;
      epi_bar_sync();
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr_row = tmem_base + static_cast<uint32_t>(acc * 256) + (static_cast<uint32_t>((warp & 3) * 32) << 16);
      for (int g = 0; g < ngroups; ++g) {
        // ---- phase A (registers only): TMEM -> alpha/bias/residual/GELU' -> packed results. The global
        // operands are requested first so their latency hides under the TMEM load.
        const int cidx = g * 64 + quarter * 16;
        const int col0 = n0 + cidx;
        const bool live = cidx < p.block_n;  // warp-uniform
        const long long off0 = row_off + col0;
        const bool inb = live && valid && col0 + 16 <= p.N;
        const bool res_fast = p.residual != nullptr && inb && ((off0 & 3) == 0);
        const bool src_fast = nsrc_ok;
        float v[16];
        __half2 h16[8], hact[8];
        float4 rres[4];
        uint4 rsrc[2];
        rsrc[0] = nsrc[0];
        rsrc[1] = nsrc[1];
        if (res_fast) {
          const float4* rp = reinterpret_cast<const float4*>(p.residual + off0);
#pragma unroll
          for (int q = 0; q < 4; ++q) rres[q] = __ldg(rp + q);
        }
        if (g + 1 < ngroups) request(g + 1);  // next group's operands travel while this one is processed
        if (live) {
          tmem_ld16(taddr_row + static_cast<uint32_t>(cidx), v);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bq = *reinterpret_cast<const float4*>(&s_bias[cidx + 4 * q]);
            v[4 * q + 0] = fmaf(v[4 * q + 0], alpha, bq.x);
            v[4 * q + 1] = fmaf(v[4 * q + 1], alpha, bq.y);
            v[4 * q + 2] = fmaf(v[4 * q + 2], alpha, bq.z);
            v[4 * q + 3] = fmaf(v[4 * q + 3], alpha, bq.w);
          }
          if (res_fast) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              v[4 * q + 0] += rres[q].x; v[4 * q + 1] += rres[q].y;
              v[4 * q + 2] += rres[q].z; v[4 * q + 3] += rres[q].w;
            }
          } else if (p.residual != nullptr && valid && col0 < p.N) {
            for (int j = 0; j < 16; ++j)
              if (col0 + j < p.N) v[j] += __ldg(p.residual + off0 + j);
          }
          if (src_fast) {
            const __half2* h0 = reinterpret_cast<const __half2*>(&rsrc[0]);
            const __half2* h1 = reinterpret_cast<const __half2*>(&rsrc[1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 a = __half22float2(h0[q]), b2 = __half22float2(h1[q]);
              v[2 * q] *= gelu_grad(a.x);
              v[2 * q + 1] *= gelu_grad(a.y);
              v[8 + 2 * q] *= gelu_grad(b2.x);
              v[8 + 2 * q + 1] *= gelu_grad(b2.y);
            }
          } else if (p.gelu_grad_src != nullptr && valid && col0 < p.N) {
            const __half* gp = reinterpret_cast<const __half*>(p.gelu_grad_src) + off0;
            for (int j = 0; j < 16; ++j)
              if (col0 + j < p.N) v[j] *= gelu_grad(__half2float(gp[j]));
          }
          if (p.out_f16 != nullptr) {
#pragma unroll
            for (int q = 0; q < 8; ++q) h16[q] = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
          }
          if (p.out_act_f16 != nullptr) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float a0 = (p.act == ACT_GELU) ? gelu_erf(v[2 * q]) : v[2 * q];
              const float a1 = (p.act == ACT_GELU) ? gelu_erf(v[2 * q + 1]) : v[2 * q + 1];
              hact[q] = __floats2half2_rn(a0, a1);
            }
          }
        }
        if (g == ngroups - 1) {
          // last TMEM read of this tile is done: hand the accumulator back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        // ---- the staging buffer may still be read by the previous group's TMA stores
        if (stores_pending) {
          if (et == 0) bulk_wait_read();
          epi_bar_sync();
        }
        // ---- phase B: registers -> 128B-swizzled staging tiles
        if (live) {
          const int cc = quarter;
          if (p.out_f32 != nullptr) {
            uint8_t* base = st32 + (cc >> 1) * 16384 + r * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t chunk = static_cast<uint32_t>((cc & 1) * 4 + q) ^ swz;
              *reinterpret_cast<float4*>(base + chunk * 16) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
          }
          if (p.out_f16 != nullptr) {
            uint8_t* base = st16 + r * 128;
            *reinterpret_cast<uint4*>(base + ((static_cast<uint32_t>(cc * 2) ^ swz) * 16)) = *reinterpret_cast<uint4*>(&h16[0]);
            *reinterpret_cast<uint4*>(base + ((static_cast<uint32_t>(cc * 2 + 1) ^ swz) * 16)) = *reinterpret_cast<uint4*>(&h16[4]);
          }
          if (p.out_act_f16 != nullptr) {
            uint8_t* base = stact + r * 128;
            *reinterpret_cast<uint4*>(base + ((static_cast<uint32_t>(cc * 2) ^ swz) * 16)) = *reinterpret_cast<uint4*>(&hact[0]);
            *reinterpret_cast<uint4*>(base + ((static_cast<uint32_t>(cc * 2 + 1) ^ swz) * 16)) = *reinterpret_cast<uint4*>(&hact[4]);
          }
        }
        fence_proxy_async();
        epi_bar_sync();
        if (et == 0) {
          const int colg = n0 + g * 64;
          if (colg < p.N) {
            if (p.out_f32 != nullptr) {
              if (p.atomic) {
                tma_reduce_add_4d(&tmO32, st32, colg, oc1, oc2, oc3);
                if (colg + 32 < p.N && g * 64 + 32 < p.block_n)
                  tma_reduce_add_4d(&tmO32, st32 + 16384, colg + 32, oc1, oc2, oc3);
              } else {
                tma_store_4d(&tmO32, st32, colg, oc1, oc2, oc3);
                if (colg + 32 < p.N && g * 64 + 32 < p.block_n)
                  tma_store_4d(&tmO32, st32 + 16384, colg + 32, oc1, oc2, oc3);
              }
            }
            if (p.out_f16 != nullptr) tma_store_4d(&tmO16, st16, colg, oc1, oc2, oc3);
            if (p.out_act_f16 != nullptr) tma_store_4d(&tmOact, stact, colg, oc1, oc2, oc3);
          }
          bulk_commit();
        }
        stores_pending = true;
      }
    }
    if (et == 0) bulk_wait_all();
    }  // legacy (operand from global memory) epilogue
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <bool A_MN, bool B_MN>
int launch_p(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmO, const CUtensorMap& tmBpart,
             const CUtensorMap& tmOp, const GemmParams& p, int m_tiles, int n_tiles, int num_tiles, int grid, size_t smem, int staging_bytes,
             cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_persistent_kernel<A_MN, B_MN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_profile) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, stream);
  }
  if (p.cluster > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(P_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = static_cast<unsigned>(p.cluster);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tc_persistent_kernel<A_MN, B_MN>, tmA, tmB, tmO[0], tmO[1], tmO[2],
                                        tmBpart, tmOp, p, m_tiles, n_tiles, num_tiles, staging_bytes);
    if (le != cudaSuccess) return static_cast<int>(le);
  } else {
    gemm_tc_persistent_kernel<A_MN, B_MN><<<grid, P_THREADS, smem, stream>>>(tmA, tmB, tmO[0], tmO[1], tmO[2], tmBpart, tmOp,
                                                                           p, m_tiles, n_tiles, num_tiles, staging_bytes);
  }
  if (g_profile) {
    cudaEventRecord(e1, stream);
    g_profile_events.emplace_back(e0, e1);
    g_profile_params.push_back(p);
    g_profile_majors.push_back((A_MN ? 2 : 0) | (B_MN ? 1 : 0) | 4);
  }
  ++g_launch_count;
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

// Called by launch_gemm when the launch is eligible (no split-K, TMA-store epilogue, not wgrad).
int launch_gemm_persistent(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmO,
                           const CUtensorMap& tmBpart, const CUtensorMap& tmOp, int cluster, int a_mn, int b_mn,
                           GemmParams p, int m_tiles, int n_tiles, cudaStream_t stream) {
  const int nb_alloc = b_mn ? ((p.block_n + 63) / 64) * 64 : p.block_n;
  const int stage_bytes = A_STAGE_BYTES + nb_alloc * 128;
  const int staging = (p.out_f32 ? 32768 : 0) + (p.out_f16 ? 16384 : 0) + (p.out_act_f16 ? 16384 : 0);  // one buffer
  if (p.epi_op != 0 && (222 * 1024 - 2 * staging - 1024) / stage_bytes < 3) p.epi_op = 0;  // keep a 3-stage pipeline
  const int nbuf = p.epi_op != 0 ? 2 : 1;
  int stages = (222 * 1024 - nbuf * staging - 1024) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return -30;
  p.num_stages = stages;
  const size_t smem = static_cast<size_t>(nbuf) * staging + static_cast<size_t>(stages) * stage_bytes + 1024;
  p.cluster = cluster > 1 ? cluster : 1;
  if (p.cluster > 1) m_tiles = (m_tiles + p.cluster - 1) / p.cluster * p.cluster;  // ghost tiles store nothing
  const int num_tiles = m_tiles * n_tiles * p.nz1 * p.nz2;
  // g_sm_reserve SMs are left to a concurrently running collective (mdm_set_sm_reserve): with a static tile stride a
  // CTA that cannot become resident would otherwise serialise its whole share of tiles behind the others
  const int sms = (148 - g_sm_reserve) / p.cluster * p.cluster;
  const int grid = num_tiles < sms ? num_tiles : sms;  // (num_tiles is a multiple of the cluster size)
  if (!a_mn && !b_mn)
    return launch_p<false, false>(tmA, tmB, tmO, tmBpart, tmOp, p, m_tiles, n_tiles, num_tiles, grid, smem, staging, stream);
  if (!a_mn && b_mn)
    return launch_p<false, true>(tmA, tmB, tmO, tmBpart, tmOp, p, m_tiles, n_tiles, num_tiles, grid, smem, staging, stream);
  if (a_mn && b_mn)
    return launch_p<true, true>(tmA, tmB, tmO, tmBpart, tmOp, p, m_tiles, n_tiles, num_tiles, grid, smem, staging, stream);
  return launch_p<true, false>(tmA, tmB, tmO, tmBpart, tmOp, p, m_tiles, n_tiles, num_tiles, grid, smem, staging, stream);
}

}  // namespace mdm

// Fused attention for the denoiser's SelfAttention blocks (reference models/unet.py:276-313):
//   h = softmax(q k^T / sqrt(d)) v  +  softmax(q k_c^T / sqrt(d)) v_c        (two independent softmaxes)
// with q,k,v = channel thirds of the qkv 1x1 conv, k_c,v_c = halves of kv_cond(LayerNorm(cond)),
// 8 heads of d = C/8 channels, optional token mask on the cross branch.
//
// Forward: one CTA per (128 queries, head, sample). Scores never leave the SM: S = Q K^T by
// tcgen05.mma into TMEM, exact two-pass softmax (pass A: row max / sum over all key chunks,
// pass B: recompute S, P = exp2(..)/l in fp16 to 128B-swizzled smem, O += P V on the tensor core).
// Backward: one CTA per (128 keys, head, sample) looping over query tiles: recomputes P from the
// saved row statistics, dV += P^T dO, dP = dO V^T, dS = P (dP - D) alpha, dK += dS^T Q, dQ += dS K
// (fp32 atomics). All five products are tcgen05.mma on operands staged by TMA; P and dS tiles are
// written once to smem and read both K-major and MN-major (transposed) through descriptors.
#include <math.h>

#include "engine.cuh"
#include "mdm_b200.h"
#include "ptx.cuh"

namespace mdm {
using namespace ptx;

namespace {

constexpr int AT_THREADS = 128;
constexpr int BWD_THREADS = 256;  // backward: two warpgroups split every tile by columns
constexpr int TILE = 128;           // queries per CTA (fwd) / keys per CTA (bwd); key chunk size
constexpr int KB_BYTES = 128 * 128; // one [128 rows][64 fp16] k-block / slab

struct AttnParams {
  int T, S, d, heads, B;
  int kblocks;        // ceil(d / 64)
  float alpha_log2e;  // (1/sqrt(d)) * log2(e)
  float alpha;
  const float* mask;  // [B][S] or null (cross branch)
  // forward outputs
  __half* h16;        // [B*T][C] summed output
  __half* oself16;    // [B*T][C] self branch only (training) or null
  float* stats;       // [B][heads][2 branches][T][2] = (m2, 1/l)
  int C;
  // backward
  const float* Dterm;  // [B][heads][2][T]
  float* dq32;         // [B*T][C] fp32, accumulated atomically (zero on entry)
  __half* dqkv16;      // [B*T][3C]: dK at +C, dV at +2C
  __half* dkv16;       // [B*S][2C]: dKc at +0, dVc at +C
};

__device__ __forceinline__ uint64_t desc_k(uint32_t base, int k16) {  // K-major, 16-element step k16
  return make_smem_desc_sw128(base + k16 * 32, 16, 1024);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t base, int k16, uint32_t slab_bytes) {  // MN-major
  return make_smem_desc_sw128(base + k16 * 2048, slab_bytes, 1024);
}

__device__ __forceinline__ void st_swz_half8(uint8_t* tile, int row, int col8, const __half2* h) {
  // tile: [128 rows][64 halves] k-blocks of 16 KB; col8 = index of the 8-half (16 B) chunk in the row of 128 keys
  const int kb = col8 >> 3, q = col8 & 7;
  uint8_t* p = tile + kb * KB_BYTES + row * 128 + ((q ^ (row & 7)) << 4);
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(h);
}

// ------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(AT_THREADS)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmQKVv,
                const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmKVv,
                const AttnParams p) {
  // No static shared memory: the dynamic segment starts the CTA's window, 1024-aligned as the swizzled
  // tiles need, so two CTAs (2 x 112 KB of tiles at d = 64) fit one SM.
  extern __shared__ __align__(1024) uint8_t attn_fwd_smem[];
  uint8_t* smem = attn_fwd_smem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int qt = blockIdx.x, hd = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * TILE;
  const int kbk = p.kblocks;
  const int stage_bytes = 2 * kbk * KB_BYTES;  // K then V
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + kbk * KB_BYTES;
  uint8_t* sP = sKV + 2 * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * KB_BYTES);
  uint64_t& q_bar = bars[0];
  uint64_t* kv_bar = bars + 1;
  uint64_t& s_bar = bars[3];
  uint64_t& o_bar = bars[4];
  uint32_t& tmem_base_smem = *reinterpret_cast<uint32_t*>(bars + 5);
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  if (tid == 0) {
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmQKVv);
    prefetch_tmap(&tmKV);
    prefetch_tmap(&tmKVv);
    mbar_init(&q_bar, 1);
    mbar_init(&kv_bar[0], 1);
    mbar_init(&kv_bar[1], 1);
    mbar_init(&s_bar, 1);
    mbar_init(&o_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  const uint32_t tS = tmem, tO = tmem + 128;
  const uint32_t trow = static_cast<uint32_t>(warp * 32) << 16;

  // Q tile
  if (tid == 0) {
    mbar_expect_tx(&q_bar, kbk * KB_BYTES);
    for (int kb = 0; kb < kbk; ++kb) tma_load_4d(sQ + kb * KB_BYTES, &tmQKV, &q_bar, kb * 64, q0, hd, b);
  }
  const uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0);
  const uint32_t idesc_pv = make_idesc_f16(128, (p.d + 15) / 16 * 16, 0, 1);
  uint32_t kv_phase[2] = {0, 0};
  uint32_t s_phase = 0;
  bool o_started = false;

  // K / V chunk loads form one queue across passes and branches (two stages), so the first chunk of
  // the next pass or branch is already in flight while the current one finishes.
  const int nself = (p.T + TILE - 1) / TILE;
  const int ncross = p.S > 0 ? (p.S + TILE - 1) / TILE : 0;
  const int self_p0 = nself > 1 ? nself : 0;   // statistics-pass items of the self branch
  const int items_self = self_p0 + nself;
  const int cross_p0 = ncross > 1 ? ncross : 0;
  const int n_items = items_self + cross_p0 + ncross;
  auto issue_item = [&](int i) {
    const bool cross = i >= items_self;
    const int k = cross ? i - items_self : i;
    const int p0 = cross ? cross_p0 : self_p0;
    const bool with_v = k >= p0;
    const int key0 = (with_v ? k - p0 : k) * TILE;
    const int st = i & 1;
    uint8_t* dst = sKV + st * stage_bytes;
    mbar_expect_tx(&kv_bar[st], (with_v ? 2 : 1) * kbk * KB_BYTES);
    for (int kb = 0; kb < kbk; ++kb) {
      if (!cross) tma_load_4d(dst + kb * KB_BYTES, &tmQKV, &kv_bar[st], kb * 64, key0, p.heads + hd, b);
      else tma_load_4d(dst + kb * KB_BYTES, &tmKV, &kv_bar[st], kb * 64, key0, hd, b);
    }
    if (with_v) {
      for (int sl = 0; sl < kbk; ++sl) {
        if (!cross) tma_load_4d(dst + (kbk + sl) * KB_BYTES, &tmQKVv, &kv_bar[st], sl * 64, key0, 2 * p.heads + hd, b);
        else tma_load_4d(dst + (kbk + sl) * KB_BYTES, &tmKVv, &kv_bar[st], sl * 64, key0, p.heads + hd, b);
      }
    }
  };
  int item = 0;  // queue position of the chunk being consumed
  if (tid == 0 && n_items > 0) issue_item(0);

  if (tid == 0) mbar_wait(&q_bar, 0);
  __syncthreads();

  for (int branch = 0; branch < 2; ++branch) {
    const bool cross = branch == 1;
    const int nkeys = cross ? p.S : p.T;
    if (nkeys <= 0) continue;
    const int nchunks = (nkeys + TILE - 1) / TILE;
    const float* mk = (cross && p.mask != nullptr) ? p.mask + static_cast<long long>(b) * p.S : nullptr;
    float m2 = -INFINITY, l = 0.f;
    // Self branch over several chunks: pass 0 finds the row maximum only, pass 1 writes unnormalised
    // P = 2^(s - m) (one exponential per score), and O is divided by l afterwards.
    const bool deferred = branch == 0 && nchunks > 1;
    // pass 0: statistics (skipped when a single chunk holds the whole row), pass 1: P and O += P V
    for (int pass = (nchunks > 1 ? 0 : 1); pass < 2; ++pass) {
      for (int c = 0; c < nchunks; ++c, ++item) {
        const int st = item & 1;
        uint8_t* sK = sKV + st * stage_bytes;
        uint8_t* sV = sK + kbk * KB_BYTES;
        if (tid == 0) {
          mbar_wait(&kv_bar[st], kv_phase[st]);
          tc_fence_after();
          const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
          for (int kb = 0; kb < kbk; ++kb)
            for (int k = 0; k < 4; ++k)
              umma_f16(tS, desc_k(qa + kb * KB_BYTES, k), desc_k(ka + kb * KB_BYTES, k), idesc_s, (kb | k) ? 1u : 0u);
          umma_commit(&s_bar);
        }
        kv_phase[st] ^= 1;  // every thread tracks the same parity
        mbar_wait(&s_bar, s_phase);
        s_phase ^= 1;
        tc_fence_after();
        // the previous item's MMAs have retired (the commit covers all earlier ones): its stage is free
        if (tid == 0 && item + 1 < n_items) issue_item(item + 1);
        const int key0 = c * TILE;
        // the whole 128-key row of this chunk in registers (one TMEM round trip), raw scores
        float v[TILE];
#pragma unroll
        for (int j = 0; j < TILE; j += 32) tmem_ld32_nowait(tS + trow + j, v + j);
        tmem_ld_wait();
        const bool tail = key0 + TILE > nkeys;
        if (tail || mk != nullptr) {
#pragma unroll
          for (int e = 0; e < TILE; ++e) {
            const int key = key0 + e;
            const bool ok = key < nkeys && (mk == nullptr || mk[key] != 0.f);
            v[e] = ok ? v[e] : -INFINITY;
          }
        }
        if (pass == 0 || nchunks == 1) {  // running row maximum (alpha > 0: max commutes with the scaling)
          float cm = -INFINITY;
#pragma unroll
          for (int e = 0; e < TILE; ++e) cm = fmaxf(cm, v[e]);
          cm *= p.alpha_log2e;
          if (pass == 0 && !deferred && cm > -INFINITY) {
            // several cross chunks: the sum is needed before P can be written, so pass 0 carries it too
            const float mn = fmaxf(m2, cm);
            float sum = 0.f;
#pragma unroll
            for (int e = 0; e < TILE; ++e) sum += ex2_approx(fmaf(v[e], p.alpha_log2e, -mn));
            l = l * (m2 > -INFINITY ? ex2_approx(m2 - mn) : 0.f) + sum;
          }
          m2 = fmaxf(m2, cm);
        }
        if (pass == 0) {
          tc_fence_before();
          __syncthreads();  // everyone is done with S before the next chunk overwrites it
        } else {
          const float mref = m2 > -INFINITY ? m2 : 0.f;
          float sum = 0.f;
#pragma unroll
          for (int e = 0; e < TILE; ++e) {
            v[e] = ex2_approx(fmaf(v[e], p.alpha_log2e, -mref));  // 2^-inf = 0 for masked keys
            sum += v[e];
          }
          if (deferred || nchunks == 1) l += sum;
          // deferred: P stays unnormalised (<= 1) and O is divided by l once after the branch
          const float pscale = deferred ? 1.f : (l > 0.f ? 1.0f / l : 0.f);
#pragma unroll
          for (int j = 0; j < TILE; j += 8) {
            __half2 h[4];
#pragma unroll
            for (int e = 0; e < 8; e += 2) h[e >> 1] = __floats2half2_rn(v[j + e] * pscale, v[j + e + 1] * pscale);
            st_swz_half8(sP, tid, j >> 3, h);
          }
          fence_proxy_async();
          tc_fence_before();
          __syncthreads();
          if (tid == 0) {
            tc_fence_after();
            const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
            for (int k = 0; k < 8; ++k)  // 128 keys = 8 x 16
              umma_f16(tO, desc_k(pa + (k >> 2) * KB_BYTES, k & 3), desc_mn(va, k, KB_BYTES), idesc_pv,
                       (o_started || k) ? 1u : 0u);
            o_started = true;
          }
        }
      }
    }
    // row statistics for the backward pass
    if (p.stats != nullptr && q0 + tid < p.T) {
      float* st = p.stats + ((((static_cast<long long>(b) * p.heads + hd) * 2 + branch) * p.T) + q0 + tid) * 2;
      st[0] = m2;
      st[1] = l > 0.f ? 1.0f / l : 0.f;
    }
    if (branch == 0 && (p.oself16 != nullptr || deferred)) {
      // self-branch output alone (the backward needs rowsum(dO * O_branch) per branch); a deferred
      // normalisation is applied here and written back so the cross branch accumulates on top of it
      if (tid == 0) umma_commit(&o_bar);
      mbar_wait(&o_bar, 0);
      tc_fence_after();
      {
        const bool valid = q0 + tid < p.T;
        const float oscale = deferred ? (l > 0.f ? 1.0f / l : 0.f) : 1.f;
        __half* dst = p.oself16 != nullptr ? p.oself16 + (static_cast<long long>(b) * p.T + q0 + tid) * p.C + hd * p.d
                                           : nullptr;
        for (int j = 0; j < p.d; j += 16) {
          float v[16];
          tmem_ld16(tO + trow + j, v);
          if (deferred) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] *= oscale;
            tmem_st16(tO + trow + j, v);
          }
          if (!valid || dst == nullptr) continue;
          __half2 h[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) h[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
          if (j + 16 <= p.d) {
            *reinterpret_cast<uint4*>(dst + j) = *reinterpret_cast<uint4*>(&h[0]);
            *reinterpret_cast<uint4*>(dst + j + 8) = *reinterpret_cast<uint4*>(&h[4]);
          } else {
            for (int e = 0; e < 16 && j + e < p.d; ++e) dst[j + e] = __float2half_rn(v[e]);
          }
        }
        if (deferred) tmem_st_wait();
      }
      tc_fence_before();
      __syncthreads();
    }
  }
  // final output
  if (tid == 0) umma_commit(&s_bar);
  mbar_wait(&s_bar, s_phase);
  tc_fence_after();
  {
    const bool valid = q0 + tid < p.T;
    __half* dst = p.h16 + (static_cast<long long>(b) * p.T + q0 + tid) * p.C + hd * p.d;
    for (int j = 0; j < p.d; j += 16) {
      float v[16];
      tmem_ld16(tO + trow + j, v);
      if (!valid) continue;
      __half2 h[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) h[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
      if (j + 16 <= p.d) {
        *reinterpret_cast<uint4*>(dst + j) = *reinterpret_cast<uint4*>(&h[0]);
        *reinterpret_cast<uint4*>(dst + j + 8) = *reinterpret_cast<uint4*>(&h[4]);
      } else {
        for (int e = 0; e < 16 && j + e < p.d; ++e) dst[j + e] = __float2half_rn(v[e]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------ backward
// D[b][h][branch][q] = sum_c dO[q][c] * O_branch[q][c]  with O_cross = h - O_self
__global__ void attn_bwd_prep_kernel(const __half* __restrict__ dO, const __half* __restrict__ h16,
                                     const __half* __restrict__ oself16, float* __restrict__ Dterm, int T, int C,
                                     int heads, int d, long long rows) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long long b = row / T;
  const int q = static_cast<int>(row - b * T);
  for (int hd = 0; hd < heads; ++hd) {
    float s0 = 0.f, s1 = 0.f;
    for (int c = lane; c < d; c += 32) {
      const long long o = row * C + hd * d + c;
      const float g = __half2float(dO[o]);
      const float os = oself16 != nullptr ? __half2float(oself16[o]) : __half2float(h16[o]);
      const float oc = __half2float(h16[o]) - os;
      s0 += g * os;
      s1 += g * oc;
    }
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane == 0) {
      Dterm[((b * heads + hd) * 2 + 0) * T + q] = s0;
      Dterm[((b * heads + hd) * 2 + 1) * T + q] = s1;
    }
  }
}

__global__ void __launch_bounds__(BWD_THREADS)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmKV,
                const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmDQ,
                const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t kv_bar, qd_bar, mma_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  // 8 warps: warp w owns TMEM lanes 32 (w % 4) .. +31 (= tile rows) and column half w / 4
  const int tid = threadIdx.x, warp = tid >> 5;
  const int lrow = tid & 127, wg = tid >> 7;
  const int hd = blockIdx.y, b = blockIdx.z;
  const int n_self = (p.T + TILE - 1) / TILE;
  const bool cross = static_cast<int>(blockIdx.x) >= n_self;
  const int jt = cross ? blockIdx.x - n_self : blockIdx.x;
  const int key0 = jt * TILE;
  const int nkeys = cross ? p.S : p.T;
  const int branch = cross ? 1 : 0;
  const int kbk = p.kblocks;
  uint8_t* sK = smem;
  uint8_t* sV = sK + kbk * KB_BYTES;
  uint8_t* sQ = sV + kbk * KB_BYTES;
  uint8_t* sDO = sQ + kbk * KB_BYTES;
  uint8_t* sP = sDO + kbk * KB_BYTES;
  uint8_t* sDS = sP + 2 * KB_BYTES;

  if (tid == 0) {
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmKV);
    prefetch_tmap(&tmDO);
    prefetch_tmap(&tmDQ);
    mbar_init(&kv_bar, 1);
    mbar_init(&qd_bar, 1);
    mbar_init(&mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  const uint32_t tS = tmem, tDP = tmem + 128, tDV = tmem + 256, tDK = tmem + 384;
  const uint32_t trow = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const int dN = (p.d + 15) / 16 * 16;
  const uint32_t id_kk = make_idesc_f16(128, 128, 0, 0);   // S, dP
  const uint32_t id_tt = make_idesc_f16(128, dN, 1, 1);    // dV, dK (A = P^T / dS^T, B = dO / Q rows)
  const uint32_t id_kt = make_idesc_f16(128, dN, 0, 1);    // dQ = dS K

  if (tid == 0) {
    mbar_expect_tx(&kv_bar, 2 * kbk * KB_BYTES);
    for (int kb = 0; kb < kbk; ++kb) {
      if (!cross) {
        tma_load_4d(sK + kb * KB_BYTES, &tmQKV, &kv_bar, kb * 64, key0, p.heads + hd, b);
        tma_load_4d(sV + kb * KB_BYTES, &tmQKV, &kv_bar, kb * 64, key0, 2 * p.heads + hd, b);
      } else {
        tma_load_4d(sK + kb * KB_BYTES, &tmKV, &kv_bar, kb * 64, key0, hd, b);
        tma_load_4d(sV + kb * KB_BYTES, &tmKV, &kv_bar, kb * 64, key0, p.heads + hd, b);
      }
    }
    mbar_wait(&kv_bar, 0);
  }
  __syncthreads();
  const float* mk = (cross && p.mask != nullptr) ? p.mask + static_cast<long long>(b) * p.S : nullptr;
  uint32_t qd_phase = 0, mma_phase = 0;
  const int nq = (p.T + TILE - 1) / TILE;
  auto load_qd = [&](int q0) {
    mbar_expect_tx(&qd_bar, 2 * kbk * KB_BYTES);
    for (int kb = 0; kb < kbk; ++kb) {
      tma_load_4d(sQ + kb * KB_BYTES, &tmQKV, &qd_bar, kb * 64, q0, hd, b);
      tma_load_4d(sDO + kb * KB_BYTES, &tmDO, &qd_bar, kb * 64, q0, hd, b);
    }
  };
  if (tid == 0) load_qd(0);
  for (int it = 0; it < nq; ++it) {
    const int q0 = it * TILE;
    if (tid == 0) {
      mbar_wait(&qd_bar, qd_phase);
      tc_fence_after();
      const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK), va = smem_u32(sV), da = smem_u32(sDO);
      for (int kb = 0; kb < kbk; ++kb)
        for (int k = 0; k < 4; ++k)
          umma_f16(tS, desc_k(qa + kb * KB_BYTES, k), desc_k(ka + kb * KB_BYTES, k), id_kk, (kb | k) ? 1u : 0u);
      for (int kb = 0; kb < kbk; ++kb)
        for (int k = 0; k < 4; ++k)
          umma_f16(tDP, desc_k(da + kb * KB_BYTES, k), desc_k(va + kb * KB_BYTES, k), id_kk, (kb | k) ? 1u : 0u);
      // the previous tile's dQ reduction must have read its staging (= the P / dS tiles) before they
      // are rewritten; the commit below is issued after this wait, so mma_bar orders it for everyone
      bulk_wait_read();
      umma_commit(&mma_bar);
    }
    qd_phase ^= 1;
    mbar_wait(&mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    // thread = query row x 64-key half
    {
      const int q = q0 + lrow;
      const bool qok = q < p.T;
      float m2 = 0.f, inv_l = 0.f, Dq = 0.f;
      if (qok) {
        const float* st = p.stats + ((((static_cast<long long>(b) * p.heads + hd) * 2 + branch) * p.T) + q) * 2;
        m2 = st[0];
        inv_l = st[1];
        Dq = p.Dterm[((static_cast<long long>(b) * p.heads + hd) * 2 + branch) * p.T + q];
      }
      if (!(inv_l > 0.f)) {  // row outside the tile or fully masked: P = 0 without inf arithmetic
        inv_l = 0.f;
        m2 = 0.f;
      }
      const float nDq = -Dq * p.alpha;
      const int j0 = wg * 64;
      float s[64], dp[64];
      tmem_ld32_nowait(tS + trow + j0, s);
      tmem_ld32_nowait(tS + trow + j0 + 32, s + 32);
      tmem_ld32_nowait(tDP + trow + j0, dp);
      tmem_ld32_nowait(tDP + trow + j0 + 32, dp + 32);
      tmem_ld_wait();
      const bool plain = mk == nullptr && key0 + TILE <= nkeys;  // uniform: no per-key predicate needed
#pragma unroll
      for (int j = 0; j < 64; j += 8) {
        __half2 hp[4], hs[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          float pv[2], ds[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            float pe = ex2_approx(fmaf(s[j + e + u], p.alpha_log2e, -m2)) * inv_l;
            if (!plain) {
              const int key = key0 + j0 + j + e + u;
              const bool ok = key < nkeys && (mk == nullptr || mk[key] != 0.f);
              pe = ok ? pe : 0.f;
            }
            pv[u] = pe;
            ds[u] = pe * fmaf(dp[j + e + u], p.alpha, nDq);  // P (dP - D) / sqrt(d)
          }
          hp[e >> 1] = __floats2half2_rn(pv[0], pv[1]);
          hs[e >> 1] = __floats2half2_rn(ds[0], ds[1]);
        }
        st_swz_half8(sP, lrow, (j0 + j) >> 3, hp);
        st_swz_half8(sDS, lrow, (j0 + j) >> 3, hs);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t pa = smem_u32(sP), sa = smem_u32(sDS), qa = smem_u32(sQ), ka = smem_u32(sK), da = smem_u32(sDO);
      // dV[key][d] += P^T dO ; dK[key][d] += dS^T Q   (contraction over the 128 query rows)
      for (int k = 0; k < 8; ++k) umma_f16(tDV, desc_mn(pa, k, KB_BYTES), desc_mn(da, k, KB_BYTES), id_tt, (it | k) ? 1u : 0u);
      for (int k = 0; k < 8; ++k) umma_f16(tDK, desc_mn(sa, k, KB_BYTES), desc_mn(qa, k, KB_BYTES), id_tt, (it | k) ? 1u : 0u);
      // dQ[q][d] = dS K   (contraction over the 128 keys), into the S columns (S is dead now)
      for (int k = 0; k < 8; ++k) umma_f16(tS, desc_k(sa + (k >> 2) * KB_BYTES, k & 3), desc_mn(ka, k, KB_BYTES), id_kt, k ? 1u : 0u);
      umma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    // all MMAs that read Q / dO have retired: fetch the next query tile under the dQ drain
    if (tid == 0 && it + 1 < nq) load_qd((it + 1) * TILE);
    {
      // dQ tile: TMEM -> 128B-swizzled fp32 staging (the dead P / dS tiles) -> one TMA reduce-add per
      // 32 columns; rows beyond T and columns beyond d are clipped by the tensor map
      uint8_t* stg = sP;
      const uint32_t swz = static_cast<uint32_t>(lrow & 7);
      for (int j = wg * 16; j < p.d; j += 32) {
        float v[16];
        tmem_ld16(tS + trow + j, v);
        uint8_t* base = stg + (j >> 5) * KB_BYTES + lrow * 128;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t chunk = static_cast<uint32_t>(((j & 16) >> 2) + e) ^ swz;
          *reinterpret_cast<float4*>(base + chunk * 16) = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      for (int g = 0; g * 32 < p.d; ++g) tma_reduce_add_4d(&tmDQ, sP + g * KB_BYTES, g * 32, q0, hd, b);
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait_all();
  // dK, dV of this key tile
  {
    const int key = key0 + lrow;
    const bool ok = key < nkeys;
    __half *dk, *dv;
    if (!cross) {
      __half* base = p.dqkv16 + (static_cast<long long>(b) * p.T + key) * 3 * p.C + hd * p.d;
      dk = base + p.C;
      dv = base + 2 * p.C;
    } else {
      __half* base = p.dkv16 + (static_cast<long long>(b) * p.S + key) * 2 * p.C + hd * p.d;
      dk = base;
      dv = base + p.C;
    }
    for (int j = wg * 16; j < p.d; j += 32) {
      float a[16], c[16];
      tmem_ld16(tDK + trow + j, a);
      tmem_ld16(tDV + trow + j, c);
      if (!ok) continue;
      if (j + 16 <= p.d) {
        __half2 ha[8], hc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          ha[e] = __floats2half2_rn(a[2 * e], a[2 * e + 1]);
          hc[e] = __floats2half2_rn(c[2 * e], c[2 * e + 1]);
        }
        *reinterpret_cast<uint4*>(dk + j) = *reinterpret_cast<uint4*>(&ha[0]);
        *reinterpret_cast<uint4*>(dk + j + 8) = *reinterpret_cast<uint4*>(&ha[4]);
        *reinterpret_cast<uint4*>(dv + j) = *reinterpret_cast<uint4*>(&hc[0]);
        *reinterpret_cast<uint4*>(dv + j + 8) = *reinterpret_cast<uint4*>(&hc[4]);
      } else {
        for (int e = 0; e < 16 && j + e < p.d; ++e) {
          dk[j + e] = __float2half_rn(a[e]);
          dv[j + e] = __float2half_rn(c[e]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

__global__ void cast_strided_kernel(const float* __restrict__ in, __half* __restrict__ out, long long rows, int C,
                                    int ld_out) {
  const long long total = rows * (C / 4);
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long gs = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += gs) {
    const long long r = i / (C / 4);
    const int c = static_cast<int>(i - r * (C / 4)) * 4;
    const float4 v = *reinterpret_cast<const float4*>(in + r * C + c);
    __half2 a = __floats2half2_rn(v.x, v.y), bq = __floats2half2_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&bq);
    *reinterpret_cast<uint2*>(out + r * ld_out + c) = o;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(q);
  }
  return fn;
}
// 4-D fp16 view (inner d, rows, slots, batch), box {64, 128, 1, 1}
void head_map(CUtensorMap* m, const void* ptr, int d, int rows, long long row_stride, int slots, long long slot_stride,
              int batch, long long batch_stride) {
  EncodeTiledFn fn = encode_fn();
  MDM_CHECK(fn != nullptr, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(slots),
                        static_cast<cuuint64_t>(batch)};
  cuuint64_t str[3] = {static_cast<cuuint64_t>(row_stride) * 2, static_cast<cuuint64_t>(slot_stride) * 2,
                       static_cast<cuuint64_t>(batch_stride) * 2};
  cuuint32_t box[4] = {64, 128, 1, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, str, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MDM_CHECK(r == CUDA_SUCCESS, "attention tensor map encode failed");
}

// fp32 [B*T][C] viewed per head: (d, rows, heads, batch), box {32, 128, 1, 1}
void head_map_f32(CUtensorMap* m, const void* ptr, int d, int rows, long long row_stride, int slots, int batch) {
  EncodeTiledFn fn = encode_fn();
  MDM_CHECK(fn != nullptr, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(slots),
                        static_cast<cuuint64_t>(batch)};
  cuuint64_t str[3] = {static_cast<cuuint64_t>(row_stride) * 4, static_cast<cuuint64_t>(d) * 4,
                       static_cast<cuuint64_t>(rows) * row_stride * 4};
  cuuint32_t box[4] = {32, 128, 1, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), dims, str, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MDM_CHECK(r == CUDA_SUCCESS, "attention dQ tensor map encode failed");
}

}  // namespace

// ------------------------------------------------------------------------------------------ host API
void attention_forward(const __half* qkv, const __half* kv, const float* mask, int B, int T, int S, int C, int heads,
                       __half* h16, __half* oself16, float* stats, cudaStream_t st) {
  const int d = C / heads;
  MDM_CHECK(d % 8 == 0 && d <= 128, "head dim must be a multiple of 8 and <= 128");
  AttnParams p{};
  p.T = T; p.S = kv != nullptr ? S : 0; p.d = d; p.heads = heads; p.B = B; p.C = C;
  p.kblocks = (d + 63) / 64;
  p.alpha = 1.0f / sqrtf(static_cast<float>(d));
  p.alpha_log2e = p.alpha * 1.4426950408889634f;
  p.mask = mask;
  p.h16 = h16; p.oself16 = oself16; p.stats = stats;
  alignas(64) CUtensorMap mq, mqv, mk, mkv;
  head_map(&mq, qkv, d, T, 3ll * C, 3 * heads, d, B, static_cast<long long>(T) * 3 * C);
  mqv = mq;
  if (kv != nullptr) head_map(&mk, kv, d, S, 2ll * C, 2 * heads, d, B, static_cast<long long>(S) * 2 * C);
  else mk = mq;
  mkv = mk;
  const int kbk = p.kblocks;
  const size_t smem = static_cast<size_t>(kbk) * KB_BYTES + 2 * (2 * kbk * KB_BYTES) + 2 * KB_BYTES + 64;
  static bool attr = false;
  if (!attr) {
    MDM_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  dim3 grid((T + TILE - 1) / TILE, heads, B);
  attn_fwd_kernel<<<grid, AT_THREADS, smem, st>>>(mq, mqv, mk, mkv, p);
  ++g_launch_count;
  MDM_CUDA(cudaGetLastError());
}

void attention_backward(const __half* qkv, const __half* kv, const float* mask, const __half* dO, const __half* h16,
                        const __half* oself16, const float* stats, int B, int T, int S, int C, int heads,
                        float* Dterm, float* dq32, __half* dqkv16, __half* dkv16, cudaStream_t st) {
  const int d = C / heads;
  AttnParams p{};
  p.T = T; p.S = kv != nullptr ? S : 0; p.d = d; p.heads = heads; p.B = B; p.C = C;
  p.kblocks = (d + 63) / 64;
  p.alpha = 1.0f / sqrtf(static_cast<float>(d));
  p.alpha_log2e = p.alpha * 1.4426950408889634f;
  p.mask = mask;
  p.stats = const_cast<float*>(stats);
  p.Dterm = Dterm; p.dq32 = dq32; p.dqkv16 = dqkv16; p.dkv16 = dkv16;
  const long long rows = static_cast<long long>(B) * T;
  attn_bwd_prep_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(dO, h16, oself16, Dterm, T, C, heads, d, rows);
  ++g_launch_count;
  MDM_CUDA(cudaMemsetAsync(dq32, 0, sizeof(float) * rows * C, st));
  alignas(64) CUtensorMap mq, mk, mdo, mdq;
  head_map_f32(&mdq, dq32, d, T, C, heads, B);
  head_map(&mq, qkv, d, T, 3ll * C, 3 * heads, d, B, static_cast<long long>(T) * 3 * C);
  if (kv != nullptr) head_map(&mk, kv, d, S, 2ll * C, 2 * heads, d, B, static_cast<long long>(S) * 2 * C);
  else mk = mq;
  head_map(&mdo, dO, d, T, C, heads, d, B, static_cast<long long>(T) * C);
  const int kbk = p.kblocks;
  const size_t smem = static_cast<size_t>(4 * kbk + 4) * KB_BYTES + 1024;
  static bool attr = false;
  if (!attr) {
    MDM_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  const int n_self = (T + TILE - 1) / TILE, n_cross = p.S > 0 ? (p.S + TILE - 1) / TILE : 0;
  dim3 grid(n_self + n_cross, heads, B);
  attn_bwd_kernel<<<grid, BWD_THREADS, smem, st>>>(mq, mk, mdo, mdq, p);
  ++g_launch_count;
  MDM_CUDA(cudaGetLastError());
  // dQ: fp32 accumulator -> fp16 into the q third of dqkv
  cast_strided_kernel<<<148 * 8, 256, 0, st>>>(dq32, dqkv16, rows, C, 3 * C);
  ++g_launch_count;
}

}  // namespace mdm

// ------------------------------------------------------------------------------------------ C ABI (tests)
#define MDM_TRY(...)                  \
  try {                               \
    __VA_ARGS__;                      \
    return 0;                         \
  } catch (const std::exception& e) { \
    mdm::set_error("%s", e.what());   \
    return -1;                        \
  }

extern "C" {

int mdm_op_attention_fwd(const void* qkv16, const void* kv16, const float* mask, int B, int T, int S, int C, int heads,
                         void* h16, void* oself16, float* stats, mdm_stream_t stream) {
  MDM_TRY(mdm::attention_forward(static_cast<const __half*>(qkv16), static_cast<const __half*>(kv16), mask, B, T, S, C,
                                 heads, static_cast<__half*>(h16), static_cast<__half*>(oself16), stats,
                                 static_cast<cudaStream_t>(stream)))
}

int mdm_op_attention_bwd(const void* qkv16, const void* kv16, const float* mask, const void* dO16, const void* h16,
                         const void* oself16, const float* stats, int B, int T, int S, int C, int heads, float* Dterm,
                         float* dq32, void* dqkv16, void* dkv16, mdm_stream_t stream) {
  MDM_TRY(mdm::attention_backward(static_cast<const __half*>(qkv16), static_cast<const __half*>(kv16), mask,
                                  static_cast<const __half*>(dO16), static_cast<const __half*>(h16),
                                  static_cast<const __half*>(oself16), stats, B, T, S, C, heads, Dterm, dq32,
                                  static_cast<__half*>(dqkv16), static_cast<__half*>(dkv16),
                                  static_cast<cudaStream_t>(stream)))
}
}

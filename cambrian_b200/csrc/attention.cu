// cambrian_b200 — tcgen05 flash attention (forward + backward) for sm_100a.
//
// Covers the two softmax-attention shapes of the hot path:
//   * ViT towers (A1-A3): non-causal, head_dim 64 (CLIP, DINOv2) / 72 (SigLIP), 577 / 729 tokens, fwd only
//   * LLaMA decoder (A9): causal + key-padding mask, GQA, head_dim 128, S = 2048, fwd + bwd
// replacing torch SDPA as called by HF CLIP/DINOv2/Llama attention and timm Attention.
//
// Layout: q/k/v/o are addressed as (b, s, head, d) with arbitrary batch / row strides and heads packed along the
// row, so the fused QKV GEMM output [B*S, (nh + 2 nkv) * hd] is consumed in place (no permute / contiguous).
//
// Forward (one CTA = 128 query rows of one head; 192 threads):
//   warp 4  TMA producer : Q once, then K_j / V_j tiles through two mbarrier rings
//   warp 5  MMA issuer   : S_j = Q K_j^T (tcgen05, fp32 in TMEM, double buffered), O += P_j V_j
//   warps 0-3 softmax    : one thread per query row (TMEM lane == row: no shuffles), online softmax in the
//                          log2 domain with lazy rescale of the TMEM-resident O, P_j written as bf16 into a
//                          128B-swizzled smem tile that the PV MMA reads as its A operand.
// Backward (one CTA = 128 keys of one KV head; loops over the query heads of the GQA group and the query tiles):
//   S^T = K Q^T and dP^T = V dO^T in TMEM; threads (one per key row) form P^T (bf16, back into TMEM as the A
//   operand of dV += P^T dO) and dS^T (bf16, smem: K-major A of dK += dS^T Q and MN-major A of dQ_i = dS K);
//   dK / dV accumulate in TMEM over the whole loop, dQ partials are staged as fp32 tiles in the Q / dO stage the
//   iteration just retired and reduce-added into global memory by TMA (cp.reduce.async.bulk.tensor .add).
#include "common.cuh"
#include <cudaTypedefs.h>

namespace cb {

int make_tmap_bf16_4d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                      uint64_t s1, uint64_t s2, uint64_t s3, uint32_t box_rows);

constexpr float LOG2E = 1.4426950408889634f;

struct AttnParams {
  bf16* o;
  long long o_bs, o_ss;  // element strides of O (batch, row); heads packed at hd
  float* lse;            // [B, nh, Sq] log2-domain, may be null
  const uint8_t* kmask;  // [B, Skv] 1 = attend, may be null
  int B, nh, nkv, Sq, Skv, hd, causal;
  float scale_log2;      // softmax scale * log2(e)
};

template <int HDP>
struct FwdCfg {
  static constexpr int ATOMS = (HDP + 63) / 64;
  static constexpr int TILE = ATOMS * 16384;         // one [128 x hd] operand tile
  static constexpr int STAGES = (HDP <= 64) ? 3 : 2;
  static constexpr int P_BYTES = 32768;               // [128 x 128] bf16
  static constexpr int SMEM = TILE * (1 + 2 * STAGES) + P_BYTES + 1024 + 256;
  static constexpr int OCH = (HDP + 31) / 32;         // 32-column chunks of O
};

template <int HDP>
__global__ void __launch_bounds__(192, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, AttnParams p) {
  using Cfg = FwdCfg<HDP>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;
  const uint32_t sK = sQ + Cfg::TILE;
  const uint32_t sV = sK + STAGES * Cfg::TILE;
  const uint32_t sP = sV + STAGES * Cfg::TILE;
  const uint32_t bars = sP + Cfg::P_BYTES;
  const uint32_t q_full = bars;
  auto k_full = [&](int s) { return bars + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bars + 8u * (1 + STAGES + s); };
  auto v_full = [&](int s) { return bars + 8u * (1 + 2 * STAGES + s); };
  auto v_empty = [&](int s) { return bars + 8u * (1 + 3 * STAGES + s); };
  auto s_full = [&](int s) { return bars + 8u * (1 + 4 * STAGES + s); };
  const uint32_t p_full = bars + 8u * (3 + 4 * STAGES);
  const uint32_t pv_done = bars + 8u * (4 + 4 * STAGES);
  const uint32_t tmem_slot = bars + 8u * (5 + 4 * STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tiles = (p.Sq + 127) / 128;
  const int qt = q_tiles - 1 - (int)blockIdx.x;  // heavy (late) causal tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.nh / p.nkv);
  const int q0 = qt * 128;
  const int kv_tiles_all = (p.Skv + 127) / 128;
  // causal: keys <= query index (+ offset when Skv > Sq, e.g. a prefilled cache)
  const int coff = p.Skv - p.Sq;
  int n_tiles = kv_tiles_all;
  if (p.causal) {
    const int last_key = q0 + 127 + coff;
    n_tiles = min(kv_tiles_all, last_key / 128 + 1);
    if (n_tiles < 1) n_tiles = 1;
  }

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
    }
    mbar_init(s_full(0), 1);
    mbar_init(s_full(1), 1);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tS0 = tmem, tO = tmem + 256;

  if (warp == 4) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Cfg::TILE);
#pragma unroll
      for (int a = 0; a < Cfg::ATOMS; ++a) tma_load_4d(sQ + a * 16384, &tmQ, q_full, a * 64, h, q0, b);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(k_empty(st), ph ^ 1u);
        mbar_arrive_expect_tx(k_full(st), Cfg::TILE);
#pragma unroll
        for (int a = 0; a < Cfg::ATOMS; ++a)
          tma_load_4d(sK + st * Cfg::TILE + a * 16384, &tmK, k_full(st), a * 64, hk, j * 128, b);
        mbar_wait(v_empty(st), ph ^ 1u);
        mbar_arrive_expect_tx(v_full(st), Cfg::TILE);
#pragma unroll
        for (int a = 0; a < Cfg::ATOMS; ++a)
          tma_load_4d(sV + st * Cfg::TILE + a * 16384, &tmV, v_full(st), a * 64, hk, j * 128, b);
        if (++st == STAGES) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 5) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, HDP, 0, 1);
      auto issue_s = [&](int j, int st) {
        const uint32_t kb = sK + st * Cfg::TILE;
#pragma unroll
        for (int kk = 0; kk < HDP / 16; ++kk) {
          const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
          umma_ss(tS0 + (j & 1) * 128, make_smem_desc_sw128(sQ + off, 0, 1024),
                  make_smem_desc_sw128(kb + off, 0, 1024), idesc_qk, kk ? 1u : 0u);
        }
        umma_commit(k_empty(st));
        umma_commit(s_full(j & 1));
      };
      mbar_wait(q_full, 0);
      int kst = 0; uint32_t kph = 0;
      int vst = 0; uint32_t vph = 0;
      mbar_wait(k_full(0), 0);
      tc_fence_after();
      issue_s(0, 0);
      if (++kst == STAGES) { kst = 0; kph ^= 1u; }
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) {
          mbar_wait(k_full(kst), kph);
          tc_fence_after();
          issue_s(j + 1, kst);
          if (++kst == STAGES) { kst = 0; kph ^= 1u; }
        }
        mbar_wait(p_full, j & 1);
        mbar_wait(v_full(vst), vph);
        tc_fence_after();
        const uint32_t vb = sV + vst * Cfg::TILE;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t aoff = (kk >> 2) * 16384 + (kk & 3) * 32;
          umma_ss(tO, make_smem_desc_sw128(sP + aoff, 0, 1024),
                  make_smem_desc_sw128(vb + kk * 2048, 16384, 1024), idesc_pv, (j | kk) ? 1u : 0u);
        }
        umma_commit(v_empty(vst));
        umma_commit(pv_done);
        if (++vst == STAGES) { vst = 0; vph ^= 1u; }
      }
    }
  } else {
    // ================================ softmax warps ===============================
    const int row = warp * 32 + lane;  // TMEM lane == query row in the tile
    const int qi = q0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    float m_ref = 0.f, l = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int k0 = j * 128;
      int kmax = p.Skv - k0;                                // keys beyond Skv
      if (p.causal) kmax = min(kmax, qi + coff - k0 + 1);   // keys beyond the diagonal
      // key-padding mask of this tile as 4 x 32 bits, built cooperatively (lane l owns keys 4l..4l+3)
      uint32_t mw[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      if (p.kmask) {
        const uint8_t* km = p.kmask + (size_t)b * p.Skv + k0;
        uint32_t nib = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int kc = lane * 4 + e;
          if (k0 + kc < p.Skv && km[kc]) nib |= 1u << e;
        }
        uint32_t v = nib << (4 * (lane & 7));
        v |= __shfl_xor_sync(0xffffffffu, v, 1);
        v |= __shfl_xor_sync(0xffffffffu, v, 2);
        v |= __shfl_xor_sync(0xffffffffu, v, 4);
#pragma unroll
        for (int c = 0; c < 4; ++c) mw[c] = __shfl_sync(0xffffffffu, v, 8 * c);
      }
      mbar_wait(s_full(j & 1), (j >> 1) & 1);
      tc_fence_after();
      const uint32_t tS = tS0 + (j & 1) * 128 + lane_off;
      // one TMEM read of the whole 128-key row into registers
      uint32_t sr[128];
      tmem_ld32(tS, sr);
      tmem_ld32(tS + 32, sr + 32);
      tmem_ld32(tS + 64, sr + 64);
      tmem_ld32(tS + 96, sr + 96);
      tmem_ld_wait();
      if (kmax < 128 || p.kmask) {  // branch-free masking: masked scores become -inf
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int rem = kmax - c * 32;
          uint32_t bits = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
          bits &= mw[c];
#pragma unroll
          for (int e = 0; e < 32; ++e) sr[c * 32 + e] = ((bits >> e) & 1u) ? sr[c * 32 + e] : 0xff800000u;
        }
      }
      float t0 = -INFINITY, t1 = -INFINITY, t2 = -INFINITY, t3 = -INFINITY;
#pragma unroll
      for (int e = 0; e < 128; e += 4) {
        t0 = fmaxf(t0, __uint_as_float(sr[e]));
        t1 = fmaxf(t1, __uint_as_float(sr[e + 1]));
        t2 = fmaxf(t2, __uint_as_float(sr[e + 2]));
        t3 = fmaxf(t3, __uint_as_float(sr[e + 3]));
      }
      float tmax = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3)) * p.scale_log2;  // scale > 0 keeps ordering; -inf stays
      // lazy rescale (log2 domain): only move the reference max when it grows by > 8 (p <= 256)
      const bool grow = (j == 0) ? true : (tmax > m_ref + 8.f);
      if (__any_sync(0xffffffffu, grow)) {
        float m_new = (j == 0) ? tmax : fmaxf(m_ref, tmax);
        if (m_new == -INFINITY) m_new = 0.f;
        if (j > 0) {
          const float corr = fast_exp2(m_ref - m_new);
          l *= corr;
          mbar_wait(pv_done, (j - 1) & 1);  // O accumulation of tile j-1 retired
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < Cfg::OCH; ++c) {
            uint32_t r[32];
            tmem_ld32(tO + lane_off + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * corr);
            tmem_st32(tO + lane_off + c * 32, r);
          }
          tmem_st_wait();
        }
        m_ref = m_new;
      } else if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);  // P smem tile free again
      }
      // p = exp2(s*scale - m_ref) (masked: exp2(-inf) = 0), row sum, bf16 P tile in the K-major SWIZZLE_128B layout
      const float neg_m = -m_ref;
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int g = 0; g < 16; ++g) {  // 16-byte chunk index along the 128 keys
        float pv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) pv[e] = fast_exp2(fmaf(__uint_as_float(sr[g * 8 + e]), p.scale_log2, neg_m));
        l0 += pv[0] + pv[4];
        l1 += pv[1] + pv[5];
        l2 += pv[2] + pv[6];
        l3 += pv[3] + pv[7];
        const uint32_t addr = sP + (g >> 3) * 16384 + row * 128 + (((g & 7) ^ (row & 7)) << 4);
        const uint4 val = pack8(pv);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(val.x), "r"(val.y), "r"(val.z),
                     "r"(val.w) : "memory");
      }
      l += (l0 + l1) + (l2 + l3);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16 -> HBM
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv = l > 0.f ? 1.f / l : 0.f;
    bf16* op = p.o + (long long)b * p.o_bs + (long long)qi * p.o_ss + (long long)h * p.hd;
#pragma unroll 1
    for (int c = 0; c < Cfg::OCH; ++c) {
      uint32_t r[32];
      tmem_ld32(tO + lane_off + c * 32, r);
      tmem_ld_wait();
      if (qi < p.Sq) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int d0 = c * 32 + u * 8;
          if (d0 < p.hd) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[u * 8 + e]) * inv;
            *reinterpret_cast<uint4*>(op + d0) = pack8(f);
          }
        }
      }
    }
    if (p.lse && qi < p.Sq)
      p.lse[((long long)b * p.nh + h) * p.Sq + qi] = l > 0.f ? m_ref + log2f(l) : INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 5) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// delta[b, h, s] = sum_d dO[b,s,h,d] * O[b,s,h,d]   (fp32).  hd / 8 lanes (one 16-byte vector each) per (b, s, h), so a
// warp covers 32 / (hd / 8) heads: with one warp per head only hd / 8 of the 32 lanes had work (2.2 TB/s).
__global__ void attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, float* __restrict__ delta,
                                  int B, int S, int nh, int hd, long long o_bs, long long o_ss, long long do_bs,
                                  long long do_ss) {
  const int lph = hd >> 3;                       // lanes per head: 16 (hd 128) or 8 (hd 64)
  const int hpw = 32 / lph;                      // heads per warp
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * S * nh;
  const long long idx = w * hpw + lane / lph;    // (b, s, h) handled by this lane group
  const int sub = lane % lph;
  float acc = 0.f;
  if (idx < total) {
    const int h = (int)(idx % nh);
    const long long t = idx / nh;
    const int s = (int)(t % S);
    const int b = (int)(t / S);
    const bf16* op = o + b * o_bs + s * o_ss + (long long)h * hd + sub * 8;
    const bf16* gp = d_o + b * do_bs + s * do_ss + (long long)h * hd + sub * 8;
    float a[8], g[8];
    unpack8(ldg_nc(op), a);
    unpack8(ldg_nc(gp), g);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += a[e] * g[e];
  }
  for (int off = lph >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (idx < total && sub == 0) {
    const int h = (int)(idx % nh);
    const long long t = idx / nh;
    delta[((long long)(t / S) * nh + h) * S + (t % S)] = acc;
  }
}

struct AttnBwdParams {
  float* dq_acc;            // [B, Sq, nh, hd] fp32, zero-initialised by the caller; holds sum of dS K (unscaled)
  bf16* dk;                 // (b, s, kv head, d) with strides below
  bf16* dv;
  long long dk_bs, dk_ss, dv_bs, dv_ss;
  const float* lse;         // [B, nh, Sq] log2 domain
  const float* delta;       // [B, nh, Sq]
  const uint8_t* kmask;
  int B, nh, nkv, Sq, Skv, hd, causal;
  float scale_log2, scale;
};

template <int HDP>
struct BwdCfg {
  static constexpr int ATOMS = (HDP + 63) / 64;
  static constexpr int TILE = ATOMS * 16384;
  static constexpr int DS_BYTES = 32768;
  static constexpr int SMEM = TILE * 6 + DS_BYTES + 1024 + 1024 /*lse+delta*/ + 256;
  static constexpr int OCH = (HDP + 31) / 32;
};

template <int HDP>
__global__ void __launch_bounds__(320, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                const __grid_constant__ CUtensorMap tmdQ, AttnBwdParams p) {
  using Cfg = BwdCfg<HDP>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = base;
  const uint32_t sV = sK + Cfg::TILE;
  const uint32_t sQ = sV + Cfg::TILE;        // 2 stages
  const uint32_t sdO = sQ + 2 * Cfg::TILE;   // 2 stages
  const uint32_t sdS = sdO + 2 * Cfg::TILE;  // [128 kv x 128 q] bf16, K-major over q
  const uint32_t sStat = sdS + Cfg::DS_BYTES;  // float lse[128], delta[128]
  const uint32_t bars = sStat + 1024;
  const uint32_t kv_full = bars;
  auto qd_full = [&](int s) { return bars + 8u * (1 + s); };
  auto qd_empty = [&](int s) { return bars + 8u * (3 + s); };
  const uint32_t sdp_full = bars + 8u * 5;   // S^T and dP^T ready in TMEM
  const uint32_t pds_full = bars + 8u * 6;   // P^T (TMEM) and dS^T (smem) written by the 128 threads
  const uint32_t dq_full = bars + 8u * 7;    // dQ partial ready in TMEM (also: dV/dK MMAs of this iteration retired)
  const uint32_t dq_done = bars + 8u * 8;    // dQ partial read out by the 128 threads
  const uint32_t dq_staged = bars + 8u * 9;  // fp32 dQ partial staged in this iteration's (retired) Q / dO stage
  const uint32_t tmem_slot = bars + 8u * 10;
  float* stat = reinterpret_cast<float*>(smem_raw + (sStat - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
  const int G = p.nh / p.nkv;
  const int k0 = jt * 128;
  const int q_tiles = (p.Sq + 127) / 128;
  const int coff = p.Skv - p.Sq;
  // causal: query i attends key k iff k <= i + coff  ->  first query tile that can see key k0
  int i_begin = 0;
  if (p.causal) {
    const int first_q = k0 - coff;
    i_begin = first_q > 0 ? first_q / 128 : 0;
  }
  const int n_i = q_tiles > i_begin ? q_tiles - i_begin : 0;
  const int n_iter = n_i * G;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmdQ);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(qd_full(s), 1);
      mbar_init(qd_empty(s), 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 256);
    mbar_init(dq_full, 1);
    mbar_init(dq_done, 256);
    mbar_init(dq_staged, 256);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tA = tmem;         // S^T fp32 [128 x 128]  -> P^T bf16 packed in the first 64 columns
  const uint32_t tB = tmem + 128;   // dP^T fp32 [128 x 128] -> dQ partial [128 q x HDP]
  const uint32_t tdV = tmem + 256;  // [128 kv x HDP]
  const uint32_t tdK = tmem + 384;  // [128 kv x HDP]

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * Cfg::TILE);
#pragma unroll
      for (int a = 0; a < Cfg::ATOMS; ++a) {
        tma_load_4d(sK + a * 16384, &tmK, kv_full, a * 64, g, k0, b);
        tma_load_4d(sV + a * 16384, &tmV, kv_full, a * 64, g, k0, b);
      }
      auto load_qd = [&](int it) {
        const int st = it & 1;
        const int h = g * G + it / n_i;
        const int qi0 = (i_begin + it % n_i) * 128;
        mbar_arrive_expect_tx(qd_full(st), 2 * Cfg::TILE);
#pragma unroll
        for (int a = 0; a < Cfg::ATOMS; ++a) {
          tma_load_4d(sQ + st * Cfg::TILE + a * 16384, &tmQ, qd_full(st), a * 64, h, qi0, b);
          tma_load_4d(sdO + st * Cfg::TILE + a * 16384, &tmdO, qd_full(st), a * 64, h, qi0, b);
        }
      };
      if (n_iter > 0) load_qd(0);
      if (n_iter > 1) load_qd(1);
      for (int it = 0; it < n_iter; ++it) {
        // The compute warps stage the fp32 dQ partial of iteration `it` in the Q / dO stage that iteration just retired
        // (its MMAs completed before dq_full).  This thread reduce-adds it into global memory with TMA — per-thread
        // `red.global` on 32 different rows per instruction cost 32 % of the kernel — waits until the engine has READ the
        // tile, and only then refills the stage with the operands of iteration it + 2 (still a full iteration ahead).
        const int st = it & 1;
        const int h = g * G + it / n_i;
        const int qi0 = (i_begin + it % n_i) * 128;
        mbar_wait(dq_staged, it & 1);
#pragma unroll
        for (int c = 0; c < Cfg::OCH; ++c) {
          const uint32_t src = (c * 16384 < Cfg::TILE) ? sQ + st * Cfg::TILE + c * 16384
                                                       : sdO + st * Cfg::TILE + (c * 16384 - Cfg::TILE);
          tma_reduce_add_4d(&tmdQ, src, c * 32, h, qi0, b);
        }
        bulk_commit_group();
        bulk_wait_group_read0();
        if (it + 2 < n_iter) load_qd(it + 2);
      }
      bulk_wait_group0();
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_dv = make_idesc_bf16(128, HDP, 0, 1);   // A: P^T from TMEM, B: dO (MN-major)
      constexpr uint32_t idesc_dk = make_idesc_bf16(128, HDP, 0, 1);   // A: dS^T smem K-major, B: Q (MN-major)
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, HDP, 1, 1);   // A: dS (MN-major view), B: K (MN-major)
      mbar_wait(kv_full, 0);
      int st = 0;
      uint32_t ph = 0;
      for (int it = 0; it < n_iter; ++it) {
        mbar_wait(qd_full(st), ph);
        if (it > 0) mbar_wait(dq_done, (it - 1) & 1);  // region B (dQ partial of it-1) has been read out
        tc_fence_after();
        const uint32_t qb = sQ + st * Cfg::TILE, dob = sdO + st * Cfg::TILE;
#pragma unroll
        for (int kk = 0; kk < HDP / 16; ++kk) {
          const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
          umma_ss(tA, make_smem_desc_sw128(sK + off, 0, 1024), make_smem_desc_sw128(qb + off, 0, 1024), idesc_s,
                  kk ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < HDP / 16; ++kk) {
          const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
          umma_ss(tB, make_smem_desc_sw128(sV + off, 0, 1024), make_smem_desc_sw128(dob + off, 0, 1024), idesc_s,
                  kk ? 1u : 0u);
        }
        umma_commit(sdp_full);
        mbar_wait(pds_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // dV += P^T dO   (K = 128 queries)
          umma_ts(tdV, tA + kk * 8, make_smem_desc_sw128(dob + kk * 2048, 16384, 1024), idesc_dv,
                  (it | kk) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // dK += dS^T Q
          const uint32_t aoff = (kk >> 2) * 16384 + (kk & 3) * 32;
          umma_ss(tdK, make_smem_desc_sw128(sdS + aoff, 0, 1024), make_smem_desc_sw128(qb + kk * 2048, 16384, 1024),
                  idesc_dk, (it | kk) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // dQ_i = dS K   (K = 128 keys)
          umma_ss(tB, make_smem_desc_sw128(sdS + kk * 2048, 16384, 1024),
                  make_smem_desc_sw128(sK + kk * 2048, 16384, 1024), idesc_dq, kk ? 1u : 0u);
        }
        umma_commit(dq_full);  // (the Q / dO stage is handed back through dq_staged: see the producer)
        if (++st == 2) { st = 0; ph ^= 1u; }
      }
    }
  } else {
    // 8 compute warps (0-3 and 6-9): two warps per TMEM lane quarter, each owning half of the columns of every
    // row — the per-element work has no row reductions, so the split needs no exchange and halves the serial thread phase
    const int row = (warp & 3) * 32 + lane;  // key row of this CTA's tile (phase 1) / query row of the tile (phase 2)
    const int ch = warp >= 6 ? 1 : 0;        // column half
    const int kidx = k0 + row;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const bool key_ok = kidx < p.Skv && (!p.kmask || p.kmask[(size_t)b * p.Skv + kidx]);
    for (int it = 0; it < n_iter; ++it) {
      const int h = g * G + it / n_i;
      const int qi0 = (i_begin + it % n_i) * 128;
      // stage lse / delta of the 128 queries of this tile (previous readers are past pds arrive of it-1)
      {
        const int qi = qi0 + row;
        const long long so = ((long long)b * p.nh + h) * p.Sq + qi;
        stat[row] = qi < p.Sq ? p.lse[so] : INFINITY;
        stat[128 + row] = qi < p.Sq ? p.delta[so] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(sdp_full, it & 1);
      tc_fence_after();
      // queries visible to this key: qi + coff >= kidx, qi < Sq
      int qmin = 0;
      if (p.causal) qmin = kidx - coff - qi0;
      const int qmax = p.Sq - qi0;
      // P^T is written in place over S^T (columns [0, 64) of region A), i.e. over columns the OTHER half's warp still has
      // to read: pull both S^T chunks of this warp into registers and meet at a barrier before anyone stores P^T
      uint32_t rs2[64];
      tmem_ld32(tA + lane_off + (2 * ch) * 32, rs2);
      tmem_ld32(tA + lane_off + (2 * ch + 1) * 32, rs2 + 32);
      tmem_ld_wait();
      asm volatile("bar.sync 2, 256;" ::: "memory");
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * ch + cc;
        const uint32_t* rs = rs2 + cc * 32;
        uint32_t rd[32];
        tmem_ld32(tB + lane_off + c * 32, rd);
        tmem_ld_wait();
        // branch-free visibility mask for the 32 queries of this chunk
        const int lo = max(qmin - c * 32, 0), hi = min(qmax - c * 32, 32);
        uint32_t bits = 0u;
        if (key_ok && hi > lo) bits = (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
        float pv[32], ds[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int qc = c * 32 + e;
          float pe = fast_exp2(fmaf(__uint_as_float(rs[e]), p.scale_log2, -stat[qc]));
          pe = ((bits >> e) & 1u) ? pe : 0.f;
          pv[e] = pe;
          ds[e] = pe * (__uint_as_float(rd[e]) - stat[128 + qc]);
        }
        // P^T -> TMEM (bf16 pairs, 16 columns per 32 queries), in place over the S^T columns already consumed
        uint32_t pk[32];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          __nv_bfloat162 t = __floats2bfloat162_rn(pv[2 * e], pv[2 * e + 1]);
          pk[e] = *reinterpret_cast<uint32_t*>(&t);
        }
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
            "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
            :
            : "r"(tA + lane_off + c * 16), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]),
              "r"(pk[6]), "r"(pk[7]), "r"(pk[8]), "r"(pk[9]), "r"(pk[10]), "r"(pk[11]), "r"(pk[12]), "r"(pk[13]),
              "r"(pk[14]), "r"(pk[15])
            : "memory");
        // dS^T -> smem, K-major over the 128 queries (row = key)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int gq = c * 4 + u;
          const uint32_t addr = sdS + (gq >> 3) * 16384 + row * 128 + (((gq & 7) ^ (row & 7)) << 4);
          const uint4 val = pack8(ds + u * 8);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(val.x), "r"(val.y), "r"(val.z),
                       "r"(val.w) : "memory");
        }
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_full);
      // ---- dQ partial: thread == query row of tile i
      mbar_wait(dq_full, it & 1);
      tc_fence_after();
      {
        // dQ partial: TMEM -> registers -> fp32 tile in this iteration's retired Q / dO stage (128B-swizzled boxes of
        // [128 rows][32 floats], matching the reduce tensor map); the producer thread issues the TMA reduce-add
        constexpr int HC = Cfg::OCH / 2;  // 32-column chunks of the dQ row owned by this warp
        uint32_t r[HC * 32];
#pragma unroll
        for (int c = 0; c < HC; ++c) tmem_ld32(tB + lane_off + (ch * HC + c) * 32, r + c * 32);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(dq_done);  // region B is free: the next iteration's S^T / dP^T MMAs overlap with the readout below
        const int st = it & 1;
#pragma unroll
        for (int rd = 0; rd < HC; ++rd) {
          const int c = ch * HC + rd;
          const uint32_t box = (c * 16384 < Cfg::TILE) ? sQ + st * Cfg::TILE + c * 16384
                                                       : sdO + st * Cfg::TILE + (c * 16384 - Cfg::TILE);
          const uint32_t dst = box + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + ((j ^ (row & 7)) << 4)),
                         "r"(r[rd * 32 + j * 4]), "r"(r[rd * 32 + j * 4 + 1]), "r"(r[rd * 32 + j * 4 + 2]),
                         "r"(r[rd * 32 + j * 4 + 3])
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(dq_staged);
      }
    }
    // ---- epilogue: dK (x softmax scale), dV -> bf16
    if (n_iter > 0) {
      // dq_full of the last iteration also covers the dV / dK MMAs issued before it
      bf16* dkp = p.dk + (long long)b * p.dk_bs + (long long)kidx * p.dk_ss + (long long)g * p.hd;
      bf16* dvp = p.dv + (long long)b * p.dv_bs + (long long)kidx * p.dv_ss + (long long)g * p.hd;
#pragma unroll 1
      for (int c = ch * (Cfg::OCH / 2); c < (ch + 1) * (Cfg::OCH / 2); ++c) {
        uint32_t r1[32], r2[32];
        tmem_ld32(tdK + lane_off + c * 32, r1);
        tmem_ld32(tdV + lane_off + c * 32, r2);
        tmem_ld_wait();
        if (kidx < p.Skv) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int d0 = c * 32 + u * 8;
            if (d0 < p.hd) {
              float f1[8], f2[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                f1[e] = __uint_as_float(r1[u * 8 + e]) * p.scale;
                f2[e] = __uint_as_float(r2[u * 8 + e]);
              }
              *reinterpret_cast<uint4*>(dkp + d0) = pack8(f1);
              *reinterpret_cast<uint4*>(dvp + d0) = pack8(f2);
            }
          }
        }
      }
    } else if (kidx < p.Skv) {
      bf16* dkp = p.dk + (long long)b * p.dk_bs + (long long)kidx * p.dk_ss + (long long)g * p.hd;
      bf16* dvp = p.dv + (long long)b * p.dv_bs + (long long)kidx * p.dv_ss + (long long)g * p.hd;
      const uint4 z = make_uint4(0, 0, 0, 0);
      for (int d0 = ch * (p.hd / 2); d0 < (ch + 1) * (p.hd / 2); d0 += 8) {
        *reinterpret_cast<uint4*>(dkp + d0) = z;
        *reinterpret_cast<uint4*>(dvp + d0) = z;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 5) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// dims (d0 = head_dim contiguous, d1 = heads, d2 = sequence, d3 = batch); strides in elements
int make_tmap_bf16_4d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                      uint64_t s1, uint64_t s2, uint64_t s3, uint32_t box_rows) {
  auto fn = encode_fn();
  if (!fn) return set_error(CB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || (s1 % 8) || (s2 % 8) || (d3 > 1 && (s3 % 8)))
    return set_error(CB_ERR_INVALID, "attention: operands must be 16-byte aligned with strides %% 8 == 0");
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  cuuint64_t strides[3] = {s1 * 2, s2 * 2, (d3 > 1 ? s3 : s2 * d2) * 2};
  cuuint32_t box[4] = {64, 1, box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_ERROR_INVALID_CONTEXT) {  // see gemm.cu: bind the primary context on lazily-initialised threads
    cudaFree(0);
    r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) return set_error(CB_ERR_CUDA, "cuTensorMapEncodeTiled(4d) failed (%d)", (int)r);
  return CB_OK;
}

static int hd_padded(int hd) { return (hd + 15) / 16 * 16; }

template <int HDP>
static int launch_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                      cudaStream_t st) {
  using Cfg = FwdCfg<HDP>;
  auto kern = attn_fwd_kernel<HDP>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return set_error(CB_ERR_CUDA, "attn fwd smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid((p.Sq + 127) / 128, p.nh, p.B);
  kern<<<grid, 192, Cfg::SMEM, st>>>(tq, tk, tv, p);
  CB_CUDA_LAUNCH_CHECK("attn_fwd");
  return CB_OK;
}

int attn_fwd_launch(const void* q, const void* k, const void* v, void* o, float* lse, const void* kmask, int B,
                    int nh, int nkv, int Sq, int Skv, int hd, long long q_bs, long long q_ss, long long k_bs,
                    long long k_ss, long long v_bs, long long v_ss, long long o_bs, long long o_ss, float scale,
                    int causal, cudaStream_t st) {
  CB_CHECK_ARG(B > 0 && nh > 0 && nkv > 0 && Sq > 0 && Skv > 0, "attention: empty problem");
  CB_CHECK_ARG(nh % nkv == 0, "attention: nh=%d not a multiple of nkv=%d", nh, nkv);
  CB_CHECK_ARG(hd % 8 == 0 && hd <= 128, "attention: head_dim=%d must be a multiple of 8 and <= 128", hd);
  CB_CHECK_ARG(o_ss % 8 == 0 && o_bs % 8 == 0 && !(reinterpret_cast<uintptr_t>(o) & 15u),
               "attention: output must be 16-byte aligned with strides %% 8 == 0");
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap_bf16_4d(&tq, q, hd, nh, Sq, B, hd, q_ss, q_bs, 128))) return rc;
  if ((rc = make_tmap_bf16_4d(&tk, k, hd, nkv, Skv, B, hd, k_ss, k_bs, 128))) return rc;
  if ((rc = make_tmap_bf16_4d(&tv, v, hd, nkv, Skv, B, hd, v_ss, v_bs, 128))) return rc;
  AttnParams p;
  p.o = (bf16*)o; p.o_bs = o_bs; p.o_ss = o_ss; p.lse = lse; p.kmask = (const uint8_t*)kmask;
  p.B = B; p.nh = nh; p.nkv = nkv; p.Sq = Sq; p.Skv = Skv; p.hd = hd; p.causal = causal;
  p.scale_log2 = scale * LOG2E;
  const int hdp = hd_padded(hd);
  if (hdp <= 64) return launch_fwd<64>(tq, tk, tv, p, st);
  if (hdp <= 80) return launch_fwd<80>(tq, tk, tv, p, st);
  if (hdp <= 96) return launch_fwd<96>(tq, tk, tv, p, st);
  return launch_fwd<128>(tq, tk, tv, p, st);
}

template <int HDP>
static int launch_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                      const CUtensorMap& tdq, const AttnBwdParams& p, cudaStream_t st) {
  using Cfg = BwdCfg<HDP>;
  auto kern = attn_bwd_kernel<HDP>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return set_error(CB_ERR_CUDA, "attn bwd smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid((p.Skv + 127) / 128, p.nkv, p.B);
  kern<<<grid, 320, Cfg::SMEM, st>>>(tq, tk, tv, tdo, tdq, p);
  CB_CUDA_LAUNCH_CHECK("attn_bwd");
  return CB_OK;
}

int attn_bwd_launch(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                    float* delta, float* dq_acc, void* dk, void* dv, const void* kmask, int B, int nh, int nkv,
                    int Sq, int Skv, int hd, long long q_bs, long long q_ss, long long k_bs, long long k_ss,
                    long long v_bs, long long v_ss, long long o_bs, long long o_ss, long long do_bs, long long do_ss,
                    long long dk_bs, long long dk_ss, long long dv_bs, long long dv_ss, float scale, int causal,
                    cudaStream_t st) {
  CB_CHECK_ARG(B > 0 && nh > 0 && nkv > 0 && Sq > 0 && Skv > 0, "attention bwd: empty problem");
  CB_CHECK_ARG(nh % nkv == 0, "attention bwd: nh=%d not a multiple of nkv=%d", nh, nkv);
  CB_CHECK_ARG(hd == 64 || hd == 128, "attention bwd: head_dim=%d unsupported (64 or 128)", hd);
  CB_CHECK_ARG(lse && delta && dq_acc, "attention bwd: lse / delta / dq_acc buffers are required");
  {
    const long long warps = ((long long)B * Sq * nh + (256 / hd) - 1) / (256 / hd);   // 32 / (hd / 8) heads per warp
    attn_delta_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>((const bf16*)o, (const bf16*)d_o, delta, B,
                                                                           Sq, nh, hd, o_bs, o_ss, do_bs, do_ss);
    CB_CUDA_LAUNCH_CHECK("attn_delta");
  }
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = make_tmap_bf16_4d(&tq, q, hd, nh, Sq, B, hd, q_ss, q_bs, 128))) return rc;
  if ((rc = make_tmap_bf16_4d(&tk, k, hd, nkv, Skv, B, hd, k_ss, k_bs, 128))) return rc;
  if ((rc = make_tmap_bf16_4d(&tv, v, hd, nkv, Skv, B, hd, v_ss, v_bs, 128))) return rc;
  if ((rc = make_tmap_bf16_4d(&tdo, d_o, hd, nh, Sq, B, hd, do_ss, do_bs, 128))) return rc;
  AttnBwdParams p;
  p.dq_acc = dq_acc; p.dk = (bf16*)dk; p.dv = (bf16*)dv;
  p.dk_bs = dk_bs; p.dk_ss = dk_ss; p.dv_bs = dv_bs; p.dv_ss = dv_ss;
  p.lse = lse; p.delta = delta; p.kmask = (const uint8_t*)kmask;
  p.B = B; p.nh = nh; p.nkv = nkv; p.Sq = Sq; p.Skv = Skv; p.hd = hd; p.causal = causal;
  p.scale_log2 = scale * LOG2E; p.scale = scale;
  CUtensorMap tdq;  // fp32 dQ accumulator [B, Sq, nh, hd]: reduce-add target, box = 32 floats x 128 query rows
  {
    auto fn = encode_fn();
    if (!fn) return set_error(CB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    CB_CHECK_ARG((reinterpret_cast<uintptr_t>(dq_acc) & 15u) == 0, "attention bwd: dq_acc must be 16-byte aligned");
    cuuint64_t dims[4] = {(cuuint64_t)hd, (cuuint64_t)nh, (cuuint64_t)Sq, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)hd * 4, (cuuint64_t)nh * hd * 4, (cuuint64_t)Sq * nh * hd * 4};
    cuuint32_t box[4] = {32, 1, 128, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = CUDA_SUCCESS;
    for (int attempt = 0; attempt < 2; ++attempt) {
      r = fn(&tdq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dq_acc, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_ERROR_INVALID_CONTEXT) break;
      cudaFree(0);
    }
    if (r != CUDA_SUCCESS) return set_error(CB_ERR_CUDA, "cuTensorMapEncodeTiled(dq) failed (%d)", (int)r);
  }
  if (hd == 64) return launch_bwd<64>(tq, tk, tv, tdo, tdq, p, st);
  return launch_bwd<128>(tq, tk, tv, tdo, tdq, p, st);
}

}  // namespace cb

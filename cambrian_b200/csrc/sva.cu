// cambrian_b200 — fused Spatial-Vision-Aggregator window attention.
//
// Reference semantics (cambrian/model/vision_sampler.py:177-234 MultiKVCrossAttention.forward,
// with the window gather of cambrian/model/cambrian_arch.py:271-287): every query n = (b, qy, qx)
// of the q_side x q_side query grid attends, with ONE joint softmax over 16 heads x 64, to the
// r_i x r_i window of tower i's (already projected) key/value grid that lies under it:
//     keys_i(n) = { K_i[b, qy*r_i + dy, qx*r_i + dx] : dy, dx < r_i },   mask_i[n, dy*r_i+dx]
// The reference materialises the windows with view/permute/contiguous and concatenates the
// towers; here the gather is index arithmetic inside the kernel, so K/V are read exactly once
// from their natural [B, side_i*side_i, 1024] layout and nothing is re-laid-out in HBM.
//
// The kernel is HBM-bound (q_len = 1, <= 19 keys per query): one warp per query,
// 16-byte coalesced loads (a warp covers a 2 KB row in 4 x 512 B requests), 8-lane shuffle
// reductions for the per-head dot products, online softmax, fp32 accumulation.
// Because windows partition each tower grid, every K/V row belongs to exactly one query, so the
// backward writes dK/dV without atomics and is deterministic.
#include "common.cuh"

namespace cb {

constexpr int SVA_MAX_TOWERS = 8;
constexpr int SVA_HIDDEN = 1024;  // 16 heads x 64 (vision_sampler.py:250 num_heads = 16)
constexpr float SVA_SCALE_LOG2E = 0.125f * 1.4426950408889634f;

struct SvaArgs {
  const bf16* k[SVA_MAX_TOWERS];
  const bf16* v[SVA_MAX_TOWERS];
  const uint8_t* mask[SVA_MAX_TOWERS];  // [N, r*r] bool or null
  bf16* dk[SVA_MAX_TOWERS];
  bf16* dv[SVA_MAX_TOWERS];
  int r[SVA_MAX_TOWERS];
  int num_towers;
  int q_side;
  int n_queries;  // B * q_side * q_side
  int windowed;   // 1: K/V are window-rearranged [N, r*r, C] (reference layout); 0: natural [B, side*side, C]
};

__device__ __forceinline__ float reduce8(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// lane handles channels {j*256 + lane*8 .. +7 : j < 4}; head(j) = 4*j + lane/8
__global__ void __launch_bounds__(128)
sva_window_attn_fwd(const bf16* __restrict__ Q, bf16* __restrict__ O, float* __restrict__ LSE, SvaArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 4 + warp;
  if (n >= a.n_queries) return;
  const int qn = a.q_side * a.q_side;
  const int b = n / qn, qi = n - b * qn;
  const int qy = qi / a.q_side, qx = qi - qy * a.q_side;

  float q[32];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(Q + (size_t)n * SVA_HIDDEN) + lane;
#pragma unroll
    for (int j = 0; j < 4; ++j) unpack8(ldg_nc(qp + j * 32), q + j * 8);
  }
  float m[4], l[4], acc[32];
#pragma unroll
  for (int j = 0; j < 4; ++j) { m[j] = -INFINITY; l[j] = 0.f; }
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;

  for (int t = 0; t < a.num_towers; ++t) {
    const int r = a.r[t];
    const int side = r * a.q_side;
    const size_t base = (size_t)b * side * side;
    const uint8_t* mk = a.mask[t] ? a.mask[t] + (size_t)n * r * r : nullptr;
    for (int w = 0; w < r * r; ++w) {
      if (mk && !mk[w]) continue;  // warp-uniform
      const int dy = w / r, dx = w - dy * r;
      const size_t row = a.windowed ? (size_t)n * r * r + w : base + (size_t)(qy * r + dy) * side + (qx * r + dx);
      const uint4* kp = reinterpret_cast<const uint4*>(a.k[t] + row * SVA_HIDDEN) + lane;
      const uint4* vp = reinterpret_cast<const uint4*>(a.v[t] + row * SVA_HIDDEN) + lane;
      uint4 kr[4], vr[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) kr[j] = ldg_nc(kp + j * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) vr[j] = ldg_nc(vp + j * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float kf[8], vf[8];
        unpack8(kr[j], kf);
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) s += q[j * 8 + e] * kf[e];
        s = reduce8(s) * SVA_SCALE_LOG2E;
        const float mn = fmaxf(m[j], s);
        const float corr = exp2f(m[j] - mn);  // m = -inf on the first key -> 0
        const float p = exp2f(s - mn);
        l[j] = l[j] * corr + p;
        m[j] = mn;
        unpack8(vr[j], vf);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[j * 8 + e] = acc[j * 8 + e] * corr + p * vf[e];
      }
    }
  }
  uint4* op = reinterpret_cast<uint4*>(O + (size_t)n * SVA_HIDDEN) + lane;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float inv = l[j] > 0.f ? 1.f / l[j] : 0.f;
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = acc[j * 8 + e] * inv;
    op[j * 32] = pack8(o);
    if (LSE && (lane & 7) == 0)
      LSE[(size_t)n * 16 + 4 * j + (lane >> 3)] = l[j] > 0.f ? m[j] + log2f(l[j]) : INFINITY;
  }
}

__global__ void __launch_bounds__(128)
sva_window_attn_bwd(const bf16* __restrict__ Q, const bf16* __restrict__ O, const bf16* __restrict__ dO,
                    const float* __restrict__ LSE, bf16* __restrict__ dQ, SvaArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 4 + warp;
  if (n >= a.n_queries) return;
  const int qn = a.q_side * a.q_side;
  const int b = n / qn, qi = n - b * qn;
  const int qy = qi / a.q_side, qx = qi - qy * a.q_side;

  float q[32], go[32], dq[32], delta[4], lse[4];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(Q + (size_t)n * SVA_HIDDEN) + lane;
    const uint4* gp = reinterpret_cast<const uint4*>(dO + (size_t)n * SVA_HIDDEN) + lane;
    const uint4* op = reinterpret_cast<const uint4*>(O + (size_t)n * SVA_HIDDEN) + lane;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float of[8];
      unpack8(ldg_nc(qp + j * 32), q + j * 8);
      unpack8(ldg_nc(gp + j * 32), go + j * 8);
      unpack8(ldg_nc(op + j * 32), of);
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) d += go[j * 8 + e] * of[e];
      delta[j] = reduce8(d);
      lse[j] = LSE[(size_t)n * 16 + 4 * j + (lane >> 3)];
    }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) dq[i] = 0.f;

  for (int t = 0; t < a.num_towers; ++t) {
    const int r = a.r[t];
    const int side = r * a.q_side;
    const size_t base = (size_t)b * side * side;
    const uint8_t* mk = a.mask[t] ? a.mask[t] + (size_t)n * r * r : nullptr;
    for (int w = 0; w < r * r; ++w) {
      const int dy = w / r, dx = w - dy * r;
      const size_t row = a.windowed ? (size_t)n * r * r + w : base + (size_t)(qy * r + dy) * side + (qx * r + dx);
      uint4* dkp = reinterpret_cast<uint4*>(a.dk[t] + row * SVA_HIDDEN) + lane;
      uint4* dvp = reinterpret_cast<uint4*>(a.dv[t] + row * SVA_HIDDEN) + lane;
      if (mk && !mk[w]) {  // masked key: no gradient, but the rows must still be defined
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) { dkp[j * 32] = z; dvp[j * 32] = z; }
        continue;
      }
      const uint4* kp = reinterpret_cast<const uint4*>(a.k[t] + row * SVA_HIDDEN) + lane;
      const uint4* vp = reinterpret_cast<const uint4*>(a.v[t] + row * SVA_HIDDEN) + lane;
      uint4 kr[4], vr[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) kr[j] = ldg_nc(kp + j * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) vr[j] = ldg_nc(vp + j * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float kf[8], vf[8];
        unpack8(kr[j], kf);
        unpack8(vr[j], vf);
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          s += q[j * 8 + e] * kf[e];
          dp += go[j * 8 + e] * vf[e];
        }
        s = reduce8(s) * SVA_SCALE_LOG2E;
        dp = reduce8(dp);
        const float p = exp2f(s - lse[j]);
        const float ds = p * (dp - delta[j]) * 0.125f;
        float dkf[8], dvf[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          dvf[e] = p * go[j * 8 + e];
          dkf[e] = ds * q[j * 8 + e];
          dq[j * 8 + e] += ds * kf[e];
        }
        dkp[j * 32] = pack8(dkf);
        dvp[j * 32] = pack8(dvf);
      }
    }
  }
  uint4* dqp = reinterpret_cast<uint4*>(dQ + (size_t)n * SVA_HIDDEN) + lane;
#pragma unroll
  for (int j = 0; j < 4; ++j) dqp[j * 32] = pack8(dq + j * 8);
}

static int fill_args(SvaArgs& a, int num_towers, const void* const* k, const void* const* v,
                     const void* const* mask, const int* r, int batch, int q_side, int windowed) {
  CB_CHECK_ARG(num_towers >= 1 && num_towers <= SVA_MAX_TOWERS, "sva: num_towers=%d out of [1,%d]", num_towers,
               SVA_MAX_TOWERS);
  CB_CHECK_ARG(batch > 0 && q_side > 0, "sva: empty query grid");
  a.num_towers = num_towers;
  a.q_side = q_side;
  a.windowed = windowed;
  a.n_queries = batch * q_side * q_side;
  for (int t = 0; t < num_towers; ++t) {
    CB_CHECK_ARG(r[t] >= 1, "sva: window side r[%d]=%d must be >= 1", t, r[t]);
    CB_CHECK_ARG(k[t] && v[t], "sva: null K/V for tower %d", t);
    a.k[t] = static_cast<const bf16*>(k[t]);
    a.v[t] = static_cast<const bf16*>(v[t]);
    a.mask[t] = mask ? static_cast<const uint8_t*>(mask[t]) : nullptr;
    a.r[t] = r[t];
    a.dk[t] = nullptr;
    a.dv[t] = nullptr;
  }
  return CB_OK;
}

int sva_window_attn_fwd_launch(const void* q, void* out, float* lse, int num_towers, const void* const* k,
                               const void* const* v, const void* const* mask, const int* r, int batch,
                               int q_side, int hidden, int windowed, cudaStream_t stream) {
  CB_CHECK_ARG(hidden == SVA_HIDDEN, "sva: hidden=%d unsupported (16 heads x 64 = 1024 only)", hidden);
  SvaArgs a;
  int rc = fill_args(a, num_towers, k, v, mask, r, batch, q_side, windowed);
  if (rc) return rc;
  const int grid = (a.n_queries + 3) / 4;
  sva_window_attn_fwd<<<grid, 128, 0, stream>>>(static_cast<const bf16*>(q), static_cast<bf16*>(out), lse, a);
  CB_CUDA_LAUNCH_CHECK("sva_window_attn_fwd");
  return CB_OK;
}

int sva_window_attn_bwd_launch(const void* q, const void* out, const void* dout, const float* lse, void* dq,
                               int num_towers, const void* const* k, const void* const* v,
                               const void* const* mask, void* const* dk, void* const* dv, const int* r,
                               int batch, int q_side, int hidden, int windowed, cudaStream_t stream) {
  CB_CHECK_ARG(hidden == SVA_HIDDEN, "sva: hidden=%d unsupported (16 heads x 64 = 1024 only)", hidden);
  CB_CHECK_ARG(lse != nullptr, "sva bwd: LSE from the forward pass is required");
  SvaArgs a;
  int rc = fill_args(a, num_towers, k, v, mask, r, batch, q_side, windowed);
  if (rc) return rc;
  for (int t = 0; t < num_towers; ++t) {
    CB_CHECK_ARG(dk[t] && dv[t], "sva bwd: null dK/dV for tower %d", t);
    a.dk[t] = static_cast<bf16*>(dk[t]);
    a.dv[t] = static_cast<bf16*>(dv[t]);
  }
  const int grid = (a.n_queries + 3) / 4;
  sva_window_attn_bwd<<<grid, 128, 0, stream>>>(static_cast<const bf16*>(q), static_cast<const bf16*>(out),
                                                static_cast<const bf16*>(dout), lse, static_cast<bf16*>(dq), a);
  CB_CUDA_LAUNCH_CHECK("sva_window_attn_bwd");
  return CB_OK;
}

}  // namespace cb

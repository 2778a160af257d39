// cambrian_b200 — kernels of the `sep` Spatial-Vision-Aggregator layer and the adjoint of the token-grid resize.
//
//  * tower_combine_fwd/bwd — VisionAggregationLayer.forward (cambrian/model/vision_sampler.py:368-398): the per-tower
//    aggregates are mixed with a per-query softmax over towers and added to the query stream,
//        w = softmax(weight_mlp(cat(q, ctx)))            [N, T]                      (:369-371)
//        out = q_in + sum_t w[:, t] * agg_t              [N, C]                      (:396-398)
//    One warp per query row, fp32 math, 16-byte coalesced accesses; the softmax weights are recomputed from the logits in
//    the backward pass (T <= 8 values per row) instead of being stored.
//  * bilinear_bwd — adjoint of bilinear_kernel (elementwise.cu; F.interpolate bilinear, align_corners=False, used for
//    the query-grid resize of cambrian_arch.py:394-401 and the tower token-grid interpolation).  Gather form: one thread
//    per (input pixel, 8 channels) sums the output pixels whose two taps per axis touch it — no atomics, deterministic.
#include "common.cuh"

namespace cb {

constexpr int AGG_MAX_TOWERS = 8;

struct AggPtrs {
  const bf16* p[AGG_MAX_TOWERS];
};
struct AggOutPtrs {
  bf16* p[AGG_MAX_TOWERS];
};

static inline unsigned agg_grid_for(long long blocks) {
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

__device__ __forceinline__ void softmax_row(const bf16* __restrict__ lg, int T, float* w) {
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < AGG_MAX_TOWERS; ++t) {
    w[t] = t < T ? __bfloat162float(lg[t]) : -INFINITY;
    mx = fmaxf(mx, w[t]);
  }
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < AGG_MAX_TOWERS; ++t) {
    w[t] = t < T ? __expf(w[t] - mx) : 0.f;
    s += w[t];
  }
  const float inv = 1.f / s;
#pragma unroll
  for (int t = 0; t < AGG_MAX_TOWERS; ++t) w[t] *= inv;
}

// out[n, :] = qin[n, :] + sum_t softmax(logits[n, :T])[t] * agg_t[n, :]
__global__ void __launch_bounds__(256)
tower_combine_fwd_kernel(const bf16* __restrict__ logits, int ldl, AggPtrs aggs, const bf16* __restrict__ qin,
                         bf16* __restrict__ out, long long N, int C, int T) {
  const int lane = threadIdx.x & 31;
  const int vpr = C >> 3;
  for (long long n = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); n < N; n += (long long)gridDim.x * 8) {
    float w[AGG_MAX_TOWERS];
    softmax_row(logits + n * ldl, T, w);
    for (int v = lane; v < vpr; v += 32) {
      float acc[8], f[8];
      unpack8(ldg_nc(reinterpret_cast<const uint4*>(qin + n * C) + v), acc);
#pragma unroll
      for (int t = 0; t < AGG_MAX_TOWERS; ++t) {
        if (t < T) {
          unpack8(ldg_nc(reinterpret_cast<const uint4*>(aggs.p[t] + n * C) + v), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(w[t], f[e], acc[e]);
        }
      }
      reinterpret_cast<uint4*>(out + n * C)[v] = pack8(acc);
    }
  }
}

// dagg_t[n, :] = w_t * dout[n, :];  dlogits[n, t] = w_t * (g_t - sum_s w_s g_s) with g_t = <dout[n, :], agg_t[n, :]>;
// padded logit columns [T, Tpad) receive zero.  (dqin = dout is handed through by the caller.)
__global__ void __launch_bounds__(256)
tower_combine_bwd_kernel(const bf16* __restrict__ logits, int ldl, AggPtrs aggs, const bf16* __restrict__ dout,
                         AggOutPtrs daggs, bf16* __restrict__ dlogits, long long N, int C, int T, int Tpad) {
  const int lane = threadIdx.x & 31;
  const int vpr = C >> 3;
  for (long long n = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); n < N; n += (long long)gridDim.x * 8) {
    float w[AGG_MAX_TOWERS], g[AGG_MAX_TOWERS];
    softmax_row(logits + n * ldl, T, w);
#pragma unroll
    for (int t = 0; t < AGG_MAX_TOWERS; ++t) g[t] = 0.f;
    for (int v = lane; v < vpr; v += 32) {
      float d[8], f[8], o[8];
      unpack8(ldg_nc(reinterpret_cast<const uint4*>(dout + n * C) + v), d);
#pragma unroll
      for (int t = 0; t < AGG_MAX_TOWERS; ++t) {
        if (t < T) {
          unpack8(ldg_nc(reinterpret_cast<const uint4*>(aggs.p[t] + n * C) + v), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            g[t] = fmaf(d[e], f[e], g[t]);
            o[e] = w[t] * d[e];
          }
          reinterpret_cast<uint4*>(daggs.p[t] + n * C)[v] = pack8(o);
        }
      }
    }
    float dot = 0.f;
#pragma unroll
    for (int t = 0; t < AGG_MAX_TOWERS; ++t) {
      g[t] = warp_sum(g[t]);
      dot += w[t] * g[t];
    }
    if (lane < Tpad) {
      float dl = 0.f;
#pragma unroll
      for (int t = 0; t < AGG_MAX_TOWERS; ++t)
        if (t == lane && t < T) dl = w[t] * (g[t] - dot);
      dlogits[n * ldl + lane] = __float2bfloat16(dl);
    }
  }
}

// din[b, iy, ix, :] = sum_{oy, ox} wy(oy, iy) wx(ox, ix) dout[b, oy, ox, :], with the per-axis tap weights of the forward
// kernel (source coordinate f = max((o + 0.5) * scale - 0.5, 0), taps i0 = floor(f) and min(i0 + 1, size - 1))
__device__ __forceinline__ float tap_weight(int o, float scale, int size, int i) {
  float f = ((float)o + 0.5f) * scale - 0.5f;
  if (f < 0.f) f = 0.f;
  const int i0 = (int)f;
  const int i1 = min(i0 + 1, size - 1);
  const float l = f - (float)i0;
  return (i0 == i ? 1.f - l : 0.f) + (i1 == i ? l : 0.f);
}

__global__ void bilinear_bwd_kernel(const bf16* __restrict__ dout, bf16* __restrict__ din, int B, int h, int w, int th,
                                    int tw, int C) {
  const int vpr = C >> 3;
  const long long total = (long long)B * h * w * vpr;
  const float sy = (float)h / th, sx = (float)w / tw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpr);
    long long t = i / vpr;
    const int ix = (int)(t % w);
    t /= w;
    const int iy = (int)(t % h);
    const int b = (int)(t / h);
    // output rows / columns whose source coordinate can fall in [i - 1, i + 1): a conservative integer range, the exact
    // tap test happens in tap_weight
    const int oy_lo = max(0, (int)floorf(((float)iy - 1.f + 0.5f) / sy - 0.5f) - 2);
    const int oy_hi = min(th - 1, (int)ceilf(((float)iy + 1.f + 0.5f) / sy - 0.5f) + 2);
    const int ox_lo = max(0, (int)floorf(((float)ix - 1.f + 0.5f) / sx - 0.5f) - 2);
    const int ox_hi = min(tw - 1, (int)ceilf(((float)ix + 1.f + 0.5f) / sx - 0.5f) + 2);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    const bf16* base = dout + (long long)b * th * tw * C;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      const float wy = tap_weight(oy, sy, h, iy);
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        const float wgt = wy * tap_weight(ox, sx, w, ix);
        if (wgt == 0.f) continue;
        float d[8];
        unpack8(ldg_nc(reinterpret_cast<const uint4*>(base + ((long long)oy * tw + ox) * C) + v), d);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(wgt, d[e], acc[e]);
      }
    }
    reinterpret_cast<uint4*>(din + (((long long)b * h + iy) * w + ix) * C)[v] = pack8(acc);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
static int check_combine(const void* logits, int ldl, const void* const* aggs, long long N, int C, int T) {
  CB_CHECK_ARG(N > 0 && C > 0 && C % 8 == 0, "tower_combine: N=%lld C=%d (C must be a positive multiple of 8)", N, C);
  CB_CHECK_ARG(T >= 1 && T <= AGG_MAX_TOWERS && ldl >= T, "tower_combine: T=%d out of [1,%d] or logits stride %d < T", T,
               AGG_MAX_TOWERS, ldl);
  CB_CHECK_ARG(logits && aggs, "tower_combine: null operand");
  for (int t = 0; t < T; ++t)
    CB_CHECK_ARG(aggs[t] && (reinterpret_cast<uintptr_t>(aggs[t]) & 15u) == 0, "tower_combine: aggregate %d null or unaligned", t);
  return CB_OK;
}

int tower_combine_fwd_launch(const void* logits, int ldl, const void* const* aggs, const void* qin, void* out, long long N,
                             int C, int T, cudaStream_t st) {
  int rc = check_combine(logits, ldl, aggs, N, C, T);
  if (rc) return rc;
  CB_CHECK_ARG(qin && out && ((reinterpret_cast<uintptr_t>(qin) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0,
               "tower_combine: query / output null or unaligned");
  AggPtrs a;
  for (int t = 0; t < AGG_MAX_TOWERS; ++t) a.p[t] = t < T ? static_cast<const bf16*>(aggs[t]) : nullptr;
  tower_combine_fwd_kernel<<<agg_grid_for((N + 7) / 8), 256, 0, st>>>((const bf16*)logits, ldl, a, (const bf16*)qin,
                                                                      (bf16*)out, N, C, T);
  CB_CUDA_LAUNCH_CHECK("tower_combine_fwd");
  return CB_OK;
}

int tower_combine_bwd_launch(const void* logits, int ldl, const void* const* aggs, const void* dout, void* const* daggs,
                             void* dlogits, long long N, int C, int T, cudaStream_t st) {
  int rc = check_combine(logits, ldl, aggs, N, C, T);
  if (rc) return rc;
  CB_CHECK_ARG(dout && dlogits && daggs && (reinterpret_cast<uintptr_t>(dout) & 15u) == 0,
               "tower_combine_bwd: null or unaligned gradient operand");
  CB_CHECK_ARG(ldl <= 32, "tower_combine_bwd: logits stride %d > 32", ldl);
  AggPtrs a;
  AggOutPtrs d;
  for (int t = 0; t < AGG_MAX_TOWERS; ++t) {
    a.p[t] = t < T ? static_cast<const bf16*>(aggs[t]) : nullptr;
    d.p[t] = t < T ? static_cast<bf16*>(daggs[t]) : nullptr;
    if (t < T)
      CB_CHECK_ARG(d.p[t] && (reinterpret_cast<uintptr_t>(d.p[t]) & 15u) == 0, "tower_combine_bwd: dagg %d null or unaligned", t);
  }
  tower_combine_bwd_kernel<<<agg_grid_for((N + 7) / 8), 256, 0, st>>>((const bf16*)logits, ldl, a, (const bf16*)dout, d,
                                                                      (bf16*)dlogits, N, C, T, ldl);
  CB_CUDA_LAUNCH_CHECK("tower_combine_bwd");
  return CB_OK;
}

int bilinear_bwd_launch(const void* dout, void* din, int B, int h, int w, int th, int tw, int C, cudaStream_t st) {
  CB_CHECK_ARG(B > 0 && h > 0 && w > 0 && th > 0 && tw > 0 && C > 0 && C % 8 == 0,
               "bilinear_bwd: empty grid or C=%d not a multiple of 8", C);
  CB_CHECK_ARG(dout && din && ((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(din)) & 15u) == 0,
               "bilinear_bwd: null or unaligned operand");
  const long long total = (long long)B * h * w * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)device_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  bilinear_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>((const bf16*)dout, (bf16*)din, B, h, w, th, tw, C);
  CB_CUDA_LAUNCH_CHECK("bilinear_bwd");
  return CB_OK;
}

}  // namespace cb

// cambrian_b200 — extern "C" entry points (the drop-in boundary, include/cambrian_b200.h).
#include "common.cuh"
#include "../../include/cambrian_b200.h"
#include <cstdarg>
#include <cstdio>

namespace cb {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int device_sm_count() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

// implemented in the kernel translation units
int gemm_bf16(const void*, const void*, void*, int, int, int, int, long long, long long, long long, long long,
              long long, long long, int, int, const void*, const void*, const void*, long long, long long, float,
              int, int, int, int, cudaStream_t);
int sva_window_attn_fwd_launch(const void*, void*, float*, int, const void* const*, const void* const*,
                               const void* const*, const int*, int, int, int, cudaStream_t);
int sva_window_attn_bwd_launch(const void*, const void*, const void*, const float*, void*, int,
                               const void* const*, const void* const*, const void* const*, void* const*,
                               void* const*, const int*, int, int, int, cudaStream_t);
int layernorm_fwd(const void*, const void*, const void*, void*, float*, float*, long long, int, float,
                  const void*, int, int, cudaStream_t);
int layernorm_bwd(const void*, const void*, const void*, const float*, const float*, void*, void*, void*,
                  float*, long long, long long, int, const void*, int, int, cudaStream_t);
int rmsnorm_fwd(const void*, const void*, void*, float*, long long, int, float, int, cudaStream_t);
int rmsnorm_bwd(const void*, const void*, const void*, const float*, void*, void*, float*, long long,
                long long, int, cudaStream_t);
long long norm_bwd_workspace_floats(long long, int);

}  // namespace cb

#define ST(s) static_cast<cudaStream_t>(s)

extern "C" {

int cb_version(void) { return 1; }
const char* cb_last_error(void) { return cb::g_err; }
int cb_sm_count(void) { return cb::device_sm_count(); }

int cb_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int batch, int64_t lda,
                 int64_t ldb, int64_t ldc, int64_t bsa, int64_t bsb, int64_t bsc, int a_mn, int b_mn,
                 const void* bias, const void* colscale, const void* residual, int64_t ldr, int64_t bsr,
                 float alpha, int act, int out_fp32, int accumulate, int force_bn, void* stream) {
  return cb::gemm_bf16(A, B, C, M, N, K, batch, lda, ldb, ldc, bsa, bsb, bsc, a_mn, b_mn, bias, colscale,
                       residual, ldr, bsr, alpha, act, out_fp32, accumulate, force_bn, ST(stream));
}

int cb_sva_window_attn_fwd(const void* q, void* out, float* lse, int num_towers, const void* const* k,
                           const void* const* v, const void* const* mask, const int* r, int batch,
                           int q_side, int hidden, void* stream) {
  return cb::sva_window_attn_fwd_launch(q, out, lse, num_towers, k, v, mask, r, batch, q_side, hidden,
                                        ST(stream));
}
int cb_sva_window_attn_bwd(const void* q, const void* out, const void* dout, const float* lse, void* dq,
                           int num_towers, const void* const* k, const void* const* v,
                           const void* const* mask, void* const* dk, void* const* dv, const int* r,
                           int batch, int q_side, int hidden, void* stream) {
  return cb::sva_window_attn_bwd_launch(q, out, dout, lse, dq, num_towers, k, v, mask, dk, dv, r, batch,
                                        q_side, hidden, ST(stream));
}

int cb_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd,
                     int64_t rows, int C, float eps, const void* pos, int side, int r, void* stream) {
  return cb::layernorm_fwd(x, gamma, beta, y, mean, rstd, rows, C, eps, pos, side, r, ST(stream));
}
int cb_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                     void* dx, void* dgamma, void* dbeta, float* workspace, int64_t workspace_floats,
                     int64_t rows, int C, const void* pos, int side, int r, void* stream) {
  return cb::layernorm_bwd(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, workspace, workspace_floats, rows, C,
                           pos, side, r, ST(stream));
}
int cb_rmsnorm_fwd(const void* x, const void* gamma, void* y, float* rstd, int64_t rows, int C, float eps,
                   int hf_cast, void* stream) {
  return cb::rmsnorm_fwd(x, gamma, y, rstd, rows, C, eps, hf_cast, ST(stream));
}
int cb_rmsnorm_bwd(const void* dy, const void* x, const void* gamma, const float* rstd, void* dx,
                   void* dgamma, float* workspace, int64_t workspace_floats, int64_t rows, int C,
                   void* stream) {
  return cb::rmsnorm_bwd(dy, x, gamma, rstd, dx, dgamma, workspace, workspace_floats, rows, C, ST(stream));
}
int64_t cb_norm_bwd_workspace_floats(int64_t rows, int C) { return cb::norm_bwd_workspace_floats(rows, C); }

}  // extern "C"

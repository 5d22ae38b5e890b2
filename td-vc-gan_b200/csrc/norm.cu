// Weight-norm and (conditional) instance-norm: HBM-bound row reductions with 128-bit loads and
// warp-shuffle + shared-memory reductions.
//   weight norm : util/__init__.py:16-20, model/generator.py:14, model/discriminator.py:11
//   instance norm / CIN : model/conditional_instance_norm.py:4-19
#include <algorithm>
#include "common.cuh"

namespace tdvc {

// one block per row
__global__ void wn_fwd_k(const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ w,
                         float* __restrict__ inv_norm, int cols) {
  pdl_prologue();
  __shared__ float sm[33];
  const int r = blockIdx.x;
  const float* vr = v + (long long)r * cols;
  float s = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) { float a = vr[i]; s = fmaf(a, a, s); }
  s = block_sum(s, sm);
  float inv = rsqrtf(s);
  // one Newton step: rsqrtf is 2 ulp, the reference divides by an exactly rounded sqrt
  inv = inv * (1.5f - 0.5f * s * inv * inv);
  float scale = g[r] * inv;
  float* wr = w + (long long)r * cols;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) wr[i] = vr[i] * scale;
  if (threadIdx.x == 0) inv_norm[r] = inv;
}

// The same for MANY weights in one launch (a step normalises ~270 weights; as separate launches that is ~530 graph nodes).
// table[j] = {v, g, offset of w_j in flat_w (floats), cols}; row_start[j] = first global row of weight j (n + 1 entries);
// one block per global row, inv_norm indexed by global row.
__global__ void wn_fwd_multi_k(const long long* __restrict__ table, const int* __restrict__ row_start, int n,
                               float* __restrict__ flat_w, float* __restrict__ flat_inv) {
  pdl_prologue();
  __shared__ float sm[33];
  const int gr = blockIdx.x;
  int lo = 0, hi = n - 1;                       // last j with row_start[j] <= gr
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(row_start + mid) <= gr) lo = mid; else hi = mid - 1;
  }
  const long long* e = table + 4LL * lo;
  const float* v = reinterpret_cast<const float*>(e[0]);
  const float* g = reinterpret_cast<const float*>(e[1]);
  const int cols = (int)e[3];
  const int r = gr - __ldg(row_start + lo);
  const float* vr = v + (long long)r * cols;
  // the partial sums are formed exactly as wn_fwd_k forms them for this row length (32 / 128 / 256 participating threads;
  // the others add zeros), so the batched and the per-weight launch give bit-identical weights
  const int nthr = cols >= 1024 ? 256 : (cols >= 128 ? 128 : 32);
  float s = 0.f;
  if ((int)threadIdx.x < nthr)
    for (int i = threadIdx.x; i < cols; i += nthr) { float a = vr[i]; s = fmaf(a, a, s); }
  s = block_sum(s, sm);
  float inv = rsqrtf(s);
  inv = inv * (1.5f - 0.5f * s * inv * inv);
  const float scale = g[r] * inv;
  float* wr = flat_w + e[2] + (long long)r * cols;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) wr[i] = vr[i] * scale;
  if (threadIdx.x == 0) flat_inv[gr] = inv;
}

// w = g v / n :  dg = <dw, v>/n ;  dv = (g/n) (dw - v <dw,v>/n^2)
__global__ void wn_bwd_k(const float* __restrict__ dw, const float* __restrict__ v, const float* __restrict__ g,
                         const float* __restrict__ inv_norm, float* __restrict__ dv, float* __restrict__ dg, int cols) {
  pdl_prologue();
  __shared__ float sm[33];
  const int r = blockIdx.x;
  const float* vr = v + (long long)r * cols;
  const float* dr = dw + (long long)r * cols;
  float s = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) s = fmaf(dr[i], vr[i], s);
  s = block_sum(s, sm);
  float inv = inv_norm[r];
  float gi = g[r] * inv;
  float c = s * inv * inv;
  float* o = dv + (long long)r * cols;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) o[i] = gi * (dr[i] - vr[i] * c);
  if (threadIdx.x == 0) dg[r] = s * inv;
}

// The backward for MANY weights in one launch: table[j] = {v, g, offset of dw_j in flat_dw, offset of dv_j in flat_dv, first
// row of dg_j in flat_dg, cols}; row_start as in wn_fwd_multi_k; one block per global row.  1 / ||v|| is recomputed from v
// (the forward's value may belong to a scope that is gone), so the kernel needs nothing from the forward.
__global__ void wn_bwd_multi_k(const long long* __restrict__ table, const int* __restrict__ row_start, int n,
                               const float* __restrict__ flat_dw, float* __restrict__ flat_dv, float* __restrict__ flat_dg) {
  pdl_prologue();
  __shared__ float sm[33];
  const int gr = blockIdx.x;
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(row_start + mid) <= gr) lo = mid; else hi = mid - 1;
  }
  const long long* e = table + 6LL * lo;
  const float* v = reinterpret_cast<const float*>(e[0]);
  const float* g = reinterpret_cast<const float*>(e[1]);
  const int cols = (int)e[5];
  const int r = gr - __ldg(row_start + lo);
  const float* vr = v + (long long)r * cols;
  const float* dr = flat_dw + e[2] + (long long)r * cols;
  float sv = 0.f, sd = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    const float a = vr[i];
    sv = fmaf(a, a, sv);
    sd = fmaf(dr[i], a, sd);
  }
  sv = block_sum(sv, sm);
  sd = block_sum(sd, sm);
  float inv = rsqrtf(sv);
  inv = inv * (1.5f - 0.5f * sv * inv * inv);
  const float gi = g[r] * inv;
  const float c = sd * inv * inv;
  float* o = flat_dv + e[3] + (long long)r * cols;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) o[i] = gi * (dr[i] - vr[i] * c);
  if (threadIdx.x == 0) flat_dg[e[4] + r] = sd * inv;
}

// per-(b,c) mean and 1/sqrt(var+eps) over T (biased variance, two-pass for fp32 accuracy)
__global__ void instnorm_stats_k(const float* __restrict__ x, float* __restrict__ mean, float* __restrict__ rstd, int T,
                                 float eps) {
  pdl_prologue();
  __shared__ float sm[33];
  const long long r = blockIdx.x;
  const float* xr = x + r * T;
  float s = 0.f;
  for (int i = threadIdx.x; i < T; i += blockDim.x) s += xr[i];
  s = block_sum(s, sm);
  float m = s / (float)T;
  float q = 0.f;
  for (int i = threadIdx.x; i < T; i += blockDim.x) { float d = xr[i] - m; q = fmaf(d, d, q); }
  q = block_sum(q, sm);
  if (threadIdx.x == 0) {
    mean[r] = m;
    rstd[r] = 1.0f / sqrtf(q / (float)T + eps);
  }
}

__global__ void cin_apply_fwd_k(const float* __restrict__ x, const float* __restrict__ mean,
                                const float* __restrict__ rstd, const float* __restrict__ gb, int Tg,
                                float* __restrict__ y, int B, int C, int T, float slope) {
  pdl_prologue();
  long long n = (long long)B * C * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long row = i / T;
    int t = (int)(i - row * T);
    float xh = (x[i] - mean[row]) * rstd[row];
    float v = xh;
    if (gb) {
      int b = (int)(row / C), c = (int)(row - (long long)b * C);
      int tg = Tg == 1 ? 0 : t;
      const float* gp = gb + ((long long)b * 2 * C) * Tg;
      v = fmaf(1.f + gp[(long long)c * Tg + tg], xh, gp[(long long)(C + c) * Tg + tg]);
    }
    y[i] = lrelu(v, slope);
  }
}

// one block per (b,c) row.  With dxh = dL/dxhat:  dx = rstd * (dxh - mean_t(dxh) - xhat * mean_t(dxh*xhat))
__global__ void cin_apply_bwd_k(const float* __restrict__ dy, const float* __restrict__ x,
                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                const float* __restrict__ gb, int Tg, const float* __restrict__ y_act,
                                float* __restrict__ dx, float* __restrict__ dgb, int C, int T, float slope) {
  pdl_prologue();
  __shared__ float sm[33];
  const long long row = blockIdx.x;
  const int b = (int)(row / C), c = (int)(row - (long long)b * C);
  const float m = mean[row], rs = rstd[row];
  const float* xr = x + row * T;
  const float* dyr = dy + row * T;
  const float* yr = y_act ? y_act + row * T : nullptr;
  const float* gam = gb ? gb + ((long long)b * 2 * C + c) * Tg : nullptr;
  float* dgam = dgb ? dgb + ((long long)b * 2 * C + c) * Tg : nullptr;
  float* dbet = dgb ? dgb + ((long long)b * 2 * C + C + c) * Tg : nullptr;
  float s1 = 0.f, s2 = 0.f, sg = 0.f, sb = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float g = dyr[t];
    if (slope != 1.0f && yr[t] <= 0.f) g *= slope;
    float xh = (xr[t] - m) * rs;
    float dxh = g;
    if (gam) {
      int tg = Tg == 1 ? 0 : t;
      dxh = g * (1.f + gam[tg]);
      if (Tg == 1) { sg = fmaf(g, xh, sg); sb += g; }
      else { dgam[t] = g * xh; dbet[t] = g; }
    }
    s1 += dxh;
    s2 = fmaf(dxh, xh, s2);
  }
  s1 = block_sum(s1, sm);
  s2 = block_sum(s2, sm);
  if (gam && Tg == 1) {
    sg = block_sum(sg, sm);
    sb = block_sum(sb, sm);
    if (threadIdx.x == 0) { dgam[0] = sg; dbet[0] = sb; }
  }
  const float m1 = s1 / (float)T, m2 = s2 / (float)T;
  float* dxr = dx + row * T;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float g = dyr[t];
    if (slope != 1.0f && yr[t] <= 0.f) g *= slope;
    float xh = (xr[t] - m) * rs;
    float dxh = gam ? g * (1.f + gam[Tg == 1 ? 0 : t]) : g;
    dxr[t] = rs * (dxh - m1 - xh * m2);
  }
}

}  // namespace tdvc
using namespace tdvc;

extern "C" int tdvc_weight_norm_fwd(const float* v, const float* g, float* w, float* inv_norm, int rows, int cols,
                                    void* stream) {
  TDVC_CHECK_ARG(rows > 0 && cols > 0 && v && g && w && inv_norm);
  int threads = cols >= 1024 ? 256 : (cols >= 128 ? 128 : 32);
  tdvc::launch_k(wn_fwd_k, rows, threads, 0, (cudaStream_t)stream, v, g, w, inv_norm, cols);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_weight_norm_fwd_multi(const void* table, const void* row_start, int n_weights, int total_rows,
                                          float* flat_w, float* flat_inv, void* stream) {
  TDVC_CHECK_ARG(table && row_start && n_weights > 0 && total_rows > 0 && flat_w && flat_inv);
  tdvc::launch_k(wn_fwd_multi_k, total_rows, 256, 0, (cudaStream_t)stream, (const long long*)table, (const int*)row_start,
                 n_weights, flat_w, flat_inv);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_weight_norm_bwd_multi(const void* table, const void* row_start, int n_weights, int total_rows,
                                          const float* flat_dw, float* flat_dv, float* flat_dg, void* stream) {
  TDVC_CHECK_ARG(table && row_start && n_weights > 0 && total_rows > 0 && flat_dw && flat_dv && flat_dg);
  tdvc::launch_k(wn_bwd_multi_k, total_rows, 128, 0, (cudaStream_t)stream, (const long long*)table, (const int*)row_start,
                 n_weights, flat_dw, flat_dv, flat_dg);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_weight_norm_bwd(const float* dw, const float* v, const float* g, const float* inv_norm, float* dv,
                                    float* dg, int rows, int cols, void* stream) {
  TDVC_CHECK_ARG(rows > 0 && cols > 0 && dw && v && g && inv_norm && dv && dg);
  int threads = cols >= 1024 ? 256 : (cols >= 128 ? 128 : 32);
  tdvc::launch_k(wn_bwd_k, rows, threads, 0, (cudaStream_t)stream, dw, v, g, inv_norm, dv, dg, cols);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_instnorm_stats(const float* x, float* mean, float* rstd, int BC, int T, float eps, void* stream) {
  TDVC_CHECK_ARG(BC >= 0 && T > 0 && x && mean && rstd);
  if (BC == 0) return TDVC_OK;
  tdvc::launch_k(instnorm_stats_k, BC, T >= 1024 ? 256 : 128, 0, (cudaStream_t)stream, x, mean, rstd, T, eps);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_cin_apply_fwd(const float* x, const float* mean, const float* rstd, const float* gb, int Tg,
                                  float* y, int B, int C, int T, float out_slope, void* stream) {
  TDVC_CHECK_ARG(B >= 0 && C > 0 && T > 0 && x && mean && rstd && y);
  TDVC_CHECK_ARG(gb == nullptr || Tg == 1 || Tg == T);
  if (B == 0) return TDVC_OK;
  long long n = (long long)B * C * T;
  int blocks = (int)std::min<long long>((n + 255) / 256, 16LL * num_sms());
  tdvc::launch_k(cin_apply_fwd_k, blocks, 256, 0, (cudaStream_t)stream, x, mean, rstd, gb, Tg, y, B, C, T, out_slope);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_cin_apply_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                                  const float* gb, int Tg, const float* y_act, float* dx, float* dgb, int B, int C,
                                  int T, float out_slope, void* stream) {
  TDVC_CHECK_ARG(B >= 0 && C > 0 && T > 0 && dy && x && mean && rstd && dx);
  TDVC_CHECK_ARG(gb == nullptr || ((Tg == 1 || Tg == T) && dgb != nullptr));
  TDVC_CHECK_ARG(out_slope == 1.0f || y_act != nullptr);
  if (B == 0) return TDVC_OK;
  tdvc::launch_k(cin_apply_bwd_k, B * C, T >= 1024 ? 256 : 128, 0, (cudaStream_t)stream, dy, x, mean, rstd, gb, Tg, y_act, dx, dgb,
                                                                             C, T, out_slope);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

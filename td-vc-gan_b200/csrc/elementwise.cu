// Bandwidth-bound elementwise / small-reduction kernels of the FiLM / MRF / encoder glue
// (model/generator.py:96-111,186-194,271,387-399; model/discriminator.py:20,31,38).
// All are grid-stride, float4-vectorised where the layout allows, coalesced along time.
#include <algorithm>
#include "common.cuh"

namespace tdvc {

static inline int ew_blocks(long long n_items) {
  long long b = (n_items + 255) / 256;
  long long cap = 16LL * num_sms();
  return (int)std::max<long long>(1, std::min(b, cap));
}

__global__ void lrelu_fwd_k(const float* __restrict__ x, float* __restrict__ y, long long n, float slope) {
  pdl_prologue();
  long long n4 = n >> 2;
  long long stride = (long long)gridDim.x * blockDim.x;
  long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float4* y4 = reinterpret_cast<float4*>(y);
  for (long long i = i0; i < n4; i += stride) {
    float4 v = x4[i];
    v.x = lrelu(v.x, slope); v.y = lrelu(v.y, slope); v.z = lrelu(v.z, slope); v.w = lrelu(v.w, slope);
    y4[i] = v;
  }
  for (long long i = (n4 << 2) + i0; i < n; i += stride) y[i] = lrelu(x[i], slope);
}

__global__ void act_bwd_k(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dz,
                          long long n, int act, float slope) {
  pdl_prologue();
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float g = dy[i], o = y[i];
    if (act == TDVC_ACT_LRELU) g = o > 0.f ? g : g * slope;
    else if (act == TDVC_ACT_TANH) g = g * (1.f - o * o);
    dz[i] = g;
  }
}

__global__ void film_fwd_k(const float* __restrict__ h, const float* __restrict__ gb, float* __restrict__ y, int B,
                           int C, int T) {
  pdl_prologue();
  long long n = (long long)B * C * T;
  long long stride = (long long)gridDim.x * blockDim.x;
  long long CT = (long long)C * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    long long b = i / CT, r = i - b * CT;
    const float* g = gb + b * 2 * CT;
    y[i] = fmaf(h[i], 1.f + g[r], g[CT + r]);
  }
}

__global__ void film_bwd_k(const float* __restrict__ dy, const float* __restrict__ h, const float* __restrict__ gb,
                           float* __restrict__ dh, float* __restrict__ dgb, int B, int C, int T) {
  pdl_prologue();
  long long n = (long long)B * C * T;
  long long stride = (long long)gridDim.x * blockDim.x;
  long long CT = (long long)C * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    long long b = i / CT, r = i - b * CT;
    float g = dy[i];
    dh[i] = g * (1.f + gb[b * 2 * CT + r]);
    float* d = dgb + b * 2 * CT;
    d[r] = g * h[i];
    d[CT + r] = g;
  }
}

__global__ void add3_scale_k(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                             float* __restrict__ y, long long n, float alpha) {
  pdl_prologue();
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = a[i];
    if (b) v += b[i];
    if (c) v += c[i];
    y[i] = alpha * v;
  }
}

// F.normalize(dim=1): one thread per (b,t) column; consecutive threads -> consecutive t (coalesced per channel)
__global__ void l2norm_fwd_k(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ inv, int B, int C,
                             int T) {
  pdl_prologue();
  long long n = (long long)B * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long b = i / T;
    int t = (int)(i - b * T);
    const float* xp = x + b * C * T + t;
    float s = 0.f;
    for (int c = 0; c < C; ++c) { float v = xp[(long long)c * T]; s = fmaf(v, v, s); }
    float r = 1.f / fmaxf(sqrtf(s), 1e-12f);
    inv[i] = r;
    float* yp = y + b * C * T + t;
    for (int c = 0; c < C; ++c) yp[(long long)c * T] = xp[(long long)c * T] * r;
  }
}

// y = x*r  =>  dx = r*(dy - y * sum_c(dy*y))   (the clamp branch has zero measure; same as ATen)
__global__ void l2norm_bwd_k(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ inv,
                             float* __restrict__ dx, int B, int C, int T) {
  pdl_prologue();
  long long n = (long long)B * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long b = i / T;
    int t = (int)(i - b * T);
    long long base = b * C * T + t;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(dy[base + (long long)c * T], y[base + (long long)c * T], s);
    float r = inv[i];
    for (int c = 0; c < C; ++c) {
      long long o = base + (long long)c * T;
      dx[o] = r * (dy[o] - y[o] * s);
    }
  }
}

__global__ void cond_concat_fwd_k(const float* __restrict__ c, const float* __restrict__ e, float* __restrict__ out,
                                  int B, int Cc, int Ce, int T, int c_first) {
  pdl_prologue();
  int Ct = Cc + Ce;
  long long n = (long long)B * Ct * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long row = i / T;
    int t = (int)(i - row * T);
    int b = (int)(row / Ct), ch = (int)(row - (long long)b * Ct);
    int cc = c_first ? ch : ch - Ce;          // index into c (valid when 0 <= cc < Cc)
    int ce = c_first ? ch - Cc : ch;          // index into e
    out[i] = (cc >= 0 && cc < Cc) ? c[(long long)b * Cc + cc] : e[((long long)b * Ce + ce) * T + t];
  }
}

// one block per (b, channel) row of dout: time-sum for the constant channels, copy for the rest
__global__ void cond_concat_bwd_k(const float* __restrict__ dout, float* __restrict__ dc, float* __restrict__ de, int B,
                                  int Cc, int Ce, int T, int c_first) {
  pdl_prologue();
  __shared__ float sm[33];
  int Ct = Cc + Ce;
  int row = blockIdx.x;
  int b = row / Ct, ch = row - b * Ct;
  int cc = c_first ? ch : ch - Ce;
  int ce = c_first ? ch - Cc : ch;
  const float* src = dout + (long long)row * T;
  if (cc >= 0 && cc < Cc) {
    float s = 0.f;
    for (int t = threadIdx.x; t < T; t += blockDim.x) s += src[t];
    s = block_sum(s, sm);
    if (threadIdx.x == 0) dc[(long long)b * Cc + cc] = s;
  } else if (de) {
    float* d = de + ((long long)b * Ce + ce) * T;
    for (int t = threadIdx.x; t < T; t += blockDim.x) d[t] = src[t];
  }
}

// WaveNet gate of the SSL content encoder (model/ssl_encoder.py:7-14, fused_add_tanh_sigmoid_multiply):
//   acts[b, c, t] = tanh(a[b, c, t] + g[b, c, t]) * sigmoid(a[b, H + c, t] + g[b, H + c, t]),  a, g: [B, 2H, T], g optional.
// The backward recomputes the two activations from a (+ g): d(a_tanh) = d * s * (1 - th^2), d(a_sig) = d * th * s * (1 - s);
// the gradient of g is the same tensor.
__global__ void gate_fwd_k(const float* __restrict__ a, const float* __restrict__ g, float* __restrict__ y, int B, int H, int T) {
  pdl_prologue();
  const long long n = (long long)B * H * T, ht = (long long)H * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / ht, r = i - b * ht;
    const long long ia = b * 2 * ht + r;
    float u = a[ia], v = a[ia + ht];
    if (g) { u += g[ia]; v += g[ia + ht]; }
    y[i] = tanhf(u) * (1.f / (1.f + expf(-v)));
  }
}

__global__ void gate_bwd_k(const float* __restrict__ dy, const float* __restrict__ a, const float* __restrict__ g,
                           float* __restrict__ da, int B, int H, int T) {
  pdl_prologue();
  const long long n = (long long)B * H * T, ht = (long long)H * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / ht, r = i - b * ht;
    const long long ia = b * 2 * ht + r;
    float u = a[ia], v = a[ia + ht];
    if (g) { u += g[ia]; v += g[ia + ht]; }
    const float th = tanhf(u), sg = 1.f / (1.f + expf(-v)), d = dy[i];
    da[ia] = d * sg * (1.f - th * th);
    da[ia + ht] = d * th * sg * (1.f - sg);
  }
}

}  // namespace tdvc
using namespace tdvc;

extern "C" int tdvc_leaky_relu_fwd(const float* x, float* y, int64_t n, float slope, void* stream) {
  TDVC_CHECK_ARG(n >= 0);
  if (n == 0) return TDVC_OK;
  TDVC_CHECK_ARG(x && y && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0));
  tdvc::launch_k(lrelu_fwd_k, ew_blocks(n / 4 + 1), 256, 0, (cudaStream_t)stream, x, y, n, slope);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_act_bwd_from_output(const float* dy, const float* y, float* dz, int64_t n, int act, float slope,
                                        void* stream) {
  TDVC_CHECK_ARG(n >= 0);
  if (n == 0) return TDVC_OK;
  TDVC_CHECK_ARG(dy && y && dz);
  tdvc::launch_k(act_bwd_k, ew_blocks(n), 256, 0, (cudaStream_t)stream, dy, y, dz, n, act, slope);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_film_fwd(const float* h, const float* gb, float* y, int B, int C, int T, void* stream) {
  TDVC_CHECK_ARG(B >= 0 && C > 0 && T > 0 && h && gb && y);
  if (B == 0) return TDVC_OK;
  tdvc::launch_k(film_fwd_k, ew_blocks((long long)B * C * T), 256, 0, (cudaStream_t)stream, h, gb, y, B, C, T);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_film_bwd(const float* dy, const float* h, const float* gb, float* dh, float* dgb, int B, int C,
                             int T, void* stream) {
  TDVC_CHECK_ARG(B >= 0 && C > 0 && T > 0 && dy && h && gb && dh && dgb);
  if (B == 0) return TDVC_OK;
  tdvc::launch_k(film_bwd_k, ew_blocks((long long)B * C * T), 256, 0, (cudaStream_t)stream, dy, h, gb, dh, dgb, B, C, T);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_add3_scale(const float* a, const float* b, const float* c, float* y, int64_t n, float alpha,
                               void* stream) {
  TDVC_CHECK_ARG(n >= 0);
  if (n == 0) return TDVC_OK;
  TDVC_CHECK_ARG(a && y);
  tdvc::launch_k(add3_scale_k, ew_blocks(n), 256, 0, (cudaStream_t)stream, a, b, c, y, n, alpha);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_l2norm_fwd(const float* x, float* y, float* inv, int B, int C, int T, void* stream) {
  TDVC_CHECK_ARG(B >= 0 && C > 0 && T > 0 && x && y && inv);
  if (B == 0) return TDVC_OK;
  tdvc::launch_k(l2norm_fwd_k, ew_blocks((long long)B * T), 256, 0, (cudaStream_t)stream, x, y, inv, B, C, T);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_l2norm_bwd(const float* dy, const float* y, const float* inv, float* dx, int B, int C, int T,
                               void* stream) {
  TDVC_CHECK_ARG(B >= 0 && C > 0 && T > 0 && dy && y && inv && dx);
  if (B == 0) return TDVC_OK;
  tdvc::launch_k(l2norm_bwd_k, ew_blocks((long long)B * T), 256, 0, (cudaStream_t)stream, dy, y, inv, dx, B, C, T);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_cond_concat_fwd(const float* c, const float* e, float* out, int B, int Cc, int Ce, int T,
                                    int c_first, void* stream) {
  TDVC_CHECK_ARG(B >= 0 && Cc >= 0 && Ce >= 0 && Cc + Ce > 0 && T > 0 && out);
  TDVC_CHECK_ARG((Cc == 0 || c) && (Ce == 0 || e));
  if (B == 0) return TDVC_OK;
  tdvc::launch_k(cond_concat_fwd_k, ew_blocks((long long)B * (Cc + Ce) * T), 256, 0, (cudaStream_t)stream, c, e, out, B, Cc, Ce, T, c_first);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_cond_concat_bwd(const float* dout, float* dc, float* de, int B, int Cc, int Ce, int T,
                                    int c_first, void* stream) {
  TDVC_CHECK_ARG(B >= 0 && Cc >= 0 && Ce >= 0 && Cc + Ce > 0 && T > 0 && dout);
  TDVC_CHECK_ARG(Cc == 0 || dc);
  if (B == 0) return TDVC_OK;
  tdvc::launch_k(cond_concat_bwd_k, B * (Cc + Ce), 256, 0, (cudaStream_t)stream, dout, dc, de, B, Cc, Ce, T, c_first);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_gate_fwd(const float* a, const float* g, float* y, int B, int H, int T, void* stream) {
  TDVC_CHECK_ARG(a && y && B >= 0 && H > 0 && T > 0);
  if (B == 0) return TDVC_OK;
  tdvc::launch_k(gate_fwd_k, ew_blocks((long long)B * H * T), 256, 0, (cudaStream_t)stream, a, g, y, B, H, T);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_gate_bwd(const float* dy, const float* a, const float* g, float* da, int B, int H, int T, void* stream) {
  TDVC_CHECK_ARG(dy && a && da && B >= 0 && H > 0 && T > 0);
  if (B == 0) return TDVC_OK;
  tdvc::launch_k(gate_bwd_k, ew_blocks((long long)B * H * T), 256, 0, (cudaStream_t)stream, dy, a, g, da, B, H, T);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

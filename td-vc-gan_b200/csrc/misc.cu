// Pooling, label gather, loss reductions and the fused multi-tensor AdamW.
//   AvgPool1d(4,2,1,count_include_pad=False): model/discriminator.py:63,72
//   x.gather(1,label): model/discriminator.py:49-51
//   LSGAN F.mse_loss: train.py:273-281,327-331 ; feature-matching F.l1_loss: util/losses.py:55-68
//   torch.optim.AdamW: train.py:188-189
#include <algorithm>
#include "common.cuh"

namespace tdvc {

static inline int ew_blocks(long long n_items) {
  long long b = (n_items + 255) / 256;
  long long cap = 16LL * num_sms();
  return (int)std::max<long long>(1, std::min(b, cap));
}

__global__ void avgpool_fwd_k(const float* __restrict__ x, float* __restrict__ y, long long rows, int Tin, int Tout) {
  pdl_prologue();
  long long n = rows * Tout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / Tout;
    int o = (int)(i - r * Tout);
    int lo = max(2 * o - 1, 0), hi = min(2 * o + 2, Tin - 1);
    const float* xr = x + r * Tin;
    float s = 0.f;
    for (int t = lo; t <= hi; ++t) s += xr[t];
    y[i] = s / (float)(hi - lo + 1);
  }
}

__global__ void avgpool_bwd_k(const float* __restrict__ dy, float* __restrict__ dx, long long rows, int Tin, int Tout) {
  pdl_prologue();
  long long n = rows * Tin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / Tin;
    int t = (int)(i - r * Tin);
    // windows o with 2o-1 <= t <= 2o+2
    int o_lo = max((t - 2 + 1) / 2, 0);
    if (t - 2 < 0) o_lo = 0;
    int o_hi = min((t + 1) / 2, Tout - 1);
    float s = 0.f;
    for (int o = o_lo; o <= o_hi; ++o) {
      int lo = max(2 * o - 1, 0), hi = min(2 * o + 2, Tin - 1);
      if (t >= lo && t <= hi) s += dy[r * Tout + o] / (float)(hi - lo + 1);
    }
    dx[i] = s;
  }
}

__global__ void select_fwd_k(const float* __restrict__ x, const int64_t* __restrict__ label, float* __restrict__ y,
                             int B, int C, int T) {
  pdl_prologue();
  long long n = (long long)B * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int b = (int)(i / T);
    int t = (int)(i - (long long)b * T);
    long long l = label[b];
    y[i] = (l >= 0 && l < C) ? x[((long long)b * C + l) * T + t] : 0.f;
  }
}

__global__ void select_bwd_k(const float* __restrict__ dy, const int64_t* __restrict__ label, float* __restrict__ dx,
                             int B, int C, int T) {
  pdl_prologue();
  long long n = (long long)B * C * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long row = i / T;
    int t = (int)(i - row * T);
    int b = (int)(row / C), c = (int)(row - (long long)b * C);
    dx[i] = (label[b] == c) ? dy[(long long)b * T + t] : 0.f;
  }
}

// The discriminator's output layer with the label gather folded in (model/discriminator.py:36,49-51: a k3 conv to one
// logit track per speaker, of which only the target speaker's row is kept).  Only that row is computed:
//   y[b, t] = sum_{c, k} w[label[b], c, k] * x[b, c, t + k - pad]          (stride 1, zero padding, no bias)
// one CTA per (sample, 32 time steps); warp = channel slice, lane = time step; the 32 partial sums meet in shared memory.
template <int K>
__global__ void __launch_bounds__(1024) select_conv_fwd_k(const float* __restrict__ x, const float* __restrict__ w,
                                                          const int64_t* __restrict__ label, float* __restrict__ y, int C, int T,
                                                          int NC) {
  pdl_prologue();
  constexpr int pad = (K - 1) / 2;
  __shared__ float sm[32][33];
  const int b = blockIdx.y, lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int t = blockIdx.x * 32 + lane;
  const long long l = label[b];
  float acc = 0.f;
  if (l >= 0 && l < NC && t < T) {
    const float* xt = x + (long long)b * C * T + t - pad;
    const float* wl = w + l * (long long)C * K;
    bool in[K];
#pragma unroll
    for (int k = 0; k < K; ++k) in[k] = (t + k - pad >= 0) && (t + k - pad < T);
    // every load of U channels is issued before the first multiply-add (a branch per tap serialised them: 96 dependent L2
    // round trips per thread; now C / (32 U) batches of 2 K U independent loads)
    constexpr int U = K <= 3 ? 8 : 4;
    for (int c0 = slice; c0 < C; c0 += 32 * U) {
      float xv[U][K], wv[U][K];
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const int c = c0 + 32 * i;
        const bool c_ok = c < C;
        const float* xr = xt + (long long)c * T;
        const float* wr = wl + (long long)c * K;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          xv[i][k] = (c_ok && in[k]) ? __ldg(xr + k) : 0.f;
          wv[i][k] = c_ok ? __ldg(wr + k) : 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < U; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc = fmaf(wv[i][k], xv[i][k], acc);
      }
    }
  }
  sm[slice][lane] = acc;
  __syncthreads();
  if (slice == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += sm[i][lane];
    if (t < T) y[(long long)b * T + t] = s;
  }
}

// Backward of the above.  One warp per (sample, channel) row, lanes over time:
//   dx[b, c, t]          = sum_k w[label[b], c, k] * dy[b, t - k + pad]                       (dx != nullptr)
//   dw[label[b], c, k]  += sum_t dy[b, t] * x[b, c, t + k - pad]                              (dw != nullptr, zero-filled by
//                                                                                             the caller; samples may share a label)
template <int K>
__global__ void __launch_bounds__(256) select_conv_bwd_k(const float* __restrict__ dy, const float* __restrict__ x,
                                                         const float* __restrict__ w, const int64_t* __restrict__ label,
                                                         float* __restrict__ dx, float* __restrict__ dw, int B, int C, int T,
                                                         int NC) {
  pdl_prologue();
  constexpr int pad = (K - 1) / 2;
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)B * C;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < rows; row += warps) {
    const int b = (int)(row / C), c = (int)(row - (long long)b * C);
    const long long l = label[b];
    const bool ok = l >= 0 && l < NC;
    const float* dyb = dy + (long long)b * T;
    const float* xr = x + row * T;
    float wk[K], dwk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      wk[k] = ok ? __ldg(w + (l * C + c) * K + k) : 0.f;
      dwk[k] = 0.f;
    }
    for (int t0 = 0; t0 < T; t0 += 32) {
      const int t = t0 + lane;
      if (t < T) {
        // dy taps t - pad .. t + pad serve both sums: dx uses dy[t - k + pad], dw uses dy[t] and x[t + k - pad]
        float dyv[K], xv[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int u = t + k - pad;
          const bool in = u >= 0 && u < T;
          dyv[k] = in ? __ldg(dyb + u) : 0.f;
          xv[k] = (in && dw) ? __ldg(xr + u) : 0.f;
        }
        if (dx) {
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < K; ++k) s = fmaf(wk[k], dyv[K - 1 - k], s);      // dy[t - k + pad] = tap K-1-k of the window
          dx[row * T + t] = s;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) dwk[k] = fmaf(dyv[pad], xv[k], dwk[k]);
      }
    }
    if (dw && ok) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float s = warp_sum(dwk[k]);
        if (lane == 0) atomicAdd(dw + (l * C + c) * K + k, s);
      }
    }
  }
}

__global__ void sq_err_const_sum_k(const float* __restrict__ a, float target, float scale, float* __restrict__ out,
                                   long long n) {
  pdl_prologue();
  __shared__ float sm[33];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float d = a[i] - target;
    s = fmaf(d, d, s);
  }
  s = block_sum(s, sm);
  if (threadIdx.x == 0) atomicAdd(out, s * scale);
}

__global__ void sq_err_const_bwd_k(const float* __restrict__ a, float target, float scale,
                                   const float* __restrict__ gscale, float* __restrict__ da, long long n) {
  pdl_prologue();
  float g = 2.f * scale * gscale[0];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    da[i] = g * (a[i] - target);
}

__global__ void abs_diff_sum_k(const float* __restrict__ a, const float* __restrict__ b, float scale,
                               float* __restrict__ out, long long n) {
  pdl_prologue();
  __shared__ float sm[33];
  float s = 0.f;
  long long n4 = n >> 2;
  long long stride = (long long)gridDim.x * blockDim.x;
  long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  for (long long i = i0; i < n4; i += stride) {
    float4 u = a4[i], v = b4[i];
    s += fabsf(u.x - v.x) + fabsf(u.y - v.y) + fabsf(u.z - v.z) + fabsf(u.w - v.w);
  }
  for (long long i = (n4 << 2) + i0; i < n; i += stride) s += fabsf(a[i] - b[i]);
  s = block_sum(s, sm);
  if (threadIdx.x == 0) atomicAdd(out, s * scale);
}

__global__ void abs_diff_bwd_k(const float* __restrict__ a, const float* __restrict__ b, float scale,
                               const float* __restrict__ gscale, float* __restrict__ da, long long n) {
  pdl_prologue();
  float g = scale * gscale[0];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float d = a[i] - b[i];
    da[i] = d > 0.f ? g : (d < 0.f ? -g : 0.f);
  }
}

// sum_j scale_j * sum_i |a_j[i] - b_j[i]| for up to TDVC_L1_MAX_JOBS tensor pairs in ONE launch (the 30 feature maps of
// util/losses.py:55-68), and its gradient.  The job table travels by value in the kernel parameters: the operands are
// transient activations, so a device-resident pointer table would need a host-to-device copy per step, which a CUDA
// graph cannot capture.  blockIdx.y = job; blockIdx.x strides over the job's elements.
struct L1Jobs {
  const float* a[TDVC_L1_MAX_JOBS];
  const float* b[TDVC_L1_MAX_JOBS];
  float* da[TDVC_L1_MAX_JOBS];
  long long n[TDVC_L1_MAX_JOBS];
  float scale[TDVC_L1_MAX_JOBS];
};

__global__ void abs_diff_sum_multi_k(const __grid_constant__ L1Jobs jobs, float* __restrict__ out) {
  pdl_prologue();
  __shared__ float sm[33];
  const int j = blockIdx.y;
  const float* a = jobs.a[j];
  const float* b = jobs.b[j];
  const long long n = jobs.n[j];
  float s = 0.f;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  for (long long i = i0; i < n4; i += stride) {
    const float4 u = a4[i], v = b4[i];
    s += fabsf(u.x - v.x) + fabsf(u.y - v.y) + fabsf(u.z - v.z) + fabsf(u.w - v.w);
  }
  for (long long i = (n4 << 2) + i0; i < n; i += stride) s += fabsf(a[i] - b[i]);
  s = block_sum(s, sm);
  if (threadIdx.x == 0 && s != 0.f) atomicAdd(out, s * jobs.scale[j]);
}

__global__ void abs_diff_bwd_multi_k(const __grid_constant__ L1Jobs jobs, const float* __restrict__ gscale) {
  pdl_prologue();
  const int j = blockIdx.y;
  const float* a = jobs.a[j];
  const float* b = jobs.b[j];
  float* da = jobs.da[j];
  const long long n = jobs.n[j];
  const float g = jobs.scale[j] * gscale[0];
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  float4* d4 = reinterpret_cast<float4*>(da);
  auto sg = [g](float d) { return d > 0.f ? g : (d < 0.f ? -g : 0.f); };
  for (long long i = i0; i < n4; i += stride) {
    const float4 u = a4[i], v = b4[i];
    d4[i] = make_float4(sg(u.x - v.x), sg(u.y - v.y), sg(u.z - v.z), sg(u.w - v.w));
  }
  for (long long i = (n4 << 2) + i0; i < n; i += stride) da[i] = sg(a[i] - b[i]);
}

// AdamW exactly as torch.optim.AdamW (decoupled decay, bias correction, eps outside the sqrt of v_hat):
//   p *= 1 - lr*wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void adamw_multi_k(float* const* __restrict__ params, const float* const* __restrict__ grads,
                              float* const* __restrict__ m1, float* const* __restrict__ m2,
                              const int64_t* __restrict__ sizes, int chunks_per_tensor, float lr, float b1, float b2,
                              float eps, float wd, float bc1, float bc2_sqrt, float gscale,
                              const float* __restrict__ step_dev) {
  pdl_prologue();
  if (step_dev) {      // step counter lives on the device (CUDA-graph replays must advance it)
    const float st = *step_dev;
    bc1 = 1.f - powf(b1, st);
    bc2_sqrt = sqrtf(1.f - powf(b2, st));
  }
  const int ti = blockIdx.x / chunks_per_tensor;
  const int ch = blockIdx.x - ti * chunks_per_tensor;
  const long long n = sizes[ti];
  float* p = params[ti];
  const float* g = grads[ti];
  float* m = m1[ti];
  float* v = m2[ti];
  if (g == nullptr) return;
  const float step_size = lr / bc1;
  for (long long i = (long long)ch * blockDim.x + threadIdx.x; i < n; i += (long long)chunks_per_tensor * blockDim.x) {
    float gi = g[i] * gscale;
    float pi = p[i] * (1.f - lr * wd);
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}


// The same update with the work cut into equal blocks of ADAMW_BLOCK elements by a host-built table (tensor index, first
// element): one CTA per block, 16 elements per thread with four independent load groups in flight.  With one grid row of
// <= 64 chunks per tensor (kernel above) the 5 M-element discriminator layers were walked by 64 CTAs of 256 threads, four
// scalar loads in flight each -- 2.1 TB/s over the 0.9 GB a step's two optimiser launches move.
constexpr int ADAMW_BLOCK = 4096;
__global__ void __launch_bounds__(256) adamw_blocks_k(float* const* __restrict__ params, const float* const* __restrict__ grads,
                                                      float* const* __restrict__ m1, float* const* __restrict__ m2,
                                                      const int64_t* __restrict__ sizes, const int2* __restrict__ blocks, float lr,
                                                      float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                                                      float gscale, const float* __restrict__ step_dev) {
  pdl_prologue();
  if (step_dev) {
    const float st = *step_dev;
    bc1 = 1.f - powf(b1, st);
    bc2_sqrt = sqrtf(1.f - powf(b2, st));
  }
  const int2 blk = blocks[blockIdx.x];
  const int ti = blk.x;
  const float* g = grads[ti];
  if (g == nullptr) return;
  const long long n = sizes[ti];
  float* p = params[ti];
  float* m = m1[ti];
  float* v = m2[ti];
  const float step_size = lr / bc1, decay = 1.f - lr * wd;
  const long long base = (long long)blk.y + threadIdx.x;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float gi[4], pi[4], mi[4], vi[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = base + 256LL * (4 * q + u);
      const bool ok = i < n;
      gi[u] = ok ? g[i] : 0.f;
      pi[u] = ok ? p[i] : 0.f;
      mi[u] = ok ? m[i] : 0.f;
      vi[u] = ok ? v[i] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = base + 256LL * (4 * q + u);
      if (i < n) {
        const float gg = gi[u] * gscale;
        const float mm = b1 * mi[u] + (1.f - b1) * gg;
        const float vv = b2 * vi[u] + (1.f - b2) * gg * gg;
        m[i] = mm;
        v[i] = vv;
        p[i] = pi[u] * decay - step_size * (mm / (sqrtf(vv) / bc2_sqrt + eps));
      }
    }
  }
}

__global__ void step_inc_k(float* step) {
  pdl_prologue(); *step += 1.f; }


// ------------------------------------------------------------------------------------------ space <-> depth
// A dense strided convolution whose kernel is a whole number of strides (the encoder's k = 2r, stride r downsamplers and,
// transposed, the decoder's upsamplers: generator.py:214-249, 299-347) is a stride-1 convolution on the input viewed
// as frames of `s` samples: xs[b, p*C + c, q] = x[b, c, s*q + p - pad] (zero outside [0, T)).  These two kernels are the
// frame view and its inverse (each is the other's gradient); the convolution itself then runs on the tcgen05 path.
constexpr int S2D_TT = 128;     // frames per block
constexpr int S2D_MAX_S = 16;

__global__ void __launch_bounds__(128) s2d_k(const float* __restrict__ x, float* __restrict__ out, int C, int T, int s, int pad,
                                             int Tq) {
  pdl_prologue();
  __shared__ float sm[S2D_MAX_S * S2D_TT];
  const int b = blockIdx.z, c = blockIdx.y, q0 = blockIdx.x * S2D_TT;
  const float* xr = x + ((long long)b * C + c) * T;
  const int u0 = s * q0 - pad, n = s * min(S2D_TT, Tq - q0);
  for (int i = threadIdx.x; i < n; i += 128) {
    const int u = u0 + i;
    sm[i] = (u >= 0 && u < T) ? __ldg(xr + u) : 0.f;
  }
  __syncthreads();
  const int q = q0 + threadIdx.x;
  if (q < Tq)
    for (int p = 0; p < s; ++p) out[((long long)b * s * C + (long long)p * C + c) * Tq + q] = sm[s * threadIdx.x + p];
}

// y[b, c, u] = in[b, p*C + c, q] with s*q + p = u + pad (zero when q >= Tq)
__global__ void __launch_bounds__(128) d2s_k(const float* __restrict__ in, float* __restrict__ y, int C, int Tq, int s, int pad,
                                             int Tout) {
  pdl_prologue();
  __shared__ float sm[S2D_MAX_S * S2D_TT];
  const int b = blockIdx.z, c = blockIdx.y, q0 = blockIdx.x * S2D_TT;
  const int q = q0 + threadIdx.x;
  for (int p = 0; p < s; ++p)
    sm[s * threadIdx.x + p] = (q < Tq) ? __ldg(in + ((long long)b * s * C + (long long)p * C + c) * Tq + q) : 0.f;
  __syncthreads();
  float* yr = y + ((long long)b * C + c) * Tout;
  const int u0 = s * q0 - pad;
  for (int i = threadIdx.x; i < s * S2D_TT; i += 128) {
    const int u = u0 + i;
    if (u >= 0 && u < Tout) yr[u] = sm[i];
  }
}

}  // namespace tdvc
using namespace tdvc;

extern "C" int tdvc_avgpool4s2_fwd(const float* x, float* y, int BC, int Tin, int Tout, void* stream) {
  TDVC_CHECK_ARG(BC >= 0 && Tin > 0 && Tout == (Tin + 2 - 4) / 2 + 1 && x && y);
  if (BC == 0) return TDVC_OK;
  tdvc::launch_k(avgpool_fwd_k, ew_blocks((long long)BC * Tout), 256, 0, (cudaStream_t)stream, x, y, BC, Tin, Tout);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_avgpool4s2_bwd(const float* dy, float* dx, int BC, int Tin, int Tout, void* stream) {
  TDVC_CHECK_ARG(BC >= 0 && Tin > 0 && Tout == (Tin + 2 - 4) / 2 + 1 && dy && dx);
  if (BC == 0) return TDVC_OK;
  tdvc::launch_k(avgpool_bwd_k, ew_blocks((long long)BC * Tin), 256, 0, (cudaStream_t)stream, dy, dx, BC, Tin, Tout);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_select_channel_fwd(const float* x, const int64_t* label, float* y, int B, int C, int T,
                                       void* stream) {
  TDVC_CHECK_ARG(B >= 0 && C > 0 && T > 0 && x && label && y);
  if (B == 0) return TDVC_OK;
  tdvc::launch_k(select_fwd_k, ew_blocks((long long)B * T), 256, 0, (cudaStream_t)stream, x, label, y, B, C, T);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_select_channel_bwd(const float* dy, const int64_t* label, float* dx, int B, int C, int T,
                                       void* stream) {
  TDVC_CHECK_ARG(B >= 0 && C > 0 && T > 0 && dy && label && dx);
  if (B == 0) return TDVC_OK;
  tdvc::launch_k(select_bwd_k, ew_blocks((long long)B * C * T), 256, 0, (cudaStream_t)stream, dy, label, dx, B, C, T);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_conv1d_select_fwd(const float* x, const float* w, const int64_t* label, float* y, int B, int C, int T,
                                      int NC, int K, int pad, void* stream) {
  TDVC_CHECK_ARG(B >= 0 && C > 0 && T > 0 && NC > 0 && (K == 1 || K == 3 || K == 5 || K == 7) && 2 * pad == K - 1);
  TDVC_CHECK_ARG(x && w && label && y);
  if (B == 0) return TDVC_OK;
  const dim3 grid((unsigned)cdiv(T, 32), (unsigned)B);
  cudaStream_t st = (cudaStream_t)stream;
  switch (K) {
    case 1: tdvc::launch_k(select_conv_fwd_k<1>, grid, 1024, 0, st, x, w, label, y, C, T, NC); break;
    case 3: tdvc::launch_k(select_conv_fwd_k<3>, grid, 1024, 0, st, x, w, label, y, C, T, NC); break;
    case 5: tdvc::launch_k(select_conv_fwd_k<5>, grid, 1024, 0, st, x, w, label, y, C, T, NC); break;
    default: tdvc::launch_k(select_conv_fwd_k<7>, grid, 1024, 0, st, x, w, label, y, C, T, NC); break;
  }
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_conv1d_select_bwd(const float* dy, const float* x, const float* w, const int64_t* label, float* dx,
                                      float* dw, int B, int C, int T, int NC, int K, int pad, void* stream) {
  TDVC_CHECK_ARG(B >= 0 && C > 0 && T > 0 && NC > 0 && (K == 1 || K == 3 || K == 5 || K == 7) && 2 * pad == K - 1);
  TDVC_CHECK_ARG(dy && x && w && label && (dx || dw));
  if (B == 0) return TDVC_OK;
  const long long rows = (long long)B * C;
  const int blocks = (int)std::max<long long>(1, std::min<long long>((rows + 7) / 8, 16LL * num_sms()));
  cudaStream_t st = (cudaStream_t)stream;
  switch (K) {
    case 1: tdvc::launch_k(select_conv_bwd_k<1>, blocks, 256, 0, st, dy, x, w, label, dx, dw, B, C, T, NC); break;
    case 3: tdvc::launch_k(select_conv_bwd_k<3>, blocks, 256, 0, st, dy, x, w, label, dx, dw, B, C, T, NC); break;
    case 5: tdvc::launch_k(select_conv_bwd_k<5>, blocks, 256, 0, st, dy, x, w, label, dx, dw, B, C, T, NC); break;
    default: tdvc::launch_k(select_conv_bwd_k<7>, blocks, 256, 0, st, dy, x, w, label, dx, dw, B, C, T, NC); break;
  }
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_sq_err_const_sum(const float* a, float target, float scale, float* out_sum, int64_t n,
                                     void* stream) {
  TDVC_CHECK_ARG(n >= 0 && out_sum);
  if (n == 0) return TDVC_OK;
  TDVC_CHECK_ARG(a);
  int blocks = (int)std::min<long long>((n + 1023) / 1024, 2LL * num_sms());
  tdvc::launch_k(sq_err_const_sum_k, blocks, 256, 0, (cudaStream_t)stream, a, target, scale, out_sum, n);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_sq_err_const_bwd(const float* a, float target, float scale, const float* gscale, float* da,
                                     int64_t n, void* stream) {
  TDVC_CHECK_ARG(n >= 0);
  if (n == 0) return TDVC_OK;
  TDVC_CHECK_ARG(a && gscale && da);
  tdvc::launch_k(sq_err_const_bwd_k, ew_blocks(n), 256, 0, (cudaStream_t)stream, a, target, scale, gscale, da, n);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_abs_diff_sum(const float* a, const float* b, float scale, float* out_sum, int64_t n,
                                 void* stream) {
  TDVC_CHECK_ARG(n >= 0 && out_sum);
  if (n == 0) return TDVC_OK;
  TDVC_CHECK_ARG(a && b && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0));
  int blocks = (int)std::min<long long>((n + 2047) / 2048, 4LL * num_sms());
  tdvc::launch_k(abs_diff_sum_k, blocks, 256, 0, (cudaStream_t)stream, a, b, scale, out_sum, n);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_abs_diff_bwd(const float* a, const float* b, float scale, const float* gscale, float* da,
                                 int64_t n, void* stream) {
  TDVC_CHECK_ARG(n >= 0);
  if (n == 0) return TDVC_OK;
  TDVC_CHECK_ARG(a && b && gscale && da);
  tdvc::launch_k(abs_diff_bwd_k, ew_blocks(n), 256, 0, (cudaStream_t)stream, a, b, scale, gscale, da, n);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

static int fill_l1_jobs(L1Jobs& J, const tdvc_l1_job* jobs, int n_jobs, bool bwd, long long* max_n) {
  TDVC_CHECK_ARG(jobs && n_jobs >= 1 && n_jobs <= TDVC_L1_MAX_JOBS);
  *max_n = 0;
  for (int j = 0; j < n_jobs; ++j) {
    TDVC_CHECK_ARG(jobs[j].n >= 0 && jobs[j].a && jobs[j].b && ((uintptr_t)jobs[j].a % 16 == 0) && ((uintptr_t)jobs[j].b % 16 == 0));
    if (bwd) TDVC_CHECK_ARG(jobs[j].da && ((uintptr_t)jobs[j].da % 16 == 0));
    J.a[j] = jobs[j].a; J.b[j] = jobs[j].b; J.da[j] = jobs[j].da; J.n[j] = jobs[j].n; J.scale[j] = jobs[j].scale;
    *max_n = std::max<long long>(*max_n, jobs[j].n);
  }
  return TDVC_OK;
}

extern "C" int tdvc_abs_diff_sum_multi(const tdvc_l1_job* jobs, int n_jobs, float* out_sum, void* stream) {
  TDVC_CHECK_ARG(out_sum);
  L1Jobs J{};
  long long max_n = 0;
  int rc = fill_l1_jobs(J, jobs, n_jobs, false, &max_n);
  if (rc) return rc;
  if (max_n == 0) return TDVC_OK;
  // enough blocks for the largest map to stream at full rate; small maps leave most of their row of blocks idle
  int bx = (int)std::max<long long>(1, std::min<long long>((max_n + 4095) / 4096, (4LL * num_sms() + n_jobs - 1) / n_jobs));
  tdvc::launch_k(abs_diff_sum_multi_k, dim3(bx, n_jobs), 256, 0, (cudaStream_t)stream, J, out_sum);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_abs_diff_bwd_multi(const tdvc_l1_job* jobs, int n_jobs, const float* gscale, void* stream) {
  TDVC_CHECK_ARG(gscale);
  L1Jobs J{};
  long long max_n = 0;
  int rc = fill_l1_jobs(J, jobs, n_jobs, true, &max_n);
  if (rc) return rc;
  if (max_n == 0) return TDVC_OK;
  int bx = (int)std::max<long long>(1, std::min<long long>((max_n + 4095) / 4096, (8LL * num_sms() + n_jobs - 1) / n_jobs));
  tdvc::launch_k(abs_diff_bwd_multi_k, dim3(bx, n_jobs), 256, 0, (cudaStream_t)stream, J, gscale);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_adamw_multi(float* const* params, const float* const* grads, float* const* exp_avg,
                                float* const* exp_avg_sq, const int64_t* sizes, int n_tensors, int64_t max_size,
                                float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                                float grad_scale, float* step_dev, void* stream) {
  TDVC_CHECK_ARG(n_tensors >= 0 && (step >= 1 || step_dev) && max_size >= 0);
  if (n_tensors == 0 || max_size == 0) return TDVC_OK;
  TDVC_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && sizes);
  float bc1 = 1.f - powf(beta1, (float)std::max(step, 1));
  float bc2 = 1.f - powf(beta2, (float)std::max(step, 1));
  if (step_dev) {
    tdvc::launch_k(step_inc_k, 1, 1, 0, (cudaStream_t)stream, step_dev);
    TDVC_LAUNCH_CHECK();
  }
  int chunks = (int)std::min<long long>((max_size + 256 * 8 - 1) / (256 * 8), 64);
  if (chunks < 1) chunks = 1;
  tdvc::launch_k(adamw_multi_k, n_tensors * chunks, 256, 0, (cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq, sizes, chunks,
                                                                      lr, beta1, beta2, eps, weight_decay, bc1,
                                                                      sqrtf(bc2), grad_scale, step_dev);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_adamw_blocks(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                                 const int64_t* sizes, const int32_t* blocks, int n_blocks, float lr, float beta1, float beta2,
                                 float eps, float weight_decay, int step, float grad_scale, float* step_dev, void* stream) {
  TDVC_CHECK_ARG(n_blocks >= 0 && (step >= 1 || step_dev));
  if (n_blocks == 0) return TDVC_OK;
  TDVC_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && sizes && blocks && ((uintptr_t)blocks % 8 == 0));
  const float bc1 = 1.f - powf(beta1, (float)std::max(step, 1));
  const float bc2 = 1.f - powf(beta2, (float)std::max(step, 1));
  if (step_dev) {
    tdvc::launch_k(step_inc_k, 1, 1, 0, (cudaStream_t)stream, step_dev);
    TDVC_LAUNCH_CHECK();
  }
  tdvc::launch_k(adamw_blocks_k, n_blocks, 256, 0, (cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq, sizes,
                 reinterpret_cast<const int2*>(blocks), lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_scale, step_dev);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_space_to_depth(const float* x, float* out, int B, int C, int T, int s, int pad, int Tq, void* stream) {
  TDVC_CHECK_ARG(x && out && B >= 0 && C > 0 && T > 0 && s >= 1 && s <= S2D_MAX_S && pad >= 0 && Tq > 0 && C <= 65535 && B <= 65535);
  if (B == 0) return TDVC_OK;
  tdvc::launch_k(s2d_k, dim3(cdiv(Tq, S2D_TT), C, B), 128, 0, (cudaStream_t)stream, x, out, C, T, s, pad, Tq);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_depth_to_space(const float* in, float* y, int B, int C, int Tq, int s, int pad, int Tout, void* stream) {
  TDVC_CHECK_ARG(in && y && B >= 0 && C > 0 && Tq > 0 && s >= 1 && s <= S2D_MAX_S && pad >= 0 && Tout > 0 && C <= 65535 && B <= 65535);
  if (B == 0) return TDVC_OK;
  // every output sample must be covered by a block: frames up to (Tout - 1 + pad) / s
  const int frames = std::max(Tq, (Tout - 1 + pad) / s + 1);
  tdvc::launch_k(d2s_k, dim3(cdiv(frames, S2D_TT), C, B), 128, 0, (cudaStream_t)stream, in, y, C, Tq, s, pad, Tout);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

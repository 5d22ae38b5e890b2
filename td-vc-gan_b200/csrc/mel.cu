// Log-mel spectrogram loss (reference util/losses.py:28-53: torchaudio MelSpectrogram(n_fft, hop = n_fft/4, hann, power 2,
// center / reflect) -> log(clamp(., 1e-5)) -> L1) without cuFFT: the short-time Fourier transform of 18 frames is a GEMM
// against a (window x cos | -sin) basis, which runs on the library's own convolution kernels (tcgen05 in bf16 mode); the
// mel projection is a second GEMM.  This file holds the bandwidth-trivial pieces around the two GEMMs:
//   stft_frames_{fwd,bwd}   framing with reflect padding and the window, as the GEMM's [n_fft][B*frames] operand, and its
//                           adjoint (overlap-add + reflect fold, written as a gather: no atomics)
//   power_{fwd,bwd}         |X|^2 from the stacked (re | im) rows
//   log_clamp_{fwd,bwd}     log(max(x, floor))
#include <cuda_bf16.h>
#include "common.cuh"

namespace tdvc {

__device__ __forceinline__ int reflect_index(int v, int pad, int T) {      // padded position -> sample index
  int u = v - pad;
  if (u < 0) u = -u;
  if (u >= T) u = 2 * (T - 1) - u;
  return u;
}

// F[k][b*NF + n] = v = win[k] * x[b][reflect(n*hop + k - pad)].
// split != 0 (bf16 tensor-core GEMM with fp32-class accuracy): three row blocks [hi | lo | hi] with hi = bf16(v), lo = bf16(v - hi)
// -- against the basis blocks [B_hi | B_hi | B_lo] the GEMM yields B_hi*(hi + lo) + B_lo*hi, i.e. every product term of
// (B_hi + B_lo) * (hi + lo) but the 2^-16-relative lo*lo one.
__global__ void stft_frames_fwd_k(const float* __restrict__ x, const float* __restrict__ win, float* __restrict__ F, int B, int T,
                                  int n_fft, int hop, int pad, int NF, int split) {
  pdl_prologue();
  const long long cols = (long long)B * NF, n = cols * n_fft;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % cols), k = (int)(i / cols);
    const int b = col / NF, f = col - b * NF;
    const float v = __ldg(win + k) * __ldg(x + (long long)b * T + reflect_index(f * hop + k, pad, T));
    if (!split) {
      F[i] = v;
    } else {
      const float hi = __bfloat162float(__float2bfloat16(v));
      const float lo = __bfloat162float(__float2bfloat16(v - hi));
      F[i] = hi;
      F[i + n] = lo;
      F[i + 2 * n] = hi;
    }
  }
}

// dx[b][u] = sum over the padded positions v that read sample u, and the frames n that contain v, of win[v - n*hop] * dF[v - n*hop][b*NF + n]
// split != 0: dF has the three row blocks of the forward; hi (blocks 0 and 2) carries the derivative (rounding is treated as
// the identity), lo = v - hi carries none.
__global__ void stft_frames_bwd_k(const float* __restrict__ dF, const float* __restrict__ win, float* __restrict__ dx, int B,
                                  int T, int n_fft, int hop, int pad, int NF, int split) {
  pdl_prologue();
  const long long cols = (long long)B * NF, n = (long long)B * T;
  const long long blk2 = 2LL * n_fft * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i % T), b = (int)(i / T);
    int vs[3];
    int nv = 0;
    vs[nv++] = u + pad;
    if (u >= 1 && u <= pad) vs[nv++] = pad - u;                                  // left reflection
    if (u <= T - 2 && u >= T - 1 - pad) vs[nv++] = pad + 2 * (T - 1) - u;        // right reflection
    float acc = 0.f;
    for (int j = 0; j < nv; ++j) {
      const int v = vs[j];
      int n_hi = v / hop;
      if (n_hi > NF - 1) n_hi = NF - 1;
      int n_lo = (v - n_fft + hop) / hop;          // ceil((v - n_fft + 1) / hop) for v - n_fft + 1 > 0
      if (v - n_fft + 1 <= 0) n_lo = 0;
      for (int f = n_lo; f <= n_hi; ++f) {
        const int k = v - f * hop;
        if (k >= 0 && k < n_fft) {
          const long long o = (long long)k * cols + (long long)b * NF + f;
          float g = __ldg(dF + o);
          if (split) g += __ldg(dF + o + blk2);
          acc = fmaf(__ldg(win + k), g, acc);
        }
      }
    }
    dx[i] = acc;
  }
}

// S[2*nf_pad ...]: rows [0, nfreq) = real parts, rows [im_off, im_off + nfreq) = imaginary parts, `cols` columns each
// split != 0: P has three row blocks [hi | lo | hi] of the power (bf16 high / low parts) for the mel projection as a bf16 GEMM
__global__ void power_fwd_k(const float* __restrict__ S, float* __restrict__ P, int nfreq, int im_off, long long cols, int split) {
  pdl_prologue();
  const long long n = (long long)nfreq * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float re = S[i], im = S[i + (long long)im_off * cols];
    const float v = fmaf(re, re, im * im);
    if (!split) {
      P[i] = v;
    } else {
      const float hi = __bfloat162float(__float2bfloat16(v));
      P[i] = hi;
      P[i + n] = __bfloat162float(__float2bfloat16(v - hi));
      P[i + 2 * n] = hi;
    }
  }
}

__global__ void power_bwd_k(const float* __restrict__ S, const float* __restrict__ dP, float* __restrict__ dS, int nfreq,
                            int im_off, int rows_total, long long cols, int split) {
  pdl_prologue();
  const long long n = (long long)rows_total * cols;
  const long long blk2 = 2LL * nfreq * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols);
    const long long c = i - (long long)r * cols;
    int f = -1;
    if (r < nfreq) f = r;
    else if (r >= im_off && r < im_off + nfreq) f = r - im_off;
    float g = 0.f;
    if (f >= 0) {
      float d = dP[(long long)f * cols + c];
      if (split) d += dP[(long long)f * cols + c + blk2];      // hi blocks carry the derivative, lo = v - hi none
      g = 2.f * S[i] * d;
    }
    dS[i] = g;       // padding rows of the GEMM output carry no gradient
  }
}

__global__ void log_clamp_fwd_k(const float* __restrict__ x, float* __restrict__ y, long long n, float floor_) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = logf(fmaxf(x[i], floor_));
}

__global__ void log_clamp_bwd_k(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, long long n,
                                float floor_) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    dx[i] = v > floor_ ? dy[i] / v : 0.f;        // torch.clamp passes no gradient below (or at) the floor
  }
}

static int blocks_for(long long n) { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 8LL * num_sms())); }

}  // namespace tdvc
using namespace tdvc;

extern "C" int tdvc_stft_frames_fwd(const float* x, const float* win, float* F, int B, int T, int n_fft, int hop, int pad,
                                    int NF, int split, void* stream) {
  TDVC_CHECK_ARG(x && win && F && B >= 0 && T > 1 && n_fft > 0 && hop > 0 && pad >= 0 && pad < T && NF > 0);
  TDVC_CHECK_ARG((NF - 1) * hop + n_fft <= T + 2 * pad);
  if (B == 0) return TDVC_OK;
  tdvc::launch_k(stft_frames_fwd_k, blocks_for((long long)B * NF * n_fft), 256, 0, (cudaStream_t)stream, x, win, F, B, T, n_fft,
                 hop, pad, NF, split);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_stft_frames_bwd(const float* dF, const float* win, float* dx, int B, int T, int n_fft, int hop, int pad,
                                    int NF, int split, void* stream) {
  TDVC_CHECK_ARG(dF && win && dx && B >= 0 && T > 1 && n_fft > 0 && hop > 0 && pad >= 0 && pad < T && NF > 0);
  if (B == 0) return TDVC_OK;
  tdvc::launch_k(stft_frames_bwd_k, blocks_for((long long)B * T), 256, 0, (cudaStream_t)stream, dF, win, dx, B, T, n_fft, hop, pad,
                 NF, split);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_power_fwd(const float* S, float* P, int nfreq, int im_off, int64_t cols, int split, void* stream) {
  TDVC_CHECK_ARG(S && P && nfreq > 0 && im_off >= nfreq && cols >= 0);
  if (cols == 0) return TDVC_OK;
  tdvc::launch_k(power_fwd_k, blocks_for((long long)nfreq * cols), 256, 0, (cudaStream_t)stream, S, P, nfreq, im_off, (long long)cols, split);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_power_bwd(const float* S, const float* dP, float* dS, int nfreq, int im_off, int rows_total, int64_t cols,
                              int split, void* stream) {
  TDVC_CHECK_ARG(S && dP && dS && nfreq > 0 && im_off >= nfreq && rows_total >= im_off + nfreq && cols >= 0);
  if (cols == 0) return TDVC_OK;
  tdvc::launch_k(power_bwd_k, blocks_for((long long)rows_total * cols), 256, 0, (cudaStream_t)stream, S, dP, dS, nfreq, im_off,
                 rows_total, (long long)cols, split);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_log_clamp_fwd(const float* x, float* y, int64_t n, float floor_, void* stream) {
  TDVC_CHECK_ARG(x && y && n >= 0 && floor_ > 0.f);
  if (n == 0) return TDVC_OK;
  tdvc::launch_k(log_clamp_fwd_k, blocks_for(n), 256, 0, (cudaStream_t)stream, x, y, (long long)n, floor_);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_log_clamp_bwd(const float* x, const float* dy, float* dx, int64_t n, float floor_, void* stream) {
  TDVC_CHECK_ARG(x && dy && dx && n >= 0 && floor_ > 0.f);
  if (n == 0) return TDVC_OK;
  tdvc::launch_k(log_clamp_bwd_k, blocks_for(n), 256, 0, (cudaStream_t)stream, x, dy, dx, (long long)n, floor_);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

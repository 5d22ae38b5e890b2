// YIN pitch estimation on the device: util/yin.py:24-140 (`estimate` with the hard threshold search), the F0 producer named
// by BASELINE config 4 / SURVEY.md section 8f row 3.
//
// Per frame of W = 2 * tau_max samples (zero padded W/2 to the left, W/2 - 1 to the right, hop `stride`):
//   r(tau)   = sum_{j < W - tau} x_j x_{j+tau}                                  (the reference gets it from an FFT)
//   d(tau)   = sum_all x^2 + sum_{tau <= j < W - tau} x_j^2 - 2 r(tau)          (= sum_{j < W - tau} (x_j - x_{j+tau})^2, eq. 6)
//   c(tau)   = d(tau) * tau / max(sum_{k=1..tau} d(k), 1e-5),  tau = 1 .. tau_max - 1        (eq. 8)
//   search over i = tau - tau_min - 1 >= 0: first i with c < threshold (none, or i = 0: unvoiced -- the reference's
//   `where(first_below > 0, ...)`), then the first i' >= i where c stops decreasing (or the last one);
//   f0 = (1 / (i' + tau_min + 1)) * sample_rate if i' > 0 else 0.
//
// One CTA per (frame, signal).  The sums are formed directly in fp64 (0.15 M products per frame: nothing next to the training
// step), c is rounded to fp32 before the comparisons because that is the precision the reference decides in.  The reference's
// fp32 FFT carries ~1e-6-relative noise in d, so a frame whose c touches the threshold or whose minimum is flat to that level
// can land one candidate earlier or later there; tests state the agreement they require.
#include "common.cuh"

namespace tdvc {

__global__ void __launch_bounds__(256) yin_k(const float* __restrict__ x, float* __restrict__ f0, int T, int n_frames, int W,
                                             int stride, int tau_min, int tau_max, float threshold, float sample_rate) {
  pdl_prologue();
  extern __shared__ double smd[];
  double* sq = smd;                      // [W + 1] prefix sums of x^2
  double* d = sq + (W + 1);              // [tau_max] difference function
  float* xs = reinterpret_cast<float*>(d + tau_max);      // [W]
  float* c = xs + W;                     // [tau_max] cumulative-mean-normalised difference, index tau - 1
  const int fr = blockIdx.x, b = blockIdx.y;
  const float* xb = x + (long long)b * T;
  const int s0 = fr * stride - W / 2;
  for (int j = threadIdx.x; j < W; j += 256) {
    const int s = s0 + j;
    xs[j] = (s >= 0 && s < T) ? __ldg(xb + s) : 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    sq[0] = 0.0;
    for (int j = 0; j < W; ++j) { a += (double)xs[j] * (double)xs[j]; sq[j + 1] = a; }
  }
  __syncthreads();
  for (int tau = threadIdx.x; tau < tau_max; tau += 256) {
    double r = 0.0;
    const int n = W - tau;
    for (int j = 0; j < n; ++j) r += (double)xs[j] * (double)xs[j + tau];
    d[tau] = sq[W] + (sq[W - tau] - sq[tau]) - 2.0 * r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double cum = 0.0;
    for (int tau = 1; tau < tau_max; ++tau) {
      cum += d[tau];
      c[tau - 1] = (float)(d[tau] * (double)tau / fmax(cum, 1e-5));
    }
    const float* cs = c + tau_min;       // the reference slices [..., tau_min:]
    const int n = tau_max - 1 - tau_min;
    int first = 0;
    for (int i = 0; i < n; ++i)
      if (cs[i] < threshold) { first = i; break; }
    int tau = 0;
    if (first > 0) {
      tau = n - 1;
      for (int i = first; i < n - 1; ++i)
        if (cs[i + 1] - cs[i] >= 0.f) { tau = i; break; }
    }
    // the reference's `sample_rate / tensor` is reciprocal(tensor) * sample_rate (Tensor.__rtruediv__): two fp32 roundings
    f0[(long long)b * n_frames + fr] = tau > 0 ? __fmul_rn(__frcp_rn((float)(tau + tau_min + 1)), sample_rate) : 0.f;
  }
}

}  // namespace tdvc
using namespace tdvc;

extern "C" int tdvc_yin_estimate(const float* signal, float* f0, int B, int T, int n_frames, int frame_length, int frame_stride,
                                 int tau_min, int tau_max, float threshold, float sample_rate, void* stream) {
  TDVC_CHECK_ARG(signal && f0 && B >= 0 && T > 0 && n_frames > 0 && frame_stride > 0 && tau_min >= 0 && tau_max > tau_min + 2);
  TDVC_CHECK_ARG(frame_length == 2 * tau_max && B <= 65535);
  if (B == 0) return TDVC_OK;
  const size_t smem = sizeof(double) * (size_t)(frame_length + 1 + tau_max) + sizeof(float) * (size_t)(frame_length + tau_max);
  TDVC_CHECK_ARG(smem <= 200 * 1024);
  if (smem > 48 * 1024) TDVC_CUDA(cudaFuncSetAttribute(yin_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  tdvc::launch_k(yin_k, dim3(n_frames, B), 256, smem, (cudaStream_t)stream, signal, f0, T, n_frames, frame_length, frame_stride,
                 tau_min, tau_max, threshold, sample_rate);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

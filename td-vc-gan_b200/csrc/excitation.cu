// F0 track -> sine + noise excitation on the device: util/__init__.py:22-50 (f0_to_excitation), the producer of the decoder's
// c_var conditioning.
//
//   f0[rows, F] (Hz, 0 = unvoiced), the last frame is dropped; omega_f = 2 pi f0_f / sr (fp32, as the reference computes it);
//   per sample t < N = (F-1)*step: nearest = omega[t / step]; linear = F.interpolate(mode='linear', align_corners=False) of omega;
//   the linear value is taken where the same interpolation of log(omega) is not -inf (i.e. no unvoiced neighbour frame carries
//   weight), else the held one;
//   phase = inclusive cumsum over t;  voiced: 0.1 sin(phase + phase0) + 0.003 n_v[t];  unvoiced (omega == 0): n_u * 0.003 * gain,
//   gain = 0.1 / (3 * 0.003).
//
// One CTA per row: every thread owns a contiguous run of samples, sums its angular frequencies in fp64, the CTA scans the
// thread totals, then each thread walks its run again.  The reference accumulates the phase in fp32 (torch.cumsum, ~1e3 rad at
// the end of a 0.56 s segment: an fp32 scan is good to ~1e-4 rad there); the fp64 scan here is exact to rounding of the final
// phase, so the two agree to the reference's own error, not bit for bit -- tests hold this kernel against an fp64 evaluation.
//
// Random numbers are inputs (drawn by the caller from torch's generator in the reference's order), so a seeded run consumes
// the generator exactly as the reference does: noise_v[rows, N]; noise_u either full [rows, N] (no host synchronisation) or
// compact -- the k-th unvoiced sample in row-major order takes noise_u[k], as `excitation[unvoiced] = randn(count)` does --
// with u_offset[row] = number of unvoiced samples in the rows before (tdvc_f0_unvoiced_count gives the per-row counts).
#include "common.cuh"

namespace tdvc {

constexpr int EXC_THREADS = 256;

// 2*pi*f0/sr as the reference evaluates it on an fp32 tensor: (fp32(2 pi) * f0) / sr, two roundings
__device__ __forceinline__ float exc_w(float f0, float sr) { return __fdiv_rn(__fmul_rn(6.283185307179586f, f0), sr); }

// per-sample angular frequency exactly as the reference's three F.interpolate calls + masked assignment produce it
__device__ __forceinline__ float exc_omega(const float* __restrict__ f0row, int nf, int t, int step, float sr,
                                           int linear) {
  const int fi = t / step;
  const float held = exc_w(__ldg(f0row + fi), sr);
  if (!linear) return held;
  // area_pixel_compute_source_index(scale = 1/step, t, align_corners = false): scale * (t + 0.5) - 0.5, clamped at 0
  float x = (1.0f / (float)step) * ((float)t + 0.5f) - 0.5f;
  if (x < 0.f) x = 0.f;
  const int i0 = (int)x;
  const int i1 = i0 + (i0 < nf - 1 ? 1 : 0);
  const float l1 = x - (float)i0, l0 = 1.f - l1;
  const float v0 = exc_w(__ldg(f0row + i0), sr), v1 = exc_w(__ldg(f0row + i1), sr);
  const float lin = l0 * v0 + l1 * v1;
  const float lg = l0 * logf(v0) + l1 * logf(v1);      // log(0) = -inf; 0 * -inf = NaN != -inf, as in the reference
  return (lg != -INFINITY) ? lin : held;
}

__global__ void __launch_bounds__(EXC_THREADS) f0_unvoiced_count_k(const float* __restrict__ f0, int* __restrict__ counts, int F,
                                                                   int step, float sr, int linear) {
  pdl_prologue();
  __shared__ float sm[33];
  const int row = blockIdx.x, nf = F - 1, N = nf * step;
  const float* f0row = f0 + (long long)row * F;
  float c = 0.f;
  for (int t = threadIdx.x; t < N; t += EXC_THREADS) c += exc_omega(f0row, nf, t, step, sr, linear) == 0.f ? 1.f : 0.f;
  c = block_sum(c, sm);
  if (threadIdx.x == 0) counts[row] = (int)(c + 0.5f);
}

__global__ void __launch_bounds__(EXC_THREADS) f0_excitation_k(const float* __restrict__ f0, const float* __restrict__ noise_v,
                                                               const float* __restrict__ noise_u,
                                                               const long long* __restrict__ u_offset,
                                                               const float* __restrict__ phase0, float* __restrict__ out, int F,
                                                               int step, float sr, int linear) {
  pdl_prologue();
  __shared__ double s_sum[EXC_THREADS];
  __shared__ int s_cnt[EXC_THREADS];
  const int row = blockIdx.x, nf = F - 1, N = nf * step;
  const float* f0row = f0 + (long long)row * F;
  const int per = (N + EXC_THREADS - 1) / EXC_THREADS;
  const int t_lo = min(N, threadIdx.x * per), t_hi = min(N, t_lo + per);
  double sum = 0.0;
  int cnt = 0;
  for (int t = t_lo; t < t_hi; ++t) {
    const float w = exc_omega(f0row, nf, t, step, sr, linear);
    sum += (double)w;
    cnt += w == 0.f;
  }
  s_sum[threadIdx.x] = sum;
  s_cnt[threadIdx.x] = cnt;
  __syncthreads();
  // exclusive scan of the 256 thread totals (Hillis-Steele on a copy; 8 rounds)
  for (int off = 1; off < EXC_THREADS; off <<= 1) {
    double a = 0.0;
    int c = 0;
    if ((int)threadIdx.x >= off) { a = s_sum[threadIdx.x - off]; c = s_cnt[threadIdx.x - off]; }
    __syncthreads();
    s_sum[threadIdx.x] += a;
    s_cnt[threadIdx.x] += c;
    __syncthreads();
  }
  double phase = s_sum[threadIdx.x] - sum;                   // exclusive prefix
  long long uidx = (long long)(s_cnt[threadIdx.x] - cnt) + (u_offset ? u_offset[row] : 0);
  const float p0 = __ldg(phase0);
  const float sin_gain = 0.1f, noise_std = 0.003f, noise_gain = (float)(0.1 / (3 * 0.003));
  const long long base = (long long)row * N;
  for (int t = t_lo; t < t_hi; ++t) {
    const float w = exc_omega(f0row, nf, t, step, sr, linear);
    phase += (double)w;
    float e;
    if (w == 0.f) {
      const float n = u_offset ? __ldg(noise_u + uidx) : __ldg(noise_u + base + t);
      e = n * noise_std * noise_gain;                        // two fp32 multiplications, in the reference's order
      ++uidx;
    } else {
      // the reference adds the start phase to the fp32 phase tensor; here the sum is formed in fp64 and rounded once
      e = sin_gain * (float)sin(phase + (double)p0) + __ldg(noise_v + base + t) * noise_std;
    }
    out[base + t] = e;
  }
}

}  // namespace tdvc
using namespace tdvc;

extern "C" int tdvc_f0_unvoiced_count(const float* f0, int* counts, int rows, int F, int step, float sampling_rate, int linear,
                                      void* stream) {
  TDVC_CHECK_ARG(f0 && counts && rows >= 0 && F >= 2 && step >= 1 && sampling_rate > 0.f);
  if (rows == 0) return TDVC_OK;
  tdvc::launch_k(f0_unvoiced_count_k, rows, EXC_THREADS, 0, (cudaStream_t)stream, f0, counts, F, step, sampling_rate, linear);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_f0_excitation(const float* f0, const float* noise_v, const float* noise_u, const int64_t* u_offset,
                                  const float* phase0, float* out, int rows, int F, int step, float sampling_rate, int linear,
                                  void* stream) {
  TDVC_CHECK_ARG(f0 && noise_v && noise_u && phase0 && out && rows >= 0 && F >= 2 && step >= 1 && sampling_rate > 0.f);
  if (rows == 0) return TDVC_OK;
  tdvc::launch_k(f0_excitation_k, rows, EXC_THREADS, 0, (cudaStream_t)stream, f0, noise_v, noise_u, (const long long*)u_offset,
                 phase0, out, F, step, sampling_rate, linear);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

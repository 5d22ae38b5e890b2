// Fused contrastive (InfoNCE) loss of util/losses.py:70-116: gather of in-utterance negatives + cosine
// similarities + cross-entropy (target = the positive) in one kernel, forward value and unit gradients
// together.  Tiny problem ([B,128,28] embeddings, 100 negatives): one CTA per (direction, batch, frame),
// one warp per candidate (lanes own channel quads), warp-shuffle dot products, softmax in shared memory.
#include "common.cuh"

namespace tdvc {

constexpr int CL_MAXN = 256;   // candidates (1 positive + negatives) supported per frame

// A: anchor signal, P: positive signal, both [B,C,T]; raw[B,T,N]: torch.randint(0, T-1) draws (self skipped below).
// loss_sum += scale * CE(frame);  dA[b,:,t] += scale * dCE/dA,  dP[b,:,t] += scale * dCE/dP  (negatives carry no
// gradient: they are gathered under no_grad in the reference).
__global__ void __launch_bounds__(128) contrastive_k(const float* __restrict__ A, const float* __restrict__ P,
                                                     const int64_t* __restrict__ raw, float* __restrict__ loss_sum,
                                                     float* __restrict__ dA, float* __restrict__ dP, int C, int T, int N,
                                                     float scale) {
  pdl_prologue();
  __shared__ float logit[CL_MAXN];
  __shared__ float vnorm[CL_MAXN];
  __shared__ int vidx[CL_MAXN];
  __shared__ float red[8];
  const int b = blockIdx.x / T, t = blockIdx.x - b * T;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const float* Ab = A + (long long)b * C * T;
  const float* Pb = P + (long long)b * C * T;
  const float eps = 1e-8f;
  // ||u||
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += 128) { float u = Ab[(long long)c * T + t]; s = fmaf(u, u, s); }
  s = warp_sum(s);
  if (lane == 0) red[wrp] = s;
  __syncthreads();
  const float un = fmaxf(sqrtf(red[0] + red[1] + red[2] + red[3]), eps);
  // candidate time indices: 0 -> positive (same frame of P), 1+n -> negative frame of A (skip self)
  for (int n = threadIdx.x; n <= N; n += 128) {
    int ti = t;
    if (n > 0) {
      long long r = raw[((long long)b * T + t) * N + (n - 1)];
      ti = (int)(r >= t ? r + 1 : r);
    }
    vidx[n] = ti;
  }
  __syncthreads();
  // cosine similarities: one warp per candidate
  for (int n = wrp; n <= N; n += 4) {
    const float* V = (n == 0 ? Pb : Ab) + vidx[n];
    float dot = 0.f, vv = 0.f;
    for (int c = lane; c < C; c += 32) {
      float v = V[(long long)c * T], u = Ab[(long long)c * T + t];
      dot = fmaf(u, v, dot);
      vv = fmaf(v, v, vv);
    }
    dot = warp_sum(dot);
    vv = warp_sum(vv);
    if (lane == 0) {
      float vn = fmaxf(sqrtf(vv), eps);
      vnorm[n] = vn;
      logit[n] = dot / (un * vn);
    }
  }
  __syncthreads();
  // softmax over the N+1 candidates (warp 0), CE with target 0
  if (wrp == 0) {
    float m = -1e30f;
    for (int n = lane; n <= N; n += 32) m = fmaxf(m, logit[n]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float z = 0.f;
    for (int n = lane; n <= N; n += 32) z += expf(logit[n] - m);
    z = warp_sum(z);
    if (lane == 0) atomicAdd(loss_sum, scale * (logf(z) + m - logit[0]));
    if (lane == 0) { red[4] = m; red[5] = z; }
  }
  __syncthreads();
  if (!dA) return;
  const float m = red[4], z = red[5];
  // d CE / d logit_n = softmax_n - [n == 0];  d cos / d u = (v_hat - cos * u_hat) / ||u||,  d cos / d v likewise
  for (int c = threadIdx.x; c < C; c += 128) {
    const float u = Ab[(long long)c * T + t];
    const float uh = u / un;
    float du = 0.f;
    for (int n = 0; n <= N; ++n) {
      const float dl = expf(logit[n] - m) / z - (n == 0 ? 1.f : 0.f);
      const float v = (n == 0 ? Pb : Ab)[(long long)c * T + vidx[n]];
      du = fmaf(dl, (v / vnorm[n] - logit[n] * uh), du);
    }
    atomicAdd(dA + ((long long)b * C + c) * T + t, scale * du / un);
    const float dl0 = expf(logit[0] - m) / z - 1.f;
    const float p = Pb[(long long)c * T + t];
    atomicAdd(dP + ((long long)b * C + c) * T + t, scale * dl0 * (uh - logit[0] * p / vnorm[0]) / vnorm[0]);
  }
}

}  // namespace tdvc
using namespace tdvc;

// One direction of util/losses.py:70-116.  loss_sum (1 float) and dA/dP ([B,C,T], may both be NULL) are accumulated
// into: the caller zeroes them and calls twice ((X,Y,raw_X) and (Y,X,raw_Y)) with scale = 1/(2*B*T).
extern "C" int tdvc_contrastive_dir(const float* A, const float* P, const int64_t* raw, float* loss_sum, float* dA,
                                    float* dP, int B, int C, int T, int N, float scale, void* stream) {
  TDVC_CHECK_ARG(A && P && raw && loss_sum && B >= 0 && C > 0 && T > 1 && N > 0 && N + 1 <= CL_MAXN);
  TDVC_CHECK_ARG((dA == nullptr) == (dP == nullptr));
  if (B == 0) return TDVC_OK;
  tdvc::launch_k(contrastive_k, B * T, 128, 0, (cudaStream_t)stream, A, P, raw, loss_sum, dA, dP, C, T, N, scale);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

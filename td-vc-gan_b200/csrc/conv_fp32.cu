// fp32 CUDA-core Conv1d / ConvTranspose1d kernels (the exact-parity path, rel 1e-5 vs the reference).
//
// Three kernels cover every convolution of td-vc-gan's Generator and Discriminators
// (model/generator.py:75-92,146-168,214-249,299-347; model/discriminator.py:17-38,100-102):
//   conv_fwd_k   : direct conv, NCW, any stride/dilation/groups, zero or reflect padding, fused input
//                  LeakyReLU, bias, residual, output activation.  Also used (weight strides swapped,
//                  taps flipped) as the stride-1 dgrad and as ConvTranspose1d's dgrad.
//   conv_tr_k    : transposed ("scatter as gather", polyphase) conv: strided dgrad and ConvTranspose1d fwd.
//   conv_wgrad_k : weight gradient as an implicit-im2col outer-product reduction over (b, t) with
//                  register accumulation and one atomicAdd per output per CTA.
// All are shared-memory tiled, coalesced along time, bank-conflict free for stride 1.
#include <algorithm>
#include "common.cuh"

namespace tdvc {

struct FwdP {
  int B, Cin, Tin, Cout, Tout, K, stride, pad, dil, groups;
  int pad_mode, out_act;
  float in_slope, out_slope;
  int cin_g, cout_g;
  long long w_sco, w_sci, w_sk;
  int w_flip;
  int ci_chunk, span;
  int seg, span_p;   // strided convs keep the input tile de-interleaved by phase: element i -> (i % stride) * seg + i / stride
  int gpb;           // groups per CTA: with few output channels per group (the discriminators' 4-in/4-out groups) a CTA's
                     // channel rows span gpb consecutive groups instead of leaving most of its threads idle
};

__device__ __forceinline__ float fetch_padded(const float* __restrict__ row, int gt, int T, int pad_mode,
                                              float slope) {
  if (gt < 0) {
    if (pad_mode != TDVC_PAD_REFLECT) return 0.f;
    gt = -gt;
    if (gt >= T) return 0.f;
  } else if (gt >= T) {
    if (pad_mode != TDVC_PAD_REFLECT) return 0.f;
    gt = 2 * (T - 1) - gt;
    if (gt < 0) return 0.f;
  }
  float v = __ldg(row + gt);
  return v > 0.f ? v : v * slope;
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == TDVC_ACT_LRELU) return v > 0.f ? v : v * slope;
  if (act == TDVC_ACT_TANH) return tanhf(v);
  return v;
}

// weight tile in shared memory: row r = (ci,k), COB output channels.  Vector (float4) readers use an
// XOR swizzle on 4-channel groups so the transposing store (consecutive r per lane) is at worst 4-way
// conflicted; scalar readers use a +1 padded pitch.
template <int CO_T, int COB>
__device__ __forceinline__ int ws_index(int r, int co) {
  if constexpr (CO_T % 4 == 0) {
    constexpr int M = COB / 4 - 1;
    return r * COB + ((((co >> 2) ^ (r & M)) << 2) | (co & 3));
  } else {
    return r * (COB + 1) + co;
  }
}

// ---------------------------------------------------------------------------------------------
// forward: each thread owns CO_T output channels x 4 time steps (time interleaved by TX so that a
// warp reads consecutive shared-memory words).  block = 256 threads = TX (time) x TY (channels).
template <int CO_T, int TX>
__global__ void __launch_bounds__(256) conv_fwd_k(FwdP p, const float* __restrict__ x, const float* __restrict__ w,
                                                   const float* __restrict__ bias, const float* __restrict__ res,
                                                   float* __restrict__ y) {
  pdl_prologue();
  constexpr int TY = 256 / TX, TT = TX * 4, COB = TY * CO_T;
  extern __shared__ float sm[];
  float* xs = sm;
  float* ws = sm + (((size_t)p.gpb * p.ci_chunk * p.span_p + 3) & ~(size_t)3);
  constexpr int WPITCH_K = (CO_T % 4 == 0) ? COB : COB + 1;
  int* kofs = reinterpret_cast<int*>(ws + (size_t)p.ci_chunk * p.K * WPITCH_K);   // [K] tile offset of tap k
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int t0 = blockIdx.x * TT;
  const int tiles_per_group = (p.cout_g + COB - 1) / COB;
  const int grp = p.gpb > 1 ? blockIdx.y * p.gpb : blockIdx.y / tiles_per_group;     // first group of this CTA
  const int co0 = p.gpb > 1 ? 0 : (blockIdx.y % tiles_per_group) * COB;
  const int co_lim = p.gpb > 1 ? min(p.gpb, p.groups - grp) * p.cout_g : p.cout_g;   // channels (from grp's first) this CTA may own
  const int g_local = p.gpb > 1 ? (ty * CO_T) / p.cout_g : 0;                        // this thread's group inside the bundle
  const int b = blockIdx.z;
  float acc[CO_T][4];
#pragma unroll
  for (int c = 0; c < CO_T; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[c][j] = 0.f;
  const float* xb = x + ((long long)b * p.Cin + (long long)grp * p.cin_g) * p.Tin;
  const int gt0 = t0 * p.stride - p.pad;
  const bool row_active = (co0 + ty * CO_T) < co_lim;
  const int x_groups = p.gpb > 1 ? min(p.gpb, p.groups - grp) : 1;                   // input-channel groups in the tile
  // With the tile stored phase-major, output step (tx + j*TX) and tap k read word  kofs[k] + tx + j*TX : consecutive
  // lanes hit consecutive banks for any stride (a plain layout gives stride-way bank conflicts).
  for (int k = threadIdx.x; k < p.K; k += 256) {
    int e = k * p.dil;
    kofs[k] = (e % p.stride) * p.seg + e / p.stride;
  }

  for (int c0 = 0; c0 < p.cin_g; c0 += p.ci_chunk) {
    const int nci = min(p.ci_chunk, p.cin_g - c0);
    __syncthreads();
    // bundled groups (gpb > 1, then ci_chunk == cin_g): the tile holds the x_groups * cin_g input channels of the bundle
    for (int idx = threadIdx.x; idx < x_groups * nci * p.span; idx += 256) {
      int ci = idx / p.span, i = idx - ci * p.span;
      float v = fetch_padded(xb + (long long)(c0 + ci) * p.Tin, gt0 + i, p.Tin, p.pad_mode, p.in_slope);
      int pos = (p.stride == 1) ? i : (i % p.stride) * p.seg + i / p.stride;
      xs[ci * p.span_p + pos] = v;
    }
    const int nwk = nci * p.K;
    for (int idx = threadIdx.x; idx < COB * nwk; idx += 256) {
      int co = idx / nwk, r = idx - co * nwk;
      int ci = r / p.K, k = r - ci * p.K;
      float v = 0.f;
      if (co0 + co < co_lim) {
        int kk = p.w_flip ? p.K - 1 - k : k;
        v = __ldg(w + (long long)(grp * p.cout_g + co0 + co) * p.w_sco + (long long)(c0 + ci) * p.w_sci +
                  (long long)kk * p.w_sk);
      }
      ws[ws_index<CO_T, COB>(r, co)] = v;
    }
    __syncthreads();
    if (row_active) {
      for (int ci = 0; ci < nci; ++ci) {
        const float* xr = xs + (g_local * p.cin_g + ci) * p.span_p + tx;
        for (int k = 0; k < p.K; ++k) {
          const float* xk = xr + kofs[k];
          float xv0 = xk[0], xv1 = xk[TX], xv2 = xk[2 * TX], xv3 = xk[3 * TX];
          const int r = ci * p.K + k;
          float wv[CO_T];
          if constexpr (CO_T % 4 == 0) {
#pragma unroll
            for (int c = 0; c < CO_T; c += 4) {
              float4 t4 = *reinterpret_cast<const float4*>(ws + ws_index<CO_T, COB>(r, ty * CO_T + c));
              wv[c] = t4.x; wv[c + 1] = t4.y; wv[c + 2] = t4.z; wv[c + 3] = t4.w;
            }
          } else {
#pragma unroll
            for (int c = 0; c < CO_T; ++c) wv[c] = ws[ws_index<CO_T, COB>(r, ty * CO_T + c)];
          }
#pragma unroll
          for (int c = 0; c < CO_T; ++c) {
            acc[c][0] = fmaf(wv[c], xv0, acc[c][0]);
            acc[c][1] = fmaf(wv[c], xv1, acc[c][1]);
            acc[c][2] = fmaf(wv[c], xv2, acc[c][2]);
            acc[c][3] = fmaf(wv[c], xv3, acc[c][3]);
          }
        }
      }
    }
  }
  if (!row_active) return;
#pragma unroll
  for (int c = 0; c < CO_T; ++c) {
    int co = co0 + ty * CO_T + c;
    if (co >= co_lim) break;
    int cog = grp * p.cout_g + co;
    float bv = bias ? __ldg(bias + cog) : 0.f;
    long long base = ((long long)b * p.Cout + cog) * p.Tout;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int t = t0 + tx + j * TX;
      if (t < p.Tout) {
        float v = acc[c][j] + bv;
        if (res) v += __ldg(res + base + t);
        y[base + t] = apply_act(v, p.out_act, p.out_slope);
      }
    }
  }
}

template <int CO_T, int TX>
static int launch_fwd_t(FwdP p, const float* x, const float* w, const float* bias, const float* res, float* y,
                        cudaStream_t st) {
  constexpr int TY = 256 / TX, TT = TX * 4, COB = TY * CO_T;
  p.span = (TT - 1) * p.stride + (p.K - 1) * p.dil + 1;
  p.seg = (p.span + p.stride - 1) / p.stride;
  p.span_p = p.seg * p.stride;
  const size_t budget = 64 * 1024;
  constexpr int WPITCH = (CO_T % 4 == 0) ? COB : COB + 1;
  size_t per_ci = ((size_t)p.span_p + (size_t)p.K * WPITCH) * sizeof(float);
  const size_t extra = 16 + (size_t)p.K * sizeof(int);
  int chunk = (int)((budget - extra) / per_ci);
  if (chunk < 1) chunk = 1;
  if (chunk > p.cin_g) chunk = p.cin_g;
  if (chunk > 32) chunk = 32;
  p.ci_chunk = chunk;
  size_t smem = per_ci * chunk + extra;
  // several groups per CTA when a group has fewer output channels than the CTA has channel rows
  p.gpb = 1;
  if (p.groups > 1 && p.cout_g < COB && COB % p.cout_g == 0 && p.cout_g % CO_T == 0 && chunk == p.cin_g) {
    const int gpb = std::min(COB / p.cout_g, p.groups);
    const size_t smem_b = ((size_t)gpb * p.cin_g * p.span_p + 4 + (size_t)p.cin_g * p.K * WPITCH) * sizeof(float) + extra;
    if (gpb > 1 && smem_b <= budget) { p.gpb = gpb; smem = smem_b; }
  }
  TDVC_CHECK_ARG(smem <= 200 * 1024);
  // the attribute is per device: set on every launch that needs it (cheap) rather than once per process
  if (smem > 48 * 1024) {
    TDVC_CUDA(cudaFuncSetAttribute(conv_fwd_k<CO_T, TX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  dim3 grid(cdiv(p.Tout, TT), p.gpb > 1 ? cdiv(p.groups, p.gpb) : p.groups * cdiv(p.cout_g, COB), p.B);
  TDVC_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535);
  tdvc::launch_k(conv_fwd_k<CO_T, TX>, grid, 256, smem, st, p, x, w, bias, res, y);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

static int launch_fwd(FwdP p, const float* x, const float* w, const float* bias, const float* res, float* y,
                      cudaStream_t st) {
  if (p.B == 0 || p.Tout <= 0) return TDVC_OK;
  g_flops[FLOP_FP32] += 2.0 * p.B * p.Tout * (double)p.Cout * p.cin_g * p.K;
  const bool small_t = p.Tout <= 48;
  const int ty = small_t ? 32 : 8;
  const int tt = small_t ? 32 : 128;
  int co_t = 8;
  if (ty * 1 >= p.cout_g) co_t = 1;
  else if (ty * 2 >= p.cout_g) co_t = 2;
  else if (ty * 4 >= p.cout_g) co_t = 4;
  // short sequences (T/320 stage, discriminator tails) would otherwise launch a handful of CTAs: trade
  // per-thread register tiling for enough CTAs to cover the 148 SMs
  auto n_ctas = [&](int c) { return (long long)cdiv(p.Tout, tt) * p.groups * cdiv(p.cout_g, ty * c) * p.B; };
  while (co_t > 1 && n_ctas(co_t) < num_sms()) co_t >>= 1;
#define TDVC_FWD_CASE(C, X) return launch_fwd_t<C, X>(p, x, w, bias, res, y, st)
  if (small_t) {
    switch (co_t) { case 1: TDVC_FWD_CASE(1, 8); case 2: TDVC_FWD_CASE(2, 8); case 4: TDVC_FWD_CASE(4, 8); default: TDVC_FWD_CASE(8, 8); }
  } else {
    switch (co_t) { case 1: TDVC_FWD_CASE(1, 32); case 2: TDVC_FWD_CASE(2, 32); case 4: TDVC_FWD_CASE(4, 32); default: TDVC_FWD_CASE(8, 32); }
  }
#undef TDVC_FWD_CASE
}

static int check_geom(const tdvc_conv_geom* g, bool transpose) {
  TDVC_CHECK_ARG(g != nullptr);
  TDVC_CHECK_ARG(g->B >= 0 && g->Cin > 0 && g->Cout > 0 && g->Tin > 0 && g->Tout > 0 && g->K > 0);
  TDVC_CHECK_ARG(g->stride > 0 && g->dilation > 0 && g->pad >= 0 && g->groups > 0);
  TDVC_CHECK_ARG(g->Cin % g->groups == 0 && g->Cout % g->groups == 0);
  if (!transpose) {
    long long tout = ((long long)g->Tin + 2LL * g->pad - (long long)g->dilation * (g->K - 1) - 1) / g->stride + 1;
    TDVC_CHECK_ARG(tout == g->Tout);
    if (g->pad_mode == TDVC_PAD_REFLECT) TDVC_CHECK_ARG(g->pad < g->Tin);
  } else {
    TDVC_CHECK_ARG(g->groups == 1 && g->dilation == 1 && g->pad_mode == TDVC_PAD_ZEROS && g->in_slope == 1.0f);
    long long tout_min = ((long long)g->Tin - 1) * g->stride - 2LL * g->pad + (g->K - 1) + 1;
    TDVC_CHECK_ARG(g->Tout >= tout_min && g->Tout < tout_min + g->stride);  // output_padding < stride
  }
  return TDVC_OK;
}

static FwdP fwd_params(const tdvc_conv_geom* g) {
  FwdP p{};
  p.B = g->B; p.Cin = g->Cin; p.Tin = g->Tin; p.Cout = g->Cout; p.Tout = g->Tout; p.K = g->K;
  p.stride = g->stride; p.pad = g->pad; p.dil = g->dilation; p.groups = g->groups;
  p.pad_mode = g->pad_mode; p.out_act = g->out_act; p.in_slope = g->in_slope; p.out_slope = g->out_slope;
  p.cin_g = g->Cin / g->groups; p.cout_g = g->Cout / g->groups;
  p.w_sco = (long long)p.cin_g * g->K; p.w_sci = g->K; p.w_sk = 1; p.w_flip = 0;
  return p;
}

// ---------------------------------------------------------------------------------------------
// transposed conv:  out[b, oc, i] = sum_{ic,k} in[b, ic, t] * w[ic, oc_in_group, k],  t*stride = i + peff - k*dil
struct TrP {
  int B, Cin, Tin, Cout, Tout, K, stride, peff, dil, groups;  // "in"/"out" are this kernel's operands
  int in_g, out_g, ic_chunk, qlen;
};

template <int OC_T>
__global__ void __launch_bounds__(256) conv_tr_k(TrP p, const float* __restrict__ in, const float* __restrict__ w,
                                                  const float* __restrict__ bias, float* __restrict__ out) {
  pdl_prologue();
  constexpr int TU = 128, TY = 2, OB = TY * OC_T;
  extern __shared__ float sm[];
  float* ins = sm;                                   // [ic_chunk][qlen]
  float* ws = sm + (size_t)p.ic_chunk * p.qlen;      // [ic_chunk][K][OB]
  const int tx = threadIdx.x % TU, ty = threadIdx.x / TU;
  const int u0 = blockIdx.x * TU;
  const int tiles_per_group = (p.out_g + OB - 1) / OB;
  const int grp = blockIdx.y / tiles_per_group;
  const int oc0 = (blockIdx.y % tiles_per_group) * OB;
  const int b = blockIdx.z;
  // first input sample any output of this tile can touch (floor division, may be negative)
  int num = u0 + p.peff - (p.K - 1) * p.dil;
  const int t_min = (num >= 0) ? num / p.stride : -((-num + p.stride - 1) / p.stride);
  const int i = u0 + tx;
  const int base = i + p.peff;  // >= 0
  float acc[OC_T];
#pragma unroll
  for (int c = 0; c < OC_T; ++c) acc[c] = 0.f;
  const float* inb = in + ((long long)b * p.Cin + (long long)grp * p.in_g) * p.Tin;
  const bool row_active = (oc0 + ty * OC_T) < p.out_g;

  for (int c0 = 0; c0 < p.in_g; c0 += p.ic_chunk) {
    const int nic = min(p.ic_chunk, p.in_g - c0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < nic * p.qlen; idx += 256) {
      int ic = idx / p.qlen, q = idx - ic * p.qlen;
      int t = t_min + q;
      ins[idx] = (t >= 0 && t < p.Tin) ? __ldg(inb + (long long)(c0 + ic) * p.Tin + t) : 0.f;
    }
    const int nko = p.K * p.out_g;  // contiguous run per input channel in global memory
    for (int idx = threadIdx.x; idx < nic * p.K * OB; idx += 256) {
      int ob = idx % OB, r = idx / OB;
      int k = r % p.K, ic = r / p.K;
      float v = 0.f;
      if (oc0 + ob < p.out_g)
        v = __ldg(w + (long long)(grp * p.in_g + c0 + ic) * nko + (long long)(oc0 + ob) * p.K + k);
      ws[idx] = v;
    }
    __syncthreads();
    if (row_active && i < p.Tout) {
      if (p.dil == 1) {
        const int phase = base % p.stride;
        const int q0 = base / p.stride - t_min;  // index of t for k = phase
        for (int ic = 0; ic < nic; ++ic) {
          const float* ir = ins + ic * p.qlen;
          const float* wr = ws + (size_t)ic * p.K * OB + ty * OC_T;
          int q = q0;
          for (int k = phase; k < p.K; k += p.stride, --q) {
            float xv = ir[q];  // zero outside [0,Tin)
#pragma unroll
            for (int c = 0; c < OC_T; ++c) acc[c] = fmaf(xv, wr[k * OB + c], acc[c]);
          }
        }
      } else {
        for (int ic = 0; ic < nic; ++ic) {
          const float* ir = ins + ic * p.qlen;
          const float* wr = ws + (size_t)ic * p.K * OB + ty * OC_T;
          for (int k = 0; k < p.K; ++k) {
            int r = base - k * p.dil;
            if (r < 0 || (r % p.stride) != 0) continue;
            float xv = ir[r / p.stride - t_min];
#pragma unroll
            for (int c = 0; c < OC_T; ++c) acc[c] = fmaf(xv, wr[k * OB + c], acc[c]);
          }
        }
      }
    }
  }
  if (!row_active || i >= p.Tout) return;
#pragma unroll
  for (int c = 0; c < OC_T; ++c) {
    int oc = oc0 + ty * OC_T + c;
    if (oc >= p.out_g) break;
    int ocg = grp * p.out_g + oc;
    float v = acc[c] + (bias ? __ldg(bias + ocg) : 0.f);
    out[((long long)b * p.Cout + ocg) * p.Tout + i] = v;
  }
}

template <int OC_T>
static int launch_tr_t(TrP p, const float* in, const float* w, const float* bias, float* out, cudaStream_t st) {
  constexpr int TU = 128, OB = 2 * OC_T;
  p.qlen = (TU - 1 + (p.K - 1) * p.dil) / p.stride + 3;
  size_t per_ic = ((size_t)p.qlen + (size_t)p.K * OB) * sizeof(float);
  int chunk = (int)((48 * 1024) / per_ic);
  if (chunk < 1) chunk = 1;
  if (chunk > p.in_g) chunk = p.in_g;
  p.ic_chunk = chunk;
  size_t smem = per_ic * chunk;
  TDVC_CHECK_ARG(smem <= 48 * 1024);
  dim3 grid(cdiv(p.Tout, TU), p.groups * cdiv(p.out_g, OB), p.B);
  TDVC_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535);
  tdvc::launch_k(conv_tr_k<OC_T>, grid, 256, smem, st, p, in, w, bias, out);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

static int launch_tr(TrP p, const float* in, const float* w, const float* bias, float* out, cudaStream_t st) {
  if (p.B == 0 || p.Tout <= 0) return TDVC_OK;
  g_flops[FLOP_FP32] += 2.0 * p.B * p.Tin * (double)p.Cout * p.in_g * p.K;
  int oc_t = p.out_g <= 2 ? 1 : (p.out_g <= 4 ? 2 : (p.out_g <= 8 ? 4 : 8));
  auto n_ctas = [&](int c) { return (long long)cdiv(p.Tout, 128) * p.groups * cdiv(p.out_g, 2 * c) * p.B; };
  while (oc_t > 1 && n_ctas(oc_t) < 2 * num_sms()) oc_t >>= 1;
  switch (oc_t) {
    case 1: return launch_tr_t<1>(p, in, w, bias, out, st);
    case 2: return launch_tr_t<2>(p, in, w, bias, out, st);
    case 4: return launch_tr_t<4>(p, in, w, bias, out, st);
    default: return launch_tr_t<8>(p, in, w, bias, out, st);
  }
}

// ---------------------------------------------------------------------------------------------
// reflect fold + input-activation mask:  dx[t] = mask(x[t]) * (stage[p+t] + mirrored halo terms)
__global__ void pad_act_bwd_k(const float* __restrict__ stage, const float* __restrict__ x, float* __restrict__ dx,
                              long long rows, int T, int p, int reflect, float slope) {
  pdl_prologue();
  long long n = rows * T;
  int Ts = T + 2 * p;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx / T;
    int t = (int)(idx - r * T);
    const float* s = stage + r * Ts;
    float v = s[p + t];
    if (reflect) {
      if (t >= 1 && t <= p) v += s[p - t];
      int m = 2 * (T - 1) - t;  // mirrored position in unpadded coordinates, lives in the right halo
      if (m >= T && m < T + p) v += s[p + m];
    }
    if (slope != 1.0f && x[idx] <= 0.f) v *= slope;
    dx[idx] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad:  out[ca, (cb,k)] = sum_{b,t} A[b, ca, t] * Bm[b, cb, t*stride + k*dil - pad]
struct WgP {
  int B, Ca, Ta, Cb, Tb, K, stride, pad, dil, groups;
  int ca_g, cb_g, pad_mode;
  float b_slope;
  int jtot, rows, span, nchunk, zsplit;
};

template <int RC>
__global__ void __launch_bounds__(256) conv_wgrad_k(WgP p, const float* __restrict__ A, const float* __restrict__ Bm,
                                                     float* __restrict__ out) {
  pdl_prologue();
  constexpr int TTW = 128, CAB = 16 * RC, JB = 64;
  extern __shared__ float sm[];
  float* a_s = sm;                  // [CAB][TTW]
  float* bs = sm + CAB * TTW;       // [rows][span]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int j0 = blockIdx.x * JB;
  const int tiles_per_group = (p.ca_g + CAB - 1) / CAB;
  const int grp = blockIdx.y / tiles_per_group;
  const int ca0 = (blockIdx.y % tiles_per_group) * CAB;
  const int cb_lo = j0 / p.K;
  int off[4];
  bool jok[4];
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    int j = j0 + tx + 16 * jj;
    jok[jj] = j < p.jtot;
    int cb = jok[jj] ? j / p.K : cb_lo;
    int k = jok[jj] ? j - cb * p.K : 0;
    off[jj] = (cb - cb_lo) * p.span + k * p.dil;
  }
  float acc[RC][4];
#pragma unroll
  for (int c = 0; c < RC; ++c)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) acc[c][jj] = 0.f;
  const int cb_hi = min((j0 + JB - 1) / p.K, p.cb_g - 1);
  const int nrows = cb_hi - cb_lo + 1;
  const int total = p.B * p.nchunk;
  for (int item = blockIdx.z; item < total; item += p.zsplit) {
    const int b = item / p.nchunk;
    const int tb = (item - b * p.nchunk) * TTW;
    const int nt = min(TTW, p.Ta - tb);
    __syncthreads();
    const float* Ab = A + ((long long)b * p.Ca + (long long)grp * p.ca_g + ca0) * p.Ta + tb;
    for (int idx = threadIdx.x; idx < CAB * TTW; idx += 256) {
      int ca = idx / TTW, t = idx - ca * TTW;
      a_s[idx] = (ca0 + ca < p.ca_g && t < nt) ? __ldg(Ab + (long long)ca * p.Ta + t) : 0.f;
    }
    const float* Bb = Bm + ((long long)b * p.Cb + (long long)grp * p.cb_g + cb_lo) * p.Tb;
    const int g0 = tb * p.stride - p.pad;
    for (int idx = threadIdx.x; idx < nrows * p.span; idx += 256) {
      int r = idx / p.span, i = idx - r * p.span;
      bs[idx] = fetch_padded(Bb + (long long)r * p.Tb, g0 + i, p.Tb, p.pad_mode, p.b_slope);
    }
    __syncthreads();
    const float* ar = a_s + ty * RC * TTW;
    for (int t = 0; t < nt; ++t) {
      float av[RC];
#pragma unroll
      for (int c = 0; c < RC; ++c) av[c] = ar[c * TTW + t];
      const int ts = t * p.stride;
      float b0 = bs[off[0] + ts], b1 = bs[off[1] + ts], b2 = bs[off[2] + ts], b3 = bs[off[3] + ts];
#pragma unroll
      for (int c = 0; c < RC; ++c) {
        acc[c][0] = fmaf(av[c], b0, acc[c][0]);
        acc[c][1] = fmaf(av[c], b1, acc[c][1]);
        acc[c][2] = fmaf(av[c], b2, acc[c][2]);
        acc[c][3] = fmaf(av[c], b3, acc[c][3]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < RC; ++c) {
    int ca = ca0 + ty * RC + c;
    if (ca >= p.ca_g) continue;
    float* orow = out + (long long)(grp * p.ca_g + ca) * p.jtot;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      if (jok[jj]) atomicAdd(orow + j0 + tx + 16 * jj, acc[c][jj]);
  }
}

template <int RC>
static int launch_wgrad_t(WgP p, const float* A, const float* Bm, float* out, cudaStream_t st) {
  constexpr int TTW = 128, CAB = 16 * RC, JB = 64;
  p.jtot = p.cb_g * p.K;
  p.span = (TTW - 1) * p.stride + (p.K - 1) * p.dil + 1;
  p.rows = min(p.cb_g, (JB + p.K - 2) / p.K + 1);
  p.nchunk = cdiv(p.Ta, TTW);
  size_t smem = ((size_t)CAB * TTW + (size_t)p.rows * p.span) * sizeof(float);
  TDVC_CHECK_ARG(smem <= 200 * 1024);
  // the attribute is per device: set on every launch that needs it (cheap) rather than once per process
  if (smem > 48 * 1024) {
    TDVC_CUDA(cudaFuncSetAttribute(conv_wgrad_k<RC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  int gx = cdiv(p.jtot, JB), gy = p.groups * cdiv(p.ca_g, CAB);
  long long total = (long long)p.B * p.nchunk;
  long long want = (4LL * num_sms() + (long long)gx * gy - 1) / ((long long)gx * gy);
  if (want < 1) want = 1;
  if (want > total) want = total;
  if (want > 65535) want = 65535;
  p.zsplit = (int)want;
  TDVC_CHECK_ARG(gy <= 65535);
  dim3 grid(gx, gy, p.zsplit);
  tdvc::launch_k(conv_wgrad_k<RC>, grid, 256, smem, st, p, A, Bm, out);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

static int launch_wgrad(WgP p, const float* A, const float* Bm, float* out, cudaStream_t st) {
  TDVC_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)p.Ca * p.cb_g * p.K, st));
  if (p.B == 0) return TDVC_OK;
  g_flops[FLOP_FP32] += 2.0 * p.B * p.Ta * (double)p.Ca * p.cb_g * p.K;
  if (p.ca_g <= 16) return launch_wgrad_t<1>(p, A, Bm, out, st);
  return launch_wgrad_t<4>(p, A, Bm, out, st);
}

// per-channel sum over (b, t): bias gradient
__global__ void channel_sum_k(const float* __restrict__ dy, float* __restrict__ db, int B, int C, int T, int zsplit) {
  pdl_prologue();
  __shared__ float sm[33];
  const int c = blockIdx.x;
  float s = 0.f;
  for (int b = blockIdx.y; b < B; b += zsplit) {
    const float* r = dy + ((long long)b * C + c) * T;
    for (int t = threadIdx.x; t < T; t += blockDim.x) s += r[t];
  }
  s = block_sum(s, sm);
  if (threadIdx.x == 0) atomicAdd(db + c, s);
}

// Weight (and bias) gradient of a 1-input-channel stride-1 conv -- the waveform stems: discriminator.0.0 (1 -> 16, k15,
// reflect), encoder.0 and excite_downsample.4 (1 -> 16 / 8, k7, reflect).  dw[co][k] = sum_{b,t} dy[b,co,t] * xpad[b, t + k*dil - pad].
// A CTA takes one (batch, 1024-step chunk): the padded input chunk and 256-step tiles of dy are staged in shared memory,
// thread p owns the pair (co, k) = (p / K, p % K) -- the 15 taps of a channel read consecutive shared-memory words, the dy
// value is a broadcast -- and adds its partial sum with one atomic.  The general kernel above spends 160 us on the
// discriminator stem at B = 32 (a few CTAs walk the whole tensor); this one reads dy once at full rate.
// 256-step chunks: a CTA's work is a chain of dependent global-load rounds (tile, then each dy sub-tile); ncu showed the kernel
// waiting on that chain (long-scoreboard stalls 13-15 per issue, 288 CTAs of 5 rounds each: 35-50 us) -- short chains in many
// CTAs overlap instead
constexpr int STEM_TT = 256, STEM_SUB = 256;

__global__ void __launch_bounds__(256) stem_wgrad_k(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                                                     float* __restrict__ db, int Cout, int T, int K, int dil, int pad, int pad_mode,
                                                     float in_slope) {
  pdl_prologue();
  extern __shared__ float sm[];
  const int span = STEM_TT + (K - 1) * dil;
  float* xs = sm;                              // [span]
  float* dys = sm + ((span + 3) & ~3);         // [Cout][STEM_SUB + 1]
  const int b = blockIdx.y, t0 = blockIdx.x * STEM_TT;
  const float* xrow = x + (long long)b * T;
  for (int base = threadIdx.x; base < span; base += 256 * 4) {      // 4 loads in flight per thread
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = base + 256 * u < span ? fetch_padded(xrow, t0 + base + 256 * u - pad, T, pad_mode, in_slope) : 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) if (base + 256 * u < span) xs[base + 256 * u] = v[u];
  }
  const int npairs = Cout * K;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};         // pairs p, p + 256, ... (Cout * K <= 1024)
  float bsum = 0.f;                            // threads < Cout: bias gradient of channel threadIdx.x
  const int tend = min(STEM_TT, T - t0);
  for (int s0 = 0; s0 < tend; s0 += STEM_SUB) {
    const int ns = min(STEM_SUB, tend - s0);
    __syncthreads();
    for (int base = threadIdx.x; base < Cout * STEM_SUB; base += 256 * 8) {      // 8 loads in flight per thread
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + 256 * u;
        const int co = i / STEM_SUB, t = i - co * STEM_SUB;
        v[u] = (i < Cout * STEM_SUB && t < ns) ? __ldg(dy + ((long long)b * Cout + co) * T + t0 + s0 + t) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + 256 * u;
        const int co = i / STEM_SUB, t = i - co * STEM_SUB;
        if (i < Cout * STEM_SUB) dys[co * (STEM_SUB + 1) + t] = v[u];
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int pidx = threadIdx.x + 256 * q;
      if (pidx < npairs) {
        const int co = pidx / K, k = pidx - co * K;
        const float* dr = dys + co * (STEM_SUB + 1);
        const float* xr = xs + s0 + k * dil;
        float a = acc[q];
        for (int t = 0; t < ns; ++t) a = fmaf(dr[t], xr[t], a);
        acc[q] = a;
      }
    }
    if (db && threadIdx.x < Cout) {
      const float* dr = dys + threadIdx.x * (STEM_SUB + 1);
      for (int t = 0; t < ns; ++t) bsum += dr[t];
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int pidx = threadIdx.x + 256 * q;
    if (pidx < npairs) atomicAdd(dw + pidx, acc[q]);
  }
  if (db && threadIdx.x < Cout) atomicAdd(db + threadIdx.x, bsum);
}


// Weight (and bias) gradient of the narrow convs -- the excitation pyramid (8 -> 8 channels: k = 2r stride r, k5, 1x1) and
// the 1-channel output heads (C -> 1, k7): Cout <= 16 and Cout * Cin * K <= 2048.  The general kernel gives such a conv a
// few CTAs that each walk the whole tensor (59 us per call on average in the step, 15 calls); here a CTA takes one (batch,
// TT-step chunk), stages the padded input span of all Cin channels and 128-step tiles of dy in shared memory, thread p owns
// the triples (co, ci, k) = p, p + 256, ... and adds its partial sums with one atomic each.
constexpr int NARROW_SUB = 128, NARROW_ACC = 8;

__global__ void __launch_bounds__(256) narrow_wgrad_k(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                                                       float* __restrict__ db, int Cin, int Cout, int Tin, int Tout, int K, int stride,
                                                       int dil, int pad, int pad_mode, float in_slope, int TT) {
  pdl_prologue();
  extern __shared__ float sm[];
  const int span = (TT - 1) * stride + (K - 1) * dil + 1;
  const int spanp = span | 1;                  // odd pitch: channels start in different banks
  float* xs = sm;                              // [Cin][spanp]
  float* dys = sm + ((Cin * spanp + 3) & ~3);  // [Cout][NARROW_SUB + 1]
  const int b = blockIdx.y, t0 = blockIdx.x * TT;
  const float* xb = x + (long long)b * Cin * Tin;
  const int g0 = t0 * stride - pad;
  // 8 independent loads in flight per thread: with a handful of warps per SM a load-then-store loop is latency bound
  for (int base = threadIdx.x; base < Cin * span; base += 256 * 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + 256 * u;
      const int ci = i / span, j = i - ci * span;
      v[u] = i < Cin * span ? fetch_padded(xb + (long long)ci * Tin, g0 + j, Tin, pad_mode, in_slope) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + 256 * u;
      const int ci = i / span, j = i - ci * span;
      if (i < Cin * span) xs[ci * spanp + j] = v[u];
    }
  }
  const int npairs = Cout * Cin * K, cik = Cin * K;
  float acc[NARROW_ACC];
  int xoff[NARROW_ACC], doff[NARROW_ACC];
#pragma unroll
  for (int q = 0; q < NARROW_ACC; ++q) {
    acc[q] = 0.f;
    const int pidx = min(threadIdx.x + 256 * q, npairs - 1);
    const int co = pidx / cik, rem = pidx - co * cik, ci = rem / K, k = rem - ci * K;
    xoff[q] = ci * spanp + k * dil;
    doff[q] = co * (NARROW_SUB + 1);
  }
  float bsum = 0.f;
  const int tend = min(TT, Tout - t0);
  for (int s0 = 0; s0 < tend; s0 += NARROW_SUB) {
    const int ns = min(NARROW_SUB, tend - s0);
    __syncthreads();
    for (int base = threadIdx.x; base < Cout * NARROW_SUB; base += 256 * 4) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + 256 * u;
        const int co = i / NARROW_SUB, t = i - co * NARROW_SUB;
        v[u] = (i < Cout * NARROW_SUB && t < ns) ? __ldg(dy + ((long long)b * Cout + co) * Tout + t0 + s0 + t) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + 256 * u;
        const int co = i / NARROW_SUB, t = i - co * NARROW_SUB;
        if (i < Cout * NARROW_SUB) dys[co * (NARROW_SUB + 1) + t] = v[u];
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NARROW_ACC; ++q) {
      if (threadIdx.x + 256 * q < npairs) {
        const float* dr = dys + doff[q];
        const float* xr = xs + xoff[q] + s0 * stride;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;      // four chains: the sum is not bound by the FMA latency
        int t = 0;
        for (; t + 4 <= ns; t += 4) {
          a0 = fmaf(dr[t], xr[t * stride], a0);
          a1 = fmaf(dr[t + 1], xr[(t + 1) * stride], a1);
          a2 = fmaf(dr[t + 2], xr[(t + 2) * stride], a2);
          a3 = fmaf(dr[t + 3], xr[(t + 3) * stride], a3);
        }
        for (; t < ns; ++t) a0 = fmaf(dr[t], xr[t * stride], a0);
        acc[q] += (a0 + a1) + (a2 + a3);
      }
    }
    if (db && threadIdx.x < Cout) {
      const float* dr = dys + threadIdx.x * (NARROW_SUB + 1);
      for (int t = 0; t < ns; ++t) bsum += dr[t];
    }
  }
#pragma unroll
  for (int q = 0; q < NARROW_ACC; ++q) {
    const int pidx = threadIdx.x + 256 * q;
    if (pidx < npairs) atomicAdd(dw + pidx, acc[q]);
  }
  if (db && threadIdx.x < Cout) atomicAdd(db + threadIdx.x, bsum);
}

// chunk length of narrow_wgrad_k (256 or 128 time steps: short load chains in many CTAs, see STEM_TT); 0 = no chunk fits
static int narrow_wgrad_tt(int Cin, int Cout, int K, int stride, int dil, size_t* smem) {
  // first choice: tiles of <= 40 KB (several CTAs per SM hide the staging latency); else whatever fits 160 KB
  for (size_t budget : {(size_t)40 * 1024, (size_t)160 * 1024})
    for (int TT = 256; TT >= 128; TT >>= 1) {
      const int span = (TT - 1) * stride + (K - 1) * dil + 1;
      const size_t need = ((size_t)((Cin * (span | 1) + 3) & ~3) + (size_t)Cout * (NARROW_SUB + 1)) * sizeof(float);
      if (need <= budget) { *smem = need; return TT; }
    }
  return 0;
}

static int launch_channel_sum(const float* dy, float* db, int B, int C, int T, cudaStream_t st) {
  TDVC_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * C, st));
  if (B == 0) return TDVC_OK;
  int zs = min(B, max(1, (2 * num_sms()) / C));
  tdvc::launch_k(channel_sum_k, dim3(C, zs), 256, 0, st, dy, db, B, C, T, zs);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

}  // namespace tdvc

using namespace tdvc;

extern "C" int tdvc_conv1d_fwd(const tdvc_conv_geom* g, const float* x, const float* w, const float* bias,
                               const float* residual, float* y, void* stream) {
  int rc = check_geom(g, false);
  if (rc) return rc;
  TDVC_CHECK_ARG(x && w && y);
  return launch_fwd(fwd_params(g), x, w, bias, residual, y, (cudaStream_t)stream);
}

static bool needs_stage(const tdvc_conv_geom* g) {
  return (g->pad_mode == TDVC_PAD_REFLECT && g->pad > 0) || g->in_slope != 1.0f;
}

extern "C" int64_t tdvc_conv1d_bwd_data_ws(const tdvc_conv_geom* g) {
  if (!g || !needs_stage(g)) return 0;
  int ph = (g->pad_mode == TDVC_PAD_REFLECT) ? g->pad : 0;
  return (int64_t)g->B * g->Cin * (g->Tin + 2 * ph);
}

extern "C" int tdvc_conv1d_bwd_data(const tdvc_conv_geom* g, const float* dy, const float* w, const float* x,
                                    float* dx, float* ws, void* stream) {
  int rc = check_geom(g, false);
  if (rc) return rc;
  TDVC_CHECK_ARG(dy && w && dx);
  cudaStream_t st = (cudaStream_t)stream;
  const bool stage = needs_stage(g);
  const int ph = (g->pad_mode == TDVC_PAD_REFLECT) ? g->pad : 0;  // halo kept in the staging buffer
  if (stage) TDVC_CHECK_ARG(ws != nullptr);
  if (g->in_slope != 1.0f) TDVC_CHECK_ARG(x != nullptr);
  float* target = stage ? ws : dx;
  const int Lout = g->Tin + 2 * ph;
  const int peff = g->pad - ph;
  if (g->stride == 1 && g->groups == 1) {
    // dgrad as a forward conv of dy with channel-swapped, tap-flipped weights
    FwdP p{};
    p.B = g->B; p.Cin = g->Cout; p.Tin = g->Tout; p.Cout = g->Cin; p.Tout = Lout; p.K = g->K;
    p.stride = 1; p.dil = g->dilation; p.groups = 1; p.pad = (g->K - 1) * g->dilation - peff;
    p.pad_mode = TDVC_PAD_ZEROS; p.out_act = TDVC_ACT_NONE; p.in_slope = 1.0f; p.out_slope = 1.0f;
    p.cin_g = p.Cin; p.cout_g = p.Cout;
    p.w_sco = g->K; p.w_sci = (long long)g->Cin * g->K; p.w_sk = 1; p.w_flip = 1;
    TDVC_CHECK_ARG(p.pad >= 0);
    rc = launch_fwd(p, dy, w, nullptr, nullptr, target, st);
  } else {
    TrP p{};
    p.B = g->B; p.Cin = g->Cout; p.Tin = g->Tout; p.Cout = g->Cin; p.Tout = Lout; p.K = g->K;
    p.stride = g->stride; p.peff = peff; p.dil = g->dilation; p.groups = g->groups;
    p.in_g = g->Cout / g->groups; p.out_g = g->Cin / g->groups;
    rc = launch_tr(p, dy, w, nullptr, target, st);
  }
  if (rc) return rc;
  if (stage) {
    long long rows = (long long)g->B * g->Cin;
    long long n = rows * g->Tin;
    int blocks = (int)std::min<long long>((n + 255) / 256, 8LL * num_sms());
    if (blocks > 0) {
      tdvc::launch_k(pad_act_bwd_k, blocks, 256, 0, st, ws, x, dx, rows, g->Tin, ph, g->pad_mode == TDVC_PAD_REFLECT, g->in_slope);
      TDVC_LAUNCH_CHECK();
    }
  }
  return TDVC_OK;
}

extern "C" int tdvc_bias_grad(const float* dy, float* dbias, int B, int C, int T, void* stream) {
  TDVC_CHECK_ARG(dy && dbias && B >= 0 && C > 0 && T > 0);
  return launch_channel_sum(dy, dbias, B, C, T, (cudaStream_t)stream);
}

extern "C" int tdvc_pad_act_bwd(const float* stage, const float* x, float* dx, int64_t rows, int T, int halo, int reflect,
                                float in_slope, void* stream) {
  TDVC_CHECK_ARG(stage && dx && rows >= 0 && T > 0 && halo >= 0);
  TDVC_CHECK_ARG(in_slope == 1.0f || x != nullptr);
  long long n = (long long)rows * T;
  if (n == 0) return TDVC_OK;
  int blocks = (int)std::min<long long>((n + 255) / 256, 8LL * num_sms());
  tdvc::launch_k(pad_act_bwd_k, blocks, 256, 0, (cudaStream_t)stream, stage, x, dx, rows, T, halo, reflect, in_slope);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_conv1d_bwd_weight(const tdvc_conv_geom* g, const float* dy, const float* x, float* dw,
                                      float* dbias, void* stream) {
  int rc = check_geom(g, false);
  if (rc) return rc;
  TDVC_CHECK_ARG(dy && x && dw);
  cudaStream_t st = (cudaStream_t)stream;
  if (g->Cin == 1 && g->groups == 1 && g->stride == 1 && g->Cout * g->K <= 1024 && g->Cout <= 256 && g->Tout == g->Tin) {
    // the waveform stems
    TDVC_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)g->Cout * g->K, st));
    if (dbias) TDVC_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)g->Cout, st));
    if (g->B == 0) return TDVC_OK;
    const int span = STEM_TT + (g->K - 1) * g->dilation;
    const size_t smem = (size_t)(((span + 3) & ~3) + g->Cout * (STEM_SUB + 1)) * sizeof(float);
    if (smem <= 200 * 1024) {
      if (smem > 48 * 1024) TDVC_CUDA(cudaFuncSetAttribute(stem_wgrad_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      tdvc::launch_k(stem_wgrad_k, dim3(cdiv(g->Tout, STEM_TT), g->B), 256, smem, st, dy, x, dw, dbias, g->Cout, g->Tin, g->K,
                     g->dilation, g->pad, g->pad_mode, g->in_slope);
      TDVC_LAUNCH_CHECK();
      g_flops[FLOP_FP32] += 2.0 * g->B * g->Tout * (double)g->Cout * g->K;
      return TDVC_OK;
    }
  }
  if (g->groups == 1 && g->Cin > 1 && g->Cout <= 16 && (long long)g->Cout * g->Cin * g->K <= 256 * NARROW_ACC && g->B > 0) {
    // the excitation pyramid and the 1-channel heads
    size_t smem = 0;
    const int TT = narrow_wgrad_tt(g->Cin, g->Cout, g->K, g->stride, g->dilation, &smem);
    if (TT > 0) {
      TDVC_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)g->Cout * g->Cin * g->K, st));
      if (dbias) TDVC_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)g->Cout, st));
      if (smem > 48 * 1024) TDVC_CUDA(cudaFuncSetAttribute(narrow_wgrad_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      tdvc::launch_k(narrow_wgrad_k, dim3(cdiv(g->Tout, TT), g->B), 256, smem, st, dy, x, dw, dbias, g->Cin, g->Cout, g->Tin,
                     g->Tout, g->K, g->stride, g->dilation, g->pad, g->pad_mode, g->in_slope, TT);
      TDVC_LAUNCH_CHECK();
      g_flops[FLOP_FP32] += 2.0 * g->B * g->Tout * (double)g->Cout * g->Cin * g->K;
      return TDVC_OK;
    }
  }
  WgP p{};
  p.B = g->B; p.Ca = g->Cout; p.Ta = g->Tout; p.Cb = g->Cin; p.Tb = g->Tin; p.K = g->K;
  p.stride = g->stride; p.pad = g->pad; p.dil = g->dilation; p.groups = g->groups;
  p.ca_g = g->Cout / g->groups; p.cb_g = g->Cin / g->groups; p.pad_mode = g->pad_mode; p.b_slope = g->in_slope;
  rc = launch_wgrad(p, dy, x, dw, st);
  if (rc) return rc;
  if (dbias) return launch_channel_sum(dy, dbias, g->B, g->Cout, g->Tout, st);
  return TDVC_OK;
}

extern "C" int tdvc_conv_transpose1d_fwd(const tdvc_conv_geom* g, const float* x, const float* w, const float* bias,
                                         float* y, void* stream) {
  int rc = check_geom(g, true);
  if (rc) return rc;
  TDVC_CHECK_ARG(x && w && y);
  TrP p{};
  p.B = g->B; p.Cin = g->Cin; p.Tin = g->Tin; p.Cout = g->Cout; p.Tout = g->Tout; p.K = g->K;
  p.stride = g->stride; p.peff = g->pad; p.dil = 1; p.groups = 1; p.in_g = g->Cin; p.out_g = g->Cout;
  return launch_tr(p, x, w, bias, y, (cudaStream_t)stream);
}

extern "C" int tdvc_conv_transpose1d_bwd_data(const tdvc_conv_geom* g, const float* dy, const float* w, float* dx,
                                              void* stream) {
  int rc = check_geom(g, true);
  if (rc) return rc;
  TDVC_CHECK_ARG(dy && w && dx);
  // dx[b,ci,t] = sum_{co,k} dy[b,co,t*s + k - p] * w[ci,co,k]: a strided forward conv of dy whose
  // "output channels" are ci -- w[Cin,Cout,K] already has that layout.
  FwdP p{};
  p.B = g->B; p.Cin = g->Cout; p.Tin = g->Tout; p.Cout = g->Cin; p.Tout = g->Tin; p.K = g->K;
  p.stride = g->stride; p.pad = g->pad; p.dil = 1; p.groups = 1;
  p.pad_mode = TDVC_PAD_ZEROS; p.out_act = TDVC_ACT_NONE; p.in_slope = 1.0f; p.out_slope = 1.0f;
  p.cin_g = p.Cin; p.cout_g = p.Cout;
  p.w_sco = (long long)g->Cout * g->K; p.w_sci = g->K; p.w_sk = 1; p.w_flip = 0;
  return launch_fwd(p, dy, w, nullptr, nullptr, dx, (cudaStream_t)stream);
}

extern "C" int tdvc_conv_transpose1d_bwd_weight(const tdvc_conv_geom* g, const float* dy, const float* x, float* dw,
                                                float* dbias, void* stream) {
  int rc = check_geom(g, true);
  if (rc) return rc;
  TDVC_CHECK_ARG(dy && x && dw);
  cudaStream_t st = (cudaStream_t)stream;
  // dw[ci,co,k] = sum_{b,t} x[b,ci,t] * dy[b,co,t*s + k - p]
  WgP p{};
  p.B = g->B; p.Ca = g->Cin; p.Ta = g->Tin; p.Cb = g->Cout; p.Tb = g->Tout; p.K = g->K;
  p.stride = g->stride; p.pad = g->pad; p.dil = 1; p.groups = 1;
  p.ca_g = g->Cin; p.cb_g = g->Cout; p.pad_mode = TDVC_PAD_ZEROS; p.b_slope = 1.0f;
  rc = launch_wgrad(p, x, dy, dw, st);
  if (rc) return rc;
  if (dbias) return launch_channel_sum(dy, dbias, g->B, g->Cout, g->Tout, st);
  return TDVC_OK;
}

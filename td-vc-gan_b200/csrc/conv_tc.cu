// bf16 tensor-core path for the dense convolutions (FiLM cond_var.0 / cond_var.2, conv.1, posconv, the
// discriminator's 1024x1024 k5 layer, the frame-view strided / transposed convs, and all their data and weight
// gradients): implicit GEMM on tcgen05.mma with the accumulator in TMEM and both operands brought in by TMA.
//
// Time-as-M kernels (conv_tc_fwd_k: one tile per CTA; conv_tc_ws_k: weight-stationary, persistent):
//   D[t, co] = sum_{tap} sum_{ci} A_tap[t, ci] * W_tap[co, ci]
//     A_tap = rows (t0 + tap*dilation + t_off ...) of the channels-last bf16 activation copy xp[B, Tp, Cp]
//             -> 3-D TMA boxes {64 ch, 128 t (+ halo), 1 b} per 64-channel chunk, K-major, SWIZZLE_128B;
//                the conv's zero padding is TMA out-of-bounds fill, reflect padding is materialised by
//                the pack kernel in the halo rows.
//     W_tap = wp[tap, co, ci] bf16 -> box {64 ch, BN co, 1 tap}, K-major, SWIZZLE_128B.
//   M = 128 time steps (TMEM lanes), N = BN <= 256 output channels (TMEM columns), K = 16 per MMA.
// Weights-as-M kernel (conv_tc_wt_k): the same sum with the roles swapped, M = 128 stacked output channels,
//   N = 256 time steps -- see the comment above it.
// Weight gradient (conv_tc_wgrad_k): M = ci, N = co, K = time, both operands MN-major views of the same packed tensors.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one lane chosen by elect.sync),
// the remaining warps (16 in the forward kernels, 4 in wgrad) = epilogue: tcgen05.ld -> bias / FiLM / residual /
// activation / LeakyReLU mask -> NCW fp32 stores (a TMEM lane is a time step: for a fixed channel a warp writes 32
// consecutive floats) or the next conv's packed bf16 operand.
#include "tc_common.cuh"

namespace tdvc {

struct TcP {
  int B, Tout, Cout, K, dil, t_off;          // Cout = valid output channels per group
  int nchunk, last_nk16, BN, stages, tmem_cols;
  // weight-stationary kernel, narrow tiles: the 16 epilogue warps form 4 / epi_spt teams of epi_spt column sets; team k takes
  // the CTA's tiles k, k + teams, ... and nacc = 2 * teams accumulators are in flight (0 = one team of 4 sets, 2 accumulators)
  int epi_spt, nacc;
  // conv_tc_fwdh_k: the activation tile is loaded ONCE per 64-channel chunk with its (K-1)*dil halo rows (a_rows rows, two
  // stages) and the weight tiles stream through their own ring of b_stages
  int a_rows, b_stages;
  int out_act;
  float out_slope;
  const float* bias;
  const float* gb;
  const float* res;
  float* y;
  int debug;   // TDVC_TC_DEBUG (development only): 1 = no epilogue stores, 2 = no main loop, 4 = no tmem loads
  // grouped / packed extensions (tdvc_conv1d_tc_fwd_ex)
  int tiles_per_group, coutp_g, a_ch_off, a_ch_stride, bias_stride;
  __nv_bfloat16* yp;                          // OUT == 1: packed bf16 channels-last output [B, tp_out, cp_out]
  int tp_out, cp_out, out_halo, out_ch_off, out_ch_stride;
  const __nv_bfloat16* maskp;                 // MASK == 1: packed activated tensor whose sign gates the result
  int tm, cm, mask_halo, mask_ch_off, mask_ch_stride;
  float mask_slope;
  long long y_grp_stride, y_b_stride;         // fp32 output: g*y_grp_stride + b*y_b_stride + n*Tout + t
  // chain epilogues (EPI >= 3): the MRF stage kept in bf16 channels-last between its convolutions
  int kg[4];                                  // per-group kernel size, centred inside the K taps of the weight tensor
  long long gb_grp_stride, res_grp_stride;    // FiLM gamma|beta and residual tensors: element offset of group g
  __nv_bfloat16* yp2;                         // EPI 3: second packed output (the un-modulated conv result), geometry of yp
  const __nv_bfloat16* auxp;                  // EPI 5: packed conv result saved by EPI 3, geometry of maskp
  __nv_bfloat16* dgbp;                        // EPI 5: packed dL/d(gamma|beta) rows [B][Tout][dgb_cp]
  int dgb_cp, dgb_ch_off, dgb_ch_stride;
  float* halo_buf;                            // EPI 6: contributions that fall on reflect-halo rows [grp][B][Cout][2*halo]
  int halo, t_valid;
  float pk_slope;                             // EPI 4: LeakyReLU slope of the packed copy
  int unframe_s, unframe_pad, unframe_T, unframe_C;   // fp32 output through the inverse frame view (see tdvc_tc_conv)
  int flat_tp, flat_halo, flat_T;                     // batch-flattened short sequences (see tdvc_tc_conv)
  // weight-stationary kernel, groups with different tap counts: CTAs are dealt out in proportion to the taps (the k = 11
  // branch gets 11/21 of them instead of a third); group g owns CTAs [grp_cta0[g], grp_cta0[g+1]) of a 1-D grid
  int balanced;
  int grp_cta0[5];
};

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;
constexpr int TC_THREADS = 192;       // wgrad kernel: TMA warp, MMA warp, 4 epilogue warps
constexpr int TC_EPI_WARPS = 16;      // forward kernels: epilogue warps (a multiple of 4: TC_EPI_WARPS / 4 per TMEM lane quadrant)
constexpr int TC_FWD_THREADS = 64 + 32 * TC_EPI_WARPS;   // + TMA warp + MMA warp

__device__ __forceinline__ void store16_bf16(__nv_bfloat16* dst, const float* v) {
  uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&h2);
  }
  reinterpret_cast<uint4*>(dst)[0] = make_uint4(w[0], w[1], w[2], w[3]);
  reinterpret_cast<uint4*>(dst)[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
// 16 packed bf16 (32 bytes) of one row: loaded as two 16-byte words, decoded later (the load can be issued long before its use)
__device__ __forceinline__ void ld32B_bf16(const __nv_bfloat16* src, uint4& a, uint4& c) {
  a = __ldg(reinterpret_cast<const uint4*>(src));
  c = __ldg(reinterpret_cast<const uint4*>(src) + 1);
}
__device__ __forceinline__ void unpack16_bf16(const uint4& a, const uint4& c, float* v) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[2 * j] = __uint_as_float(w[j] << 16);
    v[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
  }
}
// v[j] *= slope where the packed activated value is <= 0 (bf16 sign / zero test on the raw bits)
__device__ __forceinline__ void mask16_words(const uint4& m0, const uint4& m1, float* v, float slope) {
  const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const uint32_t h = (j & 1) ? (mw[j >> 1] >> 16) : (mw[j >> 1] & 0xFFFFu);
    if ((h & 0x8000u) || (h & 0x7FFFu) == 0) v[j] *= slope;
  }
}

__device__ __forceinline__ int epi_spt_of(const TcP& p) { return p.epi_spt > 0 ? p.epi_spt : TC_EPI_WARPS / 4; }

// Operands-first epilogues.  Everything an epilogue reads from global memory besides the accumulator (LeakyReLU masks, FiLM
// gamma | beta, residuals, the saved conv result) has an address that is known before the tile's MMAs retire, so those loads
// are issued BEFORE the wait on the accumulator barrier and their memory round trip overlaps the MMAs instead of following
// them: a narrow tile's epilogue was a chain of dependent round trips (TMEM load -> operand loads -> store, 2.8 us per tile
// under ncu) and the cond_var.2 data gradient ran at 3.2 TB/s on algorithmic traffic (profiles/r2/final_ncu_conv_metrics.txt).
// Same arithmetic in the same order: results are bit-identical.  TDVC_EPI_HOIST=0 compiles the loads back behind the wait.
#ifndef TDVC_EPI_HOIST
#define TDVC_EPI_HOIST 1
#endif
// plain mask path: column chunks per thread whose mask words are loaded ahead (packed output: 3, covers N <= 192 -- the
// 136-channel cond_var.2 data gradient; fp32 output: 2, which keeps conv_tc_fwdh_k<0,0,0,1> inside the 56 registers ptxas
// gives it for two resident CTAs)
template <int OUT> struct EpiPf { static constexpr int n = OUT == 1 ? 3 : 2; };

struct ChainOps {
  uint4 m0, m1;        // modes 5, 6: packed activated tensor whose sign gates the result
  uint4 a0, a1;        // mode 5: the packed un-modulated conv result h0
  float g[16];         // mode 3 / 5: gamma;  mode 4 / 6: residual
  float bt[16];        // mode 3: beta
};

template <int MODE>
__device__ __forceinline__ void tc_chain_load(const TcP& p, int b, int t, int grp, int ch, ChainOps& o) {
  if (MODE == 3) {
    if (p.gb) {
      const long long ct = p.Tout;
      const float* gp = p.gb + (long long)grp * p.gb_grp_stride + ((long long)b * 2 * p.Cout + ch) * ct + t;
      const long long beta_off = (long long)p.Cout * ct;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        o.g[j] = __ldg(gp + j * ct);
        o.bt[j] = __ldg(gp + beta_off + j * ct);
      }
    }
  } else if (MODE == 4) {
    const long long ct = p.Tout;
    const float* rp = p.res + (long long)grp * p.res_grp_stride + ((long long)b * p.Cout + ch) * ct + t;
#pragma unroll
    for (int j = 0; j < 16; ++j) o.g[j] = __ldg(rp + j * ct);
  } else if (MODE == 5) {
    const long long mrow = ((long long)b * p.tm + t + p.mask_halo) * p.cm + p.mask_ch_off + grp * p.mask_ch_stride + ch;
    ld32B_bf16(p.maskp + mrow, o.m0, o.m1);
    if (p.gb) {
      ld32B_bf16(p.auxp + mrow, o.a0, o.a1);
      const long long ct = p.Tout;
      const float* gp = p.gb + (long long)grp * p.gb_grp_stride + ((long long)b * 2 * p.Cout + ch) * ct + t;
#pragma unroll
      for (int j = 0; j < 16; ++j) o.g[j] = __ldg(gp + j * ct);
    }
  } else if (MODE == 6) {
    // rows are positions of the PADDED input (t_valid + 2 * halo of them); the packed activated input carries the same
    // halo, so row t of it gates row t here -- for a halo row that is the sign of the sample it mirrors, which is the
    // derivative the folded contribution needs
    const long long mrow = ((long long)b * p.tm + t + p.mask_halo) * p.cm + p.mask_ch_off + grp * p.mask_ch_stride + ch;
    ld32B_bf16(p.maskp + mrow, o.m0, o.m1);
    const int tt = t - p.halo;
    if (p.res && tt >= 0 && tt < p.t_valid) {
      const long long ct = p.t_valid;
      const float* rp = p.res + (long long)grp * p.res_grp_stride + ((long long)b * p.Cout + ch) * ct + tt;
#pragma unroll
      for (int j = 0; j < 16; ++j) o.g[j] = __ldg(rp + j * ct);
    }
  }
}

// Epilogues of the bf16-resident MRF stage (model/generator.py:69-111,175-194).  Every group is one kernel-size branch; all
// tensors between the convolutions are bf16 channels-last with the branches side by side in the channel dimension, the
// residual stream stays fp32 NCW [branch][B][C][T].  Channel counts are multiples of 16 (a thread's 16 columns are real).
//   MODE 3  conv.1      h0 = acc + bias -> yp2 (kept for the backward);  a1 = leaky_relu(h0 * (1 + gamma) + beta) -> yp
//   MODE 4  posconv.1   x' = acc + bias + x -> y fp32;  leaky_relu(x') -> yp with the reflect halo rows of the next conv.1
//   MODE 5  posconv^T   d = acc * lrelu'(a1);  dgamma = d * h0, dbeta = d -> dgbp;  dh0 = d * (1 + gamma) -> yp
//   MODE 6  conv.1^T    over the padded rows: d = acc * lrelu'(x);  interior rows: dx = d + dx' -> y fp32 and yp;
//                       halo rows -> halo_buf (chain_fold_k adds them onto the rows they mirror)
// wait_acc() blocks until the accumulator is complete; the first chunk's operands are in flight by then.
template <int MODE, typename Wait>
__device__ __forceinline__ void tc_epilogue_chain(const TcP& p, const float* bias_s, uint32_t acc, int b, int t0, int grp,
                                                  int n0, int q, int half, int lane, Wait wait_acc) {
  const int t = t0 + q * 32 + lane;
  const bool t_ok = t < p.Tout;
  const int nvalid = min(p.BN, p.Cout - n0);
  const int step = 16 * epi_spt_of(p), first = (half % epi_spt_of(p)) * 16;
  ChainOps o;
  if (TDVC_EPI_HOIST && t_ok && first < nvalid) tc_chain_load<MODE>(p, b, t, grp, n0 + first, o);
  wait_acc();
  for (int c0 = first; c0 < nvalid; c0 += step) {
    float v[16];
    tmem_ld16(acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
    if (!t_ok) continue;
    {
      const float4* b4 = reinterpret_cast<const float4*>(bias_s + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 bb = b4[j];
        v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
      }
    }
    const int ch = n0 + c0;                              // first of this thread's 16 channels inside the group
    if (!TDVC_EPI_HOIST || c0 != first) tc_chain_load<MODE>(p, b, t, grp, ch, o);
    if (MODE == 3) {
      const long long prow = ((long long)b * p.tp_out + t + p.out_halo) * p.cp_out + p.out_ch_off + grp * p.out_ch_stride + ch;
      if (p.yp2) store16_bf16(p.yp2 + prow, v);
      if (p.gb) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], 1.f + o.g[j], o.bt[j]);
      }
      const float sl = p.out_slope;
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], v[j] * sl);
      store16_bf16(p.yp + prow, v);
    } else if (MODE == 4) {
      const long long ct = p.Tout;
      float* yo = p.y + (long long)grp * p.y_grp_stride + (long long)b * p.y_b_stride + (long long)ch * ct + t;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        v[j] += o.g[j];
        yo[j * ct] = v[j];
      }
      if (p.yp) {
        const float sl = p.pk_slope;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], v[j] * sl);
        __nv_bfloat16* col = p.yp + (long long)b * p.tp_out * p.cp_out + p.out_ch_off + grp * p.out_ch_stride + ch;
        const int H = p.out_halo;
        store16_bf16(col + (long long)(t + H) * p.cp_out, v);
        // reflect halo of the consumer: padded row H - i (i = 1..H) mirrors x[i], padded row H + T - 1 + i mirrors x[T - 1 - i]
        if (t >= 1 && t <= H) store16_bf16(col + (long long)(H - t) * p.cp_out, v);
        if (t <= p.Tout - 2 && t >= p.Tout - 1 - H) store16_bf16(col + (long long)(H + 2 * (p.Tout - 1) - t) * p.cp_out, v);
      }
    } else if (MODE == 5) {
      mask16_words(o.m0, o.m1, v, p.mask_slope);
      if (p.gb) {
        float h0[16];
        unpack16_bf16(o.a0, o.a1, h0);
        float dg[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) dg[j] = v[j] * h0[j];
        __nv_bfloat16* dgp = p.dgbp + ((long long)b * p.Tout + t) * p.dgb_cp + p.dgb_ch_off + grp * p.dgb_ch_stride + ch;
        store16_bf16(dgp, dg);                    // dL/dgamma
        store16_bf16(dgp + p.Cout, v);            // dL/dbeta
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= 1.f + o.g[j];
      }
      store16_bf16(p.yp + ((long long)b * p.tp_out + t + p.out_halo) * p.cp_out + p.out_ch_off + grp * p.out_ch_stride + ch, v);
    } else if (MODE == 6) {
      mask16_words(o.m0, o.m1, v, p.mask_slope);
      const int tt = t - p.halo;
      if (tt >= 0 && tt < p.t_valid) {
        const long long ct = p.t_valid;
        float* yo = p.y + (long long)grp * p.y_grp_stride + (long long)b * p.y_b_stride + (long long)ch * ct + tt;
        if (p.res) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += o.g[j];
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) yo[j * ct] = v[j];
        if (p.yp) store16_bf16(p.yp + ((long long)b * p.tp_out + tt + p.out_halo) * p.cp_out + p.out_ch_off + grp * p.out_ch_stride + ch, v);
      } else {
        const int hr = t < p.halo ? t : t - p.t_valid;                 // 0 .. 2*halo - 1
        float* hb = p.halo_buf + (((long long)grp * p.B + b) * p.Cout + ch) * (2 * p.halo) + hr;
#pragma unroll
        for (int j = 0; j < 16; ++j) hb[(long long)j * 2 * p.halo] = v[j];
      }
    }
  }
}

// One 16-column chunk (columns c0 .. c0+15 of the tile, TMEM lane quadrant q) of the plain epilogues; m0 | m1 are the
// chunk's mask words (MASK == 1), loaded by the caller.
template <int ACT, int EPI, int OUT, int MASK>
__device__ __forceinline__ void tc_epilogue_chunk(const TcP& p, const float* bias_s, uint32_t acc, int b, int t, bool t_ok,
                                                  int grp, int n0, int c0, int nvalid, int q, const uint4& m0, const uint4& m1) {
  const long long ct = p.Tout;
  float v[16];
  if (p.debug & 4) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (float)(c0 + j);
  } else {
    tmem_ld16(acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
  }
  if (!t_ok) return;
  const int nj = min(16, nvalid - c0);
  {
    const float4* b4 = reinterpret_cast<const float4*>(bias_s + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 bb = b4[j];
      v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
    }
  }
  if (MASK) mask16_words(m0, m1, v, p.mask_slope);
  if (OUT == 0 && p.flat_tp > 0) {
    // batch-flattened rows: t is a row of the concatenated padded samples
    const int bb = t / p.flat_tp, tt = t - bb * p.flat_tp - p.flat_halo;
    if (tt >= 0 && tt < p.flat_T) {
      float* yp = p.y + ((long long)bb * p.Cout + n0 + c0) * p.flat_T + tt;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < nj) {
          float o = v[j];
          if (ACT == TDVC_ACT_LRELU) o = o > 0.f ? o : o * p.out_slope;
          else if (ACT == TDVC_ACT_TANH) o = tanhf(o);
          *yp = o;
        }
        yp += p.flat_T;
      }
    }
  } else if (OUT == 0 && p.unframe_s > 0) {
    // data gradient of a strided conv run over frames: this thread's 16 frame channels are 16 / s conv channels x s
    // consecutive samples
    const int s = p.unframe_s;
    const int f0 = grp * p.Cout + n0 + c0;
    const int u0 = s * t - p.unframe_pad;
    // whole frames inside the signal and 16-byte aligned rows: vector stores (a warp writes 32 consecutive frames of
    // one channel = one contiguous run); scalar stores at the edges
    const bool vec = (s % 4 == 0) && (p.unframe_T % 4 == 0) && (p.unframe_pad % 4 == 0) && u0 >= 0 && u0 + s <= p.unframe_T;
#pragma unroll
    for (int j0 = 0; j0 < 16; j0 += 4) {
      if (j0 >= nj) break;
      float* dst = p.y + ((long long)b * p.unframe_C + (f0 + j0) / s) * p.unframe_T + u0 + (j0 % s);
      if (vec) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[j0], v[j0 + 1], v[j0 + 2], v[j0 + 3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int jj = j0 + e;                                     // frame channel f0 + jj: conv channel, phase
          const int u = u0 + jj % s;
          if (jj < nj && u >= 0 && u < p.unframe_T)
            p.y[((long long)b * p.unframe_C + (f0 + jj) / s) * p.unframe_T + u] = v[jj];
        }
      }
    }
  } else if (OUT == 0) {
    const long long base = (long long)grp * p.y_grp_stride + (long long)b * p.y_b_stride + (long long)(n0 + c0) * ct + t;
    float* yp = p.y + base;
    const float* rp = p.res + base;                                   // only dereferenced when EPI asks for it
    const float* gp = p.gb + ((long long)b * 2 * p.Cout + n0 + c0) * ct + t;
    const long long beta_off = (long long)p.Cout * ct;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < nj) {
        float o = v[j];
        if (EPI == 2) {
          o = fmaf(o, 1.f + __ldg(gp), __ldg(gp + beta_off));
          if (p.res) o += __ldg(rp);
        } else if (EPI == 1) {
          o += __ldg(rp);
        }
        if (ACT == TDVC_ACT_LRELU) o = o > 0.f ? o : o * p.out_slope;
        else if (ACT == TDVC_ACT_TANH) o = tanhf(o);
        if (!(p.debug & 1)) *yp = o;
      }
      yp += ct; rp += ct; gp += ct;
    }
  } else {
    // packed output: this thread owns 16 consecutive channels of one time step = 32 contiguous bytes
    if (ACT == TDVC_ACT_LRELU) {
      const float sl = p.out_slope;       // 0 < slope < 1: leaky_relu(o) = max(o, slope * o)
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], v[j] * sl);
    }
    if (nj < 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = (j < nj) ? v[j] : 0.f;
    }
    __nv_bfloat16* op = p.yp + ((long long)b * p.tp_out + t + p.out_halo) * p.cp_out + p.out_ch_off +
                        grp * p.out_ch_stride + n0 + c0;
    if (!(p.debug & 1)) store16_bf16(op, v);
  }
}

// Epilogue of one 128 x BN accumulator tile at TMEM address `acc` (lane quadrant q, column half `half`); wait_acc() blocks
// until the accumulator is complete and is called once, by every thread, after the operand loads that can run ahead.
template <int ACT, int EPI, int OUT, int MASK, typename Wait>
__device__ __forceinline__ void tc_epilogue_tile(const TcP& p, const float* bias_s, uint32_t acc, int b, int t0, int grp,
                                                 int n0, int q, int half, int lane, Wait wait_acc) {
  if (EPI >= 3) {
    tc_epilogue_chain<EPI>(p, bias_s, acc, b, t0, grp, n0, q, half, lane, wait_acc);
    return;
  }
  const int t = t0 + q * 32 + lane;
  const bool t_ok = t < p.Tout;
  const int nvalid = min(p.BN, p.Cout - n0);          // columns of this tile that are real channels
  const int step = 16 * epi_spt_of(p), first = (half % epi_spt_of(p)) * 16;
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  if (MASK && TDVC_EPI_HOIST) {
    const __nv_bfloat16* mrow = p.maskp + ((long long)b * p.tm + t + p.mask_halo) * p.cm + p.mask_ch_off +
                                grp * p.mask_ch_stride + n0;
    constexpr int PF = EpiPf<OUT>::n;
    uint4 mk[PF][2];
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      mk[i][0] = z4; mk[i][1] = z4;
      const int c0 = first + i * step;
      if (t_ok && c0 < nvalid) ld32B_bf16(mrow + c0, mk[i][0], mk[i][1]);
    }
    wait_acc();
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int c0 = first + i * step;
      if (c0 < nvalid) tc_epilogue_chunk<ACT, EPI, OUT, MASK>(p, bias_s, acc, b, t, t_ok, grp, n0, c0, nvalid, q, mk[i][0], mk[i][1]);
    }
    for (int c0 = first + PF * step; c0 < nvalid; c0 += step) {
      uint4 m0 = z4, m1 = z4;
      if (t_ok) ld32B_bf16(mrow + c0, m0, m1);
      tc_epilogue_chunk<ACT, EPI, OUT, MASK>(p, bias_s, acc, b, t, t_ok, grp, n0, c0, nvalid, q, m0, m1);
    }
  } else {
    wait_acc();
    for (int c0 = first; c0 < nvalid; c0 += step) {
      uint4 m0 = z4, m1 = z4;
      if (MASK && t_ok) {
        const __nv_bfloat16* mp = p.maskp + ((long long)b * p.tm + t + p.mask_halo) * p.cm + p.mask_ch_off +
                                  grp * p.mask_ch_stride + n0 + c0;
        ld32B_bf16(mp, m0, m1);
      }
      tc_epilogue_chunk<ACT, EPI, OUT, MASK>(p, bias_s, acc, b, t, t_ok, grp, n0, c0, nvalid, q, m0, m1);
    }
  }
}

// ACT: tdvc_act of the epilogue; EPI: 0 = bias only, 1 = + residual, 2 = FiLM (+ residual when p.res);
// OUT: 0 = fp32 NCW [groups][B][Cout][T], 1 = bf16 channels-last (the next conv's operand, no pack pass);
// MASK: 1 = multiply by the LeakyReLU derivative taken from the sign of a packed activated tensor (dgrad).
template <int ACT, int EPI, int OUT, int MASK>
__global__ void __launch_bounds__(TC_FWD_THREADS) conv_tc_fwd_k(const __grid_constant__ CUtensorMap map_a,
                                                                const __grid_constant__ CUtensorMap map_b, TcP p) {
  pdl_prologue_top();
  __shared__ __align__(16) float bias_s[256];      // bias of this N tile (zeros when absent)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B atoms need 1024-B alignment
  const int b_bytes = p.BN * TC_BK * 2;
  const int stage_bytes = TC_A_BYTES + b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * TC_BM;
  const int grp = blockIdx.y / p.tiles_per_group;
  const int n0 = (blockIdx.y - grp * p.tiles_per_group) * p.BN;       // channel offset inside the group
  const int b = blockIdx.z;
  // a group may use fewer taps than the weight tensor holds (the k = 3 / 7 branches next to k = 11): the centred ones
  const int kgrp = p.kg[grp & 3] > 0 ? p.kg[grp & 3] : p.K;
  const int tap_lo = (p.K - kgrp) >> 1;
  const int iters = (p.debug & 2) ? 0 : kgrp * p.nchunk;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1 && !(p.debug & 16)) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  pdl_prologue_late();
  for (int i = threadIdx.x; i < p.BN; i += TC_FWD_THREADS)
    bias_s[i] = (p.bias && n0 + i < p.Cout) ? __ldg(p.bias + grp * p.bias_stride + n0 + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = (p.debug & 16) ? 0u : *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        const int tl = it / p.nchunk, ck = it - tl * p.nchunk;
        const int tap = tap_lo + tl;
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
        tma_load_3d(sa, &map_a, &full_bar[s], p.a_ch_off + grp * p.a_ch_stride + ck * TC_BK, t0 + tap * p.dil + p.t_off, b);
        tma_load_3d(sa + TC_A_BYTES, &map_b, &full_bar[s], ck * TC_BK, grp * p.coutp_g + n0, tap);
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (one elected thread)
    // instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7, 10), both K-major, N>>3 at 17, M>>4 at 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    for (int it = 0; it < iters; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const int ck = it % p.nchunk;
        const int nk = (ck == p.nchunk - 1) ? p.last_nk16 : (TC_BK / 16);
        const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
        const uint64_t da = make_sw128_kmajor_desc(a_addr);
        const uint64_t db = make_sw128_kmajor_desc(a_addr + TC_A_BYTES);
        for (int k = 0; k < nk; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);                 // frees the smem stage when these MMAs retire
        if (it == iters - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else if (!(p.debug & 8)) {
    // ---------------- epilogue: warps 2..9; TMEM lane quadrant = warp % 4, the two warps of a quadrant split
    // the 16-column chunks (even / odd).  A lane is a time step: for a fixed channel the warp stores 32
    // consecutive floats (one 128-byte line).
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    auto wait_acc = [&]() {
      if (iters > 0) mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    };
    tc_epilogue_tile<ACT, EPI, OUT, MASK>(p, bias_s, tmem_base, b, t0, grp, n0, q, half, lane, wait_acc);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1 && !(p.debug & 16)) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------ weight-stationary forward
// Persistent variant for convolutions whose whole weight tile set (K taps x Cin chunks x BN rows) fits in shared
// memory next to a small activation ring -- FiLM cond_var.0 (9 x 18 KB) and most cond_var.2 / conv.1 layers:
//   * a CTA owns one N tile and walks a strided list of 128-step time tiles: weights are fetched from L2 ONCE per
//     CTA instead of once per tile (the non-persistent kernel re-reads 166 KB of weights per 128 x 144 tile);
//   * the activation tile is loaded once per 64-channel chunk with its (K-1)*dilation halo rows; the K taps are
//     row-shifted views of it (descriptor start address + tap*dilation*128 B, swizzle phase carried in the
//     descriptor's base-offset field), so activations cross L2->SMEM once instead of K times;
//   * two TMEM accumulators: the 8 epilogue warps drain tile i while the MMA warp computes tile i+1.
struct WsP {
  int n_mtiles, mtiles_per_b, rows_a, a_stage_bytes, w_tile_bytes;
  // narrow tail: when the reduction width leaves a 16-channel last chunk (144 = 64 + 64 + 16) that chunk travels as
  // 32-byte rows (SWIZZLE_32B boxes, own small ring) instead of a zero-padded 128-byte one: 4x less shared memory
  // for its weights and activations, which buys a deeper activation ring (the kernel is load-latency bound).
  int narrow, n_full, an_stage_bytes, wn_tile_bytes, n_stages;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major SWIZZLE_32B descriptor: 32-byte rows (16 bf16), 8-row atoms 256 B apart.
__device__ __forceinline__ uint64_t make_sw32_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}

template <int ACT, int EPI, int OUT, int MASK>
__global__ void __launch_bounds__(TC_FWD_THREADS, 1) conv_tc_ws_k(const __grid_constant__ CUtensorMap map_a,
                                                                  const __grid_constant__ CUtensorMap map_b,
                                                                  const __grid_constant__ CUtensorMap map_an,
                                                                  const __grid_constant__ CUtensorMap map_bn, TcP p, WsP w) {
  pdl_prologue_top();
  __shared__ __align__(16) float bias_s[256];
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nfull = w.n_full;                                      // 64-channel chunks
  uint8_t* wsm = smem;                                             // [K][nfull] tiles of BN x 128 B
  uint8_t* wsn = wsm + (size_t)p.K * nfull * w.w_tile_bytes;       // [K] tiles of BN x 32 B (narrow tail)
  uint8_t* ring = wsn + (size_t)(w.narrow ? p.K : 0) * w.wn_tile_bytes;          // [stages] x rows_a x 128 B
  uint8_t* ringn = ring + (size_t)p.stages * w.a_stage_bytes;                   // [n_stages] x rows_a x 32 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(ringn + (size_t)(w.narrow ? w.n_stages : 0) * w.an_stage_bytes);
  uint64_t* w_full = bars;
  uint64_t* full_bar = bars + 1;
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* fulln_bar = empty_bar + p.stages;
  uint64_t* emptyn_bar = fulln_bar + w.n_stages;
  const int spt = epi_spt_of(p), nteams = (TC_EPI_WARPS / 4) / spt, nacc = p.nacc > 0 ? p.nacc : 2;
  uint64_t* tmem_full = emptyn_bar + w.n_stages;   // [nacc]
  uint64_t* tmem_empty = tmem_full + nacc;         // [nacc]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + nacc);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int grp, n0, cta_i, cta_n;          // this CTA's group / N tile, and its position among the CTAs that share them
  if (p.balanced) {
    grp = 0;
    while (grp < 3 && (int)blockIdx.x >= p.grp_cta0[grp + 1]) ++grp;
    n0 = 0;
    cta_i = (int)blockIdx.x - p.grp_cta0[grp];
    cta_n = p.grp_cta0[grp + 1] - p.grp_cta0[grp];
  } else {
    grp = blockIdx.y / p.tiles_per_group;
    n0 = (blockIdx.y - grp * p.tiles_per_group) * p.BN;
    cta_i = blockIdx.x;
    cta_n = gridDim.x;
  }

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    mbar_init(w_full, 1);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < w.n_stages; ++s) {
      mbar_init(&fulln_bar[s], 1);
      mbar_init(&emptyn_bar[s], 1);
    }
    for (int a = 0; a < nacc; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4 * spt);               // one arrival per epilogue warp of the team that drains it
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  pdl_prologue_late();
  for (int i = threadIdx.x; i < p.BN; i += TC_FWD_THREADS)
    bias_s[i] = (p.bias && n0 + i < p.Cout) ? __ldg(p.bias + grp * p.bias_stride + n0 + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int a_ch0 = p.a_ch_off + grp * p.a_ch_stride;
  const int kgrp = p.kg[grp & 3] > 0 ? p.kg[grp & 3] : p.K;      // this group's taps: the centred kgrp of the K in the tensor
  const int tap_lo = (p.K - kgrp) >> 1, tap_hi = tap_lo + kgrp;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(w_full, (uint32_t)(kgrp * nfull * w.w_tile_bytes + (w.narrow ? kgrp * w.wn_tile_bytes : 0)));
      for (int tap = tap_lo; tap < tap_hi; ++tap) {
        for (int ck = 0; ck < nfull; ++ck)
          tma_load_3d(wsm + (size_t)(tap * nfull + ck) * w.w_tile_bytes, &map_b, w_full, ck * TC_BK, grp * p.coutp_g + n0, tap);
        if (w.narrow)
          tma_load_3d(wsn + (size_t)tap * w.wn_tile_bytes, &map_bn, w_full, nfull * TC_BK, grp * p.coutp_g + n0, tap);
      }
      int it = 0, itn = 0;
      for (int m = cta_i; m < w.n_mtiles; m += cta_n) {
        const int b = m / w.mtiles_per_b, t0 = (m - b * w.mtiles_per_b) * TC_BM;
        for (int ck = 0; ck < nfull; ++ck, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], (uint32_t)w.a_stage_bytes);
          tma_load_3d(ring + (size_t)s * w.a_stage_bytes, &map_a, &full_bar[s], a_ch0 + ck * TC_BK, t0 + p.t_off, b);
        }
        if (w.narrow) {
          const int s = itn % w.n_stages;
          const uint32_t ph = (uint32_t)(itn / w.n_stages) & 1u;
          mbar_wait(&emptyn_bar[s], ph ^ 1u);
          mbar_expect_tx(&fulln_bar[s], (uint32_t)w.an_stage_bytes);
          tma_load_3d(ringn + (size_t)s * w.an_stage_bytes, &map_an, &fulln_bar[s], a_ch0 + nfull * TC_BK, t0 + p.t_off, b);
          ++itn;
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    mbar_wait(w_full, 0);
    tc_fence_after();
    // descriptor constants (units of 16 bytes for the address field)
    const uint64_t d128 = make_sw128_kmajor_desc(0), d32 = make_sw32_kmajor_desc(0);
    const uint32_t hi128 = (uint32_t)(d128 >> 32), hi32 = (uint32_t)(d32 >> 32);
    const uint32_t lo_flags128 = (uint32_t)d128, lo_flags32 = (uint32_t)d32;       // LBO field lives in the low word
    const uint32_t a_lo0 = lo_flags128 | ((smem_u32(ring) & 0x3FFFF) >> 4);
    const uint32_t b_lo0 = lo_flags128 | ((smem_u32(wsm) & 0x3FFFF) >> 4);
    const uint32_t an_lo0 = lo_flags32 | ((smem_u32(ringn) & 0x3FFFF) >> 4);
    const uint32_t bn_lo0 = lo_flags32 | ((smem_u32(wsn) & 0x3FFFF) >> 4);
    const uint32_t a_stage16 = (uint32_t)w.a_stage_bytes >> 4, an_stage16 = (uint32_t)w.an_stage_bytes >> 4;
    const uint32_t w_tile16 = (uint32_t)w.w_tile_bytes >> 4, wn_tile16 = (uint32_t)w.wn_tile_bytes >> 4;
    const uint32_t w_tap16 = w_tile16 * (uint32_t)nfull;                 // next tap's tile of the same chunk
    const uint32_t tap_step16 = (uint32_t)p.dil * 8u;                    // dil rows x 128 B
    const uint32_t tapn_step16 = (uint32_t)p.dil * 2u;                   // dil rows x 32 B
    int it = 0, itn = 0, i = 0;
    for (int m = cta_i; m < w.n_mtiles; m += cta_n, ++i) {
      const int acc = i % nacc;
      const uint32_t d_addr = tmem_base + (uint32_t)(acc * p.BN);
      mbar_wait(&tmem_empty[acc], ((uint32_t)(i / nacc) & 1u) ^ 1u);
      tc_fence_after();
      for (int ck = 0; ck < nfull; ++ck, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          // the single issuing thread must spend < ~70 cycles per MMA (its execution time at N=144): descriptors are
          // advanced with 32-bit adds on their low words only
          const int nk = (!w.narrow && ck == nfull - 1) ? p.last_nk16 : (TC_BK / 16);
          uint32_t a_lo = a_lo0 + (uint32_t)s * a_stage16 + (uint32_t)tap_lo * tap_step16;
          uint32_t b_lo = b_lo0 + (uint32_t)ck * w_tile16 + (uint32_t)tap_lo * w_tap16;
          for (int tap = tap_lo; tap < tap_hi; ++tap) {
            // row-shifted view of the haloed tile: the 128B swizzle is a function of the absolute shared-memory
            // address, so a start address moved by whole 128-byte rows needs no descriptor fix-up (verified on B200)
            umma_bf16_lohi(d_addr, a_lo, hi128, b_lo, hi128, idesc, (ck > 0 || tap > tap_lo) ? 1u : 0u);
            if (nk > 1) umma_bf16_lohi(d_addr, a_lo + 2, hi128, b_lo + 2, hi128, idesc, 1u);
            if (nk > 2) umma_bf16_lohi(d_addr, a_lo + 4, hi128, b_lo + 4, hi128, idesc, 1u);
            if (nk > 3) umma_bf16_lohi(d_addr, a_lo + 6, hi128, b_lo + 6, hi128, idesc, 1u);
            a_lo += tap_step16;
            b_lo += w_tap16;
          }
          umma_commit(&empty_bar[s]);
          if (!w.narrow && ck == nfull - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
      }
      if (w.narrow) {
        const int s = itn % w.n_stages;
        const uint32_t ph = (uint32_t)(itn / w.n_stages) & 1u;
        mbar_wait(&fulln_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          uint32_t a_lo = an_lo0 + (uint32_t)s * an_stage16 + (uint32_t)tap_lo * tapn_step16;
          uint32_t b_lo = bn_lo0 + (uint32_t)tap_lo * wn_tile16;
          for (int tap = tap_lo; tap < tap_hi; ++tap) {
            umma_bf16_lohi(d_addr, a_lo, hi32, b_lo, hi32, idesc, 1u);
            a_lo += tapn_step16;
            b_lo += wn_tile16;
          }
          umma_commit(&emptyn_bar[s]);
          umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        ++itn;
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int team = half / spt;
    int i = 0;
    for (int m = cta_i; m < w.n_mtiles; m += cta_n, ++i) {
      if (i % nteams != team) continue;               // another team's tile
      const int b = m / w.mtiles_per_b, t0 = (m - b * w.mtiles_per_b) * TC_BM;
      const int acc = i % nacc;
      const uint32_t acc_phase = (uint32_t)(i / nacc) & 1u;
      auto wait_acc = [&]() {
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
      };
      if (!(p.debug & 32))
        tc_epilogue_tile<ACT, EPI, OUT, MASK>(p, bias_s, tmem_base + (uint32_t)(acc * p.BN), b, t0, grp, n0, q, half, lane, wait_acc);
      else
        wait_acc();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------ streamed weights, haloed tile
// conv_tc_fwd_k for K > 1 with the activation side of conv_tc_ws_k: per 64-channel chunk the 128-step tile is loaded once
// with its (K-1)*dil halo rows and the K taps are row-shifted descriptor views of it, while the weight tiles (too large to be
// resident: the 1024-channel discriminator layers, the 1296-channel data gradient of cond_var.0, the C >= 128 MRF stages)
// stream through their own ring.  conv_tc_fwd_k re-loads the activation tile for every tap: (16 KB + BN x 128 B) per 4 MMAs
// is 66 B/cycle at BN = 144 -- the L2 -> SM port, not the tensor core, set its pace (428 us for the cond_var.0 data gradient
// against 230 us of MMA issue).
template <int ACT, int EPI, int OUT, int MASK>
__global__ void __launch_bounds__(TC_FWD_THREADS) conv_tc_fwdh_k(const __grid_constant__ CUtensorMap map_a,
                                                                 const __grid_constant__ CUtensorMap map_b, TcP p) {
  pdl_prologue_top();
  __shared__ __align__(16) float bias_s[256];
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int A_ST = 2;
  const int a_bytes = p.a_rows * 128, b_bytes = p.BN * TC_BK * 2;
  uint8_t* ring_a = smem;
  uint8_t* ring_b = smem + (size_t)A_ST * a_bytes;
  uint64_t* full_a = reinterpret_cast<uint64_t*>(ring_b + (size_t)p.b_stages * b_bytes);
  uint64_t* empty_a = full_a + A_ST;
  uint64_t* full_b = empty_a + A_ST;
  uint64_t* empty_b = full_b + p.b_stages;
  uint64_t* tmem_full_bar = empty_b + p.b_stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * TC_BM;
  const int grp = blockIdx.y / p.tiles_per_group;
  const int n0 = (blockIdx.y - grp * p.tiles_per_group) * p.BN;
  const int b = blockIdx.z;
  const int kgrp = p.kg[grp & 3] > 0 ? p.kg[grp & 3] : p.K;
  const int tap_lo = (p.K - kgrp) >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < A_ST; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  pdl_prologue_late();
  for (int i = threadIdx.x; i < p.BN; i += TC_FWD_THREADS)
    bias_s[i] = (p.bias && n0 + i < p.Cout) ? __ldg(p.bias + grp * p.bias_stride + n0 + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int itb = 0;
      for (int ck = 0; ck < p.nchunk; ++ck) {
        const int sa = ck % A_ST;
        mbar_wait(&empty_a[sa], ((uint32_t)(ck / A_ST) & 1u) ^ 1u);
        mbar_expect_tx(&full_a[sa], (uint32_t)a_bytes);
        tma_load_3d(ring_a + (size_t)sa * a_bytes, &map_a, &full_a[sa], p.a_ch_off + grp * p.a_ch_stride + ck * TC_BK, t0 + p.t_off, b);
        for (int tl = 0; tl < kgrp; ++tl, ++itb) {
          const int sb = itb % p.b_stages;
          mbar_wait(&empty_b[sb], ((uint32_t)(itb / p.b_stages) & 1u) ^ 1u);
          mbar_expect_tx(&full_b[sb], (uint32_t)b_bytes);
          tma_load_3d(ring_b + (size_t)sb * b_bytes, &map_b, &full_b[sb], ck * TC_BK, grp * p.coutp_g + n0, tap_lo + tl);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint64_t d128 = make_sw128_kmajor_desc(0);
    const uint32_t hi128 = (uint32_t)(d128 >> 32), lo_flags = (uint32_t)d128;
    const uint32_t a_lo0 = lo_flags | ((smem_u32(ring_a) & 0x3FFFF) >> 4);
    const uint32_t b_lo0 = lo_flags | ((smem_u32(ring_b) & 0x3FFFF) >> 4);
    const uint32_t a_stage16 = (uint32_t)a_bytes >> 4, b_stage16 = (uint32_t)b_bytes >> 4;
    const uint32_t tap_step16 = (uint32_t)p.dil * 8u;                    // dil rows x 128 B
    int itb = 0;
    for (int ck = 0; ck < p.nchunk; ++ck) {
      const int sa = ck % A_ST;
      mbar_wait(&full_a[sa], (uint32_t)(ck / A_ST) & 1u);
      tc_fence_after();
      const int nk = (ck == p.nchunk - 1) ? p.last_nk16 : (TC_BK / 16);
      for (int tl = 0; tl < kgrp; ++tl, ++itb) {
        const int sb = itb % p.b_stages;
        mbar_wait(&full_b[sb], (uint32_t)(itb / p.b_stages) & 1u);
        tc_fence_after();
        if (elect_one()) {
          // row-shifted view of the haloed tile (the 128B swizzle is a function of the absolute shared-memory address)
          const uint32_t a_lo = a_lo0 + (uint32_t)sa * a_stage16 + (uint32_t)(tap_lo + tl) * tap_step16;
          const uint32_t b_lo = b_lo0 + (uint32_t)sb * b_stage16;
          umma_bf16_lohi(tmem_base, a_lo, hi128, b_lo, hi128, idesc, (ck > 0 || tl > 0) ? 1u : 0u);
          if (nk > 1) umma_bf16_lohi(tmem_base, a_lo + 2, hi128, b_lo + 2, hi128, idesc, 1u);
          if (nk > 2) umma_bf16_lohi(tmem_base, a_lo + 4, hi128, b_lo + 4, hi128, idesc, 1u);
          if (nk > 3) umma_bf16_lohi(tmem_base, a_lo + 6, hi128, b_lo + 6, hi128, idesc, 1u);
          umma_commit(&empty_b[sb]);
          if (tl == kgrp - 1) {
            umma_commit(&empty_a[sa]);
            if (ck == p.nchunk - 1) umma_commit(tmem_full_bar);
          }
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    auto wait_acc = [&]() {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    };
    tc_epilogue_tile<ACT, EPI, OUT, MASK>(p, bias_s, tmem_base, b, t0, grp, n0, q, half, lane, wait_acc);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------ stacked-block forward
// The GEMM turned round for convolutions whose output-channel dimension is the long one -- the n FiLM cond_var.0 convs of
// an MRF stage read the same conditioning tensor, so their weights stack into one [n*136, 144*3] matrix:
//   D[co, t] = sum_{tap} sum_{ci} W_tap[co, ci] * X[t + tap*dil + t_off, ci]
//   A operand = weights, M = 128 stacked output channels (TMEM lanes), loaded once per CTA;
//   B operand = the channels-last activation tile (time steps = TMEM columns), taps = row-shifted views.
// Measured on B200: an SS-mode M=128 tcgen05.mma takes ~120 cycles for any N <= 144 (the A-operand read from shared
// memory), so the time-as-M orientation tops out near 55 % of peak at N=144.  Two forms of the turned-round GEMM:
//   TS = 0  weights resident in shared memory, N = 256 per MMA (the N at which an SS MMA reaches its 128-cycle floor);
//           A + B operand reads are 96 B/cycle of shared-memory bandwidth, which the epilogue's staging tile competes for;
//   TS = 1  weights resident in TENSOR MEMORY (K*Cin/2 <= 256 columns next to two 128-column accumulators, written once
//           with tcgen05.st): the MMA reads only B from shared memory (64 B/cycle), N = 128 already runs at the tensor
//           pipe's floor (A-in-TMEM floor = 128*N/256 cycles), and the 110 KB the weights occupied become ring stages.
// The stacked rows are dense (no per-block padding), so 9 x 136 = 1224 rows are 10 M tiles instead of 11.  Ownership is
// fixed -- CTA c keeps M tile c % m_tiles and walks time tiles c / m_tiles + i * ctas_per_m -- so the m_tiles CTAs that
// need the same activation tile ask for it at the same moment (one HBM read, the rest L2 hits).  A TMEM lane is an output
// channel: bias is one register; the packed channels-last output goes through a 1 KB per-warp staging tile so that global
// stores are 16 bytes per lane (2- and 4-byte stores of 64-byte row segments ran 3-4x slower).
struct WtP {
  int n_ttiles, ttiles_per_b, m_tiles, ctas_per_m;
  int nsplit, box_rows, a_stage_bytes, an_stage_bytes;      // activation stage = nsplit TMA boxes of box_rows rows
  int narrow, n_full, n_stages, last_nk16;
  int rows_total, rows_per_block;                           // dense stacked rows; block j = rows [j*rpb, (j+1)*rpb)
  int vec;                                                  // 16-byte output stores through the per-warp staging tile
  int R, cinp;                                              // TS: wp is [K][R][cinp]
  const __nv_bfloat16* wp;
};
constexpr int WT_W_TILE = TC_BM * TC_BK * 2;      // 128 rows x 128 B
constexpr int WT_WN_TILE = TC_BM * 32;            // 128 rows x 32 B (narrow tail)
constexpr int WT_STG_BYTES = 16 * 32 * 2;         // per-warp output staging: 16 time steps x 32 channels bf16
constexpr int WT_TS_WCOL = 256;                   // TS: first TMEM column of the weights (after two 128-column accumulators)

__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

template <int ACT, int TS>
__global__ void __launch_bounds__(TC_FWD_THREADS, 1) conv_tc_wt_k(const __grid_constant__ CUtensorMap map_x,
                                                                  const __grid_constant__ CUtensorMap map_w,
                                                                  const __grid_constant__ CUtensorMap map_xn,
                                                                  const __grid_constant__ CUtensorMap map_wn, TcP p, WtP w) {
  pdl_prologue();
  constexpr int BN = TS ? 128 : 256;                // time steps per tile = TMEM columns per accumulator
  constexpr int NG = BN / 64;                       // 16-column groups per epilogue warp (4 column quarters x NG x 16)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nfull = w.n_full;
  uint8_t* wsm = smem;                                                        // SS: [K][nfull] tiles of 128 x 128 B
  uint8_t* wsn = wsm + (TS ? 0 : (size_t)p.K * nfull * WT_W_TILE);            // SS: [K] tiles of 128 x 32 B
  uint8_t* ring = wsn + (TS ? 0 : (size_t)(w.narrow ? p.K : 0) * WT_WN_TILE); // [stages] activation chunks, 128-B rows
  uint8_t* ringn = ring + (size_t)p.stages * w.a_stage_bytes;                 // [n_stages] narrow chunks, 32-B rows
  uint8_t* stage_out = ringn + (size_t)(w.narrow ? w.n_stages : 0) * w.an_stage_bytes;   // [TC_EPI_WARPS] x 1 KB (vec epilogue)
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + (w.vec ? TC_EPI_WARPS * WT_STG_BYTES : 0));
  uint64_t* w_full = bars;
  uint64_t* full_bar = bars + 1;
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* fulln_bar = empty_bar + p.stages;
  uint64_t* emptyn_bar = fulln_bar + w.n_stages;
  uint64_t* tmem_full = emptyn_bar + w.n_stages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x % w.m_tiles;
  const int slot0 = blockIdx.x / w.m_tiles;
  const int r0 = m_tile * TC_BM;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    if (!TS) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    mbar_init(w_full, 1);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < w.n_stages; ++s) {
      mbar_init(&fulln_bar[s], 1);
      mbar_init(&emptyn_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], TC_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (TS) {
    // weights -> tensor memory, once: lane (= stacked row) x column (= two consecutive input channels of one tap).
    // Warps 2..5 cover the four lane quadrants; each thread copies its own row, 16 channels (32 B) per tcgen05.st.
    if (warp >= 2 && warp < 6) {
      const int q = warp & 3;
      const int r = r0 + q * 32 + lane;
      const int kcols = w.cinp >> 1;                               // TMEM columns per tap
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)WT_TS_WCOL;
      for (int tap = 0; tap < p.K; ++tap) {
        const uint4* src = reinterpret_cast<const uint4*>(w.wp + ((long long)tap * w.R + r) * w.cinp);
        for (int c8 = 0; c8 < (w.cinp >> 4); ++c8) {
          uint32_t v[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          if (r < w.rows_total) {
            const uint4 lo = __ldg(src + 2 * c8), hi = __ldg(src + 2 * c8 + 1);
            v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
          }
          tmem_st8(tbase + (uint32_t)(tap * kcols + c8 * 8), v);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (warp == 0) {
    if (elect_one()) {
      if (!TS) {
        mbar_expect_tx(w_full, (uint32_t)(p.K * nfull * WT_W_TILE + (w.narrow ? p.K * WT_WN_TILE : 0)));
        for (int tap = 0; tap < p.K; ++tap) {
          for (int ck = 0; ck < nfull; ++ck)
            tma_load_3d(wsm + (size_t)(tap * nfull + ck) * WT_W_TILE, &map_w, w_full, ck * TC_BK, r0, tap);
          if (w.narrow) tma_load_3d(wsn + (size_t)tap * WT_WN_TILE, &map_wn, w_full, nfull * TC_BK, r0, tap);
        }
      }
      int it = 0, itn = 0;
      for (int j = slot0; j < w.n_ttiles; j += w.ctas_per_m) {
        const int b = j / w.ttiles_per_b, t0 = (j - b * w.ttiles_per_b) * BN + p.t_off;
        if (!TS) {
          // warm L2 with the tile after this one: the SS ring only looks one tile ahead, too short for an HBM miss
          const int jn = j + w.ctas_per_m;
          if (jn < w.n_ttiles && m_tile == 0 && !(p.debug & 64)) {
            const int bn = jn / w.ttiles_per_b, tn = (jn - bn * w.ttiles_per_b) * BN + p.t_off;
            for (int ck = 0; ck < nfull; ++ck)
              for (int h = 0; h < w.nsplit; ++h) tma_prefetch_3d(&map_x, p.a_ch_off + ck * TC_BK, tn + h * w.box_rows, bn);
            if (w.narrow)
              for (int h = 0; h < w.nsplit; ++h) tma_prefetch_3d(&map_xn, p.a_ch_off + nfull * TC_BK, tn + h * w.box_rows, bn);
          }
        }
        for (int ck = 0; ck < nfull; ++ck, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], (uint32_t)w.a_stage_bytes);
          for (int h = 0; h < w.nsplit; ++h)
            tma_load_3d(ring + (size_t)s * w.a_stage_bytes + (size_t)h * w.box_rows * 128, &map_x, &full_bar[s],
                        p.a_ch_off + ck * TC_BK, t0 + h * w.box_rows, b);
        }
        if (w.narrow) {
          const int s = itn % w.n_stages;
          const uint32_t ph = (uint32_t)(itn / w.n_stages) & 1u;
          mbar_wait(&emptyn_bar[s], ph ^ 1u);
          mbar_expect_tx(&fulln_bar[s], (uint32_t)w.an_stage_bytes);
          for (int h = 0; h < w.nsplit; ++h)
            tma_load_3d(ringn + (size_t)s * w.an_stage_bytes + (size_t)h * w.box_rows * 32, &map_xn, &fulln_bar[s],
                        p.a_ch_off + nfull * TC_BK, t0 + h * w.box_rows, b);
          ++itn;
        }
      }
    }
  } else if (warp == 1) {
    // M = 128 weight rows, N = BN time steps
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    if (!TS) {
      mbar_wait(w_full, 0);
      tc_fence_after();
    }
    const uint64_t d128 = make_sw128_kmajor_desc(0), d32 = make_sw32_kmajor_desc(0);
    const uint32_t hi128 = (uint32_t)(d128 >> 32), hi32 = (uint32_t)(d32 >> 32);
    const uint32_t lo_flags128 = (uint32_t)d128, lo_flags32 = (uint32_t)d32;
    const uint32_t x_lo0 = lo_flags128 | ((smem_u32(ring) & 0x3FFFF) >> 4);
    const uint32_t w_lo0 = lo_flags128 | ((smem_u32(wsm) & 0x3FFFF) >> 4);
    const uint32_t xn_lo0 = lo_flags32 | ((smem_u32(ringn) & 0x3FFFF) >> 4);
    const uint32_t wn_lo0 = lo_flags32 | ((smem_u32(wsn) & 0x3FFFF) >> 4);
    const uint32_t a_stage16 = (uint32_t)w.a_stage_bytes >> 4, an_stage16 = (uint32_t)w.an_stage_bytes >> 4;
    const uint32_t w_tile16 = (uint32_t)WT_W_TILE >> 4, wn_tile16 = (uint32_t)WT_WN_TILE >> 4;
    const uint32_t w_tap16 = w_tile16 * (uint32_t)nfull;
    const uint32_t tap_step16 = (uint32_t)p.dil * 8u;                    // dil rows x 128 B
    const uint32_t tapn_step16 = (uint32_t)p.dil * 2u;                   // dil rows x 32 B
    const uint32_t wt_base = tmem_base + (uint32_t)WT_TS_WCOL;           // TS: weights, lane 0
    const uint32_t wt_tap = (uint32_t)(w.cinp >> 1);                     // TMEM columns per tap
    int it = 0, itn = 0, i = 0;
    for (int j = slot0; j < w.n_ttiles; j += w.ctas_per_m, ++i) {
      const int acc = i & 1;
      const uint32_t d_addr = tmem_base + (uint32_t)(acc * BN);
      mbar_wait(&tmem_empty[acc], ((uint32_t)(i >> 1) & 1u) ^ 1u);
      tc_fence_after();
      for (int ck = 0; ck < nfull; ++ck, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const int nk = (!w.narrow && ck == nfull - 1) ? w.last_nk16 : (TC_BK / 16);
          uint32_t x_lo = x_lo0 + (uint32_t)s * a_stage16;
          if (TS) {
            uint32_t wa = wt_base + (uint32_t)(ck * (TC_BK / 2));          // 32 columns per 64-channel chunk
            for (int tap = 0; tap < p.K; ++tap) {
              umma_bf16_ts(d_addr, wa, x_lo, hi128, idesc, (ck > 0 || tap > 0) ? 1u : 0u);
              if (nk > 1) umma_bf16_ts(d_addr, wa + 8, x_lo + 2, hi128, idesc, 1u);
              if (nk > 2) umma_bf16_ts(d_addr, wa + 16, x_lo + 4, hi128, idesc, 1u);
              if (nk > 3) umma_bf16_ts(d_addr, wa + 24, x_lo + 6, hi128, idesc, 1u);
              x_lo += tap_step16;
              wa += wt_tap;
            }
          } else {
            uint32_t w_lo = w_lo0 + (uint32_t)ck * w_tile16;
            for (int tap = 0; tap < p.K; ++tap) {
              umma_bf16_lohi(d_addr, w_lo, hi128, x_lo, hi128, idesc, (ck > 0 || tap > 0) ? 1u : 0u);
              if (nk > 1) umma_bf16_lohi(d_addr, w_lo + 2, hi128, x_lo + 2, hi128, idesc, 1u);
              if (nk > 2) umma_bf16_lohi(d_addr, w_lo + 4, hi128, x_lo + 4, hi128, idesc, 1u);
              if (nk > 3) umma_bf16_lohi(d_addr, w_lo + 6, hi128, x_lo + 6, hi128, idesc, 1u);
              x_lo += tap_step16;
              w_lo += w_tap16;
            }
          }
          umma_commit(&empty_bar[s]);
          if (!w.narrow && ck == nfull - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
      }
      if (w.narrow) {
        const int s = itn % w.n_stages;
        const uint32_t ph = (uint32_t)(itn / w.n_stages) & 1u;
        mbar_wait(&fulln_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          uint32_t x_lo = xn_lo0 + (uint32_t)s * an_stage16;
          if (TS) {
            uint32_t wa = wt_base + (uint32_t)(nfull * (TC_BK / 2));
            for (int tap = 0; tap < p.K; ++tap) {
              umma_bf16_ts(d_addr, wa, x_lo, hi32, idesc, 1u);
              x_lo += tapn_step16;
              wa += wt_tap;
            }
          } else {
            uint32_t w_lo = wn_lo0;
            for (int tap = 0; tap < p.K; ++tap) {
              umma_bf16_lohi(d_addr, w_lo, hi32, x_lo, hi32, idesc, 1u);
              x_lo += tapn_step16;
              w_lo += wn_tile16;
            }
          }
          umma_commit(&emptyn_bar[s]);
          umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        ++itn;
      }
    }
  } else {
    // epilogue: 16 warps; TMEM lane quadrant = warp % 4 (32 output channels), column quarter = (warp - 2) / 4
    constexpr int CQ = BN / 4;                                // time steps per epilogue warp
    const int q = warp & 3;
    const int cq = (warp - 2) >> 2;
    const int r = r0 + q * 32 + lane;                         // stacked row of this thread
    const bool r_ok = r < w.rows_total;
    const int blk = r / w.rows_per_block, ch = r - blk * w.rows_per_block;
    const float bias_v = (r_ok && p.bias) ? __ldg(p.bias + r) : 0.f;
    const int col = p.out_ch_off + blk * p.out_ch_stride + ch;
    const int npad = (r_ok && ch == w.rows_per_block - 1) ? p.out_ch_stride - w.rows_per_block : 0;   // zero columns after a block
    const float sl = p.out_slope;
    const long long row_pitch = p.cp_out;
    // vec epilogue: the warp's 32 channels x 16 steps go through a 1 KB shared-memory tile so that the global stores are
    // 16 bytes per lane (8 consecutive channels of one time step) instead of 2: lane -> (time row lane/4 (+8), chunk lane%4)
    __nv_bfloat16* stg = reinterpret_cast<__nv_bfloat16*>(stage_out + (size_t)(warp - 2) * WT_STG_BYTES);
    const int k8 = lane & 3;
    const int rk = r0 + q * 32 + 8 * k8;                      // first stacked row of this lane's 8-channel chunk
    const bool rk_ok = rk < w.rows_total;
    const int blk_k = rk / w.rows_per_block, ch_k = rk - blk_k * w.rows_per_block;
    const int col_k = p.out_ch_off + blk_k * p.out_ch_stride + ch_k;
    const int npad_k = (rk_ok && ch_k + 8 == w.rows_per_block) ? p.out_ch_stride - w.rows_per_block : 0;
    int i = 0;
    for (int j = slot0; j < w.n_ttiles; j += w.ctas_per_m, ++i) {
      const int b = j / w.ttiles_per_b, t0 = (j - b * w.ttiles_per_b) * BN;
      const int acc = i & 1;
      mbar_wait(&tmem_full[acc], (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      if (!(p.debug & 32)) {
        const uint32_t taddr = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * CQ);
        const int t_left = p.Tout - (t0 + cq * CQ);           // valid time steps from this warp's first column
        if (w.vec) {
          // Staging tile [8 step pairs][32 channels] of 32-bit words: a lane stores its channel's (t, t+1) pair as one
          // word (conflict-free, 8 stores per 16 steps); lane (m = lane / 4, k = lane % 4) then reads the 8 words of
          // pair m, channels 8k..8k+7 and splits them into the two output rows t = 2m, 2m + 1 (16 bytes each).
          // Odd m read their two 16-byte halves in the opposite order, which keeps the quarter-warp phases on 32 banks.
          uint32_t* stg32 = reinterpret_cast<uint32_t*>(stg);
          const int m = lane >> 2;
          const int first = (m & 1) ? 4 : 0;
          // TDVC_TC_DEBUG & 128 (development): all tiles write the same few rows -- same store instructions, no DRAM stream
          const long long row0 = (p.debug & 128) ? (long long)(((j / w.ctas_per_m) & 1) * BN + cq * CQ + 2 * m)
                                                 : ((long long)b * p.tp_out + t0 + cq * CQ + p.out_halo + 2 * m);
          __nv_bfloat16* op = p.yp + row0 * row_pitch + col_k;
#pragma unroll 1
          for (int cc = 0; cc < NG; ++cc) {
            float v[16];
            tmem_ld16(taddr + (uint32_t)(cc * 16), v);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float o0 = v[2 * e] + bias_v, o1 = v[2 * e + 1] + bias_v;
              if (ACT == TDVC_ACT_LRELU) { o0 = fmaxf(o0, o0 * sl); o1 = fmaxf(o1, o1 * sl); }
              const __nv_bfloat162 pr = __floats2bfloat162_rn(o0, o1);
              stg32[e * 32 + lane] = *reinterpret_cast<const uint32_t*>(&pr);
            }
            __syncwarp();
            const uint4 qa = *reinterpret_cast<const uint4*>(stg32 + m * 32 + 8 * k8 + first);
            const uint4 qb = *reinterpret_cast<const uint4*>(stg32 + m * 32 + 8 * k8 + (4 - first));
            const uint4 lo4 = first ? qb : qa, hi4 = first ? qa : qb;      // words of channels 8k..8k+3 / 8k+4..8k+7
            const uint4 r0 = make_uint4(__byte_perm(lo4.x, lo4.y, 0x5410), __byte_perm(lo4.z, lo4.w, 0x5410),
                                        __byte_perm(hi4.x, hi4.y, 0x5410), __byte_perm(hi4.z, hi4.w, 0x5410));   // step 2m
            const uint4 r1 = make_uint4(__byte_perm(lo4.x, lo4.y, 0x7632), __byte_perm(lo4.z, lo4.w, 0x7632),
                                        __byte_perm(hi4.x, hi4.y, 0x7632), __byte_perm(hi4.z, hi4.w, 0x7632));   // step 2m + 1
            if (rk_ok && !(p.debug & 1)) {
              __nv_bfloat16* o8 = op + (long long)(cc * 16) * row_pitch;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (cc * 16 + 2 * m + h < t_left) {
                  *reinterpret_cast<uint4*>(o8) = h ? r1 : r0;
                  if (npad_k) {
                    if (npad_k == 8) *reinterpret_cast<uint4*>(o8 + 8) = make_uint4(0u, 0u, 0u, 0u);
                    else for (int z = 0; z < npad_k; ++z) o8[8 + z] = __float2bfloat16(0.f);
                  }
                }
                o8 += row_pitch;
              }
            }
            __syncwarp();
          }
        } else {
          __nv_bfloat16* op = p.yp + ((long long)b * p.tp_out + t0 + cq * CQ + p.out_halo) * row_pitch + col;
#pragma unroll 1
          for (int cc = 0; cc < NG; ++cc) {
            float v[16];
            tmem_ld16(taddr + (uint32_t)(cc * 16), v);
            if (r_ok && !(p.debug & 1)) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                if (cc * 16 + e < t_left) {
                  float o = v[e] + bias_v;
                  if (ACT == TDVC_ACT_LRELU) o = fmaxf(o, o * sl);
                  op[0] = __float2bfloat16(o);
                  for (int z = 1; z <= npad; ++z) op[z] = __float2bfloat16(0.f);
                }
                op += row_pitch;
              }
            } else {
              op += 16 * row_pitch;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

// ------------------------------------------------------------------------------------------ wgrad
// dW[co, ci, tap] = sum_{b,t} xa[b, t + tap*dil + t_off, ci] * dy[b, t, co]  as a GEMM with M = ci, N = co, K = time.
// Both operands are the channels-last bf16 copies that forward / dgrad already use, read as MN-major
// (channel-contiguous) SWIZZLE_128B tiles: box {64 ch, 64 t}.  One TMEM accumulator per tap (KT taps x NT
// columns <= 512).  A CTA walks a strided subset of the (batch, 64-step time chunk) units accumulating in TMEM,
// then adds its partial tile into ws[tap][co][ci] with fp32 reductions: a TMEM lane is a ci, so each warp-level
// RED covers 32 consecutive floats (one line).  wgrad_finalize_k transposes ws into the [Cout][Cin][K] layout.
struct WgTcP {
  int B, Tout, Cout, Cin, K, dil, t_off;
  int KT, NT, nb, ntap_groups, n_ntiles, stages, tmem_cols, nchunk_t, units, splits;
  int Mp, Np;     // padded ci / co extents of the workspace
  int x_ch_off, dy_ch_off;   // first channel of this conv's slice inside the packed operands
  float* ws;
  int bias;                  // also accumulate the bias gradient (column sums of dL/dy) into ws[K*Np*Mp + co]
};

constexpr int WG_BOX_BYTES = 64 * 64 * 2;   // 64 time rows x 64 channels bf16

__global__ void __launch_bounds__(TC_THREADS) conv_tc_wgrad_k(const __grid_constant__ CUtensorMap map_x,
                                                              const __grid_constant__ CUtensorMap map_dy, WgTcP p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int a_tap_bytes = 2 * WG_BOX_BYTES;           // 128 ci x 64 t
  const int b_bytes = p.nb * WG_BOX_BYTES;            // NT co x 64 t
  const int stage_bytes = p.KT * a_tap_bytes + b_bytes;
  // bias gradient = sum over (b, t) of dL/dy: one more "tap" whose A operand is a constant tile of ones, so the column
  // sums come out of the same GEMM (every TMEM lane of accumulator KT holds them); only the CTAs of the first
  // (ci tile, tap group) do it.  Replaces the atomics the dL/dy pack pass used to spend on it.
  uint8_t* ones = smem + (size_t)p.stages * stage_bytes;                    // 128 x 64 bf16 (1024-B aligned), when p.bias
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ones + (p.bias ? a_tap_bytes : 0));
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.y enumerates (m tile, tap group, n tile)
  int yy = blockIdx.y;
  const int nt_i = yy % p.n_ntiles; yy /= p.n_ntiles;
  const int tg = yy % p.ntap_groups;
  const int mt = yy / p.ntap_groups;
  const int ci0 = mt * 128, n0 = nt_i * p.NT, tap0 = tg * p.KT;
  const int ntaps = min(p.KT, p.K - tap0);
  const int split = blockIdx.x;
  const int my_units = (p.units - split + p.splits - 1) / p.splits;   // units split, split+splits, ...
  const bool do_bias = p.bias && mt == 0 && tg == 0;
  if (do_bias) {
    uint32_t* o32 = reinterpret_cast<uint32_t*>(ones);
    for (int i = threadIdx.x; i < a_tap_bytes / 4; i += TC_THREADS) o32[i] = 0x3F803F80u;      // bf16 1.0 pairs
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA's reads
  }

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (my_units > 0) {
    if (warp == 0) {
      if (elect_one()) {
        for (int it = 0; it < my_units; ++it) {
          const int u = split + it * p.splits;
          const int b = u / p.nchunk_t, tc = (u - b * p.nchunk_t) * 64;
          const int s = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          mbar_expect_tx(&full_bar[s], (uint32_t)(ntaps * a_tap_bytes + b_bytes));
          for (int tp = 0; tp < ntaps; ++tp) {
            const int tx = tc + (tap0 + tp) * p.dil + p.t_off;
            tma_load_3d(sa + tp * a_tap_bytes, &map_x, &full_bar[s], p.x_ch_off + ci0, tx, b);
            tma_load_3d(sa + tp * a_tap_bytes + WG_BOX_BYTES, &map_x, &full_bar[s], p.x_ch_off + ci0 + 64, tx, b);
          }
          for (int j = 0; j < p.nb; ++j)
            tma_load_3d(sa + p.KT * a_tap_bytes + j * WG_BOX_BYTES, &map_dy, &full_bar[s], p.dy_ch_off + n0 + 64 * j, tc, b);
        }
      }
    } else if (warp == 1) {
      // D=f32, A=B=bf16, both MN-major (bits 15, 16), N = NT, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int it = 0; it < my_units; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t s_addr = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t b_addr = s_addr + p.KT * a_tap_bytes;
          for (int tp = 0; tp < ntaps; ++tp) {
            const uint32_t a_addr = s_addr + tp * a_tap_bytes;
            for (int k = 0; k < 4; ++k) {      // 64 time rows per stage = 4 x K16; 16 rows = 2048 B
              const uint64_t da = make_sw128_mnmajor_desc(a_addr + k * 2048, WG_BOX_BYTES);
              const uint64_t db = make_sw128_mnmajor_desc(b_addr + k * 2048, WG_BOX_BYTES);
              umma_bf16(tmem_base + (uint32_t)(tp * p.NT), da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
            }
          }
          if (do_bias) {
            const uint32_t o_addr = smem_u32(ones);
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = make_sw128_mnmajor_desc(o_addr + k * 2048, WG_BOX_BYTES);
              const uint64_t db = make_sw128_mnmajor_desc(b_addr + k * 2048, WG_BOX_BYTES);
              umma_bf16(tmem_base + (uint32_t)(p.KT * p.NT), da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[s]);
          if (it == my_units - 1) umma_commit(tmem_full_bar);
        }
        __syncwarp();
      }
    } else {
      const int q = warp & 3;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      const int ci = ci0 + q * 32 + lane;
      const bool row_ok = ci < p.Cin;
      const int nvalid = min(p.NT, p.Cout - n0);
      for (int tp = 0; tp < ntaps; ++tp) {
        float* wrow = p.ws + ((long long)(tap0 + tp) * p.Np + n0) * p.Mp + ci;
        for (int c0 = 0; c0 < nvalid; c0 += 16) {
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tp * p.NT + c0), v);
          if (row_ok) {
            const int nj = min(16, nvalid - c0);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < nj) atomicAdd(wrow + (long long)(c0 + j) * p.Mp, v[j]);
          }
        }
      }
      if (do_bias && q == 0) {
        // every lane of accumulator KT holds the column sums; lane j of the warp adds column c0 + j
        float* brow = p.ws + (long long)p.K * p.Np * p.Mp + n0;
        for (int c0 = 0; c0 < nvalid; c0 += 16) {
          float v[16];
          tmem_ld16(tmem_base + (uint32_t)(p.KT * p.NT + c0), v);
          float mine = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) mine = (lane == j) ? v[j] : mine;
          if (lane < min(16, nvalid - c0)) atomicAdd(brow + c0 + lane, mine);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// dw[co][ci][k] = ws[k][co][ci]
__global__ void wgrad_finalize_k(const float* __restrict__ ws, float* __restrict__ dw, int Cout, int Cin, int K, int Np,
                                 int Mp, float* __restrict__ db) {
  pdl_prologue();
  long long n = (long long)Cout * Cin * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i % K);
    long long r = i / K;
    int ci = (int)(r % Cin), co = (int)(r / Cin);
    dw[i] = ws[((long long)k * Np + co) * Mp + ci];
  }
  if (db && blockIdx.x == 0)
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) db[i] = ws[(long long)K * Np * Mp + i];
}

// same, and leaves the workspace zeroed again (only the entries read here were ever written): a persistent workspace
// then needs no memset node per call
__global__ void wgrad_finalize_zero_k(float* __restrict__ ws, float* __restrict__ dw, int Cout, int Cin, int K, int Np, int Mp,
                                      float* __restrict__ db, int has_bias_row) {
  pdl_prologue();
  long long n = (long long)Cout * Cin * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i % K);
    long long r = i / K;
    int ci = (int)(r % Cin), co = (int)(r / Cin);
    float* src = ws + ((long long)k * Np + co) * Mp + ci;
    dw[i] = *src;
    *src = 0.f;
  }
  if (has_bias_row && blockIdx.x == 0)
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) {
      float* src = ws + (long long)K * Np * Mp + i;
      if (db) db[i] = *src;
      *src = 0.f;
    }
}

// ------------------------------------------------------------------------------------------ pack kernels
// x[B,C,T] fp32 NCW -> xp[B,Tp,Cp] bf16 channels-last, LeakyReLU(in_slope), halo rows reflect- or zero-filled.
// x[B,C,T] fp32 NCW -> xp[B,Tp,Cp] bf16 channels-last (channels [c_off, c_off+Cw)), LeakyReLU(slope), reflect / zero
// halo rows.  A CTA moves a CH-channel x TL-step tile (CH*TL = 4096 elements whatever the channel count, so thin
// tensors still put 16 independent loads per thread in flight): reads are 128-byte lines along time, writes are
// bf16x2 per lane = 128-byte lines along channels, zero filled up to the 64-channel slice width.
template <int CH>
__global__ void __launch_bounds__(256) pack_cl_bf16_k(const float* __restrict__ x, __nv_bfloat16* __restrict__ xp, int C,
                                                     int T, int Cp, int Tp, int halo, int pad_mode, float slope,
                                                     float* __restrict__ chan_sum, int c_off, int Cw, int ones_ch,
                                                     const float* __restrict__ film_gb, const float* __restrict__ mask_y,
                                                     float mask_slope) {
  pdl_prologue();
  constexpr int TL = 4096 / CH;             // time steps per tile: 64 / 128 / 256
  constexpr int RPW = CH / 8;               // channel rows per warp
  constexpr int LPR = TL / 32;              // loads per row per lane
  __shared__ float tile[CH][TL + 1];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 64;           // first channel of this CTA's 64-wide output slice ...
  const int cin0 = c0 + (CH < 64 ? 0 : 0);  // ... and of the CH source channels it reads (CH < 64 only when C <= CH)
  const int tp0 = blockIdx.x * TL;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  float vals[RPW][LPR];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int cy = wrp + 8 * r;
    const int c = cin0 + cy;
    const float* row = x + ((long long)b * C + c) * T;
#pragma unroll
    for (int l = 0; l < LPR; ++l) {
      const int tp = tp0 + lane + 32 * l;
      float v = 0.f;
      if (c < C && tp < Tp) {
        int u = tp - halo;
        bool ok = true;
        if (u < 0) { if (pad_mode == TDVC_PAD_REFLECT) { u = -u; ok = u < T; } else ok = false; }
        else if (u >= T) { if (pad_mode == TDVC_PAD_REFLECT) { u = 2 * (T - 1) - u; ok = u >= 0; } else ok = false; }
        if (ok) {
          v = __ldg(row + u);
          if (film_gb) {       // FiLM before the activation: h*(1+gamma)+beta, gb = [B, 2C, T]
            const float* gp = film_gb + ((long long)b * 2 * C + c) * T + u;
            v = fmaf(v, 1.f + __ldg(gp), __ldg(gp + (long long)C * T));
          }
          // x = dL/dy of a LeakyReLU layer, mask_y its output: the derivative of the activation applied on the way in
          if (mask_y && !(__ldg(mask_y + ((long long)b * C + c) * T + u) > 0.f)) v *= mask_slope;
        }
      }
      vals[r][l] = v;
    }
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int cy = wrp + 8 * r;
    float sum = 0.f;
#pragma unroll
    for (int l = 0; l < LPR; ++l) {
      const int tp = tp0 + lane + 32 * l;
      float v = vals[r][l];
      if (tp >= halo && tp < halo + T) sum += v;          // each source sample counted once
      tile[cy][lane + 32 * l] = v > 0.f ? v : v * slope;
    }
    if (chan_sum) {      // per-channel sum of the fp32 source (the bias gradient when x is dL/dy)
      sum = warp_sum(sum);
      if (lane == 0 && cin0 + cy < C) atomicAdd(chan_sum + cin0 + cy, sum);
    }
  }
  __syncthreads();
  const int c = c0 + 2 * lane;                // this lane's channel pair inside the slice
  if (c >= Cw) return;
  for (int ty = wrp; ty < TL; ty += 8) {
    const int tp = tp0 + ty;
    if (tp >= Tp) break;
    float o0 = (2 * lane < CH) ? tile[2 * lane < CH ? 2 * lane : 0][ty] : 0.f;
    float o1 = (2 * lane + 1 < CH) ? tile[2 * lane + 1 < CH ? 2 * lane + 1 : 0][ty] : 0.f;
    const bool valid_row = tp >= halo && tp < halo + T;
    if (c == ones_ch) o0 = valid_row ? 1.f : 0.f;         // constant-one channel: bias grad via the wgrad GEMM
    if (c + 1 == ones_ch) o1 = valid_row ? 1.f : 0.f;
    __nv_bfloat16* dst = xp + ((long long)b * Tp + tp) * Cp + c_off + c;
    if (c + 1 < Cw) *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(o0, o1);
    else *dst = __float2bfloat16(o0);
  }
}


// The same pack for the common un-haloed case (dL/dy of the discriminator layers, the dense convs' inputs): T % 4 == 0, no halo,
// no FiLM, no ones channel.  16-byte loads along time (a 64-step row = 16 lanes), the transposed bf16x2 stores of the kernel
// above.  The scalar kernel ran these at ~0.7 TB/s (ncu: 26 us for 18 MB in + 9 MB out).
template <int CH>
__global__ void __launch_bounds__(256) pack_cl_bf16_v4_k(const float* __restrict__ x, __nv_bfloat16* __restrict__ xp, int C, int T,
                                                        int Cp, float slope, float* __restrict__ chan_sum, int c_off, int Cw,
                                                        const float* __restrict__ mask_y, float mask_slope, int halo,
                                                        int pad_mode) {
  pdl_prologue();
  constexpr int TL = 4096 / CH;              // time steps per tile: 64 / 128 / 256 (CH * TL = 4096 elements, as above)
  constexpr int LPR = TL / 4;                // lanes per channel row: 16 / 32 / 64
  constexpr int RPP = 256 / LPR;             // rows per pass: 16 / 8 / 4 (always 4 passes)
  __shared__ float tile[CH][TL + 1];
  const int b = blockIdx.z, c0 = blockIdx.y * 64, t0 = blockIdx.x * TL;
  const int col4 = threadIdx.x % LPR, r0 = threadIdx.x / LPR;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  float4 v[4], m[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + r0 + RPP * i, t = t0 + 4 * col4;
    const bool ok = c < C && t < T;
    const long long off = ((long long)b * C + c) * T + t;
    v[i] = ok ? __ldg(reinterpret_cast<const float4*>(x + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (mask_y) m[i] = ok ? __ldg(reinterpret_cast<const float4*>(mask_y + off)) : make_float4(1.f, 1.f, 1.f, 1.f);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 a = v[i];
    if (mask_y) {
      if (!(m[i].x > 0.f)) a.x *= mask_slope;
      if (!(m[i].y > 0.f)) a.y *= mask_slope;
      if (!(m[i].z > 0.f)) a.z *= mask_slope;
      if (!(m[i].w > 0.f)) a.w *= mask_slope;
    }
    if (chan_sum) {      // per-channel sum of the (masked) source: the bias gradient
      float sum = (a.x + a.y) + (a.z + a.w);
#pragma unroll
      for (int o = (LPR < 32 ? LPR : 32) / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if ((col4 & ((LPR < 32 ? LPR : 32) - 1)) == 0 && c0 + r0 + RPP * i < C) atomicAdd(chan_sum + c0 + r0 + RPP * i, sum);
    }
    float* tr = &tile[r0 + RPP * i][4 * col4];
    tr[0] = a.x > 0.f ? a.x : a.x * slope;
    tr[1] = a.y > 0.f ? a.y : a.y * slope;
    tr[2] = a.z > 0.f ? a.z : a.z * slope;
    tr[3] = a.w > 0.f ? a.w : a.w * slope;
  }
  __syncthreads();
  const int c = c0 + 2 * lane;
  if (c >= Cw) return;
  // tiles are cut in SOURCE time steps (so the 16-byte loads stay aligned whatever the halo); sample t is row halo + t, and a
  // sample that a reflect halo row mirrors (1 <= t <= halo on the left, T-1-halo <= t <= T-2 on the right) is written there too
  const int Tp = T + 2 * halo;
  const bool reflect = pad_mode == TDVC_PAD_REFLECT && halo > 0;
  for (int ty = wrp; ty < TL; ty += 8) {
    const int t = t0 + ty;
    if (t >= T) break;
    const float o0 = (2 * lane < CH) ? tile[2 * lane < CH ? 2 * lane : 0][ty] : 0.f;
    const float o1 = (2 * lane + 1 < CH) ? tile[2 * lane + 1 < CH ? 2 * lane + 1 : 0][ty] : 0.f;
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(o0, o1);
    int rows[3] = {halo + t, -1, -1};
    if (reflect) {
      if (t >= 1 && t <= halo) rows[1] = halo - t;
      if (t >= T - 1 - halo && t <= T - 2) rows[2] = halo + 2 * T - 2 - t;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      if (rows[r] < 0) continue;
      __nv_bfloat16* dst = xp + ((long long)b * Tp + rows[r]) * Cp + c_off + c;
      if (c + 1 < Cw) *reinterpret_cast<__nv_bfloat162*>(dst) = h2;
      else *dst = __low2bfloat16(h2);
    }
  }
  if (halo > 0 && !reflect) {
    // zero halo rows: the first tile writes the left ones, the last tile the right ones
    const __nv_bfloat162 z2 = __floats2bfloat162_rn(0.f, 0.f);
    const bool first = blockIdx.x == 0, last = blockIdx.x == gridDim.x - 1;
    for (int j = wrp; j < halo; j += 8) {
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        if (side == 0 ? !first : !last) continue;
        __nv_bfloat16* dst = xp + ((long long)b * Tp + (side == 0 ? j : halo + T + j)) * Cp + c_off + c;
        if (c + 1 < Cw) *reinterpret_cast<__nv_bfloat162*>(dst) = z2;
        else *dst = __low2bfloat16(z2);
      }
    }
  }
}

// The pack for short sequences (T <= 64: the discriminators' 1024-channel tail at T = 9 .. 35, the T/320 stage): the
// 64-channel x T block of one sample is CONTIGUOUS in the NCW source, so it is read as one linear run (full sectors whatever
// T % 4 is) and transposed through shared memory; halo rows (zero or reflect) and the LeakyReLU-backward mask as above.
// pack_cl_bf16_k reads such tensors as 64-step rows of which half the lanes are idle (10 us per 4.6 MB tensor, 76 launches).
__global__ void __launch_bounds__(256) pack_cl_bf16_short_k(const float* __restrict__ x, __nv_bfloat16* __restrict__ xp, int C, int T,
                                                           int Cp, float slope, float* __restrict__ chan_sum, int c_off, int Cw,
                                                           const float* __restrict__ mask_y, float mask_slope, int halo,
                                                           int pad_mode) {
  pdl_prologue();
  __shared__ float tile[64][65];
  const int b = blockIdx.y, c0 = blockIdx.x * 64;
  const int nch = min(64, C - c0);                       // source channels of this block (<= 0: padding channels only)
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const long long base = ((long long)b * C + c0) * T;
  const int n = max(nch, 0) * T;
  for (int i0 = threadIdx.x; i0 < n; i0 += 256 * 4) {
    float v[4], m[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + 256 * u;
      v[u] = i < n ? __ldg(x + base + i) : 0.f;
      m[u] = (mask_y && i < n) ? __ldg(mask_y + base + i) : 1.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + 256 * u;
      if (i < n) {
        const int c = i / T, t = i - c * T;
        tile[c][t] = (m[u] > 0.f) ? v[u] : v[u] * mask_slope;
      }
    }
  }
  __syncthreads();
  if (chan_sum && (int)threadIdx.x < nch) {              // bias gradient: per-channel sum of the (masked) source
    float sum = 0.f;
    for (int t = 0; t < T; ++t) sum += tile[threadIdx.x][t];
    atomicAdd(chan_sum + c0 + threadIdx.x, sum);
  }
  const int c = c0 + 2 * lane;
  if (c >= Cw) return;
  const int Tp = T + 2 * halo;
  for (int tp = wrp; tp < Tp; tp += 8) {
    int u = tp - halo;
    bool ok = true;
    if (u < 0) { if (pad_mode == TDVC_PAD_REFLECT) { u = -u; ok = u < T; } else ok = false; }
    else if (u >= T) { if (pad_mode == TDVC_PAD_REFLECT) { u = 2 * (T - 1) - u; ok = u >= 0; } else ok = false; }
    float o0 = 0.f, o1 = 0.f;
    if (ok) {
      if (2 * lane < nch) o0 = tile[2 * lane][u];
      if (2 * lane + 1 < nch) o1 = tile[2 * lane + 1][u];
      o0 = o0 > 0.f ? o0 : o0 * slope;
      o1 = o1 > 0.f ? o1 : o1 * slope;
    }
    __nv_bfloat16* dst = xp + ((long long)b * Tp + tp) * Cp + c_off + c;
    if (c + 1 < Cw) *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(o0, o1);
    else *dst = __float2bfloat16(o0);
  }
}

// The decoder's conditioning tensor cat([speaker code repeated over time, excitation pyramid level]) (model/generator.py:
// 387-399) written straight into the bf16 channels-last operand of the cond_var convs: cp[b, t, 0..Cc) = c[b, :] (constant
// over time), cp[b, t, Cc..Cc+Ce) = e[b, :, t], cp[b, t, Cc+Ce] = 1 (the bias-gradient channel), zero up to Cg.  The fp32
// [B, Cc+Ce, T] tensor the reference builds (156 MB at the full-rate stage) is never materialised.
__global__ void __launch_bounds__(256) cond_pack_cl_k(const float* __restrict__ c, const float* __restrict__ e,
                                                      __nv_bfloat16* __restrict__ cp, int Cc, int Ce, int T, int Cg) {
  pdl_prologue();
  extern __shared__ float sm[];               // [Cc] speaker code | [Ce][64 + 1] excitation tile
  float* cs = sm;
  float* es = sm + Cc;
  const int b = blockIdx.y, t0 = blockIdx.x * 64;
  for (int i = threadIdx.x; i < Cc; i += 256) cs[i] = __ldg(c + (long long)b * Cc + i);
  for (int i = threadIdx.x; i < Ce * 64; i += 256) {
    const int ch = i >> 6, t = i & 63;
    es[ch * 65 + t] = (t0 + t < T) ? __ldg(e + ((long long)b * Ce + ch) * T + t0 + t) : 0.f;
  }
  __syncthreads();
  const int pairs = Cg / 2;
  const int nt = min(64, T - t0);
  for (int i = threadIdx.x; i < nt * pairs; i += 256) {
    const int t = i / pairs, ch = 2 * (i - t * pairs);
    float v[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int cc = ch + q;
      v[q] = cc < Cc ? cs[cc] : (cc < Cc + Ce ? es[(cc - Cc) * 65 + t] : (cc == Cc + Ce ? 1.f : 0.f));
    }
    *reinterpret_cast<__nv_bfloat162*>(cp + ((long long)b * T + t0 + t) * Cg + ch) = __floats2bfloat162_rn(v[0], v[1]);
  }
}

__global__ void pack_weight_bf16_k(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int Cout, int Cin, int K,
                                   int Rp, int Qp, int transpose_flip, int R_total, int r_off, int Q_total, int q_off) {
  pdl_prologue();
  // writes the [K][Rp][Qp] block at (r_off, q_off) of wp[K][R_total][Q_total]
  // plain: R = co, Q = ci; transpose_flip: R = ci, Q = co, taps reversed
  long long n = (long long)K * Rp * Qp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int qq = (int)(i % Qp);
    long long r2 = i / Qp;
    int rr = (int)(r2 % Rp);
    int k = (int)(r2 / Rp);
    int co = transpose_flip ? qq : rr, ci = transpose_flip ? rr : qq;
    int ks = transpose_flip ? K - 1 - k : k;
    float v = (co < Cout && ci < Cin) ? w[((long long)co * Cin + ci) * K + ks] : 0.f;
    wp[((long long)k * R_total + r_off + rr) * Q_total + q_off + qq] = __float2bfloat16(v);
  }
}

// A list of small packs in ONE launch, the job table passed by value (the sources are this step's weight-norm outputs: a
// device-resident table would need an uncapturable host-to-device copy per step).  kind 0: the block of pack_weight_bf16_k
// (rows [r_off, r_off + Rp) x columns [q_off, q_off + Qp) of wp[K][R_total][Q_total]); kind 1: Rp floats of a bias vector,
// the first Cout from src, zeros after.  The conditioning path of an MRF stage packed its 2 x 9 weights and biases with 36 + 36
// launches per stage and direction (144 of the step's 161 pack_weight_bf16_k launches).
struct PackJobs { tdvc_pack_job j[TDVC_PACK_MAX_JOBS]; };

__global__ void __launch_bounds__(256) pack_jobs_k(const __grid_constant__ PackJobs J) {
  pdl_prologue();
  const tdvc_pack_job& e = J.j[blockIdx.y];
  if (e.kind == 1) {
    float* dst = reinterpret_cast<float*>(e.dst);
    const float* src = reinterpret_cast<const float*>(e.src);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < e.Rp; i += gridDim.x * blockDim.x)
      dst[i] = (src && i < e.Cout) ? __ldg(src + i) : 0.f;
    return;
  }
  const float* w = reinterpret_cast<const float*>(e.src);
  __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(e.dst);
  const int nrq = e.Rp * e.Qp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrq; i += gridDim.x * blockDim.x) {
    const int rr = i / e.Qp, qq = i - rr * e.Qp;
    const int co = e.flip ? qq : rr, ci = e.flip ? rr : qq;
    const bool ok = co < e.Cout && ci < e.Cin;
    const float* src = w + ((long long)co * e.Cin + ci) * e.K;
    for (int k = 0; k < e.K; ++k) {
      const float v = ok ? __ldg(src + (e.flip ? e.K - 1 - k : k)) : 0.f;
      wp[((long long)k * e.R_total + e.r_off + rr) * e.Q_total + e.q_off + qq] = __float2bfloat16(v);
    }
  }
}

// Many weights in one launch: job[j] = {src offset in flat_w (floats), dst offset in flat_wp (bf16), Cout, Cin, K, Rp, Qp,
// transpose_flip}; blockIdx.y = job, blockIdx.x strides over the job's K*Rp*Qp outputs (layout of pack_weight_bf16_k
// with R_total = Rp, Q_total = Qp).
__global__ void pack_weight_multi_k(const long long* __restrict__ jobs, const float* __restrict__ flat_w,
                                    __nv_bfloat16* __restrict__ flat_wp) {
  pdl_prologue();
  const long long* e = jobs + 8LL * blockIdx.y;
  const float* w = flat_w + e[0];
  __nv_bfloat16* wp = flat_wp + e[1];
  const int Cout = (int)e[2], Cin = (int)e[3], K = (int)e[4], Rp = (int)e[5], Qp = (int)e[6], flip = (int)e[7];
  // a thread owns one (row, column) of the operand and walks its K taps: the K source floats are contiguous (neighbouring
  // threads read neighbouring runs), each tap's store is coalesced, and the index arithmetic is 32-bit and once per K outputs
  // (one output per thread with 64-bit div/mod ran the discriminator's 36 M outputs at 0.95 TB/s)
  const int nrq = Rp * Qp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrq; i += gridDim.x * blockDim.x) {
    const int rr = i / Qp, qq = i - rr * Qp;
    const int co = flip ? qq : rr, ci = flip ? rr : qq;
    const bool ok = co < Cout && ci < Cin;
    const float* src = w + ((long long)co * Cin + ci) * K;
    for (int k = 0; k < K; ++k) {
      const float v = ok ? __ldg(src + (flip ? K - 1 - k : k)) : 0.f;
      wp[(long long)k * nrq + i] = __float2bfloat16(v);
    }
  }
}

}  // namespace tdvc
using namespace tdvc;

static int pack_cl_bf16_launch(const float* x, void* xp, int B, int C, int T, int Cp, int halo, int pad_mode, float in_slope,
                               float* chan_sum, int c_off, int Cw, int ones_ch, const float* film_gb, const float* mask_y,
                               float mask_slope, void* stream) {
  if (Cw <= 0) Cw = Cp - c_off;
  TDVC_CHECK_ARG(B >= 0 && C > 0 && T > 0 && Cw >= C && c_off >= 0 && c_off + Cw <= Cp && Cp % 8 == 0 && halo >= 0 && x && xp);
  TDVC_CHECK_ARG(ones_ch < 0 || (ones_ch >= C && ones_ch < Cw));
  if (pad_mode == TDVC_PAD_REFLECT) TDVC_CHECK_ARG(halo < T);
  if (chan_sum) TDVC_CUDA(cudaMemsetAsync(chan_sum, 0, sizeof(float) * C, (cudaStream_t)stream));
  if (B == 0) return TDVC_OK;
  int Tp = T + 2 * halo;
  TDVC_CHECK_ARG(Cw % 2 == 0 && c_off % 2 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* o = (__nv_bfloat16*)xp;
  // thin tensors: fewer channel rows, longer time tiles (same bytes in flight per CTA)
  const bool v4 = T % 4 == 0 && !film_gb && ones_ch < 0 && (uintptr_t)x % 16 == 0 && (!mask_y || (uintptr_t)mask_y % 16 == 0);
  if (T <= 64 && !film_gb && ones_ch < 0 && B <= 65535) {
    tdvc::launch_k(pack_cl_bf16_short_k, dim3(cdiv(Cw, 64), B), 256, 0, st, x, o, C, T, Cp, in_slope, chan_sum, c_off, Cw, mask_y,
                   mask_slope, halo, pad_mode);
  } else if (C <= 16 && Cw <= 64) {
    dim3 grid(cdiv(v4 ? T : Tp, 256), 1, B);
    if (v4) tdvc::launch_k(pack_cl_bf16_v4_k<16>, grid, 256, 0, st, x, o, C, T, Cp, in_slope, chan_sum, c_off, Cw, mask_y, mask_slope, halo,
                           pad_mode);
    else tdvc::launch_k(pack_cl_bf16_k<16>, grid, 256, 0, st, x, o, C, T, Cp, Tp, halo, pad_mode, in_slope, chan_sum, c_off, Cw, ones_ch,
                                             film_gb, mask_y, mask_slope);
  } else if (C <= 32 && Cw <= 64) {
    dim3 grid(cdiv(v4 ? T : Tp, 128), 1, B);
    if (v4) tdvc::launch_k(pack_cl_bf16_v4_k<32>, grid, 256, 0, st, x, o, C, T, Cp, in_slope, chan_sum, c_off, Cw, mask_y, mask_slope, halo,
                           pad_mode);
    else tdvc::launch_k(pack_cl_bf16_k<32>, grid, 256, 0, st, x, o, C, T, Cp, Tp, halo, pad_mode, in_slope, chan_sum, c_off, Cw, ones_ch,
                                             film_gb, mask_y, mask_slope);
  } else {
    dim3 grid(cdiv(v4 ? T : Tp, 64), cdiv(Cw, 64), B);
    TDVC_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535);
    if (v4) tdvc::launch_k(pack_cl_bf16_v4_k<64>, grid, 256, 0, st, x, o, C, T, Cp, in_slope, chan_sum, c_off, Cw, mask_y, mask_slope, halo,
                           pad_mode);
    else tdvc::launch_k(pack_cl_bf16_k<64>, grid, 256, 0, st, x, o, C, T, Cp, Tp, halo, pad_mode, in_slope, chan_sum, c_off, Cw, ones_ch,
                                             film_gb, mask_y, mask_slope);
  }
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_pack_cl_bf16(const float* x, void* xp, int B, int C, int T, int Cp, int halo, int pad_mode,
                                 float in_slope, float* chan_sum, int c_off, int Cw, int ones_ch, const float* film_gb,
                                 void* stream) {
  return pack_cl_bf16_launch(x, xp, B, C, T, Cp, halo, pad_mode, in_slope, chan_sum, c_off, Cw, ones_ch, film_gb, nullptr, 1.f,
                             stream);
}

extern "C" int tdvc_pack_cl_bf16_masked(const float* dy, const float* y, float slope, void* dyp, int B, int C, int T, int Cp,
                                        int halo, float* chan_sum, void* stream) {
  TDVC_CHECK_ARG(y != nullptr);
  return pack_cl_bf16_launch(dy, dyp, B, C, T, Cp, halo, TDVC_PAD_ZEROS, 1.f, chan_sum, 0, 0, -1, nullptr, y, slope, stream);
}

extern "C" int tdvc_cond_pack_cl(const float* c, const float* e, void* cp, int B, int Cc, int Ce, int T, int Cg, void* stream) {
  TDVC_CHECK_ARG(c && e && cp && B >= 0 && Cc > 0 && Ce > 0 && T > 0 && Cg % 2 == 0 && Cg > Cc + Ce && B <= 65535);
  if (B == 0) return TDVC_OK;
  const size_t smem = ((size_t)Cc + (size_t)Ce * 65) * sizeof(float);
  TDVC_CHECK_ARG(smem <= 48 * 1024);
  tdvc::launch_k(cond_pack_cl_k, dim3(cdiv(T, 64), B), 256, smem, (cudaStream_t)stream, c, e, (__nv_bfloat16*)cp, Cc, Ce, T, Cg);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_pack_weight_bf16(const float* w, void* wp, int Cout, int Cin, int K, int Coutp, int Cinp,
                                     int transpose_flip, int R_total, int r_off, int Q_total, int q_off, void* stream) {
  TDVC_CHECK_ARG(Cout > 0 && Cin > 0 && K > 0 && Coutp >= Cout && Cinp >= Cin && w && wp);
  int Rp = transpose_flip ? Cinp : Coutp, Qp = transpose_flip ? Coutp : Cinp;
  if (R_total <= 0) { R_total = Rp; r_off = 0; }
  if (Q_total <= 0) { Q_total = Qp; q_off = 0; }
  TDVC_CHECK_ARG(r_off >= 0 && q_off >= 0 && r_off + Rp <= R_total && q_off + Qp <= Q_total);
  long long n = (long long)K * Rp * Qp;
  int blocks = (int)std::min<long long>((n + 255) / 256, 8LL * num_sms());
  tdvc::launch_k(pack_weight_bf16_k, blocks, 256, 0, (cudaStream_t)stream, w, (__nv_bfloat16*)wp, Cout, Cin, K, Rp, Qp, transpose_flip, R_total,
                                                                       r_off, Q_total, q_off);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_pack_jobs(const tdvc_pack_job* jobs, int n_jobs, void* stream) {
  TDVC_CHECK_ARG(jobs && n_jobs >= 0 && n_jobs <= TDVC_PACK_MAX_JOBS);
  if (n_jobs == 0) return TDVC_OK;
  PackJobs J{};
  long long max_n = 1;
  for (int i = 0; i < n_jobs; ++i) {
    const tdvc_pack_job& e = jobs[i];
    TDVC_CHECK_ARG(e.dst && (e.kind == 0 || e.kind == 1) && e.Rp >= 0);
    if (e.kind == 0) {
      TDVC_CHECK_ARG(e.src && e.Cout > 0 && e.Cin > 0 && e.K > 0 && e.Qp > 0 && e.r_off >= 0 && e.q_off >= 0 &&
                     e.r_off + e.Rp <= e.R_total && e.q_off + e.Qp <= e.Q_total);
      max_n = std::max(max_n, (long long)e.Rp * e.Qp);
    } else {
      max_n = std::max(max_n, (long long)e.Rp);
    }
    J.j[i] = e;
  }
  const int bx = (int)std::min<long long>(cdiv(max_n, 256), 64);
  tdvc::launch_k(pack_jobs_k, dim3(bx, n_jobs), 256, 0, (cudaStream_t)stream, J);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_pack_weight_bf16_multi(const void* jobs, int n_jobs, int blocks_per_job, const float* flat_w, void* flat_wp,
                                           void* stream) {
  TDVC_CHECK_ARG(jobs && n_jobs > 0 && n_jobs <= 65535 && blocks_per_job > 0 && flat_w && flat_wp);
  tdvc::launch_k(pack_weight_multi_k, dim3(blocks_per_job, n_jobs), 256, 0, (cudaStream_t)stream, (const long long*)jobs, flat_w,
                 (__nv_bfloat16*)flat_wp);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

// taps summed over the groups of a launch
static double tc_taps(const tdvc_tc_conv* c) {
  double n = 0;
  for (int g = 0; g < c->groups; ++g) n += (c->groups <= 4 && c->kg[g] > 0) ? c->kg[g] : c->K;
  return n;
}

// chain_mode 6 follow-up: fold the reflect-halo contributions onto the samples they mirror.  One thread per TARGET sample
// (so no two threads touch the same element, also when the signal is so short that the left and the right fold overlap):
// slot s < halo is sample tt = s + 1 (mirrored by padded row halo - tt and, when the signal is short, also by a right-halo
// row), slot s >= halo is sample tt = T - 1 - halo + (s - halo) unless a left slot already owns it.
__global__ void chain_fold_k(const float* __restrict__ hb, float* __restrict__ y, __nv_bfloat16* __restrict__ yp, int groups,
                             int B, int C, int T, int halo, long long y_grp_stride, long long y_b_stride, int cp_out,
                             int out_ch_off, int out_ch_stride) {
  tdvc::pdl_prologue();
  const long long n = (long long)groups * B * C * 2 * halo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(i % (2 * halo));
    const long long r0 = i / (2 * halo);
    const int c = (int)(r0 % C);
    const long long r = r0 / C;
    const int b = (int)(r % B), g = (int)(r / B);
    int tt;
    if (s < halo) {
      tt = s + 1;
    } else {
      tt = T - 1 - halo + (s - halo);
      if (tt <= halo) continue;                      // owned by a left slot
    }
    const float* h = hb + r0 * (2 * halo);
    float add = 0.f;
    if (tt >= 1 && tt <= halo) add += h[halo - tt];                            // padded row halo - tt mirrors sample tt
    const int ir = T - 2 - tt;                                                // padded row halo + T + ir mirrors T - 2 - ir
    if (ir >= 0 && ir < halo) add += h[halo + ir];
    float* row = y + (long long)g * y_grp_stride + (long long)b * y_b_stride + (long long)c * T;
    const float v = row[tt] + add;
    row[tt] = v;
    if (yp) yp[((long long)b * T + tt) * cp_out + out_ch_off + g * out_ch_stride + c] = __float2bfloat16(v);
  }
}

extern "C" int tdvc_chain_fold(const float* halo_buf, float* y, void* yp, int groups, int B, int C, int T, int halo,
                               int64_t y_grp_stride, int64_t y_b_stride, int cp_out, int out_ch_off, int out_ch_stride,
                               void* stream) {
  TDVC_CHECK_ARG(halo_buf && y && groups > 0 && B >= 0 && C > 0 && T > 0 && halo >= 0 && halo < T);
  if (B == 0 || halo == 0) return TDVC_OK;
  const long long n = (long long)groups * B * C * 2 * halo;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 8LL * tdvc::num_sms());
  tdvc::launch_k(chain_fold_k, blocks, 256, 0, (cudaStream_t)stream, halo_buf, y, (__nv_bfloat16*)yp, groups, B, C, T, halo,
                 (long long)y_grp_stride, (long long)y_b_stride, cp_out, out_ch_off, out_ch_stride);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_conv1d_tc_fwd_ex(const tdvc_tc_conv* c, void* stream) {
  TDVC_CHECK_ARG(c && c->xp && c->wp && c->B >= 0 && c->Tp > 0 && c->Tout > 0 && c->K > 0 && c->dilation > 0);
  TDVC_CHECK_ARG(c->groups >= 1 && c->Cp_total % 8 == 0 && c->Cinp_g % 8 == 0 && c->Cinp_g > 0);
  TDVC_CHECK_ARG(c->Cout_g > 0 && c->Coutp_g >= c->Cout_g && c->Coutp_g % 16 == 0);
  TDVC_CHECK_ARG(((uintptr_t)c->xp % 16 == 0) && ((uintptr_t)c->wp % 16 == 0));
  TDVC_CHECK_ARG(c->a_ch_off % 8 == 0 && c->a_ch_stride % 8 == 0);
  TDVC_CHECK_ARG(c->out_act >= 0 && c->out_act <= 2);
  if (c->out_packed) {
    TDVC_CHECK_ARG(c->yp && c->cp_out % 8 == 0 && c->out_ch_off % 16 == 0 && c->out_ch_stride % 16 == 0 &&
                   ((uintptr_t)c->yp % 16 == 0) && c->out_act != TDVC_ACT_TANH && (c->chain_mode || (!c->gb && !c->residual)));
    // every 16-channel chunk a thread stores must lie inside the row
    TDVC_CHECK_ARG(c->out_ch_off + (c->groups - 1) * c->out_ch_stride + c->Coutp_g <= c->cp_out);
  } else {
    TDVC_CHECK_ARG(c->y != nullptr);
    if (c->groups > 1 && !c->chain_mode) TDVC_CHECK_ARG(!c->gb && !c->residual);
  }
  const int chain = c->chain_mode;
  TDVC_CHECK_ARG(chain == 0 || (chain >= 3 && chain <= 6));
  if (c->maskp) {
    TDVC_CHECK_ARG(c->cm % 8 == 0 && c->mask_ch_off % 16 == 0 && c->mask_ch_stride % 16 == 0 && ((uintptr_t)c->maskp % 16 == 0));
    TDVC_CHECK_ARG(c->mask_ch_off + (c->groups - 1) * c->mask_ch_stride + c->Coutp_g <= c->cm);
    if (!chain) TDVC_CHECK_ARG(!c->gb && !c->residual && c->out_act == TDVC_ACT_NONE);
  }
  if (chain) {
    // a thread's 16 accumulator columns must all be real channels, rows of the packed tensors 16-byte aligned
    TDVC_CHECK_ARG(c->Cout_g % 16 == 0 && c->Coutp_g == c->Cout_g);
    if (chain == 3) TDVC_CHECK_ARG(c->out_packed && c->yp && (!c->yp2 || (uintptr_t)c->yp2 % 16 == 0));
    if (chain == 4) {
      TDVC_CHECK_ARG(!c->out_packed && c->y && c->residual);
      if (c->yp) TDVC_CHECK_ARG(c->cp_out % 8 == 0 && c->out_ch_off % 8 == 0 && c->out_ch_stride % 8 == 0 && (uintptr_t)c->yp % 16 == 0 &&
                                c->out_halo >= 0 && c->out_halo < c->Tout && c->tp_out >= c->Tout + 2 * c->out_halo);
    }
    if (chain == 5) {
      TDVC_CHECK_ARG(c->out_packed && c->yp && c->maskp);
      if (c->gb) TDVC_CHECK_ARG(c->auxp && c->dgbp && (uintptr_t)c->auxp % 16 == 0 && (uintptr_t)c->dgbp % 16 == 0 && c->dgb_cp % 8 == 0 &&
                                c->dgb_ch_off % 8 == 0 && c->dgb_ch_stride % 8 == 0 &&
                                c->dgb_ch_off + (c->groups - 1) * c->dgb_ch_stride + 2 * c->Cout_g <= c->dgb_cp);
    }
    if (chain == 6) {
      TDVC_CHECK_ARG(!c->out_packed && c->y && c->maskp && c->halo >= 0 && c->t_valid > 0 && c->Tout == c->t_valid + 2 * c->halo);
      TDVC_CHECK_ARG(c->halo == 0 || c->halo_buf != nullptr);
      TDVC_CHECK_ARG(c->y_grp_stride != 0 || c->y_b_stride != 0);
      if (c->yp) TDVC_CHECK_ARG(c->cp_out % 8 == 0 && c->out_ch_off % 8 == 0 && c->out_ch_stride % 8 == 0 && (uintptr_t)c->yp % 16 == 0);
    }
  }
  for (int g = 0; g < 4; ++g) {
    TDVC_CHECK_ARG(c->kg[g] >= 0 && c->kg[g] <= c->K && (c->kg[g] == 0 || ((c->K - c->kg[g]) % 2 == 0 && c->groups <= 4)));
  }
  if (c->B == 0) return TDVC_OK;
  TcP p{};
  p.B = c->B; p.Tout = c->Tout; p.Cout = c->Cout_g; p.K = c->K; p.dil = c->dilation; p.t_off = c->t_off;
  p.nchunk = cdiv(c->Cinp_g, TC_BK);
  int rem = c->Cinp_g - (p.nchunk - 1) * TC_BK;
  p.last_nk16 = cdiv(rem, 16);
  // N tile: the widest multiple of 16 up to 256 that divides the padded per-group width
  int bn = 0;
  for (int cand = std::min(c->Coutp_g, 256); cand >= 16; cand -= 16)
    if (c->Coutp_g % cand == 0) { bn = cand; break; }
  TDVC_CHECK_ARG(bn >= 16);
  {
    // Few, fat tiles (short sequences x wide layers: the discriminator's 1024 x 1024 x 5 conv at T <= 35 is 10 x 4 tiles of
    // N = 256) leave most SMs idle while every CTA streams its 2.6 MB weight slice through one SM's L2 port (measured: 53 us
    // against an 8 us tensor roofline).  Narrower N tiles put the same traffic on more SMs.
    const long long m_tiles = (long long)cdiv(c->Tout, TC_BM) * c->B * c->groups;
    while (bn > 64 && m_tiles * (c->Coutp_g / bn) * 2 <= num_sms()) {
      int next = 0;
      for (int cand = bn - 16; cand >= 64; cand -= 16)
        if (c->Coutp_g % cand == 0) { next = cand; break; }
      if (!next) break;
      bn = next;
    }
  }
  p.BN = bn;
  p.tiles_per_group = c->Coutp_g / bn;
  p.coutp_g = c->Coutp_g;
  p.a_ch_off = c->a_ch_off; p.a_ch_stride = c->a_ch_stride; p.bias_stride = c->bias_stride;
  int cols = 32;
  while (cols < p.BN) cols <<= 1;
  p.tmem_cols = cols;
  const int stage_bytes = TC_A_BYTES + p.BN * TC_BK * 2;
  // ~100 KB of pipeline per CTA so that two CTAs share an SM: one tile's epilogue overlaps the other's main loop
  int stages = (int)((100 * 1024) / stage_bytes);
  stages = std::max(stages, 2);
  stages = std::min(stages, 6);
  stages = std::min(stages, c->K * p.nchunk);
  stages = std::max(stages, 1);
  p.stages = stages;
  p.out_act = c->out_act; p.out_slope = c->out_slope; p.bias = c->bias; p.gb = c->gb; p.res = c->residual; p.y = c->y;
  p.yp = (__nv_bfloat16*)c->yp; p.tp_out = c->tp_out; p.cp_out = c->cp_out; p.out_halo = c->out_halo;
  p.out_ch_off = c->out_ch_off; p.out_ch_stride = c->out_ch_stride;
  p.maskp = (const __nv_bfloat16*)c->maskp; p.tm = c->tm; p.cm = c->cm; p.mask_halo = c->mask_halo;
  p.mask_ch_off = c->mask_ch_off; p.mask_ch_stride = c->mask_ch_stride; p.mask_slope = c->mask_slope;
  for (int g = 0; g < 4; ++g) p.kg[g] = c->kg[g];
  p.gb_grp_stride = c->gb_grp_stride; p.res_grp_stride = c->res_grp_stride;
  p.yp2 = (__nv_bfloat16*)c->yp2; p.auxp = (const __nv_bfloat16*)c->auxp; p.dgbp = (__nv_bfloat16*)c->dgbp;
  p.dgb_cp = c->dgb_cp; p.dgb_ch_off = c->dgb_ch_off; p.dgb_ch_stride = c->dgb_ch_stride;
  p.halo_buf = c->halo_buf; p.halo = c->halo; p.t_valid = c->t_valid; p.pk_slope = c->pk_slope;
  if (c->y_grp_stride == 0 && c->y_b_stride == 0) {      // default: group-major [groups][B][Cout_g][Tout]
    p.y_grp_stride = (long long)c->B * c->Cout_g * c->Tout;
    p.y_b_stride = (long long)c->Cout_g * c->Tout;
  } else {
    p.y_grp_stride = c->y_grp_stride; p.y_b_stride = c->y_b_stride;
  }
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("TDVC_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
  const int epi = c->gb ? 2 : (c->residual ? 1 : 0);
  const int act = c->out_act;
  const bool mask = c->maskp != nullptr;
  if (c->unframe_s > 0) {
    TDVC_CHECK_ARG(!chain && !c->out_packed && !mask && epi == 0 && act == TDVC_ACT_NONE && 16 % c->unframe_s == 0 &&
                   c->Cout_g % 16 == 0 && c->unframe_T > 0 && c->unframe_C > 0 &&
                   (long long)c->groups * c->Cout_g == (long long)c->unframe_C * c->unframe_s);
    p.unframe_s = c->unframe_s; p.unframe_pad = c->unframe_pad; p.unframe_T = c->unframe_T; p.unframe_C = c->unframe_C;
  }
  if (c->flat_tp > 0) {
    TDVC_CHECK_ARG(!chain && !c->out_packed && !mask && epi == 0 && c->groups == 1 && c->B == 1 && c->unframe_s == 0 &&
                   c->flat_halo >= 0 && c->flat_T > 0 && c->flat_T + c->flat_halo <= c->flat_tp && c->Tout % c->flat_tp == 0);
    p.flat_tp = c->flat_tp; p.flat_halo = c->flat_halo; p.flat_T = c->flat_T;
  }
  // ---- weight-stationary persistent variant when the whole weight tile set fits next to an activation ring
  {
    static int ws_on = -1;
    if (ws_on < 0) { const char* e = getenv("TDVC_TC_WS"); ws_on = e ? atoi(e) : 1; }
    const int halo_rows = (c->K - 1) * c->dilation;
    const int rows_a = ((TC_BM + halo_rows + 7) / 8) * 8;
    const bool narrow = (c->Cinp_g % TC_BK) == 16 && c->Cinp_g > TC_BK;      // e.g. 144 = 64 + 64 + 16
    const int n_full = narrow ? p.nchunk - 1 : p.nchunk;
    const long long w_tile = (long long)p.BN * TC_BK * 2, wn_tile = (long long)p.BN * 32;
    const long long w_bytes = (long long)c->K * n_full * w_tile + (narrow ? (long long)c->K * wn_tile : 0);
    const long long a_stage = (long long)rows_a * TC_BK * 2, an_stage = (long long)rows_a * 32;
    const long long budget = 224LL * 1024 - 2048;
    const int mtiles_per_b = cdiv(c->Tout, TC_BM);
    const int n_mtiles = mtiles_per_b * c->B;
    const int n_tiles_total = c->groups * p.tiles_per_group;
    int ctas = std::max(1, num_sms() / n_tiles_total);
    ctas = std::min(ctas, n_mtiles);
    const int n_stages = narrow ? 3 : 0;
    const long long fixed = w_bytes + n_stages * an_stage;
    static int min_tiles = -1;     // TDVC_TC_WS_MIN_TILES: time tiles per CTA from which the persistent kernel is used
    if (min_tiles < 0) { const char* e = getenv("TDVC_TC_WS_MIN_TILES"); min_tiles = e ? std::max(1, atoi(e)) : 4; }
    if (ws_on && rows_a <= 256 && fixed + 2 * a_stage <= budget && (n_mtiles >= min_tiles * ctas || ws_on == 2) && 2 * p.BN <= 512) {   // TDVC_TC_WS=2 forces it (tests)
      WsP w{};
      w.n_mtiles = n_mtiles; w.mtiles_per_b = mtiles_per_b; w.rows_a = rows_a; w.a_stage_bytes = (int)a_stage;
      w.w_tile_bytes = (int)w_tile; w.narrow = narrow ? 1 : 0; w.n_full = n_full; w.an_stage_bytes = (int)an_stage;
      w.wn_tile_bytes = (int)wn_tile; w.n_stages = n_stages;
      int st = (int)std::min<long long>(8, (budget - fixed) / a_stage);
      {
        static int cap = -1;      // TDVC_TC_WS_STAGES: development knob (ring-depth sensitivity)
        if (cap < 0) { const char* e = getenv("TDVC_TC_WS_STAGES"); cap = e ? atoi(e) : 0; }
        if (cap >= 2) st = std::min(st, cap);
      }
      p.stages = st;
      // narrow tiles (the 16 / 32-channel MRF stages): one tile's epilogue occupies 4 / 8 of the 16 epilogue warps and is a
      // chain of dependent memory round trips (~2.5 us measured), so several tiles are drained at once by separate warp teams
      static int teams_on = -1;     // TDVC_TC_EPI_TEAMS=0: development switch, one team as before
      if (teams_on < 0) { const char* e = getenv("TDVC_TC_EPI_TEAMS"); teams_on = e ? atoi(e) : 1; }
      p.epi_spt = !teams_on ? 4 : (p.BN <= 16 ? 1 : (p.BN <= 32 ? 2 : 4));
      p.nacc = 2 * (4 / p.epi_spt);
      int cols2 = 32;
      while (cols2 < p.nacc * p.BN) cols2 <<= 1;
      p.tmem_cols = cols2;
      size_t smem_ws = (size_t)fixed + (size_t)st * a_stage + (2 * st + 2 * n_stages + 2 * p.nacc + 1) * sizeof(uint64_t) + 16 + 1024;
      typedef void (*WsFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, TcP, WsP);
      WsFn kern = nullptr;
      if (chain == 3) kern = conv_tc_ws_k<1, 3, 1, 0>;
      else if (chain == 4) kern = conv_tc_ws_k<0, 4, 0, 0>;
      else if (chain == 5) kern = conv_tc_ws_k<0, 5, 1, 1>;
      else if (chain == 6) kern = conv_tc_ws_k<0, 6, 0, 1>;
      else if (!c->out_packed && !mask) {
        static const WsFn table[3][3] = {
            {conv_tc_ws_k<0, 0, 0, 0>, conv_tc_ws_k<0, 1, 0, 0>, conv_tc_ws_k<0, 2, 0, 0>},
            {conv_tc_ws_k<1, 0, 0, 0>, conv_tc_ws_k<1, 1, 0, 0>, conv_tc_ws_k<1, 2, 0, 0>},
            {conv_tc_ws_k<2, 0, 0, 0>, conv_tc_ws_k<2, 1, 0, 0>, conv_tc_ws_k<2, 2, 0, 0>}};
        kern = table[act][epi];
      } else if (!c->out_packed && mask) {
        kern = conv_tc_ws_k<0, 0, 0, 1>;
      } else if (c->out_packed && !mask) {
        kern = act == TDVC_ACT_LRELU ? conv_tc_ws_k<1, 0, 1, 0> : conv_tc_ws_k<0, 0, 1, 0>;
      } else {
        kern = conv_tc_ws_k<0, 0, 1, 1>;
      }
      TDVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
      CUtensorMap map_a, map_b, map_an, map_bn;
      int rc = make_map_3d(&map_a, c->xp, (uint64_t)c->Cp_total, (uint64_t)c->Tp, (uint64_t)c->B, TC_BK, (uint32_t)rows_a);
      if (rc) return rc;
      rc = make_map_3d(&map_b, c->wp, (uint64_t)c->Cinp_g, (uint64_t)c->groups * c->Coutp_g, (uint64_t)c->K, TC_BK, (uint32_t)p.BN);
      if (rc) return rc;
      if (narrow) {
        rc = make_map_3d(&map_an, c->xp, (uint64_t)c->Cp_total, (uint64_t)c->Tp, (uint64_t)c->B, 16, (uint32_t)rows_a,
                         CU_TENSOR_MAP_SWIZZLE_32B);
        if (rc) return rc;
        rc = make_map_3d(&map_bn, c->wp, (uint64_t)c->Cinp_g, (uint64_t)c->groups * c->Coutp_g, (uint64_t)c->K, 16,
                         (uint32_t)p.BN, CU_TENSOR_MAP_SWIZZLE_32B);
        if (rc) return rc;
      } else {
        map_an = map_a; map_bn = map_b;
      }
      dim3 grid(ctas, n_tiles_total, 1);
      TDVC_CHECK_ARG(grid.y <= 65535);
      bool any_kg = false;
      for (int g = 0; g < 4; ++g) any_kg = any_kg || c->kg[g] > 0;
      static int balance = -1;     // TDVC_TC_BALANCE=1 (off by default: measured 2x SLOWER on the MRF chain launches -- a tile's
                                   // cost there is its epilogue, the same for every branch, not its 3 / 7 / 11 MMAs)
      if (balance < 0) { const char* e = getenv("TDVC_TC_BALANCE"); balance = e ? atoi(e) : 0; }
      if (balance && any_kg && p.tiles_per_group == 1 && c->groups >= 2 && c->groups <= 4) {
        // CTAs per group in proportion to its taps (its share of the MMAs), at least one each
        int taps[4], tot = 0, total_ctas = std::min(num_sms(), c->groups * n_mtiles), used = 0;
        for (int g = 0; g < c->groups; ++g) { taps[g] = c->kg[g] > 0 ? c->kg[g] : c->K; tot += taps[g]; }
        p.balanced = 1;
        p.grp_cta0[0] = 0;
        for (int g = 0; g < c->groups; ++g) {
          int n = std::max(1, std::min(n_mtiles, (int)((long long)total_ctas * taps[g] / tot)));
          used += n;
          p.grp_cta0[g + 1] = used;
        }
        for (int g = c->groups; g < 4; ++g) p.grp_cta0[g + 1] = used;
        grid = dim3(used, 1, 1);
      }
      tdvc::launch_k(kern, grid, TC_FWD_THREADS, smem_ws, (cudaStream_t)stream, map_a, map_b, map_an, map_bn, p, w);
      TDVC_LAUNCH_CHECK();
      g_flops[chain ? FLOP_TC_CHAIN : FLOP_TC_WS] += 2.0 * c->B * c->Tout * (double)c->Cout_g * c->Cinp_g * tc_taps(c);
      return TDVC_OK;
    }
  }
  size_t smem = (size_t)stages * stage_bytes + (2 * stages + 1) * sizeof(uint64_t) + 16 + 1024;
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, TcP);
  KernelFn kern = nullptr;
  {
    // K > 1: the haloed-tile form (conv_tc_fwdh_k) when tile + halo fit one TMA box (256 rows)
    static int halo_on = -1;     // TDVC_TC_FWD_HALO=0: development switch, per-tap activation tiles (conv_tc_fwd_k)
    if (halo_on < 0) { const char* e = getenv("TDVC_TC_FWD_HALO"); halo_on = e ? atoi(e) : 1; }
    const int a_rows = (TC_BM + (c->K - 1) * c->dilation + 7) / 8 * 8;
    const long long a_bytes = (long long)a_rows * 128, b_bytes = (long long)p.BN * TC_BK * 2;
    const int b_st = (int)std::min<long long>(6, (196LL * 1024 - 2 * a_bytes) / b_bytes);
    if (halo_on && c->K > 1 && a_rows <= 256 && b_st >= 2) {
      p.a_rows = a_rows; p.b_stages = b_st;
      const size_t smem_h = (size_t)(2 * a_bytes + b_st * b_bytes) + (2 * 2 + 2 * b_st + 1) * sizeof(uint64_t) + 16 + 1024;
      KernelFn kh = nullptr;
      if (chain == 3) kh = conv_tc_fwdh_k<1, 3, 1, 0>;
      else if (chain == 4) kh = conv_tc_fwdh_k<0, 4, 0, 0>;
      else if (chain == 5) kh = conv_tc_fwdh_k<0, 5, 1, 1>;
      else if (chain == 6) kh = conv_tc_fwdh_k<0, 6, 0, 1>;
      else if (!c->out_packed && !mask) {
        static const KernelFn table[3][3] = {
            {conv_tc_fwdh_k<0, 0, 0, 0>, conv_tc_fwdh_k<0, 1, 0, 0>, conv_tc_fwdh_k<0, 2, 0, 0>},
            {conv_tc_fwdh_k<1, 0, 0, 0>, conv_tc_fwdh_k<1, 1, 0, 0>, conv_tc_fwdh_k<1, 2, 0, 0>},
            {conv_tc_fwdh_k<2, 0, 0, 0>, conv_tc_fwdh_k<2, 1, 0, 0>, conv_tc_fwdh_k<2, 2, 0, 0>}};
        kh = table[act][epi];
      } else if (!c->out_packed && mask) {
        kh = conv_tc_fwdh_k<0, 0, 0, 1>;
      } else if (c->out_packed && !mask) {
        kh = act == TDVC_ACT_LRELU ? conv_tc_fwdh_k<1, 0, 1, 0> : conv_tc_fwdh_k<0, 0, 1, 0>;
      } else {
        kh = conv_tc_fwdh_k<0, 0, 1, 1>;
      }
      TDVC_CUDA(cudaFuncSetAttribute(kh, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      CUtensorMap map_a, map_b;
      int rc = make_map_3d(&map_a, c->xp, (uint64_t)c->Cp_total, (uint64_t)c->Tp, (uint64_t)c->B, TC_BK, (uint32_t)a_rows);
      if (rc) return rc;
      rc = make_map_3d(&map_b, c->wp, (uint64_t)c->Cinp_g, (uint64_t)c->groups * c->Coutp_g, (uint64_t)c->K, TC_BK, (uint32_t)p.BN);
      if (rc) return rc;
      dim3 grid(cdiv(c->Tout, TC_BM), c->groups * p.tiles_per_group, c->B);
      TDVC_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535);
      tdvc::launch_k(kh, grid, TC_FWD_THREADS, smem_h, (cudaStream_t)stream, map_a, map_b, p);
      TDVC_LAUNCH_CHECK();
      g_flops[chain ? FLOP_TC_CHAIN : FLOP_TC_TILE] += 2.0 * c->B * c->Tout * (double)c->Cout_g * c->Cinp_g * tc_taps(c);
      return TDVC_OK;
    }
  }
  if (chain == 3) kern = conv_tc_fwd_k<1, 3, 1, 0>;
  else if (chain == 4) kern = conv_tc_fwd_k<0, 4, 0, 0>;
  else if (chain == 5) kern = conv_tc_fwd_k<0, 5, 1, 1>;
  else if (chain == 6) kern = conv_tc_fwd_k<0, 6, 0, 1>;
  else if (!c->out_packed && !mask) {
    static const KernelFn table[3][3] = {
        {conv_tc_fwd_k<0, 0, 0, 0>, conv_tc_fwd_k<0, 1, 0, 0>, conv_tc_fwd_k<0, 2, 0, 0>},
        {conv_tc_fwd_k<1, 0, 0, 0>, conv_tc_fwd_k<1, 1, 0, 0>, conv_tc_fwd_k<1, 2, 0, 0>},
        {conv_tc_fwd_k<2, 0, 0, 0>, conv_tc_fwd_k<2, 1, 0, 0>, conv_tc_fwd_k<2, 2, 0, 0>}};
    kern = table[act][epi];
  } else if (!c->out_packed && mask) {
    kern = conv_tc_fwd_k<0, 0, 0, 1>;
  } else if (c->out_packed && !mask) {
    kern = act == TDVC_ACT_LRELU ? conv_tc_fwd_k<1, 0, 1, 0> : conv_tc_fwd_k<0, 0, 1, 0>;
  } else {
    kern = conv_tc_fwd_k<0, 0, 1, 1>;
  }
  TDVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUtensorMap map_a, map_b;
  int rc = make_map_3d(&map_a, c->xp, (uint64_t)c->Cp_total, (uint64_t)c->Tp, (uint64_t)c->B, TC_BK, TC_BM);
  if (rc) return rc;
  rc = make_map_3d(&map_b, c->wp, (uint64_t)c->Cinp_g, (uint64_t)c->groups * c->Coutp_g, (uint64_t)c->K, TC_BK, (uint32_t)p.BN);
  if (rc) return rc;
  dim3 grid(cdiv(c->Tout, TC_BM), c->groups * p.tiles_per_group, c->B);
  TDVC_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535);
  tdvc::launch_k(kern, grid, TC_FWD_THREADS, smem, (cudaStream_t)stream, map_a, map_b, p);
  TDVC_LAUNCH_CHECK();
  g_flops[chain ? FLOP_TC_CHAIN : FLOP_TC_TILE] += 2.0 * c->B * c->Tout * (double)c->Cout_g * c->Cinp_g * tc_taps(c);
  return TDVC_OK;
}

extern "C" int tdvc_conv1d_tc_fwd(const void* xp, const void* wp, const float* bias, const float* gb, const float* residual,
                                  float* y, int B, int Cinp, int Tp, int Cout, int Coutp, int Tout, int K, int dilation,
                                  int t_off, int out_act, float out_slope, void* stream) {
  tdvc_tc_conv c{};
  c.xp = xp; c.wp = wp; c.bias = bias; c.gb = gb; c.residual = residual; c.y = y;
  c.B = B; c.Tp = Tp; c.Tout = Tout; c.K = K; c.dilation = dilation; c.t_off = t_off;
  c.Cp_total = Cinp; c.groups = 1; c.a_ch_off = 0; c.a_ch_stride = 0; c.Cinp_g = Cinp;
  c.Cout_g = Cout; c.Coutp_g = Coutp; c.bias_stride = 0;
  c.out_act = out_act; c.out_slope = out_slope;
  return tdvc_conv1d_tc_fwd_ex(&c, stream);
}

extern "C" int tdvc_conv1d_tc_fwd_stacked(const void* xp, const void* wp, const float* bias, void* yp, int B, int Cp_total,
                                          int a_ch_off, int Cinp, int Tp, int Tout, int K, int dilation, int t_off, int R,
                                          int n_blocks, int rows_per_block, int out_act, float out_slope, int tp_out,
                                          int cp_out, int out_halo, int out_ch_off, int out_ch_stride, void* stream) {
  TDVC_CHECK_ARG(xp && wp && yp && B >= 0 && Tp > 0 && Tout > 0 && K > 0 && dilation > 0);
  TDVC_CHECK_ARG(Cp_total % 8 == 0 && Cinp % 16 == 0 && Cinp >= TC_BK && a_ch_off % 8 == 0 && a_ch_off + Cinp <= Cp_total);
  TDVC_CHECK_ARG(n_blocks > 0 && rows_per_block > 0 && (long long)n_blocks * rows_per_block <= R);
  TDVC_CHECK_ARG(out_act == TDVC_ACT_NONE || out_act == TDVC_ACT_LRELU);
  TDVC_CHECK_ARG(out_ch_stride >= rows_per_block && out_ch_off >= 0 && out_halo >= 0 &&
                 out_ch_off + (long long)(n_blocks - 1) * out_ch_stride + out_ch_stride <= cp_out && tp_out >= Tout + out_halo);
  TDVC_CHECK_ARG(((uintptr_t)xp % 16 == 0) && ((uintptr_t)wp % 16 == 0));
  if (B == 0) return TDVC_OK;
  TcP p{};
  p.B = B; p.Tout = Tout; p.K = K; p.dil = dilation; p.t_off = t_off; p.a_ch_off = a_ch_off;
  p.out_act = out_act; p.out_slope = out_slope; p.bias = bias;
  p.yp = (__nv_bfloat16*)yp; p.tp_out = tp_out; p.cp_out = cp_out; p.out_halo = out_halo;
  p.out_ch_off = out_ch_off; p.out_ch_stride = out_ch_stride;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("TDVC_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
  WtP w{};
  const int nchunk = cdiv(Cinp, TC_BK);
  w.narrow = ((Cinp % TC_BK) == 16 && Cinp > TC_BK) ? 1 : 0;
  w.n_full = w.narrow ? nchunk - 1 : nchunk;
  w.last_nk16 = cdiv(Cinp - (nchunk - 1) * TC_BK, 16);
  w.rows_total = n_blocks * rows_per_block; w.rows_per_block = rows_per_block;
  w.R = R; w.cinp = Cinp; w.wp = (const __nv_bfloat16*)wp;
  w.m_tiles = cdiv(w.rows_total, TC_BM);
  TDVC_CHECK_ARG(w.m_tiles <= num_sms());
  // TDVC_WT_TS=1: weights in tensor memory (when they fit next to two 128-column accumulators).  Off by default: measured
  // on B200 at the cond_var.0 shape the N=128 TS form reaches 967 TFLOP/s in its main loop against 1278 for the N=256 SS
  // form -- back-to-back MMAs into one accumulator do not run at the N/2-cycle floor at N=128 (profiles/README.md).
  static int ts_on = -1;
  if (ts_on < 0) { const char* e = getenv("TDVC_WT_TS"); ts_on = e ? atoi(e) : 0; }
  const bool ts = ts_on && K * Cinp / 2 <= 512 - WT_TS_WCOL;
  const int bn = ts ? 128 : 256;
  w.ttiles_per_b = cdiv(Tout, bn);
  w.n_ttiles = w.ttiles_per_b * B;
  w.ctas_per_m = std::max(1, std::min(num_sms() / w.m_tiles, w.n_ttiles));
  const int rows_needed = bn + (K - 1) * dilation;
  w.nsplit = cdiv(rows_needed, 256);
  w.box_rows = 8 * cdiv(cdiv(rows_needed, 8), w.nsplit);       // TMA boxes are <= 256 rows and whole 8-row swizzle atoms
  TDVC_CHECK_ARG(w.box_rows <= 256);
  w.a_stage_bytes = w.nsplit * w.box_rows * 128;
  w.an_stage_bytes = w.nsplit * w.box_rows * 32;
  // 16-byte stores need every 8-channel chunk of a warp to sit inside one block at a 16-byte aligned column
  w.vec = (rows_per_block % 8 == 0 && out_ch_stride % 8 == 0 && out_ch_off % 8 == 0 && cp_out % 8 == 0 &&
           (uintptr_t)yp % 16 == 0) ? 1 : 0;
  const long long budget = 227LL * 1024 - 1024 - 512;
  const long long w_smem = ts ? 0 : (long long)K * (w.n_full * WT_W_TILE + (w.narrow ? WT_WN_TILE : 0));
  const long long stg = w.vec ? TC_EPI_WARPS * WT_STG_BYTES : 0;
  // ring depth: whole tiles' worth of chunks, up to 4 tiles ahead (TS) / whatever is left beside the weights (SS)
  int st, nst;
  if (ts) {
    const long long per_tile = (long long)w.n_full * w.a_stage_bytes + (w.narrow ? w.an_stage_bytes : 0);
    int tiles = (int)std::min<long long>(4, (budget - stg) / per_tile);
    if (tiles < 1) { set_error("conv1d_tc_fwd_stacked: one activation tile (%lld B) does not fit in shared memory", per_tile); return TDVC_ERR_ARG; }
    st = std::max(2, tiles * w.n_full);
    nst = w.narrow ? std::max(2, tiles) : 0;
  } else {
    nst = w.narrow ? 3 : 0;
    st = (int)std::min<long long>(6, (budget - w_smem - stg - (long long)nst * w.an_stage_bytes) / w.a_stage_bytes);
  }
  w.n_stages = nst;
  const long long fixed = w_smem + stg + (long long)nst * w.an_stage_bytes;
  if (st < 2 || fixed + (long long)st * w.a_stage_bytes > budget) {
    set_error("conv1d_tc_fwd_stacked: weights (%lld B) + two activation stages do not fit in shared memory", w_smem);
    return TDVC_ERR_ARG;
  }
  p.stages = st;
  const size_t smem = (size_t)fixed + (size_t)st * w.a_stage_bytes + (2 * st + 2 * nst + 5) * sizeof(uint64_t) + 16 + 1024;
  typedef void (*WtFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, TcP, WtP);
  static const WtFn table[2][2] = {{conv_tc_wt_k<0, 0>, conv_tc_wt_k<0, 1>}, {conv_tc_wt_k<1, 0>, conv_tc_wt_k<1, 1>}};
  WtFn kern = table[out_act == TDVC_ACT_LRELU ? 1 : 0][ts ? 1 : 0];
  TDVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CUtensorMap map_x, map_w, map_xn, map_wn;
  int rc = make_map_3d(&map_x, xp, (uint64_t)Cp_total, (uint64_t)Tp, (uint64_t)B, TC_BK, (uint32_t)w.box_rows);
  if (rc) return rc;
  rc = make_map_3d(&map_w, wp, (uint64_t)Cinp, (uint64_t)R, (uint64_t)K, TC_BK, TC_BM);
  if (rc) return rc;
  if (w.narrow) {
    rc = make_map_3d(&map_xn, xp, (uint64_t)Cp_total, (uint64_t)Tp, (uint64_t)B, 16, (uint32_t)w.box_rows, CU_TENSOR_MAP_SWIZZLE_32B);
    if (rc) return rc;
    rc = make_map_3d(&map_wn, wp, (uint64_t)Cinp, (uint64_t)R, (uint64_t)K, 16, TC_BM, CU_TENSOR_MAP_SWIZZLE_32B);
    if (rc) return rc;
  } else {
    map_xn = map_x; map_wn = map_w;
  }
  tdvc::launch_k(kern, w.m_tiles * w.ctas_per_m, TC_FWD_THREADS, smem, (cudaStream_t)stream, map_x, map_w, map_xn, map_wn, p, w);
  TDVC_LAUNCH_CHECK();
  g_flops[FLOP_TC_WT] += 2.0 * B * Tout * (double)w.rows_total * Cinp * K;
  return TDVC_OK;
}

// workspace (floats) for tdvc_conv1d_tc_wgrad
extern "C" int64_t tdvc_conv1d_tc_wgrad_ws(int Cout, int Cin, int K) {
  const long long Mp = (long long)((Cin + 127) / 128) * 128, Np = (long long)((Cout + 15) / 16) * 16;
  return (int64_t)K * Np * Mp + Np;      // + the bias-gradient row
}

// dw[Cout,Cin,K] (OVERWRITTEN) from the packed operands: dyp[B,Tout,Cdp] and xp[B,Tp,Cp] (both bf16 channels-last;
// xp row = t + tap*dilation + t_off).
extern "C" int tdvc_conv1d_tc_wgrad(const void* dyp, const void* xp, float* dw, float* ws, int B, int Cdp, int Tout, int Cp,
                                    int Tp, int Cout, int Cin, int K, int dilation, int t_off, int x_ch_off,
                                    int dy_ch_off, float* db, int ws_is_zero, void* stream) {
  TDVC_CHECK_ARG(dyp && xp && dw && ws && B >= 0 && Cdp % 8 == 0 && Cp % 8 == 0 && dy_ch_off >= 0 && x_ch_off >= 0 &&
                 Cdp >= dy_ch_off + Cout && Cp >= x_ch_off + Cin && Tout > 0 && Tp > 0 && K > 0 && dilation > 0);
  cudaStream_t st = (cudaStream_t)stream;
  WgTcP p{};
  p.B = B; p.Tout = Tout; p.Cout = Cout; p.Cin = Cin; p.K = K; p.dil = dilation; p.t_off = t_off; p.ws = ws;
  p.x_ch_off = x_ch_off; p.dy_ch_off = dy_ch_off;
  p.Mp = ((Cin + 127) / 128) * 128;
  p.Np = ((Cout + 15) / 16) * 16;
  p.bias = db ? 1 : 0;
  if (!ws_is_zero) TDVC_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * ((size_t)K * p.Np * p.Mp + p.Np), st));
  if (B == 0) {
    TDVC_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * Cin * K, st));
    if (db) TDVC_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)Cout, st));
    return TDVC_OK;
  }
  const int xt = p.bias;     // accumulators beyond the taps
  const int N16 = p.Np;
  // widest co tile such that as many taps as possible share the 512 TMEM columns (fewer re-reads of the operands)
  int best_nt = 16, best_cost = 1 << 30;
  for (int cand = std::min(N16, 256); cand >= 16; cand -= 16) {
    int kt = std::min(K, 512 / cand - xt);
    if (kt < 1) continue;
    // per-stage shared memory must leave room for at least 2 stages
    long long stage = (long long)kt * 2 * WG_BOX_BYTES + (long long)cdiv(cand, 64) * WG_BOX_BYTES;
    if (2 * stage > 200 * 1024) continue;
    int cost = cdiv(K, kt) * cdiv(N16, cand);
    if (cost < best_cost) { best_cost = cost; best_nt = cand; }
  }
  p.NT = best_nt;
  p.KT = std::min(K, 512 / p.NT - xt);
  while (p.KT > 1 && 2LL * ((long long)p.KT * 2 * WG_BOX_BYTES + (long long)cdiv(p.NT, 64) * WG_BOX_BYTES) > 200 * 1024) --p.KT;
  p.ntap_groups = cdiv(K, p.KT);
  p.n_ntiles = cdiv(N16, p.NT);
  p.nb = cdiv(p.NT, 64);
  int cols = 32;
  while (cols < (p.KT + xt) * p.NT) cols <<= 1;
  TDVC_CHECK_ARG(cols <= 512);
  p.tmem_cols = cols;
  const int stage_bytes = p.KT * 2 * WG_BOX_BYTES + p.nb * WG_BOX_BYTES;
  p.nchunk_t = cdiv(Tout, 64);
  p.units = B * p.nchunk_t;
  int stages = std::min(4, (int)((200 * 1024) / stage_bytes));
  TDVC_CHECK_ARG(stages >= 1);
  stages = std::min(stages, std::max(1, p.units));
  p.stages = stages;
  const int m_tiles = p.Mp / 128;
  const int gy = m_tiles * p.ntap_groups * p.n_ntiles;
  int splits = std::max(1, num_sms() / gy);
  splits = std::min(splits, p.units);
  p.splits = splits;
  size_t smem = (size_t)stages * stage_bytes + (p.bias ? 2 * WG_BOX_BYTES : 0) + (2 * stages + 1) * sizeof(uint64_t) + 16 + 1024;
  // per-device attribute: set on every call (cheap), not once per process
  TDVC_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CUtensorMap map_x, map_dy;
  int rc = make_map_3d(&map_x, xp, (uint64_t)Cp, (uint64_t)Tp, (uint64_t)B, 64, 64);
  if (rc) return rc;
  rc = make_map_3d(&map_dy, dyp, (uint64_t)Cdp, (uint64_t)Tout, (uint64_t)B, 64, 64);
  if (rc) return rc;
  TDVC_CHECK_ARG(gy <= 65535);
  tdvc::launch_k(conv_tc_wgrad_k, dim3(splits, gy), TC_THREADS, smem, st, map_x, map_dy, p);
  TDVC_LAUNCH_CHECK();
  g_flops[FLOP_TC_WGRAD] += 2.0 * B * Tout * (double)Cout * Cin * K;
  long long n = (long long)Cout * Cin * K;
  int blocks = (int)std::min<long long>((n + 255) / 256, 4LL * num_sms());
  if (ws_is_zero) tdvc::launch_k(wgrad_finalize_zero_k, blocks, 256, 0, st, ws, dw, Cout, Cin, K, p.Np, p.Mp, db, p.bias);
  else tdvc::launch_k(wgrad_finalize_k, blocks, 256, 0, st, ws, dw, Cout, Cin, K, p.Np, p.Mp, db);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

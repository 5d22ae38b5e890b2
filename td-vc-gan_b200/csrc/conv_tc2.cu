// Second generation of the tcgen05 weight-gradient kernel, and the frame-view pack / unpack kernels of the
// discriminators' grouped strided convolutions.
//
// conv_tc_wgrad2_k -- dW[g][tap][co][ci] = sum_{b,t} dy[b, t, dy_off_g + co] * x[b, t + tap*dil + t_off_g, x_off_g + ci]
//   for up to thousands of independent groups in ONE launch (grid.z = group):
//     * the 3 kernel-size branches of an MRF depth (k = 3 / 7 / 11 at the same dilation: per-group tap count and offset),
//     * the 4 ... 256 conv groups of a discriminator layer in frame view (uniform groups),
//     * a plain conv (one group).
//   GEMM per (group, tap): M = ci (TMEM lanes), N = co (TMEM columns, one accumulator per tap), K = time.  Both operands are
//   MN-major (channel-contiguous) SWIZZLE_128B views of the packed bf16 channels-last tensors the forward / data-gradient
//   kernels already use.  What is new against conv_tc_wgrad_k:
//     * the x tile of a 64-step time unit is loaded ONCE with its (KT-1)*dil halo rows and every tap is a row-shifted view
//       of it (the 128-byte swizzle is a function of the absolute shared-memory address, so a start address moved by whole
//       128-byte rows needs no descriptor fix-up) -- the first kernel loaded one shifted copy per tap: 11 x the bytes for
//       k = 11, and a single-stage pipeline because 11 copies filled the shared memory;
//     * only the 64-channel boxes that hold real channels are loaded (Cin <= 64: one box instead of two);
//     * groups share a launch, so a 16-channel layer no longer costs one launch + one finalize per conv;
//     * TAPS ON M (Cin = 16 / 32 / 64): a tcgen05.mma costs about the same whatever part of its 128 M lanes holds real
//       channels (measured: ~128 cycles per instruction at N = 16), so a 16-channel layer used 1/8 of every instruction.
//       The x tile is loaded as Cin-channel rows (SWIZZLE_32B / 64B / 128B box) and the A descriptor's leading-dimension
//       offset -- the distance between consecutive Cin-channel atoms along M -- is set to dil rows: atom a of the operand is
//       the tile moved down by a*dil rows, i.e. tap a.  One instruction then computes 128 / Cin taps: 8 instead of 44
//       MMAs per time unit for a 16-channel k = 11 conv.
//   Split-K over (batch, 64-step chunk) with lane-coalesced fp32 reductions into an always-zero workspace, as before;
//   wgrad2_finalize_k moves the sums into the weight-gradient tensors (plain [Cout][Cin][K] per group, or the
//   discriminator's [Cout][Cin/groups][41] through the frame mapping) and re-zeroes what it read.
#include "tc_common.cuh"

namespace tdvc {

constexpr int W2_THREADS = 192;        // TMA warp, MMA warp, 4 epilogue warps
constexpr int W2_TK = 64;              // time steps per unit
constexpr int W2_BOX = 64 * 128;       // 64 rows x 64 channels bf16
constexpr int W2_MAXG = 4;             // groups with individual tap counts / offsets / output tensors

struct Wg2P {
  int B, Tout, Cout, Cin, K, dil;
  int ngroups, per_group;
  int x_ch_off, x_ch_stride, dy_ch_off, dy_ch_stride;
  int kg[W2_MAXG], toff[W2_MAXG];
  int NT, n_ntiles, KT, ntap_groups, nb_dy, nb_x, rows_x, haloed;
  int stages, tmem_cols, nchunk_t, units, splits;
  int Mp, Np;
  long long ws_grp_stride;
  float* ws;
  int bias;
  int k_inner;
  int nacc_max;        // accumulator slots per CTA (the bias accumulator follows them)
  int tm, sw;          // taps per MMA (1 = channels only on M) and swizzle / row bytes of the x tile (128 unless tm > 1)
};

// MN-major descriptor with explicit leading-dimension / stride byte offsets and layout code (2 = SWIZZLE_128B, 4 = 64B, 6 = 32B)
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

__global__ void __launch_bounds__(W2_THREADS) conv_tc_wgrad2_k(const __grid_constant__ CUtensorMap map_x,
                                                               const __grid_constant__ CUtensorMap map_dy, Wg2P p) {
  pdl_prologue_top();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // stage = [x: two 64-channel boxes (haloed: rows_x rows each; per tap: KT pairs of 64 rows)] [dy: nb_dy boxes of 64 rows]
  const int x_box = (p.haloed ? p.rows_x : W2_TK) * p.sw;
  const int x_bytes = (p.haloed ? 1 : p.KT) * 2 * x_box;
  const int dy_bytes = p.nb_dy * W2_BOX;
  const int stage_bytes = x_bytes + dy_bytes;
  uint8_t* ones = smem + (size_t)p.stages * stage_bytes;             // 2 x 64 x 64 bf16 of 1.0 (bias row), when p.bias
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ones + (p.bias ? 2 * W2_BOX : 0));
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.z;
  int yy = blockIdx.y;
  const int nt_i = yy % p.n_ntiles; yy /= p.n_ntiles;
  const int tg = yy % p.ntap_groups;
  const int mt = yy / p.ntap_groups;
  const int ci0 = mt * 128, n0 = nt_i * p.NT, tap0 = tg * p.KT;
  const int gi = p.per_group ? g : 0;
  const int kgrp = p.per_group ? p.kg[gi] : p.K;
  const int t_off = p.toff[gi];
  const int ntaps = min(p.KT, kgrp - tap0);
  const int nacc = (ntaps + p.tm - 1) / p.tm;            // accumulators in use: one per tap, or per block of tm taps
  const int split = blockIdx.x;
  const int my_units = ntaps > 0 ? (p.units - split + p.splits - 1) / p.splits : 0;
  const bool do_bias = p.bias && mt == 0 && tg == 0;
  const int xc = p.x_ch_off + g * p.x_ch_stride + ci0;
  const int dc = p.dy_ch_off + g * p.dy_ch_stride + n0;
  // real 64-channel boxes of this CTA's 128-lane ci tile
  const int nbx = p.tm > 1 ? 1 : min(p.nb_x, (p.Cin - ci0 + 63) / 64);
  if (do_bias) {
    uint32_t* o32 = reinterpret_cast<uint32_t*>(ones);
    for (int i = threadIdx.x; i < 2 * W2_BOX / 4; i += W2_THREADS) o32[i] = 0x3F803F80u;      // bf16 1.0 pairs
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
  pdl_prologue_late();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (my_units > 0) {
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t tx = (uint32_t)((p.haloed ? nbx : ntaps * nbx) * x_box + dy_bytes);
        for (int it = 0; it < my_units; ++it) {
          const int u = split + it * p.splits;
          const int b = u / p.nchunk_t, tc = (u - b * p.nchunk_t) * W2_TK;
          const int s = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* sx = smem + (size_t)s * stage_bytes;
          mbar_expect_tx(&full_bar[s], tx);
          if (p.haloed) {
            const int row0 = tc + tap0 * p.dil + t_off;
            for (int j = 0; j < nbx; ++j) tma_load_3d(sx + j * x_box, &map_x, &full_bar[s], xc + 64 * j, row0, b);
          } else {
            for (int tp = 0; tp < ntaps; ++tp)
              for (int j = 0; j < nbx; ++j)
                tma_load_3d(sx + (tp * 2 + j) * x_box, &map_x, &full_bar[s], xc + 64 * j, tc + (tap0 + tp) * p.dil + t_off, b);
          }
          for (int j = 0; j < p.nb_dy; ++j)
            tma_load_3d(sx + x_bytes + j * W2_BOX, &map_dy, &full_bar[s], dc + 64 * j, tc, b);
        }
      }
    } else if (warp == 1) {
      // D = f32, A = B = bf16, both MN-major (bits 15, 16), N = NT, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      // descriptor constants (units of 16 bytes).  A: x tile, rows of p.sw bytes; LBO = next M atom (tm > 1: dil rows below,
      // else the second 64-channel box), SBO = 8 rows.  B: dy tile, SWIZZLE_128B, LBO = next 64-channel box.
      const uint32_t row_b = (uint32_t)p.sw;
      const uint32_t layout = p.sw == 128 ? 2u : (p.sw == 64 ? 4u : 6u);
      const uint32_t lbo = p.tm > 1 ? (uint32_t)p.dil * row_b : (uint32_t)x_box;
      const uint32_t a_lo_flags = ((lbo >> 4) & 0x3FFFu) << 16;
      const uint32_t a_hi = ((8u * row_b) >> 4) | (1u << 14) | (layout << 29);
      const uint32_t b_lo_flags = (((uint32_t)W2_BOX >> 4) & 0x3FFFu) << 16;
      const uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_k16 = (16u * row_b) >> 4;                                   // 16 rows further down
      const uint32_t a_tb16 = p.haloed ? ((uint32_t)(p.tm * p.dil) * row_b) >> 4 : (uint32_t)(2 * x_box) >> 4;
      for (int it = 0; it < my_units; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t s_addr = smem_u32(smem + (size_t)s * stage_bytes);
          // one accumulator per tap block (tm taps; tm = 1: per tap).  haloed: block tb = the tile moved down by tb*tm*dil
          // rows and, inside the instruction, M atom a = the tile moved down by a further a*dil rows (LBO = dil rows).
          // The single issuing thread is the bottleneck of these small MMAs (a descriptor built from scratch costs ~100
          // cycles of dependent integer work per instruction): the high descriptor words are constants and the low
          // words advance by 32-bit adds.
          const uint32_t sa16 = (s_addr & 0x3FFFFu) >> 4;
          const uint32_t sb16 = sa16 + (uint32_t)(x_bytes >> 4);
          uint32_t a_blk = a_lo_flags + sa16;
          for (int tb = 0; tb < nacc; ++tb) {
            const uint32_t d_addr = tmem_base + (uint32_t)(tb * p.NT);
            uint32_t a_lo = a_blk, b_lo = b_lo_flags + sb16;
#pragma unroll
            for (int k = 0; k < W2_TK / 16; ++k) {      // 16 time rows per MMA
              umma_bf16_lohi(d_addr, a_lo, a_hi, b_lo, b_hi, idesc, (it > 0 || k > 0) ? 1u : 0u);
              a_lo += a_k16;
              b_lo += 2048u >> 4;
            }
            a_blk += a_tb16;
          }
          if (do_bias) {
            const uint32_t o16 = (smem_u32(ones) & 0x3FFFFu) >> 4;
            const uint32_t d_addr = tmem_base + (uint32_t)(p.nacc_max * p.NT);
#pragma unroll
            for (int k = 0; k < W2_TK / 16; ++k)
              umma_bf16_lohi(d_addr, b_lo_flags + o16 + (uint32_t)k * 128u, b_hi, b_lo_flags + sb16 + (uint32_t)k * 128u, b_hi, idesc,
                             (it > 0 || k > 0) ? 1u : 0u);
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
      const int nvalid = min(p.NT, p.Cout - n0);
      float* wsg = p.ws + (long long)g * p.ws_grp_stride;
      if (p.tm > 1) {
        // lane = (tap within the block, input channel)
        const int L = q * 32 + lane;
        const int a = L / p.Cin, ci = L - a * p.Cin;
        for (int tb = 0; tb < nacc; ++tb) {
          const int tap = tap0 + tb * p.tm + a;
          const bool row_ok = tap < tap0 + ntaps;
          float* wrow = wsg + ((long long)tap * p.Np + n0) * p.Mp + ci;
          for (int c0 = 0; c0 < nvalid; c0 += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tb * p.NT + c0), v);
            if (row_ok) {
              const int nj = min(16, nvalid - c0);
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < nj) atomicAdd(wrow + (long long)(c0 + j) * p.Mp, v[j]);
            }
          }
        }
      } else if (ci0 + q * 32 < p.Cin) {         // lanes of warps whose 32 ci are all padding have nothing to add
        const int ci = ci0 + q * 32 + lane;
        const bool row_ok = ci < p.Cin;
        for (int tp = 0; tp < ntaps; ++tp) {
          float* wrow = wsg + ((long long)(tap0 + tp) * p.Np + n0) * p.Mp + ci;
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
      }
      if (do_bias && q == 0) {
        // every lane of accumulator KT holds the column sums; lane j of the warp adds column c0 + j
        float* brow = wsg + (long long)p.K * p.Np * p.Mp + n0;
        for (int c0 = 0; c0 < nvalid; c0 += 16) {
          float v[16];
          tmem_ld16(tmem_base + (uint32_t)(p.nacc_max * p.NT + c0), v);
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

// The same GEMM with the operand roles swapped, for convolutions whose INPUT width is just over one 128-lane M tile
// (the FiLM conditioning convs: 136 / 137 channels -- a second M tile would carry 8 or 9 real channels and double the
// MMAs): M = co (dy tile, plain), N = ci <= 256 (haloed x tile; the taps are row-shifted views of the B operand), one
// accumulator of N columns per tap, bias gradient = one more 16-column accumulator against a tile of ones on the B side.
// Workspace layout [g][tap][ci][co] (a TMEM lane is a co: a warp's reduction covers 32 consecutive floats).
struct Wg2sP {
  int B, Tout, Cout, Cin, K, dil, t_off;
  int ngroups, x_ch_off, x_ch_stride, dy_ch_off, dy_ch_stride;
  int NT, nb_x, rows_x, stages, tmem_cols, nchunk_t, units, splits;
  int CinP, CoutP;
  long long ws_grp_stride;
  float* ws;
  int bias;
};

__global__ void __launch_bounds__(W2_THREADS) conv_tc_wgrad2s_k(const __grid_constant__ CUtensorMap map_x,
                                                                const __grid_constant__ CUtensorMap map_dy, Wg2sP p) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int x_box = p.rows_x * 128;
  const int x_bytes = p.nb_x * x_box;
  const int dy_bytes = 2 * W2_BOX;                                   // 128 co = two 64-channel boxes of 64 rows
  const int stage_bytes = dy_bytes + x_bytes;
  uint8_t* ones = smem + (size_t)p.stages * stage_bytes;             // 64 rows x 64 channels of 1.0, when p.bias
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ones + (p.bias ? W2_BOX : 0));
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.z;
  const int co0 = blockIdx.y * 128;
  const int split = blockIdx.x;
  const int my_units = (p.units - split + p.splits - 1) / p.splits;
  const int xc = p.x_ch_off + g * p.x_ch_stride;
  const int dc = p.dy_ch_off + g * p.dy_ch_stride + co0;
  const int nba = min(2, (p.Cout - co0 + 63) / 64);
  if (p.bias) {
    uint32_t* o32 = reinterpret_cast<uint32_t*>(ones);
    for (int i = threadIdx.x; i < W2_BOX / 4; i += W2_THREADS) o32[i] = 0x3F803F80u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
        const uint32_t tx = (uint32_t)(nba * W2_BOX + x_bytes);
        for (int it = 0; it < my_units; ++it) {
          const int u = split + it * p.splits;
          const int b = u / p.nchunk_t, tc = (u - b * p.nchunk_t) * W2_TK;
          const int s = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          mbar_expect_tx(&full_bar[s], tx);
          for (int j = 0; j < nba; ++j) tma_load_3d(sa + j * W2_BOX, &map_dy, &full_bar[s], dc + 64 * j, tc, b);
          for (int j = 0; j < p.nb_x; ++j) tma_load_3d(sa + dy_bytes + j * x_box, &map_x, &full_bar[s], xc + 64 * j, tc + p.t_off, b);
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc_b = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);                      // SBO = 8 rows, SWIZZLE_128B
      const uint32_t a_flags = (((uint32_t)W2_BOX >> 4) & 0x3FFFu) << 16;             // LBO: next 64 co
      const uint32_t b_flags = (((uint32_t)x_box >> 4) & 0x3FFFu) << 16;              // LBO: next 64 ci
      const uint32_t tap16 = ((uint32_t)p.dil * 128u) >> 4;
      for (int it = 0; it < my_units; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa16 = (smem_u32(smem + (size_t)s * stage_bytes) & 0x3FFFFu) >> 4;
          const uint32_t sb16 = sa16 + (uint32_t)(dy_bytes >> 4);
          uint32_t b_tap = b_flags + sb16;
          for (int tap = 0; tap < p.K; ++tap) {
            const uint32_t d_addr = tmem_base + (uint32_t)(tap * p.NT);
            uint32_t a_lo = a_flags + sa16, b_lo = b_tap;
#pragma unroll
            for (int k = 0; k < W2_TK / 16; ++k) {
              umma_bf16_lohi(d_addr, a_lo, hi, b_lo, hi, idesc, (it > 0 || k > 0) ? 1u : 0u);
              a_lo += 128u;
              b_lo += 128u;
            }
            b_tap += tap16;
          }
          if (p.bias) {
            const uint32_t o16 = (smem_u32(ones) & 0x3FFFFu) >> 4;
            const uint32_t d_addr = tmem_base + (uint32_t)(p.K * p.NT);
#pragma unroll
            for (int k = 0; k < W2_TK / 16; ++k)
              umma_bf16_lohi(d_addr, a_flags + sa16 + (uint32_t)k * 128u, hi, a_flags + o16 + (uint32_t)k * 128u, hi, idesc_b,
                             (it > 0 || k > 0) ? 1u : 0u);
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
      const int co = co0 + q * 32 + lane;
      const bool row_ok = co < p.Cout;
      float* wsg = p.ws + (long long)g * p.ws_grp_stride;
      if (co0 + q * 32 < p.Cout) {
        for (int tap = 0; tap < p.K; ++tap) {
          float* wrow = wsg + (long long)tap * p.CinP * p.CoutP + co;
          for (int c0 = 0; c0 < p.Cin; c0 += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tap * p.NT + c0), v);
            if (row_ok) {
              const int nj = min(16, p.Cin - c0);
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < nj) atomicAdd(wrow + (long long)(c0 + j) * p.CoutP, v[j]);
            }
          }
        }
        if (p.bias) {
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.K * p.NT), v);
          if (row_ok) atomicAdd(wsg + (long long)p.K * p.CinP * p.CoutP + co, v[0]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// swapped workspace [g][tap][ci][co] (+ bias row [co]) -> dw[g][co][ci][tap]; re-zeroes what it read
__global__ void wgrad2s_finalize_k(float* __restrict__ ws, long long ws_grp_stride, int Cout, int Cin, int K, int CinP, int CoutP,
                                   float* __restrict__ dw, long long dw_grp_stride, float* __restrict__ db, long long db_grp_stride,
                                   int bias) {
  pdl_prologue();
  const int g = blockIdx.y;
  float* wsg = ws + (long long)g * ws_grp_stride;
  const long long n = (long long)K * Cin * Cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long long r = i / Cout;
    const int ci = (int)(r % Cin), tap = (int)(r / Cin);
    float* src = wsg + ((long long)tap * CinP + ci) * CoutP + co;
    const float v = *src;
    *src = 0.f;
    if (dw) dw[(long long)g * dw_grp_stride + ((long long)co * Cin + ci) * K + tap] = v;
  }
  if (bias && blockIdx.x == 0) {
    float* brow = wsg + (long long)K * CinP * CoutP;
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) {
      if (db) db[(long long)g * db_grp_stride + i] = brow[i];
      brow[i] = 0.f;
    }
  }
}

// ws[g][tap][co][ci] (+ bias row) -> the weight-gradient tensors; what is read is written back as zero, so a persistent
// workspace needs no memset per call.
struct Fin2P {
  float* ws;
  long long ws_grp_stride;
  int ngroups, Np, Mp, Cout, Cin, K, bias, per_group;
  float* dw[W2_MAXG];        // plain mode: dw of group g = [Cout][Cin][kg[g]]  (per_group) or dw[0] + g * dw_grp_stride
  float* db[W2_MAXG];
  int kg[W2_MAXG];
  long long dw_grp_stride, db_grp_stride;
  // frame mode (frame_s > 0): the groups are bundles of `sub` conv groups of a Conv1d(k = kreal, stride = frame_s) seen as a
  // stride-1 conv over frames; ci = (local conv group, input channel c, phase p), tap = frame j, k = j * frame_s + p;
  // dw[0] is the conv's own [Cout_total][cin_conv_g][kreal] gradient, db[0] its bias gradient
  int frame_s, kreal, cin_conv_g, sub;
};

__global__ void wgrad2_finalize_k(Fin2P f) {
  pdl_prologue();
  const int g = blockIdx.y;
  float* wsg = f.ws + (long long)g * f.ws_grp_stride;
  const int gi = f.per_group ? g : 0;
  const int kg = f.per_group ? f.kg[gi] : f.K;
  const long long n = (long long)kg * f.Cout * f.Cin;
  float* dwg = f.per_group ? f.dw[gi] : (f.dw[0] ? f.dw[0] + (long long)g * f.dw_grp_stride : nullptr);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % f.Cin);
    const long long r = i / f.Cin;
    const int co = (int)(r % f.Cout), tap = (int)(r / f.Cout);
    float* src = wsg + ((long long)tap * f.Np + co) * f.Mp + ci;
    const float v = *src;
    *src = 0.f;
    if (f.frame_s > 0) {
      const int cin_b = f.Cin / f.sub, cout_b = f.Cout / f.sub;      // per conv group inside the bundle
      if (co / cout_b != ci / cin_b) continue;                      // off-diagonal block of the bundle: not a weight
      const int cl = ci % cin_b;
      const int c = cl / f.frame_s, ph = cl % f.frame_s;
      const int k = tap * f.frame_s + ph;
      if (k < f.kreal && f.dw[0])
        f.dw[0][(((long long)g * f.Cout + co) * f.cin_conv_g + c) * f.kreal + k] = v;
    } else if (dwg) {
      dwg[((long long)co * f.Cin + ci) * kg + tap] = v;
    }
  }
  if (f.bias && blockIdx.x == 0) {
    float* brow = wsg + (long long)f.K * f.Np * f.Mp;
    float* dbg = f.frame_s > 0 ? (f.db[0] ? f.db[0] + (long long)g * f.Cout : nullptr)
                               : (f.per_group ? f.db[gi] : (f.db[0] ? f.db[0] + (long long)g * f.db_grp_stride : nullptr));
    for (int i = threadIdx.x; i < f.Cout; i += blockDim.x) {
      if (dbg) dbg[i] = brow[i];
      brow[i] = 0.f;
    }
  }
}


// The same for wide layers (Cin >= 128, plain weight layout): a CTA takes one output channel and 256 input channels, reads
// their kg taps from the workspace (consecutive ci: coalesced), transposes in shared memory and writes the 256 * kg
// consecutive floats of dw[co][ci0 ..][0 .. kg) -- the element-per-thread kernel above writes with a stride of kg floats
// (207 us under ncu for the discriminator's 1024 x 1024 x 5 layer, three times per step).
constexpr int FIN_CI = 256, FIN_KMAX = 16;
__global__ void __launch_bounds__(256) wgrad2_finalize_wide_k(Fin2P f) {
  pdl_prologue();
  __shared__ float tile[FIN_CI * FIN_KMAX];
  const int g = blockIdx.z;
  float* wsg = f.ws + (long long)g * f.ws_grp_stride;
  const int gi = f.per_group ? g : 0;
  const int kg = f.per_group ? f.kg[gi] : f.K;
  float* dwg = f.per_group ? f.dw[gi] : (f.dw[0] ? f.dw[0] + (long long)g * f.dw_grp_stride : nullptr);
  const int co = blockIdx.y, ci0 = blockIdx.x * FIN_CI;
  const int nci = min(FIN_CI, f.Cin - ci0);
  if (threadIdx.x < nci) {
    for (int tap = 0; tap < kg; ++tap) {
      float* src = wsg + ((long long)tap * f.Np + co) * f.Mp + ci0 + threadIdx.x;
      tile[threadIdx.x * kg + tap] = *src;
      *src = 0.f;
    }
  }
  __syncthreads();
  if (dwg) {
    float* dst = dwg + ((long long)co * f.Cin + ci0) * kg;
    for (int i = threadIdx.x; i < nci * kg; i += 256) dst[i] = tile[i];
  }
  if (f.bias && blockIdx.x == 0 && blockIdx.y == 0) {
    float* brow = wsg + (long long)f.K * f.Np * f.Mp;
    float* dbg = f.per_group ? f.db[gi] : (f.db[0] ? f.db[0] + (long long)g * f.db_grp_stride : nullptr);
    for (int i = threadIdx.x; i < f.Cout; i += blockDim.x) {
      if (dbg) dbg[i] = brow[i];
      brow[i] = 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------ frame view (channel-major)
// x[B, C, T] fp32 NCW -> xf[B, Tq, C*s] bf16 channels-last frames, xf[b, q, c*s + p] = x[b, c, s*q + p - pad] (0 outside the
// signal): the operand of a Conv1d(k, stride = s, groups) run as a stride-1 grouped convolution over frames -- in this
// order the s*cin_g frame channels of a conv group are contiguous.  A block moves 32 channels x 32 frames.
template <int S>
__global__ void __launch_bounds__(256) frame_pack_k(const float* __restrict__ x, __nv_bfloat16* __restrict__ xf, int C, int T,
                                                    int pad, int Tq) {
  pdl_prologue();
  __shared__ float tile[32][32 * S + 1];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, q0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int u0 = q0 * S - pad;
  for (int cy = wrp; cy < 32; cy += 8) {
    const int c = c0 + cy;
    const float* row = x + ((long long)b * C + c) * T;
#pragma unroll
    for (int l = 0; l < S; ++l) {
      const int u = u0 + lane + 32 * l;
      tile[cy][lane + 32 * l] = (c < C && u >= 0 && u < T) ? __ldg(row + u) : 0.f;
    }
  }
  __syncthreads();
  const int c = c0 + lane;
  for (int qy = wrp; qy < 32; qy += 8) {
    const int q = q0 + qy;
    if (q >= Tq || c >= C) continue;
    __nv_bfloat16* dst = xf + ((long long)b * Tq + q) * ((long long)C * S) + (long long)c * S;
    if (S == 4) {
      __nv_bfloat162 a = __floats2bfloat162_rn(tile[lane][qy * 4], tile[lane][qy * 4 + 1]);
      __nv_bfloat162 d = __floats2bfloat162_rn(tile[lane][qy * 4 + 2], tile[lane][qy * 4 + 3]);
      uint2 w;
      w.x = *reinterpret_cast<uint32_t*>(&a);
      w.y = *reinterpret_cast<uint32_t*>(&d);
      *reinterpret_cast<uint2*>(dst) = w;
    } else if (S == 2) {
      *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(tile[lane][qy * 2], tile[lane][qy * 2 + 1]);
    } else {
#pragma unroll
      for (int pp = 0; pp < S; ++pp) dst[pp] = __float2bfloat16(tile[lane][qy * S + pp]);
    }
  }
}

// the inverse view for gradients: dx[b, c, u] = dxf[b, c*s + p, q] with s*q + p = u + pad; dxf is fp32 NCW over frames
// [B, C*s, Tq] (what the data-gradient conv writes)
__global__ void frame_unpack_k(const float* __restrict__ dxf, float* __restrict__ dx, int C, int T, int s, int pad, int Tq) {
  pdl_prologue();
  const int b = blockIdx.z, c = blockIdx.y;
  const float* src = dxf + ((long long)b * C + c) * s * Tq;
  float* dst = dx + ((long long)b * C + c) * T;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < T; u += gridDim.x * blockDim.x) {
    const int v = u + pad;
    const int q = v / s, ph = v - q * s;
    dst[u] = q < Tq ? __ldg(src + (long long)ph * Tq + q) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------ operand packs, one launch each
// The grouped operands of one MRF depth (G kernel-size branches, model/generator.py:175-194) from the branches' fp32
// weights w_g[C][C][k_g]: wp[Kmax][G*C][C] (forward: row = (branch, co), column = ci, branch g's taps centred inside Kmax,
// zeros around them), wtp[Kmax][G*C][C] (data gradient: row = (branch, ci), column = co, taps reversed) and the
// concatenated bias[G*C] -- instead of 3 fills + 6 per-weight pack launches + 3 bias copies.
struct ChainPackP {
  const float* w[4];
  const float* bias[4];
  int k[4];
  int G, C, Kmax;
  __nv_bfloat16* wp;
  __nv_bfloat16* wtp;
  float* bias_out;
};

__global__ void chain_pack_k(const __grid_constant__ ChainPackP p) {
  pdl_prologue();
  const long long n = (long long)p.Kmax * p.G * p.C * p.C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i % p.C);
    long long r = i / p.C;
    const int a = (int)(r % p.C); r /= p.C;
    const int g = (int)(r % p.G);
    const int tap = (int)(r / p.G);
    const int kg = p.k[g], lo = (p.Kmax - kg) >> 1, t = tap - lo;
    float f = 0.f, ft = 0.f;
    if (t >= 0 && t < kg) {
      f = p.w[g][((long long)a * p.C + b) * kg + t];                       // W[co = a][ci = b][t]
      ft = p.w[g][((long long)b * p.C + a) * kg + (kg - 1 - t)];           // W[co = b][ci = a][k - 1 - t]
    }
    p.wp[i] = __float2bfloat16(f);
    p.wtp[i] = __float2bfloat16(ft);
  }
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < p.G * p.C; i += blockDim.x) {
      const int g = i / p.C;
      p.bias_out[i] = p.bias[g] ? p.bias[g][i - g * p.C] : 0.f;
    }
}

// The bundled frame-view operands of a grouped strided conv (model/discriminator.py:26-30) from its weight
// w[Cout][cin_g][K]: wp[m][Cout][cin_b] -- frame tap j, output channel co, bundle-local frame channel (local conv group,
// c, phase p) <- w[co][c][s*j + p] on the diagonal blocks, zero elsewhere and for the appended taps -- and
// wtp[m][nb*cin_b][cout_b] = the same transposed inside each bundle with the taps reversed (data gradient).
__global__ void frame_weights_pack_k(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, __nv_bfloat16* __restrict__ wtp,
                                     int Cout, int cin_g, int K, int s, int m, int sub, int cin_b, int cout_b) {
  pdl_prologue();
  const int fpg = cin_g * s, cout_g = cout_b / sub, nb = Cout / cout_b;
  const long long n = (long long)m * Cout * cin_b;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int cib = (int)(i % cin_b);
    long long r = i / cin_b;
    const int co = (int)(r % Cout);
    const int j = (int)(r / Cout);
    const int sub_in = cib / fpg, local = cib - sub_in * fpg;
    const int c = local / s, ph = local - c * s;
    const int k = j * s + ph;
    const int sub_out = (co / cout_g) % sub;
    const float f = (sub_in == sub_out && k < K) ? w[((long long)co * cin_g + c) * K + k] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16(f);
    wp[i] = h;
    const int bundle = co / cout_b, cob = co - bundle * cout_b;
    wtp[((long long)(m - 1 - j) * nb * cin_b + (long long)bundle * cin_b + cib) * cout_b + cob] = h;
  }
}

}  // namespace tdvc
using namespace tdvc;

extern "C" int tdvc_chain_pack(const float* const* w, const float* const* bias, const int* k, int G, int C, int Kmax, void* wp,
                               void* wtp, float* bias_out, void* stream) {
  TDVC_CHECK_ARG(w && k && G >= 1 && G <= 4 && C > 0 && Kmax >= 1 && wp && wtp && bias_out);
  ChainPackP p{};
  for (int g = 0; g < G; ++g) {
    TDVC_CHECK_ARG(w[g] && k[g] >= 1 && k[g] <= Kmax && (Kmax - k[g]) % 2 == 0);
    p.w[g] = w[g]; p.bias[g] = bias ? bias[g] : nullptr; p.k[g] = k[g];
  }
  p.G = G; p.C = C; p.Kmax = Kmax; p.wp = (__nv_bfloat16*)wp; p.wtp = (__nv_bfloat16*)wtp; p.bias_out = bias_out;
  const long long n = (long long)Kmax * G * C * C;
  const int blocks = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 4LL * num_sms()));
  tdvc::launch_k(chain_pack_k, blocks, 256, 0, (cudaStream_t)stream, p);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_frame_weights_pack(const float* w, void* wp, void* wtp, int Cout, int cin_g, int K, int s, int m, int sub,
                                       int cin_b, int cout_b, void* stream) {
  TDVC_CHECK_ARG(w && wp && wtp && Cout > 0 && cin_g > 0 && K > 0 && s >= 1 && m >= 1 && m * s >= K && sub >= 1);
  TDVC_CHECK_ARG(cin_b == sub * cin_g * s && cout_b % sub == 0 && Cout % cout_b == 0);
  const long long n = (long long)m * Cout * cin_b;
  const int blocks = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 4LL * num_sms()));
  tdvc::launch_k(frame_weights_pack_k, blocks, 256, 0, (cudaStream_t)stream, w, (__nv_bfloat16*)wp, (__nv_bfloat16*)wtp, Cout, cin_g,
                 K, s, m, sub, cin_b, cout_b);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_frame_pack_bf16(const float* x, void* xf, int B, int C, int T, int s, int pad, int Tq, void* stream) {
  TDVC_CHECK_ARG(x && xf && B >= 0 && C > 0 && T > 0 && pad >= 0 && Tq > 0 && (s == 1 || s == 2 || s == 4 || s == 8));
  TDVC_CHECK_ARG(((long long)C * s) % 8 == 0 && ((uintptr_t)xf % 16) == 0);
  if (B == 0) return TDVC_OK;
  dim3 grid(cdiv(Tq, 32), cdiv(C, 32), B);
  TDVC_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* o = (__nv_bfloat16*)xf;
  switch (s) {
    case 1: tdvc::launch_k(frame_pack_k<1>, grid, 256, 0, st, x, o, C, T, pad, Tq); break;
    case 2: tdvc::launch_k(frame_pack_k<2>, grid, 256, 0, st, x, o, C, T, pad, Tq); break;
    case 4: tdvc::launch_k(frame_pack_k<4>, grid, 256, 0, st, x, o, C, T, pad, Tq); break;
    default: tdvc::launch_k(frame_pack_k<8>, grid, 256, 0, st, x, o, C, T, pad, Tq); break;
  }
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_frame_unpack(const float* dxf, float* dx, int B, int C, int T, int s, int pad, int Tq, void* stream) {
  TDVC_CHECK_ARG(dxf && dx && B >= 0 && C > 0 && T > 0 && s >= 1 && pad >= 0 && Tq > 0 && C <= 65535 && B <= 65535);
  if (B == 0) return TDVC_OK;
  tdvc::launch_k(frame_unpack_k, dim3(std::min(cdiv(T, 256), 64), C, B), 256, 0, (cudaStream_t)stream, dxf, dx, C, T, s, pad, Tq);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

// operand roles swapped (M = co, N = ci): the input width is just over one 128-lane tile and fits one N tile with all taps
static bool wgrad2_swapped(const tdvc_tc_wgrad2* c) {
  static int on = -1;        // TDVC_WGRAD2_SWAP=0: never (development / A-B switch)
  if (on < 0) { const char* e = getenv("TDVC_WGRAD2_SWAP"); on = e ? atoi(e) : 1; }
  const int nt = ((c->Cin + 15) / 16) * 16;
  return on && c->swap != 0 && !c->per_group && c->frame_s == 0 && c->Cin > 128 && c->Cin <= 256 && c->K * nt + 16 <= 512 &&
         64 + (c->K - 1) * c->dilation <= 256;
}

extern "C" int64_t tdvc_conv1d_tc_wgrad2_ws(const tdvc_tc_wgrad2* c) {
  if (!c) return 0;
  if (wgrad2_swapped(c)) {
    const long long CinP = ((c->Cin + 15) / 16) * 16, CoutP = (long long)((c->Cout + 31) / 32) * 32;
    return (int64_t)c->ngroups * ((long long)c->K * CinP * CoutP + CoutP);
  }
  const long long Mp = (long long)((c->Cin + 31) / 32) * 32, Np = (long long)((c->Cout + 15) / 16) * 16;
  return (int64_t)c->ngroups * ((long long)c->K * Np * Mp + Np);
}

static int wgrad2_launch_swapped(const tdvc_tc_wgrad2* c, cudaStream_t st) {
  Wg2sP p{};
  p.B = c->B; p.Tout = c->Tout; p.Cout = c->Cout; p.Cin = c->Cin; p.K = c->K; p.dil = c->dilation; p.t_off = c->t_off[0];
  p.ngroups = c->ngroups; p.x_ch_off = c->x_ch_off; p.x_ch_stride = c->x_ch_stride; p.dy_ch_off = c->dy_ch_off;
  p.dy_ch_stride = c->dy_ch_stride;
  p.NT = ((c->Cin + 15) / 16) * 16;
  p.nb_x = cdiv(p.NT, 64);
  p.rows_x = ((W2_TK + (c->K - 1) * c->dilation + 7) / 8) * 8;
  p.CinP = p.NT; p.CoutP = ((c->Cout + 31) / 32) * 32;
  p.ws = c->ws; p.ws_grp_stride = (long long)c->K * p.CinP * p.CoutP + p.CoutP;
  p.bias = c->want_bias ? 1 : 0;
  if (!c->ws_is_zero) TDVC_CUDA(cudaMemsetAsync(c->ws, 0, sizeof(float) * (size_t)c->ngroups * (size_t)p.ws_grp_stride, st));
  int cols = 32;
  while (cols < c->K * p.NT + (p.bias ? 16 : 0)) cols <<= 1;
  p.tmem_cols = cols;
  const long long stage_bytes = 2LL * W2_BOX + (long long)p.nb_x * p.rows_x * 128;
  p.nchunk_t = cdiv(c->Tout, W2_TK);
  p.units = c->B * p.nchunk_t;
  int stages = (int)std::min<long long>(4, (200LL * 1024) / stage_bytes);
  TDVC_CHECK_ARG(stages >= 1);
  p.stages = std::min(stages, std::max(1, p.units));
  const int m_tiles = cdiv(c->Cout, 128);
  if (c->B > 0) {
    const long long ctas_fixed = (long long)m_tiles * c->ngroups;
    int splits = (int)std::max<long long>(1, (2LL * num_sms() + ctas_fixed - 1) / ctas_fixed);
    splits = std::min(splits, std::max(1, p.units / 4));
    p.splits = std::max(1, std::min(splits, p.units));
    const size_t smem = (size_t)p.stages * stage_bytes + (p.bias ? W2_BOX : 0) + (2 * p.stages + 1) * sizeof(uint64_t) + 16 + 1024;
    TDVC_CUDA(cudaFuncSetAttribute(conv_tc_wgrad2s_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUtensorMap map_x, map_dy;
    int rc = make_map_3d(&map_x, c->xp, (uint64_t)c->Cp, (uint64_t)c->Tp, (uint64_t)c->B, 64, (uint32_t)p.rows_x);
    if (rc) return rc;
    rc = make_map_3d(&map_dy, c->dyp, (uint64_t)c->Cdp, (uint64_t)c->Tout, (uint64_t)c->B, 64, W2_TK);
    if (rc) return rc;
    TDVC_CHECK_ARG(m_tiles <= 65535);
    tdvc::launch_k(conv_tc_wgrad2s_k, dim3(p.splits, m_tiles, c->ngroups), W2_THREADS, smem, st, map_x, map_dy, p);
    TDVC_LAUNCH_CHECK();
    g_flops[FLOP_TC_WGRAD] += 2.0 * c->B * c->Tout * (double)c->Cout * c->Cin * c->K * c->ngroups;
  }
  const long long n = (long long)c->K * c->Cout * c->Cin;
  const int bx = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, std::max(1, 4 * num_sms() / c->ngroups)));
  tdvc::launch_k(wgrad2s_finalize_k, dim3(bx, c->ngroups), 256, 0, st, c->ws, p.ws_grp_stride, c->Cout, c->Cin, c->K, p.CinP, p.CoutP,
                 c->dw[0], (long long)c->dw_grp_stride, c->db[0], (long long)c->db_grp_stride, p.bias);
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

extern "C" int tdvc_conv1d_tc_wgrad2(const tdvc_tc_wgrad2* c, void* stream) {
  TDVC_CHECK_ARG(c && c->dyp && c->xp && c->ws && c->B >= 0 && c->Tout > 0 && c->Tp > 0 && c->K > 0 && c->dilation > 0);
  TDVC_CHECK_ARG(c->Cdp % 8 == 0 && c->Cp % 8 == 0 && c->ngroups >= 1 && c->ngroups <= 65535 && c->Cin > 0 && c->Cout > 0);
  TDVC_CHECK_ARG(c->x_ch_off >= 0 && c->dy_ch_off >= 0 && c->x_ch_stride >= 0 && c->dy_ch_stride >= 0);
  TDVC_CHECK_ARG(c->x_ch_off + (long long)(c->ngroups - 1) * c->x_ch_stride + c->Cin <= c->Cp);
  TDVC_CHECK_ARG(c->dy_ch_off + (long long)(c->ngroups - 1) * c->dy_ch_stride + c->Cout <= c->Cdp);
  TDVC_CHECK_ARG(((uintptr_t)c->xp % 16 == 0) && ((uintptr_t)c->dyp % 16 == 0));
  const bool per_group = c->per_group != 0;
  if (per_group) {
    TDVC_CHECK_ARG(c->ngroups <= W2_MAXG);
    for (int g = 0; g < c->ngroups; ++g) TDVC_CHECK_ARG(c->kg[g] >= 1 && c->kg[g] <= c->K);
  }
  if (c->frame_s > 0) {
    TDVC_CHECK_ARG(!per_group && c->sub >= 1 && c->Cin % c->sub == 0 && c->Cout % c->sub == 0 && c->cin_conv_g >= 1 &&
                   (c->Cin / c->sub) == c->cin_conv_g * c->frame_s && c->kreal >= 1 && c->kreal <= c->K * c->frame_s);
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (wgrad2_swapped(c)) return wgrad2_launch_swapped(c, st);
  Wg2P p{};
  p.B = c->B; p.Tout = c->Tout; p.Cout = c->Cout; p.Cin = c->Cin; p.K = c->K; p.dil = c->dilation;
  p.ngroups = c->ngroups; p.per_group = per_group ? 1 : 0;
  p.x_ch_off = c->x_ch_off; p.x_ch_stride = c->x_ch_stride; p.dy_ch_off = c->dy_ch_off; p.dy_ch_stride = c->dy_ch_stride;
  for (int g = 0; g < W2_MAXG; ++g) { p.kg[g] = per_group ? c->kg[g] : c->K; p.toff[g] = per_group ? c->t_off[g] : c->t_off[0]; }
  p.Mp = ((c->Cin + 31) / 32) * 32;
  p.Np = ((c->Cout + 15) / 16) * 16;
  p.ws = c->ws;
  p.ws_grp_stride = (long long)c->K * p.Np * p.Mp + p.Np;
  p.bias = c->want_bias ? 1 : 0;
  const size_t ws_floats = (size_t)c->ngroups * (size_t)p.ws_grp_stride;
  if (!c->ws_is_zero) TDVC_CUDA(cudaMemsetAsync(c->ws, 0, sizeof(float) * ws_floats, st));
  {
    static int h = -1;      // TDVC_WGRAD2_HALOED=0: one shifted copy of the x tile per tap (development / A-B switch)
    if (h < 0) { const char* e = getenv("TDVC_WGRAD2_HALOED"); h = e ? atoi(e) : 1; }
    p.haloed = c->haloed >= 0 ? c->haloed : h;
    static int ki = -1;
    if (ki < 0) { const char* e = getenv("TDVC_WGRAD2_KINNER"); ki = e ? atoi(e) : 1; }
    p.k_inner = ki;
  }
  const int xt = p.bias;
  p.nb_x = std::min(2, cdiv(c->Cin, 64));
  // taps on M: Cin = 16 / 32 / 64 channel rows (SWIZZLE_32B / 64B / 128B), 128 / Cin taps per instruction
  {
    static int tm_on = -1;     // TDVC_WGRAD2_TAPSM=0: channels only on M (development / A-B switch)
    if (tm_on < 0) { const char* e = getenv("TDVC_WGRAD2_TAPSM"); tm_on = e ? atoi(e) : 1; }
    const bool ok = tm_on && p.haloed && c->K > 1 && (c->Cin == 16 || c->Cin == 32 || c->Cin == 64) &&
                    (c->tapsm < 0 || c->tapsm == 1);
    p.tm = (ok && c->tapsm != 0) ? 128 / c->Cin : 1;
    p.sw = p.tm > 1 ? 2 * c->Cin : 128;
  }
  // co tile / taps per CTA: as many taps as possible share the 512 TMEM columns (the x tile is then loaded once for all of them)
  int best_nt = 16, best_kt = 1, best_cost = 1 << 30;
  const long long budget = 200 * 1024;
  auto rows_for = [&](int kt) {     // haloed x rows of one time unit for kt taps (whole tap blocks when tm > 1)
    const int taps = cdiv(kt, p.tm) * p.tm;
    return ((W2_TK + (taps - 1) * c->dilation + 7) / 8) * 8;
  };
  for (int cand = std::min(p.Np, 256); cand >= 16; cand -= 16) {
    int kt = c->K;
    while (kt >= 1) {
      const int rows_x = rows_for(kt);
      const long long xb = p.haloed ? 2LL * rows_x * p.sw : (long long)kt * 2 * W2_BOX;
      if ((cdiv(kt, p.tm) + xt) * cand <= 512 && rows_x <= 256 && 2 * (xb + (long long)cdiv(cand, 64) * W2_BOX) <= budget) break;
      --kt;
    }
    if (kt < 1) continue;
    const int cost = cdiv(c->K, kt) * cdiv(p.Np, cand);
    if (cost < best_cost) { best_cost = cost; best_nt = cand; best_kt = kt; }
  }
  TDVC_CHECK_ARG(best_cost < (1 << 30));
  p.NT = best_nt; p.KT = best_kt;
  if (p.tm > 1 && p.KT < c->K) p.KT = std::max(p.tm, p.KT / p.tm * p.tm);      // tap groups start at whole blocks
  p.ntap_groups = cdiv(c->K, p.KT);
  p.n_ntiles = cdiv(p.Np, p.NT);
  p.nb_dy = cdiv(p.NT, 64);
  p.rows_x = rows_for(p.KT);
  p.nacc_max = cdiv(p.KT, p.tm);
  int cols = 32;
  while (cols < (p.nacc_max + xt) * p.NT) cols <<= 1;
  TDVC_CHECK_ARG(cols <= 512);
  p.tmem_cols = cols;
  const long long x_bytes = p.haloed ? 2LL * p.rows_x * p.sw : (long long)p.KT * 2 * W2_BOX;
  const long long stage_bytes = x_bytes + (long long)p.nb_dy * W2_BOX;
  p.nchunk_t = cdiv(c->Tout, W2_TK);
  p.units = c->B * p.nchunk_t;
  int stages = (int)std::min<long long>(4, budget / stage_bytes);
  TDVC_CHECK_ARG(stages >= 1);
  stages = std::min(stages, std::max(1, p.units));
  p.stages = stages;
  const int m_tiles = cdiv(c->Cin, 128);
  const int gy = m_tiles * p.ntap_groups * p.n_ntiles;
  TDVC_CHECK_ARG(gy <= 65535);
  if (c->B > 0) {
    // split the time units so that the launch fills the machine about twice; one CTA reduces >= 4 units when it can
    long long ctas_fixed = (long long)gy * c->ngroups;
    int splits = (int)std::max<long long>(1, (2LL * num_sms() + ctas_fixed - 1) / ctas_fixed);
    splits = std::min(splits, std::max(1, p.units / 4));
    splits = std::min(splits, p.units);
    p.splits = std::max(1, splits);
    const size_t smem = (size_t)stages * stage_bytes + (p.bias ? 2 * W2_BOX : 0) + (2 * stages + 1) * sizeof(uint64_t) + 16 + 1024;
    TDVC_CUDA(cudaFuncSetAttribute(conv_tc_wgrad2_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUtensorMap map_x, map_dy;
    const CUtensorMapSwizzle xsw = p.sw == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.sw == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    int rc = make_map_3d(&map_x, c->xp, (uint64_t)c->Cp, (uint64_t)c->Tp, (uint64_t)c->B, (uint32_t)(p.sw / 2),
                         (uint32_t)(p.haloed ? p.rows_x : W2_TK), xsw);
    if (rc) return rc;
    rc = make_map_3d(&map_dy, c->dyp, (uint64_t)c->Cdp, (uint64_t)c->Tout, (uint64_t)c->B, 64, W2_TK);
    if (rc) return rc;
    tdvc::launch_k(conv_tc_wgrad2_k, dim3(p.splits, gy, c->ngroups), W2_THREADS, smem, st, map_x, map_dy, p);
    TDVC_LAUNCH_CHECK();
    double taps = 0;
    for (int g = 0; g < c->ngroups; ++g) taps += per_group ? c->kg[g] : c->K;
    g_flops[FLOP_TC_WGRAD] += 2.0 * c->B * c->Tout * (double)c->Cout * c->Cin * taps;
  }
  Fin2P f{};
  f.ws = c->ws; f.ws_grp_stride = p.ws_grp_stride; f.ngroups = c->ngroups; f.Np = p.Np; f.Mp = p.Mp;
  f.Cout = c->Cout; f.Cin = c->Cin; f.K = c->K; f.bias = p.bias; f.per_group = p.per_group;
  for (int g = 0; g < W2_MAXG; ++g) { f.dw[g] = c->dw[g]; f.db[g] = c->db[g]; f.kg[g] = p.kg[g]; }
  f.dw_grp_stride = c->dw_grp_stride; f.db_grp_stride = c->db_grp_stride;
  f.frame_s = c->frame_s; f.kreal = c->kreal; f.cin_conv_g = c->cin_conv_g; f.sub = c->sub > 0 ? c->sub : 1;
  const long long n = (long long)c->K * c->Cout * c->Cin;
  int kmax = c->K;
  for (int g = 0; g < W2_MAXG; ++g) kmax = std::max(kmax, p.kg[g]);
  if (c->frame_s == 0 && c->Cin >= 128 && kmax <= FIN_KMAX && c->Cout <= 65535) {
    tdvc::launch_k(wgrad2_finalize_wide_k, dim3(cdiv(c->Cin, FIN_CI), c->Cout, c->ngroups), 256, 0, st, f);
  } else {
    const int bx = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, std::max(1, 4 * num_sms() / c->ngroups)));
    tdvc::launch_k(wgrad2_finalize_k, dim3(bx, c->ngroups), 256, 0, st, f);
  }
  TDVC_LAUNCH_CHECK();
  return TDVC_OK;
}

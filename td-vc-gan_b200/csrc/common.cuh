// Shared helpers for libtdvc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/tdvc_b200.h"

namespace tdvc {

void set_error(const char* fmt, ...);

#define TDVC_CHECK_ARG(cond)                                                        \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      tdvc::set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond);        \
      return TDVC_ERR_ARG;                                                          \
    }                                                                               \
  } while (0)

#define TDVC_CUDA(expr)                                                             \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      tdvc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return TDVC_ERR_CUDA;                                                         \
    }                                                                               \
  } while (0)

// every kernel launch in the library is followed by this macro: it also feeds tdvc_launch_count()
extern unsigned long long g_launches;
// multiply-accumulate work handed to each kernel family (2 * MACs, channel padding of the packed operands included):
// bench.py divides the per-step deltas by the families' measured time.  See tdvc_flop_count().
enum FlopFamily { FLOP_TC_TILE = 0, FLOP_TC_WS = 1, FLOP_TC_WT = 2, FLOP_TC_WGRAD = 3, FLOP_FP32 = 4, FLOP_TC_CHAIN = 5, FLOP_FAMILIES = 8 };
extern double g_flops[FLOP_FAMILIES];
#define TDVC_LAUNCH_CHECK()            \
  do {                                 \
    ++tdvc::g_launches;                \
    TDVC_CUDA(cudaGetLastError());     \
  } while (0)

// Programmatic dependent launch.  The step is ~7000 short kernels replayed from a CUDA graph; between two dependent kernel
// nodes the GPU otherwise idles for the whole launch latency.  Every kernel of the library is launched with
// programmatic stream serialization and starts with pdl_prologue(): griddepcontrol.wait blocks until the preceding
// grid has completed and its writes are visible (so nothing below it can observe stale data), and launch_dependents
// then lets the NEXT kernel in the stream be scheduled -- its blocks take free SM slots and sit in their own
// griddepcontrol.wait -- while this one runs.  TDVC_PDL=0 launches plainly (the prologue is then a no-op).
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool pdl_enabled();
// The tensor-core kernels run the part of their prologue that touches no global memory -- mbarrier initialisation, tensor-map
// prefetch, TMEM allocation, shared-memory constants -- BEFORE griddepcontrol.wait, so it overlaps the tail of the preceding
// kernel instead of following it; everything that reads global memory (bias, TMA loads, epilogue operands) stays behind the
// wait.  launch_dependents is issued together with the wait as before: a dependent grid starts only when every CTA of this
// one is resident, so blocks that sit in the wait (holding shared memory / TMEM columns) cannot starve their predecessor.
// -DTDVC_PDL_LATE_WAIT=0 puts the wait back at the top of the kernel.
#ifndef TDVC_PDL_LATE_WAIT
#define TDVC_PDL_LATE_WAIT 1
#endif
__device__ __forceinline__ void pdl_prologue_top() {
  if (!TDVC_PDL_LATE_WAIT) pdl_prologue();
}
__device__ __forceinline__ void pdl_prologue_late() {
  if (TDVC_PDL_LATE_WAIT) pdl_prologue();
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<Args&&>(args)...);
}

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; every thread gets the result.  `sm` must hold >= 33 floats.
__device__ __forceinline__ float block_sum(float v, float* sm) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? sm[threadIdx.x] : 0.f;
  if (wid == 0) {
    r = warp_sum(r);
    if (lane == 0) sm[32] = r;
  }
  __syncthreads();
  return sm[32];
}

// number of SMs of the current device (cached)
int num_sms();

}  // namespace tdvc

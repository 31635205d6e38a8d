// binsort.cuh -- one-CTA stable LSD radix pass over an L2-resident segment of 64-bit keys, shared by the
// per-bin sort of the bucketed binning (bucket.cu) and the long-tie-run repair of the 32-bit-key global
// sort (project.cu).
#pragma once
#include "common.cuh"

namespace b200gs {

constexpr int BS_THREADS = 512;
constexpr int BS_WARPS = BS_THREADS / 32;
constexpr uint32_t BS_TIE_INSERTION_MAX = 24;

struct BinSortShared {
  uint32_t hist[BS_WARPS][256];
  uint32_t digit_base[256];
  uint32_t wsum[8];
  uint32_t long_run;
};

// One stable counting-sort pass of src[0,n) into dst[0,n) on the 8-bit digit at `shift`.  Warp w owns
// the contiguous slice [w*per, (w+1)*per): ranks inside a 32-element group come from __match_any_sync,
// ranks across groups from the warp's running shared-memory histogram, ranks across warps and digits
// from the scan in the middle.  Keys are re-read (L2) for the scatter instead of being kept in
// registers, so any segment length works.
__device__ __forceinline__ void bin_sort_pass(BinSortShared& sh, const uint64_t* __restrict__ src,
                                              uint64_t* __restrict__ dst, uint32_t n, int shift) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t per = (((n + BS_WARPS - 1) / BS_WARPS) + 31u) & ~31u;
  const uint32_t wbeg = min(n, warp * per), wend = min(n, wbeg + per);
  for (int i = tid; i < BS_WARPS * 256; i += BS_THREADS) (&sh.hist[0][0])[i] = 0u;
  __syncthreads();
  for (uint32_t i0 = wbeg; i0 < wend; i0 += 32) {
    const uint32_t i = i0 + lane;
    const bool ok = i < wend;
    const uint32_t d = ok ? (uint32_t)(__ldcg(src + i) >> shift) & 255u : 0xFFFFFFFFu;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (ok && lane == __ffs(peers) - 1) sh.hist[warp][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  if (tid < 256) {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < BS_WARPS; w++) {
      const uint32_t c = sh.hist[w][tid];
      sh.hist[w][tid] = run;   // items of this digit in lower warps
      run += c;
    }
    uint32_t incl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) sh.wsum[warp] = incl;
    sh.digit_base[tid] = incl - run;   // exclusive within the warp of 32 digits; warp prefix added below
  }
  __syncthreads();
  if (tid < 256) {
    uint32_t wp = 0;
    for (int w = 0; w < warp; w++) wp += sh.wsum[w];
    sh.digit_base[tid] += wp;
  }
  __syncthreads();
  const uint32_t lt = (1u << lane) - 1u;
  for (uint32_t i0 = wbeg; i0 < wend; i0 += 32) {
    const uint32_t i = i0 + lane;
    const bool ok = i < wend;
    const uint64_t key = ok ? __ldcg(src + i) : 0ull;
    const uint32_t d = ok ? (uint32_t)(key >> shift) & 255u : 0xFFFFFFFFu;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t off = 0;
    if (ok) off = sh.hist[warp][d];
    __syncwarp();
    if (ok && lane == __ffs(peers) - 1) sh.hist[warp][d] = off + __popc(peers);
    __syncwarp();
    if (ok) dst[sh.digit_base[d] + off + __popc(peers & lt)] = key;
  }
  __syncthreads();   // dst (global) is complete and visible to the whole CTA
}


}  // namespace b200gs

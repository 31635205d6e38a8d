// bucket.cu -- bucketed binning: per-bin pair lists without a global sort.
//
// Replaces rows a4 InclusiveSum, a5 duplicateWithKeys, a6 SortPairs and
// a7 identifyTileRanges of the public algorithm -- same result (every bin's Gaussian ids in
// (depth bits, index) order, which is the order a stable radix sort of (bin << 32 | depth) keys
// emitted in index order produces), different route.  The global-sort pipeline (binning.cu,
// coopsort.cu) is latency-bound on B200: 0.8 M pairs are 9 MB, yet scan + emit + 5 onesweep passes +
// ranges take 0.17 ms of a 0.39 ms C3 frame in 16 dependent launches whose decoupled look-back chains
// cannot be hidden.  Here:
//
//   k_project      counts pairs per bin while it computes the spans (one RED per pair; counters
//                  sit 256 B apart so the L2 atomic units never serialise two bins);
//   k_bin_scan     one CTA: exclusive scan of the <= few thousand bin counts -> bin_base[], the
//                  per-bin [start,end) ranges the compositing kernels read, and D;
//   k_emit_bucket  appends (depth bits << 32 | id) to the bin's segment through a per-bin cursor
//                  (order inside a segment is arbitrary);
//   k_bin_sort     ONE launch, one CTA per bin: stable LSD radix sort of the segment on the depth
//                  word, four 8-bit passes over L2-resident ping-pong buffers, ranks from
//                  __match_any_sync multi-splits and per-warp shared-memory histograms.  Depth ties
//                  (rare: equal fp32 view depths inside one bin) are put in index order afterwards --
//                  short runs by insertion, long runs by re-sorting the bin on the full 64-bit key.
//
// Five launches, no scan over Gaussians, no padding of a speculative capacity, no ranges pass; every
// bin is sorted concurrently, so the stage is bound by four L2 round trips, not by the pair count.
#include "common.cuh"
#include "kernels.cuh"
#include "spans.cuh"

namespace b200gs {

// ---- k_bin_scan ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_bin_scan(BucketArgs a) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < a.num_bins; b0 += 1024) {
    const uint32_t b = b0 + tid;
    const uint32_t c = b < a.num_bins ? a.bin_count[(size_t)b * BIN_STRIDE] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wp = 0;
    for (int w = 0; w < warp; w++) wp += s_warp[w];
    const uint32_t carry = s_carry;
    const uint32_t base = carry + wp + incl - c;
    if (b < a.num_bins) {
      a.bin_base[b] = base;
      // a bin that does not fit the pair capacity is left empty: the host learns D > capacity and
      // redoes the stage, and until then no kernel may index past the buffers
      const bool fits = (uint64_t)base + c <= (uint64_t)a.capacity;
      a.ranges[b] = fits ? make_uint2(base, base + c) : make_uint2(0u, 0u);
    }
    __syncthreads();
    if (tid == 1023) s_carry = carry + wp + incl;
    __syncthreads();
  }
  if (tid == 0) *a.total = s_carry;   // D
}

// ---- k_emit_bucket / k_emit_bucket_big ------------------------------------------------------------
constexpr uint32_t BUCKET_BIG_THRESHOLD = 12;   // bins; above this a whole warp emits the Gaussian

__device__ __forceinline__ void emit_one(const BucketArgs& a, uint32_t bin, uint64_t key) {
  const uint32_t slot = a.bin_base[bin] + atomicAdd(a.bin_cursor + (size_t)bin * BIN_STRIDE, 1u);
  if (slot < a.capacity) a.seg[slot] = key;
}

__global__ void __launch_bounds__(256) k_emit_bucket(BucketArgs a) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.P) return;
  const uint32_t n = a.tiles[r];
  if (n == 0) return;
  if (n > BUCKET_BIG_THRESHOLD) {
    a.big_queue[atomicAdd(a.big_count, 1u)] = (uint32_t)r;
    return;
  }
  const float4 q0 = a.rec[(size_t)r * REC_F4], q1 = a.rec[(size_t)r * REC_F4 + 1];
  const uint64_t key = ((uint64_t)a.depth_key[r] << 32) | (uint32_t)r;
  const TileRect rect = bin_rect(reference_rect(q0.x, q0.y, a.radii[r], a.gx, a.gy), a.bin_shift);
  SpanCtx s;
  if (!span_setup(s, q0.x, q0.y, q0.z, q0.w, q1.x, q1.z, rect, a.bin_shift)) return;
  for (int ty = s.ty0; ty < s.ty1; ty++) {
    int c0, c1;
    row_span(s, rect, ty, c0, c1);
    for (int tx = c0; tx < c1; tx++) emit_one(a, (uint32_t)(ty * a.gbx + tx), key);
  }
}

// one warp per queued Gaussian: lanes take the bin rows, then the bins of each row, in parallel
__global__ void __launch_bounds__(256) k_emit_bucket_big(BucketArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t count = *a.big_count;
  for (uint32_t w = warp_global; w < count; w += nwarps) {
    const uint32_t g = a.big_queue[w];
    const float4 q0 = a.rec[(size_t)g * REC_F4], q1 = a.rec[(size_t)g * REC_F4 + 1];
    const uint64_t key = ((uint64_t)a.depth_key[g] << 32) | g;
    const TileRect rect = bin_rect(reference_rect(q0.x, q0.y, a.radii[g], a.gx, a.gy), a.bin_shift);
    SpanCtx s;
    if (!span_setup(s, q0.x, q0.y, q0.z, q0.w, q1.x, q1.z, rect, a.bin_shift)) continue;
    for (int y_base = s.ty0; y_base < s.ty1; y_base += 32) {
      const int ty = y_base + lane;
      int c0 = 0, c1 = 0;
      if (ty < s.ty1) row_span(s, rect, ty, c0, c1);
      const int rows = min(32, s.ty1 - y_base);
      for (int i = 0; i < rows; i++) {
        const int c0_i = __shfl_sync(0xffffffffu, c0, i), c1_i = __shfl_sync(0xffffffffu, c1, i);
        for (int tx = c0_i + lane; tx < c1_i; tx += 32) emit_one(a, (uint32_t)((y_base + i) * a.gbx + tx), key);
      }
    }
  }
}

// ---- k_bin_sort ---------------------------------------------------------------------------------
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

__global__ void __launch_bounds__(BS_THREADS) k_bin_sort(BucketArgs a) {
  __shared__ BinSortShared sh;
  const uint32_t b = blockIdx.x;
  const uint2 range = a.ranges[b];
  const uint32_t n = range.y - range.x;
  if (n == 0) return;
  uint64_t* A = a.seg + range.x;
  uint64_t* B = a.seg_alt + range.x;
  const int tid = threadIdx.x;
  if (tid == 0) sh.long_run = 0u;
  // four passes over the depth word: A -> B -> A -> B -> A
  for (int p = 0; p < 4; p++) {
    bin_sort_pass(sh, (p & 1) ? B : A, (p & 1) ? A : B, n, 32 + 8 * p);
  }
  uint64_t* F = A;
  // depth ties: heads of equal-depth runs put their run in index order
  for (uint32_t i = tid; i + 1 < n; i += BS_THREADS) {
    const uint32_t d = (uint32_t)(__ldcg(F + i) >> 32);
    if ((uint32_t)(__ldcg(F + i + 1) >> 32) != d) continue;
    if (i > 0 && (uint32_t)(__ldcg(F + i - 1) >> 32) == d) continue;   // not the head
    uint32_t e = i + 2;
    while (e < n && e - i <= BS_TIE_INSERTION_MAX && (uint32_t)(__ldcg(F + e) >> 32) == d) e++;
    if (e - i > BS_TIE_INSERTION_MAX) { sh.long_run = 1u; continue; }
    for (uint32_t x = i + 1; x < e; x++) {      // insertion sort of [i, e) on the full key
      const uint64_t k = __ldcg(F + x);
      uint32_t y = x;
      while (y > i && __ldcg(F + y - 1) > k) { F[y] = __ldcg(F + y - 1); y--; }
      F[y] = k;
    }
  }
  __syncthreads();
  if (sh.long_run) {
    // many equal depths (e.g. a fronto-parallel planar scene): sort the bin on the full 64-bit key,
    // index digits first (stable LSD), then the depth word again
    int passes = 0;
    for (int s = 0; s < a.id_bits; s += 8, passes++) {
      bin_sort_pass(sh, (passes & 1) ? B : A, (passes & 1) ? A : B, n, s);
    }
    for (int p = 0; p < 4; p++, passes++) {
      bin_sort_pass(sh, (passes & 1) ? B : A, (passes & 1) ? A : B, n, 32 + 8 * p);
    }
    F = (passes & 1) ? B : A;
  }
  uint32_t* out = a.vals_sorted + range.x;
  for (uint32_t i = tid; i < n; i += BS_THREADS) out[i] = (uint32_t)__ldcg(F + i);
}

// ---- host ---------------------------------------------------------------------------------------
void launch_bin_scan(const BucketArgs& a, cudaStream_t st) {
  k_bin_scan<<<1, 1024, 0, st>>>(a);
  count_launch();
}

void launch_bucket_emit_sort_emit(const BucketArgs& a, cudaStream_t st) {
  if (a.P == 0 || a.capacity == 0) return;
  k_emit_bucket<<<(a.P + 255) / 256, 256, 0, st>>>(a);
  k_emit_bucket_big<<<148 * 2, 256, 0, st>>>(a);
  count_launch(2);
}

void launch_bucket_emit_sort_sort(const BucketArgs& a, cudaStream_t st) {
  if (a.P == 0 || a.capacity == 0) return;
  k_bin_sort<<<a.num_bins, BS_THREADS, 0, st>>>(a);
  count_launch();
}

}  // namespace b200gs

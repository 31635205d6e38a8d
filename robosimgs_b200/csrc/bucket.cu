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
#include "binsort.cuh"

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
constexpr int EMIT_WARPS = 8;
constexpr int EMIT_STAGE = 32 * BUCKET_BIG_THRESHOLD;   // staged (bin, owner lane) entries per warp

__device__ __forceinline__ void emit_one(const BucketArgs& a, uint32_t bin, uint64_t key) {
  const uint32_t slot = __ldg(a.bin_base + bin) + atomicAdd(a.bin_cursor + (size_t)bin * BIN_STRIDE, 1u);
  if (slot < a.capacity) a.seg[slot] = key;
}

// A cursor increment is an L2 round trip (~0.5 us) and a Gaussian's bins depend on nothing, so a lane
// that walked its own bins one atomic after the other would serialise up to 12 round trips while most
// lanes of the warp (culled Gaussians) idle.  Instead every lane first STAGES its bins in shared memory
// (no memory traffic), then the warp drains the staged list 32 pairs per round, one atomic per lane.
__global__ void __launch_bounds__(32 * EMIT_WARPS) k_emit_bucket(BucketArgs a) {
  __shared__ uint32_t stage[EMIT_WARPS][EMIT_STAGE];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t n = (r < a.P) ? a.tiles[r] : 0u;
  if (n > BUCKET_BIG_THRESHOLD) {
    a.big_queue[atomicAdd(a.big_count, 1u)] = (uint32_t)r;
    n = 0;
  }
  uint32_t incl = n;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;
  uint32_t depth = 0;
  if (n) {
    depth = a.depth_key[r];
    const float4 q0 = a.rec[(size_t)r * REC_F4], q1 = a.rec[(size_t)r * REC_F4 + 1];
    const TileRect rect = bin_rect(reference_rect(q0.x, q0.y, a.radii[r], a.gx, a.gy), a.bin_shift);
    SpanCtx s;
    uint32_t o = incl - n;
    const uint32_t end = incl;
    if (span_setup(s, q0.x, q0.y, q0.z, q0.w, q1.x, q1.z, rect, a.bin_shift)) {
      for (int ty = s.ty0; ty < s.ty1; ty++) {
        int c0, c1;
        row_span(s, rect, ty, c0, c1);
        for (int tx = c0; tx < c1 && o < end; tx++, o++) stage[warp][o] = (uint32_t)(ty * a.gbx + tx) | ((uint32_t)lane << 16);
      }
    }
    for (; o < end; o++) stage[warp][o] = 0xFFFFu | ((uint32_t)lane << 16);   // defensive: count == emit by construction
  }
  __syncwarp();
  for (uint32_t k0 = 0; k0 < total; k0 += 32) {
    const uint32_t k = k0 + lane;
    const uint32_t e = (k < total) ? stage[warp][k] : 0u;
    const int owner = (int)(e >> 16);
    const uint32_t od = __shfl_sync(0xffffffffu, depth, owner);
    const uint32_t oid = (uint32_t)(r - lane + owner);
    const uint32_t bin = e & 0xFFFFu;
    if (k < total && bin != 0xFFFFu) emit_one(a, bin, ((uint64_t)od << 32) | oid);
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

// ---- k_bin_sort (BinSortShared / bin_sort_pass: binsort.cuh) ---------------------------------------
// Same pass for segments of at most K*32*BS_WARPS elements, K keys per thread: the keys of a pass are
// fetched with ONE batch of independent loads and stay in registers between the ranking and the
// scatter, so a pass costs one L2 round trip instead of one per 32 elements.  Peer masks come from
// eight ballots (one per digit bit) rather than MATCH.ANY.
template <int K>
__device__ __forceinline__ void bin_sort_pass_regs(BinSortShared& sh, const uint64_t* __restrict__ src,
                                                   uint64_t* __restrict__ dst, uint32_t n, int shift) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t wbeg = (uint32_t)warp * (K * 32);
  uint64_t key[K];
#pragma unroll
  for (int it = 0; it < K; it++) {
    const uint32_t i = wbeg + it * 32 + lane;
    key[it] = (i < n) ? __ldcg(src + i) : ~0ull;
  }
  for (int i = tid; i < BS_WARPS * 256; i += BS_THREADS) (&sh.hist[0][0])[i] = 0u;
  __syncthreads();
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t local[K];   // rank among the warp's elements of the same digit
#pragma unroll
  for (int it = 0; it < K; it++) {
    const bool ok = wbeg + it * 32 + lane < n;
    const uint32_t d = (uint32_t)(key[it] >> shift) & 255u;
    uint32_t peers = __ballot_sync(0xffffffffu, ok);
    if (peers == 0u) { local[it] = 0u; continue; }     // warp-uniform: the slice ended
#pragma unroll
    for (int bit = 0; bit < 8; bit++) {
      const bool one = (d >> bit) & 1u;
      const uint32_t m = __ballot_sync(0xffffffffu, one);
      peers &= one ? m : ~m;
    }
    if (!ok) peers = 0u;
    const int leader = __ffs(peers) - 1;
    uint32_t off = 0u;
    if (ok && lane == leader) {
      off = sh.hist[warp][d];
      sh.hist[warp][d] = off + __popc(peers);
    }
    off = __shfl_sync(0xffffffffu, off, leader & 31);
    local[it] = off + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();
  if (tid < 256) {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < BS_WARPS; w++) {
      const uint32_t c = sh.hist[w][tid];
      sh.hist[w][tid] = run;
      run += c;
    }
    uint32_t incl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) sh.wsum[warp] = incl;
    sh.digit_base[tid] = incl - run;
  }
  __syncthreads();
  if (tid < 256) {
    uint32_t wp = 0;
    for (int w = 0; w < warp; w++) wp += sh.wsum[w];
    sh.digit_base[tid] += wp;
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < K; it++) {
    if (wbeg + it * 32 + lane < n) {
      const uint32_t d = (uint32_t)(key[it] >> shift) & 255u;
      dst[sh.digit_base[d] + sh.hist[warp][d] + local[it]] = key[it];
    }
  }
  __syncthreads();
}

// K = 0: any length (keys re-read per 32-element group); K > 0: segments of (lo, K*32*BS_WARPS] elements.
// One launch per class; a CTA whose bin belongs to another class exits at once, so small bins are not
// sorted by a kernel that carries the register budget of the large ones.
template <int K>
__global__ void __launch_bounds__(BS_THREADS) k_bin_sort(BucketArgs a, uint32_t lo, uint32_t hi) {
  __shared__ BinSortShared sh;
  const uint32_t b = blockIdx.x;
  const uint2 range = a.ranges[b];
  const uint32_t n = range.y - range.x;
  if (n <= lo || n > hi) return;
  uint64_t* A = a.seg + range.x;
  uint64_t* B = a.seg_alt + range.x;
  const int tid = threadIdx.x;
  if (tid == 0) sh.long_run = 0u;
  auto pass = [&](const uint64_t* src, uint64_t* dst, int shift) {
    if constexpr (K == 0) bin_sort_pass(sh, src, dst, n, shift);
    else bin_sort_pass_regs<K>(sh, src, dst, n, shift);
  };
  // four passes over the depth word: A -> B -> A -> B -> A
  for (int p = 0; p < 4; p++) pass((p & 1) ? B : A, (p & 1) ? A : B, 32 + 8 * p);
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
    for (int s = 0; s < a.id_bits; s += 8, passes++) pass((passes & 1) ? B : A, (passes & 1) ? A : B, s);
    for (int p = 0; p < 4; p++, passes++) pass((passes & 1) ? B : A, (passes & 1) ? A : B, 32 + 8 * p);
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
  constexpr uint32_t SMALL = 8 * 32 * BS_WARPS, LARGE = 24 * 32 * BS_WARPS;   // 4096, 12288
  k_bin_sort<8><<<a.num_bins, BS_THREADS, 0, st>>>(a, 0u, SMALL);
  k_bin_sort<24><<<a.num_bins, BS_THREADS, 0, st>>>(a, SMALL, LARGE);
  k_bin_sort<0><<<a.num_bins, BS_THREADS, 0, st>>>(a, LARGE, 0xFFFFFFFFu);
  count_launch(3);
}

}  // namespace b200gs

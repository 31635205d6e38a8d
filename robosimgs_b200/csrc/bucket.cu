// bucket.cu -- depth-sliced bucket binning: per-bin, depth-ordered pair lists without a global sort.
//
// Replaces rows a4 InclusiveSum, a5 duplicateWithKeys, a6 SortPairs and a7 identifyTileRanges
// of the public algorithm -- same result (every bin's Gaussian ids in (depth bits, index) order, which is the
// order a stable radix sort of (bin << 32 | depth) keys emitted in index order produces), different route.
// The global-sort pipeline (binning.cu: library scan + onesweep passes) is latency-bound on B200: 0.55 M pairs
// are 4 MB, yet scan + emit + 4 passes + ranges take 0.14 ms of a 0.36 ms C3 frame in 13 dependent launches.
//
// A pair's BUCKET is (bin, depth slice), slice = min(((depth bits - near-plane bits) >> slice_shift), S - 1):
// the IEEE bits of a positive float grow with the float, so the slices cut the depth axis into S monotone
// (logarithmic: 2^23 codes per octave, 16 octaves from the near plane) intervals, and the concatenation of a
// bin's buckets in slice order is depth-ordered as soon as every bucket is.  S ~ 512 k / bins (8192 slices of
// 0.14 % depth each at 1080p with 256-px bins), so a bucket holds about a dozen pairs.
//
//   k_project        counts pairs per bucket while it computes the spans (one RED per pair);
//   k_bucket_scan    one CTA per bin: exclusive scan of the bin's S counters, bin offsets by a decoupled look-back
//                    over the lower bins -> absolute bucket starts (also the append cursors), the per-bin
//                    [start,end) ranges the compositing kernels read, D, and the WINDOW table: a bin's list is cut
//                    every BUCKET_WINDOW pairs at the next bucket boundary, window k starts at the first bucket
//                    whose start is >= k * BUCKET_WINDOW -- equal work per sorter whatever the depth histogram;
//   k_emit_bucket    appends ((depth bits - near bits) << 32 | id) to the pair's bucket through its cursor
//                    (order inside a bucket is arbitrary; the sort key is a total order, so the result is
//                    deterministic);
//   k_bucket_sort    one warp per window: its buckets are contiguous in memory and ordered by slice, so sorting
//                    their union on the 64-bit key is exactly the (depth, index) order; up to 512 keys in
//                    registers, bitonic network with shuffles for the cross-lane stages.  A window that holds a
//                    larger bucket (a wall that faces the camera puts a whole bin at one depth) is queued for
//   k_bucket_sort_big  one CTA per queued segment: LSD radix passes over the key digits that actually differ
//                    inside the segment, L2-resident ping-pong.
//
// Four launches behind the projection kernel, no scan over Gaussians, no padding of a speculative capacity, no ranges pass, no tie repair.
#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#include "spans.cuh"
#include "binsort.cuh"

// CTAs of the one-CTA-per-fat-segment kernel (persistent over the queue).  A CTA needs a whole SM's register file (512
// threads x 128 registers), so each one waits for an SM to drain -- also when the queue is empty.
#ifndef BUCKET_SORT_BIG_CTAS
#define BUCKET_SORT_BIG_CTAS (148 * 2)
#endif
#ifndef BUCKET_SORT_BIG_SKIP      // measurement only: never launch it (wrong on scenes with fat segments)
#define BUCKET_SORT_BIG_SKIP 0
#endif

namespace b200gs {

// ---- k_bucket_scan --------------------------------------------------------------------------------
// One CTA per bin.  A bin's offsets are the sums over the lower bins, obtained by a decoupled look-back over the words
// the bins publish in bin_pub[] (CTAs are dispatched in index order, so a CTA only ever waits for CTAs that are
// already running or done).
constexpr int SCAN_THREADS = 1024;     // launched with min(1024, max(32, S / 4)) threads: four counters per thread
                                       // (eight at 8192 slices)

__global__ void __launch_bounds__(SCAN_THREADS) k_bucket_scan(BucketArgs a) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_off[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bin = blockIdx.x;
  const uint32_t S = 1u << a.slices_log2;
  const uint32_t per = S >= 8 * SCAN_THREADS ? 8u : 4u;          // counters per thread (S <= 8192)
  const uint32_t b_first = bin * S + (uint32_t)tid * per;        // global index of the thread's first bucket
  const bool active = (uint32_t)tid * per < S;
  uint32_t c[8];
#pragma unroll
  for (int k = 0; k < 8; k++) c[k] = 0u;
  if (active) {
    const uint4 v0 = __ldcg(reinterpret_cast<const uint4*>(a.bucket_count + b_first));
    c[0] = v0.x; c[1] = v0.y; c[2] = v0.z; c[3] = v0.w;
    if (per == 8u) {
      const uint4 v1 = __ldcg(reinterpret_cast<const uint4*>(a.bucket_count + b_first + 4));
      c[4] = v1.x; c[5] = v1.y; c[6] = v1.z; c[7] = v1.w;
    }
  }
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) sum += c[k];
  uint32_t incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = lane < (blockDim.x >> 5) ? s_warp[lane] : 0u;
    uint32_t wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += t;
    }
    s_warp[lane] = wi - w;                                   // exclusive prefix of the warp totals
    const uint32_t bin_total = __shfl_sync(0xffffffffu, wi, 31);
    const uint32_t nwin = (bin_total + BUCKET_WINDOW - 1) / BUCKET_WINDOW;
    // Decoupled look-back.  A bin first publishes its own AGGREGATE (pairs, windows), then -- once it knows the sums
    // over all lower bins -- its INCLUSIVE PREFIX; a bin looks back 32 words at a time until it meets a prefix, so the
    // walk is a few words long whatever the number of bins.  Word: status (2 bits) | windows (30) | pairs (32).
    volatile unsigned long long* pub = reinterpret_cast<volatile unsigned long long*>(a.bin_pub);
    constexpr unsigned long long ST_AGG = 1ull << 62, ST_PREFIX = 2ull << 62;
    if (lane == 0) pub[bin] = ST_AGG | ((unsigned long long)nwin << 32) | bin_total;
    uint32_t off = 0, woff = 0;
    for (int64_t hi = (int64_t)bin - 1; hi >= 0; hi -= 32) {
      const int64_t j = hi - lane;
      unsigned long long v = ST_PREFIX;                     // lanes past bin 0 contribute an empty prefix
      if (j >= 0) do { v = pub[j]; } while (!(v >> 62));
      // the nearest prefix (smallest lane) ends the walk: lanes nearer than it add their aggregates, it adds itself
      const uint32_t is_prefix = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
      const int first = is_prefix ? __ffs(is_prefix) - 1 : 32;
      if (lane <= first && j >= 0) {
        off += (uint32_t)v;
        woff += (uint32_t)(v >> 32) & 0x3FFFFFFFu;
      }
      if (is_prefix) break;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      off += __shfl_xor_sync(0xffffffffu, off, d);
      woff += __shfl_xor_sync(0xffffffffu, woff, d);
    }
    if (lane == 0) pub[bin] = ST_PREFIX | ((unsigned long long)(woff + nwin) << 32) | (unsigned long long)(off + bin_total);
    if (lane == 0) { s_off[0] = off; s_off[1] = woff; s_off[2] = bin_total; }
  }
  __syncthreads();
  const uint32_t off = s_off[0], woff = s_off[1], bin_total = s_off[2];
  const uint32_t nwin = (bin_total + BUCKET_WINDOW - 1) / BUCKET_WINDOW;
  if (active) {
    uint32_t lp = s_warp[warp] + incl - sum;                 // local (in-bin) start of the thread's first bucket
    uint32_t o[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      o[k] = off + lp;
      // window table: bucket b = [lp, lp + c) makes b + 1 the first bucket at or behind every window boundary
      // W k in (lp, lp + c]; the boundary at the very end of the bin belongs to the next bin's first window
      if (a.win_first && c[k] && (uint32_t)k < per) {
        const uint32_t k1 = min((lp + c[k]) / BUCKET_WINDOW, nwin ? nwin - 1u : 0u);
        for (uint32_t w = lp / BUCKET_WINDOW + 1u; w <= k1; w++)
          if (woff + w < a.win_capacity) a.win_first[woff + w] = b_first + k + 1u;     // (more pairs than capacity)
      }
      lp += c[k];
    }
    *reinterpret_cast<uint4*>(a.bucket_base + b_first) = make_uint4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<uint4*>(a.bucket_cursor + b_first) = make_uint4(o[0], o[1], o[2], o[3]);
    if (per == 8u) {
      *reinterpret_cast<uint4*>(a.bucket_base + b_first + 4) = make_uint4(o[4], o[5], o[6], o[7]);
      *reinterpret_cast<uint4*>(a.bucket_cursor + b_first + 4) = make_uint4(o[4], o[5], o[6], o[7]);
    }
  }
  if (tid == 0) {
    if (a.win_first && nwin && woff < a.win_capacity) a.win_first[woff] = bin * S;
    // A bin that does not fit the pair capacity is left empty: the caller learns D > capacity and redoes the
    // stage, and until then no kernel may index past the buffers.
    a.ranges[bin] = (off + bin_total <= a.capacity) ? make_uint2(off, off + bin_total) : make_uint2(0u, 0u);
    if (bin == a.num_bins - 1) {
      a.bucket_base[a.num_bins * S] = off + bin_total;
      *a.total = off + bin_total;                            // D
      *a.total_windows = woff + nwin;
      if (a.win_first && woff + nwin < a.win_capacity) a.win_first[woff + nwin] = a.num_bins * S;
    }
  }
}

// ---- k_emit_bucket / k_emit_bucket_big ------------------------------------------------------------
constexpr uint32_t BUCKET_BIG_THRESHOLD = 12;   // bins; above this a whole warp emits the Gaussian
constexpr int EMIT_WARPS = 8;
#ifndef EMIT_WAVES
#define EMIT_WAVES 2
#endif
constexpr int EMIT_STAGE = 32 * BUCKET_BIG_THRESHOLD;   // staged (bin, owner lane) entries per warp

__device__ __forceinline__ uint32_t rel_depth(const BucketArgs& a, uint32_t depth_bits) {
  return depth_bits > a.near_bits ? depth_bits - a.near_bits : 0u;
}
__device__ __forceinline__ uint32_t slice_of(const BucketArgs& a, uint32_t rel) {
  return min(rel >> a.slice_shift, (1u << a.slices_log2) - 1u);
}
__device__ __forceinline__ void emit_one(const BucketArgs& a, uint32_t bin, uint32_t slice, uint64_t key) {
  const uint32_t slot = atomicAdd(a.bucket_cursor + ((bin << a.slices_log2) + slice), 1u);
  if (slot < a.capacity) a.seg[slot] = key;
}
// The same with the store DEFERRED: the slot a cursor atomic returns is only needed by the 8-byte store, so a lane
// parks (slot, key) and writes them out just before its next atomic -- by then the L2 round trip is over and the
// loads of the warp's next 32 Gaussians are in flight behind it.
struct PendingPair {
  uint64_t key;
  uint32_t slot;      // 0xFFFFFFFF: nothing parked
};
__device__ __forceinline__ void flush_pending(const BucketArgs& a, PendingPair& p) {
  if (p.slot < a.capacity) a.seg[p.slot] = p.key;
  p.slot = 0xFFFFFFFFu;
}
__device__ __forceinline__ void emit_deferred(const BucketArgs& a, PendingPair& p, uint32_t bin, uint32_t slice, uint64_t key) {
  flush_pending(a, p);
  p.slot = atomicAdd(a.bucket_cursor + ((bin << a.slices_log2) + slice), 1u);
  p.key = key;
}

// A cursor increment is an L2 round trip (~0.5 us) and a Gaussian's bins depend on nothing, so a lane that
// walked its own bins one atomic after the other would serialise up to 12 round trips while most lanes of the
// warp (culled Gaussians) idle.  Instead every lane first STAGES its bins in shared memory (no memory
// traffic), then the warp drains the staged list 32 pairs per round, one atomic per lane.
__global__ void __launch_bounds__(32 * EMIT_WARPS) k_emit_bucket(BucketArgs a) {
  __shared__ uint32_t stage[EMIT_WARPS][EMIT_STAGE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int stride = gridDim.x * blockDim.x;
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  // Grid-stride loop: the footprint word and the depth of a thread's NEXT Gaussian are loaded while it works on the
  // current one (the depth travels with the footprint word, one coalesced line each: loading it only once the
  // footprint is known to be non-empty would put a second dependent round trip in front of the cursor atomics).
  uint32_t tword_n = (r < a.P) ? a.tiles[r] : 0u;     // count, or a packed footprint of up to three bins
  uint32_t dkey_n = (r < a.P) ? a.depth_key[r] : 0u;
  PendingPair pend{0ull, 0xFFFFFFFFu};
  for (; r - lane < a.P; r += stride) {
  const uint32_t tword = tword_n, dkey = dkey_n;
  tword_n = (r + stride < a.P) ? a.tiles[r + stride] : 0u;
  dkey_n = (r + stride < a.P) ? a.depth_key[r + stride] : 0u;
  uint32_t n = tiles_count(tword);
  const uint32_t big_lanes = __ballot_sync(0xffffffffu, n > BUCKET_BIG_THRESHOLD);
  if (n > BUCKET_BIG_THRESHOLD) n = 0;          // large footprints: emitted by the whole warp below
  uint32_t incl = n;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0 && big_lanes == 0u) continue;
  uint32_t rel = 0;
  if (n && (tword & TILES_PACKED)) {
    // small footprint: the projection kernel left the bin ids in tiles[] -- no record load, no span arithmetic
    rel = rel_depth(a, dkey);
    uint32_t o = incl - n;
    for (uint32_t k = 0; k < n; k++, o++) stage[warp][o] = ((tword >> (8u * k)) & 0xFFu) | ((uint32_t)lane << 16);
  } else if (n) {
    rel = rel_depth(a, dkey);
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
    const uint32_t orel = __shfl_sync(0xffffffffu, rel, owner);
    const uint32_t oid = (uint32_t)(r - lane + owner);
    const uint32_t bin = e & 0xFFFFu;
    if (k < total && bin != 0xFFFFu) emit_deferred(a, pend, bin, slice_of(a, orel), ((uint64_t)orel << 32) | oid);
  }
  // large footprints (a few screen-filling splats): one Gaussian at a time, lanes take the bin rows, then the bins of
  // each row, in parallel -- so that one of them cannot serialise a lane for hundreds of atomics
  for (uint32_t m = big_lanes; m; m &= m - 1u) {
    const uint32_t g = (uint32_t)(r - lane) + (uint32_t)(__ffs(m) - 1);
    const float4 q0 = a.rec[(size_t)g * REC_F4], q1 = a.rec[(size_t)g * REC_F4 + 1];
    const uint32_t grel = rel_depth(a, a.depth_key[g]);
    const uint32_t slice = slice_of(a, grel);
    const uint64_t key = ((uint64_t)grel << 32) | g;
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
        for (int tx = c0_i + lane; tx < c1_i; tx += 32) emit_one(a, (uint32_t)((y_base + i) * a.gbx + tx), slice, key);
      }
    }
  }
  __syncwarp();      // the staged list is rewritten by the next step
  }
  flush_pending(a, pend);
}

// ---- k_bucket_sort: one warp per bucket, keys in registers ---------------------------------------------
// Element i of the bucket lives in register i / 32 of lane i % 32.  Bitonic network: partners 32 or more
// apart are two registers of the same lane, closer partners are exchanged with shuffles.
#ifndef BUCKET_SORT_ROLLED
#define BUCKET_SORT_ROLLED 0
#endif
// Unrolled by default: the five instantiations are 21 k instructions of straight-line code, but rolling the stage loops
// (-DBUCKET_SORT_ROLLED=1: 1.9 k instructions) pays for its dynamic stage predicates with more issue slots -- measured
// on C3: the sort stage alone 0.0375 -> 0.0343 ms, the four-frames-in-flight sweep 0.1746 -> 0.1784 ms per frame.
template <int K>
__device__ __forceinline__ void warp_bitonic_sort(uint64_t (&key)[K], int lane) {
  constexpr int N = 32 * K;
#if BUCKET_SORT_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
  for (int k = 2; k <= N; k <<= 1) {
#if BUCKET_SORT_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int jr = j >> 5;
#pragma unroll
        for (int JR = 1; JR < K; JR <<= 1) {
          if (jr == JR) {
#pragma unroll
            for (int r = 0; r < K; r++) {
              if ((r & JR) == 0) {
                const bool asc = (((r << 5) & k) == 0);   // bit k of the element index: k >= 64 here, a register bit
                const uint64_t lo = key[r] < key[r | JR] ? key[r] : key[r | JR];
                const uint64_t hi = key[r] < key[r | JR] ? key[r | JR] : key[r];
                key[r] = asc ? lo : hi;
                key[r | JR] = asc ? hi : lo;
              }
            }
          }
        }
      } else {
        const bool lower = (lane & j) == 0;           // this lane keeps the smaller key when ascending
#pragma unroll
        for (int r = 0; r < K; r++) {
          const uint64_t other = __shfl_xor_sync(0xffffffffu, key[r], j);
          const bool take_min = ((((r << 5) | lane) & k) == 0) == lower;
          const bool other_less = other < key[r];
          key[r] = (take_min == other_less) ? other : key[r];
        }
      }
    }
  }
}

template <int K>
__device__ __forceinline__ void sort_bucket_in_registers(const uint64_t* __restrict__ seg, uint32_t* __restrict__ out,
                                                         uint32_t n, int lane) {
  uint64_t key[K];
#pragma unroll
  for (int r = 0; r < K; r++) {
    const uint32_t i = (uint32_t)(r * 32 + lane);
    key[r] = i < n ? __ldcg(seg + i) : ~0ull;          // padding sorts last
  }
  warp_bitonic_sort<K>(key, lane);
#pragma unroll
  for (int r = 0; r < K; r++) {
    const uint32_t i = (uint32_t)(r * 32 + lane);
    if (i < n) out[i] = (uint32_t)key[r];
  }
}

constexpr int SORT_WARPS = 8;

__global__ void __launch_bounds__(32 * SORT_WARPS) k_bucket_sort(BucketArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t w = blockIdx.x * SORT_WARPS + (threadIdx.x >> 5);
  if (w >= __ldcg(a.total_windows) || w + 1 >= a.win_capacity) return;
  const uint32_t b0 = __ldcg(a.win_first + w), b1 = __ldcg(a.win_first + w + 1);
  const uint32_t s = __ldcg(a.bucket_base + b0), e = __ldcg(a.bucket_base + b1);
  const uint32_t n = e - s;
  if (n == 0 || e > a.capacity) return;        // over capacity: the frame is redone, never index past the buffers
  const uint64_t* seg = a.seg + s;
  uint32_t* out = a.vals_sorted + s;
  if (n > WARP_SORT_MAX) {
    if (lane == 0) a.big_segs[atomicAdd(a.big_seg_count, 1u)] = make_uint2(s, n);
    return;
  }
  if (n == 1) { if (lane == 0) out[0] = (uint32_t)__ldcg(seg); }
  else if (n <= 32) sort_bucket_in_registers<1>(seg, out, n, lane);
  else if (n <= 64) sort_bucket_in_registers<2>(seg, out, n, lane);
  else if (n <= 128) sort_bucket_in_registers<4>(seg, out, n, lane);
  else if (n <= 256) sort_bucket_in_registers<8>(seg, out, n, lane);
  else if (WARP_SORT_MAX > 256) sort_bucket_in_registers<(WARP_SORT_MAX > 256 ? 16 : 8)>(seg, out, n, lane);
  if (a.slab) {     // record slab for the TMA-fed compositing ring (option "slab"): slab[s + i] = rec[out[i]]
    __syncwarp();   // the ids this warp just wrote are visible to all its lanes
    float4* dst = a.slab + (size_t)s * REC_F4;
    for (uint32_t q = lane; q < n * REC_F4; q += 32) {
      const uint32_t i = q / REC_F4;
      dst[q] = a.rec[(size_t)out[i] * REC_F4 + (q - i * REC_F4)];
    }
  }
}

// ---- k_bucket_sort_big (BinSortShared / bin_sort_pass: binsort.cuh) -----------------------------------
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

// One CTA per queued segment (a window that holds a bucket of more than a few hundred pairs).  LSD radix passes
// over the 8-bit digits of the 64-bit key that are not constant over the segment: inside one bucket only the index
// bits and the low slice_shift depth bits differ.
__global__ void __launch_bounds__(BS_THREADS) k_bucket_sort_big(BucketArgs a) {
  __shared__ BinSortShared sh;
  __shared__ unsigned long long s_or, s_and;
  const uint32_t count = *a.big_seg_count;
  const int tid = threadIdx.x, lane = tid & 31;
  for (uint32_t q = blockIdx.x; q < count; q += gridDim.x) {
    const uint2 sg = a.big_segs[q];
    const uint32_t s = sg.x, n = sg.y;
    uint64_t* A = a.seg + s;
    uint64_t* B = a.seg_alt + s;
    if (tid == 0) { s_or = 0ull; s_and = ~0ull; }
    __syncthreads();
    unsigned long long o = 0ull, an = ~0ull;
    for (uint32_t i = tid; i < n; i += BS_THREADS) {
      const unsigned long long k = __ldcg(A + i);
      o |= k; an &= k;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      o |= __shfl_xor_sync(0xffffffffu, o, d);
      an &= __shfl_xor_sync(0xffffffffu, an, d);
    }
    if (lane == 0) { atomicOr(&s_or, o); atomicAnd(&s_and, an); }
    __syncthreads();
    const unsigned long long varying = s_or ^ s_and;
    __syncthreads();
    int passes = 0;
    for (int sft = 0; sft < 64; sft += 8) {
      if (!((varying >> sft) & 0xFFull)) continue;          // CTA-uniform: this digit is the same in every key
      const uint64_t* src = (passes & 1) ? B : A;
      uint64_t* dst = (passes & 1) ? A : B;
      if (n <= 2 * 32 * BS_WARPS) bin_sort_pass_regs<2>(sh, src, dst, n, sft);
      else if (n <= 8 * 32 * BS_WARPS) bin_sort_pass_regs<8>(sh, src, dst, n, sft);
      else if (n <= 24 * 32 * BS_WARPS) bin_sort_pass_regs<24>(sh, src, dst, n, sft);
      else bin_sort_pass(sh, src, dst, n, sft);
      passes++;
    }
    const uint64_t* F = (passes & 1) ? B : A;
    uint32_t* out = a.vals_sorted + s;
    for (uint32_t i = tid; i < n; i += BS_THREADS) out[i] = (uint32_t)__ldcg(F + i);
    if (a.slab) {
      float4* dst = a.slab + (size_t)s * REC_F4;
      for (uint32_t q = tid; q < n * REC_F4; q += BS_THREADS) {
        const uint32_t i = q / REC_F4;
        dst[q] = a.rec[(size_t)(uint32_t)__ldcg(F + i) * REC_F4 + (q - i * REC_F4)];
      }
    }
    __syncthreads();
  }
}

// ---- host ---------------------------------------------------------------------------------------
void launch_bucket_scan(const BucketArgs& a, cudaStream_t st) {
  const int threads = std::min(SCAN_THREADS, std::max(32, (1 << a.slices_log2) / 4));
  k_bucket_scan<<<a.num_bins, threads, 0, st>>>(a);
  count_launch();
}

void launch_bucket_emit(const BucketArgs& a, cudaStream_t st) {
  if (a.P == 0 || a.capacity == 0) return;
  // grid-stride kernel: a few resident waves, so that every thread sees several Gaussians with the next one's loads
  // in flight (B200GS_EMIT_WAVES=0: one Gaussian per thread)
  static std::atomic<int> ctas{0}, waves{-1};
  int cpw = ctas.load(), nw = waves.load();
  if (cpw == 0) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cpw = (sms > 0 ? sms : 148) * 8;          // 8 CTAs of 256 threads per SM
    const char* e = getenv("B200GS_EMIT_WAVES");
    nw = e ? atoi(e) : EMIT_WAVES;
    ctas.store(cpw);
    waves.store(nw);
  }
  const int need = (a.P + 255) / 256;
  k_emit_bucket<<<(nw > 0 ? std::min(need, cpw * nw) : need), 256, 0, st>>>(a);
  count_launch();
}

void launch_bucket_sort(const BucketArgs& a, cudaStream_t st) {
  if (a.P == 0 || a.capacity == 0) return;
  const uint32_t max_windows = a.capacity / BUCKET_WINDOW + a.num_bins;     // sum over bins of ceil(pairs / W)
  k_bucket_sort<<<(max_windows + SORT_WARPS - 1) / SORT_WARPS, 32 * SORT_WARPS, 0, st>>>(a);
#if !BUCKET_SORT_BIG_SKIP
  k_bucket_sort_big<<<BUCKET_SORT_BIG_CTAS, BS_THREADS, 0, st>>>(a);
#endif
  count_launch(2);
}

}  // namespace b200gs

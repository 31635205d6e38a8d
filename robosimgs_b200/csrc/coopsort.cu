// coopsort.cu -- single-launch cooperative LSD radix sort for the (bin << 32 | depth) pair keys.
//
// With coarse bins the pair list is small and lives in L2, so a library onesweep sort is bound by
// its six dependent kernel launches and their decoupled look-back chains (~16 us per pass regardless
// of size).  Here one persistent cooperative kernel (one CTA per SM)
// runs all passes: per pass a block (1) turns the global [digit][block] histogram into its scatter
// bases, (2) scatters its contiguous chunk in order -- stable ranks from a warp-level
// __match_any_sync multi-split plus a per-digit scan over the 16 warps -- and, while scattering,
// (3) accumulates the NEXT pass's [digit][destination block] histogram with global REDs, so there is
// exactly one grid barrier per pass.  Same key, same stable order as the library sort it replaces
// (SURVEY.md 8(a) row a6).  Measured: faster than CUB below ~0.25 M pairs, slower above (five grid
// barriers at ~5 us each) -- api.cu picks by size.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace b200gs {

constexpr int CS_THREADS = 512;
constexpr int CS_WARPS = CS_THREADS / 32;
constexpr int CS_SLOTS = 4;                          // items per thread per tile (striped)
constexpr int CS_TILE = CS_THREADS * CS_SLOTS;       // 2048 items between CTA barriers

struct CoopSortArgs {
  uint64_t* keys[2];
  uint32_t* vals[2];
  uint32_t* hist;    // [3][G][256] rotating global histograms (block-major: coalesced per-digit reads)
  uint32_t n;
  uint32_t chunk;    // items per block (multiple of CS_TILE)
  int passes;        // 8-bit digits, starting at bit 0
};

__device__ __forceinline__ uint32_t digit_of(uint64_t key, int pass) { return (uint32_t)(key >> (8 * pass)) & 255u; }

__global__ void __launch_bounds__(CS_THREADS, 1) k_coop_radix_sort(CoopSortArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_part[2][256];
  __shared__ uint32_t s_base[256];
  __shared__ uint32_t s_running[256];
  __shared__ uint32_t s_tilebase[256];
  __shared__ uint32_t s_warp_tot[8];
  __shared__ uint16_t s_wc[CS_SLOTS * CS_WARPS][256];   // per (slot, warp) digit counts -> offsets

  const uint32_t G = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
  const uint32_t lane = tid & 31, warp = tid >> 5;
  const uint32_t start = min(a.n, b * a.chunk), end = min(a.n, start + a.chunk);
  const uint32_t HG = 256 * G;

  // ---- pass-0 histogram of this block's chunk; zero the two other rotating tables ----
  if (tid < 256) s_hist[tid] = 0;
  __syncthreads();
  for (uint32_t i = start + tid; i < end; i += CS_THREADS) atomicAdd(&s_hist[digit_of(__ldcg(a.keys[0] + i), 0)], 1u);
  __syncthreads();
  if (tid < 256) {
    a.hist[b * 256 + tid] = s_hist[tid];
    a.hist[HG + b * 256 + tid] = 0;
    a.hist[2 * HG + b * 256 + tid] = 0;
  }
  grid.sync();

  for (int p = 0; p < a.passes; p++) {
    const uint64_t* kin = a.keys[p & 1];
    const uint32_t* vin = a.vals[p & 1];
    uint64_t* kout = a.keys[(p + 1) & 1];
    uint32_t* vout = a.vals[(p + 1) & 1];
    const uint32_t* hcur = a.hist + (size_t)(p % 3) * HG;
    uint32_t* hnext = a.hist + (size_t)((p + 1) % 3) * HG;
    uint32_t* hfree = a.hist + (size_t)((p + 2) % 3) * HG;   // read last pass, accumulated into next pass
    const bool more = p + 1 < a.passes;

    // ---- scatter bases: digit start + items of the same digit in lower blocks.  Thread (h, d)
    // sums digit d over half h of the blocks; all loads of a thread are independent. ----
    {
      const uint32_t d = tid & 255u, h = tid >> 8;
      const uint32_t half = (G + 1) / 2;
      const uint32_t b0 = h * half, b1 = min(G, b0 + half);
      uint32_t tot = 0, bel = 0;
#pragma unroll 16
      for (uint32_t bb = b0; bb < b1; bb++) {
        const uint32_t c = __ldcg(hcur + bb * 256 + d);   // L2: written by other SMs during this launch
        tot += c;
        bel += (bb < b) ? c : 0u;
      }
      s_part[h][d] = tot;
      if (h == 1) s_hist[d] = bel;
      __syncthreads();
      uint32_t total = 0, below = 0;
      if (tid < 256) {
        total = s_part[0][tid] + s_part[1][tid];
        below = bel + s_hist[tid];
        hfree[b * 256 + tid] = 0;
        // exclusive scan of `total` over the 256 digits: warp scan + 8 warp totals
        uint32_t incl = total;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, incl, k);
          if (lane >= (uint32_t)k) incl += t;
        }
        if (lane == 31) s_warp_tot[warp] = incl;
        s_base[tid] = incl - total + below;   // + totals of the lower warps, added below
        s_running[tid] = 0;
      }
      __syncthreads();
      if (tid < 256) {
        uint32_t off = 0;
        for (uint32_t w = 0; w < warp; w++) off += s_warp_tot[w];
        s_base[tid] += off;
      }
      // (visibility of s_base is ordered by the first barrier of the tile loop)
    }

    // ---- stable scatter of this block's chunk: CS_SLOTS striped items per thread per tile, i.e.
    // CS_SLOTS consecutive 512-item sub-tiles ranked together between the same barriers ----
    for (uint32_t tile = start; tile < end; tile += CS_TILE) {
      uint32_t* wc32 = reinterpret_cast<uint32_t*>(&s_wc[0][0]);
      for (uint32_t k = tid; k < CS_SLOTS * CS_WARPS * 256 / 2; k += CS_THREADS) wc32[k] = 0;
      uint64_t key[CS_SLOTS];
      uint32_t val[CS_SLOTS], digit[CS_SLOTS], peers[CS_SLOTS];
      bool valid[CS_SLOTS];
#pragma unroll
      for (int j = 0; j < CS_SLOTS; j++) {
        const uint32_t i = tile + j * CS_THREADS + tid;
        valid[j] = i < end;
        key[j] = 0; val[j] = 0;
        if (valid[j]) {
          key[j] = __ldcg(kin + i);
          val[j] = __ldcg(vin + i);
        }
      }
#pragma unroll
      for (int j = 0; j < CS_SLOTS; j++) {
        digit[j] = valid[j] ? digit_of(key[j], p) : 256u + lane;   // invalid lanes never match anybody
        peers[j] = __match_any_sync(0xffffffffu, digit[j]);
      }
      __syncthreads();   // s_wc zeroed (and, first tile, s_base complete)
#pragma unroll
      for (int j = 0; j < CS_SLOTS; j++)
        if (valid[j] && (peers[j] & ((1u << lane) - 1u)) == 0) s_wc[j * CS_WARPS + warp][digit[j]] = (uint16_t)__popc(peers[j]);
      __syncthreads();
      if (tid < 256) {
        // exclusive scan of this digit's counts over (slot, warp) in list order (<= 2048: fits 16 bits)
        const uint32_t run0 = s_running[tid];
        uint32_t local = 0;
#pragma unroll
        for (int w = 0; w < CS_SLOTS * CS_WARPS; w++) {
          const uint32_t c = s_wc[w][tid];
          s_wc[w][tid] = (uint16_t)local;
          local += c;
        }
        s_tilebase[tid] = run0;
        s_running[tid] = run0 + local;
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < CS_SLOTS; j++) {
        if (valid[j]) {
          const uint32_t rank = __popc(peers[j] & ((1u << lane) - 1u));
          const uint32_t pos = s_base[digit[j]] + s_tilebase[digit[j]] + (uint32_t)s_wc[j * CS_WARPS + warp][digit[j]] + rank;
          kout[pos] = key[j];
          vout[pos] = val[j];
          if (more) atomicAdd(&hnext[(pos / a.chunk) * 256 + digit_of(key[j], p + 1)], 1u);
        }
      }
      __syncthreads();   // everyone has read s_wc / s_tilebase before the next tile rewrites them
    }
    grid.sync();
  }
}

size_t coop_sort_hist_bytes() { return (size_t)3 * 256 * COOP_SORT_MAX_BLOCKS * sizeof(uint32_t); }

// Sort n (key, value) pairs on key bits [0, key_bits).  The sorted result ends in keys[passes & 1]
// / vals[passes & 1] with passes = ceil(key_bits / 8); the caller arranges the buffers accordingly.
int coop_sort_pairs(uint64_t* keys0, uint32_t* vals0, uint64_t* keys1, uint32_t* vals1, uint32_t* hist,
                    uint32_t n, int key_bits, cudaStream_t st) {
  if (n == 0) return 0;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 1;
    if (num_sms > COOP_SORT_MAX_BLOCKS) num_sms = COOP_SORT_MAX_BLOCKS;
  }
  CoopSortArgs a;
  a.keys[0] = keys0; a.keys[1] = keys1; a.vals[0] = vals0; a.vals[1] = vals1;
  a.hist = hist;
  a.n = n;
  a.passes = (key_bits + 7) / 8;
  uint32_t grid = (uint32_t)num_sms;
  a.chunk = (n + grid - 1) / grid;
  // keep chunks a multiple of the tile so only the last block has a ragged tile
  a.chunk = (a.chunk + CS_TILE - 1) / CS_TILE * CS_TILE;
  void* args[] = {&a};
  if (check_cuda(cudaLaunchCooperativeKernel((const void*)k_coop_radix_sort, dim3(grid), dim3(CS_THREADS), args, 0, st),
                 "cooperative pair sort"))
    return B200GS_ERR_CUDA;
  count_launch();
  return 0;
}

}  // namespace b200gs

// binning.cu -- the two radix sorts and the scan of the tile-binning stage (CUB device primitives).
//
// Replaces rows a4 InclusiveSum and a6 SortPairs.  The public algorithm sorts D
// (Gaussian,tile) pairs by a 64-bit (tile | depth) key: 6 radix passes over 12-byte pairs.  Here the
// P Gaussians are first sorted by depth (4 passes over P 8-byte pairs, P << D), pairs are emitted in
// that order, and a *stable* sort on the tile id alone (ceil(log2 T) bits -> 2 passes over 8-byte
// pairs) produces exactly the same (tile, depth, index) order.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/permutation_iterator.h>

#include "common.cuh"
#include "kernels.cuh"

namespace b200gs {

size_t depth_sort_temp_bytes(int P) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, P);
  auto it = thrust::make_permutation_iterator((const uint32_t*)nullptr, (const uint32_t*)nullptr);
  cub::DeviceScan::InclusiveSum(nullptr, b, it, (uint32_t*)nullptr, P);
  return a > b ? a : b;
}

size_t tile_sort_temp_bytes(int64_t D, int tile_bits) {
  size_t a = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, D, 0, tile_bits);
  return a;
}

int sort_by_depth_and_scan(const GeomBuf& g, int P, cudaStream_t st) {
  if (P == 0) return 0;
  size_t tb = g.cub_temp_bytes;
  // positive floats order like their bit patterns; culled Gaussians carry 0xFFFFFFFF and sort last
  if (check_cuda(cub::DeviceRadixSort::SortPairs(g.cub_temp, tb, (const uint32_t*)g.depth_key, g.key_sorted,
                                                 (const uint32_t*)g.idx, g.perm, P, 0, 32, st),
                 "depth sort"))
    return B200GS_ERR_CUDA;
  count_launch(5);  // histogram + 4 onesweep passes
  tb = g.cub_temp_bytes;
  auto it = thrust::make_permutation_iterator((const uint32_t*)g.tiles, (const uint32_t*)g.perm);
  if (check_cuda(cub::DeviceScan::InclusiveSum(g.cub_temp, tb, it, g.offsets, P, st), "tile-count scan"))
    return B200GS_ERR_CUDA;
  count_launch(2);
  return 0;
}

int sort_by_tile(const BinBuf& b, int64_t D, int tile_bits, cudaStream_t st) {
  if (D == 0) return 0;
  size_t tb = b.cub_temp_bytes;
  if (check_cuda(cub::DeviceRadixSort::SortPairs(b.cub_temp, tb, (const uint32_t*)b.keys, b.keys_sorted,
                                                 (const uint32_t*)b.vals, b.vals_sorted, D, 0, tile_bits, st),
                 "tile sort"))
    return B200GS_ERR_CUDA;
  count_launch(1 + (tile_bits + 7) / 8);
  return 0;
}

}  // namespace b200gs

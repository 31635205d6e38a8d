// binning.cu -- the scan and the radix sort of the binning stage (CUB device primitives).
//
// Replaces rows a4 InclusiveSum and a6 SortPairs.  Same key as the public algorithm
// -- (bin id << 32 | depth bits), stable LSD radix sort, so ties keep index order -- but the bins are
// (16 << shift)^2 pixels and coverage is the tight ellipse span (project.cu), which cuts the number
// of sorted pairs D by an order of magnitude at 1080p (17.2 M -> ~0.6 M on config C3), and only
// 32 + ceil(log2 bins) key bits are sorted (5 passes at 1080p with 128-px bins).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "kernels.cuh"

namespace b200gs {

size_t scan_temp_bytes(int P) {
  size_t b = 0;
  cub::DeviceScan::InclusiveSum(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, P);
  return b;
}

size_t pair_sort_temp_bytes(int64_t D, int key_bits) {
  size_t a = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, D, 0, key_bits);
  size_t b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, D, 0, 32);
  return a > b ? a : b;
}

int scan_bin_counts(const GeomBuf& g, int P, cudaStream_t st) {
  if (P == 0) return 0;
  size_t tb = g.cub_temp_bytes;
  if (check_cuda(cub::DeviceScan::InclusiveSum(g.cub_temp, tb, (const uint32_t*)g.tiles, g.offsets, P, st),
                 "bin-count scan"))
    return B200GS_ERR_CUDA;
  count_launch(2);
  return 0;
}

// Stable LSD radix sort of the (bin << 32 | depth bits) keys; ties keep emission (= index) order.
int sort_pairs(const BinBuf& b, int64_t D, int key_bits, cudaStream_t st) {
  if (D == 0) return 0;
  size_t tb = b.cub_temp_bytes;
  if (check_cuda(cub::DeviceRadixSort::SortPairs(b.cub_temp, tb, (const uint64_t*)b.keys, b.keys_sorted,
                                                 (const uint32_t*)b.vals, b.vals_sorted, D, 0, key_bits, st),
                 "pair sort"))
    return B200GS_ERR_CUDA;
  count_launch(1 + (key_bits + 7) / 8);
  return 0;
}

// 32-bit keys (bin << 24 | quantised depth) stored in the first halves of the 64-bit key arrays: four passes.
int sort_pairs32(const BinBuf& b, int64_t D, int key_bits, cudaStream_t st) {
  if (D == 0) return 0;
  size_t tb = b.cub_temp_bytes;
  if (check_cuda(cub::DeviceRadixSort::SortPairs(b.cub_temp, tb, reinterpret_cast<const uint32_t*>(b.keys),
                                                 reinterpret_cast<uint32_t*>(b.keys_sorted), (const uint32_t*)b.vals,
                                                 b.vals_sorted, D, 0, key_bits, st),
                 "pair sort (32-bit keys)"))
    return B200GS_ERR_CUDA;
  count_launch(1 + (key_bits + 7) / 8);
  return 0;
}

}  // namespace b200gs

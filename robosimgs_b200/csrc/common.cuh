// common.cuh -- shared device helpers, buffer carving and PTX wrappers for libb200gs (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/b200gs.h"

namespace b200gs {

constexpr int TILE = 16;            // tile edge in pixels (fixed: parity of the 3-sigma rect culling)
constexpr int REC_F4 = 3;           // float4s per projected-splat record (48 B)
constexpr int GRAD2D_STRIDE = 12;   // floats per Gaussian in the screen-space gradient accumulator

// ---- projected-splat record (48 B, three float4) ----------------------------------------------
//   q0 = { x, y, A, B }            pixel centre, conic
//   q1 = { C, opacity, thr, idx }  thr = ln(255*opacity) + slack: 0.5*q <= thr  <=>  alpha >= 1/255
//   q2 = { r, g, b, +-radius }     radius as a float; NEGATIVE marks a record that needs the general evaluation in
//                                  the compositing kernels (opacity > 0.99: the min(0.99, .) clamp can bind; or a
//                                  conic that is not safely positive definite: `power` may round above 0)
// One record per Gaussian (geom buffer, 16-byte aligned): the unit a single 48-byte TMA bulk copy
// gathers into a compositing CTA's shared-memory ring.

// ---- buffer carving (128-byte aligned chunks inside caller-owned byte buffers) -----------------
struct Carver {
  char* p;
  size_t off;
  __host__ explicit Carver(char* base) : p(base), off(0) {}
  template <typename T>
  __host__ T* take(size_t count) {
    off = (off + 127) & ~size_t(127);
    T* r = p ? reinterpret_cast<T*>(p + off) : nullptr;
    off += count * sizeof(T);
    return r;
  }
  __host__ size_t bytes() const { return (off + 127) & ~size_t(127); }
};

struct GeomBuf {
  float4* rec;          // [P*3]
  uint32_t* depth_key;  // [P] IEEE bits of the view-space depth (positive floats order like uints)
  uint32_t* big_queue;  // [P] ids of large-footprint Gaussians (emitted one warp each)
  uint32_t* tiles;      // [P] bins touched (tight count)
  uint32_t* offsets;    // [P] inclusive scan of tiles[]
  uint8_t* clamped;     // [P] bit c set: colour channel c was clamped at 0
  uint32_t* counters;   // [32] device-side scalars (large-footprint queue length, ...)
  char* cub_temp;
  size_t cub_temp_bytes;
};

struct BinBuf {
  uint32_t* vals_sorted; // [D] Gaussian ids in (tile, depth) order -- FIRST chunk: the only part the
                         // backward pass reads, so its offset must not depend on the capacity
  uint64_t* keys_sorted; // [D]
  uint64_t* keys;        // [D] (bin << 32) | depth bits, emission (index) order
  uint32_t* vals;        // [D] Gaussian ids
  uint32_t* coop_hist;   // rotating digit histograms of the cooperative sort
  uint32_t* win_first;   // [D / BUCKET_WINDOW + BUCKET_BINS_MAX + 2] bucketed binning: first bucket of every sort window
  uint2* big_segs;       // [D / 512 + 2] bucketed binning: segments queued for the one-CTA sort
  float4* slab;          // [D * 3] records in list order (option "gather" = 2), LAST chunk: absent otherwise
  char* cub_temp;
  size_t cub_temp_bytes;
};

struct ImgBuf {
  uint2* ranges;        // [tiles] [start,end) into the slab
  float4* pix;          // [H*W] {accumulated r,g,b (no background), final transmittance}
  uint32_t* n_contrib;  // [H*W] list position behind which nothing contributes to the pixel
  // bucketed binning (bucket.cu): fixed-size tables, so the layout never depends on the bin size
  unsigned long long* bin_pub;   // [BUCKET_BINS_MAX] look-back words of the bucket scan -- directly in front of
  uint32_t* bucket_count;   // [BUCKETS_MAX] pairs per (bin, depth slice) bucket (one memset clears both)
  uint32_t* bucket_base;    // [BUCKETS_MAX + 4] exclusive scan (+ total)
  uint32_t* bucket_cursor;  // [BUCKETS_MAX]
};

// ---- small math helpers ---------------------------------------------------------------------
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(hi, fmaxf(lo, v)); }
// exp(x) for x <= 0 as one FMUL + one MUFU.EX2 (2 ulp; results below 2^-126 flush to 0, far below
// the 1/255 alpha threshold they are compared against)
__device__ __forceinline__ float exp_fast(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
  return r;
}
// 1/x for normal x as one MUFU.RCP (1 ulp)
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ---- mbarrier / TMA (1-D bulk copy) PTX wrappers ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk async copy (TMA, SASS UBLKCP); bytes multiple of 16, both 16-B aligned.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- launch accounting / error plumbing (host) -------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_cuda(cudaError_t e, const char* what);
int debug_sync(const B200GSParams* prm, cudaStream_t s, const char* what);

}  // namespace b200gs

// render4.cuh -- pieces shared by the four-pixels-per-thread compositing kernels (render4.cu: one survivor
// list per 16x8 warp block; render4q.cu: one list per 8x4 quarter of it).
#pragma once
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"

namespace b200gs {


constexpr int R4_CH = 64;       // records per ring stage == threads per CTA
#ifndef R4_STAGES_N
#define R4_STAGES_N 2
#endif
constexpr int R4_STAGES = R4_STAGES_N;
constexpr int R4_THREADS = 64;
constexpr float LOG2E = 1.4426950408889634f;

// identical to render.cu (kept local: both translation units inline it)
__device__ __forceinline__ bool r4_block_may_contribute(float x, float y, float A, float B, float C, float thr,
                                                        float x0, float y0, float x1, float y1) {
  const float cx = clampf(x, x0, x1), cy = clampf(y, y0, y1);
  const float dxe = cx - x, dye = cy - y;
  float dy1 = clampf(y - B * dxe * rcp_fast(C), y0, y1) - y;
  const float q1 = A * dxe * dxe + 2.f * B * dxe * dy1 + C * dy1 * dy1;
  float dx2 = clampf(x - B * dye * rcp_fast(A), x0, x1) - x;
  const float q2 = A * dx2 * dx2 + 2.f * B * dx2 * dye + C * dye * dye;
  const float q = fminf(q1, q2);
  const float mag = fabsf(A) * (dxe * dxe + dx2 * dx2) + fabsf(C) * (dye * dye + dy1 * dy1);
  return 0.5f * q - 4e-6f * mag <= thr;
}

// The same test on a record whose conic was pre-scaled for the exp2 evaluation (r4_prescale): hA = A log2(e)/2,
// Bs = B log2(e), hC = C log2(e)/2, thr2 = thr log2(e).  0.5 q log2(e) = hA dx^2 + Bs dx dy + hC dy^2, B/C = Bs/(2 hC).
__device__ __forceinline__ bool r4_block_may_contribute_scaled(float x, float y, float hA, float Bs, float hC, float thr2,
                                                               float x0, float y0, float x1, float y1) {
  const float cx = clampf(x, x0, x1), cy = clampf(y, y0, y1);
  const float dxe = cx - x, dye = cy - y;
  float dy1 = clampf(y - 0.5f * Bs * dxe * rcp_fast(hC), y0, y1) - y;
  const float q1 = hA * dxe * dxe + Bs * dxe * dy1 + hC * dy1 * dy1;
  float dx2 = clampf(x - 0.5f * Bs * dye * rcp_fast(hA), x0, x1) - x;
  const float q2 = hA * dx2 * dx2 + Bs * dx2 * dye + hC * dye * dye;
  const float q = fminf(q1, q2);
  const float mag = fabsf(hA) * (dxe * dxe + dx2 * dx2) + fabsf(hC) * (dye * dye + dy1 * dy1);
  return q - 8e-6f * mag <= thr2;
}

// One thread per record of a freshly landed ring stage: scale the conic (and the culling threshold) in place, ONCE,
// with exactly the products row_terms() used to form per record and per warp -- so the evaluation below is unchanged
// bit for bit and three multiplies per record leave the hot loop.  The caller synchronises the CTA afterwards.
__device__ __forceinline__ void r4_prescale(float4* st, uint32_t cnt, int tid) {
  if ((uint32_t)tid < cnt) {
    float4 q0 = st[tid * REC_F4], q1 = st[tid * REC_F4 + 1];
    q0.z = __fmul_rn(q0.z, 0.5f * LOG2E);
    q0.w = __fmul_rn(q0.w, LOG2E);
    q1.x = __fmul_rn(q1.x, 0.5f * LOG2E);
    q1.z = __fmul_rn(q1.z, LOG2E);
    st[tid * REC_F4] = q0;
    st[tid * REC_F4 + 1] = q1;
  }
}

__device__ __forceinline__ bool r4_tile_in_reference_rect(float px, float py, float fr, int tx, int ty, int gx,
                                                          int gy) {
  const int x0 = min(gx, max(0, (int)((px - fr) / TILE)));
  const int y0 = min(gy, max(0, (int)((py - fr) / TILE)));
  const int x1 = min(gx, max(0, (int)((px + fr + (TILE - 1)) / TILE)));
  const int y1 = min(gy, max(0, (int)((py + fr + (TILE - 1)) / TILE)));
  return tx >= x0 && tx < x1 && ty >= y0 && ty < y1;
}

__device__ __forceinline__ void r4_cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void r4_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void r4_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float ex2_fast(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Ring of gathered records: thread t owns slot t of every stage; the Gaussian id of the next fill
// is prefetched one chunk ahead.  One cp.async group is committed per issue() (possibly empty), so
// "chunk c resident" == "at most R4_STAGES-1 groups pending".
struct Ring4 {
  float4 (*sm)[R4_CH * REC_F4];
  const uint32_t* list;
  const float4* rec;
  uint32_t n, nchunks;
  uint32_t next_id;
  int tid;

  __device__ __forceinline__ uint32_t count(uint32_t c) const { return min((uint32_t)R4_CH, n - c * R4_CH); }
  __device__ __forceinline__ uint32_t load_id(uint32_t c) const {
    return (c < nchunks && (uint32_t)tid < count(c)) ? __ldg(list + c * R4_CH + tid) : 0u;
  }
  __device__ __forceinline__ void issue(uint32_t c, uint32_t id) {
    if (c < nchunks && (uint32_t)tid < count(c)) {
      const float4* src = rec + (size_t)id * REC_F4;
      float4* dst = &sm[c % R4_STAGES][tid * REC_F4];
      r4_cp_async16(dst, src);
      r4_cp_async16(dst + 1, src + 1);
      r4_cp_async16(dst + 2, src + 2);
    }
    r4_cp_async_commit();
  }
  __device__ __forceinline__ void prologue() {
    uint32_t ids[R4_STAGES];
#pragma unroll
    for (int s = 0; s < R4_STAGES; s++) ids[s] = load_id(s);
#pragma unroll
    for (int s = 0; s < R4_STAGES; s++) issue(s, ids[s]);
    next_id = load_id(R4_STAGES);
  }
  __device__ __forceinline__ void wait(uint32_t) {
    r4_cp_async_wait<R4_STAGES - 1>();
    __syncthreads();
  }
  __device__ __forceinline__ void drain() { r4_cp_async_wait<0>(); }   // no copy may still target this CTA's smem
  // stage c % R4_STAGES is free (caller synchronised the CTA): refill it with chunk c + R4_STAGES
  __device__ __forceinline__ void refill(uint32_t c) {
    issue(c + R4_STAGES, next_id);
    next_id = load_id(c + R4_STAGES + 1);
  }
};

// Same ring fed by TMA from the bin-ordered record SLAB (the bucket sort writes rec[point_list[i]] to slab[i], so a
// tile's list is one contiguous run of 48-byte records): thread 0 arms the stage's mbarrier with the
// chunk's byte count and issues ONE cp.async.bulk of up to 64 records (3 KB); everybody waits on the barrier's phase.
struct Ring4Slab {
  float4 (*sm)[R4_CH * REC_F4];
  uint64_t* bar;            // [R4_STAGES] in shared memory
  const float4* slab;       // first record of this tile's list
  uint32_t n, nchunks;
  int tid;

  __device__ __forceinline__ uint32_t count(uint32_t c) const { return min((uint32_t)R4_CH, n - c * R4_CH); }
  __device__ __forceinline__ void issue(uint32_t c) {
    if (tid == 0 && c < nchunks) {
      const uint32_t bytes = count(c) * (uint32_t)(REC_F4 * sizeof(float4));
      // the stage was written through the generic proxy (r4_prescale): order those writes before the bulk copy
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bar[c % R4_STAGES], bytes);
      tma_load_1d(&sm[c % R4_STAGES][0], slab + (size_t)c * R4_CH * REC_F4, bytes, &bar[c % R4_STAGES]);
    }
  }
  __device__ __forceinline__ void prologue() {
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < R4_STAGES; s++) mbar_init(&bar[s], 1);
      mbar_fence_init();
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < R4_STAGES; s++) issue(s);
  }
  // chunk c is the (c / R4_STAGES)-th use of its stage: phase parity alternates per use
  __device__ __forceinline__ void wait(uint32_t c) { mbar_wait(&bar[c % R4_STAGES], (c / R4_STAGES) & 1u); }
  // stage c % R4_STAGES is free (caller synchronised the CTA): refill it with chunk c + R4_STAGES
  __device__ __forceinline__ void refill(uint32_t c) { issue(c + R4_STAGES); }
  __device__ __forceinline__ void drain() {}
};

// Per-record terms shared by the four pixels of a thread (a row): with the conic scaled by
// log2(e), power*log2(e) = -dx*(hA*dx + B*dy) - hC*dy^2.  Explicitly rounded so the forward and the
// adjoint evaluate bit-identical alphas (the adjoint re-derives which records contributed).
struct RowTerms {
  float hA, bdy, cdy2;
};
// from a PRE-SCALED record (r4_prescale): hA = A log2(e)/2 as stored, bdy = (B log2(e)) dy, cdy2 = ((C log2(e)/2) dy) dy
__device__ __forceinline__ RowTerms row_terms(float hA, float Bs, float hC, float dy) {
  RowTerms r;
  r.hA = hA;
  r.bdy = __fmul_rn(Bs, dy);
  r.cdy2 = __fmul_rn(__fmul_rn(hC, dy), dy);
  return r;
}
__device__ __forceinline__ float power2_of(const RowTerms& r, float dx) {
  return __fmaf_rn(-dx, __fmaf_rn(r.hA, dx, r.bdy), -r.cdy2);
}


}  // namespace b200gs

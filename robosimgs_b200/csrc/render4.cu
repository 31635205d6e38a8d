// render4.cu -- compositing (K6) and its adjoint (K7) with FOUR pixels per thread.
//
// Same contract, same traversal, same culling rules and the same scratch layout as render.cu
// (arithmetic of oracle/gs_oracle_impl.h gso_render /
// gso_render_backward), re-shaped around the measured bound of those kernels: FP32 issue
// (86 % issue-active, 0.7 % of HBM peak -- profiles/r1_ncu_full_summary_d.csv).  On the opaque C3
// frame a tile walks only ~75 records of its bin list and ~56 of them touch the tile, nearly all of
// them covering every pixel of the block that evaluates them (tests/devtools/sim/run_sim.py), so three
// quarters of the issue slots of the one-pixel-per-thread kernel are the per-record loop itself
// (mask bookkeeping, three shared-memory broadcasts, dy/conic products that are identical for the
// pixels of a row) -- work that does not grow with the number of pixels a thread owns.
//
//  * one CTA of 64 threads per 16x16 tile; a warp owns a 16x8 block, a thread owns 4 horizontally
//    adjacent pixels: the row terms (dy, B dy, C dy^2) and the record loads are shared, each pixel
//    costs one subtraction and two FMAs up to `power`; pixel state lives in registers and leaves
//    the SM as 128-bit stores;
//  * the conic is pre-scaled by log2(e) per record, so alpha is one MUFU.EX2 away from `power`;
//  * 64-record ring stages (one 48-byte record per thread, three 16-byte LDGSTS), R4_STAGES deep;
//  * adjoint: the nine per-Gaussian sums are first accumulated over a thread's four pixels in
//    registers (the y-moments factor out of the row), then reduced over the warp once per record
//    with the transposing butterfly of render.cu -- one reduction per 128 pixels instead of per 32.
#include <atomic>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"
#include "render4.cuh"

namespace b200gs {

struct Tile4 {
  int px0, py;          // first of the thread's four pixels
  float pxf[4], pyf;
  float rx0, ry0, rx1, ry1;   // the warp's 16x8 block, clipped to the image
  bool in[4];
};
__device__ __forceinline__ Tile4 tile4_setup(int W, int H, int warp, int lane) {
  Tile4 t;
  const int bx = blockIdx.x * TILE, by = blockIdx.y * TILE + warp * 8;
  t.px0 = bx + 4 * (lane & 3);
  t.py = by + (lane >> 2);
  t.pyf = (float)t.py;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    t.pxf[i] = (float)(t.px0 + i);
    t.in[i] = (t.px0 + i) < W && t.py < H;
  }
  t.rx0 = (float)bx; t.ry0 = (float)by;
  t.rx1 = (float)min(bx + 15, W - 1); t.ry1 = (float)min(by + 7, H - 1);
  return t;
}

// ==================================================================================================
// K6 (four pixels per thread)
// ==================================================================================================
#ifndef R4_FWD_MINB
#define R4_FWD_MINB 16
#endif
#ifndef R4_BWD_MINB
#define R4_BWD_MINB 12
#endif
__global__ void __launch_bounds__(R4_THREADS, R4_FWD_MINB) k_render_fwd4(RenderArgs a) {
  __shared__ __align__(128) float4 sm[R4_STAGES][R4_CH * REC_F4];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint2 range = a.ranges[(blockIdx.y >> a.bin_shift) * a.gbx + (blockIdx.x >> a.bin_shift)];
  const bool coarse = a.bin_shift != 0;
  const uint32_t n = range.y - range.x;
  const uint32_t nchunks = (n + R4_CH - 1) / R4_CH;
  Ring4 ring{sm, a.point_list + range.x, a.rec, n, nchunks, 0u, tid};
  const Tile4 t = tile4_setup(a.W, a.H, warp, lane);

  ring.prologue();

  // A finished pixel (transmittance test failed once, or outside the image) is marked by a NEGATIVE
  // transmittance: |T| stays the final transmittance, T * (1 - alpha) can never pass the 1e-4 test
  // again, so the hot loop needs no separate flag.  nstop = list position of the record that
  // finished the pixel (n if none did): nothing at or behind it contributes, and every record in
  // front of it that the adjoint finds valid did contribute -- all the adjoint needs to know.
  float T[4], Cr[4], Cg[4], Cb[4];
  uint32_t nstop[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    T[i] = t.in[i] ? 1.f : -1.f;
    Cr[i] = 0.f; Cg[i] = 0.f; Cb[i] = 0.f; nstop[i] = n;
  }

  // One record against the thread's four pixels.  CLAMP = false when opacity <= 0.99: then
  // min(0.99, o*G) is the identity for every pair that can pass the tests (G <= 1 when power <= 0).
  // The accumulate is predicated, not selected: FSETP/FSEL/FMNMX share the half-rate ALU pipe, which
  // co-limits this loop with the issue rate (ncu: math-pipe throttle).
  auto eval = [&](const float4& q0, const float4& q1, const float4& q2, uint32_t pos, auto clamp_tag) {
    constexpr bool CLAMP = decltype(clamp_tag)::value;
    const RowTerms rt = row_terms(q0.z, q0.w, q1.x, q0.y - t.pyf);
    bool stop[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float dx = q0.x - t.pxf[i];
      const float p2 = power2_of(rt, dx);
      float alpha = q1.y * ex2_fast(p2);
      if (CLAMP) alpha = fminf(0.99f, alpha);
      const bool valid = p2 <= 0.f && alpha >= (1.f / 255.f);
      const float test_T = T[i] * (1.f - alpha);
      const bool upd = valid && test_T >= 0.0001f;
      stop[i] = valid && !(test_T >= 0.0001f);
      if (upd) {
        const float w = alpha * T[i];
        Cr[i] = __fmaf_rn(q2.x, w, Cr[i]);
        Cg[i] = __fmaf_rn(q2.y, w, Cg[i]);
        Cb[i] = __fmaf_rn(q2.z, w, Cb[i]);
        T[i] = test_T;
      }
    }
    if (stop[0] || stop[1] || stop[2] || stop[3]) {   // rare: at most once per pixel
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (stop[i] && T[i] > 0.f) { T[i] = -T[i]; nstop[i] = pos; }
    }
  };

  for (uint32_t c = 0; c < nchunks; c++) {
    ring.wait();
    const uint32_t cnt = ring.count(c);
    const float4* st = sm[c % R4_STAGES];
    for (uint32_t base = 0; base < cnt; base += 32) {
      if (__all_sync(0xffffffffu, fmaxf(fmaxf(T[0], T[1]), fmaxf(T[2], T[3])) < 0.f)) break;
      const uint32_t j = base + lane;
      bool hit = false;
      if (j < cnt) {
        const float4 q0 = st[j * REC_F4], q1 = st[j * REC_F4 + 1];
        hit = r4_block_may_contribute(q0.x, q0.y, q0.z, q0.w, q1.x, q1.z, t.rx0, t.ry0, t.rx1, t.ry1);
        if (coarse && hit)
          hit = r4_tile_in_reference_rect(q0.x, q0.y, st[j * REC_F4 + 2].w, blockIdx.x, blockIdx.y, gridDim.x,
                                          gridDim.y);
      }
      uint32_t mask = __ballot_sync(0xffffffffu, hit);
      while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t jj = base + b;
        const float4 q0 = st[jj * REC_F4], q1 = st[jj * REC_F4 + 1], q2 = st[jj * REC_F4 + 2];
        const uint32_t pos = c * R4_CH + jj;
        if (q1.y > 0.99f) eval(q0, q1, q2, pos, std::true_type{});
        else eval(q0, q1, q2, pos, std::false_type{});
      }
    }
    const int num_done = __syncthreads_count(fmaxf(fmaxf(T[0], T[1]), fmaxf(T[2], T[3])) < 0.f);
    if (num_done == R4_THREADS) break;
    ring.refill(c);
  }
  r4_cp_async_wait<0>();   // no asynchronous copy may still target this CTA's shared memory

#pragma unroll
  for (int i = 0; i < 4; i++) T[i] = fabsf(T[i]);
  const size_t hw = (size_t)a.H * a.W;
  const size_t pid = (size_t)t.py * a.W + t.px0;
  if (a.vec4 && t.in[3]) {
    float4* pix = a.pix + pid;
#pragma unroll
    for (int i = 0; i < 4; i++) pix[i] = make_float4(Cr[i], Cg[i], Cb[i], T[i]);
    *reinterpret_cast<uint4*>(a.n_contrib + pid) = make_uint4(nstop[0], nstop[1], nstop[2], nstop[3]);
    const float b0 = __ldg(a.bg + 0), b1 = __ldg(a.bg + 1), b2 = __ldg(a.bg + 2);
    *reinterpret_cast<float4*>(a.out_color + pid) =
        make_float4(__fmaf_rn(T[0], b0, Cr[0]), __fmaf_rn(T[1], b0, Cr[1]), __fmaf_rn(T[2], b0, Cr[2]),
                    __fmaf_rn(T[3], b0, Cr[3]));
    *reinterpret_cast<float4*>(a.out_color + hw + pid) =
        make_float4(__fmaf_rn(T[0], b1, Cg[0]), __fmaf_rn(T[1], b1, Cg[1]), __fmaf_rn(T[2], b1, Cg[2]),
                    __fmaf_rn(T[3], b1, Cg[3]));
    *reinterpret_cast<float4*>(a.out_color + 2 * hw + pid) =
        make_float4(__fmaf_rn(T[0], b2, Cb[0]), __fmaf_rn(T[1], b2, Cb[1]), __fmaf_rn(T[2], b2, Cb[2]),
                    __fmaf_rn(T[3], b2, Cb[3]));
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (!t.in[i]) continue;
      a.pix[pid + i] = make_float4(Cr[i], Cg[i], Cb[i], T[i]);
      a.n_contrib[pid + i] = nstop[i];
      a.out_color[pid + i] = __fmaf_rn(T[i], __ldg(a.bg + 0), Cr[i]);
      a.out_color[hw + pid + i] = __fmaf_rn(T[i], __ldg(a.bg + 1), Cg[i]);
      a.out_color[2 * hw + pid + i] = __fmaf_rn(T[i], __ldg(a.bg + 2), Cb[i]);
    }
  }
}

// identical to render.cu:warp_reduce9
__device__ __forceinline__ void r4_warp_reduce9(float (&v)[8], float& v8, int lane) {
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
  float r4[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float keep = h16 ? v[i + 4] : v[i];
    const float send = h16 ? v[i] : v[i + 4];
    r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float r2[2];
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const float keep = h8 ? r4[i + 2] : r4[i];
    const float send = h8 ? r4[i] : r4[i + 2];
    r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const float keep = h4 ? r2[1] : r2[0];
    const float send = h4 ? r2[0] : r2[1];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v8 += __shfl_xor_sync(0xffffffffu, v8, o);
}

// ==================================================================================================
// K7 (four pixels per thread).  Front-to-back replay, see render.cu:k_render_bwd for the algebra:
//   dL/dalpha_j = T_j <c_j, g> - (F - R_j) / (1 - alpha_j),  F = <C_final, g> + T_final <bg, g>,
//   R_j = sum_{k<=j} w_k <c_k, g>.
// ==================================================================================================
__global__ void __launch_bounds__(R4_THREADS, R4_BWD_MINB) k_render_bwd4(RenderBwdArgs a) {
  __shared__ __align__(128) float4 sm[R4_STAGES][R4_CH * REC_F4];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint2 range = a.ranges[(blockIdx.y >> a.bin_shift) * a.gbx + (blockIdx.x >> a.bin_shift)];
  const bool coarse = a.bin_shift != 0;
  const uint32_t n = range.y - range.x;
  const uint32_t nchunks = (n + R4_CH - 1) / R4_CH;
  Ring4 ring{sm, a.point_list + range.x, a.rec, n, nchunks, 0u, tid};
  const Tile4 t = tile4_setup(a.W, a.H, warp, lane);

  ring.prologue();

  float gr[4], gg[4], gb[4], F[4], T[4], R[4];
  uint32_t nc[4];
  {
    const size_t hw = (size_t)a.H * a.W;
    const size_t pid = (size_t)t.py * a.W + t.px0;
    float4 fin[4];
    if (a.vec4 && t.in[3]) {
#pragma unroll
      for (int i = 0; i < 4; i++) fin[i] = a.pix[pid + i];
      const uint4 n4 = *reinterpret_cast<const uint4*>(a.n_contrib + pid);
      nc[0] = n4.x; nc[1] = n4.y; nc[2] = n4.z; nc[3] = n4.w;
      const float4 r4 = __ldg(reinterpret_cast<const float4*>(a.dL_dpix + pid));
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.dL_dpix + hw + pid));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.dL_dpix + 2 * hw + pid));
      gr[0] = r4.x; gr[1] = r4.y; gr[2] = r4.z; gr[3] = r4.w;
      gg[0] = g4.x; gg[1] = g4.y; gg[2] = g4.z; gg[3] = g4.w;
      gb[0] = b4.x; gb[1] = b4.y; gb[2] = b4.z; gb[3] = b4.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        fin[i] = make_float4(0.f, 0.f, 0.f, 1.f);
        nc[i] = 0u; gr[i] = 0.f; gg[i] = 0.f; gb[i] = 0.f;
        if (t.in[i]) {
          fin[i] = a.pix[pid + i];
          nc[i] = a.n_contrib[pid + i];
          gr[i] = __ldg(a.dL_dpix + pid + i);
          gg[i] = __ldg(a.dL_dpix + hw + pid + i);
          gb[i] = __ldg(a.dL_dpix + 2 * hw + pid + i);
        }
      }
    }
    const float b0 = __ldg(a.bg), b1 = __ldg(a.bg + 1), b2 = __ldg(a.bg + 2);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      F[i] = fin[i].x * gr[i] + fin[i].y * gg[i] + fin[i].z * gb[i] + fin[i].w * (b0 * gr[i] + b1 * gg[i] + b2 * gb[i]);
      T[i] = 1.f; R[i] = 0.f;
    }
  }
  const uint32_t ncmax = max(max(nc[0], nc[1]), max(nc[2], nc[3]));

  for (uint32_t c = 0; c < nchunks; c++) {
    ring.wait();
    const uint32_t cnt = ring.count(c);
    const float4* st = sm[c % R4_STAGES];
    for (uint32_t base = 0; base < cnt; base += 32) {
      if (__all_sync(0xffffffffu, c * R4_CH + base >= ncmax)) break;
      const uint32_t j = base + lane;
      bool hit = false;
      if (j < cnt) {
        const float4 q0 = st[j * REC_F4], q1 = st[j * REC_F4 + 1];
        hit = r4_block_may_contribute(q0.x, q0.y, q0.z, q0.w, q1.x, q1.z, t.rx0, t.ry0, t.rx1, t.ry1);
        if (coarse && hit)
          hit = r4_tile_in_reference_rect(q0.x, q0.y, st[j * REC_F4 + 2].w, blockIdx.x, blockIdx.y, gridDim.x,
                                          gridDim.y);
      }
      uint32_t mask = __ballot_sync(0xffffffffu, hit);
      while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t jj = base + b;
        const float4 q0 = st[jj * REC_F4], q1 = st[jj * REC_F4 + 1], q2 = st[jj * REC_F4 + 2];
        const float dy = q0.y - t.pyf;
        const RowTerms rt = row_terms(q0.z, q0.w, q1.x, dy);
        const uint32_t pos = c * R4_CH + jj;
        float dx[4], G[4], alpha[4];
        bool valid[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          dx[i] = q0.x - t.pxf[i];
          const float p2 = power2_of(rt, dx[i]);
          G[i] = ex2_fast(p2);
          alpha[i] = fminf(0.99f, q1.y * G[i]);
          valid[i] = pos < nc[i] && p2 <= 0.f && alpha[i] >= (1.f / 255.f);
        }
        if (!__any_sync(0xffffffffu, valid[0] || valid[1] || valid[2] || valid[3])) continue;
        float vr = 0.f, vg = 0.f, vb = 0.f, s0 = 0.f, sx = 0.f, sxx = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          // branch-free: a record that does not contribute to this pixel acts with alpha = G = 0
          const float al = valid[i] ? alpha[i] : 0.f;
          const float Gi = valid[i] ? G[i] : 0.f;
          const float w = al * T[i];
          const float cg = q2.x * gr[i] + q2.y * gg[i] + q2.z * gb[i];
          R[i] = __fmaf_rn(w, cg, R[i]);
          const float one_m = 1.f - al;
          const float dL_dalpha = T[i] * cg - rcp_fast(one_m) * (F[i] - R[i]);
          T[i] = T[i] * one_m;
          const float m = Gi * dL_dalpha;
          vr = __fmaf_rn(w, gr[i], vr);
          vg = __fmaf_rn(w, gg[i], vg);
          vb = __fmaf_rn(w, gb[i], vb);
          s0 += m;
          const float mx = m * dx[i];
          sx += mx;
          sxx = __fmaf_rn(mx, dx[i], sxx);
        }
        // the four pixels share dy: Sy = dy S0, Sxy = dy Sx, Syy = dy^2 S0
        float v[8], v8;
        v[0] = vr; v[1] = vg; v[2] = vb;
        v[3] = s0; v[4] = sx; v[5] = s0 * dy;
        v[6] = sxx; v[7] = sx * dy;
        v8 = s0 * dy * dy;
        r4_warp_reduce9(v, v8, lane);
        const uint32_t g = __float_as_uint(q1.w);
        const bool writer = ((lane & 3) == 0) || lane == 1;
        if (writer) {
          const int k = (lane == 1) ? 8 : (lane >> 2);
          atomicAdd(a.grad2d + (size_t)g * GRAD2D_STRIDE + k, (lane == 1) ? v8 : v[0]);
        }
      }
    }
    const int num_done = __syncthreads_count(c * R4_CH + cnt >= ncmax);
    if (num_done == R4_THREADS) break;
    ring.refill(c);
  }
  r4_cp_async_wait<0>();
}

void launch_render4(const RenderArgs& a, cudaStream_t st) {
  const dim3 grid((a.W + TILE - 1) / TILE, (a.H + TILE - 1) / TILE), block(R4_THREADS);
  k_render_fwd4<<<grid, block, 0, st>>>(a);
  count_launch();
}

void launch_render_bwd4(const RenderBwdArgs& a, cudaStream_t st) {
  const dim3 grid((a.W + TILE - 1) / TILE, (a.H + TILE - 1) / TILE), block(R4_THREADS);
  k_render_bwd4<<<grid, block, 0, st>>>(a);
  count_launch();
}

}  // namespace b200gs

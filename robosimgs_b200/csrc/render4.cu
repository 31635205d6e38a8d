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
#ifndef R4_FWD_PAIRS
#define R4_FWD_PAIRS 0
#endif
template <bool SLAB>
__global__ void __launch_bounds__(R4_THREADS, R4_FWD_MINB) k_render_fwd4(RenderArgs a) {
  __shared__ __align__(128) float4 sm[R4_STAGES][R4_CH * REC_F4];
  __shared__ __align__(8) uint64_t s_bar[R4_STAGES];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint2 range = a.ranges[(blockIdx.y >> a.bin_shift) * a.gbx + (blockIdx.x >> a.bin_shift)];
  const bool coarse = a.bin_shift != 0;
  const uint32_t n = range.y - range.x;
  const uint32_t nchunks = (n + R4_CH - 1) / R4_CH;
  using Ring = std::conditional_t<SLAB, Ring4Slab, Ring4>;
  Ring ring = [&]() {
    if constexpr (SLAB) return Ring4Slab{sm, s_bar, a.slab + (size_t)range.x * REC_F4, n, nchunks, tid};
    else return Ring4{sm, a.point_list + range.x, a.rec, n, nchunks, 0u, tid};
  }();
  const Tile4 t = tile4_setup(a.W, a.H, warp, lane);

  ring.prologue();

  // A finished pixel (transmittance test failed once, or outside the image) is marked by an INFINITE x coordinate:
  // for every later record ndx = +inf, power2 = fma(inf, fma(-hA, inf, bdy), -cdy2) = -inf (hA > 0 for unmarked
  // records; marked ones skip a NaN/+inf power explicitly), so its alpha is exactly 0 and the record acts on the
  // pixel as the identity -- T keeps the final transmittance, the hot loop needs no flag and no predicate for it.
  // nstop = list position of the record that finished the pixel (n if none did): nothing at or behind it
  // contributes, and every record in front of it that the adjoint finds valid did contribute -- all the adjoint
  // needs to know.
  // Pixel state lives in register PAIRS (pixels 0|1 and 2|3) and the whole evaluation is issued as packed f32x2
  // instructions (FADD2 / FFMA2 / FMUL2, sm_100) -- one issue slot for two pixels.  Every packed
  // operation rounds exactly like the scalar expression it replaces (and like render.cu's one-pixel kernels):
  //   ndx = px - x (= -dx);  power2 = fma(ndx, fma(-hA, ndx, bdy), -cdy2);  1 - alpha = fma(alpha, -1, 1).
  const float FIN = __int_as_float(0x7f800000);
  float2 T[2], Cr[2], Cg[2], Cb[2], pxf2[2];
  uint32_t nstop[4];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    T[h] = make_float2(1.f, 1.f);
    Cr[h] = Cg[h] = Cb[h] = make_float2(0.f, 0.f);
    pxf2[h] = make_float2(t.in[2 * h] ? t.pxf[2 * h] : FIN, t.in[2 * h + 1] ? t.pxf[2 * h + 1] : FIN);
    nstop[2 * h] = n; nstop[2 * h + 1] = n;
  }

  // One record against the thread's four pixels.  A pixel the record does not reach (alpha < 1/255; for marked
  // records also power > 0) takes alpha = 0: w = 0, C + c*0 = C and T*(1 - 0) = T are exact, so the accumulate is
  // unconditional and packed.  The only data-dependent control flow is the once-per-pixel "transmittance exhausted"
  // event (T*(1 - alpha) < 1e-4 with T >= 1e-4 needs alpha > 0.99 * ... i.e. a record that does reach the pixel):
  // then the record is applied pixel by pixel and the exhausted ones are retired.
  // GENERAL = the record is marked (negative radius, project.cu:record_is_general): clamp alpha at 0.99 and skip
  // pixels whose power rounds above 0.  Unmarked records can do neither, so both tests are dropped for them.
  // every pixel of this thread is retired (or outside the image): only changes where a pixel retires, so the warp can
  // leave its list behind the very record that exhausted its last pixel for the price of one vote per record --
  // at batch granularity (32 list positions) half a batch of evaluations per tile was spent on retired pixels
  auto all_done = [&]() { return fminf(fminf(pxf2[0].x, pxf2[0].y), fminf(pxf2[1].x, pxf2[1].y)) == FIN; };
  bool mine_done = all_done();
  float pyf = t.pyf;
  asm volatile("" : "+f"(pyf));   // keep the row coordinate in its register (ptxas would re-derive it per record)
  // alpha (0 where the record does not reach the pixel) and 1 - alpha of one record for the four pixels
  auto alphas = [&](const float4& q0, const float4& q1, float2 (&alpha)[2], float2 (&one_m)[2], auto clamp_tag) {
    constexpr bool CLAMP = decltype(clamp_tag)::value;
    const RowTerms rt = row_terms(q0.z, q0.w, q1.x, q0.y - pyf);
    const float2 nx2 = make_float2(-q0.x, -q0.x), nhA2 = make_float2(-rt.hA, -rt.hA);
    const float2 bdy2 = make_float2(rt.bdy, rt.bdy), ncdy2 = make_float2(-rt.cdy2, -rt.cdy2);
    const float2 o2 = make_float2(q1.y, q1.y);
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const float2 ndx = __fadd2_rn(pxf2[h], nx2);
      const float2 p2 = __ffma2_rn(ndx, __ffma2_rn(nhA2, ndx, bdy2), ncdy2);
      float2 al = __fmul2_rn(o2, make_float2(ex2_fast(p2.x), ex2_fast(p2.y)));
      if (CLAMP) al = make_float2(fminf(0.99f, al.x), fminf(0.99f, al.y));
      const bool valid0 = (!CLAMP || p2.x <= 0.f) && al.x >= (1.f / 255.f);
      const bool valid1 = (!CLAMP || p2.y <= 0.f) && al.y >= (1.f / 255.f);
      alpha[h] = make_float2(valid0 ? al.x : 0.f, valid1 ? al.y : 0.f);
      one_m[h] = __ffma2_rn(alpha[h], make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
    }
  };
  // a pixel's transmittance is exhausted by the record at `pos` (at most once per pixel): the record does not act on
  // it (alpha := 0, the identity) and the pixel is retired; `nalpha` / `none_m` are those of a record evaluated ahead
  // of time against the same pixels (two-record step below), which must not act on the retired pixel either
  auto retire = [&](const float2& t0, const float2& t1, float2 (&alpha)[2], float2 (&one_m)[2], uint32_t pos,
                    float2 (*nalpha)[2], float2 (*none_m)[2]) {
#define B200GS_RETIRE(H, F, I, TV)                                                          \
    if (TV < 0.0001f) {                                                                       \
      alpha[H].F = 0.f; one_m[H].F = 1.f; nstop[I] = pos; pxf2[H].F = FIN;                    \
      if (nalpha) { (*nalpha)[H].F = 0.f; (*none_m)[H].F = 1.f; }                            \
    }
    B200GS_RETIRE(0, x, 0, t0.x)
    B200GS_RETIRE(0, y, 1, t0.y)
    B200GS_RETIRE(1, x, 2, t1.x)
    B200GS_RETIRE(1, y, 3, t1.y)
#undef B200GS_RETIRE
    mine_done = all_done();
  };
  auto apply = [&](float2 (&alpha)[2], float2 (&one_m)[2], const float4& q2, uint32_t pos, float2 (*nalpha)[2],
                   float2 (*none_m)[2]) {
    {
      const float2 t0 = __fmul2_rn(T[0], one_m[0]), t1 = __fmul2_rn(T[1], one_m[1]);
      if (fminf(fminf(t0.x, t0.y), fminf(t1.x, t1.y)) < 0.0001f) retire(t0, t1, alpha, one_m, pos, nalpha, none_m);
    }
    const float2 cr2 = make_float2(q2.x, q2.x), cg2 = make_float2(q2.y, q2.y), cb2 = make_float2(q2.z, q2.z);
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const float2 w = __fmul2_rn(alpha[h], T[h]);
      Cr[h] = __ffma2_rn(cr2, w, Cr[h]);
      Cg[h] = __ffma2_rn(cg2, w, Cg[h]);
      Cb[h] = __ffma2_rn(cb2, w, Cb[h]);
      T[h] = __fmul2_rn(T[h], one_m[h]);
    }
  };
  auto eval = [&](const float4& q0, const float4& q1, const float4& q2, uint32_t c, uint32_t jj, auto clamp_tag) {
    float2 alpha[2], one_m[2];
    alphas(q0, q1, alpha, one_m, clamp_tag);
    apply(alpha, one_m, q2, c * R4_CH + jj, nullptr, nullptr);
  };
#if R4_FWD_PAIRS
  // Two records per step: both alphas are evaluated first (eight independent exp2 chains instead of four -- the
  // kernel waits on its own dependent instructions, ncu: `wait` 2.65 cycles per issue), then applied in list order.
  auto eval2 = [&](const float4& a0, const float4& a1, const float4& a2, uint32_t ja, const float4& b0, const float4& b1,
                   const float4& b2, uint32_t jb, uint32_t c, auto clamp_tag) {
    float2 alA[2], omA[2], alB[2], omB[2];
    alphas(a0, a1, alA, omA, clamp_tag);
    alphas(b0, b1, alB, omB, clamp_tag);
    apply(alA, omA, a2, c * R4_CH + ja, &alB, &omB);
    apply(alB, omB, b2, c * R4_CH + jb, nullptr, nullptr);
  };
#endif

  uint32_t c = 0;
  bool early = false;
  for (; c < nchunks; c++) {
    ring.wait(c);
    const uint32_t cnt = ring.count(c);
    r4_prescale(sm[c % R4_STAGES], cnt, tid);
    __syncthreads();
    const float4* st = sm[c % R4_STAGES];
    for (uint32_t base = 0; base < cnt; base += 32) {
      if (__all_sync(0xffffffffu, mine_done)) break;
      const uint32_t j = base + lane;
      bool hit = false;
      if (j < cnt) {
        const float4 q0 = st[j * REC_F4], q1 = st[j * REC_F4 + 1];
        hit = r4_block_may_contribute_scaled(q0.x, q0.y, q0.z, q0.w, q1.x, q1.z, t.rx0, t.ry0, t.rx1, t.ry1);
        if (coarse && hit)
          hit = r4_tile_in_reference_rect(q0.x, q0.y, fabsf(st[j * REC_F4 + 2].w), blockIdx.x, blockIdx.y, gridDim.x,
                                          gridDim.y);
      }
      uint32_t mask = __ballot_sync(0xffffffffu, hit);
      while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t jj = base + b;
        const float4 q0 = st[jj * REC_F4], q1 = st[jj * REC_F4 + 1], q2 = st[jj * REC_F4 + 2];
#if R4_FWD_PAIRS
        if (mask) {
          const int b1 = __ffs(mask) - 1;
          mask &= mask - 1;
          const uint32_t jk = base + b1;
          const float4 p0 = st[jk * REC_F4], p1 = st[jk * REC_F4 + 1], p2 = st[jk * REC_F4 + 2];
          if (q2.w < 0.f || p2.w < 0.f) eval2(q0, q1, q2, jj, p0, p1, p2, jk, c, std::true_type{});
          else eval2(q0, q1, q2, jj, p0, p1, p2, jk, c, std::false_type{});
        } else
#endif
        if (q2.w < 0.f) eval(q0, q1, q2, c, jj, std::true_type{});
        else eval(q0, q1, q2, c, jj, std::false_type{});
        if (__all_sync(0xffffffffu, mine_done)) break;
      }
    }
    const int num_done = __syncthreads_count(mine_done);
    if (num_done == R4_THREADS) { early = true; break; }
    ring.refill(c);
  }
  // no asynchronous copy may still target this CTA's shared memory when it exits: LDGSTS groups are drained; bulk
  // copies that were issued ahead of a tile that finished early are waited for by the thread that issued them
  if constexpr (SLAB) {
    if (early && tid == 0)
      for (uint32_t k = c + 1; k < min(nchunks, c + R4_STAGES); k++) ring.wait(k);
  }
  ring.drain();

  const float Tf[4] = {T[0].x, T[0].y, T[1].x, T[1].y};
  const float R_[4] = {Cr[0].x, Cr[0].y, Cr[1].x, Cr[1].y};
  const float G_[4] = {Cg[0].x, Cg[0].y, Cg[1].x, Cg[1].y};
  const float B_[4] = {Cb[0].x, Cb[0].y, Cb[1].x, Cb[1].y};
  const size_t hw = (size_t)a.H * a.W;
  const size_t pid = (size_t)t.py * a.W + t.px0;
  if (a.vec4 && t.in[3]) {
    if (a.pix) {     // (null for forward-only frames: nothing is kept for the adjoint)
      float4* pix = a.pix + pid;
#pragma unroll
      for (int i = 0; i < 4; i++) pix[i] = make_float4(R_[i], G_[i], B_[i], Tf[i]);
      *reinterpret_cast<uint4*>(a.n_contrib + pid) = make_uint4(nstop[0], nstop[1], nstop[2], nstop[3]);
    }
    const float b0 = __ldg(a.bg + 0), b1 = __ldg(a.bg + 1), b2 = __ldg(a.bg + 2);
    if (a.out_rgb8) {
      // the thread's four pixels are 12 contiguous bytes of the [H][W][3] frame: three aligned 32-bit words
      uint32_t c[12];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        c[3 * i + 0] = rgb8_of(__fmaf_rn(Tf[i], b0, R_[i]));
        c[3 * i + 1] = rgb8_of(__fmaf_rn(Tf[i], b1, G_[i]));
        c[3 * i + 2] = rgb8_of(__fmaf_rn(Tf[i], b2, B_[i]));
      }
      uint32_t* o = reinterpret_cast<uint32_t*>(a.out_rgb8 + pid * 3);
#pragma unroll
      for (int w = 0; w < 3; w++) o[w] = c[4 * w] | (c[4 * w + 1] << 8) | (c[4 * w + 2] << 16) | (c[4 * w + 3] << 24);
      return;
    }
    *reinterpret_cast<float4*>(a.out_color + pid) =
        make_float4(__fmaf_rn(Tf[0], b0, R_[0]), __fmaf_rn(Tf[1], b0, R_[1]), __fmaf_rn(Tf[2], b0, R_[2]),
                    __fmaf_rn(Tf[3], b0, R_[3]));
    *reinterpret_cast<float4*>(a.out_color + hw + pid) =
        make_float4(__fmaf_rn(Tf[0], b1, G_[0]), __fmaf_rn(Tf[1], b1, G_[1]), __fmaf_rn(Tf[2], b1, G_[2]),
                    __fmaf_rn(Tf[3], b1, G_[3]));
    *reinterpret_cast<float4*>(a.out_color + 2 * hw + pid) =
        make_float4(__fmaf_rn(Tf[0], b2, B_[0]), __fmaf_rn(Tf[1], b2, B_[1]), __fmaf_rn(Tf[2], b2, B_[2]),
                    __fmaf_rn(Tf[3], b2, B_[3]));
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (!t.in[i]) continue;
      if (a.pix) {
        a.pix[pid + i] = make_float4(R_[i], G_[i], B_[i], Tf[i]);
        a.n_contrib[pid + i] = nstop[i];
      }
      const float o0 = __fmaf_rn(Tf[i], __ldg(a.bg + 0), R_[i]), o1 = __fmaf_rn(Tf[i], __ldg(a.bg + 1), G_[i]);
      const float o2 = __fmaf_rn(Tf[i], __ldg(a.bg + 2), B_[i]);
      if (a.out_rgb8) {
        uint8_t* o = a.out_rgb8 + (pid + i) * 3;
        o[0] = (uint8_t)rgb8_of(o0); o[1] = (uint8_t)rgb8_of(o1); o[2] = (uint8_t)rgb8_of(o2);
        continue;
      }
      a.out_color[pid + i] = o0;
      a.out_color[hw + pid + i] = o1;
      a.out_color[2 * hw + pid + i] = o2;
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
template <bool SLAB>
__global__ void __launch_bounds__(R4_THREADS, R4_BWD_MINB) k_render_bwd4(RenderBwdArgs a) {
  __shared__ __align__(128) float4 sm[R4_STAGES][R4_CH * REC_F4];
  __shared__ __align__(8) uint64_t s_bar[R4_STAGES];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint2 range = a.ranges[(blockIdx.y >> a.bin_shift) * a.gbx + (blockIdx.x >> a.bin_shift)];
  const bool coarse = a.bin_shift != 0;
  const uint32_t n = range.y - range.x;
  const uint32_t nchunks = (n + R4_CH - 1) / R4_CH;
  using Ring = std::conditional_t<SLAB, Ring4Slab, Ring4>;
  Ring ring = [&]() {
    if constexpr (SLAB) return Ring4Slab{sm, s_bar, a.slab + (size_t)range.x * REC_F4, n, nchunks, tid};
    else return Ring4{sm, a.point_list + range.x, a.rec, n, nchunks, 0u, tid};
  }();
  const Tile4 t = tile4_setup(a.W, a.H, warp, lane);

  ring.prologue();

  float2 gr[2], gg[2], gb[2], F[2], T[2], R[2], pxf2[2];
  uint32_t nc[4];
  {
    float gr_[4], gg_[4], gb_[4], F_[4];
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
      gr_[0] = r4.x; gr_[1] = r4.y; gr_[2] = r4.z; gr_[3] = r4.w;
      gg_[0] = g4.x; gg_[1] = g4.y; gg_[2] = g4.z; gg_[3] = g4.w;
      gb_[0] = b4.x; gb_[1] = b4.y; gb_[2] = b4.z; gb_[3] = b4.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        fin[i] = make_float4(0.f, 0.f, 0.f, 1.f);
        nc[i] = 0u; gr_[i] = 0.f; gg_[i] = 0.f; gb_[i] = 0.f;
        if (t.in[i]) {
          fin[i] = a.pix[pid + i];
          nc[i] = a.n_contrib[pid + i];
          gr_[i] = __ldg(a.dL_dpix + pid + i);
          gg_[i] = __ldg(a.dL_dpix + hw + pid + i);
          gb_[i] = __ldg(a.dL_dpix + 2 * hw + pid + i);
        }
      }
    }
    const float b0 = __ldg(a.bg), b1 = __ldg(a.bg + 1), b2 = __ldg(a.bg + 2);
#pragma unroll
    for (int i = 0; i < 4; i++)
      F_[i] = fin[i].x * gr_[i] + fin[i].y * gg_[i] + fin[i].z * gb_[i] + fin[i].w * (b0 * gr_[i] + b1 * gg_[i] + b2 * gb_[i]);
#pragma unroll
    for (int h = 0; h < 2; h++) {
      gr[h] = make_float2(gr_[2 * h], gr_[2 * h + 1]);
      gg[h] = make_float2(gg_[2 * h], gg_[2 * h + 1]);
      gb[h] = make_float2(gb_[2 * h], gb_[2 * h + 1]);
      F[h] = make_float2(F_[2 * h], F_[2 * h + 1]);
      T[h] = make_float2(1.f, 1.f);
      R[h] = make_float2(0.f, 0.f);
      pxf2[h] = make_float2(t.pxf[2 * h], t.pxf[2 * h + 1]);
    }
  }
  // nothing at or behind list position max(n_contrib) over the WARP's 128 pixels reaches any of them: those records
  // are not even tested (at batch granularity half a batch of alpha evaluations per tile was spent behind it)
  const uint32_t ncmax = __reduce_max_sync(0xffffffffu, max(max(nc[0], nc[1]), max(nc[2], nc[3])));

  uint32_t c = 0;
  bool early = false;
  for (; c < nchunks; c++) {
    ring.wait(c);
    const uint32_t cnt = ring.count(c);
    r4_prescale(sm[c % R4_STAGES], cnt, tid);
    __syncthreads();
    const float4* st = sm[c % R4_STAGES];
    for (uint32_t base = 0; base < cnt; base += 32) {
      if (c * R4_CH + base >= ncmax) break;
      const uint32_t j = base + lane;
      bool hit = false;
      if (j < cnt && c * R4_CH + j < ncmax) {
        const float4 q0 = st[j * REC_F4], q1 = st[j * REC_F4 + 1];
        hit = r4_block_may_contribute_scaled(q0.x, q0.y, q0.z, q0.w, q1.x, q1.z, t.rx0, t.ry0, t.rx1, t.ry1);
        if (coarse && hit)
          hit = r4_tile_in_reference_rect(q0.x, q0.y, fabsf(st[j * REC_F4 + 2].w), blockIdx.x, blockIdx.y, gridDim.x,
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
        // packed f32x2 over the pixel pairs (0|1, 2|3), branch-free: a record that does not contribute to a
        // pixel acts on it with alpha = G = 0.  ndx = px - x = -dx, so Sx = -sum(m ndx), Sxx = sum(m ndx^2).
        const float2 nx2 = make_float2(-q0.x, -q0.x), nhA2 = make_float2(-rt.hA, -rt.hA);
        const float2 bdy2 = make_float2(rt.bdy, rt.bdy), ncdy2 = make_float2(-rt.cdy2, -rt.cdy2);
        const float2 o2 = make_float2(q1.y, q1.y);
        float2 ndx[2], G[2], alpha[2], vmask[2];
        bool any_valid = false;
        // marked records (negative radius, project.cu:record_is_general) clamp alpha at 0.99 and skip pixels whose
        // power rounds above 0; for all others neither can happen and both tests are dropped (as in the forward)
        auto alphas = [&](auto general_tag) {
          constexpr bool GENERAL = decltype(general_tag)::value;
#pragma unroll
          for (int h = 0; h < 2; h++) {
            ndx[h] = __fadd2_rn(pxf2[h], nx2);
            const float2 p2 = __ffma2_rn(ndx[h], __ffma2_rn(nhA2, ndx[h], bdy2), ncdy2);
            G[h] = make_float2(ex2_fast(p2.x), ex2_fast(p2.y));
            const float2 og = __fmul2_rn(o2, G[h]);
            alpha[h] = GENERAL ? make_float2(fminf(0.99f, og.x), fminf(0.99f, og.y)) : og;
            const bool v0 = pos < nc[2 * h] && (!GENERAL || p2.x <= 0.f) && alpha[h].x >= (1.f / 255.f);
            const bool v1 = pos < nc[2 * h + 1] && (!GENERAL || p2.y <= 0.f) && alpha[h].y >= (1.f / 255.f);
            vmask[h] = make_float2(v0 ? 1.f : 0.f, v1 ? 1.f : 0.f);
            any_valid = any_valid || v0 || v1;
          }
        };
        if (q2.w < 0.f) alphas(std::true_type{});
        else alphas(std::false_type{});
        if (!__any_sync(0xffffffffu, any_valid)) continue;
        const float2 one2 = make_float2(1.f, 1.f), neg2 = make_float2(-1.f, -1.f);
        const float2 cr2 = make_float2(q2.x, q2.x), cg2 = make_float2(q2.y, q2.y), cb2 = make_float2(q2.z, q2.z);
        float2 vr2 = make_float2(0.f, 0.f), vg2 = vr2, vb2 = vr2, s02 = vr2, nsx2 = vr2, sxx2 = vr2;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const float2 al = __fmul2_rn(alpha[h], vmask[h]);
          const float2 Gi = __fmul2_rn(G[h], vmask[h]);
          const float2 w = __fmul2_rn(al, T[h]);
          const float2 cg = __ffma2_rn(cb2, gb[h], __ffma2_rn(cg2, gg[h], __fmul2_rn(cr2, gr[h])));
          R[h] = __ffma2_rn(w, cg, R[h]);
          const float2 one_m = __ffma2_rn(al, neg2, one2);
          const float2 rc = make_float2(rcp_fast(one_m.x), rcp_fast(one_m.y));
          // dL/dalpha = T cg - (F - R) / (1 - alpha)
          const float2 dL_dalpha = __ffma2_rn(T[h], cg, __fmul2_rn(rc, __ffma2_rn(F[h], neg2, R[h])));
          T[h] = __fmul2_rn(T[h], one_m);
          const float2 m = __fmul2_rn(Gi, dL_dalpha);
          vr2 = __ffma2_rn(w, gr[h], vr2);
          vg2 = __ffma2_rn(w, gg[h], vg2);
          vb2 = __ffma2_rn(w, gb[h], vb2);
          s02 = __fadd2_rn(s02, m);
          const float2 nmx = __fmul2_rn(m, ndx[h]);
          nsx2 = __fadd2_rn(nsx2, nmx);
          sxx2 = __ffma2_rn(nmx, ndx[h], sxx2);
        }
        const float vr = vr2.x + vr2.y, vg = vg2.x + vg2.y, vb = vb2.x + vb2.y;
        const float s0 = s02.x + s02.y, sx = -(nsx2.x + nsx2.y), sxx = sxx2.x + sxx2.y;
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
    if (num_done == R4_THREADS) { early = true; break; }
    ring.refill(c);
  }
  if constexpr (SLAB) {
    if (early && tid == 0)
      for (uint32_t k = c + 1; k < min(nchunks, c + R4_STAGES); k++) ring.wait(k);
  }
  ring.drain();
}

void launch_render4(const RenderArgs& a, cudaStream_t st) {
  const dim3 grid((a.W + TILE - 1) / TILE, (a.H + TILE - 1) / TILE), block(R4_THREADS);
  if (a.slab) k_render_fwd4<true><<<grid, block, 0, st>>>(a);
  else k_render_fwd4<false><<<grid, block, 0, st>>>(a);
  count_launch();
}

void launch_render_bwd4(const RenderBwdArgs& a, cudaStream_t st) {
  const dim3 grid((a.W + TILE - 1) / TILE, (a.H + TILE - 1) / TILE), block(R4_THREADS);
  if (a.slab) k_render_bwd4<true><<<grid, block, 0, st>>>(a);
  else k_render_bwd4<false><<<grid, block, 0, st>>>(a);
  count_launch();
}

}  // namespace b200gs

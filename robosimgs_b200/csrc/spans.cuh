// spans.cuh -- tile/bin coverage of a projected splat, shared by the projection kernel (count) and the
// pair-emission kernels (project.cu, bucket.cu).  Everything is explicitly rounded, so every
// translation unit that inlines it sees bit-identical spans.
#pragma once
#include "common.cuh"

namespace b200gs {

// ---- tile coverage --------------------------------------------------------------------------------
// The public algorithm assigns a Gaussian to every tile of the square [c - r, c + r] with
// r = ceil(3 sqrt(lambda_max)).  A pixel only receives a contribution when alpha = o*exp(power)
// >= 1/255, i.e. when 0.5*q(d) <= ln(255 o) with q the conic quadratic form.  Tiles of the square
// that the ellipse {0.5 q <= thr} cannot reach contribute nothing to the image or to any gradient,
// so they are dropped: coverage = reference rect  intersected with  per-tile-row ellipse spans.
// thr carries a +0.01 slack (alpha ratio 1%) so fp32 rounding can only add pairs, never lose one.
constexpr uint32_t EMIT_BIG_THRESHOLD = 96;  // tiles; above this a whole warp emits the Gaussian

struct TileRect {
  int x0, y0, x1, y1;  // [x0,x1) x [y0,y1) in tiles
};

__device__ __forceinline__ TileRect reference_rect(float px, float py, int radius, int gx, int gy) {
  TileRect r;
  const float fr = (float)radius;
  r.x0 = min(gx, max(0, (int)((px - fr) / TILE)));
  r.y0 = min(gy, max(0, (int)((py - fr) / TILE)));
  r.x1 = min(gx, max(0, (int)((px + fr + (TILE - 1)) / TILE)));
  r.y1 = min(gy, max(0, (int)((py + fr + (TILE - 1)) / TILE)));
  return r;
}

// Binning granularity.  Pairs are sorted per *bin* of (16 << shift)^2 pixels; every 16x16 compositing
// CTA walks the list of the bin it lies in and re-applies the reference rect (exactly) and the
// ellipse test per record.  Coarser bins mean fewer (Gaussian, bin) pairs to emit and sort.
__device__ __forceinline__ TileRect bin_rect(const TileRect& r, int shift) {
  TileRect b;
  b.x0 = r.x0 >> shift;
  b.y0 = r.y0 >> shift;
  b.x1 = (r.x1 > r.x0) ? ((r.x1 - 1) >> shift) + 1 : b.x0;
  b.y1 = (r.y1 > r.y0) ? ((r.y1 - 1) >> shift) + 1 : b.y0;
  return b;
}

// All span arithmetic uses explicitly rounded intrinsics / fixed PTX approximations, so the count
// (k_project) and the emission (k_emit_*) see bit-identical spans wherever the code is inlined.
__device__ __forceinline__ float sqrt_approx(float v) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
// one MUFU.RCP (1 ulp) instead of the ten-instruction correctly rounded reciprocal: the spans are padded outward by
// 0.5 % + 0.02 px, and count and emission inline the same instruction
__device__ __forceinline__ float rcp_approx(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

struct SpanCtx {
  float x, y, B;
  float A_tau;      // A * tau, tau = 2*thr
  float inv_A;
  float det;        // A*C - B*B
  float x_ext;      // half-extent of the ellipse in x
  float y_at_xext;  // dy at the right-most point of the ellipse
  float y_ext;
  float pad;
  float bt, inv_bt; // bin edge in pixels and its reciprocal
  int ty0, ty1;     // bin rows the ellipse can reach, clipped to the reference rect
};

__device__ __forceinline__ bool span_setup(SpanCtx& s, float x, float y, float A, float B, float C,
                                           float thr, const TileRect& r, int bin_shift) {
  s.x = x; s.y = y; s.B = B;
  s.bt = (float)(TILE << bin_shift);
  s.inv_bt = 1.f / s.bt;   // power of two: exact
  const float tau = __fmul_rn(2.f, thr);
  s.det = __fmaf_rn(A, C, -__fmul_rn(B, B));
  s.ty0 = s.ty1 = 0;
  if (!(thr > 0.f) || !(s.det > 0.f) || !(A > 0.f) || !(C > 0.f)) return false;
  s.A_tau = __fmul_rn(A, tau);
  s.inv_A = rcp_approx(A);
  const float inv_det = rcp_approx(s.det);
  // 0.5% + 0.02 px outward padding absorbs the approximations below
  s.x_ext = __fmaf_rn(sqrt_approx(__fmul_rn(__fmul_rn(tau, C), inv_det)), 1.005f, 0.02f);
  s.y_ext = __fmaf_rn(sqrt_approx(__fmul_rn(__fmul_rn(tau, A), inv_det)), 1.005f, 0.02f);
  s.y_at_xext = -__fmul_rn(__fmul_rn(B, s.x_ext), rcp_approx(C));
  s.pad = __fmaf_rn(0.01f, s.x_ext, 0.02f);
  // bin row ty holds pixel-centre rows [bt ty, bt ty + bt - 1]
  const int lo = (int)ceilf(__fmul_rn(__fadd_rn(__fadd_rn(y, -s.y_ext), 1.f - s.bt), s.inv_bt));
  const int hi = (int)floorf(__fmul_rn(__fadd_rn(y, s.y_ext), s.inv_bt)) + 1;
  s.ty0 = max(r.y0, lo);
  s.ty1 = min(r.y1, hi);
  return s.ty1 > s.ty0;
}

// bin-column span [c0,c1) of bin row ty: x-interval of the ellipse {A dx^2 + 2B dx dy + C dy^2 <=
// tau} inside the band of pixel-centre rows [bt ty, bt ty + bt - 1], clipped to the reference rect.
__device__ __forceinline__ void row_span(const SpanCtx& s, const TileRect& r, int ty, int& c0, int& c1) {
  c0 = c1 = 0;
  const float row0 = __fmul_rn((float)ty, s.bt);
  float a = __fadd_rn(row0 - 0.02f, -s.y);
  float b = __fadd_rn(row0 + (s.bt - 1.f) + 0.02f, -s.y);
  a = fmaxf(a, -s.y_ext);
  b = fminf(b, s.y_ext);
  if (a > b) return;
  // half-width at dy: sqrt(A*tau - det*dy^2)/A ; centre line: -B*dy/A
  const float da = sqrt_approx(fmaxf(0.f, __fmaf_rn(-s.det, __fmul_rn(a, a), s.A_tau)));
  const float db = sqrt_approx(fmaxf(0.f, __fmaf_rn(-s.det, __fmul_rn(b, b), s.A_tau)));
  const float ca = -__fmul_rn(s.B, a), cb = -__fmul_rn(s.B, b);
  float xmax = __fmul_rn(fmaxf(__fadd_rn(ca, da), __fadd_rn(cb, db)), s.inv_A);
  float xmin = __fmul_rn(fminf(__fadd_rn(ca, -da), __fadd_rn(cb, -db)), s.inv_A);
  if (s.y_at_xext >= a && s.y_at_xext <= b) xmax = s.x_ext;     // right-most point inside the band
  if (-s.y_at_xext >= a && -s.y_at_xext <= b) xmin = -s.x_ext;  // left-most point inside the band
  const float X0 = __fadd_rn(s.x, __fadd_rn(xmin, -s.pad)), X1 = __fadd_rn(s.x, __fadd_rn(xmax, s.pad));
  // tile tx holds pixel centres [16 tx, 16 tx + 15]: intersects [X0,X1] iff 16tx <= X1 and 16tx+15 >= X0
  int t0 = (int)ceilf(__fmul_rn(__fadd_rn(X0, 1.f - s.bt), s.inv_bt));
  int t1 = (int)floorf(__fmul_rn(X1, s.inv_bt)) + 1;
  t0 = max(t0, r.x0);
  t1 = min(t1, r.x1);
  if (t1 > t0) { c0 = t0; c1 = t1; }
}

// Number of bins the splat touches.  With bucket_count != nullptr (bucketed binning, bucket.cu) the pair counter
// of every touched bin's bucket is incremented as well (one RED per pair): bucket_count already points at the
// splat's depth slice, the buckets of one bin are `bucket_stride` counters apart.  *packed (bucketed binning with at
// most 255 bins only) receives the footprint in the form the emission kernel reads back from tiles[]:
//   TILES_PACKED | b2 << 16 | b1 << 8 | b0 with count = number of bins (1..3) in bits 24..25  -- small footprints,
//   the plain count otherwise (emission recomputes the spans).
constexpr uint32_t TILES_PACKED = 0x80000000u;
__device__ __forceinline__ uint32_t tiles_count(uint32_t t) { return (t & TILES_PACKED) ? (t >> 24) & 3u : t; }

__device__ __forceinline__ uint32_t count_tiles(float x, float y, float A, float B, float C, float thr,
                                                const TileRect& r, int bin_shift, uint32_t* bucket_count,
                                                int bucket_stride_log2, int gbx, bool pack, uint32_t* packed) {
  SpanCtx s;
  *packed = 0u;
  if (!span_setup(s, x, y, A, B, C, thr, r, bin_shift)) return 0;
  uint32_t n = 0, ids = 0;
  for (int ty = s.ty0; ty < s.ty1; ty++) {
    int c0, c1;
    row_span(s, r, ty, c0, c1);
    if (bucket_count)
      for (int tx = c0; tx < c1; tx++) {
        const uint32_t bin = (uint32_t)(ty * gbx + tx);
        atomicAdd(bucket_count + ((size_t)bin << bucket_stride_log2), 1u);
        const uint32_t k = n + (uint32_t)(tx - c0);
        if (k < 3u) ids |= bin << (8u * k);
      }
    n += (uint32_t)(c1 - c0);
  }
  *packed = (pack && n >= 1u && n <= 3u) ? (TILES_PACKED | (n << 24) | ids) : n;
  return n;
}

}  // namespace b200gs

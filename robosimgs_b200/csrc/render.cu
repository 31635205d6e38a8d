// render.cu -- per-tile front-to-back alpha compositing and its adjoint.
//
// Replaces rows a8 renderCUDA fwd and a9 renderCUDA bwd of the public
// diff-gaussian-rasterization named by BASELINE.json:north_star (third-party; the reference repo
// only delegates, /root/reference/README.md:75).  Per-pixel arithmetic follows
// oracle/gs_oracle_impl.h gso_render / gso_render_backward.
//
// B200 design:
//  * one CTA per 16x16 tile, 8 warps, each warp owns an 8x4 pixel block;
//  * a tile's splats are a contiguous range of the sorted Gaussian-id list; the 48-byte projected
//    records (L2-resident table, P_vis * 48 B) are gathered straight into a 2-stage shared-memory
//    ring, 256 records per stage, by asynchronous copies that bypass the register file: one 48-byte
//    TMA bulk copy per record (cp.async.bulk, SASS UBLKCP) completing on the stage's mbarrier
//    (GATHER_TMA) or three 16-byte cp.async (LDGSTS) per record (GATHER_LDGSTS).  No tile-ordered
//    copy of the records ("slab") is ever written to HBM;
//  * hierarchical culling: 32 lanes test 32 different records against the warp's 8x4 pixel block
//    (exact minimum of the conic form over the block), one ballot, and only the surviving records
//    are evaluated per pixel -- records are read from shared memory as 128-bit broadcasts;
//  * early-out: per-warp vote when all 32 pixels are saturated, per-CTA __syncthreads_count;
//  * adjoint: gradients of a record are reduced over the warp with a transposing butterfly
//    (14 shuffles for 9 values) and leave the SM as one 9-lane RED.ADD.F32 to a 48-byte aligned
//    accumulator row per Gaussian.
#include <atomic>
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"

namespace b200gs {

constexpr int CH = 256;      // records per ring stage
constexpr int STAGES = 2;
constexpr int CH_BYTES = CH * REC_F4 * 16;

// Upper bound of `power` over the pixel block [x0,x1]x[y0,y1] is -0.5*qmin with qmin the minimum of
// q(d) = A dx^2 + 2 B dx dy + C dy^2.  q is convex, so if the centre is outside the block the
// minimum lies on an edge facing the centre; both facing edges are minimised in closed form.
__device__ __forceinline__ bool block_may_contribute(float x, float y, float A, float B, float C, float thr,
                                                     float x0, float y0, float x1, float y1) {
  const float cx = clampf(x, x0, x1), cy = clampf(y, y0, y1);
  const float dxe = cx - x, dye = cy - y;
  // vertical edge X = cx: minimise over Y
  float dy1 = clampf(y - B * dxe * rcp_fast(C), y0, y1) - y;
  const float q1 = A * dxe * dxe + 2.f * B * dxe * dy1 + C * dy1 * dy1;
  // horizontal edge Y = cy: minimise over X
  float dx2 = clampf(x - B * dye * rcp_fast(A), x0, x1) - x;
  const float q2 = A * dx2 * dx2 + 2.f * B * dx2 * dye + C * dye * dye;
  const float q = fminf(q1, q2);
  // rounding guard: relative to the magnitude of the cancelling terms
  const float mag = fabsf(A) * (dxe * dxe + dx2 * dx2) + fabsf(C) * (dye * dye + dy1 * dy1);
  return 0.5f * q - 4e-6f * mag <= thr;
}

// Exactly the reference's tile rect (project.cu:reference_rect): is 16x16 tile (tx,ty) inside the
// square [c - r, c + r] of a splat?  Needed when pairs are binned coarser than 16 px.
__device__ __forceinline__ bool tile_in_reference_rect(float px, float py, float fr, int tx, int ty, int gx, int gy) {
  const int x0 = min(gx, max(0, (int)((px - fr) / TILE)));
  const int y0 = min(gy, max(0, (int)((py - fr) / TILE)));
  const int x1 = min(gx, max(0, (int)((px + fr + (TILE - 1)) / TILE)));
  const int y1 = min(gy, max(0, (int)((py + fr + (TILE - 1)) / TILE)));
  return tx >= x0 && tx < x1 && ty >= y0 && ty < y1;
}

enum { GATHER_TMA = 0, GATHER_LDGSTS = 1 };

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Shared-memory ring of gathered splat records, filled cooperatively by the CTA's 256 threads.
// Thread t owns slot t of every stage; the Gaussian id for the *next* fill is prefetched into a
// register one chunk ahead so the dependent gather never waits on the id load.
template <int MODE>
struct GatherRing {
  float4 (*sm)[CH * REC_F4];
  uint64_t* full;
  const uint32_t* list;   // sorted Gaussian ids of this tile
  const float4* rec;
  uint32_t n, nchunks;
  uint32_t next_id;       // id for slot tid of chunk `next_chunk`
  int tid;

  __device__ __forceinline__ uint32_t count(uint32_t c) const { return min((uint32_t)CH, n - c * CH); }
  __device__ __forceinline__ uint32_t load_id(uint32_t c) const {
    return (c < nchunks && (uint32_t)tid < count(c)) ? __ldg(list + c * CH + tid) : 0u;
  }
  // issue the gather of chunk c into stage c&1 using id (all threads call)
  __device__ __forceinline__ void issue(uint32_t c, uint32_t id) {
    const int s = c & 1;
    if (c < nchunks) {
      const uint32_t cnt = count(c);
      if (MODE == GATHER_TMA) {
        if ((uint32_t)tid < cnt) tma_load_1d(&sm[s][tid * REC_F4], rec + (size_t)id * REC_F4, REC_F4 * 16, &full[s]);
        if (tid == 0) mbar_expect_tx(&full[s], cnt * REC_F4 * 16);
      } else {
        if ((uint32_t)tid < cnt) {
          const float4* src = rec + (size_t)id * REC_F4;
          cp_async16(&sm[s][tid * REC_F4], src);
          cp_async16(&sm[s][tid * REC_F4 + 1], src + 1);
          cp_async16(&sm[s][tid * REC_F4 + 2], src + 2);
        }
      }
    }
    if (MODE == GATHER_LDGSTS) cp_async_commit();   // one (possibly empty) group per call
  }
  __device__ __forceinline__ void prologue() {
    if (MODE == GATHER_TMA) {
      if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_fence_init();
      }
      __syncthreads();
    }
    const uint32_t id0 = load_id(0), id1 = load_id(1);
    issue(0, id0);
    issue(1, id1);
    next_id = load_id(2);
  }
  // block until chunk c is resident in stage c&1 (all threads call)
  __device__ __forceinline__ void wait(uint32_t c) {
    if (MODE == GATHER_TMA) {
      mbar_wait(&full[c & 1], (c >> 1) & 1);
    } else {
      cp_async_wait<1>();
      __syncthreads();
    }
  }
  // stage c&1 is free (caller synchronised the CTA): refill it with chunk c+2
  __device__ __forceinline__ void refill(uint32_t c) {
    issue(c + 2, next_id);
    next_id = load_id(c + 3);
  }
  // leave no asynchronous copy in flight into this CTA's shared memory
  __device__ __forceinline__ void drain(uint32_t c) {
    if (MODE == GATHER_TMA) {
      if (tid == 0 && c + 1 < nchunks) mbar_wait(&full[(c + 1) & 1], ((c + 1) >> 1) & 1);
    } else {
      cp_async_wait<0>();
    }
  }
};

// ==================================================================================================
// K6: forward compositing
// ==================================================================================================
template <int MODE>
__global__ void __launch_bounds__(256) k_render_fwd(RenderArgs a) {
  __shared__ __align__(128) float4 sm[STAGES][CH * REC_F4];
  __shared__ __align__(8) uint64_t full[STAGES];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint2 range = a.ranges[(blockIdx.y >> a.bin_shift) * a.gbx + (blockIdx.x >> a.bin_shift)];
  const bool coarse = a.bin_shift != 0;
  const uint32_t n = range.y - range.x;
  const uint32_t nchunks = (n + CH - 1) / CH;
  GatherRing<MODE> ring{sm, full, a.point_list + range.x, a.rec, n, nchunks, 0u, tid};

  const int bx = blockIdx.x * TILE + (warp & 1) * 8, by = blockIdx.y * TILE + (warp >> 1) * 4;
  const int px = bx + (lane & 7), py = by + (lane >> 3);
  const bool inside = px < a.W && py < a.H;
  const float pxf = (float)px, pyf = (float)py;
  const float rx0 = (float)bx, ry0 = (float)by;
  const float rx1 = (float)min(bx + 7, a.W - 1), ry1 = (float)min(by + 3, a.H - 1);

  ring.prologue();

  bool done = !inside;
  float T = 1.f, Cr = 0.f, Cg = 0.f, Cb = 0.f;
  uint32_t last = 0;

  for (uint32_t c = 0; c < nchunks; c++) {
    const int s = c & 1;
    ring.wait(c);
    const uint32_t cnt = min((uint32_t)CH, n - c * CH);
    const float4* st = sm[s];
    for (uint32_t base = 0; base < cnt; base += 32) {
      if (__all_sync(0xffffffffu, done)) break;
      const uint32_t j = base + lane;
      bool hit = false;
      if (j < cnt) {
        const float4 q0 = st[j * REC_F4], q1 = st[j * REC_F4 + 1];
        hit = block_may_contribute(q0.x, q0.y, q0.z, q0.w, q1.x, q1.z, rx0, ry0, rx1, ry1);
        if (coarse && hit)
          hit = tile_in_reference_rect(q0.x, q0.y, fabsf(st[j * REC_F4 + 2].w), blockIdx.x, blockIdx.y, gridDim.x, gridDim.y);
      }
      uint32_t mask = __ballot_sync(0xffffffffu, hit);
      while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t jj = base + b;
        const float4 q0 = st[jj * REC_F4], q1 = st[jj * REC_F4 + 1], q2 = st[jj * REC_F4 + 2];
        const float dx = q0.x - pxf, dy = q0.y - pyf;
        const float power = -0.5f * (q0.z * dx * dx + q1.x * dy * dy) - q0.w * dx * dy;
        const float alpha = fminf(0.99f, q1.y * exp_fast(power));
        if (!done && power <= 0.f && alpha >= (1.f / 255.f)) {
          const float test_T = T * (1.f - alpha);
          if (test_T < 0.0001f) {
            done = true;
          } else {
            const float w = alpha * T;
            Cr = __fmaf_rn(q2.x, w, Cr);
            Cg = __fmaf_rn(q2.y, w, Cg);
            Cb = __fmaf_rn(q2.z, w, Cb);
            T = test_T;
            last = c * CH + jj + 1;
          }
        }
        // leave the list behind the record that finished the warp's last pixel (done only changes on the rare path)
        if (__all_sync(0xffffffffu, done)) break;
      }
    }
    const int num_done = __syncthreads_count(done);
    if (num_done == 256) {
      ring.drain(c);   // a prefetched chunk may still be in flight into our shared memory
      break;
    }
    ring.refill(c);
  }
  if (MODE == GATHER_LDGSTS) cp_async_wait<0>();

  if (inside) {
    const size_t pid = (size_t)py * a.W + px;
    const size_t hw = (size_t)a.H * a.W;
    if (a.pix) {     // (null for forward-only frames)
      a.pix[pid] = make_float4(Cr, Cg, Cb, T);
      a.n_contrib[pid] = last;
    }
    const float o0 = __fmaf_rn(T, __ldg(a.bg + 0), Cr), o1 = __fmaf_rn(T, __ldg(a.bg + 1), Cg);
    const float o2 = __fmaf_rn(T, __ldg(a.bg + 2), Cb);
    if (a.out_rgb8) {
      uint8_t* o = a.out_rgb8 + pid * 3;
      o[0] = (uint8_t)rgb8_of(o0); o[1] = (uint8_t)rgb8_of(o1); o[2] = (uint8_t)rgb8_of(o2);
    } else {
      a.out_color[pid] = o0;
      a.out_color[hw + pid] = o1;
      a.out_color[2 * hw + pid] = o2;
    }
  }
}

// Reduce 9 per-lane values over the warp.  v[0..7] go through a transposing butterfly (lane L ends
// up with the warp total of v[(L >> 2) & 7]); v8 through a plain butterfly.
__device__ __forceinline__ void warp_reduce9(float (&v)[8], float& v8, int lane) {
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
// K7: compositing adjoint.  Front-to-back replay (same traversal, same ring as the forward): with
// running transmittance T_j and the running prefix of composited colour, the colour composited behind
// splat j is S_j = C_final - prefix_j, so
//   dL/dalpha_j = T_j * <c_j, g> - (<S_j, g> + T_final * <bg, g>) / (1 - alpha_j),   g = dL/dpixel
// which is algebraically the back-to-front recursion of the public algorithm.
// Only the projection of the prefix onto g is needed, so it is carried as one scalar.
// ==================================================================================================
template <int MODE>
__global__ void __launch_bounds__(256) k_render_bwd(RenderBwdArgs a) {
  __shared__ __align__(128) float4 sm[STAGES][CH * REC_F4];
  __shared__ __align__(8) uint64_t full[STAGES];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint2 range = a.ranges[(blockIdx.y >> a.bin_shift) * a.gbx + (blockIdx.x >> a.bin_shift)];
  const bool coarse = a.bin_shift != 0;
  const uint32_t n = range.y - range.x;
  const uint32_t nchunks = (n + CH - 1) / CH;
  GatherRing<MODE> ring{sm, full, a.point_list + range.x, a.rec, n, nchunks, 0u, tid};

  const int bx = blockIdx.x * TILE + (warp & 1) * 8, by = blockIdx.y * TILE + (warp >> 1) * 4;
  const int px = bx + (lane & 7), py = by + (lane >> 3);
  const bool inside = px < a.W && py < a.H;
  const float pxf = (float)px, pyf = (float)py;
  const float rx0 = (float)bx, ry0 = (float)by;
  const float rx1 = (float)min(bx + 7, a.W - 1), ry1 = (float)min(by + 3, a.H - 1);

  ring.prologue();

  float4 fin = make_float4(0.f, 0.f, 0.f, 1.f);
  uint32_t ncontrib = 0;
  float gr = 0.f, gg = 0.f, gb = 0.f;
  if (inside) {
    const size_t pid = (size_t)py * a.W + px;
    const size_t hw = (size_t)a.H * a.W;
    fin = a.pix[pid];
    ncontrib = a.n_contrib[pid];
    gr = __ldg(a.dL_dpix + pid);
    gg = __ldg(a.dL_dpix + hw + pid);
    gb = __ldg(a.dL_dpix + 2 * hw + pid);
  }
  // F = <C_final, g> + T_final * <bg, g>.  With R_j = sum_{k<=j} w_k <c_k, g> (a running scalar) the
  // "colour behind splat j" term <S_j, g> + T_final <bg, g> is simply F - R_j.
  const float F = fin.x * gr + fin.y * gg + fin.z * gb +
                  fin.w * (__ldg(a.bg) * gr + __ldg(a.bg + 1) * gg + __ldg(a.bg + 2) * gb);
  float T = 1.f, R = 0.f;
  // nothing at or behind the largest n_contrib of the WARP's pixels reaches any of them: not even tested
  const uint32_t wmax = __reduce_max_sync(0xffffffffu, ncontrib);

  for (uint32_t c = 0; c < nchunks; c++) {
    const int s = c & 1;
    ring.wait(c);
    const uint32_t cnt = min((uint32_t)CH, n - c * CH);
    const float4* st = sm[s];
    for (uint32_t base = 0; base < cnt; base += 32) {
      // this warp is finished once the list position passes its last contributor
      if (c * CH + base >= wmax) break;
      const uint32_t j = base + lane;
      bool hit = false;
      if (j < cnt && c * CH + j < wmax) {
        const float4 q0 = st[j * REC_F4], q1 = st[j * REC_F4 + 1];
        hit = block_may_contribute(q0.x, q0.y, q0.z, q0.w, q1.x, q1.z, rx0, ry0, rx1, ry1);
        if (coarse && hit)
          hit = tile_in_reference_rect(q0.x, q0.y, fabsf(st[j * REC_F4 + 2].w), blockIdx.x, blockIdx.y, gridDim.x, gridDim.y);
      }
      uint32_t mask = __ballot_sync(0xffffffffu, hit);
      while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t jj = base + b;
        const float4 q0 = st[jj * REC_F4], q1 = st[jj * REC_F4 + 1], q2 = st[jj * REC_F4 + 2];
        const float dx = q0.x - pxf, dy = q0.y - pyf;
        const float power = -0.5f * (q0.z * dx * dx + q1.x * dy * dy) - q0.w * dx * dy;
        const float G = exp_fast(power);
        const float alpha = fminf(0.99f, q1.y * G);
        const bool valid = (c * CH + jj < ncontrib) && power <= 0.f && alpha >= (1.f / 255.f);
        if (!__any_sync(0xffffffffu, valid)) continue;
        float v[8], v8;
        if (valid) {
          const float w = alpha * T;
          const float cg = q2.x * gr + q2.y * gg + q2.z * gb;
          R = __fmaf_rn(w, cg, R);
          const float inv1ma = rcp_fast(1.f - alpha);
          const float dL_dalpha = T * cg - inv1ma * (F - R);
          T = T * (1.f - alpha);
          // Per-splat factors (opacity, conic, 0.5*W/H) are uniform over the pixels, so only the raw
          // moments of m = G * dL/dalpha about the splat centre are reduced; k_project_bwd turns them
          // into dL/d{opacity, mean2D, conic}.
          const float m = G * dL_dalpha;
          const float mx = m * dx, my = m * dy;
          v[0] = w * gr; v[1] = w * gg; v[2] = w * gb;     // dL/dcolour
          v[3] = m;                                        // S0  (= dL/dopacity)
          v[4] = mx;                                       // Sx
          v[5] = my;                                       // Sy
          v[6] = mx * dx;                                  // Sxx
          v[7] = mx * dy;                                  // Sxy
          v8 = my * dy;                                    // Syy
        } else {
#pragma unroll
          for (int k = 0; k < 8; k++) v[k] = 0.f;
          v8 = 0.f;
        }
        warp_reduce9(v, v8, lane);
        const uint32_t g = __float_as_uint(q1.w);
        const bool writer = ((lane & 3) == 0) || lane == 1;
        if (writer) {
          const int k = (lane == 1) ? 8 : (lane >> 2);
          atomicAdd(a.grad2d + (size_t)g * GRAD2D_STRIDE + k, (lane == 1) ? v8 : v[0]);
        }
      }
    }
    const int num_done = __syncthreads_count(c * CH + cnt >= ncontrib);
    if (num_done == 256) {
      ring.drain(c);
      break;
    }
    ring.refill(c);
  }
  if (MODE == GATHER_LDGSTS) cp_async_wait<0>();
}

// Frame export for the datagen sweep: CHW fp32 colour -> HWC 8-bit RGB (clamp to [0,1], round to
// nearest), 4 pixels per thread so every thread writes three aligned 32-bit words.
__global__ void __launch_bounds__(256) k_export_rgb8(const float* __restrict__ color, int H, int W,
                                                     uint8_t* __restrict__ out) {
  const size_t npx = (size_t)H * W;
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // group of 4 pixels
  const size_t p0 = q * 4;
  if (p0 >= npx) return;
  auto cvt = [](float v) { return rgb8_of(v); };
  if (p0 + 3 < npx && (npx & 3) == 0) {
    const float4 r = *reinterpret_cast<const float4*>(color + p0);
    const float4 g = *reinterpret_cast<const float4*>(color + npx + p0);
    const float4 b = *reinterpret_cast<const float4*>(color + 2 * npx + p0);
    const uint32_t w0 = cvt(r.x) | (cvt(g.x) << 8) | (cvt(b.x) << 16) | (cvt(r.y) << 24);
    const uint32_t w1 = cvt(g.y) | (cvt(b.y) << 8) | (cvt(r.z) << 16) | (cvt(g.z) << 24);
    const uint32_t w2 = cvt(b.z) | (cvt(r.w) << 8) | (cvt(g.w) << 16) | (cvt(b.w) << 24);
    uint32_t* o = reinterpret_cast<uint32_t*>(out + p0 * 3);
    o[0] = w0; o[1] = w1; o[2] = w2;
  } else {
    for (size_t p = p0; p < npx && p < p0 + 4; p++) {
      out[p * 3 + 0] = (uint8_t)cvt(color[p]);
      out[p * 3 + 1] = (uint8_t)cvt(color[npx + p]);
      out[p * 3 + 2] = (uint8_t)cvt(color[2 * npx + p]);
    }
  }
}

void launch_export_rgb8(const float* color, int H, int W, uint8_t* out, cudaStream_t st) {
  const size_t groups = ((size_t)H * W + 3) / 4;
  k_export_rgb8<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(color, H, W, out);
  count_launch();
}

static std::atomic<int> g_gather_mode{-1};

void set_gather_mode(int mode) { g_gather_mode.store(mode == GATHER_TMA ? (int)GATHER_TMA : (int)GATHER_LDGSTS); }

static int gather_mode() {
  int m = g_gather_mode.load();
  if (m < 0) {
    const char* e = getenv("B200GS_GATHER");
    m = (e && e[0] == 't') ? (int)GATHER_TMA : (int)GATHER_LDGSTS;
    g_gather_mode.store(m);
  }
  return m;
}

void launch_render(const RenderArgs& a, cudaStream_t st) {
  const dim3 grid((a.W + TILE - 1) / TILE, (a.H + TILE - 1) / TILE), block(256);
  if (gather_mode() == GATHER_TMA) k_render_fwd<GATHER_TMA><<<grid, block, 0, st>>>(a);
  else k_render_fwd<GATHER_LDGSTS><<<grid, block, 0, st>>>(a);
  count_launch();
}

void launch_render_bwd(const RenderBwdArgs& a, cudaStream_t st) {
  const dim3 grid((a.W + TILE - 1) / TILE, (a.H + TILE - 1) / TILE), block(256);
  if (gather_mode() == GATHER_TMA) k_render_bwd<GATHER_TMA><<<grid, block, 0, st>>>(a);
  else k_render_bwd<GATHER_LDGSTS><<<grid, block, 0, st>>>(a);
  count_launch();
}

}  // namespace b200gs

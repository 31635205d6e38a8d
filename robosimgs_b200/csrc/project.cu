// project.cu -- per-Gaussian kernels: projection + SH colour (forward), tile-pair emission,
// fused projection/SH/covariance adjoint (backward), near-plane visibility.
//
// Replaces: a3 preprocessCUDA fwd, a5 duplicateWithKeys, a10 computeCov2DCUDA bwd,
// a11 preprocessCUDA bwd, a12 checkFrustum -- of the public diff-gaussian-rasterization named by
// BASELINE.json:north_star (third-party; the reference repo only delegates, README.md:75).
// Arithmetic follows oracle/gs_oracle_impl.h step by step.
#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#include "spans.cuh"
#include "binsort.cuh"

namespace b200gs {

#ifndef PROJECT_MIN_BLOCKS
#define PROJECT_MIN_BLOCKS 3
#endif
#ifndef PROJECT_SH_BULK_PREFETCH
#define PROJECT_SH_BULK_PREFETCH 0
#endif
#ifndef PROJECT_WAVES
#define PROJECT_WAVES 2
#endif
#ifndef PROJECT_DENSE_WAVES
#define PROJECT_DENSE_WAVES 1
#endif
// SH row of the projection adjoint staged through the thread's slab row (LDGSTS issued with the first loads, read back
// behind the geometry): measured on C3 0.0999 -> 0.1084 ms (128-thread CTAs) -- the kernel does not wait for that round
// trip, and the copies cost shared-memory bandwidth.  Off.
#ifndef PROJECT_BWD_STAGE_SH
#define PROJECT_BWD_STAGE_SH 0
#endif
#ifndef PROJECT_BWD_EARLY_CLAMP    // clamp mask loaded with the first batch of loads: measured 0.1084 -> 0.1048 with staging on,
#define PROJECT_BWD_EARLY_CLAMP 0  // 0.1065 -> 0.111 without (256-thread CTAs)
#endif
// CTA size of the projection adjoint (its SH-gradient slab is THREADS x M x 12 bytes and leaves behind ONE barrier):
// 256 threads 0.1054 ms, 128: 0.0994, 64: 0.0968, 32: 0.0967 (same 24 warps per SM; the warps of a CTA wait for its slowest)
#ifndef PROJECT_BWD_THREADS
#define PROJECT_BWD_THREADS 64
#endif
#ifndef PROJECT_BWD_MIN_BLOCKS
#define PROJECT_BWD_MIN_BLOCKS (768 / PROJECT_BWD_THREADS)
#endif

struct CamConst {
  float v[16];
  float p[16];
  float cam[3];
};

__device__ __forceinline__ void load_cam(CamConst& c, const float* __restrict__ view,
                                         const float* __restrict__ proj,
                                         const float* __restrict__ campos) {
#pragma unroll
  for (int i = 0; i < 16; i++) {
    c.v[i] = __ldg(view + i);
    c.p[i] = __ldg(proj + i);
  }
#pragma unroll
  for (int i = 0; i < 3; i++) c.cam[i] = __ldg(campos + i);
}

// Sigma = R diag(mod*s)^2 R^T, 6 unique entries (xx,xy,xz,yy,yz,zz)
__device__ __forceinline__ void cov3d_from_scale_rot(float3 s, float mod, float4 q, float* cov) {
  const float r = q.x, x = q.y, y = q.z, z = q.w;
  const float R00 = 1.f - 2.f * (y * y + z * z), R01 = 2.f * (x * y - r * z), R02 = 2.f * (x * z + r * y);
  const float R10 = 2.f * (x * y + r * z), R11 = 1.f - 2.f * (x * x + z * z), R12 = 2.f * (y * z - r * x);
  const float R20 = 2.f * (x * z - r * y), R21 = 2.f * (y * z + r * x), R22 = 1.f - 2.f * (x * x + y * y);
  const float d0 = mod * s.x, d1 = mod * s.y, d2 = mod * s.z;
  const float m00 = R00 * d0, m01 = R01 * d1, m02 = R02 * d2;
  const float m10 = R10 * d0, m11 = R11 * d1, m12 = R12 * d2;
  const float m20 = R20 * d0, m21 = R21 * d1, m22 = R22 * d2;
  cov[0] = m00 * m00 + m01 * m01 + m02 * m02;
  cov[1] = m00 * m10 + m01 * m11 + m02 * m12;
  cov[2] = m00 * m20 + m01 * m21 + m02 * m22;
  cov[3] = m10 * m10 + m11 * m11 + m12 * m12;
  cov[4] = m10 * m20 + m11 * m21 + m12 * m22;
  cov[5] = m20 * m20 + m21 * m21 + m22 * m22;
}

struct Ewa {
  float t[3];       // view-space mean with the 1.3*tanfov clamp applied to x,y
  float m[2][3];    // J * Wr
  float xmask, ymask;
  float fx, fy;
};

__device__ __forceinline__ void ewa_jacobian(const CamConst& c, float3 mu, float W, float H,
                                             float tanfovx, float tanfovy, Ewa& e) {
  float tx = c.v[0] * mu.x + c.v[4] * mu.y + c.v[8] * mu.z + c.v[12];
  float ty = c.v[1] * mu.x + c.v[5] * mu.y + c.v[9] * mu.z + c.v[13];
  const float tz = c.v[2] * mu.x + c.v[6] * mu.y + c.v[10] * mu.z + c.v[14];
  const float limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
  const float txtz = tx / tz, tytz = ty / tz;
  e.xmask = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
  e.ymask = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
  tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
  ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
  e.t[0] = tx; e.t[1] = ty; e.t[2] = tz;
  e.fx = W / (2.f * tanfovx);
  e.fy = H / (2.f * tanfovy);
  const float j00 = e.fx / tz, j02 = -(e.fx * tx) / (tz * tz);
  const float j11 = e.fy / tz, j12 = -(e.fy * ty) / (tz * tz);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    e.m[0][k] = j00 * c.v[4 * k + 0] + j02 * c.v[4 * k + 2];
    e.m[1][k] = j11 * c.v[4 * k + 1] + j12 * c.v[4 * k + 2];
  }
}

// Sigma2D = M Sigma M^T + 0.3 I  ->  (a,b,c)
__device__ __forceinline__ void cov2d(const Ewa& e, const float* c3, float& a, float& b, float& c) {
  float ms[2][3];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    ms[r][0] = e.m[r][0] * c3[0] + e.m[r][1] * c3[1] + e.m[r][2] * c3[2];
    ms[r][1] = e.m[r][0] * c3[1] + e.m[r][1] * c3[3] + e.m[r][2] * c3[4];
    ms[r][2] = e.m[r][0] * c3[2] + e.m[r][1] * c3[4] + e.m[r][2] * c3[5];
  }
  a = ms[0][0] * e.m[0][0] + ms[0][1] * e.m[0][1] + ms[0][2] * e.m[0][2] + 0.3f;
  b = ms[0][0] * e.m[1][0] + ms[0][1] * e.m[1][1] + ms[0][2] * e.m[1][2];
  c = ms[1][0] * e.m[1][0] + ms[1][1] * e.m[1][1] + ms[1][2] * e.m[1][2] + 0.3f;
}

// Records whose alpha needs the general evaluation (see common.cuh, q2.w): opacity above the 0.99 clamp, or a conic
// whose quadratic form is not positive by a margin that dwarfs fp32 rounding (then `power` > 0 can occur and such
// pixels are skipped, as in the public algorithm).  For all others alpha = o * 2^power2 <= o <= 0.99 and power <= 0
// hold for every pixel, so the compositing kernels can drop both tests without changing a bit.
__device__ __forceinline__ bool record_is_general(float A, float B, float C, float o) {
  const bool safe = A > 0.f && C > 0.f && (A * C - B * B) > 1e-4f * (A * C);
  return o > 0.99f || !safe;
}

#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f
__device__ __constant__ float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                          -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f,  -0.4570457994644658f,
                                          0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                          -0.5900435899266435f};

template <int DEG>
__device__ __forceinline__ void sh_basis(float x, float y, float z, float* b) {
  b[0] = SH_C0;
  if (DEG > 0) {
    b[1] = -SH_C1 * y; b[2] = SH_C1 * z; b[3] = -SH_C1 * x;
  }
  if (DEG > 1) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = SH_C2[0] * xy;
    b[5] = SH_C2[1] * yz;
    b[6] = SH_C2[2] * (2.f * zz - xx - yy);
    b[7] = SH_C2[3] * xz;
    b[8] = SH_C2[4] * (xx - yy);
    if (DEG > 2) {
      b[9] = SH_C3[0] * y * (3.f * xx - yy);
      b[10] = SH_C3[1] * xy * z;
      b[11] = SH_C3[2] * y * (4.f * zz - xx - yy);
      b[12] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
      b[13] = SH_C3[4] * x * (4.f * zz - xx - yy);
      b[14] = SH_C3[5] * z * (xx - yy);
      b[15] = SH_C3[6] * x * (xx - 3.f * yy);
    }
  }
}

// Load the first NF floats of a Gaussian's SH row.  Rows of M*3 floats are 16-byte aligned when
// M % 4 == 0 (M = 16 for degree-3 storage): float4 path, one 128-bit load per 4 floats.
template <int NF>
__device__ __forceinline__ void load_sh_row(const float* __restrict__ row, bool vec_ok, float* f) {
  if (vec_ok) {
    constexpr int NV = (NF + 3) / 4;
    const float4* r4 = reinterpret_cast<const float4*>(row);
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = __ldg(r4 + i);
#pragma unroll
    for (int i = 0; i < NV; i++) {
      if (4 * i + 0 < NF) f[4 * i + 0] = v[i].x;
      if (4 * i + 1 < NF) f[4 * i + 1] = v[i].y;
      if (4 * i + 2 < NF) f[4 * i + 2] = v[i].z;
      if (4 * i + 3 < NF) f[4 * i + 3] = v[i].w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NF; i++) f[i] = __ldg(row + i);
  }
}

// ==================================================================================================
// K1: projection + SH colour.  One thread per Gaussian.  DEG = -1: colours are precomputed.
// ==================================================================================================
// The Gaussian's dense parameters (44 bytes) are fetched one grid stride AHEAD of their use: k_project is a grid-stride
// loop whose threads start the copy of their next Gaussian into a private shared-memory slot (LDGSTS, no register
// holds the data in flight) before they work on the current one, so the two dependent DRAM round trips in front of
// the geometry (means -> near-plane test -> scales / rotation / opacity) leave the critical path.  Scales, rotation
// and opacity of a Gaussian behind the near plane are fetched for nothing -- they share 32-byte sectors with their
// neighbours, the DRAM traffic is the same.   Slot: {mu.xyz, opacity} {s.xyz, -} {q}.
struct ProjIn {
  float3 mu, s;
  float4 q;
  float o;
};
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void fetch_proj_in(const ProjectArgs& a, int i, float4* slot) {
  float* f = reinterpret_cast<float*>(slot);
  const float* m = a.means + 3 * (size_t)i;
  cp_async4(f + 0, m); cp_async4(f + 1, m + 1); cp_async4(f + 2, m + 2);
  cp_async4(f + 3, a.opac + i);
  if (!a.cov3d_precomp) {
    const float* sc = a.scales + 3 * (size_t)i;
    cp_async4(f + 4, sc); cp_async4(f + 5, sc + 1); cp_async4(f + 6, sc + 2);
    cp_async16(slot + 2, reinterpret_cast<const float4*>(a.rots) + i);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ ProjIn read_proj_in(const float4* slot) {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  const float4 v0 = slot[0], v1 = slot[1], v2 = slot[2];
  ProjIn in;
  in.mu = make_float3(v0.x, v0.y, v0.z);
  in.o = v0.w;
  in.s = make_float3(v1.x, v1.y, v1.z);
  in.q = v2;
  return in;
}

template <int DEG>
__global__ void __launch_bounds__(256, PROJECT_MIN_BLOCKS) k_project(ProjectArgs a) {
  __shared__ __align__(16) float4 s_in[256][3];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.P) return;
  float4* const slot = s_in[threadIdx.x];
  fetch_proj_in(a, i, slot);
  CamConst c;
  load_cam(c, a.view, a.proj, a.campos);
  const int stride = gridDim.x * blockDim.x;
  for (; i < a.P; i += stride) {
  const ProjIn cur = read_proj_in(slot);
  if (i + stride < a.P) fetch_proj_in(a, i + stride, slot);

  uint32_t key = 0xFFFFFFFFu, ntiles = 0, tiles_word = 0;
  int radius = 0;

  const float3 mu = cur.mu;
  const float vz = c.v[2] * mu.x + c.v[6] * mu.y + c.v[10] * mu.z + c.v[14];
  if (vz > a.near_plane) {
    const float hx = c.p[0] * mu.x + c.p[4] * mu.y + c.p[8] * mu.z + c.p[12];
    const float hy = c.p[1] * mu.x + c.p[5] * mu.y + c.p[9] * mu.z + c.p[13];
    const float hw = c.p[3] * mu.x + c.p[7] * mu.y + c.p[11] * mu.z + c.p[15];
    const float pw = 1.f / (hw + 0.0000001f);
    const float ndcx = hx * pw, ndcy = hy * pw;

    float c3[6];
    const float o = cur.o;
    if (a.cov3d_precomp) {
#pragma unroll
      for (int k = 0; k < 6; k++) c3[k] = __ldg(a.cov3d_precomp + 6 * (size_t)i + k);
    } else {
      cov3d_from_scale_rot(cur.s, a.scale_modifier, cur.q, c3);
    }
    Ewa e;
    ewa_jacobian(c, mu, (float)a.W, (float)a.H, a.tanfovx, a.tanfovy, e);
    float ca, cb, cc;
    cov2d(e, c3, ca, cb, cc);
    const float det = ca * cc - cb * cb;
    if (det != 0.f) {
      const float det_inv = 1.f / det;
      const float A = cc * det_inv, B = -cb * det_inv, C = ca * det_inv;
      const float mid = 0.5f * (ca + cc);
      const float disc = sqrtf(fmaxf(0.1f, mid * mid - det));
      const float rad = ceilf(3.f * sqrtf(fmaxf(mid + disc, mid - disc)));
      const float px = ((ndcx + 1.f) * a.W - 1.f) * 0.5f;
      const float py = ((ndcy + 1.f) * a.H - 1.f) * 0.5f;
      const int irad = (int)rad;
      const TileRect r = reference_rect(px, py, irad, a.gx, a.gy);
      if ((r.x1 - r.x0) * (r.y1 - r.y0) != 0) {
        radius = irad;
        if constexpr (DEG > 0) {
          // the splat is on screen: start its SH row (up to 192 B = two lines) towards L2 now, the span walk and
          // the bucket counters below hide the DRAM latency the 12 loads would otherwise wait for
          const char* row = reinterpret_cast<const char*>(a.shs + (size_t)i * a.M * 3);
#if PROJECT_SH_BULK_PREFETCH
          if (a.sh_vec) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(row), "r"(((DEG + 1) * (DEG + 1) * 12 + 15) & ~15) : "memory");
#else
          asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 128));
#endif
        }
        // 0.5*q <= thr  <=>  o*exp(-0.5 q) >= 1/255 ; slack keeps the test conservative
        const float thr = __logf(255.f * o) + 0.01f;
        uint32_t* cnt = nullptr;
        if (a.bucket_count) {     // bucketed binning: the splat's depth slice of every bin it touches
          const uint32_t db = __float_as_uint(vz);
          const uint32_t rel = db > a.near_bits ? db - a.near_bits : 0u;
          cnt = a.bucket_count + min(rel >> a.slice_shift, (1u << a.slices_log2) - 1u);
        }
        ntiles = (o > 0.f) ? count_tiles(px, py, A, B, C, thr, bin_rect(r, a.bin_shift), a.bin_shift, cnt, a.slices_log2, a.gbx,
                                         a.pack_tiles != 0, &tiles_word) : 0u;
        if (ntiles > 0) {
          float rgb[3];
          uint32_t clampbits = 0;
          if constexpr (DEG < 0) {
            rgb[0] = __ldg(a.colors_precomp + 3 * (size_t)i);
            rgb[1] = __ldg(a.colors_precomp + 3 * (size_t)i + 1);
            rgb[2] = __ldg(a.colors_precomp + 3 * (size_t)i + 2);
          } else {
            constexpr int NB = (DEG < 0 ? 0 : (DEG + 1) * (DEG + 1));
            constexpr int NF = NB * 3;
            float f[NF > 0 ? NF : 1];
            load_sh_row<NF>(a.shs + (size_t)i * a.M * 3, a.sh_vec != 0, f);
            float dx = mu.x - c.cam[0], dy = mu.y - c.cam[1], dz = mu.z - c.cam[2];
            const float inv = 1.f / sqrtf(dx * dx + dy * dy + dz * dz);
            float b[NB > 0 ? NB : 1];
            sh_basis<DEG>(dx * inv, dy * inv, dz * inv, b);
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
              float acc = 0.f;
#pragma unroll
              for (int k = 0; k < NB; k++) acc += b[k] * f[3 * k + ch];
              acc += 0.5f;
              if (acc < 0.f) clampbits |= (1u << ch);
              rgb[ch] = fmaxf(acc, 0.f);
            }
          }
          a.clamped[i] = (uint8_t)clampbits;
          key = __float_as_uint(vz);
          float4* rec = a.rec + (size_t)i * REC_F4;
          rec[0] = make_float4(px, py, A, B);
          rec[1] = make_float4(C, o, thr, __uint_as_float((uint32_t)i));
          rec[2] = make_float4(rgb[0], rgb[1], rgb[2], record_is_general(A, B, C, o) ? -(float)irad : (float)irad);
        }
      }
    }
  }
  a.radii[i] = radius;
  a.depth_key[i] = key;
  a.tiles[i] = tiles_word;
  }
}

// ==================================================================================================
// K1 with DENSE WARPS (option "project" = 1; the default is k_project above).  Only a third of a scene's Gaussians
// reach the image, scattered at random over the index space, so in k_project every warp walks the whole visible path
// with a third of its lanes (ncu: 28 M warp-instructions on C3 at 67 % issue-active once the loads are off the critical
// path).  Here a warp owns PD_CHUNK consecutive Gaussians per step of a grid-stride loop and works in stages, with
// warp-private queues in shared memory between them (ballot + popc compaction, __syncwarp only, no block barrier):
//   0. the chunk's dense parameters (44 B per Gaussian) arrive in shared memory by coalesced 16-byte LDGSTS; the
//      lines of the warp's NEXT chunk are prefetched to L2 at the same time;
//   1. cull: near plane, then a CONSERVATIVE screen test -- radius <= 3 sqrt(|J|_F^2 |W|^2 |R|^2 s_max^2 + 0.62) + 1
//      from the Jacobian's Frobenius norm and the largest scale (on C3 it passes 37.5 % of the Gaussians against
//      37.0 % that really reach the image); the culled ones get their defaults, the survivors join the GEOMETRY QUEUE
//      with their parameters;
//   2. geometry for FULL batches of 32 queued Gaussians (the remainder is carried to the warp's next chunk): conic,
//      radius, exact rect test, tight spans, bucket counters, record q0/q1, L2 prefetch of the SH row; those that
//      touch a bin join the COLOUR QUEUE;
//   3. colour, again 32 at a time: the 12 x 128-bit SH loads of all 32 lanes in flight at once.
// Same expressions per Gaussian as k_project (the compiler contracts them per kernel: records agree to the last bit
// or two, radii and pair lists exactly -- tested).
// Measured on C3 (B200): every lane busy, but the kernel got SLOWER -- 0.058 ms (chunk 32) / 0.063 (chunk 64) against
// 0.047 ms for k_project; a first version without carried queues (one index queue per chunk, 62 % of the lanes busy
// in the geometry stage, 26.0 M warp-instructions instead of 28.0 M) took 0.052 ms.  The stages of a warp are strictly
// serial (copy -> cull -> geometry -> colour), each with its own exposed L2 latency, and the cull + queue traffic cost
// what the dense lanes save.  Kept as the second implementation the parity tests cross-check.
// ==================================================================================================
#ifndef PROJECT_DENSE_CHUNK
#define PROJECT_DENSE_CHUNK 32
#endif
constexpr int PD_CHUNK = PROJECT_DENSE_CHUNK;       // 32 or 64
constexpr int PD_WARPS = 4;
constexpr int PD_GQ = PD_CHUNK + 32;      // geometry queue: up to 31 carried entries + one chunk of survivors
constexpr int PD_CQ = 64;                 // colour queue: up to 31 carried entries + one batch
#ifndef PROJECT_DENSE_MIN_BLOCKS
#define PROJECT_DENSE_MIN_BLOCKS 6
#endif

template <int DEG>
__global__ void __launch_bounds__(32 * PD_WARPS, PROJECT_DENSE_MIN_BLOCKS) k_project_dense(ProjectArgs a) {
  __shared__ __align__(16) float s_mu[PD_WARPS][PD_CHUNK * 3];
  __shared__ __align__(16) float s_sc[PD_WARPS][PD_CHUNK * 3];
  __shared__ __align__(16) float4 s_rot[PD_WARPS][PD_CHUNK];
  __shared__ __align__(16) float s_op[PD_WARPS][PD_CHUNK];
  __shared__ __align__(16) float4 s_gq[PD_WARPS][PD_GQ][3];   // {mu, o} {s, id} {q}
  __shared__ __align__(16) float4 s_cq[PD_WARPS][PD_CQ];      // {mu, id}
  __shared__ uint32_t s_cm[PD_WARPS][PD_CQ];                  // radius | general << 31
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nchunks = (a.P + PD_CHUNK - 1) / PD_CHUNK;
  const int wstride = gridDim.x * PD_WARPS;
  int chunk = blockIdx.x * PD_WARPS + warp;
  if (chunk >= nchunks) return;
  CamConst c;
  load_cam(c, a.view, a.proj, a.campos);
  const uint32_t lt = (1u << lane) - 1u;
  const bool vec16 = ((reinterpret_cast<uintptr_t>(a.means) | reinterpret_cast<uintptr_t>(a.scales) |
                       reinterpret_cast<uintptr_t>(a.opac)) & 15u) == 0u;
  // |W|_2^2 of the view rotation: 1 for a rigid camera (checked), its Frobenius norm otherwise
  float wr2;
  {
    const float d00 = c.v[0] * c.v[0] + c.v[1] * c.v[1] + c.v[2] * c.v[2];
    const float d11 = c.v[4] * c.v[4] + c.v[5] * c.v[5] + c.v[6] * c.v[6];
    const float d22 = c.v[8] * c.v[8] + c.v[9] * c.v[9] + c.v[10] * c.v[10];
    const float d01 = c.v[0] * c.v[4] + c.v[1] * c.v[5] + c.v[2] * c.v[6];
    const float d02 = c.v[0] * c.v[8] + c.v[1] * c.v[9] + c.v[2] * c.v[10];
    const float d12 = c.v[4] * c.v[8] + c.v[5] * c.v[9] + c.v[6] * c.v[10];
    const float dev = fmaxf(fmaxf(fabsf(d00 - 1.f), fabsf(d11 - 1.f)),
                            fmaxf(fabsf(d22 - 1.f), fmaxf(fabsf(d01), fmaxf(fabsf(d02), fabsf(d12)))));
    wr2 = (dev < 1e-4f) ? 1.001f : (d00 + d11 + d22);
  }
  const float fx = (float)a.W / (2.f * a.tanfovx), fy = (float)a.H / (2.f * a.tanfovy);
  const float limx = 1.3f * a.tanfovx, limy = 1.3f * a.tanfovy;
  const float xmax = 16.f * (float)a.gx, ymax = 16.f * (float)a.gy;
  uint32_t ng = 0, nc = 0;       // entries waiting in the geometry / colour queue (warp-uniform)

  // ---- stage 3: colour of the first nv entries of the colour queue ----
  auto colour_batch = [&](uint32_t nv) {
    if ((uint32_t)lane < nv) {
      const float4 g = s_cq[warp][lane];
      const uint32_t meta = s_cm[warp][lane];
      const int i = (int)__float_as_uint(g.w);
      const float frad = (meta >> 31) ? -(float)(meta & 0x7FFFFFFFu) : (float)(meta & 0x7FFFFFFFu);
      float rgb[3];
      uint32_t clampbits = 0;
      if constexpr (DEG < 0) {
        rgb[0] = __ldg(a.colors_precomp + 3 * (size_t)i);
        rgb[1] = __ldg(a.colors_precomp + 3 * (size_t)i + 1);
        rgb[2] = __ldg(a.colors_precomp + 3 * (size_t)i + 2);
      } else {
        constexpr int NB = (DEG < 0 ? 0 : (DEG + 1) * (DEG + 1));
        constexpr int NF = NB * 3;
        float f[NF > 0 ? NF : 1];
        load_sh_row<NF>(a.shs + (size_t)i * a.M * 3, a.sh_vec != 0, f);
        float dx = g.x - c.cam[0], dy = g.y - c.cam[1], dz = g.z - c.cam[2];
        const float inv = 1.f / sqrtf(dx * dx + dy * dy + dz * dz);
        float b[NB > 0 ? NB : 1];
        sh_basis<DEG>(dx * inv, dy * inv, dz * inv, b);
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
          float acc = 0.f;
#pragma unroll
          for (int kk = 0; kk < NB; kk++) acc += b[kk] * f[3 * kk + ch];
          acc += 0.5f;
          if (acc < 0.f) clampbits |= (1u << ch);
          rgb[ch] = fmaxf(acc, 0.f);
        }
      }
      a.clamped[i] = (uint8_t)clampbits;
      a.rec[(size_t)i * REC_F4 + 2] = make_float4(rgb[0], rgb[1], rgb[2], frad);
    }
  };
  // ---- stage 2: geometry of nv entries of the geometry queue starting at `off`; survivors join the colour queue ----
  auto geom_batch = [&](uint32_t off, uint32_t nv) {
    uint32_t ntiles = 0, meta = 0;
    float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f);
    int i = 0;
    if ((uint32_t)lane < nv) {
      e0 = s_gq[warp][off + lane][0];
      const float4 e1 = s_gq[warp][off + lane][1], e2 = s_gq[warp][off + lane][2];
      i = (int)__float_as_uint(e1.w);
      uint32_t key = 0xFFFFFFFFu, tiles_word = 0;
      int radius = 0;
      const float3 mu = make_float3(e0.x, e0.y, e0.z);
      const float vz = c.v[2] * mu.x + c.v[6] * mu.y + c.v[10] * mu.z + c.v[14];
      const float hx = c.p[0] * mu.x + c.p[4] * mu.y + c.p[8] * mu.z + c.p[12];
      const float hy = c.p[1] * mu.x + c.p[5] * mu.y + c.p[9] * mu.z + c.p[13];
      const float hw = c.p[3] * mu.x + c.p[7] * mu.y + c.p[11] * mu.z + c.p[15];
      const float pw = 1.f / (hw + 0.0000001f);
      const float ndcx = hx * pw, ndcy = hy * pw;
      float c3[6];
      cov3d_from_scale_rot(make_float3(e1.x, e1.y, e1.z), a.scale_modifier, e2, c3);
      Ewa e;
      ewa_jacobian(c, mu, (float)a.W, (float)a.H, a.tanfovx, a.tanfovy, e);
      float ca, cb, cc;
      cov2d(e, c3, ca, cb, cc);
      const float det = ca * cc - cb * cb;
      if (det != 0.f) {
        const float det_inv = 1.f / det;
        const float A = cc * det_inv, B = -cb * det_inv, C = ca * det_inv;
        const float mid = 0.5f * (ca + cc);
        const float disc = sqrtf(fmaxf(0.1f, mid * mid - det));
        const float rad = ceilf(3.f * sqrtf(fmaxf(mid + disc, mid - disc)));
        const float px = ((ndcx + 1.f) * a.W - 1.f) * 0.5f;
        const float py = ((ndcy + 1.f) * a.H - 1.f) * 0.5f;
        const int irad = (int)rad;
        const TileRect r = reference_rect(px, py, irad, a.gx, a.gy);
        if ((r.x1 - r.x0) * (r.y1 - r.y0) != 0) {
          radius = irad;
          if constexpr (DEG > 0) {
            const char* row = reinterpret_cast<const char*>(a.shs + (size_t)i * a.M * 3);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 128));
          }
          const float o = e0.w;
          const float thr = __logf(255.f * o) + 0.01f;
          uint32_t* bcnt = nullptr;
          if (a.bucket_count) {
            const uint32_t db = __float_as_uint(vz);
            const uint32_t rel = db > a.near_bits ? db - a.near_bits : 0u;
            bcnt = a.bucket_count + min(rel >> a.slice_shift, (1u << a.slices_log2) - 1u);
          }
          ntiles = (o > 0.f) ? count_tiles(px, py, A, B, C, thr, bin_rect(r, a.bin_shift), a.bin_shift, bcnt, a.slices_log2, a.gbx,
                                           a.pack_tiles != 0, &tiles_word) : 0u;
          if (ntiles > 0) {
            key = __float_as_uint(vz);
            meta = (uint32_t)irad | (record_is_general(A, B, C, o) ? 0x80000000u : 0u);
            float4* rec = a.rec + (size_t)i * REC_F4;
            rec[0] = make_float4(px, py, A, B);
            rec[1] = make_float4(C, o, thr, __uint_as_float((uint32_t)i));
          }
        }
      }
      a.radii[i] = radius;
      a.depth_key[i] = key;
      a.tiles[i] = tiles_word;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, ntiles > 0);
    if (ntiles > 0) {
      const uint32_t slot = nc + __popc(m & lt);
      s_cq[warp][slot] = make_float4(e0.x, e0.y, e0.z, __uint_as_float((uint32_t)i));
      s_cm[warp][slot] = meta;
    }
    nc += __popc(m);
    __syncwarp();
    if (nc >= 32u) {
      colour_batch(32u);
      const uint32_t left = nc - 32u;
      float4 tg = make_float4(0.f, 0.f, 0.f, 0.f);
      uint32_t tm = 0;
      if ((uint32_t)lane < left) { tg = s_cq[warp][32 + lane]; tm = s_cm[warp][32 + lane]; }
      __syncwarp();
      if ((uint32_t)lane < left) { s_cq[warp][lane] = tg; s_cm[warp][lane] = tm; }
      nc = left;
      __syncwarp();
    }
  };

  for (; chunk < nchunks; chunk += wstride) {
    const int base = chunk * PD_CHUNK;
    const int cnt = min(PD_CHUNK, a.P - base);
    // ---- stage 0: chunk -> shared memory ----
    if (vec16 && cnt == PD_CHUNK) {
      const float4* gm = reinterpret_cast<const float4*>(a.means + 3 * (size_t)base);
      const float4* gs = reinterpret_cast<const float4*>(a.scales + 3 * (size_t)base);
      const float4* gq = reinterpret_cast<const float4*>(a.rots) + base;
      const float4* go = reinterpret_cast<const float4*>(a.opac + base);
      float4* dm = reinterpret_cast<float4*>(s_mu[warp]);
      float4* ds = reinterpret_cast<float4*>(s_sc[warp]);
      float4* dop = reinterpret_cast<float4*>(s_op[warp]);
#pragma unroll
      for (int e = lane; e < PD_CHUNK * 3 / 4; e += 32) { cp_async16(dm + e, gm + e); cp_async16(ds + e, gs + e); }
#pragma unroll
      for (int e = lane; e < PD_CHUNK; e += 32) cp_async16(s_rot[warp] + e, gq + e);
      if (lane < PD_CHUNK / 4) cp_async16(dop + lane, go + lane);
    } else {
      for (int e = lane; e < cnt * 3; e += 32) {
        cp_async4(&s_mu[warp][e], a.means + 3 * (size_t)base + e);
        cp_async4(&s_sc[warp][e], a.scales + 3 * (size_t)base + e);
      }
      for (int e = lane; e < cnt; e += 32) {
        cp_async16(&s_rot[warp][e], reinterpret_cast<const float4*>(a.rots) + base + e);
        cp_async4(&s_op[warp][e], a.opac + base + e);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    {   // the warp's next chunk: 44 B per Gaussian in lines of 128 B towards L2
      constexpr int LM = PD_CHUNK * 12 / 128, LR = PD_CHUNK * 16 / 128, LO = (PD_CHUNK * 4 + 127) / 128;
      const int nb = base + wstride * PD_CHUNK;
      if (nb + PD_CHUNK <= a.P && lane < 2 * LM + LR + LO) {
        const char* ptr = lane < LM       ? reinterpret_cast<const char*>(a.means + 3 * (size_t)nb) + 128 * lane
                          : lane < 2 * LM ? reinterpret_cast<const char*>(a.scales + 3 * (size_t)nb) + 128 * (lane - LM)
                          : lane < 2 * LM + LR ? reinterpret_cast<const char*>(reinterpret_cast<const float4*>(a.rots) + nb) + 128 * (lane - 2 * LM)
                                               : reinterpret_cast<const char*>(a.opac + nb) + 128 * (lane - 2 * LM - LR);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    // ---- stage 1: cull; survivors join the geometry queue with their parameters ----
#pragma unroll
    for (int it = 0; it < PD_CHUNK / 32; it++) {
      const int k = it * 32 + lane;
      bool pass = false;
      float mx = 0.f, my = 0.f, mz = 0.f;
      if (k < cnt) {
        mx = s_mu[warp][3 * k]; my = s_mu[warp][3 * k + 1]; mz = s_mu[warp][3 * k + 2];
        const float tz = c.v[2] * mx + c.v[6] * my + c.v[10] * mz + c.v[14];
        if (tz > a.near_plane) {
          pass = true;
          const float hx = c.p[0] * mx + c.p[4] * my + c.p[8] * mz + c.p[12];
          const float hy = c.p[1] * mx + c.p[5] * my + c.p[9] * mz + c.p[13];
          const float hw = c.p[3] * mx + c.p[7] * my + c.p[11] * mz + c.p[15];
          const float pw = rcp_fast(hw + 0.0000001f);
          const float px = ((hx * pw + 1.f) * (float)a.W - 1.f) * 0.5f;
          const float py = ((hy * pw + 1.f) * (float)a.H - 1.f) * 0.5f;
          const float tx = c.v[0] * mx + c.v[4] * my + c.v[8] * mz + c.v[12];
          const float ty = c.v[1] * mx + c.v[5] * my + c.v[9] * mz + c.v[13];
          const float itz = rcp_fast(tz);
          const float txc = fminf(limx, fmaxf(-limx, tx * itz)), tyc = fminf(limy, fmaxf(-limy, ty * itz));
          const float jf2 = (fx * fx * (1.f + txc * txc) + fy * fy * (1.f + tyc * tyc)) * itz * itz;
          const float smax = a.scale_modifier * fmaxf(fabsf(s_sc[warp][3 * k]), fmaxf(fabsf(s_sc[warp][3 * k + 1]), fabsf(s_sc[warp][3 * k + 2])));
          const float4 q = s_rot[warp][k];
          const float qn = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
          const float rn = (fabsf(qn - 1.f) < 1e-4f) ? 1.001f : (1.f + 2.f * qn);
          const float lam = jf2 * wr2 * (rn * rn) * (smax * smax) * 1.01f + 0.62f;
          const float rb = 3.f * sqrt_approx(lam) * 1.001f + 1.6f;      // >= ceil(3 sqrt(lambda_1)) + 0.5 px of slack
          // the reference rect is empty when px + r < 1 or px - r >= 16 gx (same in y): certain here -> radius 0
          if (px + rb < 1.f || px - rb >= xmax || py + rb < 1.f || py - rb >= ymax) pass = false;
        }
        if (!pass) { a.radii[base + k] = 0; a.depth_key[base + k] = 0xFFFFFFFFu; a.tiles[base + k] = 0u; }
      }
      const uint32_t m = __ballot_sync(0xffffffffu, pass);
      if (pass) {
        float4* e = s_gq[warp][ng + __popc(m & lt)];
        e[0] = make_float4(mx, my, mz, s_op[warp][k]);
        e[1] = make_float4(s_sc[warp][3 * k], s_sc[warp][3 * k + 1], s_sc[warp][3 * k + 2], __uint_as_float((uint32_t)(base + k)));
        e[2] = s_rot[warp][k];
      }
      ng += __popc(m);
    }
    __syncwarp();
    // ---- stage 2 on full batches; the remainder (< 32 entries) is carried to the warp's next chunk ----
    uint32_t off = 0;
    for (; ng - off >= 32u; off += 32u) geom_batch(off, 32u);
    if (off) {
      const uint32_t left = ng - off;
      float4 t0, t1, t2;
      if ((uint32_t)lane < left) { t0 = s_gq[warp][off + lane][0]; t1 = s_gq[warp][off + lane][1]; t2 = s_gq[warp][off + lane][2]; }
      __syncwarp();
      if ((uint32_t)lane < left) { s_gq[warp][lane][0] = t0; s_gq[warp][lane][1] = t1; s_gq[warp][lane][2] = t2; }
      ng = left;
    }
    __syncwarp();     // queues and the parameter stage are rewritten by the next step
  }
  // ---- flush ----
  if (ng) geom_batch(0u, ng);
  if (nc) colour_batch(nc);
}

// ==================================================================================================
// K3: emit ((bin << 32) | depth bits, Gaussian id) pairs in index order; offsets[] is the inclusive
// scan of the per-Gaussian bin counts.
// ==================================================================================================
// Pair key.  64-bit form: (bin << 32) | depth bits -- the public algorithm's key.  32-bit form (bins <=
// 255): (bin << 24) | q24, q24 = min((depth bits - near-plane bits) >> 3, 2^24 - 1), a monotone
// quantisation of the positive-float depth order that only merges depths within 8 ulps (or beyond
// 2^16 x near).  A stable sort on it needs FOUR 8-bit passes over 8-byte pairs instead of five over
// 12-byte pairs; k_tile_ranges32 then restores the exact (depth bits, index) order inside the rare runs of
// equal keys, so the result is identical.
__device__ __forceinline__ uint32_t quant24(uint32_t depth_bits, uint32_t near_bits) {
  const uint32_t d = depth_bits > near_bits ? depth_bits - near_bits : 0u;
  return min(d >> 3, 0xFFFFFFu);
}
__device__ __forceinline__ void put_key(const EmitArgs& a, uint32_t o, uint32_t bin, uint32_t depth_bits) {
  if (a.key32) reinterpret_cast<uint32_t*>(a.keys)[o] = (bin << 24) | quant24(depth_bits, a.near_bits);
  else a.keys[o] = ((uint64_t)bin << 32) | depth_bits;
}
__device__ __forceinline__ void put_invalid_key(const EmitArgs& a, uint32_t o) {
  if (a.key32) reinterpret_cast<uint32_t*>(a.keys)[o] = a.invalid_tile << 24;
  else a.keys[o] = (uint64_t)a.invalid_tile << 32;
}

__global__ void __launch_bounds__(256) k_emit_pairs(EmitArgs a) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t cap = a.capacity;
  const uint32_t g = (uint32_t)r;
  uint32_t n = 0, off = 0;
  float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
  int radius = 0;
  uint32_t depth = 0;
  if (r < a.P) {
    n = a.tiles[r];
    if (n) {
      off = a.offsets[r] - n;
      q0 = a.rec[(size_t)r * REC_F4];
      q1 = a.rec[(size_t)r * REC_F4 + 1];
      radius = a.radii[r];
      depth = a.depth_key[r];
    }
  }
  // ---- small footprints: one thread writes its own pairs ----
  const bool big = n > EMIT_BIG_THRESHOLD;
  if (n > 0 && !big) {
    const uint32_t end = off + n;
    uint32_t o = off;
    const TileRect rect = bin_rect(reference_rect(q0.x, q0.y, radius, a.gx, a.gy), a.bin_shift);
    SpanCtx s;
    if (span_setup(s, q0.x, q0.y, q0.z, q0.w, q1.x, q1.z, rect, a.bin_shift)) {
      for (int ty = s.ty0; ty < s.ty1; ty++) {
        int c0, c1;
        row_span(s, rect, ty, c0, c1);
        for (int tx = c0; tx < c1 && o < end; tx++, o++) {
          if (o < cap) {
            put_key(a, o, (uint32_t)(ty * a.gbx + tx), depth);
            a.vals[o] = g;
          }
        }
      }
    }
    // defensive: never leave unwritten slots (count and emit share row_span, so o == end)
    for (; o < end; o++) {
      if (o < cap) { put_invalid_key(a, o); a.vals[o] = g; }
    }
  }
  // ---- large footprints go to a global queue; k_emit_big emits them one warp per Gaussian so a
  // few screen-filling splats cannot serialise a warp ----
  if (big) a.big_queue[atomicAdd(a.big_count, 1u)] = (uint32_t)r;
  // ---- speculative capacity: pad [D, capacity) with the invalid tile id so the sort can run on a
  // host-known item count while D is still on the device ----
  if (cap > 0) {
    const uint32_t D = a.P > 0 ? a.offsets[a.P - 1] : 0u;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = D + (uint32_t)r; i < cap; i += stride) {
      put_invalid_key(a, i);
      a.vals[i] = 0u;
    }
  }
}

// One warp per queued large-footprint Gaussian: lanes compute the spans of 32 tile rows in parallel,
// a warp scan turns the span lengths into output offsets, then each row's tiles are written with
// coalesced 32-wide stores.
__global__ void __launch_bounds__(256) k_emit_big(EmitArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t count = *a.big_count;
  const uint32_t cap = a.capacity;
  for (uint32_t w = warp_global; w < count; w += nwarps) {
    const uint32_t g = a.big_queue[w];
    const uint32_t n = a.tiles[g];
    const uint32_t off = a.offsets[g] - n, end = off + n;
    const uint32_t depth = a.depth_key[g];
    const float4 q0 = a.rec[(size_t)g * REC_F4], q1 = a.rec[(size_t)g * REC_F4 + 1];
    const TileRect rect = bin_rect(reference_rect(q0.x, q0.y, a.radii[g], a.gx, a.gy), a.bin_shift);
    SpanCtx s;
    const bool ok = span_setup(s, q0.x, q0.y, q0.z, q0.w, q1.x, q1.z, rect, a.bin_shift);
    uint32_t o = off;
    for (int y_base = s.ty0; ok && y_base < s.ty1; y_base += 32) {
      const int ty = y_base + lane;
      int c0 = 0, c1 = 0;
      if (ty < s.ty1) row_span(s, rect, ty, c0, c1);
      const uint32_t len = (uint32_t)(c1 - c0);
      uint32_t incl = len;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      const uint32_t row_off = o + incl - len;
      const int rows = min(32, s.ty1 - y_base);
      for (int i = 0; i < rows; i++) {
        const uint32_t l_i = __shfl_sync(0xffffffffu, len, i);
        if (l_i == 0) continue;
        const uint32_t o_i = __shfl_sync(0xffffffffu, row_off, i);
        const int c_i = __shfl_sync(0xffffffffu, c0, i);
        const uint32_t tile0 = (uint32_t)((y_base + i) * a.gbx + c_i);
        for (uint32_t k = lane; k < l_i; k += 32) {
          const uint32_t idx = o_i + k;
          if (idx < end && idx < cap) {
            put_key(a, idx, tile0 + k, depth);
            a.vals[idx] = g;
          }
        }
      }
      o += __shfl_sync(0xffffffffu, incl, 31);
    }
    for (uint32_t idx = min(o, end) + lane; idx < end; idx += 32) {
      if (idx < cap) { put_invalid_key(a, idx); a.vals[idx] = g; }
    }
  }
}

// ==================================================================================================
// K5: per-tile [start,end) in the sorted pair list
// ==================================================================================================
__global__ void __launch_bounds__(256) k_tile_ranges(RangesArgs a) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.D) return;
  const uint32_t t = (uint32_t)(a.keys_sorted[j] >> 32);
  if (t >= a.num_tiles) return;   // padding of the speculative capacity
  if (j == 0 || (uint32_t)(a.keys_sorted[j - 1] >> 32) != t) a.ranges[t].x = (uint32_t)j;
  if (j == a.D - 1 || (uint32_t)(a.keys_sorted[j + 1] >> 32) != t) a.ranges[t].y = (uint32_t)(j + 1);
}

// K5 for 32-bit keys: per-bin ranges + exact order inside runs of equal keys.  A run holds pairs of one
// bin whose depths quantised alike; the stable sort left them in emission (= index) order, the exact
// order is (depth bits, index).  The head of a run repairs it: short runs by insertion on
// (depth_key[id], id); long runs (a fronto-parallel planar scene) are queued for k_fix_long_runs.
constexpr uint32_t TIE_INSERTION_MAX = 24;

__global__ void __launch_bounds__(256) k_tile_ranges32(Ranges32Args a) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.D) return;
  const uint32_t key = a.keys_sorted[j];
  const uint32_t t = key >> 24;
  if (t >= a.num_tiles) return;   // padding of the speculative capacity
  const bool has_prev = j > 0, has_next = j + 1 < a.D;
  const uint32_t prev = has_prev ? a.keys_sorted[j - 1] : 0u, next = has_next ? a.keys_sorted[j + 1] : 0u;
  if (!has_prev || (prev >> 24) != t) a.ranges[t].x = (uint32_t)j;
  if (!has_next || (next >> 24) != t) a.ranges[t].y = (uint32_t)(j + 1);
  if (!(has_next && next == key) || (has_prev && prev == key)) return;   // not the head of a run
  int64_t e = j + 2;
  while (e < a.D && a.keys_sorted[e] == key) e++;
  const uint32_t len = (uint32_t)(e - j);
  uint32_t* v = a.vals_sorted + j;
  if (len > TIE_INSERTION_MAX) {
    const uint32_t q = atomicAdd(a.run_count, 1u);
    if (q < a.run_capacity) { a.run_queue[q] = make_uint2((uint32_t)j, len); return; }
  }
  for (uint32_t x = 1; x < len; x++) {
    const uint32_t id = v[x];
    const uint64_t k = ((uint64_t)a.depth_key[id] << 32) | id;
    uint32_t y = x;
    while (y > 0) {
      const uint32_t pid = v[y - 1];
      if ((((uint64_t)a.depth_key[pid] << 32) | pid) <= k) break;
      v[y] = pid;
      y--;
    }
    v[y] = id;
  }
}

// One CTA per queued long run: sort it on the full (depth bits << 32 | index) key with the one-CTA radix
// passes of binsort.cuh (index digits first, then the depth word), scratch = the two key arrays the
// global sort no longer needs.
__global__ void __launch_bounds__(BS_THREADS) k_fix_long_runs(Ranges32Args a) {
  __shared__ BinSortShared sh;
  const uint32_t count = min(*a.run_count, a.run_capacity);
  for (uint32_t q = blockIdx.x; q < count; q += gridDim.x) {
    const uint2 run = a.run_queue[q];
    const uint32_t n = run.y;
    uint64_t* A = a.scratch_a + run.x;
    uint64_t* B = a.scratch_b + run.x;
    uint32_t* v = a.vals_sorted + run.x;
    for (uint32_t i = threadIdx.x; i < n; i += BS_THREADS) {
      const uint32_t id = v[i];
      A[i] = ((uint64_t)a.depth_key[id] << 32) | id;
    }
    __syncthreads();
    int passes = 0;
    for (int s = 0; s < a.id_bits; s += 8, passes++) bin_sort_pass(sh, (passes & 1) ? B : A, (passes & 1) ? A : B, n, s);
    for (int p = 0; p < 4; p++, passes++) bin_sort_pass(sh, (passes & 1) ? B : A, (passes & 1) ? A : B, n, 32 + 8 * p);
    const uint64_t* F = (passes & 1) ? B : A;
    for (uint32_t i = threadIdx.x; i < n; i += BS_THREADS) v[i] = (uint32_t)__ldcg(F + i);
    __syncthreads();
  }
}

// ==================================================================================================
// K10: near-plane visibility
// ==================================================================================================
__global__ void k_mark_visible(int P, const float* __restrict__ means, const float* __restrict__ view,
                               float near_plane, uint8_t* __restrict__ present) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float vz = __ldg(view + 2) * means[3 * i] + __ldg(view + 6) * means[3 * i + 1] +
                   __ldg(view + 10) * means[3 * i + 2] + __ldg(view + 14);
  present[i] = vz > near_plane ? 1 : 0;
}

// ==================================================================================================
// K8 + K9 fused: adjoint of the projection, SH colour and Sigma3D for one Gaussian per thread.
// Reads the screen-space accumulator grad2d[P][12] = {dcol r,g,b, S0, Sx, Sy, Sxx, Sxy, Syy, -, -, -}
// written by the compositing adjoint: moments of m = G * dL/dalpha about the splat centre, from which
//   dL/dopacity = S0,  dL/dmean2D = -o (A Sx + B Sy, C Sy + B Sx) * (0.5 W, 0.5 H)   (NDC-scaled),
//   dL/dA = -0.5 o Sxx,  dL/dB = -o Sxy,  dL/dC = -0.5 o Syy.
// Every output element is written.
// ==================================================================================================
template <int DEG>
__global__ void __launch_bounds__(PROJECT_BWD_THREADS, PROJECT_BWD_MIN_BLOCKS) k_project_bwd(ProjectBwdArgs a) {
  // SH-gradient rows of a block are contiguous in memory (256 rows x M*3 floats): with a.slab they are
  // staged in shared memory and leave the SM as ONE TMA bulk store (cp.async.bulk shared -> global)
  // instead of twelve 16-byte stores per thread at a 192-byte stride, whose half-written sectors cost
  // DRAM read-modify-write traffic (ncu: 203 MB read for ~140 MB of inputs).
  extern __shared__ float4 s_slab[];
  const bool use_slab = DEG >= 0 && a.slab != 0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // with the slab every thread of the block reaches the ONE barrier at the end (threads past P just
  // zero their unused slab row): a warp must never split over two different barrier instructions
  const bool in_range = i < a.P;
  if (!in_range && !use_slab) return;
  constexpr int NB = (DEG < 0 ? 0 : (DEG + 1) * (DEG + 1));
  float* const sh_row = use_slab ? reinterpret_cast<float*>(s_slab) + (size_t)threadIdx.x * a.M * 3
                                 : a.dL_dshs + (size_t)i * a.M * 3;

  float gm[3] = {0.f, 0.f, 0.f};
  float gcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float gs[3] = {0.f, 0.f, 0.f};
  float gq[4] = {0.f, 0.f, 0.f, 0.f};
  float g2x = 0.f, g2y = 0.f, gop = 0.f;
  float gcol[3] = {0.f, 0.f, 0.f};
  const bool active = in_range && a.radii[i] > 0 && a.tiles[i] != 0u;     // (tiles[] may hold a packed footprint)
  // active_only: every gradient tensor was zero-filled beforehand (on a side stream, under the compositing adjoint);
  // only the rows of Gaussians that reached the image are written here (the launcher disables the slab)
  if (a.active_only && !active) return;
  const bool vec_ok = a.sh_vec != 0;

  if (active) {
#ifdef PROJECT_BWD_PREFETCH   // measured on C3: 0.1064 -> 0.1092 ms (the adjoint's loads are not what it waits for) -- off
    if constexpr (DEG > 0) {
      const char* row = reinterpret_cast<const char*>(a.shs + (size_t)i * a.M * 3);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 128));
    }
#endif
#if PROJECT_BWD_EARLY_CLAMP
    const uint32_t cl = DEG >= 0 ? (uint32_t)a.clamped[i] : 0u;     // with the first batch of loads, not behind the geometry
#endif
#if PROJECT_BWD_STAGE_SH
    // The SH row is only needed behind the geometry part, and holding its 12 x 128-bit loads in registers across that
    // part does not fit: the compiler issues them late, a third dependent DRAM round trip.  With the slab the thread
    // owns a 192-byte shared-memory row (its OUTPUT row): the input row is copied into it by LDGSTS right away -- no
    // register holds it in flight --, read back behind the geometry and then overwritten with the gradient row.
    const bool staged = use_slab && vec_ok && DEG >= 0;
    if (staged) {
      constexpr int NV_IN = (NB * 3 + 3) / 4;
      const char* src = reinterpret_cast<const char*>(a.shs + (size_t)i * a.M * 3);
      const uint32_t dst = smem_u32(sh_row);
#pragma unroll
      for (int v = 0; v < NV_IN; v++)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * v), "l"(src + 16 * v) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
#endif
    CamConst c;
    load_cam(c, a.view, a.proj, a.campos);
    const float4* g4 = reinterpret_cast<const float4*>(a.grad2d + (size_t)i * GRAD2D_STRIDE);
    const float4 ga4 = g4[0], gb4 = g4[1], gc4 = g4[2];
    gcol[0] = ga4.x; gcol[1] = ga4.y; gcol[2] = ga4.z;
    gop = ga4.w;
    const float4 r0 = a.rec[(size_t)i * REC_F4], r1 = a.rec[(size_t)i * REC_F4 + 1];
    const float cA = r0.z, cB = r0.w, cC = r1.x, op = r1.y;
    const float Sx = gb4.x, Sy = gb4.y, Sxx = gb4.z, Sxy = gb4.w, Syy = gc4.x;
    g2x = -op * (cA * Sx + cB * Sy) * (0.5f * (float)a.W);
    g2y = -op * (cC * Sy + cB * Sx) * (0.5f * (float)a.H);
    const float gA = -0.5f * op * Sxx, gB = -op * Sxy, gC = -0.5f * op * Syy;

    const float3 mu = make_float3(__ldg(a.means + 3 * i), __ldg(a.means + 3 * i + 1), __ldg(a.means + 3 * i + 2));
    float c3[6];
    float3 s = make_float3(0.f, 0.f, 0.f);
    float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
    if (a.cov3d_precomp) {
#pragma unroll
      for (int k = 0; k < 6; k++) c3[k] = __ldg(a.cov3d_precomp + 6 * (size_t)i + k);
    } else {
      s = make_float3(__ldg(a.scales + 3 * i), __ldg(a.scales + 3 * i + 1), __ldg(a.scales + 3 * i + 2));
      q = __ldg(reinterpret_cast<const float4*>(a.rots) + i);
      cov3d_from_scale_rot(s, a.scale_modifier, q, c3);
    }
    // ---- conic -> Sigma2D -> Sigma3D, view-space mean (oracle A.2) ----
    Ewa e;
    ewa_jacobian(c, mu, (float)a.W, (float)a.H, a.tanfovx, a.tanfovy, e);
    float a_, b_, c_;
    cov2d(e, c3, a_, b_, c_);
    const float det = a_ * c_ - b_ * b_;
    const float d2inv = 1.f / (det * det + 0.0000001f);
    const float ga = d2inv * (-c_ * c_ * gA + b_ * c_ * gB + (det - a_ * c_) * gC);
    const float gc = d2inv * (-a_ * a_ * gC + a_ * b_ * gB + (det - a_ * c_) * gA);
    const float gb = d2inv * (2.f * b_ * c_ * gA - (det + 2.f * b_ * b_) * gB + 2.f * a_ * b_ * gC);
    const float G2[2][2] = {{ga, 0.5f * gb}, {0.5f * gb, gc}};
    float G2M[2][3];
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int k = 0; k < 3; k++) G2M[r][k] = G2[r][0] * e.m[0][k] + G2[r][1] * e.m[1][k];
    float GS[3][3];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int k = 0; k < 3; k++) GS[r][k] = e.m[0][r] * G2M[0][k] + e.m[1][r] * G2M[1][k];
    gcov[0] = GS[0][0]; gcov[3] = GS[1][1]; gcov[5] = GS[2][2];
    gcov[1] = 2.f * GS[0][1]; gcov[2] = 2.f * GS[0][2]; gcov[4] = 2.f * GS[1][2];
    const float S[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    float gM[2][3];
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int k = 0; k < 3; k++)
        gM[r][k] = 2.f * (G2M[r][0] * S[0][k] + G2M[r][1] * S[1][k] + G2M[r][2] * S[2][k]);
    float gJ[2][3];
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int j = 0; j < 3; j++) gJ[r][j] = gM[r][0] * c.v[j] + gM[r][1] * c.v[4 + j] + gM[r][2] * c.v[8 + j];
    const float tz = 1.f / e.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    const float gtx = e.xmask * (-e.fx * tz2 * gJ[0][2]);
    const float gty = e.ymask * (-e.fy * tz2 * gJ[1][2]);
    const float gtz = -e.fx * tz2 * gJ[0][0] - e.fy * tz2 * gJ[1][1] + (2.f * e.fx * e.t[0]) * tz3 * gJ[0][2] +
                      (2.f * e.fy * e.t[1]) * tz3 * gJ[1][2];
#pragma unroll
    for (int k = 0; k < 3; k++) gm[k] += c.v[4 * k] * gtx + c.v[4 * k + 1] * gty + c.v[4 * k + 2] * gtz;

    // ---- pixel centre through the full projection (oracle A.3) ----
    const float hx = c.p[0] * mu.x + c.p[4] * mu.y + c.p[8] * mu.z + c.p[12];
    const float hy = c.p[1] * mu.x + c.p[5] * mu.y + c.p[9] * mu.z + c.p[13];
    const float hw = c.p[3] * mu.x + c.p[7] * mu.y + c.p[11] * mu.z + c.p[15];
    const float w = 1.f / (hw + 0.0000001f);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float dnx = c.p[4 * k] * w - c.p[4 * k + 3] * hx * w * w;
      const float dny = c.p[4 * k + 1] * w - c.p[4 * k + 3] * hy * w * w;
      gm[k] += dnx * g2x + dny * g2y;
    }

    // ---- SH colour (oracle A.4) ----
    if constexpr (DEG >= 0) {
      constexpr int NF = NB * 3;
      float f[NF > 0 ? NF : 1];
#if PROJECT_BWD_STAGE_SH
      if (staged) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");      // the thread's own copies: its row is complete
        constexpr int NV_IN = (NF + 3) / 4;
        const float4* r4 = reinterpret_cast<const float4*>(sh_row);
#pragma unroll
        for (int v = 0; v < NV_IN; v++) {
          const float4 t4 = r4[v];
          if (4 * v + 0 < NF) f[4 * v + 0] = t4.x;
          if (4 * v + 1 < NF) f[4 * v + 1] = t4.y;
          if (4 * v + 2 < NF) f[4 * v + 2] = t4.z;
          if (4 * v + 3 < NF) f[4 * v + 3] = t4.w;
        }
      } else
#endif
      load_sh_row<NF>(a.shs + (size_t)i * a.M * 3, vec_ok, f);
      const float dx = mu.x - c.cam[0], dy = mu.y - c.cam[1], dz = mu.z - c.cam[2];
      const float inv = 1.f / sqrtf(dx * dx + dy * dy + dz * dz);
      const float x = dx * inv, y = dy * inv, z = dz * inv;
      float b[NB > 0 ? NB : 1];
      sh_basis<DEG>(x, y, z, b);
#if !PROJECT_BWD_EARLY_CLAMP
      const uint32_t cl = a.clamped[i];
#endif
      float gc3[3];
#pragma unroll
      for (int ch = 0; ch < 3; ch++) gc3[ch] = ((cl >> ch) & 1u) ? 0.f : gcol[ch];
      // t_k = sum_ch sh[k][ch] * dL/drgb[ch]
      float t[NB > 0 ? NB : 1];
#pragma unroll
      for (int k = 0; k < NB; k++) t[k] = f[3 * k] * gc3[0] + f[3 * k + 1] * gc3[1] + f[3 * k + 2] * gc3[2];
      float gdx = 0.f, gdy = 0.f, gdz = 0.f;
      if (DEG > 0) {
        gdy += -SH_C1 * t[1]; gdz += SH_C1 * t[2]; gdx += -SH_C1 * t[3];
      }
      if (DEG > 1) {
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        gdx += SH_C2[0] * y * t[4];                 gdy += SH_C2[0] * x * t[4];
        gdy += SH_C2[1] * z * t[5];                 gdz += SH_C2[1] * y * t[5];
        gdx += SH_C2[2] * -2.f * x * t[6];          gdy += SH_C2[2] * -2.f * y * t[6];
        gdz += SH_C2[2] * 4.f * z * t[6];
        gdx += SH_C2[3] * z * t[7];                 gdz += SH_C2[3] * x * t[7];
        gdx += SH_C2[4] * 2.f * x * t[8];           gdy += SH_C2[4] * -2.f * y * t[8];
        if (DEG > 2) {
          gdx += SH_C3[0] * 6.f * xy * t[9];        gdy += SH_C3[0] * (3.f * xx - 3.f * yy) * t[9];
          gdx += SH_C3[1] * yz * t[10];             gdy += SH_C3[1] * xz * t[10];
          gdz += SH_C3[1] * xy * t[10];
          gdx += SH_C3[2] * -2.f * xy * t[11];      gdy += SH_C3[2] * (4.f * zz - xx - 3.f * yy) * t[11];
          gdz += SH_C3[2] * 8.f * yz * t[11];
          gdx += SH_C3[3] * -6.f * xz * t[12];      gdy += SH_C3[3] * -6.f * yz * t[12];
          gdz += SH_C3[3] * (6.f * zz - 3.f * xx - 3.f * yy) * t[12];
          gdx += SH_C3[4] * (4.f * zz - 3.f * xx - yy) * t[13];
          gdy += SH_C3[4] * -2.f * xy * t[13];      gdz += SH_C3[4] * 8.f * xz * t[13];
          gdx += SH_C3[5] * 2.f * xz * t[14];       gdy += SH_C3[5] * -2.f * yz * t[14];
          gdz += SH_C3[5] * (xx - yy) * t[14];
          gdx += SH_C3[6] * (3.f * xx - 3.f * yy) * t[15];
          gdy += SH_C3[6] * -6.f * xy * t[15];
        }
      }
      const float dot = x * gdx + y * gdy + z * gdz;
      gm[0] += (gdx - x * dot) * inv;
      gm[1] += (gdy - y * dot) * inv;
      gm[2] += (gdz - z * dot) * inv;
      // dL/dsh[k][ch] = basis_k * dL/drgb[ch]; reuse f[] as the output row
#pragma unroll
      for (int k = 0; k < NB; k++) {
        f[3 * k] = b[k] * gc3[0]; f[3 * k + 1] = b[k] * gc3[1]; f[3 * k + 2] = b[k] * gc3[2];
      }
      float* row = sh_row;
      const int total = a.M * 3;
      if (vec_ok) {
        float4* r4 = reinterpret_cast<float4*>(row);
        constexpr int NV = (NF + 3) / 4;
#pragma unroll
        for (int v = 0; v < NV; v++) {
          float4 o;
          o.x = (4 * v + 0 < NF) ? f[(4 * v + 0 < NF) ? 4 * v + 0 : 0] : 0.f;
          o.y = (4 * v + 1 < NF) ? f[(4 * v + 1 < NF) ? 4 * v + 1 : 0] : 0.f;
          o.z = (4 * v + 2 < NF) ? f[(4 * v + 2 < NF) ? 4 * v + 2 : 0] : 0.f;
          o.w = (4 * v + 3 < NF) ? f[(4 * v + 3 < NF) ? 4 * v + 3 : 0] : 0.f;
          r4[v] = o;
        }
        for (int v = NV; v < total / 4; v++) r4[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
#pragma unroll
        for (int k = 0; k < NF; k++) row[k] = f[k];
        for (int k = NF; k < total; k++) row[k] = 0.f;
      }
    }

    // ---- Sigma3D -> scale, quaternion (oracle A.5) ----
    if (!a.cov3d_precomp) {
      const float r = q.x, qx = q.y, qy = q.z, qz = q.w;
      const float R[3][3] = {
          {1.f - 2.f * (qy * qy + qz * qz), 2.f * (qx * qy - r * qz), 2.f * (qx * qz + r * qy)},
          {2.f * (qx * qy + r * qz), 1.f - 2.f * (qx * qx + qz * qz), 2.f * (qy * qz - r * qx)},
          {2.f * (qx * qz - r * qy), 2.f * (qy * qz + r * qx), 1.f - 2.f * (qx * qx + qy * qy)}};
      const float mod = a.scale_modifier;
      const float d[3] = {mod * s.x, mod * s.y, mod * s.z};
      const float Gf[3][3] = {{gcov[0], 0.5f * gcov[1], 0.5f * gcov[2]},
                              {0.5f * gcov[1], gcov[3], 0.5f * gcov[4]},
                              {0.5f * gcov[2], 0.5f * gcov[4], gcov[5]}};
      float F[3][3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        float acc = 0.f;
#pragma unroll
        for (int r_ = 0; r_ < 3; r_++) {
          const float gmx = 2.f * (Gf[r_][0] * R[0][k] + Gf[r_][1] * R[1][k] + Gf[r_][2] * R[2][k]) * d[k];
          acc += gmx * R[r_][k];
          F[r_][k] = gmx * d[k];
        }
        gs[k] = mod * acc;
      }
      gq[0] = 2.f * (-qz * F[0][1] + qy * F[0][2] + qz * F[1][0] - qx * F[1][2] - qy * F[2][0] + qx * F[2][1]);
      gq[1] = 2.f * (qy * F[0][1] + qz * F[0][2] + qy * F[1][0] - 2.f * qx * F[1][1] - r * F[1][2] +
                     qz * F[2][0] + r * F[2][1] - 2.f * qx * F[2][2]);
      gq[2] = 2.f * (-2.f * qy * F[0][0] + qx * F[0][1] + r * F[0][2] + qx * F[1][0] + qz * F[1][2] -
                     r * F[2][0] + qz * F[2][1] - 2.f * qy * F[2][2]);
      gq[3] = 2.f * (-2.f * qz * F[0][0] - r * F[0][1] + qx * F[0][2] + r * F[1][0] - 2.f * qz * F[1][1] +
                     qy * F[1][2] + qx * F[2][0] + qy * F[2][1]);
    }
  } else if constexpr (DEG >= 0) {
    // culled Gaussian: zero SH gradient row
    float* row = sh_row;
    const int total = a.M * 3;
    if (vec_ok) {
      float4* r4 = reinterpret_cast<float4*>(row);
      for (int v = 0; v < total / 4; v++) r4[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int k = 0; k < total; k++) row[k] = 0.f;
    }
  }

  if (in_range) {
  a.dL_dmeans[3 * (size_t)i] = gm[0];
  a.dL_dmeans[3 * (size_t)i + 1] = gm[1];
  a.dL_dmeans[3 * (size_t)i + 2] = gm[2];
  a.dL_dmeans2D[3 * (size_t)i] = g2x;
  a.dL_dmeans2D[3 * (size_t)i + 1] = g2y;
  a.dL_dmeans2D[3 * (size_t)i + 2] = 0.f;
  a.dL_dopac[i] = gop;
  if (DEG < 0 && a.dL_dcolors) {
    a.dL_dcolors[3 * (size_t)i] = gcol[0];
    a.dL_dcolors[3 * (size_t)i + 1] = gcol[1];
    a.dL_dcolors[3 * (size_t)i + 2] = gcol[2];
  }
  if (a.cov3d_precomp) {
    if (a.dL_dcov3D) {
#pragma unroll
      for (int k = 0; k < 6; k++) a.dL_dcov3D[6 * (size_t)i + k] = gcov[k];
    }
  } else {
    a.dL_dscales[3 * (size_t)i] = gs[0];
    a.dL_dscales[3 * (size_t)i + 1] = gs[1];
    a.dL_dscales[3 * (size_t)i + 2] = gs[2];
    reinterpret_cast<float4*>(a.dL_drots)[i] = make_float4(gq[0], gq[1], gq[2], gq[3]);
  }
  }   // in_range
  if (use_slab) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
      const int rows = min((int)blockDim.x, a.P - (int)(blockIdx.x * blockDim.x));
      const uint32_t bytes = (uint32_t)rows * (uint32_t)a.M * 12u;
      float* dst = a.dL_dshs + (size_t)blockIdx.x * blockDim.x * a.M * 3;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(s_slab)),
                   "r"(bytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must outlive the read
    }
  }
}

// ---- host launchers ------------------------------------------------------------------------------
static std::atomic<int> g_project_mode{-1};
void set_project_mode(int mode) { g_project_mode.store(mode == 1 ? 1 : 0); }
static bool use_project_compact() {
  int m = g_project_mode.load();
  if (m < 0) {
    const char* e = getenv("B200GS_PROJECT");
    m = (e && e[0] == 'c') ? 1 : 0;      // B200GS_PROJECT=compact selects the dense-warp kernel (measured: no gain)
    g_project_mode.store(m);
  }
  return m == 1;
}

void launch_project(const ProjectArgs& a, int deg, cudaStream_t st) {
  if (a.P == 0) return;
  static std::atomic<int> sm_count{0};
  int sms = sm_count.load();
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    sm_count.store(sms);
  }
  if (use_project_compact() && !a.cov3d_precomp) {     // (precomputed covariances: k_project)
    const int need = (a.P + PD_CHUNK * PD_WARPS - 1) / (PD_CHUNK * PD_WARPS);
    static std::atomic<int> dwaves{-1};
    int nw = dwaves.load();
    if (nw < 0) {
      const char* e = getenv("B200GS_PROJECT_DENSE_WAVES");
      nw = e ? atoi(e) : PROJECT_DENSE_WAVES;
      dwaves.store(nw);
    }
    const dim3 grid(nw > 0 ? std::min(need, sms * PROJECT_DENSE_MIN_BLOCKS * nw) : need), block(32 * PD_WARPS);
    switch (deg) {
      case -1: k_project_dense<-1><<<grid, block, 0, st>>>(a); break;
      case 0: k_project_dense<0><<<grid, block, 0, st>>>(a); break;
      case 1: k_project_dense<1><<<grid, block, 0, st>>>(a); break;
      case 2: k_project_dense<2><<<grid, block, 0, st>>>(a); break;
      default: k_project_dense<3><<<grid, block, 0, st>>>(a); break;
    }
    count_launch();
    return;
  }
  // grid-stride kernel: PROJECT_MIN_BLOCKS resident CTAs per SM times a few "waves", so that every thread works on
  // several Gaussians with the loads of the next one in flight (load_proj_in) and the tail stays short
  static std::atomic<int> ctas_per_wave{0}, waves{0};
  int cpw = ctas_per_wave.load(), nw = waves.load();
  if (cpw == 0) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cpw = (sms > 0 ? sms : 148) * PROJECT_MIN_BLOCKS;
    const char* e = getenv("B200GS_PROJECT_WAVES");     // 0: one Gaussian per thread (no look-ahead)
    nw = e ? atoi(e) : PROJECT_WAVES;
    ctas_per_wave.store(cpw);
    waves.store(nw);
  }
  const int need = (a.P + 255) / 256;
  const dim3 grid(nw > 0 ? std::min(need, cpw * nw) : need), block(256);
  switch (deg) {
    case -1: k_project<-1><<<grid, block, 0, st>>>(a); break;
    case 0: k_project<0><<<grid, block, 0, st>>>(a); break;
    case 1: k_project<1><<<grid, block, 0, st>>>(a); break;
    case 2: k_project<2><<<grid, block, 0, st>>>(a); break;
    default: k_project<3><<<grid, block, 0, st>>>(a); break;
  }
  count_launch();
}

void launch_emit_pairs(const EmitArgs& a, cudaStream_t st) {
  if (a.P == 0) return;
  k_emit_pairs<<<(a.P + 255) / 256, 256, 0, st>>>(a);
  k_emit_big<<<148 * 4, 256, 0, st>>>(a);
  count_launch(2);
}

void launch_tile_ranges(const RangesArgs& a, cudaStream_t st) {
  if (a.D == 0) return;
  k_tile_ranges<<<(unsigned)((a.D + 255) / 256), 256, 0, st>>>(a);
  count_launch();
}

void launch_tile_ranges32(const Ranges32Args& a, cudaStream_t st) {
  if (a.D == 0) return;
  k_tile_ranges32<<<(unsigned)((a.D + 255) / 256), 256, 0, st>>>(a);
  k_fix_long_runs<<<148, BS_THREADS, 0, st>>>(a);
  count_launch(2);
}

void launch_mark_visible(int P, const float* means, const float* view, float near_plane, uint8_t* present,
                         cudaStream_t st) {
  if (P == 0) return;
  k_mark_visible<<<(P + 255) / 256, 256, 0, st>>>(P, means, view, near_plane, present);
  count_launch();
}

__global__ void k_extract_alpha(const float4* __restrict__ pix, size_t npx, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npx) out[i] = 1.f - pix[i].w;
}

void launch_extract_alpha(const float4* pix, size_t npx, float* out, cudaStream_t st) {
  if (npx == 0) return;
  k_extract_alpha<<<(unsigned)((npx + 255) / 256), 256, 0, st>>>(pix, npx, out);
  count_launch();
}

void launch_project_bwd(const ProjectBwdArgs& a_in, int deg, cudaStream_t st) {
  if (a_in.P == 0) return;
  ProjectBwdArgs a = a_in;
  const dim3 grid((a.P + PROJECT_BWD_THREADS - 1) / PROJECT_BWD_THREADS), block(PROJECT_BWD_THREADS);
  // SH-gradient slab through shared memory + TMA bulk store: rows must be 16-byte multiples and the slab
  // must fit the default 48 KB of dynamic shared memory (M <= 16)
  const size_t slab_bytes = (size_t)PROJECT_BWD_THREADS * a.M * 12;
  a.slab = (deg >= 0 && a.sh_vec && slab_bytes <= 48 * 1024 && a.slab >= 0 && !a.active_only) ? 1 : 0;
  const size_t sm = a.slab ? slab_bytes : 0;
  switch (deg) {
    case -1: k_project_bwd<-1><<<grid, block, 0, st>>>(a); break;
    case 0: k_project_bwd<0><<<grid, block, sm, st>>>(a); break;
    case 1: k_project_bwd<1><<<grid, block, sm, st>>>(a); break;
    case 2: k_project_bwd<2><<<grid, block, sm, st>>>(a); break;
    default: k_project_bwd<3><<<grid, block, sm, st>>>(a); break;
  }
  count_launch();
}

}  // namespace b200gs

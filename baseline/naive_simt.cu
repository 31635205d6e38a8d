// naive_simt.cu -- YARDSTICK ONLY, never product code (SURVEY.md 8(d) last row).
//
// A straight SIMT restatement of the PUBLISHED 3DGS rasterizer design (Kerbl et al. 2023, Sec. 6 /
// App. C), written from the algorithm description, with none of this
// repo's B200 work: one thread per Gaussian with scalar strided loads, every 16x16 tile of the 3-sigma
// square gets a pair, host read-back of D, library radix sort of 64-bit (tile | depth) keys over
// 32 + ceil(log2 T) bits, a ranges pass, one thread per pixel with 256-record cooperative fetches, and
// a back-to-front adjoint with one atomicAdd per pixel per gradient component.  It exists so the
// "1.5x the reference rasterizer" target of BASELINE.json has a measured comparator on the same GPU
// when the real third-party library cannot be installed (no network).  tools/compare_naive.py checks
// that it renders the same image as libb200gs before timing it.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

namespace naive {

constexpr int TILE = 16;
__device__ const float SH_C0 = 0.28209479177387814f;
__device__ const float SH_C1 = 0.4886025119029199f;
__device__ const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                   -1.0925484305920792f, 0.5462742152960396f};
__device__ const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                   -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

struct Cam {
  int P, deg, M, H, W;
  float tanfovx, tanfovy;
  const float *view, *proj, *campos, *bg;
};

__global__ void k_preprocess(Cam c, const float* means, const float* scales, const float* rots, const float* opac,
                             const float* shs, int* radii, float2* xy, float* depths, float4* conic_opacity, float* rgb,
                             uint32_t* tiles_touched) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.P) return;
  radii[i] = 0;
  tiles_touched[i] = 0;
  const float x = means[3 * i], y = means[3 * i + 1], z = means[3 * i + 2];
  const float* V = c.view; const float* Pm = c.proj;
  const float tz = V[2] * x + V[6] * y + V[10] * z + V[14];
  if (tz <= 0.2f) return;
  const float hx = Pm[0] * x + Pm[4] * y + Pm[8] * z + Pm[12];
  const float hy = Pm[1] * x + Pm[5] * y + Pm[9] * z + Pm[13];
  const float hw = Pm[3] * x + Pm[7] * y + Pm[11] * z + Pm[15];
  const float pw = 1.f / (hw + 0.0000001f);
  // Sigma3D = R S S^T R^T
  const float qr = rots[4 * i], qx = rots[4 * i + 1], qy = rots[4 * i + 2], qz = rots[4 * i + 3];
  const float R[3][3] = {{1.f - 2.f * (qy * qy + qz * qz), 2.f * (qx * qy - qr * qz), 2.f * (qx * qz + qr * qy)},
                         {2.f * (qx * qy + qr * qz), 1.f - 2.f * (qx * qx + qz * qz), 2.f * (qy * qz - qr * qx)},
                         {2.f * (qx * qz - qr * qy), 2.f * (qy * qz + qr * qx), 1.f - 2.f * (qx * qx + qy * qy)}};
  const float s[3] = {scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]};
  float Mx[3][3];
  for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) Mx[r][k] = R[r][k] * s[k];
  float S3[3][3];
  for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) S3[r][k] = Mx[r][0] * Mx[k][0] + Mx[r][1] * Mx[k][1] + Mx[r][2] * Mx[k][2];
  // EWA
  float tx = V[0] * x + V[4] * y + V[8] * z + V[12];
  float ty = V[1] * x + V[5] * y + V[9] * z + V[13];
  const float limx = 1.3f * c.tanfovx, limy = 1.3f * c.tanfovy;
  tx = fminf(limx, fmaxf(-limx, tx / tz)) * tz;
  ty = fminf(limy, fmaxf(-limy, ty / tz)) * tz;
  const float fx = c.W / (2.f * c.tanfovx), fy = c.H / (2.f * c.tanfovy);
  const float J[2][3] = {{fx / tz, 0.f, -(fx * tx) / (tz * tz)}, {0.f, fy / tz, -(fy * ty) / (tz * tz)}};
  float T[2][3];
  for (int r = 0; r < 2; r++) for (int k = 0; k < 3; k++) T[r][k] = J[r][0] * V[4 * k] + J[r][1] * V[4 * k + 1] + J[r][2] * V[4 * k + 2];
  float TS[2][3];
  for (int r = 0; r < 2; r++) for (int k = 0; k < 3; k++) TS[r][k] = T[r][0] * S3[0][k] + T[r][1] * S3[1][k] + T[r][2] * S3[2][k];
  const float ca = TS[0][0] * T[0][0] + TS[0][1] * T[0][1] + TS[0][2] * T[0][2] + 0.3f;
  const float cb = TS[0][0] * T[1][0] + TS[0][1] * T[1][1] + TS[0][2] * T[1][2];
  const float cc = TS[1][0] * T[1][0] + TS[1][1] * T[1][1] + TS[1][2] * T[1][2] + 0.3f;
  const float det = ca * cc - cb * cb;
  if (det == 0.f) return;
  const float di = 1.f / det;
  const float mid = 0.5f * (ca + cc);
  const float l1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det)), l2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
  const int rad = (int)ceilf(3.f * sqrtf(fmaxf(l1, l2)));
  const float px = ((hx * pw + 1.f) * c.W - 1.f) * 0.5f, py = ((hy * pw + 1.f) * c.H - 1.f) * 0.5f;
  const int gx = (c.W + TILE - 1) / TILE, gy = (c.H + TILE - 1) / TILE;
  const int x0 = min(gx, max(0, (int)((px - rad) / TILE))), y0 = min(gy, max(0, (int)((py - rad) / TILE)));
  const int x1 = min(gx, max(0, (int)((px + rad + TILE - 1) / TILE))), y1 = min(gy, max(0, (int)((py + rad + TILE - 1) / TILE)));
  if ((x1 - x0) * (y1 - y0) == 0) return;
  // SH colour, scalar strided loads
  float dx = x - c.campos[0], dy = y - c.campos[1], dz = z - c.campos[2];
  const float inv = 1.f / sqrtf(dx * dx + dy * dy + dz * dz);
  dx *= inv; dy *= inv; dz *= inv;
  const float* sh = shs + (size_t)i * c.M * 3;
  for (int ch = 0; ch < 3; ch++) {
    float r = SH_C0 * sh[ch];
    if (c.deg > 0) {
      r += -SH_C1 * dy * sh[3 + ch] + SH_C1 * dz * sh[6 + ch] - SH_C1 * dx * sh[9 + ch];
      if (c.deg > 1) {
        const float xx = dx * dx, yy = dy * dy, zz = dz * dz, xyv = dx * dy, yz = dy * dz, xz = dx * dz;
        r += SH_C2[0] * xyv * sh[12 + ch] + SH_C2[1] * yz * sh[15 + ch] + SH_C2[2] * (2.f * zz - xx - yy) * sh[18 + ch] +
             SH_C2[3] * xz * sh[21 + ch] + SH_C2[4] * (xx - yy) * sh[24 + ch];
        if (c.deg > 2) {
          r += SH_C3[0] * dy * (3.f * xx - yy) * sh[27 + ch] + SH_C3[1] * xyv * dz * sh[30 + ch] +
               SH_C3[2] * dy * (4.f * zz - xx - yy) * sh[33 + ch] + SH_C3[3] * dz * (2.f * zz - 3.f * xx - 3.f * yy) * sh[36 + ch] +
               SH_C3[4] * dx * (4.f * zz - xx - yy) * sh[39 + ch] + SH_C3[5] * dz * (xx - yy) * sh[42 + ch] +
               SH_C3[6] * dx * (xx - 3.f * yy) * sh[45 + ch];
        }
      }
    }
    rgb[3 * i + ch] = fmaxf(r + 0.5f, 0.f);
  }
  depths[i] = tz;
  radii[i] = rad;
  xy[i] = make_float2(px, py);
  conic_opacity[i] = make_float4(cc * di, -cb * di, ca * di, opac[i]);
  tiles_touched[i] = (uint32_t)((x1 - x0) * (y1 - y0));
}

__global__ void k_duplicate(int P, int W, int H, const float2* xy, const float* depths, const uint32_t* offsets,
                            const int* radii, uint64_t* keys, uint32_t* vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P || radii[i] <= 0) return;
  uint32_t off = i == 0 ? 0u : offsets[i - 1];
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const float px = xy[i].x, py = xy[i].y;
  const int rad = radii[i];
  const int x0 = min(gx, max(0, (int)((px - rad) / TILE))), y0 = min(gy, max(0, (int)((py - rad) / TILE)));
  const int x1 = min(gx, max(0, (int)((px + rad + TILE - 1) / TILE))), y1 = min(gy, max(0, (int)((py + rad + TILE - 1) / TILE)));
  for (int y = y0; y < y1; y++)
    for (int x = x0; x < x1; x++) {
      keys[off] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | __float_as_uint(depths[i]);
      vals[off] = (uint32_t)i;
      off++;
    }
}

__global__ void k_ranges(int64_t D, const uint64_t* keys, uint2* ranges) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= D) return;
  const uint32_t t = (uint32_t)(keys[j] >> 32);
  if (j == 0) ranges[t].x = 0;
  else {
    const uint32_t p = (uint32_t)(keys[j - 1] >> 32);
    if (p != t) { ranges[p].y = (uint32_t)j; ranges[t].x = (uint32_t)j; }
  }
  if (j == D - 1) ranges[t].y = (uint32_t)D;
}

__global__ void __launch_bounds__(TILE * TILE) k_render(int W, int H, const uint2* ranges, const uint32_t* list,
                                                        const float2* xy, const float* rgb, const float4* conic_opacity,
                                                        const float* bg, float* final_T, uint32_t* n_contrib, float* out) {
  __shared__ uint32_t s_id[TILE * TILE];
  __shared__ float2 s_xy[TILE * TILE];
  __shared__ float4 s_co[TILE * TILE];
  const int gx = (W + TILE - 1) / TILE;
  const int px = blockIdx.x * TILE + threadIdx.x, py = blockIdx.y * TILE + threadIdx.y;
  const int tid = threadIdx.y * TILE + threadIdx.x;
  const bool inside = px < W && py < H;
  const uint2 range = ranges[blockIdx.y * gx + blockIdx.x];
  const int rounds = (int)((range.y - range.x + TILE * TILE - 1) / (TILE * TILE));
  int todo = (int)(range.y - range.x);
  bool done = !inside;
  float T = 1.f, C[3] = {0.f, 0.f, 0.f};
  uint32_t contributor = 0, last = 0;
  for (int r = 0; r < rounds; r++, todo -= TILE * TILE) {
    if (__syncthreads_count(done) == TILE * TILE) break;
    const int prog = r * TILE * TILE + tid;
    if (range.x + prog < range.y) {
      const uint32_t id = list[range.x + prog];
      s_id[tid] = id; s_xy[tid] = xy[id]; s_co[tid] = conic_opacity[id];
    }
    __syncthreads();
    for (int j = 0; !done && j < min(TILE * TILE, todo); j++) {
      contributor++;
      const float dx = s_xy[j].x - (float)px, dy = s_xy[j].y - (float)py;
      const float4 co = s_co[j];
      const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
      if (power > 0.f) continue;
      const float alpha = fminf(0.99f, co.w * __expf(power));
      if (alpha < 1.f / 255.f) continue;
      const float test_T = T * (1.f - alpha);
      if (test_T < 0.0001f) { done = true; continue; }
      const uint32_t id = s_id[j];
      for (int ch = 0; ch < 3; ch++) C[ch] += rgb[3 * id + ch] * alpha * T;
      T = test_T;
      last = contributor;
    }
  }
  if (inside) {
    const size_t pid = (size_t)py * W + px, hw = (size_t)H * W;
    final_T[pid] = T;
    n_contrib[pid] = last;
    for (int ch = 0; ch < 3; ch++) out[ch * hw + pid] = C[ch] + T * bg[ch];
  }
}

// back-to-front adjoint, one atomicAdd per pixel per component into acc[P][12] =
// {dcol rgb, S0, Sx, Sy, Sxx, Sxy, Syy} (the accumulator layout k_project_bwd of libb200gs consumes;
// nine atomics per contribution, the same count as the published design's mean2D/conic/opacity/colour)
__global__ void __launch_bounds__(TILE * TILE) k_render_bwd(int W, int H, const uint2* ranges, const uint32_t* list,
                                                            const float2* xy, const float* rgb, const float4* conic_opacity,
                                                            const float* bg, const float* final_T, const uint32_t* n_contrib,
                                                            const float* dL_dpix, float* acc) {
  __shared__ uint32_t s_id[TILE * TILE];
  __shared__ float2 s_xy[TILE * TILE];
  __shared__ float4 s_co[TILE * TILE];
  __shared__ float s_rgb[3][TILE * TILE];
  const int gx = (W + TILE - 1) / TILE;
  const int px = blockIdx.x * TILE + threadIdx.x, py = blockIdx.y * TILE + threadIdx.y;
  const int tid = threadIdx.y * TILE + threadIdx.x;
  const bool inside = px < W && py < H;
  const size_t pid = (size_t)py * W + px, hw = (size_t)H * W;
  const uint2 range = ranges[blockIdx.y * gx + blockIdx.x];
  const int rounds = (int)((range.y - range.x + TILE * TILE - 1) / (TILE * TILE));
  int todo = (int)(range.y - range.x);
  const float T_final = inside ? final_T[pid] : 0.f;
  float T = T_final;
  uint32_t contributor = (uint32_t)todo;
  const uint32_t last = inside ? n_contrib[pid] : 0u;
  float accum[3] = {0.f, 0.f, 0.f}, g[3] = {0.f, 0.f, 0.f}, last_alpha = 0.f, last_c[3] = {0.f, 0.f, 0.f};
  if (inside) for (int ch = 0; ch < 3; ch++) g[ch] = dL_dpix[ch * hw + pid];
  for (int r = 0; r < rounds; r++, todo -= TILE * TILE) {
    __syncthreads();
    const int prog = r * TILE * TILE + tid;
    if (range.x + prog < range.y) {
      const uint32_t id = list[range.y - prog - 1];
      s_id[tid] = id; s_xy[tid] = xy[id]; s_co[tid] = conic_opacity[id];
      for (int ch = 0; ch < 3; ch++) s_rgb[ch][tid] = rgb[3 * id + ch];
    }
    __syncthreads();
    for (int j = 0; inside && j < min(TILE * TILE, todo); j++) {
      contributor--;
      if (contributor >= last) continue;
      const float dx = s_xy[j].x - (float)px, dy = s_xy[j].y - (float)py;
      const float4 co = s_co[j];
      const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
      if (power > 0.f) continue;
      const float G = __expf(power);
      const float alpha = fminf(0.99f, co.w * G);
      if (alpha < 1.f / 255.f) continue;
      T = T / (1.f - alpha);
      const float w = alpha * T;
      float dL_dalpha = 0.f;
      float* row = acc + (size_t)s_id[j] * 12;
      for (int ch = 0; ch < 3; ch++) {
        const float cch = s_rgb[ch][j];
        accum[ch] = last_alpha * last_c[ch] + (1.f - last_alpha) * accum[ch];
        last_c[ch] = cch;
        dL_dalpha += (cch - accum[ch]) * g[ch];
        atomicAdd(row + ch, w * g[ch]);
      }
      dL_dalpha *= T;
      last_alpha = alpha;
      float bgdot = 0.f;
      for (int ch = 0; ch < 3; ch++) bgdot += bg[ch] * g[ch];
      dL_dalpha += (-T_final / (1.f - alpha)) * bgdot;
      const float m = G * dL_dalpha;
      atomicAdd(row + 3, m);
      atomicAdd(row + 4, m * dx);
      atomicAdd(row + 5, m * dy);
      atomicAdd(row + 6, m * dx * dx);
      atomicAdd(row + 7, m * dx * dy);
      atomicAdd(row + 8, m * dy * dy);
    }
  }
}

}  // namespace naive

using namespace naive;

extern "C" {

// phase 1: preprocess + scan; returns D after a host synchronisation (as the published design does)
int64_t naive_preprocess(int P, int deg, int M, int H, int W, float tanfovx, float tanfovy, const float* view,
                         const float* proj, const float* campos, const float* bg, const float* means,
                         const float* scales, const float* rots, const float* opac, const float* shs, int* radii,
                         float* xy, float* depths, float* conic_opacity, float* rgb, uint32_t* tiles_touched,
                         uint32_t* offsets, void* temp, size_t temp_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  Cam c{P, deg, M, H, W, tanfovx, tanfovy, view, proj, campos, bg};
  k_preprocess<<<(P + 255) / 256, 256, 0, st>>>(c, means, scales, rots, opac, shs, radii, (float2*)xy, depths,
                                                (float4*)conic_opacity, rgb, tiles_touched);
  cub::DeviceScan::InclusiveSum(temp, temp_bytes, tiles_touched, offsets, P, st);
  uint32_t D = 0;
  cudaMemcpyAsync(&D, offsets + P - 1, 4, cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
  return (int64_t)D;
}

size_t naive_temp_bytes(int P, int64_t D) {
  size_t a = 0, b = 0;
  cub::DeviceScan::InclusiveSum(nullptr, a, (uint32_t*)nullptr, (uint32_t*)nullptr, P);
  cub::DeviceRadixSort::SortPairs(nullptr, b, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t*)nullptr, D > 0 ? D : 1);
  return a > b ? a : b;
}

// phase 2: duplicate, sort, ranges, render
int naive_bin_and_render(int P, int H, int W, int64_t D, const float* bg, const int* radii, const float* xy,
                         const float* depths, const float* conic_opacity, const float* rgb, const uint32_t* offsets,
                         uint64_t* keys, uint64_t* keys_sorted, uint32_t* vals, uint32_t* vals_sorted, uint2* ranges,
                         void* temp, size_t temp_bytes, float* final_T, uint32_t* n_contrib, float* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)gx * gy, st);
  if (D > 0) {
    k_duplicate<<<(P + 255) / 256, 256, 0, st>>>(P, W, H, (const float2*)xy, depths, offsets, radii, keys, vals);
    int bits = 1;
    while ((1 << bits) < gx * gy) bits++;
    cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_sorted, vals, vals_sorted, D, 0, 32 + bits, st);
    k_ranges<<<(unsigned)((D + 255) / 256), 256, 0, st>>>(D, keys_sorted, ranges);
  }
  k_render<<<dim3(gx, gy), dim3(TILE, TILE), 0, st>>>(W, H, ranges, vals_sorted, (const float2*)xy, rgb,
                                                      (const float4*)conic_opacity, bg, final_T, n_contrib, out);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int naive_render_backward(int H, int W, const float* bg, const float* xy, const float* conic_opacity, const float* rgb,
                          const uint32_t* vals_sorted, const uint2* ranges, const float* final_T,
                          const uint32_t* n_contrib, const float* dL_dpix, float* acc, int P, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  cudaMemsetAsync(acc, 0, sizeof(float) * 12 * (size_t)P, st);
  k_render_bwd<<<dim3(gx, gy), dim3(TILE, TILE), 0, st>>>(W, H, ranges, vals_sorted, (const float2*)xy, rgb,
                                                          (const float4*)conic_opacity, bg, final_T, n_contrib,
                                                          dL_dpix, acc);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"

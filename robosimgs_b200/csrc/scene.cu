// scene.cu -- the two data-format kernels either side of the rasterizer:
//  * k_ply_activate: 3DGS .ply vertex records (pre-activation, AoS) -> the packed post-activation
//    tensors GaussianRasterizer.forward takes (sigmoid / exp / normalise + SH repack), one pass;
//  * k_transform_gaussians: per-link rigid transform of object Gaussians (means R*mu + t, quaternions
//    q_link (x) q) for the per-frame articulated composite (BASELINE config 5).
// Anchors: /root/reference/README.md:75 (".ply" hand-off of the reconstructed background),
// /root/reference/Articulation/urdf_generation/pipeline.py:290-357 (links / hinge that drive the pose).
#include "common.cuh"
#include "kernels.cuh"

namespace b200gs {

constexpr int PLY_BLOCK = 128;

// Block-cooperative: the block's PLY_BLOCK vertex records are one contiguous byte range, staged in
// shared memory with coalesced 32-bit loads; each thread then unpacks its own record.
__global__ void __launch_bounds__(PLY_BLOCK) k_ply_activate(int P, const float* __restrict__ v, B200GSPlyLayout L,
                                                             float* __restrict__ means, float* __restrict__ shs,
                                                             float* __restrict__ opac, float* __restrict__ scales,
                                                             float* __restrict__ rots) {
  extern __shared__ float sm[];
  const int first = blockIdx.x * PLY_BLOCK;
  const int nv = min(PLY_BLOCK, P - first);
  const size_t base = (size_t)first * L.stride;
  for (int k = threadIdx.x; k < nv * L.stride; k += PLY_BLOCK) sm[k] = __ldg(v + base + k);
  __syncthreads();
  if (threadIdx.x >= nv) return;
  const int i = first + threadIdx.x;
  const float* r = sm + threadIdx.x * L.stride;
  means[3 * (size_t)i] = r[L.off_xyz];
  means[3 * (size_t)i + 1] = r[L.off_xyz + 1];
  means[3 * (size_t)i + 2] = r[L.off_xyz + 2];
  const int M = 1 + L.n_rest;
  float* sh = shs + (size_t)i * M * 3;
  for (int c = 0; c < 3; c++) sh[c] = r[L.off_fdc + c];
  // f_rest is stored channel-major [3][n_rest]; the rasterizer wants [n_rest][3]
  for (int k = 0; k < L.n_rest; k++)
    for (int c = 0; c < 3; c++) sh[3 * (1 + k) + c] = r[L.off_frest + c * L.n_rest + k];
  opac[i] = 1.f / (1.f + expf(-r[L.off_opacity]));
  for (int c = 0; c < 3; c++) scales[3 * (size_t)i + c] = expf(r[L.off_scale + c]);
  const float q0 = r[L.off_rot], q1 = r[L.off_rot + 1], q2 = r[L.off_rot + 2], q3 = r[L.off_rot + 3];
  const float inv = 1.f / fmaxf(sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3), 1e-12f);
  rots[4 * (size_t)i] = q0 * inv;
  rots[4 * (size_t)i + 1] = q1 * inv;
  rots[4 * (size_t)i + 2] = q2 * inv;
  rots[4 * (size_t)i + 3] = q3 * inv;
}

__global__ void __launch_bounds__(256) k_transform_gaussians(int n, const float* __restrict__ means_in,
                                                            const float* __restrict__ rots_in,
                                                            const int32_t* __restrict__ link_ids,
                                                            const float* __restrict__ T, const float* __restrict__ Q,
                                                            int L, float* __restrict__ means_out,
                                                            float* __restrict__ rots_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int l = link_ids ? link_ids[i] : 0;
  l = min(max(l, 0), L - 1);
  const float* t = T + 12 * l;
  const float x = means_in[3 * (size_t)i], y = means_in[3 * (size_t)i + 1], z = means_in[3 * (size_t)i + 2];
  means_out[3 * (size_t)i] = t[0] * x + t[1] * y + t[2] * z + t[3];
  means_out[3 * (size_t)i + 1] = t[4] * x + t[5] * y + t[6] * z + t[7];
  means_out[3 * (size_t)i + 2] = t[8] * x + t[9] * y + t[10] * z + t[11];
  // Hamilton product q_link (x) q, both (w,x,y,z)
  const float aw = Q[4 * l], ax = Q[4 * l + 1], ay = Q[4 * l + 2], az = Q[4 * l + 3];
  const float bw = rots_in[4 * (size_t)i], bx = rots_in[4 * (size_t)i + 1], by = rots_in[4 * (size_t)i + 2],
              bz = rots_in[4 * (size_t)i + 3];
  rots_out[4 * (size_t)i] = aw * bw - ax * bx - ay * by - az * bz;
  rots_out[4 * (size_t)i + 1] = aw * bx + ax * bw + ay * bz - az * by;
  rots_out[4 * (size_t)i + 2] = aw * by - ax * bz + ay * bw + az * bx;
  rots_out[4 * (size_t)i + 3] = aw * bz + ax * by - ay * bx + az * bw;
}

// ---- training-step neighbour: photometric loss --------------------------
// loss = sum_i w_l2 * (a_i - b_i)^2 + w_l1 * |a_i - b_i|   (caller divides by n for the mean);
// one pass over image and target, block reduction, one atomic per block.
__global__ void __launch_bounds__(256) k_photometric_loss(const float4* __restrict__ a, const float4* __restrict__ b,
                                                          size_t n4, size_t n, float w_l2, float w_l1,
                                                          float* __restrict__ out) {
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 x = __ldg(a + i), y = __ldg(b + i);
    const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    acc += w_l2 * (d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3) + w_l1 * (fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {     // up to three tail elements
    const size_t i = n4 * 4 + threadIdx.x;
    const float d = reinterpret_cast<const float*>(a)[i] - reinterpret_cast<const float*>(b)[i];
    acc += w_l2 * d * d + w_l1 * fabsf(d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; w++) t += s[w];
    atomicAdd(out, t);
  }
}

// dL/da_i = g * (2 w_l2 (a_i - b_i) + w_l1 sign(a_i - b_i)),  g = *upstream (device scalar) * scale
__global__ void __launch_bounds__(256) k_photometric_loss_bwd(const float4* __restrict__ a, const float4* __restrict__ b,
                                                              size_t n4, size_t n, float w_l2, float w_l1, float scale,
                                                              const float* __restrict__ upstream, float4* __restrict__ g) {
  const float u = __ldg(upstream) * scale;
  auto f = [&](float d) { return u * (2.f * w_l2 * d + w_l1 * ((d > 0.f) - (d < 0.f))); };
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const size_t i = n4 * 4 + threadIdx.x;
    reinterpret_cast<float*>(g)[i] = f(reinterpret_cast<const float*>(a)[i] - reinterpret_cast<const float*>(b)[i]);
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 x = __ldg(a + i), y = __ldg(b + i);
    g[i] = make_float4(f(x.x - y.x), f(x.y - y.y), f(x.z - y.z), f(x.w - y.w));
  }
}

void launch_photometric_loss(const float* a, const float* b, size_t n, float w_l2, float w_l1, float* out,
                             cudaStream_t st) {
  cudaMemsetAsync(out, 0, sizeof(float), st);
  if (n == 0) return;
  k_photometric_loss<<<148 * 8, 256, 0, st>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                              n / 4, n, w_l2, w_l1, out);
  count_launch();
}

void launch_photometric_loss_bwd(const float* a, const float* b, size_t n, float w_l2, float w_l1, float scale,
                                 const float* upstream, float* g, cudaStream_t st) {
  if (n == 0) return;
  k_photometric_loss_bwd<<<148 * 8, 256, 0, st>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                                  n / 4, n, w_l2, w_l1, scale, upstream, reinterpret_cast<float4*>(g));
  count_launch();
}

int launch_ply_activate(int P, const float* v, const B200GSPlyLayout& L, float* means, float* shs, float* opac,
                        float* scales, float* rots, cudaStream_t st) {
  if (P == 0) return 0;
  const size_t smem = (size_t)PLY_BLOCK * L.stride * sizeof(float);
  if (smem > 200 * 1024) { set_error("ply_activate: vertex stride %d too large", L.stride); return B200GS_ERR_INVALID_ARG; }
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(k_ply_activate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_ply_activate<<<(P + PLY_BLOCK - 1) / PLY_BLOCK, PLY_BLOCK, smem, st>>>(P, v, L, means, shs, opac, scales, rots);
  count_launch();
  return 0;
}

void launch_transform_gaussians(int n, const float* means_in, const float* rots_in, const int32_t* link_ids,
                                const float* T, const float* Q, int L, float* means_out, float* rots_out,
                                cudaStream_t st) {
  if (n == 0) return;
  k_transform_gaussians<<<(n + 255) / 256, 256, 0, st>>>(n, means_in, rots_in, link_ids, T, Q, L, means_out, rots_out);
  count_launch();
}

}  // namespace b200gs

// train.cu -- training-step neighbours of the rasterizer: fused SSIM
// (forward + adjoint) and a fused multi-tensor Adam step.
//
// These are the steps either side of the rasterizer in the 3DGS reconstruction loop the reference
// delegates (/root/reference/README.md:75): loss = (1 - l) L1 + l (1 - SSIM), then Adam over the
// 59 floats of every Gaussian.  In eager PyTorch the SSIM alone is five grouped 11x11 convolutions
// plus ~20 elementwise launches and their autograd twins; here it is one kernel forward and one
// backward, both HBM-streaming with the 11-tap separable window applied out of shared memory.
//
// SSIM definition (the one every public 3DGS trainer uses): 11x11 Gaussian window, sigma 1.5,
// zero padding, C1 = 0.01^2, C2 = 0.03^2, per channel; result = mean of the SSIM map.
#include "common.cuh"
#include "kernels.cuh"

namespace b200gs {

constexpr int SS_T = 16;            // output tile edge
constexpr int SS_R = 5;             // window radius
constexpr int SS_E = SS_T + 2 * SS_R;   // 26: tile + halo
constexpr float SS_C1 = 0.01f * 0.01f;
constexpr float SS_C2 = 0.03f * 0.03f;

// exp(-(i-5)^2 / (2 * 1.5^2)) / sum, rounded to fp32 (the window torch builds in fp64 and casts).  Compile-time
// constants in the kernel image: valid on every device of the process and inside stream capture.
__device__ __constant__ float c_gauss[11] = {0.001028380123898387f, 0.0075987582094967365f, 0.036000773310661316f, 0.10936068743467331f, 0.21300554275512695f, 0.26601171493530273f, 0.21300554275512695f, 0.10936068743467331f, 0.036000773310661316f, 0.0075987582094967365f, 0.001028380123898387f};

// img1 = rendered image (gradient flows to it), img2 = target.  [C][H][W].
// Writes the three derivative maps the adjoint needs (dm/dmu1, dm/dE[x^2], dm/dE[xy]) when
// `maps` != nullptr ([3][C][H][W]) and accumulates sum(ssim_map) into *out_sum.
__global__ void __launch_bounds__(SS_T * SS_T) k_ssim_fwd(const float* __restrict__ img1,
                                                          const float* __restrict__ img2, int H, int W,
                                                          float* __restrict__ maps, float* __restrict__ out_sum) {
  __shared__ float s1[SS_E][SS_E + 1], s2[SS_E][SS_E + 1];
  __shared__ float h[5][SS_E][SS_T + 1];
  __shared__ float wsum[SS_T * SS_T / 32];

  const int tid = threadIdx.y * SS_T + threadIdx.x;
  const int ch = blockIdx.z;
  const size_t plane = (size_t)H * W;
  const float* p1 = img1 + ch * plane;
  const float* p2 = img2 + ch * plane;
  const int x0 = blockIdx.x * SS_T - SS_R, y0 = blockIdx.y * SS_T - SS_R;

  for (int i = tid; i < SS_E * SS_E; i += SS_T * SS_T) {
    const int ly = i / SS_E, lx = i % SS_E;
    const int gx = x0 + lx, gy = y0 + ly;
    const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
    s1[ly][lx] = in ? __ldg(p1 + (size_t)gy * W + gx) : 0.f;
    s2[ly][lx] = in ? __ldg(p2 + (size_t)gy * W + gx) : 0.f;
  }
  __syncthreads();
  // horizontal pass: SS_E rows x SS_T columns
  for (int i = tid; i < SS_E * SS_T; i += SS_T * SS_T) {
    const int ly = i / SS_T, lx = i % SS_T;
    float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
    for (int k = 0; k < 11; k++) {
      const float g = c_gauss[k], x = s1[ly][lx + k], y = s2[ly][lx + k];
      a = __fmaf_rn(g, x, a); b = __fmaf_rn(g, y, b);
      aa = __fmaf_rn(g, x * x, aa); bb = __fmaf_rn(g, y * y, bb); ab = __fmaf_rn(g, x * y, ab);
    }
    h[0][ly][lx] = a; h[1][ly][lx] = b; h[2][ly][lx] = aa; h[3][ly][lx] = bb; h[4][ly][lx] = ab;
  }
  __syncthreads();
  const int lx = threadIdx.x, ly = threadIdx.y;
  const int gx = blockIdx.x * SS_T + lx, gy = blockIdx.y * SS_T + ly;
  float val = 0.f;
  if (gx < W && gy < H) {
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; k++) {
      const float g = c_gauss[k];
      mu1 = __fmaf_rn(g, h[0][ly + k][lx], mu1); mu2 = __fmaf_rn(g, h[1][ly + k][lx], mu2);
      e11 = __fmaf_rn(g, h[2][ly + k][lx], e11); e22 = __fmaf_rn(g, h[3][ly + k][lx], e22);
      e12 = __fmaf_rn(g, h[4][ly + k][lx], e12);
    }
    const float mu1sq = mu1 * mu1, mu2sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float sg1 = e11 - mu1sq, sg2 = e22 - mu2sq, sg12 = e12 - mu12;
    const float A1 = 2.f * mu12 + SS_C1, A2 = 2.f * sg12 + SS_C2;
    const float B1 = mu1sq + mu2sq + SS_C1, B2 = sg1 + sg2 + SS_C2;
    const float inv = 1.f / (B1 * B2);
    val = A1 * A2 * inv;
    if (maps) {
      const size_t o = ch * plane + (size_t)gy * W + gx;
      const size_t cp = (size_t)gridDim.z * plane;
      // total derivative w.r.t. mu1 with E[x^2], E[xy] held fixed (sigma terms depend on mu1 too)
      maps[o] = 2.f * mu2 * (A2 - A1) * inv - val * 2.f * mu1 * (1.f / B1 - 1.f / B2);
      maps[cp + o] = -val / B2;
      maps[2 * cp + o] = 2.f * A1 * inv;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
  if ((tid & 31) == 0) wsum[tid >> 5] = val;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < SS_T * SS_T / 32; i++) s += wsum[i];
    atomicAdd(out_sum, s);
  }
}

// dL/dimg1(p) = u * sum_q g(q-p) [ m0(q) + 2 x(p) m1(q) + y(p) m2(q) ],  u = *upstream * scale
__global__ void __launch_bounds__(SS_T * SS_T) k_ssim_bwd(const float* __restrict__ img1,
                                                          const float* __restrict__ img2,
                                                          const float* __restrict__ maps, int H, int W, float scale,
                                                          const float* __restrict__ upstream,
                                                          float* __restrict__ dL_dimg1) {
  __shared__ float s[3][SS_E][SS_E + 1];
  __shared__ float h[3][SS_E][SS_T + 1];
  const int tid = threadIdx.y * SS_T + threadIdx.x;
  const int ch = blockIdx.z;
  const size_t plane = (size_t)H * W, cp = (size_t)gridDim.z * plane;
  const int x0 = blockIdx.x * SS_T - SS_R, y0 = blockIdx.y * SS_T - SS_R;
  for (int i = tid; i < SS_E * SS_E; i += SS_T * SS_T) {
    const int ly = i / SS_E, lx = i % SS_E;
    const int gx = x0 + lx, gy = y0 + ly;
    const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
    const size_t o = ch * plane + (size_t)(in ? gy : 0) * W + (in ? gx : 0);
#pragma unroll
    for (int m = 0; m < 3; m++) s[m][ly][lx] = in ? __ldg(maps + m * cp + o) : 0.f;
  }
  __syncthreads();
  for (int i = tid; i < SS_E * SS_T; i += SS_T * SS_T) {
    const int ly = i / SS_T, lx = i % SS_T;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; k++) {
      const float g = c_gauss[k];
      a0 = __fmaf_rn(g, s[0][ly][lx + k], a0);
      a1 = __fmaf_rn(g, s[1][ly][lx + k], a1);
      a2 = __fmaf_rn(g, s[2][ly][lx + k], a2);
    }
    h[0][ly][lx] = a0; h[1][ly][lx] = a1; h[2][ly][lx] = a2;
  }
  __syncthreads();
  const int lx = threadIdx.x, ly = threadIdx.y;
  const int gx = blockIdx.x * SS_T + lx, gy = blockIdx.y * SS_T + ly;
  if (gx < W && gy < H) {
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; k++) {
      const float g = c_gauss[k];
      c0 = __fmaf_rn(g, h[0][ly + k][lx], c0);
      c1 = __fmaf_rn(g, h[1][ly + k][lx], c1);
      c2 = __fmaf_rn(g, h[2][ly + k][lx], c2);
    }
    const size_t o = ch * plane + (size_t)gy * W + gx;
    const float x = __ldg(img1 + o), y = __ldg(img2 + o);
    dL_dimg1[o] = __ldg(upstream) * scale * (c0 + 2.f * x * c1 + y * c2);
  }
}

void launch_ssim_fwd(const float* img1, const float* img2, int C, int H, int W, float* maps, float* out_sum,
                     cudaStream_t st) {
  cudaMemsetAsync(out_sum, 0, sizeof(float), st);
  const dim3 grid((W + SS_T - 1) / SS_T, (H + SS_T - 1) / SS_T, C), block(SS_T, SS_T);
  k_ssim_fwd<<<grid, block, 0, st>>>(img1, img2, H, W, maps, out_sum);
  count_launch();
}

void launch_ssim_bwd(const float* img1, const float* img2, const float* maps, int C, int H, int W, float scale,
                     const float* upstream, float* dL_dimg1, cudaStream_t st) {
  const dim3 grid((W + SS_T - 1) / SS_T, (H + SS_T - 1) / SS_T, C), block(SS_T, SS_T);
  k_ssim_bwd<<<grid, block, 0, st>>>(img1, img2, maps, H, W, scale, upstream, dL_dimg1);
  count_launch();
}

// ==================================================================================================
// Fused multi-tensor Adam (torch.optim.Adam semantics: no weight decay, no amsgrad):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// One launch updates every parameter group (means, SH, opacity, scale, rotation -- each with its own
// learning rate); a block looks its group up in a small prefix table.  Streams 16 B/float read + 12 B
// written: HBM-bound.
// ==================================================================================================
struct AdamArgs {
  float* p[B200GS_ADAM_MAX_GROUPS];
  const float* g[B200GS_ADAM_MAX_GROUPS];
  float* m[B200GS_ADAM_MAX_GROUPS];
  float* v[B200GS_ADAM_MAX_GROUPS];
  long long n[B200GS_ADAM_MAX_GROUPS];
  float step_size[B200GS_ADAM_MAX_GROUPS];   // lr / bias_correction1
  unsigned block_end[B200GS_ADAM_MAX_GROUPS]; // exclusive prefix of blocks per group
  int num_groups;
  float beta1, beta2, eps, inv_sqrt_bc2;
};

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_PER_BLOCK = ADAM_THREADS * 4 * 4;   // 4 float4 per thread

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float b1, float b2, float eps,
                                      float inv_sqrt_bc2, float step) {
  m = __fmaf_rn(b1, m, (1.f - b1) * g);
  v = __fmaf_rn(b2, v, (1.f - b2) * g * g);
  const float denom = __fmaf_rn(sqrtf(v), inv_sqrt_bc2, eps);
  p -= step * (m / denom);
}

__global__ void __launch_bounds__(ADAM_THREADS) k_adam(AdamArgs a) {
  int grp = 0;
  while (grp + 1 < a.num_groups && blockIdx.x >= a.block_end[grp]) grp++;
  const unsigned first = grp ? a.block_end[grp - 1] : 0u;
  const long long base = (long long)(blockIdx.x - first) * ADAM_PER_BLOCK;
  const long long n = a.n[grp];
  float* __restrict__ P = a.p[grp];
  const float* __restrict__ G = a.g[grp];
  float* __restrict__ M = a.m[grp];
  float* __restrict__ V = a.v[grp];
  const float step = a.step_size[grp];
  const bool vec = ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(M) |
                     reinterpret_cast<uintptr_t>(V)) & 15) == 0;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const long long i = base + ((long long)r * ADAM_THREADS + threadIdx.x) * 4;
    if (i >= n) break;
    if (vec && i + 3 < n) {
      float4 p = *reinterpret_cast<float4*>(P + i), m = *reinterpret_cast<float4*>(M + i),
             v = *reinterpret_cast<float4*>(V + i);
      const float4 g = __ldg(reinterpret_cast<const float4*>(G + i));
      adam1(p.x, g.x, m.x, v.x, a.beta1, a.beta2, a.eps, a.inv_sqrt_bc2, step);
      adam1(p.y, g.y, m.y, v.y, a.beta1, a.beta2, a.eps, a.inv_sqrt_bc2, step);
      adam1(p.z, g.z, m.z, v.z, a.beta1, a.beta2, a.eps, a.inv_sqrt_bc2, step);
      adam1(p.w, g.w, m.w, v.w, a.beta1, a.beta2, a.eps, a.inv_sqrt_bc2, step);
      *reinterpret_cast<float4*>(P + i) = p;
      *reinterpret_cast<float4*>(M + i) = m;
      *reinterpret_cast<float4*>(V + i) = v;
    } else {
      for (long long k = i; k < n && k < i + 4; k++) {
        float p = P[k], m = M[k], v = V[k];
        adam1(p, G[k], m, v, a.beta1, a.beta2, a.eps, a.inv_sqrt_bc2, step);
        P[k] = p; M[k] = m; V[k] = v;
      }
    }
  }
}

int launch_adam(const B200GSAdamGroup* groups, int num_groups, float beta1, float beta2, float eps, int step,
                cudaStream_t st) {
  AdamArgs a;
  a.num_groups = 0;
  a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  unsigned blocks = 0;
  for (int i = 0; i < num_groups; i++) {
    if (groups[i].n <= 0) continue;
    const int k = a.num_groups++;
    a.p[k] = groups[i].param; a.g[k] = groups[i].grad; a.m[k] = groups[i].exp_avg; a.v[k] = groups[i].exp_avg_sq;
    a.n[k] = groups[i].n;
    a.step_size[k] = (float)((double)groups[i].lr / bc1);
    blocks += (unsigned)((groups[i].n + ADAM_PER_BLOCK - 1) / ADAM_PER_BLOCK);
    a.block_end[k] = blocks;
  }
  if (blocks == 0) return 0;
  k_adam<<<blocks, ADAM_THREADS, 0, st>>>(a);
  count_launch();
  return 0;
}

}  // namespace b200gs

// context.cu -- B200GSContext: the no-stall protocol of the forward pass behind the C ABI (include/b200gs.h).
//
// The interface this library replaces (_C.rasterize_gaussians of the public
// diff-gaussian-rasterization) reads the pair count back and synchronises inside every call.  b200gs_forward makes
// that optional but leaves the bookkeeping -- scratch memory, the pair-capacity hint, the bin size that suits the
// scene -- to its caller; the Python operator layer (robosimgs_b200/rasterizer.py) keeps it per (device, P, H, W).
// A context is the same bookkeeping for callers that are not Python: a C / pybind host gets the speculative and the
// deferred path, the adaptive bin size and reusable scratch from one handle.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace b200gs {

// ---- policy rules (pure host functions; rasterizer.py calls the same two) ----------------------------------------
static int64_t policy_pair_capacity(int64_t D) { return D > 0 ? D + (D >> 4) + 32768 : 0; }

static int32_t policy_bin_shift(int64_t D, int64_t touching, int32_t used_shift, float coverage) {
  if (D <= 0 || touching <= 0 || used_shift < 0) return used_shift;
  const double b = (double)(16 << used_shift);
  // pairs per touching Gaussian ~ (1 + extent / bin)^2  ->  typical splat extent on screen
  const double extent = std::fmax(std::sqrt(std::fmax((double)D / (double)touching, 1.0)) - 1.0, 0.0) * b;
  int s = (int)std::lround(std::log2(std::fmax(3.0 * extent, 16.0) / 16.0));
  s = s < 1 ? 1 : (s > 4 ? 4 : s);
  // large splats AND a frame that saturates everywhere: tiles stop after the first few records of their list, so
  // one size coarser costs the compositing kernels nothing and makes emission and sort lighter
  if (s == 3 && coverage > 0.995f) s = 4;
  return s;
}

// ---- frame statistics for the bin-size policy: visible Gaussians and sum of (1 - final T) --------------------------
__global__ void __launch_bounds__(256) k_policy_stats(const int32_t* __restrict__ radii, int P, const float4* __restrict__ pix,
                                                      size_t npx, uint32_t* __restrict__ touching, float* __restrict__ cov_sum) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t t = 0;
  float c = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)P; i += stride) t += radii[i] > 0 ? 1u : 0u;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += stride) c += 1.f - pix[i].w;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    t += __shfl_xor_sync(0xffffffffu, t, d);
    c += __shfl_xor_sync(0xffffffffu, c, d);
  }
  if ((threadIdx.x & 31) == 0) {
    if (t) atomicAdd(touching, t);
    atomicAdd(cov_sum, c);
  }
}

struct Arena {
  char* p = nullptr;
  size_t cap = 0;
  cudaStream_t stream = nullptr;
};

static char* arena_resize(void* vctx, size_t bytes) {
  Arena* a = static_cast<Arena*>(vctx);
  if (bytes <= a->cap && a->p) return a->p;
  // stream-ordered free + allocation: frames still running on this stream keep their memory until they are done
  if (a->p) cudaFreeAsync(a->p, a->stream);
  a->p = nullptr; a->cap = 0;
  size_t want = bytes + bytes / 4 + 4096;
  want = (want + 255) & ~size_t(255);
  void* q = nullptr;
  if (cudaMallocAsync(&q, want, a->stream) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  a->p = static_cast<char*>(q); a->cap = want;
  return a->p;
}

constexpr int CTX_TICKETS = 64;

struct Policy {
  int64_t tracked = 0;   // decaying maximum of recent pair counts
  int shift = -1;        // -1: the library's automatic bin size
  int64_t calls = 0;     // synchronous calls seen (the policy looks at the 1st and every 256th)
};

struct Ticket {
  bool busy = false;
  int64_t id = -1, hint = 0;
  std::tuple<int, int, int> key;
  cudaEvent_t ev = nullptr;
};

}  // namespace b200gs

using namespace b200gs;

struct B200GSContext {
  int device = 0;
  std::mutex mu;
  std::map<std::tuple<int, int, int>, Policy> pol;
  Arena geom, binning, img, scratch;
  uint32_t* pinned = nullptr;      // [CTX_TICKETS] pair counts of deferred frames + [4] policy statistics
  uint32_t* dev_stats = nullptr;   // [2] device side of the policy statistics
  Ticket tickets[CTX_TICKETS];
  int64_t next_ticket = 0;
};

static int default_bin_shift(int H, int W) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  int s = 0;
  while (s < 3 && (((gx + (1 << s) - 1) >> s) * ((gy + (1 << s) - 1) >> s)) > 255) s++;
  return s;
}

extern "C" {

int64_t b200gs_policy_pair_capacity(int64_t tracked_pairs) { return policy_pair_capacity(tracked_pairs); }
int32_t b200gs_policy_bin_shift(int64_t D, int64_t touching, int32_t used_shift, float coverage) {
  return policy_bin_shift(D, touching, used_shift, coverage);
}

int b200gs_context_create(B200GSContext** out_ctx) {
  if (!out_ctx) { set_error("context_create: out_ctx is NULL"); return B200GS_ERR_INVALID_ARG; }
  B200GSContext* c = new B200GSContext();
  int rc = check_cuda(cudaGetDevice(&c->device), "context_create: cudaGetDevice");
  if (!rc) rc = check_cuda(cudaHostAlloc(reinterpret_cast<void**>(&c->pinned), sizeof(uint32_t) * (CTX_TICKETS + 4),
                                         cudaHostAllocPortable), "context_create: pinned words");
  if (!rc) rc = check_cuda(cudaMalloc(reinterpret_cast<void**>(&c->dev_stats), 2 * sizeof(uint32_t)), "context_create: stats");
  for (int i = 0; !rc && i < CTX_TICKETS; i++)
    rc = check_cuda(cudaEventCreateWithFlags(&c->tickets[i].ev, cudaEventDisableTiming), "context_create: event");
  if (rc) { b200gs_context_destroy(c); return rc; }
  *out_ctx = c;
  return 0;
}

int b200gs_context_destroy(B200GSContext* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  for (Arena* a : {&c->geom, &c->binning, &c->img, &c->scratch})
    if (a->p) { cudaStreamSynchronize(a->stream); cudaFree(a->p); }
  for (auto& t : c->tickets) if (t.ev) cudaEventDestroy(t.ev);
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->dev_stats) cudaFree(c->dev_stats);
  delete c;
  return 0;
}

int b200gs_context_query(B200GSContext* c, int32_t P, int32_t H, int32_t W, int64_t* tracked_pairs, int32_t* bin_shift) {
  if (!c) { set_error("context_query: ctx is NULL"); return B200GS_ERR_INVALID_ARG; }
  std::lock_guard<std::mutex> lk(c->mu);
  auto it = c->pol.find(std::make_tuple((int)P, (int)H, (int)W));
  if (tracked_pairs) *tracked_pairs = it == c->pol.end() ? 0 : it->second.tracked;
  if (bin_shift) *bin_shift = it == c->pol.end() ? -1 : it->second.shift;
  return 0;
}

int b200gs_context_forward(B200GSContext* c, const B200GSParams* prm, const float* bg, const float* viewmatrix,
                           const float* projmatrix, const float* campos, const float* means3D, const float* shs,
                           const float* colors_precomp, const float* opacities, const float* scales,
                           const float* rotations, const float* cov3D_precomp, float* out_color, int32_t* radii,
                           B200GSAlloc geom, B200GSAlloc binning, B200GSAlloc img, int32_t defer,
                           int32_t* num_rendered, int64_t* ticket, int32_t* used_flags, void* stream) {
  if (!c || !prm || !ticket || !used_flags || (!defer && !num_rendered)) {
    set_error("context_forward: ctx/prm/ticket/used_flags (and num_rendered unless deferred) must not be NULL");
    return B200GS_ERR_INVALID_ARG;
  }
  std::lock_guard<std::mutex> lk(c->mu);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const auto key = std::make_tuple((int)prm->P, (int)prm->image_height, (int)prm->image_width);
  Policy& pol = c->pol[key];
  const int64_t hint = prm->debug ? 0 : policy_pair_capacity(pol.tracked);
  const int used_shift = pol.shift >= 0 ? pol.shift : default_bin_shift(prm->image_height, prm->image_width);
  B200GSParams p2 = *prm;
  p2.pair_capacity_hint = hint;
  p2.flags = (pol.shift >= 0 ? B200GS_BIN_SHIFT_HINT(pol.shift) : 0) | (prm->flags & (B200GS_FORWARD_ONLY | B200GS_OUT_RGB8));
  *used_flags = p2.flags;
  *ticket = -1;
  for (Arena* a : {&c->geom, &c->binning, &c->img}) a->stream = st;
  if (!geom.resize) geom = B200GSAlloc{&c->geom, arena_resize};
  if (!binning.resize) binning = B200GSAlloc{&c->binning, arena_resize};
  if (!img.resize) img = B200GSAlloc{&c->img, arena_resize};

  if (defer && hint > 0 && prm->P > 0) {
    int slot = -1;
    for (int i = 0; i < CTX_TICKETS; i++) if (!c->tickets[i].busy) { slot = i; break; }
    if (slot < 0) { set_error("context_forward: %d deferred frames in flight, wait for a ticket first", CTX_TICKETS); return B200GS_ERR_INVALID_ARG; }
    p2.flags |= B200GS_DEFER_PAIR_CHECK;
    int rc = b200gs_forward(&p2, bg, viewmatrix, projmatrix, campos, means3D, shs, colors_precomp, opacities, scales,
                            rotations, cov3D_precomp, out_color, radii, geom, binning, img,
                            reinterpret_cast<int32_t*>(c->pinned + slot), stream);
    if (rc) return rc;
    Ticket& t = c->tickets[slot];
    if ((rc = check_cuda(cudaEventRecord(t.ev, st), "context_forward: event record"))) return rc;
    t.busy = true; t.id = c->next_ticket++; t.hint = hint; t.key = key;
    *ticket = t.id;
    return 0;
  }

  int32_t D = 0;
  int rc = b200gs_forward(&p2, bg, viewmatrix, projmatrix, campos, means3D, shs, colors_precomp, opacities, scales,
                          rotations, cov3D_precomp, out_color, radii, geom, binning, img, &D, stream);
  if (rc) return rc;
  if (num_rendered) *num_rendered = D;
  pol.tracked = std::max<int64_t>(D, (int64_t)(pol.tracked * 0.97));
  pol.calls++;
  // bin-size policy: looks at the first synchronous frame of a (P, H, W) and at every 256th
  if (!prm->debug && D > 0 && (pol.calls == 1 || pol.calls % 256 == 0)) {
    size_t img_off = 0;   // ImgBuf: ranges[tiles16] first, then pix (128-byte aligned chunks)
    {
      const size_t tiles = (size_t)((prm->image_width + TILE - 1) / TILE) * ((prm->image_height + TILE - 1) / TILE);
      img_off = (tiles * sizeof(uint2) + 127) & ~size_t(127);
    }
    char* img_p = img.resize(img.ctx, 0);   // the buffer the forward call just used (resize never shrinks)
    if (img_p) {
      const size_t npx = (size_t)prm->image_height * prm->image_width;
      uint32_t* hs = c->pinned + CTX_TICKETS;
      rc = check_cuda(cudaMemsetAsync(c->dev_stats, 0, 2 * sizeof(uint32_t), st), "policy stats clear");
      if (!rc) {
        k_policy_stats<<<148 * 4, 256, 0, st>>>(radii, prm->P, reinterpret_cast<const float4*>(img_p + img_off), npx,
                                                c->dev_stats, reinterpret_cast<float*>(c->dev_stats + 1));
        count_launch();
        rc = check_cuda(cudaMemcpyAsync(hs, c->dev_stats, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "policy stats copy");
      }
      if (!rc) rc = check_cuda(cudaStreamSynchronize(st), "policy stats sync");
      if (rc) return rc;
      float cov_sum;
      memcpy(&cov_sum, hs + 1, sizeof(float));
      const int s = policy_bin_shift(D, (int64_t)hs[0], used_shift, cov_sum / (float)npx);
      if (s != used_shift) { pol.shift = s; pol.tracked = 0; }   // the pair count changes with the bin size
    }
  }
  return 0;
}

int b200gs_context_ticket_wait(B200GSContext* c, int64_t ticket, int32_t* num_rendered, int32_t* complete) {
  if (!c || ticket < 0) { set_error("ticket_wait: invalid ticket"); return B200GS_ERR_INVALID_ARG; }
  std::lock_guard<std::mutex> lk(c->mu);
  for (int i = 0; i < CTX_TICKETS; i++) {
    Ticket& t = c->tickets[i];
    if (!t.busy || t.id != ticket) continue;
    if (int rc = check_cuda(cudaEventSynchronize(t.ev), "ticket_wait: event")) return rc;
    const int64_t D = c->pinned[i];
    if (num_rendered) *num_rendered = (int32_t)D;
    if (complete) *complete = D <= t.hint ? 1 : 0;
    Policy& pol = c->pol[t.key];
    pol.tracked = std::max<int64_t>(D, (int64_t)(pol.tracked * 0.97));
    t.busy = false;
    return 0;
  }
  set_error("ticket_wait: unknown ticket %lld", (long long)ticket);
  return B200GS_ERR_INVALID_ARG;
}

int b200gs_context_backward(B200GSContext* c, const B200GSParams* prm, int32_t used_flags, const float* bg,
                            const float* viewmatrix, const float* projmatrix, const float* campos,
                            const float* means3D, const float* shs, const float* colors_precomp,
                            const float* opacities, const float* scales, const float* rotations,
                            const float* cov3D_precomp, const int32_t* radii, const char* geom, const char* binning,
                            const char* img, int32_t num_rendered, const float* dL_dout_color, float* dL_dmeans3D,
                            float* dL_dmeans2D, float* dL_dshs, float* dL_dcolors_precomp, float* dL_dopacities,
                            float* dL_dscales, float* dL_drotations, float* dL_dcov3D, void* stream) {
  if (!c || !prm) { set_error("context_backward: ctx/prm is NULL"); return B200GS_ERR_INVALID_ARG; }
  std::lock_guard<std::mutex> lk(c->mu);
  B200GSParams p2 = *prm;
  p2.flags = used_flags & ~B200GS_DEFER_PAIR_CHECK;
  p2.pair_capacity_hint = 0;
  c->scratch.stream = static_cast<cudaStream_t>(stream);
  return b200gs_backward(&p2, bg, viewmatrix, projmatrix, campos, means3D, shs, colors_precomp, opacities, scales, rotations,
                         cov3D_precomp, radii, geom ? geom : c->geom.p, binning ? binning : c->binning.p,
                         img ? img : c->img.p, num_rendered, dL_dout_color, dL_dmeans3D, dL_dmeans2D, dL_dshs,
                         dL_dcolors_precomp, dL_dopacities, dL_dscales, dL_drotations, dL_dcov3D,
                         B200GSAlloc{&c->scratch, arena_resize}, stream);
}

// ---- captured frames with per-kernel priorities -----------------------------------------------------------------------
// A frame is a chain of short, latency-bound kernels (projection, bucket scan, pair emission, pair sort) followed by one
// long issue-bound kernel (compositing).  When several captured frames are in flight on different streams the block
// scheduler serves grids in arrival order: the next frame's chain only gets SMs once the compositing grid of the frame
// in front has placed its last CTA.  Instantiating the frame graph with cudaGraphInstantiateFlagUseNodePriority and
// marking the chain kernels HIGH, the compositing / export kernels LOW lets a chain run *underneath* another frame's
// compositing (its CTAs take the slots that compositing CTAs vacate) instead of behind it.
int b200gs_graph_instantiate(void* graph, int32_t mode, void** exec_out, int32_t* n_low, int32_t* n_high) {
  if (!graph || !exec_out) { set_error("graph_instantiate: graph/exec_out is NULL"); return B200GS_ERR_INVALID_ARG; }
  cudaGraph_t g = static_cast<cudaGraph_t>(graph);
  int lo = 0, hi = 0;
  int low_count = 0, high_count = 0;
  int rc;
  if (mode != 0) {
    if ((rc = check_cuda(cudaDeviceGetStreamPriorityRange(&lo, &hi), "graph_instantiate: priority range"))) return rc;
    size_t n = 0;
    if ((rc = check_cuda(cudaGraphGetNodes(g, nullptr, &n), "graph_instantiate: node count"))) return rc;
    std::vector<cudaGraphNode_t> nodes(n);
    if (n && (rc = check_cuda(cudaGraphGetNodes(g, nodes.data(), &n), "graph_instantiate: nodes"))) return rc;
    for (size_t i = 0; i < n; i++) {
      cudaGraphNodeType ty;
      if ((rc = check_cuda(cudaGraphNodeGetType(nodes[i], &ty), "graph_instantiate: node type"))) return rc;
      if (ty != cudaGraphNodeTypeKernel) continue;
      cudaKernelNodeParams kp;
      if ((rc = check_cuda(cudaGraphKernelNodeGetParams(nodes[i], &kp), "graph_instantiate: kernel params"))) return rc;
      const char* name = nullptr;
      if (cudaFuncGetName(&name, kp.func) != cudaSuccess) { cudaGetLastError(); name = nullptr; }
      // compositing (k_render_fwd*) and the 8-bit export run LOW; everything in front of them HIGH
      const bool low = name && (strstr(name, "k_render_fwd") || strstr(name, "k_export_rgb8"));
      cudaKernelNodeAttrValue v;
      memset(&v, 0, sizeof(v));
      v.priority = low ? lo : (mode == 2 ? (lo + hi) / 2 : hi);
      if ((rc = check_cuda(cudaGraphKernelNodeSetAttribute(nodes[i], cudaKernelNodeAttributePriority, &v),
                           "graph_instantiate: set priority")))
        return rc;
      (low ? low_count : high_count)++;
    }
  }
  cudaGraphExec_t ex = nullptr;
  if ((rc = check_cuda(cudaGraphInstantiateWithFlags(&ex, g, mode != 0 ? cudaGraphInstantiateFlagUseNodePriority : 0),
                       "graph_instantiate")))
    return rc;
  *exec_out = ex;
  if (n_low) *n_low = low_count;
  if (n_high) *n_high = high_count;
  return 0;
}

int b200gs_graph_launch(void* exec, void* stream) {
  if (!exec) { set_error("graph_launch: exec is NULL"); return B200GS_ERR_INVALID_ARG; }
  return check_cuda(cudaGraphLaunch(static_cast<cudaGraphExec_t>(exec), static_cast<cudaStream_t>(stream)), "graph_launch");
}

int b200gs_graph_exec_destroy(void* exec) {
  if (!exec) return 0;
  return check_cuda(cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(exec)), "graph_exec_destroy");
}

}  // extern "C"

// api.cu -- the C ABI of libb200gs (include/b200gs.h): argument checks, scratch carving and the
// forward / backward launch sequences.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <deque>
#include <atomic>
#include <functional>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace b200gs {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};  // process-wide: autograd runs backward on its own thread

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return B200GS_ERR_CUDA;
}
int debug_sync(const B200GSParams* prm, cudaStream_t s, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && prm->debug) e = cudaStreamSynchronize(s);
  return check_cuda(e, what);
}

// ---- optional per-stage device timing (bench / roofline accounting) ----------------------------
static const char* kStageNames[B200GS_NUM_STAGES] = {"project", "scan", "emit_pairs", "pair_sort",
                                                     "tile_ranges", "render", "render_bwd", "project_bwd"};
struct StageSpan { int stage; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<StageSpan> g_spans;
static std::vector<cudaEvent_t> g_free_events;

static cudaEvent_t prof_event() {
  cudaEvent_t e;
  if (!g_free_events.empty()) { e = g_free_events.back(); g_free_events.pop_back(); return e; }
  cudaEventCreate(&e);
  return e;
}
struct StageTimer {
  bool on; int stage; cudaStream_t st; cudaEvent_t a;
  StageTimer(int stage_, cudaStream_t st_) : stage(stage_), st(st_) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    on = g_prof_on;
    if (on) { a = prof_event(); cudaEventRecord(a, st); }
  }
  ~StageTimer() {
    if (!on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEvent_t b = prof_event();
    cudaEventRecord(b, st);
    g_spans.push_back({stage, a, b});
  }
};

// log2(bin edge / 16): pairs are binned per (16 << shift)^2 pixels (see project.cu:bin_rect).
// Default: the smallest shift that leaves at most 255 bins (bin id fits 8 key bits -> 40-bit sort
// keys), capped at 3 (64 compositing CTAs share one bin list).  B200GS_BIN_SHIFT overrides.
static std::atomic<int> g_bin_shift_override{-2};   // -2: not initialised, -1: automatic

// binning pipeline: 1 = depth-sliced buckets (bucket.cu: per-bucket counters in the projection kernel, one-CTA
// scan, cursor emission, one warp sorts one bucket; default), 0 = global radix sort of (bin | depth) keys
// (binning.cu / coopsort.cu -- the library-sort pipeline, kept as a cross-check: both give bit-identical lists).
// B200GS_BINNING=bucket|sort or b200gs_set_option("binning", 0|1).
static std::atomic<int> g_binning_mode{-1};
static bool use_bucketed() {
  int m = g_binning_mode.load();
  if (m < 0) {
    const char* e = getenv("B200GS_BINNING");
    m = (e && e[0] == 's') ? 0 : 1;
    g_binning_mode.store(m);
  }
  return m == 1;
}

static int bin_shift_for(int gx, int gy, int flags = 0) {
  int forced = g_bin_shift_override.load();
  if (forced == -2) {
    const char* e = getenv("B200GS_BIN_SHIFT");
    forced = e ? atoi(e) : -1;
    forced = forced < -1 ? -1 : (forced > 5 ? 5 : forced);
    g_bin_shift_override.store(forced);
  }
  if (forced >= 0) return forced;
  // per-call hint (B200GSParams.flags bits 8..11 = shift + 1): the caller tracks the splat extent of its
  // scene/view and asks for bins of a few splat extents; clamped so that bin ids stay below 65535
  const int hinted = ((flags >> 8) & 15) - 1;
  if (hinted >= 0) {
    int s = hinted > 5 ? 5 : hinted;
    while (s < 5 && (((gx + (1 << s) - 1) >> s) * ((gy + (1 << s) - 1) >> s)) >= 65535) s++;
    return s;
  }
  int s = 0;
  while (s < 3 && (((gx + (1 << s) - 1) >> s) * ((gy + (1 << s) - 1) >> s)) > 255) s++;
  return s;
}

// pair sort implementation.  Measured on B200 (tools/sort_ab.py): the single-launch cooperative
// sort wins on small pair lists (C1, 17 k pairs: 60 vs 76 us) where CUB's six dependent launches
// dominate, and loses above a few hundred thousand pairs (C3, 0.77 M: 127 vs 111 us; its five grid
// barriers cost ~5 us each).  Mode 0 = CUB, 1 = automatic by size (default), 2 = cooperative
// whenever it applies.
static std::atomic<int> g_sort_mode{-1};
constexpr int64_t COOP_SORT_AUTO_MAX = 256 * 1024;
static bool use_coop_sort(int64_t n) {
  int m = g_sort_mode.load();
  if (m < 0) {
    const char* e = getenv("B200GS_SORT");
    m = !e ? 1 : (!strcmp(e, "cub") ? 0 : (!strcmp(e, "coop") ? 2 : 1));
    g_sort_mode.store(m);
  }
  if (n <= 0 || m == 0) return false;
  return n <= (m == 2 ? COOP_SORT_MAX_ITEMS : COOP_SORT_AUTO_MAX);
}

// compositing kernels: 1 = four pixels per thread (render4.cu: two warps per tile, about half the
// instructions per pixel), 0 = one pixel per thread (render.cu: eight warps per tile), -1 = automatic
// (default): the four-pixel kernels need enough tiles to fill the GPU with two warps each -- at 1080p
// (8160 tiles) they are 1.6-1.9x faster, at 800x800 (2500 tiles, one partial wave of 64-thread CTAs) the
// one-pixel kernels win by up to 1.6x because four times as many warps hide the latency of each tile's
// serial list walk.  B200GS_RENDER=1px|4px|auto or b200gs_set_option("render", 0|1|-1).
static std::atomic<int> g_render_mode{-2};
constexpr int RENDER4_MIN_TILES = 4096;
static bool use_render4(int num_tiles16) {
  int m = g_render_mode.load();
  if (m == -2) {
    const char* e = getenv("B200GS_RENDER");
    m = !e ? -1 : (!strcmp(e, "1px") ? 0 : (!strcmp(e, "4px") ? 1 : -1));
    g_render_mode.store(m);
  }
  return m < 0 ? num_tiles16 >= RENDER4_MIN_TILES : m == 1;
}

// record slab: 1 = after the bucket sort the 48-byte records are copied into list order (k_build_slab) and the
// four-pixel compositing kernels fill their shared-memory ring with ONE TMA bulk copy per 64-record chunk instead of
// one LDGSTS gather per record; 0 (default) = gather from the per-Gaussian record table.  Needs the bucketed binning.
// Process-wide: must not change between a forward call and its backward call.  b200gs_set_option("slab", 0|1),
// B200GS_SLAB=1.  Measured A/B: DESIGN.md section 5.
static std::atomic<int> g_slab_mode{-1};
static bool use_slab() {
  int m = g_slab_mode.load();
  if (m < 0) {
    const char* e = getenv("B200GS_SLAB");
    m = (e && atoi(e) == 1) ? 1 : 0;
    g_slab_mode.store(m);
  }
  return m == 1;
}

// backward zero-fill: 1 = the gradient tensors are zero-filled with memsets on a side stream WHILE the compositing
// adjoint runs (it is issue-bound and leaves HBM idle), and the projection adjoint then writes only the rows of
// Gaussians that reached the image; 0 = the projection adjoint writes every row itself (one pass, no memset).
// b200gs_set_option("bwd_overlap", 0|1), B200GS_BWD_OVERLAP=0|1.
static std::atomic<int> g_bwd_overlap{-1};
static bool use_bwd_overlap() {
  int m = g_bwd_overlap.load();
  if (m < 0) {
    const char* e = getenv("B200GS_BWD_OVERLAP");
    m = (e && atoi(e) == 1) ? 1 : 0;
    g_bwd_overlap.store(m);
  }
  return m == 1;
}
struct SideStream { cudaStream_t side = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static std::mutex g_side_mu;
static std::deque<std::pair<std::pair<int, cudaStream_t>, SideStream>> g_sides;   // one per (device, caller stream);
                                                                                 // deque: entries never move
static SideStream* side_stream_for(cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_side_mu);
  for (auto& e : g_sides) if (e.first.first == dev && e.first.second == st) return &e.second;
  SideStream s;
  if (cudaStreamCreateWithFlags(&s.side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  g_sides.push_back({{dev, st}, s});
  return &g_sides.back().second;
}

// pair keys of the global sort: 32-bit (bin << 24 | quantised depth, exact order restored in the ranges
// pass; default whenever there are at most 255 bins and the library sort is used) or 64-bit.
// b200gs_set_option("sort_keys", 32|64), B200GS_SORT_KEYS=64.
static std::atomic<int> g_key_mode{-1};
static bool allow_key32() {
  int m = g_key_mode.load();
  if (m < 0) {
    const char* e = getenv("B200GS_SORT_KEYS");
    m = (e && atoi(e) == 64) ? 64 : 32;
    g_key_mode.store(m);
  }
  return m == 32;
}

static int tile_bits_for(int num_tiles) {
  int bits = 1;
  while ((1 << bits) <= num_tiles) bits++;  // room for the invalid id == num_tiles
  return bits;
}

static GeomBuf carve_geom(char* base, int P, size_t* bytes) {
  Carver c(base);
  GeomBuf g;
  g.rec = c.take<float4>((size_t)P * REC_F4);
  g.depth_key = c.take<uint32_t>(P);
  g.big_queue = c.take<uint32_t>(P);
  g.tiles = c.take<uint32_t>(P);
  g.offsets = c.take<uint32_t>(P);
  g.clamped = c.take<uint8_t>(P);
  g.counters = c.take<uint32_t>(32);
  g.cub_temp_bytes = scan_temp_bytes(P);
  g.cub_temp = c.take<char>(g.cub_temp_bytes);
  if (bytes) *bytes = c.bytes();
  return g;
}

static BinBuf carve_binning(char* base, int64_t D, int tile_bits, size_t* bytes, bool slab = false) {
  Carver c(base);
  BinBuf b;
  // the FIRST chunk is what the backward pass reads, so its offset must not depend on the capacity: the record
  // slab when there is one (it carries the Gaussian ids), else the sorted id list
  b.slab = slab ? c.take<float4>((size_t)D * REC_F4) : nullptr;
  b.vals_sorted = c.take<uint32_t>(D);
  b.keys_sorted = c.take<uint64_t>(D);
  b.keys = c.take<uint64_t>(D);
  b.vals = c.take<uint32_t>(D);
  b.coop_hist = c.take<uint32_t>(coop_sort_hist_bytes() / sizeof(uint32_t));
  b.win_first = c.take<uint32_t>((size_t)D / BUCKET_WINDOW + BUCKET_BINS_MAX + 2);
  b.big_segs = c.take<uint2>((size_t)D / WARP_SORT_MAX + 2);
  b.cub_temp_bytes = pair_sort_temp_bytes(D, 32 + tile_bits);
  b.cub_temp = c.take<char>(b.cub_temp_bytes);
  if (bytes) *bytes = c.bytes();
  return b;
}

static ImgBuf carve_img(char* base, int H, int W, size_t* bytes) {
  Carver c(base);
  ImgBuf im;
  const size_t tiles = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
  im.ranges = c.take<uint2>(tiles);
  im.pix = c.take<float4>((size_t)H * W);
  im.n_contrib = c.take<uint32_t>((size_t)H * W);
  im.bin_pub = c.take<unsigned long long>(BUCKET_BINS_MAX);   // 512 KB: bucket_count follows without a gap
  im.bucket_count = c.take<uint32_t>(BUCKETS_MAX);
  im.bucket_base = c.take<uint32_t>(BUCKETS_MAX + 4);
  im.bucket_cursor = c.take<uint32_t>(BUCKETS_MAX);
  if (bytes) *bytes = c.bytes();
  return im;
}

static int check_params(const B200GSParams* p) {
  if (!p) { set_error("params is NULL"); return B200GS_ERR_INVALID_ARG; }
  if (p->P < 0 || p->image_height <= 0 || p->image_width <= 0) {
    set_error("invalid sizes P=%d H=%d W=%d", p->P, p->image_height, p->image_width);
    return B200GS_ERR_INVALID_ARG;
  }
  if (p->sh_degree < 0 || p->sh_degree > 3) {
    set_error("sh_degree must be 0..3, got %d", p->sh_degree);
    return B200GS_ERR_INVALID_ARG;
  }
  if (!(p->tanfovx > 0.f) || !(p->tanfovy > 0.f)) {
    set_error("tanfovx/tanfovy must be positive");
    return B200GS_ERR_INVALID_ARG;
  }
  return 0;
}

static int check_inputs(const B200GSParams* p, const float* means3D, const float* shs,
                        const float* colors_precomp, const float* opacities, const float* scales,
                        const float* rotations, const float* cov3D_precomp) {
  if (p->P == 0) return 0;
  if (!means3D || !opacities) { set_error("means3D/opacities must not be NULL"); return B200GS_ERR_INVALID_ARG; }
  if ((shs == nullptr) == (colors_precomp == nullptr)) {
    set_error("provide exactly one of shs / colors_precomp");
    return B200GS_ERR_INVALID_ARG;
  }
  const bool sr = scales != nullptr && rotations != nullptr;
  if (sr == (cov3D_precomp != nullptr) || ((scales != nullptr) != (rotations != nullptr))) {
    set_error("provide exactly one of (scales, rotations) / cov3D_precomp");
    return B200GS_ERR_INVALID_ARG;
  }
  if (shs && p->M < (p->sh_degree + 1) * (p->sh_degree + 1)) {
    set_error("shs holds M=%d coefficients, sh_degree=%d needs %d", p->M, p->sh_degree,
              (p->sh_degree + 1) * (p->sh_degree + 1));
    return B200GS_ERR_INVALID_ARG;
  }
  if (rotations && (reinterpret_cast<uintptr_t>(rotations) & 15)) {
    set_error("rotations must be 16-byte aligned");
    return B200GS_ERR_INVALID_ARG;
  }
  return 0;
}

// one pinned word per host thread for the asynchronous read-back of D
static uint32_t* pinned_slot() {
  static thread_local uint32_t* slot = nullptr;
  if (!slot && cudaHostAlloc(reinterpret_cast<void**>(&slot), 64, cudaHostAllocPortable) != cudaSuccess) {
    set_error("cudaHostAlloc failed for the num_rendered slot");
    slot = nullptr;
  }
  return slot;
}

static char* grow(B200GSAlloc a, size_t bytes, const char* what) {
  if (!a.resize) { set_error("%s allocator is NULL", what); return nullptr; }
  char* p = a.resize(a.ctx, bytes);
  if (!p && bytes) set_error("%s allocator returned NULL for %zu bytes", what, bytes);
  else if (reinterpret_cast<uintptr_t>(p) & 127) { set_error("%s buffer must be 128-byte aligned", what); return nullptr; }
  return p;
}

}  // namespace b200gs

using namespace b200gs;

extern "C" {

const char* b200gs_last_error(void) { return g_err; }
int b200gs_version(void) { return B200GS_VERSION; }
int64_t b200gs_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

int b200gs_set_option(const char* name, int value) {
  g_err[0] = 0;
  if (name && !strcmp(name, "bin_shift")) {
    g_bin_shift_override.store(value < 0 ? -1 : (value > 5 ? 5 : value));
    return 0;
  }
  if (name && !strcmp(name, "sort")) {
    g_sort_mode.store(value < 0 ? 1 : (value > 2 ? 2 : value));
    return 0;
  }
  if (name && !strcmp(name, "sort_keys")) {
    g_key_mode.store(value == 64 ? 64 : 32);
    return 0;
  }
  if (name && !strcmp(name, "binning")) {
    g_binning_mode.store(value < 0 ? -1 : (value == 0 ? 0 : 1));   // -1: back to the default / environment
    return 0;
  }
  if (name && !strcmp(name, "render")) {
    g_render_mode.store(value < 0 ? -1 : (value == 0 ? 0 : 1));
    return 0;
  }
  if (name && !strcmp(name, "bwd_overlap")) {
    g_bwd_overlap.store(value == 1 ? 1 : 0);
    return 0;
  }
  if (name && !strcmp(name, "project")) {
    set_project_mode(value);
    return 0;
  }
  if (name && !strcmp(name, "slab")) {
    g_slab_mode.store(value == 1 ? 1 : 0);
    return 0;
  }
  if (name && !strcmp(name, "gather")) {
    set_gather_mode(value);
    return 0;
  }
  set_error("unknown option '%s'", name ? name : "(null)");
  return B200GS_ERR_INVALID_ARG;
}

int b200gs_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  return 0;
}
const char* b200gs_stage_name(int i) { return (i >= 0 && i < B200GS_NUM_STAGES) ? kStageNames[i] : ""; }
int b200gs_profile_read(float* stage_ms, int32_t* stage_calls, int reset) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < B200GS_NUM_STAGES; i++) { if (stage_ms) stage_ms[i] = 0.f; if (stage_calls) stage_calls[i] = 0; }
  for (auto& sp : g_spans) {
    if (cudaEventSynchronize(sp.b) != cudaSuccess) { set_error("profile_read: event sync failed"); return B200GS_ERR_CUDA; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, sp.a, sp.b);
    if (stage_ms) stage_ms[sp.stage] += ms;
    if (stage_calls) stage_calls[sp.stage] += 1;
  }
  if (reset) {
    for (auto& sp : g_spans) { g_free_events.push_back(sp.a); g_free_events.push_back(sp.b); }
    g_spans.clear();
  }
  return 0;
}

int b200gs_buffer_sizes(int32_t P, int32_t H, int32_t W, int64_t D, size_t* geom_bytes,
                        size_t* binning_bytes, size_t* img_bytes) {
  if (P < 0 || H <= 0 || W <= 0 || D < 0) { set_error("invalid sizes"); return B200GS_ERR_INVALID_ARG; }
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int bs = bin_shift_for(gx, gy);
  const int num_tiles = ((gx + (1 << bs) - 1) >> bs) * ((gy + (1 << bs) - 1) >> bs);
  if (geom_bytes) carve_geom(nullptr, P, geom_bytes);
  if (binning_bytes) carve_binning(nullptr, D, tile_bits_for(num_tiles), binning_bytes);
  if (img_bytes) carve_img(nullptr, H, W, img_bytes);
  return 0;
}

int b200gs_geom_layout(int32_t P, size_t offsets[B200GS_GEOM_FIELDS]) {
  g_err[0] = 0;
  if (P < 0 || !offsets) { set_error("geom_layout: invalid arguments"); return B200GS_ERR_INVALID_ARG; }
  GeomBuf g = carve_geom(reinterpret_cast<char*>(uintptr_t(128)), P, nullptr);   // fake 128-aligned base
  auto off = [](const void* p) { return (size_t)(reinterpret_cast<uintptr_t>(p) - 128); };
  offsets[0] = off(g.rec); offsets[1] = off(g.depth_key); offsets[2] = off(g.tiles);
  offsets[3] = off(g.offsets); offsets[4] = off(g.clamped);
  return 0;
}

int b200gs_forward(const B200GSParams* prm, const float* bg, const float* viewmatrix,
                   const float* projmatrix, const float* campos, const float* means3D,
                   const float* shs, const float* colors_precomp, const float* opacities,
                   const float* scales, const float* rotations, const float* cov3D_precomp,
                   float* out_color, int32_t* radii, B200GSAlloc geom, B200GSAlloc binning,
                   B200GSAlloc img, int32_t* num_rendered, void* stream) {
  g_err[0] = 0;
  int rc;
  if ((rc = check_params(prm))) return rc;
  if ((rc = check_inputs(prm, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp))) return rc;
  if (!bg || !viewmatrix || !projmatrix || !campos || !out_color || (prm->P && !radii) || !num_rendered) {
    set_error("bg/viewmatrix/projmatrix/campos/out_color/radii/num_rendered must not be NULL");
    return B200GS_ERR_INVALID_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int P = prm->P, H = prm->image_height, W = prm->image_width;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int bs = bin_shift_for(gx, gy, prm->flags);
  const int gbx = (gx + (1 << bs) - 1) >> bs, gby = (gy + (1 << bs) - 1) >> bs;
  const int num_tiles = gbx * gby;   // bins
  const int tile_bits = tile_bits_for(num_tiles);

  size_t geom_bytes, img_bytes;
  carve_geom(nullptr, P, &geom_bytes);
  carve_img(nullptr, H, W, &img_bytes);
  char* geom_p = grow(geom, geom_bytes, "geom");
  char* img_p = grow(img, img_bytes, "img");
  if (!geom_p || !img_p) return B200GS_ERR_ALLOC;
  GeomBuf gb = carve_geom(geom_p, P, nullptr);
  ImgBuf ib = carve_img(img_p, H, W, nullptr);

  ProjectArgs pa;
  pa.P = P; pa.M = prm->M; pa.W = W; pa.H = H; pa.gx = gx; pa.gy = gy; pa.bin_shift = bs;
  pa.sh_vec = (shs && (prm->M & 3) == 0 && (reinterpret_cast<uintptr_t>(shs) & 15) == 0) ? 1 : 0;
  pa.tanfovx = prm->tanfovx; pa.tanfovy = prm->tanfovy; pa.scale_modifier = prm->scale_modifier;
  pa.near_plane = prm->near_plane > 0.f ? prm->near_plane : 0.2f;
  pa.means = means3D; pa.scales = scales; pa.rots = rotations; pa.opac = opacities; pa.shs = shs;
  pa.colors_precomp = colors_precomp; pa.cov3d_precomp = cov3D_precomp;
  pa.view = viewmatrix; pa.proj = projmatrix; pa.campos = campos;
  pa.radii = radii; pa.rec = gb.rec; pa.depth_key = gb.depth_key; pa.tiles = gb.tiles;
  pa.clamped = gb.clamped;
  // bucketed binning: bin ids travel in 16 bits through the staged emission, and every bin needs >= 4 depth slices
  const bool bucketed = use_bucketed() && (uint32_t)num_tiles < 65535u && (uint32_t)num_tiles * 4u <= BUCKETS_MAX;
  // depth slices per bin: as many as the bucket tables hold, but no more buckets than ~4 per Gaussian (a 10 k-Gaussian
  // scene must not clear and scan half a million counters)
  const uint32_t bucket_budget = (uint32_t)std::min<uint64_t>(BUCKETS_MAX, std::max<uint64_t>(4ull * (uint64_t)P, 4096ull));
  int slices_log2 = 2;
  while (bucketed && slices_log2 < BUCKET_SLICES_LOG2_MAX && ((uint32_t)num_tiles << (slices_log2 + 1)) <= bucket_budget)
    slices_log2++;
  const uint32_t num_buckets = (uint32_t)num_tiles << slices_log2;
  uint32_t near_bits;
  {
    const float np = pa.near_plane;
    memcpy(&near_bits, &np, sizeof(uint32_t));
  }
  pa.gbx = gbx;
  pa.bucket_count = bucketed ? ib.bucket_count : nullptr;
  pa.slices_log2 = slices_log2;
  pa.slice_shift = 27 - slices_log2;      // 16 octaves of depth from the near plane over 2^slices_log2 slices
  pa.near_bits = near_bits;
  pa.pack_tiles = (bucketed && num_tiles <= 255) ? 1 : 0;
  // bucket counters and, directly in front of them, the look-back words of the bucket scan (the LAST num_tiles words
  // of the table, so that one memset of minimal length clears both)
  unsigned long long* const bin_pub = ib.bin_pub + (BUCKET_BINS_MAX - (uint32_t)(bucketed ? num_tiles : 0));
  if (bucketed && P > 0 &&
      (rc = check_cuda(cudaMemsetAsync(bin_pub, 0, sizeof(unsigned long long) * (size_t)num_tiles +
                                                     sizeof(uint32_t) * (size_t)num_buckets, st),
                       "clear bucket counters")))
    return rc;
  {
    StageTimer t(0, st);
    launch_project(pa, shs ? prm->sh_degree : -1, st);
  }
  if ((rc = debug_sync(prm, st, "project"))) return rc;
  BucketArgs ba;
  ba.P = P; ba.gx = gx; ba.gy = gy; ba.gbx = gbx; ba.bin_shift = bs;
  ba.id_bits = 1;
  while (ba.id_bits < 32 && (1ll << ba.id_bits) < (long long)P) ba.id_bits++;
  ba.slices_log2 = pa.slices_log2; ba.slice_shift = pa.slice_shift; ba.near_bits = near_bits;
  ba.num_bins = (uint32_t)num_tiles; ba.capacity = 0xFFFFFFFFu;
  ba.tiles = gb.tiles; ba.depth_key = gb.depth_key; ba.rec = gb.rec; ba.radii = radii;
  ba.bucket_count = ib.bucket_count; ba.bucket_base = ib.bucket_base; ba.bucket_cursor = ib.bucket_cursor;
  ba.bin_pub = bin_pub; ba.ranges = ib.ranges;
  ba.total = gb.counters + 1;
  ba.big_seg_count = gb.counters + 2; ba.total_windows = gb.counters + 3;
  ba.win_first = nullptr; ba.win_capacity = 0; ba.big_segs = nullptr;
  ba.seg = nullptr; ba.seg_alt = nullptr; ba.vals_sorted = nullptr; ba.slab = nullptr;
  const uint32_t* d_total = bucketed ? gb.counters + 1 : gb.offsets + (P > 0 ? P - 1 : 0);
  const int64_t hint_cap = (prm->pair_capacity_hint > 0 && prm->pair_capacity_hint < (int64_t)0x7fffffff)
                               ? prm->pair_capacity_hint : 0;
  bool bin_pub_clean = true;     // the look-back words are zero (cleared together with the counters)
  {
    StageTimer t(1, st);
    if (!bucketed) {
      if ((rc = scan_bin_counts(gb, P, st))) return rc;
    } else if (P > 0 && !hint_cap) {
      // exact path: D must reach the host before the binning buffer can be sized -- a first scan without the
      // window table; with a capacity hint the one scan inside run_binning_and_render produces everything
      launch_bucket_scan(ba, st);
      bin_pub_clean = false;
    }
  }
  if ((rc = debug_sync(prm, st, "scan"))) return rc;

  // ---- binning + compositing for a given pair capacity (D <= cap slots; [D,cap) are padding) ----
  auto run_binning_and_render = [&](uint32_t cap, const std::function<int()>* after_scan) -> int {
    size_t bin_bytes;
    const bool slab = bucketed && use_slab() && use_render4(gx * gy);
    carve_binning(nullptr, cap, tile_bits, &bin_bytes, slab);
    char* bin_p = grow(binning, bin_bytes, "binning");
    if (!bin_p && bin_bytes) return B200GS_ERR_ALLOC;
    BinBuf bb = carve_binning(bin_p, cap, tile_bits, nullptr, slab);
    int rc2;
    // global-sort pipeline: the ranges pass only writes bins that own pairs (the bucket scan writes every bin)
    if ((!bucketed || P == 0) &&
        (rc2 = check_cuda(cudaMemsetAsync(ib.ranges, 0, sizeof(uint2) * (size_t)num_tiles, st), "clear ranges")))
      return rc2;
    if (cap > 0 && bucketed) {
      // queue length of the large segments and window count; counters[1] = D is rewritten by the scan
      if ((rc2 = check_cuda(cudaMemsetAsync(gb.counters, 0, 4 * sizeof(uint32_t), st), "clear queue lengths"))) return rc2;
      if (!bin_pub_clean &&
          (rc2 = check_cuda(cudaMemsetAsync(bin_pub, 0, sizeof(unsigned long long) * (size_t)num_tiles, st),
                            "clear look-back words")))
        return rc2;
      BucketArgs b2 = ba;
      b2.capacity = cap; b2.seg = bb.keys; b2.seg_alt = bb.keys_sorted; b2.vals_sorted = bb.vals_sorted;
      b2.slab = slab ? bb.slab : nullptr;
      b2.win_first = bb.win_first; b2.win_capacity = cap / BUCKET_WINDOW + BUCKET_BINS_MAX + 2; b2.big_segs = bb.big_segs;
      {
        StageTimer t(1, st);
        launch_bucket_scan(b2, st);      // bucket starts, cursors, ranges clipped to `cap`, window table, D
        bin_pub_clean = false;
      }
      if ((rc2 = debug_sync(prm, st, "bucket scan"))) return rc2;
      if (after_scan && (rc2 = (*after_scan)())) return rc2;
      {
        StageTimer t(2, st);
        launch_bucket_emit(b2, st);
      }
      if ((rc2 = debug_sync(prm, st, "emit pairs"))) return rc2;
      {
        StageTimer t(3, st);
        launch_bucket_sort(b2, st);
      }
      if ((rc2 = debug_sync(prm, st, "bucket sort"))) return rc2;
    } else if (cap > 0) {
      if ((rc2 = check_cuda(cudaMemsetAsync(gb.counters, 0, 32 * sizeof(uint32_t), st), "clear counters"))) return rc2;
      const bool coop = use_coop_sort(cap);
      const bool key32 = !coop && num_tiles <= 255 && allow_key32();
      const int key_bits = key32 ? 24 + tile_bits : 32 + tile_bits;
      // the cooperative sort ping-pongs (passes) times; emit into whichever buffer makes the
      // result land in keys_sorted / vals_sorted
      const bool emit_into_sorted = coop && (((key_bits + 7) / 8) % 2 == 0);
      uint64_t* emit_keys = emit_into_sorted ? bb.keys_sorted : bb.keys;
      uint32_t* emit_vals = emit_into_sorted ? bb.vals_sorted : bb.vals;
      EmitArgs ea;
      ea.P = P; ea.gx = gx; ea.gy = gy; ea.gbx = gbx; ea.bin_shift = bs;
      ea.invalid_tile = (uint32_t)num_tiles; ea.capacity = cap;
      ea.key32 = key32 ? 1 : 0;
      {
        const float np = pa.near_plane;
        memcpy(&ea.near_bits, &np, sizeof(uint32_t));
      }
      ea.tiles = gb.tiles; ea.offsets = gb.offsets; ea.depth_key = gb.depth_key; ea.rec = gb.rec; ea.radii = radii;
      ea.keys = emit_keys; ea.vals = emit_vals;
      ea.big_queue = gb.big_queue; ea.big_count = gb.counters;
      {
        StageTimer t(2, st);
        launch_emit_pairs(ea, st);
      }
      if ((rc2 = debug_sync(prm, st, "emit pairs"))) return rc2;
      {
        StageTimer t(3, st);
        if (coop) {
          uint64_t* other_keys = emit_into_sorted ? bb.keys : bb.keys_sorted;
          uint32_t* other_vals = emit_into_sorted ? bb.vals : bb.vals_sorted;
          if ((rc2 = coop_sort_pairs(emit_keys, emit_vals, other_keys, other_vals, bb.coop_hist, cap, key_bits, st)))
            return rc2;
        } else if ((rc2 = key32 ? sort_pairs32(bb, cap, key_bits, st) : sort_pairs(bb, cap, key_bits, st))) {
          return rc2;
        }
      }
      if ((rc2 = debug_sync(prm, st, "pair sort"))) return rc2;
      {
        StageTimer t(4, st);
        if (key32) {
          Ranges32Args ga;
          ga.D = cap; ga.num_tiles = (uint32_t)num_tiles; ga.id_bits = ba.id_bits;
          ga.keys_sorted = reinterpret_cast<const uint32_t*>(bb.keys_sorted); ga.vals_sorted = bb.vals_sorted;
          ga.depth_key = gb.depth_key; ga.ranges = ib.ranges;
          // after the sort the unsorted value array and both key arrays are free: run queue and scratch
          ga.run_queue = reinterpret_cast<uint2*>(bb.vals); ga.run_capacity = cap / 2; ga.run_count = gb.counters + 3;
          ga.scratch_a = bb.keys; ga.scratch_b = bb.keys_sorted;
          launch_tile_ranges32(ga, st);
        } else {
          RangesArgs ga;
          ga.D = cap; ga.num_tiles = (uint32_t)num_tiles; ga.keys_sorted = bb.keys_sorted; ga.ranges = ib.ranges;
          launch_tile_ranges(ga, st);
        }
      }
      if ((rc2 = debug_sync(prm, st, "tile ranges"))) return rc2;
    }
    RenderArgs ra;
    ra.W = W; ra.H = H; ra.gbx = gbx; ra.bin_shift = bs; ra.ranges = ib.ranges; ra.point_list = bb.vals_sorted;
    ra.rec = gb.rec; ra.slab = (slab && cap > 0) ? bb.slab : nullptr; ra.bg = bg; ra.out_color = out_color;
    const bool fwd_only = (prm->flags & B200GS_FORWARD_ONLY) != 0;     // nothing is kept for the adjoint
    ra.pix = fwd_only ? nullptr : ib.pix; ra.n_contrib = fwd_only ? nullptr : ib.n_contrib;
    ra.vec4 = ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(out_color) & 15) == 0) ? 1 : 0;
    ra.out_rgb8 = (prm->flags & B200GS_OUT_RGB8) ? reinterpret_cast<uint8_t*>(out_color) : nullptr;
    {
      StageTimer t(5, st);
      if (use_render4(gx * gy)) launch_render4(ra, st);
      else launch_render(ra, st);
    }
    return debug_sync(prm, st, "render");
  };

  uint32_t D = 0;
  const int64_t hint = hint_cap;
  if (P > 0 && hint > 0 && (prm->flags & B200GS_DEFER_PAIR_CHECK)) {
    // Deferred check: D goes straight to the caller's pinned word; the host never waits here.
    const std::function<int()> queue_copy = [&]() -> int {
      return check_cuda(cudaMemcpyAsync(num_rendered, d_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, st),
                        "queue num_rendered copy (deferred)");
    };
    if (!bucketed && (rc = queue_copy())) return rc;      // global sort: D is known since the scan over Gaussians
    return run_binning_and_render((uint32_t)hint, bucketed ? &queue_copy : nullptr);
  }
  if (P > 0 && hint > 0) {
    // Speculative path: D stays on the device.  Its copy to pinned host memory is queued, the rest
    // of the frame is launched for `hint` pair slots, and only then does the host wait for the copy
    // (an event early in the stream) -- the GPU never idles while the host learns D.
    uint32_t* host_d = pinned_slot();
    if (!host_d) return B200GS_ERR_ALLOC;
    cudaEvent_t ev;
    if ((rc = check_cuda(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "event create"))) return rc;
    const std::function<int()> queue_copy = [&]() -> int {
      int r = check_cuda(cudaMemcpyAsync(host_d, d_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, st),
                         "queue num_rendered copy");
      if (!r) r = check_cuda(cudaEventRecord(ev, st), "event record");
      return r;
    };
    if (!bucketed) rc = queue_copy();
    if (!rc) rc = run_binning_and_render((uint32_t)hint, bucketed ? &queue_copy : nullptr);
    if (!rc) rc = check_cuda(cudaEventSynchronize(ev), "wait num_rendered");
    cudaEventDestroy(ev);
    if (rc) return rc;
    D = *host_d;
    *num_rendered = (int32_t)D;
    if ((int64_t)D > hint) return run_binning_and_render(D, nullptr);   // hint too small: redo exactly
    return 0;
  }
  if (P > 0) {
    if ((rc = check_cuda(cudaMemcpyAsync(&D, d_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, st),
                         "read num_rendered")))
      return rc;
    if ((rc = check_cuda(cudaStreamSynchronize(st), "sync num_rendered"))) return rc;
  }
  *num_rendered = (int32_t)D;
  return run_binning_and_render(D, nullptr);
}

int b200gs_backward(const B200GSParams* prm, const float* bg, const float* viewmatrix,
                    const float* projmatrix, const float* campos, const float* means3D,
                    const float* shs, const float* colors_precomp, const float* opacities,
                    const float* scales, const float* rotations, const float* cov3D_precomp,
                    const int32_t* radii, const char* geom, const char* binning, const char* img,
                    int32_t num_rendered, const float* dL_dout_color, float* dL_dmeans3D,
                    float* dL_dmeans2D, float* dL_dshs, float* dL_dcolors_precomp,
                    float* dL_dopacities, float* dL_dscales, float* dL_drotations,
                    float* dL_dcov3D, B200GSAlloc scratch, void* stream) {
  g_err[0] = 0;
  int rc;
  if ((rc = check_params(prm))) return rc;
  if ((rc = check_inputs(prm, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp))) return rc;
  const int P = prm->P, H = prm->image_height, W = prm->image_width;
  if (prm->flags & B200GS_FORWARD_ONLY) {
    set_error("backward: the forward call was B200GS_FORWARD_ONLY, the per-pixel state of the adjoint was not written");
    return B200GS_ERR_INVALID_ARG;
  }
  if (P == 0) return 0;
  if (!bg || !viewmatrix || !projmatrix || !campos || !radii || !geom || !img || !dL_dout_color ||
      !dL_dmeans3D || !dL_dmeans2D || !dL_dopacities || (num_rendered > 0 && !binning)) {
    set_error("backward: required pointer is NULL");
    return B200GS_ERR_INVALID_ARG;
  }
  if ((shs && !dL_dshs) || (colors_precomp && !dL_dcolors_precomp) || (scales && (!dL_dscales || !dL_drotations)) ||
      (cov3D_precomp && !dL_dcov3D)) {
    set_error("backward: gradient output missing for a provided input");
    return B200GS_ERR_INVALID_ARG;
  }
  if (dL_drotations && (reinterpret_cast<uintptr_t>(dL_drotations) & 15)) {
    set_error("dL_drotations must be 16-byte aligned");
    return B200GS_ERR_INVALID_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int bs = bin_shift_for(gx, gy, prm->flags);    // must be the flags of the forward call
  const int gbx = (gx + (1 << bs) - 1) >> bs, gby = (gy + (1 << bs) - 1) >> bs;
  const int tile_bits = tile_bits_for(gbx * gby);
  GeomBuf gb = carve_geom(const_cast<char*>(geom), P, nullptr);
  // the record slab exists iff the forward call ran the bucketed binning with option "slab" and the four-pixel kernels
  const bool slab = use_slab() && use_render4(gx * gy) && use_bucketed() && (uint32_t)(gbx * gby) < 65535u &&
                    (uint32_t)(gbx * gby) * 4u <= BUCKETS_MAX;
  BinBuf bb = carve_binning(const_cast<char*>(binning), num_rendered, tile_bits, nullptr, slab);
  ImgBuf ib = carve_img(const_cast<char*>(img), H, W, nullptr);

  const size_t g2_bytes = sizeof(float) * GRAD2D_STRIDE * (size_t)P;
  char* g2_p = grow(scratch, g2_bytes, "scratch");
  if (!g2_p) return B200GS_ERR_ALLOC;
  float* grad2d = reinterpret_cast<float*>(g2_p);
  if ((rc = check_cuda(cudaMemsetAsync(grad2d, 0, g2_bytes, st), "clear grad2d"))) return rc;

  // fork: zero-fill of the outputs on a side stream, overlapped with the compositing adjoint
  SideStream* side = (use_bwd_overlap() && num_rendered > 0) ? side_stream_for(st) : nullptr;
  if (side) {
    rc = check_cuda(cudaEventRecord(side->fork, st), "backward: fork");
    if (!rc) rc = check_cuda(cudaStreamWaitEvent(side->side, side->fork, 0), "backward: fork wait");
    auto zero = [&](void* p, size_t bytes) {
      if (!rc && p && bytes) rc = check_cuda(cudaMemsetAsync(p, 0, bytes, side->side), "backward: zero-fill");
    };
    zero(dL_dmeans3D, sizeof(float) * 3 * (size_t)P);
    zero(dL_dmeans2D, sizeof(float) * 3 * (size_t)P);
    zero(dL_dopacities, sizeof(float) * (size_t)P);
    if (shs) zero(dL_dshs, sizeof(float) * 3 * (size_t)prm->M * (size_t)P);
    if (colors_precomp) zero(dL_dcolors_precomp, sizeof(float) * 3 * (size_t)P);
    if (scales) { zero(dL_dscales, sizeof(float) * 3 * (size_t)P); zero(dL_drotations, sizeof(float) * 4 * (size_t)P); }
    if (cov3D_precomp) zero(dL_dcov3D, sizeof(float) * 6 * (size_t)P);
    if (!rc) rc = check_cuda(cudaEventRecord(side->join, side->side), "backward: join record");
    if (rc) return rc;
  }

  if (num_rendered > 0) {
    RenderBwdArgs ra;
    ra.W = W; ra.H = H; ra.gbx = gbx; ra.bin_shift = bs; ra.ranges = ib.ranges; ra.point_list = bb.vals_sorted;
    ra.rec = gb.rec; ra.slab = slab ? bb.slab : nullptr; ra.bg = bg; ra.pix = ib.pix;
    ra.n_contrib = ib.n_contrib; ra.dL_dpix = dL_dout_color; ra.grad2d = grad2d;
    ra.vec4 = ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(dL_dout_color) & 15) == 0) ? 1 : 0;
    {
      StageTimer t(6, st);
      if (use_render4(gx * gy)) launch_render_bwd4(ra, st);
      else launch_render_bwd(ra, st);
    }
    if ((rc = debug_sync(prm, st, "render backward"))) return rc;
  }

  if (side && (rc = check_cuda(cudaStreamWaitEvent(st, side->join, 0), "backward: join"))) return rc;
  ProjectBwdArgs pa;
  pa.P = P; pa.M = prm->M; pa.W = W; pa.H = H;
  pa.active_only = side ? 1 : 0;
  static const bool no_slab = getenv("B200GS_NO_SLAB") != nullptr;     // read once per process
  pa.slab = no_slab ? -1 : 0;
  pa.sh_vec = (shs && (prm->M & 3) == 0 && (reinterpret_cast<uintptr_t>(shs) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(dL_dshs) & 15) == 0) ? 1 : 0;
  pa.tanfovx = prm->tanfovx; pa.tanfovy = prm->tanfovy; pa.scale_modifier = prm->scale_modifier;
  pa.means = means3D; pa.scales = scales; pa.rots = rotations; pa.shs = shs; pa.cov3d_precomp = cov3D_precomp;
  pa.view = viewmatrix; pa.proj = projmatrix; pa.campos = campos;
  pa.radii = radii; pa.tiles = gb.tiles; pa.clamped = gb.clamped; pa.rec = gb.rec; pa.grad2d = grad2d;
  pa.dL_dmeans = dL_dmeans3D; pa.dL_dmeans2D = dL_dmeans2D; pa.dL_dshs = dL_dshs; pa.dL_dcolors = dL_dcolors_precomp;
  pa.dL_dopac = dL_dopacities; pa.dL_dscales = dL_dscales; pa.dL_drots = dL_drotations; pa.dL_dcov3D = dL_dcov3D;
  {
    StageTimer t(7, st);
    launch_project_bwd(pa, shs ? prm->sh_degree : -1, st);
  }
  return debug_sync(prm, st, "project backward");
}

static int check_loss_args(const float* a, const float* b, int64_t n, const void* o) {
  if (n < 0 || (n > 0 && (!a || !b || !o)) || (reinterpret_cast<uintptr_t>(a) & 15) ||
      (reinterpret_cast<uintptr_t>(b) & 15)) {
    set_error("photometric_loss: need 16-byte aligned non-NULL buffers");
    return B200GS_ERR_INVALID_ARG;
  }
  return 0;
}

int b200gs_photometric_loss(const float* a, const float* b, int64_t n, float w_l2, float w_l1, float* out_sum,
                            void* stream) {
  g_err[0] = 0;
  if (int rc = check_loss_args(a, b, n, out_sum)) return rc;
  launch_photometric_loss(a, b, (size_t)n, w_l2, w_l1, out_sum, static_cast<cudaStream_t>(stream));
  return check_cuda(cudaGetLastError(), "photometric_loss");
}

int b200gs_photometric_loss_backward(const float* a, const float* b, int64_t n, float w_l2, float w_l1,
                                     float scale, const float* upstream, float* dL_da, void* stream) {
  g_err[0] = 0;
  if (int rc = check_loss_args(a, b, n, dL_da)) return rc;
  if (!upstream || (reinterpret_cast<uintptr_t>(dL_da) & 15)) {
    set_error("photometric_loss_backward: upstream NULL or dL_da misaligned");
    return B200GS_ERR_INVALID_ARG;
  }
  launch_photometric_loss_bwd(a, b, (size_t)n, w_l2, w_l1, scale, upstream, dL_da, static_cast<cudaStream_t>(stream));
  return check_cuda(cudaGetLastError(), "photometric_loss_backward");
}

int b200gs_ssim_forward(const float* img1, const float* img2, int32_t C, int32_t H, int32_t W, float* maps,
                        float* out_sum, void* stream) {
  g_err[0] = 0;
  if (!img1 || !img2 || !out_sum || C <= 0 || C > 65535 || H <= 0 || W <= 0) {
    set_error("ssim_forward: invalid arguments");
    return B200GS_ERR_INVALID_ARG;
  }
  launch_ssim_fwd(img1, img2, C, H, W, maps, out_sum, static_cast<cudaStream_t>(stream));
  return check_cuda(cudaGetLastError(), "ssim_forward");
}

int b200gs_ssim_backward(const float* img1, const float* img2, const float* maps, int32_t C, int32_t H, int32_t W,
                         float scale, const float* upstream, float* dL_dimg1, void* stream) {
  g_err[0] = 0;
  if (!img1 || !img2 || !maps || !upstream || !dL_dimg1 || C <= 0 || C > 65535 || H <= 0 || W <= 0) {
    set_error("ssim_backward: invalid arguments");
    return B200GS_ERR_INVALID_ARG;
  }
  launch_ssim_bwd(img1, img2, maps, C, H, W, scale, upstream, dL_dimg1, static_cast<cudaStream_t>(stream));
  return check_cuda(cudaGetLastError(), "ssim_backward");
}

int b200gs_adam_step(const B200GSAdamGroup* groups, int32_t num_groups, float beta1, float beta2, float eps,
                     int32_t step, void* stream) {
  g_err[0] = 0;
  if (!groups || num_groups < 0 || num_groups > B200GS_ADAM_MAX_GROUPS || step < 1 || !(beta1 >= 0.f && beta1 < 1.f) ||
      !(beta2 >= 0.f && beta2 < 1.f)) {
    set_error("adam_step: need 0..%d groups, step >= 1, betas in [0,1)", B200GS_ADAM_MAX_GROUPS);
    return B200GS_ERR_INVALID_ARG;
  }
  for (int i = 0; i < num_groups; i++) {
    const B200GSAdamGroup& g = groups[i];
    if (g.n < 0 || (g.n > 0 && (!g.param || !g.grad || !g.exp_avg || !g.exp_avg_sq))) {
      set_error("adam_step: group %d has a NULL tensor or negative size", i);
      return B200GS_ERR_INVALID_ARG;
    }
  }
  if (int rc = launch_adam(groups, num_groups, beta1, beta2, eps, step, static_cast<cudaStream_t>(stream))) return rc;
  return check_cuda(cudaGetLastError(), "adam_step");
}

int b200gs_extract_alpha(const char* img, int32_t H, int32_t W, float* out_alpha, void* stream) {
  g_err[0] = 0;
  if (!img || !out_alpha || H <= 0 || W <= 0) {
    set_error("extract_alpha: invalid arguments");
    return B200GS_ERR_INVALID_ARG;
  }
  ImgBuf ib = carve_img(const_cast<char*>(img), H, W, nullptr);
  launch_extract_alpha(ib.pix, (size_t)H * W, out_alpha, static_cast<cudaStream_t>(stream));
  return check_cuda(cudaGetLastError(), "extract_alpha");
}

int b200gs_export_rgb8(const float* color, int32_t H, int32_t W, uint8_t* out_hwc, void* stream) {
  g_err[0] = 0;
  if (!color || !out_hwc || H <= 0 || W <= 0) {
    set_error("export_rgb8: invalid arguments");
    return B200GS_ERR_INVALID_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(color) & 15) || (reinterpret_cast<uintptr_t>(out_hwc) & 3)) {
    set_error("export_rgb8: color must be 16-byte aligned, out 4-byte aligned");
    return B200GS_ERR_INVALID_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  launch_export_rgb8(color, H, W, out_hwc, st);
  return check_cuda(cudaGetLastError(), "export_rgb8");
}

int b200gs_ply_activate(int32_t P, const float* vertices, const B200GSPlyLayout* layout, float* means3D,
                        float* shs, float* opacities, float* scales, float* rotations, void* stream) {
  g_err[0] = 0;
  if (P < 0 || !layout || (P > 0 && (!vertices || !means3D || !shs || !opacities || !scales || !rotations))) {
    set_error("ply_activate: invalid arguments");
    return B200GS_ERR_INVALID_ARG;
  }
  const B200GSPlyLayout& L = *layout;
  const int need = 3 + 3 + 3 * L.n_rest + 1 + 3 + 4;
  if (L.stride < need || L.n_rest < 0 || L.off_xyz < 0 || L.off_fdc < 0 || L.off_frest < 0 || L.off_opacity < 0 ||
      L.off_scale < 0 || L.off_rot < 0 || L.off_rot + 4 > L.stride || L.off_frest + 3 * L.n_rest > L.stride) {
    set_error("ply_activate: inconsistent layout (stride %d)", L.stride);
    return B200GS_ERR_INVALID_ARG;
  }
  int rc = launch_ply_activate(P, vertices, L, means3D, shs, opacities, scales, rotations,
                               static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  return check_cuda(cudaGetLastError(), "ply_activate");
}

int b200gs_transform_gaussians(int32_t n, const float* means_in, const float* rots_in,
                               const int32_t* link_ids, const float* link_transforms,
                               const float* link_quats, int32_t L, float* means_out, float* rots_out,
                               void* stream) {
  g_err[0] = 0;
  if (n < 0 || L <= 0 || (n > 0 && (!means_in || !rots_in || !link_transforms || !link_quats || !means_out || !rots_out))) {
    set_error("transform_gaussians: invalid arguments");
    return B200GS_ERR_INVALID_ARG;
  }
  launch_transform_gaussians(n, means_in, rots_in, link_ids, link_transforms, link_quats, L, means_out, rots_out,
                             static_cast<cudaStream_t>(stream));
  return check_cuda(cudaGetLastError(), "transform_gaussians");
}

int b200gs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                        const float* projmatrix, uint8_t* present, void* stream) {
  g_err[0] = 0;
  (void)projmatrix;
  if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) {
    set_error("mark_visible: invalid arguments");
    return B200GS_ERR_INVALID_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  launch_mark_visible(P, means3D, viewmatrix, 0.2f, present, st);
  return check_cuda(cudaGetLastError(), "mark_visible");
}

}  // extern "C"

// kernels.cuh -- argument blocks and host launchers shared between the translation units.
#pragma once
#include "common.cuh"

namespace b200gs {

// bucketed binning (bucket.cu): (bin, depth slice) buckets, at most BUCKET_BINS_MAX bins of 4..8192 slices each
constexpr uint32_t BUCKETS_MAX = 512 * 1024;
constexpr uint32_t BUCKET_BINS_MAX = 65536;   // table size; bin ids travel in 16 bits through the staged emission, so
                                              // the bucketed path serves images of up to 65534 bins
constexpr int BUCKET_SLICES_LOG2_MAX = 13;
#ifndef BUCKET_WINDOW_N
#define BUCKET_WINDOW_N 192
#endif
#ifndef WARP_SORT_MAX_N
#define WARP_SORT_MAX_N 512
#endif
constexpr uint32_t BUCKET_WINDOW = BUCKET_WINDOW_N;   // pairs per sorter warp (cut at the next bucket boundary)
constexpr uint32_t WARP_SORT_MAX = WARP_SORT_MAX_N;   // largest window one warp sorts in registers; larger ones take the one-CTA path

struct ProjectArgs {
  int P, M, W, H, gx, gy, sh_vec, bin_shift;
  int gbx;               // bin grid width
  uint32_t* bucket_count;  // bucketed binning: per-(bin, depth slice) pair counters, else nullptr
  int slices_log2, slice_shift;   // slice = min((depth bits - near_bits) >> slice_shift, 2^slices_log2 - 1)
  uint32_t near_bits;
  int pack_tiles;          // bucketed binning with <= 255 bins: tiles[] carries small footprints packed (spans.cuh)
  float tanfovx, tanfovy, scale_modifier, near_plane;
  const float *means, *scales, *rots, *opac, *shs, *colors_precomp, *cov3d_precomp;
  const float *view, *proj, *campos;
  int32_t* radii;
  float4* rec;
  uint32_t *depth_key, *tiles;
  uint8_t* clamped;
};

struct EmitArgs {
  int P, gx, gy;     // gx, gy: 16-px tile grid (reference rect)
  int gbx, bin_shift; // bin grid width, log2(bin edge / 16)
  uint32_t invalid_tile;
  int key32;            // 32-bit pair keys (bin << 24 | quantised depth), see project.cu:put_key
  uint32_t near_bits;   // IEEE bits of the near plane (origin of the depth quantisation)
  uint32_t capacity;   // number of pair slots in keys/vals; slots [D, capacity) are padded
  const uint32_t *tiles, *offsets, *depth_key;
  const float4* rec;
  const int32_t* radii;
  uint64_t* keys;       // (bin << 32) | depth bits
  uint32_t* vals;       // Gaussian id
  uint32_t* big_queue;  // [P] ids of large-footprint Gaussians
  uint32_t* big_count;  // [1] zeroed before the emission stage
};

// bucketed binning (bucket.cu)
struct BucketArgs {
  int P, gx, gy, gbx, bin_shift;
  int id_bits;              // bits of the largest Gaussian index
  int slices_log2, slice_shift;
  uint32_t near_bits;       // IEEE bits of the near plane (origin of the depth slices)
  uint32_t num_bins, capacity;
  const uint32_t *tiles, *depth_key;
  const float4* rec;
  const int32_t* radii;
  uint32_t* bucket_count;             // [num_bins << slices_log2] pairs per bucket (written by k_project)
  uint32_t* bucket_base;              // [(num_bins << slices_log2) + 1] exclusive scan
  uint32_t* bucket_cursor;            // [num_bins << slices_log2] append cursors (start at bucket_base)
  uint2* ranges;                      // [num_bins] per-bin [start,end) into vals_sorted
  uint32_t* total;                    // D
  unsigned long long* bin_pub;        // [num_bins] look-back words of the scan (valid | windows << 32 | pairs), zeroed
  uint32_t* win_first;                // [capacity / BUCKET_WINDOW + num_bins + 2] first bucket of every sort window
  uint32_t win_capacity;              // entries of win_first
  uint32_t* total_windows;
  uint2* big_segs;                    // [capacity / 512 + 2] (start, length) of segments too large for the warp sort
  uint32_t* big_seg_count;
  uint64_t *seg, *seg_alt;            // [capacity] ((depth bits - near bits) << 32 | id), ping-pong
  uint32_t* vals_sorted;              // [capacity] Gaussian ids in (bin, depth, index) order
  float4* slab;                       // [capacity * 3] records in list order (option "slab") or nullptr
};

struct RangesArgs {
  int64_t D;           // sorted pair slots (incl. padding of the speculative capacity)
  uint32_t num_tiles;
  const uint64_t* keys_sorted;
  uint2* ranges;
};

struct Ranges32Args {
  int64_t D;
  uint32_t num_tiles;
  int id_bits;
  const uint32_t* keys_sorted;   // (bin << 24 | q24), sorted
  uint32_t* vals_sorted;         // Gaussian ids; runs of equal keys are re-ordered in place
  const uint32_t* depth_key;     // [P] exact depth bits
  uint2* ranges;
  uint2* run_queue;              // long runs (start, length)
  uint32_t run_capacity;
  uint32_t* run_count;
  uint64_t *scratch_a, *scratch_b;   // [D] each, free after the sort
};

struct RenderArgs {
  int W, H, gbx, bin_shift;
  int vec4;                    // W % 4 == 0 and out_color 16-byte aligned: 128-bit pixel stores
  const uint2* ranges;
  const uint32_t* point_list;  // Gaussian ids in (tile, depth) order
  const float4* rec;           // [P] projected-splat records
  const float4* slab;          // [D] records in list order (written by the bucket sort) or nullptr: TMA-fed ring
  const float* bg;
  float* out_color;     // [3][H][W]
  uint8_t* out_rgb8;    // not null: the frame leaves as [H][W][3] 8-bit RGB instead (B200GS_OUT_RGB8), out_color unused
  float4* pix;          // [H*W]
  uint32_t* n_contrib;  // [H*W]
};

// fp32 colour -> 8-bit channel, the conversion of k_export_rgb8 (clamp to [0,1], round to nearest)
__device__ __forceinline__ uint32_t rgb8_of(float v) { return (uint32_t)__float2int_rn(__saturatef(v) * 255.f); }

struct RenderBwdArgs {
  int W, H, gbx, bin_shift;
  int vec4;                    // W % 4 == 0 and dL_dpix 16-byte aligned: 128-bit pixel loads
  const uint2* ranges;
  const uint32_t* point_list;
  const float4* rec;
  const float4* slab;          // see RenderArgs
  const float* bg;
  const float4* pix;
  const uint32_t* n_contrib;
  const float* dL_dpix;  // [3][H][W]
  float* grad2d;         // [P][12], zero-initialised
};

struct ProjectBwdArgs {
  int P, M, W, H, sh_vec;
  int slab;   // in: -1 disables the shared-memory SH-gradient slab; set by the launcher
  int active_only;   // outputs were zero-filled beforehand: write only the rows of Gaussians that reached the image
  float tanfovx, tanfovy, scale_modifier;
  const float *means, *scales, *rots, *shs, *cov3d_precomp;
  const float *view, *proj, *campos;
  const int32_t* radii;
  const uint32_t* tiles;
  const uint8_t* clamped;
  const float4* rec;
  const float* grad2d;
  float *dL_dmeans, *dL_dmeans2D, *dL_dshs, *dL_dcolors, *dL_dopac, *dL_dscales, *dL_drots, *dL_dcov3D;
};

void launch_project(const ProjectArgs& a, int deg, cudaStream_t st);
void launch_emit_pairs(const EmitArgs& a, cudaStream_t st);
void launch_tile_ranges(const RangesArgs& a, cudaStream_t st);
void launch_tile_ranges32(const Ranges32Args& a, cudaStream_t st);
void launch_mark_visible(int P, const float* means, const float* view, float near_plane, uint8_t* present,
                         cudaStream_t st);
void launch_extract_alpha(const float4* pix, size_t npx, float* out, cudaStream_t st);
void launch_project_bwd(const ProjectBwdArgs& a, int deg, cudaStream_t st);
void launch_render(const RenderArgs& a, cudaStream_t st);
void launch_render_bwd(const RenderBwdArgs& a, cudaStream_t st);
void set_project_mode(int mode);   // 0: one thread per Gaussian (default), 1: dense-warp projection kernel
void set_gather_mode(int mode);  // 0: TMA bulk copy per record, 1: LDGSTS
// render4.cu: four pixels per thread (default); render.cu keeps the one-pixel-per-thread kernels
void launch_render4(const RenderArgs& a, cudaStream_t st);
void launch_render_bwd4(const RenderBwdArgs& a, cudaStream_t st);
void launch_export_rgb8(const float* color, int H, int W, uint8_t* out, cudaStream_t st);

// coopsort.cu: single-launch cooperative radix sort for small pair lists
constexpr int COOP_SORT_MAX_BLOCKS = 256;
constexpr int64_t COOP_SORT_MAX_ITEMS = 3 * 1000 * 1000;
size_t coop_sort_hist_bytes();
int coop_sort_pairs(uint64_t* keys0, uint32_t* vals0, uint64_t* keys1, uint32_t* vals1, uint32_t* hist, uint32_t n,
                    int key_bits, cudaStream_t st);

// scene.cu
int launch_ply_activate(int P, const float* v, const B200GSPlyLayout& L, float* means, float* shs, float* opac,
                        float* scales, float* rots, cudaStream_t st);
void launch_transform_gaussians(int n, const float* means_in, const float* rots_in, const int32_t* link_ids,
                                const float* T, const float* Q, int L, float* means_out, float* rots_out,
                                cudaStream_t st);

void launch_photometric_loss(const float* a, const float* b, size_t n, float w_l2, float w_l1, float* out,
                             cudaStream_t st);
void launch_photometric_loss_bwd(const float* a, const float* b, size_t n, float w_l2, float w_l1, float scale,
                                 const float* upstream, float* g, cudaStream_t st);

// bucket.cu: bucketed binning (per-bin lists without a global sort)
void launch_bucket_scan(const BucketArgs& a, cudaStream_t st);
void launch_bucket_emit(const BucketArgs& a, cudaStream_t st);
void launch_bucket_sort(const BucketArgs& a, cudaStream_t st);


// train.cu: fused SSIM and multi-tensor Adam
void launch_ssim_fwd(const float* img1, const float* img2, int C, int H, int W, float* maps, float* out_sum,
                     cudaStream_t st);
void launch_ssim_bwd(const float* img1, const float* img2, const float* maps, int C, int H, int W, float scale,
                     const float* upstream, float* dL_dimg1, cudaStream_t st);
int launch_adam(const B200GSAdamGroup* groups, int num_groups, float beta1, float beta2, float eps, int step,
                cudaStream_t st);

// binning.cu (CUB): temp-storage sizing, scan and pair sort
size_t scan_temp_bytes(int P);
size_t pair_sort_temp_bytes(int64_t D, int key_bits);
int scan_bin_counts(const GeomBuf& g, int P, cudaStream_t st);
int sort_pairs(const BinBuf& b, int64_t D, int key_bits, cudaStream_t st);
int sort_pairs32(const BinBuf& b, int64_t D, int key_bits, cudaStream_t st);   // 32-bit keys in keys / keys_sorted

}  // namespace b200gs

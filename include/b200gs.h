/*
 * b200gs.h -- C ABI of the B200-native 3D Gaussian Splatting rasterizer (libb200gs.so).
 *
 * This is the drop-in boundary for the hot path named by BASELINE.json:north_star.  The reference
 * repository (Maxwell-Zhao/RoboSimGS) contains no rasterizer of its own: it delegates background
 * 3DGS reconstruction to Nerfstudio (/root/reference/README.md:75) and its per-frame datagen
 * renderer is unreleased (/root/reference/README.md:29,85).  The interface replaced here is
 * therefore the one north_star names -- the pybind module `_C` of the public
 * diff-gaussian-rasterization package (un-vendored third-party code):
 *
 *     _C.rasterize_gaussians(...)           ->  b200gs_forward()
 *     _C.rasterize_gaussians_backward(...)  ->  b200gs_backward()
 *     _C.mark_visible(...)                  ->  b200gs_mark_visible()
 *
 * Plain pointers and sizes only; no torch types.  All `const float*` / `float*` data arguments are
 * DEVICE pointers (contiguous fp32, row-major), exactly the buffers the Python operator surface
 * (GaussianRasterizer.forward) already holds.  `stream` is a cudaStream_t
 * passed as void* (NULL = legacy default stream).  Process-wide state: the tuning knobs of b200gs_set_option
 * (never change results), the optional stage timers of b200gs_profile_* and the launch counter; the error
 * string is thread-local.  Calls are re-entrant per stream.
 *
 * Scratch memory is caller-owned: the library asks the caller to size
 * byte buffers through a callback and carves 128-byte aligned sub-arrays inside them.  The three
 * forward buffers must be handed back unchanged to b200gs_backward().
 */
#ifndef B200GS_H
#define B200GS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200GS_VERSION 200

/* Error codes (0 = success).  b200gs_last_error() returns the message for the calling thread. */
#define B200GS_OK 0
#define B200GS_ERR_INVALID_ARG (-1)
#define B200GS_ERR_ALLOC (-2)
#define B200GS_ERR_CUDA (-3)

/*
 * Per-call camera/config -- the scalar fields of GaussianRasterizationSettings (SURVEY.md 8(a)
 * row a1).  The small tensors of the settings tuple (bg, viewmatrix, projmatrix, campos) stay on
 * the device and are passed as pointers to b200gs_forward/backward, so a call never forces a
 * device->host copy of the camera.
 */
#define B200GS_DEFER_PAIR_CHECK 1
#define B200GS_FORWARD_ONLY 2
#define B200GS_OUT_RGB8 4
#define B200GS_BIN_SHIFT_HINT(s) (((s) + 1) << 8)

typedef struct B200GSParams {
  int32_t P;              /* number of Gaussians */
  int32_t sh_degree;      /* active SH degree, 0..3 */
  int32_t M;              /* SH coefficients stored per Gaussian per channel (shs is [P][M][3]) */
  int32_t image_height;
  int32_t image_width;
  float tanfovx;
  float tanfovy;
  float scale_modifier;
  int32_t prefiltered;    /* accepted for signature parity; culled Gaussians are simply skipped */
  int32_t debug;          /* !=0: synchronise and check for errors after every kernel */
  float near_plane;       /* near cull on view-space z; <= 0 selects the public default 0.2 */
  int32_t flags;          /* bit 0 (B200GS_DEFER_PAIR_CHECK): with pair_capacity_hint > 0, do not wait for D at
                             all -- *num_rendered must then point to PINNED host memory, receives D
                             asynchronously on `stream`, and the CALLER checks D <= pair_capacity_hint once the
                             stream has passed the call (a frame that fails the check is incomplete and must be
                             rendered again).  Lets a sweep keep many independent frames in flight.
                             bit 1 (B200GS_FORWARD_ONLY): no backward pass will follow this frame -- the per-pixel state
                             the adjoint replays from (accumulated colour, final transmittance, list position: 20
                             bytes per pixel of `img`) is not written; out_color and radii are unchanged.
                             bit 2 (B200GS_OUT_RGB8): `out_color` points to H*W*3 BYTES and receives the frame as
                             [H][W][3] 8-bit RGB (clamp to [0,1], round to nearest -- what b200gs_export_rgb8 makes of the
                             fp32 frame, bit for bit) straight from the compositing kernel: a datagen sweep neither
                             writes nor re-reads the 12-byte-per-pixel fp32 frame.
                             bits 8..11 (B200GS_BIN_SHIFT_HINT(s) = (s + 1) << 8, 0 = none): pairs are binned per
                             (16 << s)^2 pixels for this call instead of the automatic choice -- a performance hint
                             (results are identical for every bin size); b200gs_backward must get the same bits. */
  int64_t pair_capacity_hint; /* 0: size the binning buffer exactly (host waits for D before
                                 launching the binning stage, as the replaced interface does).
                                 >0: launch the whole frame for this many (Gaussian,tile) pair
                                 slots without waiting; D is read back asynchronously and the
                                 binning stage is redone exactly only if D exceeded the hint.
                                 Results are identical either way. */
} B200GSParams;

/* Caller-owned growable byte buffer: resize(ctx, nbytes) returns a device pointer to at least
 * nbytes (the reference binding's `std::function<char*(size_t)>` resize callbacks). */
typedef struct B200GSAlloc {
  void* ctx;
  char* (*resize)(void* ctx, size_t nbytes);
} B200GSAlloc;

/*
 * Forward: projection + SH colour, tile binning + sort, per-tile front-to-back compositing.
 *
 *   bg[3], viewmatrix[16], projmatrix[16], campos[3]   device; matrices TRANSPOSED (column-major)
 *   means3D[P][3]; shs[P][M][3] or NULL; colors_precomp[P][3] or NULL (exactly one of the two)
 *   opacities[P]; scales[P][3] + rotations[P][4] (w,x,y,z; used as given) or cov3D_precomp[P][6]
 *   out_color[3][H][W]; radii[P] (int32, 0 = culled)
 *   geom/binning/img: scratch allocators; *num_rendered receives the number of (Gaussian, tile)
 *   pairs D that were sorted and composited.
 *
 * With pair_capacity_hint == 0 one device->host read (of D, to size the binning buffer)
 * synchronises the stream, as in the interface this replaces; with a hint the host only waits on
 * an event recorded before the binning stage, after the rest of the frame has been queued.
 */
int b200gs_forward(const B200GSParams* prm, const float* bg, const float* viewmatrix,
                   const float* projmatrix, const float* campos, const float* means3D,
                   const float* shs, const float* colors_precomp, const float* opacities,
                   const float* scales, const float* rotations, const float* cov3D_precomp,
                   float* out_color, int32_t* radii, B200GSAlloc geom, B200GSAlloc binning,
                   B200GSAlloc img, int32_t* num_rendered, void* stream);

/*
 * Backward: adjoint of the whole forward w.r.t. (means3D, shs | colors_precomp, opacities,
 * scales, rotations | cov3D_precomp) plus the screen-space mean gradient the densification
 * heuristics read (dL_dmeans2D, NDC-scaled units, z component 0).
 *
 *   geom/binning/img: the byte buffers produced by the matching forward call (device pointers)
 *   dL_dout_color[3][H][W]
 *   outputs (every element is written; pass NULL for tensors of the unused input alternative):
 *     dL_dmeans3D[P][3], dL_dmeans2D[P][3], dL_dshs[P][M][3], dL_dcolors_precomp[P][3],
 *     dL_dopacities[P], dL_dscales[P][3], dL_drotations[P][4], dL_dcov3D[P][6]
 *   scratch: allocator for the per-Gaussian screen-space accumulators (48 B per Gaussian)
 */
int b200gs_backward(const B200GSParams* prm, const float* bg, const float* viewmatrix,
                    const float* projmatrix, const float* campos, const float* means3D,
                    const float* shs, const float* colors_precomp, const float* opacities,
                    const float* scales, const float* rotations, const float* cov3D_precomp,
                    const int32_t* radii, const char* geom, const char* binning, const char* img,
                    int32_t num_rendered, const float* dL_dout_color, float* dL_dmeans3D,
                    float* dL_dmeans2D, float* dL_dshs, float* dL_dcolors_precomp,
                    float* dL_dopacities, float* dL_dscales, float* dL_drotations,
                    float* dL_dcov3D, B200GSAlloc scratch, void* stream);

/* present[i] = 1 iff Gaussian i passes the near-plane cull (view-space z > 0.2). */
int b200gs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                        const float* projmatrix, uint8_t* present, void* stream);

/* out_alpha[H][W] = 1 - final transmittance of the forward call that produced `img` (the opacity
 * image a gsplat-style caller expects next to the colour). */
int b200gs_extract_alpha(const char* img, int32_t image_height, int32_t image_width, float* out_alpha,
                         void* stream);

/* Frame export for the datagen sweep: color[3][H][W] fp32 -> out_hwc[H][W][3] 8-bit RGB
 * (clamp to [0,1], *255, round to nearest).  The reference's per-frame scene render feeds image
 * datasets (/root/reference/README.md:85); exporting on the device cuts the device->host read of a
 * finished 1080p frame from 24.9 MB to 6.2 MB. */
int b200gs_export_rgb8(const float* color, int32_t image_height, int32_t image_width,
                       uint8_t* out_hwc, void* stream);

/*
 * Data formats either side of the rasterizer.
 *
 * b200gs_ply_activate: vertex records of a 3DGS .ply (the hand-off format the reference names,
 * /root/reference/README.md:75; binary little-endian float properties x,y,z,nx,ny,nz,f_dc_0..2,
 * f_rest_0..3*n_rest-1,opacity,scale_0..2,rot_0..3, stored PRE-activation) -> the packed tensors
 * b200gs_forward takes: sigmoid(opacity), exp(scale), normalised quaternion, shs[P][1+n_rest][3]
 * (f_rest is channel-major in the file).  `vertices` is the device copy of the vertex block; the
 * layout gives the record stride and property offsets in floats.
 */
typedef struct B200GSPlyLayout {
  int32_t stride;      /* floats per vertex record */
  int32_t off_xyz, off_fdc, off_frest, n_rest, off_opacity, off_scale, off_rot;
} B200GSPlyLayout;
int b200gs_ply_activate(int32_t P, const float* vertices, const B200GSPlyLayout* layout, float* means3D,
                        float* shs, float* opacities, float* scales, float* rotations, void* stream);

/*
 * b200gs_transform_gaussians: per-link rigid pose update of object Gaussians for the articulated
 * composite (links / hinge as produced by /root/reference/Articulation/urdf_generation/pipeline.py:
 * 290-357): means_out = R_l * mean + t_l, rots_out = q_l (x) rot, l = link_ids[i] (NULL = link 0).
 * link_transforms[L][12] are row-major 3x4, link_quats[L][4] are (w,x,y,z); in/out may alias.
 */
int b200gs_transform_gaussians(int32_t n, const float* means_in, const float* rots_in,
                               const int32_t* link_ids, const float* link_transforms,
                               const float* link_quats, int32_t L, float* means_out, float* rots_out,
                               void* stream);

/*
 * Training-step neighbour: photometric loss between a rendered image and its
 * target, n floats each (16-byte aligned).
 *   forward : *out_sum = sum_i w_l2 (a_i-b_i)^2 + w_l1 |a_i-b_i|        (one pass, device scalar)
 *   backward: dL_da_i  = *upstream * scale * (2 w_l2 (a_i-b_i) + w_l1 sign(a_i-b_i))
 */
int b200gs_photometric_loss(const float* a, const float* b, int64_t n, float w_l2, float w_l1, float* out_sum,
                            void* stream);
int b200gs_photometric_loss_backward(const float* a, const float* b, int64_t n, float w_l2, float w_l1,
                                     float scale, const float* upstream, float* dL_da, void* stream);

/*
 * Training-step neighbour: SSIM between a rendered image img1 and its target
 * img2, both [C][H][W] fp32 -- the definition every public 3DGS trainer uses (11x11 Gaussian window,
 * sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2, per channel; the loss uses the mean of the map).
 *   forward : *out_sum = sum of the SSIM map over C*H*W (device scalar).  If maps != NULL
 *             ([3][C][H][W] floats) the three derivative maps the backward pass needs are stored.
 *   backward: dL_dimg1 = *upstream * scale * d(sum of the SSIM map)/d(img1), from the stored maps.
 */
int b200gs_ssim_forward(const float* img1, const float* img2, int32_t C, int32_t H, int32_t W, float* maps,
                        float* out_sum, void* stream);
int b200gs_ssim_backward(const float* img1, const float* img2, const float* maps, int32_t C, int32_t H, int32_t W,
                         float scale, const float* upstream, float* dL_dimg1, void* stream);

/*
 * Training-step neighbour: one fused Adam step over up to
 * B200GS_ADAM_MAX_GROUPS parameter tensors in ONE launch (torch.optim.Adam semantics, no weight
 * decay, no amsgrad): m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
 *   p -= lr / (1 - b1^step) * m / (sqrt(v) / sqrt(1 - b2^step) + eps).   step counts from 1.
 */
#define B200GS_ADAM_MAX_GROUPS 8
typedef struct B200GSAdamGroup {
  float* param;          /* [n] updated in place */
  const float* grad;     /* [n] */
  float* exp_avg;        /* [n] first moment, updated in place */
  float* exp_avg_sq;     /* [n] second moment, updated in place */
  int64_t n;
  float lr;              /* learning rate of this tensor (3DGS uses one per parameter kind) */
  float reserved;
} B200GSAdamGroup;
int b200gs_adam_step(const B200GSAdamGroup* groups, int32_t num_groups, float beta1, float beta2, float eps,
                     int32_t step, void* stream);

/*
 * ---- Context: the no-stall protocol behind the C ABI ------------------------------------------------------------
 * b200gs_forward leaves three things to its caller: scratch memory, the pair-capacity hint that lets a frame be
 * launched before its pair count D is known, and the bin size that suits the scene's splat extent.  A context owns
 * all three, so that a C / pybind caller gets the same behaviour as the Python operator layer with no state of its
 * own (what the replaced interface's caller would otherwise have to rebuild around every rasterize_gaussians call):
 *   - scratch arenas (stream-ordered device allocations, grown geometrically, reused from frame to frame);
 *   - per (P, H, W): a slowly decaying maximum of recent pair counts -> hint = b200gs_policy_pair_capacity(D);
 *   - per (P, H, W): the bin size, chosen on the first synchronous frame (and re-checked every 256th) from the
 *     pairs-per-visible-Gaussian ratio and the frame's coverage -> b200gs_policy_bin_shift();
 *   - a ring of pinned words + events for frames whose pair-count check is deferred (tickets).
 * A context belongs to the device that was current at creation and serves ONE stream at a time (frames in flight on
 * several streams: one context per stream); calls on one context are serialised by an internal mutex.
 */
typedef struct B200GSContext B200GSContext;
int b200gs_context_create(B200GSContext** out_ctx);
int b200gs_context_destroy(B200GSContext* ctx);

/*
 * Forward through a context.  Same tensors as b200gs_forward; prm->pair_capacity_hint and the bin-shift bits of
 * prm->flags are ignored (the context supplies them).  An allocator whose `resize` is NULL selects the context's own
 * arena for that buffer.
 *   defer == 0: *num_rendered receives D (exact; the host waits for D only after the whole frame is queued, and only
 *               redoes the binning stage if D exceeded the hint); *ticket = -1.
 *   defer != 0: if a hint exists the call never waits: *ticket >= 0 identifies the frame, D arrives later --
 *               b200gs_context_ticket_wait() tells whether the frame is complete (D <= hint); a frame that is not
 *               must be rendered again with defer == 0.  Without a hint (first frame of a (P,H,W)) the call behaves
 *               as defer == 0 and returns *ticket = -1.
 * *used_flags receives the bin-shift bits the frame was rendered with: b200gs_context_backward (or b200gs_backward
 * via prm->flags) must get them back.
 */
int b200gs_context_forward(B200GSContext* ctx, const B200GSParams* prm, const float* bg, const float* viewmatrix,
                           const float* projmatrix, const float* campos, const float* means3D, const float* shs,
                           const float* colors_precomp, const float* opacities, const float* scales,
                           const float* rotations, const float* cov3D_precomp, float* out_color, int32_t* radii,
                           B200GSAlloc geom, B200GSAlloc binning, B200GSAlloc img, int32_t defer,
                           int32_t* num_rendered, int64_t* ticket, int32_t* used_flags, void* stream);

/* Waits until the stream has passed the deferred frame `ticket`; *num_rendered = D, *complete = (D <= its hint). */
int b200gs_context_ticket_wait(B200GSContext* ctx, int64_t ticket, int32_t* num_rendered, int32_t* complete);

/* Backward for the LAST b200gs_context_forward of this context that used the context's own arenas (NULL geom /
 * binning / img select them); otherwise identical to b200gs_backward.  used_flags: from that forward call.
 * After a DEFERRED forward (whose pair count the host does not know yet) pass any positive num_rendered -- the adjoint
 * only needs to know that there are pairs -- and validate the ticket afterwards (b200gs_context_ticket_wait): a training
 * step then never waits inside the step; if the ticket reports an incomplete frame, discard the gradients and repeat it. */
int b200gs_context_backward(B200GSContext* ctx, const B200GSParams* prm, int32_t used_flags, const float* bg,
                            const float* viewmatrix, const float* projmatrix, const float* campos,
                            const float* means3D, const float* shs, const float* colors_precomp,
                            const float* opacities, const float* scales, const float* rotations,
                            const float* cov3D_precomp, const int32_t* radii, const char* geom, const char* binning,
                            const char* img, int32_t num_rendered, const float* dL_dout_color, float* dL_dmeans3D,
                            float* dL_dmeans2D, float* dL_dshs, float* dL_dcolors_precomp, float* dL_dopacities,
                            float* dL_dscales, float* dL_drotations, float* dL_dcov3D, void* stream);

/* What the context currently holds for (P, H, W): last tracked pair count (0 = none) and bin shift (-1 = automatic). */
int b200gs_context_query(B200GSContext* ctx, int32_t P, int32_t image_height, int32_t image_width,
                         int64_t* tracked_pairs, int32_t* bin_shift);

/*
 * Captured frames with per-kernel priorities.  `graph` is a cudaGraph_t that holds one captured frame (b200gs_forward
 * with B200GS_DEFER_PAIR_CHECK, optionally b200gs_export_rgb8 and the caller's copies).  mode 0: plain instantiation.
 * mode 1: the kernels in front of compositing (projection, bucket scan, pair emission, pair sort) get the device's
 * highest priority, compositing and the 8-bit export the lowest, and the graph is instantiated with
 * cudaGraphInstantiateFlagUseNodePriority: with several frames in flight on different streams one frame's short,
 * latency-bound binning chain runs underneath another frame's compositing instead of queueing behind its CTAs.
 * mode 2: as 1 with a middle priority for the chain.  Results are identical.  Modes 1 and 2 set the priority attribute
 * on the kernel nodes of `graph` itself (the caller's graph is modified, not copied); the graph stays the caller's and
 * must outlive nothing -- the executable graph is self-contained.  *exec_out is a cudaGraphExec_t;
 * n_low / n_high (optional) receive the number of kernel nodes marked low / high.
 */
int b200gs_graph_instantiate(void* graph, int32_t mode, void** exec_out, int32_t* n_low, int32_t* n_high);
int b200gs_graph_launch(void* exec, void* stream);
int b200gs_graph_exec_destroy(void* exec);

/* The two rules of the protocol as pure host functions (no CUDA call; the Python operator layer uses the same):
 *   pair capacity for a tracked pair count D:  D + D/16 + 32768   (0 for D <= 0);
 *   bin shift for a frame that produced D pairs from `touching` visible Gaussians with (16 << used_shift)-px bins:
 *   bins of about three splat extents, extent = (sqrt(D / touching) - 1) * bin edge, clamped to 32..256 px, and one
 *   size coarser (256 px) when 128 px came out and the frame saturates everywhere (coverage > 0.995; pass a negative
 *   coverage if unknown).  Returns used_shift unchanged when there is nothing to decide (D or touching <= 0). */
int64_t b200gs_policy_pair_capacity(int64_t tracked_pairs);
int32_t b200gs_policy_bin_shift(int64_t D, int64_t touching, int32_t used_shift, float coverage);

/* Sizes of the forward scratch buffers for given P, H, W (geom, img) and D (binning); lets a
 * caller pre-size arenas.  Any of the out pointers may be NULL. */
int b200gs_buffer_sizes(int32_t P, int32_t image_height, int32_t image_width, int64_t D,
                        size_t* geom_bytes, size_t* binning_bytes, size_t* img_bytes);

/* Byte offsets of the per-Gaussian arrays inside the geom buffer of a forward call with P Gaussians -- lets a
 * test or a debugger compare the projection stage field by field without the library
 * exposing its scratch as API.  offsets[0] rec: float[P][12] = {x, y, conic A, B | C, opacity, threshold,
 * index bits | r, g, b, radius}, written only for Gaussians that touch a bin; [1] depth_key: uint32[P] IEEE bits
 * of the view-space depth (0xFFFFFFFF = culled); [2] tiles: uint32[P] bins touched (0 = none; with the bucketed binning
 * and at most 255 bins a footprint of 1..3 bins is stored packed: bit 31 set, count in bits 24..25, bin ids in the
 * low three bytes); [3] offsets: uint32[P]
 * inclusive scan of tiles (global-sort pipeline only); [4] clamped: uint8[P], bit c = colour channel c was
 * clamped at 0. */
#define B200GS_GEOM_FIELDS 5
int b200gs_geom_layout(int32_t P, size_t offsets[B200GS_GEOM_FIELDS]);

/* Message of the last error raised on the calling thread ("" if none). */
const char* b200gs_last_error(void);

/* B200GS_VERSION the library was built from. */
int b200gs_version(void);

/* Kernel launches issued by this process since the last call with reset != 0 (bench accounting). */
int64_t b200gs_launch_count(int reset);

/*
 * Process-wide tuning knobs (never change results):
 *   "bin_shift": -1 automatic (default), 0..5 = sort pairs per (16 << shift)^2 pixel bins
 *   "binning":    1 depth-sliced bucket binning (default; bucket.cu: per-(bin, depth slice) pair counters in the
 *                 projection kernel, one-CTA scan, cursor emission, one warp sorts one bucket in registers),
 *                 0 global radix sort of (bin << 32 | depth) keys (library scan + sort; kept as a cross-check),
 *                 -1 back to the default
 *   "sort_keys":  32 (default) global sort on 32-bit keys (bin << 24 | monotone 24-bit quantisation of the depth
 *                 bits; exact (depth, index) order restored inside runs of equal keys) whenever there are at
 *                 most 255 bins and the library sort is used -- four radix passes instead of five; 64 always
 *                 sorts the public algorithm's 64-bit (bin << 32 | depth bits) keys
 *   "render":     -1 automatic (default: four pixels per thread from 4096 tiles up, one pixel per thread
 *                 below), 1 compositing kernels with four pixels per thread, 0 one pixel per thread
 *   "gather":     1 LDGSTS record gather in the one-pixel compositing kernels (default), 0 TMA bulk copies
 *   "slab":       0 (default) the four-pixel compositing kernels gather records with LDGSTS; 1 the bucket sort also
 *                 writes the records in list order and the ring is filled by one TMA bulk copy per 64-record chunk
 *                 (must not change between a forward call and its backward call)
 *   "project":    0 (default) projection kernel with one thread per Gaussian; 1 dense-warp variant (warp-level
 *                 stream compaction: cull -> geometry -> colour)
 *   "bwd_overlap": 0 (default) the projection adjoint writes every gradient row, zeros included; 1 the gradient
 *                 tensors are zero-filled on an internal side stream while the compositing adjoint runs and the
 *                 projection adjoint writes only the rows of visible Gaussians
 *   "sort":       0 CUB radix sort, 1 automatic (default: single-launch cooperative radix sort for
 *                 pair lists <= 256 k, CUB above), 2 cooperative sort whenever the list is <= 3 M
 * Environment equivalents read at first use: B200GS_BIN_SHIFT, B200GS_GATHER=tma|ldgsts,
 * B200GS_SORT=cub|auto|coop,
 * B200GS_RENDER=auto|4px|1px, B200GS_BINNING=bucket|sort, B200GS_SORT_KEYS=64, B200GS_SLAB=1, B200GS_PROJECT=compact.
 */
int b200gs_set_option(const char* name, int value);

/*
 * Optional per-stage device timing (used by bench.py for the roofline figures).  While enabled,
 * forward/backward bracket each stage with CUDA events on the caller's stream; profile_read()
 * waits for the recorded events and returns, per stage, the summed milliseconds and the number of
 * timed launches since the last reset.  Stage i is named b200gs_stage_name(i).
 */
#define B200GS_NUM_STAGES 8
int b200gs_profile_enable(int on);
int b200gs_profile_read(float* stage_ms, int32_t* stage_calls, int reset);
const char* b200gs_stage_name(int i);

#ifdef __cplusplus
}
#endif
#endif /* B200GS_H */

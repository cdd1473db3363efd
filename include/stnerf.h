/*
 * stnerf.h -- C ABI of libstnerf_b200.so: the st-nerf layered ray-march hot path on B200 (sm_100a).
 *
 * The reference (DarlingHang/st-nerf) has no FFI seam: its boundary is the Python call surface
 *   modeling/layered_rfrender.py:141   LayeredRFRender.forward(rays, labels, bboxes, only_coarse, ...)
 *   utils/batchify_rays.py:51          layered_batchify_ray(model, rays, ...)
 *   engine/render.py:30                render(model, K, T, img_size, ...)
 * The Python facade in st-nerf_b200/{modeling,utils,layers,engine} keeps that surface and marshals it
 * onto the entry points below through ctypes.  Each entry point cites the reference
 * code it replaces.  Paths are relative to the reference root.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all functions return 0 on success or a negative STNERF_E* code; nothing throws; no hidden
 *     synchronisation: work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - the context owns its weights and a workspace; the workspace is (re)allocated only when a call
 *     asks for more rays-per-chunk / samples than any previous call (stnerf_reserve does it up front);
 *   - there is no CPU fallback: without a CUDA device every call returns STNERF_ENODEVICE.
 */
#ifndef STNERF_H_
#define STNERF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STNERF_MAX_LAYERS 8      /* l = 1 background + up to 7 performers (BASELINE config #5 uses 7) */
#define STNERF_MAX_N1 128        /* coarse samples per ray per layer                                  */
#define STNERF_MAX_S 512         /* n1 + n2                                                           */

enum {
  STNERF_OK = 0,
  STNERF_EINVAL = -1,      /* bad argument (the reference prints + exit(-1), layered_rfrender.py:162) */
  STNERF_ENODEVICE = -2,   /* no CUDA device / wrong architecture                                    */
  STNERF_ECUDA = -3,       /* a CUDA runtime call failed; stnerf_last_cuda_error() has the text      */
  STNERF_ENOWEIGHTS = -4,  /* a network needed by the call was never loaded                          */
  STNERF_ENOMEM = -5
};

/* Arithmetic mode of the two MLPs (modeling/spacenet.py:101-160, modeling/motion_net.py:34-71). */
enum {
  STNERF_PREC_FP32_SIMT = 0,   /* fp32 FFMA on CUDA cores: the bit-closest mode, validation baseline      */
  STNERF_PREC_TC_3XF16 = 1,    /* tcgen05 fp16 3-term split (hi*hi + lo*hi + hi*lo), fp32 accumulate:     */
                               /* ~fp32 products, meets the 1e-3 RGB gate               */
  STNERF_PREC_TC_F16 = 2,      /* tcgen05 single fp16 pass: fastest, does NOT meet the 1e-3 gate          */
  STNERF_PREC_TC_MIXED = 3,    /* TC_3XF16 for everything the density depends on (SpaceNet trunk + sigma head, MotionNet); */
                               /* single fp16 pass for the colour-only layer rgb_net.1 (spacenet.py:81-86): ~5 % fewer     */
                               /* MMAs, colour error <= 2.5e-4, resampling untouched -- inside the 1e-3 gate               */
  STNERF_PREC_TC_3XF16_CF = 4  /* TC_3XF16 with the two correction products of every layer issued FIRST (over the whole K  */
                               /* range, then hi*hi) in the coarse pass and the MotionNets -- what the sample placement    */
                               /* depends on.  The tensor core truncates when it adds into its fp32 accumulator; this      */
                               /* order truncates at full magnitude K/16 instead of 3K/16 times: sigma error / 3 (2e-6 rel */
                               /* rms, the reference's own fp32 noise is 1e-6), about half the rays over the 1e-3 gate at  */
                               /* scale; the hi weight stages stream twice (cost: see DESIGN 3.1)                          */
};

typedef struct stnerf_ctx* stnerf_handle;

typedef struct {
  int32_t n_layers;                              /* l = L+1, layer 0 = background (layered_rfrender.py:56,209) */
  int32_t space_time[STNERF_MAX_LAYERS];         /* net of layer i consumes PE(time) in its rgb head           */
                                                 /* (cfg.MODEL.USE_SPACE_TIME, layered_rfrender.py:62-67)       */
  int32_t precision;                             /* STNERF_PREC_*                                               */
  int32_t chunk_rays;                            /* rays per internal chunk; 0 = default (65536); at most 2^22-1  */
} stnerf_model_desc;

/* Per-call scene constants: what LayeredRFRender.forward derives in its prologue
 * (layered_rfrender.py:190-242) plus the attributes the renderer mutates between frames
 * (render/layered_neural_renderer.py:435-440).  All host memory. */
typedef struct {
  float bmin[STNERF_MAX_LAYERS][3];              /* box corner 0 after scale/shift edits (:230-242)             */
  float bmax[STNERF_MAX_LAYERS][3];              /* box corner 6 after edits                                    */
  int32_t shown[STNERF_MAX_LAYERS];              /* display_layers (:104-112); entry 0 is ignored like the ref. */
  /* inverse edit applied to sample points before the nets: p -= shift; p = (p - pivot)/scale + pivot           */
  int32_t shift_on[STNERF_MAX_LAYERS];           /* (:293-298) == (:467-471)                                    */
  float shift[STNERF_MAX_LAYERS][3];
  int32_t scale_coarse_on[STNERF_MAX_LAYERS];    /* (:300-303)                                                  */
  int32_t scale_fine_on[STNERF_MAX_LAYERS];      /* (:473-475), skipped when that layer's shift entry is None   */
  float scale[STNERF_MAX_LAYERS];
  float pivot[3];                                /* (centre_layer2 + centre_layer1)/2 (:221-232)                */
  float near_plane;                              /* model.near (:143,422,605)                                   */
  float alpha_layer2;                            /* model.alpha, fine pass, layer 2 only (:575-576)             */
  float density_threshold;                       /* (:416-418, 564-566)                                         */
  float bkgd_density_threshold;                  /* (:538-547)                                                  */
  float boarder_weight;                          /* cfg.MODEL.BOARDER_WEIGHT, last delta (render_layer.py:38)   */
  int32_t apply_thresholds;                      /* 1 in the retiming branch (:416,538,564), else 0             */
  int32_t shared_frame_id;                       /* 1: 7-column rays [o,d,frame_id] of the evaluator (:157-158,171):   */
                                                 /* every layer reads column 6; 0: one column per layer (retiming)      */
} stnerf_scene;

/* ---- lifetime ------------------------------------------------------------------------------------------ */
int stnerf_create(stnerf_handle* out, const stnerf_model_desc* desc_host);   /* modeling/__init__.py:5-7 build_layered_model */
void stnerf_destroy(stnerf_handle h);
int stnerf_reserve(stnerf_handle h, int n1, int n2);                         /* size the workspace up front                  */
size_t stnerf_workspace_bytes(stnerf_handle h);
const char* stnerf_strerror(int code);
const char* stnerf_last_cuda_error(void);
int stnerf_set_precision(stnerf_handle h, int precision);

/* ---- weights: load_state_dict (render/layered_neural_renderer.py:110-117) ------------------------------ */
/* blob = the net's tensors concatenated in state_dict order, fp32, nn.Linear (out,in) row-major:
 *   SpaceNet : stage1.{0,2,4,6}.{weight,bias}, stage2.{0,2,4}.{weight,bias}, density_net.0.{weight,bias},
 *              rgb_net.1.{weight,bias}, rgb_net.3.{weight,bias}      (464260 or 466948 floats)
 *   MotionNet: motion_net.{0,2,4,6,8,10}.{weight,bias}               (77315 floats)
 * layer 0 = background; fine = 0 coarse net / 1 fine net.  MotionNets are shared by both passes.        */
int stnerf_load_spacenet(stnerf_handle h, int layer, int fine, const float* blob_host, size_t n_floats);
int stnerf_load_motionnet(stnerf_handle h, int layer, const float* blob_host, size_t n_floats);

int stnerf_set_scene(stnerf_handle h, const stnerf_scene* scene_host);
/* Per-frame boxes for rays that carry their OWN frame id: the 7-column rays of a mixed-frame batch (training / evaluation),
 * `bboxes = self.bboxes.index_select(0, rays_frame_id - 1)` (modeling/layered_rfrender.py:193).  table_host:
 * [n_frames][l][2][3] = (min, max) corner of every layer's box at every frame AFTER the scale / shift edits (:230-242), entry
 * [f][0] = the background box.  With a table set and scene.shared_frame_id = 1, ray r is clipped against row
 * (int)rays[r][6] - 1 (ids outside [1, n_frames] are clamped; the reference raises).  n_frames = 0 removes the table.      */
int stnerf_set_box_table(stnerf_handle h, const float* table_host, int n_frames);

/* ---- the hot path: LayeredRFRender.forward (layered_rfrender.py:141-734), BBOX sampling ---------------- */
/* rays: (n_rays, ray_stride) fp32, columns [o(3), d(3), frame_id_layer0 .. frame_id_layer(l-1)]
 *       (data/datasets/ray_dataset.py:276-281); ray_stride >= 6 + l  (>= 7 with scene.shared_frame_id).
 * jitter: (l, n_rays, n1) uniforms for layers/RaySamplePoint.py:98, or NULL -> in-kernel Philox(seed).
 * u:      (l, n_rays, n2) uniforms for utils/sample_pdf.py:31, or NULL -> Philox(seed).
 * out:    [2 passes: 0 coarse, 1 fine][l+1 images: 0 mixed, 1+i layer i] planes of 5*n_rays floats each,
 *         a plane = rgb (n_rays,3) | depth (n_rays) | acc (n_rays)   (the tuples of :725-734).
 *         With only_coarse the fine planes are left untouched (the facade aliases them, :721-722).
 *         The caller's current device must be the one the context was created on (else STNERF_EINVAL).
 * ray_mask: (l, n_rays) uint8, |bin_width| > 1e-5 (layers/RaySamplePoint.py:105).                        */
int stnerf_render(stnerf_handle h, const float* rays, int64_t n_rays, int ray_stride, int n1, int n2,
                  int only_coarse, const float* jitter, const float* u, uint64_t seed,
                  float* out, uint8_t* ray_mask, void* stream);

/* Keys of the in-kernel Philox stream: ray j of the following stnerf_render calls draws the uniforms of ray id
 *   base + (j / width) * row_stride + (j % width)      (width = 0: id = base + j, the default).
 * A caller that renders row-interleaved shards of an image (one process per GPU) sets base = first_row * W, width = W,
 * row_stride = n_ranks * W so that every pixel gets the same draws as in an unsharded render of the same seed.         */
int stnerf_set_ray_ids(stnerf_handle h, int64_t base, int32_t width, int64_t row_stride);

/* Same call with HOST buffers (pinned for overlap; pageable works, serialised by the driver): the rays go up chunk by chunk
 * on a copy stream while earlier chunks render, every finished chunk's slices of out / ray_mask come down on a second copy
 * stream while later chunks render; returns after everything has drained.  Replaces the `.cuda()` ... `.cpu()` bracket of
 * render_pose (render/layered_neural_renderer.py:372-392, 451-454).  This is the e2e path bench.py times.
 * Device staging is sized by stnerf_reserve_host; a call larger than anything reserved (re)allocates it first.            */
int stnerf_render_host(stnerf_handle h, const float* rays_host, int64_t n_rays, int ray_stride, int n1, int n2,
                       int only_coarse, uint64_t seed, float* out_host, uint8_t* ray_mask_host, void* stream);
/* Pre-size the device staging of the host-buffer entry points (rays, image planes, masks, per-view image buffers, copy
 * streams and events) for calls of up to `max_rays` rays, so that those calls allocate nothing.                            */
int stnerf_reserve_host(stnerf_handle h, int64_t max_rays, int ray_stride);

/* ---- the renderer-facing fast path: LayeredNeuralRenderer.render_pose / render_path ------------------------------------ */
/* (render/layered_neural_renderer.py:364-392, 401-488; data/datasets/ray_dataset.py:260-283 for the rays of a pose.)
 * One entry of a batch of poses: camera, the per-layer frame ids of the pose (`layer_frame_pair`, ray_dataset.py:276-279),
 * the scene constants of THIS frame (boxes lerped to these frame ids, per-frame shift/scale/alpha edits
 * render/layered_neural_renderer.py:435-440, thresholds) and the seed of its Philox stream.  All host memory.           */
typedef struct {
  float Kinv[9];                                 /* inverse intrinsics, row-major                                              */
  float T[16];                                   /* camera-to-world, row-major                                                  */
  float frame_ids[STNERF_MAX_LAYERS];            /* column 6+i of every ray of this view                                        */
  stnerf_scene scene;
  uint64_t seed;
} stnerf_view;
/* Renders rows row0, row0+row_step, ... (n_rows of them) of an HxW image for each of n_views poses in ONE call: rays are
 * generated on the device (never cross PCIe), and only what render_pose returns is produced -- the FINE images, mixed + one per
 * layer, pixel-interleaved: images[v*view_stride + (img*n_rows*W + pixel)*5 + {0,1,2: rgb, 3: depth, 4: acc}].  With
 * coarse_images == NULL the coarse pass only resamples (no coarse images, no merged coarse composite); with a buffer of the same
 * shape it also produces the coarse images, i.e. everything LayeredRFRender.forward returns (modeling/layered_rfrender.py:725-734).  The last requested row may lie one row_step past H-1 (equal
 * shard sizes when H is not a multiple of row_step): it is rendered as an extrapolated pixel row the caller discards.  Enqueue only; `images` is a DEVICE buffer -- e.g. one
 * rank's slice of an all-gather buffer when the rows of a view are interleaved over GPUs (row0 = rank, row_step = ranks).    */
int stnerf_render_views(stnerf_handle h, const stnerf_view* views_host, int n_views, int H, int W, int row0, int row_step,
                        int n_rows, int n1, int n2, float* images, float* coarse_images, int64_t view_stride, void* stream);
/* Full HxW frames to a HOST buffer ([n_views][l+1][H*W][5], pinned for overlap): the device->host copy of view v runs on a
 * copy stream while view v+1 renders (two device image buffers); returns after the last copy has landed.                   */
int stnerf_render_views_host(stnerf_handle h, const stnerf_view* views_host, int n_views, int H, int W, int n1, int n2,
                             float* images_host, void* stream);

/* ---- ray generation: utils/render_helpers.py:96-123 == utils/ray_sampling.py:22-72 --------------------- */
/* Kinv_host = inverse(K) (3x3 row-major), T_host = camera-to-world (4x4 row-major).  Writes rows
 * row0, row0+row_step, ... (n_rows of them) of an HxW image: rays[(k*W + j)*ray_stride + 0..5] = o,d and
 * columns 6..6+n_frame_ids-1 = frame_ids_host (data/datasets/ray_dataset.py:276-281).                     */
int stnerf_raygen(const float* Kinv_host, const float* T_host, int H, int W, int row0, int row_step, int n_rows,
                  const float* frame_ids_host, int n_frame_ids, float* rays, int ray_stride, void* stream);

/* ---- per-stage entry points (unit parity against the reference function named) ------------------------ */
/* layers/RaySamplePoint.py:8-62 + :85-105 for one box.  t (n,n1), xyz (n,n1,3) or NULL, mask (n) uint8,
 * tfar_tnear (n,2) or NULL = intersection()'s return.                                                     */
int stnerf_intersect_sample(const float* rays, int64_t n, int ray_stride, const float* bmin_host,
                            const float* bmax_host, int is_bkgd, int n1, const float* jitter,
                            float* t, float* xyz, uint8_t* mask, float* tfar_tnear, void* stream);
/* layers/render_layer.py:25-58.  t (n,S), rgb (n,S,3), sigma (n,S) -> color (n,3), depth (n), acc (n), w (n,S)|NULL */
int stnerf_composite(const float* t, const float* rgb, const float* sigma, int64_t n, int S, float boarder,
                     float* color, float* depth, float* acc, float* w, void* stream);
/* utils/sample_pdf.py:18-63 (+ the sort of layered_rfrender.py:462 when t_fine != NULL).
 * t (n,n1), w (n,n1) full weights (the [1:-1] slice is taken inside), u (n,n2) -> z (n,n2)|NULL, t_fine (n,n1+n2)|NULL */
int stnerf_sample_pdf(const float* t, const float* w, const float* u, int64_t n, int n1, int n2,
                      float* z, float* t_fine, void* stream);
/* utils/dimension_kernel.py:24-33.  x (P,dim) -> out (P, dim*(1+2*n_freq)) */
int stnerf_positional_encoding(const float* x, int64_t P, int dim, int n_freq, float* out, void* stream);
/* modeling/spacenet.py:101-160.  pos (P,3), dirs (P,3), times (P)|NULL -> rgb (P,3) raw, sigma (P) raw */
int stnerf_spacenet(stnerf_handle h, int layer, int fine, const float* pos, const float* dirs, const float* times,
                    int64_t P, float* rgb, float* sigma, void* stream);
/* modeling/motion_net.py:34-71.  xyzt (P,4) -> flow (P,3).  lerp_mode: -1 = decide like the reference
 * (any non-integer t in the batch, :53), 0/1 = force.                                                      */
int stnerf_motionnet(stnerf_handle h, int layer, const float* xyzt, int64_t P, int lerp_mode, float* flow,
                     void* stream);

/* Packed-weight image (cache next to the checkpoint; replaces re-running the state_dict -> MMA-layout packing that follows
 * render/layered_neural_renderer.py:109-117 `torch.load` + `load_state_dict`).  `export` writes every loaded network's
 * device images (fp32 SIMT layout, fp16 hi/lo tensor-core stream, fp32 bias/head block) into a HOST buffer; with
 * `host_buf == NULL` it only reports the size.  `import` validates the image against the context (layer count, per-layer
 * time inputs, sizes) before touching any network and restores the weights without re-packing.  Byte-for-byte the same
 * device state as `stnerf_load_*` on the original tensors, so renders are bit-identical.                                   */
int stnerf_weights_export(stnerf_handle h, void* host_buf, size_t capacity, size_t* bytes_needed);
int stnerf_weights_import(stnerf_handle h, const void* host_buf, size_t bytes);

/* Tensor-core plumbing self-test: one 128x128x64 fp16 UMMA through the library's descriptors, swizzled layout, bulk
 * copy and TMEM load; writes max |D - host reference| (expected < 1e-3).                                     */
int stnerf_selftest_umma(float* max_err_host);
/* Accumulation probe: the same 128x256x64 product of all-POSITIVE fp16 operands accumulated `reps` times into one TMEM
 * accumulator (4*reps MMAs of K=16).  Reports max |D - fp64 sum| and the mean SIGNED relative error: how the tensor core
 * rounds when it adds into an fp32 accumulator (a negative mean growing with reps = round-toward-zero accumulation), which
 * bounds how close the fp16x3 split can get to the reference's fp32 GEMMs.                          */
int stnerf_selftest_umma_accum(int reps, float* max_err_host, float* mean_signed_rel_err_host);
/* The same through the CTA-pair protocol (`tcgen05.mma.cta_group::2`, M = 256 over the two CTAs of a cluster: remote mbarrier
 * arrives, multicast commit, paired TMEM allocation): one 256x256x64 product; expected < 1e-3.                 */
/* The same 128x256x64 product with the A operand in TENSOR memory: written with tcgen05.st in the layout the SpaceNet epilogue
 * uses for the next layer's activations (8 columns of fp16 pairs per K=16 step at a 16-column pitch), read by
 * tcgen05.mma [d], [a], b-desc.  Pins that layout on the device the library runs on. */
int stnerf_selftest_umma_ts(float* max_err_host);
int stnerf_selftest_umma_pair(float* max_err_host);

/* Diagnostic read-back of the sample depths of the LAST chunk rendered by stnerf_render (parity tooling: which depths did
 * utils/sample_pdf.py:18-63 + the sort of modeling/layered_rfrender.py:462 produce for these rays?).
 * what = 0: coarse depths of `layer`, (n_rays, n1);  what = 1: fine depths, (n_rays, n1+n2).  dst is a DEVICE buffer of
 * n_rays*S floats; n_rays must not exceed the rays of that chunk and S must be the sample count of that call.          */
int stnerf_debug_read_depths(stnerf_handle h, int what, int layer, float* dst, int64_t n_rays, int S, void* stream);

/* Number of kernels this library has launched since load (bench.py's gpu_launches claim). */
uint64_t stnerf_launch_count(void);

/* ---- measurement: per-kernel-class device times taken with CUDA events on the launching stream ----------- */
/* classes: 0 SpaceNet MLP, 1 MotionNet MLP, 2 sampling, 3 compositing/resampling.  `points` = network
 * evaluations (classes 0,1) or rays (2) / ray-passes (3) processed, so callers can turn ms into FLOP/s or B/s. */
typedef struct {
  double ms[4];
  double points[4];
  uint64_t launches[4];
} stnerf_profile;
int stnerf_profile_begin(stnerf_handle h);                       /* start recording (adds two events per launch)   */
int stnerf_profile_end(stnerf_handle h, stnerf_profile* out_host); /* drain the device, stop recording, return totals */

#ifdef __cplusplus
}
#endif
#endif /* STNERF_H_ */

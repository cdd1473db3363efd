// Shared declarations of libstnerf_b200 (device structs, error plumbing, Philox).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/stnerf.h"

namespace stnerf {

// ---------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------
extern thread_local char g_cuda_err[512];
extern std::atomic<unsigned long long> g_launches;    // kernels launched by this library (contexts may live on several host threads)

#define STNERF_CUDA(expr)                                                                        \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      snprintf(stnerf::g_cuda_err, sizeof(stnerf::g_cuda_err), "%s:%d %s -> %s", __FILE__, __LINE__, #expr, \
               cudaGetErrorString(e_));                                                          \
      return STNERF_ECUDA;                                                                       \
    }                                                                                            \
  } while (0)

#define STNERF_LAUNCH_CHECK()                                                                    \
  do {                                                                                           \
    ++stnerf::g_launches;                                                                        \
    STNERF_CUDA(cudaGetLastError());                                                             \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// network dimensions
// ---------------------------------------------------------------------------------------------------------
constexpr int PE_POS = 63;      // 3 + 3*2*10   utils/dimension_kernel.py, modeling/spacenet.py:20
constexpr int PE_DIR = 27;      // 3 + 3*2*4
constexpr int PE_TIME = 21;     // 1 + 2*10
constexpr int PE_MOTION = 84;   // 4 + 4*2*10   modeling/motion_net.py:14
constexpr int HID = 256;        // SpaceNet backbone
constexpr int HEAD = 128;       // SpaceNet rgb head / MotionNet width
constexpr int SPACENET_FLOATS_NOTIME = 464260;
constexpr int SPACENET_FLOATS_TIME = 466948;
constexpr int MOTIONNET_FLOATS = 77315;

// fp32 weights of one SpaceNet, repacked K-major-transposed ([k][n], n contiguous) for the SIMT kernels.
struct SpaceNetW {
  const float* w[7];     // stage1.{0,2,4,6} (K=63,256,256,256), stage2.{0,2,4} (K=319,256,256), each [K][256]
  const float* b[7];     // [256]
  const float* w_sigma;  // [256]
  float b_sigma;
  const float* w_rgbh;   // [256+27(+21)][128]
  const float* b_rgbh;   // [128]
  const float* w_rgbo;   // [3][128] (row-major as in the checkpoint)
  float b_rgbo[3];
  int use_time;
};

struct MotionNetW {
  const float* w[5];     // [84][128], 4 x [128][128]
  const float* b[5];
  const float* w_out;    // [3][128]
  float b_out[3];
};

// Where a tile of points comes from.
enum { SRC_EXPLICIT = 0, SRC_MARCH = 1, SRC_XYZ = 2, SRC_XYZ_MAP = 3 };
struct PointSrc {
  int mode;
  // SRC_EXPLICIT: per-point arrays (unit entry points)
  const float* pos;          // (P,3)  | SRC_XYZ: compact (slot*S+k, 3) deformed positions
  const float* dirs;         // (P,3)
  const float* times;        // (P) or null
  int pos_stride, time_stride;   // SRC_EXPLICIT element strides (3 / 1 unless the inputs are interleaved)
  // SRC_MARCH / SRC_XYZ: points = (slot, k), ray = hit ? hit[slot] : slot
  const float* rays;         // (n, ray_stride): o, d, frame ids
  int ray_stride;
  // SRC_XYZ_MAP (fine pass with flow reuse): position k of a slot's S depths came from coarse sample m = src_map[ray*S + k] < n_first
  // (deformed position pos[(slot*n_first + m)*3]) or from new depth m - n_first (pos2[(slot*(S - n_first) + m - n_first)*3])
  const uint8_t* src_map;
  const float* pos2;
  int n_first;
  const int* hit;            // slot -> ray, or null (identity: background)
  const int* count;          // device-side number of slots, or null
  long long n_slots;         // slots when count == null; P for SRC_EXPLICIT (with S == 1)
  long long n_slots_cap;     // upper bound on *count (grid sizing of per-slot helper kernels)
  const float* t;            // (rays, S) depths of this layer
  int S;
  int layer;                 // frame id column = 6 + layer
  // inverse edit on marched points (layered_rfrender.py:293-303 / :467-475)
  int shift_on, scale_on;
  float shift[3], scale, pivot[3];
};

__device__ __forceinline__ long long src_num_points(const PointSrc& s) {
  long long slots = s.count ? (long long)(*s.count) : s.n_slots;
  return slots * (long long)s.S;
}

// Maps a ray's index within a call to the id that keys the Philox stream (identity by default).  A caller that renders
// an image in row-interleaved shards sets (base, width, row_stride) so every pixel draws the same uniforms as in an
// unsharded render:  id = base + (j / width) * row_stride + (j % width).
struct RayIdMap {
  long long base, row_stride;
  int width;
  __host__ __device__ unsigned long long operator()(long long j) const {
    return width > 0 ? (unsigned long long)(base + (j / width) * row_stride + (j % width)) : (unsigned long long)(base + j);
  }
};

// Scene constants as the kernels see them.
struct DevScene {
  float bmin[STNERF_MAX_LAYERS][3];
  float bmax[STNERF_MAX_LAYERS][3];
  int shown[STNERF_MAX_LAYERS];
  float near_plane, alpha2, thr_layer, thr_bkgd, boarder;
  int apply_thr;
  int n_layers;
  int fid_shared;            // all layers read frame-id column 6 (7-column rays)
};

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 (production RNG when no uniforms are injected; statistically equivalent to torch.rand,
// not bit-equal -- parity runs inject uniforms).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// uniform in [0,1) with 24 random bits, like torch.rand for float32
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t stream, uint64_t ray, uint32_t idx) {
  uint4 c = make_uint4((uint32_t)ray, (uint32_t)(ray >> 32), idx >> 2, stream);
  uint4 r = philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t v = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  return u01(v);
}

// ---------------------------------------------------------------------------------------------------------
// kernel launchers implemented in the .cu files (all enqueue on `st`, return STNERF_* codes)
// ---------------------------------------------------------------------------------------------------------
// geometry.cu
int launch_raygen(const float* Kinv, const float* T, int H, int W, int row0, int row_step, int n_rows,
                  const float* fids, int n_fids, float* rays, int ray_stride, cudaStream_t st);
int launch_sample(const float* rays, long long n, int ray_stride, const DevScene& scene, int n_layers, int n1,
                  const float* jitter, long long jitter_layer_stride, uint64_t seed, long long ray_base, RayIdMap idmap,
                  float* t_coarse, long long t_layer_stride, uint8_t* mask, long long mask_layer_stride,
                  int* hit, long long hit_layer_stride, int* counts, int* lerp_flags, cudaStream_t st,
                  const float* box_table = nullptr, int n_frames = 0);
int launch_intersect_sample(const float* rays, long long n, int ray_stride, const float* bmin, const float* bmax,
                            int is_bkgd, int n1, const float* jitter, float* t, float* xyz, uint8_t* mask,
                            float* tfar_tnear, cudaStream_t st);
int launch_posenc(const float* x, long long P, int dim, int n_freq, float* out, cudaStream_t st);

// composite.cu
struct CompositeArgs {
  const float* t;          // [layer][ray][S]
  long long t_layer_stride;
  const float* raw;        // [layer][ray][S][4]  (rgb raw, sigma raw)
  long long raw_layer_stride;
  const uint8_t* mask;     // [layer][ray] hit masks (chunk-local)
  long long mask_layer_stride;
  const float* u;          // [layer][ray][n2] injected uniforms or null
  long long u_layer_stride;
  float* t_fine;           // out (coarse pass with n2 > 0): [layer][ray][n1+n2]
  long long tf_layer_stride;
  float* z_new;            // out, optional: [layer][ray][n2] the new depths, ascending (flow reuse, see resample.cuh)
  long long zn_layer_stride;
  uint8_t* src_map;        // out, optional: [layer][ray][n1+n2] origin of every fine depth
  long long sm_layer_stride;
  float* out;              // images of this pass: [img][5*n_total], or null (coarse pass: resampling only, no images)
  unsigned skip_layers;    // bit i: layer i's own image + resampling were produced elsewhere (fused SpaceNet kernel): gather only
  int pixel_layout;        // 0: plane = rgb (N,3) | depth (N) | acc (N);  1: plane = (N,5) pixel-interleaved
  long long n_total;       // rays in the whole call (plane geometry)
  long long ray_base;      // first ray of this chunk within the call
  long long n;             // rays in this chunk
  int S, n2, fine;
  uint64_t seed;
  RayIdMap idmap;
};
int launch_composite_pass(const CompositeArgs& a, const DevScene& scene, int n_layers, cudaStream_t st);
int launch_composite_simple(const float* t, const float* rgb, const float* sigma, long long n, int S, float boarder,
                            float* color, float* depth, float* acc, float* w, cudaStream_t st);
int launch_sample_pdf(const float* t, const float* w, const float* u, long long n, int n1, int n2, float* z,
                      float* t_fine, cudaStream_t st);

// mlp_simt.cu
int launch_spacenet_simt(const PointSrc& src, const SpaceNetW& w, float* raw, long long raw_slot_stride,
                         float* rgb_out, float* sigma_out, int num_sms, cudaStream_t st);
int launch_motionnet_simt(const PointSrc& src, const MotionNetW& w, const int* lerp_flag_dev, int lerp_force,
                          float* xyz_out, float* flow_out, int num_sms, cudaStream_t st);
size_t simt_smem_bytes();

}  // namespace stnerf

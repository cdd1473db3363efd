// C ABI of libstnerf_b200 (include/stnerf.h): context, weight packing, workspace, chunked render orchestration.
//
// The orchestration restates the control flow of modeling/layered_rfrender.py:141-734 (BBOX sampling) as a
// stream-ordered sequence of kernels per chunk of rays -- no host synchronisation, hit counts stay on the device:
//   sample -> [bkgd SpaceNet] -> per performer [MotionNet -> SpaceNet] -> composite+resample
//          -> same nets (fine weights) on n1+n2 depths -> per-layer + merged composite.
#include <algorithm>
#include <new>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"
#include "mlp_tc.cuh"

namespace stnerf {
thread_local char g_cuda_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
}  // namespace stnerf

using namespace stnerf;

namespace {

struct SpaceNetDev {
  float* blob = nullptr;       // SIMT layout (transposed fp32)
  SpaceNetW w{};
  TcNet tc{};                  // tcgen05 packing (mlp_tc.cu)
  bool loaded = false;
};
struct MotionNetDev {
  float* blob = nullptr;
  MotionNetW w{};
  TcNet tc{};
  bool loaded = false;
};

}  // namespace

struct stnerf_ctx {
  stnerf_model_desc desc{};
  int l = 0, num_sms = 0, device = 0, precision = 0, chunk_rays = 65536;
  SpaceNetDev space[2][STNERF_MAX_LAYERS];
  MotionNetDev motion[STNERF_MAX_LAYERS];
  stnerf_scene scene{};
  DevScene dscene{};
  bool have_scene = false;
  // workspace (sized for chunk_rays rays, cap_n1 coarse and cap_s2 total samples)
  int cap_n1 = 0, cap_s2 = 0;
  long long last_chunk_rays = 0;       // geometry of the most recent chunk (stnerf_debug_read_depths)
  int last_n1 = 0, last_s2 = 0;
  float *t_coarse = nullptr, *raw_coarse = nullptr, *t_fine = nullptr, *raw_fine = nullptr, *xyz = nullptr;
  // flow reuse (fine pass): per performer layer the coarse pass' deformed points, the new depths and the origin map of t_fine
  float *xyz_coarse = nullptr, *z_new = nullptr;
  uint8_t* src_map = nullptr;
  bool no_reuse = false;       // STNERF_NO_REUSE=1 at create: the fine pass evaluates the MotionNet on all n1+n2 depths (A/B)
  float* cbuf = nullptr;       // per-slot rgb_net.1 bias of the SpaceNet being evaluated (tensor-core modes)
  uint8_t* mask_ws = nullptr;
  int *hit = nullptr, *counts = nullptr, *lerp_flags = nullptr;
  size_t ws_bytes = 0;
  // staging for stnerf_render_host
  float *h_rays = nullptr, *h_out = nullptr;
  uint8_t* h_mask = nullptr;
  size_t h_rays_bytes = 0, h_out_bytes = 0, h_mask_bytes = 0;
  // staging for stnerf_render_views(_host): rays of one view, two image buffers (double-buffered device->host copies)
  float *v_rays = nullptr, *v_img[2] = {nullptr, nullptr};
  size_t v_rays_bytes = 0, v_img_bytes[2] = {0, 0};
  cudaStream_t copy_in = nullptr, copy_out = nullptr;      // host<->device copies that overlap the kernels of other chunks
  std::vector<cudaEvent_t> ev_pool;                         // timing-disabled events, reused call after call
  float* box_table = nullptr;  // [n_frames][l][2][3] per-frame boxes for rays with their own frame id (stnerf_set_box_table)
  int box_frames = 0;
  // Order of the split MMAs (mlp_tc.cu): 0 = interleaved everywhere, 1 = correction products first in the COARSE pass and in the
  // MotionNets (what the sample placement and the positions depend on), interleaved in the fine SpaceNet pass, 2 = first everywhere.
  // -1 (default): by precision -- STNERF_PREC_TC_3XF16_CF means 1, every other mode 0.
  int lo_first_env = -1;       // STNERF_LO_FIRST=0|1|2 in the environment at create overrides (A/B: profiles/r02_ab_lo_first.json)
  int lo_first_mode() const { return lo_first_env >= 0 ? lo_first_env : (precision == STNERF_PREC_TC_3XF16_CF ? 1 : 0); }
  bool no_fuse = false;        // STNERF_NO_FUSE=1 in the environment at create: keep the coarse compositing in its own kernel (A/B)
  int* any_frac = nullptr;     // scratch flag for stnerf_motionnet(lerp_mode=-1)
  RayIdMap idmap{0, 0, 0};     // stnerf_set_ray_ids
  // profiling (stnerf_profile_begin / _end): CUDA-event pairs around every launch, on the launching stream
  struct ProfRec { int cls; cudaEvent_t a, b; double points; int count_slot; int S; };
  bool prof_on = false;
  std::vector<ProfRec> prof;
  int* prof_counts = nullptr;  // pinned host copies of the per-chunk hit counts
  int prof_chunks = 0;
};

namespace {
constexpr int PROF_MAX_CHUNKS = 1 << 16;
struct ProfScope {
  stnerf_ctx* c; int idx = -1; cudaStream_t st;
  ProfScope(stnerf_ctx* c_, int cls, double points, int count_slot, int S, cudaStream_t st_) : c(c_), st(st_) {
    if (!c->prof_on) return;
    stnerf_ctx::ProfRec r{cls, nullptr, nullptr, points, count_slot, S};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    c->prof.push_back(r);
    idx = (int)c->prof.size() - 1;
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(c->prof[idx].b, st); }
};
}  // namespace

static void free_ws(stnerf_ctx* c) {
  cudaFree(c->t_coarse); cudaFree(c->raw_coarse); cudaFree(c->t_fine); cudaFree(c->raw_fine); cudaFree(c->xyz);
  cudaFree(c->xyz_coarse); cudaFree(c->z_new); cudaFree(c->src_map);
  c->xyz_coarse = c->z_new = nullptr; c->src_map = nullptr;
  cudaFree(c->cbuf); c->cbuf = nullptr;
  cudaFree(c->mask_ws); cudaFree(c->hit); cudaFree(c->counts); cudaFree(c->lerp_flags);
  c->t_coarse = c->raw_coarse = c->t_fine = c->raw_fine = c->xyz = nullptr;
  c->mask_ws = nullptr; c->hit = c->counts = c->lerp_flags = nullptr;
  c->ws_bytes = 0; c->cap_n1 = c->cap_s2 = 0;
}

static int ensure_ws(stnerf_ctx* c, int n1, int s2) {
  if (n1 <= c->cap_n1 && s2 <= c->cap_s2 && c->t_coarse) return STNERF_OK;
  const int cn1 = std::max(n1, c->cap_n1), cs2 = std::max(s2, c->cap_s2);
  free_ws(c);
  const size_t R = (size_t)c->chunk_rays, l = (size_t)c->l;
  size_t tot = 0;
  auto A = [&](void** p, size_t bytes) -> int {
    if (cudaMalloc(p, bytes) != cudaSuccess) { cudaGetLastError(); return STNERF_ENOMEM; }
    tot += bytes;
    return STNERF_OK;
  };
  int rc = 0;
  rc |= A((void**)&c->t_coarse, l * R * cn1 * 4);
  rc |= A((void**)&c->raw_coarse, l * R * cn1 * 16);
  rc |= A((void**)&c->t_fine, l * R * cs2 * 4);
  rc |= A((void**)&c->raw_fine, l * R * cs2 * 16);
  rc |= A((void**)&c->xyz, R * cs2 * 12);
  rc |= A((void**)&c->xyz_coarse, l * R * cn1 * 12);
  rc |= A((void**)&c->z_new, l * R * cs2 * 4);
  rc |= A((void**)&c->src_map, l * R * cs2);
  rc |= A((void**)&c->cbuf, R * 128 * 4);
  rc |= A((void**)&c->mask_ws, l * R);
  rc |= A((void**)&c->hit, l * R * 4);
  rc |= A((void**)&c->counts, STNERF_MAX_LAYERS * 4);
  rc |= A((void**)&c->lerp_flags, STNERF_MAX_LAYERS * 4);
  if (rc) { free_ws(c); return STNERF_ENOMEM; }
  c->cap_n1 = cn1; c->cap_s2 = cs2; c->ws_bytes = tot;
  return STNERF_OK;
}

// ---------------------------------------------------------------------------------------------------------
// weight packing (host): state_dict order blob -> transposed [k][n] fp32 for the SIMT kernels
// ---------------------------------------------------------------------------------------------------------
static void transpose_into(std::vector<float>& dst, const float* W, int N, int K) {   // W (N,K) -> [K][N]
  const size_t base = dst.size();
  dst.resize(base + (size_t)N * K);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) dst[base + (size_t)k * N + n] = W[(size_t)n * K + k];
}

static int upload(float** dev, const std::vector<float>& host) {
  if (*dev) { cudaFree(*dev); *dev = nullptr; }
  STNERF_CUDA(cudaMalloc((void**)dev, host.size() * sizeof(float)));
  STNERF_CUDA(cudaMemcpy(*dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
  return STNERF_OK;
}

extern "C" {

const char* stnerf_strerror(int code) {
  switch (code) {
    case STNERF_OK: return "ok";
    case STNERF_EINVAL: return "invalid argument";
    case STNERF_ENODEVICE: return "no usable CUDA device (needs sm_100)";
    case STNERF_ECUDA: return "CUDA runtime error (see stnerf_last_cuda_error)";
    case STNERF_ENOWEIGHTS: return "network weights not loaded";
    case STNERF_ENOMEM: return "out of device memory";
    default: return "unknown error";
  }
}
const char* stnerf_last_cuda_error(void) { return g_cuda_err; }
uint64_t stnerf_launch_count(void) { return g_launches.load(); }

int stnerf_create(stnerf_handle* out, const stnerf_model_desc* d) {
  if (!out || !d || d->n_layers < 2 || d->n_layers > STNERF_MAX_LAYERS) return STNERF_EINVAL;
  if (d->precision < 0 || d->precision > STNERF_PREC_TC_3XF16_CF || d->chunk_rays < 0) return STNERF_EINVAL;
  // the kernels index samples of a chunk with 32-bit integers: chunk_rays * STNERF_MAX_S must stay below 2^31
  if ((long long)d->chunk_rays * STNERF_MAX_S >= (1LL << 31)) return STNERF_EINVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return STNERF_ENODEVICE; }
  int dev = 0;
  STNERF_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  STNERF_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return STNERF_ENODEVICE;     // the cubin is sm_100a only
  stnerf_ctx* c = new (std::nothrow) stnerf_ctx();
  if (!c) return STNERF_ENOMEM;
  c->desc = *d;
  c->l = d->n_layers;
  c->device = dev;
  c->num_sms = prop.multiProcessorCount;
  c->precision = d->precision;
  c->chunk_rays = d->chunk_rays > 0 ? d->chunk_rays : 65536;
  if (const char* e = getenv("STNERF_NO_FUSE")) c->no_fuse = (e[0] == '1');
  if (const char* e = getenv("STNERF_LO_FIRST")) c->lo_first_env = (e[0] == '1') ? 1 : (e[0] == '2') ? 2 : 0;
  if (const char* e = getenv("STNERF_NO_REUSE")) c->no_reuse = (e[0] == '1');
  if (cudaMalloc((void**)&c->any_frac, 4) != cudaSuccess) { delete c; return STNERF_ENOMEM; }
  *out = c;
  return STNERF_OK;
}

void stnerf_destroy(stnerf_handle c) {
  if (!c) return;
  cudaDeviceSynchronize();
  free_ws(c);
  for (int f = 0; f < 2; ++f)
    for (int i = 0; i < STNERF_MAX_LAYERS; ++i) { cudaFree(c->space[f][i].blob); tc_free(c->space[f][i].tc); }
  for (int i = 0; i < STNERF_MAX_LAYERS; ++i) { cudaFree(c->motion[i].blob); tc_free(c->motion[i].tc); }
  cudaFree(c->h_rays); cudaFree(c->h_out); cudaFree(c->h_mask); cudaFree(c->any_frac);
  cudaFree(c->v_rays); cudaFree(c->v_img[0]); cudaFree(c->v_img[1]); cudaFree(c->box_table);
  if (c->copy_in) cudaStreamDestroy(c->copy_in);
  if (c->copy_out) cudaStreamDestroy(c->copy_out);
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  if (c->prof_counts) cudaFreeHost(c->prof_counts);
  delete c;
}

int stnerf_set_precision(stnerf_handle c, int precision) {
  if (!c || precision < 0 || precision > STNERF_PREC_TC_3XF16_CF) return STNERF_EINVAL;
  c->precision = precision;
  return STNERF_OK;
}

int stnerf_reserve(stnerf_handle c, int n1, int n2) {
  if (!c || n1 < 3 || n1 > STNERF_MAX_N1 || n2 < 0 || n1 + n2 > STNERF_MAX_S) return STNERF_EINVAL;
  return ensure_ws(c, n1, n1 + n2);
}
size_t stnerf_workspace_bytes(stnerf_handle c) { return c ? c->ws_bytes : 0; }

// fp32 (SIMT) device image of a SpaceNet: [w_i^T | b_i] x 7, w_sigma, w_rgbh^T, b_rgbh, w_rgbo; the scalar biases live in SpaceNetW
static size_t bind_spacenet(SpaceNetDev& N, bool use_time) {
  const int krgb = HID + PE_DIR + (use_time ? PE_TIME : 0);
  const int Ks[7] = {PE_POS, HID, HID, HID, HID + PE_POS, HID, HID};
  size_t cur = 0;
  for (int i = 0; i < 7; ++i) {
    N.w.w[i] = N.blob + cur; cur += (size_t)HID * Ks[i];
    N.w.b[i] = N.blob + cur; cur += HID;
  }
  N.w.w_sigma = N.blob + cur; cur += HID;
  N.w.w_rgbh = N.blob + cur; cur += (size_t)HEAD * krgb;
  N.w.b_rgbh = N.blob + cur; cur += HEAD;
  N.w.w_rgbo = N.blob + cur; cur += 3 * HEAD;
  N.w.use_time = use_time ? 1 : 0;
  return cur;
}
static size_t bind_motionnet(MotionNetDev& N) {
  size_t cur = 0;
  for (int i = 0; i < 5; ++i) {
    const int K = i == 0 ? PE_MOTION : HEAD;
    N.w.w[i] = N.blob + cur; cur += (size_t)HEAD * K;
    N.w.b[i] = N.blob + cur; cur += HEAD;
  }
  N.w.w_out = N.blob + cur; cur += 3 * HEAD;
  return cur;
}

int stnerf_load_spacenet(stnerf_handle c, int layer, int fine, const float* blob, size_t n) {
  if (!c || !blob || layer < 0 || layer >= c->l || (fine != 0 && fine != 1)) return STNERF_EINVAL;
  const bool use_time = (n == (size_t)SPACENET_FLOATS_TIME);
  if (!use_time && n != (size_t)SPACENET_FLOATS_NOTIME) return STNERF_EINVAL;
  if ((c->desc.space_time[layer] != 0) != use_time) return STNERF_EINVAL;
  const int krgb = HID + PE_DIR + (use_time ? PE_TIME : 0);
  // walk the blob in state_dict order
  const float* p = blob;
  std::vector<float> host;
  const int Ks[7] = {PE_POS, HID, HID, HID, HID + PE_POS, HID, HID};
  for (int i = 0; i < 7; ++i) {
    transpose_into(host, p, HID, Ks[i]);
    p += (size_t)HID * Ks[i];
    host.insert(host.end(), p, p + HID);
    p += HID;
  }
  host.insert(host.end(), p, p + HID); p += HID;
  const float b_sigma = *p++;
  transpose_into(host, p, HEAD, krgb); p += (size_t)HEAD * krgb;
  host.insert(host.end(), p, p + HEAD); p += HEAD;
  host.insert(host.end(), p, p + 3 * HEAD); p += 3 * HEAD;
  const float b_o[3] = {p[0], p[1], p[2]};
  SpaceNetDev& N = c->space[fine][layer];
  N.loaded = false;
  int rc = upload(&N.blob, host);
  if (rc) return rc;
  if (bind_spacenet(N, use_time) != host.size()) return STNERF_EINVAL;
  N.w.b_sigma = b_sigma;
  for (int a = 0; a < 3; ++a) N.w.b_rgbo[a] = b_o[a];
  rc = tc_pack_spacenet(N.tc, blob, use_time);
  if (rc) return rc;
  N.loaded = true;
  return STNERF_OK;
}

int stnerf_load_motionnet(stnerf_handle c, int layer, const float* blob, size_t n) {
  if (!c || !blob || layer < 1 || layer >= c->l || n != (size_t)MOTIONNET_FLOATS) return STNERF_EINVAL;
  const float* p = blob;
  std::vector<float> host;
  for (int i = 0; i < 5; ++i) {
    const int K = i == 0 ? PE_MOTION : HEAD;
    transpose_into(host, p, HEAD, K); p += (size_t)HEAD * K;
    host.insert(host.end(), p, p + HEAD); p += HEAD;
  }
  host.insert(host.end(), p, p + 3 * HEAD); p += 3 * HEAD;
  MotionNetDev& N = c->motion[layer];
  N.loaded = false;
  int rc = upload(&N.blob, host);
  if (rc) return rc;
  if (bind_motionnet(N) != host.size()) return STNERF_EINVAL;
  for (int a = 0; a < 3; ++a) N.w.b_out[a] = p[a];
  rc = tc_pack_motionnet(N.tc, blob);
  if (rc) return rc;
  N.loaded = true;
  return STNERF_OK;
}

// ---- packed-weight image: every loaded network's device buffers, as they are, behind a small header ----
namespace {
constexpr char PACK_MAGIC[8] = {'S', 'T', 'N', 'B', '2', '0', '0', 'W'};
constexpr uint32_t PACK_VERSION = 3;      // 2: weight stream = correction section + main section per layer; 3: skip layer's encoding chunk first
struct PackHeader { char magic[8]; uint32_t version, n_layers, n_records, reserved; };
struct PackRec { uint32_t kind, fine, layer, use_time; uint64_t simt_floats, stream_bytes, aux_floats, tail_floats; float scalars[4]; uint32_t pad[4]; };
static size_t rec_payload(const PackRec& r) { return r.simt_floats * 4 + r.stream_bytes + r.aux_floats * 4 + r.tail_floats * 4; }
static size_t simt_floats_space(bool use_time) { return (size_t)(use_time ? SPACENET_FLOATS_TIME : SPACENET_FLOATS_NOTIME) - 4; }
static size_t simt_floats_motion() { return (size_t)MOTIONNET_FLOATS - 3; }
}  // namespace

int stnerf_weights_export(stnerf_handle c, void* buf, size_t capacity, size_t* bytes_needed) {
  if (!c) return STNERF_EINVAL;
  std::vector<PackRec> recs;
  for (int f = 0; f < 2; ++f)
    for (int i = 0; i < c->l; ++i)
      if (c->space[f][i].loaded) {
        const SpaceNetDev& N = c->space[f][i];
        PackRec r{};
        r.kind = 0; r.fine = (uint32_t)f; r.layer = (uint32_t)i; r.use_time = (uint32_t)N.w.use_time;
        r.simt_floats = simt_floats_space(N.w.use_time != 0); r.stream_bytes = tc_stream_bytes(true);
        r.aux_floats = tc_aux_floats(); r.tail_floats = tc_tail_floats(true);
        r.scalars[0] = N.w.b_sigma; for (int a = 0; a < 3; ++a) r.scalars[1 + a] = N.w.b_rgbo[a];
        recs.push_back(r);
      }
  for (int i = 1; i < c->l; ++i)
    if (c->motion[i].loaded) {
      PackRec r{};
      r.kind = 1; r.layer = (uint32_t)i;
      r.simt_floats = simt_floats_motion(); r.stream_bytes = tc_stream_bytes(false);
      r.aux_floats = tc_aux_floats(); r.tail_floats = 0;
      for (int a = 0; a < 3; ++a) r.scalars[a] = c->motion[i].w.b_out[a];
      recs.push_back(r);
    }
  if (recs.empty()) return STNERF_ENOWEIGHTS;
  size_t need = sizeof(PackHeader);
  for (const PackRec& r : recs) need += sizeof(PackRec) + rec_payload(r);
  if (bytes_needed) *bytes_needed = need;
  if (!buf) return STNERF_OK;                         // size query
  if (capacity < need) return STNERF_EINVAL;
  uint8_t* p = static_cast<uint8_t*>(buf);
  PackHeader h{};
  memcpy(h.magic, PACK_MAGIC, 8); h.version = PACK_VERSION; h.n_layers = (uint32_t)c->l; h.n_records = (uint32_t)recs.size();
  memcpy(p, &h, sizeof(h)); p += sizeof(h);
  for (const PackRec& r : recs) {
    memcpy(p, &r, sizeof(r)); p += sizeof(r);
    float* simt = reinterpret_cast<float*>(p);
    uint8_t* stream = p + r.simt_floats * 4;
    float* aux = reinterpret_cast<float*>(stream + r.stream_bytes);
    float* tail = aux + r.aux_floats;
    const float* dev = r.kind == 0 ? c->space[r.fine][r.layer].blob : c->motion[r.layer].blob;
    const TcNet& tc = r.kind == 0 ? c->space[r.fine][r.layer].tc : c->motion[r.layer].tc;
    STNERF_CUDA(cudaMemcpy(simt, dev, r.simt_floats * 4, cudaMemcpyDeviceToHost));
    const int rc = tc_export(tc, r.kind == 0, stream, aux, tail);
    if (rc) return rc;
    p += rec_payload(r);
  }
  return STNERF_OK;
}

int stnerf_weights_import(stnerf_handle c, const void* buf, size_t bytes) {
  if (!c || !buf || bytes < sizeof(PackHeader)) return STNERF_EINVAL;
  const uint8_t* p = static_cast<const uint8_t*>(buf);
  const uint8_t* end = p + bytes;
  PackHeader h;
  memcpy(&h, p, sizeof(h)); p += sizeof(h);
  if (memcmp(h.magic, PACK_MAGIC, 8) != 0 || h.version != PACK_VERSION || (int)h.n_layers != c->l) return STNERF_EINVAL;
  // validate every record against this context before touching any network
  const uint8_t* q = p;
  for (uint32_t k = 0; k < h.n_records; ++k) {
    if ((size_t)(end - q) < sizeof(PackRec)) return STNERF_EINVAL;
    PackRec r;
    memcpy(&r, q, sizeof(r)); q += sizeof(r);
    if (r.kind > 1 || r.fine > 1 || (int)r.layer >= c->l || r.aux_floats != tc_aux_floats()) return STNERF_EINVAL;
    if (r.kind == 0) {
      if ((c->desc.space_time[r.layer] != 0) != (r.use_time != 0)) return STNERF_EINVAL;
      if (r.simt_floats != simt_floats_space(r.use_time != 0) || r.stream_bytes != tc_stream_bytes(true) ||
          r.tail_floats != tc_tail_floats(true)) return STNERF_EINVAL;
    } else {
      if (r.layer < 1 || r.simt_floats != simt_floats_motion() || r.stream_bytes != tc_stream_bytes(false) || r.tail_floats != 0)
        return STNERF_EINVAL;
    }
    if ((size_t)(end - q) < rec_payload(r)) return STNERF_EINVAL;
    q += rec_payload(r);
  }
  if (q != end) return STNERF_EINVAL;
  for (uint32_t k = 0; k < h.n_records; ++k) {
    PackRec r;
    memcpy(&r, p, sizeof(r)); p += sizeof(r);
    const float* simt = reinterpret_cast<const float*>(p);
    const uint8_t* stream = p + r.simt_floats * 4;
    const float* aux = reinterpret_cast<const float*>(stream + r.stream_bytes);
    const float* tail = aux + r.aux_floats;
    const std::vector<float> host(simt, simt + r.simt_floats);
    if (r.kind == 0) {
      SpaceNetDev& N = c->space[r.fine][r.layer];
      N.loaded = false;
      int rc = upload(&N.blob, host);
      if (rc) return rc;
      bind_spacenet(N, r.use_time != 0);
      N.w.b_sigma = r.scalars[0];
      for (int a = 0; a < 3; ++a) N.w.b_rgbo[a] = r.scalars[1 + a];
      rc = tc_import(N.tc, true, (int)r.use_time, stream, aux, tail);
      if (rc) return rc;
      N.loaded = true;
    } else {
      MotionNetDev& N = c->motion[r.layer];
      N.loaded = false;
      int rc = upload(&N.blob, host);
      if (rc) return rc;
      bind_motionnet(N);
      for (int a = 0; a < 3; ++a) N.w.b_out[a] = r.scalars[a];
      rc = tc_import(N.tc, false, 0, stream, aux, nullptr);
      if (rc) return rc;
      N.loaded = true;
    }
    p += rec_payload(r);
  }
  return STNERF_OK;
}

int stnerf_set_scene(stnerf_handle c, const stnerf_scene* s) {
  if (!c || !s) return STNERF_EINVAL;
  c->scene = *s;
  DevScene& d = c->dscene;
  memset(&d, 0, sizeof(d));
  for (int i = 0; i < c->l; ++i) {
    for (int a = 0; a < 3; ++a) { d.bmin[i][a] = s->bmin[i][a]; d.bmax[i][a] = s->bmax[i][a]; }
    d.shown[i] = s->shown[i];
    if (s->scale_coarse_on[i] || s->scale_fine_on[i]) {
      if (s->scale[i] == 0.0f) return STNERF_EINVAL;
    }
  }
  d.near_plane = s->near_plane; d.alpha2 = s->alpha_layer2;
  d.thr_layer = s->density_threshold; d.thr_bkgd = s->bkgd_density_threshold;
  d.boarder = s->boarder_weight; d.apply_thr = s->apply_thresholds; d.n_layers = c->l;
  d.fid_shared = s->shared_frame_id;
  c->have_scene = true;
  return STNERF_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// one pass of the networks over a chunk
// ---------------------------------------------------------------------------------------------------------
static void fill_edit(PointSrc& s, const stnerf_scene& sc, int layer, bool fine) {
  s.shift_on = sc.shift_on[layer];
  s.scale_on = fine ? sc.scale_fine_on[layer] : sc.scale_coarse_on[layer];
  for (int a = 0; a < 3; ++a) { s.shift[a] = sc.shift[layer][a]; s.pivot[a] = sc.pivot[a]; }
  s.scale = sc.scale[layer];
  if (!s.scale_on) s.scale = 1.0f;
}

static int run_spacenet(stnerf_ctx* c, const PointSrc& src, SpaceNetDev& net, float* raw, float* rgb, float* sigma,
                        cudaStream_t st, int count_slot = -1, const FuseCoarse* fuse = nullptr, bool fine_pass = false) {
  if (!net.loaded) return STNERF_ENOWEIGHTS;
  ProfScope ps(c, 0, (double)src.n_slots * src.S, count_slot, src.S, st);
  if (c->precision == STNERF_PREC_FP32_SIMT)
    return launch_spacenet_simt(src, net.w, raw, 0, rgb, sigma, c->num_sms, st);
  if (src.mode == SRC_EXPLICIT) {      // unit entry point: one bias row per point, stream-ordered scratch
    float* cb = nullptr;
    STNERF_CUDA(cudaMallocAsync((void**)&cb, (size_t)std::max<long long>(src.n_slots, 1) * 128 * sizeof(float), st));
    const int rc = tc_launch_spacenet(src, net.tc, net.w, c->precision, cb, raw, rgb, sigma, c->num_sms, st, nullptr, c->lo_first_mode() != 0);
    STNERF_CUDA(cudaFreeAsync(cb, st));
    return rc;
  }
  const int lo_first = c->lo_first_mode() == 2 || (c->lo_first_mode() == 1 && !fine_pass);
  return tc_launch_spacenet(src, net.tc, net.w, c->precision, c->cbuf, raw, rgb, sigma, c->num_sms, st, fuse, lo_first);
}
static int run_motionnet(stnerf_ctx* c, const PointSrc& src, MotionNetDev& net, const int* lerp_flag, int lerp_force,
                         float* xyz_out, float* flow_out, cudaStream_t st, int count_slot = -1) {
  if (!net.loaded) return STNERF_ENOWEIGHTS;
  ProfScope ps(c, 1, (double)src.n_slots * src.S, count_slot, src.S, st);
  if (c->precision == STNERF_PREC_FP32_SIMT)
    return launch_motionnet_simt(src, net.w, lerp_flag, lerp_force, xyz_out, flow_out, c->num_sms, st);
  return tc_launch_motionnet(src, net.tc, net.w, c->precision, lerp_flag, lerp_force, xyz_out, flow_out, c->num_sms, st,
                             c->lo_first_mode() != 0);
}

// `fuse` (coarse pass only): template of the per-layer fusion request (everything but the layer-specific fields), or null.
// `want_raw`: the (rgb, sigma) samples must reach HBM (a later kernel composites them); false only with `fuse`.
// `reuse_n1` > 0 (fine pass): the MotionNet runs on the S - reuse_n1 NEW depths only; the flow of the coarse depths is the coarse
// pass' (xyz_coarse), stitched together through the origin map the merge wrote (z_new / src_map).
static int run_nets(stnerf_ctx* c, const float* rays, long long n, int ray_stride, bool fine, int S, cudaStream_t st,
                    int chunk_slot, const FuseCoarse* fuse = nullptr, bool want_raw = true, float* coarse_imgs = nullptr,
                    long long plane = 0, int reuse_n1 = 0) {
  const long long R = c->chunk_rays;
  const float* tbuf = fine ? c->t_fine : c->t_coarse;
  float* rawbuf = fine ? c->raw_fine : c->raw_coarse;
  const long long tl = R * (fine ? c->cap_s2 : c->cap_n1);      // layer strides of the workspace arrays
  for (int i = 0; i < c->l; ++i) {
    if (i > 0 && !c->scene.shown[i]) continue;                  // hidden performers: sigma = rgb = 0 (:401, :556)
    PointSrc s;
    memset(&s, 0, sizeof(s));
    s.mode = SRC_MARCH;
    s.rays = rays; s.ray_stride = ray_stride;
    s.t = tbuf + i * tl;
    s.S = S; s.layer = c->scene.shared_frame_id ? 0 : i;      // frame-id column offset of this layer
    s.pos_stride = 3; s.time_stride = 1;
    fill_edit(s, c->scene, i, fine);
    float* raw = want_raw ? rawbuf + (size_t)i * tl * 4 : nullptr;
    FuseCoarse f;
    memset(&f, 0, sizeof(f));
    if (fuse) {
      f = *fuse;
      f.on = 1; f.layer = i; f.is_bkgd = (i == 0);
      f.t_fine = c->t_fine + (size_t)i * R * c->cap_s2;
      f.z_new = fuse->z_new ? c->z_new + (size_t)i * R * c->cap_s2 : nullptr;
      f.src_map = fuse->z_new ? c->src_map + (size_t)i * R * c->cap_s2 : nullptr;
      f.u = fuse->u ? fuse->u + (size_t)i * fuse->n_total * fuse->n2 : nullptr;     // [layer][ray of the call][n2], chunk offset applied by the caller
      f.img = coarse_imgs ? coarse_imgs + (size_t)(1 + i) * plane : nullptr;
    }
    int rc;
    if (i == 0) {
      s.hit = nullptr; s.count = nullptr; s.n_slots = n;
      rc = run_spacenet(c, s, c->space[fine ? 1 : 0][0], raw, nullptr, nullptr, st, -1, fuse ? &f : nullptr, fine);
      if (rc) return rc;
      continue;
    }
    s.hit = c->hit + (size_t)i * R;
    s.count = c->counts + i;
    s.n_slots_cap = n;
    // the coarse pass keeps every performer's deformed points (the fine pass may reuse them); the fine pass has one scratch
    float* xyz_out = fine ? c->xyz : c->xyz_coarse + (size_t)i * R * c->cap_n1 * 3;
    // reuse needs the same inverse edit in both passes (a None shift entry skips the fine-pass scale only, layered_rfrender.py:468-469)
    const bool reuse = fine && reuse_n1 > 0 && c->scene.scale_fine_on[i] == c->scene.scale_coarse_on[i];
    if (reuse) {
      PointSrc m = s;                                            // MotionNet on the n2 new depths of every hit ray
      m.t = c->z_new + (size_t)i * R * c->cap_s2;
      m.S = S - reuse_n1;
      rc = run_motionnet(c, m, c->motion[i], c->lerp_flags + i, -1, xyz_out, nullptr, st, chunk_slot >= 0 ? chunk_slot * 8 + i : -1);
      if (rc) return rc;
      s.mode = SRC_XYZ_MAP;
      s.pos = c->xyz_coarse + (size_t)i * R * c->cap_n1 * 3;
      s.pos2 = xyz_out;
      s.src_map = c->src_map + (size_t)i * R * c->cap_s2;
      s.n_first = reuse_n1;
      rc = run_spacenet(c, s, c->space[1][i], raw, nullptr, nullptr, st, chunk_slot >= 0 ? chunk_slot * 8 + i : -1, nullptr, true);
      if (rc) return rc;
      continue;
    }
    rc = run_motionnet(c, s, c->motion[i], c->lerp_flags + i, -1, xyz_out, nullptr, st, chunk_slot >= 0 ? chunk_slot * 8 + i : -1);
    if (rc) return rc;
    s.mode = SRC_XYZ;
    s.pos = xyz_out;
    rc = run_spacenet(c, s, c->space[fine ? 1 : 0][i], raw, nullptr, nullptr, st, chunk_slot >= 0 ? chunk_slot * 8 + i : -1,
                      fuse ? &f : nullptr, fine);
    if (rc) return rc;
  }
  return STNERF_OK;
}

// Where the images of a call go.  `coarse` / `fine`: [l+1 images][5*n_total floats] each (image 0 mixed, 1+i layer i), or null =
// that pass writes no images (the coarse pass then only resamples: its merged composite is skipped altogether).
struct OutSpec {
  float* coarse;
  float* fine;
  int pixels;            // 0: image = rgb (N,3) | depth (N) | acc (N) planes;  1: image = (N,5) pixel-interleaved
};

// Called after the last kernel of a chunk has been enqueued (pipelined host copies hang their events here).
struct ChunkHook {
  int (*fn)(void* user, long long c0, long long n, cudaStream_t st);
  void* user;
};

static int render_core(stnerf_ctx* c, const float* rays, long long n_rays, int ray_stride, int n1, int n2, int only_coarse,
                       const float* jitter, const float* u, uint64_t seed, OutSpec out, uint8_t* ray_mask, cudaStream_t st,
                       const ChunkHook* before_chunk = nullptr, const ChunkHook* after_chunk = nullptr) {
  if (!c || !rays || n_rays < 0) return STNERF_EINVAL;
  if (!c->have_scene) return STNERF_EINVAL;
  {                                         // the context's weights and workspace live on the device it was created on
    int cur = -1;
    STNERF_CUDA(cudaGetDevice(&cur));
    if (cur != c->device) return STNERF_EINVAL;
  }
  if (ray_stride < 6 + (c->scene.shared_frame_id ? 1 : c->l)) return STNERF_EINVAL;   // the reference prints + exit(-1) (:162-163)
  if (n1 < 3 || n1 > STNERF_MAX_N1) return STNERF_EINVAL;
  if (only_coarse) n2 = 0;
  if (n2 < 0 || n1 + n2 > STNERF_MAX_S) return STNERF_EINVAL;
  if (n2 == 0 && !out.coarse) return STNERF_EINVAL;
  if (n2 > 0 && !out.fine) return STNERF_EINVAL;
  int rc = ensure_ws(c, n1, n1 + n2);
  if (rc) return rc;
  const int l = c->l, S2 = n1 + n2;
  const long long R = c->chunk_rays, N = n_rays;
  for (long long c0 = 0; c0 < N; c0 += R) {
    const long long n = std::min(R, N - c0);
    if (before_chunk) { rc = before_chunk->fn(before_chunk->user, c0, n, st); if (rc) return rc; }
    c->last_chunk_rays = n; c->last_n1 = n1; c->last_s2 = n2 > 0 ? S2 : 0;
    const float* rch = rays + c0 * ray_stride;
    STNERF_CUDA(cudaMemsetAsync(c->counts, 0, STNERF_MAX_LAYERS * 4, st));
    STNERF_CUDA(cudaMemsetAsync(c->lerp_flags, 0, STNERF_MAX_LAYERS * 4, st));
    uint8_t* mask = ray_mask ? ray_mask + c0 : c->mask_ws;
    const long long mask_ls = ray_mask ? N : R;
    int chunk_slot = -1;
    {
      ProfScope ps(c, 2, (double)n, -1, 1, st);
      rc = launch_sample(rch, n, ray_stride, c->dscene, l, n1, jitter ? jitter + c0 * n1 : nullptr, N * n1, seed, c0, c->idmap,
                         c->t_coarse, R * c->cap_n1, mask, mask_ls, c->hit, R, c->counts, c->lerp_flags, st, c->box_table, c->box_frames);
      if (rc) return rc;
    }
    if (c->prof_on && c->prof_counts && c->prof_chunks < PROF_MAX_CHUNKS) {
      chunk_slot = c->prof_chunks++;
      STNERF_CUDA(cudaMemcpyAsync(c->prof_counts + chunk_slot * 8, c->counts, 32, cudaMemcpyDeviceToHost, st));
    }
    // NOTE: the workspace t arrays are laid out with the *capacity* sample counts as layer stride but the
    // per-ray stride is the live n1 / S2, so a ray's samples stay contiguous.
    // Coarse-pass fusion: with a tensor-core mode and n1 = 64 the SpaceNet kernel composites and resamples every shown layer in
    // its spare warps (mlp_tc.cuh: FuseCoarse).  The stand-alone kernel is then only needed for the merged coarse image (when
    // coarse images are wanted), for the zero pixels of missed rays, and for the resampling of hit-but-hidden layers.
    const bool fuse = n2 > 0 && c->precision != STNERF_PREC_FP32_SIMT && tc_can_fuse_coarse(n1, n2) && !c->no_fuse;
    unsigned fused_layers = 0;
    bool hidden_any = false;
    for (int i = 0; i < l; ++i) {
      if (i == 0 || c->scene.shown[i]) fused_layers |= fuse ? (1u << i) : 0u;
      else hidden_any = true;
    }
    // Flow reuse: the merge (fused or stand-alone register path) also writes the new depths and the origin of every fine depth, so
    // the fine pass runs the MotionNet on the n2 new depths only (same network, same points for the other n1: SURVEY A.6).
    const bool reuse = n2 > 0 && c->precision != STNERF_PREC_FP32_SIMT && n1 <= 128 && n2 <= 256 && S2 <= 256 && !c->no_reuse;
    FuseCoarse ft;
    memset(&ft, 0, sizeof(ft));
    if (fuse) {
      ft.z_new = reuse ? c->z_new : nullptr;
      ft.n1 = n1; ft.n2 = n2; ft.u = u ? u + c0 * n2 : nullptr; ft.seed = seed; ft.idmap = c->idmap; ft.ray_base = c0;
      ft.n_total = N; ft.pixels = out.pixels; ft.near_plane = c->dscene.near_plane; ft.thr = c->dscene.thr_layer;
      ft.boarder = c->dscene.boarder; ft.apply_thr = c->dscene.apply_thr;
    }
    rc = run_nets(c, rch, n, ray_stride, false, n1, st, chunk_slot, fuse ? &ft : nullptr, /*want_raw=*/!fuse || out.coarse != nullptr,
                  out.coarse, 5 * N);
    if (rc) return rc;
    CompositeArgs a;
    memset(&a, 0, sizeof(a));
    a.skip_layers = fused_layers;
    a.t = c->t_coarse; a.t_layer_stride = R * c->cap_n1;
    a.raw = c->raw_coarse; a.raw_layer_stride = R * c->cap_n1 * 4;
    a.mask = mask; a.mask_layer_stride = mask_ls;
    a.u = u ? u + c0 * n2 : nullptr; a.u_layer_stride = N * n2;
    a.t_fine = c->t_fine; a.tf_layer_stride = R * c->cap_s2;
    if (reuse) { a.z_new = c->z_new; a.zn_layer_stride = R * c->cap_s2; a.src_map = c->src_map; a.sm_layer_stride = R * c->cap_s2; }
    a.out = out.coarse; a.pixel_layout = out.pixels; a.n_total = N; a.ray_base = c0; a.n = n;
    a.S = n1; a.n2 = n2; a.fine = 0; a.seed = seed; a.idmap = c->idmap;
    if (!fuse || out.coarse != nullptr || hidden_any) {
      ProfScope ps(c, 3, (double)n, -1, 1, st);
      rc = launch_composite_pass(a, c->dscene, l, st);
      if (rc) return rc;
    }
    if (n2 > 0) {
      rc = run_nets(c, rch, n, ray_stride, true, S2, st, chunk_slot, nullptr, true, nullptr, 0, reuse ? n1 : 0);
      if (rc) return rc;
      a.t = c->t_fine; a.t_layer_stride = R * c->cap_s2;
      a.raw = c->raw_fine; a.raw_layer_stride = R * c->cap_s2 * 4;
      a.u = nullptr; a.t_fine = nullptr;
      a.out = out.fine;
      a.S = S2; a.n2 = 0; a.fine = 1; a.skip_layers = 0;
      ProfScope ps(c, 3, (double)n, -1, 1, st);
      rc = launch_composite_pass(a, c->dscene, l, st);
      if (rc) return rc;
    }
    if (after_chunk) { rc = after_chunk->fn(after_chunk->user, c0, n, st); if (rc) return rc; }
  }
  return STNERF_OK;
}

extern "C" {

int stnerf_render(stnerf_handle c, const float* rays, int64_t n_rays, int ray_stride, int n1, int n2, int only_coarse,
                  const float* jitter, const float* u, uint64_t seed, float* out, uint8_t* ray_mask, void* stream) {
  if (!c || !out) return STNERF_EINVAL;
  OutSpec o{out, out + (size_t)(c->l + 1) * 5 * (size_t)n_rays, 0};
  return render_core(c, rays, n_rays, ray_stride, n1, n2, only_coarse, jitter, u, seed, o, ray_mask, (cudaStream_t)stream);
}

static int grow(void** p, size_t* have, size_t want) {
  if (want <= *have) return STNERF_OK;
  if (*p) cudaFree(*p);
  *p = nullptr; *have = 0;
  if (cudaMalloc(p, want) != cudaSuccess) { cudaGetLastError(); return STNERF_ENOMEM; }
  *have = want;
  return STNERF_OK;
}

}  // extern "C"

// ---- host-buffer plumbing: copy streams, an event pool, per-chunk hooks ------------------------------------------------------
static int ensure_copy_streams(stnerf_ctx* c) {
  if (!c->copy_in) STNERF_CUDA(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
  if (!c->copy_out) STNERF_CUDA(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
  return STNERF_OK;
}
static int ensure_events(stnerf_ctx* c, size_t n) {
  while (c->ev_pool.size() < n) {
    cudaEvent_t e;
    STNERF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ev_pool.push_back(e);
  }
  return STNERF_OK;
}

namespace {
// stnerf_render_host: rays arrive chunk by chunk on `copy_in` while earlier chunks render; the image slices of a finished chunk
// leave on `copy_out` while later chunks render.
struct HostPipe {
  stnerf_ctx* c;
  long long N;
  int only_coarse;
  float* out_host;
  uint8_t* mask_host;
  size_t ev_in0, ev_out0;      // first event index of the per-chunk "rays landed" / "chunk rendered" events
};
int host_before_chunk(void* user, long long c0, long long, cudaStream_t st) {
  HostPipe* p = static_cast<HostPipe*>(user);
  const size_t k = (size_t)(c0 / p->c->chunk_rays);
  STNERF_CUDA(cudaStreamWaitEvent(st, p->c->ev_pool[p->ev_in0 + k], 0));
  return STNERF_OK;
}
int host_after_chunk(void* user, long long c0, long long n, cudaStream_t st) {
  HostPipe* p = static_cast<HostPipe*>(user);
  stnerf_ctx* c = p->c;
  const size_t k = (size_t)(c0 / c->chunk_rays);
  cudaEvent_t done = c->ev_pool[p->ev_out0 + k];
  STNERF_CUDA(cudaEventRecord(done, st));
  STNERF_CUDA(cudaStreamWaitEvent(c->copy_out, done, 0));
  const size_t N = (size_t)p->N, plane = 5 * N;
  const int n_img = (p->only_coarse ? 1 : 2) * (c->l + 1);
  for (int im = 0; im < n_img; ++im) {      // plane = rgb (N,3) | depth (N) | acc (N): three contiguous slices per chunk
    const float* src = c->h_out + (size_t)im * plane;
    float* dst = p->out_host + (size_t)im * plane;
    STNERF_CUDA(cudaMemcpyAsync(dst + 3 * (size_t)c0, src + 3 * (size_t)c0, (size_t)n * 12, cudaMemcpyDeviceToHost, c->copy_out));
    STNERF_CUDA(cudaMemcpyAsync(dst + 3 * N + c0, src + 3 * N + c0, (size_t)n * 4, cudaMemcpyDeviceToHost, c->copy_out));
    STNERF_CUDA(cudaMemcpyAsync(dst + 4 * N + c0, src + 4 * N + c0, (size_t)n * 4, cudaMemcpyDeviceToHost, c->copy_out));
  }
  if (p->mask_host)
    for (int i = 0; i < c->l; ++i)
      STNERF_CUDA(cudaMemcpyAsync(p->mask_host + (size_t)i * N + c0, c->h_mask + (size_t)i * N + c0, (size_t)n,
                                  cudaMemcpyDeviceToHost, c->copy_out));
  return STNERF_OK;
}
}  // namespace

extern "C" {

int stnerf_reserve_host(stnerf_handle c, int64_t max_rays, int ray_stride) {
  if (!c || max_rays <= 0 || ray_stride < 6) return STNERF_EINVAL;
  const size_t n = (size_t)max_rays;
  int rc = grow((void**)&c->h_rays, &c->h_rays_bytes, n * ray_stride * 4);
  rc |= grow((void**)&c->h_out, &c->h_out_bytes, (size_t)2 * (c->l + 1) * 5 * n * 4);
  rc |= grow((void**)&c->h_mask, &c->h_mask_bytes, (size_t)c->l * n);
  rc |= grow((void**)&c->v_rays, &c->v_rays_bytes, n * (6 + STNERF_MAX_LAYERS) * 4);
  for (int b = 0; b < 2; ++b) rc |= grow((void**)&c->v_img[b], &c->v_img_bytes[b], (size_t)(c->l + 1) * 5 * n * 4);
  if (rc) return STNERF_ENOMEM;
  rc = ensure_copy_streams(c);
  if (rc) return rc;
  return ensure_events(c, 2 * (size_t)((max_rays + c->chunk_rays - 1) / c->chunk_rays) + 8);
}

int stnerf_render_host(stnerf_handle c, const float* rays_host, int64_t n_rays, int ray_stride, int n1, int n2,
                       int only_coarse, uint64_t seed, float* out_host, uint8_t* ray_mask_host, void* stream) {
  if (!c || !rays_host || !out_host || n_rays <= 0) return STNERF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rb = (size_t)n_rays * ray_stride * 4, ob = (size_t)2 * (c->l + 1) * 5 * n_rays * 4,
               mb = (size_t)c->l * n_rays;
  if (only_coarse) n2 = 0;
  // staging: sized by stnerf_reserve_host; grown here only when a call exceeds every earlier size
  int rc = grow((void**)&c->h_rays, &c->h_rays_bytes, rb);
  rc |= grow((void**)&c->h_out, &c->h_out_bytes, ob);
  rc |= grow((void**)&c->h_mask, &c->h_mask_bytes, mb);
  if (rc) return STNERF_ENOMEM;
  rc = ensure_copy_streams(c);
  if (rc) return rc;
  const long long R = c->chunk_rays;
  const size_t n_chunks = (size_t)((n_rays + R - 1) / R);
  rc = ensure_events(c, 2 * n_chunks + 1);
  if (rc) return rc;
  // the copy-in stream starts after everything already queued on `st` (the staging buffers may still be read by it)
  cudaEvent_t start = c->ev_pool[2 * n_chunks];
  STNERF_CUDA(cudaEventRecord(start, st));
  STNERF_CUDA(cudaStreamWaitEvent(c->copy_in, start, 0));
  STNERF_CUDA(cudaStreamWaitEvent(c->copy_out, start, 0));
  for (size_t k = 0; k < n_chunks; ++k) {
    const long long c0 = (long long)k * R, n = std::min<long long>(R, n_rays - c0);
    STNERF_CUDA(cudaMemcpyAsync(c->h_rays + (size_t)c0 * ray_stride, rays_host + (size_t)c0 * ray_stride, (size_t)n * ray_stride * 4,
                                cudaMemcpyHostToDevice, c->copy_in));
    STNERF_CUDA(cudaEventRecord(c->ev_pool[k], c->copy_in));
  }
  HostPipe pipe{c, n_rays, n2 == 0 ? 1 : 0, out_host, ray_mask_host, 0, n_chunks};
  ChunkHook before{host_before_chunk, &pipe}, after{host_after_chunk, &pipe};
  OutSpec o{c->h_out, c->h_out + (size_t)(c->l + 1) * 5 * (size_t)n_rays, 0};
  rc = render_core(c, c->h_rays, n_rays, ray_stride, n1, n2, only_coarse, nullptr, nullptr, seed, o, c->h_mask, st, &before, &after);
  // drain both streams even on an error so no copy is left in flight into the caller's buffers
  cudaError_t e1 = cudaStreamSynchronize(c->copy_out), e2 = cudaStreamSynchronize(st), e3 = cudaStreamSynchronize(c->copy_in);
  if (rc) return rc;
  STNERF_CUDA(e1); STNERF_CUDA(e2); STNERF_CUDA(e3);
  return STNERF_OK;
}

// One view of stnerf_render_views: rays of the requested rows generated on the device, scene constants of THIS view, fine images
// written pixel-interleaved to `images` ([l+1][n_rows*W][5]).  Enqueue only.
static int render_one_view(stnerf_ctx* c, const stnerf_view* v, int H, int W, int row0, int row_step, int n_rows, int n1, int n2,
                           float* images, float* coarse_images, cudaStream_t st) {
  const long long n = (long long)n_rows * W;
  const int stride = 6 + c->l;
  int rc = stnerf_set_scene(c, &v->scene);
  if (rc) return rc;
  if (c->scene.shared_frame_id) return STNERF_EINVAL;          // views carry one frame id per layer (retiming rays)
  rc = launch_raygen(v->Kinv, v->T, H, W, row0, row_step, n_rows, v->frame_ids, c->l, c->v_rays, stride, st);
  if (rc) return rc;
  const RayIdMap keep = c->idmap;
  c->idmap = RayIdMap{(long long)row0 * W, (long long)row_step * W, W};       // every pixel keeps the draws of an unsharded render
  OutSpec o{coarse_images, images, 1};
  rc = render_core(c, c->v_rays, n, stride, n1, n2, 0, nullptr, nullptr, v->seed, o, nullptr, st);
  c->idmap = keep;
  return rc;
}

int stnerf_render_views(stnerf_handle c, const stnerf_view* views_host, int n_views, int H, int W, int row0, int row_step,
                        int n_rows, int n1, int n2, float* images, float* coarse_images, int64_t view_stride, void* stream) {
  if (!c || !views_host || !images || n_views < 1 || H < 1 || W < 1 || row0 < 0 || row_step < 1 || n_rows < 1 || n2 < 1)
    return STNERF_EINVAL;
  if (row0 >= H || row0 + (long long)(n_rows - 1) * row_step >= H + row_step) return STNERF_EINVAL;   // at most one padding row
  const size_t n = (size_t)n_rows * W;
  if (view_stride < (int64_t)((size_t)(c->l + 1) * 5 * n)) return STNERF_EINVAL;
  if (grow((void**)&c->v_rays, &c->v_rays_bytes, n * (6 + c->l) * 4)) return STNERF_ENOMEM;
  for (int v = 0; v < n_views; ++v) {
    const int rc = render_one_view(c, views_host + v, H, W, row0, row_step, n_rows, n1, n2, images + (size_t)v * view_stride,
                                   coarse_images ? coarse_images + (size_t)v * view_stride : nullptr, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return STNERF_OK;
}

int stnerf_render_views_host(stnerf_handle c, const stnerf_view* views_host, int n_views, int H, int W, int n1, int n2,
                             float* images_host, void* stream) {
  if (!c || !views_host || !images_host || n_views < 1 || H < 1 || W < 1 || n2 < 1) return STNERF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)H * W, img_floats = (size_t)(c->l + 1) * 5 * n;
  int rc = grow((void**)&c->v_rays, &c->v_rays_bytes, n * (6 + c->l) * 4);
  for (int b = 0; b < 2; ++b) rc |= grow((void**)&c->v_img[b], &c->v_img_bytes[b], img_floats * 4);
  if (rc) return STNERF_ENOMEM;
  rc = ensure_copy_streams(c);
  if (rc) return rc;
  rc = ensure_events(c, 4);
  if (rc) return rc;
  // events 0/1: "view rendered into buffer b";  2/3: "buffer b copied out"
  for (int v = 0; v < n_views && rc == STNERF_OK; ++v) {
    const int b = v & 1;
    if (v >= 2) STNERF_CUDA(cudaStreamWaitEvent(st, c->ev_pool[2 + b], 0));       // buffer b is free again
    rc = render_one_view(c, views_host + v, H, W, 0, 1, H, n1, n2, c->v_img[b], nullptr, st);
    if (rc) break;
    STNERF_CUDA(cudaEventRecord(c->ev_pool[b], st));
    STNERF_CUDA(cudaStreamWaitEvent(c->copy_out, c->ev_pool[b], 0));
    STNERF_CUDA(cudaMemcpyAsync(images_host + (size_t)v * img_floats, c->v_img[b], img_floats * 4, cudaMemcpyDeviceToHost, c->copy_out));
    STNERF_CUDA(cudaEventRecord(c->ev_pool[2 + b], c->copy_out));
  }
  cudaError_t e1 = cudaStreamSynchronize(c->copy_out), e2 = cudaStreamSynchronize(st);
  if (rc) return rc;
  STNERF_CUDA(e1); STNERF_CUDA(e2);
  return STNERF_OK;
}

int stnerf_debug_read_depths(stnerf_handle c, int what, int layer, float* dst, int64_t n_rays, int S, void* stream) {
  if (!c || !dst || layer < 0 || layer >= c->l || n_rays < 0 || (what != 0 && what != 1)) return STNERF_EINVAL;
  if (!c->t_coarse || n_rays > c->last_chunk_rays) return STNERF_EINVAL;
  if (S != (what ? c->last_s2 : c->last_n1)) return STNERF_EINVAL;
  const long long R = c->chunk_rays;
  const float* src = what ? c->t_fine + (size_t)layer * R * c->cap_s2 : c->t_coarse + (size_t)layer * R * c->cap_n1;
  STNERF_CUDA(cudaMemcpyAsync(dst, src, (size_t)n_rays * S * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return STNERF_OK;
}

int stnerf_set_box_table(stnerf_handle c, const float* table_host, int n_frames) {
  if (!c || n_frames < 0 || (n_frames > 0 && !table_host)) return STNERF_EINVAL;
  STNERF_CUDA(cudaDeviceSynchronize());                  // a previous table may still be read by queued kernels
  cudaFree(c->box_table);
  c->box_table = nullptr; c->box_frames = 0;
  if (n_frames == 0) return STNERF_OK;
  const size_t bytes = (size_t)n_frames * c->l * 6 * sizeof(float);
  if (cudaMalloc((void**)&c->box_table, bytes) != cudaSuccess) { cudaGetLastError(); return STNERF_ENOMEM; }
  STNERF_CUDA(cudaMemcpy(c->box_table, table_host, bytes, cudaMemcpyHostToDevice));
  c->box_frames = n_frames;
  return STNERF_OK;
}

int stnerf_set_ray_ids(stnerf_handle c, int64_t base, int32_t width, int64_t row_stride) {
  if (!c || width < 0) return STNERF_EINVAL;
  c->idmap.base = base; c->idmap.width = width; c->idmap.row_stride = row_stride;
  return STNERF_OK;
}

int stnerf_selftest_umma(float* max_err_host) {
  if (!max_err_host) return STNERF_EINVAL;
  return tc_selftest(max_err_host);
}

int stnerf_selftest_umma_ts(float* max_err_host) {
  if (!max_err_host) return STNERF_EINVAL;
  return tc_selftest_ts(max_err_host);
}

int stnerf_selftest_umma_accum(int reps, float* max_err_host, float* mean_signed_rel_err_host) {
  if (!max_err_host || !mean_signed_rel_err_host || reps < 1 || reps > 4096) return STNERF_EINVAL;
  return tc_selftest_accum(reps, max_err_host, mean_signed_rel_err_host);
}

int stnerf_selftest_umma_pair(float* max_err_host) {
  if (!max_err_host) return STNERF_EINVAL;
  return tc_selftest_pair(max_err_host);
}

int stnerf_profile_begin(stnerf_handle c) {
  if (!c) return STNERF_EINVAL;
  if (!c->prof_counts) STNERF_CUDA(cudaHostAlloc((void**)&c->prof_counts, (size_t)PROF_MAX_CHUNKS * 32, cudaHostAllocDefault));
  for (auto& r : c->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  c->prof.clear();
  c->prof_chunks = 0;
  c->prof_on = true;
  return STNERF_OK;
}

int stnerf_profile_end(stnerf_handle c, stnerf_profile* out) {
  if (!c || !out) return STNERF_EINVAL;
  c->prof_on = false;
  STNERF_CUDA(cudaDeviceSynchronize());
  memset(out, 0, sizeof(*out));
  for (auto& r : c->prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) out->ms[r.cls] += ms;
    out->launches[r.cls] += 1;
    out->points[r.cls] += r.count_slot >= 0 ? (double)c->prof_counts[r.count_slot] * r.S : r.points;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  c->prof.clear();
  return STNERF_OK;
}

int stnerf_raygen(const float* Kinv_host, const float* T_host, int H, int W, int row0, int row_step, int n_rows,
                  const float* frame_ids_host, int n_frame_ids, float* rays, int ray_stride, void* stream) {
  if (!Kinv_host || !T_host || !rays || row_step < 1 || row0 < 0 || (n_frame_ids > 0 && !frame_ids_host))
    return STNERF_EINVAL;
  return launch_raygen(Kinv_host, T_host, H, W, row0, row_step, n_rows, frame_ids_host, n_frame_ids, rays, ray_stride,
                       (cudaStream_t)stream);
}

int stnerf_intersect_sample(const float* rays, int64_t n, int ray_stride, const float* bmin_host, const float* bmax_host,
                            int is_bkgd, int n1, const float* jitter, float* t, float* xyz, uint8_t* mask,
                            float* tfar_tnear, void* stream) {
  if (!rays || !bmin_host || !bmax_host || !jitter || n1 < 1 || ray_stride < 6) return STNERF_EINVAL;
  return launch_intersect_sample(rays, n, ray_stride, bmin_host, bmax_host, is_bkgd, n1, jitter, t, xyz, mask,
                                 tfar_tnear, (cudaStream_t)stream);
}

int stnerf_composite(const float* t, const float* rgb, const float* sigma, int64_t n, int S, float boarder, float* color,
                     float* depth, float* acc, float* w, void* stream) {
  if (!t || !rgb || !sigma || !color || !depth || !acc || S < 1) return STNERF_EINVAL;
  return launch_composite_simple(t, rgb, sigma, n, S, boarder, color, depth, acc, w, (cudaStream_t)stream);
}

int stnerf_sample_pdf(const float* t, const float* w, const float* u, int64_t n, int n1, int n2, float* z, float* t_fine,
                      void* stream) {
  if (!t || !w || !u || (!z && !t_fine) || n1 > STNERF_MAX_N1 * 4 || n1 + n2 > 4096) return STNERF_EINVAL;
  return launch_sample_pdf(t, w, u, n, n1, n2, z, t_fine, (cudaStream_t)stream);
}

int stnerf_positional_encoding(const float* x, int64_t P, int dim, int n_freq, float* out, void* stream) {
  if (!x || !out || dim < 1 || n_freq < 0 || n_freq > 16) return STNERF_EINVAL;
  return launch_posenc(x, P, dim, n_freq, out, (cudaStream_t)stream);
}

int stnerf_spacenet(stnerf_handle c, int layer, int fine, const float* pos, const float* dirs, const float* times,
                    int64_t P, float* rgb, float* sigma, void* stream) {
  if (!c || !pos || !dirs || !rgb || !sigma || layer < 0 || layer >= c->l || (fine != 0 && fine != 1)) return STNERF_EINVAL;
  SpaceNetDev& net = c->space[fine][layer];
  if (!net.loaded) return STNERF_ENOWEIGHTS;
  if (net.w.use_time && !times) return STNERF_EINVAL;
  PointSrc s;
  memset(&s, 0, sizeof(s));
  s.mode = SRC_EXPLICIT; s.pos = pos; s.dirs = dirs; s.times = times; s.pos_stride = 3; s.time_stride = 1;
  s.n_slots = P; s.S = 1; s.scale = 1.f;
  return run_spacenet(c, s, net, nullptr, rgb, sigma, (cudaStream_t)stream);
}

}  // extern "C"

__global__ void any_fraction_kernel(const float* __restrict__ xyzt, long long P, int* __restrict__ flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) {
    const float t = xyzt[4 * i + 3];
    if (floorf(t) != t) atomicOr(flag, 1);
  }
}

extern "C" int stnerf_motionnet(stnerf_handle c, int layer, const float* xyzt, int64_t P, int lerp_mode, float* flow,
                                void* stream) {
  if (!c || !xyzt || !flow || layer < 1 || layer >= c->l || lerp_mode < -1 || lerp_mode > 1) return STNERF_EINVAL;
  MotionNetDev& net = c->motion[layer];
  if (!net.loaded) return STNERF_ENOWEIGHTS;
  cudaStream_t st = (cudaStream_t)stream;
  if (lerp_mode < 0 && P > 0) {
    STNERF_CUDA(cudaMemsetAsync(c->any_frac, 0, 4, st));
    any_fraction_kernel<<<(int)((P + 255) / 256), 256, 0, st>>>(xyzt, P, c->any_frac);
    STNERF_LAUNCH_CHECK();
  }
  PointSrc s;
  memset(&s, 0, sizeof(s));
  s.mode = SRC_EXPLICIT; s.pos = xyzt; s.times = xyzt + 3; s.pos_stride = 4; s.time_stride = 4;
  s.n_slots = P; s.S = 1; s.scale = 1.f;
  return run_motionnet(c, s, net, c->any_frac, lerp_mode, nullptr, flow, st);
}

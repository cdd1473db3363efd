// tcgen05 / TMEM evaluation of SpaceNet and MotionNet (precision modes TC_3XF16 "exact" and TC_F16 "fast").
//
// One persistent CTA per SM walks tiles of 128 points.  Per tile the whole network runs on-chip:
//   * activations (A operand) live in shared memory as fp16 hi/lo pairs in the canonical 128B-swizzled K-major
//     UMMA layout; they never leave the SM between layers;
//   * weights (B operand) are pre-packed on the host into 16 KB blocks [128 out-rows x 64 k] that are already the
//     swizzled shared-memory image, in the exact order the MMA warp consumes them, and stream through a 4-stage
//     ring with 1-D bulk async copies (cp.async.bulk + mbarrier complete_tx) from L2;
//   * accumulators live in TMEM (two 128x256 fp32 buffers = all 512 columns) so the epilogue of layer k
//     (tcgen05.ld -> bias -> ReLU -> fp16 hi/lo split -> st.shared) overlaps the MMAs of layer k+1, k-chunk by k-chunk;
//   * exact mode issues three fp16 MMAs per product, D += Ahi*Whi + Alo*Whi + Ahi*Wlo (fp32 accumulate), which
//     reproduces fp32 products to ~2^-22 (the only tensor-core formulation inside the 1e-3 gate);
//   * the 1-wide density head, the 3-wide rgb / flow heads and all biases are fp32 FFMA work in the epilogue.
//
// Warp roles (384 threads): warp 0 = weight producer, warp 1 = MMA issuer + TMEM owner, warps 4..11 = epilogue /
// encoding warps (warp%4 selects the TMEM lane quarter, (warp-4)/4 the column half of every 64-column chunk).
//
// Restates modeling/spacenet.py:101-160, modeling/motion_net.py:34-71, utils/dimension_kernel.py:24-33.
#include <cuda_fp16.h>
#include <vector>
#include "mlp_tc.cuh"

namespace stnerf {

namespace {

constexpr int TILE_M = 128;
constexpr int BLOCK_BYTES = 16384;           // [128 rows x 64 k] fp16, 128B-swizzled K-major
constexpr int NSTAGE = 4;
constexpr int NTHREADS = 384;
constexpr int EPI_WARP0 = 4, N_EPI_WARPS = 8, N_EPI_THREADS = 256;

// shared memory map (bytes, from a 1024-aligned base)
constexpr int SM_ACT = 0;                    // 8 blocks: term*4 + kchunk (term 0 = hi, 1 = lo)
constexpr int SM_ENC = 8 * BLOCK_BYTES;      // 2 blocks: hi, lo (SpaceNet).  MotionNet: ACT = blocks 0-3, ENC = blocks 4-7
constexpr int SM_RING = 10 * BLOCK_BYTES;    // NSTAGE weight blocks
constexpr int SM_MISC = SM_RING + NSTAGE * BLOCK_BYTES;
constexpr int MISC_BAR = 0;                  // mbarriers (8 B each)
constexpr int BAR_WFULL = 0, BAR_WEMPTY = 4, BAR_AREADY = 8, BAR_DFULL = 13, BAR_DEMPTY = 15, N_BARS = 17;
constexpr int MISC_TMEM = 144;               // tmem base address
constexpr int MISC_OUTIDX = 160;             // int32[128]
constexpr int MISC_PART = MISC_OUTIDX + 512; // float[128][4]: partial head sums of column-half 1 (sigma | rgb / flow);
                                             // MotionNet parks the un-deformed xyz of each row here until the last layer
constexpr int SM_TOTAL = SM_MISC + MISC_PART + 2048;
static_assert(SM_TOTAL <= 232448, "shared memory budget (227 KB per CTA)");

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (kernel aborts with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 26)) {
      printf("stnerf mlp_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x,
             bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// Instruction descriptor: fp16 A/B (K-major), fp32 D, M = 128, N = 128 (cute::UMMA::InstrDescriptor).
constexpr uint32_t IDESC_N128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// byte offset of element (row, col) of a [128 x 64] fp16 block in the 128B-swizzled K-major layout
__host__ __device__ inline uint32_t sw128_offset(int row, int col) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 3) ^ (row & 7)) & 7) << 4) + ((col & 7) << 1));
}

// ---------------------------------------------------------------------------------------------------------
// network schedules (compile-time)
// ---------------------------------------------------------------------------------------------------------
enum { NET_SPACE = 0, NET_MOTION = 1 };

template <int NET> struct Sched;
template <> struct Sched<NET_SPACE> {
  static constexpr int N_LAYERS = 8;
  static constexpr int ACT_CHUNKS = 4;                      // 256-wide activations
  static constexpr int ENC_CHUNKS = 1;
  static constexpr int act_base = SM_ACT, enc_base = SM_ENC;
  static constexpr int LO_STRIDE = 4 * BLOCK_BYTES;         // ACT lo blocks follow the 4 hi blocks
  static constexpr int ENC_LO_STRIDE = BLOCK_BYTES;
  __host__ __device__ static constexpr int n_halves(int l) { return l == 7 ? 1 : 2; }
  // k-chunk sources of layer l: act chunks used (0 or 4) then enc chunks used (0 or 1)
  __host__ __device__ static constexpr int act_chunks(int l) { return l == 0 ? 0 : 4; }
  __host__ __device__ static constexpr int enc_chunks(int l) { return (l == 0 || l == 4 || l == 7) ? 1 : 0; }
  __host__ __device__ static constexpr bool enc_needs_wait(int l) { return l == 0 || l == 7; }
};
template <> struct Sched<NET_MOTION> {
  static constexpr int N_LAYERS = 5;
  static constexpr int ACT_CHUNKS = 2;                      // 128-wide activations
  static constexpr int ENC_CHUNKS = 2;                      // PE(84) padded to 128
  static constexpr int act_base = SM_ACT, enc_base = SM_ACT + 4 * BLOCK_BYTES;
  static constexpr int LO_STRIDE = 2 * BLOCK_BYTES;
  static constexpr int ENC_LO_STRIDE = 2 * BLOCK_BYTES;
  __host__ __device__ static constexpr int n_halves(int) { return 1; }
  __host__ __device__ static constexpr int act_chunks(int l) { return l == 0 ? 0 : 2; }
  __host__ __device__ static constexpr int enc_chunks(int l) { return l == 0 ? 2 : 0; }
  __host__ __device__ static constexpr bool enc_needs_wait(int l) { return l == 0; }
};

template <int NET>
__host__ __device__ constexpr int blocks_per_tile() {
  int n = 0;
  for (int l = 0; l < Sched<NET>::N_LAYERS; ++l)
    n += (Sched<NET>::act_chunks(l) + Sched<NET>::enc_chunks(l)) * Sched<NET>::n_halves(l) * 2;
  return n;
}

struct TcParams {
  PointSrc src;
  const uint8_t* wblocks;     // packed weight stream (hi/lo blocks in consumption order)
  const float* aux;           // fp32: biases per layer [8][256] | w_sigma[256] | b_sigma | w_out[3][128] | b_out[3]
  int exact;                  // 1: 3-term split, 0: single fp16 pass
  int use_time;
  // outputs
  float* raw;                 // float4 per sample (pipeline mode)
  float* rgb_out;             // explicit mode
  float* sigma_out;
  float* xyz_out;             // MotionNet: deformed position (pipeline) ...
  float* flow_out;            // ... or flow (explicit)
  const int* lerp_flag;
  int lerp_force;
};

constexpr int AUX_BIAS = 0, AUX_WSIG = 8 * 256, AUX_BSIG = AUX_WSIG + 256, AUX_WOUT = AUX_BSIG + 4,
              AUX_BOUT = AUX_WOUT + 3 * 128, AUX_FLOATS = AUX_BOUT + 4;

// ---------------------------------------------------------------------------------------------------------
// point fetch (same arithmetic as mlp_simt.cu::fetch_point)
// ---------------------------------------------------------------------------------------------------------
struct Pt { float x, y, z, dx, dy, dz, tm; int out_index; };

__device__ __forceinline__ Pt fetch_pt(const PointSrc& s, long long p, long long n_points) {
  Pt q;
  q.x = q.y = q.z = q.dx = q.dy = q.dz = q.tm = 0.f;
  q.out_index = -1;
  if (p >= n_points) return q;
  if (s.mode == SRC_EXPLICIT) {
    const float* pp = s.pos + p * s.pos_stride;
    q.x = pp[0]; q.y = pp[1]; q.z = pp[2];
    if (s.dirs) { q.dx = s.dirs[3 * p]; q.dy = s.dirs[3 * p + 1]; q.dz = s.dirs[3 * p + 2]; }
    if (s.times) q.tm = s.times[p * s.time_stride];
    q.out_index = (int)p;
    return q;
  }
  const long long slot = p / s.S;
  const int k = (int)(p - slot * s.S);
  const long long ray = s.hit ? (long long)s.hit[slot] : slot;
  const float* rp = s.rays + ray * s.ray_stride;
  q.dx = rp[3]; q.dy = rp[4]; q.dz = rp[5];
  q.tm = rp[6 + s.layer];
  q.out_index = (int)(ray * s.S + k);
  if (s.mode == SRC_XYZ) {
    q.x = s.pos[3 * p]; q.y = s.pos[3 * p + 1]; q.z = s.pos[3 * p + 2];
    return q;
  }
  const float tt = s.t[ray * s.S + k];
  float v[3] = {__fadd_rn(__fmul_rn(tt, q.dx), rp[0]), __fadd_rn(__fmul_rn(tt, q.dy), rp[1]),
                __fadd_rn(__fmul_rn(tt, q.dz), rp[2])};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (s.shift_on) v[a] = __fsub_rn(v[a], s.shift[a]);
    if (s.scale_on) v[a] = __fadd_rn(__fdiv_rn(__fsub_rn(v[a], s.pivot[a]), s.scale), s.pivot[a]);
  }
  q.x = v[0]; q.y = v[1]; q.z = v[2];
  return q;
}

// store one fp32 value as fp16 hi (+ lo) at element (row, col) of a hi block / its lo twin
__device__ __forceinline__ void put_split(uint8_t* hi_block, int lo_stride, int row, int col, float v, bool exact) {
  const __half h = __float2half_rn(v);
  const uint32_t off = sw128_offset(row, col);
  *reinterpret_cast<__half*>(hi_block + off) = h;
  if (exact) *reinterpret_cast<__half*>(hi_block + lo_stride + off) = __float2half_rn(v - __half2float(h));
}

// ---------------------------------------------------------------------------------------------------------
// epilogue helpers
// ---------------------------------------------------------------------------------------------------------
// 32 accumulator columns of one row -> bias + ReLU -> fp16 hi/lo -> four 16-byte stores each.
// `dotw` (optional): fp32 head weights for these 32 columns, accumulated into dot[0..ND).
template <int ND>
__device__ __forceinline__ void epi_store32(const float (&acc)[32], const float* __restrict__ bias, uint8_t* hi_blk,
                                            int lo_stride, int row, int col0, bool exact,
                                            const float* __restrict__ dotw, int dotw_stride, float (&dot)[ND > 0 ? ND : 1]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {               // 8 columns -> one 16-byte chunk
    uint32_t hp[4], lp[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = g * 8 + e * 2;
      float v0 = fmaxf(acc[c] + __ldg(bias + c), 0.f);
      float v1 = fmaxf(acc[c + 1] + __ldg(bias + c + 1), 0.f);
      v0 = fminf(v0, 65504.f); v1 = fminf(v1, 65504.f);
      if (ND > 0) {
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          dot[d] = fmaf(v0, __ldg(dotw + d * dotw_stride + c), dot[d]);
          dot[d] = fmaf(v1, __ldg(dotw + d * dotw_stride + c + 1), dot[d]);
        }
      }
      const __half2 h = __floats2half2_rn(v0, v1);
      hp[e] = *reinterpret_cast<const uint32_t*>(&h);
      if (exact) {
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
        lp[e] = *reinterpret_cast<const uint32_t*>(&l);
      }
    }
    if (hi_blk) {
      const uint32_t off = sw128_offset(row, col0 + g * 8);
      *reinterpret_cast<uint4*>(hi_blk + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
      if (exact) *reinterpret_cast<uint4*>(hi_blk + lo_stride + off) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------
template <int NET>
__global__ void __launch_bounds__(NTHREADS, 1) mlp_tc_kernel(const TcParams P) {
  using S = Sched<NET>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) {                    // SWIZZLE_128B operands need a 1024-byte aligned base
    if (threadIdx.x == 0) printf("stnerf mlp_tc: dynamic shared memory base %u is not 1024-byte aligned\n", sbase);
    __trap();
  }
  const uint32_t bars = sbase + SM_MISC + MISC_BAR;
  auto BAR = [bars](int i) { return bars + 8u * (uint32_t)i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_MISC + MISC_TMEM);
  int* s_outidx = reinterpret_cast<int*>(smem + SM_MISC + MISC_OUTIDX);
  float* s_part = reinterpret_cast<float*>(smem + SM_MISC + MISC_PART);     // [128][4]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool exact = P.exact != 0;
  const long long n_points = src_num_points(P.src);
  const long long n_tiles = (n_points + TILE_M - 1) / TILE_M;

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(BAR(BAR_WFULL + i), 1); mbar_init(BAR(BAR_WEMPTY + i), 1); }
    for (int i = 0; i < 5; ++i) mbar_init(BAR(BAR_AREADY + i), N_EPI_WARPS);
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(BAR_DFULL + i), 1); mbar_init(BAR(BAR_DEMPTY + i), N_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== weight producer ===============================
    if (lane == 0) {
      uint32_t cnt = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int blk = 0;
        for (int l = 0; l < S::N_LAYERS; ++l) {
          const int nch = S::act_chunks(l) + S::enc_chunks(l), nh = S::n_halves(l);
          for (int c = 0; c < nch; ++c)
            for (int term = 0; term < 2; ++term)
              for (int h = 0; h < nh; ++h, ++blk) {
                if (term == 1 && !exact) continue;           // fast mode never touches the lo blocks
                const uint32_t s = cnt % NSTAGE, n = cnt / NSTAGE;
                mbar_wait(BAR(BAR_WEMPTY + s), (n & 1) ^ 1);
                mbar_expect_tx(BAR(BAR_WFULL + s), BLOCK_BYTES);
                bulk_g2s(sbase + SM_RING + s * BLOCK_BYTES, P.wblocks + (size_t)blk * BLOCK_BYTES, BLOCK_BYTES,
                         BAR(BAR_WFULL + s));
                ++cnt;
              }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      uint32_t cnt = 0;            // weight blocks consumed
      uint32_t g = 0;              // global layer counter (selects the TMEM buffer)
      uint32_t a_uses[5] = {0, 0, 0, 0, 0};
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < S::N_LAYERS; ++l, ++g) {
          const uint32_t b = g & 1;
          mbar_wait(BAR(BAR_DEMPTY + b), ((g >> 1) & 1) ^ 1);      // accumulator buffer drained (layer g-2)
          tc_fence_after();
          const uint32_t dcol = tmem_base + b * 256;
          const int nact = S::act_chunks(l), nch = nact + S::enc_chunks(l), nh = S::n_halves(l);
          for (int c = 0; c < nch; ++c) {
            uint32_t a_hi, a_lo;
            if (c < nact) {
              a_hi = sbase + S::act_base + c * BLOCK_BYTES;
              a_lo = a_hi + S::LO_STRIDE;
              mbar_wait(BAR(BAR_AREADY + c), a_uses[c] & 1);
              ++a_uses[c];
            } else {
              const int e = c - nact;
              a_hi = sbase + S::enc_base + e * BLOCK_BYTES;
              a_lo = a_hi + S::ENC_LO_STRIDE;
              if (S::enc_needs_wait(l) && e == 0) {                // one arrival phase covers every enc chunk
                mbar_wait(BAR(BAR_AREADY + 4), a_uses[4] & 1);
                ++a_uses[4];
              }
            }
            tc_fence_after();
            for (int term = 0; term < 2; ++term) {
              if (term == 1 && !exact) continue;
              for (int h = 0; h < nh; ++h) {
                const uint32_t s = cnt % NSTAGE, n = cnt / NSTAGE;
                mbar_wait(BAR(BAR_WFULL + s), n & 1);
                tc_fence_after();
                const uint32_t wsm = sbase + SM_RING + s * BLOCK_BYTES;
                const uint32_t d = dcol + h * 128;
                // hi block: D += Ahi*Whi (+ Alo*Whi);  lo block: D += Ahi*Wlo
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_f16(d, make_desc(a_hi + ks * 32), make_desc(wsm + ks * 32), IDESC_N128,
                           (c == 0 && term == 0 && ks == 0) ? 0u : 1u);
                if (term == 0 && exact) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    umma_f16(d, make_desc(a_lo + ks * 32), make_desc(wsm + ks * 32), IDESC_N128, 1u);
                }
                umma_commit(BAR(BAR_WEMPTY + s));                   // ring slot reusable once these MMAs retire
                ++cnt;
              }
            }
          }
          umma_commit(BAR(BAR_DFULL + b));                          // accumulator of layer g complete
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // =============================== encoding + epilogue warps ===============================
    const int ew = warp - EPI_WARP0;            // 0..7
    const int q = ew & 3, hh = ew >> 2;         // TMEM lane quarter, column half
    const int row = q * 32 + lane;              // accumulator row owned in the epilogue
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float* bias_all = P.aux + AUX_BIAS;
    // encoding phase mapping: two adjacent lanes share a row
    const int erow = ew * 16 + (lane >> 1), epar = lane & 1;
    uint32_t g = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      // ---- fetch the tile's points and write the input encoding ----
      const Pt pt = fetch_pt(P.src, tile * TILE_M + erow, n_points);
      if (epar == 0) s_outidx[erow] = pt.out_index;
      uint8_t* enc_hi = smem + S::enc_base;
      if (NET == NET_SPACE) {
        // PE(pos, L=10): col 0..2 raw, 3+6f+d sin, 6+6f+d cos (utils/dimension_kernel.py:24-33); col 63 = 0
        const float xs[3] = {pt.x, pt.y, pt.z};
        if (epar == 0) {
#pragma unroll
          for (int d = 0; d < 3; ++d) put_split(enc_hi, S::ENC_LO_STRIDE, erow, d, xs[d], exact);
        } else {
          put_split(enc_hi, S::ENC_LO_STRIDE, erow, 63, 0.f, exact);
        }
#pragma unroll
        for (int ff = 0; ff < 5; ++ff) {
          const int f = epar * 5 + ff;
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            float sn, cs;
            sincosf(xs[d] * (float)(1 << f), &sn, &cs);
            put_split(enc_hi, S::ENC_LO_STRIDE, erow, 3 + 6 * f + d, sn, exact);
            put_split(enc_hi, S::ENC_LO_STRIDE, erow, 6 + 6 * f + d, cs, exact);
          }
        }
      } else {
        // PE([x,y,z,t], L=10) (+ the reference's lerp of the encodings of floor(t), floor(t)+1, motion_net.py:48-63)
        const bool lerp = P.lerp_force >= 0 ? (P.lerp_force != 0) : (P.lerp_flag && *P.lerp_flag != 0);
        const float lo_t = floorf(pt.tm), wgt = pt.tm - lo_t, omw = 1.0f - wgt;
        const float in4[4] = {pt.x, pt.y, pt.z, pt.tm};
        if (epar == 0) {
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            float v = in4[d];
            if (lerp) {
              const float a = d < 3 ? in4[d] : lo_t, b2 = d < 3 ? in4[d] : lo_t + 1.0f;
              v = __fadd_rn(__fmul_rn(omw, a), __fmul_rn(wgt, b2));
            }
            put_split(enc_hi, S::ENC_LO_STRIDE, erow, d, v, exact);
          }
        }
        // zero padding columns 84..127 (second enc chunk, cols 20..63)
        for (int c = 20 + epar; c < 64; c += 2) put_split(enc_hi + BLOCK_BYTES, S::ENC_LO_STRIDE, erow, c, 0.f, exact);
#pragma unroll
        for (int ff = 0; ff < 5; ++ff) {
          const int f = epar * 5 + ff;
          const float fr = (float)(1 << f);
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            float sn, cs;
            if (!lerp) {
              sincosf(in4[d] * fr, &sn, &cs);
            } else {
              const float a = d < 3 ? in4[d] : lo_t, b2 = d < 3 ? a : lo_t + 1.0f;
              float s0, c0, s1, c1;
              sincosf(a * fr, &s0, &c0);
              sincosf(b2 * fr, &s1, &c1);
              sn = __fadd_rn(__fmul_rn(omw, s0), __fmul_rn(wgt, s1));
              cs = __fadd_rn(__fmul_rn(omw, c0), __fmul_rn(wgt, c1));
            }
            const int cs_col = 4 + 8 * f + d, cc_col = 8 + 8 * f + d;
            put_split(enc_hi + (cs_col >> 6) * BLOCK_BYTES, S::ENC_LO_STRIDE, erow, cs_col & 63, sn, exact);
            put_split(enc_hi + (cc_col >> 6) * BLOCK_BYTES, S::ENC_LO_STRIDE, erow, cc_col & 63, cs, exact);
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(BAR_AREADY + 4));
      // MotionNet keeps the un-deformed position of the epilogue row for the final xyz + flow
      float my_xyz[3] = {0.f, 0.f, 0.f};
      if (NET == NET_MOTION) {
        float* s_xyz = s_part;                  // [128][4] reused: written now, read in the last epilogue
        if (epar == 0) { s_xyz[erow * 4 + 0] = pt.x; s_xyz[erow * 4 + 1] = pt.y; s_xyz[erow * 4 + 2] = pt.z; }
      }

      float sig_dot[1] = {0.f};
      for (int l = 0; l < S::N_LAYERS; ++l, ++g) {
        const uint32_t b = g & 1;
        mbar_wait(BAR(BAR_DFULL + b), (g >> 1) & 1);
        tc_fence_after();
        const uint32_t dcol = lane_taddr + b * 256;
        const bool last = (l == S::N_LAYERS - 1);
        if (NET == NET_SPACE && l == 4) {
          // MMAs of layer 4 are done with the position encoding: overwrite it with relu(PE(dir) | PE(time))
          // (modeling/spacenet.py:141-149 and the leading ReLU of rgb_net, :82); cols 27/48..63 = 0
          const float ds[3] = {pt.dx, pt.dy, pt.dz};
          const int ntime = P.use_time ? PE_TIME : 0;
          if (epar == 0) {
#pragma unroll
            for (int d = 0; d < 3; ++d) put_split(enc_hi, S::ENC_LO_STRIDE, erow, d, fmaxf(ds[d], 0.f), exact);
            if (P.use_time) put_split(enc_hi, S::ENC_LO_STRIDE, erow, PE_DIR, fmaxf(pt.tm, 0.f), exact);
          }
          for (int c = PE_DIR + ntime + epar; c < 64; c += 2) put_split(enc_hi, S::ENC_LO_STRIDE, erow, c, 0.f, exact);
#pragma unroll
          for (int ff = 0; ff < 2; ++ff) {
            const int f = epar * 2 + ff;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              float sn, cs;
              sincosf(ds[d] * (float)(1 << f), &sn, &cs);
              put_split(enc_hi, S::ENC_LO_STRIDE, erow, 3 + 6 * f + d, fmaxf(sn, 0.f), exact);
              put_split(enc_hi, S::ENC_LO_STRIDE, erow, 6 + 6 * f + d, fmaxf(cs, 0.f), exact);
            }
          }
          if (P.use_time) {
#pragma unroll
            for (int ff = 0; ff < 5; ++ff) {
              const int f = epar * 5 + ff;
              float sn, cs;
              sincosf(pt.tm * (float)(1 << f), &sn, &cs);
              put_split(enc_hi, S::ENC_LO_STRIDE, erow, PE_DIR + 1 + 2 * f, fmaxf(sn, 0.f), exact);
              put_split(enc_hi, S::ENC_LO_STRIDE, erow, PE_DIR + 2 + 2 * f, fmaxf(cs, 0.f), exact);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(BAR_AREADY + 4));
        }
        const int width = S::n_halves(l) * 128;                  // output features of this layer
        const float* bias = bias_all + l * 256;
        if (!last) {
          const bool sigma_layer = (NET == NET_SPACE && l == 6);
          for (int j = 0; j < width / 64; ++j) {                 // 64-column chunk j -> ACT k-chunk j
            float acc[32];
            const int col0 = j * 64 + hh * 32;
            tmem_ld32(dcol + (uint32_t)col0, acc);
            uint8_t* blk = smem + S::act_base + j * BLOCK_BYTES;
            if (sigma_layer) {
              epi_store32<1>(acc, bias + col0, blk, S::LO_STRIDE, row, hh * 32, exact, P.aux + AUX_WSIG + col0, 0, sig_dot);
            } else {
              float dummy[1];
              epi_store32<0>(acc, bias + col0, blk, S::LO_STRIDE, row, hh * 32, exact, nullptr, 0, dummy);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(BAR_AREADY + j));
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(BAR_DEMPTY + b));
        } else {
          // last layer: 128 features -> 3-wide head in fp32 (rgb_net.3 / motion_net.10), no activation store
          float dot3[3] = {0.f, 0.f, 0.f};
          for (int j = 0; j < 2; ++j) {
            float acc[32];
            const int col0 = j * 64 + hh * 32;
            tmem_ld32(dcol + (uint32_t)col0, acc);
            epi_store32<3>(acc, bias + col0, nullptr, 0, row, 0, exact, P.aux + AUX_WOUT + col0, 128, dot3);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(BAR_DEMPTY + b));
          // combine the two column halves through shared memory
          if (NET == NET_MOTION) {
            epi_bar_sync();     // the xyz parked by the encoding phase (other warps) is visible from here on
            my_xyz[0] = s_part[row * 4 + 0]; my_xyz[1] = s_part[row * 4 + 1]; my_xyz[2] = s_part[row * 4 + 2];
          }
          epi_bar_sync();
          if (hh == 1) {
            s_part[row * 4 + 0] = dot3[0]; s_part[row * 4 + 1] = dot3[1]; s_part[row * 4 + 2] = dot3[2];
            s_part[row * 4 + 3] = sig_dot[0];
          }
          epi_bar_sync();
          if (hh == 0) {
            const int oi = s_outidx[row];
            if (oi >= 0) {
              const float o0 = dot3[0] + s_part[row * 4 + 0] + P.aux[AUX_BOUT + 0];
              const float o1 = dot3[1] + s_part[row * 4 + 1] + P.aux[AUX_BOUT + 1];
              const float o2 = dot3[2] + s_part[row * 4 + 2] + P.aux[AUX_BOUT + 2];
              if (NET == NET_SPACE) {
                const float sg = sig_dot[0] + s_part[row * 4 + 3] + P.aux[AUX_BSIG];
                if (P.raw) reinterpret_cast<float4*>(P.raw)[oi] = make_float4(o0, o1, o2, sg);
                if (P.rgb_out) { P.rgb_out[3 * (size_t)oi] = o0; P.rgb_out[3 * (size_t)oi + 1] = o1; P.rgb_out[3 * (size_t)oi + 2] = o2; }
                if (P.sigma_out) P.sigma_out[oi] = sg;
              } else {
                const long long p = tile * TILE_M + row;          // compact point index
                if (P.flow_out) { P.flow_out[3 * p] = o0; P.flow_out[3 * p + 1] = o1; P.flow_out[3 * p + 2] = o2; }
                if (P.xyz_out) {
                  P.xyz_out[3 * p] = __fadd_rn(my_xyz[0], o0);
                  P.xyz_out[3 * p + 1] = __fadd_rn(my_xyz[1], o1);
                  P.xyz_out[3 * p + 2] = __fadd_rn(my_xyz[2], o2);
                }
              }
            }
          }
          epi_bar_sync();       // s_part / s_outidx are rewritten by the next tile
        }
      }
    }
  }
  // teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------
// self-test: one 128x128x64 fp16 UMMA through exactly the descriptors / swizzle / TMEM load used above
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const __half* __restrict__ A, const uint8_t* __restrict__ Bblk,
                                                              float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + 2 * BLOCK_BYTES, bar2 = bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 2 * BLOCK_BYTES + 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((sbase & 1023u) != 0) __trap();
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 128);
  // A: row `tid`, 64 columns, written with the epilogue's element mapping
  for (int c = 0; c < 64; ++c) *reinterpret_cast<__half*>(smem + sw128_offset(tid, c)) = A[tid * 64 + c];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    // B arrives through the same bulk-copy path as the weight ring
    mbar_expect_tx(bar2, BLOCK_BYTES);
    bulk_g2s(sbase + BLOCK_BYTES, Bblk, BLOCK_BYTES, bar2);
    mbar_wait(bar2, 0);
    tc_fence_after();
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      umma_f16(tmem_base, make_desc(sbase + ks * 32), make_desc(sbase + BLOCK_BYTES + ks * 32), IDESC_N128, ks ? 1u : 0u);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int j = 0; j < 4; ++j) {
    float acc[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + j * 32, acc);
#pragma unroll
    for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * 128 + j * 32 + i] = acc[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// ---------------------------------------------------------------------------------------------------------
// host: weight packing
// ---------------------------------------------------------------------------------------------------------
// Emit the hi and lo blocks of W[n0..n0+127][k0..k0+63] (W row-major (N,K), zero outside) in the SW128 image.
void emit_block_pair(std::vector<uint8_t>& hi, std::vector<uint8_t>& lo, const float* W, int N, int K, int ldw, int n0,
                     int k0, int kvalid0, int kvalid1) {
  hi.assign(BLOCK_BYTES, 0);
  lo.assign(BLOCK_BYTES, 0);
  for (int r = 0; r < 128; ++r) {
    const int n = n0 + r;
    if (n >= N) continue;
    for (int c = 0; c < 64; ++c) {
      const int k = k0 + c;
      if (k < kvalid0 || k >= kvalid1 || k >= K) continue;
      const float w = W[(size_t)n * ldw + k];
      const __half h = __float2half_rn(w);
      const __half l = __float2half_rn(w - __half2float(h));
      const uint32_t off = sw128_offset(r, c);
      memcpy(hi.data() + off, &h, 2);
      memcpy(lo.data() + off, &l, 2);
    }
  }
}

struct LayerSpec { const float* W; int N, K_total; int act_k; int enc_k0, enc_k; };   // W (N, K_total) row-major

int pack_stream(TcNet& net, const std::vector<LayerSpec>& layers, const std::vector<float>& aux) {
  std::vector<uint8_t> stream, hi, lo;
  for (const LayerSpec& L : layers) {
    const int nh = (L.N + 127) / 128;
    // k-chunks: act part (columns [0, act_k)) then enc part (columns [enc_k0, enc_k0 + enc_k)), each in 64-wide chunks
    std::vector<std::pair<int, int>> chunks;      // (k0, kend)
    for (int k = 0; k < L.act_k; k += 64) chunks.push_back({k, std::min(k + 64, L.act_k)});
    for (int k = 0; k < L.enc_k; k += 64) chunks.push_back({L.enc_k0 + k, L.enc_k0 + std::min(k + 64, L.enc_k)});
    for (auto& ch : chunks) {
      std::vector<std::vector<uint8_t>> his(nh), los(nh);
      for (int h = 0; h < nh; ++h) emit_block_pair(his[h], los[h], L.W, L.N, L.K_total, L.K_total, h * 128, ch.first, ch.first, ch.second);
      for (int h = 0; h < nh; ++h) stream.insert(stream.end(), his[h].begin(), his[h].end());
      for (int h = 0; h < nh; ++h) stream.insert(stream.end(), los[h].begin(), los[h].end());
    }
  }
  tc_free(net);
  net.n_blocks = (int)(stream.size() / BLOCK_BYTES);
  net.blob_bytes = stream.size();
  STNERF_CUDA(cudaMalloc(&net.blob, stream.size()));
  STNERF_CUDA(cudaMemcpy(net.blob, stream.data(), stream.size(), cudaMemcpyHostToDevice));
  STNERF_CUDA(cudaMalloc((void**)&net.aux, aux.size() * sizeof(float)));
  STNERF_CUDA(cudaMemcpy(net.aux, aux.data(), aux.size() * sizeof(float), cudaMemcpyHostToDevice));
  return STNERF_OK;
}

}  // namespace

void tc_free(TcNet& net) {
  if (net.blob) cudaFree(net.blob);
  if (net.aux) cudaFree(net.aux);
  net.blob = nullptr; net.aux = nullptr; net.blob_bytes = 0; net.n_blocks = 0;
}

int tc_pack_spacenet(TcNet& net, const float* p, bool use_time) {
  const int krgb = HID + PE_DIR + (use_time ? PE_TIME : 0);
  std::vector<float> aux(AUX_FLOATS, 0.f);
  std::vector<LayerSpec> layers;
  const int Ks[7] = {PE_POS, HID, HID, HID, HID + PE_POS, HID, HID};
  for (int i = 0; i < 7; ++i) {
    LayerSpec L;
    L.W = p; L.N = HID; L.K_total = Ks[i];
    if (i == 0) { L.act_k = 0; L.enc_k0 = 0; L.enc_k = PE_POS; }
    else if (i == 4) { L.act_k = HID; L.enc_k0 = HID; L.enc_k = PE_POS; }     // cat[x, PE(pos)] (spacenet.py:137)
    else { L.act_k = HID; L.enc_k0 = 0; L.enc_k = 0; }
    layers.push_back(L);
    p += (size_t)HID * Ks[i];
    memcpy(aux.data() + AUX_BIAS + i * 256, p, HID * sizeof(float));
    p += HID;
  }
  memcpy(aux.data() + AUX_WSIG, p, HID * sizeof(float)); p += HID;
  aux[AUX_BSIG] = *p++;
  LayerSpec L;
  L.W = p; L.N = HEAD; L.K_total = krgb; L.act_k = HID; L.enc_k0 = HID; L.enc_k = krgb - HID;   // cat[x, PE(dir), PE(t)]
  layers.push_back(L);
  p += (size_t)HEAD * krgb;
  memcpy(aux.data() + AUX_BIAS + 7 * 256, p, HEAD * sizeof(float)); p += HEAD;
  memcpy(aux.data() + AUX_WOUT, p, 3 * HEAD * sizeof(float)); p += 3 * HEAD;
  memcpy(aux.data() + AUX_BOUT, p, 3 * sizeof(float));
  net.use_time = use_time ? 1 : 0;
  const int rc = pack_stream(net, layers, aux);
  if (rc) return rc;
  return net.n_blocks == blocks_per_tile<NET_SPACE>() ? STNERF_OK : STNERF_EINVAL;
}

int tc_pack_motionnet(TcNet& net, const float* p) {
  std::vector<float> aux(AUX_FLOATS, 0.f);
  std::vector<LayerSpec> layers;
  for (int i = 0; i < 5; ++i) {
    LayerSpec L;
    const int K = i == 0 ? PE_MOTION : HEAD;
    L.W = p; L.N = HEAD; L.K_total = K;
    if (i == 0) { L.act_k = 0; L.enc_k0 = 0; L.enc_k = 128; }       // PE(84) zero-padded to two 64-wide chunks
    else { L.act_k = HEAD; L.enc_k0 = 0; L.enc_k = 0; }
    layers.push_back(L);
    p += (size_t)HEAD * K;
    memcpy(aux.data() + AUX_BIAS + i * 256, p, HEAD * sizeof(float));
    p += HEAD;
  }
  memcpy(aux.data() + AUX_WOUT, p, 3 * HEAD * sizeof(float)); p += 3 * HEAD;
  memcpy(aux.data() + AUX_BOUT, p, 3 * sizeof(float));
  const int rc = pack_stream(net, layers, aux);
  if (rc) return rc;
  return net.n_blocks == blocks_per_tile<NET_MOTION>() ? STNERF_OK : STNERF_EINVAL;
}

// D = A * B^T for random fp16 A (128x64), B (128x64) through the tensor-core path; returns max |D - reference|.
int tc_selftest(float* max_err_host) {
  std::vector<__half> A(128 * 64), B(128 * 64);
  std::vector<float> Af(128 * 64), Bf(128 * 64);
  uint32_t s = 12345u;
  auto rnd = [&s]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
  for (int i = 0; i < 128 * 64; ++i) {
    A[i] = __float2half_rn(rnd()); Af[i] = __half2float(A[i]);
    B[i] = __float2half_rn(rnd()); Bf[i] = __half2float(B[i]);
  }
  std::vector<uint8_t> blk(BLOCK_BYTES, 0);
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < 64; ++c) memcpy(blk.data() + sw128_offset(r, c), &B[r * 64 + c], 2);
  __half* dA = nullptr; uint8_t* dB = nullptr; float* dD = nullptr;
  STNERF_CUDA(cudaMalloc((void**)&dA, A.size() * 2));
  STNERF_CUDA(cudaMalloc((void**)&dB, BLOCK_BYTES));
  STNERF_CUDA(cudaMalloc((void**)&dD, 128 * 128 * 4));
  STNERF_CUDA(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  STNERF_CUDA(cudaMemcpy(dB, blk.data(), BLOCK_BYTES, cudaMemcpyHostToDevice));
  const int smem = 2 * BLOCK_BYTES + 64;
  STNERF_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_selftest_kernel<<<1, 128, smem>>>(dA, dB, dD);
  STNERF_LAUNCH_CHECK();
  STNERF_CUDA(cudaDeviceSynchronize());
  std::vector<float> D(128 * 128);
  STNERF_CUDA(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  float worst = 0.f;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      double ref = 0;
      for (int k = 0; k < 64; ++k) ref += (double)Af[m * 64 + k] * Bf[n * 64 + k];
      worst = fmaxf(worst, fabsf((float)ref - D[m * 128 + n]));
    }
  *max_err_host = worst;
  return STNERF_OK;
}

template <int NET>
static int launch_tc(const TcParams& P, int num_sms, cudaStream_t st) {
  static bool configured = false;
  const int smem = SM_TOTAL;
  if (!configured) {
    STNERF_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<NET>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  mlp_tc_kernel<NET><<<num_sms, NTHREADS, smem, st>>>(P);
  STNERF_LAUNCH_CHECK();
  return STNERF_OK;
}

int tc_launch_spacenet(const PointSrc& src, const TcNet& net, const SpaceNetW&, int precision, float* raw, float* rgb_out,
                       float* sigma_out, int num_sms, cudaStream_t st) {
  if (!net.blob) return STNERF_ENOWEIGHTS;
  TcParams P;
  memset(&P, 0, sizeof(P));
  P.src = src; P.wblocks = (const uint8_t*)net.blob; P.aux = net.aux;
  P.exact = precision == STNERF_PREC_TC_3XF16; P.use_time = net.use_time;
  P.raw = raw; P.rgb_out = rgb_out; P.sigma_out = sigma_out; P.lerp_force = 0;
  return launch_tc<NET_SPACE>(P, num_sms, st);
}

int tc_launch_motionnet(const PointSrc& src, const TcNet& net, const MotionNetW&, int precision, const int* lerp_flag_dev,
                        int lerp_force, float* xyz_out, float* flow_out, int num_sms, cudaStream_t st) {
  if (!net.blob) return STNERF_ENOWEIGHTS;
  TcParams P;
  memset(&P, 0, sizeof(P));
  P.src = src; P.wblocks = (const uint8_t*)net.blob; P.aux = net.aux;
  P.exact = precision == STNERF_PREC_TC_3XF16;
  P.xyz_out = xyz_out; P.flow_out = flow_out; P.lerp_flag = lerp_flag_dev; P.lerp_force = lerp_force;
  return launch_tc<NET_MOTION>(P, num_sms, st);
}

}  // namespace stnerf

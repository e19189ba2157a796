// dnn.cu -- K2: DNN-HMM forward for all frames of a batch on the 5th-generation tensor cores.
//
// Stands in for dnn_calc_outprob (libsent/src/phmm/calc_dnn.c:774-868): per frame a stack of
//   dst = W.src + b   (calc_dnn_fma.c:18-95 / sub1 calc_dnn.c:509-523)
//   hidden: logistic through a 320001-entry table with clamps (calc_dnn.c:342-369)
//   output: linear, then log-softmax via addlog_array and "- log10 prior" (calc_dnn.c:862-865)
// The reference does this one frame at a time (a GEMV stack, 130 MB of weights per frame); here
// all frames of the batch go through one GEMM per layer:  C[frames x out] = A[frames x in] . W^T.
//
// Precision.  The parity tolerance is 1e-4 relative on log-likelihoods; a single bf16/tf32 pass
// (8/10-bit mantissa) is ~1e-3.  Every operand is therefore split in two bf16 terms
// (x = hi + lo, 16 mantissa bits) and each k-block issues three MMAs into the same fp32
// accumulator in TMEM:  hi.hi + hi.lo + lo.hi   (the dropped lo.lo term is 2^-16 relative).
//
// Kernel anatomy (sm_100a), dnn_gemm_persistent<256> by default (dnn_gemm_kernel = the one-tile-per-CTA
// first version, JB200_DNN_KERNEL=0): 192 threads = warp 0 TMA producer, warp 1 tcgen05.mma issuer,
// warps 2-5 epilogue (each owns the TMEM lane quarter warp_idx%4).  Operand tiles
// 128 x 64 bf16 (K-major, 128-byte swizzle) arrive by cp.async.bulk.tensor (TMA) into a 3-stage
// shared-memory ring guarded by full/empty mbarriers; the 128 x 128 fp32 accumulator lives in
// TMEM; tcgen05.commit hands stages back to the producer and the finished tile to the epilogue,
// which reads it with tcgen05.ld, adds the bias, applies the reference's clamped table logistic and
// writes the next layer's operands already split into bf16 hi/lo.  The last layer's epilogue
// writes fp32 logits; a row kernel then does the log-softmax (with the reference's "drop terms more
// than 13.8 below the sum" rule) and subtracts the log10 prior.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <vector>
#include <cmath>

namespace jb200 {

static constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3;
static constexpr int TILE_BYTES = BM * BK * 2;                 // 16 KB (A and B tiles are the same size)
static constexpr int STAGE_BYTES = 4 * TILE_BYTES;             // A_hi A_lo B_hi B_lo
static constexpr int GEMM_THREADS = 192;
static constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
static constexpr int LOGISTIC_N = 320001;

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, 128-byte swizzle operand descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start>>4 [0,14) | LBO=1 [16,30) | SBO=1024>>4 [32,46) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_desc(const void *smem) {
  uint64_t d = (uint64_t)((smem_u32(smem) & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N=BN, M=BM
__device__ __forceinline__ uint32_t make_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// logistic_func, calc_dnn.c:362-369
__device__ __forceinline__ float logistic_ref(float x, const float *__restrict__ tbl) {
  if (x <= -8.0f) return 0.000334f;
  if (x >= 8.0f) return 0.999666f;
  const float t = __fadd_rn(x, 8.0f);
  const int idx = (int)__dadd_rn((double)__fmul_rn(t, 20000.0f), 0.5);
  return __ldg(tbl + idx);
}

struct GemmArgs {
  int M, N, K;                 // rows (frames), outputs, inputs
  const float *bias;           // [N]
  const float *logistic;       // table
  __nv_bfloat16 *out_hi, *out_lo; int ld_out;   // hidden layers: next operands [M][ld_out]
  float *logits; int ld_logits;                 // last layer: fp32 [M][ld_logits]
  int last;
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
dnn_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                const GemmArgs g) {
  extern __shared__ unsigned char dsm_raw[];
  unsigned char *dsm = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(dsm_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *full = reinterpret_cast<uint64_t *>(dsm + STAGES * STAGE_BYTES);
  uint64_t *empty = full + STAGES;
  uint64_t *tmem_full = empty + STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int nkb = (g.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) {   // TMEM: 128 fp32 columns for the 128x128 accumulator
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        unsigned char *st = dsm + s * STAGE_BYTES;
        mbar_expect_tx(&full[s], STAGE_BYTES);
        tma_load_2d(st, &map_a_hi, &full[s], kb * BK, m0);
        tma_load_2d(st + TILE_BYTES, &map_a_lo, &full[s], kb * BK, m0);
        tma_load_2d(st + 2 * TILE_BYTES, &map_b_hi, &full[s], kb * BK, n0);
        tma_load_2d(st + 3 * TILE_BYTES, &map_b_lo, &full[s], kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer: three bf16 products per k-step into one fp32 accumulator =====
      const uint32_t idesc = make_idesc();
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        unsigned char *st = dsm + s * STAGE_BYTES;
        const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + TILE_BYTES);
        const uint64_t b_hi = make_desc(st + 2 * TILE_BYTES), b_lo = make_desc(st + 3 * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; k++) {
          const uint64_t adv = (uint64_t)(k * 32 >> 4);      // 16 bf16 = 32 bytes along K inside the swizzle atom
          tc_mma_bf16(tmem_base, a_hi + adv, b_hi + adv, idesc, (kb | k) ? 1u : 0u);
          tc_mma_bf16(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
          tc_mma_bf16(tmem_base, a_lo + adv, b_hi + adv, idesc, 1u);
        }
        tc_commit(&empty[s]);                                  // frees the stage when these MMAs retire
      }
      tc_commit(tmem_full);                                    // accumulator complete
    }
  } else {
    // ===== epilogue: TMEM -> registers -> bias / logistic / split -> global =====
    const int q = warp & 3;                                    // TMEM lane quarter this warp may access
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN / 32; c++) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
      const int col0 = n0 + c * 32;
      if (row < g.M) {
        if (g.last) {
          float *dst = g.logits + (size_t)row * g.ld_logits + col0;
#pragma unroll
          for (int i = 0; i < 32; i++)
            if (col0 + i < g.N) dst[i] = __uint_as_float(r[i]) + __ldg(g.bias + col0 + i);
        } else {
          __align__(16) __nv_bfloat16 hi[32], lo[32];
#pragma unroll
          for (int i = 0; i < 32; i++) {
            float v = 0.0f;
            if (col0 + i < g.N) v = logistic_ref(__uint_as_float(r[i]) + __ldg(g.bias + col0 + i), g.logistic);
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            hi[i] = h;
            lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
          }
          // ld_out is a multiple of 8 and col0 of 32: 16-byte aligned vector stores; columns beyond N
          // (up to ld_out) are written as zeros so the next layer's K tail is clean
          uint4 *dh = reinterpret_cast<uint4 *>(g.out_hi + (size_t)row * g.ld_out + col0);
          uint4 *dl = reinterpret_cast<uint4 *>(g.out_lo + (size_t)row * g.ld_out + col0);
#pragma unroll
          for (int v4 = 0; v4 < 4; v4++)
            if (col0 + v4 * 8 < g.ld_out) { dh[v4] = reinterpret_cast<const uint4 *>(hi)[v4]; dl[v4] = reinterpret_cast<const uint4 *>(lo)[v4]; }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(128u) : "memory");
  }
}

// ---- persistent variant -----------------------------------------------------------------------------
// One CTA per SM walks the output tiles (column blocks fastest, so the CTAs that run side by side share the
// activation rows in L2); the accumulator is double-buffered in TMEM, so the four epilogue warps drain tile i
// (tcgen05.ld, bias, table logistic, bf16 hi/lo split, stores) while the MMA warp is already filling tile i+1
// and the TMA warp runs ahead through the shared-memory ring.  BN_ = 128 (3 stages of 64 KB) or 256 (2 stages of
// 96 KB: a 128x256 tile moves 1.5x the bytes for 2x the flops -- the kernel is bound by the L2 -> shared-memory
// path).
template <int BN_>
struct PersistentCfg {
  static constexpr int STAGES_ = (BN_ == 128) ? 3 : 2;
  static constexpr int B_TILE = BN_ * BK * 2;
  static constexpr int STAGE = 2 * TILE_BYTES + 2 * B_TILE;     // A_hi A_lo B_hi B_lo
  static constexpr int SMEM = STAGES_ * STAGE + 1024 + 256;
};

template <int BN_>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
dnn_gemm_persistent(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                    const GemmArgs g) {
  using Cfg = PersistentCfg<BN_>;
  constexpr int S = Cfg::STAGES_;
  extern __shared__ unsigned char dsm_raw[];
  unsigned char *dsm = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(dsm_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *full = reinterpret_cast<uint64_t *>(dsm + S * Cfg::STAGE);
  uint64_t *empty = full + S;
  uint64_t *tmem_full = empty + S;          // [2]
  uint64_t *tmem_empty = tmem_full + 2;     // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (g.K + BK - 1) / BK;
  const int n_nblk = (g.N + BN_ - 1) / BN_, n_mblk = (g.M + BM - 1) / BM;
  const int n_tiles = n_nblk * n_mblk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
    mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BN_)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_nblk) * BM, n0 = (tile % n_nblk) * BN_;
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          unsigned char *st = dsm + s * Cfg::STAGE;
          mbar_expect_tx(&full[s], Cfg::STAGE);
          tma_load_2d(st, &map_a_hi, &full[s], kb * BK, m0);
          tma_load_2d(st + TILE_BYTES, &map_a_lo, &full[s], kb * BK, m0);
          unsigned char *bh = st + 2 * TILE_BYTES, *bl = bh + Cfg::B_TILE;
#pragma unroll
          for (int h = 0; h < BN_ / 128; h++) {                 // the weight maps have 128-row boxes
            tma_load_2d(bh + h * TILE_BYTES, &map_b_hi, &full[s], kb * BK, n0 + h * 128);
            tma_load_2d(bl + h * TILE_BYTES, &map_b_lo, &full[s], kb * BK, n0 + h * 128);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN_ >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int it = 0, i = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, i++) {
        const int b = i & 1;
        mbar_wait(&tmem_empty[b], (uint32_t)((i >> 1) & 1) ^ 1u);    // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(b * BN_);
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          unsigned char *st = dsm + s * Cfg::STAGE;
          const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + TILE_BYTES);
          const uint64_t b_hi = make_desc(st + 2 * TILE_BYTES), b_lo = make_desc(st + 2 * TILE_BYTES + Cfg::B_TILE);
#pragma unroll
          for (int k = 0; k < BK / 16; k++) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            tc_mma_bf16(acc, a_hi + adv, b_hi + adv, idesc, (kb | k) ? 1u : 0u);
            tc_mma_bf16(acc, a_hi + adv, b_lo + adv, idesc, 1u);
            tc_mma_bf16(acc, a_lo + adv, b_hi + adv, idesc, 1u);
          }
          tc_commit(&empty[s]);
        }
        tc_commit(&tmem_full[b]);
      }
    }
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3;
    int i = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, i++) {
      const int b = i & 1;
      const int m0 = (tile / n_nblk) * BM, n0 = (tile % n_nblk) * BN_;
      mbar_wait(&tmem_full[b], (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN_ / 32; c++) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * BN_ + c * 32), r);
        const int col0 = n0 + c * 32;
        if (row < g.M) {
          if (g.last) {
            float *dst = g.logits + (size_t)row * g.ld_logits + col0;
#pragma unroll
            for (int e = 0; e < 32; e++)
              if (col0 + e < g.N) dst[e] = __uint_as_float(r[e]) + __ldg(g.bias + col0 + e);
          } else {
            __align__(16) __nv_bfloat16 hi[32], lo[32];
#pragma unroll
            for (int e = 0; e < 32; e++) {
              float v = 0.0f;
              if (col0 + e < g.N) v = logistic_ref(__uint_as_float(r[e]) + __ldg(g.bias + col0 + e), g.logistic);
              const __nv_bfloat16 h = __float2bfloat16_rn(v);
              hi[e] = h;
              lo[e] = __float2bfloat16_rn(v - __bfloat162float(h));
            }
            uint4 *dh = reinterpret_cast<uint4 *>(g.out_hi + (size_t)row * g.ld_out + col0);
            uint4 *dl = reinterpret_cast<uint4 *>(g.out_lo + (size_t)row * g.ld_out + col0);
#pragma unroll
            for (int v4 = 0; v4 < 4; v4++)
              if (col0 + v4 * 8 < g.ld_out) { dh[v4] = reinterpret_cast<const uint4 *>(hi)[v4]; dl[v4] = reinterpret_cast<const uint4 *>(lo)[v4]; }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[b]);               // 4 arrivals (one per epilogue warp) free the accumulator
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)(2 * BN_)) : "memory");
  }
}

// fp32 [M][K] -> bf16 hi/lo [M][ld] (zero padded)
__global__ void split_bf16_kernel(const float *__restrict__ src, int M, int K, __nv_bfloat16 *hi, __nv_bfloat16 *lo, int ld) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * ld) return;
  const int r = (int)(idx / ld), c = (int)(idx % ld);
  const float v = (c < K) ? src[(size_t)r * K + c] : 0.0f;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[idx] = h;
  lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// log-softmax + prior, one block per frame (calc_dnn.c:862-865 with addlog_array's drop rule:
// terms more than LOG_ADDMIN below the sum do not contribute, addlog.c:116)
__global__ void __launch_bounds__(256)
dnn_softmax_kernel(const float *__restrict__ logits, int ld_logits, int N, const float *__restrict__ prior,
                   float *__restrict__ rows, int row_stride) {
  __shared__ float s_red[8];
  __shared__ float s_val;
  const int t = blockIdx.x, tid = threadIdx.x;
  const float *x = logits + (size_t)t * ld_logits;
  auto block_max = [&](float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) { float w = (tid < 8) ? s_red[tid] : -INFINITY; for (int o = 4; o > 0; o >>= 1) w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, o)); if (tid == 0) s_val = w; }
    __syncthreads();
    const float r = s_val; __syncthreads(); return r;
  };
  auto block_sum = [&](float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) { float w = (tid < 8) ? s_red[tid] : 0.0f; for (int o = 4; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o); if (tid == 0) s_val = w; }
    __syncthreads();
    const float r = s_val; __syncthreads(); return r;
  };
  float mx = -INFINITY;
  for (int i = tid; i < N; i += 256) mx = fmaxf(mx, x[i]);
  mx = block_max(mx);
  float s = 0.0f;
  for (int i = tid; i < N; i += 256) s += expf(x[i] - mx);
  s = block_sum(s);
  const float lse1 = mx + logf(s);
  const float cut = lse1 + (float)JB200_LOG_ADDMIN;
  float s2 = 0.0f;
  for (int i = tid; i < N; i += 256) { const float v = x[i]; if (v >= cut) s2 += expf(v - mx); }
  s2 = block_sum(s2);
  const float lse = mx + logf(s2);
  float *out = rows + (size_t)t * row_stride;
  for (int i = tid; i < N; i += 256)
    out[i] = (float)(JB200_INV_LOG_TEN * (double)(x[i] - lse) - (double)__ldg(prior + i));
}

}  // namespace jb200

namespace jb200 {
// dnn_cluster.cu (experimental, JB200_DNN_KERNEL=2)
int dnn_launch_cluster2(const CUtensorMap &ma_hi, const CUtensorMap &ma_lo, const CUtensorMap &mw_hi, const CUtensorMap &mw_lo,
                        int M, int N, int K, const float *bias, const float *logistic, __nv_bfloat16 *out_hi, __nv_bfloat16 *out_lo,
                        int ld_out, float *logits, int ld_logits, int last, int n_sm, cudaStream_t st);
}

// =============================================================================================
using namespace jb200;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct DnnLayerDev {
  int in = 0, out = 0, ld_in = 0;          // ld_in: K padded to a multiple of 8 (16-byte row pitch)
  __nv_bfloat16 *w_hi = nullptr, *w_lo = nullptr;   // [out][ld_in]
  float *bias = nullptr;
  CUtensorMap map_w_hi, map_w_lo;
};

struct jb200_dnn {
  int device = 0, n_layers = 0, in_dim = 0, out_dim = 0, row_stride = 0;
  std::vector<DnnLayerDev> L;
  float *d_prior = nullptr, *d_logistic = nullptr;
  PFN_encodeTiled encode = nullptr;
  cudaStream_t stream = nullptr;
  // batch buffers
  int cap_frames = 0, max_width = 0, ld_logits = 0, n_sm = 0, variant = 256;
  float *d_in = nullptr, *d_logits = nullptr, *d_rows = nullptr;
  __nv_bfloat16 *act_hi[2] = {nullptr, nullptr}, *act_lo[2] = {nullptr, nullptr};
};

static int make_map(jb200_dnn *h, CUtensorMap *map, void *base, int rows, int cols_ld, int cols_valid) {
  // 2-D bf16 tensor [rows][cols_ld], box = {64 columns (128 B), 128 rows}, 128-byte swizzle, OOB -> zeros
  cuuint64_t gdim[2] = {(cuuint64_t)cols_valid, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols_ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%d ld=%d", (int)r, rows, cols_ld); return JB200_ERR_CUDA; }
  return JB200_OK;
}

extern "C" void jb200_dnn_destroy(jb200_dnn *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (auto &l : h->L) { cudaFree(l.w_hi); cudaFree(l.w_lo); cudaFree(l.bias); }
  cudaFree(h->d_prior); cudaFree(h->d_logistic); cudaFree(h->d_in); cudaFree(h->d_logits); cudaFree(h->d_rows);
  for (int i = 0; i < 2; i++) { cudaFree(h->act_hi[i]); cudaFree(h->act_lo[i]); }
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" int jb200_dnn_create(const jb200_dnn_desc *d, int device, jb200_dnn **out) {
  if (!d || !out || d->n_layers < 1 || d->n_layers > JB200_DNN_MAX_LAYERS) { set_error("jb200_dnn_create: bad argument"); return JB200_ERR_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_error("no CUDA device (libjb200 has no CPU fallback)"); return JB200_ERR_NODEVICE; }
  if (device < 0 || device >= ndev) { set_error("device %d out of range", device); return JB200_ERR_ARG; }
  JB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  JB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) { set_error("device is sm_%d%d; tcgen05 needs sm_100a", prop.major, prop.minor); return JB200_ERR_NODEVICE; }
  jb200_dnn *h = new jb200_dnn();
  h->device = device; h->n_layers = d->n_layers; h->in_dim = d->in_dim; h->out_dim = d->out_dim;
  h->row_stride = (d->out_dim + 3) & ~3;
  {
    void *fn = nullptr; cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) { set_error("cuTensorMapEncodeTiled not available"); delete h; return JB200_ERR_CUDA; }
    h->encode = (PFN_encodeTiled)fn;
  }
  JB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->L.resize(d->n_layers);
  h->max_width = (d->in_dim + 7) & ~7;
  for (int l = 0; l < d->n_layers; l++) {
    DnnLayerDev &L = h->L[l];
    L.in = d->layer_in[l]; L.out = d->layer_out[l]; L.ld_in = (L.in + 7) & ~7;
    if (l > 0 && L.in != d->layer_out[l - 1]) { set_error("layer %d input %d != previous output %d", l, L.in, d->layer_out[l - 1]); jb200_dnn_destroy(h); return JB200_ERR_ARG; }
    if (l + 1 < d->n_layers) h->max_width = std::max(h->max_width, (L.out + 7) & ~7);
    std::vector<__nv_bfloat16> hi((size_t)L.out * L.ld_in), lo((size_t)L.out * L.ld_in);
    for (int r = 0; r < L.out; r++)
      for (int c = 0; c < L.ld_in; c++) {
        const float v = (c < L.in) ? d->w[l][(size_t)r * L.in + c] : 0.0f;
        const __nv_bfloat16 b = __float2bfloat16_rn(v);
        hi[(size_t)r * L.ld_in + c] = b;
        lo[(size_t)r * L.ld_in + c] = __float2bfloat16_rn(v - __bfloat162float(b));
      }
    JB_CUDA(cudaMalloc(&L.w_hi, hi.size() * 2)); JB_CUDA(cudaMalloc(&L.w_lo, lo.size() * 2));
    JB_CUDA(cudaMemcpy(L.w_hi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice));
    JB_CUDA(cudaMemcpy(L.w_lo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
    JB_CUDA(cudaMalloc(&L.bias, sizeof(float) * L.out));
    JB_CUDA(cudaMemcpy(L.bias, d->b[l], sizeof(float) * L.out, cudaMemcpyHostToDevice));
    int rc = make_map(h, &L.map_w_hi, L.w_hi, L.out, L.ld_in, L.ld_in); if (rc) { jb200_dnn_destroy(h); return rc; }
    rc = make_map(h, &L.map_w_lo, L.w_lo, L.out, L.ld_in, L.ld_in); if (rc) { jb200_dnn_destroy(h); return rc; }
  }
  JB_CUDA(cudaMalloc(&h->d_prior, sizeof(float) * d->out_dim));
  JB_CUDA(cudaMemcpy(h->d_prior, d->state_prior, sizeof(float) * d->out_dim, cudaMemcpyHostToDevice));
  {
    // logistic_table_build, calc_dnn.c:350-360
    std::vector<float> tbl(LOGISTIC_N);
    for (int i = 0; i < LOGISTIC_N; i++) { double x = (double)i / 20000.0 - 8.0; tbl[i] = (float)(1.0 / (1.0 + exp(-x))); }
    JB_CUDA(cudaMalloc(&h->d_logistic, sizeof(float) * LOGISTIC_N));
    JB_CUDA(cudaMemcpy(h->d_logistic, tbl.data(), sizeof(float) * LOGISTIC_N, cudaMemcpyHostToDevice));
  }
  h->ld_logits = (d->out_dim + 3) & ~3;
  JB_CUDA(cudaFuncSetAttribute(dnn_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  JB_CUDA(cudaFuncSetAttribute(dnn_gemm_persistent<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, PersistentCfg<128>::SMEM));
  JB_CUDA(cudaFuncSetAttribute(dnn_gemm_persistent<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, PersistentCfg<256>::SMEM));
  h->n_sm = prop.multiProcessorCount;
  h->variant = getenv("JB200_DNN_KERNEL") ? atoi(getenv("JB200_DNN_KERNEL")) : 256;   // 0 one tile per CTA, 128 / 256 persistent, 2 = experimental 2-CTA cluster (dnn_cluster.cu)
  *out = h;
  return JB200_OK;
}

extern "C" int jb200_dnn_out_dim(const jb200_dnn *h) { return h ? h->out_dim : 0; }
extern "C" int jb200_dnn_in_dim(const jb200_dnn *h) { return h ? h->in_dim : 0; }

static int dnn_reserve(jb200_dnn *h, int T) {
  if (T <= h->cap_frames) return JB200_OK;
  cudaFree(h->d_in); cudaFree(h->d_logits); cudaFree(h->d_rows);
  for (int i = 0; i < 2; i++) { cudaFree(h->act_hi[i]); cudaFree(h->act_lo[i]); h->act_hi[i] = h->act_lo[i] = nullptr; }
  h->d_in = h->d_logits = h->d_rows = nullptr; h->cap_frames = 0;
  JB_CUDA(cudaMalloc(&h->d_in, sizeof(float) * (size_t)T * h->in_dim));
  JB_CUDA(cudaMalloc(&h->d_logits, sizeof(float) * (size_t)T * h->ld_logits));
  JB_CUDA(cudaMalloc(&h->d_rows, sizeof(float) * (size_t)T * h->row_stride));
  for (int i = 0; i < 2; i++) {
    JB_CUDA(cudaMalloc(&h->act_hi[i], 2 * (size_t)T * h->max_width));
    JB_CUDA(cudaMalloc(&h->act_lo[i], 2 * (size_t)T * h->max_width));
  }
  h->cap_frames = T;
  return JB200_OK;
}

namespace jb200 {
// d_in [T][in_dim] fp32 on device -> d_rows [T][row_stride] log10 pseudo-likelihoods
int dnn_forward_device(jb200_dnn *h, const float *d_in, int T, float *d_rows, int row_stride, cudaStream_t st) {
  if (T <= 0) return JB200_OK;
  JB_CUDA(cudaSetDevice(h->device));
  int rc = dnn_reserve(h, T); if (rc) return rc;
  int cur = 0;
  {
    const int ld = h->L[0].ld_in;
    const size_t tot = (size_t)T * ld;
    split_bf16_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_in, T, h->in_dim, h->act_hi[0], h->act_lo[0], ld);
    JB_LAUNCH_CHECK();
  }
  for (int l = 0; l < h->n_layers; l++) {
    DnnLayerDev &L = h->L[l];
    CUtensorMap ma_hi, ma_lo;
    rc = make_map(h, &ma_hi, h->act_hi[cur], T, L.ld_in, L.ld_in); if (rc) return rc;
    rc = make_map(h, &ma_lo, h->act_lo[cur], T, L.ld_in, L.ld_in); if (rc) return rc;
    GemmArgs g;
    g.M = T; g.N = L.out; g.K = L.ld_in; g.bias = L.bias; g.logistic = h->d_logistic;
    g.last = (l + 1 == h->n_layers) ? 1 : 0;
    g.out_hi = h->act_hi[cur ^ 1]; g.out_lo = h->act_lo[cur ^ 1];
    g.ld_out = g.last ? 0 : h->L[l + 1].ld_in;
    g.logits = h->d_logits; g.ld_logits = h->ld_logits;
    if (h->variant == 0) {
      dim3 grid((L.out + BN - 1) / BN, (T + BM - 1) / BM);
      dnn_gemm_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(ma_hi, ma_lo, L.map_w_hi, L.map_w_lo, g);
    } else if (h->variant == 128) {
      const int tiles = ((L.out + 127) / 128) * ((T + BM - 1) / BM);
      dnn_gemm_persistent<128><<<std::min(tiles, h->n_sm), GEMM_THREADS, PersistentCfg<128>::SMEM, st>>>(ma_hi, ma_lo, L.map_w_hi, L.map_w_lo, g);
    } else if (h->variant == 2) {
      rc = dnn_launch_cluster2(ma_hi, ma_lo, L.map_w_hi, L.map_w_lo, g.M, g.N, g.K, g.bias, g.logistic, g.out_hi, g.out_lo,
                               g.ld_out, g.logits, g.ld_logits, g.last, h->n_sm, st);
      if (rc) return rc;
      cur ^= 1;
      continue;
    } else {
      const int tiles = ((L.out + 255) / 256) * ((T + BM - 1) / BM);
      dnn_gemm_persistent<256><<<std::min(tiles, h->n_sm), GEMM_THREADS, PersistentCfg<256>::SMEM, st>>>(ma_hi, ma_lo, L.map_w_hi, L.map_w_lo, g);
    }
    JB_LAUNCH_CHECK();
    cur ^= 1;
  }
  dnn_softmax_kernel<<<T, 256, 0, st>>>(h->d_logits, h->ld_logits, h->out_dim, h->d_prior, d_rows, row_stride);
  JB_LAUNCH_CHECK();
  return JB200_OK;
}
}  // namespace jb200

extern "C" int jb200_dnn_score_host(jb200_dnn *h, const float *in, int T, float *scores) {
  if (!h || !in || !scores) { set_error("jb200_dnn_score_host: null argument"); return JB200_ERR_ARG; }
  if (T <= 0) return JB200_OK;
  JB_CUDA(cudaSetDevice(h->device));
  int rc = dnn_reserve(h, T); if (rc) return rc;
  JB_CUDA(cudaMemcpyAsync(h->d_in, in, sizeof(float) * (size_t)T * h->in_dim, cudaMemcpyHostToDevice, h->stream));
  rc = dnn_forward_device(h, h->d_in, T, h->d_rows, h->row_stride, h->stream); if (rc) return rc;
  JB_CUDA(cudaMemcpy2DAsync(scores, sizeof(float) * h->out_dim, h->d_rows, sizeof(float) * h->row_stride, sizeof(float) * h->out_dim, T,
                            cudaMemcpyDeviceToHost, h->stream));
  JB_CUDA(cudaStreamSynchronize(h->stream));
  return JB200_OK;
}

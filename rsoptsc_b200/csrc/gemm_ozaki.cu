// FP64-equivalent dense contractions on the 5th-generation tensor cores (tcgen05, kind::i8).
//
// tcgen05.mma has no f64 kind and FP32 accumulation cannot reach the accuracy the Gram route of
// the SVT needs.  The Ozaki split-integer scheme does: every operand row is
// scaled by a power of two and cut into S signed 7-bit slices (int8), all slice-pair products
// are EXACT in the int32 TMEM accumulators, and the FP64 result is the power-of-two weighted sum
// of the group sums  sum_{t+u=p} Qa_t Qb_u^T  (p = 2..S+1; lower-order pairs are below 2^-7S).
//
// Kernel shape (one CTA per SM, persistent over 128 x 128 output tiles):
//   warp 0      TMA producer: per 32-byte k-block ONE 3-D box per operand brings all S slices of
//               the A rows and B rows ([slice][row][32 B], SWIZZLE_32B), so every slice is fetched
//               once per k-block and reused by all its pairs (4.5x less L2->SM traffic than
//               streaming the 36 pair-GEMMs one after another)
//   warp 1      MMA issuer (one elected thread): tcgen05.mma.kind::i8 M128 N128 K32, the S/2 group
//               accumulators of a pass live side by side in TMEM (4 x 128 columns = 512)
//   warps 2-9   epilogue (two warps per TMEM lane quarter, 64 columns each): tcgen05.ld, int32 -> FP64, scale by
//               2^(exA_i + exB_j - 7p), store / read-modify-write C with all loads of a 32-column chunk in flight
// Two passes per tile: groups p > S/2+1 (need all slices), then the S/2 low groups (slices 1..S/2).
#include <cuda.h>

#include "gemm.h"
#include "internal.h"

namespace rs {
namespace oz {

constexpr int BM = 128, BN = 128, BK = 32;  // BK bytes == int8 elements per k-block == UMMA K
constexpr int MAXS = 8;
constexpr int STAGES = 3;
constexpr int SLICE_TILE = BM * BK;               // 4 KB: one slice of one operand in a stage
constexpr int OPER_BYTES = MAXS * SLICE_TILE;     // 32 KB
constexpr int STAGE_BYTES = 2 * OPER_BYTES;       // 64 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int EPI_WARPS = THREADS / 32 - 2;
constexpr int ROWPAD = 128;
constexpr int EX_BAD = 1 << 20;  // == EX_NONFINITE (slicing section): operand row with NaN / Inf

struct Params {
  double* C;
  const int* exA;
  const int* exB;
  const int2* tiles;
  int num_tiles, M, N, ldc, kblocks, S, accumulate;
  int b_kshift;  // B's k-block = (A's k-block + b_kshift) mod kblocks: [V | W] against [W | V] from ONE sliced operand
  int negate;    // C -= A B^T instead of +=
};

// ---- PTX helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1ll << 28)) __trap();  // a pipeline bug must fault, not hang the GPU
  }
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::i8 (s8 x s8 -> s32).  The 64-bit shared-memory matrix
// descriptors are given as (lo, hi) words: hi is constant for the whole kernel and lo advances by a
// compile-time offset per slice, so the single issuing thread spends ~6 instructions per MMA (the
// tensor pipe needs a new M128 N128 K32 MMA every 64 cycles).
__device__ __forceinline__ void umma_i8_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major operand tile in SWIZZLE_32B layout (rows at a 32-byte
// pitch, 8-row atoms of 256 B):  bits [0,14) start address >> 4 | [32,46) stride byte offset >> 4 =
// 256 >> 4 | [46,48) version = 1 (sm_100) | [61,64) layout type = 6 (SWIZZLE_32B).  The kernel builds
// the low word per slice ((addr & 0x3FFFF) >> 4) and keeps the high word constant (desc_hi).
// kind::i8 instruction descriptor: D = S32, A = B = signed 8 bit, both K-major, N = 128, M = 128
__device__ __forceinline__ uint32_t make_idesc() {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

#define TMEM_LD_32x32b_X32(taddr, r)                                                                          \
  asm volatile(                                                                                               \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                               \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, " \
      "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                             \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),       \
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),            \
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),            \
        "=r"(r[30]), "=r"(r[31])                                                                              \
      : "r"(taddr)                                                                                            \
      : "memory")

#define TMEM_LD_32x32b_X16(taddr, r)                                                                          \
  asm volatile(                                                                                               \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                               \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                        \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),       \
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) \
      : "r"(taddr)                                                                                            \
      : "memory")

__device__ __forceinline__ double pow2(int e) {
  e = max(-1022, min(1023, e));
  return __hiloint2double((e + 1023) << 20, 0);
}

// pass 0: groups p = S/2+2 .. S+1 (all slices); pass 1: groups p = 2 .. S/2+1 (slices 1..S/2)
__device__ __forceinline__ void pass_groups(int S, int pass, int& p_lo, int& p_hi) {
  if (pass == 0) { p_lo = S / 2 + 2; p_hi = S + 1; } else { p_lo = 2; p_hi = S / 2 + 1; }
}
// k-blocks per int32 accumulation unit: npairs * K * 127^2 < 2^31
__device__ __host__ __forceinline__ int max_kblocks(int S) { return (133000 / S) / BK; }

// All MMAs of one k-block of one pass, fully unrolled (S and PASS are compile-time): group p
// accumulates the pairs (t, p - t) into TMEM columns [(p - p_lo) * BN, +BN).
template <int S, int PASS>
__device__ __forceinline__ void issue_stage(uint32_t tmem_base, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                            uint32_t acc_first) {
  constexpr int P_LO = PASS == 0 ? S / 2 + 2 : 2;
  constexpr int P_HI = PASS == 0 ? S + 1 : S / 2 + 1;
  constexpr uint32_t STEP = SLICE_TILE >> 4;  // descriptor address units (16 B)
#pragma unroll
  for (int p = P_LO; p <= P_HI; ++p) {
    const uint32_t d = tmem_base + (uint32_t)(p - P_LO) * BN;
#pragma unroll
    for (int tt = 1; tt < p; ++tt) {
      const int uu = p - tt;
      if (tt > S || uu > S) continue;
      umma_i8_lohi(d, a_lo + (uint32_t)(tt - 1) * STEP, b_lo + (uint32_t)(uu - 1) * STEP, hi, idesc,
                   tt == 1 ? acc_first : 1u);
    }
  }
}

template <int S>
__global__ void __launch_bounds__(THREADS, 1)
k_ozaki_gemm(const __grid_constant__ CUtensorMap tmA_full, const __grid_constant__ CUtensorMap tmA_half,
             const __grid_constant__ CUtensorMap tmB_full, const __grid_constant__ CUtensorMap tmB_half, Params P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - raw);
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  // barrier slots (8 B each): full[STAGES], empty[STAGES], tmem_full, tmem_empty ; then the TMEM base address
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t tmem_full = bars + 8u * (2 * STAGES);
  const uint32_t tmem_empty = bars + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 2));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: all 512 columns (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int kb_total = P.kblocks;
  const int kb_unit = max_kblocks(S);

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ti = blockIdx.x; ti < P.num_tiles; ti += gridDim.x) {
        const int2 t = P.tiles[ti];
        for (int kb0 = 0; kb0 < kb_total; kb0 += kb_unit) {
          const int kb1 = min(kb_total, kb0 + kb_unit);
          for (int pass = 0; pass < 2; ++pass) {
            const int nsl = pass == 0 ? S : S / 2;
            const uint32_t bytes = 2u * nsl * SLICE_TILE;
            for (int kb = kb0; kb < kb1; ++kb) {
              mbar_wait(empty_bar(stage), phase ^ 1);
              mbar_expect_tx(full_bar(stage), bytes);
              const uint32_t sa = base + stage * STAGE_BYTES;
              tma_load_3d(sa, pass == 0 ? &tmA_full : &tmA_half, full_bar(stage), kb * BK, t.x * BM, 0);
              int kbB = kb + P.b_kshift;
              if (kbB >= kb_total) kbB -= kb_total;
              tma_load_3d(sa + OPER_BYTES, pass == 0 ? &tmB_full : &tmB_half, full_bar(stage), kbB * BK, t.y * BN, 0);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc();
      // descriptor high word: SBO = 256 B (8 rows x 32 B), version 1, SWIZZLE_32B (see smem_desc_sw32)
      const uint32_t desc_hi = (256u >> 4) | (1u << 14) | (6u << 29);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int ti = blockIdx.x; ti < P.num_tiles; ti += gridDim.x) {
        for (int kb0 = 0; kb0 < kb_total; kb0 += kb_unit) {
          const int kb1 = min(kb_total, kb0 + kb_unit);
          for (int pass = 0; pass < 2; ++pass) {
            int p_lo, p_hi;
            pass_groups(S, pass, p_lo, p_hi);
            mbar_wait(tmem_empty, acc_phase ^ 1);  // epilogue has drained the previous unit
            tc_fence_after();
            for (int kb = kb0; kb < kb1; ++kb) {
              mbar_wait(full_bar(stage), phase);
              tc_fence_after();
              const uint32_t sa = base + stage * STAGE_BYTES;
              const uint32_t a_lo = (sa & 0x3FFFF) >> 4, b_lo = ((sa + OPER_BYTES) & 0x3FFFF) >> 4;
              const uint32_t acc_first = kb > kb0 ? 1u : 0u;
              if (pass == 0) issue_stage<S, 0>(tmem_base, a_lo, b_lo, desc_hi, idesc, acc_first);
              else           issue_stage<S, 1>(tmem_base, a_lo, b_lo, desc_hi, idesc, acc_first);
              umma_commit(empty_bar(stage));  // frees the smem stage when these MMAs retire
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(tmem_full);           // accumulators of this unit are complete
            acc_phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue (warps 2..9: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4) =====
    // Two warps per SM sub-partition: while one waits for its C loads or a TMEM read the other issues.
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    uint32_t acc_phase = 0;
    for (int ti = blockIdx.x; ti < P.num_tiles; ti += gridDim.x) {
      const int2 t = P.tiles[ti];
      const int row = t.x * BM + q * 32 + lane;
      const int exa = row < P.M ? P.exA[row] : 0;
      bool first = true;
      for (int kb0 = 0; kb0 < kb_total; kb0 += kb_unit) {
        for (int pass = 0; pass < 2; ++pass) {
          int p_lo, p_hi;
          pass_groups(S, pass, p_lo, p_hi);
          mbar_wait(tmem_full, acc_phase);
          tc_fence_after();
          // Read-modify-write only where C already holds something (the second pass of a unit, or accumulate
          // mode); the first pass of a plain GEMM stores without reading.  All 32 loads of a chunk are issued
          // before the TMEM reads so that they are in flight together (the epilogue runs one warp per SM
          // sub-partition: nothing else hides their latency), and every loop over j is fully unrolled so that
          // sum[] and cold[] stay in registers.
          const bool rmw = !(first && !P.accumulate);
          for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
            const int colbase = t.y * BN + c0;
            const int exb_lane = (colbase + lane < P.N) ? P.exB[colbase + lane] : 0;
            double* const dst0 = P.C + (size_t)row + (size_t)colbase * P.ldc;
            double cold[32];
            if (rmw) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                cold[j] = (row < P.M && colbase + j < P.N) ? __ldcg(dst0 + (size_t)j * P.ldc) : 0.0;
            }
            double sum[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[j] = 0.0;
            for (int p = p_hi; p >= p_lo; --p) {  // smallest contributions first
              const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((p - p_lo) * BN + c0);
              const double w = pow2(-7 * p);
              // int32 -> FP64 without the conversion pipe: the bit pattern 0x43300000'(x ^ 0x80000000) is the
              // double 2^52 + 2^31 + x exactly, so one integer XOR and one FP64 subtraction replace I2F.F64
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint32_t r[16];
                TMEM_LD_32x32b_X16(taddr + 16u * h, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  sum[16 * h + j] += (__hiloint2double(0x43300000, (int)(r[j] ^ 0x80000000u)) - 4503601774854144.0) * w;
              }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int exb = __shfl_sync(0xffffffffu, exb_lane, j);
              if (row < P.M && colbase + j < P.N) {
                double v = (exa >= EX_BAD || exb >= EX_BAD) ? __longlong_as_double(0x7ff8000000000000LL)
                                                             : sum[j] * pow2(exa + exb);
                if (P.negate) v = -v;
                if (rmw) v += cold[j];
                dst0[(size_t)j * P.ldc] = v;
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty);
          acc_phase ^= 1;
          first = false;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

// ---- slicing ------------------------------------------------------------------------------------
// Operand row r, contraction index k.  contig: x = X[k + r*ld] (K runs down a column of X);
// otherwise x = X[r + k*ld].  Q[t][r][k] (int8, k contiguous, padded with zeros).
// A row holding NaN or Inf gets the sentinel exponent EX_NONFINITE: the epilogue then writes NaN into every output
// that row takes part in, as an FP64 GEMM would, instead of a finite but meaningless number (fmax drops NaN).
constexpr int EX_NONFINITE = EX_BAD;
__device__ __forceinline__ double abs_or_inf(double x) {
  const double a = fabs(x);
  return a == a ? a : INFINITY;  // NaN -> Inf so that it survives the fmax reductions
}
__global__ void k_rowmax_contig(const double* __restrict__ X, int64_t ld, int R, int K, int* __restrict__ ex) {
  __shared__ double sm[8];
  const int r = blockIdx.x;
  double mx = 0;
  for (int k = threadIdx.x; k < K; k += 256) mx = fmax(mx, abs_or_inf(X[(int64_t)k + (int64_t)r * ld]));
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) mx = fmax(mx, sm[w]);
    int e = 0;
    if (mx > 0 && isfinite(mx)) frexp(mx, &e);  // mx = f * 2^e, f in [0.5, 1)
    ex[r] = isfinite(mx) ? e : EX_NONFINITE;
  }
}
__global__ void k_rowmax_strided(const double* __restrict__ X, int64_t ld, int R, int K, int* __restrict__ ex) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  double mx = 0;
#pragma unroll 8
  for (int k = 0; k < K; ++k) mx = fmax(mx, abs_or_inf(X[(int64_t)r + (int64_t)k * ld]));
  int e = 0;
  if (mx > 0 && isfinite(mx)) frexp(mx, &e);
  ex[r] = isfinite(mx) ? e : EX_NONFINITE;
}
__device__ __forceinline__ void slice_value(double x, int e, int S, int8_t* out) {
  double v = x * pow2(-e);  // |v| < 1, exact
#pragma unroll
  for (int t = 0; t < MAXS; ++t) {
    if (t >= S) break;
    v *= 128.0;
    double qd = (t == S - 1) ? rint(v) : trunc(v);
    qd = fmin(127.0, fmax(-127.0, qd));
    out[t] = (int8_t)(int)qd;
    v -= qd;
  }
}
__global__ void k_slice_contig(const double* __restrict__ X, int64_t ld, int R, int K, const int* __restrict__ ex, int S,
                               int64_t Rp, int64_t Kp, int8_t* __restrict__ Q) {
  const int r = blockIdx.y;             // grid covers the padded Rp x Kp extent: pads are zero-filled
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= Kp) return;
  int8_t q[MAXS];
#pragma unroll
  for (int t = 0; t < MAXS; ++t) q[t] = 0;
  if (r < R && k < K) slice_value(X[(int64_t)k + (int64_t)r * ld], ex[r], S, q);
  for (int t = 0; t < S; ++t) Q[((int64_t)t * Rp + r) * Kp + k] = q[t];
}
// 32 rows x 32 k per block through shared memory so that both the FP64 reads (along r) and the
// int8 writes (along k) are contiguous
__global__ void k_slice_strided(const double* __restrict__ X, int64_t ld, int R, int K, const int* __restrict__ ex,
                                int S, int64_t Rp, int64_t Kp, int8_t* __restrict__ Q) {
  __shared__ int8_t tile[MAXS][32][33];
  const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  for (int kk = threadIdx.y; kk < 32; kk += blockDim.y) {
    const int r = r0 + threadIdx.x, k = k0 + kk;
    int8_t q[MAXS];
#pragma unroll
    for (int t = 0; t < MAXS; ++t) q[t] = 0;
    if (r < R && k < K) slice_value(X[(int64_t)r + (int64_t)k * ld], ex[r], S, q);
    for (int t = 0; t < S; ++t) tile[t][threadIdx.x][kk] = q[t];
  }
  __syncthreads();
  for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
    const int r = r0 + rr, k = k0 + threadIdx.x;
    if (r < Rp && k < Kp)  // the grid covers the padded extent; out-of-range entries were sliced as 0
      for (int t = 0; t < S; ++t) Q[((int64_t)t * Rp + r) * Kp + k] = tile[t][rr][threadIdx.x];
  }
}

struct Operand {
  DevBuf<int8_t> Q;
  DevBuf<int> ex;
  int64_t R = 0, K = 0, Rp = 0, Kp = 0;
  int S = 0;
};

static void slice_operand(const double* X, int64_t ld, int64_t R, int64_t K, bool contig, int S, Operand& op) {
  op.R = R; op.K = K; op.S = S;
  op.Rp = (R + ROWPAD - 1) / ROWPAD * ROWPAD;
  op.Kp = (K + BK - 1) / BK * BK;
  op.Kp = (op.Kp + 127) / 128 * 128;  // keep rows 128-byte aligned for TMA strides
  op.Q.alloc((size_t)S * op.Rp * op.Kp);
  op.ex.alloc(op.Rp);
  Phase ph("ozaki_slice");
  RS_CUDA(cudaMemsetAsync(op.ex.p, 0, sizeof(int) * op.Rp, ctx().stream));
  RS_REQUIRE(op.Rp <= 65535, "ozaki: operand has too many rows for this launch shape");
  if (contig) {
    RS_LAUNCH(k_rowmax_contig, (unsigned)R, 256, 0, X, ld, (int)R, (int)K, op.ex.p);
    dim3 grid((unsigned)cdiv(op.Kp, 256), (unsigned)op.Rp);
    RS_LAUNCH(k_slice_contig, grid, 256, 0, X, ld, (int)R, (int)K, op.ex.p, S, op.Rp, op.Kp, op.Q.p);
  } else {
    RS_LAUNCH(k_rowmax_strided, (unsigned)cdiv(R, 64), 64, 0, X, ld, (int)R, (int)K, op.ex.p);
    dim3 grid((unsigned)cdiv(op.Kp, 32), (unsigned)cdiv(op.Rp, 32)), block(32, 8);
    RS_LAUNCH(k_slice_strided, grid, block, 0, X, ld, (int)R, (int)K, op.ex.p, S, op.Rp, op.Kp, op.Q.p);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    RS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    RS_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, "ozaki: cuTensorMapEncodeTiled not available");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}
static CUtensorMap make_map(const Operand& op, int depth) {
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)op.Kp, (cuuint64_t)op.Rp, (cuuint64_t)op.S};
  cuuint64_t strides[2] = {(cuuint64_t)op.Kp, (cuuint64_t)op.Kp * op.Rp};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BM, (cuuint32_t)depth};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)op.Q.p, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RS_REQUIRE(r == CUDA_SUCCESS, "ozaki: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return m;
}

// tile_rows (optional): keep only the tiles whose tile row i has tile_rows[i] != 0 (the sharded trailing update:
// a rank's own 128-row groups)
static void run(const Operand& a, const Operand& b, double* C, int64_t M, int64_t N, int64_t ldc, bool lower_only,
                bool accumulate, int b_kshift = 0, bool negate = false, const std::vector<char>* tile_rows = nullptr) {
  RS_REQUIRE(a.Kp == b.Kp && a.S == b.S, "ozaki: operand mismatch");
  const int tm = (int)cdiv(M, BM), tn = (int)cdiv(N, BN);
  // Rasterise in G x G super-blocks: the ~148 tiles in flight then share G panels of A rows and G
  // panels of B rows and advance through k together, so each panel is fetched from HBM once per
  // super-block and served from L2 to the other CTAs.
  const int G = 12;
  std::vector<int2> tiles;
  for (int j0 = 0; j0 < tn; j0 += G)
    for (int i0 = 0; i0 < tm; i0 += G)
      for (int j = j0; j < std::min(tn, j0 + G); ++j)
        for (int i = i0; i < std::min(tm, i0 + G); ++i)
          if ((!lower_only || (int64_t)j * BN <= (int64_t)i * BM + BM - 1) && (!tile_rows || (*tile_rows)[i]))
            tiles.push_back(make_int2(i, j));
  if (tiles.empty()) return;
  DevBuf<int2> d_tiles(tiles.size());
  RS_CUDA(cudaMemcpyAsync(d_tiles.p, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice, ctx().stream));
  Params P;
  P.C = C; P.exA = a.ex.p; P.exB = b.ex.p; P.tiles = d_tiles.p; P.num_tiles = (int)tiles.size();
  P.M = (int)M; P.N = (int)N; P.ldc = (int)ldc; P.kblocks = (int)(a.Kp / BK); P.S = a.S; P.accumulate = accumulate ? 1 : 0;
  P.b_kshift = b_kshift; P.negate = negate ? 1 : 0;
  RS_REQUIRE(b_kshift >= 0 && b_kshift < std::max(P.kblocks, 1), "ozaki: bad k shift");
  const CUtensorMap mAf = make_map(a, a.S), mAh = make_map(a, a.S / 2), mBf = make_map(b, b.S), mBh = make_map(b, b.S / 2);
  static bool attr_set = false;
  if (!attr_set) {
    RS_CUDA(cudaFuncSetAttribute(k_ozaki_gemm<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    RS_CUDA(cudaFuncSetAttribute(k_ozaki_gemm<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    RS_CUDA(cudaFuncSetAttribute(k_ozaki_gemm<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  const int grid = std::min<int>((int)tiles.size(), ctx().sm_count);
  {  // work accounting for bench.py's roofline (algorithmic: unpadded K, computed tiles)
    const double macs = (double)tiles.size() * BM * BN * (double)a.K;
    add_counter("ozaki_fp64_flops", 2.0 * macs);
    add_counter("ozaki_int8_ops", 2.0 * macs * (a.S * (a.S + 1) / 2));
  }
  Phase ph("ozaki_gemm");
  switch (a.S) {
    case 4: k_ozaki_gemm<4><<<grid, THREADS, SMEM_BYTES, ctx().stream>>>(mAf, mAh, mBf, mBh, P); break;
    case 6: k_ozaki_gemm<6><<<grid, THREADS, SMEM_BYTES, ctx().stream>>>(mAf, mAh, mBf, mBh, P); break;
    default: k_ozaki_gemm<8><<<grid, THREADS, SMEM_BYTES, ctx().stream>>>(mAf, mAh, mBf, mBh, P); break;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  RS_KERNEL_CHECK();
  // d_tiles and the operand slices go back to the caching allocator when the callers return; the
  // single library stream orders any reuse after this kernel.
}

}  // namespace oz

static int ozaki_slices() {
  static int s = -1;
  if (s < 0) {
    const char* e = getenv("RSOPTSC_OZAKI_SLICES");
    s = e ? atoi(e) : 8;
    if (s != 4 && s != 6 && s != 8) s = 8;
  }
  return s;
}

void ozaki_gram_AtA(const double* A, int64_t rows, int64_t cols, double* G) {
  oz::Operand op;
  oz::slice_operand(A, rows, cols, rows, /*contig=*/true, ozaki_slices(), op);
  oz::run(op, op, G, cols, cols, cols, /*lower_only=*/true, false);
}
void ozaki_gram_AAt(const double* A, int64_t rows, int64_t cols, double* G) {  // G (rows x rows, lower) = A A^T
  oz::Operand op;
  oz::slice_operand(A, rows, rows, cols, /*contig=*/false, ozaki_slices(), op);
  oz::run(op, op, G, rows, rows, rows, /*lower_only=*/true, false);
}
void ozaki_gemm_nn(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K) {
  oz::Operand a, b;
  oz::slice_operand(A, M, M, K, /*contig=*/false, ozaki_slices(), a);
  oz::slice_operand(B, K, N, K, /*contig=*/true, ozaki_slices(), b);
  oz::run(a, b, C, M, N, M, false, false);
}
void ozaki_gemm_nn_ld(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M,
                      int64_t N, int64_t K) {
  oz::Operand a, b;
  oz::slice_operand(A, lda, M, K, /*contig=*/false, ozaki_slices(), a);
  oz::slice_operand(B, ldb, N, K, /*contig=*/true, ozaki_slices(), b);
  oz::run(a, b, C, M, N, ldc, false, false);
}
void ozaki_gemm_nn_ld_acc(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M,
                          int64_t N, int64_t K) {
  oz::Operand a, b;
  oz::slice_operand(A, lda, M, K, /*contig=*/false, ozaki_slices(), a);
  oz::slice_operand(B, ldb, N, K, /*contig=*/true, ozaki_slices(), b);
  oz::run(a, b, C, M, N, ldc, false, /*accumulate=*/true);
}
void ozaki_gemm_tn_ld(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M,
                      int64_t N, int64_t K) {
  oz::Operand a, b;
  oz::slice_operand(A, lda, M, K, /*contig=*/true, ozaki_slices(), a);
  oz::slice_operand(B, ldb, N, K, /*contig=*/true, ozaki_slices(), b);
  oz::run(a, b, C, M, N, ldc, false, false);
}
void ozaki_gemm_nt(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K) {
  oz::Operand a, b;
  oz::slice_operand(A, M, M, K, /*contig=*/false, ozaki_slices(), a);
  oz::slice_operand(B, N, N, K, /*contig=*/false, ozaki_slices(), b);
  oz::run(a, b, C, M, N, M, false, false);
}
void ozaki_gemm_nt_ld(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M,
                      int64_t N, int64_t K) {
  oz::Operand a, b;
  oz::slice_operand(A, lda, M, K, /*contig=*/false, ozaki_slices(), a);
  oz::slice_operand(B, ldb, N, K, /*contig=*/false, ozaki_slices(), b);
  oz::run(a, b, C, M, N, ldc, false, false);
}
// ---- rank-2k update of the tridiagonalisation: C (nr x nr, lower triangle) -= V W^T + W V^T ---------------------
// = -[V | W] [W | V]^T, one accumulate-mode launch over the lower tiles.  128 contraction entries are only four
// k-blocks, so the launch is bound by its epilogue (read-modify-write of C) like the cuBLAS rank-2k update it replaces.
namespace oz {
// Slicing of the paired operand [V | W] (nr rows, 2 x 64 contraction entries; V, W: nr x k with k <= 64, the unused
// entries zero) straight from the two panels -- no packed copy.
constexpr int PAIR_HALF = 64, PAIR_K = 2 * PAIR_HALF;
// The two halves share one power-of-two exponent per row, so a row whose W entries dwarf its V entries (|v| <= 1,
// |w| ~ ||A||) would slice V with fewer significant bits.  The operand is therefore [2^es V | 2^-es W] with ONE global
// es = (exponent of max|W| - exponent of max|V|) / 2: the products V W^T and W V^T are unchanged (exact scaling),
// both halves have comparable magnitude.  mx[0], mx[1]: bit patterns of max|V|, max|W| (k_pair_absmax).
__global__ void k_pair_absmax(const double* __restrict__ V, long long ldv, const double* __restrict__ W, long long ldw, int nr, int k,
                              unsigned long long* __restrict__ mx) {
  double mv = 0, mw = 0;
  for (int c = blockIdx.y; c < k; c += gridDim.y)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += gridDim.x * blockDim.x) {
      mv = fmax(mv, abs_or_inf(V[i + (long long)c * ldv]));
      mw = fmax(mw, abs_or_inf(W[i + (long long)c * ldw]));
    }
  for (int o = 16; o > 0; o >>= 1) {
    mv = fmax(mv, __shfl_xor_sync(0xffffffffu, mv, o));
    mw = fmax(mw, __shfl_xor_sync(0xffffffffu, mw, o));
  }
  if ((threadIdx.x & 31) == 0) {  // non-negative doubles order like their bit patterns
    atomicMax(mx, (unsigned long long)__double_as_longlong(mv));
    atomicMax(mx + 1, (unsigned long long)__double_as_longlong(mw));
  }
}
__device__ __forceinline__ int pair_es(const unsigned long long* __restrict__ mx) {
  const double mv = __longlong_as_double((long long)mx[0]), mw = __longlong_as_double((long long)mx[1]);
  if (!(mv > 0) || !(mw > 0) || !isfinite(mv) || !isfinite(mw)) return 0;
  int ev, ew;
  frexp(mv, &ev);
  frexp(mw, &ew);
  return max(-500, min(500, (ew - ev) / 2));
}
// operand row r = matrix row r - lead: the first `lead` operand rows are zero (they align the operand with absolute
// 128-row groups; their outputs get +0)
__device__ __forceinline__ double pair_value(const double* __restrict__ V, long long ldv, const double* __restrict__ W, long long ldw,
                                             int r, int kk, int k, int lead, int es) {
  const int c = kk & (PAIR_HALF - 1);
  if (c >= k || r < lead) return 0.0;
  return kk < PAIR_HALF ? V[(r - lead) + (long long)c * ldv] * pow2(es) : W[(r - lead) + (long long)c * ldw] * pow2(-es);
}
__global__ void k_rowmax_pair(const double* __restrict__ V, long long ldv, const double* __restrict__ W, long long ldw, int nr, int k,
                              int lead, const unsigned long long* __restrict__ amax, int* __restrict__ ex) {
  __shared__ double sm[8][32];
  const int r = blockIdx.x * 32 + (threadIdx.x & 31), g = threadIdx.x >> 5;  // 32 rows x 8 column groups
  const int es = pair_es(amax);
  double mx = 0;
  if (r < nr)
    for (int kk = g; kk < PAIR_K; kk += 8) mx = fmax(mx, abs_or_inf(pair_value(V, ldv, W, ldw, r, kk, k, lead, es)));
  sm[g][threadIdx.x & 31] = mx;
  __syncthreads();
  if (g == 0 && r < nr) {
    for (int w = 1; w < 8; ++w) mx = fmax(mx, sm[w][threadIdx.x]);
    int e = 0;
    if (mx > 0 && isfinite(mx)) frexp(mx, &e);
    ex[r] = isfinite(mx) ? e : EX_NONFINITE;
  }
}
__global__ void k_slice_pair(const double* __restrict__ V, long long ldv, const double* __restrict__ W, long long ldw, int nr, int k,
                             int lead, const unsigned long long* __restrict__ amax, const int* __restrict__ ex, int S, int64_t Rp,
                             int8_t* __restrict__ Q) {
  __shared__ int8_t tile[MAXS][32][33];
  const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int es = pair_es(amax);
  for (int kl = threadIdx.y; kl < 32; kl += blockDim.y) {
    const int r = r0 + threadIdx.x, kk = k0 + kl;
    int8_t q[MAXS];
#pragma unroll
    for (int t = 0; t < MAXS; ++t) q[t] = 0;
    if (r < nr) slice_value(pair_value(V, ldv, W, ldw, r, kk, k, lead, es), ex[r], S, q);
    for (int t = 0; t < S; ++t) tile[t][threadIdx.x][kl] = q[t];
  }
  __syncthreads();
  for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
    const int r = r0 + rr;
    if (r < Rp)
      for (int t = 0; t < S; ++t) Q[((int64_t)t * Rp + r) * PAIR_K + k0 + threadIdx.x] = tile[t][rr][threadIdx.x];
  }
}
// small trailing blocks: one thread per lower-triangle entry
__global__ void k_syr2k_small(const double* __restrict__ V, long long ldv, const double* __restrict__ W, long long ldw, int nr, int k,
                              double* __restrict__ C, long long ldc) {
  const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
  if (i >= nr || j > i) return;
  double acc = 0;
  for (int t = 0; t < k; ++t)
    acc += V[i + (long long)t * ldv] * W[j + (long long)t * ldw] + W[i + (long long)t * ldw] * V[j + (long long)t * ldv];
  C[i + (long long)j * ldc] -= acc;
}
}  // namespace oz
// r0: absolute index of the first row of the trailing block; nranks > 1: update only the 128-row groups of rank
// `rank` (group g of absolute rows [128 g, 128 g + 128) belongs to rank g mod nranks, the ownership of
// sytrd_mg.cuh).  The caller gathers the next panel's columns afterwards (comm_gather_cyclic_rows).
void ozaki_syr2k_lower_sub(const double* V, int64_t ldv, const double* W, int64_t ldw, int64_t nr, int k, double* C, int64_t ldc,
                           int64_t r0, int nranks, int rank) {
  if (nr < 1 || k < 1) return;
  if (nr < 384) {
    dim3 grid((unsigned)cdiv(nr, 32), (unsigned)cdiv(nr, 8)), block(32, 8);
    RS_LAUNCH(oz::k_syr2k_small, grid, block, 0, V, (long long)ldv, W, (long long)ldw, (int)nr, k, C, (long long)ldc);
    return;
  }
  RS_REQUIRE(k <= oz::PAIR_HALF, "syr2k: panel wider than 64 columns");
  // sharded: the tile grid starts at the absolute multiple of 128 at or below r0 (lead = 0 or 64 zero rows)
  const int lead = nranks > 1 ? (int)(r0 % oz::BM) : 0;
  const int64_t rows = nr + lead;
  // [V | W] sliced ONCE: the second factor [W | V] is the same operand with its k-blocks rotated by two (64 entries),
  // and the minus sign goes into the epilogue
  oz::Operand a;
  a.R = rows; a.K = oz::PAIR_K; a.S = ozaki_slices();
  a.Rp = (rows + oz::ROWPAD - 1) / oz::ROWPAD * oz::ROWPAD;
  a.Kp = oz::PAIR_K;
  a.Q.alloc((size_t)a.S * a.Rp * a.Kp);
  a.ex.alloc(a.Rp + 4);  // + two 64-bit slots for max|V|, max|W| (a.Rp is a multiple of 128: 8-byte aligned)
  unsigned long long* amax = reinterpret_cast<unsigned long long*>(a.ex.p + a.Rp);
  {
    Phase ph("ozaki_slice");
    RS_CUDA(cudaMemsetAsync(a.ex.p, 0, sizeof(int) * (a.Rp + 4), ctx().stream));
    dim3 gmax((unsigned)std::min<int64_t>(cdiv(nr, 256), 16), (unsigned)std::min(k, 16));
    RS_LAUNCH(oz::k_pair_absmax, gmax, 256, 0, V, (long long)ldv, W, (long long)ldw, (int)nr, k, amax);
    RS_LAUNCH(oz::k_rowmax_pair, (unsigned)cdiv(rows, 32), 256, 0, V, (long long)ldv, W, (long long)ldw, (int)rows, k, lead, amax, a.ex.p);
    dim3 grid(oz::PAIR_K / 32, (unsigned)(a.Rp / 32)), block(32, 8);
    RS_LAUNCH(oz::k_slice_pair, grid, block, 0, V, (long long)ldv, W, (long long)ldw, (int)rows, k, lead, amax, a.ex.p, a.S, a.Rp, a.Q.p);
  }
  std::vector<char> mine;
  if (nranks > 1) {
    const int64_t g0 = (r0 - lead) / oz::BM;  // absolute group of tile row 0
    mine.resize((size_t)cdiv(rows, oz::BM));
    for (size_t i = 0; i < mine.size(); ++i) mine[i] = ((g0 + (int64_t)i) % nranks) == rank;
  }
  double* C0 = C - lead - (int64_t)lead * ldc;  // absolute (r0 - lead, r0 - lead): inside the matrix, never written there
  oz::run(a, a, C0, rows, rows, ldc, /*lower_only=*/true, /*accumulate=*/true, /*b_kshift=*/oz::PAIR_HALF / oz::BK, /*negate=*/true,
          nranks > 1 ? &mine : nullptr);
}
void ozaki_gemm_tn(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K) {
  oz::Operand a, b;
  oz::slice_operand(A, K, M, K, /*contig=*/true, ozaki_slices(), a);
  oz::slice_operand(B, K, N, K, /*contig=*/true, ozaki_slices(), b);
  oz::run(a, b, C, M, N, M, false, false);
}

}  // namespace rs

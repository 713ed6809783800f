// Householder tridiagonalisation of a dense symmetric matrix (lower triangle), the memory-bound
// first half of the symmetric eigensolver behind the SVT (a4: R/SimilarityM.R:142-155 needs the
// spectrum of A^T A), the PCA (a2) and the Laplacian spectrum (a14).
//
// Same factorisation and storage as LAPACK dsytrd(uplo='L') -- Q = H(1) ... H(n-1),
// H(i) = I - tau_i v_i v_i^T, v_i(i+1) = 1, v_i(i+2:n) stored below the subdiagonal of column i --
// so a LAPACK-convention back-transformation (ormtr) applies.  Design for B200:
//
//  * one persistent cooperative kernel per panel of NB columns, one 512-thread CTA per SM, with a
//    hand-rolled grid barrier + all-reduce instead of ~15 tiny launches per column (a
//    flag-with-data variant where every CTA polls every slot was measured slower: the polls
//    saturate L2);
//  * TWO grid synchronisations per column instead of the textbook three: v^T w, which fixes the
//    symmetric correction w -= (tau/2)(w^T v) v, is obtained as tau (v^T A22 v - 2 (W^T v).(V^T v))
//    from sums that the product phase accumulates anyway, so the correction, the final W column
//    and the update of the NEXT column all happen in one phase that reads the panel once;
//  * the trailing-matrix product y = A22 v reads ONLY the lower triangle (half the bytes of a
//    gemv).  The triangle is cut into strips of 64 rows x 16 columns, enumerated along block rows;
//    every warp of the grid owns one contiguous run of strips (static, equal to within one strip).
//    Lanes own rows (coalesced 256-byte column segments), row sums stay in registers along the
//    run, column sums go through a transposing shuffle butterfly;
//  * every partial sum has its own slot in a scratch array that is reduced in a fixed order, so
//    results are bit-reproducible;
//  * while a CTA waits in the all-reduce that precedes the product phase (DRAM idle), its warps
//    stage the first strip of their run in shared memory with cp.async;
//  * the rank-2NB trailing update between panels runs on the tcgen05 split-integer GEMM (Hooks::syr2k; the
//    stand-alone probes fall back to cublasDsyr2k).
//
// HBM traffic: sum_j 4 (n-j)^2 = (4/3) n^3 bytes for the products (the floor of any one-stage
// tridiagonalisation) + 16 n^3 / (3 NB) for the trailing updates.
#pragma once
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

namespace rs {
namespace sytrd {

constexpr int NB = 64;          // panel width
constexpr int THREADS = 512;    // 16 warps per CTA, one CTA per SM
constexpr int MAX_GRID = 512;   // upper bound on the cooperative grid
constexpr int TILE = 64;        // block edge of the triangle (rows per strip)
#ifndef RS_SYTRD_SW
#define RS_SYTRD_SW 16
#endif
constexpr int SW = RS_SYTRD_SW;  // strip width (columns): 16 or 8
constexpr int SPT = TILE / SW;  // strips per 64 x 64 block
constexpr int TPR = 8;          // threads per row in the row-wise phase
constexpr int RPT = NB / TPR;   // panel columns per thread in the row-wise phase
constexpr int KB = 24;          // partial sums of a row fetched per lane in one batch
constexpr int ROW_WARPS = THREADS / 32 - 1;        // warps that sweep rows; the last warp handles row j+1
constexpr int ROWS_PER_CTA = ROW_WARPS * (32 / TPR);
constexpr int DOT_CHUNK = 32 * SW;  // rows per panel dot-product unit (about the cost of one strip)

struct Params {
  double* A;
  long long lda;
  int n;
  int j0, pw;        // panel: columns j0 .. j0+pw-1
  double* W;         // n x NB, leading dimension n, rows indexed globally
  double* d;         // n
  double* e;         // n-1
  double* tau;       // n-1
  double* rowpart;   // (warps + mB + 1) x 64 : slot (g + b) = row sums of warp g over block row b
  double* colpart;   // mB x n : colpart[I*n + c] = sum over rows of block I of A[r,c] v[r]
  double* abuf;      // 2 x n  : updated column (double-buffered by column parity)
  double* pvec;      // chunks x 2 NB : partial W^T v and V^T v per chunk of DOT_CHUNK rows
  unsigned long long* slots;  // 4 MAX_GRID words of per-CTA partials (double-buffered) + arrival counter + generation flag
  long long* prof;   // optional cycle counters (null = off): [0..15] phases of CTA 0, [16 + cta] product phase
  int mB;            // ceil(n / 64)
};

__device__ __forceinline__ double warp_sum_fixed(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Grid-wide barrier + sum of two doubles per CTA.  Valid inputs in thread 0 of each CTA; every thread
// of every CTA returns the same two sums (fixed summation order).  All global writes made by any CTA
// before the call are visible to every CTA after it (release reduction on the arrival counter, acquire
// loads of the same counter; bar.sync extends both to the whole CTA).  All CTAs are co-resident
// (cooperative launch); the arrival counter is monotonic within a launch.  The per-CTA partials are
// double-buffered by generation parity: a CTA can only write generation g+2 after every CTA has
// arrived at g+1, i.e. after every CTA has finished reading generation g.
struct NoWork {
  __device__ void operator()() const {}
};
// `while_waiting` is run by warps 1.. of every CTA while thread 0 waits for the other CTAs: free issue slots
// for work that does not depend on the barrier (staging of the next product phase).
template <class F = NoWork>
__device__ __forceinline__ void grid_allreduce2(unsigned long long* slots, unsigned nblocks, unsigned& gen, double& v0,
                                                double& v1, double* s_out, F while_waiting = F()) {
  const unsigned tid = threadIdx.x;
  ++gen;
  double* part = reinterpret_cast<double*>(slots) + (size_t)(gen & 1u) * MAX_GRID * 2;  // double-buffered partials
  unsigned* bar = reinterpret_cast<unsigned*>(slots + 4 * MAX_GRID);                    // [0] arrivals, [1] generation
  __syncthreads();
  if (tid == 0) {
    part[2 * blockIdx.x] = v0;
    part[2 * blockIdx.x + 1] = v1;
    // fire-and-forget arrival, then poll the counter itself (one hop less than ticket + flag)
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    const unsigned target = gen * nblocks;
    unsigned spins = 0, seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
      if (++spins > (1u << 25)) __trap();  // never hang the device
    } while (seen < target);
  }
  if (tid >= 32) while_waiting();
  __syncthreads();
  if (tid < 32) {
    double a = 0, b = 0;
    // all partials of a lane are requested in one batch (one L2 round trip), then added in a fixed order
    const double2* part2 = reinterpret_cast<const double2*>(part);
    for (unsigned k0 = tid; k0 < nblocks; k0 += 32 * 8) {
      double2 t[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const unsigned k = k0 + 32u * q;
        t[q] = (k < nblocks) ? part2[k] : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) { a += t[q].x; b += t[q].y; }
    }
    a = warp_sum_fixed(a);
    b = warp_sum_fixed(b);
    if (tid == 0) { s_out[0] = a; s_out[1] = b; }
  }
  __syncthreads();
  v0 = s_out[0];
  v1 = s_out[1];
}

// CTA-wide sum of two values; result valid in thread 0.
__device__ __forceinline__ void cta_reduce2(double& a, double& b, double* s_tmp /* 2 x 16 */) {
  a = warp_sum_fixed(a);
  b = warp_sum_fixed(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();  // s_tmp may still be read from the previous use
  if (l == 0) { s_tmp[w] = a; s_tmp[16 + w] = b; }
  __syncthreads();
  if (threadIdx.x < 32) {
    a = threadIdx.x < (THREADS >> 5) ? s_tmp[threadIdx.x] : 0.0;
    b = threadIdx.x < (THREADS >> 5) ? s_tmp[16 + threadIdx.x] : 0.0;
    a = warp_sum_fixed(a);
    b = warp_sum_fixed(b);
  }
}

// Transposing butterfly: every lane holds SW partial column sums; on return every lane holds the
// complete sum of ONE column: SW = 16: index = bit4*8 + bit3*4 + bit2*2 + bit1 of the lane id (lane pairs
// agree); SW = 8: index = bit4*4 + bit3*2 + bit2 (groups of four lanes agree).
__device__ __forceinline__ double butterfly(double (&p)[SW], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
  double q4[4], q2[2], q1;
  if constexpr (SW == 16) {
    double q8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double send = b4 ? p[k] : p[k + 8];
      const double keep = b4 ? p[k + 8] : p[k];
      q8[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double send = b3 ? q8[k] : q8[k + 4];
      const double keep = b3 ? q8[k + 4] : q8[k];
      q4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double send = b2 ? q4[k] : q4[k + 2];
      const double keep = b2 ? q4[k + 2] : q4[k];
      q2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const double send = b1 ? q2[0] : q2[1];
    const double keep = b1 ? q2[1] : q2[0];
    q1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  } else {
    static_assert(SW == 8 || SW == 16, "strip width");
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double send = b4 ? p[k] : p[k + 4];
      const double keep = b4 ? p[k + 4] : p[k];
      q4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double send = b3 ? q4[k] : q4[k + 2];
      const double keep = b3 ? q4[k + 2] : q4[k];
      q2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const double send = b2 ? q2[0] : q2[1];
    const double keep = b2 ? q2[1] : q2[0];
    q1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
  }
  q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
  return q1;
}
// lanes that publish a column sum, and the column (inside the strip) they hold
__device__ __forceinline__ bool col_writer(int lane) { return SW == 16 ? !(lane & 1) : !(lane & 3); }
__device__ __forceinline__ int col_of_lane(int lane) { return SW == 16 ? ((lane >> 1) & 15) : ((lane >> 2) & 7); }

// Work units of the product phase, in this order: the strips of block row b = 0, 1, ... (block row b
// has SPT (b + 1) strips, the first at unit SPT b (b + 1) / 2), then the panel dot-product chunks.
// Unit u belongs to warp owner(u) = floor(u * nw / total) with nw = min(warps, total): contiguous
// runs, equal to within one unit, and every warp below nw owns at least one unit.
struct UnitMap {
  long long strips, total, nw;
  __device__ __forceinline__ long long first_of_row(int b) const { return (long long)SPT * b * (b + 1) / 2; }
  __device__ __forceinline__ long long start(long long g) const { return (g * total + nw - 1) / nw; }  // ceil
  __device__ __forceinline__ long long owner(long long u) const { return u * nw / total; }
};

// Between the end of a product phase and the next one (row phase + two grid synchronisations, ~10 us) the
// memory system idles.  Every warp uses that window to stage the FIRST strip of its next run in shared
// memory with cp.async (8 KB per warp, 128 KB per SM, ~19 MB over the grid: a fifth of an average column).
// The strip's unit id is returned (-1: nothing staged); the product phase compares it with the unit it is
// about to process, so a mismatch only costs the staging.  (An L2-only prefetch of the same strips was
// measured slower on B200.)
#ifndef RS_SYTRD_STAGE
#define RS_SYTRD_STAGE 1
#endif
constexpr int STAGE_DOUBLES = TILE * SW;  // per warp
__device__ __forceinline__ long long stage_next_strip(const Params& P, int jnext, int jjnext, long long gwarp, long long nwarps,
                                                      int lane, double* sbuf) {
  if (!RS_SYTRD_STAGE || jnext + 1 >= P.n) return -1;
  const int n = P.n;
  const int Bmin = (jnext + 1) / TILE, mt = P.mB - Bmin;
  const int nch = (n - jnext - 1 + DOT_CHUNK - 1) / DOT_CHUNK;
  UnitMap um;
  um.strips = um.first_of_row(mt);
  um.total = um.strips + 2LL * jjnext * nch;
  um.nw = nwarps < um.total ? nwarps : um.total;
  if (gwarp >= um.nw) return -1;
  const long long u = um.start(gwarp);
  if (u >= um.strips || u >= um.start(gwarp + 1)) return -1;
  int b = (int)floor((sqrt(1.0 + 8.0 * (double)u / SPT) - 1.0) * 0.5);
  if (b < 0) b = 0;
  while (b > 0 && um.first_of_row(b) > u) --b;
  while (um.first_of_row(b + 1) <= u) ++b;
  const int k = (int)(u - um.first_of_row(b));
  const int I = Bmin + b, J = Bmin + k / SPT, pass = k % SPT;
  const int cbase = J * TILE + pass * SW;
  const int r0 = I * TILE + lane, r1 = r0 + 32;
  const unsigned s0 = (unsigned)__cvta_generic_to_shared(sbuf + lane), s1 = (unsigned)__cvta_generic_to_shared(sbuf + lane + 32);
#pragma unroll
  for (int c = 0; c < SW; ++c) {
    if (cbase + c < n) {
      const double* col = P.A + (long long)(cbase + c) * P.lda;
      if (r0 < n) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s0 + 8u * TILE * c), "l"(col + r0) : "memory");
      if (r1 < n) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s1 + 8u * TILE * c), "l"(col + r1) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  return u;
}

__device__ __forceinline__ bool more_cols(int jj, int pw) { return jj + 1 < pw; }

// Row i of the row phase, stage 1: V(i, t), W(i, t) for t = sub + u TPR < jj into vv / ww, and this
// lane's share of y(i) (the row and column partial sums of the product phase).  No dependence on the
// panel dot products, so these loads overlap their reduction.
__device__ __forceinline__ double row_load(const Params& P, const UnitMap& um, int i, bool live, int sub, int jj, int Bmin,
                                           double (&vv)[RPT], double (&ww)[RPT]) {
  const int n = P.n;
#pragma unroll
  for (int u = 0; u < RPT; ++u) {
    const int t = sub + u * TPR;
    const bool on = live && t < jj;
    vv[u] = on ? P.A[i + (long long)(P.j0 + t) * P.lda] : 0.0;
    ww[u] = on ? P.W[i + (long long)t * n] : 0.0;
  }
  double acc = 0;
  if (live) {
    // row parts: the warps whose runs intersect block row b, slot (g + b); column parts: block rows Ii..mB-1
    const int Ii = i / TILE, b = Ii - Bmin;
    const long long s0 = um.first_of_row(b);
    const int g0 = (int)um.owner(s0), g1 = (int)um.owner(s0 + (long long)SPT * (b + 1) - 1);
    const int nrow = g1 - g0 + 1, ncol = P.mB - Ii;
    const double* rp = P.rowpart + (size_t)(g0 + b) * TILE + (i - Ii * TILE);
    const double* cp = P.colpart + ((long long)(Ii - nrow) * n + i);  // k >= nrow -> block row Ii + (k - nrow)
    const int ntot = nrow + ncol;
    for (int k0 = sub; k0 < ntot; k0 += TPR * KB) {  // KB loads in flight per lane
      double t[KB];
#pragma unroll
      for (int q = 0; q < KB; ++q) {
        const int k = k0 + q * TPR;
        t[q] = (k < ntot) ? ((k < nrow) ? rp[(size_t)k * TILE] : cp[(size_t)k * n]) : 0.0;
      }
#pragma unroll
      for (int q = 0; q < KB; ++q) acc += t[q];
    }
  }
  return acc;
}
// stage 2: y(i) - (V p1)(i) - (W p2)(i), summed over the TPR lanes of the row (all lanes of the warp call)
__device__ __forceinline__ double row_combine(double acc, int sub, int jj, const double* s_p1, const double* s_p2,
                                              const double (&vv)[RPT], const double (&ww)[RPT]) {
#pragma unroll
  for (int u = 0; u < RPT; ++u) {
    const int t = sub + u * TPR;
    if (t < jj) acc -= vv[u] * s_p1[t] + ww[u] * s_p2[t];
  }
#pragma unroll
  for (int o = 1; o < TPR; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

static __global__ void __launch_bounds__(THREADS, 1) k_latrd_panel(Params P) {
  __shared__ double s_tmp[32], s_out[2], s_bc[2];
  __shared__ double s_p1[NB], s_p2[NB];      // W^T v, V^T v
  __shared__ double s_vr[NB], s_wr[NB];      // row j+1 of V and W
  __shared__ double s_vj[THREADS / 32][SW];  // per-warp copy of v over the strip's columns
  extern __shared__ double s_stage[];         // THREADS/32 x STAGE_DOUBLES: staged first strips

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned nblocks = gridDim.x;
  const int n = P.n;
  const long long lda = P.lda;
  double* A = P.A;
  double* W = P.W;
  const long long ldw = n;
  unsigned gen = 0;

  const int gthreads = nblocks * THREADS;
  const int gtid = blockIdx.x * THREADS + tid;
  // row phase: warps 0..ROW_WARPS-1 sweep rows (TPR lanes per row), the last warp recomputes row j+1
  const int sub = lane % TPR;
  const int slot = blockIdx.x * ROWS_PER_CTA + warp * (32 / TPR) + lane / TPR, nslots = nblocks * ROWS_PER_CTA;
  const bool row_warp = warp < ROW_WARPS;
  const int gwarp = blockIdx.x * (THREADS / 32) + warp;
  const long long nwarps = (long long)nblocks * (THREADS / 32);

  double* sbuf = s_stage + (size_t)warp * STAGE_DOUBLES;
  long long staged = -1;  // unit whose matrix entries sit in sbuf

  const bool prof = P.prof != nullptr && blockIdx.x == 0 && tid == 0;
  long long tc = prof ? clock64() : 0;
#define RS_PROF(k) if (prof) { const long long t_ = clock64(); P.prof[k] += t_ - tc; tc = t_; }

  // ---------------- first column of the panel: nothing to apply yet -----------------------------
  double ss = 0, alpha_part = 0;
  {
    const int j = P.j0;
    double* ab = P.abuf + (size_t)(j & 1) * n;
    for (int i = j + gtid; i < n; i += gthreads) {
      const double a = A[i + (long long)j * lda];
      ab[i] = a;
      if (i == j) P.d[j] = a;
      else if (i == j + 1) alpha_part = a;
      else ss += a * a;
    }
  }

  for (int jj = 0; jj < P.pw; ++jj) {
    const int j = P.j0 + jj;
    const double* ab = P.abuf + (size_t)(j & 1) * n;      // this column, updated
    double* abn = P.abuf + (size_t)((j + 1) & 1) * n;     // next column

    cta_reduce2(ss, alpha_part, s_tmp);
    // #1: ||x||^2 and alpha.  The row phase is over, DRAM idles until the product phase: the waiting warps stage
    // their first strip of it (column j, whose unit map does not depend on the reflector).
    {
      auto stage = [&]() {
        staged = stage_next_strip(P, j, jj, gwarp, nwarps, lane, sbuf);
      };
      grid_allreduce2(P.slots, nblocks, gen, ss, alpha_part, s_out, stage);
    }
    RS_PROF(1)

    // ---------------- reflector (every thread, identical arithmetic) ----------------------------
    double tau = 0, scale = 0;
    {
      const double alpha = alpha_part, xn2 = ss;
      double beta = alpha;
      if (xn2 > 0) {
        beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
      if (gtid == 0) { P.e[j] = beta; P.tau[j] = tau; }
    }
    // store v in the matrix (owners only; readers use ab[] * scale)
    for (int i = j + 1 + gtid; i < n; i += gthreads) A[i + (long long)j * lda] = (i == j + 1) ? 1.0 : ab[i] * scale;

    // ---------------- product phase: y = A22 v over the lower triangle; panel dot products -------
    const long long t3 = (P.prof != nullptr && tid == 0) ? clock64() : 0;
    const int Bmin = (j + 1) / TILE;
    const int mt = P.mB - Bmin;
    const int nch = (n - j - 1 + DOT_CHUNK - 1) / DOT_CHUNK;  // row chunks of the panel dot products
    UnitMap um;
    um.strips = um.first_of_row(mt);
    um.total = um.strips + 2LL * jj * nch;
    um.nw = nwarps < um.total ? nwarps : um.total;
    double yv = 0;  // this lane's share of v^T A22 v
    if (gwarp < um.nw) {
      long long u = um.start(gwarp);
      const long long uend = um.start(gwarp + 1);
      if (u < um.strips) {
        // locate the first strip: block row b (relative to Bmin), position k inside the row
        int b = (int)floor((sqrt(1.0 + 8.0 * (double)u / SPT) - 1.0) * 0.5);
        if (b < 0) b = 0;
        while (b > 0 && um.first_of_row(b) > u) --b;
        while (um.first_of_row(b + 1) <= u) ++b;
        int k = (int)(u - um.first_of_row(b));
        double rs0 = 0, rs1 = 0, vi0 = 0, vi1 = 0;
        bool newrow = true;
        const long long send = uend < um.strips ? uend : um.strips;
        for (; u < send; ++u) {
          const int I = Bmin + b, J = Bmin + k / SPT, pass = k % SPT;
          const int r0i = I * TILE + lane, r1i = r0i + 32;
          const bool ok0 = r0i < n, ok1 = r1i < n;
          const int cbase = J * TILE + pass * SW;  // first column of the strip
          // matrix loads first (the long latency), then the pieces of v
          double a0[SW], a1[SW], p[SW];
          const double* Ab = A + (long long)cbase * lda;
          if (staged == u) {
            // this strip was staged in shared memory while the grid was synchronising
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int c = 0; c < SW; ++c) {
              const bool okc = cbase + c < n;
              a0[c] = (ok0 && okc) ? sbuf[c * TILE + lane] : 0.0;
              a1[c] = (ok1 && okc) ? sbuf[c * TILE + lane + 32] : 0.0;
            }
            __syncwarp();
            staged = -1;
            if (prof) P.prof[12] += 1;
          } else {
#pragma unroll
            for (int c = 0; c < SW; ++c) {
              const bool okc = cbase + c < n;
              const double* col = Ab + (long long)c * lda;
              // streaming loads: do not evict the panel and the partial sums from L2
              a0[c] = (ok0 && okc) ? __ldcs(col + r0i) : 0.0;
              a1[c] = (ok1 && okc) ? __ldcs(col + r1i) : 0.0;
            }
          }
          if (newrow) {
            vi0 = (ok0 && r0i > j) ? (r0i == j + 1 ? 1.0 : ab[r0i] * scale) : 0.0;
            vi1 = (ok1 && r1i > j) ? (r1i == j + 1 ? 1.0 : ab[r1i] * scale) : 0.0;
            newrow = false;
          }
          double vcol;  // v at column cbase + (lane mod SW)
          {
            const int c = cbase + (lane & (SW - 1));
            vcol = (c < n && c > j) ? (c == j + 1 ? 1.0 : ab[c] * scale) : 0.0;
            __syncwarp();
            if (lane < SW) s_vj[warp][lane] = vcol;
            __syncwarp();
          }
          if (I != J) {
#pragma unroll
            for (int c = 0; c < SW; ++c) {
              const double vj = s_vj[warp][c];
              rs0 += a0[c] * vj;
              rs1 += a1[c] * vj;
              p[c] = a0[c] * vi0 + a1[c] * vi1;
            }
          } else {
#pragma unroll
            for (int c = 0; c < SW; ++c) {
              const int cl = pass * SW + c;  // column inside the 64 x 64 diagonal block
              const double vj = s_vj[warp][c];
              // local rows: lane and lane+32; lower triangle incl. diagonal feeds the row sums,
              // the strict lower triangle feeds the column sums (the upper triangle is never used)
              const double x0 = (lane >= cl) ? a0[c] : 0.0;
              const double x1 = (lane + 32 >= cl) ? a1[c] : 0.0;
              rs0 += x0 * vj;
              rs1 += x1 * vj;
              const double y0 = (lane > cl) ? a0[c] : 0.0;
              const double y1 = (lane + 32 > cl) ? a1[c] : 0.0;
              p[c] = y0 * vi0 + y1 * vi1;
            }
          }
          const double cs = butterfly(p, lane);
          if (col_writer(lane)) {
            const int cl = col_of_lane(lane);
            const int cg = cbase + cl;
            if (cg < n) P.colpart[(size_t)I * n + cg] = cs;
            yv += cs * s_vj[warp][cl];
          }
          // advance; flush the row sums at the end of a block row or of the run
          ++k;
          const bool endrow = (k == SPT * (b + 1));
          if (endrow || u + 1 == send) {
            double* rp = P.rowpart + (size_t)(gwarp + b) * TILE;
            rp[lane] = rs0;
            rp[lane + 32] = rs1;
            yv += rs0 * vi0 + rs1 * vi1;
            rs0 = rs1 = 0;
          }
          if (endrow) { ++b; k = 0; newrow = true; }
        }
      }
      // dot-product units: q < jj -> W(:,q)^T v ; q >= jj -> V(:,q-jj)^T v, over one chunk of rows
      for (; u < uend; ++u) {
        const int r = (int)(u - um.strips);
        const int q = r / nch, ch = r - q * nch;
        const double* colp = (q < jj) ? (W + (long long)q * ldw) : (A + (long long)(P.j0 + q - jj) * lda);
        const int ibeg = j + 1 + ch * DOT_CHUNK;
        const int iend = min(n, ibeg + DOT_CHUNK);
        double acc = 0;
#pragma unroll 8
        for (int i = ibeg + lane; i < iend; i += 32) {
          const double vi = (i == j + 1) ? 1.0 : ab[i] * scale;
          acc += colp[i] * vi;
        }
        acc = warp_sum_fixed(acc);
        if (lane == 0) P.pvec[(size_t)ch * (2 * NB) + q] = acc;
      }
    }
    if (staged >= 0) {  // a staged strip that was not consumed (cannot happen with identical unit maps)
      asm volatile("cp.async.wait_all;" ::: "memory");
      staged = -1;
    }
    RS_PROF(2)
    if (P.prof != nullptr && tid == 0) { P.prof[16 + blockIdx.x] += clock64() - t3; }
    {
      double z1 = 0;
      cta_reduce2(yv, z1, s_tmp);
      grid_allreduce2(P.slots, nblocks, gen, yv, z1, s_out);  // #2: v^T A22 v
    }
    RS_PROF(3)

    // ---------------- row phase: w, final W(:, jj), and the update of column j + 1 ---------------
    // One memory round trip: every warp first issues the loads of its row (the last warp: row j+1, whose
    // w every CTA needs, recomputed with the owner's arithmetic) and of the partial panel dot products.
    const bool more = jj + 1 < P.pw;
    double pa = 0;
    {
      const int q = tid >> 2, c4 = tid & 3;
      if (q < 2 * jj)
        for (int ch = c4; ch < nch; ch += 4) pa += P.pvec[(size_t)ch * (2 * NB) + q];
    }
    int i = row_warp ? j + 1 + slot : j + 1;
    bool live = i < n;
    double a_next = (row_warp && live && more && sub == 0) ? A[i + (long long)(j + 1) * lda] : 0.0;
    double vi = live ? ((i == j + 1) ? 1.0 : ab[i] * scale) : 0.0;
    double vv[RPT], ww[RPT];
    double acc = row_load(P, um, i, live, sub, jj, Bmin, vv, ww);
    {
      const int q = tid >> 2, c4 = tid & 3;
      pa += __shfl_xor_sync(0xffffffffu, pa, 1);
      pa += __shfl_xor_sync(0xffffffffu, pa, 2);
      if (c4 == 0 && q < 2 * jj) {
        if (q < jj) s_p1[q] = pa;
        else s_p2[q - jj] = pa;
      }
    }
    if (!row_warp && lane < TPR) {
#pragma unroll
      for (int u = 0; u < RPT; ++u) { s_vr[sub + u * TPR] = vv[u]; s_wr[sub + u * TPR] = ww[u]; }  // zero where t >= jj
    }
    __syncthreads();
    RS_PROF(5)
    acc = row_combine(acc, sub, jj, s_p1, s_p2, vv, ww);
    double upd = 0;
#pragma unroll
    for (int u = 0; u < RPT; ++u) upd += vv[u] * s_wr[sub + u * TPR] + ww[u] * s_vr[sub + u * TPR];
#pragma unroll
    for (int o = 1; o < TPR; o <<= 1) upd += __shfl_xor_sync(0xffffffffu, upd, o);
    if (!row_warp) {
      // w^T v = tau (v^T A22 v - 2 p1.p2)  -> alpha2;  w(j+1) before the correction
      double pp = 0;
      for (int t = lane; t < jj; t += 32) pp += s_p1[t] * s_p2[t];
      pp = warp_sum_fixed(pp);
      const double wv = tau * (yv - 2.0 * pp);
      if (lane == 0) { s_bc[0] = -0.5 * tau * wv; s_bc[1] = tau * acc; }
    }
    __syncthreads();
    RS_PROF(6)
    const double alpha2 = s_bc[0];
    const double wj1 = s_bc[1] + alpha2;  // final W(j+1, jj)   (v(j+1) = 1)
    ss = 0;
    alpha_part = 0;
    if (row_warp) {
      for (int base = j + 1;;) {  // trip count uniform within the warp (shuffles inside)
        if (sub == 0 && live) {
          const double wi = tau * acc + alpha2 * vi;  // final W(i, jj)
          W[i + (long long)jj * ldw] = wi;
          if (more) {
            // A(i, j+1) - sum_{t <= jj} V(i,t) W(j+1,t) + W(i,t) V(j+1,t)   with V(j+1, jj) = 1
            const double a = a_next - upd - (vi * wj1 + wi);
            abn[i] = a;
            if (i == j + 1) P.d[j + 1] = a;
            else if (i == j + 2) alpha_part = a;
            else ss += a * a;
          }
        }
        base += nslots;
        if (base >= n) break;
        i = base + slot;
        live = i < n;
        a_next = (live && more && sub == 0) ? A[i + (long long)(j + 1) * lda] : 0.0;
        vi = live ? ab[i] * scale : 0.0;
        acc = row_load(P, um, i, live, sub, jj, Bmin, vv, ww);
        acc = row_combine(acc, sub, jj, s_p1, s_p2, vv, ww);
        upd = 0;
#pragma unroll
        for (int u = 0; u < RPT; ++u) upd += vv[u] * s_wr[sub + u * TPR] + ww[u] * s_vr[sub + u * TPR];
#pragma unroll
        for (int o = 1; o < TPR; o <<= 1) upd += __shfl_xor_sync(0xffffffffu, upd, o);
      }
    }
    RS_PROF(4)
  }
#undef RS_PROF
}

static __global__ void k_restore_subdiag(double* A, long long lda, int n, const double* e, double* d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n - 1) A[(i + 1) + (long long)i * lda] = e[i];
  if (i == n - 1) d[n - 1] = A[(long long)(n - 1) * lda + (n - 1)];
}

constexpr size_t STAGE_BYTES = (size_t)(THREADS / 32) * STAGE_DOUBLES * sizeof(double);
inline size_t rowpart_doubles(int64_t n) {
  const int64_t mB = (n + TILE - 1) / TILE;
  return (size_t)(MAX_GRID * (THREADS / 32) + mB + 1) * TILE;
}
inline size_t workspace_doubles(int64_t n) {
  const int64_t mB = (n + TILE - 1) / TILE;
  return (size_t)n * NB + rowpart_doubles(n) + (size_t)mB * n + 2 * n + 2 * NB * (n / DOT_CHUNK + 1) + 4 * MAX_GRID + 8 + 64;
}

// A: n x n column-major, lower triangle referenced.  work: workspace_doubles(n) doubles (256-byte aligned).
// Returns cudaSuccess or the failing status; launches counted in *launches.
struct Hooks {  // optional instrumentation around every panel launch (phase timers of the library)
  void (*before)(void* user, double algorithmic_bytes) = nullptr;
  void (*after)(void* user) = nullptr;
  void* user = nullptr;
  // trailing update C(lower, nr x nr) -= V W^T + W V^T on the caller's own GEMM (the library passes its tcgen05
  // split-integer kernel, gemm_ozaki.cu); null: cublasDsyr2k (stand-alone probes)
  // r0: absolute index of the first row / column of the trailing block (the multi-GPU path shards the update by
  // absolute 128-row groups); syr2k_user: the caller's cookie
  void (*syr2k)(const double* V, long long ldv, const double* W, long long ldw, long long nr, int k, double* C,
                long long ldc, long long r0, void* user) = nullptr;
  void* syr2k_user = nullptr;
};
inline cudaError_t trailing_update(cublasHandle_t blas, const Hooks* hooks, const double* V, long long ldv, const double* W,
                                   long long ldw, long long nr, int k, double* C, long long ldc, long long r0) {
  if (hooks && hooks->syr2k) {
    hooks->syr2k(V, ldv, W, ldw, nr, k, C, ldc, r0, hooks->syr2k_user);
    return cudaSuccess;
  }
  const double mone = -1.0, one = 1.0;
  cublasStatus_t bs = cublasDsyr2k(blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, (int)nr, k, &mone, V, (int)ldv, W, (int)ldw, &one,
                                   C, (int)ldc);
  return bs == CUBLAS_STATUS_SUCCESS ? cudaSuccess : cudaErrorUnknown;
}
inline cudaError_t run(cublasHandle_t blas, cudaStream_t stream, double* A, int64_t n, int64_t lda, double* d, double* e,
                       double* tau, double* work, int64_t* launches, long long* prof = nullptr, const Hooks* hooks = nullptr) {
  if (n < 1) return cudaSuccess;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms < MAX_GRID ? sms : MAX_GRID;
  const int64_t mB = (n + TILE - 1) / TILE;
  Params P;
  P.A = A; P.lda = lda; P.n = (int)n; P.d = d; P.e = e; P.tau = tau; P.mB = (int)mB;
  double* w = work;
  P.W = w;        w += (size_t)n * NB;
  P.rowpart = w;  w += rowpart_doubles(n);
  P.colpart = w;  w += (size_t)mB * n;
  P.abuf = w;     w += 2 * n;
  P.pvec = w;     w += 2 * NB * (n / DOT_CHUNK + 1);
  if ((w - work) & 1) ++w;  // 16-byte alignment for the double2 loads of the partials
  P.slots = reinterpret_cast<unsigned long long*>(w);
  P.prof = prof;
  cublasSetStream(blas, stream);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t st = cudaFuncSetAttribute(k_latrd_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STAGE_BYTES);
    if (st != cudaSuccess) return st;
    attr_set = true;
  }
  for (int64_t j0 = 0; j0 < n - 1; j0 += NB) {
    const int pw = (int)((n - 1 - j0) < NB ? (n - 1 - j0) : NB);
    P.j0 = (int)j0; P.pw = pw;
    // generation numbers restart at 1 in every launch: clear the counter and the flag
    cudaError_t st = cudaMemsetAsync(P.slots, 0, sizeof(unsigned long long) * (4 * MAX_GRID + 8), stream);
    if (st != cudaSuccess) return st;
    void* args[] = {&P};
    if (hooks && hooks->before) {
      double bytes = 0;  // lower triangle of every trailing matrix the panel multiplies with
      for (int jj = 0; jj < pw; ++jj) { const double m = (double)(n - (j0 + jj) - 1); bytes += 4.0 * m * m; }
      hooks->before(hooks->user, bytes);
    }
    st = cudaLaunchCooperativeKernel((void*)k_latrd_panel, dim3(grid), dim3(THREADS), args, STAGE_BYTES, stream);
    if (hooks && hooks->after) hooks->after(hooks->user);
    if (st != cudaSuccess) return st;
    if (launches) ++*launches;
    const int64_t r0 = j0 + pw, nr = n - r0;  // trailing block
    if (nr > 0) {
      st = trailing_update(blas, hooks, A + r0 + j0 * lda, lda, P.W + r0, n, nr, pw, A + r0 + r0 * lda, lda, r0);
      if (st != cudaSuccess) return st;
    }
  }
  k_restore_subdiag<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(A, lda, (int)n, e, d);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace sytrd
}  // namespace rs

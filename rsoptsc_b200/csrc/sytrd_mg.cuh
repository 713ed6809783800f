// Fused multi-GPU Householder tridiagonalisation: the cooperative panel kernel of sytrd_kernel.cuh with the
// trailing-matrix product y = A22 v split across the GPUs of one NVLink/NVSwitch node and the partial products
// exchanged through peer-mapped memory INSIDE the kernel -- one compute + collective kernel per 64-column
// panel instead of two NCCL calls per column (a4: the symmetric eigensolver behind the SVT of
// R/SimilarityM.R:142-155 is 85 % of a 50,000-cell run, and this product is 70 % of it).
//
//   * The trailing matrix is owned by 128-ROW GROUPS dealt cyclically: block row I (64 rows, absolute index)
//     belongs to rank (I / 2) mod P for the whole factorisation.  Rank r streams only its block rows (1/P of the
//     4 n'^2 bytes per column, balanced to within one group for every column) with the strip / run / butterfly
//     scheme of the single-GPU kernel on its own unit map (three small lookup tables, MgTab).  Because ownership
//     never moves, a rank only ever reads its own rows of the trailing matrix -- so the rank-2NB trailing update
//     between panels is sharded as well: every rank updates its own 128-row groups and nothing else; the only
//     data every rank reads in full, the next panel's 64 columns, is all-gathered once per panel (NCCL, <= 20 MB).  The panel dot products W^T v, V^T v are dealt round-robin.
//   * After its product phase a rank reduces its OWN partial sums to one n'-vector (rows it touched, zeros
//     elsewhere) and writes that vector -- plus its share of W^T v, V^T v and v^T A v -- straight into the
//     exchange region of every rank with NVLink stores.  Every value travels as one 16-byte flag-with-data
//     packet {lo, flag, hi, flag} (flag = solve epoch << 16 | column + 1; each 8-byte half is delivered
//     atomically), and the row phase of the receiver simply spins on the packets it needs: no fence, no arrival
//     counter, no second NVLink round trip (a first version with __threadfence_system + red.release.sys
//     counters cost ~14 us per column at 7,954 rows, more than the product phase it saved on 2 GPUs).
//     Buffers alternate with the column parity: a rank can run at most one column ahead of a peer.
//   * The row phase (w, the final W column, the update of column j+1) runs replicated on bitwise identical
//     sums, so the panel, the reflectors and d/e/tau stay identical on all ranks.
// Per column: two local grid barriers (as before) + one cross-GPU exchange.
#pragma once
#include <vector>

#include "sytrd_kernel.cuh"

namespace rs {
namespace sytrd {

constexpr int MG_MAX_RANKS = 8;  // == TPR: one lane of a row group per rank in the row phase
constexpr int XCH_HDR = 32;      // doubles at the head of an exchange region (reserved)
static_assert(MG_MAX_RANKS <= TPR, "row phase reads one rank per lane of a row group");

// Ownership tables of one rank (device, built once per factorisation by run_mg):
//   ownI[l]   absolute block row of the rank's l-th block row, ascending (l < nl)
//   pref[l]   sum over t < l of SPT (ownI[t] + 1): strips of the rows before l counted from column block 0
//   lfirst[I] first l with ownI[l] >= I (nl when there is none), I = 0 .. mB
struct MgTab {
  const int* ownI;
  const long long* pref;
  const int* lfirst;
  int nl;
};
__host__ __device__ inline int mg_owner_of_block_row(int I, int nranks) { return (I >> 1) % nranks; }

struct MgParams {
  Params b;
  MgTab tab;
  int nranks, rank;
  double* xch[MG_MAX_RANKS];  // exchange region of every rank as mapped in this process
  long long xlen;             // packets per exchanged vector: y (n) | W^T v (NB) | V^T v (NB) | v^T A v | pad
  unsigned epoch;             // solve counter (1 .. 65535), upper half of every packet flag
};
inline long long mg_xlen(int64_t n) { return (n + 2 * NB + 2 + 31) / 32 * 32; }
// two parities x nranks sources x xlen packets of 16 bytes
inline size_t mg_region_doubles(int64_t n, int nranks) { return (size_t)XCH_HDR + 4 * (size_t)nranks * (size_t)mg_xlen(n); }

__device__ __forceinline__ uint4* mg_vec(double* region, int par, int src, int nranks, long long xlen) {
  return reinterpret_cast<uint4*>(region + XCH_HDR) + (long long)(par * nranks + src) * xlen;
}
// flag-with-data packet: both 8-byte halves carry the flag, so a reader that sees both flags has both halves
__device__ __forceinline__ void ll_store(uint4* slot, double v, unsigned flag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(slot), "r"((unsigned)b), "r"(flag),
               "r"((unsigned)(b >> 32)), "r"(flag)
               : "memory");
}
__device__ __forceinline__ double ll_load(const uint4* slot, unsigned flag) {
  unsigned lo, f0, hi, f1, spins = 0;
  unsigned long long t_start = 0;
  for (;;) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(slot) : "memory");
    if (f0 == flag && f1 == flag) break;
    if ((++spins & 0x3fffu) == 0) {  // a lost peer must fault, not hang the device: 20 s of wall clock
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t_start == 0) t_start = now;
      else if (now - t_start > 20000000000ull) __trap();
    }
  }
  return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

// Unit map of one rank for one column j: its block rows at or below Bmin = (j + 1) / 64 (local rows l0 .. nl-1 of
// MgTab), enumerated along block rows like the single-GPU map -- local row l holds SPT (ownI[l] - Bmin + 1) strips,
// the first at unit F(l) -- then its share of the panel dot-product units (global unit r = t * nranks + rank).
struct MgMap {
  int Bmin, mt, nch, l0, nl, bf;
  long long pref0, strips, ndot, total, nw;
  __device__ __forceinline__ long long F(const MgTab& t, int l) const {
    return __ldg(t.pref + l) - pref0 - (long long)SPT * Bmin * (l - l0);
  }
  __device__ __forceinline__ long long start(long long g) const { return (g * total + nw - 1) / nw; }
  __device__ __forceinline__ long long owner(long long u) const { return u * nw / total; }
};
__device__ __forceinline__ MgMap mg_map(const MgTab& tab, int nranks, int rank, int j, int jj, int n, int mB, long long nwarps) {
  MgMap m;
  m.Bmin = (j + 1) / TILE;
  m.mt = mB - m.Bmin;
  m.nch = (n - j - 1 + DOT_CHUNK - 1) / DOT_CHUNK;
  m.nl = tab.nl;
  m.l0 = m.Bmin <= mB ? __ldg(tab.lfirst + m.Bmin) : tab.nl;
  m.pref0 = __ldg(tab.pref + m.l0);
  m.bf = m.l0 < m.nl ? __ldg(tab.ownI + m.l0) - m.Bmin : 0;
  m.strips = m.F(tab, m.nl);
  const long long tdot = 2LL * jj * m.nch;
  m.ndot = tdot > rank ? (tdot - rank + nranks - 1) / nranks : 0;
  m.total = m.strips + m.ndot;
  m.nw = nwarps < m.total ? nwarps : m.total;
  if (m.nw < 1) m.nw = 1;  // empty share: start(0) = start(1) = 0
  return m;
}
// local row l (l0 <= l < nl) of strip unit u < strips: F(l) <= u < F(l + 1).  The rows of a rank advance by 2 nranks
// block rows per pair, so F is close to a quadratic in l - l0: solve it for the first guess, settle with the table.
__device__ __forceinline__ int mg_local_row(const MgTab& tab, const MgMap& m, int nranks, long long u) {
  const double b1 = (double)m.bf + 1.0;
  int l = m.l0 + (int)((sqrt(b1 * b1 + 2.0 * nranks * (double)u / SPT) - b1) / nranks);
  if (l > m.nl - 1) l = m.nl - 1;
  if (l < m.l0) l = m.l0;
  while (l > m.l0 && m.F(tab, l) > u) --l;
  while (l + 1 < m.nl && m.F(tab, l + 1) <= u) ++l;
  return l;
}

__device__ __forceinline__ long long mg_stage_next_strip(const Params& P, const MgTab& tab, const MgMap& um, int nranks,
                                                         long long gwarp, int lane, double* sbuf, int jnext) {
  if (!RS_SYTRD_STAGE || jnext + 1 >= P.n) return -1;
  if (um.total < 1 || gwarp >= um.nw) return -1;
  const long long u = um.start(gwarp);
  if (u >= um.strips || u >= um.start(gwarp + 1)) return -1;
  const int n = P.n;
  const int l = mg_local_row(tab, um, nranks, u);
  const int k = (int)(u - um.F(tab, l));
  const int I = __ldg(tab.ownI + l);
  const int cbase = um.Bmin * TILE + k * SW;
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

// This rank's share of y(i): its row partial sums (when row i lies in one of its block rows) and the column
// partial sums of its block rows at or below row i.  All TPR lanes of the row group call; lane `sub` adds every
// TPR-th partial.  Unreduced.
__device__ __forceinline__ double mg_row_partial(const Params& P, const MgTab& tab, const MgMap& um, int i, bool live, int sub) {
  double acc = 0;
  if (live) {
    const int n = P.n;
    const int Ii = i / TILE;
    const int lge = __ldg(tab.lfirst + max(Ii, um.Bmin));  // first own block row at or below row i
    int nrow = 0;
    const double* rp = P.rowpart;
    if (lge < um.nl && __ldg(tab.ownI + lge) == Ii) {      // row i lies in one of this rank's block rows
      const int b = Ii - um.Bmin;
      const long long s0 = um.F(tab, lge);
      const int g0 = (int)um.owner(s0), g1 = (int)um.owner(s0 + (long long)SPT * (b + 1) - 1);
      nrow = g1 - g0 + 1;
      rp = P.rowpart + (size_t)(g0 + lge - um.l0) * TILE + (i - Ii * TILE);
    }
    const int ncol = max(0, um.nl - lge);                  // column partials are stored by LOCAL block row
    const double* cp = P.colpart + ((long long)lge * n + i);
    const int ntot = nrow + ncol;
    for (int k0 = sub; k0 < ntot; k0 += TPR * KB) {
      double t[KB];
#pragma unroll
      for (int q = 0; q < KB; ++q) {
        const int k = k0 + q * TPR;
        t[q] = (k < ntot) ? ((k < nrow) ? rp[(size_t)k * TILE] : cp[(size_t)(k - nrow) * n]) : 0.0;
      }
#pragma unroll
      for (int q = 0; q < KB; ++q) acc += t[q];
    }
  }
  return acc;
}

static __global__ void __launch_bounds__(THREADS, 1) k_latrd_panel_mg(MgParams Q) {
  __shared__ double s_tmp[32], s_out[2], s_bc[2], s_yv;
  __shared__ double s_p1[NB], s_p2[NB];      // W^T v, V^T v
  __shared__ double s_vr[NB], s_wr[NB];      // row j+1 of V and W
  __shared__ double s_vj[THREADS / 32][SW];  // per-warp copy of v over the strip's columns
  extern __shared__ double s_stage[];         // THREADS/32 x STAGE_DOUBLES: staged first strips

  const Params& P = Q.b;
  const MgTab& tab = Q.tab;
  const int NR = Q.nranks, rank = Q.rank;
  const long long xlen = Q.xlen;
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
  const int sub = lane % TPR;
  const int slot = blockIdx.x * ROWS_PER_CTA + warp * (32 / TPR) + lane / TPR, nslots = nblocks * ROWS_PER_CTA;
  const bool row_warp = warp < ROW_WARPS;
  const int gwarp = blockIdx.x * (THREADS / 32) + warp;
  const long long nwarps = (long long)nblocks * (THREADS / 32);
  double* xloc = Q.xch[rank];

  double* sbuf = s_stage + (size_t)warp * STAGE_DOUBLES;
  long long staged = -1;

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
    const int par = j & 1;
    const unsigned flag = (Q.epoch << 16) | (unsigned)((j + 1) & 0xffff);
    const double* ab = P.abuf + (size_t)(j & 1) * n;
    double* abn = P.abuf + (size_t)((j + 1) & 1) * n;
    const MgMap um = mg_map(tab, NR, rank, j, jj, n, P.mB, nwarps);

    cta_reduce2(ss, alpha_part, s_tmp);
    {
      auto stage = [&]() { staged = mg_stage_next_strip(P, tab, um, NR, gwarp, lane, sbuf, j); };
      grid_allreduce2(P.slots, nblocks, gen, ss, alpha_part, s_out, stage);  // #1 (local: the data is replicated)
    }

    // ---------------- reflector (every thread of every rank, identical arithmetic) ----------------
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
    for (int i = j + 1 + gtid; i < n; i += gthreads) A[i + (long long)j * lda] = (i == j + 1) ? 1.0 : ab[i] * scale;

    // ---------------- product phase over this rank's block rows; its share of the panel dot products ----
    const int Bmin = um.Bmin;
    const int nch = um.nch;
    double yv = 0;
    if (um.total > 0 && gwarp < um.nw) {
      long long u = um.start(gwarp);
      const long long uend = um.start(gwarp + 1);
      if (u < um.strips) {
        int l = mg_local_row(tab, um, NR, u);
        int b = __ldg(tab.ownI + l) - Bmin;
        int k = (int)(u - um.F(tab, l));
        double rs0 = 0, rs1 = 0, vi0 = 0, vi1 = 0;
        bool newrow = true;
        const long long send = uend < um.strips ? uend : um.strips;
        for (; u < send; ++u) {
          const int I = Bmin + b, J = Bmin + k / SPT, pass = k % SPT;
          const int r0i = I * TILE + lane, r1i = r0i + 32;
          const bool ok0 = r0i < n, ok1 = r1i < n;
          const int cbase = J * TILE + pass * SW;
          double a0[SW], a1[SW], p[SW];
          const double* Ab = A + (long long)cbase * lda;
          if (staged == u) {
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
          } else {
#pragma unroll
            for (int c = 0; c < SW; ++c) {
              const bool okc = cbase + c < n;
              const double* col = Ab + (long long)c * lda;
              a0[c] = (ok0 && okc) ? __ldcs(col + r0i) : 0.0;
              a1[c] = (ok1 && okc) ? __ldcs(col + r1i) : 0.0;
            }
          }
          if (newrow) {
            vi0 = (ok0 && r0i > j) ? (r0i == j + 1 ? 1.0 : ab[r0i] * scale) : 0.0;
            vi1 = (ok1 && r1i > j) ? (r1i == j + 1 ? 1.0 : ab[r1i] * scale) : 0.0;
            newrow = false;
          }
          {
            const int c = cbase + (lane & (SW - 1));
            const double vcol = (c < n && c > j) ? (c == j + 1 ? 1.0 : ab[c] * scale) : 0.0;
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
              const int cl = pass * SW + c;
              const double vj = s_vj[warp][c];
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
            if (cg < n) P.colpart[(size_t)l * n + cg] = cs;
            yv += cs * s_vj[warp][cl];
          }
          ++k;
          const bool endrow = (k == SPT * (b + 1));
          if (endrow || u + 1 == send) {
            double* rp = P.rowpart + (size_t)(gwarp + l - um.l0) * TILE;
            rp[lane] = rs0;
            rp[lane + 32] = rs1;
            yv += rs0 * vi0 + rs1 * vi1;
            rs0 = rs1 = 0;
          }
          if (endrow) {
            ++l;
            if (l < um.nl) b = __ldg(tab.ownI + l) - Bmin;
            k = 0;
            newrow = true;
          }
        }
      }
      // dot-product units of this rank: global unit r = t * NR + rank
      for (; u < uend; ++u) {
        const long long r = (u - um.strips) * NR + rank;
        const int q = (int)(r / nch), ch = (int)(r - (long long)q * nch);
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
    if (staged >= 0) {
      asm volatile("cp.async.wait_all;" ::: "memory");
      staged = -1;
    }
    {
      double z1 = 0;
      cta_reduce2(yv, z1, s_tmp);
      grid_allreduce2(P.slots, nblocks, gen, yv, z1, s_out);  // #2 (local): this rank's v^T A22 v, partials visible
    }

    // ---------------- local reduction + exchange: this rank's y, W^T v, V^T v, v^T A v to every rank -------
    if (blockIdx.x == 0) {
      const int q = tid >> 2, c4 = tid & 3;
      double pa = 0;
      if (q < 2 * jj)
        for (int ch = c4; ch < nch; ch += 4)
          if (((long long)q * nch + ch) % NR == rank) pa += P.pvec[(size_t)ch * (2 * NB) + q];
      pa += __shfl_xor_sync(0xffffffffu, pa, 1);
      pa += __shfl_xor_sync(0xffffffffu, pa, 2);
      if (c4 == 0 && q < 2 * jj)
        for (int r = 0; r < NR; ++r) ll_store(mg_vec(Q.xch[r], par, rank, NR, xlen) + n + q, pa, flag);
      if (tid == THREADS - 1)
        for (int r = 0; r < NR; ++r) ll_store(mg_vec(Q.xch[r], par, rank, NR, xlen) + n + 2 * NB, yv, flag);
    }
    if (row_warp) {
      for (int base = j + 1; base < n; base += nslots) {  // trip count uniform within the warp (shuffles inside)
        const int i = base + slot;
        const bool live = i < n;
        double acc = mg_row_partial(P, tab, um, i, live, sub);
#pragma unroll
        for (int o = 1; o < TPR; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (live && sub < NR) ll_store(mg_vec(Q.xch[sub], par, rank, NR, xlen) + i, acc, flag);  // lane `sub` serves rank `sub`
      }
    }

    // ---------------- row phase: w, final W(:, jj), and the update of column j + 1 (replicated) ----------
    const bool more = jj + 1 < P.pw;
    if (tid < 2 * jj) {
      double s = 0;
      for (int r = 0; r < NR; ++r) s += ll_load(mg_vec(xloc, par, r, NR, xlen) + n + tid, flag);
      if (tid < jj) s_p1[tid] = s; else s_p2[tid - jj] = s;
    }
    if (tid == THREADS - 1) {
      double s = 0;
      for (int r = 0; r < NR; ++r) s += ll_load(mg_vec(xloc, par, r, NR, xlen) + n + 2 * NB, flag);
      s_yv = s;
    }
    int i = row_warp ? j + 1 + slot : j + 1;
    bool live = i < n;
    double a_next = (row_warp && live && more && sub == 0) ? A[i + (long long)(j + 1) * lda] : 0.0;
    double vi = live ? ((i == j + 1) ? 1.0 : ab[i] * scale) : 0.0;
    double vv[RPT], ww[RPT];
    auto load_row = [&](int row, bool on_row) -> double {
#pragma unroll
      for (int u = 0; u < RPT; ++u) {
        const int t = sub + u * TPR;
        const bool on = on_row && t < jj;
        vv[u] = on ? A[row + (long long)(P.j0 + t) * lda] : 0.0;
        ww[u] = on ? W[row + (long long)t * n] : 0.0;
      }
      return (on_row && sub < NR) ? ll_load(mg_vec(xloc, par, sub, NR, xlen) + row, flag) : 0.0;  // one rank's share per lane
    };
    double acc = load_row(i, live);
    if (!row_warp && lane < TPR) {
#pragma unroll
      for (int u = 0; u < RPT; ++u) { s_vr[sub + u * TPR] = vv[u]; s_wr[sub + u * TPR] = ww[u]; }
    }
    __syncthreads();
    acc = row_combine(acc, sub, jj, s_p1, s_p2, vv, ww);
    double upd = 0;
#pragma unroll
    for (int u = 0; u < RPT; ++u) upd += vv[u] * s_wr[sub + u * TPR] + ww[u] * s_vr[sub + u * TPR];
#pragma unroll
    for (int o = 1; o < TPR; o <<= 1) upd += __shfl_xor_sync(0xffffffffu, upd, o);
    if (!row_warp) {
      double pp = 0;
      for (int t = lane; t < jj; t += 32) pp += s_p1[t] * s_p2[t];
      pp = warp_sum_fixed(pp);
      const double wv = tau * (s_yv - 2.0 * pp);
      if (lane == 0) { s_bc[0] = -0.5 * tau * wv; s_bc[1] = tau * acc; }
    }
    __syncthreads();
    const double alpha2 = s_bc[0];
    const double wj1 = s_bc[1] + alpha2;
    ss = 0;
    alpha_part = 0;
    if (row_warp) {
      for (int base = j + 1;;) {
        if (sub == 0 && live) {
          const double wi = tau * acc + alpha2 * vi;
          W[i + (long long)jj * ldw] = wi;
          if (more) {
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
        acc = load_row(i, live);
        acc = row_combine(acc, sub, jj, s_p1, s_p2, vv, ww);
        upd = 0;
#pragma unroll
        for (int u = 0; u < RPT; ++u) upd += vv[u] * s_wr[sub + u * TPR] + ww[u] * s_vr[sub + u * TPR];
#pragma unroll
        for (int o = 1; o < TPR; o <<= 1) upd += __shfl_xor_sync(0xffffffffu, upd, o);
      }
    }
  }
}

// One rank's side of the fused factorisation.  A: n x n column-major, replicated, lower triangle referenced.
// xch[q]: exchange region of rank q as mapped here, mg_region_doubles(n, nranks) doubles.  epoch: 1 .. 65535, the same
// on every rank and different from the epoch of the previous factorisation that used the regions (packet flags);
// the regions are zeroed (and a barrier passed) whenever the epoch counter wraps.  n <= 65,535.  grid <= SM count CTAs.
// host side of MgTab: `buf` holds mg_tab_bytes(n) bytes of device memory (8-byte aligned)
inline size_t mg_tab_bytes(int64_t n) {
  const size_t mB = (size_t)((n + TILE - 1) / TILE);
  return (mB + 2) * sizeof(long long) + (2 * mB + 4) * sizeof(int);
}
// the three tables of one rank on the host (shared with the CPU test of the map, rsoptsc_host_mg_tables)
inline void mg_tables_host(int64_t n, int nranks, int rank, std::vector<int>& ownI, std::vector<long long>& pref,
                           std::vector<int>& lfirst) {
  const int mB = (int)((n + TILE - 1) / TILE);
  ownI.clear();
  lfirst.assign(mB + 1, 0);
  for (int I = 0; I < mB; ++I)
    if (mg_owner_of_block_row(I, nranks) == rank) ownI.push_back(I);
  const int nl = (int)ownI.size();
  pref.assign(nl + 1, 0);
  for (int l = 0; l < nl; ++l) pref[l + 1] = pref[l] + (long long)SPT * (ownI[l] + 1);
  for (int I = 0, l = 0; I <= mB; ++I) {
    while (l < nl && ownI[l] < I) ++l;
    lfirst[I] = l;
  }
}
inline cudaError_t mg_build_tab(int64_t n, int nranks, int rank, void* buf, cudaStream_t stream, MgTab* tab) {
  const int mB = (int)((n + TILE - 1) / TILE);
  std::vector<int> ownI, lfirst;
  std::vector<long long> pref;
  mg_tables_host(n, nranks, rank, ownI, pref, lfirst);
  const int nl = (int)ownI.size();
  char* p = static_cast<char*>(buf);
  long long* d_pref = reinterpret_cast<long long*>(p);
  int* d_own = reinterpret_cast<int*>(p + (size_t)(mB + 2) * sizeof(long long));
  int* d_lfirst = d_own + mB + 1;
  cudaError_t st = cudaMemcpyAsync(d_pref, pref.data(), sizeof(long long) * (nl + 1), cudaMemcpyHostToDevice, stream);
  if (st == cudaSuccess && nl > 0) st = cudaMemcpyAsync(d_own, ownI.data(), sizeof(int) * nl, cudaMemcpyHostToDevice, stream);
  if (st == cudaSuccess) st = cudaMemcpyAsync(d_lfirst, lfirst.data(), sizeof(int) * (mB + 1), cudaMemcpyHostToDevice, stream);
  if (st == cudaSuccess) st = cudaStreamSynchronize(stream);  // the host vectors die with this frame
  tab->ownI = d_own; tab->pref = d_pref; tab->lfirst = d_lfirst; tab->nl = nl;
  return st;
}

// `tabbuf`: mg_tab_bytes(n) bytes of device scratch for the ownership tables.  hooks->syr2k (when set) receives
// the absolute first row of the trailing block and must leave this rank's 128-row groups of it and the whole of the
// next panel's 64 columns current (eig.cu: sharded tcgen05 update + panel gather); the cuBLAS fall-back updates everything.
inline cudaError_t run_mg(cublasHandle_t blas, cudaStream_t stream, double* A, int64_t n, int64_t lda, double* d, double* e,
                          double* tau, double* work, void* tabbuf, int nranks, int rank, double* const* xch, unsigned epoch,
                          int grid, int64_t* launches, const Hooks* hooks = nullptr) {
  if (n < 1) return cudaSuccess;
  const int64_t mB = (n + TILE - 1) / TILE;
  MgParams Q;
  {
    cudaError_t st = mg_build_tab(n, nranks, rank, tabbuf, stream, &Q.tab);
    if (st != cudaSuccess) return st;
  }
  Params& P = Q.b;
  P.A = A; P.lda = lda; P.n = (int)n; P.d = d; P.e = e; P.tau = tau; P.mB = (int)mB;
  double* w = work;
  P.W = w;        w += (size_t)n * NB;
  P.rowpart = w;  w += rowpart_doubles(n);
  P.colpart = w;  w += (size_t)mB * n;
  P.abuf = w;     w += 2 * n;
  P.pvec = w;     w += 2 * NB * (n / DOT_CHUNK + 1);
  if ((w - work) & 1) ++w;
  P.slots = reinterpret_cast<unsigned long long*>(w);
  P.prof = nullptr;
  Q.nranks = nranks; Q.rank = rank; Q.xlen = mg_xlen(n); Q.epoch = epoch;
  for (int q = 0; q < MG_MAX_RANKS; ++q) Q.xch[q] = q < nranks ? xch[q] : nullptr;
  cublasSetStream(blas, stream);
  {
    cudaError_t st = cudaFuncSetAttribute(k_latrd_panel_mg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STAGE_BYTES);
    if (st != cudaSuccess) return st;
  }
  for (int64_t j0 = 0; j0 < n - 1; j0 += NB) {
    const int pw = (int)((n - 1 - j0) < NB ? (n - 1 - j0) : NB);
    P.j0 = (int)j0; P.pw = pw;
    cudaError_t st = cudaMemsetAsync(P.slots, 0, sizeof(unsigned long long) * (4 * MAX_GRID + 8), stream);
    if (st != cudaSuccess) return st;
    void* args[] = {&Q};
    if (hooks && hooks->before) {
      double bytes = 0;  // this rank's share of the lower triangles the panel multiplies with
      for (int jj = 0; jj < pw; ++jj) { const double m = (double)(n - (j0 + jj) - 1); bytes += 4.0 * m * m / nranks; }
      hooks->before(hooks->user, bytes);
    }
    st = cudaLaunchCooperativeKernel((void*)k_latrd_panel_mg, dim3(grid), dim3(THREADS), args, STAGE_BYTES, stream);
    if (hooks && hooks->after) hooks->after(hooks->user);
    if (st != cudaSuccess) return st;
    if (launches) ++*launches;
    const int64_t r0 = j0 + pw, nr = n - r0;
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

// Divide-and-conquer eigensolver for a symmetric tridiagonal matrix (Cuppen's method with the
// Gu/Eisenstat deflation and Loewner re-derivation of the rank-one weights) -- the second stage of
// the symmetric eigensolver behind the SVT (a4, R/SimilarityM.R:142-155), after sytrd_kernel.cuh.
//
// This header holds the numerics shared by the CUDA driver (stedc.cu) and the CPU test harness
// (tests/stedc_host_check.cpp): the secular-equation root finder, the Loewner weights and the host
// side merge planning (sorting, deflation, Givens rotations).  Everything is written from the
// published algorithm (Cuppen 1981; Gu & Eisenstat 1995; Li 1994 "middle way"); no LAPACK source is
// involved.
//
// One merge: T = diag(T1, T2) + rho u u^T with T1 = Q1 D1 Q1^T, T2 = Q2 D2 Q2^T gives
//     T = Q (D + rho z z^T) Q^T,  Q = diag(Q1, Q2),  z = Q^T u,
// and the eigenpairs of D + rho z z^T come from the secular equation
//     w(lambda) = 1/rho + sum_i z_i^2 / (d_i - lambda) = 0 .
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#ifdef __CUDACC__
#define RS_HD __host__ __device__
#else
#define RS_HD
#endif

namespace rs {
namespace stedc {

constexpr double EPS = 2.220446049250313e-16;

// Root j of the secular equation for k poles dl[0] < dl[1] < ... < dl[k-1] with weights z (all
// non-zero) and rho > 0.  Root j lies in (dl[j], dl[j+1]) (the last one in (dl[k-1], dl[k-1] + rho |z|^2)).
// The root is returned as lambda = dl[org] + tau with org the nearer pole, so that every difference
// dl[i] - lambda = (dl[i] - dl[org]) - tau keeps full relative accuracy.
// `lanes`/`lane`: the sums over poles can be split over a group of cooperating threads; `reduce` adds a
// value over the group (identity for one thread).
template <class Reduce>
RS_HD inline void secular_root(int k, int j, const double* dl, const double* z, double rho, int lane, int lanes,
                               Reduce reduce, int* org_out, double* tau_out) {
  const double rhoinv = 1.0 / rho;
  if (k == 1) {
    *org_out = 0;
    *tau_out = rho * z[0] * z[0];
    return;
  }
  const bool last = (j == k - 1);
  int org;
  double lo, hi;  // bracket for tau: w(lo) < 0 < w(hi) (open at the pole side)
  if (!last) {
    const double gap = dl[j + 1] - dl[j];
    const double mid = 0.5 * gap;  // relative to dl[j]
    double s = 0;
    for (int i = lane; i < k; i += lanes) s += z[i] * z[i] / ((dl[i] - dl[j]) - mid);
    s = reduce(s);
    const double wmid = rhoinv + s;
    if (wmid > 0) { org = j; lo = 0.0; hi = mid; }
    else { org = j + 1; lo = -mid; hi = 0.0; }
  } else {
    double zz = 0;
    for (int i = lane; i < k; i += lanes) zz += z[i] * z[i];
    zz = reduce(zz);
    org = k - 1;
    lo = 0.0;
    hi = rho * zz;
    // w(hi) >= 0 holds analytically; nudge for rounding
    hi *= (1.0 + 8 * EPS);
  }
  const double dorg = dl[org];
  double tau = 0.5 * (lo + hi);
  // pole(s) used by the rational model, relative to dl[org]
  const double p1 = dl[j] - dorg;                          // left pole (<= 0)
  const double p2 = last ? 0.0 : (dl[j + 1] - dorg);       // right pole (>= 0), unused for the last root
  for (int it = 0; it < 100; ++it) {
    double psi = 0, dpsi = 0, phi = 0, dphi = 0;
    for (int i = lane; i < k; i += lanes) {
      const double del = (dl[i] - dorg) - tau;
      const double t = z[i] / del;
      const double zt = z[i] * t;  // z^2 / del
      if (i <= j) { psi += zt; dpsi += t * t; }
      else { phi += zt; dphi += t * t; }
    }
    psi = reduce(psi); dpsi = reduce(dpsi); phi = reduce(phi); dphi = reduce(dphi);
    const double w = rhoinv + psi + phi;
    const double err = 8.0 * (phi - psi) + 2.0 * rhoinv + fabs(tau) * (dpsi + dphi);
    if (fabs(w) <= EPS * err) break;
    if (w < 0) lo = tau; else hi = tau;
    if (!(hi - lo > 2.0 * EPS * fmax(fabs(lo), fabs(hi)))) break;  // bracket exhausted
    const double D1 = p1 - tau;  // < 0
    double eta;
    if (!last) {
      const double D2 = p2 - tau;  // > 0
      // middle way: psi ~ s1 + a1 / (p1 - t), phi ~ s2 + a2 / (p2 - t), matched in value and slope at tau
      const double a1 = dpsi * D1 * D1, a2 = dphi * D2 * D2;
      const double c = rhoinv + (psi - dpsi * D1) + (phi - dphi * D2);
      // c eta^2 - B eta + C = 0 with the wanted root in (D1, D2)
      const double B = c * (D1 + D2) + a1 + a2;
      const double C = D1 * D2 * w;
      if (c == 0.0) {
        eta = (B != 0.0) ? C / B : 0.0;
      } else {
        double disc = B * B - 4.0 * c * C;
        if (disc < 0) disc = 0;
        const double sq = sqrt(disc);
        const double q = 0.5 * (B + (B >= 0 ? sq : -sq));
        const double e1 = (q != 0.0) ? C / q : 0.0;
        const double e2 = q / c;
        const bool in1 = e1 > D1 && e1 < D2, in2 = e2 > D1 && e2 < D2;
        if (in1 && in2) eta = fabs(e1) <= fabs(e2) ? e1 : e2;
        else if (in1) eta = e1;
        else if (in2) eta = e2;
        else eta = 0.0;
      }
    } else {
      // single pole on the left: w ~ c + a1 / (p1 - t)
      const double a1 = dpsi * D1 * D1;
      const double c = rhoinv + (psi - dpsi * D1) + phi;
      eta = (c > 0) ? D1 + a1 / c : 0.0;
    }
    double tnew = tau + eta;
    if (!(eta != 0.0) || !(tnew > lo) || !(tnew < hi)) tnew = 0.5 * (lo + hi);  // safeguard: bisection
    if (tnew == tau) break;
    tau = tnew;
  }
  *org_out = org;
  *tau_out = tau;
}

struct NoReduce {
  RS_HD double operator()(double v) const { return v; }
};

// dl[i] - lambda_j
RS_HD inline double delta(const double* dl, int i, int org_j, double tau_j) { return (dl[i] - dl[org_j]) - tau_j; }

// Loewner weight: |zhat_i|^2 = (lambda_i - dl_i) prod_{j != i} (lambda_j - dl_i) / (dl_j - dl_i)   (up to 1/rho,
// which the normalisation of the eigenvectors removes).  Partial product over j = lane, lane+lanes, ...
RS_HD inline double loewner_partial(int k, int i, const double* dl, const int* org, const double* tau, int lane, int lanes) {
  double p = 1.0;
  for (int j = lane; j < k; j += lanes) {
    const double num = -delta(dl, i, org[j], tau[j]);  // lambda_j - dl_i
    p *= (j == i) ? num : num / (dl[j] - dl[i]);
  }
  return p;
}

// ------------------------------------------------------------------------------------------------
// Host side: plan of one merge
// ------------------------------------------------------------------------------------------------
struct Rotation {
  int a, b;     // columns (a is deflated afterwards, b survives):  x' = c x + s y ; y' = c y - s x
  double c, s;
};

struct MergePlan {
  int nm = 0;                  // n1 + n2
  double rho = 0;              // of the normalised problem (|z| = 1)
  int k = 0;                   // poles that take part in the secular equation
  std::vector<int> idx_nd;     // column (0..nm) of pole i, poles ascending
  std::vector<double> dl, zz;  // poles and weights
  std::vector<int> idx_def;    // deflated columns ...
  std::vector<double> d_def;   // ... and their eigenvalues
  std::vector<Rotation> rots;  // column rotations, in the order they were applied
};

// d: the nm eigenvalues of the two children (any order); z: [last row of Q1 ; sign(beta) * first row of Q2];
// beta: the off-diagonal element that couples the children.
inline void plan_merge(const double* d_in, const double* z_in, double beta, int nm, MergePlan& P) {
  P.nm = nm;
  P.k = 0;
  P.idx_nd.clear(); P.dl.clear(); P.zz.clear(); P.idx_def.clear(); P.d_def.clear(); P.rots.clear();
  std::vector<double> d(d_in, d_in + nm), z(nm);
  const double inv = 1.0 / std::sqrt(2.0);
  double dmax = 0, zmax = 0;
  for (int i = 0; i < nm; ++i) {
    z[i] = z_in[i] * inv;
    dmax = std::max(dmax, std::fabs(d[i]));
    zmax = std::max(zmax, std::fabs(z[i]));
  }
  P.rho = 2.0 * std::fabs(beta);
  const double tol = 8.0 * EPS * std::max(dmax, zmax);
  std::vector<int> order(nm);
  for (int i = 0; i < nm; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return d[a] < d[b]; });
  if (P.rho * zmax <= tol) {  // nothing couples: plain concatenation
    for (int i : order) { P.idx_def.push_back(i); P.d_def.push_back(d[i]); }
    return;
  }
  int pj = -1;
  for (int q = 0; q < nm; ++q) {
    const int j = order[q];
    if (P.rho * std::fabs(z[j]) <= tol) {  // negligible weight: (d_j, q_j) is already an eigenpair
      P.idx_def.push_back(j);
      P.d_def.push_back(d[j]);
      continue;
    }
    if (pj < 0) { pj = j; continue; }
    // close poles: rotate the pair so that one weight vanishes, deflate if the coupling is negligible
    double s = z[pj], c = z[j];
    const double tau = std::hypot(c, s);
    const double t = d[j] - d[pj];
    c /= tau;
    s = -s / tau;
    if (std::fabs(t * c * s) <= tol) {
      z[j] = tau;
      z[pj] = 0.0;
      P.rots.push_back({pj, j, c, s});
      const double tt = d[pj] * c * c + d[j] * s * s;
      d[j] = d[pj] * s * s + d[j] * c * c;
      d[pj] = tt;
      P.idx_def.push_back(pj);
      P.d_def.push_back(d[pj]);
      pj = j;
    } else {
      P.idx_nd.push_back(pj); P.dl.push_back(d[pj]); P.zz.push_back(z[pj]);
      pj = j;
    }
  }
  if (pj >= 0) { P.idx_nd.push_back(pj); P.dl.push_back(d[pj]); P.zz.push_back(z[pj]); }
  P.k = (int)P.idx_nd.size();
}

// After the roots are known: output column of every pole / deflated value in ascending order of the new
// eigenvalues.  lam[j] = dl[org[j]] + tau[j].
inline void order_outputs(const MergePlan& P, const std::vector<double>& lam, std::vector<int>& pos_nd,
                          std::vector<int>& pos_def, std::vector<double>& d_out) {
  const int k = P.k, nd = (int)P.idx_def.size(), nm = k + nd;
  std::vector<std::pair<double, int>> v(nm);
  for (int j = 0; j < k; ++j) v[j] = {lam[j], j};
  for (int t = 0; t < nd; ++t) v[k + t] = {P.d_def[t], k + t};
  std::stable_sort(v.begin(), v.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first < b.first; });
  pos_nd.assign(k, 0);
  pos_def.assign(nd, 0);
  d_out.assign(nm, 0.0);
  for (int p = 0; p < nm; ++p) {
    d_out[p] = v[p].first;
    if (v[p].second < k) pos_nd[v[p].second] = p; else pos_def[v[p].second - k] = p;
  }
}

// Leaves of the recursion: [off, off + size) blocks of at most `leaf` rows, sizes as equal as possible.
inline void make_leaves(int n, int leaf, std::vector<int>& off) {
  const int L = std::max(1, (n + leaf - 1) / leaf);
  off.assign(L + 1, 0);
  for (int i = 0; i <= L; ++i) off[i] = (int)((long long)n * i / L);
}

}  // namespace stedc
}  // namespace rs

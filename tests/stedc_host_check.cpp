// CPU harness for the numerics of the tridiagonal divide-and-conquer eigensolver
// (rsoptsc_b200/csrc/stedc_core.h): runs the same merge plan / secular solver / Loewner weights the
// CUDA driver (stedc.cu) uses, with host loops in place of kernels, so the algorithm can be checked
// against numpy.linalg.eigh without a GPU (tests/test_stedc_host.py).
//   g++ -O2 -std=c++17 -shared -fPIC tests/stedc_host_check.cpp -o <tmp>/libstedc_host_check.so
#include <cstring>

#include "../rsoptsc_b200/csrc/stedc_core.h"

using namespace rs::stedc;

// cyclic Jacobi on a small dense symmetric matrix (leaf solver of the harness only)
static void jacobi_eig(std::vector<double>& A, int n, std::vector<double>& V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[i + (size_t)i * n] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) off += A[p + (size_t)q * n] * A[p + (size_t)q * n];
    if (off < 1e-300) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p + (size_t)q * n];
        if (apq == 0.0) continue;
        const double theta = (A[q + (size_t)q * n] - A[p + (size_t)p * n]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[k + (size_t)p * n], akq = A[k + (size_t)q * n];
          A[k + (size_t)p * n] = c * akp - s * akq;
          A[k + (size_t)q * n] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[p + (size_t)k * n], aqk = A[q + (size_t)k * n];
          A[p + (size_t)k * n] = c * apk - s * aqk;
          A[q + (size_t)k * n] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[k + (size_t)p * n], vkq = V[k + (size_t)q * n];
          V[k + (size_t)p * n] = c * vkp - s * vkq;
          V[k + (size_t)q * n] = s * vkp + c * vkq;
        }
      }
  }
}

struct Block {
  int off, size;
};

extern "C" int stedc_host(int n, const double* d_in, const double* e_in, int leaf, double* lam_out, double* Q_out,
                          int* stats /* [0] merges, [1] deflated total, [2] rotations total */) {
  std::vector<double> d(d_in, d_in + n), e(e_in, e_in + (n > 1 ? n - 1 : 0));
  std::vector<int> loff;
  make_leaves(n, leaf, loff);
  const int L = (int)loff.size() - 1;
  // rank-one tears at every leaf boundary
  for (int b = 1; b < L; ++b) {
    const int i = loff[b];  // boundary between rows i-1 and i
    d[i - 1] -= std::fabs(e[i - 1]);
    d[i] -= std::fabs(e[i - 1]);
  }
  std::vector<double> Q((size_t)n * n, 0.0), Q2((size_t)n * n, 0.0), lam(n);
  std::vector<Block> blocks;
  for (int b = 0; b < L; ++b) {
    const int off = loff[b], s = loff[b + 1] - off;
    std::vector<double> A((size_t)s * s, 0.0), V;
    for (int i = 0; i < s; ++i) {
      A[i + (size_t)i * s] = d[off + i];
      if (i + 1 < s) A[i + 1 + (size_t)i * s] = A[i + (size_t)(i + 1) * s] = e[off + i];
    }
    jacobi_eig(A, s, V);
    for (int i = 0; i < s; ++i) lam[off + i] = A[i + (size_t)i * s];
    for (int c = 0; c < s; ++c)
      for (int r = 0; r < s; ++r) Q[(off + r) + (size_t)(off + c) * n] = V[r + (size_t)c * s];
    blocks.push_back({off, s});
  }
  stats[0] = stats[1] = stats[2] = 0;
  while (blocks.size() > 1) {
    std::vector<Block> next;
    for (size_t b = 0; b + 1 < blocks.size(); b += 2) {
      const int off = blocks[b].off, n1 = blocks[b].size, n2 = blocks[b + 1].size, nm = n1 + n2;
      const double beta = e[off + n1 - 1];
      std::vector<double> z(nm);
      for (int c = 0; c < nm; ++c)
        z[c] = (c < n1) ? Q[(off + n1 - 1) + (size_t)(off + c) * n] : (beta < 0 ? -1.0 : 1.0) * Q[(off + n1) + (size_t)(off + c) * n];
      MergePlan P;
      plan_merge(&lam[off], z.data(), beta, nm, P);
      const int k = P.k;
      std::vector<int> org(k);
      std::vector<double> tau(k), lamn(k);
      for (int j = 0; j < k; ++j) {
        secular_root(k, j, P.dl.data(), P.zz.data(), P.rho, 0, 1, NoReduce(), &org[j], &tau[j]);
        lamn[j] = P.dl[org[j]] + tau[j];
      }
      std::vector<int> pos_nd, pos_def;
      std::vector<double> dnew;
      order_outputs(P, lamn, pos_nd, pos_def, dnew);
      // M = R S
      std::vector<double> M((size_t)nm * nm, 0.0), zh(k);
      for (int i = 0; i < k; ++i) {
        const double p = loewner_partial(k, i, P.dl.data(), org.data(), tau.data(), 0, 1);
        zh[i] = std::copysign(std::sqrt(std::fabs(p)), P.zz[i]);
      }
      for (int j = 0; j < k; ++j) {
        double nrm = 0;
        std::vector<double> u(k);
        for (int i = 0; i < k; ++i) {
          u[i] = zh[i] / delta(P.dl.data(), i, org[j], tau[j]);
          nrm += u[i] * u[i];
        }
        nrm = std::sqrt(nrm);
        for (int i = 0; i < k; ++i) M[P.idx_nd[i] + (size_t)pos_nd[j] * nm] = u[i] / nrm;
      }
      for (size_t t = 0; t < P.idx_def.size(); ++t) M[P.idx_def[t] + (size_t)pos_def[t] * nm] = 1.0;
      for (int r = (int)P.rots.size() - 1; r >= 0; --r) {
        const Rotation& R = P.rots[r];
        for (int c = 0; c < nm; ++c) {
          const double xa = M[R.a + (size_t)c * nm], xb = M[R.b + (size_t)c * nm];
          M[R.a + (size_t)c * nm] = R.c * xa - R.s * xb;
          M[R.b + (size_t)c * nm] = R.s * xa + R.c * xb;
        }
      }
      // Q2[block] = Q[block] * M
      for (int c = 0; c < nm; ++c)
        for (int r = 0; r < nm; ++r) {
          double acc = 0;
          for (int t = 0; t < nm; ++t) acc += Q[(off + r) + (size_t)(off + t) * n] * M[t + (size_t)c * nm];
          Q2[(off + r) + (size_t)(off + c) * n] = acc;
        }
      for (int c = 0; c < nm; ++c) lam[off + c] = dnew[c];
      stats[0] += 1;
      stats[1] += (int)P.idx_def.size();
      stats[2] += (int)P.rots.size();
      next.push_back({off, nm});
    }
    if (blocks.size() % 2 == 1) {
      const Block& B = blocks.back();
      for (int c = 0; c < B.size; ++c)
        for (int r = 0; r < B.size; ++r) Q2[(B.off + r) + (size_t)(B.off + c) * n] = Q[(B.off + r) + (size_t)(B.off + c) * n];
      next.push_back(B);
    }
    Q.swap(Q2);
    std::fill(Q2.begin(), Q2.end(), 0.0);
    blocks.swap(next);
  }
  // a single leaf is not sorted yet
  std::vector<int> ord(n);
  for (int i = 0; i < n; ++i) ord[i] = i;
  std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return lam[a] < lam[b]; });
  for (int c = 0; c < n; ++c) {
    lam_out[c] = lam[ord[c]];
    std::memcpy(Q_out + (size_t)c * n, &Q[(size_t)ord[c] * n], sizeof(double) * n);
  }
  return 0;
}

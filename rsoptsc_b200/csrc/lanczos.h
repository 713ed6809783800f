// Lanczos engine + dense matvec helper (lanczos.cu).
#pragma once
#include <functional>
#include <vector>

#include "common.h"

namespace rs {

struct DenseMatvec {  // column-major rows x cols matrix on the device
  const double* M = nullptr;
  int64_t rows = 0, cols = 0, chunk = 0;
  int nchunk = 0;
  DevBuf<double> partial;
  void init(const double* M, int64_t rows, int64_t cols);
  void mul_t(const double* x, double* y);  // y(cols) = M^T x(rows)
  void mul_n(const double* x, double* y);  // y(rows) = M x(cols)
};

struct Lanczos {  // symmetric Lanczos with full reorthogonalisation on a device operator
  int64_t dim = 0;
  int max_steps = 0, steps = 0;
  bool exhausted = false;
  bool nonfinite = false;  // NaN/Inf met in the recurrence (the operand holds non-finite values)
  std::function<void(const double*, double*)> op;  // w = B v (device pointers)
  DevBuf<double> Q, w, coef, d_alpha, d_beta, partial;
  std::vector<double> alpha, beta;
  // d_start (optional, dim values on the device): start vector instead of the built-in pseudo-random one
  void init(int64_t dim, int max_steps, std::function<void(const double*, double*)> op, const double* d_start = nullptr);
  void advance(int more);
  // eigen-decomposition of the current tridiagonal: theta ascending, S steps x steps column-major
  void ritz(std::vector<double>& theta, std::vector<double>& S, bool last_row_only = false) const;
  // d_Y (dim x sel.size(), column-major) = Q * S[:, sel]
  void ritz_vectors(const std::vector<double>& S, const std::vector<int>& sel, double* d_Y);
};

}  // namespace rs

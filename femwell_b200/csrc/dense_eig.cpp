// Small dense complex eigen-solver (host) for the projected Arnoldi matrix (ncv x ncv, ncv <~ 100).
// Plays the role of the Hessenberg QR / Ritz extraction inside ARPACK's *neupd
// (scipy arpack.py:801-853 drives it).  Householder Hessenberg reduction, explicit
// single-shift complex QR with Wilkinson shifts to Schur form, back-substitution for eigenvectors.
#include <algorithm>
#include <complex>
#include <vector>

#include "solver.h"

namespace fw {

using cd = std::complex<double>;

namespace {
struct Mat {
  int n;
  std::vector<cd> a;  // column-major
  explicit Mat(int n_) : n(n_), a((size_t)n_ * n_, cd(0)) {}
  cd& operator()(int i, int j) { return a[(size_t)j * n + i]; }
  const cd& operator()(int i, int j) const { return a[(size_t)j * n + i]; }
};

void hessenberg(Mat& H, Mat& Z) {
  const int n = H.n;
  for (int i = 0; i < n; ++i) Z(i, i) = 1.0;
  std::vector<cd> v(n);
  for (int k = 0; k + 2 < n; ++k) {
    double alpha2 = 0;
    for (int i = k + 1; i < n; ++i) alpha2 += std::norm(H(i, k));
    double below = alpha2 - std::norm(H(k + 1, k));
    if (below <= 0) continue;
    double alpha = std::sqrt(alpha2);
    cd x0 = H(k + 1, k);
    cd phase = (std::abs(x0) > 0) ? x0 / std::abs(x0) : cd(1.0);
    for (int i = 0; i < n; ++i) v[i] = 0;
    for (int i = k + 1; i < n; ++i) v[i] = H(i, k);
    v[k + 1] += phase * alpha;
    double vn2 = 0;
    for (int i = k + 1; i < n; ++i) vn2 += std::norm(v[i]);
    if (vn2 == 0) continue;
    // P = I - 2 v v^H / (v^H v) ;  H <- P H P ; Z <- Z P
    for (int j = 0; j < n; ++j) {
      cd s = 0;
      for (int i = k + 1; i < n; ++i) s += std::conj(v[i]) * H(i, j);
      s *= 2.0 / vn2;
      for (int i = k + 1; i < n; ++i) H(i, j) -= s * v[i];
    }
    for (int i = 0; i < n; ++i) {
      cd s = 0;
      for (int j = k + 1; j < n; ++j) s += H(i, j) * v[j];
      s *= 2.0 / vn2;
      for (int j = k + 1; j < n; ++j) H(i, j) -= s * std::conj(v[j]);
    }
    for (int i = 0; i < n; ++i) {
      cd s = 0;
      for (int j = k + 1; j < n; ++j) s += Z(i, j) * v[j];
      s *= 2.0 / vn2;
      for (int j = k + 1; j < n; ++j) Z(i, j) -= s * std::conj(v[j]);
    }
    for (int i = k + 2; i < n; ++i) H(i, k) = 0;
  }
}

// returns false if the QR iteration failed to converge
bool schur_qr(Mat& H, Mat& Z) {
  const int n = H.n;
  const double eps = 2.220446049250313e-16;
  double hnorm = 0;
  for (auto& x : H.a) hnorm = std::max(hnorm, std::abs(x));
  if (hnorm == 0) return true;
  int hi = n - 1, iter = 0, total = 0;
  std::vector<double> cs(n);
  std::vector<cd> sn(n);
  while (hi >= 0) {
    int l = hi;
    while (l > 0) {
      double s = std::abs(H(l - 1, l - 1)) + std::abs(H(l, l));
      if (s == 0) s = hnorm;
      if (std::abs(H(l, l - 1)) <= eps * s) {
        H(l, l - 1) = 0;
        break;
      }
      --l;
    }
    if (l == hi) {
      --hi;
      iter = 0;
      continue;
    }
    if (++total > 60 * n + 200) return false;
    ++iter;
    cd mu;
    if (iter % 11 == 10) {
      mu = H(hi, hi) + cd(std::abs(H(hi, hi - 1)), std::abs(hi >= 2 ? H(hi - 1, hi - 2) : cd(0)));
    } else {
      cd a = H(hi - 1, hi - 1), b = H(hi - 1, hi), c = H(hi, hi - 1), d = H(hi, hi);
      cd tr2 = 0.5 * (a + d);
      cd disc = std::sqrt((0.5 * (a - d)) * (0.5 * (a - d)) + b * c);
      cd e1 = tr2 + disc, e2 = tr2 - disc;
      mu = (std::abs(e1 - d) < std::abs(e2 - d)) ? e1 : e2;
    }
    for (int i = l; i <= hi; ++i) H(i, i) -= mu;
    for (int k = l; k < hi; ++k) {
      cd a = H(k, k), b = H(k + 1, k);
      double c;
      cd s;
      if (b == cd(0)) {
        c = 1, s = 0;
      } else if (a == cd(0)) {
        c = 0, s = std::conj(b) / std::abs(b);
      } else {
        double na = std::abs(a), nrm = std::hypot(na, std::abs(b));
        cd alpha = a / na;
        c = na / nrm;
        s = alpha * std::conj(b) / nrm;
      }
      cs[k] = c, sn[k] = s;
      for (int j = k; j < n; ++j) {
        cd x = H(k, j), y = H(k + 1, j);
        H(k, j) = c * x + s * y;
        H(k + 1, j) = -std::conj(s) * x + c * y;
      }
      H(k + 1, k) = 0;
    }
    for (int k = l; k < hi; ++k) {
      double c = cs[k];
      cd s = sn[k];
      int top = std::min(k + 1, hi);
      for (int i = 0; i <= top; ++i) {
        cd x = H(i, k), y = H(i, k + 1);
        H(i, k) = c * x + std::conj(s) * y;
        H(i, k + 1) = -s * x + c * y;
      }
      for (int i = 0; i < n; ++i) {
        cd x = Z(i, k), y = Z(i, k + 1);
        Z(i, k) = c * x + std::conj(s) * y;
        Z(i, k + 1) = -s * x + c * y;
      }
    }
    for (int i = l; i <= hi; ++i) H(i, i) += mu;
  }
  return true;
}
}  // namespace

// a: n x n column-major (not modified).  w: eigenvalues, v: eigenvectors (columns, unit 2-norm).
bool dense_eig(int n, const cd* a, cd* w, cd* v) {
  if (n <= 0) return true;
  Mat H(n), Z(n);
  std::copy(a, a + (size_t)n * n, H.a.begin());
  hessenberg(H, Z);
  if (!schur_qr(H, Z)) return false;
  const double eps = 2.220446049250313e-16;
  double tnorm = 0;
  for (int j = 0; j < n; ++j)
    for (int i = 0; i <= j; ++i) tnorm = std::max(tnorm, std::abs(H(i, j)));
  const double smin = std::max(eps * tnorm, 1e-300);
  std::vector<cd> x(n);
  for (int k = 0; k < n; ++k) {
    w[k] = H(k, k);
    for (int i = 0; i < n; ++i) x[i] = 0;
    x[k] = 1.0;
    for (int i = k - 1; i >= 0; --i) {
      cd s = 0;
      for (int j = i + 1; j <= k; ++j) s += H(i, j) * x[j];
      cd d = H(i, i) - H(k, k);
      if (std::abs(d) < smin) d = smin;
      x[i] = -s / d;
    }
    double nrm = 0;
    std::vector<cd> y(n, cd(0));
    for (int j = 0; j <= k; ++j)
      for (int i = 0; i < n; ++i) y[i] += Z(i, j) * x[j];
    for (int i = 0; i < n; ++i) nrm += std::norm(y[i]);
    nrm = std::sqrt(nrm);
    for (int i = 0; i < n; ++i) v[(size_t)k * n + i] = y[i] / nrm;
  }
  return true;
}

}  // namespace fw

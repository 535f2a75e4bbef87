// eigb200 -- secular-equation root finder for the divide-and-conquer merge (host/device).
//
// Role in the reference: the host ?stedc('I') call (zheevd_gpu.F90:101 / dsyevd_gpu.F90:99) ends in LAPACK
// dlaed4 for every non-deflated root; here the same mathematics runs on the device.  This is a restatement
// of the published algorithm (Li, LAPACK Working Note 89: rational "middle way" interpolation with the
// origin shifted to the nearer pole, bracketed by bisection), not a translation of dlaed4.
//
// Problem: roots of f(x) = 1/rho + sum_i z_i^2 / (d_i - x), d ascending, rho > 0, z_i != 0.
// Root j lies in (d_j, d_{j+1}) for j < k-1 and in (d_{k-1}, d_{k-1} + rho*||z||^2) for j = k-1.
// Output: the origin index K (the nearer pole) and tau with  lambda_j = d_K + tau;  the caller forms the
// differences d_i - lambda_j = (d_i - d_K) - tau without cancellation.  iters_out > 80 reports non-convergence
// (dlaed4 would return info = 1; zheevd_gpu.F90:102-106 turns that into info = -1).
//
// The sums over i are supplied by an Evaluator so that the same control flow runs serially on the host
// (unit tests against dlaed4) and warp-cooperatively on the device (every lane executes the identical
// scalar recurrence on identical reduced values => deterministic).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define EIGB_HD __host__ __device__ __forceinline__
#else
#define EIGB_HD inline
#endif

namespace eigb200 {

struct SecularSums {
  double psi, phi, dpsi, dphi;   // sums over i <= j (psi) and i > j (phi) of z^2/delta and z^2/delta^2
  double tj, tj1;                // the two nearest-pole terms z_j^2/delta_j and z_{j+1}^2/delta_{j+1} (0 if absent)
};

// Evaluator contract:  SecularSums ev(int K, double tau)  with  delta_i = (d_i - d_K) - tau.
template <class Evaluator>
EIGB_HD void secular_root(int k, int j, const double* d, const double* z, double rho, double znorm2, Evaluator ev,
                          int& Kout, double& tau_out, int& iters_out) {
  const double eps = 1.1102230246251565e-16;       // LAPACK dlamch('Epsilon') (relative rounding unit), as dlaed4
  const double rhoinv = 1.0 / rho;
  iters_out = 0;
  if (k == 1) { Kout = 0; tau_out = rho * z[0] * z[0]; return; }
  const bool last = (j == k - 1);
  int K;
  double lo, hi, tau;
  if (!last) {
    const double del = d[j + 1] - d[j];
    const double mid = 0.5 * del;
    SecularSums s = ev(j, mid);
    const double c = rhoinv + (s.psi - s.tj) + (s.phi - s.tj1);     // all but the two nearest poles
    const double w = c + s.tj + s.tj1;
    const double zj2 = z[j] * z[j], zj12 = z[j + 1] * z[j + 1];
    if (w >= 0.0) {          // root in the left half: origin d_j, tau in (0, mid]
      K = j; lo = 0.0; hi = mid;
      const double a = c * del + zj2 + zj12, b = zj2 * del;
      const double disc = sqrt(fabs(a * a - 4.0 * b * c));
      tau = (a > 0.0) ? 2.0 * b / (a + disc) : (a - disc) / (2.0 * c);
    } else {                 // origin d_{j+1}, tau in [-mid, 0)
      K = j + 1; lo = -mid; hi = 0.0;
      const double a = c * del - zj2 - zj12, b = zj12 * del;
      const double disc = sqrt(fabs(a * a + 4.0 * b * c));
      tau = (a < 0.0) ? 2.0 * b / (a - disc) : -(a + disc) / (2.0 * c);
    }
    if (!(tau > lo && tau < hi)) tau = 0.5 * (lo + hi);
  } else {
    K = k - 1; lo = 0.0; hi = rho * znorm2 * (1.0 + 4.0 * eps);
    // start from the two-pole model on (d_{k-2}, d_{k-1}) evaluated at the middle of the interval
    tau = 0.5 * hi;
    if (!(tau > lo)) { Kout = K; tau_out = hi; return; }
  }
  const int MAXIT = 80;
  for (int it = 0; it < MAXIT; ++it) {
    iters_out = it + 1;
    SecularSums s = ev(K, tau);
    const double w = rhoinv + s.psi + s.phi;
    const double dw = s.dpsi + s.dphi;
    const double erretm = 8.0 * (s.phi - s.psi) + 2.0 * rhoinv + fabs(tau) * dw;
    if (fabs(w) <= eps * erretm) break;
    if (w < 0.0) lo = fmax(lo, tau); else hi = fmin(hi, tau);
    if (!(hi - lo > 2.0 * eps * fmax(fabs(lo), fabs(hi)))) break;
    // middle-way rational step with poles d_j (Delta1 < 0) and d_{j+1} (Delta2 > 0)
    double eta;
    const double D1 = (d[j] - d[K]) - tau;
    if (!last) {
      const double D2 = (d[j + 1] - d[K]) - tau;
      const double c = w - D1 * s.dpsi - D2 * s.dphi;
      const double a = (D1 + D2) * w - D1 * D2 * dw;
      const double b = D1 * D2 * w;
      if (c == 0.0) {
        eta = b / a;
      } else {
        const double disc = sqrt(fabs(a * a - 4.0 * b * c));
        eta = (a <= 0.0) ? (a - disc) / (2.0 * c) : 2.0 * b / (a + disc);
      }
    } else {
      const double den = w - D1 * s.dpsi;
      eta = (den != 0.0) ? D1 * w / den : -w / dw;
    }
    if (!(w * eta < 0.0)) eta = -w / dw;             // must move against the sign of f (f is increasing)
    double tnew = tau + eta;
    if (!(tnew > lo && tnew < hi)) tnew = 0.5 * (lo + hi);   // bisection safeguard
    if (tnew == tau) break;
    tau = tnew;
    if (it == MAXIT - 1) iters_out = MAXIT + 1;      // ran out of iterations without meeting a stopping test
  }
  Kout = K; tau_out = tau;
}

// Serial evaluator (host tests; also usable by a single device thread).
struct SerialSecularEval {
  int k, j; const double* d; const double* z;
  EIGB_HD SecularSums operator()(int K, double tau) const {
    SecularSums s; s.psi = s.phi = s.dpsi = s.dphi = s.tj = s.tj1 = 0.0;
    const double dK = d[K];
    for (int i = 0; i < k; ++i) {
      const double delta = (d[i] - dK) - tau;
      const double t = z[i] / delta;
      const double term = z[i] * t;
      if (i <= j) { s.psi += term; s.dpsi += t * t; } else { s.phi += term; s.dphi += t * t; }
      if (i == j) s.tj = term;
      if (i == j + 1) s.tj1 = term;
    }
    return s;
  }
};

}  // namespace eigb200

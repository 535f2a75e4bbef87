// Host build of the device secular-equation core (csrc/secular.cuh) for CPU unit tests against dlaed4.
#include "secular.cuh"
using namespace eigb200;
extern "C" int secular_all(int k, const double* d, const double* z, double rho, double* lam, double* delta /*k*k col-major*/,
                           int* iters) {
  double zn2 = 0; for (int i = 0; i < k; ++i) zn2 += z[i] * z[i];
  int maxit = 0;
  for (int j = 0; j < k; ++j) {
    SerialSecularEval ev{k, j, d, z};
    int K, it; double tau;
    secular_root(k, j, d, z, rho, zn2, ev, K, tau, it);
    lam[j] = d[K] + tau;
    for (int i = 0; i < k; ++i) delta[i + (long)j * k] = (d[i] - d[K]) - tau;
    iters[j] = it;
    if (it > maxit) maxit = it;
  }
  return maxit;
}

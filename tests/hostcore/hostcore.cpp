// TEST-ONLY host build of the per-configuration core (psi_core.cuh) so that the kernel
// arithmetic can be checked against the oracle on a CPU-only box.  Never loaded by the
// product package; the product path is the CUDA library and fails loudly without it.
#include "../../ab-initio-flexible-gaussian-basis-neural-network-quantum-monte-carlo_b200/csrc/psi_core.cuh"
#include "../../ab-initio-flexible-gaussian-basis-neural-network-quantum-monte-carlo_b200/csrc/deriv_split.cuh"
#include "../../ab-initio-flexible-gaussian-basis-neural-network-quantum-monte-carlo_b200/csrc/param_grad.cuh"
#include "../../ab-initio-flexible-gaussian-basis-neural-network-quantum-monte-carlo_b200/csrc/dispatch.h"
#include <vector>

using namespace aiqmc;

template <int NE, int NA>
static void run(const AiqmcSystem* sys, const double* P, const double* pos, long n, int mode, double* phase,
                double* logabs, double* grad, double* lap) {
#pragma omp parallel for schedule(dynamic, 4)
  for (long t = 0; t < n; ++t) {
    const double* x = pos + t * 3 * NE;
    if (mode == 0) {
      Psi<NE, NA>::eval_value(*sys, P, x, phase[t], logabs[t]);
    } else if (mode == 1) {
      double dummy;
      Psi<NE, NA>::template eval_deriv<false>(*sys, P, x, phase[t], logabs[t], grad + t * 3 * NE, dummy);
    } else if (mode == 2) {
      Psi<NE, NA>::template eval_deriv<true>(*sys, P, x, phase[t], logabs[t], grad + t * 3 * NE, lap[t]);
    } else if (mode == 5) {      // gradient by the fused reverse sweep (deriv_split.cuh)
      DerivSplit<NE, NA>::grad_reverse(*sys, P, x, phase[t], logabs[t], grad + t * 3 * NE);
    } else if (mode == 6) {      // primal pass into the cache + reverse sweep ON the cache (the N > 16 sweep path)
      std::vector<double> scratch(DerivCache<NE, NA>::SIZE_GRAD);
      DerivSplit<NE, NA>::template primal<false>(*sys, P, x, scratch.data(), 1, nullptr, phase[t], logabs[t]);
      DerivSplit<NE, NA>::grad_reverse_cached(*sys, P, scratch.data(), 1, grad + t * 3 * NE);
    } else {      // 3 / 4: gradient / gradient + Laplacian through the two-pass path (deriv_split.cuh)
      std::vector<double> scratch(DerivCache<NE, NA>::SIZE_LAP);
      double dummy;
      if (mode == 3) DerivSplit<NE, NA>::template eval<false>(*sys, P, x, scratch.data(), phase[t], logabs[t], grad + t * 3 * NE, dummy);
      else DerivSplit<NE, NA>::template eval<true>(*sys, P, x, scratch.data(), phase[t], logabs[t], grad + t * 3 * NE, lap[t]);
    }
  }
}

extern "C" int hc_layout(int n, int a, AiqmcLayout* out) { *out = make_layout(n, a); return 0; }

extern "C" int hc_psi(const AiqmcSystem* sys, const double* P, const double* pos, long n, int mode, double* phase,
                      double* logabs, double* grad, double* lap) {
#define X(NE, NA) \
  if (sys->n_elec == NE && sys->n_atoms == NA) { run<NE, NA>(sys, P, pos, n, mode, phase, logabs, grad, lap); return 0; }
  AIQMC_FOR_EACH_SYSTEM(X)
#undef X
  return -1;
}

// fastmath.cuh bodies, element-wise (host build: the MUFU seeds are replaced by float-precision ones)
extern "C" void hc_fastmath(int which, const double* x, long n, double* out) {
  const double* tab = host_exp_table();
  if (which == 4 || which == 5) {           // batched variants, 4 at a time (n must be a multiple of 4)
    for (long i = 0; i + 4 <= n; i += 4) {
      if (which == 4) ftanh_n<4, 0>(x + i, out + i, tab); else ftanh_n<4, 1>(x + i, out + i, tab);
    }
    return;
  }
  for (long i = 0; i < n; ++i)
    out[i] = which == 0 ? fexp(x[i], tab) : which == 1 ? ftanh(x[i], tab) : which == 2 ? frcp(x[i]) : frsqrt(x[i]);
}

// per-walker d(alpha log|psi| + beta phase)/d(packed params) through csrc/param_grad.cuh; out (n, layout.total)
struct HostSink {
  double* acc;
  void add(int off, double v) { acc[off] += v; }
};
template <int NE, int NA>
static void run_pgrad(const AiqmcSystem* sys, const double* P, const double* pos, long n, const double* alpha,
                      const double* beta, double* out, int cached) {
  const int total = make_layout(NE, NA).total;
#pragma omp parallel for schedule(dynamic, 4)
  for (long t = 0; t < n; ++t) {
    HostSink sink{out + t * total};
    double ph, la;
    if (!cached) {
      ParamGrad<NE, NA>::run(*sys, P, pos + t * 3 * NE, alpha[t], beta[t], ph, la, sink);
    } else {     // primal pass into the derivative cache, then the sweep ON the cache (the N > 16 kernel path)
      std::vector<double> scratch(DerivCache<NE, NA>::SIZE_GRAD);
      DerivSplit<NE, NA>::template primal<false>(*sys, P, pos + t * 3 * NE, scratch.data(), 1, nullptr, ph, la);
      ParamGrad<NE, NA>::run_cached(*sys, P, scratch.data(), 1, alpha[t], beta[t], sink);
    }
  }
}
extern "C" int hc_param_grad(const AiqmcSystem* sys, const double* P, const double* pos, long n, const double* alpha,
                             const double* beta, double* out, int cached) {
#define X(NE, NA) \
  if (sys->n_elec == NE && sys->n_atoms == NA) { run_pgrad<NE, NA>(sys, P, pos, n, alpha, beta, out, cached); return 0; }
  AIQMC_FOR_EACH_SYSTEM(X)
#undef X
  return -1;
}

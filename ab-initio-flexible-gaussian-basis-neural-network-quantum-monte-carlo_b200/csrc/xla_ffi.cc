// xla_ffi.cc -- XLA FFI (jax.ffi) handlers over the C ABI of include/aiqmc_b200.h: the "thin jax.ffi / XLA custom-call"
// boundary north_star names.  The consumers in the reference are
//   jax.grad(logabs_f)            VMC/VMCmcstep.py:41, Energy/hamiltonian.py:104   -> AiqmcPsiFwd / AiqmcPsiGrad (custom_vjp)
//   jax.vmap(local_energy)        Loss/pploss.py:145-153, DMC/total_energy.py:11-19 -> AiqmcLocalEnergyEcp / AiqmcLocalEnergyAe
//   jax.pmap(mc_step)             main/main_pp_adam_muti_GPU.py:131                 -> AiqmcVmcSweep
//   compute_tmoves                DMC/Tmoves.py:32-225                              -> AiqmcDmcTmove
//   jax.grad of the loss          Loss/pploss.py:186-223                            -> AiqmcPsiParamGrad
// Every handler is natively batched (register with vmap_method="broadcast_all"), takes XLA-owned device buffers and
// XLA's stream, gets its scratch from XLA's ScratchAllocator (sized by the library's *_workspace_bytes query) and
// returns ffi::Error; nothing is allocated behind XLA's back.  The static system description travels as attributes
// (AiqmcSystem as int32 words, the ccECP table as float64 words: both are plain-old-data structs of the C ABI).
//
// Built by build.py into libaiqmc_b200_xla.so ONLY where jax.ffi.include_dir() exists; the image this repository was
// developed in has no jax, so there the file is only syntax-checked against a minimal mock of the FFI API
// (tests/ffi_mock, tests/test_xla_ffi_shim.py).  INTEGRATION.md shows the Python registration.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "xla/ffi/api/ffi.h"

#include "../../include/aiqmc_b200.h"

namespace ffi = xla::ffi;

namespace {

using F64 = ffi::Buffer<ffi::F64>;
using RF64 = ffi::ResultBuffer<ffi::F64>;
using Words = ffi::Span<const int32_t>;
using Reals = ffi::Span<const double>;

ffi::Error status(int rc, const char* what) {
  if (rc == AIQMC_OK) return ffi::Error::Success();
  if (rc == AIQMC_E_UNSUPPORTED) return ffi::Error::InvalidArgument(std::string(what) + ": no kernel instantiation for this (n_elec, n_atoms)");
  if (rc == AIQMC_E_BADARG) return ffi::Error::InvalidArgument(std::string(what) + ": bad argument");
  if (rc == AIQMC_E_WORKSPACE) return ffi::Error::Internal(std::string(what) + ": workspace too small");
  return ffi::Error::Internal(std::string(what) + ": CUDA error " + std::to_string(aiqmc_last_cuda_error()));
}

bool load_system(Words w, AiqmcSystem* sys) {
  if (w.size() * sizeof(int32_t) != sizeof(AiqmcSystem)) return false;
  std::memcpy(sys, w.begin(), sizeof(AiqmcSystem));
  return true;
}
bool load_ecp(Reals w, AiqmcEcp* ecp) {                       // the struct as float64 words (its four int32 fields packed in two)
  if (w.size() * sizeof(double) != sizeof(AiqmcEcp)) return false;
  std::memcpy(ecp, w.begin(), sizeof(AiqmcEcp));
  return true;
}
int64_t leading(const F64& pos, int n_elec) { return (int64_t)pos.element_count() / (3 * n_elec); }

#define AQ_SYS(sys_words)                                                      \
  AiqmcSystem sys;                                                             \
  if (!load_system(sys_words, &sys)) return ffi::Error::InvalidArgument("attribute `sys` must hold sizeof(AiqmcSystem) bytes")
#define AQ_SCRATCH(ptr, bytes, what)                                           \
  void* ptr = nullptr;                                                         \
  {                                                                            \
    if ((bytes) < 0) return status((int)(bytes), what);                        \
    auto mem_ = scratch.Allocate((size_t)((bytes) > 0 ? (bytes) : 16));        \
    if (!mem_.has_value()) return ffi::Error::Internal(std::string(what) + ": XLA could not provide the scratch buffer"); \
    ptr = mem_.value();                                                        \
  }

// ---- signed_network(params, pos, ...) -> (phase, log|psi|)                                   nn.py:545-551
ffi::Error PsiFwdImpl(cudaStream_t stream, F64 params, F64 pos, RF64 phase, RF64 logabs, Words sys_words) {
  AQ_SYS(sys_words);
  return status(aiqmc_psi_fwd(&sys, params.typed_data(), pos.typed_data(), leading(pos, sys.n_elec), phase->typed_data(),
                              logabs->typed_data(), stream), "aiqmc_psi_fwd");
}
// ---- + d log|psi| / d pos: the backward rule of custom_vjp(logabs_f)                         VMCmcstep.py:41
ffi::Error PsiGradImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F64 params, F64 pos, RF64 phase, RF64 logabs,
                       RF64 grad, Words sys_words) {
  AQ_SYS(sys_words);
  const int64_t n = leading(pos, sys.n_elec);
  const int64_t bytes = aiqmc_psi_workspace_bytes(&sys, n, 0);
  AQ_SCRATCH(ws, bytes, "aiqmc_psi_grad");
  return status(aiqmc_psi_grad(&sys, params.typed_data(), pos.typed_data(), n, phase->typed_data(), logabs->typed_data(),
                               grad->typed_data(), ws, bytes, stream), "aiqmc_psi_grad");
}
// ---- + Laplacian of log|psi| (forward Laplacian): local_kinetic_energy                       pphamiltonian.py:74-106
ffi::Error PsiFwdLapImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F64 params, F64 pos, RF64 phase, RF64 logabs,
                         RF64 grad, RF64 lap, Words sys_words) {
  AQ_SYS(sys_words);
  const int64_t n = leading(pos, sys.n_elec);
  const int64_t bytes = aiqmc_psi_workspace_bytes(&sys, n, 1);
  AQ_SCRATCH(ws, bytes, "aiqmc_psi_fwdlap");
  return status(aiqmc_psi_fwdlap(&sys, params.typed_data(), pos.typed_data(), n, phase->typed_data(), logabs->typed_data(),
                                 grad->typed_data(), lap->typed_data(), ws, bytes, stream), "aiqmc_psi_fwdlap");
}
// ---- one walkers_update sweep; pos_out aliases pos_in when XLA donates it (pmap(mc_step, donate_argnums=1))
//      gauss2c is the compact (B,N,3) diagonal of the reference's second normal draw            VMCmcstep.py:28-111
ffi::Error VmcSweepImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F64 params, F64 pos_in, F64 gauss1, F64 gauss2c,
                        F64 rnd, RF64 pos_out, ffi::ResultBuffer<ffi::U8> accept, Words sys_words, double tstep,
                        double acyrus, int32_t signed_ratio) {
  AQ_SYS(sys_words);
  const int64_t n = leading(pos_in, sys.n_elec);
  if (pos_out->typed_data() != pos_in.typed_data()) {
    const cudaError_t e = cudaMemcpyAsync(pos_out->typed_data(), pos_in.typed_data(), pos_in.size_bytes(), cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return ffi::Error::Internal("aiqmc_vmc_sweep: copy of the walker positions failed");
  }
  const int64_t bytes = aiqmc_vmc_workspace_bytes(&sys, n);
  AQ_SCRATCH(ws, bytes, "aiqmc_vmc_sweep");
  return status(aiqmc_vmc_sweep_compact(&sys, params.typed_data(), pos_out->typed_data(), gauss1.typed_data(),
                                        gauss2c.typed_data(), rnd.typed_data(), n, tstep, acyrus, signed_ratio,
                                        accept->typed_data(), nullptr, nullptr, ws, bytes, stream), "aiqmc_vmc_sweep");
}
// ---- the same sweep with the random inputs drawn on the device (counter-based Philox keyed by seed / step / walker)
ffi::Error VmcSweepSeededImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F64 params, F64 pos_in, RF64 pos_out,
                              Words sys_words, double tstep, double acyrus, int64_t seed, int32_t step, int64_t walker0) {
  AQ_SYS(sys_words);
  const int64_t n = leading(pos_in, sys.n_elec);
  const int N = sys.n_elec;
  if (pos_out->typed_data() != pos_in.typed_data()) {
    const cudaError_t e = cudaMemcpyAsync(pos_out->typed_data(), pos_in.typed_data(), pos_in.size_bytes(), cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return ffi::Error::Internal("aiqmc_vmc_sweep: copy of the walker positions failed");
  }
  const int64_t wb = aiqmc_vmc_workspace_bytes(&sys, n);
  if (wb < 0) return status((int)wb, "aiqmc_vmc_sweep");
  const int64_t rb = n * (3 * N + 3 * N + N) * (int64_t)sizeof(double);
  AQ_SCRATCH(mem, wb + rb + 256, "aiqmc_vmc_sweep");
  double* g1 = reinterpret_cast<double*>(mem);
  double* g2c = g1 + n * 3 * N;
  double* u = g2c + n * 3 * N;
  void* ws = reinterpret_cast<char*>(mem) + ((rb + 255) / 256) * 256;
  int rc = aiqmc_rng_sweep((uint64_t)seed, (uint32_t)step, walker0, n, N, tstep, g1, g2c, u, stream);
  if (rc != AIQMC_OK) return status(rc, "aiqmc_rng_sweep");
  return status(aiqmc_vmc_sweep_compact(&sys, params.typed_data(), pos_out->typed_data(), g1, g2c, u, n, tstep, acyrus, 0,
                                        nullptr, nullptr, nullptr, ws, wb, stream), "aiqmc_vmc_sweep");
}
// ---- local energy: all-electron (real) and ccECP (complex as (B,2))          hamiltonian.py:236-260, pphamiltonian.py:130-190
ffi::Error LocalEnergyAeImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F64 params, F64 pos, RF64 e_l, Words sys_words) {
  AQ_SYS(sys_words);
  const int64_t n = leading(pos, sys.n_elec);
  const int64_t bytes = aiqmc_energy_workspace_bytes(&sys, n, 0);
  AQ_SCRATCH(ws, bytes, "aiqmc_local_energy_ae");
  return status(aiqmc_local_energy_ae(&sys, params.typed_data(), pos.typed_data(), n, e_l->typed_data(), ws, bytes, stream),
                "aiqmc_local_energy_ae");
}
ffi::Error LocalEnergyEcpImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F64 params, F64 pos, F64 rot, RF64 e_l,
                              Words sys_words, Reals ecp_words) {
  AQ_SYS(sys_words);
  AiqmcEcp ecp;
  if (!load_ecp(ecp_words, &ecp)) return ffi::Error::InvalidArgument("attribute `ecp` must hold sizeof(AiqmcEcp) bytes");
  const int64_t n = leading(pos, sys.n_elec);
  const int64_t bytes = aiqmc_energy_workspace_bytes(&sys, n, 1);
  AQ_SCRATCH(ws, bytes, "aiqmc_local_energy_ecp");
  return status(aiqmc_local_energy_ecp(&sys, &ecp, params.typed_data(), pos.typed_data(), rot.typed_data(), n, e_l->typed_data(),
                                       ws, bytes, stream), "aiqmc_local_energy_ecp");
}
// ---- DMC T-moves                                                                               Tmoves.py:32-225
ffi::Error DmcTmoveImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F64 params, F64 pos, F64 rot, F64 u, F64 rnd,
                        RF64 pos_out, RF64 acceptance, ffi::ResultBuffer<ffi::S32> selected, Words sys_words, Reals ecp_words,
                        double tstep) {
  AQ_SYS(sys_words);
  AiqmcEcp ecp;
  if (!load_ecp(ecp_words, &ecp)) return ffi::Error::InvalidArgument("attribute `ecp` must hold sizeof(AiqmcEcp) bytes");
  const int64_t n = leading(pos, sys.n_elec);
  const int64_t bytes = aiqmc_dmc_tmove_workspace_bytes(&sys, n);
  AQ_SCRATCH(ws, bytes, "aiqmc_dmc_tmove");
  return status(aiqmc_dmc_tmove(&sys, &ecp, params.typed_data(), pos.typed_data(), rot.typed_data(), u.typed_data(),
                                rnd.typed_data(), n, tstep, pos_out->typed_data(), acceptance->typed_data(),
                                selected->typed_data(), ws, bytes, stream), "aiqmc_dmc_tmove");
}
// ---- sum_w alpha_w dlog|psi_w|/dparams + beta_w dphase_w/dparams: the parameter side of the loss JVP   pploss.py:186-223
ffi::Error PsiParamGradImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, F64 params, F64 pos, F64 alpha, F64 beta,
                            RF64 grad, RF64 phase, RF64 logabs, Words sys_words) {
  AQ_SYS(sys_words);
  const int64_t n = leading(pos, sys.n_elec);
  const int64_t bytes = aiqmc_param_grad_workspace_bytes(&sys, n);
  AQ_SCRATCH(ws, bytes, "aiqmc_psi_param_grad");
  return status(aiqmc_psi_param_grad(&sys, params.typed_data(), pos.typed_data(), n, alpha.typed_data(), beta.typed_data(),
                                     grad->typed_data(), phase->typed_data(), logabs->typed_data(), ws, bytes, stream),
                "aiqmc_psi_param_grad");
}

}  // namespace

#define AQ_STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
#define AQ_STREAM_SCRATCH AQ_STREAM.Ctx<ffi::ScratchAllocator>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(AiqmcPsiFwd, PsiFwdImpl,
                              AQ_STREAM.Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Attr<Words>("sys"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(AiqmcPsiGrad, PsiGradImpl,
                              AQ_STREAM_SCRATCH.Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Attr<Words>("sys"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(AiqmcPsiFwdLap, PsiFwdLapImpl,
                              AQ_STREAM_SCRATCH.Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Attr<Words>("sys"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(AiqmcVmcSweep, VmcSweepImpl,
                              AQ_STREAM_SCRATCH.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>()
                                  .Ret<ffi::Buffer<ffi::U8>>().Attr<Words>("sys").Attr<double>("tstep").Attr<double>("acyrus")
                                  .Attr<int32_t>("signed_ratio"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(AiqmcVmcSweepSeeded, VmcSweepSeededImpl,
                              AQ_STREAM_SCRATCH.Arg<F64>().Arg<F64>().Ret<F64>().Attr<Words>("sys").Attr<double>("tstep")
                                  .Attr<double>("acyrus").Attr<int64_t>("seed").Attr<int32_t>("step").Attr<int64_t>("walker0"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(AiqmcLocalEnergyAe, LocalEnergyAeImpl,
                              AQ_STREAM_SCRATCH.Arg<F64>().Arg<F64>().Ret<F64>().Attr<Words>("sys"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(AiqmcLocalEnergyEcp, LocalEnergyEcpImpl,
                              AQ_STREAM_SCRATCH.Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Attr<Words>("sys").Attr<Reals>("ecp"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(AiqmcDmcTmove, DmcTmoveImpl,
                              AQ_STREAM_SCRATCH.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>()
                                  .Ret<ffi::Buffer<ffi::S32>>().Attr<Words>("sys").Attr<Reals>("ecp").Attr<double>("tstep"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(AiqmcPsiParamGrad, PsiParamGradImpl,
                              AQ_STREAM_SCRATCH.Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Attr<Words>("sys"));

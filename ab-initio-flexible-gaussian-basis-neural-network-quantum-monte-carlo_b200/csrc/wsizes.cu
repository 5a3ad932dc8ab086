// workspace size functions shared by abi.cu (kept out of engine_impl.cuh's inline copies)
#include "engine_impl.cuh"
namespace aiqmc {
int64_t sweep_ws_bytes_rt(int n, int a, int64_t B) { return sweep_ws_bytes(n, a, B); }
int64_t psi_ws_bytes_rt(int n, int a, int64_t n_cfg, int with_lap) { return psi_ws_bytes(n, a, n_cfg, with_lap); }
int64_t tmove_ws_bytes_rt(int n, int a, int64_t B) { return tmove_ws_bytes(n, a, B); }
int64_t pgrad_ws_bytes_rt(int n, int a, int64_t B) { return pgrad_ws_bytes(n, a, B); }
int64_t energy_ws_bytes_rt(int n, int a, int64_t B, int with_ecp) { return energy_ws_bytes(n, a, B, with_ecp); }
}

// workspace size functions shared by abi.cu (kept out of engine_impl.cuh's inline copies)
#include "engine_impl.cuh"
namespace aiqmc {
int64_t sweep_ws_bytes_rt(int n, int64_t B) { return sweep_ws_bytes(n, B); }
int64_t energy_ws_bytes_rt(int n, int a, int64_t B, int with_ecp) { return energy_ws_bytes(n, a, B, with_ecp); }
}

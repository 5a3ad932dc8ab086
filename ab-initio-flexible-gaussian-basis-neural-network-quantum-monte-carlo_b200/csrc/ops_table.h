// ops_table.h -- per-system launcher table shared by engine_impl.cuh and abi.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aiqmc_b200.h"

namespace aiqmc {
struct OpsTable {
  int n_elec, n_atoms;
  int (*psi)(const AiqmcSystem*, const double*, const double*, int64_t, int, double*, double*, double*, double*,
             void*, int64_t, cudaStream_t);
  int (*sweep)(const AiqmcSystem*, const double*, double*, const double*, const double*, const double*, int64_t,
               double, double, int, int, uint8_t*, double*, double*, void*, int64_t, cudaStream_t);
  int (*energy)(const AiqmcSystem*, const AiqmcEcp*, const double*, const double*, const double*, int64_t, double*,
                void*, int64_t, int, cudaStream_t);
  int (*tmove)(const AiqmcSystem*, const AiqmcEcp*, const double*, const double*, const double*, const double*,
               const double*, int64_t, double, double*, double*, int32_t*, void*, int64_t, cudaStream_t);
  int (*param_grad)(const AiqmcSystem*, const double*, const double*, int64_t, const double*, const double*, double*,
                    double*, double*, void*, int64_t, cudaStream_t);
};
}  // namespace aiqmc

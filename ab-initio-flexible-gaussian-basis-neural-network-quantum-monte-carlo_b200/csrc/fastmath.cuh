// fastmath.cuh -- double-precision transcendental overrides for the device build.
// Default: CUDA libm (tanh/exp, <= 1-2 ulp).  Define AIQMC_FAST_TANH to use the in-house
// exp-based tanh below (documented max error in DESIGN.md).
#pragma once

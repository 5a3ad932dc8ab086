// fp64_pipes.cu -- B200 (sm_100a) microbenchmark: is the FP64 tensor path (DMMA, mma.sync f64) a second pipe next to
// the FP64 FMA pipe, or the same units?  Measures, chip-wide, in TFLOP/s:
//   dfma        : independent DFMA chains only
//   dmma_884    : mma.sync.aligned.m8n8k4.f64 only
//   dmma_1688   : mma.sync.aligned.m16n8k8.f64 only
//   dmma_16816  : mma.sync.aligned.m16n8k16.f64 only
//   mix_*       : DFMA chains and DMMA interleaved in the same warp (flops of both counted)
//   ffma        : FP32 FMA chains only;  mix_ffma_dfma: FP32 and FP64 FMA chains interleaved
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fp64_pipes fp64_pipes.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// NF DFMA chains, NM m8n8k4 accumulator pairs per thread
template <int NF, int NM>
__global__ void __launch_bounds__(256) k_mix884(long iters, double* sink) {
  double f[NF > 0 ? NF : 1], c0[NM > 0 ? NM : 1], c1[NM > 0 ? NM : 1];
  const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-12 * blockIdx.x;
  for (int i = 0; i < NF; ++i) f[i] = i + threadIdx.x;
  for (int i = 0; i < NM; ++i) { c0[i] = i; c1[i] = -i; }
  for (long it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < (NF > NM ? NF : NM); ++i) {
        if (i < NM) dmma884(c0[i], c1[i], x, y);
        if (i < NF) f[i] = fma(f[i], x, y);
      }
    }
  }
  double s = 0.0;
  for (int i = 0; i < NF; ++i) s += f[i];
  for (int i = 0; i < NM; ++i) s += c0[i] + c1[i];
  if (s == 123.456) sink[0] = s;
}

template <int NF, int NM>
__global__ void __launch_bounds__(256) k_mix1688(long iters, double* sink) {
  double f[NF > 0 ? NF : 1], c[NM > 0 ? NM : 1][4];
  const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-12 * blockIdx.x;
  double a[4] = {x, y, x, y}, b[2] = {y, x};
  for (int i = 0; i < NF; ++i) f[i] = i + threadIdx.x;
  for (int i = 0; i < NM; ++i) for (int q = 0; q < 4; ++q) c[i][q] = i + q;
  for (long it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < (NF > NM ? NF : NM); ++i) {
        if (i < NM) dmma1688(c[i], a, b);
        if (i < NF) f[i] = fma(f[i], x, y);
      }
    }
  }
  double s = 0.0;
  for (int i = 0; i < NF; ++i) s += f[i];
  for (int i = 0; i < NM; ++i) for (int q = 0; q < 4; ++q) s += c[i][q];
  if (s == 123.456) sink[0] = s;
}

template <int NM>
__global__ void __launch_bounds__(256) k_16816(long iters, double* sink) {
  double c[NM][4];
  const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-12 * blockIdx.x;
  double a[8] = {x, y, x, y, y, x, y, x}, b[4] = {y, x, x, y};
  for (int i = 0; i < NM; ++i) for (int q = 0; q < 4; ++q) c[i][q] = i + q;
  for (long it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < NM; ++i) dmma16816(c[i], a, b);
  }
  double s = 0.0;
  for (int i = 0; i < NM; ++i) for (int q = 0; q < 4; ++q) s += c[i][q];
  if (s == 123.456) sink[0] = s;
}

// NS FP32 chains + ND FP64 chains per thread
template <int NS, int ND>
__global__ void __launch_bounds__(256) k_ffma_dfma(long iters, double* sink) {
  float g[NS > 0 ? NS : 1]; double f[ND > 0 ? ND : 1];
  const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-12 * blockIdx.x;
  const float xs = (float)x, ys = (float)y;
  for (int i = 0; i < NS; ++i) g[i] = i + threadIdx.x;
  for (int i = 0; i < ND; ++i) f[i] = i + threadIdx.x;
  for (long it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < (NS > ND ? NS : ND); ++i) {
        if (i < ND) f[i] = fma(f[i], x, y);
        if (i < NS) g[i] = fmaf(g[i], xs, ys);
      }
    }
  }
  double s = 0.0;
  for (int i = 0; i < NS; ++i) s += g[i];
  for (int i = 0; i < ND; ++i) s += f[i];
  if (s == 123.456) sink[0] = s;
}

template <class K>
static double run(K kernel, long iters, double flops_per_thread_iter, double* sink, const char* name, int threads = 256, int ctas_per_sm = 8) {
  int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  const int grid = p.multiProcessorCount * ctas_per_sm;
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(a));
    kernel<<<grid, threads>>>(iters, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    const double tf = (double)grid * threads * flops_per_thread_iter * iters / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  printf("%-28s %8.2f TFLOP/s\n", name, best);
  fflush(stdout);
  return best;
}

int main() {
  double* sink; CK(cudaMalloc(&sink, 64));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device: %s, %d SMs, clock %.0f MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1e3);
  const long it = 20000;
  // flops per thread per iteration: DFMA chain = 2 per fma x 4 reps; m8n8k4 = 2*8*8*4/32 = 16 per mma per thread x 4 reps;
  // m16n8k8 = 2*16*8*8/32 = 64; m16n8k16 = 128
  run(k_mix884<8, 0>, it, 8 * 2 * 4, sink, "dfma (8 chains)");
  run(k_mix884<0, 4>, it, 4 * 16 * 4, sink, "dmma m8n8k4 (4 acc)");
  run(k_mix884<0, 8>, it, 8 * 16 * 4, sink, "dmma m8n8k4 (8 acc)");
  run(k_mix1688<0, 4>, it, 4 * 64 * 4, sink, "dmma m16n8k8 (4 acc)");
  run(k_16816<4>, it, 4 * 128 * 4, sink, "dmma m16n8k16 (4 acc)");
  run(k_mix884<8, 4>, it, (8 * 2 + 4 * 16) * 4, sink, "mix dfma8 + m8n8k4 x4");
  run(k_mix884<8, 8>, it, (8 * 2 + 8 * 16) * 4, sink, "mix dfma8 + m8n8k4 x8");
  run(k_mix884<8, 1>, it, (8 * 2 + 1 * 16) * 4, sink, "mix dfma8 + m8n8k4 x1");
  run(k_mix884<8, 2>, it, (8 * 2 + 2 * 16) * 4, sink, "mix dfma8 + m8n8k4 x2");
  run(k_mix1688<8, 2>, it, (8 * 2 + 2 * 64) * 4, sink, "mix dfma8 + m16n8k8 x2");
  run(k_mix1688<8, 1>, it, (8 * 2 + 1 * 64) * 4, sink, "mix dfma8 + m16n8k8 x1");
  run(k_ffma_dfma<8, 0>, it, 8 * 2 * 4, sink, "ffma (8 chains)");
  run(k_ffma_dfma<16, 0>, it, 16 * 2 * 4, sink, "ffma (16 chains)");
  run(k_ffma_dfma<8, 8>, it, (8 + 8) * 2 * 4, sink, "mix ffma8 + dfma8");
  run(k_ffma_dfma<16, 8>, it, (16 + 8) * 2 * 4, sink, "mix ffma16 + dfma8");
  // the same DFMA loop, time only (to turn the mixes into "DFMA slowdown"): reported above
  return 0;
}

// population.cu -- the HBM-bound end of the walker path: energy statistics (A24), the Q20 global minimum (A27), the
// systematic comb of DMC/branch.py:10-34 with the walker gather of DMC/main_dmc.py:208-242 (A29), and the CROSS-GPU
// population control the reference does not have (it combs per device inside pmap): SURVEY.md section 8e-3.
//
//   energy statistics / e_cut minimum : ONE launch of one 8-CTA thread-block cluster; each CTA reduces a contiguous
//       eighth of the batch, partials meet in CTA 0 through distributed shared memory (fixed order: deterministic).
//       The single-CTA version of round 1 took 31 us for 1 MB (34 GB/s).
//   comb : the inclusive prefix sum is a BLOCKED scan with a fixed block of kScanBlk walkers -- every block is scanned
//       by its own CTA, block totals are scanned sequentially, cum[i] = off[block(i)] + in_block[i].  The association
//       order depends only on kScanBlk, not on the batch size or on how many GPUs hold the walkers: the distributed
//       comb below reproduces the single-GPU comb bit for bit.
//   cross-GPU comb + migration (aiqmc_rebalance_nccl): all-gather of the ranks' BLOCK TOTALS (B/2048 doubles per rank,
//       never the weights or the positions), the same sequential scan of those totals on every rank, the comb teeth
//       located block-first (every rank can do that for every tooth) and then inside the owner's own scan; each rank
//       packs the rows its teeth select, per destination, in tooth order, and ONE grouped ncclSend/ncclRecv exchange
//       moves only the migrating walkers (8*row bytes each); rows that stay on their rank never touch the wire.
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch already loaded, or $AIQMC_NCCL_LIB): the
// library has no link-time dependency on it and loads on machines without NCCL.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <atomic>
#include <math.h>
#include <mutex>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/aiqmc_b200.h"

namespace cg = cooperative_groups;

namespace aiqmc {
extern std::atomic<int> g_last_cuda_error;
extern std::atomic<int64_t> g_launch_count;

constexpr int kScanBlk = 2048;          // walkers per scan block (fixed: part of the numerical definition of the comb)
constexpr int kScanThreads = 256;       // 8 consecutive walkers per thread
constexpr int kCl = 8;                  // CTAs per cluster of the reductions (portable cluster size)
constexpr int kClThreads = 512;

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double wmin(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// [sum Re E, sum Im E, sum |E|^2, count]: Loss/pploss.py:165-167 partials
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kClThreads)
k_energy_stats_cl(const double* __restrict__ e, int stride, int64_t B, double* __restrict__ out) {
  cg::cluster_group cl = cg::this_cluster();
  __shared__ double red[3][kClThreads / 32];
  __shared__ double part[kCl][3];                       // CTA 0's copy collects the cluster
  const unsigned r = cl.block_rank();
  const int64_t chunk = (B + kCl - 1) / kCl;
  const int64_t lo = (int64_t)r * chunk, hi = lo + chunk < B ? lo + chunk : B;
  double sr = 0.0, si = 0.0, s2 = 0.0;
  if (stride == 2) {
    const double2* e2 = reinterpret_cast<const double2*>(e);           // (re, im) pairs: one 16-byte load per walker
    for (int64_t b = lo + threadIdx.x; b < hi; b += kClThreads) { const double2 v = e2[b]; sr += v.x; si += v.y; s2 += v.x * v.x + v.y * v.y; }
  } else {
    for (int64_t b = lo + threadIdx.x; b < hi; b += kClThreads) { const double v = e[b]; sr += v; s2 += v * v; }
  }
  sr = wsum(sr); si = wsum(si); s2 = wsum(s2);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sr; red[1][threadIdx.x >> 5] = si; red[2][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b2 = 0.0, c = 0.0;
    for (int i = 0; i < kClThreads / 32; ++i) { a += red[0][i]; b2 += red[1][i]; c += red[2][i]; }
    double* dst = cl.map_shared_rank(&part[0][0], 0);
    dst[r * 3 + 0] = a; dst[r * 3 + 1] = b2; dst[r * 3 + 2] = c;
  }
  cl.sync();
  if (r == 0 && threadIdx.x == 0) {
    double a = 0.0, b2 = 0.0, c = 0.0;
    for (int i = 0; i < kCl; ++i) { a += part[i][0]; b2 += part[i][1]; c += part[i][2]; }
    out[0] = a; out[1] = b2; out[2] = c; out[3] = (double)B;
  }
}

// Large batches (beyond what one 8-CTA cluster can stream): fixed chunks of kStatChunk walkers per CTA -> partials
// [chunk][3] in the caller's workspace, then one CTA adds the partials in a fixed order.  The result depends on the
// batch size only (not on the GPU), like the single-cluster form.
constexpr int kStatChunk = 4096;
constexpr int64_t kStatSmall = (int64_t)1 << 18;           // up to here the single-cluster kernel is used
__global__ void __launch_bounds__(256) k_energy_partials(const double* __restrict__ e, int stride, int64_t B,
                                                         double* __restrict__ part) {
  __shared__ double red[3][8];
  const int64_t lo = (int64_t)blockIdx.x * kStatChunk, hi = lo + kStatChunk < B ? lo + kStatChunk : B;
  double sr = 0.0, si = 0.0, s2 = 0.0;
  if (stride == 2) {
    const double2* e2 = reinterpret_cast<const double2*>(e);
    for (int64_t b = lo + threadIdx.x; b < hi; b += 256) { const double2 v = e2[b]; sr += v.x; si += v.y; s2 += v.x * v.x + v.y * v.y; }
  } else {
    for (int64_t b = lo + threadIdx.x; b < hi; b += 256) { const double v = e[b]; sr += v; s2 += v * v; }
  }
  sr = wsum(sr); si = wsum(si); s2 = wsum(s2);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sr; red[1][threadIdx.x >> 5] = si; red[2][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double a = 0.0;
    for (int i = 0; i < 8; ++i) a += red[threadIdx.x][i];
    part[(int64_t)blockIdx.x * 3 + threadIdx.x] = a;
  }
}
__global__ void __launch_bounds__(1024) k_energy_partials_sum(const double* __restrict__ part, int nchunk, int64_t B,
                                                              double* __restrict__ out) {
  __shared__ double red[3][32];
  double a[3] = {0.0, 0.0, 0.0};
  for (int c = threadIdx.x; c < nchunk; c += 1024)
    for (int q = 0; q < 3; ++q) a[q] += part[(int64_t)c * 3 + q];
  for (int q = 0; q < 3; ++q) a[q] = wsum(a[q]);
  if ((threadIdx.x & 31) == 0) for (int q = 0; q < 3; ++q) red[q][threadIdx.x >> 5] = a[q];
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += red[threadIdx.x][i];
    out[threadIdx.x] = t;
  }
  if (threadIdx.x == 3) out[3] = (double)B;
}

// min over walkers of min(|E_est - Re E_L|, branchcut): step 1 of comput_S (DMC/S_matrix.py:22-23, quirk Q20)
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kClThreads)
k_ecut_min_cl(const double* __restrict__ e, int stride, int64_t B, double e_est, const double* __restrict__ branchcut,
              double* __restrict__ out) {
  cg::cluster_group cl = cg::this_cluster();
  __shared__ double red[kClThreads / 32];
  __shared__ double part[kCl];
  const unsigned r = cl.block_rank();
  const int64_t chunk = (B + kCl - 1) / kCl;
  const int64_t lo = (int64_t)r * chunk, hi = lo + chunk < B ? lo + chunk : B;
  double m = INFINITY;
  for (int64_t b = lo + threadIdx.x; b < hi; b += kClThreads) m = fmin(m, fmin(fabs(e_est - e[b * stride]), branchcut[b]));
  m = wmin(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = INFINITY;
    for (int i = 0; i < kClThreads / 32; ++i) v = fmin(v, red[i]);
    cl.map_shared_rank(&part[0], 0)[r] = v;
  }
  cl.sync();
  if (r == 0 && threadIdx.x == 0) {
    double v = INFINITY;
    for (int i = 0; i < kCl; ++i) v = fmin(v, part[i]);
    out[0] = v;
  }
}

// ---- blocked inclusive scan: one CTA per block of kScanBlk walkers; thread t owns 8 consecutive walkers; thread
//      totals are scanned by a fixed-shape shuffle scan.  in_block[i] = inclusive sum inside the block, totals[j] = the
//      block's last inclusive value (so that off[j+1] = off[j] + totals[j] equals the cumulative sum at the block end).
__global__ void __launch_bounds__(kScanThreads) k_scan_blocks(const double* __restrict__ w, int64_t B,
                                                              double* __restrict__ in_block, double* __restrict__ totals) {
  __shared__ double wt[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanBlk + (int64_t)threadIdx.x * 8;
  double v[8];
  if (base + 8 <= B) {
    const double2* w2 = reinterpret_cast<const double2*>(w + base);      // 4 x 16-byte loads (base is a multiple of 8 doubles)
#pragma unroll
    for (int q = 0; q < 4; ++q) { const double2 t = w2[q]; v[2 * q] = t.x; v[2 * q + 1] = t.y; }
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = base + q < B ? w[base + q] : 0.0;
  }
#pragma unroll
  for (int q = 1; q < 8; ++q) v[q] += v[q - 1];
  // exclusive scan of the thread totals: inclusive shuffle scan inside the warp, then the warp totals sequentially
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  double inc = v[7];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) wt[wp] = inc;
  __syncthreads();
  double woff = 0.0;
  for (int i = 0; i < wp; ++i) woff += wt[i];
  double excl = __shfl_up_sync(0xffffffffu, inc, 1);        // inclusive value of the previous lane: monotone by construction
  if (lane == 0) excl = 0.0;
  const double off = woff + excl;
  if (base + 8 <= B) {
    double2* o2 = reinterpret_cast<double2*>(in_block + base);
#pragma unroll
    for (int q = 0; q < 4; ++q) o2[q] = make_double2(off + v[2 * q], off + v[2 * q + 1]);
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) if (base + q < B) in_block[base + q] = off + v[q];
  }
  if (threadIdx.x == kScanThreads - 1) totals[blockIdx.x] = off + v[7];
}

// exclusive SEQUENTIAL scan of the block totals of all ranks (<= a few thousand): off[0..nb], off[nb] = total weight
__global__ void k_scan_offsets(const double* __restrict__ totals, int nb, double* __restrict__ off) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double run = 0.0;
    for (int j = 0; j < nb; ++j) { off[j] = run; run += totals[j]; }
    off[nb] = run;
  }
}

// tooth k of the comb: v = (u*wtot + k*wtot/Btot) mod wtot; source = first walker with cumulative weight >= v
// (jnp.searchsorted, side='left', DMC/branch.py:21-23).  Located block-first on the offsets every rank holds, then
// inside the block on the owner's in-block scan.  blk0 = index of this rank's first block among all blocks.
struct Tooth { int blk; int64_t idx; };      // idx: index inside the owner's LOCAL array, or -1 when the block is remote
__device__ __forceinline__ Tooth locate_tooth(int64_t k, int64_t Btot, double u, const double* __restrict__ off, int nb_all,
                                              const double* __restrict__ in_block, int64_t B_local, int blk0, int nb_local) {
  const double wtot = off[nb_all];
  double v = fmod(u * wtot + (double)k * (wtot / (double)Btot), wtot);
  if (v < 0.0) v += wtot;
  int lo = 0, hi = nb_all - 1;                 // first block j with off[j + 1] >= v
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (off[mid + 1] < v) lo = mid + 1; else hi = mid; }
  Tooth t{lo, -1};
  if (lo >= blk0 && lo < blk0 + nb_local) {
    const int64_t b0 = (int64_t)(lo - blk0) * kScanBlk;
    int64_t a = b0, b = b0 + kScanBlk < B_local ? b0 + kScanBlk : B_local;
    const double o = off[lo];
    const int64_t last = b - 1;
    while (a < b) { const int64_t mid = (a + b) >> 1; if (o + in_block[mid] < v) a = mid + 1; else b = mid; }
    t.idx = a > last ? last : a;               // a tooth that rounding pushed past the block end belongs to its last walker
  }
  return t;
}

// single GPU: newinds[k] for this device's own comb
__global__ void k_comb_blocked(const double* __restrict__ in_block, const double* __restrict__ off, int nb, int64_t B, double u,
                               int32_t* __restrict__ newinds, double* __restrict__ new_weight) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= B) return;
  const Tooth t = locate_tooth(k, B, u, off, nb, in_block, B, 0, nb);
  newinds[k] = (int32_t)t.idx;
  if (k == 0) new_weight[0] = off[nb] / (double)B;
}

// multi GPU, every tooth of the GLOBAL comb: owner rank, and the local index when this rank owns it
__global__ void k_teeth(const double* __restrict__ in_block, const double* __restrict__ off, int nb_local, int world, int rank,
                        int64_t B, double u, int32_t* __restrict__ src_rank, int32_t* __restrict__ src_local,
                        double* __restrict__ new_weight) {
  const int64_t Btot = (int64_t)world * B;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Btot) return;
  if (k == 0) new_weight[0] = off[world * nb_local] / (double)Btot;      // branch.py:24-26 made global
  const Tooth t = locate_tooth(k, Btot, u, off, world * nb_local, in_block, B, rank * nb_local, nb_local);
  src_rank[k] = t.blk / nb_local;
  src_local[k] = (int32_t)t.idx;
}

// Stable compaction, one CTA per peer p.
//   mode 0 (send plan): teeth of destination p (slots [p*B, (p+1)*B)) owned by this rank -> send_idx[p*B + pos] = local
//                       source index, in tooth order; counts[p] = how many.
//   mode 1 (recv plan): this rank's own slots fed by source p -> slot_pos[k_local] = position inside p's message;
//                       counts[world + p] = how many.
__global__ void __launch_bounds__(1024) k_plan(const int32_t* __restrict__ src_rank, const int32_t* __restrict__ src_local,
                                               int world, int rank, int64_t B, int32_t* __restrict__ send_idx,
                                               int32_t* __restrict__ slot_pos, int32_t* __restrict__ counts) {
  __shared__ int wcnt[32];
  __shared__ int base_s;
  const int p = blockIdx.x % world, mode = blockIdx.x / world;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  const int64_t k0 = (mode == 0 ? (int64_t)p : (int64_t)rank) * B;
  const int want = mode == 0 ? rank : p;
  for (int64_t c0 = 0; c0 < B; c0 += 1024) {
    const int64_t kl = c0 + threadIdx.x;
    const bool flag = kl < B && src_rank[k0 + kl] == want;
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) wcnt[wp] = __popc(m);
    __syncthreads();
    int woff = 0;
    for (int i = 0; i < wp; ++i) woff += wcnt[i];
    const int pos = base_s + woff + __popc(m & ((1u << lane) - 1u));
    if (flag) {
      if (mode == 0) send_idx[(int64_t)p * B + pos] = src_local[k0 + kl];
      else slot_pos[kl] = pos;
    }
    __syncthreads();
    if (threadIdx.x == 0) { int tot = 0; for (int i = 0; i < 32; ++i) tot += wcnt[i]; base_s += tot; }
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[mode * world + p] = base_s;
}

// Balanced mode: counts[p] = number of teeth rank p owns, and the stable list of THIS rank's teeth, owned_idx[pos] =
// local source index in tooth order (up to world*B entries).  Three small kernels over chunks of 1024 teeth: per-chunk
// counts (all ranks' totals by integer atomics -- exact, order-free; this rank's count per chunk kept), an exclusive
// scan of this rank's chunk counts, the stable scatter.  (Round 2 first had one CTA per rank walking all world*B teeth
// with three barriers per 1024: 0.1 ms at 65,536 teeth and 0.8 ms -- 4.5 % of a DMC step -- at 8 x 65,536.)
constexpr int kOwnChunk = 1024;
__global__ void __launch_bounds__(kOwnChunk) k_owned_count(const int32_t* __restrict__ src_rank, int world, int rank, int64_t Btot,
                                                           int32_t* __restrict__ chunk_cnt, int32_t* __restrict__ counts) {
  __shared__ int cnt[64];
  if (threadIdx.x < 64) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t k = (int64_t)blockIdx.x * kOwnChunk + threadIdx.x;
  const int r = k < Btot ? src_rank[k] : -1;
  const int lane = threadIdx.x & 31;
  for (int p = 0; p < world; ++p) {
    const unsigned m = __ballot_sync(0xffffffffu, r == p);
    if (lane == 0 && m) atomicAdd(&cnt[p], __popc(m));
  }
  __syncthreads();
  if ((int)threadIdx.x < world && cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], cnt[threadIdx.x]);
  if (threadIdx.x == 0) chunk_cnt[blockIdx.x] = cnt[rank];
}
// in-place exclusive scan of v[0..n) by one CTA
__global__ void __launch_bounds__(1024) k_scan_i32(int32_t* __restrict__ v, int n) {
  __shared__ int wtot[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int c0 = 0; c0 < n; c0 += 1024) {
    const int i = c0 + threadIdx.x;
    const int x = i < n ? v[i] : 0;
    int inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wtot[wp] = inc;
    __syncthreads();
    if (wp == 0) {
      int w = wtot[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
      wtot[lane] = wi - w;                                   // exclusive warp offsets
    }
    __syncthreads();
    const int excl = carry + wtot[wp] + inc - x;
    if (i < n) v[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + x;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(kOwnChunk) k_owned_list(const int32_t* __restrict__ src_rank, const int32_t* __restrict__ src_local,
                                                          int rank, int64_t Btot, const int32_t* __restrict__ chunk_off,
                                                          int32_t* __restrict__ owned_idx) {
  __shared__ int wcnt[32];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int64_t k = (int64_t)blockIdx.x * kOwnChunk + threadIdx.x;
  const bool flag = k < Btot && src_rank[k] == rank;
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) wcnt[wp] = __popc(m);
  __syncthreads();
  int woff = 0;
  for (int i = 0; i < wp; ++i) woff += wcnt[i];
  if (flag) owned_idx[chunk_off[blockIdx.x] + woff + __popc(m & ((1u << lane) - 1u))] = src_local[k];
}
// rows[t] = pos[idx[t]], t < n
__global__ void k_gather_rows(const double* __restrict__ pos, const int32_t* __restrict__ idx, int64_t n, int row,
                              double* __restrict__ out) {
  const int64_t total = n * row;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / row;
    out[t] = pos[(int64_t)idx[r] * row + (int)(t - r * row)];
  }
}
__global__ void k_fill_i32(int32_t* __restrict__ out, int64_t n, int32_t v) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = v;
}

// rows[t] = pos[idx[t]] for the concatenated send lists (seg_off[p] = first row of destination p's message)
__global__ void k_pack_rows(const double* __restrict__ pos, const int32_t* __restrict__ send_idx, const int64_t* __restrict__ seg_off,
                            const int32_t* __restrict__ counts, int world, int64_t B, int row, double* __restrict__ sendbuf) {
  const int p = blockIdx.y;
  const int64_t n = (int64_t)counts[p] * row;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / row;
    const int c = (int)(t - r * row);
    sendbuf[(seg_off[p] + r) * row + c] = pos[(int64_t)send_idx[(int64_t)p * B + r] * row + c];
  }
}
// out[k] = recvbuf[roff[src] + slot_pos[k]]
__global__ void k_unpack_rows(const double* __restrict__ recvbuf, const int32_t* __restrict__ src_rank, const int32_t* __restrict__ slot_pos,
                              const int64_t* __restrict__ roff, int rank, int64_t B, int row, double* __restrict__ out,
                              int32_t* __restrict__ src_global) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * row) return;
  const int64_t k = t / row;
  const int c = (int)(t - k * row);
  const int s = src_rank[(int64_t)rank * B + k];
  out[t] = recvbuf[(roff[s] + slot_pos[k]) * row + c];
  if (c == 0 && src_global) src_global[k] = s;
}

// gather with 16-byte accesses when the row allows it
// Row gather out[b] = in[idx[b]] (A29, the reconfiguration of DMC/main_dmc.py:208-242).  HBM-bound: a thread moves
// kGatherU elements 256 apart (all loads in flight before the first store; coalesced on both sides within a row run),
// 32-bit index arithmetic (the launcher falls back to one element per thread beyond 2^31 elements).
constexpr int kGatherU = 4;
template <class T>
__global__ void __launch_bounds__(256) k_gather_rows_u(const T* __restrict__ in, const int32_t* __restrict__ idx, uint32_t nt,
                                                       uint32_t row, T* __restrict__ out) {
  const uint32_t base = blockIdx.x * (256u * kGatherU) + threadIdx.x;
  T v[kGatherU];
#pragma unroll
  for (int u = 0; u < kGatherU; ++u) {
    const uint32_t t = base + 256u * u;
    if (t < nt) {
      const uint32_t b = t / row, c = t - b * row;
      v[u] = in[(int64_t)idx[b] * row + c];
    }
  }
#pragma unroll
  for (int u = 0; u < kGatherU; ++u) {
    const uint32_t t = base + 256u * u;
    if (t < nt) out[t] = v[u];
  }
}
template <class T>
__global__ void k_gather_rows_1(const T* __restrict__ in, const int32_t* __restrict__ idx, int64_t B, int row, T* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * row) return;
  const int64_t b = t / row;
  const int c = (int)(t - b * row);
  out[t] = in[(int64_t)idx[b] * row + c];
}
template <class T>
static void launch_gather(const T* in, const int32_t* idx, int64_t B, int row, T* out, cudaStream_t st) {
  const int64_t nt = B * row;
  if (nt < (int64_t)1 << 31) {
    const unsigned per = 256u * kGatherU;
    k_gather_rows_u<T><<<(unsigned)((nt + per - 1) / per), 256, 0, st>>>(in, idx, (uint32_t)nt, (uint32_t)row, out);
  } else {
    k_gather_rows_1<T><<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(in, idx, B, row, out);
  }
}

// ------------------------------------------------------------------ NCCL, bound at run time
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  bool ok = false;
};
static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* env = getenv("AIQMC_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
#define AQ_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name))
    AQ_SYM(GetUniqueId, "ncclGetUniqueId"); AQ_SYM(CommInitRank, "ncclCommInitRank"); AQ_SYM(CommDestroy, "ncclCommDestroy");
    AQ_SYM(AllGather, "ncclAllGather"); AQ_SYM(AllReduce, "ncclAllReduce"); AQ_SYM(Send, "ncclSend"); AQ_SYM(Recv, "ncclRecv");
    AQ_SYM(GroupStart, "ncclGroupStart"); AQ_SYM(GroupEnd, "ncclGroupEnd");
#undef AQ_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.AllReduce && api.Send && api.Recv &&
             api.GroupStart && api.GroupEnd;
  });
  return api.ok ? &api : nullptr;
}
}  // namespace aiqmc

using namespace aiqmc;
#define AQ_CUDA_OK(call)                                                                 \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) { g_last_cuda_error = (int)e_; return AIQMC_E_CUDA; }         \
  } while (0)
#define AQ_NCCL_OK(call)                                                                 \
  do {                                                                                   \
    if ((call) != ncclSuccess) return AIQMC_E_NCCL;                                      \
  } while (0)

static int64_t scan_blocks_of(int64_t B) { return (B + kScanBlk - 1) / kScanBlk; }
static int64_t al256(int64_t v) { return (v + 255) & ~(int64_t)255; }

extern "C" {

int aiqmc_energy_stats(const double* e_l, int32_t e_l_stride, int64_t n_walkers, double* stats, void* stream) {
  if (!e_l || !stats || n_walkers < 0 || (e_l_stride != 1 && e_l_stride != 2)) return AIQMC_E_BADARG;
  ++g_launch_count;
  k_energy_stats_cl<<<kCl, kClThreads, 0, (cudaStream_t)stream>>>(e_l, e_l_stride, n_walkers, stats);
  AQ_CUDA_OK(cudaGetLastError());
  return AIQMC_OK;
}

int64_t aiqmc_energy_stats_workspace_bytes(int64_t n_walkers) {
  if (n_walkers < 0) return AIQMC_E_BADARG;
  return n_walkers <= kStatSmall ? 0 : al256(((n_walkers + kStatChunk - 1) / kStatChunk) * 3 * 8);
}
int aiqmc_energy_stats_ws(const double* e_l, int32_t e_l_stride, int64_t n_walkers, double* stats, void* workspace,
                          int64_t workspace_bytes, void* stream) {
  if (!e_l || !stats || n_walkers < 0 || (e_l_stride != 1 && e_l_stride != 2)) return AIQMC_E_BADARG;
  if (n_walkers <= kStatSmall) return aiqmc_energy_stats(e_l, e_l_stride, n_walkers, stats, stream);
  if (!workspace || workspace_bytes < aiqmc_energy_stats_workspace_bytes(n_walkers)) return AIQMC_E_WORKSPACE;
  if (e_l_stride == 2 && ((uintptr_t)e_l & 15) != 0) return AIQMC_E_BADARG;
  const int nchunk = (int)((n_walkers + kStatChunk - 1) / kStatChunk);
  g_launch_count += 2;
  k_energy_partials<<<nchunk, 256, 0, (cudaStream_t)stream>>>(e_l, e_l_stride, n_walkers, (double*)workspace);
  k_energy_partials_sum<<<1, 1024, 0, (cudaStream_t)stream>>>((const double*)workspace, nchunk, n_walkers, stats);
  AQ_CUDA_OK(cudaGetLastError());
  return AIQMC_OK;
}

int aiqmc_dmc_ecut_min(const double* e_l, int32_t e_l_stride, int64_t n_walkers, double e_est, const double* branchcut,
                       double* ecut_min, void* stream) {
  if (!e_l || !branchcut || !ecut_min || n_walkers < 0 || (e_l_stride != 1 && e_l_stride != 2)) return AIQMC_E_BADARG;
  ++g_launch_count;
  k_ecut_min_cl<<<kCl, kClThreads, 0, (cudaStream_t)stream>>>(e_l, e_l_stride, n_walkers, e_est, branchcut, ecut_min);
  AQ_CUDA_OK(cudaGetLastError());
  return AIQMC_OK;
}

int64_t aiqmc_branch_workspace_bytes(int64_t n_walkers) {
  if (n_walkers < 0) return AIQMC_E_BADARG;
  return al256(n_walkers * 8) + al256((2 * scan_blocks_of(n_walkers) + 2) * 8);
}
int aiqmc_branch_comb(const double* weights, int64_t n_walkers, double u, int32_t* newinds, double* new_weight,
                      void* workspace, int64_t workspace_bytes, void* stream) {
  if (!weights || !newinds || !new_weight || !workspace || n_walkers <= 0) return AIQMC_E_BADARG;
  if (workspace_bytes < aiqmc_branch_workspace_bytes(n_walkers)) return AIQMC_E_WORKSPACE;
  if (((uintptr_t)weights & 15) != 0) return AIQMC_E_BADARG;             // 16-byte loads
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (int)scan_blocks_of(n_walkers);
  double* in_block = (double*)workspace;
  double* totals = (double*)((char*)workspace + al256(n_walkers * 8));
  double* off = totals + nb;
  g_launch_count += 3;
  k_scan_blocks<<<nb, kScanThreads, 0, st>>>(weights, n_walkers, in_block, totals);
  k_scan_offsets<<<1, 32, 0, st>>>(totals, nb, off);
  k_comb_blocked<<<(unsigned)((n_walkers + 255) / 256), 256, 0, st>>>(in_block, off, nb, n_walkers, u, newinds, new_weight);
  AQ_CUDA_OK(cudaGetLastError());
  return AIQMC_OK;
}

int aiqmc_gather_walkers(const double* pos_in, const int32_t* newinds, int64_t n_walkers, int32_t row_doubles,
                         double* pos_out, void* stream) {
  if (n_walkers < 0 || row_doubles < 1) return AIQMC_E_BADARG;
  if (n_walkers == 0) return AIQMC_OK;                               // an empty selection may come with null buffers
  if (!pos_in || !newinds || !pos_out) return AIQMC_E_BADARG;
  ++g_launch_count;
  if (row_doubles % 2 == 0 && (((uintptr_t)pos_in | (uintptr_t)pos_out) & 15) == 0)
    launch_gather<double2>((const double2*)pos_in, newinds, n_walkers, row_doubles / 2, (double2*)pos_out, (cudaStream_t)stream);
  else
    launch_gather<double>(pos_in, newinds, n_walkers, row_doubles, pos_out, (cudaStream_t)stream);
  AQ_CUDA_OK(cudaGetLastError());
  return AIQMC_OK;
}

// ---- NCCL communicator owned by the library's caller -------------------------------------------------
int aiqmc_nccl_available(void) { return nccl_api() != nullptr; }
int aiqmc_nccl_unique_id(void* id128) {
  NcclApi* a = nccl_api();
  if (!a) return AIQMC_E_NCCL;
  if (!id128) return AIQMC_E_BADARG;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  AQ_NCCL_OK(a->GetUniqueId(reinterpret_cast<ncclUniqueId*>(id128)));
  return AIQMC_OK;
}
int aiqmc_nccl_comm_init(int32_t world, int32_t rank, const void* id128, void** comm_out) {
  NcclApi* a = nccl_api();
  if (!a) return AIQMC_E_NCCL;
  if (!id128 || !comm_out || world < 1 || rank < 0 || rank >= world) return AIQMC_E_BADARG;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  AQ_NCCL_OK(a->CommInitRank(&c, world, id, rank));
  // NCCL connects a pair of ranks lazily, on the first send/recv between them (~100 ms per new pair); the comb's
  // rotation sends to different peers from step to step, so every pair is connected here, once, by one tiny grouped
  // all-to-all (measured without it: a DMC step on 8 GPUs took 348 ms instead of 19 ms while new pairs kept appearing)
  if (world > 1) {
    double* scratch = nullptr;
    AQ_CUDA_OK(cudaMalloc(&scratch, 2 * world * sizeof(double)));
    AQ_CUDA_OK(cudaMemset(scratch, 0, 2 * world * sizeof(double)));
    AQ_NCCL_OK(a->GroupStart());
    for (int q = 0; q < world; ++q) {
      if (q == rank) continue;
      AQ_NCCL_OK(a->Send(scratch + q, 1, ncclDouble, q, c, 0));
      AQ_NCCL_OK(a->Recv(scratch + world + q, 1, ncclDouble, q, c, 0));
    }
    AQ_NCCL_OK(a->GroupEnd());
    AQ_CUDA_OK(cudaStreamSynchronize(0));
    AQ_CUDA_OK(cudaFree(scratch));
  }
  *comm_out = c;
  return AIQMC_OK;
}
int aiqmc_nccl_comm_destroy(void* comm) {
  NcclApi* a = nccl_api();
  if (!a) return AIQMC_E_NCCL;
  if (comm) AQ_NCCL_OK(a->CommDestroy((ncclComm_t)comm));
  return AIQMC_OK;
}

// pmean(mean(e_l)) / variance partials (Loss/pploss.py:165-167): in-place SUM all-reduce of stats[4]
int aiqmc_energy_allreduce(double* stats, void* comm, void* stream) {
  NcclApi* a = nccl_api();
  if (!a) return AIQMC_E_NCCL;
  if (!stats || !comm) return AIQMC_E_BADARG;
  AQ_NCCL_OK(a->AllReduce(stats, stats, 4, ncclDouble, ncclSum, (ncclComm_t)comm, (cudaStream_t)stream));
  return AIQMC_OK;
}
// quirk Q20: MIN all-reduce of the e_cut scalar (DMC/S_matrix.py:23)
int aiqmc_ecut_allreduce_min(double* ecut_min, void* comm, void* stream) {
  NcclApi* a = nccl_api();
  if (!a) return AIQMC_E_NCCL;
  if (!ecut_min || !comm) return AIQMC_E_BADARG;
  AQ_NCCL_OK(a->AllReduce(ecut_min, ecut_min, 1, ncclDouble, ncclMin, (ncclComm_t)comm, (cudaStream_t)stream));
  return AIQMC_OK;
}

int64_t aiqmc_rebalance_workspace_bytes(int64_t n_walkers, int32_t row_doubles, int32_t world) {
  if (n_walkers < 0 || row_doubles < 1 || world < 1) return AIQMC_E_BADARG;
  const int64_t nb = scan_blocks_of(n_walkers), Bt = n_walkers * world;
  // the send buffer holds every tooth this rank can own: all world*B of them if one rank carries all the weight
  return al256(n_walkers * 8) + al256((world * nb + 1) * 8 * 2) + 2 * al256(Bt * 4) + al256(Bt * 4) + al256(n_walkers * 4) +
         al256(4 * world * 4 + 64) + al256(4 * (world + 1) * 8) + al256(Bt * row_doubles * 8) + al256(n_walkers * row_doubles * 8);
}

/* see include/aiqmc_b200.h */
int aiqmc_rebalance_nccl(const double* weights, const double* pos, int64_t n_walkers, int32_t row_doubles, double u,
                         int32_t world, int32_t rank, void* comm, int32_t mode, double* pos_out, double* new_weight,
                         int32_t* src_rank_out, int64_t* moved_bytes_out, void* workspace, int64_t workspace_bytes,
                         void* stream) {
  if (!weights || !pos || !pos_out || !new_weight || !workspace || n_walkers <= 0 || row_doubles < 1 || world < 1 || rank < 0 ||
      rank >= world || (mode != 0 && mode != 1))
    return AIQMC_E_BADARG;
  if (workspace_bytes < aiqmc_rebalance_workspace_bytes(n_walkers, row_doubles, world)) return AIQMC_E_WORKSPACE;
  if (((uintptr_t)weights & 15) != 0 || world > 64) return AIQMC_E_BADARG;
  NcclApi* a = world > 1 ? nccl_api() : nullptr;
  if (world > 1 && (!a || !comm)) return AIQMC_E_NCCL;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t B = n_walkers, Bt = B * world;
  const int nb = (int)scan_blocks_of(B), nba = nb * world;
  char* p = (char*)workspace;
  double* in_block = (double*)p; p += al256(B * 8);
  double* totals_all = (double*)p;
  double* off = totals_all + nba; p += al256((nba + 1) * 8 * 2);
  int32_t* src_rank = (int32_t*)p; p += al256(Bt * 4);
  int32_t* src_local = (int32_t*)p; p += al256(Bt * 4);
  int32_t* send_idx = (int32_t*)p; p += al256(Bt * 4);
  int32_t* slot_pos = (int32_t*)p; p += al256(B * 4);
  int32_t* counts = (int32_t*)p; p += al256(4 * world * 4 + 64);
  int64_t* seg = (int64_t*)p; p += al256(4 * (world + 1) * 8);           // [send offsets (world+1)] [recv offsets (world+1)]
  double* sendbuf = (double*)p; p += al256(Bt * row_doubles * 8);
  double* recvbuf = (double*)p;

  // 1. blocked scan of this rank's weights; the block totals of all ranks (the only weight data that is exchanged)
  g_launch_count += 4;
  k_scan_blocks<<<nb, kScanThreads, 0, st>>>(weights, B, in_block, totals_all + (int64_t)rank * nb);
  if (world > 1) AQ_NCCL_OK(a->AllGather(totals_all + (int64_t)rank * nb, totals_all, nb, ncclDouble, (ncclComm_t)comm, st));
  k_scan_offsets<<<1, 32, 0, st>>>(totals_all, nba, off);
  // 2. every tooth of the global comb: owner rank (+ local index when it is ours); send / receive plans
  k_teeth<<<(unsigned)((Bt + 255) / 256), 256, 0, st>>>(in_block, off, nb, world, rank, B, u, src_rank, src_local, new_weight);
  if (mode == 1) {
    // ---- balanced: every rank keeps its own selected walkers (in tooth order) and only the SURPLUS of ranks that own
    //      more than B teeth travels, to the ranks that own fewer -- the multiset of walkers is that of the global comb,
    //      their order is not.  Walkers are exchangeable, so a DMC run is unaffected; the wire carries the population
    //      imbalance only (the ordered mode moves almost every walker: the comb's base offset u*wtot rotates the teeth
    //      by a fraction u of the whole population).
    int32_t* owned_idx = send_idx;
    {
      const int nch = (int)((Bt + kOwnChunk - 1) / kOwnChunk);            // <= B for any world <= 1024: slot_pos holds them
      int32_t* chunk_off = slot_pos;                                       // (unused by this mode otherwise)
      AQ_CUDA_OK(cudaMemsetAsync(counts, 0, world * sizeof(int32_t), st));
      g_launch_count += 2;
      k_owned_count<<<nch, kOwnChunk, 0, st>>>(src_rank, world, rank, Bt, chunk_off, counts);
      k_scan_i32<<<1, 1024, 0, st>>>(chunk_off, nch);
      k_owned_list<<<nch, kOwnChunk, 0, st>>>(src_rank, src_local, rank, Bt, chunk_off, owned_idx);
    }
    AQ_CUDA_OK(cudaGetLastError());
    int32_t c[64];
    AQ_CUDA_OK(cudaMemcpyAsync(c, counts, world * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    AQ_CUDA_OK(cudaStreamSynchronize(st));
    int64_t tot = 0;
    for (int q = 0; q < world; ++q) tot += c[q];
    if (tot != Bt) return AIQMC_E_CUDA;                       // every tooth has exactly one owner
    const int64_t mine = c[rank], keep = mine < B ? mine : B;
    g_launch_count += 2;
    if (mine > 0) k_gather_rows<<<148 * 4, 256, 0, st>>>(pos, owned_idx, mine, row_doubles, sendbuf);
    if (keep > 0) AQ_CUDA_OK(cudaMemcpyAsync(pos_out, sendbuf, (size_t)keep * row_doubles * 8, cudaMemcpyDeviceToDevice, st));
    if (src_rank_out) k_fill_i32<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(src_rank_out, B, rank);
    // deterministic matching of surplus to deficit, identical on every rank: walk both lists in rank order
    int64_t moved = 0;
    if (world > 1) {
      AQ_NCCL_OK(a->GroupStart());
      int d = 0;
      int64_t dfill = 0;                                       // rows already assigned to deficit rank d
      for (int sidx = 0; sidx < world; ++sidx) {
        int64_t extra = (int64_t)c[sidx] - B, sent = 0;
        while (extra > 0) {
          while (d < world && (int64_t)c[d] + dfill >= B) { ++d; dfill = 0; }
          if (d >= world) return AIQMC_E_CUDA;
          const int64_t room = B - c[d] - dfill, n = extra < room ? extra : room;
          if (sidx == rank) {
            AQ_NCCL_OK(a->Send(sendbuf + (B + sent) * row_doubles, (size_t)n * row_doubles, ncclDouble, d, (ncclComm_t)comm, st));
            moved += n * row_doubles * 8;
          }
          if (d == rank) {
            AQ_NCCL_OK(a->Recv(pos_out + (c[d] + dfill) * row_doubles, (size_t)n * row_doubles, ncclDouble, sidx, (ncclComm_t)comm, st));
            if (src_rank_out) k_fill_i32<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src_rank_out + c[d] + dfill, n, sidx);
          }
          extra -= n; sent += n; dfill += n;
        }
      }
      AQ_NCCL_OK(a->GroupEnd());
    }
    AQ_CUDA_OK(cudaGetLastError());
    if (moved_bytes_out) *moved_bytes_out = moved;
    return AIQMC_OK;
  }
  k_plan<<<2 * world, 1024, 0, st>>>(src_rank, src_local, world, rank, B, send_idx, slot_pos, counts);
  AQ_CUDA_OK(cudaGetLastError());
  int32_t h_counts[2 * 64];
  AQ_CUDA_OK(cudaMemcpyAsync(h_counts, counts, 2 * world * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  AQ_CUDA_OK(cudaStreamSynchronize(st));                    // the message sizes must be known on the host
  int64_t h_seg[2 * 65];
  int64_t so = 0, ro = 0, moved = 0;
  for (int q = 0; q < world; ++q) {
    h_seg[q] = so; so += h_counts[q];
    h_seg[world + 1 + q] = ro; ro += h_counts[world + q];
    if (q != rank) moved += (int64_t)h_counts[q] * row_doubles * 8;
  }
  h_seg[world] = so; h_seg[2 * world + 1] = ro;
  if (ro != B) return AIQMC_E_CUDA;                         // every slot of this rank must be fed exactly once
  AQ_CUDA_OK(cudaMemcpyAsync(seg, h_seg, 2 * (world + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  // 3. pack per destination, exchange only what migrates, unpack in tooth order
  g_launch_count += 2;
  k_pack_rows<<<dim3(148, world), 256, 0, st>>>(pos, send_idx, seg, counts, world, B, row_doubles, sendbuf);
  if (world > 1) {
    AQ_NCCL_OK(a->GroupStart());
    for (int q = 0; q < world; ++q) {
      if (q == rank) continue;
      if (h_counts[q] > 0)
        AQ_NCCL_OK(a->Send(sendbuf + h_seg[q] * row_doubles, (size_t)h_counts[q] * row_doubles, ncclDouble, q, (ncclComm_t)comm, st));
      if (h_counts[world + q] > 0)
        AQ_NCCL_OK(a->Recv(recvbuf + h_seg[world + 1 + q] * row_doubles, (size_t)h_counts[world + q] * row_doubles, ncclDouble, q,
                           (ncclComm_t)comm, st));
    }
    AQ_NCCL_OK(a->GroupEnd());
  }
  if (h_counts[rank] > 0)                                   // rows that stay on this rank: device-to-device, no wire
    AQ_CUDA_OK(cudaMemcpyAsync(recvbuf + h_seg[world + 1 + rank] * row_doubles, sendbuf + h_seg[rank] * row_doubles,
                               (size_t)h_counts[rank] * row_doubles * 8, cudaMemcpyDeviceToDevice, st));
  k_unpack_rows<<<(unsigned)((B * row_doubles + 255) / 256), 256, 0, st>>>(recvbuf, src_rank, slot_pos, seg + world + 1, rank, B,
                                                                           row_doubles, pos_out, src_rank_out);
  AQ_CUDA_OK(cudaGetLastError());
  if (moved_bytes_out) *moved_bytes_out = moved;
  return AIQMC_OK;
}

}  // extern "C"

// Device pieces shared by the posterior kernels (posterior.cu) and the fused pair + posterior kernel (sot_pairs.cu).
#pragma once
#include "sot_common.cuh"

namespace bosot {

constexpr int kPostThreads = 256;
constexpr int kPostWarps = kPostThreads / 32;
constexpr int kPostCand = 64;            // candidates per CTA tile = 8 n-blocks of 8
constexpr int kNB = kPostCand / 8;
constexpr int kRowBlk = 16;              // rows of L^-1 per work unit = 2 m-blocks of 8

struct PostParams {
  const double* Kstar;
  int64_t n_cand;
  int64_t ld;
  int E;
  int ldl;  // padded order of L^-1: multiple of 16
  const double* Linv;  // FRAGMENT-MAJOR: LinvF[(rb * ldl/4 + ks) * 32 + 4 g + t] = Linv[8 rb + g][4 ks + t] (bosot_pack_linv)
  const double* beta;
  double mean_c, variance;
  int acq_kind;
  const double* cur_max;
  double xi, sqrt_ucb_beta;
  double* mu;
  double* sigma;
  double* acq;
  // optional scatter of the acquisition value into the row-sharded result vector of every GPU of the node
  // (peer memory over NVLink / NVSwitch): peers[r][peer_off + candidate] = acq, r = 0 .. n_peers-1
  double* const* peers;
  int n_peers;
  int64_t peer_off;
};

// K* tile row stride (doubles): == 4 or 12 (mod 16) so that the B-fragment read
// Ks[8*nb + lane/4][4*ks + lane%4] hits 16 distinct 8-byte banks per half warp
__host__ __device__ inline int post_tile_ld(int E) {
  const int e4 = (E + 3) & ~3;
  return (e4 % 8 == 4) ? e4 : e4 + 4;
}

__device__ __forceinline__ double norm_cdf(double z) { return 0.5 * erfc(-z * 0.70710678118654752440); }
__device__ __forceinline__ double norm_pdf(double z) { return 0.39894228040143267794 * exp(-0.5 * z * z); }

// D(8x8) += A(8x4, row) * B(4x8, col), fp64.  Lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

}  // namespace bosot

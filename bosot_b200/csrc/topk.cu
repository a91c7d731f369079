// Device-side selection for the evolutionary acquisition optimiser.
//
// Restates EvolutionaryOptimizerObjects.select (bosot/bayesian_optimization/evolutionary_optimizer_objects.py:57-64):
//     sorted_indexes = np.argsort(function_values); survivors = sorted_indexes[n - num_survive : n]
// and the np.argmax of maximize (:42-44) without moving the acquisition vector to the host: bosot_topk_f64 returns the
// last k entries of the ascending order by (value, index) -- i.e. np.argsort(values, kind="stable")[n-k:], NaN above
// +inf like NumPy, -0.0 == +0.0 -- so only k indices (or nothing, when the population stays on the GPU) cross PCIe.
// (NumPy's default sort leaves the order of EQUAL values to the build -- introsort or the AVX-512 sorter --; the stable
// order is the portable choice, and the mirror of the object EA uses it too.)
//
// Two routes, no host synchronisation in either:
//   n <= 4096  : ONE CTA sorts all (key, index) pairs in shared memory (bitonic network) and emits the last k;
//   larger n   : radix select on the order-preserving 64-bit image of the doubles (6 histogram passes of <= 11 bits, the
//                last block of every pass picks the bin that holds the k-th largest), 3 more passes on the index among
//                the values tied with the threshold when only part of them fits, one compaction, a chunk sort in shared
//                memory and log2(k / 2048) merge passes by binary search (the composite keys are unique).
#include <atomic>
#include "sot_common.cuh"
#include "launch.cuh"

namespace bosot {

constexpr int kTopkBins = 2048;
constexpr int kTopkThreads = 512;
constexpr int kTopkChunk = 2048;        // elements sorted per CTA in shared memory (radix route)
constexpr int kTopkSmallMax = 4096;     // single-CTA route (12 bytes of shared memory per element); the bitonic network of one CTA
                                        // takes 290 us at 16384 elements (profiles/r2_ea_breakdown.txt), the radix route about 60

__device__ __forceinline__ unsigned long long order_key(double v) {
  if (v != v) return ~0ULL;  // NaN sorts last (np.argsort)
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  if ((b << 1) == 0ULL) b = 0ULL;  // -0.0 == +0.0
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}

struct TopkState {
  unsigned long long key_prefix;  // digits of the threshold key found so far (right-aligned)
  long long k_rem;                // how many elements of the current bin are still to be taken
  long long ties_total;           // elements whose key equals the threshold
  unsigned int idx_prefix;        // digits of the threshold index found so far
  unsigned int ticket;            // blocks that have finished the current pass
  unsigned int out_count;         // compaction cursor
  unsigned int need_idx;          // 1: only part of the ties fits, the index passes run
  unsigned int hist[kTopkBins];
};

// One histogram pass.  kIdx = false: digit of the key at `shift` among the elements whose higher key bits equal the
// prefix; kIdx = true: digit of the index among the elements whose key equals the threshold (and whose higher index
// bits equal the index prefix).  The last block to finish picks the bin and resets the histogram for the next pass.
template <bool kIdx, bool kFirst>
__global__ void __launch_bounds__(kTopkThreads)
topk_pass_kernel(const double* __restrict__ v, long long n, int shift, int width, TopkState* st) {
  __shared__ unsigned int hist[kTopkBins];
  __shared__ unsigned int wsum[kTopkThreads / 32];
  __shared__ int is_last;
  if (kIdx && st->need_idx == 0u) return;  // uniform: every block sees the value the last key pass left
  for (int b = threadIdx.x; b < kTopkBins; b += kTopkThreads) hist[b] = 0u;
  __syncthreads();
  const unsigned long long kp = st->key_prefix;
  const unsigned int ip = st->idx_prefix;
  const unsigned int mask = (1u << width) - 1u;
  for (long long i = (long long)blockIdx.x * kTopkThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kTopkThreads) {
    const unsigned long long key = order_key(__ldg(&v[i]));
    if (!kIdx) {
      if (kFirst || (key >> (shift + width)) == kp) atomicAdd(&hist[(unsigned int)(key >> shift) & mask], 1u);
    } else {
      const unsigned int idx = (unsigned int)i;
      if (key == kp && (kFirst || (idx >> (shift + width)) == ip)) atomicAdd(&hist[(idx >> shift) & mask], 1u);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kTopkBins; b += kTopkThreads)
    if (hist[b]) atomicAdd(&st->hist[b], hist[b]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // suffix sums from the top bin down: thread t owns bins [4t, 4t + 4) of the REVERSED order
  constexpr int kPer = kTopkBins / kTopkThreads;
  unsigned int c[kPer], own = 0u;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int bin = kTopkBins - 1 - (threadIdx.x * kPer + j);
    c[j] = *((volatile unsigned int*)&st->hist[bin]);
    own += c[j];
  }
  unsigned int inc = own;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int x = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += x;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  unsigned int before = 0u;
  for (int w = 0; w < warp; ++w) before += wsum[w];
  unsigned long long above = (unsigned long long)before + inc - own;  // elements in bins above this thread's first bin
  const unsigned long long k_rem = (unsigned long long)st->k_rem;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int bin = kTopkBins - 1 - (threadIdx.x * kPer + j);
    if (above < k_rem && k_rem <= above + c[j]) {  // exactly one (thread, j) in the block
      if (!kIdx) {
        st->key_prefix = (kp << width) | (unsigned long long)bin;
        if (shift == 0) {
          st->ties_total = (long long)c[j];
          st->need_idx = (k_rem - above) < (unsigned long long)c[j] ? 1u : 0u;
          st->idx_prefix = 0u;
        }
      } else {
        st->idx_prefix = (ip << width) | (unsigned int)bin;
      }
      st->k_rem = (long long)(k_rem - above);
    }
    above += c[j];
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kTopkBins; b += kTopkThreads) st->hist[b] = 0u;
  if (threadIdx.x == 0) st->ticket = 0u;
}

// every element at or above the threshold (key, index) -> unordered list of composite keys
__global__ void __launch_bounds__(kTopkThreads)
topk_compact_kernel(const double* __restrict__ v, long long n, TopkState* st, unsigned long long* __restrict__ out_key,
                    unsigned int* __restrict__ out_idx, long long k) {
  const unsigned long long T = st->key_prefix;
  const unsigned int I = st->idx_prefix;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * kTopkThreads;
  const long long n_round = (n + stride - 1) / stride * stride;  // whole warps stay together for the ballots
  for (long long i = (long long)blockIdx.x * kTopkThreads + threadIdx.x; i < n_round; i += stride) {
    unsigned long long key = 0ULL;
    bool take = false;
    if (i < n) {
      key = order_key(__ldg(&v[i]));
      take = key > T || (key == T && (unsigned int)i >= I);
    }
    const unsigned int m = __ballot_sync(0xffffffffu, take);
    if (m == 0u) continue;
    unsigned int base = 0u;
    if (lane == __ffs(m) - 1) base = atomicAdd(&st->out_count, (unsigned int)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (take) {
      const unsigned int pos = base + (unsigned int)__popc(m & ((1u << lane) - 1u));
      if ((long long)pos < k) { out_key[pos] = key; out_idx[pos] = (unsigned int)i; }
    }
  }
}

__device__ __forceinline__ bool composite_less(unsigned long long ka, unsigned int ia, unsigned long long kb, unsigned int ib) {
  return ka < kb || (ka == kb && ia < ib);
}

// ascending bitonic sort of P (power of two) composite keys in shared memory, all threads of the CTA
__device__ __forceinline__ void bitonic_sort_smem(unsigned long long* key, unsigned int* idx, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int lo = ((t / stride) * (stride << 1)) + (t % stride), hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long ka = key[lo], kb = key[hi];
        const unsigned int ia = idx[lo], ib = idx[hi];
        if (composite_less(kb, ib, ka, ia) == up) { key[lo] = kb; key[hi] = ka; idx[lo] = ib; idx[hi] = ia; }
      }
    }
  }
  __syncthreads();
}

// single-CTA route: sort everything, emit the last k
__global__ void __launch_bounds__(1024)
topk_small_kernel(const double* __restrict__ v, int n, int k, int P, long long index_offset, long long* __restrict__ out_index,
                  double* __restrict__ out_value) {
  extern __shared__ __align__(16) uint8_t smem[];
  unsigned long long* key = reinterpret_cast<unsigned long long*>(smem);
  unsigned int* idx = reinterpret_cast<unsigned int*>(key + P);
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    key[i] = i < n ? order_key(__ldg(&v[i])) : 0ULL;  // padding below every real key (the image of -inf is 2^52 - 1)
    idx[i] = i < n ? (unsigned int)i : 0u;
  }
  bitonic_sort_smem(key, idx, P);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const unsigned int src = idx[P - k + j];
    out_index[j] = (long long)src + index_offset;
    if (out_value) out_value[j] = __ldg(&v[src]);
  }
}

// radix route, step 1 of the ordering: every CTA sorts one chunk of the selected list in shared memory
__global__ void __launch_bounds__(kTopkThreads)
topk_chunk_sort_kernel(unsigned long long* __restrict__ key_g, unsigned int* __restrict__ idx_g, long long k) {
  __shared__ unsigned long long key[kTopkChunk];
  __shared__ unsigned int idx[kTopkChunk];
  const long long c0 = (long long)blockIdx.x * kTopkChunk;
  const int len = (int)min((long long)kTopkChunk, k - c0);
  for (int i = threadIdx.x; i < kTopkChunk; i += kTopkThreads) {
    key[i] = i < len ? key_g[c0 + i] : ~0ULL;  // padding above every real composite (index 2^32 - 1 never occurs: n < 2^31)
    idx[i] = i < len ? idx_g[c0 + i] : ~0u;
  }
  bitonic_sort_smem(key, idx, kTopkChunk);
  for (int i = threadIdx.x; i < len; i += kTopkThreads) { key_g[c0 + i] = key[i]; idx_g[c0 + i] = idx[i]; }
}

// one merge level: runs of `run` sorted elements are merged pairwise; every element finds its rank in the partner run
__global__ void __launch_bounds__(256)
topk_merge_kernel(const unsigned long long* __restrict__ key_in, const unsigned int* __restrict__ idx_in,
                  unsigned long long* __restrict__ key_out, unsigned int* __restrict__ idx_out, long long k, long long run) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  const long long pair0 = (i / (2 * run)) * (2 * run);
  const long long a0 = pair0, a1 = min(pair0 + run, k), b0 = a1, b1 = min(pair0 + 2 * run, k);
  const bool in_a = i < a1;
  const unsigned long long kx = key_in[i];
  const unsigned int ix = idx_in[i];
  long long lo = in_a ? b0 : a0, hi = in_a ? b1 : a1;  // number of partner elements below x (composites are unique)
  const long long base = lo;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (composite_less(key_in[mid], idx_in[mid], kx, ix)) lo = mid + 1; else hi = mid;
  }
  const long long pos = pair0 + (i - (in_a ? a0 : b0)) + (lo - base);
  key_out[pos] = kx;
  idx_out[pos] = ix;
}

__global__ void __launch_bounds__(256)
topk_emit_kernel(const double* __restrict__ v, const unsigned int* __restrict__ idx, long long k, long long index_offset,
                 long long* __restrict__ out_index, double* __restrict__ out_value) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  const unsigned int src = idx[j];
  out_index[j] = (long long)src + index_offset;
  if (out_value) out_value[j] = __ldg(&v[src]);
}

__global__ void topk_init_kernel(TopkState* st, long long k) {
  for (int b = threadIdx.x; b < kTopkBins; b += blockDim.x) st->hist[b] = 0u;
  if (threadIdx.x == 0) {
    st->key_prefix = 0ULL; st->k_rem = k; st->ties_total = 0; st->idx_prefix = 0u; st->ticket = 0u; st->out_count = 0u; st->need_idx = 0u;
  }
}

// np.argmax: first index of the maximum, a NaN counts as the maximum (first NaN wins).  Block-level reduction of
// (key, first index); thread 0 returns the block's winner.
__device__ __forceinline__ bool argmax_better(unsigned long long ka, long long ia, unsigned long long kb, long long ib) {
  return ib < 0 ? true : (ia < 0 ? false : (ka > kb || (ka == kb && ia < ib)));
}
__device__ __forceinline__ void block_argmax(unsigned long long& best, long long& at) {
  __shared__ unsigned long long s_key[32];
  __shared__ long long s_idx[32];
  for (int o = 16; o; o >>= 1) {
    const unsigned long long kb = __shfl_xor_sync(0xffffffffu, best, o);
    const long long ib = __shfl_xor_sync(0xffffffffu, at, o);
    if (!argmax_better(best, at, kb, ib)) { best = kb; at = ib; }
  }
  if ((threadIdx.x & 31) == 0) { s_key[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = at; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = (blockDim.x + 31) >> 5;
    best = (int)threadIdx.x < nw ? s_key[threadIdx.x] : 0ULL;
    at = (int)threadIdx.x < nw ? s_idx[threadIdx.x] : -1;
    for (int o = 16; o; o >>= 1) {
      const unsigned long long kb = __shfl_xor_sync(0xffffffffu, best, o);
      const long long ib = __shfl_xor_sync(0xffffffffu, at, o);
      if (!argmax_better(best, at, kb, ib)) { best = kb; at = ib; }
    }
  }
}

// stage 1: every block the winner of its contiguous slice -> part_key / part_idx[blockIdx.x]
__global__ void __launch_bounds__(1024)
argmax_partial_kernel(const double* __restrict__ v, long long n, unsigned long long* __restrict__ part_key, long long* __restrict__ part_idx) {
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long i0 = (long long)blockIdx.x * per, i1 = min(n, i0 + per);
  unsigned long long best = 0ULL;
  long long at = -1;
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const unsigned long long key = order_key(__ldg(&v[i]));
    if (at < 0 || key > best) { best = key; at = i; }  // ascending i per thread: the first occurrence stays
  }
  block_argmax(best, at);
  if (threadIdx.x == 0) { part_key[blockIdx.x] = best; part_idx[blockIdx.x] = at; }
}

// stage 2 (or the only stage when part_key == nullptr: one block over v itself): out2[0] = value, out2[1] = index
__global__ void __launch_bounds__(1024)
argmax_final_kernel(const double* __restrict__ v, long long n, const unsigned long long* __restrict__ part_key,
                    const long long* __restrict__ part_idx, int n_part, long long index_offset, double* __restrict__ out2) {
  unsigned long long best = 0ULL;
  long long at = -1;
  if (part_key) {
    for (int j = threadIdx.x; j < n_part; j += blockDim.x) {
      const unsigned long long kb = part_key[j];
      const long long ib = part_idx[j];
      if (!argmax_better(best, at, kb, ib)) { best = kb; at = ib; }
    }
  } else {
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned long long key = order_key(__ldg(&v[i]));
      if (at < 0 || key > best) { best = key; at = i; }
    }
  }
  block_argmax(best, at);
  if (threadIdx.x == 0) { out2[0] = __ldg(&v[at]); out2[1] = (double)(at + index_offset); }
}

struct TopkLayout {
  size_t off_state, off_key0, off_key1, off_idx0, off_idx1, total;
};
static TopkLayout topk_layout(int64_t k) {
  TopkLayout L;
  size_t o = 0;
  L.off_state = o;  o = align16(o + sizeof(TopkState));
  L.off_key0 = o;   o = align16(o + sizeof(unsigned long long) * (size_t)k);
  L.off_key1 = o;   o = align16(o + sizeof(unsigned long long) * (size_t)k);
  L.off_idx0 = o;   o = align16(o + sizeof(unsigned int) * (size_t)k);
  L.off_idx1 = o;   o = align16(o + sizeof(unsigned int) * (size_t)k);
  L.total = o;
  return L;
}

}  // namespace bosot

using namespace bosot;

extern "C" size_t bosot_topk_work_bytes(int64_t n, int64_t k) {
  if (n < 1 || k < 1 || k > n || n > 0x7fffffffLL) return 0;
  return topk_layout(k).total;
}

extern "C" int bosot_topk_f64(const double* d_values, int64_t n, int64_t k, int64_t index_offset, int64_t* d_out_index,
                              double* d_out_value, void* d_work, size_t work_bytes, void* stream) {
  if (!d_values || !d_out_index || n < 1 || n > 0x7fffffffLL || k < 1 || k > n) return set_error(BOSOT_EINVAL, "bosot_topk_f64: bad argument (1 <= k <= n < 2^31)");
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= kTopkSmallMax) {
    int P = 32;
    while (P < n) P <<= 1;
    const size_t smem = (size_t)P * 12;
    static std::atomic<bool> attr_set[64];
    int dev = 0;
    if (int rc = cuda_check(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
    if (!(dev < 64 && attr_set[dev])) {
      if (int rc = cuda_check(cudaFuncSetAttribute(topk_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTopkSmallMax * 12),
                              "cudaFuncSetAttribute(topk_small_kernel)")) return rc;
      if (dev < 64) attr_set[dev] = true;
    }
    topk_small_kernel<<<1, P >= 2048 ? 1024 : (P / 2 < 32 ? 32 : P / 2), smem, st>>>(d_values, (int)n, (int)k, P, index_offset,
                                                                                   reinterpret_cast<long long*>(d_out_index), d_out_value);
    return after_launch("topk_small_kernel");
  }
  const TopkLayout L = topk_layout(k);
  if (!d_work || work_bytes < L.total || (reinterpret_cast<uintptr_t>(d_work) & 15))
    return set_error(BOSOT_EINVAL, "bosot_topk_f64: work buffer missing, too small or not 16-byte aligned (bosot_topk_work_bytes)");
  uint8_t* w = static_cast<uint8_t*>(d_work);
  TopkState* state = reinterpret_cast<TopkState*>(w + L.off_state);
  unsigned long long* key[2] = {reinterpret_cast<unsigned long long*>(w + L.off_key0), reinterpret_cast<unsigned long long*>(w + L.off_key1)};
  unsigned int* idx[2] = {reinterpret_cast<unsigned int*>(w + L.off_idx0), reinterpret_cast<unsigned int*>(w + L.off_idx1)};
  int dev = 0, sms = 0;
  if (int rc = cuda_check(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (int rc = cuda_check(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute")) return rc;
  const long long want = (n + kTopkThreads * 4 - 1) / (kTopkThreads * 4);
  const unsigned grid = (unsigned)(want < 2LL * sms ? want : 2LL * sms);

  topk_init_kernel<<<1, 256, 0, st>>>(state, k);
  if (int rc = after_launch("topk_init_kernel")) return rc;
  static const int key_shift[6] = {53, 42, 31, 20, 9, 0}, key_width[6] = {11, 11, 11, 11, 11, 9};
  for (int p = 0; p < 6; ++p) {
    if (p == 0) topk_pass_kernel<false, true><<<grid, kTopkThreads, 0, st>>>(d_values, n, key_shift[p], key_width[p], state);
    else topk_pass_kernel<false, false><<<grid, kTopkThreads, 0, st>>>(d_values, n, key_shift[p], key_width[p], state);
    if (int rc = after_launch("topk_pass_kernel")) return rc;
  }
  static const int idx_shift[3] = {21, 10, 0}, idx_width[3] = {11, 11, 10};
  for (int q = 0; q < 3; ++q) {
    if (q == 0) topk_pass_kernel<true, true><<<grid, kTopkThreads, 0, st>>>(d_values, n, idx_shift[q], idx_width[q], state);
    else topk_pass_kernel<true, false><<<grid, kTopkThreads, 0, st>>>(d_values, n, idx_shift[q], idx_width[q], state);
    if (int rc = after_launch("topk_pass_kernel(index)")) return rc;
  }
  topk_compact_kernel<<<grid, kTopkThreads, 0, st>>>(d_values, n, state, key[0], idx[0], k);
  if (int rc = after_launch("topk_compact_kernel")) return rc;
  const long long chunks = (k + kTopkChunk - 1) / kTopkChunk;
  topk_chunk_sort_kernel<<<(unsigned)chunks, kTopkThreads, 0, st>>>(key[0], idx[0], k);
  if (int rc = after_launch("topk_chunk_sort_kernel")) return rc;
  int cur = 0;
  for (long long run = kTopkChunk; run < k; run <<= 1) {
    topk_merge_kernel<<<(unsigned)((k + 255) / 256), 256, 0, st>>>(key[cur], idx[cur], key[cur ^ 1], idx[cur ^ 1], k, run);
    if (int rc = after_launch("topk_merge_kernel")) return rc;
    cur ^= 1;
  }
  topk_emit_kernel<<<(unsigned)((k + 255) / 256), 256, 0, st>>>(d_values, idx[cur], k, index_offset,
                                                               reinterpret_cast<long long*>(d_out_index), d_out_value);
  return after_launch("topk_emit_kernel");
}

constexpr int kArgmaxMaxParts = 512;
extern "C" size_t bosot_argmax_work_bytes(void) { return (size_t)kArgmaxMaxParts * 16; }

extern "C" int bosot_argmax_f64(const double* d_values, int64_t n, int64_t index_offset, double* d_out2, void* d_work,
                                size_t work_bytes, void* stream) {
  if (!d_values || !d_out2 || n < 1) return set_error(BOSOT_EINVAL, "bosot_argmax_f64: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 65536) {
    argmax_final_kernel<<<1, 1024, 0, st>>>(d_values, n, nullptr, nullptr, 0, index_offset, d_out2);
    return after_launch("argmax_final_kernel");
  }
  if (!d_work || work_bytes < bosot_argmax_work_bytes() || (reinterpret_cast<uintptr_t>(d_work) & 15))
    return set_error(BOSOT_EINVAL, "bosot_argmax_f64: work buffer missing, too small or not 16-byte aligned (bosot_argmax_work_bytes)");
  int dev = 0, sms = 0;
  if (int rc = cuda_check(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (int rc = cuda_check(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute")) return rc;
  int parts = 2 * sms < kArgmaxMaxParts ? 2 * sms : kArgmaxMaxParts;
  unsigned long long* pk = static_cast<unsigned long long*>(d_work);
  long long* pi = reinterpret_cast<long long*>(pk + kArgmaxMaxParts);
  argmax_partial_kernel<<<parts, 1024, 0, st>>>(d_values, n, pk, pi);
  if (int rc = after_launch("argmax_partial_kernel")) return rc;
  argmax_final_kernel<<<1, 1024, 0, st>>>(d_values, n, pk, pi, parts, index_offset, d_out2);
  return after_launch("argmax_final_kernel");
}

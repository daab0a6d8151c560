// rf_select.cu -- result post-processing on the device: the step AFTER the scoring
// scan, so that a one-vs-many call over 10^8 candidates ships k (index, score) pairs over PCIe instead of
// 10^8 scores.
//
//   extract  the k best candidates by (score best-first, index ascending)  -- what Python rapidfuzz calls
//            process.extract; the Rust crate leaves it to the caller's loop + sort
//   filter   every candidate whose score passed score_cutoff (not None), in index order
//
// Both run on the score vector the scan kernels produced (u32 with 0xFFFFFFFF = None, or f64 with NaN = None).
// Scores are mapped to order-preserving u64 keys (smaller = better, None = UINT64_MAX), then
//   1. sel_local_kernel   every CTA: the k smallest keys of its slice, WITH multiplicity
//   2. sel_merge_kernel   one CTA: the k-th smallest key overall = threshold v*
//   3. sel_count_kernel   every CTA: how many keys of its slice are < v* (class A) and == v* (class B)
//   4. sel_scan_kernel    one CTA: exclusive prefix sums of the per-slice counts
//   5. sel_write_kernel   every CTA: ordered compaction -- all of class A, then class B in index order until k
//   6. sel_sort_kernel    one CTA: rank sort of the <= k survivors by (key, index)
// filter is steps 3-5 with v* = "not None".  HBM-bound: the score vector is read three times (4-8 B per
// candidate each), nothing else scales with n.
#include <cstdint>
#include "rf_kernels.cuh"

namespace rfk {

constexpr int SEL_NT = 256;
constexpr int SEL_BATCH = 4096;  // keys examined per extraction batch
constexpr unsigned long long SEL_WORST = 0xFFFFFFFFFFFFFFFFull;

// order-preserving key of a score: smaller key = better candidate
template <bool F64, bool DESC>
__device__ __forceinline__ unsigned long long sel_key(const void* scores, uint64_t i) {
  if constexpr (F64) {
    const unsigned long long b = reinterpret_cast<const unsigned long long*>(scores)[i];
    const double d = __longlong_as_double((long long)b);
    if (d != d) return SEL_WORST;                                         // NaN == None
    unsigned long long k = (b >> 63) ? ~b : (b | 0x8000000000000000ull);  // ascending order of the doubles
    if (DESC) k = ~k;
    return k == SEL_WORST ? SEL_WORST - 1 : k;
  } else {
    const uint32_t s = reinterpret_cast<const uint32_t*>(scores)[i];
    if (s == NONE_U32) return SEL_WORST;
    return DESC ? (unsigned long long)(0xFFFFFFFEu - s) : (unsigned long long)s;  // s <= 0xFFFFFFFE
  }
}

struct SelParams {
  const void* scores;
  uint64_t n;
  uint32_t k;
  unsigned long long* local;        // [parts][k] per-slice smallest keys (ascending, with multiplicity)
  unsigned long long* thresh;       // [1] v*
  unsigned long long* counts;       // [parts][2] -> exclusive offsets after sel_scan_kernel; [2*parts..] totals A, B
  uint32_t* out_idx;                // [cap]
  unsigned long long* out_key;      // [cap]
  uint64_t cap;                     // entries the outputs can hold
  int filter;                       // 1: class A = every key != WORST, no class B
};

__device__ __forceinline__ void sel_slice(const SelParams& p, uint64_t& lo, uint64_t& hi) {
  // contiguous slices, multiples of SEL_NT so that a slice's tiles are aligned
  const uint64_t per = ((p.n + gridDim.x - 1) / gridDim.x + SEL_NT - 1) / SEL_NT * SEL_NT;
  lo = (uint64_t)blockIdx.x * per;
  hi = lo + per < p.n ? lo + per : p.n;
  if (lo > p.n) lo = p.n;
}

// CTA-wide min / count helpers over shared scratch
__device__ __forceinline__ unsigned long long cta_min(unsigned long long v, unsigned long long* wbuf) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, d);
    v = o < v ? o : v;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) wbuf[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned long long g = wbuf[0];
#pragma unroll
  for (int w = 1; w < SEL_NT / 32; ++w) g = wbuf[w] < g ? wbuf[w] : g;
  return g;
}
__device__ __forceinline__ uint32_t cta_sum(uint32_t v, unsigned long long* wbuf) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) wbuf[threadIdx.x >> 5] = v;
  __syncthreads();
  uint32_t g = 0;
#pragma unroll
  for (int w = 0; w < SEL_NT / 32; ++w) g += (uint32_t)wbuf[w];
  return g;
}

// the k smallest of keys[0..m) with multiplicity, ascending, into best[0..k) (padded with WORST)
__device__ void sel_extract(const unsigned long long* keys, uint32_t m, unsigned long long* best, uint32_t k,
                            unsigned long long* wbuf) {
  uint32_t filled = 0;
  unsigned long long last = 0;
  bool first = true;
  while (filled < k) {
    unsigned long long mn = SEL_WORST;
    for (uint32_t i = threadIdx.x; i < m; i += SEL_NT) {
      const unsigned long long v = keys[i];
      if ((first || v > last) && v < mn) mn = v;
    }
    const unsigned long long g = cta_min(mn, wbuf);
    if (g == SEL_WORST) break;
    uint32_t c = 0;
    for (uint32_t i = threadIdx.x; i < m; i += SEL_NT) c += keys[i] == g;
    const uint32_t cnt = cta_sum(c, wbuf);
    const uint32_t take = cnt < k - filled ? cnt : k - filled;
    for (uint32_t i = threadIdx.x; i < take; i += SEL_NT) best[filled + i] = g;
    filled += take;
    last = g;
    first = false;
  }
  for (uint32_t i = filled + threadIdx.x; i < k; i += SEL_NT) best[i] = SEL_WORST;
  __syncthreads();
}

template <bool F64, bool DESC>
__global__ void __launch_bounds__(SEL_NT) sel_local_kernel(const __grid_constant__ SelParams p) {
  extern __shared__ __align__(16) unsigned long long sel_smem[];
  unsigned long long* keys = sel_smem;                 // SEL_BATCH + k
  unsigned long long* best = keys + SEL_BATCH + p.k;   // k
  unsigned long long* tmp = best + p.k;                // k
  unsigned long long* wbuf = tmp + p.k;                // SEL_NT/32
  uint64_t lo, hi;
  sel_slice(p, lo, hi);
  for (uint32_t i = threadIdx.x; i < p.k; i += SEL_NT) best[i] = SEL_WORST;
  __syncthreads();
  for (uint64_t b = lo; b < hi; b += SEL_BATCH) {
    const uint32_t m = (uint32_t)(hi - b < SEL_BATCH ? hi - b : SEL_BATCH);
    // once the running best is full, a batch without a key below its k-th value cannot change it (the usual case)
    const unsigned long long kth = best[p.k - 1];
    uint32_t better = 0;
    for (uint32_t i = threadIdx.x; i < m; i += SEL_NT) {
      const unsigned long long key = sel_key<F64, DESC>(p.scores, b + i);
      keys[i] = key;
      better += key < kth;
    }
    if (cta_sum(better, wbuf) == 0) continue;  // uniform across the CTA
    for (uint32_t i = threadIdx.x; i < p.k; i += SEL_NT) keys[m + i] = best[i];  // merge with the running best
    __syncthreads();
    sel_extract(keys, m + p.k, tmp, p.k, wbuf);
    for (uint32_t i = threadIdx.x; i < p.k; i += SEL_NT) best[i] = tmp[i];
    __syncthreads();
  }
  unsigned long long* out = p.local + (size_t)blockIdx.x * p.k;
  for (uint32_t i = threadIdx.x; i < p.k; i += SEL_NT) out[i] = best[i];
}

__global__ void __launch_bounds__(SEL_NT) sel_merge_kernel(const __grid_constant__ SelParams p, uint32_t parts) {
  extern __shared__ __align__(16) unsigned long long sel_smem[];
  unsigned long long* keys = sel_smem;                       // parts*k
  unsigned long long* best = keys + (size_t)parts * p.k;     // k
  unsigned long long* wbuf = best + p.k;
  const uint32_t m = parts * p.k;
  for (uint32_t i = threadIdx.x; i < m; i += SEL_NT) keys[i] = p.local[i];
  __syncthreads();
  sel_extract(keys, m, best, p.k, wbuf);
  // fewer than k valid candidates: threshold = WORST-1 selects all of them (class A: key < WORST, class B empty)
  if (threadIdx.x == 0) *p.thresh = best[p.k - 1] == SEL_WORST ? SEL_WORST : best[p.k - 1];
}

template <bool F64, bool DESC>
__global__ void __launch_bounds__(SEL_NT) sel_count_kernel(const __grid_constant__ SelParams p) {
  __shared__ unsigned long long wbuf[SEL_NT / 32];
  const unsigned long long v = p.filter ? SEL_WORST : *p.thresh;
  uint64_t lo, hi;
  sel_slice(p, lo, hi);
  uint32_t a = 0, b = 0;
  for (uint64_t i = lo + threadIdx.x; i < hi; i += SEL_NT) {
    const unsigned long long key = sel_key<F64, DESC>(p.scores, i);
    a += key < v;
    b += (key == v) && (v != SEL_WORST);
  }
  const uint32_t ta = cta_sum(a, wbuf);
  const uint32_t tb = cta_sum(b, wbuf);
  if (threadIdx.x == 0) {
    p.counts[2 * blockIdx.x] = ta;
    p.counts[2 * blockIdx.x + 1] = tb;
  }
}

// one CTA: in-place exclusive prefix sums of counts[parts][2]; totals behind them
__global__ void __launch_bounds__(SEL_NT) sel_scan_kernel(unsigned long long* counts, uint32_t parts) {
  __shared__ unsigned long long carry[2];
  __shared__ unsigned long long wsum[2][SEL_NT / 32];
  if (threadIdx.x < 2) carry[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < parts; base += SEL_NT) {
    const uint32_t i = base + threadIdx.x;
    unsigned long long v[2] = {0, 0}, incl[2];
    if (i < parts) { v[0] = counts[2 * i]; v[1] = counts[2 * i + 1]; }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      unsigned long long x = v[c];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= (uint32_t)d) x += t;
      }
      incl[c] = x;
      if (lane == 31) wsum[c][warp] = x;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      unsigned long long off = carry[c];
      for (uint32_t w = 0; w < warp; ++w) off += wsum[c][w];
      if (i < parts) counts[2 * i + c] = off + incl[c] - v[c];
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      unsigned long long t = 0;
      for (int w = 0; w < SEL_NT / 32; ++w) t += wsum[threadIdx.x][w];
      carry[threadIdx.x] += t;
    }
    __syncthreads();
  }
  if (threadIdx.x < 2) counts[2 * (size_t)parts + threadIdx.x] = carry[threadIdx.x];
}

template <bool F64, bool DESC>
__global__ void __launch_bounds__(SEL_NT) sel_write_kernel(const __grid_constant__ SelParams p, uint32_t parts) {
  __shared__ uint32_t wsum[2][SEL_NT / 32];
  const unsigned long long v = p.filter ? SEL_WORST : *p.thresh;
  const unsigned long long total_a = p.counts[2 * (size_t)parts];
  // class A first (all of it), class B behind it; entries beyond `lim` are dropped
  const unsigned long long lim = p.filter ? p.cap : (unsigned long long)p.k;
  unsigned long long pos_a = p.counts[2 * blockIdx.x];
  unsigned long long pos_b = total_a + p.counts[2 * blockIdx.x + 1];
  uint64_t lo, hi;
  sel_slice(p, lo, hi);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint64_t base = lo; base < hi; base += SEL_NT) {
    const uint64_t i = base + threadIdx.x;
    unsigned long long key = SEL_WORST;
    if (i < hi) key = sel_key<F64, DESC>(p.scores, i);
    const bool fa = key < v, fb = (key == v) && (v != SEL_WORST);
    const uint32_t ma = __ballot_sync(0xffffffffu, fa), mb = __ballot_sync(0xffffffffu, fb);
    if (lane == 0) { wsum[0][warp] = __popc(ma); wsum[1][warp] = __popc(mb); }
    __syncthreads();
    uint32_t oa = 0, ob = 0, ta = 0, tb = 0;
#pragma unroll
    for (int w = 0; w < SEL_NT / 32; ++w) {
      if (w < (int)warp) { oa += wsum[0][w]; ob += wsum[1][w]; }
      ta += wsum[0][w];
      tb += wsum[1][w];
    }
    const uint32_t below = (1u << lane) - 1u;
    if (fa) {
      const unsigned long long q = pos_a + oa + __popc(ma & below);
      if (q < lim && q < p.cap) { p.out_idx[q] = (uint32_t)i; if (p.out_key) p.out_key[q] = key; }
    }
    if (fb) {
      const unsigned long long q = pos_b + ob + __popc(mb & below);
      if (q < lim && q < p.cap) { p.out_idx[q] = (uint32_t)i; if (p.out_key) p.out_key[q] = key; }
    }
    pos_a += ta;
    pos_b += tb;
    __syncthreads();
  }
}

// one CTA: sort the m <= 1024 survivors by (key, idx) with a rank sort, write (idx, score)
template <bool F64>
__global__ void __launch_bounds__(1024) sel_sort_kernel(const uint32_t* __restrict__ in_idx,
                                                        const unsigned long long* __restrict__ in_key,
                                                        const unsigned long long* __restrict__ counts, uint32_t parts,
                                                        uint32_t k, const void* __restrict__ scores,
                                                        uint32_t* __restrict__ out_idx, void* __restrict__ out_score,
                                                        uint32_t* __restrict__ out_n) {
  __shared__ unsigned long long sk[1024];
  __shared__ uint32_t si[1024];
  const unsigned long long tot = counts[2 * (size_t)parts] + counts[2 * (size_t)parts + 1];
  const uint32_t m = (uint32_t)(tot < k ? tot : k);
  const uint32_t t = threadIdx.x;
  if (t < m) { sk[t] = in_key[t]; si[t] = in_idx[t]; }
  __syncthreads();
  if (t < m) {
    const unsigned long long key = sk[t];
    const uint32_t idx = si[t];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < m; ++j) rank += (sk[j] < key) || (sk[j] == key && si[j] < idx);
    out_idx[rank] = idx;
    if (F64) reinterpret_cast<double*>(out_score)[rank] = reinterpret_cast<const double*>(scores)[idx];
    else reinterpret_cast<uint32_t*>(out_score)[rank] = reinterpret_cast<const uint32_t*>(scores)[idx];
  }
  if (t == 0) *out_n = m;
}

// filter: copy the compacted indices (already in index order) and gather their scores; the number of valid
// entries min(total, cap) is read from the device-side total
template <bool F64>
__global__ void sel_gather_kernel(const uint32_t* __restrict__ idx, const unsigned long long* __restrict__ total, uint64_t cap,
                                  const void* __restrict__ scores, uint32_t* __restrict__ out_idx, void* __restrict__ out_score) {
  const uint64_t m = *total < cap ? *total : cap;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t c = idx[i];
    out_idx[i] = c;
    if (F64) reinterpret_cast<double*>(out_score)[i] = reinterpret_cast<const double*>(scores)[c];
    else reinterpret_cast<uint32_t*>(out_score)[i] = reinterpret_cast<const uint32_t*>(scores)[c];
  }
}

static uint32_t sel_parts(int sm_count, uint64_t n, uint32_t k) {
  uint64_t parts = (uint64_t)sm_count * 4;
  const uint64_t need = (n + SEL_BATCH - 1) / SEL_BATCH;
  if (parts > need) parts = need;
  const uint64_t fit = 16384 / k;  // the merge CTA holds parts*k keys in shared memory
  if (parts > fit) parts = fit;
  return parts < 1 ? 1u : (uint32_t)parts;
}

template <bool F64, bool DESC>
static cudaError_t select_impl(const SelectLaunch& L) {
  const uint32_t k = L.filter ? 1u : L.k;
  const uint32_t parts = sel_parts(L.sm_count, L.n, k);
  const uint64_t cap = L.filter ? L.cap : (uint64_t)L.k;
  const uint64_t ncap_key = L.filter ? 0 : cap;
  unsigned long long* scratch = nullptr;
  // [local parts*k][thresh 1][counts 2*parts+2][out_key]  + out_idx cap (u32)
  const size_t n64 = (size_t)parts * k + 1 + 2 * (size_t)parts + 2 + ncap_key;
  cudaError_t e = dev_alloc(&scratch, n64 * 8 + cap * 4 + 16, L.stream);
  if (e != cudaSuccess) return e;
  SelParams p{};
  p.scores = L.scores;
  p.n = L.n;
  p.k = k;
  p.local = scratch;
  p.thresh = scratch + (size_t)parts * k;
  p.counts = p.thresh + 1;
  p.out_key = L.filter ? nullptr : p.counts + 2 * (size_t)parts + 2;
  p.out_idx = reinterpret_cast<uint32_t*>(p.counts + 2 * (size_t)parts + 2 + ncap_key);
  p.cap = cap;
  p.filter = L.filter;
  do {
    if (!L.filter) {
      const size_t sm1 = sizeof(unsigned long long) * ((size_t)SEL_BATCH + 3 * k + SEL_NT / 32);
      if (sm1 > 48 * 1024) {
        e = cudaFuncSetAttribute(sel_local_kernel<F64, DESC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
        if (e != cudaSuccess) break;
      }
      sel_local_kernel<F64, DESC><<<parts, SEL_NT, sm1, L.stream>>>(p);
      const size_t sm2 = sizeof(unsigned long long) * ((size_t)parts * k + k + SEL_NT / 32);
      if (sm2 > 48 * 1024) {
        e = cudaFuncSetAttribute(sel_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
        if (e != cudaSuccess) break;
      }
      sel_merge_kernel<<<1, SEL_NT, sm2, L.stream>>>(p, parts);
    }
    sel_count_kernel<F64, DESC><<<parts, SEL_NT, 0, L.stream>>>(p);
    sel_scan_kernel<<<1, SEL_NT, 0, L.stream>>>(p.counts, parts);
    sel_write_kernel<F64, DESC><<<parts, SEL_NT, 0, L.stream>>>(p, parts);
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    if (!L.filter) {
      sel_sort_kernel<F64><<<1, 1024, 0, L.stream>>>(p.out_idx, p.out_key, p.counts, parts, k, L.scores, L.out_idx,
                                                     L.out_score, L.out_n32);
    } else {
      // total hits = class A total; the caller learns it even when it exceeds the capacity
      e = cudaMemcpyAsync(L.out_n64, p.counts + 2 * (size_t)parts, 8, cudaMemcpyDeviceToDevice, L.stream);
      if (e != cudaSuccess) break;
      if (cap) sel_gather_kernel<F64><<<512, 256, 0, L.stream>>>(p.out_idx, p.counts + 2 * (size_t)parts, cap, L.scores,
                                                                 L.out_idx, L.out_score);
    }
    e = cudaGetLastError();
    count_launches(L.filter ? 4 : 6);
  } while (0);
  dev_free(scratch, L.stream);
  return e;
}

// ------------------------------------------------------------------------------------------------ shard merge
// Per-shard top-k lists of a sharded corpus (one part per GPU, gathered with ONE all-gather) -> global top-k.
// Part p holds [nq][k] (shard-local index, distance) rows, 0xFFFFFFFF padded; the global index of an entry is
// base[p] + idx.  Shards are contiguous candidate ranges in part order, so (distance, part, local index) orders
// exactly like (distance, global index): no wide keys are needed.  CTA per query, rank sort in shared memory
// (parts*k is a few hundred entries): every entry counts the entries that precede it; rank < k writes slot rank.
constexpr int MG_NT = 256;
__global__ void __launch_bounds__(MG_NT) topk_merge_kernel(const uint32_t* __restrict__ idx_parts,
                                                           const uint32_t* __restrict__ dist_parts, uint64_t part_stride,
                                                           const unsigned long long* __restrict__ base, uint32_t parts,
                                                           uint32_t k, unsigned long long* __restrict__ out_idx,
                                                           uint32_t* __restrict__ out_dist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* key = reinterpret_cast<unsigned long long*>(smem_raw);  // [parts*k]  dist<<32 | idx, NOKEY = padding
  const uint32_t q = blockIdx.x, m = parts * k;
  for (uint32_t e = threadIdx.x; e < m; e += MG_NT) {
    const uint32_t pt = e / k, i = e - pt * k;
    const size_t src = (size_t)pt * part_stride + (size_t)q * k + i;
    const uint32_t ix = idx_parts[src], d = dist_parts[src];
    key[e] = (ix == 0xFFFFFFFFu) ? ~0ull : (((unsigned long long)d << 32) | ix);
  }
  for (uint32_t i = threadIdx.x; i < k; i += MG_NT) {
    out_idx[(size_t)q * k + i] = ~0ull;
    out_dist[(size_t)q * k + i] = 0xFFFFFFFFu;
  }
  __syncthreads();
  for (uint32_t e = threadIdx.x; e < m; e += MG_NT) {
    const unsigned long long me = key[e];
    if (me == ~0ull) continue;
    const uint32_t md = (uint32_t)(me >> 32);
    const uint32_t pe = e / k;
    uint32_t rank = 0;
    // entry o precedes e: smaller distance, or equal distance and earlier (part, index)
    for (uint32_t po = 0, o = 0; po < parts; ++po) {
      for (uint32_t i = 0; i < k; ++i, ++o) {
        const unsigned long long ot = key[o];
        const uint32_t od = (uint32_t)(ot >> 32);
        rank += (ot != ~0ull) && (od < md || (od == md && (po < pe || (po == pe && ot < me))));
      }
    }
    if (rank < k) {
      out_idx[(size_t)q * k + rank] = base[pe] + (uint32_t)me;
      out_dist[(size_t)q * k + rank] = md;
    }
  }
}

cudaError_t launch_topk_merge(const uint32_t* idx_parts, const uint32_t* dist_parts, uint64_t part_stride,
                              const unsigned long long* base, uint32_t parts, uint32_t nq, uint32_t k,
                              unsigned long long* out_idx, uint32_t* out_dist, cudaStream_t stream) {
  if (parts == 0 || k == 0 || nq == 0) return cudaErrorInvalidValue;
  const size_t smem = sizeof(unsigned long long) * (size_t)parts * k;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  topk_merge_kernel<<<nq, MG_NT, smem, stream>>>(idx_parts, dist_parts, part_stride, base, parts, k, out_idx, out_dist);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_select(const SelectLaunch& L) {
  if (L.n == 0 || (!L.filter && (L.k == 0 || L.k > 1024))) return cudaErrorInvalidValue;
  if (L.f64) return L.desc ? select_impl<true, true>(L) : select_impl<true, false>(L);
  return L.desc ? select_impl<false, true>(L) : select_impl<false, false>(L);
}

}  // namespace rfk

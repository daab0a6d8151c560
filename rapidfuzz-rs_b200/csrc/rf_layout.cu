// rf_layout.cu -- builds the length-bucketed, warp-interleaved corpus layout (LbView) from the CSR corpus.
//
// Why: with one thread per candidate the 32 lanes of a warp must run the same number of Myers steps, and
// every lane's next 4 characters should arrive in one coalesced transaction.  Sorting candidates by length
// inside blocks of LB_BLOCK (so results still scatter into an L2-resident window of the output) and storing
// each group of 32 equal-length candidates word-interleaved gives both, once, at corpus creation.
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <mutex>
#include <vector>
#include "rf_kernels.cuh"

namespace rfk {

constexpr int LB_KEYS = 2048;  // counting-sort keys: exact for lengths < 2047, longer ones share the last bin

__device__ __forceinline__ uint64_t off_ld(const uint32_t* o32, const uint64_t* o64, uint64_t i) {
  return o64 ? o64[i] : (uint64_t)o32[i];
}

// one CTA per block of LB_BLOCK candidates: counting sort by length -> perm / lens in sorted order
__global__ void __launch_bounds__(1024) lb_sort_kernel(const uint32_t* __restrict__ o32, const uint64_t* __restrict__ o64,
                                                       uint64_t n, uint64_t n_pad, uint32_t* __restrict__ perm,
                                                       uint32_t* __restrict__ lens) {
  __shared__ uint32_t hist[LB_KEYS];
  __shared__ uint32_t wsum[32];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint64_t b0 = (uint64_t)blockIdx.x * LB_BLOCK;
  const uint64_t b1 = (b0 + LB_BLOCK < n) ? b0 + LB_BLOCK : n;
  for (uint32_t i = tid; i < LB_KEYS; i += 1024) hist[i] = 0;
  __syncthreads();
  for (uint64_t i = b0 + tid; i < b1; i += 1024) {
    const uint64_t len = off_ld(o32, o64, i + 1) - off_ld(o32, o64, i);
    atomicAdd(&hist[len < LB_KEYS - 1 ? (uint32_t)len : LB_KEYS - 1], 1u);
  }
  __syncthreads();
  {  // exclusive scan of the 2048 bins: 2 bins per thread
    const uint32_t v0 = hist[2 * tid], v1 = hist[2 * tid + 1];
    const uint32_t s = v0 + v1;
    uint32_t incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= (uint32_t)d) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = wsum[lane], wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= (uint32_t)d) wi += t;
      }
      wsum[lane] = wi - w;
    }
    __syncthreads();
    const uint32_t ex = wsum[warp] + incl - s;
    hist[2 * tid] = ex;
    hist[2 * tid + 1] = ex + v0;
  }
  __syncthreads();
  for (uint64_t i = b0 + tid; i < b1; i += 1024) {
    const uint64_t len = off_ld(o32, o64, i + 1) - off_ld(o32, o64, i);
    const uint32_t pos = atomicAdd(&hist[len < LB_KEYS - 1 ? (uint32_t)len : LB_KEYS - 1], 1u);
    perm[b0 + pos] = (uint32_t)i;
    lens[b0 + pos] = (uint32_t)len;
  }
  // padding lanes of the very last group
  const uint64_t e1 = (b0 + LB_BLOCK < n_pad) ? b0 + LB_BLOCK : n_pad;
  for (uint64_t j = b1 + tid; j < e1; j += 1024) {
    perm[j] = 0xFFFFFFFFu;
    lens[j] = 0;
  }
}

// rows (8 bytes per lane) of every group = ceil(max length in group / 8); grows[ngroups] = 0 (scan sentinel)
__global__ void lb_rows_kernel(const uint32_t* __restrict__ lens, uint64_t ngroups, uint32_t* __restrict__ grows) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t total_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t g = warp_global; g <= ngroups; g += total_warps) {
    uint32_t m = 0;
    if (g < ngroups) m = __reduce_max_sync(0xffffffffu, lens[g * 32 + lane]);
    if (lane == 0) grows[g] = (g < ngroups) ? (m + 7u) / 8u : 0u;
  }
}

// transposes each group's candidates into the interleaved rows
__global__ void lb_fill_kernel(const uint8_t* __restrict__ chars, const uint32_t* __restrict__ o32,
                               const uint64_t* __restrict__ o64, const uint32_t* __restrict__ perm,
                               const uint32_t* __restrict__ lens, const uint64_t* __restrict__ goff, uint64_t ngroups,
                               uint32_t* __restrict__ gdata) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t total_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t g = warp_global; g < ngroups; g += total_warps) {
    const uint32_t idx = perm[g * 32 + lane];
    const uint32_t len = lens[g * 32 + lane];
    const uint64_t r0 = goff[g];
    const uint32_t rows = (uint32_t)(goff[g + 1] - r0);
    uint2* dst = reinterpret_cast<uint2*>(gdata) + r0 * 32 + lane;
    if (idx == 0xFFFFFFFFu || len == 0) {
      for (uint32_t k = 0; k < rows; ++k) dst[(size_t)k * 32] = make_uint2(0u, 0u);
      continue;
    }
    const uint64_t o = off_ld(o32, o64, idx);
    ByteReader rd(chars + (o & ~3ull), (uint32_t)(o & 3ull));
    auto word = [&](uint32_t w4) -> uint32_t {  // 4-byte word w4 of the candidate, zero beyond its end
      if (4 * w4 >= len) return 0u;
      uint32_t w = rd.next4();
      const uint32_t left = len - 4 * w4;
      if (left < 4) w &= (1u << (8 * left)) - 1u;
      return w;
    };
    for (uint32_t k = 0; k < rows; ++k) {
      uint2 v;
      v.x = word(2 * k);
      v.y = word(2 * k + 1);
      dst[(size_t)k * 32] = v;
    }
  }
}

// candidate lengths (u8, as they cross PCIe in rf_batch_stream_*_len8) -> CSR starts of a chunk: off[i] = init + sum_{k<i} len[k],
// i = 0..cn (cn + 1 entries)
struct LenAt {
  const uint8_t* lens;
  uint64_t cn;
  __host__ __device__ uint32_t operator()(uint64_t i) const { return i < cn ? (uint32_t)lens[i] : 0u; }
};
size_t lens_to_offsets_tmp_bytes(uint64_t cap_n) {
  size_t bytes = 0;
  auto in = thrust::make_transform_iterator(thrust::make_counting_iterator<uint64_t>(0), LenAt{nullptr, 0});
  cub::DeviceScan::ExclusiveScan(nullptr, bytes, in, (uint32_t*)nullptr, cub::Sum(), 0u, (int64_t)(cap_n + 1), (cudaStream_t)0);
  return bytes;
}
cudaError_t lens_to_offsets(const uint8_t* d_lens, uint64_t cn, uint32_t init, uint32_t* d_off, void* tmp, size_t tmp_bytes,
                            cudaStream_t st) {
  auto in = thrust::make_transform_iterator(thrust::make_counting_iterator<uint64_t>(0), LenAt{d_lens, cn});
  return cub::DeviceScan::ExclusiveScan(tmp, tmp_bytes, in, d_off, cub::Sum(), init, (int64_t)(cn + 1), st);
}

struct CastU64 {
  __host__ __device__ uint64_t operator()(uint32_t v) const { return (uint64_t)v; }
};

// ---- pooled allocation helpers
static std::mutex g_pool_mu;
static std::vector<int> g_pool_ready;
static std::vector<cudaStream_t> g_util_streams;

static void ensure_pool(int dev) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if ((int)g_pool_ready.size() <= dev) g_pool_ready.resize(dev + 1, 0);
  if (g_pool_ready[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t never = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &never);
  }
  cudaGetLastError();
  g_pool_ready[dev] = 1;
}

cudaError_t dev_alloc(void** p, size_t bytes, cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  ensure_pool(dev);
  return cudaMallocAsync(p, bytes ? bytes : 16, st);
}
void dev_free(void* p, cudaStream_t st) {
  if (p) cudaFreeAsync(p, st);
}
cudaStream_t util_stream(int device) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if ((int)g_util_streams.size() <= device) g_util_streams.resize(device + 1, nullptr);
  if (!g_util_streams[device]) {
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    cudaStreamCreateWithFlags(&g_util_streams[device], cudaStreamNonBlocking);
    if (prev >= 0) cudaSetDevice(prev);
  }
  return g_util_streams[device];
}

void lb_free(LbAlloc* a, cudaStream_t st) {
  if (!a) return;
  dev_free(a->perm, st);
  dev_free(a->lens, st);
  dev_free(a->goff, st);
  dev_free(a->gdata, st);
  *a = LbAlloc{};
}

cudaError_t lb_build(const CorpusView& c, cudaStream_t st, LbAlloc* out) {
  *out = LbAlloc{};
  if (c.n == 0) return cudaSuccess;
  const uint64_t n_pad = (c.n + 31) / 32 * 32;
  const uint64_t ngroups = n_pad / 32;
  const uint32_t nblocks = (uint32_t)((c.n + LB_BLOCK - 1) / LB_BLOCK);
  LbAlloc a;
  a.ngroups = ngroups;
  uint32_t* grows = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  cudaError_t e;
#define LB_TRY(x)              \
  do {                         \
    e = (x);                   \
    if (e != cudaSuccess) goto fail; \
  } while (0)
  LB_TRY(dev_alloc(&a.perm, n_pad * sizeof(uint32_t), st));
  LB_TRY(dev_alloc(&a.lens, n_pad * sizeof(uint32_t), st));
  LB_TRY(dev_alloc(&a.goff, (ngroups + 1) * sizeof(uint64_t), st));
  LB_TRY(dev_alloc(&grows, (ngroups + 1) * sizeof(uint32_t), st));
  lb_sort_kernel<<<nblocks, 1024, 0, st>>>(c.off32, c.off64, c.n, n_pad, a.perm, a.lens);
  LB_TRY(cudaGetLastError());
  {
    uint64_t blocks = (ngroups + 1 + 7) / 8;
    if (blocks > 65535 * 8) blocks = 65535 * 8;
    lb_rows_kernel<<<(uint32_t)blocks, 256, 0, st>>>(a.lens, ngroups, grows);
    LB_TRY(cudaGetLastError());
  }
  {
    auto in = thrust::make_transform_iterator(grows, CastU64());
    LB_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, a.goff, (int64_t)(ngroups + 1), st));
    LB_TRY(dev_alloc(&tmp, tmp_bytes, st));
    LB_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, a.goff, (int64_t)(ngroups + 1), st));
  }
  LB_TRY(cudaMemcpyAsync(&a.total_rows, a.goff + ngroups, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  LB_TRY(cudaStreamSynchronize(st));
  // 16 rows of slack: readers load up to 3 rows and prefetch up to 8 rows past the end of the last group
  LB_TRY(dev_alloc(&a.gdata, (a.total_rows + 16) * 256, st));
  LB_TRY(cudaMemsetAsync(a.gdata + a.total_rows * 64, 0, 16 * 256, st));
  {
    uint64_t blocks = (ngroups + 7) / 8;
    if (blocks > 148 * 64) blocks = 148 * 64;
    lb_fill_kernel<<<(uint32_t)blocks, 256, 0, st>>>(c.chars, c.off32, c.off64, a.perm, a.lens, a.goff, ngroups, a.gdata);
    LB_TRY(cudaGetLastError());
  }
  LB_TRY(cudaStreamSynchronize(st));
  dev_free(grows, st);
  dev_free(tmp, st);
  *out = a;
  return cudaSuccess;
fail:
  dev_free(grows, st);
  dev_free(tmp, st);
  lb_free(&a, st);
  return e;
#undef LB_TRY
}

}  // namespace rfk

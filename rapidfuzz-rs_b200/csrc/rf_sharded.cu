// rf_sharded.cu -- one process, several GPUs: the candidate corpus sharded by candidate across the devices of one box
// (SURVEY section 8e), behind the C ABI (include/rfgpu.h, "sharded" section).
//
// The reference's BatchComparator is plain data, Clone + Send + Sync (levenshtein.rs:1635-1639): a Rust host may drive it
// from any thread over any slice of candidates.  Here the split is the library's job: contiguous candidate ranges
// balanced by BYTES, one resident shard per device, the query's tables replicated, per-device streams.  Pairs are
// independent, so the scan itself has no exchange step; the only collectives are the final ones the north star names:
//   * all-gather of the per-shard score vectors when the caller wants the full vector on every device
//     (shards have unequal counts -> one grouped ncclBroadcast per shard = all-gather-v, in place), and
//   * all-gather of the per-shard top-k lists (equal sizes -> ncclAllGather, in place) + rf_topk_merge_device.
// Host-destined results need no collective at all: every device copies its slice straight into the caller's vector.
// Devices listed more than once (tests on a one-GPU box) cannot form an NCCL communicator; the same gathers then run
// as device-to-device copies ordered by events ("sharded_collective" = 1 forces that path everywhere, for A/B timing).
#include <cuda_runtime.h>
#include <nccl.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "rf_internal.h"

using namespace rfk;

static std::atomic<int> g_coll_mode{0};  // 0: NCCL when the devices are distinct, 1: always copies
// scan + all-gather: the shard is scanned in this many pieces, piece k travels while piece k+1 is scanned.  0 = automatic:
// 4 on the copy-engine path (DMA overlaps the scan), 1 with NCCL (its kernels cannot run beside the persistent scan kernel)
static std::atomic<int> g_gather_chunks{0};
void rf__set_sharded_collective(int mode) { g_coll_mode.store(mode < 0 || mode > 2 ? 0 : mode); }
void rf__set_gather_chunks(int k) { g_gather_chunks.store(k < 0 ? 0 : (k > 16 ? 16 : k)); }
constexpr int kMaxChunks = 16;

struct rf_sharded_corpus {
  std::vector<int> devices;
  std::vector<rf_corpus*> shard;
  std::vector<uint64_t> lo;  // [ndev + 1] first candidate of every shard
  uint64_t n = 0, total = 0;
  bool distinct = true;
  bool peer_dma = true;               // every pair of devices has peer access enabled (NVLink DMA)
  std::vector<cudaStream_t> streams;  // one per shard, on its device
  std::vector<cudaStream_t> comm_streams;  // high priority: the gathers overlap the next piece's scan
  std::vector<cudaEvent_t> events;
  std::vector<cudaEvent_t> chunk_events;   // [ndev][kMaxChunks + 1]
  std::mutex coll_mu;                 // collectives on one set of communicators are issued by one thread at a time
  bool comm_ready = false;
  std::vector<ncclComm_t> comms;
};

struct rf_sharded_batch {
  std::vector<int> devices;
  std::vector<rf_batch*> per;
  rf_metric metric = RF_LEVENSHTEIN;
};

namespace {

struct DevGuard {
  int prev = -1;
  explicit DevGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    cudaSetDevice(dev);
  }
  ~DevGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

rf_status nccl_fail(ncclResult_t r, const char* what) {
  return rfi::fail(RF_ERR_NCCL, std::string(what) + ": " + ncclGetErrorString(r));
}

// runs fn(i) for every shard on its own host thread (the per-shard entry points synchronise internally); the first
// failure's status and message are re-raised on the calling thread
template <class Fn>
rf_status for_each_shard(size_t count, Fn fn) {
  std::vector<rf_status> st(count, RF_OK);
  std::vector<std::string> msg(count);
  if (count == 1) {
    return fn(0);
  }
  std::vector<std::thread> th;
  th.reserve(count);
  for (size_t i = 0; i < count; ++i) {
    try {
      th.emplace_back([&, i] {
        st[i] = fn(i);
        if (st[i] != RF_OK) msg[i] = rfi::last_error();
      });
    } catch (...) {  // no thread to be had (nothing may propagate across the C boundary): this shard runs on the caller's
      st[i] = fn(i);
      if (st[i] != RF_OK) msg[i] = rfi::last_error();
    }
  }
  for (auto& t : th) t.join();
  for (size_t i = 0; i < count; ++i)
    if (st[i] != RF_OK) return rfi::fail(st[i], "shard " + std::to_string(i) + ": " + msg[i]);
  return RF_OK;
}

bool use_nccl(const rf_sharded_corpus* c) { return c->distinct && g_coll_mode.load() == 0; }
// the score-vector gather of the one-process form: copy engines over NVLink whenever every pair has peer access (the DMA
// overlaps the scan), NCCL otherwise; "sharded_collective" = 2 forces NCCL, 1 forces copies
bool gather_with_nccl(const rf_sharded_corpus* c) {
  const int mode = g_coll_mode.load();
  if (!c->distinct || mode == 1) return false;
  if (mode == 2) return true;
  return !c->peer_dma;
}

rf_status ensure_comms(rf_sharded_corpus* c) {
  if (c->comm_ready) return RF_OK;
  c->comms.assign(c->devices.size(), nullptr);
  ncclResult_t r = ncclCommInitAll(c->comms.data(), (int)c->devices.size(), c->devices.data());
  if (r != ncclSuccess) {
    c->comms.clear();
    return nccl_fail(r, "ncclCommInitAll");
  }
  c->comm_ready = true;
  return RF_OK;
}

// In-place all-gather-v over the shards' streams: bufs[i] is device i's copy of the whole buffer; the bytes
// [off[r], off[r+1]) are valid on device r and land on every device.  Work already enqueued on streams[r] (the scan that
// produced shard r's slice) is ordered before the transfer.
rf_status allgatherv_inplace(rf_sharded_corpus* c, void* const* bufs, const uint64_t* off) {
  const int nd = (int)c->devices.size();
  if (nd == 1) return RF_OK;
  std::lock_guard<std::mutex> lk(c->coll_mu);
  if (use_nccl(c)) {
    rf_status s = ensure_comms(c);
    if (s != RF_OK) return s;
    ncclResult_t r = ncclGroupStart();
    if (r != ncclSuccess) return nccl_fail(r, "ncclGroupStart");
    for (int root = 0; root < nd && r == ncclSuccess; ++root) {
      const uint64_t bytes = off[root + 1] - off[root];
      if (!bytes) continue;
      for (int i = 0; i < nd && r == ncclSuccess; ++i) {
        char* p = (char*)bufs[i] + off[root];
        r = ncclBroadcast(p, p, bytes, ncclUint8, root, c->comms[i], c->streams[i]);
      }
    }
    ncclResult_t r2 = ncclGroupEnd();
    if (r != ncclSuccess) return nccl_fail(r, "ncclBroadcast");
    if (r2 != ncclSuccess) return nccl_fail(r2, "ncclGroupEnd");
    return RF_OK;
  }
  // copies: device i pulls every other shard's slice once that shard's stream has produced it
  for (int r = 0; r < nd; ++r) {
    DevGuard g(c->devices[r]);
    cudaError_t e = cudaEventRecord(c->events[r], c->streams[r]);
    if (e != cudaSuccess) return rfi::cuda_fail(e, "cudaEventRecord");
  }
  for (int i = 0; i < nd; ++i) {
    DevGuard g(c->devices[i]);
    for (int r = 0; r < nd; ++r) {
      if (r == i) continue;
      const uint64_t bytes = off[r + 1] - off[r];
      if (!bytes) continue;
      cudaError_t e = cudaStreamWaitEvent(c->streams[i], c->events[r], 0);
      if (e == cudaSuccess)
        e = cudaMemcpyPeerAsync((char*)bufs[i] + off[r], c->devices[i], (const char*)bufs[r] + off[r], c->devices[r], bytes,
                                c->streams[i]);
      if (e != cudaSuccess) return rfi::cuda_fail(e, "peer copy");
    }
  }
  return RF_OK;
}

rf_status sync_all(const rf_sharded_corpus* c) {
  rf_status s = RF_OK;
  for (size_t i = 0; i < c->devices.size(); ++i) {
    DevGuard g(c->devices[i]);
    cudaError_t e = cudaStreamSynchronize(c->streams[i]);
    if (e != cudaSuccess && s == RF_OK) s = rfi::cuda_fail(e, "sharded stream synchronize");
  }
  return s;
}

rf_status check_pair(const rf_sharded_batch* b, const rf_sharded_corpus* c) {
  if (!b || !c) return rfi::fail(RF_ERR_INVALID_ARG, "NULL handle");
  if (b->devices != c->devices) return rfi::fail(RF_ERR_INVALID_ARG, "sharded batch and sharded corpus were made for different device lists");
  return RF_OK;
}

// byte-balanced contiguous ranges: boundary s = first candidate starting at or after total * s / parts
void split_by_bytes(const uint64_t* offsets, uint64_t n, size_t parts, std::vector<uint64_t>* lo) {
  lo->assign(parts + 1, 0);
  const uint64_t base = offsets[0], total = offsets[n] - base;
  for (size_t s = 1; s < parts; ++s) {
    uint64_t b;
    if (total == 0) b = n * s / parts;
    else b = (uint64_t)(std::lower_bound(offsets, offsets + n, base + (uint64_t)((unsigned __int128)total * s / parts)) - offsets);
    (*lo)[s] = std::max(b, (*lo)[s - 1]);
  }
  (*lo)[parts] = n;
}


// ---- scan + all-gather with the transfer overlapped (both the one-process and the process-per-GPU form use this).
// The local shard is scanned in K pieces of whole 65536-candidate blocks on the compute stream; as soon as piece k is
// done (event) the high-priority communication stream broadcasts every rank's piece k in place (grouped ncclBroadcast =
// all-gather-v) while piece k+1 is being scanned.  The compute stream finally waits for the last transfer.
struct GatherLocal {
  const rf_batch* b;
  const rf_corpus* c;
  int rank, device;
  ncclComm_t comm;
  cudaStream_t compute, comm_stream;
  cudaEvent_t* ev;  // [kMaxChunks + 1]
  void* out;        // this device's buffer for ALL ranks' results
};

bool chunkable(const rf_batch* b, const rf_corpus* c, const rf_args* a) {
  if (b->wide || c->d_elems32 || c->compact32) return false;
  switch (b->metric) {
    case RF_LEVENSHTEIN:
      if (!a) return true;
      return a->insertion_cost == a->deletion_cost &&
             (a->insertion_cost == a->substitution_cost || a->substitution_cost >= a->insertion_cost + a->deletion_cost);
    case RF_INDEL: case RF_LCS_SEQ: case RF_OSA: case RF_JARO: case RF_JARO_WINKLER: case RF_RATIO: return true;
    default: return false;
  }
}

void piece_range(uint64_t n, int K, int k, uint64_t* a, uint64_t* z) {
  const uint64_t per = ((n + K - 1) / K + LB_BLOCK - 1) / LB_BLOCK * LB_BLOCK;
  *a = std::min<uint64_t>(n, (uint64_t)k * per);
  *z = std::min<uint64_t>(n, (uint64_t)(k + 1) * per);
}

// The same with the copy engines instead of NCCL kernels (one process, peer access): as soon as a shard's compute stream
// has produced piece k, that device's communication stream PUSHES it into every peer's buffer over NVLink.  DMA transfers
// need no SM, so they overlap the persistent scan kernel of piece k+1 completely -- NCCL's broadcast kernels cannot: the
// scan occupies every CTA slot, and the transfer of piece k only starts when piece k+1 retires (measured: no gain).
rf_status scan_allgather_copies(rf_sharded_corpus* c, const rf_sharded_batch* b, rf_kind kind, const rf_args* args,
                                void* const* out_device, bool want_f64) {
  const int nd = (int)c->devices.size();
  const size_t esz = want_f64 ? 8 : 4;
  int K = g_gather_chunks.load();
  if (K == 0) K = 4;
  for (int i = 0; i < nd; ++i)
    if (!chunkable(b->per[i], c->shard[i], args)) K = 1;
  rf_status s = RF_OK;
  auto ev = [&](int dev_i, int k) { return c->chunk_events[(size_t)dev_i * (kMaxChunks + 1) + k]; };
  for (int k = 0; k < K && s == RF_OK; ++k) {
    for (int i = 0; i < nd && s == RF_OK; ++i) {
      DevGuard dg(c->devices[i]);
      uint64_t a, z;
      piece_range(c->lo[i + 1] - c->lo[i], K, k, &a, &z);
      void* mine = (uint8_t*)out_device[i] + c->lo[i] * esz;
      if (z > a)
        s = (K == 1) ? rfi::score_device(b->per[i], c->shard[i], kind, args, mine, want_f64, c->streams[i], nullptr)
                     : rfi::score_device_range(b->per[i], c->shard[i], kind, args, mine, want_f64, c->streams[i], a, z);
      if (s != RF_OK) break;
      const cudaError_t e = cudaEventRecord(ev(i, k), c->streams[i]);
      if (e != cudaSuccess) s = rfi::cuda_fail(e, "cudaEventRecord");
    }
    // PUSH: the source's copy engine writes its piece into every peer's buffer (NVLink writes are posted; remote reads
    // pay a round trip per request -- pulling measured 2.2x slower at 8 GPUs: 6.06 vs 2.81 ms, profiles/r2_sharded_abi_one_process_n8*.json).  Staggered targets: at any moment every
    // destination receives from a different source.
    for (int r = 0; r < nd && s == RF_OK; ++r) {
      DevGuard dg(c->devices[r]);
      uint64_t a, z;
      piece_range(c->lo[r + 1] - c->lo[r], K, k, &a, &z);
      if (z == a) continue;
      const uint64_t off = (c->lo[r] + a) * esz;
      cudaError_t e = cudaStreamWaitEvent(c->comm_streams[r], ev(r, k), 0);
      for (int d = 1; d < nd && e == cudaSuccess; ++d) {
        const int i = (r + d) % nd;
        e = cudaMemcpyPeerAsync((char*)out_device[i] + off, c->devices[i], (const char*)out_device[r] + off, c->devices[r],
                                (z - a) * esz, c->comm_streams[r]);
      }
      if (e != cudaSuccess) s = rfi::cuda_fail(e, "peer copy");
    }
  }
  // every device's compute stream continues once ALL sources have delivered
  for (int r = 0; r < nd; ++r) {
    DevGuard dg(c->devices[r]);
    const cudaError_t e = cudaEventRecord(ev(r, kMaxChunks), c->comm_streams[r]);
    if (e != cudaSuccess && s == RF_OK) s = rfi::cuda_fail(e, "cudaEventRecord");
  }
  for (int i = 0; i < nd; ++i) {
    DevGuard dg(c->devices[i]);
    for (int r = 0; r < nd; ++r) {
      const cudaError_t e = cudaStreamWaitEvent(c->streams[i], ev(r, kMaxChunks), 0);
      if (e != cudaSuccess && s == RF_OK) s = rfi::cuda_fail(e, "scan/gather ordering");
    }
  }
  return s;
}

rf_status scan_allgather_nccl(GatherLocal* loc, int nloc, int nranks, const uint64_t* counts, rf_kind kind, const rf_args* args,
                              bool want_f64) {
  const size_t esz = want_f64 ? 8 : 4;
  std::vector<uint64_t> lo(nranks + 1, 0);
  for (int r = 0; r < nranks; ++r) lo[r + 1] = lo[r] + counts[r];
  int K = g_gather_chunks.load();
  if (K == 0) K = 1;
  for (int i = 0; i < nloc; ++i)
    if (!chunkable(loc[i].b, loc[i].c, args)) K = 1;
  if (nranks == 1) K = 1;
  rf_status s = RF_OK;
  for (int k = 0; k < K && s == RF_OK; ++k) {
    for (int i = 0; i < nloc && s == RF_OK; ++i) {
      GatherLocal& g = loc[i];
      DevGuard dg(g.device);
      uint64_t a, z;
      piece_range(counts[g.rank], K, k, &a, &z);
      void* mine = (uint8_t*)g.out + lo[g.rank] * esz;
      if (z > a) {
        s = (K == 1) ? rfi::score_device(g.b, g.c, kind, args, mine, want_f64, g.compute, nullptr)
                     : rfi::score_device_range(g.b, g.c, kind, args, mine, want_f64, g.compute, a, z);
        if (s != RF_OK) break;
      }
      cudaError_t e = cudaEventRecord(g.ev[k], g.compute);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(g.comm_stream, g.ev[k], 0);
      if (e != cudaSuccess) s = rfi::cuda_fail(e, "scan/gather ordering");
    }
    if (s != RF_OK || nranks == 1) continue;
    ncclResult_t r = ncclGroupStart();
    if (r != ncclSuccess) return nccl_fail(r, "ncclGroupStart");
    for (int i = 0; i < nloc && r == ncclSuccess; ++i) {
      GatherLocal& g = loc[i];
      for (int root = 0; root < nranks && r == ncclSuccess; ++root) {
        uint64_t a, z;
        piece_range(counts[root], K, k, &a, &z);
        if (z == a) continue;
        char* p = (char*)g.out + (lo[root] + a) * esz;
        r = ncclBroadcast(p, p, (z - a) * esz, ncclUint8, root, g.comm, g.comm_stream);
      }
    }
    ncclResult_t r2 = ncclGroupEnd();
    if (r != ncclSuccess) return nccl_fail(r, "ncclBroadcast");
    if (r2 != ncclSuccess) return nccl_fail(r2, "ncclGroupEnd");
  }
  for (int i = 0; i < nloc; ++i) {  // later work on the compute stream sees the gathered vector
    GatherLocal& g = loc[i];
    DevGuard dg(g.device);
    cudaError_t e = cudaEventRecord(g.ev[kMaxChunks], g.comm_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(g.compute, g.ev[kMaxChunks], 0);
    if (e != cudaSuccess && s == RF_OK) s = rfi::cuda_fail(e, "scan/gather ordering");
  }
  return s;
}
}  // namespace

extern "C" {

rf_status rf_corpus_create_sharded_u8(const uint8_t* chars, const uint64_t* offsets, uint64_t n, const int* devices, int ndev,
                                      rf_sharded_corpus** out) {
  if (!out) return rfi::fail(RF_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (!offsets) return rfi::fail(RF_ERR_INVALID_ARG, "offsets is NULL");
  if (!devices || ndev < 1 || ndev > 64) return rfi::fail(RF_ERR_INVALID_ARG, "devices / ndev (1..64)");
  if (offsets[0] != 0) return rfi::fail(RF_ERR_INVALID_ARG, "offsets[0] must be 0");
  const int have = rf_device_count();
  for (int i = 0; i < ndev; ++i)
    if (devices[i] < 0 || devices[i] >= have) return rfi::fail(RF_ERR_CUDA, "no such CUDA device");
  rf_sharded_corpus* c = new (std::nothrow) rf_sharded_corpus();
  if (!c) return rfi::fail(RF_ERR_OOM, "host allocation failed");
  c->devices.assign(devices, devices + ndev);
  c->n = n;
  c->total = offsets[n];
  for (int i = 0; i < ndev; ++i)
    for (int j = 0; j < i; ++j)
      if (devices[i] == devices[j]) c->distinct = false;
  split_by_bytes(offsets, n, (size_t)ndev, &c->lo);
  c->shard.assign(ndev, nullptr);
  c->streams.assign(ndev, nullptr);
  c->comm_streams.assign(ndev, nullptr);
  c->events.assign(ndev, nullptr);
  c->chunk_events.assign((size_t)ndev * (kMaxChunks + 1), nullptr);
  rf_status s = RF_OK;
  for (int i = 0; i < ndev && s == RF_OK; ++i) {
    DevGuard g(devices[i]);
    int lo_pri = 0, hi_pri = 0;
    cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
    cudaError_t e = cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->comm_streams[i], cudaStreamNonBlocking, hi_pri);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->events[i], cudaEventDisableTiming);
    for (int k = 0; k <= kMaxChunks && e == cudaSuccess; ++k)
      e = cudaEventCreateWithFlags(&c->chunk_events[(size_t)i * (kMaxChunks + 1) + k], cudaEventDisableTiming);
    if (e != cudaSuccess) s = rfi::cuda_fail(e, "sharded corpus streams");
  }
  // direct NVLink DMA between the shards' devices (without it cudaMemcpyPeerAsync stages through the host)
  for (int i = 0; i < ndev && s == RF_OK; ++i) {
    DevGuard g(devices[i]);
    for (int j = 0; j < ndev; ++j) {
      if (devices[j] == devices[i]) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) == cudaSuccess && can) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) c->peer_dma = false;
      } else {
        c->peer_dma = false;
      }
      cudaGetLastError();
    }
  }
  // every shard uploads and builds its layout on its own device concurrently (one host thread per shard)
  if (s == RF_OK)
    s = for_each_shard((size_t)ndev, [&](size_t i) {
      return rfi::corpus_create_sub(chars, offsets, c->lo[i], c->lo[i + 1], c->devices[i], &c->shard[i]);
    });
  if (s != RF_OK) {
    const std::string keep = rfi::last_error();
    rf_sharded_corpus_destroy(c);
    return rfi::fail(s, keep);
  }
  *out = c;
  return RF_OK;
}

rf_status rf_sharded_corpus_destroy(rf_sharded_corpus* c) {
  if (!c) return RF_OK;
  for (rf_corpus* s : c->shard) rf_corpus_destroy(s);
  if (c->comm_ready)
    for (ncclComm_t cm : c->comms)
      if (cm) ncclCommDestroy(cm);
  for (size_t i = 0; i < c->devices.size(); ++i) {
    DevGuard g(c->devices[i]);
    if (c->events[i]) cudaEventDestroy(c->events[i]);
    for (int k = 0; k <= kMaxChunks; ++k)
      if (c->chunk_events[(size_t)i * (kMaxChunks + 1) + k]) cudaEventDestroy(c->chunk_events[(size_t)i * (kMaxChunks + 1) + k]);
    if (c->comm_streams[i]) cudaStreamDestroy(c->comm_streams[i]);
    if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
  }
  delete c;
  return RF_OK;
}

uint64_t rf_sharded_corpus_size(const rf_sharded_corpus* c) { return c ? c->n : 0; }
int rf_sharded_corpus_shards(const rf_sharded_corpus* c) { return c ? (int)c->devices.size() : 0; }
rf_status rf_sharded_corpus_shard_range(const rf_sharded_corpus* c, int shard, uint64_t* first, uint64_t* end) {
  if (!c || shard < 0 || shard >= (int)c->devices.size()) return rfi::fail(RF_ERR_INVALID_ARG, "no such shard");
  if (first) *first = c->lo[shard];
  if (end) *end = c->lo[shard + 1];
  return RF_OK;
}
const rf_corpus* rf_sharded_corpus_shard(const rf_sharded_corpus* c, int shard) {
  return (c && shard >= 0 && shard < (int)c->devices.size()) ? c->shard[shard] : nullptr;
}
int rf_sharded_corpus_uses_nccl(const rf_sharded_corpus* c) { return (c && c->devices.size() > 1 && use_nccl(c)) ? 1 : 0; }

static rf_status sharded_batch_create(rf_metric metric, const void* query, uint32_t query_len, bool wide, const int* devices, int ndev,
                                      rf_sharded_batch** out) {
  if (!out) return rfi::fail(RF_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (!devices || ndev < 1 || ndev > 64) return rfi::fail(RF_ERR_INVALID_ARG, "devices / ndev (1..64)");
  rf_sharded_batch* b = new (std::nothrow) rf_sharded_batch();
  if (!b) return rfi::fail(RF_ERR_OOM, "host allocation failed");
  b->devices.assign(devices, devices + ndev);
  b->metric = metric;
  b->per.assign(ndev, nullptr);
  for (int i = 0; i < ndev; ++i) {
    rf_status s = wide ? rf_batch_create_u32(metric, (const uint32_t*)query, query_len, devices[i], &b->per[i])
                       : rf_batch_create_u8(metric, (const uint8_t*)query, query_len, devices[i], &b->per[i]);
    if (s != RF_OK) {
      const std::string keep = rfi::last_error();
      rf_sharded_batch_destroy(b);
      return rfi::fail(s, keep);
    }
  }
  *out = b;
  return RF_OK;
}

rf_status rf_sharded_batch_create_u8(rf_metric metric, const uint8_t* query, uint32_t query_len, const int* devices, int ndev,
                                     rf_sharded_batch** out) {
  return sharded_batch_create(metric, query, query_len, false, devices, ndev, out);
}
rf_status rf_sharded_batch_create_u32(rf_metric metric, const uint32_t* query, uint32_t query_len, const int* devices, int ndev,
                                      rf_sharded_batch** out) {
  return sharded_batch_create(metric, query, query_len, true, devices, ndev, out);
}
rf_status rf_sharded_batch_destroy(rf_sharded_batch* b) {
  if (!b) return RF_OK;
  for (rf_batch* p : b->per) rf_batch_destroy(p);
  delete b;
  return RF_OK;
}

// ---- scoring into the caller's HOST vector: no collective, every device downloads its own slice
static rf_status sharded_score_host(const rf_sharded_batch* b, const rf_sharded_corpus* cc, rf_kind kind, const rf_args* args,
                                    void* out_host, bool want_f64) {
  rf_status s = check_pair(b, cc);
  if (s != RF_OK) return s;
  rf_sharded_corpus* c = const_cast<rf_sharded_corpus*>(cc);
  if (c->n && !out_host) return rfi::fail(RF_ERR_INVALID_ARG, "out is NULL");
  const size_t nd = c->devices.size(), esz = want_f64 ? 8 : 4;
  std::vector<uint8_t*> d_out(nd, nullptr);
  std::vector<uint32_t> differing(nd, 0);
  for (size_t i = 0; i < nd && s == RF_OK; ++i) {
    const uint64_t cnt = c->lo[i + 1] - c->lo[i];
    DevGuard g(c->devices[i]);
    cudaStream_t st = c->streams[i];
    const size_t bytes = (size_t)cnt * esz;
    cudaError_t e = dev_alloc(&d_out[i], bytes + 16, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_out[i] + bytes, 0, 16, st);
    if (e != cudaSuccess) { s = rfi::cuda_fail(e, "result buffer"); break; }
    s = rfi::score_device(b->per[i], c->shard[i], kind, args, cnt ? d_out[i] : nullptr, want_f64, st, (uint32_t*)(d_out[i] + bytes));
    if (s != RF_OK) break;
    if (cnt) e = cudaMemcpyAsync((uint8_t*)out_host + c->lo[i] * esz, d_out[i], bytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&differing[i], d_out[i] + bytes, 4, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) s = rfi::cuda_fail(e, "result download");
  }
  const std::string keep = s != RF_OK ? rfi::last_error() : std::string();
  rf_status s2 = sync_all(c);
  for (size_t i = 0; i < nd; ++i) {
    DevGuard g(c->devices[i]);
    dev_free(d_out[i], c->streams[i]);
  }
  if (s != RF_OK) return rfi::fail(s, keep);
  if (s2 != RF_OK) return s2;
  for (size_t i = 0; i < nd; ++i)
    if (differing[i]) return rfi::fail(RF_ERR_INVALID_ARG, "Differing length arguments provided");  // hamming::Error
  return RF_OK;
}

rf_status rf_sharded_score_u32(const rf_sharded_batch* b, const rf_sharded_corpus* c, rf_kind kind, const rf_args* args,
                               uint32_t* out_host) {
  return sharded_score_host(b, c, kind, args, out_host, false);
}
rf_status rf_sharded_score_f64(const rf_sharded_batch* b, const rf_sharded_corpus* c, rf_kind kind, const rf_args* args,
                               double* out_host) {
  return sharded_score_host(b, c, kind, args, out_host, true);
}

// ---- scoring + all-gather: out_device[i] (on devices[i]) receives ALL n scores
static rf_status sharded_score_allgather(const rf_sharded_batch* b, const rf_sharded_corpus* cc, rf_kind kind, const rf_args* args,
                                         void* const* out_device, bool want_f64) {
  rf_status s = check_pair(b, cc);
  if (s != RF_OK) return s;
  rf_sharded_corpus* c = const_cast<rf_sharded_corpus*>(cc);
  if (c->n == 0) return RF_OK;
  if (!out_device) return rfi::fail(RF_ERR_INVALID_ARG, "out_device is NULL");
  const size_t nd = c->devices.size(), esz = want_f64 ? 8 : 4;
  std::vector<uint64_t> off(nd + 1);
  for (size_t i = 0; i <= nd; ++i) off[i] = c->lo[i] * esz;
  for (size_t i = 0; i < nd; ++i)
    if (!out_device[i]) return rfi::fail(RF_ERR_INVALID_ARG, "out_device[i] is NULL");
  if (nd > 1 && gather_with_nccl(c)) {
    std::lock_guard<std::mutex> lk(c->coll_mu);
    s = ensure_comms(c);
    if (s != RF_OK) return s;
    std::vector<GatherLocal> loc(nd);
    std::vector<uint64_t> counts(nd);
    for (size_t i = 0; i < nd; ++i) {
      counts[i] = c->lo[i + 1] - c->lo[i];
      loc[i] = GatherLocal{b->per[i], c->shard[i], (int)i, c->devices[i], c->comms[i], c->streams[i], c->comm_streams[i],
                           &c->chunk_events[i * (kMaxChunks + 1)], out_device[i]};
    }
    s = scan_allgather_nccl(loc.data(), (int)nd, (int)nd, counts.data(), kind, args, want_f64);
    const std::string keep = s != RF_OK ? rfi::last_error() : std::string();
    rf_status s2 = sync_all(c);
    if (s != RF_OK) return rfi::fail(s, keep);
    return s2;
  }
  {
    std::lock_guard<std::mutex> lk(c->coll_mu);
    s = scan_allgather_copies(c, b, kind, args, out_device, want_f64);
  }
  const std::string keep = s != RF_OK ? rfi::last_error() : std::string();
  rf_status s2 = sync_all(c);
  if (s != RF_OK) return rfi::fail(s, keep);
  return s2;
}

rf_status rf_sharded_score_u32_allgather_device(const rf_sharded_batch* b, const rf_sharded_corpus* c, rf_kind kind,
                                                const rf_args* args, uint32_t* const* out_device) {
  return sharded_score_allgather(b, c, kind, args, (void* const*)out_device, false);
}
rf_status rf_sharded_score_f64_allgather_device(const rf_sharded_batch* b, const rf_sharded_corpus* c, rf_kind kind,
                                                const rf_args* args, double* const* out_device) {
  return sharded_score_allgather(b, c, kind, args, (void* const*)out_device, true);
}

// ---- one process per GPU (MPI / torchrun style hosts): a communicator handle + the same overlapped scan + all-gather
struct rf_comm {
  int nranks = 1, rank = 0, device = 0;
  ncclComm_t comm = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev[kMaxChunks + 1] = {};
  uint64_t* d_counts = nullptr;  // [nranks + 1]: all ranks' candidate counts, then mine
  std::vector<uint64_t> counts;
  const rf_corpus* counts_for = nullptr;
  uint64_t counts_n = 0;
  std::mutex mu;
};

rf_status rf_comm_unique_id(void* out128) {
  if (!out128) return rfi::fail(RF_ERR_INVALID_ARG, "NULL argument");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  ncclResult_t r = ncclGetUniqueId(&id);
  if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
  memcpy(out128, &id, 128);
  return RF_OK;
}

rf_status rf_comm_create_rank(const void* id128, int nranks, int rank, int device, rf_comm** out) {
  if (!out) return rfi::fail(RF_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return rfi::fail(RF_ERR_INVALID_ARG, "id / nranks / rank");
  if (device < 0 || device >= rf_device_count()) return rfi::fail(RF_ERR_CUDA, "no such CUDA device");
  rf_comm* cm = new (std::nothrow) rf_comm();
  if (!cm) return rfi::fail(RF_ERR_OOM, "host allocation failed");
  cm->nranks = nranks;
  cm->rank = rank;
  cm->device = device;
  DevGuard g(device);
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclResult_t r = ncclCommInitRank(&cm->comm, nranks, id, rank);
  if (r != ncclSuccess) { delete cm; return nccl_fail(r, "ncclCommInitRank"); }
  int lo_pri = 0, hi_pri = 0;
  cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
  cudaError_t e = cudaStreamCreateWithPriority(&cm->comm_stream, cudaStreamNonBlocking, hi_pri);
  for (int k = 0; k <= kMaxChunks && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&cm->ev[k], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(&cm->d_counts, (size_t)(nranks + 1) * 8);
  if (e != cudaSuccess) { rf_comm_destroy(cm); return rfi::cuda_fail(e, "communicator resources"); }
  *out = cm;
  return RF_OK;
}

rf_status rf_comm_destroy(rf_comm* cm) {
  if (!cm) return RF_OK;
  DevGuard g(cm->device);
  if (cm->comm) ncclCommDestroy(cm->comm);
  for (int k = 0; k <= kMaxChunks; ++k)
    if (cm->ev[k]) cudaEventDestroy(cm->ev[k]);
  if (cm->comm_stream) cudaStreamDestroy(cm->comm_stream);
  if (cm->d_counts) cudaFree(cm->d_counts);
  delete cm;
  return RF_OK;
}
int rf_comm_rank(const rf_comm* cm) { return cm ? cm->rank : -1; }
int rf_comm_size(const rf_comm* cm) { return cm ? cm->nranks : 0; }

static rf_status comm_score_allgather(const rf_batch* b, const rf_corpus* c, rf_comm* cm, rf_kind kind, const rf_args* args,
                                      void* out_device, uint64_t out_capacity, uint64_t* counts_out, bool want_f64, cudaStream_t st) {
  if (!b || !c || !cm) return rfi::fail(RF_ERR_INVALID_ARG, "NULL handle");
  if (c->device != cm->device || b->device != cm->device) return rfi::fail(RF_ERR_INVALID_ARG, "handles live on different devices");
  std::lock_guard<std::mutex> lk(cm->mu);
  DevGuard g(cm->device);
  // every rank's candidate count (one tiny all-gather, cached per corpus)
  if (cm->counts_for != c || cm->counts_n != c->n || (int)cm->counts.size() != cm->nranks) {
    const uint64_t mine = c->n;
    cm->counts.assign(cm->nranks, 0);
    cudaError_t e = cudaMemcpyAsync(cm->d_counts + cm->nranks, &mine, 8, cudaMemcpyHostToDevice, cm->comm_stream);
    if (e != cudaSuccess) return rfi::cuda_fail(e, "count exchange");
    ncclResult_t r = ncclAllGather(cm->d_counts + cm->nranks, cm->d_counts, 1, ncclUint64, cm->comm, cm->comm_stream);
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllGather (counts)");
    e = cudaMemcpyAsync(cm->counts.data(), cm->d_counts, (size_t)cm->nranks * 8, cudaMemcpyDeviceToHost, cm->comm_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cm->comm_stream);
    if (e != cudaSuccess) return rfi::cuda_fail(e, "count exchange");
    cm->counts_for = c;
    cm->counts_n = c->n;
  }
  uint64_t total = 0;
  for (int r = 0; r < cm->nranks; ++r) total += cm->counts[r];
  if (counts_out) memcpy(counts_out, cm->counts.data(), (size_t)cm->nranks * 8);
  if (total == 0) return RF_OK;
  if (!out_device) return rfi::fail(RF_ERR_INVALID_ARG, "out_device is NULL");
  if (out_capacity < total) return rfi::fail(RF_ERR_INVALID_ARG, "out_device holds fewer than the " + std::to_string(total) + " results of all ranks");
  GatherLocal loc{b, c, cm->rank, cm->device, cm->comm, st, cm->comm_stream, cm->ev, out_device};
  return scan_allgather_nccl(&loc, 1, cm->nranks, cm->counts.data(), kind, args, want_f64);
}

rf_status rf_batch_score_u32_allgather_device(const rf_batch* b, const rf_corpus* c, rf_comm* comm, rf_kind kind, const rf_args* args,
                                              uint32_t* out_device, uint64_t out_capacity, uint64_t* counts_out, void* stream) {
  return comm_score_allgather(b, c, comm, kind, args, out_device, out_capacity, counts_out, false, (cudaStream_t)stream);
}
rf_status rf_batch_score_f64_allgather_device(const rf_batch* b, const rf_corpus* c, rf_comm* comm, rf_kind kind, const rf_args* args,
                                              double* out_device, uint64_t out_capacity, uint64_t* counts_out, void* stream) {
  return comm_score_allgather(b, c, comm, kind, args, out_device, out_capacity, counts_out, true, (cudaStream_t)stream);
}

// ---- k best of the whole sharded corpus: per-shard selection on the devices, k entries per shard to the host, merged
// by (score best-first, global index ascending)
static rf_status sharded_extract(const rf_sharded_batch* b, const rf_sharded_corpus* c, rf_kind kind, const rf_args* args, uint32_t k,
                                 uint64_t* idx_out, void* score_out, uint32_t* n_out, bool want_f64) {
  rf_status s = check_pair(b, c);
  if (s != RF_OK) return s;
  if (k == 0 || k > 1024) return rfi::fail(RF_ERR_INVALID_ARG, "k must be in 1..1024");
  if (!idx_out || !score_out || !n_out) return rfi::fail(RF_ERR_INVALID_ARG, "NULL output");
  *n_out = 0;
  const size_t nd = c->devices.size(), esz = want_f64 ? 8 : 4;
  std::vector<std::vector<uint32_t>> idx(nd, std::vector<uint32_t>(k));
  std::vector<std::vector<uint8_t>> sc(nd, std::vector<uint8_t>((size_t)k * esz));
  std::vector<uint32_t> cnt(nd, 0);
  s = for_each_shard(nd, [&](size_t i) {
    return rfi::select_host(b->per[i], c->shard[i], kind, args, want_f64, false, k, 0, idx[i].data(), sc[i].data(), &cnt[i], nullptr);
  });
  if (s != RF_OK) return s;
  const rf_kind ek = (b->metric == RF_RATIO) ? RF_NORMALIZED_SIMILARITY : kind;
  const bool desc = ek == RF_SIMILARITY || ek == RF_NORMALIZED_SIMILARITY;
  struct Ent { double key; uint64_t gidx; size_t shard; uint32_t pos; };
  std::vector<Ent> all;
  for (size_t i = 0; i < nd; ++i)
    for (uint32_t j = 0; j < cnt[i]; ++j) {
      const double v = want_f64 ? ((const double*)sc[i].data())[j] : (double)((const uint32_t*)sc[i].data())[j];
      all.push_back(Ent{desc ? -v : v, c->lo[i] + idx[i][j], i, j});
    }
  std::sort(all.begin(), all.end(), [](const Ent& a, const Ent& z) { return a.key != z.key ? a.key < z.key : a.gidx < z.gidx; });
  const size_t m = std::min<size_t>(all.size(), k);
  for (size_t t = 0; t < m; ++t) {
    idx_out[t] = all[t].gidx;
    memcpy((uint8_t*)score_out + t * esz, sc[all[t].shard].data() + (size_t)all[t].pos * esz, esz);
  }
  *n_out = (uint32_t)m;
  return RF_OK;
}

rf_status rf_sharded_extract_u32(const rf_sharded_batch* b, const rf_sharded_corpus* c, rf_kind kind, const rf_args* args, uint32_t k,
                                 uint64_t* idx_out, uint32_t* score_out, uint32_t* n_out) {
  return sharded_extract(b, c, kind, args, k, idx_out, score_out, n_out, false);
}
rf_status rf_sharded_extract_f64(const rf_sharded_batch* b, const rf_sharded_corpus* c, rf_kind kind, const rf_args* args, uint32_t k,
                                 uint64_t* idx_out, double* score_out, uint32_t* n_out) {
  return sharded_extract(b, c, kind, args, k, idx_out, score_out, n_out, true);
}

// ---- many-vs-many top-k over the sharded corpus: per-shard scan -> all-gather of the lists -> merge on the device
rf_status rf_sharded_cdist_topk_u8(const uint8_t* q_chars, const uint64_t* q_offsets, uint32_t nq, const rf_sharded_corpus* cc,
                                   const rf_args* args, uint32_t k, uint64_t* idx_host, uint32_t* dist_host) {
  if (!cc) return rfi::fail(RF_ERR_INVALID_ARG, "NULL corpus");
  if (nq == 0) return RF_OK;
  if (!idx_host || !dist_host) return rfi::fail(RF_ERR_INVALID_ARG, "NULL output");
  if (k == 0 || k > 64) return rfi::fail(RF_ERR_INVALID_ARG, "k must be in 1..64");
  rf_sharded_corpus* c = const_cast<rf_sharded_corpus*>(cc);
  const size_t nd = c->devices.size();
  if ((uint64_t)nd * k > 25600) return rfi::fail(RF_ERR_UNSUPPORTED, "shards * k must not exceed 25600");
  const size_t kk = (size_t)nq * k, part = 2 * kk;  // u32 words per shard: [idx nq*k][dist nq*k]
  std::vector<uint32_t*> d_parts(nd, nullptr);       // on every device: [nd][2][nq][k]
  uint64_t* d_base = nullptr;                         // device 0: first candidate of every shard
  uint64_t* d_oidx = nullptr;
  uint32_t* d_odist = nullptr;
  rf_status s = RF_OK;
  for (size_t i = 0; i < nd && s == RF_OK; ++i) {
    DevGuard g(c->devices[i]);
    cudaError_t e = dev_alloc(&d_parts[i], nd * part * 4, c->streams[i]);
    if (e != cudaSuccess) s = rfi::cuda_fail(e, "cdist lists");
  }
  // the scans run concurrently, one host thread per shard (rf_cdist_topk_u8_device returns after its stream has drained)
  if (s == RF_OK)
    s = for_each_shard(nd, [&](size_t i) {
      DevGuard g(c->devices[i]);
      uint32_t* mine = d_parts[i] + i * part;
      return rfi::cdist(q_chars, q_offsets, nq, c->shard[i], args, k, mine, mine + kk, true, c->streams[i]);
    });
  if (s == RF_OK) {
    std::vector<uint64_t> off(nd + 1);
    for (size_t i = 0; i <= nd; ++i) off[i] = i * part * 4;
    if (nd > 1 && use_nccl(c)) {  // equal parts: one in-place ncclAllGather per device
      std::lock_guard<std::mutex> lk(c->coll_mu);
      s = ensure_comms(c);
      if (s == RF_OK) {
        ncclResult_t r = ncclGroupStart();
        for (size_t i = 0; i < nd && r == ncclSuccess; ++i)
          r = ncclAllGather(d_parts[i] + i * part, d_parts[i], part, ncclUint32, c->comms[i], c->streams[i]);
        ncclResult_t r2 = ncclGroupEnd();
        if (r != ncclSuccess) s = nccl_fail(r, "ncclAllGather");
        else if (r2 != ncclSuccess) s = nccl_fail(r2, "ncclGroupEnd");
      }
    } else {
      s = allgatherv_inplace(c, (void* const*)d_parts.data(), off.data());
    }
  }
  if (s == RF_OK) {  // merge on the first device; its stream is ordered behind the gather
    DevGuard g(c->devices[0]);
    cudaStream_t st = c->streams[0];
    cudaError_t e = dev_alloc(&d_base, nd * 8, st);
    if (e == cudaSuccess) e = dev_alloc(&d_oidx, kk * 8, st);
    if (e == cudaSuccess) e = dev_alloc(&d_odist, kk * 4, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_base, c->lo.data(), nd * 8, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) s = rfi::cuda_fail(e, "cdist merge buffers");
    if (s == RF_OK)
      s = rf_topk_merge_device(d_parts[0], d_parts[0] + kk, part, d_base, (uint32_t)nd, nq, k, d_oidx, d_odist, c->devices[0], st);
    if (s == RF_OK) {
      e = cudaMemcpyAsync(idx_host, d_oidx, kk * 8, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(dist_host, d_odist, kk * 4, cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) s = rfi::cuda_fail(e, "cdist result download");
    }
  }
  const std::string keep = s != RF_OK ? rfi::last_error() : std::string();
  rf_status s2 = sync_all(c);
  {
    DevGuard g(c->devices[0]);
    dev_free(d_base, c->streams[0]);
    dev_free(d_oidx, c->streams[0]);
    dev_free(d_odist, c->streams[0]);
  }
  for (size_t i = 0; i < nd; ++i) {
    DevGuard g(c->devices[i]);
    dev_free(d_parts[i], c->streams[i]);
  }
  if (s != RF_OK) return rfi::fail(s, keep);
  return s2;
}

// ---- streaming from host memory over several devices: the candidate range is split by bytes, every device runs its own
// chunked H2D / scan / D2H pipeline (rf_batch_stream_*) on its own PCIe link, results land in the caller's vector
static rf_status sharded_stream(const rf_sharded_batch* b, const uint8_t* chars, const uint64_t* offsets, uint64_t n, rf_kind kind,
                                const rf_args* args, void* out_host, bool want_f64) {
  if (!b) return rfi::fail(RF_ERR_INVALID_ARG, "NULL handle");
  if (n == 0) return RF_OK;
  if (!offsets || !out_host) return rfi::fail(RF_ERR_INVALID_ARG, "NULL argument");
  if (offsets[0] != 0) return rfi::fail(RF_ERR_INVALID_ARG, "offsets[0] must be 0");
  const size_t nd = b->devices.size(), esz = want_f64 ? 8 : 4;
  std::vector<uint64_t> lo;
  split_by_bytes(offsets, n, nd, &lo);
  return for_each_shard(nd, [&](size_t i) {
    if (lo[i + 1] == lo[i]) return RF_OK;
    return rfi::stream_u64(b->per[i], chars, offsets + lo[i], lo[i + 1] - lo[i], kind, args, (uint8_t*)out_host + lo[i] * esz, want_f64);
  });
}
// ... and with one length byte per candidate on the wire (optionally 6-bit packed characters, byte results): no static
// split at all -- the devices' workers take chunks from ONE shared planner as their pipeline slots free up, so every
// PCIe link runs at whatever rate the host gives it and the call ends when the AGGREGATE is through.
static rf_status sharded_stream_len8(const rf_sharded_batch* b, const uint8_t* chars, const uint8_t* dict64, const uint8_t* lens,
                                     uint64_t n, rf_kind kind, const rf_args* args, void* out_host, bool out_u8) {
  if (!b) return rfi::fail(RF_ERR_INVALID_ARG, "NULL handle");
  if (n == 0) return RF_OK;
  if (!lens || !out_host) return rfi::fail(RF_ERR_INVALID_ARG, "NULL argument");
  void* plan = rfi::stream_len8_plan_create(lens, n, dict64 != nullptr);
  if (!plan) return rfi::fail(RF_ERR_OOM, "host allocation failed");
  // a device listed twice shares one pipeline (the per-device streaming context is exclusive): one worker per distinct device
  std::vector<size_t> workers;
  for (size_t i = 0; i < b->devices.size(); ++i) {
    bool seen = false;
    for (size_t j : workers) seen = seen || b->devices[j] == b->devices[i];
    if (!seen) workers.push_back(i);
  }
  const rf_status s = for_each_shard(workers.size(), [&](size_t w) {
    return rfi::stream_len8_shared(b->per[workers[w]], chars, dict64, lens, n, kind, args, out_host, out_u8, plan);
  });
  rfi::stream_len8_plan_destroy(plan);
  return s;
}
rf_status rf_sharded_stream_u32_len8(const rf_sharded_batch* b, const uint8_t* chars, const uint8_t* lens, uint64_t n, rf_kind kind,
                                     const rf_args* args, uint32_t* out_host) {
  return sharded_stream_len8(b, chars, nullptr, lens, n, kind, args, out_host, false);
}
rf_status rf_sharded_stream_u8_len8(const rf_sharded_batch* b, const uint8_t* chars, const uint8_t* lens, uint64_t n, rf_kind kind,
                                    const rf_args* args, uint8_t* out_host) {
  return sharded_stream_len8(b, chars, nullptr, lens, n, kind, args, out_host, true);
}
rf_status rf_sharded_stream_u8_len8_packed6(const rf_sharded_batch* b, const uint8_t* packed, const uint8_t* dict64, const uint8_t* lens,
                                            uint64_t n, rf_kind kind, const rf_args* args, uint8_t* out_host) {
  if (!dict64) return rfi::fail(RF_ERR_INVALID_ARG, "dict64 is NULL");
  return sharded_stream_len8(b, packed, dict64, lens, n, kind, args, out_host, true);
}
rf_status rf_sharded_stream_u32(const rf_sharded_batch* b, const uint8_t* chars, const uint64_t* offsets, uint64_t n, rf_kind kind,
                                const rf_args* args, uint32_t* out_host) {
  return sharded_stream(b, chars, offsets, n, kind, args, out_host, false);
}
rf_status rf_sharded_stream_f64(const rf_sharded_batch* b, const uint8_t* chars, const uint64_t* offsets, uint64_t n, rf_kind kind,
                                const rf_args* args, double* out_host) {
  return sharded_stream(b, chars, offsets, n, kind, args, out_host, true);
}

}  // extern "C"

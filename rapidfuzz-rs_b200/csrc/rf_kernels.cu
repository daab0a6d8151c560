// rf_kernels.cu -- hand-written sm_100a kernels for the one-vs-many scoring path.
//
//  scan_w1_kernel  query <= 64 elements.  One thread per candidate.  A producer warp streams candidate
//                  tiles (CSR offsets + packed bytes) HBM -> shared memory with TMA bulk copies
//                  (cp.async.bulk + mbarrier, double buffered); the compute warps bucket the tile's
//                  candidates by length (shared-memory counting sort) so that the 32 lanes of a warp run
//                  the same number of Myers/Hyyro steps, look the query's match masks up in a
//                  lane-replicated (bank-conflict-free) shared-memory table, and write results back
//                  coalesced.  Integer/bitwise work only -- no tensor cores.
//  scan_mw_kernel  query > 64 elements.  A sub-warp of G lanes per candidate, lane w owns 64-bit block(s)
//                  w of the bit-vectors; columns are skewed (lane w handles text char j-w at step j) and
//                  the horizontal carries (+ the text char itself) travel lane-to-lane in one
//                  __shfl_up_sync per step.
//  jaro_mw_kernel  Jaro / Jaro-Winkler with a multi-word query, thread per candidate.
//  cdist_*         many queries x corpus tile, per-query top-k.
#include <atomic>
#include <type_traits>
#include <cstdio>
#include <cstdlib>
#include "rf_kernels.cuh"

namespace rfk {

static std::atomic<uint64_t> g_launches{0};
uint64_t kernel_launch_count() { return g_launches.load(); }
void count_launches(uint64_t n) { g_launches.fetch_add(n); }

// Grid of a statically partitioned (grid-stride) kernel: exactly the CTAs that are resident at once.  A fixed "8 CTAs of
// 256 threads per SM" over-subscribes every kernel that needs more than 32 registers (40 registers -> 6 resident CTAs):
// the CTAs of the second wave start when the first wave retires and run their equal share of the work at a third of the
// occupancy, i.e. the kernel takes ~1.5x as long as the same work spread over one wave.
template <class K>
static uint64_t resident_ctas(K kern, int threads, size_t smem, int sm_count, int fallback_per_sm) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) {
    (void)cudaGetLastError();
    per_sm = fallback_per_sm;
  }
  return (uint64_t)sm_count * (uint64_t)per_sm;
}

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// same, but lets the hardware suspend the thread between polls (producer lane: must not burn issue slots)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(1000000u)
      : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// the same issued by ONE elected lane of a converged warp: arms the mbarrier with the byte count and starts the copy
// (operands are warp-uniform; elect.sync lets ptxas feed the uniform datapath without a per-value loop)
__device__ __forceinline__ void tma_bulk_g2s_elect(uint32_t dst_saddr, const void* src_gmem, uint32_t bytes, uint32_t bar_saddr) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "elect.sync _|P1, 0xffffffff;\n"
      "@P1 mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n"
      "@P1 cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
      "}\n" ::"r"(dst_saddr),
      "l"(src_gmem), "r"(bytes), "r"(bar_saddr)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_saddr(uint32_t bar_saddr, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar_saddr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int NT>
__device__ __forceinline__ void bar_compute() {  // named barrier 1: the NT compute threads only
  asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}

// ------------------------------------------------------------------------------------------------ w1
struct W1Params {
  const uint8_t* chars;
  const uint32_t* off32;
  const uint64_t* off64;
  uint64_t n;
  const void* tab;  // 256 x W compact match table (alignment per family, see QueryView)
  uint32_t len1;
  uint32_t T;  // candidates per tile, multiple of 4, <= TMAX
  uint32_t num_tiles;
  void* out;
  int out_f64;
  uint32_t two;  // the constant 2, opaque to ptxas: keeps x*2+1 an IMAD (FMA pipe) instead of an ALU-pipe LEA
  Epi epi;
};

template <class W, int NT, int TMAX, int BCAP>
struct W1Smem {
  alignas(16) W pm[256 * 32];            // pm[ch*32 + lane]: every lane owns a bank -> conflict-free gather
  alignas(16) uint8_t chars[2][BCAP + 32];
  alignas(16) uint8_t offs[2][(TMAX + 8) * 8];
  alignas(16) uint64_t res[TMAX];        // staged results (u32 or f64 view)
  uint32_t hist[256];
  uint16_t order[TMAX];
  alignas(8) uint64_t full[2];
  alignas(8) uint64_t empty[2];
};

// Where a candidate's bytes come from.
//  TileSrc: packed bytes at an arbitrary offset of a 4-byte aligned buffer (TMA-staged tile in shared memory, or
//           the CSR array in global memory).
//  LaneSrc: one lane's column of a length-bucketed, warp-interleaved group (8-byte row k of lane l at
//           col[k*32]): every load of a warp is 256 contiguous bytes, one row (8 chars) prefetched ahead.
struct TileReader : ByteReader {
  static constexpr bool kRow8 = false;
  __device__ __forceinline__ TileReader(const uint8_t* b, uint32_t s) : ByteReader(b, s) {}
};
struct TileSrc {
  const uint8_t* base;
  uint32_t start;
  __device__ __forceinline__ TileReader reader() const { return TileReader(base, start); }
  __device__ __forceinline__ uint32_t byte(uint32_t j) const { return base[start + j]; }
};
// STREAM: the row is read once per launch (one-vs-many) -> ld.global.cs (evict-first) so the scattered result
// sectors stay in L2; otherwise (many-vs-many re-reads the shard per query) -> plain read-only load, L2 resident.
template <bool STREAM>
__device__ __forceinline__ uint2 ld_row8(const uint2* p) {
  if constexpr (STREAM) return __ldcs(p);  // measured: .nc / .cg / .lu variants all within 0.5 % of this
  else return __ldg(p);
}
// ptxas sinks the row loads towards their first use to save registers, which shortens the look-ahead; an
// explicit L2 prefetch a few rows ahead (it needs no destination register, so nothing is gained by moving
// it) takes the DRAM latency off the critical path, the late load then only pays an L2 hit.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// L2 prefetch of the streamed rows: per-lane prefetch.global.L2 of the two rows kPfDist ahead.  Measured on B200
// (config 2): no prefetch 2.15 ms, distance 3 / 6 / 12 / 24 rows all 2.04 ms, one bulk L2 prefetch per warp 2.14 ms.
constexpr int kPfDist = 6;
template <bool STREAM>
struct LaneReaderT {
  static constexpr bool kRow8 = true;
  static constexpr bool kStream = STREAM;
  const uint2* p;  // next row to fetch
  uint2 q0, q1;    // two rows in flight / in registers (16 chars of look-ahead)
  uint2 cur;
  uint32_t half;
  __device__ __forceinline__ LaneReaderT(const uint2* third_row, uint2 first_row, uint2 second_row)
      : p(third_row), q0(first_row), q1(second_row), half(0) {}
  __device__ __forceinline__ uint2 next8() {
    const uint2 r = q0;
    q0 = q1;
    if constexpr (STREAM) prefetch_l2(p + 32 * 6);
    q1 = ld_row8<STREAM>(p);
    p += 32;
    return r;
  }
  __device__ __forceinline__ uint32_t next4() {
    if (half) { half = 0; return cur.y; }
    cur = next8();
    half = 1;
    return cur.x;
  }
};
template <bool STREAM>
struct LaneSrcT {
  const uint2* col;     // this lane's column: row k at col[k*32]
  uint2 first, second;  // rows 0 and 1, prefetched by the caller
  __device__ __forceinline__ LaneReaderT<STREAM> reader() const { return LaneReaderT<STREAM>(col + 64, first, second); }
  __device__ __forceinline__ uint32_t byte(uint32_t j) const {
    return reinterpret_cast<const uint8_t*>(col + (size_t)(j >> 3) * 32)[j & 7u];
  }
};

// Levenshtein, 32-bit words, match table in shared memory: same recurrence as rfk::lev_w1<uint32_t> with the
// address arithmetic spelled out so that it lands on the FMA pipe (the ALU pipe is the bound): per text char
//   PRMT (byte extract) . IMAD (ch*128 + lane base) . LDS . 7 LOP3 . 3 IMAD (add, HP*2+1, HN*2)
// `two` is the constant 2 passed as a kernel parameter: opaque to ptxas, so x*2+1 stays an IMAD (FMA pipe)
// instead of becoming an ALU-pipe LEA.
// OSA = true adds the transposition term of osa.rs:84-135 (TR = (((~D0_prev) & X) << 1) & X_prev, D0 |= TR):
// two more LOP3 and one more IMAD per character.
template <bool OSA, class Rd>
__device__ __forceinline__ uint32_t myers_w1_u32_fast(uint32_t pm_lane_saddr, Rd rd, uint32_t len2, uint32_t len1,
                                                      uint32_t two) {
  const uint32_t one = two >> 1;
  uint32_t VP = 0xFFFFFFFFu << (32u - len1);
  uint32_t VN = 0;
  uint32_t D0p = 0, Xp = 0;  // OSA only: previous column's D0 and match mask
  // table address of text byte K of w = byte*128 + lane base.  IDP.4A (dot product of the 4 bytes of w with a
  // selector word holding 128 in byte K, plus the base) does extract + scale + add in ONE FMA-pipe instruction;
  // the PRMT + IMAD pair it replaces costs an ALU-pipe slot, and the ALU pipe is this kernel's bound.
#ifdef RF_LEV_NO_DP4A
#define RF_LEV32_ADDR(K)                                                             \
    {                                                                                \
      const uint32_t ch = __byte_perm(w, 0u, 0x4440u + (K));                         \
      asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(addr) : "r"(ch), "r"(pm_lane_saddr)); \
    }
#else
#define RF_LEV32_ADDR(K) addr = __dp4a(w, 0x80u << (8 * (K)), pm_lane_saddr);
#endif
  // (HP << 1) | 1.  Default: one LOP3 (HP = VN | ~(D0 | VP)) + one IMAD.  RF_LEV_HP_ARITH (experiment):
  // VN and ~(D0 | VP) are disjoint and D0 | VP = D0 + VP - HN, so 2 HP + 1 = 2 (VN + HN - D0 - VP) - 1 -- four IMADs, no LOP3.
#ifdef RF_LEV_HP_ARITH
  const uint32_t minus1 = one - two;
#define RF_LEV32_HPHN                                                                \
    uint32_t HN = D0 & VP;                                                           \
    const uint32_t b2 = (VP * minus1 + VN) * two + minus1;                           \
    uint32_t HP = (D0 * minus1 + HN) * two + b2;                                     \
    HN = HN * two;
#else
#define RF_LEV32_HPHN                                                                \
    uint32_t HP = VN | ~(D0 | VP);                                                   \
    uint32_t HN = D0 & VP;                                                           \
    HP = HP * two + one;                                                             \
    HN = HN * two;
#endif
#define RF_LEV32_STEP(K)                                                             \
  {                                                                                  \
    uint32_t addr, X;                                                                \
    RF_LEV32_ADDR(K)                                                                 \
    asm("ld.shared.u32 %0, [%1];" : "=r"(X) : "r"(addr));                            \
    uint32_t D0 = ((((X & VP) + VP) ^ VP) | X) | VN;                                 \
    if constexpr (OSA) {                                                             \
      D0 |= (((~D0p) & X) * two) & Xp;                                               \
      D0p = D0;                                                                      \
      Xp = X;                                                                        \
    }                                                                                \
    RF_LEV32_HPHN                                                                    \
    VP = HN | ~(D0 | HP);                                                            \
    VN = HP & D0;                                                                    \
  }
  if constexpr (Rd::kRow8) {
    // Two rows (16 chars) always in flight: rows i+2 / i+3 are requested before rows i / i+1 are consumed, and
    // the loop is unrolled by two rows so that the loads rotate through registers without early moves.
#define RF_LEV32_ROW(R)                                                                      \
  {                                                                                          \
    { const uint32_t w = (R).x; RF_LEV32_STEP(0) RF_LEV32_STEP(1) RF_LEV32_STEP(2) RF_LEV32_STEP(3) } \
    { const uint32_t w = (R).y; RF_LEV32_STEP(0) RF_LEV32_STEP(1) RF_LEV32_STEP(2) RF_LEV32_STEP(3) } \
  }
    uint2 A = rd.q0, B = rd.q1;
    const uint2* p = rd.p;
    const uint32_t nfull = len2 >> 3;
    uint32_t i = 0;
    // four rows per iteration, the two register pairs swap roles: no register moves, half the loop overhead.
    // (the second load pair may run past this candidate's rows into the next group's or the slack: harmless)
    for (; i + 4 <= nfull; i += 4) {
      if constexpr (Rd::kStream) {
        prefetch_l2(p + 32 * kPfDist);
        prefetch_l2(p + 32 * (kPfDist + 1));
        prefetch_l2(p + 32 * (kPfDist + 2));
        prefetch_l2(p + 32 * (kPfDist + 3));
      }
      const uint2 C = ld_row8<Rd::kStream>(p);
      const uint2 D = ld_row8<Rd::kStream>(p + 32);
      RF_LEV32_ROW(A)
      RF_LEV32_ROW(B)
      A = ld_row8<Rd::kStream>(p + 64);
      B = ld_row8<Rd::kStream>(p + 96);
      p += 128;
      RF_LEV32_ROW(C)
      RF_LEV32_ROW(D)
    }
    if (i + 2 <= nfull) {
      if constexpr (Rd::kStream) {
        prefetch_l2(p + 32 * kPfDist);
        prefetch_l2(p + 32 * (kPfDist + 1));
      }
      const uint2 C = ld_row8<Rd::kStream>(p);
      const uint2 D = ld_row8<Rd::kStream>(p + 32);
      p += 64;
      RF_LEV32_ROW(A)
      RF_LEV32_ROW(B)
      A = C;
      B = D;
      i += 2;
    }
    if (i < nfull) {
      RF_LEV32_ROW(A)
      A = B;
    }
#undef RF_LEV32_ROW
    const uint32_t rem = len2 & 7u;
    if (rem) {
      const uint2 ww = A;
      { const uint32_t w = ww.x;
        RF_LEV32_STEP(0)
        if (rem > 1) RF_LEV32_STEP(1)
        if (rem > 2) RF_LEV32_STEP(2)
        if (rem > 3) RF_LEV32_STEP(3) }
      if (rem > 4) {
        const uint32_t w = ww.y;
        RF_LEV32_STEP(0)
        if (rem > 5) RF_LEV32_STEP(1)
        if (rem > 6) RF_LEV32_STEP(2)
      }
    }
  } else {
    const uint32_t nfull = len2 >> 2;
    for (uint32_t i = 0; i < nfull; ++i) {
      const uint32_t w = rd.next4();
      RF_LEV32_STEP(0)
      RF_LEV32_STEP(1)
      RF_LEV32_STEP(2)
      RF_LEV32_STEP(3)
    }
    const uint32_t rem = len2 & 3u;
    if (rem) {
      const uint32_t w = rd.next4();
      RF_LEV32_STEP(0)
      if (rem > 1) RF_LEV32_STEP(1)
      if (rem > 2) RF_LEV32_STEP(2)
    }
  }
#undef RF_LEV32_STEP
#undef RF_LEV32_HPHN
#undef RF_LEV32_ADDR
  return len2 + (uint32_t)__popc(VP) - (uint32_t)__popc(VN);
}
template <class Rd>
__device__ __forceinline__ uint32_t lev_w1_u32_fast(uint32_t pm_lane_saddr, Rd rd, uint32_t len2, uint32_t len1, uint32_t two) {
  return myers_w1_u32_fast<false>(pm_lane_saddr, rd, len2, len1, two);
}

// Row walker for the interleaved layout: the software pipeline of myers_w1_u32_fast (two rows in flight, four rows
// per iteration, L2 prefetch for streamed rows) around an arbitrary per-character step.  `st.template step<K>(w)`
// consumes byte K of the 32-bit word w.
template <class Rd, class St>
__device__ __forceinline__ void walk_rows8(Rd rd, uint32_t len2, St& st) {
  static_assert(Rd::kRow8, "interleaved rows only");
#define RF_WALK_ROW(R)                                                                                   \
  {                                                                                                      \
    { const uint32_t w = (R).x; st.template step<0>(w); st.template step<1>(w); st.template step<2>(w); st.template step<3>(w); } \
    { const uint32_t w = (R).y; st.template step<0>(w); st.template step<1>(w); st.template step<2>(w); st.template step<3>(w); } \
  }
  uint2 A = rd.q0, B = rd.q1;
  const uint2* p = rd.p;
  const uint32_t nfull = len2 >> 3;
  uint32_t i = 0;
  for (; i + 4 <= nfull; i += 4) {
    if constexpr (Rd::kStream) {
      prefetch_l2(p + 32 * kPfDist);
      prefetch_l2(p + 32 * (kPfDist + 1));
      prefetch_l2(p + 32 * (kPfDist + 2));
      prefetch_l2(p + 32 * (kPfDist + 3));
    }
    const uint2 C = ld_row8<Rd::kStream>(p);
    const uint2 D = ld_row8<Rd::kStream>(p + 32);
    RF_WALK_ROW(A)
    RF_WALK_ROW(B)
    A = ld_row8<Rd::kStream>(p + 64);
    B = ld_row8<Rd::kStream>(p + 96);
    p += 128;
    RF_WALK_ROW(C)
    RF_WALK_ROW(D)
  }
  if (i + 2 <= nfull) {
    if constexpr (Rd::kStream) {
      prefetch_l2(p + 32 * kPfDist);
      prefetch_l2(p + 32 * (kPfDist + 1));
    }
    const uint2 C = ld_row8<Rd::kStream>(p);
    const uint2 D = ld_row8<Rd::kStream>(p + 32);
    p += 64;
    RF_WALK_ROW(A)
    RF_WALK_ROW(B)
    A = C;
    B = D;
    i += 2;
  }
  if (i < nfull) {
    RF_WALK_ROW(A)
    A = B;
  }
#undef RF_WALK_ROW
  const uint32_t rem = len2 & 7u;
  if (rem) {
    const uint2 ww = A;
    { const uint32_t w = ww.x;
      st.template step<0>(w);
      if (rem > 1) st.template step<1>(w);
      if (rem > 2) st.template step<2>(w);
      if (rem > 3) st.template step<3>(w); }
    if (rem > 4) {
      const uint32_t w = ww.y;
      st.template step<0>(w);
      if (rem > 5) st.template step<1>(w);
      if (rem > 6) st.template step<2>(w);
    }
  }
}

// LCS length (Hyyro, lcs_seq.rs:222-257), 32-bit words, bottom-aligned table in shared memory.  Per text char:
// IDP.4A (address) . LDS . LOP3 (U = S & X) . IMAD (S + U; `one` is opaque so it stays on the FMA pipe) . LOP3
// ((S + U) | (S & ~U)): two ALU-pipe ops against Levenshtein's seven -- this kernel is bound by HBM, not by a pipe.
struct Lcs32Step {
  uint32_t S, base, one;
  template <int K>
  __device__ __forceinline__ void step(uint32_t w) {
    uint32_t X;
    const uint32_t addr = __dp4a(w, 0x80u << (8 * K), base);
    asm("ld.shared.u32 %0, [%1];" : "=r"(X) : "r"(addr));
    const uint32_t U = S & X;
    S = (U * one + S) | (S & ~U);
  }
};
template <class Rd>
__device__ __forceinline__ uint32_t lcs_w1_u32_fast(uint32_t pm_lane_saddr, Rd rd, uint32_t len2, uint32_t two) {
  Lcs32Step st{0xFFFFFFFFu, pm_lane_saddr, two >> 1};
  walk_rows8(rd, len2, st);
  return (uint32_t)__popc(~st.S);
}

// The same for queries of 33..64 elements on 32-bit halves: split low / high word tables (see lev_w1_u64_fast), the
// 64-bit S + U as IADD3 (carry out) + IMAD.X.
struct Lcs64Step {
  uint32_t Sl, Sh, base, one;
  template <int K>
  __device__ __forceinline__ void step(uint32_t w) {
    uint32_t Xl, Xh, al, ah;
    const uint32_t addr = __dp4a(w, 0x80u << (8 * K), base);
    asm("ld.shared.u32 %0, [%1];" : "=r"(Xl) : "r"(addr));
    asm("ld.shared.u32 %0, [%1+32768];" : "=r"(Xh) : "r"(addr));
    const uint32_t Ul = Sl & Xl, Uh = Sh & Xh;
    asm("{\n\tadd.cc.u32 %0, %2, %3;\n\tmadc.lo.u32 %1, %4, %5, %6;\n\t}"
        : "=r"(al), "=r"(ah) : "r"(Ul), "r"(Sl), "r"(Sh), "r"(one), "r"(Uh));
    Sl = al | (Sl & ~Ul);
    Sh = ah | (Sh & ~Uh);
  }
};
template <class Rd>
__device__ __forceinline__ uint32_t lcs_w1_u64_fast(uint32_t pm_lane_saddr, Rd rd, uint32_t len2, uint32_t two) {
  Lcs64Step st{0xFFFFFFFFu, 0xFFFFFFFFu, pm_lane_saddr, two >> 1};
  walk_rows8(rd, len2, st);
  return (uint32_t)__popc(~st.Sl) + (uint32_t)__popc(~st.Sh);
}

// Levenshtein, query of 33..64 elements: the 64-bit recurrence of lev_w1<uint64_t> written on 32-bit halves so that,
// as in the 32-bit routine, only the 14 LOP3 of a step remain on the ALU pipe.  ptxas expands a 64-bit add into
// IADD3 + IADD3.X and a 64-bit shift into SHF + SHL, all ALU-pipe; here
//   sum  = IMAD.WIDE.U32 (X&VP).lo * 1 + VP  ;  sum.hi += (X&VP).hi        (IMAD)
//   HP<<1|1, HN<<1 = IMAD.WIDE.U32 lo * 2 + {1|0}  (the product's high word is the bit that crosses) ; hi*2 + carry (IMAD)
// The match table is SPLIT: low words at pm_lo[ch*32 + lane], high words 32 KB further (one IDP.4A address, two LDS).
// OSA = true adds the transposition term of osa.rs:84-135 on the halves (TR = (((~D0_prev) & X) << 1) & X_prev).
template <bool OSA, class Rd>
__device__ __forceinline__ uint32_t myers_w1_u64_fast(uint32_t pm_lane_saddr, Rd rd, uint32_t len2, uint32_t len1, uint32_t two) {
  static_assert(Rd::kRow8, "interleaved rows only");
  const uint32_t one = two >> 1;
  const uint64_t vp0 = ~0ull << (64u - len1);
  uint32_t VPl = (uint32_t)vp0, VPh = (uint32_t)(vp0 >> 32), VNl = 0, VNh = 0;
  uint32_t D0pl = 0, D0ph = 0, Xpl = 0, Xph = 0;  // OSA only: previous column's D0 and match mask
#define RF_LEV64_STEP(K)                                                                              \
  {                                                                                                   \
    uint32_t Xl, Xh, sl, sh, c;                                                                       \
    const uint32_t addr = __dp4a(w, 0x80u << (8 * (K)), pm_lane_saddr);                               \
    asm("ld.shared.u32 %0, [%1];" : "=r"(Xl) : "r"(addr));                                            \
    asm("ld.shared.u32 %0, [%1+32768];" : "=r"(Xh) : "r"(addr));                                      \
    /* 64-bit (X & VP) + VP: IADD3 (carry out) + IMAD.X (VPh * 1 + (Xh & VPh) + carry) */             \
    asm("{\n\tadd.cc.u32 %0, %2, %3;\n\tmadc.lo.u32 %1, %4, %5, %6;\n\t}"                            \
        : "=r"(sl), "=r"(sh) : "r"(Xl & VPl), "r"(VPl), "r"(VPh), "r"(one), "r"(Xh & VPh));           \
    uint32_t D0l = ((sl ^ VPl) | Xl) | VNl;                                                           \
    uint32_t D0h = ((sh ^ VPh) | Xh) | VNh;                                                           \
    if constexpr (OSA) {                                                                              \
      uint32_t tl = ~D0pl & Xl, th = ~D0ph & Xh, tc;                                                  \
      asm("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}"                \
          : "=r"(tl), "=r"(tc) : "r"(tl), "r"(two));                                                  \
      th = th * two + tc;                                                                             \
      D0l |= tl & Xpl;                                                                                \
      D0h |= th & Xph;                                                                                \
      D0pl = D0l; D0ph = D0h; Xpl = Xl; Xph = Xh;                                                     \
    }                                                                                                 \
    uint32_t HPl = VNl | ~(D0l | VPl), HPh = VNh | ~(D0h | VPh);                                      \
    uint32_t HNl = D0l & VPl, HNh = D0h & VPh;                                                        \
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(c) : "r"(HPl), "r"(two));                                     \
    HPl = HPl * two + one;                                                                            \
    HPh = HPh * two + c;                                                                              \
    asm("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}"                  \
        : "=r"(HNl), "=r"(c) : "r"(HNl), "r"(two));                                                   \
    HNh = HNh * two + c;                                                                              \
    VPl = HNl | ~(D0l | HPl);                                                                         \
    VPh = HNh | ~(D0h | HPh);                                                                         \
    VNl = HPl & D0l;                                                                                  \
    VNh = HPh & D0h;                                                                                  \
  }
#define RF_LEV64_ROW(R)                                                                      \
  {                                                                                          \
    { const uint32_t w = (R).x; RF_LEV64_STEP(0) RF_LEV64_STEP(1) RF_LEV64_STEP(2) RF_LEV64_STEP(3) } \
    { const uint32_t w = (R).y; RF_LEV64_STEP(0) RF_LEV64_STEP(1) RF_LEV64_STEP(2) RF_LEV64_STEP(3) } \
  }
  uint2 A = rd.q0, B = rd.q1;
  const uint2* p = rd.p;
  const uint32_t nfull = len2 >> 3;
  uint32_t i = 0;
  for (; i + 2 <= nfull; i += 2) {  // two rows per iteration, the next two requested first
    if constexpr (Rd::kStream) {
      prefetch_l2(p + 32 * kPfDist);
      prefetch_l2(p + 32 * (kPfDist + 1));
    }
    const uint2 C = ld_row8<Rd::kStream>(p);
    const uint2 D = ld_row8<Rd::kStream>(p + 32);
    p += 64;
    RF_LEV64_ROW(A)
    RF_LEV64_ROW(B)
    A = C;
    B = D;
  }
  if (i < nfull) {
    RF_LEV64_ROW(A)
    A = B;
  }
  const uint32_t rem = len2 & 7u;
  if (rem) {
    const uint2 ww = A;
    { const uint32_t w = ww.x;
      RF_LEV64_STEP(0)
      if (rem > 1) RF_LEV64_STEP(1)
      if (rem > 2) RF_LEV64_STEP(2)
      if (rem > 3) RF_LEV64_STEP(3) }
    if (rem > 4) {
      const uint32_t w = ww.y;
      RF_LEV64_STEP(0)
      if (rem > 5) RF_LEV64_STEP(1)
      if (rem > 6) RF_LEV64_STEP(2)
    }
  }
#undef RF_LEV64_ROW
#undef RF_LEV64_STEP
  return len2 + (uint32_t)__popc(VPl) + (uint32_t)__popc(VPh) - (uint32_t)__popc(VNl) - (uint32_t)__popc(VNh);
}
template <class Rd>
__device__ __forceinline__ uint32_t lev_w1_u64_fast(uint32_t pm_lane_saddr, Rd rd, uint32_t len2, uint32_t len1, uint32_t two) {
  return myers_w1_u64_fast<false>(pm_lane_saddr, rd, len2, len1, two);
}

// The raw result of the bit-parallel kernels for one candidate: unit-cost Levenshtein / OSA distance, LCS length.
template <int FAM, class W, class Src, bool SPLIT64 = false>
__device__ __forceinline__ uint32_t raw_one(const W* __restrict__ pm_lane, const Src& src, uint32_t len2, uint32_t len1, uint32_t two) {
  static_assert(FAM == F_LEV || FAM == F_OSA || FAM == F_LCS, "integer metrics only");
  auto tab = [&](uint32_t ch) -> W { return pm_lane[ch * 32u]; };
  if (len1 == 0) return (FAM == F_LCS) ? 0u : len2;
  if constexpr (FAM == F_LEV && sizeof(W) == 4) return lev_w1_u32_fast(smem_u32(pm_lane), src.reader(), len2, len1, two);
  else if constexpr (FAM == F_LEV && SPLIT64) return lev_w1_u64_fast(smem_u32(pm_lane), src.reader(), len2, len1, two);
  else if constexpr (FAM == F_LEV) return lev_w1<W>(tab, src.reader(), len2, len1);
  else if constexpr (FAM == F_OSA && sizeof(W) == 4) return myers_w1_u32_fast<true>(smem_u32(pm_lane), src.reader(), len2, len1, two);
  else if constexpr (FAM == F_OSA && SPLIT64) return myers_w1_u64_fast<true>(smem_u32(pm_lane), src.reader(), len2, len1, two);
  else if constexpr (FAM == F_OSA) return osa_w1<W>(tab, src.reader(), len2, len1);
  else if constexpr (FAM == F_LCS && sizeof(W) == 4 && decltype(src.reader())::kRow8) return lcs_w1_u32_fast(smem_u32(pm_lane), src.reader(), len2, two);
  else if constexpr (FAM == F_LCS && SPLIT64) return lcs_w1_u64_fast(smem_u32(pm_lane), src.reader(), len2, two);
  else return lcs_w1<W>(tab, src.reader(), len2);
}

// One candidate: raw bit-parallel kernel + score algebra.  pm_lane = &pm[lane] of the lane-replicated table.
// SPLIT64 (F_LEV, 64-bit words, interleaved rows only): pm_lane points into the split low/high table of lev_w1_u64_fast.
template <int FAM, class W, class Src, bool SPLIT64 = false>
__device__ __forceinline__ void score_one(const W* __restrict__ pm_lane, const Src& src, uint32_t len2, uint32_t len1,
                                          const Epi& epi, int out_f64, uint32_t two, uint32_t& ru, double& rf) {
  auto tab = [&](uint32_t ch) -> W { return pm_lane[ch * 32u]; };
  if constexpr (FAM == F_JARO) {
    auto bytes = [&](uint32_t j) -> uint32_t { return src.byte(j); };
    auto jaro = [&](double c) { return jaro_similarity_w1(tab, bytes, len1, len2, c); };
    if (epi.metric == M_JARO) {
      rf = finish_float(epi, jaro);
    } else {
      uint32_t prefix = 0;  // common prefix, at most 4 (jaro_winkler.rs:118-123): q[i]==s[i] <=> bit i of PM[s[i]]
      while (prefix < 4 && prefix < len1 && prefix < len2 && ((tab(bytes(prefix)) >> prefix) & 1u)) ++prefix;
      const double pw = epi.prefix_weight;
      auto jw = [&](double c) { return jaro_winkler_from(jaro, prefix, pw, c); };
      rf = finish_float(epi, jw);
    }
  } else {
    const uint32_t raw = raw_one<FAM, W, Src, SPLIT64>(pm_lane, src, len2, len1, two);
    if (out_f64) rf = finish_norm(epi, raw, len1, len2);
    else ru = finish_int(epi, raw, len1, len2);
  }
}

template <int FAM, class W, int NT, int TMAX, int BCAP>
__global__ void __launch_bounds__(NT + 32) scan_w1_kernel(const __grid_constant__ W1Params p) {
  using Smem = W1Smem<W, NT, TMAX, BCAP>;
  constexpr int CPT = TMAX / NT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const uint32_t tid = threadIdx.x;
  const bool off64 = p.off64 != nullptr;
  const uint32_t osz = off64 ? 8u : 4u;

  if (tid == 0) {
    mbar_init(&S.full[0], 1);
    mbar_init(&S.full[1], 1);
    mbar_init(&S.empty[0], 1);
    mbar_init(&S.empty[1], 1);
    mbar_fence_init();
  }
  // replicate the 256-entry match table 32x: pm[ch*32 + lane]
  {
    const W* __restrict__ t = reinterpret_cast<const W*>(p.tab);
    for (uint32_t i = tid; i < 256u * 32u; i += NT + 32) S.pm[i] = t[i >> 5];
  }
  __syncthreads();

  if (tid >= NT) {
    // ===================== producer warp: one lane streams tiles with TMA bulk copies =====================
    if (tid != NT) return;
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t s = it & 1u;
      if (it >= 2) mbar_wait_sleep(&S.empty[s], ((it >> 1) - 1u) & 1u);
      const uint64_t t0 = (uint64_t)tile * p.T;
      const uint32_t tn = (uint32_t)((p.n - t0 < (uint64_t)p.T) ? (p.n - t0) : (uint64_t)p.T);
      const uint64_t o_lo = off64 ? p.off64[t0] : (uint64_t)p.off32[t0];
      const uint64_t o_hi = off64 ? p.off64[t0 + tn] : (uint64_t)p.off32[t0 + tn];
      const uint64_t a0 = o_lo & ~15ull;
      const uint64_t bytes = ((o_hi + 15ull) & ~15ull) - a0;
      const uint32_t off_bytes = ((tn + 1u) * osz + 15u) & ~15u;
      const bool fits = bytes <= (uint64_t)BCAP;
      const uint32_t tx = off_bytes + ((fits && bytes) ? (uint32_t)bytes : 0u);
      fence_proxy_async();
      mbar_expect_tx(&S.full[s], tx);
      const void* osrc = off64 ? (const void*)(p.off64 + t0) : (const void*)(p.off32 + t0);
      tma_bulk_g2s(S.offs[s], osrc, off_bytes, &S.full[s]);
      if (fits && bytes) tma_bulk_g2s(S.chars[s], p.chars + a0, (uint32_t)bytes, &S.full[s]);
    }
    return;
  }

  // ===================== compute warps =====================
  const uint32_t lane = tid & 31u;
  const W* __restrict__ pm_lane = S.pm + lane;
  uint32_t it = 0;
  for (uint32_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
    const uint32_t s = it & 1u;
    const uint64_t t0 = (uint64_t)tile * p.T;
    const uint32_t tn = (uint32_t)((p.n - t0 < (uint64_t)p.T) ? (p.n - t0) : (uint64_t)p.T);
    for (uint32_t i = tid; i < 256; i += NT) S.hist[i] = 0;
    mbar_wait(&S.full[s], (it >> 1) & 1u);
    const uint8_t* offs = S.offs[s];
    auto off_at = [&](uint32_t i) -> uint64_t {
      return off64 ? reinterpret_cast<const uint64_t*>(offs)[i] : (uint64_t) reinterpret_cast<const uint32_t*>(offs)[i];
    };
    const uint64_t a0 = off_at(0) & ~15ull;
    const bool in_smem = (((off_at(tn) + 15ull) & ~15ull) - a0) <= (uint64_t)BCAP;
    bar_compute<NT>();  // hist zeroed

    // ---- bucket by length: counting sort of the tile's candidates (keys clamp at 255)
    uint32_t key[CPT], pos[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const uint32_t i = c * NT + tid;
      if (i < tn) {
        const uint32_t len = (uint32_t)(off_at(i + 1) - off_at(i));
        key[c] = len < 255u ? len : 255u;
        pos[c] = atomicAdd(&S.hist[key[c]], 1u);
      }
    }
    bar_compute<NT>();
    if (tid < 32) {  // exclusive scan of the 256 bins by one warp
      uint32_t v[8], sum = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) { v[k] = S.hist[tid * 8 + k]; sum += v[k]; }
      uint32_t incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += t;
      }
      uint32_t run = incl - sum;
#pragma unroll
      for (int k = 0; k < 8; ++k) { S.hist[tid * 8 + k] = run; run += v[k]; }
    }
    bar_compute<NT>();
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const uint32_t i = c * NT + tid;
      if (i < tn) S.order[S.hist[key[c]] + pos[c]] = (uint16_t)i;
    }
    bar_compute<NT>();

    // ---- score: rank r*NT+tid of the length-sorted order -> a warp's 32 candidates have ~equal length
    uint32_t* res_u = reinterpret_cast<uint32_t*>(S.res);
    double* res_f = reinterpret_cast<double*>(S.res);
    // Rounds alternate direction (boustrophedon) so that every warp gets short AND long candidates:
    // without it the warp holding the longest ranks of every round is the one the barrier waits for.
    for (uint32_t r = 0; r * NT < tn; ++r) {
      const uint32_t rank = r * NT + ((r & 1u) ? (NT - 1u - tid) : tid);
      if (rank >= tn) continue;
      const uint32_t i = S.order[rank];
      const uint64_t o0 = off_at(i);
      const uint32_t len2 = (uint32_t)(off_at(i + 1) - o0);
      uint32_t ru = 0;
      double rf = 0.0;
      if (in_smem) {
        score_one<FAM, W>(pm_lane, TileSrc{S.chars[s], (uint32_t)(o0 - a0)}, len2, p.len1, p.epi, p.out_f64, p.two, ru, rf);
      } else {  // tile larger than the staging buffer (long candidates): read straight from global / L1
        score_one<FAM, W>(pm_lane, TileSrc{p.chars + (o0 & ~3ull), (uint32_t)(o0 & 3ull)}, len2, p.len1, p.epi, p.out_f64, p.two, ru, rf);
      }
      if (p.out_f64) res_f[i] = rf;
      else res_u[i] = ru;
    }
    bar_compute<NT>();  // stage s fully consumed, results staged
    if (tid == 0) mbar_arrive(&S.empty[s]);
    if (p.out_f64) {
      double* out = reinterpret_cast<double*>(p.out) + t0;
      for (uint32_t i = tid; i < tn; i += NT) out[i] = res_f[i];
    } else {
      uint32_t* out = reinterpret_cast<uint32_t*>(p.out) + t0;
      for (uint32_t i = tid; i < tn; i += NT) out[i] = res_u[i];
    }
  }
}

template <int FAM, class W, int NT, int TMAX, int BCAP>
static cudaError_t launch_w1_inst(const ScanLaunch& L, const void* tab) {
  using Smem = W1Smem<W, NT, TMAX, BCAP>;
  auto kern = scan_w1_kernel<FAM, W, NT, TMAX, BCAP>;
  const size_t smem = sizeof(Smem);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int ctas_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, NT + 32, smem);
  if (e != cudaSuccess) return e;
  if (ctas_per_sm < 1) ctas_per_sm = 1;

  W1Params p{};
  p.chars = L.corpus.chars;
  p.off32 = L.corpus.off32;
  p.off64 = L.corpus.off64;
  p.n = L.corpus.n;
  p.tab = tab;
  p.len1 = L.query.len1;
  // candidates per tile: keep the average tile ~85% of the staging buffer
  const double avg = L.corpus.n ? (double)L.corpus.total / (double)L.corpus.n : 0.0;
  uint64_t T = TMAX;
  if (avg > 0.0) {
    const uint64_t fit = (uint64_t)((double)BCAP * 0.85 / avg);
    if (fit < T) T = fit;
  }
  if (T >= NT) T = T / NT * NT;
  else T = T / 32 * 32;
  if (T < 32) T = 32;
  p.T = (uint32_t)T;
  p.num_tiles = (uint32_t)((L.corpus.n + T - 1) / T);
  p.out = L.out;
  p.out_f64 = L.out_is_f64;
  p.two = 2;
  p.epi = L.epi;
  uint32_t grid = (uint32_t)L.sm_count * (uint32_t)ctas_per_sm;
  if (grid > p.num_tiles) grid = p.num_tiles;
  kern<<<grid, NT + 32, smem, L.stream>>>(p);
  g_launches.fetch_add(1);
  return cudaGetLastError();
}

cudaError_t launch_scan_w1(const ScanLaunch& L) {
  const Family fam = family_of(L.epi.metric, L.epi.wclass);
  const bool w32 = L.query.len1 <= 32;
  static const int tune = getenv("RF_W1_TUNE") ? atoi(getenv("RF_W1_TUNE")) : 0;  // tile-shape experiments
  // 32-bit words: 32 KB table, 24 KB tiles -> 2 CTAs/SM;  64-bit words: 64 KB table, 48 KB tiles -> 1 CTA/SM
  switch (fam) {
    case F_LEV:
      if (w32 && tune == 1) return launch_w1_inst<F_LEV, uint32_t, 384, 768, 32768>(L, L.query.tab32_top);
      if (w32 && tune == 2) return launch_w1_inst<F_LEV, uint32_t, 512, 1024, 45056>(L, L.query.tab32_top);
      return w32 ? launch_w1_inst<F_LEV, uint32_t, 256, 512, 24576>(L, L.query.tab32_top)
                 : launch_w1_inst<F_LEV, uint64_t, 512, 1024, 49152>(L, L.query.tab64_top);
    case F_OSA:
      return w32 ? launch_w1_inst<F_OSA, uint32_t, 256, 512, 24576>(L, L.query.tab32_top)
                 : launch_w1_inst<F_OSA, uint64_t, 512, 1024, 49152>(L, L.query.tab64_top);
    case F_LCS:
      return w32 ? launch_w1_inst<F_LCS, uint32_t, 256, 512, 24576>(L, L.query.tab32_bot)
                 : launch_w1_inst<F_LCS, uint64_t, 512, 1024, 49152>(L, L.query.tab64_bot);
    default:
      return launch_w1_inst<F_JARO, uint64_t, 512, 1024, 49152>(L, L.query.tab64_bot);
  }
}

// ------------------------------------------------------------------------------------------------ lb
// Single-word path over the length-bucketed, warp-interleaved layout (rf_layout.cu): one warp per group of
// 32 equal-length candidates, one thread per candidate.  No tile staging, no sorting, no CTA barriers in the
// steady state: candidate words arrive as coalesced 128-byte lines, the only shared-memory traffic is the
// conflict-free match-table gather, and all 32 lanes run the same trip count.
struct LbParams {
  LbView lb;
  const void* tab;
  const double* quot;           // QueryView::quot (Jaro kernels)
  uint32_t len1;
  void* out;
  int out_f64;
  uint32_t two;
  uint32_t chunk;               // consecutive groups handed to a warp at a time
  unsigned long long* counter;  // dynamic chunk scheduler (zeroed before the launch)
  unsigned long long* flag;     // scan_jaro32_kernel: set when a group was left to jaro32_long_kernel (zeroed before the launch)
  const double* jtab;           // row-wise Jaro kernels: the whole f64 epilogue as a table (jaro_epi_table_kernel), or NULL
  uint32_t j_l2dim, j_ccdim, j_trdim, j_pfdim;
  Epi epi;
};

// RAWDIST: the launch asks for the plain unit-cost distance of Levenshtein / OSA as u32 without a cutoff
// (BatchComparator::distance, the headline call): the epilogue is the raw kernel result, so the per-group
// score algebra (a dozen uniform loads and branches on Epi) is compiled out.
template <int FAM, class W, int NT, bool RAWDIST>
__global__ void __launch_bounds__(NT) scan_lb_kernel(const __grid_constant__ LbParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr bool SPLIT64 = ((FAM == F_LEV || FAM == F_LCS || FAM == F_OSA) && sizeof(W) == 8);  // low / high words in two 32 KB tables (lev_w1_u64_fast, lcs_w1_u64_fast)
  W* pm = reinterpret_cast<W*>(smem_raw);
  {
    const W* __restrict__ t = reinterpret_cast<const W*>(p.tab);
    if constexpr (SPLIT64) {
      uint32_t* pm32 = reinterpret_cast<uint32_t*>(smem_raw);
      for (uint32_t i = threadIdx.x; i < 256u * 32u; i += NT) {
        const uint64_t v = t[i >> 5];
        pm32[i] = (uint32_t)v;
        pm32[i + 256u * 32u] = (uint32_t)(v >> 32);
      }
    } else {
      for (uint32_t i = threadIdx.x; i < 256u * 32u; i += NT) pm[i] = t[i >> 5];
    }
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const W* __restrict__ pm_lane = SPLIT64 ? reinterpret_cast<const W*>(reinterpret_cast<const uint32_t*>(smem_raw) + lane) : pm + lane;
  const uint64_t total_warps = (uint64_t)gridDim.x * (NT / 32);
  const uint64_t ngroups = p.lb.ngroups;
  const uint64_t nchunks = (ngroups + p.chunk - 1) / p.chunk;
  const uint2* __restrict__ gdata = reinterpret_cast<const uint2*>(p.lb.gdata);
  // first chunk statically, further chunks from the global counter (groups differ 8x in cost)
  uint64_t chunk = (uint64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  while (chunk < nchunks) {
    unsigned long long next_chunk = 0;
    if (lane == 0) next_chunk = total_warps + atomicAdd(p.counter, 1ull);  // latency hidden behind this chunk
    const uint64_t g0 = chunk * p.chunk;
    const uint32_t ng = (uint32_t)((g0 + p.chunk < ngroups) ? p.chunk : ngroups - g0);  // groups in this chunk
    // running per-lane pointers (32-bit group counter, no per-group 64-bit index arithmetic)
    const uint32_t* lens_p = p.lb.lens + g0 * 32 + lane;
    const uint32_t* perm_p = p.lb.perm + g0 * 32 + lane;
    const uint2* grp = gdata + __ldg(p.lb.goff + g0) * 32 + lane;  // this lane's column of the current group's first row
    uint32_t len_n = __ldg(lens_p);
    uint2 first_n = ld_row8<true>(grp);
    uint2 second_n = ld_row8<true>(grp + 32);
    for (uint32_t gi = 0; gi < ng; ++gi) {
      const uint32_t len2 = len_n;
      const uint32_t idx = __ldg(perm_p);  // only needed for the store at the end of the group
      const LaneSrcT<true> src{grp, first_n, second_n};
      // rows of this group = ceil(longest candidate / 8); the next group's rows follow immediately, so its
      // length / first rows are requested now and arrive while this group is being scored
      grp += ((__reduce_max_sync(0xffffffffu, len2) + 7u) >> 3) * 32u;
      lens_p += 32;
      perm_p += 32;
      if (gi + 1 < ng) {
        len_n = __ldg(lens_p);
        first_n = ld_row8<true>(grp);
        second_n = ld_row8<true>(grp + 32);
      }
      if constexpr (RAWDIST) {  // len1 >= 1 (checked by the launcher)
        uint32_t raw;
        if constexpr (FAM == F_LEV && sizeof(W) == 4) raw = lev_w1_u32_fast(smem_u32(pm_lane), src.reader(), len2, p.len1, p.two);
        else if constexpr (FAM == F_OSA && sizeof(W) == 4) raw = myers_w1_u32_fast<true>(smem_u32(pm_lane), src.reader(), len2, p.len1, p.two);
        else if constexpr (FAM == F_LEV && SPLIT64) raw = lev_w1_u64_fast(smem_u32(pm_lane), src.reader(), len2, p.len1, p.two);
        else if constexpr (FAM == F_OSA && SPLIT64) raw = myers_w1_u64_fast<true>(smem_u32(pm_lane), src.reader(), len2, p.len1, p.two);
        else {
          auto tab = [&](uint32_t ch) -> W { return pm_lane[ch * 32u]; };
          if constexpr (FAM == F_LEV) raw = lev_w1<W>(tab, src.reader(), len2, p.len1);
          else raw = osa_w1<W>(tab, src.reader(), len2, p.len1);
        }
        if (idx != 0xFFFFFFFFu) reinterpret_cast<uint32_t*>(p.out)[idx] = raw;
      } else {
        uint32_t ru = 0;
        double rf = 0.0;
        bool looked_up = false;
        if constexpr (FAM != F_JARO) {
          if (p.jtab != nullptr) {  // the score algebra as a table over (candidate length, raw result): int_epi_table()
            const uint32_t raw = raw_one<FAM, W, LaneSrcT<true>, SPLIT64>(pm_lane, src, len2, p.len1, p.two);
            const uint32_t l2c = len2 < p.j_l2dim ? len2 : p.j_l2dim - 1u;   // (never clamp: the dims cover the corpus)
            const uint32_t rc = raw < p.j_ccdim ? raw : p.j_ccdim - 1u;
            const uint32_t ti = l2c * p.j_ccdim + rc;
            if (p.out_f64) rf = __ldg(p.jtab + ti);
            else ru = __ldg(reinterpret_cast<const uint32_t*>(p.jtab) + ti);
            looked_up = true;
          }
        }
        if (!looked_up) score_one<FAM, W, LaneSrcT<true>, SPLIT64>(pm_lane, src, len2, p.len1, p.epi, p.out_f64, p.two, ru, rf);
        if (idx != 0xFFFFFFFFu) {
          if (p.out_f64) reinterpret_cast<double*>(p.out)[idx] = rf;
          else reinterpret_cast<uint32_t*>(p.out)[idx] = ru;
        }
      }
    }
    chunk = __shfl_sync(0xffffffffu, next_chunk, 0);
  }
}

// The score algebra of the integer metrics (MetricUsize: distance <-> similarity, normalisation, cutoff conversions and the
// final score() filter, details/distance.rs:154-275) as a per-launch TABLE over (candidate length, raw kernel result):
// with the query, kind and cutoff fixed, nothing else enters.  Built by finish_int / finish_norm themselves, so the results
// are bit-identical; the per-group epilogue (an f64 division and up to three cutoff conversions for the normalised kinds)
// becomes one IMAD and one load.  Used when every candidate is at most 255 elements long (table <= 256 x 320 entries).
struct IntTabParams {
  void* tab;
  uint32_t len1, l2dim, rawdim;
  int out_f64;
  Epi epi;
};
__global__ void __launch_bounds__(256) int_epi_table_kernel(const __grid_constant__ IntTabParams p) {
  const uint32_t total = p.l2dim * p.rawdim;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t len2 = i / p.rawdim, raw = i % p.rawdim;
    const uint32_t mx = p.len1 > len2 ? p.len1 : len2;
    const bool reachable = raw <= mx;  // distances <= max(len1, len2), LCS lengths <= min: other entries are never read
    if (p.out_f64) reinterpret_cast<double*>(p.tab)[i] = reachable ? finish_norm(p.epi, raw, p.len1, len2) : 0.0;
    else reinterpret_cast<uint32_t*>(p.tab)[i] = reachable ? finish_int(p.epi, raw, p.len1, len2) : 0u;
  }
}
constexpr uint64_t kIntTabMaxLen = 255;      // longest candidate of a corpus scored through the table
constexpr uint64_t kIntTabMinGroups = 2048;  // below 65 536 candidates the extra launch costs more than it saves
static cudaError_t int_epi_table(const ScanLaunch& L, LbParams& p, void** tab_out) {
  *tab_out = nullptr;
  if (L.corpus.max_len == 0 || L.corpus.max_len > kIntTabMaxLen || L.lb.ngroups < kIntTabMinGroups || !L.epi_table) return cudaSuccess;
  IntTabParams t{};
  t.len1 = L.query.len1;
  t.l2dim = (uint32_t)L.corpus.max_len + 1u;
  t.rawdim = (t.len1 > (uint32_t)L.corpus.max_len ? t.len1 : (uint32_t)L.corpus.max_len) + 1u;
  t.out_f64 = L.out_is_f64;
  t.epi = L.epi;
  const uint64_t total = (uint64_t)t.l2dim * t.rawdim;
  cudaError_t e = dev_alloc(&t.tab, total * (t.out_f64 ? sizeof(double) : sizeof(uint32_t)), L.stream);
  if (e != cudaSuccess) return e;
  uint64_t blocks = (total + 255) / 256;
  if (blocks > (uint64_t)L.sm_count * 8) blocks = (uint64_t)L.sm_count * 8;
  int_epi_table_kernel<<<(uint32_t)blocks, 256, 0, L.stream>>>(t);
  g_launches.fetch_add(1);
  e = cudaGetLastError();
  if (e != cudaSuccess) { dev_free(t.tab, L.stream); return e; }
  p.jtab = reinterpret_cast<const double*>(t.tab);
  p.j_l2dim = t.l2dim;
  p.j_ccdim = t.rawdim;
  *tab_out = t.tab;
  return cudaSuccess;
}

template <int FAM, class W, int NT, bool RAWDIST = false>
static cudaError_t launch_lb_inst(const ScanLaunch& L, const void* tab) {
  auto kern = scan_lb_kernel<FAM, W, NT, RAWDIST>;
  const size_t smem = sizeof(W) * 256 * 32;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // the rows are streamed (evict-first) and never re-read: give the whole unified array to shared memory so
  // that one more CTA (and its table copy) fits per SM
  e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  int ctas_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, NT, smem);
  if (e != cudaSuccess) return e;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  LbParams p{};
  p.lb = L.lb;
  p.tab = tab;
  p.len1 = L.query.len1;
  p.out = L.out;
  p.out_f64 = L.out_is_f64;
  p.two = 2;
  p.chunk = 16;
  p.counter = L.lb_counter;
  p.epi = L.epi;
  e = cudaMemsetAsync(p.counter, 0, sizeof(unsigned long long), L.stream);
  if (e != cudaSuccess) return e;
  uint64_t grid = (uint64_t)L.sm_count * ctas_per_sm;
  const uint64_t nchunks = (L.lb.ngroups + p.chunk - 1) / p.chunk;
  const uint64_t need = (nchunks + NT / 32 - 1) / (NT / 32);
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  void* etab = nullptr;
  if constexpr (!RAWDIST && FAM != F_JARO) {
    if ((e = int_epi_table(L, p, &etab)) != cudaSuccess) return e;
  }
  kern<<<(uint32_t)grid, NT, smem, L.stream>>>(p);
  g_launches.fetch_add(1);
  e = cudaGetLastError();
  dev_free(etab, L.stream);
  return e;
}

// ------------------------------------------------------------------------------------------------ jaro32
// Jaro / Jaro-Winkler with a query of at most 32 elements over the interleaved layout (BASELINE config 4).
// Thread per candidate, warp per group like scan_lb_kernel, but both passes (flagging and transposition
// counting, rf_core.cuh jaro32_rows) walk the group's 8-byte rows with warp-uniform control flow: 32-bit
// pattern flags and window masks, no per-character byte loads, no data-dependent loops.  Groups whose
// (truncated) candidates exceed 64 characters take the generic per-lane routine.
// Device form of rf_core.cuh jaro32_rows (same results; the host form is the CPU-tested specification): table
// addresses by IDP.4A, window masks advanced on the FMA pipe, per-lane predicates only in the rows that need them.
//   LO: 0 = no character of the row is past the window radius (lo stays), 1 = all are (lo shifts every step),
//       2 = mixed (per-character test);  LENP: the row may extend past this lane's candidate.
struct Jaro32Dev {
  uint32_t P, win, M;  // win = the window mask hi & lo of jaro32_rows: both shift by one per character, so
                       // win' = 2*win + (j < bound) -- one FMA-pipe op, and the match mask is ONE LOP3 (X & win & ~P)
  uint64_t T;
  template <int K>
  __device__ __forceinline__ uint32_t look(uint32_t w, uint32_t pm_lane_saddr) const {
    uint32_t X;
    const uint32_t addr = __dp4a(w, 0x80u << (8 * K), pm_lane_saddr);
    asm("ld.shared.u32 %0, [%1];" : "=r"(X) : "r"(addr));
    return X;
  }
  // rb = max(bound - j0, 0) for the row (LO == 2 only): character K of the row is before the radius iff K < rb
  template <int LO, bool LENP, int K>
  __device__ __forceinline__ void flag_step(uint32_t w, uint32_t j, uint32_t len2, uint32_t rb, uint32_t pm_lane_saddr,
                                            uint32_t two, uint32_t& t8) {
    const uint32_t X = look<(K & 3)>(w, pm_lane_saddr);
    uint32_t m;
    asm("lop3.b32 %0, %1, %2, %3, 0x40;" : "=r"(m) : "r"(X), "r"(win), "r"(P));  // X & win & ~P
    if (LENP) m = (j < len2) ? m : 0u;
    P |= m & (0u - m);
    // row-local text flags, first character in the top bit: t8 = t8 * 2 + (m != 0).  The flag is the carry of
    // m + 0xFFFFFFFF (IADD3) consumed by an IMAD.X -- one ALU-pipe op instead of a predicate-setting LOP3 + SEL.
    asm("{\n\t.reg .u32 d;\n\tadd.cc.u32 d, %1, 0xFFFFFFFF;\n\tmadc.lo.u32 %0, %0, %2, 0;\n\t}" : "+r"(t8) : "r"(m), "r"(two));
    if (LO == 0) win = win * two + (two >> 1);
    if (LO == 1) win = win * two;
    if (LO == 2)  // carry of rb + (2^32 - 1 - K) is set iff K < rb
      asm("{\n\t.reg .u32 d;\n\tadd.cc.u32 d, %1, %3;\n\tmadc.lo.u32 %0, %0, %2, 0;\n\t}" : "+r"(win) : "r"(rb), "r"(two), "n"(0xFFFFFFFFu - (uint32_t)K));
  }
  template <int LO, bool LENP>
  __device__ __forceinline__ void flag_row(uint2 v, uint32_t r, uint32_t len2, uint32_t bound, uint32_t pm_lane_saddr,
                                           uint32_t two) {
    uint32_t t8 = 0;
    const uint32_t j0 = r * 8u;
    const uint32_t rb = bound > j0 ? bound - j0 : 0u;
    flag_step<LO, LENP, 0>(v.x, j0 + 0, len2, rb, pm_lane_saddr, two, t8);
    flag_step<LO, LENP, 1>(v.x, j0 + 1, len2, rb, pm_lane_saddr, two, t8);
    flag_step<LO, LENP, 2>(v.x, j0 + 2, len2, rb, pm_lane_saddr, two, t8);
    flag_step<LO, LENP, 3>(v.x, j0 + 3, len2, rb, pm_lane_saddr, two, t8);
    flag_step<LO, LENP, 4>(v.y, j0 + 4, len2, rb, pm_lane_saddr, two, t8);
    flag_step<LO, LENP, 5>(v.y, j0 + 5, len2, rb, pm_lane_saddr, two, t8);
    flag_step<LO, LENP, 6>(v.y, j0 + 6, len2, rb, pm_lane_saddr, two, t8);
    flag_step<LO, LENP, 7>(v.y, j0 + 7, len2, rb, pm_lane_saddr, two, t8);
    T = (T << 8) | t8;  // rows pile up from the bottom byte; align_T() moves row 0 to the top byte
  }
  __device__ __forceinline__ void align_T(uint32_t nrows) { T = nrows ? T << (8u * (8u - nrows)) : 0ull; }
  // `tt` holds the INVERTED text flags of the row, current character in bit 31.  A flagged character pairs with the
  // lowest remaining pattern flag: x = P - flagged (add.cc shifts the flag out into the carry, the IMAD.X computes
  // P + 0xFFFFFFFF + !flagged), P & ~x is that pattern flag (or 0), P &= x retires it.
  template <int K>
  __device__ __forceinline__ void trans_step(uint32_t w, uint32_t& tt, uint32_t pm_lane_saddr, uint32_t one) {
    const uint32_t X = look<K>(w, pm_lane_saddr);
    uint32_t x;
    asm("{\n\tadd.cc.u32 %1, %1, %1;\n\tmadc.lo.u32 %0, %2, %3, 0xFFFFFFFF;\n\t}" : "=r"(x), "+r"(tt) : "r"(P), "r"(one));
    M |= P & ~x & ~X;  // pattern flags whose partner differs (one bit per transposed pair member)
    P &= x;
  }
  __device__ __forceinline__ void trans_row(uint2 v, uint32_t pm_lane_saddr, uint32_t one) {
    uint32_t tt = ~(uint32_t)(T >> 32) & 0xFF000000u;
    T <<= 8;
    trans_step<0>(v.x, tt, pm_lane_saddr, one);
    trans_step<1>(v.x, tt, pm_lane_saddr, one);
    trans_step<2>(v.x, tt, pm_lane_saddr, one);
    trans_step<3>(v.x, tt, pm_lane_saddr, one);
    trans_step<0>(v.y, tt, pm_lane_saddr, one);
    trans_step<1>(v.y, tt, pm_lane_saddr, one);
    trans_step<2>(v.y, tt, pm_lane_saddr, one);
    trans_step<3>(v.y, tt, pm_lane_saddr, one);
  }
};

// Groups whose (truncated) candidates exceed 64 characters: the generic per-lane routine, kept out of line so that its
// local arrays and register pressure stay out of the hot kernel body.
__device__ __noinline__ double jaro32_long_fallback(const uint32_t* __restrict__ pm_lane, const uint2* __restrict__ col,
                                                    uint32_t len1, uint32_t len2, const Epi& epi) {
  const LaneSrcT<false> src{col, __ldg(col), __ldg(col + 32)};
  auto tab64 = [&](uint32_t ch) -> uint64_t { return (uint64_t)pm_lane[ch * 32u]; };
  auto bytes = [&](uint32_t j) -> uint32_t { return src.byte(j); };
  auto jaro = [&](double c) { return jaro_similarity_w1(tab64, bytes, len1, len2, c); };
  if (epi.metric == M_JARO) return finish_float(epi, jaro);
  uint32_t prefix = 0;
  while (prefix < 4 && prefix < len1 && prefix < len2 && ((tab64(bytes(prefix)) >> prefix) & 1u)) ++prefix;
  const double pw = epi.prefix_weight;
  auto jw = [&](double c) { return jaro_winkler_from(jaro, prefix, pw, c); };
  return finish_float(epi, jw);
}

// The f64 epilogue of the row-wise Jaro kernels as a TABLE.  With the query, the kind, the cutoff and the prefix weight
// fixed for a launch, the result of a pair is a pure function of four small integers: the candidate's length, the number
// of common characters, the halved transposition count and (Jaro-Winkler) the common prefix.  The score algebra -- two
// filters, the similarity formula, the Winkler bonus, the Metricf64 chain of the four kinds, the final cutoff test:
// 150-250 executed instructions per group of 32 candidates, a quarter of them double precision, and 1700 instructions
// of code for the four kinds -- is evaluated once per launch for every reachable combination by THE SAME device
// functions (so the results stay bit-identical) and the scan kernel's epilogue becomes three IMADs and one load.
//   index = ((len2 * ccdim + cc) * trdim + transpositions / 2) * pfdim + prefix
// 1 x 1 pairs: jaro32_finish answers from `first_match`; the scan kernel stores that flag in cc for them.
struct JaroTabParams {
  double* tab;
  const double* quot;
  uint32_t len1, l2dim, ccdim, trdim, pfdim;
  Epi epi;
};
__global__ void __launch_bounds__(256) jaro_epi_table_kernel(const __grid_constant__ JaroTabParams p) {
  const uint32_t total = p.l2dim * p.ccdim * p.trdim * p.pfdim;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    uint32_t r = i;
    const uint32_t prefix = r % p.pfdim; r /= p.pfdim;
    const uint32_t th = r % p.trdim; r /= p.trdim;
    const uint32_t cc = r % p.ccdim;
    const uint32_t len2 = r / p.ccdim;
    double res = 0.0;
    // combinations no pair can produce (more common characters than characters, ...) are never looked up
    if (cc <= len2 && cc <= p.len1 && 2u * th <= cc && prefix <= len2 && prefix <= p.len1) {
      Jaro32Result jr;
      jr.cc = cc;
      jr.transpositions = 2u * th;
      const bool fm = cc != 0u;
      const uint32_t len1 = p.len1;
      const double* quot = p.quot;
      auto jaro = [&](double c) { return jaro32_finish(len1, len2, jr, fm, c, quot); };
      if (p.epi.metric == M_JARO) {
        res = finish_float(p.epi, jaro);
      } else {
        const double pw = p.epi.prefix_weight;
        auto jw = [&](double c) { return jaro_winkler_from(jaro, prefix, pw, c); };
        res = finish_float(p.epi, jw);
      }
    }
    p.tab[i] = res;
  }
}
// fills p.jtab / p.j_*dim; the caller frees *tab_out (stream-ordered) after its scan kernel is enqueued
static cudaError_t jaro_epi_table(const ScanLaunch& L, LbParams& p, double** tab_out) {
  const uint32_t len1 = L.query.len1;  // 1..64
  // longest ORIGINAL candidate the row-wise kernels score: truncated length len1 + len2/2 - 1 <= 64 (jaro.rs:553-565)
  const uint32_t l2max = (131u > 2u * len1 && 131u - 2u * len1 > 64u) ? 131u - 2u * len1 : 64u;
  JaroTabParams t{};
  t.len1 = len1;
  t.l2dim = l2max + 1u;
  t.ccdim = len1 + 1u;
  t.trdim = len1 / 2u + 1u;
  t.pfdim = L.epi.metric == M_JARO ? 1u : 5u;
  t.quot = L.query.quot;
  t.epi = L.epi;
  const uint64_t total = (uint64_t)t.l2dim * t.ccdim * t.trdim * t.pfdim;  // <= 66 * 65 * 33 * 5 = 707 850 entries (5.7 MB)
  cudaError_t e = dev_alloc(&t.tab, total * sizeof(double), L.stream);
  if (e != cudaSuccess) return e;
  uint64_t blocks = (total + 255) / 256;
  if (blocks > (uint64_t)L.sm_count * 8) blocks = (uint64_t)L.sm_count * 8;
  jaro_epi_table_kernel<<<(uint32_t)blocks, 256, 0, L.stream>>>(t);
  g_launches.fetch_add(1);
  e = cudaGetLastError();
  if (e != cudaSuccess) { dev_free(t.tab, L.stream); return e; }
  p.jtab = t.tab;
  p.j_l2dim = t.l2dim;
  p.j_ccdim = t.ccdim;
  p.j_trdim = t.trdim;
  p.j_pfdim = t.pfdim;
  *tab_out = t.tab;
  return cudaSuccess;
}
// the table lookup of a scored pair (len2 = the candidate's ORIGINAL length; the clamp never acts, see jaro_epi_table)
__device__ __forceinline__ double jaro_tab_lookup(const LbParams& p, uint32_t len2, uint32_t cc, uint32_t tr, uint32_t prefix) {
  const uint32_t l2c = len2 < p.j_l2dim ? len2 : p.j_l2dim - 1u;
  return __ldg(p.jtab + (((l2c * p.j_ccdim + cc) * p.j_trdim + (tr >> 1)) * p.j_pfdim + prefix));
}

template <int NT, bool TABEPI, int MINB = 4>
__global__ void __launch_bounds__(NT, MINB) scan_jaro32_kernel(const __grid_constant__ LbParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* pm = reinterpret_cast<uint32_t*>(smem_raw);
  {
    const uint32_t* __restrict__ t = reinterpret_cast<const uint32_t*>(p.tab);
    for (uint32_t i = threadIdx.x; i < 256u * 32u; i += NT) pm[i] = t[i >> 5];
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t* __restrict__ pm_lane = pm + lane;
  const bool zero_ok = pm[0] == 0u;  // PM[0]: no query element is the zero byte
  const uint64_t total_warps = (uint64_t)gridDim.x * (NT / 32);
  const uint64_t ngroups = p.lb.ngroups;
  const uint64_t nchunks = (ngroups + p.chunk - 1) / p.chunk;
  const uint2* __restrict__ gdata = reinterpret_cast<const uint2*>(p.lb.gdata);
  auto tab = [&](uint32_t ch) -> uint32_t { return pm_lane[ch * 32u]; };
  uint64_t chunk = (uint64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  while (chunk < nchunks) {
    unsigned long long next_chunk = 0;
    if (lane == 0) next_chunk = total_warps + atomicAdd(p.counter, 1ull);
    const uint64_t g0 = chunk * p.chunk;
    const uint64_t g1 = (g0 + p.chunk < ngroups) ? g0 + p.chunk : ngroups;
    uint64_t r = __ldg(p.lb.goff + g0);
    uint32_t len_n = __ldg(p.lb.lens + g0 * 32 + lane);
    uint32_t idx_n = __ldg(p.lb.perm + g0 * 32 + lane);
    uint2 first_n = __ldg(gdata + r * 32 + lane);
    for (uint64_t g = g0; g < g1; ++g) {
      const uint32_t len2 = len_n, idx = idx_n;
      const uint2 first = first_n;
      const uint2* col = gdata + r * 32 + lane;
      r += (__reduce_max_sync(0xffffffffu, len2) + 7u) >> 3;
      if (g + 1 < g1) {  // next group: metadata + first row into registers, its other rows into L2
        len_n = __ldg(p.lb.lens + (g + 1) * 32 + lane);
        idx_n = __ldg(p.lb.perm + (g + 1) * 32 + lane);
        const uint2* ncol = gdata + r * 32 + lane;
        first_n = __ldg(ncol);
#pragma unroll
        for (int k = 1; k < 8; ++k) prefetch_l2(ncol + k * 32);
      }
      uint32_t l1e = p.len1, l2e = len2, bound = 0;
      jaro_bounds(l1e, l2e, bound);
      const uint32_t l2max = __reduce_max_sync(0xffffffffu, l2e);
      double res;
      if (l2max <= 64) {
        const uint32_t pm_lane_saddr = smem_u32(pm_lane);
        const uint32_t nrows = (l2max + 7u) >> 3;
        const uint32_t l2min = __reduce_min_sync(0xffffffffu, l2e);
        // (a radius of 64 or more is never reached here; it also tames the wrapped radius of 1 x 1 / empty / padding lanes)
        const uint32_t bcl = bound < 64u ? bound : 64u;
        const uint32_t bmin = __reduce_min_sync(0xffffffffu, bcl), bmax = __reduce_max_sync(0xffffffffu, bcl);
        Jaro32Dev J;
        J.P = 0; J.M = 0; J.T = 0;
        J.win = (bound + 1 < 32) ? ((1u << (bound + 1)) - 1u) : 0xFFFFFFFFu;
        // pass 1: flags.  The rows fall into at most four warp-uniform segments, each with its own specialised
        // row body: before every lane's window radius, straddling it, past it, and the ragged tail.
        // rows inside every lane's candidate.  The layout pads candidates with zero bytes: when the query holds no zero
        // byte (PM[0] == 0) a padding character matches nothing, flags nothing and needs no length test at all
        const uint32_t rF = zero_ok ? nrows : (l2min >> 3);
        const uint32_t rA = (bmin < l2min ? bmin : l2min) >> 3;              // ... and before every lane's radius
        uint32_t rB = (bmax + 7u) >> 3;                                      // first row past every lane's radius
        rB = rB < rF ? rB : rF;
        rB = rB > rA ? rB : rA;
        uint2 v = first;
        uint32_t rr = 0;
#define RF_JROW(LO, LENP)                                                   \
  {                                                                         \
    const uint2 cur = v;                                                    \
    if (rr + 1 < nrows) v = __ldg(col + (rr + 1) * 32u);                    \
    J.flag_row<LO, LENP>(cur, rr, l2e, bound, pm_lane_saddr, p.two);        \
  }
        for (; rr < rA; ++rr) RF_JROW(0, false)
        for (; rr < rB; ++rr) RF_JROW(2, false)
        for (; rr < rF; ++rr) RF_JROW(1, false)
        for (; rr < nrows; ++rr) RF_JROW(2, true)
#undef RF_JROW
        J.align_T(nrows);
        Jaro32Result jr;
        jr.cc = (uint32_t)__popc(J.P);
        v = first;
        for (rr = 0; rr < nrows; ++rr) {  // pass 2: transpositions (rows now come from L1/L2)
          const uint2 cur = v;
          if (rr + 1 < nrows) v = __ldg(col + (rr + 1) * 32u);
          J.trans_row(cur, pm_lane_saddr, p.two >> 1);
        }
        jr.transpositions = (uint32_t)__popc(J.M);
        const uint32_t w0 = len2 ? first.x : 0u;
        const uint32_t len1 = p.len1;
        if constexpr (TABEPI) {
          uint32_t prefix = 0;  // common prefix, at most 4 (jaro_winkler.rs:118-123)
          if (p.j_pfdim > 1u) {
            const uint32_t lim = len1 < len2 ? (len1 < 4 ? len1 : 4) : (len2 < 4 ? len2 : 4);
            while (prefix < lim && ((tab((w0 >> (8 * prefix)) & 0xffu) >> prefix) & 1u)) ++prefix;
          }
          // 1 x 1: the wrapped radius leaves the window empty and the reference answers from the first characters
          // (jaro.rs:546-548); the table holds that answer under cc = 1 / cc = 0
          if (len1 == 1u && len2 == 1u) jr.cc = tab(w0 & 0xffu) & 1u;
          res = jaro_tab_lookup(p, len2, jr.cc, jr.transpositions, prefix);
        } else {
          const bool fm = len2 && (tab(w0 & 0xffu) & 1u);
          const double* quot = p.quot;
          auto jaro = [&](double c) { return jaro32_finish(len1, len2, jr, fm, c, quot); };
          if (p.epi.metric == M_JARO) {
            res = finish_float(p.epi, jaro);
          } else {
            uint32_t prefix = 0;
            const uint32_t lim = len1 < len2 ? (len1 < 4 ? len1 : 4) : (len2 < 4 ? len2 : 4);
            while (prefix < lim && ((tab((w0 >> (8 * prefix)) & 0xffu) >> prefix) & 1u)) ++prefix;
            const double pw = p.epi.prefix_weight;
            auto jw = [&](double c) { return jaro_winkler_from(jaro, prefix, pw, c); };
            res = finish_float(p.epi, jw);
          }
        }
      } else {
        // (truncated) candidates longer than 64 characters: left to jaro32_long_kernel, which runs right after this
        // kernel and exits at once when no group raised the flag.  A call to the generic routine from here costs the
        // hot path 16 more registers and spills around the call site (measured: 64 registers + 112 B of stack).
        if (lane == 0) *p.flag = 1ull;
        continue;
      }
      if (idx != 0xFFFFFFFFu) reinterpret_cast<double*>(p.out)[idx] = res;
    }
    chunk = __shfl_sync(0xffffffffu, next_chunk, 0);
  }
}

// The groups scan_jaro32_kernel skipped: the generic per-lane routine (any candidate length).
__global__ void __launch_bounds__(256) jaro32_long_kernel(const __grid_constant__ LbParams p) {
  if (*reinterpret_cast<const volatile unsigned long long*>(p.flag) == 0ull) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* pm = reinterpret_cast<uint32_t*>(smem_raw);
  {
    const uint32_t* __restrict__ t = reinterpret_cast<const uint32_t*>(p.tab);
    for (uint32_t i = threadIdx.x; i < 256u * 32u; i += 256) pm[i] = t[i >> 5];
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t* __restrict__ pm_lane = pm + lane;
  const uint2* __restrict__ gdata = reinterpret_cast<const uint2*>(p.lb.gdata);
  const uint64_t nwarps = (uint64_t)gridDim.x * 8, w0 = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  for (uint64_t g = w0; g < p.lb.ngroups; g += nwarps) {
    const uint32_t len2 = __ldg(p.lb.lens + g * 32 + lane);
    const uint32_t idx = __ldg(p.lb.perm + g * 32 + lane);
    uint32_t l1e = p.len1, l2e = len2, bound = 0;
    jaro_bounds(l1e, l2e, bound);
    if (__reduce_max_sync(0xffffffffu, l2e) <= 64) continue;  // scored by the fast kernel
    const uint2* col = gdata + __ldg(p.lb.goff + g) * 32 + lane;
    const double res = jaro32_long_fallback(pm_lane, col, p.len1, len2, p.epi);
    if (idx != 0xFFFFFFFFu) reinterpret_cast<double*>(p.out)[idx] = res;
  }
}

static cudaError_t launch_jaro32(const ScanLaunch& L) {
  constexpr int NT = 256;
  const bool tabepi = L.jaro32 != 2;  // 2 = the in-kernel f64 epilogue (kept for A/B runs and as a cross-check in the tests)
  // without the f64 epilogue the kernel also fits 48 registers = 5 CTAs per SM instead of 4 (jaro32 = 3): measured 3.17 ms per
  // 10^8 candidates against 3.14 ms for the 64-register build, so 4 CTAs stay the default
  auto kern = !tabepi ? scan_jaro32_kernel<NT, false> : (L.jaro32 == 3 ? scan_jaro32_kernel<NT, true, 5> : scan_jaro32_kernel<NT, true, 4>);
  const size_t smem = sizeof(uint32_t) * 256 * 32;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int ctas_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, NT, smem);
  if (e != cudaSuccess) return e;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  LbParams p{};
  p.lb = L.lb;
  p.tab = L.query.tab32_bot;
  p.quot = L.query.quot;
  p.len1 = L.query.len1;
  p.out = L.out;
  p.out_f64 = 1;
  p.two = 2;
  p.chunk = 16;  // groups per scheduler grab; 4...128 measured identical (1.99 ms): neither the atomic nor the tail matter
  p.counter = L.lb_counter;
  p.flag = L.lb_flag;
  p.epi = L.epi;
  e = cudaMemsetAsync(p.counter, 0, sizeof(unsigned long long), L.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(p.flag, 0, sizeof(unsigned long long), L.stream);
  if (e != cudaSuccess) return e;
  uint64_t grid = (uint64_t)L.sm_count * ctas_per_sm;
  const uint64_t nchunks = (L.lb.ngroups + p.chunk - 1) / p.chunk;
  const uint64_t need = (nchunks + NT / 32 - 1) / (NT / 32);
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  double* jtab = nullptr;
  if (tabepi && (e = jaro_epi_table(L, p, &jtab)) != cudaSuccess) return e;
  kern<<<(uint32_t)grid, NT, smem, L.stream>>>(p);
  g_launches.fetch_add(1);
  e = cudaGetLastError();
  dev_free(jtab, L.stream);
  if (e != cudaSuccess) return e;
  // groups with candidates beyond 64 characters (none in BASELINE config 4): a no-op launch unless the flag was raised
  e = cudaFuncSetAttribute(jaro32_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  uint64_t lgrid = (uint64_t)L.sm_count * 2;
  const uint64_t lneed = (L.lb.ngroups + 7) / 8;
  if (lgrid > lneed) lgrid = lneed;
  jaro32_long_kernel<<<(uint32_t)lgrid, 256, smem, L.stream>>>(p);
  g_launches.fetch_add(1);
  return cudaGetLastError();
}

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
  return d;
}
// x * 2 as {low word, bit that falls out}: IMAD.WIDE.U32 (`two` is opaque to ptxas, so it stays on the FMA pipe)
__device__ __forceinline__ void mul2_wide(uint32_t x, uint32_t two, uint32_t& lo, uint32_t& hi) {
  asm("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(x), "r"(two));
}

// ------------------------------------------------------------------------------------------------ jaro64
// The same row-wise scheme for queries of 33..64 elements (they used to take the per-lane generic routine: 8.5 ms per
// 10^8 candidates against 3.43 ms for queries up to 32): pattern flags, window mask and transposition marks are 64-bit,
// held as 32-bit halves with the carries on IADD3.X / IMAD.WIDE like the 64-bit Levenshtein recurrence; the match table is
// split into a low-word and a high-word copy 32 KB apart (one IDP.4A address, two LDS).  Candidates (truncated) beyond 64
// characters go to jaro64_long_kernel as before.
struct Jaro64Dev {
  uint32_t Pl, Ph, wl, wh, Ml, Mh;
  uint64_t T;
  template <int K>
  __device__ __forceinline__ void look(uint32_t w, uint32_t pm_lane_saddr, uint32_t& Xl, uint32_t& Xh) const {
    const uint32_t addr = __dp4a(w, 0x80u << (8 * K), pm_lane_saddr);
    asm("ld.shared.u32 %0, [%1];" : "=r"(Xl) : "r"(addr));
    asm("ld.shared.u32 %0, [%1+32768];" : "=r"(Xh) : "r"(addr));
  }
  template <int LO, bool LENP, int K>
  __device__ __forceinline__ void flag_step(uint32_t w, uint32_t j, uint32_t len2, uint32_t rb, uint32_t pm_lane_saddr,
                                            uint32_t two, uint32_t& t8) {
    uint32_t Xl, Xh;
    look<(K & 3)>(w, pm_lane_saddr, Xl, Xh);
    uint32_t ml = lop3<0x40>(Xl, wl, Pl), mh = lop3<0x40>(Xh, wh, Ph);  // X & win & ~P
    if (LENP) { ml = (j < len2) ? ml : 0u; mh = (j < len2) ? mh : 0u; }
    // lowest set bit of the 64-bit m: the low word's if it has one, else the high word's.  hmask = all ones iff ml == 0
    // (0xFFFFFFFF + carry of ml + 0xFFFFFFFF)
    uint32_t hmask;
    asm("{\n\t.reg .u32 d;\n\tadd.cc.u32 d, %1, 0xFFFFFFFF;\n\taddc.u32 %0, 0xFFFFFFFF, 0;\n\t}" : "=r"(hmask) : "r"(ml));
    Pl |= ml & (0u - ml);
    Ph |= mh & (0u - mh) & hmask;
    const uint32_t any = ml | mh;
    asm("{\n\t.reg .u32 d;\n\tadd.cc.u32 d, %1, 0xFFFFFFFF;\n\tmadc.lo.u32 %0, %0, %2, 0;\n\t}" : "+r"(t8) : "r"(any), "r"(two));
    // window: win' = 2 * win + (j < bound) on 64 bits
    uint32_t lo2, c;
    mul2_wide(wl, two, lo2, c);
    wh = wh * two + c;
    if (LO == 0) wl = lo2 + (two >> 1);
    if (LO == 1) wl = lo2;
    if (LO == 2) {  // carry of rb + (2^32 - 1 - K) is set iff K < rb
      uint32_t inc;
      asm("{\n\t.reg .u32 d;\n\tadd.cc.u32 d, %1, %2;\n\taddc.u32 %0, 0, 0;\n\t}" : "=r"(inc) : "r"(rb), "n"(0xFFFFFFFFu - (uint32_t)K));
      wl = lo2 + inc;
    }
  }
  template <int LO, bool LENP>
  __device__ __forceinline__ void flag_row(uint2 v, uint32_t r, uint32_t len2, uint32_t bound, uint32_t pm_lane_saddr, uint32_t two) {
    uint32_t t8 = 0;
    const uint32_t j0 = r * 8u;
    const uint32_t rb = bound > j0 ? bound - j0 : 0u;
    flag_step<LO, LENP, 0>(v.x, j0 + 0, len2, rb, pm_lane_saddr, two, t8);
    flag_step<LO, LENP, 1>(v.x, j0 + 1, len2, rb, pm_lane_saddr, two, t8);
    flag_step<LO, LENP, 2>(v.x, j0 + 2, len2, rb, pm_lane_saddr, two, t8);
    flag_step<LO, LENP, 3>(v.x, j0 + 3, len2, rb, pm_lane_saddr, two, t8);
    flag_step<LO, LENP, 4>(v.y, j0 + 4, len2, rb, pm_lane_saddr, two, t8);
    flag_step<LO, LENP, 5>(v.y, j0 + 5, len2, rb, pm_lane_saddr, two, t8);
    flag_step<LO, LENP, 6>(v.y, j0 + 6, len2, rb, pm_lane_saddr, two, t8);
    flag_step<LO, LENP, 7>(v.y, j0 + 7, len2, rb, pm_lane_saddr, two, t8);
    T = (T << 8) | t8;
  }
  __device__ __forceinline__ void align_T(uint32_t nrows) { T = nrows ? T << (8u * (8u - nrows)) : 0ull; }
  // x = P - flagged on 64 bits: the add.cc shifts the (inverted) text flag into the carry, the two addc's compute
  // P + (2^64 - 1) + !flagged; P & ~x is the lowest remaining pattern flag (or 0), P &= x retires it
  template <int K>
  __device__ __forceinline__ void trans_step(uint32_t w, uint32_t& tt, uint32_t pm_lane_saddr) {
    uint32_t Xl, Xh, xl, xh;
    look<K>(w, pm_lane_saddr, Xl, Xh);
    asm("{\n\tadd.cc.u32 %2, %2, %2;\n\taddc.cc.u32 %0, %3, 0xFFFFFFFF;\n\taddc.u32 %1, %4, 0xFFFFFFFF;\n\t}"
        : "=r"(xl), "=r"(xh), "+r"(tt) : "r"(Pl), "r"(Ph));
    Ml |= lop3<0x02>(xl, Xl, Pl);  // ~x & ~X & P
    Mh |= lop3<0x02>(xh, Xh, Ph);
    Pl &= xl;
    Ph &= xh;
  }
  __device__ __forceinline__ void trans_row(uint2 v, uint32_t pm_lane_saddr) {
    uint32_t tt = ~(uint32_t)(T >> 32) & 0xFF000000u;
    T <<= 8;
    trans_step<0>(v.x, tt, pm_lane_saddr);
    trans_step<1>(v.x, tt, pm_lane_saddr);
    trans_step<2>(v.x, tt, pm_lane_saddr);
    trans_step<3>(v.x, tt, pm_lane_saddr);
    trans_step<0>(v.y, tt, pm_lane_saddr);
    trans_step<1>(v.y, tt, pm_lane_saddr);
    trans_step<2>(v.y, tt, pm_lane_saddr);
    trans_step<3>(v.y, tt, pm_lane_saddr);
  }
};

__device__ __noinline__ double jaro64_long_fallback(const uint32_t* __restrict__ pm32_lane, const uint2* __restrict__ col,
                                                    uint32_t len1, uint32_t len2, const Epi& epi) {
  const LaneSrcT<false> src{col, __ldg(col), __ldg(col + 32)};
  auto tab64 = [&](uint32_t ch) -> uint64_t { return (uint64_t)pm32_lane[ch * 32u] | ((uint64_t)pm32_lane[8192u + ch * 32u] << 32); };
  auto bytes = [&](uint32_t j) -> uint32_t { return src.byte(j); };
  auto jaro = [&](double c) { return jaro_similarity_w1(tab64, bytes, len1, len2, c); };
  if (epi.metric == M_JARO) return finish_float(epi, jaro);
  uint32_t prefix = 0;
  while (prefix < 4 && prefix < len1 && prefix < len2 && ((tab64(bytes(prefix)) >> prefix) & 1u)) ++prefix;
  const double pw = epi.prefix_weight;
  auto jw = [&](double c) { return jaro_winkler_from(jaro, prefix, pw, c); };
  return finish_float(epi, jw);
}

template <int NT, bool TABEPI>
__global__ void __launch_bounds__(NT) scan_jaro64_kernel(const __grid_constant__ LbParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* pm32 = reinterpret_cast<uint32_t*>(smem_raw);
  {
    const uint64_t* __restrict__ t = reinterpret_cast<const uint64_t*>(p.tab);
    for (uint32_t i = threadIdx.x; i < 256u * 32u; i += NT) {
      const uint64_t v = t[i >> 5];
      pm32[i] = (uint32_t)v;
      pm32[i + 8192u] = (uint32_t)(v >> 32);
    }
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t* __restrict__ pm_lane = pm32 + lane;
  const bool zero_ok = pm32[0] == 0u && pm32[8192] == 0u;  // PM[0]: no query element is the zero byte
  const uint64_t total_warps = (uint64_t)gridDim.x * (NT / 32);
  const uint64_t ngroups = p.lb.ngroups;
  const uint64_t nchunks = (ngroups + p.chunk - 1) / p.chunk;
  const uint2* __restrict__ gdata = reinterpret_cast<const uint2*>(p.lb.gdata);
  auto tab_lo = [&](uint32_t ch) -> uint32_t { return pm_lane[ch * 32u]; };
  uint64_t chunk = (uint64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  while (chunk < nchunks) {
    unsigned long long next_chunk = 0;
    if (lane == 0) next_chunk = total_warps + atomicAdd(p.counter, 1ull);
    const uint64_t g0 = chunk * p.chunk;
    const uint64_t g1 = (g0 + p.chunk < ngroups) ? g0 + p.chunk : ngroups;
    uint64_t r = __ldg(p.lb.goff + g0);
    uint32_t len_n = __ldg(p.lb.lens + g0 * 32 + lane);
    uint32_t idx_n = __ldg(p.lb.perm + g0 * 32 + lane);
    uint2 first_n = __ldg(gdata + r * 32 + lane);
    for (uint64_t g = g0; g < g1; ++g) {
      const uint32_t len2 = len_n, idx = idx_n;
      const uint2 first = first_n;
      const uint2* col = gdata + r * 32 + lane;
      r += (__reduce_max_sync(0xffffffffu, len2) + 7u) >> 3;
      if (g + 1 < g1) {
        len_n = __ldg(p.lb.lens + (g + 1) * 32 + lane);
        idx_n = __ldg(p.lb.perm + (g + 1) * 32 + lane);
        const uint2* ncol = gdata + r * 32 + lane;
        first_n = __ldg(ncol);
#pragma unroll
        for (int k = 1; k < 8; ++k) prefetch_l2(ncol + k * 32);
      }
      uint32_t l1e = p.len1, l2e = len2, bound = 0;
      jaro_bounds(l1e, l2e, bound);
      const uint32_t l2max = __reduce_max_sync(0xffffffffu, l2e);
      double res;
      if (l2max <= 64) {
        const uint32_t pm_lane_saddr = smem_u32(pm_lane);
        const uint32_t nrows = (l2max + 7u) >> 3;
        const uint32_t l2min = __reduce_min_sync(0xffffffffu, l2e);
        const uint32_t bcl = bound < 64u ? bound : 64u;
        const uint32_t bmin = __reduce_min_sync(0xffffffffu, bcl), bmax = __reduce_max_sync(0xffffffffu, bcl);
        Jaro64Dev J;
        J.Pl = J.Ph = J.Ml = J.Mh = 0;
        J.T = 0;
        {  // bits 0 .. bound of the window (a wrapped radius -- 1 x 1 / empty / padding lanes -- opens everything)
          const uint64_t w0 = (bound + 1 < 64) ? ((1ull << (bound + 1)) - 1ull) : ~0ull;
          J.wl = (uint32_t)w0;
          J.wh = (uint32_t)(w0 >> 32);
        }
        const uint32_t rF = zero_ok ? nrows : (l2min >> 3);
        const uint32_t rA = (bmin < l2min ? bmin : l2min) >> 3;
        uint32_t rB = (bmax + 7u) >> 3;
        rB = rB < rF ? rB : rF;
        rB = rB > rA ? rB : rA;
        uint2 v = first;
        uint32_t rr = 0;
#define RF_JROW64(LO, LENP)                                                 \
  {                                                                         \
    const uint2 cur = v;                                                    \
    if (rr + 1 < nrows) v = __ldg(col + (rr + 1) * 32u);                    \
    J.flag_row<LO, LENP>(cur, rr, l2e, bound, pm_lane_saddr, p.two);        \
  }
        for (; rr < rA; ++rr) RF_JROW64(0, false)
        for (; rr < rB; ++rr) RF_JROW64(2, false)
        for (; rr < rF; ++rr) RF_JROW64(1, false)
        for (; rr < nrows; ++rr) RF_JROW64(2, true)
#undef RF_JROW64
        J.align_T(nrows);
        Jaro32Result jr;
        jr.cc = (uint32_t)__popc(J.Pl) + (uint32_t)__popc(J.Ph);
        v = first;
        for (rr = 0; rr < nrows; ++rr) {
          const uint2 cur = v;
          if (rr + 1 < nrows) v = __ldg(col + (rr + 1) * 32u);
          J.trans_row(cur, pm_lane_saddr);
        }
        jr.transpositions = (uint32_t)__popc(J.Ml) + (uint32_t)__popc(J.Mh);
        const uint32_t w0 = len2 ? first.x : 0u;
        const uint32_t len1 = p.len1;
        if constexpr (TABEPI) {
          uint32_t prefix = 0;
          if (p.j_pfdim > 1u) {
            const uint32_t lim = len1 < len2 ? (len1 < 4 ? len1 : 4) : (len2 < 4 ? len2 : 4);
            while (prefix < lim && ((tab_lo((w0 >> (8 * prefix)) & 0xffu) >> prefix) & 1u)) ++prefix;
          }
          res = jaro_tab_lookup(p, len2, jr.cc, jr.transpositions, prefix);
        } else {
          const bool fm = len2 && (tab_lo(w0 & 0xffu) & 1u);
          const double* quot = p.quot;
          auto jaro = [&](double c) { return jaro32_finish(len1, len2, jr, fm, c, quot); };
          if (p.epi.metric == M_JARO) {
            res = finish_float(p.epi, jaro);
          } else {
            uint32_t prefix = 0;
            const uint32_t lim = len1 < len2 ? (len1 < 4 ? len1 : 4) : (len2 < 4 ? len2 : 4);
            while (prefix < lim && ((tab_lo((w0 >> (8 * prefix)) & 0xffu) >> prefix) & 1u)) ++prefix;
            const double pw = p.epi.prefix_weight;
            auto jw = [&](double c) { return jaro_winkler_from(jaro, prefix, pw, c); };
            res = finish_float(p.epi, jw);
          }
        }
      } else {
        if (lane == 0) *p.flag = 1ull;
        continue;
      }
      if (idx != 0xFFFFFFFFu) reinterpret_cast<double*>(p.out)[idx] = res;
    }
    chunk = __shfl_sync(0xffffffffu, next_chunk, 0);
  }
}

__global__ void __launch_bounds__(256) jaro64_long_kernel(const __grid_constant__ LbParams p) {
  if (*reinterpret_cast<const volatile unsigned long long*>(p.flag) == 0ull) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* pm32 = reinterpret_cast<uint32_t*>(smem_raw);
  {
    const uint64_t* __restrict__ t = reinterpret_cast<const uint64_t*>(p.tab);
    for (uint32_t i = threadIdx.x; i < 256u * 32u; i += 256) {
      const uint64_t v = t[i >> 5];
      pm32[i] = (uint32_t)v;
      pm32[i + 8192u] = (uint32_t)(v >> 32);
    }
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const uint2* __restrict__ gdata = reinterpret_cast<const uint2*>(p.lb.gdata);
  const uint64_t nwarps = (uint64_t)gridDim.x * 8, w0 = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  for (uint64_t g = w0; g < p.lb.ngroups; g += nwarps) {
    const uint32_t len2 = __ldg(p.lb.lens + g * 32 + lane);
    const uint32_t idx = __ldg(p.lb.perm + g * 32 + lane);
    uint32_t l1e = p.len1, l2e = len2, bound = 0;
    jaro_bounds(l1e, l2e, bound);
    if (__reduce_max_sync(0xffffffffu, l2e) <= 64) continue;  // scored by the fast kernel
    const uint2* col = gdata + __ldg(p.lb.goff + g) * 32 + lane;
    const double res = jaro64_long_fallback(pm32 + lane, col, p.len1, len2, p.epi);
    if (idx != 0xFFFFFFFFu) reinterpret_cast<double*>(p.out)[idx] = res;
  }
}

static cudaError_t launch_jaro64(const ScanLaunch& L) {
  constexpr int NT = 256;
  const bool tabepi = L.jaro32 != 2;
  auto kern = tabepi ? scan_jaro64_kernel<NT, true> : scan_jaro64_kernel<NT, false>;
  const size_t smem = sizeof(uint64_t) * 256 * 32;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  int ctas_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, NT, smem);
  if (e != cudaSuccess) return e;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  LbParams p{};
  p.lb = L.lb;
  p.tab = L.query.tab64_bot;
  p.quot = L.query.quot;
  p.len1 = L.query.len1;
  p.out = L.out;
  p.out_f64 = 1;
  p.two = 2;
  p.chunk = 16;
  p.counter = L.lb_counter;
  p.flag = L.lb_flag;
  p.epi = L.epi;
  e = cudaMemsetAsync(p.counter, 0, sizeof(unsigned long long), L.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(p.flag, 0, sizeof(unsigned long long), L.stream);
  if (e != cudaSuccess) return e;
  uint64_t grid = (uint64_t)L.sm_count * ctas_per_sm;
  const uint64_t nchunks = (L.lb.ngroups + p.chunk - 1) / p.chunk;
  const uint64_t need = (nchunks + NT / 32 - 1) / (NT / 32);
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  double* jtab = nullptr;
  if (tabepi && (e = jaro_epi_table(L, p, &jtab)) != cudaSuccess) return e;
  kern<<<(uint32_t)grid, NT, smem, L.stream>>>(p);
  g_launches.fetch_add(1);
  e = cudaGetLastError();
  dev_free(jtab, L.stream);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(jaro64_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  uint64_t lgrid = (uint64_t)L.sm_count * 2;
  const uint64_t lneed = (L.lb.ngroups + 7) / 8;
  if (lgrid > lneed) lgrid = lneed;
  jaro64_long_kernel<<<(uint32_t)lgrid, 256, smem, L.stream>>>(p);
  g_launches.fetch_add(1);
  return cudaGetLastError();
}

cudaError_t launch_scan_lb(const ScanLaunch& L) {
  const Family fam = family_of(L.epi.metric, L.epi.wclass);
  const bool w32 = L.query.len1 <= 32;
  const bool rawdist = (fam == F_LEV || fam == F_OSA) && L.epi.unit32 && L.epi.kind == K_DISTANCE && !L.epi.has_cutoff &&
                       !L.out_is_f64 && L.query.len1 >= 1;
  switch (fam) {
    case F_LEV:
      if (rawdist)
        return w32 ? launch_lb_inst<F_LEV, uint32_t, 256, true>(L, L.query.tab32_top)
                   : launch_lb_inst<F_LEV, uint64_t, 512, true>(L, L.query.tab64_top);
      return w32 ? launch_lb_inst<F_LEV, uint32_t, 256>(L, L.query.tab32_top)
                 : launch_lb_inst<F_LEV, uint64_t, 512>(L, L.query.tab64_top);
    case F_OSA:
      if (rawdist)
        return w32 ? launch_lb_inst<F_OSA, uint32_t, 256, true>(L, L.query.tab32_top)
                   : launch_lb_inst<F_OSA, uint64_t, 512, true>(L, L.query.tab64_top);
      return w32 ? launch_lb_inst<F_OSA, uint32_t, 256>(L, L.query.tab32_top)
                 : launch_lb_inst<F_OSA, uint64_t, 512>(L, L.query.tab64_top);
    case F_LCS:
      return w32 ? launch_lb_inst<F_LCS, uint32_t, 256>(L, L.query.tab32_bot)
                 : launch_lb_inst<F_LCS, uint64_t, 512>(L, L.query.tab64_bot);
    default:
      if (L.query.len1 >= 1 && L.query.len1 <= 32 && L.jaro32) return launch_jaro32(L);
      if (L.query.len1 > 32 && L.query.len1 <= 64 && L.jaro32) return launch_jaro64(L);
      return launch_lb_inst<F_JARO, uint64_t, 512>(L, L.query.tab64_bot);
  }
}

// ------------------------------------------------------------------------------------------------ lb ring (TMA)
// Same work as scan_lb_kernel (warp per group of 32 equal-length candidates of the interleaved layout), but the
// rows reach the warp through a private shared-memory ring filled by TMA bulk copies instead of per-lane
// global loads: the groups of a chunk are one contiguous row range, so lane 0 streams it in 1 KB blocks
// (4 rows x 32 lanes x 8 bytes, cp.async.bulk + mbarrier, SASS UBLKCP), two blocks in flight per warp.  The
// look-ahead (>= 32 characters per lane) costs no registers and no per-row address arithmetic, the consumer
// reads a row with one conflict-free LDS.64.  Sequential readers only (Levenshtein, OSA, LCS family);
// Jaro needs random access to the candidate and stays on scan_lb_kernel.
// per-candidate state machines: step<K>() consumes byte K of a packed word of 4 text characters
template <class W>
struct LevSt {
  W VP, VN;
  __device__ __forceinline__ void init(uint32_t len1) { VP = (W)(~(W)0) << ((int)sizeof(W) * 8 - (int)len1); VN = 0; }
  template <int K>
  __device__ __forceinline__ void step(uint32_t w, const W* pm_lane, uint32_t, uint32_t) {
    const W X = pm_lane[((w >> (8 * K)) & 0xffu) * 32u];
    const W D0 = ((((X & VP) + VP) ^ VP) | X) | VN;
    W HP = VN | ~(D0 | VP);
    W HN = D0 & VP;
    HP = (HP << 1) | (W)1;
    HN = HN << 1;
    VP = HN | ~(D0 | HP);
    VN = HP & D0;
  }
  __device__ __forceinline__ uint32_t raw(uint32_t len2) const { return len2 + (uint32_t)popc(VP) - (uint32_t)popc(VN); }
};
// 32-bit Levenshtein with the address / shift arithmetic steered onto the FMA pipe (see lev_w1_u32_fast)
struct LevSt32Fast {
  uint32_t VP, VN;
  __device__ __forceinline__ void init(uint32_t len1) { VP = 0xFFFFFFFFu << (32u - len1); VN = 0; }
  template <int K>
  __device__ __forceinline__ void step(uint32_t w, const uint32_t*, uint32_t pm_lane_saddr, uint32_t two) {
    uint32_t X;
    const uint32_t addr = __dp4a(w, 0x80u << (8 * K), pm_lane_saddr);
    asm("ld.shared.u32 %0, [%1];" : "=r"(X) : "r"(addr));
    const uint32_t D0 = ((((X & VP) + VP) ^ VP) | X) | VN;
    uint32_t HP = VN | ~(D0 | VP);
    uint32_t HN = D0 & VP;
    HP = HP * two + (two >> 1);
    HN = HN * two;
    VP = HN | ~(D0 | HP);
    VN = HP & D0;
  }
  __device__ __forceinline__ uint32_t raw(uint32_t len2) const { return len2 + (uint32_t)__popc(VP) - (uint32_t)__popc(VN); }
};
template <class W>
struct OsaSt {
  W VP, VN, D0, PMold;
  __device__ __forceinline__ void init(uint32_t len1) { VP = (W)(~(W)0) << ((int)sizeof(W) * 8 - (int)len1); VN = 0; D0 = 0; PMold = 0; }
  template <int K>
  __device__ __forceinline__ void step(uint32_t w, const W* pm_lane, uint32_t, uint32_t) {
    const W X = pm_lane[((w >> (8 * K)) & 0xffu) * 32u];
    const W TR = (((~D0) & X) << 1) & PMold;
    D0 = (((((X & VP) + VP) ^ VP) | X) | VN) | TR;
    W HP = VN | ~(D0 | VP);
    W HN = D0 & VP;
    HP = (HP << 1) | (W)1;
    HN = HN << 1;
    VP = HN | ~(D0 | HP);
    VN = HP & D0;
    PMold = X;
  }
  __device__ __forceinline__ uint32_t raw(uint32_t len2) const { return len2 + (uint32_t)popc(VP) - (uint32_t)popc(VN); }
};
template <class W>
struct LcsSt {
  W S;
  __device__ __forceinline__ void init(uint32_t) { S = ~(W)0; }
  template <int K>
  __device__ __forceinline__ void step(uint32_t w, const W* pm_lane, uint32_t, uint32_t) {
    const W U = S & pm_lane[((w >> (8 * K)) & 0xffu) * 32u];
    S = (S + U) | (S & ~U);
  }
  __device__ __forceinline__ uint32_t raw(uint32_t) const { return (uint32_t)popc((W)~S); }
};

// BLK = rows per ring block (one TMA copy of BLK*256 bytes); the ring holds two blocks.
template <int FAM, class W, int NT, bool RAWDIST, int BLK>
__global__ void __launch_bounds__(NT) scan_lbr_kernel(const __grid_constant__ LbParams p) {
  constexpr int NW = NT / 32;
  constexpr uint32_t RING = 2 * BLK, BLK_BYTES = BLK * 256;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  W* pm = reinterpret_cast<W*>(smem_raw);
  unsigned char* rings = smem_raw + sizeof(W) * 8192;
  uint64_t* bars = reinterpret_cast<uint64_t*>(rings + NW * RING * 256);
  {
    const W* __restrict__ t = reinterpret_cast<const W*>(p.tab);
    for (uint32_t i = threadIdx.x; i < 256u * 32u; i += NT) pm[i] = t[i >> 5];
  }
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (lane == 0) {
    mbar_init(bars + warp * 2, 1);
    mbar_init(bars + warp * 2 + 1, 1);
    mbar_fence_init();
  }
  __syncthreads();
  const uint32_t ring_s = smem_u32(rings + warp * (RING * 256));
  const uint32_t bar_s = smem_u32(bars + warp * 2);
  const uint32_t ring_lane = ring_s + lane * 8u;
  const W* __restrict__ pm_lane = pm + lane;
  const uint32_t pm_lane_saddr = smem_u32(pm_lane);
  const uint64_t total_warps = (uint64_t)gridDim.x * NW;
  const uint64_t ngroups = p.lb.ngroups;
  const uint64_t nchunks = (ngroups + p.chunk - 1) / p.chunk;
  const unsigned char* __restrict__ gdata = reinterpret_cast<const unsigned char*>(p.lb.gdata);
  uint32_t Q = 0;  // ring blocks consumed so far by this warp: block q uses slot q&1 with mbarrier parity (q>>1)&1
  uint64_t chunk = (uint64_t)blockIdx.x * NW + warp;
  while (chunk < nchunks) {
    unsigned long long next_chunk = 0;
    if (lane == 0) next_chunk = total_warps + atomicAdd(p.counter, 1ull);  // latency hidden behind this chunk
    const uint64_t g0 = chunk * p.chunk;
    const uint64_t g1 = (g0 + p.chunk < ngroups) ? g0 + p.chunk : ngroups;
    // warp-uniform by construction; the shuffles say so to the compiler's uniformity analysis
    const uint64_t ra = __shfl_sync(0xffffffffu, __ldg(p.lb.goff + g0), 0);
    const uint64_t rb = __shfl_sync(0xffffffffu, __ldg(p.lb.goff + g1), 0);
    const uint32_t nblk = (uint32_t)((rb - ra + BLK - 1) / BLK);
    const unsigned char* gsrc = gdata + ra * 256;
    auto fill = [&](uint32_t b) {  // block b of this chunk -> its ring slot (whole warp calls, one lane issues)
      const uint32_t slot = (Q + b) & 1u;
      tma_bulk_g2s_elect(ring_s + slot * BLK_BYTES, gsrc + (size_t)b * BLK_BYTES, BLK_BYTES, bar_s + slot * 8u);
    };
    if (nblk > 0) fill(0);
    if (nblk > 1) fill(1);
    uint32_t len_n = __ldg(p.lb.lens + g0 * 32 + lane);
    uint32_t idx_n = __ldg(p.lb.perm + g0 * 32 + lane);
    uint32_t k = 0;  // row cursor inside the chunk
    // row kk of the chunk as (x,y) = 8 text bytes of this lane.  Entering a new block (kk % BLK == 0): every lane is
    // done with the previous block (its LDS were issued long before the copy is even requested), so that slot is
    // refilled with the block after this one; then wait for ours.
    auto row = [&](uint32_t kk) -> uint2 {
      if ((kk & (BLK - 1)) == 0) {
        const uint32_t b = kk / BLK;
        __syncwarp();
        if (b >= 1 && b + 1 < nblk) fill(b + 1);
        mbar_wait_saddr(bar_s + ((Q + b) & 1u) * 8u, ((Q + b) >> 1) & 1u);
      }
      const uint32_t a = ring_lane + (((kk + BLK * (Q & 1u)) & (RING - 1)) << 8);
      uint2 v;
      asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
      return v;
    };
    for (uint64_t g = g0; g < g1; ++g) {
      const uint32_t len2 = len_n, idx = idx_n;
      const uint32_t lmax = __reduce_max_sync(0xffffffffu, len2);
      const uint32_t lmin = __reduce_min_sync(0xffffffffu, len2);
      if (g + 1 < g1) {
        len_n = __ldg(p.lb.lens + (g + 1) * 32 + lane);
        idx_n = __ldg(p.lb.perm + (g + 1) * 32 + lane);
      }
      using St = typename std::conditional<FAM == F_LEV && sizeof(W) == 4, LevSt32Fast,
                 typename std::conditional<FAM == F_LEV, LevSt<W>,
                 typename std::conditional<FAM == F_OSA, OsaSt<W>, LcsSt<W>>::type>::type>::type;
      St st;
      st.init(p.len1);
      const uint32_t nrows = (lmax + 7u) >> 3, nfull = lmin >> 3;
#define RF_ST(K, WORD) st.template step<K>(WORD, pm_lane, pm_lane_saddr, p.two);
#define RF_ROW8(V) RF_ST(0, (V).x) RF_ST(1, (V).x) RF_ST(2, (V).x) RF_ST(3, (V).x) RF_ST(0, (V).y) RF_ST(1, (V).y) RF_ST(2, (V).y) RF_ST(3, (V).y)
      uint32_t i = 0;
      for (; i < nfull; ++i) {  // rows every lane needs in full: no predicates
        const uint2 v = row(k + i);
        RF_ROW8(v)
      }
      if (lmin == lmax) {  // (nearly every group) all 32 candidates have the same length: warp-uniform remainder
        const uint32_t rem = lmax & 7u;
        if (rem) {
          const uint2 v = row(k + i);
          RF_ST(0, v.x)
          if (rem > 1) RF_ST(1, v.x)
          if (rem > 2) RF_ST(2, v.x)
          if (rem > 3) RF_ST(3, v.x)
          if (rem > 4) RF_ST(0, v.y)
          if (rem > 5) RF_ST(1, v.y)
          if (rem > 6) RF_ST(2, v.y)
        }
      } else {
        for (; i < nrows; ++i) {  // ragged group (a length boundary of the sorted block, or padding lanes)
          const uint2 v = row(k + i);
          const uint32_t j0 = i * 8u;
#define RF_TSTEP(T, WORD, K) if (j0 + (T) < len2) RF_ST(K, WORD)
          RF_TSTEP(0, v.x, 0) RF_TSTEP(1, v.x, 1) RF_TSTEP(2, v.x, 2) RF_TSTEP(3, v.x, 3)
          RF_TSTEP(4, v.y, 0) RF_TSTEP(5, v.y, 1) RF_TSTEP(6, v.y, 2) RF_TSTEP(7, v.y, 3)
#undef RF_TSTEP
        }
      }
#undef RF_ROW8
#undef RF_ST
      k += nrows;
      const uint32_t raw = (p.len1 == 0) ? ((FAM == F_LCS) ? 0u : len2) : st.raw(len2);
      if (idx != 0xFFFFFFFFu) {
        if constexpr (RAWDIST) {
          reinterpret_cast<uint32_t*>(p.out)[idx] = raw;
        } else {
          if (p.out_f64) reinterpret_cast<double*>(p.out)[idx] = finish_norm(p.epi, raw, p.len1, len2);
          else reinterpret_cast<uint32_t*>(p.out)[idx] = finish_int(p.epi, raw, p.len1, len2);
        }
      }
    }
    Q += nblk;
    chunk = __shfl_sync(0xffffffffu, next_chunk, 0);
  }
}

template <int FAM, class W, int NT, bool RAWDIST, int BLK>
static cudaError_t launch_lbr_cfg(const ScanLaunch& L, const void* tab) {
  auto kern = scan_lbr_kernel<FAM, W, NT, RAWDIST, BLK>;
  const size_t smem = sizeof(W) * 256 * 32 + (size_t)(NT / 32) * (2 * BLK * 256 + 16);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  int ctas_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, NT, smem);
  if (e != cudaSuccess) return e;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  LbParams p{};
  p.lb = L.lb;
  p.tab = tab;
  p.len1 = L.query.len1;
  p.out = L.out;
  p.out_f64 = L.out_is_f64;
  p.two = 2;
  p.chunk = 16;
  p.counter = L.lb_counter;
  p.epi = L.epi;
  e = cudaMemsetAsync(p.counter, 0, sizeof(unsigned long long), L.stream);
  if (e != cudaSuccess) return e;
  uint64_t grid = (uint64_t)L.sm_count * ctas_per_sm;
  const uint64_t nchunks = (L.lb.ngroups + p.chunk - 1) / p.chunk;
  const uint64_t need = (nchunks + NT / 32 - 1) / (NT / 32);
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  kern<<<(uint32_t)grid, NT, smem, L.stream>>>(p);
  g_launches.fetch_add(1);
  return cudaGetLastError();
}

template <int FAM, class W, int NT, bool RAWDIST = false>
static cudaError_t launch_lbr_inst(const ScanLaunch& L, const void* tab) {
  static const int blk = getenv("RF_RING_BLK") ? atoi(getenv("RF_RING_BLK")) : 4;  // tuning knob: rows per TMA block
  if (blk == 8) return launch_lbr_cfg<FAM, W, NT, RAWDIST, 8>(L, tab);
  return launch_lbr_cfg<FAM, W, NT, RAWDIST, 4>(L, tab);
}

cudaError_t launch_scan_lbr(const ScanLaunch& L) {
  const Family fam = family_of(L.epi.metric, L.epi.wclass);
  const bool w32 = L.query.len1 <= 32;
  const bool rawdist = (fam == F_LEV || fam == F_OSA) && L.epi.unit32 && L.epi.kind == K_DISTANCE && !L.epi.has_cutoff &&
                       !L.out_is_f64 && L.query.len1 >= 1;
  switch (fam) {
    case F_LEV:
      if (rawdist)
        return w32 ? launch_lbr_inst<F_LEV, uint32_t, 512, true>(L, L.query.tab32_top)
                   : launch_lbr_inst<F_LEV, uint64_t, 512, true>(L, L.query.tab64_top);
      return w32 ? launch_lbr_inst<F_LEV, uint32_t, 512>(L, L.query.tab32_top)
                 : launch_lbr_inst<F_LEV, uint64_t, 512>(L, L.query.tab64_top);
    case F_OSA:
      if (rawdist)
        return w32 ? launch_lbr_inst<F_OSA, uint32_t, 512, true>(L, L.query.tab32_top)
                   : launch_lbr_inst<F_OSA, uint64_t, 512, true>(L, L.query.tab64_top);
      return w32 ? launch_lbr_inst<F_OSA, uint32_t, 512>(L, L.query.tab32_top)
                 : launch_lbr_inst<F_OSA, uint64_t, 512>(L, L.query.tab64_top);
    case F_LCS:
      return w32 ? launch_lbr_inst<F_LCS, uint32_t, 512>(L, L.query.tab32_bot)
                 : launch_lbr_inst<F_LCS, uint64_t, 512>(L, L.query.tab64_bot);
    default:
      return launch_scan_lb(L);  // Jaro: random access to the candidate
  }
}

// ------------------------------------------------------------------------------------------------ cdist
// many-vs-many Levenshtein top-k.  Work unit = (corpus slice, query), handed to the persistent CTAs by a global
// counter in slice-major order: at any time all CTAs scan the same (L2-resident) slice for different queries.  A
// slice is the WHOLE shard whenever it fits in L2 and there are enough queries to fill the grid (config 5: one
// slice, 10^4 units), so a CTA sees ~10^6 candidates per query: the k-best lists warm up once per unit, the
// per-unit costs (1 KB match table rebuilt in shared memory, merge of the 16 warp lists) vanish, and the k-th
// bound becomes tight early enough to skip whole groups by length (|len2 - len1| > bound  =>  d > bound).
// Slices of different units of one query are merged by a second kernel (not launched for a single slice).
constexpr int CD_NT = 512;  // 2 CTAs / SM (cdist_grid) x 16 warps: 8 warps per scheduler
constexpr int CD_KMAX = 128;
constexpr int CD_SMAX = 256;  // slices
constexpr unsigned long long CD_NOKEY = 0xFFFFFFFFFFFFFFFFull;

struct CdistParams {
  LbView lb;
  uint64_t total_rows;
  const void* tabs;        // [nq][256] top-aligned tables of W
  const uint32_t* q_len;   // [nq]
  uint32_t nq, k;
  int has_cutoff;
  uint32_t cutoff;
  uint32_t nslices;
  unsigned long long* counter;  // next unit, zeroed before the launch
  unsigned long long* scratch;  // [nq][nslices][k]   (nslices > 1)
  uint32_t* out_idx;            // [nq][k]            (nslices == 1: written directly)
  uint32_t* out_dist;
  uint32_t two;
  int skip;                     // group skipping by length on (run-time switch for measurements)
  int metric;                   // M_LEVENSHTEIN / M_OSA / M_INDEL / M_LCS_SEQ
};

// CTA-wide extraction of the k smallest keys of keys[0..m) (destroys them), ascending, into best[0..k)
__device__ __forceinline__ void cd_extract(unsigned long long* keys, uint32_t m, unsigned long long* best, uint32_t k,
                                           unsigned long long* wmin) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  unsigned long long last = CD_NOKEY;
  for (uint32_t r = 0; r < k; ++r) {
    unsigned long long mn = CD_NOKEY;
    for (uint32_t i = tid; i < m; i += CD_NT) {
      unsigned long long v = keys[i];
      if (v == last && last != CD_NOKEY) { v = CD_NOKEY; keys[i] = v; }  // retire the previous winner
      mn = v < mn ? v : mn;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, mn, d);
      mn = o < mn ? o : mn;
    }
    if (lane == 0) wmin[warp] = mn;
    __syncthreads();
    unsigned long long g = wmin[0];
#pragma unroll
    for (int w = 1; w < CD_NT / 32; ++w) g = wmin[w] < g ? wmin[w] : g;
    if (tid == 0) best[r] = g;
    last = g;
    __syncthreads();
    if (g == CD_NOKEY) {  // nothing left: pad the remainder
      for (uint32_t j = r + 1 + tid; j < k; j += CD_NT) best[j] = CD_NOKEY;
      break;
    }
  }
  __syncthreads();
}

// A warp's running k-best list, ascending, held in registers: slot i = 32*r + lane lives in v[r] of that lane
// (KR = ceil(k/32) registers of 64 bits per lane).  Inserting a key is warp-cooperative: count the slots below it
// with ballots, shift the rest up by one slot with shuffles.  No shared memory, no CTA barrier while a CTA scans
// its slice; with random data a warp inserts O(k log(n/k)) of the n keys it sees.
template <int KR>
struct WarpTopK {
  unsigned long long v[KR];
  unsigned long long kth;  // value of slot k-1 (warp-uniform): only keys below it can enter
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int r = 0; r < KR; ++r) v[r] = CD_NOKEY;
    kth = CD_NOKEY;
  }
  __device__ __forceinline__ void insert(unsigned long long key, uint32_t k, uint32_t lane) {  // key < kth, keys are unique
    uint32_t cnt = 0;
#pragma unroll
    for (int r = 0; r < KR; ++r) cnt += __popc(__ballot_sync(0xffffffffu, v[r] < key));
#pragma unroll
    for (int r = KR - 1; r >= 0; --r) {
      const uint32_t i = (uint32_t)r * 32u + lane;
      unsigned long long below = __shfl_up_sync(0xffffffffu, v[r], 1);
      if (r > 0) {
        const unsigned long long prev_top = __shfl_sync(0xffffffffu, v[r - 1], 31);
        if (lane == 0) below = prev_top;
      }
      unsigned long long nv = (i < cnt) ? v[r] : (i == cnt ? key : below);
      if (i >= k) nv = CD_NOKEY;
      v[r] = nv;
    }
    unsigned long long t = v[0];
#pragma unroll
    for (int r = 1; r < KR; ++r) t = ((k - 1) / 32 == (uint32_t)r) ? v[r] : t;  // no dynamic register indexing
    kth = __shfl_sync(0xffffffffu, t, (k - 1) & 31);
  }
  // the 32 keys of one scored group (one per lane).  `bound` = min(kth, CTA-wide bound): the smallest k-th value any
  // warp of the CTA has reached; a key at or above it is beaten by k keys of that warp, so it cannot be in the
  // CTA's k best either.  Stale bounds are safe (the bound only ever tightens).
  __device__ __forceinline__ void offer(unsigned long long key, unsigned long long bound, uint32_t k, uint32_t lane,
                                        unsigned long long* shared_kth) {
    uint32_t m = __ballot_sync(0xffffffffu, key < bound);
    if (m == 0) return;
    const unsigned long long before = kth;
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const unsigned long long kk = __shfl_sync(0xffffffffu, key, src);
      if (kk < kth) insert(kk, k, lane);  // kth may have dropped since the ballot
    }
    if (kth < before && lane == 0) atomicMin(shared_kth, kth);
  }
};

// FAM: F_LEV (Levenshtein), F_OSA, F_LCS (Indel: len1 + len2 - 2 lcs; LCSseq distance: max(len1, len2) - lcs; p.metric tells
// which).  All four are bounded below by |len1 - len2|, so the length skip stays exact.
template <class W, int KR, int FAM = F_LEV>
__global__ void __launch_bounds__(CD_NT) cdist_scan_kernel(const __grid_constant__ CdistParams p) {
  constexpr int NW = CD_NT / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  W* pm = reinterpret_cast<W*>(smem_raw);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw + sizeof(W) * 8192);  // NW * k
  unsigned long long* best = keys + NW * CD_KMAX;                                                  // CD_KMAX
  unsigned long long* wmin = best + CD_KMAX;                                                       // NW
  __shared__ uint64_t s_bounds[CD_SMAX + 1];
  __shared__ unsigned long long s_unit;
  __shared__ unsigned long long s_kth;  // CTA-wide bound on the k-th best key of the current unit
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint64_t ngroups = p.lb.ngroups;
  const uint32_t nslices = p.nslices;
  for (uint32_t b = tid; b <= nslices; b += CD_NT) {  // slice b = groups [bounds[b], bounds[b+1]): equal shares of rows
    uint64_t g = ngroups;
    if (b == 0) g = 0;
    else if (b < nslices) {
      const uint64_t target = (uint64_t)((unsigned __int128)p.total_rows * b / nslices);
      uint64_t lo = 0, hi = ngroups;
      while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (p.lb.goff[mid] < target) lo = mid + 1; else hi = mid;
      }
      g = lo;
    }
    s_bounds[b] = g;
  }
  const W* __restrict__ pm_lane = pm + lane;
  const uint2* __restrict__ gdata = reinterpret_cast<const uint2*>(p.lb.gdata);
  const uint32_t k = p.k;
  const unsigned long long units = (unsigned long long)nslices * p.nq;
  for (;;) {
    __syncthreads();  // the previous unit's shared-memory traffic is complete
    if (tid == 0) {
      s_unit = atomicAdd(p.counter, 1ull);
      s_kth = CD_NOKEY;
    }
    __syncthreads();
    const unsigned long long unit = s_unit;
    if (unit >= units) break;
    const uint32_t sl = (uint32_t)(unit / p.nq), q = (uint32_t)(unit - (unsigned long long)sl * p.nq);
    const uint64_t g_lo = s_bounds[sl], g_hi = s_bounds[sl + 1];
    {
      const W* __restrict__ t = reinterpret_cast<const W*>(p.tabs) + (size_t)q * 256;
      if constexpr (sizeof(W) == 8) {  // split low / high word tables (lev_w1_u64_fast)
        uint32_t* pm32 = reinterpret_cast<uint32_t*>(smem_raw);
        for (uint32_t i = tid; i < 8192u; i += CD_NT) {
          const uint64_t v = t[i >> 5];
          pm32[i] = (uint32_t)v;
          pm32[i + 8192u] = (uint32_t)(v >> 32);
        }
      } else {
        for (uint32_t i = tid; i < 8192u; i += CD_NT) pm[i] = t[i >> 5];
      }
    }
    __syncthreads();
    const uint32_t len1 = p.q_len[q];
    const uint32_t static_bound = p.has_cutoff ? p.cutoff : 0xFFFFFFFFu;
    WarpTopK<KR> top;
    top.reset();
    // the warps take the slice's groups round-robin (neighbouring groups have similar lengths); software pipeline:
    // the next group's length / index / first two rows are requested (L2 hits: the slice is resident) before
    // the current group is scored
    uint64_t g = g_lo + warp;
    uint32_t len_n = 0, idx_n = 0;
    const uint2* col_n = gdata;
    uint2 first_n = make_uint2(0u, 0u), second_n = make_uint2(0u, 0u);
    if (g < g_hi) {
      len_n = __ldg(p.lb.lens + g * 32 + lane);
      idx_n = __ldg(p.lb.perm + g * 32 + lane);
      col_n = gdata + __ldg(p.lb.goff + g) * 32 + lane;
      first_n = __ldg(col_n);
      second_n = __ldg(col_n + 32);
    }
    for (; g < g_hi; g += NW) {
      const uint32_t len2 = len_n, idx = idx_n;
      const LaneSrcT<false> src{col_n, first_n, second_n};
      const uint64_t gn = g + NW;
      if (gn < g_hi) {
        len_n = __ldg(p.lb.lens + gn * 32 + lane);
        idx_n = __ldg(p.lb.perm + gn * 32 + lane);
        col_n = gdata + __ldg(p.lb.goff + gn) * 32 + lane;
        first_n = __ldg(col_n);
        second_n = __ldg(col_n + 32);
      }
      unsigned long long bound = top.kth;
      {
        const unsigned long long s = *reinterpret_cast<volatile unsigned long long*>(&s_kth);
        bound = s < bound ? s : bound;
      }
      // d >= |len2 - len1|: a group whose every candidate is further than the current k-th distance (or the
      // cutoff) by length alone cannot contribute.  Equal distance can still win on the index, hence '>'.
      uint32_t bd = (uint32_t)(bound >> 32);
      bd = bd < static_bound ? bd : static_bound;
      const uint32_t ldiff = len2 > len1 ? len2 - len1 : len1 - len2;
      const bool live = idx != 0xFFFFFFFFu && ldiff <= bd;
      if (p.skip && !__any_sync(0xffffffffu, live)) continue;
      uint32_t d;
      const uint32_t saddr = sizeof(W) == 4 ? smem_u32(pm_lane) : smem_u32(reinterpret_cast<const uint32_t*>(smem_raw) + lane);
      if constexpr (FAM == F_LCS) {
        uint32_t lcs = 0;
        if (len1 != 0) {
          if constexpr (sizeof(W) == 4) lcs = lcs_w1_u32_fast(saddr, src.reader(), len2, p.two);
          else lcs = lcs_w1_u64_fast(saddr, src.reader(), len2, p.two);
        }
        d = p.metric == M_INDEL ? len1 + len2 - 2u * lcs : (len1 > len2 ? len1 : len2) - lcs;
      } else if (len1 == 0) {
        d = len2;
      } else if constexpr (sizeof(W) == 4) {
        d = myers_w1_u32_fast<FAM == F_OSA>(saddr, src.reader(), len2, len1, p.two);
      } else {
        d = myers_w1_u64_fast<FAM == F_OSA>(saddr, src.reader(), len2, len1, p.two);
      }
      const bool ok = idx != 0xFFFFFFFFu && d <= static_bound;
      top.offer(ok ? (((unsigned long long)d << 32) | idx) : CD_NOKEY, bound, k, lane, &s_kth);
    }
    // the NW warp lists -> the CTA's k best of this unit
#pragma unroll
    for (int r = 0; r < KR; ++r) {
      const uint32_t i = (uint32_t)r * 32u + lane;
      if (i < k) keys[warp * k + i] = top.v[r];
    }
    __syncthreads();
    cd_extract(keys, NW * k, best, k, wmin);
    if (nslices == 1) {
      for (uint32_t i = tid; i < k; i += CD_NT) {
        const unsigned long long v = best[i];
        p.out_idx[(size_t)q * k + i] = (v == CD_NOKEY) ? 0xFFFFFFFFu : (uint32_t)v;
        p.out_dist[(size_t)q * k + i] = (v == CD_NOKEY) ? 0xFFFFFFFFu : (uint32_t)(v >> 32);
      }
    } else {
      unsigned long long* out = p.scratch + ((size_t)q * nslices + sl) * k;
      for (uint32_t i = tid; i < k; i += CD_NT) out[i] = best[i];
    }
  }
}

// one CTA per query: k smallest of the parts*k per-slice keys -> (idx, dist) rows
__global__ void __launch_bounds__(CD_NT) cdist_merge_kernel(const unsigned long long* __restrict__ scratch, uint32_t parts,
                                                            uint32_t k, uint32_t* __restrict__ out_idx,
                                                            uint32_t* __restrict__ out_dist) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);  // parts*k
  unsigned long long* best = keys + (size_t)parts * k;
  unsigned long long* wmin = best + CD_KMAX;
  const uint32_t q = blockIdx.x, m = parts * k;
  const unsigned long long* src = scratch + (size_t)q * m;
  for (uint32_t i = threadIdx.x; i < m; i += CD_NT) keys[i] = src[i];
  __syncthreads();
  cd_extract(keys, m, best, k, wmin);
  for (uint32_t i = threadIdx.x; i < k; i += CD_NT) {
    const unsigned long long v = best[i];
    out_idx[(size_t)q * k + i] = (v == CD_NOKEY) ? 0xFFFFFFFFu : (uint32_t)v;
    out_dist[(size_t)q * k + i] = (v == CD_NOKEY) ? 0xFFFFFFFFu : (uint32_t)(v >> 32);
  }
}

uint32_t cdist_grid(int sm_count) { return (uint32_t)sm_count * 2u; }

// slices: enough units to keep the grid busy when there are few queries, and a slice small enough to stay in L2
// (the 126 MB L2 holds ~48 MB comfortably next to the tables and the other die's copy) while every query visits it
uint32_t cdist_slices(int sm_count, uint32_t nq, uint64_t layout_bytes, uint64_t ngroups) {
  const uint64_t grid = cdist_grid(sm_count);
  uint64_t s_par = (3 * grid + nq - 1) / nq;
  uint64_t s_l2 = (layout_bytes + (48ull << 20) - 1) / (48ull << 20);
  uint64_t s = s_par > s_l2 ? s_par : s_l2;
  const uint64_t cap = ngroups / 64 + 1;  // at least 64 groups (4 per warp) in a slice
  if (s > cap) s = cap;
  if (s > (uint64_t)CD_SMAX) s = CD_SMAX;
  if (s < 1) s = 1;
  return (uint32_t)s;
}

cudaError_t launch_cdist_topk(const CdistLaunch& L) {
  if (L.k == 0 || L.k > CD_KMAX || L.nslices == 0 || L.nslices > (uint32_t)CD_SMAX) return cudaErrorInvalidValue;
  CdistParams p{};
  p.lb = L.lb;
  p.total_rows = L.total_rows;
  p.tabs = L.q_tabs;
  p.q_len = L.q_len;
  p.nq = L.nq;
  p.k = L.k;
  p.has_cutoff = L.has_cutoff;
  p.cutoff = L.cutoff;
  p.nslices = L.nslices;
  p.counter = L.counter;
  p.scratch = L.scratch;
  p.out_idx = L.out_idx;
  p.out_dist = L.out_dist;
  p.two = 2;
  p.skip = L.skip;
  p.metric = L.metric;
  const size_t wsz = L.wide ? 8 : 4;
  const size_t smem = wsz * 8192 + sizeof(unsigned long long) * ((CD_NT / 32) * CD_KMAX + CD_KMAX + CD_NT / 32);
  cudaError_t e = cudaMemsetAsync(L.counter, 0, sizeof(unsigned long long), L.stream);
  if (e != cudaSuccess) return e;
  const unsigned long long units = (unsigned long long)L.nslices * L.nq;
  const uint32_t grid = (uint32_t)(units < L.grid ? units : L.grid);
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kern<<<grid, CD_NT, smem, L.stream>>>(p);
    return cudaGetLastError();
  };
  const int kr = L.k <= 32 ? 1 : (L.k <= 64 ? 2 : 4);
  auto pick = [&](auto fam_tag) -> cudaError_t {
    constexpr int FAM = decltype(fam_tag)::value;
    if (L.wide) return kr == 1 ? launch(cdist_scan_kernel<uint64_t, 1, FAM>) : kr == 2 ? launch(cdist_scan_kernel<uint64_t, 2, FAM>) : launch(cdist_scan_kernel<uint64_t, 4, FAM>);
    return kr == 1 ? launch(cdist_scan_kernel<uint32_t, 1, FAM>) : kr == 2 ? launch(cdist_scan_kernel<uint32_t, 2, FAM>) : launch(cdist_scan_kernel<uint32_t, 4, FAM>);
  };
  switch (L.metric) {
    case M_LEVENSHTEIN: e = pick(std::integral_constant<int, F_LEV>{}); break;
    case M_OSA: e = pick(std::integral_constant<int, F_OSA>{}); break;
    case M_INDEL: case M_LCS_SEQ: e = pick(std::integral_constant<int, F_LCS>{}); break;
    default: return cudaErrorInvalidValue;
  }
  if (e != cudaSuccess) return e;
  g_launches.fetch_add(1);
  if (L.nslices == 1) return cudaSuccess;
  const size_t msmem = sizeof(unsigned long long) * ((size_t)L.nslices * L.k + CD_KMAX + CD_NT / 32);
  e = cudaFuncSetAttribute(cdist_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem);
  if (e != cudaSuccess) return e;
  cdist_merge_kernel<<<L.nq, CD_NT, msmem, L.stream>>>(L.scratch, L.nslices, L.k, L.out_idx, L.out_dist);
  g_launches.fetch_add(1);
  return cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------ mw
struct MwParams {
  const uint8_t* chars;
  const uint32_t* off32;
  const uint64_t* off64;
  uint64_t n;
  const uint64_t* pm;  // [256][words]
  uint32_t len1;
  uint32_t words;
  uint32_t G;  // lanes per candidate (power of two, <= 32)
  void* out;
  int out_f64;
  Epi epi;
};

// T = uint16_t: candidates renamed to 16-bit codes (u32 queries with more than 255 distinct symbols against u32 corpora with
// more than 255 distinct symbols: p.pm then has one row per code, p.chars is an array of T)
template <int FAM, int WPL, class T = uint8_t>
__global__ void __launch_bounds__(256) scan_mw_kernel(const __grid_constant__ MwParams p) {
  constexpr uint32_t CH_BITS = sizeof(T) * 8, CH_MASK = (1u << CH_BITS) - 1u;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t G = p.G;
  const uint32_t gl = lane & (G - 1u);       // lane within the candidate's group
  const uint32_t gpw = 32u / G;              // groups (candidates) per warp pass
  const uint32_t sub = lane / G;
  const uint64_t warp_global = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint64_t total_warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  const uint32_t words = p.words;
  const uint32_t w0 = gl * WPL;                                // first 64-bit block owned by this lane
  const uint32_t act = (words + WPL - 1) / WPL;                // lanes that own at least one block
  const uint32_t last_owner = (words - 1) / WPL, last_k = (words - 1) % WPL;
  const uint32_t last_bit = (p.len1 - 1u) & 63u;
  const bool off64 = p.off64 != nullptr;
  const uint64_t* __restrict__ pm = p.pm;
  // Levenshtein distance with a cutoff: candidates are dropped on |len1-len2| (levenshtein.rs:1045-1047) and
  // abandoned as soon as the bottom-row score can no longer come back under the cutoff (the reference's
  // Ukkonen band does the same job, levenshtein.rs:897-985); both only ever turn a result into None.
  const bool lev_cut = (FAM == F_LEV) && p.epi.metric == M_LEVENSHTEIN && p.epi.kind == K_DISTANCE && p.epi.has_cutoff &&
                       p.epi.wclass == WC_UNIFORM;
  const uint64_t cut64 = lev_cut ? p.epi.cutoff_u / p.epi.w_ins : 0;
  const uint32_t cut = cut64 > 0x7FFFFFFFull ? 0x7FFFFFFFu : (uint32_t)cut64;

  // A warp takes 32 consecutive candidates at a time: every lane classifies one of them, trivial ones are
  // finished on the spot, the survivors are compacted with a ballot and scored gpw at a time.
  for (uint64_t base = warp_global * 32; base < p.n; base += total_warps * 32) {
    const uint64_t c_mine = base + lane;
    uint64_t o0_mine = 0;
    uint32_t len2_mine = 0;
    bool need = false;
    if (c_mine < p.n) {
      o0_mine = off64 ? p.off64[c_mine] : (uint64_t)p.off32[c_mine];
      const uint64_t o1 = off64 ? p.off64[c_mine + 1] : (uint64_t)p.off32[c_mine + 1];
      len2_mine = (uint32_t)(o1 - o0_mine);
      const uint32_t diff = p.len1 > len2_mine ? p.len1 - len2_mine : len2_mine - p.len1;
      if (lev_cut && diff > cut) {
        if (p.out_f64) reinterpret_cast<double*>(p.out)[c_mine] = qnan();
        else reinterpret_cast<uint32_t*>(p.out)[c_mine] = NONE_U32;
      } else if (len2_mine == 0) {
        const uint32_t raw = (FAM == F_LCS) ? 0u : p.len1;
        if (p.out_f64) reinterpret_cast<double*>(p.out)[c_mine] = finish_norm(p.epi, raw, p.len1, 0);
        else reinterpret_cast<uint32_t*>(p.out)[c_mine] = finish_int(p.epi, raw, p.len1, 0);
      } else {
        need = true;
      }
    }
    uint32_t todo = __ballot_sync(0xffffffffu, need);
#ifdef RF_DEBUG_MW
    if (lane == 0) printf("warp %llu base %llu todo %08x\n", (unsigned long long)warp_global, (unsigned long long)base, todo);
#endif
    while (todo) {
      const uint32_t src = __fns(todo, 0, sub + 1);  // this sub-group's candidate = (sub+1)-th surviving lane
      const bool have = src != 0xFFFFFFFFu;
      const uint32_t srcl = have ? src : 0u;
      const uint64_t c = __shfl_sync(0xffffffffu, c_mine, srcl);
      const uint64_t o0 = __shfl_sync(0xffffffffu, o0_mine, srcl);
      const uint32_t len2_src = __shfl_sync(0xffffffffu, len2_mine, srcl);
      const uint32_t len2 = have ? len2_src : 0u;
      for (uint32_t i = 0; i < gpw; ++i) todo &= todo - 1;  // retire the gpw candidates of this pass

      const T* __restrict__ txt = reinterpret_cast<const T*>(p.chars) + o0;
      const uint32_t my_steps = have ? len2 + act - 1u : 0u;
      const uint32_t steps = __reduce_max_sync(0xffffffffu, my_steps);

      uint64_t VP[WPL], VN[WPL], D0[WPL], PMo[WPL], Xc[WPL], Xn[WPL];
#pragma unroll
      for (int k = 0; k < WPL; ++k) {
        VP[k] = ~0ull;  // LCS: VP plays the role of S
        VN[k] = 0; D0[k] = 0; PMo[k] = 0; Xc[k] = 0; Xn[k] = 0;
      }
      int32_t score = 0;
      bool dead = false;  // abandoned: result is None
      // lane 0 of a group feeds the text: ch_cur = text[t], two bytes prefetched ahead
      uint32_t ch_cur = 0, c1 = 0, c2 = 0;
      if (gl == 0 && my_steps) {
        ch_cur = txt[0];
        c1 = len2 > 1 ? txt[1] : 0;
        c2 = len2 > 2 ? txt[2] : 0;
#pragma unroll
        for (int k = 0; k < WPL; ++k) Xc[k] = (w0 + k < words) ? __ldg(pm + (uint64_t)ch_cur * words + w0 + k) : 0ull;
      }
      uint32_t cout = 0;
      for (uint32_t t = 0; t < steps; ++t) {
        const uint32_t pk_in = __shfl_up_sync(0xffffffffu, ch_cur | (cout << CH_BITS), 1, G);
        uint32_t ch_nxt, cin;
        if (gl == 0) {
          ch_nxt = c1;
          c1 = c2;
          c2 = (t + 3 < len2) ? (uint32_t)txt[t + 3] : 0u;
          cin = (FAM == F_LCS) ? 0u : 1u;  // Levenshtein/OSA: +1 horizontal delta enters row 0; LCS: no carry
        } else {
          ch_nxt = pk_in & CH_MASK;
          cin = pk_in >> CH_BITS;
        }
        // this lane's match words for the NEXT step (text char flows one step ahead of the carries)
#pragma unroll
        for (int k = 0; k < WPL; ++k) Xn[k] = (w0 + k < words) ? __ldg(pm + (uint64_t)ch_nxt * words + w0 + k) : 0ull;

        const int32_t j = (int32_t)t - (int32_t)gl;  // text column handled by this lane in this step
        const bool active = (gl < act) && j >= 0 && j < (int32_t)len2 && my_steps && !dead;
        if (active) {
          if constexpr (FAM == F_LCS) {
            uint64_t carry = cin & 1u;
#pragma unroll
            for (int k = 0; k < WPL; ++k) {
              const uint64_t S = VP[k];
              const uint64_t u = S & Xc[k];
              const uint64_t x1 = S + u;
              const uint64_t x2 = x1 + carry;
              carry = (uint64_t)(x1 < S) | (uint64_t)(x2 < x1);
              VP[k] = x2 | (S - u);
            }
            cout = (uint32_t)carry;
          } else {
            uint64_t hp_c = cin & 1u, hn_c = (cin >> 1) & 1u, tr_c = (cin >> 2) & 1u;
#pragma unroll
            for (int k = 0; k < WPL; ++k) {
              const uint64_t X0 = Xc[k];
              const uint64_t X = X0 | hn_c;
              uint64_t d0 = ((((X & VP[k]) + VP[k]) ^ VP[k]) | X) | VN[k];
              if constexpr (FAM == F_OSA) {
                const uint64_t nd = (~D0[k]) & X0;
                d0 |= ((nd << 1) | tr_c) & PMo[k];
                tr_c = nd >> 63;
                D0[k] = d0;
                PMo[k] = X0;
              }
              uint64_t HP = VN[k] | ~(d0 | VP[k]);
              uint64_t HN = d0 & VP[k];
              if (gl == last_owner && k == (int)last_k)
                score += (int32_t)((HP >> last_bit) & 1u) - (int32_t)((HN >> last_bit) & 1u);
              const uint64_t hp_o = HP >> 63, hn_o = HN >> 63;
              HP = (HP << 1) | hp_c;
              HN = (HN << 1) | hn_c;
              VP[k] = HN | ~(d0 | HP);
              VN[k] = HP & d0;
              hp_c = hp_o;
              hn_c = hn_o;
            }
            cout = (uint32_t)hp_c | ((uint32_t)hn_c << 1) | ((uint32_t)tr_c << 2);
          }
        }
        ch_cur = ch_nxt;
#pragma unroll
        for (int k = 0; k < WPL; ++k) Xc[k] = Xn[k];

        if (lev_cut && (t & 7u) == 7u) {
          // bottom-row score after column jl: D[m][jl] = len1 + score; each remaining column lowers it by <= 1
          const int32_t jl = (int32_t)t - (int32_t)last_owner;
          bool hopeless = false;
          if (gl == last_owner && jl >= 0 && jl < (int32_t)len2 && !dead && my_steps) {
            const int64_t cur = (int64_t)p.len1 + score;
            hopeless = cur > (int64_t)cut + (int64_t)(len2 - 1 - (uint32_t)jl);
          }
          const int hopeless_g = __shfl_sync(0xffffffffu, (int)hopeless, last_owner, G);  // every lane, no short-circuit
          dead = dead || (hopeless_g != 0);
#ifdef RF_DEBUG_MW
          if (lane == 0) printf("  t %u steps %u my_steps %u dead %d score %d\n", t, steps, my_steps, (int)dead, score);
#endif
          if (__ballot_sync(0xffffffffu, my_steps && !dead && t + 1 < my_steps) == 0) break;  // whole warp done
        }
      }

      // ---- result
      uint32_t raw;
      if constexpr (FAM == F_LCS) {
        uint32_t cnt = 0;
#pragma unroll
        for (int k = 0; k < WPL; ++k)
          if (w0 + k < words) cnt += (uint32_t)__popcll(~VP[k]);
        for (uint32_t d = G >> 1; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d, G);
        raw = cnt;
      } else {
        const int32_t sc = __shfl_sync(0xffffffffu, score, last_owner, G);
        raw = (uint32_t)((int32_t)p.len1 + sc);
      }
      if (have && gl == 0) {
        if (dead) {
          if (p.out_f64) reinterpret_cast<double*>(p.out)[c] = qnan();
          else reinterpret_cast<uint32_t*>(p.out)[c] = NONE_U32;
        } else if (p.out_f64) {
          reinterpret_cast<double*>(p.out)[c] = finish_norm(p.epi, raw, p.len1, len2);
        } else {
          reinterpret_cast<uint32_t*>(p.out)[c] = finish_int(p.epi, raw, p.len1, len2);
        }
      }
    }
  }
}

template <int FAM>
static cudaError_t launch_mw_fam(const ScanLaunch& L) {
  MwParams p{};
  p.chars = L.corpus.chars;
  p.off32 = L.corpus.off32;
  p.off64 = L.corpus.off64;
  p.n = L.corpus.n;
  p.pm = L.query.pm_words;
  p.len1 = L.query.len1;
  p.words = L.query.words;
  p.out = L.out;
  p.out_f64 = L.out_is_f64;
  p.epi = L.epi;
  // blocks per lane: 1 up to 32 blocks (2048 elements), then 2/4/8 (RF_MAX_QUERY_LEN = 16384)
  int wpl = 1;
  while ((uint32_t)wpl * 32u < p.words) wpl <<= 1;
  if (wpl > 8) return cudaErrorInvalidValue;
  const uint32_t act = (p.words + wpl - 1) / wpl;
  uint32_t G = 1;
  while (G < act) G <<= 1;
  p.G = G;
  const uint32_t warps_per_block = 8;
  const uint64_t warps_needed = (p.n + 31) / 32;
  uint64_t blocks = (warps_needed + warps_per_block - 1) / warps_per_block;
  void (*kern)(MwParams);
  if (L.elem16) {
    switch (wpl) {
      case 1: kern = scan_mw_kernel<FAM, 1, uint16_t>; break;
      case 2: kern = scan_mw_kernel<FAM, 2, uint16_t>; break;
      case 4: kern = scan_mw_kernel<FAM, 4, uint16_t>; break;
      default: kern = scan_mw_kernel<FAM, 8, uint16_t>; break;
    }
  } else {
    switch (wpl) {
      case 1: kern = scan_mw_kernel<FAM, 1>; break;
      case 2: kern = scan_mw_kernel<FAM, 2>; break;
      case 4: kern = scan_mw_kernel<FAM, 4>; break;
      default: kern = scan_mw_kernel<FAM, 8>; break;
    }
  }
  const uint64_t max_blocks = resident_ctas(kern, 256, 0, L.sm_count, 2);  // one wave (48 ... 128 registers per thread)
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  kern<<<(uint32_t)blocks, 256, 0, L.stream>>>(p);
  g_launches.fetch_add(1);
  return cudaGetLastError();
}

cudaError_t launch_scan_mw(const ScanLaunch& L) {
  switch (family_of(L.epi.metric, L.epi.wclass)) {
    case F_LEV: return launch_mw_fam<F_LEV>(L);
    case F_OSA: return launch_mw_fam<F_OSA>(L);
    case F_LCS: return launch_mw_fam<F_LCS>(L);
    default: return cudaErrorInvalidValue;
  }
}

// Row walker with a SMALL loop body for steps of 100+ instructions (the multi-limb kernels): walk_rows8's four-row
// unrolling puts 32 steps = 54 KB of code into the loop, which thrashes the 32 KB L1.5 instruction cache (ncu:
// "no_instruction" was the top stall, issue 35 %).  Here one row (or one 4-character word, WORD_LOOP) per iteration,
// still two rows in flight + the L2 prefetch; the ragged tail is a rolled loop with a run-time byte selector.
// CUT: after every row the lane tests whether its bottom-row value can still come back under the cutoff (each remaining
// column lowers it by at most one) and leaves the loop for good otherwise; returns false for such a lane.
template <bool WORD_LOOP, bool CUT = false, class Rd, class St>
__device__ __forceinline__ bool walk_rows8_compact(Rd rd, uint32_t len2, St& st, uint32_t cut = 0) {
  static_assert(Rd::kRow8, "interleaved rows only");
  uint2 A = rd.q0, B = rd.q1;
  const uint2* p = rd.p;
  const uint32_t nfull = len2 >> 3;
#pragma unroll 1
  for (uint32_t i = 0; i < nfull; ++i) {
    if constexpr (CUT) {
      const uint32_t j = i * 8u;
      if (st.bottom(j) > cut + (len2 - j)) return false;
    }
    if constexpr (Rd::kStream) prefetch_l2(p + 32 * kPfDist);
    const uint2 C = ld_row8<Rd::kStream>(p);
    p += 32;
    if constexpr (WORD_LOOP) {
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const uint32_t w = h ? A.y : A.x;
        st.template step<0>(w); st.template step<1>(w); st.template step<2>(w); st.template step<3>(w);
      }
    } else {
      { const uint32_t w = A.x; st.template step<0>(w); st.template step<1>(w); st.template step<2>(w); st.template step<3>(w); }
      { const uint32_t w = A.y; st.template step<0>(w); st.template step<1>(w); st.template step<2>(w); st.template step<3>(w); }
    }
    A = B;
    B = C;
  }
  const uint32_t rem = len2 & 7u;
  uint32_t w = A.x;
#pragma unroll 1
  for (uint32_t k = 0; k < rem; ++k) {
    if (k == 4) w = A.y;
    st.step_sel(w, 0x80u << (8u * (k & 3u)));
  }
  return true;
}

// ------------------------------------------------------------------------------------------------ lbn
// Multi-word queries of 65..512 elements over the interleaved layout: ONE THREAD per candidate, warp per group of 32
// equal-length candidates like scan_lb_kernel, the whole column of the bit-parallel recurrence in registers.
// (The reference walks the 64-bit blocks of the column one by one with the carries in locals, hyrroe2003_block,
// levenshtein.rs:838-875 / lcs_blockwise, lcs_seq.rs:267-341; scan_mw_kernel spreads the blocks over the lanes of a
// sub-warp and pays a shuffle per step.)  Formulation: the column is ONE integer of L = 4Q 32-bit limbs (Q = 1..4),
// pattern TOP-aligned in it, so the single-word recurrence of lev_w1 applies unchanged and the distance is
// n + popc(VP) - popc(VN) at the end -- no per-block score bookkeeping (levenshtein.rs:885-895).  Per limb and text
// character: 7 LOP3 + 1 IADD3.X (the add of the recurrence as one carry chain) on the ALU pipe, the two 1-bit shifts
// as IMAD.WIDE (bit 31 falls out as the product's high word) + add on the FMA pipe.
// Match table in shared memory: Q planes of 32 KB, plane q holds limbs 4q..4q+3 of every symbol as one 16-byte
// vector, replicated 8x ([ch][slot][4 limbs], slot = lane & 7) so that each quarter-warp of an LDS.128 hits 8
// different bank groups whatever the 8 symbols are: conflict-free gather, one IDP.4A address for all planes.
template <int L>
__device__ __forceinline__ void add_chain(uint32_t (&s)[L], const uint32_t (&a)[L], const uint32_t (&b)[L]);
template <>
__device__ __forceinline__ void add_chain<4>(uint32_t (&s)[4], const uint32_t (&a)[4], const uint32_t (&b)[4]) {
  asm("{\n\t"
      "add.cc.u32 %0, %4, %8;\n\t"
      "addc.cc.u32 %1, %5, %9;\n\t"
      "addc.cc.u32 %2, %6, %10;\n\t"
      "addc.u32 %3, %7, %11;\n\t"
      "}"
      : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
}
template <>
__device__ __forceinline__ void add_chain<8>(uint32_t (&s)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
  asm("{\n\t"
      "add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;\n\t"
      "}"
      : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
}
template <>
__device__ __forceinline__ void add_chain<12>(uint32_t (&s)[12], const uint32_t (&a)[12], const uint32_t (&b)[12]) {
  asm("{\n\t"
      "add.cc.u32 %0, %12, %24;\n\t"
      "addc.cc.u32 %1, %13, %25;\n\t"
      "addc.cc.u32 %2, %14, %26;\n\t"
      "addc.cc.u32 %3, %15, %27;\n\t"
      "addc.cc.u32 %4, %16, %28;\n\t"
      "addc.cc.u32 %5, %17, %29;\n\t"
      "addc.cc.u32 %6, %18, %30;\n\t"
      "addc.cc.u32 %7, %19, %31;\n\t"
      "addc.cc.u32 %8, %20, %32;\n\t"
      "addc.cc.u32 %9, %21, %33;\n\t"
      "addc.cc.u32 %10, %22, %34;\n\t"
      "addc.u32 %11, %23, %35;\n\t"
      "}"
      : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]), "=r"(s[8]), "=r"(s[9]), "=r"(s[10]), "=r"(s[11])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]), "r"(b[9]), "r"(b[10]), "r"(b[11]));
}
template <>
__device__ __forceinline__ void add_chain<16>(uint32_t (&s)[16], const uint32_t (&a)[16], const uint32_t (&b)[16]) {
  asm("{\n\t"
      "add.cc.u32 %0, %16, %32;\n\t"
      "addc.cc.u32 %1, %17, %33;\n\t"
      "addc.cc.u32 %2, %18, %34;\n\t"
      "addc.cc.u32 %3, %19, %35;\n\t"
      "addc.cc.u32 %4, %20, %36;\n\t"
      "addc.cc.u32 %5, %21, %37;\n\t"
      "addc.cc.u32 %6, %22, %38;\n\t"
      "addc.cc.u32 %7, %23, %39;\n\t"
      "addc.cc.u32 %8, %24, %40;\n\t"
      "addc.cc.u32 %9, %25, %41;\n\t"
      "addc.cc.u32 %10, %26, %42;\n\t"
      "addc.cc.u32 %11, %27, %43;\n\t"
      "addc.cc.u32 %12, %28, %44;\n\t"
      "addc.cc.u32 %13, %29, %45;\n\t"
      "addc.cc.u32 %14, %30, %46;\n\t"
      "addc.u32 %15, %31, %47;\n\t"
      "}"
      : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]), "=r"(s[8]), "=r"(s[9]), "=r"(s[10]), "=r"(s[11]), "=r"(s[12]), "=r"(s[13]), "=r"(s[14]), "=r"(s[15])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]), "r"(b[9]), "r"(b[10]), "r"(b[11]), "r"(b[12]), "r"(b[13]), "r"(b[14]), "r"(b[15]));
}

template <int OFF>
__device__ __forceinline__ void lds128_off(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr), "n"(OFF));
}
template <int Q>
__device__ __forceinline__ void pmn_load(uint32_t addr, uint32_t (&X)[4 * Q]) {
  lds128_off<0>(addr, X[0], X[1], X[2], X[3]);
  if constexpr (Q > 1) lds128_off<32768>(addr, X[4], X[5], X[6], X[7]);
  if constexpr (Q > 2) lds128_off<65536>(addr, X[8], X[9], X[10], X[11]);
  if constexpr (Q > 3) lds128_off<98304>(addr, X[12], X[13], X[14], X[15]);
}
template <int Q, bool OSA>
struct LevNStep {
  static constexpr int L = 4 * Q;
  uint32_t VP[L], VN[L];
  uint32_t D0p[OSA ? L : 1], Xp[OSA ? L : 1];
  uint32_t base, two;
  __device__ __forceinline__ void init(uint32_t len1, uint32_t base_, uint32_t two_) {
    base = base_;
    two = two_;
    const uint32_t sh = 32u * L - len1;  // unused low bits
#pragma unroll
    for (int i = 0; i < L; ++i) {
      VP[i] = (32u * (i + 1) <= sh) ? 0u : (32u * i >= sh ? 0xFFFFFFFFu : 0xFFFFFFFFu << (sh - 32u * i));
      VN[i] = 0;
    }
    if constexpr (OSA) {
#pragma unroll
      for (int i = 0; i < L; ++i) { D0p[i] = 0; Xp[i] = 0; }
    }
  }
  template <int K>
  __device__ __forceinline__ void step(uint32_t w) { step_sel(w, 0x80u << (8 * K)); }
  __device__ __forceinline__ void step_sel(uint32_t w, uint32_t sel) {  // sel = 0x80 << (8 * byte index)
    uint32_t X[L], T[L], S[L];
    pmn_load<Q>(__dp4a(w, sel, base), X);
#pragma unroll
    for (int i = 0; i < L; ++i) T[i] = X[i] & VP[i];
    add_chain<L>(S, T, VP);
    uint32_t cp = two >> 1, cn = 0, ct = 0;  // +1 enters row 0 (through the unused low bits)
#pragma unroll
    for (int i = 0; i < L; ++i) {
      // the 7-LOP3 form of the step, pinned with explicit lop3 (left to itself ptxas re-associates D0 = D0' | VN into its
      // three consumers and ends up with 8): D0' (1) D0 (1) HP (1) HN (1) VP' (1) VN' (1) + the AND above
      uint32_t D0 = lop3<0xBE>(S[i], VP[i], X[i]) | VN[i];  // ((S ^ VP) | X) | VN
      if constexpr (OSA) {  // TR = (((~D0_prev) & X) << 1) & X_prev  (osa.rs:84-135), the shift across the limbs
        const uint32_t t = ~D0p[i] & X[i];
        uint32_t lo, hi;
        if (i < L - 1) { mul2_wide(t, two, lo, hi); lo += ct; ct = hi; }
        else lo = t * two + ct;
        D0 |= lo & Xp[i];
        D0p[i] = D0;
        Xp[i] = X[i];
      }
      const uint32_t HP = lop3<0xF1>(VN[i], D0, VP[i]);  // VN | ~(D0 | VP)
      const uint32_t HN = D0 & VP[i];
      uint32_t HPs, HNs;
      if (i < L - 1) {
        uint32_t lo, hi;
        mul2_wide(HP, two, lo, hi);
        HPs = lo + cp;
        cp = hi;
        mul2_wide(HN, two, lo, hi);
        HNs = lo + cn;
        cn = hi;
      } else {
        HPs = HP * two + cp;
        HNs = HN * two + cn;
      }
      VP[i] = lop3<0xF1>(HNs, D0, HPs);  // HN | ~(D0 | HP)
      VN[i] = HPs & D0;
    }
  }
  __device__ __forceinline__ uint32_t result(uint32_t len2) const {
    uint32_t r = len2;
#pragma unroll
    for (int i = 0; i < L; ++i) r += (uint32_t)__popc(VP[i]) - (uint32_t)__popc(VN[i]);
    return r;
  }
  // D[len1][j] after j columns: the top-aligned form makes the bottom-row value available at any column
  __device__ __forceinline__ uint32_t bottom(uint32_t j) const { return result(j); }
};

// LCS length (lcs_seq.rs:199-261, :267-341) on the same limbs, bottom-aligned table: S = (S + U) | (S & ~U), U = S & X.
template <int Q>
struct LcsNStep {
  static constexpr int L = 4 * Q;
  uint32_t S[L];
  uint32_t base;
  __device__ __forceinline__ void init(uint32_t, uint32_t base_, uint32_t) {
    base = base_;
#pragma unroll
    for (int i = 0; i < L; ++i) S[i] = 0xFFFFFFFFu;
  }
  template <int K>
  __device__ __forceinline__ void step(uint32_t w) { step_sel(w, 0x80u << (8 * K)); }
  __device__ __forceinline__ void step_sel(uint32_t w, uint32_t sel) {
    uint32_t X[L], U[L], A[L];
    pmn_load<Q>(__dp4a(w, sel, base), X);
#pragma unroll
    for (int i = 0; i < L; ++i) U[i] = S[i] & X[i];
    add_chain<L>(A, S, U);
#pragma unroll
    for (int i = 0; i < L; ++i) S[i] = A[i] | (S[i] & ~U[i]);
  }
  __device__ __forceinline__ uint32_t result(uint32_t) const {
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < L; ++i) r += (uint32_t)__popc(~S[i]);
    return r;
  }
};

// CUT (Levenshtein distance with a score_cutoff above 63, unit weights): lanes drop out as soon as their candidate cannot
// come back under the cutoff (the job of the reference's Ukkonen band, levenshtein.rs:897-985); only ever turns a result
// into None.
template <int FAM, int Q, int NT, bool CUT = false>
__global__ void __launch_bounds__(NT) scan_lbn_kernel(const __grid_constant__ LbParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int L = 4 * Q;
  {
    // p.tab: [256][L] u32 limbs of the (top- or bottom-aligned) match vectors -> Q planes, 8 replicas per symbol
    const uint32_t* __restrict__ t = reinterpret_cast<const uint32_t*>(p.tab);
    uint32_t* pm32 = reinterpret_cast<uint32_t*>(smem_raw);
    for (uint32_t i = threadIdx.x; i < 256u * L * 8u; i += NT) {
      const uint32_t ch = i / (L * 8u), r = i % (L * 8u), slot = r / L, limb = r % L;
      pm32[(limb >> 2) * 8192u + ch * 32u + slot * 4u + (limb & 3u)] = t[ch * L + limb];
    }
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t base = smem_u32(smem_raw) + (lane & 7u) * 16u;
  const uint64_t total_warps = (uint64_t)gridDim.x * (NT / 32);
  const uint64_t ngroups = p.lb.ngroups;
  const uint64_t nchunks = (ngroups + p.chunk - 1) / p.chunk;
  const uint2* __restrict__ gdata = reinterpret_cast<const uint2*>(p.lb.gdata);
  uint64_t chunk = (uint64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  while (chunk < nchunks) {
    unsigned long long next_chunk = 0;
    if (lane == 0) next_chunk = total_warps + atomicAdd(p.counter, 1ull);
    const uint64_t g0 = chunk * p.chunk;
    const uint32_t ng = (uint32_t)((g0 + p.chunk < ngroups) ? p.chunk : ngroups - g0);
    const uint32_t* lens_p = p.lb.lens + g0 * 32 + lane;
    const uint32_t* perm_p = p.lb.perm + g0 * 32 + lane;
    const uint2* grp = gdata + __ldg(p.lb.goff + g0) * 32 + lane;
    uint32_t len_n = __ldg(lens_p);
    uint2 first_n = ld_row8<true>(grp);
    uint2 second_n = ld_row8<true>(grp + 32);
    for (uint32_t gi = 0; gi < ng; ++gi) {
      const uint32_t len2 = len_n;
      const uint32_t idx = __ldg(perm_p);
      const LaneSrcT<true> src{grp, first_n, second_n};
      grp += ((__reduce_max_sync(0xffffffffu, len2) + 7u) >> 3) * 32u;
      lens_p += 32;
      perm_p += 32;
      if (gi + 1 < ng) {
        len_n = __ldg(lens_p);
        first_n = ld_row8<true>(grp);
        second_n = ld_row8<true>(grp + 32);
      }
      uint32_t raw;
      bool dead = false;  // CUT only: this lane's candidate is beyond the cutoff for certain
      if constexpr (FAM == F_LCS) {
        LcsNStep<Q> st;
        st.init(p.len1, base, p.two);
        walk_rows8_compact<false>(src.reader(), len2, st);
        raw = st.result(len2);
      } else {
        LevNStep<Q, FAM == F_OSA> st;
        st.init(p.len1, base, p.two);
        if constexpr (CUT) {
          const uint32_t cut = (uint32_t)(p.epi.cutoff_u / p.epi.w_ins > 0x7FFFFFFFull ? 0x7FFFFFFFull : p.epi.cutoff_u / p.epi.w_ins);
          const uint32_t diff = p.len1 > len2 ? p.len1 - len2 : len2 - p.len1;
          dead = diff > cut || !walk_rows8_compact<(Q > 2), true>(src.reader(), len2, st, cut);  // levenshtein.rs:1045-1047
        } else {
          walk_rows8_compact<(Q > 2 || FAM == F_OSA)>(src.reader(), len2, st);
        }
        raw = st.result(len2);
      }
      if (idx != 0xFFFFFFFFu) {
        if (p.out_f64) reinterpret_cast<double*>(p.out)[idx] = finish_norm(p.epi, raw, p.len1, len2);
        else reinterpret_cast<uint32_t*>(p.out)[idx] = dead ? NONE_U32 : finish_int(p.epi, raw, p.len1, len2);
      }
    }
    chunk = __shfl_sync(0xffffffffu, next_chunk, 0);
  }
}

template <int FAM, int Q, bool CUT = false>
static cudaError_t launch_lbn_inst(const ScanLaunch& L, const void* tab) {
  // 64 KB of table per CTA (Q = 2) lets 3 CTAs live on an SM: 320 threads each = 30 warps within the 64 K registers at 68
  constexpr int NT = (Q == 2 && FAM != F_OSA) ? 320 : 256;
  auto kern = scan_lbn_kernel<FAM, Q, NT, CUT>;
  const size_t smem = (size_t)Q * 32768;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  int ctas_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, NT, smem);
  if (e != cudaSuccess) return e;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  LbParams p{};
  p.lb = L.lb;
  p.tab = tab;
  p.len1 = L.query.len1;
  p.out = L.out;
  p.out_f64 = L.out_is_f64;
  p.two = 2;
  p.chunk = 4;
  p.counter = L.lb_counter;
  p.epi = L.epi;
  e = cudaMemsetAsync(p.counter, 0, sizeof(unsigned long long), L.stream);
  if (e != cudaSuccess) return e;
  uint64_t grid = (uint64_t)L.sm_count * ctas_per_sm;
  const uint64_t nchunks = (L.lb.ngroups + p.chunk - 1) / p.chunk;
  const uint64_t need = (nchunks + NT / 32 - 1) / (NT / 32);
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  kern<<<(uint32_t)grid, NT, smem, L.stream>>>(p);
  g_launches.fetch_add(1);
  return cudaGetLastError();
}

template <int FAM>
static cudaError_t launch_lbn_fam(const ScanLaunch& L, const void* tab) {
  if constexpr (FAM == F_LEV) {
    const bool lev_cut = L.epi.metric == M_LEVENSHTEIN && L.epi.kind == K_DISTANCE && L.epi.has_cutoff && L.epi.wclass == WC_UNIFORM &&
                         !L.out_is_f64;
    if (lev_cut) {
      switch (L.query.limbs / 4) {
        case 1: return launch_lbn_inst<FAM, 1, true>(L, tab);
        case 2: return launch_lbn_inst<FAM, 2, true>(L, tab);
        case 3: return launch_lbn_inst<FAM, 3, true>(L, tab);
        case 4: return launch_lbn_inst<FAM, 4, true>(L, tab);
        default: return cudaErrorInvalidValue;
      }
    }
  }
  switch (L.query.limbs / 4) {
    case 1: return launch_lbn_inst<FAM, 1>(L, tab);
    case 2: return launch_lbn_inst<FAM, 2>(L, tab);
    case 3: return launch_lbn_inst<FAM, 3>(L, tab);
    case 4: return launch_lbn_inst<FAM, 4>(L, tab);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_scan_lbn(const ScanLaunch& L) {
  if (!L.query.pmn_top || !L.lb.gdata) return cudaErrorInvalidValue;
  switch (family_of(L.epi.metric, L.epi.wclass)) {
    case F_LEV: return launch_lbn_fam<F_LEV>(L, L.query.pmn_top);
    case F_OSA: return launch_lbn_fam<F_OSA>(L, L.query.pmn_top);
    case F_LCS: return launch_lbn_fam<F_LCS>(L, L.query.pmn_bot);
    default: return cudaErrorInvalidValue;
  }
}

// ------------------------------------------------------------------------------------------------ long
// Queries beyond 16 384 elements (the reference has no limit: its own test_large_band runs 106 514 x 107 244,
// levenshtein.rs:2139-2161).  One warp per candidate, the column cut into STRIPES of 256 64-bit blocks (lane l owns
// blocks 8l..8l+7 of the stripe, columns skewed across the lanes exactly as in scan_mw_kernel).  A stripe is run over
// the whole candidate; the horizontal deltas leaving its bottom row (hp, hn, OSA's transposition bit / the LCS add
// carry: one byte per column) are parked in a per-warp scratch line and enter the next stripe's top row, which plays
// the role of the reference's hp_carry / hn_carry between blocks (levenshtein.rs:838-875) at stripe granularity.
struct LongParams {
  const uint8_t* chars;
  const uint32_t* off32;
  const uint64_t* off64;
  uint64_t n;
  const uint64_t* pm;  // [256][words]
  uint32_t len1;
  uint32_t words;
  uint8_t* scratch;         // [warps][stride] carry bytes
  uint64_t stride;
  void* out;
  int out_f64;
  Epi epi;
};

template <int FAM, class T = uint8_t>
__global__ void __launch_bounds__(128) scan_long_kernel(const __grid_constant__ LongParams p) {
  constexpr uint32_t CH_BITS = sizeof(T) * 8, CH_MASK = (1u << CH_BITS) - 1u;
  constexpr int WPL = 8;
  constexpr uint32_t SW = 32 * WPL;  // blocks per stripe
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warp_global = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint64_t total_warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  const uint32_t words = p.words;
  const uint32_t nstripes = (words + SW - 1) / SW;
  const uint32_t last_bit = (p.len1 - 1u) & 63u;
  const bool off64 = p.off64 != nullptr;
  const uint64_t* __restrict__ pm = p.pm;
  uint8_t* carry = p.scratch + warp_global * p.stride;
  const bool lev_cut = (FAM == F_LEV) && p.epi.metric == M_LEVENSHTEIN && p.epi.kind == K_DISTANCE && p.epi.has_cutoff &&
                       p.epi.wclass == WC_UNIFORM;
  const uint64_t cut64 = lev_cut ? p.epi.cutoff_u / p.epi.w_ins : 0;

  for (uint64_t c = warp_global; c < p.n; c += total_warps) {
    const uint64_t o0 = off64 ? p.off64[c] : (uint64_t)p.off32[c];
    const uint64_t o1 = off64 ? p.off64[c + 1] : (uint64_t)p.off32[c + 1];
    const uint32_t len2 = (uint32_t)(o1 - o0);
    const uint32_t diff = p.len1 > len2 ? p.len1 - len2 : len2 - p.len1;
    if (lev_cut && (uint64_t)diff > cut64) {  // levenshtein.rs:1045-1047
      if (lane == 0) {
        if (p.out_f64) reinterpret_cast<double*>(p.out)[c] = qnan();
        else reinterpret_cast<uint32_t*>(p.out)[c] = NONE_U32;
      }
      continue;
    }
    if (len2 == 0) {
      const uint32_t raw0 = (FAM == F_LCS) ? 0u : p.len1;
      if (lane == 0) {
        if (p.out_f64) reinterpret_cast<double*>(p.out)[c] = finish_norm(p.epi, raw0, p.len1, 0);
        else reinterpret_cast<uint32_t*>(p.out)[c] = finish_int(p.epi, raw0, p.len1, 0);
      }
      continue;
    }
    const T* __restrict__ txt = reinterpret_cast<const T*>(p.chars) + o0;
    int32_t score = 0;
    uint32_t lcs_cnt = 0;
    for (uint32_t s = 0; s < nstripes; ++s) {
      const uint32_t sw0 = s * SW;                                       // first block of the stripe
      const uint32_t swords = words - sw0 < SW ? words - sw0 : SW;      // blocks in this stripe
      const uint32_t act = (swords + WPL - 1) / WPL;                     // lanes that own at least one block
      const uint32_t w0 = sw0 + lane * WPL;
      const bool last_stripe = s + 1 == nstripes;
      const uint32_t last_owner = (swords - 1) / WPL, last_k = (swords - 1) % WPL;
      const uint32_t steps = len2 + act - 1u;
      uint64_t VP[WPL], VN[WPL], D0[WPL], PMo[WPL], Xc[WPL], Xn[WPL];
#pragma unroll
      for (int k = 0; k < WPL; ++k) {
        VP[k] = ~0ull;  // LCS: VP plays the role of S
        VN[k] = 0; D0[k] = 0; PMo[k] = 0; Xc[k] = 0; Xn[k] = 0;
      }
      // lane 0 feeds the text and the carries entering the stripe's top row, both fetched three columns ahead
      const uint32_t top0 = (FAM == F_LCS) ? 0u : 1u;  // matrix top row: horizontal delta +1 (Levenshtein / OSA), no carry (LCS)
      uint32_t ch_cur = 0, c1 = 0, c2 = 0, in0 = top0, in1 = top0, in2 = top0;
      if (lane == 0) {
        ch_cur = txt[0];
        c1 = len2 > 1 ? txt[1] : 0;
        c2 = len2 > 2 ? txt[2] : 0;
        if (s > 0) {
          in0 = __ldcg(carry + 0);
          in1 = len2 > 1 ? __ldcg(carry + 1) : 0;
          in2 = len2 > 2 ? __ldcg(carry + 2) : 0;
        }
#pragma unroll
        for (int k = 0; k < WPL; ++k) Xc[k] = (w0 + k < words) ? __ldg(pm + (uint64_t)ch_cur * words + w0 + k) : 0ull;
      }
      uint32_t cout = 0;
      for (uint32_t t = 0; t < steps; ++t) {
        const uint32_t pk_in = __shfl_up_sync(0xffffffffu, ch_cur | (cout << CH_BITS), 1);
        uint32_t ch_nxt, cin;
        if (lane == 0) {
          ch_nxt = c1;
          c1 = c2;
          c2 = (t + 3 < len2) ? (uint32_t)txt[t + 3] : 0u;
          cin = in0;
          in0 = in1;
          in1 = in2;
          in2 = (s > 0 && t + 3 < len2) ? (uint32_t)__ldcg(carry + t + 3) : top0;
        } else {
          ch_nxt = pk_in & CH_MASK;
          cin = pk_in >> CH_BITS;
        }
#pragma unroll
        for (int k = 0; k < WPL; ++k) Xn[k] = (w0 + k < words) ? __ldg(pm + (uint64_t)ch_nxt * words + w0 + k) : 0ull;
        const int32_t j = (int32_t)t - (int32_t)lane;  // text column handled by this lane in this step
        const bool active = (lane < act) && j >= 0 && j < (int32_t)len2;
        if (active) {
          if constexpr (FAM == F_LCS) {
            uint64_t cy = cin & 1u;
#pragma unroll
            for (int k = 0; k < WPL; ++k) {
              if (w0 + k < words) {
                const uint64_t S = VP[k];
                const uint64_t u = S & Xc[k];
                const uint64_t x1 = S + u;
                const uint64_t x2 = x1 + cy;
                cy = (uint64_t)(x1 < S) | (uint64_t)(x2 < x1);
                VP[k] = x2 | (S - u);
              }
            }
            cout = (uint32_t)cy;
          } else {
            uint64_t hp_c = cin & 1u, hn_c = (cin >> 1) & 1u, tr_c = (cin >> 2) & 1u;
#pragma unroll
            for (int k = 0; k < WPL; ++k) {
              if (w0 + k < words) {
                const uint64_t X0 = Xc[k];
                const uint64_t X = X0 | hn_c;
                uint64_t d0 = ((((X & VP[k]) + VP[k]) ^ VP[k]) | X) | VN[k];
                if constexpr (FAM == F_OSA) {
                  const uint64_t nd = (~D0[k]) & X0;
                  d0 |= ((nd << 1) | tr_c) & PMo[k];
                  tr_c = nd >> 63;
                  D0[k] = d0;
                  PMo[k] = X0;
                }
                uint64_t HP = VN[k] | ~(d0 | VP[k]);
                uint64_t HN = d0 & VP[k];
                if (last_stripe && lane == last_owner && k == (int)last_k)
                  score += (int32_t)((HP >> last_bit) & 1u) - (int32_t)((HN >> last_bit) & 1u);
                const uint64_t hp_o = HP >> 63, hn_o = HN >> 63;
                HP = (HP << 1) | hp_c;
                HN = (HN << 1) | hn_c;
                VP[k] = HN | ~(d0 | HP);
                VN[k] = HP & d0;
                hp_c = hp_o;
                hn_c = hn_o;
              }
            }
            cout = (uint32_t)hp_c | ((uint32_t)hn_c << 1) | ((uint32_t)tr_c << 2);
          }
          // the stripe's bottom row: park the deltas of column j for the next stripe (read there >= 29 steps of this
          // loop later at the earliest for the same column -- it runs after this stripe has finished)
          if (!last_stripe && lane == act - 1u) __stcg(carry + j, (uint8_t)cout);
        }
        ch_cur = ch_nxt;
#pragma unroll
        for (int k = 0; k < WPL; ++k) Xc[k] = Xn[k];
      }
      if constexpr (FAM == F_LCS) {
#pragma unroll
        for (int k = 0; k < WPL; ++k)
          if (lane < act && w0 + k < words) lcs_cnt += (uint32_t)__popcll(~VP[k]);
      }
      if (last_stripe) score = __shfl_sync(0xffffffffu, score, last_owner);
      __syncwarp();  // orders this stripe's carry stores before the next stripe's loads (same warp)
    }
    uint32_t raw;
    if constexpr (FAM == F_LCS) {
      for (uint32_t d = 16; d >= 1; d >>= 1) lcs_cnt += __shfl_xor_sync(0xffffffffu, lcs_cnt, d);
      raw = lcs_cnt;
    } else {
      raw = (uint32_t)((int32_t)p.len1 + score);
    }
    if (lane == 0) {
      if (p.out_f64) reinterpret_cast<double*>(p.out)[c] = finish_norm(p.epi, raw, p.len1, len2);
      else reinterpret_cast<uint32_t*>(p.out)[c] = finish_int(p.epi, raw, p.len1, len2);
    }
  }
}

template <int FAM>
static cudaError_t launch_long_fam(const ScanLaunch& L) {
  LongParams p{};
  p.chars = L.corpus.chars;
  p.off32 = L.corpus.off32;
  p.off64 = L.corpus.off64;
  p.n = L.corpus.n;
  p.pm = L.query.pm_words;
  p.len1 = L.query.len1;
  p.words = L.query.words;
  p.out = L.out;
  p.out_f64 = L.out_is_f64;
  p.epi = L.epi;
  // one carry byte per column and warp; no candidate is longer than the corpus.  The scratch is capped at 1 GiB, the
  // grid follows (a warp per candidate up to that many warps)
  uint64_t max_len = L.corpus.max_len ? L.corpus.max_len : L.corpus.total;
  if (max_len > L.corpus.total) max_len = L.corpus.total;
  p.stride = (max_len + 63) / 64 * 64 + 64;
  uint64_t warps = (1ull << 30) / p.stride;
  const uint64_t cap = (uint64_t)L.sm_count * 16;  // 4 CTAs of 4 warps per SM
  if (warps > cap) warps = cap;
  if (warps > p.n) warps = p.n;
  if (warps < 1) warps = 1;
  const uint32_t blocks = (uint32_t)((warps + 3) / 4);
  cudaError_t e = dev_alloc(&p.scratch, (uint64_t)blocks * 4 * p.stride, L.stream);
  if (e != cudaSuccess) return e;
  if (L.elem16) scan_long_kernel<FAM, uint16_t><<<blocks, 128, 0, L.stream>>>(p);
  else scan_long_kernel<FAM><<<blocks, 128, 0, L.stream>>>(p);
  g_launches.fetch_add(1);
  e = cudaGetLastError();
  dev_free(p.scratch, L.stream);
  return e;
}

cudaError_t launch_scan_long(const ScanLaunch& L) {
  switch (family_of(L.epi.metric, L.epi.wclass)) {
    case F_LEV: return launch_long_fam<F_LEV>(L);
    case F_OSA: return launch_long_fam<F_OSA>(L);
    case F_LCS: return launch_long_fam<F_LCS>(L);
    default: return cudaErrorInvalidValue;
  }
}

// ------------------------------------------------------------------------------------------------ band
// Levenshtein distance with score_cutoff k <= 63 for multi-word queries (BASELINE config 3).  The Ukkonen band
// of a cutoff-k problem has at most k+1 diagonals, so one 64-bit sliding window per candidate (LevBand64,
// rf_core.cuh) replaces the ceil(len1/64)-word column of the block algorithm: one THREAD per candidate, no
// shuffles.  Three dense passes, each a grid-stride loop of full warps:
//   classify  length filter |len1-len2| > k -> None (levenshtein.rs:1045-1047); survivors appended to list A
//   run<A>    first `cols_a` columns of every list-A candidate, early exit once the end diagonal exceeds k
//             (random candidates die here); finished ones are written, the still-alive go to list B
//   run<B>    list B (the true near-matches) start to end, 32 per warp
// The query's match vectors (one zero word of padding on both sides of each row) sit in shared memory.
struct BandParams {
  const uint8_t* chars;
  const uint32_t* off32;
  const uint64_t* off64;
  uint64_t i0, i1;        // candidate range of this batch
  const uint32_t* pmb;    // [256][stride] match vectors as 32-bit words, 2 zero words in front, >= 2 behind
  uint32_t stride;        // odd: the 32 lanes' rows spread over all shared-memory banks
  uint32_t len1;
  uint32_t cut;           // cutoff in unit-cost edits, <= 63
  uint32_t maxcols;       // columns to run in this pass
  const uint32_t* list_in;
  const uint32_t* cnt_in;
  uint32_t* list_out;
  uint32_t* cnt_out;
  uint32_t* out;
  Epi epi;
};

__device__ __forceinline__ uint64_t band_off(const BandParams& p, uint64_t i) {
  return p.off64 ? p.off64[i] : (uint64_t)p.off32[i];
}

// warp-aggregated append of the lanes with `flag` set
__device__ __forceinline__ void band_append(bool flag, uint32_t value, uint32_t* list, uint32_t* cnt, uint32_t lane) {
  const uint32_t m = __ballot_sync(0xffffffffu, flag);
  if (m == 0) return;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(cnt, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (flag) list[base + __popc(m & ((1u << lane) - 1u))] = value;
}

// 256 threads x 8 candidates per CTA iteration; ONE global atomic per 2048 candidates (same-address atomics
// serialise in L2: one per warp of 32 made this pass atomic-bound)
constexpr int BAND_CL_ITEMS = 8;
__global__ void __launch_bounds__(256) band_classify_kernel(const __grid_constant__ BandParams p) {
  __shared__ uint32_t wcnt[8];
  __shared__ uint32_t cta_base;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint64_t nb = p.i1 - p.i0;
  constexpr uint64_t PER_CTA = 256ull * BAND_CL_ITEMS;
  for (uint64_t t0 = (uint64_t)blockIdx.x * PER_CTA; t0 < nb; t0 += (uint64_t)gridDim.x * PER_CTA) {
    uint32_t needm = 0, mine = 0;  // per lane: bit k = item k needs scoring
#pragma unroll
    for (int k = 0; k < BAND_CL_ITEMS; ++k) {
      const uint64_t i = t0 + (uint64_t)k * 256 + threadIdx.x;
      if (i < nb) {
        const uint64_t c = p.i0 + i;
        const uint64_t o0 = band_off(p, c);
        const uint32_t len2 = (uint32_t)(band_off(p, c + 1) - o0);
        const uint32_t diff = p.len1 > len2 ? p.len1 - len2 : len2 - p.len1;
        if (diff > p.cut) p.out[c] = NONE_U32;
        else if (len2 == 0) p.out[c] = finish_int(p.epi, p.len1, p.len1, 0);
        else { needm |= 1u << k; ++mine; }
      }
    }
    // exclusive prefix of `mine` over the CTA
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= (uint32_t)d) incl += t;
    }
    if (lane == 31) wcnt[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t tot = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) { const uint32_t v = wcnt[w]; wcnt[w] = tot; tot += v; }
      cta_base = tot ? atomicAdd(p.cnt_out, tot) : 0u;
    }
    __syncthreads();
    uint32_t pos = cta_base + wcnt[warp] + incl - mine;
#pragma unroll
    for (int k = 0; k < BAND_CL_ITEMS; ++k)
      if (needm & (1u << k)) p.list_out[pos++] = (uint32_t)(t0 + (uint64_t)k * 256 + threadIdx.x);
    __syncthreads();
  }
}

// NARROW: cutoff <= 32, the band's <= 33 bits come from two table words instead of three.
template <bool FINAL, bool PM_SMEM, bool NARROW>
__global__ void __launch_bounds__(256) band_run_kernel(const __grid_constant__ BandParams p) {
  extern __shared__ __align__(16) uint32_t band_pm_s[];
  if constexpr (PM_SMEM) {
    for (uint32_t i = threadIdx.x; i < 256u * p.stride; i += blockDim.x) band_pm_s[i] = p.pmb[i];
    __syncthreads();
  }
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t cnt = *p.cnt_in;
  const uint32_t rounded = (cnt + 31u) / 32u * 32u;
  const uint32_t cut = p.cut;
  const uint32_t stride = p.stride;
  for (uint32_t li = blockIdx.x * blockDim.x + threadIdx.x; li < rounded; li += gridDim.x * blockDim.x) {
    const bool have = li < cnt;
    const uint32_t rel = have ? p.list_in[li] : 0u;  // idle lanes shadow the batch's first candidate with 0 columns
    const uint64_t c = p.i0 + rel;
    const uint64_t o0 = band_off(p, c);
    const uint32_t len2 = have ? (uint32_t)(band_off(p, c + 1) - o0) : 0u;
    LevBand64 b;
    b.init(p.len1, have ? len2 : p.len1, cut);
    const uint32_t cols = have ? (len2 < p.maxcols ? len2 : p.maxcols) : 0u;
    // each lane walks its own candidate: one 16-byte load per 16 columns
    ByteReader16 rd(p.chars + (o0 & ~15ull), (uint32_t)(o0 & 15ull));
    bool dead = false;
    auto column = [&](uint32_t ch) {
      const uint32_t sp = (uint32_t)(b.s + 64);
      const uint32_t idx = ch * stride + (sp >> 5);
      uint32_t w0, w1, w2 = 0;
      if constexpr (PM_SMEM) {
        w0 = band_pm_s[idx]; w1 = band_pm_s[idx + 1];
        if constexpr (!NARROW) w2 = band_pm_s[idx + 2];
      } else {
        w0 = __ldg(p.pmb + idx); w1 = __ldg(p.pmb + idx + 1);
        if constexpr (!NARROW) w2 = __ldg(p.pmb + idx + 2);
      }
      if constexpr (NARROW) b.step(band_window32_low33(w0, w1, sp & 31u));
      else b.step(band_window32(w0, w1, w2, sp & 31u));
    };
    auto eight = [&](uint32_t wa, uint32_t wb, uint32_t j0) {  // columns j0 .. j0+7 (clipped to cols), then the death test
      if (dead || j0 >= cols) return;
      if (cols - j0 >= 8) {  // full block: no per-column predicates
        column(wa & 0xffu); column((wa >> 8) & 0xffu); column((wa >> 16) & 0xffu); column(wa >> 24);
        column(wb & 0xffu); column((wb >> 8) & 0xffu); column((wb >> 16) & 0xffu); column(wb >> 24);
      } else {
        const uint32_t nb = cols - j0;
#pragma unroll
        for (int t = 0; t < 7; ++t)
          if ((uint32_t)t < nb) column(((t < 4 ? wa : wb) >> (8 * (t & 3))) & 0xffu);
      }
      dead = b.score() > (int32_t)cut;
    };
    const uint32_t steps = __reduce_max_sync(0xffffffffu, cols);
    for (uint32_t j0 = 0; j0 < steps; j0 += 16) {
      if (!dead && j0 < cols) {
        const Bytes16 t = rd.next16();
        eight(t.w[0], t.w[1], j0);
        eight(t.w[2], t.w[3], j0 + 8);
      }
      if (__ballot_sync(0xffffffffu, !dead && j0 + 16 < cols) == 0) break;
    }
    bool more = false;
    if (have) {
      if (dead) p.out[c] = NONE_U32;
      else if (FINAL || len2 <= p.maxcols) p.out[c] = finish_int(p.epi, (uint64_t)b.score(), p.len1, len2);
      else more = true;
    }
    if constexpr (!FINAL) band_append(more, rel, p.list_out, p.cnt_out, lane);
  }
}

cudaError_t launch_scan_band(const ScanLaunch& L, uint32_t cut) {
  constexpr uint64_t kBatch = 1ull << 24;  // candidates per batch: bounds the two index lists to 64 MiB each
  const uint64_t n = L.corpus.n;
  const uint64_t cap = n < kBatch ? n : kBatch;
  uint32_t* scratch = nullptr;  // [listA cap][listB cap][cntA][cntB] per batch parity
  cudaError_t e = dev_alloc(&scratch, (2 * cap + 16) * sizeof(uint32_t), L.stream);
  if (e != cudaSuccess) return e;
  uint32_t* listA = scratch;
  uint32_t* listB = scratch + cap;
  uint32_t* cnts = scratch + 2 * cap;
  BandParams p{};
  p.chars = L.corpus.chars;
  p.off32 = L.corpus.off32;
  p.off64 = L.corpus.off64;
  p.pmb = L.query.pm_band;
  p.stride = L.query.band_stride;
  p.len1 = L.query.len1;
  p.cut = cut;
  p.out = reinterpret_cast<uint32_t*>(L.out);
  p.epi = L.epi;
  const size_t pm_bytes = (size_t)256 * p.stride * sizeof(uint32_t);
  const bool pm_smem = pm_bytes <= 96 * 1024;  // queries up to ~5900 elements; longer ones read the table through L1
  const size_t smem = pm_smem ? pm_bytes : 0;
  const bool narrow = cut <= 32;
  void (*run_a)(BandParams);
  void (*run_b)(BandParams);
  if (pm_smem) {
    run_a = narrow ? band_run_kernel<false, true, true> : band_run_kernel<false, true, false>;
    run_b = narrow ? band_run_kernel<true, true, true> : band_run_kernel<true, true, false>;
  } else {
    run_a = narrow ? band_run_kernel<false, false, true> : band_run_kernel<false, false, false>;
    run_b = narrow ? band_run_kernel<true, false, true> : band_run_kernel<true, false, false>;
  }
  if (smem > 48 * 1024) {
    e = cudaFuncSetAttribute(run_a, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(run_b, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { dev_free(scratch, L.stream); return e; }
  }
  // columns of the first pass: random candidates leave the band after about k columns
  const uint32_t cols_a = (cut + cut / 4 + 8 + 15) / 16 * 16;
  const uint32_t grid_max = (uint32_t)resident_ctas(run_a, 256, smem, L.sm_count, 4);  // one wave (see resident_ctas)
  const uint32_t grid_cl = (uint32_t)resident_ctas(band_classify_kernel, 256, 0, L.sm_count, 8);
  for (uint64_t i0 = 0; i0 < n && e == cudaSuccess; i0 += kBatch) {
    p.i0 = i0;
    p.i1 = (i0 + kBatch < n) ? i0 + kBatch : n;
    e = cudaMemsetAsync(cnts, 0, 16 * sizeof(uint32_t), L.stream);
    if (e != cudaSuccess) break;
    const uint64_t nb = p.i1 - p.i0;
    const uint64_t cl_ctas = (nb + 256 * BAND_CL_ITEMS - 1) / (256 * BAND_CL_ITEMS);
    const uint32_t grid = (uint32_t)((nb + 255) / 256 < grid_max ? (nb + 255) / 256 : grid_max);
    p.list_out = listA; p.cnt_out = cnts;
    band_classify_kernel<<<(uint32_t)(cl_ctas < grid_cl ? cl_ctas : grid_cl), 256, 0, L.stream>>>(p);
    p.list_in = listA; p.cnt_in = cnts; p.list_out = listB; p.cnt_out = cnts + 1; p.maxcols = cols_a;
    run_a<<<grid, 256, smem, L.stream>>>(p);
    p.list_in = listB; p.cnt_in = cnts + 1; p.list_out = nullptr; p.cnt_out = nullptr; p.maxcols = 0xFFFFFFFFu;
    run_b<<<grid, 256, smem, L.stream>>>(p);
    g_launches.fetch_add(3);
    e = cudaGetLastError();
  }
  dev_free(scratch, L.stream);
  return e;
}

// ------------------------------------------------------------------------------------------------ simple
// Hamming / Prefix / Postfix over the CSR corpus: thread per candidate, query bytes in shared memory, the candidate
// walked with 16-byte loads.  HBM-bound (every candidate byte is read once, nothing else scales).
struct SimpleParams {
  const uint8_t* chars;
  const uint32_t* off32;
  const uint64_t* off64;
  uint64_t n;
  const uint8_t* qbytes;  // query, padded with zeros to a multiple of 16 (+16)
  uint32_t len1;
  void* out;
  int out_f64;
  uint32_t* err_flag;     // set to 1 when Hamming without pad meets a candidate of another length (may be NULL)
  Epi epi;
  LbView lb;              // hamming_lb_kernel
};

__global__ void __launch_bounds__(256) simple_kernel(const __grid_constant__ SimpleParams p) {
  extern __shared__ __align__(16) uint32_t simple_q[];
  const uint32_t qwords = (p.len1 + 3) / 4 + 4;
  for (uint32_t i = threadIdx.x; i < qwords; i += blockDim.x) simple_q[i] = reinterpret_cast<const uint32_t*>(p.qbytes)[i];
  __syncthreads();
  const uint8_t* qb8 = reinterpret_cast<const uint8_t*>(simple_q);
  const bool off64 = p.off64 != nullptr;
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < p.n; c += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t o0 = off64 ? p.off64[c] : (uint64_t)p.off32[c];
    const uint64_t o1 = off64 ? p.off64[c + 1] : (uint64_t)p.off32[c + 1];
    const uint32_t len2 = (uint32_t)(o1 - o0);
    bool none = false;
    uint32_t raw = 0;
    if (p.epi.metric == M_POSTFIX) {
      const uint8_t* t = p.chars + o0;
      raw = postfix_raw([&](uint32_t j) -> uint32_t { return qb8[j]; }, [&](uint32_t j) -> uint32_t { return t[j]; }, p.len1, len2);
    } else {
      ByteReader16 rd(p.chars + (o0 & ~15ull), (uint32_t)(o0 & 15ull));
      Bytes16 blk{};
      uint32_t have = 0xFFFFFFFFu;  // index (in 16-byte blocks) of the block held in blk
      auto t4 = [&](uint32_t w) -> uint32_t {  // words are requested in increasing order
        if ((w >> 2) != have) { blk = rd.next16(); have = w >> 2; }
        return blk.w[w & 3u];
      };
      auto q4 = [&](uint32_t w) -> uint32_t { return simple_q[w]; };
      if (p.epi.metric == M_HAMMING) {
        if (!p.epi.pad && len2 != p.len1) {
          none = true;
          if (p.err_flag) atomicOr(p.err_flag, 1u);
        } else {
          raw = hamming_raw(q4, t4, p.len1, len2);
        }
      } else {
        raw = prefix_raw(q4, t4, p.len1, len2);
      }
    }
    if (p.out_f64) reinterpret_cast<double*>(p.out)[c] = none ? qnan() : finish_norm(p.epi, raw, p.len1, len2);
    else reinterpret_cast<uint32_t*>(p.out)[c] = none ? NONE_U32 : finish_int(p.epi, raw, p.len1, len2);
  }
}

// Hamming over the interleaved layout: warp per group, every row load is 256 contiguous bytes (the CSR kernel's
// 16-byte per-thread loads at a 36-byte stride touch every sector several times), four rows in flight per lane.
__global__ void __launch_bounds__(256) hamming_lb_kernel(const __grid_constant__ SimpleParams p) {
  extern __shared__ __align__(16) uint32_t simple_q[];
  const uint32_t qwords = (p.len1 + 3) / 4 + 4;
  for (uint32_t i = threadIdx.x; i < qwords; i += blockDim.x) simple_q[i] = reinterpret_cast<const uint32_t*>(p.qbytes)[i];
  __syncthreads();
  const uint2* q2 = reinterpret_cast<const uint2*>(simple_q);
  const uint32_t lane = threadIdx.x & 31u, len1 = p.len1;
  const uint2* __restrict__ gdata = reinterpret_cast<const uint2*>(p.lb.gdata);
  const uint64_t nwarps = (uint64_t)gridDim.x * 8, w0 = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  auto diff2 = [](uint2 a, uint2 b) -> uint32_t { return differing_bytes(a.x, b.x) + differing_bytes(a.y, b.y); };
  // the next group's metadata is requested before the current group's rows: one DRAM latency per group instead of two
  uint32_t len_n = 0, idx_n = 0xFFFFFFFFu;
  uint64_t off_n = 0;
  if (w0 < p.lb.ngroups) {
    len_n = __ldg(p.lb.lens + w0 * 32 + lane);
    idx_n = __ldg(p.lb.perm + w0 * 32 + lane);
    off_n = __ldg(p.lb.goff + w0);
  }
  for (uint64_t g = w0; g < p.lb.ngroups; g += nwarps) {
    const uint32_t len2 = len_n, idx = idx_n;
    const uint2* col = gdata + off_n * 32 + lane;
    if (g + nwarps < p.lb.ngroups) {
      len_n = __ldg(p.lb.lens + (g + nwarps) * 32 + lane);
      idx_n = __ldg(p.lb.perm + (g + nwarps) * 32 + lane);
      off_n = __ldg(p.lb.goff + g + nwarps);
    }
    if (idx == 0xFFFFFFFFu) continue;
    const uint32_t mn = len1 < len2 ? len1 : len2, mx = len1 < len2 ? len2 : len1;
    bool none = false;
    uint32_t dist = mx - mn;  // with pad: the excess counts as mismatches (hamming.rs:156-158)
    if (!p.epi.pad && len2 != len1) {
      none = true;
      if (p.err_flag) atomicOr(p.err_flag, 1u);
    } else {
      const uint32_t nfull = mn >> 3;
      uint32_t r = 0;
      for (; r + 4 <= nfull; r += 4) {
        const uint2 a = __ldcs(col + (size_t)r * 32), b = __ldcs(col + (size_t)(r + 1) * 32);
        const uint2 c = __ldcs(col + (size_t)(r + 2) * 32), d = __ldcs(col + (size_t)(r + 3) * 32);
        dist += diff2(a, q2[r]) + diff2(b, q2[r + 1]) + diff2(c, q2[r + 2]) + diff2(d, q2[r + 3]);
      }
      for (; r < nfull; ++r) dist += diff2(__ldcs(col + (size_t)r * 32), q2[r]);
      const uint32_t rem = mn & 7u;
      if (rem) {
        uint2 v = __ldcs(col + (size_t)r * 32), qv = q2[r];
        const uint64_t mask = (1ull << (8 * rem)) - 1ull;
        const uint32_t ml = (uint32_t)mask, mh = (uint32_t)(mask >> 32);
        dist += differing_bytes(v.x & ml, qv.x & ml) + differing_bytes(v.y & mh, qv.y & mh);
      }
    }
    if (p.out_f64) reinterpret_cast<double*>(p.out)[idx] = none ? qnan() : finish_norm(p.epi, dist, len1, len2);
    else reinterpret_cast<uint32_t*>(p.out)[idx] = none ? NONE_U32 : finish_int(p.epi, dist, len1, len2);
  }
}

cudaError_t launch_simple(const ScanLaunch& L, uint32_t* err_flag) {
  if (L.lb.gdata != nullptr && L.epi.metric == M_HAMMING) {
    SimpleParams q{};
    q.qbytes = L.query.qbytes;
    q.len1 = L.query.len1;
    q.out = L.out;
    q.out_f64 = L.out_is_f64;
    q.err_flag = err_flag;
    q.epi = L.epi;
    q.lb = L.lb;
    const size_t smem = ((size_t)(q.len1 + 3) / 4 + 4) * 4 + 8;
    uint64_t blocks = (L.lb.ngroups + 7) / 8;
    const uint64_t max_blocks = resident_ctas(hamming_lb_kernel, 256, smem, L.sm_count, 4);
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    hamming_lb_kernel<<<(uint32_t)blocks, 256, smem, L.stream>>>(q);
    g_launches.fetch_add(1);
    return cudaGetLastError();
  }
  SimpleParams p{};
  p.chars = L.corpus.chars;
  p.off32 = L.corpus.off32;
  p.off64 = L.corpus.off64;
  p.n = L.corpus.n;
  p.qbytes = L.query.qbytes;
  p.len1 = L.query.len1;
  p.out = L.out;
  p.out_f64 = L.out_is_f64;
  p.err_flag = err_flag;
  p.epi = L.epi;
  const size_t smem = ((size_t)(p.len1 + 3) / 4 + 4) * 4;
  uint64_t blocks = (p.n + 255) / 256;
  const uint64_t max_blocks = resident_ctas(simple_kernel, 256, smem, L.sm_count, 4);
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  simple_kernel<<<(uint32_t)blocks, 256, smem, L.stream>>>(p);
  g_launches.fetch_add(1);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ wf
// Generic Levenshtein weights (neither uniform nor insertion/deletion-only): Wagner-Fischer, thread per candidate,
// the cost row in thread-strided global scratch (entry i of thread t at scratch[i*T + t]: coalesced across a warp).
struct WfParams {
  const uint8_t* chars;
  const uint32_t* off32;
  const uint64_t* off64;
  uint64_t n;
  const uint8_t* qbytes;
  uint32_t len1;
  uint64_t* scratch;  // [(len1+1)][T]
  uint32_t T;         // threads in the grid
  void* out;
  int out_f64;
  Epi epi;
  LbView lb;          // wf_lb_kernel
};

__global__ void __launch_bounds__(128) wf_kernel(const __grid_constant__ WfParams p) {
  extern __shared__ __align__(16) uint8_t wf_q[];
  for (uint32_t i = threadIdx.x; i < p.len1; i += blockDim.x) wf_q[i] = p.qbytes[i];
  __syncthreads();
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t* row = p.scratch + t;
  const uint32_t T = p.T;
  const bool off64 = p.off64 != nullptr;
  for (uint64_t c = t; c < p.n; c += T) {
    const uint64_t o0 = off64 ? p.off64[c] : (uint64_t)p.off32[c];
    const uint64_t o1 = off64 ? p.off64[c + 1] : (uint64_t)p.off32[c + 1];
    const uint32_t len2 = (uint32_t)(o1 - o0);
    const uint8_t* txt = p.chars + o0;
    const uint64_t raw = weighted_wagner_fischer([&](uint32_t i) -> uint32_t { return wf_q[i]; },
                                                 [&](uint32_t j) -> uint32_t { return txt[j]; }, p.len1, len2, p.epi.w_ins,
                                                 p.epi.w_del, p.epi.w_sub,
                                                 [&](uint32_t i) -> uint64_t& { return row[(size_t)i * T]; });
    if (p.out_f64) reinterpret_cast<double*>(p.out)[c] = finish_norm(p.epi, raw, p.len1, len2);
    else reinterpret_cast<uint32_t*>(p.out)[c] = finish_int(p.epi, raw, p.len1, len2);
  }
}

// Queries of at most 64 elements over the interleaved layout: warp per group of equal-length candidates, the cost row
// of every thread in shared memory (thread-strided, 64-bit cells as in the reference's usize arithmetic).
constexpr int WF_NT = 128;
__global__ void __launch_bounds__(WF_NT) wf_lb_kernel(const __grid_constant__ WfParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* rows = reinterpret_cast<uint64_t*>(smem_raw);  // [len1 + 1][WF_NT]
  uint8_t* q = reinterpret_cast<uint8_t*>(rows + (size_t)(p.len1 + 1) * WF_NT);
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  for (uint32_t i = tid; i < p.len1; i += WF_NT) q[i] = p.qbytes[i];
  __syncthreads();
  uint64_t* my_row = rows + tid;
  const uint2* __restrict__ gdata = reinterpret_cast<const uint2*>(p.lb.gdata);
  const uint64_t nwarps = (uint64_t)gridDim.x * (WF_NT / 32), w0 = (uint64_t)blockIdx.x * (WF_NT / 32) + (tid >> 5);
  for (uint64_t g = w0; g < p.lb.ngroups; g += nwarps) {
    const uint32_t len2 = __ldg(p.lb.lens + g * 32 + lane);
    const uint32_t idx = __ldg(p.lb.perm + g * 32 + lane);
    if (idx == 0xFFFFFFFFu) continue;
    const uint8_t* col = reinterpret_cast<const uint8_t*>(gdata + __ldg(p.lb.goff + g) * 32 + lane);
    const uint64_t raw = weighted_wagner_fischer([&](uint32_t i) -> uint32_t { return q[i]; },
                                                 [&](uint32_t j) -> uint32_t { return col[(size_t)(j >> 3) * 256 + (j & 7u)]; },
                                                 p.len1, len2, p.epi.w_ins, p.epi.w_del, p.epi.w_sub,
                                                 [&](uint32_t i) -> uint64_t& { return my_row[(size_t)i * WF_NT]; });
    if (p.out_f64) reinterpret_cast<double*>(p.out)[idx] = finish_norm(p.epi, raw, p.len1, len2);
    else reinterpret_cast<uint32_t*>(p.out)[idx] = finish_int(p.epi, raw, p.len1, len2);
  }
}

// Queries beyond 64 elements (thread per candidate, DP rows in global scratch): the grid shrinks with the query so that the
// rows of all threads stay within kDpScratchBytes, the query bytes sit in (opt-in) shared memory up to kDpMaxQuery elements.
constexpr uint64_t kDpScratchBytes = 1ull << 31;
static uint64_t dp_blocks(uint64_t n, uint64_t bytes_per_thread, uint64_t max_blocks) {
  uint64_t blocks = (n + 127) / 128;
  if (blocks > max_blocks) blocks = max_blocks;
  const uint64_t fit = kDpScratchBytes / (bytes_per_thread * 128);
  if (blocks > fit) blocks = fit;
  return blocks < 1 ? 1 : blocks;
}

cudaError_t launch_wf(const ScanLaunch& L) {
  if (L.query.len1 > kDpMaxQuery) return cudaErrorNotSupported;
  if (L.lb.gdata != nullptr && L.query.len1 <= 64) {
    WfParams q{};
    q.qbytes = L.query.qbytes;
    q.len1 = L.query.len1;
    q.out = L.out;
    q.out_f64 = L.out_is_f64;
    q.epi = L.epi;
    q.lb = L.lb;
    const size_t smem = (size_t)(q.len1 + 1) * WF_NT * sizeof(uint64_t) + q.len1 + 16;
    cudaError_t e = cudaFuncSetAttribute(wf_lb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int ctas_per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, wf_lb_kernel, WF_NT, smem);
    if (e != cudaSuccess) return e;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    uint64_t grid = (uint64_t)L.sm_count * ctas_per_sm;
    const uint64_t need = (L.lb.ngroups + WF_NT / 32 - 1) / (WF_NT / 32);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    wf_lb_kernel<<<(uint32_t)grid, WF_NT, smem, L.stream>>>(q);
    g_launches.fetch_add(1);
    return cudaGetLastError();
  }
  WfParams p{};
  p.chars = L.corpus.chars;
  p.off32 = L.corpus.off32;
  p.off64 = L.corpus.off64;
  p.n = L.corpus.n;
  p.qbytes = L.query.qbytes;
  p.len1 = L.query.len1;
  p.out = L.out;
  p.out_f64 = L.out_is_f64;
  p.epi = L.epi;
  const uint64_t blocks = dp_blocks(p.n, (uint64_t)(p.len1 + 1) * sizeof(uint64_t), (uint64_t)L.sm_count * 4);
  p.T = (uint32_t)blocks * 128u;
  cudaError_t e = cudaSuccess;
  if (p.len1 + 16 > 48 * 1024) e = cudaFuncSetAttribute(wf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(p.len1 + 16));
  if (e != cudaSuccess) return e;
  e = dev_alloc(&p.scratch, (size_t)(p.len1 + 1) * p.T * sizeof(uint64_t), L.stream);
  if (e != cudaSuccess) return e;
  wf_kernel<<<(uint32_t)blocks, 128, p.len1 + 16, L.stream>>>(p);
  g_launches.fetch_add(1);
  e = cudaGetLastError();
  dev_free(p.scratch, L.stream);
  return e;
}

// ------------------------------------------------------------------------------------------------ dl
// Damerau-Levenshtein: thread per candidate, the three DP rows (len1+2 entries each) and the 256-entry last-row
// table in thread-strided global scratch (entry e of thread t at scratch[e*T + t]).
struct DlParams {
  const uint8_t* chars;
  const uint32_t* off32;
  const uint64_t* off64;
  uint64_t n;
  const uint8_t* qbytes;
  uint32_t len1;
  int32_t* scratch;  // [(3*(len1+2) + 256)][T]; the last 256 rows (last-row table) start out as -1
  uint32_t T;
  void* out;
  int out_f64;
  Epi epi;
  LbView lb;                    // dl_lb_kernel: the interleaved layout
  unsigned long long* flag;     // dl_lb_kernel raises it for candidates it leaves to dl_kernel (longer than kDlLbMaxLen)
  int only_long;                // dl_kernel: score only those, and only when the flag is up
};
constexpr uint32_t kDlLbMaxLen = 32000;  // 16-bit DP cells

__global__ void dl_init_kernel(int32_t* last_row, uint64_t count) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) last_row[i] = -1;
}

__global__ void __launch_bounds__(128) dl_kernel(const __grid_constant__ DlParams p) {
  extern __shared__ __align__(16) uint8_t dl_q[];
  for (uint32_t i = threadIdx.x; i < p.len1; i += blockDim.x) dl_q[i] = p.qbytes[i];
  __syncthreads();
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t T = p.T;
  const uint32_t rowlen = p.len1 + 2;
  int32_t* rows = p.scratch + t;
  int32_t* last = p.scratch + (size_t)3 * rowlen * T + t;
  const bool off64 = p.off64 != nullptr;
  if (p.only_long && *reinterpret_cast<const volatile unsigned long long*>(p.flag) == 0ull) return;
  for (uint64_t c = t; c < p.n; c += T) {
    const uint64_t o0 = off64 ? p.off64[c] : (uint64_t)p.off32[c];
    const uint64_t o1 = off64 ? p.off64[c + 1] : (uint64_t)p.off32[c + 1];
    const uint32_t len2 = (uint32_t)(o1 - o0);
    if (p.only_long && len2 <= kDlLbMaxLen) continue;
    const uint8_t* txt = p.chars + o0;
    const uint32_t raw = damerau_zhao([&](uint32_t i) -> uint32_t { return txt[i]; }, len2,
                                      [&](uint32_t j) -> uint32_t { return dl_q[j]; }, p.len1,
                                      [&](uint32_t k, uint32_t j) -> int32_t& { return rows[((size_t)k * rowlen + j) * T]; },
                                      [&](uint32_t ch) -> int32_t& { return last[(size_t)ch * T]; });
    if (p.out_f64) reinterpret_cast<double*>(p.out)[c] = finish_norm(p.epi, raw, p.len1, len2);
    else reinterpret_cast<uint32_t*>(p.out)[c] = finish_int(p.epi, raw, p.len1, len2);
  }
}

// Damerau-Levenshtein for queries of at most 64 elements over the interleaved layout: warp per group of 32 equal-length
// candidates (no divergence between the lanes' DP loops), the three DP rows and the last-row table of every thread in
// SHARED memory as 16-bit cells, thread-strided (cell e of thread t at [e * NT + t]: conflict-free).  Only symbols of
// the query are ever looked up in the last-row table, so it is indexed by the query's symbol ids (at most 64 + one
// dummy slot for every other symbol) instead of 256 entries.  ~70 KB per 128-thread CTA, 3 CTAs per SM.
constexpr int DL_NT = 128;
__global__ void __launch_bounds__(DL_NT) dl_lb_kernel(const __grid_constant__ DlParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t len1 = p.len1, rowlen = len1 + 2;
  int16_t* rows = reinterpret_cast<int16_t*>(smem_raw);              // [3 * rowlen][DL_NT]
  int16_t* last = rows + (size_t)3 * rowlen * DL_NT;                 // [len1 + 1][DL_NT]
  uint8_t* qmap = reinterpret_cast<uint8_t*>(last + (size_t)(len1 + 1) * DL_NT);  // [256] symbol -> id (0 = not in the query)
  uint8_t* q = qmap + 256;                                           // [len1]
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  for (uint32_t i = tid; i < 256; i += DL_NT) qmap[i] = 0;
  for (uint32_t i = tid; i < len1; i += DL_NT) q[i] = p.qbytes[i];
  __syncthreads();
  if (tid == 0) {  // ids in order of first occurrence
    uint32_t d = 0;
    for (uint32_t i = 0; i < len1; ++i)
      if (!qmap[q[i]]) qmap[q[i]] = (uint8_t)++d;
  }
  for (uint32_t e = 0; e <= len1; ++e) last[(size_t)e * DL_NT + tid] = -1;
  __syncthreads();
  int16_t* my_rows = rows + tid;
  int16_t* my_last = last + tid;
  const uint2* __restrict__ gdata = reinterpret_cast<const uint2*>(p.lb.gdata);
  const uint64_t nwarps = (uint64_t)gridDim.x * (DL_NT / 32), w0 = (uint64_t)blockIdx.x * (DL_NT / 32) + (tid >> 5);
  for (uint64_t g = w0; g < p.lb.ngroups; g += nwarps) {
    const uint32_t len2 = __ldg(p.lb.lens + g * 32 + lane);
    const uint32_t idx = __ldg(p.lb.perm + g * 32 + lane);
    if (idx == 0xFFFFFFFFu) continue;
    if (len2 > kDlLbMaxLen) { *p.flag = 1ull; continue; }
    const uint8_t* col = reinterpret_cast<const uint8_t*>(gdata + __ldg(p.lb.goff + g) * 32 + lane);
    const uint32_t raw = damerau_zhao([&](uint32_t i) -> uint32_t { return col[(size_t)(i >> 3) * 256 + (i & 7u)]; }, len2,
                                      [&](uint32_t j) -> uint32_t { return q[j]; }, len1,
                                      [&](uint32_t k, uint32_t j) -> int16_t& { return my_rows[(k * rowlen + j) * DL_NT]; },
                                      [&](uint32_t ch) -> int16_t& { return my_last[(uint32_t)qmap[ch] * DL_NT]; });
    if (p.out_f64) reinterpret_cast<double*>(p.out)[idx] = finish_norm(p.epi, raw, len1, len2);
    else reinterpret_cast<uint32_t*>(p.out)[idx] = finish_int(p.epi, raw, len1, len2);
  }
}

cudaError_t launch_dl(const ScanLaunch& L) {
  if (L.query.len1 > kDpMaxQuery) return cudaErrorNotSupported;
  const bool use_lb = L.lb.gdata != nullptr && L.query.len1 <= 64 && L.lb_flag != nullptr;
  if (use_lb) {
    DlParams q{};
    q.qbytes = L.query.qbytes;
    q.len1 = L.query.len1;
    q.out = L.out;
    q.out_f64 = L.out_is_f64;
    q.epi = L.epi;
    q.lb = L.lb;
    q.flag = L.lb_flag;
    const size_t smem = ((size_t)3 * (q.len1 + 2) + q.len1 + 1) * DL_NT * sizeof(int16_t) + 256 + q.len1 + 16;
    cudaError_t e = cudaFuncSetAttribute(dl_lb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(q.flag, 0, sizeof(unsigned long long), L.stream);
    if (e != cudaSuccess) return e;
    int ctas_per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, dl_lb_kernel, DL_NT, smem);
    if (e != cudaSuccess) return e;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    uint64_t grid = (uint64_t)L.sm_count * ctas_per_sm;
    const uint64_t need = (L.lb.ngroups + DL_NT / 32 - 1) / (DL_NT / 32);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    dl_lb_kernel<<<(uint32_t)grid, DL_NT, smem, L.stream>>>(q);
    g_launches.fetch_add(1);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  DlParams p{};
  p.only_long = use_lb ? 1 : 0;
  p.flag = L.lb_flag;
  p.chars = L.corpus.chars;
  p.off32 = L.corpus.off32;
  p.off64 = L.corpus.off64;
  p.n = L.corpus.n;
  p.qbytes = L.query.qbytes;
  p.len1 = L.query.len1;
  p.out = L.out;
  p.out_f64 = L.out_is_f64;
  p.epi = L.epi;
  const size_t entries = (size_t)3 * (p.len1 + 2) + 256;
  // (only_long: the rare leftovers of dl_lb_kernel -- a small grid)
  const uint64_t blocks = dp_blocks(p.n, entries * sizeof(int32_t), (uint64_t)L.sm_count * (p.only_long ? 1 : 4));
  p.T = (uint32_t)blocks * 128u;
  cudaError_t e = cudaSuccess;
  if (p.len1 + 16 > 48 * 1024) e = cudaFuncSetAttribute(dl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(p.len1 + 16));
  if (e != cudaSuccess) return e;
  e = dev_alloc(&p.scratch, entries * p.T * sizeof(int32_t), L.stream);
  if (e != cudaSuccess) return e;
  dl_init_kernel<<<(uint32_t)blocks, 256, 0, L.stream>>>(p.scratch + (size_t)3 * (p.len1 + 2) * p.T, (uint64_t)256 * p.T);
  dl_kernel<<<(uint32_t)blocks, 128, p.len1 + 16, L.stream>>>(p);
  g_launches.fetch_add(2);
  e = cudaGetLastError();
  dev_free(p.scratch, L.stream);
  return e;
}

// ------------------------------------------------------------------------------------------------ jaro mw
template <int MAXQ, class T = uint8_t>
__global__ void __launch_bounds__(128) jaro_mw_kernel(const __grid_constant__ MwParams p) {
  const bool off64 = p.off64 != nullptr;
  const uint64_t* __restrict__ pm = p.pm;
  const uint32_t words = p.words;
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < p.n; c += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t o0 = off64 ? p.off64[c] : (uint64_t)p.off32[c];
    const uint64_t o1 = off64 ? p.off64[c + 1] : (uint64_t)p.off32[c + 1];
    const uint32_t len2 = (uint32_t)(o1 - o0);
    const T* __restrict__ txt = reinterpret_cast<const T*>(p.chars) + o0;
    auto pmw = [&](uint32_t w, uint32_t ch) -> uint64_t { return __ldg(pm + (uint64_t)ch * words + w); };
    auto bytes = [&](uint32_t j) -> uint32_t { return txt[j]; };
    const uint32_t len1 = p.len1;
    auto jaro = [&](double cut) { return jaro_similarity_generic<MAXQ, T>(pmw, bytes, len1, len2, cut); };
    double r;
    if (p.epi.metric == M_JARO) {
      r = finish_float(p.epi, jaro);
    } else {
      uint32_t prefix = 0;
      while (prefix < 4 && prefix < len1 && prefix < len2 && ((pmw(0u, bytes(prefix)) >> prefix) & 1u)) ++prefix;
      const double pw = p.epi.prefix_weight;
      auto jw = [&](double cut) { return jaro_winkler_from(jaro, prefix, pw, cut); };
      r = finish_float(p.epi, jw);
    }
    reinterpret_cast<double*>(p.out)[c] = r;
  }
}

// Jaro / Jaro-Winkler with a query beyond 2048 elements (jaro.rs:286-337, :370-420 have no cap): one WARP per
// candidate.  The pattern flags P (one bit per query element) and the matched text characters (at most len1 of them,
// see jaro_similarity_generic) live in a per-warp scratch line; per text character the lanes search the window's
// blocks 32 at a time for the first free match (ballot), the transposition pass ranks the flagged positions with a
// warp prefix sum over the blocks' popcounts.
struct JaroLongParams {
  const uint8_t* chars;
  const uint32_t* off32;
  const uint64_t* off64;
  uint64_t n;
  const uint64_t* pm;  // [256][words]
  uint32_t len1;
  uint32_t words;
  uint8_t* scratch;    // [warps][stride]: words u64 flags, then len1 matched bytes
  uint64_t stride;
  void* out;
  Epi epi;
};

template <class T>
__global__ void __launch_bounds__(128) jaro_long_kernel(const __grid_constant__ JaroLongParams p) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warp_global = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint64_t total_warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  const bool off64 = p.off64 != nullptr;
  const uint64_t* __restrict__ pm = p.pm;
  const uint32_t words = p.words;
  unsigned long long* P = reinterpret_cast<unsigned long long*>(p.scratch + warp_global * p.stride);
  T* matched = reinterpret_cast<T*>(P + words);
  for (uint64_t c = warp_global; c < p.n; c += total_warps) {
    const uint64_t o0 = off64 ? p.off64[c] : (uint64_t)p.off32[c];
    const uint64_t o1 = off64 ? p.off64[c + 1] : (uint64_t)p.off32[c + 1];
    const uint32_t len2_orig = (uint32_t)(o1 - o0), len1_orig = p.len1;
    const T* __restrict__ txt = reinterpret_cast<const T*>(p.chars) + o0;
    auto jaro = [&](double cutoff) -> double {  // warp-uniform: every lane returns the same value
      if (cutoff > 1.0) return 0.0;
      if (len1_orig == 0 && len2_orig == 0) return 1.0;
      if (!jaro_length_filter(len1_orig, len2_orig, cutoff)) return 0.0;
      if (len1_orig == 1 && len2_orig == 1) return (__ldg(pm + (uint64_t)txt[0] * words) & 1u) ? 1.0 : 0.0;
      uint32_t len1 = len1_orig, len2 = len2_orig, bound;
      jaro_bounds(len1, len2, bound);
      for (uint32_t w = lane; w < words; w += 32) __stcg(P + w, 0ull);
      __syncwarp();
      uint32_t cc = 0;
      for (uint32_t j0 = 0; j0 < len2; j0 += 32) {
        const uint32_t my_ch = (j0 + lane < len2) ? (uint32_t)txt[j0 + lane] : 0u;  // 32 text characters per load
        const uint32_t jn = len2 - j0 < 32 ? len2 - j0 : 32;
        for (uint32_t jj = 0; jj < jn; ++jj) {
          const uint32_t j = j0 + jj;
          const uint32_t ch = __shfl_sync(0xffffffffu, my_ch, jj);
          const uint32_t lo = j > bound ? j - bound : 0;
          uint32_t hi = j + bound;
          if (hi >= len1) hi = len1 - 1;
          if (lo > hi) continue;
          const uint32_t w0 = lo / 64, w1 = hi / 64;
          for (uint32_t wb = w0; wb <= w1; wb += 32) {
            const uint32_t w = wb + lane;
            unsigned long long m = 0;
            if (w <= w1 && w < words) {
              m = __ldg(pm + (uint64_t)ch * words + w) & ~__ldcg(P + w);
              if (w == w0) m &= ~0ull << (lo % 64);
              if (w == w1) m &= ~0ull >> (63 - (hi % 64));
            }
            const uint32_t ball = __ballot_sync(0xffffffffu, m != 0);
            if (ball) {
              if (lane == (uint32_t)__ffs(ball) - 1u) {
                __stcg(P + w, __ldcg(P + w) | (m & (0ull - m)));
                matched[cc] = (T)ch;
              }
              ++cc;
              __syncwarp();
              break;
            }
          }
        }
      }
      __syncwarp();
      if (!jaro_common_char_filter(len1_orig, len2_orig, cc, cutoff)) return 0.0;
      uint32_t tr = 0, running = 0;
      for (uint32_t wb = 0; wb < words; wb += 32) {
        const uint32_t w = wb + lane;
        unsigned long long pw = (w < words) ? __ldcg(P + w) : 0ull;
        const uint32_t cnt = (uint32_t)__popcll(pw);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= (uint32_t)d) incl += t;
        }
        uint32_t k = running + incl - cnt;
        while (pw) {
          const unsigned long long bit = pw & (0ull - pw);
          tr += (__ldg(pm + (uint64_t)__ldcg(matched + k) * words + w) & bit) == 0;
          ++k;
          pw ^= bit;
        }
        running += __shfl_sync(0xffffffffu, incl, 31);
      }
      for (uint32_t d = 16; d >= 1; d >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, d);
      return jaro_calculate_similarity(len1_orig, len2_orig, cc, tr);
    };
    double r;
    if (p.epi.metric == M_JARO) {
      r = finish_float(p.epi, jaro);
    } else {
      uint32_t prefix = 0;
      while (prefix < 4 && prefix < len1_orig && prefix < len2_orig &&
             ((__ldg(pm + (uint64_t)txt[prefix] * words) >> prefix) & 1u)) ++prefix;
      const double pw = p.epi.prefix_weight;
      auto jw = [&](double cut) { return jaro_winkler_from(jaro, prefix, pw, cut); };
      r = finish_float(p.epi, jw);
    }
    if (lane == 0) reinterpret_cast<double*>(p.out)[c] = r;
    __syncwarp();
  }
}

static cudaError_t launch_jaro_long(const ScanLaunch& L) {
  JaroLongParams p{};
  p.chars = L.corpus.chars;
  p.off32 = L.corpus.off32;
  p.off64 = L.corpus.off64;
  p.n = L.corpus.n;
  p.pm = L.query.pm_words;
  p.len1 = L.query.len1;
  p.words = L.query.words;
  p.out = L.out;
  p.epi = L.epi;
  p.stride = ((uint64_t)p.words * 8 + (uint64_t)p.len1 * (L.elem16 ? 2 : 1) + 63) / 64 * 64;
  uint64_t warps = (uint64_t)L.sm_count * 16;
  if (warps > p.n) warps = p.n;
  if (warps < 1) warps = 1;
  const uint32_t blocks = (uint32_t)((warps + 3) / 4);
  cudaError_t e = dev_alloc(&p.scratch, (uint64_t)blocks * 4 * p.stride, L.stream);
  if (e != cudaSuccess) return e;
  if (L.elem16) jaro_long_kernel<uint16_t><<<blocks, 128, 0, L.stream>>>(p);
  else jaro_long_kernel<uint8_t><<<blocks, 128, 0, L.stream>>>(p);
  g_launches.fetch_add(1);
  e = cudaGetLastError();
  dev_free(p.scratch, L.stream);
  return e;
}

cudaError_t launch_jaro_mw(const ScanLaunch& L) {
  if (L.query.len1 > 2048) return launch_jaro_long(L);
  MwParams p{};
  p.chars = L.corpus.chars;
  p.off32 = L.corpus.off32;
  p.off64 = L.corpus.off64;
  p.n = L.corpus.n;
  p.pm = L.query.pm_words;
  p.len1 = L.query.len1;
  p.words = L.query.words;
  p.out = L.out;
  p.out_f64 = 1;
  p.epi = L.epi;
  uint64_t blocks = (p.n + 127) / 128;
  const uint64_t max_blocks = (uint64_t)L.sm_count * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  if (L.elem16) jaro_mw_kernel<2048, uint16_t><<<(uint32_t)blocks, 128, 0, L.stream>>>(p);
  else if (p.len1 <= 256) jaro_mw_kernel<256><<<(uint32_t)blocks, 128, 0, L.stream>>>(p);
  else if (p.len1 <= 2048) jaro_mw_kernel<2048><<<(uint32_t)blocks, 128, 0, L.stream>>>(p);
  else return cudaErrorInvalidValue;
  g_launches.fetch_add(1);
  return cudaGetLastError();
}

}  // namespace rfk

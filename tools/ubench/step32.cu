// The 32-bit Hyyro step of scan_lb_kernel in isolation (dev tool): how many clocks per character and warp does the SMSP need
// when nothing but the step runs -- no row loads, no group prologue / epilogue -- and does a second independent chain per
// thread (two candidates per lane) change it?
//   REG   match masks from registers (no shared-memory gather)
//   LDS   IDP.4A address + LDS from the lane-replicated 32 KB table, like the kernel
//   ILP   1 or 2 independent (VP, VN) chains per thread
// CTAs of 256 threads, CPS CTAs per SM (6 = the kernel's 12 warps per SMSP).
// nvcc -arch=sm_100a -o step32 step32.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void step(uint32_t X, uint32_t& VP, uint32_t& VN, uint32_t two, uint32_t one) {
  const uint32_t D0 = ((((X & VP) + VP) ^ VP) | X) | VN;
  uint32_t HP = VN | ~(D0 | VP);
  uint32_t HN = D0 & VP;
  HP = HP * two + one;
  HN = HN * two;
  VP = HN | ~(D0 | HP);
  VN = HP & D0;
}

template <bool LDS, int ILP>
__global__ void __launch_bounds__(256) k(unsigned* out, unsigned a, unsigned two, int iters, long long* cyc) {
  extern __shared__ uint32_t pm[];
  for (uint32_t i = threadIdx.x; i < 256u * 32u; i += 256) pm[i] = (i >> 5) * 2654435761u ^ a;
  __syncthreads();
  const unsigned one = two >> 1;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(pm + lane);
  uint32_t W[4], W2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { W[i] = (threadIdx.x * 2654435761u) ^ (a * (i + 1) * 40503u); W2[i] = ~W[i] * 977u; }
  uint32_t VP = ~0u, VN = 0, VP2 = ~0u << 3, VN2 = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        uint32_t X, X2 = 0;
        if (LDS) {
          const uint32_t addr = __dp4a(W[j], 0x80u << (8 * b), base);
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(X) : "r"(addr));
          if (ILP == 2) {
            const uint32_t addr2 = __dp4a(W2[j], 0x80u << (8 * b), base);
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(X2) : "r"(addr2));
          }
        } else {
          X = W[j] >> b ^ W[(j + 1) & 3];
          if (ILP == 2) X2 = W2[j] >> b ^ W2[(j + 1) & 3];
        }
        step(X, VP, VN, two, one);
        if (ILP == 2) step(X2, VP2, VN2, two, one);
      }
      W[j] += VN & 0x01010101u;   // keep the text words data dependent (no hoisting), cheap
      if (ILP == 2) W2[j] += VN2 & 0x01010101u;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = VP ^ VN ^ VP2 ^ VN2;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <bool LDS, int ILP>
void run(const char* name, int cps, unsigned* out, long long* cyc, int sms) {
  const int iters = 8192;
  auto kern = k<LDS, ILP>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 32768);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    kern<<<sms * cps, 256, 32768>>>(out, 3, 2, iters, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
  }
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  static long long h[4096];
  cudaMemcpy(h, cyc, sizeof(long long) * sms * cps, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms * cps; ++i) avg += (double)h[i];
  avg /= sms * cps;
  const double warps_per_smsp = cps * 8 / 4.0;
  // clocks the SMSP spends per character of ONE warp-chain: elapsed / (chars per chain) / (chains per SMSP)
  // wall: the whole launch (CUDA events) in SM clocks at the device's reported clock rate, same normalisation
  printf("%-10s %d CTAs/SM (occ %d) %5.2f warps/SMSP: %6.2f clock64 ticks, %6.2f wall clocks @%d MHz per character and warp-chain\n", name, cps, occ,
         warps_per_smsp, avg / (iters * 16.0) / (warps_per_smsp * ILP), ms * 1e-3 * khz * 1e3 / (iters * 16.0) / (warps_per_smsp * ILP), khz / 1000);
}
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* out; long long* cyc;
  cudaMalloc(&out, sizeof(unsigned) * sms * 8 * 256);
  cudaMalloc(&cyc, sizeof(long long) * 4096);
  for (int cps : {6, 4, 3, 2}) {
    run<false, 1>("REG ILP1", cps, out, cyc, sms);
    run<false, 2>("REG ILP2", cps, out, cyc, sms);
    run<true, 1>("LDS ILP1", cps, out, cyc, sms);
    run<true, 2>("LDS ILP2", cps, out, cyc, sms);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}

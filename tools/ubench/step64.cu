// Variants of the 64-bit Hyyro step on 32-bit halves (dev tool): which mix of carry / shift forms is fastest on sm_100a?
// 12 warps per SMSP like scan_lb_kernel; match masks come from registers (no LDS), one dependent chain per thread.
//   V0  native uint64_t C++ (ptxas: IADD3 + IADD3.X, SHF funnel shifts)
//   V1  add.cc + madc (IMAD.X), mul.hi + IMAD for HP, mul.wide + IMAD for HN     (all shifts on the FMA pipe)
//   V2  add.cc + madc, low halves shifted by IMAD, high halves by one SHF.L.W funnel shift each (ALU)
//   V3  like V2 but HN's high half by mul.wide carry + IMAD (one SHF, one WIDE)
// nvcc -arch=sm_100a -o step64 step64.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
template <int V>
__global__ void __launch_bounds__(768) k(unsigned* out, unsigned a, unsigned two, int iters, long long* cyc) {
  const unsigned one = two >> 1;
  unsigned X[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) X[i] = (threadIdx.x * 2654435761u) ^ (a * (i + 1) * 40503u);
  unsigned VPl = ~0u, VPh = ~0u, VNl = 0, VNh = 0;
  uint64_t VP = ~0ull, VN = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const unsigned Xl = X[j], Xh = X[(j + 3) & 7];
      if (V == 0) {
        const uint64_t Xw = ((uint64_t)Xh << 32) | Xl;
        const uint64_t D0 = ((((Xw & VP) + VP) ^ VP) | Xw) | VN;
        uint64_t HP = VN | ~(D0 | VP);
        uint64_t HN = D0 & VP;
        HP = (HP << 1) | 1ull;
        HN = HN << 1;
        VP = HN | ~(D0 | HP);
        VN = HP & D0;
      } else {
        unsigned sl, sh, c;
        asm("{\n\tadd.cc.u32 %0, %2, %3;\n\tmadc.lo.u32 %1, %4, %5, %6;\n\t}"
            : "=r"(sl), "=r"(sh) : "r"(Xl & VPl), "r"(VPl), "r"(VPh), "r"(one), "r"(Xh & VPh));
        const unsigned D0l = ((sl ^ VPl) | Xl) | VNl;
        const unsigned D0h = ((sh ^ VPh) | Xh) | VNh;
        unsigned HPl = VNl | ~(D0l | VPl), HPh = VNh | ~(D0h | VPh);
        unsigned HNl = D0l & VPl, HNh = D0h & VPh;
        if (V == 1) {
          asm("mul.hi.u32 %0, %1, %2;" : "=r"(c) : "r"(HPl), "r"(two));
          HPl = HPl * two + one;
          HPh = HPh * two + c;
          asm("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(HNl), "=r"(c) : "r"(HNl), "r"(two));
          HNh = HNh * two + c;
        } else if (V == 2) {
          asm("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(HPh) : "r"(HPl));
          HPl = HPl * two + one;
          asm("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(HNh) : "r"(HNl));
          HNl = HNl * two;
        } else {
          asm("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(HPh) : "r"(HPl));
          HPl = HPl * two + one;
          asm("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(HNl), "=r"(c) : "r"(HNl), "r"(two));
          HNh = HNh * two + c;
        }
        VPl = HNl | ~(D0l | HPl);
        VPh = HNh | ~(D0h | HPh);
        VNl = HPl & D0l;
        VNh = HPh & D0h;
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = VPl ^ VPh ^ VNl ^ VNh ^ (unsigned)VP ^ (unsigned)(VP >> 32) ^ (unsigned)VN ^ (unsigned)(VN >> 32);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int V>
void run(const char* name, unsigned* out, long long* cyc, int sms) {
  const int iters = 2048;
  for (int rep = 0; rep < 2; ++rep) { k<V><<<sms * 2, 768>>>(out, 3, 2, iters, cyc); cudaDeviceSynchronize(); }
  long long h[512];
  cudaMemcpy(h, cyc, sizeof(long long) * sms * 2, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms * 2; ++i) avg += (double)h[i];
  avg /= sms * 2;
  printf("%-8s %.2f clk per step per SMSP-warp (12 warps/SMSP; the ALU-pipe floor of 15-17 ALU ops is 30-34)\n", name, avg / (iters * 8.0) / 12.0);
}
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* out; long long* cyc;
  cudaMalloc(&out, sizeof(unsigned) * sms * 2 * 768);
  cudaMalloc(&cyc, sizeof(long long) * 512);
  run<0>("V0", out, cyc, sms);
  run<1>("V1", out, cyc, sms);
  run<2>("V2", out, cyc, sms);
  run<3>("V3", out, cyc, sms);
  return 0;
}

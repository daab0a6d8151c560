// Pipe-throughput microbenchmark 2 (dev tool): carry / wide / funnel-shift forms used by the 64-bit Levenshtein step.
// Every mode's loop body is checked in SASS (cuobjdump) before its number is trusted: INSTR[] below is the count of
// the named instruction per loop iteration as compiled.  nvcc -arch=sm_100a -o pipes2 pipes2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
template <int MODE>
__global__ void __launch_bounds__(1024) k(unsigned* out, unsigned a, unsigned b, int iters, long long* cyc) {
  unsigned r[8], s[8], t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { r[i] = threadIdx.x * 31u + i + a; s[i] = threadIdx.x * 17u + i + b; t[i] = threadIdx.x * 7u + i * a; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#define LOP(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b));
#define MAD(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(a), "r"(b));
#define SHF(i) asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(r[i]) : "r"(t[i]));
#define WID(i) asm volatile("{\n\t.reg .b64 w;\n\tmul.wide.u32 w, %0, %2;\n\tmov.b64 {%0, %1}, w;\n\t}" : "+r"(s[i]), "=r"(t[i]) : "r"(a));
#define MHI(i) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(a));
#define CAR(i) asm volatile("{\n\tadd.cc.u32 %0, %0, %2;\n\tmadc.lo.u32 %1, %1, %3, %2;\n\t}" : "+r"(s[i]), "+r"(t[i]) : "r"(a), "r"(b));
#define AD3(i) asm volatile("{\n\t.reg .u32 x;\n\tadd.u32 x, %0, %1;\n\txor.b32 %0, x, %2;\n\t}" : "+r"(s[i]) : "r"(t[i]), "r"(a));
#define ADC(i) asm volatile("{\n\tadd.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;\n\t}" : "+r"(s[i]), "+r"(t[i]) : "r"(a), "r"(b));
    if (MODE == 0) { REP8(LOP) REP8(LOP) }
    if (MODE == 1) { REP8(SHF) REP8(SHF) }
    if (MODE == 2) { REP8(WID) REP8(WID) }
    if (MODE == 3) { REP8(MHI) REP8(MHI) }
    if (MODE == 4) { REP8(CAR) REP8(CAR) }
    if (MODE == 5) { REP8(ADC) REP8(ADC) }
    if (MODE == 6) { REP8(AD3) REP8(AD3) }
    if (MODE == 7) { REP8(LOP) REP8(SHF) }
    if (MODE == 8) { REP8(LOP) REP8(WID) }
    if (MODE == 9) { REP8(LOP) REP8(MHI) }
    if (MODE == 10) { REP8(LOP) REP8(CAR) }
    if (MODE == 11) { REP8(LOP) REP8(ADC) }
    if (MODE == 12) { REP8(MAD) REP8(ADC) }
    if (MODE == 13) { REP8(MAD) REP8(WID) }
  }
  long long t1 = clock64();
  unsigned x = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) x ^= r[i] ^ s[i] ^ t[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, unsigned* out, long long* cyc, int sms) {
  const int iters = 4096, threads = 1024;
  k<MODE><<<sms, threads>>>(out, 3, 5, iters, cyc);
  cudaDeviceSynchronize();
  k<MODE><<<sms, threads>>>(out, 3, 5, iters, cyc);
  cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; ++i) avg += (double)h[i];
  avg /= sms;
  // cycles per loop iteration per SMSP (8 warps per SMSP): divide by the SASS instruction count of the body
  printf("%-14s %.2f clk per iteration per warp-slot (8 warps/SMSP => %.2f clk per warp-iteration)\n", name, avg / iters, avg / iters / 8);
}
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* out; long long* cyc;
  cudaMalloc(&out, sizeof(unsigned) * sms * 1024);
  cudaMalloc(&cyc, sizeof(long long) * 256);
  run<0>("16 LOP3", out, cyc, sms);
  run<1>("16 SHF", out, cyc, sms);
  run<2>("16 WIDE", out, cyc, sms);
  run<3>("16 MULHI", out, cyc, sms);
  run<4>("16 ADDCC+MADC", out, cyc, sms);
  run<5>("16 ADDCC+ADDC", out, cyc, sms);
  run<6>("16 ADD+XOR", out, cyc, sms);
  run<7>("8LOP+8SHF", out, cyc, sms);
  run<8>("8LOP+8WIDE", out, cyc, sms);
  run<9>("8LOP+8MULHI", out, cyc, sms);
  run<10>("8LOP+8(CC+MADC)", out, cyc, sms);
  run<11>("8LOP+8(CC+ADDC)", out, cyc, sms);
  run<12>("8MAD+8(CC+ADDC)", out, cyc, sms);
  run<13>("8MAD+8WIDE", out, cyc, sms);
  return 0;
}

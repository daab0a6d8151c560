// Pipe-throughput microbenchmark (dev tool): which issue pipe do LOP3 / IMAD / IDP.4A / PRMT use on sm_100a and
// do they overlap?  Prints warp-instructions per clock per SM for each mix.  nvcc -arch=sm_100a -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
template <int MODE>
__global__ void __launch_bounds__(1024) k(unsigned* out, unsigned a, unsigned b, int iters, long long* cyc) {
  unsigned r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = threadIdx.x * 31u + i + a;
  unsigned s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = threadIdx.x * 17u + i + b;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#define LOP(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b));
#define MAD(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(a), "r"(b));
#define IDP(i) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(a), "r"(b));
#define PRM(i) asm volatile("prmt.b32 %0, %0, %1, 0x4442;" : "+r"(r[i]) : "r"(a));
#define ADD(i) asm volatile("add.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(a));
    if (MODE == 0) { REP8(LOP) REP8(LOP) }
    if (MODE == 1) { REP8(MAD) REP8(MAD) }
    if (MODE == 2) { REP8(IDP) REP8(IDP) }
    if (MODE == 3) { REP8(PRM) REP8(PRM) }
    if (MODE == 4) { REP8(LOP) REP8(MAD) }
    if (MODE == 5) { REP8(LOP) REP8(IDP) }
    if (MODE == 6) { REP8(LOP) REP8(PRM) }
    if (MODE == 7) { REP8(MAD) REP8(IDP) }
    if (MODE == 8) { REP8(ADD) REP8(ADD) }
    if (MODE == 9) { REP8(LOP) REP8(ADD) }
  }
  long long t1 = clock64();
  unsigned x = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) x ^= r[i] ^ s[i];
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
  const double winst = (double)iters * 16 * (threads / 32);
  printf("%-12s %.3f warp-inst/clk/SM  (%.2f per SMSP)\n", name, winst / avg, winst / avg / 4);
}
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* out; long long* cyc;
  cudaMalloc(&out, sizeof(unsigned) * sms * 1024);
  cudaMalloc(&cyc, sizeof(long long) * 256);
  run<0>("LOP3", out, cyc, sms);
  run<1>("IMAD", out, cyc, sms);
  run<2>("IDP4A", out, cyc, sms);
  run<3>("PRMT", out, cyc, sms);
  run<8>("IADD", out, cyc, sms);
  run<4>("LOP3+IMAD", out, cyc, sms);
  run<5>("LOP3+IDP4A", out, cyc, sms);
  run<6>("LOP3+PRMT", out, cyc, sms);
  run<7>("IMAD+IDP4A", out, cyc, sms);
  run<9>("LOP3+IADD", out, cyc, sms);
  return 0;
}

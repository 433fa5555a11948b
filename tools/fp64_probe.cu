// fp64_probe.cu -- what the FP64 pipe of this GPU really delivers (the ceiling the stepping kernels are
// judged against next to HBM): throughput of DFMA / DADD / DMUL streams at several ILP x occupancy points and
// the dependent-issue latency of one chain.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false tools/fp64_probe.cu -o build/fp64_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int ILP, int OP>
__global__ void Stream(double* out, double a, double b, int iters) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = a + k + threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      if (OP == 0) x[k] = __fma_rn(x[k], b, a);
      if (OP == 1) x[k] = __dadd_rn(x[k], a);
      if (OP == 2) x[k] = __dmul_rn(x[k], b);
      if (OP == 3) { x[k] = __dmul_rn(x[k], b); x[k] = __dadd_rn(x[k], a); }
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  if (s == 12345.678) out[0] = s;
}

__global__ void Latency(double* out, long long* cyc, double a, double b, int iters) {
  double x = a;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    x = __dadd_rn(x, a);
    x = __dmul_rn(x, b);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; out[0] = x; }
}

template <int ILP, int OP>
void Run(const char* name, int blocksPerSM, int threads, int sms, double clkGHz) {
  double* out;
  cudaMalloc(&out, 8);
  const int iters = 20000;
  const int grid = blocksPerSM * sms;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  Stream<ILP, OP><<<grid, threads>>>(out, 1.0000001, 0.9999999, 100);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  Stream<ILP, OP><<<grid, threads>>>(out, 1.0000001, 0.9999999, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double insts = double(grid) * threads * iters * ILP * (OP == 3 ? 2 : 1);
  const double perSmClk = insts / (ms * 1e-3) / sms / (clkGHz * 1e9);
  printf("%-6s ILP %d  warps/SM %3d : %7.2f T inst/s  %6.1f lane-inst/clk/SM (at %.3f GHz)\n", name, ILP,
         blocksPerSM * threads / 32, insts / (ms * 1e-3) / 1e12, perSmClk, clkGHz);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  printf("%s: %d SMs, %.3f GHz\n", p.name, sms, ghz);
  Run<8, 0>("DFMA", 8, 256, sms, ghz);
  Run<8, 1>("DADD", 8, 256, sms, ghz);
  Run<8, 2>("DMUL", 8, 256, sms, ghz);
  Run<8, 3>("MULADD", 8, 256, sms, ghz);
  // one dependent chain per thread: how many warps per SM does it take to fill the pipe?
  Run<1, 3>("MULADD", 1, 128, sms, ghz);
  Run<1, 3>("MULADD", 1, 256, sms, ghz);
  Run<1, 3>("MULADD", 2, 256, sms, ghz);
  Run<1, 3>("MULADD", 3, 256, sms, ghz);
  Run<1, 3>("MULADD", 4, 256, sms, ghz);
  Run<1, 3>("MULADD", 8, 256, sms, ghz);
  Run<2, 3>("MULADD", 2, 256, sms, ghz);
  Run<4, 3>("MULADD", 2, 256, sms, ghz);
  double* out;
  long long* cyc;
  cudaMalloc(&out, 8);
  cudaMalloc(&cyc, 8);
  Latency<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, 10000);
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dependent DADD->DMUL chain: %.2f cycles per instruction\n", double(h) / 20000.0);
  return 0;
}

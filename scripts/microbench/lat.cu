// Dependent-issue latencies of the operations the PEAQ kernels chain (sm_100a): cycles per
// operation of ONE warp running a dependent chain (clock64 around 4096 operations), and the same
// with 2 / 4 / 8 independent chains per thread (ILP).  Development aid; numbers go to DESIGN.md.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int ILP>
__global__ void chain(double* out, long long* cyc, double a, double b, int iters) {
  __shared__ double sm[32 * 8];
  double x[ILP];
  for (int k = 0; k < ILP; k++) x[k] = a + threadIdx.x + k;
  for (int k = 0; k < 8; k++) sm[threadIdx.x * 8 + k] = k;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) {
      if (OP == 0) x[k] = fma(x[k], b, a);
      if (OP == 1) x[k] = x[k] + b;
      if (OP == 2) x[k] = x[k] * b;
      if (OP == 3) x[k] = __shfl_down_sync(0xffffffffu, x[k], 1);
      if (OP == 4) { int idx = ((int)x[k]) & 7; x[k] = sm[threadIdx.x * 8 + idx]; }   // LDS.64, address depends on the value
      if (OP == 5) x[k] = sqrt(x[k]);
      if (OP == 6) x[k] = b / x[k];
      if (OP == 7) { x[k] = x[k] + b; x[k] = __shfl_down_sync(0xffffffffu, x[k], 1); }
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < ILP; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP, int ILP>
void run(const char* name, double b) {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * sizeof(double));
  cudaMalloc(&cyc, 8 * sizeof(long long));
  const int iters = 4096;
  chain<OP, ILP><<<1, 32>>>(out, cyc, 1.0, b, iters);
  chain<OP, ILP><<<1, 32>>>(out, cyc, 1.0, b, iters);
  long long h = 0;
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  printf("%-10s ILP %d: %.2f cycles per op-slot, %.2f per operation\n", name, ILP, (double)h / iters, (double)h / iters / ILP);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0, 1>("DFMA", 0.999999); run<0, 2>("DFMA", 0.999999); run<0, 4>("DFMA", 0.999999); run<0, 8>("DFMA", 0.999999);
  run<1, 1>("DADD", 1e-9); run<1, 4>("DADD", 1e-9);
  run<2, 1>("DMUL", 0.999999); run<2, 4>("DMUL", 0.999999);
  run<3, 1>("SHFL64", 0); run<3, 4>("SHFL64", 0);
  run<4, 1>("LDS64", 0); run<4, 4>("LDS64", 0);
  run<5, 1>("DSQRT", 0); run<5, 4>("DSQRT", 0);
  run<6, 1>("DDIV", 1.000001); run<6, 4>("DDIV", 1.000001);
  run<7, 1>("DADD+SHFL", 1e-9); run<7, 4>("DADD+SHFL", 1e-9);
  return 0;
}

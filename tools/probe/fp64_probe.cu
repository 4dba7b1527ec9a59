// Development probe: FP64 latency / throughput / division cost on the SM (calibrates the k_icp_loop cost model).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_probe fp64_probe.cu && ./fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dep_add(double* out, long long* cyc, double x, int n) {
  double a = x;
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) a = a + x;  // dependent chain
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
__global__ void k_thr_fma(double* out, long long* cyc, double x, int n) {
  double a[ILP];
#pragma unroll
  for (int u = 0; u < ILP; ++u) a[u] = x + u;
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int u = 0; u < ILP; ++u) a[u] = a[u] * x + 1.0;  // -fmad=false: DMUL + DADD
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int u = 0; u < ILP; ++u) s += a[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_div(double* out, long long* cyc, double num, double den, int n, int zero_lane) {
  double a = (threadIdx.x & 31) == zero_lane ? 0.0 : num;
  double acc = 0;
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    acc += a / den;
    den += 1e-9;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_sqrt(double* out, long long* cyc, double x, int n) {
  double a = x, acc = 0;
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    acc += sqrt(a);
    a += 1e-3;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out;
  long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 8);
  const int n = 4096;
  auto rep = [&](const char* name, double per) {
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-60s %8.2f cycles per %s\n", name, (double)h / n, per ? "iteration" : "iteration");
  };
  k_dep_add<<<1, 32>>>(out, cyc, 1.0000001, n);
  rep("dependent DADD, 1 warp", 1);
  k_dep_add<<<148, 512>>>(out, cyc, 1.0000001, n);
  rep("dependent DADD, 16 warps/SM (4/SMSP)", 1);
  k_thr_fma<1><<<1, 32>>>(out, cyc, 1.0000001, n);
  rep("DMUL+DADD dependent pair, 1 warp", 1);
  k_thr_fma<8><<<148, 512>>>(out, cyc, 1.0000001, n);
  rep("8 x (DMUL+DADD) independent, 16 warps/SM: /16 = per instr", 1);
  k_thr_fma<8><<<148, 1024>>>(out, cyc, 1.0000001, n);
  rep("8 x (DMUL+DADD) independent, 32 warps/SM: /16 = per instr", 1);
  k_div<<<1, 32>>>(out, cyc, 3.0, 7.0, n, -1);
  rep("DDIV (+DADD x2), 1 warp, normal operands", 1);
  k_div<<<1, 32>>>(out, cyc, 3.0, 7.0, n, 5);
  rep("DDIV (+DADD x2), 1 warp, one lane with a ZERO numerator", 1);
  k_div<<<148, 512>>>(out, cyc, 3.0, 7.0, n, -1);
  rep("DDIV, 16 warps/SM, normal operands", 1);
  k_div<<<148, 512>>>(out, cyc, 3.0, 7.0, n, 5);
  rep("DDIV, 16 warps/SM, one lane zero numerator", 1);
  k_sqrt<<<1, 32>>>(out, cyc, 2.0, n);
  rep("DSQRT (+2 DADD), 1 warp", 1);
  k_sqrt<<<148, 512>>>(out, cyc, 2.0, n);
  rep("DSQRT (+2 DADD), 16 warps/SM", 1);
  return 0;
}

// Dependent-issue latencies on sm_100a (cycles): DMMA.8x8x4 chain, DFMA chain, LDS.64 chain, bar.sync of 4/8 warps.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void lat(long long* out, double* sink, int reps) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 8 + 8) % 1024);  // pointer chase table (byte offsets/8)
  __syncthreads();
  double c0 = 0, c1 = 0, a = 1e-3 * threadIdx.x, b = 1.0;
  long long t0, t1;
  if (threadIdx.x < 32) {
    t0 = clock64();
    for (int i = 0; i < reps; i++) dmma(c0, c1, a, b);
    t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / reps;
    double x = a;
    t0 = clock64();
    for (int i = 0; i < reps; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(b), "d"(a));
    t1 = clock64();
    if (threadIdx.x == 0) out[1] = (t1 - t0) / reps;
    int idx = threadIdx.x;
    t0 = clock64();
    for (int i = 0; i < reps; i++) idx = (int)sm[idx & 1023] & 1023;
    t1 = clock64();
    if (threadIdx.x == 0) out[2] = (t1 - t0) / reps;
    sink[threadIdx.x] = c0 + c1 + x + idx;
    // two independent DMMA chains
    double d0 = 0, d1 = 0;
    t0 = clock64();
    for (int i = 0; i < reps; i++) { dmma(c0, c1, a, b); dmma(d0, d1, a, b); }
    t1 = clock64();
    if (threadIdx.x == 0) out[3] = (t1 - t0) / reps;
    sink[threadIdx.x] += c0 + d0 + d1;
  }
  __syncthreads();
  t0 = clock64();
  for (int i = 0; i < reps; i++) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) out[4] = (t1 - t0) / reps;
  // shuffle latency
  double s = a;
  t0 = clock64();
  for (int i = 0; i < reps; i++) s = __shfl_xor_sync(0xffffffffu, s, 1);
  t1 = clock64();
  if (threadIdx.x == 0) out[5] = (t1 - t0) / reps;
  sink[threadIdx.x] += s;
}
int main() {
  long long* out; double* sink;
  cudaMalloc(&out, 64); cudaMalloc(&sink, 8 * 1024);
  for (int nth : {128, 256}) {
    lat<<<1, nth>>>(out, sink, 1000);
    lat<<<1, nth>>>(out, sink, 1000);
    cudaDeviceSynchronize();
    long long h[6]; cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    printf("{\"threads\": %d, \"dmma_dep\": %lld, \"dfma_dep\": %lld, \"lds64_cvt_dep\": %lld, \"dmma_2chains\": %lld, \"syncthreads\": %lld, \"shfl64\": %lld}\n", nth, h[0], h[1], h[2], h[3], h[4], h[5]);
  }
  return 0;
}

// FP64 roofline denominator probe for B200 (sm_100a): DFMA register loop, DMMA m8n8k4 loop,
// and a DFMA loop fed from shared memory (LDS.128 broadcast + distinct), to decide the K1 layout.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA m8n8k4: A 8x4 (1 double/thread), B 4x8 (1 double/thread), C 8x8 (2 doubles/thread)
template<int ILP>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x * 1e-3; c1[i] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA m16n8k8 (lowered to 4x DMMA.8x8x4 by ptxas per survey probe)
template<int ILP>
__global__ void __launch_bounds__(256) dmma16_kernel(double* out, int iters, double a, double b) {
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DFMA fed from smem: per k-step 1 distinct LDS.128 (own row) + NB broadcast LDS.128, 4*NB DFMA (complex 1xNB tile)
template<int NB>
__global__ void __launch_bounds__(128) dfma_lds_kernel(double* out, int iters) {
  __shared__ double2 sa[32 * 32];
  __shared__ double2 sb[32 * 32];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) { sa[i] = make_double2(1e-3 * i, 1e-4); sb[i] = make_double2(1e-5 * i, 1.0); }
  __syncthreads();
  double cr[NB], ci[NB];
#pragma unroll
  for (int j = 0; j < NB; j++) { cr[j] = 0; ci[j] = 0; }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int it = 0; it < iters; it++) {
#pragma unroll 4
    for (int k = 0; k < 32; k++) {
      double2 a = sa[k * 32 + lane];
#pragma unroll
      for (int j = 0; j < NB; j++) {
        double2 b = sb[k * 32 + ((w * NB + j) & 31)];
        cr[j] = fma(a.x, b.x, cr[j]); cr[j] = fma(-a.y, b.y, cr[j]);
        ci[j] = fma(a.x, b.y, ci[j]); ci[j] = fma(a.y, b.x, ci[j]);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < NB; j++) s += cr[j] + ci[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<typename F> float timeit(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, clk);
  int nsm = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 16 * 256));
  const int iters = 20000;
  for (int bps : {1, 2, 4, 8}) {
    int grid = nsm * bps;
    { float ms = timeit([&]{ dfma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 8 * iters * 256.0 * grid; printf("{\"test\": \"dfma_ilp8\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", bps, ms, fl / ms * 1e-9); }
    { float ms = timeit([&]{ dmma_kernel<4><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 256 * 4 * iters * 8.0 * grid; printf("{\"test\": \"dmma_m8n8k4_ilp4\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", bps, ms, fl / ms * 1e-9); }
    { float ms = timeit([&]{ dmma16_kernel<4><<<grid, 256>>>(out, iters / 4, 1.0000001, 1e-9); });
      double fl = 2.0 * 1024 * 4 * (iters / 4) * 8.0 * grid; printf("{\"test\": \"dmma_m16n8k8_ilp4\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", bps, ms, fl / ms * 1e-9); }
  }
  for (int bps : {1, 2, 4}) {
    int grid = nsm * bps;
    { float ms = timeit([&]{ dfma_lds_kernel<4><<<grid, 128>>>(out, 200); });
      double fl = 2.0 * 4 * 4 * 32 * 200 * 128.0 * grid; printf("{\"test\": \"dfma_lds_1x4\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", bps, ms, fl / ms * 1e-9); }
    { float ms = timeit([&]{ dfma_lds_kernel<7><<<grid, 128>>>(out, 200); });
      double fl = 2.0 * 4 * 7 * 32 * 200 * 128.0 * grid; printf("{\"test\": \"dfma_lds_1x7\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", bps, ms, fl / ms * 1e-9); }
  }
  // sustained: 3 s of DFMA to see the clock/power-capped figure
  { int grid = nsm * 8; float total = 0; int n = 0; double fl = 2.0 * 8 * iters * 256.0 * grid;
    while (total < 3000.f) { total += timeit([&]{ dfma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 1) * 3; n++; }
    float ms = timeit([&]{ dfma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 3);
    printf("{\"test\": \"dfma_sustained_after_3s\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, fl / ms * 1e-9); }
  { int grid = nsm * 8; float total = 0; double fl = 2.0 * 256 * 4 * iters * 8.0 * grid;
    while (total < 3000.f) { total += timeit([&]{ dmma_kernel<4><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 1) * 3; }
    float ms = timeit([&]{ dmma_kernel<4><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 3);
    printf("{\"test\": \"dmma_sustained_after_3s\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, fl / ms * 1e-9); }
  return 0;
}

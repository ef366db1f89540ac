// Micro-benchmark of K1's service-warp inverse (service_inverse<Cfg<4,28,7>>, d = 27) alone and next to DMMA-saturated
// "compute" warps (2 per SM sub-partition, as in K1): cycles per inverse.
#include <cstdio>
#include "../quantumoptimalcontrol.jl_b200/csrc/qoc_k1.cuh"
using namespace qoc;
typedef Cfg<4, 28, 7> C;

template <int NCOMP>
__global__ void __launch_bounds__((NSW + NCOMP) * 32, 1) bench(int d, int reps, long long* out, double* sink) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* base = reinterpret_cast<double*>(smem_raw);
  const int slot_d = 2 * d * C::S;
  SvcScratch* sc = reinterpret_cast<SvcScratch*>(base + 2 * slot_d + 8 * C::S);
  __shared__ int stop;
  for (int e = threadIdx.x; e < 2 * slot_d + 8 * C::S; e += blockDim.x) base[e] = 0.0;
  if (threadIdx.x == 0) { stop = 0; sc->ok = 1; }
  __syncthreads();
  Mat N; N.re = base; N.im = base + d * C::S;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (wid < NSW) {
    long long t0 = 0, acc = 0;
    for (int it = 0; it < reps; it++) {
      // refill with a diagonally dominant complex matrix (service warps only)
      for (int e = threadIdx.x; e < d * d; e += NSW * 32) {
        const int r = e / d, c = e % d;
        N.re[r * C::S + c] = (r == c ? 3.0 : 0.0) + 0.01 * ((e * 7 + it) % 13);
        N.im[r * C::S + c] = 0.02 * ((e * 5 + it) % 11);
      }
      bar_sync_i<8>(NSW * 32);
      t0 = clock64();
      service_inverse<C>(N, d, sc, wid, lane);
      bar_sync_i<8>(NSW * 32);
      acc += clock64() - t0;
    }
    if (threadIdx.x == 0) { out[blockIdx.x] = acc; stop = 1; }
    sink[threadIdx.x] = N.re[lane];
  } else {
    double c0 = threadIdx.x, c1 = 1.0, c2 = 0.5, c3 = 2.0, a = 1e-9, b = 1.0;
    volatile int* vs = &stop;
    while (!*vs) {
#pragma unroll
      for (int i = 0; i < 32; i++) { dmma(c0, c1, a, b); dmma(c2, c3, a, b); }
    }
    sink[threadIdx.x] = c0 + c1 + c2 + c3;
  }
}

int main() {
  int d = 27, reps = 200;
  long long* out; double* sink;
  cudaMalloc(&out, 8 * 148); cudaMalloc(&sink, 8 * 1024);
  size_t smem = (size_t)(2 * 2 * d * C::S + 8 * C::S) * 8 + sizeof(SvcScratch) + 64;
#define RUN(NC)                                                                                                    \
  {                                                                                                                \
    cudaFuncSetAttribute(bench<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                       \
    bench<NC><<<148, (NSW + NC) * 32, smem>>>(d, reps, out, sink);                                                 \
    bench<NC><<<148, (NSW + NC) * 32, smem>>>(d, reps, out, sink);                                                 \
    cudaError_t e = cudaDeviceSynchronize();                                                                       \
    long long h[148]; cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);                                        \
    printf("{\"dmma_warps\": %d, \"cycles_per_inverse\": %.0f, \"err\": \"%s\"}\n", NC, (double)h[0] / reps, cudaGetErrorString(e)); \
  }
  RUN(0) RUN(4) RUN(8)
  return 0;
}

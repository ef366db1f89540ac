// Micro-benchmark of one K1 "phase" (tile loop + epilogue + compute-warp barrier) in the K1 CTA configuration:
// 8 compute warps, d = 27, S = 28, one CTA per SM, matrices resident in shared memory.  Reports cycles per phase.
#include <cstdio>
#include "../quantumoptimalcontrol.jl_b200/csrc/qoc_k1.cuh"
using namespace qoc;
typedef Cfg<4, 28, 7> C;

template <int MODE>
__global__ void __launch_bounds__(C::NTHREADS, 1) bench(int d, int reps, long long* out, double* sink) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* base = reinterpret_cast<double*>(smem_raw);
  const int slot_d = 2 * d * C::S;
  K1Ctx<C> c;
  c.d = d; c.slot_d = slot_d; c.n2 = slot_d / 2; c.tid = threadIdx.x; c.lane = threadIdx.x & 31; c.warp = threadIdx.x >> 5;
  c.mi = c.warp / (C::NT / C::BN); c.nj0 = (c.warp % (C::NT / C::BN)) * C::BN;
  c.base = base; c.plane = d * C::S; c.role_slot = c.lane;
  for (int e = threadIdx.x; e < 9 * slot_d; e += C::NTHREADS) base[e] = 0.0;
  __syncthreads();
  for (int e = threadIdx.x; e < d * d; e += C::NTHREADS) {
    int r = e / d, cc = e % d;
    for (int i = 0; i < 4; i++) { c.fixed(i).re[r * C::S + cc] = 1e-3 * ((e + i) % 7) ; c.fixed(i).im[r * C::S + cc] = 1e-3 * ((e + 2 * i) % 5); }
  }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < reps; it++) {
    if (MODE == 0) c.mm1(c.fixed(4 + (it & 1)), c.fixed(0), c.fixed(1), NoEpi());
    if (MODE == 1) c.mm2(c.fixed(4 + (it & 1)), c.fixed(0), c.fixed(1), c.fixed(2), c.fixed(3), NoEpi());
    if (MODE == 2) c.mm1(c.fixed(4 + (it & 1)), c.fixed(0), c.fixed(1), c.epi(1.0, 0.5, c.fixed(2), 0.25, c.fixed(3), 0.125, c.fixed(0), 1.0));
    if (MODE == 3) { Acc<C::BN> acc; acc.zero(); mm_acc<C, false>(acc, c.fixed(0), c.fixed(1), c.mi, c.nj0, c.lane); sink[threadIdx.x] += acc.re[0][0] + acc.im[1][1]; }
    if (MODE == 4) { c.lc(c.fixed(4), 0.5, c.fixed(0), 0.25, c.fixed(1), 0.125, c.fixed(2), 0.0); c.cbar(); }
    if (MODE == 5) { c.cbar(); }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[threadIdx.x] += c.fixed(4).re[threadIdx.x % (d * C::S)];
}

int main() {
  int d = 27, reps = 2000;
  long long* out; double* sink;
  cudaMalloc(&out, 8 * 148); cudaMalloc(&sink, 8 * 1024);
  size_t smem = (size_t)(9 * 2 * d * C::S + 8 * C::S) * 8;
  const char* names[] = {"mm1 (1 product, plain store)", "mm2 (2 products)", "mm1 + LinEpi(3 terms + I)", "mm_acc only (no store, no barrier)", "lincomb + barrier", "barrier only"};
#define RUN(M)                                                                                                     \
  {                                                                                                                \
    cudaFuncSetAttribute(bench<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                        \
    bench<M><<<148, C::NTHREADS, smem>>>(d, reps, out, sink);                                                      \
    bench<M><<<148, C::NTHREADS, smem>>>(d, reps, out, sink);                                                      \
    cudaError_t e = cudaDeviceSynchronize();                                                                       \
    long long h[148]; cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);                                        \
    printf("{\"phase\": \"%s\", \"cycles_per_phase\": %.1f, \"err\": \"%s\"}\n", names[M], (double)h[0] / reps, cudaGetErrorString(e)); \
  }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
  return 0;
}

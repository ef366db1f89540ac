// Interference micro-benchmark: 8 compute warps loop over DMMA product phases while 4 service warps run
// (mode 0) nothing, (1) the register Gauss-Jordan inverse, (2) only independent DFMA bursts, (3) only barriers + LDS,
// (4) dependent DFMA chain.  Reports cycles per mm1 phase as seen by the compute warps, and cycles per service iteration.
#include <cstdio>
#include "../quantumoptimalcontrol.jl_b200/csrc/qoc_k1.cuh"
using namespace qoc;
typedef Cfg<4, 28, 7> C;

template <int MODE, int CM>
__global__ void __launch_bounds__(C::NTHREADS + NSW * 32, 1) bench(int d, int reps, long long* out, double* sink, volatile int* stop) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* base = reinterpret_cast<double*>(smem_raw);
  const int slot_d = 2 * d * C::S;
  K1Ctx<C> c;
  c.d = d; c.slot_d = slot_d; c.n2 = slot_d / 2; c.tid = threadIdx.x; c.lane = threadIdx.x & 31; c.warp = threadIdx.x >> 5;
  c.mi = c.warp / (C::NT / C::BN); c.nj0 = (c.warp % (C::NT / C::BN)) * C::BN;
  for (int i = 0; i < 8; i++) { c.s[i].re = base + (size_t)i * slot_d; c.s[i].im = c.s[i].re + d * C::S; }
  SvcScratch* sc = reinterpret_cast<SvcScratch*>(base + 9 * slot_d + 8 * C::S);
  __shared__ int done;
  for (int e = threadIdx.x; e < 9 * slot_d + 8 * C::S; e += blockDim.x) base[e] = 0.0;
  if (threadIdx.x == 0) done = 0;
  __syncthreads();
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
    int r = e / d, cc = e % d;
    for (int i = 0; i < 4; i++) { c.s[i].re[r * C::S + cc] = 1e-3 * ((e + i) % 7) + (r == cc); c.s[i].im[r * C::S + cc] = 1e-3 * ((e + 2 * i) % 5); }
  }
  __syncthreads();
  if (c.warp >= C::NW) {
    const int sw = c.warp - C::NW;
    long long t0 = clock64(); int n = 0;
    double a0 = 1.0 + c.lane, a1 = 2.0, a2 = 3.0, a3 = 4.0, a4 = 5.0, a5 = 6.0, a6 = 7.0, a7 = 8.0;
    __shared__ int svc_stop;
    while (true) {
      if (threadIdx.x == C::NTHREADS) svc_stop = *(volatile int*)&done;
      bar_svc();
      const int st = *(volatile int*)&svc_stop;
      bar_svc();
      if (st) break;
      if (MODE == 1) { // restore a well-conditioned matrix each time, then invert
        for (int e = threadIdx.x - C::NTHREADS; e < d * d; e += NSW * 32) { int r = e / d, cc = e % d; c.s[7].re[r * C::S + cc] = 1e-3 * (e % 7) + 2.0 * (r == cc); c.s[7].im[r * C::S + cc] = 1e-3 * (e % 5); }
        bar_svc();
        service_inverse<C>(c.s[7], d, sc, sw, c.lane);
      }
      if (MODE == 2) {
#pragma unroll
        for (int r = 0; r < 16; r++) { a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9); a2 = fma(a2, 1.0000001, 1e-9); a3 = fma(a3, 1.0000001, 1e-9);
                                       a4 = fma(a4, 1.0000001, 1e-9); a5 = fma(a5, 1.0000001, 1e-9); a6 = fma(a6, 1.0000001, 1e-9); a7 = fma(a7, 1.0000001, 1e-9); }
      }
      if (MODE == 3) { for (int r = 0; r < 16; r++) { bar_svc(); a0 += sc->gbuf[r & 1][c.lane].x; } }
      if (MODE == 4) {
#pragma unroll
        for (int r = 0; r < 64; r++) a0 = fma(a0, 1.0000001, 1e-9);
      }
      if (MODE == 5) { for (int r = 0; r < 64; r++) a0 += __shfl_xor_sync(0xffffffffu, a0, 1) * 1e-9f; }
      n++;
    }
    long long t1 = clock64();
    if (threadIdx.x == C::NTHREADS) { out[148 + blockIdx.x] = (t1 - t0) / (n > 0 ? n : 1); }
    sink[threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    return;
  }
  long long t0 = clock64();
  for (int it = 0; it < reps; it++) {
    if (CM == 0) c.mm1(c.s[4 + (it & 1)], c.s[0], c.s[1], NoEpi());
    if (CM == 1) c.mm1(c.s[4 + (it & 1)], c.s[0], c.s[1], c.epi(1.0, 0.5, c.s[2], 0.25, c.s[3], 0.125, c.s[0], 1.0));
    if (CM == 2) { c.lc(c.s[4], 0.5, c.s[0], 0.25, c.s[1], 0.125, c.s[2], 0.0); c.cbar(); }
    if (CM == 3) c.mm2(c.s[4 + (it & 1)], c.s[0], c.s[1], c.s[2], c.s[3], NoEpi());
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x] = t1 - t0; done = 1; }
  sink[threadIdx.x] += c.s[4].re[threadIdx.x % (d * C::S)];
}

int main() {
  int d = 27, reps = 500;
  long long* out; double* sink; int* stop;
  cudaMalloc(&out, 8 * 296); cudaMalloc(&sink, 8 * 1024); cudaMalloc(&stop, 4);
  size_t smem = (size_t)(9 * 2 * d * C::S + 8 * C::S) * 8 + sizeof(SvcScratch) + 64;
  const char* names[] = {"service idle", "service: GJ inverse", "service: independent DFMA bursts (8-way ILP)", "service: bar + LDS only", "service: dependent DFMA chain", "service: SHFL + FADD chain"};
#define RUN2(M, CMM)                                                                                                     \
  {                                                                                                                \
    cudaFuncSetAttribute(bench<M, CMM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                        \
    bench<M, CMM><<<148, C::NTHREADS + NSW * 32, smem>>>(d, reps, out, sink, stop);                                     \
    cudaError_t e = cudaDeviceSynchronize();                                                                       \
    long long h[296]; cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);                                        \
    printf("{\"compute\": \"%s\", \"mode\": \"%s\", \"cycles_per_phase\": %.1f, \"cycles_per_service_iter\": %lld, \"err\": \"%s\"}\n", cn[CMM], names[M], (double)h[0] / reps, h[148], cudaGetErrorString(e)); \
  }
  const char* cn[] = {"mm1", "mm1+LinEpi", "lincomb", "mm2"};
  RUN2(0, 0) RUN2(1, 0) RUN2(0, 1) RUN2(1, 1) RUN2(0, 2) RUN2(1, 2) RUN2(0, 3) RUN2(1, 3)
  return 0;
}

// Micro-benchmark of variants of the 3M complex tile loop (K1 CTA shape: 8 compute warps, d = 27, S = 28, KS = 7, 1 x 2 tiles
// per warp), to find out where the gap between a product phase (~1.95 k cycles) and the DMMA issue floor (1.39 k) goes.
// Each variant runs `reps` back-to-back tile loops per warp with no store and no barrier; cycles per loop are printed.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int S = 28, KS = 7, BN = 2, NW = 8;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double dadd_v(double a, double b) {
  double r;
  asm volatile("add.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b));
  return r;
}

struct Acc {
  double re[BN][2], im[BN][2], t2[BN][2];
};

// V0: the loop as K1 has it (sums formed right where they are used)
__device__ __forceinline__ void v0(Acc& acc, const double* are, const double* aim, const double* bre, const double* bim) {
#pragma unroll
  for (int ks = 0; ks < KS; ks++) {
    double ar = are[ks * 4], ai = aim[ks * 4];
    const double as = ar + ai;
#pragma unroll
    for (int n = 0; n < BN; n++) {
      const double br = bre[ks * 4 * S + n * 8], bi = bim[ks * 4 * S + n * 8];
      const double bs = br + bi;
      dmma(acc.re[n][0], acc.re[n][1], ar, br);
      dmma(acc.t2[n][0], acc.t2[n][1], ai, bi);
      dmma(acc.im[n][0], acc.im[n][1], as, bs);
    }
  }
}
// V1: all fragments of the phase loaded and summed up front (volatile adds pin the order), then 42 DMMAs back to back
__device__ __forceinline__ void v1(Acc& acc, const double* are, const double* aim, const double* bre, const double* bim) {
  double ar[KS], ai[KS], as[KS], br[KS][BN], bi[KS][BN], bs[KS][BN];
#pragma unroll
  for (int ks = 0; ks < KS; ks++) {
    ar[ks] = are[ks * 4]; ai[ks] = aim[ks * 4];
#pragma unroll
    for (int n = 0; n < BN; n++) { br[ks][n] = bre[ks * 4 * S + n * 8]; bi[ks][n] = bim[ks * 4 * S + n * 8]; }
  }
#pragma unroll
  for (int ks = 0; ks < KS; ks++) {
    as[ks] = dadd_v(ar[ks], ai[ks]);
#pragma unroll
    for (int n = 0; n < BN; n++) bs[ks][n] = dadd_v(br[ks][n], bi[ks][n]);
  }
#pragma unroll
  for (int ks = 0; ks < KS; ks++)
#pragma unroll
    for (int n = 0; n < BN; n++) {
      dmma(acc.re[n][0], acc.re[n][1], ar[ks], br[ks][n]);
      dmma(acc.t2[n][0], acc.t2[n][1], ai[ks], bi[ks][n]);
      dmma(acc.im[n][0], acc.im[n][1], as[ks], bs[ks][n]);
    }
}
// V2: no sums at all (wrong numerics; DMMA + LDS only)
__device__ __forceinline__ void v2(Acc& acc, const double* are, const double* aim, const double* bre, const double* bim) {
#pragma unroll
  for (int ks = 0; ks < KS; ks++) {
    double ar = are[ks * 4], ai = aim[ks * 4];
#pragma unroll
    for (int n = 0; n < BN; n++) {
      const double br = bre[ks * 4 * S + n * 8], bi = bim[ks * 4 * S + n * 8];
      dmma(acc.re[n][0], acc.re[n][1], ar, br);
      dmma(acc.t2[n][0], acc.t2[n][1], ai, bi);
      dmma(acc.im[n][0], acc.im[n][1], ai, br);
    }
  }
}
// V3: the conventional 4-product form (no sums, 4 DMMAs per complex tile step)
__device__ __forceinline__ void v3(Acc& acc, const double* are, const double* aim, const double* bre, const double* bim) {
#pragma unroll
  for (int ks = 0; ks < KS; ks++) {
    double ar = are[ks * 4], ai = aim[ks * 4];
    double nai = -ai;
#pragma unroll
    for (int n = 0; n < BN; n++) {
      const double br = bre[ks * 4 * S + n * 8], bi = bim[ks * 4 * S + n * 8];
      dmma(acc.re[n][0], acc.re[n][1], ar, br);
      dmma(acc.im[n][0], acc.im[n][1], ar, bi);
      dmma(acc.re[n][0], acc.re[n][1], nai, bi);
      dmma(acc.im[n][0], acc.im[n][1], ai, br);
    }
  }
}
// V4: pure DMMA issue, operands in registers (the floor)
__device__ __forceinline__ void v4(Acc& acc, const double* are, const double* aim, const double* bre, const double* bim) {
  double ar = are[0], ai = aim[0], br = bre[0], bi = bim[0];
#pragma unroll
  for (int ks = 0; ks < KS; ks++)
#pragma unroll
    for (int n = 0; n < BN; n++) {
      dmma(acc.re[n][0], acc.re[n][1], ar, br);
      dmma(acc.t2[n][0], acc.t2[n][1], ai, bi);
      dmma(acc.im[n][0], acc.im[n][1], ai, br);
    }
}
// V5: pipelined by one k-step with pinned order: loads of ks+1, T1/T2 of ks, sums of ks+1, T3 of ks
__device__ __forceinline__ void v5(Acc& acc, const double* are, const double* aim, const double* bre, const double* bim) {
  double ar = are[0], ai = aim[0], as = dadd_v(ar, ai), br[BN], bi[BN], bs[BN];
#pragma unroll
  for (int n = 0; n < BN; n++) { br[n] = bre[n * 8]; bi[n] = bim[n * 8]; bs[n] = dadd_v(br[n], bi[n]); }
#pragma unroll
  for (int ks = 0; ks < KS; ks++) {
    double nar = 0, nai = 0, nas = 0, nbr[BN], nbi[BN], nbs[BN];
    if (ks + 1 < KS) {
      nar = are[(ks + 1) * 4]; nai = aim[(ks + 1) * 4];
#pragma unroll
      for (int n = 0; n < BN; n++) { nbr[n] = bre[(ks + 1) * 4 * S + n * 8]; nbi[n] = bim[(ks + 1) * 4 * S + n * 8]; }
    }
#pragma unroll
    for (int n = 0; n < BN; n++) {
      dmma(acc.re[n][0], acc.re[n][1], ar, br[n]);
      dmma(acc.t2[n][0], acc.t2[n][1], ai, bi[n]);
    }
    if (ks + 1 < KS) {
      nas = dadd_v(nar, nai);
#pragma unroll
      for (int n = 0; n < BN; n++) nbs[n] = dadd_v(nbr[n], nbi[n]);
    }
#pragma unroll
    for (int n = 0; n < BN; n++) dmma(acc.im[n][0], acc.im[n][1], as, bs[n]);
    ar = nar; ai = nai; as = nas;
#pragma unroll
    for (int n = 0; n < BN; n++) { br[n] = nbr[n]; bi[n] = nbi[n]; bs[n] = nbs[n]; }
  }
}
// V6: sums read from a third plane (what a stored re+im plane would cost: 3 LDS per fragment, no DADD)
__device__ __forceinline__ void v6(Acc& acc, const double* are, const double* aim, const double* bre, const double* bim) {
#pragma unroll
  for (int ks = 0; ks < KS; ks++) {
    double ar = are[ks * 4], ai = aim[ks * 4], as = are[ks * 4 + 2 * 27 * S];
#pragma unroll
    for (int n = 0; n < BN; n++) {
      const double br = bre[ks * 4 * S + n * 8], bi = bim[ks * 4 * S + n * 8], bs = bre[ks * 4 * S + n * 8 + 2 * 27 * S];
      dmma(acc.re[n][0], acc.re[n][1], ar, br);
      dmma(acc.t2[n][0], acc.t2[n][1], ai, bi);
      dmma(acc.im[n][0], acc.im[n][1], as, bs);
    }
  }
}

template <int V, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) bench(int reps, long long* out, double* sink) {
  extern __shared__ __align__(16) double sm[];
  const int d = 27;
  for (int e = threadIdx.x; e < 8 * 2 * d * S + 64 * S; e += blockDim.x) sm[e] = 1e-3 * (e % 13);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) % NW, g = lane >> 2, q = lane & 3;
  const int mi = warp / 2, nj0 = (warp % 2) * 2;
  const double* A = sm;
  const double* B = sm + 2 * d * S;
  const double* are = A + (mi * 8 + g) * S + q;
  const double* aim = are + d * S;
  const double* bre = B + q * S + nj0 * 8 + g;
  const double* bim = bre + d * S;
  Acc acc;
#pragma unroll
  for (int n = 0; n < BN; n++) acc.re[n][0] = acc.re[n][1] = acc.im[n][0] = acc.im[n][1] = acc.t2[n][0] = acc.t2[n][1] = 0.0;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < reps; it++) {
    if (V == 0) v0(acc, are, aim, bre, bim);
    if (V == 1) v1(acc, are, aim, bre, bim);
    if (V == 2) v2(acc, are, aim, bre, bim);
    if (V == 3) v3(acc, are, aim, bre, bim);
    if (V == 4) v4(acc, are, aim, bre, bim);
    if (V == 5) v5(acc, are, aim, bre, bim);
    if (V == 6) v6(acc, are, aim, bre, bim);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  double s = 0;
#pragma unroll
  for (int n = 0; n < BN; n++) s += acc.re[n][0] + acc.re[n][1] + acc.im[n][0] + acc.im[n][1] + acc.t2[n][0] + acc.t2[n][1];
  sink[threadIdx.x] = s;
}

int main() {
  int reps = 2000;
  long long* out; double* sink;
  cudaMalloc(&out, 8 * 148); cudaMalloc(&sink, 8 * 1024);
  size_t smem = (size_t)(8 * 2 * 27 * S + 64 * S) * 8;
  const char* names[] = {"v0 sums at use (K1 today)", "v1 all loads+sums up front", "v2 no sums", "v3 4M", "v4 DMMA only (registers)",
                         "v5 pipelined one k-step, pinned", "v6 sums from a third plane"};
#define RUN(V, W)                                                                                                  \
  {                                                                                                                \
    cudaFuncSetAttribute(bench<V, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                     \
    bench<V, W><<<148, W * 32, smem>>>(reps, out, sink);                                                           \
    bench<V, W><<<148, W * 32, smem>>>(reps, out, sink);                                                           \
    cudaError_t e = cudaDeviceSynchronize();                                                                       \
    long long h[148]; cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);                                        \
    printf("{\"variant\": \"%s\", \"warps\": %d, \"cycles_per_loop\": %.1f, \"err\": \"%s\"}\n", names[V], W, (double)h[0] / reps, cudaGetErrorString(e)); \
  }
  RUN(0, 8) RUN(1, 8) RUN(2, 8) RUN(3, 8) RUN(4, 8) RUN(5, 8) RUN(6, 8)
  RUN(0, 16) RUN(2, 16) RUN(4, 16) RUN(5, 16)
  return 0;
}

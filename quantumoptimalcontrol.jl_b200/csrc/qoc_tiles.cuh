// qoc_tiles.cuh -- shared-memory complex matrix primitives built on the FP64 tensor-core tile DMMA.8x8x4
// (PTX mma.sync.aligned.m8n8k4.f64) for sm_100a.
//
// Measured on this pool's B200 (profiles/r01_fp64_peak.jsonl): DMMA m8n8k4 sustains 37.1 TFLOP/s (= 148 SMs x
// 64 FMA/clk x 2 x 1.965 GHz), a pure DFMA register loop 34.1, DFMA fed from shared memory 25.8.  DMMA needs one
// issue slot per 256 FMAs, so every d x d x d contraction of the path goes through it.
//
// Layout ("planar slot"): a complex d x d matrix is two real planes (re, then im), each ROW-MAJOR with row
// stride S doubles, S = 4 (mod 8), S >= roundup(d,4).  With that stride both fragment shapes are bank-conflict
// free for 64-bit shared loads:
//    A fragment (8x4, thread t holds A[t/4][t%4])  -> address (r0 + t/4)*S + k0 + t%4
//    B fragment (4x8, thread t holds B[t%4][t/4])  -> address (k0 + t%4)*S + n0 + t/4
// (in each half-warp the 16 addresses are distinct mod 16).  The same stored matrix can therefore be used as a
// left operand, a right operand, or -- reading it with the B pattern and negating the imaginary plane -- as a
// conjugate-transposed left operand (U^dagger for the costate sweep) without any transposed copy.
// Columns d..S-1 of every row are kept exactly zero (masked stores), which is what makes the k-padding of the
// last k-step harmless; rows >= d are never stored.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qoc {

struct Mat {
  double* re;
  double* im;
};

// compile-time shape class: NT = 8x8 tiles per dimension, S = row stride, KS = k-steps of 4
template <int NT_, int S_, int KS_>
struct Cfg {
  static constexpr int NT = NT_, S = S_, KS = KS_;
  static constexpr int BN = (NT_ == 4) ? 2 : 1;       // tiles per warp block (1 x BN)
  static constexpr int NBLK = NT_ * (NT_ / BN);       // warp blocks == warps that do DMMA work
  static constexpr int NW = NBLK;                      // warps per CTA in K1
  static constexpr int NTHREADS = NW * 32;
  static constexpr int DMAX = (8 * NT_ < 4 * KS_) ? 8 * NT_ : 4 * KS_;  // largest d this class serves
  static constexpr int MINB = (NT_ <= 2) ? 3 : 1;     // K1 CTAs per SM the register budget is sized for
};

// Complex products are formed with THREE real tile products (the "3M" scheme, as in BLAS zgemm3m):
//   T1 = Ar Br, T2 = Ai Bi, T3 = (Ar + Ai)(Br + Bi);   Re C = T1 - T2,  Im C = T3 - T1 - T2
// i.e. 6 d^3 executed flops per complex product instead of 8 d^3 (-25% DMMA issue) for two extra adds per operand
// fragment.  Normwise stable (error <= c u ||A|| ||B||); the parity tests hold the result to 1e-12.
// Compile with -DQOC_4M to get the conventional four-product form.
#ifndef QOC_4M
#define QOC_3M 1
#endif

#ifdef QOC_3M
template <int BN>
struct Acc {
  double re[BN][2];   // T1 while accumulating; Re C after finish()
  double im[BN][2];   // T3 while accumulating; Im C after finish()
  double t2[BN][2];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int n = 0; n < BN; n++) re[n][0] = re[n][1] = im[n][0] = im[n][1] = t2[n][0] = t2[n][1] = 0.0;
  }
  __device__ __forceinline__ void finish() {
#pragma unroll
    for (int n = 0; n < BN; n++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const double t1 = re[n][e];
        re[n][e] = t1 - t2[n][e];
        im[n][e] = (im[n][e] - t1) - t2[n][e];
      }
  }
};
#else
template <int BN>
struct Acc {
  double re[BN][2];
  double im[BN][2];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int n = 0; n < BN; n++) re[n][0] = re[n][1] = im[n][0] = im[n][1] = 0.0;
  }
  __device__ __forceinline__ void finish() {}
};
#endif

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// acc += op(A) * B for this warp's 1 x BN block of 8x8 tiles (tile row mi, tile columns nj0..nj0+BN-1).
// ADJ = false: op(A) = A ; ADJ = true: op(A) = A^dagger (conjugate transpose of the stored matrix).
template <class C, bool ADJ>
__device__ __forceinline__ void mm_acc(Acc<C::BN>& acc, Mat A, Mat B, int mi, int nj0, int lane) {
  constexpr int S = C::S;
  const int g = lane >> 2, q = lane & 3;
  const double* are;
  const double* aim;
  int astep;
  if (ADJ) {
    are = A.re + q * S + mi * 8 + g;
    aim = A.im + q * S + mi * 8 + g;
    astep = 4 * S;
  } else {
    are = A.re + (mi * 8 + g) * S + q;
    aim = A.im + (mi * 8 + g) * S + q;
    astep = 4;
  }
  const double* bre = B.re + q * S + nj0 * 8 + g;
  const double* bim = B.im + q * S + nj0 * 8 + g;
#pragma unroll
  for (int ks = 0; ks < C::KS; ks++) {
    double ar = are[ks * astep];
    double ai = aim[ks * astep];
    if (ADJ) ai = -ai;
#ifdef QOC_3M
    const double as = ar + ai;
#pragma unroll
    for (int n = 0; n < C::BN; n++) {
      const double br = bre[ks * 4 * S + n * 8];
      const double bi = bim[ks * 4 * S + n * 8];
      const double bs = br + bi;
      dmma(acc.re[n][0], acc.re[n][1], ar, br);   // T1
      dmma(acc.t2[n][0], acc.t2[n][1], ai, bi);   // T2
      dmma(acc.im[n][0], acc.im[n][1], as, bs);   // T3
    }
#else
    double nai = -ai;
#pragma unroll
    for (int n = 0; n < C::BN; n++) {
      double br = bre[ks * 4 * S + n * 8];
      double bi = bim[ks * 4 * S + n * 8];
      dmma(acc.re[n][0], acc.re[n][1], ar, br);
      dmma(acc.im[n][0], acc.im[n][1], ar, bi);
      dmma(acc.re[n][0], acc.re[n][1], nai, bi);
      dmma(acc.im[n][0], acc.im[n][1], ai, br);
    }
#endif
  }
}

// masked store of the warp block; f(row, col, re0, im0, re1, im1) may modify the two adjacent elements
// (row, col) and (row, col+1) before they are written.  Keeps the zero padding of dst intact.
template <class C, class F>
__device__ __forceinline__ void mm_store(Mat dst, Acc<C::BN>& acc, int d, int mi, int nj0, int lane, F f) {
  constexpr int S = C::S;
  acc.finish();
  const int row = mi * 8 + (lane >> 2);
#pragma unroll
  for (int n = 0; n < C::BN; n++) {
    const int col = (nj0 + n) * 8 + 2 * (lane & 3);
    if (row < d && col < d) {
      double r0 = acc.re[n][0], r1 = acc.re[n][1], i0 = acc.im[n][0], i1 = acc.im[n][1];
      f(row, col, r0, i0, r1, i1);
      if (col + 1 >= d) { r1 = 0.0; i1 = 0.0; }
      *reinterpret_cast<double2*>(dst.re + row * S + col) = make_double2(r0, r1);
      *reinterpret_cast<double2*>(dst.im + row * S + col) = make_double2(i0, i1);
    }
  }
}

struct NoEpi {
  __device__ __forceinline__ void operator()(int, int, double&, double&, double&, double&) const {}
};

// epilogue: acc*alpha + c1*M1 + c2*M2 + c3*M3 + cI*I, read at the output position
template <int S>
struct LinEpi {
  double alpha, c1, c2, c3, cI;
  Mat m1, m2, m3;
  __device__ __forceinline__ void operator()(int row, int col, double& r0, double& i0, double& r1, double& i1) const {
    const int o = row * S + col;
    r0 *= alpha; i0 *= alpha; r1 *= alpha; i1 *= alpha;
    if (c1 != 0.0) {
      double2 a = *reinterpret_cast<const double2*>(m1.re + o), b = *reinterpret_cast<const double2*>(m1.im + o);
      r0 = fma(c1, a.x, r0); r1 = fma(c1, a.y, r1); i0 = fma(c1, b.x, i0); i1 = fma(c1, b.y, i1);
    }
    if (c2 != 0.0) {
      double2 a = *reinterpret_cast<const double2*>(m2.re + o), b = *reinterpret_cast<const double2*>(m2.im + o);
      r0 = fma(c2, a.x, r0); r1 = fma(c2, a.y, r1); i0 = fma(c2, b.x, i0); i1 = fma(c2, b.y, i1);
    }
    if (c3 != 0.0) {
      double2 a = *reinterpret_cast<const double2*>(m3.re + o), b = *reinterpret_cast<const double2*>(m3.im + o);
      r0 = fma(c3, a.x, r0); r1 = fma(c3, a.y, r1); i0 = fma(c3, b.x, i0); i1 = fma(c3, b.y, i1);
    }
    if (row == col) r0 += cI;
    if (row == col + 1) r1 += cI;
  }
};

// ---- CTA-wide elementwise helpers over whole planar slots (pad columns stay zero: lincombs of zeros) -------

// dst = c1*M1 + c2*M2 + c3*M3 + cI*I   (dst may alias any source)
template <int S>
__device__ __forceinline__ void lincomb(Mat dst, int d, double c1, Mat m1, double c2, Mat m2, double c3, Mat m3,
                                        double cI, int tid, int nthreads) {
  const int n2 = d * S / 2;
  for (int e = tid; e < n2; e += nthreads) {
    double2 r = make_double2(0.0, 0.0), i = make_double2(0.0, 0.0);
    if (c1 != 0.0) {
      double2 a = reinterpret_cast<const double2*>(m1.re)[e], b = reinterpret_cast<const double2*>(m1.im)[e];
      r.x = c1 * a.x; r.y = c1 * a.y; i.x = c1 * b.x; i.y = c1 * b.y;
    }
    if (c2 != 0.0) {
      double2 a = reinterpret_cast<const double2*>(m2.re)[e], b = reinterpret_cast<const double2*>(m2.im)[e];
      r.x = fma(c2, a.x, r.x); r.y = fma(c2, a.y, r.y); i.x = fma(c2, b.x, i.x); i.y = fma(c2, b.y, i.y);
    }
    if (c3 != 0.0) {
      double2 a = reinterpret_cast<const double2*>(m3.re)[e], b = reinterpret_cast<const double2*>(m3.im)[e];
      r.x = fma(c3, a.x, r.x); r.y = fma(c3, a.y, r.y); i.x = fma(c3, b.x, i.x); i.y = fma(c3, b.y, i.y);
    }
    if (cI != 0.0) {
      const int row = (2 * e) / S, col = (2 * e) - row * S;
      if (row == col) r.x += cI;
      if (row == col + 1) r.y += cI;
    }
    reinterpret_cast<double2*>(dst.re)[e] = r;
    reinterpret_cast<double2*>(dst.im)[e] = i;
  }
}

// (a, b) <- (a - b, a + b) in place
template <int S>
__device__ __forceinline__ void diff_sum_inplace(Mat a, Mat b, int d, int tid, int nthreads) {
  const int n = d * S;  // re plane then im plane are contiguous: treat the slot as 2*d*S doubles
  double2* pa = reinterpret_cast<double2*>(a.re);
  double2* pb = reinterpret_cast<double2*>(b.re);
  for (int e = tid; e < n; e += nthreads) {
    double2 x = pa[e], y = pb[e];
    pa[e] = make_double2(x.x - y.x, x.y - y.y);
    pb[e] = make_double2(x.x + y.x, x.y + y.y);
  }
}

// slot copy shared/global <-> shared/global with optional real scale; n2 = number of double2 in the slot
__device__ __forceinline__ void slot_copy(double* dst, const double* src, int n2, int tid, int nthreads) {
  double2* pd = reinterpret_cast<double2*>(dst);
  const double2* ps = reinterpret_cast<const double2*>(src);
  for (int e = tid; e < n2; e += nthreads) pd[e] = ps[e];
}
__device__ __forceinline__ void slot_copy_scaled(double* dst, const double* src, double sc, int n2, int tid,
                                                 int nthreads) {
  double2* pd = reinterpret_cast<double2*>(dst);
  const double2* ps = reinterpret_cast<const double2*>(src);
  for (int e = tid; e < n2; e += nthreads) {
    double2 v = ps[e];
    pd[e] = make_double2(v.x * sc, v.y * sc);
  }
}

}  // namespace qoc

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
  static constexpr int MAXE = (2 * NT_ * NT_ + NW - 1) / NW;  // matrix elements per thread in the GJ inverse
};

template <int BN>
struct Acc {
  double re[BN][2];
  double im[BN][2];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int n = 0; n < BN; n++) re[n][0] = re[n][1] = im[n][0] = im[n][1] = 0.0;
  }
};

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
  }
}

// masked store of the warp block; f(row, col, re0, im0, re1, im1) may modify the two adjacent elements
// (row, col) and (row, col+1) before they are written.  Keeps the zero padding of dst intact.
template <class C, class F>
__device__ __forceinline__ void mm_store(Mat dst, const Acc<C::BN>& acc, int d, int mi, int nj0, int lane, F f) {
  constexpr int S = C::S;
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

// 1-norm (max column sum of |a_ij|) of a planar matrix; result broadcast to every thread.
// scratch: >= 32 doubles of shared memory.  Contains two __syncthreads.
template <int S>
__device__ __forceinline__ double norm1(Mat a, int d, double* scratch, int tid, int nthreads) {
  if (tid < 32) scratch[tid] = 0.0;
  __syncthreads();
  // thread handles (col, part): parts split the rows; columns up to 32
  const int nparts = nthreads / 32 > 0 ? nthreads / 32 : 1;
  const int col = tid & 31, part = tid >> 5;
  double s = 0.0;
  if (col < d) {
    for (int r = part; r < d; r += nparts) {
      double x = a.re[r * S + col], y = a.im[r * S + col];
      s += sqrt(x * x + y * y);
    }
    // non-negative doubles order like their bit patterns, but we need a SUM over parts: use atomicAdd
    atomicAdd(&scratch[col], s);
  }
  __syncthreads();
  double mx = 0.0;
  for (int c = 0; c < d; c++) mx = fmax(mx, scratch[c]);
  return mx;
}

// In-place Gauss-Jordan inverse with partial (row) pivoting of a planar d x d matrix, CTA-wide.
// The matrix lives in registers (MAXE elements per thread) for the whole elimination; per step only the pivot
// row, the displaced row and the multiplier column are exchanged through shared memory (2 barriers / step).
// buf: 5*d complex (rowbuf, oldk, colbuf[2]) + d ints.  Returns false (uniformly) if a pivot is exactly zero.
template <class C>
__device__ __forceinline__ bool gj_inverse(Mat a, int d, double2* buf, int tid) {
  constexpr int S = C::S;
  constexpr int NTH = C::NTHREADS;
  constexpr int MAXE = C::MAXE;
  double2* rowbuf = buf;
  double2* oldk = buf + d;
  double2* colbuf0 = buf + 2 * d;
  int* idx = reinterpret_cast<int*>(buf + 4 * d);      // d ints: column bookkeeping
  int* pivs = idx + 32;                                 // d ints
  const int lane = tid & 31;

  double vr[MAXE], vi[MAXE];
  int ei[MAXE], ec[MAXE];
#pragma unroll
  for (int t = 0; t < MAXE; t++) {
    int e = tid + t * NTH;
    if (e < d * d) {
      ei[t] = e / d;
      ec[t] = e - ei[t] * d;
      vr[t] = a.re[ei[t] * S + ec[t]];
      vi[t] = a.im[ei[t] * S + ec[t]];
    } else {
      ei[t] = -1; ec[t] = -1; vr[t] = 0.0; vi[t] = 0.0;
    }
  }
  bool ok = true;
  for (int k = 0; k < d; k++) {
    double2* colbuf = colbuf0 + (k & 1) * d;
#pragma unroll
    for (int t = 0; t < MAXE; t++)
      if (ec[t] == k) colbuf[ei[t]] = make_double2(vr[t], vi[t]);
    __syncthreads();
    // every warp finds the pivot row redundantly (rows i >= k), d <= 32
    double mag = -1.0;
    int p = lane;
    if (lane >= k && lane < d) {
      double2 z = colbuf[lane];
      mag = z.x * z.x + z.y * z.y;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      double om = __shfl_xor_sync(0xffffffffu, mag, off);
      int op = __shfl_xor_sync(0xffffffffu, p, off);
      if (om > mag || (om == mag && op < p)) { mag = om; p = op; }
    }
    if (!(mag > 0.0)) ok = false;
    if (tid == 0) pivs[k] = p;
    // publish pivot row (old row p) and the displaced row (old row k)
#pragma unroll
    for (int t = 0; t < MAXE; t++) {
      if (ei[t] == p) rowbuf[ec[t]] = make_double2(vr[t], vi[t]);
      if (ei[t] == k) oldk[ec[t]] = make_double2(vr[t], vi[t]);
    }
    __syncthreads();
    const double2 pv = rowbuf[k];
    const double den = 1.0 / (pv.x * pv.x + pv.y * pv.y);
    const double pir = pv.x * den, pii = -pv.y * den;  // 1/pivot
#pragma unroll
    for (int t = 0; t < MAXE; t++) {
      const int i = ei[t], c = ec[t];
      if (i < 0) continue;
      if (i == k) {
        if (c == k) { vr[t] = pir; vi[t] = pii; }
        else {
          double2 r = rowbuf[c];
          vr[t] = r.x * pir - r.y * pii;
          vi[t] = r.x * pii + r.y * pir;
        }
      } else {
        // after the swap row p holds old row k; every other row is itself
        double sr = vr[t], si = vi[t];
        double2 f = colbuf[i];
        if (i == p) { double2 o = oldk[c]; sr = o.x; si = o.y; f = colbuf[k]; }
        // g = f / pivot
        const double gr = f.x * pir - f.y * pii, gi = f.x * pii + f.y * pir;
        if (c == k) { vr[t] = -gr; vi[t] = -gi; }
        else {
          double2 r = rowbuf[c];
          vr[t] = sr - (gr * r.x - gi * r.y);
          vi[t] = si - (gr * r.y + gi * r.x);
        }
      }
    }
    // no barrier needed here: next step writes the other colbuf; rowbuf/oldk are rewritten only after the
    // next step's first barrier, which every thread reaches after finishing this update.
  }
  __syncthreads();
  // undo the row interchanges as column interchanges in reverse order: final column position of each column
  if (tid == 0) {
    for (int c = 0; c < d; c++) idx[c] = c;
    for (int k = d - 1; k >= 0; k--) {
      int p = pivs[k];
      if (p != k) { int t = idx[k]; idx[k] = idx[p]; idx[p] = t; }
    }
    // idx[pos] = source column sitting at position pos; invert it in place into pivs
    for (int pos = 0; pos < d; pos++) pivs[idx[pos]] = pos;
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < MAXE; t++) {
    if (ei[t] >= 0) {
      const int pos = pivs[ec[t]];
      a.re[ei[t] * S + pos] = vr[t];
      a.im[ei[t] * S + pos] = vi[t];
    }
  }
  __syncthreads();
  return ok;
}

}  // namespace qoc

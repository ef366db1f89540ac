// qoc_gpath.cuh -- general path for Hilbert-space dimensions whose working set does not fit shared memory (d > 28).
//
// Same arithmetic as K1/K2/K3 (Pade [13/13] expm, structured block-triangular Frechet derivative or the reference's
// truncated Taylor Jacobian, adjoint gradient), but every matrix lives in HBM/L2 as a planar slot and the program is
// a sequence of BATCHED launches over a chunk of slices:
//   g_build_kernel    X_k = (A0 + sum_j u_jk A_j) 2^-s                         src/gradient_computations.jl:18-22
//   g_gemm_kernel     C = alpha (A1 B1 [+ A2 B2 ...]) + sum_q beta_q D_q + gamma I   on 32x32 DMMA tiles (3M products),
//                     operands staged global -> shared per 32-wide k chunk; an operand with stride 0 is broadcast
//   g_inverse_kernel  in-place Gauss-Jordan inverse with partial pivoting, one CTA per slice
//   g_sweep_kernel    serial forward sweep, cost / terminal costate, backward sweep with the gradient contraction
//                     (:27-29, :46-58, :65-74), one CTA per pulse
// This first version favours coverage over speed (no cross-launch fusion, serial sweeps); see DESIGN.md.
#pragma once
#include "qoc_tiles.cuh"

namespace qoc {

struct GOp {
  const double* p;
  long long stride;   // doubles between consecutive slices (0: same matrix for every slice)
  int inner;          // 0: flat;  > 0: slice s sits at (s / inner) * stride2 + (s % inner) * stride
  long long stride2;
  __host__ __device__ __forceinline__ long long off(int s) const {
    return inner > 0 ? (long long)(s / inner) * stride2 + (long long)(s % inner) * stride : (long long)s * stride;
  }
};

struct GGemm {
  int d, S, nb, npairs, nadd;
  GOp A[4], B[4], D[3];
  double alpha, beta[3], gamma;
  double* C;
  long long cstride;
  int cinner;
  long long cstride2;
};

#ifndef QOC_3M
#error "qoc_gpath.cuh assumes the 3M complex product (Acc carries T1/T2/T3)"
#endif

// C = alpha (A1 B1 [+ A2 B2 ...]) + sum_q beta_q D_q + gamma I  for a batch of slices.
// CTA tile (32 WM) x (16 WN): 8 warps as 4 x 2, each warp WM x WN tiles of 8 x 8 (WM, WN = 1, 2: 32 x 32 for small d;
// 2, 4: 64 x 64, half the L2 -> shared traffic per flop and half the shared loads per DMMA, for d >= 64).
// Operands are staged global -> shared per 32-wide k chunk.  8x8 tiles and k-steps that lie entirely outside the
// d x d matrix are skipped (warp-uniform predicates; d = 40 fills 25 of the 64 tiles of its 2 x 2 CTA grid and 10 of
// 16 k-steps), the zero padding of the staged tiles covers the partial ones.
template <int WM, int WN>
__global__ void __launch_bounds__(256) g_gemm_kernel(GGemm g) {
  constexpr int TM = 32 * WM, TN = 16 * WN;   // CTA tile
  constexpr int AS = 36, BS = TN + 4;         // shared row strides (= 4 mod 8: conflict-free fragment loads)
  extern __shared__ __align__(16) unsigned char gsm_raw[];
  double* smAr = reinterpret_cast<double*>(gsm_raw);
  double* smAi = smAr + TM * AS;
  double* smBr = smAi + TM * AS;
  double* smBi = smBr + 32 * BS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int s = blockIdx.z, tm = blockIdx.y, tn = blockIdx.x;
  const int d = g.d, S = g.S, plane = d * S;
  const int gq = lane >> 2, q4 = lane & 3;
  double T1[WM][WN][2], T2[WM][WN][2], T3[WM][WN][2];
#pragma unroll
  for (int a = 0; a < WM; a++)
#pragma unroll
    for (int b = 0; b < WN; b++) { T1[a][b][0] = T1[a][b][1] = T2[a][b][0] = T2[a][b][1] = T3[a][b][0] = T3[a][b][1] = 0.0; }
  bool rowv[WM], colv[WN];
#pragma unroll
  for (int a = 0; a < WM; a++) rowv[a] = tm * TM + (wm * WM + a) * 8 < d;
#pragma unroll
  for (int b = 0; b < WN; b++) colv[b] = tn * TN + (wn * WN + b) * 8 < d;
  const int nkc = (d + 31) / 32;
  for (int p = 0; p < g.npairs; p++) {
    const double* Ag = g.A[p].p + g.A[p].off(s);
    const double* Bg = g.B[p].p + g.B[p].off(s);
    for (int kc = 0; kc < nkc; kc++) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 2 * WM; i++) {          // A tile: TM rows x 32 columns
        const int e = tid + i * 256, r = e >> 4, c = (e & 15) * 2;
        const int gr = tm * TM + r, gc = kc * 32 + c;
        double2 vr = make_double2(0.0, 0.0), vi = vr;
        if (gr < d && gc < d) {
          vr = *reinterpret_cast<const double2*>(Ag + (size_t)gr * S + gc);
          vi = *reinterpret_cast<const double2*>(Ag + plane + (size_t)gr * S + gc);
        }
        *reinterpret_cast<double2*>(&smAr[r * AS + c]) = vr;
        *reinterpret_cast<double2*>(&smAi[r * AS + c]) = vi;
      }
#pragma unroll
      for (int i = 0; i < WN; i++) {              // B tile: 32 rows x TN columns
        const int e = tid + i * 256, r = e / (TN / 2), c = (e % (TN / 2)) * 2;
        const int gr = kc * 32 + r, gc = tn * TN + c;
        double2 vr = make_double2(0.0, 0.0), vi = vr;
        if (gr < d && gc < d) {
          vr = *reinterpret_cast<const double2*>(Bg + (size_t)gr * S + gc);
          vi = *reinterpret_cast<const double2*>(Bg + plane + (size_t)gr * S + gc);
        }
        *reinterpret_cast<double2*>(&smBr[r * BS + c]) = vr;
        *reinterpret_cast<double2*>(&smBi[r * BS + c]) = vi;
      }
      __syncthreads();
      const int kvalid = d - kc * 32;
      const int kmax = kvalid >= 32 ? 8 : (kvalid + 3) / 4;
      if (rowv[0] && colv[0]) {
        const double* are = smAr + (wm * WM * 8 + gq) * AS + q4;
        const double* aim = smAi + (wm * WM * 8 + gq) * AS + q4;
        const double* bre = smBr + q4 * BS + wn * WN * 8 + gq;
        const double* bim = smBi + q4 * BS + wn * WN * 8 + gq;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          if (ks < kmax) {
            double ar[WM], ai[WM], as[WM];
#pragma unroll
            for (int a = 0; a < WM; a++) { ar[a] = are[a * 8 * AS + ks * 4]; ai[a] = aim[a * 8 * AS + ks * 4]; as[a] = ar[a] + ai[a]; }
#pragma unroll
            for (int b = 0; b < WN; b++) {
              if (colv[b]) {
                const double br = bre[ks * 4 * BS + b * 8], bi = bim[ks * 4 * BS + b * 8];
                const double bs = br + bi;
#pragma unroll
                for (int a = 0; a < WM; a++) {
                  if (rowv[a]) {
                    dmma(T1[a][b][0], T1[a][b][1], ar[a], br);
                    dmma(T2[a][b][0], T2[a][b][1], ai[a], bi);
                    dmma(T3[a][b][0], T3[a][b][1], as[a], bs);
                  }
                }
              }
            }
          }
        }
      }
    }
  }
  double* Cg = g.C + (g.cinner > 0 ? (long long)(s / g.cinner) * g.cstride2 + (long long)(s % g.cinner) * g.cstride
                                   : (long long)s * g.cstride);
#pragma unroll
  for (int a = 0; a < WM; a++) {
    const int row = tm * TM + (wm * WM + a) * 8 + gq;
#pragma unroll
    for (int b = 0; b < WN; b++) {
      const int col = tn * TN + (wn * WN + b) * 8 + 2 * q4;
      if (row < d && col < d) {
        // 3M: Re = T1 - T2, Im = T3 - T1 - T2
        double r0 = g.alpha * (T1[a][b][0] - T2[a][b][0]), r1 = g.alpha * (T1[a][b][1] - T2[a][b][1]);
        double i0 = g.alpha * ((T3[a][b][0] - T1[a][b][0]) - T2[a][b][0]), i1 = g.alpha * ((T3[a][b][1] - T1[a][b][1]) - T2[a][b][1]);
        const size_t o = (size_t)row * S + col;
        for (int q = 0; q < g.nadd; q++) {
          const double* Dg = g.D[q].p + g.D[q].off(s);
          const double2 u = *reinterpret_cast<const double2*>(Dg + o), v = *reinterpret_cast<const double2*>(Dg + plane + o);
          r0 = fma(g.beta[q], u.x, r0); r1 = fma(g.beta[q], u.y, r1);
          i0 = fma(g.beta[q], v.x, i0); i1 = fma(g.beta[q], v.y, i1);
        }
        if (row == col) r0 += g.gamma;
        if (row == col + 1) r1 += g.gamma;
        if (col + 1 >= d) { r1 = 0.0; i1 = 0.0; }
        *reinterpret_cast<double2*>(Cg + o) = make_double2(r0, r1);
        *reinterpret_cast<double2*>(Cg + plane + o) = make_double2(i0, i1);
      }
    }
  }
}

// host-side launch helper.  Measured on B200 (synthetic d = 64 / 128 / 256, cavity d = 80): 64 x 64 tiles (WM, WN = 2, 4;
// 168 registers, one CTA per SM) are 10-50 % SLOWER than 32 x 32 (80 registers, three CTAs per SM) because the operand
// staging is synchronous and only co-resident CTAs hide it; 64 x 32 at two CTAs per SM is a wash.  Larger tiles need an
// asynchronous (cp.async / TMA) double buffer first -- DESIGN.md section 7.
// (A register double buffer -- global loads of chunk i+1 issued behind the hand-over of chunk i -- was measured too: 5-8 % SLOWER
// at every tile shape, 90 instead of 80 registers; the kernel is not bound by the exposed load latency.)
static inline void g_gemm_launch(const GGemm& g, int nb, cudaStream_t st) {
  constexpr int WM = 1, WN = 2;
  const size_t smem = (size_t)(2 * 32 * WM * 36 + 2 * 32 * (16 * WN + 4)) * 8;
  const int tiles = (g.d + 31) / 32;
  g_gemm_kernel<WM, WN><<<dim3(tiles, tiles, nb), 256, smem, st>>>(g);
}

// X[s] = (A0 + sum_j u[s][j] A_j) * scale ;  optionally the unscaled generator too (Taylor mode)
__global__ void g_build_kernel(int d, int S, int nc, const double* A0p, const double* Ap, const double* u /* [s][nc] */,
                               double scale, double* X, double* Xun, long long xstride) {
  const int s = blockIdx.x;
  const int n2 = d * S;  // double2 per slot
  double uj[8];
  for (int j = 0; j < nc && j < 8; j++) uj[j] = u[(size_t)s * nc + j];
  double2* out = reinterpret_cast<double2*>(X + (long long)s * xstride);
  double2* out2 = Xun ? reinterpret_cast<double2*>(Xun + (long long)s * xstride) : nullptr;
  for (int e = threadIdx.x; e < n2; e += blockDim.x) {
    double2 v = reinterpret_cast<const double2*>(A0p)[e];
    for (int j = 0; j < nc; j++) {
      const double2 w = reinterpret_cast<const double2*>(Ap + (size_t)j * 2 * n2)[e];
      v.x = fma(uj[j], w.x, v.x);
      v.y = fma(uj[j], w.y, v.y);
    }
    if (out2) out2[e] = v;
    out[e] = make_double2(v.x * scale, v.y * scale);
  }
}

// max_k |u[j][k]| per control (device-resident u): out[j]
__global__ void g_umax_kernel(const double* u, int nc, long long n, double* out) {
  __shared__ double red[8][32];
  double m[8];
  for (int j = 0; j < 8; j++) m[j] = 0.0;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
    for (int j = 0; j < nc && j < 8; j++) m[j] = fmax(m[j], fabs(u[k * nc + j]));
  for (int j = 0; j < nc && j < 8; j++) {
    double v = m[j];
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    if ((threadIdx.x & 31) == 0) red[j][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < nc && threadIdx.x < 8) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) v = fmax(v, red[threadIdx.x][w]);
    // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(out + threadIdx.x), (unsigned long long)__double_as_longlong(v));
  }
}

// In-place Gauss-Jordan inverse with partial (row) pivoting, one CTA per slice, matrix in global memory (L2).
__global__ void __launch_bounds__(256) g_inverse_kernel(int d, int S, double* N, long long stride, int* status) {
  extern __shared__ __align__(16) unsigned char gsm[];
  double2* rowbuf = reinterpret_cast<double2*>(gsm);   // d
  double2* colbuf = rowbuf + d;                        // d
  int* piv = reinterpret_cast<int*>(colbuf + d);       // d
  __shared__ double red_v[8];
  __shared__ int red_i[8];
  __shared__ int sh_p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* re = N + (long long)blockIdx.x * stride;
  double* im = re + (size_t)d * S;
  bool ok = true;
  for (int k = 0; k < d; k++) {
    // pivot search over rows i >= k
    double best = -1.0;
    int bi = k;
    for (int i = k + tid; i < d; i += 256) {
      const double x = re[(size_t)i * S + k], y = im[(size_t)i * S + k];
      const double m = x * x + y * y;
      if (m > best) { best = m; bi = i; }
    }
    for (int off = 16; off > 0; off >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { red_v[warp] = best; red_i[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      double b = red_v[0];
      int p = red_i[0];
      for (int w = 1; w < 8; w++)
        if (red_v[w] > b || (red_v[w] == b && red_i[w] < p)) { b = red_v[w]; p = red_i[w]; }
      sh_p = (b > 0.0) ? p : -1 - p;
      piv[k] = p;
    }
    __syncthreads();
    int p = sh_p;
    if (p < 0) { ok = false; p = -1 - p; }
    // stage the pivot row (old row p), the displaced row goes to row p; multipliers = column k after the swap
    for (int c = tid; c < d; c += 256) {
      const double2 rp = make_double2(re[(size_t)p * S + c], im[(size_t)p * S + c]);
      if (p != k) {
        re[(size_t)p * S + c] = re[(size_t)k * S + c];
        im[(size_t)p * S + c] = im[(size_t)k * S + c];
      }
      rowbuf[c] = rp;
    }
    __syncthreads();
    for (int i = tid; i < d; i += 256) colbuf[i] = make_double2(re[(size_t)i * S + k], im[(size_t)i * S + k]);  // row k entry unused
    __syncthreads();
    const double2 pv = rowbuf[k];
    const double den = 1.0 / (pv.x * pv.x + pv.y * pv.y);
    const double pir = pv.x * den, pii = -pv.y * den;
    for (int e = tid; e < d * d; e += 256) {
      const int i = e / d, c = e - i * d;
      double vr, vi;
      if (i == k) {
        if (c == k) { vr = pir; vi = pii; }
        else { const double2 r = rowbuf[c]; vr = r.x * pir - r.y * pii; vi = r.x * pii + r.y * pir; }
      } else {
        const double2 f = colbuf[i];
        const double gr = f.x * pir - f.y * pii, gi = f.x * pii + f.y * pir;  // f / pivot
        if (c == k) { vr = -gr; vi = -gi; }
        else {
          const double2 r = rowbuf[c];
          vr = re[(size_t)i * S + c] - (gr * r.x - gi * r.y);
          vi = im[(size_t)i * S + c] - (gr * r.y + gi * r.x);
        }
      }
      re[(size_t)i * S + c] = vr;
      im[(size_t)i * S + c] = vi;
    }
    __syncthreads();
  }
  // undo the row interchanges as column interchanges in reverse order
  for (int k = d - 1; k >= 0; k--) {
    const int p = piv[k];
    if (p != k) {
      for (int i = tid; i < d; i += 256) {
        const size_t a = (size_t)i * S + k, b = (size_t)i * S + p;
        const double tr = re[a], ti = im[a];
        re[a] = re[b]; im[a] = im[b];
        re[b] = tr; im[b] = ti;
      }
    }
    __syncthreads();
  }
  if (tid == 0 && !ok) atomicExch(status, 8);
}

struct GSweep {
  int d, S, m, nc, nt, cost, n;
  int phase;          // 0: forward + cost + backward; 1: forward (+ J) only; 2: backward only (lam_final given)
  int want_grad, store_costates;
  const double* U;    // [b*nt + k] planar slots
  const double* L;    // [(b*nt + k)*nc + j]
  long long slot;     // doubles per slot
  const double* x0;   // c128 col-major d x m (shared)  or per-pulse when x_start_ext
  const double* x_start_ext;
  const double* T;
  const double* lam_final;
  double* X;          // [(b*(nt+1) + k)] d x m c128
  double* LAM;
  double* x_final;
  double* lam_start;
  double* J;
  double* dJdu;
  unsigned long long row_mask_lo;  // penalty rows < 64 (larger d: rows list not supported on this path yet)
  unsigned col_mask;
  double mu;
};

// y (d x m, shared, [r*m + c] re/im) = op(U) x ;  8 warps: warp w takes rows/cols w, w+8, ...; lanes split the sum index
template <bool ADJ>
__device__ __forceinline__ void g_matvec(const double* Ure, const double* Uim, int d, int S, int m, const double* xr,
                                         const double* xi, double* yr, double* yi, int warp, int lane) {
  for (int r = warp; r < d; r += 8) {
    double ar[8], ai[8];
#pragma unroll
    for (int c = 0; c < 8; c++) { ar[c] = 0.0; ai[c] = 0.0; }
    for (int k = lane; k < d; k += 32) {
      double qr, qi;
      if (ADJ) { qr = Ure[(size_t)k * S + r]; qi = -Uim[(size_t)k * S + r]; }
      else     { qr = Ure[(size_t)r * S + k]; qi = Uim[(size_t)r * S + k]; }
#pragma unroll
      for (int c = 0; c < 8; c++)
        if (c < m) {
          const double vr = xr[k * m + c], vi = xi[k * m + c];
          ar[c] = fma(qr, vr, fma(-qi, vi, ar[c]));
          ai[c] = fma(qr, vi, fma(qi, vr, ai[c]));
        }
    }
#pragma unroll
    for (int c = 0; c < 8; c++)
      if (c < m) {
        double a = ar[c], b = ai[c];
        for (int off = 16; off > 0; off >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, off); b += __shfl_xor_sync(0xffffffffu, b, off); }
        if (lane == 0) { yr[r * m + c] = a; yi[r * m + c] = b; }
      }
  }
}

// One CTA (256 threads) per pulse: the serial loops of the reference, with the matrices streamed from HBM/L2.
__global__ void __launch_bounds__(256) g_sweep_kernel(GSweep g) {
  extern __shared__ __align__(16) unsigned char gsm[];
  const int d = g.d, m = g.m, dm = d * m, nt = g.nt;
  double* buf = reinterpret_cast<double*>(gsm);   // 3 state buffers x (re, im): x / lambda ping-pong + y
  double* b0 = buf; double* b1 = buf + 2 * dm; double* b2 = buf + 4 * dm;
  __shared__ double red[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const size_t plane = (size_t)d * g.S;
  const bool pen = g.col_mask != 0u && g.row_mask_lo != 0ull;
  auto load = [&](double* dst, const double* src) {
    for (int e = tid; e < dm; e += 256) { const int c = e / d, r = e - c * d; dst[r * m + c] = src[2 * e]; dst[dm + r * m + c] = src[2 * e + 1]; }
  };
  auto store = [&](double* dst, const double* src) {
    for (int e = tid; e < dm; e += 256) { const int c = e / d, r = e - c * d; dst[2 * e] = src[r * m + c]; dst[2 * e + 1] = src[dm + r * m + c]; }
  };
  auto penalised = [&](int r, int c) { return pen && r < 64 && ((g.row_mask_lo >> r) & 1ull) && ((g.col_mask >> c) & 1u); };
  double Jpen = 0.0;
  double* cur = b0; double* nxt = b1;
  if (g.phase != 2) {
    load(cur, g.x_start_ext ? g.x_start_ext + (size_t)b * 2 * dm : g.x0);
    __syncthreads();
    for (int k = 0; k < nt; k++) {
      store(g.X + ((size_t)b * (nt + 1) + k) * 2 * dm, cur);
      if (pen) for (int e = tid; e < dm; e += 256) { const int r = e / m, c = e - r * m; if (penalised(r, c)) Jpen += cur[e] * cur[e] + cur[dm + e] * cur[dm + e]; }
      const double* Uk = g.U + ((size_t)b * nt + k) * g.slot;
      g_matvec<false>(Uk, Uk + plane, d, g.S, m, cur, cur + dm, nxt, nxt + dm, warp, lane);
      __syncthreads();
      double* t = cur; cur = nxt; nxt = t;
    }
    store(g.X + ((size_t)b * (nt + 1) + nt) * 2 * dm, cur);
    if (g.x_final) store(g.x_final + (size_t)b * 2 * dm, cur);
    if (pen) for (int e = tid; e < dm; e += 256) { const int r = e / m, c = e - r * m; if (penalised(r, c)) Jpen += cur[e] * cur[e] + cur[dm + e] * cur[dm + e]; }
  } else {
    load(cur, g.X + ((size_t)b * (nt + 1) + nt) * 2 * dm);
  }
  __syncthreads();
  // ---- terminal cost / costate: x_N in cur, lambda_N -> nxt ----
  const bool builtin = (g.phase != 2) && g.cost != 2;
  if (tid < 4) red[tid] = 0.0;
  __syncthreads();
  double J = 0.0, cr_ = 0.0, ci_ = 0.0;
  if (builtin) {
    double orr = 0.0, oii = 0.0;
    for (int e = tid; e < dm; e += 256) {
      const int c = e / d, r = e - c * d;
      const double tr = g.T[2 * e], ti = g.T[2 * e + 1];
      const double xr = cur[r * m + c], xi = cur[dm + r * m + c];
      orr += tr * xr + ti * xi;
      oii += tr * xi - ti * xr;
    }
    for (int off = 16; off > 0; off >>= 1) { orr += __shfl_xor_sync(0xffffffffu, orr, off); oii += __shfl_xor_sync(0xffffffffu, oii, off); }
    if (lane == 0) { atomicAdd(&red[0], orr); atomicAdd(&red[1], oii); }
  }
  if (pen && g.phase != 2) {
    double ps = Jpen;
    for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
    if (lane == 0) atomicAdd(&red[2], ps);
  }
  __syncthreads();
  if (builtin) {
    const double Or = red[0], Oi = red[1], nn = (double)g.n * (double)g.n;
    if (g.cost == 0) { J = 1.0 - (Or * Or + Oi * Oi) / nn; cr_ = -2.0 * Or / nn; ci_ = -2.0 * Oi / nn; }
    else { const double a = sqrt(Or * Or + Oi * Oi); J = 1.0 - a; cr_ = -Or / a; ci_ = -Oi / a; }
  }
  if (pen && g.phase != 2) J += g.mu * red[2];
  if (tid == 0 && g.J && (builtin || (pen && g.phase != 2))) g.J[b] = J;
  if (g.phase == 1 || !g.want_grad) return;
  for (int e = tid; e < dm; e += 256) {
    const int c = e / d, r = e - c * d;
    double lr = 0.0, li = 0.0;
    if (g.lam_final) { lr = g.lam_final[(size_t)b * 2 * dm + 2 * e]; li = g.lam_final[(size_t)b * 2 * dm + 2 * e + 1]; }
    else if (builtin) { const double tr = g.T[2 * e], ti = g.T[2 * e + 1]; lr = cr_ * tr - ci_ * ti; li = cr_ * ti + ci_ * tr; }
    if (penalised(r, c)) { lr = fma(2.0 * g.mu, cur[r * m + c], lr); li = fma(2.0 * g.mu, cur[dm + r * m + c], li); }
    nxt[r * m + c] = lr; nxt[dm + r * m + c] = li;
  }
  __syncthreads();
  // ---- backward sweep with the gradient contraction: lam in `lam`, x_k reloaded from the stored states ----
  double* lam = nxt; double* lam2 = cur; double* xk = b2;
  // (cur / nxt roles are free now; b2 holds x_k, y goes to lam2 temporarily before the costate update)
  if (g.store_costates && g.LAM) store(g.LAM + ((size_t)b * (nt + 1) + nt) * 2 * dm, lam);
  for (int k = nt - 1; k >= 0; k--) {
    load(xk, g.X + ((size_t)b * (nt + 1) + k) * 2 * dm);
    __syncthreads();
    for (int j = 0; j < g.nc; j++) {
      const double* Lk = g.L + (((size_t)b * nt + k) * g.nc + j) * g.slot;
      g_matvec<false>(Lk, Lk + plane, d, g.S, m, xk, xk + dm, lam2, lam2 + dm, warp, lane);   // y = dU x_k
      __syncthreads();
      double s = 0.0;
      for (int e = tid; e < dm; e += 256) s += lam[e] * lam2[e] + lam[dm + e] * lam2[dm + e];  // Re(conj(lam) .* y)
      for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (tid < 4) red[tid] = 0.0;
      __syncthreads();
      if (lane == 0) atomicAdd(&red[0], s);
      __syncthreads();
      if (tid == 0) g.dJdu[((size_t)b * nt + k) * g.nc + j] = red[0];
      __syncthreads();
    }
    const double* Uk = g.U + ((size_t)b * nt + k) * g.slot;
    g_matvec<true>(Uk, Uk + plane, d, g.S, m, lam, lam + dm, lam2, lam2 + dm, warp, lane);       // lam_k = U_k' lam_{k+1}
    __syncthreads();
    if (pen) {
      for (int e = tid; e < dm; e += 256) { const int r = e / m, c = e - r * m; if (penalised(r, c)) { lam2[e] = fma(2.0 * g.mu, xk[e], lam2[e]); lam2[dm + e] = fma(2.0 * g.mu, xk[dm + e], lam2[dm + e]); } }
      __syncthreads();
    }
    double* t = lam; lam = lam2; lam2 = t;
    if (g.store_costates && g.LAM) store(g.LAM + ((size_t)b * (nt + 1) + k) * 2 * dm, lam);
  }
  if (g.lam_start) store(g.lam_start + (size_t)b * 2 * dm, lam);
}


// ---------------------------------------------------------------------------------------------------------------------
// General-path sweeps, second generation (no running penalty): the same two-level structure as K1/K2G/K3N.
//   host: segment propagators Q_seg = U_{k1-1} ... U_{k0} by batched g_gemm launches (one level per slice of a segment)
//   gs_scan_kernel      one CTA per pulse: boundary states / costates of every segment, terminal cost
//   gs_seg_kernel       one CTA per segment: x_k (forward) and lambda_k (backward) of its slices, to HBM
//   gs_contract_kernel  one CTA per (slice, control): Re tr(lambda' dU x) = <dU, lambda x'>_F, a streaming pass over dU
// Mat-vec: DMMA.8x8x4 with the operand fragments read straight from HBM/L2 (each warp request covers whole 32-byte
// sectors), the d x m state interleaved in shared memory as in qoc_sweep.cuh (two real DMMAs per complex k-step).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int GS_NW = 8;  // warps per CTA in the scan / segment kernels

// y = op(U) x; U planar in global memory (row stride S), xs / ys interleaved [row][2 m] in shared memory with rows up to
// rows_pad zeroed.  Warp w takes the 8-row output tiles w, w + GS_NW, ...
template <bool ADJ, int NCT>
__device__ __forceinline__ void gs_mv(const double* Ure, const double* Uim, int d, int S, int m, const double* xs, double* ys,
                                      int warp, int lane) {
  const int g = lane >> 2, q = lane & 3, W = 2 * m;
  const int ntile = (d + 7) / 8, nks = (d + 3) / 4;
  constexpr int KB = 8;   // k-steps per batch of loads in flight
  for (int mi = warp; mi < ntile; mi += GS_NW) {
    double p1[2][NCT][2], p2[2][NCT][2];
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int ct = 0; ct < NCT; ct++) { p1[h][ct][0] = p1[h][ct][1] = 0.0; p2[h][ct][0] = p2[h][ct][1] = 0.0; }
    const int orow = mi * 8 + g;                       // output row (non-adj) / operand column (adj) of this thread
    const int oc = orow < d ? orow : d - 1;            // clamped: results of rows >= d are discarded
    for (int k0 = 0; k0 < nks; k0 += KB) {
      double ar[KB], ai[KB];
#pragma unroll
      for (int kk = 0; kk < KB; kk++) {
        int kr = (k0 + kk) * 4 + q;                    // reduction index
        kr = kr < d ? kr : d - 1;                      // clamped: multiplied by the zero rows of x
        const size_t off = ADJ ? (size_t)kr * S + oc : (size_t)oc * S + kr;
        const bool v = (k0 + kk) < nks;
        ar[kk] = v ? __ldg(Ure + off) : 0.0;
        ai[kk] = v ? __ldg(Uim + off) : 0.0;
      }
#pragma unroll
      for (int kk = 0; kk < KB; kk++) {
        const int ks = k0 + kk;
        const int h = kk & 1;
        const int kr = ks * 4 + q;
#pragma unroll
        for (int ct = 0; ct < NCT; ct++) {
          const double bv = (ks < nks && kr < d && ct * 8 + g < W) ? xs[kr * W + ct * 8 + g] : 0.0;
          dmma(p1[h][ct][0], p1[h][ct][1], ar[kk], bv);
          dmma(p2[h][ct][0], p2[h][ct][1], ai[kk], bv);
        }
      }
    }
#pragma unroll
    for (int ct = 0; ct < NCT; ct++) {
      const int c = ct * 4 + q;
      const double u = p1[0][ct][0] + p1[1][ct][0], v = p1[0][ct][1] + p1[1][ct][1];
      const double e = p2[0][ct][0] + p2[1][ct][0], f = p2[0][ct][1] + p2[1][ct][1];
      const double yr = ADJ ? u + f : u - f;
      const double yi = ADJ ? v - e : v + e;
      if (orow < d && c < m) *reinterpret_cast<double2*>(ys + orow * W + 2 * c) = make_double2(yr, yi);
    }
  }
}
template <bool ADJ>
__device__ __forceinline__ void gs_mv_any(const double* U, int d, int S, int m, const double* xs, double* ys, int warp, int lane) {
  const double* Uim = U + (size_t)d * S;
  if (m <= 4) gs_mv<ADJ, 1>(U, Uim, d, S, m, xs, ys, warp, lane);
  else gs_mv<ADJ, 2>(U, Uim, d, S, m, xs, ys, warp, lane);
}

// interleaved (shared) <-> c128 column-major (global)
__device__ __forceinline__ void gs_to_global(double* g, const double* xs, int d, int m, int tid, int nth) {
  for (int c = 0; c < m; c++)
    for (int r = tid; r < d; r += nth)
      reinterpret_cast<double2*>(g)[r + (size_t)d * c] = *reinterpret_cast<const double2*>(xs + r * 2 * m + 2 * c);
}
__device__ __forceinline__ void gs_from_global(double* xs, const double* g, int d, int m, int tid, int nth) {
  for (int c = 0; c < m; c++)
    for (int r = tid; r < d; r += nth)
      *reinterpret_cast<double2*>(xs + r * 2 * m + 2 * c) = reinterpret_cast<const double2*>(g)[r + (size_t)d * c];
}

struct GS {
  int d, S, m, nc, nt, cost, n;
  int spp, L;          // segments per pulse, slices per segment (the last segment may be shorter)
  int mode;            // 0: forward + cost + backward, 1: forward only, 2: backward only (lam_final),
                       // 4: forward from x_start_ext and backward from lam_final, no cost (time sharding)
  int skip_bwd;        // mode 0 without the backward part (propagate with a built-in cost)
  long long slot;
  const double* U;     // [b*nt + k]
  const double* L_;    // [(b*nt + k)*nc + j]
  const double* Q;     // [b*spp + s]
  const double* x0;
  const double* x_start_ext;
  const double* T;
  const double* lam_final;
  double* xs_start;    // [b*spp + s] d x m c128
  double* lam_end;
  double* X;           // [b*(nt+1) + k]
  double* LAM;
  double* x_final;
  double* lam_start;
  double* J;
  double* dJdu;
};

// one CTA per pulse: boundary walk over the segment propagators
__global__ void __launch_bounds__(GS_NW * 32) gs_scan_kernel(GS g) {
  extern __shared__ __align__(16) unsigned char gsm[];
  const int d = g.d, m = g.m, W = 2 * m, dm = d * m;
  const int rows_pad = (d + 7) / 8 * 8;
  double* b0 = reinterpret_cast<double*>(gsm);
  double* b1 = b0 + rows_pad * W;
  __shared__ double red[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nth = blockDim.x;
  const int b = blockIdx.x;
  for (int e = tid; e < 2 * rows_pad * W; e += nth) b0[e] = 0.0;
  __syncthreads();
  double* cur = b0; double* nxt = b1;
  const bool do_fwd = g.mode != 2, do_bwd = g.mode == 2 || g.mode == 4 || (g.mode == 0 && !g.skip_bwd);
  if (do_fwd) {
    gs_from_global(cur, g.x_start_ext ? g.x_start_ext + (size_t)b * 2 * dm : g.x0, d, m, tid, nth);
    __syncthreads();
    for (int s = 0; s < g.spp; s++) {
      gs_to_global(g.xs_start + ((size_t)b * g.spp + s) * 2 * dm, cur, d, m, tid, nth);
      gs_mv_any<false>(g.Q + ((size_t)b * g.spp + s) * g.slot, d, g.S, m, cur, nxt, warp, lane);
      __syncthreads();
      double* t = cur; cur = nxt; nxt = t;
    }
    if (g.x_final) gs_to_global(g.x_final + (size_t)b * 2 * dm, cur, d, m, tid, nth);
    gs_to_global(g.X + ((size_t)b * (g.nt + 1) + g.nt) * 2 * dm, cur, d, m, tid, nth);
  }
  if (g.mode == 1) return;
  const bool builtin = do_fwd && g.cost != 2 && g.mode != 4;
  if (tid < 4) red[tid] = 0.0;
  __syncthreads();
  double cr_ = 0.0, ci_ = 0.0;
  if (builtin) {
    double orr = 0.0, oii = 0.0;
    for (int c = 0; c < m; c++)
      for (int r = tid; r < d; r += nth) {
        const double2 t = reinterpret_cast<const double2*>(g.T)[r + (size_t)d * c];
        const double2 x = *reinterpret_cast<const double2*>(cur + r * W + 2 * c);
        orr += t.x * x.x + t.y * x.y;
        oii += t.x * x.y - t.y * x.x;
      }
    for (int off = 16; off > 0; off >>= 1) { orr += __shfl_xor_sync(0xffffffffu, orr, off); oii += __shfl_xor_sync(0xffffffffu, oii, off); }
    if (lane == 0) { atomicAdd(&red[0], orr); atomicAdd(&red[1], oii); }
    __syncthreads();
    const double Or = red[0], Oi = red[1], nn = (double)g.n * (double)g.n;
    double J;
    if (g.cost == 0) { J = 1.0 - (Or * Or + Oi * Oi) / nn; cr_ = -2.0 * Or / nn; ci_ = -2.0 * Oi / nn; }
    else { const double a = sqrt(Or * Or + Oi * Oi); J = 1.0 - a; cr_ = -Or / a; ci_ = -Oi / a; }
    if (tid == 0 && g.J) g.J[b] = J;
  }
  if (!do_bwd) return;
  for (int c = 0; c < m; c++)
    for (int r = tid; r < d; r += nth) {
      double2 l = make_double2(0.0, 0.0);
      if (g.lam_final) l = reinterpret_cast<const double2*>(g.lam_final + (size_t)b * 2 * dm)[r + (size_t)d * c];
      else if (builtin) {
        const double2 t = reinterpret_cast<const double2*>(g.T)[r + (size_t)d * c];
        l = make_double2(cr_ * t.x - ci_ * t.y, cr_ * t.y + ci_ * t.x);
      }
      *reinterpret_cast<double2*>(nxt + r * W + 2 * c) = l;
    }
  { double* t = cur; cur = nxt; nxt = t; }
  __syncthreads();
  for (int s = g.spp - 1; s >= 0; s--) {
    gs_to_global(g.lam_end + ((size_t)b * g.spp + s) * 2 * dm, cur, d, m, tid, nth);
    gs_mv_any<true>(g.Q + ((size_t)b * g.spp + s) * g.slot, d, g.S, m, cur, nxt, warp, lane);
    __syncthreads();
    double* t = cur; cur = nxt; nxt = t;
  }
  if (g.lam_start) gs_to_global(g.lam_start + (size_t)b * 2 * dm, cur, d, m, tid, nth);
}

// one CTA per segment: states and costates of its slices
__global__ void __launch_bounds__(GS_NW * 32) gs_seg_kernel(GS g, int nseg_total) {
  extern __shared__ __align__(16) unsigned char gsm[];
  const int d = g.d, m = g.m, W = 2 * m, dm = d * m;
  const int rows_pad = (d + 7) / 8 * 8;
  double* b0 = reinterpret_cast<double*>(gsm);
  double* b1 = b0 + rows_pad * W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nth = blockDim.x;
  for (int e = tid; e < 2 * rows_pad * W; e += nth) b0[e] = 0.0;
  __syncthreads();
  const bool do_fwd = g.mode != 2, do_bwd = g.mode == 2 || g.mode == 4 || (g.mode == 0 && !g.skip_bwd);
  for (int seg = blockIdx.x; seg < nseg_total; seg += gridDim.x) {
    const int b = seg / g.spp, si = seg - b * g.spp;
    const int k0 = si * g.L, k1 = (k0 + g.L < g.nt) ? k0 + g.L : g.nt;
    double* cur = b0; double* nxt = b1;
    if (do_fwd) {
      gs_from_global(cur, g.xs_start + (size_t)seg * 2 * dm, d, m, tid, nth);
      __syncthreads();
      for (int k = k0; k < k1; k++) {
        gs_to_global(g.X + ((size_t)b * (g.nt + 1) + k) * 2 * dm, cur, d, m, tid, nth);
        if (k + 1 < k1) {
          gs_mv_any<false>(g.U + ((size_t)b * g.nt + k) * g.slot, d, g.S, m, cur, nxt, warp, lane);
          __syncthreads();
          double* t = cur; cur = nxt; nxt = t;
        }
      }
      __syncthreads();
    }
    if (do_bwd) {
      gs_from_global(cur, g.lam_end + (size_t)seg * 2 * dm, d, m, tid, nth);
      __syncthreads();
      for (int k = k1 - 1; k >= k0; k--) {
        gs_to_global(g.LAM + ((size_t)b * (g.nt + 1) + k + 1) * 2 * dm, cur, d, m, tid, nth);
        gs_mv_any<true>(g.U + ((size_t)b * g.nt + k) * g.slot, d, g.S, m, cur, nxt, warp, lane);
        __syncthreads();
        double* t = cur; cur = nxt; nxt = t;
      }
      if (k0 == 0) gs_to_global(g.LAM + ((size_t)b * (g.nt + 1)) * 2 * dm, cur, d, m, tid, nth);
      __syncthreads();
    }
  }
}

// one CTA per (slice, control): dJdu[j,k] = sum_{r,c} Re( dU[r][c] * w[r][c] ), w[r][c] = sum_l x_k[c][l] conj(lambda_{k+1}[r][l])
// (src/gradient_computations.jl:65-74, :217-223)
__global__ void __launch_bounds__(256) gs_contract_kernel(GS g) {
  extern __shared__ __align__(16) unsigned char gsm[];
  const int d = g.d, S = g.S, m = g.m, W = 2 * m, dm = d * m;
  double* xk = reinterpret_cast<double*>(gsm);   // S rows (zero padded)
  double* lk = xk + (size_t)S * W;               // d rows
  __shared__ double red[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t sl = blockIdx.x;                  // flat slice index b*nt + k
  const int j = blockIdx.y;
  const int b = (int)(sl / g.nt), k = (int)(sl - (size_t)b * g.nt);
  for (int e = tid; e < (S + d) * W; e += 256) xk[e] = 0.0;
  __syncthreads();
  gs_from_global(xk, g.X + ((size_t)b * (g.nt + 1) + k) * 2 * dm, d, m, tid, 256);
  gs_from_global(lk, g.LAM + ((size_t)b * (g.nt + 1) + k + 1) * 2 * dm, d, m, tid, 256);
  __syncthreads();
  const double* Lre = g.L_ + (sl * g.nc + j) * g.slot;
  const double* Lim = Lre + (size_t)d * S;
  const int units = d * S / 2, S2 = S / 2;
  double s = 0.0;
  for (int u0 = tid; u0 < units; u0 += 256 * 4) {
    double2 fr[4], fi[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int u = u0 + i * 256;
      if (u < units) { fr[i] = __ldg(reinterpret_cast<const double2*>(Lre) + u); fi[i] = __ldg(reinterpret_cast<const double2*>(Lim) + u); }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int u = u0 + i * 256;
      if (u < units) {
        const int r = u / S2, c0 = 2 * (u - r * S2);
        const double* lrow = lk + r * W;
        const double* x0p = xk + c0 * W;
        double w0r = 0.0, w0i = 0.0, w1r = 0.0, w1i = 0.0;
        for (int l = 0; l < m; l++) {
          const double2 lam = *reinterpret_cast<const double2*>(lrow + 2 * l);
          const double2 xa = *reinterpret_cast<const double2*>(x0p + 2 * l);
          const double2 xb = *reinterpret_cast<const double2*>(x0p + W + 2 * l);
          w0r = fma(xa.x, lam.x, fma(xa.y, lam.y, w0r)); w0i = fma(xa.y, lam.x, fma(-xa.x, lam.y, w0i));
          w1r = fma(xb.x, lam.x, fma(xb.y, lam.y, w1r)); w1i = fma(xb.y, lam.x, fma(-xb.x, lam.y, w1i));
        }
        s = fma(fr[i].x, w0r, fma(-fi[i].x, w0i, s));
        s = fma(fr[i].y, w1r, fma(-fi[i].y, w1i, s));
      }
    }
  }
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += red[w];
    g.dJdu[(sl * g.nc) + j] = t;
  }
}

// planar slot -> c128 column-major d x d (rank propagator hand-over of the time-sharded general path)
__global__ void planar_to_c128_kernel(const double* P, int d, int S, double* out) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < d * d; e += gridDim.x * blockDim.x) {
    const int c = e / d, r = e - c * d;
    reinterpret_cast<double2*>(out)[e] = make_double2(P[(size_t)r * S + c], P[(size_t)d * S + (size_t)r * S + c]);
  }
}

}  // namespace qoc

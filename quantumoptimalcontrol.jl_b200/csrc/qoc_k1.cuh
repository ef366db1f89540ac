// qoc_k1.cuh -- K1: per-slice generator assembly, scaling-and-squaring Pade expm, expm Jacobian
// (exact block-triangular Frechet derivative or the reference's truncated Taylor series) and the level-1
// propagator scan (running segment product), fused in one persistent kernel.
//
// Replaces, per slice k of the reference:
//   X_k = A0 + sum_j u[j,k] A_j                          src/gradient_computations.jl:18-22
//   U_k = exponential!(X_k, ExpMethodHigham2005())      src/gradient_computations.jl:24   (third-party dep)
//   dU_k/du_j = expm_jacobian!(...)                     src/gradient_computations.jl:67, :177-213
// and produces the segment propagators Q_seg = U_{k1-1} ... U_{k0} that turn the serial sweeps of
// :27-29 and :52-58 into a parallel scan (K2/K3 finish it).
//
// One CTA owns a contiguous run of slices (a segment); every matrix of a slice lives in shared memory as a
// planar slot (qoc_tiles.cuh); every d^3 contraction is a DMMA.8x8x4 tile loop; only U_k, dU_k/du_j and Q_seg
// go to HBM.
#pragma once
#include "qoc_tiles.cuh"

namespace qoc {

struct K1Params {
  int d, nc, nt, batch, order;
  int nseg, seg_per_pulse;
  int want_jac;            // 0: expm only (propagate without gradient)
  const double* A0p;       // planar slot
  const double* Ap;        // nc planar slots
  const double* u;         // nc x nt x batch
  double* U;               // [batch*nt] planar slots
  double* L;               // [(b*nt + k)*nc + j] planar slots
  double* Q;               // [nseg] planar slots
  double* flops;           // accumulated algorithmic flops (F_alg) over slices
  int* status;             // set to QOC_ERR_SINGULAR (8) on a zero pivot
  double theta13;          // scaling threshold: 5.4 (Higham-2005 / reference) or 4.74 (Frechet, Al-Mohy-Higham)
};

// Pade-13 coefficients b0..b13
__constant__ double c_b13[14] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                                 129060195264000.,   10559470521600.,    670442572800.,    33522128640.,
                                 1323241920.,        40840800.,          960960.,          16380.,
                                 182.,               1.};

// slot roles; see the schedule in k1_slice_frechet13
enum : int { sA = 0, sA2, sA4, sA6, sWZ, sW, sN, sP, sR, sM2, sM4, sM6, sT, sLw, sLv, sQ, sL0, K1_BASE_SLOTS = sL0 };

template <class C>
struct K1Ctx {
  Mat s[K1_BASE_SLOTS + 8];  // slot table (pointer-swappable)
  int d, n2;                 // n2 = double2 per slot
  int tid, lane, warp, mi, nj0;
  double2* gjbuf;
  double* scratch;

  // dst = alpha * (sum of up to 3 products) + epilogue terms; one barrier at the end
  template <class F>
  __device__ __forceinline__ void mm1(Mat dst, Mat a0, Mat b0, F f) {
    Acc<C::BN> acc; acc.zero();
    mm_acc<C, false>(acc, a0, b0, mi, nj0, lane);
    mm_store<C>(dst, acc, d, mi, nj0, lane, f);
    __syncthreads();
  }
  template <class F>
  __device__ __forceinline__ void mm2(Mat dst, Mat a0, Mat b0, Mat a1, Mat b1, F f) {
    Acc<C::BN> acc; acc.zero();
    mm_acc<C, false>(acc, a0, b0, mi, nj0, lane);
    mm_acc<C, false>(acc, a1, b1, mi, nj0, lane);
    mm_store<C>(dst, acc, d, mi, nj0, lane, f);
    __syncthreads();
  }
  template <class F>
  __device__ __forceinline__ void mm3(Mat dst, Mat a0, Mat b0, Mat a1, Mat b1, Mat a2, Mat b2, F f) {
    Acc<C::BN> acc; acc.zero();
    mm_acc<C, false>(acc, a0, b0, mi, nj0, lane);
    mm_acc<C, false>(acc, a1, b1, mi, nj0, lane);
    mm_acc<C, false>(acc, a2, b2, mi, nj0, lane);
    mm_store<C>(dst, acc, d, mi, nj0, lane, f);
    __syncthreads();
  }
  template <class F>
  __device__ __forceinline__ void mm4(Mat dst, Mat a0, Mat b0, Mat a1, Mat b1, Mat a2, Mat b2, Mat a3, Mat b3, F f) {
    Acc<C::BN> acc; acc.zero();
    mm_acc<C, false>(acc, a0, b0, mi, nj0, lane);
    mm_acc<C, false>(acc, a1, b1, mi, nj0, lane);
    mm_acc<C, false>(acc, a2, b2, mi, nj0, lane);
    mm_acc<C, false>(acc, a3, b3, mi, nj0, lane);
    mm_store<C>(dst, acc, d, mi, nj0, lane, f);
    __syncthreads();
  }
  __device__ __forceinline__ void lc(Mat dst, double c1, Mat m1, double c2, Mat m2, double c3, Mat m3, double cI) {
    lincomb<C::S>(dst, d, c1, m1, c2, m2, c3, m3, cI, tid, C::NTHREADS);
  }
  __device__ __forceinline__ LinEpi<C::S> epi(double alpha, double c1, Mat m1, double c2, Mat m2, double c3, Mat m3,
                                             double cI) {
    LinEpi<C::S> e;
    e.alpha = alpha; e.c1 = c1; e.c2 = c2; e.c3 = c3; e.cI = cI; e.m1 = m1; e.m2 = m2; e.m3 = m3;
    return e;
  }
  __device__ __forceinline__ void swap(int a, int b) { Mat t = s[a]; s[a] = s[b]; s[b] = t; }
};

// R = exp(A) by the [13/13] Pade approximant; on entry s[sA] holds the (already scaled) generator.
// On exit: s[sR] = r13(A), s[sN] = (V-U)^{-1}, and A2, A4, A6, W are kept for the Frechet part.
template <class C>
__device__ __forceinline__ bool pade13_expm(K1Ctx<C>& c) {
  const double* b = c_b13;
  Mat A = c.s[sA], A2 = c.s[sA2], A4 = c.s[sA4], A6 = c.s[sA6], WZ = c.s[sWZ], W = c.s[sW], N = c.s[sN], P = c.s[sP];
  c.mm1(A2, A, A, NoEpi());
  c.mm1(A4, A2, A2, NoEpi());
  c.mm1(A6, A2, A4, NoEpi());
  c.lc(WZ, b[13], A6, b[11], A4, b[9], A2, 0.0);  // W1
  __syncthreads();
  c.mm1(W, A6, WZ, c.epi(1.0, b[7], A6, b[5], A4, b[3], A2, b[1]));
  c.lc(WZ, b[12], A6, b[10], A4, b[8], A2, 0.0);  // Z1
  __syncthreads();
  // V -> sN and U -> sP are independent: one barrier
  {
    Acc<C::BN> acc; acc.zero();
    mm_acc<C, false>(acc, A6, WZ, c.mi, c.nj0, c.lane);
    mm_store<C>(N, acc, c.d, c.mi, c.nj0, c.lane, c.epi(1.0, b[6], A6, b[4], A4, b[2], A2, b[0]));
    acc.zero();
    mm_acc<C, false>(acc, A, W, c.mi, c.nj0, c.lane);
    mm_store<C>(P, acc, c.d, c.mi, c.nj0, c.lane, NoEpi());
    __syncthreads();
  }
  // (N, P) <- (V - U, V + U)
  diff_sum_inplace<C::S>(N, P, c.d, c.tid, C::NTHREADS);
  __syncthreads();
  bool ok = gj_inverse<C>(N, c.d, c.gjbuf, c.tid);
  c.mm1(c.s[sR], N, P, NoEpi());
  return ok;
}

// Exact Frechet derivative L(A, E) of the same Pade approximant (Al-Mohy & Higham 2009, Alg. 6.4), i.e. the
// (1,2) block of r13([[A,E],[0,A]]) evaluated with the block-triangular structure made explicit: every
// product of the augmented matrix costs the shared A-product (already done in pade13_expm) plus two d x d
// products.  On entry s[sP] holds E (scaled like A); result -> s[sL0 + j].
template <class C>
__device__ __forceinline__ void pade13_frechet(K1Ctx<C>& c, int j) {
  const double* b = c_b13;
  Mat A = c.s[sA], A2 = c.s[sA2], A4 = c.s[sA4], A6 = c.s[sA6], WZ = c.s[sWZ], W = c.s[sW], Ninv = c.s[sN],
      E = c.s[sP], R = c.s[sR], M2 = c.s[sM2], M4 = c.s[sM4], M6 = c.s[sM6], T = c.s[sT], Lw = c.s[sLw],
      Lv = c.s[sLv];
  c.mm2(M2, A, E, E, A, NoEpi());
  c.mm2(M4, A2, M2, M2, A2, NoEpi());
  c.mm2(M6, A4, M2, M4, A2, NoEpi());
  c.lc(T, b[13], M6, b[11], M4, b[9], M2, 0.0);    // Lw1
  c.lc(WZ, b[13], A6, b[11], A4, b[9], A2, 0.0);   // W1
  __syncthreads();
  c.mm2(Lw, A6, T, M6, WZ, c.epi(1.0, b[7], M6, b[5], M4, b[3], M2, 0.0));
  c.lc(T, b[12], M6, b[10], M4, b[8], M2, 0.0);    // Lz1
  c.lc(WZ, b[12], A6, b[10], A4, b[8], A2, 0.0);   // Z1
  __syncthreads();
  // Lv -> sLv and Lu -> sM2' are independent once Lw is known, but Lu's output slot must not alias inputs:
  // Lv first (frees M2/M4/M6/T afterwards), then Lu into T.
  c.mm2(Lv, A6, T, M6, WZ, c.epi(1.0, b[6], M6, b[4], M4, b[2], M2, 0.0));
  c.mm2(T, A, Lw, E, W, NoEpi());  // Lu
  // (T, Lv) <- (Lu - Lv, Lu + Lv)
  diff_sum_inplace<C::S>(T, Lv, c.d, c.tid, C::NTHREADS);
  __syncthreads();
  // rhs = (Lu + Lv) + (Lu - Lv) R  -> M4 ;  L = Ninv rhs
  c.mm1(M4, T, R, c.epi(1.0, 1.0, Lv, 0.0, Lv, 0.0, Lv, 0.0));
  c.mm1(c.s[sL0 + j], Ninv, M4, NoEpi());
}

// The reference's truncated Taylor series (src/gradient_computations.jl:177-213) with dt = 1.
// X (unscaled generator) in s[sM6]; A_j in s[sP]; result -> s[sL0 + j].  Uses sM2, sM4, sT, sLw as scratch.
template <class C>
__device__ __forceinline__ void taylor_jacobian(K1Ctx<C>& c, int j, int order) {
  Mat X = c.s[sM6], Aj = c.s[sP], AjX = c.s[sM2], XAj = c.s[sM4], X2 = c.s[sT], out = c.s[sL0 + j];
  if (order <= 1) {
    slot_copy(out.re, Aj.re, c.n2, c.tid, C::NTHREADS);
    __syncthreads();
    return;
  }
  {
    Acc<C::BN> acc; acc.zero();
    mm_acc<C, false>(acc, Aj, X, c.mi, c.nj0, c.lane);
    mm_store<C>(AjX, acc, c.d, c.mi, c.nj0, c.lane, NoEpi());
    acc.zero();
    mm_acc<C, false>(acc, X, Aj, c.mi, c.nj0, c.lane);
    mm_store<C>(XAj, acc, c.d, c.mi, c.nj0, c.lane, NoEpi());
    if (order >= 4) {
      acc.zero();
      mm_acc<C, false>(acc, X, X, c.mi, c.nj0, c.lane);
      mm_store<C>(X2, acc, c.d, c.mi, c.nj0, c.lane, NoEpi());
    }
    __syncthreads();
  }
  if (order == 2) {
    c.lc(out, 1.0, Aj, 0.5, AjX, 0.5, XAj, 0.0);
    __syncthreads();
    return;
  }
  // order >= 3:  out = Aj + (AjX + XAj)/2 + (AjX X + XAj X + X XAj)/6
  if (order == 3) {
    c.mm3(out, AjX, X, XAj, X, X, XAj, c.epi(1.0 / 6.0, 1.0, Aj, 0.5, AjX, 0.5, XAj, 0.0));
    return;
  }
  // order 4: + (AjX X2 + XAj X2 + X2 AjX + X2 XAj)/24 ; two passes through the accumulator
  c.mm3(c.s[sLw], AjX, X, XAj, X, X, XAj, c.epi(1.0 / 6.0, 1.0, Aj, 0.5, AjX, 0.5, XAj, 0.0));
  c.mm4(out, AjX, X2, XAj, X2, X2, AjX, X2, XAj, c.epi(1.0 / 24.0, 1.0, c.s[sLw], 0.0, Aj, 0.0, Aj, 0.0));
}

template <class C>
__global__ void __launch_bounds__(C::NTHREADS, 1) k1_kernel(K1Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int S = C::S;
  const int d = p.d, nc = p.nc;
  const int slot_d = 2 * d * S;  // doubles per slot
  const int nslots = K1_BASE_SLOTS + nc;
  double* base = reinterpret_cast<double*>(smem_raw);

  K1Ctx<C> c;
  c.d = d;
  c.n2 = slot_d / 2;
  c.tid = threadIdx.x;
  c.lane = threadIdx.x & 31;
  c.warp = threadIdx.x >> 5;
  c.mi = c.warp / (C::NT / C::BN);
  c.nj0 = (c.warp % (C::NT / C::BN)) * C::BN;
  for (int i = 0; i < nslots; i++) { c.s[i].re = base + (size_t)i * slot_d; c.s[i].im = c.s[i].re + d * S; }
  // tail: 8 rows of zero padding (fragment loads of the last tile row run past the last slot), then buffers
  double* tail = base + (size_t)nslots * slot_d;
  c.gjbuf = reinterpret_cast<double2*>(tail + 8 * S);
  c.scratch = reinterpret_cast<double*>(c.gjbuf + 4 * d + 32);
  // zero everything once: pad columns must be exactly zero and all pad reads finite
  {
    const int total2 = (nslots * slot_d + 8 * S) / 2;
    double2* z = reinterpret_cast<double2*>(base);
    for (int e = c.tid; e < total2; e += C::NTHREADS) z[e] = make_double2(0.0, 0.0);
  }
  __syncthreads();

  double my_flops = 0.0;
  bool all_ok = true;
  const double M = 8.0 * d * d * (double)d;

  for (int seg = blockIdx.x; seg < p.nseg; seg += gridDim.x) {
    const int b = seg / p.seg_per_pulse, si = seg - b * p.seg_per_pulse;
    const int k0 = (int)(((long long)si * p.nt) / p.seg_per_pulse);
    const int k1 = (int)(((long long)(si + 1) * p.nt) / p.seg_per_pulse);
    for (int k = k0; k < k1; k++) {
      const size_t slice = (size_t)b * p.nt + k;
      const double* uk = p.u + slice * nc;
      // ---- S1: X = A0 + sum_j u_j A_j  (planar, coalesced from L2) ----
      {
        double2* x2 = reinterpret_cast<double2*>(c.s[sA].re);
        const double2* a0 = reinterpret_cast<const double2*>(p.A0p);
        for (int e = c.tid; e < c.n2; e += C::NTHREADS) {
          double2 v = a0[e];
          for (int j = 0; j < nc; j++) {
            double2 w = reinterpret_cast<const double2*>(p.Ap + (size_t)j * slot_d)[e];
            const double uj = uk[j];
            v.x = fma(uj, w.x, v.x);
            v.y = fma(uj, w.y, v.y);
          }
          x2[e] = v;
        }
      }
      __syncthreads();
      const double nrm = norm1<S>(c.s[sA], d, c.scratch, c.tid, C::NTHREADS);
      int sq = 0;
      if (nrm > p.theta13) {
        sq = (int)ceil(log2(nrm / p.theta13));
        if (sq < 0) sq = 0;
        if (sq > 60) sq = 60;
      }
      const double sc = ldexp(1.0, -sq);
      const bool taylor = (p.order != 0);
      if (taylor && p.want_jac && p.order >= 2) {
        // keep the unscaled generator for the reference's Taylor series
        slot_copy(c.s[sM6].re, c.s[sA].re, c.n2, c.tid, C::NTHREADS);
      }
      if (sq > 0) {
        double2* x2 = reinterpret_cast<double2*>(c.s[sA].re);
        for (int e = c.tid; e < c.n2; e += C::NTHREADS) { double2 v = x2[e]; x2[e] = make_double2(v.x * sc, v.y * sc); }
      }
      __syncthreads();

      // ---- S2: U_k = r13(X / 2^s)^(2^s) ----
      all_ok &= pade13_expm<C>(c);

      // ---- S3: Jacobians ----
      if (p.want_jac) {
        if (!taylor) {
          for (int j = 0; j < nc; j++) {
            slot_copy_scaled(c.s[sP].re, p.Ap + (size_t)j * slot_d, sc, c.n2, c.tid, C::NTHREADS);
            __syncthreads();
            pade13_frechet<C>(c, j);
          }
          // squaring phase: L <- R L + L R ; R <- R R
          for (int t = 0; t < sq; t++) {
            for (int j = 0; j < nc; j++) {
              c.mm2(c.s[sT], c.s[sR], c.s[sL0 + j], c.s[sL0 + j], c.s[sR], NoEpi());
              c.swap(sT, sL0 + j);
            }
            c.mm1(c.s[sT], c.s[sR], c.s[sR], NoEpi());
            c.swap(sT, sR);
          }
        } else {
          for (int t = 0; t < sq; t++) {
            c.mm1(c.s[sT], c.s[sR], c.s[sR], NoEpi());
            c.swap(sT, sR);
          }
          for (int j = 0; j < nc; j++) {
            slot_copy(c.s[sP].re, p.Ap + (size_t)j * slot_d, c.n2, c.tid, C::NTHREADS);
            __syncthreads();
            taylor_jacobian<C>(c, j, p.order);
          }
        }
      } else {
        for (int t = 0; t < sq; t++) {
          c.mm1(c.s[sT], c.s[sR], c.s[sR], NoEpi());
          c.swap(sT, sR);
        }
      }

      // ---- store U_k and dU_k/du_j (whole slots, coalesced; pads are zero) ----
      slot_copy(p.U + slice * slot_d, c.s[sR].re, c.n2, c.tid, C::NTHREADS);
      if (p.want_jac)
        for (int j = 0; j < nc; j++)
          slot_copy(p.L + (slice * nc + j) * slot_d, c.s[sL0 + j].re, c.n2, c.tid, C::NTHREADS);

      // ---- level-1 scan: Q <- U_k Q ----
      if (k == k0) {
        slot_copy(c.s[sQ].re, c.s[sR].re, c.n2, c.tid, C::NTHREADS);
        __syncthreads();
      } else {
        c.mm1(c.s[sT], c.s[sR], c.s[sQ], NoEpi());
        c.swap(sT, sQ);
      }

      // F_alg bookkeeping: pi = 6 for q = 13
      if (c.tid == 0) {
        double G = 0.0;
        if (p.want_jac) G = taylor ? (p.order == 1 ? 0.0 : p.order == 2 ? 2.0 : p.order == 3 ? 5.0 : 10.0)
                                   : (2.0 * 6 + 2.0 * sq + 2.0);
        my_flops += M * ((6.0 + sq + 4.0 / 3.0) + nc * G);
      }
    }
    slot_copy(p.Q + (size_t)seg * slot_d, c.s[sQ].re, c.n2, c.tid, C::NTHREADS);
    __syncthreads();
  }
  if (c.tid == 0) {
    if (my_flops != 0.0) atomicAdd(p.flops, my_flops);
    if (!all_ok) atomicExch(p.status, 8);
  }
}

}  // namespace qoc

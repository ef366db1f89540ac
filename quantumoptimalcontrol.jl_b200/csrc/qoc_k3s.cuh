// qoc_k3s.cuh -- K3S: the small-dimension form of the per-segment sweeps + gradient contraction (d <= 9, m <= 4, nc <= 4).
//
// Same work as k3n_kernel (qoc_sweep.cuh) on the same buffers:
//   forward   x_{k+1} = U_k x_k                 from x_start(seg)      src/gradient_computations.jl:27-29
//   backward  lambda_k = U_k' lambda_{k+1}      from lambda_end(seg)   :52-58
//   gradient  dJ/du[j,k] = sum_l Re( lambda_{k+1}[:,l]' dU_k/du_j x_k[:,l] )   :70-73, :217-223
//
// k3n_kernel runs one CTA per segment: every recurrence step is a 16-row DMMA tile for 9 rows behind a CTA-wide named
// barrier and a TMA ring, a latency chain of ~0.5 us per step with two segments resident per SM (ncu round 1: 1.45 TB/s,
// 22 % of the HBM roofline, 61 % of its shared-memory wavefronts bank-conflicted).  At d = 9 a recurrence step is only
// 36 complex outputs: here NINE LANES own a segment (lane = row), three segments per warp, 8..16 warps per SM, no CTA
// barrier anywhere.  Each lane loads ITS row of U_k / dU_k/du_j (forward, contraction) or ITS column of U_k (backward)
// straight from HBM/L2 into registers -- the nine rows of a planar slot are contiguous, so the group's request is fully
// coalesced -- and the state / costate is exchanged through 576 bytes of group-private shared memory.  The forward states go
// to the X array (where the getters expect them anyway) and come back one slice at a time in the backward pass.
// The gradient uses the weights w[r][c] = sum_l conj(lambda[r][l]) x[c][l], formed once per slice and contracted with every
// control's Jacobian row (2 DFMA per element).
#pragma once
#include "qoc_k23.cuh"

namespace qoc {

#ifndef K3S_MINB
#define K3S_MINB 1   // (246 registers, eight warps per SM; capping at 128 / 80 registers spills: 1.17 / 1.45 ms instead of 0.86)
#endif
constexpr int K3S_WPB = 8;     // warps per CTA
constexpr int K3S_M = 4;       // state columns (compile-time bound)
constexpr int K3S_D = 9;
__host__ __device__ constexpr size_t k3s_smem_bytes() { return (size_t)K3S_WPB * 3 * 4 * K3S_D * K3S_M * 16; }

struct K3SLane {
  int d, m, S, r, g9;
  bool rowok;
};

// y[c] = sum_k a[k] * xs[k][c]   (xs: group-private shared, [k][K3S_M])
__device__ __forceinline__ void k3s_row_times(const double2 (&a)[K3S_D], const double2* xs, double2 (&y)[K3S_M]) {
#pragma unroll
  for (int c = 0; c < K3S_M; c++) y[c] = make_double2(0.0, 0.0);
#pragma unroll
  for (int k = 0; k < K3S_D; k++) {
#pragma unroll
    for (int c = 0; c < K3S_M; c++) {
      const double2 x = xs[k * K3S_M + c];
      y[c].x = fma(a[k].x, x.x, fma(-a[k].y, x.y, y[c].x));
      y[c].y = fma(a[k].x, x.y, fma(a[k].y, x.x, y[c].y));
    }
  }
}
// row r of a planar slot (columns >= d are the slot's zero padding; rows >= d: zeros)
__device__ __forceinline__ void k3s_load_row(const K3SLane& L, const double* slot, double2 (&a)[K3S_D]) {
  if (!L.rowok) {
#pragma unroll
    for (int k = 0; k < K3S_D; k++) a[k] = make_double2(0.0, 0.0);
    return;
  }
  const double* re = slot + L.r * L.S;
  const double* im = re + L.d * L.S;
  double2 pr[5], pi[5];
#pragma unroll
  for (int q = 0; q < 5; q++) { pr[q] = __ldg(reinterpret_cast<const double2*>(re) + q); pi[q] = __ldg(reinterpret_cast<const double2*>(im) + q); }
#pragma unroll
  for (int q = 0; q < 4; q++) { a[2 * q] = make_double2(pr[q].x, pi[q].x); a[2 * q + 1] = make_double2(pr[q].y, pi[q].y); }
  a[8] = make_double2(pr[4].x, pi[4].x);
}
// conj of column r of a planar slot: a[k] = conj(U[k][r])
__device__ __forceinline__ void k3s_load_col_conj(const K3SLane& L, const double* slot, double2 (&a)[K3S_D]) {
  const double* re = slot + L.r;
  const double* im = re + L.d * L.S;
#pragma unroll
  for (int k = 0; k < K3S_D; k++) {
    const bool ok = L.rowok && k < L.d;
    a[k] = ok ? make_double2(__ldg(re + k * L.S), -__ldg(im + k * L.S)) : make_double2(0.0, 0.0);
  }
}
// the nine lanes of a group pull one slot (bytes) towards L2, 128-byte line by line: no registers held, unlike a register prefetch
__device__ __forceinline__ void k3s_prefetch_l2(const void* base, int bytes, int r) {
  const char* q = reinterpret_cast<const char*>(base);
  for (int o = r * 128; o < bytes; o += 9 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + o));
}
#ifndef K3S_PF
#define K3S_PF 1   // how many slices ahead the backward pass prefetches U_k, dU_k/du_j and x_k into L2 (zz batch, K3S alone:
                   // none 0.858 ms, one slice ahead 0.806, two 0.880, three 0.957: further ahead the lines are evicted again)
#endif
// sum over the nine lanes of a group, result valid in the group's lane 0
__device__ __forceinline__ double k3s_group_sum(double v, int r) {
  const unsigned F = 0xffffffffu;
  double t = __shfl_down_sync(F, v, 8); if (r == 0) v += t;
  t = __shfl_down_sync(F, v, 4); if (r < 4) v += t;
  t = __shfl_down_sync(F, v, 2); if (r < 2) v += t;
  t = __shfl_down_sync(F, v, 1); if (r < 1) v += t;
  return v;
}

__global__ void __launch_bounds__(K3S_WPB * 32, K3S_MINB) k3s_kernel(K23Params p, int S) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l = lane < 27 ? lane : 26;
  K3SLane L;
  L.d = p.d; L.m = p.m; L.S = S; L.g9 = 9 * (l / 9); L.r = l - L.g9;
  const int g = l / 9;
  const bool act = lane < 27;
  L.rowok = L.r < p.d;
  const int d = p.d, m = p.m, nc = p.nc, dm = d * m;
  const size_t slot_d = (size_t)2 * d * S;
  // group-private exchange buffers: [buf 0..3][k][c]
  double2* ex = reinterpret_cast<double2*>(smem_raw) + (size_t)((warp * 3 + g) * 4) * (K3S_D * K3S_M);
  for (int e = L.r; e < 4 * K3S_D * K3S_M; e += 9) ex[e] = make_double2(0.0, 0.0);   // (the shadow lanes repeat lane 26's stores)
  __syncwarp();
  const int gid = (blockIdx.x * K3S_WPB + warp) * 3 + g, gstride = gridDim.x * K3S_WPB * 3;
  const bool contract = p.want_grad != 0;
  // running state penalty (src/penalty_fcns.jl:1-11): lambda_k = U_k' lambda_{k+1} + 2 mu x_k on the penalised entries
  // (src/gradient_computations.jl:55-57).  k3_mode 1 = the pre-pass of the affine segment recurrence: forward states, sum of
  // L(x_k) over the segment, and c_seg = the recurrence run from a zero costate (K2 adds it at the segment boundaries).
  const bool pen = p.row_mask != 0u && p.col_mask != 0u;
  const bool prepass = p.k3_mode == 1;
  const bool rowpen = pen && L.rowok && ((p.row_mask >> L.r) & 1u);
  const double two_mu = 2.0 * p.mu;

  for (int seg = gid; __any_sync(FULL, seg < p.nseg); seg += gstride) {
    const bool on = act && seg < p.nseg;
    const int sg = seg < p.nseg ? seg : 0;
    const int b = sg / p.seg_per_pulse, si = sg - b * p.seg_per_pulse;
    const int k0 = (int)(((long long)si * p.nt) / p.seg_per_pulse);
    const int k1 = (int)(((long long)(si + 1) * p.nt) / p.seg_per_pulse);
    const int len = on ? k1 - k0 : 0;
    int maxlen = len;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { const int o = __shfl_xor_sync(FULL, maxlen, off); maxlen = o > maxlen ? o : maxlen; }
    const size_t sl0 = (size_t)b * p.nt + k0;
    const bool wr = on && L.rowok;    // this lane owns a live row of a live segment

    // ---------------- forward: x_{k0} = x_start(seg); X[k] <- x_k; x_{k+1} = U_k x_k ----------------
    double2 x[K3S_M];
#pragma unroll
    for (int c = 0; c < K3S_M; c++)
      x[c] = (wr && c < m) ? reinterpret_cast<const double2*>(p.xs_start + (size_t)sg * 2 * dm)[L.r + d * c] : make_double2(0.0, 0.0);
    double2 a[K3S_D];
    if (len > 1) k3s_load_row(L, p.U + sl0 * slot_d, a);
    int buf = 0;
    double ps = 0.0;
    for (int i = 0; i < maxlen; i++) {
      const bool live = i < len;
      if (prepass && rowpen && wr && live) {
#pragma unroll
        for (int c = 0; c < K3S_M; c++)
          if ((p.col_mask >> c) & 1u) ps += x[c].x * x[c].x + x[c].y * x[c].y;
      }
      if (wr && live) {
        double2* Xk = reinterpret_cast<double2*>(p.X + ((size_t)b * (p.nt + 1) + k0 + i) * 2 * dm);
#pragma unroll
        for (int c = 0; c < K3S_M; c++)
          if (c < m) Xk[L.r + d * c] = x[c];
      }
      double2* xs = ex + buf * (K3S_D * K3S_M);
      if (act) {
#pragma unroll
        for (int c = 0; c < K3S_M; c++) xs[L.r * K3S_M + c] = x[c];
      }
      __syncwarp();
      if (i + 1 < maxlen) {
        double2 an[K3S_D];
        const bool more = i + 2 < len;
        if (more) k3s_load_row(L, p.U + (sl0 + i + 1) * slot_d, an);   // next slice's row: in flight during this product
        double2 y[K3S_M];
        k3s_row_times(a, xs, y);
        if (i + 1 < len) {
#pragma unroll
          for (int c = 0; c < K3S_M; c++) x[c] = y[c];
        }
        if (more) {
#pragma unroll
          for (int k = 0; k < K3S_D; k++) a[k] = an[k];
        }
      }
      buf ^= 1;
    }
    if (prepass) {   // sum_{k in segment} L(x_k)     src/penalty_fcns.jl:2-4
      ps = k3s_group_sum(ps, L.r);
      if (on && L.r == 0 && ps != 0.0) atomicAdd(&p.Jpen[b], p.mu * ps);
    }
    if (!contract && !prepass) continue;

    // ---------------- backward: lambda_{k1} = lambda_end(seg); gradient of slice k; lambda_k = U_k' lambda_{k+1} ----------------
    // (pre-pass: the same recurrence from a zero costate, no gradient: its result is c_seg)
    double2 lam[K3S_M];
#pragma unroll
    for (int c = 0; c < K3S_M; c++)
      lam[c] = (wr && c < m && !prepass) ? reinterpret_cast<const double2*>(p.lam_end + (size_t)sg * 2 * dm)[L.r + d * c] : make_double2(0.0, 0.0);
    for (int it = 0; it < maxlen; it++) {
      const bool live = it < len;
      const int k = k1 - 1 - it;                         // slice index within the pulse
      const size_t sl = live ? sl0 + (len - 1 - it) : sl0;
      if (K3S_PF > 0 && on && it + K3S_PF < len) {   // the operands of the step K3S_PF slices further down: towards L2 now
        const size_t sp = sl - K3S_PF;
        k3s_prefetch_l2(p.U + sp * slot_d, (int)(slot_d * 8), L.r);
        if (contract)
          for (int j = 0; j < nc; j++) k3s_prefetch_l2(p.L + (sp * nc + j) * slot_d, (int)(slot_d * 8), L.r);
        k3s_prefetch_l2(p.X + ((size_t)b * (p.nt + 1) + (k - K3S_PF)) * 2 * dm, 2 * dm * 8, L.r);
      }
      // operands of this step straight from HBM / L2: conj column of U_k, x_k, and (below) the Jacobian rows
      double2 uc[K3S_D];
      double2 xk[K3S_M];
      k3s_load_col_conj(L, p.U + sl * slot_d, uc);
      {
        const double2* Xk = reinterpret_cast<const double2*>(p.X + ((size_t)b * (p.nt + 1) + (live ? k : k0)) * 2 * dm);
#pragma unroll
        for (int c = 0; c < K3S_M; c++) xk[c] = (wr && c < m) ? Xk[L.r + d * c] : make_double2(0.0, 0.0);
      }
      if (wr && live && p.store_costates && p.LAM && !prepass) {
        double2* Lk = reinterpret_cast<double2*>(p.LAM + ((size_t)b * (p.nt + 1) + k + 1) * 2 * dm);
#pragma unroll
        for (int c = 0; c < K3S_M; c++)
          if (c < m) Lk[L.r + d * c] = lam[c];
      }
      double2* ls = ex + (2 + (it & 1)) * (K3S_D * K3S_M);   // lambda_{k+1}, [row][c]
      double2* xs = ex + (it & 1) * (K3S_D * K3S_M);         // x_k, [row][c]
      if (act) {
#pragma unroll
        for (int c = 0; c < K3S_M; c++) { ls[L.r * K3S_M + c] = lam[c]; xs[L.r * K3S_M + c] = xk[c]; }
      }
      __syncwarp();
      if (contract) {
      // weights of my row: w[cc] = sum_l conj(lambda[r][l]) x_k[cc][l]
      double2 w[K3S_D];
#pragma unroll
      for (int cc = 0; cc < K3S_D; cc++) {
        double wr_ = 0.0, wi_ = 0.0;
#pragma unroll
        for (int c = 0; c < K3S_M; c++) {
          const double2 xv = xs[cc * K3S_M + c];
          wr_ = fma(lam[c].x, xv.x, fma(lam[c].y, xv.y, wr_));
          wi_ = fma(lam[c].x, xv.y, fma(-lam[c].y, xv.x, wi_));
        }
        w[cc] = make_double2(wr_, wi_);
      }
      for (int j = 0; j < nc; j++) {
        double2 dj[K3S_D];
        k3s_load_row(L, p.L + (sl * nc + j) * slot_d, dj);
        double s = 0.0;
#pragma unroll
        for (int cc = 0; cc < K3S_D; cc++) s = fma(dj[cc].x, w[cc].x, fma(-dj[cc].y, w[cc].y, s));   // Re(dU[r][cc] w[cc])
        s = k3s_group_sum(s, L.r);
        if (on && live && L.r == 0) p.dJdu[sl * nc + j] = s;
      }
      }
      // costate: lambda_k[r][c] = sum_rr conj(U[rr][r]) lambda_{k+1}[rr][c]  (+ dL_dx(x_k))
      double2 y[K3S_M];
      k3s_row_times(uc, ls, y);
      if (rowpen) {
#pragma unroll
        for (int c = 0; c < K3S_M; c++)
          if ((p.col_mask >> c) & 1u) { y[c].x = fma(two_mu, xk[c].x, y[c].x); y[c].y = fma(two_mu, xk[c].y, y[c].y); }
      }
      if (live) {
#pragma unroll
        for (int c = 0; c < K3S_M; c++) lam[c] = y[c];
      }
    }
    if (prepass) {
      if (wr) {
        double2* cs = reinterpret_cast<double2*>(p.cs + (size_t)sg * 2 * dm);
#pragma unroll
        for (int c = 0; c < K3S_M; c++)
          if (c < m) cs[L.r + d * c] = lam[c];
      }
      continue;
    }
    if (wr && k0 == 0 && p.store_costates && p.LAM) {
      double2* L0 = reinterpret_cast<double2*>(p.LAM + ((size_t)b * (p.nt + 1)) * 2 * dm);
#pragma unroll
      for (int c = 0; c < K3S_M; c++)
        if (c < m) L0[L.r + d * c] = lam[c];
    }
  }
}

}  // namespace qoc

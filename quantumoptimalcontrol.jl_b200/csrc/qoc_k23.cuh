// qoc_k23.cuh -- K2: boundary scan over the segment propagators with the fidelity trace reduction and the
// terminal costate fused in; K3: per-segment forward sweep, backward (costate) sweep and the adjoint gradient
// contraction <lambda_{k+1}| dU_k/du_j |x_k> fused into the backward sweep.
//
// Together with the level-1 scan inside K1 these replace the reference's three serial loops
//   x[k+1] = U_k x[k]                              src/gradient_computations.jl:27-29
//   lambda[k] = U_k' lambda[k+1] (+ dL_dx(x[k]))   src/gradient_computations.jl:52-58
//   dJdu[j,k] = sum_l Re(dot(lambda[k+1][:,l], dU_k/du_j, x[k][:,l]))   :65-74, :217-223
// and the cost closures of src/penalty_fcns.jl:15-24 (trace infidelity), test/test_gradient_computation.jl:24
// (1-|tr|) and src/penalty_fcns.jl:1-11 (running state penalty).
//
// States are d x m with m <= 8: one 8-wide DMMA column tile; the d x d operand (U_k, U_k^dagger via the
// transposed fragment pattern, dU_k/du_j, Q_seg) is staged HBM -> shared memory through a cp.async ring.
#pragma once
#include "qoc_cost.cuh"
#include "qoc_tiles.cuh"

namespace qoc {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void stage_slot(double* dst, const double* src, int n2, int tid, int nthreads) {
  for (int e = tid; e < n2; e += nthreads) cp_async16(dst + 2 * e, src + 2 * e);
}

// y = op(A) x for the 8 x 8 output tile (mi, 0); x is compact planar in shared memory: xr[k*m + c], xi[k*m + c]
template <class C, bool ADJ>
__device__ __forceinline__ void mv_acc(double (&cr)[2], double (&ci)[2], Mat A, const double* xr, const double* xi,
                                       int d, int m, int mi, int lane) {
  constexpr int S = C::S;
  const int g = lane >> 2, q = lane & 3;
  const double* are;
  const double* aim;
  int astep;
  if (ADJ) { are = A.re + q * S + mi * 8 + g; aim = A.im + q * S + mi * 8 + g; astep = 4 * S; }
  else     { are = A.re + (mi * 8 + g) * S + q; aim = A.im + (mi * 8 + g) * S + q; astep = 4; }
  // The sweeps are latency-bound (one small product per step of a serial recurrence), so the four real products of
  // the complex tile go to separate accumulators, split once more by k-step parity: 8 independent DMMA chains of
  // length <= 4 instead of 2 chains of length 14.
  double a1[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, a2[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  double a3[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, a4[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
  for (int ks = 0; ks < C::KS; ks++) {
    double ar = are[ks * astep], ai = aim[ks * astep];
    if (ADJ) ai = -ai;
    const int k = ks * 4 + q;
    double br = 0.0, bi = 0.0;
    if (k < d && g < m) { br = xr[k * m + g]; bi = xi[k * m + g]; }
    const int h = ks & 1;
    dmma(a1[h][0], a1[h][1], ar, br);
    dmma(a3[h][0], a3[h][1], ar, bi);
    dmma(a2[h][0], a2[h][1], -ai, bi);
    dmma(a4[h][0], a4[h][1], ai, br);
  }
  cr[0] += (a1[0][0] + a1[1][0]) + (a2[0][0] + a2[1][0]);
  cr[1] += (a1[0][1] + a1[1][1]) + (a2[0][1] + a2[1][1]);
  ci[0] += (a3[0][0] + a3[1][0]) + (a4[0][0] + a4[1][0]);
  ci[1] += (a3[0][1] + a3[1][1]) + (a4[0][1] + a4[1][1]);
}

// write the tile result into a compact planar state buffer
__device__ __forceinline__ void mv_store(double* yr, double* yi, const double (&cr)[2], const double (&ci)[2], int d,
                                         int m, int mi, int lane) {
  const int row = mi * 8 + (lane >> 2), col = 2 * (lane & 3);
  if (row < d) {
    if (col < m) { yr[row * m + col] = cr[0]; yi[row * m + col] = ci[0]; }
    if (col + 1 < m) { yr[row * m + col + 1] = cr[1]; yi[row * m + col + 1] = ci[1]; }
  }
}

// compact planar (shared) <-> c128 column-major interleaved (global): element (r, c) at 2*(r + d*c)
__device__ __forceinline__ void state_to_global(double* g, const double* xr, const double* xi, int d, int m, int tid,
                                                int nthreads) {
  for (int e = tid; e < d * m; e += nthreads) {
    const int c = e / d, r = e - c * d;
    reinterpret_cast<double2*>(g)[e] = make_double2(xr[r * m + c], xi[r * m + c]);
  }
}
__device__ __forceinline__ void state_from_global(double* xr, double* xi, const double* g, int d, int m, int tid,
                                                  int nthreads) {
  for (int e = tid; e < d * m; e += nthreads) {
    const int c = e / d, r = e - c * d;
    double2 v = reinterpret_cast<const double2*>(g)[e];
    xr[r * m + c] = v.x;
    xi[r * m + c] = v.y;
  }
}

struct K23Params {
  int d, m, nc, nt, batch, nseg, seg_per_pulse;
  int cost, n;                 // qoc_cost, normalisation
  int want_grad;               // K3: run the backward sweep + contraction
  int store_states, store_costates;
  const double* U;             // planar slots [batch*nt]
  const double* L;             // planar slots [(b*nt+k)*nc + j]
  const double* Q;             // planar slots [nseg]
  const double* x0;            // c128 col-major d x m            (shared by all pulses)
  const double* x_start_ext;   // non-NULL (time sharding): per-pulse entering state, replaces x0
  const double* T;             // c128 col-major d x m target
  const double* lam_final;     // non-NULL: externally supplied terminal costate, d x m x batch
  double* xs_start;            // [nseg] d x m c128: state entering each segment
  double* lam_end;             // [nseg] d x m c128: costate leaving each segment (i.e. lambda_{k1})
  double* X;                   // [(b*(nt+1) + k)] d x m c128 states
  double* LAM;                 // same shape, costates (optional)
  double* x_final;             // [batch] d x m c128
  double* lam_start;           // [batch] d x m c128 : lambda_0 (time sharding hand-over)
  double* J;                   // [batch]
  double* dJdu;                // nc x nt x batch
  int skip_cost;               // K2: forward only, J/lambda_N supplied later
  int k2_phase;                // 0: forward + cost + backward; 1: forward only; 2: backward only (lam_final);
                               // 3: cost + backward starting from the stored x_final (penalty path)
  // running state penalty  L(x) = mu * sum |x[pen_rows, pen_cols]|^2   src/penalty_fcns.jl:1-11
  unsigned row_mask, col_mask;  // bit r / bit c set: row r / column c is penalised (d <= 32, m <= 8); 0 = no penalty
  double mu;
  double* cs;                  // [nseg] d x m c128: affine term of the segment-level costate recurrence
  double* Jpen;                // [batch] sum_k L(x_k), accumulated by K3 (atomicAdd)
  int k3_mode;                 // 0: forward (+ backward + contraction if want_grad); 1: forward + penalty pre-pass
  long long* dbg;              // developer timeline (NULL in production): clock64 stamps of CTA 0, 4 per recurrence step
  int dbg_steps;
};

// mv_store variant that adds the penalty gradient 2 mu x[r][c] (src/penalty_fcns.jl:5-9) to the stored costate
__device__ __forceinline__ void mv_store_pen(double* yr, double* yi, const double (&cr)[2], const double (&ci)[2], int d,
                                             int m, int mi, int lane, const double* xr, const double* xi,
                                             unsigned row_mask, unsigned col_mask, double two_mu) {
  const int row = mi * 8 + (lane >> 2), col = 2 * (lane & 3);
  if (row < d) {
    const bool rp = (row_mask >> row) & 1u;
    if (col < m) {
      double a = cr[0], b = ci[0];
      if (rp && ((col_mask >> col) & 1u)) { a = fma(two_mu, xr[row * m + col], a); b = fma(two_mu, xi[row * m + col], b); }
      yr[row * m + col] = a; yi[row * m + col] = b;
    }
    if (col + 1 < m) {
      double a = cr[1], b = ci[1];
      if (rp && ((col_mask >> (col + 1)) & 1u)) { a = fma(two_mu, xr[row * m + col + 1], a); b = fma(two_mu, xi[row * m + col + 1], b); }
      yr[row * m + col + 1] = a; yi[row * m + col + 1] = b;
    }
  }
}

// mu * sum |x[pen]|^2 of one state (thread-strided partial; caller reduces)
__device__ __forceinline__ double penalty_partial(const double* xr, const double* xi, int d, int m, unsigned row_mask,
                                                  unsigned col_mask, int tid, int nthreads) {
  double s = 0.0;
  for (int e = tid; e < d * m; e += nthreads) {
    const int r = e / m, c = e - r * m;
    if (((row_mask >> r) & 1u) && ((col_mask >> c) & 1u)) s += xr[e] * xr[e] + xi[e] * xi[e];
  }
  return s;
}

constexpr int K2_NST = 6;

// K2: one CTA per pulse, NT warps.
template <class C>
__global__ void __launch_bounds__(C::NT * 32, 1) k2_kernel(K23Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int S = C::S;
  constexpr int NTH = C::NT * 32;
  const int d = p.d, m = p.m, spp = p.seg_per_pulse;
  const int slot_d = 2 * d * S, n2 = slot_d / 2;
  const int tid = threadIdx.x, lane = tid & 31, mi = tid >> 5;
  const int b = blockIdx.x;
  double* ring = reinterpret_cast<double*>(smem_raw);
  double* tail = ring + (size_t)K2_NST * slot_d;  // 8*S zero pad
  double* xbuf = tail + 8 * S;                    // 2 buffers x (re, im) x d*m
  double* red = xbuf + 4 * d * m;                 // 4 doubles
  __shared__ double ov[16];                       // per-column overlaps diag(T' x_N)
  {
    double2* z = reinterpret_cast<double2*>(ring);
    const int total2 = (K2_NST * slot_d + 8 * S) / 2;
    for (int e = tid; e < total2; e += NTH) z[e] = make_double2(0.0, 0.0);
  }
  __syncthreads();
  const int dm = d * m;
  const size_t seg0 = (size_t)b * spp;
  int cur = 0;
  auto XR = [&](int w) { return xbuf + (size_t)w * 2 * dm; };
  auto XI = [&](int w) { return xbuf + (size_t)w * 2 * dm + dm; };

  const bool pen = (p.row_mask != 0u) && (p.col_mask != 0u);
  if (p.k2_phase == 3) {
    state_from_global(XR(0), XI(0), p.x_final + (size_t)b * 2 * dm, d, m, tid, NTH);
    __syncthreads();
  }
  if (p.k2_phase != 2 && p.k2_phase != 3) {
    // ---------------- forward over segments ----------------
    const double* xin = p.x_start_ext ? p.x_start_ext + (size_t)b * 2 * dm : p.x0;
    state_from_global(XR(0), XI(0), xin, d, m, tid, NTH);
    for (int i = 0; i < K2_NST - 1; i++) {
      if (i < spp) stage_slot(ring + (size_t)i * slot_d, p.Q + (seg0 + i) * slot_d, n2, tid, NTH);
      cp_async_commit();
    }
    for (int i = 0; i < spp; i++) {
      cp_async_wait<K2_NST - 2>();
      __syncthreads();
      {
        const int nx = i + K2_NST - 1;
        if (nx < spp) stage_slot(ring + (size_t)(nx % K2_NST) * slot_d, p.Q + (seg0 + nx) * slot_d, n2, tid, NTH);
        cp_async_commit();
      }
      state_to_global(p.xs_start + (seg0 + i) * 2 * dm, XR(cur), XI(cur), d, m, tid, NTH);
      Mat Qm; Qm.re = ring + (size_t)(i % K2_NST) * slot_d; Qm.im = Qm.re + d * S;
      double cr[2] = {0.0, 0.0}, ci[2] = {0.0, 0.0};
      mv_acc<C, false>(cr, ci, Qm, XR(cur), XI(cur), d, m, mi, lane);
      mv_store(XR(cur ^ 1), XI(cur ^ 1), cr, ci, d, m, mi, lane);
      cur ^= 1;
    }
    cp_async_wait<0>();
    __syncthreads();
    // x_N
    if (p.x_final) state_to_global(p.x_final + (size_t)b * 2 * dm, XR(cur), XI(cur), d, m, tid, NTH);
    if (p.X && p.store_states) state_to_global(p.X + ((size_t)b * (p.nt + 1) + p.nt) * 2 * dm, XR(cur), XI(cur), d, m, tid, NTH);
  }
  if (p.k2_phase == 1) return;

  // ---------------- terminal cost and costate ----------------
  // x_N is in buffer `cur`; lambda_N is assembled in buffer cur^1, which then becomes current.
  //   lambda_N = dJfinal_dx(x_N) (+ dL_dx(x_N))                     src/gradient_computations.jl:46-49
  //   J        = Jfinal(x_N) + sum_{k=0..N} L(x_k)                  examples/ipopt_callbacks_exp.jl:18
  {
    const bool have_x = (p.k2_phase != 2);                 // phase 2 has no forward pass in this launch
    const bool builtin = have_x && p.cost != 2;
    if (tid < 4) red[tid] = 0.0;
    if (tid < 16) ov[tid] = 0.0;
    if (!have_x && pen) state_from_global(XR(cur), XI(cur), p.x_final + (size_t)b * 2 * dm, d, m, tid, NTH);
    __syncthreads();
    double J = 0.0;
    CostCoef cc;
    if (builtin)   // per-column overlaps m_c = sum_r conj(T[r][c]) x_N[r][c]      src/penalty_fcns.jl:16,20,32
      cost_overlaps_accumulate(p.T, d, m, [&](int r, int c) { return make_double2(XR(cur)[r * m + c], XI(cur)[r * m + c]); },
                               ov, tid, NTH, lane);
    if (pen && p.k2_phase == 3) {
      double ps = penalty_partial(XR(cur), XI(cur), d, m, p.row_mask, p.col_mask, tid, NTH);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      if (lane == 0) atomicAdd(&red[2], ps);
    }
    __syncthreads();
    if (builtin) { cost_from_overlaps(p.cost, p.n, m, ov, cc); J = cc.J; }
    if (pen && p.k2_phase == 3) J += p.mu * red[2] + p.Jpen[b];
    if (tid == 0 && p.J && (builtin || (pen && p.k2_phase == 3))) p.J[b] = J;
    // terminal costate
    if (p.lam_final) state_from_global(XR(cur ^ 1), XI(cur ^ 1), p.lam_final + (size_t)b * 2 * dm, d, m, tid, NTH);
    for (int e = tid; e < dm; e += NTH) {
      const int c = e / d, r = e - c * d;
      double lr = 0.0, li = 0.0;
      if (p.lam_final) { lr = XR(cur ^ 1)[r * m + c]; li = XI(cur ^ 1)[r * m + c]; }   // same thread wrote it above
      else if (builtin) {
        double2 t = reinterpret_cast<const double2*>(p.T)[e];
        double kr = cc.cr[0], ki = cc.ci[0];   // lambda_N[:, c] = coef_c T[:, c]; only the z-calibrated cost has per-column coefficients
        if (p.cost == QOC_COST_ZCAL_) {
#pragma unroll
          for (int q = 1; q < 4; q++) if (c == q) { kr = cc.cr[q]; ki = cc.ci[q]; }
        }
        lr = kr * t.x - ki * t.y; li = kr * t.y + ki * t.x;
      }
      if (pen && ((p.row_mask >> r) & 1u) && ((p.col_mask >> c) & 1u)) {
        lr = fma(2.0 * p.mu, XR(cur)[r * m + c], lr);
        li = fma(2.0 * p.mu, XI(cur)[r * m + c], li);
      }
      XR(cur ^ 1)[r * m + c] = lr;
      XI(cur ^ 1)[r * m + c] = li;
    }
  }
  cur ^= 1;
  __syncthreads();
  if (p.skip_cost) return;

  // ---------------- backward over segments ----------------
  for (int i = 0; i < K2_NST - 1; i++) {
    const int sgi = spp - 1 - i;
    if (sgi >= 0) stage_slot(ring + (size_t)i * slot_d, p.Q + (seg0 + sgi) * slot_d, n2, tid, NTH);
    cp_async_commit();
  }
  for (int i = 0; i < spp; i++) {
    const int sgi = spp - 1 - i;
    cp_async_wait<K2_NST - 2>();
    __syncthreads();
    {
      const int nx = i + K2_NST - 1;
      if (nx < spp) stage_slot(ring + (size_t)(nx % K2_NST) * slot_d, p.Q + (seg0 + (spp - 1 - nx)) * slot_d, n2, tid, NTH);
      cp_async_commit();
    }
    state_to_global(p.lam_end + (seg0 + sgi) * 2 * dm, XR(cur), XI(cur), d, m, tid, NTH);
    Mat Qm; Qm.re = ring + (size_t)(i % K2_NST) * slot_d; Qm.im = Qm.re + d * S;
    double cr[2] = {0.0, 0.0}, ci[2] = {0.0, 0.0};
    mv_acc<C, true>(cr, ci, Qm, XR(cur), XI(cur), d, m, mi, lane);
    if (pen && p.cs) {  // lambda_{k0} = Q' lambda_{k1} + c_s : the segment-level recurrence is affine with a penalty
      const int row = mi * 8 + (lane >> 2), col = 2 * (lane & 3);
      const double* cs = p.cs + (seg0 + sgi) * 2 * dm;
      if (row < d) {
        if (col < m) { cr[0] += cs[2 * (row + d * col)]; ci[0] += cs[2 * (row + d * col) + 1]; }
        if (col + 1 < m) { cr[1] += cs[2 * (row + d * (col + 1))]; ci[1] += cs[2 * (row + d * (col + 1)) + 1]; }
      }
    }
    mv_store(XR(cur ^ 1), XI(cur ^ 1), cr, ci, d, m, mi, lane);
    cur ^= 1;
  }
  cp_async_wait<0>();
  __syncthreads();
  if (p.lam_start) state_to_global(p.lam_start + (size_t)b * 2 * dm, XR(cur), XI(cur), d, m, tid, NTH);
}

constexpr int K3_NST = 4;

// K3: one CTA per segment; warps [0,NT) run the recurrences, warps [NT*(1+j), NT*(2+j)) contract control j.
template <class C>
__global__ void __launch_bounds__(C::NT * 32 * 3, 1) k3_kernel(K23Params p, int seg_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int S = C::S;
  constexpr int NT = C::NT;
  const int NTH = blockDim.x;
  const int d = p.d, m = p.m, nc = p.nc;
  const int slot_d = 2 * d * S, n2 = slot_d / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool pen = (p.row_mask != 0u) && (p.col_mask != 0u);
  const bool prepass = (p.k3_mode == 1);               // penalty pre-pass: affine term c_s and sum_k L(x_k)
  const bool contract = p.want_grad && !prepass;        // full backward sweep with the gradient contraction
  const int nstage_slots = contract ? (1 + nc) : 1;
  const int dm = d * m;
  double* ring = reinterpret_cast<double*>(smem_raw);
  double* tail = ring + (size_t)K3_NST * nstage_slots * slot_d;
  double* xs = tail + 8 * S;                         // seg_cap states, each 2*dm (re then im)
  double* lbuf = xs + (size_t)seg_cap * 2 * dm;      // 2 costate buffers
  double* part = lbuf + 4 * dm;                      // [2][nc][NT]
  {
    double2* z = reinterpret_cast<double2*>(ring);
    const int total2 = (K3_NST * nstage_slots * slot_d + 8 * S) / 2;
    for (int e = tid; e < total2; e += NTH) z[e] = make_double2(0.0, 0.0);
  }
  __syncthreads();

  for (int seg = blockIdx.x; seg < p.nseg; seg += gridDim.x) {
    const int b = seg / p.seg_per_pulse, si = seg - b * p.seg_per_pulse;
    const int k0 = (int)(((long long)si * p.nt) / p.seg_per_pulse);
    const int k1 = (int)(((long long)(si + 1) * p.nt) / p.seg_per_pulse);
    const int len = k1 - k0;
    if (len <= 0) continue;
    const size_t sl0 = (size_t)b * p.nt + k0;  // first slice index

    // ---------------- forward sweep: xs[i] = x_{k0+i}, i = 0..len-1 ----------------
    state_from_global(xs, xs + dm, p.xs_start + (size_t)seg * 2 * dm, d, m, tid, NTH);
    for (int i = 0; i < K3_NST - 1; i++) {
      if (i < len - 1) stage_slot(ring + (size_t)i * nstage_slots * slot_d, p.U + (sl0 + i) * slot_d, n2, tid, NTH);
      cp_async_commit();
    }
    for (int i = 0; i < len - 1; i++) {
      cp_async_wait<K3_NST - 2>();
      __syncthreads();
      {
        const int nx = i + K3_NST - 1;
        if (nx < len - 1)
          stage_slot(ring + (size_t)(nx % K3_NST) * nstage_slots * slot_d, p.U + (sl0 + nx) * slot_d, n2, tid, NTH);
        cp_async_commit();
      }
      const double* xr = xs + (size_t)i * 2 * dm;
      if (p.store_states) state_to_global(p.X + ((size_t)b * (p.nt + 1) + k0 + i) * 2 * dm, xr, xr + dm, d, m, tid, NTH);
      if (warp < NT) {
        Mat Um; Um.re = ring + (size_t)(i % K3_NST) * nstage_slots * slot_d; Um.im = Um.re + d * S;
        double cr[2] = {0.0, 0.0}, ci[2] = {0.0, 0.0};
        mv_acc<C, false>(cr, ci, Um, xr, xr + dm, d, m, warp, lane);
        double* yr = xs + (size_t)(i + 1) * 2 * dm;
        mv_store(yr, yr + dm, cr, ci, d, m, warp, lane);
      }
    }
    cp_async_wait<0>();
    __syncthreads();
    if (p.store_states) {
      const double* xr = xs + (size_t)(len - 1) * 2 * dm;
      state_to_global(p.X + ((size_t)b * (p.nt + 1) + k1 - 1) * 2 * dm, xr, xr + dm, d, m, tid, NTH);
    }
    if (prepass) {
      // sum_{k in segment} L(x_k), L(x) = mu sum |x[pen]|^2     src/penalty_fcns.jl:2-4
      double ps = 0.0;
      for (int i = 0; i < len; i++) {
        const double* xr = xs + (size_t)i * 2 * dm;
        ps += penalty_partial(xr, xr + dm, d, m, p.row_mask, p.col_mask, tid, NTH);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      if (lane == 0 && ps != 0.0) atomicAdd(&p.Jpen[b], p.mu * ps);
    }
    if (!contract && !prepass) { __syncthreads(); continue; }

    // ---------------- backward sweep (+ gradient contraction) ----------------
    // prepass: lambda' from a ZERO terminal value with the penalty source terms -> c_s = lambda'_{k0}
    int cur = 0;
    auto LR = [&](int w) { return lbuf + (size_t)w * 2 * dm; };
    if (prepass) { for (int e = tid; e < 2 * dm; e += NTH) LR(0)[e] = 0.0; }
    else state_from_global(LR(0), LR(0) + dm, p.lam_end + (size_t)seg * 2 * dm, d, m, tid, NTH);
    if (contract && p.store_costates && p.LAM && k1 == p.nt)
      for (int e = tid; e < dm; e += NTH)
        reinterpret_cast<double2*>(p.LAM + ((size_t)b * (p.nt + 1) + p.nt) * 2 * dm)[e] =
            reinterpret_cast<const double2*>(p.lam_end + (size_t)seg * 2 * dm)[e];
    auto stage_bwd = [&](int it) {  // it-th backward step handles slice k1-1-it
      double* dst = ring + (size_t)(it % K3_NST) * nstage_slots * slot_d;
      const size_t sl = sl0 + (len - 1 - it);
      stage_slot(dst, p.U + sl * slot_d, n2, tid, NTH);
      if (contract)
        for (int j = 0; j < nc; j++) stage_slot(dst + (size_t)(1 + j) * slot_d, p.L + (sl * nc + j) * slot_d, n2, tid, NTH);
    };
    for (int i = 0; i < K3_NST - 1; i++) {
      if (i < len) stage_bwd(i);
      cp_async_commit();
    }
    for (int it = 0; it < len; it++) {
      const int k = k1 - 1 - it;
      cp_async_wait<K3_NST - 2>();
      __syncthreads();
      if (contract && it > 0 && tid < nc) {
        double g = 0.0;
        for (int q = 0; q < NT; q++) g += part[(((it - 1) & 1) * nc + tid) * NT + q];
        p.dJdu[((size_t)b * p.nt + (k + 1)) * nc + tid] = g;
      }
      {
        const int nx = it + K3_NST - 1;
        if (nx < len) stage_bwd(nx);
        cp_async_commit();
      }
      double* stg = ring + (size_t)(it % K3_NST) * nstage_slots * slot_d;
      const double* lr = LR(cur);
      if (warp < NT) {
        Mat Um; Um.re = stg; Um.im = stg + d * S;
        double cr[2] = {0.0, 0.0}, ci[2] = {0.0, 0.0};
        mv_acc<C, true>(cr, ci, Um, lr, lr + dm, d, m, warp, lane);
        double* yr = LR(cur ^ 1);
        if (pen) {  // lambda_k = U_k' lambda_{k+1} + dL_dx(x_k)     src/gradient_computations.jl:55-57
          const double* xk = xs + (size_t)(k - k0) * 2 * dm;
          mv_store_pen(yr, yr + dm, cr, ci, d, m, warp, lane, xk, xk + dm, p.row_mask, p.col_mask, 2.0 * p.mu);
        } else {
          mv_store(yr, yr + dm, cr, ci, d, m, warp, lane);
        }
      } else if (contract) {
        const int ngrp = (NTH >> 5) / NT - 1, grp = warp / NT - 1, mi = warp - (grp + 1) * NT;
        const double* xr = xs + (size_t)(k - k0) * 2 * dm;
        const int row = mi * 8 + (lane >> 2), col = 2 * (lane & 3);
        for (int j = grp; j < nc; j += ngrp) {
          Mat Lm; Lm.re = stg + (size_t)(1 + j) * slot_d; Lm.im = Lm.re + d * S;
          double cr[2] = {0.0, 0.0}, ci[2] = {0.0, 0.0};
          mv_acc<C, false>(cr, ci, Lm, xr, xr + dm, d, m, mi, lane);
          // Re(conj(lambda_{k+1}) .* (dU x_k)) at this thread's two elements   (:217-223)
          double s = 0.0;
          if (row < d) {
            if (col < m) s += lr[row * m + col] * cr[0] + lr[dm + row * m + col] * ci[0];
            if (col + 1 < m) s += lr[row * m + col + 1] * cr[1] + lr[dm + row * m + col + 1] * ci[1];
          }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
          if (lane == 0) part[((it & 1) * nc + j) * NT + mi] = s;
        }
      }
      if (contract && p.store_costates && p.LAM) {
        // lambda_{k+1} is final; write it (lambda_{k1} was written above / by the next segment)
        if (it > 0) state_to_global(p.LAM + ((size_t)b * (p.nt + 1) + k + 1) * 2 * dm, lr, lr + dm, d, m, tid, NTH);
      }
      cur ^= 1;
    }
    cp_async_wait<0>();
    __syncthreads();
    if (prepass) {
      const double* lr = LR(cur);
      state_to_global(p.cs + (size_t)seg * 2 * dm, lr, lr + dm, d, m, tid, NTH);
      __syncthreads();
      continue;
    }
    if (tid < nc) {
      double g = 0.0;
      for (int q = 0; q < NT; q++) g += part[(((len - 1) & 1) * nc + tid) * NT + q];
      p.dJdu[((size_t)b * p.nt + k0) * nc + tid] = g;
    }
    if (p.store_costates && p.LAM) {
      const double* lr = LR(cur);
      state_to_global(p.LAM + ((size_t)b * (p.nt + 1) + k0) * 2 * dm, lr, lr + dm, d, m, tid, NTH);
    }
    __syncthreads();
  }
}

}  // namespace qoc

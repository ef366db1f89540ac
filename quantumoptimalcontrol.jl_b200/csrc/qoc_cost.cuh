// qoc_cost.cuh -- terminal cost and terminal costate from the per-column overlaps m_c = sum_r conj(T[r][c]) x_N[r][c]:
// every built-in cost has lambda_N[:, c] = coef_c * T[:, c], so the sweep kernels (k2_kernel, k2g_kernel, gs_scan_kernel,
// g_sweep_kernel, shard_boundary_kernel) share one evaluation.
//
//   QOC_COST_INFIDELITY  J = 1 - |Omega|^2 / n^2, coef = -2 Omega / n^2, Omega = tr(T'x) = sum_c m_c   src/penalty_fcns.jl:15-24
//   QOC_COST_ABS_TRACE   J = 1 - |Omega|,         coef = -Omega / |Omega|                               test/test_gradient_computation.jl:24-25
//   QOC_COST_ZCAL        m = diag(T'x) (four numbers), F = max_theta |m1 + m2 e^{i theta}| + |m3 + m4 e^{i theta}| by the
//                        reference's golden-section search (tolerance 1e-9), J = 1 - F^2/16,
//                        coef_c = (-2F/16) dF/dm_c with theta held at its optimum (envelope theorem)
//                        src/penalty_fcns.jl:27-42, src/fidelities.jl:48-56 (rrule), :81-101 (optimal_calibration),
//                        :105-137 (_golden_section_search: same bracket updates, same comparisons)
#pragma once
#include <cuda_runtime.h>

namespace qoc {

constexpr int QOC_COST_ZCAL_ = 3;

struct CostCoef {
  double J;
  double cr[8], ci[8];
};

__device__ __forceinline__ double zc_mod2pi(double x) {
  const double tp = 6.283185307179586476925286766559;
  double r = fmod(x, tp);
  if (r < 0.0) r += tp;
  return r;
}

// F and dF/dm of abs_sum_phase_calibrated(m) (src/fidelities.jl:48-56), m = (mr[c], mi[c]), c < 4
__device__ inline void zcal_rrule(const double* mr, const double* mi, double& F, double* dFr, double* dFi) {
  const double PI = 3.141592653589793238462643383279;
  const double ab0 = hypot(mr[0], mi[0]), ab1 = hypot(mr[1], mi[1]), ab2 = hypot(mr[2], mi[2]), ab3 = hypot(mr[3], mi[3]);
  const double a1 = ab0 * ab0 + ab1 * ab1, b1 = 2.0 * ab0 * ab1;
  const double a2 = ab2 * ab2 + ab3 * ab3, b2 = 2.0 * ab2 * ab3;
  const double p1 = zc_mod2pi(atan2(mi[0], mr[0]) - atan2(mi[1], mr[1]));
  const double p2 = zc_mod2pi(atan2(mi[2], mr[2]) - atan2(mi[3], mr[3]));
  double pm, D, al;
  if (fabs(p2 - p1) <= PI) { pm = (p1 + p2) / 2; D = fabs(p2 - p1) / 2; al = p1 < p2 ? 1.0 : -1.0; }
  else { pm = (2 * PI + p1 + p2) / 2; D = PI - fabs(p2 - p1) / 2; al = p1 < p2 ? -1.0 : 1.0; }
  auto negJ = [&](double dl) { return -(sqrt(fmax(a1 + b1 * cos(dl + D), 0.0)) + sqrt(fmax(a2 + b2 * cos(dl - D), 0.0))); };
  // _golden_section_search(f, (-D, D), 1e-9)
  const double gold = 0.5 * (3.0 - sqrt(5.0));
  double lo = -D, hi = D;
  double xm = lo + gold * (hi - lo);
  double fm = negJ(xm);
  while (hi - lo >= 1e-9) {
    if (hi - xm > xm - lo) {
      const double xn = xm + gold * (hi - xm), fn = negJ(xn);
      if (fn < fm) { lo = xm; xm = xn; fm = fn; } else hi = xn;
    } else {
      const double xn = xm - gold * (xm - lo), fn = negJ(xn);
      if (fn < fm) { hi = xm; xm = xn; fm = fn; } else lo = xn;
    }
  }
  F = -fm;
  const double t1 = pm + al * xm;
  double s1, c1;
  sincos(t1, &s1, &c1);
  const double v1r = mr[0] + c1 * mr[1] - s1 * mi[1], v1i = mi[0] + c1 * mi[1] + s1 * mr[1];
  const double v2r = mr[2] + c1 * mr[3] - s1 * mi[3], v2i = mi[2] + c1 * mi[3] + s1 * mr[3];
  const double n1 = hypot(v1r, v1i), n2 = hypot(v2r, v2i);
  const double u1r = v1r / n1, u1i = v1i / n1, u2r = v2r / n2, u2i = v2i / n2;
  // dF_dm = [u1, u1 cis(-t1), u2, u2 cis(-t1)]
  dFr[0] = u1r; dFi[0] = u1i;
  dFr[1] = u1r * c1 + u1i * s1; dFi[1] = u1i * c1 - u1r * s1;
  dFr[2] = u2r; dFi[2] = u2i;
  dFr[3] = u2r * c1 + u2i * s1; dFi[3] = u2i * c1 - u2r * s1;
}

// ov: [2 c] = Re m_c, [2 c + 1] = Im m_c for c < m.  Cheap and uniform: every thread of the CTA may call it redundantly.
__device__ inline void cost_from_overlaps(int cost, int n, int m, const double* ov, CostCoef& o) {
  if (cost == QOC_COST_ZCAL_) {
    double mr[4], mi[4], dFr[4], dFi[4], F;
#pragma unroll
    for (int c = 0; c < 4; c++) { mr[c] = ov[2 * c]; mi[c] = ov[2 * c + 1]; }
    zcal_rrule(mr, mi, F, dFr, dFi);
    o.J = 1.0 - F * F / 16.0;
    const double k = -2.0 * F / 16.0;
#pragma unroll
    for (int c = 0; c < 8; c++) { o.cr[c] = c < 4 ? k * dFr[c] : 0.0; o.ci[c] = c < 4 ? k * dFi[c] : 0.0; }
    return;
  }
  double Or = 0.0, Oi = 0.0;
  for (int c = 0; c < m; c++) { Or += ov[2 * c]; Oi += ov[2 * c + 1]; }
  double cr, ci;
  if (cost == 0) {
    const double nn = (double)n * (double)n;
    o.J = 1.0 - (Or * Or + Oi * Oi) / nn; cr = -2.0 * Or / nn; ci = -2.0 * Oi / nn;
  } else {
    const double a = sqrt(Or * Or + Oi * Oi);
    o.J = 1.0 - a; cr = -Or / a; ci = -Oi / a;
  }
#pragma unroll
  for (int c = 0; c < 8; c++) { o.cr[c] = cr; o.ci[c] = ci; }
}

// lane-0-of-each-warp accumulates the CTA's per-column overlaps into shared ov[2 m] (zeroed before, barrier after by the
// caller).  xload(r, c) returns x_N[r][c] as double2; threads tid, tid + nth, ... take rows.
template <class XL>
__device__ __forceinline__ void cost_overlaps_accumulate(const double* T, int d, int m, XL xload, double* ov, int tid, int nth, int lane) {
  for (int c = 0; c < m; c++) {
    double orr = 0.0, oii = 0.0;
    for (int r = tid; r < d; r += nth) {
      const double2 t = reinterpret_cast<const double2*>(T)[r + (size_t)d * c];
      const double2 x = xload(r, c);
      orr += t.x * x.x + t.y * x.y;
      oii += t.x * x.y - t.y * x.x;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      orr += __shfl_xor_sync(0xffffffffu, orr, off);
      oii += __shfl_xor_sync(0xffffffffu, oii, off);
    }
    if (lane == 0) { atomicAdd(&ov[2 * c], orr); atomicAdd(&ov[2 * c + 1], oii); }
  }
}

}  // namespace qoc

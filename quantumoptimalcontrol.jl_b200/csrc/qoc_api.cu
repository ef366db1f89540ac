// qoc_api.cu -- C ABI (include/qoc_b200.h) over the sm_100a kernels K1/K2/K3.  No CPU fallback: every compute
// entry point requires a CUDA device and fails with QOC_ERR_NO_DEVICE / QOC_ERR_CUDA otherwise.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/qoc_b200.h"
#include "qoc_k1.cuh"
#include "qoc_k1s.cuh"
#include "qoc_k23.cuh"
#include "qoc_sweep.cuh"
#include "qoc_k3s.cuh"
#include "qoc_gpath.cuh"
#include "qoc_basis.cuh"

using namespace qoc;

static thread_local std::string g_create_error;

struct qoc_handle {
  qoc_problem prob;
  int cfg;        // shape class index
  int S;          // planar row stride
  int slot_d;     // doubles per planar slot
  int nsm;
  int spp;        // segments per pulse
  int nseg;
  int seg_cap;    // max slices per segment
  int k1_grid, k1_threads, k3_threads;
  size_t k1_smem, k2_smem, k3_smem;
  cudaStream_t stream;
  // device buffers
  double *dA0p = nullptr, *dAp = nullptr, *du = nullptr, *dU = nullptr, *dL = nullptr, *dQ = nullptr;
  double *dx0 = nullptr, *dT = nullptr, *dxs = nullptr, *dle = nullptr, *dX = nullptr, *dLAM = nullptr;
  double *dxf = nullptr, *dlam0 = nullptr, *dJ = nullptr, *dg = nullptr, *dflops = nullptr, *dlamf = nullptr;
  double *dS = nullptr;
  // more than 8 state columns (full-propagator use, m = d: src/penalty_fcns.jl:14): the columns evolve independently
  // (src/gradient_computations.jl:27-29, :52-58 act column by column), so they are swept in nch chunks of prob.m <= 8 columns
  // (zero-padded) that share one K1 pass; the terminal cost couples them through Omega = tr(T'x) only (chunk_cost_kernel)
  int m_total = 0, nch = 1;
  struct ColBufs { double *dx0, *dT, *dxs, *dle, *dX, *dLAM, *dxf, *dlam0, *dlamf, *dcs, *dg; };
  std::vector<ColBufs> cols;  // nch > 1: cols[0] = the primary buffers above
  // second-generation sweeps (csrc/qoc_sweep.cuh): two-level boundary scan K2G + K3N; used when there is no running penalty
  bool new_k2 = false, new_k3 = false;
  int G = 1;                  // groups per pulse in K2G
  double* dPg = nullptr;      // [batch * G] group propagators
  unsigned* dsync = nullptr;  // [batch] arrival counters of the per-pulse barrier (monotone, never reset)
  unsigned sync_epoch = 0;
  size_t k2g_smem = 0, k3n_smem = 0;
  int k3n_threads = 0, k3n_ngrp = 1, k3n_grid = 0;
  // general path (d > 28): matrices in HBM/L2, batched launches over chunks of slices (csrc/qoc_gpath.cuh)
  bool gpath = false;
  int gchunk = 0, gnw = 0;
  double* gW = nullptr;       // workspace: gnw slots x gchunk slices
  double* dumax = nullptr;    // max_k |u_jk| per control
  double* dubound = nullptr;  // the caller's bounds (qoc_set_control_bounds), device copy
  double* dk1s_scr = nullptr; // k1s_kernel: per-lane-group scratch (A4, W)
  int* dpiv = nullptr;        // pivot rows of the blocked Gauss-Jordan inverse: gchunk x d
  double* dbnd = nullptr;     // time sharding: x_start and lambda_end of the local segment (2 x d x m c128)
  double *dB = nullptr, *dc = nullptr, *ddc = nullptr;   // spline basis, coefficients, dJ/dc (qoc_set_basis / qoc_eval_coeffs)
  int ns = 0;
  double* dQ2 = nullptr;      // second segment-propagator buffer (ping-pong of the batched products)
  int gL = 1;                 // slices per segment on the general path (the last segment of a pulse may be shorter)
  bool gs2 = false;
  // Jacobians streamed instead of stored (general path, two-level sweeps): when U_k and all dU_k/du_j do not fit the device
  // (d = 256, Nt = 1e5, nc = 2: 106 GB + 213 GB), dL holds ONE chunk; the gradient pass re-runs K1 chunk by chunk after the
  // sweeps and contracts each chunk's Jacobians with the stored x_k, lambda_{k+1} at once.  Costs one extra expm per slice
  // ((pi + s + 4/3) M of (pi + s + 4/3 + nc G) M).  QOC_STREAM_JAC=1 forces it (tests), =0 forbids it.
  bool pen_any = false;           // a running state penalty is configured (column chunks: on the CURRENT chunk, see use_cols)
  bool pen_global = false;        // ... on any column
  unsigned col_mask_chunk[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // penalised columns of every chunk (chunk-local bit positions)
  double* dJc = nullptr;          // [nch][batch] column chunks with a penalty: the chunks' sum_k L(x_k)
  std::vector<unsigned char> penrow_host;
  bool pen_hi = false;            // a penalised row >= 64: only the two-level sweeps (byte mask) can carry it
  unsigned char* dpenrow = nullptr;   // [d] 1 = penalised row (general path, two-level sweeps)
  bool have_ubound = false;       // qoc_set_control_bounds: |u_jk| <= ubound[j] promised by the caller
  double ubound[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool stream_jac = false;
  const double* k1_u = nullptr;   // the u of the last K1 launch (device pointer), for the streamed gradient pass
  bool k1_skewh = false;      // A0 and every A_j skew-Hermitian (bitwise): k1s_kernel forms A E + E A, A2 M2 + M2 A2, X E as P + P^dagger
  bool k1s_ok = false;        // d <= 9, nc <= 4: the small-dimension kernel (nine lanes per slice, three slices per warp)
  int k1s_wpb = 0;            // its warps per CTA (what fits shared memory)
  bool shard_fwd_done = false;   // qoc_shard_forward_device ran on the current propagators (the affine call needs its c_s)
  bool k3s_ok = false;        // d <= 9, m <= 4, nc <= 4: the small-dimension sweeps (nine lanes per segment)
  int k3s_grid = 0;
  bool k1_sym = false;        // ... with symmetric H0, H_j: Pade denominator inverted through the real SPD matrix N N^dagger
  bool k1_realh = false;      // K1 real-Hamiltonian fast path (Re A0 = Re A_j = 0, Frechet mode, [13/13] instantiation)
  bool k1_low = true;         // K1 instantiation with the low-degree Pade forms (false when ||A0||_1 alone is far above theta7)           // second-generation general-path sweeps (no running penalty)
  double normA0 = 0.0, normA[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long row_mask64 = 0ull;
  long long* dbg = nullptr;  // developer timeline buffer (qoc_debug_k1_timeline)
  int dbg_slices = 0;
  int dbg_flags = 0;
  int* dstatus = nullptr;
  int *dpen_rows = nullptr, *dpen_cols = nullptr;
  double *dcs = nullptr, *dJpen = nullptr;
  unsigned row_mask = 0, col_mask = 0;
  // host state
  std::vector<double> last_u;
  bool have_u = false;       // a propagate() happened (cache valid)
  bool have_jac = false;
  bool eager_jac = false;    // qoc_propagate also produces dU_k/du_j (qoc_set_eager_jacobians); default: like the reference's
                             // propagate (:17-29), expm only -- qoc_gradient / qoc_get_jacobians add them on demand
  bool states_valid = false, costates_valid = false;
  int launches = 0;
  bool profiling = false;
  cudaEvent_t ev[4];
  double stage_ms[3] = {0, 0, 0};
  double alg_flops = 0.0, k1_exec_flops = 0.0;
  // pinned mailbox: {flops[2], status} ride to the host on the stream, in front of the call's final synchronise, instead of
  // two blocking 4 / 16-byte copies behind it (~20 us per evaluation: 10 % of a short-pulse evaluation)
  double* h_mail = nullptr;
  bool mail_valid = false;
  std::string err;
};

#define QOC_CUDA(h, call)                                                                              \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess) {                                                                          \
      char buf__[512];                                                                                 \
      snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
               __LINE__);                                                                              \
      (h)->err = buf__;                                                                                \
      return QOC_ERR_CUDA;                                                                             \
    }                                                                                                  \
  } while (0)

// ---- shape classes ------------------------------------------------------------------------------------------
typedef Cfg<1, 12, 2> Cfg0;  // d <= 8
typedef Cfg<2, 12, 3> Cfg1;  // d <= 12
typedef Cfg<2, 20, 4> Cfg2;  // d <= 16
typedef Cfg<3, 20, 5> Cfg3;  // d <= 20
typedef Cfg<3, 28, 6> Cfg4;  // d <= 24
typedef Cfg<4, 28, 7> Cfg5;  // d <= 28

static int pick_cfg(int d) {
  if (d <= 8) return 0;
  if (d <= 12) return 1;
  if (d <= 16) return 2;
  if (d <= 20) return 3;
  if (d <= 24) return 4;
  if (d <= 28) return 5;
  return -1;
}

template <class F>
static auto with_cfg(int cfg, F f) {
  switch (cfg) {
    case 0: return f(Cfg0());
    case 1: return f(Cfg1());
    case 2: return f(Cfg2());
    case 3: return f(Cfg3());
    case 4: return f(Cfg4());
    default: return f(Cfg5());
  }
}

static size_t k1_smem_bytes(int d, int nc, int S, int pad_rows) {
  const size_t slot = (size_t)2 * d * S * 8;
  return (size_t)k1_num_slots(nc) * slot + (size_t)pad_rows * S * 8 + sizeof(SvcScratch) + 16;
}
static size_t k2_smem_bytes(int d, int m, int S) {
  const size_t slot = (size_t)2 * d * S * 8;
  return (size_t)K2_NST * slot + (size_t)8 * S * 8 + (size_t)4 * d * m * 8 + 64;
}
static size_t k3_smem_bytes(int d, int m, int nc, int S, int NT, int seg_cap, bool grad) {
  const size_t slot = (size_t)2 * d * S * 8;
  const int nss = grad ? 1 + nc : 1;
  return (size_t)K3_NST * nss * slot + (size_t)8 * S * 8 + (size_t)seg_cap * 2 * d * m * 8 + (size_t)4 * d * m * 8 +
         (size_t)2 * nc * NT * 8 + 64;
}

template <class C>
static size_t k2g_smem_bytes(int d, int m) {
  const size_t slot = (size_t)2 * d * C::S * 8;
  return (size_t)(SW_NST + 2) * slot + (size_t)k1_pad_rows<C>(d) * C::S * 8 + (size_t)4 * state_rows<C>() * 2 * m * 8 + 4 * 8 +
         2 * SW_NST * 8 + 64;
}
template <class C>
static size_t k3n_smem_bytes(int d, int m, int nc, int seg_cap) {
  const size_t slot = (size_t)2 * d * C::S * 8;
  return (size_t)SW_NST * slot + (size_t)k1_pad_rows<C>(d) * C::S * 8 + (size_t)(2 * seg_cap + 1) * state_rows<C>() * 2 * m * 8 +
         2 * SW_NST * 8 + 64;
}

static void to_planar(const double* M, int d, int S, double* out) {  // c128 col-major -> planar slot
  memset(out, 0, sizeof(double) * 2 * d * S);
  for (int c = 0; c < d; c++)
    for (int r = 0; r < d; r++) {
      out[r * S + c] = M[2 * (r + (size_t)d * c)];
      out[d * S + r * S + c] = M[2 * (r + (size_t)d * c) + 1];
    }
}
static void from_planar(const double* P, int d, int S, double* M) {
  for (int c = 0; c < d; c++)
    for (int r = 0; r < d; r++) {
      M[2 * (r + (size_t)d * c)] = P[r * S + c];
      M[2 * (r + (size_t)d * c) + 1] = P[d * S + r * S + c];
    }
}

extern "C" const char* qoc_status_string(int s) {
  switch (s) {
    case QOC_OK: return "ok";
    case QOC_ERR_INVALID: return "invalid argument";
    case QOC_ERR_DIMENSION: return "Error when creating cache, A0 and x0 have incompatiable dimensions";
    case QOC_ERR_STALE_CACHE: return "Cache data from other control signal u";
    case QOC_ERR_UNSUPPORTED: return "size or mode not supported by the CUDA kernels";
    case QOC_ERR_CUDA: return "CUDA runtime error";
    case QOC_ERR_NO_DEVICE: return "no CUDA device of compute capability 10.x (no CPU fallback exists)";
    case QOC_ERR_NOT_FINITE: return "J is not finite";
    case QOC_ERR_SINGULAR: return "Pade denominator numerically singular";
    default: return "unknown status";
  }
}
extern "C" const char* qoc_last_error(const qoc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
extern "C" int qoc_last_launch_count(const qoc_handle* h) { return h ? h->launches : 0; }
extern "C" int qoc_version(void) { return 200; }   // 200: round 2 (lazy Jacobians, z-calibrated cost, in-library sharding)
extern "C" int qoc_set_profiling(qoc_handle* h, int on) {
  if (!h) return QOC_ERR_INVALID;
  h->profiling = on != 0;
  return QOC_OK;
}
extern "C" double qoc_stage_ms(const qoc_handle* h, int stage) {
  return (h && stage >= 0 && stage < 3) ? h->stage_ms[stage] : -1.0;
}
extern "C" double qoc_last_alg_flops(const qoc_handle* h) { return h ? h->alg_flops : 0.0; }
extern "C" double qoc_last_exec_flops(const qoc_handle* h) { return h ? h->k1_exec_flops : 0.0; }

extern "C" int qoc_destroy(qoc_handle* h) {
  if (!h) return QOC_OK;
  cudaSetDevice(h->prob.device);
  double* bufs[] = {h->dA0p, h->dAp, h->du, h->dU, h->dL, h->dQ, h->dx0, h->dT, h->dxs, h->dle, h->dX,
                    h->dLAM, h->dxf, h->dlam0, h->dJ, h->dg, h->dflops, h->dlamf, h->dS, h->dcs, h->dJpen, h->gW, h->dumax, h->dPg, h->dQ2, h->dB, h->dc, h->ddc, h->dbnd, h->dJc};
  if (h->nch > 1 && !h->cols.empty()) {   // the primary pointers may currently alias any chunk: free through the table only
    for (auto& cb : h->cols) {
      double* cbufs[] = {cb.dx0, cb.dT, cb.dxs, cb.dle, cb.dX, cb.dLAM, cb.dxf, cb.dlam0, cb.dlamf, cb.dcs, cb.dg};
      for (double* b : cbufs) if (b) cudaFree(b);
    }
    h->dx0 = h->dT = h->dxs = h->dle = h->dX = h->dLAM = h->dxf = h->dlam0 = h->dlamf = h->dcs = h->dg = nullptr;
    bufs[6] = bufs[7] = bufs[8] = bufs[9] = bufs[10] = bufs[11] = bufs[12] = bufs[13] = bufs[15] = bufs[17] = bufs[19] = nullptr;
  }
  for (double* b : bufs)
    if (b) cudaFree(b);
  if (h->dstatus) cudaFree(h->dstatus);
  if (h->h_mail) cudaFreeHost(h->h_mail);
  if (h->dsync) cudaFree(h->dsync);
  if (h->dpiv) cudaFree(h->dpiv);
  if (h->dk1s_scr) cudaFree(h->dk1s_scr);
  if (h->dubound) cudaFree(h->dubound);
  if (h->dpenrow) cudaFree(h->dpenrow);
  if (h->dpen_rows) cudaFree(h->dpen_rows);
  if (h->dpen_cols) cudaFree(h->dpen_cols);
  for (int i = 0; i < 4; i++) cudaEventDestroy(h->ev[i]);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return QOC_OK;
}

extern "C" int qoc_create(const qoc_problem* prob, const double* A0, const double* A, const double* x0, const double* T,
                          qoc_handle** out) {
  g_create_error.clear();
  if (!prob || !A0 || !x0 || !out || (prob->nc > 0 && !A)) { g_create_error = "NULL argument"; return QOC_ERR_INVALID; }
  *out = nullptr;
  qoc_problem p_local = *prob;
  int m_total = prob->m, nch = 1;
  if (p_local.n <= 0) p_local.n = m_total;
  if (m_total > 8) {   // column chunks of equal width <= 8 (the last one zero-padded)
    if (m_total > 64) { g_create_error = "m > 64 state columns not supported"; return QOC_ERR_UNSUPPORTED; }
    nch = (m_total + 7) / 8;
    p_local.m = (m_total + nch - 1) / nch;
  }
  const qoc_problem& p = p_local;
  if (p.d <= 0 || p.m <= 0 || p.nc <= 0 || p.nt <= 0 || p.batch <= 0) { g_create_error = "non-positive size"; return QOC_ERR_DIMENSION; }
  if (p.order < 0 || p.order > 4) { g_create_error = "order must be 0 (Frechet) or 1..4"; return QOC_ERR_INVALID; }
  if (p.cost < 0 || p.cost > 3) { g_create_error = "unknown cost kind"; return QOC_ERR_INVALID; }
  if (p.cost == QOC_COST_ZCAL && p.m != 4) {   // src/penalty_fcns.jl:28-30
    g_create_error = "Only works for two-qubit gates, x_target must have four columns";
    return QOC_ERR_DIMENSION;
  }
  if (p.cost != QOC_COST_NONE && !T) { g_create_error = "built-in cost needs a target T"; return QOC_ERR_INVALID; }
  if (p.nc > 8) { g_create_error = "nc > 8 controls not supported yet"; return QOC_ERR_UNSUPPORTED; }
  if ((p.n_pen_rows > 0) != (p.n_pen_cols > 0) || (p.n_pen_rows > 0 && (!p.pen_rows || !p.pen_cols))) {
    g_create_error = "penalty index lists inconsistent"; return QOC_ERR_INVALID;
  }
  int cfg = pick_cfg(p.d);
  const char* force_g = getenv("QOC_FORCE_GPATH");
  bool use_gpath = (cfg < 0) || (force_g && force_g[0] == '1');
  if (p.d > 512) { g_create_error = "d > 512 not supported"; return QOC_ERR_UNSUPPORTED; }

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || p.device < 0 || p.device >= ndev) {
    cudaGetLastError();
    g_create_error = "no usable CUDA device (no CPU fallback exists)";
    return QOC_ERR_NO_DEVICE;
  }
  cudaDeviceProp dp;
  if (cudaGetDeviceProperties(&dp, p.device) != cudaSuccess || dp.major != 10) {
    g_create_error = "device is not compute capability 10.x (kernels are built for sm_100a only)";
    return QOC_ERR_NO_DEVICE;
  }
  if (!use_gpath) {
    // the shared-memory-resident kernels need 17 + 3 nc - 2 matrices on chip; otherwise take the general path
    const size_t need = with_cfg(cfg, [&](auto c) -> size_t { typedef decltype(c) C; return k1_smem_bytes(p.d, p.nc, C::S, k1_pad_rows<C>(p.d)); });
    if (need > (size_t)dp.sharedMemPerBlockOptin) use_gpath = true;
  }
  if (use_gpath) {
    cfg = 5;
    // dynamic shared memory of the general-path sweep kernels (gs_scan / gs_seg: two interleaved states; gs_contract: x_k and
    // lambda_{k+1}; g_sweep: three planar states; shard_boundary: two column-major states): rejected here, not at launch
    // (three states with the running penalty: x_k rides along with the costate)
    const size_t need = (size_t)6 * (p.d + 8) * p.m * 8 > (size_t)(3 * p.d + 24) * 2 * p.m * 8 ? (size_t)6 * (p.d + 8) * p.m * 8
                                                                                           : (size_t)(3 * p.d + 24) * 2 * p.m * 8;
    if (need > (size_t)dp.sharedMemPerBlockOptin) {
      g_create_error = "d x m state working set of the general-path sweeps exceeds shared memory";
      return QOC_ERR_UNSUPPORTED;
    }
  }
  qoc_handle* h = new qoc_handle();
  h->prob = p;
  h->m_total = m_total; h->nch = nch;
  h->gpath = use_gpath;
  for (int i = 0; i < p.n_pen_rows; i++) {
    if (p.pen_rows[i] < 0 || p.pen_rows[i] >= p.d) { g_create_error = "penalty row index out of range"; delete h; return QOC_ERR_INVALID; }
    if (p.pen_rows[i] >= 64) h->pen_hi = true;   // beyond the 64-bit row mask of the serial general-path sweep (checked below)
    if (p.pen_rows[i] < 32) h->row_mask |= 1u << p.pen_rows[i];
    if (p.pen_rows[i] < 64) h->row_mask64 |= 1ull << p.pen_rows[i];
    h->penrow_host.resize(p.d, 0);
    h->penrow_host[p.pen_rows[i]] = 1;
  }
  h->pen_global = p.n_pen_rows > 0 && p.n_pen_cols > 0 && p.mu != 0.0;
  for (int i = 0; i < p.n_pen_cols; i++) {   // (p.m is the chunk width when there are more than 8 columns)
    if (p.pen_cols[i] < 0 || p.pen_cols[i] >= m_total) { g_create_error = "penalty column index out of range"; delete h; return QOC_ERR_INVALID; }
    h->col_mask_chunk[p.pen_cols[i] / p.m] |= 1u << (p.pen_cols[i] % p.m);
  }
  h->col_mask = h->col_mask_chunk[0];
  h->pen_any = h->pen_global && h->col_mask != 0u;
  h->prob.pen_rows = nullptr;
  h->prob.pen_cols = nullptr;
  h->cfg = cfg;
  h->nsm = dp.multiProcessorCount;
  h->stream = nullptr;
  for (int i = 0; i < 4; i++) h->ev[i] = nullptr;
  int rc = QOC_OK;
  if (use_gpath) {
    int S = (p.d + 3) / 4 * 4;
    if (S % 8 != 4) S += 4;
    h->S = S;
    h->slot_d = 2 * p.d * S;
    h->spp = 1; h->nseg = p.batch; h->seg_cap = p.nt; h->k1_grid = 0; h->k3_threads = 256;
    {
      const char* old_sw = getenv("QOC_OLD_SWEEPS");
      if (!(old_sw && old_sw[0] == '1') && p.nt >= 4) {   // (with or without the running penalty)
        // two-level sweeps: L slices per segment ~ sqrt(nt) balances the boundary walk (2 spp steps) against the
        // per-segment sweeps (2 L steps)
        int L = (int)ceil(sqrt((double)p.nt));
        if (L < 2) L = 2;
        const int spp = (p.nt + L - 1) / L;
        if ((long long)spp * p.batch <= 65535) { h->gs2 = true; h->gL = L; h->spp = spp; h->nseg = spp * p.batch; h->seg_cap = L; }
      }
    }
    if (h->pen_hi && !h->gs2) {
      g_create_error = "penalty rows >= 64 need the two-level sweeps of the general path (nt >= 4)";
      delete h;
      return QOC_ERR_UNSUPPORTED;
    }
    h->k1_smem = h->k2_smem = h->k3_smem = 0;
    h->gnw = 25 + p.nc;
    const size_t slotBg = (size_t)h->slot_d * 8;
    // workspace: up to ~8 GB (an HBM3e part has 180), at least two slices per SM so that the one-CTA-per-slice kernels
    // (Gauss-Jordan panels) fill the chip; d = 256: 296 slices = 8.2 GB, d = 64: 2048 slices = 3.7 GB
    size_t chunk = (size_t)(8192ull << 20) / ((size_t)h->gnw * slotBg);
    if (chunk < (size_t)2 * dp.multiProcessorCount) chunk = (size_t)2 * dp.multiProcessorCount;
    if (chunk > 2048) chunk = 2048;
    if (chunk > (size_t)p.batch * p.nt) chunk = (size_t)p.batch * p.nt;
    h->gchunk = (int)chunk;
  } else
  rc = with_cfg(cfg, [&](auto c) -> int {
    typedef decltype(c) C;
    h->S = C::S;
    h->slot_d = 2 * p.d * C::S;
    h->k1_threads = C::NTHREADS;
    h->k1_smem = k1_smem_bytes(p.d, p.nc, C::S, k1_pad_rows<C>(p.d));
    h->k2_smem = k2_smem_bytes(p.d, p.m, C::S);
    if (h->k1_smem > (size_t)dp.sharedMemPerBlockOptin) {
      g_create_error = "K1 working set (16+2*nc matrices) exceeds shared memory for this d / nc";
      return QOC_ERR_UNSUPPORTED;
    }
    QOC_CUDA(h, cudaSetDevice(p.device));
    QOC_CUDA(h, cudaFuncSetAttribute(k1_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->k1_smem));
    QOC_CUDA(h, cudaFuncSetAttribute(k1_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->k1_smem));
    QOC_CUDA(h, cudaFuncSetAttribute(k1_kernel<C, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->k1_smem));
    QOC_CUDA(h, cudaFuncSetAttribute(k2_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->k2_smem));
    {
      const char* off = getenv("QOC_NO_K1S");
      h->k1s_wpb = k1s_warps_per_block(p.nc, (size_t)dp.sharedMemPerBlockOptin);
      { const char* w = getenv("QOC_K1S_WPB"); if (w && atoi(w) > 0 && atoi(w) < h->k1s_wpb) h->k1s_wpb = atoi(w); }
      h->k1s_ok = p.d <= K1S_DMAX && p.nc <= K1S_MAXNC && C::S == 12 && !(off && off[0] == '1') && h->k1s_wpb >= 1;
      if (h->k1s_ok)
        QOC_CUDA(h, cudaFuncSetAttribute(k1s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)k1s_smem_bytes(p.nc, h->k1s_wpb)));
    }
    int occ = 1;
    QOC_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k1_kernel<C, true>, C::NTHREADS + NSW * 32, h->k1_smem));
    if (occ < 1) occ = 1;
    long long target = (long long)h->nsm * occ;
    const long long k1s_workers = (long long)h->nsm * h->k1s_wpb * K1S_GPW;   // lane groups on the chip
    if (h->k1s_ok && target < k1s_workers) target = k1s_workers;
    long long spp = (target + p.batch - 1) / p.batch;
    if (spp < 1) spp = 1;
    if (spp > p.nt) spp = p.nt;
    if (h->k1s_ok && p.batch > 1) {
      // K1S workers (warps) take whole segments: pick the segment count that wastes least of the last wave
      // (4096 pulses on 1184 workers: 1 segment per pulse fills 3.46 waves, 2 fill 6.92)
      const long long W = k1s_workers;
      double best = 1e30; long long bs = spp;
      for (long long s2 = spp; s2 < spp + 6 && s2 <= p.nt; s2++) {
        const long long ns = s2 * p.batch, waves = (ns + W - 1) / W;
        const double waste = (double)(waves * W) / (double)ns + 0.01 * (double)(s2 - spp);
        if (waste < best - 1e-9) { best = waste; bs = s2; }
      }
      spp = bs;
    }
    // K3 keeps the forward states of a segment in shared memory: cap the segment length
    const int ngrp = p.nc < 2 ? p.nc : 2;
    h->k3_threads = C::NT * 32 * (1 + ngrp);
    for (;;) {
      h->seg_cap = (int)((p.nt + spp - 1) / spp);
      h->k3_smem = k3_smem_bytes(p.d, p.m, p.nc, C::S, C::NT, h->seg_cap, true);
      if (h->k3_smem <= (size_t)dp.sharedMemPerBlockOptin || spp >= p.nt) break;
      spp = spp * 2 > p.nt ? p.nt : spp * 2;
    }
    if (h->k3_smem > (size_t)dp.sharedMemPerBlockOptin) {
      g_create_error = "K3 working set exceeds shared memory";
      return QOC_ERR_UNSUPPORTED;
    }
    QOC_CUDA(h, cudaFuncSetAttribute(k3_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->k3_smem));
    h->spp = (int)spp;
    h->nseg = (int)(spp * p.batch);
    h->k1_grid = (int)(h->nseg < target ? h->nseg : target);
    // second-generation sweeps (no running penalty)
    const char* old_sw = getenv("QOC_OLD_SWEEPS");
    const bool pen = h->pen_global;
    {
      // small-dimension sweeps (with or without the running penalty: they carry the affine recurrence and its pre-pass)
      const char* off3 = getenv("QOC_NO_K3S");
      h->k3s_ok = p.d <= K3S_D && p.m <= K3S_M && p.nc <= 4 && C::S == 12 && !(off3 && off3[0] == '1') && !(old_sw && old_sw[0] == '1');
      if (h->k3s_ok) {
        QOC_CUDA(h, cudaFuncSetAttribute(k3s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k3s_smem_bytes()));
        int occs = 1;
        QOC_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occs, k3s_kernel, K3S_WPB * 32, k3s_smem_bytes()));
        h->k3s_grid = h->nsm * (occs < 1 ? 1 : occs);
      }
    }
    if (!pen && !(old_sw && old_sw[0] == '1')) {
      h->k3n_threads = (C::NT + 1 + K3N_CW) * 32;
      h->k3n_smem = k3n_smem_bytes<C>(p.d, p.m, p.nc, h->seg_cap);
      if (h->k3n_smem <= (size_t)dp.sharedMemPerBlockOptin) {
        QOC_CUDA(h, cudaFuncSetAttribute(k3n_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->k3n_smem));
        int occ3 = 1;
        QOC_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3, k3n_kernel<C>, h->k3n_threads, h->k3n_smem));
        if (occ3 < 1) occ3 = 1;
        h->k3n_grid = h->nseg < h->nsm * occ3 ? h->nseg : h->nsm * occ3;
        h->new_k3 = true;
      }
      // K2G: G groups per pulse, chosen to balance the spp/G tile products against the 2 G + 2 spp/G mat-vec steps;
      // every CTA of the launch must be resident at once (per-pulse barrier)
      h->k2g_smem = k2g_smem_bytes<C>(p.d, p.m);
      if (spp >= 8 && h->k2g_smem <= (size_t)dp.sharedMemPerBlockOptin) {
        QOC_CUDA(h, cudaFuncSetAttribute(k2g_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->k2g_smem));
        int occ2 = 0;
        QOC_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k2g_kernel<C>, C::NTHREADS + 32, h->k2g_smem));
        const long long cap = (long long)h->nsm * occ2;
        long long gs = llround(sqrt(0.29 * (double)spp));
        if (gs < 1) gs = 1;
        long long G = (spp + gs - 1) / gs;
        if (G * p.batch > cap) G = cap / p.batch;
        if (G >= 2) { h->G = (int)G; h->new_k2 = true; }
      }
    }
    return QOC_OK;
  });
  if (rc != QOC_OK) { if (g_create_error.empty()) g_create_error = h->err; delete h; return rc; }

  auto fail = [&](int code) { g_create_error = h->err; qoc_destroy(h); return code; };
#define CR(call) do { int r__ = [&]() -> int { QOC_CUDA(h, call); return QOC_OK; }(); if (r__ != QOC_OK) return fail(r__); } while (0)
  CR(cudaSetDevice(p.device));
  CR(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 4; i++) CR(cudaEventCreate(&h->ev[i]));
  const size_t slotB = (size_t)h->slot_d * 8;
  const size_t nsl = (size_t)p.batch * p.nt;
  const size_t dmB = (size_t)2 * p.d * p.m * 8;
  CR(cudaMalloc(&h->dA0p, slotB));
  CR(cudaMalloc(&h->dAp, slotB * p.nc));
  CR(cudaMalloc(&h->du, nsl * p.nc * 8));
  CR(cudaMalloc(&h->dU, nsl * slotB));
  if (h->gpath && h->gs2) {
    size_t fr = 0, tot = 0;
    cudaMemGetInfo(&fr, &tot);
    const char* sj = getenv("QOC_STREAM_JAC");
    // (U_k is allocated already) the Jacobians + the chunk workspace + states, costates and boundary buffers still to come
    const double need = (double)nsl * p.nc * slotB + (double)h->gnw * h->gchunk * slotB + 4.0 * (double)p.batch * (p.nt + 1) * dmB;
    h->stream_jac = (sj && sj[0] == '1') || (!(sj && sj[0] == '0') && need > 0.9 * (double)fr);
  }
  CR(cudaMalloc(&h->dL, (h->stream_jac ? (size_t)h->gchunk : nsl) * p.nc * slotB));
  CR(cudaMalloc(&h->dQ, (size_t)h->nseg * slotB));
  CR(cudaMalloc(&h->dx0, dmB));
  CR(cudaMalloc(&h->dT, dmB));
  CR(cudaMalloc(&h->dxs, (size_t)h->nseg * dmB));
  CR(cudaMalloc(&h->dle, (size_t)h->nseg * dmB));
  CR(cudaMalloc(&h->dX, (size_t)p.batch * (p.nt + 1) * dmB));
  if (p.store_costates || h->gs2) CR(cudaMalloc(&h->dLAM, (size_t)p.batch * (p.nt + 1) * dmB));
  if (h->gs2) CR(cudaMalloc(&h->dQ2, (size_t)h->nseg * slotB));
  CR(cudaMalloc(&h->dxf, (size_t)p.batch * dmB));
  CR(cudaMalloc(&h->dlam0, (size_t)p.batch * dmB));
  CR(cudaMalloc(&h->dlamf, (size_t)p.batch * dmB));
  CR(cudaMalloc(&h->dbnd, 2 * dmB));
  CR(cudaMalloc(&h->dJ, (size_t)p.batch * 8));
  CR(cudaMalloc(&h->dg, nsl * p.nc * 8));
  CR(cudaMalloc(&h->dflops, 16));
  if (h->k1s_ok) CR(cudaMalloc(&h->dk1s_scr, (size_t)h->nsm * K1S_MAXWPB * K1S_GPW * 2 * K1S_MSZ * 16));
  CR(cudaMalloc(&h->dS, slotB));
  {
    auto norm1 = [&](const double* M) {
      double mx = 0;
      for (int c = 0; c < p.d; c++) { double sm_ = 0; for (int r = 0; r < p.d; r++) sm_ += hypot(M[2 * (r + (size_t)p.d * c)], M[2 * (r + (size_t)p.d * c) + 1]); if (sm_ > mx) mx = sm_; }
      return mx;
    };
    h->normA0 = norm1(A0);
    for (int j = 0; j < p.nc; j++) h->normA[j] = norm1(A + (size_t)j * 2 * p.d * p.d);
    // a performance heuristic only: the [13/13]-only instantiation is always correct
    h->k1_low = h->normA0 <= 2.0;
    // real-Hamiltonian fast path of K1: every generator plane Re(A0), Re(A_j) exactly zero (X = -i H dt with a real H),
    // [13/13] instantiation, every gradient mode.  QOC_NO_REALH=1 switches it off (A/B measurements).
    {
      bool re0 = true;
      for (size_t e = 0; e < (size_t)p.d * p.d && re0; e++) re0 = (A0[2 * e] == 0.0);
      for (size_t e = 0; e < (size_t)p.nc * p.d * p.d && re0; e++) re0 = (A[2 * e] == 0.0);
      const char* off = getenv("QOC_NO_REALH");
      h->k1_realh = re0 && !h->k1_low && !(off && off[0] == '1');   // every gradient mode (exact Frechet, Taylor 1..4, expm only)
      // ... and all of them symmetric (H real symmetric, X_k skew-Hermitian): the Pade denominator is inverted through the
      // real SPD matrix N N^dagger.  QOC_NO_REALH=2 keeps the real-plane path but switches this off.
      bool sy = h->k1_realh && !(off && off[0] == '2');
      for (int r = 0; r < p.d && sy; r++)
        for (int cc = 0; cc < r && sy; cc++) {
          sy = (A0[2 * (r + (size_t)p.d * cc) + 1] == A0[2 * (cc + (size_t)p.d * r) + 1]);
          for (int j = 0; j < p.nc && sy; j++) {
            const double* Aj = A + (size_t)j * 2 * p.d * p.d;
            sy = (Aj[2 * (r + (size_t)p.d * cc) + 1] == Aj[2 * (cc + (size_t)p.d * r) + 1]);
          }
        }
      h->k1_sym = sy;
      // skew-Hermitian generators (every physical problem: X = -i H dt with H Hermitian), checked bitwise
      const char* offs = getenv("QOC_NO_SKEWH");
      bool sk = !(offs && offs[0] == '1');
      auto skew = [&](const double* M) {
        for (int r = 0; r < p.d; r++)
          for (int cc = 0; cc <= r; cc++) {
            const double ar = M[2 * (r + (size_t)p.d * cc)], ai = M[2 * (r + (size_t)p.d * cc) + 1];
            const double br = M[2 * (cc + (size_t)p.d * r)], bi = M[2 * (cc + (size_t)p.d * r) + 1];
            if (!(ar == -br && ai == bi)) return false;
          }
        return true;
      };
      sk = sk && skew(A0);
      for (int j = 0; j < p.nc && sk; j++) sk = skew(A + (size_t)j * 2 * p.d * p.d);
      h->k1_skewh = sk;
    }
  }
  if (h->gpath) {
    CR(cudaSetDevice(p.device));
    CR(cudaMalloc(&h->gW, (size_t)h->gnw * h->gchunk * slotB));
    CR(cudaMalloc(&h->dumax, 8 * 8));
    CR(cudaMalloc(&h->dpiv, (size_t)h->gchunk * p.d * 4));
    if (h->pen_global) {
      h->penrow_host.resize(p.d, 0);
      CR(cudaMalloc(&h->dpenrow, p.d));
      CR(cudaMemcpy(h->dpenrow, h->penrow_host.data(), p.d, cudaMemcpyHostToDevice));
    }
  }
  if (h->new_k2) {
    CR(cudaMalloc(&h->dPg, (size_t)p.batch * h->G * slotB));
    CR(cudaMalloc(&h->dsync, (size_t)p.batch * 4));
    CR(cudaMemset(h->dsync, 0, (size_t)p.batch * 4));
  }
  CR(cudaMalloc(&h->dcs, (size_t)h->nseg * dmB));
  CR(cudaMalloc(&h->dJpen, (size_t)p.batch * 8));
  CR(cudaMalloc(&h->dstatus, 4));
  CR(cudaMemset(h->dstatus, 0, 4));
  CR(cudaMemset(h->dJ, 0, (size_t)p.batch * 8));
  {
    std::vector<double> tmp((size_t)h->slot_d * (1 + p.nc));
    to_planar(A0, p.d, h->S, tmp.data());
    for (int j = 0; j < p.nc; j++) to_planar(A + (size_t)j * 2 * p.d * p.d, p.d, h->S, tmp.data() + (size_t)(1 + j) * h->slot_d);
    CR(cudaMemcpy(h->dA0p, tmp.data(), slotB, cudaMemcpyHostToDevice));
    CR(cudaMemcpy(h->dAp, tmp.data() + h->slot_d, slotB * p.nc, cudaMemcpyHostToDevice));
  }
  if (h->nch == 1) {
    CR(cudaMemcpy(h->dx0, x0, dmB, cudaMemcpyHostToDevice));
    if (T) CR(cudaMemcpy(h->dT, T, dmB, cudaMemcpyHostToDevice));
    else CR(cudaMemset(h->dT, 0, dmB));
  } else {
    if (h->pen_global) CR(cudaMalloc(&h->dJc, (size_t)h->nch * p.batch * 8));
    h->cols.resize(h->nch);
    h->cols[0] = qoc_handle::ColBufs{h->dx0, h->dT, h->dxs, h->dle, h->dX, h->dLAM, h->dxf, h->dlam0, h->dlamf, h->dcs, h->dg};
    for (int c = 1; c < h->nch; c++) {
      qoc_handle::ColBufs& cb = h->cols[c];
      memset(&cb, 0, sizeof cb);
      CR(cudaMalloc(&cb.dx0, dmB));
      CR(cudaMalloc(&cb.dT, dmB));
      CR(cudaMalloc(&cb.dxs, (size_t)h->nseg * dmB));
      CR(cudaMalloc(&cb.dle, (size_t)h->nseg * dmB));
      CR(cudaMalloc(&cb.dX, (size_t)p.batch * (p.nt + 1) * dmB));
      if (h->dLAM) CR(cudaMalloc(&cb.dLAM, (size_t)p.batch * (p.nt + 1) * dmB));
      CR(cudaMalloc(&cb.dxf, (size_t)p.batch * dmB));
      CR(cudaMalloc(&cb.dlam0, (size_t)p.batch * dmB));
      CR(cudaMalloc(&cb.dlamf, (size_t)p.batch * dmB));
      CR(cudaMalloc(&cb.dcs, (size_t)h->nseg * dmB));
      CR(cudaMalloc(&cb.dg, nsl * p.nc * 8));
    }
    for (int c = 0; c < h->nch; c++) {   // columns [c mc, (c+1) mc) of x0 / T, zero beyond m_total
      const int c0 = c * p.m, ncol = (c0 + p.m <= h->m_total) ? p.m : (h->m_total - c0 > 0 ? h->m_total - c0 : 0);
      CR(cudaMemset(h->cols[c].dx0, 0, dmB));
      CR(cudaMemset(h->cols[c].dT, 0, dmB));
      if (ncol > 0) {
        CR(cudaMemcpy(h->cols[c].dx0, x0 + (size_t)2 * p.d * c0, (size_t)2 * p.d * ncol * 8, cudaMemcpyHostToDevice));
        if (T) CR(cudaMemcpy(h->cols[c].dT, T + (size_t)2 * p.d * c0, (size_t)2 * p.d * ncol * 8, cudaMemcpyHostToDevice));
      }
    }
  }
#undef CR
  *out = h;
  return QOC_OK;
}

extern "C" int qoc_set_order(qoc_handle* h, int order) {
  if (!h) return QOC_ERR_INVALID;
  if (order < 0 || order > 4) { h->err = "order must be 0 (Frechet) or 1..4"; return QOC_ERR_INVALID; }
  if (order != h->prob.order) { h->prob.order = order; h->have_jac = false; }
  return QOC_OK;
}
extern "C" int qoc_set_control_bounds(qoc_handle* h, const double* umax) {
  if (!h) return QOC_ERR_INVALID;
  if (!umax) { h->have_ubound = false; return QOC_OK; }
  for (int j = 0; j < h->prob.nc && j < 8; j++) {
    if (!(umax[j] >= 0.0) || !std::isfinite(umax[j])) { h->err = "control bounds must be finite and non-negative"; return QOC_ERR_INVALID; }
    h->ubound[j] = umax[j];
  }
  QOC_CUDA(h, cudaSetDevice(h->prob.device));
  if (!h->dubound) QOC_CUDA(h, cudaMalloc(&h->dubound, 64));
  QOC_CUDA(h, cudaMemcpy(h->dubound, h->ubound, 64, cudaMemcpyHostToDevice));
  h->have_ubound = true;
  return QOC_OK;
}
extern "C" int qoc_set_eager_jacobians(qoc_handle* h, int on) {
  if (!h) return QOC_ERR_INVALID;
  h->eager_jac = on != 0;
  return QOC_OK;
}
extern "C" int qoc_set_cost(qoc_handle* h, int cost, const double* T, int n) {
  if (!h) return QOC_ERR_INVALID;
  if (cost < 0 || cost > 3) { h->err = "unknown cost kind"; return QOC_ERR_INVALID; }
  if (cost == QOC_COST_ZCAL && h->m_total != 4) {   // src/penalty_fcns.jl:28-30
    h->err = "Only works for two-qubit gates, x_target must have four columns";
    return QOC_ERR_DIMENSION;
  }
  if (cost != QOC_COST_NONE && !T) { h->err = "built-in cost needs a target T"; return QOC_ERR_INVALID; }
  QOC_CUDA(h, cudaSetDevice(h->prob.device));
  if (T && h->nch == 1) QOC_CUDA(h, cudaMemcpy(h->dT, T, (size_t)2 * h->prob.d * h->prob.m * 8, cudaMemcpyHostToDevice));
  if (T && h->nch > 1)
    for (int c = 0; c < h->nch; c++) {
      const int c0 = c * h->prob.m, ncol = (c0 + h->prob.m <= h->m_total) ? h->prob.m : (h->m_total - c0 > 0 ? h->m_total - c0 : 0);
      if (ncol > 0) QOC_CUDA(h, cudaMemcpy(h->cols[c].dT, T + (size_t)2 * h->prob.d * c0, (size_t)2 * h->prob.d * ncol * 8, cudaMemcpyHostToDevice));
    }
  h->prob.cost = cost;
  h->prob.n = n > 0 ? n : h->m_total;
  return QOC_OK;
}

// ---- launches -------------------------------------------------------------------------------------------------

static K23Params base_k23(qoc_handle* h) {
  const qoc_problem& p = h->prob;
  K23Params q;
  memset(&q, 0, sizeof q);
  q.d = p.d; q.m = p.m; q.nc = p.nc; q.nt = p.nt; q.batch = p.batch; q.nseg = h->nseg; q.seg_per_pulse = h->spp;
  q.cost = p.cost; q.n = p.n;
  q.U = h->dU; q.L = h->dL; q.Q = h->dQ; q.x0 = h->dx0; q.T = h->dT;
  q.xs_start = h->dxs; q.lam_end = h->dle; q.X = h->dX; q.LAM = h->dLAM; q.x_final = h->dxf; q.lam_start = h->dlam0;
  q.J = h->dJ; q.dJdu = h->dg;
  q.store_costates = p.store_costates;
  q.row_mask = h->row_mask; q.col_mask = h->col_mask; q.mu = p.mu; q.cs = h->dcs; q.Jpen = h->dJpen;
  q.k3_mode = 0;
  return q;
}


// ---- general path (d > 28): host-orchestrated batched launches -------------------------------------------------------
namespace {
struct GRun {
  qoc_handle* h;
  cudaStream_t st;
  int nb;           // slices in this chunk
  int tiles;
  GOp W(int slot) const { return GOp{h->gW + (size_t)slot * h->gchunk * h->slot_d, (long long)h->slot_d}; }
  double* Wp(int slot) const { return h->gW + (size_t)slot * h->gchunk * h->slot_d; }
  // C = alpha * sum_p A_p B_p + sum_q beta_q D_q + gamma I
  void gemm(double* C, long long cstride, int npairs, const GOp* A, const GOp* B, double alpha, int nadd, const GOp* D,
            const double* beta, double gamma) const {
    GGemm g;
    memset(&g, 0, sizeof g);
    g.d = h->prob.d; g.S = h->S; g.nb = nb; g.npairs = npairs; g.nadd = nadd; g.alpha = alpha; g.gamma = gamma;
    for (int i = 0; i < npairs; i++) { g.A[i] = A[i]; g.B[i] = B[i]; }
    for (int i = 0; i < nadd; i++) { g.D[i] = D[i]; g.beta[i] = beta[i]; }
    g.C = C; g.cstride = cstride;
    g_gemm_launch(g, nb, st);
    h->launches++;
  }
  // the same with up to two extra outputs (workspace slots) that reuse the product and the D operands:
  //   slot xs[k] = ax[k] P + sum_q bx[k][q] D_q + gx[k] I
  void gemm_x(double* C, long long cstride, int npairs, const GOp* A, const GOp* B, double alpha, int nadd, const GOp* D,
              const double* beta, double gamma, int nx, const int* xs, const double* ax, const double (*bx)[3], const double* gx) const {
    GGemm g;
    memset(&g, 0, sizeof g);
    g.d = h->prob.d; g.S = h->S; g.nb = nb; g.npairs = npairs; g.nadd = nadd; g.alpha = alpha; g.gamma = gamma;
    for (int i = 0; i < npairs; i++) { g.A[i] = A[i]; g.B[i] = B[i]; }
    for (int i = 0; i < nadd; i++) { g.D[i] = D[i]; g.beta[i] = beta[i]; }
    g.C = C; g.cstride = cstride;
    g.nextra = nx;
    for (int k = 0; k < nx; k++) {
      g.Cx[k] = Wp(xs[k]); g.cxstride[k] = h->slot_d; g.alphax[k] = ax[k]; g.gammax[k] = gx[k];
      for (int q = 0; q < 3; q++) g.betax[k][q] = q < nadd ? bx[k][q] : 0.0;
    }
    g_gemm_launch(g, nb, st);
    h->launches++;
  }
  void mm1(int c, GOp a, GOp b, double alpha = 1.0, int nadd = 0, const GOp* D = nullptr, const double* beta = nullptr, double gamma = 0.0) const {
    gemm(Wp(c), h->slot_d, 1, &a, &b, alpha, nadd, D, beta, gamma);
  }
  void mm2(int c, GOp a0, GOp b0, GOp a1, GOp b1, double alpha = 1.0, int nadd = 0, const GOp* D = nullptr, const double* beta = nullptr, double gamma = 0.0) const {
    GOp A[2] = {a0, a1}, B[2] = {b0, b1};
    gemm(Wp(c), h->slot_d, 2, A, B, alpha, nadd, D, beta, gamma);
  }
  void lin(int c, int nadd, const GOp* D, const double* beta, double gamma = 0.0) const { gemm(Wp(c), h->slot_d, 0, nullptr, nullptr, 0.0, nadd, D, beta, gamma); }
};
}  // namespace

// Segment propagators of the general path: Q_seg = U_{k1-1} ... U_{k0}, one batched product launch per slice position of a
// segment (all full-length segments of all pulses in one launch, the shorter last segments of the pulses in a second one).
static int gpath_build_Q(qoc_handle* h, cudaStream_t st) {
  const qoc_problem& p = h->prob;
  const int L = h->gL, spp = h->spp, d = p.d;
  const long long slot = h->slot_d;
  const int len_last = p.nt - (spp - 1) * L;          // 1..L
  struct Grp { int inner, count, len; long long first; };   // segments si in [first, first + inner) of every pulse
  Grp grp[2] = {{spp - 1, (spp - 1) * p.batch, L, 0}, {1, p.batch, len_last, spp - 1}};
  if (len_last == L) { grp[0] = Grp{spp, spp * p.batch, L, 0}; grp[1].count = 0; }
  for (int gi = 0; gi < 2; gi++) {
    const Grp& G = grp[gi];
    if (G.count <= 0) continue;
    double* cur = h->dQ; double* oth = h->dQ2;
    auto launch = [&](int npairs, const GOp* A, const GOp* B, int nadd, const GOp* D, double* C) {
      GGemm g;
      memset(&g, 0, sizeof g);
      g.d = d; g.S = h->S; g.nb = G.count; g.npairs = npairs; g.nadd = nadd; g.alpha = 1.0; g.gamma = 0.0;
      for (int i = 0; i < npairs; i++) { g.A[i] = A[i]; g.B[i] = B[i]; }
      for (int i = 0; i < nadd; i++) { g.D[i] = D[i]; g.beta[i] = 1.0; }
      g.C = C + G.first * slot; g.cstride = slot; g.cinner = G.inner; g.cstride2 = (long long)spp * slot;
      g_gemm_launch(g, G.count, st);
      h->launches++;
    };
    auto Uop = [&](int t) { return GOp{h->dU + (G.first * L + t) * slot, (long long)L * slot, G.inner, (long long)p.nt * slot}; };
    auto Qop = [&](const double* buf) { return GOp{buf + G.first * slot, slot, G.inner, (long long)spp * slot}; };
    { GOp D0 = Uop(0); launch(0, nullptr, nullptr, 1, &D0, cur); }                 // Q = U_{k0}
    for (int t = 1; t < G.len; t++) {
      GOp A = Uop(t), B = Qop(cur);
      launch(1, &A, &B, 0, nullptr, oth);                                           // Q <- U_{k0+t} Q
      double* x = cur; cur = oth; oth = x;
    }
    if (cur != h->dQ) { GOp D0 = Qop(cur); launch(0, nullptr, nullptr, 1, &D0, h->dQ); }
  }
  QOC_CUDA(h, cudaGetLastError());
  return QOC_OK;
}

static int gpath_sweep2(qoc_handle* h, int mode, bool skip_bwd, bool want_grad, const double* d_lam_final, const double* d_x_start,
                        double* d_J, double* d_dJdu, cudaStream_t st, bool scan_only = false) {
  const qoc_problem& p = h->prob;
  GS g;
  memset(&g, 0, sizeof g);
  g.d = p.d; g.S = h->S; g.m = p.m; g.nc = p.nc; g.nt = p.nt; g.cost = p.cost; g.n = p.n;
  g.spp = h->spp; g.L = h->gL; g.mode = mode; g.skip_bwd = skip_bwd ? 1 : 0; g.slot = h->slot_d;
  g.U = h->dU; g.L_ = h->dL; g.Q = h->dQ; g.x0 = h->dx0; g.x_start_ext = d_x_start; g.T = h->dT; g.lam_final = d_lam_final;
  g.xs_start = h->dxs; g.lam_end = h->dle; g.X = h->dX; g.LAM = h->dLAM; g.x_final = h->dxf; g.lam_start = h->dlam0;
  g.J = d_J ? d_J : h->dJ; g.dJdu = d_dJdu ? d_dJdu : h->dg;
  const bool pen = h->pen_any;
  const size_t st_smem = (size_t)(pen ? 3 : 2) * ((p.d + 7) / 8 * 8) * 2 * p.m * 8;
  if (st_smem > 40 * 1024) {   // qoc_create checked it against sharedMemPerBlockOptin
    QOC_CUDA(h, cudaFuncSetAttribute(gs_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st_smem));
    QOC_CUDA(h, cudaFuncSetAttribute(gs_seg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st_smem));
  }
  const int grid = h->nseg < h->nsm * 4 ? h->nseg : h->nsm * 4;
  if (pen) {
    // running penalty: the costate recurrence is affine (src/gradient_computations.jl:55-57), a segment is (Q_seg, c_seg):
    //   boundary walk forward -> per-segment forward states + c_seg + sum L(x_k) -> cost + affine boundary walk -> per-segment
    //   affine backward sweeps -> contraction.  Everything but the two boundary walks is parallel over the segments.
    if (mode == 4) { h->err = "time sharding does not carry the running state penalty"; return QOC_ERR_UNSUPPORTED; }
    g.pen_row = h->dpenrow; g.col_mask = h->col_mask; g.mu = p.mu; g.cs = h->dcs; g.Jpen = h->dJpen;
    if (mode != 2) {
      QOC_CUDA(h, cudaMemsetAsync(h->dJpen, 0, (size_t)p.batch * 8, st));
      g.mode = 1;
      gs_scan_kernel<<<p.batch, GS_NW * 32, st_smem, st>>>(g);
      g.pen_prepass = 1;
      gs_seg_kernel<<<grid, GS_NW * 32, st_smem, st>>>(g, h->nseg);
      g.pen_prepass = 0;
      h->launches += 2;
    }
    g.mode = (mode == 2) ? 2 : 3;
    gs_scan_kernel<<<p.batch, GS_NW * 32, st_smem, st>>>(g);
    h->launches++;
    if (!(mode == 0 && skip_bwd) && mode != 1 && !scan_only) {
      gs_seg_kernel<<<grid, GS_NW * 32, st_smem, st>>>(g, h->nseg);
      h->launches++;
    }
    g.mode = mode;
  } else {
  gs_scan_kernel<<<p.batch, GS_NW * 32, st_smem, st>>>(g);
  h->launches++;
  gs_seg_kernel<<<grid, GS_NW * 32, st_smem, st>>>(g, h->nseg);
  h->launches++;
  }
  if (want_grad && (mode == 2 || mode == 4 || (mode == 0 && !skip_bwd))) {
    const size_t c_smem = (size_t)(h->S + p.d) * 2 * p.m * 8;
    if (c_smem > 40 * 1024) QOC_CUDA(h, cudaFuncSetAttribute(gs_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c_smem));
    gs_contract_kernel<<<dim3((unsigned)((size_t)p.batch * p.nt), p.nc), 256, c_smem, st>>>(g);
    h->launches++;
    if (p.store_costates) h->costates_valid = true;
  }
  QOC_CUDA(h, cudaGetLastError());
  if (mode != 2) h->states_valid = true;
  return QOC_OK;
}

static const double kB13[14] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800., 129060195264000.,
                                10559470521600., 670442572800., 33522128640., 1323241920., 40840800., 960960., 16380., 182., 1.};

static int gpath_k1(qoc_handle* h, const double* d_u, bool want_jac, cudaStream_t st, bool stream_contract = false,
                    double* d_dJdu = nullptr) {
  const qoc_problem& p = h->prob;
  h->k1_u = d_u;
  // Jacobian slots of the chunk at c0: the full array, or (streamed) the one-chunk buffer that is contracted right away
  auto Lbase = [&](size_t c0) { return h->stream_jac ? (size_t)0 : c0; };
  const int nc = p.nc, d = p.d;
  const size_t nsl = (size_t)p.batch * p.nt;
  const bool taylor = p.order != 0;
  // number of squarings from the bound ||X_k||_1 <= ||A0||_1 + sum_j max_k|u_jk| ||A_j||_1 (uniform over the launch)
  double umax[8];
  if (h->have_ubound) {
    // the caller promised |u_jk| <= ubound[j] (qoc_set_control_bounds): nothing to read back, the call stays asynchronous;
    // the promise is checked on the device and a violation is reported like a singular Pade denominator, at the next sync
    for (int j = 0; j < 8; j++) umax[j] = h->ubound[j];
    g_ubound_check_kernel<<<64, 256, 0, st>>>(d_u, nc, (long long)nsl, h->dubound, h->dstatus);
    h->launches++;
  } else {
    QOC_CUDA(h, cudaMemsetAsync(h->dumax, 0, 64, st));
    g_umax_kernel<<<64, 256, 0, st>>>(d_u, nc, (long long)nsl, h->dumax);
    h->launches++;
    QOC_CUDA(h, cudaMemcpyAsync(umax, h->dumax, 64, cudaMemcpyDeviceToHost, st));
    QOC_CUDA(h, cudaStreamSynchronize(st));
  }
  double bound = h->normA0;
  for (int j = 0; j < nc; j++) bound += umax[j] * h->normA[j];
  const double theta = taylor ? 5.4 : 4.74;
  int sq = 0;
  if (bound > theta) sq = (int)ceil(log2(bound / theta));
  if (sq < 0) sq = 0;
  const double sc = ldexp(1.0, -sq);
  // Pade degree from the same bound (uniform over the launch): [5/5] / [7/7] without scaling for small norms
  static const double kB5[6] = {30240., 15120., 3360., 420., 30., 1.};
  static const double kB7[8] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
  int q = 13;
  {
    const char* f13 = getenv("QOC_PADE13");
    if (!(f13 && f13[0] == '1')) {
      if (bound <= (taylor ? 0.25 : 0.2)) q = 5;
      else if (bound <= (taylor ? 0.95 : 0.783)) q = 7;
    }
  }
  const double pi_q = q == 13 ? 6.0 : q == 7 ? 4.0 : 3.0;
  const double* b = q == 13 ? kB13 : q == 7 ? kB7 : kB5;
  enum { A = 0, X, A2, A4, A6, W1, Z1, Wm, V, U, R, M2, M4, M6, T1, T2, Lw, Lv, Dd, Ss, RH, TMP, TMPR, NI, L0 };
  GRun g{h, st, 0, (d + 31) / 32};
  for (size_t c0 = 0; c0 < nsl; c0 += h->gchunk) {
    const int nb = (int)((nsl - c0 < (size_t)h->gchunk) ? nsl - c0 : h->gchunk);
    g.nb = nb;
    g_build_kernel<<<nb, 256, 0, st>>>(d, h->S, nc, h->dA0p, h->dAp, d_u + c0 * nc, sc, g.Wp(A),
                                       (taylor && want_jac && p.order >= 2) ? g.Wp(X) : nullptr, h->slot_d);
    h->launches++;
    // (the linear combinations of the program ride in the epilogues of the products that complete their last term: gemm_x)
    const double zero3[3] = {0.0, 0.0, 0.0};
    g.mm1(A2, g.W(A), g.W(A));
    if (q == 5) {          // A4 = A2 A2;  W = b5 A4 + b3 A2 + b1 I,  V = b4 A4 + b2 A2 + b0 I
      GOp a1 = g.W(A2), b1 = g.W(A2), D[1] = {g.W(A2)};
      const int xs[2] = {Wm, V}; const double ax[2] = {b[5], b[4]}, bx[2][3] = {{b[3], 0, 0}, {b[2], 0, 0}}, gx[2] = {b[1], b[0]};
      g.gemm_x(g.Wp(A4), h->slot_d, 1, &a1, &b1, 1.0, 1, D, zero3, 0.0, 2, xs, ax, bx, gx);
    } else {
      g.mm1(A4, g.W(A2), g.W(A2));
      GOp a1 = g.W(A2), b1 = g.W(A4), D[2] = {g.W(A4), g.W(A2)};
      if (q == 13) {       // A6 = A2 A4;  W1 = b13 A6 + b11 A4 + b9 A2,  Z1 = b12 A6 + b10 A4 + b8 A2
        const int xs[2] = {W1, Z1}; const double ax[2] = {b[13], b[12]}, bx[2][3] = {{b[11], b[9], 0}, {b[10], b[8], 0}}, gx[2] = {0.0, 0.0};
        g.gemm_x(g.Wp(A6), h->slot_d, 1, &a1, &b1, 1.0, 2, D, zero3, 0.0, 2, xs, ax, bx, gx);
        { GOp D3[3] = {g.W(A6), g.W(A4), g.W(A2)}; double be[3] = {b[7], b[5], b[3]}; g.mm1(Wm, g.W(A6), g.W(W1), 1.0, 3, D3, be, b[1]); }
        { GOp D3[3] = {g.W(A6), g.W(A4), g.W(A2)}; double be[3] = {b[6], b[4], b[2]}; g.mm1(V, g.W(A6), g.W(Z1), 1.0, 3, D3, be, b[0]); }
      } else {             // A6 = A2 A4;  W = b7 A6 + b5 A4 + b3 A2 + b1 I,  V = b6 A6 + b4 A4 + b2 A2 + b0 I
        const int xs[2] = {Wm, V}; const double ax[2] = {b[7], b[6]}, bx[2][3] = {{b[5], b[3], 0}, {b[4], b[2], 0}}, gx[2] = {b[1], b[0]};
        g.gemm_x(g.Wp(A6), h->slot_d, 1, &a1, &b1, 1.0, 2, D, zero3, 0.0, 2, xs, ax, bx, gx);
      }
    }
    {   // U = A W, and N = V - U in the same epilogue (in place on V: every element is read before it is written, by one thread)
      GOp a1 = g.W(A), b1 = g.W(Wm), D[1] = {g.W(V)};
      const int xs[1] = {V}; const double ax[1] = {-1.0}, bx[1][3] = {{1.0, 0, 0}}, gx[1] = {0.0};
      g.gemm_x(g.Wp(U), h->slot_d, 1, &a1, &b1, 1.0, 1, D, zero3, 0.0, 1, xs, ax, bx, gx);
    }
    // N^-1: blocked Gauss-Jordan (panels in shared memory, rank-32 DMMA updates), V and NI as the two buffers
    const GOp Ninv{g_inverse_blocked(d, h->S, h->slot_d, nb, g.Wp(V), g.Wp(NI), h->dpiv, h->dstatus, st, &h->launches), (long long)h->slot_d};
    // R = I + 2 N^-1 U
    double* Rout = (sq == 0) ? h->dU + c0 * h->slot_d : g.Wp(R);
    { GOp a1 = Ninv, b1 = g.W(U); g.gemm(Rout, h->slot_d, 1, &a1, &b1, 2.0, 0, nullptr, nullptr, 1.0); }
    GOp Rop{Rout, (long long)h->slot_d};
    if (want_jac) {
      for (int j = 0; j < nc; j++) {
        const GOp E{h->dAp + (size_t)j * h->slot_d, 0};
        double* Lout = h->dL + (Lbase(c0) * nc + j) * h->slot_d;
        const long long lstride = (long long)nc * h->slot_d;
        if (taylor) {
          // expm_jacobian!  src/gradient_computations.jl:177-213 (dt = 1), X = unscaled generator
          if (p.order == 1) { double be[1] = {1.0}; g.gemm(Lout, lstride, 0, nullptr, nullptr, 0.0, 1, &E, be, 0.0); continue; }
          g.mm1(M2, E, g.W(X));   // AjX
          g.mm1(M4, g.W(X), E);   // XAj
          if (p.order == 2) {
            GOp D[3] = {E, g.W(M2), g.W(M4)}; double be[3] = {1.0, 0.5, 0.5};
            g.gemm(Lout, lstride, 0, nullptr, nullptr, 0.0, 3, D, be, 0.0);
            continue;
          }
          GOp A3[3] = {g.W(M2), g.W(M4), g.W(X)}, B3[3] = {g.W(X), g.W(X), g.W(M4)};
          GOp D[3] = {E, g.W(M2), g.W(M4)}; double be[3] = {1.0, 0.5, 0.5};
          if (p.order == 3) { g.gemm(Lout, lstride, 3, A3, B3, 1.0 / 6.0, 3, D, be, 0.0); continue; }
          g.gemm(g.Wp(Lw), h->slot_d, 3, A3, B3, 1.0 / 6.0, 3, D, be, 0.0);
          g.mm1(M6, g.W(X), g.W(X));  // X2
          GOp A4_[4] = {g.W(M2), g.W(M4), g.W(M6), g.W(M6)}, B4_[4] = {g.W(M6), g.W(M6), g.W(M2), g.W(M4)};
          GOp D4[1] = {g.W(Lw)}; double b4[1] = {1.0};
          g.gemm(Lout, lstride, 4, A4_, B4_, 1.0 / 24.0, 1, D4, b4, 0.0);
          continue;
        }
        // exact Frechet derivative, structured block-triangular evaluation (Al-Mohy & Higham 2009, Alg. 6.4)
        g.mm2(M2, g.W(A), E, E, g.W(A));
        if (q == 5) {      // M4 = A2 M2 + M2 A2;  Lw = b5 M4 + b3 M2,  Lv = b4 M4 + b2 M2
          GOp Aa[2] = {g.W(A2), g.W(M2)}, Bb[2] = {g.W(M2), g.W(A2)}, D[1] = {g.W(M2)};
          const int xs[2] = {Lw, Lv}; const double ax[2] = {b[5], b[4]}, bx[2][3] = {{b[3], 0, 0}, {b[2], 0, 0}}, gx[2] = {0.0, 0.0};
          g.gemm_x(g.Wp(M4), h->slot_d, 2, Aa, Bb, 1.0, 1, D, zero3, 0.0, 2, xs, ax, bx, gx);
        } else {
          g.mm2(M4, g.W(A2), g.W(M2), g.W(M2), g.W(A2));
          GOp Aa[2] = {g.W(A4), g.W(M4)}, Bb[2] = {g.W(M2), g.W(A2)}, D[2] = {g.W(M4), g.W(M2)};
          if (q == 13) {   // M6 = A4 M2 + M4 A2;  T1 = b13 M6 + b11 M4 + b9 M2,  T2 = b12 M6 + b10 M4 + b8 M2
            const int xs[2] = {T1, T2}; const double ax[2] = {b[13], b[12]}, bx[2][3] = {{b[11], b[9], 0}, {b[10], b[8], 0}}, gx[2] = {0.0, 0.0};
            g.gemm_x(g.Wp(M6), h->slot_d, 2, Aa, Bb, 1.0, 2, D, zero3, 0.0, 2, xs, ax, bx, gx);
            { GOp D3[3] = {g.W(M6), g.W(M4), g.W(M2)}; double be[3] = {b[7], b[5], b[3]}; g.mm2(Lw, g.W(A6), g.W(T1), g.W(M6), g.W(W1), 1.0, 3, D3, be); }
            { GOp D3[3] = {g.W(M6), g.W(M4), g.W(M2)}; double be[3] = {b[6], b[4], b[2]}; g.mm2(Lv, g.W(A6), g.W(T2), g.W(M6), g.W(Z1), 1.0, 3, D3, be); }
          } else {         // M6 = A4 M2 + M4 A2;  Lw = b7 M6 + b5 M4 + b3 M2,  Lv = b6 M6 + b4 M4 + b2 M2
            const int xs[2] = {Lw, Lv}; const double ax[2] = {b[7], b[6]}, bx[2][3] = {{b[5], b[3], 0}, {b[4], b[2], 0}}, gx[2] = {0.0, 0.0};
            g.gemm_x(g.Wp(M6), h->slot_d, 2, Aa, Bb, 1.0, 2, D, zero3, 0.0, 2, xs, ax, bx, gx);
          }
        }
        {   // D = Lu - Lv and S = Lu + Lv from one product pair (Lu = A Lw + E W)
          GOp Aa[2] = {g.W(A), E}, Bb[2] = {g.W(Lw), g.W(Wm)}, D[1] = {g.W(Lv)};
          const double be[3] = {-1.0, 0, 0};
          const int xs[1] = {Ss}; const double ax[1] = {1.0}, bx[1][3] = {{1.0, 0, 0}}, gx[1] = {0.0};
          g.gemm_x(g.Wp(Dd), h->slot_d, 2, Aa, Bb, 1.0, 1, D, be, 0.0, 1, xs, ax, bx, gx);
        }
        { GOp D[1] = {g.W(Ss)}; double be[1] = {1.0}; g.mm1(RH, g.W(Dd), Rop, 1.0, 1, D, be); }                       // rhs = S + D R
        double* Lj = (sq == 0) ? Lout : g.Wp(L0 + j);
        { GOp a1 = Ninv, b1 = g.W(RH); g.gemm(Lj, (sq == 0) ? lstride : (long long)h->slot_d, 1, &a1, &b1, sc, 0, nullptr, nullptr, 0.0); }
      }
    }
    // squaring phase: L <- R L + L R ; R <- R R
    for (int t = 0; t < sq; t++) {
      const bool last = (t == sq - 1);
      if (want_jac && !taylor) {
        for (int j = 0; j < nc; j++) {
          GOp Lj = g.W(L0 + j);
          GOp Aa[2] = {Rop, Lj}, Bb[2] = {Lj, Rop};
          if (last) g.gemm(h->dL + (Lbase(c0) * nc + j) * h->slot_d, (long long)nc * h->slot_d, 2, Aa, Bb, 1.0, 0, nullptr, nullptr, 0.0);
          else {
            g.gemm(g.Wp(TMP), h->slot_d, 2, Aa, Bb, 1.0, 0, nullptr, nullptr, 0.0);
            GOp D[1] = {g.W(TMP)}; double be[1] = {1.0}; g.lin(L0 + j, 1, D, be);
          }
        }
      }
      double* Rn = last ? h->dU + c0 * h->slot_d : g.Wp(TMPR);
      g.gemm(Rn, h->slot_d, 1, &Rop, &Rop, 1.0, 0, nullptr, nullptr, 0.0);
      if (!last) { GOp D[1] = {g.W(TMPR)}; double be[1] = {1.0}; g.lin(R, 1, D, be); }
    }
    if (stream_contract) {   // dJ/du of this chunk's slices from its Jacobians and the stored x_k, lambda_{k+1} (:65-74, :217-223)
      GS gc;
      memset(&gc, 0, sizeof gc);
      gc.d = p.d; gc.S = h->S; gc.m = p.m; gc.nc = p.nc; gc.nt = p.nt; gc.slot = h->slot_d;
      gc.L_ = h->dL; gc.X = h->dX; gc.LAM = h->dLAM; gc.dJdu = d_dJdu ? d_dJdu : h->dg;
      gc.sl0 = (long long)c0; gc.L_chunked = 1;
      const size_t c_smem = (size_t)(h->S + p.d) * 2 * p.m * 8;
      if (c_smem > 40 * 1024) QOC_CUDA(h, cudaFuncSetAttribute(gs_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c_smem));
      gs_contract_kernel<<<dim3((unsigned)nb, p.nc), 256, c_smem, st>>>(gc);
      h->launches++;
    }
    QOC_CUDA(h, cudaGetLastError());
  }
  // F_alg bookkeeping, q = 13
  {
    const double M = 8.0 * d * d * (double)d;
    double G = 0.0;
    if (want_jac) G = taylor ? (p.order == 1 ? 0.0 : p.order == 2 ? 2.0 : p.order == 3 ? 5.0 : 10.0) : (2.0 * pi_q + 2.0 * sq + 2.0);
    const double f = M * ((pi_q + sq + 4.0 / 3.0) + nc * G) * (double)nsl;
    // (executed flops are not tracked on the general path; a kernel, not a copy from the host stack: graph-capturable)
    g_set2_kernel<<<1, 1, 0, st>>>(h->dflops, f, 0.0);
  }
  h->have_jac = want_jac || h->stream_jac;   // streamed: they are re-formed by the gradient pass, never stored
  if (stream_contract) return QOC_OK;        // (the segment products exist already)
  if (h->gs2) return gpath_build_Q(h, st);
  return QOC_OK;
}

static int gpath_sweep(qoc_handle* h, int phase, bool want_grad, const double* d_lam_final, const double* d_x_start,
                       double* d_J, double* d_dJdu, cudaStream_t st) {
  const qoc_problem& p = h->prob;
  GSweep g;
  memset(&g, 0, sizeof g);
  g.d = p.d; g.S = h->S; g.m = p.m; g.nc = p.nc; g.nt = p.nt; g.cost = p.cost; g.n = p.n;
  g.phase = phase; g.want_grad = want_grad ? 1 : 0; g.store_costates = p.store_costates;
  g.U = h->dU; g.L = h->dL; g.slot = h->slot_d; g.x0 = h->dx0; g.x_start_ext = d_x_start; g.T = h->dT;
  g.lam_final = d_lam_final; g.X = h->dX; g.LAM = h->dLAM; g.x_final = h->dxf; g.lam_start = h->dlam0;
  g.J = d_J ? d_J : h->dJ; g.dJdu = d_dJdu ? d_dJdu : h->dg;
  g.row_mask_lo = (p.mu != 0.0) ? h->row_mask64 : 0ull; g.col_mask = h->col_mask; g.mu = p.mu;
  const size_t sweep_smem = (size_t)6 * p.d * p.m * 8;
  if (sweep_smem > 40 * 1024)
    QOC_CUDA(h, cudaFuncSetAttribute(g_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem));
  g_sweep_kernel<<<p.batch, 256, sweep_smem, st>>>(g);
  h->launches++;
  QOC_CUDA(h, cudaGetLastError());
  h->states_valid = (phase != 2);
  if (want_grad && p.store_costates) h->costates_valid = true;
  return QOC_OK;
}

static int launch_k1(qoc_handle* h, const double* d_u, bool want_jac, cudaStream_t st) {
  if (h->gpath) return gpath_k1(h, d_u, want_jac && !h->stream_jac, st);
  const qoc_problem& p = h->prob;
  K1Params k;
  k.d = p.d; k.nc = p.nc; k.nt = p.nt; k.batch = p.batch; k.order = p.order;
  k.nseg = h->nseg; k.seg_per_pulse = h->spp; k.want_jac = want_jac ? 1 : 0; k.sym = h->k1_sym ? 1 : 0; k.skewh = h->k1_skewh ? 1 : 0;
  k.A0p = h->dA0p; k.Ap = h->dAp; k.u = d_u; k.U = h->dU; k.L = h->dL; k.Q = h->dQ;
  k.flops = h->dflops; k.status = h->dstatus; k.scr = h->dk1s_scr;
  k.theta13 = (p.order == QOC_ORDER_FRECHET) ? 4.74 : 5.4;
  // degree switch points: Higham-2005 table as rounded by the reference's dependency (Taylor mode: only expm is
  // approximated), Al-Mohy-Higham l_m for the Frechet pair.  QOC_PADE13=1 forces the [13/13] form (A/B measurements).
  k.theta5 = (p.order == QOC_ORDER_FRECHET) ? 0.2 : 0.25;
  k.theta7 = (p.order == QOC_ORDER_FRECHET) ? 0.783 : 0.95;
  { const char* f13 = getenv("QOC_PADE13"); if (f13 && f13[0] == '1') { k.theta5 = -1.0; k.theta7 = -1.0; } }
  k.dbg = h->dbg; k.dbg_slices = h->dbg_slices; k.dbg_flags = h->dbg_flags;
  QOC_CUDA(h, cudaMemsetAsync(h->dflops, 0, 16, st));
  with_cfg(h->cfg, [&](auto c) {
    typedef decltype(c) C;
    // NW compute warps + 4 service warps
    if (h->k1s_ok) {
      // small-dimension form: one warp per segment, no CTA barriers (qoc_k1s.cuh)
      const int per_cta = h->k1s_wpb * K1S_GPW;
      const int ctas = (h->nseg + per_cta - 1) / per_cta;
      k1s_kernel<<<ctas < h->nsm ? ctas : h->nsm, h->k1s_wpb * 32, k1s_smem_bytes(p.nc, h->k1s_wpb), st>>>(k, h->S);
    } else if (h->k1_realh) k1_kernel<C, false, true><<<h->k1_grid, C::NTHREADS + NSW * 32, h->k1_smem, st>>>(k);
    else if (h->k1_low) k1_kernel<C, true><<<h->k1_grid, C::NTHREADS + NSW * 32, h->k1_smem, st>>>(k);
    else k1_kernel<C, false><<<h->k1_grid, C::NTHREADS + NSW * 32, h->k1_smem, st>>>(k);
    return 0;
  });
  h->launches += 1;  // kernels only (the flop-counter memset is not counted)
  QOC_CUDA(h, cudaGetLastError());
  h->have_jac = want_jac;
  return QOC_OK;
}

// phase: 0 forward + cost + backward, 1 forward only, 2 backward only (needs lam_final)
static int launch_k2(qoc_handle* h, int phase, bool no_backward, const double* d_lam_final, const double* d_x_start,
                     double* d_J, cudaStream_t st, double* d_S_out = nullptr) {
  K23Params q = base_k23(h);
  q.k2_phase = phase;
  q.skip_cost = no_backward ? 1 : 0;
  q.lam_final = d_lam_final;
  q.x_start_ext = d_x_start;
  q.store_states = 1;
  if (d_J) q.J = d_J;
  if (h->new_k2 && phase != 3) {
    K2GParams P;
    P.q = q; P.G = h->G; P.Pg = h->dPg; P.sync = h->dsync; P.S_out = d_S_out;
    P.sync_target = (unsigned)h->G * (h->sync_epoch + 1);
    const cudaError_t le = with_cfg(h->cfg, [&](auto c) {
      typedef decltype(c) C;
      // cooperative launch: the CTAs of a pulse wait on one another (per-pulse barrier), so they must all be resident
      void* args[] = {&P};
      return cudaLaunchCooperativeKernel((void*)k2g_kernel<C>, dim3(h->prob.batch * h->G), dim3(C::NTHREADS + 32), args, h->k2g_smem, st);
    });
    // the device counters only advance when the launch was accepted: a refused launch must not move the epoch, or every later
    // K2G launch on this handle would spin on a target the counters can never reach
    if (le != cudaSuccess) {
      cudaGetLastError();
      h->err = std::string("cooperative launch of k2g_kernel failed: ") + cudaGetErrorString(le);
      return QOC_ERR_CUDA;
    }
    h->sync_epoch += 1;
  } else
  with_cfg(h->cfg, [&](auto c) {
    typedef decltype(c) C;
    k2_kernel<C><<<h->prob.batch, C::NT * 32, h->k2_smem, st>>>(q);
    return 0;
  });
  h->launches += 1;
  QOC_CUDA(h, cudaGetLastError());
  return QOC_OK;
}

static int launch_k3(qoc_handle* h, bool want_grad, bool store_states, double* d_dJdu, cudaStream_t st, int mode = 0) {
  K23Params q = base_k23(h);
  q.dbg = h->dbg; q.dbg_steps = h->dbg_slices;
  q.k3_mode = mode;
  q.want_grad = want_grad ? 1 : 0;
  q.store_states = store_states ? 1 : 0;
  if (d_dJdu) q.dJdu = d_dJdu;
  const int grid = h->nseg < h->nsm * 2 ? h->nseg : h->nsm * 2;
  if (h->k3s_ok && (mode == 0 || mode == 1) && !h->dbg) {
    // small-dimension form: nine lanes per segment, no CTA barriers (qoc_k3s.cuh)
    const int per_cta = K3S_WPB * 3;
    const int ctas = (h->nseg + per_cta - 1) / per_cta;
    k3s_kernel<<<ctas < h->k3s_grid ? ctas : h->k3s_grid, K3S_WPB * 32, k3s_smem_bytes(), st>>>(q, h->S);
  } else
  if (h->new_k3 && mode == 0)
    with_cfg(h->cfg, [&](auto c) {
      typedef decltype(c) C;
      k3n_kernel<C><<<h->k3n_grid, h->k3n_threads, h->k3n_smem, st>>>(q, h->seg_cap);
      return 0;
    });
  else
  with_cfg(h->cfg, [&](auto c) {
    typedef decltype(c) C;
    k3_kernel<C><<<grid, h->k3_threads, h->k3_smem, st>>>(q, h->seg_cap);
    return 0;
  });
  h->launches += 1;
  QOC_CUDA(h, cudaGetLastError());
  if (store_states) h->states_valid = true;
  if (want_grad && h->prob.store_costates) h->costates_valid = true;
  return QOC_OK;
}

static bool has_penalty(const qoc_handle* h) { return h->pen_any; }

// ---- more than 8 state columns: column chunks sharing one K1 pass ---------------------------------------------------
static void use_cols(qoc_handle* h, int c) {
  const qoc_handle::ColBufs& cb = h->cols[c];
  h->dx0 = cb.dx0; h->dT = cb.dT; h->dxs = cb.dxs; h->dle = cb.dle; h->dX = cb.dX; h->dLAM = cb.dLAM;
  h->dxf = cb.dxf; h->dlam0 = cb.dlam0; h->dlamf = cb.dlamf; h->dcs = cb.dcs; h->dg = cb.dg;
  // the running penalty is separable over columns: every chunk carries its own columns of it
  h->col_mask = h->col_mask_chunk[c];
  h->pen_any = h->pen_global && h->col_mask != 0u;
}
struct ChunkCost {
  int d, mc, nch, cost, n, want_lam, batch;
  const double* Jc;   // [nch][batch] the chunks' running-penalty sums (NULL: no penalty)
  unsigned pen_chunks; // bit c: chunk c has penalised columns
  const double* xf[8];
  const double* T[8];
  double* lamf[8];
  double* J;
};
// one CTA per pulse: Omega = tr(T'x_N) over ALL columns, J and lambda_N = coef T for every chunk
// (src/penalty_fcns.jl:15-24; test/test_gradient_computation.jl:24-25)
__global__ void __launch_bounds__(256) chunk_cost_kernel(ChunkCost q) {
  __shared__ double ov[16];
  const int tid = threadIdx.x, lane = tid & 31, b = blockIdx.x;
  const size_t dm = (size_t)q.d * q.mc;
  if (tid < 16) ov[tid] = 0.0;
  __syncthreads();
  double orr = 0.0, oii = 0.0;
  for (int c = 0; c < q.nch; c++) {
    const double2* x = reinterpret_cast<const double2*>(q.xf[c]) + (size_t)b * dm;
    const double2* t = reinterpret_cast<const double2*>(q.T[c]);
    for (size_t e = tid; e < dm; e += 256) { orr += t[e].x * x[e].x + t[e].y * x[e].y; oii += t[e].x * x[e].y - t[e].y * x[e].x; }
  }
  for (int off = 16; off > 0; off >>= 1) { orr += __shfl_xor_sync(0xffffffffu, orr, off); oii += __shfl_xor_sync(0xffffffffu, oii, off); }
  if (lane == 0) { atomicAdd(&ov[0], orr); atomicAdd(&ov[1], oii); }
  __syncthreads();
  CostCoef cc;
  cc.J = 0.0;
  if (q.cost != QOC_COST_NONE) cost_from_overlaps(q.cost, q.n, 1, ov, cc);
  if (tid == 0 && q.J) {
    double J = cc.J;
    if (q.Jc)
      for (int c = 0; c < q.nch; c++) if ((q.pen_chunks >> c) & 1u) J += q.Jc[(size_t)c * q.batch + b];
    q.J[b] = J;
  }
  if (!q.want_lam || q.cost == QOC_COST_NONE) return;
  for (int c = 0; c < q.nch; c++) {
    const double2* t = reinterpret_cast<const double2*>(q.T[c]);
    double2* l = reinterpret_cast<double2*>(q.lamf[c]) + (size_t)b * dm;
    for (size_t e = tid; e < dm; e += 256) l[e] = make_double2(cc.cr[0] * t[e].x - cc.ci[0] * t[e].y, cc.cr[0] * t[e].y + cc.ci[0] * t[e].x);
  }
}
static int run_sweeps(qoc_handle* h, bool want_grad, const double* d_lam_final, double* d_J, double* d_dJdu,
                      bool store_states, cudaStream_t st);
__global__ void chunk_sum_kernel(double* out, const double* const g0, const double* const g1, const double* const g2, const double* const g3,
                                 const double* const g4, const double* const g5, const double* const g6, const double* const g7, int nch, size_t n) {
  const double* g[8] = {g0, g1, g2, g3, g4, g5, g6, g7};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < nch; c++) s += g[c][i];
    out[i] = s;
  }
}
// forward sweeps of every chunk (cost switched off: it needs all columns), then J (and lambda_N when want_lam)
static int chunked_forward(qoc_handle* h, bool store_states, bool want_lam, double* d_J, cudaStream_t st) {
  const int cost = h->prob.cost;
  int rc = QOC_OK;
  h->prob.cost = QOC_COST_NONE;
  unsigned pen_chunks = 0;
  for (int c = 0; c < h->nch && rc == QOC_OK; c++) {
    use_cols(h, c);
    // (a chunk with penalised columns leaves its sum_k L(x_k) in its own slot: with the cost off, that is all its "J" is)
    if (has_penalty(h)) pen_chunks |= 1u << c;
    rc = run_sweeps(h, false, nullptr, has_penalty(h) ? h->dJc + (size_t)c * h->prob.batch : nullptr, nullptr, store_states, st);
  }
  h->prob.cost = cost;
  use_cols(h, 0);
  if (rc != QOC_OK) return rc;
  if (cost != QOC_COST_NONE || pen_chunks) {
    ChunkCost q;
    memset(&q, 0, sizeof q);
    q.d = h->prob.d; q.mc = h->prob.m; q.nch = h->nch; q.cost = cost; q.n = h->prob.n; q.want_lam = want_lam ? 1 : 0;
    q.batch = h->prob.batch; q.Jc = pen_chunks ? h->dJc : nullptr; q.pen_chunks = pen_chunks;
    for (int c = 0; c < h->nch; c++) { q.xf[c] = h->cols[c].dxf; q.T[c] = h->cols[c].dT; q.lamf[c] = h->cols[c].dlamf; }
    q.J = d_J ? d_J : h->dJ;
    chunk_cost_kernel<<<h->prob.batch, 256, 0, st>>>(q);
    h->launches++;
    QOC_CUDA(h, cudaGetLastError());
  }
  return QOC_OK;
}
// backward sweeps + gradient contraction of every chunk from its lambda_N (cols[c].dlamf), summed into d_dJdu
static int chunked_backward(qoc_handle* h, bool store_states, double* d_dJdu, cudaStream_t st) {
  int rc = QOC_OK;
  for (int c = 0; c < h->nch && rc == QOC_OK; c++) { use_cols(h, c); rc = run_sweeps(h, true, h->cols[c].dlamf, nullptr, h->cols[c].dg, store_states, st); }
  use_cols(h, 0);
  if (rc != QOC_OK) return rc;
  const double* g[8];
  for (int c = 0; c < 8; c++) g[c] = h->cols[c < h->nch ? c : 0].dg;
  const size_t n = (size_t)h->prob.batch * h->prob.nt * h->prob.nc;
  // (the sum may land in chunk 0's own buffer: every element is read before it is written, by the same thread)
  chunk_sum_kernel<<<(int)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024), 256, 0, st>>>(d_dJdu ? d_dJdu : h->cols[0].dg, g[0], g[1], g[2], g[3],
                                                                                            g[4], g[5], g[6], g[7], h->nch, n);
  h->launches++;
  QOC_CUDA(h, cudaGetLastError());
  return QOC_OK;
}
// d x m_total x batch host array <-> the chunks' d x mc x batch device arrays (column blocks)
static int chunked_copy(qoc_handle* h, double* host, bool to_host, size_t per_pulse_items, int which /*0 xf, 1 lamf*/, cudaStream_t st) {
  const qoc_problem& p = h->prob;
  (void)per_pulse_items;
  for (int c = 0; c < h->nch; c++) {
    const int c0 = c * p.m, ncol = (c0 + p.m <= h->m_total) ? p.m : (h->m_total - c0 > 0 ? h->m_total - c0 : 0);
    double* dev = which == 0 ? h->cols[c].dxf : h->cols[c].dlamf;
    if (!to_host) QOC_CUDA(h, cudaMemsetAsync(dev, 0, (size_t)p.batch * 2 * p.d * p.m * 8, st));
    if (ncol == 0) continue;
    const size_t hp = (size_t)2 * p.d * h->m_total * 8, dpitch = (size_t)2 * p.d * p.m * 8, w = (size_t)2 * p.d * ncol * 8;
    double* hptr = host + (size_t)2 * p.d * c0;
    if (to_host) QOC_CUDA(h, cudaMemcpy2DAsync(hptr, hp, dev, dpitch, w, p.batch, cudaMemcpyDeviceToHost, st));
    else QOC_CUDA(h, cudaMemcpy2DAsync(dev, dpitch, hptr, hp, w, p.batch, cudaMemcpyHostToDevice, st));
  }
  return QOC_OK;
}

// Everything after K1.  Without a running penalty: K2 (forward, cost, backward) then K3.  With it the costate
// recurrence is affine, lambda_k = U_k' lambda_{k+1} + dL_dx(x_k) (src/gradient_computations.jl:55-57), so the
// segment-level scan needs the per-segment affine terms first:
//   K2 forward -> K3 forward + pre-pass (c_s, sum L(x_k)) -> K2 cost + affine backward -> K3 forward + backward.
static int run_sweeps(qoc_handle* h, bool want_grad, const double* d_lam_final, double* d_J, double* d_dJdu,
                      bool store_states, cudaStream_t st) {
  int rc;
  const bool builtin = h->prob.cost != QOC_COST_NONE;
  if (h->gpath && h->gs2 && h->stream_jac && (want_grad || d_lam_final)) {
    // sweeps first (states and costates to HBM, no contraction), then the streamed Jacobian pass contracts chunk by chunk
    if ((rc = gpath_sweep2(h, d_lam_final ? 2 : 0, false, false, d_lam_final, nullptr, d_J, d_dJdu, st)) != QOC_OK) return rc;
    return gpath_k1(h, h->k1_u, true, st, true, d_dJdu);
  }
  if (h->gpath && h->gs2) {
    if (d_lam_final) return gpath_sweep2(h, 2, false, true, d_lam_final, nullptr, d_J, d_dJdu, st);
    if (want_grad) return gpath_sweep2(h, 0, false, true, nullptr, nullptr, d_J, d_dJdu, st);
    return gpath_sweep2(h, builtin ? 0 : 1, true, false, nullptr, nullptr, d_J, d_dJdu, st);
  }
  if (h->gpath) {
    if (d_lam_final) return gpath_sweep(h, 2, true, d_lam_final, nullptr, d_J, d_dJdu, st);
    return gpath_sweep(h, want_grad ? 0 : 1, want_grad, nullptr, nullptr, d_J, d_dJdu, st);
  }
  if (!has_penalty(h)) {
    if (d_lam_final) rc = launch_k2(h, 2, false, d_lam_final, nullptr, nullptr, st);
    else rc = launch_k2(h, builtin ? 0 : 1, !want_grad, nullptr, nullptr, d_J, st);
    if (rc != QOC_OK) return rc;
    if (want_grad || store_states) rc = launch_k3(h, want_grad, store_states, d_dJdu, st);
    return rc;
  }
  QOC_CUDA(h, cudaMemsetAsync(h->dJpen, 0, (size_t)h->prob.batch * 8, st));
  if ((rc = launch_k2(h, 1, true, nullptr, nullptr, nullptr, st)) != QOC_OK) return rc;      // forward boundary scan
  if ((rc = launch_k3(h, false, store_states, nullptr, st, 1)) != QOC_OK) return rc;        // c_s and sum_k L(x_k)
  if (d_lam_final) rc = launch_k2(h, 2, false, d_lam_final, nullptr, nullptr, st);           // affine backward
  else rc = launch_k2(h, 3, !want_grad, nullptr, nullptr, d_J, st);                          // cost (+ affine backward)
  if (rc != QOC_OK) return rc;
  if (want_grad) rc = launch_k3(h, true, store_states, d_dJdu, st);
  return rc;
}

// queue the mailbox copies on the stream (call right before the final cudaStreamSynchronize of an evaluation)
static int queue_mail(qoc_handle* h) {
  if (!h->h_mail) QOC_CUDA(h, cudaMallocHost(&h->h_mail, 32));
  QOC_CUDA(h, cudaMemcpyAsync(h->h_mail, h->dflops, 16, cudaMemcpyDeviceToHost, h->stream));
  QOC_CUDA(h, cudaMemcpyAsync(h->h_mail + 2, h->dstatus, 4, cudaMemcpyDeviceToHost, h->stream));
  h->mail_valid = true;
  return QOC_OK;
}

static int check_status(qoc_handle* h) {
  int st = 0;
  if (h->mail_valid) memcpy(&st, h->h_mail + 2, 4);
  else QOC_CUDA(h, cudaMemcpy(&st, h->dstatus, 4, cudaMemcpyDeviceToHost));
  if (st != 0) {
    h->mail_valid = false;
    cudaMemset(h->dstatus, 0, 4);
    if (st == 9) { h->err = "a control amplitude exceeds the bound given to qoc_set_control_bounds"; return QOC_ERR_INVALID; }
    if (st == 10) { h->err = "time-sharded evaluation: a peer rank never published its propagator (timeout)"; return QOC_ERR_CUDA; }
    h->err = "zero pivot while inverting the Pade denominator";
    return QOC_ERR_SINGULAR;
  }
  return QOC_OK;
}

static double sweep_flops(const qoc_problem& p0, bool grad, int m_total = 0) {
  qoc_problem p = p0;
  if (m_total > 0) p.m = m_total;
  const double d2 = (double)p.d * p.d;
  const double per = grad ? 8.0 * d2 * p.m * (2 + p.nc) + 4.0 * p.nc * d2 : 8.0 * d2 * p.m;
  return per * (double)p.nt * (double)p.batch;
}

extern "C" int qoc_eval_device(qoc_handle* h, const double* d_u, double* d_J, double* d_dJdu, void* stream) {
  if (!h || !d_u) return QOC_ERR_INVALID;
  if (h->prob.cost == QOC_COST_NONE) { h->err = "qoc_eval needs a built-in cost"; return QOC_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  QOC_CUDA(h, cudaSetDevice(h->prob.device));
  h->launches = 0;
  int rc;
  if (h->profiling) QOC_CUDA(h, cudaEventRecord(h->ev[0], st));
  if ((rc = launch_k1(h, d_u, true, st)) != QOC_OK) return rc;
  if (h->profiling) QOC_CUDA(h, cudaEventRecord(h->ev[1], st));
  if (h->nch > 1) {
    if (h->profiling) QOC_CUDA(h, cudaEventRecord(h->ev[2], st));
    if ((rc = chunked_forward(h, false, true, d_J, st)) != QOC_OK) return rc;
    if ((rc = chunked_backward(h, h->prob.store_costates != 0, d_dJdu ? d_dJdu : h->cols[0].dg, st)) != QOC_OK) return rc;
  } else
  if (!has_penalty(h) && !h->gpath) {
    if ((rc = launch_k2(h, 0, false, nullptr, nullptr, d_J, st)) != QOC_OK) return rc;
    if (h->profiling) QOC_CUDA(h, cudaEventRecord(h->ev[2], st));
    if ((rc = launch_k3(h, true, h->prob.store_costates != 0, d_dJdu, st)) != QOC_OK) return rc;
  } else {
    if (h->profiling) QOC_CUDA(h, cudaEventRecord(h->ev[2], st));  // K2 and K3 interleave: reported together as K3
    if ((rc = run_sweeps(h, true, nullptr, d_J, d_dJdu, h->prob.store_costates != 0, st)) != QOC_OK) return rc;
  }
  if (h->profiling) {
    QOC_CUDA(h, cudaEventRecord(h->ev[3], st));
    QOC_CUDA(h, cudaEventSynchronize(h->ev[3]));
    for (int i = 0; i < 3; i++) {
      float ms = 0;
      QOC_CUDA(h, cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]));
      h->stage_ms[i] = ms;
    }
  }
  return QOC_OK;
}

static int fetch_flops(qoc_handle* h, bool grad) {
  double f[2] = {0, 0};
  if (h->mail_valid) { f[0] = h->h_mail[0]; f[1] = h->h_mail[1]; h->mail_valid = false; }
  else QOC_CUDA(h, cudaMemcpy(f, h->dflops, 16, cudaMemcpyDeviceToHost));
  h->alg_flops = f[0] + sweep_flops(h->prob, grad, h->m_total);
  h->k1_exec_flops = f[1];
  return QOC_OK;
}

extern "C" int qoc_eval(qoc_handle* h, const double* u, double* J_out, double* dJdu_out) {
  if (!h || !u) return QOC_ERR_INVALID;
  const qoc_problem& p = h->prob;
  const size_t nu = (size_t)p.nc * p.nt * p.batch;
  QOC_CUDA(h, cudaSetDevice(p.device));
  QOC_CUDA(h, cudaMemcpyAsync(h->du, u, nu * 8, cudaMemcpyHostToDevice, h->stream));
  int rc = qoc_eval_device(h, h->du, nullptr, nullptr, h->stream);
  if (rc != QOC_OK) return rc;
  if (J_out) QOC_CUDA(h, cudaMemcpyAsync(J_out, h->dJ, (size_t)p.batch * 8, cudaMemcpyDeviceToHost, h->stream));
  if (dJdu_out) QOC_CUDA(h, cudaMemcpyAsync(dJdu_out, h->dg, nu * 8, cudaMemcpyDeviceToHost, h->stream));
  if ((rc = queue_mail(h)) != QOC_OK) return rc;
  h->last_u.assign(u, u + nu);   // (host copy for the stale-cache check: overlaps the device work queued above)
  QOC_CUDA(h, cudaStreamSynchronize(h->stream));
  h->have_u = true;
  h->states_valid = p.store_costates != 0;
  if ((rc = check_status(h)) != QOC_OK) return rc;
  if ((rc = fetch_flops(h, true)) != QOC_OK) return rc;
  if (J_out)
    for (int b = 0; b < p.batch; b++)
      if (!std::isfinite(J_out[b])) { h->err = "J is not finite"; return QOC_ERR_NOT_FINITE; }
  return QOC_OK;
}

extern "C" int qoc_propagate(qoc_handle* h, const double* u, double* J_out, double* x_final_out) {
  if (!h || !u) return QOC_ERR_INVALID;
  const qoc_problem& p = h->prob;
  const size_t nu = (size_t)p.nc * p.nt * p.batch;
  QOC_CUDA(h, cudaSetDevice(p.device));
  h->launches = 0;
  h->states_valid = false;
  h->costates_valid = false;
  QOC_CUDA(h, cudaMemcpyAsync(h->du, u, nu * 8, cudaMemcpyHostToDevice, h->stream));
  int rc;
  // the reference's propagate does the matrix exponentials only (src/gradient_computations.jl:17-25): an f-only call (line
  // search) does not pay for the Jacobians; qoc_gradient re-runs K1 with them on the cached u when they are missing.
  // qoc_set_eager_jacobians(h, 1) produces them here (they share the Pade powers) when f_grad always follows f.
  if ((rc = launch_k1(h, h->du, h->eager_jac, h->stream)) != QOC_OK) return rc;
  const bool builtin = p.cost != QOC_COST_NONE || h->pen_global;  // J (or its penalty part) is formed on the device
  if (h->nch > 1) {
    if ((rc = chunked_forward(h, false, false, nullptr, h->stream)) != QOC_OK) return rc;
    if (x_final_out && (rc = chunked_copy(h, x_final_out, true, 0, 0, h->stream)) != QOC_OK) return rc;
  } else {
    if ((rc = run_sweeps(h, false, nullptr, nullptr, nullptr, false, h->stream)) != QOC_OK) return rc;
    if (x_final_out)
      QOC_CUDA(h, cudaMemcpyAsync(x_final_out, h->dxf, (size_t)p.batch * 2 * p.d * p.m * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  if (J_out && builtin) QOC_CUDA(h, cudaMemcpyAsync(J_out, h->dJ, (size_t)p.batch * 8, cudaMemcpyDeviceToHost, h->stream));
  if ((rc = queue_mail(h)) != QOC_OK) return rc;
  h->last_u.assign(u, u + nu);   // (host copy for the stale-cache check: overlaps the device work queued above)
  QOC_CUDA(h, cudaStreamSynchronize(h->stream));
  h->have_u = true;
  if ((rc = check_status(h)) != QOC_OK) return rc;
  if ((rc = fetch_flops(h, false)) != QOC_OK) return rc;
  if (J_out && builtin)
    for (int b = 0; b < p.batch; b++)
      if (!std::isfinite(J_out[b])) { h->err = "J is not finite"; return QOC_ERR_NOT_FINITE; }
  return QOC_OK;
}

extern "C" int qoc_gradient(qoc_handle* h, const double* u, const double* lambda_final, double* dJdu_out) {
  if (!h || !dJdu_out) return QOC_ERR_INVALID;
  const qoc_problem& p = h->prob;
  const size_t nu = (size_t)p.nc * p.nt * p.batch;
  if (!h->have_u) { h->err = "qoc_gradient called before qoc_propagate"; return QOC_ERR_STALE_CACHE; }
  if (u && memcmp(u, h->last_u.data(), nu * 8) != 0) {
    h->err = "Cache data from other control signal u";  // src/gradient_computations.jl:37-39
    return QOC_ERR_STALE_CACHE;
  }
  if (!lambda_final && p.cost == QOC_COST_NONE) { h->err = "no built-in cost: lambda_final required"; return QOC_ERR_INVALID; }
  QOC_CUDA(h, cudaSetDevice(p.device));
  h->launches = 0;
  int rc;
  bool k1_rerun = false;
  if (!h->have_jac) {  // propagate ran without Jacobians, or the order changed since: K1 on the cached u (h->du still holds it)
    if ((rc = launch_k1(h, h->du, true, h->stream)) != QOC_OK) return rc;
    k1_rerun = true;
  }
  if (h->nch > 1) {
    if (lambda_final) { if ((rc = chunked_copy(h, const_cast<double*>(lambda_final), false, 0, 1, h->stream)) != QOC_OK) return rc; }
    else if ((rc = chunked_forward(h, false, true, nullptr, h->stream)) != QOC_OK) return rc;   // lambda_N of the built-in cost
    if ((rc = chunked_backward(h, true, h->cols[0].dg, h->stream)) != QOC_OK) return rc;
    QOC_CUDA(h, cudaMemcpyAsync(dJdu_out, h->cols[0].dg, nu * 8, cudaMemcpyDeviceToHost, h->stream));
  } else {
  if (lambda_final)
    QOC_CUDA(h, cudaMemcpyAsync(h->dlamf, lambda_final, (size_t)p.batch * 2 * p.d * p.m * 8, cudaMemcpyHostToDevice, h->stream));
  if ((rc = run_sweeps(h, true, lambda_final ? h->dlamf : nullptr, nullptr, nullptr, true, h->stream)) != QOC_OK) return rc;
  QOC_CUDA(h, cudaMemcpyAsync(dJdu_out, h->dg, nu * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  if (k1_rerun && (rc = queue_mail(h)) != QOC_OK) return rc;
  QOC_CUDA(h, cudaStreamSynchronize(h->stream));
  if (k1_rerun) {
    if ((rc = check_status(h)) != QOC_OK) return rc;
    if ((rc = fetch_flops(h, true)) != QOC_OK) return rc;
  } else
    h->alg_flops += sweep_flops(p, true, h->m_total) - sweep_flops(p, false, h->m_total);
  return QOC_OK;
}

// ---- getters ----------------------------------------------------------------------------------------------------

static int recompute_states(qoc_handle* h) {
  return (h->gpath && h->gs2) ? gpath_sweep2(h, 1, true, false, nullptr, nullptr, nullptr, nullptr, h->stream)
         : h->gpath ? gpath_sweep(h, 1, false, nullptr, nullptr, nullptr, nullptr, h->stream)
                  : launch_k3(h, false, true, nullptr, h->stream);
}
// chunked handles: the d x mc x (Nt+1) x batch arrays of the chunks -> d x m_total x (Nt+1) x batch
static int gather_chunks(qoc_handle* h, bool costates, double* out) {
  const qoc_problem& p = h->prob;
  const size_t nk = (size_t)p.batch * (p.nt + 1), dm = (size_t)2 * p.d * p.m;
  std::vector<double> tmp(nk * dm);
  for (int c = 0; c < h->nch; c++) {
    const int c0 = c * p.m, ncol = (c0 + p.m <= h->m_total) ? p.m : (h->m_total - c0 > 0 ? h->m_total - c0 : 0);
    if (ncol == 0) continue;
    QOC_CUDA(h, cudaMemcpy(tmp.data(), costates ? h->cols[c].dLAM : h->cols[c].dX, tmp.size() * 8, cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < nk; k++)
      memcpy(out + (k * h->m_total + c0) * 2 * p.d, tmp.data() + k * dm, (size_t)2 * p.d * ncol * 8);
  }
  return QOC_OK;
}

extern "C" int qoc_get_states(qoc_handle* h, double* x_out) {
  if (!h || !x_out) return QOC_ERR_INVALID;
  if (!h->have_u) { h->err = "no propagation cached"; return QOC_ERR_STALE_CACHE; }
  const qoc_problem& p = h->prob;
  QOC_CUDA(h, cudaSetDevice(p.device));
  if (h->nch > 1) {
    if (!h->states_valid) {
      for (int c = 0; c < h->nch; c++) { use_cols(h, c); const int rc = recompute_states(h); if (rc != QOC_OK) { use_cols(h, 0); return rc; } }
      use_cols(h, 0);
    }
    QOC_CUDA(h, cudaStreamSynchronize(h->stream));
    return gather_chunks(h, false, x_out);
  }
  if (!h->states_valid) {
    int rc = recompute_states(h);
    if (rc != QOC_OK) return rc;
  }
  QOC_CUDA(h, cudaMemcpyAsync(x_out, h->dX, (size_t)p.batch * (p.nt + 1) * 2 * p.d * p.m * 8, cudaMemcpyDeviceToHost, h->stream));
  QOC_CUDA(h, cudaStreamSynchronize(h->stream));
  return QOC_OK;
}

extern "C" int qoc_get_costates(qoc_handle* h, double* lam_out) {
  if (!h || !lam_out) return QOC_ERR_INVALID;
  if (!h->prob.store_costates || !h->costates_valid) { h->err = "costates not stored (store_costates=0 or no gradient yet)"; return QOC_ERR_STALE_CACHE; }
  const qoc_problem& p = h->prob;
  QOC_CUDA(h, cudaSetDevice(p.device));
  if (h->nch > 1) return gather_chunks(h, true, lam_out);
  QOC_CUDA(h, cudaMemcpy(lam_out, h->dLAM, (size_t)p.batch * (p.nt + 1) * 2 * p.d * p.m * 8, cudaMemcpyDeviceToHost));
  return QOC_OK;
}

static int get_slots(qoc_handle* h, const double* dsrc, size_t nslots, double* out) {
  const qoc_problem& p = h->prob;
  QOC_CUDA(h, cudaSetDevice(p.device));
  std::vector<double> tmp(nslots * h->slot_d);
  QOC_CUDA(h, cudaMemcpy(tmp.data(), dsrc, tmp.size() * 8, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < nslots; i++) from_planar(tmp.data() + i * h->slot_d, p.d, h->S, out + i * 2 * p.d * p.d);
  return QOC_OK;
}
extern "C" int qoc_get_propagators(qoc_handle* h, double* U_out) {
  if (!h || !U_out) return QOC_ERR_INVALID;
  if (!h->have_u) { h->err = "no propagation cached"; return QOC_ERR_STALE_CACHE; }
  return get_slots(h, h->dU, (size_t)h->prob.batch * h->prob.nt, U_out);
}
extern "C" int qoc_get_jacobians(qoc_handle* h, double* dU_out) {
  if (!h || !dU_out) return QOC_ERR_INVALID;
  if (!h->have_u) { h->err = "no propagation cached"; return QOC_ERR_STALE_CACHE; }
  if (h->stream_jac) { h->err = "this handle streams its Jacobians (they do not fit the device): nothing is stored"; return QOC_ERR_UNSUPPORTED; }
  if (!h->have_jac) {   // lazy: K1 with Jacobians on the cached u
    QOC_CUDA(h, cudaSetDevice(h->prob.device));
    const int rc = launch_k1(h, h->du, true, h->stream);
    if (rc != QOC_OK) return rc;
    QOC_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return get_slots(h, h->dL, (size_t)h->prob.batch * h->prob.nt * h->prob.nc, dU_out);
}

// ---- time-segment sharding (phase API) ---------------------------------------------------------------------------
// Sequential product of the local segment propagators: S = Q_{n-1} ... Q_0 (one CTA; the K1 tile machinery).
template <class C>
__global__ void __launch_bounds__(C::NTHREADS, 1) kq_reduce_kernel(const double* Q, int nq, int d, double* out_c128) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int S = C::S;
  const int slot_d = 2 * d * S, n2 = slot_d / 2;
  double* base = reinterpret_cast<double*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int mi = warp / (C::NT / C::BN), nj0 = (warp % (C::NT / C::BN)) * C::BN;
  {
    double2* z = reinterpret_cast<double2*>(base);
    for (int e = tid; e < (3 * slot_d + 8 * S) / 2; e += C::NTHREADS) z[e] = make_double2(0.0, 0.0);
  }
  __syncthreads();
  Mat P, T, Qm;
  P.re = base; P.im = P.re + d * S;
  T.re = base + slot_d; T.im = T.re + d * S;
  Qm.re = base + 2 * slot_d; Qm.im = Qm.re + d * S;
  slot_copy(P.re, Q, n2, tid, C::NTHREADS);
  __syncthreads();
  for (int i = 1; i < nq; i++) {
    slot_copy(Qm.re, Q + (size_t)i * slot_d, n2, tid, C::NTHREADS);
    __syncthreads();
    Acc<C::BN> acc; acc.zero();
    mm_acc<C, false>(acc, Qm, P, mi, nj0, lane);
    mm_store<C>(T, acc, d, mi, nj0, lane, NoEpi());
    __syncthreads();
    Mat t = P; P = T; T = t;
  }
  for (int e = tid; e < d * d; e += C::NTHREADS) {
    const int c = e / d, r = e - c * d;
    reinterpret_cast<double2*>(out_c128)[e] = make_double2(P.re[r * S + c], P.im[r * S + c]);
  }
}

// Preconditions shared by every qoc_shard_* entry point: one pulse and, on the general path, the two-level sweeps.  The running
// state penalty makes the costate recurrence affine: the three-call form (forward / affine / backward) carries it, the caller
// exchanging the affine terms c_p next to the S_p; the one-call phase 2 does not (allow_pen = false).  Never a silent fallback:
// anything else is QOC_ERR_UNSUPPORTED.
static int shard_guard(qoc_handle* h, bool allow_pen = true) {
  if (h->prob.batch != 1) { h->err = "time sharding handles one pulse (batch == 1)"; return QOC_ERR_INVALID; }
  if (h->nch > 1) { h->err = "time sharding with more than 8 state columns is not supported yet"; return QOC_ERR_UNSUPPORTED; }
  if (!allow_pen && has_penalty(h)) {
    h->err = "qoc_shard_phase2_device does not carry the running state penalty: use forward / affine / backward";
    return QOC_ERR_UNSUPPORTED;
  }
  if (h->gpath && !h->gs2) { h->err = "time sharding on the general path needs the two-level sweeps (nt >= 4)"; return QOC_ERR_UNSUPPORTED; }
  return QOC_OK;
}

extern "C" int qoc_shard_phase1_device(qoc_handle* h, const double* d_u, double* d_S_out, void* stream) {
  if (!h || !d_u || !d_S_out) return QOC_ERR_INVALID;
  { const int gr = shard_guard(h); if (gr != QOC_OK) return gr; }
  cudaStream_t st = (cudaStream_t)stream;
  QOC_CUDA(h, cudaSetDevice(h->prob.device));
  h->launches = 0;
  h->shard_fwd_done = false;
  int rc = launch_k1(h, d_u, true, st);
  if (rc != QOC_OK) return rc;
  if (h->gpath) {
    // S = Q_{spp-1} ... Q_0: spp - 1 single products through the batched GEMM (the chain is short: spp ~ sqrt(nt))
    const long long slot = h->slot_d;
    double* bufs[2] = {h->dQ2, h->dQ2 + slot};   // dQ2 is free once the Q_seg are in dQ (nseg >= 2 slots)
    const double* cur = h->dQ;
    for (int s = 1; s < h->spp; s++) {
      GGemm g;
      memset(&g, 0, sizeof g);
      g.d = h->prob.d; g.S = h->S; g.nb = 1; g.npairs = 1; g.alpha = 1.0;
      g.A[0] = GOp{h->dQ + (size_t)s * slot, 0}; g.B[0] = GOp{cur, 0};
      g.C = bufs[s & 1]; g.cstride = 0;
      g_gemm_launch(g, 1, st);
      h->launches++;
      cur = bufs[s & 1];
    }
    planar_to_c128_kernel<<<32, 256, 0, st>>>(cur, h->prob.d, h->S, d_S_out);
    h->launches++;
    QOC_CUDA(h, cudaGetLastError());
    h->have_u = true;
    return QOC_OK;
  }
  if (h->new_k2) {   // two-level product: group products in parallel, then G - 1 products by one CTA
    if ((rc = launch_k2(h, 5, true, nullptr, nullptr, nullptr, st, d_S_out)) != QOC_OK) return rc;
    h->have_u = true;
    return QOC_OK;
  }
  with_cfg(h->cfg, [&](auto c) {
    typedef decltype(c) C;
    const size_t smem = (size_t)(3 * h->slot_d + 8 * C::S) * 8;
    cudaFuncSetAttribute(kq_reduce_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kq_reduce_kernel<C><<<1, C::NTHREADS, smem, st>>>(h->dQ, h->nseg, h->prob.d, d_S_out);
    return 0;
  });
  h->launches += 1;
  QOC_CUDA(h, cudaGetLastError());
  h->have_u = true;
  return QOC_OK;
}

extern "C" int qoc_shard_forward_device(qoc_handle* h, const double* d_x_start, double* d_x_end, void* stream) {
  if (!h || !d_x_start) return QOC_ERR_INVALID;
  { const int gr = shard_guard(h); if (gr != QOC_OK) return gr; }
  cudaStream_t st = (cudaStream_t)stream;
  QOC_CUDA(h, cudaSetDevice(h->prob.device));
  // general path: boundary walk + segment sweeps of the two-level form, forward only from the external state
  // (with a running penalty both forms also leave the segments' affine terms c_s and sum_k L(x_k) behind)
  int rc;
  if (h->gpath) rc = gpath_sweep2(h, 1, true, false, nullptr, d_x_start, nullptr, nullptr, st);
  else {
    if (has_penalty(h)) QOC_CUDA(h, cudaMemsetAsync(h->dJpen, 0, 8, st));
    rc = launch_k2(h, 1, true, nullptr, d_x_start, nullptr, st);
    if (rc == QOC_OK && has_penalty(h)) rc = launch_k3(h, false, true, nullptr, st, 1);
  }
  if (rc != QOC_OK) return rc;
  h->shard_fwd_done = true;
  if (d_x_end)
    QOC_CUDA(h, cudaMemcpyAsync(d_x_end, h->dxf, (size_t)2 * h->prob.d * h->prob.m * 8, cudaMemcpyDeviceToDevice, st));
  return QOC_OK;
}

// sum_k L(x_k) over the nt_local + 1 local states: what the pre-pass accumulated over x_0 .. x_{nt-1} (mu already applied on the
// shared-memory path, not on the general path) plus L(x_nt)
__global__ void __launch_bounds__(256) shard_jpen_kernel(const double* Jpen, const double* xf, int d, int m, const unsigned char* pen_row,
                                                         unsigned long long row_mask, unsigned col_mask, double mu, int scale_acc,
                                                         double* out) {
  __shared__ double red[8];
  const int tid = threadIdx.x;
  double s = 0.0;
  for (int e = tid; e < d * m; e += 256) {
    const int c = e / d, r = e - c * d;
    const bool on = (pen_row ? pen_row[r] != 0 : (r < 64 && ((row_mask >> r) & 1ull))) && ((col_mask >> c) & 1u);
    if (on) s += xf[2 * e] * xf[2 * e] + xf[2 * e + 1] * xf[2 * e + 1];
  }
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += red[w];
    out[0] = (scale_acc ? mu * Jpen[0] : Jpen[0]) + mu * t;
  }
}

// Running penalty under time sharding: the local costate recurrence is affine in the costate entering from the right,
//   lambda_start = S_p' lambda_end + c_p,
// c_p being what the backward boundary walk returns for lambda_end = 0 (the segments' c_s are there since the forward call;
// dL_dx of the last local state is added by the walk itself, as for any terminal costate: src/gradient_computations.jl:47-49).
// One boundary-scan launch; also hands out the local sum_k L(x_k) over the nt_local + 1 local states.
extern "C" int qoc_shard_affine_device(qoc_handle* h, double* d_c_out, double* d_Jpen_out, void* stream) {
  if (!h || !d_c_out) return QOC_ERR_INVALID;
  { const int gr = shard_guard(h); if (gr != QOC_OK) return gr; }
  if (!has_penalty(h)) { h->err = "qoc_shard_affine_device: the problem has no running penalty"; return QOC_ERR_INVALID; }
  if (!h->shard_fwd_done || !h->states_valid) { h->err = "qoc_shard_affine_device needs qoc_shard_forward_device first"; return QOC_ERR_STALE_CACHE; }
  cudaStream_t st = (cudaStream_t)stream;
  const qoc_problem& p = h->prob;
  QOC_CUDA(h, cudaSetDevice(p.device));
  const size_t sb = (size_t)2 * p.d * p.m * 8;
  double* zero = h->dbnd + (size_t)2 * p.d * p.m;      // lambda_end = 0 (the lambda_end half of the boundary scratch)
  QOC_CUDA(h, cudaMemsetAsync(zero, 0, sb, st));
  int rc;
  if (h->gpath) rc = gpath_sweep2(h, 2, false, false, zero, nullptr, nullptr, nullptr, st, true);
  else rc = launch_k2(h, 2, false, zero, nullptr, nullptr, st);
  if (rc != QOC_OK) return rc;
  QOC_CUDA(h, cudaMemcpyAsync(d_c_out, h->dlam0, sb, cudaMemcpyDeviceToDevice, st));
  if (d_Jpen_out) {
    // sum over the local states: the pre-pass has x_0 .. x_{nt-1} (+ mu folded in or not, see below), x_nt is added here
    shard_jpen_kernel<<<1, 256, 0, st>>>(h->dJpen, h->dxf, p.d, p.m, h->gpath ? h->dpenrow : nullptr, h->row_mask64, h->col_mask, p.mu,
                                         h->gpath ? 1 : 0, d_Jpen_out);
    h->launches++;
    QOC_CUDA(h, cudaGetLastError());
  }
  return QOC_OK;
}

extern "C" int qoc_shard_backward_device(qoc_handle* h, const double* d_lambda_end, double* d_dJdu, double* d_lambda_start,
                                         void* stream) {
  if (!h || !d_lambda_end) return QOC_ERR_INVALID;
  { const int gr = shard_guard(h); if (gr != QOC_OK) return gr; }
  cudaStream_t st = (cudaStream_t)stream;
  QOC_CUDA(h, cudaSetDevice(h->prob.device));
  int rc;
  if (h->gpath) {   // backward from the external costate over the states the forward call left in HBM
    if (!h->states_valid) { h->err = "qoc_shard_backward_device needs qoc_shard_forward_device first"; return QOC_ERR_STALE_CACHE; }
    if (h->stream_jac) {
      rc = gpath_sweep2(h, 2, false, false, d_lambda_end, nullptr, nullptr, d_dJdu, st);
      if (rc == QOC_OK) rc = gpath_k1(h, h->k1_u, true, st, true, d_dJdu);
    } else
    rc = gpath_sweep2(h, 2, false, true, d_lambda_end, nullptr, nullptr, d_dJdu, st);
  } else {
    rc = launch_k2(h, 2, false, d_lambda_end, nullptr, nullptr, st);
    if (rc != QOC_OK) return rc;
    rc = launch_k3(h, true, true, d_dJdu, st);
  }
  if (rc != QOC_OK) return rc;
  if (d_lambda_start)
    QOC_CUDA(h, cudaMemcpyAsync(d_lambda_start, h->dlam0, (size_t)2 * h->prob.d * h->prob.m * 8, cudaMemcpyDeviceToDevice, st));
  return QOC_OK;
}

// Boundary algebra over the all-gathered rank propagators on the device: x_start -> h->dbnd, lambda_end -> h->dbnd + 2 d m, J.
// C_all / xs: the affine terms of the running penalty and the boundary-state scratch (NULL: no penalty, or the first of the two
// calls of the penalty route, which only needs x_start).
static int launch_shard_boundary(qoc_handle* h, const double* d_S_all, int nranks, int rank, const double* d_C_all, double* d_xs,
                                 double* d_J, cudaStream_t st) {
  const qoc_problem& p = h->prob;
  ShardBoundary q;
  memset(&q, 0, sizeof q);
  q.d = p.d; q.m = p.m; q.nranks = nranks; q.rank = rank; q.cost = p.cost; q.n = p.n;
  q.S_all = d_S_all; q.x0 = h->dx0; q.T = h->dT; q.x_start = h->dbnd; q.lam_end = h->dbnd + (size_t)2 * p.d * p.m; q.J = d_J ? d_J : h->dJ;
  q.pen = has_penalty(h) ? 1 : 0; q.C_all = d_C_all; q.xs = d_xs;
  q.pen_row = h->dpenrow; q.row_mask = h->row_mask64; q.col_mask = h->col_mask; q.mu = p.mu;
  const size_t b_smem = (size_t)2 * p.d * p.m * 16;
  if (b_smem > 40 * 1024) QOC_CUDA(h, cudaFuncSetAttribute(shard_boundary_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b_smem));
  shard_boundary_kernel<<<1, 256, b_smem, st>>>(q);
  QOC_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return QOC_OK;
}

// Phase 2 of the time-sharded evaluation in one call: boundary algebra over the all-gathered rank propagators (on the
// device, redundantly on every rank), then the local boundary scan and sweeps.  Needs a built-in cost.
extern "C" int qoc_shard_phase2_device(qoc_handle* h, const double* d_S_all, int nranks, int rank, double* d_J, double* d_dJdu,
                                       void* stream) {
  if (!h || !d_S_all || nranks <= 0 || rank < 0 || rank >= nranks) return QOC_ERR_INVALID;
  { const int gr = shard_guard(h, false); if (gr != QOC_OK) return gr; }
  if (h->prob.cost == QOC_COST_NONE) { h->err = "qoc_shard_phase2_device needs a built-in cost"; return QOC_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const qoc_problem& p = h->prob;
  QOC_CUDA(h, cudaSetDevice(p.device));
  { const int rb = launch_shard_boundary(h, d_S_all, nranks, rank, nullptr, nullptr, d_J, st); if (rb != QOC_OK) return rb; }
  const double* lam_end_bnd = h->dbnd + (size_t)2 * p.d * p.m;
  const double* x_start_bnd = h->dbnd;
  int rc;
  if (h->gpath && h->stream_jac) {
    if ((rc = gpath_sweep2(h, 4, false, false, lam_end_bnd, x_start_bnd, nullptr, d_dJdu, st)) != QOC_OK) return rc;
    return gpath_k1(h, h->k1_u, true, st, true, d_dJdu);
  }
  if (h->gpath) return gpath_sweep2(h, 4, false, true, lam_end_bnd, x_start_bnd, nullptr, d_dJdu, st);
  if (h->new_k2) {
    if ((rc = launch_k2(h, 4, false, lam_end_bnd, x_start_bnd, nullptr, st)) != QOC_OK) return rc;
  } else {
    if ((rc = launch_k2(h, 1, true, nullptr, x_start_bnd, nullptr, st)) != QOC_OK) return rc;
    if ((rc = launch_k2(h, 2, false, lam_end_bnd, nullptr, nullptr, st)) != QOC_OK) return rc;
  }
  return launch_k3(h, true, true, d_dJdu, st);
}

// Developer aid (not part of include/qoc_b200.h): runs K1 once on the cached u with clock64() stamps recorded by CTA 0
// at every hand-off between the compute warps and the service warp.  out: nslices x 16 long longs.
extern "C" int qoc_debug_k1_timeline(qoc_handle* h, long long* out, int nslices, int want_jac) {
  if (!h || !out || nslices <= 0 || !h->have_u) return QOC_ERR_INVALID;
  if (h->gpath || h->k1s_ok) { h->err = "qoc_debug_k1_timeline: k1_kernel only"; return QOC_ERR_UNSUPPORTED; }
  QOC_CUDA(h, cudaSetDevice(h->prob.device));
  QOC_CUDA(h, cudaMalloc(&h->dbg, sizeof(long long) * (16 * nslices + 4096)));
  QOC_CUDA(h, cudaMemset(h->dbg, 0, sizeof(long long) * (16 * nslices + 4096)));
  h->dbg_slices = nslices;
  h->dbg_flags = 0;
  if (want_jac >= 8) { h->dbg_flags |= 4; want_jac -= 8; }  // timing experiment: skip the U_k / dU_k stores
  if (want_jac >= 4) { h->dbg_flags |= 2; want_jac -= 4; }  // also record one stamp per compute-warp barrier
  if (want_jac >= 2) { h->dbg_flags |= 1; want_jac -= 2; }  // timing experiment: skip the inverse (results invalid)
  int rc = launch_k1(h, h->du, want_jac != 0, h->stream);
  cudaStreamSynchronize(h->stream);
  if (rc == QOC_OK) cudaMemcpy(out, h->dbg, sizeof(long long) * (16 * nslices + ((h->dbg_flags & 2) ? 4096 : 0)), cudaMemcpyDeviceToHost);
  cudaFree(h->dbg);
  h->dbg = nullptr; h->dbg_slices = 0; h->dbg_flags = 0;
  return rc;
}

// Developer aid: K3N once on the cached evaluation with clock64 stamps of CTA 0's recurrence warp 0 (4 per step).
extern "C" int qoc_debug_k3_timeline(qoc_handle* h, long long* out, int nsteps) {
  if (!h || !out || nsteps <= 0 || !h->have_u) return QOC_ERR_INVALID;
  if (h->gpath) { h->err = "qoc_debug_k3_timeline: shared-memory path only"; return QOC_ERR_UNSUPPORTED; }
  QOC_CUDA(h, cudaSetDevice(h->prob.device));
  QOC_CUDA(h, cudaMalloc(&h->dbg, sizeof(long long) * 4 * nsteps));
  QOC_CUDA(h, cudaMemset(h->dbg, 0, sizeof(long long) * 4 * nsteps));
  h->dbg_slices = nsteps;
  int rc = launch_k3(h, true, false, nullptr, h->stream);
  cudaStreamSynchronize(h->stream);
  if (rc == QOC_OK) cudaMemcpy(out, h->dbg, sizeof(long long) * 4 * nsteps, cudaMemcpyDeviceToHost);
  cudaFree(h->dbg);
  h->dbg = nullptr; h->dbg_slices = 0;
  return rc;
}

// ---- pulse parameterisation around the path ------------------------------------------------------
extern "C" int qoc_set_basis(qoc_handle* h, const double* B, int ns) {
  if (!h || !B || ns <= 0) return QOC_ERR_INVALID;
  const qoc_problem& p = h->prob;
  QOC_CUDA(h, cudaSetDevice(p.device));
  if (h->dB) { cudaFree(h->dB); cudaFree(h->dc); cudaFree(h->ddc); h->dB = h->dc = h->ddc = nullptr; }
  QOC_CUDA(h, cudaMalloc(&h->dB, (size_t)p.nt * ns * 8));
  QOC_CUDA(h, cudaMalloc(&h->dc, (size_t)ns * p.nc * p.batch * 8));
  QOC_CUDA(h, cudaMalloc(&h->ddc, (size_t)ns * p.nc * p.batch * 8));
  QOC_CUDA(h, cudaMemcpy(h->dB, B, (size_t)p.nt * ns * 8, cudaMemcpyHostToDevice));
  h->ns = ns;
  return QOC_OK;
}

extern "C" int qoc_eval_coeffs(qoc_handle* h, const double* c, double* J_out, double* dJdc_out) {
  if (!h || !c) return QOC_ERR_INVALID;
  if (!h->dB) { h->err = "qoc_eval_coeffs needs qoc_set_basis first"; return QOC_ERR_INVALID; }
  const qoc_problem& p = h->prob;
  const size_t ncoef = (size_t)h->ns * p.nc * p.batch;
  QOC_CUDA(h, cudaSetDevice(p.device));
  QOC_CUDA(h, cudaMemcpyAsync(h->dc, c, ncoef * 8, cudaMemcpyHostToDevice, h->stream));
  const long long total = (long long)p.batch * p.nt * p.nc;
  const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  basis_expand_kernel<<<grid, 256, 0, h->stream>>>(h->dB, h->dc, h->du, p.nt, h->ns, p.nc, p.batch);
  int rc = qoc_eval_device(h, h->du, nullptr, nullptr, h->stream);
  if (rc != QOC_OK) return rc;
  basis_project_kernel<<<dim3(h->ns, p.nc, p.batch), 256, 0, h->stream>>>(h->dB, h->dg, h->ddc, p.nt, h->ns, p.nc);
  h->launches += 2;
  QOC_CUDA(h, cudaGetLastError());
  if (J_out) QOC_CUDA(h, cudaMemcpyAsync(J_out, h->dJ, (size_t)p.batch * 8, cudaMemcpyDeviceToHost, h->stream));
  if (dJdc_out) QOC_CUDA(h, cudaMemcpyAsync(dJdc_out, h->ddc, ncoef * 8, cudaMemcpyDeviceToHost, h->stream));
  if ((rc = queue_mail(h)) != QOC_OK) return rc;
  QOC_CUDA(h, cudaStreamSynchronize(h->stream));
  h->have_u = false;   // the host never saw this u: a later qoc_gradient(u) has nothing to compare against
  h->states_valid = p.store_costates != 0;
  if ((rc = check_status(h)) != QOC_OK) return rc;
  if ((rc = fetch_flops(h, true)) != QOC_OK) return rc;
  if (J_out)
    for (int b = 0; b < p.batch; b++)
      if (!std::isfinite(J_out[b])) { h->err = "J is not finite"; return QOC_ERR_NOT_FINITE; }
  return QOC_OK;
}

// =====================================================================================================================
// One process, several GPUs (include/qoc_b200.h: qoc_create_sharded / qoc_sharded_eval).
//
//   QOC_SHARD_BATCH  the pulses of a batch are block-partitioned over the devices; no exchange at all.
//   QOC_SHARD_TIME   ONE pulse, contiguous time segments (the serial loops of src/gradient_computations.jl:27-29, :52-58,
//                    :65-74 cut into P pieces).  Per evaluation and device p:
//                      phase 1 (local): U_k, dU_k/du_j, rank propagator S_p
//                      publish: ONE kernel on device p stores S_p straight into slot p of every device's S_all buffer over
//                               NVLink peer mappings (no NCCL, no host hop: the message is 16 d^2 bytes), then an event
//                      phase 2 (local, after the P events): boundary state / cost / boundary costate from S_all, sweeps,
//                               gradient columns of the segment
//                      the gradient segment goes from device p straight to ITS slice of the caller's host buffer: there
//                      is no gather step because the destination is host memory anyway.
//                    Every device runs on its own stream; the host thread only enqueues.
// A device may appear several times in `devices` (virtual ranks on one GPU: how the single-GPU CI exercises this code).
// =====================================================================================================================
struct qoc_sharded {
  int n = 0, kind = 0;
  qoc_problem prob;                 // the global problem
  std::vector<qoc_handle*> h;
  std::vector<int> dev, lo, hi;     // per rank: device, unit range (slices or pulses)
  std::vector<cudaEvent_t> ev_pub, ev_t0, ev_t1;
  std::vector<double*> dS_loc, dS_all, du, dJ, dg;
  std::vector<double**> d_dst;      // per rank: device array of n pointers (slot p of every rank's S_all)
  std::vector<unsigned*> d_flags, d_done;   // per rank: n arrival flags (written by the peers), n block counters of its publish kernel
  std::vector<unsigned**> d_dstflag;        // per rank: device array of n pointers (flag p of every rank)
  // running state penalty under time sharding: second exchange of (c_p, sum_k L) records of 2 d m + 2 doubles
  bool pen = false;
  std::vector<double*> dC_loc, dC_all, dXs;
  std::vector<double**> d_dstc;     // per rank: device array of n pointers (record p of every rank's C_all)
  std::vector<cudaEvent_t> ev_pub2;
  int arrived2 = 0;
  unsigned long long arrived2_gen = 0;
  unsigned epoch = 0;               // advances by two per evaluation: epoch - 1 marks the S_p exchange, epoch the c_p exchange
  bool peer = true;                 // every pair of distinct devices has a peer mapping
  bool flags = false;               // exchange by peer stores + flags (every rank on its own device, all pairs mapped); else events:
                                    // ranks that share a device must not park a spinning kernel in front of each other's
                                    // cooperative launches
  float last_ms = 0.f;
  std::string err;
  // one host thread per rank (persistent): job hand-over and the barrier between the two halves of an evaluation
  bool threads = true;
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv;
  unsigned long long gen = 0, arrived_gen = 0;
  int arrived = 0, done = 0;
  bool quit = false;
  // low-latency hand-over for back-to-back evaluations (an optimiser loop): workers and the caller spin on these for a short
  // while (~100 us) before they fall back to the condition variable
  std::atomic<unsigned long long> gen_a{0};
  std::atomic<int> done_a{0};
  std::atomic<int> sleepers{0};
  const double* job_u = nullptr;
  double *job_J = nullptr, *job_g = nullptr;
  std::vector<int> rc_rank;
};

static thread_local std::string g_sharded_error;
static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  asm volatile("pause" ::: "memory");
#else
  std::this_thread::yield();
#endif
}

// Fused exchange: rank p stores its S_p into slot p of EVERY rank's S_all buffer through the NVLink peer mappings and then
// raises its flag there (release at system scope); nothing on the host takes part.  grid = (blocks, ranks); the last block of
// a target column to finish publishes the flag.
__global__ void shard_publish_kernel(const double* S, int n2, double* const* dst, unsigned* const* dstflag, unsigned* done,
                                     unsigned epoch) {
  double2* o = reinterpret_cast<double2*>(dst[blockIdx.y]);
  const double2* s = reinterpret_cast<const double2*>(S);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) o[i] = s[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(&done[blockIdx.y], 1u) == gridDim.x - 1) {
      done[blockIdx.y] = 0u;
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dstflag[blockIdx.y]), "r"(epoch) : "memory");
    }
  }
}
// flags only (a rank whose phase 1 failed still releases its peers)
__global__ void shard_flag_kernel(unsigned* const* dstflag, int nranks, unsigned epoch) {
  if (threadIdx.x < nranks) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dstflag[threadIdx.x]), "r"(epoch) : "memory");
}
// phase 2 of a rank starts behind this kernel: one lane per peer spins (acquire at system scope) until that peer's flag
// has reached this evaluation's epoch.  Bounded: ~4 s without progress sets status 10 instead of hanging the device.
__global__ void shard_wait_kernel(const unsigned* flags, int nranks, unsigned epoch, int* status) {
  if (threadIdx.x >= nranks) return;
  const long long t0 = clock64();
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
    if ((int)(v - epoch) >= 0) break;
    if (clock64() - t0 > 8000000000ll) { atomicExch(status, 10); break; }
    __nanosleep(200);
  }
}

extern "C" const char* qoc_sharded_last_error(const qoc_sharded* s) { return s ? s->err.c_str() : g_sharded_error.c_str(); }
extern "C" double qoc_sharded_last_ms(const qoc_sharded* s) { return s ? (double)s->last_ms : 0.0; }
extern "C" int qoc_sharded_ranks(const qoc_sharded* s) { return s ? s->n : 0; }

extern "C" int qoc_sharded_destroy(qoc_sharded* s) {
  if (!s) return QOC_OK;
  if (!s->workers.empty()) {
    { std::lock_guard<std::mutex> lk(s->mu); s->quit = true; }
    s->cv.notify_all();
    for (auto& t : s->workers) t.join();
  }
  for (int p = 0; p < s->n; p++) {
    if (!(p < (int)s->h.size() && s->h[p])) continue;   // nothing was created for this rank (its device may not even exist)
    cudaSetDevice(s->dev[p]);
    if (p < (int)s->dS_loc.size() && s->dS_loc[p]) cudaFree(s->dS_loc[p]);
    if (p < (int)s->dS_all.size() && s->dS_all[p]) cudaFree(s->dS_all[p]);
    if (p < (int)s->du.size() && s->du[p]) cudaFree(s->du[p]);
    if (p < (int)s->dJ.size() && s->dJ[p]) cudaFree(s->dJ[p]);
    if (p < (int)s->dg.size() && s->dg[p]) cudaFree(s->dg[p]);
    if (p < (int)s->d_dst.size() && s->d_dst[p]) cudaFree(s->d_dst[p]);
    if (p < (int)s->d_flags.size() && s->d_flags[p]) cudaFree(s->d_flags[p]);
    if (p < (int)s->d_done.size() && s->d_done[p]) cudaFree(s->d_done[p]);
    if (p < (int)s->d_dstflag.size() && s->d_dstflag[p]) cudaFree(s->d_dstflag[p]);
    if (p < (int)s->dC_loc.size()) { cudaFree(s->dC_loc[p]); cudaFree(s->dC_all[p]); cudaFree(s->dXs[p]); cudaFree(s->d_dstc[p]); }
    if (p < (int)s->ev_pub2.size()) cudaEventDestroy(s->ev_pub2[p]);
    if (p < (int)s->ev_pub.size()) { cudaEventDestroy(s->ev_pub[p]); cudaEventDestroy(s->ev_t0[p]); cudaEventDestroy(s->ev_t1[p]); }
    if (p < (int)s->h.size() && s->h[p]) qoc_destroy(s->h[p]);
  }
  cudaGetLastError();   // leave no stale error behind for an unrelated later call to pick up
  delete s;
  return QOC_OK;
}

#define QOC_SH(s, call)                                                                                       \
  do {                                                                                                        \
    cudaError_t e__ = (call);                                                                                 \
    if (e__ != cudaSuccess) {                                                                                 \
      char buf__[512];                                                                                        \
      snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      (s)->err = buf__;                                                                                       \
      return QOC_ERR_CUDA;                                                                                    \
    }                                                                                                         \
  } while (0)

extern "C" int qoc_create_sharded(const qoc_problem* prob, const double* A0, const double* A, const double* x0, const double* T,
                                  int n_ranks, const int* devices, int shard_kind, qoc_sharded** out) {
  if (!prob || !out || n_ranks <= 0) { g_sharded_error = "null argument or n_ranks <= 0"; return QOC_ERR_INVALID; }
  if (shard_kind != QOC_SHARD_BATCH && shard_kind != QOC_SHARD_TIME) { g_sharded_error = "unknown shard kind"; return QOC_ERR_INVALID; }
  const long long units = shard_kind == QOC_SHARD_TIME ? prob->nt : prob->batch;
  if (shard_kind == QOC_SHARD_TIME && prob->batch != 1) { g_sharded_error = "time sharding handles one pulse (batch == 1)"; return QOC_ERR_INVALID; }
  if (shard_kind == QOC_SHARD_TIME && prob->cost == QOC_COST_NONE) { g_sharded_error = "time sharding needs a built-in cost"; return QOC_ERR_INVALID; }
  if (units < n_ranks) { g_sharded_error = "fewer slices / pulses than ranks"; return QOC_ERR_DIMENSION; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_sharded_error = "no CUDA device"; return QOC_ERR_NO_DEVICE; }
  qoc_sharded* s = new qoc_sharded();
  s->n = n_ranks; s->kind = shard_kind; s->prob = *prob;
  { const char* t = getenv("QOC_SHARD_THREADS"); s->threads = !(t && t[0] == '0') && n_ranks > 1; }
  s->dev.resize(n_ranks); s->lo.resize(n_ranks); s->hi.resize(n_ranks);
  s->h.assign(n_ranks, nullptr);
  s->dS_loc.assign(n_ranks, nullptr); s->dS_all.assign(n_ranks, nullptr); s->du.assign(n_ranks, nullptr);
  s->dJ.assign(n_ranks, nullptr); s->dg.assign(n_ranks, nullptr); s->d_dst.assign(n_ranks, nullptr);
  s->d_flags.assign(n_ranks, nullptr); s->d_done.assign(n_ranks, nullptr); s->d_dstflag.assign(n_ranks, nullptr);
  auto fail = [&](int rc, const std::string& msg) { g_sharded_error = msg; qoc_sharded_destroy(s); return rc; };
  for (int p = 0; p < n_ranks; p++) {
    s->dev[p] = devices ? devices[p] : p;
    if (s->dev[p] < 0 || s->dev[p] >= ndev) return fail(QOC_ERR_INVALID, "device ordinal out of range");
    s->lo[p] = (int)((long long)p * units / n_ranks);          // the partition of sharding.py (block_partition / time_partition)
    s->hi[p] = (int)((long long)(p + 1) * units / n_ranks);
  }
  // peer mappings between every pair of distinct devices (NVLink / NVSwitch); without them the exchange falls back to
  // cudaMemcpyPeerAsync
  for (int p = 0; p < n_ranks && shard_kind == QOC_SHARD_TIME; p++)
    for (int q = 0; q < n_ranks; q++) {
      if (s->dev[p] == s->dev[q]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, s->dev[p], s->dev[q]);
      if (!can) { s->peer = false; continue; }
      cudaSetDevice(s->dev[p]);
      const cudaError_t e = cudaDeviceEnablePeerAccess(s->dev[q], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) s->peer = false;
      cudaGetLastError();
    }
  {
    bool distinct = true;
    for (int p = 0; p < n_ranks; p++)
      for (int q = 0; q < p; q++) distinct = distinct && s->dev[p] != s->dev[q];
    const char* nf = getenv("QOC_SHARD_NO_FLAGS");
    s->flags = s->peer && distinct && shard_kind == QOC_SHARD_TIME && !(nf && nf[0] == '1');
  }
  const size_t d2 = (size_t)2 * prob->d * prob->d;   // doubles of one c128 d x d matrix
  for (int p = 0; p < n_ranks; p++) {
    qoc_problem pp = *prob;
    pp.device = s->dev[p];
    if (shard_kind == QOC_SHARD_TIME) pp.nt = s->hi[p] - s->lo[p]; else pp.batch = s->hi[p] - s->lo[p];
    const int rc = qoc_create(&pp, A0, A, x0, T, &s->h[p]);
    if (rc != QOC_OK) return fail(rc, std::string("rank ") + std::to_string(p) + ": " + g_create_error);
    if (cudaSetDevice(s->dev[p]) != cudaSuccess) return fail(QOC_ERR_CUDA, "cudaSetDevice failed");
    const size_t nu = (size_t)prob->nc * (shard_kind == QOC_SHARD_TIME ? pp.nt : (size_t)prob->nt * pp.batch);
    bool ok = cudaMalloc(&s->du[p], nu * 8) == cudaSuccess && cudaMalloc(&s->dg[p], nu * 8) == cudaSuccess &&
              cudaMalloc(&s->dJ[p], (size_t)pp.batch * 8) == cudaSuccess;
    if (shard_kind == QOC_SHARD_TIME)
      ok = ok && cudaMalloc(&s->dS_loc[p], d2 * 8) == cudaSuccess && cudaMalloc(&s->dS_all[p], d2 * 8 * n_ranks) == cudaSuccess &&
           cudaMalloc(&s->d_dst[p], sizeof(double*) * n_ranks) == cudaSuccess &&
           cudaMalloc(&s->d_flags[p], 4 * n_ranks) == cudaSuccess && cudaMemset(s->d_flags[p], 0, 4 * n_ranks) == cudaSuccess &&
           cudaMalloc(&s->d_done[p], 4 * n_ranks) == cudaSuccess && cudaMemset(s->d_done[p], 0, 4 * n_ranks) == cudaSuccess &&
           cudaMalloc(&s->d_dstflag[p], sizeof(unsigned*) * n_ranks) == cudaSuccess;
    if (!ok) return fail(QOC_ERR_CUDA, "device allocation failed");
  }
  s->pen = shard_kind == QOC_SHARD_TIME && has_penalty(s->h[0]);
  if (s->pen) {
    const size_t rec = (size_t)2 * prob->d * prob->m + 2;
    s->dC_loc.assign(n_ranks, nullptr); s->dC_all.assign(n_ranks, nullptr); s->dXs.assign(n_ranks, nullptr); s->d_dstc.assign(n_ranks, nullptr);
    for (int p = 0; p < n_ranks; p++) {
      if (cudaSetDevice(s->dev[p]) != cudaSuccess) return fail(QOC_ERR_CUDA, "cudaSetDevice failed");
      if (cudaMalloc(&s->dC_loc[p], rec * 8) != cudaSuccess || cudaMalloc(&s->dC_all[p], rec * 8 * n_ranks) != cudaSuccess ||
          cudaMalloc(&s->dXs[p], rec * 8 * n_ranks) != cudaSuccess || cudaMalloc(&s->d_dstc[p], sizeof(double*) * n_ranks) != cudaSuccess)
        return fail(QOC_ERR_CUDA, "device allocation failed");
    }
    s->ev_pub2.resize(n_ranks);
    for (int p = 0; p < n_ranks; p++) {
      cudaSetDevice(s->dev[p]);
      cudaEventCreateWithFlags(&s->ev_pub2[p], cudaEventDisableTiming);
      std::vector<double*> dst(n_ranks);
      for (int q = 0; q < n_ranks; q++) dst[q] = s->dC_all[q] + (size_t)p * rec;
      if (cudaMemcpy(s->d_dstc[p], dst.data(), sizeof(double*) * n_ranks, cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(QOC_ERR_CUDA, "upload of the peer table failed");
    }
  }
  s->ev_pub.resize(n_ranks); s->ev_t0.resize(n_ranks); s->ev_t1.resize(n_ranks);
  for (int p = 0; p < n_ranks; p++) {
    cudaSetDevice(s->dev[p]);
    cudaEventCreateWithFlags(&s->ev_pub[p], cudaEventDisableTiming);
    cudaEventCreate(&s->ev_t0[p]);
    cudaEventCreate(&s->ev_t1[p]);
    if (shard_kind == QOC_SHARD_TIME) {
      std::vector<double*> dst(n_ranks);
      for (int q = 0; q < n_ranks; q++) dst[q] = s->dS_all[q] + (size_t)p * d2;
      std::vector<unsigned*> df(n_ranks);
      for (int q = 0; q < n_ranks; q++) df[q] = s->d_flags[q] + p;
      if (cudaMemcpy(s->d_dst[p], dst.data(), sizeof(double*) * n_ranks, cudaMemcpyHostToDevice) != cudaSuccess ||
          cudaMemcpy(s->d_dstflag[p], df.data(), sizeof(unsigned*) * n_ranks, cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(QOC_ERR_CUDA, "upload of the peer table failed");
    }
  }
  *out = s;
  return QOC_OK;
}

extern "C" int qoc_sharded_set_order(qoc_sharded* s, int order) {
  if (!s) return QOC_ERR_INVALID;
  for (int p = 0; p < s->n; p++) {
    const int rc = qoc_set_order(s->h[p], order);
    if (rc != QOC_OK) { s->err = s->h[p]->err; return rc; }
  }
  s->prob.order = order;
  return QOC_OK;
}

// ---- per-rank halves of an evaluation (each runs on the rank's own host thread: a stream's launch queue is finite, and one
//      thread issuing the ~1000 launches of a long general-path phase 1 for rank 0 would not reach rank 1 before rank 0's
//      GPU had drained most of them -- measured: 2 GPUs, d = 64, Nt = 1e5: 317 ms single-threaded vs 217 ms one process per GPU) ----
static int sharded_rank_A(qoc_sharded* s, int p, const double* u) {
  const qoc_problem& gp = s->prob;
  qoc_handle* h = s->h[p];
  auto sub = [&](int rc) { return rc; };
  int rc;
  QOC_CUDA(h, cudaSetDevice(s->dev[p]));
  QOC_CUDA(h, cudaEventRecord(s->ev_t0[p], h->stream));
  if (s->kind == QOC_SHARD_BATCH) {
    const size_t per = (size_t)gp.nc * gp.nt, n = (size_t)(s->hi[p] - s->lo[p]);
    QOC_CUDA(h, cudaMemcpyAsync(s->du[p], u + (size_t)s->lo[p] * per, n * per * 8, cudaMemcpyHostToDevice, h->stream));
    return sub(qoc_eval_device(h, s->du[p], s->dJ[p], s->dg[p], h->stream));
  }
  const size_t d2 = (size_t)2 * gp.d * gp.d;
  const size_t n = (size_t)(s->hi[p] - s->lo[p]) * gp.nc;
  QOC_CUDA(h, cudaMemcpyAsync(s->du[p], u + (size_t)s->lo[p] * gp.nc, n * 8, cudaMemcpyHostToDevice, h->stream));
  if ((rc = qoc_shard_phase1_device(h, s->du[p], s->dS_loc[p], h->stream)) != QOC_OK) return rc;
  if (s->peer) {
    int bx = (int)((d2 / 2 + 255) / 256);
    if (bx > 16) bx = 16;
    shard_publish_kernel<<<dim3(bx, s->n), 256, 0, h->stream>>>(s->dS_loc[p], (int)(d2 / 2), s->d_dst[p], s->d_dstflag[p], s->d_done[p],
                                                              s->epoch - 1);
    QOC_CUDA(h, cudaGetLastError());
  } else {
    for (int q = 0; q < s->n; q++)
      QOC_CUDA(h, cudaMemcpyPeerAsync(s->dS_all[q] + (size_t)p * d2, s->dev[q], s->dS_loc[p], s->dev[p], d2 * 8, h->stream));
  }
  if (!s->flags) QOC_CUDA(h, cudaEventRecord(s->ev_pub[p], h->stream));
  return QOC_OK;
}

// running penalty, between the halves: once every S_q has landed, x_start, the local forward sweep, the affine term c_p of the
// local costate recurrence and the local sum of L; (c_p, sum L) published to every rank like the S_p were
static int sharded_rank_B1(qoc_sharded* s, int p) {
  const qoc_problem& gp = s->prob;
  qoc_handle* h = s->h[p];
  int rc;
  QOC_CUDA(h, cudaSetDevice(s->dev[p]));
  if (s->flags) {
    shard_wait_kernel<<<1, 32, 0, h->stream>>>(s->d_flags[p], s->n, s->epoch - 1, h->dstatus);
    QOC_CUDA(h, cudaGetLastError());
  } else
  for (int q = 0; q < s->n; q++)
    if (q != p) QOC_CUDA(h, cudaStreamWaitEvent(h->stream, s->ev_pub[q], 0));
  if ((rc = launch_shard_boundary(h, s->dS_all[p], s->n, p, nullptr, nullptr, s->dJ[p], h->stream)) != QOC_OK) return rc;
  if ((rc = qoc_shard_forward_device(h, h->dbnd, nullptr, h->stream)) != QOC_OK) return rc;
  const size_t rec = (size_t)2 * gp.d * gp.m + 2;
  if ((rc = qoc_shard_affine_device(h, s->dC_loc[p], s->dC_loc[p] + rec - 2, h->stream)) != QOC_OK) return rc;
  if (s->peer) {
    shard_publish_kernel<<<dim3(1, s->n), 256, 0, h->stream>>>(s->dC_loc[p], (int)(rec / 2), s->d_dstc[p], s->d_dstflag[p], s->d_done[p],
                                                             s->epoch);
    QOC_CUDA(h, cudaGetLastError());
  } else {
    for (int q = 0; q < s->n; q++)
      QOC_CUDA(h, cudaMemcpyPeerAsync(s->dC_all[q] + (size_t)p * rec, s->dev[q], s->dC_loc[p], s->dev[p], rec * 8, h->stream));
  }
  if (!s->flags) QOC_CUDA(h, cudaEventRecord(s->ev_pub2[p], h->stream));
  return QOC_OK;
}

// second half: (time sharding) phase 2 once every S_q has landed; results straight to the caller's host arrays; synchronise
static int sharded_rank_B(qoc_sharded* s, int p, double* J_out, double* dJdu_out) {
  const qoc_problem& gp = s->prob;
  qoc_handle* h = s->h[p];
  int rc;
  QOC_CUDA(h, cudaSetDevice(s->dev[p]));
  if (s->kind == QOC_SHARD_BATCH) {
    const size_t per = (size_t)gp.nc * gp.nt, n = (size_t)(s->hi[p] - s->lo[p]);
    if (J_out) QOC_CUDA(h, cudaMemcpyAsync(J_out + s->lo[p], s->dJ[p], n * 8, cudaMemcpyDeviceToHost, h->stream));
    if (dJdu_out) QOC_CUDA(h, cudaMemcpyAsync(dJdu_out + (size_t)s->lo[p] * per, s->dg[p], n * per * 8, cudaMemcpyDeviceToHost, h->stream));
  } else {
    const size_t n = (size_t)(s->hi[p] - s->lo[p]) * gp.nc;
    if (s->pen) {
      // the c_p of every rank have landed (second exchange): boundary costates through (S_q', c_q), J with the partial sums of L
      if (s->flags) {
        shard_wait_kernel<<<1, 32, 0, h->stream>>>(s->d_flags[p], s->n, s->epoch, h->dstatus);
        QOC_CUDA(h, cudaGetLastError());
      } else
      for (int q = 0; q < s->n; q++)
        if (q != p) QOC_CUDA(h, cudaStreamWaitEvent(h->stream, s->ev_pub2[q], 0));
      if ((rc = launch_shard_boundary(h, s->dS_all[p], s->n, p, s->dC_all[p], s->dXs[p], s->dJ[p], h->stream)) != QOC_OK) return rc;
      if ((rc = qoc_shard_backward_device(h, h->dbnd + (size_t)2 * gp.d * gp.m, s->dg[p], nullptr, h->stream)) != QOC_OK) return rc;
    } else {
    if (s->flags) {   // behind a kernel that spins on the peers' flags: no event, no host rendezvous
      shard_wait_kernel<<<1, 32, 0, h->stream>>>(s->d_flags[p], s->n, s->epoch - 1, h->dstatus);
      QOC_CUDA(h, cudaGetLastError());
    } else
    for (int q = 0; q < s->n; q++)
      if (q != p) QOC_CUDA(h, cudaStreamWaitEvent(h->stream, s->ev_pub[q], 0));
    if ((rc = qoc_shard_phase2_device(h, s->dS_all[p], s->n, p, s->dJ[p], s->dg[p], h->stream)) != QOC_OK) return rc;
    }
    if (dJdu_out) QOC_CUDA(h, cudaMemcpyAsync(dJdu_out + (size_t)s->lo[p] * gp.nc, s->dg[p], n * 8, cudaMemcpyDeviceToHost, h->stream));
    if (p == 0 && J_out) QOC_CUDA(h, cudaMemcpyAsync(J_out, s->dJ[0], 8, cudaMemcpyDeviceToHost, h->stream));
  }
  QOC_CUDA(h, cudaEventRecord(s->ev_t1[p], h->stream));
  if ((rc = queue_mail(h)) != QOC_OK) return rc;
  QOC_CUDA(h, cudaStreamSynchronize(h->stream));
  rc = check_status(h);
  h->mail_valid = false;
  return rc;
}

static void sharded_worker(qoc_sharded* s, int p) {
  unsigned long long seen = 0;
  for (;;) {
    {
      bool got = false;
      for (int spin = 0; spin < 20000 && !got; spin++) {   // ~100 us of polling
        got = s->gen_a.load(std::memory_order_acquire) != seen;
        if (!got) cpu_relax();
      }
      if (!got) {
        std::unique_lock<std::mutex> lk(s->mu);
        s->sleepers.fetch_add(1);
        s->cv.wait(lk, [&] { return s->quit || s->gen != seen; });
        s->sleepers.fetch_sub(1);
        if (s->quit) return;
      }
      if (s->quit) return;
      seen = s->gen_a.load(std::memory_order_acquire);
    }
    int rc = sharded_rank_A(s, p, s->job_u);
    if (rc != QOC_OK && s->flags) {   // release the peers that will spin on this rank's flag
      cudaSetDevice(s->dev[p]);
      shard_flag_kernel<<<1, 32, 0, s->h[p]->stream>>>(s->d_dstflag[p], s->n, s->epoch);
    }
    bool any_bad = rc != QOC_OK;
    if (rc != QOC_OK) { std::lock_guard<std::mutex> lk(s->mu); s->rc_rank[p] = rc; }
    if (!(s->flags || s->kind == QOC_SHARD_BATCH)) {
      // event path: host barrier, every ev_pub of this evaluation is recorded before anybody waits on one
      std::unique_lock<std::mutex> lk(s->mu);
      if (++s->arrived == s->n) { s->arrived_gen = seen; s->cv.notify_all(); }
      else s->cv.wait(lk, [&] { return s->arrived_gen == seen; });
      for (int q = 0; q < s->n; q++) any_bad |= s->rc_rank[q] != QOC_OK;
    }
    if (s->pen) {
      rc = any_bad ? QOC_OK : sharded_rank_B1(s, p);
      if ((any_bad || rc != QOC_OK) && s->flags) {   // release the peers that will spin on this rank's second flag
        cudaSetDevice(s->dev[p]);
        shard_flag_kernel<<<1, 32, 0, s->h[p]->stream>>>(s->d_dstflag[p], s->n, s->epoch);
      }
      if (rc != QOC_OK) { std::lock_guard<std::mutex> lk(s->mu); s->rc_rank[p] = rc; any_bad = true; }
      if (!s->flags) {
        std::unique_lock<std::mutex> lk(s->mu);
        if (++s->arrived2 == s->n) { s->arrived2_gen = seen; s->cv.notify_all(); }
        else s->cv.wait(lk, [&] { return s->arrived2_gen == seen; });
        for (int q = 0; q < s->n; q++) any_bad |= s->rc_rank[q] != QOC_OK;
      }
    }
    if (!any_bad) {
      rc = sharded_rank_B(s, p, s->job_J, s->job_g);
      if (rc != QOC_OK) { std::lock_guard<std::mutex> lk(s->mu); s->rc_rank[p] = rc; }
    }
    if (s->done_a.fetch_add(1, std::memory_order_acq_rel) + 1 == s->n) {
      std::lock_guard<std::mutex> lk(s->mu);
      s->done = s->n;
      s->cv.notify_all();
    }
  }
}

extern "C" int qoc_sharded_eval(qoc_sharded* s, const double* u, double* J_out, double* dJdu_out) {
  if (!s || !u) return QOC_ERR_INVALID;
  const qoc_problem& gp = s->prob;
  const int P = s->n;
  s->rc_rank.assign(P, QOC_OK);
  s->epoch += 2;
  if (s->threads) {
    if (s->workers.empty())
      for (int p = 0; p < P; p++) s->workers.emplace_back(sharded_worker, s, p);
    {
      std::lock_guard<std::mutex> lk(s->mu);
      s->job_u = u; s->job_J = J_out; s->job_g = dJdu_out;
      s->arrived = 0; s->arrived2 = 0; s->done = 0;
      s->done_a.store(0, std::memory_order_relaxed);
      s->gen++;
      s->gen_a.store(s->gen, std::memory_order_release);
    }
    if (s->sleepers.load() > 0) s->cv.notify_all();
    {
      bool fin = false;
      for (int spin = 0; spin < 2000000 && !fin; spin++) {   // a few ms of polling, then sleep
        fin = s->done_a.load(std::memory_order_acquire) == P;
        if (!fin) cpu_relax();
      }
      if (!fin) {
        std::unique_lock<std::mutex> lk(s->mu);
        s->cv.wait(lk, [&] { return s->done == P; });
      }
    }
  } else {
    bool bad = false;
    for (int p = 0; p < P; p++) {
      s->rc_rank[p] = sharded_rank_A(s, p, u);
      bad |= s->rc_rank[p] != QOC_OK;
    }
    for (int p = 0; p < P && !bad && s->pen; p++) { s->rc_rank[p] = sharded_rank_B1(s, p); bad |= s->rc_rank[p] != QOC_OK; }
    for (int p = 0; p < P && !bad; p++) s->rc_rank[p] = sharded_rank_B(s, p, J_out, dJdu_out);
  }
  for (int p = 0; p < P; p++)
    if (s->rc_rank[p] != QOC_OK) { s->err = std::string("rank ") + std::to_string(p) + ": " + s->h[p]->err; return s->rc_rank[p]; }
  float worst = 0.f;
  for (int p = 0; p < P; p++) {
    float ms = 0.f;
    cudaSetDevice(s->dev[p]);
    if (cudaEventElapsedTime(&ms, s->ev_t0[p], s->ev_t1[p]) == cudaSuccess && ms > worst) worst = ms;
  }
  cudaGetLastError();
  s->last_ms = worst;   // device time of the slowest rank (H2D of its inputs to D2H of its results)
  const int nJ = s->kind == QOC_SHARD_BATCH ? gp.batch : 1;
  if (J_out)
    for (int b = 0; b < nJ; b++)
      if (!std::isfinite(J_out[b])) { s->err = "J is not finite"; return QOC_ERR_NOT_FINITE; }
  return QOC_OK;
}

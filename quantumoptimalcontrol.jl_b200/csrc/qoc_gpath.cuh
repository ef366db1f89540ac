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
#include "qoc_cost.cuh"

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
  int gj_k0, gj_nb;   // gj_nb > 0: rank-nb update of the blocked Gauss-Jordan inverse (k range [gj_k0, gj_k0 + gj_nb), D[0] = cur)
  // up to two more outputs from the same product P = sum_p A_p B_p and the same D operands (the linear combinations of the
  // Pade / Frechet program that used to be launches of their own):  Cx_k = alphax_k P + sum_q betax_k[q] D_q + gammax_k I
  int nextra;
  double* Cx[2];
  long long cxstride[2];
  double alphax[2], betax[2][3], gammax[2];
};

#ifndef QOC_GEMM_KC64
#define QOC_GEMM_KC64 32
#endif
#ifndef QOC_GEMM_NST64
#define QOC_GEMM_NST64 3
#endif
#ifndef QOC_GEMM_MINBX
#define QOC_GEMM_MINBX 2    // ... of its extra-output instantiation (137 registers uncapped: one CTA per SM)
#endif
#ifndef QOC_GEMM_MINBX40
#define QOC_GEMM_MINBX40 3  // extra-output instantiations of the 40 / 48 tiles: 167 -> 128 and 184 -> 166 registers, i.e. three and two
#endif                      // CTAs per SM like their plain forms instead of two and ONE (d = 48: 10.8 -> 9.7 ms, d = 96: 14.7 -> 13.1,
#ifndef QOC_GEMM_MINBX48    // cavity-40: 4.30 -> 4.11)
#define QOC_GEMM_MINBX48 2
#endif
#ifndef QOC_GEMM_NSTS
#define QOC_GEMM_NSTS 3     // ring depth of the small tiles (2 measured worse: 84 registers, d = 32 9.1 -> 10.3 ms)
#endif
#ifndef QOC_GEMM_MINB
#define QOC_GEMM_MINB 4     // __launch_bounds__ minimum CTAs per SM of the 32 x 32 tile (8 warps): 64 registers instead of 70, four
#endif                      // CTAs per SM instead of three (d = 32, 8000 slices: 9.1 -> 8.2 ms; on the 40 / 48 tiles it only hurt)
#ifndef QOC_GEMM_KCS
#define QOC_GEMM_KCS 16
#endif
#ifndef QOC_3M
#error "qoc_gpath.cuh assumes the 3M complex product (Acc carries T1/T2/T3)"
#endif

__device__ __forceinline__ unsigned g_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// 16-byte asynchronous global -> shared copy (SASS LDGSTS); bytes beyond src_bytes (0, 8 or 16) are zero-filled
__device__ __forceinline__ void g_cp_async16(void* sdst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(g_smem_u32(sdst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void g_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void g_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// C = alpha (A1 B1 [+ A2 B2 ...]) + sum_q beta_q D_q + gamma I  for a batch of slices (blockIdx.z), 3M DMMA tiles.
// CTA tile (8 WM NWM) x (8 WN NWN): NWM x NWN warps, each WM x WN tiles of 8 x 8.  The operands stream through a
// three-stage cp.async ring in 16-wide k chunks (row strides = 4 mod 8 doubles: conflict-free fragment loads), one
// __syncthreads per chunk; edges are zero-filled by the copy itself (src-size form), 8 x 8 tiles and k-steps that lie
// entirely outside the d x d matrix are skipped (warp-uniform predicates).
// Round 1's kernel staged synchronously through registers (two barriers per chunk, nothing in flight while the tile loop
// ran): 0.41 of the FP64 peak at best.  gj_nb > 0 selects the rank-nb update of the blocked Gauss-Jordan inverse below.
template <int WM, int WN, int NWM, int NWN>
struct GemmShape {
  // k-chunk: 32 wide for the 64 x 64 tile (3 x 72 KB of ring: one CTA per SM anyway, half as many barriers per flop), 16 otherwise
  static constexpr int NTH = 32 * NWM * NWN, TM = 8 * WM * NWM, TN = 8 * WN * NWN, KC = (WM * NWM == 8 && WN * NWN == 8) ? QOC_GEMM_KC64 : QOC_GEMM_KCS,
                       NST = (WM * NWM == 8 && WN * NWN == 8) ? QOC_GEMM_NST64 : QOC_GEMM_NSTS;
  static constexpr int AS = KC + 4, BS = TN + 4;
  static constexpr int A_PLANE = TM * AS, B_PLANE = KC * BS, STAGE = 2 * A_PLANE + 2 * B_PLANE;   // doubles
  static constexpr size_t SMEM = (size_t)NST * STAGE * 8;
};

// PERSISTENT: the grid is one wave of CTAs (SM count x occupancy); CTA b walks the work items b, b + grid, ... (item =
// one CTA tile of one slice) and its cp.async ring runs straight THROUGH the item boundaries: while the last chunks of an
// item are multiplied and its epilogue (D reads, C stores) runs, the first chunks of the next item are already in flight.
// One-CTA-per-tile launches exposed a prologue (first chunk: an L2 / HBM round trip) and an epilogue per tile with
// nothing to overlap them when only one CTA fits an SM (the 64 x 64 tile: 16 warps x ~100 registers): at d = 64, where a
// tile is only four chunks long, that was half the kernel.
template <int WM, int WN, int NWM, int NWN, bool XTRA>   // XTRA: the epilogue also writes the extra outputs (more registers)
__global__ void __launch_bounds__(32 * NWM * NWN, (NWM * NWN == 8) ? (XTRA ? QOC_GEMM_MINBX : QOC_GEMM_MINB) : (XTRA && NWM * NWN == 5) ? QOC_GEMM_MINBX40 : (XTRA && NWM * NWN == 6) ? QOC_GEMM_MINBX48 : 1) g_gemm2_kernel(GGemm g, int ntm, int ntn, int nitems) {
  typedef GemmShape<WM, WN, NWM, NWN> G;
  constexpr int NTH = G::NTH, TM = G::TM, TN = G::TN, KC = G::KC, NST = G::NST, AS = G::AS, BS = G::BS;
  extern __shared__ __align__(16) unsigned char gsm_raw[];
  double* sm = reinterpret_cast<double*>(gsm_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / NWN, wn = warp - wm * NWN;
  const int d = g.d, S = g.S, plane = d * S;
  const int gq = lane >> 2, q4 = lane & 3;
  const bool gj = g.gj_nb > 0;
  const int kbeg = gj ? g.gj_k0 : 0;
  const int kend = gj ? ((g.gj_k0 + g.gj_nb < d) ? g.gj_k0 + g.gj_nb : d) : d;
  const int nkc = (kend - kbeg + KC - 1) / KC;
  const int total = g.npairs * nkc;   // chunks per item
  const int tps = ntm * ntn;          // tiles per slice

  // Ring hand-over by mbarriers instead of a CTA-wide barrier per chunk: full[s] completes when every thread's cp.async of the
  // chunk in stage s has landed (cp.async.mbarrier.arrive.noinc), done[s] when every thread has finished reading it.  A warp
  // consumes chunk c as soon as it is there and refills the stage of chunk c - 1 afterwards, by which time everybody has
  // long left it: warps drift by up to a chunk instead of meeting at a barrier 2 x (d / 16) times per product.
  __shared__ __align__(8) unsigned long long ring_bar[2 * NST];
  unsigned long long* full = ring_bar;
  unsigned long long* done = ring_bar + NST;
  if (tid == 0) {
    for (int i = 0; i < NST; i++) { mbar_init(full + i, NTH); mbar_init(done + i, NTH); }
    fence_mbar_init();
  }
  __syncthreads();
  // producer cursor: (item, chunk) of the next chunk to put in flight, and its running number (stage = number % NST)
  int pw = blockIdx.x, pc = 0, pgc = 0;
  auto issue_next = [&]() {
    if (pw < nitems) {
      const int pstage = pgc % NST, use = pgc / NST;
      if (use >= 1) mbar_wait(done + pstage, (unsigned)((use - 1) & 1));
      const int s = pw / tps, r = pw - s * tps, tm = r / ntn, tn = r - tm * ntn;
      const int p = pc / nkc, kc = pc - p * nkc, kb = kbeg + kc * KC;
      const double* Ag = g.A[p].p + g.A[p].off(s);
      const double* Bg = g.B[p].p + g.B[p].off(s);
      double* st = sm + (size_t)pstage * G::STAGE;
      for (int e = tid; e < TM * (KC / 2); e += NTH) {          // A tile: TM rows x KC columns, both planes
        const int rr = e / (KC / 2), cc = (e - rr * (KC / 2)) * 2;
        const int gr = tm * TM + rr, gc = kb + cc;
        int bytes = (gr < d) ? (kend - gc) * 8 : 0;
        bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
        const double* src = bytes ? Ag + (size_t)gr * S + gc : Ag;
        g_cp_async16(st + rr * AS + cc, src, bytes);
        g_cp_async16(st + G::A_PLANE + rr * AS + cc, bytes ? src + plane : Ag, bytes);
      }
      double* sb = st + 2 * G::A_PLANE;
      for (int e = tid; e < KC * (TN / 2); e += NTH) {          // B tile: KC rows x TN columns, both planes
        const int rr = e / (TN / 2), cc = (e - rr * (TN / 2)) * 2;
        const int gr = kb + rr, gc = tn * TN + cc;
        const int bytes = (gr < kend && gc < S) ? 16 : 0;
        const double* src = bytes ? Bg + (size_t)gr * S + gc : Bg;
        g_cp_async16(sb + rr * BS + cc, src, bytes);
        g_cp_async16(sb + G::B_PLANE + rr * BS + cc, bytes ? src + plane : Bg, bytes);
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(full + pstage)) : "memory");
      pgc++;
      if (++pc == total) { pc = 0; pw += gridDim.x; }
    }
  };

#pragma unroll
  for (int c = 0; c < NST - 1; c++) issue_next();
  int cgc = 0;   // consumer: running chunk number
  for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
    const int s = w / tps, r = w - s * tps, tm = r / ntn, tn = r - tm * ntn;
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
    const bool work = rowv[0] && colv[0];
    bool fullw = true;   // every 8 x 8 tile of this warp lies inside the matrix
#pragma unroll
    for (int a = 0; a < WM; a++) fullw = fullw && rowv[a];
#pragma unroll
    for (int b = 0; b < WN; b++) fullw = fullw && colv[b];
    for (int it = 0; it < total; it++) {
      const int cstage = cgc % NST;
      mbar_wait(full + cstage, (unsigned)((cgc / NST) & 1));
      if (work) {
        const double* st = sm + (size_t)cstage * G::STAGE;
        const int kc = it % nkc;
        const int kleft = kend - (kbeg + kc * KC);
        const int ksmax = kleft >= KC ? KC / 4 : (kleft + 3) / 4;
        const double* are = st + (wm * WM * 8 + gq) * AS + q4;
        const double* aim = are + G::A_PLANE;
        const double* bre = st + 2 * G::A_PLANE + q4 * BS + wn * WN * 8 + gq;
        const double* bim = bre + G::B_PLANE;
        if (fullw && ksmax == KC / 4) {
          // interior warp tile, whole chunk: straight-line code, no predicate between the DMMAs (the predicated form below
          // spends ~9 instructions per DMMA on warp-uniform branches, WARPSYNC and NOP padding: ncu, d = 256)
#pragma unroll
          for (int ks = 0; ks < KC / 4; ks++) {
            double ar[WM], ai[WM], as[WM], br[WN], bi[WN], bs[WN];
#pragma unroll
            for (int a = 0; a < WM; a++) { ar[a] = are[a * 8 * AS + ks * 4]; ai[a] = aim[a * 8 * AS + ks * 4]; }
#pragma unroll
            for (int b = 0; b < WN; b++) { br[b] = bre[ks * 4 * BS + b * 8]; bi[b] = bim[ks * 4 * BS + b * 8]; }
#pragma unroll
            for (int a = 0; a < WM; a++) as[a] = ar[a] + ai[a];
#pragma unroll
            for (int b = 0; b < WN; b++) bs[b] = br[b] + bi[b];
#pragma unroll
            for (int b = 0; b < WN; b++)
#pragma unroll
              for (int a = 0; a < WM; a++) {
                dmma(T1[a][b][0], T1[a][b][1], ar[a], br[b]);
                dmma(T2[a][b][0], T2[a][b][1], ai[a], bi[b]);
                dmma(T3[a][b][0], T3[a][b][1], as[a], bs[b]);
              }
          }
        } else {
#pragma unroll
        for (int ks = 0; ks < KC / 4; ks++) {
          if (ks < ksmax) {
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
      mbar_arrive(done + cstage);
      cgc++;
      issue_next();
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
          const double p0 = T1[a][b][0] - T2[a][b][0], p1 = T1[a][b][1] - T2[a][b][1];
          const double q0 = (T3[a][b][0] - T1[a][b][0]) - T2[a][b][0], q1 = (T3[a][b][1] - T1[a][b][1]) - T2[a][b][1];
          double r0 = g.alpha * p0, r1 = g.alpha * p1, i0 = g.alpha * q0, i1 = g.alpha * q1;
          const size_t o = (size_t)row * S + col;
          if (gj) {
            // blocked Gauss-Jordan step on the panel K = [kbeg, kend): the K columns of `cur` already hold the transformation
            // G; every other column becomes (cur with its K rows zeroed) + G cur[K, :]
            const double* Dg = g.D[0].p + g.D[0].off(s);
            const double2 u = *reinterpret_cast<const double2*>(Dg + o), v = *reinterpret_cast<const double2*>(Dg + plane + o);
            if (col >= kbeg && col < kend) { r0 = u.x; r1 = u.y; i0 = v.x; i1 = v.y; }
            else if (!(row >= kbeg && row < kend)) { r0 += u.x; r1 += u.y; i0 += v.x; i1 += v.y; }
          } else {
            double2 du[3], dv[3];
#pragma unroll
            for (int q = 0; q < 3; q++) {
              du[q] = make_double2(0.0, 0.0); dv[q] = du[q];
              if (q < g.nadd) {
                const double* Dg = g.D[q].p + g.D[q].off(s);
                du[q] = *reinterpret_cast<const double2*>(Dg + o); dv[q] = *reinterpret_cast<const double2*>(Dg + plane + o);
                r0 = fma(g.beta[q], du[q].x, r0); r1 = fma(g.beta[q], du[q].y, r1);
                i0 = fma(g.beta[q], dv[q].x, i0); i1 = fma(g.beta[q], dv[q].y, i1);
              }
            }
            if (row == col) r0 += g.gamma;
            if (row == col + 1) r1 += g.gamma;
            for (int x = 0; XTRA && x < g.nextra; x++) {
              double xr0 = g.alphax[x] * p0, xr1 = g.alphax[x] * p1, xi0 = g.alphax[x] * q0, xi1 = g.alphax[x] * q1;
#pragma unroll
              for (int q = 0; q < 3; q++)
                if (q < g.nadd) {
                  xr0 = fma(g.betax[x][q], du[q].x, xr0); xr1 = fma(g.betax[x][q], du[q].y, xr1);
                  xi0 = fma(g.betax[x][q], dv[q].x, xi0); xi1 = fma(g.betax[x][q], dv[q].y, xi1);
                }
              if (row == col) xr0 += g.gammax[x];
              if (row == col + 1) xr1 += g.gammax[x];
              if (col + 1 >= d) { xr1 = 0.0; xi1 = 0.0; }
              double* Xg = g.Cx[x] + (long long)s * g.cxstride[x];
              *reinterpret_cast<double2*>(Xg + o) = make_double2(xr0, xr1);
              *reinterpret_cast<double2*>(Xg + plane + o) = make_double2(xi0, xi1);
            }
          }
          if (col + 1 >= d) { r1 = 0.0; i1 = 0.0; }
          *reinterpret_cast<double2*>(Cg + o) = make_double2(r0, r1);
          *reinterpret_cast<double2*>(Cg + plane + o) = make_double2(i0, i1);
        }
      }
    }
  }
  g_cp_async_wait<0>();
}

// C = sum_q beta_q D_q + gamma I over whole planar slots (no product): a streaming elementwise pass, 16-byte accesses
__global__ void __launch_bounds__(256) g_lin_kernel(GGemm g) {
  const int s = blockIdx.y;
  const int d = g.d, S = g.S, plane = d * S, n2 = plane / 2, S2 = S / 2;
  double* Cg = g.C + (g.cinner > 0 ? (long long)(s / g.cinner) * g.cstride2 + (long long)(s % g.cinner) * g.cstride
                                   : (long long)s * g.cstride);
  const double* Dp[3];
  for (int q = 0; q < 3; q++) Dp[q] = q < g.nadd ? g.D[q].p + g.D[q].off(s) : nullptr;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += gridDim.x * blockDim.x) {
    double2 r = make_double2(0.0, 0.0), i = r;
    for (int q = 0; q < g.nadd; q++) {
      const double2 u = reinterpret_cast<const double2*>(Dp[q])[e], v = reinterpret_cast<const double2*>(Dp[q] + plane)[e];
      r.x = fma(g.beta[q], u.x, r.x); r.y = fma(g.beta[q], u.y, r.y);
      i.x = fma(g.beta[q], v.x, i.x); i.y = fma(g.beta[q], v.y, i.y);
    }
    if (g.gamma != 0.0) {
      const int row = e / S2, col = 2 * (e - row * S2);
      if (row == col) r.x += g.gamma;
      if (row == col + 1) r.y += g.gamma;
    }
    reinterpret_cast<double2*>(Cg)[e] = r;
    reinterpret_cast<double2*>(Cg + plane)[e] = i;
  }
}

// host-side launch helper: the CTA tile that wastes least of the d x d matrix (ties: the larger tile)
//   32 x 32 (8 warps of 1 x 2 tiles) | 40 x 40 (5 warps of 1 x 5) | 48 x 48 (6 warps of 1 x 6) | 64 x 64 (16 warps of 2 x 2) |
//   80 x 80 (10 warps of 1 x 10)
template <int WM, int WN, int NWM, int NWN, bool XTRA>
static inline void g_gemm2_launch_x(const GGemm& g, int nb, cudaStream_t st) {
  typedef GemmShape<WM, WN, NWM, NWN> G;
  static int wave_of[64] = {0};   // CTAs in one wave of this instantiation, per device (function attributes are per device)
  int dev = 0;
  cudaGetDevice(&dev);
  int& wave = wave_of[dev & 63];
  if (!wave) {
    cudaFuncSetAttribute(g_gemm2_kernel<WM, WN, NWM, NWN, XTRA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    int nsm = 148, occ = 1;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, g_gemm2_kernel<WM, WN, NWM, NWN, XTRA>, G::NTH, G::SMEM);
    wave = nsm * (occ < 1 ? 1 : occ);
  }
  const int tm = (g.d + G::TM - 1) / G::TM, tn = (g.d + G::TN - 1) / G::TN;
  const long long items = (long long)tm * tn * nb;
  // small tiles (d < 64: two or three chunks per item) run one CTA per item: several CTAs are resident per SM anyway and the
  // per-chunk bookkeeping of the persistent walk costs more than it hides (d = 32, Nt = 1e5: 74 ms vs 84 ms persistent)
  static const int persist_all = [] { const char* e = getenv("QOC_GEMM_PERSIST"); return (e && e[0] == '1') ? 1 : 0; }();
  const long long grid = ((G::TM >= 64 || persist_all) && items > wave) ? wave : items;
  g_gemm2_kernel<WM, WN, NWM, NWN, XTRA><<<(int)grid, G::NTH, G::SMEM, st>>>(g, tm, tn, (int)items);
}
template <int WM, int WN, int NWM, int NWN>
static inline void g_gemm2_launch_t(const GGemm& g, int nb, cudaStream_t st) {
  if (g.nextra > 0) g_gemm2_launch_x<WM, WN, NWM, NWN, true>(g, nb, st);
  else g_gemm2_launch_x<WM, WN, NWM, NWN, false>(g, nb, st);
}
static inline int g_gemm_tile(int d) {
  static const int cand[4] = {64, 48, 40, 32};   // (80 x 80 exists for A/B runs, QOC_GEMM_TILE=80: 5.2 vs 4.1 ms on cavity-40)
  int best = 64, bestpad = 1 << 30;
  for (int i = 0; i < 4; i++) {
    const int t = cand[i], pad = (d + t - 1) / t * t;
    if (pad < bestpad) { bestpad = pad; best = t; }
  }
  const char* f = getenv("QOC_GEMM_TILE");
  if (f && atoi(f) > 0) best = atoi(f);
  return best;
}
static inline void g_gemm_launch(const GGemm& g, int nb, cudaStream_t st) {
  if (g.npairs == 0 && g.gj_nb == 0) {
    const int n2 = g.d * g.S / 2;
    int bx = (n2 + 255) / 256;
    if (bx > 8) bx = 8;
    g_lin_kernel<<<dim3(bx, nb), 256, 0, st>>>(g);
    return;
  }
  switch (g_gemm_tile(g.d)) {
    case 32: g_gemm2_launch_t<1, 2, 4, 2>(g, nb, st); break;
    case 40: g_gemm2_launch_t<1, 5, 5, 1>(g, nb, st); break;
    case 48: g_gemm2_launch_t<1, 6, 6, 1>(g, nb, st); break;
    case 80: g_gemm2_launch_t<1, 10, 10, 1>(g, nb, st); break;
    default: g_gemm2_launch_t<2, 2, 4, 4>(g, nb, st); break;
  }
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

__global__ void g_set2_kernel(double* p, double a, double b) { p[0] = a; p[1] = b; }

// |u[k][j]| <= bound[j] for every slice (the promise behind qoc_set_control_bounds): status 9 otherwise
__global__ void g_ubound_check_kernel(const double* u, int nc, long long n, const double* bound, int* status) {
  bool bad = false;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
    for (int j = 0; j < nc && j < 8; j++) bad |= !(fabs(u[k * nc + j]) <= bound[j]);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicExch(status, 9);
}

// ---- blocked Gauss-Jordan inverse with partial (row) pivoting -------------------------------------------------------------
// Replaces round 1's g_inverse_kernel (whole matrix eliminated in L2, d steps of d^2 work behind four barriers each: 23 % of
// the step at d = 64, 74 % at d = 256).  The matrix is processed in column panels K of nb <= 32 columns:
//   g_gj_panel_kernel  one CTA per slice: the d x nb panel in shared memory, nb pivot searches / eliminations on the panel
//                      only; leaves the transformation G in the K columns and applies the panel's row interchanges to the
//                      other columns.  O(d nb^2) work per panel, latency-bound, negligible.
//   g_gemm2_kernel     (gj_nb > 0) every other column block: (cur with its K rows zeroed) + G cur[K, :] -- a rank-nb DMMA
//                      update, out of place (cur -> oth); all d^3 flops of the inversion are here.
//   g_gj_perm_kernel   undoes the row interchanges as one column gather (A^-1 = (P A)^-1 P).
// With d <= 32 the single panel is the whole matrix and no update is needed.
__global__ void __launch_bounds__(256) g_gj_panel_kernel(int d, int S, int k0, int nb, double* M, long long stride, int* piv, int* status) {
  extern __shared__ __align__(16) unsigned char gsm[];
  const int NBP = nb <= 8 ? 8 : (nb <= 16 ? 16 : 32);   // lanes per panel row
  double2* pan = reinterpret_cast<double2*>(gsm);       // [d][NBP]
  double2* prow = pan + (size_t)d * NBP;                // old pivot row      [NBP]
  double2* krow = prow + 32;                            // old row k          [NBP]
  __shared__ double red_v[8];
  __shared__ int red_i[8];
  __shared__ int sh_piv[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* re = M + (long long)blockIdx.x * stride;
  double* im = re + (size_t)d * S;
  int* pv = piv + (size_t)blockIdx.x * d;
  const int rpw = 32 / NBP;                 // panel rows per warp pass
  const int c = lane % NBP, rsub = lane / NBP;
  // load the panel (columns k0 .. k0+nb-1 of every row)
  for (int i = warp * rpw + rsub; i < d; i += 8 * rpw)
    pan[(size_t)i * NBP + c] = (c < nb) ? make_double2(re[(size_t)i * S + k0 + c], im[(size_t)i * S + k0 + c]) : make_double2(0.0, 0.0);
  __syncthreads();
  bool ok = true;
  for (int kk = 0; kk < nb; kk++) {
    const int k = k0 + kk;
    // pivot search over the rows that have not pivoted yet (i >= k)
    double best = -1.0;
    int bi = k;
    if (k + warp * 32 < d) {   // (warp-uniform: a warp none of whose threads has a candidate row skips the reduction)
      for (int i = k + tid; i < d; i += 256) {
        const double2 v = pan[(size_t)i * NBP + kk];
        const double m = v.x * v.x + v.y * v.y;
        if (m > best) { best = m; bi = i; }
      }
      for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
    }
    if (lane == 0) { red_v[warp] = best; red_i[warp] = bi; }
    __syncthreads();
    double b = red_v[0];
    int p = red_i[0];
#pragma unroll
    for (int w = 1; w < 8; w++)
      if (red_v[w] > b || (red_v[w] == b && red_i[w] < p)) { b = red_v[w]; p = red_i[w]; }
    if (!(b > 0.0)) ok = false;
    // stage the two rows that trade places (old values), then update in place
    if (warp == 0 && lane < NBP) prow[lane] = pan[(size_t)p * NBP + lane];
    if (warp == 1 && lane < NBP) krow[lane] = pan[(size_t)k * NBP + lane];
    if (tid == 64) sh_piv[kk] = p;
    __syncthreads();
    const double2 pvv = prow[kk];
    const double den = 1.0 / (pvv.x * pvv.x + pvv.y * pvv.y);
    const double pir = pvv.x * den, pii = -pvv.y * den;   // 1 / pivot
    const double2 pr = prow[c];
    for (int i0 = warp * rpw; i0 < d; i0 += 8 * rpw) {   // warp-uniform trip count (__syncwarp inside)
      const int i = (i0 + rsub < d) ? i0 + rsub : d - 1;
      const bool live = (i0 + rsub < d) && c < nb;
      double2 x = pan[(size_t)i * NBP + c];
      if (i == p) x = krow[c];                 // the displaced row k lands in row p
      double2 val;
      if (i == k) {                            // pivot row: scaled
        if (c == kk) val = make_double2(pir, pii);
        else val = make_double2(pr.x * pir - pr.y * pii, pr.x * pii + pr.y * pir);
      } else {
        const double2 f = (i == p) ? krow[kk] : pan[(size_t)i * NBP + kk];   // multiplier: this row's entry in column kk
        const double gr = f.x * pir - f.y * pii, gi = f.x * pii + f.y * pir;
        if (c == kk) val = make_double2(-gr, -gi);
        else val = make_double2(x.x - (gr * pr.x - gi * pr.y), x.y - (gr * pr.y + gi * pr.x));
      }
      // every lane of the row group has read column kk of its row before anybody overwrites it
      __syncwarp();
      if (live) pan[(size_t)i * NBP + c] = val;
    }
    __syncthreads();
  }
  // panel back to the K columns; pivots of this panel to the list
  for (int i = warp * rpw + rsub; i < d; i += 8 * rpw)
    if (c < nb) { const double2 v = pan[(size_t)i * NBP + c]; re[(size_t)i * S + k0 + c] = v.x; im[(size_t)i * S + k0 + c] = v.y; }
  if (tid < nb) pv[k0 + tid] = sh_piv[tid];
  // the panel's row interchanges on every other column: columns are independent, each thread replays the nb swaps on its own
  for (int col = tid; col < d; col += 256) {
    if (col >= k0 && col < k0 + nb) continue;
    for (int kk = 0; kk < nb; kk++) {
      const int p = sh_piv[kk], k = k0 + kk;
      if (p != k) {
        const size_t a = (size_t)k * S + col, b2 = (size_t)p * S + col;
        const double tr = re[a], ti = im[a];
        re[a] = re[b2]; im[a] = im[b2];
        re[b2] = tr; im[b2] = ti;
      }
    }
  }
  if (tid == 0 && !ok) atomicExch(status, 8);
}

// dst[:, c] = src[:, idx[c]] with idx = the column interchanges k <-> piv[k], k = d-1 .. 0, composed
__global__ void __launch_bounds__(256) g_gj_perm_kernel(int d, int S, const double* src, double* dst, long long stride, const int* piv) {
  extern __shared__ __align__(16) unsigned char gsm[];
  int* idx = reinterpret_cast<int*>(gsm);
  const int tid = threadIdx.x;
  const int* pv = piv + (size_t)blockIdx.x * d;
  for (int i = tid; i < d; i += 256) idx[i] = i;
  __syncthreads();
  if (tid == 0)
    for (int k = d - 1; k >= 0; k--) {
      const int p = pv[k];
      if (p != k) { const int t = idx[k]; idx[k] = idx[p]; idx[p] = t; }
    }
  __syncthreads();
  const double* sre = src + (long long)blockIdx.x * stride;
  const double* sim = sre + (size_t)d * S;
  double* dre = dst + (long long)blockIdx.x * stride;
  double* dim_ = dre + (size_t)d * S;
  for (int e = tid; e < d * S; e += 256) {
    const int r = e / S, col = e - r * S;
    double vr = 0.0, vi = 0.0;
    if (col < d) { vr = sre[(size_t)r * S + idx[col]]; vi = sim[(size_t)r * S + idx[col]]; }
    dre[e] = vr; dim_[e] = vi;
  }
}

static inline int g_gj_panel_width(int d) { return d <= 384 ? 32 : 16; }
static inline size_t g_gj_panel_smem(int d) {
  const int nb = g_gj_panel_width(d) < d ? g_gj_panel_width(d) : d;
  const int NBP = nb <= 8 ? 8 : (nb <= 16 ? 16 : 32);
  return ((size_t)d * NBP + 64) * 16;
}

// N^-1 for `nsl` slices: cur (holds N, destroyed) and oth are two workspace slots; returns the buffer that holds the inverse.
static inline double* g_inverse_blocked(int d, int S, long long slot_d, int nsl, double* cur, double* oth, int* piv, int* status,
                                        cudaStream_t st, int* launches) {
  const int nbw = g_gj_panel_width(d);
  const size_t psm = g_gj_panel_smem(d);
  static size_t attr_of[64] = {0};   // per device
  int dev = 0;
  cudaGetDevice(&dev);
  size_t& attr_set = attr_of[dev & 63];
  if (psm > 40 * 1024 && psm > attr_set) {
    cudaFuncSetAttribute(g_gj_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);
    attr_set = psm;
  }
  for (int k0 = 0; k0 < d; k0 += nbw) {
    const int nb = (d - k0 < nbw) ? d - k0 : nbw;
    g_gj_panel_kernel<<<nsl, 256, psm, st>>>(d, S, k0, nb, cur, slot_d, piv, status);
    (*launches)++;
    if (d > nbw) {
      GGemm g;
      memset(&g, 0, sizeof g);
      g.d = d; g.S = S; g.nb = nsl; g.npairs = 1; g.nadd = 1; g.alpha = 1.0;
      g.A[0] = GOp{cur, slot_d, 0, 0}; g.B[0] = g.A[0]; g.D[0] = g.A[0]; g.beta[0] = 1.0;
      g.C = oth; g.cstride = slot_d;
      g.gj_k0 = k0; g.gj_nb = nb;
      g_gemm_launch(g, nsl, st);
      (*launches)++;
      double* t = cur; cur = oth; oth = t;
    }
  }
  g_gj_perm_kernel<<<nsl, 256, (size_t)d * 4, st>>>(d, S, cur, oth, slot_d, piv);
  (*launches)++;
  return oth;
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
  __shared__ double ov[16];
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
  if (tid < 16) ov[tid] = 0.0;
  __syncthreads();
  double J = 0.0;
  CostCoef cc;
  if (builtin)
    cost_overlaps_accumulate(g.T, d, m, [&](int r, int c) { return make_double2(cur[r * m + c], cur[dm + r * m + c]); }, ov, tid, 256, lane);
  if (pen && g.phase != 2) {
    double ps = Jpen;
    for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
    if (lane == 0) atomicAdd(&red[2], ps);
  }
  __syncthreads();
  if (builtin) { cost_from_overlaps(g.cost, g.n, m, ov, cc); J = cc.J; }
  if (pen && g.phase != 2) J += g.mu * red[2];
  if (tid == 0 && g.J && (builtin || (pen && g.phase != 2))) g.J[b] = J;
  if (g.phase == 1 || !g.want_grad) return;
  for (int e = tid; e < dm; e += 256) {
    const int c = e / d, r = e - c * d;
    double lr = 0.0, li = 0.0;
    if (g.lam_final) { lr = g.lam_final[(size_t)b * 2 * dm + 2 * e]; li = g.lam_final[(size_t)b * 2 * dm + 2 * e + 1]; }
    else if (builtin) {
      const double tr = g.T[2 * e], ti = g.T[2 * e + 1];
      double kr = cc.cr[0], ki = cc.ci[0];
      if (g.cost == QOC_COST_ZCAL_) {
#pragma unroll
        for (int q = 1; q < 4; q++) if (c == q) { kr = cc.cr[q]; ki = cc.ci[q]; }
      }
      lr = kr * tr - ki * ti; li = kr * ti + ki * tr;
    }
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
// the CTA pulls the planar slot of the NEXT step towards L2 while this step's mat-vec runs (the steps of a sweep are a latency
// chain: one L2 / HBM round trip per k-batch of operand fragments, nothing else in flight)
#ifndef GS_PF
#define GS_PF 1
#endif
__device__ __forceinline__ void gs_prefetch_slot(const double* slot, long long doubles, int tid, int nth) {
  if (!GS_PF) return;
  const char* q = reinterpret_cast<const char*>(slot);
  for (long long o = (long long)tid * 128; o < doubles * 8; o += (long long)nth * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + o));
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
  long long sl0;       // gs_contract_kernel: flat index of the first slice of this launch
  int L_chunked;       // ... and L_ holds the Jacobians of the slices [sl0, sl0 + gridDim.x) only (streamed Jacobians)
  // running state penalty L(x) = mu sum |x[pen rows, pen cols]|^2 (src/penalty_fcns.jl:1-11): the costate recurrence becomes
  // affine, lambda_k = U_k' lambda_{k+1} + dL_dx(x_k) (src/gradient_computations.jl:47-49, :55-57), so a segment is summarised
  // by (Q_seg, c_seg): lambda_{k0} = Q_seg' lambda_{k1} + c_seg, c_seg = the recurrence run from a zero costate.
  const unsigned char* pen_row;   // [d] 1 = penalised row (NULL: no penalty)
  unsigned col_mask;
  double mu;
  double* cs;          // [b*spp + s] d x m c128 affine terms
  double* Jpen;        // [batch] sum over the slices' states of |x[pen]|^2 (atomicAdd)
  int pen_prepass;     // gs_seg_kernel: forward states + c_seg + the penalty sum, nothing else
};

// xs[r][c] += 2 mu x[r][c] on the penalised entries (interleaved shared layouts, W = 2 m doubles per row)
__device__ __forceinline__ void gs_add_penalty(double* ys, const double* xs, const GS& g, int tid, int nth) {
  const int W = 2 * g.m;
  for (int e = tid; e < g.d * g.m; e += nth) {
    const int r = e / g.m, c = e - r * g.m;
    if (g.pen_row[r] && ((g.col_mask >> c) & 1u)) {
      ys[r * W + 2 * c] = fma(2.0 * g.mu, xs[r * W + 2 * c], ys[r * W + 2 * c]);
      ys[r * W + 2 * c + 1] = fma(2.0 * g.mu, xs[r * W + 2 * c + 1], ys[r * W + 2 * c + 1]);
    }
  }
}
// sum of |x[r][c]|^2 over the penalised entries, this thread's share
__device__ __forceinline__ double gs_penalty_sum(const double* xs, const GS& g, int tid, int nth) {
  const int W = 2 * g.m;
  double s = 0.0;
  for (int e = tid; e < g.d * g.m; e += nth) {
    const int r = e / g.m, c = e - r * g.m;
    if (g.pen_row[r] && ((g.col_mask >> c) & 1u)) s += xs[r * W + 2 * c] * xs[r * W + 2 * c] + xs[r * W + 2 * c + 1] * xs[r * W + 2 * c + 1];
  }
  return s;
}

// one CTA per pulse: boundary walk over the segment propagators
__global__ void __launch_bounds__(GS_NW * 32) gs_scan_kernel(GS g) {
  extern __shared__ __align__(16) unsigned char gsm[];
  const int d = g.d, m = g.m, W = 2 * m, dm = d * m;
  const int rows_pad = (d + 7) / 8 * 8;
  double* b0 = reinterpret_cast<double*>(gsm);
  double* b1 = b0 + rows_pad * W;
  __shared__ double ov[16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nth = blockDim.x;
  const int b = blockIdx.x;
  for (int e = tid; e < 2 * rows_pad * W; e += nth) b0[e] = 0.0;
  __syncthreads();
  double* cur = b0; double* nxt = b1;
  // mode 3 (penalty): cost + affine boundary walk from the stored x_N (the forward walk ran as mode 1, the c_seg in between)
  const bool pen = g.pen_row != nullptr;
  const bool do_fwd = g.mode != 2 && g.mode != 3, do_bwd = g.mode == 2 || g.mode == 4 || ((g.mode == 0 || g.mode == 3) && !g.skip_bwd);
  if (g.mode == 3 || (g.mode == 2 && pen)) {
    gs_from_global(cur, g.X + ((size_t)b * (g.nt + 1) + g.nt) * 2 * dm, d, m, tid, nth);
    __syncthreads();
  }
  if (do_fwd) {
    gs_from_global(cur, g.x_start_ext ? g.x_start_ext + (size_t)b * 2 * dm : g.x0, d, m, tid, nth);
    __syncthreads();
    for (int s = 0; s < g.spp; s++) {
      gs_to_global(g.xs_start + ((size_t)b * g.spp + s) * 2 * dm, cur, d, m, tid, nth);
      if (s + 1 < g.spp) gs_prefetch_slot(g.Q + ((size_t)b * g.spp + s + 1) * g.slot, g.slot, tid, nth);
      gs_mv_any<false>(g.Q + ((size_t)b * g.spp + s) * g.slot, d, g.S, m, cur, nxt, warp, lane);
      __syncthreads();
      double* t = cur; cur = nxt; nxt = t;
    }
    if (g.x_final) gs_to_global(g.x_final + (size_t)b * 2 * dm, cur, d, m, tid, nth);
    gs_to_global(g.X + ((size_t)b * (g.nt + 1) + g.nt) * 2 * dm, cur, d, m, tid, nth);
  }
  if (g.mode == 1) return;
  const bool builtin = (do_fwd || g.mode == 3) && g.cost != 2 && g.mode != 4;
  if (tid < 16) ov[tid] = 0.0;
  __syncthreads();
  CostCoef cc;
  cc.J = 0.0;
  if (builtin) {
    cost_overlaps_accumulate(g.T, d, m, [&](int r, int c) { return *reinterpret_cast<const double2*>(cur + r * W + 2 * c); },
                             ov, tid, nth, lane);
    __syncthreads();
    cost_from_overlaps(g.cost, g.n, m, ov, cc);
  }
  if (pen && g.mode == 3) {   // J = Jfinal(x_N) + sum_k L(x_k), k = 0 .. N (examples/ipopt_callbacks_exp.jl:18)
    double ps = gs_penalty_sum(cur, g, tid, nth);
    // (its own accumulator: ov[0 .. 2m) holds the per-column overlaps, m up to 8)
    __shared__ double pen_xN;
    if (tid == 0) pen_xN = 0.0;
    __syncthreads();
    for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
    if (lane == 0) atomicAdd(&pen_xN, ps);
    __syncthreads();
    if (tid == 0 && g.J) g.J[b] = cc.J + g.mu * (g.Jpen[b] + pen_xN);
  } else if (builtin && tid == 0 && g.J) g.J[b] = cc.J;
  if (!do_bwd) return;
  for (int c = 0; c < m; c++)
    for (int r = tid; r < d; r += nth) {
      double2 l = make_double2(0.0, 0.0);
      if (g.lam_final) l = reinterpret_cast<const double2*>(g.lam_final + (size_t)b * 2 * dm)[r + (size_t)d * c];
      else if (builtin) {
        const double2 t = reinterpret_cast<const double2*>(g.T)[r + (size_t)d * c];
        double kr = cc.cr[0], ki = cc.ci[0];
        if (g.cost == QOC_COST_ZCAL_) {
#pragma unroll
          for (int q = 1; q < 4; q++) if (c == q) { kr = cc.cr[q]; ki = cc.ci[q]; }
        }
        l = make_double2(kr * t.x - ki * t.y, kr * t.y + ki * t.x);
      }
      *reinterpret_cast<double2*>(nxt + r * W + 2 * c) = l;
    }
  __syncthreads();
  if (pen) { gs_add_penalty(nxt, cur, g, tid, nth); __syncthreads(); }   // lambda_N += dL_dx(x_N)  (:47-49)
  { double* t = cur; cur = nxt; nxt = t; }
  for (int s = g.spp - 1; s >= 0; s--) {
    gs_to_global(g.lam_end + ((size_t)b * g.spp + s) * 2 * dm, cur, d, m, tid, nth);
    if (s > 0) gs_prefetch_slot(g.Q + ((size_t)b * g.spp + s - 1) * g.slot, g.slot, tid, nth);
    gs_mv_any<true>(g.Q + ((size_t)b * g.spp + s) * g.slot, d, g.S, m, cur, nxt, warp, lane);
    __syncthreads();
    if (pen) {   // + c_seg
      const double2* cg = reinterpret_cast<const double2*>(g.cs + ((size_t)b * g.spp + s) * 2 * dm);
      for (int c = 0; c < m; c++)
        for (int r = tid; r < d; r += nth) { const double2 v = cg[r + (size_t)d * c]; nxt[r * W + 2 * c] += v.x; nxt[r * W + 2 * c + 1] += v.y; }
      __syncthreads();
    }
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
  double* b2 = b1 + rows_pad * W;   // penalty only: x_k next to the costate
  __shared__ double psum;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nth = blockDim.x;
  const bool pen = g.pen_row != nullptr;
  for (int e = tid; e < (pen ? 3 : 2) * rows_pad * W; e += nth) b0[e] = 0.0;
  __syncthreads();
  const bool do_fwd = (g.mode != 2 && g.mode != 3) || g.pen_prepass;
  const bool do_bwd = !g.pen_prepass && (g.mode == 2 || g.mode == 3 || g.mode == 4 || (g.mode == 0 && !g.skip_bwd));
  for (int seg = blockIdx.x; seg < nseg_total; seg += gridDim.x) {
    const int b = seg / g.spp, si = seg - b * g.spp;
    const int k0 = si * g.L, k1 = (k0 + g.L < g.nt) ? k0 + g.L : g.nt;
    double* cur = b0; double* nxt = b1;
    if (do_fwd) {
      double ps = 0.0;
      gs_from_global(cur, g.xs_start + (size_t)seg * 2 * dm, d, m, tid, nth);
      __syncthreads();
      for (int k = k0; k < k1; k++) {
        gs_to_global(g.X + ((size_t)b * (g.nt + 1) + k) * 2 * dm, cur, d, m, tid, nth);
        if (g.pen_prepass) ps += gs_penalty_sum(cur, g, tid, nth);
        if (k + 1 < k1) {
          if (k + 1 < k1) gs_prefetch_slot(g.U + ((size_t)b * g.nt + k + 1) * g.slot, g.slot, tid, nth);
          gs_mv_any<false>(g.U + ((size_t)b * g.nt + k) * g.slot, d, g.S, m, cur, nxt, warp, lane);
          __syncthreads();
          double* t = cur; cur = nxt; nxt = t;
        }
      }
      __syncthreads();
      if (g.pen_prepass) {
        // sum_k L(x_k)/mu of this segment, and c_seg: the affine costate recurrence from a zero costate
        if (tid == 0) psum = 0.0;
        __syncthreads();
        for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
        if (lane == 0) atomicAdd(&psum, ps);
        __syncthreads();
        if (tid == 0) atomicAdd(g.Jpen + b, psum);
        for (int e = tid; e < rows_pad * W; e += nth) cur[e] = 0.0;
        __syncthreads();
        for (int k = k1 - 1; k >= k0; k--) {
          gs_from_global(b2, g.X + ((size_t)b * (g.nt + 1) + k) * 2 * dm, d, m, tid, nth);
          if (k > k0) gs_prefetch_slot(g.U + ((size_t)b * g.nt + k - 1) * g.slot, g.slot, tid, nth);
        gs_mv_any<true>(g.U + ((size_t)b * g.nt + k) * g.slot, d, g.S, m, cur, nxt, warp, lane);
          __syncthreads();
          gs_add_penalty(nxt, b2, g, tid, nth);
          __syncthreads();
          double* t = cur; cur = nxt; nxt = t;
        }
        gs_to_global(g.cs + (size_t)seg * 2 * dm, cur, d, m, tid, nth);
        __syncthreads();
      }
    }
    if (do_bwd) {
      gs_from_global(cur, g.lam_end + (size_t)seg * 2 * dm, d, m, tid, nth);
      __syncthreads();
      for (int k = k1 - 1; k >= k0; k--) {
        gs_to_global(g.LAM + ((size_t)b * (g.nt + 1) + k + 1) * 2 * dm, cur, d, m, tid, nth);
        if (pen) gs_from_global(b2, g.X + ((size_t)b * (g.nt + 1) + k) * 2 * dm, d, m, tid, nth);
        if (k > k0) gs_prefetch_slot(g.U + ((size_t)b * g.nt + k - 1) * g.slot, g.slot, tid, nth);
        gs_mv_any<true>(g.U + ((size_t)b * g.nt + k) * g.slot, d, g.S, m, cur, nxt, warp, lane);
        __syncthreads();
        if (pen) { gs_add_penalty(nxt, b2, g, tid, nth); __syncthreads(); }   // + dL_dx(x_k)  (:55-57)
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
  const size_t sl = (size_t)g.sl0 + blockIdx.x;  // flat slice index b*nt + k
  const int j = blockIdx.y;
  const int b = (int)(sl / g.nt), k = (int)(sl - (size_t)b * g.nt);
  for (int e = tid; e < (S + d) * W; e += 256) xk[e] = 0.0;
  __syncthreads();
  gs_from_global(xk, g.X + ((size_t)b * (g.nt + 1) + k) * 2 * dm, d, m, tid, 256);
  gs_from_global(lk, g.LAM + ((size_t)b * (g.nt + 1) + k + 1) * 2 * dm, d, m, tid, 256);
  __syncthreads();
  const double* Lre = g.L_ + ((g.L_chunked ? (size_t)blockIdx.x : sl) * g.nc + j) * g.slot;
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

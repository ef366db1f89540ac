// qoc_sweep.cuh -- second-generation sweep kernels (no running penalty): K2G, the boundary scan over the segment
// propagators as a TWO-LEVEL scan in one launch, and K3N, the per-segment forward / costate sweeps with the adjoint
// gradient contraction.  They replace the reference's three serial loops
//   x[k+1] = U_k x[k]                              src/gradient_computations.jl:27-29
//   lambda[k] = U_k' lambda[k+1]                   src/gradient_computations.jl:52-58
//   dJdu[j,k] = sum_l Re(dot(lambda[k+1][:,l], dU_k/du_j, x[k][:,l]))   :65-74, :217-223
// and the terminal cost closures of src/penalty_fcns.jl:15-24 and test/test_gradient_computation.jl:24.
//
// What is different from qoc_k23.cuh (which stays for the affine, running-penalty recurrence):
//  * mat-vec step: the d x m complex state is kept INTERLEAVED in shared memory (row k: re0 im0 re1 im1 ...), so one
//    8-wide DMMA column tile carries [x_re, x_im] of up to 4 state columns and a complex mat-vec costs TWO real DMMAs per
//    k-step (U_re * [x_re x_im], U_im * [x_re x_im]) instead of four; the cross terms are recombined in registers.
//    The serial recurrences are bound by the FP64 pipe occupancy of these DMMAs (16.5 cycles each, measured), so this
//    halves the step.
//  * operands arrive through a TMA ring (cp.async.bulk + mbarrier): one thread issues one bulk copy per slot.
//  * K2G: G CTAs per pulse.  Each multiplies its group's segment propagators (P_g, d^3 tile products), a per-pulse
//    barrier publishes them, every CTA then walks the G group propagators (redundantly, 2 G short steps) to get the
//    state entering / the costate leaving its group, and finally walks its own segments.  Serial depth
//    ~ spp/G products + 2 G + 2 spp/G mat-vecs instead of 2 spp mat-vecs.
//  * K3N: the gradient contraction Re tr(lambda' dU x) = <dU, lambda x'>_F is evaluated ELEMENTWISE by separate warps
//    that read dU_k/du_j straight from global memory (coalesced 16-byte loads, register prefetch one step ahead, L2
//    bulk prefetch further ahead): no DMMA-pipe time and no shared-memory staging for the Jacobians.
#pragma once
#include "qoc_k23.cuh"

namespace qoc {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = smem_u32(bar);
  unsigned ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ unsigned mbar_try(unsigned long long* bar, unsigned parity) {   // non-blocking probe
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
// global -> shared bulk copy (TMA, SASS UBLKCP) completing on an mbarrier; bytes % 16 == 0, 16-byte aligned
__device__ __forceinline__ void tma_load(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void* gsrc, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// rows of one interleaved state buffer: covers the k-steps of the tile loop (4 KS) and the row stride of a slot (S)
template <class C>
__host__ __device__ constexpr int state_rows() { return (4 * C::KS > C::S) ? 4 * C::KS : C::S; }

// ---- mat-vec step ----------------------------------------------------------------------------------------------------
// y = op(A) x for the 8-row output tile mi.  xs / ys: interleaved state buffers, row stride W = 2 m doubles, rows >= d of
// xs are zero.  The operand fragment (thread (g, q) holds A[8 mi + g][4 ks + q], or the transposed element for A^dagger)
// does not depend on x: it is loaded BEFORE the barrier that publishes x, off the critical path of the recurrence.
template <class C>
struct AFrag {
  double ar[C::KS], ai[C::KS];
  __device__ __forceinline__ void load(Mat A, int mi, int lane, bool adj) {
    constexpr int S = C::S;
    const int g = lane >> 2, q = lane & 3;
    const int off = adj ? q * S + mi * 8 + g : (mi * 8 + g) * S + q;
    const int astep = adj ? 4 * S : 4;
    const double* are = A.re + off;
    const double* aim = A.im + off;
#pragma unroll
    for (int ks = 0; ks < C::KS; ks++) { ar[ks] = are[ks * astep]; ai[ks] = aim[ks * astep]; }
  }
};

// DMMA form: the 8-wide column tile carries [re, im] of 4 state columns; NCT = ceil(2 m / 8) column tiles.
template <class C, bool ADJ, int NCT>
__device__ __forceinline__ void mv2(const AFrag<C>& a, const double* xs, double* ys, int d, int m, int mi, int lane) {
  const int g = lane >> 2, q = lane & 3, W = 2 * m;
  double p1[2][NCT][2], p2[2][NCT][2];
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int ct = 0; ct < NCT; ct++) { p1[h][ct][0] = p1[h][ct][1] = 0.0; p2[h][ct][0] = p2[h][ct][1] = 0.0; }
  const double* xp = xs + q * W + g;
#pragma unroll
  for (int ks = 0; ks < C::KS; ks++) {
    const int h = ks & 1;
#pragma unroll
    for (int ct = 0; ct < NCT; ct++) {
      const double bv = (ct * 8 + g < W) ? xp[ks * 4 * W + ct * 8] : 0.0;
      dmma(p1[h][ct][0], p1[h][ct][1], a.ar[ks], bv);
      dmma(p2[h][ct][0], p2[h][ct][1], a.ai[ks], bv);
    }
  }
  const int row = mi * 8 + g;
#pragma unroll
  for (int ct = 0; ct < NCT; ct++) {
    const int c = ct * 4 + q;  // state column held by this thread: (re, im) pair
    const double u = p1[0][ct][0] + p1[1][ct][0], v = p1[0][ct][1] + p1[1][ct][1];   // A_re x_re, A_re x_im
    const double e = p2[0][ct][0] + p2[1][ct][0], f = p2[0][ct][1] + p2[1][ct][1];   // A_im x_re, A_im x_im
    const double yr = ADJ ? u + f : u - f;
    const double yi = ADJ ? v - e : v + e;
    if (row < d && c < m) *reinterpret_cast<double2*>(ys + row * W + 2 * c) = make_double2(yr, yi);
  }
}

// DFMA form for m <= 2 (a DMMA tile would be 3/4 or 1/2 empty and the recurrence is bound by the FP64 pipe time of the
// step): the same fragment ownership, the four q-lanes of a row each sum their quarter of the k range, two shuffle
// stages finish the row.
template <class C, bool ADJ, int M>
__device__ __forceinline__ void mvf(const AFrag<C>& a, const double* xs, double* ys, int d, int mi, int lane) {
  constexpr int W = 2 * M;
  const int g = lane >> 2, q = lane & 3;
  double yr[2][M], yi[2][M];
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int c = 0; c < M; c++) { yr[h][c] = 0.0; yi[h][c] = 0.0; }
  const double* xp = xs + q * W;
#pragma unroll
  for (int ks = 0; ks < C::KS; ks++) {
    const int h = ks & 1;
    const double ar = a.ar[ks], ai = ADJ ? -a.ai[ks] : a.ai[ks];
#pragma unroll
    for (int c = 0; c < M; c++) {
      const double2 x = *reinterpret_cast<const double2*>(xp + ks * 4 * W + 2 * c);
      yr[h][c] = fma(ar, x.x, fma(-ai, x.y, yr[h][c]));
      yi[h][c] = fma(ar, x.y, fma(ai, x.x, yi[h][c]));
    }
  }
  const int row = mi * 8 + g;
#pragma unroll
  for (int c = 0; c < M; c++) {
    double r = yr[0][c] + yr[1][c], i = yi[0][c] + yi[1][c];
    r += __shfl_xor_sync(0xffffffffu, r, 1); i += __shfl_xor_sync(0xffffffffu, i, 1);
    r += __shfl_xor_sync(0xffffffffu, r, 2); i += __shfl_xor_sync(0xffffffffu, i, 2);
    if (q == c % 4 && row < d) *reinterpret_cast<double2*>(ys + row * W + 2 * c) = make_double2(r, i);
  }
}

template <class C, bool ADJ>
__device__ __forceinline__ void mv_any(const AFrag<C>& a, const double* xs, double* ys, int d, int m, int mi, int lane) {
  if (m == 1) mvf<C, ADJ, 1>(a, xs, ys, d, mi, lane);
  else if (m == 2) mvf<C, ADJ, 2>(a, xs, ys, d, mi, lane);
  else if (m <= 4) mv2<C, ADJ, 1>(a, xs, ys, d, m, mi, lane);
  else mv2<C, ADJ, 2>(a, xs, ys, d, m, mi, lane);
}

// interleaved (shared) <-> c128 column-major (global)
__device__ __forceinline__ void ist_to_global(double* g, const double* xs, int d, int m, int tid, int nthreads) {
  for (int c = 0; c < m; c++)
    for (int r = tid; r < d; r += nthreads)
      reinterpret_cast<double2*>(g)[r + d * c] = *reinterpret_cast<const double2*>(xs + r * 2 * m + 2 * c);
}
__device__ __forceinline__ void ist_from_global(double* xs, const double* g, int d, int m, int tid, int nthreads) {
  for (int c = 0; c < m; c++)
    for (int r = tid; r < d; r += nthreads)
      *reinterpret_cast<double2*>(xs + r * 2 * m + 2 * c) = reinterpret_cast<const double2*>(g)[r + d * c];
}

#ifndef QOC_SW_NST
#define QOC_SW_NST 8
#endif
constexpr int SW_NST = QOC_SW_NST;  // TMA ring depth of both kernels

// Ring of SW_NST operand slots filled by TMA.  Stream positions are numbered from 0 over the whole kernel; position pos
// uses stage pos % NST for the (pos / NST)-th time.  full[s]: completes when the bulk copy has landed (expect_tx);
// empty[s]: completes when the NT consumer warps have pulled their fragments of the stage into registers.  The producer
// (one lane) runs ahead on its own: it never takes part in the consumers' barriers.
struct TmaRing {
  double* slots;
  unsigned long long* full;
  unsigned long long* empty;
  int slot_d;
  unsigned bytes;
  __device__ __forceinline__ void init(int nconsumers) const {   // one thread, before a __syncthreads
    for (int i = 0; i < SW_NST; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], nconsumers); }
    fence_mbar_init();
  }
  __device__ __forceinline__ double* stage(int pos) const { return slots + (size_t)(pos % SW_NST) * slot_d; }
  __device__ __forceinline__ unsigned try_full(int pos) const { return mbar_try(&full[pos % SW_NST], (unsigned)((pos / SW_NST) & 1)); }
  __device__ __forceinline__ void wait_full(int pos) const { mbar_wait(&full[pos % SW_NST], (unsigned)((pos / SW_NST) & 1)); }
  __device__ __forceinline__ void release(int pos) const { mbar_arrive(&empty[pos % SW_NST]); }   // one lane per consumer warp
  __device__ __forceinline__ void produce(int pos, const double* src) const {   // one thread
    const int n = pos / SW_NST;
    if (n > 0) mbar_wait(&empty[pos % SW_NST], (unsigned)((n - 1) & 1));
    mbar_expect_tx(&full[pos % SW_NST], bytes);
    tma_load(stage(pos), src, bytes, &full[pos % SW_NST]);
  }
};

// named barrier 1: the NT recurrence warps
__device__ __forceinline__ void bar_rec(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// One step of a serial recurrence, executed by the NT recurrence warps: out = op(M) in, M = the operand of stream position
// pos.  The operand fragments go to registers first (they do not depend on `in`), the stage is handed back to the producer,
// and only then the barrier that orders the previous step's writes of `in` before this step's reads is taken.
// The caller issues one more bar_rec() before anybody else reads the last `out`.
// (A variant that probes the next stage and loads its fragments behind the mat-vec of the current step measured slower.)
template <class C, bool ADJ>
__device__ __forceinline__ void rec_step(const TmaRing& ring, int pos, const double* in, double* out, int d, int m, int warp,
                                         int lane) {
  constexpr int S = C::S;
  ring.wait_full(pos);
  Mat Mm; Mm.re = ring.stage(pos); Mm.im = Mm.re + d * S;
  AFrag<C> a;
  a.load(Mm, warp, lane, ADJ);
  __syncwarp();
  if (lane == 0) ring.release(pos);
  bar_rec(C::NT * 32);
  mv_any<C, ADJ>(a, in, out, d, m, warp, lane);
}

struct K2GParams {
  K23Params q;
  int G;                 // groups per pulse
  double* Pg;            // [batch * G] planar slots: group propagators
  unsigned* sync;        // [batch] arrival counters, never reset
  unsigned sync_target;  // G * (launch index): value the counter reaches when every CTA of the pulse has published
  double* S_out;         // k2_phase 5: c128 column-major d x d product of ALL segment propagators of pulse 0 (time sharding)
};

// K2G: grid = batch * G CTAs (all co-resident: the host guarantees batch * G <= resident capacity);
// C::NTHREADS product threads + one producer warp.  Warps [0, NT) run the walks, warp NT is the TMA producer.
template <class C>
__global__ void __launch_bounds__(C::NTHREADS + 64, 1) k2g_kernel(K2GParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int S = C::S;
  constexpr int NT = C::NT;
  constexpr int RP = state_rows<C>();
  const int NTH = blockDim.x;
  const K23Params& p = P.q;
  const int d = p.d, m = p.m, spp = p.seg_per_pulse, G = P.G, W = 2 * m;
  const int slot_d = 2 * d * S, n2 = slot_d / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / G, g = blockIdx.x - b * G;
  const int s0 = (int)(((long long)g * spp) / G), s1 = (int)(((long long)(g + 1) * spp) / G);
  const int ngs = s1 - s0;
  const size_t seg0 = (size_t)b * spp;
  const int dm = d * m;

  double* base = reinterpret_cast<double*>(smem_raw);
  double* ringp = base;                                   // SW_NST slots
  double* prod = ringp + (size_t)SW_NST * slot_d;         // 2 slots (running product ping-pong)
  double* pad = prod + (size_t)2 * slot_d;                // zero rows behind the last slot
  const int pad_rows = k1_pad_rows<C>(d);
  double* sbuf = pad + pad_rows * S;                      // 4 state buffers: xa, xb, save_x, save_l
  const int sb = RP * W;
  double* red = sbuf + 4 * sb;                            // 4 doubles
  __shared__ double ov[16];                               // per-column overlaps diag(T' x_N)
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(red + 4);
  TmaRing ring;
  ring.slots = ringp; ring.full = bars; ring.empty = bars + SW_NST; ring.slot_d = slot_d; ring.bytes = (unsigned)slot_d * 8u;
  {
    double2* z = reinterpret_cast<double2*>(base);
    const int total2 = (int)((red + 4 - base) / 2);
    for (int e = tid; e < total2; e += NTH) z[e] = make_double2(0.0, 0.0);
    if (tid == 0) ring.init(NT);
  }
  __syncthreads();
  int gi = 0;  // global stream position (the mbarrier phases run on across the phases of the kernel)

  // ---------------- phase A: P_g = Q_{s1-1} ... Q_{s0} ----------------
  {
    auto srcA = [&](int i) { return p.Q + (seg0 + s0 + i) * slot_d; };
    if (tid == 0) for (int i = 0; i < SW_NST - 1 && i < ngs; i++) ring.produce(gi + i, srcA(i));
    int cur = 0;
    const int mi = warp / (C::NT / C::BN), nj0 = (warp % (C::NT / C::BN)) * C::BN;
    for (int i = 0; i < ngs; i++) {
      ring.wait_full(gi + i);
      if (tid == 0 && i + SW_NST - 1 < ngs) ring.produce(gi + i + SW_NST - 1, srcA(i + SW_NST - 1));   // stage of step i-1
      Mat Qm; Qm.re = ring.stage(gi + i); Qm.im = Qm.re + d * S;
      Mat Pc; Pc.re = prod + (size_t)cur * slot_d; Pc.im = Pc.re + d * S;
      Mat Pn; Pn.re = prod + (size_t)(cur ^ 1) * slot_d; Pn.im = Pn.re + d * S;
      if (i == 0) {
        slot_copy(Pn.re, Qm.re, n2, tid, NTH);
      } else if (warp < C::NBLK) {
        Acc<C::BN> acc; acc.zero();
        mm_acc<C, false>(acc, Qm, Pc, mi, nj0, lane);
        mm_store<C>(Pn, acc, d, mi, nj0, lane, NoEpi());
      }
      cur ^= 1;
      __syncthreads();
      if (warp < NT && lane == 0) ring.release(gi + i);
    }
    gi += ngs;
    slot_copy(P.Pg + ((size_t)b * G + g) * slot_d, prod + (size_t)cur * slot_d, n2, tid, NTH);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      atomicAdd(&P.sync[b], 1u);
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(P.sync + b) : "memory");
      } while ((int)(v - P.sync_target) < 0);
      __threadfence();
    }
    __syncthreads();
  }

  const int mode = p.k2_phase;   // 0: forward + cost + backward, 1: forward only, 2: backward only (lam_final),
                                 // 4: forward from x_start_ext and backward from lam_final, no cost (time sharding),
                                 // 5: only the product of all segment propagators -> S_out (time sharding, phase 1)
  if (mode == 5) {
    // S = P_{G-1} ... P_0 by CTA 0 of the pulse: G - 1 sequential tile products over the freshly published group products
    if (g != 0) return;
    auto srcP = [&](int i) { return P.Pg + ((size_t)b * G + i) * slot_d; };
    if (tid == 0) for (int i = 0; i < SW_NST - 1 && i < G; i++) ring.produce(gi + i, srcP(i));
    int cur = 0;
    const int mi = warp / (C::NT / C::BN), nj0 = (warp % (C::NT / C::BN)) * C::BN;
    for (int i = 0; i < G; i++) {
      ring.wait_full(gi + i);
      if (tid == 0 && i + SW_NST - 1 < G) ring.produce(gi + i + SW_NST - 1, srcP(i + SW_NST - 1));
      Mat Qm; Qm.re = ring.stage(gi + i); Qm.im = Qm.re + d * S;
      Mat Pc; Pc.re = prod + (size_t)cur * slot_d; Pc.im = Pc.re + d * S;
      Mat Pn; Pn.re = prod + (size_t)(cur ^ 1) * slot_d; Pn.im = Pn.re + d * S;
      if (i == 0) {
        slot_copy(Pn.re, Qm.re, n2, tid, NTH);
      } else if (warp < C::NBLK) {
        Acc<C::BN> acc; acc.zero();
        mm_acc<C, false>(acc, Qm, Pc, mi, nj0, lane);
        mm_store<C>(Pn, acc, d, mi, nj0, lane, NoEpi());
      }
      cur ^= 1;
      __syncthreads();
      if (warp < NT && lane == 0) ring.release(gi + i);
    }
    const double* Pre = prod + (size_t)cur * slot_d;
    for (int e = tid; e < d * d; e += NTH) {
      const int c = e / d, r = e - c * d;
      reinterpret_cast<double2*>(P.S_out)[e] = make_double2(Pre[r * S + c], Pre[d * S + r * S + c]);
    }
    return;
  }

  // ---------------- phases B (walk the groups) and C (walk this group's segments) ----------------
  const bool do_fwd = (mode != 2), do_bwd = (mode == 2) || (mode == 4) || (mode == 0 && !p.skip_cost);
  const int nB1 = do_fwd ? G : 0, nB2 = do_bwd ? (G - 1 - g) : 0, nC1 = do_fwd ? ngs : 0, nC2 = do_bwd ? ngs : 0;
  const int total = nB1 + nB2 + nC1 + nC2;
  const int gi0 = gi;
  auto src = [&](int pos) -> const double* {
    int i = pos - gi0;
    if (i < nB1) return P.Pg + ((size_t)b * G + i) * slot_d;
    i -= nB1;
    if (i < nB2) return P.Pg + ((size_t)b * G + (G - 1 - i)) * slot_d;
    i -= nB2;
    if (i < nC1) return p.Q + (seg0 + s0 + i) * slot_d;
    i -= nC1;
    return p.Q + (seg0 + s1 - 1 - i) * slot_d;
  };
  if (warp == NT) {        // producer: runs the whole stream ahead of the walks
    if (lane == 0) for (int pos = gi0; pos < gi0 + total; pos++) ring.produce(pos, src(pos));
    return;
  }
  if (warp > NT) return;   // the walks are run by the NT recurrence warps
  const int RT = NT * 32;  // recurrence threads
  const bool rec = true;
  int li = 0;  // local stream index
  double* xa = sbuf; double* xb = sbuf + sb; double* save_x = sbuf + 2 * sb; double* save_l = sbuf + 3 * sb;
  auto step = [&](bool adj, const double* in, double* out) {
    if (adj) rec_step<C, true>(ring, gi0 + li, in, out, d, m, warp, lane);
    else rec_step<C, false>(ring, gi0 + li, in, out, d, m, warp, lane);
    li++;
  };

  double* cur = xa; double* nxt = xb;
  if (do_fwd) {
    const double* xin = p.x_start_ext ? p.x_start_ext + (size_t)b * 2 * dm : p.x0;
    if (rec) ist_from_global(cur, xin, d, m, tid, RT);
    for (int i = 0; i < G; i++) {
      step(false, cur, nxt);
      // after the barrier inside step(): `cur` is complete and stable -> the group's entering state can be saved
      if (i == g && rec) for (int e = tid; e < d * W; e += RT) save_x[e] = cur[e];
      double* t = cur; cur = nxt; nxt = t;
    }
    bar_rec(NT * 32);
    if (g == 0 && rec) {
      if (p.x_final) ist_to_global(p.x_final + (size_t)b * 2 * dm, cur, d, m, tid, RT);
      if (p.X && p.store_states) ist_to_global(p.X + ((size_t)b * (p.nt + 1) + p.nt) * 2 * dm, cur, d, m, tid, RT);
    }
  }
  if (mode == 0 || mode == 2 || mode == 4) {
    // terminal cost and costate:  lambda_N = dJfinal_dx(x_N)     src/gradient_computations.jl:46, penalty_fcns.jl:15-24
    const bool builtin = do_fwd && p.cost != 2 && mode != 4;
    if (tid < 16) ov[tid] = 0.0;
    bar_rec(NT * 32);
    CostCoef cc;
    if (builtin) {
      // per-column overlaps m_c = sum_r conj(T[r][c]) x_N[r][c] (Omega = tr(T'x) is their sum)
      cost_overlaps_accumulate(p.T, d, m, [&](int r, int c) { return *reinterpret_cast<const double2*>(cur + r * W + 2 * c); },
                               ov, tid, RT, lane);
      bar_rec(NT * 32);
      cost_from_overlaps(p.cost, p.n, m, ov, cc);
      if (tid == 0 && g == 0 && p.J) p.J[b] = cc.J;
    }
    if (do_bwd) {
      if (rec)
        for (int c = 0; c < m; c++)
          for (int r = tid; r < d; r += RT) {
            double2 l = make_double2(0.0, 0.0);
            if (p.lam_final) l = reinterpret_cast<const double2*>(p.lam_final + (size_t)b * 2 * dm)[r + d * c];
            else if (builtin) {
              const double2 t = reinterpret_cast<const double2*>(p.T)[r + d * c];
              double kr = cc.cr[0], ki = cc.ci[0];   // per-column coefficients only for the z-calibrated cost
              if (p.cost == QOC_COST_ZCAL_) {
#pragma unroll
                for (int q = 1; q < 4; q++) if (c == q) { kr = cc.cr[q]; ki = cc.ci[q]; }
              }
              l = make_double2(kr * t.x - ki * t.y, kr * t.y + ki * t.x);
            }
            *reinterpret_cast<double2*>(nxt + r * W + 2 * c) = l;
          }
      { double* t = cur; cur = nxt; nxt = t; }
      for (int i = 0; i < nB2; i++) {
        step(true, cur, nxt);
        double* t = cur; cur = nxt; nxt = t;
      }
      bar_rec(NT * 32);
      if (rec) for (int e = tid; e < d * W; e += RT) save_l[e] = cur[e];
    }
  }
  if (do_fwd) {
    const double* in = save_x;
    for (int i = 0; i < ngs; i++) {
      double* out = (in == xa) ? xb : xa;
      step(false, in, out);
      if (rec) ist_to_global(p.xs_start + (seg0 + s0 + i) * 2 * dm, in, d, m, tid, RT);
      in = out;
    }
  }
  if (do_bwd) {
    const double* in = save_l;
    for (int i = 0; i < ngs; i++) {
      double* out = (in == xa) ? xb : xa;
      step(true, in, out);
      if (rec) ist_to_global(p.lam_end + (seg0 + s1 - 1 - i) * 2 * dm, in, d, m, tid, RT);
      in = out;
    }
    bar_rec(NT * 32);
    if (g == 0 && p.lam_start && rec) ist_to_global(p.lam_start + (size_t)b * 2 * dm, in, d, m, tid, RT);
  }
}

// Boundary algebra of the time-sharded evaluation, redundantly on every rank, one CTA:
//   x_start(rank) = S_{rank-1} ... S_0 x0,  x_N,  J,  lambda_N = dJfinal_dx(x_N),  lambda_end(rank) = S_{rank+1}' ... S_{P-1}' lambda_N
// S_all: P consecutive c128 column-major d x d rank propagators (the all-gathered output of phase 1).
struct ShardBoundary {
  int d, m, nranks, rank, cost, n;
  const double* S_all;
  const double* x0;
  const double* T;
  double* x_start;   // c128 d x m
  double* lam_end;   // c128 d x m
  double* J;
  // running state penalty (src/penalty_fcns.jl:1-11): the rank recurrence is affine, lambda_start(p) = S_p' lambda_end(p) + c_p.
  // C_all: P records of 2 d m + 2 doubles (c_p, then the rank's sum_k L(x_k) over its nt_p + 1 states); xs: P x (d x m c128)
  // scratch for the boundary states.  A boundary state is the last state of rank p - 1 and the first of rank p: its L and
  // dL_dx are taken out once.  C_all == NULL (first call, before the forward sweeps): only x_start is meaningful.
  int pen;
  const double* C_all;
  double* xs;
  const unsigned char* pen_row;   // [d] byte mask (general path) or NULL -> row_mask
  unsigned long long row_mask;
  unsigned col_mask;
  double mu;
};

__global__ void __launch_bounds__(256) shard_boundary_kernel(ShardBoundary q) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = q.d, m = q.m, dm = d * m, tid = threadIdx.x;
  double2* xa = reinterpret_cast<double2*>(smem_raw);   // column-major d x m, like the ABI
  double2* xb = xa + dm;
  __shared__ double ov[16];
  for (int e = tid; e < dm; e += 256) xa[e] = reinterpret_cast<const double2*>(q.x0)[e];
  __syncthreads();
  auto matvec = [&](const double* Sp, bool adj, const double2* in, double2* out) {
    for (int e = tid; e < dm; e += 256) {
      const int c = e / d, r = e - c * d;
      double yr = 0.0, yi = 0.0;
      for (int k = 0; k < d; k++) {
        const double2 s = reinterpret_cast<const double2*>(Sp)[adj ? (k + (size_t)d * r) : (r + (size_t)d * k)];
        const double sr = s.x, si = adj ? -s.y : s.y;
        const double2 x = in[k + d * c];
        yr = fma(sr, x.x, fma(-si, x.y, yr));
        yi = fma(sr, x.y, fma(si, x.x, yi));
      }
      out[e] = make_double2(yr, yi);
    }
    __syncthreads();
  };
  double2* cur = xa; double2* nxt = xb;
  const bool pen = q.pen && q.C_all != nullptr;
  auto penalised = [&](int r, int c) {
    return (q.pen_row ? q.pen_row[r] != 0 : (r < 64 && ((q.row_mask >> r) & 1ull))) && ((q.col_mask >> c) & 1u);
  };
  double lsub = 0.0;   // this thread's share of sum_{p >= 1} |x_start(p)[pen]|^2
  for (int p = 0; p < q.nranks; p++) {
    if (p == q.rank) for (int e = tid; e < dm; e += 256) reinterpret_cast<double2*>(q.x_start)[e] = cur[e];
    if (pen) for (int e = tid; e < dm; e += 256) {
      reinterpret_cast<double2*>(q.xs)[(size_t)p * dm + e] = cur[e];   // re-read below by the same thread
      if (p >= 1 && penalised(e % d, e / d)) lsub += cur[e].x * cur[e].x + cur[e].y * cur[e].y;
    }
    matvec(q.S_all + (size_t)p * 2 * d * d, false, cur, nxt);
    double2* t = cur; cur = nxt; nxt = t;
  }
  // cost and terminal costate (src/penalty_fcns.jl:15-24, test/test_gradient_computation.jl:24-25)
  if (tid < 16) ov[tid] = 0.0;
  __syncthreads();
  cost_overlaps_accumulate(q.T, d, m, [&](int r, int c) { return cur[r + d * c]; }, ov, tid, 256, tid & 31);
  __syncthreads();
  CostCoef cc;
  cost_from_overlaps(q.cost, q.n, m, ov, cc);
  if (pen) {
    __shared__ double lred;
    if (tid == 0) lred = 0.0;
    __syncthreads();
    for (int off = 16; off > 0; off >>= 1) lsub += __shfl_xor_sync(0xffffffffu, lsub, off);
    if ((tid & 31) == 0) atomicAdd(&lred, lsub);
    __syncthreads();
    if (tid == 0 && q.J) {
      double Jp = 0.0;
      for (int p = 0; p < q.nranks; p++) Jp += q.C_all[(size_t)p * (2 * dm + 2) + 2 * dm];
      *q.J = cc.J + Jp - q.mu * lred;
    }
  } else
  if (tid == 0 && q.J) *q.J = cc.J;
  for (int e = tid; e < dm; e += 256) {
    const double2 t = reinterpret_cast<const double2*>(q.T)[e];
    const int c = e / d;
    double kr = cc.cr[0], ki = cc.ci[0];
    if (q.cost == QOC_COST_ZCAL_) {
#pragma unroll
      for (int k = 1; k < 4; k++) if (c == k) { kr = cc.cr[k]; ki = cc.ci[k]; }
    }
    nxt[e] = make_double2(kr * t.x - ki * t.y, kr * t.y + ki * t.x);
  }
  __syncthreads();
  { double2* t = cur; cur = nxt; nxt = t; }
  for (int p = q.nranks - 1; p >= 0; p--) {
    if (p == q.rank) for (int e = tid; e < dm; e += 256) reinterpret_cast<double2*>(q.lam_end)[e] = cur[e];
    if (p == q.rank) break;
    matvec(q.S_all + (size_t)p * 2 * d * d, true, cur, nxt);
    if (pen) {   // + c_p - dL_dx(x_start(p)): the costate leaving rank p - 1, before dL_dx of ITS last state (added by its sweep)
      const double2* cp = reinterpret_cast<const double2*>(q.C_all + (size_t)p * (2 * dm + 2));
      for (int e = tid; e < dm; e += 256) {
        double2 v = nxt[e];
        v.x += cp[e].x; v.y += cp[e].y;
        if (penalised(e % d, e / d)) {
          const double2 x = reinterpret_cast<const double2*>(q.xs)[(size_t)p * dm + e];
          v.x = fma(-2.0 * q.mu, x.x, v.x); v.y = fma(-2.0 * q.mu, x.y, v.y);
        }
        nxt[e] = v;
      }
      __syncthreads();
    }
    double2* t = cur; cur = nxt; nxt = t;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// K3N: one CTA per segment.  Warps [0, NT): the recurrences; warp NT: TMA producer; the remaining K3N_CW warps contract the
// Jacobians, each warp one (slice, control) item at a time, reading dU straight from global memory with many loads in
// flight.  The whole costate history of the segment stays in shared memory, so the contraction warps only wait for the
// recurrence through a progress counter and never hold it back.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int K3N_CW = 7;  // contraction warps

template <class C>
// (small shape classes: 2 CTAs per SM -- the per-segment recurrence is a latency chain, co-resident segments are what
// keeps the SM busy; d = 9 batch: 2.72 ms at one CTA per SM, 2.09 ms at two (96 registers), 2.55 ms at three (64 registers, spills);
// three contraction warps instead of seven at four CTAs per SM: 3.2 ms -- the contraction, not the recurrence, is what is slow at d = 9)
__global__ void __launch_bounds__((C::NT + 1 + K3N_CW) * 32, (C::NT <= 2) ? 2 : 1) k3n_kernel(K23Params p, int seg_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int S = C::S;
  constexpr int NT = C::NT;
  constexpr int RP = state_rows<C>();
  const int NTH = blockDim.x;
  const int d = p.d, m = p.m, nc = p.nc, W = 2 * m;
  const int slot_d = 2 * d * S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dm = d * m;
  const int sb = RP * W;
  const bool contract = p.want_grad != 0;

  double* base = reinterpret_cast<double*>(smem_raw);
  double* ringp = base;                                       // SW_NST slots (U_k)
  double* pad = ringp + (size_t)SW_NST * slot_d;
  const int pad_rows = k1_pad_rows<C>(d);
  double* xs = pad + pad_rows * S;                            // seg_cap states x_{k0+i}
  double* ls = xs + (size_t)seg_cap * sb;                     // seg_cap + 1 costates: ls[it] = lambda_{k1-it}
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(ls + (size_t)(seg_cap + 1) * sb);
  volatile int* ready = reinterpret_cast<volatile int*>(bars + 2 * SW_NST);
  TmaRing ring;
  ring.slots = ringp; ring.full = bars; ring.empty = bars + SW_NST; ring.slot_d = slot_d; ring.bytes = (unsigned)slot_d * 8u;
  {
    double2* z = reinterpret_cast<double2*>(base);
    const int total2 = (int)((ls + (size_t)(seg_cap + 1) * sb - base) / 2);
    for (int e = tid; e < total2; e += NTH) z[e] = make_double2(0.0, 0.0);
    if (tid == 0) { ring.init(NT); *ready = -1; }
  }
  __syncthreads();
  int gi = 0;
  const int units = d * S / 2;   // double2 units per plane

  for (int seg = blockIdx.x; seg < p.nseg; seg += gridDim.x) {
    const int b = seg / p.seg_per_pulse, si = seg - b * p.seg_per_pulse;
    const int k0 = (int)(((long long)si * p.nt) / p.seg_per_pulse);
    const int k1 = (int)(((long long)(si + 1) * p.nt) / p.seg_per_pulse);
    const int len = k1 - k0;
    if (len <= 0) continue;
    const size_t sl0 = (size_t)b * p.nt + k0;
    const int nF = len - 1, nB = contract ? len : 0, total = nF + nB;
    const int gi0 = gi;
    auto src = [&](int pos) -> const double* {
      const int i = pos - gi0;
      if (i < nF) return p.U + (sl0 + i) * slot_d;
      return p.U + (sl0 + (len - 1 - (i - nF))) * slot_d;
    };
    ist_from_global(xs, p.xs_start + (size_t)seg * 2 * dm, d, m, tid, NTH);
    if (contract) ist_from_global(ls, p.lam_end + (size_t)seg * 2 * dm, d, m, tid, NTH);
    __syncthreads();

    if (warp == NT) {
      // ---------------- producer: the whole operand stream of the segment, as far ahead as the ring allows ----------------
      if (lane == 0) for (int pos = gi0; pos < gi0 + total; pos++) ring.produce(pos, src(pos));
    } else if (warp < NT) {
      // ---------------- recurrence warps ----------------
      int li = 0;
      for (int i = 0; i < nF; i++, li++)   // forward sweep: xs[i+1] = U_{k0+i} xs[i]
        rec_step<C, false>(ring, gi0 + li, xs + (size_t)i * sb, xs + (size_t)(i + 1) * sb, d, m, warp, lane);
      for (int it = 0; it < nB; it++, li++) {   // backward sweep: ls[it+1] = U_k' ls[it], k = k1-1-it
        rec_step<C, true>(ring, gi0 + li, ls + (size_t)it * sb, ls + (size_t)(it + 1) * sb, d, m, warp, lane);
        // past the barrier of this step: all x (forward sweep) and ls[0..it] are complete
        if (tid == 0) { __threadfence_block(); *ready = it; }
        if (p.dbg && blockIdx.x == 0 && tid == 0 && li < p.dbg_steps) p.dbg[li] = clock64();
      }
      bar_rec(NT * 32);
    } else if (contract) {
      // ---------------- contraction warps ----------------
      const int cw = warp - NT - 1;
      if (nc == 1) {
      // a single control: one item per slice, four units in flight per lane (HBM-bound at d = 27)
      for (int item = cw; item < len * nc; item += K3N_CW) {
        const int it = item / nc, j = item - it * nc;
        const int k = k1 - 1 - it;
        const double* Lp = p.L + ((sl0 + (len - 1 - it)) * nc + j) * slot_d;
        // Re tr(lambda' dU x) = sum_{r,c} Re( dU[r][c] * w[r][c] ),  w[r][c] = sum_l x[c][l] conj(lambda[r][l])   (:217-223)
        // the Jacobian loads do not depend on lambda: issue the first batch before waiting for the recurrence
        constexpr int UB = 4;   // units in flight per lane
        double2 fre[UB], fim[UB];
#pragma unroll
        for (int i = 0; i < UB; i++) {
          const int u = lane + i * 32;
          if (u < units) {
            fre[i] = __ldg(reinterpret_cast<const double2*>(Lp) + u);
            fim[i] = __ldg(reinterpret_cast<const double2*>(Lp + d * S) + u);
          }
        }
        while (*ready < it) __nanosleep(64);
        __threadfence_block();
        const double* lr = ls + (size_t)it * sb;              // lambda_{k+1}
        const double* xk = xs + (size_t)(k - k0) * sb;        // x_k
        double s = 0.0;
        for (int u0 = 0; u0 < units; u0 += UB * 32) {
          double2 cre[UB], cim[UB];
#pragma unroll
          for (int i = 0; i < UB; i++) { cre[i] = fre[i]; cim[i] = fim[i]; }
#pragma unroll
          for (int i = 0; i < UB; i++) {
            const int u = u0 + UB * 32 + lane + i * 32;
            if (u < units) {
              fre[i] = __ldg(reinterpret_cast<const double2*>(Lp) + u);
              fim[i] = __ldg(reinterpret_cast<const double2*>(Lp + d * S) + u);
            }
          }
#pragma unroll
          for (int i = 0; i < UB; i++) {
            const int u = u0 + lane + i * 32;
            if (u < units) {
              const int r = u / (S / 2), c0 = 2 * (u - r * (S / 2));
              const double* lrow = lr + r * W;
              const double* x0p = xk + c0 * W;
              double w0r = 0.0, w0i = 0.0, w1r = 0.0, w1i = 0.0;
              for (int l = 0; l < m; l++) {
                const double2 lam = *reinterpret_cast<const double2*>(lrow + 2 * l);
                const double2 xa_ = *reinterpret_cast<const double2*>(x0p + 2 * l);
                const double2 xb_ = *reinterpret_cast<const double2*>(x0p + W + 2 * l);
                w0r = fma(xa_.x, lam.x, fma(xa_.y, lam.y, w0r)); w0i = fma(xa_.y, lam.x, fma(-xa_.x, lam.y, w0i));
                w1r = fma(xb_.x, lam.x, fma(xb_.y, lam.y, w1r)); w1i = fma(xb_.y, lam.x, fma(-xb_.x, lam.y, w1i));
              }
              s = fma(cre[i].x, w0r, fma(-cim[i].x, w0i, s));
              s = fma(cre[i].y, w1r, fma(-cim[i].y, w1i, s));
            }
          }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) p.dJdu[((size_t)b * p.nt + k) * nc + j] = s;
      }
          } else {
      // one item = (slice, pair of controls): w[r][c] = sum_l x[c][l] conj(lambda[r][l]) does not depend on the control, so
      // it is formed once per unit and contracted with both Jacobians of the pair
      const int npair = (nc + 1) / 2;
      for (int item = cw; item < len * npair; item += K3N_CW) {
        const int it = item / npair, j0 = 2 * (item - it * npair);
        const bool two = (j0 + 1 < nc);
        const int k = k1 - 1 - it;
        const double* Lp0 = p.L + ((sl0 + (len - 1 - it)) * nc + j0) * slot_d;
        const double* Lp1 = Lp0 + slot_d;   // read only when the pair is complete
        // Re tr(lambda' dU x) = sum_{r,c} Re( dU[r][c] * w[r][c] )   (:217-223)
        // the Jacobian loads do not depend on lambda: issue the first batch before waiting for the recurrence
        constexpr int UB = 2;   // units in flight per lane and control
        double2 fre[2][UB], fim[2][UB];
#pragma unroll
        for (int i = 0; i < UB; i++) { fre[1][i] = make_double2(0.0, 0.0); fim[1][i] = fre[1][i]; fre[0][i] = fre[1][i]; fim[0][i] = fre[1][i]; }
#pragma unroll
        for (int i = 0; i < UB; i++) {
          const int u = lane + i * 32;
          if (u < units) {
            fre[0][i] = __ldg(reinterpret_cast<const double2*>(Lp0) + u);
            fim[0][i] = __ldg(reinterpret_cast<const double2*>(Lp0 + d * S) + u);
            if (two) {
              fre[1][i] = __ldg(reinterpret_cast<const double2*>(Lp1) + u);
              fim[1][i] = __ldg(reinterpret_cast<const double2*>(Lp1 + d * S) + u);
            }
          }
        }
        while (*ready < it) __nanosleep(64);
        __threadfence_block();
        const double* lr = ls + (size_t)it * sb;              // lambda_{k+1}
        const double* xk = xs + (size_t)(k - k0) * sb;        // x_k
        double s0 = 0.0, s1 = 0.0;
        for (int u0 = 0; u0 < units; u0 += UB * 32) {
          double2 cre[2][UB], cim[2][UB];
#pragma unroll
          for (int i = 0; i < UB; i++) { cre[0][i] = fre[0][i]; cim[0][i] = fim[0][i]; cre[1][i] = fre[1][i]; cim[1][i] = fim[1][i]; }
#pragma unroll
          for (int i = 0; i < UB; i++) {
            const int u = u0 + UB * 32 + lane + i * 32;
            if (u < units) {
              fre[0][i] = __ldg(reinterpret_cast<const double2*>(Lp0) + u);
              fim[0][i] = __ldg(reinterpret_cast<const double2*>(Lp0 + d * S) + u);
              if (two) {
                fre[1][i] = __ldg(reinterpret_cast<const double2*>(Lp1) + u);
                fim[1][i] = __ldg(reinterpret_cast<const double2*>(Lp1 + d * S) + u);
              }
            }
          }
#pragma unroll
          for (int i = 0; i < UB; i++) {
            const int u = u0 + lane + i * 32;
            if (u < units) {
              const int r = u / (S / 2), c0 = 2 * (u - r * (S / 2));
              const double* lrow = lr + r * W;
              const double* x0p = xk + c0 * W;
              double w0r = 0.0, w0i = 0.0, w1r = 0.0, w1i = 0.0;
              for (int l = 0; l < m; l++) {
                const double2 lam = *reinterpret_cast<const double2*>(lrow + 2 * l);
                const double2 xa_ = *reinterpret_cast<const double2*>(x0p + 2 * l);
                const double2 xb_ = *reinterpret_cast<const double2*>(x0p + W + 2 * l);
                w0r = fma(xa_.x, lam.x, fma(xa_.y, lam.y, w0r)); w0i = fma(xa_.y, lam.x, fma(-xa_.x, lam.y, w0i));
                w1r = fma(xb_.x, lam.x, fma(xb_.y, lam.y, w1r)); w1i = fma(xb_.y, lam.x, fma(-xb_.x, lam.y, w1i));
              }
              s0 = fma(cre[0][i].x, w0r, fma(-cim[0][i].x, w0i, s0));
              s0 = fma(cre[0][i].y, w1r, fma(-cim[0][i].y, w1i, s0));
              s1 = fma(cre[1][i].x, w0r, fma(-cim[1][i].x, w0i, s1));
              s1 = fma(cre[1][i].y, w1r, fma(-cim[1][i].y, w1i, s1));
            }
          }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, off); s1 += __shfl_xor_sync(0xffffffffu, s1, off); }
        if (lane == 0) {
          p.dJdu[((size_t)b * p.nt + k) * nc + j0] = s0;
          if (two) p.dJdu[((size_t)b * p.nt + k) * nc + j0 + 1] = s1;
        }
      }
          }
    }
    __syncthreads();
    // states / costates of the segment to global memory (debug getters, time-sharding hand-over)
    if (p.store_states)
      for (int i = 0; i < len; i++) ist_to_global(p.X + ((size_t)b * (p.nt + 1) + k0 + i) * 2 * dm, xs + (size_t)i * sb, d, m, tid, NTH);
    if (contract && p.store_costates && p.LAM)
      for (int it = (k1 == p.nt ? 0 : 1); it <= len; it++)
        ist_to_global(p.LAM + ((size_t)b * (p.nt + 1) + (k1 - it)) * 2 * dm, ls + (size_t)it * sb, d, m, tid, NTH);
    gi += total;
    if (tid == 0) *ready = -1;
    __syncthreads();
  }
}

}  // namespace qoc

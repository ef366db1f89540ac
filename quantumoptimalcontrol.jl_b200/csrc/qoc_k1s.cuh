// qoc_k1s.cuh -- K1S: the small-dimension form of K1 (d <= 9: two qutrits, qubit pairs, ...), every gradient mode.
//
// Same arithmetic and the same outputs as k1_kernel (generator, Pade expm with scaling and squaring, block-triangular
// Frechet derivative or the reference's truncated Taylor Jacobian per control, running segment product; U_k, dU_k/du_j
// and Q_seg in the planar-slot format the sweeps read), for:  src/gradient_computations.jl:18-24 (X_k, exponential!) and
// :67, :177-213 (expm_jacobian!) or its exact counterpart.
//
// Why a second form: at d = 9 the DMMA classes run 16 x 16 x 12 tiles for 9 x 9 x 9 products (24 % of the tensor-pipe
// flops are useful) and every one of the ~26 products of a slice ends in a CTA-wide barrier that costs more than the
// product.  Here a group of NINE LANES owns a slice from generator to store, THREE slices per warp (27 of 32 lanes busy):
// every matrix lives in group-private shared memory (interleaved complex, 9 x 9), products are plain DFMA with a 3 x 3
// register block per lane (lane = 9 g + 3 br + bc), the Pade denominator is inverted in registers by the same nine lanes
// (partial pivoting, shuffles only), and phases are separated by __syncwarp.  Warps never talk to each other.
//
// Round 1 used one warp per slice with a 1 x 3 strip per lane: 4 LDS.128 per 12 DFMA.  A 128-bit shared load costs four
// wavefronts whatever the lanes broadcast, so that form was bound by shared-memory bandwidth (ncu: 1.84 G wavefronts on
// the 4096-pulse batch, 79 % of the kernel's SM cycles; FP64 pipe 29 %).  The 3 x 3 block needs 6 LDS.128 per 36 DFMA
// -- 2.7 bytes per DFMA instead of 5.3 -- and the matrix set is trimmed to nine per slice ([5/5] and [7/7] Pade only:
// beyond theta_7 the generator is scaled down further, which costs the same number of products as [13/13] would, see
// below; A4, W, the running segment product and, when squarings follow, the L_j pass through L2-resident global slots)
// so that 24 slices stay resident per SM: eight warps, two per scheduler.  One warp per scheduler cannot hide the FP64
// issue latency (measured: 2.9 cycles per DFMA in the dense product loops of a lone warp; 4 / 6 warps: 6.67 / 5.69 ms).
//
// Degrees: ||X||_1 <= theta_5 -> [5/5]; otherwise [7/7] with s = ceil(log2(||X||_1 / theta_7))+ squarings.  ([13/13] buys
// a factor 5.7 in norm for two extra products and four more per control, i.e. the price of two squarings: the same
// algorithmic flop count for norms in (theta_7, 4 theta_7], and it would need three more operand matrices per slice.)
// The three slices of a warp run in lock step, so they use one (degree, s): the largest any of them needs.
#pragma once
#include "qoc_k1.cuh"

namespace qoc {

constexpr int K1S_DMAX = 9;
constexpr int K1S_MSZ = K1S_DMAX * K1S_DMAX;  // complex elements per matrix
// Shared-memory layout of a matrix: BLOCK-major -- the 3 x 3 block (I, J) occupies the 9 consecutive elements from
// 9 (3 I + J), row-major inside the block.  A lane's own block is one contiguous 144-byte run (conflict-free block loads,
// stores and adjoint loads: the 8 lanes of a quarter-warp start in 8 different 16-byte bank quads), and the operand
// loads of a product step through it with unit strides.
__host__ __device__ __forceinline__ constexpr int k1s_idx(int r, int c) { return 9 * (3 * (r / 3) + c / 3) + 3 * (r % 3) + c % 3; }
constexpr int K1S_GPW = 3;                  // lane groups (slices) per warp
constexpr int K1S_MAXWPB = 8;
constexpr int K1S_MAXNC = 4;
// group-private matrices: seven, whatever nc.  Everything that is an operand only once per control passes through a
// temporary instead of owning a slot: A4 and W (homes: sX3_ / sX1_ while the Pade approximant is formed, then a
// per-group global scratch that stays in L2, prefetched into registers one product ahead), the running segment product Q
// and (only when squarings follow) the L_j (their own global slots).
enum : int { sA_ = 0, sA2_, sN_, sR_, sX1_, sX2_, sX3_, K1S_FIXED };
__host__ __device__ constexpr int k1s_mats(int) { return K1S_FIXED; }
// complex elements between the matrix sets of consecutive groups: (stride in 16-byte units) = 4 mod 8, so that the
// groups that share a quarter-warp hit disjoint bank quads in the operand loads (block rows are 27 = 3 mod 8 units apart,
// block columns 9 = 1 mod 8)
__host__ __device__ constexpr int k1s_group_stride(int nc) {
  int s = k1s_mats(nc) * K1S_MSZ;
  while (s % 8 != 4) s++;
  return s;
}
__host__ __device__ constexpr size_t k1s_smem_bytes(int nc, int wpb) {
  return ((size_t)wpb * K1S_GPW * k1s_group_stride(nc) + (size_t)(1 + nc) * K1S_MSZ) * 16 + 64;
}
static inline int k1s_warps_per_block(int nc, size_t smem_optin) {
  int w = K1S_MAXWPB;
  while (w > 0 && k1s_smem_bytes(nc, w) > smem_optin) w--;
  return w;
}

struct C9 {  // the lane's 3 x 3 block of a matrix
  double2 v[3][3];
};

struct K1SCtx {
  double2* base;       // group-private matrices
  const double2* E;    // CTA-shared control operators (nc matrices)
  int d, kd, lane, g, br, bc;   // kd: contraction length, d rounded up to whole blocks
  bool act;            // lane < 27: owns a block (lanes 27..31 shadow lane 26 and never store)
  __device__ __forceinline__ double2* M(int i) const { return base + i * K1S_MSZ; }
  __device__ __forceinline__ C9 ld(const double2* m) const {
    C9 x;
    const double2* q = m + 9 * (3 * br + bc);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) x.v[i][j] = q[3 * i + j];
    return x;
  }
  // the lane's block of m^dagger
  __device__ __forceinline__ C9 ldH(const double2* m) const {
    C9 x;
    const double2* q = m + 9 * (3 * bc + br);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const double2 v = q[3 * j + i];
        x.v[i][j] = make_double2(v.x, -v.y);
      }
    return x;
  }
  // Unmasked, unpredicated: rows / columns >= d of everything that is ever stored are zero by construction (zero-padded
  // inputs, products and linear combinations of zero-padded matrices, add_eye masked), and the shadow lanes 27..31
  // write the very values lane 26 writes to the very same addresses.
  __device__ __forceinline__ void st(double2* m, const C9& x) const {
    double2* q = m + 9 * (3 * br + bc);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) q[3 * i + j] = x.v[i][j];
  }
  // acc += A * B on the lane's block.  The k loop runs over whole block columns (rows / columns >= d of every matrix are
  // exactly zero), three k-steps per iteration, NOT fully unrolled: the kernel has ~30 product sites and the fully
  // unrolled form was 440 KB of SASS -- four warps walking it thrash the instruction cache.
  __device__ __forceinline__ void macc(C9& acc, const double2* A, const double2* B) const {
    const double2* ap = A + 27 * br;   // block (br, kb): + 9 per iteration
    const double2* bp = B + 9 * bc;    // block (kb, bc): + 27 per iteration
#pragma unroll 1
    for (int k0 = 0; k0 < kd; k0 += 3) {
#pragma unroll
      for (int kk = 0; kk < 3; kk++) {
        double2 a[3], b[3];
#pragma unroll
        for (int i = 0; i < 3; i++) a[i] = ap[3 * i + kk];
#pragma unroll
        for (int j = 0; j < 3; j++) b[j] = bp[3 * kk + j];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) {
            acc.v[i][j].x = fma(a[i].x, b[j].x, acc.v[i][j].x);
            acc.v[i][j].y = fma(a[i].x, b[j].y, acc.v[i][j].y);
            acc.v[i][j].x = fma(-a[i].y, b[j].y, acc.v[i][j].x);
            acc.v[i][j].y = fma(a[i].y, b[j].x, acc.v[i][j].y);
          }
      }
      ap += 9;
      bp += 27;
    }
  }
  static __device__ __forceinline__ C9 zero() {
    C9 x;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) x.v[i][j] = make_double2(0.0, 0.0);
    return x;
  }
  // x + cI * I
  __device__ __forceinline__ void add_eye(C9& x, double cI) const {
    if (br == bc) {
#pragma unroll
      for (int i = 0; i < 3; i++)
        if (3 * br + i < d) x.v[i][i].x += cI;
    }
  }
};

__device__ __forceinline__ C9 lin3(double c1, const C9& a, double c2, const C9& b, double c3, const C9& c) {
  C9 x;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      x.v[i][j].x = fma(c1, a.v[i][j].x, fma(c2, b.v[i][j].x, c3 * c.v[i][j].x));
      x.v[i][j].y = fma(c1, a.v[i][j].y, fma(c2, b.v[i][j].y, c3 * c.v[i][j].y));
    }
  return x;
}
__device__ __forceinline__ C9 lin2(double c1, const C9& a, double c2, const C9& b) {
  C9 x;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      x.v[i][j].x = fma(c1, a.v[i][j].x, c2 * b.v[i][j].x);
      x.v[i][j].y = fma(c1, a.v[i][j].y, c2 * b.v[i][j].y);
    }
  return x;
}
__device__ __forceinline__ void axpy9(C9& y, double c, const C9& a) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) { y.v[i][j].x = fma(c, a.v[i][j].x, y.v[i][j].x); y.v[i][j].y = fma(c, a.v[i][j].y, y.v[i][j].y); }
}
__device__ __forceinline__ void scale9(C9& y, double c) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) { y.v[i][j].x *= c; y.v[i][j].y *= c; }
}
// s ? a : b, element-wise (s is uniform over the lanes of a group)
__device__ __forceinline__ C9 sel9(bool s, const C9& a, const C9& b) {
  C9 x;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) x.v[i][j] = s ? a.v[i][j] : b.v[i][j];
  return x;
}
// v[idx] for idx in 0..2 without dynamic register indexing
__device__ __forceinline__ double2 pick3(const double2& a, const double2& b, const double2& c, int idx) {
  return idx == 0 ? a : (idx == 1 ? b : c);
}

// In-register Gauss-Jordan inverse with partial pivoting of the d x d matrix held as 3 x 3 blocks (lane = 9 g + 3 br + bc),
// one matrix per lane group, the three groups in lock step.  Rows are never swapped: the row that pivots column k
// remembers it (mycol); with W the array after the last step, A^-1[mycol_l][prow_j] = W[l][j].  Writes the inverse to
// `out`; returns false on an exactly zero pivot.
__device__ __forceinline__ bool k1s_inverse(const K1SCtx& c, C9 w, double2* out) {
  const unsigned FULL = 0xffffffffu;
  const int d = c.d, g9 = 9 * c.g, br = c.br, bc = c.bc;
  bool used[3];
  int mycol[3];
#pragma unroll
  for (int i = 0; i < 3; i++) { used[i] = (3 * br + i >= d); mycol[i] = -1; }   // rows >= d never pivot
  int prow_of[K1S_DMAX];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < K1S_DMAX; k++) {
    prow_of[k] = 0;
    if (k < d) {
      const int kb = k / 3, kj = k % 3;   // block column / element column that hold column k
      // ---- pivot search: the holder lanes (bc == kb) propose their best unused row, everybody reads the three proposals ----
      unsigned key = 0u;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double2 cv = w.v[i][kj];
        const double mag = cv.x * cv.x + cv.y * cv.y;
        // for mag >= 0 the high word of the double is a monotone key; the low 4 bits carry the row (smaller row wins ties)
        const unsigned ki = used[i] ? 0u : (((unsigned)__double2hiint(mag) & ~15u) | (unsigned)(8 - (3 * br + i)));
        key = ki > key ? ki : key;
      }
      unsigned best = __shfl_sync(FULL, key, g9 + kb);
      { const unsigned k1 = __shfl_sync(FULL, key, g9 + 3 + kb); best = k1 > best ? k1 : best; }
      { const unsigned k2 = __shfl_sync(FULL, key, g9 + 6 + kb); best = k2 > best ? k2 : best; }
      if ((best >> 4) == 0u) ok = false;
      const int prow = 8 - (int)(best & 15u);
      prow_of[k] = prow;
      const int pb = prow / 3, pi = prow - 3 * pb;
      // ---- 1 / pivot from the lane (pb, kb) ----
      double pir, pii;
      {
        const double2 cv = pick3(w.v[0][kj], w.v[1][kj], w.v[2][kj], pi);
        const double den = fast_rcp(cv.x * cv.x + cv.y * cv.y);
        pir = __shfl_sync(FULL, cv.x * den, g9 + 3 * pb + kb);
        pii = __shfl_sync(FULL, -cv.y * den, g9 + 3 * pb + kb);
      }
      // ---- multipliers of my rows (from the holder lane of my block row) and the pivot row's entries of my columns ----
      double gr[3], gi[3], rx[3], ry[3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double2 cv = w.v[i][kj];
        gr[i] = __shfl_sync(FULL, cv.x * pir - cv.y * pii, g9 + 3 * br + kb);
        gi[i] = __shfl_sync(FULL, cv.x * pii + cv.y * pir, g9 + 3 * br + kb);
      }
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const double2 rv = pick3(w.v[0][j], w.v[1][j], w.v[2][j], pi);
        rx[j] = __shfl_sync(FULL, rv.x, g9 + 3 * pb + bc);
        ry[j] = __shfl_sync(FULL, rv.y, g9 + 3 * pb + bc);
      }
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const bool isp = (3 * br + i == prow);
        if (isp) { used[i] = true; mycol[i] = k; gr[i] = -pir; gi[i] = -pii; }
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const double bx = isp ? 0.0 : w.v[i][j].x, by = isp ? 0.0 : w.v[i][j].y;
          w.v[i][j].x = fma(-gr[i], rx[j], fma(gi[i], ry[j], bx));
          w.v[i][j].y = fma(-gr[i], ry[j], fma(-gi[i], rx[j], by));
        }
        if (bc == kb) {
          w.v[i][kj].x = isp ? pir : -gr[i];
          w.v[i][kj].y = isp ? pii : -gi[i];
        }
      }
    }
  }
  if (c.act) {
#pragma unroll
    for (int i = 0; i < 3; i++)
      if (mycol[i] >= 0) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const int col = 3 * bc + j;
          if (col < d) {
            int pj = 0;
#pragma unroll
            for (int k = 0; k < K1S_DMAX; k++)
              if (k == col) pj = prow_of[k];
            out[k1s_idx(mycol[i], pj)] = w.v[i][j];
          }
        }
      }
  }
  return ok;
}

// planar-slot store (re plane then im plane, row stride S doubles, pad columns zero) of the lane's block
__device__ __forceinline__ void k1s_store_slot(const K1SCtx& c, bool on, double* slot, int S, const C9& x) {
  if (!c.act || !on) return;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int row = 3 * c.br + i;
    if (row < c.d) {
      double* re = slot + row * S;
      double* im = re + c.d * S;
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int col = 3 * c.bc + j;
        const bool v = col < c.d;
        re[col] = v ? x.v[i][j].x : 0.0;
        im[col] = v ? x.v[i][j].y : 0.0;
      }
      if (c.bc == 2)
        for (int col = K1S_DMAX; col < S; col++) { re[col] = 0.0; im[col] = 0.0; }
    }
  }
}

// the lane's block of a planar slot (rows / columns >= d read as zero; `on` false: zeros)
__device__ __forceinline__ C9 k1s_load_slot(const K1SCtx& c, bool on, const double* slot, int S) {
  C9 x;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int row = 3 * c.br + i;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int col = 3 * c.bc + j;
      const bool v = on && row < c.d && col < c.d;
      x.v[i][j] = v ? make_double2(slot[row * S + col], slot[c.d * S + row * S + col]) : make_double2(0.0, 0.0);
    }
  }
  return x;
}

// per-group global scratch (block-major like shared memory: a lane's block is 144 contiguous bytes, and every lane only
// ever re-reads the block it wrote itself)
__device__ __forceinline__ void k1s_scr_store(const K1SCtx& c, double2* m, const C9& x) {
  if (!c.act) return;
  double2* q = m + 9 * (3 * c.br + c.bc);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) q[3 * i + j] = x.v[i][j];
}
__device__ __forceinline__ C9 k1s_scr_load(const K1SCtx& c, const double2* m) {
  C9 x;
  const double2* q = m + 9 * (3 * c.br + c.bc);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) x.v[i][j] = q[3 * i + j];
  return x;
}

__global__ void __launch_bounds__(K1S_MAXWPB * 32, 1) k1s_kernel(K1Params p, int S) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned FULL = 0xffffffffu;
  const int d = p.d, nc = p.nc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  double2* sm = reinterpret_cast<double2*>(smem_raw);
  const int gstride = k1s_group_stride(nc);
  double2* shA0 = sm + (size_t)wpb * K1S_GPW * gstride;
  double2* shE = shA0 + K1S_MSZ;

  // zero everything once (pad rows / columns stay finite), then load A0 and the control operators (planar slots -> interleaved)
  for (int e = threadIdx.x; e < wpb * K1S_GPW * gstride + (1 + nc) * K1S_MSZ; e += blockDim.x) sm[e] = make_double2(0.0, 0.0);
  __syncthreads();
  for (int e = threadIdx.x; e < (1 + nc) * d * d; e += blockDim.x) {
    const int m = e / (d * d), q = e - m * d * d, r = q / d, cc = q - r * d;
    const double* src = (m == 0) ? p.A0p : p.Ap + (size_t)(m - 1) * 2 * d * S;
    shA0[m * K1S_MSZ + k1s_idx(r, cc)] = make_double2(src[r * S + cc], src[d * S + r * S + cc]);
  }
  __syncthreads();

  K1SCtx c;
  {
    const int l = lane < 27 ? lane : 26;
    c.g = l / 9;
    c.br = (l - 9 * c.g) / 3;
    c.bc = l - 9 * c.g - 3 * c.br;
  }
  c.base = sm + (size_t)(warp * K1S_GPW + c.g) * gstride;
  c.E = shE;
  double2* gA4 = reinterpret_cast<double2*>(p.scr) + (size_t)((blockIdx.x * wpb + warp) * K1S_GPW + c.g) * 2 * K1S_MSZ;
  double2* gW = gA4 + K1S_MSZ;
  c.d = d; c.lane = lane;
  c.kd = (d + 2) / 3 * 3;
  c.act = lane < 27;
  const size_t slot_d = (size_t)2 * d * S;

  WorkIter it;   // per lane group: its own sequence of segments
  it.init((blockIdx.x * wpb + warp) * K1S_GPW + c.g, p.nseg, p.seg_per_pulse, p.nt, gridDim.x * wpb * K1S_GPW);
  long long my_thirds = 0, my_exec = 0;
  bool all_ok = true;
  // (nothing is held in registers across a slice -- a 3 x 3 block is 36 of them: R is re-read from sR_, the running segment
  //  product from its global slot)
  const bool taylor = (p.order != 0);
  // (QOC_PADE13 = 1 sets the switch points negative to force [13/13] in k1_kernel: this form has no [13/13], keep its tables)
  const float theta7 = p.theta7 > 0.0 ? (float)p.theta7 : (taylor ? 0.95f : 0.783f);
  const float theta5 = p.theta7 > 0.0 ? (float)p.theta5 : (taylor ? 0.25f : 0.2f);

  double un[K1S_MAXNC];   // control amplitudes of the next work item (prefetched a whole slice ahead)
  auto load_u = [&](const WorkIter& w) {
#pragma unroll
    for (int j = 0; j < K1S_MAXNC; j++) un[j] = (j < nc && w.valid()) ? __ldg(p.u + ((size_t)w.b * p.nt + w.k) * nc + j) : 0.0;
  };
  load_u(it);
  while (__any_sync(FULL, it.valid())) {
    const bool on = it.valid();                       // this group has a slice in this round
    const bool first_of_seg = (it.k == it.k0);
    const size_t slice = on ? (size_t)it.b * p.nt + it.k : 0;
    const int seg = on ? it.seg : 0;
    // ---- generator X = A0 + sum_j u_j E_j, 1-norm, degree and scaling ----
    C9 x = c.ld(shA0);
#pragma unroll
    for (int j = 0; j < K1S_MAXNC; j++)
      if (j < nc) axpy9(x, un[j], c.ld(c.E + (size_t)j * K1S_MSZ));
    if (on) it.next();
    load_u(it);
    float ps;
    {
      // column sums: own 3 rows, then the three block rows of the group, then the max over columns / block columns / groups
      float cs[3];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const float ax = (float)x.v[i][j].x, ay = (float)x.v[i][j].y;
          s += (3 * c.br + i < d && 3 * c.bc + j < d) ? sqrtf(ax * ax + ay * ay) : 0.f;
        }
        cs[j] = s;
      }
      float cm = 0.f;
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const float t = __shfl_sync(FULL, cs[j], 9 * c.g + c.bc) + __shfl_sync(FULL, cs[j], 9 * c.g + 3 + c.bc) +
                        __shfl_sync(FULL, cs[j], 9 * c.g + 6 + c.bc);
        cm = fmaxf(cm, t);
      }
      if (!on) cm = 0.f;
      ps = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(cm)));   // one (degree, s) for the three slices of the warp
    }
    int sq = 0;
    const int qd = (ps <= theta5) ? 5 : 7;
    { float t = theta7; while (ps > t && sq < 60) { t *= 2.f; sq++; } }
    const double scl = __hiloint2double((1023 - sq) << 20, 0);  // 2^-sq
    {
      C9 a = x;
      scale9(a, scl);
      c.st(c.M(sA_), a);
    }
    __syncwarp();

    // ---- Pade numerator / denominator: U = A W, N = V - U ----
    const double* b = (qd == 7) ? c_b7 : c_b5;
    C9 a2 = K1SCtx::zero(), a4 = K1SCtx::zero(), w, nn, uu = K1SCtx::zero();
    c.macc(a2, c.M(sA_), c.M(sA_));
    c.st(c.M(sA2_), a2);
    __syncwarp();
    c.macc(a4, c.M(sA2_), c.M(sA2_));
    const bool frechet_on = p.want_jac && !taylor;
    c.st(c.M(sX3_), a4);
    if (frechet_on && qd == 7) k1s_scr_store(c, gA4, a4);
    __syncwarp();
    if (qd == 7) {
      C9 a6 = K1SCtx::zero();
      c.macc(a6, c.M(sA2_), c.M(sX3_));
      w = lin3(b[7], a6, b[5], a4, b[3], a2); c.add_eye(w, b[1]);
      nn = lin3(b[6], a6, b[4], a4, b[2], a2); c.add_eye(nn, b[0]);
    } else {
      w = lin2(b[5], a4, b[3], a2); c.add_eye(w, b[1]);
      nn = lin2(b[4], a4, b[2], a2); c.add_eye(nn, b[0]);
    }
    c.st(c.M(sX1_), w);
    if (frechet_on) k1s_scr_store(c, gW, w);
    __syncwarp();
    c.macc(uu, c.M(sA_), c.M(sX1_));
    c.st(c.M(sX2_), uu);      // U lives in a temporary: it is dead once R has been formed
    axpy9(nn, -1.0, uu);
    // ---- N^-1 (registers -> sN_) and R = I + 2 N^-1 U ----
    all_ok &= (k1s_inverse(c, nn, c.M(sN_)) || !on);
    __syncwarp();
    C9 rr9 = K1SCtx::zero();
    c.macc(rr9, c.M(sN_), c.M(sX2_));
    scale9(rr9, 2.0);
    c.add_eye(rr9, 1.0);
    c.st(c.M(sR_), rr9);
    if (sq == 0) k1s_store_slot(c, on, p.U + slice * slot_d, S, rr9);
    __syncwarp();

    // ---- the reference's truncated Taylor Jacobian (src/gradient_computations.jl:177-213, dt = 1, same association
    //      order): E + (EX + XE)/2 + (EX X + XE X + X XE)/6 + (EX X2 + XE X2 + X2 EX + X2 XE)/24 ----
    // slots: X (unscaled) in sA_ (X = 2^s A exactly: rescaled in place, the Pade operands are dead), EX in sX2_, XE in
    // sX3_, X2 in sA2_, scratch in sX1_
    if (p.want_jac && taylor) {
      if (sq > 0 && p.order >= 2) {
        C9 a = c.ld(c.M(sA_));
        scale9(a, __hiloint2double((1023 + sq) << 20, 0));
        c.st(c.M(sA_), a);      // (each lane rewrites the block it read)
        __syncwarp();
      }
      if (p.order >= 4) {
        C9 x2 = K1SCtx::zero();
        c.macc(x2, c.M(sA_), c.M(sA_));
        c.st(c.M(sA2_), x2);
      }
      for (int j = 0; j < nc; j++) {
        const double2* E = c.E + (size_t)j * K1S_MSZ;
        C9 out = c.ld(E);
        if (p.order >= 2) {
          C9 ex = K1SCtx::zero(), xe = K1SCtx::zero();
          c.macc(ex, E, c.M(sA_));
          c.st(c.M(sX2_), ex);
          if (p.skewh) {          // X, E skew-Hermitian: X E = (E X)^dagger
            __syncwarp();
            xe = c.ldH(c.M(sX2_));
          } else c.macc(xe, c.M(sA_), E);
          c.st(c.M(sX3_), xe);
          __syncwarp();
          axpy9(out, 0.5, ex);
          axpy9(out, 0.5, xe);
          if (p.order >= 3) {
            C9 t3 = K1SCtx::zero();
            c.macc(t3, c.M(sX2_), c.M(sA_));
            if (p.skewh) {        // X (X E) = -((E X) X)^dagger
              c.st(c.M(sX1_), t3);
              __syncwarp();
              axpy9(t3, -1.0, c.ldH(c.M(sX1_)));
            } else c.macc(t3, c.M(sA_), c.M(sX3_));
            c.macc(t3, c.M(sX3_), c.M(sA_));
            axpy9(out, 1.0 / 6.0, t3);
          }
          if (p.order >= 4) {
            C9 t4 = K1SCtx::zero();
            c.macc(t4, c.M(sX2_), c.M(sA2_));
            c.macc(t4, c.M(sX3_), c.M(sA2_));
            c.macc(t4, c.M(sA2_), c.M(sX2_));
            c.macc(t4, c.M(sA2_), c.M(sX3_));
            axpy9(out, 1.0 / 24.0, t4);
          }
          __syncwarp();   // every lane has read EX / XE before the next control overwrites them
        }
        k1s_store_slot(c, on, p.L + (slice * nc + j) * slot_d, S, out);
      }
    }
    // ---- exact Frechet derivative per control (Al-Mohy & Higham 2009, Alg. 6.4; E unscaled, 2^-s on the result) ----
    // temporaries: sX1_ = M2 then D, sX2_ = M4 then W then rhs, sX3_ = scratch of the adjoint shortcuts then A4 then Lw
    if (p.want_jac && !taylor) {
      for (int j = 0; j < nc; j++) {
        const double2* E = c.E + (size_t)j * K1S_MSZ;
        C9 m2 = K1SCtx::zero(), m4 = K1SCtx::zero(), lw, lv, lu = K1SCtx::zero();
        // skew-Hermitian generators: A E + E A = P + P^dagger with P = A E, and A2 M2 + M2 A2 = P + P^dagger with
        // P = A2 M2 (A2 and M2 Hermitian): one product each instead of two, the adjoint read back from shared memory
        c.macc(m2, c.M(sA_), E);
        if (p.skewh) {
          c.st(c.M(sX3_), m2);
          __syncwarp();
          axpy9(m2, 1.0, c.ldH(c.M(sX3_)));
        } else c.macc(m2, E, c.M(sA_));
        c.st(c.M(sX1_), m2);
        __syncwarp();
        C9 a4g;
        if (qd == 7) a4g = k1s_scr_load(c, gA4);   // in flight during the next product
        c.macc(m4, c.M(sA2_), c.M(sX1_));
        if (p.skewh) {
          c.st(c.M(sX3_), m4);
          __syncwarp();
          axpy9(m4, 1.0, c.ldH(c.M(sX3_)));
        } else c.macc(m4, c.M(sX1_), c.M(sA2_));
        c.st(c.M(sX2_), m4);
        __syncwarp();
        if (qd == 7) {
          c.st(c.M(sX3_), a4g);      // A4 back from its scratch
          __syncwarp();
          C9 m6 = K1SCtx::zero();
          c.macc(m6, c.M(sX3_), c.M(sX1_));
          c.macc(m6, c.M(sX2_), c.M(sA2_));
          lw = lin3(b[7], m6, b[5], m4, b[3], m2);
          lv = lin3(b[6], m6, b[4], m4, b[2], m2);
          __syncwarp();              // every lane has read A4
        } else {
          lw = lin2(b[5], m4, b[3], m2);
          lv = lin2(b[4], m4, b[2], m2);
        }
        c.st(c.M(sX3_), lw);
        __syncwarp();
        {
          const C9 wg = k1s_scr_load(c, gW);   // in flight during A Lw
          c.macc(lu, c.M(sA_), c.M(sX3_));
          c.st(c.M(sX2_), wg);                 // (M4 is dead)
        }
        __syncwarp();
        c.macc(lu, E, c.M(sX2_));
        // rhs = (Lu + Lv) + (Lu - Lv) R ;  L = 2^-s N^-1 rhs
        C9 dd = lu, rhs = lu;
        axpy9(dd, -1.0, lv);
        axpy9(rhs, 1.0, lv);
        c.st(c.M(sX1_), dd);       // D (M2 is dead)
        __syncwarp();
        c.macc(rhs, c.M(sX1_), c.M(sR_));
        c.st(c.M(sX2_), rhs);      // (M4 is dead)
        __syncwarp();
        C9 L = K1SCtx::zero();
        c.macc(L, c.M(sN_), c.M(sX2_));
        scale9(L, scl);
        k1s_store_slot(c, on, p.L + (slice * nc + j) * slot_d, S, L);   // final when sq == 0, else squared below
        __syncwarp();
      }
    }
    // ---- squarings: L <- R L + L R ; R <- R R ----
    for (int t2 = 0; t2 < sq; t2++) {
      if (p.want_jac && !taylor)
        for (int j = 0; j < nc; j++) {
          // L_j comes back from its global slot (each lane re-reads the block it wrote itself) through a temporary
          double* Lslot = p.L + (slice * nc + j) * slot_d;
          c.st(c.M(sX1_), k1s_load_slot(c, on, Lslot, S));
          __syncwarp();
          C9 ln = K1SCtx::zero();
          c.macc(ln, c.M(sR_), c.M(sX1_));
          c.macc(ln, c.M(sX1_), c.M(sR_));
          k1s_store_slot(c, on, Lslot, S, ln);
          __syncwarp();   // every lane has read L_j
        }
      C9 r2 = K1SCtx::zero();
      c.macc(r2, c.M(sR_), c.M(sR_));
      __syncwarp();
      c.st(c.M(sR_), r2);
      if (t2 + 1 == sq) k1s_store_slot(c, on, p.U + slice * slot_d, S, r2);
      __syncwarp();
    }

    // ---- level-1 scan: Q <- U_k Q (a group that starts a segment takes U_k itself) ----
    {
      // the running product sits in the segment's own Q slot (L2): each lane re-reads the block it wrote a slice ago
      double* Qslot = p.Q + (size_t)seg * slot_d;
      C9 qn = K1SCtx::zero();
      if (__any_sync(FULL, on && !first_of_seg)) {
        c.st(c.M(sX1_), k1s_load_slot(c, on && !first_of_seg, Qslot, S));
        __syncwarp();
        c.macc(qn, c.M(sR_), c.M(sX1_));
      }
      if (first_of_seg) qn = c.ld(c.M(sR_));
      k1s_store_slot(c, on, Qslot, S, qn);
      __syncwarp();
    }
    if (on) {
      const int pi_q = qd == 7 ? 4 : 3;
      const int G = !p.want_jac ? 0 : taylor ? (p.order == 1 ? 0 : p.order == 2 ? 2 : p.order == 3 ? 5 : 10) : (2 * pi_q + 2 * sq + 2);
      my_thirds += 3 * (pi_q + sq) + 4 + 3 * nc * G;
      // executed: the skew-Hermitian shortcuts save two products per control (Frechet) / one or two (Taylor order 2 / >= 3);
      // Taylor order 4 forms X^2 once per slice instead of once per control (:206)
      const int saved = (!p.want_jac || !p.skewh) ? 0 : taylor ? (p.order >= 3 ? 2 : p.order == 2 ? 1 : 0) : 2;
      my_exec += 3 * (pi_q + sq) + 4 + 3 * nc * (G - saved) - ((p.want_jac && taylor && p.order >= 4) ? 3 * (nc - 1) : 0);
    }
  }
  {
    const bool leader = c.act && c.br == 0 && c.bc == 0;
    if (__any_sync(FULL, !all_ok) && lane == 0) atomicExch(p.status, 8);
    if (leader && my_thirds != 0) {
      const double f = (8.0 * d * d * (double)d) * ((double)my_thirds / 3.0);
      atomicAdd(p.flops, f);
      atomicAdd(p.flops + 1, (8.0 * d * d * (double)d) * ((double)my_exec / 3.0));   // scalar DFMA products: nothing is padded
    }
  }
}

}  // namespace qoc

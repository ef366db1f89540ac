// qoc_k1s.cuh -- K1S: the small-dimension form of K1 (d <= 9: two qutrits, qubit pairs, ...), every gradient mode.
//
// Same arithmetic and the same outputs as k1_kernel (generator, Pade [5/5] / [7/7] / [13/13] expm with scaling and
// squaring, block-triangular Frechet derivative per control, running segment product; U_k, dU_k/du_j and Q_seg in the
// planar-slot format the sweeps read), for:  src/gradient_computations.jl:18-24 (X_k, exponential!) and the exact
// counterpart of :67, :177-213 (expm_jacobian!).
//
// Why a second form: at d = 9 the DMMA classes run 16 x 16 x 12 tiles for 9 x 9 x 9 products (24 % of the tensor-pipe
// flops are useful) and every one of the ~26 products of a slice ends in a CTA-wide barrier that costs more than the
// product.  Here ONE WARP owns a slice from generator to store: every matrix lives in warp-private shared memory
// (interleaved complex, 9 x 9), products are plain DFMA with a 1 x 3 register strip per lane (lane = 3 * row + strip, 27
// of 32 lanes busy, 4 LDS.128 per 12 DFMA), the Pade denominator is inverted in registers by the same 27 lanes, and
// phases are separated by __syncwarp only.  Warps never talk to each other; a CTA is just 8 independent workers sharing
// the control operators.  (Measured DFMA peak 34 TFLOP/s vs 37 for DMMA: without the padding the scalar pipe wins 3x.)
#pragma once
#include "qoc_k1.cuh"

namespace qoc {

constexpr int K1S_DMAX = 9;
constexpr int K1S_RS = 9;                 // row stride (complex elements)
constexpr int K1S_MSZ = K1S_DMAX * K1S_RS;  // complex elements per matrix
constexpr int K1S_WPB = 8;                // warps (independent workers) per CTA
constexpr int K1S_MAXNC = 4;
// warp-private matrices
enum : int { sA_ = 0, sA2_, sA4_, sA6_, sW1_, sZ1_, sW_, sU_, sN_, sR_, sQ_, sM2_, sM4_, sM6_, sT1_, sLw_, sS_, K1S_FIXED };
// sU_ doubles as T2 (Lz1) once R has been formed; L_j of control j sits behind the fixed ones
__host__ __device__ constexpr int k1s_mats(int nc) { return K1S_FIXED + nc; }
__host__ __device__ constexpr size_t k1s_smem_bytes(int nc) {
  return (size_t)(K1S_WPB * k1s_mats(nc) + 1 + nc) * K1S_MSZ * 16 + (size_t)K1S_WPB * 96 * 4 + 64;
}

struct C3 {  // the lane's 1 x 3 strip of a matrix
  double2 v[3];
};

struct K1SCtx {
  double2* base;       // warp-private matrices
  const double2* E;    // CTA-shared control operators (nc matrices)
  float* nrm;          // warp-private 96 floats (norm scratch)
  int d, lane, r, rr, cs;
  bool act;            // this lane owns a row of the matrix
  __device__ __forceinline__ double2* M(int i) const { return base + i * K1S_MSZ; }
  __device__ __forceinline__ C3 ld(const double2* m) const {
    C3 x;
#pragma unroll
    for (int e = 0; e < 3; e++) x.v[e] = m[rr * K1S_RS + cs + e];
    return x;
  }
  // the lane's strip of m^dagger
  __device__ __forceinline__ C3 ldH(const double2* m) const {
    C3 x;
#pragma unroll
    for (int e = 0; e < 3; e++) {
      const double2 v = m[(cs + e) * K1S_RS + rr];
      x.v[e] = make_double2(v.x, -v.y);
    }
    return x;
  }
  __device__ __forceinline__ void st(double2* m, const C3& x) const {
    if (act) {
#pragma unroll
      for (int e = 0; e < 3; e++) m[r * K1S_RS + cs + e] = (cs + e < d) ? x.v[e] : make_double2(0.0, 0.0);
    }
  }
  // acc += A * B on the lane's strip.  The k loop always runs over the full 9 (rows / columns >= d of every matrix are
  // exactly zero: masked stores, zero-initialised shared memory), so it is branch-free and the loads can be hoisted.
  // KD = 5 / 9: compile-time contraction length (d <= 5 stops early).
  template <int KD>
  __device__ __forceinline__ void macc_k(C3& acc, const double2* A, const double2* B) const {
    const double2* ap = A + rr * K1S_RS;
    const double2* bp = B + cs;
#pragma unroll
    for (int k = 0; k < KD; k++) {
      const double2 a = ap[k];
#pragma unroll
      for (int e = 0; e < 3; e++) {
        const double2 b = bp[k * K1S_RS + e];
        acc.v[e].x = fma(a.x, b.x, acc.v[e].x);
        acc.v[e].y = fma(a.x, b.y, acc.v[e].y);
        acc.v[e].x = fma(-a.y, b.y, acc.v[e].x);
        acc.v[e].y = fma(a.y, b.x, acc.v[e].y);
      }
    }
  }
  __device__ __forceinline__ void macc(C3& acc, const double2* A, const double2* B) const {
    if (d <= 5) macc_k<5>(acc, A, B);
    else macc_k<K1S_DMAX>(acc, A, B);
  }
  static __device__ __forceinline__ C3 zero() {
    C3 x;
#pragma unroll
    for (int e = 0; e < 3; e++) x.v[e] = make_double2(0.0, 0.0);
    return x;
  }
  // x + cI * I
  __device__ __forceinline__ void add_eye(C3& x, double cI) const {
#pragma unroll
    for (int e = 0; e < 3; e++)
      if (cs + e == r) x.v[e].x += cI;
  }
};

__device__ __forceinline__ C3 lin3(double c1, const C3& a, double c2, const C3& b, double c3, const C3& c) {
  C3 x;
#pragma unroll
  for (int e = 0; e < 3; e++) {
    x.v[e].x = fma(c1, a.v[e].x, fma(c2, b.v[e].x, c3 * c.v[e].x));
    x.v[e].y = fma(c1, a.v[e].y, fma(c2, b.v[e].y, c3 * c.v[e].y));
  }
  return x;
}
__device__ __forceinline__ C3 lin2(double c1, const C3& a, double c2, const C3& b) {
  C3 x;
#pragma unroll
  for (int e = 0; e < 3; e++) {
    x.v[e].x = fma(c1, a.v[e].x, c2 * b.v[e].x);
    x.v[e].y = fma(c1, a.v[e].y, c2 * b.v[e].y);
  }
  return x;
}
__device__ __forceinline__ void axpy3(C3& y, double c, const C3& a) {
#pragma unroll
  for (int e = 0; e < 3; e++) { y.v[e].x = fma(c, a.v[e].x, y.v[e].x); y.v[e].y = fma(c, a.v[e].y, y.v[e].y); }
}

// In-register Gauss-Jordan inverse with partial pivoting of the d x d matrix held as strips (lane = 3 * row + strip).
// Rows are never swapped: the row that pivots column k remembers it (mycol); with W the array after the last step,
// A^-1[mycol_l][prow_j] = W[l][j].  Writes the inverse to `out`; returns false on an exactly zero pivot.
__device__ __forceinline__ bool k1s_inverse(const K1SCtx& c, C3 w, double2* out) {
  const int d = c.d, lane = c.lane, strip = lane - 3 * (lane / 3);
  const int row = lane / 3;
  bool used = !c.act;     // rows >= d never pivot
  int mycol = -1;
  int prow_of[K1S_DMAX];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < K1S_DMAX; k++) {
    prow_of[k] = 0;
    if (k < d) {
      const int ks = k / 3, ke = k % 3;   // strip / element that hold column k
      const double2 cv = w.v[ke];
      const double mag = cv.x * cv.x + cv.y * cv.y;
      const bool holder = (strip == ks);
      const unsigned key = (holder && !used) ? (((unsigned)__double2hiint(mag) & ~31u) | (unsigned)(31 - lane)) : 0u;
      const unsigned best = __reduce_max_sync(0xffffffffu, key);
      const int pl = 31 - (int)(best & 31u);     // holder lane of the pivot row
      if ((best >> 5) == 0u) ok = false;
      const int prow = pl / 3;
      prow_of[k] = prow;
      const double den = fast_rcp(mag);
      const double ir = cv.x * den, ii = -cv.y * den;
      const double pir = __shfl_sync(0xffffffffu, ir, pl), pii = __shfl_sync(0xffffffffu, ii, pl);   // 1 / pivot
      // multiplier of my row: from the holder lane of my row
      double gr = cv.x * pir - cv.y * pii, gi = cv.x * pii + cv.y * pir;
      const int hl = 3 * row + ks;
      gr = __shfl_sync(0xffffffffu, gr, hl & 31);
      gi = __shfl_sync(0xffffffffu, gi, hl & 31);
      const bool isp = (row == prow);
      if (isp) { used = true; mycol = k; gr = -pir; gi = -pii; }
      // pivot row entries of my strip
      const int src = 3 * prow + strip;
#pragma unroll
      for (int e = 0; e < 3; e++) {
        const double rx = __shfl_sync(0xffffffffu, w.v[e].x, src), ry = __shfl_sync(0xffffffffu, w.v[e].y, src);
        const double bx = isp ? 0.0 : w.v[e].x, by = isp ? 0.0 : w.v[e].y;
        w.v[e].x = fma(-gr, rx, fma(gi, ry, bx));
        w.v[e].y = fma(-gr, ry, fma(-gi, rx, by));
      }
      if (holder) {
        w.v[ke].x = isp ? pir : -gr;
        w.v[ke].y = isp ? pii : -gi;
      }
    }
  }
  if (c.act && mycol >= 0) {
#pragma unroll
    for (int e = 0; e < 3; e++) {
      const int j = c.cs + e;
      if (j < d) {
        int pj = 0;
#pragma unroll
        for (int k = 0; k < K1S_DMAX; k++)
          if (k == j) pj = prow_of[k];
        out[mycol * K1S_RS + pj] = w.v[e];
      }
    }
  }
  return ok;
}

// planar-slot store (re plane then im plane, row stride S doubles, pad columns zero) of the lane's strip
__device__ __forceinline__ void k1s_store_slot(const K1SCtx& c, double* slot, int S, const C3& x) {
  if (!c.act) return;
  double* re = slot + c.r * S;
  double* im = re + c.d * S;
#pragma unroll
  for (int e = 0; e < 3; e++) {
    const int col = c.cs + e;
    const bool v = col < c.d;
    re[col] = v ? x.v[e].x : 0.0;
    im[col] = v ? x.v[e].y : 0.0;
  }
  if (c.cs == 6)
    for (int col = K1S_DMAX; col < S; col++) { re[col] = 0.0; im[col] = 0.0; }
}

__global__ void __launch_bounds__(K1S_WPB * 32, 1) k1s_kernel(K1Params p, int S) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = p.d, nc = p.nc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double2* sm = reinterpret_cast<double2*>(smem_raw);
  const int nm = k1s_mats(nc);
  double2* shA0 = sm + (size_t)K1S_WPB * nm * K1S_MSZ;
  double2* shE = shA0 + K1S_MSZ;
  float* nrm_all = reinterpret_cast<float*>(shE + (size_t)nc * K1S_MSZ);

  // zero everything once (pad rows / columns stay finite), then load A0 and the control operators (planar slots -> interleaved)
  for (int e = threadIdx.x; e < (K1S_WPB * nm + 1 + nc) * K1S_MSZ; e += blockDim.x) sm[e] = make_double2(0.0, 0.0);
  __syncthreads();
  for (int e = threadIdx.x; e < (1 + nc) * d * d; e += blockDim.x) {
    const int m = e / (d * d), q = e - m * d * d, r = q / d, cc = q - r * d;
    const double* src = (m == 0) ? p.A0p : p.Ap + (size_t)(m - 1) * 2 * d * S;
    shA0[m * K1S_MSZ + r * K1S_RS + cc] = make_double2(src[r * S + cc], src[d * S + r * S + cc]);
  }
  __syncthreads();

  K1SCtx c;
  c.base = sm + (size_t)warp * nm * K1S_MSZ;
  c.E = shE;
  c.nrm = nrm_all + warp * 96;
  c.d = d; c.lane = lane; c.r = lane / 3; c.cs = 3 * (lane - 3 * c.r);
  c.act = c.r < d;
  c.rr = c.act ? c.r : 0;
  const C3 a0 = c.ld(shA0);
  const size_t slot_d = (size_t)2 * d * S;

  WorkIter it;
  it.init(blockIdx.x * K1S_WPB + warp, p.nseg, p.seg_per_pulse, p.nt, gridDim.x * K1S_WPB);
  long long my_thirds = 0, my_exec = 0;
  bool all_ok = true;
  C3 q3 = K1SCtx::zero();   // running segment product, the lane's strip (also in sQ_ for use as an operand)

  double un[K1S_MAXNC];   // control amplitudes of the next work item (prefetched a whole slice ahead)
  auto load_u = [&](const WorkIter& w) {
#pragma unroll
    for (int j = 0; j < K1S_MAXNC; j++) un[j] = (j < nc && w.valid()) ? __ldg(p.u + ((size_t)w.b * p.nt + w.k) * nc + j) : 0.0;
  };
  load_u(it);
  while (it.valid()) {
    const bool first_of_seg = (it.k == it.k0), last_of_seg = (it.k + 1 >= it.k1);
    const size_t slice = (size_t)it.b * p.nt + it.k;
    const int seg = it.seg;
    // ---- generator X = A0 + sum_j u_j E_j, 1-norm, degree and scaling ----
    C3 x = a0;
#pragma unroll
    for (int j = 0; j < K1S_MAXNC; j++)
      if (j < nc) axpy3(x, un[j], c.ld(c.E + (size_t)j * K1S_MSZ));
    it.next();
    load_u(it);
    {
#pragma unroll
      for (int e = 0; e < 3; e++) {
        const float ax = (float)x.v[e].x, ay = (float)x.v[e].y;
        c.nrm[lane * 3 + e] = (c.act && c.cs + e < d) ? sqrtf(ax * ax + ay * ay) : 0.f;
      }
      __syncwarp();
      float cs_ = 0.f;   // lane l < 9 sums column l: element (row rw, col l) sits at nrm[(3 rw + l / 3) * 3 + l % 3]
      if (lane < K1S_DMAX)
        for (int rw = 0; rw < K1S_DMAX; rw++) cs_ += c.nrm[(3 * rw + lane / 3) * 3 + lane % 3];
      const float ps = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(cs_)));
      __syncwarp();
      int sq = 0, qd = 13;
      float t = (float)p.theta13;
      if (ps <= (float)p.theta5) qd = 5;
      else if (ps <= (float)p.theta7) qd = 7;
      else { while (ps > t && sq < 60) { t *= 2.f; sq++; } }
      const double scl = __hiloint2double((1023 - sq) << 20, 0);  // 2^-sq
      C3 a = x;
#pragma unroll
      for (int e = 0; e < 3; e++) { a.v[e].x *= scl; a.v[e].y *= scl; }
      c.st(c.M(sA_), a);
      const bool taylor = (p.order != 0);
      if (taylor && p.want_jac && p.order >= 2) c.st(c.M(sS_), x);   // the unscaled generator, operand of the Taylor Jacobian
      __syncwarp();

      // ---- Pade numerator / denominator: U = A W, N = V - U ----
      const double* b = (qd == 13) ? c_b13 : (qd == 7) ? c_b7 : c_b5;
      C3 a2 = K1SCtx::zero(), a4 = K1SCtx::zero(), a6 = K1SCtx::zero(), w = K1SCtx::zero(), uu = K1SCtx::zero(), nn;
      c.macc(a2, c.M(sA_), c.M(sA_));
      c.st(c.M(sA2_), a2);
      __syncwarp();
      c.macc(a4, c.M(sA2_), c.M(sA2_));
      c.st(c.M(sA4_), a4);
      __syncwarp();
      if (qd >= 7) {
        c.macc(a6, c.M(sA2_), c.M(sA4_));
        c.st(c.M(sA6_), a6);
      }
      if (qd == 13) {
        c.st(c.M(sW1_), lin3(b[13], a6, b[11], a4, b[9], a2));
        c.st(c.M(sZ1_), lin3(b[12], a6, b[10], a4, b[8], a2));
        __syncwarp();
        C3 v = K1SCtx::zero();
        c.macc(w, c.M(sA6_), c.M(sW1_));
        c.macc(v, c.M(sA6_), c.M(sZ1_));
        { C3 l = lin3(b[7], a6, b[5], a4, b[3], a2); axpy3(w, 1.0, l); c.add_eye(w, b[1]); }
        { C3 l = lin3(b[6], a6, b[4], a4, b[2], a2); axpy3(v, 1.0, l); c.add_eye(v, b[0]); }
        nn = v;
      } else if (qd == 7) {
        w = lin3(b[7], a6, b[5], a4, b[3], a2); c.add_eye(w, b[1]);
        nn = lin3(b[6], a6, b[4], a4, b[2], a2); c.add_eye(nn, b[0]);
      } else {
        w = lin2(b[5], a4, b[3], a2); c.add_eye(w, b[1]);
        nn = lin2(b[4], a4, b[2], a2); c.add_eye(nn, b[0]);
      }
      c.st(c.M(sW_), w);
      __syncwarp();
      c.macc(uu, c.M(sA_), c.M(sW_));
      c.st(c.M(sU_), uu);
      axpy3(nn, -1.0, uu);
      // ---- N^-1 (registers -> sN_) and R = I + 2 N^-1 U ----
      all_ok &= k1s_inverse(c, nn, c.M(sN_));
      __syncwarp();
      C3 rr3 = K1SCtx::zero();
      c.macc(rr3, c.M(sN_), c.M(sU_));
#pragma unroll
      for (int e = 0; e < 3; e++) { rr3.v[e].x *= 2.0; rr3.v[e].y *= 2.0; }
      c.add_eye(rr3, 1.0);
      c.st(c.M(sR_), rr3);
      __syncwarp();

      // ---- the reference's truncated Taylor Jacobian (src/gradient_computations.jl:177-213, dt = 1, same association
      //      order): E + (EX + XE)/2 + (EX X + XE X + X XE)/6 + (EX X2 + XE X2 + X2 EX + X2 XE)/24 ----
      if (p.want_jac && taylor) {
        for (int j = 0; j < nc; j++) {
          const double2* E = c.E + (size_t)j * K1S_MSZ;
          C3 out = c.ld(E);
          if (p.order >= 2) {
            C3 ex = K1SCtx::zero(), xe = K1SCtx::zero();
            c.macc(ex, E, c.M(sS_));
            c.st(c.M(sM2_), ex);
            if (p.skewh) {          // X, E skew-Hermitian: X E = (E X)^dagger
              __syncwarp();
              xe = c.ldH(c.M(sM2_));
            } else c.macc(xe, c.M(sS_), E);
            c.st(c.M(sM4_), xe);
            if (p.order >= 4) {
              C3 x2 = K1SCtx::zero();
              c.macc(x2, c.M(sS_), c.M(sS_));
              c.st(c.M(sM6_), x2);
            }
            __syncwarp();
            axpy3(out, 0.5, ex);
            axpy3(out, 0.5, xe);
            if (p.order >= 3) {
              C3 t3 = K1SCtx::zero();
              c.macc(t3, c.M(sM2_), c.M(sS_));
              if (p.skewh) {        // X (X E) = -((E X) X)^dagger
                c.st(c.M(sT1_), t3);
                __syncwarp();
                axpy3(t3, -1.0, c.ldH(c.M(sT1_)));
              } else c.macc(t3, c.M(sS_), c.M(sM4_));
              c.macc(t3, c.M(sM4_), c.M(sS_));
              axpy3(out, 1.0 / 6.0, t3);
            }
            if (p.order >= 4) {
              C3 t4 = K1SCtx::zero();
              c.macc(t4, c.M(sM2_), c.M(sM6_));
              c.macc(t4, c.M(sM4_), c.M(sM6_));
              c.macc(t4, c.M(sM6_), c.M(sM2_));
              c.macc(t4, c.M(sM6_), c.M(sM4_));
              axpy3(out, 1.0 / 24.0, t4);
            }
            __syncwarp();   // every lane has read EX / XE / X2 before the next control overwrites them
          }
          k1s_store_slot(c, p.L + (slice * nc + j) * slot_d, S, out);
        }
      }
      // ---- exact Frechet derivative per control (Al-Mohy & Higham 2009, Alg. 6.4; E unscaled, 2^-s on the result) ----
      if (p.want_jac && !taylor) {
        for (int j = 0; j < nc; j++) {
          const double2* E = c.E + (size_t)j * K1S_MSZ;
          C3 m2 = K1SCtx::zero(), m4 = K1SCtx::zero(), m6 = K1SCtx::zero(), lw, lv, lu = K1SCtx::zero();
          // skew-Hermitian generators: A E + E A = P + P^dagger with P = A E, and A2 M2 + M2 A2 = P + P^dagger with
          // P = A2 M2 (A2 and M2 Hermitian): one product each instead of two, the adjoint read back from shared memory
          c.macc(m2, c.M(sA_), E);
          if (p.skewh) {
            c.st(c.M(sM6_), m2);
            __syncwarp();
            axpy3(m2, 1.0, c.ldH(c.M(sM6_)));
          } else c.macc(m2, E, c.M(sA_));
          c.st(c.M(sM2_), m2);
          __syncwarp();
          c.macc(m4, c.M(sA2_), c.M(sM2_));
          if (p.skewh) {
            c.st(c.M(sM6_), m4);
            __syncwarp();
            axpy3(m4, 1.0, c.ldH(c.M(sM6_)));
          } else c.macc(m4, c.M(sM2_), c.M(sA2_));
          c.st(c.M(sM4_), m4);
          __syncwarp();
          if (qd >= 7) {
            c.macc(m6, c.M(sA4_), c.M(sM2_));
            c.macc(m6, c.M(sM4_), c.M(sA2_));
            c.st(c.M(sM6_), m6);
          }
          if (qd == 13) {
            c.st(c.M(sT1_), lin3(b[13], m6, b[11], m4, b[9], m2));   // Lw1
            c.st(c.M(sU_), lin3(b[12], m6, b[10], m4, b[8], m2));    // Lz1 (U is dead: R has been formed)
            __syncwarp();
            lw = K1SCtx::zero(); lv = K1SCtx::zero();
            c.macc(lw, c.M(sA6_), c.M(sT1_));
            c.macc(lw, c.M(sM6_), c.M(sW1_));
            c.macc(lv, c.M(sA6_), c.M(sU_));
            c.macc(lv, c.M(sM6_), c.M(sZ1_));
            axpy3(lw, 1.0, lin3(b[7], m6, b[5], m4, b[3], m2));
            axpy3(lv, 1.0, lin3(b[6], m6, b[4], m4, b[2], m2));
          } else if (qd == 7) {
            lw = lin3(b[7], m6, b[5], m4, b[3], m2);
            lv = lin3(b[6], m6, b[4], m4, b[2], m2);
          } else {
            lw = lin2(b[5], m4, b[3], m2);
            lv = lin2(b[4], m4, b[2], m2);
          }
          c.st(c.M(sLw_), lw);
          __syncwarp();
          c.macc(lu, c.M(sA_), c.M(sLw_));
          c.macc(lu, E, c.M(sW_));
          // rhs = (Lu + Lv) + (Lu - Lv) R ;  L = 2^-s N^-1 rhs
          C3 dd = lu, ss = lu;
          axpy3(dd, -1.0, lv);
          axpy3(ss, 1.0, lv);
          c.st(c.M(sM2_), dd);       // D (M2 is dead)
          __syncwarp();
          C3 rhs = ss;
          c.macc(rhs, c.M(sM2_), c.M(sR_));
          c.st(c.M(sM4_), rhs);
          __syncwarp();
          C3 L = K1SCtx::zero();
          c.macc(L, c.M(sN_), c.M(sM4_));
#pragma unroll
          for (int e = 0; e < 3; e++) { L.v[e].x *= scl; L.v[e].y *= scl; }
          if (sq == 0) k1s_store_slot(c, p.L + (slice * nc + j) * slot_d, S, L);
          else c.st(c.M(K1S_FIXED + j), L);
          __syncwarp();
        }
      }
      // ---- squarings: L <- R L + L R ; R <- R R ----
      for (int t2 = 0; t2 < sq; t2++) {
        if (p.want_jac && !taylor)
          for (int j = 0; j < nc; j++) {
            C3 ln = K1SCtx::zero();
            c.macc(ln, c.M(sR_), c.M(K1S_FIXED + j));
            c.macc(ln, c.M(K1S_FIXED + j), c.M(sR_));
            __syncwarp();   // every lane has read L_j
            if (t2 + 1 == sq) k1s_store_slot(c, p.L + (slice * nc + j) * slot_d, S, ln);
            else c.st(c.M(K1S_FIXED + j), ln);
            __syncwarp();
          }
        C3 r2 = K1SCtx::zero();
        c.macc(r2, c.M(sR_), c.M(sR_));
        __syncwarp();
        c.st(c.M(sR_), r2);
        rr3 = r2;
        __syncwarp();
      }
      k1s_store_slot(c, p.U + slice * slot_d, S, rr3);

      // ---- level-1 scan: Q <- U_k Q ----
      if (first_of_seg) q3 = rr3;
      else {
        C3 qn = K1SCtx::zero();
        c.macc(qn, c.M(sR_), c.M(sQ_));
        q3 = qn;
        __syncwarp();
      }
      c.st(c.M(sQ_), q3);
      __syncwarp();
      if (last_of_seg) k1s_store_slot(c, p.Q + (size_t)seg * slot_d, S, q3);
      {
        const int pi_q = qd == 13 ? 6 : qd == 7 ? 4 : 3;
        const int G = !p.want_jac ? 0 : taylor ? (p.order == 1 ? 0 : p.order == 2 ? 2 : p.order == 3 ? 5 : 10) : (2 * pi_q + 2 * sq + 2);
        my_thirds += 3 * (pi_q + sq) + 4 + 3 * nc * G;
        // executed: the skew-Hermitian shortcuts save two products per control (Frechet) / one or two (Taylor order 2 / >= 3)
        const int saved = (!p.want_jac || !p.skewh) ? 0 : taylor ? (p.order >= 3 ? 2 : p.order == 2 ? 1 : 0) : 2;
        my_exec += 3 * (pi_q + sq) + 4 + 3 * nc * (G - saved);
      }
    }
  }
  if (lane == 0) {
    if (!all_ok) atomicExch(p.status, 8);
    if (my_thirds != 0) {
      const double f = (8.0 * d * d * (double)d) * ((double)my_thirds / 3.0);
      atomicAdd(p.flops, f);
      atomicAdd(p.flops + 1, (8.0 * d * d * (double)d) * ((double)my_exec / 3.0));   // scalar DFMA products: nothing is padded
    }
  }
}

}  // namespace qoc

// qoc_basis.cuh -- the pulse-parameterisation chain rule around the GRAPE path, kept on the device so
// that only the ns x nc spline coefficients cross PCIe per evaluation:
//   u    = transpose(B * c)                 examples/ipopt_callbacks_exp.jl:13-14
//   dJdc = B' * transpose(dJdu)             examples/ipopt_callbacks_exp.jl:28
// B: Nt x ns column-major (basis functions sampled at the slice midpoints, examples/zz_coupling_ipopt_exp.jl:29-38),
// c: ns x nc (x batch) column-major, u / dJdu: nc x Nt (x batch) column-major.
#pragma once
#include <cuda_runtime.h>

namespace qoc {

__global__ void basis_expand_kernel(const double* __restrict__ B, const double* __restrict__ c, double* __restrict__ u, int nt,
                                    int ns, int nc, int batch) {
  const long long total = (long long)batch * nt * nc;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % nc);
    const long long bk = e / nc;
    const int k = (int)(bk % nt), b = (int)(bk / nt);
    const double* cj = c + ((size_t)b * nc + j) * ns;
    double s = 0.0;
    for (int q = 0; q < ns; q++) s = fma(B[k + (size_t)nt * q], cj[q], s);
    u[e] = s;
  }
}

// one CTA per (spline s, control j, pulse b)
__global__ void __launch_bounds__(256) basis_project_kernel(const double* __restrict__ B, const double* __restrict__ g,
                                                            double* __restrict__ dJdc, int nt, int ns, int nc) {
  __shared__ double red[8];
  const int s = blockIdx.x, j = blockIdx.y, b = blockIdx.z;
  double acc = 0.0;
  for (int k = threadIdx.x; k < nt; k += 256) acc = fma(B[k + (size_t)nt * s], g[((size_t)b * nt + k) * nc + j], acc);
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += red[w];
    dJdc[((size_t)b * nc + j) * ns + s] = t;
  }
}

}  // namespace qoc

// CSR sparse products Y = alpha*A@X + beta*Y for one or many right-hand sides.
// Replaces scipy's csr_matvec / csr_matvecs behind `B @ x`, `A @ x` in the reference
// (eigd/eigenvector_derivatives.py:255-265, 519, 609, 800-857, 975-1010, 1173-1252, 1500).
//
// k == 1 : a group of G lanes walks one row's non-zeros (coalesced index/value loads),
//          shuffle reduction inside the group.
// k  > 1 : a group of G >= k lanes owns one row; lane c accumulates column c, so each
//          matrix entry is loaded once (broadcast) and the k-wide slice of X is one
//          contiguous, coalesced read in the reference's (n, N) row-major layout.
// Algorithmic bytes: nnz*12 + (n+1)*4 + 2*n*k*8.
#include "common.cuh"
#include <cstdlib>
#include "../../include/eigd_b200.h"

namespace {

template <int G>
__global__ void __launch_bounds__(256)
spmv_kernel(int n, const int* __restrict__ indptr, const int* __restrict__ indices, const double* __restrict__ vals,
            const double* __restrict__ X, int64_t xrs, double* __restrict__ Y, int64_t yrs, double alpha, double beta) {
  const int lane = threadIdx.x & (G - 1);
  const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / G;
  // the loop bound is uniform across the warp (first group of the warp), so the shuffles
  // below always run with all 32 lanes converged
  const int gw = (threadIdx.x & 31) / G;  // group index inside the warp
  for (int64_t row = gid; row - gw < n; row += ngroups) {
    const bool valid = row < n;
    int p0 = 0, p1 = 0;
    if (valid) { p0 = indptr[row]; p1 = indptr[row + 1]; }
    double acc = 0.0;
    // three entries per lane and trip: all index / value loads first, then the three gathers, then the products
    // (same summation order as a plain loop; two memory round trips per trip instead of six)
    for (int p = p0 + lane; p < p1; p += 3 * G) {
      const bool b1 = p + G < p1, b2 = p + 2 * G < p1;
      const int i0 = indices[p];
      const int i1 = b1 ? indices[p + G] : 0;
      const int i2 = b2 ? indices[p + 2 * G] : 0;
      const double v0 = vals[p];
      const double v1 = b1 ? vals[p + G] : 0.0;
      const double v2 = b2 ? vals[p + 2 * G] : 0.0;
      const double x0 = X[(int64_t)i0 * xrs];
      const double x1 = b1 ? X[(int64_t)i1 * xrs] : 0.0;
      const double x2 = b2 ? X[(int64_t)i2 * xrs] : 0.0;
      acc = fma(v0, x0, acc);
      if (b1) acc = fma(v1, x1, acc);
      if (b2) acc = fma(v2, x2, acc);
    }
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, G);
    if (valid && lane == 0) {
      double y0 = (beta == 0.0) ? 0.0 : beta * Y[row * yrs];
      Y[row * yrs] = fma(alpha, acc, y0);
    }
  }
}

template <int G>
__global__ void __launch_bounds__(256)
spmm_kernel(int n, const int* __restrict__ indptr, const int* __restrict__ indices, const double* __restrict__ vals,
            const double* __restrict__ X, int64_t xrs, int64_t xcs, double* __restrict__ Y, int64_t yrs, int64_t ycs,
            int k, double alpha, double beta) {
  const int lane = threadIdx.x & (G - 1);
  const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / G;
  for (int64_t row = gid; row < n; row += ngroups) {
    int p0 = indptr[row], p1 = indptr[row + 1];
    double acc = 0.0;
    // four entries per trip: the (broadcast) index / value loads first, then the four gathers, then the products in
    // the order of the plain loop (one entry per trip, each with its own index -> value chain, is latency bound)
    if (lane < k) {
      const int64_t xoff = (int64_t)lane * xcs;
      int p = p0;
      for (; p + 4 <= p1; p += 4) {
        const int i0 = indices[p], i1 = indices[p + 1], i2 = indices[p + 2], i3 = indices[p + 3];
        const double v0 = vals[p], v1 = vals[p + 1], v2 = vals[p + 2], v3 = vals[p + 3];
        const double x0 = X[(int64_t)i0 * xrs + xoff], x1 = X[(int64_t)i1 * xrs + xoff];
        const double x2 = X[(int64_t)i2 * xrs + xoff], x3 = X[(int64_t)i3 * xrs + xoff];
        acc = fma(v0, x0, acc);
        acc = fma(v1, x1, acc);
        acc = fma(v2, x2, acc);
        acc = fma(v3, x3, acc);
      }
      for (; p < p1; ++p) acc = fma(vals[p], X[(int64_t)indices[p] * xrs + xoff], acc);
    }
    if (lane < k) {
      int64_t yi = row * yrs + (int64_t)lane * ycs;
      double y0 = (beta == 0.0) ? 0.0 : beta * Y[yi];
      Y[yi] = fma(alpha, acc, y0);
    }
  }
}

// y = A x and, in the same launch, dot = sum_i y_i x_i (the B-norm of a new Lanczos direction: one launch instead of
// an SpMV, a one-row dot kernel and its reduction).  One row per group of G lanes, no row loop; block sums go to
// partial[cta], the last CTA to finish (ticket) adds them in a fixed order -> bitwise reproducible.
__device__ unsigned int g_spmv_ticket = 0;

template <int G>
__global__ void __launch_bounds__(256)
spmv_dot_kernel(int n, const int* __restrict__ indptr, const int* __restrict__ indices, const double* __restrict__ vals,
                const double* __restrict__ X, double* __restrict__ Y, double* __restrict__ partial,
                double* __restrict__ out, unsigned int* __restrict__ ticket) {
  constexpr int GPB = 256 / G;                              // groups (rows) per CTA
  __shared__ double red[256];
  __shared__ bool last;
  const int tid = threadIdx.x, lane = tid & (G - 1), grp = tid / G;
  const int64_t row = (int64_t)blockIdx.x * GPB + grp;
  const bool valid = row < n;
  int p0 = 0, p1 = 0;
  if (valid) { p0 = indptr[row]; p1 = indptr[row + 1]; }
  double acc = 0.0;
  for (int p = p0 + lane; p < p1; p += 3 * G) {             // same batching and summation order as spmv_kernel
    const bool b1 = p + G < p1, b2 = p + 2 * G < p1;
    const int i0 = indices[p];
    const int i1 = b1 ? indices[p + G] : 0;
    const int i2 = b2 ? indices[p + 2 * G] : 0;
    const double v0 = vals[p];
    const double v1 = b1 ? vals[p + G] : 0.0;
    const double v2 = b2 ? vals[p + 2 * G] : 0.0;
    const double x0 = X[i0];
    const double x1 = b1 ? X[i1] : 0.0;
    const double x2 = b2 ? X[i2] : 0.0;
    acc = fma(v0, x0, acc);
    if (b1) acc = fma(v1, x1, acc);
    if (b2) acc = fma(v2, x2, acc);
  }
#pragma unroll
  for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, G);
  double t = 0.0;
  if (valid && lane == 0) {
    Y[row] = acc;
    t = acc * X[row];
  }
  red[tid] = t;                                             // zero for every lane but the group leaders
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  if (tid == 0) partial[blockIdx.x] = red[0];
  __threadfence();
  __syncthreads();
  if (tid == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  double s = 0.0;
  for (int c = tid; c < (int)gridDim.x; c += 256) s += __ldcg(partial + c);
  red[tid] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  if (tid == 0) { out[0] = red[0]; *ticket = 0u; }
}

inline int grid_for(int64_t groups_needed, int groups_per_block) {
  int64_t g = (groups_needed + groups_per_block - 1) / groups_per_block;
  int64_t cap = 148 * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

// internal (krylov.cu): y = A x, out[0] = y . x; work must hold ceil(n / 64) doubles
int eigd_csr_spmv_dot(int n, const int* indptr, const int* indices, const double* vals, const double* x, double* y,
                      double* out, double* work) {
  if (n <= 0) return 0;
  unsigned int* ticket = nullptr;
  EIGD_CUDA(cudaGetSymbolAddress((void**)&ticket, g_spmv_ticket));
  const int grid = (n + 63) / 64;
  EIGD_LAUNCH(spmv_dot_kernel<4>, grid, 256, 0, n, indptr, indices, vals, x, y, work, out, ticket);
  EIGD_CHECK_LAUNCH();
  return 0;
}

extern "C" int eigd_csr_spmm(int n, const int* indptr, const int* indices, const double* vals, const double* X,
                             int64_t xrs, int64_t xcs, double* Y, int64_t yrs, int64_t ycs, int k, double alpha,
                             double beta) {
  if (n <= 0 || k <= 0) return 0;
  if (k == 1) {
    static int G = -1;                      // lanes per row; EIGD_SPMV_G overrides (developer tuning)
    if (G < 0) { const char* e = getenv("EIGD_SPMV_G"); G = e ? atoi(e) : 4; }   // measured (cold L2, 251k rows): 9 nnz/row 17 us (8 lanes: 22 us), 18 nnz/row 22.5 us (27 us)
    switch (G) {
      case 1: EIGD_LAUNCH(spmv_kernel<1>, grid_for(n, 256), 256, 0, n, indptr, indices, vals, X, xrs, Y, yrs, alpha, beta); break;
      case 2: EIGD_LAUNCH(spmv_kernel<2>, grid_for(n, 128), 256, 0, n, indptr, indices, vals, X, xrs, Y, yrs, alpha, beta); break;
      case 4: EIGD_LAUNCH(spmv_kernel<4>, grid_for(n, 64), 256, 0, n, indptr, indices, vals, X, xrs, Y, yrs, alpha, beta); break;
      case 16: EIGD_LAUNCH(spmv_kernel<16>, grid_for(n, 16), 256, 0, n, indptr, indices, vals, X, xrs, Y, yrs, alpha, beta); break;
      default: EIGD_LAUNCH(spmv_kernel<8>, grid_for(n, 32), 256, 0, n, indptr, indices, vals, X, xrs, Y, yrs, alpha, beta); break;
    }
    EIGD_CHECK_LAUNCH();
    return 0;
  }
  for (int c0 = 0; c0 < k; c0 += 32) {
    int kc = k - c0 < 32 ? k - c0 : 32;
    const double* Xc = X + (int64_t)c0 * xcs;
    double* Yc = Y + (int64_t)c0 * ycs;
    if (kc <= 2) EIGD_LAUNCH(spmm_kernel<2>, grid_for(n, 128), 256, 0, n, indptr, indices, vals, Xc, xrs, xcs, Yc, yrs, ycs, kc, alpha, beta);
    else if (kc <= 4) EIGD_LAUNCH(spmm_kernel<4>, grid_for(n, 64), 256, 0, n, indptr, indices, vals, Xc, xrs, xcs, Yc, yrs, ycs, kc, alpha, beta);
    else if (kc <= 8) EIGD_LAUNCH(spmm_kernel<8>, grid_for(n, 32), 256, 0, n, indptr, indices, vals, Xc, xrs, xcs, Yc, yrs, ycs, kc, alpha, beta);
    else if (kc <= 16) EIGD_LAUNCH(spmm_kernel<16>, grid_for(n, 16), 256, 0, n, indptr, indices, vals, Xc, xrs, xcs, Yc, yrs, ycs, kc, alpha, beta);
    else EIGD_LAUNCH(spmm_kernel<32>, grid_for(n, 8), 256, 0, n, indptr, indices, vals, Xc, xrs, xcs, Yc, yrs, ycs, kc, alpha, beta);
    EIGD_CHECK_LAUNCH();
  }
  return 0;
}

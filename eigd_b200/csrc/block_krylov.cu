// Device-resident BLOCK shift-and-invert Lanczos recurrence (block size P = 2 .. 4).
//
// Why a block method: on B200 a sparse triangular solve is bound by the dependency depth of the assembly tree,
// not by bandwidth, so P right-hand sides cost little more than one -- and the tall-skinny
// orthogonalisation products read the Krylov basis once for all P new vectors.  The block recurrence needs about
// the same number of operator applications as the single-vector one (measured on the bench problems), i.e. 1/P of
// the sequential solves.  Replaces, like krylov.cu, the reverse-communication loop of reference eigd/arpack.py:438-442
// (ARPACK dsaupd is a single-vector method; the (d, z, Tm, v) contract of eigd/arpack.py:58-101 -- v^T B v = I,
// Tm = v^T B OP v, z = v y -- holds for the block basis as well; Tm is block tridiagonal plus the restart arrow).
//
// One call runs the block steps of a restart cycle without returning to the host.  Per step (m = basis vectors so far,
// j = first vector of the block being expanded):
//     W   = factor^{-1} BV[j : j+P]                      (one P-column solve, + refinement steps)
//     H1  = BV[lo:m] W, W -= V[lo:m]^T H1                (local Gram-Schmidt pass in the B inner product: lo = j - P, the two
//                                                         blocks of the three-term recurrence; lo = 0 after a restart)
//     H2  = BV[0:m] W,  W -= V[0:m]^T H2                 (unconditional full pass: full reorthogonalisation)
//     BW  = B W,  G = W^T BW = R1^T R1,  W <- W R1^-1, BW <- BW R1^-1      (Cholesky QR ...
//     G2  = W^T BW = R2^T R2,  W <- W R2^-1, BW <- BW R2^-1                 ... twice: orthonormal to rounding)
//     V[m : m+P] = W, BV[m : m+P] = BW;  A_j = (H1 + H2)[j : j+P] (diagonal block), R_j = R2 R1 (sub-diagonal block)
// The small P x P factorisations run in one warp on the device (no host round trip inside a cycle).
//
// Storage as in krylov.cu: V and BV are (ncv + P) x n row-major (one vector per row, leading dimension ld).
#include "common.cuh"
#include "../../include/eigd_b200.h"

#include <cstdlib>

namespace {

constexpr int BK_THREADS = 256;
constexpr int BK_EPT = 4;                        // elements per thread and chunk
constexpr int BK_CHUNK = BK_THREADS * BK_EPT;
constexpr int BK_ROWS = 4;                       // basis rows in flight per thread
constexpr int BK_JMAX = 128;                     // basis vectors per launch
constexpr int BK_PMAX = 4;

__device__ unsigned int g_block_ticket = 0;

// out[a * P + c] = sum_i Vb[a][i] * W[c][i]  (a < j rows of the basis, c < P new vectors, W rows at W + c * ldw).
// CTA-strided chunks of 1024 elements: the P vectors' entries live in registers, BK_ROWS basis rows are streamed at a
// time; per-warp partial sums accumulate in shared memory, per-CTA partials go to `partial`, the last CTA to finish
// (ticket) adds them in a fixed order -- bitwise reproducible, one launch.
template <int P>
__device__ void block_chol(const double* __restrict__ G, double* __restrict__ Rinv, double* __restrict__ Racc, int pass);

// chol_pass >= 0 (Gram-matrix call, j == P): the last CTA also factors the P x P result (block_chol) -- one launch less
// per Cholesky-QR pass.
template <int P>
__global__ void __launch_bounds__(BK_THREADS)
block_dots_kernel(int64_t n, int j, const double* __restrict__ Vb, int64_t ldv, const double* __restrict__ W, int64_t ldw,
                  double* __restrict__ partial, double* __restrict__ out, unsigned int* __restrict__ ticket, int chol_pass,
                  double* __restrict__ Rinv, double* __restrict__ Racc) {
  extern __shared__ double red[];                // [BK_THREADS / 32][j * P]
  __shared__ bool last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int JP = j * P;
  for (int e = tid; e < (BK_THREADS / 32) * JP; e += BK_THREADS) red[e] = 0.0;
  __syncthreads();
  double* myred = red + warp * JP;
  for (int64_t c0 = (int64_t)blockIdx.x * BK_CHUNK; c0 < n; c0 += (int64_t)gridDim.x * BK_CHUNK) {
    const int64_t i0 = c0 + tid;
    double wv[P][BK_EPT];
#pragma unroll
    for (int c = 0; c < P; ++c)
#pragma unroll
      for (int q = 0; q < BK_EPT; ++q) {
        const int64_t i = i0 + (int64_t)q * BK_THREADS;
        wv[c][q] = (i < n) ? W[(int64_t)c * ldw + i] : 0.0;
      }
    for (int a0 = 0; a0 < j; a0 += BK_ROWS) {
      double v[BK_ROWS][BK_EPT];
#pragma unroll
      for (int r = 0; r < BK_ROWS; ++r) {
        const double* row = Vb + (int64_t)min(a0 + r, j - 1) * ldv;
#pragma unroll
        for (int q = 0; q < BK_EPT; ++q) {
          const int64_t i = i0 + (int64_t)q * BK_THREADS;
          v[r][q] = (i < n) ? __ldg(row + i) : 0.0;
        }
      }
#pragma unroll
      for (int r = 0; r < BK_ROWS; ++r)
#pragma unroll
        for (int c = 0; c < P; ++c) {
          double s = 0.0;
#pragma unroll
          for (int q = 0; q < BK_EPT; ++q) s = fma(v[r][q], wv[c][q], s);
          s = warp_sum(s);
          if (lane == 0 && a0 + r < j) myred[(a0 + r) * P + c] += s;
        }
    }
  }
  __syncthreads();
  for (int e = tid; e < JP; e += BK_THREADS) {
    double s = 0.0;
#pragma unroll
    for (int g = 0; g < BK_THREADS / 32; ++g) s += red[g * JP + e];
    partial[(int64_t)blockIdx.x * JP + e] = s;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  const int ncta = gridDim.x;
  for (int e = warp; e < JP; e += BK_THREADS / 32) {      // one warp per output, lanes stride over the CTAs
    double s = 0.0;
    for (int c = lane; c < ncta; c += 32) s += __ldcg(partial + (int64_t)c * JP + e);
    s = warp_sum(s);
    if (lane == 0) out[e] = s;
  }
  if (tid == 0) *ticket = 0u;
  if (chol_pass >= 0) {
    __syncthreads();
    if (tid == 0) block_chol<P>(out, Rinv, Racc, chol_pass);
  }
}

// W[c][i] += sign * sum_a H[a * P + c] * Vb[a][i]
// H1 / Aout / jblk (second Gram-Schmidt pass only, else NULL): CTA 0 also writes the diagonal block of the projected
// operator, Aout = (H1 + H)[jblk : jblk + P].
template <int P>
__global__ void __launch_bounds__(BK_THREADS)
block_axpy_kernel(int64_t n, int j, const double* __restrict__ Vb, int64_t ldv, const double* __restrict__ H, double sign,
                  double* __restrict__ W, int64_t ldw, const double* __restrict__ H1, double* __restrict__ Aout, int jblk) {
  extern __shared__ double hs[];                 // [j * P]
  const int tid = threadIdx.x;
  if (Aout && blockIdx.x == 0 && tid < P * P) Aout[tid] = H1[jblk * P + tid] + H[jblk * P + tid];
  for (int e = tid; e < j * P; e += BK_THREADS) hs[e] = sign * H[e];
  __syncthreads();
  for (int64_t c0 = (int64_t)blockIdx.x * BK_CHUNK; c0 < n; c0 += (int64_t)gridDim.x * BK_CHUNK) {
    const int64_t i0 = c0 + tid;
    double acc[P][BK_EPT];
#pragma unroll
    for (int c = 0; c < P; ++c)
#pragma unroll
      for (int q = 0; q < BK_EPT; ++q) {
        const int64_t i = i0 + (int64_t)q * BK_THREADS;
        acc[c][q] = (i < n) ? W[(int64_t)c * ldw + i] : 0.0;
      }
    for (int a0 = 0; a0 < j; a0 += BK_ROWS) {
      double v[BK_ROWS][BK_EPT];
#pragma unroll
      for (int r = 0; r < BK_ROWS; ++r) {
        const double* row = Vb + (int64_t)min(a0 + r, j - 1) * ldv;
#pragma unroll
        for (int q = 0; q < BK_EPT; ++q) {
          const int64_t i = i0 + (int64_t)q * BK_THREADS;
          v[r][q] = (i < n) ? __ldg(row + i) : 0.0;
        }
      }
#pragma unroll
      for (int r = 0; r < BK_ROWS; ++r) {
        if (a0 + r >= j) break;
#pragma unroll
        for (int c = 0; c < P; ++c) {
          const double h = hs[(a0 + r) * P + c];
#pragma unroll
          for (int q = 0; q < BK_EPT; ++q) acc[c][q] = fma(h, v[r][q], acc[c][q]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < P; ++c)
#pragma unroll
      for (int q = 0; q < BK_EPT; ++q) {
        const int64_t i = i0 + (int64_t)q * BK_THREADS;
        if (i < n) W[(int64_t)c * ldw + i] = acc[c][q];
      }
  }
}

// Cholesky G = R^T R of the P x P Gram matrix (row-major G[a * P + c]); writes the inverse Rinv (upper triangular,
// row-major) for the scaling kernel and accumulates Racc <- R * Racc (Racc = identity on the first pass: pass 0).
// A non-positive pivot (rank-deficient block: breakdown) is flagged by NaNs, which the host checks once per cycle.
template <int P>
__device__ void block_chol(const double* __restrict__ G, double* __restrict__ Rinv, double* __restrict__ Racc, int pass) {
  double R[P][P], X[P][P];
#pragma unroll
  for (int a = 0; a < P; ++a)
#pragma unroll
    for (int c = 0; c < P; ++c) { R[a][c] = 0.0; X[a][c] = 0.0; }
  bool ok = true;
#pragma unroll
  for (int c = 0; c < P; ++c) {                  // column-by-column (upper factor): R[a][c], a <= c
#pragma unroll
    for (int a = 0; a <= c; ++a) {
      double s = 0.5 * (G[a * P + c] + G[c * P + a]);
#pragma unroll
      for (int t = 0; t < a; ++t) s -= R[t][a] * R[t][c];
      if (a < c) R[a][c] = s / R[a][a];
      else {
        if (!(s > 0.0)) ok = false;
        R[c][c] = sqrt(s);
      }
    }
  }
  // X = R^-1 (upper triangular) by back substitution
#pragma unroll
  for (int c = 0; c < P; ++c) {
    X[c][c] = 1.0 / R[c][c];
#pragma unroll
    for (int a = c - 1; a >= 0; --a) {
      double s = 0.0;
#pragma unroll
      for (int t = a + 1; t <= c; ++t) s += R[a][t] * X[t][c];
      X[a][c] = -s / R[a][a];
    }
  }
  const double bad = ok ? 0.0 : __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
  for (int a = 0; a < P; ++a)
#pragma unroll
    for (int c = 0; c < P; ++c) Rinv[a * P + c] = X[a][c] + bad;
  double A[P][P];
#pragma unroll
  for (int a = 0; a < P; ++a)
#pragma unroll
    for (int c = 0; c < P; ++c) {
      if (pass == 0) A[a][c] = R[a][c];
      else {
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < P; ++t) s += R[a][t] * Racc[t * P + c];
        A[a][c] = s;
      }
    }
#pragma unroll
  for (int a = 0; a < P; ++a)
#pragma unroll
    for (int c = 0; c < P; ++c) Racc[a * P + c] = A[a][c] + bad;
}

// W <- W Rinv, BW <- BW Rinv (Rinv upper triangular, row-major); vectors are rows (W + c * ld)
template <int P>
__global__ void __launch_bounds__(256)
block_scale_kernel(int64_t n, const double* __restrict__ Rinv, double* __restrict__ W, double* __restrict__ BW, int64_t ld) {
  __shared__ double x[P * P];
  if (threadIdx.x < P * P) x[threadIdx.x] = Rinv[threadIdx.x];
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double w[P], b[P];
#pragma unroll
    for (int c = 0; c < P; ++c) { w[c] = W[(int64_t)c * ld + i]; b[c] = BW[(int64_t)c * ld + i]; }
#pragma unroll
    for (int c = P - 1; c >= 0; --c) {
      double sw = 0.0, sb = 0.0;
#pragma unroll
      for (int t = 0; t <= c; ++t) { sw = fma(w[t], x[t * P + c], sw); sb = fma(b[t], x[t * P + c], sb); }
      W[(int64_t)c * ld + i] = sw;
      BW[(int64_t)c * ld + i] = sb;
    }
  }
}

// r = b - r for P strided vectors; x += dx
__global__ void block_sub_from_kernel(int64_t n, int P, const double* __restrict__ b, int64_t ldb, double* __restrict__ r, int64_t ldr) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * P; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e / n);
    const int64_t i = e - (int64_t)c * n;
    r[(int64_t)c * ldr + i] = b[(int64_t)c * ldb + i] - r[(int64_t)c * ldr + i];
  }
}
__global__ void block_add_to_kernel(int64_t n, int P, const double* __restrict__ dx, int64_t ldd, double* __restrict__ x, int64_t ldx) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * P; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e / n);
    const int64_t i = e - (int64_t)c * n;
    x[(int64_t)c * ldx + i] += dx[(int64_t)c * ldd + i];
  }
}

inline int ew_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

template <int P>
int dots(int64_t n, int j, const double* Vb, int64_t ldv, const double* W, int64_t ldw, double* out, double* work,
         int chol_pass = -1, double* Rinv = nullptr, double* Racc = nullptr) {
  unsigned int* ticket = nullptr;
  EIGD_CUDA(cudaGetSymbolAddress((void**)&ticket, g_block_ticket));
  int64_t want = (n + BK_CHUNK - 1) / BK_CHUNK;
  const int64_t cap = eigd_gemm_tn_workspace(32, 32) / ((int64_t)j * P);     // partial[grid][j * P] must fit the workspace
  int grid = (int)(want < cap ? want : cap);
  if (grid < 1) { eigd_set_error("block_dots: workspace too small"); return 8; }
  const size_t smem = (size_t)(BK_THREADS / 32) * j * P * sizeof(double);
  EIGD_LAUNCH(block_dots_kernel<P>, grid, BK_THREADS, smem, n, j, Vb, ldv, W, ldw, work, out, ticket, chol_pass, Rinv, Racc);
  EIGD_CHECK_LAUNCH();
  return 0;
}

template <int P>
int axpy(int64_t n, int j, const double* Vb, int64_t ldv, const double* H, double sign, double* W, int64_t ldw,
         const double* H1 = nullptr, double* Aout = nullptr, int jblk = 0) {
  int64_t want = (n + BK_CHUNK - 1) / BK_CHUNK;
  int grid = (int)(want < 148 * 16 ? want : 148 * 16);
  EIGD_LAUNCH(block_axpy_kernel<P>, grid, BK_THREADS, (size_t)j * P * sizeof(double), n, j, Vb, ldv, H, sign, W, ldw, H1, Aout, jblk);
  EIGD_CHECK_LAUNCH();
  return 0;
}

// Cholesky QR (twice) of the P vectors W in the B inner product; BW = B W on entry and on exit; Racc <- R2 R1.
// scratch: 2 * P * P doubles (Gram matrix, inverse factor)
template <int P>
int cholqr2(int64_t n, double* W, double* BW, int64_t ld, double* Racc, double* scratch, double* work) {
  double* G = scratch;
  double* Rinv = scratch + P * P;
  int rc;
  for (int pass = 0; pass < 2; ++pass) {
    if ((rc = dots<P>(n, P, BW, ld, W, ld, G, work, pass, Rinv, Racc))) return rc;   // G = BW^T W, factored by the last CTA
    EIGD_LAUNCH(block_scale_kernel<P>, ew_grid(n), 256, 0, n, Rinv, W, BW, ld);
    EIGD_CHECK_LAUNCH();
  }
  return 0;
}

template <int P>
int extend(eigd_factor* f, int refine, int n, const int* mp, const int* mi, const double* mv, const int* bp, const int* bi,
           const double* bv, double* V, double* BV, int64_t ld, int j0, int m0, int ncv, double* Ablk, double* Rblk,
           double* H1, double* H2, double* scratch, double* work, double* work2) {
  int rc;
  int step = 0;
  static int local_first = -1;          // EIGD_LANCZOS_LOCAL=0: two full passes (developer comparison runs)
  if (local_first < 0) { const char* e = getenv("EIGD_LANCZOS_LOCAL"); local_first = e ? atoi(e) : 1; }
  for (int j = j0, m = m0; j < ncv; j += P, m += P, ++step) {
    const double* bvj = BV + (int64_t)j * ld;
    double* W = V + (int64_t)m * ld;
    double* BW = BV + (int64_t)m * ld;
    // W = OP V_j = factor^{-1} (B V_j): the P vectors are rows (stride ld between vectors, 1 inside)
    if ((rc = eigd_factor_solve(f, bvj, 1, ld, W, 1, ld, P))) return rc;
    for (int it = 0; it < refine; ++it) {                        // X += F (B V_j - mat X)
      double* r = work2;
      double* dx = work2 + (int64_t)P * n;
      if ((rc = eigd_csr_spmm(n, mp, mi, mv, W, 1, ld, r, 1, n, P, 1.0, 0.0))) return rc;
      EIGD_LAUNCH(block_sub_from_kernel, ew_grid((int64_t)n * P), 256, 0, (int64_t)n, P, bvj, ld, r, (int64_t)n);
      EIGD_CHECK_LAUNCH();
      if ((rc = eigd_factor_solve(f, r, 1, n, dx, 1, n, P))) return rc;
      EIGD_LAUNCH(block_add_to_kernel, ew_grid((int64_t)n * P), 256, 0, (int64_t)n, P, dx, (int64_t)n, W, ld);
      EIGD_CHECK_LAUNCH();
    }
    // Gram-Schmidt in the B inner product: a LOCAL pass against the vectors W is coupled to in exact arithmetic -- the
    // block being expanded and its predecessor (block three-term recurrence), or the whole basis in the first step
    // after a thick restart (the locked Ritz vectors: the arrow) -- then one FULL classical pass against V[0:m].  The
    // local pass takes out the large components (where the cancellation is), the full pass the O(eps)-relative ones
    // left along the older vectors: full reorthogonalisation at half the basis traffic of two full passes.
    const int lo = (step == 0 || local_first == 0) ? 0 : (j - P > 0 ? j - P : 0);
    if ((rc = dots<P>(n, m - lo, BV + (int64_t)lo * ld, ld, W, ld, H1, work))) return rc;
    if ((rc = axpy<P>(n, m - lo, V + (int64_t)lo * ld, ld, H1, -1.0, W, ld))) return rc;
    if ((rc = dots<P>(n, m, BV, ld, W, ld, H2, work))) return rc;
    // diagonal block A_j = (H1 + H2)[j : j + P]; H1 holds rows lo .. m-1
    if ((rc = axpy<P>(n, m, V, ld, H2, -1.0, W, ld, H1 - (int64_t)lo * P, Ablk + (int64_t)step * P * P, j))) return rc;
    // BW = B W, then orthonormalise the block
    if ((rc = eigd_csr_spmm(n, bp, bi, bv, W, 1, ld, BW, 1, ld, P, 1.0, 0.0))) return rc;
    if ((rc = cholqr2<P>(n, W, BW, ld, Rblk + (int64_t)step * P * P, scratch, work))) return rc;
  }
  return 0;
}

template <int P>
int start_block(int n, const int* bp, const int* bi, const double* bv, double* V, double* BV, int64_t ld, double* scratch,
                double* work) {
  int rc;
  if ((rc = eigd_csr_spmm(n, bp, bi, bv, V, 1, ld, BV, 1, ld, P, 1.0, 0.0))) return rc;
  return cholqr2<P>(n, V, BV, ld, scratch + 2 * P * P, scratch, work);
}

}  // namespace

extern "C" int eigd_block_lanczos_extend(eigd_factor* f, int refine, int n, int P, const int* d_mat_indptr,
                                         const int* d_mat_indices, const double* d_mat_vals, const int* d_b_indptr,
                                         const int* d_b_indices, const double* d_b_vals, double* d_V, double* d_BV, int64_t ld,
                                         int j0, int m0, int ncv, double* d_Ablk, double* d_Rblk, double* d_H1, double* d_H2,
                                         double* d_scratch, double* d_work, double* d_work2) {
  if (j0 >= ncv) return 0;
  if ((ncv - j0) % P != 0 || m0 < j0 + P) { eigd_set_error("block_lanczos_extend: the cycle does not end on a block boundary"); return 7; }
  if (m0 + (ncv - j0) > BK_JMAX + BK_PMAX) { eigd_set_error("block_lanczos_extend: basis too large"); return 7; }
  if (refine > 0 && (!d_mat_indptr || !d_work2)) { eigd_set_error("block_lanczos_extend: refinement needs the shifted matrix and 2 P n doubles of work"); return 7; }
  switch (P) {
    case 2: return extend<2>(f, refine, n, d_mat_indptr, d_mat_indices, d_mat_vals, d_b_indptr, d_b_indices, d_b_vals, d_V, d_BV, ld, j0, m0, ncv, d_Ablk, d_Rblk, d_H1, d_H2, d_scratch, d_work, d_work2);
    case 3: return extend<3>(f, refine, n, d_mat_indptr, d_mat_indices, d_mat_vals, d_b_indptr, d_b_indices, d_b_vals, d_V, d_BV, ld, j0, m0, ncv, d_Ablk, d_Rblk, d_H1, d_H2, d_scratch, d_work, d_work2);
    case 4: return extend<4>(f, refine, n, d_mat_indptr, d_mat_indices, d_mat_vals, d_b_indptr, d_b_indices, d_b_vals, d_V, d_BV, ld, j0, m0, ncv, d_Ablk, d_Rblk, d_H1, d_H2, d_scratch, d_work, d_work2);
    default: eigd_set_error("block_lanczos_extend: block size must be 2, 3 or 4"); return 7;
  }
}

// B-orthonormalise the P start vectors V[0:P] (rows) and set BV[0:P] = B V[0:P]; scratch: 3 * P * P doubles
extern "C" int eigd_block_lanczos_start(int n, int P, const int* d_b_indptr, const int* d_b_indices, const double* d_b_vals,
                                        double* d_V, double* d_BV, int64_t ld, double* d_scratch, double* d_work) {
  switch (P) {
    case 2: return start_block<2>(n, d_b_indptr, d_b_indices, d_b_vals, d_V, d_BV, ld, d_scratch, d_work);
    case 3: return start_block<3>(n, d_b_indptr, d_b_indices, d_b_vals, d_V, d_BV, ld, d_scratch, d_work);
    case 4: return start_block<4>(n, d_b_indptr, d_b_indices, d_b_vals, d_V, d_BV, ld, d_scratch, d_work);
    default: eigd_set_error("block_lanczos_start: block size must be 2, 3 or 4"); return 7;
  }
}

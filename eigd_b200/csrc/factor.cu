// Supernodal multifrontal LDL^T of the shifted matrix and its multi-RHS triangular solves.
// Replaces scipy.sparse.linalg.splu (SuperLU gstrf) and SuperLU.solve behind SpLuOperator
// (reference eigd/eigenvector_derivatives.py:11-23): x <- (A - sigma B)^{-1} x, or
// (B + sigma A)^{-1} x in buckling mode (:786-790).
//
// Data layout in HBM
//   fronts : one dense column-major f x f block per supernode (lower triangle used);
//            columns [0, ncols) hold L (unit diagonal implied, D on the diagonal), the
//            trailing (f-ncols)^2 block is the contribution passed to the parent.
//   linv   : explicit inverses of the 32x32 unit-lower diagonal blocks of every L11, so the
//            triangular solves inside a front become small dense products.
//   wbuf   : per-front work vectors (f x k, row-major) used by the solve; the part below the
//            pivot rows carries the front's update to its parent (multifrontal solve: no
//            atomics, deterministic sums).
// Execution: supernodes are grouped by height in the assembly tree; every level is a handful
// of batched launches driven by host-built task lists (front, tile).
#include "common.cuh"
#include "symbolic.hpp"
#include "../../include/eigd_b200.h"

#include <algorithm>
#include <cmath>
#include <vector>

namespace {

constexpr int NB = 32;        // pivot block width
constexpr int TRSM_ROWS = 128;
constexpr int UPD_TILE = 64;
constexpr int EA_TILE = 32;
constexpr int FWD_ROWS = 128;
constexpr int BWD_COLS = 8;

struct SymDev {
  int n, nsuper;
  int *perm, *iperm, *col2sn;
  int* sn_first;
  int64_t* sn_rowptr;
  int* sn_rows;
  int* rel;
  int64_t* front_off;
  int64_t* w_off;
  int64_t* linv_off;
  int* sn_parent;
  int *child_ptr, *child_idx;
};

struct Launch {
  int kind;      // 0 extend-add, 1 diag, 2 trsm, 3 update | solve: 4 fwd_diag, 5 fwd_update, 6 bwd_update, 7 bwd_diag
  int kb;        // pivot block index (factor kinds)
  int64_t off;   // offset into the task array (int2 entries)
  int count;
};

struct SymDevHolder {
  SymDev d;
  std::vector<void*> allocs;
  int2* tasks = nullptr;
  std::vector<Launch> factor_plan, fwd_plan, bwd_plan;
  int64_t linv_total = 0;
};

template <class T>
int upload(SymDevHolder* h, const std::vector<T>& v, T** out) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  EIGD_CUDA(cudaMalloc(&p, bytes));
  h->allocs.push_back(p);
  if (!v.empty()) EIGD_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = (T*)p;
  return 0;
}

void free_symdev(void* p) {
  auto* h = (SymDevHolder*)p;
  for (void* a : h->allocs) cudaFree(a);
  if (h->tasks) cudaFree(h->tasks);
  delete h;
}

inline int tri_tiles(int m, int tile) {
  int t = (m + tile - 1) / tile;
  return t * (t + 1) / 2;
}

int build_symdev(eigd_symbolic* S) {
  if (S->dev) return 0;
  auto* h = new SymDevHolder();
  S->dev = h;
  S->dev_free = free_symdev;
  int ns = S->nsuper;
  h->d.n = S->n;
  h->d.nsuper = ns;
  std::vector<int64_t> linv_off(ns + 1, 0);
  for (int k = 0; k < ns; ++k) linv_off[k + 1] = linv_off[k] + (int64_t)((sn_ncols(S, k) + NB - 1) / NB) * NB * NB;
  h->linv_total = linv_off[ns];
  int rc = 0;
  rc |= upload(h, S->perm, &h->d.perm);
  rc |= upload(h, S->iperm, &h->d.iperm);
  rc |= upload(h, S->col2sn, &h->d.col2sn);
  rc |= upload(h, S->sn_first, &h->d.sn_first);
  rc |= upload(h, S->sn_rowptr, &h->d.sn_rowptr);
  rc |= upload(h, S->sn_rows, &h->d.sn_rows);
  rc |= upload(h, S->rel, &h->d.rel);
  rc |= upload(h, S->front_off, &h->d.front_off);
  rc |= upload(h, S->w_off, &h->d.w_off);
  rc |= upload(h, linv_off, &h->d.linv_off);
  rc |= upload(h, S->sn_parent, &h->d.sn_parent);
  rc |= upload(h, S->child_ptr, &h->d.child_ptr);
  rc |= upload(h, S->child_idx, &h->d.child_idx);
  if (rc) return rc;

  // ---- task lists ---------------------------------------------------------------------
  std::vector<int2> tasks;
  auto push_launch = [&](std::vector<Launch>& plan, int kind, int kb, int64_t off) {
    int count = (int)((int64_t)tasks.size() - off);
    if (count > 0) plan.push_back({kind, kb, off, count});
  };
  // rank of each child among its siblings
  std::vector<int> rank(ns, 0);
  for (int p = 0; p < ns; ++p)
    for (int q = S->child_ptr[p]; q < S->child_ptr[p + 1]; ++q) rank[S->child_idx[q]] = q - S->child_ptr[p];
  for (int l = 0; l < S->nlevels; ++l) {
    const int* sn = S->level_sn.data() + S->level_ptr[l];
    int cnt = S->level_ptr[l + 1] - S->level_ptr[l];
    // extend-add: one launch per sibling rank so that no two tasks touch the same parent
    int maxrank = 0;
    for (int i = 0; i < cnt; ++i) maxrank = std::max(maxrank, S->child_ptr[sn[i] + 1] - S->child_ptr[sn[i]]);
    for (int r = 0; r < maxrank; ++r) {
      int64_t off = (int64_t)tasks.size();
      for (int i = 0; i < cnt; ++i) {
        int p = sn[i];
        if (S->child_ptr[p] + r >= S->child_ptr[p + 1]) continue;
        int c = S->child_idx[S->child_ptr[p] + r];
        int nt = tri_tiles(sn_nbelow(S, c), EA_TILE);
        for (int t = 0; t < nt; ++t) tasks.push_back(make_int2(c, t));
      }
      push_launch(h->factor_plan, 0, r, off);
    }
    int maxsteps = 0;
    for (int i = 0; i < cnt; ++i) maxsteps = std::max(maxsteps, (sn_ncols(S, sn[i]) + NB - 1) / NB);
    for (int kb = 0; kb < maxsteps; ++kb) {
      int64_t off = (int64_t)tasks.size();
      for (int i = 0; i < cnt; ++i)
        if (sn_ncols(S, sn[i]) > kb * NB) tasks.push_back(make_int2(sn[i], 0));
      push_launch(h->factor_plan, 1, kb, off);
      off = (int64_t)tasks.size();
      for (int i = 0; i < cnt; ++i) {
        int k = sn[i], nc = sn_ncols(S, k), f = sn_fsize(S, k);
        if (nc <= kb * NB) continue;
        int r0 = std::min(nc, (kb + 1) * NB);
        int nt = (f - r0 + TRSM_ROWS - 1) / TRSM_ROWS;
        for (int t = 0; t < nt; ++t) tasks.push_back(make_int2(k, t));
      }
      push_launch(h->factor_plan, 2, kb, off);
      off = (int64_t)tasks.size();
      for (int i = 0; i < cnt; ++i) {
        int k = sn[i], nc = sn_ncols(S, k), f = sn_fsize(S, k);
        if (nc <= kb * NB) continue;
        int r0 = std::min(nc, (kb + 1) * NB);
        int nt = tri_tiles(f - r0, UPD_TILE);
        for (int t = 0; t < nt; ++t) tasks.push_back(make_int2(k, t));
      }
      push_launch(h->factor_plan, 3, kb, off);
    }
    // forward solve
    int64_t off = (int64_t)tasks.size();
    for (int i = 0; i < cnt; ++i) tasks.push_back(make_int2(sn[i], 0));
    push_launch(h->fwd_plan, 4, 0, off);
    off = (int64_t)tasks.size();
    for (int i = 0; i < cnt; ++i) {
      int nt = (sn_nbelow(S, sn[i]) + FWD_ROWS - 1) / FWD_ROWS;
      for (int t = 0; t < nt; ++t) tasks.push_back(make_int2(sn[i], t));
    }
    push_launch(h->fwd_plan, 5, 0, off);
  }
  for (int l = S->nlevels - 1; l >= 0; --l) {
    const int* sn = S->level_sn.data() + S->level_ptr[l];
    int cnt = S->level_ptr[l + 1] - S->level_ptr[l];
    int64_t off = (int64_t)tasks.size();
    for (int i = 0; i < cnt; ++i) {
      int nt = (sn_ncols(S, sn[i]) + BWD_COLS - 1) / BWD_COLS;
      for (int t = 0; t < nt; ++t) tasks.push_back(make_int2(sn[i], t));
    }
    push_launch(h->bwd_plan, 6, 0, off);
    off = (int64_t)tasks.size();
    for (int i = 0; i < cnt; ++i) tasks.push_back(make_int2(sn[i], 0));
    push_launch(h->bwd_plan, 7, 0, off);
  }
  EIGD_CUDA(cudaMalloc((void**)&h->tasks, std::max<size_t>(tasks.size(), 1) * sizeof(int2)));
  EIGD_CUDA(cudaMemcpy(h->tasks, tasks.data(), tasks.size() * sizeof(int2), cudaMemcpyHostToDevice));
  return 0;
}

// ---------------------------------------------------------------------------------------
// integer kernel: CSR non-zero -> slot in front storage (bit-exact twin of the host routine)
// ---------------------------------------------------------------------------------------
__global__ void assembly_map_kernel(SymDev d, const int* __restrict__ indptr, const int* __restrict__ indices,
                                    int64_t* __restrict__ map, int* __restrict__ err) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= d.n) return;
  int pr = d.iperm[r];
  for (int p = indptr[r]; p < indptr[r + 1]; ++p) {
    int pc = d.iperm[indices[p]];
    if (pr < pc) { map[p] = -1; continue; }
    int k = d.col2sn[pc];
    int first = d.sn_first[k], nc = d.sn_first[k + 1] - first;
    int64_t rb = d.sn_rowptr[k], re = d.sn_rowptr[k + 1];
    int64_t f = nc + (re - rb);
    int64_t lr;
    if (pr < first + nc) lr = pr - first;
    else {
      int64_t lo = rb, hi = re;
      while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (d.sn_rows[mid] < pr) lo = mid + 1; else hi = mid;
      }
      if (lo >= re || d.sn_rows[lo] != pr) { atomicExch(err, 1); map[p] = -1; continue; }
      lr = nc + (lo - rb);
    }
    map[p] = d.front_off[k] + lr + (int64_t)(pc - first) * f;
  }
}

// ---------------------------------------------------------------------------------------
// numeric factorisation kernels
// ---------------------------------------------------------------------------------------
__global__ void absmax_kernel(int64_t nnz, const double* __restrict__ v, unsigned long long* __restrict__ out) {
  double m = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
    double a = fabs(v[e]);
    if (a == a) m = fmax(m, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

__global__ void scatter_values_kernel(int64_t nnz, const double* __restrict__ vals, const int64_t* __restrict__ map,
                                      double* __restrict__ fronts) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t m = map[e];
    if (m >= 0) fronts[m] = vals[e];
  }
}

__device__ __forceinline__ void tri_decode(int t, int& ti, int& tj) {
  int i = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((i + 1) * (i + 2) / 2 <= t) ++i;
  while (i * (i + 1) / 2 > t) --i;
  ti = i;
  tj = t - i * (i + 1) / 2;
}

// parent(rel[i], rel[j]) += child contribution (i, j), lower triangle, 32x32 tiles
__global__ void __launch_bounds__(256)
extend_add_kernel(SymDev d, const int2* __restrict__ tasks, double* __restrict__ fronts) {
  int2 tk = tasks[blockIdx.x];
  int c = tk.x;
  int ti, tj;
  tri_decode(tk.y, ti, tj);
  int p = d.sn_parent[c];
  int ncc = d.sn_first[c + 1] - d.sn_first[c];
  int nb = (int)(d.sn_rowptr[c + 1] - d.sn_rowptr[c]);
  int64_t fc = ncc + nb;
  int64_t fp = (d.sn_first[p + 1] - d.sn_first[p]) + (d.sn_rowptr[p + 1] - d.sn_rowptr[p]);
  const double* C = fronts + d.front_off[c] + ncc + (int64_t)ncc * fc;
  double* P = fronts + d.front_off[p];
  const int* rel = d.rel + d.sn_rowptr[c];
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int i = ti * EA_TILE + tx;
  if (i >= nb) return;
  int ri = rel[i];
  for (int jj = ty; jj < EA_TILE; jj += 8) {
    int j = tj * EA_TILE + jj;
    if (j < nb && j <= i) P[ri + (int64_t)rel[j] * fp] += C[i + (int64_t)j * fc];
  }
}

// LDL^T of one 32-wide pivot block + inverse of its unit-lower factor; one CTA per front
__global__ void __launch_bounds__(256)
diag_factor_kernel(SymDev d, const int2* __restrict__ tasks, int kb, double* __restrict__ fronts,
                   double* __restrict__ linv, double* __restrict__ dval, double* __restrict__ dinv,
                   const unsigned long long* __restrict__ amax_bits, double piv_tol,
                   unsigned long long* __restrict__ info) {
  __shared__ double T[NB][NB + 1];
  __shared__ double Li[NB][NB + 1];
  __shared__ double colj[NB];
  __shared__ double sh_d;
  __shared__ int sh_neg, sh_pert, sh_bad;
  int s = tasks[blockIdx.x].x;
  int first = d.sn_first[s];
  int nc = d.sn_first[s + 1] - first;
  int64_t f = nc + (d.sn_rowptr[s + 1] - d.sn_rowptr[s]);
  int j0 = kb * NB;
  int bs = min(NB, nc - j0);
  double* F = fronts + d.front_off[s];
  const int tid = threadIdx.x;
  if (tid == 0) { sh_neg = 0; sh_pert = 0; sh_bad = 0; }
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    int i = e % NB, j = e / NB;
    T[i][j] = (i < bs && j < bs && i >= j) ? F[(j0 + i) + (int64_t)(j0 + j) * f] : 0.0;
    Li[i][j] = (i == j) ? 1.0 : 0.0;
  }
  double thr = piv_tol * __longlong_as_double((long long)(*amax_bits));
  __syncthreads();
  for (int j = 0; j < bs; ++j) {
    if (tid == 0) {
      double dj = T[j][j];
      if (!(dj == dj) || fabs(dj) > 1e300) { sh_bad++; dj = thr > 0.0 ? thr : 1.0; }
      if (fabs(dj) < thr) { dj = (dj >= 0.0) ? thr : -thr; sh_pert++; }
      if (dj < 0.0) sh_neg++;
      T[j][j] = dj;
      sh_d = dj;
    }
    if (tid > j && tid < bs) colj[tid] = T[tid][j];
    __syncthreads();
    double rd = 1.0 / sh_d;
    for (int e = tid; e < bs * bs; e += blockDim.x) {
      int i = e % bs, c = e / bs;
      if (c > j && i >= c) T[i][c] = fma(-colj[i], colj[c] * rd, T[i][c]);
    }
    if (tid > j && tid < bs) T[tid][j] = colj[tid] * rd;
    __syncthreads();
  }
  // inverse of the unit lower triangle, one column per thread
  if (tid < bs) {
    int c = tid;
    for (int i = c + 1; i < bs; ++i) {
      double sum = 0.0;
      for (int k = c; k < i; ++k) sum = fma(T[i][k], Li[k][c], sum);
      Li[i][c] = -sum;
    }
  }
  __syncthreads();
  for (int e = tid; e < bs * bs; e += blockDim.x) {
    int i = e % bs, j = e / bs;
    if (i >= j) F[(j0 + i) + (int64_t)(j0 + j) * f] = T[i][j];
  }
  double* Lo = linv + d.linv_off[s] + (int64_t)kb * NB * NB;
  for (int e = tid; e < NB * NB; e += blockDim.x) Lo[e] = Li[e / NB][e % NB];
  if (tid < bs) {
    dval[first + j0 + tid] = T[tid][tid];
    dinv[first + j0 + tid] = 1.0 / T[tid][tid];
  }
  if (tid == 0) {
    if (sh_neg) atomicAdd(&info[0], (unsigned long long)sh_neg);
    if (sh_pert) atomicAdd(&info[1], (unsigned long long)sh_pert);
    if (sh_bad) atomicAdd(&info[2], (unsigned long long)sh_bad);
  }
}

// rows below the pivot block: L_row = (F_row * Linv^T) * D^-1 ; one thread per row
__global__ void __launch_bounds__(TRSM_ROWS)
trsm_kernel(SymDev d, const int2* __restrict__ tasks, int kb, double* __restrict__ fronts,
            const double* __restrict__ linv, const double* __restrict__ dinv) {
  __shared__ double Li[NB][NB + 1];
  __shared__ double di[NB];
  int2 tk = tasks[blockIdx.x];
  int s = tk.x;
  int first = d.sn_first[s];
  int nc = d.sn_first[s + 1] - first;
  int64_t f = nc + (d.sn_rowptr[s + 1] - d.sn_rowptr[s]);
  int j0 = kb * NB;
  int bs = min(NB, nc - j0);
  const double* Lg = linv + d.linv_off[s] + (int64_t)kb * NB * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) Li[e / NB][e % NB] = Lg[e];
  if (threadIdx.x < NB) di[threadIdx.x] = (threadIdx.x < bs) ? dinv[first + j0 + threadIdx.x] : 0.0;
  __syncthreads();
  int64_t i = (int64_t)j0 + bs + (int64_t)tk.y * TRSM_ROWS + threadIdx.x;
  if (i >= f) return;
  double* F = fronts + d.front_off[s] + i + (int64_t)j0 * f;
  double x[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) x[k] = (k < bs) ? F[(int64_t)k * f] : 0.0;
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k <= c; ++k) sum = fma(x[k], Li[c][k], sum);
    if (c < bs) F[(int64_t)c * f] = sum * di[c];
  }
}

// trailing update C(i,c) -= sum_k L(i,k) d_k L(c,k) over one pivot block; 64x64 lower tiles
__global__ void __launch_bounds__(256)
trailing_update_kernel(SymDev d, const int2* __restrict__ tasks, int kb, double* __restrict__ fronts,
                       const double* __restrict__ dval) {
  __shared__ double As[UPD_TILE][NB + 1];
  __shared__ double Bs[UPD_TILE][NB + 1];
  int2 tk = tasks[blockIdx.x];
  int s = tk.x;
  int ti, tj;
  tri_decode(tk.y, ti, tj);
  int first = d.sn_first[s];
  int nc = d.sn_first[s + 1] - first;
  int64_t f = nc + (d.sn_rowptr[s + 1] - d.sn_rowptr[s]);
  int j0 = kb * NB;
  int bs = min(NB, nc - j0);
  int64_t r0 = j0 + bs;
  double* F = fronts + d.front_off[s];
  int64_t ib = r0 + (int64_t)ti * UPD_TILE, cb = r0 + (int64_t)tj * UPD_TILE;
  for (int e = threadIdx.x; e < UPD_TILE * NB; e += blockDim.x) {
    int r = e % UPD_TILE, k = e / UPD_TILE;
    double a = 0.0, b = 0.0;
    if (k < bs) {
      if (ib + r < f) a = F[(ib + r) + (int64_t)(j0 + k) * f];
      if (cb + r < f) b = F[(cb + r) + (int64_t)(j0 + k) * f] * dval[first + j0 + k];
    }
    As[r][k] = a;
    Bs[r][k] = b;
  }
  __syncthreads();
  int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 8
  for (int k = 0; k < NB; ++k) {
    double av[4], bv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) av[a] = As[tx + 16 * a][k];
#pragma unroll
    for (int b = 0; b < 4; ++b) bv[b] = Bs[ty + 16 * b][k];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
  }
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    int64_t c = cb + ty + 16 * b;
    if (c >= f) continue;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      int64_t i = ib + tx + 16 * a;
      if (i < f && i >= c) F[i + c * f] -= acc[a][b];
    }
  }
}

// ---------------------------------------------------------------------------------------
// solve kernels (k right-hand sides; work vectors row-major with stride k)
// ---------------------------------------------------------------------------------------
// forward, pivot rows: gather b, add the children's updates, w1 <- L11^{-1} w1
__global__ void __launch_bounds__(256)
fwd_diag_kernel(SymDev d, const int2* __restrict__ tasks, const double* __restrict__ fronts,
                const double* __restrict__ linv, double* __restrict__ wbuf, const double* __restrict__ B,
                int64_t brs, int64_t bcs, int k) {
  extern __shared__ double W1[];  // nc * k
  __shared__ double tmp[NB * 32];
  int s = tasks[blockIdx.x].x;
  int first = d.sn_first[s];
  int nc = d.sn_first[s + 1] - first;
  int nb = (int)(d.sn_rowptr[s + 1] - d.sn_rowptr[s]);
  int64_t f = nc + nb;
  const double* F = fronts + d.front_off[s];
  double* w = wbuf + d.w_off[s] * k;
  const int tid = threadIdx.x;
  for (int e = tid; e < nc * k; e += blockDim.x) {
    int i = e / k, r = e - i * k;
    W1[e] = B[(int64_t)d.perm[first + i] * brs + (int64_t)r * bcs];
  }
  for (int64_t e = tid; e < (int64_t)nb * k; e += blockDim.x) w[(int64_t)nc * k + e] = 0.0;
  __syncthreads();
  for (int q = d.child_ptr[s]; q < d.child_ptr[s + 1]; ++q) {
    int c = d.child_idx[q];
    int ncc = d.sn_first[c + 1] - d.sn_first[c];
    int nbc = (int)(d.sn_rowptr[c + 1] - d.sn_rowptr[c]);
    const double* wc = wbuf + (d.w_off[c] + ncc) * k;
    const int* rel = d.rel + d.sn_rowptr[c];
    for (int e = tid; e < nbc * k; e += blockDim.x) {
      int i = e / k, r = e - i * k;
      int t = rel[i];
      if (t < nc) W1[t * k + r] += wc[e];
      else w[(int64_t)t * k + r] += wc[e];
    }
    __syncthreads();
  }
  int nblk = (nc + NB - 1) / NB;
  for (int kb = 0; kb < nblk; ++kb) {
    int j0 = kb * NB;
    int bs = min(NB, nc - j0);
    const double* Li = linv + d.linv_off[s] + (int64_t)kb * NB * NB;
    for (int e = tid; e < bs * k; e += blockDim.x) {
      int i = e / k, r = e - i * k;
      double sum = 0.0;
      for (int kk = 0; kk <= i; ++kk) sum = fma(Li[i * NB + kk], W1[(j0 + kk) * k + r], sum);
      tmp[e] = sum;
    }
    __syncthreads();
    for (int e = tid; e < bs * k; e += blockDim.x) W1[j0 * k + e] = tmp[e];
    __syncthreads();
    int rest = nc - j0 - bs;
    for (int e = tid; e < rest * k; e += blockDim.x) {
      int i = j0 + bs + e / k, r = e % k;
      double sum = 0.0;
      for (int kk = 0; kk < bs; ++kk) sum = fma(F[i + (int64_t)(j0 + kk) * f], W1[(j0 + kk) * k + r], sum);
      W1[i * k + r] -= sum;
    }
    __syncthreads();
  }
  for (int e = tid; e < nc * k; e += blockDim.x) w[e] = W1[e];
}

// forward, rows below the pivots: w2 -= L21 * w1 ; one thread per row
template <int KT>
__global__ void __launch_bounds__(FWD_ROWS)
fwd_update_kernel(SymDev d, const int2* __restrict__ tasks, const double* __restrict__ fronts,
                  double* __restrict__ wbuf, int k) {
  extern __shared__ double X1[];  // nc * k
  int2 tk = tasks[blockIdx.x];
  int s = tk.x;
  int nc = d.sn_first[s + 1] - d.sn_first[s];
  int nb = (int)(d.sn_rowptr[s + 1] - d.sn_rowptr[s]);
  int64_t f = nc + nb;
  double* w = wbuf + d.w_off[s] * k;
  for (int e = threadIdx.x; e < nc * k; e += blockDim.x) X1[e] = w[e];
  __syncthreads();
  int i = tk.y * FWD_ROWS + threadIdx.x;
  if (i >= nb) return;
  const double* Lrow = fronts + d.front_off[s] + nc + i;
  double acc[KT];
#pragma unroll
  for (int r = 0; r < KT; ++r) acc[r] = 0.0;
#pragma unroll 4
  for (int j = 0; j < nc; ++j) {
    double l = Lrow[(int64_t)j * f];
#pragma unroll
    for (int r = 0; r < KT; ++r)
      if (r < k) acc[r] = fma(l, X1[j * k + r], acc[r]);
  }
  double* wr = w + (int64_t)(nc + i) * k;
#pragma unroll
  for (int r = 0; r < KT; ++r)
    if (r < k) wr[r] -= acc[r];
}

// backward, w1 <- D^{-1} w1 - L21^T x(below rows) ; one warp per pivot column
template <int KT>
__global__ void __launch_bounds__(BWD_COLS * 32)
bwd_update_kernel(SymDev d, const int2* __restrict__ tasks, const double* __restrict__ fronts,
                  double* __restrict__ wbuf, const double* __restrict__ xperm, const double* __restrict__ dinv, int k) {
  int2 tk = tasks[blockIdx.x];
  int s = tk.x;
  int first = d.sn_first[s];
  int nc = d.sn_first[s + 1] - first;
  int nb = (int)(d.sn_rowptr[s + 1] - d.sn_rowptr[s]);
  int64_t f = nc + nb;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int j = tk.y * BWD_COLS + warp;
  if (j >= nc) return;
  const double* Lcol = fronts + d.front_off[s] + nc + (int64_t)j * f;
  const int* rows = d.sn_rows + d.sn_rowptr[s];
  double acc[KT];
#pragma unroll
  for (int r = 0; r < KT; ++r) acc[r] = 0.0;
  for (int i = lane; i < nb; i += 32) {
    double l = Lcol[i];
    const double* xr = xperm + (int64_t)rows[i] * k;
#pragma unroll
    for (int r = 0; r < KT; ++r)
      if (r < k) acc[r] = fma(l, xr[r], acc[r]);
  }
#pragma unroll
  for (int r = 0; r < KT; ++r) acc[r] = warp_sum(acc[r]);
  if (lane == 0) {
    double* w = wbuf + (d.w_off[s] + j) * k;
    double di = dinv[first + j];
#pragma unroll
    for (int r = 0; r < KT; ++r)
      if (r < k) w[r] = w[r] * di - acc[r];
  }
}

// backward, pivot rows: x1 <- L11^{-T} w1, scatter to the permuted and the original ordering
__global__ void __launch_bounds__(256)
bwd_diag_kernel(SymDev d, const int2* __restrict__ tasks, const double* __restrict__ fronts,
                const double* __restrict__ linv, const double* __restrict__ wbuf, double* __restrict__ xperm,
                double* __restrict__ X, int64_t xrs, int64_t xcs, int k) {
  extern __shared__ double W1[];
  __shared__ double tmp[NB * 32];
  int s = tasks[blockIdx.x].x;
  int first = d.sn_first[s];
  int nc = d.sn_first[s + 1] - first;
  int64_t f = nc + (d.sn_rowptr[s + 1] - d.sn_rowptr[s]);
  const double* F = fronts + d.front_off[s];
  const double* w = wbuf + d.w_off[s] * k;
  const int tid = threadIdx.x;
  for (int e = tid; e < nc * k; e += blockDim.x) W1[e] = w[e];
  __syncthreads();
  int nblk = (nc + NB - 1) / NB;
  for (int kb = nblk - 1; kb >= 0; --kb) {
    int j0 = kb * NB;
    int bs = min(NB, nc - j0);
    const double* Li = linv + d.linv_off[s] + (int64_t)kb * NB * NB;
    for (int e = tid; e < bs * k; e += blockDim.x) {
      int i = e / k, r = e - i * k;
      double sum = 0.0;
      for (int kk = i; kk < bs; ++kk) sum = fma(Li[kk * NB + i], W1[(j0 + kk) * k + r], sum);
      tmp[e] = sum;
    }
    __syncthreads();
    for (int e = tid; e < bs * k; e += blockDim.x) W1[j0 * k + e] = tmp[e];
    __syncthreads();
    for (int e = tid; e < j0 * k; e += blockDim.x) {
      int j = e / k, r = e - j * k;
      const double* Lc = F + j0 + (int64_t)j * f;
      double sum = 0.0;
      for (int i = 0; i < bs; ++i) sum = fma(Lc[i], W1[(j0 + i) * k + r], sum);
      W1[e] -= sum;
    }
    __syncthreads();
  }
  for (int e = tid; e < nc * k; e += blockDim.x) {
    int i = e / k, r = e - i * k;
    double v = W1[e];
    xperm[(int64_t)(first + i) * k + r] = v;
    X[(int64_t)d.perm[first + i] * xrs + (int64_t)r * xcs] = v;
  }
}

}  // namespace

struct eigd_factor {
  eigd_symbolic* sym = nullptr;
  SymDevHolder* h = nullptr;
  int max_rhs = 1;
  double* fronts = nullptr;
  double* linv = nullptr;
  double* dval = nullptr;
  double* dinv = nullptr;
  double* wbuf = nullptr;
  double* xperm = nullptr;
  unsigned long long* amax = nullptr;  // 1 value
  unsigned long long* info = nullptr;  // 4 values
  double piv_tol = 1e-11;
  int64_t bytes = 0;
  bool attrs_set = false;
  char* base = nullptr;
  bool owns = true;
};

extern "C" int eigd_symbolic_assembly_map_device(eigd_symbolic* s, int n, const int* d_indptr, const int* d_indices,
                                                 int64_t* d_map) {
  if (n != s->n) { eigd_set_error("assembly_map_device: n mismatch"); return 1; }
  int rc = build_symdev(s);
  if (rc) return rc;
  auto* h = (SymDevHolder*)s->dev;
  int* err = nullptr;
  EIGD_CUDA(cudaMalloc((void**)&err, sizeof(int)));
  EIGD_CUDA(cudaMemsetAsync(err, 0, sizeof(int), g_eigd_stream));
  EIGD_LAUNCH(assembly_map_kernel, (n + 127) / 128, 128, 0, h->d, d_indptr, d_indices, d_map, err);
  EIGD_CHECK_LAUNCH();
  int herr = 0;
  EIGD_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, g_eigd_stream));
  EIGD_CUDA(cudaStreamSynchronize(g_eigd_stream));
  cudaFree(err);
  if (herr) { eigd_set_error("assembly_map_device: matrix entry outside the symbolic pattern"); return 2; }
  return 0;
}

static inline int64_t align256(int64_t b) { return (b + 255) / 256 * 256; }

// sizes (bytes, 256-aligned) of the eight device arrays of a factor, in carving order
static void factor_layout(const eigd_symbolic* s, int64_t linv_total, int max_rhs, int64_t sz[8]) {
  int64_t nfront = s->front_off[s->nsuper], sumf = s->w_off[s->nsuper];
  sz[0] = align256(nfront * 8);                        // fronts
  sz[1] = align256(linv_total * 8);                    // linv
  sz[2] = align256((int64_t)s->n * 8);                 // dval
  sz[3] = align256((int64_t)s->n * 8);                 // dinv
  sz[4] = align256(sumf * max_rhs * 8);                // wbuf
  sz[5] = align256((int64_t)s->n * max_rhs * 8);       // xperm
  sz[6] = 256;                                         // amax
  sz[7] = 256;                                         // info
}

extern "C" int64_t eigd_factor_workspace_bytes(eigd_symbolic* s, int max_rhs) {
  if (build_symdev(s)) return -1;
  int64_t sz[8], tot = 0;
  factor_layout(s, ((SymDevHolder*)s->dev)->linv_total, std::max(1, max_rhs), sz);
  for (int i = 0; i < 8; ++i) tot += sz[i];
  return tot + 256;
}

// d_workspace == NULL: the library allocates (cudaMalloc) and owns the arrays; otherwise they are carved
// out of the caller's buffer (a torch tensor: the caching allocator then recycles it between factors)
extern "C" int eigd_factor_create_in(eigd_symbolic* s, int max_rhs, void* d_workspace, int64_t workspace_bytes,
                                     eigd_factor** out) {
  int rc = build_symdev(s);
  if (rc) return rc;
  auto* f = new eigd_factor();
  f->sym = s;
  f->h = (SymDevHolder*)s->dev;
  f->max_rhs = std::max(1, max_rhs);
  int64_t sz[8], tot = 0;
  factor_layout(s, f->h->linv_total, f->max_rhs, sz);
  for (int i = 0; i < 8; ++i) tot += sz[i];
  char* base = nullptr;
  if (d_workspace) {
    base = (char*)(((uintptr_t)d_workspace + 255) / 256 * 256);
    if ((base - (char*)d_workspace) + tot > workspace_bytes) {
      eigd_set_error("factor_create_in: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)(tot + 256));
      delete f;
      return 3;
    }
    f->owns = false;
  } else {
    cudaError_t e = cudaMalloc((void**)&base, (size_t)tot);
    if (e != cudaSuccess) { eigd_set_error("factor_create: cudaMalloc(%lld) -> %s", (long long)tot, cudaGetErrorString(e)); delete f; return 100 + (int)e; }
    f->owns = true;
  }
  f->base = base;
  f->bytes = tot;
  char* p = base;
  f->fronts = (double*)p; p += sz[0];
  f->linv = (double*)p; p += sz[1];
  f->dval = (double*)p; p += sz[2];
  f->dinv = (double*)p; p += sz[3];
  f->wbuf = (double*)p; p += sz[4];
  f->xperm = (double*)p; p += sz[5];
  f->amax = (unsigned long long*)p; p += sz[6];
  f->info = (unsigned long long*)p;
  *out = f;
  return 0;
}

extern "C" int eigd_factor_create(eigd_symbolic* s, int max_rhs, eigd_factor** out) {
  return eigd_factor_create_in(s, max_rhs, nullptr, 0, out);
}

extern "C" void eigd_factor_destroy(eigd_factor* f) {
  if (!f) return;
  if (f->owns && f->base) cudaFree(f->base);
  delete f;
}

extern "C" int64_t eigd_factor_bytes(const eigd_factor* f) { return f->bytes; }

extern "C" int eigd_factor_numeric(eigd_factor* f, int64_t nnz, const double* d_vals, const int64_t* d_map) {
  eigd_symbolic* s = f->sym;
  SymDevHolder* h = f->h;
  int64_t nfront = s->front_off[s->nsuper];
  EIGD_CUDA(cudaMemsetAsync(f->fronts, 0, (size_t)nfront * 8, g_eigd_stream));
  EIGD_CUDA(cudaMemsetAsync(f->amax, 0, 8, g_eigd_stream));
  EIGD_CUDA(cudaMemsetAsync(f->info, 0, 32, g_eigd_stream));
  int g = (int)std::min<int64_t>((nnz + 255) / 256, 148 * 16);
  if (g < 1) g = 1;
  EIGD_LAUNCH(absmax_kernel, g, 256, 0, nnz, d_vals, f->amax);
  EIGD_CHECK_LAUNCH();
  EIGD_LAUNCH(scatter_values_kernel, g, 256, 0, nnz, d_vals, d_map, f->fronts);
  EIGD_CHECK_LAUNCH();
  for (const Launch& L : h->factor_plan) {
    const int2* t = h->tasks + L.off;
    switch (L.kind) {
      case 0: EIGD_LAUNCH(extend_add_kernel, L.count, 256, 0, h->d, t, f->fronts); break;
      case 1: EIGD_LAUNCH(diag_factor_kernel, L.count, 256, 0, h->d, t, L.kb, f->fronts, f->linv, f->dval, f->dinv, f->amax, f->piv_tol, f->info); break;
      case 2: EIGD_LAUNCH(trsm_kernel, L.count, TRSM_ROWS, 0, h->d, t, L.kb, f->fronts, f->linv, f->dinv); break;
      case 3: EIGD_LAUNCH(trailing_update_kernel, L.count, 256, 0, h->d, t, L.kb, f->fronts, f->dval); break;
      default: break;
    }
    EIGD_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int eigd_factor_info(eigd_factor* f, int64_t* info3) {
  unsigned long long hinfo[4];
  EIGD_CUDA(cudaMemcpyAsync(hinfo, f->info, 32, cudaMemcpyDeviceToHost, g_eigd_stream));
  EIGD_CUDA(cudaStreamSynchronize(g_eigd_stream));
  for (int i = 0; i < 3; ++i) info3[i] = (int64_t)hinfo[i];
  return 0;
}

template <int KT>
static int solve_chunk(eigd_factor* f, const double* B, int64_t brs, int64_t bcs, double* X, int64_t xrs, int64_t xcs, int k) {
  SymDevHolder* h = f->h;
  size_t smem = (size_t)f->sym->maxcols * k * sizeof(double);
  if (!f->attrs_set) {
    int maxs = 256 * 32 * 8;
    EIGD_CUDA(cudaFuncSetAttribute(fwd_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    EIGD_CUDA(cudaFuncSetAttribute(bwd_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    EIGD_CUDA(cudaFuncSetAttribute(fwd_update_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    EIGD_CUDA(cudaFuncSetAttribute(fwd_update_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    EIGD_CUDA(cudaFuncSetAttribute(fwd_update_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    EIGD_CUDA(cudaFuncSetAttribute(fwd_update_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    EIGD_CUDA(cudaFuncSetAttribute(fwd_update_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
    f->attrs_set = true;
  }
  for (const Launch& L : h->fwd_plan) {
    const int2* t = h->tasks + L.off;
    if (L.kind == 4) EIGD_LAUNCH(fwd_diag_kernel, L.count, 256, smem, h->d, t, f->fronts, f->linv, f->wbuf, B, brs, bcs, k);
    else EIGD_LAUNCH(fwd_update_kernel<KT>, L.count, FWD_ROWS, smem, h->d, t, f->fronts, f->wbuf, k);
    EIGD_CHECK_LAUNCH();
  }
  for (const Launch& L : h->bwd_plan) {
    const int2* t = h->tasks + L.off;
    if (L.kind == 6) EIGD_LAUNCH(bwd_update_kernel<KT>, L.count, BWD_COLS * 32, 0, h->d, t, f->fronts, f->wbuf, f->xperm, f->dinv, k);
    else EIGD_LAUNCH(bwd_diag_kernel, L.count, 256, smem, h->d, t, f->fronts, f->linv, f->wbuf, f->xperm, X, xrs, xcs, k);
    EIGD_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int eigd_factor_solve(eigd_factor* f, const double* B, int64_t brs, int64_t bcs, double* X, int64_t xrs,
                                 int64_t xcs, int k) {
  if (k <= 0) return 0;
  int chunk = std::min(32, f->max_rhs);
  for (int c0 = 0; c0 < k; c0 += chunk) {
    int kc = std::min(chunk, k - c0);
    const double* Bc = B + (int64_t)c0 * bcs;
    double* Xc = X + (int64_t)c0 * xcs;
    int rc;
    if (kc == 1) rc = solve_chunk<1>(f, Bc, brs, bcs, Xc, xrs, xcs, kc);
    else if (kc <= 4) rc = solve_chunk<4>(f, Bc, brs, bcs, Xc, xrs, xcs, kc);
    else if (kc <= 8) rc = solve_chunk<8>(f, Bc, brs, bcs, Xc, xrs, xcs, kc);
    else if (kc <= 16) rc = solve_chunk<16>(f, Bc, brs, bcs, Xc, xrs, xcs, kc);
    else rc = solve_chunk<32>(f, Bc, brs, bcs, Xc, xrs, xcs, kc);
    if (rc) return rc;
  }
  return 0;
}

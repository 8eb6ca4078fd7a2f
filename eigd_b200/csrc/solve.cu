// Multi-RHS sparse triangular solves x = (L D L^T)^{-1} b as ONE persistent kernel.
// Replaces SuperLU.solve behind SpLuOperator._matvec (reference eigd/eigenvector_derivatives.py:18-23),
// which the reference calls once per right-hand-side column.
//
// Formulation.  The factorisation leaves, for every front (supernode with nc pivot columns and nb
// rows below them, f = nc + nb), the dense solve panel S = [L11^{-1} ; -L21 L11^{-1}] (f x nc) and
// its transpose.  With w1 = (b restricted to the pivot rows) + (updates gathered from the children):
//     forward :  [y1 ; u2] = S w1,   z1 = D^{-1} y1 kept, u2 (+ gathered child rows) passed up;
//     backward:  x1 = S^T [z1 ; x2], x2 = solution entries of the front's below rows (ancestors).
// Neither sweep has a dependency inside a front, so a level of the assembly tree is one batch of
// independent dense products: the kernel walks 2 * nlevels phases separated by grid barriers.
//
// Mapping.  512-thread CTAs, one (or two) per SM, cooperative launch.  A warp tile is 32 consecutive
// outputs of one front (lane = output), the reduction dimension is streamed in chunks of 32 whose
// input vectors are staged in shared memory ([k][32] per warp) while the panel entries are read
// straight from HBM, coalesced across the lanes (column-major panels: consecutive lanes = consecutive
// addresses).  On the upper levels, where a level has few tiles, ws warps share a tile and split its
// reduction dimension; partial sums meet in shared memory.  Update vectors are gathered (pull lists
// built on the host), never scattered: no atomics, bitwise reproducible sums.
//
// Algorithmic bytes per call: 2*(nnz(L)*8 + idx) + n*8 + 4*n*k*8.
#include "factor_internal.cuh"
#include "../../include/eigd_b200.h"

#include <algorithm>

namespace {

struct SolveArgs {
  const TileRec* tiles;
  const PhaseRec* phases;
  int nphases;
  const int2* pull2;
  const int* ovf;
  const int* perm;
  const int* sn_rows;
  const double* sfwd;
  const double* sbwd;
  const double* dinv;
  double* wbuf;
  double* ybuf;
  double* xperm;
  const double* B;
  int64_t brs, bcs;
  double* X;
  int64_t xrs, xcs;
  int k;
  unsigned long long* barrier;
  unsigned long long bar_base;
};

// vectors produced earlier in the same launch by other SMs are read through L2 (ld.global.cg): L1 is
// not coherent across SMs and a line fetched in an earlier phase may be stale
template <int KT>
__device__ __forceinline__ void add_row(const double* __restrict__ base, int64_t row, int k, double* v) {
  const double* p = base + row * k;
#pragma unroll
  for (int r = 0; r < KT; ++r)
    if (r < k) v[r] += __ldcg(p + r);
}

// v += sum of the forward-sweep updates addressed to w-row t (fixed order: deterministic)
template <int KT>
__device__ __forceinline__ void pull_add(const SolveArgs& a, int64_t t, double* v) {
  const int2 pp = __ldg(&a.pull2[t]);
  if (pp.x >= 0) add_row<KT>(a.wbuf, pp.x, a.k, v);
  if (pp.y >= 0) add_row<KT>(a.wbuf, pp.y, a.k, v);
  else if (pp.y <= -2) {
    const int o = -2 - pp.y;
    const int cnt = __ldg(&a.ovf[o]);
    for (int q = 0; q < cnt; ++q) add_row<KT>(a.wbuf, __ldg(&a.ovf[o + 1 + q]), a.k, v);
  }
}

__device__ __forceinline__ TileRec load_tile(const TileRec* p) {
  const int4* q = reinterpret_cast<const int4*>(p);
  int4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  TileRec t;
  t.first = a.x; t.nc = a.y; t.nb = a.z; t.tile = a.w;
  t.soff = ((int64_t)(unsigned)b.x) | ((int64_t)b.y << 32);
  t.w_off = ((int64_t)(unsigned)b.z) | ((int64_t)b.w << 32);
  t.row_off = ((int64_t)(unsigned)c.x) | ((int64_t)c.y << 32);
  t.pad = 0;
  return t;
}

// acc += M[out, c0:c1) * in[c0:c1) for this lane's output; M column-major with leading dimension ld.
// stage_fn(c, v) fills v[0:k) with input vector entry c (lane-parallel), the chunk is then shared
// through the warp's staging buffer.
template <int KT, class StageFn>
__device__ __forceinline__ void warp_panel_product(const double* __restrict__ M, int64_t ld, int c0, int c1, int lane,
                                                   double* stage, double* acc, StageFn stage_fn) {
  for (int cc = c0; cc < c1; cc += 32) {
    const int ncol = min(32, c1 - cc);
    double v[KT];
#pragma unroll
    for (int r = 0; r < KT; ++r) v[r] = 0.0;
    if (lane < ncol) stage_fn(cc + lane, v);
#pragma unroll
    for (int r = 0; r < KT; ++r) stage[r * 32 + lane] = v[r];
    __syncwarp();
    const double* Mc = M + (int64_t)cc * ld;
#pragma unroll 8
    for (int t = 0; t < ncol; ++t) {
      const double m = __ldg(Mc + (int64_t)t * ld);
#pragma unroll
      for (int r = 0; r < KT; ++r) acc[r] = fma(m, stage[r * 32 + t], acc[r]);
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void grid_barrier(unsigned long long* ctr, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1ULL);
    unsigned long long v;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

template <int KT>
__global__ void __launch_bounds__(SOLVE_WARPS * 32, 1) solve_kernel(SolveArgs a) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* stage = smem + warp * (32 * KT);                 // [KT][32] per warp
  double* part = smem + SOLVE_WARPS * 32 * KT;             // [SOLVE_WARPS][KT][32] partial sums
  const int k = a.k;
  unsigned long long target = a.bar_base;
  for (int p = 0; p < a.nphases; ++p) {
    const int4 ph4 = __ldg(reinterpret_cast<const int4*>(a.phases + p));
    const int64_t tile_off = __ldg(&a.phases[p].tile_off);
    const int dir = ph4.x, ws = ph4.y, ntiles = ph4.z;
    const int tpc = SOLVE_WARPS / ws;
    const int sub = warp / ws, slice = warp - sub * ws;
    const int nct = (ntiles + tpc - 1) / tpc;
    for (int ct = blockIdx.x; ct < nct; ct += gridDim.x) {
      const int te = ct * tpc + sub;
      const bool have = te < ntiles;
      double acc[KT];
#pragma unroll
      for (int r = 0; r < KT; ++r) acc[r] = 0.0;
      TileRec tr;
      tr.first = tr.nc = tr.nb = tr.tile = 0;
      tr.soff = tr.w_off = tr.row_off = 0;
      int out = 0;          // this lane's output index inside the front
      if (have) {
        tr = load_tile(a.tiles + tile_off + te);
        const int nc = tr.nc, f = tr.nc + tr.nb;
        const int o0 = tr.tile * SOLVE_TILE;
        out = o0 + lane;
        if (dir == 0) {
          // ---- forward: outputs are front rows; acc = S[row, 0:cend) w1 (+ gathered updates of the row)
          if (slice == 0 && out >= nc && out < f) pull_add<KT>(a, tr.w_off + out, acc);
          const int cend = min(nc, o0 + SOLVE_TILE);          // S is lower triangular inside the pivot rows
          const int per = (cend + ws - 1) / ws;
          const int c0 = slice * per, c1 = min(cend, c0 + per);
          const double* M = a.sfwd + tr.soff + min(out, f - 1);
          warp_panel_product<KT>(M, f, c0, c1, lane, stage, acc, [&](int c, double* v) {
            const int64_t po = __ldg(&a.perm[tr.first + c]);
            const double* bp = a.B + po * a.brs;
#pragma unroll
            for (int r = 0; r < KT; ++r)
              if (r < k) v[r] = bp[(int64_t)r * a.bcs];
            pull_add<KT>(a, tr.w_off + c, v);
          });
        } else {
          // ---- backward: outputs are pivot columns; acc = S^T[col, o0:f) [z1 ; x2]
          const int len = f - o0;
          const int per = (len + ws - 1) / ws;
          const int i0 = o0 + slice * per, i1 = min(f, i0 + per);
          const double* M = a.sbwd + tr.soff + min(out, nc - 1);
          warp_panel_product<KT>(M, nc, i0, i1, lane, stage, acc, [&](int i, double* v) {
            if (i < nc) add_row<KT>(a.ybuf, tr.first + i, k, v);
            else add_row<KT>(a.xperm, __ldg(&a.sn_rows[tr.row_off + i - nc]), k, v);
          });
        }
      }
      if (ws > 1) {
#pragma unroll
        for (int r = 0; r < KT; ++r) part[(warp * KT + r) * 32 + lane] = acc[r];
        __syncthreads();
        if (have && slice == 0)
          for (int s = 1; s < ws; ++s)
#pragma unroll
            for (int r = 0; r < KT; ++r) acc[r] += part[((warp + s) * KT + r) * 32 + lane];
      }
      if (have && slice == 0) {
        const int nc = tr.nc, f = tr.nc + tr.nb;
        if (dir == 0) {
          if (out < nc) {
            const double di = __ldg(&a.dinv[tr.first + out]);
            double* y = a.ybuf + (int64_t)(tr.first + out) * k;
#pragma unroll
            for (int r = 0; r < KT; ++r)
              if (r < k) y[r] = di * acc[r];
          } else if (out < f) {
            double* w = a.wbuf + (tr.w_off + out) * k;
#pragma unroll
            for (int r = 0; r < KT; ++r)
              if (r < k) w[r] = acc[r];
          }
        } else if (out < nc) {
          double* xp = a.xperm + (int64_t)(tr.first + out) * k;
          double* xo = a.X + (int64_t)__ldg(&a.perm[tr.first + out]) * a.xrs;
#pragma unroll
          for (int r = 0; r < KT; ++r)
            if (r < k) { xp[r] = acc[r]; xo[(int64_t)r * a.xcs] = acc[r]; }
        }
      }
      if (ws > 1) __syncthreads();
    }
    target += gridDim.x;
    grid_barrier(a.barrier, target);
  }
}

template <class T>
int upload_vec(SymDevHolder* h, const std::vector<T>& v, T** out) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  EIGD_CUDA(cudaMalloc(&p, bytes));
  h->allocs.push_back(p);
  if (!v.empty()) EIGD_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = (T*)p;
  return 0;
}

struct KernelCfg {
  bool ready = false;
  int grid = 0;
  size_t smem = 0;
};
KernelCfg g_cfg[5];
int g_num_sms = 0;

template <int KT>
int configure(int slot) {
  KernelCfg& c = g_cfg[slot];
  if (c.ready) return 0;
  c.smem = (size_t)SOLVE_WARPS * 32 * KT * 8 * 2;
  EIGD_CUDA(cudaFuncSetAttribute(solve_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
  int occ = 0;
  EIGD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, solve_kernel<KT>, SOLVE_WARPS * 32, c.smem));
  if (occ < 1) { eigd_set_error("solve: kernel does not fit on an SM"); return 5; }
  c.grid = g_num_sms * std::min(occ, 2);
  c.ready = true;
  return 0;
}

template <int KT>
int launch_solve(int slot, eigd_factor* f, SolveArgs& a) {
  int rc = configure<KT>(slot);
  if (rc) return rc;
  const KernelCfg& c = g_cfg[slot];
  a.barrier = f->barrier;
  a.bar_base = f->bar_base;
  void* params[] = {(void*)&a};
  EIGD_CUDA(cudaLaunchCooperativeKernel((void*)solve_kernel<KT>, dim3(c.grid), dim3(SOLVE_WARPS * 32), params, c.smem,
                                        g_eigd_stream));
  ++g_eigd_launches;
  f->bar_base += (unsigned long long)a.nphases * (unsigned long long)c.grid;
  return 0;
}

}  // namespace

int build_solve_plan_dev(eigd_symbolic* S, SymDevHolder* h) {
  if (!g_num_sms) {
    int dev = 0;
    EIGD_CUDA(cudaGetDevice(&dev));
    EIGD_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  SolvePlanHost P;
  build_solve_plan_host(S, g_num_sms * 2 * SOLVE_WARPS, P);
  SolvePlanDev& d = h->solve;
  int rc = 0;
  rc |= upload_vec(h, P.tiles, &d.tiles);
  rc |= upload_vec(h, P.phases, &d.phases);
  std::vector<int2> p2(P.pull2.size() / 2);
  for (size_t i = 0; i < p2.size(); ++i) p2[i] = make_int2(P.pull2[2 * i], P.pull2[2 * i + 1]);
  rc |= upload_vec(h, p2, &d.pull2);
  rc |= upload_vec(h, P.ovf, &d.ovf);
  d.nphases = (int)P.phases.size();
  d.host_phases = P.phases;
  return rc;
}

void free_solve_plan_dev(SymDevHolder*) {}   // the arrays are owned by SymDevHolder::allocs

extern "C" int eigd_factor_solve(eigd_factor* f, const double* B, int64_t brs, int64_t bcs, double* X, int64_t xrs,
                                 int64_t xcs, int k) {
  if (k < 1) return 0;
  SymDevHolder* h = f->h;
  const int kmax = std::min(f->max_rhs, 16);
  // balanced chunks of at most kmax columns (20 columns -> 10 + 10, not 16 + 4)
  const int nchunks = (k + kmax - 1) / kmax;
  const int per = (k + nchunks - 1) / nchunks;
  for (int c0 = 0; c0 < k; c0 += per) {
    const int kc = std::min(per, k - c0);
    SolveArgs a;
    a.tiles = h->solve.tiles;
    a.phases = h->solve.phases;
    a.nphases = h->solve.nphases;
    a.pull2 = h->solve.pull2;
    a.ovf = h->solve.ovf;
    a.perm = h->d.perm;
    a.sn_rows = h->d.sn_rows;
    a.sfwd = f->sfwd;
    a.sbwd = f->sbwd;
    a.dinv = f->dinv;
    a.wbuf = f->wbuf;
    a.ybuf = f->ybuf;
    a.xperm = f->xperm;
    a.B = B + (int64_t)c0 * bcs;
    a.brs = brs;
    a.bcs = bcs;
    a.X = X + (int64_t)c0 * xcs;
    a.xrs = xrs;
    a.xcs = xcs;
    a.k = kc;
    int rc;
    if (kc == 1) rc = launch_solve<1>(0, f, a);
    else if (kc == 2) rc = launch_solve<2>(1, f, a);
    else if (kc <= 4) rc = launch_solve<4>(2, f, a);
    else if (kc <= 8) rc = launch_solve<8>(3, f, a);
    else rc = launch_solve<16>(4, f, a);
    if (rc) return rc;
  }
  return 0;
}

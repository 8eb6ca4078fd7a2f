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
#include <vector>

namespace {

constexpr int MAX_PHASES_SMEM = 96;
constexpr int MAX_SUB_LEVELS = 32;
constexpr int REC_CAP = 512;          // subtree tile records staged in shared memory per slot (48 B each)

struct SolveArgs {
  const TileRec* tiles;
  const PhaseRec* phases;
  int nphases;
  const int* ovf_row;
  const int* ovf;
  const int* sub_ptr;
  const int* perm;
  const int* sn_rows;
  const int* rel;
  const double* sfwd;
  const double* sbwd;
  const double* dinv;
  double* wbuf;          // three slabs x kmax planes x sumf rows
  int64_t sumf;          // rows per plane
  int64_t slab_stride;   // kmax * sumf
  double* bperm;         // permuted right-hand side, n x k
  double* ybuf;
  double* xperm;
  const double* B;
  int64_t brs, bcs;
  double* X;
  int64_t xrs, xcs;
  int n, k;
  unsigned long long* barrier;
  unsigned long long bar_base;
  unsigned long long* times;      // developer profiling: %globaltimer of CTA 0 after every phase (NULL: off)
};

// Vectors produced earlier in the same launch by other SMs are read through L2 (ld.global.cg): L1 is not
// coherent across SMs and a line fetched in an earlier phase may be stale.
//
// Every operand loader below first ISSUES all its loads into registers and only then combines them: the
// per-tile critical path is one memory round trip, not one per operand (measured: `v += ld(..)` chains cost
// three to four exposed L2/DRAM latencies per tile, 1.5 - 4 us, which is what bounded the solve).
template <int KT>
struct RhsGroup {           // right-hand sides handled per batch of loads (register budget: 5 loads each)
  static constexpr int value = KT % 4 == 0 ? 4 : (KT % 2 == 0 ? 2 : 1);
};

// v[0:k) = (brow >= 0 ? bperm[brow, :] : 0) + the forward-sweep updates addressed to w-row t: the direct
// child slabs the front's children actually write (fixed order), then the overflow list
template <int KT>
__device__ __forceinline__ void fwd_operand(const SolveArgs& a, int64_t t, int64_t brow, int64_t link, double* v) {
  constexpr int G = RhsGroup<KT>::value;
  const int k = a.k;
  const int nsl = (int)((link >> LINK_NSLAB_SHIFT) & 7);
  const double* w0 = a.wbuf + t;
  const double* bp = a.bperm + (brow >= 0 ? brow : 0) * k;
#pragma unroll
  for (int r0 = 0; r0 < KT; r0 += G) {
    double b[G], s[SOLVE_NSLAB][G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int r = r0 + g;
      const bool ok = r < k;
      b[g] = (ok && brow >= 0) ? __ldcg(bp + r) : 0.0;
#pragma unroll
      for (int sl = 0; sl < SOLVE_NSLAB; ++sl)
        s[sl][g] = (ok && sl < nsl) ? __ldcg(w0 + sl * a.slab_stride + (int64_t)r * a.sumf) : 0.0;
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      double acc = b[g];
#pragma unroll
      for (int sl = 0; sl < SOLVE_NSLAB; ++sl) acc += s[sl][g];
      v[r0 + g] = acc;
    }
  }
  if (link & LINK_HAS_OVF) {
    const int o = __ldg(&a.ovf_row[t]);
    if (o >= 0) {
      const int cnt = __ldg(&a.ovf[o]);
      const double* p2 = a.wbuf + SOLVE_NSLAB * a.slab_stride;
#pragma unroll 1
      for (int q = 0; q < cnt; ++q) {
        const int64_t src = __ldg(&a.ovf[o + 1 + q]);
#pragma unroll
        for (int r = 0; r < KT; ++r)
          if (r < k) v[r] += __ldcg(p2 + (int64_t)r * a.sumf + src);
      }
    }
  }
}

// v[0:k) = base[row, :]
template <int KT>
__device__ __forceinline__ void load_row(const double* __restrict__ base, int64_t row, int k, double* v) {
  const double* p = base + row * k;
#pragma unroll
  for (int r = 0; r < KT; ++r) v[r] = r < k ? __ldcg(p + r) : 0.0;
}

__device__ __forceinline__ TileRec unpack_tile(int4 a, int4 b, int4 c) {
  TileRec t;
  t.first = a.x; t.nc = a.y; t.nb = a.z; t.tile = a.w;
  t.soff = ((int64_t)(unsigned)b.x) | ((int64_t)b.y << 32);
  t.w_off = ((int64_t)(unsigned)b.z) | ((int64_t)b.w << 32);
  t.row_off = ((int64_t)(unsigned)c.x) | ((int64_t)c.y << 32);
  t.link = ((int64_t)(unsigned)c.z) | ((int64_t)c.w << 32);
  return t;
}

__device__ __forceinline__ TileRec load_tile(const TileRec* p) {
  const int4* q = reinterpret_cast<const int4*>(p);
  return unpack_tile(__ldg(q), __ldg(q + 1), __ldg(q + 2));
}

// acc += M[out, c0:c1) * in[c0:c1) for this lane's output; M column-major with leading dimension ld.
// stage_fn(c, v) fills v[0:k) with input vector entry c (lane-parallel over the chunk of 32 columns),
// the chunk is then shared through the warp's staging buffer ([32 columns][KT], so that one column's KT
// values are one or a few 16-byte shared loads).
// The loop is written for a SHORT instruction stream: ncu showed the previous fully unrolled, predicated
// 32-slot body (~2500 SASS instructions per chunk, 64-bit multiplies for every address) made the solve
// issue-bound -- 16 warps x 2500 instructions per round of tiles is 5 us on an SM.  Here a column costs one
// pointer add, one predicated load and KT (LDS +) DFMA.  Memory-level parallelism still comes from the
// instruction stream: MB independent 8-byte panel loads per lane (256 B per warp each) are in flight while
// the previous batch is consumed, and the first batch is requested BEFORE the gather of the input vector.
template <int KT, class StageFn>
__device__ __forceinline__ void warp_panel_product(const double* __restrict__ M, int64_t ld, int c0, int c1, int lane,
                                                   double* stage, double* acc, StageFn stage_fn) {
  constexpr int MB = KT <= 2 ? 16 : 8;
  for (int cc = c0; cc < c1; cc += 32) {
    const int ncol = min(32, c1 - cc);
    const double* p = M + (int64_t)cc * ld;
    double m[MB];
#pragma unroll
    for (int j = 0; j < MB; ++j) {
      m[j] = j < ncol ? __ldg(p) : 0.0;
      p += ld;
    }
    double v[KT];
#pragma unroll
    for (int r = 0; r < KT; ++r) v[r] = 0.0;
    if (lane < ncol) stage_fn(cc + lane, v);
    if constexpr (KT % 2 == 0) {
      double2* s2 = reinterpret_cast<double2*>(stage + lane * KT);
#pragma unroll
      for (int r = 0; r < KT / 2; ++r) s2[r] = make_double2(v[2 * r], v[2 * r + 1]);
    } else {
#pragma unroll
      for (int r = 0; r < KT; ++r) stage[lane * KT + r] = v[r];
    }
    __syncwarp();
    for (int j0 = 0; j0 < ncol; j0 += MB) {
      double mn[MB];
      const int rem = ncol - j0 - MB;                      // columns left after this batch
      if (rem > 0) {
#pragma unroll
        for (int j = 0; j < MB; ++j) {
          mn[j] = j < rem ? __ldg(p) : 0.0;
          p += ld;
        }
      }
      const double* sj = stage + j0 * KT;
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        if constexpr (KT % 2 == 0) {
          const double2* s2 = reinterpret_cast<const double2*>(sj + j * KT);
#pragma unroll
          for (int r = 0; r < KT / 2; ++r) {
            const double2 t = s2[r];
            acc[2 * r] = fma(m[j], t.x, acc[2 * r]);
            acc[2 * r + 1] = fma(m[j], t.y, acc[2 * r + 1]);
          }
        } else {
#pragma unroll
          for (int r = 0; r < KT; ++r) acc[r] = fma(m[j], sj[j * KT + r], acc[r]);
        }
      }
      if (rem > 0) {
#pragma unroll
        for (int j = 0; j < MB; ++j) m[j] = mn[j];
      }
    }
    __syncwarp();
  }
}

// L2 prefetch of [p, p + bytes): a hint, so the 16-byte alignment the instruction asks for is met by widening
// the range (panel storage is padded to 256 B, the widened range stays inside the allocation)
__device__ __forceinline__ void prefetch_l2(const void* p, unsigned bytes) {
  const unsigned long long addr = (unsigned long long)p;
  const unsigned long long a0 = addr & ~15ull;
  const unsigned long long a1 = (addr + bytes + 15ull) & ~15ull;
  const unsigned sz = (unsigned)(a1 - a0);
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(sz) : "memory");
}

__device__ __forceinline__ void grid_barrier(unsigned long long* ctr, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long v;
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(ctr) : "memory");
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

// One warp tile: 32 consecutive outputs of one front, slice `slice` of `ws` of its reduction dimension.
// Forward (dir 0): outputs are front rows, acc = S[row, 0:cend) w1 (+ the children's updates of the row);
// backward (dir 1): outputs are pivot columns, acc = S^T[col, o0:f) [z1 ; x2].
// Both directions go through ONE instance of warp_panel_product (the direction is a run-time branch in the
// gather), and both kinds of phase through one call site: the kernel's code footprint and register
// pressure, not DRAM, were what bounded the previous version.
// aux: what the store of this lane's output needs (forward pivot row: D^-1 entry; forward update row: its
// position in the parent front; backward: the original index of the column), requested here, at the top of
// the tile, so that the store does not start with a dependent look-up.
// panel slice of a tile: M = first output row (forward) / pivot column (backward) of the tile in the column-major
// panel with leading dimension ld, [c0, c1) = the part of the reduction dimension slice `slice` of `ws` covers
__device__ __forceinline__ void tile_range(const SolveArgs& a, int dir, const TileRec& tr, int slice, int ws,
                                           const double*& M, int& ld, int& c0, int& c1) {
  const int nc = tr.nc, f = tr.nc + tr.nb;
  const int o0 = tr.tile * SOLVE_TILE;
  if (dir == 0) {
    const int cend = min(nc, o0 + SOLVE_TILE);          // S is lower triangular inside the pivot rows
    const int per = (cend + ws - 1) / ws;
    c0 = slice * per;
    c1 = min(cend, c0 + per);
    M = a.sfwd + tr.soff + o0;
    ld = f;
  } else {
    const int len = f - o0;
    const int per = (len + ws - 1) / ws;
    c0 = o0 + slice * per;
    c1 = min(f, c0 + per);
    M = a.sbwd + tr.soff + o0;
    ld = nc;
  }
}

struct TileAux {
  double d;
  int i;
};

template <int KT>
__device__ __forceinline__ TileAux tile_compute(const SolveArgs& a, int dir, const TileRec& tr, int lane, int slice, int ws,
                                                double* stage, double* acc) {
  const int k = a.k;
  const int nc = tr.nc, f = tr.nc + tr.nb;
  const int o0 = tr.tile * SOLVE_TILE;
  const int out = o0 + lane;
  const double* M;
  int ld, c0, c1;
  TileAux aux;
  aux.d = 0.0;
  aux.i = 0;
  const bool upd = dir == 0 && slice == 0 && out >= nc && out < f;      // forward update row owned by this lane
  if (slice == 0) {
    if (dir == 0) {
      if (out < nc) aux.d = __ldg(&a.dinv[tr.first + out]);
      else if (out < f) aux.i = __ldg(&a.rel[tr.row_off + out - nc]);
    } else if (out < nc) {
      aux.i = __ldg(&a.perm[tr.first + out]);
    }
  }
  double u[KT];
#pragma unroll
  for (int r = 0; r < KT; ++r) u[r] = 0.0;
  if (KT <= 4 && upd && (tr.link & LINK_HAS_CHILDREN)) fwd_operand<KT>(a, tr.w_off + out, -1, tr.link, u);
  tile_range(a, dir, tr, slice, ws, M, ld, c0, c1);
  M += min(lane, (dir == 0 ? f : nc) - 1 - o0);
  warp_panel_product<KT>(M, ld, c0, c1, lane, stage, acc, [&](int c, double* v) {
    if (dir == 0) fwd_operand<KT>(a, tr.w_off + c, tr.first + c, tr.link, v);
    else if (c < nc) load_row<KT>(a.ybuf, tr.first + c, k, v);
    else load_row<KT>(a.xperm, __ldg(&a.sn_rows[tr.row_off + c - nc]), k, v);
  });
  if (KT > 4 && upd && (tr.link & LINK_HAS_CHILDREN)) fwd_operand<KT>(a, tr.w_off + out, -1, tr.link, u);
#pragma unroll
  for (int r = 0; r < KT; ++r) acc[r] += u[r];
  return aux;
}

template <int KT>
__device__ __forceinline__ void tile_store(const SolveArgs& a, int dir, const TileRec& tr, int lane, const TileAux& aux,
                                           const double* acc) {
  const int k = a.k;
  const int nc = tr.nc, f = tr.nc + tr.nb;
  const int out = tr.tile * SOLVE_TILE + lane;
  if (dir == 0) {
    if (out < nc) {
      double* y = a.ybuf + (int64_t)(tr.first + out) * k;
#pragma unroll
      for (int r = 0; r < KT; ++r)
        if (r < k) y[r] = aux.d * acc[r];
    } else if (out < f) {
      const int slab = (int)((tr.link >> LINK_SLAB_SHIFT) & 0xff);
      int64_t dst;                                            // row in the plane of the destination slab
      if (slab < SOLVE_NSLAB) dst = slab * a.slab_stride + (tr.link & LINK_WOFF_MASK) + aux.i;
      else dst = SOLVE_NSLAB * a.slab_stride + tr.w_off + out;
      double* w = a.wbuf + dst;
#pragma unroll
      for (int r = 0; r < KT; ++r)
        if (r < k) w[(int64_t)r * a.sumf] = acc[r];
    }
  } else if (out < nc) {
    double* xp = a.xperm + (int64_t)(tr.first + out) * k;
    double* xo = a.X + (int64_t)aux.i * a.xrs;
#pragma unroll
    for (int r = 0; r < KT; ++r)
      if (r < k) { xp[r] = acc[r]; xo[(int64_t)r * a.xcs] = acc[r]; }
  }
}

template <int KT>
__global__ void __launch_bounds__(SOLVE_WARPS * 32, 1) solve_kernel(const __grid_constant__ SolveArgs a) {
  extern __shared__ double smem[];
  __shared__ PhaseRec s_phase[MAX_PHASES_SMEM];
  __shared__ int s_tab[MAX_SUB_LEVELS];                    // level table of the slot of a subtree phase
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* stage = smem + warp * (32 * KT);                 // [32][KT] per warp
  double* part = smem + SOLVE_WARPS * 32 * KT;             // [SOLVE_WARPS][KT][32] partial sums
  int4* s_rec = reinterpret_cast<int4*>(smem + 2 * SOLVE_WARPS * 32 * KT);   // tile records, REC_CAP x 48 B
  int4* s_first = s_rec + 3 * REC_CAP;                     // per warp: first tile record of the next level phase
  unsigned long long target = a.bar_base;
  if (a.times && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    a.times[0] = t;
  }
  {
    const int4* src = reinterpret_cast<const int4*>(a.phases);
    int4* dst = reinterpret_cast<int4*>(s_phase);
    for (int e = threadIdx.x; e < 2 * min(a.nphases, MAX_PHASES_SMEM); e += blockDim.x) dst[e] = __ldg(src + e);
  }
  // ---- pre-phase: permuted copy of the right-hand side (coalesced writes, gathered reads), then a grid barrier
  {
    const int k = a.k;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)a.n * k; e += (int64_t)gridDim.x * blockDim.x) {
      const int64_t i = e / k;
      const int r = (int)(e - i * k);
      a.bperm[e] = a.B[(int64_t)__ldg(&a.perm[i]) * a.brs + (int64_t)r * a.bcs];
    }
    target += gridDim.x;
    grid_barrier(a.barrier, target);
  }
  bool have_first = false;                                 // s_first[warp] holds this warp's first tile of phase p
  for (int p = 0; p < a.nphases; ++p) {
    const PhaseRec ph = p < MAX_PHASES_SMEM ? s_phase[p] : a.phases[p];
    const int dir = ph.dir;
    const bool subtree = ph.ws == 0;
    const int ws = subtree ? 1 : ph.ws;
    // request this warp's first tile record of the NEXT phase now (static schedule): the load completes
    // behind this phase's work instead of in front of the next phase's
    bool nxt_have = false;
    int4 nxt = make_int4(0, 0, 0, 0);
    if (p + 1 < a.nphases) {
      const PhaseRec nx = (p + 1) < MAX_PHASES_SMEM ? s_phase[p + 1] : a.phases[p + 1];
      if (nx.ws > 0) {
        const int te = (int)blockIdx.x * (SOLVE_WARPS / nx.ws) + warp / nx.ws;
        if (te < nx.ntiles) {
          nxt_have = true;
          if (lane < 3) nxt = __ldg(reinterpret_cast<const int4*>(a.tiles + nx.tile_off + te) + lane);
        }
      }
    }
    // A phase is a sequence of steps; a step is a strided range of tile records [base, end) handled by this
    // warp group.  Level phase: one step, tiles spread over the grid.  Subtree phase: one step per local
    // level of the CTA's own subtrees (CTA barrier between steps), tile records staged in shared memory.
    int nsteps = 1, nrec = 0, tb = 0;
    const int* tab = nullptr;
    if (subtree) {
      nsteps = (int)blockIdx.x < ph.level ? ph.ntiles : 0;   // slots beyond the grid do not exist (nslots == grid)
      tab = a.sub_ptr + ph.tile_off + (int64_t)blockIdx.x * (ph.ntiles + 1);
      if (nsteps) {
        tb = __ldg(&tab[0]);
        nrec = min(__ldg(&tab[nsteps]) - tb, REC_CAP);
        if (threadIdx.x <= nsteps && threadIdx.x < MAX_SUB_LEVELS) s_tab[threadIdx.x] = __ldg(&tab[threadIdx.x]);
        const int4* src = reinterpret_cast<const int4*>(a.tiles + tb);
        for (int e = threadIdx.x; e < 3 * nrec; e += blockDim.x) s_rec[e] = __ldg(src + e);
      }
      __syncthreads();
      // pull the panels of the slot's fronts into L2 now (one bulk prefetch per front, ~3 - 10 KB each): the
      // tiles' own loads then see L2 latency, and DRAM streams at full rate instead of in per-tile bursts
      {
        const double* pan = dir == 0 ? a.sfwd : a.sbwd;
        for (int q = threadIdx.x; q < nrec; q += blockDim.x) {
          const int4 r0 = s_rec[3 * q];
          if (r0.w == 0) {                                   // first tile of its front
            const int4 r1 = s_rec[3 * q + 1];
            const int64_t soff = ((int64_t)(unsigned)r1.x) | ((int64_t)r1.y << 32);
            prefetch_l2(pan + soff, (unsigned)((r0.y + r0.z) * r0.y) * 8u);
          }
        }
      }
    }
    for (int step = 0; step < nsteps; ++step) {
      int base, stride, end, niter;
      if (subtree) {
        const int l = dir == 0 ? step : nsteps - 1 - step;
        const int t0 = l + 1 < MAX_SUB_LEVELS ? s_tab[l] : __ldg(&tab[l]);
        end = l + 1 < MAX_SUB_LEVELS ? s_tab[l + 1] : __ldg(&tab[l + 1]);
        base = t0 + warp;
        stride = SOLVE_WARPS;
        niter = (end - t0 + SOLVE_WARPS - 1) / SOLVE_WARPS;
      } else {
        const int tpc = SOLVE_WARPS / ws;
        const int nct = (ph.ntiles + tpc - 1) / tpc;
        base = (int)ph.tile_off + (int)blockIdx.x * tpc + warp / ws;
        stride = (int)gridDim.x * tpc;
        end = (int)ph.tile_off + ph.ntiles;
        niter = nct > (int)blockIdx.x ? (nct - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
      }
      const int slice = subtree ? 0 : warp % ws;
      unsigned long long* dbg = nullptr;
      if (a.times && blockIdx.x == 0 && warp == 0 && subtree)
        dbg = a.times + a.nphases + 1 + ((dir * MAX_SUB_LEVELS) + step) * 8;
      if (dbg && lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        dbg[0] = t;
        dbg[1] = niter;
      }
      for (int it = 0; it < niter; ++it) {
        long long c0 = 0, c1 = 0;
        if (dbg) c0 = clock64();
        const int te = base + it * stride;
        const bool have = te < end;
        double acc[KT];
#pragma unroll
        for (int r = 0; r < KT; ++r) acc[r] = 0.0;
        TileRec tr;
        tr.first = tr.nc = tr.nb = tr.tile = 0;
        tr.soff = tr.w_off = tr.row_off = tr.link = 0;
        TileAux aux;
        aux.d = 0.0;
        aux.i = 0;
        if (have) {
          const int q = te - tb;
          if (subtree && q < nrec) tr = unpack_tile(s_rec[3 * q], s_rec[3 * q + 1], s_rec[3 * q + 2]);
          else if (!subtree && it == 0 && have_first) tr = unpack_tile(s_first[3 * warp], s_first[3 * warp + 1], s_first[3 * warp + 2]);
          else tr = load_tile(a.tiles + te);
          aux = tile_compute<KT>(a, dir, tr, lane, slice, ws, stage, acc);
        }
        if (dbg) c1 = clock64();
        if (ws > 1) {
#pragma unroll
          for (int r = 0; r < KT; ++r) part[(warp * KT + r) * 32 + lane] = acc[r];
          __syncthreads();
          if (have && slice == 0)
            for (int s = 1; s < ws; ++s)
#pragma unroll
              for (int r = 0; r < KT; ++r) acc[r] += part[((warp + s) * KT + r) * 32 + lane];
        }
        if (have && slice == 0) tile_store<KT>(a, dir, tr, lane, aux, acc);
        if (ws > 1) __syncthreads();
        if (dbg && lane == 0 && it < 3) {
          dbg[2 + 2 * it] = (unsigned long long)(c1 - c0);
          dbg[3 + 2 * it] = (unsigned long long)(clock64() - c1);
        }
      }
      if (subtree) __syncthreads();     // CTA-scope ordering: the level's results are visible to the whole slot
    }
    // ---- hand the prefetched first tile record of the next phase to the whole warp
    have_first = nxt_have;
    if (nxt_have && lane < 3) s_first[3 * warp + lane] = nxt;
    __syncwarp();
    if (nxt_have) {
      // ... and pull that tile's panel slice into L2 behind the barrier: after it, the tile's loads see L2 latency
      const PhaseRec nx = (p + 1) < MAX_PHASES_SMEM ? s_phase[p + 1] : a.phases[p + 1];
      const TileRec nt = unpack_tile(s_first[3 * warp], s_first[3 * warp + 1], s_first[3 * warp + 2]);
      const double* M;
      int ld, c0, c1;
      tile_range(a, nx.dir, nt, warp % nx.ws, nx.ws, M, ld, c0, c1);
      const int rows = min(SOLVE_TILE, (nx.dir == 0 ? nt.nc + nt.nb : nt.nc) - nt.tile * SOLVE_TILE);
      for (int c = c0 + lane; c < c1; c += 32) prefetch_l2(M + (int64_t)c * ld, (unsigned)rows * 8u);
    }
    target += gridDim.x;
    if (p + 1 < a.nphases) grid_barrier(a.barrier, target);
    if (a.times && blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      a.times[p + 1] = t;
    }
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
KernelCfg g_cfg[8];
int g_num_sms = 0;

template <int KT>
int configure(int slot) {
  KernelCfg& c = g_cfg[slot];
  if (c.ready) return 0;
  c.smem = (size_t)SOLVE_WARPS * 32 * KT * 8 * 2 + (size_t)(REC_CAP + SOLVE_WARPS) * 48;
  EIGD_CUDA(cudaFuncSetAttribute(solve_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
  int occ = 0;
  EIGD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, solve_kernel<KT>, SOLVE_WARPS * 32, c.smem));
  if (occ < 1) { eigd_set_error("solve: kernel does not fit on an SM"); return 5; }
  c.grid = g_num_sms;
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
  f->bar_base += (unsigned long long)a.nphases * (unsigned long long)c.grid;   // pre-phase + (nphases - 1) barriers
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
  build_solve_plan_host(S, g_num_sms * SOLVE_WARPS, g_num_sms, -2, P);
  SolvePlanDev& d = h->solve;
  int rc = 0;
  rc |= upload_vec(h, P.tiles, &d.tiles);
  rc |= upload_vec(h, P.phases, &d.phases);
  rc |= upload_vec(h, P.ovf_row, &d.ovf_row);
  rc |= upload_vec(h, P.ovf, &d.ovf);
  rc |= upload_vec(h, P.sub_ptr, &d.sub_ptr);
  d.nphases = (int)P.phases.size();
  d.host_phases = P.phases;
  return rc;
}

void free_solve_plan_dev(SymDevHolder*) {}   // the arrays are owned by SymDevHolder::allocs

// ---- live timing of every solve launch (bench.py roofline): CUDA events on the launching stream ----------
struct SolveTiming {
  cudaEvent_t e0, e1;
  int k;
};
static bool g_timing_on = false;
static std::vector<SolveTiming> g_timings;

extern "C" int eigd_solve_timing_begin(void) {
  for (auto& t : g_timings) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
  g_timings.clear();
  g_timing_on = true;
  return 0;
}

// calls_by_k / ms_by_k: 33 entries each (index = number of right-hand sides of the launch, 1 .. 32)
extern "C" int eigd_solve_timing_end(int64_t* calls_by_k, double* ms_by_k) {
  g_timing_on = false;
  for (int i = 0; i < 33; ++i) { calls_by_k[i] = 0; ms_by_k[i] = 0.0; }
  for (auto& t : g_timings) {
    EIGD_CUDA(cudaEventSynchronize(t.e1));
    float ms = 0.f;
    EIGD_CUDA(cudaEventElapsedTime(&ms, t.e0, t.e1));
    int k = t.k < 32 ? t.k : 32;
    calls_by_k[k] += 1;
    ms_by_k[k] += ms;
    cudaEventDestroy(t.e0);
    cudaEventDestroy(t.e1);
  }
  g_timings.clear();
  return 0;
}

// developer profiling hook: device buffer of (nphases + 1) u64 receiving per-phase timestamps of the next solves
static unsigned long long* g_phase_times = nullptr;
extern "C" int eigd_solve_set_phase_times(void* d_buf) { g_phase_times = (unsigned long long*)d_buf; return 0; }
extern "C" int eigd_solve_num_phases(const eigd_factor* f) { return f->h->solve.nphases; }

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
    a.ovf_row = h->solve.ovf_row;
    a.ovf = h->solve.ovf;
    a.sub_ptr = h->solve.sub_ptr;
    a.perm = h->d.perm;
    a.sn_rows = h->d.sn_rows;
    a.rel = h->d.rel;
    a.sumf = f->sym->w_off[f->sym->nsuper];
    a.slab_stride = a.sumf * kmax;
    a.bperm = f->bperm;
    a.n = f->sym->n;
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
    a.times = g_phase_times;
    int rc;
    SolveTiming tm;
    if (g_timing_on) {
      tm.k = kc;
      EIGD_CUDA(cudaEventCreate(&tm.e0));
      EIGD_CUDA(cudaEventCreate(&tm.e1));
      EIGD_CUDA(cudaEventRecord(tm.e0, g_eigd_stream));
    }
    if (kc == 1) rc = launch_solve<1>(0, f, a);
    else if (kc == 2) rc = launch_solve<2>(1, f, a);
    else if (kc <= 4) rc = launch_solve<4>(2, f, a);
    else if (kc <= 6) rc = launch_solve<6>(3, f, a);
    else if (kc <= 8) rc = launch_solve<8>(4, f, a);
    else if (kc <= 10) rc = launch_solve<10>(5, f, a);
    else if (kc <= 12) rc = launch_solve<12>(6, f, a);
    else rc = launch_solve<16>(7, f, a);
    if (rc) return rc;
    if (g_timing_on) {
      EIGD_CUDA(cudaEventRecord(tm.e1, g_eigd_stream));
      g_timings.push_back(tm);
    }
  }
  return 0;
}

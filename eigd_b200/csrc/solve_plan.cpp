// Host side of the triangular-solve plan: tile records, per-level phases, gather lists.
// Pure C++ (no CUDA) so that the plan can be checked on a machine without a GPU
// (tests/test_solve_plan.py emulates the kernel on these arrays).
#include "solve_plan.hpp"

#include <algorithm>

namespace {

int pick_ws(int ntiles, int redmax, int target_warps) {
  // split the reduction dimension of a tile over ws warps while (i) the level would otherwise leave
  // most resident warps idle and (ii) every warp still gets >= 8 terms
  int ws = 1;
  while (ws < SOLVE_WARPS && (int64_t)ntiles * ws * 2 <= target_warps && redmax / (2 * ws) >= 8) ws *= 2;
  return ws;
}

}  // namespace

void build_solve_plan_host(const eigd_symbolic* S, int target_warps, SolvePlanHost& P) {
  const int ns = S->nsuper;
  P.soff.assign(ns + 1, 0);
  for (int k = 0; k < ns; ++k) P.soff[k + 1] = P.soff[k] + (int64_t)sn_fsize(S, k) * sn_ncols(S, k);

  // ---- gather lists -------------------------------------------------------------------------
  const int64_t sumf = S->w_off[ns];
  std::vector<std::vector<int>> extra;          // overflow sources, rare (more than two children share a row)
  std::vector<int> cnt(sumf, 0);
  P.pull2.assign(2 * sumf, -1);
  std::vector<int64_t> extra_of(sumf, -1);
  for (int p = 0; p < ns; ++p)
    for (int q = S->child_ptr[p]; q < S->child_ptr[p + 1]; ++q) {   // children in ascending order: fixed sum order
      int c = S->child_idx[q];
      int ncc = sn_ncols(S, c), nbc = sn_nbelow(S, c);
      for (int i = 0; i < nbc; ++i) {
        int64_t t = S->w_off[p] + S->rel[S->sn_rowptr[c] + i];
        int src = (int)(S->w_off[c] + ncc + i);
        if (cnt[t] < 2) P.pull2[2 * t + cnt[t]] = src;
        else {
          if (extra_of[t] < 0) { extra_of[t] = (int64_t)extra.size(); extra.emplace_back(); }
          extra[extra_of[t]].push_back(src);
        }
        cnt[t]++;
      }
    }
  P.ovf.clear();
  for (int64_t t = 0; t < sumf; ++t)
    if (extra_of[t] >= 0) {
      // keep source 0 in place, move source 1 to the head of the overflow list
      std::vector<int>& e = extra[extra_of[t]];
      int o = (int)P.ovf.size();
      P.ovf.push_back((int)e.size() + 1);
      P.ovf.push_back(P.pull2[2 * t + 1]);
      P.ovf.insert(P.ovf.end(), e.begin(), e.end());
      P.pull2[2 * t + 1] = -2 - o;
    }

  // ---- tiles and phases ---------------------------------------------------------------------
  P.tiles.clear();
  P.phases.clear();
  auto rec = [&](int k, int tile) {
    TileRec r;
    r.first = S->sn_first[k];
    r.nc = sn_ncols(S, k);
    r.nb = sn_nbelow(S, k);
    r.tile = tile;
    r.soff = P.soff[k];
    r.w_off = S->w_off[k];
    r.row_off = S->sn_rowptr[k];
    r.pad = 0;
    return r;
  };
  for (int l = 0; l < S->nlevels; ++l) {                 // forward: outputs are the f rows of each front
    PhaseRec ph{0, 1, 0, l, (int64_t)P.tiles.size(), 0};
    int redmax = 1;
    for (int q = S->level_ptr[l]; q < S->level_ptr[l + 1]; ++q) {
      int k = S->level_sn[q], f = sn_fsize(S, k);
      redmax = std::max(redmax, sn_ncols(S, k));
      for (int t = 0; t * SOLVE_TILE < f; ++t) P.tiles.push_back(rec(k, t));
    }
    ph.ntiles = (int)((int64_t)P.tiles.size() - ph.tile_off);
    ph.ws = pick_ws(ph.ntiles, redmax, target_warps);
    P.phases.push_back(ph);
  }
  P.nfwd = (int)P.phases.size();
  for (int l = S->nlevels - 1; l >= 0; --l) {            // backward: outputs are the nc pivot columns
    PhaseRec ph{1, 1, 0, l, (int64_t)P.tiles.size(), 0};
    int redmax = 1;
    for (int q = S->level_ptr[l]; q < S->level_ptr[l + 1]; ++q) {
      int k = S->level_sn[q], nc = sn_ncols(S, k);
      redmax = std::max(redmax, sn_fsize(S, k));
      for (int t = 0; t * SOLVE_TILE < nc; ++t) P.tiles.push_back(rec(k, t));
    }
    ph.ntiles = (int)((int64_t)P.tiles.size() - ph.tile_off);
    ph.ws = pick_ws(ph.ntiles, redmax, target_warps);
    P.phases.push_back(ph);
  }
}

// ---- inspection entry point (tests; not used by the product path) ----------------------------
template <class T>
static int64_t copy_out64(const std::vector<T>& v, int64_t* out, int64_t cap) {
  int64_t m = std::min<int64_t>((int64_t)v.size(), cap);
  if (out) for (int64_t i = 0; i < m; ++i) out[i] = (int64_t)v[i];
  return (int64_t)v.size();
}

extern "C" int64_t eigd_solve_plan_get(const eigd_symbolic* s, int target_warps, int which, int64_t* out, int64_t cap) {
  SolvePlanHost P;
  build_solve_plan_host(s, target_warps, P);
  switch (which) {
    case 0: return copy_out64(P.soff, out, cap);
    case 1: return copy_out64(P.pull2, out, cap);
    case 2: return copy_out64(P.ovf, out, cap);
    case 3: {
      std::vector<int64_t> flat;
      flat.reserve(P.tiles.size() * 7);
      for (const TileRec& r : P.tiles) {
        int64_t a[7] = {r.first, r.nc, r.nb, r.tile, r.soff, r.w_off, r.row_off};
        flat.insert(flat.end(), a, a + 7);
      }
      return copy_out64(flat, out, cap);
    }
    case 4: {
      std::vector<int64_t> flat;
      for (const PhaseRec& p : P.phases) {
        int64_t a[5] = {p.dir, p.ws, p.ntiles, p.level, p.tile_off};
        flat.insert(flat.end(), a, a + 5);
      }
      return copy_out64(flat, out, cap);
    }
    default: return -1;
  }
}

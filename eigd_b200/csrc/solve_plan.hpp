// Host-built execution plan of the persistent triangular-solve kernel (solve.cu).
// No reference counterpart: the reference calls SuperLU.solve once per right-hand side
// (eigd/eigenvector_derivatives.py:18-23).  See DESIGN.md "Triangular solves".
#pragma once
#include <cstdint>
#include <vector>

#include "symbolic.hpp"

constexpr int SOLVE_WARPS = 16;   // warps per CTA of the solve kernel
constexpr int SOLVE_TILE = 32;    // outputs (front rows / pivot columns) per warp tile

// One warp tile: 32 consecutive outputs of one front.  Everything the warp needs about the front
// travels in this one 48-byte record, so the dependent-load chain per tile is
// record -> {perm, pull2, rows} -> {b, w, x} instead of five chained index look-ups.
struct TileRec {
  int first, nc, nb, tile;        // first pivot column (permuted), #pivot columns, #rows below, tile index
  int64_t soff;                   // offset of the front's f x nc solve panel (same in S and S^T storage)
  int64_t w_off;                  // offset of the front in the w-row space (prefix sum of front sizes)
  int64_t row_off;                // offset of the front's below-row list in sn_rows
  int64_t pad;
};
static_assert(sizeof(TileRec) == 48, "TileRec is read as three 16-byte words");

// One phase = the tiles of one level of the assembly tree in one direction; phases are separated
// by a grid-wide barrier.  ws warps share a tile (they split its reduction dimension).
struct PhaseRec {
  int dir, ws, ntiles, level;     // dir 0 forward, 1 backward
  int64_t tile_off;
  int64_t pad;
};
static_assert(sizeof(PhaseRec) == 32, "PhaseRec layout");

struct SolvePlanHost {
  std::vector<int64_t> soff;      // nsuper + 1, prefix sum of f * nc
  // deterministic gather form of the multifrontal extend-add of the forward sweep: row t of front p
  // receives w[pull2[2t]] + w[pull2[2t+1]] (-1 = none); if pull2[2t+1] <= -2 the remaining sources
  // are ovf[o+1 .. o+1+ovf[o]) with o = -2 - pull2[2t+1]
  std::vector<int> pull2;         // 2 * sum_front
  std::vector<int> ovf;
  std::vector<TileRec> tiles;
  std::vector<PhaseRec> phases;   // forward phases (leaves -> root) then backward phases (root -> leaves)
  int nfwd = 0;
};

void build_solve_plan_host(const eigd_symbolic* S, int target_warps, SolvePlanHost& P);

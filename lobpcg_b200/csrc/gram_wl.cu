// lobpcg_b200/csrc/gram_wl.cu — K2/K3 f64 Gram, work-list ("stream-K") version.
//
//   G (ma x mb) = A^H B,  A: n x ma, B: n x mb, column-major, n >> ma, mb          (reference: syrk/herk and
//   gemm_tn at src/gram/gram_impl.inc:54-63,92-101, src/rayleigh/rayleigh_ritz_modified_impl.inc:75-77,193-195)
//
// The first Gram kernel (dense.cu: gram_dmma_kernel) runs its main loop at ~85 % of the DMMA issue rate, but on the
// solver's shapes a third of the DMMAs it issues are wasted: (i) diagonal tiles of a Hermitian product compute the
// full 128x128 square, (ii) the last tile row/column is padded to the tile size (m = 900 -> 1024), (iii) the grid is
// tiles x equal n-splits, which rarely fills 148 SMs.  This version removes all three:
//   * the unit of MMA work is the 8x8 output block; every warp owns an (8 x 4)-block window of its CTA's tile and a
//     32-bit mask of the blocks it really computes.  Diagonal tiles use a hand-balanced cover of the upper triangle
//     (136 of 256 blocks, at most 18 per warp instead of 32); ragged edge tiles use a warp grid chosen for their
//     block extents.  A window may be "transposed" (the B panel feeds the MMA's row operand), which gives 4 x 8
//     windows with the same accumulator registers — both operand fragments have the same shared-memory layout.
//   * work = (tile, row range) items; tiles are laid end to end, weighted by their cost, and the 148 CTAs cut that
//     line into equal pieces (a CTA may finish one tile and start the next).  Every item writes its partial tile to
//     its own scratch slot; gram_wl_reduce_kernel sums the slots of a tile in item order — deterministic, no atomics.
// The schedule (items, warp layouts) is built on the host once per shape and cached on the context.
#include <algorithm>
#include <cstring>
#include <cmath>
#include <map>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "context.h"
#include "kernels.h"
#include "tile_loader.cuh"

namespace lb2 {

namespace {

constexpr int WL_T = 128;       // tile edge (columns of A / B staged per CTA)
constexpr int WL_NT = 256;      // threads per CTA (8 warps)
constexpr int WL_BLK = WL_T / 8;

struct WlWarp {                 // per-warp window of a tile layout
  uint32_t mask;                // bit i*4+j: compute block (a-operand block a0+i, b-operand block b0+j)
  uint8_t a0, b0;               // first 8-column block of the MMA row operand / column operand inside its panel
  uint8_t transposed;           // 1: row operand comes from the B panel (G columns), column operand from the A panel
  uint8_t shape;                // compile-time window shape id of `mask` (wl_stage_dispatch)
};
struct WlItem {
  int64_t r_begin, r_end;       // rows of this piece
  int32_t a_col0, a_cols;       // A panel: columns [a_col0, a_col0 + a_cols)
  int32_t b_col0, b_cols;
  int32_t layout;               // index of the first of 8 WlWarp entries
  int32_t same_panel;           // 1: A and B panels are the same columns of the same matrix -> staged once
  int32_t c_start;              // first K chunk to process; the piece is walked cyclically from there (plan_schedule: phase)
  int32_t b_sel;                // which column operand the B panel comes from (0: B, 1: B1; gram_wl_cols_f64)
};

// Window shapes.  Per-DMMA predicates cost more than the DMMAs they skip (ptxas wraps every predicated mma.sync in
// WARPSYNC), so the block mask of a warp is a COMPILE-TIME constant: 32 rectangles (na x nb blocks, na <= 8, nb <= 4)
// and the two staircase patterns of the balanced diagonal cover; the kernel switches on a warp-uniform shape id once
// per K chunk.  The host planner only emits masks from this list.
__host__ __device__ constexpr uint32_t wl_rect_mask(int na, int nb) {
  uint32_t m = 0;
  for (int i = 0; i < na; i++)
    for (int j = 0; j < nb; j++) m |= 1u << (i * 4 + j);
  return m;
}
// staircase A (diag16 warps 0 and 6, transposed window): i = G column block 0..5, j = G row block 0..3, j <= i
__host__ __device__ constexpr uint32_t wl_tri_a_mask() {
  uint32_t m = 0;
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 4; j++)
      if (j <= i) m |= 1u << (i * 4 + j);
  return m;
}
// staircase B (diag16 warps 1 and 7): i = G row block 0..7, j -> G column block 4 + j; upper part minus the
// (rows 0..3) x (columns 4..5) corner that belongs to staircase A
__host__ __device__ constexpr uint32_t wl_tri_b_mask() {
  uint32_t m = 0;
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 4; j++)
      if (i <= 4 + j && !(i <= 3 && j <= 1)) m |= 1u << (i * 4 + j);
  return m;
}
constexpr int WL_SHAPE_TRI_A = 32, WL_SHAPE_TRI_B = 33, WL_SHAPE_NONE = 34;

template <uint32_t MASK, int KS0, int KS1, int LDS>
__device__ __forceinline__ void wl_stage(double (&acc)[8][4][2], const double* __restrict__ as,
                                         const double* __restrict__ bs) {
#pragma unroll
  for (int ks = KS0; ks < KS1; ks++) {
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++)
      if ((MASK >> (4 * i)) & 0xFu) a[i] = as[i * 8 * LDS + ks * 4];
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (MASK & (0x11111111u << j)) b[j] = bs[j * 8 * LDS + ks * 4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 4; j++)
        if ((MASK >> (i * 4 + j)) & 1u) dmma884(acc[i][j], a[i], b[j]);
  }
}

template <int KS0, int KS1, int LDS>
__device__ __forceinline__ void wl_stage_dispatch(int shape, double (&acc)[8][4][2], const double* __restrict__ as,
                                                  const double* __restrict__ bs) {
#define WL_CASE(NA, NB) \
  case ((NA - 1) * 4 + (NB - 1)): wl_stage<wl_rect_mask(NA, NB), KS0, KS1, LDS>(acc, as, bs); break;
#define WL_ROW(NA) WL_CASE(NA, 1) WL_CASE(NA, 2) WL_CASE(NA, 3) WL_CASE(NA, 4)
  switch (shape) {
    WL_ROW(8) WL_ROW(7) WL_ROW(6) WL_ROW(5) WL_ROW(4) WL_ROW(3) WL_ROW(2) WL_ROW(1)
    case WL_SHAPE_TRI_A: wl_stage<wl_tri_a_mask(), KS0, KS1, LDS>(acc, as, bs); break;
    case WL_SHAPE_TRI_B: wl_stage<wl_tri_b_mask(), KS0, KS1, LDS>(acc, as, bs); break;
    default: break;
  }
#undef WL_ROW
#undef WL_CASE
}

template <int BK, int STAGES, bool VEC>
__global__ void __launch_bounds__(WL_NT, 1)
    gram_wl_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B0, int64_t ldb0,
                   const double* __restrict__ B1, int64_t ldb1, const double* __restrict__ B2, int64_t ldb2,
                   const WlItem* __restrict__ items, const int* __restrict__ cta_first, const WlWarp* __restrict__ layouts,
                   double* __restrict__ part, int split) {
  constexpr int LDS = BK + 4;
  extern __shared__ __align__(16) double smem_wl[];
  double* As = smem_wl;
  double* Bs = smem_wl + (size_t)STAGES * WL_T * LDS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int it_end = cta_first[blockIdx.x + 1];
  for (int it = cta_first[blockIdx.x]; it < it_end; ++it) {
    const WlItem item = items[it];
    const WlWarp cfg = layouts[item.layout + warp];
    const uint32_t mask = cfg.mask;
    const int shape = cfg.shape;
    const int64_t rows = item.r_end - item.r_begin;
    const int nchunks = (int)((rows + BK - 1) / BK);

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    const bool same = item.same_panel != 0;   // diagonal tile of S^H S: one panel feeds both operands
    const double* __restrict__ B = item.b_sel == 2 ? B2 : (item.b_sel ? B1 : B0);
    const int64_t ldb = item.b_sel == 2 ? ldb2 : (item.b_sel ? ldb1 : ldb0);
    TileLoaderF64<WL_T, BK, LDS, WL_NT, VEC> la, lb;
    int ci = item.c_start;   // chunk index of the next copy; chunks are walked c_start .. nchunks-1, 0 .. c_start-1
    la.init(A, lda, item.r_begin + (int64_t)ci * BK, item.a_col0, item.a_col0 + item.a_cols, tid);
    lb.init(B, ldb, item.r_begin + (int64_t)ci * BK, item.b_col0, item.b_col0 + item.b_cols, tid);
    int issued = 0, wstage = 0;
    auto issue = [&]() {
      if (issued < nchunks) {
        const int64_t valid = rows - (int64_t)ci * BK;
        const bool wrap = (ci + 1 == nchunks);
        const int64_t step = wrap ? -(int64_t)(nchunks - 1) * BK : (int64_t)BK;
        ci = wrap ? 0 : ci + 1;
        la.issue(As + wstage * (WL_T * LDS), A, valid);
        la.advance(step);
        if (!same) {
          lb.issue(Bs + wstage * (WL_T * LDS), B, valid);
          lb.advance(step);
        }
      }
      issued++;
      wstage = (wstage + 1 == STAGES) ? 0 : wstage + 1;
      cp_async_commit();
    };
    __syncthreads();   // every warp is done reading the stages of the previous item
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) issue();

    const int aoff = (cfg.a0 * 8 + g) * LDS + t;
    const int boff = (cfg.b0 * 8 + g) * LDS + t;
    const double* Bp = same ? As : Bs;
    const double* abase = (cfg.transposed ? Bp : As) + aoff;
    const double* bbase = (cfg.transposed ? As : Bp) + boff;
    int rstage = 0;
    const bool late = warp >= split;
    for (int chunk = 0; chunk < nchunks; chunk++) {
      // The copy instructions of a chunk keep a warp away from the DMMA pipe for several hundred clocks.  Of the two
      // warps of a scheduler (w, w + 4) the first issues the copies of chunk + 2 right after the barrier, the second
      // ("late") issues those of chunk + 1 just before it, i.e. after its MMAs of the previous chunk, so that one of the
      // two always feeds the pipe.  Both orders write a stage that every warp left before the previous barrier.
      if (!late) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
      }
      if (!late || chunk > 0) issue();
      if (late) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
      }
      const double* as = abase + rstage * (WL_T * LDS);
      const double* bs = bbase + rstage * (WL_T * LDS);
      rstage = (rstage + 1 == STAGES) ? 0 : rstage + 1;
      wl_stage_dispatch<0, BK / 4, LDS>(shape, acc, as, bs);
    }
    cp_async_wait<0>();

    // partial tile: dense 128 x 128, column-major (row = G row inside the tile = A-panel column)
    double* o = part + (int64_t)it * (WL_T * WL_T);
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (!((mask >> (i * 4 + j)) & 1u)) continue;
        const int ra = (cfg.a0 + i) * 8 + g;          // index along the MMA row operand
        const int cb = (cfg.b0 + j) * 8 + 2 * t;      // index along the MMA column operand
        if (!cfg.transposed) {
          o[ra + cb * WL_T] = acc[i][j][0];
          o[ra + (cb + 1) * WL_T] = acc[i][j][1];
        } else {
          o[cb + ra * WL_T] = acc[i][j][0];
          o[cb + 1 + ra * WL_T] = acc[i][j][1];
        }
      }
  }
}

// G[r, c] = sum over the item slots of its tile, in item order; upper: the lower triangle is the mirrored upper one.
__global__ void gram_wl_reduce_kernel(const double* __restrict__ part, const int* __restrict__ tile_first, int ntm,
                                      int ma, int mb, int upper, double* __restrict__ G, int ldg) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)ma * mb) return;
  const int r = (int)(idx % ma), c = (int)(idx / ma);
  const bool flip = upper && r > c;
  const int sr = flip ? c : r, sc = flip ? r : c;
  const int ti = sr / WL_T, tj = sc / WL_T;
  const int tile = upper ? ti + tj * (tj + 1) / 2 : ti + tj * ntm;
  const int s0 = tile_first[tile], s1 = tile_first[tile + 1];
  const int64_t off = (sr - ti * WL_T) + (int64_t)(sc - tj * WL_T) * WL_T;
  double s = 0.0;
  for (int k = s0; k < s1; k++) s += part[(int64_t)k * (WL_T * WL_T) + off];
  G[r + (int64_t)c * ldg] = s;
}

// ------------------------------------------------------------------------------------------------- schedule
struct WlSchedule {
  int nitems = 0, ncta = 0, ntm = 0, ntn = 0;
  void* dev = nullptr;           // one allocation: items | cta_first | layouts | tile_first
  const WlItem* items = nullptr;
  const int* cta_first = nullptr;
  const WlWarp* layouts = nullptr;
  const int* tile_first = nullptr;
};
using WlKey = std::tuple<int, int, int, int64_t, int, int, int, int>;   // ma, mb, upper + 2 * same_ab, n, ncta, BK, load_pct, phase
constexpr int WL_LOAD_PCT = 70;   // default staging cost of a 256-column tile relative to its DMMA time, in %
struct WlCache {
  std::map<WlKey, WlSchedule> map;
};

inline int popc(uint32_t v) { return __builtin_popcount(v); }

// shape id of a mask (the kernel only has code for these); -1 = not representable
int shape_of(uint32_t mask) {
  if (mask == 0) return WL_SHAPE_NONE;
  for (int na = 1; na <= 8; na++)
    for (int nb = 1; nb <= 4; nb++)
      if (mask == wl_rect_mask(na, nb)) return (na - 1) * 4 + (nb - 1);
  if (mask == wl_tri_a_mask()) return WL_SHAPE_TRI_A;
  if (mask == wl_tri_b_mask()) return WL_SHAPE_TRI_B;
  return -1;
}

// warp grid for a rectangular tile of ra x cb 8-blocks: every warp gets an (na x nb) rectangle of blocks.  Ragged
// diagonal tiles (the last one of a Hermitian product) also come here and compute their few sub-diagonal blocks.
int rect_layout(int ra, int cb, WlWarp (&w)[8], int& cost) {
  int best = 1 << 30, bWA = 2, bTr = 0;
  for (int WA = 1; WA <= 8; WA *= 2) {
    const int WB = 8 / WA;
    const int a = (ra + WA - 1) / WA, b = (cb + WB - 1) / WB;
    for (int tr = 0; tr < 2; tr++) {
      const bool ok = tr ? (a <= 4 && b <= 8) : (a <= 8 && b <= 4);
      if (!ok) continue;
      if (a * b < best) { best = a * b; bWA = WA; bTr = tr; }
    }
  }
  const int WA = bWA, WB = 8 / WA;
  const int a = (ra + WA - 1) / WA, b = (cb + WB - 1) / WB;
  cost = 0;
  for (int q = 0; q < 8; q++) {
    const int wa = q % WA, wb = q / WA;
    WlWarp c{};
    c.transposed = (uint8_t)bTr;
    const int r0 = wa * a, c0 = wb * b;      // first G-row block / G-column block of this warp
    const int nr = std::max(0, std::min(a, ra - r0)), nc = std::max(0, std::min(b, cb - c0));
    if (nr > 0 && nc > 0) {
      c.a0 = (uint8_t)(bTr ? c0 : r0);
      c.b0 = (uint8_t)(bTr ? r0 : c0);
      c.mask = bTr ? wl_rect_mask(nc, nr) : wl_rect_mask(nr, nc);
    }
    const int sh = shape_of(c.mask);
    if (sh < 0) return -4;
    c.shape = (uint8_t)sh;
    w[q] = c;
    cost = std::max(cost, popc(c.mask));
  }
  return 0;
}

// balanced cover of the upper triangle of a full 16 x 16-block diagonal tile: 136 blocks, at most 18 per warp
int diag16_layout(WlWarp (&w)[8], int& cost) {
  auto owner = [](int r, int c) {
    if (r <= 3) return c <= 5 ? 0 : c <= 7 ? 1 : c <= 11 ? 2 : 4;
    if (r <= 7) return c <= 7 ? 1 : c <= 11 ? 3 : 5;
    if (r <= 11) return c <= 13 ? 6 : 7;
    return 7;
  };
  //                    a0  b0  transposed      (transposed: a0 = first G-column block, b0 = first G-row block)
  const int cfg[8][3] = {{0, 0, 1}, {0, 4, 0}, {0, 8, 0}, {4, 8, 0}, {0, 12, 0}, {4, 12, 0}, {8, 8, 1}, {8, 12, 0}};
  cost = 0;
  for (int q = 0; q < 8; q++) {
    WlWarp c{};
    c.a0 = (uint8_t)cfg[q][0];
    c.b0 = (uint8_t)cfg[q][1];
    c.transposed = (uint8_t)cfg[q][2];
    for (int i = 0; i < 8; i++)
      for (int j = 0; j < 4; j++) {
        const int row = c.transposed ? c.b0 + j : c.a0 + i, col = c.transposed ? c.a0 + i : c.b0 + j;
        if (row < 16 && col < 16 && row <= col && owner(row, col) == q) c.mask |= 1u << (i * 4 + j);
      }
    const int sh = shape_of(c.mask);
    if (sh < 0) return -4;
    c.shape = (uint8_t)sh;
    w[q] = c;
    cost = std::max(cost, popc(c.mask));
  }
  return 0;
}

struct WlPlan {   // host-side schedule
  int ntm = 0, ntn = 0, ntiles = 0;
  std::vector<WlItem> items;
  std::vector<int> item_cta, item_tile, cta_first, tile_first;
  std::vector<WlWarp> layouts;
  std::vector<double> tile_cost;
};

struct WlTile {      // one output tile of a schedule
  int ti, tj;        // tile coordinates (standard products: G rows ti*128.., columns tj*128..)
  int a_col0, a_cols, b_col0, b_cols;
  int layout, same, b_sel;
  double cost;
};
using WlLayoutCache = std::map<std::tuple<int, int, int>, std::pair<int, int>>;   // (kind, ra, cb) -> (layout index, max blocks)

// warp layout + cost of a tile of a_cols x b_cols; kind 1 = diagonal tile of a Hermitian product
int make_tile(WlTile& tl, int kind, int same, int load_pct, std::vector<WlWarp>& layouts, WlLayoutCache& lcache) {
  const int ra = (tl.a_cols + 7) / 8, cb = (tl.b_cols + 7) / 8;
  auto key = std::make_tuple(kind, ra, cb);
  auto f = lcache.find(key);
  if (f == lcache.end()) {
    WlWarp w[8];
    int cost = 0;
    const int lrc = (kind == 1 && ra == WL_BLK && cb == WL_BLK) ? diag16_layout(w, cost) : rect_layout(ra, cb, w, cost);
    if (lrc) return lrc;
    const int idx = (int)layouts.size();
    layouts.insert(layouts.end(), w, w + 8);
    f = lcache.emplace(key, std::make_pair(idx, cost)).first;
  }
  tl.layout = f->second.first;
  // cost relative to a full tile: DMMA blocks of the busiest warp, bounded below by the staging traffic
  const double mma = f->second.second / 32.0;
  tl.same = same;
  const double load = 0.01 * load_pct * (tl.a_cols + (tl.same ? 0 : tl.b_cols)) / (2.0 * WL_T);
  tl.cost = std::max(mma, load);
  return 0;
}

// tiles laid end to end, weighted by their cost, cut into ncta equal pieces
int schedule_tiles(const std::vector<WlTile>& tiles, int64_t n, int ncta, int BK, int phase, WlPlan& P) {
  const int ntiles = (int)tiles.size();
  P.tile_cost.clear();
  for (auto& tl : tiles) P.tile_cost.push_back(tl.cost);
  double total = 0;
  for (auto& tl : tiles) total += tl.cost * (double)n;
  const double L = total / ncta;
  const int64_t min_rows = 16 * (int64_t)BK;
  std::vector<WlItem>& items = P.items;
  P.cta_first.assign(ncta + 1, 0);
  P.tile_first.assign(ntiles + 1, 0);
  double U = 0;
  for (int tix = 0; tix < ntiles; tix++) {
    const WlTile& tl = tiles[tix];
    P.tile_first[tix] = (int)items.size();
    const double span = tl.cost * (double)n;
    int b_lo = (int)std::floor(U / L), b_hi = (int)std::floor((U + span) / L);
    b_lo = std::min(std::max(b_lo, 0), ncta - 1);
    b_hi = std::min(std::max(b_hi, 0), ncta - 1);
    auto boundary = [&](int b) -> int64_t {   // first row of tile `tix` that belongs to CTA b (or later)
      if (b <= b_lo) return 0;
      if (b > b_hi) return n;
      const double u = (double)b * L - U;
      int64_t r = (int64_t)std::llround(u / tl.cost / BK) * BK;
      if (r < min_rows) r = 0;
      if (n - r < min_rows) r = n;
      return std::min<int64_t>(std::max<int64_t>(r, 0), n);
    };
    for (int b = b_lo; b <= b_hi; b++) {
      const int64_t r0 = boundary(b), r1 = boundary(b + 1);
      if (r1 <= r0) continue;
      WlItem itx{};
      itx.r_begin = r0; itx.r_end = r1;
      itx.a_col0 = tl.a_col0; itx.a_cols = tl.a_cols;
      itx.b_col0 = tl.b_col0; itx.b_cols = tl.b_cols;
      itx.layout = tl.layout;
      itx.same_panel = tl.same;
      itx.b_sel = tl.b_sel;
      // Phase alignment: a piece starts at the chunk whose row is a multiple of its own length, so that all pieces of
      // (nearly) equal length — the full-CTA pieces of the rectangular tiles — sit at rows congruent modulo that length
      // at every moment; tiles that share a panel then read the same rows at the same time and the panel comes out of
      // L2 instead of DRAM (without it every one of the m/128 tiles of a panel streams it from HBM on its own).
      if (phase) {
        const int64_t nch = (r1 - r0 + BK - 1) / BK, a = r0 / BK;
        const int64_t len = std::max<int64_t>(1, std::llround(L / tl.cost / BK));   // chunks of a full-CTA piece of this tile
        itx.c_start = (int32_t)(((len - a % len) % len) % nch);
      }
      items.push_back(itx);
      P.item_cta.push_back(b);
      P.item_tile.push_back(tix);
    }
    U += span;
  }
  P.tile_first[ntiles] = (int)items.size();
  const int nitems = (int)items.size();
  // items are ordered by (tile, CTA) = by position on the line, so each CTA owns a contiguous run
  int it = 0;
  for (int b = 0; b < ncta; b++) {
    while (it < nitems && P.item_cta[it] < b) it++;
    P.cta_first[b] = it;
  }
  P.cta_first[ncta] = nitems;
  for (int i = 1; i < nitems; i++)
    if (P.item_cta[i] < P.item_cta[i - 1]) return -3;   // cannot happen
  P.ntiles = ntiles;
  return 0;
}

int plan_schedule(int ma, int mb, int upper, int same_ab, int64_t n, int ncta, int BK, int load_pct, WlPlan& P,
                  int phase = 1) {
  const int ntm = (ma + WL_T - 1) / WL_T, ntn = (mb + WL_T - 1) / WL_T;
  std::vector<WlTile> tiles;
  WlLayoutCache lcache;
  for (int tj = 0; tj < ntn; tj++)
    for (int ti = 0; ti < (upper ? tj + 1 : ntm); ti++) {
      WlTile tl{};
      tl.ti = ti; tl.tj = tj;
      tl.a_col0 = ti * WL_T; tl.b_col0 = tj * WL_T;
      tl.a_cols = std::min(WL_T, ma - ti * WL_T);
      tl.b_cols = std::min(WL_T, mb - tj * WL_T);
      const int kind = (upper && ti == tj) ? 1 : 0;
      if (int rc = make_tile(tl, kind, (same_ab && upper && ti == tj) ? 1 : 0, load_pct, P.layouts, lcache)) return rc;
      tiles.push_back(tl);
    }
  if (int rc = schedule_tiles(tiles, n, ncta, BK, phase, P)) return rc;
  P.ntm = ntm; P.ntn = ntn;
  return 0;
}

// Column-block products (gram_wl_cols_f64): G_q[0:m, 0:nw] = S^H W_q for q < nprod with S = [.. | W-block ..] whose
// columns tri_c0 .. tri_c0 + nw - 1 are the W block of a Hermitian product (tri_c0 < 0: plain rectangular products).
// Tiles that lie entirely below the diagonal of that block (first row - tri_c0 > last column) are left out: the caller
// mirrors them from the upper part.  P.tiles_out lists the tiles in schedule order.
struct WlTileOut { int32_t a_col0, a_cols, b_col0, b_cols, b_sel, first, last, diag, msplit; };   // b_sel 2: merged remainder tile, columns < msplit -> G0, the rest -> G1
int plan_schedule_cols(int m, int nw, int nprod, int tri_c0, int64_t n, int ncta, int BK, int load_pct, int phase,
                       WlPlan& P, std::vector<WlTileOut>& out, int merge = 1) {
  // Row tiles: 128-column steps over [0, tri_c0) and, separately, over the Hermitian block [tri_c0, tri_c0 + nw), so
  // that the tiles of W^H W / W^H A W are aligned with the column tiles: tiles below the diagonal are left out, full
  // diagonal tiles use the balanced upper-triangle cover (diag16_layout), exactly as in a Hermitian product.
  std::vector<std::pair<int, int>> rows;   // (first column of S, columns)
  const int split = (tri_c0 >= 0 && tri_c0 + nw == m) ? tri_c0 : m;
  for (int c = 0; c < split; c += WL_T) rows.emplace_back(c, std::min(WL_T, split - c));
  for (int c = split; c < m; c += WL_T) rows.emplace_back(c, std::min(WL_T, m - c));
  const int ntn = (nw + WL_T - 1) / WL_T;
  std::vector<WlTile> tiles;
  std::vector<int> is_diag;
  WlLayoutCache lcache;
  // MERGED REMAINDER TILE: the last, ragged column tile of k = 300 is 44 columns wide — a 128 x 44 tile stages 172 columns for 96
  // MMA blocks and is bound by its copies, twice per S panel (both products).  With both products present and a remainder of
  // at most 64 columns the two remainder panels are gathered into one scratch block [W0 rem | W1 rem] (gram_wl_cols_f64) and
  // computed as ONE tile of 2 * rem columns (b_sel = 2); the reduction scatters its halves to the two outputs.
  const int nfull = nw / WL_T, rem = nw - nfull * WL_T;
  const bool merged = merge && nprod == 2 && rem > 0 && 2 * rem <= WL_T;
  // row tile outermost: the tiles of one S panel (both products, every column tile) are neighbours on the line
  for (size_t ri = 0; ri < rows.size(); ri++)
    for (int q = 0; q < (merged ? nprod + 1 : nprod); q++)
      for (int tj = (q == nprod ? nfull : 0); tj < (q == nprod ? nfull + 1 : (merged ? nfull : ntn)); tj++) {
        WlTile tl{};
        tl.ti = (int)ri; tl.tj = tj; tl.b_sel = q;
        tl.a_col0 = rows[ri].first; tl.a_cols = rows[ri].second;
        tl.b_col0 = (q == nprod) ? 0 : tj * WL_T;                       // merged panel: column 0 of the gathered block
        tl.b_cols = (q == nprod) ? 2 * rem : std::min(WL_T, nw - tj * WL_T);
        int kind = 0;
        if (tl.a_col0 >= split) {   // inside the Hermitian block (aligned tiles)
          const int wi = (tl.a_col0 - split) / WL_T;
          if (wi > tj) continue;
          kind = (wi == tj && q != nprod) ? 1 : 0;
        }
        if (int rc = make_tile(tl, kind, 0, load_pct, P.layouts, lcache)) return rc;
        tiles.push_back(tl);
        is_diag.push_back(kind == 1 && (tl.a_cols + 7) / 8 == WL_BLK && (tl.b_cols + 7) / 8 == WL_BLK);   // same rule as make_tile: narrower diagonal tiles compute the full square
      }
  if (tiles.empty()) return -5;
  if (int rc = schedule_tiles(tiles, n, ncta, BK, phase, P)) return rc;
  P.ntm = (int)rows.size(); P.ntn = ntn;
  out.clear();
  for (size_t i = 0; i < tiles.size(); i++) {
    const WlTile& tl = tiles[i];
    const bool mt = tl.b_sel == nprod;   // merged remainder tile
    out.push_back(WlTileOut{tl.a_col0, tl.a_cols, mt ? nfull * WL_T : tl.b_col0, tl.b_cols, mt ? 2 : tl.b_sel, P.tile_first[i],
                            P.tile_first[i + 1], is_diag[i], mt ? rem : 0});
  }
  return 0;
}

int build_schedule(lb2_ctx* ctx, int ma, int mb, int upper, int same_ab, int64_t n, int ncta, int BK, int load_pct,
                   int phase, WlSchedule& S) {
  WlPlan P;
  const int rc = plan_schedule(ma, mb, upper, same_ab, n, ncta, BK, load_pct, P, phase);
  if (rc) return rc;
  const int nitems = (int)P.items.size(), ntiles = P.ntiles;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t o_items = 0, o_cta = al(sizeof(WlItem) * nitems), o_lay = o_cta + al(sizeof(int) * (ncta + 1)),
               o_tile = o_lay + al(sizeof(WlWarp) * P.layouts.size()), tot = o_tile + al(sizeof(int) * (ntiles + 1));
  std::vector<char> host(tot, 0);
  memcpy(host.data() + o_items, P.items.data(), sizeof(WlItem) * nitems);
  memcpy(host.data() + o_cta, P.cta_first.data(), sizeof(int) * (ncta + 1));
  memcpy(host.data() + o_lay, P.layouts.data(), sizeof(WlWarp) * P.layouts.size());
  memcpy(host.data() + o_tile, P.tile_first.data(), sizeof(int) * (ntiles + 1));
  LB2_CUDA_OK(cudaMalloc(&S.dev, tot));
  LB2_CUDA_OK(cudaMemcpyAsync(S.dev, host.data(), tot, cudaMemcpyHostToDevice, ctx->stream));
  LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  const char* d = (const char*)S.dev;
  S.items = (const WlItem*)(d + o_items);
  S.cta_first = (const int*)(d + o_cta);
  S.layouts = (const WlWarp*)(d + o_lay);
  S.tile_first = (const int*)(d + o_tile);
  S.nitems = nitems; S.ncta = ncta; S.ntm = P.ntm; S.ntn = P.ntn;
  return 0;
}

template <int BK, int STAGES>
int launch_wl(lb2_ctx* ctx, const WlSchedule& S, int64_t n, int ma, int mb, const double* A, int64_t lda,
              const double* B, int64_t ldb, double* G, int ldg, int upper) {
  double* part = (double*)ctx_scratch(ctx, sizeof(double) * (size_t)S.nitems * WL_T * WL_T);
  if (!part) return -1;
  const bool vec = (lda % 2 == 0) && (ldb % 2 == 0) && ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0);
  constexpr size_t smem = sizeof(double) * (size_t)STAGES * 2 * WL_T * (BK + 4);
  if (vec) {
    auto k = gram_wl_kernel<BK, STAGES, true>;
    LB2_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<S.ncta, WL_NT, smem, ctx->stream>>>(A, lda, B, ldb, B, ldb, B, ldb, S.items, S.cta_first, S.layouts, part,
                                            ctx->nn_stagger ? 4 : 8);
  } else {
    auto k = gram_wl_kernel<BK, STAGES, false>;
    LB2_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<S.ncta, WL_NT, smem, ctx->stream>>>(A, lda, B, ldb, B, ldb, B, ldb, S.items, S.cta_first, S.layouts, part,
                                            ctx->nn_stagger ? 4 : 8);
  }
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  const int64_t tot = (int64_t)ma * mb;
  gram_wl_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(part, S.tile_first, S.ntm, ma, mb,
                                                                               upper, G, ldg);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// G_q[a_col0 + r, b_col0 + c] = sum over the item slots of the tile, in item order (column-block products)
__global__ void gram_wl_reduce_tiles_kernel(const double* __restrict__ part, const WlTileOut* __restrict__ tiles,
                                            double* __restrict__ G0, int ldg0, double* __restrict__ G1, int ldg1) {
  const WlTileOut t = tiles[blockIdx.x];
  double* __restrict__ G = t.b_sel == 1 ? G1 : G0;
  const int ldg = t.b_sel == 1 ? ldg1 : ldg0;
  const int tot = t.a_cols * t.b_cols;
  for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < tot; idx += gridDim.y * blockDim.x) {
    const int r = idx % t.a_cols, c = idx / t.a_cols;
    if (t.diag && r > c) continue;   // balanced upper-triangle cover: the blocks below the diagonal were not computed
    const int64_t off = r + (int64_t)c * WL_T;
    double s = 0.0;
    for (int k = t.first; k < t.last; k++) s += part[(int64_t)k * (WL_T * WL_T) + off];
    if (t.b_sel == 2) {   // merged remainder tile: [W0 rem | W1 rem]
      if (c < t.msplit) G0[(t.a_col0 + r) + (int64_t)(t.b_col0 + c) * ldg0] = s;
      else G1[(t.a_col0 + r) + (int64_t)(t.b_col0 + c - t.msplit) * ldg1] = s;
    } else {
      G[(t.a_col0 + r) + (int64_t)(t.b_col0 + c) * ldg] = s;
    }
  }
}

struct WlColsSchedule {
  WlSchedule base;
  const WlTileOut* tiles_dev = nullptr;
  int ntiles = 0;
  int merged_rem = 0;   // > 0: the schedule has merged remainder tiles of 2 * merged_rem columns (b_sel = 2)
};
using WlColsKey = std::tuple<int, int, int, int, int64_t, int, int, int, int, int>;   // m, nw, nprod, tri_c0, n, ncta, BK, load_pct, phase, merge

int build_schedule_cols(lb2_ctx* ctx, int m, int nw, int nprod, int tri_c0, int64_t n, int ncta, int BK, int load_pct,
                        int phase, int merge, WlColsSchedule& S) {
  WlPlan P;
  std::vector<WlTileOut> tout;
  const int rc = plan_schedule_cols(m, nw, nprod, tri_c0, n, ncta, BK, load_pct, phase, P, tout, merge);
  if (rc) return rc;
  for (auto& t : tout)
    if (t.b_sel == 2) S.merged_rem = t.msplit;
  const int nitems = (int)P.items.size(), ntiles = P.ntiles;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t o_items = 0, o_cta = al(sizeof(WlItem) * nitems), o_lay = o_cta + al(sizeof(int) * (ncta + 1)),
               o_tile = o_lay + al(sizeof(WlWarp) * P.layouts.size()), tot = o_tile + al(sizeof(WlTileOut) * ntiles);
  std::vector<char> host(tot, 0);
  memcpy(host.data() + o_items, P.items.data(), sizeof(WlItem) * nitems);
  memcpy(host.data() + o_cta, P.cta_first.data(), sizeof(int) * (ncta + 1));
  memcpy(host.data() + o_lay, P.layouts.data(), sizeof(WlWarp) * P.layouts.size());
  memcpy(host.data() + o_tile, tout.data(), sizeof(WlTileOut) * ntiles);
  LB2_CUDA_OK(cudaMalloc(&S.base.dev, tot));
  LB2_CUDA_OK(cudaMemcpyAsync(S.base.dev, host.data(), tot, cudaMemcpyHostToDevice, ctx->stream));
  LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  const char* d = (const char*)S.base.dev;
  S.base.items = (const WlItem*)(d + o_items);
  S.base.cta_first = (const int*)(d + o_cta);
  S.base.layouts = (const WlWarp*)(d + o_lay);
  S.tiles_dev = (const WlTileOut*)(d + o_tile);
  S.base.nitems = nitems; S.base.ncta = ncta; S.base.ntm = P.ntm; S.base.ntn = P.ntn;
  S.ntiles = ntiles;
  return 0;
}

template <int BK, int STAGES>
int launch_wl_cols(lb2_ctx* ctx, const WlColsSchedule& S, int64_t n, int nw, const double* A, int64_t lda, const double* B0,
                   int64_t ldb0, double* G0, int ldg0, const double* B1, int64_t ldb1, double* G1, int ldg1) {
  const size_t part_bytes = (sizeof(double) * (size_t)S.base.nitems * WL_T * WL_T + 255) & ~(size_t)255;
  const int rem = S.merged_rem;
  const int64_t ldm = (n + 1) & ~(int64_t)1;                                  // even leading dimension: 16-byte copies
  double* part = (double*)ctx_scratch(ctx, part_bytes + (rem ? sizeof(double) * (size_t)ldm * 2 * rem : 0));
  if (!part) return -1;
  double* Bm = (double*)((char*)part + part_bytes);
  if (rem) {   // gather the two ragged remainder panels side by side: [W0[:, nw - rem :] | W1[:, nw - rem :]]
    const int c0 = nw - rem;
    LB2_CUDA_OK(cudaMemcpy2DAsync(Bm, sizeof(double) * ldm, B0 + (int64_t)c0 * ldb0, sizeof(double) * ldb0, sizeof(double) * n, rem,
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    LB2_CUDA_OK(cudaMemcpy2DAsync(Bm + (int64_t)rem * ldm, sizeof(double) * ldm, B1 + (int64_t)c0 * ldb1, sizeof(double) * ldb1,
                                  sizeof(double) * n, rem, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  const bool vec = (lda % 2 == 0) && (ldb0 % 2 == 0) && (ldb1 % 2 == 0) && ((uintptr_t)A % 16 == 0) &&
                   ((uintptr_t)B0 % 16 == 0) && ((uintptr_t)B1 % 16 == 0);
  constexpr size_t smem = sizeof(double) * (size_t)STAGES * 2 * WL_T * (BK + 4);
  if (vec) {
    auto k = gram_wl_kernel<BK, STAGES, true>;
    LB2_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<S.base.ncta, WL_NT, smem, ctx->stream>>>(A, lda, B0, ldb0, B1, ldb1, Bm, ldm, S.base.items, S.base.cta_first,
                                                  S.base.layouts, part, ctx->nn_stagger ? 4 : 8);
  } else {
    auto k = gram_wl_kernel<BK, STAGES, false>;
    LB2_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<S.base.ncta, WL_NT, smem, ctx->stream>>>(A, lda, B0, ldb0, B1, ldb1, Bm, ldm, S.base.items, S.base.cta_first,
                                                  S.base.layouts, part, ctx->nn_stagger ? 4 : 8);
  }
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  gram_wl_reduce_tiles_kernel<<<dim3((unsigned)S.ntiles, 8), 256, 0, ctx->stream>>>(part, S.tiles_dev, G0, ldg0, G1, ldg1);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

struct WlColsCache {
  std::map<WlColsKey, WlColsSchedule> map;
};

}  // namespace

// Host-only self-check of a schedule (no CUDA calls; tests/test_abi.py runs it on CPU for many shapes): every needed
// 8x8 block of every tile is owned by exactly one warp, the items of a tile partition [0, n) exactly, every CTA owns a
// contiguous run of items.  stats: [0] items, [1] busiest CTA cost / mean CTA cost, [2] issued DMMA blocks x rows over
// needed blocks x rows (1 = no wasted MMA), [3] tiles.
int gram_wl_plan_check(int ma, int mb, int upper, int64_t n, int ncta, int BK, double* stats) {
  WlPlan P;
  int rc = plan_schedule(ma, mb, upper, upper, n, ncta, BK, WL_LOAD_PCT, P);
  if (rc) return rc;
  if (upper && ma != mb) return -2;
  std::vector<double> cta_cost(ncta, 0.0);
  double issued = 0, needed = 0;
  for (int tix = 0; tix < P.ntiles; tix++) {
    const int i0 = P.tile_first[tix], i1 = P.tile_first[tix + 1];
    if (i1 <= i0) return 10;
    int64_t r = 0;
    for (int i = i0; i < i1; i++) {
      if (P.items[i].r_begin != r || P.items[i].r_end <= r || P.item_tile[i] != tix) return 11;
      if (i > i0 && P.items[i].r_begin % BK) return 12;
      if (P.items[i].c_start < 0 || (int64_t)P.items[i].c_start * BK >= P.items[i].r_end - P.items[i].r_begin) return 18;
      r = P.items[i].r_end;
      cta_cost[P.item_cta[i]] += P.tile_cost[tix] * (double)(P.items[i].r_end - P.items[i].r_begin);
    }
    if (r != n) return 13;
    const WlItem& it0 = P.items[i0];
    const int ra = (it0.a_cols + 7) / 8, cb = (it0.b_cols + 7) / 8;
    const bool diag = upper && it0.a_col0 == it0.b_col0 && ra == WL_BLK && cb == WL_BLK;   // ragged diagonal tiles compute the full square
    int owner[16][16];
    for (auto& row : owner) for (int& v : row) v = 0;
    int busiest = 0;
    for (int w = 0; w < 8; w++) {
      const WlWarp& c = P.layouts[it0.layout + w];
      if (shape_of(c.mask) != c.shape) return 17;
      busiest = std::max(busiest, popc(c.mask));
      for (int i = 0; i < 8; i++)
        for (int j = 0; j < 4; j++)
          if ((c.mask >> (i * 4 + j)) & 1u) {
            const int row = c.transposed ? c.b0 + j : c.a0 + i, col = c.transposed ? c.a0 + i : c.b0 + j;
            if (row >= 16 || col >= 16) return 14;
            owner[row][col]++;
          }
    }
    int need = 0;
    for (int row = 0; row < 16; row++)
      for (int col = 0; col < 16; col++) {
        const bool want = row < ra && col < cb && (!diag || row <= col);
        if (owner[row][col] != (want ? 1 : 0)) return 15;
        need += want;
      }
    issued += 8.0 * busiest * (double)n;
    needed += (double)need * (double)n;
  }
  for (int b = 0; b < ncta; b++)
    for (int i = P.cta_first[b]; i < P.cta_first[b + 1]; i++)
      if (P.item_cta[i] != b) return 16;
  double mx = 0, sum = 0;
  for (double c : cta_cost) { mx = std::max(mx, c); sum += c; }
  if (stats) {
    stats[0] = (double)P.items.size();
    stats[1] = mx / (sum / ncta);
    stats[2] = issued / needed;
    stats[3] = (double)P.ntiles;
  }
  return 0;
}

// Host-only model of the operand sharing of a schedule (tests/test_gram_plan.py, tools/gram_phase_sim.py): every CTA
// advances through its pieces at exactly the cost rate of their tiles; at `samples` instants the (matrix, panel, row
// window) triples the CTAs are reading are collected, windows of `window` chunks counting as one L2-resident region.
// *share = distinct requests / all requests: the fraction of the panel traffic that has to come from DRAM if L2 serves
// every repeated request (1 = nothing shared).
int gram_wl_plan_sharing(int ma, int mb, int upper, int64_t n, int ncta, int BK, int phase, int window, int samples,
                         double* share) {
  if (!share || window < 1 || samples < 1) return -1;
  WlPlan P;
  const int rc = plan_schedule(ma, mb, upper, upper, n, ncta, BK, WL_LOAD_PCT, P, phase);
  if (rc) return rc;
  double T = 0;
  for (int b = 0; b < ncta; b++) {
    double t = 0;
    for (int i = P.cta_first[b]; i < P.cta_first[b + 1]; i++)
      t += P.tile_cost[P.item_tile[i]] * (double)(P.items[i].r_end - P.items[i].r_begin);
    T = std::max(T, t);
  }
  double total = 0, uniq = 0;
  std::vector<std::tuple<int, int, int64_t>> reads;
  for (int sidx = 0; sidx < samples; sidx++) {
    const double t = (sidx + 0.5) / samples * T;
    reads.clear();
    for (int b = 0; b < ncta; b++) {
      double tt = t;
      for (int i = P.cta_first[b]; i < P.cta_first[b + 1]; i++) {
        const WlItem& it = P.items[i];
        const double cost = P.tile_cost[P.item_tile[i]];
        const int64_t rows = it.r_end - it.r_begin, nch = (rows + BK - 1) / BK;
        const double dur = cost * (double)rows;
        if (tt < dur) {
          const int64_t c = (it.c_start + (int64_t)(tt / (cost * BK))) % nch;
          const int64_t w = (it.r_begin / BK + c) / window;
          reads.emplace_back(0, it.a_col0 / WL_T, w);
          if (!it.same_panel) reads.emplace_back(upper ? 0 : 1, it.b_col0 / WL_T, w);
          break;
        }
        tt -= dur;
      }
    }
    total += (double)reads.size();
    std::sort(reads.begin(), reads.end());
    uniq += (double)(std::unique(reads.begin(), reads.end()) - reads.begin());
  }
  *share = total > 0 ? uniq / total : 1.0;
  return 0;
}

void gram_wl_cache_free(lb2_ctx* ctx) {
  WlCache* c = (WlCache*)ctx->gram_wl_cache;
  if (c) {
    for (auto& kv : c->map)
      if (kv.second.dev) cudaFree(kv.second.dev);
    delete c;
    ctx->gram_wl_cache = nullptr;
  }
  WlColsCache* cc = (WlColsCache*)ctx->gram_wl_cols_cache;
  if (cc) {
    for (auto& kv : cc->map)
      if (kv.second.base.dev) cudaFree(kv.second.base.dev);
    delete cc;
    ctx->gram_wl_cols_cache = nullptr;
  }
}

static int run_wl(lb2_ctx* ctx, int64_t n, int ma, int mb, const double* A, int64_t lda, const double* B,
                  int64_t ldb, double* G, int ldg, int upper) {
  if (!ctx->gram_wl_cache) ctx->gram_wl_cache = new WlCache();
  WlCache* cache = (WlCache*)ctx->gram_wl_cache;
  const int BK = (ctx->gram_bk == 16) ? 16 : 32;
  // one CTA per SM; small problems use fewer CTAs so that a piece is never shorter than ~1024 rows of a full tile
  const int64_t tiles_full = (int64_t)((ma + WL_T - 1) / WL_T) * ((mb + WL_T - 1) / WL_T);
  const int ncta = (int)std::min<int64_t>(ctx->sm_count, std::max<int64_t>(1, tiles_full * n / 4096));
  const int load_pct = ctx->gram_load_pct > 0 ? ctx->gram_load_pct : WL_LOAD_PCT;
  const int same_ab = (upper && A == B && lda == ldb) ? 1 : 0;
  const int phase = ctx->gram_phase != 0 ? 1 : 0;
  const WlKey key(ma, mb, (upper ? 1 : 0) + 2 * same_ab, n, ncta, BK, load_pct, phase);
  auto f = cache->map.find(key);
  if (f == cache->map.end()) {
    WlSchedule S;
    const int rc = build_schedule(ctx, ma, mb, upper ? 1 : 0, same_ab, n, ncta, BK, load_pct, phase, S);
    if (rc) return rc;
    f = cache->map.emplace(key, S).first;
  }
  if (BK == 32) return launch_wl<32, 3>(ctx, f->second, n, ma, mb, A, lda, B, ldb, G, ldg, upper);
  return launch_wl<16, 4>(ctx, f->second, n, ma, mb, A, lda, B, ldb, G, ldg, upper);
}

// ------------------------------------------------------------------------------------------------- narrow strips
// G[0:ma, c0:c0+r] = A^H B[:, c0:c0+r] for a NARROW strip (r <= 8; wider ones would re-read the strip from L2 too often): a few hundred MFLOP per streamed GB, so the only
// thing that matters is how A is read.  The tile kernels read BK x 8-byte column pieces (one DRAM page each, ~2 TB/s);
// here a CTA owns CB columns of A and one contiguous row range and streams each column in long runs (coalesced 2 KB
// per CTA and iteration), keeps a CB x RB block of dot products in FP64 FMA registers, and re-reads the r strip values of
// a row from L2 (grid.x = column groups is the fast index, so the CTAs that run together share the same strip rows).
template <int CB, int RB>
__global__ void __launch_bounds__(256)
    strip_gram_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ Bs, int64_t ldb, int ma, int r,
                      int64_t n, int64_t rows_per_split, double* __restrict__ part) {
  const int c0 = blockIdx.x * CB;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(n, r_begin + rows_per_split);
  double acc[CB][RB];
#pragma unroll
  for (int c = 0; c < CB; c++)
#pragma unroll
    for (int j = 0; j < RB; j++) acc[c][j] = 0.0;
  const double* ap[CB];
#pragma unroll
  for (int c = 0; c < CB; c++) ap[c] = A + (int64_t)min(c0 + c, ma - 1) * lda;   // columns past ma: duplicates, never stored
  for (int64_t row = r_begin + threadIdx.x; row < r_end; row += 256) {
    double b[RB], a[CB];
#pragma unroll
    for (int c = 0; c < CB; c++) a[c] = ap[c][row];
#pragma unroll
    for (int j = 0; j < RB; j++) b[j] = (j < r) ? Bs[row + (int64_t)j * ldb] : 0.0;
#pragma unroll
    for (int c = 0; c < CB; c++)
#pragma unroll
      for (int j = 0; j < RB; j++) acc[c][j] = fma(a[c], b[j], acc[c][j]);
  }
  __shared__ double red[8][CB * RB];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < CB; c++)
#pragma unroll
    for (int j = 0; j < RB; j++) {
      const double v = warp_sum(acc[c][j]);
      if (lane == 0) red[warp][c * RB + j] = v;
    }
  __syncthreads();
  if (threadIdx.x < CB * RB) {
    const int c = threadIdx.x / RB, j = threadIdx.x % RB;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) v += red[w][threadIdx.x];
    if (c0 + c < ma && j < r) part[(int64_t)blockIdx.y * ((int64_t)ma * r) + (c0 + c) + (int64_t)j * ma] = v;
  }
}

__global__ void strip_reduce_kernel(const double* __restrict__ part, int nsplit, int ma, int r, double* __restrict__ G,
                                    int ldg) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ma * r) return;
  double s = 0.0;
  for (int k = 0; k < nsplit; k++) s += part[(int64_t)k * ((int64_t)ma * r) + idx];
  G[(idx % ma) + (int64_t)(idx / ma) * ldg] = s;
}

// G[0:ma, 0:r] (leading dimension ldg) = A^H Bs, Bs = the r strip columns
static int strip_gram(lb2_ctx* ctx, int64_t n, int ma, int r, const double* A, int64_t lda, const double* Bs, int64_t ldb,
                      double* G, int ldg) {
  const int CB = 8;
  const int ngroups = (ma + CB - 1) / CB;
  // ~8 resident CTAs per SM; splits of at least 4096 rows
  int nsplit = std::max(1, (ctx->sm_count * 8 + ngroups - 1) / ngroups);
  nsplit = (int)std::min<int64_t>(nsplit, std::max<int64_t>(1, n / 4096));
  const int64_t rps = ((n + nsplit - 1) / nsplit + 255) / 256 * 256;
  nsplit = (int)((n + rps - 1) / rps);
  double* part = (double*)ctx_scratch(ctx, sizeof(double) * (size_t)nsplit * ma * r);
  if (!part) return -1;
  dim3 grid(ngroups, nsplit);
  if (r <= 4) strip_gram_kernel<8, 4><<<grid, 256, 0, ctx->stream>>>(A, lda, Bs, ldb, ma, r, n, rps, part);
  else strip_gram_kernel<8, 8><<<grid, 256, 0, ctx->stream>>>(A, lda, Bs, ldb, ma, r, n, rps, part);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  strip_reduce_kernel<<<(ma * r + 255) / 256, 256, 0, ctx->stream>>>(part, nsplit, ma, r, G, ldg);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// G[c, i] = G[i, c] for the strip columns c = c0 .. c0 + r - 1 and every row i < c: lower mirror of a strip that was
// computed as a rectangular product (its r x r diagonal block is made exactly symmetric, like the mirrored tiles)
__global__ void mirror_strip_kernel(double* __restrict__ G, int ldg, int c0, int r) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = c0 + r;
  if (idx >= m * r) return;
  const int i = idx % m, c = c0 + idx / m;
  if (i < c) G[c + (int64_t)i * ldg] = G[i + (int64_t)c * ldg];
}

// f64 Gram through the work-list kernel.  Same contract as gram<double>() in dense.cu.
//
// The work-list schedule has no two CTAs on the same rows at the same time, so every operand byte comes from HBM in
// BK x 8-byte pieces; measured on B200 that pattern tops out near 2 TB/s, which a full 128 x 128 tile just stays under
// (it needs 32 KB per 2 us and SM) but a narrow ragged tile does not (same panel traffic, a few DMMAs).  A Hermitian
// product whose last tile column is ragged (m mod 128 != 0) is therefore split: the leading multiple of 128 goes
// through the work list; the strip G[:, c0:m] through strip_gram_kernel when it is at most 8 columns wide (streams A
// once in long runs, FP64 FMA) and through the lock-step kernel of dense.cu otherwise (its CTAs share the panel
// through L2); the strip is then mirrored.
int gram_wl_f64(lb2_ctx* ctx, int64_t n, int ma, int mb, const double* A, int64_t lda, const double* B, int64_t ldb,
                double* G, int ldg, int upper) {
  const int strip_max = ctx->gram_strip_max > 0 ? ctx->gram_strip_max : (ctx->gram_strip_max < 0 ? 0 : WL_T);
  const int r = ma % WL_T, c0 = ma - r;
  if (upper && c0 > 0 && r > 0 && r < strip_max) {
    int rc = run_wl(ctx, n, c0, c0, A, lda, B, ldb, G, ldg, 1);
    if (rc) return rc;
    if (r <= 8 && ctx->gram_strip_fma >= 0)
      rc = strip_gram(ctx, n, ma, r, A, lda, B + (int64_t)c0 * ldb, ldb, G + (int64_t)c0 * ldg, ldg);
    else
      rc = gram_tiles_f64(ctx, n, ma, r, A, lda, B + (int64_t)c0 * ldb, ldb, G + (int64_t)c0 * ldg, ldg, 0);
    if (rc) return rc;
    mirror_strip_kernel<<<(ma * r + 255) / 256, 256, 0, ctx->stream>>>(G, ldg, c0, r);
    ctx->launches++;
    LB2_CUDA_OK(cudaGetLastError());
    return 0;
  }
  return run_wl(ctx, n, ma, mb, A, lda, B, ldb, G, ldg, upper);
}

// Column-block Gram products of the cached-Gram pass: with S = [X | P | W] (n x m) only the W columns of
// S^H B S and S^H A S are contracted over n,
//     G0[0:m, 0:nw] = S^H W0   (W0 = W or B W)        G1[0:m, 0:nw] = S^H W1   (W1 = A W),
// both in ONE launch (the S panel of a row tile feeds every tile of both products); W1 == nullptr: one product.
// tri_c0 >= 0 says that the rows tri_c0 .. tri_c0 + nw - 1 of the results are a Hermitian nw x nw block (W^H W, W^H A W):
// tiles entirely below its diagonal are not computed and those entries of G0 / G1 are left untouched (the caller
// mirrors them).  2 n m nw flop per product minus the skipped tiles.
int gram_wl_cols_f64(lb2_ctx* ctx, int64_t n, int m, int nw, const double* S, int64_t lds, const double* W0, int64_t ldw0,
                     double* G0, int ldg0, const double* W1, int64_t ldw1, double* G1, int ldg1, int tri_c0) {
  if (m <= 0 || nw <= 0) return 0;
  if (!W0 || !G0) return -2;
  const int nprod = (W1 && G1) ? 2 : 1;
  if (!ctx->gram_wl_cols_cache) ctx->gram_wl_cols_cache = new WlColsCache();
  WlColsCache* cache = (WlColsCache*)ctx->gram_wl_cols_cache;
  const int BK = (ctx->gram_bk == 16) ? 16 : 32;
  const int64_t tiles_full = (int64_t)((m + WL_T - 1) / WL_T) * ((nw + WL_T - 1) / WL_T) * nprod;
  const int ncta = (int)std::min<int64_t>(ctx->sm_count, std::max<int64_t>(1, tiles_full * n / 4096));
  const int load_pct = ctx->gram_load_pct > 0 ? ctx->gram_load_pct : WL_LOAD_PCT;
  const int phase = ctx->gram_phase != 0 ? 1 : 0;
  const int merge = ctx->gram_merge != 0 ? 1 : 0;
  const WlColsKey key(m, nw, nprod, tri_c0, n, ncta, BK, load_pct, phase, merge);
  auto f = cache->map.find(key);
  if (f == cache->map.end()) {
    WlColsSchedule Sc;
    const int rc = build_schedule_cols(ctx, m, nw, nprod, tri_c0, n, ncta, BK, load_pct, phase, merge, Sc);
    if (rc) return rc;
    f = cache->map.emplace(key, Sc).first;
  }
  const double* B1 = nprod == 2 ? W1 : W0;
  const int64_t ldb1 = nprod == 2 ? ldw1 : ldw0;
  double* Gq1 = nprod == 2 ? G1 : G0;
  const int ldq1 = nprod == 2 ? ldg1 : ldg0;
  if (BK == 32) return launch_wl_cols<32, 3>(ctx, f->second, n, nw, S, lds, W0, ldw0, G0, ldg0, B1, ldb1, Gq1, ldq1);
  return launch_wl_cols<16, 4>(ctx, f->second, n, nw, S, lds, W0, ldw0, G0, ldg0, B1, ldb1, Gq1, ldq1);
}

// Host-only self-check of a column-block schedule (no CUDA calls; tests/test_gram_plan.py): the items of every tile
// partition [0, n), every 8x8 block of a tile is owned by exactly one warp, every output entry that is not strictly
// below the diagonal of the Hermitian block is covered by exactly one tile, the skipped tiles lie entirely below it.
// stats: [0] items, [1] busiest CTA cost / mean, [2] tiles, [3] computed tile area / full rectangular area.
int gram_wl_cols_plan_check(int m, int nw, int nprod, int tri_c0, int64_t n, int ncta, int BK, double* stats) {
  WlPlan P;
  std::vector<WlTileOut> tout;
  int rc = plan_schedule_cols(m, nw, nprod, tri_c0, n, ncta, BK, WL_LOAD_PCT, 1, P, tout);
  if (rc) return rc;
  std::vector<double> cta_cost(ncta, 0.0);
  std::vector<int> cover((size_t)nprod * m * nw, 0);
  double area = 0;
  for (int tix = 0; tix < P.ntiles; tix++) {
    const WlTileOut& t = tout[tix];
    if (t.first != P.tile_first[tix] || t.last != P.tile_first[tix + 1] || t.last <= t.first) return 10;
    int64_t r = 0;
    for (int i = t.first; i < t.last; i++) {
      const WlItem& it = P.items[i];
      if (it.r_begin != r || it.r_end <= r || P.item_tile[i] != tix) return 11;
      if (i > t.first && it.r_begin % BK) return 12;
      if (it.c_start < 0 || (int64_t)it.c_start * BK >= it.r_end - it.r_begin) return 18;
      if (it.a_col0 != t.a_col0 || it.b_col0 != (t.b_sel == 2 ? 0 : t.b_col0) || it.a_cols != t.a_cols || it.b_cols != t.b_cols ||
          it.b_sel != t.b_sel || it.same_panel != 0)
        return 19;
      r = it.r_end;
      cta_cost[P.item_cta[i]] += P.tile_cost[tix] * (double)(it.r_end - it.r_begin);
    }
    if (r != n) return 13;
    const int ra = (t.a_cols + 7) / 8, cb = (t.b_cols + 7) / 8;
    int owner[16][16];
    for (auto& row : owner) for (int& v : row) v = 0;
    for (int w = 0; w < 8; w++) {
      const WlWarp& c = P.layouts[P.items[t.first].layout + w];
      if (shape_of(c.mask) != c.shape) return 17;
      for (int i = 0; i < 8; i++)
        for (int j = 0; j < 4; j++)
          if ((c.mask >> (i * 4 + j)) & 1u) {
            const int row = c.transposed ? c.b0 + j : c.a0 + i, col = c.transposed ? c.a0 + i : c.b0 + j;
            if (row >= 16 || col >= 16) return 14;
            owner[row][col]++;
          }
    }
    for (int row = 0; row < 16; row++)
      for (int col = 0; col < 16; col++)
        if (owner[row][col] != ((row < ra && col < cb && (!t.diag || row <= col)) ? 1 : 0)) return 15;
    if (t.diag && (tri_c0 < 0 || t.a_col0 - tri_c0 != t.b_col0)) return 22;
    if (t.b_sel == 2) {   // merged remainder tile: columns [0, msplit) -> product 0, [msplit, 2 msplit) -> product 1
      if (nprod != 2 || t.diag || t.msplit <= 0 || t.b_cols != 2 * t.msplit || t.b_col0 + t.msplit != nw || t.a_col0 < 0 ||
          t.a_col0 + t.a_cols > m)
        return 23;
      for (int i = 0; i < t.a_cols; i++)
        for (int j = 0; j < t.b_cols; j++)
          cover[((size_t)(j / t.msplit) * m + t.a_col0 + i) * nw + t.b_col0 + j % t.msplit]++;
      area += (double)t.a_cols * t.b_cols;
      continue;
    }
    if (t.a_col0 < 0 || t.b_col0 < 0 || t.a_col0 + t.a_cols > m || t.b_col0 + t.b_cols > nw || t.b_sel < 0 || t.b_sel >= nprod)
      return 20;
    for (int i = 0; i < t.a_cols; i++)
      for (int j = 0; j < t.b_cols; j++)
        if (!t.diag || i <= j) cover[((size_t)t.b_sel * m + t.a_col0 + i) * nw + t.b_col0 + j]++;
    area += (double)t.a_cols * t.b_cols * (t.diag ? 136.0 / 256.0 : 1.0);
  }
  for (int q = 0; q < nprod; q++)
    for (int i = 0; i < m; i++)
      for (int j = 0; j < nw; j++) {
        const int c = cover[((size_t)q * m + i) * nw + j];
        const bool below = tri_c0 >= 0 && (i - tri_c0 > j);
        if (c > 1 || (c == 0 && !below)) return 21;
      }
  for (int b = 0; b < ncta; b++)
    for (int i = P.cta_first[b]; i < P.cta_first[b + 1]; i++)
      if (P.item_cta[i] != b) return 16;
  double mx = 0, sum = 0;
  for (double c : cta_cost) { mx = std::max(mx, c); sum += c; }
  if (stats) {
    stats[0] = (double)P.items.size();
    stats[1] = mx / (sum / ncta);
    stats[2] = (double)P.ntiles;
    stats[3] = area / ((double)nprod * m * nw);
  }
  return 0;
}

}  // namespace lb2

// lobpcg_b200/csrc/spmm.cu — K1: block operator application Y = Op X on n x nc column-major block vectors.
//
// Replaces the reference's serial single-vector loop `for j<k: Op->matvec(Op, &X[j*n], &Y[j*n])`
// (src/gram/gram_impl.inc:29-33) for the built-in operators:
//   * matrix-free Dirichlet stencil (1-D/2-D/3-D, 3/5/7 point) with optional diagonal potential, optional
//     z-slab halo planes (multi-GPU row partition: the halo pointers are NVLink peer mappings of the
//     neighbour rank's boundary plane, read directly by the kernel — no separate exchange pass), and the
//     BdG-style 2x2 block variant of config C4;
//   * CSR (int64 row pointers, int32 column indices);
//   * real diagonal (mass matrix B, Jacobi preconditioner T).
// All are HBM-bound: algorithmic traffic 2*n*nc*s (+ matrix once for CSR), see DESIGN.md.
#include "common.cuh"
#include "context.h"
#include "kernels.h"

namespace lb2 {

// =====================================================================================================
// Stencil: CTA = TX x TY points of the xy-plane marching a z-chunk.  Whole (TX+2)x(TY+2) plane tiles
// (with their x/y halo ring, zero-filled outside the grid = Dirichlet) are staged global->shared by
// cp.async (LDGSTS, 16-byte chunks on the aligned interior) into a ring of D+2 planes, D planes ahead
// of the compute, so no thread ever waits on a global load it has just issued.  Per output the thread
// reads its 4 in-plane neighbours and the centre of the next plane from shared memory; z-1 / z / z+1
// centres rotate through registers.  One __syncthreads per plane.  Planes z=-1 and z=gz come from the
// halo pointers (peer memory on a row-partitioned run) or are zero.
// =====================================================================================================
template <typename T, int N> struct alignas(N * sizeof(T)) Pack { T v[N]; };

// EPI = 1: Chebyshev-step epilogue (preconditioner T = p(A)).  X is the search direction d; instead of
// storing A d the kernel finishes the step in registers:  r = r_in - A d;  d' = c1 d + c2 r;  y += d'  (6 block streams
// per step instead of 9 for SpMM + separate update; d' goes to another block because neighbours still read d's halo).
template <typename T, int TXT, int TY, int EPT, int D, bool VECP, bool HAS_POT, bool BDG, int EPI = 0>
__global__ void __launch_bounds__(TXT* TY)
    stencil_kernel(StencilDesc d, const T* __restrict__ X, int64_t ldx, T* __restrict__ Y, int64_t ldy, int ntx,
                   int nty, int zchunk, ChebEpilogue<T> ep) {
  using R = real_t<T>;
  using PK = Pack<T, EPT>;
  constexpr int TX = TXT * EPT;                // tile width in points; each thread owns EPT consecutive x
  constexpr int NT = TXT * TY;
  constexpr int VEC = 16 / sizeof(T);          // elements per 16-byte cp.async chunk
  constexpr int PADL = (VEC > EPT) ? VEC : EPT;  // interior starts 16-byte (and pack) aligned
  constexpr int RS = TX + 2 * PADL;            // row stride (elements)
  constexpr int ROWS = TY + 2;
  constexpr int NS = D + 2;                    // ring slots
  constexpr int PLANE_ELEMS = ROWS * RS;
  // copy slots of one plane tile: VECP: 16-byte interior chunks + 2 single halo elements per row; else elementwise
  constexpr int NCH = VECP ? ROWS * (TX / VEC) : ROWS * (TX + 2);
  constexpr int NHL = VECP ? 2 * ROWS : 0;
  constexpr int NSLOT = (NCH + NHL + NT - 1) / NT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* ring = reinterpret_cast<T*>(smem_raw);    // [NS][ROWS][RS]

  const int tid = threadIdx.x;
  const int tx = tid % TXT, ty = tid / TXT;
  const int bx = blockIdx.x % ntx, by = (blockIdx.x / ntx) % nty, bz = blockIdx.x / (ntx * nty);
  const int x0 = bx * TX, y0 = by * TY;
  const int x = x0 + tx * EPT, y = y0 + ty;
  const int z0 = bz * zchunk, z1 = min(d.gz, z0 + zchunk);
  const int col = blockIdx.y;
  const int64_t plane = (int64_t)d.gx * d.gy;
  const int64_t m = plane * d.gz;              // points per field (BdG: n = 2m)
  const int half = BDG ? blockIdx.z : 0;       // BdG: 0 = u block, 1 = v block
  const int64_t hoff = (int64_t)half * m;
  const R cd = (R)(d.cdiag + d.shift), co = (R)d.coff;
  const T* xcol = X + (int64_t)col * ldx + hoff;
  const T* hlo = d.halo_lo ? (const T*)d.halo_lo + (int64_t)col * d.halo_ld + hoff : nullptr;
  const T* hhi = d.halo_hi ? (const T*)d.halo_hi + (int64_t)col * d.halo_ld + hoff : nullptr;
  const int nplanes = (z1 - z0) + 2;           // planes z0-1 .. z1

  // z-invariant copy descriptors of this thread: offset inside a plane (or -1 = zero fill) and inside a tile
  int goff[NSLOT], soff[NSLOT];
#pragma unroll
  for (int s = 0; s < NSLOT; s++) {
    const int id = tid + s * NT;
    goff[s] = -1;
    soff[s] = -1;
    if (id < NCH) {
      int r, xx, sx;
      if (VECP) { r = id / (TX / VEC); const int c = id % (TX / VEC); xx = x0 + c * VEC; sx = PADL + c * VEC; }
      else { r = id / (TX + 2); const int c = id % (TX + 2); xx = x0 - 1 + c; sx = PADL - 1 + c; }
      const int yy = y0 - 1 + r;
      soff[s] = r * RS + sx;
      if (yy >= 0 && yy < d.gy && xx >= 0 && xx < d.gx) goff[s] = yy * d.gx + xx;
    } else if (id < NCH + NHL) {
      const int h = id - NCH;
      const int r = h >> 1, right = h & 1;
      const int yy = y0 - 1 + r, xx = right ? x0 + TX : x0 - 1;
      soff[s] = r * RS + (right ? PADL + TX : PADL - 1);
      if (yy >= 0 && yy < d.gy && xx >= 0 && xx < d.gx) goff[s] = yy * d.gx + xx;
    }
  }

  auto issue = [&](int q) {
    if (q < nplanes) {
      const int z = z0 - 1 + q;
      const T* src = (z < 0) ? hlo : (z >= d.gz) ? hhi : xcol + (int64_t)z * plane;
      T* dst = ring + (q % NS) * PLANE_ELEMS;
#pragma unroll
      for (int s = 0; s < NSLOT; s++) {
        if (soff[s] < 0) continue;
        const bool ok = (src != nullptr) && goff[s] >= 0;
        const void* g = ok ? (const void*)(src + goff[s]) : (const void*)X;
        if (VECP && tid + s * NT < NCH) cp_async_zfill<16>(dst + soff[s], g, ok ? 16 : 0);
        else cp_async_zfill<(int)sizeof(T)>(dst + soff[s], g, ok ? (int)sizeof(T) : 0);
      }
    }
    cp_async_commit();
  };

#pragma unroll
  for (int q = 0; q <= D; q++) issue(q);

  const int ctr = (ty + 1) * RS + PADL + tx * EPT;   // my first centre inside a plane tile (pack aligned)
  const bool rowin = (y < d.gy);
  const bool full = rowin && (x + EPT <= d.gx);
  const bool any = rowin && (x < d.gx);
  PK prev, cur;
#pragma unroll
  for (int e = 0; e < EPT; e++) prev.v[e] = cur.v[e] = zero<T>();
  const int64_t idx0 = (int64_t)(z0 - 1) * plane + (int64_t)y * d.gx + x;
  T* yp = Y + (int64_t)col * ldy + hoff + idx0;                       // advanced by `plane` per step
  const T* e_rin = EPI ? ep.rin + (int64_t)col * ep.ldrin + hoff + idx0 : nullptr;
  T* e_rout = EPI ? ep.rout + (int64_t)col * ep.ldw + hoff + idx0 : nullptr;
  T* e_dout = EPI ? ep.dout + (int64_t)col * ep.ldw + hoff + idx0 : nullptr;
  const R* pp = HAS_POT ? (const R*)d.potential + idx0 : nullptr;
  const T* op = BDG ? X + (int64_t)col * ldx + (half ? 0 : m) + idx0 : nullptr;
  int slot = 0;
  for (int q = 0; q < nplanes - 1; q++) {
    cp_async_wait<D - 1>();
    __syncthreads();
    issue(q + D + 1);
    const T* pl = ring + slot * PLANE_ELEMS;
    slot = (slot + 1 == NS) ? 0 : slot + 1;
    const T* pn = ring + slot * PLANE_ELEMS;
    if (q == 0) cur = *reinterpret_cast<const PK*>(pl + ctr);   // priming step: plane z0-1 only feeds `prev`
    const PK next = *reinterpret_cast<const PK*>(pn + ctr);
    if (q > 0 && any) {
      const PK up = *reinterpret_cast<const PK*>(pl + ctr - RS);
      const PK dn = *reinterpret_cast<const PK*>(pl + ctr + RS);
      const T left = pl[ctr - 1], right = pl[ctr + EPT];
      PK out;
#pragma unroll
      for (int e = 0; e < EPT; e++) {
        const T xl = (e == 0) ? left : cur.v[e - 1];
        const T xr = (e == EPT - 1) ? right : cur.v[e + 1];
        T nb = add_(add_(xl, xr), add_(up.v[e], dn.v[e]));
        nb = add_(nb, add_(prev.v[e], next.v[e]));
        R dg = cd;
        if (HAS_POT) { if (full || x + e < d.gx) dg += pp[e]; }
        T r = add_(rscale_(cur.v[e], dg), rscale_(nb, co));
        if (BDG) {
          if (full || x + e < d.gx) {
            const T other = op[e];             // coupling: u rows get d v, v rows get conj(d) u
            if constexpr (Sc<T>::cplx) {
              T dd;
              dd.re = (R)d.dre;
              dd.im = half ? (R)(-d.dim) : (R)d.dim;
              fma_(r, dd, other);
            } else {
              r = add_(r, rscale_(other, (R)d.dre));
            }
          }
        }
        out.v[e] = r;
      }
      if (EPI == 0) {
        if (VECP && full) {
          *reinterpret_cast<PK*>(yp) = out;
        } else {
#pragma unroll
          for (int e = 0; e < EPT; e++)
            if (x + e < d.gx) yp[e] = out.v[e];
        }
      } else {   // out = A d at my points, cur = d; yp walks the accumulated solution y
        if (VECP && full) {
          const PK rin = *reinterpret_cast<const PK*>(e_rin);
          PK yv = *reinterpret_cast<const PK*>(yp);
          PK rn, dn;
#pragma unroll
          for (int e = 0; e < EPT; e++) {
            rn.v[e] = sub_(rin.v[e], out.v[e]);
            dn.v[e] = add_(rscale_(cur.v[e], ep.c1), rscale_(rn.v[e], ep.c2));
            yv.v[e] = add_(yv.v[e], dn.v[e]);
          }
          if (ep.write_r) *reinterpret_cast<PK*>(e_rout) = rn;
          *reinterpret_cast<PK*>(e_dout) = dn;
          *reinterpret_cast<PK*>(yp) = yv;
        } else {
#pragma unroll
          for (int e = 0; e < EPT; e++)
            if (x + e < d.gx) {
              const T rn = sub_(e_rin[e], out.v[e]);
              const T dn = add_(rscale_(cur.v[e], ep.c1), rscale_(rn, ep.c2));
              if (ep.write_r) e_rout[e] = rn;
              e_dout[e] = dn;
              yp[e] = add_(yp[e], dn);
            }
        }
      }
    }
    prev = cur;
    cur = next;
    yp += plane;
    if (EPI) { e_rin += plane; e_rout += plane; e_dout += plane; }
    if (HAS_POT) pp += plane;
    if (BDG) op += plane;
  }
  cp_async_wait<0>();
}

template <typename T, int TXT, int TY, int EPT, int D>
static int launch_stencil(lb2_ctx* ctx, const StencilDesc& d, int nc, const T* X, int64_t ldx, T* Y, int64_t ldy,
                          const ChebEpilogue<T>* ep = nullptr) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int PADL = (VEC > EPT) ? VEC : EPT;
  constexpr int TX = TXT * EPT;
  constexpr size_t smem = sizeof(T) * (size_t)(D + 2) * (TY + 2) * (TX + 2 * PADL);
  const int halves = d.bdg ? 2 : 1;
  const int ntx = (d.gx + TX - 1) / TX, nty = (d.gy + TY - 1) / TY;
  // z chunks: long marches amortise the 2 extra planes per chunk; split only to fill the machine
  int zchunk = d.gz;
  const int64_t base_ctas = (int64_t)ntx * nty * nc * halves;
  while (zchunk > 32 && base_ctas * ((d.gz + zchunk - 1) / zchunk) < 16LL * ctx->sm_count) zchunk = (zchunk + 1) / 2;
  const int nz = (d.gz + zchunk - 1) / zchunk;
  const int64_t plane = (int64_t)d.gx * d.gy;
  if (plane >= (1LL << 31)) return -2;
  constexpr int AL = (VEC > EPT) ? VEC : EPT;
  auto al16 = [](const void* p) { return ((uintptr_t)p % 16) == 0; };
  bool vec_ok = (d.gx % AL == 0) && (ldx % AL == 0) && (ldy % AL == 0) && (plane % AL == 0) && al16(X) && al16(Y) &&
                (!d.halo_lo || (al16(d.halo_lo) && d.halo_ld % AL == 0)) &&
                (!d.halo_hi || (al16(d.halo_hi) && d.halo_ld % AL == 0)) && (!d.bdg || (plane * d.gz) % AL == 0);
  if (ep) vec_ok = vec_ok && al16(ep->rin) && al16(ep->rout) && al16(ep->dout) && (ep->ldrin % AL == 0) && (ep->ldw % AL == 0);
  dim3 grid(ntx * nty * nz, nc, halves);
  if (ep) {   // Chebyshev-step epilogue (plain stencil only)
    if (d.bdg) return -2;
#define LB2_STE(VP, HP)                                                                                      \
  {                                                                                                          \
    auto kern = stencil_kernel<T, TXT, TY, EPT, D, VP, HP, false, 1>;                                        \
    if (smem > 48 * 1024)                                                                                    \
      LB2_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    kern<<<grid, TXT * TY, smem, ctx->stream>>>(d, X, ldx, Y, ldy, ntx, nty, zchunk, *ep);                   \
  }
    if (d.potential) { if (vec_ok) LB2_STE(true, true) else LB2_STE(false, true) }
    else { if (vec_ok) LB2_STE(true, false) else LB2_STE(false, false) }
#undef LB2_STE
    ctx->launches++;
    LB2_CUDA_OK(cudaGetLastError());
    return 0;
  }
#define LB2_ST(VP, HP, BD)                                                                                   \
  {                                                                                                          \
    auto kern = stencil_kernel<T, TXT, TY, EPT, D, VP, HP, BD>;                                              \
    if (smem > 48 * 1024)                                                                                    \
      LB2_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    kern<<<grid, TXT * TY, smem, ctx->stream>>>(d, X, ldx, Y, ldy, ntx, nty, zchunk, ChebEpilogue<T>());     \
  }
  const bool hp = d.potential != nullptr;
  if (d.bdg) { if (vec_ok) LB2_ST(true, false, true) else LB2_ST(false, false, true) }
  else if (hp) { if (vec_ok) LB2_ST(true, true, false) else LB2_ST(false, true, false) }
  else { if (vec_ok) LB2_ST(true, false, false) else LB2_ST(false, false, false) }
#undef LB2_ST
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T>
int spmm_stencil(lb2_ctx* ctx, const StencilDesc& d, int nc, const T* X, int64_t ldx, T* Y, int64_t ldy) {
  if (nc <= 0) return 0;
  if (nc > 65535) return -2;
  constexpr int EPT = Sc<T>::cplx ? 1 : 2;   // consecutive x points per thread (16-byte packs for double)
  if (d.gy == 1 && d.gz == 1) return launch_stencil<T, 128, 1, EPT, 2>(ctx, d, nc, X, ldx, Y, ldy);
  return launch_stencil<T, 32, 8, EPT, 4>(ctx, d, nc, X, ldx, Y, ldy);
}

// One Chebyshev step with the stencil as A: r_out = r_in - A D, d_out = c1 D + c2 r_out, Yacc += d_out.
template <typename T>
int spmm_stencil_cheb(lb2_ctx* ctx, const StencilDesc& d, int nc, const T* Din, int64_t ldd, T* Yacc, int64_t ldy,
                      const ChebEpilogue<T>& ep) {
  if (nc <= 0) return 0;
  if (nc > 65535 || d.bdg) return -2;
  constexpr int EPT = Sc<T>::cplx ? 1 : 2;
  if (d.gy == 1 && d.gz == 1) return launch_stencil<T, 128, 1, EPT, 2>(ctx, d, nc, Din, ldd, Yacc, ldy, &ep);
  return launch_stencil<T, 32, 8, EPT, 4>(ctx, d, nc, Din, ldd, Yacc, ldy, &ep);
}

// =====================================================================================================
// CSR: one thread per row, NCOL block-vector columns accumulated in registers per pass; consecutive
// threads own consecutive rows, so for banded matrices the gathers X[col, c] are coalesced across the warp
// and the (col,val) stream of a warp is one contiguous range.  grid.x = row blocks (fast), grid.y = column
// groups, so concurrently resident CTAs share a small column window of X in L2.
// Round-2 sweep on the 128^3 7-point matrix x 128 columns (profiles/kernel_bench_csr_r02.jsonl; algorithmic bytes
// nnz (s + 4) + 8 (n + 1) + 2 n nc s): plain loop, 16 columns per thread 1.29 ms = 53 % of HBM; the pipelined loop
// (next (col, val) pair requested ahead of the current gathers) with 8 columns per thread in the chunked launch order
// 1.17 ms = 59 % (default).  What did NOT help, each measured: staging the matrix stream in shared memory
// (csr_staged_kernel, 1.71 ms), a row-block-major launch order (1.71 ms: short runs scattered over all columns of X,
// poor DRAM page locality), a 2 x 4 / 4 x 2 blocked row order that puts the +-gx / +-gx*gy neighbours into the same CTA
// (1.9 ms, same reason), two couplings per step (2.0-2.4 ms) and the pipelined loop with 16 columns (1.97 ms): more
// gathers in flight evict the lines the +-1 / +-gx neighbours would have hit in L1.  Every output gathers 7 X values of
// which L1 serves the +-1 neighbours only, so ~4x the bytes of X cross the L2 -> L1 path; removing the 1.2 GB of
// matrix re-reads (chunked order) changed nothing by itself.  Getting past ~60 % needs explicit reuse of X (the stencil
// kernel's plane ring), i.e. structure the CSR arrays do not carry.  Matrices that ARE Dirichlet stencils never get
// here (capi.cu: detect_stencil).
// =====================================================================================================
// HALO (row-partitioned matrix): column indices are relative to the first local row; an index below 0 is
// row (index + n_lo) of the lower neighbour's block, an index >= n is row (index - n) of the upper neighbour's block, and
// both are read in place from the neighbour's arena over NVLink (same idea as the stencil's halo planes: no exchange
// pass, no gathered copy).
struct CsrMap {
  int ch = 0;        // != 0: 1-D grid, chunked order with ch row blocks per chunk
  int ncg = 0;       // column groups
  int64_t nrb = 0;   // row blocks
};

template <typename T, int NCOL, int RPT, bool HALO, int PIPE>
__global__ void __launch_bounds__(256)
    csr_kernel(int64_t n, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
               const T* __restrict__ val, int nc, const T* __restrict__ X, int64_t ldx, T* __restrict__ Y,
               int64_t ldy, CsrHalo halo, CsrMap map) {
  // a CTA owns 256 * RPT consecutive rows (thread t: rows base + t + 256 i).  RPT > 1 was measured (r01: 1/2/4/8 rows per
  // thread x 4/8/16/32 columns, 128^3 7-point matrix): no effect, the kernel stays at ~49 % of HBM with 8 or 16 columns
  // per thread — it is bound by L2 sector reads (5 gathered lines per output), not by L1 reuse inside the CTA.
  // map.ch != 0: 1-D grid in CHUNKED order — ch consecutive row blocks for column group 0, the same ch row blocks for
  // column group 1, ... then the next ch row blocks.  A chunk's slice of the (col, val) stream (ch * 256 rows) then stays
  // in L2 while the column groups take their turns (the 2-D column-group-major grid re-reads the whole matrix once per
  // group: 1.4 GB of its 3.7 GB of DRAM reads at 128^3 x 128), while the CTAs that run together still read long
  // contiguous runs of the same few columns of X (a row-block-major order, one row block for all column groups, measured
  // 25 % SLOWER than the 2-D grid: short runs scattered over all columns, poor DRAM page locality).
  int cg = blockIdx.y;
  int64_t blk = blockIdx.x;
  if (map.ch) {
    const int64_t per_chunk = (int64_t)map.ch * map.ncg;
    const int64_t chunk = blockIdx.x / per_chunk, in = blockIdx.x % per_chunk;
    const int64_t left = map.nrb - chunk * map.ch;              // row blocks in this (possibly last, shorter) chunk
    const int chn = (int)min((int64_t)map.ch, left);
    cg = (int)(in / chn);
    blk = chunk * map.ch + in % chn;
  }
  const int c0 = cg * NCOL;
  const int ncol = min(NCOL, nc - c0);
  const T* xb = X + (int64_t)c0 * ldx;
#pragma unroll 1
  for (int i = 0; i < RPT; i++) {
    const int64_t row = (blk * RPT + i) * 256 + threadIdx.x;
    if (row >= n) return;
    T acc[NCOL];
#pragma unroll
    for (int c = 0; c < NCOL; c++) acc[c] = zero<T>();
    const int64_t p0 = rowptr[row], p1 = rowptr[row + 1];
    if constexpr (HALO) {
      const T* xlo = (const T*)halo.lo + (int64_t)c0 * halo.ld_lo + halo.ld_lo;   // row -1 of the local block = last row below
      const T* xhi = (const T*)halo.hi + (int64_t)c0 * halo.ld_hi - n;
      for (int64_t p = p0; p < p1; p++) {
        const int64_t cj = col[p];
        const T v = val[p];
        const T* src = cj < 0 ? xlo + cj : (cj >= n ? xhi + cj : xb + cj);
        const int64_t ld = cj < 0 ? halo.ld_lo : (cj >= n ? halo.ld_hi : ldx);
#pragma unroll
        for (int c = 0; c < NCOL; c++)
          if (c < ncol) fma_(acc[c], v, src[(int64_t)c * ld]);
      }
    } else if (ncol == NCOL && PIPE == 1) {
      // software pipeline: the (col, val) pair of the NEXT coupling is requested before the gathers of the current one,
      // so a row costs one memory latency per coupling instead of two dependent ones (col[p] -> X[col[p], :])
      int64_t p = p0;
      int32_t cn = 0;
      T vn = zero<T>();
      if (p < p1) { cn = col[p]; vn = val[p]; }
      while (p < p1) {
        const int64_t cj = cn;
        const T v = vn;
        ++p;
        if (p < p1) { cn = col[p]; vn = val[p]; }
        T x[NCOL];
#pragma unroll
        for (int c = 0; c < NCOL; c++) x[c] = xb[cj + (int64_t)c * ldx];
#pragma unroll
        for (int c = 0; c < NCOL; c++) fma_(acc[c], v, x[c]);
      }
    } else if (ncol == NCOL && PIPE == 2) {
      // two couplings per step (2 x NCOL gathers in flight per thread), the next pair requested ahead of them
      int64_t p = p0;
      int32_t ca = 0, cb = 0;
      T va = zero<T>(), vb = zero<T>();
      if (p < p1) { ca = col[p]; va = val[p]; }
      if (p + 1 < p1) { cb = col[p + 1]; vb = val[p + 1]; }
      while (p < p1) {
        const int64_t a = ca, b = cb;
        const T wa = va, wb = vb;
        const bool two = p + 1 < p1;
        p += 2;
        if (p < p1) { ca = col[p]; va = val[p]; }
        if (p + 1 < p1) { cb = col[p + 1]; vb = val[p + 1]; }
        T x[NCOL], y[NCOL];
#pragma unroll
        for (int c = 0; c < NCOL; c++) x[c] = xb[a + (int64_t)c * ldx];
        if (two) {
#pragma unroll
          for (int c = 0; c < NCOL; c++) y[c] = xb[b + (int64_t)c * ldx];
        }
#pragma unroll
        for (int c = 0; c < NCOL; c++) fma_(acc[c], wa, x[c]);
        if (two) {
#pragma unroll
          for (int c = 0; c < NCOL; c++) fma_(acc[c], wb, y[c]);
        }
      }
    } else if (ncol == NCOL) {
      for (int64_t p = p0; p < p1; p++) {
        const int64_t cj = col[p];
        const T v = val[p];
#pragma unroll
        for (int c = 0; c < NCOL; c++) fma_(acc[c], v, xb[cj + (int64_t)c * ldx]);
      }
    } else {
      for (int64_t p = p0; p < p1; p++) {
        const int64_t cj = col[p];
        const T v = val[p];
#pragma unroll
        for (int c = 0; c < NCOL; c++)
          if (c < ncol) fma_(acc[c], v, xb[cj + (int64_t)c * ldx]);
      }
    }
    T* yb = Y + (int64_t)c0 * ldy + row;
#pragma unroll
    for (int c = 0; c < NCOL; c++)
      if (c < ncol) yb[(int64_t)c * ldy] = acc[c];
  }
}

// =====================================================================================================
// CSR with a shared-memory X window ("banded" matrices: most couplings of row r lie in [r - H, r + H]).
// A CTA owns RB = 512 consecutive rows and NCOL columns of the block vector; it stages X[r0 - H, r0 + RB + H) of its
// columns in shared memory with coalesced loads (each X element is fetched from L2 once per CTA instead of once per
// coupling), every thread then walks the nonzeros of its rows: a column index inside the window reads shared memory,
// anything else (far couplings, e.g. the +-gx*gy neighbours of a 3-D stencil) is gathered from global memory as before.
// blockIdx.x = row block * ncg + column group: the CTAs of one row block are neighbours in launch order, so its
// (col, val) stream comes from DRAM once and from L2 for the other column groups.
// H is chosen per matrix on the host (capi.cu: csr_window_halo) from the histogram of |col - row|.
// =====================================================================================================
constexpr int CSRW_RB = 512;
template <typename T, int NCOL>
__global__ void __launch_bounds__(256)
    csr_win_kernel(int64_t n, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                   const T* __restrict__ val, int nc, const T* __restrict__ X, int64_t ldx, T* __restrict__ Y,
                   int64_t ldy, int H, int ncg) {
  extern __shared__ __align__(16) unsigned char csrw_smem[];
  T* win = reinterpret_cast<T*>(csrw_smem);
  const int WL = CSRW_RB + 2 * H;
  const int cg = blockIdx.x % ncg;
  const int64_t rb = blockIdx.x / ncg;
  const int c0 = cg * NCOL;
  const int ncol = min(NCOL, nc - c0);
  const int64_t r0 = rb * CSRW_RB, wlo = r0 - H;
  const T* xb = X + (int64_t)c0 * ldx;
  // window fill with cp.async (zero-filling outside the matrix / beyond the last column): every copy is in flight before the
  // first one lands — plain loads + stores serialise on the load latency (in-order issue), which made the first version of
  // this kernel slower than the plain one
  for (int c = 0; c < NCOL; c++) {
    T* w = win + (size_t)c * WL;
    const T* xc = xb + (int64_t)c * ldx;
    for (int i = threadIdx.x; i < WL; i += 256) {
      const int64_t row = wlo + i;
      const bool ok = c < ncol && row >= 0 && row < n;
      cp_async_zfill<sizeof(T)>(w + i, ok ? xc + row : X, ok ? (int)sizeof(T) : 0);
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
#pragma unroll 1
  for (int rr = 0; rr < CSRW_RB / 256; rr++) {
    const int64_t row = r0 + rr * 256 + threadIdx.x;
    if (row >= n) break;
    T acc[NCOL];
#pragma unroll
    for (int c = 0; c < NCOL; c++) acc[c] = zero<T>();
    const int64_t p0 = rowptr[row], p1 = rowptr[row + 1];
    for (int64_t p = p0; p < p1; p++) {
      const int64_t cj = col[p];
      const T v = val[p];
      const int64_t d = cj - wlo;
      if ((uint64_t)d < (uint64_t)WL) {
        const T* sp = win + d;
#pragma unroll
        for (int c = 0; c < NCOL; c++) fma_(acc[c], v, sp[(size_t)c * WL]);
      } else {
        const T* gp = xb + cj;
#pragma unroll
        for (int c = 0; c < NCOL; c++)
          if (c < ncol) fma_(acc[c], v, gp[(int64_t)c * ldx]);
      }
    }
    T* yb = Y + (int64_t)c0 * ldy + row;
#pragma unroll
    for (int c = 0; c < NCOL; c++)
      if (c < ncol) yb[(int64_t)c * ldy] = acc[c];
  }
}

// =====================================================================================================
// CSR with the matrix stream staged in shared memory (default general kernel).
// A CTA owns RB = 256 / LPR consecutive rows and NCOL columns of the block vector.  The (col, val) entries of those rows
// are ONE contiguous range of the CSR arrays: the CTA copies it to shared memory with 16-byte cp.async (fully coalesced,
// every copy in flight at once), and only then walks its rows.  That removes the dependent global-load chain of the plain
// kernel (col[p] -> X[col[p], :]: two DRAM/L2 latencies per nonzero with nothing else in flight, ncu: long_scoreboard
// 21 warps per issue at 52 % of DRAM throughput) — the gathers of several nonzeros are issued back to back — and the row
// walk reads the indices at shared-memory latency.  blockIdx.x = row block * ncg + column group: the CTAs of a row block
// run together, so the matrix stream comes from DRAM once and from L2 for the other column groups (the plain kernel's
// column-group-major grid re-read it nc / NCOL times: 1.4 GB of its 3.7 GB of DRAM reads at 128^3 x 128), and the rows
// X[r +- bandwidth] that later row blocks need again stay in L2 for all columns (2 * bandwidth * nc * s bytes).
// LPR lanes share a row (rows with many couplings): lane l takes entries l, l + LPR, ... and the partial sums are
// combined with shuffles.  Rows longer than the staging buffer are handled by walking the range in several passes.
// =====================================================================================================
constexpr int CSRS_CAP = 4096;   // staged entries per pass

template <typename R>
__device__ __forceinline__ R shfl_xor_(R v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
template <typename R>
__device__ __forceinline__ Cx<R> shfl_xor_(Cx<R> v, int o) {
  return Cx<R>{__shfl_xor_sync(0xffffffffu, v.re, o), __shfl_xor_sync(0xffffffffu, v.im, o)};
}

template <typename T, int NCOL, int LPR, bool HALO>
__global__ void __launch_bounds__(256)
    csr_staged_kernel(int64_t n, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                      const T* __restrict__ val, int64_t nnz, int nc, const T* __restrict__ X, int64_t ldx,
                      T* __restrict__ Y, int64_t ldy, CsrHalo halo, int ncg) {
  constexpr int RB = 256 / LPR;
  constexpr int VPU = 16 / (int)sizeof(T);   // values per 16-byte copy
  extern __shared__ __align__(16) unsigned char csrs_smem[];
  T* sval = reinterpret_cast<T*>(csrs_smem);
  int32_t* scol = reinterpret_cast<int32_t*>(sval + CSRS_CAP);
  const int cg = blockIdx.x % ncg;
  const int64_t r0 = (int64_t)(blockIdx.x / ncg) * RB;
  const int c0 = cg * NCOL;
  const int ncol = min(NCOL, nc - c0);
  const int sub = threadIdx.x % LPR;
  const int64_t row = r0 + threadIdx.x / LPR;
  int64_t p0 = 0, p1 = 0;
  if (row < n) { p0 = rowptr[row]; p1 = rowptr[row + 1]; }
  const int64_t pb = rowptr[r0] & ~(int64_t)3;      // 16-byte aligned start of the copies
  const int64_t pe = rowptr[min(r0 + RB, n)];
  const T* xb = X + (int64_t)c0 * ldx;
  T acc[NCOL];
#pragma unroll
  for (int c = 0; c < NCOL; c++) acc[c] = zero<T>();
  for (int64_t cb = pb; cb < pe; cb += CSRS_CAP) {
    const int cnt = (int)min((int64_t)CSRS_CAP, pe - cb);
    if (cb != pb) __syncthreads();   // every lane is done with the previous pass
    for (int u = threadIdx.x; u * 4 < cnt; u += 256) {
      const int64_t left = nnz - (cb + 4 * u);
      cp_async_zfill<16>(scol + 4 * u, col + cb + 4 * u, left >= 4 ? 16 : (int)left * 4);
    }
    for (int u = threadIdx.x; u * VPU < cnt; u += 256) {
      const int64_t left = nnz - (cb + VPU * u);
      cp_async_zfill<16>(sval + VPU * u, val + cb + VPU * u, left >= VPU ? 16 : (int)left * (int)sizeof(T));
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int lo = (int)(max(p0, cb) - cb) + sub;
    const int hi = (int)(min(p1, cb + cnt) - cb);
    if constexpr (HALO) {
      const T* xlo = (const T*)halo.lo + (int64_t)c0 * halo.ld_lo + halo.ld_lo;   // row -1 of the local block = last row below
      const T* xhi = (const T*)halo.hi + (int64_t)c0 * halo.ld_hi - n;
      for (int q = lo; q < hi; q += LPR) {
        const int64_t cj = scol[q];
        const T v = sval[q];
        const T* src = cj < 0 ? xlo + cj : (cj >= n ? xhi + cj : xb + cj);
        const int64_t ld = cj < 0 ? halo.ld_lo : (cj >= n ? halo.ld_hi : ldx);
#pragma unroll
        for (int c = 0; c < NCOL; c++)
          if (c < ncol) fma_(acc[c], v, src[(int64_t)c * ld]);
      }
    } else if (ncol == NCOL) {
#pragma unroll 4
      for (int q = lo; q < hi; q += LPR) {
        const T* src = xb + scol[q];
        const T v = sval[q];
#pragma unroll
        for (int c = 0; c < NCOL; c++) fma_(acc[c], v, src[(int64_t)c * ldx]);
      }
    } else {
      for (int q = lo; q < hi; q += LPR) {
        const T* src = xb + scol[q];
        const T v = sval[q];
#pragma unroll
        for (int c = 0; c < NCOL; c++)
          if (c < ncol) fma_(acc[c], v, src[(int64_t)c * ldx]);
      }
    }
  }
  if constexpr (LPR > 1) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
      for (int c = 0; c < NCOL; c++) acc[c] = add_(acc[c], shfl_xor_(acc[c], o));
  }
  if (row < n && sub == 0) {
    T* yb = Y + (int64_t)c0 * ldy + row;
#pragma unroll
    for (int c = 0; c < NCOL; c++)
      if (c < ncol) yb[(int64_t)c * ldy] = acc[c];
  }
}

template <typename T, int NCOL, int LPR>
static int launch_csr_staged(lb2_ctx* ctx, int64_t n, const int64_t* rowptr, const int32_t* col, const T* val,
                             int64_t nnz, int nc, const T* X, int64_t ldx, T* Y, int64_t ldy, const CsrHalo* halo) {
  constexpr int RB = 256 / LPR;
  const int ncg = (nc + NCOL - 1) / NCOL;
  const int64_t nrb = (n + RB - 1) / RB;
  if (nrb * ncg > 0x7fffffffLL) return -100;
  constexpr size_t smem = (sizeof(T) + sizeof(int32_t)) * (size_t)CSRS_CAP;
  const CsrHalo h = halo ? *halo : CsrHalo{};
  if (halo) {
    auto k = csr_staged_kernel<T, NCOL, LPR, true>;
    LB2_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)(nrb * ncg), 256, smem, ctx->stream>>>(n, rowptr, col, val, nnz, nc, X, ldx, Y, ldy, h, ncg);
  } else {
    auto k = csr_staged_kernel<T, NCOL, LPR, false>;
    LB2_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)(nrb * ncg), 256, smem, ctx->stream>>>(n, rowptr, col, val, nnz, nc, X, ldx, Y, ldy, h, ncg);
  }
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T>
int spmm_csr_window(lb2_ctx* ctx, int64_t n, const int64_t* rowptr, const int32_t* col, const T* val, int nc,
                    const T* X, int64_t ldx, T* Y, int64_t ldy, int H) {
  if (n <= 0 || nc <= 0) return 0;
  constexpr int NCOL = sizeof(T) >= 16 ? 4 : (sizeof(T) == 8 ? 8 : 16);
  const int ncg = (nc + NCOL - 1) / NCOL;
  const int64_t nrb = (n + CSRW_RB - 1) / CSRW_RB;
  if (nrb * ncg > 0x7fffffffLL) return -100;
  const size_t smem = sizeof(T) * (size_t)NCOL * (CSRW_RB + 2 * H);
  auto k = csr_win_kernel<T, NCOL>;
  LB2_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<(unsigned)(nrb * ncg), 256, smem, ctx->stream>>>(n, rowptr, col, val, nc, X, ldx, Y, ldy, H, ncg);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T>
int spmm_csr(lb2_ctx* ctx, int64_t n, const int64_t* rowptr, const int32_t* col, const T* val, int nc,
             const T* X, int64_t ldx, T* Y, int64_t ldy, const CsrHalo* halo, int64_t nnz) {
  if (n <= 0 || nc <= 0) return 0;
  // 64 bytes of a block-vector row per thread (8 doubles) with the pipelined loop: r02 sweep at 128^3 x 128, f64 — plain loop
  // 16 columns 1.29 ms, pipelined 16 columns 1.97 ms (more gathers in flight thrash L1), pipelined 8 columns 1.17 ms
  int ncol = ctx->spmm_cols ? ctx->spmm_cols : (ctx->csr_pipe == 1 ? 64 / (int)sizeof(T) : 16);
  if (nnz >= 0 && ctx->csr_staged > 0 && (uintptr_t)col % 16 == 0 && (uintptr_t)val % 16 == 0) {
    // staged kernel: lanes per row from the mean row length (1 for stencil-like rows, 4 / 16 for FEM-like and denser rows)
    const double mean = (double)nnz / (double)n;
    const int lpr = ctx->csr_lpr > 0 ? ctx->csr_lpr : (mean <= 12.0 ? 1 : (mean <= 48.0 ? 4 : 16));
    constexpr int NC = sizeof(T) >= 16 ? 4 : 8;
    constexpr int NCW = sizeof(T) >= 16 ? 8 : 16;
    int rc;
    if (lpr >= 16) rc = launch_csr_staged<T, NC, 16>(ctx, n, rowptr, col, val, nnz, nc, X, ldx, Y, ldy, halo);
    else if (lpr >= 4) rc = launch_csr_staged<T, NC, 4>(ctx, n, rowptr, col, val, nnz, nc, X, ldx, Y, ldy, halo);
    else if (ncol <= 8 || nc <= NC) rc = launch_csr_staged<T, NC, 1>(ctx, n, rowptr, col, val, nnz, nc, X, ldx, Y, ldy, halo);
    else rc = launch_csr_staged<T, NCW, 1>(ctx, n, rowptr, col, val, nnz, nc, X, ldx, Y, ldy, halo);
    if (rc != -100) return rc;
  }
  const CsrHalo h = halo ? *halo : CsrHalo{};
  CsrMap map{};
#define LB2_CSR2(NC, RP)                                                                                        \
  do {                                                                                                          \
    const int64_t nrb = (n + 256 * RP - 1) / (256 * RP);                                                        \
    const int ncg = (nc + NC - 1) / NC;                                                                         \
    dim3 grid((unsigned)nrb, ncg);                                                                              \
    if (ctx->csr_order > 0 && ncg > 1 && nrb * ncg <= 0x7fffffffLL) {                                           \
      map.ch = ctx->csr_order; map.ncg = ncg; map.nrb = nrb; grid = dim3((unsigned)(nrb * ncg));                \
    }                                                                                                           \
    if (halo)                                                                                                   \
      csr_kernel<T, NC, RP, true, 0><<<grid, 256, 0, ctx->stream>>>(n, rowptr, col, val, nc, X, ldx, Y, ldy, h, map);   \
    else if (ctx->csr_pipe == 1)                                                                                \
      csr_kernel<T, NC, RP, false, 1><<<grid, 256, 0, ctx->stream>>>(n, rowptr, col, val, nc, X, ldx, Y, ldy, h, map);  \
    else if (ctx->csr_pipe == 2)                                                                                \
      csr_kernel<T, NC, RP, false, 2><<<grid, 256, 0, ctx->stream>>>(n, rowptr, col, val, nc, X, ldx, Y, ldy, h, map);  \
    else                                                                                                        \
      csr_kernel<T, NC, RP, false, 0><<<grid, 256, 0, ctx->stream>>>(n, rowptr, col, val, nc, X, ldx, Y, ldy, h, map);  \
  } while (0)
#define LB2_CSR(NC) LB2_CSR2(NC, 1)
  if (nc <= 4 || ncol <= 4) LB2_CSR(4);
  else if (nc <= 8 || ncol <= 8) LB2_CSR(8);
  else if (ncol <= 16 || sizeof(T) >= 16) LB2_CSR(16);
  else LB2_CSR(32);
#undef LB2_CSR
#undef LB2_CSR2
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// =====================================================================================================
// Real diagonal operator: Y[:,j] = d .* X[:,j]
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
    diag_kernel(int64_t n, const real_t<T>* __restrict__ dg, int nc, const T* __restrict__ X, int64_t ldx,
                T* __restrict__ Y, int64_t ldy, int cols_per_cta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const real_t<T> f = dg[i];
  const int c0 = blockIdx.y * cols_per_cta;
  const int c1 = min(nc, c0 + cols_per_cta);
  for (int c = c0; c < c1; c++) Y[i + (int64_t)c * ldy] = rscale_(X[i + (int64_t)c * ldx], f);
}
template <typename T>
int spmm_diag(lb2_ctx* ctx, int64_t n, const real_t<T>* d, int nc, const T* X, int64_t ldx, T* Y, int64_t ldy) {
  if (n <= 0 || nc <= 0) return 0;
  const int cpc = 8;
  dim3 grid((unsigned)((n + 255) / 256), (nc + cpc - 1) / cpc);
  diag_kernel<T><<<grid, 256, 0, ctx->stream>>>(n, d, nc, X, ldx, Y, ldy, cpc);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

#define LB2_INST(T)                                                                                       \
  template int spmm_stencil<T>(lb2_ctx*, const StencilDesc&, int, const T*, int64_t, T*, int64_t);         \
  template int spmm_stencil_cheb<T>(lb2_ctx*, const StencilDesc&, int, const T*, int64_t, T*, int64_t, const ChebEpilogue<T>&); \
  template int spmm_csr<T>(lb2_ctx*, int64_t, const int64_t*, const int32_t*, const T*, int, const T*, int64_t, T*, int64_t, const CsrHalo*, int64_t); \
  template int spmm_csr_window<T>(lb2_ctx*, int64_t, const int64_t*, const int32_t*, const T*, int, const T*, int64_t, T*, int64_t, int); \
  template int spmm_diag<T>(lb2_ctx*, int64_t, const real_t<T>*, int, const T*, int64_t, T*, int64_t);
LB2_INST(float)
LB2_INST(double)
LB2_INST(c32)
LB2_INST(c64)
#undef LB2_INST

}  // namespace lb2

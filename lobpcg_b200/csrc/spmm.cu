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
// Stencil: CTA = TX x TY points of the xy-plane, marches a z-chunk keeping (z-1, z, z+1) in registers;
// the current plane goes through a double-buffered shared tile for the x/y neighbours.  Each thread
// carries NCOL columns of the block vector => NCOL independent load streams in flight.
// =====================================================================================================
template <typename T, int TX, int TY, int NCOL>
__global__ void __launch_bounds__(TX* TY)
    stencil_kernel(StencilDesc d, int nc, const T* __restrict__ X, int64_t ldx, T* __restrict__ Y,
                   int64_t ldy, int ntx, int nty, int zchunk) {
  using R = real_t<T>;
  __shared__ T tile[2][NCOL][TY + 2][TX + 2];
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int bx = blockIdx.x % ntx, by = (blockIdx.x / ntx) % nty, bz = blockIdx.x / (ntx * nty);
  const int x = bx * TX + tx, y = by * TY + ty;
  const int z0 = bz * zchunk, z1 = min(d.gz, z0 + zchunk);
  const int c0 = blockIdx.y * NCOL;
  const bool inside = (x < d.gx) && (y < d.gy);
  const int64_t plane = (int64_t)d.gx * d.gy;
  const int64_t m = plane * d.gz;                 // points per field (BdG: n = 2m)
  const int half = d.bdg ? blockIdx.z : 0;        // BdG: 0 = u block, 1 = v block
  const int64_t hoff = (int64_t)half * m;
  const int64_t pxy = (int64_t)y * d.gx + x;
  const R cd = (R)(d.cdiag + d.shift), co = (R)d.coff;
  const R* pot = (const R*)d.potential;
  const T* hlo = (const T*)d.halo_lo;
  const T* hhi = (const T*)d.halo_hi;

  const T* xc[NCOL];
  T* yc[NCOL];
  bool cv[NCOL];
#pragma unroll
  for (int c = 0; c < NCOL; c++) {
    cv[c] = (c0 + c < nc);
    const int cc = cv[c] ? c0 + c : c0;
    xc[c] = X + (int64_t)cc * ldx + hoff;
    yc[c] = Y + (int64_t)cc * ldy + hoff;
  }

  auto ld_plane = [&](int c, int z) -> T {
    // value at (x,y,z) of column c; z may be -1 or gz (halo / Dirichlet zero)
    if (!inside || !cv[c]) return zero<T>();
    if (z < 0) return hlo ? hlo[(int64_t)(c0 + c) * d.halo_ld + hoff + pxy] : zero<T>();
    if (z >= d.gz) return hhi ? hhi[(int64_t)(c0 + c) * d.halo_ld + hoff + pxy] : zero<T>();
    return xc[c][(int64_t)z * plane + pxy];
  };
  auto ld_xy = [&](int c, int xx, int yy, int z) -> T {
    if (xx < 0 || xx >= d.gx || yy < 0 || yy >= d.gy || !cv[c]) return zero<T>();
    return xc[c][(int64_t)z * plane + (int64_t)yy * d.gx + xx];
  };

  T prev[NCOL], cur[NCOL], next[NCOL];
#pragma unroll
  for (int c = 0; c < NCOL; c++) {
    prev[c] = ld_plane(c, z0 - 1);
    cur[c] = ld_plane(c, z0);
    next[c] = ld_plane(c, z0 + 1);
  }

  for (int z = z0; z < z1; z++) {
    const int buf = (z - z0) & 1;
    T nn[NCOL];
#pragma unroll
    for (int c = 0; c < NCOL; c++) {
      nn[c] = (z + 2 <= z1) ? ld_plane(c, z + 2) : zero<T>();  // prefetch z+2 (z1 may be gz => halo)
      tile[buf][c][ty + 1][tx + 1] = cur[c];
      if (tx == 0) tile[buf][c][ty + 1][0] = ld_xy(c, x - 1, y, z);
      if (tx == TX - 1) tile[buf][c][ty + 1][TX + 1] = ld_xy(c, x + 1, y, z);
      if (ty == 0) tile[buf][c][0][tx + 1] = ld_xy(c, x, y - 1, z);
      if (ty == TY - 1) tile[buf][c][TY + 1][tx + 1] = ld_xy(c, x, y + 1, z);
    }
    __syncthreads();
    if (inside) {
      const int64_t idx = (int64_t)z * plane + pxy;
      const R dg = cd + (pot ? pot[idx] : R(0));
#pragma unroll
      for (int c = 0; c < NCOL; c++) {
        if (!cv[c]) continue;
        T nb = add_(add_(tile[buf][c][ty + 1][tx], tile[buf][c][ty + 1][tx + 2]),
                    add_(tile[buf][c][ty][tx + 1], tile[buf][c][ty + 2][tx + 1]));
        nb = add_(nb, add_(prev[c], next[c]));
        T r = add_(rscale_(cur[c], dg), rscale_(nb, co));
        if (d.bdg) {
          // coupling block: u rows get d * v, v rows get conj(d) * u
          const T other = X[(int64_t)(c0 + c) * ldx + (half ? 0 : m) + idx];
          if constexpr (Sc<T>::cplx) {
            T dd;
            dd.re = (R)d.dre;
            dd.im = half ? (R)(-d.dim) : (R)d.dim;
            fma_(r, dd, other);
          } else {
            r = add_(r, rscale_(other, (R)d.dre));
          }
        }
        yc[c][idx] = r;
      }
    }
#pragma unroll
    for (int c = 0; c < NCOL; c++) {
      prev[c] = cur[c];
      cur[c] = next[c];
      next[c] = nn[c];
    }
  }
}

template <typename T>
int spmm_stencil(lb2_ctx* ctx, const StencilDesc& d, int nc, const T* X, int64_t ldx, T* Y, int64_t ldy) {
  if (nc <= 0) return 0;
  constexpr int NCOL = (sizeof(T) >= 16) ? 2 : 4;
  const int halves = d.bdg ? 2 : 1;
  if (d.gy == 1 && d.gz == 1) {
    constexpr int TX = 128, TY = 1;
    const int ntx = (d.gx + TX - 1) / TX;
    dim3 grid(ntx, (nc + NCOL - 1) / NCOL, halves);
    stencil_kernel<T, TX, TY, NCOL><<<grid, TX * TY, 0, ctx->stream>>>(d, nc, X, ldx, Y, ldy, ntx, 1, 1);
  } else {
    constexpr int TX = 32, TY = 8;
    const int ntx = (d.gx + TX - 1) / TX, nty = (d.gy + TY - 1) / TY;
    // z chunks: enough CTAs to fill the machine, but long enough marches to amortise the 3-plane prologue
    int zchunk = d.gz;
    const int64_t base_ctas = (int64_t)ntx * nty * ((nc + NCOL - 1) / NCOL) * halves;
    while (zchunk > 16 && base_ctas * ((d.gz + zchunk - 1) / zchunk) < 8LL * ctx->sm_count) zchunk = (zchunk + 1) / 2;
    const int nz = (d.gz + zchunk - 1) / zchunk;
    dim3 grid(ntx * nty * nz, (nc + NCOL - 1) / NCOL, halves);
    stencil_kernel<T, TX, TY, NCOL><<<grid, TX * TY, 0, ctx->stream>>>(d, nc, X, ldx, Y, ldy, ntx, nty, zchunk);
  }
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// =====================================================================================================
// CSR: one thread per row, NCOL block-vector columns accumulated in registers per pass; consecutive
// threads own consecutive rows, so for banded matrices the gathers X[col, c] are coalesced across the warp
// and the (col,val) stream of a warp is one contiguous range.  grid.x = row blocks (fast), grid.y = column
// groups, so concurrently resident CTAs share a small column window of X in L2.
// =====================================================================================================
template <typename T, int NCOL>
__global__ void __launch_bounds__(128)
    csr_kernel(int64_t n, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
               const T* __restrict__ val, int nc, const T* __restrict__ X, int64_t ldx, T* __restrict__ Y,
               int64_t ldy) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = blockIdx.y * NCOL;
  if (row >= n) return;
  const int ncol = min(NCOL, nc - c0);
  T acc[NCOL];
#pragma unroll
  for (int c = 0; c < NCOL; c++) acc[c] = zero<T>();
  const int64_t p0 = rowptr[row], p1 = rowptr[row + 1];
  const T* xb = X + (int64_t)c0 * ldx;
  if (ncol == NCOL) {
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

template <typename T>
int spmm_csr(lb2_ctx* ctx, int64_t n, const int64_t* rowptr, const int32_t* col, const T* val, int nc,
             const T* X, int64_t ldx, T* Y, int64_t ldy) {
  if (n <= 0 || nc <= 0) return 0;
  int ncol = ctx->spmm_cols ? ctx->spmm_cols : 16;
  const unsigned gx = (unsigned)((n + 127) / 128);
#define LB2_CSR(NC)                                                                                   \
  csr_kernel<T, NC><<<dim3(gx, (nc + NC - 1) / NC), 128, 0, ctx->stream>>>(n, rowptr, col, val, nc, X, ldx, Y, ldy)
  if (nc <= 4 || ncol <= 4) LB2_CSR(4);
  else if (nc <= 8 || ncol <= 8) LB2_CSR(8);
  else if (ncol <= 16 || sizeof(T) >= 16) LB2_CSR(16);
  else LB2_CSR(32);
#undef LB2_CSR
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
  template int spmm_csr<T>(lb2_ctx*, int64_t, const int64_t*, const int32_t*, const T*, int, const T*, int64_t, T*, int64_t); \
  template int spmm_diag<T>(lb2_ctx*, int64_t, const real_t<T>*, int, const T*, int64_t, T*, int64_t);
LB2_INST(float)
LB2_INST(double)
LB2_INST(c32)
LB2_INST(c64)
#undef LB2_INST

}  // namespace lb2

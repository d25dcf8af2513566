// Tables and workspaces of the batched OSQP-style QP solver.
#pragma once
#include <stdint.h>

#include "plm_types.h"

namespace plm {

// The symmetric inverse S^-1 of a stage block (s x s) is stored by cyclic diagonals: row j of the stored array M
// (j = 0 .. s/2, s doubles each) holds M[j][k] = S^-1[k][(k + j) mod s].  Every unordered pair {a, b} appears exactly
// once (for even s the second half of row s/2 repeats the first and is stored as zeros), so the array has the size of
// the packed triangle, and the product out = S^-1 in needs no masks:
//   out[k] = M[0][k] in[k] + sum_{j >= 1} ( M[j][k] in[(k + j) mod s] + M[j][(k - j) mod s] in[(k - j) mod s] ).
// Element (r, c), r >= c, lives at plm_sinv_index(s, r, c).  The array is streamed through shared memory in panels of
// consecutive rows j.
#define PLM_PANEL_DOUBLES 2192        // 17 KB: three panels per B2G stage in either sweep, four CTAs per SM
// schedule step: {offset in the instance's factor (doubles), doubles to copy (even), first row, end row, stage,
// flags (bit 0 direction: 0 forward / 1 backward, bit 1 first panel of the stage, bit 2 last panel of the stage,
// bit 3 coupling panel (dense integrator rows: rows [first, end) of the zero-filled block diag(rho n) A_int of node
// stage - 1, streamed from the Gc workspace ahead of the stage's own panels), bit 4 first coupling panel of the stage;
// forward steps: bits 8.. size of the previous stage),
// offset of the copy inside the stage block, stage size | x_off << 8}
// A backward step streams columns [first, end) of B_i instead (all s rows of each); the seventh int is the column stride sp.
#define PLM_SCHED_INTS 8

#if defined(__CUDACC__)
#define PLM_QP_HD __host__ __device__
#else
#define PLM_QP_HD
#endif
PLM_QP_HD inline int plm_sinv_rows(int s) { return (s >> 1) + 1; }
PLM_QP_HD inline int plm_sinv_index(int s, int r, int c) {      // r >= c
  const int d = r - c;
  return d <= (s >> 1) ? d * s + c : (s - d) * s + r;
}

// Per node-type local sparsity tables (offsets into one int16 pool).  Local columns of a node block are
// [0, s) = this stage (DX_i | U_i) and [s, s + ndx) = DX_{i+1}; local rows are the node's rows in g order.
struct QpTypeIdx {
  int32_t rptr;   // [nrows+1]  CSR row pointers (positions in the node's value block)
  int32_t ccol;   // [nnz]      local column of each value
  int32_t cptr;   // [ncols+1]  CSC column pointers
  int32_t cpos;   // [nnz]      value position of each CSC entry
  int32_t crow;   // [nnz]      local row of each CSC entry
  int32_t ncols;  // s + ndx
  int32_t s;      // stage size ndx + nu
  // general coupling (QpLayout::general_coupling): the rows with entries in DX_{i+1}
  int32_t gc_rows;   // [ncoup]  local row of coupling row q
  int32_t gc_rowq;   // [nrows]  q of a local row, -1 if the row does not touch DX_{i+1}
  int32_t ncoup;
};

struct QpLayout {
  QpTypeIdx type[4];
  // flat (whole-problem) index tables, offsets into the int32 pool: CSR row pointers, CSC column pointers,
  // CSC source positions (into the CSR value order); and into the int16 pool: CSR global columns, CSC global rows
  int32_t f_rptr, f_tptr, f_tsrc, f_rcol, f_trow;
  int32_t f_rperm, f_cperm;            // int16 pool: rows / columns sorted by descending length (balanced warps)
  int32_t sparse_coupling;             // every integrator row has at most 4 entries in its own stage (all but whole_body_aba / centroidal_vel)
  // general_coupling = 0: the first ndx rows of every node are the integrator rows, each with exactly one entry in
  // DX_{i+1} (its last one, at next-column r), and no other row touches DX_{i+1}.  1: any row may touch DX_{i+1}
  // (whole_body_rnea without acceleration inputs: the RNEA rows depend on dv_{i+1}); the coupling block
  // H_{i+1,i} = sum_q rho_q n_q a_q^T and the carry sum_q rho_q n_q n_q^T are formed from the coupling rows q
  // (a_q: own-stage part, n_q: DX_{i+1} part of row q)
  int32_t general_coupling, ncoup_max;
  // sliced-ELL copies of A^ for the ADMM products (see plm_host.cpp): slice bases / source positions (int32 pool), indices (int16 pool)
  int32_t f_rell_base, f_rell_src, f_rell_ind, n_rslices, rell_total;
  int32_t f_cell_base, f_cell_src, f_cell_ind, n_cslices, cell_total;
  int32_t f_sched, n_sched;            // int32 pool: panel schedule of one ADMM iteration, 8 ints per step
  int32_t panel_doubles;               // capacity of one shared-memory panel buffer (doubles)
  // the same for the latency kernel (whole stages as panels)
  int32_t f_sched_lat, n_sched_lat, panel_doubles_lat;
  int32_t gdense_ld;                   // dense integrator rows (neither sparse nor general coupling): leading dimension of a stage's
                                       // zero-filled coupling block [ndx][gdense_ld]; 0 otherwise
  int32_t g_doubles;                   // doubles of one stage's compact coupling block (5 per integrator row: 4 values and their
                                       // own-stage columns packed into the fifth word)
  int32_t fac_off[PLM_MAXNODES + 2];   // offset (doubles) of stage i's inverse block (cyclic diagonals, see above)
  int32_t bk_off[PLM_MAXNODES + 1];    // offset (doubles) of stage i's back-substitution block B_i = S_i^-1 G_i^T, i < N:
                                       // column major [ndx][sp], sp = s rounded up to even
  int32_t fac_total;                   // everything the ADMM iterations stream: the S_i^-1 and the B_i
  int32_t smax;                        // largest stage size
  // OSQP settings
  int32_t max_iter, check_termination, scaling;
  double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
};

struct QpWork {
  int allocated = 0;
  QpLayout* d_ql = nullptr;
  int16_t* d_idx = nullptr;
  int32_t* d_idx32 = nullptr;
  // per instance, [max_batch][...]
  double* Ahat = nullptr;    // [nnz]   E A D, CSR order (value order of J)
  double* AhatT = nullptr;   // [nnz]   the same values in CSC order
  double* AhatR = nullptr;   // [rell_total] sliced-ELL by rows (z~ = A x~)
  double* AhatC = nullptr;   // [cell_total] sliced-ELL by columns (A^T w)
  double* D = nullptr;       // [n]
  double* E = nullptr;       // [m]
  double* Eprev = nullptr;   // [m]     row scaling in force when the bounds are classified (osqp update order)
  double* cscale = nullptr;  // [1]
  double* Ph = nullptr;      // [n]     c D P D
  double* qh = nullptr;      // [n]     c D q
  double* lh = nullptr;      // [m]     E l
  double* uh = nullptr;      // [m]     E u
  double* rho = nullptr;     // [m]     rho_vec
  double* Linv = nullptr;    // [fac_total] inverse stage blocks S_i^-1 (cyclic diagonals), then the back-substitution blocks B_i
  double* Gc = nullptr;      // [nodes][g_doubles] compact coupling blocks diag(rho n) A_int (sparse couplings only)
  double* x = nullptr;       // [n]     persistent scaled ADMM iterates
  double* z = nullptr;       // [m]
  double* y = nullptr;       // [m]
};

}  // namespace plm

// Batched OSQP-style ADMM QP solver: one CTA per MPC instance.
//
// Replaces osqp.OSQP.setup / update / solve as called at optimization/ocp.py:312-313, :395, :401
// (third-party osqp 0.6.x semantics restated in SURVEY.md appendix A.8):
//   qp_scale_kernel   Ruiz equilibration (scale_data), cost scaling, rho_vec classification
//   qp_factor_kernel  reduced KKT  H = P + sigma I + A^T diag(rho) A  is block tridiagonal over the shooting
//                     stages (only the integrator rows couple stage i with DX_{i+1}); block Cholesky, the inverse
//                     S_i^-1 of every Schur-complemented stage block is stored packed together with the
//                     back-substitution block B_i = S_i^-1 G_i^T, so the ADMM sweeps are plain mat-vecs
//   qp_admm_kernel    x~ = H^-1 (sigma x - q + A^T(rho z - y)), z~ = A x~, relaxation, projection on [l,u],
//                     dual update, termination tests every check_termination iterations (unscaled residuals,
//                     primal / dual infeasibility certificates), persistent warm-started iterates.
#include <math.h>
#include <stdlib.h>

#include "plm_handle.cuh"

using namespace plm;

#define OSQP_INFTY 1e30
#define MIN_SCALING 1e-4
#define MAX_SCALING 1e4
#define RHO_MIN 1e-6
#define RHO_TOL 1e-4
#define RHO_EQ_OVER_RHO_INEQ 1e3
#define QP_THREADS 256

__device__ __forceinline__ double limit_scaling(double v) {
  v = v < MIN_SCALING ? 1.0 : v;
  return v > MAX_SCALING ? MAX_SCALING : v;
}

// deterministic block reductions (fixed tree)
__device__ __forceinline__ double block_reduce(double v, double* red, bool is_max) {
  for (int o = 16; o > 0; o >>= 1) {
    double t = __shfl_down_sync(0xffffffffu, v, o);
    v = is_max ? fmax(v, t) : v + t;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = is_max ? fmax(r, red[w]) : r + red[w];
  __syncthreads();
  return r;
}

struct StageView {
  const int16_t *rptr, *ccol, *cptr, *cpos, *crow;
  int nrows, ncols, s;
};

__device__ __forceinline__ StageView stage_view(const PlmLayout& L, const QpLayout& Q, const int16_t* idx, int node) {
  const int t = L.node_type[node];
  const QpTypeIdx& I = Q.type[t];
  StageView v;
  v.rptr = idx + I.rptr; v.ccol = idx + I.ccol; v.cptr = idx + I.cptr; v.cpos = idx + I.cpos; v.crow = idx + I.crow;
  v.nrows = L.types[t].nrows; v.ncols = I.ncols; v.s = I.s;
  return v;
}

// ------------------------------------------------------------------------------------------------------------
// Scaling.  mode 0: data update (A = J values, q, l, u given).  mode 1: setup-time scaling with the dummy data of
// optimization/ocp.py:305-310 (A = ones on the pattern, q = 1, l = -1, u = 1); only E is kept (as Eprev).
// ------------------------------------------------------------------------------------------------------------
#ifndef PLM_SCALE_THREADS
#define PLM_SCALE_THREADS 256
#define PLM_SCALE_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(PLM_SCALE_THREADS, PLM_SCALE_MIN_CTAS)
qp_scale_kernel(DeviceTables tab, const QpLayout* __restrict__ Qp, const int16_t* __restrict__ idx, const int32_t* __restrict__ idx32, int mode, int first,
                const double* __restrict__ hess, const double* __restrict__ qin, const double* __restrict__ Jv,
                const double* __restrict__ lin, const double* __restrict__ uin, QpWork W) {
  extern __shared__ double sm[];
  const PlmLayout& L = *tab.layout;
  const QpLayout& Q = *Qp;
  const int b = first + blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int n = L.n, m = L.m, nnz = L.nnz;
  double* D = sm;            // [n]
  double* E = D + n;         // [m]
  double* red = E + m;       // [32]
  const int32_t* rptr = idx32 + Q.f_rptr;
  const int32_t* tptr = idx32 + Q.f_tptr;
  const int32_t* tsrc = idx32 + Q.f_tsrc;
  const int16_t* rcol = idx + Q.f_rcol;
  const int16_t* trow = idx + Q.f_trow;
  const int16_t* rperm = idx + Q.f_rperm;      // rows / columns by decreasing length: the lanes of a warp get equal work
  const int16_t* cperm = idx + Q.f_cperm;
  const double* P = hess + (size_t)b * n;
  const double* Ag = Jv ? Jv + (size_t)b * nnz : nullptr;     // mode 1: A = ones on the pattern
  const double* q = qin ? qin + (size_t)b * n : nullptr;
  double* Dg = W.D + (size_t)b * n;      // also the exchange buffers of the Jacobi-style update
  double* Eg = W.E + (size_t)b * m;
  // The passes walk the two sliced-ELL copies of A that the ADMM products use (rows / columns sorted by length, 32 per
  // slice, warp per slice, lane per item: coalesced value and index streams).  The copies are first filled with the
  // unscaled values (one gather pass each), read by the ten passes, and scaled in place at the end; the CSR / CSC
  // copies for the factorisation and the termination tests are written once from the final scaling.
  // (Before: thread per row / column over the CSR values and a CSC gather, 16.6 ms at 8192 instances; alternatives
  // measured then: values resident in shared memory with one 1024-thread CTA per SM 17.7 ms, warp per row 16.4 ms.
  // Measured on this form, both slower: 512- / 1024-thread CTAs that keep the copies of the instances in flight in L2
  // (+2 / +3 ms), one walk of the row copy per pass with the column maxima gathered by shared-memory atomicMax on the
  // bit patterns (64-bit shared atomicMax is a compare-and-swap loop: +4.5 ms).)
  const double* A = Ag;
  const int32_t* rbase = idx32 + Q.f_rell_base;
  const int32_t* cbase = idx32 + Q.f_cell_base;
  const int16_t* rind = idx + Q.f_rell_ind;
  const int16_t* cind = idx + Q.f_cell_ind;
  double* AR = W.AhatR + (size_t)b * Q.rell_total;
  double* AC = W.AhatC + (size_t)b * Q.cell_total;
  {
    const int32_t* rsrc = idx32 + Q.f_rell_src;
    const int32_t* csrc = idx32 + Q.f_cell_src;
    for (int e = tid; e < Q.rell_total; e += nth) { const int sp = rsrc[e]; AR[e] = sp >= 0 ? (A ? A[sp] : 1.0) : 0.0; }
    for (int e = tid; e < Q.cell_total; e += nth) { const int sp = csrc[e]; AC[e] = sp >= 0 ? (A ? A[sp] : 1.0) : 0.0; }
  }
  for (int j = tid; j < n; j += nth) D[j] = 1.0;
  for (int r = tid; r < m; r += nth) E[r] = 1.0;
  double c = 1.0;
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nw = nth >> 5;
  // max_j |vals| * v[ind] of every item (padding slots hold 0 and index 0)
  auto ell_max = [&](const int32_t* __restrict__ base, const int16_t* __restrict__ ind, int sl, const double* vals, const double* v) {      // (vals: written by this kernel, coherent loads)
    const int b0 = base[sl] + lane, b1 = base[sl + 1];
    double acc = 0.0;
    int p = b0;
    for (; p + 32 * 7 < b1; p += 32 * 8) {
      double a[8];
      int c[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { a[j] = fabs(vals[p + 32 * j]); c[j] = ind[p + 32 * j]; }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmax(acc, a[j] * v[c[j]]);
    }
    if (p < b1) {      // the remaining rows as one predicated group (all loads in flight at once)
      double a[7];
      int c[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const bool ok = p + 32 * j < b1;
        a[j] = ok ? fabs(vals[p + 32 * j]) : 0.0;
        c[j] = ok ? (int)ind[p + 32 * j] : 0;
      }
#pragma unroll
      for (int j = 0; j < 7; ++j) acc = fmax(acc, a[j] * v[c[j]]);
    }
    return acc;
  };
  for (int pass = 0; pass < Q.scaling; ++pass) {
    // row norms of the current scaled A: E_r * max_k |A_rk| D_col ; column norms: max(|P^_jj|, D_j * max_r E_r |A_rj|)
    for (int sl = warp; sl < Q.n_rslices; sl += nw) {
      const double v = ell_max(rbase, rind, sl, AR, D);
      const int item = 32 * sl + lane;
      if (item < m) { const int r = rperm[item]; Eg[r] = E[r] / sqrt(limit_scaling(E[r] * v)); }
    }
    for (int sl = warp; sl < Q.n_cslices; sl += nw) {
      double v = ell_max(cbase, cind, sl, AC, E);
      const int item = 32 * sl + lane;
      if (item < n) {
        const int j = cperm[item];
        v = fmax(D[j] * v, c * D[j] * D[j] * fabs(P[j]));
        Dg[j] = D[j] / sqrt(limit_scaling(v));
      }
    }
    __syncthreads();
    for (int j = tid; j < n; j += nth) D[j] = Dg[j];
    for (int r = tid; r < m; r += nth) E[r] = Eg[r];
    __syncthreads();
    // cost scaling: c_temp = 1 / limit(max(mean_j |P^_jj|, limit(||q^||_inf)))
    double sp_ = 0.0, mq = 0.0;
    for (int j = tid; j < n; j += nth) {
      sp_ += c * D[j] * D[j] * fabs(P[j]);
      mq = fmax(mq, fabs(c * D[j] * (q ? q[j] : 1.0)));
    }
    const double psum = block_reduce(sp_, red, false);
    const double qmax = block_reduce(mq, red, true);
    const double ct = 1.0 / limit_scaling(fmax(psum / (double)n, limit_scaling(qmax)));
    c *= ct;
  }
  if (mode == 1) {
    for (int r = tid; r < m; r += nth) W.Eprev[(size_t)b * m + r] = E[r];
    return;
  }
  // scaled data: CSR-ordered and CSC-ordered copies of E A D
  double* Ah = W.Ahat + (size_t)b * nnz;
  double* AT = W.AhatT + (size_t)b * nnz;
  for (int i = tid; i < m; i += nth) {
    const int r = rperm[i];
    for (int e = rptr[r]; e < rptr[r + 1]; ++e) Ah[e] = E[r] * A[e] * D[rcol[e]];
  }
  for (int i = tid; i < n; i += nth) {
    const int j = cperm[i];
    for (int e = tptr[j]; e < tptr[j + 1]; ++e) AT[e] = E[trow[e]] * A[tsrc[e]] * D[j];
  }
  for (int j = tid; j < n; j += nth) {
    W.Ph[(size_t)b * n + j] = c * D[j] * D[j] * P[j];
    W.qh[(size_t)b * n + j] = c * D[j] * q[j];
  }
  if (tid == 0) W.cscale[b] = c;
  // the sliced-ELL copies are scaled in place: (E_r a) D_c, the same product as the CSR copy
  for (int sl = warp; sl < Q.n_rslices; sl += nw) {
    const int item = 32 * sl + lane;
    const double er = item < m ? E[rperm[item]] : 0.0;
    for (int p2 = rbase[sl] + lane; p2 < rbase[sl + 1]; p2 += 32) AR[p2] = er * AR[p2] * D[rind[p2]];
  }
  for (int sl = warp; sl < Q.n_cslices; sl += nw) {
    const int item = 32 * sl + lane;
    const double dc = item < n ? D[cperm[item]] : 0.0;
    for (int p2 = cbase[sl] + lane; p2 < cbase[sl + 1]; p2 += 32) AC[p2] = E[cind[p2]] * AC[p2] * dc;
  }
  const double* l = lin + (size_t)b * m;
  const double* u = uin + (size_t)b * m;
  for (int r = tid; r < m; r += nth) {
    const double lc = fmax(l[r], -OSQP_INFTY), uc = fmin(u[r], OSQP_INFTY);
    // osqp_update_bounds runs before osqp_update_A: rows are classified with the previous row scaling
    const double ep = W.Eprev[(size_t)b * m + r];
    const double lp = ep * lc, up = ep * uc;
    double rho;
    if (lp < -OSQP_INFTY * MIN_SCALING && up > OSQP_INFTY * MIN_SCALING) rho = RHO_MIN;
    else if (up - lp < RHO_TOL) rho = RHO_EQ_OVER_RHO_INEQ * Q.rho;
    else rho = Q.rho;
    W.rho[(size_t)b * m + r] = rho;
    W.lh[(size_t)b * m + r] = E[r] * lc;
    W.uh[(size_t)b * m + r] = E[r] * uc;
    W.Eprev[(size_t)b * m + r] = E[r];
  }
}

// ------------------------------------------------------------------------------------------------------------
// Factorisation.  Packed lower-triangular storage: element (i, j <= i) at i(i+1)/2 + j.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int tri(int i, int j) { return ((i * (i + 1)) >> 1) + j; }      // (i (i+1) is even and >= 0: shift, not signed division)

// D = A B + C on the FP64 tensor cores (DMMA m8n8k4).  Lane l holds A[l/4][l%4] (8x4, row major), B[l%4][l/4] (4x8,
// column major) and C / D [l/4][2 (l%4) + {0, 1}] (8x8).
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

#ifndef PLM_FACTOR_MIN_CTAS
#define PLM_FACTOR_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(QP_THREADS, PLM_FACTOR_MIN_CTAS)
qp_factor_kernel(DeviceTables tab, const QpLayout* __restrict__ Qp, const int16_t* __restrict__ idx, QpWork W, int* __restrict__ fail) {
  extern __shared__ double sm[];
  const PlmLayout& L = *tab.layout;
  const QpLayout& Q = *Qp;
  const int b = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int n = L.n, m = L.m, ndx = L.ndx, N = L.nodes, smax = Q.smax;
  const int tsz = smax * (smax + 1) / 2;
  double* H = sm;                      // [tsz]  stage block -> (in place) inverse X = L^-1 of its Cholesky factor, packed lower
  const int wld = ((ndx + 15) & ~15) + 8;      // leading dimension of W: = 8 mod 16, so that the four rows of a DMMA fragment fall in distinct banks
  double* Wm = H + tsz;                // [smax][wld]  W = X G^T (dense coupling only)
  // K: [ndx (ndx+1)/2] Schur term for the next stage, packed lower.  It is written at the end of a stage and consumed by
  // the assembly of the next one; in between (factorisation) the same shared memory is the scratch of the blocked steps:
  // Tm [8][smax] and dsc [72].
  double* K = Wm + (Q.sparse_coupling ? 0 : smax * wld);
  const int ksz = max(ndx * (ndx + 1) / 2, 8 * smax + 72);
  double* gsc = K + ksz;               // [ndx]  rho_r * (next entry)^2 of the integrator rows
  double* rs = gsc + ndx;              // [max_rows] rho of the node rows
  // The scaled A values of the node block are staged behind the stage block when the stage is smaller than the largest
  // one (17 of the 20 B2G stages), else read from global memory (L1 / L2): what this saves lets four CTAs share an SM.
  const double* Ah = W.Ahat + (size_t)b * L.nnz;
  const double* Ph = W.Ph + (size_t)b * n;
  const double* rho = W.rho + (size_t)b * m;
  double* Lout = W.Linv + (size_t)b * Q.fac_total;
  if (tid == 0) fail[b] = 0;

  for (int i = 0; i <= N; ++i) {
    const bool last = (i == N);
    StageView sv;
    if (!last) sv = stage_view(L, Q, idx, i);
    const int s = last ? ndx : sv.s;
    const int xo = L.x_off[i];
    const double* As = nullptr;
    if (!last) {   // stage the node's values in shared memory: every entry is used many times below
      const double* An = Ah + L.nnz_off[i];
      const double* rh = rho + L.row_off[i];
      const int nnz_i = L.types[L.node_type[i]].nnz;
      const int hsz = (s * (s + 1) / 2 + 1) & ~1;
      if (tsz - hsz >= nnz_i) {
        double* Asm = H + hsz;
        for (int e = tid; e < nnz_i; e += nth) Asm[e] = An[e];
        As = Asm;
      } else As = An;
      for (int r = tid; r < sv.nrows; r += nth) rs[r] = rh[r];
    }
    __syncthreads();
    // ---- H_ii (lower triangle): thread j owns row j
    for (int j = tid; j < s; j += nth) {
      double* Hj = H + tri(j, 0);
      for (int k = 0; k <= j; ++k) Hj[k] = 0.0;
      Hj[j] = Ph[xo + j] + Q.sigma;
      if (j < ndx) {
        if (i == 0) Hj[j] += rho[j] * Ah[j] * Ah[j];           // DX_0 == 0 rows
        else {
          for (int k = 0; k <= j; ++k) Hj[k] -= K[tri(j, k)];    // Schur complement of stage i-1
          Hj[j] += gsc[j];                                       // rho_r n_r^2 of the integrator row feeding DX_i[j]
        }
      }
      if (!last) {
        for (int e = sv.cptr[j]; e < sv.cptr[j + 1]; ++e) {
          const int r = sv.crow[e];
          const double w = rs[r] * As[sv.cpos[e]];
          // the columns of one row are distinct (ascending): four read-modify-writes of H in flight at a time
          int e2 = sv.rptr[r];
          const int e2e = sv.rptr[r + 1];
          for (; e2 + 3 < e2e && sv.ccol[e2 + 3] <= j; e2 += 4) {
            const int k0 = sv.ccol[e2], k1 = sv.ccol[e2 + 1], k2 = sv.ccol[e2 + 2], k3 = sv.ccol[e2 + 3];
            const double h0 = Hj[k0], h1 = Hj[k1], h2 = Hj[k2], h3 = Hj[k3];
            Hj[k0] = h0 + w * As[e2];
            Hj[k1] = h1 + w * As[e2 + 1];
            Hj[k2] = h2 + w * As[e2 + 2];
            Hj[k3] = h3 + w * As[e2 + 3];
          }
          for (; e2 < e2e; ++e2) {
            const int k = sv.ccol[e2];
            if (k > j) break;
            Hj[k] += w * As[e2];
          }
        }
      }
    }
    __syncthreads();
    // ---- Cholesky H = L L^T and X = L^-1, in place (packed lower), blocked by eight columns with the inner products on
    // the FP64 tensor cores (DMMA m8n8k4) and the accumulators in registers over the whole k-loop:
    //   P1 (left-looking, block column J): H[j0:, J] -= L[j0:, 0:j0] L[J, 0:j0]^T (8x8 tiles dealt to the warps); warp 0
    //      factors the 8x8 diagonal block and stores ITS INVERSE in its place (the block of L itself is not needed
    //      again); one thread per row below solves L[r, J] = H[r, J] L_JJ^-T with that inverse.
    //   P2 (row block I, top down): X[I, 0:i0] = -L_II^-1 (L[I, 0:i0] X[0:i0, 0:i0]), X[I, I] = L_II^-1 (already there):
    //      tiles of the product into a scratch row block, then one thread per column applies -L_II^-1.
    // Rows / columns beyond s (ragged last block) are masked; Tm (8 x smax scratch) and dsc alias K.
    {
      const int warp = tid >> 5, lane = tid & 31;
      constexpr int nw = QP_THREADS >> 5;
      const int fr = lane >> 2, fk = lane & 3;
      const int nb8 = (s + 7) >> 3;
      double* Tm = K;                       // [8][smax] (K is dead until the end of the stage)
      double* dsc = K + 8 * smax;           // [64 + 8] scratch of the diagonal-block step
      for (int Jb = 0; Jb < nb8; ++Jb) {
        const int j0 = 8 * Jb;
        // (1) update of block column J: warp 0 takes the diagonal tile and goes straight on to factor it (2) while the
        // other warps share the tiles below (the factorisation of the 8x8 block is a serial chain of one warp)
        const int ntile = nb8 - Jb;
        for (int t = (warp == 0) ? 0 : warp; t < (warp == 0 ? 1 : ntile); t += nw - 1) {
          const int r0 = j0 + 8 * t;
          const int ar = r0 + fr, br = j0 + fr;
          const bool aok = ar < s, bok = br < s;
          const double* Arow = H + tri(aok ? ar : 0, 0);
          const double* Brow = H + tri(bok ? br : 0, 0);
          const int cc = j0 + 2 * fk;
          double* Cp = H + tri(aok ? ar : 0, 0) + cc;
          const bool v0 = aok && cc <= ar && cc < s, v1 = aok && cc + 1 <= ar && cc + 1 < s;
          double c0v = v0 ? Cp[0] : 0.0, c1v = v1 ? Cp[1] : 0.0;
          for (int kk = 0; kk < j0; kk += 4) {
            const double av = aok ? -Arow[kk + fk] : 0.0;
            const double bv = bok ? Brow[kk + fk] : 0.0;
            dmma884(c0v, c1v, av, bv, c0v, c1v);
          }
          if (v0) Cp[0] = c0v;
          if (v1) Cp[1] = c1v;
        }
        // (2) diagonal block: lane i = row i of the 8x8 block (identity beyond s); Cholesky by columns with shuffles, then
        // lane c solves L x = e_c for column c of the inverse.  (The other warps only read columns < j0 of these rows.)
        if (warp == 0) {
          __syncwarp();
          const int mrows = min(8, s - j0);
          const int ri = lane & 7;
          double a8[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) a8[c] = (ri < mrows && c <= ri) ? H[tri(j0 + ri, 0) + j0 + c] : (c == ri ? 1.0 : 0.0);
          double myinv = 1.0;
          bool bad = false;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const double dcc = __shfl_sync(0xffffffffu, a8[c], c);
            if (!(dcc > 0.0)) bad = true;
            const double inv = rsqrt(dcc > 0.0 ? dcc : 1.0);
            const double lic = a8[c] * inv;           // lane c: sqrt(d_cc); lanes i > c: L[i][c]
            a8[c] = lic;
            if (ri == c) myinv = inv;
#pragma unroll
            for (int j = c + 1; j < 8; ++j) {
              const double ljc = __shfl_sync(0xffffffffu, lic, j);
              if (ri >= j) a8[j] -= lic * ljc;
            }
          }
          if (bad && lane == 0) atomicExch(&fail[b], i + 1);
          if (lane < 8) {
#pragma unroll
            for (int c = 0; c < 8; ++c) dsc[8 * lane + c] = a8[c];
            dsc[64 + lane] = myinv;
          }
          __syncwarp();
          if (lane < 8) {
            // column `lane` of L_JJ^-1: x[r] for r >= lane
            double x8[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              double acc = (r == lane) ? 1.0 : 0.0;
#pragma unroll
              for (int d = 0; d < 8; ++d)
                if (d < r && d >= lane) acc -= dsc[8 * r + d] * x8[d];
              x8[r] = (r >= lane) ? acc * dsc[64 + r] : 0.0;
            }
#pragma unroll
            for (int r = 0; r < 8; ++r)
              if (r >= lane && r < mrows && lane < mrows) H[tri(j0 + r, 0) + j0 + lane] = x8[r];
          }
        }
        __syncthreads();
        // (3) panel: L[R, J] = H[R, J] L_JJ^-T for the 8-row tiles R below the block, two DMMAs per tile:
        // D[r][c] = sum_d H[r][j0 + d] (L_JJ^-1)[c][d]  (rows below exist only for full blocks)
        if (ntile > 1) {
          for (int t = 1 + warp; t < ntile; t += nw) {
            const int ar = j0 + 8 * t + fr;
            const bool aok = ar < s;
            double* Hr = H + tri(aok ? ar : 0, 0) + j0;
            const double* Li = H + tri(j0 + fr, 0) + j0;          // row fr of the inverse block (column index of D)
            const double a0 = aok ? Hr[fk] : 0.0, a1 = aok ? Hr[4 + fk] : 0.0;
            const double b0 = (fk <= fr) ? Li[fk] : 0.0, b1 = (4 + fk <= fr) ? Li[4 + fk] : 0.0;
            double d0 = 0.0, d1 = 0.0;
            dmma884(d0, d1, a0, b0, d0, d1);
            dmma884(d0, d1, a1, b1, d0, d1);
            __syncwarp();
            if (aok) { Hr[2 * fk] = d0; Hr[2 * fk + 1] = d1; }
          }
          __syncthreads();
        }
      }
      // ---- P2: X = L^-1 row block by row block
      for (int Ib = 1; Ib < nb8; ++Ib) {
        const int i0 = 8 * Ib;
        for (int t = warp; t < Ib; t += nw) {
          const int c0 = 8 * t;
          const int ar = i0 + fr;
          const bool aok = ar < s;
          const double* Arow = H + tri(aok ? ar : 0, 0);
          const int bc = c0 + fr;                 // column of X in the B fragment
          double c0v = 0.0, c1v = 0.0;
          for (int kk = c0; kk < i0; kk += 4) {
            const int xr = kk + fk;               // row of X in the B fragment (xr < i0 <= s)
            const double av = aok ? Arow[kk + fk] : 0.0;
            const double bv = (bc <= xr) ? H[tri(xr, 0) + bc] : 0.0;
            dmma884(c0v, c1v, av, bv, c0v, c1v);
          }
          Tm[fr * smax + c0 + 2 * fk] = c0v;
          Tm[fr * smax + c0 + 2 * fk + 1] = c1v;
        }
        __syncthreads();
        {
          // X[I, 0:i0] = -L_II^-1 Tm, 8-column tiles, two DMMAs each: D[r][c] = sum_d (L_II^-1)[r][d] Tm[d][c]
          const int ar = i0 + fr;
          const bool aok = ar < s;
          const double* Li = H + tri(aok ? ar : 0, 0) + i0;
          const double a0 = (aok && fk <= fr) ? Li[fk] : 0.0, a1 = (aok && 4 + fk <= fr) ? Li[4 + fk] : 0.0;
          for (int t = warp; t < Ib; t += nw) {
            const int c0 = 8 * t;
            const double b0 = Tm[fk * smax + c0 + fr], b1 = Tm[(4 + fk) * smax + c0 + fr];
            double d0 = 0.0, d1 = 0.0;
            dmma884(d0, d1, a0, b0, d0, d1);
            dmma884(d0, d1, a1, b1, d0, d1);
            if (aok) {
              double* Xo = H + tri(ar, 0) + c0 + 2 * fk;
              Xo[0] = -d0;
              Xo[1] = -d1;
            }
          }
        }
        __syncthreads();
      }
    }
    // ---- S_i^-1 = X^T X, stored by cyclic diagonals (plm_qp_types.h): what the ADMM sweeps multiply with (one
    // symmetric product per stage visit).
    // 8x8 output tiles on the FP64 tensor cores: S[r][c] = sum_{t >= r} X[t][r] X[t][c], four rows t of X per DMMA.
    {
      const int warp = tid >> 5, lane = tid & 31;
      constexpr int nw = QP_THREADS >> 5;
      const int nt8 = (s + 7) >> 3;
      const int fr = lane >> 2, fk = lane & 3;
      double* So = Lout + Q.fac_off[i];
      int turn = warp;                           // lower tiles dealt round robin to the warps (see the rank-4 update)
      for (int tr = 0; tr < nt8; ++tr) {
       const int nc = tr + 1;
       int tc = turn;
       turn = (turn >= nc) ? turn - nc : (nw - 1) - ((nc - 1 - turn) % nw);
       for (; tc < nc; tc += nw) {
        const int r0 = 8 * tr, c0 = 8 * tc;
        const int ra = r0 + fr, cb = c0 + fr;    // this lane's column of X in the A / B fragments
        double d0 = 0.0, d1 = 0.0;
        for (int t = r0; t < s; t += 4) {
          const int tt = t + fk;
          const double* Xt = H + tri(tt < s ? tt : 0, 0);
          const double a = (tt < s && ra <= tt) ? Xt[ra] : 0.0;
          const double bv = (tt < s && cb <= tt) ? Xt[cb] : 0.0;
          dmma884(d0, d1, a, bv, d0, d1);
        }
        const int cc = c0 + 2 * fk;
        if (ra < s) {
          if (cc <= ra) So[plm_sinv_index(s, ra, cc)] = d0;
          if (cc + 1 <= ra) So[plm_sinv_index(s, ra, cc + 1)] = d1;
        }
       }
      }
      // even s: the second half of the last cyclic diagonal repeats the first; it is stored as zeros
      if (!(s & 1) && tid < (s >> 1)) So[(s >> 1) * s + (s >> 1) + tid] = 0.0;
    }
    if (last) break;
    // ---- coupling to stage i+1: G = diag(g) * A_int,loc ; W = X G^T ; K = W^T W
    for (int c2 = tid; c2 < ndx && !Q.general_coupling; c2 += nth) {
      const int elast = sv.rptr[c2 + 1] - 1;              // next entry of integrator row c2
      const double nn = As[elast];
      gsc[c2] = rs[c2] * nn * nn;
    }
    if (Q.sparse_coupling) {
      // K = G S^-1 G^T straight from the block just written (L2-hot): a handful of terms per entry
      __syncthreads();     // the S^-1 block written above is visible to the whole CTA
      const double* Sg = Lout + Q.fac_off[i];
      for (int o = tid; o < ndx * (ndx + 1) / 2; o += nth) {
        int r = (int)((sqrt(8.0 * o + 1.0) - 1.0) * 0.5);
        while (tri(r + 1, 0) <= o) ++r;
        while (tri(r, 0) > o) --r;
        const int c2 = o - tri(r, 0);
        const int ea0 = sv.rptr[r], ea1 = sv.rptr[r + 1] - 1, eb0 = sv.rptr[c2], eb1 = sv.rptr[c2 + 1] - 1;
        double acc = 0.0;
        for (int ea = ea0; ea < ea1; ++ea) {
          const int ja = sv.ccol[ea];
          for (int eb = eb0; eb < eb1; ++eb) {
            const int jb = sv.ccol[eb];
            acc += As[ea] * As[eb] * Sg[ja >= jb ? plm_sinv_index(s, ja, jb) : plm_sinv_index(s, jb, ja)];
          }
        }
        K[o] = rs[r] * As[ea1] * rs[c2] * As[eb1] * acc;
      }
      // back-substitution block B_i = S_i^-1 G_i^T (column major [ndx][sp]): the backward sweep of the ADMM
      // iterations is x_i = tv_i - B_i x_{i+1}[0:ndx], a plain product in which every stored element is used once
      {
        const int sp = (s + 1) & ~1;
        double* Bo = Lout + Q.bk_off[i];
        for (int o = tid; o < ndx * sp; o += nth) {
          const int c2 = o / sp, kk = o - c2 * sp;
          double acc = 0.0;
          if (kk < s) {
            const int e0 = sv.rptr[c2], e1 = sv.rptr[c2 + 1] - 1;
            for (int e = e0; e < e1; ++e) {
              const int j = sv.ccol[e];
              acc += As[e] * Sg[kk >= j ? plm_sinv_index(s, kk, j) : plm_sinv_index(s, j, kk)];
            }
            acc *= rs[c2] * As[e1];
          }
          Bo[o] = acc;
        }
      }
      __syncthreads();
      continue;
    }
    if (Q.general_coupling) {
      // Any row of the node may touch DX_{i+1} (whole_body_rnea without acceleration inputs: RNEA rows depend on dv_{i+1}).
      // With a_q / n_q the own-stage / DX_{i+1} part of coupling row q:  G = sum_q rho_q n_q a_q^T,  W = X G^T, i.e.
      // W[t][c] = sum_q n_q[c] Y[t][q],  Y[t][q] = rho_q (X a_q)[t];  the carry sum_q rho_q n_q n_q^T replaces the diagonal gsc.
      const QpTypeIdx& I = Q.type[L.node_type[i]];
      const int16_t* crows = idx + I.gc_rows;
      const int nc = I.ncoup, ncm = Q.ncoup_max;
      double* Y = rs + L.max_rows;          // [smax][ncm]
      double* Nn = Y + smax * ncm;          // [ncm][ndx]
      for (int o = tid; o < nc * ndx; o += nth) Nn[o] = 0.0;
      for (int c2 = tid; c2 < ndx; c2 += nth) gsc[c2] = 0.0;
      __syncthreads();
      for (int q = tid; q < nc; q += nth) {
        const int r = crows[q];
        for (int e = sv.rptr[r]; e < sv.rptr[r + 1]; ++e)
          if (sv.ccol[e] >= s) Nn[q * ndx + sv.ccol[e] - s] = As[e];
      }
      const int s32 = (s + 31) & ~31;
      for (int o = tid; o < s32 * nc; o += nth) {
        const int q = o / s32, t = o - q * s32;
        if (t >= s) continue;
        const int r = crows[q];
        const double* Xt = H + tri(t, 0);
        double acc = 0.0;
        for (int e = sv.rptr[r]; e < sv.rptr[r + 1]; ++e) {
          const int k = sv.ccol[e];
          if (k > t) break;                 // (columns ascending: the rest is above the diagonal of X or in DX_{i+1})
          acc += As[e] * Xt[k];
        }
        Y[t * ncm + q] = rs[r] * acc;
      }
      __syncthreads();
      for (int o = tid; o < s * ndx; o += nth) {
        const int t = o / ndx, c2 = o - t * ndx;
        double acc = 0.0;
        for (int q = 0; q < nc; ++q) acc += Nn[q * ndx + c2] * Y[t * ncm + q];
        Wm[t * wld + c2] = acc;
      }
    } else
    {   // lanes of a warp share the integrator row c2 (uniform entry loop: dense and sparse rows do not mix) and take
        // consecutive rows t of X
      const int s32 = (s + 31) & ~31;
      for (int o = tid; o < s32 * ndx; o += nth) {
        const int c2 = o / s32, t = o - c2 * s32;
        if (t >= s) continue;
        const int e0 = sv.rptr[c2], e1 = sv.rptr[c2 + 1] - 1;
        const double g = rs[c2] * As[e1];
        double acc = 0.0;
        const double* Xt = H + tri(t, 0);
        for (int e = e0; e < e1; ++e) {
          const int k = sv.ccol[e];
          if (k <= t) acc += As[e] * Xt[k];
        }
        Wm[t * wld + c2] = g * acc;
      }
    }
    __syncthreads();
    {   // B_i = S_i^-1 G_i^T = X^T W and K = W^T W (- carry) as 8x8 tiles on the FP64 tensor cores
      const int warp = tid >> 5, lane = tid & 31, fr = lane >> 2, fk = lane & 3;
      constexpr int nw = QP_THREADS >> 5;
      const int sp = (s + 1) & ~1;
      double* Bo = Lout + Q.bk_off[i];
      const int nkt = (s + 7) >> 3, nct = (ndx + 7) >> 3;
      // D[kk][c2] = sum_{t >= kk} X[t][kk] W[t][c2]
      for (int tile = warp; tile < nkt * nct; tile += nw) {
        const int kt = tile / nct, ct = tile - kt * nct;
        const int kk = 8 * kt + fr;           // row of the A fragment
        const int cb = 8 * ct + fr;           // column of the B fragment
        double d0 = 0.0, d1 = 0.0;
        for (int t0 = 8 * kt; t0 < s; t0 += 4) {
          const int t = t0 + fk;
          const double a = (t < s && t >= kk) ? H[tri(t, kk)] : 0.0;      // (t >= kk and t < s imply kk < s)
          const double bv = (t < s && cb < ndx) ? Wm[t * wld + cb] : 0.0;
          dmma884(d0, d1, a, bv, d0, d1);
        }
        const int c2 = 8 * ct + 2 * fk;
        if (kk < s) {
          if (c2 < ndx) Bo[c2 * sp + kk] = d0;
          if (c2 + 1 < ndx) Bo[(c2 + 1) * sp + kk] = d1;
        }
      }
      if (sp > s)
        for (int c2 = tid; c2 < ndx; c2 += nth) Bo[c2 * sp + s] = 0.0;      // padding row of the column-major block
      // K[r][c2] = sum_t W[t][r] W[t][c2] - carry, lower tiles
      const QpTypeIdx& I = Q.type[L.node_type[i]];
      const int16_t* crows = idx + I.gc_rows;
      const double* Nn = rs + L.max_rows + smax * Q.ncoup_max;
      for (int tile = warp; tile < nct * (nct + 1) / 2; tile += nw) {
        int rt = 0;
        while ((rt + 1) * (rt + 2) / 2 <= tile) ++rt;
        const int ct = tile - rt * (rt + 1) / 2;
        const int ra = 8 * rt + fr, cb = 8 * ct + fr;
        double d0 = 0.0, d1 = 0.0;
        for (int t0 = 0; t0 < s; t0 += 4) {
          const int t = t0 + fk;
          const double a = (t < s && ra < ndx) ? Wm[t * wld + ra] : 0.0;
          const double bv = (t < s && cb < ndx) ? Wm[t * wld + cb] : 0.0;
          dmma884(d0, d1, a, bv, d0, d1);
        }
#pragma unroll
        for (int q2 = 0; q2 < 2; ++q2) {
          const int c2 = 8 * ct + 2 * fk + q2;
          if (ra < ndx && c2 <= ra) {
            double carry = 0.0;               // general coupling: H_{i+1,i+1} += sum_q rho_q n_q n_q^T (K is subtracted from H)
            if (Q.general_coupling)
              for (int q = 0; q < I.ncoup; ++q) carry += rs[crows[q]] * Nn[q * ndx + ra] * Nn[q * ndx + c2];
            K[tri(ra, c2)] = (q2 ? d1 : d0) - carry;
          }
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------
// ADMM iterations.
// ------------------------------------------------------------------------------------------------------------
// Optional in-kernel phase timers (build with -DPLM_ADMM_PROFILE): cycles of thread 0 of CTA 0 per phase class.
#ifdef PLM_ADMM_PROFILE
__device__ long long g_admm_prof[16];
#define PROF_T0() long long _pt = clock64()
#define PROF_ADD(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) { long long _n = clock64(); g_admm_prof[k] += _n - _pt; _pt = _n; } } while (0)
#else
#define PROF_T0()
#define PROF_ADD(k)
#endif

// Flat index tables of the whole pattern (shared by all instances).
struct FlatIdx {
  const int32_t *rptr, *tptr;
  const int16_t *rcol, *trow, *rperm, *cperm;
};

// Sliced-ELL product: warp per slice of 32 items, lane per item, coalesced value / index streams.
//   out[perm[item]] = sum_j vals[slot] v[ind[slot]] (+ sigma x - q for the column product)
#ifndef ELL_B
#define ELL_B 8     // rows of a slice in flight per lane (4, 12, 16 measured: equal or slower)
#endif
template <bool ADD>
__device__ __forceinline__ void spmv_ell(const int32_t* __restrict__ base, const int16_t* __restrict__ ind, const int16_t* __restrict__ perm, int nitems,
                                         int nsl, const double* __restrict__ vals, const double* v, double* out, double sigma,
                                         const double* x, const double* __restrict__ q, const double* addv = nullptr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  // (the slice bounds and the output index of the next slice are loaded one slice ahead: they head the dependency chain
  // bounds -> values / indices -> gather of every slice)
  int nb0 = 0, nb1 = 0, no = -1;
  if (warp < nsl) {
    nb0 = __ldg(base + warp); nb1 = __ldg(base + warp + 1);
    no = 32 * warp + lane < nitems ? (int)__ldg(perm + 32 * warp + lane) : -1;
  }
  for (int sl = warp; sl < nsl; sl += nw) {
    const int b0 = nb0 + lane, b1 = nb1;
    const int o = no;
    if (sl + nw < nsl) {
      nb0 = __ldg(base + sl + nw); nb1 = __ldg(base + sl + nw + 1);
      no = 32 * (sl + nw) + lane < nitems ? (int)__ldg(perm + 32 * (sl + nw) + lane) : -1;
    }
    double add = 0.0;
    if (ADD && o >= 0) add = addv ? addv[o] : sigma * x[o] - __ldg(q + o);      // (x is rewritten by this kernel: coherent load)
    double acc = 0.0;
    // every slot row of a slice is 32 wide (padded), so the row count is warp uniform: full groups of ELL_B rows run
    // without predicates
    int p = b0;
    for (; p + 32 * (ELL_B - 1) < b1; p += 32 * ELL_B) {
      double a[ELL_B];
      int c[ELL_B];
#pragma unroll
      for (int j = 0; j < ELL_B; ++j) {
        a[j] = __ldg(vals + p + 32 * j);
        c[j] = (int)__ldg(ind + p + 32 * j);
      }
#pragma unroll
      for (int j = 0; j < ELL_B; ++j) acc += a[j] * v[c[j]];
    }
    if (p < b1) {                                 // the remaining rows as one predicated group (all loads in flight at once)
      double a[ELL_B - 1];
      int c[ELL_B - 1];
#pragma unroll
      for (int j = 0; j < ELL_B - 1; ++j) {
        const bool ok = p + 32 * j < b1;
        a[j] = ok ? __ldg(vals + p + 32 * j) : 0.0;
        c[j] = ok ? (int)__ldg(ind + p + 32 * j) : 0;
      }
#pragma unroll
      for (int j = 0; j < ELL_B - 1; ++j) acc += a[j] * v[c[j]];
    }
    if (o >= 0) out[o] = acc + add;
  }
}

// ---- 1-D bulk asynchronous copies global -> shared (TMA, UBLKCP) tracked by an mbarrier
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Experiment switch: L2 prefetch of the panel PLM_ADMM_PREFETCH_STEPS schedule steps ahead of the shared-memory ring (no
// shared memory needed), so that the ring's own bulk copies hit L2.  It does not pay: the kernel is bound by the
// instruction / barrier chain of a CTA, not by the latency of its panel loads.
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
#ifndef PLM_ADMM_PREFETCH_STEPS
#define PLM_ADMM_PREFETCH_STEPS 0   // measured (round 2): 4 steps ahead 592 instances 24.2 -> 25.6 ms, 148 instances (latency kernel) 10.2 -> 10.4 ms; again with
                                    // the cyclic-diagonal layout at 8192 instances: 2 steps ahead +4 %, 4 steps ahead +5 %: no gain, off
#endif
#ifndef PLM_MBAR_SUSPEND_NS
#define PLM_MBAR_SUSPEND_NS 1000      // (250 / 4000 / 20000 ns measured: no difference at 8192 instances, 250 ns slower for one wave)
#endif
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(PLM_MBAR_SUSPEND_NS)      // suspend-time hint: the warp sleeps instead of spinning on issue slots
      : "memory");
}

// The inverse stage blocks S_i^-1 (stored by cyclic diagonals, plm_qp_types.h) are streamed through shared memory in
// panels of consecutive rows j of the stored array (<= PLM_PANEL_DOUBLES) by bulk asynchronous copies, NBUF panels deep,
// following a host-built schedule of one ADMM iteration.  A sweep step is one symmetric product out = S_i^-1 in; each
// stored element M[j][c] = S^-1[c][(c + j) mod s] is used twice (out[c] += M in[c + j], out[c + j] += M in[c]):
//   forward  stage i: tv_i = S_i^-1 (b_i - G_{i-1} tv_{i-1})
//   backward stage i: x_i  = tv_i - B_i x_{i+1}[0:ndx],  B_i = S_i^-1 G_i^T (s x ndx, from the factor kernel; x_N = tv_N):
//                     a plain product, column panels of B_i, every element read from shared memory once
// Two instantiations: throughput (256 threads, 4 CTAs per SM) and, for batches that leave SMs idle, latency (512
// threads, one CTA per SM, no register pressure).
#define ADMM_THREADS 256
#ifndef ADMM_MIN_CTAS
#define ADMM_MIN_CTAS 4     // (three CTAs per SM with 80 registers and two panels per stage: 313 -> 332 ms at 8192 instances)
#endif
#ifndef ADMM_THREADS_LAT
#define ADMM_THREADS_LAT 512     // (one B2G instance, 100 iterations: 256 threads 7.81 ms, 512 threads 7.85 ms, 1024 threads 9.23 ms; 148 instances: 9.88 / 9.44 / 10.64 ms)
#endif
#ifndef NBUF
#define NBUF 2        // ring of panel buffers, throughput kernel (three buffers of 1368 doubles within the same shared memory: +12 %)
#endif
#define NBUF_LAT 3    // latency kernel (whole stages)
#define SYM_K 128    // threads per part: thread (k, part) owns output k of the stage (stage size <= SYM_K), part = tid / SYM_K
#define SYM_PARTS_MAX (ADMM_THREADS_LAT / SYM_K)
static_assert(ADMM_THREADS % SYM_K == 0 && ADMM_THREADS_LAT % SYM_K == 0 && true, "thread layout of sym_panel");

// Thread (k, part) accumulates output k of out = S^-1 in over the rows j = r0 + part, r0 + part + P, ... < r1 of the
// resident panel: the term M[j][k] in[k + j] and the term M[j][k - j] in[k - j] (indices mod s) of every row; row 0 (the
// diagonal) has the first term only.  Lanes read consecutive addresses in all four loads, every thread of the stage has
// the same trip count, nothing is masked.  `vd` is the input vector stored twice in a row ([in, in], 2 s doubles), which
// takes the index wrap off the vector loads; the wrap of the matrix column is one select.  `pan` is the panel buffer
// moved back by the offset of the copy inside the stage block (element (j, c) at pan[j s + c]).  No cross-thread
// reduction inside a stage: the sums live in registers across the panels of the stage.
template <int P>
__device__ __forceinline__ void sym_panel(const double* __restrict__ pan, int s, int r0, int r1, const double* __restrict__ vd, int k, int part,
                                          double& acc0, double& acc1) {
  int j = r0 + part;
  if (j == 0) {
    acc0 += pan[k] * vd[k];
    j = P;
  }
  const double* pf = pan + j * s + k;         // M[j][k]
  const double* vf = vd + k + j;               // in[(k + j) mod s]
  const double* vb = vd + s + k - j;           // in[(k - j) mod s]
  const int ps = P * s;
#pragma unroll 4
  for (; j < r1; j += P) {
    const int ob = (k < j ? s : 0) - j;        // M[j][(k - j) mod s] relative to M[j][k]
    const double af = pf[0], ab = pf[ob];
    acc0 += af * vf[0];
    acc1 += ab * vb[0];
    pf += ps; vf += P; vb -= P;
  }
}

// Backward step: out[k] += sum_j B[k][j] in[j] over the resident columns [j0, j1) of B_i (column major, stride sp).  One
// warp per 32 outputs (k = 32 warp + lane) takes all the columns; the sums live in registers across the panels of the
// stage (four independent chains).  Lanes read consecutive addresses, `in` is a broadcast.
__device__ __forceinline__ void rect_panel(const double* __restrict__ pan, int sp, int j0, int j1, const double* __restrict__ vin, int k,
                                           double& acc0, double& acc1) {
  const double* a = pan + k;
  double s2 = 0.0, s3 = 0.0;
  int j = j0;
  for (; j + 3 < j1; j += 4) {
    const double a0 = a[0], a1 = a[sp], a2 = a[2 * sp], a3 = a[3 * sp];
    acc0 += a0 * vin[j];
    acc1 += a1 * vin[j + 1];
    s2 += a2 * vin[j + 2];
    s3 += a3 * vin[j + 3];
    a += 4 * sp;
  }
  for (; j < j1; ++j) { acc0 += a[0] * vin[j]; a += sp; }
  acc0 += s2;
  acc1 += s3;
}
template <int NT, int MINB, int NB, bool LAT>
__global__ void __launch_bounds__(NT, MINB)
qp_admm_kernel(DeviceTables tab, const QpLayout* __restrict__ Qp, const int16_t* __restrict__ idx, const int32_t* __restrict__ idx32, QpWork W,
               double* __restrict__ dx_out, int* __restrict__ iters_out, int* __restrict__ status_out, const int* __restrict__ fail) {
  extern __shared__ double sm[];
  const PlmLayout& L = *tab.layout;
  const QpLayout& Q = *Qp;
  const int b = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int n = L.n, m = L.m, ndx = L.ndx, N = L.nodes, smax = Q.smax;
  const int pdb = LAT ? Q.panel_doubles_lat : Q.panel_doubles;
  const int gd = Q.g_doubles;
  const bool sparse = Q.sparse_coupling != 0, general = Q.general_coupling != 0;
  // panel buffers and coupling-block buffers first (16-byte aligned), then the mbarriers, then the gathered vectors.
  // Throughput kernel: the row vector w is dead during the sweeps and no panel is resident outside them, so w ALIASES
  // the panel buffers (the ring is filled at the start of each iteration's sweeps and runs empty at their end); this
  // is what lets four CTAs share an SM.  The latency kernel has shared memory to spare and prefetches across iterations.
  constexpr bool ALIAS = !LAT;
  const int ring = NB * (pdb + gd);
  const int ring_al = ALIAS ? ((max(ring, m) + 1) & ~1) : ring;
  double* pbuf = sm;                                   // [NB][pdb]
  double* gbuf = sm + NB * pdb;                      // [NB][gd] compact coupling block travelling with a stage's first panel
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + ring_al);  // [NB] "panel landed" barriers
  int* cnt = reinterpret_cast<int*>(sm + ring_al + NB);                          // [NB] warps done with the panel (running count)
  // the schedule of one iteration, 16 bytes per step, lives in shared memory: the step loop and the refills then run
  // without global loads (packed from the host table at kernel start)
  const int nsched = LAT ? Q.n_sched_lat : Q.n_sched;
  uint4* sch = reinterpret_cast<uint4*>(sm + ring_al + 2 * NB);     // [nsched] {offset, len | r0 << 16 | r1 << 24, stage | flags << 8 | s_prev << 16 | s << 24, shift | x_off << 16}
  double* xt = sm + ring_al + 2 * NB + 2 * nsched;        // [n]  rhs -> forward solution y -> x~ -> delta_x
  double* w = ALIAS ? sm : xt + n;          // [m]  rho z - y, then z~ = A x~, then delta_y
  constexpr int SYM_PARTS = NT / SYM_K;
  double* cpart = xt + n + (ALIAS ? 0 : m);       // [SYM_PARTS][smax] partial sums of the forward product, one slice per part
  double* vd = cpart + SYM_PARTS * smax;      // [2 smax] input of the forward product, stored twice in a row (sym_panel)
  double* red = vd + 2 * smax;                // [16] (one entry per warp, at most 16 warps)
  double* gcs = red + 16;      // [ncoup_max] general coupling only: rho_q (a_q . tv_{i-1}) of the coupling rows
  const double* Ah = W.Ahat + (size_t)b * L.nnz;
  const double* AT = W.AhatT + (size_t)b * L.nnz;
  const double* AR = W.AhatR + (size_t)b * Q.rell_total;
  const double* AC = W.AhatC + (size_t)b * Q.cell_total;
  FlatIdx F;
  F.rptr = idx32 + Q.f_rptr; F.tptr = idx32 + Q.f_tptr; F.rcol = idx + Q.f_rcol; F.trow = idx + Q.f_trow; F.rperm = idx + Q.f_rperm; F.cperm = idx + Q.f_cperm;
  const int32_t* sched = idx32 + (LAT ? Q.f_sched_lat : Q.f_sched);
  const double* Ph = W.Ph + (size_t)b * n;
  const double* qh = W.qh + (size_t)b * n;
  const double* lh = W.lh + (size_t)b * m;
  const double* uh = W.uh + (size_t)b * m;
  const double* rho = W.rho + (size_t)b * m;
  const double* Dv = W.D + (size_t)b * n;
  const double* Ev = W.E + (size_t)b * m;
  const double* Lf = W.Linv + (size_t)b * Q.fac_total;
  const double cs = W.cscale[b];
  // persistent scaled iterates live in HBM: they are only streamed (element-wise), never gathered
  double* x = W.x + (size_t)b * n;
  double* z = W.z + (size_t)b * m;
  double* y = W.y + (size_t)b * m;
  const double alpha = Q.alpha, sigma = Q.sigma;
  // ---- compact coupling blocks G_i[c][j] = rho_c n_c A_int[c][j-th own-stage entry] (the ADMM iterations only ever
  // need these products); written once per solve, then streamed with the panels.
  const int gld = Q.gdense_ld;      // dense integrator rows: leading dimension of the zero-filled coupling blocks
  double* Gc = W.Gc + (size_t)b * N * max(gd, ndx * gld);
  if (sparse) {
    for (int q = tid; q < N * ndx; q += nth) {
      const int i = q / ndx, c2 = q - i * ndx;
      const StageView sp = stage_view(L, Q, idx, i);
      const double* Ap = Ah + L.nnz_off[i];
      const int e0 = sp.rptr[c2], e1 = sp.rptr[c2 + 1] - 1;
      const double coef = rho[L.row_off[i] + c2] * Ap[e1];
#pragma unroll
      for (int j = 0; j < 4; ++j) Gc[(size_t)i * gd + 5 * c2 + j] = (e0 + j < e1) ? coef * Ap[e0 + j] : 0.0;
      // the own-stage columns of the (up to) four entries ride along in the fifth word: the coupling step of the
      // iterations then runs without a single global load
      unsigned long long pk = 0ull;
#pragma unroll
      for (int j = 0; j < 4; ++j) pk |= (unsigned long long)(unsigned short)sp.ccol[e0 + (e0 + j < e1 ? j : 0)] << (16 * j);
      Gc[(size_t)i * gd + 5 * c2 + 4] = __longlong_as_double((long long)pk);
    }
    asm volatile("fence.proxy.async;" ::: "memory");     // the bulk copies below read Gc through the async proxy
  }
  if (gld > 0) {
    // dense integrator rows (whole_body_aba, centroidal_vel): zero-filled blocks G_i[c][k] = rho_c n_c A_int[c][k], k < s_i,
    // streamed through the panel ring ahead of stage i + 1 (coupling panels of the schedule)
    for (int q = tid; q < N * ndx * gld; q += nth) Gc[q] = 0.0;
    __syncthreads();
    for (int q = tid >> 3; q < N * ndx; q += nth >> 3) {
      const int i = q / ndx, c2 = q - i * ndx;
      const StageView sp = stage_view(L, Q, idx, i);
      const double* Ap = Ah + L.nnz_off[i];
      const int e1 = sp.rptr[c2 + 1] - 1;
      const double coef = rho[L.row_off[i] + c2] * Ap[e1];
      for (int e = sp.rptr[c2] + (tid & 7); e < e1; e += 8) Gc[(size_t)q * gld + sp.ccol[e]] = coef * Ap[e];
    }
    __syncthreads();
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  // ---- panel pipeline.  `used` counts schedule steps (uniform across the CTA); step q lives in buffer q % NB.  Warps
  // consume the panels of a stage at their own pace (no CTA barrier between panels): every warp waits on the buffer's
  // mbarrier, and the last warp to finish step q refills the buffer with step q + NB.
  if (tid == 0) {
    for (int k = 0; k < NB; ++k) { mbar_init(&bars[k], 1); cnt[k] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int j = tid; j < (int)(gcs - xt); j += nth) xt[j] = 0.0;
  for (int st = tid; st < nsched; st += nth) {
    const int4 S0 = __ldg(reinterpret_cast<const int4*>(sched + st * PLM_SCHED_INTS));
    const int4 S1 = __ldg(reinterpret_cast<const int4*>(sched + st * PLM_SCHED_INTS) + 1);
    sch[st] = make_uint4((unsigned)S0.x, (unsigned)S0.y | ((unsigned)S0.z << 16) | ((unsigned)S0.w << 24),
                         (unsigned)S1.x | (((unsigned)S1.y & 255u) << 8) | (((unsigned)S1.y >> 8) << 16) | (((unsigned)S1.w & 255u) << 24),
                         (unsigned)S1.z | (((unsigned)S1.w >> 8) << 16));
  }
  if (ALIAS) for (int j = tid; j < ring_al; j += nth) sm[j] = 0.0;
  __syncthreads();
  unsigned used = 0;
  auto issue_step = [&](int st, int buf) {            // one thread: schedule step st into buffer buf
    const uint4 E = sch[st];
    const unsigned bytes = (E.y & 0xffffu) * 8u;
    const int i = (int)(E.z & 255u), fl = (int)((E.z >> 8) & 255u);
    const bool with_g = sparse && (fl & 2) && !(fl & 1) && i > 0;     // forward coupling b_i -= G_{i-1} tv_{i-1}
    mbar_expect_tx(&bars[buf], bytes + (with_g ? (unsigned)gd * 8u : 0u));
    bulk_g2s(pbuf + (size_t)buf * pdb, ((fl & 8) ? Gc : Lf) + E.x, bytes, &bars[buf]);
    if (with_g) bulk_g2s(gbuf + (size_t)buf * gd, Gc + (size_t)(i - 1) * gd, (unsigned)gd * 8u, &bars[buf]);
    if (PLM_ADMM_PREFETCH_STEPS > 0) {      // the panel PLM_ADMM_PREFETCH_STEPS steps further on (wraps into the next iteration)
      int pf = st + PLM_ADMM_PREFETCH_STEPS;
      if (pf >= nsched) pf -= nsched;
      const uint4 P0 = sch[pf];
      bulk_prefetch_l2(Lf + P0.x, (P0.y & 0xffffu) * 8u);
    }
  };
  if (!ALIAS && tid == 0)
    for (int k = 0; k < NB; ++k) issue_step(k, k);
  constexpr int nwarps = NT / 32;
  int status = 0, it = 0;
  bool w_ready = false;   // shared w already holds rho z - y of the current iterates
  bool xq_ready = false;  // shared xt already holds sigma x - q of the current iterates
  double ndx_max = 0.0;   // ||D dx||_inf of the last iteration (dual infeasibility test)
  PROF_T0();
  for (it = 1; it <= Q.max_iter; ++it) {
    PROF_ADD(15);
    // ---- rhs = sigma x - q + A^T (rho z - y); w = rho z - y is left in shared memory by the update phase of the
    // previous iteration unless that one ran the termination tests (which need delta_y there)
    if (!w_ready) {
    for (int r0 = tid; r0 < m; r0 += 4 * nth) {
      double a[4], c[4], d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = r0 + j * nth;
        const bool ok = r < m;
        a[j] = ok ? __ldg(rho + r) : 0.0;
        c[j] = ok ? z[r] : 0.0;
        d[j] = ok ? y[r] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = r0 + j * nth;
        if (r < m) w[r] = a[j] * c[j] - d[j];
      }
    }
    __syncthreads();
    }
    // (sigma x - q is left in xt by the update phase of the previous iteration, like w)
    spmv_ell<true>(idx32 + Q.f_cell_base, idx + Q.f_cell_ind, F.cperm, n, Q.n_cslices, AC, w, xt, sigma, x, qh, xq_ready ? xt : nullptr);
    __syncthreads();
    if (ALIAS && tid == 0) {     // w is dead: the ring takes over its shared memory (generic accesses before async writes)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      for (int k = 0; k < NB; ++k) issue_step(k, (int)((used + k) % NB));
    }
    PROF_ADD(0);
    // ---- forward and backward sweeps, one schedule step = one row panel of one inverse stage block
    double acc0 = 0.0, acc1 = 0.0;
    int bk = -1;                              // backward steps: output of this thread (-1: none)
    int pend = -1, pend_st = 0;               // lane 0: deferred refill check of the previous step
    for (int st = 0; st < nsched; ++st) {
      const uint4 E = sch[st];
      const int r0 = (int)((E.y >> 16) & 255u), r1 = (int)(E.y >> 24), i = (int)(E.z & 255u), fl = (int)((E.z >> 8) & 255u);
      const int dir = fl & 1, first = fl & 2, last = fl & 4, shift = (int)(E.w & 0xffffu);
      const int sprev_sched = (int)((E.z >> 16) & 255u);      // forward steps: size of the previous stage
      const int s = (int)(E.z >> 24);
      double* bi = xt + (E.w >> 16);
      const int bsel = (int)(used % NB);
      if (first && dir) bk = tid < s ? tid : -1;     // backward: one warp per 32 outputs over all the columns (no partial
                                                     // sums: the result goes straight into x_i and the stage needs one CTA barrier;
                                                     // dealing the columns to the two halves of every warp, all eight warps busy and one
                                                     // shuffle at the end, measured +1 %)
      mbar_wait(&bars[bsel], (used / NB) & 1u);
      if (pend >= 0) {      // the previous step's buffer: refill it if this warp was the last one out
        if ((pend + 1) % nwarps == 0) {
          const int nx = pend_st + NB;
          if (!ALIAS || nx < nsched) issue_step(nx >= nsched ? nx - nsched : nx, (int)((used - 1) % NB));
        }
        pend = -1;
      }
      PROF_ADD(2);
      // tv_{i-1} still sits in the partial sums of the parts: they are added up on read, which saves the combine pass and
      // its CTA barrier; the sum also goes to stage i-1's slice of xt, where the backward sweep expects it
      const double* pp = cpart;
      auto tprev = [&](int k) {
        double v = pp[k];
#pragma unroll
        for (int w2 = 1; w2 < SYM_PARTS; ++w2) v += pp[w2 * smax + k];
        return v;
      };
      const bool coup = (fl & 8) != 0;
      double csum = 0.0;
      if (coup) {
        // coupling panel (dense integrator rows): rows [r0, r1) of G_{i-1}, b_i[c] -= G_{i-1}[c][:] . tv_{i-1}, eight lanes
        // per row; the first one of a stage also moves tv_{i-1} to xt and the rest of b_i to vd
        const int sprev = sprev_sched;
        if (fl & 16) {
          if (tid < sprev) bi[tid - sprev] = tprev(tid);
          const int c = ndx + tid;
          if (c < s) { const double v = bi[c]; vd[c] = v; vd[c + s] = v; }
        }
        const double* pan = pbuf + bsel * pdb;
        const int sub = tid & 7;
        for (int c0 = r0; c0 < r1; c0 += nth >> 3) {
          const int c2 = c0 + (tid >> 3);
          double acc = 0.0;
          if (c2 < r1) {
            const double* gr = pan + (c2 - r0) * shift;      // (the seventh schedule word of a coupling panel: leading dimension)
            for (int k = sub; k < sprev; k += 8) acc += gr[k] * tprev(k);
          }
          acc += __shfl_xor_sync(0xffffffffu, acc, 4);
          acc += __shfl_xor_sync(0xffffffffu, acc, 2);
          acc += __shfl_xor_sync(0xffffffffu, acc, 1);
          if (c2 < r1 && sub == 0) {
            const double nv = bi[c2] - acc;
            vd[c2] = nv; vd[c2 + s] = nv;
          }
          csum += acc;
        }
        PROF_ADD(11);
      } else if (first) {
        const double* g = gbuf + bsel * gd;
        if (dir == 0) {
          // the input b_i of the product goes to vd twice in a row (sym_panel); its first ndx entries come out of the
          // coupling step below (dense integrator rows: out of the coupling panels, which also moved the rest)
          if (!(gld > 0 && i > 0)) {
            const int c = (i > 0 ? ndx : 0) + tid;
            if (c < s) { const double v = bi[c]; vd[c] = v; vd[c + s] = v; }
          }
          if (i > 0 && gld == 0) {     // b_i -= G_{i-1} tv_{i-1}
            const int sprev = sprev_sched;      // size of stage i - 1 (schedule entry)
            if (tid < sprev) bi[tid - sprev] = tprev(tid);      // (stage i - 1's slice of xt ends where b_i starts)
            if (general) {
              const StageView sp = stage_view(L, Q, idx, i - 1);
              // b_i -= sum_q n_q rho_q (a_q . tv_{i-1}) over the coupling rows q of node i-1 (see the factor kernel)
              const QpTypeIdx& I = Q.type[L.node_type[i - 1]];
              const int16_t* crows = idx + I.gc_rows;
              const int16_t* rowq = idx + I.gc_rowq;
              const double* Ap = Ah + L.nnz_off[i - 1];
              const double* rp = rho + L.row_off[i - 1];
              const int sub = tid & 7;
              for (int q0 = 0; q0 < I.ncoup; q0 += nth >> 3) {
                const int q = q0 + (tid >> 3);
                double acc = 0.0;
                int r = 0;
                if (q < I.ncoup) {
                  r = crows[q];
                  for (int e = sp.rptr[r] + sub; e < sp.rptr[r + 1]; e += 8) {
                    const int k = sp.ccol[e];
                    if (k < sprev) acc += Ap[e] * tprev(k);
                  }
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                if (q < I.ncoup && sub == 0) gcs[q] = rp[r] * acc;
              }
              __syncthreads();
              if (tid < ndx) {
                double acc = 0.0;
                for (int e = sp.cptr[sprev + tid]; e < sp.cptr[sprev + tid + 1]; ++e) acc += Ap[sp.cpos[e]] * gcs[rowq[sp.crow[e]]];
                const double nv = bi[tid] - acc;
                vd[tid] = nv; vd[tid + s] = nv;
              }
            } else if (sparse) {
              if (tid < ndx) {
                const double* gr = g + 5 * tid;
                const unsigned long long pk = (unsigned long long)__double_as_longlong(gr[4]);
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) acc += gr[j] * tprev((int)((pk >> (16 * j)) & 0xffffull));
                const double nv = bi[tid] - acc;
                vd[tid] = nv; vd[tid + s] = nv;
              }
            }
          }
          __syncthreads();
          PROF_ADD(1);
        }
        acc0 = 0.0; acc1 = 0.0;
      }
      if (coup) {
      } else if (dir == 0) {
        if ((tid & (SYM_K - 1)) < s) sym_panel<SYM_PARTS>(pbuf + bsel * pdb - shift, s, r0, r1, vd, tid & (SYM_K - 1), tid / SYM_K, acc0, acc1);
      }
      else if (bk >= 0) rect_panel(pbuf + bsel * pdb, shift, r0, r1, bi + s, bk, acc0, acc1);   // x_i = tv_i - B_i x_{i+1}[0:ndx]
      if (!coup) { if (dir == 0) PROF_ADD(9); else PROF_ADD(5); }
      {
        // release the buffer: the count is bumped by an instruction that depends on the sums, i.e. after every
        // shared-memory read of the panel by this warp has returned; the refill check is deferred past the next wait
        const double sum = coup ? csum : acc0 + acc1;
        int zero;
        asm volatile("and.b32 %0, %1, 0;" : "=r"(zero) : "r"(__double2loint(sum)));
        __syncwarp();
        if ((tid & 31) == 0) { pend = atomicAdd(&cnt[bsel], 1 + zero); pend_st = st; }
        ++used;
        PROF_ADD(3);
        if (last) {
          // forward: the sum of this part goes to its slice (the parts are added up by the reader)
          if (dir == 0) {
            if ((tid & (SYM_K - 1)) < s) cpart[(tid / SYM_K) * smax + (tid & (SYM_K - 1))] = sum;
          }
          else if (bk >= 0) bi[bk] -= sum;       // nothing else reads stage i's slice of xt during its backward step
        }
      }
      if (last) {
        if (pend >= 0) {
          if ((pend + 1) % nwarps == 0) {
            const int nx = pend_st + NB;
            if (!ALIAS || nx < nsched) issue_step(nx >= nsched ? nx - nsched : nx, bsel);
          }
          pend = -1;
        }
        __syncthreads();
        PROF_ADD(10);
        // backward stages are complete; forward stages leave the partial sums to the next stage's coupling step, except
        // the last one (x_N = tv_N is the first input of the backward sweep)
        if (dir == 1 || i < N) continue;
        if (tid < s) {
          double o = cpart[tid];
#pragma unroll
          for (int w2 = 1; w2 < SYM_PARTS; ++w2) o += cpart[w2 * smax + tid];
          bi[tid] = o;
        }
        __syncthreads();
        PROF_ADD(4);
      }
    }
    // ---- z~ = A x~ ; relaxation, projection, dual update
    const bool check = (Q.check_termination > 0 && it % Q.check_termination == 0) || it == Q.max_iter;
    spmv_ell<false>(idx32 + Q.f_rell_base, idx + Q.f_rell_ind, F.rperm, m, Q.n_rslices, AR, xt, w, 0.0, nullptr, nullptr);
    __syncthreads();
    PROF_ADD(6);
    double mdx = 0.0;
    for (int j0 = tid; j0 < n; j0 += 4 * nth) {
      double xo[4], dv[4];      // dv: D (iterations that run the termination tests) or q (the others)
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int j = j0 + q4 * nth;
        const bool ok = j < n;
        xo[q4] = ok ? x[j] : 0.0;
        dv[q4] = ok ? __ldg((check ? Dv : qh) + j) : 0.0;
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int j = j0 + q4 * nth;
        if (j < n) {
          const double xn = alpha * xt[j] + (1.0 - alpha) * xo[q4];
          if (check) {
            mdx = fmax(mdx, fabs(dv[q4] * (xn - xo[q4])));
            xt[j] = xn - xo[q4];      // delta_x (kept for the dual infeasibility test)
          } else xt[j] = sigma * xn - dv[q4];      // the x part of the next right-hand side
          x[j] = xn;
        }
      }
    }
    for (int r0 = tid; r0 < m; r0 += 4 * nth) {
      double zo[4], yo[4], rr[4], lo[4], up[4];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int r = r0 + q4 * nth;
        const bool ok = r < m;
        zo[q4] = ok ? z[r] : 0.0;
        yo[q4] = ok ? y[r] : 0.0;
        rr[q4] = ok ? __ldg(rho + r) : 1.0;      // (reading a one-byte class of rho_vec instead, three possible values: measured 21.7 -> 22.5 ms per wave, the selects and spills cost more than the 19 KB per iteration saved)
        lo[q4] = ok ? __ldg(lh + r) : 0.0;
        up[q4] = ok ? __ldg(uh + r) : 0.0;
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int r = r0 + q4 * nth;
        if (r < m) {
          const double zr = alpha * w[r] + (1.0 - alpha) * zo[q4];
          double zn = zr + yo[q4] / rr[q4];
          zn = fmin(fmax(zn, lo[q4]), up[q4]);
          const double dy = rr[q4] * (zr - zn);
          const double yn = yo[q4] + dy;
          y[r] = yn;
          z[r] = zn;
          w[r] = check ? dy : rr[q4] * zn - yn;     // delta_y for the primal infeasibility test, else the next rhs term
        }
      }
    }
    __syncthreads();
    PROF_ADD(7);
    w_ready = !check;
    xq_ready = !check;
    if (!check) continue;
    const bool approx = !(Q.check_termination > 0 && it % Q.check_termination == 0);
    ndx_max = block_reduce(mdx, red, true);
    // two passes when the last iteration is also a regular check: exact first, then approximate
    for (int pass = 0; pass < 2 && status == 0; ++pass) {
      const bool apx = approx || pass == 1;
      if (pass == 1 && it != Q.max_iter) break;
      const double kk = apx ? 10.0 : 1.0;
      const double eps_abs = Q.eps_abs * kk, eps_rel = Q.eps_rel * kk, eps_pinf = Q.eps_prim_inf * kk, eps_dinf = Q.eps_dual_inf * kk;
      // primal residual ||E^-1 (A x - z)||, norms of E^-1 z and E^-1 A x
      double pr = 0.0, nz = 0.0, nax = 0.0;
      for (int i2 = tid; i2 < m; i2 += nth) {       // rows by decreasing length: balanced warps (max-reductions only)
        const int r = F.rperm[i2];
        double ax = 0.0;
        for (int e = F.rptr[r]; e < F.rptr[r + 1]; ++e) ax += Ah[e] * x[F.rcol[e]];
        const double ei = 1.0 / Ev[r];
        pr = fmax(pr, fabs((ax - z[r]) * ei));
        nz = fmax(nz, fabs(z[r] * ei));
        nax = fmax(nax, fabs(ax * ei));
      }
      pr = block_reduce(pr, red, true);
      nz = block_reduce(nz, red, true);
      nax = block_reduce(nax, red, true);
      // dual residual ||D^-1 (P x + q + A^T y)|| / c
      double dr = 0.0, nq = 0.0, naty = 0.0, npx = 0.0;
      for (int i2 = tid; i2 < n; i2 += nth) {
        const int j = F.cperm[i2];
        double acc = 0.0;
        for (int e = F.tptr[j]; e < F.tptr[j + 1]; ++e) acc += AT[e] * y[F.trow[e]];
        const double di = 1.0 / Dv[j];
        const double px = Ph[j] * x[j];
        dr = fmax(dr, fabs((px + qh[j] + acc) * di));
        nq = fmax(nq, fabs(qh[j] * di));
        naty = fmax(naty, fabs(acc * di));
        npx = fmax(npx, fabs(px * di));
      }
      dr = block_reduce(dr, red, true) / cs;
      nq = block_reduce(nq, red, true);
      naty = block_reduce(naty, red, true);
      npx = block_reduce(npx, red, true);
      const double eps_pri = eps_abs + eps_rel * fmax(nz, nax);
      const double eps_dua = eps_abs + eps_rel * fmax(nq, fmax(naty, npx)) / cs;
      const bool prim_ok = pr < eps_pri, dual_ok = dr < eps_dua;
      bool prim_inf = false, dual_inf = false;
      if (!prim_ok) {
        // primal infeasibility certificate: delta_y (w) is projected in place on the polar of the recession cone of
        // [l,u], exactly as osqp does with work->delta_y
        double ndy = 0.0, lhs = 0.0;
        for (int r = tid; r < m; r += nth) {
          double dy = w[r];
          const bool up = uh[r] > OSQP_INFTY * MIN_SCALING, lo = lh[r] < -OSQP_INFTY * MIN_SCALING;
          if (up && lo) dy = 0.0;
          else if (up) dy = fmin(dy, 0.0);
          else if (lo) dy = fmax(dy, 0.0);
          w[r] = dy;
          ndy = fmax(ndy, fabs(Ev[r] * dy));
          lhs += uh[r] * fmax(dy, 0.0) + lh[r] * fmin(dy, 0.0);
        }
        ndy = block_reduce(ndy, red, true);     // (barriers inside make w visible)
        lhs = block_reduce(lhs, red, false);
        if (ndy > eps_pinf && lhs < -eps_pinf * ndy) {
          double na = 0.0;
          for (int j = tid; j < n; j += nth) {
            double acc = 0.0;
            for (int e = F.tptr[j]; e < F.tptr[j + 1]; ++e) acc += AT[e] * w[F.trow[e]];
            na = fmax(na, fabs(acc / Dv[j]));
          }
          na = block_reduce(na, red, true);
          prim_inf = na < eps_pinf * ndy;
        }
      }
      if (!dual_ok && !prim_inf) {
        // dual infeasibility certificate on delta_x (xt)
        if (ndx_max > eps_dinf) {
          double qd = 0.0, npd = 0.0;
          for (int j = tid; j < n; j += nth) {
            qd += qh[j] * xt[j];
            npd = fmax(npd, fabs(Ph[j] * xt[j] / Dv[j]));
          }
          qd = block_reduce(qd, red, false);
          npd = block_reduce(npd, red, true);
          if (qd < -cs * eps_dinf * ndx_max && npd < cs * eps_dinf * ndx_max) {
            double bad = 0.0;
            for (int r = tid; r < m; r += nth) {
              double ax = 0.0;
              for (int e = F.rptr[r]; e < F.rptr[r + 1]; ++e) ax += Ah[e] * xt[F.rcol[e]];
              ax /= Ev[r];
              if ((uh[r] < OSQP_INFTY * MIN_SCALING && ax > eps_dinf * ndx_max) ||
                  (lh[r] > -OSQP_INFTY * MIN_SCALING && ax < -eps_dinf * ndx_max)) bad = 1.0;
            }
            bad = block_reduce(bad, red, true);
            dual_inf = bad == 0.0;
          }
        }
      }
      if (prim_ok && dual_ok) status = apx ? 2 : 1;
      else if (prim_inf) status = apx ? 3 : -3;
      else if (dual_inf) status = apx ? 4 : -4;
    }
    PROF_ADD(8);
    if (status != 0) break;
  }
  if (!ALIAS)
    for (unsigned k = used; k < used + NB; ++k) mbar_wait(&bars[k % NB], (k / NB) & 1u);   // drain prefetches in flight
  if (it > Q.max_iter) it = Q.max_iter;
  if (status == 0) status = -2;   // maximum iterations reached
  // failure detection beyond osqp's own codes: -10 the stage factorisation met a non-positive pivot, -11 NaN iterates
  {
    double bad = 0.0;      // max-reductions ignore NaN, so look for it explicitly once at the end
    for (int j = tid; j < n; j += nth) if (x[j] != x[j]) bad = 1.0;
    bad = block_reduce(bad, red, true);
    if (fail[b] != 0) status = -10;
    else if (bad != 0.0) status = -11;
  }
  const bool no_solution = (status == 3 || status == -3 || status == 4 || status == -4);
  __syncthreads();
  for (int j = tid; j < n; j += nth) {
    if (dx_out) dx_out[(size_t)b * n + j] = no_solution ? nan("") : Dv[j] * x[j];
    if (no_solution) x[j] = 0.0;      // store_solution(): no solution -> cold start
  }
  if (no_solution)
    for (int r = tid; r < m; r += nth) { z[r] = 0.0; y[r] = 0.0; }
  if (tid == 0) {
    if (iters_out) iters_out[b] = it;
    if (status_out) status_out[b] = status;
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
#define QP_CUDA(h, expr)                                                               \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      (h)->error = std::string(#expr) + ": " + cudaGetErrorString(_e);                 \
      return 7;                                                                        \
    }                                                                                  \
  } while (0)

int plm_qp_alloc(plm_handle* h) {
  const PlmLayout& L = h->host.layout;
  const QpLayout& Q = h->host.qp;
  QpWork& W = h->qp;
  const size_t B = (size_t)h->max_batch;
  QP_CUDA(h, cudaMalloc(&W.d_ql, sizeof(QpLayout)));
  QP_CUDA(h, cudaMemcpy(W.d_ql, &Q, sizeof(QpLayout), cudaMemcpyHostToDevice));
  QP_CUDA(h, cudaMalloc(&W.d_idx, h->host.qp_idx.size() * sizeof(int16_t)));
  QP_CUDA(h, cudaMemcpy(W.d_idx, h->host.qp_idx.data(), h->host.qp_idx.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
  QP_CUDA(h, cudaMalloc(&W.d_idx32, h->host.qp_idx32.size() * sizeof(int32_t)));
  QP_CUDA(h, cudaMemcpy(W.d_idx32, h->host.qp_idx32.data(), h->host.qp_idx32.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  auto al = [&](double** p, size_t per) { return cudaMalloc(p, B * per * sizeof(double)); };
  QP_CUDA(h, al(&W.Ahat, L.nnz)); QP_CUDA(h, al(&W.AhatT, L.nnz)); QP_CUDA(h, al(&W.AhatR, Q.rell_total)); QP_CUDA(h, al(&W.AhatC, Q.cell_total)); QP_CUDA(h, al(&W.D, L.n)); QP_CUDA(h, al(&W.E, L.m)); QP_CUDA(h, al(&W.Eprev, L.m));
  QP_CUDA(h, al(&W.cscale, 1)); QP_CUDA(h, al(&W.Ph, L.n)); QP_CUDA(h, al(&W.qh, L.n)); QP_CUDA(h, al(&W.lh, L.m));
  QP_CUDA(h, al(&W.uh, L.m)); QP_CUDA(h, al(&W.rho, L.m)); QP_CUDA(h, al(&W.Linv, Q.fac_total)); QP_CUDA(h, al(&W.Gc, (size_t)L.nodes * std::max((int)Q.g_doubles, (int)(L.ndx * Q.gdense_ld))));
  QP_CUDA(h, al(&W.x, L.n)); QP_CUDA(h, al(&W.z, L.m)); QP_CUDA(h, al(&W.y, L.m));
  QP_CUDA(h, cudaMalloc(&h->d_qp_fail, B * sizeof(int)));
  QP_CUDA(h, cudaMemset(W.x, 0, B * L.n * sizeof(double)));
  QP_CUDA(h, cudaMemset(W.z, 0, B * L.m * sizeof(double)));
  QP_CUDA(h, cudaMemset(W.y, 0, B * L.m * sizeof(double)));
  QP_CUDA(h, cudaMemset(h->d_qp_fail, 0, B * sizeof(int)));
  {
    // Eprev = 1 until plm_qp_setup computes the setup-time scaling
    std::vector<double> ones(B * L.m, 1.0);
    QP_CUDA(h, cudaMemcpy(W.Eprev, ones.data(), ones.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  W.allocated = 1;
  h->qp_factor_doubles = Q.fac_total;
  const int smax = Q.smax, ndx = L.ndx;
  h->smem_scale = (size_t)(L.n + L.m + 32) * 8;
  // staging J in shared memory (1 CTA/SM) loses against reading it from L2 with 5-6 resident CTAs per SM (measured)
  const int wld = ((ndx + 15) & ~15) + 8;
  h->smem_factor = (size_t)(smax * (smax + 1) / 2 + smax * wld + std::max(ndx * (ndx + 1) / 2, 8 * smax + 72) + ndx + L.max_rows + 2) * 8;
  if (Q.sparse_coupling) h->smem_factor -= (size_t)smax * wld * 8;     // no W buffer
  if (Q.general_coupling) h->smem_factor += (size_t)(smax * Q.ncoup_max + Q.ncoup_max * ndx) * 8;     // Y, Nn
  // throughput kernel: w aliases the panel ring
  const int gcn = Q.general_coupling ? Q.ncoup_max : 0;
  h->smem_admm = (size_t)(((std::max(NBUF * (Q.panel_doubles + Q.g_doubles), L.m) + 1) & ~1) + 2 * NBUF + 2 * Q.n_sched + L.n + (ADMM_THREADS / SYM_K + 2) * smax + 16 + gcn) * 8;
  h->smem_admm_lat = (size_t)(NBUF_LAT * (Q.panel_doubles_lat + Q.g_doubles) + 2 * NBUF_LAT + 2 * Q.n_sched_lat + L.n + L.m + (ADMM_THREADS_LAT / SYM_K + 2) * smax + 16 + gcn) * 8;
  if (smax > SYM_K) { h->error = "stage size exceeds the thread-column capacity of the ADMM kernel"; return 7; }
  if (h->smem_scale > 227 * 1024 || h->smem_factor > 227 * 1024 || h->smem_admm > 227 * 1024 || h->smem_admm_lat > 227 * 1024) {
    h->error = "QP workspace exceeds shared memory";
    return 7;
  }
  QP_CUDA(h, cudaFuncSetAttribute(qp_scale_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  QP_CUDA(h, cudaFuncSetAttribute(qp_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  QP_CUDA(h, cudaFuncSetAttribute(qp_admm_kernel<ADMM_THREADS, ADMM_MIN_CTAS, NBUF, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  QP_CUDA(h, cudaFuncSetAttribute(qp_admm_kernel<ADMM_THREADS_LAT, 1, NBUF_LAT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return 0;
}

void plm_qp_free(plm_handle* h) {
  QpWork& W = h->qp;
  cudaFree(W.d_ql); cudaFree(W.d_idx); cudaFree(W.d_idx32); cudaFree(W.AhatT); cudaFree(W.Ahat); cudaFree(W.D); cudaFree(W.E); cudaFree(W.Eprev);
  cudaFree(W.cscale); cudaFree(W.Ph); cudaFree(W.qh); cudaFree(W.lh); cudaFree(W.uh); cudaFree(W.rho); cudaFree(W.Linv); cudaFree(W.Gc); cudaFree(W.AhatR); cudaFree(W.AhatC);
  cudaFree(W.x); cudaFree(W.z); cudaFree(W.y); cudaFree(h->d_qp_fail);
}

// osqp setup of instances [first, first + count): d_hess is the base of the [max_batch][n] array
int plm_qp_setup_impl(plm_handle* h, int first, int count, const double* d_hess, cudaStream_t s) {
  const PlmLayout& L = h->host.layout;
  QpWork& W = h->qp;
  QP_CUDA(h, cudaMemsetAsync(W.x + (size_t)first * L.n, 0, (size_t)count * L.n * sizeof(double), s));
  QP_CUDA(h, cudaMemsetAsync(W.z + (size_t)first * L.m, 0, (size_t)count * L.m * sizeof(double), s));
  QP_CUDA(h, cudaMemsetAsync(W.y + (size_t)first * L.m, 0, (size_t)count * L.m * sizeof(double), s));
  qp_scale_kernel<<<count, PLM_SCALE_THREADS, h->smem_scale, s>>>(h->tab, W.d_ql, W.d_idx, W.d_idx32, 1, first, d_hess, nullptr, nullptr, nullptr, nullptr, W);
  PLM_LAUNCH_CHECK(h);
  return 0;
}

int plm_qp_update_impl(plm_handle* h, int batch, const double* d_hess, const double* d_q, const double* d_J,
                       const double* d_l, const double* d_u, cudaStream_t s) {
  QpWork& W = h->qp;
  qp_scale_kernel<<<batch, PLM_SCALE_THREADS, h->smem_scale, s>>>(h->tab, W.d_ql, W.d_idx, W.d_idx32, 0, 0, d_hess, d_q, d_J, d_l, d_u, W);
  PLM_LAUNCH_CHECK(h);
  qp_factor_kernel<<<batch, QP_THREADS, h->smem_factor, s>>>(h->tab, W.d_ql, W.d_idx, W, h->d_qp_fail);
  PLM_LAUNCH_CHECK(h);
  return 0;
}

int plm_qp_solve_impl(plm_handle* h, int batch, double* d_dx, int* d_iters, int* d_status, cudaStream_t s) {
  QpWork& W = h->qp;
  // fewer instances than SMs: one 512-thread CTA per instance (latency); else three 256-thread CTAs per SM (throughput)
  // (PLM_ADMM_LATENCY_MAX_BATCH overrides the switch-over: 0 forces the throughput kernel; used by the tests)
  int lat_max = h->num_sms;
  if (const char* ev = getenv("PLM_ADMM_LATENCY_MAX_BATCH")) lat_max = atoi(ev);
  if (batch <= lat_max)
    qp_admm_kernel<ADMM_THREADS_LAT, 1, NBUF_LAT, true><<<batch, ADMM_THREADS_LAT, h->smem_admm_lat, s>>>(h->tab, W.d_ql, W.d_idx, W.d_idx32, W, d_dx, d_iters, d_status, h->d_qp_fail);
  else
    qp_admm_kernel<ADMM_THREADS, ADMM_MIN_CTAS, NBUF, false><<<batch, ADMM_THREADS, h->smem_admm, s>>>(h->tab, W.d_ql, W.d_idx, W.d_idx32, W, d_dx, d_iters, d_status, h->d_qp_fail);
  PLM_LAUNCH_CHECK(h);
  return 0;
}

extern "C" {

int plm_qp_setup(plm_handle* h, int32_t batch, const double* d_hess, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  h->qp_setup_batch = batch;
  return plm_qp_setup_impl(h, 0, batch, d_hess, (cudaStream_t)stream);
}

int plm_qp_update(plm_handle* h, int32_t batch, const double* d_hess, const double* d_q, const double* d_J,
                  const double* d_l, const double* d_u, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  return plm_qp_update_impl(h, batch, d_hess, d_q, d_J, d_l, d_u, (cudaStream_t)stream);
}

int plm_qp_solve(plm_handle* h, int32_t batch, double* d_dx, int32_t* d_iters, int32_t* d_status, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  return plm_qp_solve_impl(h, batch, d_dx, d_iters, d_status, (cudaStream_t)stream);
}

int plm_qp_get_iterates(plm_handle* h, int32_t batch, double* d_x, double* d_z, double* d_y, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  const PlmLayout& L = h->host.layout;
  cudaStream_t s = (cudaStream_t)stream;
  QP_CUDA(h, cudaMemcpyAsync(d_x, h->qp.x, (size_t)batch * L.n * 8, cudaMemcpyDeviceToDevice, s));
  QP_CUDA(h, cudaMemcpyAsync(d_z, h->qp.z, (size_t)batch * L.m * 8, cudaMemcpyDeviceToDevice, s));
  QP_CUDA(h, cudaMemcpyAsync(d_y, h->qp.y, (size_t)batch * L.m * 8, cudaMemcpyDeviceToDevice, s));
  return 0;
}

int plm_qp_set_iterates(plm_handle* h, int32_t batch, const double* d_x, const double* d_z, const double* d_y, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  const PlmLayout& L = h->host.layout;
  cudaStream_t s = (cudaStream_t)stream;
  QP_CUDA(h, cudaMemcpyAsync(h->qp.x, d_x, (size_t)batch * L.n * 8, cudaMemcpyDeviceToDevice, s));
  QP_CUDA(h, cudaMemcpyAsync(h->qp.z, d_z, (size_t)batch * L.m * 8, cudaMemcpyDeviceToDevice, s));
  QP_CUDA(h, cudaMemcpyAsync(h->qp.y, d_y, (size_t)batch * L.m * 8, cudaMemcpyDeviceToDevice, s));
  return 0;
}

#ifdef PLM_ADMM_PROFILE
int plm_debug_admm_profile(long long* out16, int reset) {
  if (reset) { long long z[16] = {0}; return (int)cudaMemcpyToSymbol(g_admm_prof, z, sizeof(z)); }
  return (int)cudaMemcpyFromSymbol(out16, g_admm_prof, 16 * sizeof(long long));
}
#endif

/* Debug / test access to the scaling of the last plm_qp_update: D [batch][n], E [batch][m], c [batch]. */
int plm_qp_get_scaling(plm_handle* h, int32_t batch, double* d_D, double* d_E, double* d_c, void* stream) {
  if (batch < 1 || batch > h->max_batch) { h->error = "batch exceeds max_batch of the handle"; return 4; }
  const PlmLayout& L = h->host.layout;
  cudaStream_t s = (cudaStream_t)stream;
  QP_CUDA(h, cudaMemcpyAsync(d_D, h->qp.D, (size_t)batch * L.n * 8, cudaMemcpyDeviceToDevice, s));
  QP_CUDA(h, cudaMemcpyAsync(d_E, h->qp.E, (size_t)batch * L.m * 8, cudaMemcpyDeviceToDevice, s));
  QP_CUDA(h, cudaMemcpyAsync(d_c, h->qp.cscale, (size_t)batch * 8, cudaMemcpyDeviceToDevice, s));
  return 0;
}

}  // extern "C"

"""Prototype (numpy) of the stage-structured ADMM linear algebra used by the CUDA QP kernels:
reduced system (P + sigma I + A^T R A) x = rhs solved by a block-tridiagonal Cholesky whose diagonal
blocks are stored as explicit inverses of their Cholesky factors.  Checked against the oracle's KKT solve."""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from emu_util import *
from oracle.ocp import OracleOCP
from oracle.sqp import OracleSQP
from scipy import sparse

prod, ora = make_robots()
rng = np.random.default_rng(0)
o = OracleOCP(ora['b2'], 'whole_body_rnea', 5)
x, p = random_problem(o, rng)
s = OracleSQP(o); s.init_solver()
grad, J, g, lb, ub = o.sqp_data(x, p)
s.osqp.update(q=grad, Ax=s.csc_values(J), l=lb - g, u=ub - g)
Q = s.osqp
A = Q.A.toarray(); P = Q.P; rho = Q.rho_vec; sigma = 1e-6
n, m = o.n, o.m
H = np.diag(P + sigma) + A.T @ (rho[:, None] * A)
off = list(o.x_off) + [o.n]
N = o.nodes
# block tridiagonal check
for i in range(N + 1):
    for j in range(N + 1):
        blk = H[off[i]:off[i+1], off[j]:off[j+1]]
        if abs(i - j) > 1: assert np.abs(blk).max() == 0
# factor
Linv = []; Kprev = None
for i in range(N + 1):
    S = H[off[i]:off[i+1], off[i]:off[i+1]].copy()
    if i > 0:
        G = H[off[i]:off[i+1], off[i-1]:off[i]]          # (s_i x s_{i-1}), only first ndx rows nonzero
        W = Linv[i-1] @ G.T
        S -= W.T @ W
    L = np.linalg.cholesky(S)
    Linv.append(np.linalg.inv(L))
def solve(b):
    # forward: y_i = Linv_i (b_i - G_{i-1} Linv_{i-1}^T y_{i-1})
    y = []
    for i in range(N + 1):
        r = b[off[i]:off[i+1]].copy()
        if i > 0:
            G = H[off[i]:off[i+1], off[i-1]:off[i]]
            r -= G @ (Linv[i-1].T @ y[i-1])
        y.append(Linv[i] @ r)
    xs = [None] * (N + 1)
    for i in range(N, -1, -1):
        r = y[i].copy()
        if i < N:
            G = H[off[i+1]:off[i+2], off[i]:off[i+1]]
            r -= Linv[i] @ (G.T @ xs[i+1])
        xs[i] = Linv[i].T @ r
    return np.concatenate(xs)
b = rng.normal(size=n)
xs = solve(b)
print('block solve residual', np.abs(H @ xs - b).max() / np.abs(b).max(), 'cond', np.linalg.cond(H))
# compare one ADMM x-update with the oracle KKT solve
xk, zk, yk = rng.normal(size=n), rng.normal(size=m), rng.normal(size=m)
rhs = np.concatenate([sigma * xk - Q.q, zk - yk / rho])
sol = Q.lu.solve(rhs); xt_ref = sol[:n]; zt_ref = zk + (sol[n:] - yk) / rho
xt = solve(sigma * xk - Q.q + A.T @ (rho * zk - yk))
print('x_tilde err', np.abs(xt - xt_ref).max() / np.abs(xt_ref).max(), 'z_tilde err', np.abs(A @ xt - zt_ref).max() / np.abs(zt_ref).max())

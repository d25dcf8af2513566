"""One SQP iteration: sqp_data -> OSQP -> Armijo line search (oracle; test infrastructure only).

Restates optimization/ocp.py:265-319 (init_solver, osqp branch), :375-422 (solve, osqp branch) and
:430-496 (_armijo_line_search, _constraint_violation_metric/_max) including the quirk that f and
g_metric are overwritten with the rejected trial's values (ocp.py:470-471).
"""
import numpy as np
from scipy import sparse

from .osqp_admm import OSQP

OSQP_OPTS = dict(max_iter=100, alpha=1.4, rho=2e-2, warm_start=True, adaptive_rho=False)  # ocp.py:267-273


def constraint_violation_metric(g, lbg, ubg):   # ocp.py:482-488
    v = np.concatenate((np.maximum(0, lbg - g), np.maximum(0, g - ubg)))
    return float(np.linalg.norm(v))


def constraint_violation_max(g, lbg, ubg):      # ocp.py:490-496
    v = np.concatenate((np.maximum(0, lbg - g), np.maximum(0, g - ubg)))
    return float(np.max(np.abs(v)))


def armijo_line_search(ocp, dx, current_x, p):  # ocp.py:430-480
    armijo_factor, a, a_min, a_decay = 1e-4, 1.0, 1e-4, 0.5
    g_max, g_min, gamma = 1e-3, 1e-5, 1e-5
    f, grad_f = ocp.f_data(current_x, p)
    g, lbg, ubg = ocp.g_data(current_x, p)
    g_metric = constraint_violation_metric(g, lbg, ubg)
    armijo_metric = float(grad_f @ dx)
    accepted = False
    trials = 0
    new_x = current_x
    while not accepted and a > a_min:
        new_x = current_x + a * dx
        new_f, _ = ocp.f_data(new_x, p)
        new_g, lbg, ubg = ocp.g_data(new_x, p)
        trials += 1
        new_g_metric = constraint_violation_metric(new_g, lbg, ubg)
        if new_g_metric > g_max:
            if new_g_metric < (1 - gamma) * g_metric:
                accepted = True
        elif max(new_g_metric, g_metric) < g_min and armijo_metric < 0:
            if new_f <= f + armijo_factor * armijo_metric:
                accepted = True
        elif new_f <= f - gamma * new_g_metric or new_g_metric < (1 - gamma) * g_metric:
            accepted = True
        a *= a_decay
        f = new_f
        g_metric = new_g_metric
    info = dict(accepted=accepted, alpha=a / a_decay, trials=trials, f=f, g_metric=g_metric)
    return (new_x if accepted else current_x), info


class OracleSQP:
    def __init__(self, ocp):
        self.ocp = ocp
        self.osqp = None

    def jac_pattern(self, seed=0):
        """Structural pattern of J_g: union of numeric nonzeros at random points with 0<contact<1."""
        o = self.ocp
        rng = np.random.default_rng(seed)
        saved = {k: v.copy() for k, v in o.params.items()}
        pat = None
        for _ in range(2):
            o.params["contact_schedule"][:] = rng.uniform(0.3, 0.7, o.params["contact_schedule"].shape)
            o.params["swing_schedule"][:] = rng.uniform(0.1, 0.9, o.params["swing_schedule"].shape)
            x = o.initial_guess() + rng.normal(size=o.n)
            J = o.jac_g(x, o.p_vector())
            pat = (J != 0) if pat is None else (pat | (J != 0))
        o.params.update(saved)
        return pat

    def init_solver(self):   # ocp.py:292-313
        o = self.ocp
        p = o.p_vector()
        self.hess_diag = o.hess_diag(p)
        self.pattern = self.jac_pattern()
        A = sparse.csc_matrix(self.pattern.astype(float))
        self.A_rows, self.A_cols = A.nonzero()  # not used for ordering; CSC order below
        A.sort_indices()
        self._csc = A
        self.osqp = OSQP()
        self.osqp.setup(self.hess_diag, np.ones(o.n), A, -np.ones(o.m), np.ones(o.m), **OSQP_OPTS)

    def csc_values(self, J):
        """J_g.nonzeros(): values in CSC order of the fixed pattern."""
        A = self._csc
        cols = np.repeat(np.arange(A.shape[1]), np.diff(A.indptr))
        return J[A.indices, cols]

    def solve(self, x, p):
        """Body of the loop at ocp.py:383-406 plus the violation print at :412-414."""
        o = self.ocp
        grad_f, J, g, lbg, ubg = o.sqp_data(x, p)
        self.osqp.update(q=grad_f, Ax=self.csc_values(J), l=lbg - g, u=ubg - g)
        sol_dx = self.osqp.solve()
        new_x, info = armijo_line_search(o, sol_dx, x, p)
        g, lbg, ubg = o.g_data(new_x, p)
        info.update(sol_dx=sol_dx, qp_iters=self.osqp.iters, qp_status=self.osqp.status,
                    violation_max=constraint_violation_max(g, lbg, ubg))
        return new_x, info

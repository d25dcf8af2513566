"""OSQP's ADMM restated (oracle; test infrastructure only).

Third-party algorithm: osqp (conda-forge, UNPINNED in the reference, README.md:16-19); this follows
the published algorithm and the 0.6.x C sources' behaviour [3P] as summarised in SURVEY.md appendix
A.8, for exactly the call pattern of optimization/ocp.py:305-313 (setup with dummy data),
:391-395 (update(q=, Ax=, l=, u=)) and :401 (solve().x):

  * Ruiz equilibration (``scaling`` passes, MIN/MAX_SCALING clamps, cost scaling c) recomputed from
    scratch on every ``update(Ax=)`` (unscale -> overwrite -> scale_data);
  * rho_vec = rho * 1e3 on equality rows (u - l < 1e-4), 1e-6 on free rows, rho otherwise,
    re-derived in update_bounds;
  * iterates (x, z, y) live in the scaled space and persist across solve() calls untouched
    (warm_start=True), even when the scaling changes;
  * termination tested every ``check_termination`` iterations on unscaled residuals, including the
    primal/dual infeasibility certificates; a max-iter iterate is still returned.

The linear system is the quasi-definite KKT system of OSQP solved with a sparse LU (scipy SuperLU)
instead of QDLDL: same solution up to rounding.
"""
import numpy as np
from scipy import sparse
from scipy.sparse.linalg import splu

OSQP_INFTY = 1e30
MIN_SCALING = 1e-4
MAX_SCALING = 1e4
RHO_MIN = 1e-6
RHO_TOL = 1e-4
RHO_EQ_OVER_RHO_INEQ = 1e3


def _limit_scaling(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.minimum(v, MAX_SCALING)


class OSQP:
    def __init__(self):
        self.iters = 0
        self.status = "unsolved"

    # ------------------------------------------------------------------ setup
    def setup(self, P_diag, q, A, l, u, max_iter=4000, alpha=1.6, rho=0.1, sigma=1e-6, eps_abs=1e-3,
              eps_rel=1e-3, eps_prim_inf=1e-4, eps_dual_inf=1e-4, scaling=10, check_termination=25,
              warm_start=True, adaptive_rho=False):
        assert not adaptive_rho
        self.n, self.m = A.shape[1], A.shape[0]
        self.st = dict(max_iter=max_iter, alpha=alpha, rho=rho, sigma=sigma, eps_abs=eps_abs, eps_rel=eps_rel,
                       eps_prim_inf=eps_prim_inf, eps_dual_inf=eps_dual_inf, scaling=scaling,
                       check_termination=check_termination, warm_start=warm_start)
        A = sparse.csc_matrix(A)
        A.sort_indices()
        self.A_pattern = (A.indices.copy(), A.indptr.copy())
        # unscaled problem data
        self.P0 = np.asarray(P_diag, dtype=float).copy()
        self.q0 = np.asarray(q, dtype=float).copy()
        self.A0 = A.data.astype(float).copy()
        self.l0 = np.maximum(np.asarray(l, dtype=float), -OSQP_INFTY)
        self.u0 = np.minimum(np.asarray(u, dtype=float), OSQP_INFTY)
        self.x = np.zeros(self.n)
        self.z = np.zeros(self.m)
        self.y = np.zeros(self.m)
        self.constr_type = np.zeros(self.m, dtype=int)
        self._scale_data()
        self._update_rho_vec(force=True)
        self._factor()

    def _mat(self, data):
        return sparse.csc_matrix((data, self.A_pattern[0], self.A_pattern[1]), shape=(self.m, self.n))

    def _scale_data(self):
        """scale_data(): fresh Ruiz equilibration of (P0, q0, A0, l0, u0)."""
        n, m = self.n, self.m
        P, q = self.P0.copy(), self.q0.copy()
        A = self._mat(self.A0.copy())
        D, E, c = np.ones(n), np.ones(m), 1.0
        for _ in range(self.st["scaling"]):
            absA = abs(A)
            Dt = np.maximum(np.abs(P), absA.max(axis=0).toarray().ravel() if m else 0.0)
            Et = absA.max(axis=1).toarray().ravel()
            Dt = 1.0 / np.sqrt(_limit_scaling(Dt))
            Et = 1.0 / np.sqrt(_limit_scaling(Et))
            P = Dt * P * Dt
            A = sparse.diags(Et) @ A @ sparse.diags(Dt)
            A = sparse.csc_matrix(A)
            q = Dt * q
            D, E = D * Dt, E * Et
            c_temp = float(np.mean(np.abs(P)))          # P diagonal: column inf-norms are |P_ii|
            inf_norm_q = float(_limit_scaling(np.array([np.max(np.abs(q))]))[0])
            c_temp = float(_limit_scaling(np.array([max(c_temp, inf_norm_q)]))[0])
            c_temp = 1.0 / c_temp
            P, q = P * c_temp, q * c_temp
            c *= c_temp
        self.D, self.E, self.c = D, E, c
        self.P, self.q = P, q
        A.sort_indices()
        self.A = A
        self.l, self.u = E * self.l0, E * self.u0

    def _update_rho_vec(self, force=False):
        lo = self.l < -OSQP_INFTY * MIN_SCALING
        up = self.u > OSQP_INFTY * MIN_SCALING
        ct = np.where(lo & up, -1, np.where(self.u - self.l < RHO_TOL, 1, 0))
        changed = force or np.any(ct != self.constr_type)
        self.constr_type = ct
        rho = self.st["rho"]
        self.rho_vec = np.where(ct == -1, RHO_MIN, np.where(ct == 1, RHO_EQ_OVER_RHO_INEQ * rho, rho))
        return changed

    def _factor(self):
        n, m = self.n, self.m
        K = sparse.bmat([[sparse.diags(self.P + self.st["sigma"]), self.A.T],
                         [self.A, -sparse.diags(1.0 / self.rho_vec)]], format="csc")
        self.lu = splu(K)

    # ------------------------------------------------------------------ update (python wrapper order: q, bounds, Ax)
    def update(self, q=None, l=None, u=None, Ax=None):
        if q is not None:       # osqp_update_lin_cost
            self.q0 = np.asarray(q, dtype=float).copy()
            self.q = self.c * self.D * self.q0
        if l is not None and u is not None:   # osqp_update_bounds
            l = np.maximum(np.asarray(l, dtype=float), -OSQP_INFTY)
            u = np.minimum(np.asarray(u, dtype=float), OSQP_INFTY)
            if np.any(l > u):
                raise ValueError("lower bound must be lower than or equal to upper bound")
            self.l0, self.u0 = l.copy(), u.copy()
            self.l, self.u = self.E * l, self.E * u
            refactor = self._update_rho_vec()
        else:
            refactor = False
        if Ax is not None:      # osqp_update_A: unscale_data, overwrite, scale_data, refactor
            self.A0 = np.asarray(Ax, dtype=float).copy()
            self._scale_data()
            refactor = True
        if refactor:
            self._factor()

    # ------------------------------------------------------------------ solve
    def solve(self):
        st = self.st
        n, m = self.n, self.m
        alpha, sigma = st["alpha"], st["sigma"]
        if not st["warm_start"]:
            self.x[:], self.z[:], self.y[:] = 0, 0, 0
        x, z, y = self.x, self.z, self.y
        A, P, q, rho = self.A, self.P, self.q, self.rho_vec
        status = "unsolved"
        it = 0
        for it in range(1, st["max_iter"] + 1):
            x_prev, z_prev = x, z
            rhs = np.concatenate([sigma * x_prev - q, z_prev - y / rho])
            sol = self.lu.solve(rhs)
            xt, nu = sol[:n], sol[n:]
            zt = z_prev + (nu - y) / rho
            x = alpha * xt + (1 - alpha) * x_prev
            delta_x = x - x_prev
            z = np.clip(alpha * zt + (1 - alpha) * z_prev + y / rho, self.l, self.u)
            delta_y = rho * (alpha * zt + (1 - alpha) * z_prev - z)
            y = y + delta_y
            if st["check_termination"] and it % st["check_termination"] == 0:
                status = self._check_termination(x, z, y, delta_x, delta_y, approx=False)
                if status != "unsolved":
                    break
        if status == "unsolved":
            status = self._check_termination(x, z, y, delta_x, delta_y, approx=True)
            if status == "unsolved":
                status = "maximum iterations reached"
        self.iters, self.status = it, status
        if "infeasible" in status:
            # store_solution(): no solution -> NaN and cold start
            self.x, self.z, self.y = np.zeros(n), np.zeros(m), np.zeros(m)
            return np.full(n, np.nan)
        self.x, self.z, self.y = x, z, y
        return self.D * x

    def residuals(self, x, z, y):
        Ax = self.A @ x
        pri = np.max(np.abs((Ax - z) / self.E)) if self.m else 0.0
        Px, Aty = self.P * x, self.A.T @ y
        dua = np.max(np.abs((Px + self.q + Aty) / self.D)) / self.c
        return pri, dua, Ax, Px, Aty

    def _check_termination(self, x, z, y, delta_x, delta_y, approx):
        st = self.st
        k = 10.0 if approx else 1.0
        eps_abs, eps_rel = st["eps_abs"] * k, st["eps_rel"] * k
        eps_pinf, eps_dinf = st["eps_prim_inf"] * k, st["eps_dual_inf"] * k
        pri, dua, Ax, Px, Aty = self.residuals(x, z, y)
        self.pri_res, self.dua_res = pri, dua
        eps_pri = eps_abs + eps_rel * max(np.max(np.abs(z / self.E)), np.max(np.abs(Ax / self.E)))
        eps_dua = eps_abs + eps_rel * max(np.max(np.abs(self.q / self.D)), np.max(np.abs(Aty / self.D)),
                                          np.max(np.abs(Px / self.D))) / self.c
        prim_ok = pri < eps_pri
        dual_ok = dua < eps_dua
        prim_inf = (not prim_ok) and self._is_primal_infeasible(delta_y, eps_pinf)
        dual_inf = (not dual_ok) and self._is_dual_infeasible(delta_x, eps_dinf)
        if prim_ok and dual_ok:
            return "solved inaccurate" if approx else "solved"
        if prim_inf:
            return "primal infeasible inaccurate" if approx else "primal infeasible"
        if dual_inf:
            return "dual infeasible inaccurate" if approx else "dual infeasible"
        return "unsolved"

    def _is_primal_infeasible(self, delta_y, eps):
        up = self.u > OSQP_INFTY * MIN_SCALING
        lo = self.l < -OSQP_INFTY * MIN_SCALING
        dy = delta_y.copy()
        dy[up & lo] = 0.0
        only_up = up & ~lo
        dy[only_up] = np.minimum(dy[only_up], 0.0)
        only_lo = lo & ~up
        dy[only_lo] = np.maximum(dy[only_lo], 0.0)
        norm = np.max(np.abs(self.E * dy))
        if norm > eps:
            lhs = np.sum(self.u * np.maximum(dy, 0) + self.l * np.minimum(dy, 0))
            if lhs < -eps * norm:
                return np.max(np.abs((self.A.T @ dy) / self.D)) < eps * norm
        return False

    def _is_dual_infeasible(self, delta_x, eps):
        norm = np.max(np.abs(self.D * delta_x))
        c = self.c
        if norm > eps:
            if self.q @ delta_x < -c * eps * norm:
                if np.max(np.abs((self.P * delta_x) / self.D)) < c * eps * norm:
                    Adx = (self.A @ delta_x) / self.E
                    bad = ((self.u < OSQP_INFTY * MIN_SCALING) & (Adx > eps * norm)) | \
                          ((self.l > -OSQP_INFTY * MIN_SCALING) & (Adx < -eps * norm))
                    return not np.any(bad)
        return False

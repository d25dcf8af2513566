"""CasADi-external shim (libplm_casadi.so): the symbols casadi's external loader resolves, the sparsities it reads, and
(GPU) the values, called the way casadi calls a generated function: name(arg, res, iw, w, mem)."""
import ctypes
import os

import numpy as np
import pytest

from emu_util import make_robots

KIND, N = "whole_body_rnea", 6
NAMES = {"sqp_data": 5, "hess_data": 1, "f_data": 2, "g_data": 3}


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    from pino_locoman_b200.casadi_shim import SHIM_PATH, export_problem
    prod, _ = make_robots()
    path = str(tmp_path_factory.mktemp("plm") / "b2_rnea.plm")
    export_problem(prod["b2"], KIND, N, path, tau_nodes=3)
    os.environ["PLM_CASADI_PROBLEM"] = path
    lib = ctypes.CDLL(SHIM_PATH)
    for name in NAMES:
        for suffix in ("_n_in", "_n_out"):
            getattr(lib, name + suffix).restype = ctypes.c_longlong
        for suffix in ("_sparsity_in", "_sparsity_out"):
            getattr(lib, name + suffix).restype = ctypes.POINTER(ctypes.c_longlong)
            getattr(lib, name + suffix).argtypes = [ctypes.c_longlong]
        for suffix in ("_name_in", "_name_out"):
            getattr(lib, name + suffix).restype = ctypes.c_char_p
            getattr(lib, name + suffix).argtypes = [ctypes.c_longlong]
    return lib, prod["b2"]


def _sparsity(ptr):
    nrow, ncol = ptr[0], ptr[1]
    colind = [ptr[2 + c] for c in range(ncol + 1)]
    rows = [ptr[3 + ncol + e] for e in range(colind[-1])]
    return nrow, ncol, np.array(colind), np.array(rows)


def test_symbols_and_sparsity(shim):
    from pino_locoman_b200.handle import Handle
    lib, robot = shim
    h = Handle(robot, KIND, N, max_batch=0, tau_nodes=3)
    for name, nout in NAMES.items():
        assert getattr(lib, name + "_n_in")() == 2 and getattr(lib, name + "_n_out")() == nout
        for sym in ("", "_work", "_incref", "_decref"):
            assert hasattr(lib, name + sym)
        assert getattr(lib, name + "_name_in")(0) == b"x" and getattr(lib, name + "_name_in")(1) == b"p"
        assert _sparsity(getattr(lib, name + "_sparsity_in")(0))[:2] == (h.n, 1)
        assert _sparsity(getattr(lib, name + "_sparsity_in")(1))[:2] == (h.np, 1)
    assert [lib.sqp_data_name_out(i) for i in range(5)] == [b"grad_f", b"J_g", b"g", b"lbg", b"ubg"]
    nrow, ncol, colind, rows = _sparsity(lib.sqp_data_sparsity_out(1))
    assert (nrow, ncol) == (h.m, h.n) and colind[-1] == h.nnz
    pr, pc = h.jac_rows, h.jac_cols
    order = np.lexsort((pr, pc))                       # column major, rows ascending: compressed column storage
    assert np.array_equal(rows, pr[order])
    assert np.array_equal(np.diff(colind), np.bincount(pc, minlength=h.n))
    for c in range(ncol):                              # strictly ascending rows inside a column
        assert np.all(np.diff(rows[colind[c]:colind[c + 1]]) > 0)
    nrow, ncol, colind, rows = _sparsity(lib.hess_data_sparsity_out(0))
    assert (nrow, ncol) == (h.n, h.n) and np.array_equal(rows, np.arange(h.n)) and np.array_equal(colind, np.arange(h.n + 1))
    sz = [ctypes.c_longlong(-1) for _ in range(4)]
    assert lib.sqp_data_work(*[ctypes.byref(v) for v in sz]) == 0 and [v.value for v in sz] == [2, 5, 0, 0]


@pytest.mark.gpu
def test_values_match_the_library(shim):
    torch = pytest.importorskip("torch")
    from emu_util import random_problem
    from oracle.ocp import OracleOCP
    from pino_locoman_b200.handle import Handle
    lib, robot = shim
    _, ora = make_robots()
    o = OracleOCP(ora["b2"], KIND, N)
    x, p = random_problem(o, np.random.default_rng(2))
    h = Handle(robot, KIND, N, max_batch=1, tau_nodes=3)
    xd, pd = torch.tensor(x[None], device="cuda"), torch.tensor(p[None], device="cuda")
    grad, J, g, lbg, ubg = [t[0].cpu().numpy() for t in h.sqp_data(xd, pd)]
    pr, pc = h.jac_rows, h.jac_cols
    order = np.lexsort((pr, pc))
    dptr = ctypes.POINTER(ctypes.c_double)

    def call(name, nout, sizes):
        outs = [np.full(s, np.nan) for s in sizes]
        arg = (dptr * 2)(x.ctypes.data_as(dptr), p.ctypes.data_as(dptr))
        res = (dptr * nout)(*[o_.ctypes.data_as(dptr) for o_ in outs])
        assert getattr(lib, name)(arg, res, None, None, 0) == 0
        return outs

    got = call("sqp_data", 5, [h.n, h.nnz, h.m, h.m, h.m])
    assert np.array_equal(got[0], grad) and np.array_equal(got[1], J[order])
    assert np.array_equal(got[2], g) and np.array_equal(got[3], lbg) and np.array_equal(got[4], ubg)
    f, grad2 = [t[0].cpu().numpy() for t in h.f_data(xd, pd)]
    got = call("f_data", 2, [1, h.n])
    assert got[0][0] == f and np.array_equal(got[1], grad2)
    got = call("g_data", 3, [h.m, h.m, h.m])
    assert np.array_equal(got[0], g)
    got = call("hess_data", 1, [h.n])
    assert np.array_equal(got[0], h.hess_diag(pd)[0].cpu().numpy())
    # casadi passes NULL for outputs it does not need
    res = (dptr * 5)(None, None, got[0].ctypes.data_as(dptr), None, None)
    arg = (dptr * 2)(x.ctypes.data_as(dptr), p.ctypes.data_as(dptr))
    g_only = np.full(h.m, np.nan)
    res[2] = g_only.ctypes.data_as(dptr)
    assert lib.sqp_data(arg, res, None, None, 0) == 0 and np.array_equal(g_only, g)

"""Spatial algebra with pinocchio conventions (oracle; test infrastructure only).

Conventions restated from pinocchio [3P]:
motion = [linear; angular], force = [linear; angular], SE3 M=(R,p) maps child
coordinates into the parent (x -> R x + p), quaternion order [x, y, z, w].
All functions broadcast over leading batch dimensions and work for float64 and
complex128 (complex-step differentiation), so nothing here uses abs/conj.
"""
import numpy as np


def skew(v):
    v = np.asarray(v)
    z = np.zeros_like(v[..., 0])
    return np.stack([
        np.stack([z, -v[..., 2], v[..., 1]], -1),
        np.stack([v[..., 2], z, -v[..., 0]], -1),
        np.stack([-v[..., 1], v[..., 0], z], -1),
    ], -2)


def cross(a, b):
    a, b = np.broadcast_arrays(a, b)
    return np.stack([
        a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
        a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
        a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0],
    ], -1)


def mv(M, v):
    """Batched matrix @ vector."""
    return np.einsum("...ij,...j->...i", M, v)


def mtv(M, v):
    """Batched matrix.T @ vector."""
    return np.einsum("...ji,...j->...i", M, v)


def mm(A, B):
    return np.einsum("...ij,...jk->...ik", A, B)


def rpy_to_R(r, p, y):
    """URDF rpy -> R = Rz(y) Ry(p) Rx(r)."""
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def quat_to_R(q):
    """Unit quaternion [x,y,z,w] -> rotation matrix (analytic in the components)."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1),
    ], -2)


def quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz,
    ], -1)


def R_to_quat(R):
    """Real rotation matrix -> unit quaternion [x,y,z,w] with w >= 0 (non-batched use)."""
    R = np.asarray(R, dtype=float)
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax([R[0, 0], R[1, 1], R[2, 2]]))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


# ---------------------------------------------------------------------------
# SO(3)/SE(3) exponential and logarithm, series near zero (the first SQP iterate
# has DX = 0 exactly; pinocchio's casadi instantiation uses Taylor branches [3P]).
# ---------------------------------------------------------------------------
_SERIES_T2 = 0.25  # theta^2 below which the power series is used (theta < 0.5)


def _series(t2, first_k, terms=12):
    """sum_{i>=0} (-1)^i t2^i / (2i + first_k)!   (first_k=2: (1-cos)/t2, 3: (t-sin)/t^3, 1: sin/t)."""
    out = np.zeros_like(t2)
    fact = 1.0
    for n in range(1, first_k + 1):
        fact *= n
    term = np.ones_like(t2) / fact
    k = first_k
    for i in range(terms):
        out = out + term
        term = -term * t2 / ((k + 1) * (k + 2))
        k += 2
    return out


def exp_coeffs(w):
    """A = sin(t)/t, B = (1-cos t)/t^2, C = (t - sin t)/t^3 for t = |w|."""
    t2 = np.sum(w * w, -1)
    small = np.real(t2) < _SERIES_T2
    t2s = np.where(small, t2, 0.0)
    A_s, B_s, C_s = _series(t2s, 1), _series(t2s, 2), _series(t2s, 3)
    t2l = np.where(small, 1.0, t2)
    t = np.sqrt(t2l)
    A_l = np.sin(t) / t
    B_l = (1 - np.cos(t)) / t2l
    C_l = (t - np.sin(t)) / (t2l * t)
    return np.where(small, A_s, A_l), np.where(small, B_s, B_l), np.where(small, C_s, C_l)


def exp3(w):
    A, B, _ = exp_coeffs(w)
    W = skew(w)
    I = np.eye(3)
    return I + A[..., None, None] * W + B[..., None, None] * mm(W, W)


def exp3_quat(w):
    """Quaternion of exp3(w): [sin(t/2)/t * w, cos(t/2)]."""
    t2 = np.sum(w * w, -1)
    small = np.real(t2) < _SERIES_T2
    h2 = np.where(small, t2, 0.0) / 4
    s_s = 0.5 * _series(h2, 1)   # sin(t/2)/t = 0.5 * sin(h)/h
    c_s = 1 - h2 * _series(h2, 2)  # cos h = 1 - h^2 * (1-cos h)/h^2
    t = np.sqrt(np.where(small, 1.0, t2))
    s_l = np.sin(t / 2) / t
    c_l = np.cos(t / 2)
    s = np.where(small, s_s, s_l)
    c = np.where(small, c_s, c_l)
    return np.concatenate([s[..., None] * w, c[..., None]], -1)


def exp6(nu):
    """SE(3) exponential of nu=[rho; w]: returns (R, p) with p = V(w) rho."""
    rho, w = nu[..., :3], nu[..., 3:]
    A, B, C = exp_coeffs(w)
    W = skew(w)
    WW = mm(W, W)
    I = np.eye(3)
    R = I + A[..., None, None] * W + B[..., None, None] * WW
    V = I + B[..., None, None] * W + C[..., None, None] * WW
    return R, mv(V, rho)


def log3(R):
    """Real-valued SO(3) logarithm (used on parameters only, never differentiated)."""
    R = np.asarray(R, dtype=float)
    q = R_to_quat(R)
    n = np.linalg.norm(q[:3])
    if n < 1e-12:
        return 2.0 * q[:3] / q[3]
    ang = 2.0 * np.arctan2(n, q[3])
    return q[:3] / n * ang


def log6(R, p):
    w = log3(R)
    t2 = float(w @ w)
    W = skew(w)
    if t2 < 1e-6:
        beta = 1.0 / 12 + t2 / 720 + t2 * t2 / 30240
    else:
        t = np.sqrt(t2)
        beta = 1.0 / t2 - (1 + np.cos(t)) / (2 * t * np.sin(t))
    Vinv = np.eye(3) - 0.5 * W + beta * (W @ W)
    return np.concatenate([Vinv @ p, w])


# ---------------------------------------------------------------------------
# spatial transforms / cross products, [lin; ang]
# ---------------------------------------------------------------------------
def act_motion(R, p, m):
    v, w = m[..., :3], m[..., 3:]
    Rw = mv(R, w)
    return np.concatenate([mv(R, v) + cross(p, Rw), Rw], -1)


def actinv_motion(R, p, m):
    v, w = m[..., :3], m[..., 3:]
    return np.concatenate([mtv(R, v - cross(p, w)), mtv(R, w)], -1)


def act_force(R, p, f):
    l, n = f[..., :3], f[..., 3:]
    Rl = mv(R, l)
    return np.concatenate([Rl, mv(R, n) + cross(p, Rl)], -1)


def cross_mm(a, b):
    """motion x motion."""
    return np.concatenate([cross(a[..., 3:], b[..., :3]) + cross(a[..., :3], b[..., 3:]),
                           cross(a[..., 3:], b[..., 3:])], -1)


def cross_mf(a, f):
    """motion x* force."""
    return np.concatenate([cross(a[..., 3:], f[..., :3]),
                           cross(a[..., 3:], f[..., 3:]) + cross(a[..., :3], f[..., :3])], -1)


def inertia_mul(mass, com, Ic, m):
    """Y m for body inertia (mass, com, Ic about com) in its own frame."""
    v, w = m[..., :3], m[..., 3:]
    lin = mass * (v - cross(com, w))
    ang = mv(Ic, w) + cross(com, lin)
    return np.concatenate([lin, ang], -1)


def inertia_matrix(mass, com, Ic):
    C = skew(com)
    M = np.zeros((6, 6))
    M[:3, :3] = mass * np.eye(3)
    M[:3, 3:] = -mass * C
    M[3:, :3] = mass * C
    M[3:, 3:] = Ic - mass * C @ C
    return M


def motion_xform(R, p):
    """6x6 matrix of act_motion (child -> parent)."""
    z = np.zeros_like(R)
    top = np.concatenate([R, mm(skew(p), R)], -1)
    bot = np.concatenate([z, R], -1)
    return np.concatenate([top, bot], -2)


def force_xform(R, p):
    """6x6 matrix of act_force (child -> parent)."""
    z = np.zeros_like(R)
    top = np.concatenate([R, z], -1)
    bot = np.concatenate([mm(skew(p), R), R], -1)
    return np.concatenate([top, bot], -2)

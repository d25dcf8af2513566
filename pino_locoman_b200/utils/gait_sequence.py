"""Gait / contact schedules and the swing-foot z-velocity spline (host side, bit-exact).

Mirrors the reference's ``utils/gait_sequence.py``: ``GaitSequence`` (:5-77), ``get_spline_vel_z`` and
``CubicSpline`` (:96-133).  The schedule is built on the host in double precision with the reference's
operation order (time accumulated node by node, Python/IEEE ``%`` for the phase) and is generalised to a
batch of start times: ``t_current`` may be a scalar (reference behaviour, returns ``(4, N)`` arrays) or an
array of shape ``[B]`` (returns ``[B, 4, N]``).  The device only ever sees the resulting 0/1 contact flags and
swing phases inside the parameter vector ``p``.
"""
import numpy as np

# swing groups per gait: (upper bound of the gait phase, feet in swing); feet indexed FR, FL, RR, RL
_SWING_GROUPS = {
    "trot": ((0.5, (0, 3)), (None, (1, 2))),
    "walk": ((0.25, (1,)), (0.5, (2,)), (0.75, (0,)), (None, (3,))),
    "stand": (),
}
_N_CONTACTS = {"trot": 2, "walk": 3, "stand": 4}
_SWING_FRACTION = {"trot": 0.5, "walk": 0.25, "stand": 1.0}


class GaitSequence:
    def __init__(self, gait_type="trot", gait_period=0.5):
        if gait_type not in _SWING_GROUPS:
            raise ValueError(f"Gait: {gait_type} not supported")
        self.feet = ["FR_foot", "FL_foot", "RR_foot", "RL_foot"]
        self.gait_type = gait_type
        self.gait_period = gait_period
        self.n_contacts = _N_CONTACTS[gait_type]
        frac = _SWING_FRACTION[gait_type]
        self.swing_period = gait_period if frac == 1.0 else frac * gait_period

    def get_gait_schedule(self, t_current, dts, nodes):
        """Contact (0/1) and swing-phase (0..1) schedules over the horizon."""
        scalar = np.ndim(t_current) == 0
        t = np.atleast_1d(np.asarray(t_current, dtype=np.float64)).copy()
        B = t.shape[0]
        contact = np.ones((B, 4, nodes))
        swing = np.zeros((B, 4, nodes))
        groups = _SWING_GROUPS[self.gait_type]
        for i in range(nodes):
            if i > 0:
                t = t + dts[i - 1]
            if not groups:
                continue
            gait_phase = t % self.gait_period / self.gait_period
            swing_phase = t % self.swing_period / self.swing_period
            unassigned = np.ones(B, dtype=bool)
            for bound, feet in groups:
                sel = unassigned if bound is None else (unassigned & (gait_phase < bound))
                for f in feet:
                    contact[sel, f, i] = 0
                    swing[sel, f, i] = swing_phase[sel]
                unassigned = unassigned & ~sel
        if scalar:
            return contact[0], swing[0]
        return contact, swing


def horizon_dts(dt_min, dt_max, nodes):
    """Geometric step sizes dt_i = dt_min * gamma**i of optimization/ocp.py:71-74 (same expression order)."""
    ratio = dt_max / dt_min
    gamma = ratio ** (1 / (nodes - 1))
    return [dt_min * gamma ** i for i in range(nodes)]


class CubicSpline:
    """Cubic through (t0, pos0, vel0) and (t1, pos1, vel1), normalised time (gait_sequence.py:112-133)."""

    def __init__(self, t0, t1, pos0, vel0, pos1, vel1):
        self.t0, self.t1, self.dt = t0, t1, t1 - t0
        dpos, dvel = pos1 - pos0, vel1 - vel0
        self.c0 = pos0
        self.c1 = vel0 * self.dt
        self.c2 = -(3.0 * vel0 + dvel) * self.dt + 3.0 * dpos
        self.c3 = (2.0 * vel0 + dvel) * self.dt - 2.0 * dpos

    def position(self, t):
        tn = (t - self.t0) / self.dt
        return self.c3 * tn ** 3 + self.c2 * tn ** 2 + self.c1 * tn + self.c0

    def velocity(self, t):
        tn = (t - self.t0) / self.dt
        return (3.0 * self.c3 * tn ** 2 + 2.0 * self.c2 * tn + self.c1) / self.dt


def get_spline_vel_z(swing_phase, swing_period, h_max=0.1, v_liftoff=0.1, v_touchdown=-0.2):
    """Desired swing-foot z velocity: up-spline for phase < 0.5, down-spline after (gait_sequence.py:96-109)."""
    mid_time = swing_period / 2
    up = CubicSpline(0, mid_time, 0, v_liftoff, h_max, 0)
    down = CubicSpline(mid_time, swing_period, h_max, 0, 0, v_touchdown)
    t = swing_phase * swing_period
    return np.where(np.asarray(swing_phase) < 0.5, up.velocity(t), down.velocity(t))

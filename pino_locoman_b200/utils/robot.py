"""URDF/SRDF -> kinematic-tree tables for the GPU kernels.

Host-side replacement for the reference's ``utils/robot.py:10-118`` (``Robot``, ``Go2``, ``B2``, ``B2G``): instead of a
``pinocchio.Model`` the loader produces flat tables (parent index, joint placement, revolute axis, spatial
inertia, contact / end-effector frames, limits, reference pose) that ``plm_create`` uploads to HBM.  The tree is
built with pinocchio's URDF rules so that joint order, merged inertias and frame placements match the
reference model: free-flyer root joint, children visited depth-first in joint-name order, fixed joints (and
joints locked at the neutral configuration, ``buildReducedRobot``) folded into the parent body.
"""
import ctypes
import os
import xml.etree.ElementTree as ET

import numpy as np

from .gait_sequence import GaitSequence

_ROBOT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "robots")


def _vec(text, default):
    return np.array(default if text is None else [float(t) for t in text.split()], dtype=np.float64)


def _rpy(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def _hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


def _spatial_inertia(mass, com, I_com):
    """6x6 spatial inertia ([lin; ang]) about the frame origin."""
    C = _hat(com)
    Y = np.zeros((6, 6))
    Y[:3, :3] = mass * np.eye(3)
    Y[:3, 3:] = -mass * C
    Y[3:, :3] = mass * C
    Y[3:, 3:] = I_com - mass * C @ C
    return Y


def _motion_inverse_xform(T):
    """6x6 matrix taking motions of the parent frame to the child frame placed at T=(R,p) in the parent."""
    R, p = T[:3, :3], T[:3, 3]
    X = np.zeros((6, 6))
    X[:3, :3] = R.T
    X[:3, 3:] = -R.T @ _hat(p)
    X[3:, 3:] = R.T
    return X


class KinematicTree:
    """Flat tables of the movable-joint tree (body 0 = free-flyer root)."""

    def __init__(self, urdf_path, locked=()):
        xml = ET.parse(urdf_path).getroot()
        link_inertia = {}
        for link in xml.findall("link"):
            node = link.find("inertial")
            if node is None:
                link_inertia[link.get("name")] = None
                continue
            org = node.find("origin")
            com = _vec(None if org is None else org.get("xyz"), [0, 0, 0])
            Rin = _rpy(_vec(None if org is None else org.get("rpy"), [0, 0, 0]))
            i = node.find("inertia")
            I = np.array([[float(i.get("ixx")), float(i.get("ixy")), float(i.get("ixz"))],
                          [float(i.get("ixy")), float(i.get("iyy")), float(i.get("iyz"))],
                          [float(i.get("ixz")), float(i.get("iyz")), float(i.get("izz"))]])
            link_inertia[link.get("name")] = _spatial_inertia(float(node.find("mass").get("value")), com, Rin @ I @ Rin.T)
        joints = {}
        children_of = {name: [] for name in link_inertia}
        is_child = set()
        for j in xml.findall("joint"):
            org = j.find("origin")
            T = np.eye(4)
            T[:3, :3] = _rpy(_vec(None if org is None else org.get("rpy"), [0, 0, 0]))
            T[:3, 3] = _vec(None if org is None else org.get("xyz"), [0, 0, 0])
            ax = j.find("axis")
            joints[j.get("name")] = (j.get("type"), j.find("child").get("link"), T,
                                     _vec(None if ax is None else ax.get("xyz"), [1, 0, 0]))
            children_of[j.find("parent").get("link")].append(j.get("name"))
            is_child.add(j.find("child").get("link"))
        (root,) = [name for name in link_inertia if name not in is_child]

        self.joint_names = ["root_joint"]
        self.parent = [-1]
        self.placement = [np.eye(4)]
        self.axis = [np.zeros(3)]
        self._Y = [np.zeros((6, 6))]
        self.frames = {}     # name -> (body, 4x4 placement in the body's joint frame, kind)

        def attach(link, body, T):
            Y = link_inertia[link]
            if Y is not None:
                X = _motion_inverse_xform(T)
                self._Y[body] = self._Y[body] + X.T @ Y @ X
            self.frames.setdefault(link, (body, T.copy(), "BODY"))

        def descend(link, body, T):
            for jname in sorted(children_of[link]):   # urdfdom keeps joints in a name-sorted map
                jtype, child, Tj, axis = joints[jname]
                Tc = T @ Tj
                if jtype in ("revolute", "continuous") and jname not in locked:
                    self.joint_names.append(jname)
                    self.parent.append(body)
                    self.placement.append(Tc)
                    self.axis.append(axis / np.linalg.norm(axis))
                    self._Y.append(np.zeros((6, 6)))
                    new = len(self.parent) - 1
                    attach(child, new, np.eye(4))
                    descend(child, new, np.eye(4))
                elif jtype in ("fixed", "revolute", "continuous"):
                    self.frames.setdefault(jname, (body, Tc.copy(), "FIXED_JOINT"))
                    attach(child, body, Tc)
                    descend(child, body, Tc)
                else:
                    raise ValueError(f"unsupported URDF joint type '{jtype}' ({jname})")

        attach(root, 0, np.eye(4))
        descend(root, 0, np.eye(4))
        self.nbody = len(self.parent)
        self.nq, self.nv = 7 + self.nbody - 1, 6 + self.nbody - 1
        self.root_link = root
        # (mass, com, I_com) per body from the accumulated 6x6 inertias
        self.inertia = np.zeros((self.nbody, 10))
        for b, Y in enumerate(self._Y):
            m = Y[0, 0]
            c = np.array([Y[5, 1], Y[3, 2], Y[4, 0]]) / m     # m*[c]x block
            Ic = Y[3:, 3:] + m * _hat(c) @ _hat(c)
            self.inertia[b] = [m, *c, Ic[0, 0], Ic[0, 1], Ic[0, 2], Ic[1, 1], Ic[1, 2], Ic[2, 2]]
        self.total_mass = float(self.inertia[:, 0].sum())

    def frame(self, name):
        if name not in self.frames:
            raise KeyError(f"frame '{name}' not in model")
        return self.frames[name]

    def reference_configuration(self, srdf_path, pose):
        q = np.zeros(self.nq)
        q[6] = 1.0
        for gs in ET.parse(srdf_path).getroot().findall("group_state"):
            if gs.get("name") != pose:
                continue
            for j in gs.findall("joint"):
                vals = [float(t) for t in j.get("value").split()]
                if j.get("name") == "root_joint":
                    q[:7] = vals
                elif j.get("name") in self.joint_names:
                    q[7 + self.joint_names.index(j.get("name")) - 1] = vals[0]
            return q
        raise KeyError(f"reference pose '{pose}' not in {srdf_path}")


class Robot:
    """Mirror of the reference ``Robot`` (utils/robot.py:10-42): same attributes the OCP layer reads."""

    def __init__(self, urdf_path, srdf_path, reference_pose, lock_joints=None):
        self.tree = KinematicTree(urdf_path, tuple(lock_joints or ()))
        self.model = self.tree          # the tables play the role of pin.Model
        self.tree.robot = self          # Dynamics*(robot.model, mass, foot_frames) finds the limits / frames here
        if srdf_path and reference_pose:
            self.q0 = self.tree.reference_configuration(srdf_path, reference_pose)
        else:
            self.q0 = np.zeros(self.tree.nq)
            self.q0[6] = 1.0
        self.nq, self.nv = self.tree.nq, self.tree.nv
        self.nj = self.nq - 7
        self.nf = 12
        self.mass = self.tree.total_mass
        self.ext_force_frame = None
        self.arm_ee_frame = None
        self.gait_sequence = None
        self.foot_frames = None

    def set_gait_sequence(self, gait_type, gait_period):
        self.gait_sequence = GaitSequence(gait_type, gait_period)
        self.foot_frames = list(self.gait_sequence.feet)
        for f in self.foot_frames:
            self.tree.frame(f)

    # -- tables for the C ABI -------------------------------------------------------------------
    def tables(self):
        """Arrays in the layout of ``plm_robot_desc`` (include/pino_locoman_b200.h)."""
        t = self.tree
        feet = self.foot_frames or list(GaitSequence("trot", 0.8).feet)
        contacts = list(feet) + ([self.ext_force_frame] if self.ext_force_frame else [])
        placement = np.zeros((t.nbody, 12))
        for b in range(t.nbody):
            placement[b, :9] = t.placement[b][:3, :3].reshape(9)
            placement[b, 9:] = t.placement[b][:3, 3]
        out = dict(
            nbody=t.nbody,
            parent=np.array(t.parent, dtype=np.int32),
            placement=placement,
            axis=np.array(t.axis, dtype=np.float64),
            inertia=t.inertia.copy(),
            nfeet=len(feet),
            has_ext_force=int(self.ext_force_frame is not None),
            contact_body=np.array([t.frame(c)[0] for c in contacts], dtype=np.int32),
            contact_offset=np.array([t.frame(c)[1][:3, 3] for c in contacts], dtype=np.float64),
            arm_body=-1, arm_offset=np.zeros(3),
            joint_pos_min=np.asarray(self.joint_pos_min, dtype=np.float64),
            joint_pos_max=np.asarray(self.joint_pos_max, dtype=np.float64),
            joint_vel_max=np.asarray(self.joint_vel_max, dtype=np.float64),
            joint_torque_max=np.asarray(self.joint_torque_max, dtype=np.float64),
            q0=np.asarray(self.q0, dtype=np.float64),
        )
        if self.arm_ee_frame:
            body, T, _ = t.frame(self.arm_ee_frame)
            out["arm_body"], out["arm_offset"] = body, T[:3, 3].copy()
        return out


class Go2(Robot):
    def __init__(self, reference_pose="standing"):
        super().__init__(os.path.join(_ROBOT_DIR, "go2.urdf"), os.path.join(_ROBOT_DIR, "go2.srdf"), reference_pose)
        # joint limits tiled hip, thigh, calf (utils/robot.py:52-55)
        self.joint_pos_min = np.tile([-1.0472, -1.5708, -2.7227], 4)
        self.joint_pos_max = np.tile([1.0472, 3.4907, -0.83776], 4)
        self.joint_vel_max = np.tile([30.1, 30.1, 15.70], 4)
        self.joint_torque_max = np.tile([23.7, 23.7, 45.43], 4)


class B2(Robot):
    def __init__(self, reference_pose="standing", payload=None):
        super().__init__(os.path.join(_ROBOT_DIR, "b2.urdf"), os.path.join(_ROBOT_DIR, "b2.srdf"), reference_pose)
        self.joint_pos_min = np.tile([-0.87, -0.94, -2.82], 4)     # utils/robot.py:65-68
        self.joint_pos_max = np.tile([0.87, 4.69, -0.43], 4)
        self.joint_vel_max = np.tile([23.0, 23.0, 14.0], 4)
        self.joint_torque_max = np.tile([200.0, 200.0, 320.0], 4)
        if payload in ("front", "rear"):                           # utils/robot.py:70-76
            self.ext_force_frame = f"payload_joint_{payload}"
            self.tree.frame(self.ext_force_frame)
            self.nf += 3


class B2G(Robot):
    def __init__(self, reference_pose="standing_with_arm_up", ignore_arm=False):
        arm = ["joint1", "joint2", "joint3", "joint4", "joint5", "joint6", "jointGripper"]
        lock = arm if ignore_arm else ["jointGripper"]             # utils/robot.py:83-86 (ids 14..20 / 20)
        super().__init__(os.path.join(_ROBOT_DIR, "b2g.urdf"), os.path.join(_ROBOT_DIR, "b2g.srdf"), reference_pose,
                         lock_joints=lock)
        self.joint_pos_min = np.tile([-0.87, -0.94, -2.82], 4)
        self.joint_pos_max = np.tile([0.87, 4.69, -0.43], 4)
        self.joint_vel_max = np.tile([23.0, 23.0, 14.0], 4)
        self.joint_torque_max = np.tile([200.0, 200.0, 320.0], 4)
        if not ignore_arm:                                         # utils/robot.py:96-118
            self.ext_force_frame = "gripperStator"
            self.arm_ee_frame = "gripperStator"
            self.nf += 3
            self.joint_pos_min = np.concatenate((self.joint_pos_min, [-2.62, 0.0, -2.88, -1.52, -1.34, -2.79]))
            self.joint_pos_max = np.concatenate((self.joint_pos_max, [2.62, 2.97, 0.0, 1.52, 1.34, 2.79]))
            self.joint_vel_max = np.concatenate((self.joint_vel_max, [3.14] * 6))
            self.joint_torque_max = np.concatenate((self.joint_torque_max, [30.0, 60.0, 30.0, 30.0, 30.0, 30.0]))


class RobotDesc(ctypes.Structure):
    """ctypes image of ``plm_robot_desc``."""
    _fields_ = [
        ("nbody", ctypes.c_int32), ("parent", ctypes.POINTER(ctypes.c_int32)),
        ("placement", ctypes.POINTER(ctypes.c_double)), ("axis", ctypes.POINTER(ctypes.c_double)),
        ("inertia", ctypes.POINTER(ctypes.c_double)), ("nfeet", ctypes.c_int32), ("has_ext_force", ctypes.c_int32),
        ("contact_body", ctypes.POINTER(ctypes.c_int32)), ("contact_offset", ctypes.POINTER(ctypes.c_double)),
        ("arm_body", ctypes.c_int32), ("arm_offset", ctypes.c_double * 3),
        ("joint_pos_min", ctypes.POINTER(ctypes.c_double)), ("joint_pos_max", ctypes.POINTER(ctypes.c_double)),
        ("joint_vel_max", ctypes.POINTER(ctypes.c_double)), ("joint_torque_max", ctypes.POINTER(ctypes.c_double)),
        ("q0", ctypes.POINTER(ctypes.c_double)),
    ]


def robot_desc(robot):
    """Build a ``plm_robot_desc`` (and keep the numpy buffers alive on the returned object)."""
    tab = robot.tables()
    d = RobotDesc()
    keep = {}

    def dptr(name):
        keep[name] = np.ascontiguousarray(tab[name], dtype=np.float64)
        return keep[name].ctypes.data_as(ctypes.POINTER(ctypes.c_double))

    def iptr(name):
        keep[name] = np.ascontiguousarray(tab[name], dtype=np.int32)
        return keep[name].ctypes.data_as(ctypes.POINTER(ctypes.c_int32))

    d.nbody = tab["nbody"]
    d.parent = iptr("parent")
    d.placement, d.axis, d.inertia = dptr("placement"), dptr("axis"), dptr("inertia")
    d.nfeet, d.has_ext_force = tab["nfeet"], tab["has_ext_force"]
    d.contact_body, d.contact_offset = iptr("contact_body"), dptr("contact_offset")
    d.arm_body = tab["arm_body"]
    d.arm_offset = (ctypes.c_double * 3)(*tab["arm_offset"])
    d.joint_pos_min, d.joint_pos_max = dptr("joint_pos_min"), dptr("joint_pos_max")
    d.joint_vel_max, d.joint_torque_max = dptr("joint_vel_max"), dptr("joint_torque_max")
    d.q0 = dptr("q0")
    d._keep = keep
    return d

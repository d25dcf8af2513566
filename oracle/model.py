"""URDF/SRDF -> kinematic tree with pinocchio's model-building rules (oracle; test infrastructure only).

Follows the reference's utils/robot.py:10-118 (RobotWrapper.BuildFromURDF with a
JointModelFreeFlyer root, optional buildReducedRobot(lock_joints) at the neutral
configuration, SRDF reference pose) with the pinocchio/urdfdom semantics of
SURVEY.md section 2.4 [3P]: children visited depth-first in joint-name order,
fixed-joint children merged into the nearest movable ancestor, one BODY frame per
link and one FIXED_JOINT frame per fixed joint.
"""
import os
import xml.etree.ElementTree as ET

import numpy as np

from .spatial import rpy_to_R

ROBOT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pino_locoman_b200", "robots")


class Frame:
    def __init__(self, name, parent, R, p, kind):
        self.name, self.parent, self.R, self.p, self.kind = name, parent, R, p, kind


class Model:
    """joints[0] is the universe, joints[1] the free-flyer root."""

    def __init__(self):
        self.names = ["universe", "root_joint"]
        self.parents = [0, 0]
        self.placement_R = [np.eye(3), np.eye(3)]
        self.placement_p = [np.zeros(3), np.zeros(3)]
        self.axis = [None, None]              # revolute axis in the joint frame
        self.idx_q = [0, 0]
        self.idx_v = [0, 0]
        # body inertia accumulated as (mass, first moment m*c, inertia about the joint origin)
        self._m = [0.0, 0.0]
        self._mc = [np.zeros(3), np.zeros(3)]
        self._Io = [np.zeros((3, 3)), np.zeros((3, 3))]
        self.frames = [Frame("universe", 0, np.eye(3), np.zeros(3), "FIXED_JOINT"),
                       Frame("root_joint", 1, np.eye(3), np.zeros(3), "JOINT")]
        self.nq, self.nv = 7, 6
        self.gravity = np.array([0.0, 0.0, -9.81])

    # --- construction -----------------------------------------------------
    def add_revolute(self, name, parent, R, p, axis):
        self.names.append(name)
        self.parents.append(parent)
        self.placement_R.append(R)
        self.placement_p.append(p)
        self.axis.append(axis / np.linalg.norm(axis))
        self.idx_q.append(self.nq)
        self.idx_v.append(self.nv)
        self._m.append(0.0)
        self._mc.append(np.zeros(3))
        self._Io.append(np.zeros((3, 3)))
        self.nq += 1
        self.nv += 1
        return len(self.names) - 1

    def append_body(self, joint, R, p, mass, com, Ic):
        """Add a link inertia (mass, com, Ic in the link frame) placed at (R,p) in the joint frame."""
        c = R @ com + p
        I = R @ Ic @ R.T
        self._m[joint] += mass
        self._mc[joint] = self._mc[joint] + mass * c
        self._Io[joint] = self._Io[joint] + I + mass * ((c @ c) * np.eye(3) - np.outer(c, c))

    def finalize(self):
        n = len(self.names)
        self.njoints = n
        self.mass = np.array(self._m)
        self.com = np.zeros((n, 3))
        self.Ic = np.zeros((n, 3, 3))
        for i in range(1, n):
            m = self._m[i]
            c = self._mc[i] / m
            self.com[i] = c
            self.Ic[i] = self._Io[i] - m * ((c @ c) * np.eye(3) - np.outer(c, c))
        self.total_mass = float(sum(self._m[1:]))
        self.children = [[] for _ in range(n)]
        for i in range(2, n):
            self.children[self.parents[i]].append(i)
        self.subtree = [[i] for i in range(n)]
        for i in range(n - 1, 1, -1):
            self.subtree[self.parents[i]] += self.subtree[i]
        self.nj = self.nq - 7

    def getFrameId(self, name, kind=None):
        for i, f in enumerate(self.frames):
            if f.name == name and (kind is None or f.kind == kind):
                return i
        return len(self.frames)

    def neutral(self):
        q = np.zeros(self.nq)
        q[6] = 1.0
        return q


def _floats(s, default):
    if s is None:
        return np.array(default, dtype=float)
    return np.array([float(t) for t in s.split()], dtype=float)


def build_model(urdf_path, lock_joint_names=()):
    root = ET.parse(urdf_path).getroot()
    links = {}
    for l in root.findall("link"):
        inertial = l.find("inertial")
        if inertial is None:
            links[l.get("name")] = None
            continue
        org = inertial.find("origin")
        xyz = _floats(org.get("xyz") if org is not None else None, [0, 0, 0])
        rpy = _floats(org.get("rpy") if org is not None else None, [0, 0, 0])
        mass = float(inertial.find("mass").get("value"))
        ie = inertial.find("inertia")
        ixx, ixy, ixz, iyy, iyz, izz = (float(ie.get(k)) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz"))
        I = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
        links[l.get("name")] = (mass, xyz, rpy_to_R(*rpy), I)

    joints = {}
    child_links = set()
    for j in root.findall("joint"):
        org = j.find("origin")
        xyz = _floats(org.get("xyz") if org is not None else None, [0, 0, 0])
        rpy = _floats(org.get("rpy") if org is not None else None, [0, 0, 0])
        ax = j.find("axis")
        axis = _floats(ax.get("xyz") if ax is not None else None, [1, 0, 0])
        joints[j.get("name")] = dict(type=j.get("type"), parent=j.find("parent").get("link"),
                                     child=j.find("child").get("link"), p=xyz, R=rpy_to_R(*rpy), axis=axis)
        child_links.add(j.find("child").get("link"))
    roots = [n for n in links if n not in child_links]
    assert len(roots) == 1, roots
    root_link = roots[0]

    # urdfdom: joints_ is a std::map (sorted by name); children appended in that order
    children = {n: [] for n in links}
    for jn in sorted(joints):
        children[joints[jn]["parent"]].append(jn)

    model = Model()

    def add_link_body(link_name, joint_id, R, p):
        inert = links[link_name]
        if inert is not None:
            mass, com, Ri, I = inert
            # inertial frame (com, Ri) expressed in the link frame
            model.append_body(joint_id, R, p, mass, com, Ri @ I @ Ri.T)
        model.frames.append(Frame(link_name, joint_id, R, p, "BODY"))

    def visit(link_name, joint_id, R, p):
        """(R,p): placement of this link's frame in joint_id's frame."""
        for jn in children[link_name]:
            jd = joints[jn]
            Rj, pj = R @ jd["R"], R @ jd["p"] + p
            movable = jd["type"] in ("revolute", "continuous") and jn not in lock_joint_names
            if movable:
                jid = model.add_revolute(jn, joint_id, Rj, pj, jd["axis"])
                model.frames.append(Frame(jn, jid, np.eye(3), np.zeros(3), "JOINT"))
                add_link_body(jd["child"], jid, np.eye(3), np.zeros(3))
                visit(jd["child"], jid, np.eye(3), np.zeros(3))
            elif jd["type"] in ("fixed", "revolute", "continuous"):
                # fixed joint, or joint locked at the neutral configuration (angle 0)
                model.frames.append(Frame(jn, joint_id, Rj, pj, "FIXED_JOINT"))
                add_link_body(jd["child"], joint_id, Rj, pj)
                visit(jd["child"], joint_id, Rj, pj)
            else:
                raise ValueError(f"unsupported joint type {jd['type']}")

    add_link_body(root_link, 1, np.eye(3), np.zeros(3))
    visit(root_link, 1, np.eye(3), np.zeros(3))
    model.finalize()
    return model


def load_reference_configuration(model, srdf_path, pose):
    root = ET.parse(srdf_path).getroot()
    q = model.neutral()
    for gs in root.findall("group_state"):
        if gs.get("name") != pose:
            continue
        for j in gs.findall("joint"):
            vals = [float(t) for t in j.get("value").split()]
            name = j.get("name")
            if name == "root_joint":
                q[:7] = vals
            elif name in model.names:
                q[model.idx_q[model.names.index(name)]] = vals[0]
        return q
    raise KeyError(pose)


class OracleRobot:
    """Mirror of utils/robot.py Robot/Go2/B2/B2G (limits at robot.py:52-55, 65-68, 91-118)."""

    FEET = ["FR_foot", "FL_foot", "RR_foot", "RL_foot"]  # gait_sequence.py:7

    def __init__(self, name, reference_pose=None, payload=None, ignore_arm=False, robot_dir=ROBOT_DIR):
        self.name = name
        lock = ()
        if name == "b2g":
            lock = ("joint1", "joint2", "joint3", "joint4", "joint5", "joint6", "jointGripper") if ignore_arm \
                else ("jointGripper",)
        self.model = build_model(os.path.join(robot_dir, f"{name}.urdf"), lock)
        if reference_pose is None:
            reference_pose = "standing_with_arm_up" if name == "b2g" else "standing"
        self.q0 = load_reference_configuration(self.model, os.path.join(robot_dir, f"{name}.srdf"), reference_pose)
        m = self.model
        self.nq, self.nv, self.nj, self.nf = m.nq, m.nv, m.nq - 7, 12
        self.mass = m.total_mass
        self.ext_force_frame = None
        self.arm_ee_frame = None
        if name == "go2":
            self.joint_pos_min = np.tile([-1.0472, -1.5708, -2.7227], 4)
            self.joint_pos_max = np.tile([1.0472, 3.4907, -0.83776], 4)
            self.joint_vel_max = np.tile([30.1, 30.1, 15.70], 4)
            self.joint_torque_max = np.tile([23.7, 23.7, 45.43], 4)
        else:
            self.joint_pos_min = np.tile([-0.87, -0.94, -2.82], 4)
            self.joint_pos_max = np.tile([0.87, 4.69, -0.43], 4)
            self.joint_vel_max = np.tile([23.0, 23.0, 14.0], 4)
            self.joint_torque_max = np.tile([200.0, 200.0, 320.0], 4)
        if name == "b2" and payload in ("front", "rear"):
            self.ext_force_frame = m.getFrameId(f"payload_joint_{payload}", "FIXED_JOINT")
            self.nf += 3
        if name == "b2g" and not ignore_arm:
            self.ext_force_frame = m.getFrameId("gripperStator", "FIXED_JOINT")
            self.arm_ee_frame = self.ext_force_frame
            self.nf += 3
            self.joint_pos_min = np.concatenate((self.joint_pos_min, [-2.62, 0.0, -2.88, -1.52, -1.34, -2.79]))
            self.joint_pos_max = np.concatenate((self.joint_pos_max, [2.62, 2.97, 0.0, 1.52, 1.34, 2.79]))
            self.joint_vel_max = np.concatenate((self.joint_vel_max, [3.14] * 6))
            self.joint_torque_max = np.concatenate((self.joint_torque_max, [30.0, 60, 30, 30, 30, 30]))
        self.foot_frames = [m.getFrameId(f) for f in self.FEET]
        self.base_frame = m.getFrameId("base_link")  # dynamics/dynamics.py:17

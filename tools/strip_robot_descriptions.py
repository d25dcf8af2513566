#!/usr/bin/env python3
"""Derive kinematics+inertia-only robot descriptions from full URDF/SRDF packages.

The hot path only needs the kinematic tree (joints, origins, axes), the link
inertials and the SRDF reference poses (utils/robot.py:14-30 in the reference).
Visual/collision/mesh/transmission/gazebo content is dropped, numbers are kept
as the exact source strings so the derived model is bit-identical.

Usage: tools/strip_robot_descriptions.py <robots_dir> <out_dir>
"""
import sys
import os
import xml.etree.ElementTree as ET


def _attrs(el, names):
    return " ".join(f'{k}="{" ".join(el.get(k).split())}"' for k in names if el is not None and el.get(k) is not None)


def strip_urdf(src, dst):
    root = ET.parse(src).getroot()
    out = [f'<robot name="{root.get("name")}">']
    for link in root.findall("link"):
        inertial = link.find("inertial")
        if inertial is None:
            out.append(f'  <link name="{link.get("name")}"/>')
            continue
        out.append(f'  <link name="{link.get("name")}">')
        out.append("    <inertial>")
        org = inertial.find("origin")
        if org is not None:
            out.append(f"      <origin {_attrs(org, ['xyz', 'rpy'])}/>")
        out.append(f"      <mass {_attrs(inertial.find('mass'), ['value'])}/>")
        out.append(f"      <inertia {_attrs(inertial.find('inertia'), ['ixx', 'ixy', 'ixz', 'iyy', 'iyz', 'izz'])}/>")
        out.append("    </inertial>")
        out.append("  </link>")
    for joint in root.findall("joint"):
        out.append(f'  <joint name="{joint.get("name")}" type="{joint.get("type")}">')
        org = joint.find("origin")
        if org is not None:
            out.append(f"    <origin {_attrs(org, ['xyz', 'rpy'])}/>")
        out.append(f'    <parent link="{joint.find("parent").get("link")}"/>')
        out.append(f'    <child link="{joint.find("child").get("link")}"/>')
        ax = joint.find("axis")
        if ax is not None:
            out.append(f"    <axis {_attrs(ax, ['xyz'])}/>")
        lim = joint.find("limit")
        if lim is not None:
            out.append(f"    <limit {_attrs(lim, ['lower', 'upper', 'effort', 'velocity'])}/>")
        out.append("  </joint>")
    out.append("</robot>")
    with open(dst, "w") as f:
        f.write("\n".join(out) + "\n")


def strip_srdf(src, dst):
    root = ET.parse(src).getroot()
    out = [f'<robot name="{root.get("name")}">']
    for gs in root.findall("group_state"):
        out.append(f'  <group_state name="{gs.get("name")}" group="{gs.get("group")}">')
        for j in gs.findall("joint"):
            out.append(f'    <joint name="{j.get("name")}" value="{" ".join(j.get("value").split())}"/>')
        out.append("  </group_state>")
    out.append("</robot>")
    with open(dst, "w") as f:
        f.write("\n".join(out) + "\n")


if __name__ == "__main__":
    src_dir, out_dir = sys.argv[1], sys.argv[2]
    os.makedirs(out_dir, exist_ok=True)
    for name in ("go2", "b2", "b2g"):
        strip_urdf(os.path.join(src_dir, f"{name}_description/urdf/{name}.urdf"), os.path.join(out_dir, f"{name}.urdf"))
        strip_srdf(os.path.join(src_dir, f"{name}_description/srdf/{name}.srdf"), os.path.join(out_dir, f"{name}.srdf"))
        print("wrote", name)

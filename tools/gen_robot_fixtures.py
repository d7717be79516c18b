#!/usr/bin/env python3
"""Generate the robot description fixtures under smpl_b200/data/.

Runs ONLY in the authoring container (reads /root/reference); the generated
`.robot` text files are committed and are what the oracle, the product and the
tests read.  Nothing at test/bench time touches /root/reference.

What comes from the reference (data fixtures):
  * PR2 sphere models, voxel-link list and collision groups:
      sbpl_collision_checking_test/config/collision_model_pr2.yaml
  * UBR1 sphere model (old-format yaml, re-grouped per link):
      sbpl_collision_checking_test/config/ubr1_model.yaml
  * PR2 allowed collision matrix (1081 setEntry lines):
      smpl_test/src/call_planner.cpp:441-1527

What is OUR fixture (the reference ships no URDF; pr2_description and
ubr1_description are external ROS packages): joint origins / axes / limits of
the PR2 and UBR1 kinematic trees, written down from the public robot
descriptions, and coarse boxes standing in for the link meshes that the
reference voxelises (robot_collision_model.cpp:565-575).

File format (one record per line, '#' comments):
  robot NAME
  root LINK
  world_joint NAME TYPE
  joint NAME TYPE PARENT CHILD x y z roll pitch yaw ax ay az has_limits lower upper has_safety soft_lower soft_upper
  spheres_model LINK
  sphere NAME x y z radius priority
  voxels_model LINK res cx cy cz sx sy sz       (box centre/size in link frame; 0 size = no geometry)
  group NAME
  group_link NAME
  group_chain BASE TIP
  group_sub NAME
  acm A B allowed(0/1)
Child joints of a link are listed to the builder in joint-NAME order, which is
the order urdfdom's initTree() produces (it walks a std::map keyed by joint
name) and therefore the order RobotCollisionModel::initRobotModel sees
(robot_collision_model.cpp:160-197).
"""
import os
import re
import sys

import yaml

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "smpl_b200", "data")


def J(name, typ, parent, child, xyz=(0, 0, 0), rpy=(0, 0, 0), axis=(0, 0, 0),
      limits=None, safety=None):
    return dict(name=name, type=typ, parent=parent, child=child, xyz=xyz, rpy=rpy,
                axis=axis, limits=limits, safety=safety)


def pr2_arm(side, y):
    """One PR2 arm + gripper (public pr2_description values)."""
    s = side
    lim = {
        "shoulder_pan": ((-2.2853981634, 0.714601836603), (-2.1353981634, 0.564601836603)) if s == "r"
        else ((-0.714601836603, 2.2853981634), (-0.564601836603, 2.1353981634)),
        "shoulder_lift": ((-0.5236, 1.3963), (-0.3536, 1.2963)),
        "upper_arm_roll": ((-3.9, 0.8), (-3.75, 0.65)) if s == "r" else ((-0.8, 3.9), (-0.65, 3.75)),
        "elbow_flex": ((-2.3213, 0.0), (-2.1213, -0.15)),
        "wrist_flex": ((-2.18, 0.0), (-2.0, -0.1)),
    }
    j = []
    j.append(J(f"{s}_shoulder_pan_joint", "revolute", "torso_lift_link", f"{s}_shoulder_pan_link",
               (0.0, y, 0.0), axis=(0, 0, 1), limits=lim["shoulder_pan"][0], safety=lim["shoulder_pan"][1]))
    j.append(J(f"{s}_shoulder_lift_joint", "revolute", f"{s}_shoulder_pan_link", f"{s}_shoulder_lift_link",
               (0.1, 0.0, 0.0), axis=(0, 1, 0), limits=lim["shoulder_lift"][0], safety=lim["shoulder_lift"][1]))
    j.append(J(f"{s}_upper_arm_roll_joint", "revolute", f"{s}_shoulder_lift_link", f"{s}_upper_arm_roll_link",
               axis=(1, 0, 0), limits=lim["upper_arm_roll"][0], safety=lim["upper_arm_roll"][1]))
    j.append(J(f"{s}_upper_arm_joint", "fixed", f"{s}_upper_arm_roll_link", f"{s}_upper_arm_link"))
    j.append(J(f"{s}_elbow_flex_joint", "revolute", f"{s}_upper_arm_link", f"{s}_elbow_flex_link",
               (0.4, 0.0, 0.0), axis=(0, 1, 0), limits=lim["elbow_flex"][0], safety=lim["elbow_flex"][1]))
    j.append(J(f"{s}_forearm_roll_joint", "continuous", f"{s}_elbow_flex_link", f"{s}_forearm_roll_link",
               axis=(1, 0, 0)))
    j.append(J(f"{s}_forearm_joint", "fixed", f"{s}_forearm_roll_link", f"{s}_forearm_link"))
    j.append(J(f"{s}_forearm_cam_frame_joint", "fixed", f"{s}_forearm_roll_link", f"{s}_forearm_cam_frame",
               (0.135, 0.0, 0.044), rpy=(-1.5708 if s == "r" else 1.5708, -0.562868683369, 0.0)))
    j.append(J(f"{s}_wrist_flex_joint", "revolute", f"{s}_forearm_link", f"{s}_wrist_flex_link",
               (0.321, 0.0, 0.0), axis=(0, 1, 0), limits=lim["wrist_flex"][0], safety=lim["wrist_flex"][1]))
    j.append(J(f"{s}_wrist_roll_joint", "continuous", f"{s}_wrist_flex_link", f"{s}_wrist_roll_link",
               axis=(1, 0, 0)))
    j.append(J(f"{s}_gripper_palm_joint", "fixed", f"{s}_wrist_roll_link", f"{s}_gripper_palm_link"))
    fl = ((0.0, 0.548), (0.0, 0.548))
    j.append(J(f"{s}_gripper_l_finger_joint", "revolute", f"{s}_gripper_palm_link", f"{s}_gripper_l_finger_link",
               (0.07691, 0.01, 0.0), axis=(0, 0, 1), limits=fl[0]))
    j.append(J(f"{s}_gripper_l_finger_tip_joint", "revolute", f"{s}_gripper_l_finger_link",
               f"{s}_gripper_l_finger_tip_link", (0.09137, 0.00495, 0.0), axis=(0, 0, -1), limits=fl[0]))
    j.append(J(f"{s}_gripper_r_finger_joint", "revolute", f"{s}_gripper_palm_link", f"{s}_gripper_r_finger_link",
               (0.07691, -0.01, 0.0), axis=(0, 0, -1), limits=fl[0]))
    j.append(J(f"{s}_gripper_r_finger_tip_joint", "revolute", f"{s}_gripper_r_finger_link",
               f"{s}_gripper_r_finger_tip_link", (0.09137, -0.00495, 0.0), axis=(0, 0, 1), limits=fl[0]))
    j.append(J(f"{s}_gripper_tool_joint", "fixed", f"{s}_gripper_palm_link", f"{s}_gripper_tool_frame",
               (0.18, 0.0, 0.0)))
    return j


def pr2_joints():
    j = [
        J("base_footprint_joint", "fixed", "base_footprint", "base_link", (0.0, 0.0, 0.051)),
        J("base_bellow_joint", "fixed", "base_link", "base_bellow_link", (-0.29, 0.0, 0.8)),
        J("torso_lift_joint", "prismatic", "base_link", "torso_lift_link", (-0.05, 0.0, 0.739675),
          axis=(0, 0, 1), limits=(0.0, 0.33), safety=(0.0115, 0.325)),
        J("head_pan_joint", "revolute", "torso_lift_link", "head_pan_link", (-0.01707, 0.0, 0.38145),
          axis=(0, 0, 1), limits=(-3.007, 3.007), safety=(-2.857, 2.857)),
        J("head_tilt_joint", "revolute", "head_pan_link", "head_tilt_link", (0.068, 0.0, 0.0),
          axis=(0, 1, 0), limits=(-0.4712, 1.39626), safety=(-0.3712, 1.29626)),
        J("laser_tilt_mount_joint", "revolute", "torso_lift_link", "laser_tilt_mount_link",
          (0.09893, 0.0, 0.227), axis=(0, 1, 0), limits=(-0.7854, 1.48353), safety=(-0.7354, 1.43353)),
        J("sensor_mount_frame_joint", "fixed", "head_tilt_link", "sensor_mount_link", (0.0232, 0.0, 0.0645)),
    ]
    # four casters, two wheels each
    for cn, (cx, cy) in {"fl": (0.2246, 0.2246), "fr": (0.2246, -0.2246),
                         "bl": (-0.2246, 0.2246), "br": (-0.2246, -0.2246)}.items():
        j.append(J(f"{cn}_caster_rotation_joint", "continuous", "base_link", f"{cn}_caster_rotation_link",
                   (cx, cy, 0.0282), axis=(0, 0, 1)))
        j.append(J(f"{cn}_caster_l_wheel_joint", "continuous", f"{cn}_caster_rotation_link",
                   f"{cn}_caster_l_wheel_link", (0.0, 0.049, 0.0), axis=(0, 1, 0)))
        j.append(J(f"{cn}_caster_r_wheel_joint", "continuous", f"{cn}_caster_rotation_link",
                   f"{cn}_caster_r_wheel_link", (0.0, -0.049, 0.0), axis=(0, 1, 0)))
    j += pr2_arm("r", -0.188)
    j += pr2_arm("l", 0.188)
    return j


# Coarse boxes in the link frame (centre, size) standing in for the PR2 meshes
# that the reference voxelises for out-of-group links.  OUR fixture.
PR2_LINK_BOXES = {
    "base_link": ((0.0, 0.0, 0.14), (0.65, 0.65, 0.28)),
    "base_bellow_link": ((0.0, 0.0, -0.22), (0.05, 0.37, 0.3)),
    "torso_lift_link": ((-0.25, 0.0, -0.1), (0.25, 0.5, 0.8)),
    "head_pan_link": ((0.0, 0.0, 0.05), (0.2, 0.25, 0.1)),
    "head_tilt_link": ((0.05, 0.0, 0.08), (0.2, 0.3, 0.16)),
    "laser_tilt_mount_link": ((0.0, 0.0, 0.0), (0.08, 0.1, 0.08)),
    "shoulder_pan_link": ((0.05, 0.0, -0.1), (0.25, 0.2, 0.4)),
    "shoulder_lift_link": ((0.0, 0.0, 0.0), (0.15, 0.15, 0.15)),
    "upper_arm_roll_link": ((0.1, 0.0, 0.0), (0.2, 0.12, 0.12)),
    "upper_arm_link": ((0.3, 0.0, -0.02), (0.3, 0.14, 0.14)),
    "elbow_flex_link": ((0.0, 0.0, 0.0), (0.12, 0.14, 0.12)),
    "forearm_roll_link": ((0.1, 0.0, 0.0), (0.1, 0.1, 0.1)),
    "forearm_link": ((0.2, 0.0, 0.0), (0.3, 0.1, 0.1)),
    "wrist_flex_link": ((0.0, 0.0, 0.0), (0.08, 0.08, 0.08)),
    "wrist_roll_link": ((0.02, 0.0, 0.0), (0.04, 0.06, 0.06)),
    "gripper_palm_link": ((0.07, 0.0, 0.0), (0.1, 0.1, 0.05)),
    "gripper_l_finger_link": ((0.045, 0.01, 0.0), (0.09, 0.03, 0.03)),
    "gripper_r_finger_link": ((0.045, -0.01, 0.0), (0.09, 0.03, 0.03)),
    "gripper_l_finger_tip_link": ((0.02, 0.0, 0.0), (0.04, 0.02, 0.02)),
    "gripper_r_finger_tip_link": ((0.02, 0.0, 0.0), (0.04, 0.02, 0.02)),
    "caster_rotation_link": ((0.0, 0.0, 0.03), (0.15, 0.1, 0.1)),
    "caster_l_wheel_link": ((0.0, 0.0, 0.0), (0.15, 0.03, 0.15)),
    "caster_r_wheel_link": ((0.0, 0.0, 0.0), (0.15, 0.03, 0.15)),
}


def box_for(link):
    if link in PR2_LINK_BOXES:
        return PR2_LINK_BOXES[link]
    m = re.match(r"^(?:[lr]|fl|fr|bl|br)_(.*)$", link)
    if m and m.group(1) in PR2_LINK_BOXES:
        return PR2_LINK_BOXES[m.group(1)]
    return ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0))


def fmt(v):
    return repr(float(v))


def emit_joint(f, j):
    lim = j["limits"]
    saf = j["safety"]
    vals = [j["name"], j["type"], j["parent"], j["child"]]
    vals += [fmt(v) for v in j["xyz"]] + [fmt(v) for v in j["rpy"]] + [fmt(v) for v in j["axis"]]
    vals += ["1", fmt(lim[0]), fmt(lim[1])] if lim else ["0", "0.0", "0.0"]
    vals += ["1", fmt(saf[0]), fmt(saf[1])] if saf else ["0", "0.0", "0.0"]
    f.write("joint " + " ".join(vals) + "\n")


def gen_pr2():
    cfg = yaml.safe_load(open(f"{REF}/sbpl_collision_checking_test/config/collision_model_pr2.yaml"))
    rcm = cfg["robot_collision_model"]
    joints = pr2_joints()
    links = {"base_footprint"} | {j["child"] for j in joints}
    with open(os.path.join(OUT, "pr2.robot"), "w") as f:
        f.write("# generated by tools/gen_robot_fixtures.py -- do not edit\n")
        f.write("robot pr2\nroot base_footprint\n")
        f.write(f"world_joint {rcm['world_joint']['name']} {rcm['world_joint']['type']}\n")
        for j in joints:
            emit_joint(f, j)
        for sm in rcm["spheres_models"]:
            assert sm["link_name"] in links, sm["link_name"]
            assert not sm.get("auto", False)
            f.write(f"spheres_model {sm['link_name']}\n")
            for s in sm["spheres"]:
                f.write("sphere %s %s %s %s %s %d\n" % (
                    s["name"], fmt(s["x"]), fmt(s["y"]), fmt(s["z"]), fmt(s["radius"]), s["priority"]))
        for vm in rcm["voxels_models"]:
            ln = vm["link_name"]
            assert ln in links, ln
            c, sz = box_for(ln)
            f.write("voxels_model %s %s %s %s\n" % (
                ln, fmt(vm["res"]), " ".join(fmt(v) for v in c), " ".join(fmt(v) for v in sz)))
        for g in rcm["collision_groups"]:
            f.write(f"group {g['name']}\n")
            for l in g.get("links") or []:
                assert l["name"] in links, l["name"]
                f.write(f"group_link {l['name']}\n")
            for c in g.get("chains") or []:
                f.write(f"group_chain {c['base']} {c['tip']}\n")
            for sg in g.get("groups") or []:
                f.write(f"group_sub {sg}\n")
        # ACM: call_planner.cpp:441-1527
        src = open(f"{REF}/smpl_test/src/call_planner.cpp").read()
        body = src[src.index("void initAllowedCollisionsPR2"):src.index("int main(")]
        n = 0
        for a, b, v in re.findall(r'acm\.setEntry\("([^"]+)",\s*"([^"]+)",\s*(true|false)\)', body):
            f.write(f"acm {a} {b} {1 if v == 'true' else 0}\n")
            n += 1
        assert n == 1081, n
    print("pr2.robot: %d joints, %d acm entries" % (len(joints), n))


def gen_ubr1():
    cfg = yaml.safe_load(open(f"{REF}/sbpl_collision_checking_test/config/ubr1_model.yaml"))
    spheres = {s["name"]: s for s in cfg["collision_spheres"]}
    arm = [g for g in cfg["collision_groups"] if g["name"] == "arm"][0]
    body = [g for g in cfg["collision_groups"] if g["name"] == "body"][0]
    # OUR fixture: UBR1 kinematics (public ubr1_description values)
    joints = [
        J("torso_lift_joint", "prismatic", "base_link", "torso_lift_link", (-0.086875, 0.0, 0.37743),
          axis=(0, 0, 1), limits=(0.0, 0.35)),
        J("head_pan_joint", "revolute", "torso_lift_link", "head_pan_link", (0.053125, 0.0, 0.603001),
          axis=(0, 0, 1), limits=(-1.57, 1.57)),
        J("head_tilt_joint", "revolute", "head_pan_link", "head_tilt_link", (0.14253, 0.0, 0.057999),
          axis=(0, 1, 0), limits=(-0.76, 1.45)),
        J("head_camera_joint", "fixed", "head_tilt_link", "head_camera_link", (0.055, 0.0, 0.0225)),
        J("shoulder_pan_joint", "revolute", "torso_lift_link", "shoulder_pan_link", (0.119525, 0.0, 0.34858),
          axis=(0, 0, 1), limits=(-1.6056, 1.6056)),
        J("shoulder_lift_joint", "revolute", "shoulder_pan_link", "shoulder_lift_link", (0.117, 0.0, 0.06),
          axis=(0, 1, 0), limits=(-1.221, 1.518)),
        J("upperarm_roll_joint", "continuous", "shoulder_lift_link", "upperarm_roll_link", (0.219, 0.0, 0.0),
          axis=(1, 0, 0)),
        J("elbow_flex_joint", "revolute", "upperarm_roll_link", "elbow_flex_link", (0.133, 0.0, 0.0),
          axis=(0, 1, 0), limits=(-2.251, 2.251)),
        J("forearm_roll_joint", "continuous", "elbow_flex_link", "forearm_roll_link", (0.197, 0.0, 0.0),
          axis=(1, 0, 0)),
        J("wrist_flex_joint", "revolute", "forearm_roll_link", "wrist_flex_link", (0.1245, 0.0, 0.0),
          axis=(0, 1, 0), limits=(-2.16, 2.16)),
        J("wrist_roll_joint", "continuous", "wrist_flex_link", "wrist_roll_link", (0.1385, 0.0, 0.0),
          axis=(1, 0, 0)),
        J("gripper_joint", "fixed", "wrist_roll_link", "gripper_link", (0.16645, 0.0, 0.0)),
        J("left_gripper_finger_joint", "prismatic", "gripper_link", "left_gripper_finger_link",
          (0.0, 0.015425, 0.0), rpy=(-1.5707963267948966, 0.0, 0.0), axis=(0, 0, 1), limits=(0.0, 0.05)),
        J("right_gripper_finger_joint", "prismatic", "gripper_link", "right_gripper_finger_link",
          (0.0, -0.015425, 0.0), rpy=(1.5707963267948966, 0.0, 0.0), axis=(0, 0, 1), limits=(0.0, 0.05)),
    ]
    links = {"base_link"} | {j["child"] for j in joints}
    boxes = {
        "base_link": ((0.0, 0.0, 0.18), (0.56, 0.56, 0.36)),
        "torso_lift_link": ((-0.05, 0.0, 0.3), (0.3, 0.35, 0.6)),
        "head_pan_link": ((0.05, 0.0, 0.03), (0.25, 0.25, 0.06)),
        "head_tilt_link": ((0.02, 0.0, 0.03), (0.12, 0.26, 0.12)),
        "head_camera_link": ((0.0, 0.0, 0.0), (0.05, 0.2, 0.05)),
    }
    with open(os.path.join(OUT, "ubr1.robot"), "w") as f:
        f.write("# generated by tools/gen_robot_fixtures.py -- do not edit\n")
        f.write("robot ubr1\nroot base_link\nworld_joint world_joint fixed\n")
        for j in joints:
            emit_joint(f, j)
        group_links = []
        for cl in arm["collision_links"]:
            ln = cl["root"]
            assert ln in links, ln
            group_links.append(ln)
            f.write(f"spheres_model {ln}\n")
            for sn in cl["spheres"].split():
                s = spheres[sn]
                f.write("sphere %s %s %s %s %s %d\n" % (
                    s["name"], fmt(s["x"]), fmt(s["y"]), fmt(s["z"]), fmt(s["radius"]), s["priority"]))
        for cl in body["collision_links"]:
            ln = cl["root"]
            c, sz = boxes[ln]
            f.write("voxels_model %s 0.01 %s %s\n" % (
                ln, " ".join(fmt(v) for v in c), " ".join(fmt(v) for v in sz)))
        # group "arm": chain torso_lift_link..wrist_roll_link (ubr1_model.yaml root_name/tip_name) + fingers
        f.write("group arm\n")
        f.write("group_chain shoulder_pan_link wrist_roll_link\n")
        for ln in ("gripper_link", "left_gripper_finger_link", "right_gripper_finger_link"):
            f.write(f"group_link {ln}\n")
    print("ubr1.robot: %d joints" % len(joints))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (authoring container only)")
    gen_pr2()
    gen_ubr1()

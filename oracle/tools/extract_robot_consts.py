#!/usr/bin/env python3
"""Robot constants extractor (test infrastructure; run once per robot, in the build container).

Restates, without pinocchio, what the reference obtains from pinocchio at construction:
  * total mass                      -- mpc/models/model.cpp:27  (pinocchio::computeTotalMass)
  * composite rotational inertia about the whole-body CoM at the nominal configuration,
    expressed in the floating-base frame
                                    -- mpc/models/single_rigid_body_model.cpp:33-37
                                       (computeCentroidalMap; oMi[1].actInv(oYcrb[0]).inertia())
  * root->hip joint translations    -- mpc/models/single_rigid_body_model.cpp:258-308 (GetCOMToHip),
    including the +0.025 x / +-0.1 y offsets applied there.

pinocchio is not installed here, so these values are "parity unpinned" against pinocchio itself;
they follow the URDF arithmetic (fixed joints merged into their parent, revolute joints rotated by
the nominal joint angle about their axis).  Sibling order FL, FR, RL, RR is pinocchio's alphabetical
child ordering, which is what apps/a1_configuration.yaml:init_config assumes.

Usage: extract_robot_consts.py <urdf> <out.json>
"""
import json
import sys
import xml.etree.ElementTree as ET

import numpy as np


def inverse3_cofactor(M):
    """Ir_.inverse() as Eigen evaluates it for a fixed 3 x 3 matrix (Eigen/src/LU/InverseImpl.h, compute_inverse<.., 3>): cofactors
    times 1 / det, the determinant expanded along the first column.  Plain Python floats, same operation order."""
    m = [[float(M[i][j]) for j in range(3)] for i in range(3)]

    def cof(i, j):
        i1, i2, j1, j2 = (i + 1) % 3, (i + 2) % 3, (j + 1) % 3, (j + 2) % 3
        return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1]
    c0, c1, c2 = cof(0, 0), cof(1, 0), cof(2, 0)
    det = (c0 * m[0][0] + c1 * m[1][0]) + c2 * m[2][0]
    invdet = 1.0 / det
    r = [[0.0] * 3 for _ in range(3)]
    r[0][0], r[0][1], r[0][2] = c0 * invdet, c1 * invdet, c2 * invdet
    r[1][0], r[1][1], r[2][2] = cof(0, 1) * invdet, cof(1, 1) * invdet, cof(2, 2) * invdet
    r[1][2], r[2][1], r[2][0] = cof(2, 1) * invdet, cof(1, 2) * invdet, cof(0, 2) * invdet
    return r


def rpy_to_R(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def axis_angle_R(axis, q):
    a = np.asarray(axis, float)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(q) * K + (1 - np.cos(q)) * (K @ K)


def parse_origin(el):
    if el is None:
        return np.zeros(3), np.eye(3)
    xyz = np.array([float(v) for v in el.get("xyz", "0 0 0").split()])
    rpy = [float(v) for v in el.get("rpy", "0 0 0").split()]
    return xyz, rpy_to_R(*rpy)


def main(urdf_path, out_path, joint_cfg):
    root = ET.parse(urdf_path).getroot()
    links = {}
    for l in root.findall("link"):
        inertial = l.find("inertial")
        if inertial is None:
            links[l.get("name")] = None
            continue
        c, Rc = parse_origin(inertial.find("origin"))
        m = float(inertial.find("mass").get("value"))
        i = inertial.find("inertia")
        I = np.array([[float(i.get("ixx")), float(i.get("ixy")), float(i.get("ixz"))],
                      [float(i.get("ixy")), float(i.get("iyy")), float(i.get("iyz"))],
                      [float(i.get("ixz")), float(i.get("iyz")), float(i.get("izz"))]])
        links[l.get("name")] = (m, c, Rc @ I @ Rc.T)
    children = {}
    child_names = set()
    for j in root.findall("joint"):
        if j.find("parent") is None:
            continue  # transmission <joint> stubs
        p = j.find("parent").get("link")
        c = j.find("child").get("link")
        xyz, R = parse_origin(j.find("origin"))
        ax = j.find("axis")
        axis = [float(v) for v in ax.get("xyz").split()] if ax is not None else [1, 0, 0]
        children.setdefault(p, []).append((j.get("name"), j.get("type"), c, xyz, R, axis))
        child_names.add(c)
    roots = [n for n in links if n not in child_names]
    assert len(roots) == 1, roots
    bodies = []      # (mass, com in base frame, inertia about own com in base frame)
    joint_pos = {}   # joint name -> translation in base frame

    def walk(link, p_w, R_w):
        if links[link] is not None:
            m, c, I = links[link]
            bodies.append((m, p_w + R_w @ c, R_w @ I @ R_w.T))
        for (jn, jt, c, xyz, R, axis) in sorted(children.get(link, []), key=lambda t: t[0]):
            pj = p_w + R_w @ xyz
            Rj = R_w @ R
            if jt in ("revolute", "continuous"):
                Rj = Rj @ axis_angle_R(axis, joint_cfg[jn])
            joint_pos[jn] = pj
            walk(c, pj, Rj)

    walk(roots[0], np.zeros(3), np.eye(3))
    mass = sum(b[0] for b in bodies)
    com = sum(b[0] * b[1] for b in bodies) / mass
    Ir = np.zeros((3, 3))
    for m, c, I in bodies:
        d = c - com
        Ir += I + m * (d @ d * np.eye(3) - np.outer(d, d))
    hips, hips_raw = [], []
    for name in ("FL_hip_joint", "FR_hip_joint", "RL_hip_joint", "RR_hip_joint"):
        t = joint_pos[name].copy()                      # root joint sits at the base origin
        hips_raw.append([float(v) for v in t])
        t[1] += 0.1 if t[1] >= 0 else -0.1              # single_rigid_body_model.cpp:291-297
        t[0] += 0.025                                   # :299-305 (both branches add 0.025)
        hips.append([float(t[0]), float(t[1])])
    # leg chains for the inverse kinematics (single_rigid_body_model.cpp:314-455): hip / thigh / calf joint and foot frame, each
    # placed in the frame of the movable joint before it (fixed joints in between merged, as pinocchio's URDF parser does)
    by_child = {}
    for plink, lst in children.items():
        for (jn, jt, c, xyz, R, axis) in lst:
            by_child[c] = (jn, jt, plink, xyz, R, axis)
    by_name = {v[0]: (c,) + v for c, v in by_child.items()}

    def placement(jname, ancestor_link):
        c, jn, jt, plink, xyz, R, axis = by_name[jname]
        t, Rt = xyz.copy(), R.copy()
        link = plink
        while link != ancestor_link:
            jn2, jt2, plink2, xyz2, R2, _ = by_child[link]
            assert jt2 == "fixed", (jname, jn2)
            t, Rt = xyz2 + R2 @ t, R2 @ Rt
            link = plink2
        return t, Rt, axis, c

    legs = []
    for leg in ("FL", "FR", "RL", "RR"):
        anc = roots[0]
        ts, Rs, axes = [], [], []
        for jn in (f"{leg}_hip_joint", f"{leg}_thigh_joint", f"{leg}_calf_joint", f"{leg}_foot_fixed"):
            t, Rt, axis, child = placement(jn, anc)
            ts.append([float(v) for v in t])
            Rs.append([float(v) for v in Rt.reshape(-1)])
            if not jn.endswith("_fixed"):
                a = np.asarray(axis, float)
                axes.append([float(v) for v in a / np.linalg.norm(a)])
            anc = child
        legs.append({"t": ts, "R": Rs, "axis": axes})
    out = {
        "robot": "a1",
        "source": "models/a1_description/urdf/a1.urdf + apps/a1_configuration.yaml:init_config",
        "mass": float(mass),
        "com_in_base": [float(v) for v in com],
        "Ir": [[float(v) for v in row] for row in Ir],
        "Ir_inv": [[float(v) for v in row] for row in inverse3_cofactor(Ir)],
        "hip_offsets_xy": hips,
        "hip_joint_translation": hips_raw,   # oMi[hip joint] - oMi[root] before the reference's own offsets
        "gravity": [0.0, 0.0, -9.81],
        "legs": legs,   # FL, FR, RL, RR: placements of hip / thigh / calf joint and of the foot frame; joint axes
    }
    with open(out_path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    # apps/a1_configuration.yaml:init_config joint part, pinocchio order FL, FR, RL, RR
    legs = {"FL": (-0.02, 0.9, -1.6), "FR": (0.02, 0.9, -1.6), "RL": (0.02, 0.9, -1.6), "RR": (-0.02, 0.9, -1.6)}
    cfg = {}
    for leg, (h, t, c) in legs.items():
        cfg[f"{leg}_hip_joint"] = h
        cfg[f"{leg}_thigh_joint"] = t
        cfg[f"{leg}_calf_joint"] = c
    main(sys.argv[1], sys.argv[2], cfg)

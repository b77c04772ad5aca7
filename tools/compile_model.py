#!/usr/bin/env python
"""Offline model compiler: reference MJCF -> constant tables for the oracle and the CUDA kernel.

Reads  packages/biped_assets/biped_assets/models/h12/scene/h12_12dof.xml  (reference, read-only)
Writes include/h1v2_model_h12.h        (C tables, float64 literals; shared DATA, not code)
       h1v2_isaac_b200/model/h12_12dof.json  (same numbers for Python)

What it does:
  * applies the <default> joint/geom attributes (h12_12dof.xml:4-7), including to the free joint,
    which is declared with <joint type="free"> and therefore inherits damping/armature/frictionloss;
  * fuses the 43 joint-less bodies under torso_link (h12_12dof.xml:143-343) into the pelvis composite;
  * emits the 13 moving bodies (root + 2 x 6 leg links) with parent, offset, axis, mass, COM, inertia;
  * computes MuJoCo's qpos0 constants that enter the constraint regulariser
    (dof_invweight0, body_invweight0, meaninertia) with an independent dense numpy CRBA.

Nothing under tests/-m gpu, smoke() or bench.py reads /root/reference: the generated files are committed.
"""
from __future__ import annotations

import argparse
import json
import os
import xml.etree.ElementTree as ET

import numpy as np

REF_XML = "/root/reference/packages/biped_assets/biped_assets/models/h12/scene/h12_12dof.xml"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

LEG_BODIES = ["hip_yaw_link", "hip_pitch_link", "hip_roll_link", "knee_link", "ankle_pitch_link", "ankle_roll_link"]


def quat_to_mat(q):
    w, x, y, z = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
        ]
    )


def fvec(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    v = np.array([float(t) for t in s.split()], dtype=np.float64)
    if n is not None:
        assert len(v) == n, (s, n)
    return v


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


class Body:
    def __init__(self, name, parent, pos, quat, mass, ipos, inertia_b, joint):
        self.name, self.parent, self.pos, self.quat = name, parent, pos, quat
        self.mass, self.ipos, self.inertia_b, self.joint = mass, ipos, inertia_b, joint
        self.geoms = []
        self.children = []


def parse(xml_path):
    root = ET.parse(xml_path).getroot()
    dflt = root.find("default")
    jd = dict(dflt.find("joint").attrib)
    gd = dict(dflt.find("geom").attrib)
    bodies = []

    def walk(elem, parent):
        inert = elem.find("inertial")
        q = fvec(inert.get("quat"), 4, [1, 0, 0, 0])
        Rq = quat_to_mat(q / np.linalg.norm(q))
        Ib = Rq @ np.diag(fvec(inert.get("diaginertia"), 3)) @ Rq.T
        j = elem.find("joint")
        joint = None
        if j is not None:
            a = dict(jd)
            a.update(j.attrib)
            joint = a
        b = Body(
            elem.get("name"),
            parent,
            fvec(elem.get("pos"), 3, [0, 0, 0]),
            fvec(elem.get("quat"), 4, [1, 0, 0, 0]),
            float(inert.get("mass")),
            fvec(inert.get("pos"), 3),
            Ib,
            joint,
        )
        for g in elem.findall("geom"):
            a = dict(gd)
            a.update(g.attrib)
            if a.get("contype", "1") == "0" and a.get("conaffinity", "1") == "0":
                continue  # visual only
            b.geoms.append(a)
        bodies.append(b)
        idx = len(bodies) - 1
        if parent is not None:
            bodies[parent].children.append(idx)
        for c in elem.findall("body"):
            walk(c, idx)

    for e in root.find("worldbody").findall("body"):
        walk(e, None)
    key = fvec(root.find("keyframe").find("key").get("qpos"))
    return bodies, jd, gd, key


def fuse_static(bodies, idx):
    """Composite (mass, com, inertia about com) of body idx plus all joint-less descendants, in idx's frame."""
    parts = []  # (mass, com in frame, inertia about own com in frame)

    def rec(i, R, p):
        b = bodies[i]
        parts.append((b.mass, p + R @ b.ipos, R @ b.inertia_b @ R.T))
        for c in b.children:
            cb = bodies[c]
            if cb.joint is None:
                Rc = R @ quat_to_mat(cb.quat / np.linalg.norm(cb.quat))
                rec(c, Rc, p + R @ cb.pos)

    rec(idx, np.eye(3), np.zeros(3))
    m = sum(x[0] for x in parts)
    com = sum(x[0] * x[1] for x in parts) / m
    I = np.zeros((3, 3))
    for mi, ci, Ii in parts:
        d = ci - com
        I += Ii + mi * (d @ d * np.eye(3) - np.outer(d, d))
    return m, com, I, len(parts)


def build_model(xml_path=REF_XML):
    bodies, jd, gd, key = parse(xml_path)
    names = [b.name for b in bodies]
    pelvis = names.index("pelvis")
    order = [pelvis] + [names.index(f"{s}_{n}") for s in ("left", "right") for n in LEG_BODIES]
    m0, c0, I0, nfused = fuse_static(bodies, pelvis)

    mb = []
    for k, bi in enumerate(order):
        b = bodies[bi]
        if k == 0:
            mass, ipos, Ib, parent = m0, c0, I0, -1
        else:
            mass, ipos, Ib = b.mass, b.ipos, b.inertia_b
            parent = order.index(b.parent)
            assert np.allclose(b.quat, [1, 0, 0, 0])
        j = b.joint
        mb.append(
            dict(
                name=b.name,
                parent=parent,
                pos=b.pos.tolist(),
                mass=mass,
                ipos=np.asarray(ipos).tolist(),
                inertia=np.asarray(Ib).tolist(),
                axis=(fvec(j.get("axis"), 3).tolist() if j.get("type", "hinge") != "free" else [0, 0, 0]),
                range=(fvec(j.get("range"), 2).tolist() if j.get("range") else [0, 0]),
                frcrange=(fvec(j.get("actuatorfrcrange"), 2).tolist() if j.get("actuatorfrcrange") else [0, 0]),
                damping=float(j.get("damping", 0)),
                armature=float(j.get("armature", 0)),
                frictionloss=float(j.get("frictionloss", 0)),
            )
        )
    total_mass = sum(b.mass for b in bodies)
    assert abs(total_mass - sum(x["mass"] for x in mb)) < 1e-9

    # ---- original-body COMs needed for body_invweight0 (MuJoCo keeps un-fused bodies) ----
    torso = names.index("torso_link")
    tb = bodies[torso]
    assert tb.parent == pelvis and np.allclose(tb.pos, 0) and np.allclose(tb.quat, [1, 0, 0, 0])
    pelvis_ipos_own = bodies[pelvis].ipos
    torso_ipos_own = tb.pos + tb.ipos

    model = dict(
        source="packages/biped_assets/biped_assets/models/h12/scene/h12_12dof.xml",
        nbody=13,
        nv=18,
        nq=19,
        total_mass=total_mass,
        n_fused_into_root=nfused,
        bodies=mb,
        key_qpos=key.tolist(),
        qpos0=[0, 0, float(bodies[pelvis].pos[2]), 1, 0, 0, 0] + [0.0] * 12,
        geom_solref=fvec(gd.get("solref"), 2).tolist(),
        pelvis_ipos_own=pelvis_ipos_own.tolist(),
        torso_ipos_own=torso_ipos_own.tolist(),
    )
    # colliders (SURVEY 8(a) a5 + Appendix A): sole corners from the STL sole polygon == URDF rectangle
    model["colliders"] = collider_table()
    model.update(qpos0_constants(model))
    return model


def collider_table():
    """Point/sphere-vs-plane contact candidates.  body index is into the 13 moving bodies.

    feet  : 4 sole corners of ankle_roll_link (STL sole plane z=-0.045; h12_12dof.urdf:168-191 rectangle)
    shins : knee_link cylinder r=0.02 half-length 0.1 at z=-0.2 (h12_12dof.xml:90, urdf:116-121) as 2 end spheres
    torso : box half-extents (0.04,0.08,0.05) at z=0.15 (h12_12dof.xml:146, urdf:387-392) as 8 corners
    pelvis: sphere r=0.05 (h12_12dof.xml:70)
    Each row: body, x, y, z, radius, tracked-body slot (0,1 feet L/R; 2,3 shins L/R; 4 torso; 5 pelvis).
    """
    rows = []
    for side, foot, shin in ((0, 6, 4), (1, 12, 10)):
        for x, y in ((-0.081, 0.038), (-0.081, -0.038), (0.169, 0.021), (0.169, -0.021)):
            rows.append([foot, x, y, -0.045, 0.0, side])
        for z in (-0.1, -0.3):
            rows.append([shin, 0.0, 0.0, z, 0.02, 2 + side])
    for sx in (-1, 1):
        for sy in (-1, 1):
            for sz in (-1, 1):
                rows.append([0, 0.04 * sx, 0.08 * sy, 0.15 + 0.05 * sz, 0.0, 4])
    rows.append([0, 0.0, 0.0, 0.0, 0.05, 5])
    return rows


# ----------------------------------------------------------------------------------------------
# independent dense dynamics at qpos0 (numpy, world frame about the world origin)
# ----------------------------------------------------------------------------------------------
def fk(model, qpos):
    nb = model["nbody"]
    R = [None] * nb
    x = [None] * nb
    p = np.array(qpos[0:3])
    q = np.array(qpos[3:7])
    R[0] = quat_to_mat(q / np.linalg.norm(q))
    x[0] = p
    for i in range(1, nb):
        b = model["bodies"][i]
        pa = b["parent"]
        a = np.array(b["axis"])
        th = qpos[7 + i - 1]
        K = skew(a)
        Rj = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
        R[i] = R[pa] @ Rj
        x[i] = x[pa] + R[pa] @ np.array(b["pos"])
    return R, x


def dof_columns(model, R, x):
    """6x18 motion-subspace columns [ang; lin at world origin] and the body each dof belongs to."""
    S = np.zeros((6, 18))
    dof_body = [0] * 6 + list(range(1, 13))
    for k in range(3):
        S[3 + k, k] = 1.0
        w = R[0][:, k]
        S[0:3, 3 + k] = w
        S[3:6, 3 + k] = np.cross(x[0], w)
    for i in range(1, 13):
        w = R[i] @ np.array(model["bodies"][i]["axis"])
        S[0:3, 5 + i] = w
        S[3:6, 5 + i] = np.cross(x[i], w)
    return S, dof_body


def ancestors(model, b):
    out = []
    while b >= 0:
        out.append(b)
        b = model["bodies"][b]["parent"]
    return out


def mass_matrix(model, qpos):
    R, x = fk(model, qpos)
    S, dof_body = dof_columns(model, R, x)
    M = np.zeros((18, 18))
    for bi, b in enumerate(model["bodies"]):
        c = x[bi] + R[bi] @ np.array(b["ipos"])
        Iw = R[bi] @ np.array(b["inertia"]) @ R[bi].T
        m = b["mass"]
        I6 = np.zeros((6, 6))
        I6[0:3, 0:3] = Iw + m * (c @ c * np.eye(3) - np.outer(c, c))
        I6[0:3, 3:6] = m * skew(c)
        I6[3:6, 0:3] = m * skew(c).T
        I6[3:6, 3:6] = m * np.eye(3)
        anc = set(ancestors(model, bi))
        J = S * np.array([1.0 if dof_body[d] in anc else 0.0 for d in range(18)])
        M += J.T @ I6 @ J
    arm = np.array([model["bodies"][0]["armature"]] * 6 + [model["bodies"][i]["armature"] for i in range(1, 13)])
    return M + np.diag(arm), S, dof_body, R, x


def qpos0_constants(model):
    M, S, dof_body, R, x = mass_matrix(model, model["qpos0"])
    Minv = np.linalg.inv(M)
    d = np.diag(Minv).copy()
    dof_invweight0 = d.copy()
    dof_invweight0[0:3] = d[0:3].mean()
    dof_invweight0[3:6] = d[3:6].mean()

    def body_invw(bi, ipos):
        c = x[bi] + R[bi] @ np.array(ipos)
        anc = set(ancestors(model, bi))
        J = np.zeros((6, 18))
        for dd in range(18):
            if dof_body[dd] in anc:
                w, v0 = S[0:3, dd], S[3:6, dd]
                J[0:3, dd] = v0 + np.cross(w, c)  # translational jacobian of the COM
                J[3:6, dd] = w
        A = J @ Minv @ J.T
        return [float(np.trace(A[0:3, 0:3]) / 3), float(np.trace(A[3:6, 3:6]) / 3)]

    biw = {
        "pelvis": body_invw(0, model["pelvis_ipos_own"]),
        "torso_link": body_invw(0, model["torso_ipos_own"]),
    }
    for i in range(1, 13):
        biw[model["bodies"][i]["name"]] = body_invw(i, model["bodies"][i]["ipos"])
    # tracked-body slot -> translational invweight (0,1 feet; 2,3 shins; 4 torso; 5 pelvis)
    slot_body = ["left_ankle_roll_link", "right_ankle_roll_link", "left_knee_link", "right_knee_link", "torso_link", "pelvis"]
    return dict(
        dof_invweight0=dof_invweight0.tolist(),
        body_invweight0=biw,
        slot_invweight_tran=[biw[n][0] for n in slot_body],
        meaninertia=float(np.mean(np.diag(M))),
    )


# ----------------------------------------------------------------------------------------------
def c_array(name, arr, ctype="double"):
    a = np.asarray(arr, dtype=np.float64)
    flat = ", ".join(repr(float(v)) for v in a.reshape(-1))
    dims = "".join(f"[{d}]" for d in a.shape)
    return f"static const {ctype} {name}{dims} = {{{flat}}};\n"


def emit_header(model, path):
    B = model["bodies"]
    s = "/* GENERATED by tools/compile_model.py from the reference MJCF\n"
    s += f" * ({model['source']}); DATA ONLY. Do not edit. */\n"
    s += "#ifndef H1V2_MODEL_H12_H\n#define H1V2_MODEL_H12_H\n"
    s += "#define H1V2_NBODY 13\n#define H1V2_NV 18\n#define H1V2_NQ 19\n#define H1V2_NJ 12\n"
    s += f"#define H1V2_NCOLL {len(model['colliders'])}\n#define H1V2_NSLOT 6\n"
    s += f"#define H1V2_TOTAL_MASS {model['total_mass']!r}\n#define H1V2_MEANINERTIA {model['meaninertia']!r}\n"
    s += "static const int h1v2_body_parent[13] = {" + ", ".join(str(b["parent"]) for b in B) + "};\n"
    s += c_array("h1v2_body_pos", [b["pos"] for b in B])
    s += c_array("h1v2_body_mass", [b["mass"] for b in B])
    s += c_array("h1v2_body_ipos", [b["ipos"] for b in B])
    s += c_array("h1v2_body_inertia", [np.array(b["inertia"]).reshape(9) for b in B])
    s += c_array("h1v2_jnt_axis", [b["axis"] for b in B[1:]])
    s += c_array("h1v2_jnt_range", [b["range"] for b in B[1:]])
    s += c_array("h1v2_jnt_frcrange", [b["frcrange"][1] for b in B[1:]])
    s += c_array("h1v2_dof_damping", [B[0]["damping"]] * 6 + [b["damping"] for b in B[1:]])
    s += c_array("h1v2_dof_armature", [B[0]["armature"]] * 6 + [b["armature"] for b in B[1:]])
    s += c_array("h1v2_dof_frictionloss", [B[0]["frictionloss"]] * 6 + [b["frictionloss"] for b in B[1:]])
    s += c_array("h1v2_dof_invweight0", model["dof_invweight0"])
    s += c_array("h1v2_slot_invweight_tran", model["slot_invweight_tran"])
    s += c_array("h1v2_key_qpos", model["key_qpos"])
    s += c_array("h1v2_coll", model["colliders"])
    s += "#endif\n"
    with open(path, "w") as f:
        f.write(s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--xml", default=REF_XML)
    ap.add_argument("--check", action="store_true", help="compare against the committed files instead of writing")
    args = ap.parse_args()
    model = build_model(args.xml)
    jpath = os.path.join(ROOT, "h1v2_isaac_b200", "model", "h12_12dof.json")
    hpath = os.path.join(ROOT, "include", "h1v2_model_h12.h")
    if args.check:
        old = json.load(open(jpath))
        assert json.dumps(old, sort_keys=True) == json.dumps(json.loads(json.dumps(model)), sort_keys=True), "model drift"
        print("model matches committed json")
        return
    with open(jpath, "w") as f:
        json.dump(model, f, indent=1)
    emit_header(model, hpath)
    print(f"total mass {model['total_mass']:.7f} kg, fused {model['n_fused_into_root']} bodies into root")
    print("root mass", model["bodies"][0]["mass"], "com", model["bodies"][0]["ipos"])
    print("dof_invweight0", np.round(model["dof_invweight0"], 5))
    print("slot_invweight_tran", np.round(model["slot_invweight_tran"], 5), "meaninertia", model["meaninertia"])


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Offline model compiler: reference XML + STL + pose tables -> assets/model_card.bin.

Runs ONLY in the authoring container (it reads /root/reference); the blob it
writes is committed so that nothing at run time needs the reference tree.

What it restates (the reference delegates all of this to the closed MuJoCo 1.50
model compiler; see SURVEY.md Appendix A, "Mesh compile"):
  * STL -> legacy (surface-pyramid, |volume|) centre of mass + inertia,
    principal axes by the Jacobi/quaternion scheme MuJoCo's mju_eig3 uses
    (eigenvalues descending, proper rotations only), geom frame = that frame,
    geom_size = half sizes of the equivalent inertia box, rbound from the AABB;
  * convex hull (qhull 'Qt') + vertex adjacency graph for hill-climb support;
  * the kinematic tree / inertial / site constants of
    kinova_description/j2s7s300_end_effector_v1_<Shape>.xml  (XML:54-153);
  * body/dof/tendon invweight0 at qpos0 (used by the constraint regulariser);
  * the with_noise pose tables (ENV:1008-1054,1245) incl. the 5-character euler
    truncation of write_xml (ENV:870-874);
  * raw triangles + a flattened BVH per mesh for the rangefinder rays.
"""
import os
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np
from scipy.spatial import ConvexHull

REF = "/root/reference/gym-kinova-gripper/gym_kinova_gripper/envs/kinova_description"
SHAPES = ["CubeS", "CubeB", "CylinderS", "CylinderB", "Cube45S", "Cube45B", "Cone1S", "Cone1B",
          "Cone2S", "Cone2B", "Vase1S", "Vase1B", "Vase2S", "Vase2B",
          # the medium sizes (ENV:150-178; MAIN:1271 test_sizes = ["M"]) are appended so that the ids of the 14 above never change
          "CubeM", "CylinderM", "Cube45M", "Cone1M", "Cone2M", "Vase1M", "Vase2M",
          # the lemon (ENV:198-201): the one single-geom object among the later additions; its XMLs are named <size>lemon, the
          # three sizes are one STL under three mesh scales, and the geom is rotated in the body (euler="1.57 0 0")
          "LemonS", "LemonM", "LemonB"]
XML_FILE = {"LemonS": "slemon", "LemonM": "mlemon", "LemonB": "blemon"}
ORIENTS = ["normal", "rotated", "top"]
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kinovagrasping_b200", "assets",
                   "model_card.bin")


# ----------------------------------------------------------------------------- quaternion helpers
def qmul(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def q2m(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def euler2quat(e):
    """MuJoCo default eulerseq 'xyz' (intrinsic): q = qx * qy * qz."""
    q = np.array([1.0, 0, 0, 0])
    for i in range(3):
        h = 0.5 * e[i]
        r = np.array([np.cos(h), 0, 0, 0])
        r[i + 1] = np.sin(h)
        q = qmul(q, r)
    return q / np.linalg.norm(q)


def eig3(mat):
    """Restatement of MuJoCo's mju_eig3: Jacobi sweeps accumulated in a quaternion,
    then a 3-pass bubble sort to descending eigenvalues done with 90-degree rotations."""
    eps = 1e-12
    quat = np.array([1.0, 0, 0, 0])
    for _ in range(500):
        ev = q2m(quat)
        D = ev.T @ mat @ ev
        eigval = np.array([D[0, 0], D[1, 1], D[2, 2]])
        if abs(D[0, 1]) > abs(D[0, 2]) and abs(D[0, 1]) > abs(D[1, 2]):
            rk, ck, rotk = 0, 1, 2
        elif abs(D[0, 2]) > abs(D[1, 2]):
            rk, ck, rotk = 0, 2, 1
        else:
            rk, ck, rotk = 1, 2, 0
        if abs(D[rk, ck]) < eps:
            break
        tau = (D[ck, ck] - D[rk, rk]) / (2 * D[rk, ck])
        if tau >= 0:
            t = 1.0 / (tau + np.sqrt(1 + tau * tau))
        else:
            t = -1.0 / (-tau + np.sqrt(1 + tau * tau))
        c = 1.0 / np.sqrt(1 + t * t)
        if c > 1.0 - eps:
            break
        tmp = np.zeros(4)
        tmp[rotk + 1] = -np.sqrt(0.5 - 0.5 * c) if tau >= 0 else np.sqrt(0.5 - 0.5 * c)
        if rotk == 1:
            tmp[rotk + 1] = -tmp[rotk + 1]
        tmp[0] = np.sqrt(1.0 - tmp[rotk + 1] ** 2)
        tmp /= np.linalg.norm(tmp)
        quat = qmul(quat, tmp)
        quat /= np.linalg.norm(quat)
    for j in range(3):
        j1 = j % 2
        if eigval[j1] + eps < eigval[j1 + 1]:
            eigval[j1], eigval[j1 + 1] = eigval[j1 + 1], eigval[j1]
            tmp = np.zeros(4)
            tmp[0] = 0.707106781186548
            tmp[(j1 + 2) % 3 + 1] = tmp[0]
            quat = qmul(quat, tmp)
            quat /= np.linalg.norm(quat)
    # eigenvalues must be re-read from the final frame (the swap only relabels them)
    ev = q2m(quat)
    D = ev.T @ mat @ ev
    return np.array([D[0, 0], D[1, 1], D[2, 2]]), quat


# ----------------------------------------------------------------------------- meshes
def load_stl(path, scale=1.0):
    b = open(path, "rb").read()
    if b[:5] == b"solid" and b"facet normal" in b[:400]:
        # ASCII STL (only CubeM.stl, written by stlwrite.m).  MuJoCo's loader is documented as binary-only, so whether the
        # reference's CubeM ever loaded under MuJoCo 1.50 is unknown; its triangles are taken as written, in float32.
        v = np.array([ln.split()[1:4] for ln in b.decode("ascii", "replace").splitlines() if ln.strip().startswith("vertex")], dtype=np.float64)
        assert len(v) % 3 == 0 and len(v) > 0, path
        return (v.reshape(-1, 3, 3).astype(np.float32) * np.float32(scale)).astype(np.float64)
    n = struct.unpack("<I", b[80:84])[0]
    assert 84 + 50 * n == len(b), path
    a = np.frombuffer(b[84:], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
    # MuJoCo keeps mesh vertices in float32; the scale is applied in float32 as well
    return (a["v"].astype(np.float32) * np.float32(scale)).astype(np.float64)


def mesh_inertial(tris):
    """Legacy (pre-2.1.2) MuJoCo mesh inertia: pyramids from the area-weighted face
    centroid with |volume| (over-counts concave meshes), then exact tetra second moments
    about the resulting centre of mass."""
    v0, v1, v2 = tris[:, 0], tris[:, 1], tris[:, 2]
    cr = np.cross(v1 - v0, v2 - v0)
    area = 0.5 * np.linalg.norm(cr, axis=1)
    ok = area > 1e-20
    nrm = np.zeros_like(cr)
    nrm[ok] = cr[ok] / (2 * area[ok, None])
    cen = (v0 + v1 + v2) / 3.0
    facecen = (area[:, None] * cen).sum(0) / area.sum()
    vol = np.abs(((cen - facecen) * nrm).sum(1) * area / 3.0)
    volume = vol.sum()
    com = (vol[:, None] * (0.75 * cen + 0.25 * facecen)).sum(0) / volume
    a, b, c = v0 - com, v1 - com, v2 - com
    det = np.abs(np.einsum("ij,ij->i", a, np.cross(b, c))) / 6.0
    P = np.zeros((3, 3))
    for j in range(3):
        for k in range(3):
            s = (2 * (a[:, j] * a[:, k] + b[:, j] * b[:, k] + c[:, j] * c[:, k])
                 + a[:, j] * b[:, k] + a[:, k] * b[:, j]
                 + a[:, j] * c[:, k] + a[:, k] * c[:, j]
                 + b[:, j] * c[:, k] + b[:, k] * c[:, j])
            P[j, k] = (det / 20.0 * s).sum()
    vol2 = det.sum()
    inertia = np.trace(P) * np.eye(3) - P
    return com, vol2, inertia


class Mesh:
    def __init__(self, name, tris):
        self.name = name
        com, vol, inertia = mesh_inertial(tris)
        eigval, quat = eig3(inertia)
        R = q2m(quat)
        self.com, self.volume, self.eigval, self.quat = com, vol, eigval, quat
        # vertices in the mesh's own (inertial) frame
        self.tris = ((tris.reshape(-1, 3) - com) @ R).reshape(-1, 3, 3)
        pts = self.tris.reshape(-1, 3)
        # equivalent inertia box half sizes
        e = eigval
        self.boxsz = 0.5 * np.sqrt(6.0 * np.array([e[1] + e[2] - e[0], e[0] + e[2] - e[1], e[0] + e[1] - e[2]]) / vol)
        amax = np.maximum(np.abs(pts.min(0)), np.abs(pts.max(0)))
        self.rbound = float(np.linalg.norm(amax))
        upts = np.unique(pts, axis=0)
        hull = ConvexHull(upts)
        vid = hull.vertices
        remap = -np.ones(len(upts), dtype=np.int64)
        remap[vid] = np.arange(len(vid))
        self.hull = upts[vid]
        faces = remap[hull.simplices]
        adj = [set() for _ in vid]
        for f in faces:
            for i in range(3):
                adj[f[i]].add(int(f[(i + 1) % 3]))
                adj[f[i]].add(int(f[(i + 2) % 3]))
        self.adj = [sorted(a) for a in adj]
        self.nface = len(faces)


# ----------------------------------------------------------------------------- BVH for rays
def build_bvh(tris, leaf=4):
    """Flattened median-split AABB tree.  nodes: [lo(3), hi(3)], meta: [left|first, count] with
    count==0 for inner nodes (children are left, left+1)."""
    n = len(tris)
    cen = tris.mean(1)
    lo_t, hi_t = tris.min(1), tris.max(1)
    order = np.arange(n)
    nodes_lo, nodes_hi, meta = [], [], []

    def alloc():
        nodes_lo.append(None); nodes_hi.append(None); meta.append(None)
        return len(meta) - 1

    root = alloc()
    stack = [(root, 0, n)]
    while stack:
        node, s, e = stack.pop()
        idx = order[s:e]
        lo, hi = lo_t[idx].min(0), hi_t[idx].max(0)
        nodes_lo[node], nodes_hi[node] = lo, hi
        if e - s <= leaf:
            meta[node] = (s, e - s)
            continue
        ax = int(np.argmax(cen[idx].max(0) - cen[idx].min(0)))
        sub = idx[np.argsort(cen[idx, ax], kind="stable")]
        order[s:e] = sub
        mid = (s + e) // 2
        l = alloc(); r = alloc()
        assert r == l + 1
        meta[node] = (l, 0)
        stack.append((l, s, mid)); stack.append((r, mid, e))
    return (np.array(nodes_lo), np.array(nodes_hi), np.array(meta, dtype=np.int32), tris[order])


# ----------------------------------------------------------------------------- blob writer
class Blob:
    def __init__(self):
        self.sections = []

    def add(self, name, arr, dtype):
        code = {"f64": 0, "i32": 1, "f32": 2}[dtype]
        npdt = {"f64": "<f8", "i32": "<i4", "f32": "<f4"}[dtype]
        a = np.ascontiguousarray(np.asarray(arr, dtype=npdt)).ravel()
        assert len(name) < 24
        self.sections.append((name, code, a))

    def write(self, path):
        n = len(self.sections)
        head = 16 + 40 * n
        off = head
        table = b""
        datas = []
        for name, code, a in self.sections:
            raw = a.tobytes()
            pad = (-len(raw)) % 8
            table += struct.pack("<24sIIQ", name.encode(), code, a.size, off)
            datas.append(raw + b"\0" * pad)
            off += len(raw) + pad
        with open(path, "wb") as f:
            f.write(struct.pack("<8sII", b"KGMC0001", n, 0))
            f.write(table)
            for d in datas:
                f.write(d)
        return off


# ----------------------------------------------------------------------------- XML
def parse_xml(shape):
    root = ET.parse(f"{REF}/j2s7s300_end_effector_v1_{XML_FILE.get(shape, shape)}.xml").getroot()
    fv = lambda s: np.array([float(x) for x in s.split()])
    opt = root.find("option")
    out = {"timestep": float(opt.get("timestep")), "impratio": float(opt.get("impratio"))}
    dflt = root.find("default")
    out["margin"] = float(dflt.find("geom").get("margin"))
    out["damping"] = float(dflt.find("joint").get("damping"))
    out["armature"] = float(dflt.find("joint").get("armature"))
    wb = root.find("worldbody")
    rootb = wb.find("body[@name='root']")
    l7 = rootb.find("body[@name='j2s7s300_link_7']")
    out["l7_pos"] = fv(l7.get("pos"))
    ine = l7.find("inertial")
    out["l7_inertial"] = np.concatenate([fv(ine.get("pos")), [float(ine.get("mass"))], fv(ine.get("diaginertia"))])
    out["slide_axes"] = np.array([fv(j.get("axis")) for j in l7.findall("joint")])
    out["slide_range"] = np.array([fv(j.get("range")) for j in l7.findall("joint")])
    sites = []
    for s in l7.findall("site"):
        q = fv(s.get("quat"))
        sites.append((s.get("name"), 0, fv(s.get("pos")), q / np.linalg.norm(q)))
    fingers = []
    for k, fb in enumerate(l7.findall("body")):
        q = fv(fb.get("quat"))
        ine = fb.find("inertial")
        jr = fv(fb.find("joint").get("range"))
        tb = fb.find("body")
        ine2 = tb.find("inertial")
        fingers.append(dict(
            pos=fv(fb.get("pos")), quat=q / np.linalg.norm(q),
            inertial=np.concatenate([fv(ine.get("pos")), [float(ine.get("mass"))], fv(ine.get("diaginertia"))]),
            range=jr, tip_pos=fv(tb.get("pos")),
            tip_inertial=np.concatenate([fv(ine2.get("pos")), [float(ine2.get("mass"))], fv(ine2.get("diaginertia"))])))
        for s in fb.findall("site"):
            sites.append((s.get("name"), 1 + k, fv(s.get("pos")), euler2quat(fv(s.get("euler")))))
        for s in tb.findall("site"):
            sites.append((s.get("name"), 4 + k, fv(s.get("pos")), euler2quat(fv(s.get("euler")))))
    out["fingers"] = fingers
    out["sites"] = sites
    og = wb.find("body[@name='object']").find("geom")
    out["obj_mass"] = float(og.get("mass"))
    assert len(wb.find("body[@name='object']").findall(".//geom")) == 1, shape + ": multi-geom objects are not supported"
    am = {a.get("name"): a for a in root.find("asset").findall("mesh")}[og.get("mesh")]
    out["obj_mesh_file"], out["obj_mesh_scale"] = am.get("file"), fv(am.get("scale"))
    out["obj_geom_euler"] = fv(og.get("euler")) if og.get("euler") else np.zeros(3)
    out["xml_file"] = f"j2s7s300_end_effector_v1_{XML_FILE.get(shape, shape)}.xml"
    pairs = []
    for p in root.find("contact").findall("pair"):
        pairs.append((p.get("geom1"), p.get("geom2"), int(p.get("condim")), fv(p.get("friction"))))
    out["pairs"] = pairs
    acts = []
    for a in root.find("actuator"):
        acts.append((a.tag, a.get("joint"), float(a.get("kv", "0")), float(a.get("gear", "1")),
                     fv(a.get("ctrlrange")) if a.get("ctrlrange") else None))
    out["actuators"] = acts
    tend = []
    for t in root.find("tendon").findall("fixed"):
        tend.append([(j.get("joint"), float(j.get("coef"))) for j in t.findall("joint")])
    out["tendons"] = tend
    sens = [(s.tag, s.get("joint") or s.get("site")) for s in root.find("sensor")]
    out["sensors"] = sens
    return out


# ----------------------------------------------------------------------------- invweight0
def hand_mass_matrix0(x, palm_unused=None):
    """Joint-space inertia of the 9-dof hand at qpos0 from body Jacobians (independent of the
    CRBA used by the oracle and the kernels).  Orientation of link_7 is irrelevant."""
    nv = 9
    bodies = []  # (mass, com_world, Rworld, diagI, jacp(3xnv), jacr(3xnv))
    axes = x["slide_axes"]
    p7 = x["l7_pos"]
    def add(mass, com, R, diagI, chain):
        jp = np.zeros((3, nv)); jr = np.zeros((3, nv))
        for d in range(3):
            jp[:, d] = axes[d]
        for (dof, anchor, axis) in chain:
            jr[:, dof] = axis
            jp[:, dof] = np.cross(axis, com - anchor)
        bodies.append((mass, com, R, diagI, jp, jr))
    I7 = x["l7_inertial"]
    add(I7[3], p7 + I7[0:3], np.eye(3), I7[4:7], [])
    for k, f in enumerate(x["fingers"]):
        Rf = q2m(f["quat"]); pf = p7 + f["pos"]
        ax = Rf[:, 2]
        fi = f["inertial"]
        add(fi[3], pf + Rf @ fi[0:3], Rf, fi[4:7], [(3 + 2 * k, pf, ax)])
        pt = pf + Rf @ f["tip_pos"]
        ti = f["tip_inertial"]
        add(ti[3], pt + Rf @ ti[0:3], Rf, ti[4:7], [(3 + 2 * k, pf, ax), (4 + 2 * k, pt, ax)])
    M = np.zeros((nv, nv))
    for (m, com, R, dI, jp, jr) in bodies:
        Iw = R @ np.diag(dI) @ R.T
        M += m * jp.T @ jp + jr.T @ Iw @ jr
    M += x["armature"] * np.eye(nv)
    return M, bodies


def main():
    x = parse_xml("CubeS")
    for s in SHAPES[1:]:
        y = parse_xml(s)
        for key in ("timestep", "impratio", "margin", "damping", "armature", "obj_mass"):
            assert x[key] == y[key]
        assert np.allclose(x["l7_pos"], y["l7_pos"]) and len(y["pairs"]) == 8
    # the reference's qpos order (XML:62-64,82,91,...) : slides, f1, ft1, f2, ft2, f3, ft3, object
    blob = Blob()
    blob.add("opt", [x["timestep"], x["impratio"], x["margin"], x["damping"], x["armature"], x["obj_mass"],
                     9.81, 0.02, 1.0, 0.9, 0.95, 0.001], "f64")
    blob.add("l7_pos", x["l7_pos"], "f64")
    blob.add("l7_inertial", x["l7_inertial"], "f64")
    blob.add("slide_axes", x["slide_axes"], "f64")
    blob.add("slide_range", x["slide_range"], "f64")
    blob.add("finger_pos", [f["pos"] for f in x["fingers"]], "f64")
    blob.add("finger_quat", [f["quat"] for f in x["fingers"]], "f64")
    blob.add("finger_inertial", [f["inertial"] for f in x["fingers"]], "f64")
    blob.add("finger_range", [f["range"] for f in x["fingers"]], "f64")
    blob.add("tip_pos", [f["tip_pos"] for f in x["fingers"]], "f64")
    blob.add("tip_inertial", [f["tip_inertial"] for f in x["fingers"]], "f64")
    # sites in sensor order: palm, palm_1..4, then per finger prox, prox_1, dist, dist_1  (XML:265-288)
    sd = {s[0]: s for s in x["sites"]}
    order = [s[1] for s in x["sensors"] if s[0] == "rangefinder"]
    assert len(order) == 17
    blob.add("site_body", [sd[n][1] for n in order], "i32")
    blob.add("site_pos", [sd[n][2] for n in order], "f64")
    blob.add("site_zaxis", [q2m(sd[n][3])[:, 2] for n in order], "f64")
    # _get_finger_obj_dist site order (ENV:539) as indices into the sensor order above
    fo = ["f1_prox", "f1_prox_1", "f2_prox", "f2_prox_1", "f3_prox", "f3_prox_1",
          "f1_dist", "f1_dist_1", "f2_dist", "f2_dist_1", "f3_dist", "f3_dist_1"]
    blob.add("fod_site", [order.index(n) for n in fo], "i32")
    # actuators (XML:214-222)
    assert [a[0] for a in x["actuators"]] == ["velocity", "motor"] * 3 + ["velocity"] * 3
    blob.add("act_kv", [a[2] for a in x["actuators"]], "f64")
    blob.add("act_gear", [a[3] for a in x["actuators"]], "f64")
    blob.add("act_ctrlrange", [a[4] if a[4] is not None else [0, 0] for a in x["actuators"]], "f64")
    blob.add("pair_friction", [p[3][0] for p in x["pairs"]], "f64")  # object-ground, then 6 fingers, palm

    # ---- meshes: 0 palm, 1 proximal, 2 distal, 3.. objects
    meshes = [Mesh("palm", load_stl(f"{REF}/meshes/hand_3finger.STL")),
              Mesh("prox", load_stl(f"{REF}/meshes/finger_proximal.STL")),
              Mesh("dist", load_stl(f"{REF}/meshes/finger_distal.STL"))]
    flags = []
    for s in SHAPES:
        y = parse_xml(s)
        sc = y["obj_mesh_scale"]
        assert sc[0] == sc[1] == sc[2]
        if s not in XML_FILE:
            assert y["obj_mesh_file"] == f"{s}.stl" and sc[0] == 0.001 and not y["obj_geom_euler"].any(), s
        msh = Mesh(s, load_stl(f"{REF}/meshes/{y['obj_mesh_file']}", float(sc[0])))
        if y["obj_geom_euler"].any():      # geom frame in the body = the geom's own (euler) frame composed with the mesh's inertial frame
            qe = euler2quat(y["obj_geom_euler"])
            msh.com, msh.quat = q2m(qe) @ msh.com, qmul(qe, msh.quat)
        meshes.append(msh)
        flags.append(1 if "lemon" in y["xml_file"] else 0)      # _get_obj_size: `'lemon' in self.filename` swaps size[0], size[2] (ENV:721)
    blob.add("shape_flags", flags, "i32")
    mi, hv, adj_start, adj, tri_v, bvh_lo, bvh_hi, bvh_meta = [], [], [], [], [], [], [], []
    geomc = []
    for m in meshes:
        lo, hi, meta, tris = build_bvh(m.tris)
        nb = sum(len(b) for b in bvh_meta)
        mi.append([sum(len(h) for h in hv), len(m.hull), len(adj), sum(len(t) for t in tri_v), len(tris), nb, len(meta), m.nface])
        hv.append(m.hull)
        for a in m.adj:
            adj_start.append(len(adj)); adj.extend(a)
        tri_v.append(tris); bvh_lo.append(lo); bvh_hi.append(hi); bvh_meta.append(meta)
        geomc.append(np.concatenate([m.com, m.quat, [m.rbound], m.boxsz, [m.volume], m.eigval]))
        print(f"{m.name:10s} hull {len(m.hull):4d} v {m.nface:4d} f  maxdeg {max(len(a) for a in m.adj):3d} tris {len(tris):6d} bvh {len(meta):5d} "
              f"com {np.round(m.com, 5)} quat {np.round(m.quat, 4)} size {np.round(m.boxsz, 5)} rbound {m.rbound:.4f}")
    adj_start.append(len(adj))
    blob.add("mesh_index", mi, "i32")
    blob.add("mesh_geom", geomc, "f64")       # [com3, quat4, rbound, size3, volume, eig3] per mesh
    blob.add("hull_vert", np.concatenate(hv), "f64")
    blob.add("hull_adj_start", adj_start, "i32")
    blob.add("hull_adj", adj, "i32")
    blob.add("tri_vert", np.concatenate(tri_v), "f32")
    blob.add("bvh_lo", np.concatenate(bvh_lo), "f32")
    blob.add("bvh_hi", np.concatenate(bvh_hi), "f32")
    blob.add("bvh_meta", np.concatenate(bvh_meta), "i32")

    # ---- invweight0 (MuJoCo setInertia at qpos0)
    M, bodies = hand_mass_matrix0(x)
    Minv = np.linalg.inv(M)
    binv = []
    for (m_, com, R, dI, jp, jr) in bodies:
        A = np.vstack([jp, jr]) @ Minv @ np.vstack([jp, jr]).T
        binv.append([np.trace(A[:3, :3]) / 3, np.trace(A[3:, 3:]) / 3])
    blob.add("hand_M0", M, "f64")
    blob.add("hand_body_invw0", binv, "f64")                 # link7, f1, t1, f2, t2, f3, t3
    blob.add("hand_dof_invw0", np.diag(Minv), "f64")
    ten = []
    for k in range(3):
        J = np.zeros(9); J[3 + 2 * k] = 1; J[4 + 2 * k] = -2
        ten.append(J @ Minv @ J)
    blob.add("tendon_invw0", ten, "f64")
    oinv = []
    for m in meshes[3:]:
        inertia = m.eigval * x["obj_mass"] / m.volume
        # free joint at the body origin, inertial frame at com/quat; dofs: 3 world-frame translations,
        # 3 body-frame rotations (body frame == world at qpos0)
        R = q2m(m.quat); Iw = R @ np.diag(inertia) @ R.T
        jp = np.hstack([np.eye(3), -np.array([[0, -m.com[2], m.com[1]], [m.com[2], 0, -m.com[0]], [-m.com[1], m.com[0], 0]])])
        jr = np.hstack([np.zeros((3, 3)), np.eye(3)])
        Mo = x["obj_mass"] * jp.T @ jp + jr.T @ Iw @ jr + x["armature"] * np.eye(6)
        Mi = np.linalg.inv(Mo)
        A = np.vstack([jp, jr]) @ Mi @ np.vstack([jp, jr]).T
        d = np.diag(Mi)
        oinv.append([np.trace(A[:3, :3]) / 3, np.trace(A[3:, 3:]) / 3, d[:3].mean(), d[3:].mean()])
    blob.add("obj_invw0", oinv, "f64")                       # body trans, body rot, dof trans, dof rot

    # ---- pose tables (ENV:1008-1054): raw rows + the euler after write_xml's 5-char truncation
    def trunc(v):
        s = str(np.float64(v))
        return float(s[0:min(5, len(s))])
    for mode, nrow in (("train", 4500), ("test", 500)):
        tab = np.zeros((len(SHAPES), len(ORIENTS), nrow, 9), dtype=np.float64)
        for si, s in enumerate(SHAPES):
            for oi, o in enumerate(ORIENTS):
                rows = np.loadtxt(f"{REF}/obj_hand_coords/with_noise/{mode}_coords/{o}/{s}.txt", delimiter=",")
                assert rows.shape == (nrow, 6)
                tab[si, oi, :, 0:6] = rows
                tab[si, oi, :, 6:9] = [[trunc(r[3]), trunc(r[4]), trunc(r[5])] for r in rows]
        blob.add(f"pose_{mode}", tab, "f32")
    size = blob.write(OUT)
    print("wrote", OUT, size, "bytes")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generate tests/golden/reset_golden.npz by executing the REFERENCE's own reset helpers (unmodified, imported from
/root/reference): Generate_Latin_Square, objects_file_to_list / get_object, select_orienation,
sample_initial_object_hand_pos, determine_hand_location, write_xml's 5-character euler truncation
(kinova_gripper_env.py:851-881, 895-1054, 1180-1222, 1286-1307).  Runs only in the authoring container; the fixture is
committed so that tests (CPU and GPU box) never need /root/reference.  SURVEY 8a row I1."""
import contextlib
import io
import json
import os
import re
import shutil
import sys
import tempfile

import numpy as np

REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
import gen_golden as GG      # the stubbing machinery (mujoco_py, gym, ... stand-ins)

REF = GG.REF
OUT = os.path.join(REPO, "tests", "golden")
SHAPES = ["CubeS", "CubeB", "CylinderS", "CylinderB", "Cube45S", "Cube45B", "Cone1S", "Cone1B",
          "Cone2S", "Cone2B", "Vase1S", "Vase1B", "Vase2S", "Vase2B"]


def new_env(K):
    env = object.__new__(K)
    env.objects = {}
    env.obj_keys = []
    env.obj_size = "m"
    env.orientation = "normal"
    env.Tfw = np.zeros([4, 4])
    env.obj_coord_region = None
    env.file_dir = os.path.join(REF, "gym_kinova_gripper/envs")
    # all_objects as __init__ builds it for the CubeS..Vase2B set (ENV:138-179)
    env.all_objects = {k: f"/kinova_description/j2s7s300_end_effector_v1_{k}.xml" for k in SHAPES}
    return env


def main():
    GG.install_stubs()
    sys.path.insert(0, REF)
    import importlib.util
    spec = importlib.util.spec_from_file_location("kinova_gripper_env", os.path.join(REF, "gym_kinova_gripper/envs/kinova_gripper_env.py"))
    kge = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(kge)
    K = kge.KinovaGripper_Env
    out = {}
    tmp = tempfile.mkdtemp()
    cwd = os.getcwd()
    quiet = lambda: contextlib.redirect_stdout(io.StringIO())

    # ---- 1. Generate_Latin_Square
    cases = [(["CubeS", "CubeB"], 7), (SHAPES, 40), (["CubeS", "CylinderB", "Cone2S"], 10), (["Vase1S"], 3), (SHAPES[:5], 25)]
    os.chdir(tmp)
    ls = []
    for keys, n in cases:
        env = new_env(K)
        with quiet():
            env.Generate_Latin_Square(n, "objects.csv", keys)
        ls.append({"keys": keys, "n": n, "out": list(env.obj_keys)})
        os.remove("objects.csv")
    out["latin_square"] = json.dumps(ls)

    # ---- 2. the shape sequence a run of fresh env objects sees (MAIN:381): select_object -> objects_file_to_list -> get_object
    keys = ["CubeS", "CylinderB", "Cone2S"]
    seq, nkeys = [], []
    for ep in range(9):
        env = new_env(K)
        with quiet():
            empty = env.check_obj_file_empty("objects.csv")
            shape = env.select_object("env", keys, None)
        seq.append(shape); nkeys.append(len(env.obj_keys))
        assert empty is False
    out["select_object_seq"] = json.dumps({"keys": keys, "seq": seq, "nkeys": nkeys})
    os.remove("objects.csv")
    os.chdir(cwd)

    # ---- 3. select_orienation
    so = []
    env = new_env(K)
    for seed in range(40):
        np.random.seed(seed)
        so.append([env.select_orienation("CubeS", "random"), env.select_orienation("Vase2B", "normal"), float(np.random.rand())])
    out["select_orienation"] = json.dumps(so)

    # ---- 4. sample_initial_object_hand_pos on the real coordinate files (CWD = gym-kinova-gripper, as the reference requires)
    os.chdir(REF)
    sp = []
    for shape, orient, mode, region, seed in [("CubeS", "normal", "train", None, 0), ("CubeB", "top", "train", None, 1), ("Cone1S", "rotated", "test", None, 2),
                                              ("Vase2B", "normal", "train", "left", 3), ("CylinderS", "top", "train", "center", 4),
                                              ("Cube45B", "rotated", "train", "right", 5), ("CubeS", "normal", "test", "target", 6),
                                              ("CubeB", "normal", "train", "origin", 7)]:
        env = new_env(K)
        fn = f"gym_kinova_gripper/envs/kinova_description/obj_hand_coords/with_noise/{mode}_coords/{orient}/{shape}.txt"
        np.random.seed(seed)
        r = env.sample_initial_object_hand_pos(fn, with_noise=True, orient_idx=None, region=region)
        sp.append({"shape": shape, "orient": orient, "mode": mode, "region": region, "seed": seed,
                   "out": [float(v) for v in r[:6]], "idx": None if r[6] is None else int(r[6])})
    out["sample_pose"] = json.dumps(sp)
    os.chdir(cwd)

    # ---- 5. write_xml's truncation: what ends up in the XML for the hand eulers of table rows
    tr_in, tr_out = [], []
    work = os.path.join(tmp, "kd")
    os.makedirs(work + "/kinova_description", exist_ok=True)
    src = os.path.join(REF, "gym_kinova_gripper/envs/kinova_description/j2s7s300_end_effector_v1_CubeB.xml")
    rng = np.random.default_rng(5)
    from kinovagrasping_b200.modelcard import ModelCard
    pose = ModelCard().pose("train")
    rows = [(int(rng.integers(0, 14)), int(rng.integers(0, 3)), int(rng.integers(1, 4500))) for _ in range(300)]
    rows += [(s, o, 1) for s in range(14) for o in range(3)]
    for (s, o, r) in rows:
        env = new_env(K)
        env.file_dir = work
        env.filename = "/kinova_description/x.xml"
        env.wrist_pose = np.zeros(3)
        shutil.copy(src, work + env.filename)
        coords = os.path.join(REF, "gym_kinova_gripper/envs/kinova_description/obj_hand_coords/with_noise/train_coords", ("normal", "rotated", "top")[o], SHAPES[s] + ".txt")
        line = open(coords).read().splitlines()[r]          # file line r <-> pose-table row r
        vals = [float(v) for v in re.split(r"[ ,]+", line.strip())]
        assert np.allclose(vals[3:6], pose[s, o, r, 3:6], atol=1e-6)
        rot = np.array(vals[3:6])
        try:
            with quiet():
                env.write_xml(rot)          # the file is rewritten before the (stubbed) model reload fails
        except Exception:
            pass
        txt = open(work + env.filename).read()
        m = re.search(r'<body name="j2s7s300_link_7"[^>]*euler="([^"]*)"', txt)
        tr_in.append([s, o, r]); tr_out.append([float(v) for v in m.group(1).split()])
    out["trunc_rows"] = np.array(tr_in); out["trunc_euler"] = np.array(tr_out)

    # ---- 6. determine_hand_location for given (stale) Tfw
    hl = []
    for k in range(12):
        env = new_env(K)
        A = rng.normal(size=(3, 3)); Q, _ = np.linalg.qr(A)
        env.Tfw = np.eye(4); env.Tfw[:3, :3] = Q
        env._sim = GG.FakeSim(None, GG.FakeModel(np.ones((9, 3)) * 0.03)); env.filename = "/kinova_description/x.xml"
        for orient in ("normal", "rotated", "top"):
            env.orientation = orient
            hl.append({"R": Q.tolist(), "orient": orient, "out": [float(v) for v in env.determine_hand_location()]})
    out["hand_location"] = json.dumps(hl)
    np.savez_compressed(os.path.join(OUT, "reset_golden.npz"), **out)
    print("latin cases", len(ls), "shape seq", seq, "orient draws", len(so), "pose samples", len(sp), "truncations", len(tr_in), "hand locations", len(hl))
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()

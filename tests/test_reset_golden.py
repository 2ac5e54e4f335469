"""Reset bookkeeping against fixtures written by the reference's OWN functions
(tools/gen_golden_reset.py imports kinova_gripper_env.py unmodified): Latin-square shape lists, the objects.csv round trip
across env objects, orientation thresholds, pose-row sampling incl. the region quirk, the 5-character euler truncation, and
determine_hand_location -- held by the facade (host logic) and, where the device does the same job, by the kernels."""
import json
import os

import numpy as np
import pytest

from conftest import _emu_lib
from kinovagrasping_b200 import env as kenv
from kinovagrasping_b200.handle import Handle
from kinovagrasping_b200.modelcard import SHAPES, ORIENTATIONS


@pytest.fixture(scope="module")
def G(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "reset_golden.npz")))


class _NoHandle:
    def set_frame_skip(self, n): pass
    def set_action_dim(self, n): pass
    def set_max_episode_steps(self, n): pass
    def close(self): pass


def _env():
    return kenv.KinovaGripper_Env(handle=_NoHandle())


def test_latin_square_matches_reference(G):
    for case in json.loads(str(G["latin_square"])):
        assert kenv.latin_square_keys(case["keys"], case["n"]) == case["out"]
        e = _env()
        kenv._MEM_FILES.clear()
        e.Generate_Latin_Square(case["n"], "objects.csv", case["keys"])
        assert e.get_obj_keys() == case["out"] and kenv._MEM_FILES["objects.csv"] == case["out"]


def test_shape_sequence_over_fresh_env_objects_matches_reference(G):
    ref = json.loads(str(G["select_object_seq"]))
    kenv._MEM_FILES.clear()
    seq, nkeys = [], []
    for _ in ref["seq"]:
        e = _env()
        assert e.check_obj_file_empty("objects.csv") is False
        seq.append(e.select_object("env", ref["keys"], None)); nkeys.append(len(e.get_obj_keys()))
    assert seq == ref["seq"] and nkeys == ref["nkeys"]


def test_select_orienation_matches_reference(G):
    e = _env()
    for seed, (a, b, nxt) in enumerate(json.loads(str(G["select_orienation"]))):
        np.random.seed(seed)
        assert [e.select_orienation("CubeS", "random"), e.select_orienation("Vase2B", "normal")] == [a, b]
        assert float(np.random.rand()) == nxt                    # same number of draws consumed


def test_sample_initial_object_hand_pos_matches_reference(G):
    for c in json.loads(str(G["sample_pose"])):
        e = _env()
        e._shape_id = SHAPES.index(c["shape"]); e.orientation = c["orient"]
        fn = f"gym_kinova_gripper/envs/kinova_description/obj_hand_coords/with_noise/{c['mode']}_coords/{c['orient']}/{c['shape']}.txt"
        np.random.seed(c["seed"])
        r = e.sample_initial_object_hand_pos(fn, with_noise=True, orient_idx=None, region=c["region"])
        assert r[6] == c["idx"]
        assert np.allclose(r[:6], c["out"], rtol=0, atol=1e-7)      # the card stores the file's numbers as fp32


def test_euler_truncation_matches_reference_write_xml(G, card):
    rows, eul = G["trunc_rows"], G["trunc_euler"]
    pose = card.pose("train")
    for (s, o, r), e in zip(rows, eul):
        assert np.array_equal(pose[s, o, r, 6:9], e.astype(np.float32)), (s, o, r, pose[s, o, r, 3:9], e)
        assert kenv.truncate_euler([float(v) for v in np.float64(e)]) == [float(v) for v in e]      # idempotent on its own output


def test_determine_hand_location_matches_reference(G):
    for c in json.loads(str(G["hand_location"])):
        e = _env()
        e.Tfw = np.eye(4); e.Tfw[:3, :3] = np.array(c["R"]); e.orientation = c["orient"]
        assert np.allclose(e.determine_hand_location(), c["out"], rtol=0, atol=1e-15)


def _mix(x):
    M = (1 << 64) - 1
    x = (x + 0x9E3779B97F4A7C15) & M
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
    return x ^ (x >> 31)


def _expected_draw(seed, env_global, episode, rows, random_orient):
    """The device sampler's counter-based stream (csrc/kg_api.cu kg_sample_pose), restated: file line 1..rows-1 uniformly
    (the reference's np.random.randint(0, len(data)) over data = file lines 1.., ENV:1012,1046), orientation from one
    U(0,1) draw against the reference's 0.333 / 0.667 thresholds (ENV:1208-1220)."""
    r = _mix(seed ^ _mix((env_global * 0x100000001B3 + episode) & ((1 << 64) - 1)))
    r2 = _mix(r)
    o = None
    if random_orient:
        u = np.float32((r2 >> 40) * (1.0 / 16777216.0))
        o = 0 if u < np.float32(0.333) else (2 if u > np.float32(0.667) else 1)
    return 1 + int(r % (rows - 1)), o


@pytest.mark.parametrize("hand_location", [0, 1])
def test_device_sampled_reset_against_oracle(backend, oracle, card, hand_location):
    """kg_reset_sampled per env and episode: the pose row it must have drawn for (seed, global env id, episode), the
    5-char-truncated euler of that row, the orientation thresholds, and -- with kg_set_hand_location(1) -- the slide offsets
    of determine_hand_location from the PREVIOUS episode's hand frame (zero for the first); state and all 82 obs entries
    against the oracle reset from those numbers (1e-5)."""
    n, seed, offset = 36, 4242, 1000
    h = Handle(n, lib=backend)
    h.set_env_offset(offset)
    h.set_hand_location(hand_location)
    shape = (np.arange(n) % 14).astype(np.int32)
    pose = card.pose("train")
    import ctypes as C
    sp = shape.ctypes.data_as(C.c_void_p)
    cuda = not backend._name.endswith("libkg_emu.so")
    if cuda:
        import torch
        shape_d = torch.from_numpy(shape).cuda(); obs_d = torch.zeros((82, n), device="cuda")
        sp, op = shape_d.data_ptr(), obs_d.data_ptr()
    else:
        obs_h = np.zeros((82, n), dtype=np.float32); op = obs_h.ctypes.data_as(C.c_void_p)
    prev = [None] * n
    seen_orient = set()
    for episode in range(3):
        h.reset_sampled(0, sp, 3, 0, 0, seed, op)
        h.sync()
        obs = obs_d.cpu().numpy() if cuda else obs_h
        qpos, qvel = h.get_state_host()
        for i in range(n):
            row, o = _expected_draw(seed, offset + i, episode, pose.shape[2], True)
            seen_orient.add(o)
            p = pose[shape[i], o, row]
            hand = oracle.hand_location(prev[i] if hand_location else None, o)
            assert np.abs(qpos[0:3, i] - hand).max() < 1e-6, (episode, i, o, qpos[0:3, i], hand)
            # the obs comparison starts from the SAME (fp32) state on both sides, like every other obs check
            e, ref_obs = oracle.env_reset(int(shape[i]), p[6:9].astype(np.float64), p[0:3].astype(np.float64), hand_pos=qpos[0:3, i].astype(np.float64))
            if hand_location and episode > 0:
                assert (o == 0) == (np.abs(qpos[0:3, i]).max() == 0.0)
            assert np.abs(qpos[9:12, i] - np.array(e.st.qpos[9:12])).max() < 1e-6
            err = np.abs(obs[:, i] - ref_obs) / np.maximum(1.0, np.abs(ref_obs))
            # _get_dot_product (ENV:591-608) normalises |0 - hand_xy|: the 'top' hand offset parks the hand within millimetres
            # of the world z axis, where that direction -- raised to the 20th power -- amplifies the 1.5e-8 m resolution of an
            # fp32 slide position by 20 / |hand_xy|.  The 7 dot-product entries get that conditioning factor, nothing else does.
            hxy = float(np.hypot(*e.d.arr("xpos", (9, 3))[1][:2]))
            tol = np.full(82, 1e-5); tol[75:82] *= max(1.0, 0.05 / hxy)
            assert (err < tol).all(), (episode, i, int(np.argmax(err / tol)), err.max(), hxy)
            prev[i] = e.d
    assert seen_orient == {0, 1, 2}
    h.close()


def test_device_shape_cycle_is_a_latin_square(backend):
    """kg_set_shape_cycle: with no shape array, episode e of global env i is reset to keys[(i + e) % n] -- every env walks
    through all keys, every episode index sees every key equally often (what Generate_Latin_Square provides per env object)."""
    import ctypes as C
    n, offset, keys = 12, 5, [0, 3, 9, 13]
    h = Handle(n, lib=backend)
    h.set_env_offset(offset)
    h.set_shape_cycle(keys)
    cuda = not backend._name.endswith("libkg_emu.so")
    seen = []
    for e in range(6):
        h.reset_sampled(0, 0, 0, 0, 0, 1, 0)
        h.sync()
        if cuda:
            import torch
            obs = torch.zeros((82, n), device="cuda"); rew = torch.zeros(n, device="cuda"); done = torch.zeros(n, dtype=torch.int32, device="cuda")
            h.step(0, 0, obs.data_ptr(), rew.data_ptr(), done.data_ptr()); h.sync()
            o = obs.cpu().numpy()
        else:
            o, _, _ = h.step_host(None, 0)
        seen.append(o[33:36].copy())                        # obs[33:36] = the object's size: identifies the shape
    card_sizes = {}
    ref = Handle(1, lib=backend)
    for k in keys:
        ob = ref.reset_host(np.array([k], dtype=np.int32), np.zeros((3, 1), dtype=np.float32), np.array([[0.0], [0.0], [0.1]], dtype=np.float32))
        card_sizes[k] = ob[33:36, 0].copy()
    for e in range(6):
        for i in range(n):
            assert np.array_equal(seen[e][:, i], card_sizes[keys[(offset + i + e) % len(keys)]]), (e, i)
    h.close(); ref.close()

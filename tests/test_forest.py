"""Grasp-classifier reward (ENV:248, 643-659, 791, 1327, 1372).

* the oracle's numpy forest (oracle/forest.py) against tests/golden/forest.npz -- leaves picked by scikit-learn's own
  compiled traversal on the reference's trees (tools/gen_golden_forest.py), labels confirmed by the installed
  RandomForestClassifier.predict: exact;
* the kernels through the C ABI (host emulation here, CUDA under -m gpu): kg_grasp_predict_host on the golden inputs -- exact
  labels, on-threshold inputs and 10:10 ties included; kg_step with the reward switched on against the oracle's
  _get_reward restatement applied to the kernel's own observations -- reward, info and the once-per-episode flag exact;
* the extraction tool: refuses any pickle global outside its seven, and re-extracting the reference's file reproduces the
  committed asset byte for byte (only where /root/reference exists)."""
import io
import os
import pickle
import sys

import numpy as np
import pytest

from kinovagrasping_b200.handle import Handle, KgError

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
FOREST = os.path.join(ROOT, "kinovagrasping_b200", "assets", "gc_forest.bin")
REF_PKL = "/root/reference/gym-kinova-gripper/gym_kinova_gripper/envs/kinova_description/gc_model.pkl"


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "forest.npz"))


@pytest.fixture(scope="module")
def forest():
    from oracle.forest import Forest
    return Forest(FOREST)


def test_oracle_forest_against_sklearn_golden(forest, golden):
    X, leaves = golden["X"], golden["leaves"]
    assert forest.n_trees == 20 and forest.n_features == 75 and list(forest.classes) == [0.0, 1.0]
    idx = np.r_[0:400, int(golden["n_real"]):int(golden["n_real"]) + 100, len(X) - 1500:len(X)]      # real, Gaussian, on-threshold
    for r in idx:
        got = [forest.leaf(t, X[r]) - int(forest.root[t]) for t in range(forest.n_trees)]
        assert got == list(leaves[r]), r
    P = forest.predict_proba(X[idx])
    assert np.array_equal(P, golden["proba"][idx])                  # same fp64 sums in estimator order
    assert np.array_equal(forest.predict(X[idx]), golden["label"][idx])
    ties = np.flatnonzero(golden["proba"][:, 0] == golden["proba"][:, 1])
    assert len(ties) > 100 and (golden["label"][ties] == 0.0).all()   # numpy argmax: the first class wins a 10:10 vote
    assert np.array_equal(forest.predict(X[ties[:50]]), golden["label"][ties[:50]])


def test_kernel_labels_equal_sklearn_on_all_golden_inputs(backend, golden):
    X, label = golden["X"], golden["label"]
    h = Handle(2048, device=0, lib=backend)
    with pytest.raises(KgError):
        h.set_grasp_reward(True, forest_path=os.path.join(ROOT, "kinovagrasping_b200", "assets", "model_card.bin"))   # not a forest
    with pytest.raises(KgError):
        h.grasp_predict_host(X[:4])                                  # nothing loaded yet
    h.load_grasp_forest()
    out = np.concatenate([h.grasp_predict_host(X[i:i + 2048]) for i in range(0, len(X), 2048)])
    assert out.shape == label.shape and np.array_equal(out.astype(np.float64), label)
    assert 500 < int((out >= 0.3).sum()) < len(out) - 500             # both classes well represented
    h.close()


def test_forest_loader_rejects_walks_that_could_not_end(backend, tmp_path):
    """kg_set_grasp_forest validates the blob before any kernel may walk it: children must lie behind their parent inside their
    own tree (sklearn numbers nodes in construction order), features inside obs[0:75], sections consistent and inside the blob."""
    import struct
    raw = bytearray(open(FOREST, "rb").read())
    n = struct.unpack_from("<I", raw, 8)[0]
    sec = {}
    for i in range(n):
        name, dt, cnt, off = struct.unpack_from("<24sIIQ", raw, 16 + 40 * i)
        sec[name.rstrip(b"\0").decode()] = (16 + 40 * i, dt, cnt, off)

    def broken(edit):
        b = bytearray(raw)
        edit(b)
        path = tmp_path / "bad_forest.bin"
        path.write_bytes(bytes(b))
        h = Handle(4, device=0, lib=backend)
        with pytest.raises(KgError):
            h.load_grasp_forest(str(path))
        h.close()
    broken(lambda b: struct.pack_into("<i", b, sec["node_left"][3], 0))                       # root's child = root: a cycle
    broken(lambda b: struct.pack_into("<i", b, sec["node_right"][3], 10 ** 7))                # child outside the tree
    broken(lambda b: struct.pack_into("<i", b, sec["node_feature"][3], 80))                   # feature beyond obs[0:75]
    broken(lambda b: struct.pack_into("<I", b, sec["node_thresh"][0] + 28, sec["node_thresh"][2] - 1))      # section count mismatch
    broken(lambda b: struct.pack_into("<Q", b, sec["node_proba"][0] + 32, len(raw) - 64))    # section runs past the end of the blob
    h = Handle(4, device=0, lib=backend)
    h.load_grasp_forest()                                                                      # the committed blob itself loads
    h.close()


def test_step_reward_with_grasp_classifier(backend, oracle, forest, card):
    """kg_step with kg_set_grasp_reward(1): total reward = lift + grasp, grasp = 5.0 the first time per episode the forest
    labels the step's obs[0:75] >= 0.3 (ENV:651-659), flag cleared by reset (ENV:1372) -- against oracle/forest.py's
    grasp_reward() fed with the observations the kernels returned (the obs themselves are held by test_parity.py)."""
    from oracle.forest import grasp_reward
    n = 48
    h = Handle(n, device=0, lib=backend)
    shape = (np.arange(n) % 14).astype(np.int32)
    pose = card.pose("train")
    rng = np.random.default_rng(5)
    rows = rng.integers(1, pose.shape[2], size=n)
    # variant B of SURVEY 8d: the side-grasp hand euler ('top' table) with the 'normal' object position
    euler = np.stack([pose[shape[i], 2, rows[i], 6:9] for i in range(n)], 1).astype(np.float32)
    objp = np.stack([pose[shape[i], 0, rows[i], 0:3] for i in range(n)], 1).astype(np.float32)
    h.set_grasp_reward(True)
    given = np.zeros(n, dtype=bool)
    paid = 0
    for episode in range(2):
        h.reset_host(shape, euler, objp)
        g0, f0 = h.get_grasp_reward_host()
        assert not f0.any()                                          # Grasp_Reward = False after reset
        given[:] = False
        for step in range(30):
            obs, rew, done = h.step_host(None, 0)
            g, flags = h.get_grasp_reward_host()
            for i in range(n):
                want, given[i] = grasp_reward(forest, obs[:, i].astype(np.float64), bool(given[i]))
                z_done = rew[i] >= 50.0
                assert g[i] == want and flags[i] == given[i], (episode, step, i)
                assert rew[i] == (50.0 if z_done else 0.0) + want
                paid += want > 0
    assert paid >= n // 4, paid                                      # the expert's grasps are recognised in a good share of the envs
    # a partial reset re-arms only the envs it touches
    mask = (np.arange(n) % 2).astype(np.int32)
    h.reset_host(shape, euler, objp, mask=mask)
    _, flags = h.get_grasp_reward_host()
    assert not flags[mask == 1].any() and np.array_equal(flags[mask == 0], given[mask == 0])
    # switched off: lift reward only, flags untouched
    h.set_grasp_reward(False)
    obs, rew, done = h.step_host(None, 0)
    assert set(np.unique(rew)) <= {0.0, 50.0}
    h.close()


def test_data_only_unpickler_refuses_code():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from extract_forest import DataOnlyUnpickler

    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))
    with pytest.raises(pickle.UnpicklingError):
        DataOnlyUnpickler(io.BytesIO(pickle.dumps(Evil()))).load()
    with pytest.raises(pickle.UnpicklingError):
        DataOnlyUnpickler(io.BytesIO(pickle.dumps(np.random.default_rng))).load()


@pytest.mark.skipif(not os.path.exists(REF_PKL), reason="needs /root/reference (build container only)")
def test_committed_forest_is_the_reference_model():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from extract_forest import pack, read_forest
    fr = read_forest(REF_PKL)
    assert fr["params"]["n_jobs"] is None and fr["params"]["n_estimators"] == 20     # estimator-order summation is the reference's
    with open(FOREST, "rb") as f:
        assert f.read() == pack(fr)

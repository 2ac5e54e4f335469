"""Extracts the grasp classifier's decision forest (ENV:248: pickle.load of kinova_description/gc_model.pkl, used by
_get_reward ENV:643-659) into assets/gc_forest.bin WITHOUT executing anything the pickle names.

The file is an sklearn 0.22 RandomForestClassifier.  A stock pickle.load would import and call whatever the stream
asks for; this reader is a pickle.Unpickler whose find_class knows exactly seven names (the seven the stream holds, see
`ALLOWED`): numpy's array / dtype / scalar reconstructors are the real ones (pure data), and the three sklearn classes map
to inert record types defined HERE, so no sklearn code (of any version) runs and any other global raises.

Blob layout (model-card container of tools/compile_model.py, magic KGMC0001):
  forest_meta  i4[4]   n_trees, n_features, n_classes, n_nodes_total
  tree_root    i4[T+1] first node of tree t (nodes of a tree are contiguous, children indices are tree-relative as in sklearn)
  node_feature i4[n]   feature tested at the node (-2 at a leaf, sklearn's TREE_UNDEFINED)
  node_left    i4[n]   tree-relative index of the `<=` child, -1 at a leaf (TREE_LEAF)
  node_right   i4[n]
  node_thresh  f8[n]   threshold (sklearn compares float32(x) <= float64 threshold)
  node_proba   f8[n*C] value[node, 0, :] / sum (what DecisionTreeClassifier.predict_proba returns for a sample in that node)
  classes      f8[C]   RandomForestClassifier.classes_
"""
import io
import os
import pickle
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/gym-kinova-gripper/gym_kinova_gripper/envs/kinova_description/gc_model.pkl"
DST = os.path.join(ROOT, "kinovagrasping_b200", "assets", "gc_forest.bin")


class _Record:
    """Inert stand-in for an sklearn estimator: keeps constructor arguments and pickled state, runs nothing."""
    def __init__(self, *args):
        self.args = args
        self.state = None

    def __setstate__(self, state):
        self.state = state


class RandomForestClassifier(_Record):
    pass


class DecisionTreeClassifier(_Record):
    pass


class Tree(_Record):
    pass


def _np(name):
    import numpy.core.multiarray as ma  # noqa: F401  (numpy >= 2 keeps the alias)
    return {"_reconstruct": ma._reconstruct, "scalar": ma.scalar}[name]


ALLOWED = {
    ("sklearn.ensemble._forest", "RandomForestClassifier"): lambda: RandomForestClassifier,
    ("sklearn.tree._classes", "DecisionTreeClassifier"): lambda: DecisionTreeClassifier,
    ("sklearn.tree._tree", "Tree"): lambda: Tree,
    ("numpy.core.multiarray", "_reconstruct"): lambda: _np("_reconstruct"),
    ("numpy.core.multiarray", "scalar"): lambda: _np("scalar"),
    ("numpy", "ndarray"): lambda: np.ndarray,
    ("numpy", "dtype"): lambda: np.dtype,
}


class DataOnlyUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        try:
            return ALLOWED[(module, name)]()
        except KeyError:
            raise pickle.UnpicklingError(f"refused global {module}.{name}") from None


def read_forest(path=SRC):
    with open(path, "rb") as f:
        top = DataOnlyUnpickler(io.BytesIO(f.read())).load()
    if not isinstance(top, RandomForestClassifier):
        raise ValueError("not a RandomForestClassifier record")
    st = top.state
    trees = []
    for est in st["estimators_"]:
        t = est.state["tree_"]
        nodes, values = t.state["nodes"], t.state["values"]
        trees.append(dict(left=np.asarray(nodes["left_child"], dtype=np.int64), right=np.asarray(nodes["right_child"], dtype=np.int64),
                          feature=np.asarray(nodes["feature"], dtype=np.int64), threshold=np.asarray(nodes["threshold"], dtype=np.float64),
                          value=np.asarray(values, dtype=np.float64)))
    meta = {k: v for k, v in st.items() if k not in ("estimators_", "base_estimator_", "base_estimator")}
    return dict(trees=trees, classes=np.asarray(st["classes_"], dtype=np.float64), n_features=int(st["n_features_"]),
                n_outputs=int(st["n_outputs_"]), params=meta)


def pack(fr):
    T = len(fr["trees"])
    C = len(fr["classes"])
    if fr["n_outputs"] != 1:
        raise ValueError("single-output forests only")
    root = np.zeros(T + 1, dtype=np.int32)
    feat, left, right, thr, proba = [], [], [], [], []
    for t, tr in enumerate(fr["trees"]):
        n = len(tr["left"])
        root[t + 1] = root[t] + n
        v = tr["value"].reshape(n, -1)[:, :C]
        norm = v.sum(axis=1)                    # DecisionTreeClassifier.predict_proba: value / value.sum(), 0 -> 1
        norm[norm == 0.0] = 1.0
        proba.append(v / norm[:, None])
        feat.append(tr["feature"]); left.append(tr["left"]); right.append(tr["right"]); thr.append(tr["threshold"])
    secs = [("forest_meta", np.array([T, fr["n_features"], C, root[-1]], dtype=np.int32)), ("tree_root", root),
            ("node_feature", np.concatenate(feat).astype(np.int32)), ("node_left", np.concatenate(left).astype(np.int32)),
            ("node_right", np.concatenate(right).astype(np.int32)), ("node_thresh", np.concatenate(thr).astype(np.float64)),
            ("node_proba", np.concatenate(proba).astype(np.float64).ravel()), ("classes", fr["classes"])]
    code = {np.dtype("<f8"): 0, np.dtype("<i4"): 1, np.dtype("<f4"): 2}
    head = 16 + 40 * len(secs)
    off = (head + 63) // 64 * 64
    table, body = b"", b""
    for name, a in secs:
        a = np.ascontiguousarray(a)
        table += struct.pack("<24sIIQ", name.encode(), code[a.dtype], a.size, off + len(body))
        raw = a.tobytes()
        body += raw + b"\0" * ((-len(raw)) % 64)
    return struct.pack("<8sII", b"KGMC0001", len(secs), 0) + table + b"\0" * (off - head) + body


if __name__ == "__main__":
    fr = read_forest(sys.argv[1] if len(sys.argv) > 1 else SRC)
    print("trees", len(fr["trees"]), "features", fr["n_features"], "classes", fr["classes"], "nodes", sum(len(t["left"]) for t in fr["trees"]),
          "max depth-ish params", {k: fr["params"][k] for k in ("n_estimators", "n_jobs", "max_depth", "_sklearn_version") if k in fr["params"]})
    blob = pack(fr)
    with open(DST, "wb") as f:
        f.write(blob)
    print(DST, len(blob), "bytes")

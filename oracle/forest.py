"""TEST INFRASTRUCTURE ONLY (oracle): numpy restatement of the grasp classifier the reference calls in _get_reward
(ENV:248 loads kinova_description/gc_model.pkl; ENV:651 `self.Grasp_net.predict(local_obs[0:75].reshape(1,-1))`;
ENV:655-659 `outputs >= 0.3 and not self.Grasp_Reward` -> +5.0 once per episode, flag cleared by reset ENV:1372).

The model is an sklearn 0.22.2 RandomForestClassifier (20 trees, 75 features, classes [0., 1.]).  scikit-learn is a
third-party dependency that is not vendored in /root/reference (README.md installs it through pip); its published
algorithm, restated here:
  * sklearn/ensemble/_forest.py ForestClassifier.predict: classes_.take(argmax(predict_proba(X), axis=1))
  * ForestClassifier.predict_proba: X -> float32; all_proba = sum over estimators_ (in list order for n_jobs=None) of
    tree.predict_proba(X); all_proba /= n_estimators
  * sklearn/tree/_classes.py DecisionTreeClassifier.predict_proba: value[leaf, 0, :] / value[leaf, 0, :].sum()
  * sklearn/tree/_tree.pyx Tree.apply_dense: from node 0, while left_child != TREE_LEAF (-1):
        node = left_child if float32(X[feature]) <= float64(threshold) else right_child
Pinned by tests/test_forest.py against tests/golden/forest.npz, which tools/gen_golden_forest.py writes by running the
INSTALLED scikit-learn's own compiled Tree.predict on the same node arrays (tools/extract_forest.py reads them from the
reference's file with a data-only unpickler)."""
import struct

import numpy as np

_DT = {0: "<f8", 1: "<i4", 2: "<f4"}


class Forest:
    def __init__(self, path):
        raw = open(path, "rb").read()
        magic, n, _ = struct.unpack_from("<8sII", raw, 0)
        assert magic == b"KGMC0001"
        sec = {}
        for i in range(n):
            name, dt, cnt, off = struct.unpack_from("<24sIIQ", raw, 16 + 40 * i)
            sec[name.rstrip(b"\0").decode()] = np.frombuffer(raw, dtype=_DT[dt], count=cnt, offset=off)
        self.n_trees, self.n_features, self.n_classes, self.n_nodes = (int(v) for v in sec["forest_meta"])
        self.root, self.feature, self.left, self.right = sec["tree_root"], sec["node_feature"], sec["node_left"], sec["node_right"]
        self.thresh, self.proba, self.classes = sec["node_thresh"], sec["node_proba"].reshape(-1, self.n_classes), sec["classes"]

    def leaf(self, t, x32):
        base, node = int(self.root[t]), 0
        while self.left[base + node] != -1:
            i = base + node
            node = int(self.left[i]) if np.float64(x32[self.feature[i]]) <= self.thresh[i] else int(self.right[i])
        return base + node

    def predict_proba(self, X):
        X = np.atleast_2d(np.asarray(X)).astype(np.float32)
        out = np.zeros((X.shape[0], self.n_classes))
        for r, x in enumerate(X):
            for t in range(self.n_trees):
                out[r] += self.proba[self.leaf(t, x)]
        return out / self.n_trees

    def predict(self, X):
        return self.classes.take(np.argmax(self.predict_proba(X), axis=1))


def grasp_reward(forest, local_obs, already_given):
    """ENV:651-659 -> (grasp_reward, Grasp_Reward flag after the call)."""
    out = forest.predict(np.asarray(local_obs[0:75]).reshape(1, -1))
    if bool(out >= 0.3) and not already_given:
        return 5.0, True
    return 0.0, already_given

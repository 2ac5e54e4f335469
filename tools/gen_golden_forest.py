"""Writes tests/golden/forest.npz: inputs and the outputs scikit-learn's OWN compiled code gives for the reference's
grasp-classifier forest (ENV:248, 651).  Run in the build container (needs /root/reference and scikit-learn):

    python tools/gen_golden_forest.py

The node arrays come out of gc_model.pkl through tools/extract_forest.py's data-only unpickler (the file is sklearn
0.22.2; nothing it names is executed).  They are loaded into the INSTALLED scikit-learn's sklearn.tree._tree.Tree through
Tree.__setstate__ (its node record gained a `missing_go_to_left` field since 0.22; set to 0, no input is NaN), and
Tree.apply -- sklearn's Cython traversal -- picks the leaves.  The class probabilities are the 0.22 formulas on the
pickled `value` array (DecisionTreeClassifier.predict_proba / ForestClassifier.predict_proba / predict, cited in
oracle/forest.py); a RandomForestClassifier shell of the installed version with these trees as estimators_ is run as a
second opinion on predict().

Inputs: observation vectors of the reference's own _get_obs (tests/golden/obs_golden.npz, expert_golden.npz), Gaussian
vectors, and vectors placed ON the split thresholds of the nodes a sample visits (float32 neighbours on either side of
each float64 threshold), which is where a traversal can go wrong."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from extract_forest import read_forest  # noqa: E402


def sk_trees(fr, normalised=False):
    """normalised: leaf values as class fractions -- the form the installed version (>= 1.3) stores and averages directly."""
    from sklearn.tree._tree import NODE_DTYPE, Tree
    out = []
    for tr in fr["trees"]:
        n = len(tr["left"])
        val = tr["value"].reshape(n, 1, -1).copy()
        if normalised:
            s = val.sum(axis=2, keepdims=True); s[s == 0.0] = 1.0
            val = val / s
        nodes = np.zeros(n, dtype=NODE_DTYPE)
        nodes["left_child"], nodes["right_child"], nodes["feature"], nodes["threshold"] = tr["left"], tr["right"], tr["feature"], tr["threshold"]
        nodes["n_node_samples"] = 1
        nodes["weighted_n_node_samples"] = 1.0
        t = Tree(fr["n_features"], np.array([len(fr["classes"])], dtype=np.intp), 1)
        depth = np.zeros(n, dtype=np.int64)
        for i in range(n):
            if tr["left"][i] != -1:
                depth[tr["left"][i]] = depth[tr["right"][i]] = depth[i] + 1
        t.__setstate__({"max_depth": int(depth.max()), "node_count": n, "nodes": nodes, "values": np.ascontiguousarray(val)})
        out.append(t)
    return out


def main():
    fr = read_forest()
    trees = sk_trees(fr)
    C = len(fr["classes"])
    rng = np.random.default_rng(20201018)
    og = np.load(os.path.join(ROOT, "tests", "golden", "obs_golden.npz"))["obs_local"][:, :75]
    eg = np.load(os.path.join(ROOT, "tests", "golden", "expert_golden.npz"))["obs"].reshape(-1, 82)[:, :75]
    eg = eg[np.abs(eg).sum(axis=1) > 0]
    real = np.concatenate([og, eg[rng.choice(len(eg), 1500, replace=False)]])
    gauss = rng.normal(0.0, 0.08, size=(600, 75))
    # on-threshold cases: walk a real sample down a tree, and at every node visited emit the sample with that node's feature
    # set to the float32 numbers just below / at / just above the float64 threshold
    edge = []
    for k in range(60):
        x = real[rng.integers(len(real))].astype(np.float32)
        tr = fr["trees"][k % len(trees)]
        node = 0
        while tr["left"][node] != -1 and len(edge) < 6000:
            f, th = int(tr["feature"][node]), tr["threshold"][node]
            c = np.float32(th)
            for v in (np.nextafter(c, np.float32(-np.inf)), c, np.nextafter(c, np.float32(np.inf))):
                y = x.copy(); y[f] = v
                edge.append(y)
            node = int(tr["left"][node]) if np.float64(x[f]) <= th else int(tr["right"][node])
    X = np.concatenate([real, gauss, np.asarray(edge, dtype=np.float64)]).astype(np.float32)
    leaves = np.stack([t.apply(X) for t in trees], axis=1)                     # sklearn's compiled traversal
    proba = np.zeros((len(X), C))
    for t, tr in enumerate(fr["trees"]):                                        # ForestClassifier.predict_proba, estimator order
        v = tr["value"].reshape(len(tr["left"]), -1)[leaves[:, t]]
        norm = v.sum(axis=1); norm[norm == 0.0] = 1.0
        proba += v / norm[:, None]
    proba /= len(trees)
    label = fr["classes"].take(np.argmax(proba, axis=1))
    # second opinion: the installed RandomForestClassifier.predict over DecisionTreeClassifier shells holding these trees
    from sklearn.ensemble import RandomForestClassifier
    from sklearn.tree import DecisionTreeClassifier
    rf = RandomForestClassifier(n_estimators=len(trees), n_jobs=1)
    ests = []
    for t in sk_trees(fr, normalised=True):
        d = DecisionTreeClassifier()
        d.tree_, d.n_outputs_, d.n_classes_, d.classes_, d.n_features_in_, d.max_features_ = t, 1, C, fr["classes"], fr["n_features"], fr["n_features"]
        ests.append(d)
    rf.estimators_, rf.n_outputs_, rf.n_classes_, rf.classes_, rf.n_features_in_ = ests, 1, C, fr["classes"], fr["n_features"]
    label2 = rf.predict(X)
    margin = np.abs(proba[:, 0] - proba[:, 1])
    agree = label2 == label
    print("samples", len(X), "positive", int((label >= 0.3).sum()), "ties", int((margin == 0).sum()), "installed-RF predict agrees on", int(agree.sum()),
          "disagreements all at margin <", float(margin[~agree].max()) if (~agree).any() else 0.0)
    # the installed version averages NORMALISED leaf values exactly as 0.22 does; any disagreement could only be a rounding tie
    assert (margin[~agree] < 1e-12).all()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "forest.npz"), X=X, leaves=leaves.astype(np.int32), proba=proba, label=label,
                        n_real=len(real), n_gauss=len(gauss), sklearn_version=np.array(__import__("sklearn").__version__))
    print("wrote tests/golden/forest.npz")


if __name__ == "__main__":
    main()

// Grasp-classifier reward (ENV:248, 643-659): the reference unpickles an sklearn RandomForestClassifier
// (kinova_description/gc_model.pkl: 20 trees, 75 features, classes {0, 1}) and, when with_grasp_reward is on, adds 5.0 to
// the reward of the first step of an episode whose local observation [0:75] the forest labels >= 0.3.
// Here the forest is a device table (assets/gc_forest.bin, written by tools/extract_forest.py) and the reward is two small
// kernels behind the last per-env kernel of an RL step -- they read the obs that kernel just wrote; the hot kernels are
// untouched:
//   kg_forest_leaf:   a thread per (tree, env) walks one tree (sklearn/tree/_tree.pyx apply_dense: float32 feature <=
//                     float64 threshold -> left child, else right; a leaf has left == -1)
//   kg_grasp_decide:  a thread per env adds the leaves' class fractions in estimator order in fp64 (ForestClassifier.
//                     predict_proba with n_jobs=None), takes the first arg-max (numpy argmax: class 0 on a 10:10 tie), maps it
//                     through classes_, applies `>= 0.3 and not Grasp_Reward`, adds 5.0 and sets the per-episode flag.
// Discrete output: the label is bit-exact against scikit-learn's own traversal (tests/golden/forest.npz).
#pragma once
struct KgForest {
    int n_trees, n_features, n_classes, n_nodes;
    const int *root, *feature, *left, *right;     // root[T+1]; children are tree-relative
    const double *thresh, *proba, *classes;        // proba[n_nodes][n_classes]
};
#define KG_GRASP_REWARD 5.0f
#define KG_GRASP_LABEL_MIN 0.3

// obs: [82][N] field-major (KgEnvArrays::obs); returns the global node index of the leaf
KG_DEV int kg_forest_leaf(const KgForest &F, int tree, const float *obs, int N, int env) {
    const int base = F.root[tree];
    int node = 0;
    for (;;) {
        const int i = base + node;
        const int l = F.left[i];
        if (l < 0) return i;
        const double x = (double)obs[(size_t)F.feature[i] * N + env];
        node = (x <= F.thresh[i]) ? l : F.right[i];
    }
}
// RandomForestClassifier.predict of one sample from its leaves, leaf: [n_trees][N]
KG_DEV double kg_forest_label(const KgForest &F, const int *leaf, int N, int env) {
    int best = 0;
    double pbest = 0.0;
    for (int c = 0; c < F.n_classes; c++) {
        double p = 0.0;
        for (int t = 0; t < F.n_trees; t++) p += F.proba[(size_t)leaf[(size_t)t * N + env] * F.n_classes + c];
        p /= (double)F.n_trees;
        if (c == 0 || p > pbest) { pbest = p; best = c; }
    }
    return F.classes[best];
}
// given: per-episode flag (ENV:103, cleared by reset ENV:1372).  reward may be NULL.
KG_DEV void kg_grasp_decide(const KgForest &F, const int *leaf, int N, int env, int *given, float *reward, float *grasp_out) {
    float g = 0.f;
    if (kg_forest_label(F, leaf, N, env) >= KG_GRASP_LABEL_MIN && !given[env]) { g = KG_GRASP_REWARD; given[env] = 1; }
    grasp_out[env] = g;
    if (reward) reward[env] += g;        // reward = 0.2 * finger_reward (0) + lift_reward + grasp_reward, ENV:682
}
#if !defined(KG_EMU)
__global__ void __launch_bounds__(256) kg_forest_leaf_kernel(KgForest F, const float *obs, int N, int *leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // consecutive threads = consecutive envs of one tree
    if (i >= F.n_trees * N) return;
    const int tree = i / N, env = i - tree * N;
    leaf[i] = kg_forest_leaf(F, tree, obs, N, env);
}
__global__ void __launch_bounds__(256) kg_grasp_decide_kernel(KgForest F, const int *leaf, int N, int *given, float *reward, float *grasp_out) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env < N) kg_grasp_decide(F, leaf, N, env, given, reward, grasp_out);
}
__global__ void __launch_bounds__(256) kg_forest_label_kernel(KgForest F, const int *leaf, int N, float *label) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env < N) label[env] = (float)kg_forest_label(F, leaf, N, env);
}
#endif

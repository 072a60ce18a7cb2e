"""Loader for tests/golden/*.npz (outputs of the UNMODIFIED reference, see
tests/golden/make_golden.py)."""
import glob
import os
import types

import numpy as np

from oracle import qrf_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.splitext(os.path.basename(p))[0]
               for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class ArrayTree:
    """Duck-typed stand-in for sklearn.tree._tree.Tree built from stored arrays."""

    def __init__(self, left, right, feature, threshold, value):
        self.children_left, self.children_right = left, right
        self.feature, self.threshold = feature, threshold
        self.value = value.reshape(-1, 1, 1)
        self.node_count = len(left)

    def apply(self, X):                       # the oracle's restatement of _apply_dense
        return O.tree_apply(self, X)


class ArrayEstimator:
    def __init__(self, tree, seed):
        self.tree_ = tree
        self.random_state = int(seed)


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    off = z["node_offsets"]
    ests = []
    for t in range(len(off) - 1):
        s = slice(off[t], off[t + 1])
        ests.append(ArrayEstimator(ArrayTree(z["children_left"][s], z["children_right"][s],
                                             z["feature"][s], z["threshold"][s], z["value"][s]),
                                   z["seeds"][t]))
    g = types.SimpleNamespace(**{k: z[k] for k in z.files})
    g.name = name
    g.estimators = ests
    g.bootstrap = bool(z["bootstrap"])
    g.kind = str(z["kind"])
    return g


def flat_from_golden(g):
    """Flatten the stored trees with the product flattener (host side, no GPU)."""
    import skgarden_b200 as sg
    N = len(g.y_train)
    trees = [e.tree_ for e in g.estimators]
    train_leaves = [t.apply(g.X_train) for t in trees]
    counts = [sg.bootstrap_counts(e.random_state, N) for e in g.estimators] if g.bootstrap else None
    return sg.flatten_forest(trees, train_leaves, counts, g.y_train, g.X_train.shape[1],
                             sorter=g.sorter)

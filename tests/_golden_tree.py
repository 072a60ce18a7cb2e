"""Loader for tests/golden_tree/*.npz (outputs of the UNMODIFIED reference's tree-level
estimators, see tests/golden_tree/make_golden_tree.py)."""
import glob
import os
import types

import numpy as np

from _golden import ArrayTree

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_tree")
CASES = sorted(os.path.splitext(os.path.basename(p))[0]
               for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = types.SimpleNamespace(**{k: z[k] for k in z.files})
    g.name = name
    g.tree = ArrayTree(z["children_left"], z["children_right"], z["feature"], z["threshold"], z["value"])
    return g


def flat_from_golden(g):
    """The stored tree through the product flattener: T = 1, unit weights (host side, no GPU)."""
    import skgarden_b200 as sg
    return sg.flatten_forest([g.tree], [g.y_train_leaves_], None, g.y_train, g.X_train.shape[1],
                             sorter=np.argsort(g.y_train), unit_weights=True)

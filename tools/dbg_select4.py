"""Bisect helper: one golden case through select4 with the given options (separate process per variant)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import _golden
from skgarden_b200 import device_forest as DF

name = sys.argv[1]
opts = dict(kv.split("=") for kv in sys.argv[2:])
g = _golden.load(name)
flat = _golden.flat_from_golden(g)
print(name, "T", flat.n_trees if hasattr(flat, "n_trees") else "?", "N", len(g.y_train), "rows", len(g.X_test), "opts", opts, flush=True)
dev = DF.DeviceForest(flat, device=0)
dev.set_option("warp_elems", 0)
dev.set_option("select_impl", 4)
for k, v in opts.items():
    dev.set_option(k, int(v))
X = torch.as_tensor(np.ascontiguousarray(g.X_test, dtype=np.float32)).cuda()
qs = list(g.q[:3])
print("kernel", dev.quantile_kernel_name(len(g.X_test), 3, 1), flush=True)
for rows in (1, 8, 9, len(g.X_test)):
    out = dev.predict_quantiles(X[:rows], qs, mode=1)
    torch.cuda.synchronize()
    print("rows", rows, "ok", out.cpu().numpy()[:2].tolist(), flush=True)

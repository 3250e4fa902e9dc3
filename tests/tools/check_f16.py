#!/usr/bin/env python
"""A/B check of the fp16-split tensor-core kernels (k_tower16, k_fc16) against the fp32 CUDA-core kernels, layer by
layer, and of every mode combination against the fp64 CPU oracle.  Usage: check_f16.py [n_positions]"""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("OMK_TOWER_STAMPS", "1")  # the instrumented instantiation of k_tower16 (phase timestamps)
omk = importlib.import_module("omok-ai_b200")
from oracle import net_oracle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
rng = np.random.default_rng(0)
boards = np.zeros((n, 81), np.uint8)
turns = np.zeros(n, np.uint8)
for b in range(n):
    k = int(rng.integers(0, 60))
    for j, c in enumerate(rng.permutation(81)[:k]):
        boards[b, c] = 1 + (j % 2)
    turns[b] = k % 2
ctx = omk.Context(device=0, capacity_envs=4, capacity_trees=4, capacity_nodes=64, seed=0)
params = net_oracle.random_params(0)
ctx.net_load_params(params)


def run(tower, fc):
    ctx.debug_set_tower_mode(tower)
    ctx.debug_set_fc0_mode(fc)
    return ctx.net_eval(boards, turns)


def report(name, a, ref):
    d = np.abs(a - ref)
    print(f"{name}: max|ref| {np.abs(ref).max():.4g}  max abs diff {d.max():.3g}  rel to max {d.max() / np.abs(ref).max():.3g}", flush=True)
    return d


# 1. the tower alone: fp16-split tower -> fp32 act0 -> CUDA-core fc
run(0, 0)
x0 = ctx.debug_get_buffer(0, n * 10368).reshape(n, 81, 128).astype(np.float64)
a1_0 = ctx.debug_get_buffer(1, n * 512).reshape(n, 512).astype(np.float64)
run(1, 0)
x2 = ctx.debug_get_buffer(0, n * 10368).reshape(n, 81, 128).astype(np.float64)
d = report("tower f16 vs simt (act0)", x2, x0)
bad = np.argwhere(d > 1e-4 * np.abs(x0).max())
print("  bad entries:", len(bad), bad[:6].tolist())
if len(bad):
    print("  bad by position%3:", np.bincount(bad[:, 0] % 3, minlength=3).tolist())
    print("  bad by pixel:", np.bincount(bad[:, 1], minlength=81).tolist())
    print("  bad by channel/16:", np.bincount(bad[:, 2] // 16, minlength=8).tolist())
# 2. fc0 alone: CUDA-core tower -> split -> fp16-split fc0/fc1
run(0, 1)
a1_2 = ctx.debug_get_buffer(9, n * 512).reshape(n, 512).astype(np.float64)
d = report("fc0 f16 vs simt (act1)", a1_2, a1_0)
bad = np.argwhere(d > 1e-4 * np.abs(a1_0).max())
print("  bad entries:", len(bad), bad[:6].tolist())
# 3. every combination against the fp64 oracle
m = min(n, 96)
import torch  # noqa: E402

rp, rv, _ = net_oracle.forward_boards(params, boards[:m], turns[:m], dtype=torch.float64)
for tower, fc in ((0, 0), (1, 0), (0, 1), (1, 1)):
    p, v = run(tower, fc)
    big = rp > 1e-12
    print(f"tower {tower} fc {fc}: max rel P vs fp64 {np.max(np.abs(p[:m][big] - rp[big]) / rp[big]):.3g}  "
          f"max rel V {np.max(np.abs(v[:m] - rv) / np.maximum(np.abs(rv), 1e-3)):.3g}", flush=True)
# phase timing of one iteration (clock64 stamps inside k_tower16)
run(1, 1)
ts = ctx.debug_tower_timing()
names = ["start", "stem"] + [f"b{r}:{nm}" for r in range(3) for nm in ("conv0", "E1+sync", "dw+st", "conv1", "E2+st", "conv2", "E3+st", "-")]
prev = ts[0]
for i in list(range(0, 2)) + [2 + r * 8 + k for r in range(3) for k in range(7)] + [30, 31]:
    nm = names[i] if i < len(names) else ("before store" if i == 30 else "stored")
    print(f"{nm:12s} +{int(ts[i] - prev):6d}  (t={int(ts[i] - ts[0])})")
    prev = ts[i]

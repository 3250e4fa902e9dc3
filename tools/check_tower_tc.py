#!/usr/bin/env python
"""A/B check of the tensor-core tower (k_tower_tc) against the fp32 CUDA-core tower and the fp64 CPU oracle."""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omk = importlib.import_module("omok-ai_b200")
from oracle import net_oracle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
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
ctx.debug_set_fc0_mode(1)
out = {}
for mode in (0, 1):
    ctx.debug_set_tower_mode(mode)
    t = time.time()
    p, v = ctx.net_eval(boards, turns)
    x = ctx.debug_get_buffer(4, n * 10368).astype(np.float64) + ctx.debug_get_buffer(5, n * 10368)
    out[mode] = (p, v, x.reshape(n, 81, 128), time.time() - t)
    print("tower mode", mode, "ok", round(out[mode][3], 4), flush=True)
x0, x1 = out[0][2], out[1][2]
print("tower out: max|simt|", np.abs(x0).max(), "max abs diff", np.abs(x0 - x1).max(), "rel to max", np.abs(x0 - x1).max() / np.abs(x0).max())
bad = np.argwhere(np.abs(x0 - x1) > 1e-3 * np.abs(x0).max())
print("bad entries:", len(bad), bad[:8].tolist())
if len(bad):
    print("bad by pixel:", np.bincount(bad[:, 1], minlength=81).tolist())
    print("bad by channel (first 32):", np.bincount(bad[:, 2], minlength=128)[:32].tolist())
rp, rv, _ = net_oracle.forward_boards(params, boards[:64], turns[:64], dtype=__import__("torch").float64)
for mode in (0, 1):
    p, v = out[mode][0][:64], out[mode][1][:64]
    big = rp > 1e-12
    print("tower mode", mode, "max rel P vs fp64", np.max(np.abs(p[big] - rp[big]) / rp[big]), "max rel V", np.max(np.abs(v - rv) / np.maximum(np.abs(rv), 1e-3)))

# phase timing of one position (clock64 stamps inside k_tower_tc)
ts = ctx.debug_tower_timing()
names = ["start", "stem"] + [f"b{r}:{n}" for r in range(3) for n in ("conv0", "E1+sync", "dw+st", "conv1", "E2+st", "conv2", "E3+st", "-")]
prev = ts[0]
for i in list(range(0, 2)) + [2 + r * 8 + k for r in range(3) for k in range(7)] + [30]:
    nm = names[i] if i < len(names) else "before store"
    print(f"{nm:12s} +{int(ts[i] - prev):6d}  (t={int(ts[i] - ts[0])})")
    prev = ts[i]
print("stem detail: img+sync", int(ts[40]-ts[0]), "math", int(ts[41]-ts[40]), "split+st issue", int(ts[42]-ts[41]), "wait::st", int(ts[1]-ts[42]))
print("store phase", int(ts[31] - ts[30]))

#!/usr/bin/env python
"""A/B check of the tensor-core fc0 (tcgen05 3xTF32) against the fp32 CUDA-core GEMM and the fp64 CPU oracle."""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omk = importlib.import_module("omok-ai_b200")
from oracle import net_oracle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
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
out = {}
for mode in (0, 1):
    ctx.debug_set_fc0_mode(mode)
    t = time.time()
    p, v = ctx.net_eval(boards, turns)
    a1 = ctx.debug_get_buffer(1, n * 512) if mode == 0 else ctx.debug_get_buffer(6, n * 512).astype(np.float64) + ctx.debug_get_buffer(7, n * 512)
    out[mode] = (p, v, np.asarray(a1, np.float64).reshape(n, 512), time.time() - t, ctx.debug_get_buffer(2, n * 512).reshape(n, 512))
    print("mode", mode, "ok", out[mode][3], flush=True)
a0, a1 = out[0][2], out[1][2]
print("fc0 out: max|simt|", np.abs(a0).max(), "max abs diff", np.abs(a0 - a1).max(), "rel to max", np.abs(a0 - a1).max() / np.abs(a0).max())
d = (a1.astype(np.float64) - a0) * np.sign(a0)
print("signed diff toward larger magnitude: mean", d.mean(), "median", np.median(d), "frac negative", (d < 0).mean(), "rms", np.sqrt((d**2).mean()), "typical |a0|", np.median(np.abs(a0)))
print("fc1 out: max abs diff", np.abs(out[0][4] - out[1][4]).max(), "max|.|", np.abs(out[0][4]).max())
bad = np.argwhere(np.abs(a0 - a1) > 1e-3 * np.abs(a0).max())
print("bad entries:", len(bad), bad[:10].tolist())
rp, rv, _ = net_oracle.forward_boards(params, boards[:64], turns[:64], dtype=__import__("torch").float64)
for mode in (0, 1):
    p, v = out[mode][0][:64], out[mode][1][:64]
    big = rp > 1e-12
    print("mode", mode, "max rel P vs fp64", np.max(np.abs(p[big] - rp[big]) / rp[big]), "max rel V", np.max(np.abs(v - rv) / np.maximum(np.abs(rv), 1e-3)))

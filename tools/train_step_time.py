#!/usr/bin/env python
"""tools/train_step_time.py -- wall-clock of omk_train_step (forward, losses, backward, Adadelta, operand re-pack, the
reference's second forward; alpha-zero/src/agent_model.rs:136-168) on one GPU for a minibatch of N positions.

  python tools/train_step_time.py [--n 128] [--steps 20] [--warm 3]

Prints one JSON line.  Under `ncu --metrics gpu__time_duration.sum` (with --steps 2 --warm 1) the launch list gives the
per-kernel split of a step (tools/ncu_launch_shares.py).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warm", type=int, default=3)
    args = ap.parse_args()
    omk = importlib.import_module("omok-ai_b200")
    ctx = omk.Context(device=0, capacity_envs=4, capacity_trees=2, capacity_nodes=64, seed=1)
    ctx.net_init_random(0)
    rng = np.random.default_rng(0)
    n = args.n
    images = (rng.random((n, 243)) < 0.3).astype(np.float32)
    pi = rng.random((n, 81)).astype(np.float32)
    pi /= pi.sum(axis=1, keepdims=True)
    z = rng.choice([-1.0, 0.0, 1.0], size=n).astype(np.float32)
    losses = []
    for _ in range(args.warm):
        losses.append(ctx.train_step(images, pi, z))
    ctx.synchronize()
    l0 = ctx.launch_count
    t0 = time.perf_counter()
    for _ in range(args.steps):
        losses.append(ctx.train_step(images, pi, z))
    ctx.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"what": "omk_train_step wall-clock, one GPU", "positions": n, "steps": args.steps,
                      "ms_per_step": 1e3 * dt / args.steps, "launches_per_step": (ctx.launch_count - l0) / args.steps,
                      "loss_first": losses[0][2], "loss_last": losses[-1][2]}))
    ctx.close()


if __name__ == "__main__":
    main()

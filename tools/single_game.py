#!/usr/bin/env python
"""BASELINE configs[0]: ONE self-play game on one tree pair, 800 simulations per move in rounds of 8, epsilon 0, alpha 1,
Best move (benchmark/src/main.rs:9-10, benchmark/src/agent.rs:14-47) -- the reference's own CPU-runnable case, here as a
latency measurement of the GPU path (network calls of 8 rows).  Usage: single_game.py [max_plies]"""
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omk = importlib.import_module("omok-ai_b200")
max_plies = int(sys.argv[1]) if len(sys.argv) > 1 else 81
ctx = omk.Context(device=0, capacity_envs=4, capacity_trees=2, capacity_nodes=4096, seed=0)
ctx.net_init_random(0)
ctx.pool_new_games(n=2, evaluator=omk.EVAL_NET)
moves, t_search = [], []
status, ply = 0, 0
while status == 0 and ply < max_plies:
    mover, other = (0, 1) if ply % 2 == 0 else (1, 0)
    t0 = time.perf_counter()
    ctx.pool_search(ids=[mover], count=800, batch_size=8, epsilon=0.0, alpha=1.0, evaluator=omk.EVAL_NET)
    t_search.append(time.perf_counter() - t0)
    act, _ = ctx.pool_sample(ids=[mover], modes=[omk.SAMPLE_BEST])
    status = int(ctx.pool_play(act, ids=[mover])[0])
    ctx.pool_ensure_action(act, ids=[other], evaluator=omk.EVAL_NET)
    ctx.pool_play(act, ids=[other])
    moves.append(int(act[0]))
    ply += 1
ms = 1e3 * np.array(t_search[1:] if len(t_search) > 1 else t_search)
print(json.dumps({"plies": ply, "status": status, "ms_per_move_median": float(np.median(ms)), "ms_per_move_max": float(ms.max()),
                  "sims_per_s": 800 / (float(np.median(ms)) * 1e-3), "moves": moves[:12]}))

#!/usr/bin/env python
"""One short self-play run of a tree pool with the hash evaluator (the ncu target for the K2 tree kernels).
Usage: tree_pool_run.py [games] [capacity_nodes] [sims_per_move]"""
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omk = importlib.import_module("omok-ai_b200")
games = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 768
count = int(sys.argv[3]) if len(sys.argv) > 3 else 160
ctx = omk.Context(device=0, capacity_envs=4, capacity_trees=2 * games, capacity_nodes=cap, seed=1)
ctx.debug_set_lane_min_trees(0)
ctx.selfplay_begin(games, count, 16, 0.25, 0.03, 1.0, 30, omk.EVAL_HASH)
stats, *_ = ctx.selfplay_run(2, profile=2, want_transitions=False)
print(json.dumps({k: v for k, v in stats.by_kind().items()}))

#!/usr/bin/env python
"""Smallest run that touches every kernel of the self-play path (target of `compute-sanitizer --tool memcheck`)."""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omk = importlib.import_module("omok-ai_b200")
ctx = omk.Context(device=0, capacity_envs=8, capacity_trees=8, capacity_nodes=256, seed=0)
ctx.net_init_random(0)
rng = np.random.default_rng(0)
boards = ((rng.random((7, 81)) < 0.3) * rng.integers(1, 3, size=(7, 81))).astype(np.uint8)
p, v = ctx.net_eval(boards, np.zeros(7, np.uint8))
ctx.env_reset(n=8)
ctx.env_step(np.arange(8, dtype=np.uint8))
ctx.selfplay_begin(4, 32, 16, 0.25, 0.03, 1.0, 30, omk.EVAL_NET)
stats, *_ = ctx.selfplay_run(3, profile=0, want_transitions=True)
# round 2 additions: batched tree-env read, the trainer step (forward / backward / Adadelta / re-pack), the opt-in search mode
ctx.pool_get_envs(ids=[0, 3, 5])
imgs = rng.random((5, 243)).astype(np.float32)
pi = rng.random((5, 81)).astype(np.float32)
pi /= pi.sum(1, keepdims=True)
losses = ctx.train_step(imgs, pi, rng.choice(np.array([-1.0, 1.0], np.float32), 5))
ctx.debug_set_fc0_chunk(3)
ctx.net_eval_images(imgs)
ctx.search_set_virtual_loss(True)
ctx.pool_new_games(n=4, evaluator=omk.EVAL_NET)
ctx.pool_search(n=4, count=32, batch_size=16, epsilon=0.25, alpha=0.03, evaluator=omk.EVAL_NET)
print("ok", float(p.sum()), int(stats.simulations), losses)
ctx.close()

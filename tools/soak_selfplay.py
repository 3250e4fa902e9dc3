#!/usr/bin/env python
"""Soak run of the self-play driver: many plies on a full pool (games finish and restart, trees re-root, both search lanes
busy), then consistency checks on the streamed transitions.  Usage: soak_selfplay.py [games] [plies] [sims_per_move]"""
import importlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omk = importlib.import_module("omok-ai_b200")
trainer = importlib.import_module("omok-ai_b200.trainer")
games = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
plies = int(sys.argv[2]) if len(sys.argv) > 2 else 120
count = int(sys.argv[3]) if len(sys.argv) > 3 else 160
ctx = omk.Context(device=0, capacity_envs=4, capacity_trees=2 * games, capacity_nodes=4096, seed=9)
ctx.net_init_random(0)
ctx.selfplay_begin(games, count, 16, 0.25, 0.03, 1.0, 30, omk.EVAL_NET)
carry, finished, transitions, sims = None, 0, 0, 0
lengths = []
for chunk in range(plies // 20):
    stats, boards, policy, status, actions = ctx.selfplay_run(20, profile=0, want_transitions=True)
    sims += int(stats.simulations)
    # every recorded move is legal on the recorded board, every policy is a distribution over empty cells
    occupied = np.take_along_axis(boards, actions[..., None].astype(np.int64), axis=2)[..., 0]
    assert (occupied == 0).all(), "an illegal move was played"
    assert np.allclose(policy.sum(2), 1.0, atol=1e-4)
    assert (policy[boards != 0] == 0).all(), "visit policy puts mass on an occupied cell"
    assert ((boards == 1).sum(2) - (boards == 2).sum(2) <= 1).all() and ((boards == 1).sum(2) >= (boards == 2).sum(2)).all()
    episodes, carry = trainer.split_episodes(boards, policy, status, carry)
    finished += len(episodes)
    assert finished >= int(stats.games_finished) > 0 or chunk == 0
    for b, p, z in episodes:
        lengths.append(len(b))
        assert (b[0] == 0).all(), "a finished game's first recorded board must be empty"
        assert 9 <= len(b) <= 81
    transitions += boards.shape[0] * boards.shape[1]
print(json.dumps({"games": games, "plies": plies, "sims_per_move": count, "simulations": sims, "expected_simulations": games * plies * count,
                  "episodes_finished": finished, "mean_game_length": float(np.mean(lengths)) if lengths else None,
                  "transitions": transitions}))
assert sims == games * plies * count
ctx.close()

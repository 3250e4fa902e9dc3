#!/usr/bin/env python
"""BASELINE configs[0] on the CPU: the oracle port (C search + PyTorch-CPU fp32 network standing in for TensorFlow-CPU)
plays one game at 800 simulations per move in rounds of 8, epsilon 0, alpha 1, Best move (benchmark/src/main.rs:9-10,
benchmark/src/agent.rs:14-47).  Companion of tools/single_game.py (the GPU path).  Usage: single_game_cpu.py [max_plies]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import net_oracle, oracle as orc  # noqa: E402

max_plies = int(sys.argv[1]) if len(sys.argv) > 1 else 6
torch.set_num_threads(os.cpu_count() or 1)
ev = orc.TorchEvaluator(net_oracle.random_params(0))
agents = [orc.Agent(ev, 0, 0), orc.Agent(ev, 0, 1)]
t_move, status, ply = [], 0, 0
while status == 0 and ply < max_plies:
    mover, other = agents[ply % 2], agents[1 - ply % 2]
    t0 = time.perf_counter()
    orc.execute([mover], 800, 8, 0.0, 1.0, ev)
    t_move.append(time.perf_counter() - t0)
    act, _ = mover.sample_action(0, 1.0)
    status = mover.play_action(act)
    other.ensure_action_exists(act, ev)
    other.play_action(act)
    ply += 1
ms = 1e3 * np.array(t_move)
print(json.dumps({"plies": ply, "cores": os.cpu_count(), "ms_per_move_median": float(np.median(ms)),
                  "sims_per_s": 800 / (float(np.median(ms)) * 1e-3), "kind": "port (C oracle + PyTorch-CPU fp32 network)"}))
del mover, other
agents.clear()  # free the oracle agents before the interpreter tears the library down

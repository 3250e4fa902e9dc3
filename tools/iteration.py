#!/usr/bin/env python
"""One AlphaZero iteration at BASELINE config 5 scale-down: sharded self-play on every GPU (no collective), then the
trainer's gradient steps with ONE NCCL all-reduce of the 22.6 MB gradient per step, then the weight refresh of the
self-play kernels.  Single GPU: `python tools/iteration.py`; N GPUs of one node:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/iteration.py
Prints one JSON line (rank 0): self-play throughput, trainer step time (device timed, max over ranks) and losses."""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(games=256, plies=24, count=160, batch=16, steps=20, minibatch=128, seed=0, quiet=False):
    omk = importlib.import_module("omok-ai_b200")
    trainer = importlib.import_module("omok-ai_b200.trainer")
    sharding = importlib.import_module("omok-ai_b200.sharding")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _, my_games = sharding.shard_games(games, world, rank)
    ctx = omk.Context(device=local, capacity_envs=4, capacity_trees=2 * my_games, capacity_nodes=2048, seed=sharding.rank_seed(seed, rank))
    ctx.net_init_random(seed)                       # identical weights on every rank
    step = trainer.TrainStep(ctx.net_get_params(), device=f"cuda:{local}")
    mem = trainer.ReplayMemory(capacity=600_000, seed=seed + rank)

    # ---- self-play phase: per-GPU game pools, transitions streamed to the host replay memory ----
    ctx.selfplay_begin(my_games, count, batch, 0.25, 0.03, 1.0, 30, omk.EVAL_NET)
    t0 = time.perf_counter()
    stats, boards, policy, status, _ = ctx.selfplay_run(plies, profile=0, want_transitions=True)
    sp_s = time.perf_counter() - t0
    episodes, carry = trainer.split_episodes(boards, policy, status)
    for b, p, z in episodes:
        mem.add_episode(b, p, z)
    for bs, ps in carry:                            # unfinished tails still train the policy head; z = 0 (draw-like)
        if bs:
            mem.add_episode(np.stack(bs), np.stack(ps), 0.0)

    # ---- training phase: data parallel over the minibatch, one gradient all-reduce per step ----
    per_rank = max(1, minibatch // world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    losses = []
    for i in range(steps + 2):
        if i == 2:                                  # two untimed warm-up steps
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
            e0.record()
        b, t, pi, z = mem.sample(per_rank)
        losses.append(step.train(trainer.encode_nn_input(b, t), pi, z))
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / steps
    step.sync_to(ctx)                               # refreshed weights for the next self-play phase
    p, _ = ctx.net_eval_images(trainer.encode_nn_input(b[:4], t[:4]))
    times, work = sharding.reduce_measurements({"step_ms": step_ms, "selfplay_s": sp_s},
                                               {"sims": int(stats.simulations), "positions": int(stats.positions),
                                                "episodes": len(episodes), "replay": len(mem)},
                                               device="cuda" if dist else "cpu")
    out = {"config": "AlphaZero iteration (BASELINE configs[4] shape)", "n_gpus": world, "games": games, "plies": plies,
           "sims_per_move": count, "selfplay_sims_per_s": work["sims"] / times["selfplay_s"],
           "selfplay_positions_per_s": work["positions"] / times["selfplay_s"], "episodes_finished": work["episodes"],
           "replay_transitions": work["replay"], "train_steps": steps, "global_minibatch": per_rank * world,
           "train_step_ms": times["step_ms"], "allreduce_bytes_per_step": 5_643_250 * 4 if world > 1 else 0,
           "first_loss": losses[0][2], "last_loss": losses[-1][2], "policy_sums_to_one": bool(abs(float(p[0].sum()) - 1) < 1e-3)}
    if rank == 0 and not quiet:
        print(json.dumps(out), flush=True)
    ctx.close()
    if dist:
        dist.destroy_process_group()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=256)
    ap.add_argument("--plies", type=int, default=24)
    ap.add_argument("--count", type=int, default=160)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--minibatch", type=int, default=128)
    a = ap.parse_args()
    run(games=a.games, plies=a.plies, count=a.count, steps=a.steps, minibatch=a.minibatch)

#!/usr/bin/env python
"""Per-kernel-family time split of the self-play step (CUDA events, profile level 2) and env-step roofline sweep.
Usage: python tools/profile_step.py [--games 1024] [--plies 2] [--evaluator net|hash] [--env]"""
import argparse
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omk = importlib.import_module("omok-ai_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=1024)
ap.add_argument("--plies", type=int, default=2)
ap.add_argument("--warm", type=int, default=2)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--evaluator", default="net")
ap.add_argument("--lanes", type=int, default=0, help="1: keep the two search lanes (kernel spans then include cross-stream queueing)")
ap.add_argument("--env", action="store_true")
ap.add_argument("--tree-sweep", action="store_true", help="tree-kernel bandwidth vs pool size (hash evaluator, no network)")
args = ap.parse_args()

if args.env:
    import torch

    out = []
    for n in (4096, 65536, 1 << 20, 1 << 24):
        ctx = omk.Context(device=0, capacity_envs=n, capacity_trees=1, capacity_nodes=16, seed=0)
        ctx.env_reset(n=n)
        acts = [torch.randint(0, 81, (n,), dtype=torch.uint8, device="cuda") for _ in range(8)]
        st = torch.empty(n, dtype=torch.int8, device="cuda")
        legal = torch.empty((n, 3), dtype=torch.int32, device="cuda")
        stream = torch.cuda.ExternalStream(ctx.stream)
        with torch.cuda.stream(stream):
            for a in acts[:3]:
                ctx.env_step_device(a.data_ptr(), n, st.data_ptr(), legal.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for a in acts[3:]:
                ctx.env_step_device(a.data_ptr(), n, st.data_ptr(), legal.data_ptr())
            e1.record(stream)
        ctx.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out.append({"boards": n, "ms_per_step": ms, "board_steps_per_s": n / ms * 1e3, "algorithmic_GBs": n * 78 / ms / 1e6})
        ctx.close()
    print(json.dumps({"env_step_sweep": out}))
    sys.exit(0)

if args.tree_sweep:
    # K2 roofline sweep: the tree kernels are warp-per-tree, so bandwidth grows with the number of concurrent trees until
    # HBM saturates.  Hash evaluator (no network), 160 simulations per move in rounds of 16, two plies timed.
    # Algorithmic bytes per simulation: 12*S + 16*(L+2) + 1720 - 972 (packed boards, no 972-byte image), S ~ 73, L ~ 1.
    B_SIM = 12 * 73 + 16 * 3 + 1720 - 972
    out = []
    for games in (1024, 4096, 16384, 32768):
        ctx = omk.Context(device=0, capacity_envs=4, capacity_trees=2 * games, capacity_nodes=768, seed=1)
        ctx.debug_set_lane_min_trees(0)  # one stream: the spans then time the kernels alone
        ctx.selfplay_begin(games, 160, 16, 0.25, 0.03, 1.0, 30, omk.EVAL_HASH)
        ctx.selfplay_run(2, profile=0, want_transitions=False)
        stats, *_ = ctx.selfplay_run(2, profile=2, want_transitions=False)
        kinds = stats.by_kind()
        tree_ms = kinds["select_expand"][0] + kinds["apply"][0]
        sims = int(stats.simulations)
        out.append({"games": games, "trees": 2 * games, "sims": sims, "select_expand_ms": round(kinds["select_expand"][0], 3),
                    "apply_ms": round(kinds["apply"][0], 3), "tree_sims_per_s": sims / tree_ms * 1e3,
                    "algorithmic_GBs": sims * B_SIM / tree_ms / 1e6, "bytes_per_sim": B_SIM})
        ctx.close()
    print(json.dumps({"tree_sweep": out}))
    sys.exit(0)

ev = omk.EVAL_NET if args.evaluator == "net" else omk.EVAL_HASH
ctx = omk.Context(device=0, capacity_envs=4, capacity_trees=2 * args.games, capacity_nodes=4096, seed=1)
if not args.lanes:
    ctx.debug_set_lane_min_trees(0)
if ev == omk.EVAL_NET:
    ctx.net_init_random(0)
ctx.selfplay_begin(args.games, 800, args.batch, 0.25, 0.03, 1.0, 30, ev)
ctx.selfplay_run(args.warm, profile=0, want_transitions=False)
stats, *_ = ctx.selfplay_run(args.plies, profile=2, want_transitions=False)
kinds = stats.by_kind()
tot = float(stats.gpu_ms)
res = {"games": args.games, "plies": args.plies, "evaluator": args.evaluator, "gpu_ms": tot, "sims": int(stats.simulations),
       "nn_evals": int(stats.nn_evals), "sims_per_s": stats.simulations / tot * 1e3,
       "kinds": {k: {"ms": round(v[0], 3), "launches": v[1], "share": round(v[0] / tot, 4)} for k, v in kinds.items()}}
print(json.dumps(res))

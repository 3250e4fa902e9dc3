#!/usr/bin/env python
"""Writes the golden fixtures of tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

* reference_env_vectors.json -- TRANSCRIBED from the reference's own unit tests (environment/src/lib.rs:196-427) and
  src/utils.rs:70-108 (board symmetries); nothing here is computed, the oracle and the CUDA path are both checked
  against it.
* search_hash_golden.json   -- GENERATED with the CPU oracle (oracle/omok_oracle.c) and the exact hash evaluator: root
  children (actions in creation order, visit counts, w and p as raw f32 bits), root statistics, node count and
  random-stream position after searches / plies with fixed seeds.  The reference (Rust + TensorFlow) cannot run here, so
  this pins the ORACLE against drift and gives the GPU tests an oracle-independent target; it is not a reference output.
* net_fp64_golden.npz       -- GENERATED with oracle/net_oracle.py in float64 on the N3 random-init weights (seed 0):
  eight positions, priors and values.  Same caveat.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

IP, DRAW, BW, WW = 0, 1, 2, 3

ENV_VECTORS = {
    "source": "environment/src/lib.rs:196-427 (unit tests of the reference), transcribed by hand",
    "status_codes": {"InProgress": IP, "Draw": DRAW, "BlackWin": BW, "WhiteWin": WW},
    "games": [
        {"cite": "lib.rs:201-252 test_place_stone", "moves": list(range(12)), "status": [IP] * 12},
        {"cite": "lib.rs:255-298 horizontal", "moves": [0, 9, 1, 10, 2, 11, 3, 12, 4], "status": [IP] * 8 + [BW]},
        {"cite": "lib.rs:301-344 vertical", "moves": [0, 2, 9, 11, 18, 20, 27, 29, 36], "status": [IP] * 8 + [BW]},
        {"cite": "lib.rs:347-358 diagonal lt-rb", "moves": list(range(36)) + [40], "status": [IP] * 36 + [BW]},
        {"cite": "lib.rs:361-372 diagonal lb-rt", "moves": list(range(36)) + [36], "status": [IP] * 36 + [BW]},
    ],
    "encodings": [
        {"cite": "lib.rs:375-388", "moves": [0], "perspective": "Black", "ones": [0]},
        {"cite": "lib.rs:391-408", "moves": [0, 10, 2, 30], "perspective": "Black", "ones": [0, 21, 4, 61]},
        {"cite": "lib.rs:409-426", "moves": [0, 10, 2, 30], "perspective": "White", "ones": [1, 20, 5, 60]},
    ],
    "symmetries": {
        "cite": "src/utils.rs:70-108 (3x3 example boards of the reference's tests are applied to index maps here)",
        "note": "see tests/test_trainer_cpu.py::test_oracle_symmetries_on_the_reference_golden_vectors for the 3x3 vectors",
    },
}


def f32_bits(a):
    return [int(x) for x in np.asarray(a, np.float32).view(np.uint32)]


def tree_record(agent):
    a, n, w, p = agent.root_children()
    rn, rw, rp, rst, rpol = agent.root_stats()
    return {"actions": [int(x) for x in a], "n": [int(x) for x in n], "w_bits": f32_bits(w), "p_bits": f32_bits(p),
            "root_n": int(rn), "root_w_bits": f32_bits([rw])[0], "root_p_bits": f32_bits([rp])[0], "root_status": int(rst),
            "root_policy_bits": f32_bits(rpol), "nodes": int(agent.node_count), "rng_counter": int(agent.rng_counter),
            "board": [int(x) for x in agent.board()], "turn": int(agent.env.turn)}


def make_search():
    from oracle import oracle as orc

    orc.lib()
    ev = orc.NativeHashEvaluator()
    seed = 2024
    out = {"seed": seed, "evaluator": "hash", "searches": [], "self_play": None}
    for count, batch, eps, alpha in [(800, 16, 0.0, 1.0), (800, 8, 0.0, 1.0), (600, 16, 0.25, 0.03), (50, 1, 0.25, 0.3), (96, 32, 0.25, 2.5)]:
        streams = list(range(100, 104))
        agents = [orc.Agent(ev, seed, s) for s in streams]
        orc.execute(agents, count, batch, eps, alpha, ev)
        out["searches"].append({"count": count, "batch": batch, "epsilon": eps, "alpha": alpha, "streams": streams,
                                "trees": [tree_record(a) for a in agents]})
    # trainer-shaped self-play (src/trainer.rs:86-204) of two games, eight plies: Boltzmann(1.0) for plies < 4, then Best
    G, plies, threshold = 2, 8, 4
    black = [orc.Agent(ev, seed, 2 * g) for g in range(G)]
    white = [orc.Agent(ev, seed, 2 * g + 1) for g in range(G)]
    steps = []
    for ply in range(plies):
        movers = [black[g] if ply % 2 == 0 else white[g] for g in range(G)]
        others = [white[g] if ply % 2 == 0 else black[g] for g in range(G)]
        orc.execute(movers, 200, 16, 0.25, 0.03, ev)
        mode = 1 if ply < threshold else 0
        ref = [m.sample_action(mode, 1.0) for m in movers]
        acts = [int(r[0]) for r in ref]
        st = [m.play_action(a) for m, a in zip(movers, acts)]
        for o, a in zip(others, acts):
            o.ensure_action_exists(a, ev)
        st2 = [o.play_action(a) for o, a in zip(others, acts)]
        steps.append({"ply": ply, "mode": mode, "actions": acts, "policy_bits": [f32_bits(r[1]) for r in ref],
                      "status": [int(s) for s in st], "status_other": [(-1 if s is None else int(s)) for s in st2],
                      "movers": [tree_record(m) for m in movers], "others": [tree_record(o) for o in others]})
        assert all(s == 0 for s in st), "fixture games must not end inside the fixture"
    out["self_play"] = {"games": G, "count": 200, "batch": 16, "epsilon": 0.25, "alpha": 0.03, "temperature": 1.0,
                        "threshold": threshold, "steps": steps}
    return out


def make_net():
    import torch

    from oracle import net_oracle

    params = net_oracle.random_params(0)
    rng = np.random.default_rng(7)
    boards = np.zeros((8, 81), np.uint8)
    turns = np.zeros(8, np.uint8)
    for b in range(8):
        k = int(rng.integers(0, 40))
        for j, c in enumerate(rng.permutation(81)[:k]):
            boards[b, c] = 1 + (j % 2)
        turns[b] = k % 2
    p, v, _ = net_oracle.forward_boards(params, boards, turns, dtype=torch.float64)
    return {"boards": boards, "turns": turns, "p": np.asarray(p, np.float64), "v": np.asarray(v, np.float64)}


if __name__ == "__main__":
    json.dump(ENV_VECTORS, open(os.path.join(HERE, "reference_env_vectors.json"), "w"), indent=1)
    json.dump(make_search(), open(os.path.join(HERE, "search_hash_golden.json"), "w"), separators=(",", ":"))
    np.savez_compressed(os.path.join(HERE, "net_fp64_golden.npz"), **make_net())
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))

#!/usr/bin/env python
"""Per-layer error of the network kernels against an fp64 forward, on a LARGE sample ("what's weak" 1).

Positions: thirds of early (0-30 stones), middle (30-60) and late (60-80 stones) random boards, both EnvTurnMode
encodings; weights: the reference's random-init recipe under several seeds plus one "trained-scale" set (policy/value
head and fc1 weights scaled down so the logits have a standard deviation of a few units instead of ~25).  For every
weight set the GPU path (tensor-core kernels, and the fp32 CUDA-core A/B kernels for comparison) is compared layer by
layer with oracle/net_oracle.forward_layers in fp64: tower output, fc0 output, fc1 output, priors, values.

This is test tooling (it imports oracle/): python tools/net_error_study.py [--positions 20000] [--out FILE.json]"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def positions(n, seed):
    rng = np.random.default_rng(seed)
    boards = np.zeros((n, 81), np.uint8)
    turns = np.zeros(n, np.uint8)
    bands = ((0, 31), (30, 61), (60, 81))
    for b in range(n):
        lo, hi = bands[b % 3]
        k = int(rng.integers(lo, hi))
        cells = rng.permutation(81)[:k]
        boards[b, cells[0::2]] = 1
        boards[b, cells[1::2]] = 2
        turns[b] = k % 2
    return boards, turns


def weight_sets(no, seeds):
    for s in seeds:
        yield f"init_seed{s}", no.random_params(s)
    p = no.random_params(seeds[0])
    names = [n for n, _ in no.PARAM_SPECS]
    for nm, f in (("fc1_w", 0.5), ("p_w", 0.25), ("v_w", 0.25)):
        p[names.index(nm)] = (p[names.index(nm)] * np.float32(f)).astype(np.float32)
    yield "trained_scale", p


def summarize(err):
    err = np.asarray(err, dtype=np.float64).reshape(-1)
    q = np.quantile(err, [0.5, 0.9, 0.99, 0.999])
    edges = [0, 1e-7, 3e-7, 1e-6, 3e-6, 1e-5, 3e-5, 1e-4, 2e-4, 3e-4, 5e-4, 1e-3, np.inf]
    hist, _ = np.histogram(err, bins=edges)
    return {"max": float(err.max()), "mean": float(err.mean()), "p50": float(q[0]), "p90": float(q[1]), "p99": float(q[2]),
            "p999": float(q[3]), "hist_edges": [float(e) if np.isfinite(e) else "inf" for e in edges], "hist": hist.tolist()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--positions", type=int, default=20000)
    ap.add_argument("--seeds", type=int, nargs="+", default=[0, 1, 2])
    ap.add_argument("--chunk", type=int, default=4000)
    ap.add_argument("--layer-rows", type=int, default=600, help="rows per weight set whose tower output is compared (166 KB per row)")
    ap.add_argument("--paths", nargs="+", default=["tc", "simt"], help="tc = tensor-core kernels (product), simt = fp32 CUDA-core A/B kernels, cpu32 = PyTorch-CPU fp32 (yardstick)")
    ap.add_argument("--fc0-chunk", type=int, default=9, help="k-blocks of fc0 per TMEM drain: 9 (default) or 3 (omk_debug_set_fc0_chunk)")
    ap.add_argument("--label", default="")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch

    omk = importlib.import_module("omok-ai_b200")
    from oracle import net_oracle as no

    torch.set_num_threads(os.cpu_count() or 8)
    boards, turns = positions(args.positions, 12345)
    modes = (np.arange(args.positions) // 3) % 2  # EnvTurnMode::Player / ::Opponent alternate within each band
    ctx = omk.Context(device=0, capacity_envs=4, capacity_trees=4, capacity_nodes=64, seed=0)
    ctx.debug_set_fc0_chunk(args.fc0_chunk)
    result = {"label": args.label, "fc0_chunk": args.fc0_chunk, "positions": args.positions, "stones": {"early": "0-30", "middle": "30-60", "late": "60-80"},
              "tolerance": 1e-3, "sets": {}}
    t_start = time.time()
    for set_name, params in weight_sets(no, args.seeds):
        ctx.net_load_params(params)
        per_path = {p: {k: [] for k in ("P", "V", "Vabs", "tower", "fc0", "fc1")} for p in args.paths}
        scale = {}
        ref_logit_std = []
        for c0 in range(0, args.positions, args.chunk):
            sl = slice(c0, min(c0 + args.chunk, args.positions))
            b, t, m = boards[sl], turns[sl], modes[sl]
            imgs = np.stack([no.encode_image(bb, int(tt), bool(mm)) for bb, tt, mm in zip(b, t, m)])
            ref = no.forward_layers(params, imgs, torch.float64)
            ref_logit_std.append(float(ref["logits"].std()))
            for k in ("tower", "fc0", "fc1"):
                scale[k] = max(scale.get(k, 0.0), float(np.abs(ref[k]).max()))
            n = b.shape[0]
            for path in args.paths:
                if path == "cpu32":  # the yardstick: a plain fp32 forward on the CPU (PyTorch), as any fp32 implementation errs
                    c32 = no.forward_layers(params, imgs, torch.float32)
                    for k in ("tower", "fc0", "fc1"):
                        per_path[path][k].append(np.abs(c32[k].astype(np.float64) - ref[k]).reshape(-1))
                    big = ref["P"] > 1e-12
                    per_path[path]["P"].append((np.abs(c32["P"].astype(np.float64) - ref["P"])[big] / ref["P"][big]).reshape(-1))
                    per_path[path]["V"].append(np.abs(c32["V"].astype(np.float64) - ref["V"]) / np.maximum(np.abs(ref["V"]), 1e-3))
                    per_path[path]["Vabs"].append(np.abs(c32["V"].astype(np.float64) - ref["V"]))
                    continue
                mode_flag = 1 if path == "tc" else 0
                ctx.debug_set_tower_mode(mode_flag)
                ctx.debug_set_fc0_mode(mode_flag)
                # the two encodings go through the boards entry point separately (mode is per call)
                p = np.zeros((n, 81), np.float32)
                v = np.zeros(n, np.float32)
                lay = {k: None for k in ("tower", "fc0", "fc1")}
                for md in (0, 1):
                    idx = np.flatnonzero(m == md)
                    if not idx.size:
                        continue
                    pp, vv = ctx.net_eval(b[idx], t[idx], mode=md)
                    p[idx], v[idx] = pp, vv
                    k_rows = idx.size
                    ids = (8, 9, 10) if path == "tc" else (0, 1, 2)
                    rows_t = min(k_rows, args.layer_rows // 2) if c0 == 0 else 0
                    got = {"fc0": ctx.debug_get_buffer(ids[1], k_rows * 512).reshape(k_rows, 512),
                           "fc1": ctx.debug_get_buffer(ids[2], k_rows * 512).reshape(k_rows, 512)}
                    if rows_t:
                        got["tower"] = ctx.debug_get_buffer(ids[0], rows_t * 10368).reshape(rows_t, 81, 128)
                    for k, g in got.items():
                        r = ref[k][idx[: g.shape[0]]]
                        per_path[path][k].append(np.abs(g.astype(np.float64) - r).reshape(-1))
                big = ref["P"] > 1e-12
                per_path[path]["P"].append((np.abs(p.astype(np.float64) - ref["P"])[big] / ref["P"][big]).reshape(-1))
                per_path[path]["V"].append(np.abs(v.astype(np.float64) - ref["V"]) / np.maximum(np.abs(ref["V"]), 1e-3))
                per_path[path]["Vabs"].append(np.abs(v.astype(np.float64) - ref["V"]))
        entry = {"ref_logit_std": float(np.mean(ref_logit_std)), "layer_max_abs": scale, "paths": {}}
        for path in args.paths:
            e = {}
            for k in ("P", "V"):
                e[k + "_rel"] = summarize(np.concatenate(per_path[path][k]))
            vabs = np.concatenate(per_path[path]["Vabs"])
            vrel = np.concatenate(per_path[path]["V"])
            # values are tanh(logit) with logits of scale ~20: near a zero crossing a RELATIVE error is ill-conditioned for
            # any fp32 implementation, so the absolute error and the count beyond 1e-3 relative are reported next to it
            e["V_abs"] = {"max": float(vabs.max()), "p99": float(np.quantile(vabs, 0.99))}
            e["V_rel_over_1e-3"] = int(np.count_nonzero(vrel > 1e-3))
            for k in ("tower", "fc0", "fc1"):
                d = np.concatenate(per_path[path][k])
                e[k + "_abs_over_layer_max"] = {"max": float(d.max() / scale[k]), "rms": float(np.sqrt(np.mean(d * d)) / scale[k])}
            entry["paths"][path] = e
            print(f"[{args.label}] {set_name:14s} {path:4s} P rel max {e['P_rel']['max']:.3e} p99 {e['P_rel']['p99']:.2e} | "
                  f"V rel max {e['V_rel']['max']:.3e} (>{1e-3:g}: {e['V_rel_over_1e-3']}, abs max {e['V_abs']['max']:.2e}) | tower {e['tower_abs_over_layer_max']['max']:.2e} "
                  f"fc0 {e['fc0_abs_over_layer_max']['max']:.2e} fc1 {e['fc1_abs_over_layer_max']['max']:.2e} (of the layer's max)", flush=True)
        result["sets"][set_name] = entry
    result["seconds"] = time.time() - t_start
    ctx.close()
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(result, f, indent=1)
    print(json.dumps({"label": args.label, "worst_P_rel": max(s["paths"][args.paths[0]]["P_rel"]["max"] for s in result["sets"].values()),
                      "worst_V_rel": max(s["paths"][args.paths[0]]["V_rel"]["max"] for s in result["sets"].values())}))


if __name__ == "__main__":
    main()

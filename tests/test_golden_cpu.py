"""The committed fixtures of tests/golden/ against the CPU oracle: the reference's own environment test vectors
(transcribed, environment/src/lib.rs:196-427) pin the oracle; the oracle-generated search / network fixtures pin it
against drift (tests/golden/make_golden.py wrote them; the same files are the GPU path's target in test_golden_gpu.py)."""
import json
import os

import numpy as np

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return [int(x) for x in np.asarray(a, np.float32).view(np.uint32)]


def check_tree(agent, rec, where):
    a, n, w, p = agent.root_children()
    assert [int(x) for x in a] == rec["actions"], f"{where}: child actions / creation order"
    assert [int(x) for x in n] == rec["n"], f"{where}: visit counts"
    assert bits(w) == rec["w_bits"] and bits(p) == rec["p_bits"], f"{where}: w / p bits"
    rn, rw, rp, rst, rpol = agent.root_stats()
    assert (int(rn), int(rst)) == (rec["root_n"], rec["root_status"]), where
    assert bits([rw])[0] == rec["root_w_bits"] and bits([rp])[0] == rec["root_p_bits"], where
    assert bits(rpol) == rec["root_policy_bits"], f"{where}: root policy"
    assert (agent.node_count, agent.rng_counter) == (rec["nodes"], rec["rng_counter"]), f"{where}: nodes / stream position"
    assert [int(x) for x in agent.board()] == rec["board"] and agent.env.turn == rec["turn"], where


def test_reference_env_vectors_pin_the_oracle(orc):
    v = json.load(open(os.path.join(G, "reference_env_vectors.json")))
    for g in v["games"]:
        env = orc.Environment()
        got = [env.place_stone(m) for m in g["moves"]]
        assert got == g["status"], g["cite"]
    for e in v["encodings"]:
        env = orc.Environment()
        for m in e["moves"]:
            env.place_stone(m)
        exp = np.zeros(162, np.float32)
        exp[e["ones"]] = 1
        assert np.array_equal(env.encode_board(0 if e["perspective"] == "Black" else 1), exp), e["cite"]


def test_search_fixture_reproduced_by_the_oracle(orc):
    f = json.load(open(os.path.join(G, "search_hash_golden.json")))
    ev = orc.NativeHashEvaluator()
    for s in f["searches"]:
        agents = [orc.Agent(ev, f["seed"], st) for st in s["streams"]]
        orc.execute(agents, s["count"], s["batch"], s["epsilon"], s["alpha"], ev)
        for a, rec in zip(agents, s["trees"]):
            check_tree(a, rec, f"search {s['count']}/{s['batch']}")


def test_self_play_fixture_reproduced_by_the_oracle(orc):
    f = json.load(open(os.path.join(G, "search_hash_golden.json")))
    sp = f["self_play"]
    ev = orc.NativeHashEvaluator()
    black = [orc.Agent(ev, f["seed"], 2 * g) for g in range(sp["games"])]
    white = [orc.Agent(ev, f["seed"], 2 * g + 1) for g in range(sp["games"])]
    for step in sp["steps"]:
        ply = step["ply"]
        movers = [black[g] if ply % 2 == 0 else white[g] for g in range(sp["games"])]
        others = [white[g] if ply % 2 == 0 else black[g] for g in range(sp["games"])]
        orc.execute(movers, sp["count"], sp["batch"], sp["epsilon"], sp["alpha"], ev)
        ref = [m.sample_action(step["mode"], sp["temperature"]) for m in movers]
        assert [int(r[0]) for r in ref] == step["actions"]
        assert [bits(r[1]) for r in ref] == step["policy_bits"]
        assert [m.play_action(a) for m, a in zip(movers, step["actions"])] == step["status"]
        for o, a in zip(others, step["actions"]):
            o.ensure_action_exists(a, ev)
        assert [(-1 if s is None else int(s)) for s in (o.play_action(a) for o, a in zip(others, step["actions"]))] == step["status_other"]
        for m, o, rm, ro in zip(movers, others, step["movers"], step["others"]):
            check_tree(m, rm, f"ply {ply} mover")
            check_tree(o, ro, f"ply {ply} other")


def test_network_fixture_reproduced_by_the_fp32_oracle():
    from oracle import net_oracle

    z = np.load(os.path.join(G, "net_fp64_golden.npz"))
    p, v, _ = net_oracle.forward_boards(net_oracle.random_params(0), z["boards"], z["turns"])  # float32 restatement
    big = z["p"] > 1e-12
    assert np.max(np.abs(p[big] - z["p"][big]) / z["p"][big]) < 1e-3  # north_star tolerance: 1e-3 relative
    assert np.max(np.abs(v - z["v"]) / np.maximum(np.abs(z["v"]), 1e-3)) < 1e-3

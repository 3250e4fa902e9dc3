"""The CUDA path (through the C ABI) against the committed fixtures of tests/golden/ -- no oracle in the loop: the
reference's environment vectors, the hash-evaluator search / self-play fixture (bit-exact) and the fp64 network fixture
(1e-3 relative, the north-star tolerance)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return [int(x) for x in np.asarray(a, np.float32).view(np.uint32)]


@pytest.fixture(scope="module")
def ctx(omk):
    seed = json.load(open(os.path.join(G, "search_hash_golden.json")))["seed"]
    c = omk.Context(device=0, capacity_envs=16, capacity_trees=16, capacity_nodes=4096, seed=seed)
    yield c
    c.close()


def check_tree(ctx, tree, rec, where):
    a, n, w, p = ctx.pool_root_children(tree)
    assert [int(x) for x in a] == rec["actions"], f"{where}: child actions / creation order"
    assert [int(x) for x in n] == rec["n"], f"{where}: visit counts"
    assert bits(w) == rec["w_bits"] and bits(p) == rec["p_bits"], f"{where}: w / p bits"
    rn, rw, rp, rst, rpol = ctx.pool_root_stats(tree)
    assert (int(rn), int(rst)) == (rec["root_n"], rec["root_status"]), where
    assert bits([rw])[0] == rec["root_w_bits"] and bits([rp])[0] == rec["root_p_bits"], where
    assert bits(rpol) == rec["root_policy_bits"], f"{where}: root policy"
    nodes, ctr = ctx.pool_tree_info(tree)
    assert (nodes, ctr) == (rec["nodes"], rec["rng_counter"]), f"{where}: nodes / stream position"
    board, turn, _ = ctx.pool_get_env(tree)
    assert [int(x) for x in board] == rec["board"] and turn == rec["turn"], where


def test_reference_env_vectors(ctx, omk):
    v = json.load(open(os.path.join(G, "reference_env_vectors.json")))
    for i, g in enumerate(v["games"]):
        ctx.env_reset(ids=[i])
        got = [int(ctx.env_step([m], ids=[i])[0][0]) for m in g["moves"]]
        assert got == g["status"], g["cite"]
    for e in v["encodings"]:
        env = omk.Environment(ctx)  # host mirror of the reference's Environment over the env-pool kernels
        for m in e["moves"]:
            env.place_stone(m)
        exp = np.zeros(162, np.float32)
        exp[e["ones"]] = 1
        assert np.array_equal(env.encode_board(omk.Turn.Black if e["perspective"] == "Black" else omk.Turn.White), exp), e["cite"]


def test_search_fixture_bit_exact(ctx, omk):
    f = json.load(open(os.path.join(G, "search_hash_golden.json")))
    for s in f["searches"]:
        T = len(s["streams"])
        ctx.pool_new_games(n=T, streams=np.array(s["streams"], np.uint32), evaluator=omk.EVAL_HASH)
        ctx.pool_search(n=T, count=s["count"], batch_size=s["batch"], epsilon=s["epsilon"], alpha=s["alpha"], evaluator=omk.EVAL_HASH)
        for t, rec in enumerate(s["trees"]):
            check_tree(ctx, t, rec, f"search {s['count']}/{s['batch']} tree {t}")


def test_self_play_fixture_bit_exact(ctx, omk):
    f = json.load(open(os.path.join(G, "search_hash_golden.json")))
    sp = f["self_play"]
    Gm = sp["games"]
    ctx.pool_new_games(n=2 * Gm, evaluator=omk.EVAL_HASH)  # stream == tree id
    for step in sp["steps"]:
        ply = step["ply"]
        mids = [2 * g + (ply % 2) for g in range(Gm)]
        oids = [2 * g + 1 - (ply % 2) for g in range(Gm)]
        ctx.pool_search(ids=mids, count=sp["count"], batch_size=sp["batch"], epsilon=sp["epsilon"], alpha=sp["alpha"], evaluator=omk.EVAL_HASH)
        acts, pol = ctx.pool_sample(ids=mids, modes=[step["mode"]] * Gm, temperatures=[sp["temperature"]] * Gm)
        assert [int(a) for a in acts] == step["actions"], f"ply {ply}: sampled actions"
        assert [bits(r) for r in pol] == step["policy_bits"], f"ply {ply}: visit policy"
        assert [int(s) for s in ctx.pool_play(acts, ids=mids)] == step["status"]
        ctx.pool_ensure_action(acts, ids=oids, evaluator=omk.EVAL_HASH)
        assert [int(s) for s in ctx.pool_play(acts, ids=oids)] == step["status_other"]
        for g in range(Gm):
            check_tree(ctx, mids[g], step["movers"][g], f"ply {ply} game {g} mover")
            check_tree(ctx, oids[g], step["others"][g], f"ply {ply} game {g} other")


def test_network_fixture_within_tolerance(omk):
    z = np.load(os.path.join(G, "net_fp64_golden.npz"))
    c = omk.Context(device=0, capacity_envs=4, capacity_trees=4, capacity_nodes=64, seed=0)
    from oracle import net_oracle  # only for the seeded N3 weights the fixture was made with (22 MB: not committed)

    c.net_load_params(net_oracle.random_params(0))
    p, v = c.net_eval(z["boards"], z["turns"])
    c.close()
    big = z["p"] > 1e-12
    assert np.max(np.abs(p[big] - z["p"][big]) / z["p"][big]) < 1e-3  # north_star tolerance: 1e-3 relative
    assert np.max(np.abs(v - z["v"]) / np.maximum(np.abs(z["v"]), 1e-3)) < 1e-3

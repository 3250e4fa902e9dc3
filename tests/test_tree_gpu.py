"""K2 parity (through the C ABI): search with the exact hash evaluator must reproduce the oracle's
visit counts, w, p, root statistics, random-stream position and node counts BIT-EXACTLY, including
across sample / play / ensure_action_exists / re-rooting."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(omk):
    c = omk.Context(device=0, capacity_envs=4, capacity_trees=256, capacity_nodes=4096, seed=2024)
    yield c
    c.close()


def assert_tree_equal(ctx, tree, agent, where=""):
    ga, gn, gw, gp = ctx.pool_root_children(tree)
    oa, on, ow, op = agent.root_children()
    assert np.array_equal(ga, oa), f"{where}: child actions / creation order"
    assert np.array_equal(gn, on), f"{where}: visit counts"
    assert gw.tobytes() == ow.tobytes(), f"{where}: w"
    assert gp.tobytes() == op.tobytes(), f"{where}: p"
    n, w, p, st, pol = ctx.pool_root_stats(tree)
    rn, rw, rp, rst, rpol = agent.root_stats()
    assert (n, st) == (rn, rst), where
    assert np.float32(w).tobytes() == np.float32(rw).tobytes(), where
    assert np.float32(p).tobytes() == np.float32(rp).tobytes(), where
    assert pol.tobytes() == rpol.tobytes(), f"{where}: root policy"
    nodes, ctr = ctx.pool_tree_info(tree)
    assert nodes == agent.node_count, f"{where}: node count"
    assert ctr == agent.rng_counter, f"{where}: rng counter"
    board, turn, legal = ctx.pool_get_env(tree)
    assert np.array_equal(board, agent.board()) and turn == agent.env.turn and legal == agent.env.legal_move_count


@pytest.mark.parametrize("count,batch,eps,alpha", [(800, 16, 0.0, 1.0), (800, 8, 0.0, 1.0), (600, 16, 0.25, 0.03), (50, 1, 0.25, 0.3), (100, 64, 0.5, 1.0), (96, 32, 0.25, 2.5)])
def test_search_matches_oracle_bit_exact(ctx, omk, orc, count, batch, eps, alpha):
    T = 24
    ev = orc.NativeHashEvaluator()
    streams = np.arange(100, 100 + T, dtype=np.uint32)
    agents = [orc.Agent(ev, ctx.seed, int(s)) for s in streams]
    ctx.pool_new_games(n=T, streams=streams, evaluator=omk.EVAL_HASH)
    for t in range(T):
        assert_tree_equal(ctx, t, agents[t], "new")
    orc.execute(agents, count, batch, eps, alpha, ev)
    ctx.pool_search(n=T, count=count, batch_size=batch, epsilon=eps, alpha=alpha, evaluator=omk.EVAL_HASH)
    for t in range(T):
        assert_tree_equal(ctx, t, agents[t], f"tree {t}")


def test_full_self_play_games_match_oracle(ctx, omk, orc):
    """Trainer-shaped self-play (src/trainer.rs:86-204): two agents per game, Boltzmann for the first plies then
    Best, ensure_action_exists + play on the opponent tree; compared after every step until all games end."""
    G = 12
    ev = orc.NativeHashEvaluator()
    black = [orc.Agent(ev, ctx.seed, 2 * g) for g in range(G)]
    white = [orc.Agent(ev, ctx.seed, 2 * g + 1) for g in range(G)]
    ctx.pool_new_games(n=2 * G, evaluator=omk.EVAL_HASH)  # stream == tree id
    live = list(range(G))
    ply = 0
    threshold = 6
    while live and ply < 81:
        movers = [black[g] if ply % 2 == 0 else white[g] for g in live]
        others = [white[g] if ply % 2 == 0 else black[g] for g in live]
        mids = [2 * g + (ply % 2) for g in live]
        oids = [2 * g + 1 - (ply % 2) for g in live]
        orc.execute(movers, 200, 16, 0.25, 0.03, ev)
        ctx.pool_search(ids=mids, count=200, batch_size=16, epsilon=0.25, alpha=0.03, evaluator=omk.EVAL_HASH)
        mode = 1 if ply < threshold else 0
        acts, pol = ctx.pool_sample(ids=mids, modes=[mode] * len(live), temperatures=[1.0] * len(live))
        ref = [m.sample_action(mode, 1.0) for m in movers]
        assert [int(a) for a in acts] == [r[0] for r in ref], f"ply {ply}: sampled actions"
        assert np.stack([r[1] for r in ref]).tobytes() == pol.tobytes(), f"ply {ply}: visit policy"
        st = ctx.pool_play(acts, ids=mids)
        rst = [m.play_action(int(a)) for m, a in zip(movers, acts)]
        assert [int(s) for s in st] == rst
        ctx.pool_ensure_action(acts, ids=oids, evaluator=omk.EVAL_HASH)
        for o, a in zip(others, acts):
            o.ensure_action_exists(int(a), ev)
        st2 = ctx.pool_play(acts, ids=oids)
        rst2 = [o.play_action(int(a)) for o, a in zip(others, acts)]
        assert [int(s) for s in st2] == [(-1 if r is None else r) for r in rst2]
        for g, mid, oid in zip(live, mids, oids):
            assert_tree_equal(ctx, mid, black[g] if ply % 2 == 0 else white[g], f"ply {ply} game {g} mover")
            assert_tree_equal(ctx, oid, white[g] if ply % 2 == 0 else black[g], f"ply {ply} game {g} other")
        live = [g for g, s in zip(live, st) if s == 0]
        ply += 1
    assert not live or ply == 81


@pytest.mark.parametrize("batch", [16, 32])
def test_late_game_positions_match_oracle(ctx, omk, orc, batch):
    """Late positions -- the reference-derived draw / win-on-move-81 boards (tests/test_oracle_env.py) up to 66..77 stones --
    where terminal children (backed up at once, parallel_mcts_executor.rs:177-181), terminal leaves (:92-97), nodes that fill
    after a handful of expansions and roots with fewer empty cells than a round has simulations are the rule (the regime in
    which a run of lane-parallel expansions is cut behind a terminal child and its draws are handed back): fed in move by
    move through ensure_action_exists + play_action, then searched and played to the end, compared after every step."""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_oracle_env import draw_moves, win81_moves

    openings = [draw_moves()[:66], draw_moves()[:75], win81_moves()[:70], win81_moves()[:77]]
    G = len(openings)
    ev = orc.NativeHashEvaluator()
    black = [orc.Agent(ev, ctx.seed, 2 * g) for g in range(G)]
    white = [orc.Agent(ev, ctx.seed, 2 * g + 1) for g in range(G)]
    ctx.pool_new_games(n=2 * G, evaluator=omk.EVAL_HASH)  # stream == tree id
    for k in range(max(len(o) for o in openings)):
        games = [g for g in range(G) if k < len(openings[g])]
        ids = [2 * g for g in games] + [2 * g + 1 for g in games]
        acts = [openings[g][k] for g in games] * 2
        ctx.pool_ensure_action(acts, ids=ids, evaluator=omk.EVAL_HASH)
        st = ctx.pool_play(acts, ids=ids)
        assert not np.any(st), f"opening move {k}"
        for g in games:
            for agent in (black[g], white[g]):
                agent.ensure_action_exists(openings[g][k], ev)
                assert agent.play_action(openings[g][k]) == 0
    for g in range(G):
        assert_tree_equal(ctx, 2 * g, black[g], f"game {g} opening, black")
        assert_tree_equal(ctx, 2 * g + 1, white[g], f"game {g} opening, white")
    ply = [len(o) for o in openings]
    live = list(range(G))
    for _ in range(16):
        if not live:
            break
        movers = [black[g] if ply[g] % 2 == 0 else white[g] for g in live]
        others = [white[g] if ply[g] % 2 == 0 else black[g] for g in live]
        mids = [2 * g + (ply[g] % 2) for g in live]
        oids = [2 * g + 1 - (ply[g] % 2) for g in live]
        orc.execute(movers, 48, batch, 0.25, 0.3, ev)
        ctx.pool_search(ids=mids, count=48, batch_size=batch, epsilon=0.25, alpha=0.3, evaluator=omk.EVAL_HASH)
        for g, mid, m in zip(live, mids, movers):
            assert_tree_equal(ctx, mid, m, f"game {g} ply {ply[g]} searched")
        acts, pol = ctx.pool_sample(ids=mids, modes=[0] * len(live), temperatures=[1.0] * len(live))
        ref = [m.sample_action(0, 1.0) for m in movers]
        assert [int(a) for a in acts] == [r[0] for r in ref]
        assert np.stack([r[1] for r in ref]).tobytes() == pol.tobytes()
        st = ctx.pool_play(acts, ids=mids)
        assert [int(x) for x in st] == [m.play_action(int(a)) for m, a in zip(movers, acts)]
        ctx.pool_ensure_action(acts, ids=oids, evaluator=omk.EVAL_HASH)
        for o, a in zip(others, acts):
            o.ensure_action_exists(int(a), ev)
        st2 = ctx.pool_play(acts, ids=oids)
        rst2 = [o.play_action(int(a)) for o, a in zip(others, acts)]
        assert [int(x) for x in st2] == [(-1 if r is None else r) for r in rst2]
        for g, mid, oid, m, o in zip(live, mids, oids, movers, others):
            assert_tree_equal(ctx, mid, m, f"game {g} ply {ply[g]} mover re-rooted")
            assert_tree_equal(ctx, oid, o, f"game {g} ply {ply[g]} other re-rooted")
        for g in live:
            ply[g] += 1
        live = [g for g, x in zip(live, st) if x == 0]
    assert not live, "every game must end on a board with at most fifteen empty cells"


def test_play_edge_cases(ctx, omk):
    ctx.pool_new_games(ids=[200], evaluator=omk.EVAL_HASH)
    assert ctx.pool_play([5], ids=[200])[0] == -1  # action not in the tree
    assert ctx.pool_policy(ids=[200])[1][0] == 0  # no children -> None
    assert ctx.pool_sample(ids=[200])[0][0] == -1
    ctx.pool_ensure_action([5], ids=[200], evaluator=omk.EVAL_HASH)
    assert ctx.pool_policy(ids=[200])[1][0] == 0  # child exists but no visits -> None (sum < EPSILON)
    assert ctx.pool_play([5], ids=[200])[0] == 0
    ctx.pool_ensure_action([5], ids=[200], evaluator=omk.EVAL_HASH)  # occupied cell: child for an illegal action
    assert ctx.pool_play([5], ids=[200])[0] == -1  # place_stone refuses -> None, tree untouched
    assert ctx.pool_tree_info(200)[0] == 2
    ctx.pool_ensure_action([81], ids=[200], evaluator=omk.EVAL_HASH)  # out of range: no-op
    assert ctx.pool_tree_info(200)[0] == 2


def test_capacity_error_is_loud(omk):
    c = omk.Context(device=0, capacity_envs=1, capacity_trees=2, capacity_nodes=64, seed=1)
    c.pool_new_games(n=2, evaluator=omk.EVAL_HASH)
    with pytest.raises(omk.OmkError) as e:
        c.pool_search(n=2, count=800, batch_size=16, epsilon=0.0, alpha=1.0, evaluator=omk.EVAL_HASH)
    assert e.value.code == -3
    c.close()


def test_reference_shaped_api(ctx, omk, orc):
    """The host mirror reads like the reference: Agent / ParallelMCTSExecutor / MCTSExecutor."""
    ev = orc.NativeHashEvaluator()
    a = omk.Agent(ctx, evaluator=omk.EVAL_HASH, stream=7)
    ref = orc.Agent(ev, ctx.seed, 7)
    assert a.compute_policy() is None and ref.compute_policy() is None
    omk.MCTSExecutor().run(800, 8, 0.0, 1.0, a)  # benchmark crate constants (benchmark/src/main.rs:9-10)
    orc.execute([ref], 800, 8, 0.0, 1.0, ev)
    act, pol = a.sample_action(omk.ActionSamplingMode.Best)
    ract, rpol = ref.sample_action(0)
    assert act == ract and pol.tobytes() == rpol.tobytes()
    assert a.play_action(act) == omk.GameStatus.InProgress
    assert a.play_action(act) is None
    assert a.env["turn"] == omk.Turn.White


@pytest.mark.parametrize("ids_kind", ["identity", "explicit"])
def test_two_lane_search_matches_oracle_bit_exact(omk, orc, ids_kind):
    """Large searches split their trees over two lanes (two streams, two evaluator workspaces, omk_api.cu LaneScope).
    Forced here for a small pool (OMK_LANE_MIN_TREES=2): every tree must still match the oracle bit for bit, with the
    identity id list (lane 1 starts at an id offset) and with an explicit, shuffled id list; and the self-play driver
    must produce the same transitions with and without lanes."""
    import os

    T = 11  # odd: the lanes get 6 and 5 trees
    ev = orc.NativeHashEvaluator()
    os.environ["OMK_LANE_MIN_TREES"] = "2"
    try:
        c = omk.Context(device=0, capacity_envs=4, capacity_trees=32, capacity_nodes=2048, seed=77)
    finally:
        del os.environ["OMK_LANE_MIN_TREES"]
    ids = None if ids_kind == "identity" else np.array([20, 3, 9, 14, 1, 30, 7, 8, 25, 2, 11], dtype=np.int32)
    trees = list(range(T)) if ids is None else [int(i) for i in ids]
    agents = [orc.Agent(ev, c.seed, t) for t in trees]  # stream == tree id
    c.pool_new_games(ids=ids, n=T, evaluator=omk.EVAL_HASH)
    for rnd in range(2):
        orc.execute(agents, 96, 16, 0.25, 0.03, ev)
        c.pool_search(ids=ids, n=T, count=96, batch_size=16, epsilon=0.25, alpha=0.03, evaluator=omk.EVAL_HASH)
        for t, a in zip(trees, agents):
            assert_tree_equal(c, t, a, f"round {rnd} tree {t}")
    c.close()
    if ids_kind == "identity":
        out = []
        for lanes in ("2", "0"):
            os.environ["OMK_LANE_MIN_TREES"] = lanes
            try:
                c = omk.Context(device=0, capacity_envs=4, capacity_trees=2 * 9, capacity_nodes=2048, seed=5)
            finally:
                del os.environ["OMK_LANE_MIN_TREES"]
            c.selfplay_begin(9, 64, 16, 0.25, 0.03, 1.0, 4, omk.EVAL_HASH)
            stats, boards, policy, status, actions = c.selfplay_run(14, profile=0, want_transitions=True)
            out.append((boards.tobytes(), policy.tobytes(), status.tobytes(), actions.tobytes(), int(stats.simulations), int(stats.nn_evals)))
            c.close()
        assert out[0] == out[1]


def test_config3_full_size_pool_properties(omk, orc):
    """BASELINE config 3 at full size -- 1024 concurrent trees x 800 simulations per move (rounds of 16, eps 0.25,
    alpha 0.03) -- with the exact hash evaluator, checked through properties that do not need the oracle to replay a
    million simulations: (1) the simulation counter, (2) root.n == sum of child visits + the root's own evaluation-free
    start, (3) compute_policy == n_child * recip(sum n) and sums to one, (4) the two-lane and the one-lane runs of the same
    pool agree bit for bit on every tree, (5) eight trees picked at random match the oracle bit for bit."""
    import os

    T, count, batch = 1024, 800, 16
    runs = []
    for lanes in ("512", "0"):  # 1024 searching trees: two lanes of 512 / one lane
        os.environ["OMK_LANE_MIN_TREES"] = lanes
        try:
            c = omk.Context(device=0, capacity_envs=4, capacity_trees=T, capacity_nodes=2048, seed=31)
        finally:
            del os.environ["OMK_LANE_MIN_TREES"]
        c.pool_new_games(n=T, evaluator=omk.EVAL_HASH)
        c.selfplay_begin(1, 16, 16, 0.0, 1.0, 1.0, 30, omk.EVAL_HASH)  # only to read the simulation counter through stats
        c.pool_new_games(n=T, evaluator=omk.EVAL_HASH)
        c.pool_search(n=T, count=count, batch_size=batch, epsilon=0.25, alpha=0.03, evaluator=omk.EVAL_HASH)
        pol, valid = c.pool_policy(n=T)
        acts, _ = c.pool_sample(n=T)
        per_tree = []
        for t in range(T):
            a, n, w, p = c.pool_root_children(t)
            rn = c.pool_root_stats(t)[0]
            per_tree.append((a.tobytes(), n.tobytes(), w.tobytes(), p.tobytes(), rn, c.pool_tree_info(t)))
            if lanes == "512":
                assert int(n.sum()) == rn == count, f"tree {t}: every simulation backs up through the root exactly once"
                assert valid[t] and abs(float(pol[t].sum()) - 1.0) < 1e-5
                # agent.rs:56-76: pi = n * recip(sum) in f32
                assert np.array_equal(pol[t][a], n.astype(np.float32) * (np.float32(1.0) / np.float32(n.sum())))
                assert int(acts[t]) == int(a[n == n.max()].max())  # Best: last maximum over the 81 cells (agent.rs:98-105)
        runs.append((per_tree, pol.tobytes(), acts.tobytes()))
        if lanes == "0":
            ev = orc.NativeHashEvaluator()
            rng = np.random.default_rng(0)
            for t in rng.choice(T, size=8, replace=False):
                agent = orc.Agent(ev, c.seed, int(t))
                orc.execute([agent], count, batch, 0.25, 0.03, ev)
                assert_tree_equal(c, int(t), agent, f"full-size tree {t}")
        c.close()
    assert runs[0] == runs[1], "two search lanes must not change any tree"


def test_selfplay_driver_ring_matches_granular_calls(omk):
    """The device-resident self-play driver streams its transitions through a ring of four ply slots (pinned host mirror,
    copy stream).  Seven plies (> the ring depth) of the driver must give, ply by ply, the boards / visit policies /
    actions / statuses of the same games played through the granular calls (search, sample, play, ensure_action, play)."""
    G, plies, count, batch, eps, alpha, temp, threshold = 6, 7, 64, 16, 0.25, 0.03, 1.0, 3
    a = omk.Context(device=0, capacity_envs=4, capacity_trees=2 * G, capacity_nodes=2048, seed=77)
    a.selfplay_begin(G, count, batch, eps, alpha, temp, threshold, omk.EVAL_HASH)
    stats, boards, policy, status, actions = a.selfplay_run(plies, profile=0, want_transitions=True)
    a.close()
    assert int(stats.positions) == G * plies and int(stats.d2h_bytes) == G * plies * (81 + 81 * 4 + 4 + 1)
    boards, policy = boards.reshape(plies, G, 81), policy.reshape(plies, G, 81)
    status, actions = status.reshape(plies, G), actions.reshape(plies, G)

    b = omk.Context(device=0, capacity_envs=4, capacity_trees=2 * G, capacity_nodes=2048, seed=77)
    b.pool_new_games(n=2 * G, evaluator=omk.EVAL_HASH)  # black = tree 2g, white = 2g + 1, stream id = tree id
    for ply in range(plies):
        mids = [2 * g + (ply % 2) for g in range(G)]
        oids = [2 * g + 1 - (ply % 2) for g in range(G)]
        b.pool_search(ids=mids, count=count, batch_size=batch, epsilon=eps, alpha=alpha, evaluator=omk.EVAL_HASH)
        before = np.stack([b.pool_get_env(t)[0] for t in mids])
        acts, pol = b.pool_sample(ids=mids, modes=[1 if ply < threshold else 0] * G, temperatures=[temp] * G)
        st = b.pool_play(acts, ids=mids)
        b.pool_ensure_action(acts, ids=oids, evaluator=omk.EVAL_HASH)
        b.pool_play(acts, ids=oids)
        assert [int(x) for x in acts] == [int(x) for x in actions[ply]], f"ply {ply}: actions"
        assert pol.tobytes() == policy[ply].tobytes(), f"ply {ply}: visit policies"
        assert np.array_equal(before, boards[ply]), f"ply {ply}: boards (state the move was chosen in)"
        assert [int(x) for x in st] == [int(x) for x in status[ply]], f"ply {ply}: status"
        assert all(int(x) == 0 for x in st), "the fixture horizon is too short for a game to end"
    b.close()

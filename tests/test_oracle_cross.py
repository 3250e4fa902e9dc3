"""Cross-check of the search oracle (oracle/omok_oracle.c) against a second, independently written restatement of the
reference's Rust (tests/pyref.py): same parameters, same specified random stream, same exact evaluator -> every tree must
agree bit for bit (child creation order, visit counts, w, p, root statistics, policies, stream position, node counts).

The reference holds no test for its mcts / alpha-zero crates and cannot be compiled here, so the search oracle stays
"parity unpinned"; this removes the single-reader risk: the two restatements share no code for the reference's algorithm
(alpha-zero/src/parallel_mcts_executor.rs:80-189,222-265, mcts/src/node.rs:39-99, alpha-zero/src/agent.rs:43-232)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pyref  # noqa: E402  (tests/pyref.py)


def bits(x):
    return np.asarray(x, dtype=np.float32).tobytes()


def assert_same_tree(py: "pyref.Agent", c, where=""):
    root = py.mcts.root
    a, n, w, p = c.root_children()
    assert [ch.action for ch in root.children] == [int(x) for x in a], f"{where}: children / creation order"
    assert [ch.n for ch in root.children] == [int(x) for x in n], f"{where}: visit counts"
    assert bits([ch.w for ch in root.children]) == w.tobytes(), f"{where}: w"
    assert bits([ch.p for ch in root.children]) == p.tobytes(), f"{where}: p"
    rn, rw, rp, rst, rpol = c.root_stats()
    assert (root.n, root.state.status) == (rn, rst), f"{where}: root n / status"
    assert bits(root.w) == bits(rw) and bits(root.p) == bits(rp), f"{where}: root w / p"
    assert bits(root.state.policy) == rpol.tobytes(), f"{where}: root policy"
    assert py.rng.counter == c.rng_counter, f"{where}: random-stream position"
    assert py.mcts.node_count() == c.node_count, f"{where}: node count"
    assert py.env.board == [int(x) for x in c.board()] and py.env.turn == c.env.turn
    assert py.env.legal_move_count == c.env.legal_move_count


@pytest.mark.parametrize("count,batch,eps,alpha", [(800, 16, 0.0, 1.0), (600, 16, 0.25, 0.03), (50, 1, 0.25, 0.3), (128, 64, 0.5, 1.0)])
def test_search_agrees_with_the_c_oracle(orc, count, batch, eps, alpha):
    """The golden-fixture parameter sets (tests/golden/make_golden.py) on two trees each."""
    seed = 2024
    ev = orc.NativeHashEvaluator()
    streams = (100, 101)
    c_agents = [orc.Agent(ev, seed, s) for s in streams]
    py_agents = [pyref.Agent(pyref.hash_evaluate_p, seed, s) for s in streams]
    for p, c in zip(py_agents, c_agents):
        assert_same_tree(p, c, "new")
    orc.execute(c_agents, count, batch, eps, alpha, ev)
    pyref.execute(py_agents, count, batch, eps, alpha, pyref.hash_evaluate_pv)
    rounds = -(-count // batch)
    for p, c in zip(py_agents, c_agents):
        assert_same_tree(p, c, f"after {count}/{batch}")
        assert p.mcts.root.n == rounds * batch  # sims per move = ceil(count / b) * b (parallel_mcts_executor.rs:39-42,207)


def test_self_play_sequence_agrees_with_the_c_oracle(orc):
    """Trainer-shaped play (src/trainer.rs:86-204): two agents per game, search, Boltzmann(1.0) then Best, play_action on
    the mover, ensure_action_exists + play_action on the other agent; compared after every step for several plies."""
    seed, count, batch, eps, alpha, threshold = 7, 64, 16, 0.25, 0.03, 3
    ev = orc.NativeHashEvaluator()
    c_pair = [orc.Agent(ev, seed, 0), orc.Agent(ev, seed, 1)]
    py_pair = [pyref.Agent(pyref.hash_evaluate_p, seed, 0), pyref.Agent(pyref.hash_evaluate_p, seed, 1)]
    for ply in range(9):
        m, o = ply % 2, 1 - ply % 2
        orc.execute([c_pair[m]], count, batch, eps, alpha, ev)
        pyref.execute([py_pair[m]], count, batch, eps, alpha, pyref.hash_evaluate_pv)
        assert_same_tree(py_pair[m], c_pair[m], f"ply {ply} searched")
        if ply < threshold:
            ca, cpol = c_pair[m].sample_action(1, 1.0)
            pa, ppol = py_pair[m].sample_action(1.0)
        else:
            ca, cpol = c_pair[m].sample_action(0)
            pa, ppol = py_pair[m].sample_action(None)
        assert pa == ca and bits(ppol) == cpol.tobytes(), f"ply {ply}: sampled action / visit policy"
        assert py_pair[m].play_action(pa) == c_pair[m].play_action(ca)
        c_pair[o].ensure_action_exists(ca, ev)
        py_pair[o].ensure_action_exists(pa, pyref.hash_evaluate_p)
        assert py_pair[o].play_action(pa) == c_pair[o].play_action(ca)
        assert_same_tree(py_pair[m], c_pair[m], f"ply {ply} mover re-rooted")
        assert_same_tree(py_pair[o], c_pair[o], f"ply {ply} other re-rooted")


def test_edge_semantics_agree(orc):
    """ensure_action_exists on an occupied cell / an existing child / out of range, play_action of an absent action, the
    epsilon = 0 noise pass (pure renormalisation), compute_policy == None before any visit."""
    seed = 3
    ev = orc.NativeHashEvaluator()
    c, p = orc.Agent(ev, seed, 5), pyref.Agent(pyref.hash_evaluate_p, seed, 5)
    assert p.compute_policy() is None and c.compute_policy() is None
    assert p.play_action(40) is None and c.play_action(40) is None
    for a in (40, 40, 81):  # new child, existing child (no-op), out of range (no-op)
        c.ensure_action_exists(a, ev)
        p.ensure_action_exists(a, pyref.hash_evaluate_p)
        assert_same_tree(p, c, f"ensure {a}")
    assert p.compute_policy() is None and c.compute_policy() is None  # child without visits: sum < EPSILON
    assert p.play_action(40) == c.play_action(40) == 0
    c.ensure_action_exists(40, ev)  # occupied cell: the child's env is the root's env, unchanged (agent.rs:154-155)
    p.ensure_action_exists(40, pyref.hash_evaluate_p)
    assert_same_tree(p, c, "ensure on an occupied cell")
    assert p.play_action(40) is None and c.play_action(40) is None  # place_stone refuses
    orc.execute([c], 32, 8, 0.0, 1.0, ev)
    pyref.execute([p], 32, 8, 0.0, 1.0, pyref.hash_evaluate_pv)
    assert_same_tree(p, c, "search from a root with an illegal child")


def _scripted(moves, orc, seed):
    """Both restatements' agent pairs after a scripted opening (ensure_action_exists + play_action, as the trainer feeds the
    opponent's moves in: src/trainer.rs:147-167)."""
    ev = orc.NativeHashEvaluator()
    c_pair = [orc.Agent(ev, seed, 0), orc.Agent(ev, seed, 1)]
    py_pair = [pyref.Agent(pyref.hash_evaluate_p, seed, 0), pyref.Agent(pyref.hash_evaluate_p, seed, 1)]
    for mv in moves:
        for c, p in zip(c_pair, py_pair):
            c.ensure_action_exists(mv, ev)
            p.ensure_action_exists(mv, pyref.hash_evaluate_p)
            assert p.play_action(mv) == c.play_action(mv) == 0
    return ev, c_pair, py_pair


@pytest.mark.parametrize("which,keep", [("draw", 66), ("draw", 75), ("win81", 70), ("win81", 77)])
def test_late_game_search_agrees_with_the_c_oracle(orc, which, keep):
    """Late positions, where terminal children (backed up at once, parallel_mcts_executor.rs:177-181), terminal leaves
    (:92-97), nodes that fill up after a handful of expansions and roots with fewer empty cells than a round has
    simulations are the rule: the boards of the reference-derived draw / win-on-move-81 constructions (test_oracle_env.py)
    up to `keep` stones, then searched and played to the end by both restatements, compared after every step."""
    from test_oracle_env import draw_moves, win81_moves

    moves = (draw_moves() if which == "draw" else win81_moves())[:keep]
    ev, c_pair, py_pair = _scripted(moves, orc, 11)
    count, batch, eps, alpha = 48, 16, 0.25, 0.3
    ply, status = keep, 0
    while status == 0 and ply < 81:
        m, o = ply % 2, 1 - ply % 2
        orc.execute([c_pair[m]], count, batch, eps, alpha, ev)
        pyref.execute([py_pair[m]], count, batch, eps, alpha, pyref.hash_evaluate_pv)
        assert_same_tree(py_pair[m], c_pair[m], f"{which}/{keep} ply {ply} searched")
        ca, cpol = c_pair[m].sample_action(0)
        pa, ppol = py_pair[m].sample_action(None)
        assert pa == ca and bits(ppol) == cpol.tobytes(), f"ply {ply}: action / visit policy"
        status = c_pair[m].play_action(ca)
        assert py_pair[m].play_action(pa) == status
        c_pair[o].ensure_action_exists(ca, ev)
        py_pair[o].ensure_action_exists(pa, pyref.hash_evaluate_p)
        assert py_pair[o].play_action(pa) == c_pair[o].play_action(ca)
        assert_same_tree(py_pair[m], c_pair[m], f"ply {ply} mover re-rooted")
        assert_same_tree(py_pair[o], c_pair[o], f"ply {ply} other re-rooted")
        ply += 1
    assert status != 0, "the game must end on a board with at most fifteen empty cells"


def test_environment_agrees_on_random_playouts(orc):
    rng = np.random.default_rng(0)
    for game in range(30):
        pe, ce = pyref.Environment(), orc.Environment()
        while True:
            a = int(rng.integers(0, 81))
            ps, cs = pe.place_stone(a), ce.place_stone(a)
            assert ps == cs and pe.board == [int(x) for x in ce.board] and pe.turn == ce.turn
            assert pe.legal_move_count == ce.legal_move_count
            if ps is not None and ps != 0:
                break


# ---- the network oracle against a second restatement (tests/pyref_net.py: numpy, explicit tap / channel loops) ----
def _random_net_params(seed):
    from oracle import net_oracle

    rng = np.random.default_rng(seed)
    params = net_oracle.random_params(seed)
    for i, (name, shape) in enumerate(net_oracle.PARAM_SPECS):  # the recipe's biases are zero: give every bias a value
        if len(shape) == 1:
            params[i] = rng.normal(0.0, 0.3, size=shape).astype(np.float32)
    return params


def _random_positions(n, seed):
    rng = np.random.default_rng(seed)
    boards = np.zeros((n, 81), np.uint8)
    turns = np.zeros(n, np.uint8)
    for b in range(n):
        k = int(rng.integers(0, 81))
        for j, c in enumerate(rng.permutation(81)[:k]):
            boards[b, c] = 1 + (j % 2)
        turns[b] = int(rng.integers(0, 2))
    return boards, turns


def test_input_image_agrees_with_second_restatement():
    import pyref_net
    from oracle import net_oracle

    boards, turns = _random_positions(40, 5)
    for b, t in zip(boards, turns):
        for opp in (False, True):
            a = net_oracle.encode_image(b, int(t), opponent_mode=opp)
            r = pyref_net.encode_nn_input(b, black_to_move=(t == 0), opponent_mode=opp)
            assert a.astype(np.float64).tobytes() == r.tobytes()


@pytest.mark.parametrize("seed", [0, 3])
def test_network_forward_agrees_with_second_restatement(seed):
    """Every layer of oracle/net_oracle.py (torch conv2d graph, fp64) against the numpy restatement written from the Rust
    graph builder and TensorFlow's op semantics: stem, the three bottleneck blocks (depthwise filter layout [kh][kw][c],
    SAME padding, one bias after the pointwise conv, residual add before the last leaky-relu), NHWC flatten order, fc0, fc1,
    both heads.  fp64 on both sides: agreement to rounding."""
    import torch

    import pyref_net
    from oracle import net_oracle

    params = _random_net_params(seed)
    boards, turns = _random_positions(10, 100 + seed)
    imgs = np.stack([net_oracle.encode_image(b, int(t), opponent_mode=bool(i % 2)) for i, (b, t) in enumerate(zip(boards, turns))])
    a = net_oracle.forward_layers(params, imgs, dtype=torch.float64)
    r = pyref_net.forward(params, imgs)
    for k in ("tower", "fc0", "fc1", "logits", "vlogit", "P", "V"):
        scale = max(np.abs(r[k]).max(), 1e-30)
        assert np.abs(np.asarray(a[k], np.float64) - r[k]).max() <= 1e-11 * scale, k
    # the fp32 forward the GPU tests compare against is the same graph
    p32, v32, lg32 = net_oracle.forward(params, imgs)
    assert np.abs(p32 - r["P"]).max() <= 2e-4 * r["P"].max() and np.abs(v32 - r["V"]).max() <= 1e-4

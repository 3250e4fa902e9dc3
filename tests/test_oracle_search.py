"""Known-answer / property tests for the oracle's tree search.  The reference holds
no test for mcts / alpha-zero, so these are derived from the code."""
import numpy as np


def test_f32_epsilon_facts():
    # compute_ucb_1 (parallel_mcts_executor.rs:282): n as f32 + EPSILON changes n==0 and n==1 only
    eps = np.float32(1.1920929e-7)
    assert np.float32(1) + eps != np.float32(1)
    for n in (2, 3, 100, 799):
        assert np.float32(n) + eps == np.float32(n)


def test_round_count_and_root_visits(orc):
    ev = orc.HashEvaluator()
    for count, b in ((800, 8), (800, 16), (600, 16), (5, 4)):
        a = orc.Agent(ev, seed=7, stream=1)
        sims = orc.execute([a], count, b, 0.0, 1.0, ev)
        assert sims == -(-count // b) * b  # ceil(count/b)*b  (parallel_mcts_executor.rs:39-42,207)
        n, w, p, st, pol = a.root_stats()
        assert n == sims  # every simulation back-propagates through the root exactly once
        acts, cn, cw, cp = a.root_children()
        assert cn.sum() == sims  # and through exactly one root child
        assert len(set(acts.tolist())) == len(acts)


def test_noise_pass_renormalises_with_epsilon_zero(orc):
    ev = orc.HashEvaluator()
    a = orc.Agent(ev, seed=1, stream=0)
    raw = a.root_stats()[4]
    assert raw.sum() > 10  # the fake net is unnormalised
    orc.execute([a], 8, 8, 0.0, 1.0, ev)
    pol = a.root_stats()[4]
    assert abs(pol.sum() - 1) < 1e-5
    s = np.float32(0)
    for x in raw:
        s = np.float32(s + x)
    assert np.array_equal(pol, raw * (np.float32(1) / s))


def test_last_max_tie_break_prefers_last_created(orc):
    # uniform priors and V=0 -> all root children tie at every visit count; max_by keeps the LAST maximum
    ev = orc.CallbackEvaluator(lambda b, t, m: (np.full((len(b), 81), 1 / 81, np.float32), np.zeros(len(b), np.float32)))
    a = orc.Agent(ev, seed=3, stream=5)
    orc.execute([a], 81, 1, 0.0, 1.0, ev)  # 81 sims, one per round: root becomes fully expanded
    acts, cn, cw, cp = a.root_children()
    assert len(acts) == 81 and (cn == 1).all()
    orc.execute([a], 1, 1, 0.0, 1.0, ev)  # next sim must descend into the last-created child
    acts2, cn2, _, _ = a.root_children()
    assert cn2[-1] == 2 and (cn2[:-1] == 1).all()


def test_transition_sets_root_n_to_children_sum(orc):
    ev = orc.HashEvaluator()
    a = orc.Agent(ev, seed=11, stream=2)
    orc.execute([a], 800, 16, 0.0, 1.0, ev)
    act, pol = a.sample_action(0)
    assert abs(pol.sum() - 1) < 1e-5
    nodes_before = a.node_count
    assert a.play_action(act) == 0
    n, *_ = a.root_stats()
    _, cn, _, _ = a.root_children()
    assert n == cn.sum()
    assert a.node_count < nodes_before
    assert a.play_action(act) is None  # occupied now / not in tree


def test_thread_count_does_not_change_results(orc):
    ev = orc.HashEvaluator()
    res = []
    for nt in (1, 4):
        agents = [orc.Agent(ev, seed=5, stream=i) for i in range(6)]
        orc.execute(agents, 64, 16, 0.25, 0.03, ev, n_threads=nt)
        res.append([tuple(map(bytes, a.root_children())) for a in agents])
    assert res[0] == res[1]


def test_dirichlet_is_a_distribution(orc):
    import ctypes as C

    out = np.zeros(81, np.float32)
    for alpha in (0.03, 0.3, 1.0, 2.5):
        for epoch in range(3):
            orc.lib().orc_dirichlet81(9, 4, epoch, alpha, out.ctypes.data_as(C.POINTER(C.c_float)))
            assert np.isfinite(out).all() and (out >= 0).all() and abs(out.sum() - 1) < 1e-4
    # moments of Gamma via many Dirichlet draws are loose; check det_log/det_exp accuracy instead
    for x in (1e-9, 0.3, 1.0, 2.0, 77.7, 1e12):
        assert abs(orc.lib().orc_det_log(x) - np.log(x)) <= 4e-16 * max(1, abs(np.log(x)))
    for x in (-700.0, -30.5, -1.0, 0.0, 0.5, 1.0, 33.3, 700.0):
        assert abs(orc.lib().orc_det_exp(x) / np.exp(x) - 1) < 1e-15


def test_self_play_game_terminates(orc):
    ev = orc.HashEvaluator()
    black, white = orc.Agent(ev, 1, 0), orc.Agent(ev, 1, 1)
    status, plies = 0, 0
    while status == 0:
        mover, other = (black, white) if plies % 2 == 0 else (white, black)
        orc.execute([mover], 64, 16, 0.25, 0.03, ev)
        act, _ = mover.sample_action(1 if plies < 30 else 0, 1.0)
        status = mover.play_action(act)
        other.ensure_action_exists(act, ev)
        assert other.play_action(act) == status or status != 0  # ensure_action marks InProgress always
        plies += 1
    assert status in (1, 2, 3) and plies <= 81
    assert np.array_equal(black.board(), white.board())

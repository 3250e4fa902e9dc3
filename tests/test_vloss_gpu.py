"""Virtual loss (north_star: "PUCT selection with virtual loss") is an OPT-IN search mode with no counterpart in the
reference (alpha-zero/src/parallel_mcts_executor.rs:80-90 reads the statistics as they are), so it has no parity target:
property tests only.  Default off; refused with the parity evaluator."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_virtual_loss_conserves_visits_and_leaves_no_residue(omk):
    T, count, batch = 12, 320, 16
    stats = {}
    for on in (False, True):
        ctx = omk.Context(device=0, capacity_envs=1, capacity_trees=T, capacity_nodes=2048, seed=5)
        ctx.net_init_random(3)
        ctx.search_set_virtual_loss(on)
        ctx.pool_new_games(n=T, evaluator=omk.EVAL_NET)
        ctx.pool_search(n=T, count=count, batch_size=batch, epsilon=0.25, alpha=0.03, evaluator=omk.EVAL_NET)
        spread = []
        for t in range(T):
            a, n, w, p = ctx.pool_root_children(t)
            root_n = ctx.pool_root_stats(t)[0]
            assert root_n == count == int(n.sum()), "every simulation is backed up through the root exactly once"
            assert np.all(np.abs(w) <= n.astype(np.float32) + 1e-3), "|w| <= n: no virtual loss is left behind (values are in [-1, 1])"
            spread.append(ctx.pool_tree_info(t)[0])
        pol, valid = ctx.pool_policy(n=T)
        assert valid.all() and np.allclose(pol.sum(1), 1, atol=1e-5)
        acts, _ = ctx.pool_sample(n=T)
        assert (ctx.pool_play(acts) == 0).all()
        stats[on] = spread
        ctx.close()
    # same number of expansions either way (one node per non-terminal simulation) ...
    assert stats[True] == stats[False] or abs(sum(stats[True]) - sum(stats[False])) <= T * 4


def test_virtual_loss_spreads_a_round_and_is_refused_with_the_parity_evaluator(omk):
    """One round of 16 simulations from a fully expanded root: without virtual loss the reference's rule sends all 16 to the
    same child (statistics cannot change inside a round); with it they spread over several children."""
    def distinct_children_visited_in_last_round(on):
        ctx = omk.Context(device=0, capacity_envs=1, capacity_trees=4, capacity_nodes=2048, seed=9)
        ctx.net_init_random(1)
        ctx.pool_new_games(n=4, evaluator=omk.EVAL_NET)
        ctx.pool_search(n=4, count=96, batch_size=16, epsilon=0.0, alpha=1.0, evaluator=omk.EVAL_NET)  # fills the root (81 children)
        before = [ctx.pool_root_children(t)[1].copy() for t in range(4)]
        ctx.search_set_virtual_loss(on)
        ctx.pool_search(n=4, count=16, batch_size=16, epsilon=0.0, alpha=1.0, evaluator=omk.EVAL_NET)
        after = [ctx.pool_root_children(t)[1] for t in range(4)]
        ctx.close()
        return [int(np.count_nonzero(a - b)) for a, b in zip(after, before)]

    off, on = distinct_children_visited_in_last_round(False), distinct_children_visited_in_last_round(True)
    assert all(k == 1 for k in off), off
    assert all(k > 1 for k in on), on
    ctx = omk.Context(device=0, capacity_envs=1, capacity_trees=2, capacity_nodes=256, seed=1)
    ctx.search_set_virtual_loss(True)
    with pytest.raises(omk.OmkError) as e:
        ctx.pool_new_games(n=2, evaluator=omk.EVAL_HASH)
    assert e.value.code == -4
    ctx.search_set_virtual_loss(False)
    ctx.pool_new_games(n=2, evaluator=omk.EVAL_HASH)
    ctx.close()


def test_self_play_driver_runs_with_virtual_loss(omk):
    """The device-resident driver with the opt-in mode: positions, simulation counts and the transition stream stay
    consistent (every policy sums to one and is zero on occupied cells)."""
    G, plies = 16, 6
    ctx = omk.Context(device=0, capacity_envs=4, capacity_trees=2 * G, capacity_nodes=2048, seed=2)
    ctx.net_init_random(0)
    ctx.search_set_virtual_loss(True)
    ctx.selfplay_begin(G, 64, 16, 0.25, 0.03, 1.0, 30, omk.EVAL_NET)
    stats, boards, policy, status, actions = ctx.selfplay_run(plies, profile=0, want_transitions=True)
    assert int(stats.positions) == G * plies and int(stats.simulations) == G * plies * 64
    assert np.allclose(policy.sum(-1), 1.0, atol=1e-5) and (policy[boards != 0] == 0).all()
    assert (status >= 0).all()
    ctx.close()

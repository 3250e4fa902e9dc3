"""Self-play driver (src/trainer.rs:86-204 restated on the device) with the real network: weight changes between
self-play phases, and the streamed transitions of a long run."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_restarted_games_use_the_current_weights(omk):
    """Agent::new evaluates the empty board with the CURRENT weights (alpha-zero/src/agent.rs:16-25).  The driver caches
    that prior for restarted games, so the cache must follow omk_net_load_params / omk_net_init_random (the trainer's
    weight refresh between self-play phases).  Load other weights after omk_selfplay_begin, play until a game ends, and
    compare the restarted trees' root prior with omk_net_eval of the empty board under the new weights."""
    from oracle import net_oracle

    G = 6
    ctx = omk.Context(device=0, capacity_envs=4, capacity_trees=2 * G, capacity_nodes=1024, seed=9)
    ctx.net_load_params(net_oracle.random_params(0))
    empty = np.zeros((1, 81), np.uint8), np.zeros(1, np.uint8)
    p_old, _ = ctx.net_eval(*empty)
    ctx.selfplay_begin(G, 16, 8, 0.25, 0.03, 1.0, 30, omk.EVAL_NET)
    ctx.net_load_params(net_oracle.random_params(1))  # "trainer step": new weights while the driver is active
    p_new, _ = ctx.net_eval(*empty)
    assert not np.array_equal(p_old, p_new)
    finished = None
    for _ in range(81):
        _, _, _, status, _ = ctx.selfplay_run(1, profile=0, want_transitions=True)
        done = np.flatnonzero(status[0] > 0)
        if done.size:
            finished = int(done[0])
            break
    assert finished is not None, "no game ended within 81 plies"
    for tree in (2 * finished, 2 * finished + 1):  # both colours' trees were restarted
        board, turn, legal = ctx.pool_get_env(tree)
        assert legal == 81 and turn == 0 and not board.any()
        pol = ctx.pool_root_stats(tree)[4]
        assert pol.tobytes() == p_new[0].tobytes(), "restarted root prior must come from the current network"
    ctx.net_init_random(5)  # the other weight-changing entry point refreshes the cache too
    p_rand, _ = ctx.net_eval(*empty)
    for _ in range(200):
        _, _, _, status, _ = ctx.selfplay_run(1, profile=0, want_transitions=True)
        done = np.flatnonzero(status[0] > 0)
        if done.size:
            pol = ctx.pool_root_stats(2 * int(done[0]))[4]
            assert pol.tobytes() == p_rand[0].tobytes()
            break
    else:
        pytest.fail("no game ended after the second weight change")
    # the trainer's own entry point: omk_train_step marks the tensor-core operand images stale and the NEXT evaluation --
    # here the driver itself, with no omk_net_eval in between -- rebuilds them together with the cached prior
    rng = np.random.default_rng(3)
    images = (rng.random((32, 243)) < 0.3).astype(np.float32)
    pi = rng.random((32, 81)).astype(np.float32)
    pi /= pi.sum(axis=1, keepdims=True)
    z = rng.choice([-1.0, 1.0], size=32).astype(np.float32)
    for _ in range(3):
        ctx.train_step(images, pi, z)
    for _ in range(200):
        _, _, _, status, _ = ctx.selfplay_run(1, profile=0, want_transitions=True)
        done = np.flatnonzero(status[0] > 0)
        if done.size:
            pol = ctx.pool_root_stats(2 * int(done[0]))[4]
            p_trained, _ = ctx.net_eval(*empty)
            assert not np.array_equal(p_trained, p_rand)
            assert pol.tobytes() == p_trained[0].tobytes(), "restarted root prior must follow omk_train_step"
            break
    else:
        pytest.fail("no game ended after the training steps")
    ctx.close()


def test_streamed_transitions_replay_on_the_environment(omk, orc):
    """BASELINE config 4's stream: every transition (board the move was chosen in, visit policy, action, status after the
    move) of a 40-ply run over 64 games is consistent with the reference rules replayed on the oracle environment: the
    board of ply t+1 is the board of ply t plus the action (or the empty board after a terminal status), the status is
    place_stone's, the visit policy sums to one and is zero on occupied cells."""
    G, plies = 64, 40
    ctx = omk.Context(device=0, capacity_envs=4, capacity_trees=2 * G, capacity_nodes=1024, seed=4)
    ctx.selfplay_begin(G, 32, 16, 0.25, 0.03, 1.0, 6, omk.EVAL_HASH)
    stats, boards, policy, status, actions = ctx.selfplay_run(plies, profile=0, want_transitions=True)
    ctx.close()
    assert int(stats.positions) == G * plies
    assert int(stats.d2h_bytes) == G * plies * (81 + 81 * 4 + 4 + 1)
    envs = [orc.Environment() for _ in range(G)]
    ended = 0
    for t in range(plies):
        for g in range(G):
            assert np.array_equal(boards[t, g], envs[g].board), f"ply {t} game {g}: board"
            a = int(actions[t, g])
            assert abs(float(policy[t, g].sum()) - 1.0) < 1e-5 and policy[t, g][boards[t, g] != 0].max(initial=0.0) == 0.0
            assert policy[t, g][a] > 0.0
            st = envs[g].place_stone(a)
            assert st == int(status[t, g]), f"ply {t} game {g}: status"
            if st != 0:
                envs[g] = orc.Environment()  # finished games restart
                ended += 1
    assert ended == int(stats.games_finished)

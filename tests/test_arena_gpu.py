"""the naive player (a pure environment-kernel workload) and batched evaluation play."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def arena():
    return importlib.import_module("omok-ai_b200.arena")


def test_naive_player_wins_blocks_then_plays_at_random(omk, arena):
    ctx = omk.Context(device=0, capacity_envs=162 * 4, capacity_trees=4, capacity_nodes=64, seed=0)
    boards = np.zeros((4, 81), np.uint8)
    turns = np.zeros(4, np.uint8)
    boards[0, [0, 1, 2, 3]] = 1          # black to move with four in a row: completes five at cell 4
    boards[0, [9, 10, 11]] = 2
    boards[1, [20, 21, 22, 23]] = 1      # white to move: must block; the first blocking cell in ascending order is 19
    boards[1, [40, 41, 50]] = 2
    turns[1] = 1
    boards[2, [0, 1, 2, 3, 5]] = 1       # cell 4 would make SIX for black: not a win (lib.rs:151-154) -> no forced move
    boards[2, [30, 31, 32, 33]] = 2      # ... but white's four is a threat black must answer: cells 29 / 34 -> 29
    boards[2, 60] = 2
    boards[3, 40] = 1                    # nothing to win or block: a uniformly random legal move
    turns[3] = 1
    rng = np.random.default_rng(0)
    got = arena.naive_moves(ctx, boards, turns, rng)
    assert got[:3].tolist() == [4, 19, 29]
    assert boards[3, got[3]] == 0
    seen = {int(arena.naive_moves(ctx, boards[3:], turns[3:], rng)[0]) for _ in range(12)}
    assert len(seen) > 3 and 40 not in seen
    ctx.close()


def test_batched_evaluation_play_accounts_for_every_game(omk, arena):
    games = 8
    ctx = omk.Context(device=0, capacity_envs=162 * games, capacity_trees=games, capacity_nodes=1024, seed=1)
    b, w, d = arena.play_against_naive_player(ctx, episode_count=games, count=48, batch_size=16, evaluator=omk.EVAL_HASH, seed=2)
    assert b + w + d == games
    left = omk.Context(device=0, capacity_envs=1, capacity_trees=games // 2, capacity_nodes=1024, seed=3)
    right = omk.Context(device=0, capacity_envs=1, capacity_trees=games // 2, capacity_nodes=1024, seed=4)
    lw, rw, dr = arena.play_match(left, right, game_count=games, count=32, batch_size=8, evaluator=omk.EVAL_HASH)
    assert lw + rw + dr == games
    for c in (ctx, left, right):
        c.close()


def _oracle_naive_move(orc, env, rng):
    """trainer.rs:506-537 restated on the oracle's environment: the first legal action (ascending) that ends the game for
    the mover, or that would end it for the opponent; otherwise a uniformly random legal move."""
    import ctypes as C

    def clone(flip_turn):
        e = orc.Environment()
        C.memmove(C.byref(e.e), C.byref(env), C.sizeof(env))
        if flip_turn:
            e.e.turn = 1 - e.e.turn
        return e

    legal = [a for a in range(81) if env.board[a] == 0]
    for a in legal:
        if clone(False).place_stone(a) != 0 or clone(True).place_stone(a) != 0:  # GameStatus::is_terminal
            return a
    return legal[rng.integers(0, len(legal))]


def test_play_against_naive_player_matches_oracle_loop(omk, orc, arena):
    """The whole evaluation loop (src/trainer.rs:487-603) against the oracle: same W/L/D and the same move lists, game by
    game, with the exact hash evaluator and the same generator for the naive player's random moves."""
    games, count, batch, eps, alpha, seed, rseed = 10, 64, 16, 0.25, 0.03, 11, 5
    ctx = omk.Context(device=0, capacity_envs=162 * games, capacity_trees=games, capacity_nodes=2048, seed=seed)
    moves = []
    got = arena.play_against_naive_player(ctx, episode_count=games, count=count, batch_size=batch, epsilon=eps, alpha=alpha,
                                          evaluator=omk.EVAL_HASH, seed=rseed, moves_out=moves)
    ctx.close()
    ev = orc.NativeHashEvaluator()
    rng = np.random.default_rng(rseed)
    agents = {g: orc.Agent(ev, seed, g) for g in range(games)}
    want_moves = [[] for _ in range(games)]
    tally = {1: 0, 2: 0, 3: 0}
    live = list(range(games))
    while live:
        nxt = []
        for g in live:
            a = _oracle_naive_move(orc, agents[g].env, rng)
            want_moves[g].append(a)
            agents[g].ensure_action_exists(a, ev)
            st = agents[g].play_action(a)
            if st == 0:
                nxt.append(g)
            else:
                tally[st] += 1
        live = nxt
        if not live:
            break
        orc.execute([agents[g] for g in live], count, batch, eps, alpha, ev)
        nxt = []
        for g in live:
            a, _ = agents[g].sample_action(0)
            want_moves[g].append(a)
            st = agents[g].play_action(a)
            if st == 0:
                nxt.append(g)
            else:
                tally[st] += 1
        live = nxt
    assert moves == want_moves
    assert got == (tally[2], tally[3], tally[1])


def test_play_match_matches_oracle_loop(omk, orc, arena):
    """benchmark/src/main.rs:14-108 against the oracle: two agents per game (one per model), MCTSExecutor::run(count, batch,
    eps 0, alpha 1) + Best for the mover, ensure_action_exists + play_action for the other; move lists and result."""
    games, count, batch = 6, 48, 8
    half = games // 2
    seeds = (21, 22)
    left = omk.Context(device=0, capacity_envs=1, capacity_trees=half, capacity_nodes=2048, seed=seeds[0])
    right = omk.Context(device=0, capacity_envs=1, capacity_trees=half, capacity_nodes=2048, seed=seeds[1])
    moves = []
    got = arena.play_match(left, right, game_count=games, count=count, batch_size=batch, evaluator=omk.EVAL_HASH, moves_out=moves)
    left.close()
    right.close()
    ev = orc.NativeHashEvaluator()
    want_moves = []
    wins = {0: 0, 1: 0, "draw": 0}  # 0 = left, 1 = right
    for leg in range(2):
        first, second = (0, 1) if leg == 0 else (1, 0)
        ag = {side: [orc.Agent(ev, seeds[side], g) for g in range(half)] for side in (0, 1)}
        logs = [[] for _ in range(half)]
        live = list(range(half))
        mover, other = first, second
        while live:
            orc.execute([ag[mover][g] for g in live], count, batch, 0.0, 1.0, ev)
            nxt = []
            for g in live:
                a, _ = ag[mover][g].sample_action(0)
                logs[g].append(a)
                st = ag[mover][g].play_action(a)
                if st >= 2:
                    wins[mover] += 1
                elif st == 1:
                    wins["draw"] += 1
                else:
                    ag[other][g].ensure_action_exists(a, ev)
                    assert ag[other][g].play_action(a) == 0
                    nxt.append(g)
            live = nxt
            mover, other = other, mover
        want_moves += logs
    assert moves == want_moves
    assert got == (wins[0], wins[1], wins["draw"])


def test_pool_get_envs_matches_single_tree_reads(omk):
    ctx = omk.Context(device=0, capacity_envs=1, capacity_trees=16, capacity_nodes=256, seed=3)
    ctx.pool_new_games(n=16, evaluator=omk.EVAL_HASH)
    ctx.pool_search(n=16, count=32, batch_size=8, epsilon=0.0, alpha=1.0, evaluator=omk.EVAL_HASH)
    acts, _ = ctx.pool_sample(n=16, modes=np.zeros(16, np.uint8))
    ctx.pool_play(acts)
    ids = np.array([5, 0, 15, 7], np.int32)
    boards, turns, legal, status = ctx.pool_get_envs(ids=ids)
    for k, t in enumerate(ids):
        b, tu, lg = ctx.pool_get_env(int(t))
        assert np.array_equal(boards[k], b) and turns[k] == tu and legal[k] == lg and status[k] == 0
        assert b[acts[t]] == 1 and lg == 80
    with pytest.raises(omk.OmkError):
        ctx.pool_get_envs(ids=[1, 1])  # ids of one call must be unique
    ctx.close()

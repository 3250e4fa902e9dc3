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

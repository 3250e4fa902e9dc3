"""Evaluation play, batched over games on the tree / env pools of `binding.Context`.

  * `play_match`                 benchmark/src/main.rs:14-108 + benchmark/src/agent.rs:14-47: two models, GAME_COUNT games,
                                 half with each side moving first; every move = MCTSExecutor::run(800, 8, eps 0, alpha 1)
                                 then `sample_action(Best)`; the opponent follows with ensure_action_exists + play_action.
  * `play_against_naive_player`  src/trainer.rs:487-603: the naive player takes the first legal move (ascending) that ends
                                 the game for itself, or that would end it for the opponent (a block); else a random
                                 legal move.  The model answers with ParallelMCTSExecutor::execute + Best.

The reference plays its games one after another; games are independent, so all of them run concurrently here (one tree
per game and model), which changes nothing for any single game.  The naive player's "does this move end the game?" test
is the environment kernel itself: every (game, empty cell, perspective) candidate is one board of the env pool, stepped
in a single `omk_env_step` call -- a pure K1 workload.
"""
from __future__ import annotations

import numpy as np

from . import binding as B

CELLS = 81


def _play_moves(ctx, ids, actions, evaluator):
    """ensure_action_exists + play_action on the trees `ids` (the opponent's move arriving in this model's trees)."""
    ctx.pool_ensure_action(actions, ids=ids, evaluator=evaluator)
    return ctx.pool_play(actions, ids=ids)


def play_match(ctx_left, ctx_right, game_count: int = 100, count: int = 800, batch_size: int = 8, epsilon: float = 0.0,
               alpha: float = 1.0, evaluator: int = B.EVAL_NET, moves_out: list | None = None):
    """Returns (left_wins, right_wins, draws).  Each context holds one model; it needs `game_count // 2` trees.
    `moves_out`, if given, receives one move list per game: first the `half` games left opens, then those right opens."""
    half = game_count // 2
    wins = {"left": 0, "right": 0, "draw": 0}
    log = [[] for _ in range(2 * half)]
    for leg, (first, second, first_name, second_name) in enumerate(((ctx_left, ctx_right, "left", "right"), (ctx_right, ctx_left, "right", "left"))):
        ids = np.arange(half, dtype=np.int32)
        first.pool_new_games(ids=ids, evaluator=evaluator)
        second.pool_new_games(ids=ids, evaluator=evaluator)
        live = ids.copy()
        mover, other, mover_name, other_name = first, second, first_name, second_name
        while live.size:
            mover.pool_search(ids=live, count=count, batch_size=batch_size, epsilon=epsilon, alpha=alpha, evaluator=evaluator)
            actions, _ = mover.pool_sample(ids=live, modes=np.full(live.size, B.SAMPLE_BEST, np.uint8))
            for g, a in zip(live, actions):
                log[leg * half + int(g)].append(int(a))
            status = mover.pool_play(actions, ids=live)
            # BlackWin / WhiteWin: the player who just moved made five (main.rs:61-75 maps it to the side to move first)
            done = status != 0
            wins[mover_name] += int(np.count_nonzero(status >= 2))
            wins["draw"] += int(np.count_nonzero(status == 1))
            keep = ~done
            if keep.any():
                _play_moves(other, live[keep], actions[keep], evaluator)
            live = live[keep]
            mover, other, mover_name, other_name = other, mover, other_name, mover_name
    if moves_out is not None:
        moves_out[:] = log
    return wins["left"], wins["right"], wins["draw"]


def naive_moves(ctx, boards: np.ndarray, turns: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    """The naive player's move for every board (trainer.rs:506-534).  Needs 162 * n env slots in `ctx`."""
    boards = np.ascontiguousarray(boards, dtype=np.uint8).reshape(-1, CELLS)
    turns = np.ascontiguousarray(turns, dtype=np.uint8).reshape(-1)
    n = boards.shape[0]
    # candidate (game g, perspective s in {own, opponent}, cell a): board g with the turn of perspective s, move a
    cand_boards = np.repeat(boards, 2 * CELLS, axis=0)
    cand_turns = np.repeat(np.stack([turns, turns ^ 1], axis=1).reshape(-1), CELLS)
    cand_actions = np.tile(np.arange(CELLS, dtype=np.uint8), 2 * n)
    ids = np.arange(cand_boards.shape[0], dtype=np.int32)
    ctx.env_set(cand_boards, cand_turns, ids=ids)
    status, _ = ctx.env_step(cand_actions, ids=ids, want_legal=False)
    status = status.reshape(n, 2, CELLS)
    terminal = (status != B.NONE) & (status != 0)          # occupied cells answer NONE; Draw counts as terminal (:517,:525)
    ends = terminal[:, 0, :] | terminal[:, 1, :]            # per action: own win checked first, then the block -- same action
    out = np.empty(n, dtype=np.int32)
    for g in range(n):
        hits = np.flatnonzero(ends[g])
        if hits.size:
            out[g] = hits[0]
        else:
            legal = np.flatnonzero(boards[g] == 0)
            out[g] = legal[rng.integers(0, legal.size)]     # :536-537 uniform over the legal moves
    return out


def play_against_naive_player(ctx, episode_count: int = 100, count: int = 800, batch_size: int = 16, epsilon: float = 0.25,
                              alpha: float = 0.03, evaluator: int = B.EVAL_NET, seed: int = 0, moves_out: list | None = None):
    """Returns (black_win, white_win, draw): the naive player moves first (Black), the model answers (White).

    The naive player's random choices come from `numpy.random.default_rng(seed)`, one draw per live game without a
    forced move, games in ascending id order (the reference draws from thread_rng() in its swap_remove order,
    trainer.rs:498-563: unreproducible; the order of independent games does not change any single game).
    `moves_out`, if given, receives one list of actions per game (both players' moves in playing order)."""
    rng = np.random.default_rng(seed)
    ids = np.arange(episode_count, dtype=np.int32)
    ctx.pool_new_games(ids=ids, evaluator=evaluator)
    result = np.zeros(4, dtype=np.int64)
    log = [[] for _ in range(episode_count)]
    live = ids
    while live.size:
        boards, turns, _, _ = ctx.pool_get_envs(ids=live)  # one launch + one synchronisation for all live games
        actions = naive_moves(ctx, boards, turns, rng)
        for g, a in zip(live, actions):
            log[int(g)].append(int(a))
        status = _play_moves(ctx, live, actions, evaluator)
        np.add.at(result, status[status != 0].astype(np.int64), 1)
        live = live[status == 0]
        if not live.size:
            break
        ctx.pool_search(ids=live, count=count, batch_size=batch_size, epsilon=epsilon, alpha=alpha, evaluator=evaluator)
        actions, _ = ctx.pool_sample(ids=live, modes=np.full(live.size, B.SAMPLE_BEST, np.uint8))
        for g, a in zip(live, actions):
            log[int(g)].append(int(a))
        status = ctx.pool_play(actions, ids=live)
        np.add.at(result, status[status != 0].astype(np.int64), 1)
        live = live[status == 0]
    if moves_out is not None:
        moves_out[:] = log
    return int(result[2]), int(result[3]), int(result[1])

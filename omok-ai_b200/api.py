"""Host-side mirror of the reference's public API for the self-play path.

Same names, argument meaning and None/Err behaviour as the Rust crates, so the parity
tests read like the reference's own tests:

  environment::{Environment, Stone, Turn, GameStatus}        environment/src/lib.rs:4-193
  alpha_zero::{Agent, ActionSamplingMode, AgentModel,
               EnvTurnMode, encode_nn_input,
               MCTSExecutor, ParallelMCTSExecutor}           alpha-zero/src/*.rs

Every object is a handle onto a slot of a `binding.Context`; every method is one call
into libomok_b200.so.  No game logic, tree logic or network arithmetic lives here.
"""
from __future__ import annotations

import enum

import numpy as np

from . import binding as B


class Stone(enum.IntEnum):  # environment/src/lib.rs:4-9
    Empty = 0
    Black = 1
    White = 2


class Turn(enum.IntEnum):  # :21-25
    Black = 0
    White = 1

    def opponent(self) -> "Turn":
        return Turn(1 - int(self))


class GameStatus(enum.IntEnum):  # :45-51
    InProgress = 0
    Draw = 1
    BlackWin = 2
    WhiteWin = 3

    def is_terminal(self) -> bool:
        return self != GameStatus.InProgress


class EnvTurnMode(enum.IntEnum):  # alpha-zero/src/encoder.rs:4-8
    Player = 0
    Opponent = 1


class ActionSamplingMode:  # alpha-zero/src/agent.rs:236-241
    Best = ("best", 1.0)

    @staticmethod
    def Boltzmann(temperature: float):
        return ("boltzmann", float(temperature))


class _SlotAllocator:
    def __init__(self, capacity: int):
        self.free = list(range(capacity - 1, -1, -1))

    def take(self) -> int:
        if not self.free:
            raise B.OmkError(-3, "pool exhausted: raise the context capacity")
        return self.free.pop()

    def give(self, i: int):
        self.free.append(i)


def _allocator(ctx: B.Context, kind: str) -> _SlotAllocator:
    name = f"_alloc_{kind}"
    if not hasattr(ctx, name):
        setattr(ctx, name, _SlotAllocator(ctx.capacity_envs if kind == "env" else ctx.capacity_trees))
    return getattr(ctx, name)


class Environment:
    """environment::Environment (environment/src/lib.rs:62-193) on one slot of the env pool."""

    BOARD_SIZE = 9
    SERIAL_STONE_COUNT = 5

    def __init__(self, ctx: B.Context, _slot: int | None = None):
        self.ctx = ctx
        self.slot = _allocator(ctx, "env").take() if _slot is None else _slot
        if _slot is None:
            ctx.env_reset(ids=[self.slot])

    def __del__(self):
        try:
            _allocator(self.ctx, "env").give(self.slot)
        except Exception:
            pass

    def _get(self):
        return self.ctx.env_get(ids=[self.slot])

    @property
    def turn(self) -> Turn:
        return Turn(int(self._get()[1][0]))

    @property
    def legal_move_count(self) -> int:
        return int(self._get()[2][0])

    @property
    def board(self):
        return [Stone(int(s)) for s in self._get()[0][0]]

    def place_stone(self, index: int):
        status, _ = self.ctx.env_step([index], ids=[self.slot], want_legal=False)
        return None if status[0] == B.NONE else GameStatus(int(status[0]))

    def encode_board(self, turn: Turn) -> np.ndarray:
        """The 162-float two-plane interleave of encode_board (:81-102), perspective `turn`."""
        env_turn = int(self._get()[1][0])
        mode = 0 if int(turn) == env_turn else 1
        return self.ctx.env_encode(ids=[self.slot], mode=mode)[0, :162].copy()

    def clone(self) -> "Environment":
        boards, turns, _ = self._get()
        other = Environment(self.ctx)
        self.ctx.env_set(boards, turns, ids=[other.slot])
        return other


def encode_nn_input(ctx: B.Context, env_turn_mode: EnvTurnMode, envs) -> np.ndarray:
    """alpha_zero::encode_nn_input (encoder.rs:10-46): [n, 9, 9, 3] f32 memory image."""
    ids = [e.slot for e in envs]
    return ctx.env_encode(ids=ids, mode=int(env_turn_mode)).reshape(len(ids), 9, 9, 3)


class AgentModel:
    """alpha_zero::AgentModel forward side (agent_model.rs:105-134).  `session` is implicit (the Context)."""

    def __init__(self, ctx: B.Context, params=None, seed: int | None = None):
        self.ctx = ctx
        if params is not None:
            ctx.net_load_params(params)
        else:
            ctx.net_init_random(0 if seed is None else seed)

    @property
    def variables(self):
        return self.ctx.net_get_params()

    def evaluate_p(self, input_tensor: np.ndarray) -> np.ndarray:
        p, _ = self.ctx.net_eval_images(np.asarray(input_tensor, np.float32).reshape(-1, 243), want_v=False)
        return p.reshape(-1, 9, 9)

    def evaluate_pv(self, input_tensor: np.ndarray):
        p, v = self.ctx.net_eval_images(np.asarray(input_tensor, np.float32).reshape(-1, 243))
        return p.reshape(-1, 9, 9), v.reshape(-1, 1)


class Agent:
    """alpha_zero::Agent (agent.rs:10-232): an environment plus its search tree, on one tree slot."""

    def __init__(self, ctx: B.Context, evaluator: int = B.EVAL_NET, stream: int | None = None):
        self.ctx = ctx
        self.evaluator = evaluator
        self.slot = _allocator(ctx, "tree").take()
        ctx.pool_new_games(ids=[self.slot], streams=None if stream is None else [stream], evaluator=evaluator)

    def __del__(self):
        try:
            _allocator(self.ctx, "tree").give(self.slot)
        except Exception:
            pass

    @property
    def env(self):
        board, turn, legal = self.ctx.pool_get_env(self.slot)
        return {"board": board, "turn": Turn(turn), "legal_move_count": legal}

    def compute_policy(self):
        pol, valid = self.ctx.pool_policy(ids=[self.slot])
        return pol[0] if valid[0] else None

    def sample_action(self, mode=ActionSamplingMode.Best):
        kind, temperature = mode
        actions, pol = self.ctx.pool_sample(
            ids=[self.slot], modes=[B.SAMPLE_BOLTZMANN if kind == "boltzmann" else B.SAMPLE_BEST], temperatures=[temperature]
        )
        return None if actions[0] == B.NONE else (int(actions[0]), pol[0])

    def ensure_action_exists(self, action: int):
        self.ctx.pool_ensure_action([action], ids=[self.slot], evaluator=self.evaluator)

    def play_action(self, action: int):
        st = self.ctx.pool_play([action], ids=[self.slot])[0]
        return None if st == B.NONE else GameStatus(int(st))


class ParallelMCTSExecutor:
    """alpha_zero::ParallelMCTSExecutor (parallel_mcts_executor.rs:13-270)."""

    C_PUCT = 1.0

    def execute(self, count: int, batch_size: int, epsilon: float, alpha: float, agents):
        if not agents:
            return
        ctx = agents[0].ctx
        ctx.pool_search(ids=[a.slot for a in agents], count=count, batch_size=batch_size, epsilon=epsilon, alpha=alpha,
                        evaluator=agents[0].evaluator)


class MCTSExecutor:
    """alpha_zero::MCTSExecutor (mcts_executor.rs:16-255).  The reference runs its ceil(count/batch)
    rounds concurrently on one tree (racy, nondeterministic); the deterministic restatement is the same
    rounds executed in order, i.e. `execute` with a single agent."""

    C_PUCT = 1.0

    def run(self, count: int, batch_size: int, epsilon: float, alpha: float, agent: Agent):
        agent.ctx.pool_search(ids=[agent.slot], count=count, batch_size=batch_size, epsilon=epsilon, alpha=alpha,
                              evaluator=agent.evaluator)

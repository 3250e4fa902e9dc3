"""ORACLE (test infrastructure, NOT product code) -- ctypes view of libomok_oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs import this module.  It wraps oracle/omok_oracle.c, the CPU
restatement of the reference's environment / mcts / alpha-zero search
(citations inside the C file), and adds evaluator plumbing:
  * HashEvaluator  -> orc_hash_eval (exact fake net, bit-identical on the GPU)
  * TorchEvaluator -> oracle/net_oracle.py forward on CPU fp32
  * CallbackEvaluator(fn) -> any python callable (e.g. the GPU net, "recorded" mode)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libomok_oracle.so")

CELLS = 81
NONE = -1
IN_PROGRESS, DRAW, BLACK_WIN, WHITE_WIN = 0, 1, 2, 3


class Env(C.Structure):
    _fields_ = [("turn", C.c_uint8), ("legal_move_count", C.c_uint16), ("board", C.c_uint8 * CELLS)]


EVAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(Env), C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float))


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "omok_oracle.c")
    hdr = os.path.join(_HERE, "omok_oracle.h")
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or (os.path.exists(src) and os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libomok_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    P = C.POINTER
    L.orc_env_new.argtypes = [P(Env)]
    L.orc_env_place_stone.argtypes = [P(Env), C.c_int]
    L.orc_env_place_stone.restype = C.c_int
    L.orc_env_encode_board.argtypes = [P(Env), C.c_int, P(C.c_float)]
    L.orc_encode_nn_input.argtypes = [P(Env), C.c_int, C.c_int, P(C.c_float)]
    L.orc_random_playout.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, P(Env)]
    L.orc_rng_u32.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
    L.orc_rng_u32.restype = C.c_uint32
    L.orc_rng_below.argtypes = [C.c_uint64, C.c_uint32, P(C.c_uint32), C.c_uint32]
    L.orc_rng_below.restype = C.c_uint32
    L.orc_det_log.argtypes = [C.c_double]
    L.orc_det_log.restype = C.c_double
    L.orc_det_exp.argtypes = [C.c_double]
    L.orc_det_exp.restype = C.c_double
    L.orc_dirichlet81.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_float, P(C.c_float)]
    L.orc_hash_eval.argtypes = [P(Env), C.c_int, P(C.c_float), P(C.c_float)]
    L.orc_agent_new.argtypes = [EVAL_FN, C.c_void_p, C.c_uint64, C.c_uint32]
    L.orc_agent_new.restype = C.c_void_p
    L.orc_agent_free.argtypes = [C.c_void_p]
    L.orc_agent_env.argtypes = [C.c_void_p]
    L.orc_agent_env.restype = P(Env)
    L.orc_agent_rng_counter.argtypes = [C.c_void_p]
    L.orc_agent_rng_counter.restype = C.c_uint32
    L.orc_agent_node_count.argtypes = [C.c_void_p]
    L.orc_agent_node_count.restype = C.c_int
    L.orc_execute.argtypes = [
        P(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, EVAL_FN, C.c_void_p, C.c_int,
    ]
    L.orc_execute.restype = C.c_int64
    L.orc_agent_root_children.argtypes = [C.c_void_p, P(C.c_int32), P(C.c_uint64), P(C.c_float), P(C.c_float)]
    L.orc_agent_root_children.restype = C.c_int
    L.orc_agent_root_stats.argtypes = [
        C.c_void_p, P(C.c_uint64), P(C.c_float), P(C.c_float), P(C.c_int), P(C.c_float),
    ]
    L.orc_agent_compute_policy.argtypes = [C.c_void_p, P(C.c_float)]
    L.orc_agent_compute_policy.restype = C.c_int
    L.orc_agent_sample_action.argtypes = [C.c_void_p, C.c_int, C.c_float, P(C.c_float)]
    L.orc_agent_sample_action.restype = C.c_int
    L.orc_agent_ensure_action_exists.argtypes = [C.c_void_p, C.c_int, EVAL_FN, C.c_void_p]
    L.orc_agent_play_action.argtypes = [C.c_void_p, C.c_int]
    L.orc_agent_play_action.restype = C.c_int
    _lib = L
    return L


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


# --------------------------------------------------------------------------
# Environment (environment crate)
# --------------------------------------------------------------------------
class Environment:
    def __init__(self):
        self.e = Env()
        lib().orc_env_new(C.byref(self.e))

    @property
    def turn(self) -> int:
        return int(self.e.turn)

    @property
    def legal_move_count(self) -> int:
        return int(self.e.legal_move_count)

    @property
    def board(self) -> np.ndarray:
        return np.frombuffer(bytes(self.e.board), dtype=np.uint8).copy()

    def place_stone(self, index: int):
        r = lib().orc_env_place_stone(C.byref(self.e), int(index))
        return None if r == NONE else r

    def encode_board(self, turn: int) -> np.ndarray:
        out = np.zeros(162, dtype=np.float32)
        lib().orc_env_encode_board(C.byref(self.e), int(turn), _fp(out))
        return out

    def encode_nn_input(self, mode: int = 0) -> np.ndarray:
        out = np.zeros(243, dtype=np.float32)
        lib().orc_encode_nn_input(C.byref(self.e), 1, int(mode), _fp(out))
        return out

    def clone(self) -> "Environment":
        o = Environment()
        C.memmove(C.byref(o.e), C.byref(self.e), C.sizeof(Env))
        return o


def random_playout(n: int, plies: int, seed: int):
    """(actions [plies,n] u8, status [plies,n] i8, final boards [n,81], final turns [n])."""
    actions = np.zeros((plies, n), dtype=np.uint8)
    status = np.zeros((plies, n), dtype=np.int8)
    envs = (Env * n)()
    lib().orc_random_playout(n, plies, seed, actions.ctypes.data, status.ctypes.data, envs)
    boards, turns = envs_to_arrays(envs, n)
    return actions, status, boards, turns


def envs_to_arrays(envs_ptr, n: int):
    """(boards [n,81] u8, turns [n] u8) from an orc_env array pointer."""
    boards = np.empty((n, CELLS), dtype=np.uint8)
    turns = np.empty(n, dtype=np.uint8)
    for i in range(n):
        boards[i] = np.frombuffer(bytes(envs_ptr[i].board), dtype=np.uint8)
        turns[i] = envs_ptr[i].turn
    return boards, turns


# --------------------------------------------------------------------------
# Evaluators
# --------------------------------------------------------------------------
class Evaluator:
    """Holds the ctypes callback alive; subclasses implement eval_arrays."""

    def __init__(self):
        self.calls = 0
        self.positions = 0
        self._cb = EVAL_FN(self._trampoline)

    def _trampoline(self, user, envs, n, mode, out_p, out_v):
        self.calls += 1
        self.positions += n
        boards, turns = envs_to_arrays(envs, n)
        p, v = self.eval_arrays(boards, turns, int(mode))
        p = np.ascontiguousarray(p, dtype=np.float32).reshape(n, CELLS)
        v = np.ascontiguousarray(v, dtype=np.float32).reshape(n)
        C.memmove(out_p, p.ctypes.data, p.nbytes)
        if out_v:
            C.memmove(out_v, v.ctypes.data, v.nbytes)

    def eval_arrays(self, boards, turns, mode):  # pragma: no cover
        raise NotImplementedError

    @property
    def fn(self):
        return self._cb


class HashEvaluator(Evaluator):
    """orc_hash_eval through the generic callback path (keeps one code path)."""

    def eval_arrays(self, boards, turns, mode):
        n = boards.shape[0]
        p = np.zeros((n, CELLS), dtype=np.float32)
        v = np.zeros(n, dtype=np.float32)
        e = Env()
        for i in range(n):
            e.turn = int(turns[i])
            C.memmove(e.board, boards[i].ctypes.data, CELLS)
            vv = C.c_float()
            lib().orc_hash_eval(C.byref(e), mode, _fp(p[i]), C.byref(vv))
            v[i] = vv.value
        return p, v


class NativeHashEvaluator:
    """orc_eval_hash_batch passed straight to C (no Python in the loop); same results as HashEvaluator."""

    def __init__(self):
        self._cb = C.cast(lib().orc_eval_hash_batch, EVAL_FN)

    @property
    def fn(self):
        return self._cb


class TorchEvaluator(Evaluator):
    """oracle/net_oracle.py on CPU fp32 (stands in for TensorFlow-CPU)."""

    def __init__(self, params):
        super().__init__()
        self.params = params

    def eval_arrays(self, boards, turns, mode):
        from . import net_oracle  # noqa: PLC0415

        p, v, _ = net_oracle.forward_boards(self.params, boards, turns, opponent_mode=bool(mode))
        return p, v


class CallbackEvaluator(Evaluator):
    def __init__(self, fn):
        super().__init__()
        self._fn = fn

    def eval_arrays(self, boards, turns, mode):
        return self._fn(boards, turns, mode)


# --------------------------------------------------------------------------
# Agent + executor (alpha-zero crate)
# --------------------------------------------------------------------------
class Agent:
    def __init__(self, evaluator: Evaluator, seed: int = 0, stream: int = 0):
        self.h = lib().orc_agent_new(evaluator.fn, None, seed, stream)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_agent_free(self.h)
            self.h = None

    @property
    def env(self) -> Env:
        return lib().orc_agent_env(self.h).contents

    def board(self) -> np.ndarray:
        return np.frombuffer(bytes(self.env.board), dtype=np.uint8).copy()

    @property
    def rng_counter(self) -> int:
        return lib().orc_agent_rng_counter(self.h)

    @property
    def node_count(self) -> int:
        return lib().orc_agent_node_count(self.h)

    def root_children(self):
        a = np.zeros(CELLS, dtype=np.int32)
        n = np.zeros(CELLS, dtype=np.uint64)
        w = np.zeros(CELLS, dtype=np.float32)
        p = np.zeros(CELLS, dtype=np.float32)
        k = lib().orc_agent_root_children(
            self.h, a.ctypes.data_as(C.POINTER(C.c_int32)), n.ctypes.data_as(C.POINTER(C.c_uint64)), _fp(w), _fp(p)
        )
        return a[:k], n[:k], w[:k], p[:k]

    def root_stats(self):
        n = C.c_uint64()
        w = C.c_float()
        p = C.c_float()
        st = C.c_int()
        pol = np.zeros(CELLS, dtype=np.float32)
        lib().orc_agent_root_stats(self.h, C.byref(n), C.byref(w), C.byref(p), C.byref(st), _fp(pol))
        return n.value, w.value, p.value, st.value, pol

    def compute_policy(self):
        pol = np.zeros(CELLS, dtype=np.float32)
        ok = lib().orc_agent_compute_policy(self.h, _fp(pol))
        return pol if ok else None

    def sample_action(self, mode: int = 0, temperature: float = 1.0):
        pol = np.zeros(CELLS, dtype=np.float32)
        a = lib().orc_agent_sample_action(self.h, mode, temperature, _fp(pol))
        return None if a == NONE else (a, pol)

    def ensure_action_exists(self, action: int, evaluator: Evaluator):
        lib().orc_agent_ensure_action_exists(self.h, action, evaluator.fn, None)

    def play_action(self, action: int):
        r = lib().orc_agent_play_action(self.h, action)
        return None if r == NONE else r


def execute(agents, count, batch_size, epsilon, alpha, evaluator: Evaluator, n_threads: int = 1) -> int:
    arr = (C.c_void_p * len(agents))(*[a.h for a in agents])
    return lib().orc_execute(arr, len(agents), count, batch_size, epsilon, alpha, evaluator.fn, None, n_threads)

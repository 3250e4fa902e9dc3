"""omok-ai_b200 -- B200-native (sm_100a) self-play hot path of AcrylicShrimp/omok-ai.

This package is a thin ctypes view of `libomok_b200.so` (C ABI: include/omok_b200.h).
All compute happens in hand-written CUDA kernels (csrc/); there is no CPU fallback:
importing works anywhere (so the ABI can be inspected), but creating a Context
without the built library or without a CUDA device raises.

The directory name contains a hyphen, so import it with
    importlib.import_module("omok-ai_b200")
"""
from __future__ import annotations

import os
import subprocess

from .binding import (  # noqa: F401
    CELLS,
    EVAL_HASH,
    EVAL_NET,
    NONE,
    SAMPLE_BEST,
    SAMPLE_BOLTZMANN,
    Context,
    OmkError,
    SelfPlayStats,
    declared_symbols,
    lib_path,
    load_library,
)
from .api import (  # noqa: F401
    ActionSamplingMode,
    Agent,
    AgentModel,
    Environment,
    EnvTurnMode,
    GameStatus,
    MCTSExecutor,
    ParallelMCTSExecutor,
    Stone,
    Turn,
    encode_nn_input,
)

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu into libomok_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
        print(out.stderr)
    if out.returncode != 0:
        raise RuntimeError("building libomok_b200.so failed")
    return lib_path()

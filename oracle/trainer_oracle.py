"""ORACLE (test infrastructure, NOT product code) -- trainer-side pieces of an AlphaZero iteration.

Plain-Python / numpy restatement of what the reference does between two self-play phases.  Only `tests/` may import
this file; the product path (omok-ai_b200/) never does.  Citations are relative to /root/reference/.

  * rotate_90 / rotate_180 / rotate_270 / flip_horizontal / flip_vertical   src/utils.rs:1-64
      PINNED on the reference's own golden vectors (src/utils.rs:70-108, the 2x2 cases), see tests/test_trainer_cpu.py
  * value-target back-fill and the 5 augmented copies per transition         src/trainer.rs:208-318
  * SavedData bincode image                                                  alpha-zero/src/model_io.rs:20-24,59-91
      third-party: bincode 1.3.3 (Cargo.lock), default options = little-endian, fixed-width ints, u64 length prefixes.
      PARITY UNPINNED: the reference holds no checkpoint fixture; anchored on the published format.
  * Adadelta update                                                          alpha-zero/src/agent_model.rs:75-83
      third-party: tensorflow 0.21.0 `train::AdadeltaOptimizer` (defaults rho 0.95, epsilon 1e-8) -> ApplyAdadelta:
          accum        = rho * accum + (1 - rho) * g^2
          update       = sqrt(accum_update + eps) / sqrt(accum + eps) * g
          var         -= lr * update
          accum_update = rho * accum_update + (1 - rho) * update^2
      PARITY UNPINNED (no reference test); anchored on the op's published semantics.
  * losses                                                                   agent_model.rs:60-73, network.rs:249-253
"""
from __future__ import annotations

import struct

import numpy as np


def rotate_90(src, size):
    dst = [None] * (size * size)
    for i in range(size):
        for j in range(size):
            dst[i * size + j] = src[(size - j - 1) * size + i]
    return dst


def rotate_180(src, size):
    dst = [None] * (size * size)
    for i in range(size):
        for j in range(size):
            dst[i * size + j] = src[(size - i - 1) * size + (size - j - 1)]
    return dst


def rotate_270(src, size):
    dst = [None] * (size * size)
    for i in range(size):
        for j in range(size):
            dst[i * size + j] = src[j * size + (size - i - 1)]
    return dst


def flip_horizontal(src, size):
    dst = [None] * (size * size)
    for i in range(size):
        for j in range(size):
            dst[i * size + j] = src[i * size + (size - j - 1)]
    return dst


def flip_vertical(src, size):
    dst = [None] * (size * size)
    for i in range(size):
        for j in range(size):
            dst[i * size + j] = src[(size - i - 1) * size + j]
    return dst


SYMMETRY_FUNCS = (rotate_90, rotate_180, rotate_270, flip_horizontal, flip_vertical)  # order of trainer.rs:223-315


def episode_to_replay(boards, policies, final_z):
    """trainer.rs:208-318: z back-fill from the last ply with alternating sign; originals first, then for every
    transition its five symmetric copies.  Returns a list of (board[81], policy[81], z)."""
    T = len(boards)
    zs = [0.0] * T
    z = float(final_z)
    for t in range(T - 1, -1, -1):
        zs[t] = z
        z = -z
    out = [(list(boards[t]), list(policies[t]), zs[t]) for t in range(T)]
    for t in range(T):
        for fn in SYMMETRY_FUNCS:
            out.append((fn(list(boards[t]), 9), fn(list(policies[t]), 9), zs[t]))
    return out


def bincode_saved_data(names, params) -> bytes:
    """bincode 1.x default image of SavedData { variable_names: Vec<String>, parameters: Vec<Vec<f32>> }."""
    out = bytearray()
    out += struct.pack("<Q", len(names))
    for n in names:
        raw = n.encode("utf-8")
        out += struct.pack("<Q", len(raw))
        out += raw
    out += struct.pack("<Q", len(params))
    for p in params:
        flat = np.asarray(p, dtype=np.float32).reshape(-1)
        out += struct.pack("<Q", flat.size)
        for v in flat.tolist():
            out += struct.pack("<f", v)
    return bytes(out)


class Adadelta:
    """tensorflow ApplyAdadelta, one state pair per variable, in float64 (a reference value for fp32 runs)."""

    def __init__(self, params, lr=0.01, rho=0.95, eps=1e-8):
        self.p = [np.asarray(p, dtype=np.float64).copy() for p in params]
        self.accum = [np.zeros_like(p) for p in self.p]
        self.accum_update = [np.zeros_like(p) for p in self.p]
        self.lr, self.rho, self.eps = lr, rho, eps

    def step(self, grads):
        for i, g in enumerate(grads):
            g = np.asarray(g, dtype=np.float64)
            self.accum[i] = self.rho * self.accum[i] + (1 - self.rho) * g * g
            update = np.sqrt(self.accum_update[i] + self.eps) / np.sqrt(self.accum[i] + self.eps) * g
            self.p[i] = self.p[i] - self.lr * update
            self.accum_update[i] = self.rho * self.accum_update[i] + (1 - self.rho) * update * update
        return self.p


def losses_fp64(params, images, pi, z):
    """(p_loss, v_loss, loss) in float64 through the network oracle (oracle/net_oracle.py)."""
    import torch

    from . import net_oracle

    _, v, logits = net_oracle.forward(params, images, dtype=torch.float64)
    # net_oracle.forward returns float32 views of its float64 results; recompute the log-softmax in float64
    logits = np.asarray(logits, dtype=np.float64)
    m = logits.max(axis=1, keepdims=True)
    logp = logits - m - np.log(np.exp(logits - m).sum(axis=1, keepdims=True))
    p_loss = float(np.mean(-(np.asarray(pi, np.float64).reshape(-1, 81) * logp).sum(axis=1)))
    v_loss = float(np.mean((np.asarray(z, np.float64).reshape(-1) - np.asarray(v, np.float64)) ** 2))
    return p_loss, v_loss, p_loss + v_loss

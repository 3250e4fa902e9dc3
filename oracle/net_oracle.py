"""ORACLE (test infrastructure, NOT product code) -- policy/value network forward.

CPU fp32 restatement of the reference network for the self-play path.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl
reference legs may import this file.  The product path (omok-ai_b200/) never does.

What it restates (all citations relative to /root/reference/):
  * graph:      alpha-zero/src/network.rs:51-262  (stem 1x1 3->128, 3x bottleneck
                residual, flatten NHWC, fc0 10368->512, fc1 512->512, value fc
                512->1 + tanh, policy fc 512->81 + softmax)
  * layers:     network-utils/src/lib.rs:95-170 (conv2d + bias_add),
                :172-262 (depthwise 3x3 + pointwise 1x1 + bias_add),
                :285-330 (fc), :386-461 (bottleneck residual: 1x1 -> lrelu ->
                separable -> lrelu -> 1x1 -> +x; the caller applies the last lrelu,
                network.rs:108-111)
  * init:       network-utils/src/lib.rs:86-92 (He 2/sqrt(fan_in), Xavier
                2/sqrt(fan_in+fan_out)), biases zero
  * activation: tensorflow `LeakyRelu` op with its default alpha = 0.2 (the
                reference never sets alpha; network.rs:77,108,148,161)
  * input:      alpha-zero/src/encoder.rs:10-46 + environment/src/lib.rs:81-102.
                NB the reference writes a 2-channel interleave into floats [0,162)
                and the turn plane into floats [162,243) of each 243-float slot and
                TensorFlow then *reads* that slot as [9,9,3].  `encode_image`
                reproduces the memory image; `forward` reads it as NHWC [B,9,9,3].

Third-party arithmetic: the reference delegates to tensorflow 0.21.0 /
tensorflow-sys 0.24.0 (Cargo.lock:3530-3568), absent from /root/reference and not
installable here.  PARITY UNPINNED for the network: the reference holds no test
or golden vector for network outputs; this restatement is anchored on the call
sites above and on TensorFlow's published op semantics (Conv2D / DepthwiseConv2dNative
NHWC SAME stride 1, BiasAdd, LeakyRelu(0.2), MatMul, Tanh, Softmax).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

BOARD = 9
CELLS = 81
CH = 128
MID = 32
NRES = 3
FC = 512
LRELU = 0.2

# Variable order == checkpoint order (network.rs:78-79,113-122,149-150,162-163,201-202,240-241)
PARAM_SPECS = (
    [("conv_w", (1, 1, 3, CH)), ("conv_b", (CH,))]
    + sum(
        [
            [
                (f"res{i}_w0", (1, 1, CH, MID)),
                (f"res{i}_b0", (MID,)),
                (f"res{i}_dw", (3, 3, MID, 1)),
                (f"res{i}_pw", (1, 1, MID, MID)),
                (f"res{i}_b1", (MID,)),
                (f"res{i}_w2", (1, 1, MID, CH)),
                (f"res{i}_b2", (CH,)),
            ]
            for i in range(NRES)
        ],
        [],
    )
    + [
        ("fc0_w", (CELLS * CH, FC)),
        ("fc0_b", (FC,)),
        ("fc1_w", (FC, FC)),
        ("fc1_b", (FC,)),
        ("v_w", (FC, 1)),
        ("v_b", (1,)),
        ("p_w", (FC, CELLS)),
        ("p_b", (CELLS,)),
    ]
)
assert len(PARAM_SPECS) == 31
N_PARAMS = sum(int(np.prod(s)) for _, s in PARAM_SPECS)
assert N_PARAMS == 5_643_250

FLOP_PER_POSITION = 2 * (
    CELLS * 3 * CH
    + NRES * CELLS * (CH * MID + 9 * MID + MID * MID + MID * CH)
    + CELLS * CH * FC
    + FC * FC
    + FC * 1
    + FC * CELLS
)


def _he(fan_in: int) -> float:
    return float(np.float32(2.0) / np.sqrt(np.float32(fan_in)))


def _xavier(fan_in: int, fan_out: int) -> float:
    return float(np.float32(2.0) / np.sqrt(np.float32(fan_in + fan_out)))


def init_scales() -> dict[str, float]:
    """Scale constants of the N3 recipe (network-utils/src/lib.rs:86-92,116-137,189-230,298-304)."""
    s = {"conv_w": _he(1 * 1 * 3)}
    for i in range(NRES):
        s[f"res{i}_w0"] = _he(CH)
        s[f"res{i}_dw"] = _he(3 * 3 * MID)  # depthwise fan_in = kh*kw*cin (lib.rs:199-202)
        s[f"res{i}_pw"] = _he(MID)
        s[f"res{i}_w2"] = _he(MID)
    s["fc0_w"] = _he(CELLS * CH)
    s["fc1_w"] = _he(FC)
    s["v_w"] = _xavier(FC, 1)
    s["p_w"] = _xavier(FC, CELLS)
    return s


def random_params(seed: int = 0) -> list[np.ndarray]:
    """Random-init weights, `w = N(0,1) * c`, biases 0, in checkpoint order.

    The reference draws N(0,1) from TensorFlow's RandomStandardNormal with an
    unseeded graph (unreproducible); our stream is numpy's PCG64 keyed by
    `seed`, drawn tensor by tensor in checkpoint order.  The same arrays are
    handed to the CUDA path by the tests (omk_net_load_params), so the stream
    choice never affects parity.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    scales = init_scales()
    out = []
    for name, shape in PARAM_SPECS:
        if name in scales:
            w = rng.standard_normal(size=shape, dtype=np.float32) * np.float32(scales[name])
            out.append(np.ascontiguousarray(w, dtype=np.float32))
        else:
            out.append(np.zeros(shape, dtype=np.float32))
    return out


def encode_image(board: np.ndarray, turn: int, opponent_mode: bool = False) -> np.ndarray:
    """243-float memory image of ONE position (encoder.rs:22-43, environment lib.rs:81-102).

    board: 81 x {0 empty, 1 black, 2 white}; turn: 0 black to move, 1 white to move.
    Perspective = env.turn (EnvTurnMode::Player) or its opponent (::Opponent).
    """
    img = np.zeros(243, dtype=np.float32)
    persp = turn ^ 1 if opponent_mode else turn
    black_off = 0 if persp == 0 else 1
    white_off = 1 - black_off
    for i in range(CELLS):
        if board[i] == 1:
            img[2 * i + black_off] = 1.0
        elif board[i] == 2:
            img[2 * i + white_off] = 1.0
    img[162:] = 1.0 if turn == 0 else 0.0
    return img


def forward(params: list[np.ndarray], images: np.ndarray, dtype=torch.float32):
    """images: [B,243] float32 memory images -> (P [B,81], V [B], logits [B,81])."""
    p = {name: torch.from_numpy(np.asarray(a)).to(dtype) for (name, _), a in zip(PARAM_SPECS, params)}
    x = torch.from_numpy(np.asarray(images, dtype=np.float32)).to(dtype)
    B = x.shape[0]
    x = x.reshape(B, BOARD, BOARD, 3).permute(0, 3, 1, 2)  # TF reads the slot as NHWC

    def conv1x1(t, w, b):
        # TF filter [1,1,cin,cout] -> torch [cout,cin,1,1]
        return F.conv2d(t, w[0, 0].t().reshape(w.shape[3], w.shape[2], 1, 1), b)

    x = F.leaky_relu(conv1x1(x, p["conv_w"], p["conv_b"]), LRELU)
    for i in range(NRES):
        h = F.leaky_relu(conv1x1(x, p[f"res{i}_w0"], p[f"res{i}_b0"]), LRELU)
        dw = p[f"res{i}_dw"]  # [3,3,32,1] -> torch depthwise [32,1,3,3]
        h = F.conv2d(h, dw[:, :, :, 0].permute(2, 0, 1).unsqueeze(1), None, padding=1, groups=MID)
        h = F.leaky_relu(conv1x1(h, p[f"res{i}_pw"], p[f"res{i}_b1"]), LRELU)
        h = conv1x1(h, p[f"res{i}_w2"], p[f"res{i}_b2"])
        x = F.leaky_relu(h + x, LRELU)
    flat = x.permute(0, 2, 3, 1).reshape(B, CELLS * CH)  # NHWC flatten: (y*9+x)*128+c
    h = F.leaky_relu(flat @ p["fc0_w"] + p["fc0_b"], LRELU)
    h = F.leaky_relu(h @ p["fc1_w"] + p["fc1_b"], LRELU)
    v = torch.tanh(h @ p["v_w"] + p["v_b"]).reshape(B)
    logits = h @ p["p_w"] + p["p_b"]
    pol = torch.softmax(logits, dim=1)
    return pol.float().numpy(), v.float().numpy(), logits.float().numpy()


def loss_and_grads(params: list[np.ndarray], images: np.ndarray, pi: np.ndarray, z: np.ndarray, dtype=torch.float64):
    """The reference's training loss (alpha-zero/src/agent_model.rs:60-73, network.rs:249-253) and its gradient with
    respect to the 31 tensors, by autograd over the same graph as `forward` in `dtype`:
    loss = mean((z - v)^2) + mean_i(-sum_j pi_ij log_softmax(logits)_ij).  Returns ((p_loss, v_loss, loss), grads)."""
    p = {name: torch.from_numpy(np.asarray(a)).to(dtype).requires_grad_(True) for (name, _), a in zip(PARAM_SPECS, params)}
    x = torch.from_numpy(np.asarray(images, dtype=np.float32)).to(dtype)
    B = x.shape[0]
    x = x.reshape(B, BOARD, BOARD, 3).permute(0, 3, 1, 2)

    def conv1x1(t, w, b):
        return F.conv2d(t, w[0, 0].t().reshape(w.shape[3], w.shape[2], 1, 1), b)

    x = F.leaky_relu(conv1x1(x, p["conv_w"], p["conv_b"]), LRELU)
    for i in range(NRES):
        h = F.leaky_relu(conv1x1(x, p[f"res{i}_w0"], p[f"res{i}_b0"]), LRELU)
        dw = p[f"res{i}_dw"]
        h = F.conv2d(h, dw[:, :, :, 0].permute(2, 0, 1).unsqueeze(1), None, padding=1, groups=MID)
        h = F.leaky_relu(conv1x1(h, p[f"res{i}_pw"], p[f"res{i}_b1"]), LRELU)
        h = conv1x1(h, p[f"res{i}_w2"], p[f"res{i}_b2"])
        x = F.leaky_relu(h + x, LRELU)
    flat = x.permute(0, 2, 3, 1).reshape(B, CELLS * CH)
    h = F.leaky_relu(flat @ p["fc0_w"] + p["fc0_b"], LRELU)
    h = F.leaky_relu(h @ p["fc1_w"] + p["fc1_b"], LRELU)
    v = torch.tanh(h @ p["v_w"] + p["v_b"])
    logits = h @ p["p_w"] + p["p_b"]
    zt = torch.from_numpy(np.asarray(z, dtype=np.float32)).to(dtype).reshape(B, 1)
    pt = torch.from_numpy(np.asarray(pi, dtype=np.float32)).to(dtype).reshape(B, CELLS)
    v_loss = torch.mean((zt - v) ** 2)
    p_loss = torch.mean(-(pt * F.log_softmax(logits, dim=1)).sum(dim=1))
    loss = v_loss + p_loss
    loss.backward()
    grads = [p[name].grad.detach().numpy() for name, _ in PARAM_SPECS]
    return (p_loss.item(), v_loss.item(), loss.item()), grads


def forward_layers(params: list[np.ndarray], images: np.ndarray, dtype=torch.float64) -> dict[str, np.ndarray]:
    """The same graph as `forward`, returning every layer the CUDA path can be inspected at, in `dtype` precision (no
    cast to fp32): tower [B,81,128] (the NHWC flatten fc0 reads), fc0 [B,512], fc1 [B,512], logits [B,81], vlogit [B],
    P [B,81], V [B].  Used by the per-layer error study (tools/net_error_study.py, tests/test_net_gpu.py)."""
    p = {name: torch.from_numpy(np.asarray(a)).to(dtype) for (name, _), a in zip(PARAM_SPECS, params)}
    x = torch.from_numpy(np.asarray(images, dtype=np.float32)).to(dtype)
    B = x.shape[0]
    x = x.reshape(B, BOARD, BOARD, 3).permute(0, 3, 1, 2)

    def conv1x1(t, w, b):
        return F.conv2d(t, w[0, 0].t().reshape(w.shape[3], w.shape[2], 1, 1), b)

    x = F.leaky_relu(conv1x1(x, p["conv_w"], p["conv_b"]), LRELU)
    for i in range(NRES):
        h = F.leaky_relu(conv1x1(x, p[f"res{i}_w0"], p[f"res{i}_b0"]), LRELU)
        dw = p[f"res{i}_dw"]
        h = F.conv2d(h, dw[:, :, :, 0].permute(2, 0, 1).unsqueeze(1), None, padding=1, groups=MID)
        h = F.leaky_relu(conv1x1(h, p[f"res{i}_pw"], p[f"res{i}_b1"]), LRELU)
        h = conv1x1(h, p[f"res{i}_w2"], p[f"res{i}_b2"])
        x = F.leaky_relu(h + x, LRELU)
    tower = x.permute(0, 2, 3, 1).reshape(B, CELLS, CH)
    fc0 = F.leaky_relu(tower.reshape(B, CELLS * CH) @ p["fc0_w"] + p["fc0_b"], LRELU)
    fc1 = F.leaky_relu(fc0 @ p["fc1_w"] + p["fc1_b"], LRELU)
    vlogit = (fc1 @ p["v_w"] + p["v_b"]).reshape(B)
    logits = fc1 @ p["p_w"] + p["p_b"]
    out = {"tower": tower, "fc0": fc0, "fc1": fc1, "logits": logits, "vlogit": vlogit,
           "P": torch.softmax(logits, dim=1), "V": torch.tanh(vlogit)}
    return {k: v.numpy() for k, v in out.items()}


def forward_boards(params, boards: np.ndarray, turns: np.ndarray, opponent_mode: bool = False, dtype=torch.float32):
    imgs = np.stack([encode_image(b, int(t), opponent_mode) for b, t in zip(boards, turns)])
    return forward(params, imgs, dtype)

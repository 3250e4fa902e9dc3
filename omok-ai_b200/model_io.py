"""Checkpoint wire format of the reference: alpha_zero::ModelIO::{save, load} (alpha-zero/src/model_io.rs:20-24,59-120).

`SavedData { variable_names: Vec<String>, parameters: Vec<Vec<f32>> }` serialised with bincode 1.3.3's default
configuration (Cargo.lock: bincode 1.3.3): little-endian, fixed-width integers, every sequence prefixed by its length
as u64, strings as u64 length + UTF-8 bytes, f32 as 4 raw bytes.  Optimizer slots are not stored (only
`network.variables`, model_io.rs:59 via agent_model.rs:82).  On load the reference ignores the names and zips the
parameter vectors with its variables in graph order (model_io.rs:98), so this order is the contract:

    conv_w conv_b  {residual_i_conv0_w _b  residual_i_conv1_w (depthwise) _w_1 (pointwise) _b  residual_i_conv2_w _b} x 3
    fc0_w fc0_b  fc1_w fc1_b  v_fc0_w v_fc0_b  p_fc0_w p_fc0_b

which is exactly the tensor order of `omk_net_load_params` / `omk_net_get_params` (include/omok_b200.h).
"""
from __future__ import annotations

import struct

import numpy as np

PARAM_SHAPES = (
    [(1, 1, 3, 128), (128,)]
    + [s for _ in range(3) for s in ((1, 1, 128, 32), (32,), (3, 3, 32, 1), (1, 1, 32, 32), (32,), (1, 1, 32, 128), (128,))]
    + [(10368, 512), (512,), (512, 512), (512,), (512, 1), (1,), (512, 81), (81,)]
)
# TensorFlow op names the reference's builders produce (network-utils/src/lib.rs:138,147,203,231,250,305,314;
# network.rs:66,97,140,153,189,228); the depthwise and pointwise filters of a separable conv share "{name}_w", so the
# graph uniquifies the second one to "{name}_w_1".
VARIABLE_NAMES = (
    ["conv_w", "conv_b"]
    + [n for i in range(3) for n in (f"residual_{i}_conv0_w", f"residual_{i}_conv0_b", f"residual_{i}_conv1_w",
                                     f"residual_{i}_conv1_w_1", f"residual_{i}_conv1_b", f"residual_{i}_conv2_w",
                                     f"residual_{i}_conv2_b")]
    + ["fc0_w", "fc0_b", "fc1_w", "fc1_b", "v_fc0_w", "v_fc0_b", "p_fc0_w", "p_fc0_b"]
)
assert len(PARAM_SHAPES) == len(VARIABLE_NAMES) == 31


class ModelIOError(Exception):
    pass


def dumps(params, names=VARIABLE_NAMES) -> bytes:
    """Serialise 31 parameter tensors (any shape, C order) to the reference's bincode image."""
    if len(params) != len(names):
        raise ModelIOError("variable_names and parameters differ in length")
    out = [struct.pack("<Q", len(names))]
    for n in names:
        b = n.encode("utf-8")
        out.append(struct.pack("<Q", len(b)))
        out.append(b)
    out.append(struct.pack("<Q", len(params)))
    for p in params:
        a = np.ascontiguousarray(p, dtype="<f4").reshape(-1)
        out.append(struct.pack("<Q", a.size))
        out.append(a.tobytes())
    return b"".join(out)


def loads(data: bytes, shapes=PARAM_SHAPES):
    """Parse a bincode image -> (variable_names, parameters reshaped to `shapes`).  Like the reference, parameters are
    matched to variables by position; a length that does not fit its variable is an error (the reference panics in
    `copy_from_slice`, model_io.rs:106)."""
    mv = memoryview(data)
    pos = 0

    def u64():
        nonlocal pos
        if pos + 8 > len(mv):
            raise ModelIOError("unexpected end of file")
        (v,) = struct.unpack_from("<Q", mv, pos)
        pos += 8
        return v

    names = []
    for _ in range(u64()):
        n = u64()
        if pos + n > len(mv):
            raise ModelIOError("unexpected end of file")
        names.append(bytes(mv[pos:pos + n]).decode("utf-8"))
        pos += n
    params = []
    count = u64()
    for i in range(count):
        n = u64()
        if pos + 4 * n > len(mv):
            raise ModelIOError("unexpected end of file")
        a = np.frombuffer(mv, dtype="<f4", count=n, offset=pos).astype(np.float32)
        pos += 4 * n
        if shapes is not None and i < len(shapes):
            if a.size != int(np.prod(shapes[i])):
                raise ModelIOError(f"parameter {i}: {a.size} values do not fit variable shape {shapes[i]}")
            a = a.reshape(shapes[i])
        params.append(a)
    if shapes is not None and count < len(shapes):
        raise ModelIOError(f"checkpoint holds {count} parameters, the network has {len(shapes)}")
    return names, params[: len(shapes)] if shapes is not None else params


def save(path, params, names=VARIABLE_NAMES) -> None:
    with open(path, "wb") as f:
        f.write(dumps(params, names))


def load(path, shapes=PARAM_SHAPES):
    with open(path, "rb") as f:
        return loads(f.read(), shapes)

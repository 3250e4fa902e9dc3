"""Second, independently written restatement of the reference's policy/value network forward -- TEST INFRASTRUCTURE, used
only by tests/test_oracle_cross.py to cross-check oracle/net_oracle.py (which builds the graph on torch.conv2d).

Written straight from the Rust graph builder and TensorFlow's published op semantics, in plain numpy with explicit tap and
channel loops (no torch, no shared code with oracle/):
  graph        /root/reference/alpha-zero/src/network.rs:51-262
  conv2d       /root/reference/network-utils/src/lib.rs:95-170   Conv2D NHWC, filter [fh, fw, cin, cout], SAME, + BiasAdd
  separable    /root/reference/network-utils/src/lib.rs:172-262  DepthwiseConv2dNative NHWC filter [fh, fw, cin, 1] SAME, then
                                                                 a 1x1 Conv2D [1,1,cin,cout], then ONE BiasAdd (no activation between)
  bottleneck   /root/reference/network-utils/src/lib.rs:386-461  conv0 1x1 -> LeakyRelu -> separable -> LeakyRelu -> conv2 1x1 -> Add(x)
  input image  /root/reference/alpha-zero/src/encoder.rs:10-46, /root/reference/environment/src/lib.rs:81-102
TensorFlow's LeakyRelu default alpha is 0.2 (the reference never sets it)."""
import numpy as np

SIDE, CELLS = 9, 81


def leaky_relu(x):
    return np.where(x > 0, x, 0.2 * x)


def conv2d_1x1(x, w, b):
    """x [B,H,W,Cin] NHWC, w [1,1,Cin,Cout], SAME (a 1x1 window never pads): out[..., o] = sum_c x[..., c] w[0,0,c,o] + b[o]."""
    out = np.zeros(x.shape[:3] + (w.shape[3],), dtype=x.dtype)
    for c in range(w.shape[2]):
        out += x[..., c:c + 1] * w[0, 0, c, :]
    return out + b


def depthwise_3x3_same(x, w):
    """DepthwiseConv2dNative, NHWC, stride 1, SAME, channel multiplier 1: out[b,y,x,c] = sum_{i,j} in[b, y+i-1, x+j-1, c] w[i,j,c,0];
    SAME with a 3x3 window and stride 1 pads one row / column of zeros on every side."""
    B, H, W, C = x.shape
    out = np.zeros_like(x)
    for i in range(3):
        for j in range(3):
            for y in range(H):
                yy = y + i - 1
                if yy < 0 or yy >= H:
                    continue
                for xx_out in range(W):
                    xx = xx_out + j - 1
                    if xx < 0 or xx >= W:
                        continue
                    out[:, y, xx_out, :] += x[:, yy, xx, :] * w[i, j, :, 0]
    return out


def encode_nn_input(board, black_to_move, opponent_mode=False):
    """One 243-float slot as encoder.rs writes it: `encode_board(perspective)` fills floats [0,162), two floats per cell --
    a black stone sets the float at offset 0 of its cell when the perspective turn is Black and at offset 1 when it is White,
    a white stone the other one (environment lib.rs:81-102) -- where the perspective is env.turn (EnvTurnMode::Player) or its
    opponent; then the turn plane, 1.0 when env.turn is Black, fills floats [162,243)."""
    slot = np.zeros(243, dtype=np.float64)
    perspective_is_black = black_to_move if not opponent_mode else not black_to_move
    black_offset, white_offset = (0, 1) if perspective_is_black else (1, 0)
    for index in range(CELLS):
        stone = int(board[index])  # 0 empty, 1 black, 2 white
        if stone == 1:
            slot[index * 2 + black_offset] = 1.0
        elif stone == 2:
            slot[index * 2 + white_offset] = 1.0
    slot[162:243] = 1.0 if black_to_move else 0.0
    return slot


def forward(params, slots):
    """params: the 31 tensors in checkpoint order (network.rs pushes them in this order); slots [B,243].  float64 throughout.
    Returns a dict of the layers the CUDA path can be inspected at."""
    it = iter([np.asarray(p, dtype=np.float64) for p in params])
    x = np.asarray(slots, dtype=np.float64).reshape(-1, SIDE, SIDE, 3)  # the placeholder's shape [-1, 9, 9, 3]
    B = x.shape[0]
    conv_w, conv_b = next(it), next(it)
    x = leaky_relu(conv2d_1x1(x, conv_w, conv_b))
    for _ in range(3):
        w0, b0, dw, pw, b1, w2, b2 = (next(it) for _ in range(7))
        h = leaky_relu(conv2d_1x1(x, w0, b0))
        h = depthwise_3x3_same(h, dw)
        h = leaky_relu(conv2d_1x1(h, pw, b1))
        h = conv2d_1x1(h, w2, b2)
        x = leaky_relu(h + x)
    tower = x.reshape(B, CELLS, 128)
    flat = x.reshape(B, 128 * SIDE * SIDE)  # Reshape of the NHWC tensor: row-major (y, x, c)
    fc0_w, fc0_b, fc1_w, fc1_b, v_w, v_b, p_w, p_b = (next(it) for _ in range(8))
    fc0 = leaky_relu(flat @ fc0_w + fc0_b)
    fc1 = leaky_relu(fc0 @ fc1_w + fc1_b)
    vlogit = (fc1 @ v_w + v_b).reshape(B)
    logits = fc1 @ p_w + p_b
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    return {"tower": tower, "fc0": fc0, "fc1": fc1, "logits": logits, "vlogit": vlogit, "P": e / e.sum(axis=1, keepdims=True),
            "V": np.tanh(vlogit)}

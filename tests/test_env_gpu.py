"""K1 parity (through the C ABI): the reference's own golden vectors, the derived KATs, and
BASELINE config 2 (4096 random-play boards) bit-exact against the oracle."""
import numpy as np
import pytest

from test_oracle_env import draw_moves, win81_moves

pytestmark = pytest.mark.gpu
IP, DRAW, BW, WW = 0, 1, 2, 3


@pytest.fixture(scope="module")
def ctx(omk):
    c = omk.Context(device=0, capacity_envs=1 << 16, capacity_trees=2, capacity_nodes=64, seed=0)
    yield c
    c.close()


def play(ctx, moves, slot=0):
    ctx.env_reset(ids=[slot])
    out = []
    for m in moves:
        st, legal = ctx.env_step([m], ids=[slot])
        out.append(int(st[0]))
    return out


def test_reference_golden_vectors(ctx, omk):
    # environment/src/lib.rs:201-252
    env = omk.Environment(ctx)
    assert env.turn == omk.Turn.Black
    for i in range(12):
        assert env.place_stone(i) == omk.GameStatus.InProgress
        assert env.board[i] == (omk.Stone.Black if i % 2 == 0 else omk.Stone.White)
        assert env.turn == (omk.Turn.White if i % 2 == 0 else omk.Turn.Black)
    # :255-298, :301-344
    assert play(ctx, [0, 9, 1, 10, 2, 11, 3, 12, 4]) == [IP] * 8 + [BW]
    assert play(ctx, [0, 2, 9, 11, 18, 20, 27, 29, 36]) == [IP] * 8 + [BW]
    # :347-358, :361-372
    assert play(ctx, list(range(36)) + [40])[-1] == BW
    assert play(ctx, list(range(36)) + [36])[-1] == BW


def test_reference_encoding_vectors(ctx, omk):
    # :375-426
    env = omk.Environment(ctx)
    env.place_stone(0)
    exp = np.zeros(162, np.float32)
    exp[0] = 1
    assert np.array_equal(env.encode_board(omk.Turn.Black), exp)
    env = omk.Environment(ctx)
    for m in [0, 10, 2, 30]:
        env.place_stone(m)
    exp = np.zeros(162, np.float32)
    exp[[0, 21, 4, 61]] = 1
    assert np.array_equal(env.encode_board(omk.Turn.Black), exp)
    exp = np.zeros(162, np.float32)
    exp[[1, 20, 5, 60]] = 1
    assert np.array_equal(env.encode_board(omk.Turn.White), exp)


def test_derived_kats(ctx, orc):
    # overline is not a win
    seq = []
    for b, w in zip([0, 1, 2, 4, 5], [18, 19, 20, 22, 23]):
        seq += [b, w]
    assert play(ctx, seq + [3]) == [IP] * 11
    # occupied cell -> None, no mutation
    ctx.env_reset(ids=[1])
    ctx.env_step([40], ids=[1])
    b0, t0, l0 = ctx.env_get(ids=[1])
    st, _ = ctx.env_step([40], ids=[1])
    assert st[0] == -1
    b1, t1, l1 = ctx.env_get(ids=[1])
    assert np.array_equal(b0, b1) and t0 == t1 and l0 == l1
    # draw on move 81; win on move 81 beats draw; white win; post-terminal mutation
    assert play(ctx, draw_moves()) == [IP] * 80 + [DRAW]
    assert play(ctx, win81_moves()) == [IP] * 80 + [BW]
    seq = []
    for b, w in zip([0, 1, 2, 3, 80], [9, 10, 11, 12, 13]):
        seq += [b, w]
    assert play(ctx, seq + [40]) == [IP] * 9 + [WW, IP]
    # run of nine is not five
    seq = []
    for b, w in zip([0, 1, 2, 3, 5, 6, 7, 8], [18, 20, 22, 24, 26, 36, 38, 40]):
        seq += [b, w]
    assert play(ctx, seq + [4])[-1] == IP


def test_legal_mask_and_fields_match_oracle_every_ply(ctx, orc):
    rng = np.random.default_rng(5)
    n = 512
    ctx.env_reset(n=n)
    envs = [orc.Environment() for _ in range(n)]
    for ply in range(90):
        acts = rng.integers(0, 81, n).astype(np.uint8)  # includes occupied cells -> None
        st, legal = ctx.env_step(acts)
        boards, turns, counts = ctx.env_get(n=n)
        for i, e in enumerate(envs):
            r = e.place_stone(int(acts[i]))
            assert (-1 if r is None else r) == st[i]
        ob = np.stack([e.board for e in envs])
        assert np.array_equal(ob, boards)
        assert np.array_equal(turns, [e.turn for e in envs])
        assert np.array_equal(counts, [e.legal_move_count for e in envs])
        mask = np.zeros((n, 3), np.uint32)
        for c in range(81):
            mask[:, c >> 5] |= (ob[:, c] == 0).astype(np.uint32) << np.uint32(c & 31)
        assert np.array_equal(mask, legal)


def test_config2_random_playout_4096_boards_bit_exact(ctx, orc):
    n, plies = 4096, 256
    a, s = ctx.env_random_playout(n, plies)
    oa, os_, ob, ot = orc.random_playout(n, plies, ctx.seed)
    assert np.array_equal(a, oa)
    assert np.array_equal(s, os_)
    boards, turns, _ = ctx.env_get(n=n)
    assert np.array_equal(boards, ob) and np.array_equal(turns, ot)
    assert (s != 0).sum() > n  # several finished games per board on average


def test_nn_input_image_matches_oracle(ctx, orc):
    rng = np.random.default_rng(9)
    n = 64
    ctx.env_reset(n=n)
    envs = [orc.Environment() for _ in range(n)]
    for ply in range(20):
        acts = rng.integers(0, 81, n).astype(np.uint8)
        ctx.env_step(acts)
        for i, e in enumerate(envs):
            e.place_stone(int(acts[i]))
    for mode in (0, 1):
        img = ctx.env_encode(n=n, mode=mode)
        ref = np.stack([e.encode_nn_input(mode) for e in envs])
        assert np.array_equal(img, ref)


def test_step_device_large_batch_idempotent_checksum(ctx):
    """Full-size property (no oracle): stepping the same action twice is None the second time and the
    board population grows by exactly one per successful step."""
    import torch

    n = 1 << 16
    ctx.env_reset(n=n)
    acts = torch.randint(0, 81, (n,), dtype=torch.uint8, device="cuda")
    st = torch.empty(n, dtype=torch.int8, device="cuda")
    legal = torch.empty((n, 3), dtype=torch.int32, device="cuda")
    ctx.env_step_device(acts.data_ptr(), n, st.data_ptr(), legal.data_ptr())
    ctx.synchronize()
    assert (st == 0).all()
    ctx.env_step_device(acts.data_ptr(), n, st.data_ptr(), legal.data_ptr())
    ctx.synchronize()
    assert (st == -1).all()
    _, _, counts = ctx.env_get(n=n)
    assert (counts == 80).all()

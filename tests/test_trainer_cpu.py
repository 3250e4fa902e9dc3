"""Trainer-side rows of SURVEY.md 8f (symmetries, replay memory, checkpoint format, loss, Adadelta step, data-parallel
gradient averaging) against the oracle restatement and the reference's own golden vectors.  CPU only."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
trainer = importlib.import_module("omok-ai_b200.trainer")
model_io = importlib.import_module("omok-ai_b200.model_io")
from oracle import net_oracle, trainer_oracle as TO  # noqa: E402


def test_oracle_symmetries_on_the_reference_golden_vectors():
    """src/utils.rs:70-108."""
    src = [1, 2, 3, 4]
    assert TO.rotate_90(src, 2) == [3, 1, 4, 2]
    assert TO.rotate_180(src, 2) == [4, 3, 2, 1]
    assert TO.rotate_270(src, 2) == [2, 4, 1, 3]
    assert TO.flip_horizontal(src, 2) == [2, 1, 4, 3]
    assert TO.flip_vertical(src, 2) == [3, 4, 1, 2]


def test_symmetries_match_the_oracle_and_compose():
    rng = np.random.default_rng(0)
    x = rng.integers(0, 3, size=(7, 81)).astype(np.uint8)
    for name, fn in zip(trainer.AUGMENT_ORDER, TO.SYMMETRY_FUNCS):
        got = trainer.apply_symmetry(name, x)
        for r in range(7):
            assert got[r].tolist() == fn(x[r].tolist(), 9), name
    r90 = lambda a: trainer.apply_symmetry("rotate_90", a)  # noqa: E731
    assert np.array_equal(r90(r90(x)), trainer.apply_symmetry("rotate_180", x))
    assert np.array_equal(r90(r90(r90(x))), trainer.apply_symmetry("rotate_270", x))
    assert np.array_equal(r90(r90(r90(r90(x)))), x)


def test_replay_memory_backfill_augmentation_order_and_cap():
    rng = np.random.default_rng(1)
    T = 5
    boards = np.zeros((T, 81), np.uint8)
    cells = rng.permutation(81)[:T]
    for t in range(1, T):
        boards[t] = boards[t - 1]
        boards[t, cells[t - 1]] = 1 + ((t - 1) % 2)
    pol = rng.random((T, 81)).astype(np.float32)
    mem = trainer.ReplayMemory(capacity=10_000)
    assert mem.add_episode(boards, pol, 1.0) == 6 * T
    ref = TO.episode_to_replay(boards.tolist(), pol.tolist(), 1.0)
    assert len(mem) == len(ref) == 30
    for i, (b, p, z) in enumerate(ref):
        assert mem.boards[i].tolist() == b and z == mem.zs[i]
        assert np.array_equal(mem.policies[i], np.array(p, np.float32))
    assert [mem.zs[t] for t in range(T)] == [1.0, -1.0, 1.0, -1.0, 1.0]
    assert [mem.turns[t] for t in range(T)] == [0, 1, 0, 1, 0]  # black first, alternating
    small = trainer.ReplayMemory(capacity=12)
    small.add_episode(boards, pol, 0.0)
    assert len(small) == 12 and small.zs[-1] == 0.0
    b, t, p, z = small.sample(128)
    assert b.shape == (12, 81) and p.shape == (12, 81) and len(set(map(bytes, p))) == 12  # without replacement


def test_split_episodes_cuts_games_at_terminal_status():
    P, G = 6, 2
    boards = np.arange(P * G * 81, dtype=np.int64).reshape(P, G, 81) % 3
    pol = np.random.default_rng(2).random((P, G, 81)).astype(np.float32)
    status = np.zeros((P, G), np.int8)
    status[2, 0] = 2   # game 0 ends on ply 2 with a win, restarts
    status[4, 1] = 1   # game 1 ends on ply 4 with a draw
    eps, carry = trainer.split_episodes(boards, pol, status)
    assert [(e[0].shape[0], e[2]) for e in eps] == [(3, 1.0), (5, 0.0)]
    assert len(carry[0][0]) == 3 and len(carry[1][0]) == 1
    status2 = np.zeros((1, G), np.int8)
    status2[0, 0] = 3
    eps2, _ = trainer.split_episodes(boards[:1], pol[:1], status2, carry)
    assert [(e[0].shape[0], e[2]) for e in eps2] == [(4, 1.0)]


def test_checkpoint_image_matches_bincode_and_roundtrips(tmp_path):
    rng = np.random.default_rng(3)
    params = [rng.standard_normal(s).astype(np.float32) for s in model_io.PARAM_SHAPES]
    small_names, small = ["a", "résumé"], [np.array([1.5, -2.0], np.float32), np.zeros(0, np.float32)]
    assert model_io.dumps(small, small_names) == TO.bincode_saved_data(small_names, small)
    # hand-written image of {["w"], [[1.0]]}: u64 1 | u64 1 "w" | u64 1 | u64 1 | 0x3f800000
    assert model_io.dumps([np.array([1.0], np.float32)], ["w"]) == bytes.fromhex(
        "0100000000000000" "0100000000000000" "77" "0100000000000000" "0100000000000000" "0000803f")
    path = tmp_path / "alpha-zero"
    model_io.save(path, params)
    names, got = model_io.load(path)
    assert names == model_io.VARIABLE_NAMES
    assert all(a.shape == tuple(s) and a.tobytes() == p.tobytes() for a, p, s in zip(got, params, model_io.PARAM_SHAPES))
    assert os.path.getsize(path) == 8 + sum(8 + len(n) for n in names) + 8 + sum(8 + 4 * p.size for p in params)
    with pytest.raises(model_io.ModelIOError):
        model_io.loads(open(path, "rb").read()[:-3])
    with pytest.raises(model_io.ModelIOError):
        model_io.loads(model_io.dumps(params[:-1] + [np.zeros(80, np.float32)]))


def test_encode_nn_input_matches_the_network_oracle_image():
    rng = np.random.default_rng(4)
    boards = np.zeros((20, 81), np.uint8)
    turns = np.zeros(20, np.uint8)
    for b in range(20):
        k = int(rng.integers(0, 50))
        for j, c in enumerate(rng.permutation(81)[:k]):
            boards[b, c] = 1 + (j % 2)
        turns[b] = k % 2
    for opp in (False, True):
        got = trainer.encode_nn_input(boards, turns, opponent_mode=opp)
        ref = np.stack([net_oracle.encode_image(b, int(t), opp) for b, t in zip(boards, turns)])
        assert np.array_equal(got, ref)
    # SURVEY 8c KAT: after moves [0, 10] (black to move, Player): 1.0 at floats 0, 21 and 162..242
    kat = np.zeros((1, 81), np.uint8)
    kat[0, 0], kat[0, 10] = 1, 2
    img = trainer.encode_nn_input(kat, np.array([0], np.uint8))[0]
    assert np.flatnonzero(img).tolist() == [0, 21] + list(range(162, 243))


def _batch(n, seed):
    rng = np.random.default_rng(seed)
    boards = (rng.random((n, 81)) < 0.3) * rng.integers(1, 3, size=(n, 81))
    turns = rng.integers(0, 2, size=n).astype(np.uint8)
    pi = rng.random((n, 81)).astype(np.float32)
    pi /= pi.sum(1, keepdims=True)
    z = rng.choice([-1.0, 0.0, 1.0], size=n).astype(np.float32)
    return trainer.encode_nn_input(boards.astype(np.uint8), turns), pi, z


def test_losses_match_the_oracle_network():
    params = net_oracle.random_params(0)
    images, pi, z = _batch(6, 5)
    step = trainer.TrainStep(params, dtype=torch.float64)
    with torch.no_grad():
        got = [float(x) for x in trainer.losses(step.params, step._t(images), step._t(pi), step._t(z))]
    ref = TO.losses_fp64(params, images, pi, z)
    assert np.allclose(got, ref, rtol=1e-6)  # the network oracle hands its float64 logits back as float32


def test_train_step_is_the_tensorflow_adadelta_update():
    params = net_oracle.random_params(1)
    images, pi, z = _batch(8, 6)
    step = trainer.TrainStep(params, dtype=torch.float64)
    ref = TO.Adadelta(params)
    for _ in range(2):  # two steps: the second one exercises both accumulators
        loss = trainer.losses(step.params, step._t(images), step._t(pi), step._t(z))[2]
        grads = [g.numpy() for g in torch.autograd.grad(loss, step.params)]
        expect = ref.step(grads)
        p_loss, v_loss, total = step.train(images, pi, z)
        for got, want in zip(step.params, expect):
            assert np.allclose(got.detach().numpy(), want, rtol=1e-10, atol=1e-14)
        after = TO.losses_fp64([p.astype(np.float32) for p in expect], images, pi, z)
        assert abs(total - (p_loss + v_loss)) < 1e-12 and np.allclose([p_loss, v_loss], after[:2], rtol=1e-4)
    assert step.steps == 2


def _dp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    params = net_oracle.random_params(2)
    images, pi, z = _batch(8, 7)
    lo, hi = rank * 4, rank * 4 + 4
    step = trainer.TrainStep(params, dtype=torch.float64)
    out = step.train(images[lo:hi], pi[lo:hi], z[lo:hi])
    ret[rank] = (step.numpy_params(), out)
    torch.distributed.destroy_process_group()


def test_data_parallel_step_equals_the_full_batch_step():
    """Config 5: two ranks, half a minibatch each, one all-reduce of the gradients == one step on the whole minibatch."""
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ret = mp.Manager().dict()
    mp.spawn(_dp_worker, args=(2, port, ret), nprocs=2, join=True)
    params = net_oracle.random_params(2)
    images, pi, z = _batch(8, 7)
    single = trainer.TrainStep(params, dtype=torch.float64)
    want_losses = single.train(images, pi, z)
    for rank in (0, 1):
        got_params, got_losses = ret[rank]
        for a, b in zip(got_params, single.numpy_params()):
            assert np.allclose(a, b, rtol=1e-6, atol=1e-9)
        assert np.allclose(got_losses, want_losses, rtol=1e-6)

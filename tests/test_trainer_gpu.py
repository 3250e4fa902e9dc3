"""SURVEY.md 8f on the GPU: the gradient step's weights feed the CUDA self-play kernels, and a whole (small) AlphaZero
iteration -- self-play -> episodes -> replay -> train -> weight refresh -- runs end to end."""
import importlib
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_trained_weights_reach_the_cuda_network(omk):
    import torch

    from oracle import net_oracle

    trainer = importlib.import_module("omok-ai_b200.trainer")
    params = net_oracle.random_params(3)
    rng = np.random.default_rng(0)
    boards = ((rng.random((32, 81)) < 0.3) * rng.integers(1, 3, size=(32, 81))).astype(np.uint8)
    turns = rng.integers(0, 2, size=32).astype(np.uint8)
    pi = rng.random((32, 81)).astype(np.float32)
    pi /= pi.sum(1, keepdims=True)
    z = rng.choice([-1.0, 1.0], size=32).astype(np.float32)
    images = trainer.encode_nn_input(boards, turns)
    step = trainer.TrainStep(params, device="cuda:0")
    before = [p.copy() for p in step.numpy_params()]
    l0 = step.train(images, pi, z)
    for _ in range(3):
        l1 = step.train(images, pi, z)
    assert l1[2] < l0[2]  # the same minibatch four times: the loss must fall
    assert any(not np.array_equal(a, b) for a, b in zip(before, step.numpy_params()))
    ctx = omk.Context(device=0, capacity_envs=1, capacity_trees=1, capacity_nodes=16, seed=0)
    step.sync_to(ctx)
    p, v = ctx.net_eval(boards, turns)
    p2, v2 = ctx.net_eval_images(images)
    assert p.tobytes() == p2.tobytes()
    with torch.no_grad():
        logits, tv = trainer.forward_logits([q.double() for q in step.params], torch.as_tensor(images, device="cuda:0").double())
        tp = torch.softmax(logits, 1).cpu().numpy()
    big = tp > 1e-12
    assert np.max(np.abs(p[big] - tp[big]) / tp[big]) < 1e-3  # north-star tolerance on priors
    assert np.max(np.abs(v - tv.reshape(-1).cpu().numpy())) < 1e-3
    ctx.close()


def test_one_small_iteration_end_to_end():
    """self-play (transitions streamed to the host) -> episodes -> replay with the 6x augmentation -> omk_train_step x 3
    -> the refreshed weights answer omk_net_eval_images, on one GPU."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    iteration = importlib.import_module("iteration")
    out = iteration.run(games_per_gpu=16, plies=30, count=32, batch=16, steps=3, minibatch=32, quiet=True)
    assert out["replay_transitions"] >= 32 and out["d2h_bytes"] == 16 * 30 * (81 + 81 * 4 + 4 + 1)
    assert out["policy_sums_to_one"] and np.isfinite(out["last_loss"]) and out["train_step_ms"] > 0
    assert out["selfplay_sims_per_s"] > 0 and out["train_step"].startswith("omk_train_step")

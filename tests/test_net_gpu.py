"""K3 parity: network forward vs the fp32 CPU oracle (oracle/net_oracle.py).
Tolerance (north_star): priors and values within 1e-3 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REL = 1e-3


def random_positions(n, seed, max_stones=60):
    rng = np.random.default_rng(seed)
    boards = np.zeros((n, 81), np.uint8)
    turns = np.zeros(n, np.uint8)
    for b in range(n):
        k = int(rng.integers(0, max_stones))
        for j, c in enumerate(rng.permutation(81)[:k]):
            boards[b, c] = 1 + (j % 2)
        turns[b] = k % 2
    return boards, turns


@pytest.fixture(scope="module")
def net(omk):
    from oracle import net_oracle

    c = omk.Context(device=0, capacity_envs=256, capacity_trees=64, capacity_nodes=2048, seed=3)
    params = net_oracle.random_params(0)
    c.net_load_params(params)
    yield c, params, net_oracle
    c.close()


def check(p, v, rp, rv, v_floor=1e-3):
    """1e-3 RELATIVE on priors and values.  Priors below 1e-12 (the softmax tail of logits ~ +-80) compare
    absolutely; values are relative down to `v_floor` (a pre-tanh error of 1e-4 on logits whose scale is ~20
    is fp32 rounding itself, so values closer to zero than the floor compare against the floor)."""
    big = rp > 1e-12
    assert np.max(np.abs(p[big] - rp[big]) / rp[big]) < REL
    assert np.max(np.abs(p[~big] - rp[~big])) < 1e-12 if (~big).any() else True
    assert np.max(np.abs(v - rv) / np.maximum(np.abs(rv), v_floor)) < REL
    assert np.allclose(p.sum(1), 1, atol=1e-4)


def test_forward_matches_oracle(net):
    ctx, params, no = net
    boards, turns = random_positions(200, 1)
    p, v = ctx.net_eval(boards, turns)
    rp, rv, _ = no.forward_boards(params, boards, turns, dtype=__import__("torch").float64)
    check(p, v, rp, rv)


def test_opponent_mode_and_images_path(net):
    ctx, params, no = net
    boards, turns = random_positions(40, 2)
    p, v = ctx.net_eval(boards, turns, mode=1)
    rp, rv, _ = no.forward_boards(params, boards, turns, opponent_mode=True, dtype=__import__("torch").float64)
    check(p, v, rp, rv)
    imgs = np.stack([no.encode_image(b, int(t)) for b, t in zip(boards, turns)])
    p2, v2 = ctx.net_eval_images(imgs)
    p3, v3 = ctx.net_eval(boards, turns, mode=0)
    assert p2.tobytes() == p3.tobytes() and v2.tobytes() == v3.tobytes()
    # arbitrary float images (AgentModel::evaluate_pv takes any tensor)
    rng = np.random.default_rng(0)
    fimgs = rng.standard_normal((16, 243)).astype(np.float32)
    p4, v4 = ctx.net_eval_images(fimgs)
    rp4, rv4, _ = no.forward(params, fimgs, dtype=__import__("torch").float64)
    check(p4, v4, rp4, rv4, v_floor=0.5)  # gaussian images: hidden activations ~10x larger than for boards


def test_batch_invariance(net):
    """A row's result must not depend on batch size or position (needed for recorded-mode parity)."""
    ctx, params, no = net
    boards, turns = random_positions(300, 3)
    p, v = ctx.net_eval(boards, turns)
    for lo, hi in ((0, 1), (17, 18), (5, 133), (299, 300)):
        q, w = ctx.net_eval(boards[lo:hi], turns[lo:hi])
        assert q.tobytes() == p[lo:hi].tobytes() and w.tobytes() == v[lo:hi].tobytes()
    # across the two fc0 paths: batches of <= 2048 rows take the split-K kernel (chunks reduced in order), larger ones the
    # CTA-pair kernel (chunks accumulated in registers); the sums are the same sums in the same order
    big_b, big_t = random_positions(4500, 33)
    P, V = ctx.net_eval(big_b, big_t)
    for lo, hi in ((0, 8), (100, 1124), (404, 4500), (4499, 4500)):
        q, w = ctx.net_eval(big_b[lo:hi], big_t[lo:hi])
        assert q.tobytes() == P[lo:hi].tobytes() and w.tobytes() == V[lo:hi].tobytes()


@pytest.mark.parametrize("n", [4500, 8040, 9400, 17000])
def test_fc0_tail_balancing_keeps_every_row(net, n):
    """The CTA-pair fc0 shares the K range of its last, partial wave of tiles with the otherwise idle pairs (fc_f16.cu FcBal:
    helper pairs compute the tail chunks, the tile's own pair adds them in order).  Every row must keep the bits it gets
    from the split-K path (batches of <= 2048 rows: every chunk a CTA, chunks added in order).  4500 rows = 36 tiles beside
    38 helper pairs, 8040 = a search lane's 64 tiles, 9400 = 74 tiles (a full wave: nothing to balance), 17000 = one whole
    wave + a balanced second one."""
    ctx, params, no = net
    boards, turns = random_positions(n, 1000 + n)
    ctx.debug_set_fc0_balance(True)  # opt-in (default off: see include/omok_b200.h)
    try:
        P, V = ctx.net_eval(boards, turns)
        P2, V2 = ctx.net_eval(boards, turns)  # again: re-armed counters, same scratch buffer
    finally:
        ctx.debug_set_fc0_balance(False)
    assert P2.tobytes() == P.tobytes() and V2.tobytes() == V.tobytes()
    P0, V0 = ctx.net_eval(boards, turns)  # the unbalanced pair kernel
    assert P0.tobytes() == P.tobytes() and V0.tobytes() == V.tobytes()
    for lo in range(0, n, 2048):  # ... and the split-K path
        hi = min(n, lo + 2048)
        q, w = ctx.net_eval(boards[lo:hi], turns[lo:hi])
        assert q.tobytes() == P[lo:hi].tobytes() and w.tobytes() == V[lo:hi].tobytes(), (n, lo)


@pytest.mark.parametrize("n", [1, 7, 8, 127, 128, 129, 255, 256, 257, 513])
def test_small_and_ragged_batches(omk, n):
    """Tile edges: batches below / at / above the 128-row MMA tile and the 256-row CTA-pair tile, each in a FRESH context
    so the workspace is sized by this very call (regression: a pair tile once stored past a 128-row workspace)."""
    import torch

    from oracle import net_oracle

    params = net_oracle.random_params(0)
    boards, turns = random_positions(n, 100 + n)
    c = omk.Context(device=0, capacity_envs=1, capacity_trees=1, capacity_nodes=16, seed=0)
    c.net_load_params(params)
    p, v = c.net_eval(boards, turns)
    p2, v2 = c.net_eval(boards, turns)  # same workspace again: results must not depend on stale rows
    c.close()
    assert p.tobytes() == p2.tobytes() and v.tobytes() == v2.tobytes()
    m = min(n, 48)
    rp, rv, _ = net_oracle.forward_boards(params, boards[-m:], turns[-m:], dtype=torch.float64)
    check(p[-m:], v[-m:], rp, rv)


def test_get_params_roundtrip_and_random_init(net, omk):
    ctx, params, no = net
    got = ctx.net_get_params()
    for a, b in zip(got, params):
        assert a.tobytes() == np.asarray(b, np.float32).tobytes()
    c2 = omk.Context(device=0, capacity_envs=1, capacity_trees=1, capacity_nodes=16, seed=0)
    c2.net_init_random(5)
    ws = c2.net_get_params()
    scales = no.init_scales()
    for (name, shape), w in zip(no.PARAM_SPECS, ws):
        assert w.shape == tuple(shape)
        if name in scales:  # w = N(0,1)*c : std ~ c
            assert abs(w.std() / scales[name] - 1) < (0.2 if w.size < 2000 else 0.05), name
            assert abs(w.mean()) < 4 * scales[name] / np.sqrt(w.size) + 1e-6
        else:
            assert not w.any()
    boards, turns = random_positions(8, 4)
    p, v = c2.net_eval(boards, turns)
    rp, rv, _ = no.forward_boards(ws, boards, turns, dtype=__import__("torch").float64)
    check(p, v, rp, rv)
    c2.close()


def test_search_with_real_net_recorded_mode(net, omk, orc):
    """MCTS visit counts bit-exact when the oracle is fed the GPU network's own outputs (SURVEY 8c 'recorded')."""
    ctx, params, no = net
    T = 6
    ev = orc.CallbackEvaluator(lambda b, t, m: ctx.net_eval(b, t, mode=m))
    streams = np.arange(T, dtype=np.uint32)
    agents = [orc.Agent(ev, ctx.seed, int(s)) for s in streams]
    ctx.pool_new_games(n=T, streams=streams, evaluator=omk.EVAL_NET)
    for rnd in range(2):
        orc.execute(agents, 160, 16, 0.25, 0.03, ev)
        ctx.pool_search(n=T, count=160, batch_size=16, epsilon=0.25, alpha=0.03, evaluator=omk.EVAL_NET)
        for t in range(T):
            ga, gn, gw, gp = ctx.pool_root_children(t)
            oa, on, ow, op = agents[t].root_children()
            assert np.array_equal(ga, oa) and np.array_equal(gn, on)
            assert gw.tobytes() == ow.tobytes() and gp.tobytes() == op.tobytes()
        acts, _ = ctx.pool_sample(n=T)
        assert [int(a) for a in acts] == [a.sample_action(0)[0] for a in agents]
        ctx.pool_play(acts, ids=list(range(T)))
        for a, act in zip(agents, acts):
            a.play_action(int(act))


def test_search_with_real_net_vs_cpu_net_tolerance(net, omk, orc):
    """Independent check: oracle search driven by the CPU fp32 network; priors/Q within 1e-3 where the visit
    pattern agrees (tiny float differences can legitimately flip an argmax, so counts are compared loosely)."""
    ctx, params, no = net
    ev = orc.TorchEvaluator(params)
    a = orc.Agent(ev, ctx.seed, 50)
    ctx.pool_new_games(ids=[10], streams=[50], evaluator=omk.EVAL_NET)
    pol_gpu = ctx.pool_root_stats(10)[4]
    pol_cpu = a.root_stats()[4]
    big = pol_cpu > 1e-12
    assert np.max(np.abs(pol_gpu[big] - pol_cpu[big]) / pol_cpu[big]) < REL


@pytest.mark.parametrize("tower_mode,fc_mode", [(1, 1), (1, 0), (0, 1), (0, 0)])
def test_all_kernel_paths_hold_the_tolerance(net, tower_mode, fc_mode):
    """The tcgen05 3xFP16 kernels (mode 1, the product path) and the fp32 CUDA-core kernels kept in the library as an A/B
    check (mode 0), for the tower and for fc0/fc1, in every combination, against the fp64 oracle; and the tensor-core
    tower / fc0 against the CUDA-core ones on the layer outputs themselves."""
    import torch

    ctx, params, no = net
    boards, turns = random_positions(160, 11)
    rp, rv, _ = no.forward_boards(params, boards, turns, dtype=torch.float64)

    def layer_outputs():  # (tower output, fc0 output) as the fc path in use sees them
        if fc_mode == 1:
            return ctx.debug_get_buffer(8, 160 * 10368).astype(np.float64), ctx.debug_get_buffer(9, 160 * 512).astype(np.float64)
        return ctx.debug_get_buffer(0, 160 * 10368).astype(np.float64), ctx.debug_get_buffer(1, 160 * 512).astype(np.float64)

    try:
        ctx.debug_set_tower_mode(tower_mode)
        ctx.debug_set_fc0_mode(fc_mode)
        p, v = ctx.net_eval(boards, turns)
        check(p, v, rp, rv)
        x, a1 = layer_outputs()
        ctx.debug_set_tower_mode(0)
        ctx.debug_set_fc0_mode(0)
        ctx.net_eval(boards, turns)
        x0 = ctx.debug_get_buffer(0, 160 * 10368).astype(np.float64)
        a0 = ctx.debug_get_buffer(1, 160 * 512).astype(np.float64)
        assert np.abs(x - x0).max() <= 2e-5 * np.abs(x0).max()
        assert np.abs(a1 - a0).max() <= 5e-5 * np.abs(a0).max()
    finally:
        ctx.debug_set_tower_mode(1)
        ctx.debug_set_fc0_mode(1)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 80, 443, 444, 445, 1000])
def test_f16_tower_is_independent_of_the_grid(omk, n):
    """k_tower16 walks position triples with CTA pairs; the result of a position must not depend on how many pairs are
    resident (1 pair = every triple on the same two CTAs, 5 = ragged tail, default = two CTAs per SM), including
    partial last triples."""
    import os

    from oracle import net_oracle

    params = net_oracle.random_params(0)
    boards, turns = random_positions(n, 300 + n)
    res = []
    for pairs in ("1", "5", None):
        if pairs:
            os.environ["OMK_TOWER_PAIRS"] = pairs
        c = omk.Context(device=0, capacity_envs=1, capacity_trees=1, capacity_nodes=16, seed=0)
        c.net_load_params(params)
        res.append(c.net_eval(boards, turns))
        c.close()
        os.environ.pop("OMK_TOWER_PAIRS", None)
    for r in res[1:]:
        assert res[0][0].tobytes() == r[0].tobytes() and res[0][1].tobytes() == r[1].tobytes()


def test_operand_overflow_is_a_loud_error(omk):
    """The tensor-core path splits activations into fp16 halves: |x| >= 65504 cannot be represented.  Such a network must
    fail with OMK_ERR_NUMERIC, never return garbage priors; and the context must keep working afterwards."""
    from oracle import net_oracle

    params = net_oracle.random_params(0)
    boards, turns = random_positions(5, 77)
    c = omk.Context(device=0, capacity_envs=1, capacity_trees=1, capacity_nodes=16, seed=0)
    bad = [p.copy() for p in params]
    bad[0] = bad[0] * np.float32(1e5)  # stem weights x 1e5: the residual stream leaves the fp16 range
    c.net_load_params(bad)
    with pytest.raises(omk.OmkError) as ei:
        c.net_eval(boards, turns)
    assert ei.value.code == -5
    c.net_load_params(params)
    p, v = c.net_eval(boards, turns)
    assert np.isfinite(p).all() and np.allclose(p.sum(1), 1, atol=1e-4)
    c.close()


def test_tensor_core_kernels_are_in_the_library(omk):
    """SASS evidence: tcgen05.mma -> UTC*MMA, TMA -> UTMALDG / UBLKCP, tcgen05.ld/st -> LDTM / STTM."""
    import subprocess

    sass = subprocess.run(["cuobjdump", "-sass", omk.lib_path()], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UBLKCP", "LDTM", "STTM"):
        assert mnemonic in sass, mnemonic


def banded_positions(n, seed):
    """Thirds of early (0-30 stones), middle (30-60) and late (60-80 stones) boards."""
    rng = np.random.default_rng(seed)
    boards = np.zeros((n, 81), np.uint8)
    turns = np.zeros(n, np.uint8)
    for b in range(n):
        lo, hi = ((0, 31), (30, 61), (60, 81))[b % 3]
        k = int(rng.integers(lo, hi))
        cells = rng.permutation(81)[:k]
        boards[b, cells[0::2]] = 1
        boards[b, cells[1::2]] = 2
        turns[b] = k % 2
    return boards, turns


@pytest.mark.parametrize("chunk,bar", [(9, 3.0e-4), (3, 2.4e-4)])
def test_large_sample_error_margin(omk, chunk, bar):
    """the margin to the 1e-3 bar on a LARGE sample -- 6 000 positions per weight seed including
    60-80-stone boards, both EnvTurnMode encodings, two seeds (profiles/r02_net_error_study.md has 20 000 x 3 seeds + a
    trained-scale set: max 2.1e-4 with the default chunking, 1.6e-4 with fc0_chunk = 3).  Priors: relative error
    below `bar` (3.3x / 4.2x inside the north-star tolerance).  Values: tanh of a logit of scale ~20 -- near a zero crossing a
    relative error is ill-conditioned for ANY fp32 forward, so values are held to 2e-4 absolute everywhere and to 1e-3
    relative where |v| >= 0.1."""
    import torch

    from oracle import net_oracle as no

    ctx = omk.Context(device=0, capacity_envs=1, capacity_trees=1, capacity_nodes=16, seed=0)
    ctx.debug_set_fc0_chunk(chunk)
    boards, turns = banded_positions(6000, 777)
    modes = (np.arange(6000) // 3) % 2
    worst = 0.0
    for seed in (0, 5):
        params = no.random_params(seed)
        ctx.net_load_params(params)
        for md in (0, 1):
            idx = np.flatnonzero(modes == md)
            p, v = ctx.net_eval(boards[idx], turns[idx], mode=md)
            imgs = np.stack([no.encode_image(b, int(t), bool(md)) for b, t in zip(boards[idx], turns[idx])])
            ref = no.forward_layers(params, imgs, torch.float64)
            big = ref["P"] > 1e-12
            rel = np.abs(p.astype(np.float64) - ref["P"])[big] / ref["P"][big]
            worst = max(worst, float(rel.max()))
            assert rel.max() < bar, f"seed {seed} mode {md}: max relative prior error {rel.max():.2e}"
            dv = np.abs(v.astype(np.float64) - ref["V"])
            assert dv.max() < 2e-4
            far = np.abs(ref["V"]) >= 0.1
            assert (dv[far] / np.abs(ref["V"][far])).max() < 1e-3
    ctx.close()
    print(f"fc0 chunk {chunk}: worst relative prior error {worst:.2e}")


def test_fc0_fine_chunk_keeps_batch_invariance(omk):
    """fc0_chunk = 3 (finer TMEM accumulation): rows stay bit-identical whatever the batch size and whichever fc0 path the
    batch takes (split-K up to 1024 rows with this chunking, CTA pairs above), and differ from the default chunking only
    in the last bits."""
    from oracle import net_oracle

    ctx = omk.Context(device=0, capacity_envs=1, capacity_trees=1, capacity_nodes=16, seed=0)
    ctx.net_load_params(net_oracle.random_params(0))
    boards, turns = random_positions(2600, 44)
    p9, v9 = ctx.net_eval(boards, turns)
    ctx.debug_set_fc0_chunk(3)
    P, V = ctx.net_eval(boards, turns)  # 2600 rows: CTA-pair path
    for lo, hi in ((0, 8), (3, 900), (1000, 2024), (1500, 2600), (2599, 2600)):  # split-K (<= 1024 rows) and pair sub-batches
        q, w = ctx.net_eval(boards[lo:hi], turns[lo:hi])
        assert q.tobytes() == P[lo:hi].tobytes() and w.tobytes() == V[lo:hi].tobytes()
    assert np.max(np.abs(P - p9) / np.maximum(p9, 1e-12)) < 1e-3 and P.tobytes() != p9.tobytes()
    with pytest.raises(omk.OmkError):
        ctx.debug_set_fc0_chunk(5)
    ctx.close()


def test_tower_biases_through_the_mma_survive_large_biases_and_tiny_weights(omk):
    """The tower's biases travel through the MMA as an fp16 hi/lo tile of b * 2^s / 64 (tower_f16.cu): a network whose
    1x1 weights are tiny (large power-of-two scale) and whose biases are of order one must not overflow that tile -- the
    scale is limited by the bias magnitude instead -- and must still meet the tolerance."""
    import torch

    from oracle import net_oracle as no

    rng = np.random.default_rng(5)
    params = no.random_params(2)
    names = [n for n, _ in no.PARAM_SPECS]
    for i, (name, shape) in enumerate(no.PARAM_SPECS):
        if name.startswith("res") and name[5:] in ("w0", "pw", "w2"):
            params[i] = (params[i] * np.float32(0.02)).astype(np.float32)      # tiny 1x1 weights
        if name.startswith("res") and name[5:] in ("b0", "b1", "b2"):
            params[i] = rng.uniform(-3.0, 3.0, size=shape).astype(np.float32)  # biases of order one
    # keep the logits in a sane range for a relative comparison of priors
    params[names.index("p_w")] = (params[names.index("p_w")] * np.float32(0.2)).astype(np.float32)
    ctx = omk.Context(device=0, capacity_envs=1, capacity_trees=1, capacity_nodes=16, seed=0)
    ctx.net_load_params(params)
    boards, turns = random_positions(300, 9)
    p, v = ctx.net_eval(boards, turns)
    rp, rv, _ = no.forward_boards(params, boards, turns, dtype=torch.float64)
    big = rp > 1e-12
    assert np.isfinite(p).all() and np.max(np.abs(p[big] - rp[big]) / rp[big]) < REL
    assert np.max(np.abs(v - rv)) < 5e-4
    imgs = np.stack([no.encode_image(b, int(t)) for b, t in zip(boards, turns)])
    p2, v2 = ctx.net_eval_images(imgs)
    assert p2.tobytes() == p.tobytes() and v2.tobytes() == v.tobytes()  # boards path (tables) == image path, bit for bit
    ctx.close()

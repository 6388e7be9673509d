"""GPU tests of the BASELINE configurations at size, of the sync-free / graph-captured train step and of multi-GPU parity.

* C1 (configs[0]) at its full 4096 rays against the oracle, fp32 (1e-5 / 2e-4) and bf16 (stated relative Frobenius);
* 300-step convergence: the bf16 loss curve stays inside a stated band around the fp32 one;
* a train step captured in a CUDA graph replays to the same parameters as the eager step, with lr / eps / alpha / timestep
  changing between replays; the loss value is bit-reproducible; no host synchronisation inside the step;
* `max_obj_rays` overflow is flagged on the device;
* 2 GPUs (skipped on a 1-GPU box): the all-reduced fp32 gradient equals the mean of the two single-rank gradients.
"""
import os
import socket

import numpy as np
import pytest
import torch

import durf_test_helpers as H
import ref_cases as C
from oracle import durf_oracle as O

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _model(**kw):
    from durf_b200.obbpose_model import MipNerfModel
    return MipNerfModel(**kw)


def test_c1_full_size_4096_rays_fp32_and_bf16():
    """BASELINE configs[0]: static background, no contraction, deterministic sampling, 4096 rays x (128 + 128) samples,
    compared with the oracle IN FULL (the oracle needs ~3 s for it)."""
    sc = H.scene(B=4096, K=1, seed=41, behind=True, far=40.0)
    cfg = O.ModelConfig(dynamics=False, contraction=False)
    want = O.model_forward(H.oracle_params(sc), H.oracle_rays(sc), torch.from_numpy(sc['ext']), 0, False, False, False, 10.0, cfg=cfg)
    rays, ext = H.cuda_rays(sc), cu(sc['ext'])
    m32 = _model(dynamics=False, contraction=False, precision='fp32')
    got = m32.apply(H.cuda_variables(sc, m32), None, rays, None, ext, torch.tensor([0]), False, False, False, 10.0)
    names = ('comp_rgb', 'distance', 'acc', 'weights', 't_vals', 't_mids', 't_dists')
    for lvl, (g, w) in enumerate(zip(got, want)):
        rt = 1e-5 if lvl == 0 else 2e-4
        for i, nm in enumerate(names):
            H.assert_close(g[i], w[i], rtol=rt if nm != 'weights' else 5 * rt, what=f"C1 fp32 level{lvl}.{nm}")
    m16 = _model(dynamics=False, contraction=False, precision='bf16')
    got = m16.apply(H.cuda_variables(sc, m16), None, rays, None, ext, torch.tensor([0]), False, False, False, 10.0)
    rel = lambda a, b: float((a.cpu() - b).norm() / b.norm())
    # bf16 tolerance of the raw outputs -> composited quantities, stated: rgb 5e-3, acc 1e-2, coarse weights 2e-2
    assert rel(got[0][0], want[0].comp_rgb) <= 5e-3 and rel(got[1][0], want[1].comp_rgb) <= 5e-3
    assert rel(got[0][2], want[0].acc) <= 1e-2
    assert rel(got[0][3], want[0].weights) <= 2e-2
    mse = float(((got[1][0].cpu() - want[1].comp_rgb) ** 2).mean())
    assert -10.0 * np.log10(max(mse, 1e-20)) >= 45.0


def _learnable_batches(B, K, steps, seed):
    """A fixed small scene whose targets are a smooth function of the ray (so the loss can actually go down)."""
    rng = np.random.default_rng(seed)
    sc = C._scene(B=B, K=K, seed=seed)
    out = []
    from durf_b200 import synthetic as S
    for i in range(steps):
        rays, _ = S.random_rays(rng, B, c2w=sc['c2w'], far=40.0)
        v = rays.viewdirs
        pixels = (0.5 + 0.5 * np.stack([np.sin(3 * v[:, 0]), np.cos(2 * v[:, 1]), np.sin(4 * v[:, 2] + 1.0)], -1)).astype(np.float32)
        depth = (6.0 + 3.0 * np.sin(5 * v[:, :1])).astype(np.float32)
        depth[rng.uniform(size=(B, 1)) < 0.3] = 0.0
        sky = np.where((depth == 0) & (rng.uniform(size=(B, 1)) < 0.5), 0.995, 0.0).astype(np.float32)
        out.append(dict(rays=rays, pixels=pixels, depth=depth, sky=sky,
                        t_rand=rng.uniform(size=(B, 129)).astype(np.float32), u_rand=rng.uniform(size=(B, 129)).astype(np.float32)))
    return sc, out


def test_convergence_bf16_tracks_fp32_over_300_steps():
    """Train the same model on the same 300 batches with precision='fp32' and 'bf16'.  Stated band: the 25-step moving
    average of the bf16 loss stays within 5 % (+1e-3) of the fp32 one over the whole run, and both fall by > 40 %."""
    from durf_b200.train import TrainState, train_step
    from durf_b200.utils import Config, Rays
    from durf_b200 import math as dmath
    B, K, steps = 384, 2, 300
    sc, batches = _learnable_batches(B, K, steps, seed=77)
    curves = {}
    for precision in ('fp32', 'bf16'):
        model = _model(precision=precision)
        v = H.cuda_variables(sc, model)
        state = TrainState.create(v)
        config = Config()
        losses = []
        for step, b in enumerate(batches, start=1):
            batch = dict(rays=Rays(*[cu(a) for a in b['rays']]), ext=cu(sc['ext']), ts=torch.tensor([step % 5]), pixels=cu(b['pixels']),
                         depth=cu(b['depth']), sky=cu(b['sky']))
            rng = dict(t_rand=cu(b['t_rand']), u_rand=cu(b['u_rand']))
            lr = dmath.learning_rate_decay(step, 2e-3, 2e-4, steps, 20, 0.1)
            state, st = train_step(model, config, rng, state, batch, lr=lr, eps=3.0, alpha=10.0)
            losses.append(st['loss'])
        curves[precision] = torch.stack(losses).double().cpu().numpy()
        assert np.isfinite(curves[precision]).all()
    k = np.ones(25) / 25
    a, b = np.convolve(curves['fp32'], k, 'valid'), np.convolve(curves['bf16'], k, 'valid')
    assert a[-1] < 0.6 * a[0] and b[-1] < 0.6 * b[0], (a[0], a[-1], b[0], b[-1])
    worst = float(np.max(np.abs(b - a) / (a + 1e-3 / 0.05)))
    assert worst <= 0.05, f"bf16 loss curve leaves the 5 % band around fp32: worst {worst:.3f}"


def _train_inputs(B, K, seed, pose_opt=False):
    sc = C._scene(B=B, K=K, seed=seed)
    tg = sc['targets']
    mk = lambda ts: dict(rays=H.cuda_rays(sc), ext=cu(sc['ext']), ts=torch.tensor([ts]), pixels=cu(tg['pixels']), depth=cu(tg['depth']),
                         sky=cu(tg['sky']))
    rng = dict(t_rand=cu(sc['t_rand']), u_rand=cu(sc['u_rand']))
    return sc, mk, rng


@pytest.mark.parametrize("pose_opt", [False, True])
def test_graph_captured_step_replays_like_eager(pose_opt):
    """GraphedTrainStep: capture once, replay 3 steps with different lr / eps / alpha / timestep; parameters after the 3 replays
    agree with 3 eager steps fed the same inputs (wgrad reduces with float atomics, so not bit-for-bit: 1e-5 of the update),
    the step counter travels with the other per-step scalars, and the library made no launch outside the graph during replay."""
    from durf_b200 import ops
    from durf_b200.train import TrainState, train_step, GraphedTrainStep
    from durf_b200.utils import Config
    B, K = 512, 2
    sc, mk, rng = _train_inputs(B, K, seed=91)
    kw = dict(no_pose_opt=not pose_opt, no_yaw_opt=not pose_opt)
    config = Config(tv_loss_mult=0.01 if pose_opt else 0.0)
    sched = [(1e-3, 3.0, 4.0, 2), (8e-4, 2.5, 6.5, 0), (5e-4, 2.0, 10.0, 4)]
    prev = cu(sc['centers'][1])[None]

    first = {}

    def eager(tag):
        model_e = _model(precision='bf16', **kw)
        v_e = H.cuda_variables(sc, model_e)
        st_e = TrainState.create(v_e)
        for i, (lr, eps, alpha, ts) in enumerate(sched):
            st_e, stats_e = train_step(model_e, config, rng, st_e, mk(ts), lr=lr, eps=eps, alpha=alpha, prev=prev)
            if i == 0:
                first[tag] = (stats_e['grad'].clone(), stats_e['loss'].clone())
        return v_e, stats_e
    v_e, stats_e = eager('a')
    v_e2, _ = eager('b')       # run-to-run noise of the eager step itself (wgrad / bias reductions use float atomics)

    model_g = _model(precision='bf16', **kw)
    v_g = H.cuda_variables(sc, model_g)
    st_g = TrainState.create(v_g)
    start = v_g.flat.clone()
    step = GraphedTrainStep(model_g, config, st_g, B, K, use_prev=pose_opt)
    for i, (lr, eps, alpha, ts) in enumerate(sched):
        if i == 1:
            ops.reset_launch_count()
        stats_g = step(mk(ts), lr, eps, alpha, rng=rng, prev=prev)
        if i == 0:
            first['g'] = (stats_g['grad'].clone(), stats_g['loss'].clone())
    torch.cuda.synchronize()
    # the FIRST step's raw gradient and loss: nothing amplifies summation-order noise yet
    ga, gb, gg = first['a'][0].double(), first['b'][0].double(), first['g'][0].double()
    n1 = float((ga - gb).norm() / ga.norm())
    r1 = float((ga - gg).norm() / ga.norm())
    assert r1 <= max(1e-5, 3.0 * n1), f"first replayed step: gradient differs from eager by {r1:.3e} (eager vs eager {n1:.3e})"
    assert torch.equal(first['a'][1], first['g'][1]), "the deterministic loss value must be bit-identical eager vs replay"
    assert ops.launch_count() == 0, "replays must not launch kernels from the host side of the library"
    assert int(step.scalars.step.item()) == 2 and st_g.step == 3      # the 0-based step the last replay's Adam used
    upd_e, upd_g = (v_e.flat - start).double(), (v_g.flat - start).double()
    assert float(upd_e.norm()) > 0
    rel = float((upd_e - upd_g).norm() / upd_e.norm())
    noise = float(((v_e2.flat - start).double() - upd_e).norm() / upd_e.norm())
    # three Adam steps from zero moments are sign-like in g where |g| ~ 1e-8, so summation-order noise is amplified: the graph
    # must be no further from an eager run than two eager runs are from each other (x5, one pair is a noisy estimate), floor 2e-3
    assert rel <= max(2e-3, 5.0 * noise), f"graph replay vs eager: {rel:.3e} of the 3-step update (eager vs eager: {noise:.3e})"
    assert abs(float(stats_g['loss']) - float(stats_e['loss'])) <= 1e-4 * max(1.0, abs(float(stats_e['loss'])))


@pytest.mark.parametrize("B", [300, 1500])
def test_overlapped_backward_equals_the_serial_backward(B, monkeypatch):
    """ops.OverlappedBackward (the dZ chain and the weight-gradient kernel of the background network on disjoint SMs, dZ
    blocks handed over through the tile_done counters) gives the gradient of the serial durf_mlp_bwd: same kernels, same
    operands, only the order of the fp32 reductions differs (<= 1e-5 relative, against the run-to-run noise of the serial
    path).  B = 300 leaves most weight-gradient CTAs waiting for tiles; B = 1500 spans several waves of the chain.  Also
    inside a captured CUDA graph (fork / join of the side stream)."""
    from durf_b200.train import TrainState, train_step, GraphedTrainStep
    from durf_b200.utils import Config
    sc, mk, rng = _train_inputs(B, 2, seed=97)
    config = Config()

    from durf_b200 import ops
    wcap = ops.OverlappedBackward().weight_ctas

    def grad(overlap, shared=False):
        model = _model(precision='bf16', overlap_backward=overlap, overlap_min_rays=0, shared_level_backward=shared)
        st = TrainState.create(H.cuda_variables(sc, model))
        _, stats = train_step(model, config, rng, st, mk(1), lr=1e-3, eps=3.0, alpha=10.0)
        torch.cuda.synchronize()
        return stats['grad'].double().clone(), stats['loss'].clone()
    # the serial reference runs its weight-gradient kernel on the same number of CTAs (same split of the tiles over fp32
    # accumulators); NaN-poisoned dZ workspaces turn any block read before it was written into a non-finite gradient
    monkeypatch.setenv('DURF_BWD_POISON', '1')
    monkeypatch.setenv('DURF_WGRAD_CTAS', str(wcap))
    g_serial, l_serial = grad(False)
    g_serial2, _ = grad(False)
    monkeypatch.delenv('DURF_WGRAD_CTAS')
    g_full, _ = grad(False)                        # all SMs: a different partition, fp32 summation-order noise only
    g_over, l_over = grad(True)
    assert float((g_serial - g_full).norm() / g_serial.norm()) <= 2e-4
    # the default path: both levels in one data-gradient and one weight-gradient launch (again another partition of the sums)
    g_shared, l_shared = grad(False, shared=True)
    assert float((g_serial - g_shared).norm() / g_serial.norm()) <= 2e-4 and torch.equal(l_serial, l_shared)
    noise = float((g_serial - g_serial2).norm() / g_serial.norm())
    rel = float((g_serial - g_over).norm() / g_serial.norm())
    assert float(g_serial.norm()) > 0 and torch.isfinite(g_over).all()
    assert rel <= max(1e-5, 3.0 * noise), f"overlapped vs serial backward: {rel:.3e} (serial vs serial: {noise:.3e})"
    assert torch.equal(l_serial, l_over)
    # captured: the fork / join of the side stream becomes two parallel branches of the graph
    model_g = _model(precision='bf16', overlap_backward=True, overlap_min_rays=0)
    st_g = TrainState.create(H.cuda_variables(sc, model_g))
    step = GraphedTrainStep(model_g, config, st_g, B, 2)
    stats_g = step(mk(1), 1e-3, 3.0, 10.0, rng=rng)
    torch.cuda.synchronize()
    rel_g = float((g_serial - stats_g['grad'].double()).norm() / g_serial.norm())
    assert rel_g <= max(1e-5, 3.0 * noise), f"graph-captured overlapped backward: {rel_g:.3e}"


@pytest.mark.parametrize("randomized", [False, True])
def test_fused_raymarch_is_the_separate_raymarch(randomized):
    """with `fuse_raymarch` the tcgen05 MLP kernel generates its own input tiles (mip.py:155-282, mip360.py:47-79
    inside K2) - the same device functions as the stand-alone ray-march kernel, so the whole model output must be IDENTICAL,
    bit for bit, to the path with separate durf_raymarch_fwd launches: background (contraction, ray multiplier) and object
    networks (weighted IPE on compacted hit rays), sampled and resampled levels, inference and the training forward (where
    the generated tile is also stored for the weight-gradient kernel)."""
    sc = H.scene(B=700, K=2, seed=41)
    outs = {}
    for fuse in (True, False):
        model = _model(precision='bf16', fuse_raymarch=fuse)
        v = H.cuda_variables(sc, model)
        rng = dict(t_rand=cu(sc['t_rand']), u_rand=cu(sc['u_rand'])) if randomized else None
        for training in (False, True):
            ctx = {} if training else None
            ret = model.apply(v, rng, H.cuda_rays(sc), None, cu(sc['ext']), torch.tensor([1]), randomized, False, False, 4.5, ctx=ctx)
            torch.cuda.synchronize()
            outs[(fuse, training)] = [(lv[0].clone(), lv[1].clone(), lv[2].clone(), lv[3].clone(), lv[4].clone()) for lv in ret]
            if training:
                outs[(fuse, 'feat')] = [lvl['feat_bg'].clone() for lvl in ctx['levels']]
    for training in (False, True):
        for a, b in zip(outs[(True, training)], outs[(False, training)]):
            for x, y, name in zip(a, b, ('comp_rgb', 'depth', 'acc', 'weights', 't_vals')):
                assert torch.equal(x, y), f"{name} differs between the fused and the separate ray-march (training={training}): " \
                                          f"max |d| = {float((x - y).abs().max()):.3e}"
    for x, y in zip(outs[(True, 'feat')], outs[(False, 'feat')]):
        assert torch.equal(x.view(torch.int16), y.view(torch.int16)), "stored feature tiles differ"


@pytest.mark.parametrize("randomized", [False, True])
def test_concurrent_object_networks_are_the_serial_ones(randomized):
    """`concurrent_objects`: the object networks' forward runs on side streams next to the background network, into compact
    rows (DurfMlpArgs.accumulate == 2, fenceposts re-formed in the kernel at level 0: DURF_RM_NO_TVALS_OUT) that
    durf_mlp_merge_raw adds in object order.  raw = MLP_0 + mask_0 * BoxMLP_0 + mask_1 * BoxMLP_1 in the reference's order
    (obbpose_model.py:203-204, 233-234), so model outputs, stored tiles and the gradient of a whole train step must be
    IDENTICAL to the serial accumulate == 1 path."""
    from durf_b200.train import TrainState, train_step
    from durf_b200.utils import Config
    sc = H.scene(B=700, K=2, seed=47)
    outs = {}
    for conc in (True, False):
        model = _model(precision='bf16', concurrent_objects=conc)
        v = H.cuda_variables(sc, model)
        rng = dict(t_rand=cu(sc['t_rand']), u_rand=cu(sc['u_rand'])) if randomized else None
        for training in (False, True):
            ctx = {} if training else None
            ret = model.apply(v, rng, H.cuda_rays(sc), None, cu(sc['ext']), torch.tensor([1]), randomized, False, False, 4.5, ctx=ctx)
            torch.cuda.synchronize()
            outs[(conc, training)] = [tuple(lv[i].clone() for i in range(5)) for lv in ret]
            if training:
                outs[(conc, 'feat')] = [o['feat'].clone() for lvl in ctx['levels'] for o in lvl['obj']]
    for training in (False, True):
        for a, b in zip(outs[(True, training)], outs[(False, training)]):
            for x, y, name in zip(a, b, ('comp_rgb', 'depth', 'acc', 'weights', 't_vals')):
                assert torch.equal(x, y), f"{name} differs (training={training}): max |d| = {float((x - y).abs().max()):.3e}"
    assert len(outs[(True, 'feat')]) == 4
    # compacted indices are unordered (warp-aggregated atomics): compare the object tiles as sets of rows per ray elsewhere;
    # here the whole step's gradient, which every stored record feeds
    sc2, mk, rng2 = _train_inputs(1024, 2, seed=95)
    grads = []
    for conc in (True, False):
        model = _model(precision='bf16', concurrent_objects=conc)
        st = TrainState.create(H.cuda_variables(sc2, model))
        _, stats = train_step(model, Config(), rng2, st, mk(1), lr=1e-3, eps=3.0, alpha=10.0)
        torch.cuda.synchronize()
        grads.append((stats['grad'].clone(), stats['loss'].clone()))
    assert torch.equal(grads[0][1], grads[1][1]), "loss differs"
    d = (grads[0][0] - grads[1][0]).abs().max()
    # the object networks' weight gradients sum their tiles in the order of the (unordered) compaction: equal up to fp32
    # summation order, and exactly equal for the background network
    assert float(d) <= 1e-6 * float(grads[1][0].abs().max()) + 1e-9, float(d)


def test_loss_value_is_bit_reproducible_and_step_has_no_host_sync():
    """The loss reduction is deterministic (fixed-order block partials instead of float atomics), and a bf16 train step issues
    no device->host read (checked with torch's sync debug mode)."""
    from durf_b200.train import TrainState, train_step
    from durf_b200.utils import Config
    sc, mk, rng = _train_inputs(1024, 2, seed=93)
    vals = []
    for _ in range(3):
        model = _model(precision='bf16')
        st = TrainState.create(H.cuda_variables(sc, model))
        _, stats = train_step(model, Config(), rng, st, mk(1), lr=1e-3, eps=3.0, alpha=10.0)
        vals.append(stats['loss'].clone())
    assert torch.equal(vals[0], vals[1]) and torch.equal(vals[1], vals[2]), [float(x) for x in vals]
    model = _model(precision='bf16')
    st = TrainState.create(H.cuda_variables(sc, model))
    batch = mk(1)
    train_step(model, Config(), rng, st, batch, lr=1e-3, eps=3.0, alpha=10.0)          # warm: allocator, weight images
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        train_step(model, Config(), rng, st, batch, lr=1e-3, eps=3.0, alpha=10.0)
    finally:
        torch.cuda.set_sync_debug_mode("default")
    torch.cuda.synchronize()


def test_max_obj_rays_overflow_is_flagged():
    sc, mk, rng = _train_inputs(1024, 2, seed=95)
    hits = None
    for cap, expect in ((1024, False), (4, True)):
        model = _model(precision='bf16', max_obj_rays=cap)
        v = H.cuda_variables(sc, model)
        ctx = {}
        b = mk(0)
        ret = model.apply(v, rng, b['rays'], None, b['ext'], b['ts'], True, False, False, 10.0, ctx=ctx)
        hits = int(ret[0][8].sum())
        flag = ctx['obj_overflow']
        assert (flag is not None and bool(flag)) == expect or (not expect and flag is None)
    assert hits > 8


# ------------------------------------------------------------------------------------------------ 2 GPUs
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _two_gpu_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from durf_b200.train import TrainState, train_step
    from durf_b200.utils import Config, Rays
    from durf_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        sc = C._scene(B=512, K=2, seed=97)
        tg = sc['targets']
        s, e = parallel.shard_range(512, rank, world)
        sl = lambda a: cu(a[s:e])
        model = _model(precision='fp32')
        v = H.cuda_variables(sc, model)
        st = TrainState.create(v)
        batch = dict(rays=Rays(*[sl(a) for a in sc['rays']]), ext=cu(sc['ext']), ts=torch.tensor([1]), pixels=sl(tg['pixels']),
                     depth=sl(tg['depth']), sky=sl(tg['sky']))
        rng = dict(t_rand=sl(sc['t_rand']), u_rand=sl(sc['u_rand']))
        cfg = Config(grad_max_val=0.0, grad_max_norm=0.0)
        _, single = train_step(model, cfg, rng, TrainState.create(H.cuda_variables(sc, model)), batch, lr=1e-3, eps=3.0, alpha=10.0)
        _, multi = train_step(model, cfg, rng, st, batch, lr=1e-3, eps=3.0, alpha=10.0, world_size=world)
        torch.cuda.synchronize()
        torch.save(dict(single=single['grad'].cpu(), multi=multi['grad'].cpu(), params=v.flat.cpu()), os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_allreduced_gradient_is_the_mean_of_rank_gradients(tmp_path):
    """jax.lax.pmean(grad, 'batch') (train_boxpose.py:253): each rank differentiates its own half of the batch (per-device
    normalisers, like pmap), the bucketed NCCL all-reduce + 1/world scale gives the mean of the two single-rank gradients, and
    both ranks hold bit-identical gradients and parameters afterwards."""
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_two_gpu_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    want = (outs[0]['single'] + outs[1]['single']) * 0.5
    assert float(outs[0]['single'].abs().max()) > 0 and not torch.equal(outs[0]['single'], outs[1]['single'])
    # `single` and the local half of `multi` are two executions of the backward pass (the box-pose and bias reductions use
    # float atomics), so the comparison with the mean holds to rounding; what must be exact is that every rank steps with the
    # IDENTICAL gradient and ends with identical parameters.
    err = float((outs[0]['multi'] - want).abs().max())
    assert err <= 1e-6 * float(want.abs().max()), err
    assert torch.equal(outs[0]['multi'], outs[1]['multi'])
    assert torch.equal(outs[0]['params'], outs[1]['params'])


def test_merged_launch_entry_points_equal_their_single_forms():
    """durf_compact_hits_all == durf_compact_hits per object (as sets: the lists are unordered), durf_mlp_pack_weights_multi ==
    durf_mlp_pack_weights per network (byte for byte), durf_mlp_merge_raw == an index_add of the compact rows (bit for bit,
    rows beyond the device count untouched), and the ray multiplier formed in the kernel (DURF_RM_MULT_IS_NHIT) == the
    multiplier 1 - nhit handed in (obbpose_model.py:205)."""
    from durf_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(5)
    B, K, N = 777, 3, 128
    hit = (torch.rand(B, K, device='cuda', generator=g) < 0.3).int()
    idx_all, cnt_all = ops.compact_hits_all(hit)
    for k in range(K):
        idx, cnt = ops.compact_hits(hit, k)
        n = int(cnt)
        assert n == int(cnt_all[k]) == int(hit[:, k].sum())
        assert torch.equal(idx[:n].sort().values, idx_all[k, :n].sort().values)
        assert torch.equal(idx_all[k, :n].sort().values, torch.nonzero(hit[:, k]).flatten().int())
    # pack: two topologies, three networks
    topos = [(60, 256, 8, 4, 27, 128), (63, 128, 8, 4, 27, 128), (63, 128, 8, 4, 27, 128)]
    blobs = [torch.randn(ops.mlp_param_count(t), device='cuda', generator=g) * 0.1 for t in topos]
    multi = ops.mlp_pack_multi(topos, blobs, [None] * 3)
    for t, b, m in zip(topos, blobs, multi):
        assert torch.equal(ops.mlp_pack(t, b), m)
    # merge
    rows, n_valid = 200, 150
    src_rgb = torch.randn(rows, N, 3, device='cuda', generator=g)
    src_den = torch.randn(rows, N, device='cuda', generator=g)
    raw_rgb = torch.randn(B, N, 3, device='cuda', generator=g)
    raw_den = torch.randn(B, N, device='cuda', generator=g)
    ray_index = torch.randperm(B, device='cuda', generator=g)[:rows].int()
    count = torch.tensor([n_valid], device='cuda', dtype=torch.int32)
    want_rgb, want_den = raw_rgb.clone(), raw_den.clone()
    want_rgb[ray_index[:n_valid].long()] += src_rgb[:n_valid]
    want_den[ray_index[:n_valid].long()] += src_den[:n_valid]
    ops.mlp_merge_raw(src_rgb, src_den, ray_index, count, raw_rgb, raw_den)
    assert torch.equal(raw_rgb, want_rgb) and torch.equal(raw_den, want_den)
    # multiplier from nhit
    sc = H.scene(B=300, K=2, seed=12)
    cr = H.cuda_rays(sc)
    nhit = torch.randint(0, 3, (300,), device='cuda', generator=g).float()
    kw = dict(near=cr.near, far=cr.far, contract=True)
    a = ops.raymarch(cr.origins, cr.directions, cr.radii, N, ray_mult=1.0 - nhit, **kw)
    b = ops.raymarch(cr.origins, cr.directions, cr.radii, N, ray_mult=nhit, ray_mult_is_nhit=True, **kw)
    assert torch.equal(a['features'], b['features']) and torch.equal(a['t_vals'], b['t_vals'])

"""CPU checks of the training-step oracle (BASELINE configs[3]) — the checker the backward kernels
of the next round will be held against: the sigma sampler against outputs of the reference's own functions, the loss
arithmetic, where gradients may flow, the literal `sample[ran_idx]` slicing, and data-parallel gradient averaging over
a world-size-2 gloo group."""
import json
import os
import socket

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "train_sigma_golden.json")
TINY = dict(in_channels=8, out_channels=4, block_out_channels=(32, 64, 64, 64), addition_time_embed_dim=32,
            projection_class_embeddings_input_dim=96, layers_per_block=1, cross_attention_dim=64,
            num_attention_heads=(1, 2, 2, 2), num_frames=2)


def _models(seed=0, randomize_zero_convs=True):
    from oracle.models import build_models
    unet, cnet = build_models(seed=seed, randomize_zero_convs=randomize_zero_convs, **TINY)
    unet.requires_grad_(False)          # train...cam_concat.py:984-987, 1101: only the ControlNet trains
    cnet.requires_grad_(True)
    return unet, cnet


def _batch(b=1, F=2, h=8, w=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    return dict(latents=torch.randn(b, F, 4, h, w, generator=g) * 0.18215 * 5,
                noise=torch.randn(b, F, 4, h, w, generator=g),
                sigmas=torch.tensor([1.3, 0.4, 7.0, 0.05][:b]),
                image_embeddings=torch.randn(b, 1, TINY["cross_attention_dim"], generator=g),
                trajectories=torch.rand(b, F, 3, 8 * h, 8 * w, generator=g) * 2 - 1,
                motion_values=torch.tensor([127.0, 90.0, 10.0, 200.0][:b]))


def test_sigma_sampler_matches_reference_functions():
    """tests/golden/train_sigma_golden.json = outputs of the reference's own stratified_uniform /
    rand_cosine_interpolated (scripts/train_svd_traj_VIPSeg_14_cam_concat.py:289-336)."""
    from oracle.train import logsnr_to_sigma, rand_cosine_interpolated
    for c in json.load(open(GOLDEN)):
        torch.manual_seed(c["seed"])
        sig = rand_cosine_interpolated([c["n"]])
        assert torch.allclose(sig, torch.tensor(c["sigmas"]), rtol=1e-6, atol=0), c["seed"]
        n = c["n"]
        u = (torch.arange(n, dtype=torch.float32) + torch.tensor(c["u"])) / n
        assert torch.allclose(logsnr_to_sigma(u), torch.tensor(c["sigmas"]), rtol=1e-6, atol=0)
    # the distribution covers (0.002, 700) monotonically in u
    s = logsnr_to_sigma(torch.tensor([1e-6, 0.5, 1 - 1e-6]))
    assert s[0] > 600 and s[2] < 0.0022 and s[0] > s[1] > s[2]


def test_time_ids_order_and_dropout_masks():
    from oracle.train import add_time_ids, dropout_masks
    ids = add_time_ids(6, torch.tensor([127.0, 90.0]), 0.02, 2)
    assert ids.tolist() == [[6.0, pytest.approx(0.02), 127.0], [6.0, pytest.approx(0.02), 90.0]]   # not [6, 128, 0.02]
    with pytest.raises(ValueError):
        add_time_ids(6, torch.tensor([1.0]), 0.02, 2)
    p = 0.1
    prompt, image = dropout_masks(torch.tensor([0.05, 0.15, 0.25, 0.5]), p)
    assert prompt.flatten().tolist() == [True, True, False, False]       # < 2p: drop the embedding
    assert image.flatten().tolist() == [1.0, 0.0, 0.0, 1.0]              # p <= r < 3p: drop the conditioning latent


def test_loss_arithmetic_and_gradient_flow():
    from oracle.train import training_step
    unet, cnet = _models()
    batch = _batch()
    out = training_step(unet, cnet, ran_idx=1, **batch)
    # loss_main recomputed from the returned prediction with the formulas of :1423-1436
    s = batch["sigmas"].reshape(-1, 1, 1, 1, 1)
    noisy = batch["latents"] + batch["noise"] * s
    den = out["model_pred"] * (-s / (s ** 2 + 1) ** 0.5) + noisy / (s ** 2 + 1)
    want = (((1 + s ** 2) / s ** 2) * (den - batch["latents"]) ** 2).mean()
    assert torch.allclose(out["loss_main"], want, rtol=1e-5)
    assert torch.allclose(out["loss"], out["loss_main"] + 0.5 * out["loss_spatial"])
    out["loss"].backward()
    assert all(p.grad is None for p in unet.parameters())
    g = {n: p.grad for n, p in cnet.named_parameters()}
    assert all(v is not None and torch.isfinite(v).all() for v in g.values())
    assert sum(float(v.abs().sum()) for v in g.values()) > 0
    # a perfect v-prediction gives zero loss: v* = (noisy c_skip - z) / (-c_out)
    v_star = (noisy / (s ** 2 + 1) - batch["latents"]) / (s / (s ** 2 + 1) ** 0.5)
    den = v_star * (-s / (s ** 2 + 1) ** 0.5) + noisy / (s ** 2 + 1)
    assert float(((den - batch["latents"]) ** 2).max()) < 1e-9


def test_zero_init_controlnet_blocks_gradients_behind_the_zero_convs():
    """With the faithful zero-initialised zero-convs (controlnet_sdv.py:394-410) the residuals are zero, the zero-convs
    receive gradients and everything behind them receives exactly none — the property that makes ControlNet training
    start from the frozen model."""
    from oracle.train import training_step
    unet, cnet = _models(randomize_zero_convs=False)
    training_step(unet, cnet, use_spatial=False, **_batch())["loss"].backward()
    zero_names = ("controlnet_down_blocks", "controlnet_mid_block")
    behind = [n for n, p in cnet.named_parameters() if not n.startswith(zero_names)]
    front = [n for n, p in cnet.named_parameters() if n.startswith(zero_names) and n.endswith("weight")]
    grads = dict((n, p.grad) for n, p in cnet.named_parameters())
    assert all(grads[n] is None or float(grads[n].abs().max()) == 0.0 for n in behind)
    assert any(float(grads[n].abs().max()) > 0 for n in front)


def test_spatial_pass_slices_the_flattened_axis():
    """`sample[ran_idx]` indexes the flattened (b*F) axis of the residuals (:1449-1452): for b = 1 that is frame ran_idx,
    for b = 2 and ran_idx = 1 it is still sample 0's frame 1, broadcast to both samples (replicated literally)."""
    from oracle.train import training_step
    unet, cnet = _models()
    one = training_step(unet, cnet, ran_idx=1, **_batch(b=1))
    assert torch.isfinite(one["loss_spatial"])
    two = training_step(unet, cnet, ran_idx=1, **_batch(b=2))
    assert torch.isfinite(two["loss_spatial"]) and two["model_pred"].shape[0] == 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dp_worker(rank, port, q):
    import torch.distributed as dist
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE="2")
    dist.init_process_group("gloo", rank=rank, world_size=2)
    try:
        from oracle.train import training_step
        torch.set_num_threads(2)
        unet, cnet = _models()
        batch = _batch(b=1, seed=10 + rank)               # every rank its own sample
        training_step(unet, cnet, ran_idx=0, **batch)["loss"].backward()
        flat = torch.cat([p.grad.reshape(-1) for p in cnet.parameters()])
        dist.all_reduce(flat)                              # the NCCL all-reduce of configs[3], here over gloo
        flat /= 2
        q.put((rank, flat[:: max(1, flat.numel() // 4096)].clone().numpy(), float(flat.norm())))
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_average_world2():
    """configs[3] is data-parallel over samples: the all-reduced mean of the per-rank ControlNet gradients equals the
    mean of the per-sample gradients computed in one process (training DP row)."""
    import torch.multiprocessing as mp
    from oracle.train import training_step
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    try:
        for _ in range(2):
            r, sample, norm = q.get(timeout=400)
            res[r] = (torch.from_numpy(sample), norm)
    finally:
        for p in procs:
            p.join(timeout=180)
            if p.is_alive():
                p.kill()
    grads = []
    threads = torch.get_num_threads()
    torch.set_num_threads(2)            # same thread count as the workers: oneDNN's summation order depends on it
    try:
        for r in range(2):
            unet, cnet = _models()
            training_step(unet, cnet, ran_idx=0, **_batch(b=1, seed=10 + r))["loss"].backward()
            grads.append(torch.cat([p.grad.reshape(-1) for p in cnet.parameters()]))
    finally:
        torch.set_num_threads(threads)
    mean = (grads[0] + grads[1]) / 2
    step = max(1, mean.numel() // 4096)
    for r in range(2):
        assert torch.allclose(res[r][0], mean[::step], rtol=1e-3, atol=1e-6)
        assert res[r][1] == pytest.approx(float(mean.norm()), rel=1e-3)
    assert torch.equal(res[0][0], res[1][0])


def test_oracle_reproduces_committed_training_golden():
    """tests/golden/train_golden.safetensors (losses, prediction, ControlNet gradient norms and a few full gradients of
    one training step, minted by tests/golden/gen_train_golden.py) is what the backward kernels will be checked against;
    here: the oracle still produces it."""
    import importlib.util
    from safetensors.torch import load_file
    here = os.path.dirname(__file__)
    spec = importlib.util.spec_from_file_location("gen_train_golden", os.path.join(here, "golden", "gen_train_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    threads = torch.get_num_threads()
    try:
        res, names = mod.run()
    finally:
        torch.set_num_threads(threads)
    gold = load_file(os.path.join(here, "golden", "train_golden.safetensors"))
    assert set(gold) == set(res) and len(names) == gold["grad_norms"].numel()
    for k in ("loss", "loss_main", "loss_spatial"):
        assert torch.allclose(res[k], gold[k], rtol=1e-4), k
    assert torch.allclose(res["model_pred"], gold["model_pred"], rtol=1e-3, atol=1e-4)
    assert torch.allclose(res["grad_norms"], gold["grad_norms"], rtol=2e-3, atol=1e-7)
    for k in gold:
        if k.startswith("grad."):
            rel = (res[k] - gold[k]).norm() / gold[k].norm().clamp_min(1e-12)
            assert rel < 2e-3, (k, float(rel))
    # every ControlNet parameter gets a gradient, and the mix factors (scalars) are among them
    assert float(gold["grad_norms"].min()) >= 0 and float((gold["grad_norms"] > 0).float().mean()) > 0.95

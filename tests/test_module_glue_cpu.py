"""CPU: the Python glue of the module-level seams (RaySamples / PDFSampler / ProposalNetworkSampler / NeuRADField /
NeuRADProposalField mirrors, the per-module walk of NeuRADModel.get_nff_outputs(fused=False), B200Backend.field_forward's
composition) run over tests/fake_backend.py -- the new device code through the host emulation, the older operators through
the oracle -- against the reference's golden values.  The same bodies run on the GPU in test_zz_module_seams_gpu.py."""
import pytest

from tests import module_seam_cases as C
from tests.fake_backend import FakeBackend


@pytest.fixture(autouse=True)
def fake_backend(monkeypatch):
    from neurad_studio_b200 import nerfstudio_api

    be = FakeBackend()
    monkeypatch.setattr(nerfstudio_api, "get_backend", lambda device: be)
    return be


@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz", "nff_sharp.npz"])
def test_field_forward_glue(name):
    C.field_forward_matches_reference_golden(name, "cpu")


@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz"])
def test_proposal_density_and_encoding_glue(name):
    C.proposal_density_and_encoding_match_reference_golden(name, "cpu")


@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz", "nff_sharp.npz"])
def test_proposal_sampler_and_module_walk_glue(name):
    C.proposal_sampler_and_module_walk_match_reference_golden(name, "cpu")


@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz"])
def test_training_gradients_glue(name):
    C.training_gradients_match_reference_golden(name, "cpu")


def test_proposal_density_backward_clamp_glue():
    C.proposal_density_backward_clamps_like_trunc_exp("cpu")


def test_backward_stage_operators_glue():
    C.backward_stage_operators_match_torch_autograd("cpu")


def test_stratified_sampling_glue():
    C.stratified_sampling_matches_reference_golden("cpu")


@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz"])
def test_training_mode_walk_glue(name):
    C.training_mode_walk_runs(name, "cpu")


def test_training_losses_glue():
    C.training_losses_match_reference_golden("cpu")


def test_metric_entry_points_glue():
    C.metric_entry_points("cpu")


def test_lidar_carving_masks_glue():
    C.lidar_carving_masks_and_training_outputs("cpu")


def test_get_outputs_and_decode_features_glue():
    C.get_outputs_and_decode_features("cpu")


def test_train_mode_encoding_glue():
    C.train_mode_encoding_matches_reference_golden("cpu")


def test_train_probe_step_glue():
    """tools/train_probe.py's training step (module walk, both regularisers, carving terms, backward) over the fake
    backend: keeps the round-2 measurement script from rotting."""
    import os
    import sys

    import torch

    import neurad_studio_b200 as nsb
    from neurad_studio_b200 import scene
    from neurad_studio_b200.nerfstudio_api import NeuRADModel

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import train_probe as T

    cfg = nsb.small_config(n_actors=2, log2_main=10, log2_prop=10)
    trajs = scene.make_trajectories(2, cfg.duration)
    model = NeuRADModel(cfg, trajs)
    model.load_reference_state_dict(scene.make_params(cfg, seed=1, beta=3.0, sdf_bias=0.6, trajectories=trajs))
    model.requires_grad_(True)
    model.train()
    rays, rb = T.make_batch(cfg, 40, 24, trajs, "cpu")
    with torch.no_grad():
        ref = model.get_nff_outputs(rb, fused=True)
    out, loss = T.step(model, rb, {"features": ref["features"] + 0.1, "depth": ref["depth"] * 1.1})
    loss.backward()
    assert torch.isfinite(loss) and model._param("field.mlp_geo.layers.0.weight").grad.abs().max().item() > 0
    assert model._param("proposal_fields.1.density_decoder.weight").grad.abs().max().item() > 0


def test_few_optimizer_steps_reduce_the_loss(fake_backend):
    """End to end: module walk + regularisers + loss.backward() + Adam on every trained tensor (tables, MLPs, decoders,
    beta, trajectories), re-binding the updated parameters each step (version tracking in NeuRADModel._bind): the loss of
    a small fitting problem goes down."""
    import torch

    import neurad_studio_b200 as nsb
    from neurad_studio_b200 import losses as L
    from neurad_studio_b200 import scene
    from neurad_studio_b200.nerfstudio_api import NeuRADModel, RayBundle

    cfg = nsb.small_config(n_actors=2, log2_main=10, log2_prop=10)
    trajs = scene.make_trajectories(2, cfg.duration)
    teacher = scene.make_params(cfg, seed=1, beta=3.0, sdf_bias=0.6, trajectories=trajs, table_scale=1.0)
    rays = scene.random_rays(48, cfg, seed=3, trajectories=trajs)
    rb = RayBundle(origins=rays["origins"], directions=rays["directions"], pixel_area=rays["pixel_area"], times=rays["times"],
                   metadata={"is_lidar": rays["is_lidar"], "sensor_idxs": rays["sensor_idx"]})
    model = NeuRADModel(cfg, trajs)
    model.load_reference_state_dict(teacher)
    with torch.no_grad():
        target = model.get_nff_outputs(rb, fused=False)
    student = {k: (v + 0.05 * torch.randn(v.shape, generator=torch.Generator().manual_seed(9)) if v.dtype.is_floating_point and
                   (k.endswith("hash_table") or ".layers." in k) else v) for k, v in teacher.items()}
    model.load_reference_state_dict(student)
    model.requires_grad_(True)
    binds = []
    orig = fake_backend.load_params
    fake_backend.load_params = lambda *a, **k: (binds.append(1), orig(*a, **k))[1]
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=2e-3)
    history = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        out = model.get_nff_outputs(rb)
        loss = ((out["features"] - target["features"]) ** 2).mean() + 1e-4 * (out["depth"] - target["depth"]).abs().mean()
        loss = loss + 1e-3 * L.zipnerf_interlevel_loss(out["weights_list"], out["ray_samples_list"])
        loss = loss + 2e-3 * L.distortion_loss(out["weights_list"], out["ray_samples_list"])
        loss.backward()
        opt.step()
        history.append(float(loss.detach()))
    assert history[-1] < 0.8 * history[0], history
    assert len(binds) == 6  # parameters are re-bound exactly once per step (after the optimizer changed them)

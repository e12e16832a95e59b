"""CPU: the device code of the BACKWARD operators (csrc/nff_modules.h) run by the host emulation, against
torch autograd through the oracle (whose forward is pinned bit-for-bit to the reference, so its autograd is the
reference's autograd in torch mode)."""
import pytest
import torch

from oracle import neurad_oracle as O
from oracle.convert import to_oracle_cfg
from tests.helpers import cfg_from_meta, load_golden
from tests.host_emul import emul


def rel_to_max(a, b):
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)


def test_weights_backward_rows():
    gen = torch.Generator().manual_seed(3)
    n, s = 37, 45
    alphas = (torch.rand(n, s, generator=gen) * 0.9).requires_grad_(True)
    g = torch.randn(n, s, generator=gen)
    (O.render_weight_from_alpha(alphas) * g).sum().backward()
    assert rel_to_max(emul.weights_bwd(True, alphas.detach(), None, g), alphas.grad) < 1e-5
    deltas = torch.rand(n, s, generator=gen) + 0.05
    dens = (torch.rand(n, s, generator=gen) * 2).requires_grad_(True)
    (O.weights_from_density(deltas, dens) * g).sum().backward()
    assert rel_to_max(emul.weights_bwd(False, deltas, dens.detach(), g), dens.grad) < 1e-5
    # a saturated sample (alpha == 1) gives finite gradients (the guarded quotient)
    a1 = alphas.detach().clone()
    a1[:, 7] = 1.0
    assert torch.isfinite(emul.weights_bwd(True, a1, None, g)).all()


@pytest.mark.parametrize("k,n,relu", [(32, 33, False), (48, 32, True), (64, 64, True), (6, 1, False), (50, 57, True)])
def test_linear_wgrad_tiling(k, n, relu):
    gen = torch.Generator().manual_seed(k * 100 + n)
    rows = 32 * 7 + 5
    x, dy = torch.randn(rows, k, generator=gen), torch.randn(rows, n, generator=gen)
    dW, db = torch.zeros(n, k), torch.zeros(n)
    emul.linear_wgrad(x, dy, relu, dW, db, n_ctas=3)
    xa = torch.relu(x) if relu else x
    assert rel_to_max(dW, dy.t() @ xa) < 1e-5 and rel_to_max(db, dy.sum(0)) < 1e-5
    emul.linear_wgrad(x, dy, relu, dW, db, n_ctas=1)  # accumulates
    assert rel_to_max(dW, 2 * (dy.t() @ xa)) < 1e-5


def _grad_params(p, prefix, n_actors):
    """Leaf copies (requires_grad) of one field's tables (+ decoder) inside a copy of the parameter dict."""
    q = dict(p)
    keys = [f"{prefix}.hashgrid.static_grid.hash_table"] + [f"{prefix}.hashgrid.actor_grids.{a}.hash_table" for a in range(n_actors)]
    if f"{prefix}.density_decoder.weight" in p:
        keys.append(f"{prefix}.density_decoder.weight")
    for k in keys:
        q[k] = p[k].clone().requires_grad_(True)
    return q, keys


@pytest.fixture(params=["fast", "generic"])
def bwd_variant(request):
    """Both variants of neurad_encoding_bwd_kernel: the register-resident fast paths (NeuRAD's grid shapes) and the generic
    one (any L x F)."""
    emul.set_bwd_generic(request.param == "generic")
    yield request.param
    emul.set_bwd_generic(False)


@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz"])
def test_encoding_backward_features_mode(name, bwd_variant):
    meta, g = load_golden(name)
    cfg = cfg_from_meta(meta)
    ocfg = to_oracle_cfg(cfg)
    p, r, ref = g["param"], g["ray"], g["ref"]
    n = r["origins"].shape[0]
    starts, ends = ref["starts"].reshape(n, -1), ref["ends"].reshape(n, -1)
    s = starts.shape[1]
    lidar = r["is_lidar"].reshape(-1).bool()
    area = r["pixel_area"].reshape(-1) * torch.where(lidar, 1.0, float(cfg.rgb_upsample_factor**2))
    q, keys = _grad_params(p, "field", meta["n_actors"])
    mean, std = O.fast_isotropic_gaussian(r["origins"][:, None, :], r["directions"][:, None, :], area[:, None, None],
                                          starts[..., None], ends[..., None])
    t = r["times"].reshape(n, 1, 1).expand(n, s, 1)
    feats, _ = O.hashgrid_forward(q, "field", ocfg.main, ocfg, mean, std, t, None)
    G = torch.randn(feats.shape, generator=torch.Generator().manual_seed(1))
    (feats * G).sum().backward()
    em_mean, em_std = emul.gaussian(r["origins"], r["directions"], area, torch.cat([starts, ends[:, -1:]], 1))
    grads = {"static": torch.zeros_like(p[keys[0]]), "actors": [torch.zeros_like(p[k]) for k in keys[1:]]}
    emul.encoding_bwd(cfg, p, O.pdf_u, 0, em_mean, em_std, r["times"], grads, dfeatures=G)
    assert rel_to_max(grads["static"], q[keys[0]].grad) < 1e-4
    hit = 0
    for a, k in enumerate(keys[1:]):
        want = q[k].grad if q[k].grad is not None else torch.zeros_like(p[k])
        assert (grads["actors"][a] - want).abs().max().item() <= 1e-4 * max(want.abs().max().item(), 1e-6)
        hit += int(want.abs().max().item() > 0)
    assert hit > 0 or meta["n_actors"] == 0


@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz"])
def test_encoding_backward_density_mode(name, bwd_variant):
    meta, g = load_golden(name)
    cfg = cfg_from_meta(meta)
    ocfg = to_oracle_cfg(cfg)
    p, r, ref = g["param"], g["ray"], g["ref"]
    n = r["origins"].shape[0]
    edges = ref["bins_e_1"].reshape(n, -1)
    lidar = r["is_lidar"].reshape(-1).bool()
    area = r["pixel_area"].reshape(-1) * torch.where(lidar, 1.0, float(cfg.rgb_upsample_factor**2))
    q, keys = _grad_params(p, "proposal_fields.1", meta["n_actors"])
    dens = O.proposal_density(q, 1, ocfg, r["origins"], r["directions"], area, r["times"].reshape(-1), edges[:, :-1], edges[:, 1:])
    G = torch.randn(dens.shape, generator=torch.Generator().manual_seed(2))
    (dens * G).sum().backward()
    em_mean, em_std = emul.gaussian(r["origins"], r["directions"], area, edges)
    n_act = meta["n_actors"]
    grads = {"static": torch.zeros_like(p[keys[0]]), "actors": [torch.zeros_like(p[k]) for k in keys[1:1 + n_act]],
             "decoder": torch.zeros(p[keys[-1]].numel())}
    emul.encoding_bwd(cfg, p, O.pdf_u, 2, em_mean, em_std, r["times"], grads, density=dens.detach(), ddensity=G)
    assert rel_to_max(grads["static"], q[keys[0]].grad) < 1e-4
    assert rel_to_max(grads["decoder"], q[keys[-1]].grad.reshape(-1)) < 1e-4
    for a, k in enumerate(keys[1:1 + n_act]):
        want = q[k].grad if q[k].grad is not None else torch.zeros_like(p[k])
        assert (grads["actors"][a] - want).abs().max().item() <= 1e-4 * max(want.abs().max().item(), 1e-6)


@pytest.mark.parametrize("field,mode", [(0, "features"), (2, "density")])
def test_encoding_backward_aggregation_equals_per_sample_scatter(field, mode):
    """The kernel's device functions (contiguous ray segments, run-length aggregation of the coarse cells, x-pair vector
    reductions) against the per-sample generic scatter on rays whose consecutive samples SHARE coarse cells (5 cm steps):
    the same sums in a different order."""
    meta, g = load_golden("nff_actors.npz")
    cfg = cfg_from_meta(meta)
    p, r = g["param"], g["ray"]
    n, s = 24, 64
    gen = torch.Generator().manual_seed(11)
    o, d = r["origins"][:n], r["directions"][:n]
    t = 0.5 + 0.05 * torch.arange(s + 1).float()[None, :] + torch.rand(n, 1, generator=gen)
    area = r["pixel_area"].reshape(-1)[:n]
    mean, std = emul.gaussian(o, d, area, t)
    prefix = "field" if field == 0 else "proposal_fields.1"
    keys = [f"{prefix}.hashgrid.static_grid.hash_table"] + [f"{prefix}.hashgrid.actor_grids.{a}.hash_table" for a in range(meta["n_actors"])]
    width = p[keys[0]].shape[-1] * (8 if field == 0 else 6)
    kw = {}
    if mode == "features":
        kw["dfeatures"] = torch.randn(n * s, width, generator=gen)
    else:
        kw["density"] = torch.rand(n, s, generator=gen) + 0.1
        kw["ddensity"] = torch.randn(n, s, generator=gen)
    out = {}
    for variant in ("fast", "generic"):
        emul.set_bwd_generic(variant == "generic")
        grads = {"static": torch.zeros_like(p[keys[0]]), "actors": [torch.zeros_like(p[k]) for k in keys[1:]]}
        if mode == "density":
            grads["decoder"] = torch.zeros(6)
        emul.encoding_bwd(cfg, p, O.pdf_u, field, mean, std, r["times"][:n], grads, **kw)
        out[variant] = grads
    emul.set_bwd_generic(False)
    a, b = out["fast"], out["generic"]
    assert b["static"].abs().max().item() > 0
    assert rel_to_max(a["static"], b["static"]) < 2e-6
    for x, y in zip(a["actors"], b["actors"]):
        assert (x - y).abs().max().item() <= 2e-6 * max(y.abs().max().item(), 1e-6)
    if mode == "density":
        assert rel_to_max(a["decoder"], b["decoder"]) < 1e-5  # 1536 terms per weight, summed per segment vs per sample


def test_encoding_backward_density_mode_clamps_like_trunc_exp():
    """Pre-activations outside [-15, 15]: the reference's trunc_exp backward is g * exp(clamp(x, -15, 15))
    (field_components/activations.py:38-41), not g * exp(x)."""
    meta, g = load_golden("nff_static.npz")
    cfg = cfg_from_meta(meta)
    ocfg = to_oracle_cfg(cfg)
    p, r, ref = dict(g["param"]), g["ray"], g["ref"]
    key = "proposal_fields.1.density_decoder.weight"
    p[key] = p[key] * 400.0  # drives |x| far beyond 15 on many samples
    n = r["origins"].shape[0]
    edges = ref["bins_e_1"].reshape(n, -1)
    lidar = r["is_lidar"].reshape(-1).bool()
    area = r["pixel_area"].reshape(-1) * torch.where(lidar, 1.0, float(cfg.rgb_upsample_factor**2))
    q, keys = _grad_params(p, "proposal_fields.1", 0)
    dens = O.proposal_density(q, 1, ocfg, r["origins"], r["directions"], area, r["times"].reshape(-1), edges[:, :-1], edges[:, 1:])
    x = dens.detach().log()
    assert (x > 15).any() and (x < -15).any()
    G = torch.randn(dens.shape, generator=torch.Generator().manual_seed(2))
    (dens * G).sum().backward()
    em_mean, em_std = emul.gaussian(r["origins"], r["directions"], area, edges)
    grads = {"static": torch.zeros_like(p[keys[0]]), "actors": [], "decoder": torch.zeros(p[keys[-1]].numel())}
    emul.encoding_bwd(cfg, p, O.pdf_u, 2, em_mean, em_std, r["times"], grads, density=dens.detach(), ddensity=G)
    assert torch.isfinite(grads["static"]).all()
    assert rel_to_max(grads["static"], q[keys[0]].grad) < 1e-4
    assert rel_to_max(grads["decoder"], q[keys[-1]].grad.reshape(-1)) < 1e-4


def test_training_losses_match_reference_golden():
    """distortion_loss_ray / zipnerf_interlevel_ray (csrc/nff_modules.h) against the reference's own loss values and
    autograd gradients (tests/golden/losses.npz, oracle/make_golden_losses.py: loss = 3 * interlevel + 5 * distortion)."""
    meta, g = load_golden("losses.npz")
    sd = [g["in"][f"sdist_{i}"] for i in range(3)]
    w = [g["in"][f"weights_{i}"] for i in range(3)]
    n = sd[0].shape[0]
    dl, ddw = emul.distortion_loss(sd[2], w[2])
    assert abs(dl.mean().item() - g["ref"]["distortion"].item()) < 1e-5 * abs(g["ref"]["distortion"].item())
    assert rel_to_max(ddw * 5 / n, g["ref"]["grad_2"]) < 1e-4  # only the distortion loss reaches the final level's weights
    total = 0.0
    for i in range(2):
        li, dwp = emul.zipnerf_interlevel(sd[2], w[2], sd[i], w[i], meta["pulse_widths"][i])
        total += li.mean().item()
        # relu(w_s - wp)^2 / (wp + 1e-5) amplifies rounding in the blurred cdf: the reference's own fp32 result is 1-2e-4
        # away from a float64 evaluation, and so is this one (cumulative sums in double, like torch's CPU cumsum)
        assert rel_to_max(dwp * 3 / n, g["ref"][f"grad_{i}"]) < 5e-4, i
    assert abs(total - g["ref"]["interlevel"].item()) < 5e-4 * abs(g["ref"]["interlevel"].item())


@pytest.mark.parametrize("with_flip", [False, True])
def test_encoding_backward_to_actor_trajectories(with_flip):
    """dL/d(actor_positions, actor_rotations_6d) of the main field's grid features (require_actor_grad): the device chain
    (position gradient of the lookup -> box transform -> rotation_6d_to_matrix -> keyframe lerp -> keyframe Gram-Schmidt)
    against torch autograd through the oracle's NeuRADHashEncoding restatement."""
    meta, g = load_golden("nff_actors.npz")
    cfg = cfg_from_meta(meta)
    ocfg = to_oracle_cfg(cfg)
    p, r, ref = dict(g["param"]), g["ray"], g["ref"]
    n = r["origins"].shape[0]
    starts, ends = ref["starts"].reshape(n, -1), ref["ends"].reshape(n, -1)
    s = starts.shape[1]
    lidar = r["is_lidar"].reshape(-1).bool()
    area = r["pixel_area"].reshape(-1) * torch.where(lidar, 1.0, float(cfg.rgb_upsample_factor**2))
    for k in ("dynamic_actors.actor_positions", "dynamic_actors.actor_rotations_6d"):
        p[k] = p[k].clone().requires_grad_(True)
    mean, std = O.fast_isotropic_gaussian(r["origins"][:, None, :], r["directions"][:, None, :], area[:, None, None],
                                          starts[..., None], ends[..., None])
    t = r["times"].reshape(n, 1, 1).expand(n, s, 1)
    flip = None
    if with_flip:
        flip = torch.where(torch.arange(n) % 2 == 0, -1.0, 1.0)
    feats, _ = O.hashgrid_forward(p, "field", ocfg.main, ocfg, mean, std, t, None, flip=flip)
    G = torch.randn(feats.shape, generator=torch.Generator().manual_seed(4))
    (feats * G).sum().backward()
    want_pos, want_rot = p["dynamic_actors.actor_positions"].grad, p["dynamic_actors.actor_rotations_6d"].grad
    assert want_pos.abs().max().item() > 0 and want_rot.abs().max().item() > 0
    em_mean, em_std = emul.gaussian(r["origins"], r["directions"], area, torch.cat([starts, ends[:, -1:]], 1))
    g_rot, g_pos = torch.zeros_like(want_rot), torch.zeros_like(want_pos)
    q = {k: v.detach() for k, v in p.items()}
    emul.encoding_pose_bwd(cfg, q, O.pdf_u, 0, em_mean, em_std, r["times"], G, g_rot, g_pos, flip=flip)
    assert rel_to_max(g_pos, want_pos) < 2e-4, rel_to_max(g_pos, want_pos)
    assert rel_to_max(g_rot, want_rot) < 2e-4, rel_to_max(g_rot, want_rot)

"""CPU: the device code of the MODULE-LEVEL operators (csrc/nff_modules.h: NeuRADHashEncoding.forward as a stand-alone
op, the proposal density head) executed by the host emulation, against the reference's own per-stage golden values
(density at the reference's proposal samples, actor ids) and the oracle's NeuRADHashEncoding restatement.  The GPU
tests of the same operators through the C ABI are in tests/test_zz_module_seams_gpu.py."""
import pytest
import torch

from oracle import neurad_oracle as O
from oracle.convert import to_oracle_cfg
from tests.helpers import cfg_from_meta, load_golden
from tests.host_emul import emul


def rel_to_max(a, b):
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)


def scaled_area(cfg, r):
    lidar = r["is_lidar"].reshape(-1).bool()
    return r["pixel_area"].reshape(-1) * torch.where(lidar, 1.0, float(cfg.rgb_upsample_factor**2))


@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz"])
def test_proposal_density_matches_reference_golden(name):
    meta, g = load_golden(name)
    cfg = cfg_from_meta(meta)
    p, r, ref = g["param"], g["ray"], g["ref"]
    for rd, key in ((0, "bins_e_0"), (1, "bins_e_1")):
        mean, std = emul.gaussian(r["origins"], r["directions"], scaled_area(cfg, r), ref[key])
        out = emul.encoding(cfg, p, O.pdf_u, 2, mean, std, r["times"], want_density=True)
        assert rel_to_max(out["density"], ref[f"density_{rd}"].reshape(out["density"].shape)) < 1e-4
        assert torch.equal(out["actor_id"].long(), ref[f"actor_id_{rd}"].reshape(out["actor_id"].shape).long())


@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz", "nff_sharp.npz"])
def test_main_encoding_matches_oracle(name):
    meta, g = load_golden(name)
    cfg = cfg_from_meta(meta)
    ocfg = to_oracle_cfg(cfg)
    p, r, ref = g["param"], g["ray"], g["ref"]
    n = r["origins"].shape[0]
    starts, ends = ref["starts"].reshape(n, -1), ref["ends"].reshape(n, -1)
    edges = torch.cat([starts, ends[:, -1:]], dim=1)
    area = scaled_area(cfg, r)
    mean, std = emul.gaussian(r["origins"], r["directions"], area, edges)
    out = emul.encoding(cfg, p, O.pdf_u, 0, mean, std, r["times"], r["directions"])
    trace = {}
    with torch.no_grad():
        O.main_field(p, ocfg, r["origins"], r["directions"], area, r["times"].reshape(-1), starts, ends, trace)
    s = starts.shape[1]
    assert torch.equal(out["actor_id"].long(), ref["actor_id_main"].reshape(n, s).long())
    assert torch.equal(out["actor_id"].long(), trace["actor_id"].reshape(n, s))
    assert rel_to_max(out["features"].view(n, s, -1), trace["grid_features"]) < 1e-4
    if meta["n_actors"]:
        inside = out["actor_id"] >= 0
        assert inside.any() and (~inside).any()
        # directions of actor samples are in the box frame and unit length; the others are untouched
        d = out["directions"]
        assert torch.equal(d[~inside], r["directions"][:, None, :].expand(n, s, 3)[~inside])
        assert (d[inside].norm(dim=-1) - 1).abs().max() < 1e-5


def test_overlapping_actor_boxes_highest_index_wins():
    """Two actors share a trajectory, so their padded boxes coincide: the reference's `features[ray_idx, sample_idx] = ...`
    (neurad_encoding.py:185) is a sequential index_put on CPU and the LAST triple (highest actor index) wins; the operator
    picks the same actor."""
    meta, g = load_golden("nff_actors.npz")
    cfg = cfg_from_meta(meta)
    ocfg = to_oracle_cfg(cfg)
    p, r, ref = dict(g["param"]), g["ray"], g["ref"]
    n = r["origins"].shape[0]
    base = ref["actor_id_main"].reshape(n, -1)
    a_lo = int(base[base >= 0].min())  # an actor that is hit in this case
    a_hi = a_lo + 1 if a_lo + 1 < meta["n_actors"] else a_lo - 1
    for k in ("actor_positions", "actor_rotations_6d", "actor_present_at_time"):
        t = p[f"dynamic_actors.{k}"].clone()
        t[:, a_hi] = t[:, a_lo]
        p[f"dynamic_actors.{k}"] = t
    sz = p["dynamic_actors.actor_sizes"].clone()
    sz[a_hi] = sz[a_lo]
    p["dynamic_actors.actor_sizes"] = sz
    starts, ends = ref["starts"].reshape(n, -1), ref["ends"].reshape(n, -1)
    area = scaled_area(cfg, r)
    trace = {}
    with torch.no_grad():
        O.main_field(p, ocfg, r["origins"], r["directions"], area, r["times"].reshape(-1), starts, ends, trace)
    mean, std = emul.gaussian(r["origins"], r["directions"], area, torch.cat([starts, ends[:, -1:]], 1))
    out = emul.encoding(cfg, p, O.pdf_u, 0, mean, std, r["times"], r["directions"])
    want = trace["actor_id"].reshape(n, -1)
    winner = max(a_lo, a_hi)
    assert (want == winner).any() and not (want == min(a_lo, a_hi)).any()
    assert torch.equal(out["actor_id"].long(), want)
    assert rel_to_max(out["features"].view(n, starts.shape[1], -1), trace["grid_features"]) < 1e-4

"""CPU: the oracle restatement vs the golden vectors produced by the REAL reference (oracle/make_golden.py).

At generation time the oracle matched the reference bit for bit; here (possibly on another CPU, where MKL may
pick other GEMM kernels) the elementwise stages must still be bit-exact and the GEMM-fed ones within 1e-6.
"""
import pytest
import torch

from oracle import neurad_oracle as O
from oracle.convert import to_oracle_cfg
from tests.helpers import cfg_from_meta, load_golden

CASES = ["nff_static.npz", "nff_actors.npz", "nff_sharp.npz"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    meta, g = load_golden(name)
    cfg = to_oracle_cfg(cfg_from_meta(meta))
    p, r, ref = g["param"], g["ray"], g["ref"]
    with torch.no_grad():
        out = O.nff_outputs(p, cfg, r["origins"], r["directions"], r["pixel_area"], r["times"], r["sensor_idx"],
                            r["is_lidar"], want_trace=True)
        inten, drop = O.decode_lidar(p, out["features"])
    tr = out.pop("trace")
    got = {**out, **tr, "intensity": inten, "ray_drop_logits": drop}
    # integer / index work: bit-exact
    for k in ("inds_1", "inds_2", "actor_id_main", "actor_id_0", "actor_id_1"):
        assert torch.equal(got[k], ref[k]), k
    # initial bins are pure elementwise fp32: bit-exact
    for k in ("bins_s_0", "bins_e_0"):
        assert torch.equal(got[k], ref[k]), k
    for k in ("features", "depth", "accumulation", "prop_depth_0", "prop_depth_1", "prop_weights_0", "prop_weights_1",
              "bins_s_1", "bins_e_1", "bins_s_2", "bins_e_2", "sdf", "alpha", "field_feature", "intensity",
              "ray_drop_logits"):
        torch.testing.assert_close(got[k], ref[k], rtol=1e-6, atol=1e-6, msg=lambda m, k=k: f"{k}: {m}")


def test_hash_encode_matches_reference_golden():
    meta, g = load_golden("nff_static.npz")
    cfg = to_oracle_cfg(cfg_from_meta(meta))
    p, ref = g["param"], g["ref"]
    out = O.hash_encode(ref["hash_in"], p["field.hashgrid.static_grid.hash_table"],
                        p["field.hashgrid.static_grid.scalings"], cfg.main.static.table_size)
    assert torch.equal(out, ref["hash_out"])


def test_density_fn_closure_quirk_is_pinned():
    """neurad.py:248: late-binding closure -> both proposal rounds evaluate proposal_fields[1].  If the oracle used
    proposal_fields[0] for round 0 the golden weights could not match (they are random, distinct tables)."""
    meta, g = load_golden("nff_static.npz")
    p = g["param"]
    assert not torch.equal(p["proposal_fields.0.hashgrid.static_grid.hash_table"],
                           p["proposal_fields.1.hashgrid.static_grid.hash_table"])
    assert O.density_fn_field_index(0, 2) == 1 and O.density_fn_field_index(1, 2) == 1


def test_raygen_matches_reference_golden():
    meta, g = load_golden("raygen.npz")
    for cam in ("cam0", "cam3"):
        c = g[cam]
        h, w = (int(v) for v in c["hw"])
        ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        coords = torch.stack([ys, xs], -1) + 0.5  # cameras.py:296-330 (pixel_offset 0.5)
        fx, fy, cx, cy = (float(v) for v in c["intr"])
        o = O.generate_rays_pinhole(c["c2w"], fx, fy, cx, cy, h, w, coords, float(c["time"]), c["velocity"],
                                    float(c["rs"][0]), float(c["rs"][1]))
        for k in ("origins", "directions", "pixel_area", "times"):
            assert torch.equal(o[k], c[k]), (cam, k)
    li = g["lidar"]
    o = O.generate_rays_lidar_points(li["l2w"], li["points"], float(li["time"]), li["velocity"])
    for k in ("origins", "directions", "pixel_area", "times"):
        assert torch.equal(o[k], li[k]), k


def test_config1_oracle_matches_reference_golden():
    """BASELINE config 1 (64x64 pinhole, UniformSampler(32), 16x2 hash grid, MLP 32->64->4, RGB / expected-depth /
    accumulation renderers): the restatement reproduces the reference's outputs bit for bit."""
    from oracle import simple_oracle as S
    from tests.helpers import load_config1

    meta, p, r, ref = load_config1()
    with torch.no_grad():
        out = S.config1_render(p, r["origins"], r["directions"], r["nears"], r["fars"], 32, want_trace=True)
    for k in ("rgb", "depth", "depth_median", "accumulation", "bins_e"):
        assert torch.equal(out[k].reshape(ref[k].shape), ref[k]), k
    for k in ("positions", "encoding", "raw", "density", "rgb_samples", "weights"):
        assert torch.equal(out[k][::8].reshape(ref[k + "_sub"].shape), ref[k + "_sub"]), k
    assert 0.5 < float(ref["accumulation"].mean()) < 1.0  # the case exercises compositing, not a saturated ray


def test_spacing_functions_match_reference_samplers():
    """SpacedSampler family (ray_samplers.py:135-228, 838-852), pinned by closed forms on a tiny case."""
    from oracle import simple_oracle as S

    nears, fars = torch.tensor([[1.0], [2.0]]), torch.tensor([[4.0], [10.0]])
    _, e = S.spaced_sample(nears, fars, 3, S.SPACING_UNIFORM)
    assert torch.allclose(e, torch.tensor([[1.0, 2.0, 3.0, 4.0], [2.0, 14 / 3, 22 / 3, 10.0]]))
    _, e = S.spaced_sample(nears, fars, 2, S.SPACING_LINDISP)
    assert torch.allclose(e[:, 1], 1 / (0.5 * (1 / nears[:, 0] + 1 / fars[:, 0])))
    _, e = S.spaced_sample(nears, fars, 2, S.SPACING_SQRT)
    assert torch.allclose(e[:, 1], (0.5 * (nears[:, 0].sqrt() + fars[:, 0].sqrt())) ** 2)
    _, e = S.spaced_sample(nears, fars, 2, S.SPACING_LOG)
    assert torch.allclose(e[:, 1], (nears[:, 0] * fars[:, 0]).sqrt())
    for kind in range(5):
        _, e = S.spaced_sample(nears, fars, 8, kind)
        assert torch.allclose(e[:, 0], nears[:, 0], rtol=1e-5) and torch.allclose(e[:, -1], fars[:, 0], rtol=1e-5)
        assert (e[:, 1:] > e[:, :-1]).all()


@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz"])
def test_oracle_autograd_matches_reference_gradients(name):
    """torch autograd through the oracle reproduces the gradients the unmodified reference computed
    (tests/golden/grads_*.npz, oracle/make_golden_grads.py) for every parameter the path trains."""
    from tests.module_seam_cases import OUT_KEYS, check_grads, load_golden_grads, loss_weights

    gmeta, want = load_golden_grads(name)
    meta, g = load_golden(name)
    cfg = cfg_from_meta(meta)
    n = gmeta["n_rays"]
    p = {k: v.clone() for k, v in g["param"].items()}
    pose = ("dynamic_actors.actor_positions", "dynamic_actors.actor_rotations_6d")  # optimize_trajectories (dynamic_actors.py:37)
    keys = [k for k, v in p.items() if v.dtype.is_floating_point and (not k.startswith("dynamic_actors.") or k in pose)
            and not k.endswith("scalings") and k != "static_scale"]
    for k in keys:
        p[k].requires_grad_(True)
    r = g["ray"]
    out = O.nff_outputs(p, to_oracle_cfg(cfg), r["origins"][:n], r["directions"][:n], r["pixel_area"][:n], r["times"][:n],
                        r["sensor_idx"][:n], r["is_lidar"][:n])
    G = loss_weights({k: out[k].shape for k in OUT_KEYS}, gmeta["loss_seed"])
    sum((out[k] * G[k]).sum() for k in OUT_KEYS).backward()
    check_grads({k: p[k].grad for k in keys}, want, min_checked=15)

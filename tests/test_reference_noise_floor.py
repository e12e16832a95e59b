"""CPU: how reproducible is the REFERENCE's own arithmetic on this path?

The parity bar is "1e-4 rel fp32" against vectors the unmodified reference produced on a CPU (tests/golden/nff_*.npz).
Every sampled position is a continuous function of the proposal weights, and those are only defined up to the rounding
of exp() / sum orders.  This test evaluates the reference's math (the oracle, pinned bit for bit to the reference when the
goldens were written) once more with ONE thing changed -- torch.exp replaced by a correctly rounded exp (computed in
double, rounded to fp32) instead of the vectorised SLEEF routine torch uses on CPU -- and records how far the outputs
move.  That distance is a floor no independent implementation (the reference's own CUDA path included) can be expected to
stay under; it is what the tolerances in tests/test_parity_gpu.py are set against:

  * beta = 3 / 4 goldens: depth moves ~1e-5, everything else < 1e-6       -> bar 1e-4 holds with margin
  * beta = 20 golden (the reference's default init): depth moves ~0.7e-4  -> depth bar 2e-4 there, the rest 1e-4
  * per-sample sdf / alpha / features at samples whose edges moved by an ulp: ~1e-4 (finest level = 8191 cells)
"""
import pytest
import torch

from oracle import neurad_oracle as O
from oracle.convert import to_oracle_cfg
from tests.helpers import cfg_from_meta, load_golden


def rel_to_max(a, b):
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)


def _oracle_with_exact_exp(name):
    meta, g = load_golden(name)
    cfg, r = cfg_from_meta(meta), g["ray"]
    orig = torch.exp
    torch.exp = lambda x, *a, **k: orig(x.double()).float() if x.dtype == torch.float32 else orig(x, *a, **k)
    try:
        with torch.no_grad():
            out = O.nff_outputs(g["param"], to_oracle_cfg(cfg), r["origins"], r["directions"], r["pixel_area"], r["times"],
                                r.get("sensor_idx"), r.get("is_lidar"), want_trace=True)
    finally:
        torch.exp = orig
    out.update(out.pop("trace"))
    return meta, out, g["ref"]


@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz", "nff_sharp.npz"])
def test_reference_math_with_a_different_exp_stays_inside_the_parity_bars(name):
    meta, out, ref = _oracle_with_exact_exp(name)
    moved = {k: rel_to_max(out[k], ref[k]) for k in ("features", "accumulation", "depth", "prop_depth_0", "prop_depth_1")}
    print(name, "beta", meta["beta"], {k: f"{v:.1e}" for k, v in moved.items()})
    for k, v in moved.items():
        assert v < (2e-4 if (k == "depth" and meta["beta"] >= 20) else 1e-4), (k, v)
    # sample indices / actor assignment survive the perturbation on these fixtures (they are asserted exactly on the GPU)
    for k in ("inds_1", "inds_2", "actor_id_main"):
        assert int((out[k].long() != ref[k].long()).sum()) == 0, k


def test_beta20_depth_floor_is_close_to_the_bar():
    """Documents WHY depth at beta = 20 is held to 2e-4 and not 1e-4: the reference's math itself moves by more than
    0.3e-4 (measured 0.7e-4) when exp() is rounded differently; per-sample sdf moves by ~1e-4 where an edge moved an ulp."""
    meta, out, ref = _oracle_with_exact_exp("nff_sharp.npz")
    assert meta["beta"] >= 20
    assert rel_to_max(out["depth"], ref["depth"]) > 3e-5
    eq = out["bins_s_2"] == ref["bins_s_2"]
    same = eq[:, :-1] & eq[:, 1:]
    err = (out["sdf"].reshape(same.shape) - ref["sdf"].reshape(same.shape)).abs() / ref["sdf"].abs().max()
    assert err[same].max().item() < 1e-5  # identical sample position -> fp32-exact
    assert err[~same].max().item() > 1e-5  # moved by an ulp -> amplified by the fine grid levels

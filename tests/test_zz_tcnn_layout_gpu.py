"""GPU: tiny-cuda-nn layout through the C ABI against the CPU restatement of the library's published
algorithm (oracle/tcnn_oracle.py; PARITY UNPINNED -- tiny-cuda-nn itself is not installed anywhere in this environment)."""
import pytest

from tests import tcnn_cases as C

pytestmark = pytest.mark.gpu


def test_grid_matches_oracle_and_interpolates_linear_functions():
    C.grid_matches_oracle_and_interpolates_linear_functions("cuda")


@pytest.mark.parametrize("n_actors", [0, 3])
def test_fused_render_matches_tcnn_oracle(n_actors):
    C.fused_render_matches_tcnn_oracle("cuda", n_actors=n_actors, n_rays=256)


def test_torch_layout_still_selected_after_a_tcnn_bind():
    """Binding a torch-layout parameter set after a tcnn one switches the kernels back (per-context layout flag)."""
    import torch

    import neurad_studio_b200 as nsb
    from neurad_studio_b200 import scene
    from neurad_studio_b200.nerfstudio_api import get_backend
    from oracle import neurad_oracle as O
    from oracle.convert import to_oracle_cfg

    C.fused_render_matches_tcnn_oracle("cuda", n_actors=0, n_rays=64)
    be = get_backend(torch.device("cuda", 0))
    cfg = nsb.small_config(n_actors=0, log2_main=12, log2_prop=11)
    params = scene.make_params(cfg, seed=4, beta=3.0, sdf_bias=0.5)
    rays = scene.random_rays(64, cfg, seed=6)
    be.load_params(cfg, params)
    assert be.layout == "torch"
    out = be.render(rays)
    with torch.no_grad():
        ref = O.nff_outputs(params, to_oracle_cfg(cfg), rays["origins"], rays["directions"], rays["pixel_area"], rays["times"],
                            rays["sensor_idx"], rays["is_lidar"])
    assert C.rel_to_max(out["features"], ref["features"]) < 1e-4

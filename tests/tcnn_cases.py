"""Shared bodies of the tiny-cuda-nn-layout tests.  PARITY UNPINNED: tiny-cuda-nn is absent, so the
reference for these tests is oracle/tcnn_oracle.py, an independent CPU restatement of the published algorithm, plus
properties that hold whatever the layout details are.  `dev` = "cuda" (through the C ABI) or "cpu" (host emulation of the
same device functions)."""
import math

import torch

import neurad_studio_b200 as nsb
from neurad_studio_b200 import scene
from neurad_studio_b200 import tcnn_compat as T
from oracle import neurad_oracle as O
from oracle import tcnn_oracle as TO
from oracle.convert import to_oracle_cfg


def rel_to_max(a, b):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    return (a.reshape(b.shape) - b).abs().max().item() / (b.abs().max().item() + 1e-30)


def tcnn_test_config(n_actors=3):
    """NeuRAD's grid SHAPES (8x4 / 6x1 / 4 levels) at resolutions small enough that every kind of level occurs in every
    grid: dense coarse levels, hashed fine levels, and a dense level of the 4-D actor grid."""
    cfg = nsb.NeuRADConfig(n_actors=n_actors)
    cfg.grid.static = nsb.HashGridSettings(4, 8, 8, 1024, 15)         # 8^3 .. 25^3 dense, finer levels hashed into 2^15
    cfg.grid.actor = nsb.HashGridSettings(4, 4, 6, 96, 12)             # 4-D: 6^4 = 1296 dense, the rest hashed into 2^12
    for g in cfg.proposal_grids:
        g.static = nsb.HashGridSettings(1, 6, 16, 512, 14)
        g.actor = nsb.HashGridSettings(1, 4, 6, 96, 11)
    return cfg


def _grid_fn(dev):
    """(layout, flat params, x) -> features, through the C ABI or the host emulation."""
    if dev == "cuda":
        from neurad_studio_b200.nerfstudio_api import get_backend

        be = get_backend(torch.device("cuda", 0))
        return lambda lay, p, x, sc: be.tcnn_hashgrid_fwd(lay, p.cuda(), x.cuda(), sc).cpu()
    from tests.host_emul import emul

    return lambda lay, p, x, sc: emul.tcnn_hashgrid(lay, p, x)


def layout_has_every_kind_of_level():
    cfg = tcnn_test_config()
    for g, nd in ((cfg.grid.static, 3), (cfg.grid.actor, 4), (cfg.proposal_grid_1.static, 3), (cfg.proposal_grid_1.actor, 4)):
        lay = T.layout_of(g, nd)
        assert any(lay["dense"]) and not all(lay["dense"]), (lay["resolution"], lay["size"], lay["dense"])
        for l in range(lay["n_levels"]):
            r, n = lay["resolution"][l], lay["size"][l]
            assert n % 8 == 0 and n <= 2**g.log2_hashmap_size
            assert lay["dense"][l] == (r**nd <= n)
            if not lay["dense"][l]:
                assert n == 2**g.log2_hashmap_size  # hashed levels fill their share exactly (a power of two: `& mask`)
            assert lay["offset"][l] == sum(lay["size"][:l])
        # product-side layout == the oracle's independent restatement
        ol = TO.grid_layout(g.num_levels, g.hashgrid_dim, g.log2_hashmap_size, g.base_res, T.growth_factor(g), nd)
        assert ol.resolution == lay["resolution"] and ol.size == lay["size"] and ol.offset == lay["offset"] and ol.dense == lay["dense"]
        assert all(abs(a - b) < 1e-6 * max(1, abs(b)) for a, b in zip(ol.scale, lay["scale"]))
    # NeuRAD's real sizes: the three coarsest levels of the main grid are dense (32^3, 71^3, 157^3 <= 2^22)
    lay = T.layout_of(nsb.NeuRADConfig().grid.static, 3)
    assert lay["resolution"][:3] == [32, 71, 157] and lay["dense"] == [True, True, True] + [False] * 5


def grid_matches_oracle_and_interpolates_linear_functions(dev):
    grid = _grid_fn(dev)
    gen = torch.Generator().manual_seed(7)
    for g, nd in ((nsb.HashGridSettings(4, 8, 8, 1024, 15), 3), (nsb.HashGridSettings(1, 6, 16, 512, 14), 3),
                  (nsb.HashGridSettings(4, 4, 6, 96, 12), 4), (nsb.HashGridSettings(2, 3, 5, 20, 10), 4)):
        lay = T.layout_of(g, nd)
        n = lay["n_entries"] * lay["n_features"]
        params = T.half_round(torch.rand(n, generator=gen) * 2 - 1)
        x = torch.rand(777, nd, generator=gen)
        x[0], x[1] = 0.0, 1.0  # the corners of the unit cube (pos = 0.5 and scale + 0.5)
        ol = TO.grid_layout(g.num_levels, g.hashgrid_dim, g.log2_hashmap_size, g.base_res, T.growth_factor(g), nd)
        ref = TO.hashgrid_encode(ol, params, x)
        out = grid(lay, params, x, g.scalings())
        assert out.shape == ref.shape and rel_to_max(out, ref) < 2e-6, (nd, rel_to_max(out, ref))
        # layout-independent property: fill every DENSE level from an affine function of the vertex position; N-linear
        # interpolation must reproduce the function exactly (up to rounding) at arbitrary points
        coef = torch.tensor([0.3, -0.7, 0.45, 0.2][:nd])
        p2 = torch.zeros(lay["n_entries"], lay["n_features"])
        for l in range(lay["n_levels"]):
            if not lay["dense"][l]:
                continue
            r = lay["resolution"][l]
            axes = torch.meshgrid(*[torch.arange(r, dtype=torch.float32)] * nd, indexing="ij")
            # linear index = x + y*r + z*r^2 (+ w*r^3): x is the FASTEST axis
            vert = torch.stack([a.permute(*reversed(range(nd))).reshape(-1) for a in axes], -1)
            val = ((vert - 0.5) / lay["scale"][l]) @ coef  # vertex v sits at x = (v - 0.5) / scale
            p2[lay["offset"][l]: lay["offset"][l] + r**nd] = val[:, None] + torch.arange(lay["n_features"])[None, :]
        xi = torch.rand(300, nd, generator=gen) * 0.75 + 0.05  # away from x ~ 1, where the "+1" vertex wraps (no clamping)
        out2 = grid(lay, p2.reshape(-1), xi, g.scalings()).reshape(300, lay["n_levels"], lay["n_features"])
        want = (xi @ coef)[:, None] + torch.arange(lay["n_features"])[None, :]
        for l in range(lay["n_levels"]):
            if lay["dense"][l]:
                assert (out2[:, l] - want).abs().max().item() < 2e-5, (nd, l)


def mlp_unpack_strips_the_padding():
    gen = torch.Generator().manual_seed(3)
    for in_dim, width, n_layers, out_dim in ((32, 32, 2, 33), (48, 32, 3, 32), (48, 32, 3, 2), (20, 16, 2, 5)):
        shapes = T.mlp_shapes(in_dim, width, n_layers - 1, out_dim)
        assert shapes[0] == (width, (in_dim + 15) // 16 * 16) and shapes[-1] == ((out_dim + 15) // 16 * 16, width)
        flat = torch.randn(sum(o * k for o, k in shapes), generator=gen)
        ws = T.mlp_unpack(flat, in_dim, width, n_layers - 1, out_dim)
        ows = TO.mlp_unpack(flat, in_dim, width, n_layers - 1, out_dim)
        assert [tuple(w.shape) for w in ws] == [(width, in_dim)] + [(width, width)] * (n_layers - 2) + [(out_dim, width)]
        assert all(torch.equal(a, b) for a, b in zip(ws, ows))
        assert all(torch.equal(w, w.half().float()) for w in ws)  # fp16-representable
        # the padded evaluation (zero-padded input, all padded output rows) restricted to the real rows is the same function
        x = torch.randn(9, in_dim, generator=gen)
        h, off = torch.nn.functional.pad(x, (0, shapes[0][1] - in_dim)), 0
        for i, (o, k) in enumerate(shapes):
            w = T.half_round(flat[off:off + o * k]).reshape(o, k)
            off += o * k
            h = h @ w.t()
            h = torch.relu(h) if i < len(shapes) - 1 else h
        assert torch.allclose(h[:, :out_dim], TO.mlp_forward(ws, x), atol=1e-5)


def fused_render_matches_tcnn_oracle(dev, n_actors=3, n_rays=96, beta=1.5):
    """The whole NFF path on a tcnn-layout parameter set: CUDA (C ABI) or host emulation vs the oracle's tcnn mode."""
    cfg = tcnn_test_config(n_actors)
    trajs = scene.make_trajectories(n_actors, cfg.duration, seed=2) if n_actors else None
    params = scene.make_params_tcnn(cfg, seed=3, beta=beta, trajectories=trajs, mlp_gain=0.6)
    rays = scene.random_rays(n_rays, cfg, seed=5, trajectories=trajs)
    with torch.no_grad():
        ref = O.nff_outputs(params, to_oracle_cfg(cfg), rays["origins"], rays["directions"], rays["pixel_area"], rays["times"],
                            rays["sensor_idx"], rays["is_lidar"], want_trace=True)
        ref_i, ref_d = O.decode_lidar(params, ref["features"])
    tr = ref.pop("trace")
    if n_actors:
        assert int((tr["actor_id_main"] >= 0).sum()) > 0  # the 4-D actor grid is exercised
    if dev == "cuda":
        from neurad_studio_b200.nerfstudio_api import get_backend

        be = get_backend(torch.device("cuda", 0))
        be.load_params(cfg, params)
        assert be.layout == "tcnn"
        out = be.render(rays, want_trace=True, want_intensity=True)
        be.check_status()
        assert rel_to_max(out["intensity"], ref_i) < 1e-4 and rel_to_max(out["ray_drop_logits"], ref_d) < 1e-4
    else:
        from tests.host_emul import emul

        out = emul.render(cfg, params, rays, O.pdf_u, lane_mode=True)
    for k in ("actor_id_0", "actor_id_1", "actor_id_main"):
        assert int((out[k].cpu().long() != tr[k].long()).sum()) == 0, k
    for k in ("inds_1", "inds_2"):
        assert (out[k].cpu().long() != tr[k].long()).float().mean().item() <= 2e-3, k
    for k in ("features", "accumulation", "depth", "prop_depth_0", "prop_depth_1"):
        assert rel_to_max(out[k], ref[k]) < 1e-4, (k, rel_to_max(out[k], ref[k]))
    return out, ref

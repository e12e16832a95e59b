"""TEST INFRASTRUCTURE ONLY -- golden vectors of BASELINE config 1 from the REAL reference (needs /root/reference).

    python -m oracle.make_golden_config1      ->  tests/golden/config1.npz

Builds the reference's own modules (Cameras, UniformSampler, HashEncoding(torch), MLP(torch), trunc_exp, RaySamples
.get_weights, RGBRenderer("black"), DepthRenderer("expected"/"median"), AccumulationRenderer), runs them on the CPU and
asserts that oracle/simple_oracle.py reproduces every stage BIT FOR BIT before writing the fixture.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from oracle import simple_oracle as S  # noqa: E402
from oracle.make_golden import _save  # noqa: E402


def main(log2_hashmap_size: int = 19, near: float = 0.05, far: float = 4.0, table_scale: float = 1.0):
    ref_import.install()
    from nerfstudio.cameras.cameras import Cameras, CameraType
    from nerfstudio.data.scene_box import SceneBox
    from nerfstudio.field_components.activations import trunc_exp
    from nerfstudio.field_components.encodings import HashEncoding
    from nerfstudio.field_components.mlp import MLP
    from nerfstudio.model_components.ray_samplers import UniformSampler
    from nerfstudio.model_components.renderers import AccumulationRenderer, DepthRenderer, RGBRenderer

    torch.manual_seed(0)
    c2w = torch.eye(4)[:3]
    cam = Cameras(camera_to_worlds=c2w[None], fx=64.0, fy=64.0, cx=32.0, cy=32.0, width=64, height=64,
                  camera_type=CameraType.PERSPECTIVE)
    rb = cam.generate_rays(camera_indices=0, keep_shape=False)
    n = rb.origins.shape[0]
    rb.nears = torch.full((n, 1), near)
    rb.fars = torch.full((n, 1), far)
    enc = HashEncoding(num_levels=16, min_res=16, max_res=1024, log2_hashmap_size=log2_hashmap_size,
                       features_per_level=2, implementation="torch")
    with torch.no_grad():
        enc.hash_table.copy_(S.synthetic_table(enc.hash_table.shape[0], enc.hash_table.shape[1], table_scale))
    mlp = MLP(in_dim=32, num_layers=2, layer_width=64, out_dim=4, implementation="torch")
    aabb = torch.tensor([[-4.0, -4.0, -4.0], [4.0, 4.0, 4.0]])
    sampler = UniformSampler(num_samples=32).eval()
    rgb_r, dep_r, med_r, acc_r = RGBRenderer("black").eval(), DepthRenderer("expected"), DepthRenderer("median"), AccumulationRenderer()
    with torch.no_grad():
        rs = sampler(rb)
        pos = SceneBox.get_normalized_positions(rs.frustums.get_positions(), aabb)
        feat = enc(pos.view(-1, 3))
        raw = mlp(feat).view(n, 32, 4)
        density = trunc_exp(raw[..., 0:1])
        rgb_s = torch.sigmoid(raw[..., 1:4])
        w = rs.get_weights(density)
        ref = dict(rgb=rgb_r(rgb_s, w), depth=dep_r(w, rs), depth_median=med_r(w, rs), accumulation=acc_r(w),
                   positions=pos, encoding=feat.view(n, 32, -1), raw=raw, density=density, rgb_samples=rgb_s, weights=w,
                   bins_e=torch.cat([rs.frustums.starts[..., 0], rs.frustums.ends[:, -1:, 0]], -1))
    lin = [m for m in mlp.layers if isinstance(m, torch.nn.Linear)]
    p = dict(hash_table=enc.hash_table.detach(), scalings=enc.scalings.detach(), w0=lin[0].weight.detach(),
             b0=lin[0].bias.detach(), w1=lin[1].weight.detach(), b1=lin[1].bias.detach(), aabb=aabb)
    with torch.no_grad():
        out = S.config1_render(p, rb.origins, rb.directions, rb.nears, rb.fars, 32, want_trace=True)
    for k, v in ref.items():
        assert torch.equal(v.reshape(out[k].shape), out[k]), f"oracle != reference for {k}"
    print(f"config1: oracle == reference bit-for-bit on {len(ref)} tensors; accumulation mean "
          f"{ref['accumulation'].mean():.3f}, depth range [{ref['depth'].min():.3f}, {ref['depth'].max():.3f}]")
    # the table (16 x 2^19 x 2 fp32 = 64 MB) is not committed: S.synthetic_table re-creates it bit-identically
    arrays = {f"param/{k}": v for k, v in p.items() if k != "hash_table"}
    arrays.update({"ray/origins": rb.origins, "ray/directions": rb.directions, "ray/nears": rb.nears, "ray/fars": rb.fars})
    per_ray = ("rgb", "depth", "depth_median", "accumulation", "bins_e")
    arrays.update({f"ref/{k}": ref[k] for k in per_ray})
    arrays.update({f"ref/{k}_sub": v[::8] for k, v in ref.items() if k not in per_ray})  # per-sample stages: every 8th ray
    arrays["param/hash_table_sub"] = p["hash_table"][::4099]
    meta = dict(log2_hashmap_size=log2_hashmap_size, near=near, far=far, table_scale=table_scale, torch=torch.__version__)
    return arrays, meta, p


if __name__ == "__main__":
    arrays, meta, _ = main()
    _save("config1.npz", arrays, meta)

"""TEST INFRASTRUCTURE ONLY -- pin the TRAINING-MODE sampling restatements to the real reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_stratified

Runs the reference's own ``PowerSampler`` and ``PDFSampler`` modules in ``.train()`` mode (train_stratified, both
single_jitter settings) on seeded inputs.  The modules draw their jitter with ``torch.rand`` internally; re-seeding the
global generator before the call and drawing a tensor of the same shape afterwards recovers exactly those numbers, which
are then fed to the oracle restatements (oracle/neurad_oracle.py::pdf_resample(rand=), oracle/simple_oracle.py::
spaced_sample(t_rand=)).  The script asserts bit-equality and writes tests/golden/stratified.npz (inputs, the jitter,
reference outputs) for the GPU tests of b200nerf_spaced_sample_stratified / b200nerf_pdf_resample_stratified.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import neurad_oracle as O  # noqa: E402
from oracle import ref_import  # noqa: E402
from oracle import simple_oracle as S  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    ref_import.install()
    from nerfstudio.cameras.rays import RayBundle
    from nerfstudio.model_components.ray_samplers import PDFSampler, PowerSampler

    gen = torch.Generator().manual_seed(3)
    n, s0, s1 = 64, 128, 64
    origins = torch.randn(n, 3, generator=gen)
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1)
    nears = torch.rand(n, 1, generator=gen) * 0.5
    fars = 50 + torch.rand(n, 1, generator=gen) * 500
    rb = RayBundle(origins=origins, directions=dirs, pixel_area=torch.full((n, 1), 1e-6), nears=nears, fars=fars)
    arrays = {"in/nears": nears, "in/fars": fars}
    for tag, single in (("single", True), ("full", False)):
        init = PowerSampler(lambda_=-1.0, scaling=0.1, single_jitter=single).train()
        torch.manual_seed(100 + int(single))
        rs = init(rb, num_samples=s0)
        torch.manual_seed(100 + int(single))
        t_rand = torch.rand((n, 1 if single else s0 + 1))
        bins_s = torch.cat([rs.spacing_starts[..., 0], rs.spacing_ends[..., -1:, 0]], -1)
        bins_e = torch.cat([rs.frustums.starts[..., 0], rs.frustums.ends[..., -1:, 0]], -1)
        ob_s, ob_e = S.spaced_sample(nears, fars, s0, S.SPACING_POWER, -1.0, 0.1, t_rand=t_rand)
        assert torch.equal(ob_s.expand_as(bins_s), bins_s) and torch.equal(ob_e, bins_e), tag
        weights = torch.rand(n, s0, 1, generator=gen) ** 4
        weights[3] = 0.0  # a ray without any weight (the eps padding branch)
        pdf = PDFSampler(include_original=False, single_jitter=single).train()
        torch.manual_seed(200 + int(single))
        rs2 = pdf(rb, rs, weights, num_samples=s1)
        torch.manual_seed(200 + int(single))
        rand = torch.rand((n, 1 if single else s1 + 1))
        new_s = torch.cat([rs2.spacing_starts[..., 0], rs2.spacing_ends[..., -1:, 0]], -1)
        new_e = torch.cat([rs2.frustums.starts[..., 0], rs2.frustums.ends[..., -1:, 0]], -1)
        r = O.pdf_resample(weights[..., 0], bins_s, s1, 0.01, rand=rand)
        assert torch.equal(r["bins"], new_s), tag
        arrays.update({f"{tag}/t_rand": t_rand, f"{tag}/bins_s": bins_s, f"{tag}/bins_e": bins_e, f"{tag}/weights": weights[..., 0],
                       f"{tag}/rand": rand, f"{tag}/new_bins_s": new_s, f"{tag}/new_bins_e": new_e, f"{tag}/inds": r["inds"]})
        print(f"{tag}: oracle == reference bit-for-bit (stratified PowerSampler + PDFSampler)")
    out = {k: v.numpy() for k, v in arrays.items()}
    out["__meta__"] = np.array(repr(dict(n=n, s0=s0, s1=s1, power_lambda=-1.0, power_scaling=0.1, histogram_padding=0.01,
                                         torch=torch.__version__)))
    path = os.path.join(GOLDEN, "stratified.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)/1e6:.2f} MB")


if __name__ == "__main__":
    main()

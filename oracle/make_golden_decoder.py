"""TEST INFRASTRUCTURE ONLY -- golden vectors of NeuRADModel.rgb_decoder from the REAL reference modules.

    python -m oracle.make_golden_decoder      ->  tests/golden/rgb_decoder.npz

Builds the reference's nn.Sequential exactly as NeuRADModel.populate_modules does (models/neurad.py:201-216, with
model_components/cnns.py BasicBlock), loads seeded parameters, runs it in eval mode on a small feature image and
asserts that oracle/decoder_oracle.py reproduces it BIT FOR BIT before writing the fixture.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import decoder_oracle as D  # noqa: E402
from oracle import ref_import  # noqa: E402
from oracle.make_golden import _save  # noqa: E402


def reference_decoder(in_dim=48, hidden_dim=32, upsample=3):
    ref_import.install()
    from nerfstudio.model_components.cnns import BasicBlock

    return torch.nn.Sequential(  # neurad.py:201-216
        torch.nn.Conv2d(in_dim, hidden_dim, kernel_size=1, padding=0),
        torch.nn.ReLU(inplace=True),
        BasicBlock(hidden_dim, hidden_dim, kernel_size=7, padding=3, use_bn=True),
        BasicBlock(hidden_dim, hidden_dim, kernel_size=7, padding=3, use_bn=True),
        torch.nn.ConvTranspose2d(hidden_dim, hidden_dim, kernel_size=upsample, stride=upsample),
        BasicBlock(hidden_dim, hidden_dim, kernel_size=7, padding=3, use_bn=True),
        BasicBlock(hidden_dim, hidden_dim, kernel_size=7, padding=3, use_bn=True),
        torch.nn.Conv2d(hidden_dim, 3, kernel_size=1, padding=0),
        torch.nn.Sigmoid(),
    )


def main():
    torch.manual_seed(0)
    dec = reference_decoder().eval()
    p = D.random_decoder_params(seed=11)
    sd = {k[len("rgb_decoder."):]: v for k, v in p.items()}
    missing, unexpected = dec.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
    g = torch.Generator().manual_seed(12)
    # two "cameras" of 19 x 45 feature pixels (odd sizes: strips narrower than a tile, ragged rows)
    feats = torch.randn(2, 19, 45, 48, generator=g) * 0.7
    with torch.no_grad():
        ref = dec(feats.permute(0, 3, 1, 2)).permute(0, 2, 3, 1).contiguous()  # neurad.py:362-365
        out = D.rgb_decoder(p, feats)
    assert torch.equal(ref, out), "oracle != reference for rgb"
    print(f"rgb_decoder: oracle == reference bit-for-bit; rgb in [{ref.min():.3f}, {ref.max():.3f}], std {ref.std():.3f}")
    arrays = {f"param/{k}": v for k, v in p.items()}
    arrays.update({"in/features": feats, "ref/rgb": ref})
    _save("rgb_decoder.npz", arrays, dict(seed=11, in_dim=48, hidden=32, upsample=3, torch=torch.__version__))


if __name__ == "__main__":
    main()

"""Device-side timing of the rgb decoder on config-2 sized input (6 cameras x 360 x 640 feature pixels -> 6 x 1080 x 1920
rgb).  Development aid, not the bench."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neurad_studio_b200.backend import B200Backend
from neurad_studio_b200 import scene

B, H, W = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (6, 360, 640)))
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
be = B200Backend(torch.device("cuda", 0))
be.set_rgb_decoder(scene.make_rgb_decoder_params(seed=1))
feats = torch.randn(B, H, W, 48, device="cuda") * 0.7
for impl in os.environ.get("DEC_IMPLS", "tc,tc_ldgsts,ref").split(","):
    out = be.rgb_decode(feats, impl)
    torch.cuda.synchronize()
    be.check_status()
    ts = []
    for _ in range(reps if impl != "ref" else 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = be.rgb_decode(feats, impl); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    flop = 2.0 * B * H * W * (48 * 32 + 4 * 50176 + 32 * 288 + 9 * (4 * 50176 + 96))
    print(f"{impl}: {ms:.3f} ms  ({flop / ms / 1e9:.1f} algorithmic TFLOP/s, {B * H * W / ms / 1e3:.2f} M camera rays/s)  all={['%.2f' % t for t in ts]}")
    if impl == "tc":
        keep = out
    else:
        print("max |tc - ref| =", (keep - out).abs().max().item())

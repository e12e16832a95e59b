"""Per-image CUDA-event timeline of bench.py's API arm (development aid for the decoder-stream question).

For every step and image it records, relative to the step's start: when the render finished on the main stream, when the rgb
decoder started / finished (main stream, or the side stream with --decoder-stream), when the image's D2H copies finished on
the copy stream, and the host time the Python thread needed to ENQUEUE the image -- so a slow step shows whether a stream
waited, a copy dragged, or the host fell behind.  Prints the median step and the slowest step side by side.

  python tools/e2e_trace.py [--steps 20] [--warmup 3] [--decoder-stream]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--decoder-stream", action="store_true")
    a = ap.parse_args()
    numa = bench.bind_to_gpu_numa_node(0)
    import neurad_studio_b200 as nsb
    from neurad_studio_b200 import scene
    from neurad_studio_b200.nerfstudio_api import NeuRADModel

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = nsb.NeuRADConfig(n_actors=0)
    params = scene.make_params(cfg, seed=1, beta=3.0, sdf_bias=0.6, device=dev)
    model = NeuRADModel(cfg)
    model.load_reference_state_dict(params)
    model.rgb_decoder.load_state_dict({k[len("rgb_decoder."):]: v for k, v in scene.make_rgb_decoder_params(seed=2).items()}, strict=False)
    model = model.to(dev).eval()
    del params
    cams, scan = bench.build_workload(cfg, 0)
    step = bench.Step(model, cfg, cams, scan, 1, 0)
    ds = torch.cuda.Stream(device=dev) if a.decoder_stream else None
    model.set_decoder_stream(ds)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def one_step():
        rec, keep = [], []
        t0 = ev()
        t0.record()
        for i, host in enumerate(step.host_cam):
            h0 = time.perf_counter()
            rb = step.cameras.generate_rays(camera_indices=i, keep_shape=True)
            out = model.get_outputs_for_camera_ray_bundle(rb)
            main_done = ev()
            main_done.record()  # single stream: render + decoder; with a decoder stream: the render (+ lidar head) only
            dec_done = out.get("rgb_ready")
            step._to_host([(host[k], out[k]) for k in host], keep, ready=dec_done)
            copy_done = ev()
            copy_done.record(step.copy_stream)
            rec.append((main_done, dec_done, copy_done, (time.perf_counter() - h0) * 1e3))
        out, _ = model.get_outputs_for_lidar(step.lidars, {"lidar": step.points_pinned, "lidar_idx": 0})
        step._to_host([(step.host_lidar[k], out[k]) for k in step.host_lidar], keep)
        end = ev()
        end.record(step.copy_stream)
        if ds is not None:
            ds.synchronize()
        step.copy_stream.synchronize()
        torch.cuda.current_stream(dev).synchronize()
        rows = []
        for main_done, dec_done, copy_done, host_ms in rec:
            rows.append((t0.elapsed_time(main_done), None if dec_done is None else _elapsed(t0, dec_done), t0.elapsed_time(copy_done), host_ms))
        return t0.elapsed_time(end), rows

    def _elapsed(t0, e):
        # rgb_ready is created without timing in the API; re-recording is not possible, so measure through a timed twin
        return float("nan") if not getattr(e, "_timed", False) else t0.elapsed_time(e)

    for _ in range(a.warmup):
        one_step()
    runs = [one_step() for _ in range(a.steps)]
    totals = sorted(t for t, _ in runs)
    med = totals[len(totals) // 2]
    slow_t, slow_rows = max(runs, key=lambda r: r[0])
    med_rows = min(runs, key=lambda r: abs(r[0] - med))[1]
    print(f"numa {numa}; decoder_stream {a.decoder_stream}; steps {a.steps}: median {med:.2f} ms, min {totals[0]:.2f}, max {slow_t:.2f}")
    print("image | main-stream done (ms) | D2H done (ms) | host enqueue (ms)    [median step | slowest step]")
    for i, (m, s) in enumerate(zip(med_rows, slow_rows)):
        print(f"  {i}   | {m[0]:8.2f} {s[0]:8.2f} | {m[2]:8.2f} {s[2]:8.2f} | {m[3]:6.2f} {s[3]:6.2f}")


if __name__ == "__main__":
    main()

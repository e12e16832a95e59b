#!/usr/bin/env python
"""bench.py -- rays/sec of NeuRAD's volumetric-rendering hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one synthetic PandaSet-shaped time step (BASELINE config 2): ray
generation for 6 x 1920x1080 pinhole cameras at NeuRAD's render stride ([1::3,1::3] -> 6 x 230 400 rays) and for
one 64-beam x 1800-azimuth lidar sweep (115 200 rays), then `get_nff_outputs` for all 1 497 600 rays with the
reference's default grids / MLPs (random-init, tables U(-1,1), 0 actors).  Definition of the metric as in the
reference: rays / time between device synchronisations (nerfstudio/pipelines/ad_pipeline.py:198-208, 296-304).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 is launched by torchrun, one rank per GPU.  Rays shard with no data-path collective (every rank renders
its own time step = weak scaling); the only communication is the per-step NCCL all-gather of the per-ray
outputs, into which the kernel's epilogue writes directly.

Beside the headline keys the N = 1 line carries two secondary figures: "with_rgb_decoder" (the step followed by the
camera rgb decoder) and "train_step" (one NFF training step through the hand-written backward operators,
SURVEY 8f f2, measured by tools/train_probe.py in a child process once every headline measurement is done).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

ALGO_BYTES_PER_RAY = 69_900  # SURVEY.md section 8(d) / BASELINE.md section 2: fp32 tables, no actor hits
CAM_RAYS = 640 * 360
WORKLOAD = "neurad-default config2: 6x1920x1080 pinhole @stride3 (6x230400 rays) + 64x1800 lidar (115200 rays), 0 actors"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- workload
def build_workload(cfg, frame: int):
    """Host-side description of one time step: 6 cameras + 1 lidar sweep at t = 1 + 0.5*frame seconds."""
    from neurad_studio_b200 import scene

    t = 1.0 + 0.5 * (frame % 12)
    return scene.pandaset_rig(time=t), scene.pandar64_scan(time=t, seed=frame)


class Step:
    """The public-API call sequence of one step (what a user of the backend does to render one time step)."""

    def __init__(self, be, cfg, cams, scan, world, rank, gather="p2p"):
        self.be, self.cfg, self.cams, self.scan = be, cfg, cams, scan
        self.n_cam = len(cams) * CAM_RAYS
        self.n = self.n_cam + scan.points.shape[0]
        dev = be.device
        self.world, self.rank = world, rank
        fdim = cfg.feature_dim
        # the gather buffers [world, n, w]: every rank's slice is written directly by the render kernel; with
        # gather="p2p" they live in symmetric memory and the kernel also stores each row into the peers' copies
        self.p2p = world > 1 and gather == "p2p"
        if self.p2p:
            from neurad_studio_b200.dist import PeerGatherBuffers

            self.pg = PeerGatherBuffers(self.n, fdim, dev)
            self.pg.bind(be)
            self.gather = self.pg.buf
        else:
            self.gather = {k: torch.empty(world, self.n, w, device=dev) for k, w in (("features", fdim), ("depth", 1), ("accumulation", 1))}
        self.local = {k: torch.empty(self.n, 1, device=dev) for k in ("prop_depth_0", "prop_depth_1")}
        self.sensor = torch.cat([torch.full((CAM_RAYS,), c.sensor_idx, dtype=torch.long) for c in cams] +
                                [torch.full((scan.points.shape[0],), scan.sensor_idx, dtype=torch.long)]).to(dev)
        self.is_lidar = torch.cat([torch.zeros(self.n_cam, dtype=torch.uint8), torch.ones(scan.points.shape[0], dtype=torch.uint8)]).to(dev)
        self.rays = {k: torch.empty(self.n, w, device=dev) for k, w in (("origins", 3), ("directions", 3), ("pixel_area", 1), ("times", 1))}
        self.points_dev = scan.points.to(dev)
        self.points_pinned = scan.points.clone().pin_memory()
        self.host_out = {k: torch.empty(self.n, w).pin_memory() for k, w in (("features", fdim), ("depth", 1), ("accumulation", 1))}
        self.kernel_events = []
        self.launches = 0
        # peer lists: device peers (fused multi-GPU gather) and, for the e2e arm, the pinned host buffers as one more
        # "peer" (pinned host memory is device-mapped under UVA)
        keys = ("features", "depth", "accumulation")
        widths = {"features": fdim, "depth": 1, "accumulation": 1}
        host_ptrs = {k: self.host_out[k].data_ptr() for k in keys}
        if self.p2p:
            self._dev_peers = {k: [int(p) for p in self.pg.hdl[k].buffer_ptrs] for k in keys}
            # GPU peers hold [world, n, w] (this rank's rows at row_offset = rank * n); the host buffer is [n, w], so
            # its base pointer is shifted back by that offset
            self._e2e_row_offset = self.rank * self.n
            self._e2e_peers = {k: self._dev_peers[k] + [host_ptrs[k] - self._e2e_row_offset * widths[k] * 4] for k in keys}
            self._e2e_self = self.rank
        else:
            self._dev_peers = None
            self._e2e_row_offset = 0
            self._e2e_peers, self._e2e_self = {k: [host_ptrs[k]] for k in keys}, -1

    def _raygen(self, points):
        be = self.be
        off = 0
        for cam in self.cams:
            be.raygen_pinhole(cam, 1, 3, 1, 3, out={k: v[off:off + CAM_RAYS] for k, v in self.rays.items()})
            off += CAM_RAYS
            self.launches += 1
        be.raygen_lidar_points(self.scan, points, out={k: v[off:] for k, v in self.rays.items()})
        self.launches += 1

    def run_device(self, time_kernel: bool):
        """inputs already resident in HBM"""
        self._raygen(self.points_dev)
        rays = dict(self.rays, sensor_idx=self.sensor, is_lidar=self.is_lidar)
        out = {k: self.gather[k][self.rank] for k in self.gather}
        out.update(self.local)
        if time_kernel:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self.be.render(rays, out=out, image_width=640)
        self.launches += 2  # nff_sample_lane_kernel + nff_shade_lane_kernel
        if time_kernel:
            e1.record()
            self.kernel_events.append((e0, e1))
        self._finish_gather()
        return out

    def _finish_gather(self):
        if self.world == 1:
            return
        if self.p2p:
            self.pg.barrier()  # rows were stored into every peer by the render kernel; only a barrier is left
        else:
            import torch.distributed as dist

            for k, buf in self.gather.items():
                dist.all_gather_into_tensor(buf.view(-1), buf[self.rank].reshape(-1))

    def run_e2e(self):
        """host buffers in, host buffers out, inside the call: H2D of the step's inputs (camera descriptors + the lidar
        sweep from pinned memory) and the step's results landing in pinned host memory.  The results are not copied
        after the fact: the pinned output buffers are handed to the library as one more "peer" (set_peer_outputs), so the
        render epilogue streams every finished row to the host over PCIe while the kernel is still rendering."""
        dev = self.be.device
        pts = self.points_pinned.to(dev, non_blocking=True)
        self._raygen(pts)
        rays = dict(self.rays, sensor_idx=self.sensor, is_lidar=self.is_lidar)
        out = {k: self.gather[k][self.rank] for k in self.gather}
        out.update(self.local)
        self.be.set_peer_outputs(self._e2e_peers, self_rank=self._e2e_self, row_offset=self._e2e_row_offset)
        self.be.render(rays, out=out, image_width=640)
        self.launches += 2  # nff_sample_lane_kernel + nff_shade_lane_kernel
        self._finish_gather()
        torch.cuda.current_stream(dev).synchronize()
        if self.p2p:
            self.be.set_peer_outputs(self._dev_peers, self_rank=self.rank, row_offset=self.rank * self.n)
        else:
            self.be.set_peer_outputs(None)
        return self.host_out

    @property
    def h2d_bytes(self):
        cam_desc = len(self.cams) * (12 + 4 + 3 + 4) * 4
        return self.points_pinned.numel() * 4 + cam_desc

    @property
    def d2h_bytes(self):
        return sum(t.numel() * 4 for t in self.host_out.values())


_BEST_THREADS = None


def pick_cpu_threads(cfg):
    """The torch CPU path is made of small ops and stops scaling (or regresses) on many-core hosts: probe a few
    thread counts on a tiny sample and keep the fastest, so the baseline is the reference at its best."""
    global _BEST_THREADS
    if _BEST_THREADS is not None:
        return _BEST_THREADS
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (ncpu, 64, 32, 16, 8) if c <= ncpu})
    best, best_v = cands[0], -1.0
    for c in cands:
        torch.set_num_threads(c)
        v, _, _ = oracle_rays_per_sec(cfg, 1024, _threads_fixed=True)
        if v > best_v:
            best, best_v = c, v
    _BEST_THREADS = best
    torch.set_num_threads(best)
    return best


def oracle_rays_per_sec(cfg, n_sample: int, repeats: int = 1, _threads_fixed: bool = False):
    """The reference's PyTorch path (oracle port) on the host cores, on a bounded sample of the same workload."""
    from neurad_studio_b200 import scene
    from oracle import neurad_oracle as O
    from oracle.convert import to_oracle_cfg

    if not _threads_fixed:
        pick_cpu_threads(cfg)
    params = scene.make_params(cfg, seed=1, beta=3.0, sdf_bias=0.6)
    cams, scan = build_workload(cfg, 0)
    n_l = n_sample // 13  # same camera : lidar proportion as the workload (12 : 1)
    n_c = n_sample - n_l
    cam = cams[0]
    ys, xs = torch.meshgrid(torch.arange(1, cam.height, 3), torch.arange(1, cam.width, 3), indexing="ij")
    coords = (torch.stack([ys, xs], -1).reshape(-1, 2)[:: max(1, CAM_RAYS // n_c)][:n_c] + 0.5).float()
    ocfg = to_oracle_cfg(cfg)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        with torch.no_grad():
            rc = O.generate_rays_pinhole(cam.c2w, cam.fx, cam.fy, cam.cx, cam.cy, cam.height, cam.width, coords, cam.time,
                                         cam.velocity, cam.rolling_shutter_time, cam.time_to_center_pixel)
            rl = O.generate_rays_lidar_points(scan.l2w, scan.points[:n_l], scan.time, scan.velocity)
            rays = {k: torch.cat([rc[k], rl[k]]) for k in ("origins", "directions", "pixel_area", "times")}
            n = rays["origins"].shape[0]
            sensor = torch.cat([torch.zeros(n_c, 1, dtype=torch.long), torch.full((n_l, 1), 6)])
            is_lidar = torch.cat([torch.zeros(n_c, 1, dtype=torch.bool), torch.ones(n_l, 1, dtype=torch.bool)])
            O.nff_outputs(params, ocfg, rays["origins"], rays["directions"], rays["pixel_area"], rays["times"], sensor, is_lidar)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n / best, n, best


_SAVED_STDOUT = None


def train_step_probe(timeout_s: float = 240.0) -> dict:
    """Secondary figure for SURVEY 8(f) row f2: one NFF TRAINING step (NeuRAD's 40 960 camera + 16 384 lidar ray batch
    through the module walk, both regularisers, loss.backward() through the hand-written backward operators), timed
    with CUDA events by tools/train_probe.py in a CHILD process after every headline measurement is finished -- a
    failure, crash or time-out of that young code path can only turn this entry into {"error": ...}, never the line."""
    import subprocess

    cmd = [sys.executable, os.path.join(ROOT, "tools", "train_probe.py"), "--steps", "5", "--warmup", "2"]
    try:
        res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout_s)
        lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        if res.returncode != 0 or not lines:
            tail = (res.stderr or res.stdout).strip().splitlines()[-3:]
            return {"error": f"tools/train_probe.py rc={res.returncode}: " + " | ".join(tail)[:400]}
        out = json.loads(lines[-1])
        out["note"] = "secondary figure (not the headline metric); measured in a child process after the timed arms"
        return out
    except Exception as e:  # time-out, missing file, malformed output ...
        return {"error": f"{type(e).__name__}: {e}"[:400]}


def _emit(line: dict):
    sys.stdout.flush()
    if _SAVED_STDOUT is not None:
        os.dup2(_SAVED_STDOUT, 1)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=16384)
    ap.add_argument("--no-decoder", action="store_true", help="skip the extra 'with_rgb_decoder' measurement (N = 1)")
    ap.add_argument("--no-train", action="store_true", help="skip the extra 'train_step' measurement (N = 1, child process)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: p2p = render epilogue stores rows into every peer's buffer over NVLink (default); "
                         "nccl = all_gather_into_tensor after the render")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3)

    import neurad_studio_b200 as nsb

    cfg = nsb.NeuRADConfig(n_actors=0)
    base = {
        "metric": "rays/sec (camera+lidar)", "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": 6 * CAM_RAYS + 115200, "tables": "fp32, main 8x2^22x4 + proposal 6x2^20x1 (U(-1,1))",
                   "l2": "inputs larger than L2 (560 MB of tables, 300 MB of outputs per step); no explicit flush", "kernel": "ray-per-lane, 2-D tile walk (image_width=640), tcgen05 3xTF32 MLPs", "parallelism": f"ray-shard dp{world}" + ("" if world == 1 else f", gather={args.gather}")},
    }

    if args.impl == "reference":
        # the reference's own (PyTorch, CPU) implementation of the path = the oracle port, on the host cores
        if rank != 0:
            return
        n_sample = max(2048, args.cpu_sample // 4)
        for _ in range(min(args.warmup, 1)):
            oracle_rays_per_sec(cfg, n_sample)
        t0 = time.perf_counter()
        tot = 0
        for _ in range(args.steps):
            _, n, _ = oracle_rays_per_sec(cfg, n_sample)
            tot += n
        dt = time.perf_counter() - t0
        v = tot / dt
        line = dict(base, impl="reference", value=v, ms_per_step=dt / args.steps * 1e3, n_gpus=world,
                    cpu_baseline={"value": v, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                                  "sample": f"{n_sample} rays/step of the same workload (12:1 camera:lidar), oracle port of the reference torch path"},
                    e2e={"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, gpu_launches=0)
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU baseline")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        # NCCL / c10d print a version banner on stdout at first use; rank 0 must print ONE JSON line, so stdout is
        # pointed at stderr until the result line is written (restored in _emit)
        sys.stdout.flush()
        global _SAVED_STDOUT
        _SAVED_STDOUT = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    from neurad_studio_b200 import scene
    from neurad_studio_b200.backend import B200Backend

    be = B200Backend(dev)
    params = scene.make_params(cfg, seed=1, beta=3.0, sdf_bias=0.6, device=dev)
    be.load_params(cfg, params)
    cams, scan = build_workload(cfg, rank)
    step = Step(be, cfg, cams, scan, world, rank, args.gather)

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        step.run_device(False)
        step.run_e2e()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    step.launches = 0
    ms = timed(lambda: step.run_device(True), args.steps)
    launches = step.launches // args.steps
    kern_ms = sorted(a.elapsed_time(b) for a, b in step.kernel_events)
    kern_ms = sum(kern_ms) / len(kern_ms)
    ms_e2e = timed(step.run_e2e, args.steps)
    # the host buffers the e2e arm filled must hold exactly what the device buffers hold
    e2e_ok = all(torch.equal(step.host_out[k], step.gather[k][rank].cpu()) for k in step.host_out)
    # (ii) of SURVEY 8(d): the same step followed by the camera rgb decoder (NeuRADModel.rgb_decoder, tcgen05 implicit-GEMM
    # convolutions) on the six rendered feature images -> 6 x 1080 x 1920 rgb.  Reported beside the headline, N = 1 only
    # (the decoder needs whole images; the multi-GPU arm shards rays, not images).
    dec_line = None
    if world == 1 and not args.no_decoder:
        be.set_rgb_decoder(scene.make_rgb_decoder_params(seed=2, device=dev))
        n_cams = len(cams)
        rgb = torch.empty(n_cams, 1080, 1920, 3, device=dev)
        dec_events = []

        def run_with_decoder():
            out = step.run_device(False)
            feats = out["features"][: step.n_cam].view(n_cams, 360, 640, -1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            be.rgb_decode(feats, out=rgb)
            e1.record()
            dec_events.append((e0, e1))

        for _ in range(2):
            run_with_decoder()
        dec_events.clear()
        ms_dec = timed(run_with_decoder, args.steps)
        be.check_status()
        d_ms = sorted(a.elapsed_time(b) for a, b in dec_events)
        d_ms = sum(d_ms) / len(d_ms)
        mac_per_ray = 48 * 32 + 4 * 50176 + 32 * 288 + 9 * (4 * 50176 + 3 * 32)
        tf = 2.0 * mac_per_ray * step.n_cam / (d_ms * 1e-3) / 1e12
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        tpeak = float(peaks.get("bf16_tflops_sustained", 1443.2))
        dec_line = {"value": step.n * args.steps / (ms_dec * 1e-3), "unit": "rays/s", "ms_per_step": ms_dec / args.steps,
                    "decoder_ms": d_ms, "decoder_camera_rays_per_s": step.n_cam / (d_ms * 1e-3), "gpu_launches_decoder": 10,
                    "roofline": {"bound": "tensor", "achieved": tf, "executed": 3 * tf, "peak": tpeak, "unit": "TFLOP/s",
                                 "frac": tf / tpeak, "frac_executed": 3 * tf / tpeak,
                                 "note": "achieved = algorithmic 4.04 MFLOP/camera ray; executed = 3x (bf16 hi/lo split: three MMAs per product for fp32-level accuracy); peak = measured sustained dense bf16"},
                    "what": "render step + NeuRADModel.rgb_decoder on the 6 feature images (6x360x640x48 -> 6x1080x1920x3 rgb)"}
    clocks = sampler.stop() if rank == 0 else None
    rays_total = step.n * world * args.steps
    value = rays_total / (ms * 1e-3)
    e2e_value = rays_total / (ms_e2e * 1e-3)
    if rank != 0:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    achieved = step.n * ALGO_BYTES_PER_RAY / (kern_ms * 1e-3) / 1e9
    line = dict(base, value=value, ms_per_step=ms / args.steps, clocks=clocks, gpu_launches=launches,
                e2e={"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": step.h2d_bytes, "d2h_bytes_per_step": step.d2h_bytes,
                     "ms_per_step": ms_e2e / args.steps, "host_buffers_verified": bool(e2e_ok),
                     "how": "host descriptors -> ray generation -> fused render; the render epilogue stores every finished row into pinned host memory (device-mapped) while rendering"},
                roofline={"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                          "traffic": None, "kernel": "nff_sample_lane_kernel + nff_shade_lane_kernel (one render)", "kernel_ms": kern_ms, "peak_source": peak_src,
                          "algorithmic_bytes_per_ray": ALGO_BYTES_PER_RAY})
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        line["roofline"]["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_launch")
    if dec_line is not None:
        line["with_rgb_decoder"] = dec_line
    if world == 1 and not args.no_train:
        torch.cuda.synchronize()
        torch.cuda.empty_cache()  # the child process needs ~3 GB of its own
        line["train_step"] = train_step_probe()
    if world == 1 and args.cpu_sample > 0:
        v, n, dt = oracle_rays_per_sec(cfg, args.cpu_sample)
        line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{n} rays (12:1 camera:lidar) of the same workload in {dt:.1f} s, oracle port of the reference torch path, best of thread counts probed, host has {os.cpu_count()} cpus"}
    _emit(line)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()

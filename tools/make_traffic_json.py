"""profiles/traffic.json from an `ncu --set full` report of the two render kernels (one launch each).

  ncu -i gpurun_out/<rep>.ncu-rep --page raw --csv > raw.csv
  python tools/make_traffic_json.py raw.csv "profiles/<summary file>" > profiles/traffic.json

Records the DRAM traffic per render (read + write of both kernels), the limiter figures bench.py copies into
`roofline.limiter`, and the hash of the kernel sources the capture belongs to: bench.py reports the traffic only while the
sources are unchanged (a capture of another binary is not this run's traffic)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    raw, source = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, key, scale_unit=True):
        v = float(r[idx[key]].replace(",", ""))
        u = units[idx[key]]
        if scale_unit:
            v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        return v

    per, tot_r, tot_w = {}, 0.0, 0.0
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        name = "nff_sample_lane_kernel" if "sample" in name else "nff_shade_lane_kernel" if "shade" in name else name
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        per[name] = {
            "read": rd, "write": wr, "ncu_ms": val(r, "gpu__time_duration.sum", False) * {"ms": 1, "us": 1e-3, "ns": 1e-6}.get(units[idx["gpu__time_duration.sum"]], 1),
            "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
            "l1_lsu_wavefronts_pct": val(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", False),
            "dram_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
            "tensor_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
            "l1_hit_pct": val(r, "l1tex__t_sector_hit_rate.pct", False), "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct", False),
            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active", False),
            "registers": int(val(r, "launch__registers_per_thread", False)),
        }
        tot_r += rd
        tot_w += wr
    from bench import kernel_sources_sha

    out = {
        "kernel": "nff_sample_lane_kernel + nff_shade_lane_kernel (one render = the pair)", "source": source,
        "kernel_sources_sha": kernel_sources_sha(), "rays_per_launch": 1497600,
        "dram_bytes_per_launch": tot_r + tot_w, "dram_bytes_read": tot_r, "dram_bytes_write": tot_w, "per_kernel": per,
        "limiter": {
            "what": "physical limiter of the pair (ncu --set full): instruction issue, not HBM",
            **{f"{k.split('_')[1]}_{m}": v[m] for k, v in per.items() for m in ("issue_active_pct", "l1_lsu_wavefronts_pct", "dram_pct", "tensor_pipe_pct")},
        },
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

"""Summarise an ncu report exported as CSV pages (raw + source) into a short text block for profiles/."""
import csv, collections, re, sys

def summarise(raw_csv, src_csv, n_rays, title):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    keys = ["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","lts__t_sector_hit_rate.pct","l1tex__t_sector_hit_rate.pct","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__grid_size","launch__block_size","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct_of_peak_sustained_elapsed","lts__throughput.avg.pct_of_peak_sustained_elapsed","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","lts__t_bytes.sum","l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum","l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
    out = [title, ""]
    for h,u,v in zip(hdr,units,vals):
        if h in keys: out.append(f"{h:75s} {u:16s} {v}")
    rows = list(csv.reader(open(src_csv)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)  # header row (position varies with ncu options)
    hdr = rows[hi]; rows = [None, hdr] + rows[hi + 1:]; idx = {h:i for i,h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter(); ops = collections.Counter(); n_inst=0; n_sass=0
    for r in rows[2:]:
        if r == hdr: break  # the page repeats itself (second view): one section is the kernel
        if len(r) < len(hdr): continue
        ie = int(r[idx["Instructions Executed"]] or 0); n_inst += ie; n_sass += 1
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[idx["Source"]]); op = m.group(2).split(".")[0] if m else "?"
        ops[op] += ie
        for c in stall_cols: tot[c] += int(r[idx[c]] or 0)
    out += ["", f"SASS instructions in kernel: {n_sass}; warp instructions executed: {n_inst} ({n_inst/n_rays:.0f} per ray)", "", "warp stall samples:"]
    s = sum(tot.values())
    for k,v in tot.most_common(9): out.append(f"  {k:26s} {100*v/s:5.1f}%")
    out += ["", "executed instruction mix:"]
    for k,v in ops.most_common(16): out.append(f"  {k:8s} {100*v/n_inst:5.1f}%  ({v/n_rays:7.0f} per ray)")
    return "\n".join(out) + "\n"

if __name__ == "__main__":
    print(summarise(sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4] if len(sys.argv) > 4 else ""))

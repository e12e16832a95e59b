"""Attribute an ncu source-page CSV (SASS view) to source FUNCTIONS of nff_device.h / tc_mlp.cuh using the
inline line info of the cubin (nvdisasm -gi).  Usage:
  python tools/ncu_by_function.py <lib.so> <kernel-name-substring> <ncu_source.csv> <n_rays>
"""
import collections, csv, os, re, subprocess, sys, tempfile

def function_ranges(path):
    """[(start_line, end_line, name)] of function bodies in a C++ header (brace matching on NFF_D / template lines)."""
    src = open(path).read().split("\n")
    out, i = [], 0
    pat = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:NFF_D|NFF_HD|__device__|__host__|inline|static)[^;{]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;]*$")
    while i < len(src):
        m = pat.match(src[i])
        if m and "(" in src[i]:
            name = m.group(1); depth = 0; j = i; seen = False
            while j < len(src):
                depth += src[j].count("{") - src[j].count("}")
                if "{" in src[j]: seen = True
                if seen and depth <= 0: break
                j += 1
            out.append((i + 1, j + 1, name)); i = j + 1
        else:
            i += 1
    return out

def main():
    so, kname, ncu_csv, n_rays = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    asm = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout.split("\n")
    csrc = os.path.join(os.path.dirname(os.path.abspath(so)), "..", "csrc")
    ranges = {}
    for f in ("nff_device.h", "nff_lane.h", "tc_mlp.cuh"):
        ranges[f] = function_ranges(os.path.join(csrc, f))
    def fn_of(file, line):
        for a, b, n in ranges.get(os.path.basename(file), []):
            if a <= line <= b: return n
        return None
    # parse kernel
    in_k = False; frames = []; last = []; off2chain = {}
    for ln in asm:
        if ln.startswith("\t.section") or ln.startswith(".text."):
            pass
        if re.match(r"^\.text\..*", ln):
            in_k = kname in ln
            continue
        if not in_k: continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            frames.append((m.group(1), int(m.group(2)))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            # nvdisasm prints the location only when it changes: instructions without one inherit the previous
            if frames:
                last = list(frames)
            off2chain[int(m.group(1), 16)] = last; frames = []
    rows = list(csv.reader(open(ncu_csv)))
    hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
    base = None
    agg_inst = collections.Counter(); agg_samp = collections.Counter(); agg_stall = collections.defaultdict(collections.Counter)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_i = tot_s = 0
    for r in rows[2:]:
        if len(r) < len(hdr): continue
        addr = int(r[idx["Address"]], 16)
        if base is None: base = addr
        chain = off2chain.get(addr - base, [])
        names = [fn_of(f, l) for f, l in chain]
        names = [n for n in names if n]
        # innermost meaningful function, skipping tiny wrappers
        skip = {"fmul", "fadd", "fsub", "fdiv", "frcp", "fsqrt", "ldg", "lane", "shfl", "shfl_up", "blend", "smem_u32", "tf32_hi"}
        inner = next((n for n in names if n not in skip), "kernel-body")
        phase = "proposal" if ("proposal_round" in names or "lane_proposal_round" in names) else ("mlp" if ("run" in names or "layer" in names) else "main/other")
        key = f"{phase:10s} {inner}"
        ie = int(r[idx["Instructions Executed"]] or 0); sm = int(r[idx["# Samples"]] or 0)
        agg_inst[key] += ie; agg_samp[key] += sm; tot_i += ie; tot_s += sm
        for c in stall_cols: agg_stall[key][c] += int(r[idx[c]] or 0)
    print(f"{'phase / function':42s} {'inst/ray':>9s} {'inst%':>6s} {'samples%':>9s}  top stalls")
    for k, v in sorted(agg_samp.items(), key=lambda kv: -kv[1])[:28]:
        st = agg_stall[k]; s = sum(st.values()) or 1
        top = ", ".join(f"{n.replace('stall_','')} {100*c/s:.0f}%" for n, c in st.most_common(3))
        print(f"{k:42s} {agg_inst[k]/n_rays:9.0f} {100*agg_inst[k]/tot_i:5.1f}% {100*v/tot_s:8.1f}%  {top}")

if __name__ == "__main__":
    main()

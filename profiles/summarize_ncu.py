#!/usr/bin/env python
"""
Turns an ``ncu --set full --import-source on`` capture of k_run into the text summary committed under profiles/.

Usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep [path/to/libitsim_b200.so] > profiles/<name>.txt
Needs ncu (to read the report) and, for the per-function table, cuobjdump + nvdisasm of the *same* build.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

RAW_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    # instruction supply: per-SM instruction cache (icc) and the GPC-level cache behind it (gcc)
    "sm__icc_requests.sum", "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum",
    "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "gcc__average_cache_request_hit_rate.pct",
    "smsp__inst_executed_op_branch.sum",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    lib = sys.argv[2] if len(sys.argv) > 2 else None
    raw = ncu_csv(rep, "raw")
    head, units, vals = raw[0], raw[1], raw[2]
    kernel = vals[head.index('Kernel Name')].split("(")[0]
    print(f"# ncu summary of {os.path.basename(rep)} (kernel {kernel})")
    print("\n## raw metrics")
    for name in RAW_METRICS:
        if name in head:
            i = head.index(name)
            print(f"{name:70s} {vals[i]:>18s} {units[i]}")
    src = ncu_csv(rep, "source")
    hdr, data = src[1], src[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {s: sum(int(r[idx[s]] or 0) for r in data) for s in stalls}
    total = sum(tot.values())
    print(f"\n## warp stall samples (all samples; {total} total; {len(data)} SASS instructions in k_run + its functions)")
    for s, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v:
            print(f"{s:28s} {v:10d} {100.0 * v / total:5.1f}%")
    executed = [int(r[idx["Instructions Executed"]]) for r in data]
    order = sorted(range(len(executed)), key=lambda i: -executed[i])
    print(f"\n## hot code footprint ({sum(executed)} warp-instructions executed)")
    for frac in (0.5, 0.8, 0.9, 0.95, 0.99):
        acc = n = 0
        for i in order:
            acc += executed[i]
            n += 1
            if acc >= frac * sum(executed):
                break
        print(f"{100 * frac:5.1f}% of executed instructions come from {n:5d} SASS instructions ({n * 16 / 1024:6.1f} KB)")
    if lib:
        with tempfile.TemporaryDirectory() as tmp:
            subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
            cubins = sorted(f for f in os.listdir(tmp) if f.endswith(".cubin"))
            texts = [subprocess.run(["nvdisasm", "-c", os.path.join(tmp, c)], capture_output=True, text=True).stdout
                     for c in cubins]
        seq = []
        for text in texts:        # the cubin (translation unit) that holds the profiled kernel
            fn, cand = None, []
            for line in text.splitlines():
                m = re.match(r"^([_a-zA-Z$.][A-Za-z0-9_$.]*):", line)
                if m and not m.group(1).startswith(".L"):
                    fn = m.group(1)
                elif re.match(r"^ +/\*[0-9a-f]+\*/", line):
                    cand.append(fn)
            cand = [f for f in cand if f == ".text." + kernel or f.startswith("$" + kernel + "$")
                    or f.startswith("$__internal")]
            if len(cand) != len(data):      # an $__internal helper of another kernel of the cubin (a division routine)
                for drop in sorted({f for f in cand if f.startswith("$__internal")}):
                    trial = [f for f in cand if f != drop]
                    if len(trial) == len(data):
                        cand = trial
                        break
            if len(cand) == len(data):
                seq = cand
        if len(seq) == len(data):
            ex, sm, ni, size = (collections.Counter() for _ in range(4))
            for f, r in zip(seq, data):
                m = re.search(r"4itsb\d+([a-z_0-9]+?)E", f)
                name = m.group(1) if m else f
                ex[name] += int(r[idx["Instructions Executed"]])
                sm[name] += int(r[idx["# Samples"]])
                ni[name] += int(r[idx["stall_no_inst"]])
                size[name] += 1
            t = sum(sm.values())
            print("\n## per function (k_run = the inlined hot loop; the rest are out-of-line device functions)")
            print(f"{'function':24s} {'SASS':>6s} {'executed (M)':>13s} {'samples':>8s} {'no_inst share':>14s}")
            for k, v in sm.most_common(16):
                print(f"{k:24s} {size[k]:6d} {ex[k] / 1e6:13.1f} {100.0 * v / t:7.1f}% {100.0 * ni[k] / max(1, v):13.0f}%")
        else:
            print(f"\n(per-function table skipped: {len(seq)} disassembled vs {len(data)} profiled instructions)")


if __name__ == "__main__":
    main()

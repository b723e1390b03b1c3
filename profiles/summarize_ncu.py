#!/usr/bin/env python
"""Condense an Nsight Compute report into the few numbers the roofline discussion uses.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.md

Per profiled launch: duration, DRAM bytes (read/write), issue-slot utilisation, IPC, registers,
occupancy, stall-reason shares from warp sampling, and the ten hottest SASS instructions.
"""
import collections
import csv
import io
import subprocess
import sys


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(path):
    raw = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum"]
    print(f"# ncu summary of `{path}`\n")
    for r in raw[2:]:
        print(f"## {r[hdr.index('Kernel Name')]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in want:
            if k in hdr:
                print(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        print()
    src = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "source", "--csv", "--print-source", "sass"]))))
    kernels, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    seen = set()
    for kd in kernels:
        if not kd["rows"]:
            continue
        fingerprint = (kd["name"], tuple(tuple(r[:4]) for r in kd["rows"][:40]))       # (ncu lists a launch's source page more than once)
        if fingerprint in seen:
            continue
        seen.add(fingerprint)
        h = kd["rows"][0]
        data = [r for r in kd["rows"][1:] if len(r) >= len(h) - 2]
        i_s, i_src = h.index("# Samples"), h.index("Source")
        stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        tot = sum(int(r[i_s]) for r in data) or 1
        agg = collections.Counter()
        for r in data:
            for c in stall:
                if r[c]:
                    agg[h[c][6:]] += int(r[c])
        print(f"### warp-state sampling: {kd['name'][:90]}\n")
        print(f"{len(data)} SASS instructions, {tot} samples; stall shares: " +
              ", ".join(f"{k} {v / tot:.1%}" for k, v in agg.most_common(9)) + "\n")
        print("| # | samples | instruction | top stalls |\n|---|---|---|---|")
        for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][i_s]))[:10]):
            r = data[i]
            st = {h[c][6:]: int(r[c]) for c in stall if r[c] not in ("", "0")}
            top = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
            print(f"| {i} | {r[i_s]} | `{r[i_src].strip()[:70]}` | {top} |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])

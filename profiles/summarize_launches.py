#!/usr/bin/env python
"""Kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python profiles/summarize_launches.py profiles/<name>.csv "<command that was profiled>" > profiles/<name>_summary.md
"""
import collections
import csv
import re
import sys


def main(path, command):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki])
        tot[name] += float(r[vi].replace(",", "")) / 1e3
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# ncu launch list of `{command}` ({sum(cnt.values())} launches captured)\n")
    print("Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's `kernels[*].share`, not absolutes.\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for name, t in tot.most_common():
        print(f"| `{name}` | {cnt[name]} | {t:.1f} | {t / cnt[name]:.1f} | {t / total:.3f} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "?")

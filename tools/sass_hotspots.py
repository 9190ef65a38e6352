#!/usr/bin/env python3
"""Per-opcode and per-instruction hot spots of one kernel in an .ncu-rep (source page, SASS view).
    python tools/sass_hotspots.py <report.ncu-rep> <kernel-regex> [top_n]
"""
import csv
import subprocess
import sys
from collections import Counter


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                          f"regex:{pat}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # first kernel instance only
    end = next((i for i in range(2, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
    hdr, data = rows[1], [r for r in rows[2:end] if len(r) == len(rows[1])]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
    print(f"kernel {rows[0][1][:80]}\nsamples {tot}, static instructions {len(data)}, warp instructions executed {inst}")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = sorted(((s, sum(int(r[ix[s]]) for r in data)) for s in stalls), key=lambda x: -x[1])[:7]
    print("stall reasons:", ", ".join(f"{s[6:]} {100 * n / max(tot, 1):.1f}%" for s, n in agg))
    c, cs = Counter(), Counter()
    for r in data:
        parts = r[ix["Source"]].split()
        op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
        c[op] += int(r[ix["Instructions Executed"]])
        cs[op] += int(r[ix["# Samples"]])
    print("opcode            executed   %inst  %samples")
    for op, n in c.most_common(16):
        print(f"{op:<12} {n:>13} {100 * n / inst:>7.1f} {100 * cs[op] / max(tot, 1):>8.1f}")
    print("hottest instructions (samples, executed, sass, top stall)")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top_n]:
        st = max(stalls, key=lambda s: int(r[ix[s]]))
        print(f"{r[ix['# Samples']]:>7} {r[ix['Instructions Executed']]:>10}  {r[ix['Source']][:64]:<64} {st[6:]}")


if __name__ == "__main__":
    main()

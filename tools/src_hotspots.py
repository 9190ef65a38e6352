#!/usr/bin/env python3
"""Per-SOURCE-LINE hot spots of one kernel in an .ncu-rep (needs -lineinfo and --import-source on):
    python tools/src_hotspots.py <report.ncu-rep> <kernel-regex> [top_n [inst]]
Prints, per source line: stall samples, share, warp instructions executed, top stall reasons."""
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    by_inst = len(sys.argv) > 4 and sys.argv[4] == "inst"     # rank the lines by warp instructions instead of stall samples
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name",
                          f"regex:{pat}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ix = {h: i for i, h in enumerate(hdr)}
            continue
        if r[0] in ("Function Name", "Kernel Name") or hdr is None or len(r) != len(hdr):
            continue
        if r[0] == "":      # a SASS row under the current source line
            continue
        stalls = {h[6:]: int(r[i] or 0) for h, i in ix.items() if h.startswith("stall_") and "Not Issued" not in h}
        lines.append((cur_file, int(r[0]), r[1].strip(), int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0), stalls))
    tot = sum(l[3] for l in lines) or 1
    inst = sum(l[4] for l in lines) or 1
    print(f"samples {tot}, warp instructions {inst}")
    print(f"{'file:line':<28}{'samples':>8}{'%':>6}{'inst%':>7}  top stalls / source")
    for f, ln, src, smp, ins, st in sorted(lines, key=lambda l: -(l[4] if by_inst else l[3]))[:top_n]:
        top = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2] if v)
        print(f"{f + ':' + str(ln):<28}{smp:>8}{100 * smp / tot:>6.1f}{100 * ins / inst:>7.1f}  [{top}] {src[:90]}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` export: instruction mix, stall reasons, hottest SASS lines."""
import collections
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    ops, samp, stall = collections.Counter(), collections.Counter(), collections.Counter()
    lines = []
    tot_i = tot_s = 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[0] == "Address":
            continue
        src = r[ix["Source"]].strip()
        tok = src.split()
        op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
        try:
            n = int(r[ix["Instructions Executed"]])
            s = int(r[ix["# Samples"]])
        except ValueError:
            continue
        ops[op] += n
        samp[op] += s
        tot_i += n
        tot_s += s
        for c in stall_cols:
            stall[c] += int(r[ix[c]] or 0)
        lines.append((s, n, r[ix["Avg. Threads Executed"]], src))
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    for op, n in ops.most_common(top):
        print("%-10s inst %10d %5.1f%%   samples %5.1f%%" % (op, n, 100.0 * n / tot_i, 100.0 * samp[op] / max(1, tot_s)))
    print("stalls:", ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100.0 * v / max(1, tot_s))
                               for k, v in sorted(stall.items(), key=lambda x: -x[1])[:8]))
    print("hottest instructions (samples, executed, avg threads):")
    for s, n, t, src in sorted(lines, reverse=True)[:top]:
        print("  %6d %9d %5s  %s" % (s, n, t, src))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)

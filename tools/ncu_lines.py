#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source cuda,sass` output per source line:
share of executed instructions, stall samples and L2 sectors (first kernel instance only)."""
import csv
import sys


def num(x):
    try:
        return int(float(x))
    except Exception:
        return 0


def main(path, thresh=0.008):
    rows = list(csv.reader(open(path)))
    sections, cur, fp = [], None, None
    for r in rows:
        if r and r[0] == "File Path":
            fp = r[1]
        if r and r[0] == "Line No" and len(r) > 5:
            cur = {"file": fp, "hdr": r, "rows": []}
            sections.append(cur)
            continue
        if cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    seen, agg = set(), {}
    for sec in sections:
        if sec["file"] in seen:
            break                      # second kernel instance starts
        seen.add(sec["file"])
        h = sec["hdr"]
        iS, iI = h.index("# Samples"), h.index("Instructions Executed")
        iL, iG = h.index("L2 Theoretical Sectors Local"), h.index("L2 Theoretical Sectors Global")
        iT = h.index("Thread Instructions Executed")
        for r in sec["rows"]:
            if not r[0].strip().isdigit():
                continue
            k = (sec["file"].split("/")[-1], int(r[0]))
            a = agg.setdefault(k, [0, 0, 0, 0, "", 0])
            a[0] += num(r[iI]); a[1] += num(r[iS]); a[2] += num(r[iL]); a[3] += num(r[iG]); a[4] = r[1]; a[5] += num(r[iT])
    ti = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    print("total warp-inst %.1fM, samples %d, SIMD eff %.1f%%" % (ti / 1e6, ts, 100.0 * sum(a[5] for a in agg.values()) / ti / 32))
    for k, a in sorted(agg.items()):
        if a[0] > thresh * ti or a[1] > thresh * ts:
            print("%-18s:%4d inst=%5.1f%% samp=%5.1f%% thr/inst=%4.1f L2loc=%7.1fM L2glob=%7.1fM | %s" %
                  (k[0][:18], k[1], 100 * a[0] / ti, 100 * a[1] / ts, a[5] / max(1, a[0]), a[2] / 1e6, a[3] / 1e6, a[4].strip()[:90]))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.008)

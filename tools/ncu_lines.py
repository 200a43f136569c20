#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source cuda,sass` output per source line:
share of executed instructions, stall samples and L2 sectors (first instance of every distinct kernel)."""
import csv
import sys


def num(x):
    try:
        return int(float(x))
    except Exception:
        return 0


def main(path, thresh=0.008):
    rows = list(csv.reader(open(path)))
    # the page is a sequence of (File Path, Function Name, table) blocks: one per source file of every kernel instance
    sections, cur, fp, fn = [], None, None, None
    for r in rows:
        if r and r[0] == "File Path":
            fp = r[1]
        if r and r[0] == "Function Name":
            fn = r[1]
        if r and r[0] == "Line No" and len(r) > 5:
            cur = {"file": fp, "func": fn, "hdr": r, "rows": []}
            sections.append(cur)
            continue
        if cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    # first instance of every distinct kernel: a (function, file) pair seen again starts a later instance
    done_funcs, order, per_func = set(), [], {}
    seen_pairs = set()
    for sec in sections:
        key = (sec["func"], sec["file"])
        if key in seen_pairs:
            done_funcs.add(sec["func"])
        if sec["func"] in done_funcs:
            continue
        seen_pairs.add(key)
        if sec["func"] not in per_func:
            per_func[sec["func"]] = {}
            order.append(sec["func"])
        agg = per_func[sec["func"]]
        h = sec["hdr"]
        col = lambda name: h.index(name) if name in h else None
        iS, iI, iT = col("# Samples"), col("Instructions Executed"), col("Thread Instructions Executed")
        iL, iG = col("L2 Theoretical Sectors Local"), col("L2 Theoretical Sectors Global")
        get = lambda r, i: num(r[i]) if i is not None else 0
        for r in sec["rows"]:
            if not r[0].strip().isdigit():
                continue
            k = (sec["file"].split("/")[-1], int(r[0]))
            a = agg.setdefault(k, [0, 0, 0, 0, "", 0])
            a[0] += get(r, iI); a[1] += get(r, iS); a[2] += get(r, iL); a[3] += get(r, iG); a[4] = r[1]; a[5] += get(r, iT)
    for func in order:
        agg = per_func[func]
        ti = sum(a[0] for a in agg.values()) or 1
        ts = sum(a[1] for a in agg.values()) or 1
        if len(order) > 1:
            print("== %s" % (func or "?")[:110])
        print("total warp-inst %.1fM, samples %d, SIMD eff %.1f%%" % (ti / 1e6, ts, 100.0 * sum(a[5] for a in agg.values()) / ti / 32))
        for k, a in sorted(agg.items()):
            if a[0] > thresh * ti or a[1] > thresh * ts:
                print("%-18s:%4d inst=%5.1f%% samp=%5.1f%% thr/inst=%4.1f L2loc=%7.1fM L2glob=%7.1fM | %s" %
                      (k[0][:18], k[1], 100 * a[0] / ti, 100 * a[1] / ts, a[5] / max(1, a[0]), a[2] / 1e6, a[3] / 1e6, a[4].strip()[:90]))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.008)

#!/usr/bin/env python
"""Reduce `ncu --page source --csv` (one row per SASS/source line, dozens of MB) to the lines that matter:
per kernel, the top-N lines by warp stall samples, with the stall-reason columns that are non-zero.

    ncu -i prof.ncu-rep --page source --csv | python tools/top_stalls.py [N] > profiles/X_top_stalls.txt
"""
import csv
import sys


def main():
    n_top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rd = csv.reader(sys.stdin)
    hdr = None
    rows = []
    kernel = "?"
    out = []

    def flush():
        if not rows or hdr is None:
            return
        idx = {h: i for i, h in enumerate(hdr)}
        samp = next((h for h in hdr if h.startswith("# Samples") or h == "Warp Stall Sampling (All Samples)"), None)
        if samp is None:
            samp = next((h for h in hdr if "Samples" in h), None)
        if samp is None:
            return
        def val(r, h):
            try:
                return float(r[idx[h]].replace(",", ""))
            except Exception:
                return 0.0
        total = sum(val(r, samp) for r in rows) or 1.0
        stall_cols = [h for h in hdr if h.startswith("stall_") or h.lower().startswith("warp stall")]
        top = sorted(rows, key=lambda r: -val(r, samp))[:n_top]
        out.append("== %s  (total samples %.0f)" % (kernel, total))
        src = next((h for h in hdr if h in ("Source", "SASS", "Instruction")), hdr[1])
        for r in top:
            reasons = sorted(((val(r, h), h) for h in stall_cols if val(r, h) > 0), reverse=True)[:4]
            out.append("%6.2f%%  %-70s %s" % (100 * val(r, samp) / total, r[idx[src]][:70],
                                               " ".join("%s=%.0f" % (h.replace("stall_", ""), v) for v, h in reasons)))

    for row in rd:
        if not row:
            continue
        if row[0].startswith("Kernel Name") or (len(row) == 1 and "(" in row[0]):
            flush()
            rows = []
            hdr = None
            kernel = row[-1] if len(row) > 1 else row[0]
            continue
        if hdr is None:
            hdr = row
            continue
        if len(row) == len(hdr):
            rows.append(row)
    flush()
    print("\n".join(out))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel launch from an ncu report:
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > both.csv ; python tools/ncu_lines.py both.csv [launch_index] [top]"""
import csv
import sys


def main(path, launch=0, top=30):
    rows = list(csv.reader(open(path)))
    marks = [i for i, r in enumerate(rows) if r and r[0] == "File Path"]
    # launches repeat the same sequence of files; a launch starts whenever the first file name reappears
    first = rows[marks[0]][1]
    starts = [m for m in marks if rows[m][1] == first]
    lo = starts[launch]
    hi = starts[launch + 1] if launch + 1 < len(starts) else len(rows)
    agg, filep, hdr = {}, None, None
    for r in rows[lo:hi]:
        if not r:
            continue
        if r[0] == "File Path":
            filep = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif r[0].isdigit() and hdr:
            try:
                n = int(r[hdr.index("Instructions Executed")])
                w = int(r[hdr.index("Warp Stall Sampling (All Samples)")] or 0)
            except ValueError:
                continue
            agg[(filep, int(r[0]), r[1].strip()[:100])] = (n, w)
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[1] for v in agg.values()) or 1
    print("launch %d: %d warp instructions, %d stall samples" % (launch, tot, tots))
    for k, v in sorted(agg.items(), key=lambda kv: -(kv[1][0] / tot + kv[1][1] / tots))[:top]:
        print("%-16s %4d  instr %5.1f%%  stall %5.1f%%  %s" % (k[0], k[1], 100 * v[0] / tot, 100 * v[1] / tots, k[2]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 30)

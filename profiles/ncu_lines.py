"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line.

usage: python profiles/ncu_lines.py <dump.csv> [top_n]
Prints the top lines by executed warp instructions and by stall samples."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows, cur_file, hdr = [], None, None
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] in ("Function Name",) or hdr is None or r[0] == "":
        continue
    try:
        rows.append((cur_file, int(r[0]), r[1].strip(), int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]),
                     int(r[hdr.index("Thread Instructions Executed")])))
    except (ValueError, IndexError):
        pass
ti = sum(x[4] for x in rows)
ts = sum(x[3] for x in rows)
print(f"total warp instructions {ti}, samples {ts}")
print("== by instructions executed")
for f, ln, src, s, i, t in sorted(rows, key=lambda x: -x[4])[:top]:
    print(f"{100 * i / ti:5.1f}% inst {100 * s / max(1, ts):5.1f}% smpl  thr/inst {t / max(1, i):5.1f}  {f}:{ln}  {src[:90]}")
print("== by stall samples")
for f, ln, src, s, i, t in sorted(rows, key=lambda x: -x[3])[:top]:
    print(f"{100 * s / max(1, ts):5.1f}% smpl {100 * i / ti:5.1f}% inst  {f}:{ln}  {src[:90]}")

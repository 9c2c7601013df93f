"""Per-source-line instruction / stall-sample totals of a `ncu --set full --import-source on` report.
Usage: python profiles/src_lines.py report.ncu-rep [top_n]   (needs -lineinfo at compile time)"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, lines, tot_i, tot_s = "", [], 0, 0
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] and r[0].isdigit() and len(r) >= 8 and r[2] == "-":
        inst, samp = int(r[7] or 0), int(r[4] or 0)
        lines.append((inst, samp, cur_file, int(r[0]), r[1].strip()[:110]))
        tot_i += inst
        tot_s += samp
print(f"total warp instructions {tot_i}, stall samples {tot_s}")
for inst, samp, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{inst:>11} {100.0 * inst / max(tot_i, 1):5.1f}%  samples {100.0 * samp / max(tot_s, 1):5.1f}%  {f}:{ln}  {src}")

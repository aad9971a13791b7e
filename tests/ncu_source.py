"""Development aid: per-source-line totals (instructions, stall samples) of the FIRST result in an .ncu-rep.
  python tests/ncu_source.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lines = []
cur_file = ""
nfunc = 0
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Function Name":
        nfunc += 1
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if nfunc > 1 and r and r[0] == "Function Name":
        break
    if r and r[0].isdigit():
        try:
            lines.append((cur_file, int(r[0]), r[1].strip()[:110], int(r[4] or 0), int(r[7] or 0)))
        except ValueError:
            pass
tot_s = sum(x[3] for x in lines) or 1
tot_i = sum(x[4] for x in lines) or 1
print(f"total samples {tot_s}, warp instructions {tot_i}")
print("--- by stall samples")
for f, ln, src, s, i in sorted(lines, key=lambda x: -x[3])[:top]:
    print(f"{100 * s / tot_s:5.1f}% smp {100 * i / tot_i:5.1f}% ins  {f}:{ln}  {src}")

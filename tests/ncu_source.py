"""Development aid: per-source-line totals (warp instructions, stall samples) of one result in an .ncu-rep.
  python tests/ncu_source.py gpurun_out/prof.ncu-rep [result_index=0] [top_n=40]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the dump is: for every result, for every file: "File Path", "Function Name", header, rows
results = []
smp = None
seen_files = set()
cur_file = ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        fn = r[1].split("(")[0]
        if not results or results[-1]["fn"] != fn or cur_file in results[-1]["files"]:
            results.append({"fn": fn, "files": set(), "smp": collections.Counter(), "ins": collections.Counter(), "src": {}})
        results[-1]["files"].add(cur_file)
        continue
    if r[0].isdigit() and results:
        try:
            k = (cur_file, int(r[0]))
            results[-1]["smp"][k] += int(r[4] or 0)
            results[-1]["ins"][k] += int(r[7] or 0)
            results[-1]["src"][k] = r[1].strip()[:105]
        except (ValueError, IndexError):
            pass
print("results:", [(i, x["fn"]) for i, x in enumerate(results)])
R = results[which]
ts, ti = sum(R["smp"].values()) or 1, sum(R["ins"].values()) or 1
print(f"{R['fn']}: samples {ts}, warp instructions {ti}")
print("--- by stall samples")
for k, v in R["smp"].most_common(top):
    print(f"{100 * v / ts:5.1f}% smp {100 * R['ins'][k] / ti:5.1f}% ins  {k[0]}:{k[1]}  {R['src'][k]}")
print("--- by instructions")
for k, v in R["ins"].most_common(top // 2):
    print(f"{100 * v / ti:5.1f}% ins {100 * R['smp'][k] / ts:5.1f}% smp  {k[0]}:{k[1]}  {R['src'][k]}")

"""Development aid: print selected raw metrics of an .ncu-rep (needs ncu on PATH; no GPU needed).
  python tests/ncu_raw.py gpurun_out/prof.ncu-rep [extra_metric_regex]"""
import csv
import io
import re
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
pat = re.compile(
    r"^(Kernel Name|gpu__time_duration.sum|dram__bytes_(read|write).sum|gpu__dram_throughput.avg.pct|l1tex__throughput.avg.pct|lts__throughput.avg.pct"
    r"|sm__throughput.avg.pct|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__occupancy_limit"
    r"|smsp__issue_active.avg.pct|smsp__inst_executed.sum$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|smsp__average_warps_issue_stalled.*_per_issue_active"
    r"|launch__shared_mem_per_block|sm__inst_executed_pipe_lsu|smsp__inst_executed_op_shared" + (("|" + sys.argv[2]) if len(sys.argv) > 2 else "") + ")")
idx = [i for i, x in enumerate(h) if pat.search(x)]
for r in rows[2:]:
    for i in idx:
        print(f"{h[i]} = {r[i]} {rows[1][i]}")
    print("---")

"""Summarise an ncu report (raw page) into the handful of numbers the profiles/ notes quote, plus the top stall reasons
and (with --source) the instructions that collect the most stall samples.  python tools/ncu_summary.py rep.ncu-rep [--source N]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__icc_request_hit_rate.pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg"]
for d in data:
    print("===", d[hdr.index("Kernel Name")][:60])
    for k in KEYS:
        if k in hdr:
            print(f"  {k:86s} {d[hdr.index(k)]} {units[hdr.index(k)]}")
    st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(d[i]))
          for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    st.sort(key=lambda x: -x[1])
    print("  warp stall reasons per issue (top 7): " + ", ".join(f"{n}={v:.2f}" for n, v in st[:7]))
if "--source" in sys.argv:
    n = int(sys.argv[sys.argv.index("--source") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    secs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    for si, s0 in enumerate(secs[:2]):
        h = rows[s0]
        end = secs[si + 1] - 1 if si + 1 < len(secs) else len(rows)
        dd = rows[s0 + 1:end]
        iS, iSrc, iEx = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
        stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        print("--- kernel", rows[s0 - 1][1][:60] if s0 > 0 else "")
        agg = defaultdict(lambda: [0, 0])
        tot = 0
        for r in dd:
            if len(r) <= iS or not r[iS].isdigit():
                continue
            p = r[iSrc].strip().split()
            op = (p[1] if p[0].startswith("@") else p[0]).split(".")[0]
            agg[op][0] += int(r[iS]); agg[op][1] += int(r[iEx]); tot += int(r[iS])
        print("  samples by opcode: " + ", ".join(f"{op} {100 * a / max(tot, 1):.1f}% ({b} exec)" for op, (a, b) in sorted(agg.items(), key=lambda x: -x[1][0])[:14]))
        top = sorted(((int(r[iS]), k) for k, r in enumerate(dd) if len(r) > iS and r[iS].isdigit()), reverse=True)[:n]
        for sm, k in top:
            r = dd[k]
            stl = sorted([(h[i].replace("stall_", ""), int(r[i] or 0)) for i in stall_cols], key=lambda x: -x[1])[:2]
            print(f"  {k:5d} {sm:5d} {r[iSrc].strip()[:64]:64s} {stl}")

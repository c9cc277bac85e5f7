"""Summarise an `ncu --page raw --csv` export: one line per launch with duration, DRAM rate, occupancy, issue
utilisation, registers and the top three stall reasons (warps stalled per issue-active cycle)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
H, U, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(H)}


def val(r, k):
    v = r[ix[k]].replace(",", "") if k in ix else ""
    try:
        return float(v)
    except ValueError:
        return float("nan")


stall = [(i, h.split("issue_stalled_")[1].split("_per")[0]) for i, h in enumerate(H) if "issue_stalled" in h and "per_issue_active" in h]
for r in data:
    name = r[ix["Kernel Name"]].split("(")[0].replace("pcb::", "").replace("void ", "")[:30]
    t = val(r, "gpu__time_duration.sum")
    un = U[ix["gpu__time_duration.sum"]]
    t_us = t / 1e3 if un in ("ns", "nsecond") else (t if un in ("us", "usecond") else t * 1e3)
    bw = val(r, "dram__bytes.sum.per_second")
    ub = U[ix["dram__bytes.sum.per_second"]]
    bw *= {"byte/second": 1e-9, "Kbyte/second": 1e-6, "Mbyte/second": 1e-3, "Gbyte/second": 1.0, "Tbyte/second": 1e3}.get(ub, 1.0)
    st = sorted(((float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else 0.0, n) for i, n in stall), reverse=True)[:3]
    print(f"{name:30s} grid={r[ix['launch__grid_size']]:>7s} {t_us:8.1f}us {bw:6.0f}GB/s dram%={val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
          f"warps%={val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} issue%={val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} "
          f"ipc={val(r, 'sm__inst_executed.avg.per_cycle_elapsed'):4.2f} inst={val(r, 'smsp__inst_executed.sum') / 1e6:7.2f}M regs={r[ix['launch__registers_per_thread']]:>3s} "
          f"lsu%={val(r, 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):4.1f} fma%={val(r, 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):4.1f} alu%={val(r, 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'):4.1f} "
          f"l1%={val(r, 'l1tex__throughput.avg.pct_of_peak_sustained_active'):4.1f} l2%={val(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):4.1f} | "
          + " ".join(f"{n}:{v:.1f}" for v, n in st))

"""Per-kernel table from an ncu raw-page CSV export (ncu -i X.ncu-rep --page raw --csv > X_raw.csv):
duration, achieved DRAM GB/s and % of the measured HBM peak, sectors per global load request, lanes per instruction,
issue-slot utilisation, branch uniformity, registers, the main stall reasons.
  python tools/ncu_table.py gpurun_out/r2_kernels_raw.csv > profiles/r2_kernels_metrics.txt"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    peak = 6650.0


def get(r, name, default=float("nan")):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:  # noqa: BLE001
        return default


def scale(r, name):  # bytes with their unit
    v = get(r, name)
    u = units[col[name]] if name in col else ""
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def dur_s(r):
    v = get(r, "gpu__time_duration.sum")
    u = units[col["gpu__time_duration.sum"]]
    return v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(u, 1e-9)


STALLS = ["long_scoreboard", "short_scoreboard", "wait", "no_instruction", "barrier", "branch_resolving", "lg_throttle", "mio_throttle", "math_pipe_throttle", "not_selected"]
print("ncu --set full --clock-control none; HBM peak = %.1f GB/s (MEASURED_PEAKS.json); one launch per kernel (tools/prof_kernels.py)" % peak)
print("%-44s %9s %8s %7s %6s %6s %6s %6s %5s  %s" % ("kernel", "time", "DRAM", "%HBM", "sec/rq", "thr/in", "issue%", "unif%", "regs", "stall cycles per issued instruction"))
for r in body:
    name = r[col["Kernel Name"]]
    t = dur_s(r)
    dram = scale(r, "dram__bytes_read.sum") + scale(r, "dram__bytes_write.sum")
    req = get(r, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")
    sec = get(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
    st = sorted(((get(r, "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s, 0.0), s) for s in STALLS), reverse=True)[:3]
    print("%-44s %7.1fus %6.1fGB/s %6.2f%% %6.1f %6.2f %6.1f %6.1f %5d  %s" % (
        name[:44], t * 1e6, dram / t / 1e9, 100 * dram / t / 1e9 / peak, sec / req if req else float("nan"),
        get(r, "smsp__thread_inst_executed_per_inst_executed.ratio"), get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        get(r, "smsp__sass_average_branch_targets_threads_uniform.pct"), int(get(r, "launch__registers_per_thread", 0)),
        ", ".join("%s %.1f" % (s, v) for v, s in st)))

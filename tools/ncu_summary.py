"""Summarise ncu output brought back in gpurun_out/ into small tracked files under profiles/.

    python tools/ncu_summary.py full   gpurun_out/prof.ncu-rep   profiles/r01_x_full.md   ["note"]
    python tools/ncu_summary.py list   gpurun_out/launches.csv   profiles/r01_x_launches.md ["note"]
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM (register limit)"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("sm__warps_active.avg.per_cycle_active", "warps active / SM"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.per_cycle_active", "issue slots busy / cycle / SMSP"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads / warp instruction"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (per issue)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
]


def full(rep, out, note):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# ncu --set full summary: %s\n\n%s\n\n" % (rep.split("/")[-1], note))
        for r in data:
            f.write("## %s\n\n| metric | value |\n|---|---|\n" % r[ix["Kernel Name"]])
            for k, label in KEYS:
                if k in ix:
                    f.write("| %s (`%s`) | %s %s |\n" % (label, k, r[ix[k]], units[ix[k]]))
            f.write("\n")
    print("wrote", out)


def launch_list(path, out, note):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(out, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, cold-cache, serialised): %s\n\n%s\n\n" % (path.split("/")[-1], note))
        f.write("| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|\n")
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.3f | %.3f | %.1f %% |\n" % (name, n, ms, ms / n, 100 * ms / tot))
    print("wrote", out)


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    (full if mode == "full" else launch_list)(src, dst, note)

"""Summarise an .ncu-rep (ncu --set full) into a small markdown table for profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more.ncu-rep ...] > profiles/<name>.md
Reads the report with `ncu -i … --page raw --csv` (works without a GPU)."""
import csv
import os
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors (from SMs)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy % (active cycles)"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy % (elapsed)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp cycles per issued instruction"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (global/L2)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (smem/MIO)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall: membar"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction (icache)"),
]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def traffic(reps, out_path):
    """profiles/ncu_traffic.json: mean dram__bytes_read.sum + dram__bytes_write.sum per launch and kernel."""
    import collections
    import json
    acc = collections.defaultdict(list)
    for rep in reps:
        hdr, units, data = load(rep)
        col = {h: i for i, h in enumerate(hdr)}
        ir, iw = col["dram__bytes_read.sum"], col["dram__bytes_write.sum"]
        for r in data:
            name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
            acc[name].append(float(r[ir].replace(",", "")) * UNIT[units[ir]] + float(r[iw].replace(",", "")) * UNIT[units[iw]])
    out = {"source": [os.path.basename(r) for r in reps], "what": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean per launch",
           "bytes_per_launch": {k: int(sum(v) / len(v)) for k, v in acc.items()}, "launches": {k: len(v) for k, v in acc.items()}}
    with open(out_path, "w") as f:
        json.dump(out, f, indent=1)
    print(out)


def main():
    if sys.argv[1] == "--traffic":
        return traffic(sys.argv[3:], sys.argv[2])
    for rep in sys.argv[1:]:
        hdr, units, data = load(rep)
        col = {h: i for i, h in enumerate(hdr)}
        print(f"## {rep.split('/')[-1]}\n")
        names = [r[col["Kernel Name"]].split("(")[0].replace("void ", "") for r in data]
        print("| metric | " + " | ".join(f"{n} #{i}" for i, n in enumerate(names)) + " |")
        print("|---|" + "---|" * len(names))
        for m, label in METRICS:
            if m not in col:
                continue
            i = col[m]
            vals = []
            for r in data:
                v = r[i]
                try:
                    fv = float(v.replace(",", ""))
                    v = f"{fv:.3g}" if abs(fv) < 1e6 else f"{fv:.4g}"
                except ValueError:
                    pass
                vals.append(f"{v} {units[i]}".strip())
            print(f"| {label} (`{m}`) | " + " | ".join(vals) + " |")
        print()


if __name__ == "__main__":
    main()

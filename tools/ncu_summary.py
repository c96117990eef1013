"""Summary of an ncu report (`ncu --set full --clock-control none`): one block per captured launch with the counters DESIGN.md and
profiles/*_summary.md quote (duration, instructions, issue rate, FP64 pipe, shared-memory pipe, instruction cache, dram bytes,
stall reasons).

usage: python tools/ncu_summary.py report.ncu-rep [--traffic-json out.json kernel-substring]
"""
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__icc_requests.sum", "sm__icc_requests_lookup_hit.sum", "sm__icc_requests_lookup_miss.sum",
        "gcc__requests.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    ix = {n: i for i, n in enumerate(h)}
    stall = [n for n in h if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued")]
    blocks = []
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        print(f"kernel: {name}")
        for n in WANT:
            if n in ix:
                print(f"  {n:85s} {r[ix[n]]} {units[ix[n]]}")
        tot = sum(float(r[ix[n]] or 0) for n in stall) or 1.0
        top = sorted(stall, key=lambda n: -float(r[ix[n]] or 0))[:10]
        print("  stall reasons (share of samples): " + ", ".join(f"{n.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * float(r[ix[n]] or 0) / tot:.1f} %" for n in top))
        blocks.append((name, r))
    if "--traffic-json" in sys.argv:
        k = sys.argv.index("--traffic-json")
        path, sub = sys.argv[k + 1], sys.argv[k + 2]
        for name, r in blocks:
            if sub in name:
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd = float(r[ix["dram__bytes_read.sum"]]) * scale[units[ix["dram__bytes_read.sum"]]]
                wr = float(r[ix["dram__bytes_write.sum"]]) * scale[units[ix["dram__bytes_write.sum"]]]
                json.dump({"kernel": name, "grid": r[ix["launch__grid_size"]], "dram_bytes_per_launch": rd + wr,
                           "source": f"ncu --set full capture {rep.split('/')[-1]} (summary committed under profiles/): dram__bytes_read.sum {rd / 1e6:.2f} MB + "
                                     f"dram__bytes_write.sum {wr / 1e6:.2f} MB per launch"}, open(path, "w"), indent=1)
                break


if __name__ == "__main__":
    main()

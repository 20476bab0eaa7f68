"""Turn ncu outputs into the small text summaries committed under profiles/.

    python profiles/summarize.py rep  gpurun_out/prof.ncu-rep   > profiles/<name>.txt     (needs ncu on PATH, no GPU)
    python profiles/summarize.py rep  gpurun_out/asm_raw.csv    > profiles/<name>.txt     (raw page exported by profiles/capture.sh)
    python profiles/summarize.py list gpurun_out/launches.csv   > profiles/<name>.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__ops_path_tensor_src_fp64.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def rep(path):
    if path.endswith(".csv"):          # already reduced on the GPU box (profiles/capture.sh)
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    for r in data:
        print("kernel:", r[name_col][:110])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:90s} {r[i]:>18s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if "smsp__pcsamp_warps_issue_stalled" in h and not h.endswith("_not_issued"):
                try:
                    stalls.append((float(r[i].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        print("  warp stall samples:", ", ".join(f"{h} {v / tot:.1%}" for v, h in sorted(stalls, reverse=True)[:8]))
        print()


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("bsde::<unnamed>::", "").replace("unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':60s} {'launches':>9s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:60s} {v[0]:9d} {v[1] / 1e3:10.2f} {v[1] / v[0]:9.1f} {v[1] / tot:7.1%}")


if __name__ == "__main__":
    (rep if sys.argv[1] == "rep" else launches)(sys.argv[2])

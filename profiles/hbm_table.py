"""Achieved DRAM bandwidth of the HBM-bound kernels from `ncu --set full` raw pages (profiles/capture_r02.sh):
    python profiles/hbm_table.py gpurun_out/r02_hbm_raw.csv gpurun_out/r02_iface_left_raw.csv > profiles/hbm_kernels_r02.txt
Peak = MEASURED_PEAKS.json hbm_gbs (6452.8 GB/s, copy bandwidth measured on this pool's B200s)."""
import csv
import sys

PEAK = 6452.8


def load(path):
    rows = list(csv.reader(open(path)))
    return rows[0], rows[1], rows[2:]


print(f"{'kernel':58s} {'us':>9s} {'read MB':>9s} {'write MB':>9s} {'GB/s':>8s} {'/ 6452.8':>9s} {'ncu dram %':>10s} {'grid':>6s} {'regs':>5s}")
for f in sys.argv[1:]:
    hdr, units, data = load(f)
    g = lambda r, k: r[hdr.index(k)].replace(",", "")
    for r in data:
        name = g(r, "Kernel Name").replace("bsde::<unnamed>::", "").replace("void ", "")[:58]
        dur = float(g(r, "gpu__time_duration.sum"))
        u = units[hdr.index("gpu__time_duration.sum")]
        dur_us = dur / 1e3 if u == "ns" else (dur if u == "us" else dur * 1e3)

        def nbytes(k):
            return float(g(r, k)) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[hdr.index(k)]]

        rd, wr = nbytes("dram__bytes_read.sum"), nbytes("dram__bytes_write.sum")
        gbs = (rd + wr) / dur_us / 1e3
        print(f"{name:58s} {dur_us:9.1f} {rd / 1e6:9.1f} {wr / 1e6:9.1f} {gbs:8.0f} {gbs / PEAK:9.2f} "
              f"{float(g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):10.1f} {g(r, 'launch__grid_size'):>6s} {g(r, 'launch__registers_per_thread'):>5s}")

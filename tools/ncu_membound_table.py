"""One line per launch of an `ncu --set full` report of memory-bound kernels: duration, DRAM bytes moved, achieved
DRAM GB/s and its fraction of the measured HBM peak (MEASURED_PEAKS.json), SM throughput, registers.
Usage: python tools/ncu_membound_table.py report.ncu-rep > profiles/<name>.txt"""
import csv, json, os, subprocess, sys

PEAK = 6543.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def col(hdr, name):
    hit = [i for i, h in enumerate(hdr) if h == name or h.endswith("." + name)]
    return hit[0] if hit else None


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def to_sec(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0}[unit]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {k: col(hdr, k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                   "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
                                   "Grid Size", "Block Size", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                                   "sm__warps_active.avg.pct_of_peak_sustained_active")}
    print(f"# ncu --set full --clock-control none; HBM peak (measured) = {PEAK} GB/s; one cold-L2 launch per row")
    print(f"{'kernel':44s} {'grid':>16s} {'regs':>4s} {'us':>8s} {'DRAM rd MB':>10s} {'DRAM wr MB':>10s} {'GB/s':>7s} {'of peak':>7s} {'SM %':>5s} {'warps %':>7s}")
    for r in rows[2:]:
        t = to_sec(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]])
        rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
        wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        gbs = (rd + wr) / t / 1e9
        name = r[ix["Kernel Name"]].replace("void ", "").split("(")[0][:44]
        print(f"{name:44s} {r[ix['Grid Size']].replace(' ', ''):>16s} {r[ix['launch__registers_per_thread']]:>4s} {t * 1e6:8.1f} "
              f"{rd / 1e6:10.1f} {wr / 1e6:10.1f} {gbs:7.0f} {gbs / PEAK:7.2f} "
              f"{float(r[ix['sm__throughput.avg.pct_of_peak_sustained_elapsed']]):5.1f} "
              f"{float(r[ix['sm__warps_active.avg.pct_of_peak_sustained_active']]):7.1f}")


if __name__ == "__main__":
    main(sys.argv[1])

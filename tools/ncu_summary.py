"""Summarise an ncu --set full report (.ncu-rep) into the few numbers the roofline discussion needs.
Usage: python tools/ncu_summary.py report.ncu-rep > profiles/<name>.txt   (needs `ncu` on PATH, no GPU)."""
import csv, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes (TMA loads)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor (UTCHMMA) inst % of peak"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "tensor pipe busy cycles (per TPC = 2 SMs)"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor-memory cycles active %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts by tensor core % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for n, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"--- launch {n}: {d.get('Kernel Name', '?')[:60]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for key, label in KEYS:
            hit = [h for h in hdr if h == key or h.endswith("." + key)]
            if hit:
                print(f"    {label:48s} {d[hit[0]]} {u[hit[0]]}")


if __name__ == "__main__":
    main(sys.argv[1])

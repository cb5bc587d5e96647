"""Aggregate an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list by kernel:
launches, device time (cold, serialised: compare shares), DRAM bytes and GB/s.  Usage: launch_summary_dram.py X.csv [top]"""
import collections, csv, re, sys


def main(path, top=30):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [set(), 0.0, 0.0])
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"]))[:64]
        a = agg[name]
        a[0].add(r["ID"])
        if r["Metric Name"].startswith("gpu__time"):
            a[1] += v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
        else:
            a[2] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    tot = sum(a[1] for a in agg.values())
    print(f"total kernel time {tot:.3f} ms over {sum(len(a[0]) for a in agg.values())} launches (cold, serialised: shares, not absolutes)")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{a[1]:9.3f} ms {100 * a[1] / tot:5.1f}%  x{len(a[0]):4d}  {a[2] / 1e6:9.1f} MB DRAM  {a[2] / 1e6 / max(a[1], 1e-9):7.0f} GB/s  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)

"""profiles/rNN_conv_dram_traffic_per_step.json from an ncu metrics pass over ONE training step:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --profile-from-start off --csv --log-file gpurun_out/traffic.csv python bench.py --eager --profile-step
    python tools/conv_traffic_from_ncu.py gpurun_out/traffic.csv > profiles/r02_conv_dram_traffic_per_step.json

bench.py reports `roofline.traffic` from this file only when its launch count equals the launches of the timed binary."""
import collections, csv, json, re, sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: dict(ids=set(), dram_read_bytes=0.0, dram_write_bytes=0.0, ncu_time_s=0.0))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6,
             "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0}
    for r in csv.DictReader(lines):
        name = re.sub(r"^void ", "", re.sub(r"[<(].*", "", r["Kernel Name"])).split("::")[-1]
        if name not in ("igemm_kernel", "wgrad_kernel"):
            continue
        v = float(r["Metric Value"].replace(",", "")) * scale[r["Metric Unit"]]
        a = agg[name]
        a["ids"].add(r["ID"])
        if r["Metric Name"] == "dram__bytes_read.sum":
            a["dram_read_bytes"] += v
        elif r["Metric Name"] == "dram__bytes_write.sum":
            a["dram_write_bytes"] += v
        elif r["Metric Name"] == "gpu__time_duration.sum":
            a["ncu_time_s"] += v
    out = {"what": "sum over all launches of one BaselineModel batch-32 training step (ncu --metrics dram__bytes_read.sum,"
                   "dram__bytes_write.sum,gpu__time_duration.sum, bench.py --eager --profile-step)", "kernels": {}}
    for k, a in agg.items():
        out["kernels"][k] = dict(launches=len(a["ids"]), dram_read_bytes=a["dram_read_bytes"],
                                 dram_write_bytes=a["dram_write_bytes"], ncu_time_s=a["ncu_time_s"])
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])

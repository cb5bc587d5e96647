"""Top warp-stall lines of one launch of an ncu --set full --import-source on report.
Usage: python tools/ncu_top_stalls.py report.ncu-rep <launch index> [n lines] [context]"""
import csv, subprocess, sys

rep, launch = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 14
ctx = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:80])
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
seen, data = set(), []
for r in rows[2:]:
    if len(r) != len(hdr) or r[idx["Address"]] in seen:
        continue
    seen.add(r[idx["Address"]])
    data.append(r)


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


stalls = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
tot = sum(num(r[idx["# Samples"]]) for r in data)
agg = sorted(((s, sum(num(r[idx[s]]) for r in data)) for s in stalls), key=lambda kv: -kv[1])
print("samples", int(tot), "instructions", len(data), "| " + ", ".join(f"{s[6:]} {v / tot:.0%}" for s, v in agg[:8]))
order = sorted(range(len(data)), key=lambda i: -num(data[i][idx["# Samples"]]))[:n]
for i in order:
    r = data[i]
    st = sorted(((s[6:], num(r[idx[s]])) for s in stalls), key=lambda kv: -kv[1])[:2]
    print(f'{num(r[idx["# Samples"]]) / tot:6.1%}  {r[idx["Address"]][-5:]}  {r[idx["Source"]][:72]:72s} {st[0][0]}')
    for k in range(max(0, i - ctx), i):
        print(f'          {data[k][idx["Address"]][-5:]}  {data[k][idx["Source"]][:72]}')

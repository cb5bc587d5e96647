"""Opcode histogram of the SASS of every object of the library (cuobjdump -sass), per kernel for the tensor-core
kernels: the proof that the hot kernels are Blackwell-native (UTCHMMA = tcgen05.mma, UTCHMMA.2CTA = cta_group::2,
LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor loads / stores, UTCBAR = tcgen05.commit, SYNCS = mbarrier, UCGABAR =
cluster barrier).  Usage: python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "multimodal_uav_det_b200", "build")
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "UCGABAR", "UTCATOMSWS", "HMMA", "REDG", "ATOMG",
       "LDG", "STG", "LDS", "STS", "FFMA", "FFMA2", "FMUL2", "FADD2", "BAR")


def main():
    for obj in sorted(f for f in os.listdir(OBJ) if f.endswith(".o")):
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
        kernels, cur = collections.OrderedDict(), None
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = kernels.setdefault(m.group(1), collections.Counter())
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
            if m and cur is not None:
                cur[m.group(1)] += 1
        print(f"==== {obj}")
        for name, cnt in kernels.items():
            demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
            total = sum(cnt.values())
            picks = []
            for k in KEY:
                hits = {op: n for op, n in cnt.items() if op == k or op.startswith(k + ".")}
                if hits:
                    picks.append(", ".join(f"{op} x{n}" for op, n in sorted(hits.items(), key=lambda kv: -kv[1])[:6]))
            print(f"  {demangled}: {total} instructions")
            for p in picks:
                print(f"      {p}")


if __name__ == "__main__":
    main()

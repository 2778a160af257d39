"""Per-kernel histogram of the Blackwell-specific SASS opcodes in libcambrian_b200.so (no GPU needed):
UTC*MMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store /
reduce-add, UGETNEXTWORKID = Cluster Launch Control try_cancel, SYNCS = mbarrier ops, plus HMMA (legacy mma.sync: must be 0).

    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cambrian_b200", "libcambrian_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UGETNEXTWORKID", "SYNCS", "HMMA", "MUFU.EX2"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    demangle = {}
    names = re.findall(r"Function : (\S+)", sass)
    if names:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, out))
    cur, hist = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle.get(m.group(1), m.group(1))
            cur = re.sub(r"\((CUtensorMap_st|const |__nv_bfloat16|float|void|unsigned|int|long|cb::)[^<>]*$", "", cur)   # drop the parameter list
            cur = cur.replace("void cb::", "").replace("(int)", "").replace("(bool)", "")
            hist[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        hist[cur]["_total"] += 1
        for key in OPS:
            if key == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    hist[cur][key] += 1
            elif key == "SYNCS":
                if op.startswith("SYNCS"):
                    hist[cur][key] += 1
            elif key == "HMMA":
                if op.startswith("HMMA"):
                    hist[cur][key] += 1
            elif op.startswith(key):
                hist[cur][key] += 1
    print(f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} ({len(hist)} kernels); columns: " + " ".join(OPS) + " | instructions")
    tot = collections.Counter()
    for k, c in hist.items():
        tot.update(c)
        if any(c[o] for o in OPS if o not in ("SYNCS", "MUFU.EX2")):
            print(f"{k[:70]:70s} " + " ".join(f"{c[o]:5d}" for o in OPS) + f" | {c['_total']}")
    print("# kernels without tensor/TMA opcodes (elementwise, norms, SVA window attention, preprocessing): "
          + str(sum(1 for c in hist.values() if not any(c[o] for o in OPS if o not in ('SYNCS', 'MUFU.EX2')))))
    print("TOTAL".ljust(70) + " " + " ".join(f"{tot[o]:5d}" for o in OPS) + f" | {tot['_total']}")
    if tot["HMMA"]:
        sys.exit("legacy HMMA found")


if __name__ == "__main__":
    main()

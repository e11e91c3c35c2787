"""Per-kernel counts of the Blackwell-specific SASS instructions in libshasta_b200.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG (cp.async.bulk.tensor), UBLKCP (cp.async.bulk),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), FFMA2 (packed fp32). Writes a table to stdout.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "shasta_b200", "libshasta_b200.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "FFMA2", "USETMAXREG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["total"] += 1
            if op in OPS:
                counts[cur][op] += 1
    names = list(counts)
    try:
        dm = subprocess.run(["c++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except Exception:  # noqa: BLE001
        pass
    print("# cuobjdump -sass %s  (sm_100a)" % os.path.relpath(LIB, ROOT))
    print("%-58s %7s " % ("kernel", "instrs") + " ".join("%8s" % o for o in OPS))
    tot = collections.Counter()
    for k in names:
        c = counts[k]
        if not any(c[o] for o in OPS):
            continue
        name = re.sub(r"\(.*", "", demangle.get(k, k)).replace("void shasta::", "")
        print("%-58s %7d " % (name[:58], c["total"]) + " ".join("%8d" % c[o] for o in OPS))
        tot.update(c)
    print("%-58s %7d " % ("TOTAL (kernels with any of these)", tot["total"]) + " ".join("%8d" % tot[o] for o in OPS))


if __name__ == "__main__":
    sys.exit(main())

"""Per-kernel counts of the Blackwell-specific SASS instructions in libpaule_b200.so (evidence that the tensor-core path is
tcgen05 / TMEM / bulk-copy code, not recompiled mma.sync): UTCHMMA (tcgen05.mma kind::f16), LDTM / STTM (tcgen05.ld / .st),
UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk, 1-D TMA), UTMALDG / UTMASTG (tensor-map TMA: none -- every bulk movement is a
1-D copy of a pre-swizzled image), SYNCS (mbarrier), LDGSTS (cp.async), HMMA (mma.sync: none expected).

    python tools/sass_summary.py > profiles/sass_summary.txt        (build container: cuobjdump, no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "paule_b200", "lib", "libpaule_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "HMMA", "FFMA", "MUFU",
       "RED", "ATOM", "UCGABAR", "ELECT"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    counts[cur][o] += 1
            counts[cur]["_total"] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    for k, n in zip(counts, names):
        demangle[k] = re.sub(r"\(.*", "", n).replace("void ", "").replace("paule::", "")
    print("# SASS instruction counts per kernel of paule_b200/lib/libpaule_b200.so (cuobjdump -sass, sm_100a); tools/sass_summary.py")
    print(f"{'kernel':58s} " + " ".join(f"{o:>7s}" for o in OPS) + f" {'total':>7s}")
    for k, c in counts.items():
        if c["_total"] == 0:
            continue
        print(f"{demangle[k][:58]:58s} " + " ".join(f"{c[o]:7d}" for o in OPS) + f" {c['_total']:7d}")
    tc = [k for k, c in counts.items() if c["UTCHMMA"] > 0]
    print(f"\n# kernels that issue tcgen05.mma (UTCHMMA): {len(tc)}; kernels with mma.sync (HMMA): "
          f"{sum(1 for c in counts.values() if c['HMMA'] > 0)}; tensor-map TMA (UTMALDG/UTMASTG): "
          f"{sum(1 for c in counts.values() if c['UTMALDG'] + c['UTMASTG'] > 0)} (all bulk movement is 1-D cp.async.bulk of pre-swizzled images)")


if __name__ == "__main__":
    main()

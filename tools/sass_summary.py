"""SASS evidence of the sm_100a instruction classes per kernel of libeditor_b200.so (cuobjdump -sass), written to
profiles/<tag>_sass_summary.txt:  UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA
tensor load / store, UTMAPF = TMA L2 prefetch, UTCBAR = tcgen05.commit, SYNCS = mbarrier, FFMA2 / FMUL2 = packed fp32."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
LIB = os.path.join(ROOT, "editor_b200", "lib", "libeditor_b200.so")
CLASSES = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMALDG.2CTA", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "FFMA2", "FMUL2",
           "MUFU.TANH", "MUFU.EX2", "RED", "ATOMG"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = per.setdefault(re.sub(r"\(CUtensorMap_st.*|\(.*", "", name).replace("void ", ""), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(2)
            cur["_total"] += 1
            for c in CLASSES:
                if op == c or op.startswith(c + ".") or (c.count(".") and c in op):
                    cur[c] += 1
    out = os.path.join(ROOT, "profiles", "%s_sass_summary.txt" % TAG)
    with open(out, "w") as f:
        f.write("# cuobjdump -sass editor_b200/lib/libeditor_b200.so -- static instruction counts per kernel (sm_100a)\n")
        f.write("# %-78s %7s " % ("kernel", "SASS") + " ".join("%12s" % c for c in CLASSES) + "\n")
        tot = collections.Counter()
        for k, c in per.items():
            if not any(c[x] for x in CLASSES):
                continue
            f.write("%-80s %7d " % (k[:80], c["_total"]) + " ".join("%12d" % c[x] for x in CLASSES) + "\n")
            tot.update(c)
        f.write("%-80s %7d " % ("TOTAL (kernels listed)", tot["_total"]) + " ".join("%12d" % tot[x] for x in CLASSES) + "\n")
    print(open(out).read()[:3000])


if __name__ == "__main__":
    main()

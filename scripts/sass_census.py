"""Per-kernel census of the Blackwell-specific SASS in lib/libflowdiff.so (cuobjdump -sass, sm_100a):
UTCHMMA (tcgen05.mma), UTCCP (tcgen05.cp), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA load / store / reduce),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), plus the legacy tensor path HMMA (mma.sync) and LDGSTS (cp.async).
Usage: python scripts/sass_census.py > profiles/r2_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "opticalflowdiffusion_b200", "lib", "libflowdiff.so")
OPS = ("UTCHMMA", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "FFMA2", "MUFU")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    names = re.findall(r"Function : (\S+)", sass)
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for n, d in zip(names, out):
        d = re.sub(r"\((int|bool|unsigned int)\)", "", d).replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        demangle[n] = re.sub(r"\(.*", "", d).replace("void ", "")
    rows = []
    for blk in re.split(r"\n\s*Function : ", sass)[1:]:
        name = blk.split("\n")[0].strip()
        cnt = collections.Counter()
        total = 0
        for line in blk.split("\n"):
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                total += 1
                op = m.group(1)
                for o in OPS:
                    if op.startswith(o):
                        cnt[o] += 1
        rows.append((demangle.get(name, name), total, cnt))
    rows.sort(key=lambda r: (-r[2]["UTCHMMA"], -r[2]["HMMA"], r[0]))
    arch = set(re.findall(r"arch = (sm_\w+)", sass))
    print(f"{os.path.relpath(LIB, ROOT)}: {len(rows)} kernels, architectures {sorted(arch)}")
    print("%-64s %6s " % ("kernel", "instr") + " ".join("%7s" % o for o in OPS))
    tot = collections.Counter()
    for name, total, cnt in rows:
        if not any(cnt[o] for o in OPS[:10]):
            continue
        print("%-64s %6d " % (name[:64], total) + " ".join("%7d" % cnt[o] for o in OPS))
        tot.update(cnt)
    print("%-64s %6s " % ("TOTAL (all kernels)", "") + " ".join("%7d" % tot[o] for o in OPS))


if __name__ == "__main__":
    sys.exit(main())

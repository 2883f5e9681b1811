"""Opcode evidence of the built library (no GPU needed): `cuobjdump -sass` of premvos_b200/lib/libpremvos_b200.so, per kernel the counts
of the Blackwell-specific mnemonics (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP,
tcgen05.commit -> UTCBAR, mbarrier -> SYNCS) plus the legacy tensor path (HMMA: must be absent).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "premvos_b200", "lib", "libpremvos_b200.so")
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR.2CTA.MULTICAST", "UTCBAR", "LDTM", "UTMALDG.4D.2CTA", "UTMALDG.2D.2CTA", "UTMALDG.4D", "UTMALDG.5D",
        "UBLKCP", "UTMACCTL", "SYNCS", "UCGABAR", "FFMA", "HMMA", "LDGSTS", "F2FP"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::|premvos::", "", name)
            name = re.sub(r"\(.*", "", name)
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    cur[k] += 1
                    break
    print("cuobjdump -sass %s (sm_100a), instructions per kernel; Blackwell mnemonics only where present" % os.path.relpath(LIB, ROOT))
    print("%-58s %7s  %s" % ("kernel", "instr", "mnemonic counts"))
    for name, c in kernels.items():
        parts = ["%s x%d" % (k, c[k]) for k in KEYS if c[k] and k not in ("FFMA", "F2FP", "SYNCS") or (k in ("SYNCS",) and c[k])]
        print("%-58s %7d  %s" % (name[:58], c["_total"], "  ".join(parts)))
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print("\nwhole library: " + "  ".join("%s x%d" % (k, tot[k]) for k in KEYS))
    print("legacy tensor path (HMMA / mma.sync): %s" % ("ABSENT" if tot["HMMA"] == 0 else "PRESENT x%d" % tot["HMMA"]))


if __name__ == "__main__":
    main()

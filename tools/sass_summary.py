"""Per-kernel SASS evidence of the Blackwell-native paths in libvgposp.so (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): counts of the FP64 tensor instruction (DMMA), TMA (UTMALDG / UBLKCP), tcgen05.mma
(UTC*MMA), tcgen05.ld (LDTM) and cp.async (LDGSTS) per kernel.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vgposp_b200", "lib", "libvgposp.so")
MNEMONICS = ["DMMA", "UTMALDG", "UBLKCP", "UTCIMMA", "UTCHMMA", "LDTM", "STTM", "UTCBAR", "LDGSTS", "SYNCS", "HMMA", "DFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            per[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            full = m.group(1)
            op = full.split(".")[0]
            if op in ("LDG", "STG") and ".128" in full:
                per[name][op + ".128"] += 1
            if op in ("LDG", "STG") and ".64" in full and ".128" not in full:
                per[name][op + ".64"] += 1
            if op in ("ST", "STG", "RED", "ATOM") and ".SYS" in full:
                per[name]["sys-scope store"] += 1
            for mn in MNEMONICS:
                if op.startswith(mn):
                    per[name][mn] += 1
            per[name]["_total"] += 1
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    print("libvgposp.so SASS summary (cuobjdump -sass; architectures: %s)" % ", ".join(arch))
    print("%-86s %7s " % ("kernel", "instrs") + " ".join("%7s" % m for m in MNEMONICS))
    totals = collections.Counter()
    for k, c in per.items():
        if not any(c[m] for m in MNEMONICS if m not in ("DFMA",)):
            continue
        print("%-86s %7d " % (k[:86], c["_total"]) + " ".join("%7d" % c[m] for m in MNEMONICS))
        totals.update(c)
    print("%-86s %7d " % ("TOTAL over the kernels listed", totals["_total"]) + " ".join("%7d" % totals[m] for m in MNEMONICS))
    print("kernels in the library: %d" % len(per))
    print()
    print("HBM-bound kernels: vector width of their global accesses, system-scope stores (peer mailboxes / flags)")
    cols = ["LDG.128", "STG.128", "LDG.64", "STG.64", "sys-scope store", "DFMA"]
    print("%-86s %7s " % ("kernel", "instrs") + " ".join("%15s" % c for c in cols))
    for k, c in per.items():
        if re.search(r"downdate|trigemv|kernel_(sym|rect)_kernel<0, 3,|peer_step|lazy_step|score_kernel|digit_planes|dist_barrier", k):
            print("%-86s %7d " % (k[:86], c["_total"]) + " ".join("%15d" % c[x] for x in cols))


if __name__ == "__main__":
    sys.exit(main())

"""Blackwell evidence from the built library: per kernel, the counts of the SASS mnemonics that prove tcgen05 / TMEM /
TMA / cp.async use (B200_PROFILING.md "What proves a Blackwell-native kernel").

    python tools/sass_extract.py > profiles/sass_extract.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "no-time-to-train_b200", "libnttt_b200.so")
PATTERNS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "VIMNMX3", "REDUX",
            "VOTE", "POPC", "HMMA", "HGMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        per[cur]["instructions"] += 1 if re.search(r"/\*[0-9a-f]{4}\*/", line) else 0
        for p in PATTERNS:
            if re.search(r"\b" + p, line):
                per[cur][p] += 1
    arch = re.findall(r"arch = (sm_\w+)", sass)
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}   (arch: {sorted(set(arch))})")
    print(f"# {'kernel':58s} {'instr':>6s}  " + " ".join(f"{p:>8s}" for p in PATTERNS))
    tot = collections.Counter()
    for k, c in per.items():
        print(f"{k[:60]:60s} {c['instructions']:6d}  " + " ".join(f"{c[p]:8d}" for p in PATTERNS))
        tot.update(c)
    print(f"{'TOTAL':60s} {tot['instructions']:6d}  " + " ".join(f"{tot[p]:8d}" for p in PATTERNS))
    print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM), UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk,")
    print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier try_wait / arrive, LDGSTS = cp.async, HMMA / HGMMA (legacy mma) must be 0")


if __name__ == "__main__":
    main()

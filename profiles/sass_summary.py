"""Per-kernel count of the SASS mnemonics that prove the Blackwell-native path (B200_PROFILING.md): UTCHMMA
(tcgen05.mma, incl. .2CTA), LDTM (tcgen05.ld), UTMALDG (TMA tensor loads, incl. .MULTICAST), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), and the legacy HMMA (mma.sync -- must be 0). Run in the build container (no GPU needed):

    python profiles/sass_summary.py > profiles/rNN_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "robot_aware_control_b200", "lib", "libracb200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMALDG.MULTICAST", "UTCBAR", "SYNCS", "HMMA", "FFMA", "total"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\((?!anonymous).*", "", name)
            per.setdefault(cur, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            c = per[cur]
            c["total"] += 1
            base = op.split(".")[0]
            if base in ("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "SYNCS", "FFMA"):
                c[base] += 1
            if op.startswith("HMMA"):
                c["HMMA"] += 1
            if base == "UTCHMMA" and ".2CTA" in op:
                c["UTCHMMA.2CTA"] += 1
            if base == "UTMALDG" and ".MULTICAST" in op:
                c["UTMALDG.MULTICAST"] += 1
    print(f"# SASS mnemonic counts per kernel of {os.path.relpath(SO, ROOT)} (cuobjdump -sass, sm_100a)")
    print(f"{'kernel':72s} " + " ".join(f"{k:>9s}" for k in KEYS))
    tot = collections.Counter()
    for name, c in per.items():
        if not any(c[k] for k in ("UTCHMMA", "LDTM", "UTMALDG", "HMMA")) and "--all" not in sys.argv:
            continue
        print(f"{name[:72]:72s} " + " ".join(f"{c[k]:9d}" for k in KEYS))
    for c in per.values():
        tot.update(c)
    print(f"{'ALL KERNELS (' + str(len(per)) + ')':72s} " + " ".join(f"{tot[k]:9d}" for k in KEYS))


if __name__ == "__main__":
    main()

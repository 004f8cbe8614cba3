"""Count the Blackwell-specific SASS mnemonics of every kernel in libqsft_b200.so (cuobjdump -sass; runs without a GPU).
usage: python tools/sass_evidence.py > profiles/r1_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "qsft_b200", "libqsft_b200.so")
WATCH = [("UTC[A-Z]*MMA", "tcgen05.mma (UTCIMMA = kind::i8; .SP = 2:4 sparse)"), ("UTCCP", "tcgen05.cp (smem -> TMEM)"),
         ("UTCBAR", "tcgen05.commit (mbarrier arrive)"), ("LDTM", "tcgen05.ld"), ("STTM", "tcgen05.st"),
         ("UTMALDG", "TMA tensor load"), ("UTMASTG", "TMA tensor store"), ("UBLKCP", "bulk copy"), ("SYNCS", "mbarrier ops"),
         ("IDP", "dp4a integer dot product"), ("HMMA|IMMA", "legacy mma.sync tensor path"), ("ATOMG|REDG|RED\\b|ATOM\\b", "global atomics / reductions"),
         ("LDG", "global loads"), ("STG", "global stores"), ("SHFL", "warp shuffles"), ("DFMA|DADD|DMUL", "fp64 arithmetic")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
            per[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            per[name]["_total"] += 1
            for pat, _ in WATCH:
                if re.match(f"(?:{pat})(?:\\.|$)", op):
                    per[name][pat] += 1
            if op.startswith("UTC") and "MMA" in op:
                per[name]["variant:" + op] += 1
    print("SASS evidence for qsft_b200/libqsft_b200.so (sm_100a), one line per kernel: instruction counts by mnemonic class")
    print("legend: " + "; ".join(f"{p} = {d}" for p, d in WATCH))
    print()
    for k, c in per.items():
        parts = [f"{p}={c[p]}" for p, _ in WATCH if c[p]]
        variants = [f"{v[8:]}x{n}" for v, n in c.items() if v.startswith("variant:")]
        print(f"{k}: total={c['_total']} " + " ".join(parts) + (("  [" + ", ".join(variants) + "]") if variants else ""))
    ptx_strings()


def ptx_strings():
    """Distinct tcgen05 / TMA / cluster PTX instructions written in the sources (the library carries SASS only)."""
    import glob
    print()
    print("inline PTX in the sources (sparsity, cta_group and multicast are PTX qualifiers / descriptor bits, not SASS mnemonics):")
    for path in sorted(glob.glob(os.path.join(ROOT, "qsft_b200", "csrc", "*.cu"))):
        src = open(path).read()
        toks = sorted(set(re.findall(r"(?:tcgen05|cp\.async\.bulk|mbarrier|mapa|barrier\.cluster|fence\.proxy)[a-z0-9_.:]*", src)))
        if toks:
            print(f"{os.path.basename(path)}: " + ", ".join(toks))


if __name__ == "__main__":
    sys.exit(main())

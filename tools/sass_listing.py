"""SASS evidence per kernel of the built library: counts of the Blackwell tensor-core / TMEM / TMA / mbarrier mnemonics.
usage: python tools/sass_listing.py [lib.so] > profiles/rN_sass_tcgen05.txt"""
import collections, os, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "hallucidet_b200", "libhallucidet_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
KEEP = re.compile(r"^(UTC|UTMA|LDTM|STTM|SYNCS|ACQBULK|ELECT|LDGSTS|HMMA|LDSM|STSM|REDG|RED\.|ATOMS|ATOMG|UBLKCP|UCGABAR|MEMBAR)")
print("# SASS evidence (cuobjdump -sass hallucidet_b200/libhallucidet_b200.so), per kernel: Blackwell tensor-core / TMEM / TMA mnemonics")
print("# tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, TMA -> UTMALDG / UTMASTG, mbarrier -> SYNCS, cp.async -> LDGSTS,")
print("# mma.sync -> HMMA (the HBM-bound 16/32-channel layers and the 7x7 stems only), ldmatrix / stmatrix -> LDSM / STSM, red.global -> REDG\n")
blocks = re.split(r"\n\s*Function : ", sass)[1:]
for name, block in zip(names, blocks):
    ops = collections.Counter()
    n = 0
    for m in re.finditer(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", block):
        n += 1
        op = m.group(1)
        if KEEP.match(op):
            ops[op] += 1
    if not ops:
        continue
    print(f"{name.strip()}  ({n} instructions)")
    for op, c in sorted(ops.items()):
        print(f"    {op:40s} x{c}")
    print()

"""Blackwell-specific / notable SASS mnemonics per kernel of libavctc_b200.so (cuobjdump -sass, runs without a GPU):
the evidence that the tcgen05 / TMEM / TMA / mbarrier / cluster / PDL paths are really in the binary.

    python tools/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodal-av-model_b200", "csrc", "libavctc_b200.so")
NOTABLE = re.compile(r"^(UTCHMMA|UTCQMMA|UTCBAR|UTCCP|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|UBLKPF|SYNCS|HMMA|REDUX|CREDUX|MATCH|"
                     r"PREEXIT|ACQBULK|NANOSLEEP|CCTL|UCGABAR|CGAERRBAR|ERRBAR|ATOMS|ATOMG|RED|MEMBAR|FENCE|DSETP|LDGSTS|MAPA|ST\.E\..*CLUSTER)")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            sym = m.group(1)
            dem = subprocess.run(["c++filt", sym], capture_output=True, text=True).stdout.strip()
            k = re.search(r"avctc::(\w+)", dem)
            name = k.group(1) if k else dem.split("(")[0]
            per.setdefault(name, [0, collections.Counter()])
            per[name][0] += 1
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
        if m and name and NOTABLE.match(m.group(1)):
            op = m.group(1)
            if op.startswith(("RED.", "ATOMG.", "ATOMS.", "MEMBAR.", "DSETP.", "FENCE.", "CCTL.")):
                op = op.split(".")[0] + "." + op.split(".")[1]
            per[name][1][op] += 1
    print("# cuobjdump -sass of libavctc_b200.so (sm_100a): Blackwell-specific / notable SASS mnemonics per kernel (all instantiations summed)")
    print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM), UTCBAR = tcgen05.commit, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk,")
    print("# SYNCS.* = mbarrier, UCGABAR = barrier.cluster, MAPA = mapa (DSMEM address), HMMA = mma.sync, (C)REDUX = redux.sync,")
    print("# PREEXIT/ACQBULK = griddepcontrol (PDL), LDGSTS = cp.async")
    for name, (n, c) in per.items():
        if not c:
            continue
        print(f"{name} ({n} instantiation{'s' if n != 1 else ''}): " + ", ".join(f"{op} x{k}" for op, k in c.most_common()))


if __name__ == "__main__":
    main()

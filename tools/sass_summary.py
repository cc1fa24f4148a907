#!/usr/bin/env python
"""Per-kernel SASS evidence: counts of the tensor-core / TMEM / TMA / reduction mnemonics in the built library (cuobjdump -sass).

usage: python tools/sass_summary.py [lib.so] > profiles/rNN/sass_mnemonics.txt
UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc,
UTMALDG / UTMASTG = cp.async.bulk.tensor load / store, SYNCS = mbarrier, REDG = red.global, ELECT = elect.sync.
"""
import collections
import os
import re
import subprocess
import sys

KEEP = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "REDG", "RED", "ATOMG", "ATOM", "ATOMS", "ELECT",
        "SYNCS", "LDGSTS", "HMMA", "UBLKCP")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "jnerf_b200", "libngp_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_]*(?:\.[A-Za-z0-9_\.]+)?)", line)
        if m:
            op = m.group(1)
            if op.split(".")[0] in KEEP:
                counts[cur][op] += 1
            counts[cur]["_total"] += 1
    print(f"# {os.path.basename(lib)}: mnemonic counts per kernel (static SASS, sm_100a)")
    for k, v in counts.items():
        if not v["_total"]:
            continue
        name = subprocess.run(["cu++filt", k], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\((?:anonymous namespace|unnamed)\)::|<unnamed>::", "", name)
        name = re.sub(r"\(.*", "", name)
        items = ", ".join(f"{o} x{c}" for o, c in sorted(v.items()) if o != "_total")
        print(f"{name}: {v['_total']} instructions; {items or '-'}")


if __name__ == "__main__":
    main()

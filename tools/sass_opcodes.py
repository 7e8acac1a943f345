"""Opcode histogram per kernel of the built library: evidence that the hot kernels carry tcgen05 / TMEM / TMA instructions.
    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt
Counts, per kernel symbol of libatmvfi_b200.so (cuobjdump -sass): UTCHMMA (tcgen05.mma kind::tf32 / kind::f16, incl. .2CTA),
LDTM / STTM (tcgen05.ld / st), UTMALDG (cp.async.bulk.tensor loads), UTMASTG, UTCBAR (tcgen05.commit), SYNCS (mbarrier),
UTCATOMSWS / UTCCP, plus FFMA / HMMA for contrast."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "atm-vfi_b200", "atmvfi", "libatmvfi_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "UTCATOMSWS", "ELECT", "HMMA", "FFMA", "FFMA2", "LDG", "STG", "LDS", "STS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P(?:\d+|T)\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m:
            op, mods = m.group(1), m.group(2)
            cur[op] += 1
            if op == "UTCHMMA" and ".2CTA" in mods:
                cur["UTCHMMA.2CTA"] += 1
            if op == "UTMALDG":
                dims = re.search(r"\.(\dD)", mods)
                cur["UTMALDG." + (dims.group(1) if dims else "?")] += 1
            if op == "UTCBAR" and "MULTICAST" in mods:
                cur["UTCBAR.MULTICAST"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a), {len(kernels)} kernels")
    cols = OPS + ["UTCHMMA.2CTA", "UTMALDG.2D", "UTMALDG.4D", "UTCBAR.MULTICAST"]
    tot = collections.Counter()
    groups = collections.OrderedDict()
    for (sym, cnt), name in zip(kernels.items(), demangle):
        base = re.sub(r"[<(].*", "", name.replace("(anonymous namespace)::", "").replace("void ", ""))
        g = groups.setdefault(base, [0, collections.Counter()])
        g[0] += 1
        g[1].update(cnt)
        tot.update(cnt)
    print("kernel (instantiations) | " + " | ".join(cols))
    for base, (n, cnt) in groups.items():
        print(f"{base} ({n}) | " + " | ".join(str(cnt.get(c, 0)) for c in cols))
    print("TOTAL | " + " | ".join(str(tot.get(c, 0)) for c in cols))


if __name__ == "__main__":
    main()

"""Per-kernel mnemonic counts of the shipped library: python tools/sass_digest.py > profiles/sass_digest_r02.txt"""
import collections
import os
import re
import subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "nwchem_b200", "lib", "libnwc_triples.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = ("BAR.SYNC", "BRA", "BRX", "DADD", "DFMA", "DMMA.8x8x4", "DMUL", "LDG.E", "LDL", "LDS.128", "LDS.64", "MUFU.RCP64H",
        "STG.E", "STL", "STS.128", "STS.64", "SYNCS.ARRIVE.TRANS64", "SYNCS.EXCH", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "UBLKCP",
        "WARPSYNC", "UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG")
print("# SASS digest of nwchem_b200/lib/libnwc_triples.so (sm_100a), round 2 final build: `cuobjdump -sass`, per kernel the count of the mnemonics that prove the path.")
print("# DMMA.8x8x4 = FP64 tensor-core MMA (mma.sync m8n8k4 f64); UBLKCP = cp.async.bulk (TMA bulk copy, G->S); SYNCS.* = mbarrier arrive / expect-tx / try_wait.")
print("# No UTCxMMA / LDTM / UTMALDG: tcgen05 has no f64 kind and the operand blocks are 1-D 4 KiB bulk copies, so DMMA via mma.sync is the FP64 tensor path of sm_100a.")
print("# fused_kernel<DUMP,TIMING,RAGGED,ORDER,LAMBDA>: <0,0,0,*,0> = aligned (T) tuples (no spills), <0,0,1,*,0> = ragged (T) tuples (guarded DMMA groups),")
print("# <1,..> = validation dump, <0,1,..> = phase-clock build, LAMBDA 1 = two-sided tuples of Lambda-CCSD(T), LAMBDA 2 = CR-CCSD(T) (dual-energy tuples).")
print("# The LAMBDA 0 / 1 instantiations are byte-identical to the builds every number of DESIGN 6-8 was measured with.\n")
cur, counts, total = None, None, 0


def flush():
    if cur:
        print(cur)
        print("    " + ", ".join(f"{k}={counts[k]}" for k in sorted(counts)) + f", total={total}")


for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        cur, counts, total = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m and cur:
        total += 1
        op = m.group(1)
        for w in want:
            if op == w or op.startswith(w + "."):
                counts[w] += 1
                break
flush()

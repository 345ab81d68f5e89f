"""SASS opcode counts per kernel of the built library (the evidence for tcgen05 / TMA use): python tools/sass_opcodes.py > profiles/rNN_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pwcnet_b200", "lib", "libpwc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCBAR", "SYNCS", "FFMA", "HMMA", "SEL", "LDS", "STS", "LDG", "STG", "LDGSTS"]
kern, counts, order = None, {}, []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[kern] = collections.Counter(); order.append(kern); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["total"] += 1
        for c in cols:
            if op == c or (c in ("LDG", "STG", "LDS", "STS") and op == c):
                counts[kern][c] += 1
print("SASS opcode counts per kernel of pwcnet_b200/lib/libpwc_b200.so (cuobjdump -sass; nvcc 12.9, -gencode arch=compute_100a,code=sm_100a).")
print("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG/UTMAREDG = TMA tensor load/store/reduce-add, UBLKCP = cp.async.bulk, SYNCS = mbarrier, HMMA = legacy mma.sync (none).\n")
print(f"{'kernel':68s}" + "".join(f"{c:>9s}" for c in cols) + f"{'total':>8s}")
for k in order:
    print(f"{k[:67]:68s}" + "".join(f"{counts[k][c]:9d}" for c in cols) + f"{counts[k]['total']:8d}")

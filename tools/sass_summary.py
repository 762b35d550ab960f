#!/usr/bin/env python
"""Static SASS instruction counts per kernel of libpolycracker_b200.so (cuobjdump -sass | this script)."""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "polycracker-unofficial-mirror_b200/libpolycracker_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
cols = ["UBLKCP", "SYNCS", "ATOMS", "ATOMG", "REDG", "SHFL", "BREV", "LDG", "STG", "LDS", "STS", "BAR", "MATCH", "UTMALDG", "UTCHMMA", "UTCBAR", "LDTM", "HMMA"]
out = []
parts = sass.split("Function : ")[1:]
for part, name in zip(parts, names):
    name = re.sub(r"\(.*", "", name).replace("void pc::", "").replace("pc::", "")
    ops = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", part, re.M)
    cnt = collections.Counter(o.split(".")[0] for o in ops)
    out.append((name, len(ops), [cnt.get(c, 0) for c in cols]))
print("SASS instruction counts per kernel of libpolycracker_b200.so (cuobjdump -sass, sm_100a), static counts.")
print("UBLKCP/SYNCS = 1-D TMA bulk copy + mbarrier (k_pack); UTMALDG = 2-D TMA tensor load, UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit,")
print("LDTM = tcgen05.ld (all four: gram::k_gram_tf32, the one dense contraction of the path); ATOMS = shared-memory atomics; ATOMG/REDG = global atomics.\n")
print("%-72s %6s " % ("kernel", "total") + " ".join("%7s" % c for c in cols))
for name, tot, v in sorted(out, key=lambda x: -x[1]):
    print("%-72s %6d " % (name[:72], tot) + " ".join("%7d" % x for x in v))

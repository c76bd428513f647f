"""profiles/sass_summary.txt: per kernel of libbetapose_b200.so, how many tcgen05 / TMEM / TMA instructions its SASS holds
(B200_PROFILING.md: `tcgen05.mma` -> UTC*MMA, `tcgen05.ld` -> LDTM, TMA -> UTMALDG / UTMASTG, legacy tensor path -> HMMA).
    python scripts/sass_summary.py > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "betapose_b200", "libbetapose_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
ops = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMALDG.4D.IM2COL", "UTMASTG", "UTMAPF", "SYNCS", "UCGABAR", "HMMA", "DFMA", "DADD", "DMUL", "IMAD", "SHFL"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    per[cur]["total"] += 1
    base = op.split(".")[0]
    per[cur][base] += 1
    if op.startswith("UTCHMMA.2CTA"):
        per[cur]["UTCHMMA.2CTA"] += 1
    if op.startswith("UTMALDG.4D.IM2COL"):
        per[cur]["UTMALDG.4D.IM2COL"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"SASS opcode counts per kernel of {os.path.basename(so)} (cuobjdump -sass, sm_100a); columns: " + " ".join(ops))
tot = collections.Counter()
for (name, c), dn in zip(per.items(), demangle):
    short = re.sub(r"\(.*", "", dn).replace("void ", "").replace("(anonymous namespace)::", "")
    print(f"{short[:70]:70s} total {c['total']:6d} | " + " ".join(f"{o}={c[o]}" for o in ops if c[o]))
    tot.update(c)
print("ALL KERNELS: " + " ".join(f"{o}={tot[o]}" for o in ops))

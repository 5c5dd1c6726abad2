"""Per-kernel census of the Blackwell-specific SASS mnemonics in libasr_b200.so -> profiles/r2_sass_census.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "adaptive-surface-reconstruction_b200", "asr_b200", "libasr_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
out = ["Round 2: SASS census of adaptive-surface-reconstruction_b200/asr_b200/libasr_b200.so (sm_100a), `cuobjdump -sass`",
       "mnemonics: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (1-D TMA), UTMALDG = cp.async.bulk.tensor",
       "(.GATHER4 = tile::gather4, the optional gather path of the gx kernel), UBLKRED = cp.reduce.async.bulk, LDGSTS = cp.async,",
       "SYNCS = mbarrier operations, UTCBAR = tcgen05.commit, HMMA = legacy mma.sync (none).  Kernels without any of these are omitted.", ""]
rows = []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    c = collections.Counter()
    for m in re.finditer(r"\b(UTCHMMA|UTCQMMA|LDTM|STTM|UBLKCP|UBLKRED|UTMALDG[A-Z0-9.]*|UTMASTG|LDGSTS|SYNCS|HMMA|UTCBAR)\b", f):
        k = m.group(1)
        if k.startswith("UTMALDG"):
            k = "UTMALDG" + (".GATHER4" if "GATHER4" in k else "")
        c[k] += 1
    if any(c.get(k) for k in ("UTCHMMA", "UBLKCP", "UTMALDG", "UTMALDG.GATHER4", "LDGSTS", "UBLKRED", "HMMA")):
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        d = d.replace("(anonymous namespace)::", "").replace("asrb::", "")
        d = re.sub(r"\((?!bool|int).*", "", d)[:70]
        rows.append((d, c))
for d, c in sorted(rows, key=lambda r: (-r[1].get("UTCHMMA", 0), r[0])):
    out.append("%-72s %s" % (d, "  ".join("%s=%d" % kv for kv in sorted(c.items()))))
open(os.path.join(ROOT, "profiles", "r2_sass_census.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))

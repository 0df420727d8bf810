"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the built library (profiles/sass_summary.txt):
   python scripts/sass_summary.py > profiles/sass_summary.txt
UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA engine, 1-D),
SYNCS = mbarrier, IMMA = mma.sync int8 (tensor cores, decode matvec of the integer formats), HMMA = mma.sync f16,
IDP = dp4a, FFMA2 = fma.rn.f32x2."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "xalm_b200", "libxalm_cuda.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "IMMA", "HMMA", "IDP", "FFMA2"]
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
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        for k in keys:
            if op.startswith(k):
                per[cur][k] += 1
dem = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
print(f"# {os.path.relpath(lib, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass), sm_100a")
print("# " + " ".join(f"{k:>8s}" for k in keys) + "  kernel")
for (name, c), d in zip(per.items(), dem):
    tot.update(c)
    if sum(c.values()) == 0:
        continue
    d = re.sub(r"\(.*", "", d).replace("xalm::", "").replace("void ", "")
    print("  " + " ".join(f"{c[k]:8d}" for k in keys) + "  " + d[:110])
print("# " + " ".join(f"{tot[k]:8d}" for k in keys) + "  TOTAL")

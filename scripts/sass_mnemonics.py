"""Counts of the Blackwell-specific SASS mnemonics per kernel of lib/libsurgvid.so (evidence that the tcgen05 / TMEM / TMA paths are in
the shipped binary).  usage: cuobjdump -sass lib/libsurgvid.so > lib.sass; python scripts/sass_mnemonics.py lib.sass > profiles/rNN/sass_mnemonics.txt"""
import collections, re, subprocess, sys
name, cnt = None, collections.OrderedDict()
for line in open(sys.argv[1]):
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); cnt[name] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        cnt[name][m.group(1).split(".")[0]] += 1
ops = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "HMMA", "LDGSTS", "MUFU", "FFMA2")
print("SASS mnemonic counts per kernel of lib/libsurgvid.so (cuobjdump -sass)")
print("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA load / store, UTMAPF = tensor-map prefetch, HMMA = legacy mma.sync, LDGSTS = cp.async\n")
tot = collections.Counter()
agg = collections.OrderedDict()
for k, c in cnt.items():
    dem = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"^void ", "", dem).replace("sv::(anonymous namespace)::", "").replace("(anonymous namespace)::", "")
    short = re.sub(r"\((CUtensorMap|sv::|float|__nv|unsigned|int|long|const).*$", "", dem)
    base = re.sub(r"<.*", "", short)
    a = agg.setdefault(base, [0, collections.Counter(), 0])
    a[0] += 1; a[2] += sum(c.values())
    for o in ops:
        a[1][o] += c[o]; tot[o] += c[o]
for base, (n, c, instr) in agg.items():
    sel = {o: c[o] for o in ops if c[o]}
    print(f"{base:38s} {n:3d} variant(s) {instr:7d} instr  {sel}")
print("\ntotal", dict(tot))

"""Per-kernel SASS instruction histogram of the built library (cuobjdump -sass): the mnemonics that show what the
kernels run on -- UTCHMMA / UTCBAR (tcgen05.mma / commit), UTMALDG (TMA), LDTM (tcgen05.ld), SYNCS (mbarrier), FMNMX3, REDUX,
LDGSTS (cp.async), FFMA ... Usage: sass_digest.py [lib.so] > profiles/rNN_sass_digest.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fpsg_b200", "libfpsg_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print(f"library: {os.path.basename(lib)}   cubins: {', '.join(arch)}")
KEY = ("UTC", "UTMA", "LDTM", "SYNCS", "FMNMX3", "REDUX", "CREDUX", "LDGSTS", "FFMA", "FADD", "FMUL", "MUFU", "HMMA", "DFMA", "ATOM", "RED")
for k, c in hist.items():
    total = sum(c.values())
    base = collections.Counter()
    for op, n in c.items():
        base[op.split(".")[0]] += n
    sel = {op: n for op, n in c.items() if op.startswith(("UTC", "UTMA", "LDTM", "SYNCS"))}
    top = ", ".join(f"{op} {n}" for op, n in base.most_common(10))
    print(f"\n== {k[:150]}\n   {total} instructions; top: {top}")
    keyed = ", ".join(f"{op} {n}" for op, n in sorted(base.items()) if op.startswith(KEY))
    if keyed:
        print(f"   of note: {keyed}")
    if sel:
        print("   tensor core / TMA / mbarrier: " + ", ".join(f"{op} {n}" for op, n in sorted(sel.items())))

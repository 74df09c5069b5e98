"""cuobjdump -sass over libplas.so -> profiles/<tag>_sass_summary.txt: per-kernel counts of the instructions that prove the
sm_100a paths (tcgen05.mma / ld / st / commit, TMA, bulk copies, cluster launch control, mbarriers)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(ROOT, "phones_las_b200", "csrc", "libplas.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cols = [("UTCHMMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"),
        ("UTCBAR", r"\bUTCBAR"), ("CLC", r"\bUCGABAR_|\bUCLC|CANCEL|\bUGETNEXTWORKID|NEXTWORK"), ("HMMA", r"\bHMMA"), ("MUFU.TANH", r"MUFU\.TANH"),
        ("SYNCS", r"\bSYNCS"), ("STAS", r"\bSTAS"), ("LDGSTS", r"\bLDGSTS")]
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
        for name, pat in cols:
            if re.search(pat, line):
                counts[cur][name] += 1
with open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt"), "w") as f:
    f.write("# cuobjdump -sass phones_las_b200/csrc/libplas.so: instruction counts per kernel (tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM,\n"
            "# cp.async.bulk.tensor = UTMALDG, cp.async.bulk = UBLKCP, tcgen05.commit = UTCBAR, CLC = cluster launch control / cluster barrier ops,\n"
            "# mma.sync = HMMA, st.async = STAS, cp.async = LDGSTS, mbarrier ops = SYNCS); scripts/sass_summary.py\n")
    f.write(f"{'kernel':72s}" + "".join(f"{n:>10s}" for n, _ in cols) + "\n")
    for k, c in counts.items():
        if any(c[n] for n, _ in cols):
            f.write(f"{k[:72]:72s}" + "".join(f"{c[n]:10d}" for n, _ in cols) + "\n")
print(open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt")).read()[:3000])

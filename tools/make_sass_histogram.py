"""profiles/r02_sass_opcodes.csv: Blackwell-specific and atomic SASS opcodes per kernel of the built library
(`cuobjdump -sass`; runs without a GPU).  `python tools/make_sass_histogram.py`"""
import collections
import csv
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "salib_b200", "lib", "libsalib_b200.so")
PAT = re.compile(r"\b(UTC\w*MMA[\w.]*|UTMALDG[\w.]*|UTMASTG[\w.]*|LDTM[\w.]*|STTM[\w.]*|UBLKCP[\w.]*|UTCBAR[\w.]*|"
                 r"UTCATOMSWS[\w.]*|SYNCS[\w.]*|ATOMS[\w.]*|ATOMG[\w.]*|RED[\w.]*|REDUX[\w.]*|LDGSTS[\w.]*|DMMA[\w.]*)\b")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = {}
hist = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist.setdefault(cur, collections.Counter())
        continue
    if cur is None or "/*" not in line:
        continue
    body = line.split("*/", 1)[1] if "*/" in line else line
    for op in PAT.findall(body.split(";")[0]):
        hist[cur][op] += 1
names = list(hist)
out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
demangle = dict(zip(names, out)) if len(out) == len(names) else {n: n for n in names}
path = os.path.join(ROOT, "profiles", "r02_sass_opcodes.csv")
with open(path, "w") as f:
    f.write("# SASS opcode histogram of salib_b200/lib/libsalib_b200.so (cuobjdump -sass, sm_100a; final round-2 build), "
            "Blackwell-specific and atomic opcodes per kernel (tools/make_sass_histogram.py)\n")
    f.write("# UTC*MMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, "
            "UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, ATOMS = shared-memory atomics / red.shared\n")
    w = csv.writer(f)
    w.writerow(["kernel", "opcode", "count"])
    for n in names:
        short = re.sub(r"\(.*", "", demangle[n])
        for op, c in sorted(hist[n].items()):
            w.writerow([short, op, c])
print("wrote", path, sum(len(h) for h in hist.values()), "rows")

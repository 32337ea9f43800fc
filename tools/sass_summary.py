"""Per-kernel SASS mnemonic counts of the built library (evidence that the hot path is tcgen05 / TMA / tensor memory):
    python tools/sass_summary.py [drawingspinup_b200/lib/libdsu_b200.so] > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "drawingspinup_b200/lib/libdsu_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, out = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        out[cur] = []
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m:
        out[cur].append(m.group(1).strip())
dem = subprocess.run(["c++filt"] + list(out.keys()), capture_output=True, text=True).stdout.split("\n")
KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMAPF", "UBLKCP", "STTM", "LDTM", "UTCATOM", "USETMAXREG", "SYNCS", "HFMA2", "FFMA2", "FMUL2",
        "FFMA", "LDS", "LDGSTS", "STG", "LDG", "ELECT", "NANOSLEEP"]
print("SASS summary of %s (cuobjdump -sass, sm_100a; tools/sass_summary.py)" % so)
print("UTCHMMA = tcgen05.mma (first source operand tmem[..] = TS form, A from tensor memory; gdesc[..] = SS form), UTCBAR = tcgen05.commit,")
print("UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, STTM / LDTM = tcgen05.st / tcgen05.ld, USETMAXREG = setmaxnreg, SYNCS = mbarrier ops\n")
for (k, v), d in zip(out.items(), dem):
    c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0].split(".")[0] for ins in v)
    print("%s\n    %d instructions; %s" % (d[:120], len(v), ", ".join("%s %d" % (kk, c[kk]) for kk in KEYS if c[kk])))
    ex = ([i for i in v if "UTCHMMA" in i][:2] + [i for i in v if "UTMALDG" in i][:1] + [i for i in v if "STTM" in i][:1] +
          [i for i in v if "USETMAXREG" in i][:3])
    for e in ex:
        print("      e.g. " + e)

"""Static tcgen05 / TMEM / TMA instruction counts per kernel instantiation of the built engine (runs anywhere: cuobjdump only).
   python tools/sass_counts.py [build/tc_engine.o]   -> the markdown table of profiles/r2_sass_tcgen05.md"""
import collections
import re
import subprocess
import sys

obj = sys.argv[1] if len(sys.argv) > 1 else "build/tc_engine.o"
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
ops = ["UTCHMMA", "STTM", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR"]
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        k = re.search(r"(tc_ts_kernel|tc_igemm_kernel)<deft::(\w+), (\d+), (\d+)", name)
        cur = f"{k.group(1)}<{k.group(2)},{k.group(3)},{k.group(4)}>" if k else None
        if cur:
            counts[cur] = collections.Counter()
        continue
    if cur:
        for o in ops:
            if re.search(r"\b" + o + r"\b|\b" + o + r"\.", line):
                counts[cur][o] += 1
print("| kernel | " + " | ".join(ops) + " |\n|---|" + "---:|" * len(ops))
tot = collections.Counter()
for k, c in counts.items():
    print(f"| `{k}` | " + " | ".join(str(c[o]) for o in ops) + " |")
    tot.update(c)
print("| **total** | " + " | ".join(str(tot[o]) for o in ops) + " |")

"""Instruction mix per device function of a built .so (no GPU needed): python profiles/sass_mix.py <lib.so> [filter]"""
import collections, os, re, subprocess, sys, tempfile
lib = os.path.abspath(sys.argv[1]); flt = sys.argv[2] if len(sys.argv) > 2 else ""
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
cur, cnt = None, collections.defaultdict(collections.Counter)
for l in dis.splitlines():
    m = re.match(r"(\$?[_A-Za-z][^\s:]*):\s*$", l)
    if m and not m.group(1).startswith(".L"):
        cur = m.group(1).split("$")[-1] if not m.group(1).startswith("$__internal") else cur; continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and cur:
        cnt[cur][m.group(1).split(".")[0]] += 1
for f, c in cnt.items():
    if flt in f:
        print(f"{f[:70]:70s} {sum(c.values()):6d}", dict(c.most_common(9)))

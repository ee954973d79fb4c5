"""Aggregate an ncu source-page CSV (SASS level) by CUDA source line / function using nvdisasm -g line info.

usage: python profiles/ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-section-substring> [top]
Needs: ncu, cuobjdump, nvdisasm on PATH (no GPU).  The report must come from a build with -lineinfo of the same .so.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kname = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
sec, line, seq = None, None, []
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        sec = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and sec and kname in sec:
        seq.append((line, m.group(2).strip()))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = rows[hi + 1:]
assert len(data) == len(seq), (len(data), len(seq))
# function table from the sources
funcs = {}
srcdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "paper_romualdi_2022_icra_centroidal-mpc-walking_b200", "csrc")
for fn in os.listdir(srcdir):
    starts = []
    for i, l in enumerate(open(os.path.join(srcdir, fn)), 1):
        m = re.match(r"(?:CMPC_FN|CMPC_HD|__global__|static|template.*\)\s*$)?\s*(?:CMPC_FN|CMPC_HD)\s+[\w:<>]+\s+(\w+)\(", l)
        if m:
            starts.append((i, m.group(1)))
        m = re.match(r"(\w+)\(Config cfg", l)
        if m:
            starts.append((i, m.group(1)))
    funcs[fn] = starts


def func_of(file, ln):
    best = "?"
    for s, name in funcs.get(file, []):
        if s <= ln:
            best = name
    return best


STALLS = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_selected", "stall_branch_resolving", "stall_no_inst", "stall_math", "stall_mio", "stall_lg", "stall_not_selected"]
sidx = [hdr.index(s_) if s_ in hdr else None for s_ in STALLS]
stall_line = collections.defaultdict(lambda: [0] * len(STALLS))
wf_line = collections.Counter()
iwf = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
by_line, by_func = collections.Counter(), collections.Counter()
samp_line, samp_func = collections.Counter(), collections.Counter()
tot_i = tot_s = 0
for (ln, txt), r in zip(seq, data):
    n = int(float(r[ci] or 0)); s = int(float(r[cs] or 0))
    key = ln or ("?", 0)
    by_line[key] += n; samp_line[key] += s
    if iwf is not None and r[iwf]:
        wf_line[key] += int(float(r[iwf]))
    for q, ix in enumerate(sidx):
        if ix is not None and r[ix]:
            stall_line[key][q] += int(float(r[ix]))
    f = func_of(*key)
    by_func[f] += n; samp_func[f] += s
    tot_i += n; tot_s += s
print(f"total warp instructions {tot_i:.3e}, samples {tot_s}")
print("---- by function (instr %, samples %)")
for f, n in by_func.most_common(25):
    print(f"{f:22s} {100*n/tot_i:6.2f}% {100*samp_func[f]/max(tot_s,1):6.2f}%")
print("---- by line (samples %)")
for k, s in samp_line.most_common(top):
    print(f"{k[0]}:{k[1]:<5d} samples {100*s/max(tot_s,1):5.2f}%  instr {100*by_line[k]/tot_i:5.2f}%")
print("---- by line (instr %)")
for k, n in by_line.most_common(top):
    print(f"{k[0]}:{k[1]:<5d} instr {100*n/tot_i:5.2f}%  samples {100*samp_line[k]/max(tot_s,1):5.2f}%")
# optional phase table: CMPC_PHASES="file:lo-hi=name,..."
ph = os.environ.get("CMPC_PHASES")
if ph:
    print("---- by phase (instr %, samples %) | samples % by stall: " + " ".join(x.replace("stall_", "") for x in STALLS))
    for item in ph.split(","):
        rng, name = item.split("=")
        fn, lh = rng.split(":")
        lo, hi = map(int, lh.split("-"))
        n = sum(v for (f, l), v in by_line.items() if f == fn and lo <= l <= hi)
        s = sum(v for (f, l), v in samp_line.items() if f == fn and lo <= l <= hi)
        st = [sum(v[q] for (f, l), v in stall_line.items() if f == fn and lo <= l <= hi) for q in range(len(STALLS))]
        wf = sum(v for (f, l), v in wf_line.items() if f == fn and lo <= l <= hi)
        print(f"{name:28s} {100*n/tot_i:6.2f}% {100*s/max(tot_s,1):6.2f}% smem-wf {100*wf/max(sum(wf_line.values()),1):5.1f}%  | " + " ".join(f"{100*x/max(tot_s,1):5.2f}" for x in st))

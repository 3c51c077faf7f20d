"""Per-kernel SASS digest of libseqj_b200.so (cuobjdump -sass): instruction counts that identify the hardware paths
each kernel uses.  python tools/sass_digest.py > profiles/r02_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sequencerj.jl_b200", "libseqj_b200.so")
MNEMONICS = ["UTMALDG", "UTMAPF", "SYNCS", "DMMA", "DADD", "DFMA", "DMUL", "CREDUX", "REDUX", "LDGSTS", "LDG", "STG", "LDS", "STS", "ATOM",
             "RED", "SHFL", "VOTE", "BAR", "MUFU", "UCGABAR", "ACQBULK", "ST.E", "UTC", "HMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    counts = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("seqj::", "")
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1)
            counts[cur]["total"] += 1
            for mn in MNEMONICS:
                if op.startswith(mn):
                    counts[cur][mn] += 1
    print(f"# SASS digest of {os.path.relpath(LIB, ROOT)}: cubin architectures {arch}")
    print("# TMA = UTMALDG (cp.async.bulk.tensor); FP64 tensor core = DMMA (tcgen05 has no f64 kind); LDGSTS = cp.async;")
    print("# CREDUX.MIN = warp-wide integer min (__reduce_min_sync), REDUX = warp-wide sum / or; SYNCS = mbarrier; UTC*/HMMA counts are expected to be 0 for this FP64 path")
    cols = ["total", "UTMALDG", "SYNCS", "DMMA", "DADD", "DFMA", "DMUL", "CREDUX", "REDUX", "LDGSTS", "LDG", "STG", "LDS", "STS", "ATOM", "RED",
            "SHFL", "MUFU", "BAR", "UTC", "HMMA"]
    print(f"{'kernel':58s}" + "".join(f"{c:>8s}" for c in cols))
    for k, c in counts.items():
        print(f"{k[:58]:58s}" + "".join(f"{c.get(col, 0):8d}" for col in cols))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print(f"{'ALL KERNELS':58s}" + "".join(f"{tot.get(col, 0):8d}" for col in cols))


if __name__ == "__main__":
    main()

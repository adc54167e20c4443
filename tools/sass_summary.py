"""Per-kernel SASS opcode summary of libprfl_b200.so (cuobjdump -sass): counts of the Blackwell-native instructions that
prove which hardware path a kernel uses — UTCHMMA (tcgen05.mma; `.2CTA` = cta_group::2), LDTM / STTM (tcgen05.ld / st),
UTMALDG (TMA tensor load), UBLKCP (TMA bulk copy), UTCBAR (tcgen05.commit), SYNCS (mbarrier) — next to the legacy
HMMA.16816 (mma.sync) count, which must be zero.  Usage: python tools/sass_summary.py > profiles/rNN_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hy-video-prfl_b200", "libprfl_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU.EX2", "FFMA2", "LDG", "STG", "LDS", "STS", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_total"] += 1
        base = op.split(".")[0]
        for o in OPS:
            if "." in o:
                if op.startswith(o.split(".")[0]) and o.split(".", 1)[1] in op:
                    kernels[cur][o] += 1
            elif base == o:
                kernels[cur][o] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, demangle)) if len(demangle) == len(kernels) else {k: k for k in kernels}
    print("# SASS opcode summary of `hy-video-prfl_b200/libprfl_b200.so` (sm_100a)\n")
    print("`python tools/sass_summary.py` (cuobjdump -sass, CUDA 12.9).  UTCHMMA = tcgen05.mma (`.2CTA` subset = cta_group::2), LDTM / STTM = "
          "tcgen05.ld / st, UTMALDG = TMA tensor load, UBLKCP = TMA bulk copy, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, "
          "HMMA = legacy mma.sync (must be 0).\n")
    cols = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU.EX2", "FFMA2", "_total"]
    print("| kernel | " + " | ".join(c.replace("_total", "instructions") for c in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    tot = collections.Counter()
    for k, c in kernels.items():
        m = re.match(r"^(?:void )?(?:prfl::)?(\w+(?:<.*>)?)\(", names[k])
        nm = (m.group(1) if m else names[k]).replace("prfl::", "").replace("(bool)", "").replace("(int)", "")
        print(f"| `{nm}` | " + " | ".join(str(c.get(x, 0)) for x in cols) + " |")
        tot.update(c)
    print("| **all kernels** | " + " | ".join(f"**{tot.get(x, 0)}**" for x in cols) + " |")
    if tot.get("HMMA", 0):
        print("\nWARNING: legacy HMMA instructions present", file=sys.stderr)


if __name__ == "__main__":
    main()

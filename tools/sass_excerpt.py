"""profiles/r02_sass_excerpt.txt: Blackwell-specific (and legacy tensor-core) SASS mnemonics per kernel of libdmv3d.so.

    python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "dynamic_multiview_3d_b200", "libdmv3d.so")
PAT = re.compile(r"\b(UTC[A-Z]+|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|SYNCS|LDGMC|STGMC|REDG?MC|HMMA|LDSM|LDGSTS|MEMBAR\.ALL\.SYS)\b")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for k in PAT.findall(line.split("/*")[1] if "/*" in line and line.strip().startswith("/*") else ""):
            funcs[cur][k] += 1
    names = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass dynamic_multiview_3d_b200/libdmv3d.so (sm_100a): tensor-core / TMA / TMEM / multimem mnemonics per kernel")
    print("# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG = TMA tensor load/store (cp.async.bulk.tensor),")
    print("# SYNCS = mbarrier ops, LDGSTS = cp.async, LDSM = ldmatrix, LDGMC = multimem.ld_reduce (in-switch reduction), MEMBAR.ALL.SYS = fence.sys of the")
    print("# peer exchange.  HMMA (warp-level mma.sync) appears ONLY in the HBM-bound kernels that use it on purpose: the thin layers")
    print("# (thin_mma.cu: 3 / 2 image-side channels) and the fused FC weight-gradient + Adam kernel (fc_adam.cu: 64-deep contraction feeding a")
    print("# 26 B/parameter stream); every convolution / deconvolution / linear contraction of the graph is tcgen05.")
    print()
    tot = collections.Counter()
    for (mangled, cnt), name in zip(funcs.items(), names):
        if not cnt:
            continue
        tot.update(cnt)
        short = re.sub(r"\(anonymous namespace\)::|<unnamed>::|dmv::", "", name).replace("void ", "")
        short = re.sub(r"_GLOBAL__N__\w+::", "", short).split("(")[0]
        print("%-70s %s" % (short[:70], "  ".join("%s=%d" % kv for kv in sorted(cnt.items()))))
    print()
    print("# totals: " + "  ".join("%s=%d" % kv for kv in sorted(tot.items())))


if __name__ == "__main__":
    main()

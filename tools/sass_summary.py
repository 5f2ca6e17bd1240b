"""Evidence helper: per-kernel summary of the tcgen05 / TMEM / TMA instructions in the shipped library's SASS.

    python tools/sass_summary.py > profiles/r02_sass_tcgen05_tma.txt

For every kernel of libmcm_b200.so (cuobjdump -sass): the number of SASS instructions, the counts of the mnemonics that
prove the Blackwell paths (B200_PROFILING.md: UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG / UTMASTG /
UTMAREDG = cp.async.bulk.tensor load / store / reduce, UBLKPF = bulk L2 prefetch, UTCBAR = tcgen05.commit, SYNCS = mbarrier)
and the first occurrence of each as it appears in the listing.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "motioncraft_b200", "libmcm_b200.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UTMAREDG[.\w]*|UBLKPF[.\w]*|UBLKCP[.\w]*|"
                 r"UTCBAR[.\w]*|UTCATOMSWS[.\w]*|SYNCS[.\w]*|FFMA2|MUFU[.\w]*|F2FP[.\w]*|REDG?[.\w]*|ATOMG?[.\w]*)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = dict(n=0, counts=collections.Counter(), first={})
            continue
        if cur is None or "/*" not in line:
            continue
        body = line.split("*/", 1)[-1]
        if not re.search(r"[A-Z]{2,}", body):
            continue
        kernels[cur]["n"] += 1
        for mm in PAT.finditer(body):
            op = mm.group(1)
            key = op.split(".")[0] if op.startswith(("SYNCS", "MUFU", "F2FP", "RED", "ATOM")) else op
            kernels[cur]["counts"][key] += 1
            kernels[cur]["first"].setdefault(key, body.strip().rstrip(";").strip())
    demangle = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, cuobjdump -sass (sm_100a)")
    for (name, k), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dn.replace("(anonymous namespace)::", ""))
        print(f"\n## {short}   [{k['n']} SASS instructions]")
        tc = [f"{op} x{n}" for op, n in sorted(k["counts"].items())]
        print("   " + (", ".join(tc) if tc else "(no tcgen05 / TMA / MUFU instructions)"))
        for op in sorted(k["first"]):
            if op.startswith(("UTC", "LDTM", "STTM", "UTMA", "UBLK")):
                print(f"     {op:28s} e.g.  {k['first'][op]}")


if __name__ == "__main__":
    sys.exit(main())

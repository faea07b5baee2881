"""Per-kernel counts of the SASS instructions that prove the Blackwell-native paths (B200_PROFILING.md "What proves a
Blackwell-native kernel"): UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA load /
store / reduce), UTCBAR (tcgen05.commit), and the legacy HMMA (mma.sync -- must be absent).

usage: python tools/sass_summary.py [path/to/libhvlm_b200.so] > profiles/sass_summary.txt"""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "handsonvlm-release_b200", "libhvlm_b200.so")
PATTERNS = OrderedDict([
    ("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"),
    ("UTMALDG", r"\bUTMALDG\b"), ("UTMASTG", r"\bUTMASTG\b"), ("UTMAREDG", r"\bUTMAREDG\b"), ("UTCBAR", r"\bUTCBAR\b"),
    ("MUFU", r"\bMUFU\b"), ("HMMA", r"\bHMMA\b")])


def demangle(names):
    try:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        return out if len(out) == len(names) else names
    except Exception:
        return names


def summarize(lib=LIB):
    """-> OrderedDict kernel (demangled, shortened) -> {mnemonic: count}"""
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True)
    if sass.returncode != 0:
        raise RuntimeError(sass.stderr[-500:])
    kernels, cur = OrderedDict(), None
    for line in sass.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {k: 0 for k in PATTERNS}
            continue
        if cur is None or "/*" not in line:
            continue
        for k, pat in PATTERNS.items():
            if re.search(pat, line):
                kernels[cur][k] += 1
    names = list(kernels)
    short = [re.sub(r"\(.*", "", d).replace("hvlm::", "").replace("void ", "") for d in demangle(names)]
    return OrderedDict((s, kernels[n]) for s, n in zip(short, names))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    ks = summarize(lib)
    cols = list(PATTERNS)
    print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} -- instruction counts per kernel (sm_100a)")
    print(f"# {'kernel':78s} " + " ".join(f"{c:>12s}" for c in cols))
    tot = {c: 0 for c in cols}
    for k, v in ks.items():
        if any(v[c] for c in cols if c != "MUFU"):
            print(f"{k[:80]:80s} " + " ".join(f"{v[c]:12d}" for c in cols))
        for c in cols:
            tot[c] += v[c]
    print(f"{'TOTAL (all ' + str(len(ks)) + ' kernels)':80s} " + " ".join(f"{tot[c]:12d}" for c in cols))


if __name__ == "__main__":
    main()

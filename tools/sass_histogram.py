"""Per-kernel SASS opcode histogram of libvd3d_b200.so (cuobjdump -sass): the evidence that the hot kernels are tcgen05 / TMA code.
usage: python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "visualdet3d_b200", "libvd3d_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "REDUX", "HMMA", "FFMA", "DFMA",
        "LDG", "STG", "LDS", "STS", "ATOMG", "RED", "SHFL", "MUFU", "BAR"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, hist = None, collections.OrderedDict()
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            op = m.group(1).split(".")[0]
            hist[cur][op] += 1
            hist[cur]["_total"] += 1
    dm = demangle(list(hist))
    short = lambda n: re.sub(r"\(.*", "", dm[n]).replace("vd3d::", "")
    print("# cuobjdump -sass visualdet3d_b200/libvd3d_b200.so (sm_100a): instruction counts per kernel; tcgen05 = UTCHMMA (mma) / LDTM (tcgen05.ld) /")
    print("# UTCBAR (tcgen05.commit); TMA = UTMALDG (cp.async.bulk.tensor) / UBLKCP (cp.async.bulk); SYNCS = mbarrier; REDUX = uniform-register broadcast")
    cols = [k for k in KEYS if any(h[k] for h in hist.values())]
    print(f"{'kernel':58s} {'total':>7s} " + " ".join(f"{c:>8s}" for c in cols))
    for n, h in sorted(hist.items(), key=lambda kv: -(kv[1]["UTCHMMA"] * 1000000 + kv[1]["_total"])):
        print(f"{short(n)[:58]:58s} {h['_total']:7d} " + " ".join(f"{h[c]:8d}" for c in cols))
    tc = [short(n) for n, h in hist.items() if h["UTCHMMA"]]
    tma = [short(n) for n, h in hist.items() if h["UTMALDG"] or h["UBLKCP"]]
    print(f"\n# kernels with tcgen05.mma (UTCHMMA): {len(tc)};  kernels with TMA (UTMALDG / UBLKCP): {len(tma)};  kernels in the library: {len(hist)}")


if __name__ == "__main__":
    main()

"""SASS evidence for the built library: per kernel, the Blackwell-specific mnemonics it contains (with counts) and one
example line of each of the telling ones.   python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multi-level-indoor-slam_b200", "semgate", "libsemgate.so")
KEEP = re.compile(r"^(UTCHMMA|UTMALDG|UTMAPF|UBLKCP|LDTM|UTCBAR|UTCATOMSWS|SYNCS|UCGABAR|REDUX|ELECT|MEMBAR|DADD|DSETP|HMMA|ATOMG|ATOMS|REDS|RED\b|SHFL)")
SHOW = ("LDTM", "UTCBAR", "UTCHMMA", "UTMALDG", "UBLKCP")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    dem = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    names = iter(dem)
    print("SASS evidence for libsemgate.so (sm_100a), round 2 (final code)")
    print("command: cuobjdump -sass multi-level-indoor-slam_b200/semgate/libsemgate.so   (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a); tools/sass_evidence.py")
    print("mnemonics: UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = TMA tensor load (cp.async.bulk.tensor; .MULTICAST = multicast::cluster),")
    print("           UBLKCP = cp.async.bulk (non-tensor TMA bulk copy global -> shared, mbarrier completion; K3's dense-list kernel),")
    print("           LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit (mbarrier arrive), SYNCS = mbarrier ops, UCGABAR = barrier.cluster,")
    print("           REDUX = warp reduce, SHFL = warp shuffle, HMMA = legacy mma.sync (none expected)\n")
    cur, counts, ex = None, None, None

    def flush():
        if cur is None:
            return
        if not any(k.startswith(("UTCHMMA", "UTMALDG", "LDTM", "UBLKCP")) for k in counts):
            print(f"== {cur}\n   (no tensor-core / TMA instructions) " + ", ".join(f"{k} x{v}" for k, v in sorted(counts.items()) if k.startswith(("REDUX", "SHFL", "ATOM", "RED", "DADD"))))
            return
        print(f"== {cur}")
        print("   " + ", ".join(f"{k} x{v}" for k, v in sorted(counts.items())))
        for k in SHOW:
            for m, line in ex.items():
                if m.startswith(k):
                    print("      e.g. " + line.strip())
                    break

    for line in sass.split("\n"):
        if "Function :" in line:
            flush()
            cur = next(names).replace("(anonymous namespace)::", "").replace("(bool)", "").replace("(int)", "").split("(")[0]
            counts, ex = collections.Counter(), {}
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None and KEEP.match(m.group(1)):
            counts[m.group(1)] += 1
            ex.setdefault(m.group(1), line)
    flush()


if __name__ == "__main__":
    main()

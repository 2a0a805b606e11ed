"""Per-kernel SASS mnemonic counts of the shipped library -> profiles/r2_sass_evidence.txt.

    python tools/sass_evidence.py            # needs cuobjdump + c++filt (CUDA toolkit), no GPU

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, HMMA = legacy
mma.sync, LDSM = ldmatrix, LDGSTS = cp.async, SYNCS = mbarrier ops, STAS = st.async (DSMEM store + complete_tx),
UCGABAR = barrier.cluster, MEMBAR = fence."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorch-kaldi_b200", "libpk_b200.so")
KEYS = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UBLKCP", "HMMA", "LDSM", "LDGSTS", "SYNCS",
        "STAS", "UCGABAR", "MEMBAR")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    counts[cur][k] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True, check=True).stdout.splitlines()
    rows = []
    for mangled, name in zip(counts, names):
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"^void ", "", name.split("(")[0])
        c = counts[mangled]
        if c:
            rows.append((name, " ".join(f"{k}={c[k]}" for k in KEYS if c[k])))
    path = os.path.join(ROOT, "profiles", "r2_sass_evidence.txt")
    with open(path, "w") as f:
        f.write("# SASS evidence (cuobjdump -sass pytorch-kaldi_b200/libpk_b200.so, mnemonic counts per kernel; tools/sass_evidence.py).\n"
                "# B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk,\n"
                "# HMMA = legacy mma.sync, LDSM = ldmatrix, LDGSTS = cp.async, SYNCS = mbarrier ops, STAS = st.async (DSMEM store +\n"
                "# complete_tx), UCGABAR = barrier.cluster (only at kernel start / exit in the persistent kernels), MEMBAR = fence\n")
        for name, c in sorted(rows):
            f.write(f"{name:84s} {c}\n")
    print(f"wrote {path}: {len(rows)} kernels")


if __name__ == "__main__":
    sys.exit(main())

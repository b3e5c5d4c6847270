#!/usr/bin/env python
"""Per-kernel SASS instruction histogram of libdce_b200.so (no GPU needed):  python tools/sass_histogram.py [out.txt]
Counts the mnemonics that prove what a kernel is built from (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UBLKCP = 1-D bulk TMA (cp.async.bulk), UTMALDG = tensor-map TMA, SYNCS = mbarrier ops, HMMA = the
legacy mma.sync path (none expected), FFMA = fp32 CUDA-core arithmetic."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "deep_contact_estimator_b200", "libdce_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UBLKPF", "UTMALDG", "SYNCS", "ELECT", "HMMA", "FFMA", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
    out, k = [], -1
    hist = []
    for line in sass.splitlines():
        if "Function :" in line:
            k += 1
            hist.append(collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and k >= 0:
            op = m.group(1)
            hist[k]["total"] += 1
            for key in KEYS:
                if op.startswith(key):
                    hist[k][key] += 1
                    break
    arch = re.search(r"arch = (sm_\w+)", sass)
    out.append(f"libdce_b200.so: {len(hist)} kernels, {arch.group(1) if arch else '?'}; columns: " + " ".join(KEYS) + " | total instructions")
    tot = collections.Counter()
    for name, h in sorted(zip(names, hist), key=lambda t: -t[1]["total"]):
        short = name[:name.rfind("(")] if "(" in name else name
        short = short.replace("void ", "").replace("dce::", "").replace("(int)", "").replace("(bool)", "")
        out.append(f"{short[:78]:78s} " + " ".join(f"{h[key]:5d}" for key in KEYS) + f" | {h['total']:6d}")
        tot.update(h)
    out.append(f"{'ALL':78s} " + " ".join(f"{tot[key]:5d}" for key in KEYS) + f" | {tot['total']:6d}")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()

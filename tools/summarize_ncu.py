#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.
   python tools/summarize_ncu.py <launches.csv> <full.ncu-rep> <out.txt>"""
import csv, re, subprocess, sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
           "smsp__inst_executed.sum"]


def short(name):
    m = re.search(r"tapgemm_kernel<\(int\)(\d+), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+)>", name) or \
        re.search(r"tapgemm_kernel<(\d+), (\d+), (\d+), (\d+), (\d+)>", name)
    if m:
        return "tapgemm<BN=%s,TAPS=%s,KSA=%s,NSTAGE=%s,EPI=%s>" % m.groups()
    name = name[:name.rfind("(")] if "(" in name else name
    return name.replace("void ", "").replace("dce::", "").replace("(int)", "").replace("(bool)", "")


def main():
    launches, rep, out = sys.argv[1], sys.argv[2], sys.argv[3]
    lines = []
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else v
        agg.setdefault(short(r[ki]), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    lines.append(f"== launch list ({launches}): gpu__time_duration per kernel, ncu-serialised, cold cache: compare SHARES ==")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"{k:52s} n={len(v):3d}  avg {sum(v) / len(v):8.1f} us  share {sum(v) / tot * 100:5.1f} %")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(txt.splitlines()))
    h = rr[0]
    lines.append("")
    lines.append(f"== ncu --set full ({rep}) ==")
    seen = set()
    for r in rr[2:]:
        name = short(r[h.index("Kernel Name")])
        if name in seen:
            continue
        seen.add(name)
        lines.append(name)
        for m in METRICS:
            if m in h:
                lines.append(f"    {m:70s} {r[h.index(m)]} {rr[1][h.index(m)]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()

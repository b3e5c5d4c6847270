#!/usr/bin/env python
"""DRAM traffic of the batch kernels as a function of the chunk size, natural cache state.  Run under
    ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file out.csv python tools/traffic_by_batch.py 1024 2048 4096
(two metrics = one pass, so nothing is replayed and the L2 keeps what the previous kernel left); summarise with --summarise out.csv."""
import collections
import csv
import os
import sys

if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
    rows = list(csv.reader(open(sys.argv[2])))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, mi, vi, gi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Grid Size")
    acc = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        acc.setdefault((r[ki][:44], r[gi]), collections.defaultdict(list))[r[mi]].append(float(r[vi].replace(",", "")) / 1e6)
    for k, v in acc.items():
        print(k, {m: [round(x, 1) for x in xs[-3:]] for m, xs in v.items()})
    sys.exit(0)

import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce          # noqa: E402
from deep_contact_estimator_b200 import synth      # noqa: E402

dev = torch.device("cuda", 0)
eng = dce.ContactEngine(synth.make_params(0), dev, "bf16x3")
for B in [int(a) for a in sys.argv[1:]] or [4096]:
    xs = [synth.make_windows(B, seed=5 + i).to(dev) for i in range(3)]
    for i in range(6):
        eng.classify(xs[i % 3])
    torch.cuda.synchronize()

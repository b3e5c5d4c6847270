#!/usr/bin/env python
"""clock64 timeline of CTA 0 of one tapgemm layer: python tools/trace_tapgemm.py <layer 2..5>"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce
from deep_contact_estimator_b200 import synth
layer = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
eng = dce.ContactEngine(synth.make_params(0), dev, "bf16x3")
x = synth.make_windows(4096, seed=1).to(dev)
for _ in range(3): eng.classify(x)
eng.set_option(b"trace", 1); eng.set_option(b"trace_layer", layer)
eng.classify(x); torch.cuda.synchronize()
buf = eng.read_trace(60 * 16)
t = np.array(buf, dtype=np.int64).reshape(60, 16)
t0 = t[t > 0].min()
names = ["ld:first", "ld:last", "mma:start", "mma:tempty", "mma:full0", "mma:fullL", "ep:start", "ep:tfull", "ep:done", "s1:pre", "s1:post", "s2:pre", "s2:post", "s3:pre", "s3:post"]
print("tile " + " ".join(f"{n:>10s}" for n in names))
for k in range(0, 10):
    print(f"{k:4d} " + " ".join(f"{(t[k, e] - t0) if t[k, e] else 0:10d}" for e in range(15)))
k1 = max(k for k in range(60) if t[k, 2] > 0)
print("tiles", k1 + 1, "SM clock during kernel: %.3f GHz" % ((t[k1, 2] - t[1, 2]) / (t[k1, 15] - t[1, 15])), " tile period cycles %.0f, ns %.0f" % ((t[k1, 2] - t[1, 2]) / (k1 - 1), (t[k1, 15] - t[1, 15]) / (k1 - 1)))

# per-CTA wall clock (globaltimer, ns): kernel entry, after griddepcontrol.wait, exit — CTAs 0..59
ent, dep, ext = t[:, 12], t[:, 13], t[:, 14]
ok = ext > 0
if ok.any():
    z = ent[ok].min()
    print("CTAs traced:", int(ok.sum()), " entry spread %.1f us, dependency wait ends %.1f .. %.1f us after the first entry, exits %.1f .. %.1f us, CTA 0 exit %.1f us"
          % ((ent[ok].max() - z) / 1e3, (dep[ok].min() - z) / 1e3, (dep[ok].max() - z) / 1e3, (ext[ok].min() - z) / 1e3, (ext[ok].max() - z) / 1e3, (ext[0] - z) / 1e3))

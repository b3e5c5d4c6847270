import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce
from deep_contact_estimator_b200 import synth
dev = torch.device("cuda", 0)
eng = dce.ContactEngine(synth.make_params(0), dev, "bf16x3")
xs = [synth.make_windows(4096, seed=10 + i).to(dev) for i in range(3)]
ref = None
for fuse in (0, 1):
    eng.set_option(b"fuse_block2", fuse)
    for i in range(3): out = eng.classify(xs[0])
    torch.cuda.synchronize()
    if ref is None: ref = [t.clone() for t in out]
    else: print("max |dlogit| fused vs layerwise:", (out[0] - ref[0]).abs().max().item(), "classes equal:", bool(torch.equal(out[1], ref[1])))
    tot = {}
    for i in range(20):
        for name, ms in eng.profile_forward(xs[i % 3]): tot[name] = tot.get(name, 0) + ms / 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50): eng.classify(xs[i % 3])
    e1.record(); torch.cuda.synchronize()
    print(f"fuse_block2={fuse}  step {e0.elapsed_time(e1) / 50 * 1e3:.1f} us  " + "  ".join(f"{k} {v*1e3:.1f}" for k, v in tot.items()))

import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce
from deep_contact_estimator_b200 import synth
dev = torch.device("cuda", 0)
eng = dce.ContactEngine(synth.make_params(0), dev, "bf16x3")
for B in [int(a) for a in sys.argv[1:]] or [512, 997, 1024, 1994, 2048, 4096]:
    xs = [synth.make_windows(B, seed=5 + i).to(dev) for i in range(2)]
    for i in range(3): eng.classify(xs[i % 2])
    tot = {}
    for i in range(10):
        for n, ms in eng.profile_forward(xs[i % 2]): tot[n] = tot.get(n, 0) + ms / 10
    print(B, {k: round(v * 1e3, 1) for k, v in tot.items()}, "sum", round(sum(tot.values()) * 1e3, 1), "per4096", round(sum(tot.values()) * 1e3 * 4096 / B, 1))

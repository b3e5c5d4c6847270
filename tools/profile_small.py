import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce
from deep_contact_estimator_b200 import synth
dev = torch.device("cuda", 0)
eng = dce.ContactEngine(synth.make_params(0), dev, "bf16x3")
for B in (1, 4):
    x = synth.make_windows(B, seed=5).to(dev)
    for _ in range(3): eng.classify(x)
    tot = {}
    for _ in range(20):
        for n, ms in eng.profile_forward(x): tot[n] = tot.get(n, 0) + ms / 20
    print(B, {k: round(v * 1e3, 1) for k, v in tot.items()}, "sum us", round(sum(tot.values()) * 1e3, 1))
log = synth.make_sensor_log(4096 + 149, seed=2).to(dev)
for _ in range(3): eng.stream(log)
tot = {}
for _ in range(10):
    for n, ms in eng.profile_stream(log, 0, 4096): tot[n] = tot.get(n, 0) + ms / 10
print("stream4096", {k: round(v * 1e3, 1) for k, v in tot.items()}, "sum us", round(sum(tot.values()) * 1e3, 1))

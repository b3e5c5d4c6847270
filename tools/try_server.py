#!/usr/bin/env python
"""First contact with the resident latency server (LatencyRunner(persistent=True)): parity with the launch-per-step
runner on changing windows, host latency percentiles, idle retirement + restart.  Run under a timeout:
    timeout 120 python tools/try_server.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce
from deep_contact_estimator_b200 import synth

dev = torch.device("cuda", 0)
eng = dce.ContactEngine(synth.make_params(0), dev, "bf16x3")
xs = synth.make_windows(64, seed=3)
ref = eng.latency_runner(1, want_logits=True)
want = []
for i in range(64):
    ref.step(xs[i]); want.append((int(ref.cls_host[0]), ref.bits_host[0].tolist(), ref.logits_host.clone()))
t = []
for _ in range(300):
    t0 = time.perf_counter(); ref.step(); t.append((time.perf_counter() - t0) * 1e6)
print("launch-per-step runner: host p50 %.1f us p99 %.1f us" % (np.percentile(t, 50), np.percentile(t, 99)), flush=True)

run = eng.latency_runner(1, want_logits=True, persistent=True, idle_timeout_s=0.5)
print("server up, starts:", run.server_starts, flush=True)
ok = True
for rep in range(3):
    for i in range(64):
        run.step(xs[i])
        same = int(run.cls_host[0]) == want[i][0] and run.bits_host[0].tolist() == want[i][1] and torch.equal(run.logits_host, want[i][2])
        ok &= same
        if not same and rep == 0:
            print("mismatch at", i, int(run.cls_host[0]), want[i][0], (run.logits_host - want[i][2]).abs().max().item(), flush=True)
print("bit-identical to the launch-per-step kernel over 192 steps:", ok, flush=True)
for label, pace in (("back to back", 0.0), ("paced at 1 kHz", 1e-3)):
    t, d = [], []
    t_next = time.perf_counter()
    for i in range(2000 if pace == 0 else 500):
        if pace:
            t_next += pace
            while time.perf_counter() < t_next:
                pass
        t0 = time.perf_counter(); run.step(); t.append((time.perf_counter() - t0) * 1e6); d.append(int(run._c[run._DEVICE_NS]) / 1e3)
    print("persistent %s: host p50 %.1f us p90 %.1f p99 %.1f max %.1f | device (doorbell seen -> results written) p50 %.1f us p99 %.1f" % (label, np.percentile(t, 50), np.percentile(t, 90), np.percentile(t, 99), max(t), np.percentile(d, 50), np.percentile(d, 99)), flush=True)
t = []
for i in range(500):
    t0 = time.perf_counter(); run.step(xs[i % 64]); t.append((time.perf_counter() - t0) * 1e6)
print("persistent incl. the 32 KB window copy: host p50 %.1f us p99 %.1f" % (np.percentile(t, 50), np.percentile(t, 99)), flush=True)
run.close()
fast = eng.latency_runner(1, persistent=True, idle_timeout_s=0.5)          # class + bits only: results arrive with the step number in one store
okf = True
for i in range(64):
    fast.step(xs[i])
    okf &= int(fast.cls_host[0]) == want[i][0] and fast.bits_host[0].tolist() == want[i][1]
t, d = [], []
for i in range(2000):
    t0 = time.perf_counter(); fast.step(); t.append((time.perf_counter() - t0) * 1e6); d.append(int(fast._c[fast._DEVICE_NS]) / 1e3)
print("persistent, results in the control block (one 16-byte store): correct %s, host p50 %.1f us p99 %.1f | device p50 %.1f us" % (okf, np.percentile(t, 50), np.percentile(t, 99), np.percentile(d, 50)), flush=True)
xn = [xs[i % 64].numpy() for i in range(64)]
xv = fast.x_host.numpy()
t = []
for i in range(1000):
    t0 = time.perf_counter(); np.copyto(xv[0], xn[i % 64]); fast.step(); t.append((time.perf_counter() - t0) * 1e6)
print("  ... incl. a fresh 32 KB window (numpy copy): host p50 %.1f us p99 %.1f" % (np.percentile(t, 50), np.percentile(t, 99)), flush=True)
fast.close()
run = eng.latency_runner(1, want_logits=True, persistent=True, idle_timeout_s=0.5)
run.step(xs[1])
time.sleep(1.0)
print("after 1 s idle: alive =", int(run._c[run._ALIVE]), flush=True)
# the GPU is free again: an ordinary call goes through, then the next step restarts the server
lo, cl, bi = eng.classify(xs[:8].to(dev)); torch.cuda.synchronize()
run.step(xs[5])
print("restarted:", run.server_starts, "result ok:", int(run.cls_host[0]) == want[5][0], flush=True)
run.close()
print("closed: alive =", int(run._c[run._ALIVE]), flush=True)

# ---- row server: one new 54-float row per step, ring + z-score on the device ----
from oracle import contact_oracle as oracle
log = synth.make_sensor_log(150 + 400, seed=9)
_, wc, wb = oracle.inference_stream(synth.make_params(0), log)
rr = eng.row_runner(idle_timeout_s=0.5)
rows = log.numpy()
got, t, d = [], [], []
for i in range(rows.shape[0]):
    t0 = time.perf_counter(); out = rr.push(rows[i]); t.append((time.perf_counter() - t0) * 1e6); d.append(rr.device_us)
    if rr.ready:
        got.append(out)
okr = [g[0] for g in got] == wc.tolist() and [list(g[1]) for g in got] == wb.tolist()
print("row server: %d windows, classes + bits equal to the reference loop: %s; host p50 %.1f us p99 %.1f | device p50 %.1f us" %
      (len(got), okr, np.percentile(t[150:], 50), np.percentile(t[150:], 99), np.percentile(d[150:], 50)), flush=True)
rr.close()
# clock64 timeline of the last step (first and last CTA; written behind the barrier counters of the workspace header)
tr = rr._ws[64:64 + 24 * 8].view(torch.int64).cpu().numpy().reshape(2, 12)
NAMES = ["A done", "bar1", "B done", "bar2", "C done", "bar3", "D done", "E done (last CTA only)", "exit"]
for which, row in zip(("cta 0", "last cta"), tr):
    d = (row[1:10] - row[10]) / 1.965e3
    print(f"row server, {which}: us since 'row seen' (at 1.965 GHz): " + "  ".join(f"{n} {v:.2f}" for n, v in zip(NAMES, d)), flush=True)
time.sleep(0.3)
out = rr.push(rows[0])                       # restarted server, ring kept
print("row server after close + push: starts", rr.server_starts, flush=True)
rr.close()

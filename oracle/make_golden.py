#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE in the build container.

    python oracle/make_golden.py            # needs /root/reference (read-only)

/root/reference does not exist on the GPU box, so the fixtures are committed.
For every fixture this script also checks that ``oracle/contact_oracle.py``
reproduces the reference bit-for-bit here (same torch build, same ATen CPU
kernels) and prints the fp64 arbiter's distance; it aborts if they disagree.

Fixtures (small by design; weights are regenerated from seeds, never stored):
  forward_seed0.npz     logits of 64 z-scored windows, params seed 0
  forward_scaled.npz    same with the last layer x50 (logits O(1))
  stream_seed2.npz      reference contact_dataset + inference() over a 420-step
                        log: normalised-window checksum, logits, classes, bits
  bits_table.npz        decimal2binary truth table 0..15
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DCE_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, REF)

from deep_contact_estimator_b200 import synth            # noqa: E402
from oracle import contact_oracle as oracle              # noqa: E402

# the reference model + dataset import cleanly; inference_one_seq needs a stub lcm
from contact_cnn import contact_cnn as ref_contact_cnn   # noqa: E402  (reference)
from utils.data_handler import contact_dataset as ref_contact_dataset  # noqa: E402  (reference)

sys.modules.setdefault("lcm", types.ModuleType("lcm"))
import inference_one_seq as ref_infer                    # noqa: E402  (reference)


def ref_model(params):
    m = ref_contact_cnn()
    m.load_state_dict(params)
    return m.eval()


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    # ---- forward fixtures -------------------------------------------------
    for tag, seed, scale in (("forward_seed0", 0, 1.0), ("forward_scaled", 5, 50.0)):
        params = synth.make_params(seed, logit_scale=scale)
        x = synth.make_windows(64, seed=1)
        with torch.no_grad():
            want = ref_model(params)(x)
            got = oracle.forward_torch(params, x)
        assert torch.equal(want, got), f"{tag}: oracle differs from the reference"
        arb = oracle.forward_numpy64(params, x)
        print(f"{tag}: ref vs oracle bit-exact; ref vs fp64 arbiter normwise {oracle.normwise_rel_err(want.numpy(), arb):.2e}")
        np.savez(os.path.join(out_dir, tag + ".npz"),
                 logits=want.numpy(), logits_fp64=arb,
                 cls=torch.max(want, 1)[1].numpy().astype(np.int64),
                 param_seed=seed, logit_scale=scale, input_seed=1, batch=64)

    # ---- stream fixture: reference dataset + inference() loop --------------
    steps = 420
    log = synth.make_sensor_log(steps, seed=2)
    labels = synth.make_labels(steps, seed=3)
    params = synth.make_params(0)
    with tempfile.TemporaryDirectory() as d:
        dp, lp = os.path.join(d, "data.npy"), os.path.join(d, "label.npy")
        np.save(dp, log.double().numpy())            # on-disk format is float64 (utils/mat2numpy.py:73)
        np.save(lp, labels.numpy().reshape(-1, 1))   # (T,1) as utils/mat2numpy.py:76 writes it
        ds = ref_contact_dataset(dp, lp, window_size=150, device="cpu")
        assert len(ds) == oracle.num_windows(steps)
        loader = torch.utils.data.DataLoader(ds, batch_size=30)
        model = ref_model(params)
        bits = ref_infer.inference(loader, model, "cpu")
        wins = torch.stack([ds[i]["data"] for i in range(len(ds))])
        lab = torch.stack([ds[i]["label"] for i in range(len(ds))]).reshape(-1)
        with torch.no_grad():
            logits = torch.cat([model(wins[i:i + 30]) for i in range(0, len(ds), 30)])
    # f64 on disk -> f32 must round-trip to the synthetic log exactly
    o_wins = oracle.extract_windows(log, 0, len(ds))
    assert torch.equal(o_wins, wins), "oracle window extraction differs from contact_dataset"
    o_logits, o_cls, o_bits = oracle.inference_stream(params, log, batch_size=30)
    assert torch.equal(o_logits, logits) and torch.equal(o_bits, bits)
    assert torch.equal(lab, torch.stack([oracle.window_label(labels, i) for i in range(len(ds))]))
    print(f"stream_seed2: {len(ds)} windows, oracle == reference inference() bit-exact")
    np.savez(os.path.join(out_dir, "stream_seed2.npz"),
             logits=logits.numpy(), cls=o_cls.numpy(), bits=bits.numpy(), labels=lab.numpy(),
             win_sum=wins.double().sum(dim=(1, 2)).numpy(), win0=wins[0].numpy(), win_last=wins[-1].numpy(),
             steps=steps, log_seed=2, label_seed=3, param_seed=0)

    # ---- decimal2binary truth table ---------------------------------------
    tbl = ref_infer.decimal2binary(torch.arange(16))
    assert torch.equal(tbl, oracle.decimal2binary(torch.arange(16)))
    assert np.array_equal(tbl.numpy(), oracle.decimal2binary_numpy(np.arange(16)))
    np.savez(os.path.join(out_dir, "bits_table.npz"), table=tbl.numpy())
    print("bits_table: ok")

    # ---- LCM wire bytes from the reference's generated Python types (lcm_types/python) ----------
    from lcm_types.python import contact_t, leg_control_data_lcmt, microstrain_lcmt   # (reference)
    rng = np.random.RandomState(0)
    c = contact_t(); c.num_legs = 4; c.timestamp = 12.5; c.contact = [1, 0, 0, 1]
    leg = leg_control_data_lcmt()
    vals = {k: rng.randn(12).astype(np.float32) for k in ("q", "qd", "p", "v", "tau_est")}
    for k, v in vals.items():
        setattr(leg, k, v.tolist())
    imu = microstrain_lcmt()
    iv = {"quat": rng.randn(4).astype(np.float32), "rpy": rng.randn(3).astype(np.float32),
          "omega": rng.randn(3).astype(np.float32), "acc": rng.randn(3).astype(np.float32)}
    for k, v in iv.items():
        setattr(imu, k, v.tolist())
    imu.good_packets, imu.bad_packets = 7, 2
    np.savez(os.path.join(out_dir, "lcm_bytes.npz"),
             contact=np.frombuffer(c.encode(), dtype=np.uint8), leg=np.frombuffer(leg.encode(), dtype=np.uint8),
             imu=np.frombuffer(imu.encode(), dtype=np.uint8),
             **{"leg_" + k: v for k, v in vals.items()}, **{"imu_" + k: v for k, v in iv.items()})
    print("lcm_bytes: ok")


if __name__ == "__main__":
    main()

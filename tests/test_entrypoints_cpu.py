"""Entry-point scripts and the LCM wire shim, on CPU with synthetic files."""
import os

import numpy as np
import pytest
import scipy.io as sio
import torch
import yaml

from deep_contact_estimator_b200 import lcm_wire, synth
from deep_contact_estimator_b200.scripts import inference_one_seq as ios
from deep_contact_estimator_b200.scripts import test as test_script
from oracle import contact_oracle as oracle


def test_lcm_encoders_match_reference_generated_types(golden_dir):
    g = np.load(os.path.join(golden_dir, "lcm_bytes.npz"))
    assert lcm_wire.encode_contact(4, 12.5, [1, 0, 0, 1]) == g["contact"].tobytes()
    assert lcm_wire.encode_leg_control_data(g["leg_q"], g["leg_qd"], g["leg_p"], g["leg_v"], g["leg_tau_est"]) == g["leg"].tobytes()
    assert lcm_wire.encode_microstrain(g["imu_quat"], g["imu_rpy"], g["imu_omega"], g["imu_acc"], 7, 2) == g["imu"].tobytes()
    assert lcm_wire.decode_contact(g["contact"].tobytes()) == (4, 12.5, [1, 0, 0, 1])


def _make_files(tmp_path, steps=200):
    log, lab = synth.make_sensor_log(steps, seed=2), synth.make_labels(steps, seed=3)
    np.save(tmp_path / "data.npy", log.double().numpy())
    np.save(tmp_path / "label.npy", lab.numpy().reshape(-1, 1))
    np.save(tmp_path / "test.npy", log.double().numpy())
    np.save(tmp_path / "test_label.npy", lab.numpy())
    torch.save({"epoch": 1, "model_state_dict": synth.make_params(0), "val_acc": np.float64(0.5)}, tmp_path / "model.pt")
    rng = np.random.RandomState(1)
    sio.savemat(tmp_path / "raw.mat", {"control_time": np.arange(steps) * 1e-3, "imu_time": np.arange(steps) * 1e-3,
                                        "tau_est": rng.randn(steps, 12), "F": rng.randn(steps, 12), "q": rng.randn(steps, 12),
                                        "qd": rng.randn(steps, 12), "p": rng.randn(steps, 12), "v": rng.randn(steps, 12),
                                        "imu_acc": rng.randn(steps, 3), "imu_omega": rng.randn(steps, 3),
                                        "imu_rpy": rng.randn(steps, 3), "imu_quat": rng.randn(steps, 4)})
    return log, lab


def test_inference_one_seq_script(tmp_path):
    log, lab = _make_files(tmp_path)
    cfg = {"data_path": str(tmp_path / "data.npy"), "label_path": str(tmp_path / "label.npy"),
           "mat_data_path": str(tmp_path / "raw.mat"), "model_load_path": str(tmp_path / "model.pt"),
           "window_size": 150, "batch_size": 1, "calculate_accuracy": True, "save_mat": True,
           "mat_save_path": str(tmp_path / "out.mat"), "save_lcm": True, "lcm_save_path": str(tmp_path / "out.lcm")}
    with open(tmp_path / "cfg.yaml", "w") as f:
        yaml.safe_dump(cfg, f)
    pred = ios.main(["--config_name", str(tmp_path / "cfg.yaml")])
    _, _, want = oracle.inference_stream(synth.make_params(0), log, batch_size=16)
    assert np.array_equal(pred.cpu().numpy(), want.numpy())
    out = sio.loadmat(tmp_path / "out.mat")
    assert np.array_equal(out["contacts_est"], want.numpy()) and out["q"].shape == (51, 12)
    events = list(lcm_wire.read_events(str(tmp_path / "out.lcm")))
    assert len(events) == 3 * 51 and [e[2] for e in events[:3]] == ["leg_control_data", "contact", "microstrain"]
    n, ts, contact = lcm_wire.decode_contact(events[1][3])
    assert n == 4 and contact == want[0].tolist() and abs(ts - 0.149) < 1e-9
    assert [e[0] for e in events] == list(range(len(events)))


def test_test_script(tmp_path):
    log, lab = _make_files(tmp_path)
    cfg = {"data_folder": str(tmp_path) + "/", "model_load_path": str(tmp_path / "model.pt"), "window_size": 150, "batch_size": 30}
    with open(tmp_path / "cfg.yaml", "w") as f:
        yaml.safe_dump(cfg, f)
    acc, per_leg, m = test_script.main(["--config_name", str(tmp_path / "cfg.yaml")])
    _, cls, bits = oracle.inference_stream(synth.make_params(0), log, batch_size=16)
    gt = lab[149:].numpy()
    assert abs(acc - float((cls.numpy() == gt).mean())) < 1e-12
    assert np.allclose(per_leg, (bits.numpy() == oracle.decimal2binary_numpy(gt)).mean(axis=0))
    assert m["confusion_mat"]["total"].sum() == 4 * 51


def test_lcm_decoders_invert_the_reference_encoders(golden_dir):
    g = np.load(os.path.join(golden_dir, "lcm_bytes.npz"))
    q, qd, p, v, tau = lcm_wire.decode_leg_control_data(g["leg"].tobytes())
    for got, key in ((q, "leg_q"), (qd, "leg_qd"), (p, "leg_p"), (v, "leg_v"), (tau, "leg_tau_est")):
        assert np.array_equal(np.asarray(got, dtype=np.float32), g[key])
    quat, rpy, omega, acc, good, bad = lcm_wire.decode_microstrain(g["imu"].tobytes())
    for got, key in ((quat, "imu_quat"), (rpy, "imu_rpy"), (omega, "imu_omega"), (acc, "imu_acc")):
        assert np.array_equal(np.asarray(got, dtype=np.float32), g[key])
    assert (good, bad) == (7, 2)
    with pytest.raises(ValueError):
        lcm_wire.decode_microstrain(g["leg"].tobytes())
    with pytest.raises(ValueError):
        lcm_wire.decode_leg_control_data(g["contact"].tobytes())


class _OracleRunner:
    """CPU stand-in for LatencyRunner (x_host + step()), classifying with the oracle."""

    def __init__(self, params):
        self.params = params
        self.x_host = torch.zeros((1, 150, 54), dtype=torch.float32)

    def step(self):
        logits = oracle.forward_torch(self.params, self.x_host)
        cls = oracle.argmax_class(logits)
        return cls.to(torch.int32), oracle.decimal2binary(cls)


def test_realtime_estimator_ring_and_messages_match_the_reference_loop():
    """RealtimeContactEstimator: rows arrive as leg_control_data + microstrain messages, the
    150-row ring and z-score reproduce contact_dataset.__getitem__, and contact_t messages carry the same bits
    as the reference loop at batch_size 1 over the same log."""
    from deep_contact_estimator_b200.realtime import RealtimeContactEstimator
    params = synth.make_params(0)
    T = 150 + 160                                  # wraps the ring once
    log = synth.make_sensor_log(T, seed=4)
    _, wc, wb = oracle.inference_stream(params, log)
    est = RealtimeContactEstimator(runner=_OracleRunner(params))
    outs = []
    for t in range(T):
        r = log[t]
        leg = lcm_wire.encode_leg_control_data(r[0:12], r[12:24], r[30:42], r[42:54], np.zeros(12))   # q, qd, p, v
        imu = lcm_wire.encode_microstrain([1, 0, 0, 0], [0, 0, 0], r[27:30], r[24:27])                # omega, acc
        msg = est.push_messages(leg, imu, timestamp=t * 1e-3)
        assert (msg is None) == (t < 149)
        if msg is not None:
            outs.append(lcm_wire.decode_contact(msg))
    assert len(outs) == T - 149
    assert [o[2] for o in outs] == wb.tolist()
    assert outs[0][0] == 4 and abs(outs[5][1] - (149 + 5) * 1e-3) < 1e-12
    # push_row gives class + bits
    est2 = RealtimeContactEstimator(runner=_OracleRunner(params))
    res = [est2.push_row(log[t]) for t in range(151)]
    assert res[148] is None and res[149] == (int(wc[0]), tuple(wb[0].tolist())) and res[150][0] == int(wc[1])
    with pytest.raises(ValueError):
        RealtimeContactEstimator()

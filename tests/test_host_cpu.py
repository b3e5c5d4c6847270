"""Host-side mirror of the reference interface, on CPU: module surface,
state_dict compatibility, dataset semantics, the inference loops, sharding.
The CPU path of ``contact_cnn`` is the stock PyTorch layers (the reference's
own behaviour on a CPU); the CUDA path is covered by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch.utils.data import DataLoader

import deep_contact_estimator_b200 as dce
from deep_contact_estimator_b200 import synth, sharding
from oracle import contact_oracle as oracle


def test_state_dict_keys_and_shapes_match_reference():
    m = dce.contact_cnn()
    sd = m.state_dict()
    assert list(sd.keys()) == synth.PARAM_NAMES
    for k, v in sd.items():
        assert tuple(v.shape) == synth.PARAM_SHAPES[k]
    m.load_state_dict(synth.make_params(0))          # strict load of a reference-shaped checkpoint


def test_cpu_forward_equals_oracle(golden_dir):
    g = np.load(os.path.join(golden_dir, "forward_seed0.npz"))
    m = dce.contact_cnn()
    m.load_state_dict(synth.make_params(0))
    m = m.eval()
    with torch.no_grad():
        y = m(synth.make_windows(64, seed=1))
    assert oracle.normwise_rel_err(y.numpy(), g["logits"]) <= 2e-6
    assert np.array_equal(y.argmax(1).numpy(), g["cls"])


def test_training_path_keeps_autograd():
    m = dce.contact_cnn().train()
    y = m(synth.make_windows(2, seed=4))
    y.sum().backward()
    assert m.fc[6].weight.grad is not None


def test_dataset_semantics():
    log, lab = synth.make_sensor_log(200, seed=2), synth.make_labels(200, seed=3)
    ds = dce.contact_dataset(data=log.double().numpy(), label=lab.numpy().reshape(-1, 1), window_size=150, device="cpu")
    assert len(ds) == 51 and ds.data.dtype == torch.float32 and ds.label.dtype == torch.int64
    s = ds[7]
    assert torch.equal(s["data"], oracle.normalize_window(log[7:157]))
    assert int(s["label"]) == int(lab[7 + 149])


def test_inference_loops_cpu(golden_dir):
    g = np.load(os.path.join(golden_dir, "stream_seed2.npz"))
    log, lab = synth.make_sensor_log(420, seed=2), synth.make_labels(420, seed=3)
    ds = dce.contact_dataset(data=log, label=lab.reshape(-1, 1), window_size=150, device="cpu")
    m = dce.contact_cnn(); m.load_state_dict(synth.make_params(0)); m = m.eval()
    loader = DataLoader(ds, batch_size=30)
    bits = dce.inference(loader, m, "cpu")
    assert bits.dtype == torch.uint8 and np.array_equal(bits.numpy(), g["bits"])
    bits2, acc, per_leg = dce.inference_and_compute_acc(loader, m, "cpu")
    assert np.array_equal(bits2.numpy(), g["bits"])
    assert abs(acc - float((g["cls"] == g["labels"]).mean())) < 1e-12 and 0.0 <= acc <= 1.0
    gt_bits = oracle.decimal2binary_numpy(g["labels"])
    assert np.allclose(per_leg, (g["bits"] == gt_bits).mean(axis=0))
    acc3, per_leg3, bp, bg, pa, ga = dce.compute_accuracy(loader, m)
    assert acc3 == acc and np.allclose(per_leg3, per_leg)
    assert bp.dtype == np.float64 and bp.shape == (271, 4) and np.array_equal(pa, g["cls"]) and np.array_equal(ga, g["labels"])


def test_decimal2binary_matches_table(golden_dir):
    tbl = np.load(os.path.join(golden_dir, "bits_table.npz"))["table"]
    assert np.array_equal(dce.decimal2binary(torch.arange(16)).numpy(), tbl)


def test_engine_refuses_cpu_device():
    with pytest.raises(RuntimeError):
        dce.ContactEngine(None, "cpu")


def test_window_range_partition():
    for n in (0, 1, 7, 4096, 9_999_851):
        for world in (1, 2, 3, 8):
            spans = [sharding.window_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.rows_for_windows(10, 20) == (10, 169)      # 149-row halo
    assert sharding.rows_for_windows(5, 5) == (5, 5)


def _gloo_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        log = synth.make_sensor_log(150 + 36, seed=2)                  # 37 windows: ragged over 2 ranks
        n = oracle.num_windows(log.shape[0])
        # weights: rank 0 owns them, one-time broadcast (the NCCL broadcast on GPUs)
        params = synth.make_params(0) if rank == 0 else {k: torch.zeros(s) for k, s in synth.PARAM_SHAPES.items()}
        for k in synth.PARAM_NAMES:
            dist.broadcast(params[k], src=0)
        m = dce.contact_cnn(); m.load_state_dict(params); m = m.eval()
        s, e = sharding.window_range(n, rank, world)
        r0, r1 = sharding.rows_for_windows(s, e)
        ds = dce.contact_dataset(data=log[r0:r1], label=torch.zeros(r1 - r0, dtype=torch.int64), device="cpu")
        assert len(ds) == e - s
        bits = dce.inference(DataLoader(ds, batch_size=8), m, "cpu")
        full = sharding.all_gather_bits(bits, n)
        if rank == 0:
            torch.save(full, os.path.join(tmp, "full.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_inference_equals_single(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    full = torch.load(os.path.join(str(tmp_path), "full.pt"))
    log = synth.make_sensor_log(150 + 36, seed=2)
    _, _, want = oracle.inference_stream(synth.make_params(0), log, batch_size=8)
    assert torch.equal(full, want)


def test_upload_schedule_covers_every_window_once_on_tile_boundaries():
    """ContactEngine.stream_host's chunk schedule: every window exactly once, in order, only after its rows (and,
    except for the last call, its whole 32-window statistics tile) have been uploaded."""
    from deep_contact_estimator_b200 import sharding
    for T, chunk in ((1_000_003, 1 << 18), (5000, 700), (149, 64), (150, 64), (181, 10), (400, 1), (10_000, 10_000), (0, 5)):
        sched = sharding.upload_schedule(T, chunk)
        n = max(T - 149, 0)
        pos = 0
        for k, (rows_up, first, end) in enumerate(sched):
            assert first == pos and end > first and rows_up <= T and end - 1 + 150 <= rows_up
            if k < len(sched) - 1:
                assert end % 32 == 0 and (end - 32) + 181 <= rows_up
            pos = end
        assert pos == n and (not sched or sched[-1][0] == T)
    with pytest.raises(ValueError):
        sharding.upload_schedule(10, 0)


def test_module_copies_and_pickles_without_its_engine():
    """copy.deepcopy / pickle of the module never carries the device handle (it is rebuilt from the parameters)."""
    import copy
    import pickle
    m = dce.contact_cnn().eval()
    m._engine, m._engine_key = object(), ("cuda:0",)          # stand-ins for a live engine
    for clone in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert clone._engine is None and clone._engine_key is None
        assert all(torch.equal(a, b) for a, b in zip(clone.state_dict().values(), m.state_dict().values()))
    assert m._engine is not None


def test_metrics_from_counts_match_the_reference_sklearn_formulation():
    """every metric /root/reference/src/test.py:19-70 prints is a function of the counters
    dce_accuracy_counts accumulates on the device (4 per-leg 2x2 confusion matrices, 16x16 class confusion matrix).
    Checked against the reference's own sklearn formulation (scripts/test.py: compute_metrics) on skewed random
    predictions, including classes that never occur / are never predicted (sklearn's zero-division rule)."""
    import warnings
    import numpy as np
    import torch
    import deep_contact_estimator_b200 as dce
    from deep_contact_estimator_b200.scripts.test import compute_metrics
    from oracle import contact_oracle as oracle
    rng = np.random.default_rng(3)
    for n, classes in ((5000, 16), (300, 5), (64, 16)):
        gt = rng.integers(0, classes, n)
        pred = np.where(rng.random(n) < 0.7, gt, rng.integers(0, 16, n))
        if classes == 5:
            pred[pred == 3] = 4                         # class 3 occurs but is never predicted
        c = dce.counts_from_arrays(pred, gt)
        assert c.shape == (277,) and c[0] == (pred == gt).sum() and c[21:].sum() == n and c[5:21].sum() == 4 * n
        m = dce.metrics_from_counts(c)
        bp, bg = oracle.decimal2binary_numpy(pred).astype(np.float64), oracle.decimal2binary_numpy(gt).astype(np.float64)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = compute_metrics(bp, bg, pred.astype(np.float64), gt.astype(np.float64))
        for key in ("precision_of_class", "precision_of_all_legs", "jaccard_of_class", "jaccard_of_all_legs"):
            assert abs(m[key] - want[key]) < 1e-12, key
        assert np.allclose(m["precision_of_legs"], want["precision_of_legs"], atol=1e-12)
        assert np.allclose(m["jaccard_of_legs"], want["jaccard_of_legs"], atol=1e-12)
        for leg in ("leg_rf", "leg_lf", "leg_rh", "leg_lh", "total"):
            assert np.array_equal(m["confusion_mat"][leg], want["confusion_mat"][leg])
            assert np.isclose(m["fn_rate"][leg], want["fn_rate"][leg], equal_nan=True)
            assert np.isclose(m["fp_rate"][leg], want["fp_rate"][leg], equal_nan=True)
        assert np.allclose(c[1:5], (bp == bg).sum(0))

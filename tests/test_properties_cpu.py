"""Property tests (hypothesis) of the host-side logic: sharding, LCM wire format, oracle invariants."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from deep_contact_estimator_b200 import lcm_wire, sharding, synth
from oracle import contact_oracle as oracle


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 10_000_000), st.integers(1, 16))
def test_window_ranges_partition_exactly(n, world):
    spans = [sharding.window_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [e - s for s, e in spans]
    assert all(x >= 0 for x in sizes) and max(sizes) - min(sizes) <= 1
    for s, e in spans:                                   # every rank's rows cover exactly its windows + 149-row halo
        r0, r1 = sharding.rows_for_windows(s, e)
        assert (r1 - r0) == (0 if e == s else (e - s) + 149)


@settings(max_examples=100, deadline=None)
@given(st.integers(0, 127), st.floats(-1e6, 1e6, allow_nan=False), st.lists(st.integers(0, 1), min_size=4, max_size=4))
def test_contact_message_roundtrip(n_pad, ts, contact):
    n, t, c = lcm_wire.decode_contact(lcm_wire.encode_contact(4, ts, contact))
    assert (n, t, c) == (4, ts, contact)


@settings(max_examples=30, deadline=None)
@given(st.integers(0, 15))
def test_bits_are_msb_first(cls):
    bits = oracle.decimal2binary(torch.tensor([cls]))[0].tolist()
    assert sum(b << (3 - i) for i, b in enumerate(bits)) == cls


@settings(max_examples=10, deadline=None)
@given(st.integers(150, 200), st.integers(0, 2**31 - 1))
def test_stream_windows_equal_batch_windows(steps, seed):
    """inference over a log == forward over the explicitly extracted windows (the oracle's two entry points agree)."""
    log = synth.make_sensor_log(steps, seed=seed % 1000)
    params = synth.make_params(0)
    n = oracle.num_windows(steps)
    logits, cls, bits = oracle.inference_stream(params, log, batch_size=16)
    with torch.no_grad():
        direct = oracle.forward_torch(params, oracle.extract_windows(log, 0, n))
    assert torch.allclose(logits, direct, atol=1e-6) and logits.shape == (n, 16)
    assert torch.equal(bits, oracle.decimal2binary(cls))


@settings(max_examples=20, deadline=None)
@given(st.floats(0.01, 100.0), st.floats(-50.0, 50.0))
def test_zscore_is_affine_invariant(scale, shift):
    """(a*x + b) z-scores to the same window as x: the property that makes per-window normalisation
    remove sensor offsets (utils/data_handler.py:55-56)."""
    w = synth.make_sensor_log(150, seed=4)
    z0 = oracle.normalize_window(w.double())
    z1 = oracle.normalize_window((w.double() * scale + shift))
    assert torch.allclose(z0, z1, atol=1e-6)


def test_fused_kernel_protocols_simulated():
    """tools/simulate_protocols.py: randomised discrete-event simulation of the mbarrier protocols of block2_kernel (weight
    ring, single-buffered accumulators, staging tile) and of block1_kernel's CTA pair (leader issues, peer relays, multicast
    commits): no deadlock, no parity aliasing, no operand hazard over many interleavings — and every wait that is removed
    (one at a time) IS detected, which is what shows the checks can fail."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("simulate_protocols", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                                   "tools", "simulate_protocols.py"))
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    n, missed = sim.check(runs=60)
    assert n == 60 * 4 * 2 + 3 * 60 * 3 and missed == []

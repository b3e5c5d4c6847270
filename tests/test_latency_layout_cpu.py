"""The weight layouts and index arithmetic of the fused latency kernel (csrc/dce_latency.cuh: per-CTA fc.0 / fc.3
slices, the conv4 output-channel quarters, halo rows of the two conv phases), restated in numpy by
tools/emulate_latency.py and checked against the oracle on the CPU.  Pins the formulas, not the synchronisation."""
import importlib.util
import os


def test_latency_kernel_layouts_match_oracle():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "emulate_latency.py")
    spec = importlib.util.spec_from_file_location("emulate_latency", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    e_batch, e_stream = mod.check(batch=1)
    assert e_batch <= 1e-6 and e_stream <= 1e-6


def test_f16f8_fc_layouts_match_float64():
    """Experimental fp16 + e4m3 FC mode (option "fc_f16f8", off by default): weight images, both tape writers and the
    F8 producer / issuer addressing of tapgemm_kernel, restated in numpy by tools/emulate_f16f8.py.  The bound is the
    arithmetic's own error for this data (plain-matrix fp16 + e4m3: 1.2e-5; fp16 alone: 3e-4)."""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "emulate_f16f8.py")
    spec = importlib.util.spec_from_file_location("emulate_f16f8", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    e_fc0, e_fc3 = mod.check(rows=3)
    assert e_fc0 <= 3e-5 and e_fc3 <= 3e-5
    # option "conv_f16f8": block1's X2 writer, the conv weight blocks and block2_kernel<true, true> (slabs, issuer, pool)
    assert mod.check_block2(windows=2) <= 3e-5
    # option "conv_f16f8" = 2: block1's converter, both resident weight images, slab1 and the pooled X2 writer
    assert mod.check_block1(windows=1) <= 4e-5


def test_block2_barrier_protocol_simulation():
    """tools/simulate_block2_protocol.py: every warp role of block2_kernel as a coroutine under a random scheduler, with
    and without clusters (option "block2_cluster": multicast weight blocks, multicast slot release, dummy last rounds):
    no deadlock, no parity aliasing, no operand overwritten under a pending MMA, one L2 fetch per block and cluster."""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "simulate_block2_protocol.py")
    spec = importlib.util.spec_from_file_location("simulate_block2_protocol", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.check(runs=80, seed=3) == 80
    assert mod.check_tapgemm(runs=60, seed=4) == 60            # fc.0 / fc.3 / conv configurations, plain and as CTA pairs
    import pytest
    with pytest.raises(AssertionError, match="deadlock"):      # the one combination launch_layer() refuses (see its comment)
        mod.simulate_tapgemm(1, 2, 3, 4, 2, 5, nbuf=1)
    # the simulation must be able to fail: releasing a ring slot to the local CTA only deadlocks or corrupts a cluster
    import random
    src = open(path).read()
    bad = {}
    exec(compile(src.replace("for dst in (ctas if cl > 1 else [c]):", "for dst in [c]:", 1), "mutant", "exec"), bad)
    with pytest.raises(AssertionError):
        bad["simulate"](2, 3, [3, 3], random.Random(1).randrange(1 << 30))

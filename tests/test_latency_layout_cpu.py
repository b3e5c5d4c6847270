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

"""bench.py's reference arm (the only arm that runs without a GPU): one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--batch", "64"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "windows/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("contact windows/sec") and d["vs_baseline"] is None and d["data"] == "synthetic"
    from oracle import make_ref
    assert d["cpu_baseline"]["kind"] == ("reference" if make_ref.available() else "port")      # oracle/_ref when it is there
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0 and d["latency_b1"]["host_us_p50"] > 0


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_kernel_flops_cover_the_whole_path():
    """roofline.achieved is algorithmic FLOPs of the dominant kernel: the per-kernel figures, keyed by the profiler
    names of dce_forward_profile, must add up to SURVEY.md §8d's 39 368 192 per window — for the default kernels and
    for the layer-wise ablation alike."""
    sys.path.insert(0, ROOT)
    import bench
    default = ["tc_block1", "tc_block2", "tc_fc1", "tc_fc2_fc3_argmax"]
    layerwise = ["tc_ingest", "tc_conv1", "tc_conv2_pool", "tc_conv3", "tc_conv4_pool", "tc_fc1", "tc_fc2", "fc3_argmax_bits"]
    for names in (default, layerwise):
        assert sum(bench.kernel_flops(n) for n in names) == bench.FLOP_PER_WINDOW, names
    assert bench.kernel_flops("tc_fc1") == 19_398_656 and bench.kernel_flops("tc_window_stats") == 0

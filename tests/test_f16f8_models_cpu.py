"""CPU-side checks of the option-gated kernels, none of which has run on a GPU yet:
numpy emulation of the fp16 + e4m3 layouts and issue plans (tools/emulate_f16f8.py), the CUDA source's own packers /
helpers / plans executed on the host (tools/host_check_f16f8.cu), and a randomised simulation of the mbarrier
protocols of block2 and tapgemm with and without clusters (tools/simulate_block2_protocol.py)."""
import importlib.util
import os


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


def test_cuda_source_matches_numpy_emulation_byte_for_byte(tmp_path):
    """The numpy emulation above consumes ITS OWN packers; this ties them to the CUDA source: split16_f16f8,
    pack_b_f16f8_elem and pack_conv_f16f8_elem of csrc/dce_tc.cuh are __host__ __device__, tools/host_check_f16f8.cu
    runs them on the CPU (nvcc-compiled, no CUDA call) and their bytes must equal the emulation's."""
    import shutil
    import subprocess
    import numpy as np
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "host_check_f16f8")
    res = subprocess.run([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                          os.path.join(root, "tools", "host_check_f16f8.cu")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    spec = importlib.util.spec_from_file_location("emulate_f16f8", os.path.join(root, "tools", "emulate_f16f8.py"))
    emu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(emu)
    rng = np.random.default_rng(5)
    for signed, fc_kind in ((0, 4), (1, 3)):
        y = (rng.standard_normal((64, 16)) * np.exp(rng.uniform(-9, 5, (64, 1)))).astype(np.float32)
        y[3, 5], y[7, 0], y[9, 9] = np.nan, 7.0e4, 0.0                        # NaN propagates, saturation at 65504, zero
        if not signed:
            y = np.where(np.isnan(y), y, np.abs(y))
        else:
            y[11, 2] = -9.0e4
        fc_n, fc_k, fc_bn = (256, 4736, 256) if fc_kind == 3 else (256, 2048, 128)
        wfc = (rng.uniform(-1, 1, (fc_n, fc_k)) / np.sqrt(fc_k)).astype(np.float32)
        cout, cin, cin_pad = (64, 54, 64) if signed else (128, 128, 128)
        wcv = (rng.uniform(-1, 1, (cout, cin, 3)) / np.sqrt(3 * cin)).astype(np.float32)
        sw_fc, _ = emu.weight_scale(wfc)
        sw_cv, _ = emu.weight_scale(wcv)
        src, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
        with open(src, "wb") as f:
            f.write(np.array([64, signed, fc_n, fc_k, fc_bn, fc_kind, cout, cin, cin_pad], np.int32).tobytes())
            f.write(np.array([sw_fc, sw_cv], np.float32).tobytes())
            f.write(y.tobytes()); f.write(wfc.tobytes()); f.write(wcv.tobytes())
        assert subprocess.run([exe, src, out]).returncode == 0
        got = np.fromfile(out, np.uint8)
        f16, lo, hi = emu.split16(y, signed=bool(signed))
        want_split = np.concatenate([f16, lo, hi], axis=1).reshape(-1)
        n = want_split.size
        nan_rows = np.isnan(y).any(axis=1)
        ok = (got[:n].reshape(64, 64) == want_split.reshape(64, 64)) | nan_rows[:, None]     # NaN payload bits may differ
        assert ok.all()
        g_nan = got[:n].reshape(64, 64)[3]
        assert np.isnan(g_nan[:32].view(np.float16)[5]) and g_nan[32 + 5] & 0x7F == 0x7F and g_nan[48 + 5] & 0x7F == 0x7F
        want_fc = emu.pack_b_f16f8(wfc, fc_n // fc_bn, fc_k // 32, fc_bn, fc_kind, fc_k, sw_fc)
        assert np.array_equal(got[n:n + want_fc.size], want_fc)
        want_cv = emu.pack_conv_f16f8(wcv, cout, cin, cin_pad, sw_cv)
        assert np.array_equal(got[n + want_fc.size:n + want_fc.size + want_cv.size], want_cv)
        # the layout helpers every f16f8 writer and reader goes through (f8_tape_dst, f8_slab_dst, f8_stage_src)
        tab = got[n + want_fc.size + want_cv.size:].view(np.uint64).astype(np.int64)
        P, K, SLAB = 1 << 40, 1 << 20, 130 * 16
        want_tab = []
        for C in (64, 128, 2048, 4736):
            for g in range(C // 16):
                # fp16 chunks in tape part 0 (chunk 2 g), lo8 chunk g and hi8 chunk C/16 + g in part 1; slab: C/8 fp16, C/16 lo8, C/16 hi8
                want_tab += [2 * g * K, P + g * K, P + (C // 16 + g) * K, 2 * g * SLAB, (C // 8 + g) * SLAB, (C // 8 + C // 16 + g) * SLAB]
        for stages in (148, 64):
            half = stages // 2
            for s_ in range(stages):
                for part in range(2):
                    for j in range(4):
                        want_tab.append(P + (part * 2 * stages + s_ * 4 + j) * K if s_ < half else ((s_ - half) * 8 + part * 4 + j) * K)
        # B-operand offsets inside a conv weight block, as the numpy issuers use them (conv64_mmas / conv_blocks)
        want_tab += [0, 2048, 4096, 6144, 8192, 10240] + [0, 2048, 4096, 6144, 8192, 10240]                 # cout 64: e4m3 [img][tap], fp16 [tap][kk]
        want_tab += [0, 4096, 8192, 12288, 16384, 20480] + [0, 4096, 8192, 12288, 16384, 20480]             # cout 128
        n_tab = len(want_tab)
        assert np.array_equal(tab[:n_tab], np.array(want_tab, np.int64))
        plans = tab[n_tab:]
    # the issue plans dumped from the CUDA source (f8_fc_mma, f8_conv_mma): first they must equal the emulation's own
    # formulas, then the emulation runs ON them — packed bytes, writers, producer addressing and the kernels' own MMA
    # plan together must reproduce float64 x @ W^T
    off = 0
    dumped = {}
    for stages, bn in ((148, 256), (64, 128)):
        k = stages * 4 * 4
        dumped[("fc", stages, bn)] = plans[off:off + k].reshape(stages, 4, 4); off += k
    for half, C, cout in ((2, 64, 64), (2, 64, 128), (4, 128, 128)):
        k = 2 * half * 3 * 2 * 4
        dumped[("conv", half, C, cout)] = plans[off:off + k].reshape(2 * half, 3, 2, 4); off += k
    assert off == plans.size
    for (kind, *cfg), table in dumped.items():
        for idx in np.ndindex(*table.shape[:-1]):
            own = emu.fc_plan(idx[0], cfg[0], idx[1], 4 * emu.SLAB, 4 * cfg[1] * 16, cfg[1] * 16) if kind == "fc" \
                else emu.conv_plan(idx[0], cfg[0], idx[1], idx[2], cfg[1], cfg[2])
            assert tuple(int(v) for v in table[idx]) == own, (kind, cfg, idx)
    emu.PLANS.update(dumped)
    try:
        e_fc0, e_fc3 = emu.check(rows=2)
        assert e_fc0 <= 3e-5 and e_fc3 <= 3e-5 and emu.check_block2(windows=2) <= 3e-5 and emu.check_block1(windows=1) <= 4e-5
    finally:
        emu.PLANS.clear()

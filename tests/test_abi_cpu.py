"""The C-ABI library builds, loads, and exports every symbol include/dce.h
declares.  No compute calls (no GPU here)."""
import ctypes
import os

import pytest

from deep_contact_estimator_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build_library()
    return _lib.load()


def test_library_is_in_tree(lib):
    assert os.path.dirname(_lib.LIB_PATH).endswith("deep_contact_estimator_b200")
    assert os.path.exists(_lib.LIB_PATH)


def test_exports_every_header_symbol(lib):
    names = _lib.header_symbols()
    assert len(names) >= 15 and set(names) == set(_lib._SIGNATURES)
    for n in names:
        assert hasattr(lib, n), n


def test_version_and_strerror(lib):
    assert lib.dce_version() == 100
    assert lib.dce_strerror(0) == b"ok"
    for code in range(-7, 0):
        assert lib.dce_strerror(code) not in (b"ok", b"unknown error")
    assert lib.dce_strerror(-99) == b"unknown error"


def test_argument_errors_without_gpu(lib):
    # argument validation happens before any CUDA call
    assert lib.dce_weights_create(None, 0) == -1
    assert lib.dce_forward(None, None, 1, None, None, None, None, 0, 0, None) == -1
    assert lib.dce_weights_pack(None, None, None) == -1
    assert lib.dce_workspace_bytes(0, 0) == 512         # the header: latency-kernel counters [0,256) + fc.3 tickets [256,512)
    assert lib.dce_workspace_bytes(4096, 0) > 4096 * 4736 * 4
    assert lib.dce_decimal2binary(None, -1, None, None) == -1
    assert lib.dce_weights_destroy(None) == 0
    assert lib.dce_ingest_f64(None, None, -1, None) == -1 and lib.dce_ingest_f64(None, None, 0, None) == 0
    assert lib.dce_ingest_f64(None, None, 8, None) == -1
    assert lib.dce_weights_set_option(None, b"fuse_block1", 1) == -1


def test_workspace_covers_the_latency_kernel(lib):
    """Calls of <= 4 windows run the fused latency kernel, whose buffers (256-byte counter header, P1, A4, H1,
    logit shares for 4 windows) must fit the workspace of ANY call size in both precision modes."""
    need = 256 + 4 * (75 * 64 + 4736 + 2048 + 128 * 16) * 4
    for precision in (0, 1):
        sizes = [lib.dce_workspace_bytes(n, precision) for n in (1, 2, 4, 5, 4096, 1 << 20)]
        assert all(s >= need for s in sizes)
        assert sizes == sorted(sizes) and sizes[-1] == sizes[-2]          # saturates at one internal chunk
    assert lib.dce_workspace_bytes(1, 7) == 0                                # unknown precision


def test_no_cpu_fallback_on_missing_library(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no fallback"):
        _lib.load()


def test_header_is_plain_c_and_links_from_c(lib, tmp_path):
    """include/dce.h compiles as C99 (no C++ or torch types at the boundary) and a C program links against the
    library: examples/realtime_step.c is the control-loop caller a C/C++ front end would write."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not found")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda_inc = "/usr/local/cuda/include"
    cuda_lib = "/usr/local/cuda/lib64"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime_api.h")):
        pytest.skip("CUDA toolkit headers not found")
    exe = str(tmp_path / "realtime_step")
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"), "-I", cuda_inc,
           os.path.join(root, "examples", "realtime_step.c"), "-L", os.path.dirname(_lib.LIB_PATH), "-ldce_b200",
           "-L", cuda_lib, "-lcudart", "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH), "-Wl,-rpath," + cuda_lib, "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0 and "version 100" in run.stdout and "invalid argument" in run.stdout


def test_torch_extension_registers_the_ops():
    """contact_cnn.forward -> torch.ops.dce.forward -> dce_forward.  _dce_torch.so (TORCH_LIBRARY `dce`)
    loads without a GPU, registers forward / stream / accuracy_counts, answers shape queries on the Meta backend
    (fake-tensor tracing) and refuses CPU tensors (no CPU kernel exists: the product path has no fallback)."""
    import pytest
    import torch
    from deep_contact_estimator_b200 import _lib, build
    build.build_torch_extension()
    ops = _lib.torch_ops()
    assert ops is not None
    x = torch.empty(7, 150, 54, device="meta")
    ws = torch.empty(16, dtype=torch.uint8, device="meta")
    lo, cl, bi = ops.forward(0, x, ws, 1, True, True, False)
    assert lo.shape == (7, 16) and lo.dtype == torch.float32 and cl.shape == (7,) and cl.dtype == torch.int32 and bi.shape == (0, 4)
    lo, cl, bi = ops.stream(0, torch.empty(1000, 54, device="meta"), 3, 500, ws, 1, False, True, True)
    assert lo.shape == (0, 16) and cl.shape == (500,) and bi.shape == (500, 4) and bi.dtype == torch.uint8
    with pytest.raises((RuntimeError, NotImplementedError)):
        ops.forward(0, torch.zeros(2, 150, 54), torch.zeros(16, dtype=torch.uint8), 1, True, True, True)
    schema = str(torch.ops.dce.forward.default._schema)
    assert "Tensor x" in schema and "int handle" in schema


def test_server_control_blocks_match_the_python_runners(tmp_path):
    """The resident servers talk through plain structs in pinned memory (dce_latency_ctrl, dce_latency_row_ctrl): the word
    indices the Python runners poke must be the header's offsets (the CUDA side static_asserts the same against its own
    constants, csrc/dce.cu)."""
    import shutil
    import subprocess
    from deep_contact_estimator_b200.engine import LatencyRunner, RowRunner
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not found")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "dce.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(dce_latency_ctrl), offsetof(dce_latency_ctrl, seq_in) / 4, offsetof(dce_latency_ctrl, quit) / 4,
           offsetof(dce_latency_ctrl, seq_out) / 4, offsetof(dce_latency_ctrl, cls0) / 4, offsetof(dce_latency_ctrl, bits0) / 4,
           offsetof(dce_latency_ctrl, device_ns) / 4, offsetof(dce_latency_ctrl, alive) / 4);
    printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(dce_latency_row_ctrl), offsetof(dce_latency_row_ctrl, chunk) / 4,
           offsetof(dce_latency_row_ctrl, seq_out) / 4, offsetof(dce_latency_row_ctrl, cls0) / 4, offsetof(dce_latency_row_ctrl, bits0) / 4,
           offsetof(dce_latency_row_ctrl, device_ns) / 4, offsetof(dce_latency_row_ctrl, alive) / 4);
    return 0;
}
''')
    exe = str(tmp_path / "layout")
    res = subprocess.run([gcc, "-std=c99", "-I", os.path.join(root, "include"), str(src), "-o", exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    a, b = [list(map(int, line.split())) for line in subprocess.run([exe], capture_output=True, text=True).stdout.splitlines()]
    L = LatencyRunner
    assert a == [128, L._SEQ_IN, L._QUIT, L._SEQ_OUT, L._CLS0, L._BITS0, L._DEVICE_NS, L._ALIVE]
    R = RowRunner
    assert b == [512, 0, R._SEQ_OUT, R._CLS0, R._BITS0, R._DEVICE_NS, R._ALIVE]

#!/usr/bin/env python
"""Build ``oracle/_ref/``: the reference's OWN implementation of the path, compiled where its sources lie.

    python oracle/make_ref.py            # needs /root/reference (this container; read-only)

TEST / BENCH INFRASTRUCTURE ONLY (as the rest of ``oracle/``).  The reference is pure Python; its two modules on
the hot path are byte-compiled from ``/root/reference`` by CPython's own compiler

    /root/reference/src/contact_cnn.py        ->  oracle/_ref/contact_cnn.bc      (contact_cnn, src/contact_cnn.py:7-66)
    /root/reference/utils/data_handler.py     ->  oracle/_ref/data_handler.bc     (contact_dataset, utils/data_handler.py:13-61)

the way a C reference would be compiled into ``oracle/_ref/*.so``: no reference SOURCE enters the repo, the
outputs are git-ignored and travel to the GPU box with the snapshot (same image, same interpreter, so the byte code
loads there; the files are CPython ``.pyc`` images under another suffix, because snapshot tools skip ``*.pyc``).  ``load()`` imports them back (sourceless loader).  Users:

  * ``bench.py --impl reference`` and ``cpu_baseline``: time the reference module itself on the box's host cores
    (``kind: "reference"``); when ``oracle/_ref`` is absent they fall back to the oracle port (``kind: "port"``);
  * ``tests/test_oracle_cpu.py``: the oracle port must agree with it bit for bit (skipped when absent).
The loop functions of ``src/inference_one_seq.py`` are not compiled (the module imports ``lcm`` and ``yaml`` at the
top and cannot load here); ``oracle/make_golden.py`` runs them with a stub ``lcm`` to make fixtures.
"""
import importlib.machinery
import importlib.util
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("DCE_REFERENCE", "/root/reference")
UNITS = {"contact_cnn": os.path.join("src", "contact_cnn.py"), "data_handler": os.path.join("utils", "data_handler.py")}


def make(quiet: bool = False) -> str:
    """Compile the reference modules into oracle/_ref/ when /root/reference is present; returns what happened."""
    if not os.path.isdir(REF):
        return "kept (no /root/reference here)" if available() else "absent (no /root/reference here)"
    os.makedirs(OUT, exist_ok=True)
    for name, rel in UNITS.items():
        py_compile.compile(os.path.join(REF, rel), cfile=os.path.join(OUT, name + ".bc"), dfile=rel, doraise=True)
    with open(os.path.join(OUT, "README"), "w") as f:
        f.write("byte-compiled from /root/reference by oracle/make_ref.py (python %d.%d); git-ignored, not source\n" % sys.version_info[:2])
    if not quiet:
        print("compiled", ", ".join(UNITS.values()), "->", OUT)
    return "compiled"


def available() -> bool:
    return all(os.path.exists(os.path.join(OUT, n + ".bc")) for n in UNITS)


_mods = {}


def load():
    """-> (contact_cnn class, contact_dataset class) of the REFERENCE, from oracle/_ref/*.bc; raises if absent."""
    if not available():
        raise FileNotFoundError("oracle/_ref is empty: run `python oracle/make_ref.py` where /root/reference exists")
    for name in UNITS:
        if name not in _mods:
            path = os.path.join(OUT, name + ".bc")
            loader = importlib.machinery.SourcelessFileLoader("dce_reference_" + name, path)
            spec = importlib.util.spec_from_loader(loader.name, loader)
            mod = importlib.util.module_from_spec(spec)
            loader.exec_module(mod)
            _mods[name] = mod
    return _mods["contact_cnn"].contact_cnn, _mods["data_handler"].contact_dataset


if __name__ == "__main__":
    print(make())

"""The drop-in boundary is a C ABI: ``include/pcb200.h`` must be valid strict C (not only C++), and a host that is neither
Python nor torch must be able to link ``libpcb200.so`` and call it.  ``examples/c_host.c`` is compiled with
``gcc -std=c99 -pedantic -Werror`` and run: the integer planning of BASELINE configs[4] (2048^3 / 160^3 / 0.5 — reference
``window.py:57-134``, ``lazy.py:269-365``) comes back with the reference's window counts, bad arguments come back as
PCB_ERR_INVALID with a message, and without a GPU the device probe says so instead of crashing."""

import os
import shutil
import subprocess

import pytest

from conftest import ROOT
from pytorch_connectomics_b200 import _lib


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_c_host_compiles_as_c99_and_runs(tmp_path):
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    exe = tmp_path / "c_host"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_host.c"), "-L", _lib.CSRC, "-lpcb200", f"-Wl,-rpath,{_lib.CSRC}", "-o", str(exe)]
    built = subprocess.run(cmd, capture_output=True, text=True)
    assert built.returncode == 0, built.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "interval 80 80 80" in run.stdout and "eager windows 15625 (last start 1888 1888 1888)" in run.stdout
    assert "lazy windows 19683" in run.stdout and 'invalid roi -> "roi_size must contain positive values"' in run.stdout


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_header_alone_is_strict_c(tmp_path):
    src = tmp_path / "only_header.c"
    src.write_text('#include "pcb200.h"\nint main(void) { return PCB_COMM_ID_BYTES == 128 ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I",
                        os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def window_goldens():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "window_goldens.npz"))


@pytest.fixture(scope="session")
def mednext_tiny_golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "mednext_tiny.npz"))


def free_port() -> int:
    """A TCP port the kernel reports as free right now (gloo rendezvous of the multi-process CPU tests)."""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return int(sk.getsockname()[1])


# ----------------------------------------------------------------------------- hang isolation of first-run GPU tests
# Tests marked ``isolated`` (the late-sorted files holding code whose first B200 run is the driver's) do not run in this
# process.  The first one reached starts ONE child ``pytest -v --runxfail`` over its whole file and every test of the file then
# reports what the child reported for it.  A kernel that hangs cannot be interrupted from Python (the SIGALRM of
# pytest-timeout is only served once the blocked CUDA call returns), so without this a single hang in new code would stall
# the whole ``pytest -m gpu`` run and lose the verdict of every measured-kernel test before it.  The parent watches the
# child's per-test lines; a test silent for longer than the stall limit gets the child's process group killed (the driver
# tears the hung context down) and the tests the child never reached are reported as failed.
ISOLATED_CHILD_ENV = "PCB_ISOLATED_CHILD"
_ISOLATED_RESULTS = {}


def _run_isolated_file(path: str, stall_s: float, marker_expr: str, total_s: float = 1500.0):
    import queue
    import re
    import signal
    import subprocess
    import threading
    env = dict(os.environ, **{ISOLATED_CHILD_ENV: "1", "PYTHONUNBUFFERED": "1"})
    cmd = [sys.executable, "-m", "pytest", path, "-v", "--runxfail", "-p", "no:cacheprovider", "-m", marker_expr]
    proc = subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                            start_new_session=True)
    lines, q = [], queue.Queue()

    def pump():
        for line in proc.stdout:
            q.put(line)
        q.put(None)

    threading.Thread(target=pump, daemon=True).start()
    results, killed = {}, None
    import time
    t0 = time.time()
    pat = re.compile(r"^(\S+\.py::.+?) (PASSED|FAILED|ERROR|SKIPPED|XFAIL|XPASS)\b")
    while True:
        try:
            if time.time() - t0 > total_s:
                raise queue.Empty
            line = q.get(timeout=stall_s)
        except queue.Empty:
            killed = (f"no test finished within {stall_s:.0f} s" if time.time() - t0 <= total_s
                      else f"file not done after {total_s:.0f} s") + ": child process group killed"
            try:
                os.killpg(proc.pid, signal.SIGKILL)
            except ProcessLookupError:
                pass
            break
        if line is None:
            break
        lines.append(line.rstrip("\n"))
        m = pat.match(line)
        if m:
            results[m.group(1).split("::", 1)[1]] = m.group(2)
    try:
        proc.wait(timeout=30)
    except Exception:
        pass
    return {"results": results, "killed": killed, "tail": "\n".join(lines[-80:]), "returncode": proc.returncode}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "isolated(stall=300): run the test's file in a child pytest process (hang isolation)")


@pytest.hookimpl(tryfirst=True)
def pytest_pyfunc_call(pyfuncitem):
    mark = pyfuncitem.get_closest_marker("isolated")
    if mark is None or os.environ.get(ISOLATED_CHILD_ENV):
        return None                     # run normally (in the child, or an unmarked test)
    path = str(pyfuncitem.fspath)
    if path not in _ISOLATED_RESULTS:
        import signal
        if hasattr(signal, "setitimer"):        # this test waits for the whole file's child run: the stall limit is the bound,
            signal.setitimer(signal.ITIMER_REAL, 0)      # not this one test's pytest-timeout alarm
        expr = pyfuncitem.config.getoption("-m") or "gpu"
        _ISOLATED_RESULTS[path] = _run_isolated_file(path, float(mark.kwargs.get("stall", 300)), expr)
    rep = _ISOLATED_RESULTS[path]
    key = pyfuncitem.nodeid.split("::", 1)[1]
    status = rep["results"].get(key)
    if status in ("PASSED", "XPASS"):
        return True
    if status == "SKIPPED":
        pytest.skip("skipped in the isolated child run")
    why = rep["killed"] or f"child exit code {rep['returncode']}"
    shown = "" if rep.get("shown") else "\n---- tail of the isolated child run ----\n" + rep["tail"]
    rep["shown"] = True
    pytest.fail(f"isolated child run: {key} -> {status or 'NOT REACHED'} ({why}){shown}", pytrace=False)

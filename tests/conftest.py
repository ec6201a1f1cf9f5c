import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (plain-C restatement).  Built on demand with gcc."""
    import subprocess
    so = os.path.join(ROOT, "oracle", "libmorpho_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libmorpho_oracle.so"])
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference behind our harness, when oracle/_ref has been built."""
    from oracle.oracle import Reference, reference_available
    if not reference_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    try:
        return Reference(0)
    except OSError as e:      # e.g. OpenBLAS of the image not found
        pytest.skip(f"reference harness not loadable: {e}")


@pytest.fixture(scope="session")
def cases():
    from tests.golden.cases import all_cases
    return all_cases()


def pytest_sessionfinish(session, exitstatus):
    """Worst achieved parity errors of this session (tests/parity.py record()) → gpurun_out/parity_errors.json."""
    import json
    from tests import parity
    rec = parity.recorded()
    if not rec:
        return
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_errors.json"), "w") as fh:
            json.dump({g: {str(k): v for k, v in sorted(d.items(), key=lambda kv: str(kv[0]))} for g, d in rec.items()}, fh, indent=1)
    except OSError:
        pass

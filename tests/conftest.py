import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The shared library is a build artefact (git-ignored): compile it if a fresh checkout has none.  nvcc cross-compiles
    sm_100a without a GPU; when nvcc is absent too, the tests that need the library fail with its own loud message."""
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "kvmatch_b200", "libkvmatch_gpu.so")
    if not os.path.exists(lib) and shutil.which("nvcc") and shutil.which("make"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "kvmatch_b200", "csrc")], check=False,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle():
    from oracle import kvm_oracle
    kvm_oracle.lib()
    return kvm_oracle

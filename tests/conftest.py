import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

# The C-ABI library is a build artefact (git-ignored).  If the tests are run on a fresh checkout before
# __graft_entry__.build(), build it here (nvcc cross-compiles without a GPU) rather than failing at import.
_LIB = ROOT / "srsran_ce_pytorch_b200" / "libsrsce.so"
if not _LIB.exists():
    import subprocess
    subprocess.run(["make", "-C", str(ROOT / "srsran_ce_pytorch_b200" / "csrc")], check=True)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

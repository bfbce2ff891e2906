import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    # tests marked `reference` are skipped wherever the reference sources are absent (e.g. the GPU box)
    from oracle.refharness import reference_available
    if reference_available():
        return
    skip = pytest.mark.skip(reason="reference sources not present")
    for it in items:
        if "reference" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(autouse=True, scope="session")
def _oracle_pow_mode():
    """RRTFND_ORACLE_LIBM_POW=1 runs every oracle comparison with the reference's `**` evaluated as libm pow() (csrc/libm_pow.h):
    the setting that goes with a -DRRTFND_LIBM_POW=1 build of the library selected through RRTFND_LIB (scripts/gpu_variants.sh)."""
    if os.environ.get("RRTFND_ORACLE_LIBM_POW") == "1":
        import oracle
        oracle.set_libm_pow(True)
    yield

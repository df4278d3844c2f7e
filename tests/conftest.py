import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    # a fresh checkout has no built libraries (they are git-ignored): build product + oracle once, like the driver's build()
    need = [os.path.join(ROOT, "mgcod_b200", "libmgcod_b200.so"), os.path.join(ROOT, "oracle", "liboracle.so"),
            os.path.join(ROOT, "workload", "libmgc_workload.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def model_dir(tmp_path_factory):
    """Parametric 82-section model set (41 bacterial + 41 archaeal GC bins) for codes 11/4/15/101."""
    from mgcod_b200 import synth
    d = tmp_path_factory.mktemp("models")
    synth.write_model_set(str(d), kind="parametric")
    return str(d)


@pytest.fixture(scope="session")
def small_model_dir(tmp_path_factory):
    """Few-section variants with odd orders to exercise ramp / gap-fill rules."""
    from mgcod_b200 import synth
    d = tmp_path_factory.mktemp("models_small")
    synth.write_model_set(str(d), kind="parametric", cod_order=3, non_order=1, bins=(28, 35, 44, 57), domains=("B", "A"))
    return str(d)

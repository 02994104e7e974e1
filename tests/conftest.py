import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "tests", ROOT / "tests" / "emu"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build (or refresh) the native pieces once per session: product .so, host layer, oracle, emulation."""
    import __graft_entry__ as g
    missing = [p for p in (ROOT / "rustracer_b200/csrc/_build/librt_b200.so", ROOT / "host/_build/libgltf_host.so",
                           ROOT / "oracle/_build/liborc.so", ROOT / "tests/emu/_build/librt_emu.so") if not p.exists()]
    if missing:
        g.build()
    yield


@pytest.fixture(scope="session")
def cornell_desc():
    import util
    return util.load_scene_npz(util.GOLDEN / "cornell_box_scene.npz")


@pytest.fixture(scope="session")
def cornell_oracle(cornell_desc):
    from oracle import orc
    return orc.OracleScene(cornell_desc)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    import util
    return np.load(util.GOLDEN / "cornell_box_golden.npz")

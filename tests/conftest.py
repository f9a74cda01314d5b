import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Native pieces are built once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def ref_available(built):
    from oracle import ref
    return ref.available("default") and ref.available("strict")


def pytest_collection_modifyitems(config, items):
    """tests/test_emulated_cuda.py runs its tests on two contexts (two-kernel / fused); only some tests mean
    something on both (see FUSED_TESTS / FUSED_ONLY there): the other combinations are not collected."""
    try:
        from tests.test_emulated_cuda import FUSED_ONLY, FUSED_TESTS
    except Exception:  # noqa: BLE001
        return
    keep, drop = [], []
    for it in items:
        cs = getattr(it, "callspec", None)
        which = cs.params.get("emu_pipe") if cs is not None and it.fspath.basename == "test_emulated_cuda.py" else None
        name = getattr(it, "originalname", None) or it.name
        if (which == "fused" and name not in FUSED_TESTS) or (which == "two-kernel" and name in FUSED_ONLY):
            drop.append(it)
        else:
            keep.append(it)
    if drop:
        items[:] = keep
        config.hook.pytest_deselected(items=drop)

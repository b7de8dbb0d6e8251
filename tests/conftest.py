"""pytest configuration: `gpu` marker + import paths.

`-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI loading.
`-m gpu` runs on a B200: the parity tests proper, all through the C-ABI of libb200vs.so.
"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "mlx-vector-db_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def native_lib():
    """Build (if stale) and load libb200vs.so."""
    from b200vs import build, _cabi
    build.build_library()
    return _cabi.lib()


@pytest.fixture
def make_store(tmp_path, native_lib):
    """Factory for engine stores; closes them after the test."""
    from b200vs import MLXVectorStore, MLXVectorStoreConfig
    made = []

    def _make(dim, metric="cosine", **kw):
        kw.setdefault("persist", False)
        kw.setdefault("max_vectors", 2_000_000)
        cfg = MLXVectorStoreConfig(dimension=dim, metric=metric, **kw)
        s = MLXVectorStore(str(tmp_path / f"store_{len(made)}"), cfg)
        made.append(s)
        return s

    yield _make
    for s in made:
        s.close()

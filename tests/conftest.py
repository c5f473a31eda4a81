import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "gimp-fix-ca_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def restatement():
    import oracle

    return oracle.Restatement()


@pytest.fixture(scope="session")
def reference():
    import oracle

    if not oracle.Reference.available():
        pytest.skip("oracle/_ref/libfixca_ref.so not built (needs /root/reference at build time)")
    return oracle.Reference()


@pytest.fixture(scope="session")
def checker():
    """Strongest CPU checker present: the reference's own code, else our restatement."""
    import oracle

    return oracle.best_checker()


@pytest.fixture(scope="session")
def fx():
    """The product binding; the library must exist (there is no fallback)."""
    import fixca

    fixca.load()
    return fixca


@pytest.fixture
def tuning(fx, monkeypatch):
    """Set a FIXCA_* tuning variable for one test.  The library reads them once per process, so the change
    (and its undo) is followed by fixca_cuda_reload_tuning()."""
    def set_var(name, value):
        monkeypatch.setenv(name, value)
        fx.reload_tuning()

    yield set_var
    monkeypatch.undo()
    fx.reload_tuning()

import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Native artefacts (nvcc cross-compiles without a GPU; the prebuilt files travel to the GPU box)."""
    from qtorch_b200 import build
    build.build_all()
    from oracle import oracle
    oracle.build()
    return True


@pytest.fixture(scope="session")
def networks():
    return json.load(open(os.path.join(GOLDEN, "networks.json")))


@pytest.fixture(scope="session")
def engine(built):
    import qtorch_b200 as qt
    e = qt.Engine(0)
    yield e
    e.close()


def golden_paths(rec):
    """absolute (cwd, qasm, measure, ordering) for a networks.json record"""
    cwd = os.path.join(GOLDEN, rec.get("cwd", ""))
    qasm = os.path.join(GOLDEN, rec["qasm"])
    if rec.get("cwd"):
        qasm = os.path.relpath(qasm, cwd)
    meas = os.path.join(GOLDEN, rec["measure"])
    ordering = os.path.join(GOLDEN, rec["ordering"]) if "ordering" in rec else None
    return cwd, qasm, meas, ordering


def plan_file(rec, tmpdir):
    """write the golden plan (mCreatedFrom pairs) as an 'a b' per line file for the `seq` modes"""
    path = os.path.join(str(tmpdir), "plan.txt")
    with open(path, "w") as f:
        for p in rec["plan"]:
            a, b = p.split(",")
            f.write("%s %s\n" % (a, b))
    return path

import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dist-gnn_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _load_reference():
    """The unmodified reference compiled for sm_100a by oracle/build_ref.sh (differential oracle)."""
    import glob
    cands = glob.glob(os.path.join(ROOT, "oracle", "_ref", "dgs.cpython-*.so"))
    if not cands:
        return None
    import torch  # noqa: F401  (the extension links libtorch)
    # The reference module is also called `dgs` and pybind's def_submodule() registers
    # "dgs.classes" / "dgs.ops" through sys.modules: hide this repo's package while it loads.
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "dgs" or k.startswith("dgs.")}
    try:
        spec = importlib.util.spec_from_file_location("dgs", cands[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k in list(sys.modules):
            if k == "dgs" or k.startswith("dgs."):
                sys.modules.pop(k)
        sys.modules.update(saved)
    return mod


_REF = {}


@pytest.fixture(scope="session")
def ref():
    """Reference `dgs` module with its NCCL context initialised for one rank (its classes read the
    rank from the context, src/sampling/sampler.cc:72)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    if "mod" not in _REF:
        try:
            mod = _load_reference()
        except Exception as e:  # pragma: no cover
            mod = None
            _REF["err"] = repr(e)
        if mod is not None:
            torch.cuda.set_device(0)
            mod.ops._CAPI_set_nccl(1, mod.ops._CAPI_get_unique_id(), 0)
        _REF["mod"] = mod
    if _REF["mod"] is None:
        pytest.skip("oracle/_ref not built: " + _REF.get("err", "run oracle/build_ref.sh"))
    return _REF["mod"]


@pytest.fixture(scope="session")
def dgs():
    import dgs as m
    return m


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)

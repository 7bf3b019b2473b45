"""``dgs`` - drop-in for the reference's pybind11 extension module of the same name
(src/pybind.cc:17-78): ``dgs.classes.{P2PCacheSampler, P2PCacheFeatureServer, TensorP2PServer}``
and ``dgs.ops._CAPI_* / _Test_*``, implemented over the sm_100a C-ABI library
``libdgs_b200.so`` (include/dgs_b200.h).  There is no CPU or PyTorch fallback."""
from . import _lib
from . import ops
from . import classes

__all__ = ["ops", "classes"]


def launch_count():
    """Kernels launched by the native library since load (bench.py's ``gpu_launches``)."""
    return int(_lib.lib().dgs_launch_count())

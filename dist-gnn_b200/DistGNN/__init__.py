"""``DistGNN`` - the reference's Python package (python/DistGNN/__init__.py:1-4) over the
B200-native ``dgs`` module: ``DistGNN.capi`` is ``dgs``."""
from . import cache
from . import dataloading
from . import dist
import dgs as capi

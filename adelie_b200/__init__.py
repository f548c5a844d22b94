"""adelie_b200 -- B200-native (sm_100a) group-elastic-net coordinate-descent path.

Keeps the ``adelie.solver.grpnet`` / ``adelie.state`` / ``adelie.matrix`` / ``adelie.glm`` /
``adelie.bcd`` / ``adelie.configs`` surface of JamesYang007/adelie for the naive-method path and
runs it on hand-written CUDA kernels through the C ABI in ``include/adelie_b200.h``.
There is no CPU fallback: every operator raises if ``libadelie_b200.so`` or a GPU is missing.
"""
from . import _lib
from . import bcd
from . import configs
from . import cv
from . import data
from . import diagnostic
from . import dist
from . import glm
from . import io
from . import matrix
from . import solver
from . import state
from .configs import set_configs
from .solver import grpnet, gaussian_cov
from .cv import cv_grpnet

__version__ = "0.1.0"


def __getattr__(name):
    # `adelie_b200.sklearn` needs scikit-learn: imported on first use so that the package itself does not depend on it
    if name == "sklearn":
        import importlib
        return importlib.import_module(".sklearn", __name__)
    raise AttributeError(name)

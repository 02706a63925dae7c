"""gwin_b200 -- B200-native batched log-likelihood-ratio path of gwastro/gwin.

The package holds only what that path needs: the CUDA library and its C ABI
(``csrc/``, ``include/gwin_b200.h``), the Python models that keep gwin's model /
``CallModel`` / pool protocols, and the ensemble stepper that calls them.
"""
from . import _lib  # noqa: F401
from ._lib import GwinCudaError, build, pinned_empty  # noqa: F401

__version__ = "0.1.0"

"""tensor_networks_b200 -- B200-native tensor-train core sweeps behind the pytens API.

Only the hot path named in BASELINE.json is implemented (TT inner product / norm,
TT rounding, TT-SVD, and their batched forms); see DESIGN.md.  Importing this
package does not need a GPU, but every numerical entry point does: there is no
CPU fallback.
"""

from .types import Index, SVDConfig  # noqa: F401
from .tt import TensorTrain  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["Index", "SVDConfig", "TensorTrain"]
__version__ = "0.2.0"

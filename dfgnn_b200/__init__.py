"""dfgnn_b200 -- B200-native (sm_100a) fused attention-convolution for GT / GAT / AGNN.

Drop-in for the hot path of zli96/DF-GNN: ``operators`` mirrors
``DFGNN/operators``, ``layers`` mirrors ``DFGNN/layers``; underneath is the C-ABI
library ``libdfgnn_b200.so`` (``include/dfgnn_b200.h``).  No CPU fallback."""
from . import _lib  # noqa: F401

__all__ = ["_lib", "formats", "graphs", "layers", "operators", "utils"]
__version__ = "0.1.0"

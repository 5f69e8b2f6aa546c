"""Mirror of ``DFGNN/layers/__init__.py`` (+ the per-format modules the factories return)."""
from .AGNN import *  # noqa: F401,F403
from .GAT import *  # noqa: F401,F403
from .GT import *  # noqa: F401,F403
from .util import (load_graphconv_layer, load_layer_AGNN, load_layer_GAT, load_layer_GT,  # noqa: F401
                   load_prepfunc, preprocess_CSR, preprocess_dglsp, preprocess_gat_fw_bw,
                   preprocess_Hyper, preprocess_Hyper_fw_bw, preprocess_softmax)

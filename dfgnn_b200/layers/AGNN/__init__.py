from .agnn_layer import (AGNNConv_csr, AGNNConv_csr_gm, AGNNConv_forward, AGNNConv_hyper,
                         AGNNConv_softmax, AGNNConv_softmax_gm, AGNNConv_tiling, AGNNConvDGL)

from .gatconv_layer import (GATConv_dgNN, GATConv_forward, GATConv_hyper, GATConv_hyper_recompute,
                            GATConv_hyper_v2, GATConv_softmax, GATConv_softmax_gm, GATConv_tiling,
                            GATConvDGL)

from .gtconv_layer import (SparseMHA, SparseMHA_CSR, SparseMHA_CSR_GM, SparseMHA_forward,
                           SparseMHA_forward_timing, SparseMHA_hyper, SparseMHA_softmax,
                           SparseMHA_softmax_gm, SparseMHA_tiling)

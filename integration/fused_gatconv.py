"""Drop-in for the reference's pybind module ``fused_gatconv``
(``DFGNN/src/fused_gatconv/fused_gatconv.cpp:355-372``); see fused_gtconv.py here."""
from dfgnn_b200.operators._native import (  # noqa: F401
    gat_backward,
    gat_forward,
    gat_inference,
    gat_inference_hyper,
    gat_inference_hyper_recompute,
    gat_inference_hyper_v2,
    gat_inference_softmax,
    gat_inference_softmax_gm,
    gat_inference_tiling,
)

# same maths as gat_inference_hyper (fused_gatconv.cpp:371)
gat_inference_hyper_ablation = gat_inference_hyper

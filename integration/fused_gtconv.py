"""Drop-in for the reference's pybind module ``fused_gtconv``
(``DFGNN/src/fused_gtconv/fused_gtconv.cpp:577-602``).

Put this directory on ``sys.path`` (ahead of any built reference extension) and
the UNMODIFIED ``DFGNN/operators/fused_gtconv.py`` -- which does
``import fused_gtconv as fused_gt`` -- runs on the B200 kernels.  See
INTEGRATION.md."""
from dfgnn_b200.operators._native import (  # noqa: F401
    gt_backward,
    gt_csr_gm_inference,
    gt_csr_inference,
    gt_hyper_forward,
    gt_hyper_inference,
    gt_softmax_gm_inference,
    gt_softmax_inference,
    gt_tiling_inference,
)

# the ablation export (fused_gtconv.cpp:595) computes the same function as gt_hyper_inference
gt_hyper_inference_ablation = gt_hyper_inference

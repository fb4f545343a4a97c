"""decaf_b200 — B200-native (sm_100a) DeCaf-Grounder inference hot path.

Host-side mirror of the reference's API surface for this path:
    decaf_b200.modeling   <-> libs/modeling   (make_video_net / make_text_net / make_fusion /
                                               make_head, PtTransformerEarlyFusionIterative, PtGenerator)
    decaf_b200.nms        <-> libs/nms        (batched_nms, NMSop, SoftNMSop, nms_1d_gpu)
    decaf_b200.worker_v2  <-> libs/worker_v2  (Evaluator, create_model)
All arithmetic runs in libdecaf_b200.so (include/decaf_b200.h); importing the modules above fails
loudly when the library has not been built.  `decaf_b200.synth` (synthetic data) is pure Python.
"""
__all__ = ['synth']

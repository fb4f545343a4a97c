from .nms import batched_nms, NMSop, SoftNMSop, nms_1d_gpu

"""pfst_b200 — B200-native (sm_100a) self-training hot path of zhu-xlab/PFST.

Host code is Python/PyTorch (device memory, streams, torch.distributed); the
arithmetic is in hand-written CUDA kernels behind a C ABI
(include/pfst_sm100.h -> pfst_b200/csrc/libpfst_sm100.so). No CPU fallback.
"""
__version__ = "0.1.0"

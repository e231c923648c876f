"""pemp_b200 - B200-native (sm_100a) prototype-matching head for Jarvis73/PEMP's few-shot models.

Host code is Python/PyTorch; every hot op is a hand-written CUDA kernel reached through the C ABI
declared in `include/pemp_b200.h` (built in-tree as `pemp_b200/libpemp_b200.so`).  There is no CPU
fallback: operators raise if the library is missing or a tensor is not on a CUDA device.
"""
__version__ = "0.1.0"

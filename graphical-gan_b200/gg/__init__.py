"""Host side of the B200-native Graphical-GAN hot path: graph IR (graph.py, ops.py), plan compiler / CUDA-graph
executor (executor.py), ctypes binding of the C-ABI (cabi.py), data-parallel plumbing (dist.py)."""

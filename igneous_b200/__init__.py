"""igneous_b200 -- B200-native (sm_100a) implementation of the igneous
per-chunk hot path: DownsampleTask pooling, 6-connected CCL, MeshTask
marching cubes, behind igneous's own task API.

Importing this package never touches the GPU; the native library is loaded on
first use (`igneous_b200._shim.load()`), and compute calls raise if it or a
CUDA device is missing -- there is no CPU fallback.
"""
__version__ = "0.1.0"

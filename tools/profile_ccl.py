"""One CCL call for ncu: python tools/profile_ccl.py [size] [in_dtype]"""
import ctypes as c, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from igneous_b200 import _shim
size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dtype = np.dtype(sys.argv[2]) if len(sys.argv) > 2 else np.dtype(np.uint32)
ctx = _shim.default_context()
n = size ** 3
d_in = ctx.alloc(n * dtype.itemsize); d_out = ctx.alloc(n * 8)
code = _shim.dtype_code(dtype)
_shim.check(ctx.lib.ign_synth_seg_dev(ctx.handle, _shim.ptr(d_in), c.c_int(code), c.c_uint64(size), c.c_uint64(size), c.c_uint64(size), c.c_int64(0), c.c_int64(0), c.c_int64(0), c.c_uint32(64), c.c_uint64(1 << 20), c.c_uint64(0), c.c_uint64(0)))
N = c.c_uint64(0)
for _ in range(2):
  _shim.check(ctx.lib.ign_ccl6_dev(ctx.handle, _shim.ptr(d_in), c.c_int(code), c.c_uint64(size), c.c_uint64(size), c.c_uint64(size), _shim.ptr(d_out), c.c_int(_shim.IGN_U64), c.byref(N)))
ctx.sync()
print("N", N.value)

"""Host-side cost of one RandomVariable operation (Python mirror + ctypes + C ABI + launch), measured on tiny vectors so that the
GPU is never the bottleneck (default), or on 1 M-element vectors (argument 1000000) where "incl. drain" is the device time per operation.
Prints microseconds per operation for each layer."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
nv = pkg.native
nv.init(0)
import numpy as np  # noqa: E402

SIZE = int(sys.argv[1]) if len(sys.argv) > 1 else 1000         # vector length: 1000 = pure host cost, 1000000 = the C5 per-GPU size
N = 20000 if SIZE <= 10000 else 3000
x = pkg.RandomVariableCuda(0.0, np.arange(float(SIZE)) + 1.0)
y = pkg.RandomVariableCuda(0.0, np.arange(float(SIZE)) + 2.0)


def bench(name, fn):
    fn()
    nv.synchronize()
    t0 = time.perf_counter()
    for _ in range(N):
        fn()
    t1 = time.perf_counter()
    nv.synchronize()
    t2 = time.perf_counter()
    print("%-44s %6.2f us/op issued, %6.2f us/op incl. drain" % (name, (t1 - t0) / N * 1e6, (t2 - t0) / N * 1e6))


lib = nv.load()
out = C.c_uint64()


def raw_unary():
    lib.fmb_rv_unary(nv.U_SQUARED, x.dv.h, 0.0, C.byref(out))
    lib.fmb_rv_free(out.value)


bench("C ABI unary + free (two ctypes calls)", raw_unary)
bench("native.unary (DeviceVector, __del__ frees)", lambda: nv.unary(nv.U_SQUARED, x.dv))
bench("RandomVariableCuda.squared()", lambda: x.squared())
bench("RandomVariableCuda.mult(2.0)", lambda: x.mult(2.0))
bench("RandomVariableCuda.mult(rv)", lambda: x.mult(y))
bench("RandomVariableCuda.addProduct(rv, rv)", lambda: x.addProduct(y, y))
bench("RandomVariableCuda.addProduct(rv, 2.0)", lambda: x.addProduct(y, 2.0))
bench("x.sub(0.03).mult(0.5).div(y)  (3 ops)", lambda: x.sub(0.03).mult(0.5).div(y))
bench("getAverage() (reduction + D2H)", lambda: x.getAverage())

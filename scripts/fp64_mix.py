"""FP64 DFMA throughput with 0/2/4/8 integer multiply-adds per 8 DFMAs (issue-slot pressure probe)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ad_mpc_b200 import _lib
L = _lib.load()
for n in (0, 2, 4, 8, 16):
    v = C.c_double()
    _lib.check(L.admpc_measure_fp64_mix(0, n, C.byref(v)), "mix")
    print("int ops per 8 DFMA: %d -> %.2f TFLOP/s" % (n, v.value))

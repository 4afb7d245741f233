"""Phase stamps of tcg_sparse_kernel (timeline build) on the C5 pose-graph workload: CTA 0, fourth iteration."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("OB200_TIMELINE", "5")
import bench
from optimization_b200.device import Context
ctx = Context(0)
ctx.lib.ob200_debug_phase_times(ctx.h, 1, None, None)
bench.so3_c5(ctx)
mx = (C.c_uint64 * 4)(); mn = (C.c_uint64 * 4)()
ctx.lib.ob200_debug_phase_times(ctx.h, 0, mx, mn)

"""optimization_b200: B200-native fused Steihaug-Toint truncated-CG path of
david-m-rosen/Optimization (reference IterativeSolvers.h:166-426 driven by
TNT.h:242-689) behind a C ABI (include/optimization_b200.h).

`optimization_b200.capi` binds the shared library; `optimization_b200.device`
is the Python host handle; `optimization_b200.problems` holds the synthetic
problem generators (pure numpy).  Importing `device` requires the built
library and a GPU: there is no CPU fallback.
"""
__all__ = ["capi", "device", "problems"]

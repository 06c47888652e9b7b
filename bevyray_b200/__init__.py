"""bevyray_b200 — B200-native (sm_100a) drop-in for bevyray's path-tracing hot path.

The product is libbevyray_b200.so (CUDA kernels + C ABI + C++ host layer); this package is the thin
ctypes binding used by the tests and the benchmark."""
from . import _capi as capi  # noqa: F401  (raises ImportError when the library is not built)
from .api import (BVH_NODE_DTYPE, INF, MATERIAL_DTYPE, MISS_ID, MODEL_DTYPE, BvrError, Context, Scene,  # noqa: F401
                  build_ploc, make_camera, make_level, make_options, make_window, traversal_ranks, validate_bvh)

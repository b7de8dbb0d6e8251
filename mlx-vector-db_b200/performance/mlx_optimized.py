"""Import-path shim for the reference's performance/mlx_optimized.py function surface."""
from b200vs.ops import *  # noqa: F401,F403
from b200vs.ops import (PerformanceMonitor, performance_monitor,  # noqa: F401
                        warmup_compiled_functions)

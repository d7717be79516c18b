"""smpl_b200 -- B200-native validity + heuristic hot path behind smpl's plugin interfaces.

Layout
  csrc/   hand-written sm_100a kernels + the C ABI (include/smplgpu.h) -> lib/libsmplgpu.so
  host/   host-side C++: model builder, adapters mirroring the reference's
          CollisionChecker / RobotHeuristic / RobotModel interfaces -> lib/libsmplhost.so
  api.py  ctypes plumbing used by tests/ and bench.py
  scenes.py synthetic scenes / query sets of the shapes BASELINE.json names
"""
from .api import GpuContext, RobotTables, SmplGpuError, build_tables, scene_cells, setup_context, world_to_grid  # noqa: F401

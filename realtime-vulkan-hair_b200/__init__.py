"""realtime-vulkan-hair_b200: B200-native guide-strand physics step (FTL / PBD compute pass of
clach/Realtime-Vulkan-Hair) behind a C ABI (include/rvh.h, librvh.so).

This Python layer is a thin ctypes harness for tests and benchmarks; the product is the
CUDA library.  There is no CPU fallback: importing works without a GPU (so symbols can be
checked), creating a simulation does not.
"""
from .binding import (  # noqa: F401
    GRID_ON, WIND_A, WIND_B, GRID_INT32_WRAP, KEEP_CORRECTION, KEEP_ORDER, SDF_ON, REPULSION_ON, SDF_TMA,
    RvhConfig, RvhError, HairSim, load_library, library_path, default_config,
    collider_build, collider_translate, wind_fbm, nccl_unique_id, EXPORTED_SYMBOLS,
)
from . import scenes  # noqa: F401

"""cell_image_analysis_b200 -- sm_100a implementation of the per-cell screening hot
path of Kmatsuo57/cell-image-analysis (improved_detection.py:66-111, 117-153).

Layout: ``csrc/`` CUDA kernels + the C-ABI (``include/cia.h``) built into
``libcia.so``; ``_lib`` ctypes binding; ``artifacts`` (.keras / .pkl loading);
``screening`` the host-side mirror of ``ProductionMutantScreening``; ``distributed``
field sharding + the per-strain all-reduce; ``synth`` seeded synthetic fields.
"""
from . import synth  # noqa: F401  (pure NumPy)

__all__ = ["ProductionMutantScreening", "Engine", "synth"]


def __getattr__(name):
    # torch and the CUDA library are only needed by the engine classes
    if name in ("ProductionMutantScreening", "Engine", "PRECISION_FP32", "PRECISION_TC"):
        from . import screening
        return getattr(screening, name)
    raise AttributeError(name)

"""pymra_b200: B200-native (sm_100a) implementation of pyMRA's MRATree hot path.

    from pymra_b200.MRATree import MRATree
    import pymra_b200.MRATools as mt
"""
__version__ = "0.1"

# One process per GPU on a shared host: divide the host threads of the tree builder / planner between the local
# ranks (torchrun exports LOCAL_WORLD_SIZE) unless the user already chose MRA_HOST_THREADS.
import os as _os
if "MRA_HOST_THREADS" not in _os.environ and _os.environ.get("LOCAL_WORLD_SIZE", "1").isdigit():
    _lw = max(1, int(_os.environ.get("LOCAL_WORLD_SIZE", "1")))
    if _lw > 1:
        _os.environ["MRA_HOST_THREADS"] = str(max(1, (_os.cpu_count() or 8) // _lw - 1))

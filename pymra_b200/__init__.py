"""pymra_b200: B200-native (sm_100a) implementation of pyMRA's MRATree hot path.

    from pymra_b200.MRATree import MRATree
    import pymra_b200.MRATools as mt
"""
__version__ = "0.1"

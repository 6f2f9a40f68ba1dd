"""Empty import stub (the reference imports matplotlib at module scope,
MRATree.py:6-8, MRATools.py:7-8; nothing on the hot path uses it)."""

#!/usr/bin/env python
"""Parity of the CUDA path against the oracle port at the sizes BASELINE.json quotes (run on the GPU box):

    python tools/parity_fullsize.py cfg4 g700_r16_m7 cfg3 cfg5_exp cfg5 > gpurun_out/r04_parity_fullsize.jsonl

One JSON line per case: achieved likelihood / mean / sd errors, port seconds and peak RSS, GPU e2e seconds.
The port is test infrastructure (oracle/); this tool is a checker, not a product path."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import logging  # noqa: E402

logging.getLogger("pymra_b200.MRATree").setLevel(logging.ERROR)

from _util import fullsize_parity  # noqa: E402

if __name__ == "__main__":
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    for case in sys.argv[1:]:
        rec = fullsize_parity(case)
        rec["host_cores"] = os.cpu_count()
        print(json.dumps(rec), flush=True)

#!/usr/bin/env python3
"""Record the exception type and text the UNMODIFIED reference raises for invalid arguments of the mirrored
functions:  python tests/golden/make_golden_errors.py  -> tests/golden/errors.json
(checked by tests/test_layout_cpu.py::test_argument_errors_mirror_reference; the cases live in
tests/golden_io.py so that the test replays exactly the same calls)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.environ.get("OPTPRICER_REF", "/root/reference/src"))
import optpricer  # noqa: E402
from golden_io import error_cases  # noqa: E402

out = {}
for name, call in error_cases(optpricer):
    try:
        call()
        out[name] = None
    except Exception as e:  # noqa: BLE001
        out[name] = [type(e).__name__, str(e)]
with open(os.path.join(HERE, "errors.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))

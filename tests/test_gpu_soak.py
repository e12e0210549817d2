"""GPU: randomised soak of the search path (tools/soak_search.py) — random sizes, dims, k, segment
sizes, add splits and eight data families (heavy-tailed norms, duplicates, planted huge rows,
integer near-duplicates, anisotropic, tiny / large scales) against a float64 brute force."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_randomised_search_soak():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "soak_search.py"), "40", "11"], capture_output=True, text=True,
                       timeout=900)
    line = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert line, p.stdout[-2000:] + p.stderr[-2000:]
    res = json.loads(line[-1])
    assert p.returncode == 0 and res["failures"] == 0 and res["agg"]["flagged"] == 0, res

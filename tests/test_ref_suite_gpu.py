"""The reference's own 64 tests (tests/ref_suite/, unmodified copies of /root/reference/tests/*.py) run against the
drop-in alias ``import isoext`` -> isoext_b200.  Includes the reference's UNFILTERED sparse population recipe
(tests/conftest.py:39-61 of the reference feeds X*Y*Z candidate ids, i.e. ids past the (X-1)(Y-1)(Z-1) cell range,
through get_points_by_cell_indices / filter_cell_indices / add_cells)."""
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
SUITE = ROOT / "tests" / "ref_suite"


def test_reference_suite_passes_unchanged(iso):
    env = dict(os.environ, PYTHONPATH=str(ROOT) + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "pytest", str(SUITE), "-q", "-p", "no:cacheprovider", f"--rootdir={SUITE}",
                        f"--confcutdir={SUITE}", "-c", os.devnull], cwd=str(SUITE), env=env, capture_output=True, text=True,
                       timeout=1500)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) == 64 and "failed" not in r.stdout and "skipped" not in r.stdout, tail

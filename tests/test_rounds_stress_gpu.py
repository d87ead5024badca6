"""Race regression for the lock-step rounds (csrc/rounds.cuh): the once-per-call evaluation (leaf de-duplication +
evaluation cache) changes which slot waits for which batch row and when, never a game.  The same self-play call runs in
child processes with both mechanisms off (once) and on (three times): all four must hash to the same actions, root
counts and samples (scripts/stress_dedup.py; a single de-duplication table instead of one per round parity failed this
about once in ten runs)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_repeated_runs_are_identical_with_and_without_dedup_and_cache(azb):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "stress_dedup.py"), "2048", "80", "1", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.strip().splitlines()[-1] == "IDENTICAL"

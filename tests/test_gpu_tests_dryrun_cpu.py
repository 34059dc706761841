"""Every `-m gpu` test (validated and gated) executed on the CPU with a no-compute library (tools/gated_tests_dryrun.py --all): a
Python-level error in a GPU test or in the host / oracle code it drives (a missing attribute, a changed signature) must not wait for a GPU
box to be found -- the GPU suite runs with -x, so one such error hides every test after it."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_python_level_errors_in_the_gpu_tests():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'gated_tests_dryrun.py'), '--all'], capture_output=True, text=True, timeout=1500)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    errors = [ln for ln in out.splitlines() if ln.startswith('PYTHON ERROR')]
    assert not errors, '\n'.join(errors)
    assert 'python errors: 0' in out
    assert out.count('numeric stage') + out.count('ran through') > 60

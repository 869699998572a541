"""Development aid: print the per-function parity report (see tests/povar_gpu_checks.py)."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
runpy.run_path(os.path.join(ROOT, "tests", "povar_gpu_checks.py"), run_name="__main__")

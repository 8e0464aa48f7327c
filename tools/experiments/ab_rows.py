"""EXPERIMENT RECORD (profiles/r02_k_rows2_experiment.md): the CCO_ROWS_IMPL / CCO_V2_TUNE switches this script sets existed
only while tools/experiments/cco_rows2.cuh was compiled into the library; kept to show how the A/B numbers were taken.
Development: A/B of the row-kernel generations on one workload (CCO_ROWS_IMPL is read once per process).
usage: python tools/ab_rows.py [workload=C3] [trains=4]"""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
n = sys.argv[2] if len(sys.argv) > 2 else "4"
for impl in ("1", "2"):
    env = dict(os.environ, CCO_ROWS_IMPL=impl)
    print(f"==== CCO_ROWS_IMPL={impl}", flush=True)
    subprocess.run([sys.executable, os.path.join(here, "prof_rows.py"), wl, n], env=env)

#!/usr/bin/env python3
"""Run a Python script inside the environment of the reference runtime snapshot (oracle/_ref):
    python oracle/run_ref.py <script.py> [args...]
Test infrastructure (fixture generators, reference timing); never used by the product."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_env import reference_env

env = reference_env()
if env is None:
    sys.exit("no reference runtime under oracle/_ref (run oracle/build_ref.sh)")
sys.exit(subprocess.call([sys.executable] + sys.argv[1:], env=env))

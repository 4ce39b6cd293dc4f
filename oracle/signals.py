"""Re-export of the synthetic signal generators (they live in pycusdr_b200.benchmark.signals so that
bench.py can build its workload without importing the oracle)."""
from pycusdr_b200.benchmark.signals import *  # noqa: F401,F403
from pycusdr_b200.benchmark.signals import MODULATORS, BENCH_BW  # noqa: F401

"""Synthetic workloads mirroring the reference's examples/benchmark (signal generators only)."""

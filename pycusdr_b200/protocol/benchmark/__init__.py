from .bench_protocols import Bench_base, Bench_GMSK, Bench_FSK, Bench_GFSK, Bench_BPSK

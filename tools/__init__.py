"""Developer measurement scripts (sweeps, ncu drivers, workload replays, sanitizer workload).

Not part of the product: nothing under crypto_b200/ imports from here.  Like the tests, these scripts
use oracle/ only as an input generator (seeded scalars, k_i * G bases), as the checker of every GPU
result they time, and - in replay_workloads.py / wire_bench.py - as the timed CPU baseline next to the
GPU number.  The product library (crypto_b200/libdockgpu.so) never links or calls the oracle."""

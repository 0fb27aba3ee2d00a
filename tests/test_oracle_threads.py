"""Thread-count calibration of the CPU baseline arms (oracle/oracle.py): test infrastructure for bench.py."""
from oracle import oracle
from registrar_b200 import synth


def test_usable_cpus_is_positive_and_bounded(built):
    import os
    n = oracle.usable_cpus()
    assert 1 <= n <= (os.cpu_count() or 1)


def test_calibration_picks_a_measured_candidate(built):
    batch = synth.generate("config2", n=20_000)
    threads, rates = oracle.calibrate_threads(batch, sample=20_000, seconds=0.02)
    assert threads in rates and all(v > 0 for v in rates.values())
    # near-ties go to the smaller count: nothing smaller is within 7 % of the best rate
    best = max(rates.values())
    assert all(v < 0.93 * best for t, v in rates.items() if t < threads)

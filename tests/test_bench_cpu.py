"""CPU: the CPU legs of bench.py (cpu_baseline / --impl reference) run the oracle chain of every configuration."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.mark.parametrize("config,unit", [("c1", "monocular frames/s"), ("c2", "stereo frames/s"), ("c3", "stereo frames/s")])
def test_cpu_streams_run_the_chain_of_the_configuration(config, unit):
    import bench
    import oracle
    try:
        cfg = bench.set_config(config)
        assert bench.UNIT == unit and bench.MONO == (config == "c1") and bench.INERTIAL == (config == "c3")
        assert cfg["streams"] % 148 == 0                      # a multiple of the SM count (bench.py CONFIGS)
        cs = bench._CpuStreams(npool=1, n_map=600)
        ex = (oracle.Extractor(bench.NFEAT, 1.2, bench.NLEVELS, 20, 7), oracle.Extractor(bench.NFEAT, 1.2, bench.NLEVELS, 20, 7))
        for i in range(2):                                    # c3: frame 0 runs LastKeyFrame, frame 1 LastFrame
            T, st = cs.frame(i, ex)
            assert st[0] > 900 and st[6] > 300 and np.isfinite(T).all()
            assert (st[1] == 0) == (config == "c1")
            assert np.abs(T[:3, 3] - cs.Tt[0][:3, 3]).max() < 5e-2
    finally:
        bench.set_config("c2")


def test_algorithmic_bytes_match_survey_8d():
    import bench
    bench.set_config("c2")
    a = bench.algorithmic_bytes(1000)
    assert a["pyramid"] == 1845634 and a["fast"] == 1117367 and a["blur"] == 2 * 1117367   # SURVEY.md 8(d), DESIGN.md 4
    assert abs(a["total"] - 7179735) < 3000

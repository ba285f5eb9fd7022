"""The lane-level model of the CUDA SCL kernel (tests/model_scl_lanes.py) must reproduce the
oracle's final path list exactly — this checks the kernel DESIGN (slot pointers without refcounts,
bit-packed partial sums, stable candidate ranking) on the CPU."""
import numpy as np
import pytest
from _inputs import awgn_llr_set, detector_like_llr_set
from _polar_host import polar_transform
from echoseal_b200.polar_tables import frozen_mask, data_positions
from model_scl_lanes import decode_lanes
from oracle import polar_oracle as po


@pytest.mark.parametrize("kind,L", [("awgn", 8), ("awgn", 4), ("awgn", 1), ("tie", 8)])
def test_lane_model_matches_oracle(kind, L):
    fr = frozen_mask()
    pos = data_positions()
    if kind == "awgn":
        llr, _ = awgn_llr_set(8, seed=7)
        llr = llr[[1, 2, 3]]          # sigma 0.3, 0.4, 0.5 (list decoder proper)
    else:
        llr = detector_like_llr_set(48, seed=11)[[0, 30]]
    res = po.scl_batch(llr, L=L)
    for w in range(llr.shape[0]):
        paths = decode_lanes(llr[w].astype(np.float64), fr, L)
        assert len(paths) == int(res["npaths"][w])
        for a, (metric, xhat) in enumerate(paths):
            u = polar_transform(xhat)
            assert (u[fr] == 0).all()
            assert (u[pos][:440] == res["path_info"][w, a]).all(), (w, a)
            assert metric == res["path_metric"][w, a], (w, a)

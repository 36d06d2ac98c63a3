"""CPU: the C++ restatement of DetectionMatching against the reference's OWN
det_matching.cc (compiled unmodified into oracle/_ref by oracle/Makefile)."""
import numpy as np
import pytest

from oracle import det_matching_oracle as dm


def random_case(rs, n, g, tie_scores=False, tie_iou=False):
    iou = rs.uniform(0, 1, (n, g)).astype(np.float32)
    iou[rs.uniform(0, 1, (n, g)) < 0.6] = 0.0
    if tie_iou:
        iou = np.round(iou * 8) / np.float32(8)
    score = rs.uniform(0, 1, n).astype(np.float32)
    if tie_scores:
        score = np.round(score * 10) / np.float32(10)
    ignore = rs.uniform(0, 1, g) < 0.3
    return iou, score, ignore


@pytest.mark.parametrize('tie_scores,tie_iou', [(False, False), (True, False), (False, True),
                                               (True, True)])
def test_restatement_equals_reference_build(oracle_built, tie_scores, tie_iou):
    if not dm.have_reference_build():
        pytest.skip('oracle/_ref not built (no /root/reference here)')
    rs = np.random.RandomState(123)
    for n, g in [(0, 0), (5, 0), (0, 4), (1, 1), (17, 3), (40, 40), (300, 12), (1000, 40),
                 (2500, 97)]:
        for _ in range(3):
            iou, score, ignore = random_case(rs, n, g, tie_scores, tie_iou)
            got = dm.detection_matching(iou, score, ignore)
            ref = dm.ref_detection_matching(iou, score, ignore)
            for a, b in zip(got, ref):
                assert np.array_equal(a, b)
            assert got[0].dtype == np.float32 and got[2].dtype == np.int32


def test_reference_build_rejects_bad_shapes(oracle_built):
    if not dm.have_reference_build():
        pytest.skip('oracle/_ref not built')
    with pytest.raises(ValueError):
        dm.ref_detection_matching(np.zeros((3, 2), np.float32), np.zeros(4, np.float32),
                                  np.zeros(2, bool))

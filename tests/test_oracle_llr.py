"""The oracle's LogLikelihood against Mahout's own known-answer values and basic identities."""
import math

import pytest

from conftest import load_golden


def test_llr_known_answers(orc):
    for k11, k12, k21, k22, want in load_golden("llr_kats.json")["kats"]:
        for flags in (0, orc.FLAG_ENTROPY_VARARGS):
            got = orc.llr(k11, k12, k21, k22, flags)
            assert got == pytest.approx(want, rel=5e-7), (k11, k12, k21, k22)


def test_llr_survey_worked_cells(orc):
    # SURVEY.md Appendix B.1 worked cells
    assert orc.llr(1, 0, 0, 2) == pytest.approx(3.819085009768877, rel=1e-14)
    assert orc.llr(1, 1, 0, 1) == pytest.approx(1.046496287529096, rel=1e-14)
    assert orc.llr(2, 0, 0, 1) == pytest.approx(3.819085009768877, rel=1e-14)
    assert orc.llr(3, 0, 0, 0) == 0.0


def test_llr_matches_python_restatement(orc):
    import random
    rng = random.Random(5)
    for _ in range(2000):
        k = [rng.randrange(0, 10 ** rng.randrange(1, 7)) for _ in range(4)]
        assert orc.llr(*k) == orc.llr_py(*k)          # same formula, same libm -> identical bits


def test_llr_properties(orc):
    assert orc.xlogx(0) == 0.0
    assert orc.xlogx(1) == 0.0
    assert orc.xlogx(7) == 7 * math.log(7)
    # symmetric under swapping rows / columns of the 2x2 table (up to round-off)
    assert orc.llr(5, 7, 11, 1000) == pytest.approx(orc.llr(11, 1000, 5, 7), rel=1e-9)
    assert orc.llr(5, 7, 11, 1000) == pytest.approx(orc.llr(7, 5, 1000, 11), rel=1e-9)
    # independence -> ~0, never negative
    for n in (10, 1000, 10 ** 6):
        v = orc.llr(n, n, n, n)
        assert 0.0 <= v < 1e-6
    # Preconditions.checkArgument(k >= 0)
    assert math.isnan(orc.llr(-1, 1, 1, 1))


def test_entropy_order_variants_agree_to_roundoff(orc):
    import random
    rng = random.Random(11)
    for _ in range(500):
        n = 10 ** rng.randrange(3, 8)
        ra, cb = rng.randrange(1, 500), rng.randrange(1, 500)
        k11 = rng.randrange(1, min(ra, cb) + 1)
        a = orc.llr(k11, ra - k11, cb - k11, n - ra - cb + k11, 0)
        b = orc.llr(k11, ra - k11, cb - k11, n - ra - cb + k11, orc.FLAG_ENTROPY_VARARGS)
        assert a == pytest.approx(b, rel=1e-6, abs=1e-6)


def test_dominance_property_behind_the_device_filter(orc):
    """DESIGN.md 3.1 "dominance filter": on the positively associated side (rowA*colB < k11*N) the LLR grows with k11
    and shrinks with colB, so a cell (k', c') with k' <= k and c' >= c never scores above (k, c)."""
    import random
    rng = random.Random(21)
    checked = 0
    for _ in range(4000):
        n = 10 ** rng.randrange(3, 8)
        ra = rng.randrange(1, min(600, n // 2))
        c = rng.randrange(1, min(600, n // 2))
        k = rng.randrange(1, min(ra, c) + 1)
        kp = rng.randrange(1, k + 1)
        cp = rng.randrange(c, min(c + 50, n - ra) + 1)
        if kp > min(ra, cp) or not ra * cp < kp * n:          # the dominated cell must be on the positive side
            continue
        hi = orc.llr(k, ra - k, c - k, n - ra - c + k)
        lo = orc.llr(kp, ra - kp, cp - kp, n - ra - cp + kp)
        assert lo <= hi * (1 + 1e-12) + 1e-7, (n, ra, k, c, kp, cp)
        if (kp, cp) != (k, c):
            checked += 1
    assert checked > 1000

"""Host-side tap design twin (multirate.jl_b200/firdesign.py) against the reference's formulas
(src/FIRDesign.jl:18-95) and the numbers the survey verified."""
import numpy as np
import pytest

import multirate_b200 as mr
import multirate_oracle as mo


def test_kaiserlength_matches_reference_test_recipe():
    # test/runtests.jl:336-341 -- kaiserlength(0.05, samplerate=32) -> 2321 taps, beta 5.6533
    n, beta = mr.kaiserlength(0.05, samplerate=32)
    assert n == 2321 and abs(beta - 5.65326) < 1e-4
    assert mr.kaiserlength(0.1, 30)[1] == pytest.approx(0.5842 * 9 ** 0.4 + 0.07886 * 9)
    assert mr.kaiserlength(0.1, 10)[1] == 0.0
    assert (n, beta) == mo.kaiserlength(0.05, samplerate=32)


def test_lowpass_firdes_is_windowed_sinc_and_matches_the_oracle_twin():
    # README.md:177-179: firdes(numTaps, cutoff, kaiser, beta = 7.8562), numTaps = 24*147, cutoff = 0.5/147
    h = mr.firdes(24 * 147, 0.5 / 147, mr.kaiser, beta=7.8562)
    assert h.shape == (3528,) and np.allclose(h, h[::-1])
    M = 3527
    n = np.arange(3528) - M / 2
    assert np.allclose(h, 2 * (0.5 / 147) * np.sinc(2 * (0.5 / 147) * n) * np.kaiser(3528, 7.8562), rtol=0, atol=1e-18)
    assert np.array_equal(h, mo.firdes(3528, 0.5 / 147, 7.8562))
    assert abs(h.sum() - 1.0) < 1e-3                                   # unity DC gain
    # samplerate scales the cutoff (src/FIRDesign.jl:78)
    assert np.array_equal(mr.firdes(128, 11025.0, mr.kaiser, samplerate=44100.0, beta=5.0), mr.firdes(128, 0.25, mr.kaiser, beta=5.0))


def test_responses_and_second_method():
    lp = mr.firprototype(65, 0.2)
    hp = mr.firprototype(65, 0.2, response=mr.HIGHPASS)
    assert np.allclose(lp + hp, np.sinc(np.arange(65) - 32))            # complementary
    assert len(mr.firprototype(64, 0.2, response=mr.HIGHPASS)) == 65   # made type 1 (src/FIRDesign.jl:55)
    bp = mr.firprototype(65, [0.3, 0.1], response=mr.BANDPASS)
    bs = mr.firprototype(65, [0.3, 0.1], response=mr.BANDSTOP)
    assert np.allclose(bp, -bs)
    assert np.allclose(bp, mr.firprototype(65, 0.3) - mr.firprototype(65, 0.1))
    h = mr.firdes(0.25, 0.05, 60)                                        # firdes(cutoff, transitionwidth, attenuation)
    n, beta = mr.kaiserlength(0.05, 60)
    assert len(h) == n and np.array_equal(h, mr.firdes(n, 0.25, mr.kaiser, beta=beta))
    assert np.array_equal(mr.firdes(33, 0.2, mr.hamming), mr.firprototype(33, 0.2) * np.hamming(33))
    with pytest.raises(TypeError):
        mr.firdes(1)

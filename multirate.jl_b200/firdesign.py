"""Tap design on the host: the Python twin of the reference's FIRDesign.jl.

`kaiserlength` (src/FIRDesign.jl:18-32), `firprototype` (:49-65) and both `firdes` methods (:76-95).  Taps are an
INPUT of the filtering path, so nothing here runs on the device; it exists so that the README / example calls
(`firdes(numTaps, cutoff, kaiser, beta=...)`, README.md:177-179) have a drop-in on this side too.

Window convention: the reference takes its windows from DSP.jl (`using DSP.Windows`, src/Multirate.jl:9), which is
not vendored under /root/reference.  `kaiser(n, beta)` here is numpy's (beta is the textbook Kaiser beta, the value
`kaiserlength` returns); DSP.jl's `kaiser(n, alpha)` of that era took alpha = beta/pi -- pass beta/pi there.
"""
import enum
import math

import numpy as np


class FIRResponse(enum.Enum):                  # @enum(FIRResponse, ...), src/FIRDesign.jl:7
    LOWPASS = 0
    BANDPASS = 1
    HIGHPASS = 2
    BANDSTOP = 3


LOWPASS, BANDPASS, HIGHPASS, BANDSTOP = FIRResponse


def kaiser(n, beta):
    return np.kaiser(int(n), float(beta))


def hanning(n):
    return np.hanning(int(n))


def hamming(n):
    return np.hamming(int(n))


def blackman(n):
    return np.blackman(int(n))


def kaiserlength(transition, attenuation=60, samplerate=1.0):
    """(numtaps, beta) of a Kaiser-window design, src/FIRDesign.jl:18-32."""
    transition = transition / samplerate
    numtaps = int(math.ceil((attenuation - 7.95) / (2 * math.pi * 2.285 * transition)))
    if attenuation > 50:
        beta = 0.1102 * (attenuation - 8.7)
    elif attenuation >= 21:
        beta = 0.5842 * (attenuation - 21) ** 0.4 + 0.07886 * (attenuation - 21)
    else:
        beta = 0.0
    return numtaps, beta


def firprototype(numtaps, F, response=LOWPASS):
    """Ideal (unwindowed) impulse response, src/FIRDesign.jl:49-65.  F is a scalar for low / high pass, a pair for
    band pass / band stop; HIGHPASS may return one more tap to make the filter type 1 (:55)."""
    M = int(numtaps) - 1
    if response in (LOWPASS, HIGHPASS):
        F = float(F)
    else:
        F = (float(F[0]), float(F[1]))
    if response == HIGHPASS and M % 2 == 1:
        M += 1
    n = np.arange(M + 1, dtype=np.float64) - M / 2
    if response == LOWPASS:
        return 2 * F * np.sinc(2 * F * n)
    if response == BANDPASS:
        return 2 * (F[0] * np.sinc(2 * F[0] * n) - F[1] * np.sinc(2 * F[1] * n))
    if response == HIGHPASS:
        return np.sinc(n) - 2 * F * np.sinc(2 * F * n)
    if response == BANDSTOP:
        return 2 * (F[1] * np.sinc(2 * F[1] * n) - F[0] * np.sinc(2 * F[0] * n))
    raise ValueError("Not a valid FIR_TYPE")                                      # :61


def firdes(*args, response=LOWPASS, samplerate=1.0, beta=6.75):
    """firdes(numtaps, cutoff, windowfunction; response, samplerate, beta)          src/FIRDesign.jl:76-88
    firdes(cutoff, transitionwidth[, stopbandAttenuation=60]; response, samplerate)  src/FIRDesign.jl:90-95"""
    if len(args) == 3 and callable(args[2]):
        numtaps, cutoff, window = args
        cutoff = np.asarray(cutoff, dtype=np.float64) / samplerate
        proto = firprototype(numtaps, cutoff if cutoff.ndim else float(cutoff), response=response)
        n = len(proto)
        return proto * (kaiser(n, beta) if window is kaiser else window(n))
    if len(args) in (2, 3):
        cutoff, transitionwidth = args[0], args[1]
        att = args[2] if len(args) == 3 else 60
        numtaps, b = kaiserlength(transitionwidth, att, samplerate=samplerate)
        return firdes(numtaps, cutoff, kaiser, response=response, samplerate=samplerate, beta=b)
    raise TypeError("firdes(numtaps, cutoff, windowfunction; ...) or firdes(cutoff, transitionwidth[, attenuation]; ...)")

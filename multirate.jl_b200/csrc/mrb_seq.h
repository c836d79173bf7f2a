// mrb_seq.h -- host-side sequencing: output counts, phase/deficit carry and the exact
// replay of the arbitrary-rate phase accumulators.  Pure C++ (no CUDA), shared by the
// API layer; everything here is data independent and channel independent.
//
// Reference: src/Filters.jl:352-385 (outputlength), :433-439 (nextphase), :536-575
// (rational loop), :598-631 (decimator loop), :663-673 / :780-786 (update).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace mrb {

static inline int64_t ceil_div(int64_t a, int64_t b) {   // b > 0
    return a >= 0 ? (a + b - 1) / b : -((-a) / b);
}
static inline int64_t floor_div(int64_t a, int64_t b) {  // b > 0
    return a >= 0 ? a / b : -((-a + b - 1) / b);
}

// Integer schedule shared by standard (L=M=1), interpolator (M=1), decimator (L=1) and
// rational.  State: p = phiIdx-1 (0-based phase), d = inputDeficit (1-based).
// Output k (0-based) of a chunk reads the window ending at 1-based input index
//   n_k = d + floor((p + k*M)/L)          with branch  phi_k = (p + k*M) mod L.
// This is the closed form of the loop at src/Filters.jl:558-569 (:567 advances inputIdx by
// floor((phiIdx+M-1)/L), :568 is nextphase); tests/test_host_logic.py checks it against the
// literal loop of the oracle.
struct IntSeq {
    int64_t L, M;
    // number of outputs for xLen inputs from state (p, d); src/Filters.jl:352-357,371-373
    static int64_t count(int64_t L, int64_t M, int64_t p, int64_t d, int64_t xLen) {
        if (xLen < d) return 0;
        return ceil_div((xLen - d + 1) * L - p, M);
    }
    // state after the chunk
    static void advance(int64_t L, int64_t M, int64_t &p, int64_t &d, int64_t xLen) {
        if (xLen < d) { d -= xLen; return; }                     // :543-547
        const int64_t N = count(L, M, p, d, xLen);
        const int64_t t = p + N * M;
        d = d + t / L - xLen;                                     // :571
        p = t % L;
    }
};

// FIRArbitrary.update / FIRFarrow.update, src/Filters.jl:663-673, 780-786.  Sequentially
// rounded Float64 recurrence: replayed literally (compile with -ffp-contract=off).
struct ArbState {
    double acc;     // phiAccumulator (arbitrary) / Float64 phiIdx (farrow), in [1, Nphi+1)
    int64_t xIdx;   // 1-based
};
static inline void arb_update(ArbState &s, double delta, int64_t Nphi) {
    s.acc += delta;
    const double nphi = (double)Nphi;
    if (s.acc > nphi) {
        const double t = s.acc - 1.0;
        s.xIdx += (int64_t)(t / nphi);                           // the reference floors the ROUNDED quotient (t > 0)
        // mod(t, Nphi), exact and bit-identical to fmod on every path (checked against fmod over 8e7 updates of random
        // rates and Nphi): t - q*Nphi is exact for every integer q with q*Nphi <= t, because Nphi is an integer --
        // hence a multiple of ulp(t) -- and the difference is no larger than t.  The common quotients are compared
        // for; otherwise q is estimated with a multiplication and an estimate that is off by one is repaired, exactly
        // for the same reason.  fmod itself costs ~50 ns and sat on 9 % of the updates of a 0.92 resampler.
        double r;
        if (t < nphi) r = t;
        else if (t < 2.0 * nphi) r = t - nphi;
        else if (t < 3.0 * nphi) r = t - 2.0 * nphi;
        else if (t < 9.0e15) {
            r = t - (double)(int64_t)(t * (1.0 / nphi)) * nphi;
            if (r < 0.0) r += nphi; else if (r >= nphi) r -= nphi;
        } else {
            r = std::fmod(t, nphi);
        }
        s.acc = r + 1.0;
    }
}

}  // namespace mrb

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
// One update, bit-identical to the reference's
//     acc += delta; if acc > Nphi { xIdx += ifloor((acc-1)/Nphi); acc = mod(acc-1, Nphi) + 1 }
// (checked against the literal fmod form over 2e8 updates of random rates and Nphi, rational rates included), but with
// a dependency chain of two floating-point operations instead of five.  Why it is exact, with a1 = fl(acc + delta) > Nphi:
//  * t = a1 - 1 is exact (a1 >= 1 is a multiple of its ulp, and so is 1);
//  * q = floor(t / Nphi) as a real number is found by comparing a1 with 1 + q Nphi; t - q Nphi is exact (Nphi is an
//    integer, hence a multiple of ulp(t), and the difference is no larger than t);
//  * (t - q Nphi) + 1 = a1 - q Nphi is a multiple of ulp(a1) no larger than a1, hence representable: the reference's
//    final "+ 1" never rounds, and acc' = a1 - q Nphi is ONE exact subtraction.
// The reference floors the ROUNDED quotient fl(t / Nphi) for xIdx, which exceeds q only when t sits within rounding
// distance below a multiple of Nphi: there (a 2^-48 relative guard band) the division is done for real.
struct ArbStepper {
    double delta, nphi, k1, k2, k3, g1, g2, g3;
    ArbStepper(double delta_, int64_t Nphi) : delta(delta_), nphi((double)Nphi) {
        k1 = 1.0 + nphi; k2 = 1.0 + 2.0 * nphi; k3 = 1.0 + 3.0 * nphi;
        const double sh = 1.0 - 0x1p-48;
        g1 = nphi * sh; g2 = 2.0 * nphi * sh; g3 = 3.0 * nphi * sh;
    }
    inline void step(ArbState &s) const {
        const double a1 = s.acc + delta;
        if (a1 > nphi) {
            if (a1 < k3) {
                const double t = a1 - 1.0;
                if (a1 < k1) { s.acc = a1; s.xIdx += t >= g1 ? (int64_t)(t / nphi) : 0; }
                else if (a1 < k2) { s.acc = a1 - nphi; s.xIdx += t >= g2 ? (int64_t)(t / nphi) : 1; }
                else { s.acc = a1 - 2.0 * nphi; s.xIdx += t >= g3 ? (int64_t)(t / nphi) : 2; }
            } else {                                             // steep decimation: the general form
                const double t = a1 - 1.0;
                s.xIdx += (int64_t)(t / nphi);                   // the reference floors the ROUNDED quotient (t > 0)
                s.acc = std::fmod(t, nphi) + 1.0;
            }
        } else {
            s.acc = a1;
        }
    }
};

static inline void arb_update(ArbState &s, double delta, int64_t Nphi) { ArbStepper(delta, Nphi).step(s); }

}  // namespace mrb

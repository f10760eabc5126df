// Slot layout of the theta-function blocks in the per-ray workspace (see kg_theta.cuh).
#pragma once
namespace kg {

constexpr int kThJM = 8;   // stored harmonics per series

// slot offsets of one motion's block (relative to its base slot)
enum ThetaBaseSlot {
    TB_OK = 0,     // 1.0 when the theta path may be used for this motion
    TB_W,          // pi / (2K)
    TB_EOKM1,      // E/K - 1  (accurate for m -> 0)
    TB_J,          // number of harmonics to sum
    TB_R1, TB_R2, TB_R3,                 // th3(0)/th2(0), th4(0)/th2(0), th4(0)/th3(0)
    TB_R13, TB_R123,                     // R1/R3, R1 R2/R3 (polar motion: sn/dn and sn cn/dn with one reciprocal)
    TB_QE,                               // 2 (-1)^j q^(j^2), j = 1..JM
    TB_QO = TB_QE + kThJM,               // 2 (-1)^j q^((j+1/2)^2), j = 0..JM-1
    TB_SIZE = TB_QO + kThJM
};
// one Pi: Pi(n; am u | m) = TP_A * u + TP_P * L(v)
enum ThetaPiSlot {
    TP_KIND = 0,   // 0: real shift, even harmonics (0<n<m)   1: real shift, odd harmonics (n>1)
                   // 2: imaginary shift (m<n<1, n<0): L = atan2(Y, X)
    TP_A, TP_P,
    TP_C,                                // X coefficients
    TP_S = TP_C + kThJM,                 // Y coefficients
    TP_SIZE = TP_S + kThJM
};

}  // namespace kg

// SPDX-License-Identifier: MIT
//
// DEEP-quotient pieces (stwo-verifier/src/deep/quotients.simf:15-44) shared by the verifier's query kernel and the prover.
#pragma once
#include "field.cuh"

namespace ssym {

struct LineCoeffs {
    QM31 a, b, c;
};
// deep_quotient_interpolant_coefficients                      deep/quotients.simf:25-35
static __device__ __noinline__ LineCoeffs interpolant_coefficients(QM31 py, QM31 sv, QM31 alpha_i) {
    QM31 a = qm31c(cm31(0, 0), cm31_neg(cm31_dbl(sv.i)));
    QM31 b = qm31c(cm31(0, 0), cm31_neg(cm31_dbl(py.i)));
    QM31 a_py = qm31_mul(a, py);
    QM31 b_val = qm31_mul(b, sv);
    QM31 c = qm31_sub(b_val, a_py);
    LineCoeffs r;
    r.a = qm31_mul(alpha_i, a);
    r.b = qm31_mul(alpha_i, b);
    r.c = qm31_mul(alpha_i, c);
    return r;
}

__device__ __forceinline__ CM31 denominator_inverse(QM31 px, QM31 py, M31Point r, bool &fail) { // deep/quotients.simf:15-22
    CM31 dx = cm31_sub_m31(px.r, r.x);
    CM31 dy = cm31_sub_m31(py.r, r.y);
    CM31 d = cm31_sub(cm31_mul(dx, py.i), cm31_mul(dy, px.i));
    return cm31_inv(d, fail);
}

} // namespace ssym

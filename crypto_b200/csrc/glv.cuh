// GLV endomorphism of BLS12-381 shared by the MSM digit kernel (msm_kernels.cuh) and the batch scalar
// multiplication (batch_kernels.cuh).  Replaces nothing in the reference by name: arkworks' msm_bigint / mul_bigint
//  do not use it; the group elements produced are the same.
#pragma once
#include "fp2.cuh"

namespace dg {

// ---- GLV split ------------------------------------------------------------------------------------------------
// BLS12-381 has the endomorphism phi(x, y) = (beta x, y) acting on the prime-order subgroups as multiplication by
// lambda = -x^2 mod r (x the curve parameter; lambda^2 + lambda + 1 = r).  A scalar s' < r / 2 is written
// s' = q * x^2 + rem = q * x^2 +- k1 with k1 <= x^2 / 2, i.e.  [s'] P = +-[k1] P - [q] phi(P): two 127-bit scalars
// instead of one 254-bit one, so the windows only have to cover 128 bits -- half the bucket sets to reduce and half the
// doublings in the window combination, at the same number of bucket additions.
__device__ __forceinline__ void glv_split(const uint32_t sp[8], uint32_t k1[4], bool &neg1, uint32_t k2[4]) {
    constexpr uint32_t X2[4] = {DG_GLV_X2_0, DG_GLV_X2_1, DG_GLV_X2_2, DG_GLV_X2_3};
    constexpr uint32_t MU[5] = {DG_GLV_MU_0, DG_GLV_MU_1, DG_GLV_MU_2, DG_GLV_MU_3, DG_GLV_MU_4};       // floor(2^256 / x^2)
    constexpr uint32_t HALF[4] = {DG_GLV_HALF_0, DG_GLV_HALF_1, DG_GLV_HALF_2, DG_GLV_HALF_3};           // floor(x^2 / 2)
    // Barrett: q' = floor(sp * MU / 2^256) satisfies q - 2 <= q' <= q
    uint32_t q[5];
    {
        uint64_t lo = 0;
        uint32_t hi = 0;
#pragma unroll
        for (int col = 0; col < 13; col++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int j = col - i;
                if (j < 0 || j > 4) continue;
                uint64_t t = (uint64_t)sp[i] * MU[j];
                lo += t;
                hi += lo < t;
            }
            if (col >= 8) q[col - 8] = (uint32_t)lo;
            lo = (lo >> 32) | ((uint64_t)hi << 32);
            hi = 0;
        }
    }
    // rem = sp - q' * x^2, which is < 3 x^2 < 2^130: five limbs are enough
    uint32_t rem[5];
    {
        uint32_t qx[5];
        uint64_t lo = 0;
        uint32_t hi = 0;
#pragma unroll
        for (int col = 0; col < 5; col++) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int j = col - i;
                if (j < 0 || j > 3) continue;
                uint64_t t = (uint64_t)q[i] * X2[j];
                lo += t;
                hi += lo < t;
            }
            qx[col] = (uint32_t)lo;
            lo = (lo >> 32) | ((uint64_t)hi << 32);
            hi = 0;
        }
        uint32_t br = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            uint64_t d = (uint64_t)sp[k] - qx[k] - br;
            rem[k] = (uint32_t)d;
            br = (uint32_t)(d >> 63);
        }
    }
#pragma unroll
    for (int t = 0; t < 2; t++) {                                   // rem >= x^2: rem -= x^2, q += 1
        uint32_t d[5], br = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            uint64_t v = (uint64_t)rem[k] - (k < 4 ? X2[k] : 0u) - br;
            d[k] = (uint32_t)v;
            br = (uint32_t)(v >> 63);
        }
        const bool ge = br == 0;
        uint32_t cy = ge ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            rem[k] = ge ? d[k] : rem[k];
            uint64_t v = (uint64_t)q[k] + cy;
            q[k] = (uint32_t)v;
            cy = (uint32_t)(v >> 32);
        }
    }
    // balance: rem > x^2 / 2  ->  k1 = x^2 - rem with a minus sign, q += 1
    bool gt = false, eq = true;
#pragma unroll
    for (int k = 3; k >= 0; k--) {
        if (eq && rem[k] != HALF[k]) { gt = rem[k] > HALF[k]; eq = false; }
    }
    neg1 = gt;
    {
        uint32_t d[4], br = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint64_t v = (uint64_t)X2[k] - rem[k] - br;
            d[k] = (uint32_t)v;
            br = (uint32_t)(v >> 63);
        }
        uint32_t cy = gt ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            k1[k] = gt ? d[k] : rem[k];
            uint64_t v = (uint64_t)q[k] + cy;
            k2[k] = (uint32_t)v;
            cy = (uint32_t)(v >> 32);
        }
    }
}

// phi on a coordinate: beta is an Fp constant for both groups (DGC_GLV_BETA_G1 / _G2, oracle/gen_constants.py)
__device__ __forceinline__ Fp glv_mul_beta(const Fp &x, const Fp &beta) { return fp_mul(x, beta); }
__device__ __forceinline__ Fp2 glv_mul_beta(const Fp2 &x, const Fp &beta) { return {fp_mul(x.c0, beta), fp_mul(x.c1, beta)}; }

}  // namespace dg

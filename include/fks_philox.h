/*
 * fks_philox.h -- counter-based actuator noise shared by the device kernels and the CPU oracle.
 *
 * Replaces (in FKS_NOISE_PHILOX mode) the reference's per-thread std::mt19937_64 +
 * arc_helpers::TruncatedNormalDistribution draw (simple_uncertainty_models.hpp:86, generators
 * seeded at simple_particle_contact_simulator.hpp:431-441).  The distribution is the same --
 * Normal(0, sigma) truncated to [-1, 1], sampled by accept/reject on standard normals -- but the
 * stream is a pure function of (seed, particle, step, microstep, dof, attempt), so results do not
 * depend on thread count, GPU count or scheduling.
 *
 * Philox4x32-10 (Salmon et al., SC'11), Box-Muller on two 53-bit uniforms.
 */
#ifndef FKS_PHILOX_H
#define FKS_PHILOX_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FKS_HD __host__ __device__ __forceinline__
#else
#define FKS_HD static inline
#endif

FKS_HD void fks_mulhilo32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    *hi = (uint32_t)(p >> 32);
    *lo = (uint32_t)p;
}

FKS_HD void fks_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                              uint32_t* out) {
    for (int round = 0; round < 10; round++) {
        uint32_t hi0, lo0, hi1, lo1;
        fks_mulhilo32(0xD2511F53u, c0, &hi0, &lo0);
        fks_mulhilo32(0xCD9E8D57u, c2, &hi1, &lo1);
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n1 = lo1;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        const uint32_t n3 = lo0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

/* candidate number `attempt` of a draw: a standard normal variate by Box-Muller from one Philox block */
FKS_HD double fks_philox_normal_candidate(uint64_t seed, uint64_t particle, uint32_t step, uint32_t microstep, uint32_t dof,
                                          uint32_t attempt) {
    uint32_t r[4];
    fks_philox4x32_10((uint32_t)particle, (uint32_t)(particle >> 32), (step << 16) | (microstep & 0xFFFFu), (dof << 16) | attempt,
                      (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const uint64_t a = (((uint64_t)r[0] << 32) | r[1]) >> 11; /* 53 bits */
    const uint64_t b = (((uint64_t)r[2] << 32) | r[3]) >> 11;
    const double u1 = ((double)a + 1.0) * (1.0 / 9007199254740992.0); /* (0, 1] */
    const double u2 = (double)b * (1.0 / 9007199254740992.0);         /* [0, 1) */
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

/* One draw of sigma * z, z ~ N(0,1) conditioned on |sigma z| <= 1; 0 when sigma == 0: the first accepted candidate.
 * (Two candidates per pass side by side, so that a warp waits less often for its unluckiest lane, was measured 3-7 % slower.) */
FKS_HD double fks_philox_truncated_normal(uint64_t seed, uint64_t particle, uint32_t step, uint32_t microstep,
                                          uint32_t dof, double sigma) {
    double s = fabs(sigma);
    s = s > 1.0 ? 1.0 : s; /* ClampValue(|percent_variance|, 0, 1), unc.hpp:61 */
    if (s == 0.0) return 0.0;
    const double bound = 1.0 / s;
    double z = 0.0;
    for (uint32_t attempt = 0; attempt < 64u; attempt++) {
        z = fks_philox_normal_candidate(seed, particle, step, microstep, dof, attempt);
        if (z <= bound && z >= -bound) return s * z;
    }
    /* 64 consecutive rejections has probability < 1e-80 for sigma = 0.5; clamp to stay in range */
    z = z > bound ? bound : (z < -bound ? -bound : z);
    return s * z;
}

#endif /* FKS_PHILOX_H */

// Device-side building blocks of libptb_b200: counter-based RNG, variate sources and the physics of one
// particle at one plane.  Everything in here is written from the behavioural spec in SURVEY.md
// Appendix A; the reference lines each formula answers to are cited next to it.
//
// Arithmetic contract (exact build, -fmad=false): every expression is evaluated in IEEE double in the
// operation order of the reference's Numba code, so that with identical injected variates the results
// differ from the CPU only through the <=2 ulp of CUDA's tan / exp / pow against glibc's.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "ptb.h"

namespace ptb {

// ---------------------------------------------------------------------------------------------------
// configuration derived on the host (ptb_create) and passed as a __grid_constant__ kernel parameter,
// i.e. it lives in the constant bank and is read with uniform LDC
// ---------------------------------------------------------------------------------------------------
// geometry of one pixel axis (x / columns, y / rows): pitch, offset, number / 2, pitch / 2
struct AxisK {
    double pitch, delta, half_n, half_p;
};

struct PlaneK {
    double z;
    double x_pos, x_neg, y_pos, y_neg;   // acceptance window, main/hit.py:56-59
    double eps, corr2;                   // thickness/X0 and (1+0.038 ln eps)^2, main/hit.py:294-296
    double hl_c;                         // 13.6 * sqrt(eps * corr2): energy-independent factor of the Highland width
    double pitch_c, pitch_r, delta_x, delta_y;
    double rcp_pc, rcp_pr;               // RN(1/pitch), for div_rcp
    double half_nc, half_nr;             // number / 2 (true division), main/device.py:356,372
    double half_pc, half_pr;             // pitch / 2
    double thr_ke;                       // threshold * 1e-3, main/device.py:152
    double noise_sigma;                  // sqrt(k_b*300*C)/e, main/device.py:410-415
    double lan_loc, lan_scale;           // energy loss = loc + scale * lambda (Moyal variable), main/hit.py:399-400
    double lan_ulo, lan_uspan;           // window of the Moyal CDF that maps onto [0, 0.5] MeV, main/hit.py:99-100
    AxisK   ax[2];                       // the same geometry once more, by axis index (phase A of k_digitise)
    int32_t ncol, nrow, triggered;
    int32_t lan_direct;                  // 1: the window holds essentially all of the Moyal law: sample -ln(Z^2) directly
};

struct AliasEntry {
    double  prob;    // probability of keeping the drawn column
    int32_t alias;   // the column's alias otherwise
    int32_t _pad;
};

struct ConfigK {
    int32_t n_planes, pool_entries;
    double  energy, loc_x, sigma_x, loc_y, sigma_y, disp_x, disp_y, angle_x, angle_y;
    double  theta0_first;                // Highland width of plane 0 at beam energy: event 0 only, main/hit.py:75-80
    double  accept_p;                    // main/device.py:56
    // Poisson gap, main/hit.py:33,239
    double  lam, p_loglam, p_b, p_a, p_log_invalpha, p_vr, p_explam;
    // Walker alias table of the Poisson law on [alias_kmin, alias_kmin + alias_n) (device memory owned by the
    // handle; null when the law is too wide for a table and the rejection samplers are used instead)
    const struct AliasEntry *alias;
    long long                alias_kmin;
    int32_t                  alias_n, _pad1;
    // charge-cloud constants, main/device.py:186-202
    double  diff2, qe, c3mu, t90, cden;
    double  sigma0;                      // cloud width for a vanishing deposit
    double  rcp_3p6, rcp_cden;           // RN(1/3.6), RN(1/cden), for div_rcp
    double  exp2_tab[64];                // RN(2^(j/64)), for exp_nonpos (copied to shared memory by the digitisation kernels)
    PlaneK  pl[PTB_MAX_PLANES];
};

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  key = 64-bit seed, counter = (particle id lo, hi, site, index)
// ---------------------------------------------------------------------------------------------------
struct U4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo)
{
#ifdef __CUDA_ARCH__
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

// The ten round keys (key + r * Weyl constants) of a seed, worked out once on the host: as part of a kernel parameter
// they sit in the constant bank and enter the round's XOR as an operand, instead of being re-derived by every call.
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
};

__host__ __device__ inline PhiloxKeys philox_keys(uint64_t seed)
{
    PhiloxKeys k;
    uint32_t   a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        k.k0[r] = a;
        k.k1[r] = b;
        a += 0x9E3779B9u;
        b += 0xBB67AE85u;
    }
    return k;
}

__host__ __device__ __forceinline__ U4 philox4x32_10(U4 c, const PhiloxKeys &key)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0, l0, h1, l1;
        mulhilo(0xD2511F53u, c.x, h0, l0);
        mulhilo(0xCD9E8D57u, c.z, h1, l1);
        c = U4{h1 ^ c.y ^ key.k0[r], l1, h0 ^ c.w ^ key.k1[r], l0};
    }
    return c;
}

// Philox2x32-10 (same paper): 64-bit counter, 32-bit key.  Used for the one variate that is drawn per pixel rather than
// per particle (the pixel noise), where a four-word block would be three quarters wasted.
__device__ __forceinline__ void philox2x32_10(uint32_t &c0, uint32_t &c1, uint32_t key)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi = __umulhi(0xD256D193u, c0), lo = 0xD256D193u * c0;
        c0 = hi ^ (key + (uint32_t)r * 0x9E3779B9u) ^ c1;
        c1 = lo;
    }
}

// draw sites of one particle; plane-dependent sites are SITE_PLANE + 4*plane + {STEP,DIG,PIX}
enum : uint32_t { SITE_EVENT = 0, SITE_START = 1, SITE_PLANE = 4, SUB_STEP = 0, SUB_DIG = 1, SUB_PIX = 2 };

__device__ __forceinline__ U4 philox_at(const PhiloxKeys &seed, uint64_t particle, uint32_t site, uint32_t index)
{
    return philox4x32_10(U4{(uint32_t)particle, (uint32_t)(particle >> 32), site, index}, seed);
}

// 53-bit uniform in [0,1) the way numpy/numba build it from two words
__device__ __forceinline__ double u53_closed_open(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
// 53-bit uniform in (0,1)
__device__ __forceinline__ double u53_open(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}

// lg2 / sqrt straight from the special-function unit (no denormal fix-up, no correctly-rounded square root: their
// arguments here are uniforms in [2^-33, 1] and 0 <= -2 ln u < 46)
__device__ __forceinline__ float lg2_sfu(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_sfu(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Box-Muller pair in fp32 arithmetic, returned as doubles (sites that tolerate 24-bit normals: scattering
// kicks, cluster radii, pixel noise).  Special-function-unit approximations throughout: lg2 / sin / cos are
// good to ~2^-21 absolute on these ranges, far below what any of the consumers resolves.
__device__ __forceinline__ void normal_pair_f32(uint32_t a, uint32_t b, double &n0, double &n1)
{
    float u = fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // (0,1]
    float v = fmaf((float)b, 1.4629180792671596e-09f, 7.3145903963357980e-10f);   // (0,2pi]
    float r = sqrt_sfu(-1.3862943611198906f * lg2_sfu(u));                         // sqrt(-2 ln u)
    float s, c;
    __sincosf(v, &s, &c);
    n0 = (double)(r * c);
    n1 = (double)(r * s);
}
// ln Gamma(k+1) for the PTRS acceptance test: Stirling's series above 16 (truncation < 2e-12, against a
// continuous uniform on the other side of the inequality), the library routine below
__device__ __forceinline__ double log_gamma_kp1(double k)
{
    if (k < 16.0) return lgamma(k + 1.0);
    const double z = k + 1.0, iz = 1.0 / z, iz2 = iz * iz;
    return (z - 0.5) * log(z) - z + 0.9189385332046727 + iz * (1.0 / 12.0 - iz2 * (1.0 / 360.0 - iz2 * (1.0 / 1260.0)));
}

// ---------------------------------------------------------------------------------------------------
// variate sources.  Both answer the same questions; the kernels are templated on the source so that the
// deterministic (injected) and the production (Philox) mode run the very same physics code.
// ---------------------------------------------------------------------------------------------------
struct PhiloxSource {
    PhiloxKeys seed;

    __device__ bool accepted(const ConfigK &cfg, uint64_t k) const
    {
        U4 w = philox_at(seed, k, SITE_EVENT, 0);
        return u53_closed_open(w.x, w.y) < cfg.accept_p;
    }
    // Poisson(lam): multiplication method below 10, Hoermann's PTRS above (same laws as numpy's legacy
    // generator that the reference draws from, main/hit.py:239)
    __device__ int64_t gap(const ConfigK &cfg, uint64_t k) const
    {
        if (cfg.alias) {
            // alias method: exact for the tabulated law (the table spans the mean +- 10 sigma), one Philox block
            U4               w = philox_at(seed, k, SITE_EVENT, 1);
            const int        col = min((int)(u53_closed_open(w.x, w.y) * (double)cfg.alias_n), cfg.alias_n - 1);
            const AliasEntry e = cfg.alias[col];
            return cfg.alias_kmin + (u53_closed_open(w.z, w.w) < e.prob ? col : e.alias);
        }
        const double lam = cfg.lam;
        if (lam >= 10.0) {
            for (uint32_t it = 1;; ++it) {
                U4     w = philox_at(seed, k, SITE_EVENT, it);
                double U = u53_closed_open(w.x, w.y) - 0.5;
                double V = u53_closed_open(w.z, w.w);
                double us = 0.5 - fabs(U);
                double kk = floor((2.0 * cfg.p_a / us + cfg.p_b) * U + lam + 0.43);
                if (us >= 0.07 && V <= cfg.p_vr) return (int64_t)kk;
                if (kk < 0.0 || (us < 0.013 && V > us)) continue;
                if (log(V) + cfg.p_log_invalpha - log(cfg.p_a / (us * us) + cfg.p_b) <=
                    -lam + kk * cfg.p_loglam - log_gamma_kp1(kk))
                    return (int64_t)kk;
            }
        }
        if (lam == 0.0) return 0;
        int64_t n = 0;
        double  prod = 1.0;
        for (uint32_t it = 1;; ++it) {
            U4 w = philox_at(seed, k, SITE_EVENT, it);
            prod *= u53_closed_open(w.x, w.y);
            if (prod <= cfg.p_explam) return n;
            ++n;
            prod *= u53_closed_open(w.z, w.w);
            if (prod <= cfg.p_explam) return n;
            ++n;
        }
    }
    // one Philox block: the two angle normals from the fp32 generator (dispersion of 1e-4 rad), the two position
    // normals from a double-precision Box-Muller on 32-bit uniforms (millimetre beam against micrometre pixels)
    __device__ void start(uint64_t k, double z[4]) const
    {
        U4 w = philox_at(seed, k, SITE_START, 0);
        normal_pair_f32(w.x, w.y, z[0], z[1]);
        const double u = ((double)w.z + 0.5) * 2.3283064365386963e-10;   // (0,1)
        const double v = (double)w.w * 4.656612873077393e-10;            // [0,2)
        const double r = sqrt(-2.0 * log(u));
        double       sn, cs;
        sincospi(v, &sn, &cs);
        z[2] = r * cs;
        z[3] = r * sn;
    }
    // scattering kicks and the Landau loss of plane d come out of one Philox block
    __device__ void step(const ConfigK &cfg, uint64_t k, int d, double &zkx, double &zky, double &eloss) const
    {
        U4 w = philox_at(seed, k, SITE_PLANE + 4 * d + SUB_STEP, 0);
        normal_pair_f32(w.x, w.y, zkx, zky);
        eloss = landau(cfg.pl[d], k, d, w.z, w.w);
    }
    __device__ double eloss_only(const ConfigK &cfg, uint64_t k, int d) const
    {
        U4 w = philox_at(seed, k, SITE_PLANE + 4 * d + SUB_STEP, 0);
        return landau(cfg.pl[d], k, d, w.z, w.w);
    }
    // Energy loss: the Moyal-form law the reference's accept/reject pool follows (main/hit.py:345-359, 399-400,
    // 431), restricted to [0, 0.5] MeV.  If Z ~ N(0,1) then -ln(Z^2) is Moyal distributed, so when
    // the window holds practically all of the law two normals are tried and the (astronomically rare) leftovers
    // go to the inversion below; planes whose window is a thin slice of the law (thick absorbers) always invert.
    __device__ double landau(const PlaneK &p, uint64_t k, int d, uint32_t a, uint32_t b) const
    {
        if (p.lan_direct) {
            double z0, z1;
            normal_pair_f32(a, b, z0, z1);
            double e = p.lan_loc + p.lan_scale * (double)(-1.3862943611198906f * lg2_sfu(fabsf((float)z0)));
            if (e >= 0.0 && e <= 0.5) return e;
            e = p.lan_loc + p.lan_scale * (double)(-1.3862943611198906f * lg2_sfu(fabsf((float)z1)));
            if (e >= 0.0 && e <= 0.5) return e;
            U4 w = philox_at(seed, k, SITE_PLANE + 4 * d + SUB_STEP, 1);
            a = w.x;
            b = w.y;
        }
        return landau_inverse(p, a, b);
    }
    // inversion of F(lambda) = erfc(exp(-lambda/2)/sqrt2) on the window
    __device__ static __noinline__ double landau_inverse(const PlaneK &p, uint32_t a, uint32_t b)
    {
        double U = p.lan_ulo + p.lan_uspan * u53_open(a, b);
        double s = erfcinv(U);
        double lam = -2.0 * log(1.4142135623730951 * s);
        double e = p.lan_loc + p.lan_scale * lam;
        return fmin(fmax(e, 0.0), 0.5);
    }
    // pixkey: key of the pair's pixel-noise stream (below); a function of (seed, particle, plane) like everything else here
    __device__ void dig(uint64_t k, int d, double &zrx, double &zry, double &zseed, uint32_t &pixkey) const
    {
        U4     w = philox_at(seed, k, SITE_PLANE + 4 * d + SUB_DIG, 0);
        double spare;
        normal_pair_f32(w.x, w.y, zrx, zry);
        normal_pair_f32(w.z, w.w, zseed, spare);
        pixkey = w.z ^ __funnelshift_l(w.w, w.w, 16);
    }
    static constexpr bool kPixelKeyed = true;
#ifndef PTB_PIXEL_PHILOX4
    // Noise normal of the p-th candidate pixel of particle k on plane d: Philox2x32-10 under the pair's key, counter =
    // (particle id, plane, pixel index) -- 34 + 5 + 20 bits, so no two pixels of a run share a counter whatever their
    // keys -- and the cosine branch of a Box-Muller pair (the sine branch would belong to no pixel).
    __device__ double pixel(uint64_t k, int d, uint32_t p, uint32_t key) const
    {
        uint32_t a = (uint32_t)k, b = p | ((uint32_t)d << 20) | ((uint32_t)(k >> 32) << 25);
        philox2x32_10(a, b, key);
        const float u = fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // as normal_pair_f32
        const float v = fmaf((float)b, 1.4629180792671596e-09f, 7.3145903963357980e-10f);
        return (double)(sqrt_sfu(-1.3862943611198906f * lg2_sfu(u)) * __cosf(v));
    }
#else
    // (round-1 scheme, kept for A/B timing: four consecutive pixels share one Philox4x32 block, every lane works out the
    // one normal it needs)
    __device__ double pixel(uint64_t k, int d, uint32_t p, uint32_t /*key*/) const
    {
        const U4       w = philox_at(seed, k, SITE_PLANE + 4 * d + SUB_PIX, p >> 2);
        const uint32_t a = (p & 2u) ? w.z : w.x, b = (p & 2u) ? w.w : w.y;
        const float    u = fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // as normal_pair_f32
        const float    v = fmaf((float)b, 1.4629180792671596e-09f, 7.3145903963357980e-10f);
        const float    r = sqrt_sfu(-1.3862943611198906f * lg2_sfu(u));
        float          s, c;
        __sincosf(v, &s, &c);
        return (double)(r * ((p & 1u) ? s : c));
    }
#endif
};

struct InjectedSource {
    PtbInjected v;
    uint64_t    first, n;

    __device__ bool accepted(const ConfigK &, uint64_t k) const { return v.accepted[k - first] != 0; }
    __device__ int64_t gap(const ConfigK &, uint64_t k) const { return v.gap[k - first]; }
    __device__ void start(uint64_t k, double z[4]) const
    {
        const double *p = v.zstart + 4 * (k - first);
        z[0] = p[0]; z[1] = p[1]; z[2] = p[2]; z[3] = p[3];
    }
    __device__ void step(const ConfigK &cfg, uint64_t k, int d, double &zkx, double &zky, double &eloss) const
    {
        const double *p = v.zkick + ((k - first) * (uint64_t)(cfg.n_planes - 1) + (uint64_t)(d - 1)) * 2;
        zkx = p[0];
        zky = p[1];
        eloss = v.eloss[(uint64_t)d * n + (k - first)];
    }
    // k may be first-1 (the look-back of SURVEY.md Q4)
    __device__ double eloss_only(const ConfigK &, uint64_t k, int d) const
    {
        return (k + 1 == first) ? v.eloss0_prev : v.eloss[(uint64_t)d * n + (k - first)];
    }
    __device__ void dig(uint64_t k, int d, double &zrx, double &zry, double &zseed, uint32_t &pixkey) const
    {
        const double *p = v.dig_z + v.dig_off[(uint64_t)d * (n + 1) + (k - first)];
        zrx = p[0]; zry = p[1]; zseed = p[2];
        pixkey = 0;
    }
    // jf: index of the pixel in the reference's walk over its full bounding box = draw order
    static constexpr bool kPixelKeyed = false;
    __device__ double pixel(uint64_t k, int d, uint32_t /*p*/, uint32_t jf) const
    {
        return v.dig_z[v.dig_off[(uint64_t)d * (n + 1) + (k - first)] + 3 + jf];
    }
};

// ---------------------------------------------------------------------------------------------------
// physics
// ---------------------------------------------------------------------------------------------------

// a / b with y = RN(1/b) known in advance: q0 = RN(a*y), r = a - q0*b (exact in an FMA), q = RN(q0 + r*y).
// Markstein's correction step: with a correctly rounded reciprocal the result is the correctly rounded quotient,
// i.e. bit-identical to the IEEE division the reference executes (tests/test_division.py checks the recurrence in
// exact rational arithmetic; the GPU parity tests check the implementation).  3 instructions instead of ~28.
__host__ __device__ __forceinline__ double div_rcp(double a, double b, double y)
{
    const double q0 = a * y;
    const double r = fma(-q0, b, a);
    return fma(r, y, q0);
}

// tan for the flight step (main/hit.py:322-323).  Beam angles are of order 1e-4..1e-1 rad: below 2^-5 the odd
// Taylor polynomial through x^13 is accurate to ~0.6 ulp (truncation 1e-24 relative), cheaper than the generic
// routine's range reduction; larger angles take the library tan.
__device__ __forceinline__ double tan_flight(double x)
{
    if (fabs(x) < 0.03125) {
        const double t = x * x;
        double       p = 21844.0 / 6081075.0;
        p = fma(p, t, 1382.0 / 155925.0);
        p = fma(p, t, 62.0 / 2835.0);
        p = fma(p, t, 17.0 / 315.0);
        p = fma(p, t, 2.0 / 15.0);
        p = fma(p, t, 1.0 / 3.0);
        return fma(x, t * p, x);
    }
    return tan(x);
}

// Highland width for electrons, main/hit.py:290-297 (the charge factor z=-1 only flips a sign that is squared)
__host__ __device__ __forceinline__ double highland_theta0(double energy, double eps, double corr2)
{
    double p = sqrt(energy * energy - 0.511 * 0.511);
    double g = div_rcp(energy, 0.511, 1.0 / 0.511);
    double beta = sqrt(1.0 - 1.0 / (1.0 + g * g));
    double lead = 13.6 / (beta * p);
    return sqrt(lead * lead * eps * corr2);
}

// The same width on the device, regrouped so that it costs one reciprocal square root instead of three square
// roots, a division and a reciprocal: with g = E/0.511, beta = g/sqrt(1+g^2) and lead > 0,
//     theta0 = 13.6 sqrt(eps corr2) / (beta p) = hl_c * (1+g^2) * rsqrt((1+g^2) * g^2 * (E^2 - 0.511^2)).
// Agrees with the reference's operation order to <= 3 ulp (7e-16 relative; tests/test_highland.py), i.e. 1e-19 rad
// on a kick: far below anything the pixel indices resolve.  NaN below the electron mass, as the reference (Q17).
__device__ __forceinline__ double highland_theta0_dev(double energy, double hl_c)
{
    const double g = div_rcp(energy, 0.511, 1.0 / 0.511);
    const double g2 = g * g;
    const double num = 1.0 + g2;
    const double den = g2 * (energy * energy - 0.511 * 0.511);
    return hl_c * num * rsqrt(num * den);
}

// 1 / x for a positive x of ordinary magnitude: the special-function-unit seed and the library's own refinement
// (a cubic and a quadratic Newton step), without its detour for denormal / huge arguments.  A zero x gives NaN where
// the division gives inf; the callers only ever compare the results, and both fail every comparison.
__device__ __forceinline__ double rcp_pos(double x)
{
#ifdef PTB_LIBRARY_MATH
    return 1.0 / x;
#else
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    e = fma(e, e, e);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
#endif
}

__device__ __forceinline__ float ex2_sfu(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// width of the collected charge cloud in um, main/device.py:186-202.  np.cbrt under Numba is pow(x, 1/3).
// pow(x, 1.0/3.0) for x > 0, i.e. with the exponent rounded to double as the reference evaluates it:
// x^(1/3 + delta) = cbrt(x) * (1 + delta ln x), delta = double(1/3) - 1/3 = -1.85e-17 (the correction is a fraction
// of an ulp, so a float logarithm is ample).  ~6x fewer instructions than the generic pow and no far call.
__device__ __forceinline__ double pow_one_third(double x)
{
#ifndef PTB_LIBRARY_MATH
    // Arguments here are 1e-19 .. 1e-13 (m^3): inside the float range, so the cube root's seed is 2^(lg2(x)/3) on the
    // special-function unit (~2^-21 relative) and one Halley step in double brings it to the last bit or two -- the
    // library routine's own scheme without its exponent split, and the logarithm serves the exponent correction too.
    if (x > 1e-30 && x < 1e30) {
        const float  lg = lg2_sfu((float)x);
        const double y0 = (double)ex2_sfu(lg * 0.333333343f);
        const double y2 = y0 * y0;
        const double num = fma(-y0, y2, x);          // x - y0^3
        const double den = fma(y0, y2 + y2, x);      // x + 2 y0^3
        double       r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
        double e = fma(-den, r, 1.0);
        e = fma(e, e, e);
        r = fma(r, e, r);
        const double c = fma(y0, r * num, y0);       // y0 + y0 (x - y0^3) / (x + 2 y0^3)
        return fma(c, -1.850371707708594e-17 * (double)(0.6931471805599453f * lg), c);
    }
#endif
    const double c = cbrt(x);
    // ln x to ~1e-6 relative is ample for a correction of a few ulp: lg2 on the special-function unit
    return fma(c, -1.850371707708594e-17 * (double)(0.6931471805599453f * __log2f((float)x)), c);
}

__device__ __forceinline__ double cloud_sigma(const ConfigK &c, double eloss)
{
    if (eloss == 0.0) return c.sigma0;   // misses deposit nothing; also keeps 0/x off the division slow path
    double ev = eloss * 1e6;
    double q = div_rcp(c.qe * ev, 3.6, c.rcp_3p6);
    double carg = div_rcp(c.c3mu * q * c.t90, c.cden, c.rcp_cden);
    double coul = 0.2 * pow_one_third(carg);
    return sqrt(c.diff2 + coul * coul) * 1e6;
}

// 1-based pixel index with truncation toward zero, main/device.py:356
__device__ __forceinline__ int pixel_index(double pos, double delta, double pitch, double rcp_pitch, double half_n)
{
    const double t = div_rcp(pos + delta, pitch, rcp_pitch) + half_n;
    // the conversion saturates (NaN -> 0); a saturated index wraps to INT_MIN and fails every matrix test
    return (int)((unsigned)__double2int_rz(t) + 1u);
}

// centre of pixel idx, main/device.py:372
__device__ __forceinline__ double pixel_centre(int idx, double delta, double pitch, double half_n, double half_pitch)
{
    return pitch * ((double)idx - half_n - 1.0) - delta + half_pitch;
}

#define PTB_SQRT_2PI 2.5066282746310002 /* sqrt(2*pi) rounded as numpy does: np.sqrt(2*np.pi) */

// RN(1 / (2 rc^2)) for the clamped radius rc = 1/sqrt(2 pi), folded by the compiler in IEEE double
#define PTB_RCP_2RC2_CLAMPED (1.0 / (2.0 * ((1.0 / PTB_SQRT_2PI) * (1.0 / PTB_SQRT_2PI))))

// radius entering the Gaussian profile: clamped below 1 um, main/device.py:400-403
__device__ __forceinline__ double profile_radius(double r) { return r < 1.0 ? 1.0 / PTB_SQRT_2PI : r; }

// 1/(sigma*sqrt(2 pi)), first factor of main/device.py:378
__device__ __forceinline__ double profile_norm(double rc) { return rcp_pos(rc * PTB_SQRT_2PI); }

// exp(x) for x <= 0 (the Gaussian profile's exponent): x = (64 m + j) ln2/64 + r with |r| <= ln2/128, so that
//     exp(x) = 2^m * 2^(j/64) * (1 + r + r^2/2 + ... + r^5/120)          (the r^6 term is below 4e-17)
// with 2^(j/64) from a 64-entry table in shared memory: about half the instructions of the library routine (whose
// polynomial has eleven terms, each coefficient a pair of moves), ~1 ulp.  Arguments below -700 (profiles that vanish
// against any charge) are clamped so that the result stays a normal number for the exponent arithmetic; a NaN argument
// comes out as that tiny number too (the library gives NaN) -- no caller produces one from finite positions.
__device__ __forceinline__ double exp_nonpos(double x, const double *tab)
{
#ifdef PTB_LIBRARY_MATH
    return exp(x);
#else
    x = fmax(x, -700.0);
    const double t = fma(x, 92.332482616893657, 6755399441055744.0);     // 64/ln2; 1.5 * 2^52: round to nearest integer
    const int    k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double       r = fma(kd, -1.0830424696249145e-02, x);                 // ln2/64, head ...
    r = fma(kd, -3.623510646634843e-19, r);                               // ... and tail (ln2/64 - head)
    double p = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r * r, r);                                                 // exp(r) - 1
    const double T = tab[k & 63];
    const double y = fma(T, p, T);
    return __hiloint2double(__double2hiint(y) + ((k >> 6) << 20), __double2loint(y));
#endif
}

// exp(-(d^2)/(2 sigma^2)), second factor of main/device.py:378; y2 = RN(1/(2 sigma^2))
__device__ __forceinline__ double profile_shape(double d, double rc, double y2, const double *tab)
{
    return exp_nonpos(div_rcp(-(d * d), 2.0 * (rc * rc), y2), tab);
}

}  // namespace ptb

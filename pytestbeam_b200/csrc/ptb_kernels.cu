// libptb_b200 -- kernels and C ABI (include/ptb.h).
//
// One call of ptb_simulate runs these kernels on the caller's stream:
//
//   k_tracks    one lane per particle, one warp per group of 32 consecutive particles: beam draw, plane-by-
//               plane propagation, Highland kick, Landau loss, acceptance test.  Hands the per-plane
//               (x, y, energy loss) of the group to k_digitise through a coalesced, L2-resident park.
//   k_digitise  a persistent grid whose warps take one group after the other, warp-cooperative: the 32 particles' pixel boxes are flattened into work
//               lists that the lanes walk 32 items at a time (first the separable per-column / per-row
//               profiles, then the pixels, one per lane), and hits are compacted in reference order (particle,
//               column, row) with ballots.  Only the box pixels that can pass the ellipse test and lie on the
//               matrix are evaluated.  Before the pixels of a plane are evaluated the group reserves room for
//               all of them (an upper bound on its hits) in that plane's scratch slab with one atomic, so
//               that the kept pixels go straight to their place in the segment.  (Two kernels rather than
//               one: fused, the code no longer fits the instruction cache and ran 26 % slower -- profiles/.)
//   k_scan_sums, k_scan   exclusive prefix over the per-group counters (hits per plane, accepted events,
//               time), 16 blocks per counter column.
//   k_gather    copies each group's scratch segment to its final, particle-ordered position in the
//               structure-of-arrays hit slab and stamps the event numbers.
//
// The kernels never wait on each other inside a launch (no look-back spinning), so forward progress
// does not depend on block scheduling order.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <unistd.h>
#include <vector>

#include "ptb_device.cuh"

namespace ptb {

#ifndef PTB_DIGI_WARPS
#define PTB_DIGI_WARPS 4
#endif
constexpr int      WARPS_PER_BLOCK = PTB_DIGI_WARPS;    // k_digitise
constexpr int      BLOCK = 32 * WARPS_PER_BLOCK;
#ifndef PTB_DIGI_MIN_BLOCKS
#define PTB_DIGI_MIN_BLOCKS 8
#endif
#ifndef PTB_TRACK_MIN_BLOCKS
#define PTB_TRACK_MIN_BLOCKS 7
#endif
constexpr int      DIGI_MIN_BLOCKS = PTB_DIGI_MIN_BLOCKS;
// shared-memory sizing of k_digitise (compile-time).  The small variant exists so that the tests can drive the
// multi-round pool path with ordinary geometries.
constexpr int      POOL_ENTRIES = 128;
constexpr int      SMALL_POOL = 96;
constexpr int      TRACK_BLOCK = 128;                   // k_tracks
constexpr unsigned FULL = 0xffffffffu;
constexpr unsigned long long NO_SEGMENT = ~0ull;      // table entry of a group whose scratch reservation did not fit
constexpr int ALLOC_LINES = PTB_MAX_PLANES + 2;       // counters of Workspace::alloc, one per 128-byte line
constexpr int ALLOC_DEFERRED = PTB_MAX_PLANES, ALLOC_TICKET = PTB_MAX_PLANES + 1;
constexpr int      SCAN_SEGMENTS = 16;                  // every counter column is scanned by this many blocks

static unsigned long long g_launches = 0;

// Index checks of the warp-cooperative phases (compute-sanitizer is not available on the GPU pool): compiled in with
// -DPTB_DEBUG_CHECKS (tools/build_variants.py dbg=-DPTB_DEBUG_CHECKS), a violated check raises PTB_DEV_INTERNAL in the totals.
#define PTB_DEV_INTERNAL 0x80000000u
#ifdef PTB_DEBUG_CHECKS
#define PTB_DBG_CHECK(cond)                                           \
    do {                                                              \
        if (!(cond)) atomicOr(&P.totals->error, PTB_DEV_INTERNAL);    \
    } while (0)
#else
#define PTB_DBG_CHECK(cond) \
    do {                    \
    } while (0)
#endif

// ---------------------------------------------------------------------------------------------------
// workspace layout (device memory owned by the caller)
// ---------------------------------------------------------------------------------------------------
struct Workspace {
    unsigned long long *alloc;      // [ALLOC_LINES * 16] one counter per 128-byte line: scratch rows handed out per plane;
                                    // then the deferred (group, plane) items and the group tickets of k_digitise
    unsigned long long *deferred;   // [n_groups * D] (group << 8 | plane) of the pairs k_digitise left to k_digitise_big
    unsigned long long *table;      // [(2D+5)][n_groups]: hits[D], accepted, tsum, pairs, inside, pixels, base[D]
    unsigned long long *prefix;     // [(D+2)][n_groups]
    PtbBases           *bases_used; // copy of bases_in taken by k_scan before bases_out is written
    double             *park;       // [n_groups][D][3][32] digitisation inputs handed from k_tracks to k_digitise
    unsigned           *accmask;    // [n_groups] trigger ballot of every group
    unsigned long long *segsum;     // [(D+5) * SCAN_SEGMENTS] partial sums of the counter columns
    // scratch slab of every plane: segments in arbitrary (atomic) order, rows inside a segment in reference order
    uint2              *s_hit[PTB_MAX_PLANES];   // packed (col | row << 16), f32 charge bits
    uint8_t            *s_ev[PTB_MAX_PLANES];    // accepted particles of the group before the hit's particle
    double             *s_q64[PTB_MAX_PLANES];
    unsigned long long  cap[PTB_MAX_PLANES];     // rows
};

struct KParams {
    unsigned long long first, n, n_groups;
    uint32_t           flags, _pad;
    PtbTrackTable      trk;
    Workspace          ws;
    PtbTotals         *totals;
    PtbCompactOut      cmp;   // PTB_COMPACT
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t carve_workspace(int D, unsigned long long n, const unsigned long long *scratch_rows, uint32_t flags,
                              char *base, Workspace *ws)
{
    unsigned long long ng = (n + 31) / 32;
    size_t             off = 0;
    auto take = [&](size_t bytes) {
        size_t at = off;
        off = align_up(off + bytes, 256);
        return base ? base + at : (char *)nullptr;
    };
    char *p;
    p = take(sizeof(unsigned long long) * ALLOC_LINES * 16);
    if (ws) ws->alloc = (unsigned long long *)p;
    p = take((flags & PTB_DIGITISE) ? sizeof(unsigned long long) * (size_t)D * ng : 0);
    if (ws) ws->deferred = (unsigned long long *)p;
    p = take(sizeof(unsigned long long) * (size_t)(2 * D + 5) * ng);
    if (ws) ws->table = (unsigned long long *)p;
    p = take(sizeof(unsigned long long) * (size_t)(D + 2) * ng);
    if (ws) ws->prefix = (unsigned long long *)p;
    p = take(sizeof(PtbBases));
    if (ws) ws->bases_used = (PtbBases *)p;
    const bool parked = (flags & PTB_DIGITISE) && !(flags & PTB_TRACKS_FROM_TABLE);
    p = take(parked ? sizeof(double) * 96 * (size_t)D * ng : 0);
    if (ws) ws->park = (double *)p;
    p = take(sizeof(unsigned) * ng);
    if (ws) ws->accmask = (unsigned *)p;
    p = take(sizeof(unsigned long long) * (size_t)(D + 5) * SCAN_SEGMENTS);
    if (ws) ws->segsum = (unsigned long long *)p;
    for (int d = 0; d < D; ++d) {
        unsigned long long c = (flags & PTB_DIGITISE) ? scratch_rows[d] : 0;
        p = take(8 * (size_t)c);
        if (ws) ws->s_hit[d] = (uint2 *)p;
        p = take((size_t)c);
        if (ws) ws->s_ev[d] = (uint8_t *)p;
        p = take((flags & PTB_KEEP_CHARGE64) ? 8 * (size_t)c : 0);
        if (ws) ws->s_q64[d] = (double *)p;
        if (ws) ws->cap[d] = c;
    }
    return off;
}


// shared memory of one warp: the per-particle cluster descriptors
struct WarpShared {
    double   in[3][32];      // (x, y, energy loss) of the next plane, filled by cp.async while the current one is worked on
    // per particle and axis ([0] = x / columns, [1] = y / rows, so that phase A picks its operands by index):
    double   pos[2][32], rad[2][32], nrm[2][32];       // hit position, cluster radius, 1/(sigma sqrt(2 pi))
    double   ye[2][32];                                // RN(1/r^2) for div_rcp
    double   q[32], seedq[32];
    int      lo[2][32];                                // trimmed box: first column / row ...
    uint16_t ncols[32], nrows[32];                     // ... and extent (ncols + nrows <= pool entries, or the pair is
                                                       //     left to k_digitise_big)
    int      jbase[32], nrfull[32];                    // position of the trimmed box inside the reference's box
    int      qexcl[32], qprefix[32], poff[32];     // work lists: first / one-past-last pixel slot, pool offset
    uint8_t  ownA[32], ownB[32];                   // the particles that own profile entries / pixel slots, in order
    float    rnr[32];                              // 1 / rows of the trimmed box
    uint32_t seedcr[32];
    uint8_t  evoff[32], accb[32];
    uint8_t  cnt[32];                              // PTB_COMPACT: rows of every particle on the current plane (saturating)
    unsigned long long census;                     // census of the group (packed), kept by lane 0
};

// bytes of one warp's region: descriptors, profile pool (2 doubles per column / row)
__host__ __device__ constexpr size_t warp_shared_bytes(int pool)
{
    size_t b = sizeof(WarpShared);
    b += (size_t)pool * 16;               // profile value and ellipse term per box column / row
    return (b + 15) / 16 * 16;
}

// 8-byte asynchronous global -> shared copy (LDGSTS) and its completion
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---------------------------------------------------------------------------------------------------
// k_tracks: one lane per particle, one warp per group of 32 consecutive particles.  Event draws (trigger
// mask, Poisson gap), the source state and the plane-by-plane propagation (main/hit.py:70-80, 104, 209-239).
// Leaves behind, per group: the digitisation inputs (x, y, energy loss) of every plane in the workspace
// ("park", [D][3][32] doubles, coalesced), the trigger ballot, and the group's accepted / time counters.
// ---------------------------------------------------------------------------------------------------
template <class Src>
__global__ void __launch_bounds__(TRACK_BLOCK, PTB_TRACK_MIN_BLOCKS) k_tracks(const __grid_constant__ ConfigK cfg,
                                                       const __grid_constant__ KParams P, const Src src)
{
    const int                lane = threadIdx.x & 31;
    const unsigned long long group = (unsigned long long)blockIdx.x * (TRACK_BLOCK / 32) + (threadIdx.x >> 5);
    if (group >= P.n_groups) return;

    const int  D = cfg.n_planes;
    const bool digitise = (P.flags & PTB_DIGITISE) != 0;
    const bool write_tracks = (P.flags & PTB_WRITE_TRACKS) != 0;
    double    *park = P.ws.park + group * (unsigned long long)(D * 96) + lane;

    const unsigned long long kl = group * 32 + lane;  // index inside this call's range
    const bool               valid = kl < P.n;
    const unsigned long long k = P.first + kl;        // global particle id

    // ---- event-level draws: trigger mask and time gap ------------------------------------------------
    bool    acc = false;
    int64_t gap = 0;
    if (valid) {
        acc = src.accepted(cfg, k);
        gap = src.gap(cfg, k);
    }
    const unsigned accmask = __ballot_sync(FULL, acc);
    int64_t        t_rel = gap;                              // inclusive scan: time relative to the group start
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        int64_t o = __shfl_up_sync(FULL, t_rel, s);
        if (lane >= s) t_rel += o;
    }

    // ---- source state (main/hit.py:70-80, 104, 234-238) ----------------------------------------------
    double E = 0, ax = 0, ay = 0, x = 0, y = 0, e0_self = 0, e0_prev = 0;
    if (valid) e0_self = src.eloss_only(cfg, k, 0);
    e0_prev = __shfl_up_sync(FULL, e0_self, 1);
    if (valid) {
        if (k == 0)
            e0_prev = e0_self;                           // event 0 starts from its own plane-0 loss
        else if (lane == 0)
            e0_prev = src.eloss_only(cfg, k - 1, 0);     // neighbour group: recomputed / injected look-back
        double z[4];
        src.start(k, z);
        const bool first_event = (k == 0);
        ax = cfg.angle_x + (first_event ? cfg.theta0_first : cfg.disp_x) * z[0];
        ay = cfg.angle_y + (first_event ? cfg.theta0_first : cfg.disp_y) * z[1];
        x = cfg.loc_x + cfg.sigma_x * z[2];
        y = cfg.loc_y + cfg.sigma_y * z[3];
        E = cfg.energy - e0_prev;
    }

    // ---- the track through all planes (main/hit.py:211-231)
#pragma unroll 1
    for (int d = 0; d < D; ++d) {
        const PlaneK &pl = cfg.pl[d];
        double        eloss = 0;
        if (valid) {
            if (d == 0) {
                eloss = e0_self;   // plane 0 is the start point: never propagated to, never bounds-checked (Q2)
            } else {
                x = x + tan_flight(ax) * pl.z;   // absolute z of the target plane (Q1)
                y = y + tan_flight(ay) * pl.z;
                const double th = highland_theta0_dev(E, pl.hl_c);
                double       zkx, zky;
                src.step(cfg, k, d, zkx, zky, eloss);
                double kx = 0.0 + th * zkx;
                double ky = 0.0 + th * zky;
                if (x > pl.x_pos || x < pl.x_neg || y > pl.y_pos || y < pl.y_neg) {
                    kx = 0.0;
                    ky = 0.0;
                    eloss = 0.0;
                }
                ax = ax + kx;
                ay = ay + ky;
                E = E - eloss;
            }
            if (write_tracks) {
                const unsigned long long o = kl * D + d;
                P.trk.energy[o] = E;
                P.trk.angle_x[o] = ax;
                P.trk.angle_y[o] = ay;
                P.trk.x[o] = x;
                P.trk.y[o] = y;
                P.trk.z[o] = pl.z;
                P.trk.time_stamp[o] = t_rel;   // k_gather adds the group's base
                P.trk.eloss[(unsigned long long)d * P.n + kl] = eloss;
                if (d == 0) P.trk.accepted[kl] = acc ? 1 : 0;
            }
        }
        if (digitise) {
            park[d * 96] = x;
            park[d * 96 + 32] = y;
            park[d * 96 + 64] = eloss;
        }
    }


    const int64_t tsum = __shfl_sync(FULL, t_rel, 31);
    if (lane == 0) {
        unsigned long long *t = P.ws.table + group;
        t[(unsigned long long)(D + 0) * P.n_groups] = (unsigned long long)__popc(accmask);
        t[(unsigned long long)(D + 1) * P.n_groups] = (unsigned long long)tsum;
        P.ws.accmask[group] = accmask;
        if (!digitise) {
            for (int d = 0; d < D; ++d) t[(unsigned long long)d * P.n_groups] = 0;
            for (int c = D + 2; c < D + 5; ++c) t[(unsigned long long)c * P.n_groups] = 0;
        }
    }
}

// Cluster descriptor of one (particle, plane) pair, main/device.py:137-140, 241-266: radii, seed pixel and charge, the
// reference's bounding box and its trimmed version.  The per-particle values the later phases read go to the warp's
// shared memory (as soon as they exist: the fewer values live at once, the better); the counts come back.
struct Cluster {
    int  cnt;        // pixels of the reference's bounding box (census)
    int  ncT, nrT;   // columns / rows of the trimmed box
    bool inside;     // seed pixel on the matrix
    bool too_large;  // bounding box beyond 2^20 pixels: refused
};

template <class Src>
__device__ __forceinline__ Cluster describe(const ConfigK &cfg, const PlaneK &pl, uint32_t flags, const Src &src,
                                            unsigned long long k, int d, double x, double y, double eloss, WarpShared &W,
                                            int lane)
{
    Cluster      out{0, 0, 0, false, false};
    const double sig = cloud_sigma(cfg, eloss);
    double       zrx, zry, zseed;
    uint32_t     pixkey;
    src.dig(k, d, zrx, zry, zseed, pixkey);
    // What selects a pixel's noise variate: the key of the pair's stream (Philox; stored at once, it need not live on),
    // or the position of the trimmed box inside the reference's box (injected variates come in the reference's draw order)
    if (Src::kPixelKeyed) W.jbase[lane] = (int)pixkey;
    const bool   injected_charge = (flags & PTB_ZERO_RADIUS) != 0;   // threshold scan: radius 0
    const double rx = injected_charge ? 0.0 : fabs(0.0 + sig * zrx);
    const double ry = injected_charge ? 0.0 : fabs(0.0 + sig * zry);
    const int    sc = pixel_index(x, pl.delta_x, pl.pitch_c, pl.rcp_pc, pl.half_nc);
    const int    sr = pixel_index(y, pl.delta_y, pl.pitch_r, pl.rcp_pr, pl.half_nr);
    // a NaN position (energy below the electron mass upstream, Q17) converts to 0 here but to INT64_MIN under Numba,
    // which fails the matrix test: such a track leaves no hit
    if (sc >= 1 && sc <= pl.ncol && sr >= 1 && sr <= pl.nrow && x == x && y == y) {
        out.inside = true;
        const double inx = profile_norm(profile_radius(rx));
        const double iny = profile_norm(profile_radius(ry));
        const double q = eloss == 0.0 ? 0.0 : div_rcp(eloss * 1e6, 3.6, cfg.rcp_3p6);
        const double noise = 0.0 + pl.noise_sigma * zseed;
        // seed pixel: both profile factors are exp(-0) = 1; no pixel-area factor (Q9)
        const double seedq = (q * (inx * iny) + noise) * 1e-3;
        W.q[lane] = q; W.nrm[0][lane] = inx; W.nrm[1][lane] = iny; W.seedq[lane] = seedq;
        W.seedcr[lane] = (uint32_t)sc | ((uint32_t)sr << 16);
        const int    c_hi = pixel_index(x + rx, pl.delta_x, pl.pitch_c, pl.rcp_pc, pl.half_nc);
        const int    c_lo = pixel_index(x - rx, pl.delta_x, pl.pitch_c, pl.rcp_pc, pl.half_nc);
        const int    r_hi = pixel_index(y + ry, pl.delta_y, pl.pitch_r, pl.rcp_pr, pl.half_nr);
        const int    r_lo = pixel_index(y - ry, pl.delta_y, pl.pitch_r, pl.rcp_pr, pl.half_nr);
        const long long nc = (long long)c_hi - c_lo + 1, nr = (long long)r_hi - r_lo + 1;
        long long       box = (nc > 0 && nr > 0) ? nc * nr : 0;
        if (box > (1ll << 20)) {
            out.too_large = true;
            box = 0;
        }
        out.cnt = (int)box;
        // A box pixel can only be kept if its column passes dx^2/rx^2 <= 1, its row dy^2/ry^2 <= 1 (both terms
        // of the ellipse test are non-negative) and it lies on the matrix (main/device.py:290-295).  The box
        // is clipped to the matrix, and its edge columns / rows -- the usual casualties, their centres lie up
        // to half a pitch outside x +- rx -- are dropped when they fail the very expression the pixel test
        // uses.  Whatever else fails is still rejected by the full test of every pixel.
        const double yx = rcp_pos(rx * rx), yy = rcp_pos(ry * ry);
        int ca = 0, cb = -1, ra = 0, rb = -1;      // box-relative, inclusive
        if (box > 0) {
            ca = max(0, 1 - c_lo); cb = min((int)nc - 1, pl.ncol - c_lo);
            ra = max(0, 1 - r_lo); rb = min((int)nr - 1, pl.nrow - r_lo);
            auto term = [&](int idx, double delta, double pitch, double hn, double hp, double pos, double r,
                            double y_) {
                const double dd = pixel_centre(idx, delta, pitch, hn, hp) - pos;
                return div_rcp(dd * dd, r * r, y_);
            };
            const bool f0 = !(term(c_lo + ca, pl.delta_x, pl.pitch_c, pl.half_nc, pl.half_pc, x, rx, yx) <= 1.0);
            const bool f1 = !(term(c_lo + cb, pl.delta_x, pl.pitch_c, pl.half_nc, pl.half_pc, x, rx, yx) <= 1.0);
            const bool f2 = !(term(r_lo + ra, pl.delta_y, pl.pitch_r, pl.half_nr, pl.half_pr, y, ry, yy) <= 1.0);
            const bool f3 = !(term(r_lo + rb, pl.delta_y, pl.pitch_r, pl.half_nr, pl.half_pr, y, ry, yy) <= 1.0);
            ca += f0; cb -= f1; ra += f2; rb -= f3;
        }
        int ncT = max(cb - ca + 1, 0), nrT = max(rb - ra + 1, 0);
        if (ncT == 0 || nrT == 0) ncT = nrT = 0;
        out.ncT = ncT;
        out.nrT = nrT;
        W.pos[0][lane] = x; W.pos[1][lane] = y; W.rad[0][lane] = rx; W.rad[1][lane] = ry;
        W.ye[0][lane] = yx; W.ye[1][lane] = yy;
        W.lo[0][lane] = c_lo + ca; W.lo[1][lane] = r_lo + ra; W.ncols[lane] = (uint16_t)ncT; W.nrows[lane] = (uint16_t)nrT;
        if (!Src::kPixelKeyed) W.jbase[lane] = ca * (int)nr + ra;
        W.nrfull[lane] = (int)nr;
        W.rnr[lane] = __fdividef(1.0f, (float)max(nrT, 1));
    }
    return out;
}

// PTB_COMPACT_NARROW: a count that does not fit its nibble goes to the escape list (include/ptb.h)
__device__ __forceinline__ void escape_count(const KParams &P, int d, unsigned long long kl, unsigned count)
{
    const unsigned slot = atomicAdd(&P.totals->escapes, 1u);
    if (slot < P.cmp.escape_capacity)
        P.cmp.escapes[slot] = ((unsigned long long)count << 48) | ((unsigned long long)d << 40) | kl;
    else
        atomicOr(&P.totals->error, PTB_DEV_COUNT_OVERFLOW);
}

// The separable parts of a pixel's charge and ellipse test for one column (axis 0) or row (axis 1):
//      g = 1/(sigma sqrt(2pi)) exp(-(d^2)/(2 sigma^2))  (main/device.py:378),   e = d^2 / r^2  (main/device.py:291-293)
__device__ __forceinline__ void profile_entry(const AxisK &ax, int idx, double pos, double r, double nrm, double ye,
                                              const double *exp2_tab, double &g, double &e)
{
    const double dd = pixel_centre(idx, ax.delta, ax.pitch, ax.half_n, ax.half_p) - pos;
    // 1/(2 rc^2): half of 1/r^2 (exact scaling) unless the radius is clamped (main/device.py:400-403)
    const double y2 = r < 1.0 ? PTB_RCP_2RC2_CLAMPED : 0.5 * ye;
    g = nrm * profile_shape(dd, profile_radius(r), y2, exp2_tab);
    e = div_rcp(dd * dd, r * r, ye);
}

// ---------------------------------------------------------------------------------------------------
// k_digitise: one warp per group of 32 particles, plane by plane (main/device.py:132-168).  Inputs come from
// k_tracks' park or, with PTB_TRACKS_FROM_TABLE, from the caller's hit_tables arrays exactly as
// main/device.py:132-138 reads them.
// ---------------------------------------------------------------------------------------------------
// PC = profile-pool entries, MODE = output form (0: slabs, 1: slabs + f8 charge, 2: compressed rows): compile-time, so that every shared-memory address
// is base + immediate (as run-time values their re-derivation was 7 % of the executed instructions).
// One group of 32 particles through all planes: the body of k_digitise.  Leaves the group's rows in its scratch segments
// and its counters in the workspace table.
template <class Src, int PC, int MODE>
__device__ __forceinline__ void digitise_group(const ConfigK &cfg, const KParams &P, const Src &src,
                                               const unsigned long long group, WarpShared &W, double *pool_g,
                                               double *pool_e, const double *exp2_tab, const int lane)
{
    const int    D = cfg.n_planes;
    constexpr bool   keep64 = MODE == 1, compact = MODE == 2;
    const bool       from_table = (P.flags & PTB_TRACKS_FROM_TABLE) != 0;
    const double  *park = P.ws.park + group * (unsigned long long)(D * 96) + lane;

    const unsigned long long kl = group * 32 + lane;  // index inside this call's range
    const bool               valid = kl < P.n;
    const unsigned long long k = P.first + kl;        // global particle id

    // trigger mask of the group: from k_tracks, or from the caller's table
    unsigned accmask;
    if (from_table)   // the caller's mask, or (when it passes none) the source's own trigger draw
        accmask = __ballot_sync(FULL, valid && (P.trk.accepted ? P.trk.accepted[kl] != 0 : src.accepted(cfg, k)));
    else
        accmask = P.ws.accmask[group];
    // the trigger bit of the lane lives in shared memory (read back only on triggered planes) rather than in a
    // register that would have to survive the whole plane loop
    W.accb[lane] = (uint8_t)((accmask >> lane) & 1u);
    W.evoff[lane] = (uint8_t)__popc(accmask & lanemask_lt());   // accepted particles of this group before me
    if (from_table && lane == 0) {
        P.ws.table[(unsigned long long)(D + 0) * P.n_groups + group] = (unsigned long long)__popc(accmask);
        P.ws.table[(unsigned long long)(D + 1) * P.n_groups + group] = 0;
        P.ws.accmask[group] = accmask;
    }

    // census of the group (digitised pairs, in-acceptance pairs, box pixels): summed over the warp plane by plane and
    // kept in shared memory by lane 0, so that it costs no registers across the plane loop
    if (lane == 0) W.census = 0ull;

    // Digitisation inputs (x, y, energy loss) of plane d: fetched one plane ahead with asynchronous copies into the
    // lane's shared-memory slot (free again as soon as the current plane's values sit in registers), so that the
    // loads' latency hides behind the previous plane's work without holding registers.
    auto fetch = [&](int d) {
        double *slot = &W.in[0][lane];
        if (from_table) {
            if (valid) {
                cp_async8(slot, &P.trk.x[kl * D + d]);
                cp_async8(slot + 32, &P.trk.y[kl * D + d]);
                cp_async8(slot + 64, &P.trk.eloss[(unsigned long long)d * P.n + kl]);
            } else {
                slot[0] = slot[32] = slot[64] = 0.0;
            }
        } else {
            cp_async8(slot, &park[d * 96]);
            cp_async8(slot + 32, &park[d * 96 + 32]);
            cp_async8(slot + 64, &park[d * 96 + 64]);
        }
        cp_async_commit();
    };
    fetch(0);
    int                my_rows = 0;    // lane d: hits of this group on plane d ...
    unsigned long long my_base = 0;    // ... and the start of its scratch segment

#pragma unroll 1
    for (int d = 0; d < D; ++d) {
        const PlaneK &pl = cfg.pl[d];
        cp_async_wait_all();   // every lane reads only what it copied itself
        const double x = W.in[0][lane], y = W.in[1][lane], eloss = W.in[2][lane];
        if (d + 1 < D) fetch(d + 1);

        // ---- cluster descriptors, one lane per particle (main/device.py:137-140, 241-266) -----------
        const bool digitised = valid && (!pl.triggered || W.accb[lane]);
        Cluster    cl{0, 0, 0, false, false};
        if (compact) W.cnt[lane] = 0;
        if (digitised) cl = describe(cfg, pl, P.flags, src, k, d, x, y, eloss, W, lane);
        const int  cnt = cl.cnt;                  // pixels of the reference's bounding box (census)
        bool       inside = cl.inside;            // seed pixel on the matrix: the particle owns a work item even with an empty box
        int        nprof = cl.ncT + cl.nrT;       // columns + rows that can pass the ellipse test and lie on the matrix
        int        cntT = cl.ncT * cl.nrT;        // pixels of that trimmed box: the only ones that can become hits
        const bool too_large = cl.too_large;
        const bool big = nprof > PC;              // more columns + rows than the profile pool holds
        {   // census, one packed word: box pixels (<= 2^30 per group), in-acceptance pairs << 32, digitised pairs << 44
            const unsigned long long pairs_d = __popc(__ballot_sync(FULL, digitised)), inside_d = __popc(__ballot_sync(FULL, inside));
            const unsigned           pixels_d = __reduce_add_sync(FULL, (unsigned)cnt);   // <= 32 boxes of <= 2^20 pixels
            if (lane == 0) W.census += (unsigned long long)pixels_d | (inside_d << 32) | (pairs_d << 44);
        }

        // A pair whose trimmed box has more columns + rows than the pool holds (pitch far below the cloud width) takes the
        // whole (group, plane) out of this kernel: it is listed for k_digitise_big and counts as empty here.
        if (__any_sync(FULL, big)) {
            if (lane == 0)
                P.ws.deferred[atomicAdd(&P.ws.alloc[ALLOC_DEFERRED * 16], 1ull)] = (group << 8) | (unsigned long long)d;
            inside = false;
            nprof = cntT = 0;
        }

        // ---- room for every pixel that can become a hit (all trimmed-box pixels, or the seed fallback), reserved
        //      in the plane's scratch slab before the pixels are evaluated: the kept ones go straight to their place
        //      in the segment, and the atomic's latency hides behind phase A.  The segment's unused tail stays a hole
        //      (k_gather copies the first n_plane rows).
        const int          ub = __reduce_add_sync(FULL, inside ? max(cntT, 1) : 0);
        unsigned long long base = 0;         // start of the group's segment in the plane's scratch slab
        if (lane == 0 && ub > 0) base = atomicAdd(&P.ws.alloc[d * 16], (unsigned long long)ub);
        uint2   *seg_hit = P.ws.s_hit[d];    // become the group's segment once the reservation has come back
        uint8_t *seg_ev = P.ws.s_ev[d];
        double  *seg_q64 = P.ws.s_q64[d];
        bool     room = true;

        // The particles of the group are served in rounds so that the 1-D profiles of one round fit the
        // shared-memory pool (one round in all but pathological geometries).
        int n_plane = 0;      // hits of this group on this plane (warp-uniform)
        int o_begin = 0;
        bool first_round = true;
        while (true) {
            // inclusive prefixes of the profile entries (<= 128 per particle: 13 bits hold the warp's sum) and of the
            // pixel slots (<= 4096 per particle) of lanes >= o_begin, scanned as one packed word
            unsigned both = (lane >= o_begin) ? (unsigned)nprof | ((inside ? (unsigned)max(cntT, 1) : 0u) << 13) : 0u;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                unsigned o = __shfl_up_sync(FULL, both, s);
                if (lane >= s) both += o;
            }
            const int      pin = (int)(both & 0x1fffu);
            const unsigned fits = __ballot_sync(FULL, pin <= PC);           // a prefix: monotone in the lane
            const int      o_end = (fits == FULL) ? 32 : __ffs(~fits) - 1;  // first lane that does not fit
            const bool     mine = lane >= o_begin && lane < o_end;
            const int      my_prof = mine ? nprof : 0;
            const int      my_pix = (mine && inside) ? max(cntT, 1) : 0;   // an empty box keeps a slot for the fallback
            int            qin = (int)(both >> 13);
            if (o_end < 32) {   // the round ends early: the pixel slots of the lanes left out do not count
                qin = my_pix;
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) {
                    int o = __shfl_up_sync(FULL, qin, s);
                    if (lane >= s) qin += o;
                }
            }
            const int prof_total = __shfl_sync(FULL, pin, o_end - 1);   // o_end > o_begin: one box always fits
            const int pix_total = __shfl_sync(FULL, qin, 31);
            // Work-list look-up: the owner of item t is the first particle (with items) whose inclusive prefix exceeds
            // t.  The particles with items are listed once per round; inside a step every such particle whose items end
            // there raises the bit of its last item, so that a lane finds its owner's rank by counting the bits below
            // its own position -- one warp-wide OR and one shared-memory read instead of a five-level binary search.
            const unsigned actA = __ballot_sync(FULL, my_prof > 0), actB = __ballot_sync(FULL, my_pix > 0);
            __syncwarp();
            if (my_prof > 0) W.ownA[__popc(actA & lanemask_lt())] = (uint8_t)lane;
            if (my_pix > 0) W.ownB[__popc(actB & lanemask_lt())] = (uint8_t)lane;
            W.qexcl[lane] = qin - my_pix;
            W.qprefix[lane] = qin;
            W.poff[lane] = pin - my_prof;
            const int endA = my_prof > 0 ? pin : 0;   // one past my last profile entry; 0: none
            __syncwarp();

            // ---- phase A: the separable parts of every pixel of the round, once per column and per row:
            //      profile = 1/(sigma sqrt(2pi)) exp(-(d^2)/(2 sigma^2))  (main/device.py:378)
            //      ellipse = d^2 / r^2                                    (main/device.py:291-293)
            int doneA = 0;   // particles whose entries end before this step
            for (int t0 = 0; t0 < prof_total; t0 += 32) {
                const int      t = t0 + lane;
                const unsigned relA = (unsigned)(endA - t0 - 1);   // my last entry's position in this step, if below 32
                const unsigned endsA = __reduce_or_sync(FULL, relA < 32u ? 1u << relA : 0u);
                const int      rankA = doneA + __popc(endsA & lanemask_lt());
                doneA += __popc(endsA);
                if (t < prof_total) {
                    PTB_DBG_CHECK(rankA >= 0 && rankA < 32 && t < PC);
                    const int    own = W.ownA[rankA];
                    const int    e = t - W.poff[own];
                    PTB_DBG_CHECK(own < 32 && e >= 0 && e < (int)W.ncols[own] + (int)W.nrows[own]);
                    const int    ncl = W.ncols[own];
                    const int    a = e >= ncl ? 1 : 0;            // axis of the entry: a column (x) or a row (y)
                    const int    idx = W.lo[a][own] + (e - (a ? ncl : 0));
                    profile_entry(pl.ax[a], idx, W.pos[a][own], W.rad[a][own], W.nrm[a][own], W.ye[a][own], exp2_tab,
                                  pool_g[t], pool_e[t]);
                }
            }
            __syncwarp();
            if (first_round) {
                base = __shfl_sync(FULL, base, 0);
                room = base + (unsigned long long)ub <= P.ws.cap[d];   // else: count only, the host retries
                seg_hit += base;
                seg_ev += base;
                if (keep64) seg_q64 += base;
                first_round = false;
            }

            // ---- phase B: the pixels of the round in (particle, column, row) order, one per lane and step.  The
            //      order of the work list is the order of the output, so a kept pixel's place in the group's segment is
            //      a ballot and a population count away.
            int carry_owner = -1;   // particle whose pixels straddle the previous step, and its kept count
            int carry_kept = 0;
            int doneB = 0;          // particles whose pixels end before this step
            // one past my last pixel slot, 0: none (re-read rather than kept in a register through phase A)
            const int endB = W.qprefix[lane] != W.qexcl[lane] ? W.qprefix[lane] : 0;
            for (int w0 = 0; w0 < pix_total; w0 += 32) {
                const int      w = w0 + lane;
                const bool     has = w < pix_total;
                const unsigned relB = (unsigned)(endB - w0 - 1);
                const unsigned endsB = __reduce_or_sync(FULL, relB < 32u ? 1u << relB : 0u);
                const int      own = has ? W.ownB[doneB + __popc(endsB & lanemask_lt())] : 31;
                doneB += __popc(endsB);
                const int excl = W.qexcl[own];
                const int end = W.qprefix[own];
                const int p = w - excl;             // index of the pixel in its particle's trimmed box, column-major
                bool      keep = false;
                double    hq = 0.0;
                uint32_t  hcr = 0;
                if (has) {
                    const int ncl = W.ncols[own], nrw = W.nrows[own];
                    if (p < ncl * nrw) {            // the trimmed box may be empty (seed fallback only)
                        const int po = W.poff[own];
                        // p / nrw through the reciprocal: (p + 0.5) / nrw stays >= 1/256 away from an integer, far
                        // more than the rounding of the product for p < 4096
                        const int ci = __float2int_rz(((float)p + 0.5f) * W.rnr[own]);
                        const int ri = p - ci * nrw;
                        PTB_DBG_CHECK(own < 32 && p >= 0 && ci >= 0 && ci < ncl && ri >= 0 && ri < nrw && po >= 0 &&
                                      po + ncl + nrw <= PC);
                        // index of the pixel in the reference's (column-major) walk over its full box: the injected
                        // noise variates are keyed by it
                        const int    jf = Src::kPixelKeyed ? W.jbase[own] : W.jbase[own] + ci * W.nrfull[own] + ri;
                        const double zn = src.pixel(P.first + group * 32 + own, d, (uint32_t)p, (uint32_t)jf);
                        const double noise = 0.0 + pl.noise_sigma * zn;
                        const double signal = W.q[own] * (pool_g[po + ci] * pool_g[po + ncl + ri]);
                        hq = (signal + noise) * 1e-3 * pl.pitch_c * pl.pitch_r;   // device.py:275-288
                        // the matrix bounds are implied by the clipped box; a zero radius makes the ellipse terms
                        // NaN or inf, which fail the comparison as rx > 0 and ry > 0 would (device.py:289-295)
                        keep = hq >= pl.thr_ke && pool_e[po + ci] + pool_e[po + ncl + ri] <= 1.0;
                        hcr = (uint32_t)(W.lo[0][own] + ci) | ((uint32_t)(W.lo[1][own] + ri) << 16);
                    }
                }
                // kept pixels of my particle in this step: the lanes of one particle are contiguous in the work list
                const unsigned kballot = __ballot_sync(FULL, keep);
                const int      lo_lane = has ? max(excl - w0, 0) : 0;
                const int      hi_lane = has ? min(end - 1 - w0, 31) : 0;
                const unsigned span = (2u << hi_lane) - (1u << lo_lane);   // lanes lo..hi (2u << 31 wraps to 0)
                const int      kept_here = __popc(kballot & span);
                const int      kept_before = (own == carry_owner) ? carry_kept : 0;
                // seed fallback (main/device.py:299-302): the lane holding the particle's last pixel emits the seed
                // pixel when nothing in the box survived (its own pixel was not kept either, then)
                const bool fb = has && w == end - 1 && kept_before + kept_here == 0 && W.seedq[own] >= pl.thr_ke;
                if (fb) {
                    hq = W.seedq[own];
                    hcr = W.seedcr[own];
                }
                const unsigned emask = kballot | __ballot_sync(FULL, fb);
                if ((keep || fb) && room) {
                    const unsigned at = (unsigned)n_plane + __popc(emask & lanemask_lt());
                    PTB_DBG_CHECK((int)at < ub && base + at < P.ws.cap[d]);
                    seg_hit[at] = make_uint2(hcr, __float_as_uint((float)hq));
                    if (!compact) seg_ev[at] = W.evoff[own];
                    if (keep64) seg_q64[at] = hq;
                }
                if (compact && has && w == end - 1) {   // the lane holding a particle's last slot knows its row count
                    const int rows_k = kept_before + kept_here + (fb ? 1 : 0);
                    W.cnt[own] = (uint8_t)min(rows_k, 255);
                    if (rows_k > 255) atomicOr(&P.totals->error, PTB_DEV_COUNT_OVERFLOW);
                }
                n_plane += __popc(emask);
                // carry for a particle that continues into the next step
                const int last_lane = min(pix_total - w0, 32) - 1;
                carry_owner = __shfl_sync(FULL, own, last_lane);
                carry_kept = __shfl_sync(FULL, kept_before + kept_here, last_lane);
            }
            if (o_end >= 32) break;
            o_begin = o_end;
        }
        __syncwarp();
        if (compact) {
            if (P.flags & (PTB_COMPACT_NARROW | PTB_COMPACT_PACKED)) {   // 4 bits per particle: the even lane stores the pair's byte
                const unsigned full = W.cnt[lane], c = min(full, 15u);
                if (full >= 15u) escape_count(P, d, kl, full);   // (invalid lanes count 0)
                const unsigned pair = c | (__shfl_down_sync(FULL, c, 1) << 4);
                if (valid && !(lane & 1)) P.cmp.counts[(unsigned long long)d * ((P.n + 1) / 2) + kl / 2] = (uint8_t)pair;
            } else if (valid) {
                P.cmp.counts[(unsigned long long)d * P.n + kl] = W.cnt[lane];
            }
        }
        // lane d keeps the plane's row count and segment start; the table is written once, after the last plane.  A group
        // whose reservation did not fit wrote nothing: its count stays exact (the host retries with it), its segment
        // start tells k_gather that there is nothing to copy
        if (lane == d) {
            my_rows = n_plane;
            my_base = room ? base : NO_SEGMENT;
        }
        // device-side conditions go straight to the totals (rare), so that no error word has to live across the planes
        {
            const unsigned e = (__any_sync(FULL, too_large) ? PTB_DEV_BOX_TOO_LARGE : 0u) |
                               (room ? 0u : PTB_DEV_SCRATCH_OVERFLOW);   // totals stay exact; the host retries
            if (e && lane == 0) atomicOr(&P.totals->error, e);
        }
    }


    // ---- per-group counters for k_scan ---------------------------------------------------------------
    if (lane < D) {
        P.ws.table[(unsigned long long)lane * P.n_groups + group] = (unsigned long long)my_rows;
        P.ws.table[(unsigned long long)(D + 5 + lane) * P.n_groups + group] = my_base;
    }
    if (lane == 0) {
        unsigned long long *t = P.ws.table + group;
        const unsigned long long census = W.census;
        t[(unsigned long long)(D + 2) * P.n_groups] = census >> 44;
        t[(unsigned long long)(D + 3) * P.n_groups] = (census >> 32) & 0xfffull;
        t[(unsigned long long)(D + 4) * P.n_groups] = census & 0xffffffffull;
    }
}

// ---------------------------------------------------------------------------------------------------
// k_digitise: a persistent grid (one wave of blocks); every warp takes group after group by ticket, so that the warps
// of an SM drift apart instead of marching through the phases in step and the last wave has no ragged tail (1.3 %
// against one block per four groups).  The rows stay in the scratch segments for k_gather.
// ---------------------------------------------------------------------------------------------------
template <class Src, int PC, int MODE>
__global__ void __launch_bounds__(BLOCK, DIGI_MIN_BLOCKS) k_digitise(const __grid_constant__ ConfigK cfg,
                                                                    const __grid_constant__ KParams P, const Src src)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int        lane = threadIdx.x & 31;
    const int        wib = threadIdx.x >> 5;
    constexpr size_t wbytes = warp_shared_bytes(PC);
    unsigned char   *wbase = smem_raw + wbytes * wib;
    WarpShared      &W = *reinterpret_cast<WarpShared *>(wbase);
    double          *pool_g = reinterpret_cast<double *>(wbase + sizeof(WarpShared));
    double          *pool_e = pool_g + PC;
    // the block's copy of the 2^(j/64) table (exp_nonpos), behind the warps' regions
    double          *exp2_tab = reinterpret_cast<double *>(smem_raw + wbytes * WARPS_PER_BLOCK);
    if (threadIdx.x < 64) exp2_tab[threadIdx.x] = cfg.exp2_tab[threadIdx.x];
    __syncthreads();   // the only block-wide barrier: before the warps go their own ways
    for (;;) {
        unsigned long long group = 0;
        if (lane == 0) group = atomicAdd(&P.ws.alloc[ALLOC_TICKET * 16], 1ull);
        group = __shfl_sync(FULL, group, 0);
        if (group >= P.n_groups) return;   // whole warps leave; no block-wide barrier is used anywhere
        digitise_group<Src, PC, MODE>(cfg, P, src, group, W, pool_g, pool_e, exp2_tab, lane);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------
// k_digitise_big: the (group, plane) pairs k_digitise left aside because a cluster of theirs has more box columns + rows
// than the profile pool holds (pixel pitch far below the cloud width).  One warp per pair; the particles are served one
// after the other, the warp walking the pixels of one particle 32 per step and working out both profiles per pixel
// with the very expressions of phase A -- no pool, no limit but the 2^20 pixels of a bounding box.  Launched only when
// the geometry or the caller's own variates make such clusters possible.
// ---------------------------------------------------------------------------------------------------
constexpr int BIG_WARPS = 2;

template <class Src, bool KEEP64>
__global__ void __launch_bounds__(32 * BIG_WARPS) k_digitise_big(const __grid_constant__ ConfigK cfg,
                                                                 const __grid_constant__ KParams P, const Src src)
{
    __shared__ WarpShared Ws[BIG_WARPS];
    __shared__ double     exp2_tab[64];
    static_assert(32 * BIG_WARPS == 64, "one thread per table entry");
    exp2_tab[threadIdx.x] = cfg.exp2_tab[threadIdx.x];   // 32 * BIG_WARPS = 64 threads
    __syncthreads();
    const int                lane = threadIdx.x & 31;
    WarpShared              &W = Ws[threadIdx.x >> 5];
    const int                D = cfg.n_planes;
    const bool               from_table = (P.flags & PTB_TRACKS_FROM_TABLE) != 0;
    const unsigned long long n_items = P.ws.alloc[ALLOC_DEFERRED * 16];
    for (unsigned long long item = (unsigned long long)blockIdx.x * BIG_WARPS + (threadIdx.x >> 5); item < n_items;
         item += (unsigned long long)gridDim.x * BIG_WARPS) {
        const unsigned long long entry = P.ws.deferred[item];
        const unsigned long long group = entry >> 8;
        const int                d = (int)(entry & 0xffu);
        const PlaneK            &pl = cfg.pl[d];
        const unsigned long long kl = group * 32 + lane;
        const bool               valid = kl < P.n;
        const unsigned long long k = P.first + kl;
        const bool               compact = (P.flags & PTB_COMPACT) != 0;
        const unsigned           accmask = P.ws.accmask[group];   // k_tracks' ballot, or k_digitise's copy of the table's
        __syncwarp();
        W.evoff[lane] = (uint8_t)__popc(accmask & lanemask_lt());
        double x = 0.0, y = 0.0, eloss = 0.0;
        if (from_table) {
            if (valid) {
                x = P.trk.x[kl * D + d];
                y = P.trk.y[kl * D + d];
                eloss = P.trk.eloss[(unsigned long long)d * P.n + kl];
            }
        } else {
            const double *park = P.ws.park + group * (unsigned long long)(D * 96) + lane;
            x = park[d * 96];
            y = park[d * 96 + 32];
            eloss = park[d * 96 + 64];
        }
        const bool digitised = valid && (!pl.triggered || ((accmask >> lane) & 1u));
        Cluster    cl{0, 0, 0, false, false};
        if (digitised) cl = describe(cfg, pl, P.flags, src, k, d, x, y, eloss, W, lane);
        __syncwarp();
        const unsigned     insidemask = __ballot_sync(FULL, cl.inside);
        const int          ub = __reduce_add_sync(FULL, cl.inside ? max(cl.ncT * cl.nrT, 1) : 0);
        unsigned long long base = 0;
        if (lane == 0 && ub > 0) base = atomicAdd(&P.ws.alloc[d * 16], (unsigned long long)ub);
        base = __shfl_sync(FULL, base, 0);
        const bool room = base + (unsigned long long)ub <= P.ws.cap[d];
        uint2     *seg_hit = P.ws.s_hit[d] + base;
        uint8_t   *seg_ev = P.ws.s_ev[d] + base;
        double    *seg_q64 = KEEP64 ? P.ws.s_q64[d] + base : nullptr;
        int        n_plane = 0;
        for (unsigned left = insidemask; left; left &= left - 1) {
            const int    b = __ffs(left) - 1;   // the particles in order
            const int    ncl = W.ncols[b], nrw = W.nrows[b], box = ncl * nrw;
            const double px = W.pos[0][b], py = W.pos[1][b], rx = W.rad[0][b], ry = W.rad[1][b], q = W.q[b];
            int          kept = 0;
            for (int p0 = 0; p0 < box; p0 += 32) {
                const int p = p0 + lane;
                bool      keep = false;
                double    hq = 0.0;
                uint32_t  hcr = 0;
                if (p < box) {
                    const int ci = p / nrw, ri = p - ci * nrw;
                    double    gc, ec, gr, er;
                    profile_entry(pl.ax[0], W.lo[0][b] + ci, px, rx, W.nrm[0][b], W.ye[0][b], exp2_tab, gc, ec);
                    profile_entry(pl.ax[1], W.lo[1][b] + ri, py, ry, W.nrm[1][b], W.ye[1][b], exp2_tab, gr, er);
                    const double zn = src.pixel(P.first + group * 32 + b, d, (uint32_t)p,
                                                (uint32_t)(Src::kPixelKeyed ? W.jbase[b] : W.jbase[b] + ci * W.nrfull[b] + ri));
                    const double noise = 0.0 + pl.noise_sigma * zn;
                    const double signal = q * (gc * gr);
                    hq = (signal + noise) * 1e-3 * pl.pitch_c * pl.pitch_r;   // device.py:275-288
                    keep = hq >= pl.thr_ke && ec + er <= 1.0;
                    hcr = (uint32_t)(W.lo[0][b] + ci) | ((uint32_t)(W.lo[1][b] + ri) << 16);
                }
                const unsigned kballot = __ballot_sync(FULL, keep);
                if (keep && room) {
                    const unsigned at = (unsigned)(n_plane + kept) + __popc(kballot & lanemask_lt());
                    PTB_DBG_CHECK((int)at < ub && base + at < P.ws.cap[d]);
                    seg_hit[at] = make_uint2(hcr, __float_as_uint((float)hq));
                    if (!compact) seg_ev[at] = W.evoff[b];
                    if (KEEP64) seg_q64[at] = hq;
                }
                kept += __popc(kballot);
            }
            if (kept == 0 && W.seedq[b] >= pl.thr_ke) {   // seed fallback, main/device.py:299-302
                if (lane == 0 && room) {
                    seg_hit[n_plane] = make_uint2(W.seedcr[b], __float_as_uint((float)W.seedq[b]));
                    if (!compact) seg_ev[n_plane] = W.evoff[b];
                    if (KEEP64) seg_q64[n_plane] = W.seedq[b];
                }
                kept = 1;
            }
            if (compact && lane == 0) {
                if (P.flags & (PTB_COMPACT_NARROW | PTB_COMPACT_PACKED)) {   // the particle's nibble of the byte k_digitise left (zero) for the pair
                    uint8_t  *byte = &P.cmp.counts[(unsigned long long)d * ((P.n + 1) / 2) + (group * 32 + b) / 2];
                    const int sh = (b & 1) * 4;
                    *byte = (uint8_t)((*byte & ~(15u << sh)) | ((unsigned)min(kept, 15) << sh));
                    if (kept >= 15) escape_count(P, d, group * 32 + b, (unsigned)min(kept, 255));
                    if (kept > 255) atomicOr(&P.totals->error, PTB_DEV_COUNT_OVERFLOW);
                } else {
                    P.cmp.counts[(unsigned long long)d * P.n + group * 32 + b] = (uint8_t)min(kept, 255);
                    if (kept > 255) atomicOr(&P.totals->error, PTB_DEV_COUNT_OVERFLOW);
                }
            }
            n_plane += kept;
        }
        if (lane == 0) {
            P.ws.table[(unsigned long long)d * P.n_groups + group] = (unsigned long long)n_plane;
            P.ws.table[(unsigned long long)(D + 5 + d) * P.n_groups + group] = room ? base : NO_SEGMENT;
            if (!room) atomicOr(&P.totals->error, PTB_DEV_SCRATCH_OVERFLOW);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------
// k_scan: block c scans column c of the group table (exclusive prefix into ws.prefix for c < D+2), writes
// this call's totals, the bases the gather pass has to use, and the bases of the following range
// ---------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;

__device__ __forceinline__ void scan_segment_bounds(unsigned long long n_groups, int seg, unsigned long long &lo,
                                                    unsigned long long &hi)
{
    const unsigned long long len = (n_groups + SCAN_SEGMENTS - 1) / SCAN_SEGMENTS;
    lo = min(n_groups, len * (unsigned long long)seg);
    hi = min(n_groups, lo + len);
}

// pass 1: sum of every segment of every column
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(unsigned long long n_groups, Workspace ws)
{
    __shared__ unsigned long long warp_sum[32];
    const int                 c = blockIdx.x / SCAN_SEGMENTS, seg = blockIdx.x % SCAN_SEGMENTS;
    const unsigned long long *col = ws.table + (unsigned long long)c * n_groups;
    unsigned long long        lo, hi, sum = 0;
    scan_segment_bounds(n_groups, seg, lo, hi);
    for (unsigned long long i = lo + threadIdx.x; i < hi; i += SCAN_THREADS) sum += col[i];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(FULL, sum, s);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x < 32) {
        sum = warp_sum[threadIdx.x];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(FULL, sum, s);
        if (threadIdx.x == 0) ws.segsum[blockIdx.x] = sum;
    }
}

// pass 2: block (c, seg) scans its segment of column c, seeded with the sums of the segments before it
__global__ void __launch_bounds__(SCAN_THREADS) k_scan(int D, unsigned long long n_groups, uint32_t flags, Workspace ws,
                                                       const PtbBases *bases_in, PtbBases *bases_out, PtbTotals *totals)
{
    __shared__ unsigned long long warp_sum[32];
    __shared__ unsigned long long carry_s;
    const int                     c = blockIdx.x / SCAN_SEGMENTS, seg = blockIdx.x % SCAN_SEGMENTS;
    const unsigned long long     *col = ws.table + (unsigned long long)c * n_groups;
    unsigned long long           *out = (c < D + 2) ? ws.prefix + (unsigned long long)c * n_groups : nullptr;
    const int                     lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long            lo, hi;
    scan_segment_bounds(n_groups, seg, lo, hi);
    if (threadIdx.x == 0) {
        unsigned long long before = 0;
        for (int s2 = 0; s2 < seg; ++s2) before += ws.segsum[c * SCAN_SEGMENTS + s2];
        carry_s = before;
    }
    __syncthreads();
    if (out == nullptr && seg != SCAN_SEGMENTS - 1) return;   // census columns only need their total
    for (unsigned long long tile = lo; tile < hi; tile += (unsigned long long)SCAN_THREADS * SCAN_ITEMS) {
        const unsigned long long i0 = tile + (unsigned long long)threadIdx.x * SCAN_ITEMS;
        unsigned long long       v[SCAN_ITEMS], sum = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            v[i] = (i0 + i < hi) ? col[i0 + i] : 0ull;
            sum += v[i];
        }
        unsigned long long incl = sum;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            unsigned long long o = __shfl_up_sync(FULL, incl, s);
            if (lane >= s) incl += o;
        }
        if (lane == 31) warp_sum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            unsigned long long ws_incl = warp_sum[lane];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                unsigned long long o = __shfl_up_sync(FULL, ws_incl, s);
                if (lane >= s) ws_incl += o;
            }
            warp_sum[lane] = ws_incl;
        }
        __syncthreads();
        const unsigned long long carry = carry_s;
        unsigned long long       excl = carry + (wid ? warp_sum[wid - 1] : 0ull) + (incl - sum);
        if (out) {
#pragma unroll
            for (int i = 0; i < SCAN_ITEMS; ++i) {
                if (i0 + i < hi) out[i0 + i] = excl;
                excl += v[i];
            }
        }
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) carry_s = carry + warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0 && seg == SCAN_SEGMENTS - 1) {
        const unsigned long long total = carry_s;
        if (c < D) {
            const bool               recycle = (flags & PTB_RECYCLE_SLABS) != 0;
            const unsigned long long b = (bases_in && !recycle) ? bases_in->hits[c] : 0ull;
            ws.bases_used->hits[c] = b;
            totals->hits[c] = total;
            totals->reserved[c] = ws.alloc[c * 16];   // scratch rows the groups asked for (k_digitise has finished)
            if (bases_out) bases_out->hits[c] = recycle ? 0ull : b + total;
        } else if (c == D) {
            const int64_t b = bases_in ? bases_in->event : 0;
            ws.bases_used->event = b;
            totals->accepted = total;
            if (bases_out) bases_out->event = b + (int64_t)total;
        } else if (c == D + 1) {
            const int64_t b = bases_in ? bases_in->time : 0;
            ws.bases_used->time = b;
            totals->time_sum = (int64_t)total;
            if (bases_out) bases_out->time = b + (int64_t)total;
        } else if (c == D + 2) {
            totals->pairs = total;
        } else if (c == D + 3) {
            totals->inside = total;
        } else if (c == D + 4) {
            totals->pixels = total;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// k_gather: one warp per group; scratch segment -> final particle-ordered slab rows, event numbers,
// absolute time stamps of the track records
// ---------------------------------------------------------------------------------------------------
struct GatherParams {
    int                D;
    uint32_t           flags;
    unsigned long long n, n_groups;
    Workspace          ws;
    PtbHitSlab         slabs[PTB_MAX_PLANES];
    PtbCompactOut      cmp;
    PtbPackedLayout    packed;   // PTB_COMPACT_PACKED
    int64_t           *time_stamp;
    PtbTotals         *totals;
};

__global__ void __launch_bounds__(256, 8) k_gather(const __grid_constant__ GatherParams G)
{
    // per-plane pointers in shared memory: lanes of one warp address different planes at the same time
    __shared__ const uint2    *s_hit[PTB_MAX_PLANES];
    __shared__ const uint8_t  *s_ev[PTB_MAX_PLANES];
    __shared__ const double   *s_q64[PTB_MAX_PLANES];
    __shared__ int64_t        *d_event[PTB_MAX_PLANES];
    __shared__ uint16_t       *d_col[PTB_MAX_PLANES], *d_row[PTB_MAX_PLANES];
    __shared__ float          *d_charge[PTB_MAX_PLANES];
    __shared__ double         *d_q64[PTB_MAX_PLANES];
    const int  D = G.D;
    const bool compact = (G.flags & PTB_COMPACT) != 0;
    if (threadIdx.x < D) {
        const int d = threadIdx.x;
        s_hit[d] = G.ws.s_hit[d]; s_ev[d] = G.ws.s_ev[d]; s_q64[d] = G.ws.s_q64[d];
        d_event[d] = G.slabs[d].event; d_col[d] = G.slabs[d].col; d_row[d] = G.slabs[d].row;
        d_charge[d] = G.slabs[d].charge; d_q64[d] = G.slabs[d].charge64;
    }
    __syncthreads();
    const int                lane = threadIdx.x & 31;
    const unsigned long long group = (unsigned long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (group >= G.n_groups) return;
    const PtbBases &B = *G.ws.bases_used;
    if (G.flags & PTB_DIGITISE) {
        // PTB_COMPACT_NARROW: the planes' rows lie back to back, plane d starts where the rows of planes 0..d-1 end
        // (k_scan has written this call's totals)
        const bool         packed = (G.flags & PTB_COMPACT_PACKED) != 0;
        const bool         narrow = (G.flags & PTB_COMPACT_NARROW) != 0 || packed;   // dense rows, planes back to back
        unsigned long long dense_l = 0;
        if (narrow) {
            const unsigned long long mine = lane < D ? G.totals->hits[lane] : 0ull;
            unsigned long long       incl = mine;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const unsigned long long o = __shfl_up_sync(FULL, incl, s);
                if (lane >= s) incl += o;
            }
            dense_l = incl - mine;
        }
        // lane d fetches the segment descriptor of plane d: all table look-ups of the group are in flight at once
        unsigned long long n_l = 0, src_l = 0, dst_l = 0;
        if (lane < D) {
            n_l = G.ws.table[(unsigned long long)lane * G.n_groups + group];
            src_l = G.ws.table[(unsigned long long)(D + 5 + lane) * G.n_groups + group];
            dst_l = (compact ? 0ull : B.hits[lane]) + G.ws.prefix[(unsigned long long)lane * G.n_groups + group];
            const unsigned long long room_l = compact ? G.cmp.plane_first[lane + 1] - G.cmp.plane_first[lane]
                                                      : G.slabs[lane].capacity;
            // a group whose scratch reservation did not fit wrote no rows (NO_SEGMENT): nothing of it may be copied
            const bool no_src = src_l == NO_SEGMENT || src_l + n_l > G.ws.cap[lane];
            // (narrow form: a plane that outgrew its share pushes the later planes' rows towards the end of the arrays)
            const bool no_room = dst_l + n_l > room_l || (narrow && dense_l + dst_l + n_l > G.cmp.plane_first[D]);
            if (n_l != 0 && (no_room || no_src)) {
                atomicOr(&G.totals->error, no_room ? PTB_DEV_SLAB_OVERFLOW : PTB_DEV_SCRATCH_OVERFLOW);
                n_l = 0;
            }
        }
        if (compact && lane == 0) G.cmp.accepted[group] = G.ws.accmask[group];
        const int64_t ev_base = B.event + (int64_t)G.ws.prefix[(unsigned long long)D * G.n_groups + group] + 1;
        const bool    keep64 = (G.flags & PTB_KEEP_CHARGE64) != 0;
        // the group's segments of all planes form one work list, walked 32 rows at a time: every step is a full
        // warp of independent loads, whatever the sizes of the individual segments
        int cum = (int)n_l;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            int o = __shfl_up_sync(FULL, cum, s);
            if (lane >= s) cum += o;
        }
        const int total = __shfl_sync(FULL, cum, 31);
        // The planes that have rows move to the low lanes, in order: lane r holds the r-th of them -- one past its last
        // row in the work list, and its source / destination positions shifted by its first row, so that row w of the
        // list is read at src + w and written at dst + w.  With the empty planes gone the ends are distinct, and the
        // plane of a row is found as in k_digitise: count the planes that ended before the step (a ballot) plus those
        // ending inside it below the row (one warp-wide OR of their end bits).
        const unsigned nz = __ballot_sync(FULL, n_l != 0);
        const int      R = __popc(nz);
        const int      from = lane < R ? (int)__fns(nz, 0, lane + 1) : 0;
        const int      end_from = __shfl_sync(FULL, cum, from);           // (every lane takes part in the shuffles)
        const int      endR = lane < R ? end_from : 0x7fffffff;
        const int      firstR = end_from - (int)__shfl_sync(FULL, (unsigned)n_l, from);
        const unsigned long long dstAll = dst_l + (narrow ? dense_l : (compact && lane < D) ? G.cmp.plane_first[lane] : 0ull);
        const unsigned long long srcR = __shfl_sync(FULL, src_l, from) - (unsigned long long)firstR;
        const unsigned long long dstR = __shfl_sync(FULL, dstAll, from) - (unsigned long long)firstR;
        for (int w0 = 0; w0 < total; w0 += 32) {
            const int      w = w0 + lane;
            const unsigned before = __ballot_sync(FULL, endR <= w0);
            const int      rel = endR - w0;
            const unsigned ends = __reduce_or_sync(FULL, (rel >= 1 && rel <= 31) ? 1u << rel : 0u);
            const int      rank = min(__popc(before) + __popc(ends & ((2u << lane) - 1u)), R - 1);
            const int      d = __shfl_sync(FULL, from, rank);
            const unsigned long long si = __shfl_sync(FULL, srcR, rank) + (unsigned long long)w;
            const unsigned long long o = __shfl_sync(FULL, dstR, rank) + (unsigned long long)w;
            if (w >= total) continue;
            const uint2 hit = s_hit[d][si];
            if (packed) {
                // 48 bits: f32 mantissa | exponent above the plane's threshold << 23 | column << 27 | row << (27 + col_bits)
                const unsigned ex = (hit.y >> 23) & 0xffu, ex_min = G.packed.exp_min[d];
                unsigned       off = ex - ex_min;
                if (ex < ex_min || (hit.y >> 31)) atomicOr(&G.totals->error, PTB_DEV_PACK_OVERFLOW);   // (cannot happen)
                if (off > 15u) {
                    // a charge beyond the 16 binades of the row format (the far tail of the deposit law): the row
                    // keeps exponent 15 and the charge travels on the escape list under the row's index
                    const unsigned slot = atomicAdd(&G.totals->escapes, 1u);
                    if (slot < G.cmp.escape_capacity)
                        G.cmp.escapes[slot] = (1ull << 63) | ((o & 0x7fffffffull) << 32) | (unsigned long long)hit.y;
                    else
                        atomicOr(&G.totals->error, PTB_DEV_COUNT_OVERFLOW);
                    off = 15u;
                }
                const unsigned long long word = (unsigned long long)(hit.y & 0x7fffffu) | ((unsigned long long)(off & 15u) << 23) |
                                                ((unsigned long long)(hit.x & 0xffffu) << 27) |
                                                ((unsigned long long)(hit.x >> 16) << (27 + G.packed.col_bits[d]));
                G.cmp.word0[o] = (uint16_t)word;
                G.cmp.word1[o] = (uint16_t)(word >> 16);
                G.cmp.word2[o] = (uint16_t)(word >> 32);
            } else if (narrow) {
                const unsigned pos = (hit.x & 0xfffu) | ((hit.x >> 16) << 12);   // column | row << 12
                G.cmp.charge[o] = __uint_as_float(hit.y);
                G.cmp.pos_lo[o] = (uint16_t)pos;
                G.cmp.pos_hi[o] = (uint8_t)(pos >> 16);
            } else if (compact) {
                G.cmp.hits[o] = (unsigned long long)hit.x | ((unsigned long long)hit.y << 32);
            } else {
                d_event[d][o] = ev_base + (int64_t)s_ev[d][si];
                d_col[d][o] = (uint16_t)(hit.x & 0xffffu);
                d_row[d][o] = (uint16_t)(hit.x >> 16);
                d_charge[d][o] = __uint_as_float(hit.y);
                if (keep64) d_q64[d][o] = s_q64[d][si];
            }
        }
    }
    if ((G.flags & PTB_WRITE_TRACKS) && !(G.flags & PTB_TRACKS_FROM_TABLE)) {
        const int64_t            t_base = B.time + (int64_t)G.ws.prefix[(unsigned long long)(D + 1) * G.n_groups + group];
        const unsigned long long lo = group * 32 * D;
        const unsigned long long hi = min((group + 1) * 32, G.n) * D;
        for (unsigned long long i = lo + lane; i < hi; i += 32) G.time_stamp[i] += t_base;
    }
}

// ---------------------------------------------------------------------------------------------------
// k_prefix: accepted-event count and gap sum of a particle range (ptb_prefix)
// ---------------------------------------------------------------------------------------------------
template <class Src>
__global__ void __launch_bounds__(256) k_prefix(const __grid_constant__ ConfigK cfg, unsigned long long first,
                                                unsigned long long n, const Src src, long long *out)
{
    long long acc = 0, tsum = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        acc += src.accepted(cfg, first + i) ? 1 : 0;
        tsum += src.gap(cfg, first + i);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        acc += __shfl_xor_sync(FULL, acc, s);
        tsum += __shfl_xor_sync(FULL, tsum, s);
    }
    __shared__ long long sa[8], st[8];
    if ((threadIdx.x & 31) == 0) {
        sa[threadIdx.x >> 5] = acc;
        st[threadIdx.x >> 5] = tsum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) {
            acc += sa[i];
            tsum += st[i];
        }
        atomicAdd((unsigned long long *)&out[0], (unsigned long long)acc);
        atomicAdd((unsigned long long *)&out[1], (unsigned long long)tsum);
    }
}

// ---------------------------------------------------------------------------------------------------
// k_pack: SoA slab rows -> the reference's 18-byte records (main/device.py:20-28), 256 rows per block
// staged in shared memory so that the global stores are contiguous
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack(PtbHitSlab s, unsigned long long n, uint16_t *out)
{
    __shared__ __align__(16) uint16_t tile[256 * 9];
    const unsigned long long          r0 = (unsigned long long)blockIdx.x * 256;
    const unsigned long long          r = r0 + threadIdx.x;
    if (r < n) {
        const unsigned long long ev = (unsigned long long)s.event[r];
        const uint32_t           qb = __float_as_uint(s.charge[r]);
        uint16_t                *t = tile + threadIdx.x * 9;
        t[0] = (uint16_t)ev; t[1] = (uint16_t)(ev >> 16); t[2] = (uint16_t)(ev >> 32); t[3] = (uint16_t)(ev >> 48);
        t[4] = 0;
        t[5] = s.col[r];
        t[6] = s.row[r];
        t[7] = (uint16_t)qb; t[8] = (uint16_t)(qb >> 16);
    }
    __syncthreads();
    const unsigned long long rows = min((unsigned long long)256, n - r0);
    uint16_t                *dst = out + r0 * 9;
    if (rows == 256 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const uint4 *t4 = reinterpret_cast<const uint4 *>(tile);
        uint4       *d4 = reinterpret_cast<uint4 *>(dst);
        for (int i = threadIdx.x; i < 256 * 18 / 16; i += 256) d4[i] = t4[i];
    } else {
        for (unsigned long long i = threadIdx.x; i < rows * 9; i += 256) dst[i] = tile[i];
    }
}

// ---------------------------------------------------------------------------------------------------
// k_digest: checksums of a hit table (ptb_digest): five plain sums and two sums over a 64-bit mix of every row, the
// second weighted with the row's position so that it notices rows in the wrong order
// ---------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long row_mix(long long ev, unsigned col, unsigned row, unsigned qbits)
{
    unsigned long long h = (unsigned long long)ev * 0x9E3779B97F4A7C15ull;
    h ^= (unsigned long long)col | ((unsigned long long)row << 16) | ((unsigned long long)qbits << 32);
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    return h;
}

__global__ void __launch_bounds__(256) k_digest(PtbHitSlab s, unsigned long long n, unsigned long long row_base,
                                                unsigned long long *out)
{
    unsigned long long a[7] = {0, 0, 0, 0, 0, 0, 0};
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const long long ev = s.event[i];
        const unsigned  c = s.col[i], r = s.row[i], q = __float_as_uint(s.charge[i]);
        const unsigned long long m = row_mix(ev, c, r, q);
        a[0] += 1ull;
        a[1] += (unsigned long long)ev;
        a[2] += c;
        a[3] += r;
        a[4] += q;
        a[5] += m;
        a[6] += (row_base + i + 1ull) * m;
    }
    __shared__ unsigned long long part[8][7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) a[j] += __shfl_xor_sync(FULL, a[j], sft);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5][j] = a[j];
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += part[w][threadIdx.x];
        atomicAdd(&out[threadIdx.x], t);
    }
}

// ---------------------------------------------------------------------------------------------------
// k_correlate: one thread per event number; the hits of that event in both tables (plus the first hit of the next
// event, as the reference's loop does) are found by binary search and their pairs counted with atomics
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long first_at_least(const int64_t *ev, unsigned long long n, int64_t v)
{
    unsigned long long lo = 0, hi = n;
    while (lo < hi) {
        const unsigned long long mid = (lo + hi) >> 1;
        if (ev[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_correlate(const int64_t *ev1, const uint16_t *col1, const uint16_t *row1,
                                                   unsigned long long n1, const int64_t *ev2, const uint16_t *col2,
                                                   const uint16_t *row2, unsigned long long n2, int64_t ev_min,
                                                   int64_t ev_max, int *xh, int xd0, int xd1, int *yh, int yd0, int yd1)
{
    for (int64_t i = ev_min + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ev_max;
         i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long a0 = first_at_least(ev1, n1, i), a1 = first_at_least(ev1, n1, i + 1);
        const unsigned long long b0 = first_at_least(ev2, n2, i), b1 = first_at_least(ev2, n2, i + 1);
        for (unsigned long long a = a0; a <= a1 && a < n1; ++a) {
            const int c1 = col1[a], r1 = row1[a];
            for (unsigned long long b = b0; b <= b1 && b < n2; ++b) {
                const int c2 = col2[b], r2 = row2[b];
                if (c1 < xd0 && c2 < xd1) atomicAdd(&xh[(long long)c1 * xd1 + c2], 1);
                if (r1 < yd0 && r2 < yd1) atomicAdd(&yh[(long long)r1 * yd1 + r2], 1);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// k_dfma: FP64 pipe probe for the roofline denominator
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b)
{
    double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
    for (int i = 0; i < iters; ++i) {
        v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
        v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
}

}  // namespace ptb

template <class Src, int PC, int MODE>
static cudaError_t launch_digitise(unsigned blocks, cudaStream_t st, const ptb::ConfigK &c, const ptb::KParams &P,
                                   const Src &src)
{
    using namespace ptb;
    constexpr size_t smem = warp_shared_bytes(PC) * WARPS_PER_BLOCK + 64 * sizeof(double);   // + the exp2 table
    k_digitise<Src, PC, MODE><<<blocks, BLOCK, smem, st>>>(c, P, src);
    return cudaGetLastError();
}

template <class Src, int PC>
static cudaError_t launch_digitise_mode(int mode, unsigned blocks, cudaStream_t st, const ptb::ConfigK &c,
                                        const ptb::KParams &P, const Src &src)
{
    return mode == 2 ? launch_digitise<Src, PC, 2>(blocks, st, c, P, src)
         : mode == 1 ? launch_digitise<Src, PC, 1>(blocks, st, c, P, src)
                     : launch_digitise<Src, PC, 0>(blocks, st, c, P, src);
}

// Shared-memory carve-out of every k_digitise variant: asked for once per handle (ptb_create), not per launch.  The
// dynamic shared memory stays below the 48 KB that need no opt-in.
template <class Src, int PC, int MODE>
static cudaError_t prefer_shared()
{
    static_assert(ptb::warp_shared_bytes(PC) * ptb::WARPS_PER_BLOCK + 512 <= 48 * 1024, "k_digitise needs the shared-memory opt-in");
    return cudaFuncSetAttribute(ptb::k_digitise<Src, PC, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared);
}

static cudaError_t prefer_shared_all()
{
    using namespace ptb;
    cudaError_t e = cudaSuccess;
    auto also = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    also(prefer_shared<PhiloxSource, POOL_ENTRIES, 0>());
    also(prefer_shared<PhiloxSource, POOL_ENTRIES, 1>());
    also(prefer_shared<PhiloxSource, POOL_ENTRIES, 2>());
    also(prefer_shared<PhiloxSource, SMALL_POOL, 0>());
    also(prefer_shared<PhiloxSource, SMALL_POOL, 1>());
    also(prefer_shared<PhiloxSource, SMALL_POOL, 2>());
    also(prefer_shared<InjectedSource, POOL_ENTRIES, 0>());
    also(prefer_shared<InjectedSource, POOL_ENTRIES, 1>());
    also(prefer_shared<InjectedSource, POOL_ENTRIES, 2>());
    also(prefer_shared<InjectedSource, SMALL_POOL, 0>());
    also(prefer_shared<InjectedSource, SMALL_POOL, 1>());
    also(prefer_shared<InjectedSource, SMALL_POOL, 2>());
    return e;
}

// =====================================================================================================
// host side: handle, derived constants, C ABI
// =====================================================================================================
using namespace ptb;

struct PtbContext {
    ConfigK cfg;
    void   *alias_dev;   // device copy of the Poisson alias table (cfg.alias points at it)
    int     device;
    int     sm_count;
    int     may_be_big;   // a production-mode cluster can outgrow the profile pool (sizes k_digitise_big's grid)
    int     timing;
    std::vector<cudaEvent_t> ev;   // TIMING_EVENTS per ptb_simulate (start, after k_tracks, after k_digitise, after
                                   // k_digitise_big + scans, after k_gather), when timing is enabled
    size_t  ev_used;
    char    err[512];
};

static int fail(PtbHandle h, int code, const char *msg)
{
    if (h) snprintf(h->err, sizeof(h->err), "%s", msg);
    return code;
}

static int fail_cuda(PtbHandle h, cudaError_t e, const char *where)
{
    if (h) snprintf(h->err, sizeof(h->err), "%s: %s", where, cudaGetErrorString(e));
    return PTB_E_CUDA;
}

// Moyal-form density of the reference in the energy-loss variable (unnormalised), main/hit.py:399-400,431
static double moyal_density(double delta, double loc, double scale)
{
    double lam = (delta - loc) / scale;
    return exp(-(lam + exp(-lam)) / 2);
}

// F(lambda) = P(Lambda <= lambda) for the Moyal law = erfc(exp(-lambda/2)/sqrt 2)
static double moyal_cdf(double lam) { return erfc(exp(-lam / 2) / sqrt(2.0)); }

// Where the expanded 18-byte records of one thread go: straight to memory, or block by block into a file.  A thread
// writes its rows strictly in order, so the file writer only ever holds one partly filled block.
struct RowsToMemory {
    unsigned char *dst;
    unsigned char *at(uint64_t i) { return dst + i * 18; }      // row i, final
    unsigned char *peek2(uint64_t i) { return dst + i * 18; }   // room for rows i and i + 1, not final yet
    void           commit(unsigned) {}
    bool           finish() { return true; }
};

#ifndef PTB_FILE_BLOCK_ROWS
#define PTB_FILE_BLOCK_ROWS 233016   /* 4 MiB of records: measured best on a disk-backed page cache (256 KiB .. 16 MiB tried) */
#endif
struct RowsToFile {
    static constexpr size_t BLOCK_ROWS = PTB_FILE_BLOCK_ROWS;
    int                     fd;
    uint64_t                offset;                // file position of row 0 of the plane
    std::vector<unsigned char> buf;
    uint64_t                first = 0;
    size_t                  used = 0;
    bool                    ok = true;
    RowsToFile(int fd_, uint64_t offset_) : fd(fd_), offset(offset_), buf(BLOCK_ROWS * 18) {}
    void flush()
    {
        const unsigned char *p = buf.data();
        size_t               left = used * 18;
        uint64_t             at = offset + first * 18;
        while (left && ok) {
            const ssize_t done = pwrite(fd, p, left, (off_t)at);
            if (done <= 0) {
                ok = false;
                break;
            }
            p += done;
            at += (uint64_t)done;
            left -= (size_t)done;
        }
        used = 0;
    }
    unsigned char *at(uint64_t i)
    {
        if (used == BLOCK_ROWS) flush();
        if (used == 0) first = i;
        return buf.data() + 18 * used++;
    }
    unsigned char *peek2(uint64_t i)
    {
        if (used + 2 > BLOCK_ROWS) flush();
        if (used == 0) first = i;
        return buf.data() + 18 * used;
    }
    void commit(unsigned c) { used += c; }
    bool finish()
    {
        if (used) flush();
        return ok;
    }
};

// Rows: how a compressed row is read (hit i of the plane -> column, row, charge bits); Counts: rows of particle k;
// make_out(): the per-thread destination (RowsToMemory / RowsToFile).
template <class Rows, class Counts, class MakeOut>
static int64_t expand_rows(const Rows &rows_of, const Counts &count_of, const uint32_t *accepted, uint64_t n,
                           int64_t event_base, uint64_t n_rows_expected, const MakeOut &make_out, int32_t threads)
{
    int T = threads > 0 ? threads : (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
    const uint64_t words = (n + 31) / 32;
    T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)T, (words + 4095) / 4096));   // >= 128k particles per thread
    const uint64_t          per = (words + T - 1) / T;   // words per thread
    std::vector<uint64_t>   rows(T + 1, 0), acc(T + 1, 0);
    std::vector<int>        failed(T, 0);
    auto range = [&](int t, uint64_t &k0, uint64_t &k1) {
        k0 = std::min<uint64_t>(n, per * 32 * (uint64_t)t);
        k1 = std::min<uint64_t>(n, k0 + per * 32);
    };
    auto count_pass = [&](int t) {
        uint64_t k0, k1, r = 0, a = 0;
        range(t, k0, k1);
        for (uint64_t k = k0; k < k1; ++k) r += count_of(k);
        for (uint64_t w = k0 / 32; w < (k1 + 31) / 32; ++w) {
            uint32_t m = accepted[w];
            if ((w + 1) * 32 > n) m &= (uint32_t)((1ull << (n - w * 32)) - 1ull);   // bits beyond the range do not count
            a += (uint64_t)__builtin_popcount(m);
        }
        rows[t + 1] = r;
        acc[t + 1] = a;
    };
    auto run_threads = [&](auto fn) {
        std::vector<std::thread> pool;
        for (int t = 1; t < T; ++t) pool.emplace_back(fn, t);
        fn(0);
        for (auto &th : pool) th.join();
    };
    run_threads(count_pass);
    for (int t = 0; t < T; ++t) {
        rows[t + 1] += rows[t];
        acc[t + 1] += acc[t];
    }
    if (rows[T] != n_rows_expected) return PTB_E_INVALID;
    auto expand_pass = [&](int t) {
        uint64_t k0, k1;
        range(t, k0, k1);
        uint64_t i = rows[t];
        int64_t  ev = event_base + 1 + (int64_t)acc[t];
        auto     out = make_out();
        // event_number <i8 | frame <u2 = 0 | column <u2 | row <u2 | charge <f4 (main/device.py:20-28) as 8 + 8 + 2 bytes
        auto put = [&](unsigned char *r, uint64_t at_row) {
            uint16_t col, row;
            uint32_t q;
            rows_of(at_row, col, row, q);
            const uint64_t mid = ((uint64_t)col << 16) | ((uint64_t)row << 32) | ((uint64_t)(q & 0xffffu) << 48);
            const uint16_t top = (uint16_t)(q >> 16);
            memcpy(r, &ev, 8);
            memcpy(r + 8, &mid, 8);
            memcpy(r + 16, &top, 2);
        };
        const uint64_t end = rows[t + 1];
        uint64_t       k = k0;
        // Most particles leave 0, 1 or 2 rows, in no predictable order: two rows are written whatever the count, and
        // the position then moves on by the count -- what was written in excess is overwritten by the rows that
        // follow (1.6 times the rate of a loop over the count, whose trip count the branch predictor cannot learn).
        for (; k < k1 && i + 2 <= end; ++k) {
            const unsigned c = count_of(k);
            if (c <= 2) {
                unsigned char *r = out.peek2(i);
                put(r, i);
                put(r + 18, i + 1);
                out.commit(c);
                i += c;
            } else {
                for (unsigned j = 0; j < c; ++j, ++i) put(out.at(i), i);
            }
            ev += (accepted[k >> 5] >> (k & 31)) & 1u;
        }
        for (; k < k1; ++k) {   // the last rows of the range: nothing may be written beyond them
            const unsigned c = count_of(k);
            for (unsigned j = 0; j < c; ++j, ++i) put(out.at(i), i);
            ev += (accepted[k >> 5] >> (k & 31)) & 1u;
        }
        failed[t] = out.finish() ? 0 : 1;
    };
    run_threads(expand_pass);
    for (int t = 0; t < T; ++t)
        if (failed[t]) return PTB_E_IO;
    return (int64_t)rows[T];
}

// one plane of the wide / narrow compressed form as (rows_of, count_of) pairs handed to `go`
template <class Go>
static int64_t with_wide_plane(const uint64_t *hits, const uint8_t *counts, const Go &go)
{
    return go(
        [=](uint64_t i, uint16_t &col, uint16_t &row, uint32_t &q) {
            const uint64_t h = hits[i];
            col = (uint16_t)h;
            row = (uint16_t)(h >> 16);
            q = (uint32_t)(h >> 32);
        },
        [=](uint64_t k) -> unsigned { return counts[k]; });
}

// the 4-bit counts of the narrow and the packed form: a nibble of 15 is looked up among the plane's escape entries
struct NibbleCounts {
    const uint8_t                             *counts4;
    std::vector<std::pair<uint64_t, unsigned>> esc;   // this plane's escaped counts, sorted by particle
    NibbleCounts(const uint8_t *c4, const uint64_t *escapes, uint64_t n_escapes, int32_t plane) : counts4(c4)
    {
        for (uint64_t i = 0; i < n_escapes; ++i)
            if (!(escapes[i] >> 63) && (int32_t)((escapes[i] >> 40) & 0xffu) == plane)   // (bit 63: a charge, not a count)
                esc.emplace_back(escapes[i] & 0xffffffffffull, (unsigned)(escapes[i] >> 48));
        std::sort(esc.begin(), esc.end());
    }
};

template <class Go>
static int64_t with_packed_plane(const uint16_t *w0, const uint16_t *w1, const uint16_t *w2, const uint8_t *counts4,
                                 const uint64_t *escapes, uint64_t n_escapes, int32_t plane, const PtbPackedLayout &L,
                                 uint64_t row_base, uint64_t n_rows, const Go &go)
{
    const NibbleCounts nc(counts4, escapes, n_escapes, plane);
    const auto        *eb = nc.esc.data();
    const size_t       ne = nc.esc.size();
    // charges beyond the row format's 16 binades, by row of this plane
    std::vector<std::pair<uint64_t, uint32_t>> big;
    for (uint64_t i = 0; i < n_escapes; ++i)
        if (escapes[i] >> 63) {
            const uint64_t at = (escapes[i] >> 32) & 0x7fffffffull;
            if (at >= row_base && at - row_base < n_rows) big.emplace_back(at - row_base, (uint32_t)escapes[i]);
        }
    std::sort(big.begin(), big.end());
    const auto    *bb = big.data();
    const size_t   nb = big.size();
    const unsigned ex = L.exp_min[plane], cb = L.col_bits[plane];
    return go(
        [=](uint64_t i, uint16_t &col, uint16_t &row, uint32_t &q) {
            const uint64_t w = (uint64_t)w0[i] | ((uint64_t)w1[i] << 16) | ((uint64_t)w2[i] << 32);
            const uint32_t off = (uint32_t)((w >> 23) & 15u);
            q = (uint32_t)(w & 0x7fffffu) | ((ex + off) << 23);
            if (off == 15u && nb) {   // the top binade, or beyond it: then the charge stands on the escape list
                const auto *it = std::lower_bound(bb, bb + nb, std::make_pair(i, (uint32_t)0));
                if (it != bb + nb && it->first == i) q = it->second;
            }
            col = (uint16_t)((w >> 27) & ((1u << cb) - 1u));
            row = (uint16_t)(w >> (27 + cb));
        },
        [=](uint64_t k) -> unsigned {
            const unsigned c = (counts4[k >> 1] >> ((k & 1) * 4)) & 15u;
            if (c < 15u) return c;
            const auto *it = std::lower_bound(eb, eb + ne, std::make_pair(k, 0u));
            return (it != eb + ne && it->first == k) ? it->second : 0xffffffffu;   // no entry: the sum check fails
        });
}

template <class Go>
static int64_t with_narrow_plane(const float *charge, const uint16_t *pos_lo, const uint8_t *pos_hi, const uint8_t *counts4,
                                 const uint64_t *escapes, uint64_t n_escapes, int32_t plane, const Go &go)
{
    // this plane's escaped counts, sorted by particle: a nibble of 15 is looked up here
    std::vector<std::pair<uint64_t, unsigned>> esc;
    for (uint64_t i = 0; i < n_escapes; ++i)
        if ((int32_t)((escapes[i] >> 40) & 0xffu) == plane)
            esc.emplace_back(escapes[i] & 0xffffffffffull, (unsigned)(escapes[i] >> 48));
    std::sort(esc.begin(), esc.end());
    const auto  *eb = esc.data();
    const size_t ne = esc.size();
    return go(
        [=](uint64_t i, uint16_t &col, uint16_t &row, uint32_t &q) {
            const uint32_t pos = (uint32_t)pos_lo[i] | ((uint32_t)pos_hi[i] << 16);   // column | row << 12
            col = (uint16_t)(pos & 0xfffu);
            row = (uint16_t)(pos >> 12);
            memcpy(&q, &charge[i], 4);
        },
        [=](uint64_t k) -> unsigned {
            const unsigned c = (counts4[k >> 1] >> ((k & 1) * 4)) & 15u;
            if (c < 15u) return c;
            const auto *it = std::lower_bound(eb, eb + ne, std::make_pair(k, 0u));
            return (it != eb + ne && it->first == k) ? it->second : 0xffffffffu;   // no entry: the sum check fails
        });
}

extern "C" {

int ptb_create(const PtbBeam *beam, const PtbDevice *dev, int32_t n_dev, const PtbOptions *opt, PtbHandle *out)
{
    if (!beam || !dev || !out || n_dev < 1 || n_dev > PTB_MAX_PLANES) return PTB_E_INVALID;
    PtbContext *h = new (std::nothrow) PtbContext();
    if (!h) return PTB_E_NOMEM;
    memset(&h->cfg, 0, sizeof(h->cfg));
    h->timing = 0;
    h->alias_dev = nullptr;
    h->ev_used = 0;
    h->err[0] = 0;
    *out = h;
    // the reference's stride-D bookkeeping silently assumes exactly one record per plane and particle,
    // which only holds for z[0] == 0 and distinct consecutive z; refuse anything else
    if (dev[0].z != 0.0) return fail(h, PTB_E_INVALID, "first device must sit at z_position 0");
    for (int d = 1; d < n_dev; ++d)
        if (dev[d].z == dev[d - 1].z) return fail(h, PTB_E_INVALID, "consecutive devices share a z_position");
    for (int d = 0; d < n_dev; ++d)
        if (dev[d].ncol < 1 || dev[d].nrow < 1 || dev[d].ncol > 65535 || dev[d].nrow > 65535 ||
            !(dev[d].pitch_c > 0) || !(dev[d].pitch_r > 0) || !(dev[d].thickness > 0) || !(dev[d].rad_length > 0))
            return fail(h, PTB_E_INVALID, "device geometry out of range (1..65535 pixels, positive pitch/thickness)");
    if (!(beam->particle_rate > 0)) return fail(h, PTB_E_INVALID, "particle_rate must be positive");

    ConfigK &c = h->cfg;
    c.n_planes = n_dev;
    c.pool_entries = (opt && opt->pool_entries > 0) ? opt->pool_entries : POOL_ENTRIES;
    if (c.pool_entries != POOL_ENTRIES && c.pool_entries != SMALL_POOL)
        return fail(h, PTB_E_INVALID, "pool_entries: built variants are 128 (default) and 96 (tests)");
    c.energy = beam->energy;
    c.loc_x = beam->loc_x; c.sigma_x = beam->sigma_x;
    c.loc_y = beam->loc_y; c.sigma_y = beam->sigma_y;
    c.disp_x = beam->disp_x; c.disp_y = beam->disp_y;
    c.angle_x = beam->angle_x; c.angle_y = beam->angle_y;
    c.accept_p = (opt && opt->trigger_accept > 0) ? opt->trigger_accept : 0.7;

    // Poisson gap: lam = 1/rate*1e9 ns (main/hit.py:33) and the PTRS constants
    c.lam = 1 / beam->particle_rate * 1e9;
    {
        double slam = sqrt(c.lam);
        c.p_loglam = log(c.lam);
        c.p_b = 0.931 + 2.53 * slam;
        c.p_a = -0.059 + 0.02483 * c.p_b;
        c.p_log_invalpha = log(1.1239 + 1.1328 / (c.p_b - 3.4));
        c.p_vr = 0.9277 - 3.6224 / (c.p_b - 2);
        c.p_explam = exp(-c.lam);
    }
    // charge cloud (main/device.py:186-197), same operation order as the reference
    {
        double kb = 1.38 * 1e-23, T = 300, qe = 1.602 * 1e-19, mu = 0.14;
        double Dc = mu * kb * T / qe;
        double t = 90 * 1e-9;
        double diffusion = 0.8 * sqrt(2 * Dc * t);
        double eps0 = 8.854 * 1e-12, epsr = 11.7;
        c.diff2 = diffusion * diffusion;
        c.qe = qe;
        c.c3mu = 3 * mu;
        c.t90 = t;
        c.cden = 4 * M_PI * eps0 * epsr;
        c.sigma0 = sqrt(c.diff2 + 0.0) * 1e6;
        c.rcp_3p6 = 1.0 / 3.6;
        c.rcp_cden = 1.0 / c.cden;
        for (int j = 0; j < 64; ++j) c.exp2_tab[j] = (double)exp2l((long double)j / 64.0L);   // rounded once, from 64 bits
    }
    double mean_prev = 0.0;   // mean loss in the previous plane; the reference indexes row -1 (all zero) for plane 0
    for (int d = 0; d < n_dev; ++d) {
        const PtbDevice &v = dev[d];
        PlaneK          &p = c.pl[d];
        p.z = v.z;
        p.x_pos = v.pitch_c * v.ncol / 2 + v.delta_x;
        p.x_neg = -v.pitch_c * v.ncol / 2 + v.delta_x;
        p.y_pos = v.pitch_r * v.nrow / 2 + v.delta_y;
        p.y_neg = -v.pitch_r * v.nrow / 2 + v.delta_y;
        p.eps = v.thickness / v.rad_length;
        double corr = 1 + 0.038 * log(p.eps);
        p.corr2 = corr * corr;
        p.hl_c = 13.6 * sqrt(p.eps * p.corr2);
        p.ax[0] = AxisK{v.pitch_c, v.delta_x, v.ncol / 2.0, v.pitch_c / 2};
        p.ax[1] = AxisK{v.pitch_r, v.delta_y, v.nrow / 2.0, v.pitch_r / 2};
        p.pitch_c = v.pitch_c; p.pitch_r = v.pitch_r;
        p.delta_x = v.delta_x; p.delta_y = v.delta_y;
        p.half_nc = v.ncol / 2.0; p.half_nr = v.nrow / 2.0;
        p.half_pc = v.pitch_c / 2; p.half_pr = v.pitch_r / 2;
        p.rcp_pc = 1.0 / v.pitch_c; p.rcp_pr = 1.0 / v.pitch_r;
        p.thr_ke = v.threshold_e * 1e-3;
        {
            double kb = 1.38 * 1e-23, qe = 1.602 * 1e-19, eps0 = 8.854 * 1e-12, epsr = 11.7, T = 300;
            double cap = eps0 * epsr * v.pitch_r * v.pitch_c / v.thickness * 1e-6;
            p.noise_sigma = sqrt(kb * T * cap) / qe;
        }
        p.ncol = v.ncol; p.nrow = v.nrow; p.triggered = v.triggered ? 1 : 0;
        // Landau (Moyal form): xi = 0.1535 z^2 Z rho x / (A beta^2), x in cm; lambda = (delta-0.025)/xi - beta^2
        // - ln(xi/1e-6) - 1 + e   =>   delta = loc + xi*lambda    (main/hit.py:385, 399-400)
        {
            double en = beam->energy - mean_prev;
            double g = en / 0.511;
            double beta2 = 1 - 1 / (1 + g * g);
            double xcm = v.thickness * 1e-4;
            double xi = 0.1535 * v.Z * v.rho * xcm / (v.A * beta2);
            p.lan_scale = xi;
            p.lan_loc = 0.025 + xi * (beta2 + log(xi / 1e-6) + 1 - M_E);
            double lo = (0.0 - p.lan_loc) / xi, hi = (0.5 - p.lan_loc) / xi;
            p.lan_ulo = moyal_cdf(lo);
            p.lan_uspan = moyal_cdf(hi) - p.lan_ulo;
            if (!(p.lan_uspan > 0)) return fail(h, PTB_E_INVALID, "Landau window [0,0.5] MeV carries no probability");
            p.lan_direct = (p.lan_uspan > 0.999) ? 1 : 0;
            // mean of the truncated law by Simpson's rule, feeds beta of the next plane (main/hit.py:91)
            const int M = 20000;
            double    num = 0, den = 0, hstep = 0.5 / M;
            for (int i = 0; i <= M; ++i) {
                double wgt = (i == 0 || i == M) ? 1 : ((i & 1) ? 4 : 2);
                double dl = i * hstep;
                double f = moyal_density(dl, p.lan_loc, xi);
                num += wgt * f * dl;
                den += wgt * f;
            }
            mean_prev = den > 0 ? num / den : 0.0;
        }
    }
    c.theta0_first = highland_theta0(beam->energy, c.pl[0].eps, c.pl[0].corr2);
    // Can a production-mode cluster outgrow the profile pool?  The widest cloud belongs to the largest deposit (0.5 MeV,
    // main/hit.py:99-100), the 24-bit normals stop at sqrt(2 ln 2^33) = 6.8, and a radius r spans 2r/pitch + 2 pixels.
    {
        const double q = c.qe * (0.5 * 1e6) / 3.6;
        const double coul = 0.2 * cbrt(c.c3mu * q * c.t90 / c.cden);
        const double r_max = 6.8 * sqrt(c.diff2 + coul * coul) * 1e6 * 1.01;
        h->may_be_big = 0;
        for (int d = 0; d < n_dev; ++d)
            if (2 * r_max / dev[d].pitch_c + 2 * r_max / dev[d].pitch_r + 6 > c.pool_entries) h->may_be_big = 1;
    }

    cudaError_t e = cudaGetDevice(&h->device);
    if (e != cudaSuccess) return fail_cuda(h, e, "cudaGetDevice");
    // Poisson gap as a Walker alias table over mean +- 10 sigma (tail mass < 1e-22), when that is a sane size
    {
        const double sd = sqrt(c.lam);
        const double klo = floor(fmax(0.0, c.lam - 10.0 * sd - 10.0)), khi = ceil(c.lam + 10.0 * sd + 10.0);
        const double span = khi - klo + 1.0;
        if (c.lam > 0.0 && span <= (double)(1 << 20)) {
            const int           m = (int)span;
            std::vector<double> pmf(m);
            double              sum = 0;
            for (int i = 0; i < m; ++i) {
                const double k = klo + i;
                pmf[i] = exp(k * log(c.lam) - c.lam - lgamma(k + 1.0));
                sum += pmf[i];
            }
            std::vector<AliasEntry> tab(m);
            std::vector<int>        small, large;
            std::vector<double>     scaled(m);
            for (int i = 0; i < m; ++i) {
                scaled[i] = pmf[i] / sum * m;
                (scaled[i] < 1.0 ? small : large).push_back(i);
            }
            while (!small.empty() && !large.empty()) {
                const int s_ = small.back(), l_ = large.back();
                small.pop_back();
                tab[s_].prob = scaled[s_];
                tab[s_].alias = l_;
                tab[s_]._pad = 0;
                scaled[l_] = (scaled[l_] + scaled[s_]) - 1.0;
                if (scaled[l_] < 1.0) {
                    large.pop_back();
                    small.push_back(l_);
                }
            }
            for (int i : large) tab[i] = AliasEntry{1.0, i, 0};
            for (int i : small) tab[i] = AliasEntry{1.0, i, 0};
            e = cudaMalloc(&h->alias_dev, sizeof(AliasEntry) * m);
            if (e != cudaSuccess) return fail_cuda(h, e, "cudaMalloc(alias table)");
            e = cudaMemcpy(h->alias_dev, tab.data(), sizeof(AliasEntry) * m, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) return fail_cuda(h, e, "cudaMemcpy(alias table)");
            c.alias = (const AliasEntry *)h->alias_dev;
            c.alias_n = m;
            c.alias_kmin = (long long)klo;
        }
    }
    e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device);
    if (e != cudaSuccess) return fail_cuda(h, e, "cudaDeviceGetAttribute");
    e = prefer_shared_all();
    if (e != cudaSuccess) return fail_cuda(h, e, "cudaFuncSetAttribute(k_digitise, shared-memory carve-out)");
    return PTB_OK;
}

void ptb_destroy(PtbHandle h)
{
    if (!h) return;
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    if (h->alias_dev) cudaFree(h->alias_dev);
    delete h;
}

int ptb_timing_enable(PtbHandle h, int32_t on)
{
    if (!h) return PTB_E_INVALID;
    h->timing = on ? 1 : 0;
    return PTB_OK;
}

constexpr size_t TIMING_EVENTS = 5;

int ptb_timing_read_stages(PtbHandle h, double *stage_ms, uint64_t *calls)
{
    if (!h || !stage_ms || !calls) return PTB_E_INVALID;
    for (size_t j = 0; j + 1 < TIMING_EVENTS; ++j) stage_ms[j] = 0;
    for (size_t i = 0; i + TIMING_EVENTS <= h->ev_used; i += TIMING_EVENTS) {
        cudaError_t e = cudaEventSynchronize(h->ev[i + TIMING_EVENTS - 1]);
        if (e != cudaSuccess) return fail_cuda(h, e, "cudaEventSynchronize");
        for (size_t j = 0; j + 1 < TIMING_EVENTS; ++j) {
            float a = 0;
            e = cudaEventElapsedTime(&a, h->ev[i + j], h->ev[i + j + 1]);
            if (e != cudaSuccess) return fail_cuda(h, e, "cudaEventElapsedTime");
            stage_ms[j] += a;
        }
    }
    *calls = h->ev_used / TIMING_EVENTS;
    h->ev_used = 0;
    return PTB_OK;
}

int ptb_timing_read(PtbHandle h, double *tracks_ms, double *digitise_ms, uint64_t *calls)
{
    if (!tracks_ms || !digitise_ms) return PTB_E_INVALID;
    double    st[TIMING_EVENTS - 1];
    const int rc = ptb_timing_read_stages(h, st, calls);
    if (rc != PTB_OK) return rc;
    *tracks_ms = st[0];
    *digitise_ms = st[1];
    return PTB_OK;
}

// next group of events of the timing record, or nullptr when timing is off
static cudaEvent_t *timing_pair(PtbHandle h)
{
    if (!h->timing) return nullptr;
    if (h->ev_used + TIMING_EVENTS > h->ev.size()) {
        for (size_t i = 0; i < TIMING_EVENTS; ++i) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
            h->ev.push_back(e);
        }
    }
    cudaEvent_t *p = &h->ev[h->ev_used];
    h->ev_used += TIMING_EVENTS;
    return p;
}

const char *ptb_last_error(PtbHandle h) { return h ? h->err : "null handle"; }

int32_t ptb_n_planes(PtbHandle h) { return h ? h->cfg.n_planes : 0; }

// rows of a plane's scratch slab: the caller's figure, or the default rule of include/ptb.h
static unsigned long long scratch_rows_of(uint64_t capacity, uint64_t given)
{
    return given ? given : PTB_DEFAULT_SCRATCH_ROWS(capacity);
}

size_t ptb_workspace_bytes(PtbHandle h, uint64_t n, const uint64_t *capacity, const uint64_t *scratch_rows,
                           uint32_t flags)
{
    if (!h) return 0;
    unsigned long long rows[PTB_MAX_PLANES] = {0};
    for (int d = 0; d < h->cfg.n_planes; ++d)
        rows[d] = scratch_rows_of(capacity ? capacity[d] : 0, scratch_rows ? scratch_rows[d] : 0);
    return carve_workspace(h->cfg.n_planes, n, rows, flags, nullptr, nullptr);
}

uint64_t ptb_launch_count(void) { return g_launches; }

static bool injected_complete(const PtbInjected *v, uint32_t flags)
{
    if (flags & PTB_TRACKS_FROM_TABLE) return v->dig_z && v->dig_off;   /* accepted: table's, else v->accepted */
    bool ok = v->zstart && v->gap && v->zkick && v->eloss && v->accepted;
    if (flags & PTB_DIGITISE) ok = ok && v->dig_z && v->dig_off;
    return ok;
}

int ptb_prefix(PtbHandle h, uint64_t first, uint64_t n, uint64_t seed, const PtbInjected *inj, int64_t *out_dev,
               void *stream)
{
    if (!h || !out_dev) return PTB_E_INVALID;
    if (n == 0) return PTB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int          blocks = (int)std::min<unsigned long long>((n + 255) / 256, (unsigned long long)h->sm_count * 8);
    if (inj) {
        if (!inj->accepted || !inj->gap) return fail(h, PTB_E_INVALID, "ptb_prefix: injected accepted/gap missing");
        InjectedSource src{*inj, first, n};
        k_prefix<InjectedSource><<<blocks, 256, 0, st>>>(h->cfg, first, n, src, (long long *)out_dev);
    } else {
        PhiloxSource src{philox_keys(seed)};
        k_prefix<PhiloxSource><<<blocks, 256, 0, st>>>(h->cfg, first, n, src, (long long *)out_dev);
    }
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PTB_OK : fail_cuda(h, e, "k_prefix launch");
}

int ptb_simulate(PtbHandle h, const PtbRun *run, void *stream)
{
    if (!h || !run || !run->totals) return PTB_E_INVALID;
    const ConfigK &c = h->cfg;
    const int      D = c.n_planes;
    cudaStream_t   st = (cudaStream_t)stream;
    const uint32_t flags = run->flags;
    if (run->n == 0) return fail(h, PTB_E_INVALID, "ptb_simulate: empty particle range");
    if ((flags & PTB_TRACKS_FROM_TABLE) && !(flags & PTB_DIGITISE))
        return fail(h, PTB_E_INVALID, "PTB_TRACKS_FROM_TABLE without PTB_DIGITISE does nothing");
    if ((flags & PTB_TRACKS_FROM_TABLE) && (flags & PTB_WRITE_TRACKS))
        return fail(h, PTB_E_INVALID, "PTB_TRACKS_FROM_TABLE and PTB_WRITE_TRACKS are exclusive");
    if (flags & (PTB_TRACKS_FROM_TABLE | PTB_WRITE_TRACKS)) {
        const PtbTrackTable &t = run->tracks;
        bool ok = t.x && t.y && t.eloss && (t.accepted || (flags & PTB_TRACKS_FROM_TABLE));
        if (flags & PTB_WRITE_TRACKS) ok = ok && t.energy && t.angle_x && t.angle_y && t.z && t.time_stamp;
        if (!ok) return fail(h, PTB_E_INVALID, "track table pointers missing");
    }
    if (run->injected && !injected_complete(run->injected, flags))
        return fail(h, PTB_E_INVALID, "injected variate tables incomplete for the requested flags");
    if (run->injected && (flags & PTB_TRACKS_FROM_TABLE) && !run->tracks.accepted && !run->injected->accepted)
        return fail(h, PTB_E_INVALID, "no trigger mask: neither the track table nor the injected tables carry one");
    unsigned long long cap[PTB_MAX_PLANES] = {0};   // rows of every plane's scratch slab
    const bool         compact = (flags & PTB_COMPACT) != 0;
    if (compact && (!(flags & PTB_DIGITISE) || (flags & PTB_KEEP_CHARGE64)))
        return fail(h, PTB_E_INVALID, "PTB_COMPACT needs PTB_DIGITISE and excludes PTB_KEEP_CHARGE64");
    const bool         narrow = (flags & PTB_COMPACT_NARROW) != 0;
    if (narrow && !compact) return fail(h, PTB_E_INVALID, "PTB_COMPACT_NARROW needs PTB_COMPACT");
    if (compact && (!run->compact.counts || !run->compact.accepted ||
                    (narrow ? (!run->compact.charge || !run->compact.pos_lo || !run->compact.pos_hi ||
                               (run->compact.escape_capacity && !run->compact.escapes))
                            : (!(flags & PTB_COMPACT_PACKED) && !run->compact.hits))))
        return fail(h, PTB_E_INVALID, "compact output pointers missing");
    const bool packed = (flags & PTB_COMPACT_PACKED) != 0;
    PtbPackedLayout layout;
    memset(&layout, 0, sizeof(layout));
    if (packed) {
        if (!compact || narrow) return fail(h, PTB_E_INVALID, "PTB_COMPACT_PACKED needs PTB_COMPACT and excludes PTB_COMPACT_NARROW");
        if (!run->compact.word0 || !run->compact.word1 || !run->compact.word2 ||
            (run->compact.escape_capacity && !run->compact.escapes))
            return fail(h, PTB_E_INVALID, "compact output pointers missing");
        if (ptb_packed_layout(h, &layout) != PTB_OK) return PTB_E_INVALID;   // (the message is set there)
        if (run->compact.plane_first[D] > 0x7fffffffull)
            return fail(h, PTB_E_INVALID, "PTB_COMPACT_PACKED: at most 2^31 - 1 rows per call");
    }
    if (narrow)
        for (int d = 0; d < D; ++d)
            if (c.pl[d].ncol > 4095 || c.pl[d].nrow > 4095)
                return fail(h, PTB_E_INVALID, "PTB_COMPACT_NARROW holds 12 bits per pixel index: a matrix exceeds 4095 columns or rows");
    if (flags & PTB_DIGITISE) {
        for (int d = 0; d < D; ++d) {
            const PtbHitSlab &s = run->slabs[d];
            if (!compact && (!s.event || !s.col || !s.row || !s.charge || ((flags & PTB_KEEP_CHARGE64) && !s.charge64)))
                return fail(h, PTB_E_INVALID, "hit slab pointers missing");
            if (compact && run->compact.plane_first[d + 1] < run->compact.plane_first[d])
                return fail(h, PTB_E_INVALID, "compact.plane_first must not decrease");
            cap[d] = scratch_rows_of(compact ? run->compact.plane_first[d + 1] - run->compact.plane_first[d] : s.capacity,
                                     run->scratch_rows[d]);
        }
    }
    KParams P;
    memset(&P, 0, sizeof(P));
    P.first = run->first;
    P.n = run->n;
    P.n_groups = (run->n + 31) / 32;
    P.flags = flags;
    P.trk = run->tracks;
    P.totals = run->totals;
    P.cmp = run->compact;
    size_t need = carve_workspace(D, run->n, cap, flags, (char *)run->workspace, &P.ws);
    if (!run->workspace || run->workspace_bytes < need) return fail(h, PTB_E_INVALID, "workspace too small");
    if (P.n_groups > 0x7fffffffull * WARPS_PER_BLOCK) return fail(h, PTB_E_INVALID, "range too large for one call");

    cudaError_t e;
    e = cudaMemsetAsync(P.ws.alloc, 0, sizeof(unsigned long long) * ALLOC_LINES * 16, st);
    if (e != cudaSuccess) return fail_cuda(h, e, "memset alloc");
    // the error word and the escape counter collect over the call's kernels: every call starts from zero
    static_assert(offsetof(PtbTotals, escapes) == offsetof(PtbTotals, error) + 4, "error and escapes are cleared together");
    e = cudaMemsetAsync(&run->totals->error, 0, 8, st);
    if (e != cudaSuccess) return fail_cuda(h, e, "memset totals");

    const bool keep64 = (flags & PTB_KEEP_CHARGE64) != 0;
    cudaEvent_t *tp = timing_pair(h);
    if (tp) cudaEventRecord(tp[0], st);
    if (!(flags & PTB_TRACKS_FROM_TABLE)) {
        const unsigned tb = (unsigned)((P.n_groups + TRACK_BLOCK / 32 - 1) / (TRACK_BLOCK / 32));
        if (run->injected)
            k_tracks<InjectedSource><<<tb, TRACK_BLOCK, 0, st>>>(c, P, InjectedSource{*run->injected, run->first, run->n});
        else
            k_tracks<PhiloxSource><<<tb, TRACK_BLOCK, 0, st>>>(c, P, PhiloxSource{philox_keys(run->seed)});
        ++g_launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return fail_cuda(h, e, "k_tracks launch");
    }
    if (tp) cudaEventRecord(tp[1], st);
    if (flags & PTB_DIGITISE) {
        const unsigned blocks = (unsigned)std::min<unsigned long long>((P.n_groups + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK,
                                                                       (unsigned long long)h->sm_count * DIGI_MIN_BLOCKS);   // one wave
        const bool     small = c.pool_entries == SMALL_POOL;
        const int mode = compact ? 2 : keep64 ? 1 : 0;
        if (run->injected) {
            InjectedSource src{*run->injected, run->first, run->n};
            e = small ? launch_digitise_mode<InjectedSource, SMALL_POOL>(mode, blocks, st, c, P, src)
                      : launch_digitise_mode<InjectedSource, POOL_ENTRIES>(mode, blocks, st, c, P, src);
        } else {
            PhiloxSource src{philox_keys(run->seed)};
            e = small ? launch_digitise_mode<PhiloxSource, SMALL_POOL>(mode, blocks, st, c, P, src)
                      : launch_digitise_mode<PhiloxSource, POOL_ENTRIES>(mode, blocks, st, c, P, src);
        }
        ++g_launches;
        if (e != cudaSuccess) return fail_cuda(h, e, "k_digitise launch");
        if (tp) cudaEventRecord(tp[2], st);
        // clusters wider than the profile pool (possible with the caller's own variates / deposits, or when the geometry
        // says so, ptb_create): always launched -- with an empty list the blocks read one counter and leave -- so that
        // a pair k_digitise defers can never stay unserved; a small grid when no such cluster is expected
        {
            const bool               expected = run->injected || (flags & PTB_TRACKS_FROM_TABLE) || h->may_be_big;
#ifdef PTB_BIG_ONLY_IF_EXPECTED
            if (expected)
#endif
            {
            const unsigned long long items = P.n_groups * (unsigned long long)D;
            const unsigned bb = (unsigned)std::min<unsigned long long>((items + BIG_WARPS - 1) / BIG_WARPS,
                                                                       (unsigned long long)h->sm_count * (expected ? 16 : 2));
            if (run->injected) {
                InjectedSource src{*run->injected, run->first, run->n};
                if (keep64) k_digitise_big<InjectedSource, true><<<bb, 32 * BIG_WARPS, 0, st>>>(c, P, src);
                else        k_digitise_big<InjectedSource, false><<<bb, 32 * BIG_WARPS, 0, st>>>(c, P, src);
            } else {
                PhiloxSource src{philox_keys(run->seed)};
                if (keep64) k_digitise_big<PhiloxSource, true><<<bb, 32 * BIG_WARPS, 0, st>>>(c, P, src);
                else        k_digitise_big<PhiloxSource, false><<<bb, 32 * BIG_WARPS, 0, st>>>(c, P, src);
            }
            ++g_launches;
            e = cudaGetLastError();
            if (e != cudaSuccess) return fail_cuda(h, e, "k_digitise_big launch");
            }
        }
    }
    if (tp && !(flags & PTB_DIGITISE)) cudaEventRecord(tp[2], st);

    k_scan_sums<<<(D + 5) * SCAN_SEGMENTS, SCAN_THREADS, 0, st>>>(P.n_groups, P.ws);
    k_scan<<<(D + 5) * SCAN_SEGMENTS, SCAN_THREADS, 0, st>>>(D, P.n_groups, flags, P.ws, run->bases_in, run->bases_out, run->totals);
    g_launches += 2;
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(h, e, "k_scan launch");
    if (tp) cudaEventRecord(tp[3], st);

    if ((flags & PTB_DIGITISE) || (flags & PTB_WRITE_TRACKS)) {
        GatherParams G;
        memset(&G, 0, sizeof(G));
        G.D = D;
        G.flags = flags;
        G.n = run->n;
        G.n_groups = P.n_groups;
        G.ws = P.ws;
        for (int d = 0; d < D; ++d) G.slabs[d] = run->slabs[d];
        G.cmp = run->compact;
        G.packed = layout;
        G.time_stamp = run->tracks.time_stamp;
        G.totals = run->totals;
        k_gather<<<(unsigned)((P.n_groups + 7) / 8), 256, 0, st>>>(G);
        ++g_launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return fail_cuda(h, e, "k_gather launch");
    }
    if (tp) cudaEventRecord(tp[4], st);
    return PTB_OK;
}

int ptb_pack_hits(const PtbHitSlab *slab, uint64_t n, void *out_dev, void *stream)
{
    if (!slab || !out_dev) return PTB_E_INVALID;
    if (n == 0) return PTB_OK;
    if (n > slab->capacity) return PTB_E_INVALID;
    k_pack<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*slab, n, (uint16_t *)out_dev);
    ++g_launches;
    return cudaGetLastError() == cudaSuccess ? PTB_OK : PTB_E_CUDA;
}

int ptb_correlate(const int64_t *ev1, const uint16_t *col1, const uint16_t *row1, uint64_t n1, const int64_t *ev2,
                  const uint16_t *col2, const uint16_t *row2, uint64_t n2, int64_t ev_min, int64_t ev_max,
                  int32_t *x_hist, int32_t x_dim0, int32_t x_dim1, int32_t *y_hist, int32_t y_dim0, int32_t y_dim1,
                  void *stream)
{
    if (!ev1 || !col1 || !row1 || !ev2 || !col2 || !row2 || !x_hist || !y_hist) return PTB_E_INVALID;
    if (x_dim0 < 1 || x_dim1 < 1 || y_dim0 < 1 || y_dim1 < 1) return PTB_E_INVALID;
    if (n1 == 0 || n2 == 0 || ev_max <= ev_min) return PTB_OK;
    const unsigned long long events = (unsigned long long)(ev_max - ev_min);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned           blocks = (unsigned)std::min<unsigned long long>((events + 255) / 256, (unsigned long long)sms * 16);
    k_correlate<<<blocks, 256, 0, (cudaStream_t)stream>>>(ev1, col1, row1, n1, ev2, col2, row2, n2, ev_min, ev_max,
                                                         x_hist, x_dim0, x_dim1, y_hist, y_dim0, y_dim1);
    ++g_launches;
    return cudaGetLastError() == cudaSuccess ? PTB_OK : PTB_E_CUDA;
}

int ptb_digest(const PtbHitSlab *slab, uint64_t n, uint64_t row_base, uint64_t *out_dev, void *stream)
{
    if (!slab || !out_dev) return PTB_E_INVALID;
    if (n == 0) return PTB_OK;
    if (n > slab->capacity || !slab->event || !slab->col || !slab->row || !slab->charge) return PTB_E_INVALID;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned blocks = (unsigned)std::min<unsigned long long>((n + 255) / 256, (unsigned long long)sms * 8);
    k_digest<<<blocks, 256, 0, (cudaStream_t)stream>>>(*slab, n, row_base, (unsigned long long *)out_dev);
    ++g_launches;
    return cudaGetLastError() == cudaSuccess ? PTB_OK : PTB_E_CUDA;
}

// Host side of PtbCompactOut: counts + trigger mask -> event numbers, 8-byte rows -> the reference's 18-byte records.
// The particle range is cut into blocks of whole 32-particle words; a first pass sums every block's rows and accepted
// particles, the second pass lets each thread expand its blocks from those prefixes.
int ptb_packed_layout(PtbHandle h, PtbPackedLayout *out)
{
    if (!h || !out) return PTB_E_INVALID;
    PtbPackedLayout L;
    memset(&L, 0, sizeof(L));
    auto bits_for = [](int32_t n) {
        int b = 0;
        while ((1ll << b) <= (long long)n) ++b;   // indices run from 1 to n
        return b;
    };
    for (int d = 0; d < h->cfg.n_planes; ++d) {
        const PlaneK &pl = h->cfg.pl[d];
        const int     cb = bits_for(pl.ncol), rb = bits_for(pl.nrow);
        const float   thr = (float)pl.thr_ke;
        uint32_t      tb;
        memcpy(&tb, &thr, 4);
        const unsigned ex = (tb >> 23) & 0xffu;
        // a row exists because its charge reached the threshold (as doubles; rounding to f32 keeps the order): with a
        // positive, normal threshold every charge is positive and its exponent is at least the threshold's
        if (cb + rb > 21 || !(pl.thr_ke > 0.0) || ex == 0 || ex > 239)
            return fail(h, PTB_E_INVALID, "PTB_COMPACT_PACKED: a plane needs more than 21 bits of column + row index, or its "
                                          "threshold is not a positive number");
        L.exp_min[d] = (uint8_t)ex;
        L.col_bits[d] = (uint8_t)cb;
    }
    *out = L;
    return PTB_OK;
}

int64_t ptb_expand_hits(const uint64_t *hits, const uint8_t *counts, const uint32_t *accepted, uint64_t n,
                        int64_t event_base, uint64_t n_rows_expected, void *out, int32_t threads)
{
    if ((!hits && n_rows_expected) || !counts || !accepted || (!out && n_rows_expected)) return PTB_E_INVALID;
    return with_wide_plane(hits, counts, [&](const auto &rows_of, const auto &count_of) {
        return expand_rows(rows_of, count_of, accepted, n, event_base, n_rows_expected,
                           [=] { return RowsToMemory{(unsigned char *)out}; }, threads);
    });
}

int64_t ptb_expand_hits_narrow(const float *charge, const uint16_t *pos_lo, const uint8_t *pos_hi, const uint8_t *counts4,
                               const uint64_t *escapes, uint64_t n_escapes, int32_t plane, const uint32_t *accepted,
                               uint64_t n, int64_t event_base, uint64_t n_rows_expected, void *out, int32_t threads)
{
    if (((!charge || !pos_lo || !pos_hi) && n_rows_expected) || !counts4 || !accepted || (!out && n_rows_expected) ||
        (n_escapes && !escapes))
        return PTB_E_INVALID;
    return with_narrow_plane(charge, pos_lo, pos_hi, counts4, escapes, n_escapes, plane,
                             [&](const auto &rows_of, const auto &count_of) {
                                 return expand_rows(rows_of, count_of, accepted, n, event_base, n_rows_expected,
                                                    [=] { return RowsToMemory{(unsigned char *)out}; }, threads);
                             });
}

int64_t ptb_expand_hits_packed(const uint16_t *word0, const uint16_t *word1, const uint16_t *word2, const uint8_t *counts4,
                               const uint64_t *escapes, uint64_t n_escapes, int32_t plane, const PtbPackedLayout *layout,
                               uint64_t row_base, const uint32_t *accepted, uint64_t n, int64_t event_base,
                               uint64_t n_rows_expected, void *out, int32_t threads)
{
    if (((!word0 || !word1 || !word2) && n_rows_expected) || !counts4 || !accepted || (!out && n_rows_expected) ||
        (n_escapes && !escapes) || !layout || plane < 0 || plane >= PTB_MAX_PLANES)
        return PTB_E_INVALID;
    return with_packed_plane(word0, word1, word2, counts4, escapes, n_escapes, plane, *layout, row_base, n_rows_expected,
                             [&](const auto &rows_of, const auto &count_of) {
                                 return expand_rows(rows_of, count_of, accepted, n, event_base, n_rows_expected,
                                                    [=] { return RowsToMemory{(unsigned char *)out}; }, threads);
                             });
}

int64_t ptb_expand_hits_packed_fd(const uint16_t *word0, const uint16_t *word1, const uint16_t *word2,
                                  const uint8_t *counts4, const uint64_t *escapes, uint64_t n_escapes, int32_t plane,
                                  const PtbPackedLayout *layout, uint64_t row_base, const uint32_t *accepted, uint64_t n,
                                  int64_t event_base, uint64_t n_rows_expected, int32_t fd, uint64_t file_offset,
                                  int32_t threads)
{
    if (((!word0 || !word1 || !word2) && n_rows_expected) || !counts4 || !accepted || fd < 0 || (n_escapes && !escapes) ||
        !layout || plane < 0 || plane >= PTB_MAX_PLANES)
        return PTB_E_INVALID;
    return with_packed_plane(word0, word1, word2, counts4, escapes, n_escapes, plane, *layout, row_base, n_rows_expected,
                             [&](const auto &rows_of, const auto &count_of) {
                                 return expand_rows(rows_of, count_of, accepted, n, event_base, n_rows_expected,
                                                    [=] { return RowsToFile(fd, file_offset); }, threads);
                             });
}

int64_t ptb_expand_hits_fd(const uint64_t *hits, const uint8_t *counts, const uint32_t *accepted, uint64_t n,
                           int64_t event_base, uint64_t n_rows_expected, int32_t fd, uint64_t file_offset, int32_t threads)
{
    if ((!hits && n_rows_expected) || !counts || !accepted || fd < 0) return PTB_E_INVALID;
    return with_wide_plane(hits, counts, [&](const auto &rows_of, const auto &count_of) {
        return expand_rows(rows_of, count_of, accepted, n, event_base, n_rows_expected,
                           [=] { return RowsToFile(fd, file_offset); }, threads);
    });
}

int64_t ptb_expand_hits_narrow_fd(const float *charge, const uint16_t *pos_lo, const uint8_t *pos_hi,
                                  const uint8_t *counts4, const uint64_t *escapes, uint64_t n_escapes, int32_t plane,
                                  const uint32_t *accepted, uint64_t n, int64_t event_base, uint64_t n_rows_expected,
                                  int32_t fd, uint64_t file_offset, int32_t threads)
{
    if (((!charge || !pos_lo || !pos_hi) && n_rows_expected) || !counts4 || !accepted || fd < 0 || (n_escapes && !escapes))
        return PTB_E_INVALID;
    return with_narrow_plane(charge, pos_lo, pos_hi, counts4, escapes, n_escapes, plane,
                             [&](const auto &rows_of, const auto &count_of) {
                                 return expand_rows(rows_of, count_of, accepted, n, event_base, n_rows_expected,
                                                    [=] { return RowsToFile(fd, file_offset); }, threads);
                             });
}

double ptb_measure_fp64_peak(void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    int          dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256, iters = 1 << 14;
    double   *buf = nullptr;
    if (cudaMalloc(&buf, sizeof(double) * blocks * threads) != cudaSuccess) return -1;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k_dfma<<<blocks, threads, 0, st>>>(buf, iters, 0.999999, 1e-9);
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a, st);
        k_dfma<<<blocks, threads, 0, st>>>(buf, iters, 0.999999, 1e-9);
        cudaEventRecord(b, st);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        double flops = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3);
        if (flops > best) best = flops;
    }
    g_launches += 6;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(buf);
    return cudaGetLastError() == cudaSuccess ? best : -1;
}

}  // extern "C"

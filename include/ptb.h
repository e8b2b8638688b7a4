/*
 * ptb.h -- C ABI of libptb_b200: the per-particle Monte Carlo hot path of pytestbeam as sm_100a CUDA
 * kernels.  Plain pointers and sizes only; no torch, no C++ types.  Every entry point is
 * stream-ordered and non-blocking; the caller owns all device buffers (hit slabs, track tables,
 * injected variates, workspace); the library owns only the handle (host memory plus one small device table of the
 * Poisson gap law).
 *
 * What each entry point replaces in the reference (rpartzsch/pytestbeam has no FFI layer; the boundary
 * is two Python calls and the two Numba kernels behind them -- SURVEY.md section 8b):
 *
 *   ptb_create      the flattening of setup.yml / material.yml into per-plane scalar lists:
 *                   pytestbeam/main/hit.py:24-61 and pytestbeam/main/device.py:30-53
 *   ptb_prefix      the two running counters the sequential loops carry across particles: the accepted-
 *                   event counter (main/device.py:134,164-165) and the cumulative time stamp
 *                   (main/hit.py:73,239); needed up front only when particles are sharded
 *   ptb_simulate    generate_tracks (main/hit.py:141-241, called at :107), the Landau table it consumes
 *                   (main/hit.py:84-102 with sample_dist_fast :327-368), the trigger mask
 *                   (main/device.py:56) and calc_position (main/device.py:91-168, called at :68) for
 *                   every plane, fused; with PTB_TRACKS_FROM_TABLE it is calc_position alone, fed from
 *                   the hit_tables tuple exactly as main/device.py:132-138 reads it
 *   PtbTrackTable   the 8-tuple returned by hit.tracks (main/hit.py:138)
 *   PtbHitSlab      the per-device hit table (main/device.py:20-28,66,156-163), as structure of arrays
 *   ptb_pack_hits   the structured-array rows handed to create_hit_file (main/device.py:317-327)
 *   PtbCompactOut   the same hit tables in compressed-row form for the trip to the host: per (plane, particle) the
 *                   number of rows calc_position appended for it (main/device.py:156-163) instead of a repeated event
 *                   number, the trigger mask (main/device.py:56) and 8-byte (column, row, charge) rows -- or, with
 *                   PTB_COMPACT_NARROW, 7-byte rows in three arrays and 4-bit counts
 *   ptb_expand_hits, ptb_expand_hits_narrow   that form back into the 18-byte rows of main/device.py:20-28 (host code,
 *                   multi-threaded); ptb_expand_hits_fd, ptb_expand_hits_narrow_fd write them into a file at a byte
 *                   offset instead: the /Hits dataset that create_hit_file fills (main/device.py:306-340)
 *   ptb_digest      no counterpart: order-sensitive checksums of a hit table, for comparing sharded runs
 *   PTB_ZERO_RADIUS the calc_cluster_hits call of the threshold scan (main/threshold_scan.py:55-69)
 *   ptb_correlate   _eventloop_fast of the correlation plot (main/plotting.py:730-764)
 *
 * Error convention: 0 on success, negative PTB_E_* otherwise; nothing is thrown across the ABI.
 * Conditions detected on the device (slab / scratch overflow, a cluster bounding box beyond 2^20 pixels) are reported
 * in PtbTotals.error, which the caller reads after synchronising the stream; on an overflow the totals still hold the
 * exact sizes, so the call can be repeated with them.
 * One handle per GPU; a handle is not thread-safe, distinct handles are independent.
 */
#ifndef PTB_H_
#define PTB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTB_MAX_PLANES 32

/* return codes */
#define PTB_OK 0
#define PTB_E_INVALID (-1)   /* bad argument / unsupported geometry (see ptb_last_error) */
#define PTB_E_CUDA (-2)      /* a CUDA runtime call failed */
#define PTB_E_NOMEM (-3)
#define PTB_E_IO (-4)        /* a write to the caller's file descriptor failed (ptb_expand_hits*_fd) */

/* bits of PtbTotals.error (device-detected) */
#define PTB_DEV_SCRATCH_OVERFLOW 1u /* the workspace's scratch rows of a plane ran out; PtbTotals.reserved is exact */
#define PTB_DEV_SLAB_OVERFLOW 2u  /* a hit slab's capacity was exceeded; totals still hold the exact counts */
#define PTB_DEV_BOX_TOO_LARGE 4u  /* a cluster bounding box exceeded 2^20 pixels */
#define PTB_DEV_PACK_OVERFLOW 16u /* PTB_COMPACT_PACKED: a negative charge or one below the plane's threshold reached the
                                     row packer (an internal inconsistency: a row exists because it reached the threshold) */
#define PTB_DEV_COUNT_OVERFLOW 8u /* PTB_COMPACT: one particle left more than 255 rows on a plane (the count saturated),
                                     or PTB_COMPACT_NARROW ran out of escape entries; rerun the range without
                                     PTB_COMPACT */

/* flags of PtbRun.flags */
#define PTB_DIGITISE 1u          /* run the cluster digitisation and fill the hit slabs */
#define PTB_TRACKS_FROM_TABLE 2u /* do not generate tracks: read x, y, eloss from PtbRun.tracks (calc_position alone) */
#define PTB_WRITE_TRACKS 4u      /* store the per-plane track records into PtbRun.tracks */
#define PTB_KEEP_CHARGE64 8u     /* also store the f8 charge of every hit (PtbHitSlab.charge64) */
#define PTB_ZERO_RADIUS 32u      /* charge injection (threshold scan, threshold_scan.py:55-69): the cluster radii are 0
                                    instead of being drawn, so only the seed-pixel fallback of calc_cluster_hits fires */
#define PTB_RECYCLE_SLABS 16u    /* rows start at 0 of each slab: ignore bases_in->hits and write bases_out->hits = 0 */
#define PTB_COMPACT 64u          /* with PTB_DIGITISE: write PtbRun.compact instead of PtbRun.slabs (always recycled) */
#define PTB_COMPACT_NARROW 128u  /* with PTB_COMPACT: 7-byte rows in three arrays and 4-bit counts (PtbCompactOut) */
#define PTB_COMPACT_PACKED 256u  /* with PTB_COMPACT: 6-byte rows in three arrays and 4-bit counts (PtbCompactOut); for
                                    set-ups ptb_packed_layout() accepts */

/* beam block of setup.yml; lengths um, angles rad, energy MeV, rate Hz */
typedef struct PtbBeam {
    double energy, sigma_x, sigma_y, loc_x, loc_y, disp_x, disp_y, angle_x, angle_y, particle_rate;
} PtbBeam;

/* one device of setup.yml with its material.yml entry, in beam order */
typedef struct PtbDevice {
    int32_t ncol, nrow;
    double  pitch_c, pitch_r, thickness, z, delta_x, delta_y;
    double  threshold_e;  /* electrons, as written in setup.yml */
    int32_t triggered;    /* trigger: "triggered" -> 1 */
    int32_t _pad;
    double  rad_length, Z, A, rho;
} PtbDevice;

typedef struct PtbOptions {
    int32_t pool_entries; /* box columns + rows one round of a 32-particle group may hold in shared memory; 0 = default
                             (128).  Compile-time variants: 128 and 96 (drives the multi-round path in the tests).
                             No limit on the clusters: a single cluster beyond it is served without the pool */
    int32_t _pad;
    double  trigger_accept; /* main/device.py:56 hard-codes 0.7; 0 = use that */
} PtbOptions;

/*
 * Deterministic mode: every random variate of particles [first, first+n) pre-drawn by the caller
 * (device pointers, indexed by k - first).
 */
typedef struct PtbInjected {
    const double  *zstart;   /* [n][4]  standard normals for (angle_x, angle_y, x, y) at the source */
    const int64_t *gap;      /* [n]     Poisson gap to the previous particle, ns */
    const double  *zkick;    /* [n][D-1][2] standard normals of the scattering kicks at planes 1..D-1 */
    const double  *eloss;    /* [D][n]  Landau energy loss per plane before the acceptance zeroing */
    const uint8_t *accepted; /* [n]     trigger mask */
    const double  *dig_z;    /* digitisation normals, concatenated */
    const int64_t *dig_off;  /* [D][n+1] offsets into dig_z: (radius_x, radius_y, seed noise, pixel noise ...) */
    double         eloss0_prev; /* eloss[0][first-1]; ignored when first == 0 */
} PtbInjected;

/* the reference's hit_tables tuple (main/hit.py:138) for particles [first, first+n), particle-major */
typedef struct PtbTrackTable {
    double  *energy, *angle_x, *angle_y, *x, *y, *z; /* [n*D] */
    int64_t *time_stamp;                              /* [n*D] */
    double  *eloss;                                   /* [D][n] after the acceptance zeroing */
    uint8_t *accepted;                                /* [n] trigger mask; with PTB_TRACKS_FROM_TABLE an optional input
                                                         (NULL = draw it from the variate source, main/device.py:56) */
} PtbTrackTable;

/* hit table of one plane as structure of arrays; rows in particle order, then column-major in the cluster */
typedef struct PtbHitSlab {
    int64_t  *event;    /* event_number */
    uint16_t *col;      /* column, 1-based */
    uint16_t *row;      /* row, 1-based */
    float    *charge;   /* ke */
    double   *charge64; /* optional (PTB_KEEP_CHARGE64) */
    uint64_t  capacity; /* rows */
} PtbHitSlab;

/*
 * Compressed-row form of the hit tables of particles [first, first+n) (PTB_COMPACT), device memory owned by the caller.
 * Plane d's totals.hits[d] rows sit at hits[plane_first[d] ...], in the reference's order (particle, column, row); the
 * plane's region ends at plane_first[d+1].  Row i of plane d belongs to the particle k whose running count range covers
 * i:  sum(counts[d][0..k-1]) <= i < ... + counts[d][k];  its event number is  bases_in->event + 1 + (accepted particles
 * among first .. first+k-1)  (main/device.py:159,164-165), frame is 0.  66 instead of 118 bytes per electron on the
 * default telescope.
 *
 * PTB_COMPACT_NARROW (matrices of at most 4095 x 4095 pixels): the rows of all planes lie back to back -- plane d's
 * rows start at row  totals.hits[0] + ... + totals.hits[d-1]  of the three row arrays, whose common capacity is
 * plane_first[n_planes] -- as  charge[i]  (f32),  pos_lo[i] | pos_hi[i] << 16 = column | row << 12;  counts holds 4 bits
 * per (plane, particle): the count of particle k on plane d is nibble (k & 1) of counts[d * ((n + 1) / 2) + k / 2].
 * A nibble of 15 means "15 or more": the exact count (up to 255) then stands in one of the totals.escapes entries of
 * the escapes array, in no particular order, as  count << 48 | plane << 40 | (k - first).
 * 55 bytes per electron on the default telescope.
 *
 * PTB_COMPACT_PACKED (set-ups ptb_packed_layout accepts: every plane's column and row indices fit 21 bits together and
 * its threshold is positive): counts, escapes, trigger mask and row order as in the narrow form; a row is 48 bits,
 * word0[i] | word1[i] << 16 | word2[i] << 32:
 *     bits 0-22   mantissa of the f32 charge (its sign is +: a row exists because the charge reached the threshold)
 *     bits 23-26  biased exponent of the charge minus exp_min[plane] (the exponent of the threshold as f32); a charge
 *                 beyond these 16 binades (the far tail of the deposit law) keeps 15 here and stands on the escape list
 *                 as  1 << 63 | row index (counted over all planes, < 2^31) << 32 | f32 bits of the charge
 *     bits 27 ..  column (col_bits[plane] bits), then row
 * 48 bytes per electron on the default telescope: this is what travels over PCIe.
 */
typedef struct PtbCompactOut {
    uint64_t *hits;     /* column | row << 16 | (f32 charge bits) << 32; unused with PTB_COMPACT_NARROW / _PACKED */
    uint8_t  *counts;   /* [D][n] rows of particle k on plane d; [D][(n+1)/2] nibble pairs with PTB_COMPACT_NARROW / _PACKED */
    uint32_t *accepted; /* [(n+31)/32] trigger mask, bit j of word g = particle first + 32 g + j */
    uint64_t  plane_first[PTB_MAX_PLANES + 1]; /* first row of every plane's region in hits; [D] = rows of hits in all */
    float    *charge;   /* PTB_COMPACT_NARROW: the three row arrays, plane_first[D] rows each */
    uint16_t *pos_lo;
    uint8_t  *pos_hi;
    uint64_t *escapes;  /* PTB_COMPACT_NARROW / _PACKED: counts of 15 and more, out-of-window charges (see above);
                           escape_capacity entries */
    uint64_t  escape_capacity;
    uint16_t *word0, *word1, *word2; /* PTB_COMPACT_PACKED: the three row arrays, plane_first[D] rows each */
} PtbCompactOut;

/* bit layout of the PTB_COMPACT_PACKED rows of every plane (see above) */
typedef struct PtbPackedLayout {
    uint8_t exp_min[PTB_MAX_PLANES];  /* biased f32 exponent of the plane's threshold in ke */
    uint8_t col_bits[PTB_MAX_PLANES]; /* bits of the column index; the row index takes the bits above it */
} PtbPackedLayout;

/* running prefixes carried from one particle range to the next (device memory) */
typedef struct PtbBases {
    int64_t  event;            /* accepted particles before `first` */
    int64_t  time;             /* time stamp of particle first-1 (0 when first == 0) */
    uint64_t hits[PTB_MAX_PLANES]; /* rows already in each slab (row offset at which this range starts) */
} PtbBases;

/* per-call results (device memory) */
typedef struct PtbTotals {
    uint64_t hits[PTB_MAX_PLANES]; /* rows produced per plane by this call */
    uint64_t accepted;
    int64_t  time_sum;
    uint64_t pairs, inside, pixels; /* census: digitised (particle,plane) pairs, seed-in-matrix pairs, pixel evaluations */
    uint32_t error;
    uint32_t escapes;  /* PTB_COMPACT_NARROW: entries written to PtbCompactOut.escapes (more than its capacity: the
                          rest was dropped and PTB_DEV_COUNT_OVERFLOW is set) */
    uint64_t reserved[PTB_MAX_PLANES]; /* scratch rows the call needs per plane (>= hits: every 32-particle group
                                          reserves room for all pixels that can still become hits) */
} PtbTotals;

/* scratch rows per plane when the caller names none: four times the slab capacity (measured need on the default telescope: 2.2 x the hits on the MIMOSA26 planes, 3.9 x on ITkPix) */
#define PTB_DEFAULT_SCRATCH_ROWS(capacity) (4ull * (capacity) + 4096ull)

typedef struct PtbRun {
    uint64_t first, n;   /* global particle index range */
    uint64_t seed;       /* Philox key (production mode) */
    uint32_t flags, _pad;
    const PtbInjected *injected;   /* host pointer to the struct, or NULL for production (Philox) mode */
    PtbTrackTable      tracks;     /* used with PTB_TRACKS_FROM_TABLE / PTB_WRITE_TRACKS */
    PtbHitSlab         slabs[PTB_MAX_PLANES];
    const PtbBases    *bases_in;   /* device; NULL = all zero */
    PtbBases          *bases_out;  /* device; bases_in + this call's totals; may alias bases_in; may be NULL */
    PtbTotals         *totals;     /* device; required */
    void              *workspace;  /* device; ptb_workspace_bytes(handle, n, capacities, scratch_rows, flags) bytes */
    size_t             workspace_bytes;
    uint64_t           scratch_rows[PTB_MAX_PLANES]; /* rows of the workspace's scratch slab per plane; 0 = the default
                                                        rule.  On PTB_DEV_SCRATCH_OVERFLOW retry with totals.reserved */
    PtbCompactOut      compact;    /* used with PTB_COMPACT; scratch_rows[d] = 0 then means the default rule on the
                                      rows of plane d's region */
} PtbRun;

typedef struct PtbContext *PtbHandle;

int         ptb_create(const PtbBeam *beam, const PtbDevice *devices, int32_t n_devices, const PtbOptions *opt,
                       PtbHandle *out);
void        ptb_destroy(PtbHandle h);
const char *ptb_last_error(PtbHandle h);
int32_t     ptb_n_planes(PtbHandle h);

/* the row layout PTB_COMPACT_PACKED uses for this handle's planes; PTB_E_INVALID (and *out untouched) when a plane does
 * not fit it: more than 21 bits of column + row index, or a threshold that is not positive */
int ptb_packed_layout(PtbHandle h, PtbPackedLayout *out);

/* bytes of workspace ptb_simulate needs for n particles, the given slab capacities and scratch rows (NULL = default) */
size_t ptb_workspace_bytes(PtbHandle h, uint64_t n, const uint64_t *capacity /*[n_planes]*/,
                           const uint64_t *scratch_rows /*[n_planes] or NULL*/, uint32_t flags);

/* out_dev[0] += accepted particles in [first, first+n); out_dev[1] += sum of their Poisson gaps */
int ptb_prefix(PtbHandle h, uint64_t first, uint64_t n, uint64_t seed, const PtbInjected *inj, int64_t *out_dev,
               void *stream);

int ptb_simulate(PtbHandle h, const PtbRun *run, void *stream);

/*
 * Correlation histograms of two hit tables: the event loop of the reference's correlation plot (_eventloop_fast,
 * main/plotting.py:730-764).  For every event number i in [ev_min, ev_max) each hit of table 1 with that number --
 * plus the first hit of the following event, which the reference's loop also uses before it breaks -- is paired with
 * the same set of table 2; x_hist[col1][col2] and y_hist[row1][row2] are incremented (int32, row-major with the given
 * strides, zeroed by the caller, pairs outside the given extents are skipped).  Tables must be sorted by event
 * number and hold the same set of event numbers (the reference filters them so, main/plotting.py:531-538).
 */
int ptb_correlate(const int64_t *ev1, const uint16_t *col1, const uint16_t *row1, uint64_t n1, const int64_t *ev2,
                  const uint16_t *col2, const uint16_t *row2, uint64_t n2, int64_t ev_min, int64_t ev_max,
                  int32_t *x_hist, int32_t x_dim0, int32_t x_dim1, int32_t *y_hist, int32_t y_dim0, int32_t y_dim1,
                  void *stream);

/* SoA slab rows [0, n) -> packed 18-byte records (event_number <i8, frame <u2 = 0, column <u2, row <u2, charge <f4) */
int ptb_pack_hits(const PtbHitSlab *slab, uint64_t n, void *out_dev, void *stream);

/*
 * PtbCompactOut (copied to the host) -> the 18-byte rows of one plane.  hits: the plane's rows; counts: the plane's
 * [n] counts; accepted: [(n+31)/32] words; event_base: bases_in->event of the call.  Writes sum(counts) records to out
 * (host memory, e.g. the mapped dataset region of an HDF5 file) with `threads` host threads (<= 0: one per core, at most
 * 16) and returns the number of rows, or a negative PTB_E_* when n_rows_expected differs from sum(counts).
 */
int64_t ptb_expand_hits(const uint64_t *hits, const uint8_t *counts, const uint32_t *accepted, uint64_t n,
                        int64_t event_base, uint64_t n_rows_expected, void *out, int32_t threads);
/* the same for plane `plane` of the PTB_COMPACT_NARROW form: the plane's slices of the three row arrays, its
 * [(n+1)/2] nibble pairs, and the call's escape entries (all planes; the entries of other planes are skipped) */
int64_t ptb_expand_hits_narrow(const float *charge, const uint16_t *pos_lo, const uint8_t *pos_hi, const uint8_t *counts4,
                               const uint64_t *escapes, uint64_t n_escapes, int32_t plane, const uint32_t *accepted,
                               uint64_t n, int64_t event_base, uint64_t n_rows_expected, void *out, int32_t threads);

/* the same for plane `plane` of the PTB_COMPACT_PACKED form (layout: ptb_packed_layout of the handle that produced it;
 * row_base: the plane's first row counted over all planes = totals.hits[0] + ... + totals.hits[plane-1]) */
int64_t ptb_expand_hits_packed(const uint16_t *word0, const uint16_t *word1, const uint16_t *word2, const uint8_t *counts4,
                               const uint64_t *escapes, uint64_t n_escapes, int32_t plane, const PtbPackedLayout *layout,
                               uint64_t row_base, const uint32_t *accepted, uint64_t n, int64_t event_base,
                               uint64_t n_rows_expected, void *out, int32_t threads);

/* Both expansions with a file instead of a buffer as destination: the plane's records go to file descriptor fd from
 * byte file_offset on, written block by block (positioned writes of 4 MiB per worker thread)
 * -- the dataset region of a *_dut.h5 file is filled without staging the table in memory.  PTB_E_IO when a write fails. */
int64_t ptb_expand_hits_fd(const uint64_t *hits, const uint8_t *counts, const uint32_t *accepted, uint64_t n,
                           int64_t event_base, uint64_t n_rows_expected, int32_t fd, uint64_t file_offset, int32_t threads);
int64_t ptb_expand_hits_narrow_fd(const float *charge, const uint16_t *pos_lo, const uint8_t *pos_hi,
                                  const uint8_t *counts4, const uint64_t *escapes, uint64_t n_escapes, int32_t plane,
                                  const uint32_t *accepted, uint64_t n, int64_t event_base, uint64_t n_rows_expected,
                                  int32_t fd, uint64_t file_offset, int32_t threads);

int64_t ptb_expand_hits_packed_fd(const uint16_t *word0, const uint16_t *word1, const uint16_t *word2,
                                  const uint8_t *counts4, const uint64_t *escapes, uint64_t n_escapes, int32_t plane,
                                  const PtbPackedLayout *layout, uint64_t row_base, const uint32_t *accepted, uint64_t n,
                                  int64_t event_base, uint64_t n_rows_expected, int32_t fd, uint64_t file_offset,
                                  int32_t threads);

/*
 * Checksums of rows [0, n) of a hit table, added (modulo 2^64) to out_dev[0..6]:
 *   rows, sum event_number, sum column, sum row, sum of the charges' bit patterns, S1 = sum m_i, S2 = sum (row_base+i+1) m_i
 * with m_i = a 64-bit mix of row i's (event, column, row, charge bits).  S2 depends on the order of the rows.  For a table
 * cut into pieces (chunks, ranks) the pieces' words add up when every piece is given its global row_base; a piece
 * digested with row_base 0 is moved to base b afterwards by S2 += b * S1.
 */
int ptb_digest(const PtbHitSlab *slab, uint64_t n, uint64_t row_base, uint64_t *out_dev, void *stream);
/*
 * Optional per-kernel timing: while enabled, every ptb_simulate brackets its k_tracks and k_digitise launches
 * with CUDA events on the caller's stream.  ptb_timing_read waits for the recorded events, returns the summed
 * duration of each kernel in milliseconds and the number of ptb_simulate calls covered, and clears the record.
 */
int ptb_timing_enable(PtbHandle h, int32_t on);
int ptb_timing_read(PtbHandle h, double *tracks_ms, double *digitise_ms, uint64_t *calls);
/* the same record by stage: stage_ms[0..3] = k_tracks, k_digitise, k_digitise_big + the two scan kernels, k_gather */
int ptb_timing_read_stages(PtbHandle h, double *stage_ms /*[4]*/, uint64_t *calls);

/* number of kernels this library has launched on the calling process since load (bench bookkeeping) */
uint64_t ptb_launch_count(void);

/* in-job FP64 peak probe: dependent DFMA chains on every SM; returns FLOP/s (2 per DFMA), <0 on error */
double ptb_measure_fp64_peak(void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PTB_H_ */

/*
 * The C ABI of libptb_b200 used from plain C: no Python, no torch -- device memory from the CUDA runtime, the
 * default telescope of the reference's setup.yml, production (Philox) mode.  Prints the hit rows per plane and the
 * first rows of the last plane.
 *
 *   nvcc -x c -I include -o c_abi_example examples/c_abi_example.c -L pytestbeam_b200/_lib -lptb_b200 -lcudart
 *   LD_LIBRARY_PATH=pytestbeam_b200/_lib ./c_abi_example [n_electrons] [seed] [file for the last plane's records] [narrow|packed]
 *
 * With a file name the same particles run once more in a compressed-row form (what a caller ships over PCIe: 7-byte
 * "narrow" rows by default, 6-byte "packed" rows on request), and the library expands the last plane's rows straight
 * into that file (ptb_expand_hits_narrow_fd / ptb_expand_hits_packed_fd).
 */
#include <fcntl.h>
#include <unistd.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ptb.h"

#define CK(call)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));               \
            return 2;                                                                 \
        }                                                                             \
    } while (0)

int main(int argc, char **argv)
{
    const uint64_t n = argc > 1 ? strtoull(argv[1], NULL, 10) : 1000000ull;
    const uint64_t seed = argc > 2 ? strtoull(argv[2], NULL, 10) : 1ull;

    /* pytestbeam/setup.yml:2-109 and material.yml (silicon), flattened the way ptb_create expects them */
    const double z[7] = {0, 30250, 61800, 202550, 225000, 250700, 288150};
    const double dx[7] = {100, 0, 0, 100, 0, 0, 0};
    PtbDevice    dev[7];
    memset(dev, 0, sizeof(dev));
    for (int d = 0; d < 7; ++d) {
        const int itk = d == 6;
        dev[d].ncol = itk ? 400 : 1152;
        dev[d].nrow = itk ? 384 : 576;
        dev[d].pitch_c = dev[d].pitch_r = itk ? 50.0 : 18.4;
        dev[d].thickness = itk ? 300.0 : 50.0;
        dev[d].z = z[d];
        dev[d].delta_x = dx[d];
        dev[d].threshold_e = itk ? 1000.0 : 100.0;
        dev[d].triggered = itk;
        dev[d].rad_length = 93650;
        dev[d].Z = 14;
        dev[d].A = 24;
        dev[d].rho = 2.33;
    }
    PtbBeam beam = {5000.0, 5000.0, 5000.0, 0.0, 0.0, 1e-4, 1e-4, 0.0, 0.0, 5000.0};

    PtbHandle h = NULL;
    int       rc = ptb_create(&beam, dev, 7, NULL, &h);
    if (rc != PTB_OK) {
        fprintf(stderr, "ptb_create: %d %s\n", rc, ptb_last_error(h));
        return 1;
    }

    PtbRun run;
    memset(&run, 0, sizeof(run));
    run.first = 0;
    run.n = n;
    run.seed = seed;
    run.flags = PTB_DIGITISE | PTB_RECYCLE_SLABS;
    uint64_t cap[PTB_MAX_PLANES] = {0};
    for (int d = 0; d < 7; ++d) {
        cap[d] = (uint64_t)(1.5 * (double)n) + 4096; /* 1.2 hits per electron and plane on this telescope */
        run.slabs[d].capacity = cap[d];
        CK(cudaMalloc((void **)&run.slabs[d].event, cap[d] * sizeof(int64_t)));
        CK(cudaMalloc((void **)&run.slabs[d].col, cap[d] * sizeof(uint16_t)));
        CK(cudaMalloc((void **)&run.slabs[d].row, cap[d] * sizeof(uint16_t)));
        CK(cudaMalloc((void **)&run.slabs[d].charge, cap[d] * sizeof(float)));
    }
    CK(cudaMalloc((void **)&run.totals, sizeof(PtbTotals)));
    run.workspace_bytes = ptb_workspace_bytes(h, n, cap, NULL, run.flags);
    CK(cudaMalloc(&run.workspace, run.workspace_bytes));

    rc = ptb_simulate(h, &run, NULL); /* default stream */
    if (rc != PTB_OK) {
        fprintf(stderr, "ptb_simulate: %d %s\n", rc, ptb_last_error(h));
        return 1;
    }
    PtbTotals tot;
    CK(cudaMemcpy(&tot, run.totals, sizeof(tot), cudaMemcpyDeviceToHost)); /* synchronises */
    if (tot.error) {
        fprintf(stderr, "device error bits %#x\n", tot.error);
        return 1;
    }
    printf("electrons %llu accepted %llu pixels %llu\n", (unsigned long long)n, (unsigned long long)tot.accepted,
           (unsigned long long)tot.pixels);
    for (int d = 0; d < 7; ++d) printf("plane %d rows %llu\n", d, (unsigned long long)tot.hits[d]);

    /* the first rows of the last plane, as the reference's (event_number, frame, column, row, charge) records */
    const uint64_t show = tot.hits[6] < 5 ? tot.hits[6] : 5;
    void          *packed = NULL;
    CK(cudaMalloc(&packed, 18 * (show ? show : 1)));
    if (show) {
        rc = ptb_pack_hits(&run.slabs[6], show, packed, NULL);
        if (rc != PTB_OK) return 1;
        unsigned char rows[5 * 18];
        CK(cudaMemcpy(rows, packed, 18 * show, cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < show; ++i) {
            int64_t  ev;
            uint16_t f, c, r;
            float    q;
            memcpy(&ev, rows + 18 * i, 8);
            memcpy(&f, rows + 18 * i + 8, 2);
            memcpy(&c, rows + 18 * i + 10, 2);
            memcpy(&r, rows + 18 * i + 12, 2);
            memcpy(&q, rows + 18 * i + 14, 4);
            printf("itkpix row %llu: event %lld frame %u column %u row %u charge %.6f\n", (unsigned long long)i,
                   (long long)ev, f, c, r, q);
        }
    }
    if (argc > 3) {
        /* the same range as compressed rows: 7 (narrow) or 6 (packed) bytes per hit + 4 bits per (particle, plane) + 1 bit per particle */
        uint64_t rows_cap = 0;
        memset(&run.compact, 0, sizeof(run.compact));
        for (int d = 0; d < 7; ++d) {
            run.compact.plane_first[d] = rows_cap;
            rows_cap += cap[d];
        }
        run.compact.plane_first[7] = rows_cap;
        const uint64_t half = (n + 1) / 2, words = (n + 31) / 32;
        const int       packed = argc > 4 && strcmp(argv[4], "packed") == 0;
        PtbPackedLayout layout;
        if (packed && ptb_packed_layout(h, &layout) != PTB_OK) {
            fprintf(stderr, "ptb_packed_layout: %s\n", ptb_last_error(h));
            return 1;
        }
        run.compact.escape_capacity = 4096;
        if (packed) {
            CK(cudaMalloc((void **)&run.compact.word0, rows_cap * sizeof(uint16_t)));
            CK(cudaMalloc((void **)&run.compact.word1, rows_cap * sizeof(uint16_t)));
            CK(cudaMalloc((void **)&run.compact.word2, rows_cap * sizeof(uint16_t)));
        } else {
            CK(cudaMalloc((void **)&run.compact.charge, rows_cap * sizeof(float)));
            CK(cudaMalloc((void **)&run.compact.pos_lo, rows_cap * sizeof(uint16_t)));
            CK(cudaMalloc((void **)&run.compact.pos_hi, rows_cap));
        }
        CK(cudaMalloc((void **)&run.compact.counts, 7 * half));
        CK(cudaMalloc((void **)&run.compact.accepted, words * sizeof(uint32_t)));
        CK(cudaMalloc((void **)&run.compact.escapes, run.compact.escape_capacity * sizeof(uint64_t)));
        run.flags = PTB_DIGITISE | PTB_RECYCLE_SLABS | PTB_COMPACT | (packed ? PTB_COMPACT_PACKED : PTB_COMPACT_NARROW);
        rc = ptb_simulate(h, &run, NULL);
        if (rc != PTB_OK) {
            fprintf(stderr, "ptb_simulate (compressed rows): %d %s\n", rc, ptb_last_error(h));
            return 1;
        }
        CK(cudaMemcpy(&tot, run.totals, sizeof(tot), cudaMemcpyDeviceToHost));
        if (tot.error) {
            fprintf(stderr, "device error bits %#x (compressed rows)\n", tot.error);
            return 1;
        }
        uint64_t first6 = 0; /* the planes' rows lie back to back */
        for (int d = 0; d < 6; ++d) first6 += tot.hits[d];
        const uint64_t m = tot.hits[6], mm = m ? m : 1;
        uint8_t       *cnt = malloc(half);
        uint32_t      *acc = malloc(words * sizeof(uint32_t));
        uint64_t      *esc = malloc((tot.escapes ? tot.escapes : 1) * sizeof(uint64_t));
        CK(cudaMemcpy(cnt, run.compact.counts + 6 * half, half, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(acc, run.compact.accepted, words * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(esc, run.compact.escapes, tot.escapes * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        const int fd = open(argv[3], O_CREAT | O_TRUNC | O_WRONLY, 0644);
        if (fd < 0) {
            perror(argv[3]);
            return 1;
        }
        int64_t wrote;
        if (packed) {
            uint16_t *w0 = malloc(mm * 2), *w1 = malloc(mm * 2), *w2 = malloc(mm * 2);
            CK(cudaMemcpy(w0, run.compact.word0 + first6, m * 2, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(w1, run.compact.word1 + first6, m * 2, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(w2, run.compact.word2 + first6, m * 2, cudaMemcpyDeviceToHost));
            wrote = ptb_expand_hits_packed_fd(w0, w1, w2, cnt, esc, tot.escapes, 6, &layout, first6, acc, n, 0, m, fd, 0, 0);
        } else {
            float    *q = malloc(mm * sizeof(float));
            uint16_t *lo = malloc(mm * sizeof(uint16_t));
            uint8_t  *hi = malloc(mm);
            CK(cudaMemcpy(q, run.compact.charge + first6, m * sizeof(float), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(lo, run.compact.pos_lo + first6, m * sizeof(uint16_t), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hi, run.compact.pos_hi + first6, m, cudaMemcpyDeviceToHost));
            wrote = ptb_expand_hits_narrow_fd(q, lo, hi, cnt, esc, tot.escapes, 6, acc, n, 0, m, fd, 0, 0);
        }
        close(fd);
        printf("file rows %lld of %llu (%s rows: %llu bytes of compressed rows for all planes)\n", (long long)wrote,
               (unsigned long long)m, packed ? "packed" : "narrow",
               (unsigned long long)((packed ? 6 : 7) * (first6 + m) + 7 * half + 4 * words + 8 * tot.escapes));
        if (wrote != (int64_t)m) return 1;
    }
    ptb_destroy(h);
    return 0;
}

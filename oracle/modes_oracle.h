/*
 * modes_oracle.h — CPU restatement of readsb's 2.4 MSPS Mode-S demodulator path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle for the CUDA path: a plain, sequential C
 * restatement of the reference algorithm (convert.c, demod_2400.c, crc.c, mode_s.c accept logic,
 * icao_filter.c semantics).  Nothing in the product (readsb_b200/, include/) may call, link or import
 * it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Pinning: the reference ships no golden vectors for this path, so this
 * oracle is pinned against the reference itself: oracle/Makefile compiles the reference's own
 * convert.c / demod_2400.c / crc.c / mode_s.c / icao_filter.c (from /root/reference, untouched)
 * into oracle/_ref/libreadsb_ref.so behind oracle/ref_harness.c, and tests/test_oracle_vs_ref.py
 * checks frame-for-frame and counter-for-counter equality on seeded synthetic captures; the
 * outputs of that run are committed as fixtures under tests/golden/ (made by
 * tests/golden/make_golden.py) so the pin also holds where /root/reference is absent.
 */
#ifndef MODES_ORACLE_H
#define MODES_ORACLE_H

#include <stdint.h>
#include "../include/b200_demod.h" /* only for the plain result structs b200_frame etc. */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_ctx oracle_ctx;

oracle_ctx *oracle_create(int preamble_threshold, int nfix_crc, int fix_df, int icao_ttl_ms);
void oracle_destroy(oracle_ctx *o);
void oracle_set_preamble_threshold(oracle_ctx *o, int preamble_threshold);

/* convert.c:35-62 */
void oracle_uc8_lut(uint16_t *out65536);
/* convert.c:64-108 (integer sums instead of the final fp64 divides) */
void oracle_convert_uc8(const uint8_t *iq, uint16_t *mag, unsigned nsamples,
                        uint64_t *sum_level, uint64_t *sum_power);
/* convert.c:212-250 (q11 = 0) / 329-367 (q11 = 1): magnitudes + the two float accumulators */
void oracle_convert_sc16(const int16_t *iq, uint16_t *mag, unsigned nsamples, int q11, float *sum_level, float *sum_power);
/* crc.c:67-82 */
uint32_t oracle_crc24(const uint8_t *msg, int bits);
/* crc.c:383-406 with nfix_crc=1 tables: returns corrected bit (5..bits-1), -1 if syndrome==0, -2 if none */
int oracle_crc_diagnose1(uint32_t syndrome, int bits);
/* --aggressive (nfix_crc = 2) error tables, crc.c:180-378 with max_correct 2 / max_detect 4: preparation, not used by the path yet */
int oracle_crc_diagnose2(uint32_t syndrome, int bits, int *b0, int *b1);
int oracle_crc_table2_digest(int bits, uint64_t *digest);

/* demod_2400.c:264-482 for one mag_buf: data = 326 halo + length new magnitudes.
 * Appends accepted frames to out[*n_out...] (cap entries); returns 0, or -1 if cap was hit. */
int oracle_demodulate2400(oracle_ctx *o, const uint16_t *data, unsigned length, int64_t sample_timestamp,
                          uint64_t sum_level, uint64_t sum_power,
                          b200_frame *out, unsigned cap, unsigned *n_out, b200_buffer_result *res);

/* sdr_ifile.c:169-259 + readsb.c:853-901,1227-1231: replay a uc8 capture as consecutive buffers of
 * buf_samples (last one partial), carrying the 326-sample halo and flipping the ICAO filter on the
 * synthetic clock.  first_ts = 12 MHz timestamp of iq[0].  Returns number of frames, or -1 on overflow. */
long oracle_run_stream_uc8(oracle_ctx *o, const uint8_t *iq, uint64_t nsamples, unsigned buf_samples,
                           int64_t first_ts, b200_frame *frames, unsigned frame_cap,
                           b200_buffer_result *bufres, unsigned bufres_cap, unsigned *n_bufres);

/* The halo kept between oracle_run_stream_uc8 calls is dropped: the next call starts like a fresh stream (zero halo). */
void oracle_stream_restart(oracle_ctx *o);

/* demod_2400.c:575-761 for one mag_buf (needs the converter's sums for mean_level / mean_power).  Appends to out. */
int oracle_demodulate2400AC(oracle_ctx *o, const uint16_t *data, unsigned length, int64_t sample_timestamp,
                            uint64_t sum_level, uint64_t sum_power, b200_modeac *out, unsigned cap, unsigned *n_out);
int oracle_demodulate2400AC_levels(oracle_ctx *o, const uint16_t *data, unsigned length, int64_t sample_timestamp,
                                   double mean_level, double mean_power, b200_modeac *out, unsigned cap, unsigned *n_out);

/* net_io.c:1655-1714 modesSendBeastOutput (+ netTimestamp :1617-1648): one Beast record, at most 44 bytes; returns its length */
unsigned oracle_beast_frame(const b200_frame *f, int verbatim, uint8_t *out);
unsigned oracle_beast_modeac(const b200_modeac *a, uint8_t *out);

void oracle_get_stats(const oracle_ctx *o, b200_demod_stats *out);
void oracle_icao_add(oracle_ctx *o, uint32_t addr);
int  oracle_icao_test(const oracle_ctx *o, uint32_t addr);
void oracle_icao_expire(oracle_ctx *o);

#ifdef __cplusplus
}
#endif
#endif

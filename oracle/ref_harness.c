/*
 * ref_harness.c — drives the UNMODIFIED reference demodulator as a library (TEST INFRASTRUCTURE).
 *
 * Compiled by oracle/Makefile together with the reference's own translation units, taken where
 * they lie under /root/reference (convert.c, demod_2400.c, crc.c, mode_s.c, icao_filter.c and
 * their leaf dependencies mode_ac.c, comm_b.c, ais_charset.c), into oracle/_ref/libreadsb_ref.so.
 * No reference source is copied into this repository; this file only supplies what those objects
 * import from the rest of readsb:
 *   - the `Modes` global (readsb.c:61) with the hot-path options set as configSetDefaults /
 *     configAfterParse do (readsb.c:150,194,2268-2270),
 *   - the message sink netGetMM / netUseMessage / netDrainMessageBuffers (net_io.c:5978-6012),
 *     here a capture buffer,
 *   - the ifile replay loop (sdr_ifile.c:169-259) and the decode loop's filter flip
 *     (readsb.c:901,1227-1231), restated around the real converter and the real demodulate2400(),
 *   - receiveclock_ms_elapsed (util.c:149-151) and two display-only helpers as stubs.
 * The results are reported in the same plain structs the product uses (include/b200_demod.h) so
 * that tests compare reference, oracle and CUDA output field by field.
 */
#include "readsb.h"
#include "../include/b200_demod.h"

struct _Modes Modes;

void setExit(int arg) { (void)arg; }
int64_t receiveclock_ms_elapsed(int64_t t1, int64_t t2) { return (t2 - t1) / 12000U; }
void printACASInfoShort(uint32_t addr, unsigned char *MV, struct aircraft *a, struct modesMessage *mm, int64_t now) {
    (void)addr; (void)MV; (void)a; (void)mm; (void)now;
}
char *sprint_uuid1(uint64_t id1, char *p) { (void)id1; return p; }

/* ---- message sink ------------------------------------------------------------------------- */
static struct modesMessage g_mm;          /* the slot netGetMM hands out */
static struct messageBuffer g_mb;
static b200_frame *g_out;                 /* capture array for the current call */
static double *g_levels;                  /* mm->signalLevel per captured frame */
static unsigned g_cap, g_n, g_overflow;
static int64_t g_cur_sample_ts;
static uint32_t g_buffer_seq;
static unsigned g_buf_frames;
static b200_modeac *g_ac_out;            /* Mode A/C capture (decodeModeAMessage marks mm->msgtype DFTYPE_MODEAC) */
static unsigned g_ac_cap, g_ac_n, g_ac_overflow;

struct modesMessage *netGetMM(struct messageBuffer *buf) {
    memset(&g_mm, 0, sizeof g_mm);
    g_mm.messageBuffer = buf;
    return &g_mm;
}

void netUseMessage(struct modesMessage *mm) {
    if (mm->msgtype == DFTYPE_MODEAC) {
        if (g_ac_n >= g_ac_cap) { g_ac_overflow = 1; return; }
        b200_modeac *a = &g_ac_out[g_ac_n++];
        a->timestamp = mm->timestamp; a->f1_sample = 0; a->buffer_idx = 0;
        a->modeac = (uint16_t)((mm->msg[0] << 8) | mm->msg[1]);           /* mode_ac.c:171-173 */
        return;
    }
    g_buf_frames++;
    if (g_n >= g_cap) { g_overflow = 1; return; }
    b200_frame *f = &g_out[g_n];
    memset(f, 0, sizeof *f);
    f->timestamp = mm->timestamp;
    /* timestamp = sampleTimestamp + 5*j + 768 + phase, phase in 4..8 (demod_2400.c:406) */
    int64_t x = mm->timestamp - g_cur_sample_ts - 768;
    int phase = (int)((x + 1) % 5) + 4;   /* x%5: 4->4, 0->5, 1->6, 2->7, 3->8 */
    f->phase = (uint8_t)phase;
    f->j = (uint32_t)((x - phase) / 5);
    f->crc = mm->crc;
    f->addr = mm->addr;
    f->score = mm->score;
    f->buffer_seq = g_buffer_seq;
    f->msgtype = (uint8_t)mm->msgtype;
    f->msgbits = (uint8_t)mm->msgbits;
    f->correctedbits = (uint8_t)mm->correctedbits;
    memcpy(f->msg, mm->msg, (size_t)mm->msgbits / 8);
    /* verbatim = frame as sliced (Modes.net_verbatim, mode_s.c:444-447): locate the corrected bit */
    f->fix_bit = -1;
    for (int b = 0; b < 112; b++) {
        int byte = b >> 3, mask = 1 << (7 - (b & 7));
        if (byte < mm->msgbits / 8 && ((mm->msg[byte] ^ mm->verbatim[byte]) & mask)) { f->fix_bit = (int8_t)b; break; }
    }
    f->signal_len = (uint16_t)(((mm->verbatim[0] >> 3) & 0x10) ? 268 : 134);
    if (g_levels) g_levels[g_n] = mm->signalLevel;
    g_n++;
}

void netDrainMessageBuffers() {}

/* ---- setup --------------------------------------------------------------------------------- */
static iq_convert_fn g_conv;
static struct converter_state *g_conv_state;
static int g_ttl_ms;
static int64_t g_next_flip;
static uint16_t g_halo[B200_TRAILING_SAMPLES];
static int g_halo_valid;
static uint64_t g_flips, g_buffers;

int ref_init(int preamble_threshold, int nfix_crc, int fix_df, int icao_ttl_ms) {
    memset(&Modes, 0, sizeof Modes);
    Modes.sample_rate = 2400000.0;
    Modes.trailing_samples = B200_TRAILING_SAMPLES;       /* readsb.c:288 */
    Modes.nfix_crc = (int8_t)(nfix_crc ? 1 : 0);          /* readsb.c:150 */
    Modes.fixDF = (int8_t)(fix_df ? 1 : 0);               /* readsb.c:194 */
    Modes.preambleThreshold = (uint32_t)(preamble_threshold ? preamble_threshold : PREAMBLE_THRESHOLD_DEFAULT);
    Modes.net_verbatim = 1;                               /* keep the uncorrected frame in mm->verbatim */
    Modes.sdr_type = SDR_NONE;                            /* silence init_converter's banner */
    Modes.startup_time = 1000000;
    Modes.decodeThreads = 1;
    Modes.netMessageBuffer = &g_mb;
    modesChecksumInit(Modes.nfix_crc);                    /* readsb.c:306 */
    icaoFilterInit();                                     /* readsb.c:307 */
    g_conv = init_converter(INPUT_UC8, Modes.sample_rate, 0, &g_conv_state); /* sdr_ifile.c:156 */
    Modes.sdr_type = SDR_IFILE;
    Modes.synthetic_now = Modes.startup_time;             /* sdr_ifile.c:132 (a fixed epoch here) */
    g_ttl_ms = icao_ttl_ms == 0 ? MODES_ICAO_FILTER_TTL : icao_ttl_ms;
    g_next_flip = 0;
    g_halo_valid = 0;
    g_buffer_seq = 0;
    g_flips = g_buffers = 0;
    return g_conv ? 0 : -1;
}

/* Modes.stats_15min.samples_dropped != 0 makes demodulate2400 use max(75, preambleThreshold) (demod_2400.c:334-338). */
void ref_set_samples_dropped(unsigned n) { Modes.stats_15min.samples_dropped = n; }

void ref_uc8_lut(uint16_t *out65536) {
    uint8_t *iq = malloc(65536 * 2);
    for (int i = 0; i < 256; i++)
        for (int q = 0; q < 256; q++) { iq[2 * (i * 256 + q)] = (uint8_t)i; iq[2 * (i * 256 + q) + 1] = (uint8_t)q; }
    g_conv(iq, out65536, 65536, g_conv_state, NULL, NULL);
    free(iq);
}

/* The reference's own float-path converters (static in convert.c, reached through init_converter like a frontend does). */
int ref_convert_sc16(const int16_t *iq, uint16_t *mag, unsigned n, int q11, double *mean_level, double *mean_power) {
    static iq_convert_fn fn[2];
    static struct converter_state *st[2];
    if (!fn[q11 != 0]) fn[q11 != 0] = init_converter(q11 ? INPUT_SC16Q11 : INPUT_SC16, Modes.sample_rate, 0, &st[q11 != 0]);
    if (!fn[q11 != 0]) return -1;
    fn[q11 != 0]((void *)iq, mag, n, st[q11 != 0], mean_level, mean_power);
    return 0;
}

void ref_convert_uc8(const uint8_t *iq, uint16_t *mag, unsigned n, double *mean_level, double *mean_power) {
    g_conv((void *)iq, mag, n, g_conv_state, mean_level, mean_power);
}

uint32_t ref_crc24(const uint8_t *msg, int bits) { return modesChecksum((uint8_t *)msg, bits); }

int ref_crc_diagnose1(uint32_t syndrome, int bits) {
    struct errorinfo *ei = modesChecksumDiagnose(syndrome, bits);
    if (!ei) return -2;
    if (ei->errors == 0) return -1;
    return ei->bit[0];
}

/* The reference's own --aggressive tables (modesChecksumInit(2)): number of syndromes modesChecksumDiagnose knows and a digest
 * over (syndrome, errors, bits) in ascending syndrome order.  Re-initialises the tables: call ref_init again afterwards. */
int ref_crc_table2_digest(int bits, uint64_t *digest) {
    modesChecksumInit(2);
    uint64_t h = 1469598103934665603ull;
    int n = 0;
    for (uint32_t syn = 1; syn < (1u << 24); syn++) {
        struct errorinfo *ei = modesChecksumDiagnose(syn, bits);
        if (!ei) continue;
        const int b0 = ei->bit[0], b1 = ei->errors > 1 ? ei->bit[1] : -1;
        const uint64_t v = ((uint64_t)syn << 16) | ((uint64_t)(uint8_t)b0 << 8) | (uint8_t)b1;
        h = (h ^ v) * 1099511628211ull;
        n++;
    }
    *digest = h;
    return n;
}

int ref_score(const uint8_t *msg14, int validbits) {
    uint8_t tmp[14];
    memcpy(tmp, msg14, 14);
    return scoreModesMessage(tmp, validbits);
}

void ref_icao_add(uint32_t a) { icaoFilterAdd(a); }
int ref_icao_test(uint32_t a) { return icaoFilterTest(a); }
void ref_icao_expire(void) { icaoFilterExpire(); g_flips++; }

/* One call of the real demodulate2400() on one mag_buf, then the decode loop's flip check. */
int ref_demodulate2400(uint16_t *data, unsigned length, int64_t sample_ts, double mean_level, double mean_power,
                       b200_frame *out, double *levels, unsigned cap, unsigned *n_out, b200_buffer_result *res) {
    struct mag_buf mb;
    memset(&mb, 0, sizeof mb);
    mb.sampleTimestamp = sample_ts;
    mb.sysTimestamp = sample_ts / 12000U + Modes.startup_time;   /* sdr_ifile.c:216 */
    mb.mean_level = mean_level;
    mb.mean_power = mean_power;
    mb.length = length;
    mb.data = data;
    g_out = out; g_levels = levels; g_cap = cap; g_n = *n_out; g_overflow = 0;
    g_cur_sample_ts = sample_ts; g_buf_frames = 0;
    const struct stats before = Modes.stats_current;
    demodulate2400(&mb);                                         /* readsb.c:871 */
    Modes.stats_current.samples_processed += length;             /* readsb.c:876 */
    *n_out = g_n;
    g_buffers++;
    int flipped = 0;
    int64_t now = Modes.synthetic_now;                           /* mstime(), util.c:58-60 */
    if (g_ttl_ms > 0 && now >= g_next_flip) {                    /* readsb.c:1227-1231 */
        icaoFilterExpire();
        g_next_flip = now + g_ttl_ms;
        flipped = 1; g_flips++;
    }
    if (res) {
        memset(res, 0, sizeof *res);
        res->sample_timestamp = sample_ts; res->length = length; res->n_frames = g_buf_frames;
        res->buffer_seq = g_buffer_seq; res->icao_flipped = (uint32_t)flipped;
        const struct stats *s = &Modes.stats_current;             /* what this call added (stats.h:62-83) */
        res->demod_preambles = s->demod_preambles - before.demod_preambles;
        res->demod_rejected_bad = s->demod_rejected_bad - before.demod_rejected_bad;
        res->demod_rejected_unknown_icao = s->demod_rejected_unknown_icao - before.demod_rejected_unknown_icao;
        for (int i = 0; i < 2; i++) res->demod_accepted[i] = s->demod_accepted[i] - before.demod_accepted[i];
        for (int i = 0; i < 5; i++) {
            res->demod_preamblePhase[i] = s->demod_preamblePhase[i] - before.demod_preamblePhase[i];
            res->demod_bestPhase[i] = s->demod_bestPhase[i] - before.demod_bestPhase[i];
        }
    }
    g_buffer_seq++;
    return g_overflow ? -1 : 0;
}

/* The real demodulate2400AC() on one mag_buf (readsb.c:872-874). */
int ref_demodulate2400AC(uint16_t *data, unsigned length, int64_t sample_ts, double mean_level, double mean_power,
                         b200_modeac *out, unsigned cap, unsigned *n_out) {
    struct mag_buf mb;
    memset(&mb, 0, sizeof mb);
    mb.sampleTimestamp = sample_ts;
    mb.sysTimestamp = sample_ts / 12000U + Modes.startup_time;
    mb.mean_level = mean_level; mb.mean_power = mean_power;
    mb.length = length; mb.data = data;
    g_ac_out = out; g_ac_cap = cap; g_ac_n = *n_out; g_ac_overflow = 0;
    demodulate2400AC(&mb);
    *n_out = g_ac_n;
    return g_ac_overflow ? -1 : 0;
}

uint64_t ref_modeac_count(void) { return Modes.stats_current.demod_modeac; }

/* The ifile replay loop (sdr_ifile.c:169-259) around the real converter and demodulator.
 * mean_levels/mean_powers (optional) receive the converter's two doubles per buffer. */
long ref_run_stream_uc8(const uint8_t *iq, uint64_t nsamples, unsigned buf_samples, int64_t first_ts,
                        b200_frame *frames, double *levels, unsigned frame_cap,
                        b200_buffer_result *bufres, double *mean_levels, double *mean_powers,
                        unsigned bufres_cap, unsigned *n_bufres) {
    uint16_t *data = calloc((size_t)buf_samples + B200_TRAILING_SAMPLES, 2);
    unsigned nf = 0, nb = 0;
    int bad = 0;
    for (uint64_t off = 0; off < nsamples; off += buf_samples) {
        unsigned len = (unsigned)((nsamples - off < buf_samples) ? nsamples - off : buf_samples);
        if (g_halo_valid) memcpy(data, g_halo, sizeof g_halo); else memset(data, 0, sizeof g_halo);
        double ml = 0, mp = 0;
        g_conv((void *)(iq + off * 2), data + B200_TRAILING_SAMPLES, len, g_conv_state, &ml, &mp);
        b200_buffer_result r;
        if (ref_demodulate2400(data, len, first_ts + (int64_t)off * 5, ml, mp, frames, levels, frame_cap, &nf, &r) < 0) bad = 1;
        if (nb < bufres_cap) {
            if (bufres) bufres[nb] = r;
            if (mean_levels) mean_levels[nb] = ml;
            if (mean_powers) mean_powers[nb] = mp;
        }
        nb++;
        if (len >= B200_TRAILING_SAMPLES) { memcpy(g_halo, data + len, sizeof g_halo); g_halo_valid = 1; } else g_halo_valid = 0;
    }
    free(data);
    if (n_bufres) *n_bufres = nb;
    return bad ? -1 : (long)nf;
}

/* Modes.stats_current demod counters; the three fp64 accumulators are returned separately. */
void ref_get_stats(b200_demod_stats *out, double *signal_power_sum, double *noise_power_sum, double *peak_signal_power) {
    struct stats *s = &Modes.stats_current;
    memset(out, 0, sizeof *out);
    out->samples_processed = s->samples_processed;
    out->demod_preambles = s->demod_preambles;
    out->demod_rejected_bad = s->demod_rejected_bad;
    out->demod_rejected_unknown_icao = s->demod_rejected_unknown_icao;
    out->demod_accepted[0] = s->demod_accepted[0];
    out->demod_accepted[1] = s->demod_accepted[1];
    for (int i = 0; i < 5; i++) { out->demod_preamblePhase[i] = s->demod_preamblePhase[i]; out->demod_bestPhase[i] = s->demod_bestPhase[i]; }
    out->signal_power_count = s->signal_power_count;
    out->strong_signal_count = s->strong_signal_count;
    out->peak_signal_power = s->peak_signal_power;
    out->buffers = g_buffers;
    out->icao_flips = g_flips;
    if (signal_power_sum) *signal_power_sum = s->signal_power_sum;
    if (noise_power_sum) *noise_power_sum = s->noise_power_sum;
    if (peak_signal_power) *peak_signal_power = s->peak_signal_power;
}

/* Thread CPU seconds spent in converter + demodulate2400 for a whole capture: the figure readsb's
 * --stats prints as "ms for demodulation" + "ms for reading from USB" (stats.c:183-188), used by
 * bench.py as the reference CPU baseline. */
double ref_time_stream_uc8(const uint8_t *iq, uint64_t nsamples, unsigned buf_samples, unsigned *n_frames) {
    static b200_frame scratch[4096];
    struct timespec a, b;
    uint16_t *data = calloc((size_t)buf_samples + B200_TRAILING_SAMPLES, 2);
    unsigned total = 0;
    clock_gettime(CLOCK_THREAD_CPUTIME_ID, &a);
    for (uint64_t off = 0; off < nsamples; off += buf_samples) {
        unsigned len = (unsigned)((nsamples - off < buf_samples) ? nsamples - off : buf_samples);
        if (g_halo_valid) memcpy(data, g_halo, sizeof g_halo); else memset(data, 0, sizeof g_halo);
        double ml = 0, mp = 0;
        g_conv((void *)(iq + off * 2), data + B200_TRAILING_SAMPLES, len, g_conv_state, &ml, &mp);
        unsigned nf = 0;
        ref_demodulate2400(data, len, (int64_t)off * 5, ml, mp, scratch, NULL, 4096, &nf, NULL);
        total += g_buf_frames;
        if (len >= B200_TRAILING_SAMPLES) { memcpy(g_halo, data + len, sizeof g_halo); g_halo_valid = 1; } else g_halo_valid = 0;
    }
    clock_gettime(CLOCK_THREAD_CPUTIME_ID, &b);
    free(data);
    if (n_frames) *n_frames = total;
    return (double)(b.tv_sec - a.tv_sec) + (double)(b.tv_nsec - a.tv_nsec) * 1e-9;
}

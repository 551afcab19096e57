/*
 * readsb_shim_iq.c — optional second hook of the reference-side binding: the CONVERTER call site.
 *
 * Every SDR frontend of readsb hands its raw samples to an `iq_convert_fn` obtained from init_converter()
 * (convert.h:34-46; call sites sdr_ifile.c:156,241, sdr_rtlsdr.c:271,395 ...) and gets magnitudes back.  Linked with
 *
 *     -Wl,--wrap=init_converter        (in addition to the wraps of readsb_shim.c)
 *
 * this file returns, for INPUT_UC8 without DC filter, a converter that does NOT convert: it keeps the raw IQ bytes of the
 * mag_buf being filled (pinned staging, one slot per ring buffer) and the demodulator hook (readsb_shim.c) later submits
 * those bytes with b200_demod_submit_iq_uc8 — the uc8 -> magnitude conversion (convert.c:64-108), its exact level / power
 * sums included, happens on the GPU, fused into the scan kernel.  The reader thread's per-buffer work drops from a table
 * lookup per sample to a memcpy.  Other formats / DC filtering keep the reference's own converter.
 *
 * mag_buf.data of such a buffer is NOT filled (nothing on this path reads it: demodulate2400 and demodulate2400AC are both
 * redirected); mag_buf.mean_level / mean_power are set by the demodulator hook from the library's sums before anything reads them.
 */
#include "readsb.h"
#include "b200_demod.h"

iq_convert_fn __real_init_converter(input_format_t format, double sample_rate, int filter_dc, struct converter_state **out_state);

#define IQ_SLOTS (MODES_MAG_BUFFERS + 4)
static struct { const uint16_t *mag_data; uint8_t *iq; unsigned n; } g_slot[IQ_SLOTS];
static int g_nslot;

/* Called by readsb_shim.c: the raw IQ captured for the mag_buf whose new samples start at `mag_data`, or 0. */
int b200_shim_iq_lookup(const uint16_t *mag_data, const uint8_t **iq, unsigned *n) {
    const int cnt = __atomic_load_n(&g_nslot, __ATOMIC_ACQUIRE);
    for (int i = 0; i < cnt; i++)
        if (g_slot[i].mag_data == mag_data) { *iq = g_slot[i].iq; *n = g_slot[i].n; return 1; }
    return 0;
}

static void capture_uc8(void *iq_data, uint16_t *mag_data, unsigned nsamples, struct converter_state *state, double *out_mean_level, double *out_mean_power) {
    (void) state;
    int i, cnt = g_nslot;
    for (i = 0; i < cnt; i++) if (g_slot[i].mag_data == mag_data) break;
    if (i == cnt) {                                     /* first time this ring buffer is filled (reader thread only) */
        if (cnt == IQ_SLOTS) { fprintf(stderr, "b200 shim: more mag_bufs than expected\n"); setExit(2); return; }
        g_slot[i].iq = b200_demod_host_alloc((size_t) Modes.sdr_buf_samples * 2);      /* pinned: the upload is a direct DMA */
        if (!g_slot[i].iq) { setExit(2); return; }
        g_slot[i].mag_data = mag_data;
        __atomic_store_n(&g_nslot, cnt + 1, __ATOMIC_RELEASE);
    }
    memcpy(g_slot[i].iq, iq_data, (size_t) nsamples * 2);
    g_slot[i].n = nsamples;
    *out_mean_level = 0; *out_mean_power = 0;           /* the demodulator hook fills them in from the GPU's exact sums */
}

iq_convert_fn __wrap_init_converter(input_format_t format, double sample_rate, int filter_dc, struct converter_state **out_state) {
    iq_convert_fn real = __real_init_converter(format, sample_rate, filter_dc, out_state);     /* state, tables, error handling as ever */
    if (real && format == INPUT_UC8 && !filter_dc) {
        if (Modes.sdr_type == SDR_IFILE) fprintf(stderr, "init_converter: UC8 conversion moved to the GPU (b200)\n");
        return capture_uc8;
    }
    return real;
}

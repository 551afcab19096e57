/* latency_probe.c — bench.py's stopwatch for the drop-in's own call shape (BASELINE configs[1]): one receiver, one 65536-sample
 * mag_buf per call, blocking submit -> run -> fetch, exactly what integration/readsb_shim.c does per demodulate2400() call
 * (readsb.c:866-878).  Written in C so that the per-call time is the library's, not the Python wrapper's.  Measurement
 * infrastructure: it only calls the public C ABI (include/b200_demod.h) through the function pointers bench.py hands it. */
#define _POSIX_C_SOURCE 200809L
#include <stdint.h>
#include <string.h>
#include <time.h>
#include "b200_demod.h"

typedef int (*submit_mag_fn)(b200_demod_ctx *, uint32_t, const uint16_t *, uint32_t, int64_t);
typedef int (*submit_iq_fn)(b200_demod_ctx *, uint32_t, const uint8_t *, uint32_t, int64_t);
typedef int (*run_fn)(b200_demod_ctx *);
typedef int (*fetch_fn)(b200_demod_ctx *, uint32_t, b200_frame *, uint32_t, uint32_t *);

static double now_us(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e6 + t.tv_nsec * 1e-3;
}

/* bufs: nbuf buffers back to back, each `stride_bytes` apart: mag hand-off = (326 + n) uint16 (halo first), iq = 2 n bytes.
 * Returns 0, or the library's error code; us_out[reps] per-call latencies, *frames = frames fetched over all calls. */
int probe_latency(b200_demod_ctx *ctx, submit_mag_fn submit_mag, submit_iq_fn submit_iq, run_fn run, fetch_fn fetch,
                  const void *bufs, uint64_t stride_bytes, uint32_t n, uint32_t nbuf, uint32_t reps, int is_iq,
                  double *us_out, uint64_t *frames) {
    static b200_frame out[4096];
    uint64_t total = 0;
    for (uint32_t r = 0; r < reps; r++) {
        const uint8_t *p = (const uint8_t *)bufs + (uint64_t)(r % nbuf) * stride_bytes;
        const int64_t ts = (int64_t)r * n * 5;
        uint32_t got = 0;
        const double t0 = now_us();
        int rc = is_iq ? submit_iq(ctx, 0, p, n, ts) : submit_mag(ctx, 0, (const uint16_t *)p, n, ts);
        if (rc == 0) rc = run(ctx);
        if (rc == 0) rc = fetch(ctx, 0, out, 4096, &got);
        us_out[r] = now_us() - t0;
        if (rc != 0) return rc;
        total += got;
    }
    *frames = total;
    return 0;
}

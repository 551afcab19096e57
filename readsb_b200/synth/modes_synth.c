/*
 * modes_synth.c — deterministic synthetic 2.4 MSPS uc8 IQ generator (see modes_synth.h).
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fPIC -shared modes_synth.c -o libmodes_synth.so -lm
 */
#include "modes_synth.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t sm64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t *s) { return (double)(sm64(s) >> 11) * (1.0 / 9007199254740992.0); }

/* CRC-24 remainder of the first nbits_total-24 bits (Mode-S generator 0xFFF409, MSB first). */
uint32_t synth_crc24(const uint8_t *msg, int nbits_total) {
    uint32_t rem = 0;
    for (int b = 0; b < nbits_total - 24; b++) {
        uint32_t in = (msg[b >> 3] >> (7 - (b & 7))) & 1u;
        uint32_t fb = ((rem >> 23) & 1u) ^ in;
        rem = (rem << 1) & 0xffffffu;
        if (fb) rem ^= 0xfff409u;
    }
    return rem;
}

static void put_pi(uint8_t *msg, int nbits, uint32_t overlay) {
    uint32_t pi = synth_crc24(msg, nbits) ^ overlay;
    int n = nbits / 8;
    msg[n - 3] = (uint8_t)(pi >> 16); msg[n - 2] = (uint8_t)(pi >> 8); msg[n - 1] = (uint8_t)pi;
}

static uint32_t icao_of(uint64_t seed, uint32_t i) {
    uint64_t s = seed * 0x100000001B3ull + i * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t a = (uint32_t)(sm64(&s) & 0xffffff);
    return a ? a : 0x4840d6;
}

/* Build one reply of the given kind; returns bit length. */
static int build_frame(uint32_t kind, uint64_t *rng, uint32_t addr, uint8_t *msg) {
    memset(msg, 0, 14);
    uint64_t r = sm64(rng), r2 = sm64(rng);
    if (kind == SYNTH_DF17 || kind == SYNTH_DF18) {
        int df = kind == SYNTH_DF17 ? 17 : 18;
        msg[0] = (uint8_t)((df << 3) | (kind == SYNTH_DF17 ? 5 : (int)(r & 1)));
        msg[1] = (uint8_t)(addr >> 16); msg[2] = (uint8_t)(addr >> 8); msg[3] = (uint8_t)addr;
        for (int i = 0; i < 7; i++) msg[4 + i] = (uint8_t)(r2 >> (8 * i));
        put_pi(msg, 112, 0);
        return 112;
    }
    if (kind == SYNTH_DF11 || kind == SYNTH_DF11_IID) {
        msg[0] = (uint8_t)((11 << 3) | 5);
        msg[1] = (uint8_t)(addr >> 16); msg[2] = (uint8_t)(addr >> 8); msg[3] = (uint8_t)addr;
        put_pi(msg, 56, kind == SYNTH_DF11_IID ? (uint32_t)(1 + (r % 15)) : 0);
        return 56;
    }
    /* address/parity kinds */
    static const int ap_df[6] = {0, 4, 5, 16, 20, 21};
    int df = ap_df[r % 6];
    int nbits = (df & 0x10) ? 112 : 56;
    msg[0] = (uint8_t)((df << 3) | ((r >> 8) & 7));
    for (int i = 1; i < nbits / 8 - 3; i++) msg[i] = (uint8_t)(r2 >> (8 * (i % 8))) ^ (uint8_t)(r >> (16 + (i % 5)));
    put_pi(msg, nbits, addr);
    return nbits;
}

long synth_generate_uc8(const synth_params *p, uint64_t nsamples, uint8_t *iq,
                        synth_truth *truth, uint64_t truth_cap) {
    uint64_t rng = p->seed * 0xD6E8FEB86659FD93ull + 0x5851F42D4C957F2Dull;
    float *acc = calloc((size_t)nsamples * 2 + 16, sizeof(float));
    if (!acc) return -1;

    uint32_t kinds[6]; int nk = 0;
    for (uint32_t k = 1; k <= SYNTH_MODEAC; k <<= 1) if (p->df_mask & k) kinds[nk++] = k;
    uint64_t nframes = nk ? (uint64_t)(p->frames_per_sec * (double)nsamples / 2.4e6 + 0.5) : 0;
    int64_t total_ticks = (int64_t)nsamples * 5;
    uint32_t n_icao = p->n_icao ? p->n_icao : 1;
    long injected = 0;

    for (uint64_t f = 0; f < nframes; f++) {
        uint8_t msg[14];
        uint32_t kind = kinds[f % (uint64_t)nk];
        if (kind == SYNTH_MODEAC) {
            /* 20 bit cells of 87 cycles at 60 MHz (one sample = 25 cycles): F1 C1 A1 C2 A2 C4 A4 X B1 D1 B2 D2 B4 D4 F2 X X SPI X X,
             * each "on" cell = 27 cycles high.  Start uniformly random at 60 MHz resolution. */
            uint64_t r = sm64(&rng);
            uint32_t cells = 0x80020u | ((uint32_t)(r & 0x3f) << 13) | ((uint32_t)((r >> 6) & 0x3f) << 6) | (((r >> 12) & 7) == 0 ? 0x4u : 0u);
            double us = u01(&rng), ua = u01(&rng), uph = u01(&rng);
            int64_t total_cyc = (int64_t)nsamples * 25;
            if (total_cyc <= 20 * 87 + 100) break;
            int64_t c0 = (int64_t)(us * (double)(total_cyc - 20 * 87 - 100));
            double amp = (p->amp_min + (p->amp_max - p->amp_min) * ua) * 127.5;
            float ci = (float)(amp * cos(6.283185307179586 * uph)), cq = (float)(amp * sin(6.283185307179586 * uph));
            for (int b = 0; b < 20; b++) {
                if (!((cells >> (19 - b)) & 1u)) continue;
                int64_t a0 = c0 + 87 * b, a1 = a0 + 27;                 /* high during [a0, a1) cycles */
                for (int64_t n = a0 / 25; n <= (a1 - 1) / 25 && n < (int64_t)nsamples; n++) {
                    int64_t lo = n * 25 > a0 ? n * 25 : a0, hi = (n + 1) * 25 < a1 ? (n + 1) * 25 : a1;
                    if (hi > lo) { float w = (float)(hi - lo) * 0.04f; acc[2 * n] += w * ci; acc[2 * n + 1] += w * cq; }
                }
            }
            if (truth && (uint64_t)injected < truth_cap) {
                synth_truth *t = &truth[injected];
                t->start_tick = c0 / 5; memset(t->msg, 0, 14); t->msg[0] = (uint8_t)(cells >> 16); t->msg[1] = (uint8_t)(cells >> 8); t->msg[2] = (uint8_t)cells;
                t->nbits = 24; t->errors = 0;
            }
            injected++;
            continue;
        }
        uint32_t addr = icao_of(p->seed, (uint32_t)(sm64(&rng) % n_icao));
        int nbits = build_frame(kind, &rng, addr, msg);
        int errors = 0;
        double pe = u01(&rng);
        if (pe < p->p_two_bit_error) errors = 2; else if (pe < p->p_two_bit_error + p->p_bit_error) errors = 1;
        for (int e = 0; e < errors; e++) { int b = (int)(sm64(&rng) % (uint64_t)nbits); msg[b >> 3] ^= (uint8_t)(1u << (7 - (b & 7))); }

        int len_ticks = 96 + 12 * nbits;
        double us = u01(&rng), ua = u01(&rng), uph = u01(&rng);
        if (total_ticks <= len_ticks + 10) break;
        int64_t t0 = (int64_t)(us * (double)(total_ticks - len_ticks - 10));
        double amp = (p->amp_min + (p->amp_max - p->amp_min) * ua) * 127.5;
        float ci = (float)(amp * cos(6.283185307179586 * uph)), cq = (float)(amp * sin(6.283185307179586 * uph));

        /* tick-level envelope */
        uint8_t env[96 + 12 * 112 + 8];
        memset(env, 0, sizeof env);
        static const int pre[4] = {0, 12, 42, 54};
        for (int k = 0; k < 4; k++) memset(env + pre[k], 1, 6);
        for (int b = 0; b < nbits; b++) {
            int bit = (msg[b >> 3] >> (7 - (b & 7))) & 1;
            memset(env + 96 + 12 * b + (bit ? 0 : 6), 1, 6);
        }
        int64_t n0 = t0 / 5, n1 = (t0 + len_ticks + 4) / 5;
        for (int64_t n = n0; n <= n1 && n < (int64_t)nsamples; n++) {
            int cnt = 0;
            for (int t = 0; t < 5; t++) {
                int64_t rel = n * 5 + t - t0;
                if (rel >= 0 && rel < len_ticks) cnt += env[rel];
            }
            if (cnt) {
                float w = (float)cnt * 0.2f;
                acc[2 * n] += w * ci; acc[2 * n + 1] += w * cq;
            }
        }
        if (truth && (uint64_t)injected < truth_cap) {
            synth_truth *t = &truth[injected];
            t->start_tick = t0; memcpy(t->msg, msg, 14); t->nbits = (uint8_t)nbits; t->errors = (uint8_t)errors;
        }
        injected++;
    }

    /* noise: Irwin-Hall(4 bytes) per component, sd of the byte sum = sqrt(4*(256^2-1)/12) */
    const float nscale = (float)(p->noise_sigma / 147.80054127098534);
    for (uint64_t n = 0; n < nsamples; n++) {
        uint64_t r = sm64(&rng);
        int si = (int)(r & 0xff) + (int)((r >> 8) & 0xff) + (int)((r >> 16) & 0xff) + (int)((r >> 24) & 0xff) - 510;
        int sq = (int)((r >> 32) & 0xff) + (int)((r >> 40) & 0xff) + (int)((r >> 48) & 0xff) + (int)((r >> 56) & 0xff) - 510;
        float xi = 128.0f + acc[2 * n] + (float)si * nscale;   /* 127.5 + s, then +0.5 to round */
        float xq = 128.0f + acc[2 * n + 1] + (float)sq * nscale;
        int vi = (int)floorf(xi), vq = (int)floorf(xq);
        iq[2 * n] = (uint8_t)(vi < 0 ? 0 : vi > 255 ? 255 : vi);
        iq[2 * n + 1] = (uint8_t)(vq < 0 ? 0 : vq > 255 ? 255 : vq);
    }
    free(acc);
    return injected;
}

// scan_kernel.cu — stage A of the Mode-S demodulator: the stateless, data-parallel part of
// demodulate2400() as one persistent sm_100a kernel.
//
//   per tile of SCAN_TILE preamble start positions (+ SCAN_LOOKAHEAD samples of look-ahead):
//     1. coalesced 16-byte HBM loads of uc8 IQ, magnitude through a folded, bank-swizzled
//        lookup table in shared memory (convert.c:35-108), exact per-buffer level/power sums;
//        magnitudes staged in shared memory as uint16
//     2. pre-check of every position (demod_2400.c:311-320), warp ballots -> ordered compaction
//     3. noise-relative thresholds of the three preamble correlations (demod_2400.c:330-378)
//     4. per (position, phase): PPM/Manchester slicer (demod_2400.c:74-213 in closed form), DF gate,
//        CRC-24 syndrome, DF17 repair / single-bit-fix classification (crc.c, mode_s.c:276-419)
//     5. ordered emission of PosEntry + Rec lists for the per-receiver resolver (stage B)
//
// No tensor cores: integer scan/correlate work bounded by HBM reads and instruction issue.
#include "common.h"
#include "device_utils.cuh"

struct ScanSmem {
    uint16_t lut[128 * 128];            // folded, bank-swizzled UC8 magnitude table
    uint16_t mag[SCAN_NMAG + 8];        // magnitudes of the tile, index = tile coordinate x - x0
    uint32_t crc_tab[256];
    uint32_t bit_syn[112];
    uint32_t syn_hash[512];
    uint32_t pre_bits[SCAN_TILE / 32];  // pre-check result, one bit per position
    uint16_t q1[SCAN_Q1_CAP];           // positions (tile relative) that passed the pre-check, ascending
    uint8_t  q1_tried[SCAN_Q1_CAP];     // phases whose correlator reached the threshold
    uint8_t  q1_live[SCAN_Q1_CAP];      // phases with a filter-dependent score (a Rec exists)
    uint32_t items[SCAN_ITEM_CAP];      // q1 index << 3 | phase index, ascending (position, phase)
    Rec      recs[SCAN_FULL_CAP];       // live records in final order
    uint32_t scratch[40];
    uint32_t syn_mul;
    uint32_t rec_off, tile, overflow;
};

// Correlator `row` (= u % 5) on four consecutive magnitudes; demod_2400.c:74-93.
__device__ __forceinline__ int correlate(int row, int m0, int m1, int m2, int m3) {
    switch (row) {
        case 0: return 18 * m0 - 15 * m1 - 3 * m2;
        case 1: return 14 * m0 - 5 * m1 - 9 * m2;
        case 2: return 16 * m0 + 5 * m1 - 20 * m2;
        case 3: return 7 * m0 + 11 * m1 - 18 * m2;
        default: return 4 * m0 + 15 * m1 - 20 * m2 + m3;
    }
}

// Slice message bits [k0, k0+nb) of try_phase t for the preamble at `pa` (pointer to mag[p]),
// MSB first; closed form of slice_byte (demod_2400.c:133-213): u = t + 12k, sample 19 + u/5, row u%5.
__device__ __forceinline__ uint32_t slice_bits(const uint16_t *pa, int t, int k0, int nb) {
    uint32_t v = 0;
    int u = t + 12 * k0;
    int o = u / 5, r = u - 5 * o;
    const uint16_t *s = pa + 19 + o;
    for (int k = 0; k < nb; k++) {
        int c = correlate(r, s[0], s[1], s[2], s[3]);
        v = (v << 1) | (c > 0 ? 1u : 0u);
        r += 2; s += 2;                   // u += 12: two samples and two rows further...
        if (r >= 5) { r -= 5; s += 1; }   // ...with carry
    }
    return v;
}

__device__ __forceinline__ uint32_t msg_byte(const uint32_t w[4], int i) { return (w[i >> 2] >> (24 - 8 * (i & 3))) & 0xffu; }

__device__ __forceinline__ uint32_t crc24(const ScanSmem &S, const uint32_t w[4], int nbytes) {
    uint32_t rem = 0;
    for (int i = 0; i < nbytes - 3; i++) rem = ((rem << 8) ^ S.crc_tab[msg_byte(w, i) ^ ((rem >> 16) & 0xffu)]) & 0xffffffu;
    return rem ^ (msg_byte(w, nbytes - 3) << 16) ^ (msg_byte(w, nbytes - 2) << 8) ^ msg_byte(w, nbytes - 1);
}

// crc.c:383-406 for nfix_crc = 1: message bit (>= 5) whose single-bit syndrome equals `syn`, or -2.
__device__ __forceinline__ int diagnose1(const ScanSmem &S, uint32_t syn, int bits) {
    uint32_t e = S.syn_hash[(syn * S.syn_mul) >> 23];
    if ((e >> 8) != syn) return -2;
    int b = (int)(e & 0xffu) - (112 - bits);
    return b >= 5 ? b : -2;
}

// Filter-independent part of scoreModesMessage (mode_s.c:309-419) for a fully sliced frame.
// Returns RecKind, or 0 when the score is -2 whatever the filter holds.
__device__ __forceinline__ uint32_t classify(const ScanSmem &S, const ScanParams &P, const uint32_t w[4], int df, int nbytes,
                                             uint32_t *crc_out, uint32_t *addr_out, int *fixbit_out) {
    const uint32_t aa = w[0] & 0xffffffu;
    *fixbit_out = -1;
    if (nbytes == 14) {
        const uint32_t crc = crc24(S, w, 14);
        *crc_out = crc;
        if (P.fixdf && P.nfix && (df == 1 || df == 25 || df == 21 || df == 19 || df == 16)) {
            // fixDF17msgtype (mode_s.c:276-301): forcing DF=17 flips exactly one DF bit, so the repaired
            // frame is CRC-clean iff the syndrome equals that bit's single-bit syndrome.
            const int bit = __clz((uint32_t)(df ^ 17)) - 27;     // 16->0, 8->1, 4->2, 2->3, 1->4
            if (crc == S.bit_syn[bit]) { *addr_out = aa; *fixbit_out = bit; return K_DFREPAIR; }
        }
        if (df == 16 || df == 20 || df == 21) { *addr_out = crc; return K_AP; }
        if (df == 17 || df == 18) {
            if (crc == 0) { *addr_out = aa; return K_ES_OK; }
            if (!P.nfix) return 0;
            const int b = diagnose1(S, crc, 112);
            if (b < 0) return 0;
            *fixbit_out = b;
            *addr_out = (b >= 8 && b <= 31) ? (aa ^ (1u << (31 - b))) : aa;   // correct_aa_field, mode_s.c:230-245
            return K_ES_FIX;
        }
        return 0;   // DF1/19/25 without a repair: unknown message type
    }
    // short frames; all-zero check mode_s.c:336-338 (only DF0 can start with a zero byte)
    if (w[0] == 0 && (w[1] >> 8) == 0) return 0;
    const uint32_t crc = crc24(S, w, 7);
    *crc_out = crc;
    if (df == 11) {
        if (crc & 0xffff80u) {
            if (!P.nfix) return 0;
            const int b = diagnose1(S, crc, 56);
            if (b < 0) return 0;
            *fixbit_out = b;
            *addr_out = (b >= 8 && b <= 31) ? (aa ^ (1u << (31 - b))) : aa;
            return K_DF11_FIX;
        }
        *addr_out = aa;
        return (crc & 0x7fu) ? K_DF11_IID : K_DF11_IID0;
    }
    *addr_out = crc;   // DF0/4/5
    return K_AP;
}

// Phases 2b-5 for one tile: from the pre-check bitmap to the emitted lists.  The candidate queues live in
// shared memory (SLOW = false, capacities sized for real traffic) or, when a tile is denser than that, in this
// CTA's slice of a global scratch arena sized for the worst case (SLOW = true) — same code, never a dropped
// candidate.  Returns false (block-uniform) if a shared-memory queue overflowed; nothing has been written then.
template <bool SLOW>
__device__ __forceinline__ bool process_candidates(ScanSmem &S, const ScanParams &P, uint32_t tile, uint8_t *scratch) {
    const uint32_t tid = threadIdx.x;
    uint16_t *q1 = SLOW ? reinterpret_cast<uint16_t *>(scratch) : S.q1;
    uint8_t *q1_tried = SLOW ? scratch + 2 * SCAN_TILE : S.q1_tried;
    uint8_t *q1_live = SLOW ? scratch + 3 * SCAN_TILE : S.q1_live;
    uint32_t *items = SLOW ? reinterpret_cast<uint32_t *>(scratch + 4 * SCAN_TILE) : S.items;
    Rec *recs = SLOW ? reinterpret_cast<Rec *>(scratch + 24 * SCAN_TILE) : S.recs;
    const uint32_t q1_cap = SLOW ? SCAN_TILE : SCAN_Q1_CAP;
    const uint32_t item_cap = SLOW ? 5 * SCAN_TILE : SCAN_ITEM_CAP;
    const uint32_t rec_cap = SLOW ? 5 * SCAN_TILE : SCAN_FULL_CAP;

    if (tid == 0) S.overflow = 0;
    uint32_t n_q1;
    {   // ordered compaction of the pre-check bitmap into q1
        const uint32_t word = tid < SCAN_TILE / 32 ? S.pre_bits[tid] : 0;
        uint32_t off = block_excl_scan(__popc(word), S.scratch, &n_q1);
        if (n_q1 > q1_cap) return false;
        uint32_t wbits = word;
        while (wbits) { const uint32_t b = __ffs(wbits) - 1; wbits &= wbits - 1; q1[off++] = (uint16_t)(tid * 32 + b); }
    }
    __syncthreads();

    // ---- noise-relative thresholds, three correlations (demod_2400.c:330-378) -------------------------
    uint32_t n_items = 0;
    for (uint32_t r0 = 0; r0 < n_q1; r0 += SCAN_THREADS) {
        const uint32_t e = r0 + tid;
        uint32_t tried = 0;
        if (e < n_q1) {
            const uint16_t *pa = &S.mag[q1[e]];
            const int base_noise = pa[5] + pa[8] + pa[16] + pa[17] + pa[18];
            const int ref_level = (base_noise * P.thr) >> 5;
            const int d23 = (int)pa[2] - (int)pa[3], s14 = pa[1] + pa[4], d1011 = (int)pa[10] - (int)pa[11];
            const int common = s14 - d23 + pa[9] + pa[12];
            if (common - d1011 >= ref_level) tried |= 0x03;                        // try_phase 4, 5
            if (common + d1011 >= ref_level) tried |= 0x0c;                        // try_phase 6, 7
            if (s14 + 2 * d23 + d1011 + pa[12] >= ref_level) tried |= 0x10;        // try_phase 8
            q1_tried[e] = (uint8_t)tried;
            q1_live[e] = 0;
        }
        uint32_t total;
        uint32_t off = n_items + block_excl_scan(__popc(tried), S.scratch, &total);
        if (n_items + total > item_cap) { if (tid == 0) S.overflow = 1; }
        else { uint32_t tb = tried; while (tb) { const uint32_t ph = __ffs(tb) - 1; tb &= tb - 1; items[off++] = (e << 3) | ph; } }
        n_items += total;
    }
    __syncthreads();
    if (S.overflow) return false;

    // ---- slice, CRC, classify each (position, phase) (demod_2400.c:215-258) ---------------------------
    uint32_t n_recs = 0;
    for (uint32_t r0 = 0; r0 < n_items; r0 += SCAN_THREADS) {
        const uint32_t i = r0 + tid;
        uint32_t kind = 0, crc = 0, addr = 0, w[4] = {0, 0, 0, 0}, item = 0;
        int fixbit = -1;
        if (i < n_items) {
            item = items[i];
            const uint16_t *pa = &S.mag[q1[item >> 3]];
            const int t = 4 + (int)(item & 7);
            const uint32_t b0 = slice_bits(pa, t, 0, 8);
            const int df = (int)(b0 >> 3);
            const int nbytes = ((P.long_set >> df) & 1) ? 14 : ((P.short_set >> df) & 1) ? 7 : 0;
            if (nbytes) {
                w[0] = (b0 << 24) | slice_bits(pa, t, 8, 24);
                if (nbytes == 7) w[1] = slice_bits(pa, t, 32, 24) << 8;
                else { w[1] = slice_bits(pa, t, 32, 32); w[2] = slice_bits(pa, t, 64, 32); w[3] = slice_bits(pa, t, 96, 16) << 16; }
                kind = classify(S, P, w, df, nbytes, &crc, &addr, &fixbit);
            }
        }
        uint32_t total;
        const uint32_t r = n_recs + block_excl_scan(kind ? 1u : 0u, S.scratch, &total);
        if (n_recs + total > rec_cap) { if (tid == 0) S.overflow = 1; }
        else if (kind) {
            uint32_t *rw = reinterpret_cast<uint32_t *>(&recs[r]);
            // bytes 0..13 = message, byte 14 = kind, byte 15 = fixbit (little-endian words, big-endian message)
            rw[0] = __byte_perm(w[0], 0, 0x0123); rw[1] = __byte_perm(w[1], 0, 0x0123); rw[2] = __byte_perm(w[2], 0, 0x0123);
            rw[3] = (__byte_perm(w[3], 0, 0x0123) & 0xffffu) | (kind << 16) | ((uint32_t)(fixbit & 0xff) << 24);
            rw[4] = crc; rw[5] = addr; rw[6] = 0; rw[7] = 0;
            const uint32_t e = item >> 3;
            atomicOr(reinterpret_cast<uint32_t *>(&q1_live[e & ~3u]), (1u << (item & 7)) << (8 * (e & 3)));
        }
        n_recs += total;
    }
    __syncthreads();
    if (S.overflow) return false;

    // ---- ordered emission ---------------------------------------------------------------------------------
    if (tid == 0) {
        const uint32_t off = n_recs ? atomicAdd(&P.ctl->rec_alloc, n_recs) : 0;
        if (off + n_recs > P.ctl->rec_cap) { atomicOr(&P.ctl->overflow, 1u); S.overflow = 1; }   // host regrows the pool and reruns
        S.rec_off = off;
    }
    uint32_t n_pos = 0;
    PosEntry *pos_out = P.pos_pool + (size_t)tile * SCAN_TILE;
    for (uint32_t r0 = 0; r0 < n_q1; r0 += SCAN_THREADS) {
        const uint32_t e = r0 + tid;
        const uint32_t tried = e < n_q1 ? q1_tried[e] : 0;
        uint32_t total;
        const uint32_t r = n_pos + block_excl_scan(tried ? 1u : 0u, S.scratch, &total);
        if (tried) pos_out[r] = (uint32_t)q1[e] | (tried << 16) | ((uint32_t)q1_live[e] << 21);
        n_pos += total;
    }
    __syncthreads();   // S.rec_off / S.overflow visible
    const bool pool_ok = !S.overflow;
    if (pool_ok) {
        const uint4 *src = reinterpret_cast<const uint4 *>(recs);
        uint4 *dst = reinterpret_cast<uint4 *>(P.rec_pool + S.rec_off);
        for (uint32_t i = tid; i < n_recs * 2; i += SCAN_THREADS) dst[i] = src[i];
    }
    if (tid == 0) { TileOut t; t.n_pos = n_pos; t.n_rec = pool_ok ? n_recs : 0; t.rec_off = S.rec_off; t.pad_ = 0; P.tile_out[tile] = t; }
    return true;
}

__global__ void __launch_bounds__(SCAN_THREADS, 2) scan_kernel(const ScanParams P, const DeviceTables *__restrict__ tables) {
    extern __shared__ uint4 smem_raw[];
    ScanSmem &S = *reinterpret_cast<ScanSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31;

    // one-time table staging (persistent CTA)
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(tables->lut_fold);
        uint4 *dst = reinterpret_cast<uint4 *>(S.lut);
        for (uint32_t i = tid; i < sizeof(S.lut) / 16; i += SCAN_THREADS) dst[i] = src[i];
        for (uint32_t i = tid; i < 256; i += SCAN_THREADS) S.crc_tab[i] = tables->crc_tab[i];
        for (uint32_t i = tid; i < 112; i += SCAN_THREADS) S.bit_syn[i] = tables->bit_syn[i];
        for (uint32_t i = tid; i < 512; i += SCAN_THREADS) S.syn_hash[i] = tables->syn_hash[i];
        if (tid == 0) S.syn_mul = tables->syn_hash_mul;
    }

    for (;;) {
        __syncthreads();
        if (tid == 0) S.tile = atomicAdd(&P.ctl->tile_counter, 1u);
        __syncthreads();
        const uint32_t tile = S.tile;
        if (tile >= P.n_tiles) break;

        const Segment seg = P.segs[P.tile_seg[tile]];
        const uint32_t x0 = (tile - seg.tile_begin) * SCAN_TILE;         // tile origin in tile coordinates
        // byte address of tile coordinate x: seg.base + 2*(x - lead); x0 multiple of 8 => 16B aligned
        const uint8_t *tile_base = seg.base + 2 * ((int64_t)x0 - (int64_t)seg.lead);
        const uint32_t x_data_end = seg.lead + seg.npos + B200_TRAIL;    // first x without data
        const uint32_t x_zero_end = (seg.flags & SEG_HALO_ZERO) ? seg.lead + B200_TRAIL : seg.lead; // x below: magnitude 0, memory not read
        const bool is_mag = seg.flags & SEG_MAG;
        const bool last_tile = tile + 1 == seg.tile_begin + seg.n_tiles;

        // Power statistics (convert.c:75-79) are per reference buffer: new sample n = x - lead - 326 belongs to
        // buffer n / buf_len and is counted by the tile whose position range holds x (the last tile also owns
        // the tail).  A tile touches at most two buffers when buf_len >= SCAN_NMAG; other layouts take the slow branch.
        const int64_t n_first = (int64_t)x0 - seg.lead - B200_TRAIL;    // new-sample index of tile coordinate x0
        const uint32_t nb0 = n_first > 0 ? (uint32_t)n_first / seg.buf_len : 0;
        const int64_t bound1 = (int64_t)(nb0 + 1) * seg.buf_len;         // first new sample of buffer nb0 + 1
        const int64_t bound2 = bound1 + seg.buf_len;

        // ---- phase 1: load + convert ----------------------------------------------------------------
        unsigned long long acc_level = 0, acc_power = 0;
        uint32_t acc_buf = 0xffffffffu;
        for (uint32_t c = tid; c < SCAN_NMAG / 8; c += SCAN_THREADS) {
            const uint32_t xc = x0 + c * 8;
            uint32_t m[8];
            if (xc + 8 <= x_zero_end || xc >= x_data_end) {
#pragma unroll
                for (int i = 0; i < 8; i++) m[i] = 0;
            } else {
                const uint4 raw = ldg_stream_u4(tile_base + (size_t)c * 16);
                const uint32_t wv[4] = {raw.x, raw.y, raw.z, raw.w};
                if (is_mag) {
#pragma unroll
                    for (int i = 0; i < 4; i++) { m[2 * i] = wv[i] & 0xffffu; m[2 * i + 1] = wv[i] >> 16; }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint32_t w = wv[i];                               // bytes I0 Q0 I1 Q1
                        const uint32_t sgn = prmt(w, 0, 0xba98);                // 0xff where the byte is >= 128
                        const uint32_t f = (w ^ ~sgn) & 0x7f7f7f7fu;            // fold: v>=128 ? v-128 : 127-v
                        uint32_t off = f + (f & 0x007f007fu);                   // per half: fq*256 + 2*fi
                        off ^= (f >> 5) & 0x00780078u;                          // bank swizzle (see modes_tables.h)
                        m[2 * i] = *reinterpret_cast<const uint16_t *>(reinterpret_cast<const uint8_t *>(S.lut) + (off & 0xffffu));
                        m[2 * i + 1] = *reinterpret_cast<const uint16_t *>(reinterpret_cast<const uint8_t *>(S.lut) + (off >> 16));
                    }
                }
                if (xc < x_zero_end || xc + 8 > x_data_end) {   // boundary chunk: mask the samples that are not data
#pragma unroll
                    for (int i = 0; i < 8; i++) if (xc + i < x_zero_end || xc + i >= x_data_end) m[i] = 0;
                }
                if (c < SCAN_TILE / 8 || last_tile) {
                    const int64_t n0 = n_first + (int64_t)c * 8;
                    uint32_t b = 0xffffffffu;
                    if (n0 >= 0 && n0 + 8 <= (int64_t)seg.npos) {
                        if (n0 + 8 <= bound1) b = nb0; else if (n0 >= bound1 && n0 + 8 <= bound2) b = nb0 + 1;
                    }
                    if (b != 0xffffffffu) {           // all eight samples belong to buffer b
                        b += seg.first_buf;
                        if (b != acc_buf) {
                            if (acc_buf != 0xffffffffu) { atomicAdd(&P.buf_acc[acc_buf].sum_level, acc_level); atomicAdd(&P.buf_acc[acc_buf].sum_power, acc_power); }
                            acc_buf = b; acc_level = 0; acc_power = 0;
                        }
                        acc_level += (m[0] + m[1]) + (m[2] + m[3]) + (m[4] + m[5]) + (m[6] + m[7]);
#pragma unroll
                        for (int i = 0; i < 8; i++) mad_wide(acc_power, m[i], m[i]);
                    } else if (n0 + 8 > 0 && n0 < (int64_t)seg.npos) {   // chunk straddles a boundary: per sample
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int64_t n = n0 + i;
                            if (n >= 0 && n < (int64_t)seg.npos) {
                                const uint32_t bb = seg.first_buf + (uint32_t)n / seg.buf_len;
                                if (bb != acc_buf) {
                                    if (acc_buf != 0xffffffffu) { atomicAdd(&P.buf_acc[acc_buf].sum_level, acc_level); atomicAdd(&P.buf_acc[acc_buf].sum_power, acc_power); }
                                    acc_buf = bb; acc_level = 0; acc_power = 0;
                                }
                                acc_level += m[i];
                                mad_wide(acc_power, m[i], m[i]);
                            }
                        }
                    }
                }
            }
            uint4 packed;
            packed.x = m[0] | (m[1] << 16); packed.y = m[2] | (m[3] << 16);
            packed.z = m[4] | (m[5] << 16); packed.w = m[6] | (m[7] << 16);
            *reinterpret_cast<uint4 *>(&S.mag[c * 8]) = packed;
        }
        {   // flush the statistics: one atomic pair per warp when the whole warp fed the same buffer
            const uint32_t b0 = __shfl_sync(FULLMASK, acc_buf, 0);
            if (__all_sync(FULLMASK, acc_buf == b0)) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { acc_level += __shfl_xor_sync(FULLMASK, acc_level, o); acc_power += __shfl_xor_sync(FULLMASK, acc_power, o); }
                if (lane == 0 && b0 != 0xffffffffu) { atomicAdd(&P.buf_acc[b0].sum_level, acc_level); atomicAdd(&P.buf_acc[b0].sum_power, acc_power); }
            } else if (acc_buf != 0xffffffffu) {
                atomicAdd(&P.buf_acc[acc_buf].sum_level, acc_level); atomicAdd(&P.buf_acc[acc_buf].sum_power, acc_power);
            }
        }
        __syncthreads();

        // ---- phase 2: pre-check every position (demod_2400.c:311-320) ------------------------------------
        const uint32_t p_lo = seg.lead > x0 ? seg.lead - x0 : 0;                              // first real position
        const uint32_t p_hi = min((uint32_t)SCAN_TILE, seg.lead + seg.npos > x0 ? seg.lead + seg.npos - x0 : 0u);
        for (uint32_t it = 0; it < SCAN_TILE / SCAN_THREADS; it++) {
            const uint32_t p = it * SCAN_THREADS + tid;
            const uint16_t *pa = &S.mag[p];
            const bool ok = pa[1] > pa[7] && pa[12] > pa[14] && pa[12] > pa[15] && p >= p_lo && p < p_hi;
            const uint32_t bal = __ballot_sync(FULLMASK, ok);
            if (lane == 0) S.pre_bits[p >> 5] = bal;
        }
        __syncthreads();

        // ---- phases 2b-5 ----------------------------------------------------------------------------------------
        if (!process_candidates<false>(S, P, tile, nullptr)) {
            __syncthreads();
            if (P.scratch) {
                process_candidates<true>(S, P, tile, P.scratch + (size_t)blockIdx.x * SCAN_SCRATCH_BYTES);
            } else if (tid == 0) {      // no scratch arena yet: tell the host, which allocates one and reruns
                atomicOr(&P.ctl->overflow, 2u);
                TileOut t = {0, 0, 0, 0};
                P.tile_out[tile] = t;
            }
        }
    }
}

extern "C" int b200_scan_grid(int n_sm) {
    static int per_sm = 0;
    if (!per_sm) {
        if (cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanSmem)) != cudaSuccess) return -1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_kernel, SCAN_THREADS, sizeof(ScanSmem)) != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    return n_sm * per_sm;
}

extern "C" int b200_launch_scan(const ScanParams *p, const DeviceTables *d_tables, int n_sm, void *stream) {
    int grid = b200_scan_grid(n_sm);
    if (grid < 0) return (int)cudaGetLastError();
    if ((uint32_t)grid > p->n_tiles) grid = (int)p->n_tiles;
    if (grid == 0) return 0;
    scan_kernel<<<grid, SCAN_THREADS, sizeof(ScanSmem), (cudaStream_t)stream>>>(*p, d_tables);
    return (int)cudaGetLastError();
}

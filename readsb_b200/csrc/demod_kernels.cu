// demod_kernels.cu — hand-written sm_100a kernels of the Mode-S demodulator pipeline.
//
//   scan_kernel     (stage A) uc8 IQ -> magnitude (shared-memory folded LUT, coalesced 16-byte HBM
//                   loads), sliding preamble pre-check + three correlator thresholds, 5-phase
//                   PPM/Manchester bit slicer for 56/112-bit frames, CRC-24 syndrome, DF17 repair and
//                   single-bit-fix classification.  Stateless per position; persistent CTAs pull
//                   tiles of SCAN_TILE positions from a dynamic counter.
//   resolve_kernel  (stage B) one warp per receiver: the sequential part of demodulate2400()
//                   (ICAO-filter dependent scoring, best-phase pick, accept, skip-ahead) walked
//                   speculatively 32 candidates at a time, filter tables in shared memory.
//   finalize_kernel one warp per accepted frame: signal power, per-buffer / per-receiver power
//                   statistics, packing of the frame list for the single D2H copy.
//
// Behavioural references (reference tree): convert.c:35-108, demod_2400.c:74-482, crc.c:42-418,
// mode_s.c:230-419 and :443-596,:766-779, icao_filter.c:96-154, readsb.c:1227-1231.
// No tensor cores: this is an integer scan/correlate path bounded by HBM reads and instruction issue.
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.h"

#define WARP 32
#define FULLMASK 0xffffffffu

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// prmt.b32 in its default mode: selector nibble bit 3 replicates the sign of the selected byte
// (the __byte_perm intrinsic masks that bit away).
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t lane, uint32_t *total) {
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < WARP; o <<= 1) {
        uint32_t y = __shfl_up_sync(FULLMASK, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    *total = __shfl_sync(FULLMASK, x, WARP - 1);
    return x - v;
}

// Exclusive prefix sum over the block (all threads must call). scratch: >= 33 uint32 of shared memory.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *scratch, uint32_t *total) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t wtot;
    uint32_t ex = warp_excl_scan(v, lane, &wtot);
    __syncthreads();                       // protect scratch reuse from a previous call
    if (lane == 0) scratch[wid] = wtot;
    __syncthreads();
    if (wid == 0) {
        uint32_t t = lane < nw ? scratch[lane] : 0, tt;
        uint32_t e = warp_excl_scan(t, lane, &tt);
        if (lane < nw) scratch[lane] = e;
        if (lane == 0) scratch[32] = tt;
    }
    __syncthreads();
    *total = scratch[32];
    return ex + scratch[wid];
}

// ------------------------------------------------------------------------------------------------
// stage A
// ------------------------------------------------------------------------------------------------
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
    uint32_t n_q1, n_items, n_recs, n_pos, rec_off, tile, overflow;
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

__global__ void __launch_bounds__(SCAN_THREADS, 2) scan_kernel(const ScanParams P, const DeviceTables *__restrict__ tables) {
    extern __shared__ uint4 smem_raw[];
    ScanSmem &S = *reinterpret_cast<ScanSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

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
        if (tid == 0) { S.tile = atomicAdd(&P.ctl->tile_counter, 1u); S.overflow = 0; }
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

        // ---- phase 1: load + convert ----------------------------------------------------------------
        // Each chunk = 8 samples = 16 bytes.  Per-buffer level/power sums are exact integers (convert.c:75-79).
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
                        const uint32_t sgn = prmt(w, 0, 0xba98);         // 0xff where the byte is >= 128
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
                // ownership of the power statistics: new sample n = x - lead - 326 belongs to buffer n / buf_len,
                // and is counted by the tile whose position range contains x (the last tile also owns the tail).
                if (c < SCAN_TILE / 8 || tile + 1 == seg.tile_begin + seg.n_tiles) {
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int64_t n = (int64_t)xc + i - seg.lead - B200_TRAIL;
                        if (n >= 0 && n < (int64_t)seg.npos) {
                            const uint32_t b = seg.first_buf + (uint32_t)n / seg.buf_len;
                            if (b != acc_buf) {
                                if (acc_buf != 0xffffffffu) { atomicAdd(&P.buf_acc[acc_buf].sum_level, acc_level); atomicAdd(&P.buf_acc[acc_buf].sum_power, acc_power); }
                                acc_buf = b; acc_level = 0; acc_power = 0;
                            }
                            acc_level += m[i];
                            acc_power += (unsigned long long)(m[i] * m[i]);
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
        if (tid == 0) { S.n_q1 = 0; S.n_items = 0; S.n_recs = 0; S.n_pos = 0; }
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
        {   // ordered compaction of the set bits into q1
            const uint32_t word = tid < SCAN_TILE / 32 ? S.pre_bits[tid] : 0;
            uint32_t total;
            uint32_t off = block_excl_scan(__popc(word), S.scratch, &total);
            if (total > SCAN_Q1_CAP) { if (tid == 0) S.overflow = 1; total = 0; }
            else { uint32_t wbits = word; while (wbits) { const uint32_t b = __ffs(wbits) - 1; wbits &= wbits - 1; S.q1[off++] = (uint16_t)(tid * 32 + b); } }
            if (tid == 0) S.n_q1 = total;
        }
        __syncthreads();
        const uint32_t n_q1 = S.n_q1;

        // ---- phase 3: noise-relative thresholds, three correlations (demod_2400.c:330-378) -------------
        uint32_t item_base = 0;
        for (uint32_t r0 = 0; r0 < n_q1; r0 += SCAN_THREADS) {
            const uint32_t e = r0 + tid;
            uint32_t tried = 0;
            if (e < n_q1) {
                const uint16_t *pa = &S.mag[S.q1[e]];
                const int base_noise = pa[5] + pa[8] + pa[16] + pa[17] + pa[18];
                const int ref_level = (base_noise * P.thr) >> 5;
                const int d23 = (int)pa[2] - (int)pa[3], s14 = pa[1] + pa[4], d1011 = (int)pa[10] - (int)pa[11];
                const int common = s14 - d23 + pa[9] + pa[12];
                if (common - d1011 >= ref_level) tried |= 0x03;                        // try_phase 4, 5
                if (common + d1011 >= ref_level) tried |= 0x0c;                        // try_phase 6, 7
                if (s14 + 2 * d23 + d1011 + pa[12] >= ref_level) tried |= 0x10;        // try_phase 8
                S.q1_tried[e] = (uint8_t)tried;
                S.q1_live[e] = 0;
            }
            uint32_t total;
            uint32_t off = item_base + block_excl_scan(__popc(tried), S.scratch, &total);
            if (item_base + total > SCAN_ITEM_CAP) { if (tid == 0) S.overflow = 1; }
            else { uint32_t tb = tried; while (tb) { const uint32_t ph = __ffs(tb) - 1; tb &= tb - 1; S.items[off++] = (e << 3) | ph; } }
            item_base += total;
        }
        __syncthreads();
        const uint32_t n_items = S.overflow ? 0 : item_base;

        // ---- phase 4: slice, CRC, classify each (position, phase) (demod_2400.c:215-258) ----------------
        uint32_t rec_base = 0;
        for (uint32_t r0 = 0; r0 < n_items; r0 += SCAN_THREADS) {
            const uint32_t i = r0 + tid;
            uint32_t kind = 0, crc = 0, addr = 0, w[4] = {0, 0, 0, 0}, item = 0;
            int fixbit = -1;
            if (i < n_items) {
                item = S.items[i];
                const uint16_t *pa = &S.mag[S.q1[item >> 3]];
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
            const uint32_t r = rec_base + block_excl_scan(kind ? 1u : 0u, S.scratch, &total);
            if (rec_base + total > SCAN_FULL_CAP) { if (tid == 0) S.overflow = 1; }
            else if (kind) {
                Rec &R = S.recs[r];
                uint32_t *rw = reinterpret_cast<uint32_t *>(&R);
                // bytes 0..13 = message, byte 14 = kind, byte 15 = fixbit (little-endian words, big-endian message)
                rw[0] = __byte_perm(w[0], 0, 0x0123); rw[1] = __byte_perm(w[1], 0, 0x0123); rw[2] = __byte_perm(w[2], 0, 0x0123);
                rw[3] = (__byte_perm(w[3], 0, 0x0123) & 0xffffu) | (kind << 16) | ((uint32_t)(fixbit & 0xff) << 24);
                rw[4] = crc; rw[5] = addr; rw[6] = 0; rw[7] = 0;
                const uint32_t e = item >> 3;
                atomicOr(reinterpret_cast<uint32_t *>(&S.q1_live[e & ~3u]), (1u << (item & 7)) << (8 * (e & 3)));
            }
            rec_base += total;
        }
        __syncthreads();
        if (S.overflow) {       // per-tile capacity exceeded: fail loudly, never drop silently
            if (tid == 0) { atomicOr(&P.ctl->overflow, 2u); P.tile_out[tile].n_pos = 0; P.tile_out[tile].n_rec = 0; P.tile_out[tile].rec_off = 0; }
            continue;
        }
        const uint32_t n_recs = rec_base;

        // ---- phase 5: ordered emission ---------------------------------------------------------------------
        if (tid == 0) {
            uint32_t off = n_recs ? atomicAdd(&P.ctl->rec_alloc, n_recs) : 0;
            if (off + n_recs > P.ctl->rec_cap) { atomicOr(&P.ctl->overflow, 1u); S.overflow = 1; }
            S.rec_off = off;
        }
        uint32_t pos_base = 0;
        PosEntry *pos_out = P.pos_pool + (size_t)tile * SCAN_TILE;
        for (uint32_t r0 = 0; r0 < n_q1; r0 += SCAN_THREADS) {
            const uint32_t e = r0 + tid;
            const uint32_t tried = e < n_q1 ? S.q1_tried[e] : 0;
            uint32_t total;
            const uint32_t r = pos_base + block_excl_scan(tried ? 1u : 0u, S.scratch, &total);
            if (tried) pos_out[r] = (uint32_t)S.q1[e] | (tried << 16) | ((uint32_t)S.q1_live[e] << 21);
            pos_base += total;
        }
        __syncthreads();   // S.rec_off / S.overflow visible
        if (!S.overflow) {
            const uint4 *src = reinterpret_cast<const uint4 *>(S.recs);
            uint4 *dst = reinterpret_cast<uint4 *>(P.rec_pool + S.rec_off);
            for (uint32_t i = tid; i < n_recs * 2; i += SCAN_THREADS) dst[i] = src[i];
        }
        if (tid == 0) { TileOut t; t.n_pos = pos_base; t.n_rec = S.overflow ? 0 : n_recs; t.rec_off = S.rec_off; t.pad_ = 0; P.tile_out[tile] = t; }
    }
}

// ------------------------------------------------------------------------------------------------
// stage B
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t icao_slot(uint32_t a) { return (a * 0x9E3779B1u) >> (32 - ICAO_CAP_LOG2); }

__device__ __forceinline__ bool gen_has(const uint32_t *g, uint32_t a) {
    uint32_t h = icao_slot(a);
    for (;;) {
        const uint32_t v = g[h];
        if (v == a) return true;
        if (v == ICAO_EMPTY) return false;
        h = (h + 1) & (ICAO_CAP - 1);
    }
}

__device__ __forceinline__ bool gen_add(uint32_t *g, uint32_t *count, uint32_t a) {   // false when full
    uint32_t h = icao_slot(a);
    for (;;) {
        const uint32_t v = g[h];
        if (v == a) return true;
        if (v == ICAO_EMPTY) break;
        h = (h + 1) & (ICAO_CAP - 1);
    }
    if (*count >= ICAO_CAP / 2) return false;
    g[h] = a; (*count)++;
    return true;
}

struct ResolveSmem {
    uint32_t gen[2][ICAO_CAP];
};

// Score of one record under the current filter (mode_s.c:309-419).
__device__ __forceinline__ int rec_score(uint32_t kind, bool known) {
    switch (kind) {
        case K_AP: return known ? 1000 : -1;
        case K_DFREPAIR: return known ? 900 : 700;
        case K_DF11_FIX: return known ? 800 : -1;
        case K_DF11_IID0: return known ? 1600 : 750;
        case K_DF11_IID: return known ? 1000 : -1;
        case K_ES_OK: return known ? 1800 : 1400;
        case K_ES_FIX: return known ? 900 : 700;
        default: return -2;
    }
}

__global__ void __launch_bounds__(32) resolve_kernel(const ResolveParams P) {
    __shared__ ResolveSmem S;
    const uint32_t stream = blockIdx.x, lane = threadIdx.x;
    StreamState *st = &P.state[stream];
    if (P.ctl->overflow & 3u) return;     // stage A failed: leave every receiver's state untouched, the host redoes the run

    for (uint32_t i = lane; i < 2 * ICAO_CAP; i += 32) (&S.gen[0][0])[i] = (&st->gen[0][0])[i];
    uint32_t gcount[2] = {st->gen_count[0], st->gen_count[1]};
    uint32_t active = st->active, armed = st->flip_armed, seq = st->buffer_seq, err = st->error;
    int64_t next_flip = st->next_flip_ms;
    bool dirty[2] = {false, false};
    __syncwarp();

    // per-lane partial counters, reduced at the end
    uint32_t c_pre = 0, c_bad = 0, c_unk = 0, c_acc0 = 0, c_acc1 = 0, c_tp[5] = {0, 0, 0, 0, 0}, c_bp[5] = {0, 0, 0, 0, 0};
    unsigned long long c_samples = 0;
    uint32_t c_bufs = 0, c_flips = 0;
    uint32_t nframes = 0;
    b200_frame *fout = P.frames + (size_t)stream * P.frame_cap;

    for (uint32_t si = P.stream_seg_begin[stream]; si < P.stream_seg_begin[stream + 1]; si++) {
        const Segment seg = P.segs[si];
        uint32_t tile = seg.tile_begin, idx = 0;
        const uint32_t tile_end = seg.tile_begin + seg.n_tiles;
        TileOut to = {0, 0, 0, 0};
        if (seg.n_tiles) to = P.tile_out[tile];
        uint32_t rec_cursor = to.rec_off;

        for (uint32_t b = 0; b < seg.n_bufs; b++) {
            const uint32_t d_begin = b * seg.buf_len;
            const uint32_t d_end = min(d_begin + seg.buf_len, seg.npos);
            const int64_t buf_ts = seg.first_ts + (int64_t)d_begin * 5;
            int64_t now_ms = buf_ts / 12000;          // demod_2400.c:283-285
            uint32_t skip_until = d_begin;            // data-index form of the reference's `pa` skip
            uint32_t nfr_buf = 0;

            for (;;) {
                while (idx >= to.n_pos && tile + 1 < tile_end) { tile++; idx = 0; to = P.tile_out[tile]; rec_cursor = to.rec_off; }
                if (idx >= to.n_pos) break;
                const uint32_t x0 = (tile - seg.tile_begin) * SCAN_TILE;
                const bool has = idx + lane < to.n_pos;
                const PosEntry pe = has ? P.pos_pool[(size_t)tile * SCAN_TILE + idx + lane] : 0;
                const uint32_t d = x0 + (pe & 0x1fffu) - seg.lead;          // data index = position in the segment
                const bool inbuf = has && d < d_end;                        // entries are ascending: a prefix of lanes
                const uint32_t n_in = __popc(__ballot_sync(FULLMASK, inbuf));
                if (n_in == 0) break;                                       // next entry belongs to the next buffer
                const uint32_t tried = (pe >> 16) & 31u, live = (pe >> 21) & 31u;
                const uint32_t nlive = inbuf ? __popc(live) : 0;
                uint32_t dummy;
                const uint32_t rprefix = warp_excl_scan(nlive, lane, &dummy);
                const bool valid = inbuf && d >= skip_until;

                // score every tried phase with the current filter; first strictly greatest wins (demod_2400.c:243)
                int best = -2; uint32_t best_rec = 0, best_phase = 0, best_kind = 0; bool best_known = false; int best_fix = -1;
                if (valid && live) {
                    uint32_t lb = live, k = 0;
                    while (lb) {
                        const uint32_t ph = __ffs(lb) - 1; lb &= lb - 1;
                        const uint32_t ri = rec_cursor + rprefix + k; k++;
                        const uint32_t *rw = reinterpret_cast<const uint32_t *>(&P.rec_pool[ri]);
                        const uint32_t meta = rw[3], addr = rw[5];
                        const uint32_t kind = (meta >> 16) & 0xffu;
                        const bool known = gen_has(S.gen[0], addr) || gen_has(S.gen[1], addr);
                        const int sc = rec_score(kind, known);
                        if (sc > best) { best = sc; best_rec = ri; best_phase = ph; best_kind = kind; best_known = known; best_fix = (int)(int8_t)(meta >> 24); }
                    }
                }
                // accept test of decodeModesMessage (mode_s.c:443-596): only a corrected AA that is unknown rejects
                const bool decode_ok = best >= 0 && !(best_kind == K_ES_FIX && best_fix >= 8 && best_fix <= 31 && !best_known);
                const uint32_t acc_mask = __ballot_sync(FULLMASK, valid && decode_ok);
                const uint32_t f = acc_mask ? (uint32_t)__ffs(acc_mask) - 1 : 32u;
                const uint32_t consumed = acc_mask ? f + 1 : n_in;

                if (valid && lane < f) {       // rejected preambles before the first accepted one
                    c_pre++;
#pragma unroll
                    for (int p = 0; p < 5; p++) c_tp[p] += (tried >> p) & 1u;
                    if (best == -2) c_bad++; else c_unk++;       // -1, or decode result -1
                }
                if (acc_mask) {
                    uint32_t msglen = 0;
                    if (lane == f) {
                        c_pre++;
#pragma unroll
                        for (int p = 0; p < 5; p++) c_tp[p] += (tried >> p) & 1u;
                        const uint4 r0 = reinterpret_cast<const uint4 *>(&P.rec_pool[best_rec])[0];
                        const uint4 r1 = reinterpret_cast<const uint4 *>(&P.rec_pool[best_rec])[1];
                        uint8_t msg[16];
                        *reinterpret_cast<uint4 *>(msg) = r0;
                        const uint32_t crc_raw = r1.x;
                        const uint32_t df_raw = msg[0] >> 3;
                        msglen = (df_raw & 0x10) ? 112 : 56;                 // demod_2400.c:399 (DF as sliced)
                        uint32_t msgtype = df_raw, corrected = 0, crc = crc_raw;
                        int fix_bit = -1;
                        bool add = false;
                        if (best_kind == K_DFREPAIR) { msg[0] = (uint8_t)((msg[0] & 7) | (17 << 3)); msgtype = 17; corrected = 1; fix_bit = best_fix; crc = 0; }
                        else if (best_kind == K_DF11_FIX || best_kind == K_ES_FIX) { corrected = 1; fix_bit = best_fix; msg[fix_bit >> 3] ^= (uint8_t)(1u << (7 - (fix_bit & 7))); }
                        else if (best_kind == K_DF11_IID0 || (best_kind == K_ES_OK && msgtype == 17)) add = true;   // mode_s.c:766-779
                        const uint32_t msgbits = (msgtype & 0x10) ? 112 : 56;
                        const uint32_t aa = ((uint32_t)msg[1] << 16) | ((uint32_t)msg[2] << 8) | msg[3];
                        const uint32_t addr = best_kind == K_AP ? crc : aa;
                        const uint32_t j = d - d_begin;
                        const int64_t ts = buf_ts + (int64_t)j * 5 + (8 + 56) * 12 + (4 + best_phase);   // demod_2400.c:406
                        if (nframes < P.frame_cap) {
                            b200_frame fr;
                            fr.timestamp = ts; fr.sigpow_sum = 0; fr.j = j; fr.crc = crc; fr.addr = addr; fr.score = best;
                            fr.buffer_seq = seq; fr.signal_len = (uint16_t)(msglen * 12 / 5); fr.phase = (uint8_t)(4 + best_phase);
                            fr.msgtype = (uint8_t)msgtype; fr.msgbits = (uint8_t)msgbits; fr.correctedbits = (uint8_t)corrected;
                            fr.fix_bit = (int8_t)fix_bit; fr.flags = add ? B200_FRAME_ICAO_ADDED : 0;
#pragma unroll
                            for (int i = 0; i < 14; i++) fr.msg[i] = (uint32_t)i < msgbits / 8 ? msg[i] : 0;
                            // pad_: segment index and data index for finalize_kernel (cleared there)
                            fr.pad_[0] = 0; fr.pad_[1] = 0;
                            *reinterpret_cast<uint16_t *>(&fr.pad_[0]) = (uint16_t)(si & 0xffffu);
                            *reinterpret_cast<uint32_t *>(&fr.pad_[2]) = d;
                            // pad_[0..1] hold only 16 bits of the segment index; the upper bits ride in flags' spare bits
                            fout[nframes] = fr;
                        } else atomicOr(&P.ctl->overflow, 4u);
                        if (corrected) c_acc1++; else c_acc0++;
                        c_bp[best_phase]++;
                        if (add) { if (!gen_add(S.gen[active], &gcount[active], addr)) err = 1; dirty[active] = true; }
                        now_ms = buf_ts / 12000 + (ts - buf_ts) / 12000;      // demod_2400.c:409-414
                    }
                    __syncwarp();
                    // broadcast the state the accepting lane changed
                    msglen = __shfl_sync(FULLMASK, msglen, f);
                    now_ms = __shfl_sync(FULLMASK, now_ms, f);
                    gcount[0] = __shfl_sync(FULLMASK, gcount[0], f); gcount[1] = __shfl_sync(FULLMASK, gcount[1], f);
                    dirty[0] = __shfl_sync(FULLMASK, (int)dirty[0], f); dirty[1] = __shfl_sync(FULLMASK, (int)dirty[1], f);
                    err = __shfl_sync(FULLMASK, err, f);
                    const uint32_t d_f = __shfl_sync(FULLMASK, d, f);
                    skip_until = d_f + msglen * 2 + 1;                          // demod_2400.c:468 + loop increment
                    nframes++; nfr_buf++;
                }
                // advance the cursors past the consumed entries
                const uint32_t last = consumed - 1;
                rec_cursor += __shfl_sync(FULLMASK, rprefix + nlive, last);
                idx += consumed;
            }

            // end of buffer: readsb.c:876, then backgroundTasks' filter flip (readsb.c:1227-1231)
            c_samples += d_end - d_begin; c_bufs++;
            uint32_t flipped = 0;
            if (P.ttl_ms > 0 && (!armed || now_ms >= next_flip)) {
                const uint32_t other = active ^ 1u;
                for (uint32_t i = lane; i < ICAO_CAP; i += 32) S.gen[other][i] = ICAO_EMPTY;
                gcount[other] = 0; dirty[other] = true; active = other;
                next_flip = now_ms + P.ttl_ms; armed = 1; flipped = 1; c_flips++;
                __syncwarp();
            }
            if (lane == 0) {
                b200_buffer_result r;
                r.sample_timestamp = buf_ts; r.sum_level = 0; r.sum_power = 0; r.sum_signal_power = 0;
                r.length = d_end - d_begin; r.n_frames = nfr_buf; r.buffer_seq = seq; r.icao_flipped = flipped;
                P.buf_out[seg.first_buf + b] = r;
            }
            seq++;
        }
    }

    // write back
    for (int g = 0; g < 2; g++)
        if (dirty[g]) for (uint32_t i = lane; i < ICAO_CAP; i += 32) st->gen[g][i] = S.gen[g][i];
    uint32_t red[15] = {c_pre, c_bad, c_unk, c_acc0, c_acc1, c_tp[0], c_tp[1], c_tp[2], c_tp[3], c_tp[4], c_bp[0], c_bp[1], c_bp[2], c_bp[3], c_bp[4]};
#pragma unroll
    for (int k = 0; k < 15; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) red[k] += __shfl_xor_sync(FULLMASK, red[k], o);
    if (lane == 0) {
        st->gen_count[0] = gcount[0]; st->gen_count[1] = gcount[1];
        st->active = active; st->flip_armed = armed; st->next_flip_ms = next_flip; st->buffer_seq = seq; st->error = err;
        b200_demod_stats &s = st->stats;
        s.samples_processed += c_samples; s.demod_preambles += red[0]; s.demod_rejected_bad += red[1];
        s.demod_rejected_unknown_icao += red[2]; s.demod_accepted[0] += red[3]; s.demod_accepted[1] += red[4];
        for (int p = 0; p < 5; p++) { s.demod_preamblePhase[p] += red[5 + p]; s.demod_bestPhase[p] += red[10 + p]; }
        s.buffers += c_bufs; s.icao_flips += c_flips;
        P.frame_count[stream] = min(nframes, P.frame_cap);
        if (err) atomicOr(&P.ctl->overflow, 8u);
    }
}

// ------------------------------------------------------------------------------------------------
// finalize: prefix of per-stream frame counts, then one warp per frame
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) frame_prefix_kernel(const uint32_t *count, uint32_t *prefix, uint32_t n, RunCtl *ctl) {
    const uint32_t n_all = n;
    __shared__ uint32_t scratch[40];
    uint32_t base = 0;
    if (ctl->overflow & 3u) n = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {
        const uint32_t i = i0 + threadIdx.x;
        const uint32_t v = i < n ? count[i] : 0;
        uint32_t total;
        const uint32_t ex = block_excl_scan(v, scratch, &total);
        if (i < n) prefix[i] = base + ex;
        base += total;
    }
    if (threadIdx.x == 0) { prefix[n_all] = base; ctl->total_frames = base; }
}

__device__ __forceinline__ unsigned long long dmax_bits(double v) { return (unsigned long long)__double_as_longlong(v); }

__global__ void __launch_bounds__(256) finalize_kernel(const FinalizeParams P) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t total = P.frame_prefix[P.n_streams];
    for (uint32_t fi = warp_global; fi < total; fi += n_warps) {
        // stream = last s with prefix[s] <= fi
        uint32_t lo = 0, hi = P.n_streams;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (P.frame_prefix[mid] <= fi) lo = mid; else hi = mid; }
        const uint32_t stream = lo, k = fi - P.frame_prefix[lo];
        b200_frame *src = &P.frames[(size_t)stream * P.frame_cap + k];
        const uint32_t d = *reinterpret_cast<const uint32_t *>(&src->pad_[2]);
        // locate the segment: frames carry the low 16 bits of the segment index; segments of one stream are few
        uint32_t seg_i = P.stream_seg_begin[stream];
        {
            const uint32_t low = *reinterpret_cast<const uint16_t *>(&src->pad_[0]);
            while ((seg_i & 0xffffu) != low) seg_i++;
        }
        const Segment seg = P.segs[seg_i];
        const uint32_t len = src->signal_len;
        unsigned long long sum = 0;
        for (uint32_t i = lane; i < len; i += 32) {
            const uint32_t dd = d + 19 + i;                      // data index of the sample (demod_2400.c:443)
            uint32_t m;
            if ((seg.flags & SEG_HALO_ZERO) && dd < B200_TRAIL) m = 0;
            else {
                const uint16_t raw = *reinterpret_cast<const uint16_t *>(seg.base + 2 * (size_t)dd);
                m = (seg.flags & SEG_MAG) ? raw : P.lut_full[(raw & 0xffu) * 256 + (raw >> 8)];
            }
            sum += (unsigned long long)(m * m);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULLMASK, sum, o);
        if (lane == 0) {
            b200_frame fr = *src;
            fr.sigpow_sum = sum;
#pragma unroll
            for (int i = 0; i < 6; i++) fr.pad_[i] = 0;
            P.packed[fi] = fr;
            const uint32_t b = seg.first_buf + d / seg.buf_len;
            atomicAdd(&P.buf_acc[b].sum_signal_power, sum);
            b200_demod_stats &s = P.state[stream].stats;
            atomicAdd((unsigned long long *)&s.signal_power_count, (unsigned long long)len);
            atomicAdd((unsigned long long *)&s.sum_signal_power, sum);
            const double level = (double)sum / 65535.0 / 65535.0 / (double)len;      // demod_2400.c:448-449
            if (level > 0.50119) atomicAdd((unsigned long long *)&s.strong_signal_count, 1ull);
            atomicMax((unsigned long long *)&s.peak_signal_power, dmax_bits(level));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// tiny control-plane kernels: ICAO filter operations from the host API
// ------------------------------------------------------------------------------------------------
__global__ void icao_op_kernel(StreamState *state, uint32_t stream, int op, uint32_t addr, int *result) {
    StreamState *st = &state[stream];
    if (threadIdx.x != 0) return;
    int r = 0;
    if (op == 0) {            // add (icao_filter.c:112-130)
        r = gen_add(st->gen[st->active], &st->gen_count[st->active], addr) ? 0 : -1;
    } else if (op == 1) {     // test (icao_filter.c:132-154)
        r = (gen_has(st->gen[0], addr) || gen_has(st->gen[1], addr)) ? 1 : 0;
    } else if (op == 2) {     // expire (icao_filter.c:96-110)
        const uint32_t other = st->active ^ 1u;
        for (uint32_t i = 0; i < ICAO_CAP; i++) st->gen[other][i] = ICAO_EMPTY;
        st->gen_count[other] = 0; st->active = other; st->stats.icao_flips++;
    } else if (op == 3) {     // reset (icaoFilterInit)
        for (uint32_t i = 0; i < ICAO_CAP; i++) { st->gen[0][i] = ICAO_EMPTY; st->gen[1][i] = ICAO_EMPTY; }
        st->gen_count[0] = st->gen_count[1] = 0; st->active = 0; st->flip_armed = 0; st->next_flip_ms = 0;
    }
    if (result) *result = r;
}

// ------------------------------------------------------------------------------------------------
// launch wrappers
// ------------------------------------------------------------------------------------------------
extern "C" int b200_launch_scan(const ScanParams *p, const DeviceTables *d_tables, int n_sm, void *stream) {
    static bool configured = false;
    const size_t smem = sizeof(ScanSmem);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_kernel, SCAN_THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    uint32_t grid = (uint32_t)(n_sm * per_sm);
    if (grid > p->n_tiles) grid = p->n_tiles;
    if (grid == 0) return 0;
    scan_kernel<<<grid, SCAN_THREADS, smem, (cudaStream_t)stream>>>(*p, d_tables);
    return (int)cudaGetLastError();
}

extern "C" int b200_launch_resolve(const ResolveParams *p, void *stream) {
    if (p->n_streams == 0) return 0;
    resolve_kernel<<<p->n_streams, 32, 0, (cudaStream_t)stream>>>(*p);
    return (int)cudaGetLastError();
}

extern "C" int b200_launch_finalize(const FinalizeParams *p, uint32_t *d_frame_prefix, RunCtl *ctl, void *stream) {
    frame_prefix_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(p->frame_count, d_frame_prefix, p->n_streams, ctl);
    finalize_kernel<<<148 * 2, 256, 0, (cudaStream_t)stream>>>(*p);
    return (int)cudaGetLastError();
}

extern "C" int b200_launch_icao_op(StreamState *state, uint32_t stream, int op, uint32_t addr, int *d_result, void *cstream) {
    icao_op_kernel<<<1, 32, 0, (cudaStream_t)cstream>>>(state, stream, op, addr, d_result);
    return (int)cudaGetLastError();
}

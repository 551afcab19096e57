// scan_kernel.cu — stage A of the Mode-S demodulator: the stateless, data-parallel part of
// demodulate2400() as one persistent sm_100a kernel.
//
//   per tile of SCAN_TILE preamble start positions (+ SCAN_LOOKAHEAD samples of look-ahead):
//     1. coalesced 16-byte HBM loads of uc8 IQ, magnitude through a folded, bank-swizzled
//        lookup table in shared memory (convert.c:35-108), exact per-buffer level/power sums;
//        magnitudes staged in shared memory as uint16
//     2. one pass over the magnitudes, 16 consecutive samples per thread in a register window:
//          - pre-check of every position (demod_2400.c:311-320) -> ordered bitmap
//          - the TICK MAP: for every sample s and every correlator row r (demod_2400.c:74-93) the sign
//            bit [row r applied at s > 0], stored at tick U = 5*s + r.  PPM bit k of a frame that starts
//            at sample j with try_phase t is the single tick 5*(j+19) + t + 12*k (the closed form of
//            slice_byte, demod_2400.c:133-213), so slicing any (position, phase) later is a stride-12
//            bit gather and its cost no longer depends on how dense the candidates are.
//     3. noise-relative thresholds of the three preamble correlations (demod_2400.c:330-378)
//     4. DF gate: five ticks per (position, phase) (demod_2400.c:215-239)
//     5. full slice: one thread per surviving (position, phase) gathers 8 message bits per 96 ticks with
//        three multiply-gathers, CRC-24 by table (crc.c:67-82), then DF17 repair / single-bit-fix
//        classification (crc.c:383-418, mode_s.c:276-419)
//     6. ordered emission of PosEntry + Rec lists for the per-receiver resolver (stage B)
//
// No tensor cores: integer scan/correlate work bounded by HBM reads and instruction issue.
#include "common.h"
#include "device_utils.cuh"

#define SCAN_WARPS (SCAN_THREADS / 32)
#define MAIN_WARPS (SCAN_MAIN_THREADS / 32)
#define TICK_WORDS ((5 * (SCAN_NMAG + 24) + 31) / 32 + 8)
#define MAG_PAD 40                      // the register-window pass may read this far past SCAN_NMAG

#define WQ1_CAP   448                   // pre-check passers per warp range kept in shared memory (of 512 positions)
#define WPASS_CAP 64                    // threshold passers per warp range kept in shared memory
#define WFULL_CAP 96                    // DF-gate survivors per warp range kept in shared memory

struct TileInfo { uint32_t x0, interior, p_lo, p_hi; };

struct WarpQueues {                     // candidate discovery is warp-local: warp w owns positions [512w, 512w+512)
    uint16_t q1[MAIN_WARPS][WQ1_CAP];
    uint16_t pass_pos[MAIN_WARPS][WPASS_CAP];
    uint8_t  pass_tried[MAIN_WARPS][WPASS_CAP];
    uint32_t full[MAIN_WARPS][WFULL_CAP];
};

struct ScanSmem {
    uint16_t lut[128 * 128];            // folded, bank-swizzled UC8 magnitude table
    uint16_t mag[SCAN_NMAG + MAG_PAD];  // magnitudes of the tile, index = tile coordinate x - x0
    uint32_t tick[TICK_WORDS];          // tick U = 5*(x - x0) + row at word U >> 5, bit U & 31
    uint32_t crc_tab[256];
    uint32_t bit_syn[112];
    uint32_t syn_hash[512];
    uint32_t pre_bits[SCAN_TILE / 32];  // pre-check result, one bit per position (input of the dense-tile slow path)
    uint16_t pass_pos[SCAN_PASS_CAP];   // positions that reached a preamble threshold, ascending
    uint8_t  pass_tried[SCAN_PASS_CAP]; // phases tried
    uint8_t  pass_live[SCAN_PASS_CAP];  // phases with a filter-dependent score (a Rec exists)
    uint32_t full[SCAN_FULL_CAP];       // pass index << 4 | phase << 1 | long, for the phases that passed the DF gate (ascending)
    union {                             // the warp queues are dead once published to the block lists; the records reuse them
        WarpQueues wq;
        Rec recs[SCAN_FULL_CAP];        // one record per fully sliced phase (kind 0 = score -2 regardless of the filter)
    };
    Segment seg, seg_next;              // descriptor of the current / prefetched tile's segment
    TileInfo info, info_next;           // per-tile scalars, computed once by the helper warp
    uint32_t wcnt[MAIN_WARPS];          // per warp: passers | survivors << 16
    uint32_t scratch[40];
    uint32_t syn_mul;
    uint32_t rec_off, tile_next, overflow;
};

// mixed-sign two-way dot products: a = two unsigned 16-bit magnitudes, b = four signed 8-bit coefficients
__device__ __forceinline__ int dp2a_lo(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// Exclusive rank of a 0/1 flag over the block plus the block total (all threads must call).
__device__ __forceinline__ uint32_t block_flag_scan(bool flag, uint32_t *scratch, uint32_t *total) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t bal = __ballot_sync(FULLMASK, flag);
    __syncthreads();
    if (lane == 0) scratch[wid] = __popc(bal);
    __syncthreads();
    uint32_t tot;
    const uint32_t ex = warp_excl_scan(lane < SCAN_WARPS ? scratch[lane] : 0u, lane, &tot);
    *total = tot;
    return __shfl_sync(FULLMASK, ex, wid) + __popc(bal & ((1u << lane) - 1u));
}

// crc.c:383-406 for nfix_crc = 1: message bit (>= 5) whose single-bit syndrome equals `syn`, or -2.
__device__ __forceinline__ int diagnose1(const ScanSmem &S, uint32_t syn, int bits) {
    uint32_t e = S.syn_hash[(syn * S.syn_mul) >> 23];
    if ((e >> 8) != syn) return -2;
    int b = (int)(e & 0xffu) - (112 - bits);
    return b >= 5 ? b : -2;
}

// Filter-independent part of scoreModesMessage (mode_s.c:309-419) for a fully sliced frame with syndrome `crc`.
// Returns RecKind, or 0 when the score is -2 whatever the filter holds.
__device__ __forceinline__ uint32_t classify(const ScanSmem &S, const ScanParams &P, const uint32_t w[4], int df, bool is_long,
                                             uint32_t crc, uint32_t *addr_out, int *fixbit_out) {
    const uint32_t aa = w[0] & 0xffffffu;
    *fixbit_out = -1;
    *addr_out = 0;
    if (is_long) {
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

// 64 ticks starting at tick B as two words (lo = ticks B..B+31).
__device__ __forceinline__ void ticks64(const ScanSmem &S, uint32_t B, uint32_t &lo, uint32_t &hi) {
    const uint32_t *t = &S.tick[B >> 5];
    const uint32_t sh = B & 31u, a = t[0], b = t[1], c = t[2];
    lo = __funnelshift_r(a, b, sh); hi = __funnelshift_r(b, c, sh);
}

// Ticks at stride 12 -> adjacent bits, earliest tick in the most significant position.
__device__ __forceinline__ uint32_t gather3(uint32_t x) {   // ticks 0, 12, 24 of x -> 3 bits
    return (((x & 0x01001001u) * 0x04002001u) >> 24) & 7u;
}
__device__ __forceinline__ uint32_t gather2(uint32_t x) {   // ticks 0, 12 of x -> 2 bits
    return (((x & 0x00001001u) * 0x00002001u) >> 12) & 3u;
}

// Thresholds of the three preamble correlations at position p (demod_2400.c:330-378): phases to try.
__device__ __forceinline__ uint32_t threshold_phases(const ScanSmem &S, const ScanParams &P, uint32_t p) {
    const uint16_t *pa = &S.mag[p];
    const int base_noise = pa[5] + pa[8] + pa[16] + pa[17] + pa[18];
    const int ref_level = (base_noise * P.thr) >> 5;
    const int d23 = (int)pa[2] - (int)pa[3], s14 = pa[1] + pa[4], d1011 = (int)pa[10] - (int)pa[11];
    const int common = s14 - d23 + pa[9] + pa[12];
    uint32_t tried = 0;
    if (common - d1011 >= ref_level) tried |= 0x03;                        // try_phase 4, 5
    if (common + d1011 >= ref_level) tried |= 0x0c;                        // try_phase 6, 7
    if (s14 + 2 * d23 + d1011 + pa[12] >= ref_level) tried |= 0x10;        // try_phase 8
    return tried;
}

// DF gate (demod_2400.c:222-239): the five DF ticks of try_phase 4+ph at position p -> keep | long << 1.
__device__ __forceinline__ uint32_t df_gate(const ScanSmem &S, const ScanParams &P, uint32_t p, uint32_t ph) {
    uint32_t lo, hi;
    ticks64(S, 5u * (p + 19u) + 4u + ph, lo, hi);          // bit k at tick 12k: 0, 12, 24, 36, 48
    const uint32_t df = (gather3(lo) << 2) | gather2(hi >> 4);
    const uint32_t is_long = (P.long_set >> df) & 1u;
    return (is_long | ((P.short_set >> df) & 1u)) | (is_long << 1);
}

// Full slice of one (position, phase): 8 message bits per 96 ticks, CRC by table, classification; writes *rec.
// Returns the RecKind (0 = score -2 regardless of the filter).
__device__ __forceinline__ uint32_t slice_and_classify(const ScanSmem &S, const ScanParams &P, uint32_t p, uint32_t ph, bool is_long, Rec *rec) {
    const int nbytes = is_long ? 14 : 7;
    uint32_t B = 5u * (p + 19u) + 4u + ph;
    uint32_t w[4] = {0, 0, 0, 0}, rem = 0, tail = 0;
#pragma unroll
    for (int by = 0; by < 14; by++, B += 96) {
        if (by >= nbytes) break;
        const uint32_t *t = &S.tick[B >> 5];
        const uint32_t sh = B & 31u, a = t[0], b = t[1], c = t[2], d = t[3];
        const uint32_t x0 = __funnelshift_r(a, b, sh), x1 = __funnelshift_r(b, c, sh), x2 = __funnelshift_r(c, d, sh);
        // message bits at ticks 0,12,24 | 36,48,60 | 72,84 of this 96-tick group, MSB first
        const uint32_t byte = (gather3(x0) << 5) | (gather3(x1 >> 4) << 2) | gather2(x2 >> 8);
        w[by >> 2] |= byte << (24 - 8 * (by & 3));
        if (by < nbytes - 3) rem = ((rem << 8) ^ S.crc_tab[byte ^ ((rem >> 16) & 0xffu)]) & 0xffffffu;   // crc.c:74-77
        else tail = (tail << 8) | byte;
    }
    const uint32_t syn = rem ^ tail;                                                                  // crc.c:79-80
    const int df = (int)(w[0] >> 27);
    uint32_t addr; int fixbit;
    const uint32_t kind = classify(S, P, w, df, is_long, syn, &addr, &fixbit);
    uint32_t *rw = reinterpret_cast<uint32_t *>(rec);
    // bytes 0..13 = message, byte 14 = kind, byte 15 = fixbit (little-endian words, big-endian message)
    rw[0] = __byte_perm(w[0], 0, 0x0123); rw[1] = __byte_perm(w[1], 0, 0x0123); rw[2] = __byte_perm(w[2], 0, 0x0123);
    rw[3] = (__byte_perm(w[3], 0, 0x0123) & 0xffffu) | (kind << 16) | ((uint32_t)(fixbit & 0xff) << 24);
    rw[4] = syn; rw[5] = addr; rw[6] = 0; rw[7] = 0;
    return kind;
}

// Ordered emission of one tile's lists (block-wide): PosEntry per passer, live records packed at the front of a
// record-pool chunk sized for every sliced phase.
__device__ __forceinline__ void emit_tile(ScanSmem &S, const ScanParams &P, uint32_t tile, const uint16_t *pass_pos, const uint8_t *pass_tried,
                                          const uint8_t *pass_live, const Rec *recs, uint32_t n_pass, uint32_t n_full,
                                          bool have_chunk, uint32_t chunk_off) {
    const uint32_t tid = threadIdx.x;
    if (tid == 0) {   // the fast path issued this atomic before slicing so that nobody waits for its round trip here
        const uint32_t off = have_chunk ? chunk_off : (n_full ? atomicAdd(&P.ctl->rec_alloc, n_full) : 0);
        S.overflow = 0;
        if (off + n_full > P.ctl->rec_cap) { atomicOr(&P.ctl->overflow, 1u); S.overflow = 1; }   // host regrows the pool and reruns
        S.rec_off = off;
    }
    PosEntry *pos_out = P.pos_pool + (size_t)tile * SCAN_TILE;
    for (uint32_t r = tid; r < n_pass; r += SCAN_THREADS)
        pos_out[r] = (uint32_t)pass_pos[r] | ((uint32_t)pass_tried[r] << 16) | ((uint32_t)pass_live[r] << 21);
    __syncthreads();   // S.rec_off / S.overflow visible
    const bool pool_ok = !S.overflow;
    uint32_t n_recs = 0;
    for (uint32_t r0 = 0; r0 < n_full; r0 += SCAN_THREADS) {
        const uint32_t q = r0 + tid;
        uint4 a = {0, 0, 0, 0}, b = {0, 0, 0, 0};
        bool live = false;
        if (q < n_full) {
            a = reinterpret_cast<const uint4 *>(&recs[q])[0]; b = reinterpret_cast<const uint4 *>(&recs[q])[1];
            live = (a.w >> 16) & 0xffu;
        }
        uint32_t total;
        const uint32_t r = n_recs + block_flag_scan(live, S.scratch, &total);
        if (live && pool_ok) {
            uint4 *dst = reinterpret_cast<uint4 *>(P.rec_pool + S.rec_off + r);
            dst[0] = a; dst[1] = b;
            const uint32_t kind = (a.w >> 16) & 0xffu;
            const int fixbit = (int)(int8_t)(a.w >> 24);
            const uint32_t aa_changed = (kind == K_ES_FIX && fixbit >= 8 && fixbit <= 31) ? 1u : 0u;   // mode_s.c:560
            const uint32_t df = (a.x & 0xffu) >> 3;                     // byte 0 of the frame as sliced
            P.key_pool[S.rec_off + r] = (aa_changed ? KEY_AA_CHANGED : 0u) | (df == 17 ? KEY_DF17 : 0u) | ((df & 0x10u) ? KEY_LONG : 0u) |
                                        (kind << 24) | (b.y & 0xffffffu);
        }
        n_recs += total;
    }
    if (tid == 0) { TileOut t; t.n_pos = n_pass; t.n_rec = pool_ok ? n_recs : 0; t.rec_off = S.rec_off; t.pad_ = n_full; P.tile_out[tile] = t; }
}

// Dense-tile slow path: the same steps block-wide with every queue in this CTA's slice of a global scratch arena
// sized for the worst case (every position a candidate in all five phases) — never a dropped candidate.
__device__ __noinline__ void process_candidates_slow(ScanSmem &S, const ScanParams &P, uint32_t tile, uint8_t *scratch) {
    const uint32_t tid = threadIdx.x;
    uint16_t *q1 = reinterpret_cast<uint16_t *>(scratch);
    uint16_t *pass_pos = reinterpret_cast<uint16_t *>(scratch + 2 * SCAN_TILE);
    uint8_t *pass_tried = scratch + 4 * SCAN_TILE;
    uint8_t *pass_live = scratch + 5 * SCAN_TILE;
    uint32_t *full = reinterpret_cast<uint32_t *>(scratch + 6 * SCAN_TILE);
    Rec *recs = reinterpret_cast<Rec *>(scratch + 26 * SCAN_TILE);

    uint32_t n_q1;
    {   // ordered compaction of the pre-check bitmap into q1
        const uint32_t word = tid < SCAN_TILE / 32 ? S.pre_bits[tid] : 0;
        uint32_t off = block_excl_scan(__popc(word), S.scratch, &n_q1);
        uint32_t wbits = word;
        while (wbits) { const uint32_t b = __ffs(wbits) - 1; wbits &= wbits - 1; q1[off++] = (uint16_t)(tid * 32 + b); }
    }
    __syncthreads();
    uint32_t n_pass = 0;
    for (uint32_t r0 = 0; r0 < n_q1; r0 += SCAN_THREADS) {
        const uint32_t e = r0 + tid;
        const uint32_t p = e < n_q1 ? q1[e] : 0;
        const uint32_t tried = e < n_q1 ? threshold_phases(S, P, p) : 0;
        uint32_t total;
        const uint32_t r = n_pass + block_flag_scan(tried != 0, S.scratch, &total);
        if (tried) { pass_pos[r] = (uint16_t)p; pass_tried[r] = (uint8_t)tried; pass_live[r] = 0; }
        n_pass += total;
    }
    __syncthreads();
    uint32_t n_full = 0;
    for (uint32_t r0 = 0; r0 < 5 * n_pass; r0 += SCAN_THREADS) {
        const uint32_t i = r0 + tid, r = i / 5, ph = i - 5 * r;
        uint32_t g = 0;
        if (r < n_pass && ((pass_tried[r] >> ph) & 1u)) g = df_gate(S, P, pass_pos[r], ph);
        uint32_t total;
        const uint32_t off = n_full + block_flag_scan(g & 1u, S.scratch, &total);
        if (g & 1u) full[off] = (r << 4) | (ph << 1) | (g >> 1);
        n_full += total;
    }
    __syncthreads();
    for (uint32_t q = tid; q < n_full; q += SCAN_THREADS) {
        const uint32_t fe = full[q], r = fe >> 4, ph = (fe >> 1) & 7u;
        if (slice_and_classify(S, P, pass_pos[r], ph, fe & 1u, &recs[q]))
            atomicOr(reinterpret_cast<uint32_t *>(&pass_live[r & ~3u]), (1u << ph) << (8 * (r & 3)));
    }
    __syncthreads();
    emit_tile(S, P, tile, pass_pos, pass_tried, pass_live, recs, n_pass, n_full, false, 0);
}

// Phase 1: 16-byte HBM loads -> magnitudes in shared memory + per-buffer sums.  INTERIOR tiles (all samples are data,
// all owned samples belong to one reference buffer) skip every per-chunk boundary test.
template <bool INTERIOR>
__device__ __forceinline__ void load_convert(ScanSmem &S, const ScanParams &P, const Segment &seg, uint32_t tile, uint32_t x0) {
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    // byte address of tile coordinate x: seg.base + 2*(x - lead); x0 multiple of 8 => 16B aligned
    const uint8_t *tile_base = seg.base + 2 * ((int64_t)x0 - (int64_t)seg.lead);
    const uint32_t x_data_end = seg.lead + seg.npos + B200_TRAIL;    // first x without data
    const uint32_t x_zero_end = (seg.flags & SEG_HALO_ZERO) ? seg.lead + B200_TRAIL : seg.lead; // x below: magnitude 0, memory not read
    const bool is_mag = seg.flags & SEG_MAG;
    const bool last_tile = tile + 1 == seg.tile_begin + seg.n_tiles;
    // Power statistics (convert.c:75-79) are per reference buffer: new sample n = x - lead - 326 belongs to
    // buffer n / buf_len and is counted by the tile whose position range holds x (the last tile also owns the tail).
    const int64_t n_first = (int64_t)x0 - seg.lead - B200_TRAIL;    // new-sample index of tile coordinate x0
    const uint32_t nb0 = n_first > 0 ? (uint32_t)n_first / seg.buf_len : 0;
    const int64_t bound1 = (int64_t)(nb0 + 1) * seg.buf_len;         // first new sample of buffer nb0 + 1
    const int64_t bound2 = bound1 + seg.buf_len;

    unsigned long long acc_level = 0, acc_power = 0;
    uint32_t acc_buf = INTERIOR ? seg.first_buf + nb0 : 0xffffffffu;
    // 16 main warps: chunks [0, SCAN_TILE/8) in two rounds; helper warp: the look-ahead chunks — two rounds each, balanced
    const bool helper = tid >= SCAN_MAIN_THREADS;
#pragma unroll 1
    for (uint32_t k = 0; k < 2; k++) {
        const uint32_t c = helper ? SCAN_TILE / 8 + k * 32 + lane : k * SCAN_MAIN_THREADS + tid;
        if (c >= SCAN_NMAG / 8) continue;
        const uint32_t xc = x0 + c * 8;
        uint32_t m[8];
        if (!INTERIOR && (xc + 8 <= x_zero_end || xc >= x_data_end)) {
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
            if (INTERIOR) {
                if (c < SCAN_TILE / 8) {
                    acc_level += (m[0] + m[1]) + (m[2] + m[3]) + (m[4] + m[5]) + (m[6] + m[7]);
#pragma unroll
                    for (int i = 0; i < 8; i++) mad_wide(acc_power, m[i], m[i]);
                }
            } else {
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
        }
        uint4 packed;
        packed.x = m[0] | (m[1] << 16); packed.y = m[2] | (m[3] << 16);
        packed.z = m[4] | (m[5] << 16); packed.w = m[6] | (m[7] << 16);
        *reinterpret_cast<uint4 *>(&S.mag[c * 8]) = packed;
    }
    // flush the statistics: one atomic pair per warp when the whole warp fed the same buffer
    const uint32_t b0 = __shfl_sync(FULLMASK, acc_buf, 0);
    if (INTERIOR || __all_sync(FULLMASK, acc_buf == b0)) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { acc_level += __shfl_xor_sync(FULLMASK, acc_level, o); acc_power += __shfl_xor_sync(FULLMASK, acc_power, o); }
        if (lane == 0 && b0 != 0xffffffffu) { atomicAdd(&P.buf_acc[b0].sum_level, acc_level); atomicAdd(&P.buf_acc[b0].sum_power, acc_power); }
    } else if (acc_buf != 0xffffffffu) {
        atomicAdd(&P.buf_acc[acc_buf].sum_level, acc_level); atomicAdd(&P.buf_acc[acc_buf].sum_power, acc_power);
    }
}

// Correlator rows (demod_2400.c:74-93) NEGATED and packed as four signed bytes (c0, c1, c2, c3):
// the sign bit of the negated correlation is exactly [correlation > 0].
#define NEG_ROW0 0x00030feeu   // -18,  15,   3,  0
#define NEG_ROW1 0x000905f2u   // -14,   5,   9,  0
#define NEG_ROW2 0x0014fbf0u   // -16,  -5,  20,  0
#define NEG_ROW3 0x0012f5f9u   //  -7, -11,  18,  0
#define NEG_ROW4 0xff14f1fcu   //  -4, -15,  20, -1

// Phase 2: thread = 16 consecutive samples/positions starting at i0 (window of 32 magnitudes in registers).
// Returns the 16-bit pre-check mask; writes the 80 ticks of its samples (two lanes share five words).
__device__ __forceinline__ uint32_t window_pass(ScanSmem &S, uint32_t i0, uint32_t lane, bool store) {
    const uint4 *src = reinterpret_cast<const uint4 *>(&S.mag[i0]);
    const uint4 q0 = src[0], q1v = src[1], q2 = src[2], q3 = src[3];
    const uint32_t wv[16] = {q0.x, q0.y, q0.z, q0.w, q1v.x, q1v.y, q1v.z, q1v.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
    uint32_t mask = 0;
    {   // pre-check (demod_2400.c:311-320): pa[1] > pa[7] && pa[12] > pa[14] && pa[12] > pa[15]
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 16; i++) { v[2 * i] = wv[i] & 0xffffu; v[2 * i + 1] = wv[i] >> 16; }
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (v[i + 1] > v[i + 7] && v[i + 12] > v[i + 14] && v[i + 12] > v[i + 15]) mask |= 1u << i;
    }
    // tick map: sample i needs the pairs (m[i], m[i+1]) and (m[i+2], m[i+3]); odd i takes them from the 16-bit-shifted words
    uint32_t xs[10];
#pragma unroll
    for (int j = 0; j < 10; j++) xs[j] = __funnelshift_r(wv[j], wv[j + 1], 16);
    uint32_t acc[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t A = (i & 1) ? xs[i >> 1] : wv[i >> 1];
        const uint32_t Bp = (i & 1) ? xs[(i >> 1) + 1] : wv[(i >> 1) + 1];
        const uint32_t rows[5] = {NEG_ROW0, NEG_ROW1, NEG_ROW2, NEG_ROW3, NEG_ROW4};
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const int nv = dp2a_hi(Bp, rows[r], dp2a_lo(A, rows[r], 0));
            const int g = 5 * i + r;                                           // tick within this thread's 80
            acc[g >> 5] = __funnelshift_l((uint32_t)nv, acc[g >> 5], 1);       // shift the sign bit in, earliest tick ends up highest
        }
    }
    // acc[0], acc[1] hold 32 ticks each, earliest in bit 31 -> bit-reverse; acc[2] holds 16 ticks in its low half
    const uint32_t t0 = __brev(acc[0]), t1 = __brev(acc[1]), t2 = __brev(acc[2]) >> 16;
    // lanes 2j, 2j+1 own ticks [160j', 160j'+160): five words
    uint32_t *dst = &S.tick[(i0 >> 5) * 5];
    const uint32_t other_t0 = __shfl_down_sync(FULLMASK, t0, 1);
    if (store) {
        if ((lane & 1) == 0) { dst[0] = t0; dst[1] = t1; dst[2] = t2 | (other_t0 << 16); }
        else { dst[3] = __funnelshift_r(t0, t1, 16); dst[4] = __funnelshift_r(t1, t2, 16); }
    }
    return mask;
}

// Per-tile scalars every thread needs, computed by one lane while the previous tile is being processed.
__device__ __forceinline__ TileInfo make_tile_info(const Segment &seg, uint32_t tile) {
    TileInfo t;
    t.x0 = (tile - seg.tile_begin) * SCAN_TILE;                           // tile origin in tile coordinates
    const int64_t n_first = (int64_t)t.x0 - seg.lead - B200_TRAIL;
    const uint32_t x_zero_end = (seg.flags & SEG_HALO_ZERO) ? seg.lead + B200_TRAIL : seg.lead;
    bool interior = t.x0 >= x_zero_end && t.x0 + SCAN_NMAG <= seg.lead + seg.npos + B200_TRAIL && n_first >= 0 &&
                    tile + 1 != seg.tile_begin + seg.n_tiles;
    if (interior) {   // all owned samples in one reference buffer?
        const uint32_t nb0 = (uint32_t)n_first / seg.buf_len;
        interior = (uint64_t)n_first + SCAN_TILE <= (uint64_t)(nb0 + 1) * seg.buf_len;
    }
    t.interior = interior;
    t.p_lo = seg.lead > t.x0 ? seg.lead - t.x0 : 0;                        // first real position
    t.p_hi = min((uint32_t)SCAN_TILE, seg.lead + seg.npos > t.x0 ? seg.lead + seg.npos - t.x0 : 0u);
    return t;
}

__global__ void __maxnreg__(56) scan_kernel(const ScanParams P, const DeviceTables *__restrict__ tables) {
    extern __shared__ uint4 smem_raw[];
    ScanSmem &S = *reinterpret_cast<ScanSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    // one-time table staging (persistent CTA) and the first tile
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(tables->lut_fold);
        uint4 *dst = reinterpret_cast<uint4 *>(S.lut);
        for (uint32_t i = tid; i < sizeof(S.lut) / 16; i += SCAN_THREADS) dst[i] = src[i];
        for (uint32_t i = tid; i < 256; i += SCAN_THREADS) S.crc_tab[i] = tables->crc_tab[i];
        for (uint32_t i = tid; i < 112; i += SCAN_THREADS) S.bit_syn[i] = tables->bit_syn[i];
        for (uint32_t i = tid; i < 512; i += SCAN_THREADS) S.syn_hash[i] = tables->syn_hash[i];
        for (uint32_t i = tid; i < MAG_PAD; i += SCAN_THREADS) S.mag[SCAN_NMAG + i] = 0;
        if (tid == 0) {
            S.syn_mul = tables->syn_hash_mul;
            const uint32_t t = atomicAdd(&P.ctl->tile_counter, 1u);
            S.tile_next = t;
            if (t < P.n_tiles) { const Segment sg = P.segs[P.tile_seg[t]]; S.seg_next = sg; S.info_next = make_tile_info(sg, t); }
        }
    }
    const bool helper_lane0 = tid == SCAN_MAIN_THREADS;

    for (;;) {
        __syncthreads();
        const uint32_t tile = S.tile_next;
        if (tile >= P.n_tiles) break;
        if (tid < sizeof(Segment) / 4) reinterpret_cast<uint32_t *>(&S.seg)[tid] = reinterpret_cast<const uint32_t *>(&S.seg_next)[tid];
        if (tid >= 32 && tid < 32 + sizeof(TileInfo) / 4) reinterpret_cast<uint32_t *>(&S.info)[tid - 32] = reinterpret_cast<const uint32_t *>(&S.info_next)[tid - 32];
        __syncthreads();
        // Software-pipelined tile fetch (lane 0 of the helper warp only): the atomic is issued here, its result is first
        // used after phase 1, the segment it names is loaded after phase 2 — three round trips hidden behind this tile's work.
        uint32_t nxt = 0;
        if (helper_lane0) nxt = atomicAdd(&P.ctl->tile_counter, 1u);

        const Segment &seg = S.seg;
        const uint32_t x0 = S.info.x0;

        // ---- phase 1: load + convert ----------------------------------------------------------------
        if (S.info.interior) load_convert<true>(S, P, seg, tile, x0); else load_convert<false>(S, P, seg, tile, x0);
        uint32_t nxt_seg = 0;
        if (helper_lane0) { S.tile_next = nxt; if (nxt < P.n_tiles) nxt_seg = P.tile_seg[nxt]; }
        __syncthreads();                                                   // B: magnitudes complete

        // ---- phase 2: pre-check masks + tick map, 16 positions per lane; main warp w owns positions [512w, 512w+512),
        //      the helper warp the look-ahead samples SCAN_TILE .. SCAN_NMAG-1 (ticks only) ------------------------------
        const bool is_main = tid < SCAN_MAIN_THREADS;
        const uint32_t i0 = is_main ? tid * 16 : SCAN_TILE + lane * 16;
        uint32_t mask = window_pass(S, i0, lane, is_main || lane < (SCAN_LOOKAHEAD + 31) / 32 * 2);
        if (is_main) {
            const uint32_t p_lo = S.info.p_lo, p_hi = S.info.p_hi;
            // positions outside [p_lo, p_hi) are not preamble starts of this segment
            const uint32_t lo_cut = p_lo > i0 ? min(p_lo - i0, 16u) : 0u, hi_cut = p_hi > i0 ? min(p_hi - i0, 16u) : 0u;
            mask &= (0xffffu << lo_cut) & ((1u << hi_cut) - 1u);
            const uint32_t both = (mask << (16 * (lane & 1))) | __shfl_xor_sync(FULLMASK, mask << (16 * (lane & 1)), 1);
            if ((lane & 1) == 0) S.pre_bits[i0 >> 5] = both;
        } else {
            mask = 0;
            if (lane == 0 && nxt < P.n_tiles) { const Segment sg = P.segs[nxt_seg]; S.seg_next = sg; S.info_next = make_tile_info(sg, nxt); }
        }

        // ---- warp-local discovery: pre-check passers -> thresholds (magnitudes only, no tick needed yet) ----------------
        uint32_t n_wpass = 0;
        bool wover = false;
        if (is_main) {
            uint16_t *wq1 = S.wq.q1[wid];
            uint32_t n_wq1;
            uint32_t off = warp_excl_scan(__popc(mask), lane, &n_wq1);
            uint32_t mb = mask;
            if (n_wq1 > WQ1_CAP) { wover = true; n_wq1 = 0; mb = 0; }
            while (mb) { const uint32_t b = __ffs(mb) - 1; mb &= mb - 1; wq1[off++] = (uint16_t)(i0 + b); }
            __syncwarp();
            for (uint32_t r0 = 0; r0 < n_wq1; r0 += 32) {
                const uint32_t e = r0 + lane;
                const uint32_t p = e < n_wq1 ? wq1[e] : 0;
                const uint32_t tried = e < n_wq1 ? threshold_phases(S, P, p) : 0;
                const uint32_t bal = __ballot_sync(FULLMASK, tried != 0);
                const uint32_t r = n_wpass + __popc(bal & ((1u << lane) - 1u));
                if (tried) { if (r < WPASS_CAP) { S.wq.pass_pos[wid][r] = (uint16_t)p; S.wq.pass_tried[wid][r] = (uint8_t)tried; } else wover = true; }
                n_wpass += __popc(bal);
            }
        }
        __syncthreads();                                                   // C: tick map complete

        // ---- warp-local DF gate over the warp's own passers ------------------------------------------------------------------
        uint32_t n_wfull = 0;
        if (is_main && !__any_sync(FULLMASK, wover)) {
            for (uint32_t r0 = 0; r0 < 5 * n_wpass; r0 += 32) {
                const uint32_t i = r0 + lane, r = i / 5, ph = i - 5 * r;
                uint32_t g = 0;
                if (r < n_wpass && ((S.wq.pass_tried[wid][r] >> ph) & 1u)) g = df_gate(S, P, S.wq.pass_pos[wid][r], ph);
                const uint32_t bal = __ballot_sync(FULLMASK, g & 1u);
                const uint32_t q = n_wfull + __popc(bal & ((1u << lane) - 1u));
                if (g & 1u) { if (q < WFULL_CAP) S.wq.full[wid][q] = (r << 4) | (ph << 1) | (g >> 1); else wover = true; }
                n_wfull += __popc(bal);
            }
        }
        wover = __any_sync(FULLMASK, wover);
        if (lane == 0 && is_main) S.wcnt[wid] = wover ? 0xffffffffu : (n_wpass | (n_wfull << 16));
        __syncthreads();                                                   // D: per-warp counts published

        // ---- publish to the ordered block lists -----------------------------------------------------------------------------------
        uint32_t n_pass, n_full;
        bool dense;
        {
            const uint32_t c = lane < MAIN_WARPS ? S.wcnt[lane] : 0u;
            dense = __any_sync(FULLMASK, c == 0xffffffffu);
            uint32_t tot;
            const uint32_t ex = warp_excl_scan(dense ? 0u : c, lane, &tot);   // both 16-bit counters in one add: no carry (<= 8192 each)
            n_pass = tot & 0xffffu; n_full = tot >> 16;
            dense = dense || n_pass > SCAN_PASS_CAP || n_full > SCAN_FULL_CAP;
            if (!dense && is_main) {
                const uint32_t mine = __shfl_sync(FULLMASK, ex, wid), base_p = mine & 0xffffu, base_f = mine >> 16;
                for (uint32_t e = lane; e < n_wpass; e += 32) {
                    S.pass_pos[base_p + e] = S.wq.pass_pos[wid][e]; S.pass_tried[base_p + e] = S.wq.pass_tried[wid][e]; S.pass_live[base_p + e] = 0;
                }
                for (uint32_t e = lane; e < n_wfull; e += 32) S.full[base_f + e] = S.wq.full[wid][e] + (base_p << 4);
            }
        }
        __syncthreads();                                                   // E: block lists complete, warp queues dead

        if (dense) {      // denser than the shared-memory queues: redo the candidate steps on the global scratch arena
            if (P.scratch) {
                process_candidates_slow(S, P, tile, P.scratch + (size_t)blockIdx.x * SCAN_SCRATCH_BYTES);
            } else if (tid == 0) {      // no scratch arena yet: tell the host, which allocates one and reruns
                atomicOr(&P.ctl->overflow, 2u);
                TileOut t = {0, 0, 0, 0};
                P.tile_out[tile] = t;
            }
            continue;
        }

        // ---- full slice, one thread per (position, phase) ---------------------------------------------------------------------------
        uint32_t chunk_off = 0;
        if (tid == 0 && n_full) chunk_off = atomicAdd(&P.ctl->rec_alloc, n_full);   // record-pool chunk: used only after the slicing below
        for (uint32_t q = tid; q < n_full; q += SCAN_THREADS) {
            const uint32_t fe = S.full[q], r = fe >> 4, ph = (fe >> 1) & 7u;
            if (slice_and_classify(S, P, S.pass_pos[r], ph, fe & 1u, &S.recs[q]))
                atomicOr(reinterpret_cast<uint32_t *>(&S.pass_live[r & ~3u]), (1u << ph) << (8 * (r & 3)));
        }
        __syncthreads();                                                   // F: records staged
        emit_tile(S, P, tile, S.pass_pos, S.pass_tried, S.pass_live, S.recs, n_pass, n_full, true, chunk_off);
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

// modeac_kernel.cu — Mode A/C demodulator (reference demodulate2400AC, demod_2400.c:575-761).
//
// Runs only for contexts created with B200_CFG_MODE_AC (readsb --modeac), after the scan kernel of the same run:
// the reply detector needs the buffer's mean level and mean power (noise floor, demod_2400.c:580-581), i.e. the
// exact per-buffer sums the scan kernel produces, so it cannot share that kernel's single pass over the samples.
//
//   modeac_noise_kernel   one thread per reference buffer: the fp64 noise floor, once (not per position)
//   modeac_scan_kernel    stateless per position, like stage A: is there a well-formed reply whose F1 pulse starts here?
//                         Output is ONE BIT per position (1 KB per tile), so its size does not depend on the input:
//                         no candidate pool, no overflow, no repeat.
//   modeac_walk_kernel    the reference's sequential part, one warp per reference buffer (the skip state does not
//                         survive a buffer): greedy walk over the bit map (a reply hides the next 69 positions), then
//                         the lanes decode the accepted replies in parallel straight from the samples.
//
// Exactness notes: float products and divisions are single IEEE operations (--fmad=false keeps them un-fused), the
// sqrt(2) factors are fp64 exactly as the C expressions promote them, sqrt / sqrtf are the correctly rounded CUDA
// versions, `noise_level * f1f2_level` wraps in 32 bits like the reference's unsigned product, and
// (unsigned)((double)x + 0.5) of a non-negative float x is evaluated as trunc(x) + (x - trunc(x) >= 0.5), which is the
// same integer for every x (both subtractions are exact in fp32).
#include "common.h"
#include "device_utils.cuh"

#define AC_THREADS 512
#define AC_WARPS (AC_THREADS / 32)
#define AC_BEHIND 8            // magnitudes kept before the tile origin (1 needed, 8 keeps 16-byte alignment)
#define AC_AHEAD 80            // after the last position (<= f1 + 69 is read)
#define AC_TILE 8192           // positions per block iteration = AC_TILE / SCAN_TILE consecutive scan tiles of one segment
#define AC_NMAG (AC_BEHIND + AC_TILE + AC_AHEAD)
#define AC_SKIP (20 * 87 / 25 + 1)   // positions hidden by an accepted reply (demod_2400.c:753 + the loop increment)

struct AcSmem {
    uint16_t lut[128 * 128];                 // folded + swizzled uc8 table, as in the scan kernel
    alignas(16) uint16_t mag[AC_NMAG + 8];
    uint16_t q1[AC_TILE];                    // positions that passed the F1 tests (at most 2 of 3 can), then in place: F2 survivors
    uint32_t bitmap[AC_TILE / 32];
    uint32_t q1n;
};

// ---- the reply detector, shared by the scan (magnitudes in shared memory) and the walk (samples from global memory) ----

struct SmemMag {              // m[k] = data[f1_sample + k]
    const uint16_t *m;
    __device__ __forceinline__ uint32_t operator()(int k) const { return m[k]; }
};

struct GlobalMag {
    const uint8_t *base;      // segment data index 0
    const uint16_t *lut;      // full 65536-entry table
    int64_t d;                // data index of f1_sample in the segment
    int64_t zero_end, data_end;
    bool is_mag;
    __device__ __forceinline__ uint32_t operator()(int k) const {
        const int64_t i = d + k;
        if (i < zero_end || i >= data_end) return 0;
        const uint32_t v = __ldg(reinterpret_cast<const uint16_t *>(base + 2 * i));
        return is_mag ? v : __ldg(&lut[v]);      // the table is symmetric in I and Q: the little-endian pair indexes it directly
    }
};

__device__ __forceinline__ uint32_t round_half_up(float x) {     // (unsigned)((double)x + 0.5), x >= 0
    const uint32_t r = (uint32_t)x;
    return r + ((x - (float)r) >= 0.5f ? 1u : 0u);
}

// demod_2400.c:630-672: F1 edge / quiet / level, clock phase, F2 edge / quiet / level.  j = f1_sample.
template <class M>
__device__ __forceinline__ bool ac_front(const M &m, uint32_t j, uint32_t noise_level, uint32_t *f1_clock_out, uint32_t *f1f2_out) {
    const uint32_t a0 = m(0), a1 = m(1), a2 = m(2);
    if (!(m(-1) < a0) || a2 > a0 || a2 > a1) return false;
    const uint32_t f1_level = (a0 + a1) / 2;
    if (noise_level * 2 > f1_level) return false;
    // :651-654 clock phase from the power that spilled into the second sample
    const float f1a = (float)a0 * (float)a0, f1b = (float)a1 * (float)a1;
    const float fraction = f1b / (f1a + f1b);
    const uint32_t f1_clock = round_half_up(25.0f * ((float)j + fraction * fraction));
    const int o2 = (int)((f1_clock + 87 * 14) / 25 - j);
    const uint32_t b0 = m(o2), b1 = m(o2 + 1), b2 = m(o2 + 2);
    if (!(m(o2 - 1) < b0) || b2 > b0 || b2 > b1) return false;
    const uint32_t f2_level = (b0 + b1) / 2;
    if (noise_level * 2 > f2_level) return false;
    *f1_clock_out = f1_clock;
    *f1f2_out = max(f1_level, f2_level);
    return true;
}

// demod_2400.c:674-731: thresholds, 20 bit cells, framing; returns the Mode A/C word or 0xffffffff.
template <class M>
__device__ __forceinline__ uint32_t ac_bits(const M &m, uint32_t j, uint32_t noise_level, uint32_t f1_clock, uint32_t f1f2) {
    const float midpoint = sqrtf((float)(noise_level * f1f2));            // :676, 32-bit wrap included
    const uint32_t signal_threshold = (uint32_t)((double)midpoint * 1.41421356237309504880 + 0.5);
    const uint32_t noise_threshold = (uint32_t)((double)midpoint / 1.41421356237309504880 + 0.5);
    uint32_t bits = 0, bad = 0, clock = f1_clock;
#pragma unroll 4
    for (int bit = 0; bit < 20; ++bit, clock += 87) {
        const int o = (int)(clock / 25 - j);
        const uint32_t s0 = m(o), s1 = m(o + 1), s2 = m(o + 2);
        bits <<= 1;
        if (s2 >= signal_threshold) bad = 1;                                   // noisy quiet period
        if (s0 >= signal_threshold || s1 >= signal_threshold) bits |= 1;
        else if (s0 > noise_threshold && s1 > noise_threshold) bad = 1;        // uncertain
    }
    if ((bits & 0x80020) != 0x80020 || (bits & 0x0101B) != 0 || bad) return 0xffffffffu;
    return ((bits & 0x40000) ? 0x0010 : 0) | ((bits & 0x20000) ? 0x1000 : 0) | ((bits & 0x10000) ? 0x0020 : 0) |      // :716-731
           ((bits & 0x08000) ? 0x2000 : 0) | ((bits & 0x04000) ? 0x0040 : 0) | ((bits & 0x02000) ? 0x4000 : 0) |
           ((bits & 0x00800) ? 0x0100 : 0) | ((bits & 0x00400) ? 0x0001 : 0) | ((bits & 0x00200) ? 0x0200 : 0) |
           ((bits & 0x00100) ? 0x0002 : 0) | ((bits & 0x00080) ? 0x0400 : 0) | ((bits & 0x00040) ? 0x0004 : 0) |
           ((bits & 0x00004) ? 0x0080 : 0);
}

// ---- noise floor per reference buffer --------------------------------------------------------------------------------
__global__ void modeac_noise_kernel(const AcScanParams P) {
    for (uint32_t si = blockIdx.x; si < P.n_segs; si += gridDim.x) {
        const Segment seg = P.segs[si];
        for (uint32_t b = threadIdx.x; b < seg.n_bufs; b += blockDim.x) {
            const uint32_t len = min(seg.buf_len, seg.npos - b * seg.buf_len);
            if (len == 0) { P.noise[seg.first_buf + b] = 0; continue; }       // empty buffer: nothing will ask for its noise floor
            const BufAcc &a = P.buf_acc[seg.first_buf + b];     // sum_signal_power may still be accumulating: not read
            const AcLevel lv = P.levels[seg.first_buf + b];
            double mean_level, mean_power;
            if (lv.mode == AC_LEVEL_GIVEN) { mean_level = lv.mean_level; mean_power = lv.mean_power; }
            else if (lv.mode == AC_LEVEL_FSUM) {                // `sum_level / nsamples`: a float division, widened afterwards (convert.c:243-249)
                const float2 f = P.fsum[lv.idx];
                mean_level = (double)(f.x / (float)len); mean_power = (double)(f.y / (float)len);
            } else {
                mean_level = (double)a.sum_level / 65536.0 / (double)len;            // convert.c:100-102
                mean_power = (double)a.sum_power / 65535.0 / 65535.0 / (double)len;  // convert.c:104-106
            }
            const double noise_stddev = sqrt(mean_power - mean_level * mean_level);           // demod_2400.c:580
            P.noise[seg.first_buf + b] = (uint32_t)((mean_power + noise_stddev) * 65535 + 0.5);   // :581
        }
    }
}

// ---- stateless scan: one bit per position ----------------------------------------------------------------------------
__global__ void __launch_bounds__(AC_THREADS, 2) modeac_scan_kernel(const AcScanParams P) {
    extern __shared__ __align__(16) unsigned char ac_smem_raw[];
    AcSmem &S = *reinterpret_cast<AcSmem *>(ac_smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (P.ctl->overflow & 3u) return;
    {   // the folded table, once per CTA
        const uint4 *src = reinterpret_cast<const uint4 *>(P.tables->lut_fold);
        uint4 *dst = reinterpret_cast<uint4 *>(S.lut);
        for (uint32_t i = tid; i < 128 * 128 * 2 / 16; i += AC_THREADS) dst[i] = src[i];
    }
    for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        const uint32_t ts = P.tile_seg[tile];
        if (!(ts & TILE_QUAD_START)) continue;                              // a block iteration covers one quad: the scan tiles [tile, tile + 4) of the segment
        const Segment seg = P.segs[ts & ~TILE_QUAD_START];
        const uint32_t x0 = (tile - seg.tile_begin) * SCAN_TILE;
        const uint32_t x_data_end = seg.lead + seg.npos + B200_TRAIL;
        const uint32_t x_zero_end = (seg.flags & SEG_HALO_ZERO) ? seg.lead + B200_TRAIL : seg.lead;
        const bool is_mag = seg.flags & SEG_MAG;
        __syncthreads();                                                   // the previous iteration's bitmap has been written out
        if (tid < AC_TILE / 32) S.bitmap[tid] = 0;
        if (tid == 0) S.q1n = 0;
        // magnitudes of tile coordinates [x0 - AC_BEHIND, x0 + AC_TILE + AC_AHEAD): shared index = x - x0 + AC_BEHIND
        for (uint32_t c = tid; c < AC_NMAG / 8; c += AC_THREADS) {
            const int64_t xc = (int64_t)x0 - AC_BEHIND + (int64_t)c * 8;
            uint32_t m[8];
            if (xc < 0 || xc + 8 <= (int64_t)x_zero_end || xc >= (int64_t)x_data_end) {
#pragma unroll
                for (int i = 0; i < 8; i++) m[i] = 0;
            } else {
                const uint4 raw = ldg_stream_u4(seg.base + 2 * (xc - (int64_t)seg.lead));
                const uint32_t wv[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (is_mag) { m[2 * i] = wv[i] & 0xffffu; m[2 * i + 1] = wv[i] >> 16; }
                    else uc8_pair_to_mag(S.lut, wv[i], m[2 * i], m[2 * i + 1]);
                }
                if (xc < (int64_t)x_zero_end || xc + 8 > (int64_t)x_data_end) {
#pragma unroll
                    for (int i = 0; i < 8; i++) if (xc + i < (int64_t)x_zero_end || xc + i >= (int64_t)x_data_end) m[i] = 0;
                }
            }
            uint4 packed;
            packed.x = m[0] | (m[1] << 16); packed.y = m[2] | (m[3] << 16);
            packed.z = m[4] | (m[5] << 16); packed.w = m[6] | (m[7] << 16);
            *reinterpret_cast<uint4 *>(&S.mag[c * 8]) = packed;
        }
        __syncthreads();

        // position p of the tile = data index d_tile0 + p of the segment = f1_sample (d mod buf_len) of buffer d / buf_len
        const int64_t d_tile0 = (int64_t)x0 - (int64_t)seg.lead;
        const uint32_t bt = d_tile0 > 0 ? (uint32_t)d_tile0 / seg.buf_len : 0;     // buffer of the tile's first position
        const int64_t bt_d0 = (int64_t)bt * seg.buf_len;

        // ---- phase A: F1 edge / quiet / level for 8 consecutive positions per thread, two passes ------------------
#pragma unroll 1
        for (uint32_t pass = 0; pass < AC_TILE / (8 * AC_THREADS); pass++) {
            const uint32_t p = 8 * (pass * AC_THREADS + tid);
            const uint16_t *mp = &S.mag[p + AC_BEHIND];
            uint32_t v[11];
            v[0] = mp[-1];
            {
                const uint4 B = *reinterpret_cast<const uint4 *>(mp);
                const uint32_t Cw = *reinterpret_cast<const uint32_t *>(mp + 8);
                v[1] = B.x & 0xffffu; v[2] = B.x >> 16; v[3] = B.y & 0xffffu; v[4] = B.y >> 16;
                v[5] = B.z & 0xffffu; v[6] = B.z >> 16; v[7] = B.w & 0xffffu; v[8] = B.w >> 16;
                v[9] = Cw & 0xffffu; v[10] = Cw >> 16;
            }
            // which of the 8 positions are f1_sample values of this segment (1 <= f1_sample < length), and their noise floor
            const int64_t d0 = d_tile0 + p;
            uint32_t valid = 0, noise2 = 0;
            if (d0 + 8 > 0 && d0 < (int64_t)seg.npos) {
                uint32_t b = bt;
                int64_t jj = d0 - bt_d0;
                while (jj >= (int64_t)seg.buf_len) { jj -= seg.buf_len; b++; }
                const uint32_t len_b = b < seg.n_bufs ? min(seg.buf_len, seg.npos - b * seg.buf_len) : 0;
                if (d0 >= 0 && jj + 8 <= (int64_t)len_b) {          // all eight in buffer b: level test here, with its noise floor
                    valid = jj == 0 ? 0xfeu : 0xffu;
                    noise2 = 2 * P.noise[seg.first_buf + b];
                } else {                                             // straddles a buffer edge: level test left to phase B
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int64_t d = d0 + i;
                        if (d < 0 || d >= (int64_t)seg.npos) continue;
                        const int64_t ji = jj + i < (int64_t)seg.buf_len ? jj + i : jj + i - seg.buf_len;
                        if (ji >= 1) valid |= 1u << i;
                    }
                }
            }
            uint32_t mask = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t a = v[i], b = v[i + 1], c = v[i + 2], e = v[i + 3];
                const bool ok = a < b && e <= b && e <= c && ((b + c) >> 1) >= noise2;
                mask |= ok ? 1u << i : 0u;
            }
            mask &= valid;
            uint32_t wtot;
            uint32_t off = warp_excl_scan(__popc(mask), lane, &wtot);
            uint32_t base = 0;
            if (lane == 0 && wtot) base = atomicAdd(&S.q1n, wtot);
            off += __shfl_sync(FULLMASK, base, 0);
            while (mask) { const uint32_t i = __ffs(mask) - 1; mask &= mask - 1; S.q1[off++] = (uint16_t)(p + i); }
        }
        __syncthreads();

        // ---- phase B: clock phase + F2 tests over the warp's slice of the queue, survivors compacted in place;
        //      phase C: bit cells of the survivors ----------------------------------------------------------------
        {
            const uint32_t n1 = S.q1n;
            const uint32_t lo = n1 * wid / AC_WARPS, hi = n1 * (wid + 1) / AC_WARPS;
            uint32_t wr = lo;
            for (uint32_t r0 = lo; r0 < hi; r0 += 32) {
                const bool has = r0 + lane < hi;
                const uint32_t p = has ? S.q1[r0 + lane] : 0;
                bool surv = false;
                if (has) {
                    uint32_t b = bt;
                    int64_t jj = d_tile0 + p - bt_d0;
                    while (jj >= (int64_t)seg.buf_len) { jj -= seg.buf_len; b++; }
                    uint32_t f1_clock, f1f2;
                    surv = ac_front(SmemMag{&S.mag[p + AC_BEHIND]}, (uint32_t)jj, P.noise[seg.first_buf + b], &f1_clock, &f1f2);
                }
                const uint32_t bal = __ballot_sync(FULLMASK, surv);
                if (surv) S.q1[wr + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)p;
                wr += __popc(bal);
                __syncwarp();
            }
            for (uint32_t r0 = lo; r0 < wr; r0 += 32) {
                if (r0 + lane < wr) {
                    const uint32_t p = S.q1[r0 + lane];
                    uint32_t b = bt;
                    int64_t jj = d_tile0 + p - bt_d0;
                    while (jj >= (int64_t)seg.buf_len) { jj -= seg.buf_len; b++; }
                    const uint32_t noise = P.noise[seg.first_buf + b];
                    const SmemMag m{&S.mag[p + AC_BEHIND]};
                    uint32_t f1_clock = 0, f1f2 = 0;
                    ac_front(m, (uint32_t)jj, noise, &f1_clock, &f1f2);
                    if (ac_bits(m, (uint32_t)jj, noise, f1_clock, f1f2) != 0xffffffffu) atomicOr(&S.bitmap[p >> 5], 1u << (p & 31));
                }
            }
        }
        __syncthreads();
        {   // one bit per position; only the words of this segment's own scan tiles
            const uint32_t words = min((uint32_t)(AC_TILE / 32), (seg.tile_begin + seg.n_tiles - tile) * (SCAN_TILE / 32));
            if (tid < words) P.bitmap[(size_t)tile * (SCAN_TILE / 32) + tid] = S.bitmap[tid];
        }
    }
}

// ---- the sequential part, one warp per reference buffer ---------------------------------------------------------------
__global__ void __launch_bounds__(256) modeac_walk_kernel(const AcWalkParams P) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (P.ctl->overflow & 3u) return;
    for (uint32_t si = blockIdx.x; si < P.n_segs; si += gridDim.x) {
        const Segment seg = P.segs[si];
        const uint32_t stream_first_buf = P.segs[P.stream_seg_begin[seg.stream]].first_buf;
        const uint32_t *bits = P.bitmap + (size_t)seg.tile_begin * (SCAN_TILE / 32);    // bit x = tile coordinate x of the segment
        for (uint32_t b = wid; b < seg.n_bufs; b += nw) {
            const uint32_t len_b = min(seg.buf_len, seg.npos - b * seg.buf_len);
            if (len_b == 0) {       // an empty buffer (a frontend that had nothing to deliver): no positions, no tiles, no bit-map words
                if (lane == 0) P.ac_count[seg.first_buf + b] = 0;
                continue;
            }
            const uint32_t xa = seg.lead + b * seg.buf_len, xb = xa + len_b;
            const uint32_t w_first = xa >> 5, w_last = (xb - 1) >> 5;
            b200_modeac *out = P.ac_out + (size_t)(seg.first_buf + b) * P.per_buf_cap;
            uint32_t n = 0, next_ok = 0;
            uint32_t word_next = w_first + lane <= w_last ? bits[w_first + lane] : 0;
            for (uint32_t w0 = w_first; w0 <= w_last; w0 += 32) {
                const uint32_t wi = w0 + lane, wx = wi * 32;
                uint32_t word = word_next;
                word_next = wi + 32 <= w_last ? bits[wi + 32] : 0;
                if (wi == w_first) word &= ~0u << (xa & 31);
                if (wi == w_last && (xb & 31)) word &= (1u << (xb & 31)) - 1u;
                for (;;) {
                    uint32_t cand = word;
                    if (next_ok > wx) cand = next_ok - wx >= 32 ? 0 : word & (~0u << (next_ok - wx));
                    const uint32_t bal = __ballot_sync(FULLMASK, cand != 0);
                    if (!bal) break;
                    const uint32_t pos = __shfl_sync(FULLMASK, wx + __ffs(cand) - 1, __ffs(bal) - 1);
                    if (lane == 0) { if (n < P.per_buf_cap) out[n].f1_sample = pos - xa; else atomicOr(&P.ctl->overflow, 32u); }
                    n++;
                    next_ok = pos + AC_SKIP;
                }
            }
            n = min(n, P.per_buf_cap);
            __syncwarp();
            // decode the accepted replies, one per lane (the scan proved each of them well-formed)
            const uint32_t noise = P.noise[seg.first_buf + b];
            for (uint32_t k = lane; k < n; k += 32) {
                const uint32_t j = out[k].f1_sample;
                GlobalMag m;
                m.base = seg.base; m.lut = P.lut_full; m.d = (int64_t)b * seg.buf_len + j;
                m.zero_end = (seg.flags & SEG_HALO_ZERO) ? B200_TRAIL : 0; m.data_end = (int64_t)seg.npos + B200_TRAIL;
                m.is_mag = seg.flags & SEG_MAG;
                uint32_t f1_clock = 0, f1f2 = 0;
                ac_front(m, j, noise, &f1_clock, &f1f2);
                const uint32_t modeac = ac_bits(m, j, noise, f1_clock, f1f2);
                b200_modeac a;
                a.timestamp = seg.first_ts + (int64_t)b * seg.buf_len * 5 + (f1_clock + 87 * 14) / 5;          // demod_2400.c:745
                a.f1_sample = j; a.modeac = (uint16_t)modeac;
                a.buffer_idx = (uint16_t)(seg.first_buf + b - stream_first_buf);
                out[k] = a;
            }
            if (lane == 0) P.ac_count[seg.first_buf + b] = n;
        }
    }
}

// after the count prefix: receiver statistics (the only state Mode A/C touches; skipped with the rest of stage B on a failed run)
__global__ void modeac_stats_kernel(const AcWalkParams P, const uint32_t *prefix) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.n_streams || (P.ctl->overflow & (RUN_REPEAT_BITS | 32u))) return;
    const uint32_t sb = P.stream_seg_begin[s], se = P.stream_seg_begin[s + 1];
    if (sb == se) return;
    const uint32_t b0 = P.segs[sb].first_buf, b1 = P.segs[se - 1].first_buf + P.segs[se - 1].n_bufs;
    P.state[s].stats.demod_modeac += prefix[b1] - prefix[b0];
}

extern "C" int b200_prepare_modeac(void) {      // per device, from b200_demod_create
    return (int)cudaFuncSetAttribute(modeac_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AcSmem));
}

extern "C" int b200_launch_modeac(const AcScanParams *sp, const AcWalkParams *wp, int n_sm, void *stream) {
    if (sp->n_segs) {       // also for a run of empty buffers only (no tiles): the walk is what sets every buffer's reply count, zero included
        modeac_noise_kernel<<<min(sp->n_segs, 1024u), 32, 0, (cudaStream_t)stream>>>(*sp);
        if (sp->n_tiles) {
            // three of four scan tiles are skipped (AC_TILE = 4 scan tiles): an odd grid gives every block the same share of the fourth
            uint32_t grid = (uint32_t)n_sm * 2 - 1;       // odd and not more than one wave (2 CTAs per SM)
            if (grid > sp->n_tiles) grid = sp->n_tiles;
            modeac_scan_kernel<<<grid, AC_THREADS, sizeof(AcSmem), (cudaStream_t)stream>>>(*sp);
        }
        modeac_walk_kernel<<<wp->n_segs, 256, 0, (cudaStream_t)stream>>>(*wp);
    }
    return (int)cudaGetLastError();
}

extern "C" int b200_launch_modeac_stats(const AcWalkParams *wp, const uint32_t *prefix, void *stream) {
    modeac_stats_kernel<<<(wp->n_streams + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*wp, prefix);
    return (int)cudaGetLastError();
}

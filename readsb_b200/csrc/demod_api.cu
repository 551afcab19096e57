// demod_api.cu — the C ABI declared in include/b200_demod.h: context, device memory, the
// host-buffer (drop-in) path and the device-resident path around the kernels in demod_kernels.cu.
//
// Host side of the boundary (reference tree): a frontend's converter call + mag_buf hand-off
// (sdr_ifile.c:194-259, convert.h:34-39) becomes b200_demod_submit_iq_uc8; the decode thread's
// demodulate2400(buf) (readsb.c:871, demod_2400.h:38) becomes submit_mag_u16 / run / fetch.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.h"
#include "modes_tables.h"

#define API extern "C" __attribute__((visibility("default")))

static std::string g_create_error;

struct Pending {
    uint32_t n;          // new samples
    int64_t ts;
    size_t off;          // byte offset of data index 0 (mag) or of the first new sample (iq) in the stream's arena region
};

struct b200_demod_ctx {
    b200_demod_config cfg;
    int device = 0, n_sm = 148;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    std::string err;

    DeviceTables *d_tables = nullptr;
    uint16_t *d_lut_full = nullptr;
    StreamState *d_state = nullptr;
    RunCtl *d_ctl = nullptr, *h_ctl = nullptr;

    uint8_t *d_arena = nullptr;
    size_t stream_stride = 0;
    std::vector<std::vector<Pending>> pending;
    std::vector<uint8_t> kind;        // per stream this run: 0 none, 1 iq, 2 mag
    std::vector<uint8_t> halo_valid;  // iq streams: saved 326-sample tail is valid
    std::vector<size_t> cursor;       // append offset in the stream's arena region

    uint32_t seg_cap = 0, tile_cap = 0, buf_cap = 0, frame_cap = 0, rec_cap = 0;
    Segment *d_segs = nullptr, *h_segs = nullptr;
    uint32_t *d_tile_seg = nullptr, *h_tile_seg = nullptr;
    uint32_t *d_stream_seg_begin = nullptr, *h_stream_seg_begin = nullptr;
    uint32_t cached_tiles = 0;        // tile_seg on the device is valid for this many tiles (device-resident path)
    uint64_t cached_layout_key = 0;
    PosEntry *d_pos_pool = nullptr;
    Rec *d_rec_pool = nullptr;
    uint32_t *d_key_pool = nullptr;
    TileOut *d_tile_out = nullptr;
    BufAcc *d_buf_acc = nullptr, *h_buf_acc = nullptr;
    b200_buffer_result *d_buf_out = nullptr, *h_buf_out = nullptr;
    b200_frame *d_frames = nullptr, *d_packed = nullptr, *h_packed = nullptr;
    uint32_t *d_frame_count = nullptr, *d_frame_prefix = nullptr, *h_frame_prefix = nullptr;
    uint32_t *d_carry_src = nullptr, *h_carry_src = nullptr;
    uint8_t *d_scratch = nullptr;     // dense-input slow path arena, allocated on first need
    int *d_result = nullptr;

    // results of the last run
    uint32_t run_segs = 0, run_tiles = 0, run_bufs = 0, run_frames = 0;
    std::vector<uint32_t> stream_buf_begin;   // [n_streams+1] into h_buf_out
    cudaEvent_t ev[6] = {};
    float ms[5] = {0, 0, 0, 0, 0};
    uint32_t launches = 0;
};

static int fail(b200_demod_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CU(c, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(c, B200_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

template <class T> static cudaError_t dev_alloc(T **p, size_t n) { return cudaMalloc((void **)p, n * sizeof(T)); }
template <class T> static cudaError_t pin_alloc(T **p, size_t n) { return cudaHostAlloc((void **)p, n * sizeof(T), cudaHostAllocDefault); }

// Copies each IQ stream's last 326 samples to the front of its arena region (the mag_buf overlap copy,
// sdr_ifile.c:209-213), one block per stream; src[s] = byte offset of the tail, 0xffffffff = nothing to do.
__global__ void carry_halo_kernel(uint8_t *arena, size_t stride, const uint32_t *src, uint32_t n_streams) {
    const uint32_t s = blockIdx.x;
    if (s >= n_streams || src[s] == 0xffffffffu) return;
    uint8_t *region = arena + (size_t)s * stride;
    const uint16_t *from = reinterpret_cast<const uint16_t *>(region + src[s]);
    uint16_t *to = reinterpret_cast<uint16_t *>(region);
    for (uint32_t i = threadIdx.x; i < B200_TRAIL; i += blockDim.x) to[i] = from[i];
}

__global__ void init_state_kernel(StreamState *st, uint32_t n) {
    const uint32_t s = blockIdx.x;
    if (s >= n) return;
    for (uint32_t i = threadIdx.x; i < 2 * ICAO_CAP; i += blockDim.x) (&st[s].gen[0][0])[i] = ICAO_EMPTY;
    if (threadIdx.x == 0) {
        st[s].gen_count[0] = st[s].gen_count[1] = 0; st[s].active = 0; st[s].flip_armed = 0; st[s].next_flip_ms = 0;
        st[s].buffer_seq = 0; st[s].error = 0;
        memset(&st[s].stats, 0, sizeof(st[s].stats));
    }
}

API int b200_demod_abi_version(void) { return B200_DEMOD_ABI_VERSION; }

API const char *b200_demod_last_error(const b200_demod_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

API void *b200_demod_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
API void b200_demod_host_free(void *p) { if (p) cudaFreeHost(p); }

API int b200_demod_uc8_lut(uint16_t *out) {
    if (!out) return B200_E_INVAL;
    b200_build_uc8_lut(out);
    return B200_OK;
}

API void b200_demod_destroy(b200_demod_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_tables); cudaFree(c->d_lut_full); cudaFree(c->d_state); cudaFree(c->d_ctl); cudaFree(c->d_arena);
    cudaFree(c->d_segs); cudaFree(c->d_tile_seg); cudaFree(c->d_stream_seg_begin); cudaFree(c->d_pos_pool);
    cudaFree(c->d_rec_pool); cudaFree(c->d_key_pool); cudaFree(c->d_tile_out); cudaFree(c->d_buf_acc); cudaFree(c->d_buf_out);
    cudaFree(c->d_frames); cudaFree(c->d_packed); cudaFree(c->d_frame_count); cudaFree(c->d_frame_prefix);
    cudaFree(c->d_carry_src); cudaFree(c->d_result); cudaFree(c->d_scratch);
    cudaFreeHost(c->h_ctl); cudaFreeHost(c->h_segs); cudaFreeHost(c->h_tile_seg); cudaFreeHost(c->h_stream_seg_begin);
    cudaFreeHost(c->h_buf_acc); cudaFreeHost(c->h_buf_out); cudaFreeHost(c->h_packed); cudaFreeHost(c->h_frame_prefix);
    cudaFreeHost(c->h_carry_src);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

API int b200_demod_create(const b200_demod_config *cfg, b200_demod_ctx **out) {
    if (!cfg || !out || cfg->struct_size != sizeof(b200_demod_config)) return fail(nullptr, B200_E_INVAL, "bad config (struct_size)");
    if (cfg->n_streams == 0 || cfg->buf_samples == 0 || cfg->max_buffers_per_run == 0) return fail(nullptr, B200_E_INVAL, "n_streams, buf_samples and max_buffers_per_run must be > 0");
    if (cfg->nfix_crc < 0 || cfg->nfix_crc > 1) return fail(nullptr, B200_E_INVAL, "nfix_crc must be 0 or 1 (2-bit --aggressive tables are not on this path)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, B200_E_NODEV, "no CUDA device: this library has no CPU path");
    b200_demod_ctx *c = new b200_demod_ctx();
    c->cfg = *cfg;
    if (c->cfg.preamble_threshold == 0) c->cfg.preamble_threshold = B200_PREAMBLE_THRESHOLD_DEFAULT;
    if (c->cfg.icao_ttl_ms == 0) c->cfg.icao_ttl_ms = B200_ICAO_TTL_MS;
    int dev = cfg->device;
    if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) { delete c; return fail(nullptr, B200_E_NODEV, "cudaGetDevice failed"); }
    if (dev >= ndev) { delete c; return fail(nullptr, B200_E_INVAL, "device %d out of range", dev); }
    c->device = dev;
#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(nullptr, B200_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); b200_demod_destroy(c); return e_ == cudaErrorMemoryAllocation ? B200_E_NOMEM : B200_E_CUDA; } } while (0)
    CUC(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CUC(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) { fail(nullptr, B200_E_NODEV, "device %d is sm_%d%d; the kernels are built for sm_100a only", dev, prop.major, prop.minor); b200_demod_destroy(c); return B200_E_NODEV; }
    c->n_sm = prop.multiProcessorCount;
    CUC(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    for (auto &e : c->ev) CUC(cudaEventCreate(&e));

    const uint32_t S = cfg->n_streams, K = cfg->max_buffers_per_run, BUF = cfg->buf_samples;
    // tables
    {
        DeviceTables *t = new DeviceTables();
        std::vector<uint16_t> lut(65536);
        if (b200_build_tables(t, lut.data()) != 0) { delete t; fail(nullptr, B200_E_INVAL, "syndrome hash construction failed"); b200_demod_destroy(c); return B200_E_INVAL; }
        // the table must be bit-identical to the reference's: CRC32 of the LE table
        if (b200_crc32_ieee(lut.data(), 65536 * 2) != 0x8e9d21e1u) { delete t; fail(nullptr, B200_E_INVAL, "UC8 lookup table does not match the reference arithmetic (host FP contraction?)"); b200_demod_destroy(c); return B200_E_INVAL; }
        CUC(dev_alloc(&c->d_tables, 1));
        CUC(dev_alloc(&c->d_lut_full, 65536));
        CUC(cudaMemcpy(c->d_tables, t, sizeof(DeviceTables), cudaMemcpyHostToDevice));
        CUC(cudaMemcpy(c->d_lut_full, lut.data(), 65536 * 2, cudaMemcpyHostToDevice));
        delete t;
    }
    CUC(dev_alloc(&c->d_state, S));
    init_state_kernel<<<S, 256, 0, c->stream>>>(c->d_state, S);
    CUC(cudaGetLastError());
    CUC(dev_alloc(&c->d_ctl, 1));
    CUC(pin_alloc(&c->h_ctl, 1));
    CUC(dev_alloc(&c->d_result, 1));

    // arena for host submits: [326-sample halo][K buffers, each with room for its own halo when magnitudes are submitted]
    c->stream_stride = (((size_t)K * ((size_t)BUF + B200_TRAIL + 16) * 2 + 1024) + 255) & ~(size_t)255;
    CUC(cudaMalloc((void **)&c->d_arena, c->stream_stride * S + 256));
    CUC(cudaMemsetAsync(c->d_arena, 0, c->stream_stride * S + 256, c->stream));
    c->pending.resize(S); c->kind.assign(S, 0); c->halo_valid.assign(S, 0); c->cursor.assign(S, 0);

    const uint32_t tiles_per_buf = (BUF + 16 + SCAN_TILE - 1) / SCAN_TILE + 1;
    c->seg_cap = S * K;
    c->tile_cap = S * K * tiles_per_buf;
    c->buf_cap = S * K;
    c->frame_cap = K * (BUF / 113 + 2);
    const size_t positions = (size_t)S * K * BUF;
    c->rec_cap = (uint32_t)std::max<size_t>(65536, positions / 16);
    CUC(dev_alloc(&c->d_segs, c->seg_cap)); CUC(pin_alloc(&c->h_segs, c->seg_cap));
    CUC(dev_alloc(&c->d_tile_seg, c->tile_cap)); CUC(pin_alloc(&c->h_tile_seg, c->tile_cap));
    CUC(dev_alloc(&c->d_stream_seg_begin, S + 1)); CUC(pin_alloc(&c->h_stream_seg_begin, S + 1));
    CUC(dev_alloc(&c->d_pos_pool, (size_t)c->tile_cap * SCAN_TILE));
    CUC(dev_alloc(&c->d_rec_pool, c->rec_cap));
    CUC(dev_alloc(&c->d_key_pool, c->rec_cap));
    CUC(dev_alloc(&c->d_tile_out, c->tile_cap));
    CUC(dev_alloc(&c->d_buf_acc, c->buf_cap)); CUC(pin_alloc(&c->h_buf_acc, c->buf_cap));
    CUC(dev_alloc(&c->d_buf_out, c->buf_cap)); CUC(pin_alloc(&c->h_buf_out, c->buf_cap));
    CUC(dev_alloc(&c->d_frames, (size_t)S * c->frame_cap));
    CUC(dev_alloc(&c->d_packed, (size_t)S * c->frame_cap)); CUC(pin_alloc(&c->h_packed, (size_t)S * c->frame_cap));
    CUC(dev_alloc(&c->d_frame_count, S));
    CUC(cudaMemsetAsync(c->d_frame_count, 0, S * 4, c->stream));
    CUC(dev_alloc(&c->d_frame_prefix, S + 1)); CUC(pin_alloc(&c->h_frame_prefix, S + 1));
    CUC(dev_alloc(&c->d_carry_src, S)); CUC(pin_alloc(&c->h_carry_src, S));
    c->stream_buf_begin.assign(S + 1, 0);
    memset(c->h_frame_prefix, 0, (S + 1) * 4);
    CUC(cudaStreamSynchronize(c->stream));
#undef CUC
    *out = c;
    return B200_OK;
}

// ---- submits ---------------------------------------------------------------------------------
static int submit_common(b200_demod_ctx *c, uint32_t s, const void *host, uint32_t n, int64_t ts, bool mag) {
    if (!c) return B200_E_INVAL;
    if (s >= c->cfg.n_streams || (!host && n)) return fail(c, B200_E_INVAL, "bad stream or buffer");
    if (n > c->cfg.buf_samples) return fail(c, B200_E_INVAL, "buffer of %u samples exceeds buf_samples=%u", n, c->cfg.buf_samples);
    if (c->pending[s].size() >= c->cfg.max_buffers_per_run) return fail(c, B200_E_STATE, "stream %u already has max_buffers_per_run buffers queued", s);
    const uint8_t want = mag ? 2 : 1;
    if (c->kind[s] && c->kind[s] != want) return fail(c, B200_E_STATE, "stream %u mixes IQ and magnitude submits in one run", s);
    CU(c, cudaSetDevice(c->device));
    if (!c->kind[s]) { c->kind[s] = want; c->cursor[s] = mag ? 0 : (size_t)B200_TRAIL * 2; }
    uint8_t *region = c->d_arena + (size_t)s * c->stream_stride;
    Pending p;
    p.n = n; p.ts = ts;
    if (mag) {
        p.off = (c->cursor[s] + 15) & ~(size_t)15;
        const size_t bytes = ((size_t)n + B200_TRAIL) * 2;
        CU(c, cudaMemcpyAsync(region + p.off, host, bytes, cudaMemcpyHostToDevice, c->stream));
        c->cursor[s] = p.off + bytes;
    } else {
        p.off = c->cursor[s];
        if (n) CU(c, cudaMemcpyAsync(region + p.off, host, (size_t)n * 2, cudaMemcpyHostToDevice, c->stream));
        c->cursor[s] = p.off + (size_t)n * 2;
    }
    c->pending[s].push_back(p);
    return B200_OK;
}

API int b200_demod_set_stream(b200_demod_ctx *c, void *cuda_stream) {
    if (!c) return B200_E_INVAL;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return B200_OK;
}

API int b200_demod_submit_iq_uc8_strided(b200_demod_ctx *c, uint32_t first, uint32_t ns, const uint8_t *iq, uint64_t host_stride,
                                         uint32_t n_buffers, uint32_t buf_len, int64_t ts) {
    if (!c || !iq) return B200_E_INVAL;
    if (ns == 0 || first + ns > c->cfg.n_streams || n_buffers == 0 || buf_len == 0 || buf_len > c->cfg.buf_samples) return fail(c, B200_E_INVAL, "bad stream range or buffer length");
    if (n_buffers > 1 && buf_len != c->cfg.buf_samples) return fail(c, B200_E_INVAL, "several buffers per call need buf_len == buf_samples");
    const size_t row = (size_t)n_buffers * buf_len * 2;
    if (host_stride < row) return fail(c, B200_E_INVAL, "host_stride_bytes smaller than one stream's data");
    for (uint32_t s = first; s < first + ns; s++) {
        if (c->kind[s] == 2) return fail(c, B200_E_STATE, "stream %u mixes IQ and magnitude submits in one run", s);
        if (c->pending[s].size() + n_buffers > c->cfg.max_buffers_per_run) return fail(c, B200_E_STATE, "stream %u would exceed max_buffers_per_run", s);
        if (c->kind[s] && c->cursor[s] != c->cursor[first]) return fail(c, B200_E_STATE, "strided submit needs all streams of the range at the same fill level");
    }
    CU(c, cudaSetDevice(c->device));
    const size_t off0 = c->kind[first] ? c->cursor[first] : (size_t)B200_TRAIL * 2;
    CU(c, cudaMemcpy2DAsync(c->d_arena + (size_t)first * c->stream_stride + off0, c->stream_stride, iq, host_stride, row, ns,
                            cudaMemcpyHostToDevice, c->stream));
    for (uint32_t s = first; s < first + ns; s++) {
        c->kind[s] = 1;
        for (uint32_t b = 0; b < n_buffers; b++) {
            Pending p; p.n = buf_len; p.ts = ts + (int64_t)b * buf_len * 5; p.off = off0 + (size_t)b * buf_len * 2;
            c->pending[s].push_back(p);
        }
        c->cursor[s] = off0 + row;
    }
    return B200_OK;
}

API int b200_demod_submit_iq_uc8(b200_demod_ctx *c, uint32_t s, const uint8_t *iq, uint32_t n, int64_t ts) { return submit_common(c, s, iq, n, ts, false); }
API int b200_demod_submit_mag_u16(b200_demod_ctx *c, uint32_t s, const uint16_t *data, uint32_t n, int64_t ts) { return submit_common(c, s, data, n, ts, true); }

// ---- the pipeline ------------------------------------------------------------------------------
static void add_segment(b200_demod_ctx *c, uint32_t &nseg, uint32_t &ntile, uint32_t &nbuf, uint32_t stream, const uint8_t *base,
                        uint32_t npos, uint32_t buf_len, uint32_t n_bufs, uint32_t flags, int64_t first_ts) {
    Segment &g = c->h_segs[nseg];
    memset(&g, 0, sizeof g);
    g.base = base; g.first_ts = first_ts; g.npos = npos; g.buf_len = buf_len ? buf_len : 1;
    g.lead = (uint32_t)(((uintptr_t)base & 15) / 2); g.flags = flags; g.stream = stream;
    g.first_buf = nbuf; g.n_bufs = n_bufs; g.tile_begin = ntile;
    g.n_tiles = npos ? (g.lead + npos + SCAN_TILE - 1) / SCAN_TILE : 0;
    for (uint32_t t = 0; t < g.n_tiles; t++) c->h_tile_seg[ntile + t] = nseg;
    ntile += g.n_tiles; nbuf += n_bufs; nseg++;
}

static int execute(b200_demod_ctx *c, uint32_t nseg, uint32_t ntile, uint32_t nbuf, bool upload_tiles) {
    const uint32_t S = c->cfg.n_streams;
    c->run_segs = nseg; c->run_tiles = ntile; c->run_bufs = nbuf; c->run_frames = 0;
    CU(c, cudaMemcpyAsync(c->d_segs, c->h_segs, nseg * sizeof(Segment), cudaMemcpyHostToDevice, c->stream));
    if (upload_tiles && ntile) CU(c, cudaMemcpyAsync(c->d_tile_seg, c->h_tile_seg, ntile * 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->d_stream_seg_begin, c->h_stream_seg_begin, (S + 1) * 4, cudaMemcpyHostToDevice, c->stream));

    for (int attempt = 0; attempt < 6; attempt++) {
        memset(c->h_ctl, 0, sizeof(RunCtl));
        c->h_ctl->rec_cap = c->rec_cap;
        CU(c, cudaMemcpyAsync(c->d_ctl, c->h_ctl, sizeof(RunCtl), cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemsetAsync(c->d_buf_acc, 0, (size_t)nbuf * sizeof(BufAcc), c->stream));
        c->launches = 0;

        ScanParams sp;
        sp.segs = c->d_segs; sp.tile_seg = c->d_tile_seg; sp.n_tiles = ntile; sp.pos_pool = c->d_pos_pool; sp.rec_pool = c->d_rec_pool; sp.key_pool = c->d_key_pool;
        sp.tile_out = c->d_tile_out; sp.buf_acc = c->d_buf_acc; sp.ctl = c->d_ctl; sp.thr = c->cfg.preamble_threshold;
        sp.nfix = c->cfg.nfix_crc; sp.fixdf = c->cfg.fix_df; sp.scratch = c->d_scratch;
        // demod_2400.c:112-127
        sp.short_set = (1u << 0) | (1u << 4) | (1u << 5) | (1u << 11);
        sp.long_set = (1u << 16) | (1u << 17) | (1u << 18) | (1u << 20) | (1u << 21);
        if (sp.nfix && sp.fixdf) for (int b = 0; b < 5; b++) sp.long_set |= 1u << (17 ^ (1 << b));
        CU(c, cudaEventRecord(c->ev[1], c->stream));
        if (ntile) { int r = b200_launch_scan(&sp, c->d_tables, c->n_sm, c->stream); if (r) return fail(c, B200_E_CUDA, "scan launch: %s", cudaGetErrorString((cudaError_t)r)); c->launches++; }
        CU(c, cudaEventRecord(c->ev[2], c->stream));

        ResolveParams rp;
        rp.segs = c->d_segs; rp.stream_seg_begin = c->d_stream_seg_begin; rp.n_streams = S; rp.pos_pool = c->d_pos_pool;
        rp.rec_pool = c->d_rec_pool; rp.key_pool = c->d_key_pool; rp.tile_out = c->d_tile_out; rp.buf_acc = c->d_buf_acc; rp.buf_out = c->d_buf_out;
        rp.state = c->d_state; rp.frames = c->d_frames; rp.frame_count = c->d_frame_count; rp.frame_cap = c->frame_cap;
        rp.ctl = c->d_ctl; rp.ttl_ms = c->cfg.icao_ttl_ms;
        { int r = b200_launch_resolve(&rp, c->stream); if (r) return fail(c, B200_E_CUDA, "resolve launch: %s", cudaGetErrorString((cudaError_t)r)); c->launches++; }
        CU(c, cudaEventRecord(c->ev[3], c->stream));

        FinalizeParams fp;
        fp.segs = c->d_segs; fp.stream_seg_begin = c->d_stream_seg_begin; fp.n_streams = S; fp.frames = c->d_frames;
        fp.frame_count = c->d_frame_count; fp.frame_prefix = c->d_frame_prefix; fp.frame_cap = c->frame_cap; fp.packed = c->d_packed;
        fp.buf_acc = c->d_buf_acc; fp.state = c->d_state; fp.lut_full = c->d_lut_full; fp.rec_pool = c->d_rec_pool;
        { int r = b200_launch_finalize(&fp, c->d_frame_prefix, c->d_ctl, c->stream); if (r) return fail(c, B200_E_CUDA, "finalize launch: %s", cudaGetErrorString((cudaError_t)r)); c->launches += 2; }
        CU(c, cudaEventRecord(c->ev[4], c->stream));

        CU(c, cudaMemcpyAsync(c->h_ctl, c->d_ctl, sizeof(RunCtl), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaMemcpyAsync(c->h_frame_prefix, c->d_frame_prefix, (S + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
        if (nbuf) {
            CU(c, cudaMemcpyAsync(c->h_buf_out, c->d_buf_out, (size_t)nbuf * sizeof(b200_buffer_result), cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaMemcpyAsync(c->h_buf_acc, c->d_buf_acc, (size_t)nbuf * sizeof(BufAcc), cudaMemcpyDeviceToHost, c->stream));
        }
        CU(c, cudaStreamSynchronize(c->stream));
        if (c->h_ctl->overflow & 1u) {        // record pool too small: stage A is stateless, stage B did nothing -> grow and redo
            const uint32_t need = c->h_ctl->rec_alloc + c->h_ctl->rec_alloc / 4 + 65536;
            cudaFree(c->d_rec_pool); c->d_rec_pool = nullptr;
            cudaFree(c->d_key_pool); c->d_key_pool = nullptr;
            c->rec_cap = need;
            if (dev_alloc(&c->d_rec_pool, c->rec_cap) != cudaSuccess || dev_alloc(&c->d_key_pool, c->rec_cap) != cudaSuccess) return fail(c, B200_E_NOMEM, "cannot grow the record pool to %u records", need);
            continue;
        }
        if ((c->h_ctl->overflow & 2u) && !c->d_scratch) {   // a tile denser than the shared-memory queues: give the kernel its slow-path arena
            const int grid = b200_scan_grid(c->n_sm);
            if (grid <= 0 || cudaMalloc((void **)&c->d_scratch, (size_t)grid * SCAN_SCRATCH_BYTES) != cudaSuccess)
                return fail(c, B200_E_NOMEM, "cannot allocate the dense-input scratch arena");
            continue;
        }
        break;
    }
    if (c->h_ctl->overflow & 1u) return fail(c, B200_E_NOMEM, "record pool still too small after regrowth");
    if (c->h_ctl->overflow & 2u) return fail(c, B200_E_OVERFLOW, "a tile exceeded the in-kernel candidate capacity (input denser than the kernel is sized for)");
    if (c->h_ctl->overflow & 4u) return fail(c, B200_E_OVERFLOW, "per-stream frame capacity exceeded");
    if (c->h_ctl->overflow & 8u) return fail(c, B200_E_OVERFLOW, "a receiver's ICAO filter generation is full (%u addresses)", ICAO_CAP / 2);
    const uint32_t total = c->h_frame_prefix[S];
    c->run_frames = total;
    if (total) CU(c, cudaMemcpyAsync(c->h_packed, c->d_packed, (size_t)total * sizeof(b200_frame), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaEventRecord(c->ev[5], c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (uint32_t b = 0; b < nbuf; b++) {
        c->h_buf_out[b].sum_level = c->h_buf_acc[b].sum_level;
        c->h_buf_out[b].sum_power = c->h_buf_acc[b].sum_power;
        c->h_buf_out[b].sum_signal_power = c->h_buf_acc[b].sum_signal_power;
    }
    cudaEventElapsedTime(&c->ms[1], c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&c->ms[2], c->ev[2], c->ev[3]);
    cudaEventElapsedTime(&c->ms[0], c->ev[1], c->ev[5]);
    cudaEventElapsedTime(&c->ms[4], c->ev[4], c->ev[5]);
    c->ms[3] = 0;
    return B200_OK;
}

API int b200_demod_run(b200_demod_ctx *c) {
    if (!c) return B200_E_INVAL;
    CU(c, cudaSetDevice(c->device));
    const uint32_t S = c->cfg.n_streams, BUF = c->cfg.buf_samples;
    uint32_t nseg = 0, ntile = 0, nbuf = 0;
    for (uint32_t s = 0; s < S; s++) {
        c->h_stream_seg_begin[s] = nseg;
        c->stream_buf_begin[s] = nbuf;
        c->h_carry_src[s] = 0xffffffffu;
        const auto &pl = c->pending[s];
        if (pl.empty()) continue;
        const uint8_t *region = c->d_arena + (size_t)s * c->stream_stride;
        if (c->kind[s] == 2) {
            for (const Pending &p : pl) add_segment(c, nseg, ntile, nbuf, s, region + p.off, p.n, p.n, 1, SEG_MAG, p.ts);
        } else {
            // consecutive full buffers with contiguous timestamps form one segment; a partial buffer ends it
            size_t i = 0;
            bool halo_ok = c->halo_valid[s];
            while (i < pl.size()) {
                size_t j = i;
                uint32_t npos = 0;
                for (;;) {
                    npos += pl[j].n;
                    const bool more = j + 1 < pl.size() && pl[j].n == BUF && pl[j + 1].ts == pl[j].ts + (int64_t)BUF * 5;
                    if (!more) break;
                    j++;
                }
                const uint32_t nb = (uint32_t)(j - i + 1);
                add_segment(c, nseg, ntile, nbuf, s, region + pl[i].off - (size_t)B200_TRAIL * 2, npos, nb > 1 ? BUF : pl[i].n, nb,
                            halo_ok ? 0 : SEG_HALO_ZERO, pl[i].ts);
                halo_ok = pl[j].n >= B200_TRAIL;      // sdr_ifile.c:209-213
                i = j + 1;
            }
            c->halo_valid[s] = halo_ok;
            if (halo_ok) c->h_carry_src[s] = (uint32_t)(c->cursor[s] - (size_t)B200_TRAIL * 2);
        }
    }
    c->h_stream_seg_begin[S] = nseg;
    c->stream_buf_begin[S] = nbuf;
    c->cached_tiles = 0;
    int rc = execute(c, nseg, ntile, nbuf, true);
    // next run: move each IQ stream's tail to the front of its region
    if (rc == B200_OK) {
        cudaMemcpyAsync(c->d_carry_src, c->h_carry_src, S * 4, cudaMemcpyHostToDevice, c->stream);
        carry_halo_kernel<<<S, 128, 0, c->stream>>>(c->d_arena, c->stream_stride, c->d_carry_src, S);
        c->launches++;
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(c, B200_E_CUDA, "halo carry failed");
    }
    for (uint32_t s = 0; s < S; s++) { c->pending[s].clear(); c->kind[s] = 0; c->cursor[s] = 0; }
    return rc;
}

API int b200_demod_run_device_uc8(b200_demod_ctx *c, const uint8_t *d_iq, uint64_t stride, uint32_t n_buffers, uint32_t buf_len,
                                  int continues, int64_t first_ts) {
    if (!c || !d_iq) return B200_E_INVAL;
    const uint32_t S = c->cfg.n_streams;
    if (((uintptr_t)d_iq & 15) || (stride & 15) || (buf_len & 7)) return fail(c, B200_E_INVAL, "d_iq and stream_stride_bytes must be 16-byte aligned and buf_len a multiple of 8");
    if (n_buffers == 0 || n_buffers > c->cfg.max_buffers_per_run || buf_len == 0 || buf_len > c->cfg.buf_samples) return fail(c, B200_E_INVAL, "n_buffers/buf_len exceed the context's configuration");
    if ((uint64_t)n_buffers * buf_len * 2 > stride && S > 1) return fail(c, B200_E_INVAL, "stream_stride_bytes smaller than one stream's data");
    CU(c, cudaSetDevice(c->device));
    uint32_t nseg = 0, ntile = 0, nbuf = 0;
    for (uint32_t s = 0; s < S; s++) {
        c->h_stream_seg_begin[s] = nseg;
        c->stream_buf_begin[s] = nbuf;
        add_segment(c, nseg, ntile, nbuf, s, d_iq + (size_t)s * stride - (size_t)B200_TRAIL * 2, n_buffers * buf_len, buf_len, n_buffers,
                    continues ? 0 : SEG_HALO_ZERO, first_ts);
    }
    c->h_stream_seg_begin[S] = nseg;
    c->stream_buf_begin[S] = nbuf;
    // the tile -> segment table only depends on the layout; skip its upload when nothing changed
    const uint64_t key = ((uint64_t)n_buffers << 40) ^ ((uint64_t)buf_len << 8) ^ (((uintptr_t)d_iq >> 4) & 15) ^ (stride << 20);
    const bool upload = !(c->cached_tiles == ntile && c->cached_layout_key == key);
    c->cached_tiles = ntile; c->cached_layout_key = key;
    return execute(c, nseg, ntile, nbuf, upload);
}

// ---- results -----------------------------------------------------------------------------------
API int b200_demod_total_frames(b200_demod_ctx *c, uint64_t *n) { if (!c || !n) return B200_E_INVAL; *n = c->run_frames; return B200_OK; }

API int b200_demod_frame_count(b200_demod_ctx *c, uint32_t s, uint32_t *n) {
    if (!c || !n || s >= c->cfg.n_streams) return B200_E_INVAL;
    *n = c->h_frame_prefix[s + 1] - c->h_frame_prefix[s];
    return B200_OK;
}

API int b200_demod_fetch(b200_demod_ctx *c, uint32_t s, b200_frame *out, uint32_t cap, uint32_t *n) {
    if (!c || !n || s >= c->cfg.n_streams) return B200_E_INVAL;
    const uint32_t cnt = c->h_frame_prefix[s + 1] - c->h_frame_prefix[s];
    *n = cnt;
    if (cnt > cap) return fail(c, B200_E_OVERFLOW, "stream %u has %u frames, output holds %u", s, cnt, cap);
    if (cnt) memcpy(out, c->h_packed + c->h_frame_prefix[s], (size_t)cnt * sizeof(b200_frame));
    return B200_OK;
}

API int b200_demod_buffer_results(b200_demod_ctx *c, uint32_t s, b200_buffer_result *out, uint32_t cap, uint32_t *n) {
    if (!c || !n || s >= c->cfg.n_streams) return B200_E_INVAL;
    const uint32_t b0 = c->stream_buf_begin[s], cnt = c->stream_buf_begin[s + 1] - b0;
    *n = cnt;
    if (cnt > cap) return fail(c, B200_E_OVERFLOW, "stream %u has %u buffer results, output holds %u", s, cnt, cap);
    if (cnt) memcpy(out, c->h_buf_out + b0, (size_t)cnt * sizeof(b200_buffer_result));
    return B200_OK;
}

API int b200_demod_get_stats(b200_demod_ctx *c, uint32_t s, b200_demod_stats *out) {
    if (!c || !out || s >= c->cfg.n_streams) return B200_E_INVAL;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemcpyAsync(out, &c->d_state[s].stats, sizeof(b200_demod_stats), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return B200_OK;
}

API int b200_demod_last_timing(b200_demod_ctx *c, float ms[5], uint32_t *launches) {
    if (!c) return B200_E_INVAL;
    if (ms) memcpy(ms, c->ms, sizeof(c->ms));
    if (launches) *launches = c->launches;
    return B200_OK;
}

// Instrumentation for tests: out[0] = tiles of the last run, [1] = sum of PosEntry counts, [2] = sum of Rec counts,
// [3] = RunCtl.rec_alloc, [4] = RunCtl.overflow, [5] = segments, [6] = buffers, [7] = frames.
API int b200_demod_debug_counters(b200_demod_ctx *c, uint64_t out[8]) {
    if (!c || !out) return B200_E_INVAL;
    CU(c, cudaSetDevice(c->device));
    std::vector<TileOut> t(c->run_tiles);
    if (c->run_tiles) CU(c, cudaMemcpy(t.data(), c->d_tile_out, c->run_tiles * sizeof(TileOut), cudaMemcpyDeviceToHost));
    uint64_t np = 0, nr = 0;
    for (auto &x : t) { np += x.n_pos; nr += x.n_rec; }
    out[0] = c->run_tiles; out[1] = np; out[2] = nr; out[3] = c->h_ctl->rec_alloc; out[4] = c->h_ctl->overflow;
    out[5] = c->run_segs; out[6] = c->run_bufs; out[7] = c->run_frames;
    return B200_OK;
}

// ---- ICAO filter control -------------------------------------------------------------------------
static int icao_op(b200_demod_ctx *c, uint32_t s, int op, uint32_t addr, int *result) {
    if (!c || s >= c->cfg.n_streams) return B200_E_INVAL;
    CU(c, cudaSetDevice(c->device));
    int r = b200_launch_icao_op(c->d_state, s, op, addr & 0xffffffu, c->d_result, c->stream);
    if (r) return fail(c, B200_E_CUDA, "icao op launch: %s", cudaGetErrorString((cudaError_t)r));
    int host = 0;
    CU(c, cudaMemcpyAsync(&host, c->d_result, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (result) *result = host;
    if (op == 0 && host < 0) return fail(c, B200_E_OVERFLOW, "ICAO filter generation full");
    return B200_OK;
}
API int b200_demod_icao_add(b200_demod_ctx *c, uint32_t s, uint32_t addr) { return icao_op(c, s, 0, addr, nullptr); }
API int b200_demod_icao_test(b200_demod_ctx *c, uint32_t s, uint32_t addr, int *present) { return icao_op(c, s, 1, addr, present); }
API int b200_demod_icao_expire(b200_demod_ctx *c, uint32_t s) { return icao_op(c, s, 2, 0, nullptr); }
API int b200_demod_icao_reset(b200_demod_ctx *c, uint32_t s) { return icao_op(c, s, 3, 0, nullptr); }

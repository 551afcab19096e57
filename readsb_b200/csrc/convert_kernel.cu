// convert_kernel.cu — the float-path IQ converters: convert_sc16_nodc convert.c:212-250 and
// convert_sc16q11_nodc convert.c:329-367 (the default build has no SC16Q11 table), for bladeRF / Pluto / Soapy style input.
//
//   sc16_mag_kernel   data parallel: int16 I, Q -> uint16 magnitude, exactly the reference's float expression
//                     (I / 32768 or I / 2048 is an exact scaling, the two products and their sum are rounded separately —
//                     the library is built with --fmad=false —, sqrtf is correctly rounded, * 65535 + 0.5 are two operations)
//   sc16_sum_kernel   the reference also returns mean_level / mean_power from FLOAT accumulators that are added to in
//                     sample order (convert.c:225,238-239); a float sum is not associative, so one warp per buffer keeps the
//                     order: the lanes compute 32 samples' magnitude / power in parallel, then every lane adds the 32 values
//                     in sample order (shuffles), all lanes holding the same running sums.
//
// The magnitudes go into the receiver's arena like a magnitude hand-off, and the Mode S pipeline runs on them unchanged.
#include "common.h"
#include "device_utils.cuh"

__device__ __forceinline__ void sc16_sample(uint32_t iq, float scale, float &mag, float &magsq) {
    const float fI = (float)(int16_t)(iq & 0xffffu) * scale, fQ = (float)(int16_t)(iq >> 16) * scale;   // I / 32768.0f resp. I / 2048.0f: exact
    const float a = fI * fI, b = fQ * fQ;
    magsq = a + b;
    if (magsq > 1.0f) magsq = 1.0f;
    mag = sqrtf(magsq);
}

__global__ void __launch_bounds__(256) sc16_mag_kernel(const uint32_t *__restrict__ iq, uint16_t *__restrict__ mag_out, uint32_t n, float scale) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float mag, magsq;
        sc16_sample(iq[i], scale, mag, magsq);
        const float scaled = mag * 65535.0f;
        mag_out[i] = (uint16_t)(scaled + 0.5f);
    }
}

__global__ void __launch_bounds__(32) sc16_sum_kernel(const uint32_t *__restrict__ iq, uint32_t n, float scale, float2 *out) {
    const uint32_t lane = threadIdx.x;
    float sum_level = 0.0f, sum_power = 0.0f;
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
        float mag = 0.0f, magsq = 0.0f;
        if (i0 + lane < n) sc16_sample(iq[i0 + lane], scale, mag, magsq);
        const uint32_t cnt = min(32u, n - i0);
        if (cnt == 32) {
#pragma unroll
            for (int k = 0; k < 32; k++) { sum_power += __shfl_sync(FULLMASK, magsq, k); sum_level += __shfl_sync(FULLMASK, mag, k); }
        } else {
            for (uint32_t k = 0; k < cnt; k++) { sum_power += __shfl_sync(FULLMASK, magsq, k); sum_level += __shfl_sync(FULLMASK, mag, k); }
        }
    }
    if (lane == 0) *out = make_float2(sum_level, sum_power);
}

extern "C" int b200_launch_sc16_convert(const void *d_iq, uint16_t *d_mag, uint32_t n, int q11, float2 *d_sums, int n_sm, void *stream) {
    if (!n) return 0;
    const float scale = q11 ? 1.0f / 2048.0f : 1.0f / 32768.0f;
    uint32_t grid = (n + 255) / 256;
    if (grid > (uint32_t)n_sm * 4) grid = (uint32_t)n_sm * 4;
    sc16_mag_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t *>(d_iq), d_mag, n, scale);
    sc16_sum_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t *>(d_iq), n, scale, d_sums);
    return (int)cudaGetLastError();
}

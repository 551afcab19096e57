// tests/emu/selftest/selftest.cu — TEST INFRASTRUCTURE: kernels that exercise every primitive the emulator provides (collectives,
// block barrier, integer intrinsics, the inline-PTX stand-ins, atomics, dynamic / static shared memory) with results a test can
// predict from the CUDA / PTX definitions.  Built by build_emu.py's rewrite like the product sources; also valid CUDA.
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t st_prmt(uint32_t a, uint32_t b, uint32_t sel) { uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d; }
__device__ __forceinline__ int st_dp2a_lo(uint32_t a, uint32_t b, int c) { int d; asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int st_dp2a_hi(uint32_t a, uint32_t b, int c) { int d; asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

// out[lane * 16 + k] for one warp; in[lane] = per-lane input
__global__ void st_warp_kernel(const uint32_t *in, uint32_t *out) {
    const uint32_t lane = threadIdx.x & 31, v = in[lane];
    uint32_t *o = out + lane * 16;
    o[0] = __shfl_sync(0xffffffffu, v, (lane * 7 + 3) & 31);
    o[1] = __shfl_up_sync(0xffffffffu, v, 3);
    o[2] = __shfl_down_sync(0xffffffffu, v, 5);
    o[3] = __shfl_xor_sync(0xffffffffu, v, 9);
    o[4] = __ballot_sync(0xffffffffu, (v >> 3) & 1u);
    o[5] = __reduce_add_sync(0xffffffffu, v & 0xffffu);
    o[6] = __popc(v) | ((uint32_t)__ffs((int)v) << 8) | ((uint32_t)__clz((int)v) << 16);
    o[7] = __brev(v);
    o[8] = __funnelshift_r(v, ~v, lane);
    o[9] = __funnelshift_l(v, ~v, lane);
    o[10] = __byte_perm(v, ~v, 0x5410 + (lane & 3) + ((lane & 4) << 2) + ((lane & 24) << 5));     // selector nibbles stay in 0..7
    o[11] = st_prmt(v, 0, 0xba98);
    o[12] = (uint32_t)st_dp2a_lo(v, 0x00030feeu, -7);
    o[13] = (uint32_t)st_dp2a_hi(v, 0xff14f1fcu, 11);
    // a divergent section followed by a converged collective: lanes must reconverge at the __syncwarp
    uint32_t acc = 0;
    for (uint32_t i = 0; i < (lane & 3u); i++) acc += i + lane;
    __syncwarp();
    o[14] = __shfl_sync(0xffffffffu, acc, 31 - lane);
    unsigned long long w = 5; asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w) : "r"(v), "r"(v));
    o[15] = (uint32_t)(w >> 32) ^ (uint32_t)w;
}

// block-level: dynamic + static shared memory, __syncthreads, shared and global atomics, cp.async
__global__ void st_block_kernel(const uint32_t *in, uint32_t *out, uint32_t *counter) {
    extern __shared__ uint4 st_smem[];
    __shared__ uint32_t total;
    uint32_t *buf = reinterpret_cast<uint32_t *>(st_smem);
    const uint32_t tid = threadIdx.x, n = blockDim.x;
    if (tid == 0) total = 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(&buf[tid])), "l"(&in[blockIdx.x * n + tid]));
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const uint32_t mine = buf[n - 1 - tid];              // written by another thread (another warp for n > 32)
    atomicAdd(&total, mine & 0xffu);
    __syncthreads();
    out[blockIdx.x * n + tid] = mine + total;
    if (tid == 0) atomicAdd(counter, total);
}

extern "C" __attribute__((visibility("default"))) int b200_emu_selftest(const uint32_t *in, uint32_t *out_warp, uint32_t *out_block, uint32_t *counter, uint32_t n_blocks, uint32_t n_threads) {
    st_warp_kernel<<<1, 32>>>(in, out_warp);
    st_block_kernel<<<n_blocks, n_threads, n_threads * 4>>>(in, out_block, counter);
    return (int)cudaGetLastError();
}

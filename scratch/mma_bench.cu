// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128 single CTA, K=16) issued back to back from resident shared
// memory tiles, as a function of the instruction width N and of the number of independent TMEM accumulators the
// instruction stream alternates between.  Answers: is there a fixed bubble between dependent accumulations into one tile?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I riffusion-hobby_b200/csrc scratch/mma_bench.cu -o scratch/mma_bench.bin
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "rf_tc.cuh"

template <int N, int NACC, int KB_STAGES>
__global__ void __launch_bounds__(128, 1) k_bench(int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    // KB_STAGES distinct k-blocks of A (128x64) and B (Nx64) so that operand fetch does not always hit the same lines
    uint8_t* sA = smem;
    uint8_t* sB = smem + KB_STAGES * 16384;
    for (int i = threadIdx.x; i < (KB_STAGES * (16384 + N * 128)) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (threadIdx.x < 32) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = tc::make_idesc_f16(128, N);
        constexpr int ACC_STRIDE = NACC == 1 ? 0 : (NACC == 3 ? 160 : 512 / NACC);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        long long t0 = clock64();
        int cnt = 0;
        for (int it = 0; it < iters; ++it) {
            const int st = it % KB_STAGES;
            const uint32_t a = tc::smem_u32(sA + st * 16384), b = tc::smem_u32(sB + st * N * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t d = tmem + ((cnt % NACC) * ACC_STRIDE);
                tc::mma_f16(d, tc::make_desc_sw128(a + k * 32), tc::make_desc_sw128(b + k * 32), idesc, 1u);
                ++cnt;
            }
        }
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        long long t1 = clock64();
        out[0] = t1 - t0;
        out[1] = cnt;
    }
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}

template <int N, int NACC>
void run(const char* name) {
    constexpr int KS = 4;
    long long* d;
    cudaMalloc(&d, 16);
    const size_t smem = KS * (16384 + N * 128) + 1024;
    cudaFuncSetAttribute(k_bench<N, NACC, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) k_bench<N, NACC, KS><<<1, 128, smem>>>(2000, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    // all SMs busy variant: power / clock effects
    k_bench<N, NACC, KS><<<148, 128, smem>>>(2000, d);
    cudaDeviceSynchronize();
    long long h2[2] = {0, 0};
    cudaMemcpy(h2, d, 16, cudaMemcpyDeviceToHost);
    printf("%-28s N=%3d acc=%d : %7.2f cyc/MMA (ideal %5.1f) -> %5.1f %% of peak | 148 CTAs: %7.2f cyc/MMA  [%s]\n", name, N, NACC,
           double(h[0]) / h[1], N / 2.0, 100.0 * (N / 2.0) / (double(h[0]) / h[1]), double(h2[0]) / h2[1], cudaGetErrorString(e));
    cudaFree(d);
}


// Same MMA stream, but driven through the real kernel's producer/consumer barrier ring WITHOUT any data movement: a
// "producer" thread waits empty[s] and arrives full[s]; the issuer waits full[s], issues 4 MMAs, commits to empty[s].
// Cycles per K slab above 4 * N/2 are the cost of the synchronisation structure itself.
template <int N, int STAGES, int SLABS, bool RING>
__global__ void __launch_bounds__(128, 1) k_ring(int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[STAGES], empty[STAGES], done;
    __shared__ uint32_t slot;
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * SLABS * 16384;
    for (int i = threadIdx.x; i < (STAGES * SLABS * (16384 + N * 128)) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(&done, 1);
        tc::fence_barrier_init();
    }
    if (threadIdx.x < 32) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = slot;
    if (threadIdx.x == 64 && RING) {            // producer (no TMA)
        for (int it = 0; it < iters; ++it) {
            const int st = it % STAGES;
            tc::mbar_wait(&empty[st], ((it / STAGES) & 1) ^ 1);
            tc::mbar_arrive(&full[st]);
        }
    } else if (threadIdx.x == 32) {     // issuer
        constexpr uint32_t idesc = tc::make_idesc_f16(128, N);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int st = it % STAGES;
            if (RING) {
                tc::mbar_wait(&full[st], (it / STAGES) & 1);
                tc::fence_after_sync();
            }
#pragma unroll
            for (int sl = 0; sl < SLABS; ++sl) {
                const uint32_t a = tc::smem_u32(sA + (st * SLABS + sl) * 16384), b = tc::smem_u32(sB + (st * SLABS + sl) * N * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc::mma_f16(tmem, tc::make_desc_sw128(a + k * 32), tc::make_desc_sw128(b + k * 32), idesc, 1u);
            }
            tc::mma_commit(&empty[st]);
        }
        tc::mma_commit(&done);
        tc::mbar_wait(&done, 0);
        out[0] = clock64() - t0;
        out[1] = iters;
    }
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}

template <int N, int STAGES, int SLABS = 1, bool RING = true>
void run_ring() {
    long long* d;
    cudaMalloc(&d, 16);
    const size_t smem = STAGES * SLABS * (16384 + N * 128) + 1024;
    cudaFuncSetAttribute(k_ring<N, STAGES, SLABS, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) k_ring<N, STAGES, SLABS, RING><<<148, 128, smem>>>(4000, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%s, no data movement  N=%3d stages=%d, %d MMAs per stage : %7.1f cyc per stage (ideal %6.1f) -> %5.1f %% of peak [%s]\n",
           RING ? "barrier ring" : "commit only ", N, STAGES, 4 * SLABS, double(h[0]) / h[1], 2.0 * N * SLABS,
           100.0 * 2.0 * N * SLABS / (double(h[0]) / h[1]), cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run_ring<160, 6>();
    run_ring<160, 6, 1, false>();
    run_ring<160, 3, 2>();
    run_ring<160, 2, 3>();
    run_ring<128, 6>();
    run_ring<128, 3, 2>();
    run_ring<128, 6, 1, false>();
    run_ring<256, 4>();
    run_ring<160, 2>();
    run_ring<160, 3>();
    run_ring<160, 4>();
    run<128, 1>("single accumulator");
    run<128, 2>("two accumulators");
    run<160, 1>("single accumulator");
    run<160, 2>("two accumulators");
    run<160, 3>("three accumulators");
    run<192, 1>("single accumulator");
    run<192, 2>("two accumulators");
    run<256, 1>("single accumulator");
    run<256, 2>("two accumulators");
    run<64, 1>("single accumulator");
    run<64, 4>("four accumulators");
    return 0;
}

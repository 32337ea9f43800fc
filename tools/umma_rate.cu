// Hardware micro-benchmark (sm_100a): sustained issue/execute rate of tcgen05.mma cta_group::1
// kind::f16, M=128, SS mode (A and B from shared memory, SWIZZLE_128B K-major), as a function of
//   N (64/128/256), the number of K=16 MMAs issued between commits ("group"), whether the issuer
//   waits for each group's commit before the next (sync=1: a 1-deep pipeline; sync=0: commits are
//   only tracked), the number of distinct accumulators cycled through, and CTAs per SM.
// Used to separate "tensor pipe / smem operand bandwidth" limits from per-chunk synchronisation cost
// in conv_umma.cu / conv_halo_persist.cu.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate.bin umma_rate.cu
#include <cstdio>
#include <vector>
#include "../drawingspinup_b200/csrc/ptx.cuh"

using namespace dsu;

__global__ void rate_kernel(int n, int group, int sync, int naccum, int iters, int a_rows_shift, long long* out_cycles) {
    extern __shared__ uint8_t raw[];
    const uint32_t raw_u = smem_u32(raw);
    const uint32_t base = (raw_u + 1023u) & ~1023u;
    uint8_t* smem = raw + (base - raw_u);
    // A: 4 tiles of 128 rows (16 KB each), B: 4 tiles of n rows
    const uint32_t a_bytes = 128 * 128, b_bytes = n * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * a_bytes + 4 * b_bytes);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 10);
    const int tid = threadIdx.x;
    for (uint32_t i = tid; i < (4 * a_bytes + 4 * b_bytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
    fence_proxy_async_smem();
    const int cols = n * naccum <= 32 ? 32 : (n * naccum <= 64 ? 64 : (n * naccum <= 128 ? 128 : (n * naccum <= 256 ? 256 : 512)));
    if (tid < 32) {
        if (tid == 0) {
            mbar_init(smem_u32(bar), 1);                                   // bar[0]: waited on
            for (int i = 1; i < 9; ++i) mbar_init(smem_u32(bar + i), 1000000);  // dummies: commits only tracked, never complete
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(slot), cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (tid < 32) {
        const uint32_t idesc = umma_idesc_f16(128, n);
        const uint32_t bar_a = smem_u32(bar);
        long long t0 = clock64();
        int issued = 0;
        uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
            const uint32_t a_addr = base + (it & 3) * a_bytes + a_rows_shift * 128;
            const uint32_t b_addr = base + 4 * a_bytes + (it & 3) * b_bytes;
            if (elect_one()) {
                const uint64_t da = umma_desc_sw128(a_addr, 1024), db = umma_desc_sw128(b_addr, 1024);
                for (int g = 0; g < group; ++g) {
                    const uint32_t d = tmem + ((issued / 4) % naccum) * n;
                    umma_f16(d, da + 2 * (g & 3), db + 2 * (g & 3), idesc, issued >= 4 * naccum ? 1u : 0u);
                    ++issued;
                }
                umma_commit(sync ? bar_a : bar_a + 8 * (1 + (it & 7)));
                if (!sync && it == iters - 1) umma_commit(bar_a);
            }
            __syncwarp();
            issued = (it + 1) * group;
            if (sync) { mbar_wait(bar_a, phase); phase ^= 1; tc_fence_after(); }
        }
        if (!sync) mbar_wait(bar_a, 0);      // drain: the single final commit
        long long t1 = clock64();
        if (tid == 0 && blockIdx.x == 0) *out_cycles = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, cols);
}

int main() {
    long long* d_c;
    cudaMalloc(&d_c, 8);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, sms);
    printf("%4s %6s %5s %7s %8s %8s | %10s %12s %9s\n", "N", "group", "sync", "naccum", "cta/SM", "shift", "cyc/MMA", "ideal", "TFLOP/s*");
    const int iters = 2000;
    struct Cfg { int n, group, sync, naccum, ctas, shift; };
    std::vector<Cfg> cfgs;
    for (int n : {64, 128, 256})
        for (int group : {4, 8, 16, 32})
            for (int sync : {0, 1})
                cfgs.push_back({n, group, sync, 1, 1, 0});
    for (int n : {64, 128}) {
        cfgs.push_back({n, 4, 0, 2, 1, 0});
        cfgs.push_back({n, 8, 0, 2, 1, 0});
        cfgs.push_back({n, 4, 0, 1, 2, 0});
        cfgs.push_back({n, 4, 1, 1, 2, 0});
        cfgs.push_back({n, 16, 0, 1, 2, 0});
        cfgs.push_back({n, 16, 0, 1, 1, 3});
    }
    for (const Cfg& c : cfgs) {
        const size_t smem = 4 * 128 * 128 + 4 * c.n * 128 + 128 + 1024;
        if (c.ctas == 2 && smem > 110 * 1024) continue;
        rate_kernel<<<sms * c.ctas, 128, smem>>>(c.n, c.group, c.sync, c.naccum, iters, c.shift, d_c);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long cyc = 0;
        cudaMemcpy(&cyc, d_c, 8, cudaMemcpyDeviceToHost);
        const double per = static_cast<double>(cyc) / (static_cast<double>(iters) * c.group);
        const double ideal = 128.0 * c.n * 16 / 4096.0;
        // chip rate if every SM sustains this: 2*128*n*16 flop per MMA per CTA
        const double tf = 2.0 * 128 * c.n * 16 / per * 1.9e9 * sms * c.ctas / 1e12;
        printf("%4d %6d %5d %7d %8d %8d | %10.1f %12.1f %9.0f\n", c.n, c.group, c.sync, c.naccum, c.ctas, c.shift, per, ideal, tf);
    }
    printf("* assuming 1.9 GHz; cyc/MMA is per CTA\n");
    return 0;
}

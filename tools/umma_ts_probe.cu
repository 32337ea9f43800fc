// Hardware probe (sm_100a) for the TS form of tcgen05.mma (A operand read from TENSOR MEMORY, B from shared memory):
//   1. layout: which (lane, column, half) of TMEM holds A[m][k] for kind::f16 - written with tcgen05.st.32x32b,
//      read back through an identity B so that D[m][n] = A[m][n];
//   2. a NO-SWIZZLE K-major B tile of N rows x 16 K (32 B per row, 8-row x 16-byte core matrices, LBO = 128, SBO = 256)
//      as the B operand of such an MMA (the per-(tap, K-step) weight tile of the RIC kernel);
//   3. several warps issuing into the SAME accumulator (K-split without partial sums): exact integer data, the result
//      must equal the single-issuer result bit for bit, repeated to catch ordering races;
//   4. issue rate of TS MMAs: cycles per MMA per CTA for N = 64 / 128, 1-4 issuing warps, shared or private accumulators.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_ts_probe.bin umma_ts_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../drawingspinup_b200/csrc/ptx.cuh"

using namespace dsu;

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------ parts 1-3
// mode 0: A[m][k] = k, B = SW128 identity        -> expect D[m][n] = n  (n < 16)
// mode 1: A[m][k] = m % 64, B = SW128 identity   -> expect D[m][n] = m % 64
// mode 2: A[m][k] = k, B = no-swizzle identity (LBO 128, SBO 256)
// mode 3: integer GEMM, K = 16 * ksteps, `issuers` warps issue k-steps s % issuers == w into ONE zero-initialised accumulator
__global__ void probe_kernel(int mode, int issuers, int ksteps, float* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t raw_u = smem_u32(raw);
    const uint32_t base = (raw_u + 1023u) & ~1023u;
    uint8_t* smem = raw + (base - raw_u);
    uint8_t* Bsw = smem;                       // 64 rows x 128 B, SWIZZLE_128B
    uint8_t* Bns = smem + 64 * 128;            // up to 32 k-steps x (64 rows x 32 B), no swizzle
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 64 * 128 + 32 * 2048);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 64 * 64; i += blockDim.x) {
        const int row = i / 64, col = i % 64, chunk = col / 8, within = col % 8;
        reinterpret_cast<__half*>(Bsw + row * 128 + ((chunk ^ (row & 7)) << 4))[within] = __float2half((row == col && row < 16) ? 1.0f : 0.0f);
    }
    for (int s = 0; s < 32; ++s)
        for (int i = tid; i < 64 * 16; i += blockDim.x) {
            const int n = i / 16, k = i % 16;
            float v = (n == k) ? 1.0f : 0.0f;
            if (mode == 3) v = static_cast<float>(((n * 7 + k * 3 + s * 5) % 5) - 2);        // small integers
            reinterpret_cast<__half*>(Bns + s * 2048 + (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16)[k % 8] = __float2half(v);
        }
    fence_proxy_async_smem();
    if (warp == 0) {
        if (tid == 0) { mbar_init(smem_u32(bar), mode == 3 ? issuers : 1); fence_mbar_init(); }
        __syncwarp();
        tmem_alloc(smem_u32(slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t a_col = 64;                 // A lives at columns [64, 64 + 8 * ksteps)
    const int m = (warp & 3) * 32 + (tid & 31);
    if (warp < 4) {
        // every thread writes its own row of A: 8 columns (= one K16 step) per tcgen05.st
        for (int s = 0; s < ksteps; ++s) {
            uint32_t v[8];
            for (int c = 0; c < 8; ++c) {
                const int k = 2 * c;
                float lo, hi;
                if (mode == 0 || mode == 2) { lo = static_cast<float>(k); hi = static_cast<float>(k + 1); }
                else if (mode == 1) { lo = hi = static_cast<float>(m % 64); }
                else { lo = static_cast<float>(((m + k + s) % 7) - 3); hi = static_cast<float>(((m + k + 1 + s) % 7) - 3); }
                v[c] = pack2(lo, hi);
            }
            tmem_st8(tmem + lane_base + a_col + 8 * s, v);
        }
        if (mode == 3) {                        // zero the accumulator so that every issuer can accumulate
            uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int c = 0; c < 64; c += 8) tmem_st8(tmem + lane_base + c, z);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(128, 64);
    if (mode != 3) {
        if (tid == 0) {
            const uint64_t db = mode == 2 ? umma_desc_noswizzle(base + 64 * 128, 128, 256) : umma_desc_sw128(base, 1024);
            umma_f16_ts(tmem, tmem + a_col, db, idesc, 0u);
            umma_commit(smem_u32(bar));
        }
    } else if (warp < issuers) {
        if (elect_one()) {
            for (int s = warp; s < ksteps; s += issuers)
                umma_f16_ts(tmem, tmem + a_col + 8 * s, umma_desc_noswizzle(base + 64 * 128 + s * 2048, 128, 256), idesc, 1u);
            umma_commit(smem_u32(bar));
        }
        __syncwarp();
    }
    __syncthreads();
    mbar_wait(smem_u32(bar), 0);
    tc_fence_after();
    if (warp < 4) {
        for (int cb = 0; cb < 64; cb += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + lane_base + cb, v);
            tmem_ld_wait();
            for (int c = 0; c < 32; ++c) out[m * 64 + cb + c] = __uint_as_float(v[c]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ part 4: issue rate
// `issuers` warps, each issues `group` TS MMAs (N = n, K = 16, B no-swizzle tiles cycled) per commit, `iters` times.
// shared = 1: all into the same accumulator; 0: accumulator w * n.  ss = 1: the same with the A operand in shared memory
// (SWIZZLE_128B tile) for comparison.
__global__ void rate_kernel(int n, int issuers, int group, int shared_acc, int ss, int iters, long long* out_cycles) {
    extern __shared__ uint8_t raw[];
    const uint32_t raw_u = smem_u32(raw);
    const uint32_t base = (raw_u + 1023u) & ~1023u;
    uint8_t* smem = raw + (base - raw_u);
    const uint32_t b_tile = static_cast<uint32_t>(n) * 32;             // one (tap, K-step) weight tile
    const uint32_t a_off = 36 * b_tile;                                // SS comparison: 4 A tiles of 16 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + a_off + 4 * 16384);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 12);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < (a_off + 4 * 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    if (warp == 0) {
        if (tid == 0) {
            for (int i = 0; i < 4; ++i) mbar_init(smem_u32(bar + i), 1);                 // final commit of issuer i
            for (int i = 4; i < 12; ++i) mbar_init(smem_u32(bar + i), 1000000);           // tracked only
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t a_col = 512 - 72;                                    // 9 taps x 8 columns
    long long t0 = 0, t1 = 0;
    if (warp < issuers) {
        const uint32_t idesc = umma_idesc_f16(128, n);
        const uint32_t d = tmem + (shared_acc ? 0 : warp * n);
        __syncwarp();
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
                for (int g = 0; g < group; ++g) {
                    const int q = (it * group + g) % 36;
                    const uint64_t db = umma_desc_noswizzle(base + q * b_tile, 128, 256);
                    if (ss) umma_f16(d, umma_desc_sw128(base + a_off + (q & 3) * 16384, 1024) + 2 * (g & 3), db, idesc, 1u);
                    else umma_f16_ts(d, tmem + a_col + 8 * (q % 9), db, idesc, 1u);
                }
                umma_commit(smem_u32(bar + 4 + (it & 7)));
                if (it == iters - 1) umma_commit(smem_u32(bar + warp));
            }
            __syncwarp();
        }
        mbar_wait(smem_u32(bar + warp), 0);
        t1 = clock64();
        if ((tid & 31) == 0 && blockIdx.x == 0) out_cycles[warp] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ part 5: MMAs under producer traffic
// 6 issuing warps (warps 0-5) run TS MMAs exactly like part 4 (shared accumulator, `group` MMAs per commit) while 12 other
// warps (6-17) keep doing what the RIC producers do: 9 x ld.shared.v4 + 9 x tcgen05.st.x4 per item into OTHER columns
// (bg & 1) and / or packed-half math (bg & 2).  Does the producers' traffic slow the tensor pipe down?
__global__ void rate_bg_kernel(int n, int group, int bg, int iters, long long* out_cycles) {
    extern __shared__ uint8_t raw[];
    const uint32_t raw_u = smem_u32(raw);
    const uint32_t base = (raw_u + 1023u) & ~1023u;
    uint8_t* smem = raw + (base - raw_u);
    const uint32_t b_tile = static_cast<uint32_t>(n) * 32;
    const uint32_t halo_off = 36 * b_tile;                               // 24 KB "halo tile" for the background loads
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + halo_off + 24576);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 16);
    volatile int* stop = reinterpret_cast<volatile int*>(slot + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (uint32_t i = tid; i < (halo_off + 24576) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0) *stop = 0;
    fence_proxy_async_smem();
    if (warp == 0) {
        if (tid == 0) {
            for (int i = 0; i < 6; ++i) mbar_init(smem_u32(bar + i), 1);
            for (int i = 6; i < 14; ++i) mbar_init(smem_u32(bar + i), 1000000);
            mbar_init(smem_u32(bar + 14), 6 * 32);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t a_col = 512 - 72;
    if (warp < 6) {
        const uint32_t idesc = umma_idesc_f16(128, n);
        __syncwarp();
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
                for (int g = 0; g < group; ++g) {
                    const int q = (it * group + g) % 36;
                    umma_f16_ts(tmem, tmem + a_col + 8 * (q % 9), umma_desc_noswizzle(base + q * b_tile, 128, 256), idesc, 1u);
                }
                umma_commit(smem_u32(bar + 6 + (it & 7)));
                if (it == iters - 1) umma_commit(smem_u32(bar + warp));
            }
            __syncwarp();
        }
        mbar_wait(smem_u32(bar + warp), 0);
        const long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) out_cycles[warp] = t1 - t0;
        mbar_arrive(smem_u32(bar + 14));
        if (warp == 0) { mbar_wait(smem_u32(bar + 14), 0); if (lane == 0) *stop = 1; }
    } else if (bg) {
        const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t hb = base + halo_off + ((warp - 6) * 32 + lane) % 160 * 128;
        uint32_t acc = lane;
        long long n_items = 0;
        while (!*stop) {
            uint32_t v[9][4];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                if (bg & 1) {
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[k][0]), "=r"(v[k][1]), "=r"(v[k][2]), "=r"(v[k][3]) : "r"(hb + ((k * 144) % 2048)));
                } else { v[k][0] = acc + k; v[k][1] = acc; v[k][2] = k; v[k][3] = acc ^ k; }
            }
            if (bg & 2) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int k = 0; k < 9; ++k)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            __half2 x = *reinterpret_cast<__half2*>(&v[k][c]);
                            x = __hfma2(x, x, x);
                            v[k][c] = *reinterpret_cast<uint32_t*>(&x);
                        }
            }
            if (bg & 1) {
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tmem + lane_base + 256 + k * 8), "r"(v[k][0]), "r"(v[k][1]), "r"(v[k][2]), "r"(v[k][3]) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            acc += v[0][0];
            ++n_items;
        }
        if (lane == 0 && blockIdx.x == 0 && warp == 6) out_cycles[6] = n_items + (acc & 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ part 6: tcgen05.st cost
// `nw` warps store `per_wait` x (32 lanes x W columns) and then tcgen05.wait::st, in a loop for a fixed number of iterations;
// optionally 6 other warps keep the tensor pipe saturated with TS MMAs (N = 128).  cycles per store instruction per warp.
template <int W>
__device__ __forceinline__ void st_w(uint32_t taddr, uint32_t v) {
    if constexpr (W == 2) asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %1};" ::"r"(taddr), "r"(v) : "memory");
    if constexpr (W == 4) asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(taddr), "r"(v) : "memory");
    if constexpr (W == 8) asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(v) : "memory");
    if constexpr (W == 16) asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(v) : "memory");
    if constexpr (W == 32) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(v) : "memory");
}
template <int W>
__global__ void st_cost_kernel(int nw, int per_wait, int with_mma, int iters, long long* out_cycles) {
    extern __shared__ uint8_t raw[];
    const uint32_t raw_u = smem_u32(raw);
    const uint32_t base = (raw_u + 1023u) & ~1023u;
    uint8_t* smem = raw + (base - raw_u);
    const uint32_t b_tile = 128 * 32;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 36 * b_tile);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 16);
    volatile int* stop = reinterpret_cast<volatile int*>(slot + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (uint32_t i = tid; i < 36 * b_tile / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0) *stop = 0;
    fence_proxy_async_smem();
    if (warp == 0) {
        if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(bar + i), 1000000); mbar_init(smem_u32(bar + 8), nw * 32); fence_mbar_init(); }
        __syncwarp();
        tmem_alloc(smem_u32(slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (warp < 6) {
        if (with_mma) {
            const uint32_t idesc = umma_idesc_f16(128, 128);
            int it = 0;
            while (!*stop) {
                if (elect_one()) {
                    for (int g = 0; g < 9; ++g) {
                        const int q = (it * 9 + g) % 36;
                        umma_f16_ts(tmem, tmem + 440 + 8 * (q % 9), umma_desc_noswizzle(base + q * b_tile, 128, 256), idesc, 1u);
                    }
                    umma_commit(smem_u32(bar + (it & 7)));
                }
                __syncwarp();
                ++it;
            }
        }
    } else if (warp < 6 + nw) {
        const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
        __syncwarp();
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int k = 0; k < per_wait; ++k) st_w<W>(tmem + lane_base + 128 + ((k * W) % 256), static_cast<uint32_t>(it + k));
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        const long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0 && warp == 6) out_cycles[0] = t1 - t0;
        mbar_arrive(smem_u32(bar + 8));
        if (warp == 6) { mbar_wait(smem_u32(bar + 8), 0); if (lane == 0) *stop = 1; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    float* d_out;
    cudaMalloc(&d_out, 128 * 64 * 4);
    const size_t smem = 64 * 128 + 32 * 2048 + 64 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<float> h(128 * 64);
    auto run = [&](int mode, int issuers, int ksteps) {
        cudaMemset(d_out, 0xff, 128 * 64 * 4);
        probe_kernel<<<1, 128, smem>>>(mode, issuers, ksteps, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost);
    };
    // ---- 1/2: layout
    for (int mode = 0; mode < 3; ++mode) {
        run(mode, 1, 1);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 16; ++n) {
                const float want = mode == 1 ? static_cast<float>(m % 64) : static_cast<float>(n);
                if (h[m * 64 + n] != want) ++bad;
            }
        printf("TS_PROBE mode %d (%s): mismatches vs 'A[m][k] = lane m, column k/2, half k%%2' = %d  %s\n", mode,
               mode == 0 ? "A=k, B SW128 identity" : (mode == 1 ? "A=m, B SW128 identity" : "A=k, B no-swizzle identity LBO128 SBO256"),
               bad, bad == 0 ? "OK" : "MISMATCH");
        if (bad) {
            for (int m : {0, 1, 2, 31, 32, 33, 64, 127}) {
                printf("   m=%3d :", m);
                for (int n = 0; n < 16; ++n) printf(" %4.0f", h[m * 64 + n]);
                printf("\n");
            }
        }
    }
    // ---- 3: several issuers into one accumulator (exact integers)
    const int ksteps = 32;
    std::vector<float> ref(128 * 64, 0.0f);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            float acc = 0;
            for (int s = 0; s < ksteps; ++s)
                for (int k = 0; k < 16; ++k)
                    acc += static_cast<float>(((m + k + s) % 7) - 3) * static_cast<float>(((n * 7 + k * 3 + s * 5) % 5) - 2);
            ref[m * 64 + n] = acc;
        }
    for (int issuers = 1; issuers <= 4; ++issuers) {
        int bad_runs = 0, worst = 0;
        for (int rep = 0; rep < 50; ++rep) {
            run(3, issuers, ksteps);
            int bad = 0;
            for (size_t i = 0; i < ref.size(); ++i) if (h[i] != ref[i]) ++bad;
            if (bad) { ++bad_runs; if (bad > worst) worst = bad; }
        }
        printf("TS_PROBE shared accumulator, %d issuing warp(s), K = %d: %d of 50 runs differ from the exact integer result (worst %d elements)  %s\n",
               issuers, 16 * ksteps, bad_runs, worst, bad_runs == 0 ? "OK" : "MISMATCH");
    }
    // ---- 4: issue rate
    long long* d_c;
    cudaMalloc(&d_c, 4 * 8);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("%4s %8s %6s %7s %4s | %12s %8s %9s\n", "N", "issuers", "group", "shared", "A", "cyc/MMA/CTA", "ideal", "TFLOP/s*");
    const int iters = 400;
    for (int n : {64, 128})
        for (int ss : {0, 1})
            for (int shared_acc : {1, 0})
                for (int issuers : {1, 2, 3, 4})
                    for (int group : {9, 36}) {
                        if (!shared_acc && issuers * n > 512 - 72) continue;
                        if (issuers == 1 && !shared_acc) continue;
                        const size_t sm = 36 * n * 32 + 4 * 16384 + 128 + 1024;
                        rate_kernel<<<sms, 128, sm>>>(n, issuers, group, shared_acc, ss, iters, d_c);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("rate: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
                        long long cyc[4] = {0, 0, 0, 0};
                        cudaMemcpy(cyc, d_c, 32, cudaMemcpyDeviceToHost);
                        long long mx = 0;
                        for (int i = 0; i < issuers; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
                        const double per = static_cast<double>(mx) / (static_cast<double>(iters) * group * issuers);
                        const double tf = 2.0 * 128 * n * 16 / per * 1.9e9 * sms / 1e12;
                        printf("%4d %8d %6d %7d %4s | %12.1f %8.1f %9.0f\n", n, issuers, group, shared_acc, ss ? "smem" : "tmem", per,
                               128.0 * n * 16 / 4096.0, tf);
                    }
    printf("* assuming 1.9 GHz\n");
    // ---- 5: six issuers under producer-like background traffic of 12 warps
    cudaFuncSetAttribute(rate_bg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    long long* d_c2;
    cudaMalloc(&d_c2, 8 * 8);
    printf("%4s %6s %28s | %12s %8s | %s\n", "N", "group", "background (12 warps)", "cyc/MMA/CTA", "ideal", "bg items per warp");
    for (int n : {64, 128})
        for (int group : {3, 9})
            for (int bg : {0, 1, 2, 3}) {
                const size_t sm = 36 * n * 32 + 24576 + 256 + 1024;
                cudaMemset(d_c2, 0, 64);
                rate_bg_kernel<<<sms, 18 * 32, sm>>>(n, group, bg, 600, d_c2);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("rate_bg: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
                long long cyc[8] = {0};
                cudaMemcpy(cyc, d_c2, 64, cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (int i = 0; i < 6; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
                const double per = static_cast<double>(mx) / (600.0 * group * 6);
                printf("%4d %6d %28s | %12.1f %8.1f | %lld\n", n, group,
                       bg == 0 ? "none" : (bg == 1 ? "ld.shared + tcgen05.st" : (bg == 2 ? "HFMA2 only" : "ld.shared + HFMA2 + tcgen05.st")), per,
                       128.0 * n * 16 / 4096.0, cyc[6]);
            }
    // ---- 6: cost of tcgen05.st by width, with / without a saturated tensor pipe
    printf("%6s %6s %9s %9s | %14s %14s\n", "width", "warps", "per_wait", "MMAs", "cyc/st/warp", "B/clk/SM");
    auto st_run = [&](auto kern, int W, int nw, int per_wait, int with_mma) {
        const size_t sm = 36 * 128 * 32 + 256 + 1024;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaMemset(d_c2, 0, 64);
        const int iters = 2000;
        kern<<<sms, 18 * 32, sm>>>(nw, per_wait, with_mma, iters, d_c2);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("st_cost: CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
        long long cyc = 0;
        cudaMemcpy(&cyc, d_c2, 8, cudaMemcpyDeviceToHost);
        const double per = static_cast<double>(cyc) / (static_cast<double>(iters) * per_wait);
        printf("%6d %6d %9d %9s | %14.1f %14.1f\n", W, nw, per_wait, with_mma ? "N=128" : "none", per, nw * 32.0 * W * 4 / per);
    };
    for (int with_mma : {0, 1})
        for (int nw : {4, 12})
            for (int per_wait : {1, 9}) {
                st_run(st_cost_kernel<2>, 2, nw, per_wait, with_mma);
                st_run(st_cost_kernel<4>, 4, nw, per_wait, with_mma);
                st_run(st_cost_kernel<8>, 8, nw, per_wait, with_mma);
                st_run(st_cost_kernel<16>, 16, nw, per_wait, with_mma);
                st_run(st_cost_kernel<32>, 32, nw, per_wait, with_mma);
            }
    return 0;
}

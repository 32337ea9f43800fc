// Hardware probe (sm_100a): which smem rows does tcgen05.mma read for a K-major SWIZZLE_128B A
// descriptor whose start address is NOT 1024-byte aligned (shifted by whole 128-byte rows), and
// whose 8-row-group stride (SBO) is not 1024?  The answer decides whether a halo tile loaded once
// can feed all taps of a convolution through shifted descriptors (DESIGN.md "halo reuse").
//
// B = identity (64x64), so D[m][n] = A[m][n]; A is filled either with its absolute smem row index
// or with its logical column, written with the absolute-address XOR swizzle (chunk ^ (row & 7))
// that TMA / cp.async producers use.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../drawingspinup_b200/csrc/ptx.cuh"

using namespace dsu;

constexpr int kRows = 256;   // smem rows available to A

__global__ void probe_kernel(int start_row, int sbo_bytes, int base_off_mode, int fill_mode, float* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t raw_u = smem_u32(raw);
    const uint32_t base = (raw_u + 1023u) & ~1023u;
    uint8_t* smem = raw + (base - raw_u);
    uint8_t* A = smem;                         // kRows x 128 B
    uint8_t* Bm = smem + kRows * 128;          // 64 rows x 128 B (1024-aligned)
    uint64_t* bar = reinterpret_cast<uint64_t*>(Bm + 64 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x;
    for (int i = tid; i < kRows * 64; i += blockDim.x) {
        const int row = i / 64, col = i % 64;
        const float v = fill_mode == 0 ? static_cast<float>(row) : static_cast<float>(col);
        const int chunk = col / 8, within = col % 8;
        const int phys = chunk ^ (row & 7);
        reinterpret_cast<__half*>(A + row * 128 + phys * 16)[within] = __float2half(v);
    }
    for (int i = tid; i < 64 * 64; i += blockDim.x) {
        const int row = i / 64, col = i % 64;
        const int chunk = col / 8, within = col % 8;
        const int phys = chunk ^ (row & 7);
        reinterpret_cast<__half*>(Bm + row * 128 + phys * 16)[within] = __float2half(row == col ? 1.0f : 0.0f);
    }
    fence_proxy_async_smem();
    if (tid < 32) {
        if (tid == 0) { mbar_init(smem_u32(bar), 1); fence_mbar_init(); }
        __syncwarp();
        tmem_alloc(smem_u32(slot), 64);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const uint32_t a_addr = base + start_row * 128;
        const uint32_t b_addr = base + kRows * 128;
        const uint32_t bo = base_off_mode ? ((a_addr >> 7) & 7u) : 0u;
        for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_desc_sw128(a_addr + k * 32, sbo_bytes, bo);
            const uint64_t db = umma_desc_sw128(b_addr + k * 32, 1024, 0);
            umma_f16(tmem, da, db, umma_idesc_f16(128, 64), k > 0);
        }
        umma_commit(smem_u32(bar));
    }
    __syncthreads();
    mbar_wait(smem_u32(bar), 0);
    tc_fence_after();
    const int warp = tid >> 5;
    for (int cb = 0; cb < 64; cb += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + cb, v);
        tmem_ld_wait();
        for (int c = 0; c < 32; ++c) out[tid * 64 + cb + c] = __uint_as_float(v[c]);
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 64);
}

int main() {
    float* d_out;
    cudaMalloc(&d_out, 128 * 64 * 4);
    const size_t smem = kRows * 128 + 64 * 128 + 64 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    struct Case { int start, sbo, bo; };
    const Case cases[] = {{0, 1024, 0}, {1, 1024, 0}, {1, 1024, 1}, {3, 1024, 0}, {3, 1024, 1}, {8, 1024, 0},
                          {0, 1280, 0}, {0, 1280, 1}, {11, 1280, 0}, {11, 1280, 1}, {2, 2304, 0}, {2, 2304, 1}};
    std::vector<float> rows(128 * 64), cols(128 * 64);
    for (const Case& c : cases) {
        probe_kernel<<<1, 128, smem>>>(c.start, c.sbo, c.bo, 0, d_out);
        cudaMemcpy(rows.data(), d_out, rows.size() * 4, cudaMemcpyDeviceToHost);
        probe_kernel<<<1, 128, smem>>>(c.start, c.sbo, c.bo, 1, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(cols.data(), d_out, cols.size() * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("case start=%d sbo=%d bo=%d: CUDA error %s\n", c.start, c.sbo, c.bo, cudaGetErrorString(e)); return 1; }
        int bad_row = 0, bad_col = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
                const float exp_row = static_cast<float>(c.start + (m / 8) * (c.sbo / 128) + (m % 8));
                if (rows[m * 64 + n] != exp_row) ++bad_row;
                if (cols[m * 64 + n] != static_cast<float>(n)) ++bad_col;
            }
        printf("PROBE start_row=%2d sbo=%4d base_offset=%s : row mismatches %5d, col mismatches %5d  %s\n", c.start, c.sbo,
               c.bo ? "(addr>>7)&7" : "0", bad_row, bad_col, (bad_row == 0 && bad_col == 0) ? "OK (linear rows + absolute swizzle)" : "MISMATCH");
        if (bad_row || bad_col) {
            printf("   m : row(n=0) row(n=8) row(n=63) | col(n=0) col(n=8) col(n=63)\n");
            for (int m = 0; m < 20; ++m)
                printf("  %2d : %6.0f %6.0f %6.0f | %6.0f %6.0f %6.0f\n", m, rows[m * 64], rows[m * 64 + 8], rows[m * 64 + 63],
                       cols[m * 64], cols[m * 64 + 8], cols[m * 64 + 63]);
        }
    }
    return 0;
}

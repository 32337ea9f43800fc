// Hardware probe (sm_100a), companion of umma_probe.cu: K-major SWIZZLE_64B operands (64-byte rows = 32 fp16 of K,
// 8-row groups 512 B apart, 16-byte chunk XOR ((row >> 1) & 3)).  B = identity (32x32 inside N=64), so D[m][n] = A[m][n]
// for n < 32.  Confirms the layout the double-buffered RIC producer writes (ric_producer.cuh, fast path).
#include <cstdio>
#include <vector>
#include "../drawingspinup_b200/csrc/ptx.cuh"
using namespace dsu;

__global__ void probe64(int start_row, int fill_mode, float* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t raw_u = smem_u32(raw);
    const uint32_t base = (raw_u + 1023u) & ~1023u;
    uint8_t* smem = raw + (base - raw_u);
    uint8_t* A = smem;                    // 256 rows x 64 B
    uint8_t* Bm = smem + 256 * 64;        // 64 rows x 64 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(Bm + 64 * 64);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x;
    for (int i = tid; i < 256 * 32; i += blockDim.x) {
        const int row = i / 32, col = i % 32, chunk = col / 8, within = col % 8;
        const float v = fill_mode == 0 ? static_cast<float>(row) : static_cast<float>(col);
        reinterpret_cast<__half*>(A + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4))[within] = __float2half(v);
    }
    for (int i = tid; i < 64 * 32; i += blockDim.x) {
        const int row = i / 32, col = i % 32, chunk = col / 8, within = col % 8;
        reinterpret_cast<__half*>(Bm + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4))[within] = __float2half(row == col ? 1.0f : 0.0f);
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
    if (tid < 32) {
        if (elect_one()) {
            const uint64_t da = umma_desc_sw64(base + start_row * 64, 512), db = umma_desc_sw64(base + 256 * 64, 512);
            umma_f16(tmem, da, db, umma_idesc_f16(128, 64), 0u);
            umma_f16(tmem, da + 2, db + 2, umma_idesc_f16(128, 64), 1u);
            umma_commit(smem_u32(bar));
        }
        __syncwarp();
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
    const size_t smem = 256 * 64 + 64 * 64 + 64 + 1024;
    std::vector<float> rows(128 * 64), cols(128 * 64);
    for (int start : {0, 8, 16}) {
        probe64<<<1, 128, smem>>>(start, 0, d_out);
        cudaMemcpy(rows.data(), d_out, rows.size() * 4, cudaMemcpyDeviceToHost);
        probe64<<<1, 128, smem>>>(start, 1, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(cols.data(), d_out, cols.size() * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        int bad_row = 0, bad_col = 0, bad_zero = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
                if (n < 32) {
                    if (rows[m * 64 + n] != static_cast<float>(start + m)) ++bad_row;
                    if (cols[m * 64 + n] != static_cast<float>(n)) ++bad_col;
                } else if (rows[m * 64 + n] != 0.0f) ++bad_zero;
            }
        printf("PROBE64 start_row=%2d: row mismatches %d, col mismatches %d, nonzero beyond K %d  %s\n", start, bad_row, bad_col, bad_zero,
               (bad_row | bad_col | bad_zero) ? "MISMATCH" : "OK");
        if (bad_row | bad_col) for (int m = 0; m < 10; ++m)
            printf("  %2d : rows %4.0f %4.0f %4.0f | cols %4.0f %4.0f %4.0f\n", m, rows[m * 64], rows[m * 64 + 8], rows[m * 64 + 31], cols[m * 64], cols[m * 64 + 8], cols[m * 64 + 31]);
    }
    return 0;
}

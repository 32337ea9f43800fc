// Hardware probe (sm_100a): a K-major NO-SWIZZLE A descriptor over a pixel-linear buffer (16 bytes = 8 fp16
// channels per pixel) with OVERLAPPING core matrices: leading-dimension byte offset 16 (the next 8 K elements
// are the next pixel) and stride byte offset = the row pitch of a halo tile.  If tcgen05.mma simply reads
//     addr(m, k) = start + (m / 8) * SBO + (k / 8) * LBO + (m % 8) * 16 + (k % 8) * 2
// then the 7 taps of one kernel row of an 8-channel 7x7 convolution are ONE contiguous K = 56 (+8 zero-weight)
// slice of the halo tile and conv0 needs no im2col at all (conv_first.cu).
//
// B (SWIZZLE_128B, N = 64) selects K element n into column n for n < 16, so one K=16 MMA returns the 16 A
// elements each row read.  Fill 0 stores the pixel index, fill 1 the channel index.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe_ns umma_probe_ns.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../drawingspinup_b200/csrc/ptx.cuh"

using namespace dsu;

constexpr int kPix = 1024;   // 16 KB of pixels

__device__ __forceinline__ uint64_t desc_noswizzle(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;   // layout type 0 = no swizzle
}

__global__ void probe_kernel(int start_pix, int lbo, int sbo, int fill_mode, float* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t raw_u = smem_u32(raw);
    const uint32_t base = (raw_u + 1023u) & ~1023u;
    uint8_t* smem = raw + (base - raw_u);
    uint8_t* P = smem;                          // kPix x 16 B
    uint8_t* Bm = smem + kPix * 16;             // 64 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(Bm + 64 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x;
    for (int i = tid; i < kPix * 8; i += blockDim.x)
        reinterpret_cast<__half*>(P)[i] = __float2half(fill_mode == 0 ? static_cast<float>(i / 8) : static_cast<float>(i % 8));
    for (int i = tid; i < 64 * 64; i += blockDim.x) {
        const int row = i / 64, col = i % 64;
        const int chunk = col / 8, within = col % 8;
        reinterpret_cast<__half*>(Bm + row * 128 + ((chunk ^ (row & 7)) << 4))[within] = __float2half((row == col && row < 16) ? 1.0f : 0.0f);
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
        const uint64_t da = desc_noswizzle(base + start_pix * 16, lbo, sbo);
        const uint64_t db = umma_desc_sw128(base + kPix * 16, 1024, 0);
        umma_f16(tmem, da, db, umma_idesc_f16(128, 64), 0u);
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
    const size_t smem = kPix * 16 + 64 * 128 + 64 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    struct Case { int start, lbo, sbo; };
    // canonical (core matrices 128 B apart in K, 256 B apart in M), then the overlapping halo-tile forms
    const Case cases[] = {{0, 128, 256}, {0, 256, 128}, {0, 16, 224}, {5, 16, 224}, {0, 224, 16}, {3, 16, 352}, {1, 16, 128}, {0, 32, 224}};
    std::vector<float> pix(128 * 64), ch(128 * 64);
    for (const Case& c : cases) {
        probe_kernel<<<1, 128, smem>>>(c.start, c.lbo, c.sbo, 0, d_out);
        cudaMemcpy(pix.data(), d_out, pix.size() * 4, cudaMemcpyDeviceToHost);
        probe_kernel<<<1, 128, smem>>>(c.start, c.lbo, c.sbo, 1, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(ch.data(), d_out, ch.size() * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("case start=%d lbo=%d sbo=%d: CUDA error %s\n", c.start, c.lbo, c.sbo, cudaGetErrorString(e)); return 1; }
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int k = 0; k < 16; ++k) {
                const int byte = c.start * 16 + (m / 8) * c.sbo + (k / 8) * c.lbo + (m % 8) * 16 + (k % 8) * 2;
                if (pix[m * 64 + k] != static_cast<float>(byte / 16) || ch[m * 64 + k] != static_cast<float>((byte % 16) / 2)) ++bad;
            }
        printf("PROBE_NS start_pix=%d lbo=%3d sbo=%3d : mismatches vs start+(m/8)*SBO+(k/8)*LBO+(m%%8)*16+(k%%8)*2 = %4d  %s\n",
               c.start, c.lbo, c.sbo, bad, bad == 0 ? "OK" : "MISMATCH");
        if (bad) {
            printf("   m : pixel read for k=0 / k=8 | channel for k=0,1,7,8,15\n");
            for (int m : {0, 1, 2, 7, 8, 9, 16, 127})
                printf("  %3d : %5.0f %5.0f | %2.0f %2.0f %2.0f %2.0f %2.0f\n", m, pix[m * 64], pix[m * 64 + 8], ch[m * 64], ch[m * 64 + 1],
                       ch[m * 64 + 7], ch[m * 64 + 8], ch[m * 64 + 15]);
        }
    }
    return 0;
}

// Persistent variant of the RIC convolution (conv_umma.cu, modes 1-3) for the Cout <= 64 layers of stage 1
// (conv1, conv_11, conv_11_a.3): their tiles have only 1-3 channel blocks, so a one-tile CTA spends as long in
// prologue + epilogue as in production.  One CTA per SM walks a static tile list; accumulators are
// double-buffered in TMEM (2 sets x ks K-split partials x Cout columns <= 512) and 4 dedicated epilogue warps
// drain tile i while the 8 producer warps gather/blend tile i+1.
// Registers: 512 threads cap the launch at 128 regs/thread; the producers need ~168, so the warpgroups
// rebalance with setmaxnreg (producers up, epilogue / issuers / loader down).
//
// Roles: warps 0-7 producers (ric_producer.cuh), 8-11 epilogue, 12-14 MMA issuers (taps t % ks), 15 weight loader.
#include "conv_device.cuh"
#include "ric_producer.cuh"

namespace dsu {

namespace {

constexpr int kRpThreads = 512;
constexpr int kRpBars = 2 * 9 + 2 * kMaxStagesB + 4;

struct RpSmem {
    uint32_t a0, b0, par, bars, total;
};

__host__ __device__ inline RpSmem rp_smem(int sb, int b_bytes, int cout) {
    RpSmem L;
    L.a0 = 0;
    L.b0 = 9 * kABytes;
    L.par = L.b0 + sb * b_bytes;
    L.bars = (L.par + (7 * cout + 4) * 4 + 15u) & ~15u;
    L.total = L.bars + (kRpBars + 1) * 8;
    return L;
}

template <int kRegs>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }

}  // namespace

// kMode: 1 = fp16 / fp32 blend, 2 = split fp16 hi|lo, 3 = fp16 / packed half2 blend
template <int kMode>
__global__ void __launch_bounds__(kRpThreads, 1)
conv_ric_persist_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_u32);
    const int SB = p.sb, C = p.Cout, NI = p.ks;
    const int SBK = SB / NI;
    const RpSmem L = rp_smem(SB, p.b_bytes, C);
    float* s_par = reinterpret_cast<float*>(smem + L.par);
    const uint32_t bar_full_a = base + L.bars;
    const uint32_t bar_empty_a = bar_full_a + 9 * 8;
    const uint32_t bar_full_b = bar_empty_a + 9 * 8;
    const uint32_t bar_empty_b = bar_full_b + kMaxStagesB * 8;
    const uint32_t bar_acc_full = bar_empty_b + kMaxStagesB * 8;
    const uint32_t bar_acc_empty = bar_acc_full + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bars + kRpBars * 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_x = (p.Wout + kTileW - 1) / kTileW, tiles_y = (p.Hout + kTileH - 1) / kTileH;
    const int tiles_per_frame = tiles_x * tiles_y;
    const int total_tiles = tiles_per_frame * p.B;
    const int my_tiles = (total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    constexpr int kEpi = 128;

    if (warp == 12) {
        if (lane == 0) {
            for (int s = 0; s < 9; ++s) {
                mbar_init(bar_full_a + 8 * s, kWorkers / 32);     // one arrival per producer warp
                mbar_init(bar_empty_a + 8 * s, 1);
            }
            for (int s = 0; s < SB; ++s) {
                mbar_init(bar_full_b + 8 * s, 1);
                mbar_init(bar_empty_b + 8 * s, 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(bar_acc_full + 8 * s, NI);
                mbar_init(bar_acc_empty + 8 * s, kEpi);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
        tmem_relinquish();
    } else if (warp >= 8 && warp < 12) {
        load_epilogue_params(p, s_par, tid - 256, kEpi);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto tile_coords = [&](int it, int& n, int& ty0, int& tx0) {
        const int t = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
        n = t / tiles_per_frame;
        const int r = t - n * tiles_per_frame;
        ty0 = (r / tiles_x) * kTileH;
        tx0 = (r % tiles_x) * kTileW;
    };

    if (warp < 8) {
        // ======================================================== producers (2 warpgroups)
        reg_inc<176>();
        for (int it = 0; it < my_tiles; ++it) {
            int n, ty0, tx0;
            tile_coords(it, n, ty0, tx0);
            ric_produce<kMode == 2, kMode == 3>(p, smem + L.a0, bar_full_a, bar_empty_a, tid, n, ty0, tx0, it * p.nblocks);
        }
    } else if (warp < 12) {
        // ======================================================== epilogue warpgroup (one warp per TMEM lane quadrant)
        reg_dec<112>();
        const int quad = warp - 8;
        const int r = quad * 32 + lane;
        for (int it = 0; it < my_tiles; ++it) {
            int n, ty0, tx0;
            tile_coords(it, n, ty0, tx0);
            const int set = it & 1;
            mbar_wait(bar_acc_full + 8 * set, (it >> 1) & 1);
            tc_fence_after();
            epilogue_row<(kMode == 2 ? 0x0702u : 0x0007u)>(p, s_par, tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(set * NI * C), n,
                         ty0 + (r >> 4), tx0 + (r & 15), 0, NI, C, 1);
            tc_fence_before();
            mbar_arrive(bar_acc_empty + 8 * set);
        }
    } else {
        // ======================================================== issuers (warps 12-14) and weight loader (warp 15)
        reg_dec<40>();
        const int wi = warp - 12;
        if (wi < 3) {
            if (wi < NI) {
                const uint32_t idesc = umma_idesc_f16(kTileM, C);
                int cnt = 0;
                for (int it = 0; it < my_tiles; ++it) {
                    const int set = it & 1;
                    if (it >= 2) {
                        mbar_wait(bar_acc_empty + 8 * set, ((it >> 1) - 1) & 1);
                        tc_fence_after();
                    }
                    const uint32_t d_addr = tmem_base + static_cast<uint32_t>((set * NI + wi) * C);
                    uint32_t acc = 0;
                    for (int b = 0; b < p.nblocks; ++b) {
                        const int g = it * p.nblocks + b;
                        const uint32_t km = b == p.nblocks - 1 ? p.kmask_last : p.kmask_full;
                        const uint32_t km2 = b == p.nblocks - 1 ? p.kmask2_last : p.kmask2_full;
                        for (int t = wi; t < 9; t += NI, ++cnt) {
                            const int s_b = wi * SBK + cnt % SBK;
                            mbar_wait(bar_full_b + 8 * s_b, (cnt / SBK) & 1);
                            mbar_wait(bar_full_a + 8 * t, g & 1);
                            tc_fence_after();
                            const uint32_t a_addr = base + L.a0 + t * kABytes;
                            const uint32_t b_addr = base + L.b0 + s_b * p.b_bytes;
                            if (elect_one()) {
                                const uint64_t da0 = umma_desc_sw128(a_addr, 1024), db0 = umma_desc_sw128(b_addr, 1024);
                                if (km2) {
                                    // split-fp16: A = [a_hi | a_lo] (steps 0-1 | 2-3), B = [W_hi | W_lo]; a_hi*W_hi + a_lo*W_hi + a_hi*W_lo
                                    uint32_t a2 = acc;
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        if ((km >> k) & 1) { umma_f16(d_addr, da0 + 2 * k, db0 + 2 * (k & 1), idesc, a2); a2 = 1u; }
#pragma unroll
                                    for (int k = 0; k < 2; ++k)
                                        if ((km2 >> k) & 1) umma_f16(d_addr, da0 + 2 * k, db0 + 4 + 2 * k, idesc, 1u);
                                } else if (km == 0xFu) {
                                    umma_f16(d_addr, da0, db0, idesc, acc);
                                    umma_f16(d_addr, da0 + 2, db0 + 2, idesc, 1u);
                                    umma_f16(d_addr, da0 + 4, db0 + 4, idesc, 1u);
                                    umma_f16(d_addr, da0 + 6, db0 + 6, idesc, 1u);
                                } else {
                                    uint32_t a2 = acc;
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        if ((km >> k) & 1) { umma_f16(d_addr, da0 + 2 * k, db0 + 2 * k, idesc, a2); a2 = 1u; }
                                }
                                umma_commit(bar_empty_a + 8 * t);
                                umma_commit(bar_empty_b + 8 * s_b);
                            }
                            acc = 1u;
                            __syncwarp();
                        }
                    }
                    if (elect_one()) umma_commit(bar_acc_full + 8 * set);
                    __syncwarp();
                }
            }
            tc_fence_before();
        } else {
            int cnt[3] = {0, 0, 0};
            for (int it = 0; it < my_tiles; ++it)
                for (int q = 0; q < p.nchunks; ++q) {
                    const int k = (q % 9) % NI;
                    int c = 0;
#pragma unroll
                    for (int i = 0; i < 3; ++i) if (i == k) { c = cnt[i]; cnt[i] = c + 1; }
                    const int s = k * SBK + c % SBK;
                    if (c >= SBK) mbar_wait(bar_empty_b + 8 * s, ((c / SBK) - 1) & 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bar_full_b + 8 * s, static_cast<uint32_t>(p.b_bytes));
                        bulk_g2s(base + L.b0 + s * p.b_bytes, p.wpack + static_cast<size_t>(q) * p.b_bytes,
                                 static_cast<uint32_t>(p.b_bytes), bar_full_b + 8 * s);
                    }
                    __syncwarp();
                }
        }
    }

    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

size_t conv_ric_persist_smem_bytes(const ConvParams& p) { return rp_smem(p.sb, p.b_bytes, p.Cout).total + 1024; }

cudaError_t launch_conv_ric_persist(const ConvParams& p, cudaStream_t stream) {
    static bool attr_set[64] = {};
    static int sm_count[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64) dev = 0;
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv_ric_persist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_ric_persist_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_ric_persist_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    if (!p.ric || p.ks < 1 || p.ks > 3 || p.sb / p.ks < 2 || p.sb > kMaxStagesB || 2 * p.ks * p.Cout > 512 ||
        p.tmem_cols < 2 * p.ks * p.Cout || p.nchunks != p.nblocks * 9 || conv_ric_persist_smem_bytes(p) > 227 * 1024)
        return cudaErrorInvalidConfiguration;
    const int tiles = ((p.Wout + kTileW - 1) / kTileW) * ((p.Hout + kTileH - 1) / kTileH) * p.B;
    const int ctas = tiles < sm_count[dev] ? tiles : sm_count[dev];
    const size_t smem = conv_ric_persist_smem_bytes(p);
    if (p.exact) conv_ric_persist_kernel<2><<<ctas, kRpThreads, smem, stream>>>(p);
    else if (p.ric == 2) conv_ric_persist_kernel<3><<<ctas, kRpThreads, smem, stream>>>(p);
    else conv_ric_persist_kernel<1><<<ctas, kRpThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace dsu

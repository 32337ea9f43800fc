// Persistent first-layer convolution (conv0 of GeneratorJ, models.py:44-46 / 92-94: 7x7, stride 1, <= 8 input
// channels) WITHOUT im2col: the tensor core reads the A operand straight out of a pixel-linear halo tile.
//
// With <= 8 input channels a pixel is ONE 16-byte K slot.  A K-major no-swizzle UMMA operand is made of
// 8-row x 16-byte core matrices and the hardware computes (tools/umma_probe_ns.cu, profiles/r01m_umma_probe_ns.log)
//     addr(row, k) = start + (row / 8) * SBO + (k / 8) * LBO + (row % 8) * 16 + (k % 8) * 2
// for any 16-byte aligned start / LBO / SBO - overlapping core matrices included.  Store the halo tile of a
// 16 x 8 output tile as [16 + k - 1 rows][16 pixels][8 ch] fp16 (row pitch 256 B) and set LBO = 16, SBO = 256,
// start = &halo[kh][2 * step]: accumulator row (y, x) then reads pixels (y + kh, x + 2*step .. +1), i.e. the 7
// taps of kernel row kh are one contiguous K = 56 (+8 zero-weight) slice - 4 K=16 MMAs per kernel row, 28 per
// tile, no im2col copy, no per-chunk producer/consumer handshake.  The tap-mode kernel (conv_umma.cu) gathers the
// same 49 slots per output pixel with 7168 16-byte cp.async per tile (1.07 ms for 16 x 512 x 512), a shared-memory
// im2col variant of this kernel measured 0.86-1.0 ms (0.35 ms of it barrier skeleton).
//
// One CTA per SM walks a static tile list.  Warps 0-3 stage halo tiles (cp.async, zero fill outside the image,
// ring of p.sa buffers), warps 12-15 issue the MMAs (kernel rows kh % ks, one TMEM accumulator each, summed by the
// epilogue; warp 12 loads the whole weight matrix once), warps 4-7 / 8-11 are two epilogue groups that take even /
// odd tiles (2 or 4 accumulator sets in TMEM).
#include "conv_device.cuh"

namespace dsu {

namespace {

constexpr int kFirstIssuers = 4;
constexpr int kFirstThreads = (4 + 8 + kFirstIssuers) * 32;
constexpr int kFirstMaxHalo = 6;
constexpr int kFirstMaxSets = 4;
constexpr int kFirstBars = 2 * kFirstMaxHalo + 2 * kFirstMaxSets + 1;
constexpr int kFirstHaloW = 16;                 // pixels per halo row (8 outputs + 7 taps, padded): SBO = 256 B

struct FirstSmem {
    uint32_t h0, b0, par, bars, total;
};

__host__ __device__ inline FirstSmem first_smem(int na, int halo_bytes, int nchunks, int b_bytes, int cout) {
    FirstSmem L;
    L.b0 = 0;                                   // weight tiles (SWIZZLE_128B) at the 1024-aligned base
    L.h0 = nchunks * b_bytes;
    L.par = L.h0 + na * halo_bytes;
    L.bars = (L.par + (7 * cout + 4) * 4 + 15u) & ~15u;
    L.total = L.bars + (kFirstBars + 1) * 8;
    return L;
}

}  // namespace

__global__ void __launch_bounds__(kFirstThreads, 1)
conv_first_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_u32);
    const int NA = p.sa, C = p.Cout, NI = p.ks, KR = p.nchunks;    // KR kernel rows = weight chunks
    const FirstSmem L = first_smem(NA, p.halo_bytes, KR, p.b_bytes, C);
    float* s_par = reinterpret_cast<float*>(smem + L.par);
    const uint32_t bar_full = base + L.bars;
    const uint32_t bar_empty = bar_full + kFirstMaxHalo * 8;
    const uint32_t bar_acc_full = bar_empty + kFirstMaxHalo * 8;                // [kFirstMaxSets]
    const uint32_t bar_acc_empty = bar_acc_full + kFirstMaxSets * 8;            // [kFirstMaxSets]
    const uint32_t bar_w = bar_acc_empty + kFirstMaxSets * 8;
    // accumulator sets in TMEM: 4 when they fit (each epilogue group then alternates between two sets and never waits
    // for the issuer round trip of the set it just released), else 2
    const int NSETS = p.ns;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bars + kFirstBars * 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_x = (p.Wout + 7) / 8, tiles_y = (p.Hout + 15) / 16;
    const int tiles_per_frame = tiles_x * tiles_y;
    const int total_tiles = tiles_per_frame * p.B;
    const int my_tiles = (total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    constexpr int kProd = 128, kEpi = 128;

    if (warp == 12) {
        if (lane == 0) {
            for (int s = 0; s < NA; ++s) {
                mbar_init(bar_full + 8 * s, kProd / 32);        // one arrive per producer warp
                mbar_init(bar_empty + 8 * s, NI);
            }
            for (int s = 0; s < NSETS; ++s) {
                mbar_init(bar_acc_full + 8 * s, NI);
                mbar_init(bar_acc_empty + 8 * s, kEpi);
            }
            mbar_init(bar_w, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
        tmem_relinquish();
    } else if (warp >= 4 && warp < 8) {
        load_epilogue_params(p, s_par, tid - 128, kEpi);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto tile_coords = [&](int it, int& n, int& ty0, int& tx0) {
        const int t = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
        n = t / tiles_per_frame;
        const int r = t - n * tiles_per_frame;
        ty0 = (r / tiles_x) * 16;
        tx0 = (r % tiles_x) * 8;
    };

    if (warp < 4) {
        // ======================================================== halo producers (completion lags the issue by LAG tiles)
        const int HR = p.halo_rows * kFirstHaloW;           // halo entries (pixels) per tile
        const Slot sl0 = p.slots[0];                        // every valid slot reads the same 8-channel group
        const Seg sg = p.seg[sl0.seg];
        const __half* sbase = sg.ptr + sl0.choff;
        // split-fp16: the lo plane of the same pixels goes to the second half of the halo buffer
        const __half* sbase_lo = p.exact ? p.seg[sl0.seg + kMaxSeg / 2].ptr + sl0.choff : nullptr;
        const uint32_t plane_bytes = static_cast<uint32_t>(HR) * 16u;
        const int LAG = NA >= 3 ? 2 : 1;
        auto publish = [&](int it) {                        // this thread's copies of tile `it` have landed
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * (it % NA));
        };
        for (int it = 0; it < my_tiles; ++it) {
            const int hs = it % NA;
            if (it >= NA) mbar_wait(bar_empty + 8 * hs, ((it / NA) - 1) & 1);
            int n, ty0, tx0;
            tile_coords(it, n, ty0, tx0);
            const size_t frame_in = static_cast<size_t>(n) * p.Hin * p.Win;
            const uint32_t dst0 = base + L.h0 + hs * p.halo_bytes;
            for (int e = tid; e < HR; e += kProd) {
                const int hy = e / kFirstHaloW, hx = e % kFirstHaloW;
                const int vy = ty0 - p.pad + hy, vx = tx0 - p.pad + hx;
                const bool ok = static_cast<unsigned>(vy) < static_cast<unsigned>(p.Hv) &&
                                static_cast<unsigned>(vx) < static_cast<unsigned>(p.Wv);
                const size_t pix = frame_in + static_cast<size_t>(vy) * p.Win + vx;
                const __half* src = ok ? sbase + pix * sg.pitch : sbase;
                cp_async16(dst0 + e * 16, src, ok ? 16u : 0u);
                if (sbase_lo) cp_async16(dst0 + plane_bytes + e * 16, ok ? sbase_lo + pix * sg.pitch : sbase_lo, ok ? 16u : 0u);
            }
            cp_async_commit();
            if (it >= LAG) {
                if (LAG == 2) cp_async_wait<2>(); else cp_async_wait<1>();
                publish(it - LAG);
            }
        }
        cp_async_wait<0>();
        for (int it = (my_tiles > LAG ? my_tiles - LAG : 0); it < my_tiles; ++it) publish(it);
    } else if (warp < 12) {
        // ======================================================== epilogue: group 0 (warps 4-7) even tiles, group 1 odd tiles
        const int grp = (warp - 4) >> 2;
        const int quad = warp & 3;                          // TMEM lane quadrant this warp may read
        const int r = quad * 32 + lane;
        for (int it = grp; it < my_tiles; it += 2) {
            int n, ty0, tx0;
            tile_coords(it, n, ty0, tx0);
            const int set = it % NSETS;
            mbar_wait(bar_acc_full + 8 * set, (it / NSETS) & 1);
            tc_fence_after();
            const uint32_t t_set = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(set * NI * C);
            if (p.exact) epilogue_row<kEpiSplit>(p, s_par, t_set, n, ty0 + (r >> 3), tx0 + (r & 7), 0, NI, C, 1);
            else epilogue_row<kEpiFp16>(p, s_par, t_set, n, ty0 + (r >> 3), tx0 + (r & 7), 0, NI, C, 1);
            tc_fence_before();
            mbar_arrive(bar_acc_empty + 8 * set);
        }
    } else {
        // ======================================================== MMA issuers: kernel row kh belongs to issuer kh % NI
        const int wi = warp - 12;
        if (wi == 0 && elect_one()) {                       // the whole weight matrix, once per CTA
            mbar_arrive_expect_tx(bar_w, static_cast<uint32_t>(KR * p.b_bytes));
            for (int q = 0; q < KR; ++q)
                bulk_g2s(base + L.b0 + q * p.b_bytes, p.wpack + static_cast<size_t>(q) * p.b_bytes, static_cast<uint32_t>(p.b_bytes), bar_w);
        }
        __syncwarp();
        if (wi < NI) {
            const uint32_t idesc = umma_idesc_f16(kTileM, C);
            const uint32_t km = p.kmask_full;
            constexpr uint32_t kPitch = kFirstHaloW * 16;   // halo row pitch in bytes = SBO
            mbar_wait(bar_w, 0);
            for (int it = 0; it < my_tiles; ++it) {
                const int set = it % NSETS, hs = it % NA;
                if (it >= NSETS) {
                    mbar_wait(bar_acc_empty + 8 * set, ((it / NSETS) - 1) & 1);
                    tc_fence_after();
                }
                mbar_wait(bar_full + 8 * hs, (it / NA) & 1);
                tc_fence_after();
                const uint32_t d_addr = tmem_base + static_cast<uint32_t>((set * NI + wi) * C);
                const uint32_t h_addr = base + L.h0 + hs * p.halo_bytes;
                const uint32_t lo_off = static_cast<uint32_t>(p.halo_rows) * kPitch;     // lo plane of the halo tile (split-fp16)
                const uint32_t wlo_off = static_cast<uint32_t>(C) * 128u;                // W_lo tile of a kernel row
                uint32_t acc = 0;
                for (int kh = wi; kh < KR; kh += NI) {
                    const uint32_t a_addr = h_addr + kh * kPitch;
                    const uint32_t b_addr = base + L.b0 + kh * p.b_bytes;
                    if (elect_one()) {
                        // K step s: pixels x + 2s, x + 2s + 1 (A: +32 B) against weight slots 2s, 2s + 1 (B: +32 B inside the swizzled row)
                        const uint64_t da0 = umma_desc_noswizzle(a_addr, 16, kPitch), db0 = umma_desc_sw128(b_addr, 1024);
                        uint32_t a2 = acc;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if ((km >> k) & 1) { umma_f16(d_addr, da0 + 2 * k, db0 + 2 * k, idesc, a2); a2 = 1u; }
                        if (p.exact) {                      // + a_lo x W_hi + a_hi x W_lo
                            const uint64_t dal = umma_desc_noswizzle(a_addr + lo_off, 16, kPitch), dbl = umma_desc_sw128(b_addr + wlo_off, 1024);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if ((km >> k) & 1) {
                                    umma_f16(d_addr, dal + 2 * k, db0 + 2 * k, idesc, 1u);
                                    umma_f16(d_addr, da0 + 2 * k, dbl + 2 * k, idesc, 1u);
                                }
                        }
                    }
                    acc = 1u;
                    __syncwarp();
                }
                if (elect_one()) {
                    umma_commit(bar_empty + 8 * hs);        // halo buffer free when this issuer's MMAs retire
                    umma_commit(bar_acc_full + 8 * set);
                }
                __syncwarp();
            }
        }
        tc_fence_before();
    }

    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

size_t conv_first_smem_bytes(const ConvParams& p) {
    return first_smem(p.sa, p.halo_bytes, p.nchunks, p.b_bytes, p.Cout).total + 1024;
}

// Expects: stride 1 (split-fp16: second halo plane + a W_lo tile behind every W_hi tile, 3 MMAs per K step), no fused upsampling, one weight chunk per kernel row (slot j of chunk kh = tap (kh, j),
// all on the same 8-channel group), ksize <= 8, p.ks = issuers (1..4, <= ksize), p.sa = halo buffers (2..6),
// p.ns = accumulator sets (2 or 4), p.halo_rows = 16 + ksize - 1, p.halo_bytes >= halo_rows * 256,
// p.tmem_cols >= p.ns * p.ks * Cout.
cudaError_t launch_conv_first(const ConvParams& p, cudaStream_t stream) {
    static bool attr_set[64] = {};
    static int sm_count[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    if (p.stride != 1 || p.up != 0 || p.ric || p.ksize < 1 || p.ksize > 8 || p.nchunks != p.ksize || p.pad != (p.ksize - 1) / 2 ||
        p.sa < 2 || p.sa > kFirstMaxHalo || p.ks < 1 || p.ks > kFirstIssuers || p.ks > p.ksize || (p.ns != 2 && p.ns != 4) || p.ns * p.ks * p.Cout > 512 ||
        p.tmem_cols < p.ns * p.ks * p.Cout || p.halo_rows != 16 + p.ksize - 1 || p.halo_bytes < p.halo_rows * kFirstHaloW * 16 * (p.exact ? 2 : 1) || p.b_bytes != p.Cout * 128 * (p.exact ? 2 : 1) ||
        (p.halo_bytes & 127) || p.kmask_full != p.kmask_last || conv_first_smem_bytes(p) > 227 * 1024)
        return cudaErrorInvalidConfiguration;
    const int tiles = ((p.Wout + 7) / 8) * ((p.Hout + 15) / 16) * p.B;
    const int ctas = tiles < sm_count[dev < 64 ? dev : 0] ? tiles : sm_count[dev < 64 ? dev : 0];
    conv_first_kernel<<<ctas, kFirstThreads, conv_first_smem_bytes(p), stream>>>(p);
    return cudaGetLastError();
}

}  // namespace dsu

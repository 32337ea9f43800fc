// EXPERIMENTAL (DSU_RIC_FIRST=1, fp16 mode; compiled in but not yet run on hardware - see DESIGN.md 7b):
// stage-1 conv0 (GeneratorJ_RIC, models.py:302: rotation-invariant deformable 3x3 over the <= 8-channel network input)
// as ONE persistent kernel, without the 72-channel tap-expanded buffer that `ric_expand` writes to HBM and the 1x1
// contraction then reads back (0.30 + 0.50 ms per 16 x 512 x 512 batch, 0.6 GB each way).
//
// With <= 8 input channels a tap is one 16-byte K slot, so the A operand of a 128-pixel tile is 128 x (9 taps + 1 zero pad)
// slots = 20 KB.  The producers (one thread per pixel) do exactly what frames.cu ric_expand_kernel does - 3x3 neighbourhood,
// zero outside the image, fp32 bilinear blend of the 8 circle taps in octant-rotated order, fp16 rounding - but store the
// slots into shared memory in the K-major NO-SWIZZLE core-matrix layout  [pixel / 8][tap][pixel % 8][16 B]
// (tools/umma_probe_ns.cu: addr = start + (row/8) SBO + (k/8) LBO + (row%8) 16 + (k%8) 2 with LBO = 128, SBO = 1280).
// Five K=16 MMAs (2 taps each) per tile against the weight tiles the "expanded" layer already packs (SWIZZLE_128B:
// chunk 0 = taps 0-7, chunk 1 = tap 8), N = Cout.
//
// Roles (416 threads): warps 0-3 producers (ring of p.sa A tiles), warps 4-7 / 8-11 two epilogue groups (even / odd tiles,
// p.ns accumulator sets in TMEM), warp 12 MMA issuer + one-time weight load.
#include "conv_device.cuh"

namespace dsu {

namespace {

constexpr int kRfThreads = 13 * 32;                // 152 registers per thread: the producers keep 72 fp32 neighbours live
constexpr int kRfMaxStages = 6;
constexpr int kRfMaxSets = 4;
constexpr int kRfBars = 2 * kRfMaxStages + 2 * kRfMaxSets + 1;
constexpr int kRfTaps = 10;                         // 9 taps + one zero slot: K = 80 = 5 MMA K-steps
constexpr int kRfGroupBytes = kRfTaps * 128;        // one 8-pixel row group of the A tile = SBO
constexpr int kRfABytes = 16 * kRfGroupBytes;       // 128 pixels

struct RfSmem {
    uint32_t b0, a0, par, bars, total;
};

__host__ __device__ inline RfSmem rf_smem(int na, int nchunks, int b_bytes, int cout) {
    RfSmem L;
    L.b0 = 0;                                       // SWIZZLE_128B weight tiles at the 1024-aligned base
    L.a0 = nchunks * b_bytes;
    L.par = L.a0 + na * kRfABytes;
    L.bars = (L.par + (7 * cout + 4) * 4 + 15u) & ~15u;
    L.total = L.bars + (kRfBars + 1) * 8;
    return L;
}

}  // namespace

__global__ void __launch_bounds__(kRfThreads, 1)
conv_ric_first_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_u32);
    const int NA = p.sa, C = p.Cout, NSETS = p.ns, NQ = p.nchunks;
    const RfSmem L = rf_smem(NA, NQ, p.b_bytes, C);
    float* s_par = reinterpret_cast<float*>(smem + L.par);
    const uint32_t bar_full = base + L.bars;
    const uint32_t bar_empty = bar_full + kRfMaxStages * 8;
    const uint32_t bar_acc_full = bar_empty + kRfMaxStages * 8;
    const uint32_t bar_acc_empty = bar_acc_full + kRfMaxSets * 8;
    const uint32_t bar_w = bar_acc_empty + kRfMaxSets * 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bars + kRfBars * 8);

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
                mbar_init(bar_empty + 8 * s, 1);
            }
            for (int s = 0; s < NSETS; ++s) {
                mbar_init(bar_acc_full + 8 * s, 1);
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
        // ======================================================== producers: thread m = tile pixel (m >> 3, m & 7)
        const int m = tid;
        const Seg sg = p.seg[0];
        const __half* sbase = sg.ptr + p.raw_choff;
        const uint32_t row_off = static_cast<uint32_t>(m >> 3) * kRfGroupBytes + static_cast<uint32_t>(m & 7) * 16u;
        for (int it = 0; it < my_tiles; ++it) {
            const int s = it % NA;
            int n, ty0, tx0;
            tile_coords(it, n, ty0, tx0);
            const int oy = ty0 + (m >> 3), ox = tx0 + (m & 7);
            const bool live = oy < p.Hout && ox < p.Wout;
            const size_t frame_in = static_cast<size_t>(n) * p.Hin * p.Win;
            // ---- 3x3 neighbourhood, zero outside the image (torchvision's border rule), as fp32
            float nf[9][8];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int vy = oy + k / 3 - 1, vx = ox + k % 3 - 1;
                const bool ok = live && static_cast<unsigned>(vy) < static_cast<unsigned>(p.Hin) &&
                                static_cast<unsigned>(vx) < static_cast<unsigned>(p.Win);
                const uint4 raw = ldg128_if(sbase + (frame_in + static_cast<size_t>(ok ? vy : 0) * p.Win + (ok ? vx : 0)) * sg.pitch, ok);
                unpack8(raw, nf[k]);
            }
            int oct = 0;
            float2 lyx[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) lyx[t] = make_float2(0.0f, 0.0f);
            if (live) {
                const size_t e = static_cast<size_t>(oy) * p.Wout + ox;
                const float4* tp = reinterpret_cast<const float4*>(p.ric_lyx + e * 8);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float4 v = __ldg(tp + t);
                    lyx[2 * t] = make_float2(v.x, v.y);
                    lyx[2 * t + 1] = make_float2(v.z, v.w);
                }
                oct = __ldg(p.ric_oct + e);
            }
            if (it >= NA) mbar_wait(bar_empty + 8 * s, ((it / NA) - 1) & 1);
            uint8_t* dst = smem + L.a0 + s * kRfABytes + row_off;
            auto put = [&](int tap, const float* v) {
                *reinterpret_cast<uint4*>(dst + tap * 128) =
                    make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
            };
            put(4, nf[4]);                                                          // centre tap
            *reinterpret_cast<uint4*>(dst + 9 * 128) = make_uint4(0, 0, 0, 0);       // K padding slot (zero weights, finite data)
#pragma unroll
            for (int mm = 0; mm < 8; ++mm) {                                         // same arithmetic as frames.cu ric_expand_kernel
                const float ly = lyx[mm].x, lx = lyx[mm].y;
                const float hy = 1.0f - ly, hx = 1.0f - lx;
                const float w00 = hy * hx, w01 = hy * lx, w10 = ly * hx, w11 = ly * lx;
                const int r0 = (mm >= 2 && mm <= 5) ? 0 : 1, c0 = (mm >= 4) ? 0 : 1;
                float o[8];
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    o[c] = fmaf(w11, nf[(r0 + 1) * 3 + c0 + 1][c], fmaf(w10, nf[(r0 + 1) * 3 + c0][c],
                           fmaf(w01, nf[r0 * 3 + c0 + 1][c], w00 * nf[r0 * 3 + c0][c])));
                const int kq = (mm - oct) & 7;
                put(kq + (kq >> 2), o);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * s);
        }
    } else if (warp < 12) {
        // ======================================================== epilogue: group 0 (warps 4-7) even tiles, group 1 odd tiles
        const int grp = (warp - 4) >> 2;
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        for (int it = grp; it < my_tiles; it += 2) {
            int n, ty0, tx0;
            tile_coords(it, n, ty0, tx0);
            const int set = it % NSETS;
            mbar_wait(bar_acc_full + 8 * set, (it / NSETS) & 1);
            tc_fence_after();
            const uint32_t t_set = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(set * C);
            epilogue_row<kEpiFp16>(p, s_par, t_set, n, ty0 + (r >> 3), tx0 + (r & 7), 0, 1, 0, 1);
            tc_fence_before();
            mbar_arrive(bar_acc_empty + 8 * set);
        }
    } else if (warp == 12) {
        // ======================================================== MMA issuer (+ the whole weight matrix, once per CTA)
        if (elect_one()) {
            mbar_arrive_expect_tx(bar_w, static_cast<uint32_t>(NQ * p.b_bytes));
            for (int q = 0; q < NQ; ++q)
                bulk_g2s(base + L.b0 + q * p.b_bytes, p.wpack + static_cast<size_t>(q) * p.b_bytes, static_cast<uint32_t>(p.b_bytes), bar_w);
        }
        __syncwarp();
        const uint32_t idesc = umma_idesc_f16(kTileM, C);
        mbar_wait(bar_w, 0);
        for (int it = 0; it < my_tiles; ++it) {
            const int set = it % NSETS, s = it % NA;
            if (it >= NSETS) {
                mbar_wait(bar_acc_empty + 8 * set, ((it / NSETS) - 1) & 1);
                tc_fence_after();
            }
            mbar_wait(bar_full + 8 * s, (it / NA) & 1);
            tc_fence_after();
            const uint32_t d_addr = tmem_base + static_cast<uint32_t>(set * C);
            const uint32_t a_addr = base + L.a0 + s * kRfABytes;
            if (elect_one()) {
                // K-step j = taps 2j, 2j+1: A start + 2 core matrices (256 B); B = chunk j / 4 (taps 0-7 | 8-9), step j % 4
                const uint64_t da0 = umma_desc_noswizzle(a_addr, 128, kRfGroupBytes);
                const uint64_t db0 = umma_desc_sw128(base + L.b0, 1024), db1 = umma_desc_sw128(base + L.b0 + p.b_bytes, 1024);
                umma_f16(d_addr, da0, db0, idesc, 0u);
                umma_f16(d_addr, da0 + 16, db0 + 2, idesc, 1u);
                umma_f16(d_addr, da0 + 32, db0 + 4, idesc, 1u);
                umma_f16(d_addr, da0 + 48, db0 + 6, idesc, 1u);
                umma_f16(d_addr, da0 + 64, db1, idesc, 1u);
                umma_commit(bar_empty + 8 * s);
                umma_commit(bar_acc_full + 8 * set);
            }
            __syncwarp();
        }
        tc_fence_before();
    }

    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

size_t conv_ric_first_smem_bytes(const ConvParams& p) { return rf_smem(p.sa, p.nchunks, p.b_bytes, p.Cout).total + 1024; }

// Expects the weight tiles of the tap-expanded 1x1 layer (2 chunks: taps 0-7, tap 8; kmask_last = 1), fp16 mode, p.seg[0] = the
// raw network-input buffer with p.raw_choff = its first channel, p.ric_lyx / p.ric_oct = level-0 stencil, p.sa = A-tile ring
// (2..6), p.ns = accumulator sets (2 or 4), p.tmem_cols >= p.ns * Cout.
cudaError_t launch_conv_ric_first(const ConvParams& p, cudaStream_t stream) {
    static bool attr_set[64] = {};
    static int sm_count[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64) dev = 0;
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv_ric_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    if (p.exact || p.nchunks != 2 || p.kmask_full != 0xFu || p.kmask_last != 0x1u || p.up != 0 || p.sa < 2 || p.sa > kRfMaxStages ||
        (p.ns != 2 && p.ns != 4) || p.ns * p.Cout > 512 || p.tmem_cols < p.ns * p.Cout || !p.ric_lyx || !p.ric_oct ||
        p.Hin != p.Hout || p.Win != p.Wout || conv_ric_first_smem_bytes(p) > 227 * 1024)
        return cudaErrorInvalidConfiguration;
    const int tiles = ((p.Wout + 7) / 8) * ((p.Hout + 15) / 16) * p.B;
    const int ctas = tiles < sm_count[dev] ? tiles : sm_count[dev];
    conv_ric_first_kernel<<<ctas, kRfThreads, conv_ric_first_smem_bytes(p), stream>>>(p);
    return cudaGetLastError();
}

}  // namespace dsu

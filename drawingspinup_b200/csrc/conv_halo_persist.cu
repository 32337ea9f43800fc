// Halo-reuse convolution for the plain (stage-2) stride-1 layers, persistent: one CTA per SM walks a static list of output
// tiles.  A (16 + 2p) x (8 ns + 2p)-pixel halo tile per 64-channel block (split-fp16: 32 channels as [hi | lo]) is staged once
// in shared memory, one pixel per 128-byte row, and EVERY tap reads a shifted window of it through its UMMA descriptor (start
// row = kh W + kw + 8 s, 8-row-group stride = W x 128 B; tools/umma_probe.cu established that the descriptor reads linear rows
// from any 128-byte-aligned start and swizzles on absolute address bits).  The accumulators are double-buffered in TMEM so the
// epilogue of tile i (4 dedicated warps) overlaps the halo staging, weight streaming and MMAs of tile i + 1, and the prologue
// (barrier init, TMEM allocation, parameter load) is paid once per SM instead of once per tile.
//
// Roles: warps 0-3 halo producers, warps 4-7 epilogue, warps 8-11 MMA issuers (sub-tile x K-split; every K-split group has a
// private weight ring - a ring shared by consumer warps lets one that is a revolution ahead pass a parity wait on the previous
// phase), warp 12 weight loader.  TMEM: nsets (2, or 1 when two do not fit) x (ns*ks) accumulators x Cout columns (2 Cout in
// the n128 form of the split-fp16 Cout = 64 layers, see ConvParams::n128).
#include "conv_device.cuh"

namespace dsu {

namespace {

constexpr int kMaxHaloBufs = 3;
constexpr int kPersistBars = 2 * kMaxHaloBufs + 2 * kMaxStagesB + 4;

struct PersistSmem {
    uint32_t a0, b0, par, bars, total;
};

__host__ __device__ inline PersistSmem persist_smem(int na, int halo_bytes, int sb, int b_bytes, int cout) {
    PersistSmem L;
    L.a0 = 0;
    L.b0 = na * halo_bytes;
    L.par = L.b0 + sb * b_bytes;
    L.bars = (L.par + (7 * cout + 4) * 4 + 15u) & ~15u;
    L.total = L.bars + (kPersistBars + 1) * 8;
    return L;
}

}  // namespace

// kSub = true: one sub-pixel class of a fused nearest-x2 + 3x3 convolution (see ConvParams::sub) - the halo origin uses the
// asymmetric (pad_y, pad_x) and the epilogue scatters to every second output pixel; everything else is the same kernel.
template <bool kSub>
__global__ void __launch_bounds__(kThreadsHalo, 1)
conv_halo_persist_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_u32);
    const int NA = p.sa, SB = p.sb, C = p.Cout, NS = p.ns, KS = p.ks;
    const int NI = NS * KS;
    const int AC = p.n128 ? 2 * C : C;                         // accumulator columns per issuer
    const int NSETS = p.nsets;
    const int SBK = SB / KS;
    const int TPS = p.tps;                                   // taps per weight stage (> 1 only with ks == 1)
    const int stage_bytes = TPS * p.b_bytes;
    const PersistSmem L = persist_smem(NA, p.halo_bytes, SB, stage_bytes, C);
    float* s_par = reinterpret_cast<float*>(smem + L.par);
    const uint32_t bar_full_a = base + L.bars;
    const uint32_t bar_empty_a = bar_full_a + kMaxHaloBufs * 8;
    const uint32_t bar_full_b = bar_empty_a + kMaxHaloBufs * 8;
    const uint32_t bar_empty_b = bar_full_b + kMaxStagesB * 8;
    const uint32_t bar_acc_full = bar_empty_b + kMaxStagesB * 8;     // [2]
    const uint32_t bar_acc_empty = bar_acc_full + 16;                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bars + kPersistBars * 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int taps = p.ksize * p.ksize;
    const int tiles_x = (p.Wout + 8 * NS - 1) / (8 * NS), tiles_y = (p.Hout + 15) / 16;
    const int tiles_per_frame = tiles_x * tiles_y;
    const int total_tiles = tiles_per_frame * p.B;
    const int my_tiles = (total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    constexpr int kProd = 128, kEpi = 128;

    if (warp == 8) {
        if (lane == 0) {
            for (int s = 0; s < NA; ++s) {
                mbar_init(bar_full_a + 8 * s, kProd);
                mbar_init(bar_empty_a + 8 * s, NI);
            }
            for (int s = 0; s < SB; ++s) {
                mbar_init(bar_full_b + 8 * s, 1);
                mbar_init(bar_empty_b + 8 * s, NS);
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
    } else if (warp >= 4 && warp < 8) {
        load_epilogue_params(p, s_par, tid - 128, kEpi);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile it of this CTA -> (frame, tile row, tile column)
    auto tile_coords = [&](int it, int& n, int& ty0, int& tx0) {
        const int t = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
        n = t / tiles_per_frame;
        const int r = t - n * tiles_per_frame;
        ty0 = (r / tiles_x) * 16;
        tx0 = (r % tiles_x) * 8 * NS;
    };

    if (warp < 4) {
        // ======================================================== halo producers
        const int j = tid & 7;
        const int HW = p.halo_w, HR = p.halo_rows;
        int g = 0;                                       // halo blocks staged so far (ring position)
        for (int it = 0; it < my_tiles; ++it) {
            int n, ty0, tx0;
            tile_coords(it, n, ty0, tx0);
            const size_t frame_in = static_cast<size_t>(n) * p.Hin * p.Win;
            for (int b = 0; b < p.nblocks; ++b, ++g) {
                const int s = g % NA;
                if (g >= NA) mbar_wait(bar_empty_a + 8 * s, ((g / NA) - 1) & 1);
                const Slot sl = p.slots[b * 8 + j];
                const Seg sg = p.seg[sl.seg];
                const __half* sbase = sg.ptr + sl.choff;
                const uint32_t dst0 = base + L.a0 + s * p.halo_bytes;
                for (int row = tid >> 3; row < HR; row += kProd / 8) {
                    const int hy = row / HW, hx = row - hy * HW;
                    const int vy = ty0 - (kSub ? p.pad_y : p.pad) + hy, vx = tx0 - (kSub ? p.pad_x : p.pad) + hx;
                    const bool ok = sl.valid && static_cast<unsigned>(vy) < static_cast<unsigned>(p.Hv) &&
                                    static_cast<unsigned>(vx) < static_cast<unsigned>(p.Wv);
                    const size_t pix = frame_in + static_cast<size_t>(vy >> p.up) * p.Win + (vx >> p.up);
                    const __half* src = ok ? sbase + pix * sg.pitch : sbase;
                    cp_async16(dst0 + row * 128 + (static_cast<uint32_t>(j ^ (row & 7)) << 4), src, ok ? 16u : 0u);
                }
                cp_async_commit();
                cp_async_wait<0>();
                fence_proxy_async_smem();
                mbar_arrive(bar_full_a + 8 * s);
            }
        }
    } else if (warp < 8) {
        // ======================================================== epilogue warps (one per TMEM lane quadrant)
        const int quad = warp - 4;
        const int r = quad * 32 + lane;
        int set = 0;
        uint32_t sph = 0;                                // accumulator set of tile `it` and the phase of its barriers
        for (int it = 0; it < my_tiles; ++it) {
            int n, ty0, tx0;
            tile_coords(it, n, ty0, tx0);
            mbar_wait(bar_acc_full + 8 * set, sph);
            tc_fence_after();
            const uint32_t t_set = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(set * NI * AC);
            for (int s = 0; s < NS; ++s) {
                if (p.n128) epilogue_row<kEpiAll, kSub>(p, s_par, t_set + s * AC, n, ty0 + (r >> 3), tx0 + 8 * s + (r & 7), 0, 2, C, 1);
                else epilogue_row<kEpiAll, kSub>(p, s_par, t_set + s * C, n, ty0 + (r >> 3), tx0 + 8 * s + (r & 7), 0, KS, NS * C, 1);
            }
            tc_fence_before();
            mbar_arrive(bar_acc_empty + 8 * set);        // this accumulator set may be overwritten
            if (++set == NSETS) { set = 0; sph ^= 1u; }
        }
    } else if (warp < 8 + kIssuersHalo) {
        // ======================================================== MMA issuers (warp-uniform loops, elected lane issues)
        const int wi = warp - 8;
        if (wi < NI) {
            const int sub = wi % NS, ksp = wi / NS;
            const uint32_t idesc = umma_idesc_f16(kTileM, C);
            const uint32_t sbo = static_cast<uint32_t>(p.halo_w) * 128u;
            const uint32_t idesc2 = umma_idesc_f16(kTileM, 2 * C);
            int g = 0, cnt = 0;
            int set = 0;
            uint32_t sph = 1;                            // phase of the PREVIOUS use of this set's empty barrier
            for (int it = 0; it < my_tiles; ++it) {
                if (it >= NSETS) {                       // the epilogue of the tile that used this set must be done
                    mbar_wait(bar_acc_empty + 8 * set, sph);
                    tc_fence_after();
                }
                const uint32_t d_addr = tmem_base + static_cast<uint32_t>((set * NI + wi) * AC);
                uint32_t acc = 0;
                for (int b = 0; b < p.nblocks; ++b, ++g) {
                    const int s_a = g % NA;
                    const uint32_t km = b == p.nblocks - 1 ? p.kmask_last : p.kmask_full;
                    const uint32_t km2 = b == p.nblocks - 1 ? p.kmask2_last : p.kmask2_full;
                    mbar_wait(bar_full_a + 8 * s_a, (g / NA) & 1);
                    const uint32_t sub_addr = base + L.a0 + s_a * p.halo_bytes + static_cast<uint32_t>(sub) * 1024u;
                    // a weight stage holds TPS consecutive taps (one wait + one commit per stage: the ~450-cycle commit
                    // group cost measured in tools/umma_rate.cu is paid once per TPS taps)
                    for (int t0 = ksp * TPS; t0 < taps; t0 += KS * TPS, ++cnt) {
                        const int s_b = ksp * SBK + cnt % SBK;
                        const int nt = (taps - t0) < TPS ? (taps - t0) : TPS;
                        mbar_wait(bar_full_b + 8 * s_b, (cnt / SBK) & 1);
                        tc_fence_after();
                        for (int u = 0; u < nt; ++u) {           // warp-uniform address math, elected lane issues
                            const int t = t0 + u;
                            const int kh = t / p.ksize, kw = t - kh * p.ksize;
                            const uint32_t b_addr = base + L.b0 + s_b * stage_bytes + u * p.b_bytes;
                            const uint32_t a_addr = sub_addr + static_cast<uint32_t>(kh * p.halo_w + kw) * 128u;
                            if (elect_one()) {
                                const uint64_t da0 = umma_desc_sw128(a_addr, sbo);
                                const uint64_t db0 = umma_desc_sw128(b_addr, 1024);
                                if (p.n128) {
                                    // B tile [128 rows = W_hi ; W_lo][32 channels], no swizzle: K16 step k starts 256 B further
                                    uint32_t a2 = acc;
#pragma unroll
                                    for (int k = 0; k < 2; ++k)
                                        if ((km >> k) & 1) {
                                            const uint64_t dbn = umma_desc_noswizzle(b_addr + k * 256, 128, 512);
                                            umma_f16(d_addr, da0 + 2 * k, dbn, idesc2, a2);          // a_hi x [W_hi | W_lo]
                                            umma_f16(d_addr, da0 + 4 + 2 * k, dbn, idesc, 1u);       // a_lo x W_hi
                                            a2 = 1u;
                                        }
                                } else if (km2) {
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
                            }
                            acc = 1u;
                            __syncwarp();
                        }
                        if (elect_one()) umma_commit(bar_empty_b + 8 * s_b);   // same elected lane as the MMAs it tracks
                        acc = 1u;
                        __syncwarp();
                    }
                    if (elect_one()) {
                        umma_commit(bar_empty_a + 8 * s_a);
                        if (b == p.nblocks - 1) umma_commit(bar_acc_full + 8 * set);
                    }
                    __syncwarp();
                }
                if (++set == NSETS) { set = 0; sph ^= 1u; }
            }
        }
        tc_fence_before();
    } else {
        // ======================================================== weight loader (rings run on across tiles)
        int cnt[kIssuersHalo] = {0, 0, 0, 0};
        for (int it = 0; it < my_tiles; ++it)
            for (int b = 0; b < p.nblocks; ++b)
                for (int t0 = 0, grp = 0; t0 < taps; t0 += TPS, ++grp) {
                    const int k = grp % KS;
                    const int nt = (taps - t0) < TPS ? (taps - t0) : TPS;
                    int c = 0;
#pragma unroll
                    for (int i = 0; i < kIssuersHalo; ++i) if (i == k) { c = cnt[i]; cnt[i] = c + 1; }
                    const int s = k * SBK + c % SBK;
                    if (c >= SBK) mbar_wait(bar_empty_b + 8 * s, ((c / SBK) - 1) & 1);
                    if (elect_one()) {
                        const uint32_t bytes = static_cast<uint32_t>(nt * p.b_bytes);
                        mbar_arrive_expect_tx(bar_full_b + 8 * s, bytes);
                        bulk_g2s(base + L.b0 + s * stage_bytes, p.wpack + static_cast<size_t>(b * taps + t0) * p.b_bytes, bytes, bar_full_b + 8 * s);
                    }
                    __syncwarp();
                }
    }

    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

size_t conv_halo_persist_smem_bytes(const ConvParams& p) {
    return persist_smem(p.sa, p.halo_bytes, p.sb, p.tps * p.b_bytes, p.Cout).total + 1024;
}

cudaError_t launch_conv_halo_persist(const ConvParams& p, cudaStream_t stream) {
    static bool attr_set[64] = {};
    static int sm_count[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv_halo_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(conv_halo_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    if (p.sa < 1 || p.sa > kMaxHaloBufs || p.sb < 2 || p.sb > kMaxStagesB || p.ns < 1 || p.ks < 1 ||
        p.ns * p.ks > kIssuersHalo || p.sb / p.ks < 2 || p.stride != 1 || p.tps < 1 || p.tps > 4 || p.nsets < 1 || p.nsets > 2 ||
        p.nsets * p.ns * p.ks * (p.n128 ? 2 : 1) * p.Cout > 512 || p.tmem_cols < p.nsets * p.ns * p.ks * (p.n128 ? 2 : 1) * p.Cout ||
        (p.n128 && (p.ks != 1 || !p.exact || p.Cout != 64)) || conv_halo_persist_smem_bytes(p) > 227 * 1024 ||
        p.nchunks != p.nblocks * p.ksize * p.ksize || p.ks > p.ksize * p.ksize)
        return cudaErrorInvalidConfiguration;
    const int tiles = ((p.Wout + 8 * p.ns - 1) / (8 * p.ns)) * ((p.Hout + 15) / 16) * p.B;
    const int ctas = tiles < sm_count[dev < 64 ? dev : 0] ? tiles : sm_count[dev < 64 ? dev : 0];
    if (p.sub) {
        if (p.up != 0 || p.ksize != 2 || p.pad_y < 0 || p.pad_y > 1 || p.pad_x < 0 || p.pad_x > 1 || (p.sub_py | p.sub_px) & ~1)
            return cudaErrorInvalidConfiguration;
        conv_halo_persist_kernel<true><<<ctas, kThreadsHalo, conv_halo_persist_smem_bytes(p), stream>>>(p);
    } else {
        conv_halo_persist_kernel<false><<<ctas, kThreadsHalo, conv_halo_persist_smem_bytes(p), stream>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace dsu

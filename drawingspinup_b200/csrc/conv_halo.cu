// Halo-reuse implicit-GEMM convolution for plain stride-1 convs (stage 2: residual trunk, upconvs,
// the 7x7 conv_11 and the smoothers; models.py:118-127).
//
// Measured on B200 (tools/umma_probe.cu, profiles/r01_umma_probe.log): a K-major SWIZZLE_128B UMMA
// descriptor reads LINEAR 128-byte rows from its start address - which may be any multiple of
// 128 B - with an arbitrary 8-row-group stride, and applies the swizzle on absolute shared-memory
// address bits.  So an input halo tile stored one pixel per 128-byte row, (16+2p) x (8*ns+2p)
// pixels, row pitch = halo width, serves EVERY tap of a k x k convolution: tap (kh, kw) of
// sub-tile s is the same buffer read from row (kh*W + kw + 8*s) with group stride W*128.
// Each input pixel is fetched from L2 once per 64-channel block instead of once per tap (9x / 49x
// fewer cp.async), and each weight tile is shared by the ns side-by-side 16x8 sub-tiles of the CTA.
//
// Issue rate (tools/umma_rate.cu, profiles/r01_umma_rate.log): one warp sustains ~100 cycles per
// tcgen05.mma in SS mode, 3x the execution time of an N=64 MMA, so the CTA runs ns*ks issuing warps:
// warp (s, k) owns sub-tile s and the taps t with t % ks == k, and accumulates into its own TMEM
// column range; the epilogue adds the ks partial sums.
//
// Roles: warps 0-7 stage halo tiles (cp.async, zero fill outside the image; nearest-x2 upsampling,
// concat and the hi|lo split-fp16 layout are address arithmetic) and run the epilogue, warps 8-11
// issue tcgen05.mma, warp 12 streams weight tiles with bulk async copies.
#include "conv_device.cuh"

namespace dsu {

namespace {

constexpr int kMaxHaloBufs = 3;
constexpr int kHaloBars = 2 * kMaxHaloBufs + 2 * kMaxStagesB + 1;

struct HaloSmem {
    uint32_t a0, b0, par, bars, total;
};

__host__ __device__ inline HaloSmem halo_smem(int na, int halo_bytes, int sb, int b_bytes, int cout) {
    HaloSmem L;
    L.a0 = 0;
    L.b0 = na * halo_bytes;
    L.par = L.b0 + sb * b_bytes;
    L.bars = (L.par + (7 * cout + 4) * 4 + 15u) & ~15u;
    L.total = L.bars + (kHaloBars + 1) * 8;
    return L;
}

}  // namespace

__global__ void __launch_bounds__(kThreadsHalo, 1)
conv_halo_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_u32);
    const int NA = p.sa, SB = p.sb, C = p.Cout, NS = p.ns, KS = p.ks;
    const int NI = NS * KS;                                  // issuing warps
    const int SBK = SB / KS;                                 // weight stages per K-split ring (each ring has exactly one
                                                             // consumer group that sees every phase in order: no mbarrier parity aliasing)
    const HaloSmem L = halo_smem(NA, p.halo_bytes, SB, p.b_bytes, C);
    float* s_par = reinterpret_cast<float*>(smem + L.par);
    const uint32_t bar_full_a = base + L.bars;
    const uint32_t bar_empty_a = bar_full_a + kMaxHaloBufs * 8;
    const uint32_t bar_full_b = bar_empty_a + kMaxHaloBufs * 8;
    const uint32_t bar_empty_b = bar_full_b + kMaxStagesB * 8;
    const uint32_t bar_accum = bar_empty_b + kMaxStagesB * 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bars + kHaloBars * 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = blockIdx.z;
    const int ty0 = blockIdx.y * 16;
    const int tx0 = blockIdx.x * 8 * NS;
    const int taps = p.ksize * p.ksize;

    if (warp == 8) {
        if (lane == 0) {
            for (int s = 0; s < NA; ++s) {
                mbar_init(bar_full_a + 8 * s, kWorkers);
                mbar_init(bar_empty_a + 8 * s, NI);          // every issuer reads every halo block
            }
            for (int s = 0; s < SB; ++s) {
                mbar_init(bar_full_b + 8 * s, 1);
                mbar_init(bar_empty_b + 8 * s, NS);          // a tap's weight tile is read by the ns issuers of its K-split
            }
            mbar_init(bar_accum, NI);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
        tmem_relinquish();
    } else if (warp < 8) {
        load_epilogue_params(p, s_par, tid, kWorkers);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ======================================================== halo producers
        const int j = tid & 7;                    // 16-byte slot of the pixel row
        const size_t frame_in = static_cast<size_t>(n) * p.Hin * p.Win;
        const int HW = p.halo_w, HR = p.halo_rows;
        for (int b = 0; b < p.nblocks; ++b) {
            const int s = b % NA;
            if (b >= NA) mbar_wait(bar_empty_a + 8 * s, ((b / NA) - 1) & 1);
            const Slot sl = p.slots[b * 8 + j];
            const Seg sg = p.seg[sl.seg];
            const __half* sbase = sg.ptr + sl.choff;
            const uint32_t dst0 = base + L.a0 + s * p.halo_bytes;
            for (int row = tid >> 3; row < HR; row += kWorkers / 8) {
                const int hy = row / HW, hx = row - hy * HW;
                const int vy = ty0 - p.pad + hy, vx = tx0 - p.pad + hx;
                const bool ok = sl.valid && static_cast<unsigned>(vy) < static_cast<unsigned>(p.Hv) &&
                                static_cast<unsigned>(vx) < static_cast<unsigned>(p.Wv);
                const size_t pix = frame_in + static_cast<size_t>(vy >> p.up) * p.Win + (vx >> p.up);
                const __half* src = ok ? sbase + pix * sg.pitch : sbase;
                cp_async16(dst0 + row * 128 + (static_cast<uint32_t>(j ^ (row & 7)) << 4), src, ok ? 16u : 0u);
            }
            cp_async_commit();
            cp_async_wait<0>();                    // a block feeds k*k taps of MMA work: no need to lag the hand-off
            fence_proxy_async_smem();
            mbar_arrive(bar_full_a + 8 * s);
        }

        // ======================================================== epilogue
        mbar_wait(bar_accum, 0);
        tc_fence_after();
        const int quad = warp & 3;
        const int r = quad * 32 + lane;           // accumulator row: patch pixel (r >> 3, r & 7) of each sub-tile
        for (int s = 0; s < NS; ++s)
            epilogue_row(p, s_par, tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + s * C, n,
                         ty0 + (r >> 3), tx0 + 8 * s + (r & 7), warp >> 2, KS, NS * C);
        tc_fence_before();
    } else if (warp < 8 + kIssuersHalo) {
        // ======================================================== MMA issuers: warp (sub, ksp) = ((warp-8) % NS, (warp-8) / NS)
        // Warp-uniform loops, one elected lane issues (keeps descriptors in uniform registers).
        const int wi = warp - 8;
        if (wi < NI) {
            const int sub = wi % NS, ksp = wi / NS;
            const uint32_t idesc = umma_idesc_f16(kTileM, C);
            const uint32_t sbo = static_cast<uint32_t>(p.halo_w) * 128u;
            const uint32_t d_addr = tmem_base + static_cast<uint32_t>(wi * C);      // partial accumulator (ksp, sub)
            uint32_t acc = 0;
            int cnt = 0;                                         // taps consumed from this group's ring
            for (int b = 0; b < p.nblocks; ++b) {
                const int s_a = b % NA;
                const uint32_t km = b == p.nblocks - 1 ? p.kmask_last : p.kmask_full;
                const uint32_t km2 = b == p.nblocks - 1 ? p.kmask2_last : p.kmask2_full;
                mbar_wait(bar_full_a + 8 * s_a, (b / NA) & 1);
                const uint32_t sub_addr = base + L.a0 + s_a * p.halo_bytes + static_cast<uint32_t>(sub) * 1024u;   // 8 pixels per sub-tile
                for (int t = ksp; t < taps; t += KS, ++cnt) {
                    const int s_b = ksp * SBK + cnt % SBK;
                    const int kh = t / p.ksize, kw = t - kh * p.ksize;
                    mbar_wait(bar_full_b + 8 * s_b, (cnt / SBK) & 1);
                    tc_fence_after();
                    const uint32_t b_addr = base + L.b0 + s_b * p.b_bytes;
                    const uint32_t a_addr = sub_addr + static_cast<uint32_t>(kh * p.halo_w + kw) * 128u;
                    if (elect_one()) {
                        // descriptors are built once per tap; a K=16 step advances the start-address field by 32 B (= 2)
                        const uint64_t da0 = umma_desc_sw128(a_addr, sbo);
                        const uint64_t db0 = umma_desc_sw128(b_addr, 1024);
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
                        umma_commit(bar_empty_b + 8 * s_b);
                    }
                    acc = 1u;
                    __syncwarp();
                }
                if (elect_one()) {
                    umma_commit(bar_empty_a + 8 * s_a);
                    if (b == p.nblocks - 1) umma_commit(bar_accum);
                }
                __syncwarp();
            }
        }
        tc_fence_before();
    } else {
        // ======================================================== weight loader: tap t goes to ring t % KS
        int cnt[kIssuersHalo] = {0, 0, 0, 0};
        for (int b = 0; b < p.nblocks; ++b)
            for (int t = 0; t < taps; ++t) {
                const int k = t % KS;
                int c = 0;
#pragma unroll
                for (int i = 0; i < kIssuersHalo; ++i) if (i == k) { c = cnt[i]; cnt[i] = c + 1; }
                const int s = k * SBK + c % SBK;
                if (c >= SBK) mbar_wait(bar_empty_b + 8 * s, ((c / SBK) - 1) & 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(bar_full_b + 8 * s, static_cast<uint32_t>(p.b_bytes));
                    bulk_g2s(base + L.b0 + s * p.b_bytes, p.wpack + static_cast<size_t>(b * taps + t) * p.b_bytes,
                             static_cast<uint32_t>(p.b_bytes), bar_full_b + 8 * s);
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

size_t conv_halo_smem_bytes(const ConvParams& p) {
    return halo_smem(p.sa, p.halo_bytes, p.sb, p.b_bytes, p.Cout).total + 1024;
}

cudaError_t launch_conv_halo(const ConvParams& p, cudaStream_t stream) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    if (p.sa < 1 || p.sa > kMaxHaloBufs || p.sb < 2 || p.sb > kMaxStagesB || p.ns < 1 || p.ks < 1 ||
        p.ns * p.ks > kIssuersHalo || p.sb / p.ks < 2 || p.stride != 1 || p.ns * p.ks * p.Cout > 512 || p.tmem_cols < p.ns * p.ks * p.Cout ||
        conv_halo_smem_bytes(p) > 227 * 1024 || p.nchunks != p.nblocks * p.ksize * p.ksize || p.ks > p.ksize * p.ksize)
        return cudaErrorInvalidConfiguration;
    dim3 grid((p.Wout + 8 * p.ns - 1) / (8 * p.ns), (p.Hout + 15) / 16, p.B);
    conv_halo_kernel<<<grid, kThreadsHalo, conv_halo_smem_bytes(p), stream>>>(p);
    return cudaGetLastError();
}

}  // namespace dsu

// Fused convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05, accumulators in TMEM).
//
// One CTA computes a kTileH x kTileW patch (128 output pixels = UMMA M) of one frame for all
// Cout channels (UMMA N).  K = taps x input channels, walked in 64-element chunks:
//   * warps 0-7 (256 threads) PRODUCE the A operand: one smem row (128 B, SWIZZLE_128B, K-major)
//     per output pixel.  Each 16-byte slot of a row is 8 channels of one tap of one concat
//     segment (slot table), so plain 3x3 / 7x7 taps, stride 2, fused nearest-x2 upsampling and
//     channel concatenation are pure address arithmetic (cp.async with zero fill at the border).
//     (Used for the stride-2 convolutions of stage 2 and as the generic fallback of its first layer; the stride-1 layers run
//     the halo-reuse kernel, stage 1 the tensor-memory RIC kernel conv_ric_tm.cu.)
//   * warp 9 streams the pre-swizzled weight tile (B operand) with 1-D bulk async copies.
//   * warp 8 issues tcgen05.mma (one thread) and commits stage-release / accumulator-ready
//     mbarriers.
//   * the producer warps then become the epilogue: tcgen05.ld the accumulator row of their pixel,
//     apply folded BN / activation / residual, write fp16 NHWC (hi [+lo] planes), the fp32 residual
//     stream, or the fused conv_12 1x1 + tanh + uint8 composite tail.
// "Exact" mode (split fp16): a K chunk holds 32 channels as [a_hi | a_lo], its weight tile [W_hi | W_lo] in one
// 128-byte row; A steps 0-3 are issued against B steps 0,1,0,1 and A steps 0,1 again against B steps 2,3 into the
// same accumulator: a_hi*W_hi + a_lo*W_hi + a_hi*W_lo, fp32-grade products at 3x the tensor work.
#include "conv_device.cuh"

namespace dsu {

namespace {

constexpr int kNumBars = 2 * kMaxStagesA + 2 * kMaxStagesB + 1;

struct SmemLayout {
    uint32_t a0, b0, par, bars;   // byte offsets from the 1024-aligned base
    uint32_t total;
};

__host__ __device__ inline SmemLayout smem_layout(int sa, int sb, int b_bytes, int cout, int nchunks) {
    SmemLayout L;
    L.a0 = 0;
    L.b0 = L.a0 + sa * kABytes;
    L.par = L.b0 + sb * b_bytes;
    (void)nchunks;
    L.bars = (L.par + (7 * cout + 4) * 4 + 15u) & ~15u;
    L.total = L.bars + (kNumBars + 1) * 8;
    return L;
}

}  // namespace

__global__ void __launch_bounds__(kThreadsTap, 2)
conv_umma_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_u32);
    const SmemLayout L = smem_layout(p.sa, p.sb, p.b_bytes, p.Cout, p.nchunks);
    float* s_par = reinterpret_cast<float*>(smem + L.par);
    const uint32_t bar_full_a = base + L.bars;
    const uint32_t bar_empty_a = bar_full_a + kMaxStagesA * 8;
    const uint32_t bar_full_b = bar_empty_a + kMaxStagesA * 8;
    const uint32_t bar_empty_b = bar_full_b + kMaxStagesB * 8;
    const uint32_t bar_accum = bar_empty_b + kMaxStagesB * 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bars + kNumBars * 8);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int SA = p.sa, SB = p.sb;
    const int C = p.Cout;
    const int n = blockIdx.z;
    const int ty0 = blockIdx.y * kTileH;
    const int tx0 = blockIdx.x * kTileW;

    // ------------------------------------------------------------ setup
    if (warp == 8) {
        if (lane == 0) {
            for (int s = 0; s < SA; ++s) {
                mbar_init(bar_full_a + 8 * s, kWorkers);
                mbar_init(bar_empty_a + 8 * s, 1);
            }
            for (int s = 0; s < SB; ++s) {
                mbar_init(bar_full_b + 8 * s, 1);
                mbar_init(bar_empty_b + 8 * s, 1);
            }
            mbar_init(bar_accum, 1);
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
        // ======================================================== A producers
        const int j = tid & 7;          // slot (16 B column) of the row
        const int prow = tid >> 3;      // 0..31; item i covers accumulator row prow + 32*i
        const int swz = prow & 7;       // (row & 7) of all of this thread's rows
        const size_t frame_in = static_cast<size_t>(n) * p.Hin * p.Win;

        {
            // ---- plain taps: cp.async (LDGSTS) with zero fill, completion lagging by LAG chunks
            const uint32_t row_off = static_cast<uint32_t>(prow) * 128u + (static_cast<uint32_t>(j ^ swz) << 4);
            const int ox = tx0 + (prow & 15);
            const int oy0 = ty0 + (prow >> 4);
            const int LAG = (SA >= 3) ? 2 : 1;
            Slot sl_next = p.slots[j];
            for (int q = 0; q < p.nchunks; ++q) {
                const int s = q % SA;
                const Slot sl = sl_next;
                if (q + 1 < p.nchunks) sl_next = p.slots[(q + 1) * 8 + j];
                if (q >= SA) mbar_wait(bar_empty_a + 8 * s, ((q / SA) - 1) & 1);
                const Seg sg = p.seg[sl.seg];
                const __half* sbase = sg.ptr + sl.choff;
                const uint32_t dst0 = base + L.a0 + s * kABytes + row_off;
                const int vx = ox * p.stride + sl.dx;
                const bool okx = sl.valid && static_cast<unsigned>(vx) < static_cast<unsigned>(p.Wv);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int vy = (oy0 + 2 * i) * p.stride + sl.dy;
                    const bool ok = okx && static_cast<unsigned>(vy) < static_cast<unsigned>(p.Hv);
                    const size_t pix = frame_in + static_cast<size_t>(vy >> p.up) * p.Win + (vx >> p.up);
                    const __half* src = ok ? sbase + pix * sg.pitch : sbase;
                    cp_async16(dst0 + i * 4096u, src, ok ? 16u : 0u);
                }
                cp_async_commit();
                if (q >= LAG) {
                    if (LAG == 2) cp_async_wait<2>(); else cp_async_wait<1>();
                    fence_proxy_async_smem();
                    mbar_arrive(bar_full_a + 8 * ((q - LAG) % SA));
                }
            }
            cp_async_wait<0>();
            fence_proxy_async_smem();
            for (int q = (p.nchunks > LAG ? p.nchunks - LAG : 0); q < p.nchunks; ++q)
                mbar_arrive(bar_full_a + 8 * (q % SA));
        }

        // ======================================================== epilogue (warps 0-7)
        mbar_wait(bar_accum, 0);
        tc_fence_after();
        {
            const int quad = warp & 3;
            const int r = quad * 32 + lane;      // accumulator row = TMEM lane = patch pixel
            epilogue_row<kEpiAll>(p, s_par, tmem_base + (static_cast<uint32_t>(quad * 32) << 16), n, ty0 + (r >> 4), tx0 + (r & 15), warp >> 2,
                         1, C);
        }
        tc_fence_before();
    } else if (warp < 8 + kIssuersTap) {
        // ======================================================== MMA issuers
        // Warp-uniform loops (loop counters, launch constants) and one elected lane issues: keeps the
        // descriptors in uniform registers (a divergent single-lane loop makes the compiler wrap every
        // UTCHMMA in an R2UR / ELECT / BRA.U.ANY sequence).  One warp sustains only ~100 cycles per MMA
        // (tools/umma_rate.cu), so RIC splits the 9 taps of every block over 3 issuers (tap % 3) that
        // accumulate into separate TMEM column ranges, summed in the epilogue.
        // The number of RIC issuers (p.ks, 1..3) is chosen by the planner so that every issuer owns a private
        // weight ring of >= 2 stages: a ring shared by several consumers would let a warp that is one ring
        // revolution ahead pass mbarrier.try_wait.parity on the previous phase.
        const int NI = 1;
        const int SBK = SB / NI;
        const int wi = warp - 8;
        if (wi < NI) {
        const uint32_t idesc = umma_idesc_f16(kTileM, C);
        const uint32_t d_addr = tmem_base + static_cast<uint32_t>(wi * C);
        const int tail_from = p.nchunks - 1;
        uint32_t acc = 0;
        int cnt = 0;
        for (int q = 0; q < p.nchunks; ++q) {
            const int s_a = q % SA;
            const int s_b = wi * SBK + cnt % SBK;
            const uint32_t b_par = (cnt / SBK) & 1;
            ++cnt;
            const uint32_t km = q >= tail_from ? p.kmask_last : p.kmask_full;
            const uint32_t km2 = q >= tail_from ? p.kmask2_last : p.kmask2_full;
            mbar_wait(bar_full_b + 8 * s_b, b_par);
            mbar_wait(bar_full_a + 8 * s_a, (q / SA) & 1);
            tc_fence_after();
            const uint32_t a_addr = base + L.a0 + s_a * kABytes;
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
                umma_commit(bar_empty_a + 8 * s_a);      // frees the stages when these MMAs retire
                umma_commit(bar_empty_b + 8 * s_b);
            }
            acc = 1u;
            __syncwarp();
        }
        if (elect_one()) umma_commit(bar_accum);         // this issuer's partial accumulator is complete
        __syncwarp();
        }
        tc_fence_before();
    } else {
        // ======================================================== weight (B operand) loader: chunk q -> stage q % SB
        for (int q = 0; q < p.nchunks; ++q) {
            const int s = q % SB;
            if (q >= SB) mbar_wait(bar_empty_b + 8 * s, ((q / SB) - 1) & 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_full_b + 8 * s, static_cast<uint32_t>(p.b_bytes));
                bulk_g2s(base + L.b0 + s * p.b_bytes, p.wpack + static_cast<size_t>(q) * p.b_bytes,
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

size_t conv_smem_bytes(const ConvParams& p) {
    return smem_layout(p.sa, p.sb, p.b_bytes, p.Cout, p.nchunks).total + 1024;
}

cudaError_t launch_conv(const ConvParams& p, cudaStream_t stream) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    if (p.ric || p.sa < 2 || p.sa > kMaxStagesA || p.sb < 2 || p.sb > kMaxStagesB || conv_smem_bytes(p) > 227 * 1024)
        return cudaErrorInvalidConfiguration;
    dim3 grid((p.Wout + kTileW - 1) / kTileW, (p.Hout + kTileH - 1) / kTileH, p.B);
    conv_umma_kernel<<<grid, kThreadsTap, conv_smem_bytes(p), stream>>>(p);
    return cudaGetLastError();
}

}  // namespace dsu

// Fused convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05, accumulators in TMEM).
//
// One CTA computes a kTileH x kTileW patch (128 output pixels = UMMA M) of one frame for all
// Cout channels (UMMA N).  K = taps x input channels, walked in 64-element chunks:
//   * warps 0-7 (256 threads) PRODUCE the A operand: one smem row (128 B, SWIZZLE_128B, K-major)
//     per output pixel.  Each 16-byte slot of a row is 8 channels of one tap of one concat
//     segment (slot table), so plain 3x3 / 7x7 taps, stride 2, fused nearest-x2 upsampling and
//     channel concatenation are pure address arithmetic (cp.async with zero fill at the border).
//   * RIC mode (stage-1 rotation-invariant deformable conv, models.py:302-351): a thread loads the
//     3x3 neighbourhood of its pixel once (9 x 16 B, zero outside the image) and blends all 8
//     non-centre taps from it - every tap samples on the unit circle around the pixel - writing one
//     A buffer per tap.  Taps are visited in octant-rotated order so the 2x2 corner set of each
//     tap is a compile-time constant (no dynamic register indexing).
//   * warp 9 streams the pre-swizzled weight tile (B operand) with 1-D bulk async copies.
//   * warp 8 issues tcgen05.mma (one thread) and commits stage-release / accumulator-ready
//     mbarriers.
//   * the producer warps then become the epilogue: tcgen05.ld the accumulator row of their pixel,
//     apply folded BN / activation / residual, write fp16 NHWC (hi [+lo] planes), the fp32 residual
//     stream, or the fused conv_12 1x1 + tanh + uint8 composite tail.
// "Exact" mode (split fp16): a K chunk holds 32 channels as [a_hi | a_lo]; it is multiplied by
// [W_hi | W_hi] (4 K-steps) and its hi half again by W_lo (2 K-steps) into the same accumulator:
// a_hi*W_hi + a_lo*W_hi + a_hi*W_lo, fp32-grade products at 3x the tensor work.
#include "conv.cuh"
#include "ptx.cuh"

namespace dsu {

namespace {

constexpr int kNumBars = 2 * kMaxStagesA + 2 * kMaxStagesB + 1;

struct SmemLayout {
    uint32_t a0, b0, par, hdr, bars;   // byte offsets from the 1024-aligned base
    uint32_t total;
};

__host__ __device__ inline SmemLayout smem_layout(int sa, int sb, int b_bytes, int cout, int nchunks) {
    SmemLayout L;
    L.a0 = 0;
    L.b0 = L.a0 + sa * kABytes;
    L.par = L.b0 + sb * b_bytes;
    L.hdr = (L.par + (7 * cout + 4) * 4 + 15u) & ~15u;
    L.bars = (L.hdr + nchunks * 8 + 15u) & ~15u;
    L.total = L.bars + (kNumBars + 1) * 8;
    return L;
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ void unpack8(const uint4& raw, float* f) {
    float2 a = unpack_h2(raw.x), b = unpack_h2(raw.y), c = unpack_h2(raw.z), d = unpack_h2(raw.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
// 8 fp32 -> packed fp16 hi and residual lo = fp16(v - hi)
__device__ __forceinline__ void split8(const float* f, uint4& hi, uint4& lo) {
    hi.x = pack_h2(f[0], f[1]); hi.y = pack_h2(f[2], f[3]); hi.z = pack_h2(f[4], f[5]); hi.w = pack_h2(f[6], f[7]);
    float r[8];
    unpack8(hi, r);
    lo.x = pack_h2(f[0] - r[0], f[1] - r[1]); lo.y = pack_h2(f[2] - r[2], f[3] - r[3]);
    lo.z = pack_h2(f[4] - r[4], f[5] - r[5]); lo.w = pack_h2(f[6] - r[6], f[7] - r[7]);
}

// exact fp32 -> uint8 of custom_transforms.py:7-8: ((clip(x,-1,1)+1)/2*255) truncated, fp32 ops in order
__device__ __forceinline__ uint8_t to_u8(float x) {
    x = fminf(fmaxf(x, -1.0f), 1.0f);
    float t = __fmul_rn(__fmul_rn(__fadd_rn(x, 1.0f), 0.5f), 255.0f);
    return static_cast<uint8_t>(static_cast<int>(t));
}

// Static corner set of rotated tap m (sample angle in [m*45, m*45+45) degrees):
// row offset cos<0 for m in 2..5, column offset sin<0 for m in 4..7.
__device__ __forceinline__ constexpr int quad_r0(int m) { return (m >= 2 && m <= 5) ? 0 : 1; }
__device__ __forceinline__ constexpr int quad_c0(int m) { return (m >= 4) ? 0 : 1; }

}  // namespace

template <bool kRic>
__global__ void __launch_bounds__(kThreads, kRic ? 1 : 2)
conv_umma_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_u32);
    const SmemLayout L = smem_layout(p.sa, p.sb, p.b_bytes, p.Cout, p.nchunks);
    float* s_par = reinterpret_cast<float*>(smem + L.par);
    ChunkHdr* s_hdr = reinterpret_cast<ChunkHdr*>(smem + L.hdr);
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
        // epilogue parameters -> smem: [scale C][shift C][scale2 C][shift2 C][w12 3C][b12 4]
        for (int i = tid; i < C; i += kWorkers) {
            s_par[i] = p.epi.scale[i];
            s_par[C + i] = p.epi.shift[i];
            s_par[2 * C + i] = p.epi.scale2 ? p.epi.scale2[i] : 1.0f;
            s_par[3 * C + i] = p.epi.shift2 ? p.epi.shift2[i] : 0.0f;
            if (p.epi.w12) {
                s_par[4 * C + i] = p.epi.w12[i];
                s_par[5 * C + i] = p.epi.w12[C + i];
                s_par[6 * C + i] = p.epi.w12[2 * C + i];
            }
        }
        if (p.epi.w12 && tid < 3) s_par[7 * C + tid] = p.epi.b12[tid];
        for (int i = tid; i < p.nchunks; i += kWorkers) s_hdr[i] = p.hdrs[i];
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

        if constexpr (!kRic) {
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
        } else {
            // ---- RIC: one (pixel, 8-channel group) item = 9 neighbour loads -> 8 blended taps + centre.
            // chunk q = block * 9 + tap lives in A buffer `tap` (SA == 9).
            const bool exact = p.exact != 0;
            const int cg = exact ? (j & 3) : j;                 // data slot handled by this thread
            const int i_lo = exact ? 2 * (j >> 2) : 0;          // exact: the two threads of a slot pair split the rows
            const int i_hi = exact ? i_lo + 2 : 4;
            const uint32_t slot_hi = static_cast<uint32_t>(cg ^ swz) << 4;
            const uint32_t slot_lo = static_cast<uint32_t>((cg + 4) ^ swz) << 4;
            for (int b = 0; b < p.nblocks; ++b) {
                const Slot sl = p.slots[b * 8 + cg];
                const Seg sg = p.seg[sl.seg];
                const __half* sbase = sg.ptr + sl.choff;
                const __half* sbase_lo = exact ? p.seg[sl.seg + kMaxSeg / 2].ptr + sl.choff : nullptr;
                for (int i = i_lo; i < i_hi; ++i) {
                    const int r = prow + 32 * i;
                    const int oy = ty0 + (r >> 4), ox = tx0 + (r & 15);
                    const bool live = sl.valid && oy < p.Hout && ox < p.Wout;
                    // ---- stencil of this pixel
                    float2 lyx[8];
                    int oct = 0;
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
                    } else {
#pragma unroll
                        for (int t = 0; t < 8; ++t) lyx[t] = make_float2(0.0f, 0.0f);
                    }
                    // ---- 3x3 neighbourhood (virtual coordinates; nearest-x2 folded into the address)
                    uint4 nb[9], nbl[9];
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        const int vy = oy + k / 3 - 1, vx = ox + k % 3 - 1;
                        const bool inb = live && static_cast<unsigned>(vy) < static_cast<unsigned>(p.Hv) &&
                                         static_cast<unsigned>(vx) < static_cast<unsigned>(p.Wv);
                        nb[k] = make_uint4(0, 0, 0, 0);
                        nbl[k] = make_uint4(0, 0, 0, 0);
                        if (inb) {
                            const size_t pix = frame_in + static_cast<size_t>(vy >> p.up) * p.Win + (vx >> p.up);
                            nb[k] = __ldg(reinterpret_cast<const uint4*>(sbase + pix * sg.pitch));
                            if (exact) nbl[k] = __ldg(reinterpret_cast<const uint4*>(sbase_lo + pix * sg.pitch));
                        }
                    }
                    // first item of a block: the previous block's MMAs must have drained the tap buffers
                    if (b > 0 && i == i_lo) {
#pragma unroll 1
                        for (int t = 0; t < 9; ++t) mbar_wait(bar_empty_a + 8 * t, (b - 1) & 1);
                    }
                    uint8_t* rowp = smem + L.a0 + r * 128;
                    // ---- centre tap (raster tap 4): the pixel itself
                    *reinterpret_cast<uint4*>(rowp + 4 * kABytes + slot_hi) = nb[4];
                    if (exact) *reinterpret_cast<uint4*>(rowp + 4 * kABytes + slot_lo) = nbl[4];
                    // ---- 8 circle taps, two channel halves to bound register use
                    uint32_t keep[8][2];       // half-0 results (packed) while half 1 is computed
                    float keepf[8][4];         // exact mode keeps fp32 to split hi/lo over all 8 channels
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        float nf[9][4];
#pragma unroll
                        for (int k = 0; k < 9; ++k) {
                            const float2 a = unpack_h2(hf ? nb[k].z : nb[k].x), c = unpack_h2(hf ? nb[k].w : nb[k].y);
                            nf[k][0] = a.x; nf[k][1] = a.y; nf[k][2] = c.x; nf[k][3] = c.y;
                            if (exact) {
                                const float2 al = unpack_h2(hf ? nbl[k].z : nbl[k].x), cl = unpack_h2(hf ? nbl[k].w : nbl[k].y);
                                nf[k][0] += al.x; nf[k][1] += al.y; nf[k][2] += cl.x; nf[k][3] += cl.y;
                            }
                        }
#pragma unroll
                        for (int m = 0; m < 8; ++m) {
                            const float ly = lyx[m].x, lx = lyx[m].y;
                            const float hy = 1.0f - ly, hx = 1.0f - lx;
                            const float w00 = hy * hx, w01 = hy * lx, w10 = ly * hx, w11 = ly * lx;
                            const int r0 = quad_r0(m), c0 = quad_c0(m);
                            float o[4];
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                o[c] = fmaf(w11, nf[(r0 + 1) * 3 + c0 + 1][c],
                                       fmaf(w10, nf[(r0 + 1) * 3 + c0][c],
                                       fmaf(w01, nf[r0 * 3 + c0 + 1][c], w00 * nf[r0 * 3 + c0][c])));
                            if (hf == 0) {
                                if (exact) {
#pragma unroll
                                    for (int c = 0; c < 4; ++c) keepf[m][c] = o[c];
                                } else {
                                    keep[m][0] = pack_h2(o[0], o[1]);
                                    keep[m][1] = pack_h2(o[2], o[3]);
                                }
                            } else {
                                const int kq = (m - oct) & 7;               // reference rotation index of this tap
                                const int tap = kq + (kq >= 4 ? 1 : 0);     // raster tap -> A buffer
                                uint8_t* dst = rowp + tap * kABytes;
                                if (exact) {
                                    const float f8[8] = {keepf[m][0], keepf[m][1], keepf[m][2], keepf[m][3], o[0], o[1], o[2], o[3]};
                                    uint4 hi, lo;
                                    split8(f8, hi, lo);
                                    *reinterpret_cast<uint4*>(dst + slot_hi) = hi;
                                    *reinterpret_cast<uint4*>(dst + slot_lo) = lo;
                                } else {
                                    uint4 v;
                                    v.x = keep[m][0]; v.y = keep[m][1]; v.z = pack_h2(o[0], o[1]); v.w = pack_h2(o[2], o[3]);
                                    *reinterpret_cast<uint4*>(dst + slot_hi) = v;
                                }
                            }
                        }
                    }
                }
                fence_proxy_async_smem();
#pragma unroll 1
                for (int t = 0; t < 9; ++t) mbar_arrive(bar_full_a + 8 * t);
            }
        }

        // ======================================================== epilogue (warps 0-7)
        mbar_wait(bar_accum, 0);
        tc_fence_after();
        const EpiParams& e = p.epi;
        const int quad = warp & 3, chalf = warp >> 2;
        const int r = quad * 32 + lane;          // accumulator row = TMEM lane = patch pixel
        const int oy = ty0 + (r >> 4);
        const int oxe = tx0 + (r & 15);
        const bool pix_ok = oy < p.Hout && oxe < p.Wout;
        const size_t opix = (static_cast<size_t>(n) * p.Hout + oy) * p.Wout + oxe;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        const int ncb = C / 32;
        // column batches are split between the two warps of a lane quadrant; the 1x1 tail needs a
        // whole row in one thread, so only the first warp of each quadrant runs it
        const bool tail = e.w12 != nullptr;
        const int cb_first = tail ? 0 : chalf, cb_step = tail ? 1 : 2;
        const bool active = tail ? (chalf == 0) : (chalf < ncb);
        float y3[3] = {0.0f, 0.0f, 0.0f};
        if (active) {
            for (int cbi = cb_first; cbi < ncb; cbi += cb_step) {
                const int cb = cbi * 32;
                uint32_t v[32];
                tmem_ld32(t_row + cb, v);
                tmem_ld_wait();
                if (!pix_ok) continue;
                float f[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    float x = fmaf(__uint_as_float(v[c]), s_par[cb + c], s_par[C + cb + c]);
                    if (e.act == 1) x = fmaxf(x, 0.0f);
                    else if (e.act == 2) x = x > 0.0f ? x : 0.2f * x;
                    if (e.scale2) x = fmaf(x, s_par[2 * C + cb + c], s_par[3 * C + cb + c]);
                    f[c] = x;
                }
                if (e.resid_in) {
                    const float4* rp = reinterpret_cast<const float4*>(e.resid + opix * C + cb);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 rv = rp[c];
                        f[4 * c] += rv.x; f[4 * c + 1] += rv.y; f[4 * c + 2] += rv.z; f[4 * c + 3] += rv.w;
                    }
                }
                if (e.resid_out) {
                    float4* rp = reinterpret_cast<float4*>(e.resid + opix * C + cb);
#pragma unroll
                    for (int c = 0; c < 8; ++c) rp[c] = make_float4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
                }
                if (e.out2_hi) {
                    uint4* o = reinterpret_cast<uint4*>(e.out2_hi + opix * e.out2_pitch + e.out2_choff + cb);
                    uint4* ol = e.out2_lo ? reinterpret_cast<uint4*>(e.out2_lo + opix * e.out2_pitch + e.out2_choff + cb) : nullptr;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 hi, lo;
                        split8(f + 8 * c, hi, lo);
                        o[c] = hi;
                        if (ol) ol[c] = lo;
                    }
                }
                if (e.out_relu) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) f[c] = fmaxf(f[c], 0.0f);
                }
                if (e.out_hi) {
                    uint4* o = reinterpret_cast<uint4*>(e.out_hi + opix * e.out_pitch + e.out_choff + cb);
                    uint4* ol = e.out_lo ? reinterpret_cast<uint4*>(e.out_lo + opix * e.out_pitch + e.out_choff + cb) : nullptr;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 hi, lo;
                        split8(f + 8 * c, hi, lo);
                        o[c] = hi;
                        if (ol) ol[c] = lo;
                    }
                }
                if (tail) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        y3[0] = fmaf(f[c], s_par[4 * C + cb + c], y3[0]);
                        y3[1] = fmaf(f[c], s_par[5 * C + cb + c], y3[1]);
                        y3[2] = fmaf(f[c], s_par[6 * C + cb + c], y3[2]);
                    }
                }
            }
            if (tail && pix_ok) {
                const size_t plane = static_cast<size_t>(p.Hout) * p.Wout;
                const size_t pin = static_cast<size_t>(oy) * p.Wout + oxe;
                uint8_t rgb[3];
#pragma unroll
                for (int o = 0; o < 3; ++o) {
                    float yv = y3[o] + s_par[7 * C + o];
                    if (e.tanh_flag) yv = tanhf(yv);
                    if (e.y_nchw) e.y_nchw[(static_cast<size_t>(n) * 3 + o) * plane + pin] = yv;
                    rgb[o] = to_u8(yv);
                }
                if (e.y_rgba) {
                    const uint8_t a = e.alpha_src ? e.alpha_src[opix * e.alpha_stride] : 255;
                    reinterpret_cast<uchar4*>(e.y_rgba)[opix] = make_uchar4(rgb[0], rgb[1], rgb[2], a);
                }
            }
        }
        tc_fence_before();
    } else if (warp == 8) {
        // ======================================================== MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(kTileM, C);
            uint32_t acc = 0;
            for (int q = 0; q < p.nchunks; ++q) {
                const int s_a = q % SA, s_b = q % SB;
                const ChunkHdr h = s_hdr[q];
                mbar_wait(bar_full_b + 8 * s_b, (q / SB) & 1);
                mbar_wait(bar_full_a + 8 * s_a, (q / SA) & 1);
                tc_fence_after();
                const uint32_t a_addr = base + L.a0 + s_a * kABytes;
                const uint32_t b_addr = base + L.b0 + s_b * p.b_bytes;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if ((h.kmask >> k) & 1) {
                        umma_f16(tmem_base, umma_desc_sw128(a_addr + k * 32, 1024), umma_desc_sw128(b_addr + k * 32, 1024), idesc, acc);
                        acc = 1;
                    }
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    if ((h.kmask2 >> k) & 1)
                        umma_f16(tmem_base, umma_desc_sw128(a_addr + k * 32, 1024),
                                 umma_desc_sw128(b_addr + C * 128 + k * 32, 1024), idesc, 1u);
                umma_commit(bar_empty_a + 8 * s_a);      // frees the stages when these MMAs retire
                umma_commit(bar_empty_b + 8 * s_b);
            }
            umma_commit(bar_accum);                      // accumulator complete -> epilogue
        }
        __syncwarp();
        tc_fence_before();
    } else {
        // ======================================================== weight (B operand) loader
        if (lane == 0) {
            for (int q = 0; q < p.nchunks; ++q) {
                const int s = q % SB;
                if (q >= SB) mbar_wait(bar_empty_b + 8 * s, ((q / SB) - 1) & 1);
                const ChunkHdr h = s_hdr[q];
                const uint32_t bytes = static_cast<uint32_t>(h.kmask2 ? 2 * C : C) * 128u;
                mbar_arrive_expect_tx(bar_full_b + 8 * s, bytes);
                bulk_g2s(base + L.b0 + s * p.b_bytes, p.wpack + h.b_off, bytes, bar_full_b + 8 * s);
            }
        }
        __syncwarp();
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
        cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(conv_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    if (p.sa < 2 || p.sa > kMaxStagesA || p.sb < 2 || p.sb > kMaxStagesB || (p.ric && p.sa != 9) ||
        conv_smem_bytes(p) > 227 * 1024)
        return cudaErrorInvalidConfiguration;
    dim3 grid((p.Wout + kTileW - 1) / kTileW, (p.Hout + kTileH - 1) / kTileH, p.B);
    if (p.ric) conv_umma_kernel<true><<<grid, kThreads, conv_smem_bytes(p), stream>>>(p);
    else conv_umma_kernel<false><<<grid, kThreads, conv_smem_bytes(p), stream>>>(p);
    return cudaGetLastError();
}

}  // namespace dsu

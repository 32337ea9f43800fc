// Fused convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05, accumulators in TMEM).
//
// One CTA computes a kTileH x kTileW patch (128 output pixels = UMMA M) of one frame for all
// Cout channels (UMMA N).  K = taps x input channels, walked in 64-element chunks:
//   * warps 0-3 (128 threads) PRODUCE the A operand: one smem row (128 B, SWIZZLE_128B, K-major)
//     per output pixel.  Each 16-byte slot of a row is 8 channels of one tap of one concat
//     segment (slot table), so plain 3x3 / 7x7 taps, stride 2, fused nearest-x2 upsampling and
//     channel concatenation are pure address arithmetic (cp.async with zero fill at the border).
//     In RIC mode (stage-1 rotation-invariant deformable conv, models.py:302-351) a slot is the
//     bilinear blend of 4 neighbours with weights from a per-level stencil table.
//   * warp 5 streams the pre-swizzled weight tile (B operand) with 1-D bulk async copies.
//   * warp 4 issues tcgen05.mma (one thread) and commits stage-release / accumulator-ready
//     mbarriers.
//   * the producer warps then become the epilogue: tcgen05.ld the accumulator row of their pixel,
//     apply folded BN / activation / residual, write fp16 NHWC (hi [+lo] planes), the fp32 residual
//     stream, or the fused conv_12 1x1 + tanh + uint8 composite tail.
// "Exact" mode (split fp16): activations and weights are hi+lo fp16 pairs; K chunks alternate
// A_hi x [W_hi;W_lo] (N = 2*Cout, two accumulator halves) and A_lo x W_hi.
#include "conv.cuh"
#include "ptx.cuh"

namespace dsu {

namespace {

struct SmemLayout {
    uint32_t a0, b0, par, hdr, bars;   // byte offsets from the 1024-aligned base
    uint32_t total;
};

__host__ __device__ inline SmemLayout smem_layout(int nstages, int a_bytes, int b_bytes, int cout, int nchunks) {
    SmemLayout L;
    L.a0 = 0;
    L.b0 = L.a0 + nstages * a_bytes;
    L.par = L.b0 + nstages * b_bytes;
    L.hdr = (L.par + (7 * cout + 4) * 4 + 15u) & ~15u;
    L.bars = (L.hdr + nchunks * 8 + 15u) & ~15u;
    L.total = L.bars + (2 * kMaxStages + 2) * 8;
    return L;
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void unpack8(const uint4& raw, float* f) {
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 v = __half22float2(h[i]);
        f[2 * i] = v.x; f[2 * i + 1] = v.y;
    }
}

// exact fp32 -> uint8 of custom_transforms.py:7-8: ((clip(x,-1,1)+1)/2*255) truncated, fp32 ops in order
__device__ __forceinline__ uint8_t to_u8(float x) {
    x = fminf(fmaxf(x, -1.0f), 1.0f);
    float t = __fmul_rn(__fmul_rn(__fadd_rn(x, 1.0f), 0.5f), 255.0f);
    return static_cast<uint8_t>(static_cast<int>(t));
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 2)
conv_umma_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_u32);
    const SmemLayout L = smem_layout(p.nstages, p.a_bytes, p.b_bytes, p.Cout, p.nchunks);
    float* s_par = reinterpret_cast<float*>(smem + L.par);
    ChunkHdr* s_hdr = reinterpret_cast<ChunkHdr*>(smem + L.hdr);
    const uint32_t bar_full = base + L.bars;
    const uint32_t bar_empty = bar_full + kMaxStages * 8;
    const uint32_t bar_accum = bar_empty + kMaxStages * 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bars + (2 * kMaxStages + 1) * 8);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int S = p.nstages;
    const int C = p.Cout;
    const int n = blockIdx.z;
    const int ty0 = blockIdx.y * kTileH;
    const int tx0 = blockIdx.x * kTileW;

    // ------------------------------------------------------------ setup
    if (warp == 4) {
        if (lane == 0) {
            for (int s = 0; s < S; ++s) {
                mbar_init(bar_full + 8 * s, kWorkers + 1);
                mbar_init(bar_empty + 8 * s, 1);
            }
            mbar_init(bar_accum, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
        tmem_relinquish();
    } else if (warp < 4) {
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

    if (warp < 4) {
        // ======================================================== A producers
        const int j = tid & 7;          // slot (16 B column) of the row
        const int prow = tid >> 3;      // patch column 0..15; item i covers patch row i
        const uint32_t row_off = static_cast<uint32_t>(prow) * 128u + (static_cast<uint32_t>(j ^ (prow & 7)) << 4);
        const int ox = tx0 + prow;
        const size_t frame_in = static_cast<size_t>(n) * p.Hin * p.Win;

        if (!p.ric) {
            // ---- plain taps: cp.async (LDGSTS) with zero fill, completion lagging by LAG chunks
            const int LAG = (S >= 3) ? 2 : 1;
            Slot sl_next = p.slots[j];
            for (int q = 0; q < p.nchunks; ++q) {
                const int s = q % S;
                const Slot sl = sl_next;
                if (q + 1 < p.nchunks) sl_next = p.slots[(q + 1) * 8 + j];
                if (q >= S) mbar_wait(bar_empty + 8 * s, ((q / S) - 1) & 1);
                const Seg sg = p.seg[sl.seg];
                const __half* sbase = sg.ptr + sl.choff;
                const uint32_t dst0 = base + L.a0 + s * p.a_bytes + row_off;
#pragma unroll
                for (int i = 0; i < kTileH; ++i) {
                    const int vy = (ty0 + i) * p.stride + sl.dy;
                    const int vx = ox * p.stride + sl.dx;
                    const bool ok = sl.valid && static_cast<unsigned>(vy) < static_cast<unsigned>(p.Hv) &&
                                    static_cast<unsigned>(vx) < static_cast<unsigned>(p.Wv);
                    const size_t pix = frame_in + static_cast<size_t>(vy >> p.up) * p.Win + (vx >> p.up);
                    const __half* src = ok ? sbase + pix * sg.pitch : sbase;
                    cp_async16(dst0 + i * 2048u, src, ok ? 16u : 0u);
                }
                cp_async_commit();
                if (q >= LAG) {
                    if (LAG == 2) cp_async_wait<2>(); else cp_async_wait<1>();
                    fence_proxy_async_smem();
                    mbar_arrive(bar_full + 8 * ((q - LAG) % S));
                }
            }
            cp_async_wait<0>();
            fence_proxy_async_smem();
            for (int q = (p.nchunks > LAG ? p.nchunks - LAG : 0); q < p.nchunks; ++q)
                mbar_arrive(bar_full + 8 * (q % S));
        } else {
            // ---- RIC taps: bilinear blend of 4 neighbours (fp32 math), register path.
            // exact mode: chunks come in (hi, lo) pairs built from one set of loads.
            const int step = p.exact ? 2 : 1;
            const int HW = p.Hout * p.Wout;
            for (int q = 0; q < p.nchunks; q += step) {
                const int s = q % S;
                if (q >= S) mbar_wait(bar_empty + 8 * s, ((q / S) - 1) & 1);
                const int s2 = (q + 1) % S;
                if (p.exact && q + 1 >= S) mbar_wait(bar_empty + 8 * s2, (((q + 1) / S) - 1) & 1);
                const Slot sl = p.slots[q * 8 + j];
                const Seg sg = p.seg[sl.seg];
                const __half* sbase = sg.ptr + sl.choff;
                const __half* sbase_lo = p.exact ? p.seg[sl.seg + kMaxSeg / 2].ptr + sl.choff : nullptr;
                uint8_t* dst = smem + L.a0 + s * p.a_bytes + row_off;
                uint8_t* dst_lo = smem + L.a0 + s2 * p.a_bytes + row_off;
                const int tap = sl.dy;
#pragma unroll 2
                for (int i = 0; i < kTileH; ++i) {
                    const int oy = ty0 + i;
                    float acc[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[c] = 0.0f;
                    if (sl.valid && oy < p.Hout && ox < p.Wout) {
                        float w[4] = {1.0f, 0.0f, 0.0f, 0.0f};
                        int dyl = 0, dyh = 0, dxl = 0, dxh = 0;
                        if (tap != 4) {
                            const int e = (tap < 4 ? tap : tap - 1) * HW + oy * p.Wout + ox;
                            const float4 wv = __ldg(p.ric_w + e);
                            const char4 ov = __ldg(p.ric_off + e);
                            w[0] = wv.x; w[1] = wv.y; w[2] = wv.z; w[3] = wv.w;
                            dyl = ov.x; dyh = ov.y; dxl = ov.z; dxh = ov.w;
                        }
                        const int ncorner = (tap == 4) ? 1 : 4;
                        uint4 raw[4], raw_lo[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (c < ncorner) {
                                const int vy = oy + ((c & 2) ? dyh : dyl);
                                const int vx = ox + ((c & 1) ? dxh : dxl);
                                const size_t pix = frame_in + static_cast<size_t>(vy >> p.up) * p.Win + (vx >> p.up);
                                raw[c] = __ldg(reinterpret_cast<const uint4*>(sbase + pix * sg.pitch));
                                if (p.exact) raw_lo[c] = __ldg(reinterpret_cast<const uint4*>(sbase_lo + pix * sg.pitch));
                            }
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (c < ncorner) {
                                float f[8];
                                unpack8(raw[c], f);
#pragma unroll
                                for (int k = 0; k < 8; ++k) acc[k] = fmaf(w[c], f[k], acc[k]);
                                if (p.exact) {
                                    unpack8(raw_lo[c], f);
#pragma unroll
                                    for (int k = 0; k < 8; ++k) acc[k] = fmaf(w[c], f[k], acc[k]);
                                }
                            }
                        }
                    }
                    uint4 hi;
                    hi.x = pack_h2(acc[0], acc[1]); hi.y = pack_h2(acc[2], acc[3]);
                    hi.z = pack_h2(acc[4], acc[5]); hi.w = pack_h2(acc[6], acc[7]);
                    *reinterpret_cast<uint4*>(dst + i * 2048) = hi;
                    if (p.exact) {
                        float r[8];
                        unpack8(hi, r);
                        uint4 lo;
                        lo.x = pack_h2(acc[0] - r[0], acc[1] - r[1]); lo.y = pack_h2(acc[2] - r[2], acc[3] - r[3]);
                        lo.z = pack_h2(acc[4] - r[4], acc[5] - r[5]); lo.w = pack_h2(acc[6] - r[6], acc[7] - r[7]);
                        *reinterpret_cast<uint4*>(dst_lo + i * 2048) = lo;
                    }
                }
                fence_proxy_async_smem();
                mbar_arrive(bar_full + 8 * s);
                if (p.exact) mbar_arrive(bar_full + 8 * s2);
            }
        }

        // ======================================================== epilogue (same 4 warps)
        mbar_wait(bar_accum, 0);
        tc_fence_after();
        const int r = tid;                       // accumulator row = TMEM lane = patch pixel
        const int oy = ty0 + (r >> 4);
        const int oxe = tx0 + (r & 15);
        const bool pix_ok = oy < p.Hout && oxe < p.Wout;
        const size_t opix = (static_cast<size_t>(n) * p.Hout + oy) * p.Wout + oxe;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        const EpiParams& e = p.epi;
        float y3[3] = {0.0f, 0.0f, 0.0f};
        for (int cb = 0; cb < C; cb += 32) {
            uint32_t v[32];
            tmem_ld32(t_row + cb, v);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) f[c] = __uint_as_float(v[c]);
            if (p.exact) {                       // second accumulator half: A_hi x W_lo
                tmem_ld32(t_row + C + cb, v);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c) f[c] += __uint_as_float(v[c]);
            }
            if (pix_ok) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    float x = fmaf(f[c], s_par[cb + c], s_par[C + cb + c]);
                    if (e.act == 1) x = fmaxf(x, 0.0f);
                    else if (e.act == 2) x = x > 0.0f ? x : 0.2f * x;
                    if (e.scale2) x = fmaf(x, s_par[2 * C + cb + c], s_par[3 * C + cb + c]);
                    f[c] = x;
                }
                if (e.resid_in) {
                    const float4* rp = reinterpret_cast<const float4*>(e.resid + opix * C + cb);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float4 rv = rp[c];
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
                        uint4 hi;
                        hi.x = pack_h2(f[8 * c], f[8 * c + 1]); hi.y = pack_h2(f[8 * c + 2], f[8 * c + 3]);
                        hi.z = pack_h2(f[8 * c + 4], f[8 * c + 5]); hi.w = pack_h2(f[8 * c + 6], f[8 * c + 7]);
                        o[c] = hi;
                        if (ol) {
                            float rr[8];
                            unpack8(hi, rr);
                            uint4 lo;
                            lo.x = pack_h2(f[8 * c] - rr[0], f[8 * c + 1] - rr[1]); lo.y = pack_h2(f[8 * c + 2] - rr[2], f[8 * c + 3] - rr[3]);
                            lo.z = pack_h2(f[8 * c + 4] - rr[4], f[8 * c + 5] - rr[5]); lo.w = pack_h2(f[8 * c + 6] - rr[6], f[8 * c + 7] - rr[7]);
                            ol[c] = lo;
                        }
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
                        uint4 hi;
                        hi.x = pack_h2(f[8 * c], f[8 * c + 1]); hi.y = pack_h2(f[8 * c + 2], f[8 * c + 3]);
                        hi.z = pack_h2(f[8 * c + 4], f[8 * c + 5]); hi.w = pack_h2(f[8 * c + 6], f[8 * c + 7]);
                        o[c] = hi;
                        if (ol) {
                            float rr[8];
                            unpack8(hi, rr);
                            uint4 lo;
                            lo.x = pack_h2(f[8 * c] - rr[0], f[8 * c + 1] - rr[1]); lo.y = pack_h2(f[8 * c + 2] - rr[2], f[8 * c + 3] - rr[3]);
                            lo.z = pack_h2(f[8 * c + 4] - rr[4], f[8 * c + 5] - rr[5]); lo.w = pack_h2(f[8 * c + 6] - rr[6], f[8 * c + 7] - rr[7]);
                            ol[c] = lo;
                        }
                    }
                }
                if (e.w12) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        y3[0] = fmaf(f[c], s_par[4 * C + cb + c], y3[0]);
                        y3[1] = fmaf(f[c], s_par[5 * C + cb + c], y3[1]);
                        y3[2] = fmaf(f[c], s_par[6 * C + cb + c], y3[2]);
                    }
                }
            }
        }
        if (e.w12 && pix_ok) {
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
                uchar4 px = make_uchar4(rgb[0], rgb[1], rgb[2], a);
                reinterpret_cast<uchar4*>(e.y_rgba)[opix] = px;
            }
        }
        tc_fence_before();
    } else if (warp == 4) {
        // ======================================================== MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc_n = umma_idesc_f16(kTileM, C);
            const uint32_t idesc_w = umma_idesc_f16(kTileM, 2 * C);
            for (int q = 0; q < p.nchunks; ++q) {
                const int s = q % S;
                const ChunkHdr h = s_hdr[q];
                mbar_wait(bar_full + 8 * s, (q / S) & 1);
                tc_fence_after();
                const uint32_t a_addr = base + L.a0 + s * p.a_bytes;
                const uint32_t b_addr = base + L.b0 + s * p.b_bytes;
                const uint32_t idesc = h.wide ? idesc_w : idesc_n;
                for (int k = 0; k < h.ksteps; ++k) {
                    const uint64_t da = umma_desc_sw128(a_addr + k * 32, 1024);
                    const uint64_t db = umma_desc_sw128(b_addr + k * 32, 1024);
                    umma_f16(tmem_base, da, db, idesc, (q > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(bar_empty + 8 * s);      // frees the stage when these MMAs retire
            }
            umma_commit(bar_accum);                  // accumulator complete -> epilogue
        }
        __syncwarp();
        tc_fence_before();
    } else {
        // ======================================================== weight (B operand) loader
        if (lane == 0) {
            for (int q = 0; q < p.nchunks; ++q) {
                const int s = q % S;
                if (q >= S) mbar_wait(bar_empty + 8 * s, ((q / S) - 1) & 1);
                const ChunkHdr h = s_hdr[q];
                const uint32_t bytes = static_cast<uint32_t>(h.wide ? 2 * C : C) * 128u;
                mbar_arrive_expect_tx(bar_full + 8 * s, bytes);
                bulk_g2s(base + L.b0 + s * p.b_bytes, p.wpack + h.b_off, bytes, bar_full + 8 * s);
            }
        }
        __syncwarp();
    }

    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

size_t conv_smem_bytes(const ConvParams& p) {
    return smem_layout(p.nstages, p.a_bytes, p.b_bytes, p.Cout, p.nchunks).total + 1024;
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
    dim3 grid((p.Wout + kTileW - 1) / kTileW, (p.Hout + kTileH - 1) / kTileH, p.B);
    conv_umma_kernel<<<grid, kThreads, conv_smem_bytes(p), stream>>>(p);
    return cudaGetLastError();
}

}  // namespace dsu

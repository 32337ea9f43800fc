// A-operand producer of the RIC (rotation-invariant deformable) convolution, stage 1
// (training/models.py:302-351 with the offset field of generate_coordinates, :551-604).
//
// Every non-centre tap k samples at pixel + (cos, sin)(theta + k*pi/4): inside the 3x3 neighbourhood.
// One work item = (output pixel, 8-channel group): load the 3x3 neighbourhood once (9 x 16 B, zero
// outside the image, which reproduces torchvision's border rule), blend the 8 circle taps in fp32
// from registers and write one 16-byte slot into each of the 9 per-tap A buffers.  Taps are
// visited in octant-rotated order m = (octant + k) & 7, for which the 2x2 corner set is a
// compile-time constant: rows {-1,0} for m in 2..5 else {0,+1}; cols {-1,0} for m >= 4 else {0,+1}.
#pragma once
#include "conv_device.cuh"

namespace dsu {

__device__ __forceinline__ constexpr int ric_r0(int m) { return (m >= 2 && m <= 5) ? 0 : 1; }
__device__ __forceinline__ constexpr int ric_c0(int m) { return (m >= 4) ? 0 : 1; }

// kExact = false: 256 threads = 32 pixel rows x 8 slots, 4 items per thread, fp16 result.
// kExact = true : slots are [hi x4 | lo x4]; thread (prow, j) handles channel group j & 3 for two of the
//                 four pixel rows, reads hi + lo planes and writes fp16 hi and lo = fp16(v - hi).
// kHalfBlend (fp16 mode only): blend with packed half2 math instead of fp32.
template <bool kExact, bool kHalfBlend>
// g0 = number of blocks this CTA has already produced (persistent kernel: the tap-buffer barriers keep counting across tiles).
__device__ __forceinline__ void ric_produce(const ConvParams& p, uint8_t* a_smem, uint32_t bar_full_a, uint32_t bar_empty_a,
                                            int tid, int n, int ty0, int tx0, int g0 = 0) {
    const int j = tid & 7, prow = tid >> 3, swz = prow & 7;
    const int cg = kExact ? (j & 3) : j;
    const int i_lo = kExact ? 2 * (j >> 2) : 0;
    const int i_hi = kExact ? i_lo + 2 : 4;
    const uint32_t slot_hi = static_cast<uint32_t>(cg ^ swz) << 4;
    const uint32_t slot_lo = static_cast<uint32_t>((cg + 4) ^ swz) << 4;
    const size_t frame_in = static_cast<size_t>(n) * p.Hin * p.Win;
    // column geometry is the same for every item of this thread
    const int ox = tx0 + (prow & 15);
    int cx[3];
    bool okx[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int vx = ox + c - 1;
        okx[c] = static_cast<unsigned>(vx) < static_cast<unsigned>(p.Wv);
        cx[c] = vx >> p.up;
    }
    for (int b = 0; b < p.nblocks; ++b) {
        const Slot sl = p.slots[b * 8 + cg];
        const Seg sg = p.seg[sl.seg];
        const int pitch = sg.pitch;
        const __half* fb = sg.ptr + sl.choff + frame_in * pitch;
        const __half* fb_lo = kExact ? p.seg[sl.seg + kMaxSeg / 2].ptr + sl.choff + frame_in * pitch : nullptr;
        // ---- item loaders / consumers (forced inline; the item loop is fully unrolled so that everything
        // stays in registers).  In fp16 mode the loads of item i+1 are issued before item i is blended
        // (software pipelining: global latency ~1k cycles per item with only 2 producer warps per scheduler).
        auto load_item = [&](int i, uint4 (&nb)[9], uint4 (&nbl)[9], float2 (&lyx)[8], int& oct) {
            const int r = prow + 32 * i;
            const int oy = ty0 + (r >> 4);
            const bool live = sl.valid && oy < p.Hout && ox < p.Wout;
            oct = 0;
            if (live) {
                const size_t e = static_cast<size_t>(oy) * p.Wout + ox;
                const float4* tp = kHalfBlend ? reinterpret_cast<const float4*>(p.ric_wh + e * 8)      // fp16 weights, same 64 B / pixel
                                              : reinterpret_cast<const float4*>(p.ric_lyx + e * 8);
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
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
                const int vy = oy + rr - 1;
                const bool oky = live && static_cast<unsigned>(vy) < static_cast<unsigned>(p.Hv);
                const int rowoff = (vy >> p.up) * p.Win;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    nb[rr * 3 + c] = make_uint4(0, 0, 0, 0);
                    if (kExact) nbl[rr * 3 + c] = make_uint4(0, 0, 0, 0);
                    if (oky && okx[c]) {
                        const int off = (rowoff + cx[c]) * pitch;
                        nb[rr * 3 + c] = __ldg(reinterpret_cast<const uint4*>(fb + off));
                        if (kExact) nbl[rr * 3 + c] = __ldg(reinterpret_cast<const uint4*>(fb_lo + off));
                    }
                }
            }
        };
        auto blend_item = [&](int i, const uint4 (&nb)[9], const uint4 (&nbl)[9], const float2 (&lyx)[8], int oct) {
            const int r = prow + 32 * i;
            // first item of a block: the previous block's MMAs must have drained the tap buffers
            if (g0 + b > 0 && i == i_lo) {
                // every lane polls (warp-uniform): a single polling lane followed by __syncwarp parks the other 31 lanes on a
                // WARPSYNC while lane 0 sleeps in NANOSLEEP.SYNCS - measured 5 % slower on upconv1 (profiles/r01m_ric_upconv1.ncu-rep)
#pragma unroll 1
                for (int t = 0; t < 9; ++t) mbar_wait(bar_empty_a + 8 * t, (g0 + b - 1) & 1);
            }
            uint8_t* rowp = a_smem + r * 128;
            // ---- centre tap (raster tap 4): the pixel itself
            *reinterpret_cast<uint4*>(rowp + 4 * kABytes + slot_hi) = nb[4];
            if (kExact) *reinterpret_cast<uint4*>(rowp + 4 * kABytes + slot_lo) = nbl[4];
            if constexpr (!kExact && kHalfBlend) {
                // ---- fp16 mode, packed math: the neighbours stay half2, the 4 bilinear weights of a tap are
                // rounded to fp16 and the blend is 1 HMUL2 + 3 HFMA2 per channel pair (no unpack / repack).
                // Costs ~1.5 fp16 ulp more rounding on the A operand than the fp32 blend below.
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    // table entry = {w00,w01 | w10,w11} in fp16 (host: fp32 products of (1-ly),(1-lx),ly,lx rounded once)
                    const __half2 wa = *reinterpret_cast<const __half2*>(&lyx[m].x), wb = *reinterpret_cast<const __half2*>(&lyx[m].y);
                    const __half2 w00 = __low2half2(wa), w01 = __high2half2(wa), w10 = __low2half2(wb), w11 = __high2half2(wb);
                    const int r0 = ric_r0(m), c0 = ric_c0(m);
                    const __half2* n00 = reinterpret_cast<const __half2*>(&nb[r0 * 3 + c0]);
                    const __half2* n01 = reinterpret_cast<const __half2*>(&nb[r0 * 3 + c0 + 1]);
                    const __half2* n10 = reinterpret_cast<const __half2*>(&nb[(r0 + 1) * 3 + c0]);
                    const __half2* n11 = reinterpret_cast<const __half2*>(&nb[(r0 + 1) * 3 + c0 + 1]);
                    uint4 v;
                    __half2* o = reinterpret_cast<__half2*>(&v);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        o[c] = __hfma2(w11, n11[c], __hfma2(w10, n10[c], __hfma2(w01, n01[c], __hmul2(w00, n00[c]))));
                    const int kq = (m - oct) & 7;
                    const int tap = kq + (kq >> 2);
                    *reinterpret_cast<uint4*>(rowp + tap * kABytes + slot_hi) = v;
                }
            } else if constexpr (!kExact) {
                // ---- all 8 channels at once: 72 fp32 neighbours in registers, weights computed once per tap
                float nf[9][8];
#pragma unroll
                for (int k = 0; k < 9; ++k) unpack8(nb[k], nf[k]);
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const float ly = lyx[m].x, lx = lyx[m].y;
                    const float hy = 1.0f - ly, hx = 1.0f - lx;
                    const float w00 = hy * hx, w01 = hy * lx, w10 = ly * hx, w11 = ly * lx;
                    const int r0 = ric_r0(m), c0 = ric_c0(m);
                    float o[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        o[c] = fmaf(w11, nf[(r0 + 1) * 3 + c0 + 1][c],
                               fmaf(w10, nf[(r0 + 1) * 3 + c0][c],
                               fmaf(w01, nf[r0 * 3 + c0 + 1][c], w00 * nf[r0 * 3 + c0][c])));
                    const int kq = (m - oct) & 7;               // reference rotation index of this tap
                    const int tap = kq + (kq >> 2);             // raster tap (centre skipped) -> A buffer
                    uint4 v;
                    v.x = pack_h2(o[0], o[1]); v.y = pack_h2(o[2], o[3]); v.z = pack_h2(o[4], o[5]); v.w = pack_h2(o[6], o[7]);
                    *reinterpret_cast<uint4*>(rowp + tap * kABytes + slot_hi) = v;
                }
            } else {
                // ---- exact: hi + lo planes, two channel halves to bound register use
                float keepf[8][4];
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    float nf[9][4];
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        const float2 a = unpack_h2(hf ? nb[k].z : nb[k].x), c = unpack_h2(hf ? nb[k].w : nb[k].y);
                        const float2 al = unpack_h2(hf ? nbl[k].z : nbl[k].x), cl = unpack_h2(hf ? nbl[k].w : nbl[k].y);
                        nf[k][0] = a.x + al.x; nf[k][1] = a.y + al.y; nf[k][2] = c.x + cl.x; nf[k][3] = c.y + cl.y;
                    }
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        const float ly = lyx[m].x, lx = lyx[m].y;
                        const float hy = 1.0f - ly, hx = 1.0f - lx;
                        const float w00 = hy * hx, w01 = hy * lx, w10 = ly * hx, w11 = ly * lx;
                        const int r0 = ric_r0(m), c0 = ric_c0(m);
                        float o[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            o[c] = fmaf(w11, nf[(r0 + 1) * 3 + c0 + 1][c],
                                   fmaf(w10, nf[(r0 + 1) * 3 + c0][c],
                                   fmaf(w01, nf[r0 * 3 + c0 + 1][c], w00 * nf[r0 * 3 + c0][c])));
                        if (hf == 0) {
#pragma unroll
                            for (int c = 0; c < 4; ++c) keepf[m][c] = o[c];
                        } else {
                            const int kq = (m - oct) & 7;
                            const int tap = kq + (kq >> 2);
                            const float f8[8] = {keepf[m][0], keepf[m][1], keepf[m][2], keepf[m][3], o[0], o[1], o[2], o[3]};
                            uint4 hi, lo;
                            split8(f8, hi, lo);
                            *reinterpret_cast<uint4*>(rowp + tap * kABytes + slot_hi) = hi;
                            *reinterpret_cast<uint4*>(rowp + tap * kABytes + slot_lo) = lo;
                        }
                    }
                }
            }
        };
        if constexpr (kExact) {
            for (int i = i_lo; i < i_hi; ++i) {
                uint4 nb[9], nbl[9];
                float2 lyx[8];
                int oct;
                load_item(i, nb, nbl, lyx, oct);
                blend_item(i, nb, nbl, lyx, oct);
            }
        } else {
            uint4 nbA[9], nbB[9], nbl[9];
            float2 lyA[8], lyB[8];
            int octA, octB;
            load_item(0, nbA, nbl, lyA, octA);
            load_item(1, nbB, nbl, lyB, octB);
            blend_item(0, nbA, nbl, lyA, octA);
            load_item(2, nbA, nbl, lyA, octA);
            blend_item(1, nbB, nbl, lyB, octB);
            load_item(3, nbB, nbl, lyB, octB);
            blend_item(2, nbA, nbl, lyA, octA);
            blend_item(3, nbB, nbl, lyB, octB);
        }
        // every lane fences its own smem writes, then one arrival per warp and tap (barrier count = 8 producer warps)
        fence_proxy_async_smem();
        __syncwarp();
        if ((tid & 31) == 0) {
#pragma unroll 1
            for (int t = 0; t < 9; ++t) mbar_arrive(bar_full_a + 8 * t);
        }
    }
}

}  // namespace dsu

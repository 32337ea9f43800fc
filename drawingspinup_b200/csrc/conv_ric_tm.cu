// RIC (rotation-invariant deformable) convolution of stage 1 (training/models.py:302-351 with the offset field of
// generate_coordinates, :551-604) with the blended A operand in TENSOR MEMORY.
//
// Why (round-2 measurements, profiles/r02b_umma_ts_probe.log, the r01n_ric_persist_upconv1 row of profiles/r01_kernels.csv): the round-1 kernel wrote
// the 9 blended taps of every (pixel, 8 channels) item to shared memory and let tcgen05.mma read them back - 144 KB of
// single-buffered A tiles per 64-channel block (production and MMAs serialised), three shared-memory passes over the same
// bytes (neighbour reads, A writes, A reads: the SM's 128 B/clk were the limit before the tensor pipe), and N = 64 layers
// capped at 2/3 of the MMA rate by operand bandwidth.  Here
//   * the source pixels of a tile are staged ONCE per 128-byte channel block by a TMA tensor load (4-D NHWC box with
//     out-of-bounds zero fill = torchvision's border rule; the fused nearest x2 of the up-convolutions is a half-size box);
//   * a producer thread owns one pixel (= one TMEM lane), keeps its stencil weights / neighbour addresses in registers for
//     the whole tile, blends from ld.shared and writes the taps with tcgen05.st straight into tensor memory;
//   * tcgen05.mma reads A from TMEM (TS form): shared memory only carries the weight tiles, and N = 64 issues at the full rate;
//   * several warps issue MMAs into ONE accumulator (exact, probe part 3), so there are no K-split partial sums; production
//     of stage j+1 overlaps the MMAs of stage j (two 144-column A stages in TMEM).
// The tap rotation of round 1 (rotated tap m = octant + k has a compile-time 2x2 corner set) moves from the DESTINATION
// address (impossible with the warp-uniform column of tcgen05.st) to the CODE: eight blend variants selected by the pixel's
// octant, writing raster-ordered taps; a warp whose pixels straddle an octant boundary runs two variants.
//
// Roles (20 warps): 0-7 producers (two per TMEM lane quadrant: chunk pairs g = 0 / 1 of a stage), 8-11 epilogue,
// 12-17 MMA issuers, 18 halo loader (TMA tensor loads), 19 weight loader (bulk copies).
#include "conv_device.cuh"

namespace dsu {

namespace {

constexpr int kTmMaxB = 8;                       // weight stages in shared memory
constexpr int kTmBars = 2 + 2 + 3 + kTmMaxB + kTmMaxB + 2 + 2;

struct TmSmem {
    uint32_t halo0, b0, par, bars, total;
};
__host__ __device__ inline TmSmem tm_smem(int sb, int b_stage_bytes, int cout) {
    TmSmem L;
    L.halo0 = 0;
    L.b0 = 2 * kTmHaloBytes;
    L.par = L.b0 + sb * b_stage_bytes;
    L.bars = (L.par + (7 * cout + 4) * 4 + 15u) & ~15u;
    L.total = L.bars + (kTmBars + 1) * 8;
    return L;
}

// Every barrier wait of this kernel carries a watchdog: a wait that lasts longer than ~1 s (a protocol bug, never a
// legitimate state) records (tag, a, b, block) in a pinned host buffer and traps, so a deadlock becomes a diagnosable
// launch failure instead of a hung GPU.  dbg may be null (no record, still traps).
__device__ __noinline__ void tm_watchdog_fire(unsigned long long* dbg, uint32_t tag, int a, int b) {
    if (dbg && (threadIdx.x & 31) == 0) {
        dbg[threadIdx.x >> 5] = 0xD5ull << 56 | static_cast<unsigned long long>(tag & 0xFF) << 48 |
                                static_cast<unsigned long long>(a & 0xFFFF) << 32 | static_cast<unsigned long long>(b & 0xFFFF) << 16 |
                                static_cast<unsigned long long>(blockIdx.x & 0xFFFF);
        __threadfence_system();
    }
}
// kSleepNs = 0: tight poll (the waits on the producer <-> MMA critical path).  kSleepNs > 0: test, then sleep between polls -
// the waits of the roles that idle for most of a tile (epilogue, loaders, issuers).  The first version polled everything
// with the blocking try_wait: NANOSLEEP.SYNCS wakes on every barrier event of the CTA, the idle roles re-tested ~170 times
// per tile and three quarters of all issued instructions were wait loops competing with the producers for issue slots
// (profiles/r02d_tm_upconv1_fp16 source page).
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
template <int kSleepNs>
__device__ __forceinline__ void mbar_wait_poll(uint32_t bar, uint32_t parity, unsigned long long* dbg, uint32_t tag, int a, int b) {
    if (mbar_test(bar, parity)) return;
    long long t0 = 0;
    uint32_t iter = 0;
    bool reported = false;
    while (true) {
        if constexpr (kSleepNs > 0) {
            asm volatile("nanosleep.u32 %0;" ::"n"(kSleepNs));
            if (mbar_test(bar, parity)) return;
        } else {
            if (mbar_try_wait(bar, parity)) return;
        }
        if ((++iter & 0x3FFu) == 0) {                 // the clock is only read every 1024 polls
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            const long long dt = now - t0;
            if (!reported && dt > 2000000000LL) { tm_watchdog_fire(dbg, tag, a, b); reported = true; }
            if (dt > 2800000000LL) __trap();
        }
    }
}

// Every wait of this kernel is made by all 32 lanes of a warp; lanes may leave the polling loop in different iterations, and
// what follows a wait is often a warp-collective .sync.aligned tcgen05 instruction that must not be reached diverged.
template <int kSleepNs>
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity, unsigned long long* dbg, uint32_t tag, int a, int b) {
    mbar_wait_poll<kSleepNs>(bar, parity, dbg, tag, a, b);
    __syncwarp();
}

template <int kRegs>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }

// Development trace (knob tm_trace): CTA 0 records (event, clock64) pairs of one warp per role into pinned host memory,
// 1024 slots per role: [0] producer warp 0, [1] issuer 0, [2] epilogue warp 0, [3] weight loader, [4] halo loader.
struct Trace {
    unsigned long long* p = nullptr;
    int n = 0;
    __device__ __forceinline__ void init(unsigned long long* base, int role, bool on) { p = on && base ? base + role * 1024 : nullptr; }
    __device__ __forceinline__ void ev(int id, int a) {
        if (p && n < 1024) { p[n++] = static_cast<unsigned long long>(id) << 56 | static_cast<unsigned long long>(a & 0xFFFF) << 40 |
                                      (static_cast<unsigned long long>(clock64()) & 0xFFFFFFFFFFull); }
    }
};

// position in a ring of n barriers / buffers and the parity of the current round (no divisions in the per-stage loops)
struct Ring {
    int idx = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance(int n) {
        if (++idx == n) { idx = 0; phase ^= 1u; }
    }
};

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ constexpr int tm_r0(int m) { return (m >= 2 && m <= 5) ? 0 : 1; }
__device__ __forceinline__ constexpr int tm_c0(int m) { return (m >= 4) ? 0 : 1; }

// ---- fp16 mode: 8 channels of one pixel, packed half2 blend (same operation order as round 1's ric_produce<false, true>:
// w00*n00, then fma w01*n01, w10*n10, w11*n11), raster-ordered output.  wq = the pixel's 8 x {w00,w01 | w10,w11} fp16 table
// entry in rotated tap order; octant O maps raster tap t (circle index kq) to rotated tap m = (kq + O) & 7.
template <int O>
__device__ __forceinline__ void blend_h(const uint4 (&nb)[9], const uint4 (&wq)[4], uint4 (&out)[9]) {
    out[4] = nb[4];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        if (t == 4) continue;
        const int kq = t < 4 ? t : t - 1;
        const int m = (kq + O) & 7;
        const uint32_t a = (m & 1) ? wq[m >> 1].z : wq[m >> 1].x, b = (m & 1) ? wq[m >> 1].w : wq[m >> 1].y;
        const __half2 wa = *reinterpret_cast<const __half2*>(&a), wb = *reinterpret_cast<const __half2*>(&b);
        const __half2 w00 = __low2half2(wa), w01 = __high2half2(wa), w10 = __low2half2(wb), w11 = __high2half2(wb);
        const int r0 = tm_r0(m), c0 = tm_c0(m);
        const __half2* n00 = reinterpret_cast<const __half2*>(&nb[r0 * 3 + c0]);
        const __half2* n01 = reinterpret_cast<const __half2*>(&nb[r0 * 3 + c0 + 1]);
        const __half2* n10 = reinterpret_cast<const __half2*>(&nb[(r0 + 1) * 3 + c0]);
        const __half2* n11 = reinterpret_cast<const __half2*>(&nb[(r0 + 1) * 3 + c0 + 1]);
        __half2* o = reinterpret_cast<__half2*>(&out[t]);
#pragma unroll
        for (int c = 0; c < 4; ++c)
            o[c] = __hfma2(w11, n11[c], __hfma2(w10, n10[c], __hfma2(w01, n01[c], __hmul2(w00, n00[c]))));
    }
}

// ---- split-fp16 mode: 4 fp32 channels of one pixel, fp32 blend with the reference's weights (1-ly)(1-lx), (1-ly)lx,
// ly(1-lx), ly*lx, then split into fp16 hi and lo = fp16(v - hi); out[t] = {hi01, hi23, lo01, lo23}
__device__ __forceinline__ float sub_half(float v, uint32_t packed, int hi_half) {
    // v - float(half): one mixed-precision FMA (FHFMA) instead of a convert and a subtract; exact in fp32
    float d;
    if (hi_half)
        asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tfma.rn.f32.f16 %0, hi, %2, %3;\n\t}" : "=f"(d) : "r"(packed), "h"(static_cast<unsigned short>(0xBC00)), "f"(v));
    else
        asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tfma.rn.f32.f16 %0, lo, %2, %3;\n\t}" : "=f"(d) : "r"(packed), "h"(static_cast<unsigned short>(0xBC00)), "f"(v));
    return d;
}
__device__ __forceinline__ uint4 split4(float v0, float v1, float v2, float v3) {
    uint4 r;
    r.x = pack_h2(v0, v1);
    r.y = pack_h2(v2, v3);
    r.z = pack_h2(sub_half(v0, r.x, 0), sub_half(v1, r.x, 1));
    r.w = pack_h2(sub_half(v2, r.y, 0), sub_half(v3, r.y, 1));
    return r;
}
// packed fp32 pairs (FFMA2 / FMUL2, sm_100): two channels per instruction, the scalar weight is broadcast by the
// instruction's operand selector (R.F32), so the fp32 blend costs 64 instead of 128 issue slots per 4-channel chunk
__device__ __forceinline__ unsigned long long f2pack(uint32_t lo, uint32_t hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long f2bcast(float w) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(w));
    return r;
}
__device__ __forceinline__ unsigned long long f2mul(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long f2fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
template <int O>
__device__ __forceinline__ void blend_f(const uint4 (&nb)[9], const float (&w)[8][4], uint4 (&out)[9]) {
    out[4] = split4(__uint_as_float(nb[4].x), __uint_as_float(nb[4].y), __uint_as_float(nb[4].z), __uint_as_float(nb[4].w));
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        if (t == 4) continue;
        const int kq = t < 4 ? t : t - 1;
        const int m = (kq + O) & 7;
        const int r0 = tm_r0(m), c0 = tm_c0(m);
        const uint4 &a = nb[r0 * 3 + c0], &b = nb[r0 * 3 + c0 + 1], &c = nb[(r0 + 1) * 3 + c0], &d = nb[(r0 + 1) * 3 + c0 + 1];
        // same operation order per channel as the scalar form: w00*a, then fma w01*b, w10*c, w11*d (each one rounding)
        const unsigned long long w0 = f2bcast(w[m][0]), w1 = f2bcast(w[m][1]), w2 = f2bcast(w[m][2]), w3 = f2bcast(w[m][3]);
        const unsigned long long v01 = f2fma(w3, f2pack(d.x, d.y), f2fma(w2, f2pack(c.x, c.y), f2fma(w1, f2pack(b.x, b.y), f2mul(w0, f2pack(a.x, a.y)))));
        const unsigned long long v23 = f2fma(w3, f2pack(d.z, d.w), f2fma(w2, f2pack(c.z, c.w), f2fma(w1, f2pack(b.z, b.w), f2mul(w0, f2pack(a.z, a.w)))));
        float v0, v1, v2, v3;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(v0), "=f"(v1) : "l"(v01));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(v2), "=f"(v3) : "l"(v23));
        out[t] = split4(v0, v1, v2, v3);
    }
}

}  // namespace

template <bool kExact>
__global__ void __launch_bounds__(kTmThreads, 1)
conv_ric_tm_kernel(const __grid_constant__ TmParams P) {
    const ConvParams& p = P.c;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_u32);
    const int C = p.Cout, SA = P.sa, SB = P.sb, NI = P.ni, NSETS = P.nsets;
    const TmSmem L = tm_smem(SB, P.b_stage_bytes, C);
    float* s_par = reinterpret_cast<float*>(smem + L.par);
    const uint32_t bar_halo_full = base + L.bars;                    // [2]
    const uint32_t bar_halo_empty = bar_halo_full + 2 * 8;           // [2]
    const uint32_t bar_a_full = bar_halo_empty + 2 * 8;              // [3]
    const uint32_t bar_b_full = bar_a_full + 3 * 8;                  // [kTmMaxB]
    const uint32_t bar_done = bar_b_full + kTmMaxB * 8;              // [kTmMaxB] MMAs of stage j complete: bar_done[j % SB]
    const uint32_t bar_acc_full = bar_done + kTmMaxB * 8;            // [2]
    const uint32_t bar_acc_empty = bar_acc_full + 2 * 8;             // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bars + kTmBars * 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_x = (p.Wout + kTileW - 1) / kTileW, tiles_y = (p.Hout + kTileH - 1) / kTileH;
    const int tiles_per_frame = tiles_x * tiles_y;
    const int total_tiles = tiles_per_frame * p.B;
    const int my_tiles = (total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    constexpr int kEpi = kTmEpilogueWarps * 32;
    constexpr int kWarpEpi0 = kTmProducerWarps, kWarpIss0 = kWarpEpi0 + kTmEpilogueWarps, kWarpHalo = kWarpIss0 + kTmIssuerWarps,
                  kWarpWgt = kWarpHalo + 1;

    if (warp == kWarpIss0) {
        if (lane == 0) {
            for (int s = 0; s < 2; ++s) {
                mbar_init(bar_halo_full + 8 * s, 1);
                mbar_init(bar_halo_empty + 8 * s, kTmProducerWarps);
                mbar_init(bar_acc_full + 8 * s, NI);
                mbar_init(bar_acc_empty + 8 * s, kEpi);
            }
            for (int s = 0; s < 3; ++s) mbar_init(bar_a_full + 8 * s, kTmProducerWarps);
            for (int s = 0; s < kTmMaxB; ++s) {
                mbar_init(bar_b_full + 8 * s, 1);
                mbar_init(bar_done + 8 * s, NI);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), 512);
        tmem_relinquish();
    } else if (warp >= kWarpEpi0 && warp < kWarpIss0) {
        load_epilogue_params(p, s_par, tid - kWarpEpi0 * 32, kEpi);
    } else if (warp == kWarpHalo && lane == 0) {
        for (int i = 0; i < kTmMaxMaps; ++i) tma_prefetch_desc(&P.tmap[i]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_col0 = 512u - static_cast<uint32_t>(SA * kTmStageCols);   // A stages at the top of the 512 columns

    auto tile_coords = [&](int it, int& n, int& ty0, int& tx0) {
        const int t = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
        n = t / tiles_per_frame;
        const int r = t - n * tiles_per_frame;
        ty0 = (r / tiles_x) * kTileH;
        tx0 = (r % tiles_x) * kTileW;
    };

    if (warp < kTmProducerWarps) {
        // ================================================================ producers
        reg_inc<144>();      // CTA register pool: 8 light warps release (96 - 40) x 32 each = 14336 = 8 x (144 - 96) x 32 + 4 x (112 - 96) x 32
        const int g = warp >> 2;                                     // chunk pair of a stage this warp produces
        const int r = (warp & 3) * 32 + lane;                        // tile pixel = accumulator row = TMEM lane
        const int py = r >> 4, px = r & 15;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        const uint32_t halo_u32 = base + L.halo0;
        int j = 0, gb = 0;                                           // stages / blocks produced so far by this CTA
        Ring ra, rd;                                                 // A stage being written; completion barrier of step j - SA
        Trace tr;
        tr.init(P.trace, 0, blockIdx.x == 0 && tid == 0);
        for (int it = 0; it < my_tiles; ++it) {
            int n, ty0, tx0;
            tile_coords(it, n, ty0, tx0);
            const int oy = ty0 + py, ox = tx0 + px;
            const bool live = oy < p.Hout && ox < p.Wout;
            const size_t e = live ? static_cast<size_t>(oy) * p.Wout + ox : 0;
            // ---- per-pixel state for the whole tile
            const int oct = __ldg(p.ric_oct + e);
            uint4 wq[4];                                             // fp16 mode: 8 x {w00,w01 | w10,w11}
            float w[8][4];                                           // split-fp16 mode: fp32 weights in rotated tap order
            if constexpr (!kExact) {
                const uint4* tp = reinterpret_cast<const uint4*>(p.ric_wh + e * 8);
#pragma unroll
                for (int i = 0; i < 4; ++i) wq[i] = __ldg(tp + i);
            } else {
                const float4* tp = reinterpret_cast<const float4*>(p.ric_lyx + e * 8);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 v = __ldg(tp + i);
                    const float ly0 = v.x, lx0 = v.y, ly1 = v.z, lx1 = v.w;
                    const float hy0 = 1.0f - ly0, hx0 = 1.0f - lx0, hy1 = 1.0f - ly1, hx1 = 1.0f - lx1;
                    w[2 * i][0] = hy0 * hx0; w[2 * i][1] = hy0 * lx0; w[2 * i][2] = ly0 * hx0; w[2 * i][3] = ly0 * lx0;
                    w[2 * i + 1][0] = hy1 * hx1; w[2 * i + 1][1] = hy1 * lx1; w[2 * i + 1][2] = ly1 * hx1; w[2 * i + 1][3] = ly1 * lx1;
                }
            }
            // neighbour (dy, dx) -> 128-byte line of the halo tile (source pixel, nearest x2 folded in) with its swizzle key in
            // bits 4-6: the 16-byte chunk c of that pixel sits at line * 128 + ((c ^ (line & 7)) << 4)  (TMA SWIZZLE_128B)
            uint32_t nbl[9];
            {
                const int sy0 = (ty0 - 1) >> p.up, sx0 = (tx0 - 1) >> p.up;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const int line = (((oy + dy - 1) >> p.up) - sy0) * P.halo_w + (((ox + dx - 1) >> p.up) - sx0);
                        nbl[dy * 3 + dx] = static_cast<uint32_t>(line) * 128u + (static_cast<uint32_t>(line & 7) << 4);
                    }
            }
            for (int b = 0; b < P.nblocks; ++b, ++gb) {
                const TmBlock blk = P.blk[b];
                const int hb = gb & 1;
                mbar_wait_wd<0>(bar_halo_full + 8 * hb, (gb >> 1) & 1, P.dbg, 1, gb, it);
                const uint32_t hbase = halo_u32 + static_cast<uint32_t>(hb) * kTmHaloBytes;
                for (int h = 0; h < blk.nstages; ++h, ++j) {
                    const int slot = ra.idx;
                    if (j >= SA) {                                   // the MMAs that read this A stage last must be complete
                        mbar_wait_wd<0>(bar_done + 8 * rd.idx, rd.phase, P.dbg, 2, j, it);
                        rd.advance(SB);
                        tc_fence_after();
                    }
                    ra.advance(SA);
                    tr.ev(1, j);
                    const uint32_t col0 = a_col0 + static_cast<uint32_t>(slot * kTmStageCols);
                    const int nvalid = blk.chunks[h];
#pragma unroll 1
                    for (int u = 0; u < 2; ++u) {
                        const int ci = 2 * g + u;                    // chunk of the stage; chunk of the line = 4 * h + ci
                        if (ci >= nvalid) {                          // K padding of a ragged last stage: the MMA reads these columns
                            if constexpr (!kExact) {
                                const uint32_t cb = lane_addr + col0 + static_cast<uint32_t>((ci >> 1) * 72 + (ci & 1) * 4);
#pragma unroll
                                for (int t = 0; t < 9; ++t) tmem_st4(cb + t * 8, 0u, 0u, 0u, 0u);
                            } else {
                                const uint32_t cb = lane_addr + col0 + static_cast<uint32_t>(ci * 2);
#pragma unroll
                                for (int t = 0; t < 9; ++t) { tmem_st2(cb + t * 8, 0u, 0u); tmem_st2(cb + 72 + t * 8, 0u, 0u); }
                            }
                            continue;
                        }
                        uint4 out[9];
                        {
                            const uint32_t cx = static_cast<uint32_t>(4 * h + ci) << 4;
                            uint4 nb[9];
#pragma unroll
                            for (int k = 0; k < 9; ++k) nb[k] = lds128(hbase + (nbl[k] ^ cx));
                            if constexpr (!kExact) {
                                switch (oct) {
                                    case 0: blend_h<0>(nb, wq, out); break;
                                    case 1: blend_h<1>(nb, wq, out); break;
                                    case 2: blend_h<2>(nb, wq, out); break;
                                    case 3: blend_h<3>(nb, wq, out); break;
                                    case 4: blend_h<4>(nb, wq, out); break;
                                    case 5: blend_h<5>(nb, wq, out); break;
                                    case 6: blend_h<6>(nb, wq, out); break;
                                    default: blend_h<7>(nb, wq, out); break;
                                }
                            } else {
                                switch (oct) {
                                    case 0: blend_f<0>(nb, w, out); break;
                                    case 1: blend_f<1>(nb, w, out); break;
                                    case 2: blend_f<2>(nb, w, out); break;
                                    case 3: blend_f<3>(nb, w, out); break;
                                    case 4: blend_f<4>(nb, w, out); break;
                                    case 5: blend_f<5>(nb, w, out); break;
                                    case 6: blend_f<6>(nb, w, out); break;
                                    default: blend_f<7>(nb, w, out); break;
                                }
                            }
                        }
                        __syncwarp();                                // reconverge before the warp-collective stores
                        if constexpr (!kExact) {
                            // 8 channels = 4 columns of K16 step ci >> 1: tap t at col0 + (ci >> 1) * 72 + t * 8 + (ci & 1) * 4
                            const uint32_t cb = lane_addr + col0 + static_cast<uint32_t>((ci >> 1) * 72 + (ci & 1) * 4);
#pragma unroll
                            for (int t = 0; t < 9; ++t) tmem_st4(cb + t * 8, out[t].x, out[t].y, out[t].z, out[t].w);
                        } else {
                            // 4 channels = 2 columns: hi at col0 + t * 8 + ci * 2, lo 72 columns further
                            const uint32_t cb = lane_addr + col0 + static_cast<uint32_t>(ci * 2);
#pragma unroll
                            for (int t = 0; t < 9; ++t) {
                                tmem_st2(cb + t * 8, out[t].x, out[t].y);
                                tmem_st2(cb + 72 + t * 8, out[t].z, out[t].w);
                            }
                        }
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_a_full + 8 * slot);
                    tr.ev(2, j);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_halo_empty + 8 * hb);    // this warp has read the halo tile for the last time
            }
        }
    } else if (warp < kWarpIss0) {
        // ================================================================ epilogue (one warp per TMEM lane quadrant)
        reg_inc<112>();
        const int quad = warp - kWarpEpi0;
        const int r = quad * 32 + lane;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        // every MMA accumulates: the accumulators start as zeros and are zeroed again after each drain
        for (int s = 0; s < NSETS; ++s) {
            for (int c = 0; c < C; c += 32) tmem_st_zero32(lane_addr + static_cast<uint32_t>(s * C + c));
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_acc_empty + 8 * s);
        }
        Ring rs;
        Trace tr;
        tr.init(P.trace, 2, blockIdx.x == 0 && tid == kWarpEpi0 * 32);
        for (int it = 0; it < my_tiles; ++it) {
            int n, ty0, tx0;
            tile_coords(it, n, ty0, tx0);
            const int set = rs.idx;
            mbar_wait_wd<256>(bar_acc_full + 8 * set, rs.phase, P.dbg, 3, it, set);
            tr.ev(5, it);
            rs.advance(NSETS);
            tc_fence_after();
            const uint32_t t_acc = lane_addr + static_cast<uint32_t>(set * C);
            // stage 1 only needs the plain variants (no activation / ReLU / LeakyReLU; never a second affine or a lo plane)
            epilogue_row<0x0007u>(p, s_par, t_acc, n, ty0 + (r >> 4), tx0 + (r & 15), 0, 1, 0, 1);
            tr.ev(8, it);
            __syncwarp();            // lanes of pixels outside the image took a shorter path through the epilogue
            for (int c = 0; c < C; c += 32) tmem_st_zero32(t_acc + c);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_acc_empty + 8 * set);
            tr.ev(6, it);
        }
    } else {
        reg_dec<40>();
        if (warp < kWarpHalo) {
            // ============================================================ MMA issuers: MMA q of a stage goes to issuer q % NI
            const int wi = warp - kWarpIss0;
            if (wi < NI) {
                const uint32_t idesc = umma_idesc_f16(kTileM, C);
                const uint32_t tile_b = static_cast<uint32_t>(C) * 32u;          // one (tap, part) weight tile
                int j = 0;
                Ring rs, ra, rb;
                Trace tr;
                tr.init(P.trace, 1, blockIdx.x == 0 && tid == kWarpIss0 * 32);
                for (int it = 0; it < my_tiles; ++it) {
                    const int set = rs.idx;
                    mbar_wait_wd<96>(bar_acc_empty + 8 * set, rs.phase, P.dbg, 4, it, set);        // drained and zeroed
                    rs.advance(NSETS);
                    tc_fence_after();
                    const uint32_t d_addr = tmem_base + static_cast<uint32_t>(set * C);
                    for (int b = 0; b < P.nblocks; ++b) {
                        const TmBlock blk = P.blk[b];
                        for (int h = 0; h < blk.nstages; ++h, ++j) {
                            const int slot = ra.idx, sb = rb.idx;
                            // fp16: parts = live K16 steps (2 chunks each); split-fp16: hi*Whi, lo*Whi, hi*Wlo of one K16 step
                            const int nparts = kExact ? 3 : (blk.chunks[h] + 1) >> 1;
                            const int nq = 9 * nparts;
                            mbar_wait_wd<32>(bar_b_full + 8 * sb, rb.phase, P.dbg, 5, j, it);
                            mbar_wait_wd<32>(bar_a_full + 8 * slot, ra.phase, P.dbg, 6, j, it);
                            ra.advance(SA);
                            rb.advance(SB);
                            tc_fence_after();
                            tr.ev(3, j);
                            const uint32_t a_stage = tmem_base + a_col0 + static_cast<uint32_t>(slot * kTmStageCols);
                            const uint32_t b_stage = base + L.b0 + static_cast<uint32_t>(sb) * P.b_stage_bytes;
                            for (int q = wi; q < nq; q += NI) {
                                const int part = q / 9, t = q - part * 9;
                                // A columns: fp16 part = K16 step; split: part 1 reads the lo columns.  B tile: fp16 (t, part); split (t, part == 2)
                                const uint32_t a_addr = a_stage + static_cast<uint32_t>((kExact ? (part == 1 ? 72 : 0) : part * 72) + t * 8);
                                const uint32_t b_addr = b_stage + static_cast<uint32_t>(t * 2 + (kExact ? (part == 2 ? 1 : 0) : part)) * tile_b;
                                if (elect_one()) umma_f16_ts(d_addr, a_addr, umma_desc_noswizzle(b_addr, 128, 256), idesc, 1u);
                                __syncwarp();
                            }
                            if (elect_one()) umma_commit(bar_done + 8 * sb);     // frees the A stage (producers) and the weight stage (loader)
                            __syncwarp();
                            tr.ev(4, j);
                        }
                    }
                    if (elect_one()) umma_commit(bar_acc_full + 8 * set);
                    __syncwarp();
                }
            }
            tc_fence_before();
        } else if (warp == kWarpHalo) {
            // ============================================================ halo loader: one TMA tensor load per (tile, block)
            const uint32_t box_bytes = static_cast<uint32_t>(P.halo_w * P.halo_h) * 128u;
            int gb = 0;
            for (int it = 0; it < my_tiles; ++it) {
                int n, ty0, tx0;
                tile_coords(it, n, ty0, tx0);
                const int sy0 = (ty0 - 1) >> p.up, sx0 = (tx0 - 1) >> p.up;
                for (int b = 0; b < P.nblocks; ++b, ++gb) {
                    const int hb = gb & 1;
                    if (gb >= 2) mbar_wait_wd<160>(bar_halo_empty + 8 * hb, ((gb >> 1) - 1) & 1, P.dbg, 7, gb, it);
                    if (elect_one()) {
                        const TmBlock blk = P.blk[b];
                        mbar_arrive_expect_tx(bar_halo_full + 8 * hb, box_bytes);
                        tma_load_4d(base + L.halo0 + static_cast<uint32_t>(hb) * kTmHaloBytes, &P.tmap[blk.map], blk.c0, sx0, sy0, n,
                                    bar_halo_full + 8 * hb);
                    }
                    __syncwarp();
                }
            }
        } else {
            // ============================================================ weight loader: one bulk copy per stage
            int j = 0;
            Ring rb;
            Trace tr;
            tr.init(P.trace, 3, blockIdx.x == 0 && lane == 0);
            for (int it = 0; it < my_tiles; ++it)
                for (int s = 0; s < P.nstages; ++s, ++j) {
                    const int sb = rb.idx;
                    if (j >= SB) mbar_wait_wd<160>(bar_done + 8 * sb, rb.phase ^ 1u, P.dbg, 8, j, it);
                    rb.advance(SB);
                    tr.ev(7, j);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bar_b_full + 8 * sb, static_cast<uint32_t>(P.b_stage_bytes));
                        bulk_g2s(base + L.b0 + static_cast<uint32_t>(sb) * P.b_stage_bytes, P.wpack + static_cast<size_t>(s) * P.b_stage_bytes,
                                 static_cast<uint32_t>(P.b_stage_bytes), bar_b_full + 8 * sb);
                    }
                    __syncwarp();
                }
        }
    }

    __syncthreads();
    if (warp == kWarpIss0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

size_t conv_ric_tm_smem_bytes(int cout, int sb) { return tm_smem(sb, cout * 576, cout).total + 1024; }

cudaError_t launch_conv_ric_tm(const TmParams& P, cudaStream_t stream) {
    static bool attr_set[64] = {};
    static int sm_count[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64) dev = 0;
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv_ric_tm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_ric_tm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    const ConvParams& p = P.c;
    const size_t smem = conv_ric_tm_smem_bytes(p.Cout, P.sb);
    if (P.sa < 2 || P.sa > 3 || P.sb < P.sa || P.sb > kTmMaxB || (P.sb % P.sa) || P.nsets < 1 || P.nsets > 2 || P.ni < 1 ||
        P.ni > kTmIssuerWarps || P.nsets * p.Cout + P.sa * kTmStageCols > 512 || (p.Cout % 32) || p.Cout < 32 || p.Cout > 224 ||
        P.nblocks < 1 || P.nblocks > kTmMaxBlocks || P.b_stage_bytes != p.Cout * 576 || smem > 227 * 1024 || p.up < 0 || p.up > 1)
        return cudaErrorInvalidConfiguration;
    const int tiles = ((p.Wout + kTileW - 1) / kTileW) * ((p.Hout + kTileH - 1) / kTileH) * p.B;
    const int ctas = tiles < sm_count[dev] ? tiles : sm_count[dev];
    if (p.exact) conv_ric_tm_kernel<true><<<ctas, kTmThreads, smem, stream>>>(P);
    else conv_ric_tm_kernel<false><<<ctas, kTmThreads, smem, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace dsu

// HBM-bound per-pixel kernels around the convolution stack: frame ingest (uint8 / fp32 -> fp16 NHWC
// hi[/lo] planes), 2x2 max-pool, and the stand-alone uint8 steps of the reference's frame loop
// (custom_transforms.py:7-35, data.py:23-47, test_stage1.py:68-70, run_render.py:31-57).
// One thread per pixel (or per 8-channel group), 128-bit accesses where the layout allows.
#include "frames.cuh"

namespace dsu {

namespace {

__device__ __forceinline__ uint32_t pack_h2f(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// write 8 fp32 values as fp16 hi (and optional lo = fp16(v - hi)) 16-byte groups
__device__ __forceinline__ void store8(__half* hi, __half* lo, const float* v) {
    uint4 h;
    h.x = pack_h2f(v[0], v[1]); h.y = pack_h2f(v[2], v[3]); h.z = pack_h2f(v[4], v[5]); h.w = pack_h2f(v[6], v[7]);
    *reinterpret_cast<uint4*>(hi) = h;
    if (lo) {
        const __half2* hh = reinterpret_cast<const __half2*>(&h);
        float r[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 t = __half22float2(hh[i]); r[2 * i] = t.x; r[2 * i + 1] = t.y; }
        uint4 l;
        l.x = pack_h2f(v[0] - r[0], v[1] - r[1]); l.y = pack_h2f(v[2] - r[2], v[3] - r[3]);
        l.z = pack_h2f(v[4] - r[4], v[5] - r[5]); l.w = pack_h2f(v[6] - r[6], v[7] - r[7]);
        *reinterpret_cast<uint4*>(lo) = l;
    }
}

// ToTensor + Normalize(0.5, 0.5) of custom_transforms.py:18-22: (u8/255 - 0.5)/0.5, fp32 ops in order
__device__ __forceinline__ float norm_u8(uint8_t v) {
    return __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), 0.5f), 0.5f);
}

// 8 fp32 values -> fp32 NHWC (stage 1 of the split-fp16 mode keeps fp32 activations)
__device__ __forceinline__ void store8_f32(float* dst, const float* v) {
    reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

__global__ void ingest_f32_kernel(const float* __restrict__ x, int cin, int cpad, size_t npix_frame, size_t npix,
                                  __half* hi, __half* lo, float* f32, int pitch, int choff) {
    const size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (p >= npix) return;
    const size_t n = p / npix_frame, q = p % npix_frame;
    const float* xp = x + n * cin * npix_frame + q;
    for (int g = 0; g < cpad; g += 8) {
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = (g + c < cin) ? xp[static_cast<size_t>(g + c) * npix_frame] : 0.0f;
        if (f32) store8_f32(f32 + p * pitch + choff + g, v);
        else store8(hi + p * pitch + choff + g, lo ? lo + p * pitch + choff + g : nullptr, v);
    }
}

__global__ void frames_to_tensor_kernel(const uchar4* __restrict__ color, const uchar4* __restrict__ pos,
                                        const uint8_t* __restrict__ edge, size_t npix_frame, size_t npix,
                                        float* __restrict__ pre, float* __restrict__ mask_out) {
    const size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (p >= npix) return;
    const size_t n = p / npix_frame, r = p % npix_frame;
    uchar4 c = color[p];
    const uchar4 q = pos[p];
    const float mask = __fdiv_rn(static_cast<float>(c.w), 255.0f);
    if (edge && edge[p] < 255) { c.x = 0; c.y = 0; c.z = 0; }
    float* o = pre + n * 6 * npix_frame + r;
    o[0] = norm_u8(c.x); o[npix_frame] = norm_u8(c.y); o[2 * npix_frame] = norm_u8(c.z);
    o[3 * npix_frame] = mask; o[4 * npix_frame] = norm_u8(q.x); o[5 * npix_frame] = norm_u8(q.y);
    if (mask_out) mask_out[p] = mask;
}

// 2x2 / stride 2 max-pool over NHWC fp16 (8 channels per thread).  With a lo plane the winner is
// chosen on hi+lo and carries its own (hi, lo) pair.
__global__ void maxpool2_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, int in_pitch, int in_choff,
                                int B, int Hin, int Win, int C,
                                __half* out_hi, __half* out_lo, int out_pitch) {
    const int Ho = Hin / 2, Wo = Win / 2, G = C / 8;
    const size_t total = static_cast<size_t>(B) * Ho * Wo * G;
    const size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (t >= total) return;
    const int g = static_cast<int>(t % G);
    const size_t op = t / G;
    const int ox = static_cast<int>(op % Wo);
    const int oy = static_cast<int>((op / Wo) % Ho);
    const int n = static_cast<int>(op / (static_cast<size_t>(Wo) * Ho));
    float best[8], bh[8], bl[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t ip = (static_cast<size_t>(n) * Hin + 2 * oy + (k >> 1)) * Win + 2 * ox + (k & 1);
        const uint4 rh = *reinterpret_cast<const uint4*>(in_hi + ip * in_pitch + in_choff + g * 8);
        uint4 rl = make_uint4(0, 0, 0, 0);
        if (in_lo) rl = *reinterpret_cast<const uint4*>(in_lo + ip * in_pitch + in_choff + g * 8);
        const __half2* ph = reinterpret_cast<const __half2*>(&rh);
        const __half2* pl = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float2 h2 = __half22float2(ph[c]);
            const float2 l2 = __half22float2(pl[c]);
            const float v0 = h2.x + l2.x, v1 = h2.y + l2.y;
            if (k == 0 || v0 > best[2 * c]) { best[2 * c] = v0; bh[2 * c] = h2.x; bl[2 * c] = l2.x; }
            if (k == 0 || v1 > best[2 * c + 1]) { best[2 * c + 1] = v1; bh[2 * c + 1] = h2.y; bl[2 * c + 1] = l2.y; }
        }
    }
    uint4 oh, ol;
    oh.x = pack_h2f(bh[0], bh[1]); oh.y = pack_h2f(bh[2], bh[3]); oh.z = pack_h2f(bh[4], bh[5]); oh.w = pack_h2f(bh[6], bh[7]);
    *reinterpret_cast<uint4*>(out_hi + op * out_pitch + g * 8) = oh;
    if (out_lo) {
        ol.x = pack_h2f(bl[0], bl[1]); ol.y = pack_h2f(bl[2], bl[3]); ol.z = pack_h2f(bl[4], bl[5]); ol.w = pack_h2f(bl[6], bl[7]);
        *reinterpret_cast<uint4*>(out_lo + op * out_pitch + g * 8) = ol;
    }
}

// the same pool over fp32 NHWC (4 channels per thread)
__global__ void maxpool2_f32_kernel(const float* __restrict__ in, int in_pitch, int in_choff, int B, int Hin, int Win, int C,
                                    float* out, int out_pitch) {
    const int Ho = Hin / 2, Wo = Win / 2, G = C / 4;
    const size_t total = static_cast<size_t>(B) * Ho * Wo * G;
    const size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (t >= total) return;
    const int g = static_cast<int>(t % G);
    const size_t op = t / G;
    const int ox = static_cast<int>(op % Wo);
    const int oy = static_cast<int>((op / Wo) % Ho);
    const int n = static_cast<int>(op / (static_cast<size_t>(Wo) * Ho));
    float4 best = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t ip = (static_cast<size_t>(n) * Hin + 2 * oy + (k >> 1)) * Win + 2 * ox + (k & 1);
        const float4 v = *reinterpret_cast<const float4*>(in + ip * in_pitch + in_choff + g * 4);
        if (k == 0) best = v;
        else {      // strict > keeps the first maximum, like the fp16 kernel above and F.max_pool2d's value
            if (v.x > best.x) best.x = v.x;
            if (v.y > best.y) best.y = v.y;
            if (v.z > best.z) best.z = v.z;
            if (v.w > best.w) best.w = v.w;
        }
    }
    *reinterpret_cast<float4*>(out + op * out_pitch + g * 4) = best;
}

__device__ __forceinline__ uint8_t to_u8_dev(float x) {
    x = fminf(fmaxf(x, -1.0f), 1.0f);
    const float t = __fmul_rn(__fmul_rn(__fadd_rn(x, 1.0f), 0.5f), 255.0f);
    return static_cast<uint8_t>(static_cast<int>(t));
}

__global__ void to_image_space_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, size_t n) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i < n) out[i] = to_u8_dev(x[i]);
}

__global__ void overlap_edge_kernel(const uint8_t* __restrict__ edge, uchar4* rgba, size_t npix) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i < npix && edge[i] < 255) rgba[i] = make_uchar4(0, 0, 0, 255);
}

__global__ void compose_rgba_kernel(const float* __restrict__ y, const float* __restrict__ mask,
                                    size_t npix_frame, size_t npix, uchar4* __restrict__ out) {
    const size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (p >= npix) return;
    const size_t n = p / npix_frame, r = p % npix_frame;
    const float* yp = y + n * 3 * npix_frame + r;
    const uint8_t a = static_cast<uint8_t>(static_cast<int>(__fmul_rn(mask[p], 255.0f)));   // (mask*255).astype(uint8)
    out[p] = make_uchar4(to_u8_dev(yp[0]), to_u8_dev(yp[npix_frame]), to_u8_dev(yp[2 * npix_frame]), a);
}

// pos2edge (run_render.py:31-57): per channel Sobel-3 (BORDER_REFLECT_101) in float64 on u8/255 with
// the background (alpha < 255) forced to 2, max magnitude over the 3 channels > 0.3.
__device__ __forceinline__ bool pos_is_edge(const uchar4* __restrict__ f, int x, int y, int H, int W) {
    double v[3][3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int yy = y + i - 1;
        yy = yy < 0 ? -yy : (yy >= H ? 2 * H - 2 - yy : yy);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            int xx = x + j - 1;
            xx = xx < 0 ? -xx : (xx >= W ? 2 * W - 2 - xx : xx);
            const uchar4 q = f[static_cast<size_t>(yy) * W + xx];
            const bool bg = q.w < 255;
            v[0][i][j] = bg ? 2.0 : static_cast<double>(__fdiv_rn(static_cast<float>(q.x), 255.0f));
            v[1][i][j] = bg ? 2.0 : static_cast<double>(__fdiv_rn(static_cast<float>(q.y), 255.0f));
            v[2][i][j] = bg ? 2.0 : static_cast<double>(__fdiv_rn(static_cast<float>(q.z), 255.0f));
        }
    }
    double best = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double gx = (v[c][0][2] - v[c][0][0]) + 2.0 * (v[c][1][2] - v[c][1][0]) + (v[c][2][2] - v[c][2][0]);
        const double gy = (v[c][2][0] - v[c][0][0]) + 2.0 * (v[c][2][1] - v[c][0][1]) + (v[c][2][2] - v[c][0][2]);
        best = fmax(best, sqrt(gx * gx + gy * gy));
    }
    return best > 0.3;
}
__global__ void pos2edge_kernel(const uchar4* __restrict__ pos, int B, int H, int W, uint8_t* __restrict__ edge) {
    const size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t npf = static_cast<size_t>(H) * W;
    if (p >= npf * B) return;
    edge[p] = pos_is_edge(pos + (p / npf) * npf, static_cast<int>(p % W), static_cast<int>((p / W) % H), H, W) ? 255 : 0;
}


__global__ void ingest_u8_kernel(const uchar4* __restrict__ color, const uchar4* __restrict__ pos,
                                 const uint8_t* __restrict__ edge, int derive_edge, int H, int W, size_t npix,
                                 __half* hi, __half* lo, float* f32, int pitch, int choff) {
    const size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (p >= npix) return;
    uchar4 c = color[p];
    const uchar4 q = pos[p];
    const float mask = __fdiv_rn(static_cast<float>(c.w), 255.0f);      // alpha BEFORE the edge burn-in (data.py:28)
    // overlap_edge_on_img: burn where the stored edge map (255 - pos2edge, run_render.py:117-120) is < 255.  derive_edge: no edge
    // map given - the same predicate straight from the pos frame (pos2edge fused into the ingest, nothing crosses PCIe)
    bool burn = false;
    if (edge) burn = edge[p] < 255;
    else if (derive_edge) {
        const size_t npf = static_cast<size_t>(H) * W;
        burn = pos_is_edge(pos + (p / npf) * npf, static_cast<int>(p % W), static_cast<int>((p / W) % H), H, W);
    }
    if (burn) { c.x = 0; c.y = 0; c.z = 0; }
    float v[8] = {norm_u8(c.x), norm_u8(c.y), norm_u8(c.z), mask, norm_u8(q.x), norm_u8(q.y), 0.0f, 0.0f};
    if (f32) store8_f32(f32 + p * pitch + choff, v);
    else store8(hi + p * pitch + choff, lo ? lo + p * pitch + choff : nullptr, v);
}


// ---- nn.InstanceNorm2d (norm_layer='instance_norm', models.py:34-35; defaults: affine=False, no running stats, eps 1e-5,
// biased variance) between a convolution and its activation.  The convolution's epilogue leaves its raw fp32 output
// x[B][HW][C] in a scratch buffer; (1) per-(frame, channel) sum and sum of squares, accumulated in fp64 (slices of the
// frame per block, one atomicAdd per (block, channel)); (2) mean / 1/sqrt(var + eps); (3) normalise, activation, and the
// stores the fused epilogue would have done (residual stream, un-ReLU'd second copy, fp16 hi[/lo] or fp32 NHWC output).
__global__ void instnorm_stats_kernel(const float* __restrict__ x, int HW, int C, int rows_per_slice, double* acc) {
    __shared__ double s_sum[8][33], s_sq[8][33];
    const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;              // 8 pixel rows x 32 channels per block step
    const int c = blockIdx.x * 32 + lane, n = blockIdx.z;
    const int p0 = blockIdx.y * rows_per_slice, p1 = min(HW, p0 + rows_per_slice);
    double s = 0.0, q = 0.0;
    if (c < C) {
        const float* base = x + static_cast<size_t>(n) * HW * C + c;
        for (int p = p0 + row; p < p1; p += 8) {
            const double v = static_cast<double>(base[static_cast<size_t>(p) * C]);
            s += v; q += v * v;
        }
    }
    s_sum[row][lane] = s; s_sq[row][lane] = q;
    __syncthreads();
    if (row == 0 && c < C) {
#pragma unroll
        for (int r = 1; r < 8; ++r) { s += s_sum[r][lane]; q += s_sq[r][lane]; }
        atomicAdd(acc + (static_cast<size_t>(n) * C + c) * 2, s);
        atomicAdd(acc + (static_cast<size_t>(n) * C + c) * 2 + 1, q);
    }
}

__global__ void instnorm_finish_kernel(const double* __restrict__ acc, int total, int HW, float2* stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double mean = acc[2 * i] / HW;
    double var = acc[2 * i + 1] / HW - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    stats[i] = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + 1e-5)));
}

__global__ void instnorm_apply_kernel(InstNormApply a) {
    const int G = a.C / 8;
    const size_t total = static_cast<size_t>(a.B) * a.HW * G;
    const size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (t >= total) return;
    const int g = static_cast<int>(t % G);
    const size_t pix = t / G;
    const int n = static_cast<int>(pix / a.HW);
    const float* src = a.x + pix * a.C + g * 8;
    const float4 x0 = reinterpret_cast<const float4*>(src)[0], x1 = reinterpret_cast<const float4*>(src)[1];
    float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    const float2* st = a.stats + static_cast<size_t>(n) * a.C + g * 8;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float2 ms = st[c];
        v[c] = (v[c] - ms.x) * ms.y;
        if (a.act == 1) v[c] = fmaxf(v[c], 0.0f);
        else if (a.act == 2) v[c] = fmaxf(v[c], 0.2f * v[c]);
    }
    if (a.resid) store8_f32(a.resid + pix * a.C + g * 8, v);
    if (a.out2_f32) store8_f32(a.out2_f32 + pix * a.out2_pitch + a.out2_choff + g * 8, v);
    else if (a.out2_hi) store8(a.out2_hi + pix * a.out2_pitch + a.out2_choff + g * 8,
                               a.out2_lo ? a.out2_lo + pix * a.out2_pitch + a.out2_choff + g * 8 : nullptr, v);
    if (a.out_relu) {
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = fmaxf(v[c], 0.0f);
    }
    if (a.out_f32) store8_f32(a.out_f32 + pix * a.out_pitch + a.out_choff + g * 8, v);
    else if (a.out_hi) store8(a.out_hi + pix * a.out_pitch + a.out_choff + g * 8,
                              a.out_lo ? a.out_lo + pix * a.out_pitch + a.out_choff + g * 8 : nullptr, v);
}

inline unsigned blocks_for(size_t n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

}  // namespace

cudaError_t ingest_f32(const float* x, int B, int cin, int cpad, int H, int W, __half* hi, __half* lo, float* f32, int pitch,
                       int choff, cudaStream_t st) {
    const size_t npf = static_cast<size_t>(H) * W, np = npf * B;
    ingest_f32_kernel<<<blocks_for(np, 256), 256, 0, st>>>(x, cin, cpad, npf, np, hi, lo, f32, pitch, choff);
    return cudaGetLastError();
}
cudaError_t ingest_u8(const uint8_t* color, const uint8_t* pos, const uint8_t* edge, int derive_edge, int B, int H, int W,
                      __half* hi, __half* lo, float* f32, int pitch, int choff, cudaStream_t st) {
    const size_t np = static_cast<size_t>(H) * W * B;
    ingest_u8_kernel<<<blocks_for(np, 256), 256, 0, st>>>(reinterpret_cast<const uchar4*>(color), reinterpret_cast<const uchar4*>(pos), edge,
                                                          derive_edge, H, W, np, hi, lo, f32, pitch, choff);
    return cudaGetLastError();
}
cudaError_t frames_to_tensor(const uint8_t* color, const uint8_t* pos, const uint8_t* edge, int B, int H, int W,
                             float* pre, float* mask, cudaStream_t st) {
    const size_t npf = static_cast<size_t>(H) * W, np = npf * B;
    frames_to_tensor_kernel<<<blocks_for(np, 256), 256, 0, st>>>(reinterpret_cast<const uchar4*>(color),
                                                                 reinterpret_cast<const uchar4*>(pos), edge, npf, np, pre, mask);
    return cudaGetLastError();
}
cudaError_t maxpool2(const __half* in_hi, const __half* in_lo, int in_pitch, int in_choff, int B, int Hin, int Win, int C,
                     __half* out_hi, __half* out_lo, int out_pitch, cudaStream_t st) {
    const size_t total = static_cast<size_t>(B) * (Hin / 2) * (Win / 2) * (C / 8);
    maxpool2_kernel<<<blocks_for(total, 256), 256, 0, st>>>(in_hi, in_lo, in_pitch, in_choff, B, Hin, Win, C, out_hi, out_lo, out_pitch);
    return cudaGetLastError();
}
cudaError_t maxpool2_f32(const float* in, int in_pitch, int in_choff, int B, int Hin, int Win, int C, float* out, int out_pitch,
                         cudaStream_t st) {
    const size_t total = static_cast<size_t>(B) * (Hin / 2) * (Win / 2) * (C / 4);
    maxpool2_f32_kernel<<<blocks_for(total, 256), 256, 0, st>>>(in, in_pitch, in_choff, B, Hin, Win, C, out, out_pitch);
    return cudaGetLastError();
}
cudaError_t to_image_space(const float* x, uint8_t* out, size_t n, cudaStream_t st) {
    to_image_space_kernel<<<blocks_for(n, 256), 256, 0, st>>>(x, out, n);
    return cudaGetLastError();
}
cudaError_t overlap_edge(const uint8_t* edge, uint8_t* rgba, size_t npix, cudaStream_t st) {
    overlap_edge_kernel<<<blocks_for(npix, 256), 256, 0, st>>>(edge, reinterpret_cast<uchar4*>(rgba), npix);
    return cudaGetLastError();
}
cudaError_t compose_rgba(const float* y, const float* mask, int B, int H, int W, uint8_t* out, cudaStream_t st) {
    const size_t npf = static_cast<size_t>(H) * W, np = npf * B;
    compose_rgba_kernel<<<blocks_for(np, 256), 256, 0, st>>>(y, mask, npf, np, reinterpret_cast<uchar4*>(out));
    return cudaGetLastError();
}
cudaError_t pos2edge(const uint8_t* pos, int B, int H, int W, uint8_t* edge, cudaStream_t st) {
    const size_t np = static_cast<size_t>(H) * W * B;
    pos2edge_kernel<<<blocks_for(np, 256), 256, 0, st>>>(reinterpret_cast<const uchar4*>(pos), B, H, W, edge);
    return cudaGetLastError();
}

cudaError_t instance_norm(const InstNormApply& a, double* acc, cudaStream_t st) {
    if (a.C % 8 || a.B < 1 || a.HW < 1 || !a.x || !a.stats || !acc) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(acc, 0, static_cast<size_t>(a.B) * a.C * 2 * sizeof(double), st);
    if (e != cudaSuccess) return e;
    int slices = (a.HW + 4095) / 4096;                       // >= 4096 pixels per block and channel group, at most 64 slices
    slices = slices > 64 ? 64 : slices;
    const int rows = (a.HW + slices - 1) / slices;
    instnorm_stats_kernel<<<dim3((a.C + 31) / 32, slices, a.B), 256, 0, st>>>(a.x, a.HW, a.C, rows, acc);
    instnorm_finish_kernel<<<blocks_for(static_cast<size_t>(a.B) * a.C, 128), 128, 0, st>>>(acc, a.B * a.C, a.HW, a.stats);
    const size_t total = static_cast<size_t>(a.B) * a.HW * (a.C / 8);
    instnorm_apply_kernel<<<blocks_for(total, 256), 256, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace dsu

// Parameter blocks shared by the fused convolution kernel and the engine that plans a network.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dsu {

constexpr int kTileH = 8;      // output patch rows per CTA
constexpr int kTileW = 16;     // output patch cols per CTA
constexpr int kTileM = 128;    // = UMMA M
constexpr int kChunkK = 64;    // fp16 K elements per smem row (128 B, SWIZZLE_128B)
constexpr int kWorkers = 128;  // producer / epilogue threads (warps 0-3)
constexpr int kThreads = 192;  // + warp 4 (MMA issue, TMEM alloc) + warp 5 (weight loader)
constexpr int kMaxSeg = 6;     // concat segments (x2 for the lo planes in exact mode)
constexpr int kMaxStages = 8;

// One 16-byte (8-channel) K slot of a chunk: which tap of which source segment fills it.
struct Slot {
    int8_t dy, dx;      // plain: tap offset (kh - pad, kw - pad).  RIC: dy = raster tap index 0..8
    uint8_t seg;        // source segment index
    uint8_t valid;      // 0 -> zero fill (K padding)
    uint16_t choff;     // first channel inside the segment buffer
    uint16_t pad_;
};
static_assert(sizeof(Slot) == 8, "Slot must be 8 bytes");

// One K chunk (8 slots = 64 K elements) of the implicit GEMM.
struct ChunkHdr {
    uint8_t ksteps;     // UMMA K=16 steps actually issued (1..4)
    uint8_t wide;       // 1: B tile holds [W_hi ; W_lo] (N = 2*Cout) - exact-mode hi-plane chunk
    uint16_t pad_;
    uint32_t b_off;     // byte offset of this chunk's B tile inside wpack
};
static_assert(sizeof(ChunkHdr) == 8, "ChunkHdr must be 8 bytes");

struct Seg {
    const __half* ptr;
    int pitch;          // channels per pixel in the buffer
    int pad_;
};

// Per bilinear stencil entry of a RIC level: weights of the 4 corners (invalid corners already 0)
// and the CLAMPED corner offsets relative to the output pixel {dy_lo, dy_hi, dx_lo, dx_hi}.
struct EpiParams {
    const float* scale;     // [Cout] pre-activation affine (folded BN / bias); never null
    const float* shift;
    const float* scale2;    // [Cout] post-activation affine (conv_11_a.2) or null
    const float* shift2;
    int act;                // 0 none, 1 ReLU, 2 LeakyReLU(0.2)
    int resid_in, resid_out;
    float* resid;           // fp32 residual stream [pix][Cout]
    __half* out_hi;         // main fp16 output (after residual), optional ReLU
    __half* out_lo;         // lo plane (exact mode) or null
    int out_pitch, out_choff, out_relu;
    __half* out2_hi;        // second, un-ReLU'd copy (skip connection) or null
    __half* out2_lo;
    int out2_pitch, out2_choff;
    // fused conv_12 (1x1, +bias, optional tanh) tail
    const float* w12;       // [3][Cout] or null
    const float* b12;       // [3]
    int tanh_flag;
    float* y_nchw;          // [B,3,H,W] fp32 or null
    uint8_t* y_rgba;        // [B,H,W,4] uint8 or null (to_image_space + alpha)
    const uint8_t* alpha_src;   // alpha byte of pixel p at alpha_src[p * alpha_stride]
    int alpha_stride;
};

struct ConvParams {
    int B, Hout, Wout;      // output geometry
    int Hin, Win;           // source buffer geometry (before the fused nearest x2)
    int Hv, Wv;             // virtual conv-input geometry = (Hin << up, Win << up)
    int stride, up, ric, exact;
    int nchunks, Cout, nstages, tmem_cols;
    int a_bytes, b_bytes;   // bytes per A / B stage
    const Slot* slots;      // [nchunks][8]
    const ChunkHdr* hdrs;   // [nchunks]
    const uint8_t* wpack;   // pre-swizzled B tiles
    Seg seg[kMaxSeg];
    const float4* ric_w;    // [8][Hout*Wout]
    const char4* ric_off;   // [8][Hout*Wout]
    EpiParams epi;
};

cudaError_t launch_conv(const ConvParams& p, cudaStream_t stream);
size_t conv_smem_bytes(const ConvParams& p);

}  // namespace dsu

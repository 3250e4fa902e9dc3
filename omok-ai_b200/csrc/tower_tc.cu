// tower_tc.cu -- K3 tensor-core path for the residual tower (stem + 3 bottleneck blocks).
//
// Replaces k_tower's fp32 CUDA-core GEMMs (network.rs:65-125, network-utils lib.rs:386-461) with
// tcgen05.mma kind::tf32, 3-pass hi/lo error compensation like fc0.  Per position the
// three 1x1 convolutions of a block are [81(->128) x K] . [K x N] GEMMs with tiny K and N, so the design
// is about latency and data placement, not MMA throughput:
//   * the ACTIVATIONS are the A operand and live in TENSOR MEMORY (lane = pixel, column = channel,
//     hi and lo parts side by side); threads write them with tcgen05.st straight from registers --
//     no swizzled shared-memory stores, no 8-byte-per-element smem footprint;
//   * the WEIGHTS are the B operand in shared memory, K-major SWIZZLE_128B, hi and lo parts,
//     pre-swizzled once on the device (k_tower_pack) so one block's 72 KB image arrives by plain bulk
//     copies (cp.async.bulk) into a double buffer, one block ahead;
//   * the residual stream x stays in REGISTERS (each thread owns one pixel row x 64 channels) across all
//     three blocks; depthwise 3x3 goes through a small fp32 shared tile.
// One CTA (8 warps) walks positions.  Warp w works on TMEM lane quadrant w%4 (pixels 32q..32q+31) and
// channel half w/4.  Thread 0 issues the MMAs; completion is a tcgen05.commit on an mbarrier.
//
// TMEM columns: X_hi [0,128) X_lo [128,256) H_hi [256,288) H_lo [288,320) D12 [320,352) D3 [352,480).
#include <cstdio>

#include "omk_internal.h"

namespace omk {

constexpr int TW_THREADS = 256;
constexpr uint32_t TW_TMEM_COLS = 512;
constexpr uint32_t C_XHI = 0, C_XLO = 128, C_HHI = 256, C_HLO = 288, C_D12 = 320, C_D3 = 352;
// per-block weight image (bytes): W0^T hi/lo [32 n][128 k] as 4 k-atoms of 4 KB, PW^T hi/lo [32][32], W2^T hi/lo [128][32]
constexpr int WI_W0HI = 0, WI_W0LO = 16384, WI_PWHI = 32768, WI_PWLO = 36864, WI_W2HI = 40960, WI_W2LO = 57344;
constexpr int WI_BYTES = 73728;
// small fp32 parameters image (floats): stem W[3][128], stem b[128], then per block b0[32] dw[9][32] b1[32] b2[128]
constexpr int PI_WSTEM = 0, PI_BSTEM = 384, PI_BLK0 = 512, PI_BLK = 32 + 288 + 32 + 128;  // 480
constexpr int PI_B0 = 0, PI_DW = 32, PI_B1 = 320, PI_B2 = 352;
constexpr int PI_FLOATS = PI_BLK0 + 3 * PI_BLK;  // 1952
constexpr int TW_HSTRIDE = 36;                   // fp32 depthwise tile row stride (floats)
// shared memory map (bytes from the 1024-aligned base)
constexpr int SM_W = 0;                                    // 2 x 72 KB weight images
constexpr int SM_H0 = 2 * WI_BYTES;                        // 147456: [128][36] fp32
constexpr int SM_PAR = SM_H0 + 128 * TW_HSTRIDE * 4;       // 165888
constexpr int SM_IMG = SM_PAR + PI_FLOATS * 4;             // 173696
constexpr int SM_BAR = SM_IMG + 256 * 4;                   // 174720
constexpr int TW_SMEM_BYTES = SM_BAR + 64 + 1024;
// pair kernel: H tile = 2 buffers x 162 rows (two positions), IMG = 3 images
constexpr int SM3_PAR = SM_H0 + 2 * 162 * TW_HSTRIDE * 4;   // 194112
constexpr int SM3_IMG = SM3_PAR + PI_FLOATS * 4;             // 201920
constexpr int SM3_BAR = SM3_IMG + 736 * 4;                   // 204864
constexpr int TW3_SMEM_BYTES = SM3_BAR + 64 + 1024;
constexpr uint32_t TW_IDESC_N32 = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t TW_IDESC_N128 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint32_t tw_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tw_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tw_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ uint64_t tw_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void tw_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tw_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// asynchronous TMEM load of 16 columns of this thread's lane; results are valid after tw_wait_ld()
__device__ __forceinline__ void tw_ld16(uint32_t taddr, float *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tw_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tw_st16(uint32_t taddr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// split 16 fp32 values into TF32-exact high parts and residuals and store them as A-operand columns
__device__ __forceinline__ void tw_store_split16(uint32_t taddr_hi, uint32_t taddr_lo, const float *v) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        hi[i] = __float_as_uint(v[i]) & 0xFFFFE000u;
        lo[i] = __float_as_uint(v[i] - __uint_as_float(hi[i]));
    }
    tw_st16(taddr_hi, hi);
    tw_st16(taddr_lo, lo);
}
__device__ long long g_tw_dbg[64];  // phase timestamps of CTA 0's second position (omk_debug_tower_timing)
#define TW_STAMP(i) do { if (dbg_on && t == 0) g_tw_dbg[(i)] = clock64(); } while (0)
__device__ __forceinline__ float tw_lrelu(float v) { return fmaxf(v, 0.2f * v); }  // alpha < 1: max(v, alpha v)
// column of the a-th 32-column partial accumulator used by conv0 / conv1: D12, four slices of the (then idle) D3
// region, and the 32 spare columns at the top
__device__ __forceinline__ constexpr uint32_t conv0_acc(int a) { return a == 0 ? C_D12 : (a < 5 ? C_D3 + 32u * (uint32_t)(a - 1) : 480u); }


// Coalesced write-out of one warp's 32 pixel rows x 64 channels.  A thread owns a 256-byte row segment, so direct
// per-thread stores make every warp instruction touch 32 different cache lines (measured: ~16k clk per iteration in
// the LSU).  Instead each warp transposes through a 4.6 KB staging tile: rows in, 128-byte line segments out.
// `off` = element offset of this lane's segment in the output arrays, or -1 for a padded / out-of-batch row.
__device__ __forceinline__ void tw_store_part(float *stage, int lane, const long long *orow, const float *v, float *dst) {
#pragma unroll
    for (int cp = 0; cp < 2; ++cp) {  // two passes of 32 channels
#pragma unroll
        for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4 *>(stage + lane * TW_HSTRIDE + k * 4) =
                make_float4(v[cp * 32 + k * 4 + 0], v[cp * 32 + k * 4 + 1], v[cp * 32 + k * 4 + 2], v[cp * 32 + k * 4 + 3]);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + (lane >> 3), ch = lane & 7;
            const float4 val = *reinterpret_cast<const float4 *>(stage + rr * TW_HSTRIDE + ch * 4);
            const long long o = orow[it];
            if (o >= 0) *reinterpret_cast<float4 *>(dst + o + cp * 32 + ch * 4) = val;
        }
        __syncwarp();
    }
}
__device__ __forceinline__ void tw_store_out(float *stage, int lane, long long off, const float *x, float *act0, float *act0_hi,
                                             float *act0_lo) {
    long long orow[8];  // output offsets of the eight rows this lane writes (row it*4 + lane/8), fetched once
#pragma unroll
    for (int it = 0; it < 8; ++it) orow[it] = __shfl_sync(0xffffffffu, off, it * 4 + (lane >> 3));
    if (act0_hi) {
        float part[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) part[c] = __uint_as_float(__float_as_uint(x[c]) & 0xFFFFE000u);
        tw_store_part(stage, lane, orow, part, act0_hi);
#pragma unroll
        for (int c = 0; c < 64; ++c) part[c] = x[c] - part[c];
        tw_store_part(stage, lane, orow, part, act0_lo);
    } else {
        tw_store_part(stage, lane, orow, x, act0);
    }
}

__device__ __forceinline__ void tw_bulk_load(uint32_t dst, const uint8_t *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__global__ void __launch_bounds__(TW_THREADS, 1)
    k_tower_tc(const uint8_t *__restrict__ wimg, const float *__restrict__ pimg, const NNIn *__restrict__ nn_in,
               const float *__restrict__ images, const uint32_t *n_req, int max_rows, float *__restrict__ act0,
               float *__restrict__ act0_hi, float *__restrict__ act0_lo) {
    extern __shared__ uint8_t tw_smem_raw[];
    const int rows = (int)min(*n_req, (uint32_t)max_rows);
    if ((int)blockIdx.x >= rows) return;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int q = warp & 3, half = warp >> 2;
    // pixel row owned by this thread == its TMEM lane (a warp reaches only its own lane quadrant).  Dealing pixels
    // round-robin over the quadrants was measured slower: SIMT width is wasted either way, and it adds smem traffic.
    const int p = q * 32 + lane;
    const bool real = p < kCells;

    uint8_t *sm = tw_smem_raw + ((1024u - (tw_smem_u32(tw_smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = tw_smem_u32(sm);
    float *H0 = reinterpret_cast<float *>(sm + SM_H0);
    float *PAR = reinterpret_cast<float *>(sm + SM_PAR);
    float *IMG = reinterpret_cast<float *>(sm + SM_IMG);
    const uint32_t bar_w0 = sbase + SM_BAR, bar_mma = sbase + SM_BAR + 16, tmem_slot = sbase + SM_BAR + 24;

    for (int i = t; i < PI_FLOATS; i += TW_THREADS) PAR[i] = pimg[i];
    if (t == 0) {
        tw_mbar_init(bar_w0, 1);
        tw_mbar_init(bar_w0 + 8, 1);
        tw_mbar_init(bar_mma, 8);  // one tcgen05.commit per warp: every warp's lane 0 issues its own MMA chain
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TW_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);  // this warp's lane quadrant

    uint32_t g = 0;          // running block counter: weight buffer = g & 1, its phase parity = (g >> 1) & 1
    uint32_t mma_uses = 0;   // completed uses of bar_mma
    auto issue_weights = [&](uint32_t gg) {  // thread 0 only
        const uint32_t buf = gg & 1u, bar = bar_w0 + 8 * buf;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer may have served as a staging tile (generic proxy)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)WI_BYTES) : "memory");
        const uint8_t *src = wimg + (size_t)(gg % 3u) * WI_BYTES;
        for (int c = 0; c < WI_BYTES; c += 8192) tw_bulk_load(sbase + SM_W + buf * WI_BYTES + c, src + c, 8192, bar);
    };
    if (t == 0) issue_weights(0);

    NNIn cur_in{};
    if (!images) cur_in = nn_in[blockIdx.x];
    int pos_iter = 0;
    for (int row = blockIdx.x; row < rows; row += gridDim.x, ++pos_iter) {
        const bool dbg_on = blockIdx.x == 0 && pos_iter == 1;
        TW_STAMP(0);
        // ---- input image (the reference's 243-float slot read as [81][3]) ----
        if (t < 243) {
            if (images) {
                IMG[t] = images[(size_t)row * 243 + t];
            } else {
                IMG[t] = image_value(cur_in.black, cur_in.white, cur_in.meta & 1u, (cur_in.meta >> 1) & 1u, t);
            }
        }
        // prefetch the next position's request row: its global-load latency hides behind this whole position
        if (!images && row + (int)gridDim.x < rows) cur_in = nn_in[row + gridDim.x];
        __syncthreads();
        TW_STAMP(40);
        // ---- stem 1x1 conv 3 -> 128 (network.rs:65-79): this thread's pixel, its 64 channels ----
        float x[64];
        {
            const int pc = real ? p : 0;  // padded rows recompute pixel 0 (harmless, never stored)
            const float v0 = IMG[3 * pc], v1 = IMG[3 * pc + 1], v2 = IMG[3 * pc + 2];
#pragma unroll
            for (int c = 0; c < 64; c += 4) {
                const int ch = half * 64 + c;
                const float4 b = *reinterpret_cast<const float4 *>(PAR + PI_BSTEM + ch);
                const float4 w0 = *reinterpret_cast<const float4 *>(PAR + PI_WSTEM + ch);
                const float4 w1 = *reinterpret_cast<const float4 *>(PAR + PI_WSTEM + 128 + ch);
                const float4 w2 = *reinterpret_cast<const float4 *>(PAR + PI_WSTEM + 256 + ch);
                float2 s01 = __ffma2_rn(make_float2(v0, v0), make_float2(w0.x, w0.y), make_float2(b.x, b.y));
                float2 s23 = __ffma2_rn(make_float2(v0, v0), make_float2(w0.z, w0.w), make_float2(b.z, b.w));
                s01 = __ffma2_rn(make_float2(v1, v1), make_float2(w1.x, w1.y), s01);
                s23 = __ffma2_rn(make_float2(v1, v1), make_float2(w1.z, w1.w), s23);
                s01 = __ffma2_rn(make_float2(v2, v2), make_float2(w2.x, w2.y), s01);
                s23 = __ffma2_rn(make_float2(v2, v2), make_float2(w2.z, w2.w), s23);
                x[c + 0] = tw_lrelu(s01.x);
                x[c + 1] = tw_lrelu(s01.y);
                x[c + 2] = tw_lrelu(s23.x);
                x[c + 3] = tw_lrelu(s23.y);
            }
            TW_STAMP(41);
#pragma unroll
            for (int c = 0; c < 64; c += 16) tw_store_split16(tlane + C_XHI + half * 64 + c, tlane + C_XLO + half * 64 + c, x + c);
            TW_STAMP(42);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        TW_STAMP(1);

        for (int r = 0; r < 3; ++r, ++g) {
            const float *bp = PAR + PI_BLK0 + r * PI_BLK;
            const uint32_t wb = sbase + SM_W + (g & 1u) * WI_BYTES;
            // ================= conv0: 1x1 128 -> 32 (A = X in TMEM) =================
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (lane == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp == 0) issue_weights(g + 1);  // the other buffer's last reader (block g-1) has completed
                if (warp < 6) {
                    // six independent accumulator chains (3 products x 2 K-halves), one per issuing warp: the single-thread
                    // issue path costs ~60 clk per MMA, so the chains are issued in parallel from six threads
                    tw_mbar_wait(bar_w0 + 8 * (g & 1u), (g >> 1) & 1u);
                    const int pass = warp >> 1, h = warp & 1;
                    const uint32_t acol = pass == 0 ? C_XLO : C_XHI;
                    const uint32_t bimg = pass == 1 ? WI_W0LO : WI_W0HI;
#pragma unroll 1
                    for (int ks = 0; ks < 8; ++ks) {
                        const int kk = ks + 8 * h;
                        const uint32_t boff = (uint32_t)((kk >> 2) * 4096 + (kk & 3) * 32);
                        tw_umma_ts(tmem_base + conv0_acc(warp), tmem_base + acol + kk * 8, tw_desc_sw128(wb + bimg + boff),
                                   TW_IDESC_N32, ks != 0);
                    }
                }
                tw_commit(bar_mma);
            }
            tw_mbar_wait(bar_mma, mma_uses & 1u);
            ++mma_uses;
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            TW_STAMP(2 + r * 8 + 0);
            // epilogue 1: + b0, lrelu -> fp32 tile for the depthwise stencil (this thread: 16 channels of its pixel)
            {
                float d[16], e[5][16];
                tw_ld16(tlane + conv0_acc(0) + half * 16, d);
#pragma unroll
                for (int a = 1; a < 6; ++a) tw_ld16(tlane + conv0_acc(a) + half * 16, e[a - 1]);
                tw_wait_ld();  // one wait for all six partial accumulators
#pragma unroll
                for (int c = 0; c < 16; ++c) d[c] = ((d[c] + e[0][c]) + (e[1][c] + e[2][c])) + (e[3][c] + e[4][c]);
#pragma unroll
                for (int c = 0; c < 16; c += 4) {
                    const float4 b = *reinterpret_cast<const float4 *>(bp + PI_B0 + half * 16 + c);
                    float4 o;
                    o.x = tw_lrelu(d[c + 0] + b.x);
                    o.y = tw_lrelu(d[c + 1] + b.y);
                    o.z = tw_lrelu(d[c + 2] + b.z);
                    o.w = tw_lrelu(d[c + 3] + b.w);
                    *reinterpret_cast<float4 *>(H0 + p * TW_HSTRIDE + half * 16 + c) = o;
                }
            }
            __syncthreads();
            TW_STAMP(2 + r * 8 + 1);
            // conv1 depthwise 3x3, SAME zero padding, no bias (lib.rs:204-216) -> A operand of the pointwise conv
            {
                float a[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) a[c] = 0.0f;
                if (real) {
                    const int y = p / kSide, xx0 = p % kSide;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int yy = y + ky - 1;
                        if (yy < 0 || yy >= kSide) continue;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const int xx = xx0 + kx - 1;
                            if (xx < 0 || xx >= kSide) continue;
                            const float *hp = H0 + (yy * kSide + xx) * TW_HSTRIDE + half * 16;
                            const float *wp = bp + PI_DW + (ky * 3 + kx) * 32 + half * 16;
#pragma unroll
                            for (int c = 0; c < 16; c += 4) {
                                const float4 h = *reinterpret_cast<const float4 *>(hp + c);
                                const float4 w = *reinterpret_cast<const float4 *>(wp + c);
                                // packed fp32x2 FMA (sm_100): same results as two FFMAs, half the issue slots
                                float2 lo2 = __ffma2_rn(make_float2(h.x, h.y), make_float2(w.x, w.y), make_float2(a[c + 0], a[c + 1]));
                                float2 hi2 = __ffma2_rn(make_float2(h.z, h.w), make_float2(w.z, w.w), make_float2(a[c + 2], a[c + 3]));
                                a[c + 0] = lo2.x; a[c + 1] = lo2.y; a[c + 2] = hi2.x; a[c + 3] = hi2.y;
                            }
                        }
                    }
                }
                tw_store_split16(tlane + C_HHI + half * 16, tlane + C_HLO + half * 16, a);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            TW_STAMP(2 + r * 8 + 2);
            // ================= conv1 pointwise: 1x1 32 -> 32 (A = H in TMEM) =================
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (lane == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp < 3) {  // one product per issuing warp, three accumulators
                    const uint32_t acol = warp == 0 ? C_HLO : C_HHI;
                    const uint32_t bimg = warp == 1 ? WI_PWLO : WI_PWHI;
#pragma unroll 1
                    for (int ks = 0; ks < 4; ++ks)
                        tw_umma_ts(tmem_base + conv0_acc(warp), tmem_base + acol + ks * 8, tw_desc_sw128(wb + bimg + ks * 32),
                                   TW_IDESC_N32, ks != 0);
                }
                tw_commit(bar_mma);
            }
            tw_mbar_wait(bar_mma, mma_uses & 1u);
            ++mma_uses;
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            TW_STAMP(2 + r * 8 + 3);
            {   // epilogue 2: + b1, lrelu -> A operand of conv2
                float d[16], e[16], f[16];
                tw_ld16(tlane + conv0_acc(0) + half * 16, d);
                tw_ld16(tlane + conv0_acc(1) + half * 16, e);
                tw_ld16(tlane + conv0_acc(2) + half * 16, f);
                tw_wait_ld();
#pragma unroll
                for (int c = 0; c < 16; c += 4) {
                    const float4 b = *reinterpret_cast<const float4 *>(bp + PI_B1 + half * 16 + c);
                    d[c + 0] = tw_lrelu((d[c + 0] + e[c + 0]) + f[c + 0] + b.x);
                    d[c + 1] = tw_lrelu((d[c + 1] + e[c + 1]) + f[c + 1] + b.y);
                    d[c + 2] = tw_lrelu((d[c + 2] + e[c + 2]) + f[c + 2] + b.z);
                    d[c + 3] = tw_lrelu((d[c + 3] + e[c + 3]) + f[c + 3] + b.w);
                }
                tw_store_split16(tlane + C_HHI + half * 16, tlane + C_HLO + half * 16, d);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            TW_STAMP(2 + r * 8 + 4);
            // ================= conv2: 1x1 32 -> 128 (A = H in TMEM), + x, lrelu =================
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (lane == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp == 0) {  // N = 128 MMAs are tensor-bound (64 clk each): one chain
#pragma unroll 1
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t bhi = tw_desc_sw128(wb + WI_W2HI + ks * 32), blo = tw_desc_sw128(wb + WI_W2LO + ks * 32);
                        tw_umma_ts(tmem_base + C_D3, tmem_base + C_HLO + ks * 8, bhi, TW_IDESC_N128, ks != 0);
                        tw_umma_ts(tmem_base + C_D3, tmem_base + C_HHI + ks * 8, blo, TW_IDESC_N128, 1u);
                        tw_umma_ts(tmem_base + C_D3, tmem_base + C_HHI + ks * 8, bhi, TW_IDESC_N128, 1u);
                    }
                }
                tw_commit(bar_mma);
            }
            tw_mbar_wait(bar_mma, mma_uses & 1u);
            ++mma_uses;
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            TW_STAMP(2 + r * 8 + 5);
            {   // epilogue 3: x = lrelu(conv2 + b2 + x), next block's A operand
                float d[64];
#pragma unroll
                for (int c = 0; c < 64; c += 16) tw_ld16(tlane + C_D3 + half * 64 + c, d + c);
                tw_wait_ld();
#pragma unroll
                for (int c = 0; c < 64; c += 16) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {  // padded rows (p >= 81) carry harmless garbage; they are never stored
                        const float4 b = *reinterpret_cast<const float4 *>(bp + PI_B2 + half * 64 + c + i);
                        const float2 t01 = __fadd2_rn(__fadd2_rn(make_float2(d[c + i + 0], d[c + i + 1]), make_float2(b.x, b.y)),
                                                      make_float2(x[c + i + 0], x[c + i + 1]));
                        const float2 t23 = __fadd2_rn(__fadd2_rn(make_float2(d[c + i + 2], d[c + i + 3]), make_float2(b.z, b.w)),
                                                      make_float2(x[c + i + 2], x[c + i + 3]));
                        const float2 u01 = __fmul2_rn(t01, make_float2(0.2f, 0.2f)), u23 = __fmul2_rn(t23, make_float2(0.2f, 0.2f));
                        x[c + i + 0] = fmaxf(t01.x, u01.x);
                        x[c + i + 1] = fmaxf(t01.y, u01.y);
                        x[c + i + 2] = fmaxf(t23.x, u23.x);
                        x[c + i + 3] = fmaxf(t23.y, u23.y);
                    }
                    if (r < 2) tw_store_split16(tlane + C_XHI + half * 64 + c, tlane + C_XLO + half * 64 + c, x + c);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            TW_STAMP(2 + r * 8 + 6);
        }
        // ---- flatten NHWC (network.rs:127-137): this thread's pixel row, its 64 channels ----
        TW_STAMP(30);
        {   // the weight buffer of the block just finished is idle until the next conv0 phase: staging tiles live there
            float *stage = reinterpret_cast<float *>(sm + SM_W + ((g - 1u) & 1u) * WI_BYTES) + warp * (32 * TW_HSTRIDE);
            const long long off = real ? (long long)row * 10368 + (long long)p * 128 + half * 64 : -1ll;
            tw_store_out(stage, lane, off, x, act0, act0_hi, act0_lo);
        }
        TW_STAMP(31);
    }
    // (timestamps: 0 start, 1 stem done, then per block: conv0 done, E1+sync, dw stored, conv1 done, E2 stored, conv2 done, E3 stored)
    // drain the weight prefetch that was issued one block ahead, then release TMEM
    if (t == 0) tw_mbar_wait(bar_w0 + 8 * (g & 1u), (g >> 1) & 1u);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TW_TMEM_COLS) : "memory");
    }
}

__global__ void __launch_bounds__(TW_THREADS, 1)
    k_tower_tc3(const uint8_t *__restrict__ wimg, const float *__restrict__ pimg, const NNIn *__restrict__ nn_in,
               const float *__restrict__ images, const uint32_t *n_req, int max_rows, float *__restrict__ act0,
               float *__restrict__ act0_hi, float *__restrict__ act0_lo) {
    extern __shared__ uint8_t tw_smem_raw[];
    const int rows = (int)min(*n_req, (uint32_t)max_rows);
    const int n_triples = (rows + 2) / 3, n_pairs = (int)gridDim.x >> 1, pair = (int)blockIdx.x >> 1;
    if (pair >= n_triples) return;  // uniform for both CTAs of the pair, before any cluster barrier
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int q = warp & 3, half = warp >> 2;
    // A CTA PAIR walks position TRIPLES: the 243 pixel rows of three positions fill 243 of the pair's 256 TMEM lanes
    // (one position alone fills only 81 of a CTA's 128).  Row R of the pair = 128*rank + TMEM lane; position j = R/81.
    const int R = (int)rank * 128 + q * 32 + lane;
    const int j = R / kCells, p = R - j * kCells;       // position within the triple, pixel within the position
    const int slot = j - (int)rank;                      // H tile slot: CTA r keeps positions r and r+1 (162 rows)

    uint8_t *sm = tw_smem_raw + ((1024u - (tw_smem_u32(tw_smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = tw_smem_u32(sm);
    float *H0 = reinterpret_cast<float *>(sm + SM_H0);
    float *PAR = reinterpret_cast<float *>(sm + SM3_PAR);
    float *IMG = reinterpret_cast<float *>(sm + SM3_IMG);
    const uint32_t bar_w0 = sbase + SM3_BAR, bar_mma = sbase + SM3_BAR + 16, tmem_slot = sbase + SM3_BAR + 24;
    // the pair's middle position straddles the two CTAs: its H rows are mirrored into the peer's tile through DSMEM
    uint32_t peer_h0;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_h0) : "r"(sbase + SM_H0), "r"(rank ^ 1u));

    for (int i = t; i < PI_FLOATS; i += TW_THREADS) PAR[i] = pimg[i];
    if (t == 0) {
        tw_mbar_init(bar_w0, 1);
        tw_mbar_init(bar_w0 + 8, 1);
        tw_mbar_init(bar_mma, 8);  // one tcgen05.commit per warp: every warp's lane 0 issues its own MMA chain
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TW_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);  // this warp's lane quadrant

    uint32_t g = 0;          // running block counter: weight buffer = g & 1, its phase parity = (g >> 1) & 1
    uint32_t mma_uses = 0;   // completed uses of bar_mma
    auto issue_weights = [&](uint32_t gg) {  // thread 0 only
        const uint32_t buf = gg & 1u, bar = bar_w0 + 8 * buf;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer may have served as a staging tile (generic proxy)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)WI_BYTES) : "memory");
        const uint8_t *src = wimg + (size_t)(gg % 3u) * WI_BYTES;
        for (int c = 0; c < WI_BYTES; c += 8192) tw_bulk_load(sbase + SM_W + buf * WI_BYTES + c, src + c, 8192, bar);
    };
    if (t == 0) issue_weights(0);

    NNIn cur0{}, cur1{}, cur2{};
    if (!images) {
        cur0 = nn_in[min(pair * 3 + 0, rows - 1)];
        cur1 = nn_in[min(pair * 3 + 1, rows - 1)];
        cur2 = nn_in[min(pair * 3 + 2, rows - 1)];
    }
    int pos_iter = 0;
    for (int tr = pair; tr < n_triples; tr += n_pairs, ++pos_iter) {
        const bool dbg_on = blockIdx.x == 0 && pos_iter == 1;
        const int row = tr * 3 + j;                       // this thread's position (global row of the batch)
        const bool real = R < 3 * kCells && row < rows;   // 13 padding lanes per pair; the last triple may be partial
        TW_STAMP(0);
        // ---- input images of the triple (the reference's 243-float slot read as [81][3]) ----
        if (t < 243) {
            if (images) {
                IMG[t] = images[(size_t)min(tr * 3 + 0, rows - 1) * 243 + t];
                IMG[243 + t] = images[(size_t)min(tr * 3 + 1, rows - 1) * 243 + t];
                IMG[486 + t] = images[(size_t)min(tr * 3 + 2, rows - 1) * 243 + t];
            } else {
                IMG[t] = image_value(cur0.black, cur0.white, cur0.meta & 1u, (cur0.meta >> 1) & 1u, t);
                IMG[243 + t] = image_value(cur1.black, cur1.white, cur1.meta & 1u, (cur1.meta >> 1) & 1u, t);
                IMG[486 + t] = image_value(cur2.black, cur2.white, cur2.meta & 1u, (cur2.meta >> 1) & 1u, t);
            }
        }
        // prefetch the next triple's request rows: their global-load latency hides behind this whole iteration
        if (!images && tr + n_pairs < n_triples) {
            cur0 = nn_in[min((tr + n_pairs) * 3 + 0, rows - 1)];
            cur1 = nn_in[min((tr + n_pairs) * 3 + 1, rows - 1)];
            cur2 = nn_in[min((tr + n_pairs) * 3 + 2, rows - 1)];
        }
        __syncthreads();
        TW_STAMP(40);
        // ---- stem 1x1 conv 3 -> 128 (network.rs:65-79): this thread's pixel, its 64 channels ----
        float x[64];
        {
            const int pc = real ? p : 0;  // padded rows recompute pixel 0 (harmless, never stored)
            const float *im = IMG + (real ? j : 0) * 243;
            const float v0 = im[3 * pc], v1 = im[3 * pc + 1], v2 = im[3 * pc + 2];
#pragma unroll
            for (int c = 0; c < 64; c += 4) {
                const int ch = half * 64 + c;
                const float4 b = *reinterpret_cast<const float4 *>(PAR + PI_BSTEM + ch);
                const float4 w0 = *reinterpret_cast<const float4 *>(PAR + PI_WSTEM + ch);
                const float4 w1 = *reinterpret_cast<const float4 *>(PAR + PI_WSTEM + 128 + ch);
                const float4 w2 = *reinterpret_cast<const float4 *>(PAR + PI_WSTEM + 256 + ch);
                float2 s01 = __ffma2_rn(make_float2(v0, v0), make_float2(w0.x, w0.y), make_float2(b.x, b.y));
                float2 s23 = __ffma2_rn(make_float2(v0, v0), make_float2(w0.z, w0.w), make_float2(b.z, b.w));
                s01 = __ffma2_rn(make_float2(v1, v1), make_float2(w1.x, w1.y), s01);
                s23 = __ffma2_rn(make_float2(v1, v1), make_float2(w1.z, w1.w), s23);
                s01 = __ffma2_rn(make_float2(v2, v2), make_float2(w2.x, w2.y), s01);
                s23 = __ffma2_rn(make_float2(v2, v2), make_float2(w2.z, w2.w), s23);
                x[c + 0] = tw_lrelu(s01.x);
                x[c + 1] = tw_lrelu(s01.y);
                x[c + 2] = tw_lrelu(s23.x);
                x[c + 3] = tw_lrelu(s23.y);
            }
            TW_STAMP(41);
#pragma unroll
            for (int c = 0; c < 64; c += 16) tw_store_split16(tlane + C_XHI + half * 64 + c, tlane + C_XLO + half * 64 + c, x + c);
            TW_STAMP(42);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        TW_STAMP(1);

        for (int r = 0; r < 3; ++r, ++g) {
            const float *bp = PAR + PI_BLK0 + r * PI_BLK;
            float *H0b = H0 + (r & 1) * (162 * TW_HSTRIDE);      // double-buffered: the peer may still read the other one
            const uint32_t peer_h0b = peer_h0 + (uint32_t)((r & 1) * (162 * TW_HSTRIDE * 4));
            const uint32_t wb = sbase + SM_W + (g & 1u) * WI_BYTES;
            // ================= conv0: 1x1 128 -> 32 (A = X in TMEM) =================
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (lane == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp == 0) issue_weights(g + 1);  // the other buffer's last reader (block g-1) has completed
                if (warp < 6) {
                    // six independent accumulator chains (3 products x 2 K-halves), one per issuing warp: the single-thread
                    // issue path costs ~60 clk per MMA, so the chains are issued in parallel from six threads
                    tw_mbar_wait(bar_w0 + 8 * (g & 1u), (g >> 1) & 1u);
                    const int pass = warp >> 1, h = warp & 1;
                    const uint32_t acol = pass == 0 ? C_XLO : C_XHI;
                    const uint32_t bimg = pass == 1 ? WI_W0LO : WI_W0HI;
#pragma unroll 1
                    for (int ks = 0; ks < 8; ++ks) {
                        const int kk = ks + 8 * h;
                        const uint32_t boff = (uint32_t)((kk >> 2) * 4096 + (kk & 3) * 32);
                        tw_umma_ts(tmem_base + conv0_acc(warp), tmem_base + acol + kk * 8, tw_desc_sw128(wb + bimg + boff),
                                   TW_IDESC_N32, ks != 0);
                    }
                }
                tw_commit(bar_mma);
            }
            tw_mbar_wait(bar_mma, mma_uses & 1u);
            ++mma_uses;
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            TW_STAMP(2 + r * 8 + 0);
            // epilogue 1: + b0, lrelu -> fp32 tile for the depthwise stencil (this thread: 16 channels of its pixel)
            {
                float d[16], e[5][16];
                tw_ld16(tlane + conv0_acc(0) + half * 16, d);
#pragma unroll
                for (int a = 1; a < 6; ++a) tw_ld16(tlane + conv0_acc(a) + half * 16, e[a - 1]);
                tw_wait_ld();  // one wait for all six partial accumulators
#pragma unroll
                for (int c = 0; c < 16; ++c) d[c] = ((d[c] + e[0][c]) + (e[1][c] + e[2][c])) + (e[3][c] + e[4][c]);
#pragma unroll
                for (int c = 0; c < 16; c += 4) {
                    const float4 b = *reinterpret_cast<const float4 *>(bp + PI_B0 + half * 16 + c);
                    float4 o;
                    o.x = tw_lrelu(d[c + 0] + b.x);
                    o.y = tw_lrelu(d[c + 1] + b.y);
                    o.z = tw_lrelu(d[c + 2] + b.z);
                    o.w = tw_lrelu(d[c + 3] + b.w);
                    if (real) {
                        *reinterpret_cast<float4 *>(H0b + (slot * kCells + p) * TW_HSTRIDE + half * 16 + c) = o;
                        // only the band the peer's stencil can reach (pixels 37..46 from CTA 0, 47..56 from CTA 1) is mirrored
                        if (j == 1 && (rank == 0 ? p >= 37 : p <= 56)) {  // the peer keeps this position in its slot 1 - peer_rank = rank
                            const uint32_t ra = peer_h0b + (uint32_t)((((int)rank * kCells + p) * TW_HSTRIDE + half * 16 + c) * 4);
                            asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ra), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
                        }
                    }
                }
            }
            // both CTAs' rows (local and mirrored) must be visible before the stencil: cluster-wide barrier
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
            TW_STAMP(2 + r * 8 + 1);
            // conv1 depthwise 3x3, SAME zero padding, no bias (lib.rs:204-216) -> A operand of the pointwise conv
            {
                float a[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) a[c] = 0.0f;
                if (real) {
                    const int y = p / kSide, xx0 = p % kSide;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int yy = y + ky - 1;
                        if (yy < 0 || yy >= kSide) continue;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const int xx = xx0 + kx - 1;
                            if (xx < 0 || xx >= kSide) continue;
                            const float *hp = H0b + (slot * kCells + yy * kSide + xx) * TW_HSTRIDE + half * 16;
                            const float *wp = bp + PI_DW + (ky * 3 + kx) * 32 + half * 16;
#pragma unroll
                            for (int c = 0; c < 16; c += 4) {
                                const float4 h = *reinterpret_cast<const float4 *>(hp + c);
                                const float4 w = *reinterpret_cast<const float4 *>(wp + c);
                                // packed fp32x2 FMA (sm_100): same results as two FFMAs, half the issue slots
                                float2 lo2 = __ffma2_rn(make_float2(h.x, h.y), make_float2(w.x, w.y), make_float2(a[c + 0], a[c + 1]));
                                float2 hi2 = __ffma2_rn(make_float2(h.z, h.w), make_float2(w.z, w.w), make_float2(a[c + 2], a[c + 3]));
                                a[c + 0] = lo2.x; a[c + 1] = lo2.y; a[c + 2] = hi2.x; a[c + 3] = hi2.y;
                            }
                        }
                    }
                }
                tw_store_split16(tlane + C_HHI + half * 16, tlane + C_HLO + half * 16, a);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            TW_STAMP(2 + r * 8 + 2);
            // ================= conv1 pointwise: 1x1 32 -> 32 (A = H in TMEM) =================
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (lane == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp < 3) {  // one product per issuing warp, three accumulators
                    const uint32_t acol = warp == 0 ? C_HLO : C_HHI;
                    const uint32_t bimg = warp == 1 ? WI_PWLO : WI_PWHI;
#pragma unroll 1
                    for (int ks = 0; ks < 4; ++ks)
                        tw_umma_ts(tmem_base + conv0_acc(warp), tmem_base + acol + ks * 8, tw_desc_sw128(wb + bimg + ks * 32),
                                   TW_IDESC_N32, ks != 0);
                }
                tw_commit(bar_mma);
            }
            tw_mbar_wait(bar_mma, mma_uses & 1u);
            ++mma_uses;
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            TW_STAMP(2 + r * 8 + 3);
            {   // epilogue 2: + b1, lrelu -> A operand of conv2
                float d[16], e[16], f[16];
                tw_ld16(tlane + conv0_acc(0) + half * 16, d);
                tw_ld16(tlane + conv0_acc(1) + half * 16, e);
                tw_ld16(tlane + conv0_acc(2) + half * 16, f);
                tw_wait_ld();
#pragma unroll
                for (int c = 0; c < 16; c += 4) {
                    const float4 b = *reinterpret_cast<const float4 *>(bp + PI_B1 + half * 16 + c);
                    d[c + 0] = tw_lrelu((d[c + 0] + e[c + 0]) + f[c + 0] + b.x);
                    d[c + 1] = tw_lrelu((d[c + 1] + e[c + 1]) + f[c + 1] + b.y);
                    d[c + 2] = tw_lrelu((d[c + 2] + e[c + 2]) + f[c + 2] + b.z);
                    d[c + 3] = tw_lrelu((d[c + 3] + e[c + 3]) + f[c + 3] + b.w);
                }
                tw_store_split16(tlane + C_HHI + half * 16, tlane + C_HLO + half * 16, d);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            TW_STAMP(2 + r * 8 + 4);
            // ================= conv2: 1x1 32 -> 128 (A = H in TMEM), + x, lrelu =================
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (lane == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp == 0) {  // N = 128 MMAs are tensor-bound (64 clk each): one chain
#pragma unroll 1
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t bhi = tw_desc_sw128(wb + WI_W2HI + ks * 32), blo = tw_desc_sw128(wb + WI_W2LO + ks * 32);
                        tw_umma_ts(tmem_base + C_D3, tmem_base + C_HLO + ks * 8, bhi, TW_IDESC_N128, ks != 0);
                        tw_umma_ts(tmem_base + C_D3, tmem_base + C_HHI + ks * 8, blo, TW_IDESC_N128, 1u);
                        tw_umma_ts(tmem_base + C_D3, tmem_base + C_HHI + ks * 8, bhi, TW_IDESC_N128, 1u);
                    }
                }
                tw_commit(bar_mma);
            }
            tw_mbar_wait(bar_mma, mma_uses & 1u);
            ++mma_uses;
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            TW_STAMP(2 + r * 8 + 5);
            {   // epilogue 3: x = lrelu(conv2 + b2 + x), next block's A operand
                float d[64];
#pragma unroll
                for (int c = 0; c < 64; c += 16) tw_ld16(tlane + C_D3 + half * 64 + c, d + c);
                tw_wait_ld();
#pragma unroll
                for (int c = 0; c < 64; c += 16) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {  // padded rows (p >= 81) carry harmless garbage; they are never stored
                        const float4 b = *reinterpret_cast<const float4 *>(bp + PI_B2 + half * 64 + c + i);
                        const float2 t01 = __fadd2_rn(__fadd2_rn(make_float2(d[c + i + 0], d[c + i + 1]), make_float2(b.x, b.y)),
                                                      make_float2(x[c + i + 0], x[c + i + 1]));
                        const float2 t23 = __fadd2_rn(__fadd2_rn(make_float2(d[c + i + 2], d[c + i + 3]), make_float2(b.z, b.w)),
                                                      make_float2(x[c + i + 2], x[c + i + 3]));
                        const float2 u01 = __fmul2_rn(t01, make_float2(0.2f, 0.2f)), u23 = __fmul2_rn(t23, make_float2(0.2f, 0.2f));
                        x[c + i + 0] = fmaxf(t01.x, u01.x);
                        x[c + i + 1] = fmaxf(t01.y, u01.y);
                        x[c + i + 2] = fmaxf(t23.x, u23.x);
                        x[c + i + 3] = fmaxf(t23.y, u23.y);
                    }
                    if (r < 2) tw_store_split16(tlane + C_XHI + half * 64 + c, tlane + C_XLO + half * 64 + c, x + c);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            TW_STAMP(2 + r * 8 + 6);
        }
        // ---- flatten NHWC (network.rs:127-137): this thread's pixel row, its 64 channels ----
        TW_STAMP(30);
        {   // the weight buffer of the block just finished is idle until the next conv0 phase: staging tiles live there
            float *stage = reinterpret_cast<float *>(sm + SM_W + ((g - 1u) & 1u) * WI_BYTES) + warp * (32 * TW_HSTRIDE);
            const long long off = real ? (long long)row * 10368 + (long long)p * 128 + half * 64 : -1ll;
            tw_store_out(stage, lane, off, x, act0, act0_hi, act0_lo);
        }
        TW_STAMP(31);
    }
    // (timestamps: 0 start, 1 stem done, then per block: conv0 done, E1+sync, dw stored, conv1 done, E2 stored, conv2 done, E3 stored)
    // drain the weight prefetch that was issued one block ahead, then release TMEM
    if (t == 0) tw_mbar_wait(bar_w0 + 8 * (g & 1u), (g >> 1) & 1u);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TW_TMEM_COLS) : "memory");
    }
}


// Build the per-block B-operand images (K-major SWIZZLE_128B, hi/lo) and the fp32 parameter image.
__device__ __forceinline__ uint32_t swz128(int n, int k) {  // byte offset of element (row n, k in [0,32)) inside an atom
    return (uint32_t)((n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 2) ^ (n & 7)) & 7) << 4) + (k & 3) * 4);
}
struct TowerPackArgs {
    const float *conv_w, *conv_b;
    const float *w0[3], *b0[3], *dw[3], *pw[3], *b1[3], *w2[3], *b2[3];
};
__global__ void k_tower_pack(TowerPackArgs a, uint8_t *wimg, float *pimg) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nth = gridDim.x * blockDim.x;
    for (int r = 0; r < 3; ++r) {
        uint8_t *img = wimg + (size_t)r * WI_BYTES;
        for (int i = tid; i < 128 * 32; i += nth) {  // W0[k][n]: k = cin 0..127, n = cout 0..31
            const int k = i >> 5, n = i & 31;
            const float v = a.w0[r][k * 32 + n];
            const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
            const uint32_t off = (uint32_t)(k >> 5) * 4096u + swz128(n, k & 31);
            *reinterpret_cast<float *>(img + WI_W0HI + off) = h;
            *reinterpret_cast<float *>(img + WI_W0LO + off) = v - h;
        }
        for (int i = tid; i < 32 * 32; i += nth) {  // PW[k][n]
            const int k = i >> 5, n = i & 31;
            const float v = a.pw[r][k * 32 + n];
            const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
            *reinterpret_cast<float *>(img + WI_PWHI + swz128(n, k)) = h;
            *reinterpret_cast<float *>(img + WI_PWLO + swz128(n, k)) = v - h;
        }
        for (int i = tid; i < 32 * 128; i += nth) {  // W2[k][n]: k = cin 0..31, n = cout 0..127
            const int k = i >> 7, n = i & 127;
            const float v = a.w2[r][k * 128 + n];
            const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
            *reinterpret_cast<float *>(img + WI_W2HI + swz128(n, k)) = h;
            *reinterpret_cast<float *>(img + WI_W2LO + swz128(n, k)) = v - h;
        }
        float *pb = pimg + PI_BLK0 + r * PI_BLK;
        for (int i = tid; i < 32; i += nth) {
            pb[PI_B0 + i] = a.b0[r][i];
            pb[PI_B1 + i] = a.b1[r][i];
        }
        for (int i = tid; i < 288; i += nth) pb[PI_DW + i] = a.dw[r][i];
        for (int i = tid; i < 128; i += nth) pb[PI_B2 + i] = a.b2[r][i];
    }
    for (int i = tid; i < 384; i += nth) pimg[PI_WSTEM + i] = a.conv_w[i];
    for (int i = tid; i < 128; i += nth) pimg[PI_BSTEM + i] = a.conv_b[i];
}

bool tower_tc_prepare_weights(omk_ctx *c) {
    if (!c->net.tower_wimg) {
        if (cudaMalloc(&c->net.tower_wimg, 3 * WI_BYTES) != cudaSuccess) return false;
        if (cudaMalloc(&c->net.tower_pimg, sizeof(float) * PI_FLOATS) != cudaSuccess) return false;
    }
    TowerPackArgs a;
    a.conv_w = c->net.t[0];
    a.conv_b = c->net.t[1];
    for (int r = 0; r < 3; ++r) {
        const int b = 2 + 7 * r;
        a.w0[r] = c->net.t[b + 0]; a.b0[r] = c->net.t[b + 1]; a.dw[r] = c->net.t[b + 2]; a.pw[r] = c->net.t[b + 3];
        a.b1[r] = c->net.t[b + 4]; a.w2[r] = c->net.t[b + 5]; a.b2[r] = c->net.t[b + 6];
    }
    k_tower_pack<<<16, 256, 0, c->stream>>>(a, c->net.tower_wimg, c->net.tower_pimg);
    c->launches++;
    return true;
}

void tower_tc_read_timing(long long *out64) { cudaMemcpyFromSymbol(out64, g_tw_dbg, sizeof(long long) * 64); }

void launch_tower_tc(omk_ctx *c, const float *images_dev, int rows_bound, bool split_out) {
    if (c->tower_pair) {  // CTA pairs walking position triples (243 of 256 TMEM lanes busy)
        cudaFuncSetAttribute(k_tower_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, TW3_SMEM_BYTES);
        const int triples = (rows_bound + 2) / 3;
        const int pairs = triples < c->n_sms / 2 ? triples : c->n_sms / 2;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(TW_THREADS);
        cfg.dynamicSmemBytes = TW3_SMEM_BYTES;
        cfg.stream = c->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const uint8_t *wimg = c->net.tower_wimg;
        const float *pimg = c->net.tower_pimg;
        const NNIn *nn_in = c->ws.nn_in;
        const uint32_t *n_req = c->ws.n_req;
        float *a0 = c->ws.act0, *ah = split_out ? c->ws.act0_hi : nullptr, *al = split_out ? c->ws.act0_lo : nullptr;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, k_tower_tc3, wimg, pimg, nn_in, images_dev, n_req, rows_bound, a0, ah, al);
        if (e != cudaSuccess) fprintf(stderr, "omok_b200: cudaLaunchKernelEx(k_tower_tc3): %s\n", cudaGetErrorString(e));
        c->launches++;
        return;
    }
    cudaFuncSetAttribute(k_tower_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TW_SMEM_BYTES);
    const int grid = rows_bound < c->n_sms ? rows_bound : c->n_sms;
    k_tower_tc<<<grid, TW_THREADS, TW_SMEM_BYTES, c->stream>>>(c->net.tower_wimg, c->net.tower_pimg, c->ws.nn_in, images_dev,
                                                                c->ws.n_req, rows_bound, c->ws.act0,
                                                                split_out ? c->ws.act0_hi : nullptr, split_out ? c->ws.act0_lo : nullptr);
    c->launches++;
}

}  // namespace omk

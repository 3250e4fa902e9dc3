// fc0_tc.cu -- K3 tensor-core path for fc0 (10368 -> 512, 67 % of the network's flops).
//
//   C[M x 512] = lrelu( A[M x 10368] . W[10368 x 512] + b )        (alpha-zero/src/network.rs:139-150)
//
// fp32-class accuracy on the 5th-gen tensor cores: every operand is split into a TF32-exact
// high part and a residual low part (x = hi + lo, hi = x with the 13 low mantissa bits cleared)
// and each k-step issues three tcgen05.mma kind::tf32 into one TMEM accumulator:
//        hi.hi + hi.lo + lo.hi        (the dropped lo.lo term is ~2^-22 relative)
// which holds the north-star 1e-3 tolerance where one-pass TF32 misses it by 50x.
//
// The tensor-core accumulator truncates (measured: 86 % of outputs biased toward zero, 6e-5 relative
// after 3 x 1296 accumulations), so K is accumulated in TMEM only over short chunks of TC_CHUNK
// k-blocks; each chunk is drained into fp32 registers with round-to-nearest adds ("promotion") while
// the tensor cores fill the other TMEM buffer.
//
// Structure (one CTA per 128 x 256 output tile, 10 warps):
//   warp 0     TMA producer: four 2-D tiled loads per k-block (A_hi, A_lo, W_hi, W_lo; 128-byte
//              rows, SWIZZLE_128B) into a 2-stage ring, completion on an mbarrier (expect_tx)
//   warp 1     TMEM allocator (512 columns = 2 accumulator buffers) + single-thread MMA issuer:
//              4 k-steps x 3 products of 128 x 256 x 8 per k-block; tcgen05.commit releases the
//              smem stage, and at each chunk end signals the drain warps
//   warps 2-9  drain / epilogue: two warps per TMEM lane quadrant (128 columns each): tcgen05.ld
//              32x32b, accumulate in registers; at the end + bias, lrelu, store
// Both operands are K-major: A rows are positions, and W is transposed once at load time.
#include <cstdio>
#include <cuda.h>

#include "omk_internal.h"

namespace omk {

constexpr int TC_BM = 128, TC_BN = 256, TC_BK = 32, TC_STAGES = 2;
constexpr int TC_K = 10368, TC_N = 512;   // fc0
constexpr int TC_K1 = 512;                // fc1 (512 -> 512) runs through the same kernel template
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;  // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;  // 32 KB
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;  // 96 KB
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*barriers*/ + 1024 /*alignment slack*/;
constexpr int TC_THREADS = 320;
constexpr uint32_t TC_TMEM_COLS = 512;
constexpr int TC_CHUNK0 = 9;   // fc0: k-blocks accumulated in TMEM before promotion (324 = 36 x 9)
constexpr int TC_CHUNK1 = 8;   // fc1: 16 = 2 x 8
// instruction descriptor: D=F32 (bit 4), A=B=TF32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// spin on the phase parity; a bounded spin turns a pipeline bug into a trap instead of a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);       // start address, 16-byte units
    d |= (uint64_t)0 << 16;                        // leading byte offset: unused for swizzled K-major
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset between 8-row core groups
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K = inner dimension, CHUNK = k-blocks per TMEM accumulation chunk.  Output: C (fp32) and/or its TF32 hi/lo split
// (C_hi/C_lo, the A operand of the next tensor-core layer); any of the three may be null.
template <int K, int CHUNK>
__global__ void __launch_bounds__(TC_THREADS, 1)
    k_fc_tc(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
            const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
            const float *__restrict__ bias, float *__restrict__ C, float *__restrict__ C_hi, float *__restrict__ C_lo,
            const uint32_t *n_req, int max_rows) {
    constexpr int TC_NKB = K / TC_BK;
    constexpr int TC_CHUNK = CHUNK;
    constexpr int TC_NCHUNK = TC_NKB / TC_CHUNK;
    static_assert(TC_NKB % TC_CHUNK == 0, "chunking must tile K");
    extern __shared__ uint8_t smem_raw[];
    const int rows = (int)min(*n_req, (uint32_t)max_rows);
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;
    if (m0 >= rows) return;  // uniform for the whole CTA, before any barrier or TMEM use

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t bars = base + TC_STAGES * TC_STAGE_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * TC_STAGES, tmem_full0 = bars + 16 * TC_STAGES, tmem_empty0 = tmem_full0 + 16,
                   tmem_slot = tmem_empty0 + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo));
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tmem_full0 + 8 * b, 1);
            mbar_init(tmem_empty0 + 8 * b, 8);  // one arrival per drain warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            for (int kb = 0; kb < TC_NKB; ++kb) {
                const int s = kb % TC_STAGES;
                const uint32_t ph = (uint32_t)(kb / TC_STAGES) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                const uint32_t st = base + s * TC_STAGE_BYTES;
                const uint32_t bar = full0 + 8 * s;
                mbar_expect_tx(bar, TC_STAGE_BYTES);
                tma_load_2d(st, &map_a_hi, bar, kb * TC_BK, m0);
                tma_load_2d(st + TC_A_BYTES, &map_a_lo, bar, kb * TC_BK, m0);
                tma_load_2d(st + 2 * TC_A_BYTES, &map_b_hi, bar, kb * TC_BK, n0);
                tma_load_2d(st + 2 * TC_A_BYTES + TC_B_BYTES, &map_b_lo, bar, kb * TC_BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            for (int ch = 0; ch < TC_NCHUNK; ++ch) {
                const int buf = ch & 1;
                const uint32_t use = (uint32_t)(ch >> 1);
                mbar_wait(tmem_empty0 + 8 * buf, (use & 1u) ^ 1u);  // drained (passes at once for the first use)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * TC_BN);
                for (int kc = 0; kc < TC_CHUNK; ++kc) {
                    const int kb = ch * TC_CHUNK + kc;
                    const int s = kb % TC_STAGES;
                    const uint32_t ph = (uint32_t)(kb / TC_STAGES) & 1u;
                    mbar_wait(full0 + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * TC_STAGE_BYTES;
                    const uint64_t a_hi = make_desc_sw128(st), a_lo = make_desc_sw128(st + TC_A_BYTES);
                    const uint64_t b_hi = make_desc_sw128(st + 2 * TC_A_BYTES), b_lo = make_desc_sw128(st + 2 * TC_A_BYTES + TC_B_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);  // 32 bytes per k-step inside the swizzle atom
                        umma_tf32(tmem_acc, a_lo + adv, b_hi + adv, TC_IDESC, (kc | k) != 0 ? 1u : 0u);
                        umma_tf32(tmem_acc, a_hi + adv, b_lo + adv, TC_IDESC, 1u);
                        umma_tf32(tmem_acc, a_hi + adv, b_hi + adv, TC_IDESC, 1u);
                    }
                    umma_commit(empty0 + 8 * s);  // stage reusable once these MMAs have read it
                }
                umma_commit(tmem_full0 + 8 * buf);  // chunk accumulator complete
            }
        }
    } else {  // ===== drain + epilogue: warps 2..9; quadrant = warp % 4, column half = (warp - 2) / 4 =====
        const int q = warp & 3, half = (warp - 2) >> 2;
        float acc[128];
#pragma unroll
        for (int j = 0; j < 128; ++j) acc[j] = 0.0f;
        for (int ch = 0; ch < TC_NCHUNK; ++ch) {
            const int buf = ch & 1;
            const uint32_t use = (uint32_t)(ch >> 1);
            mbar_wait(tmem_full0 + 8 * buf, use & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 128; c += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC_BN + half * 128 + c);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[c + j] += __uint_as_float(v[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty0 + 8 * buf) : "memory");
        }
        const int row = m0 + q * 32 + lane;
        const size_t coff = (size_t)row * TC_N + n0 + half * 128;
        const float *brow = bias + n0 + half * 128;
#pragma unroll
        for (int j = 0; j < 128; j += 4) {
            const float4 b = *reinterpret_cast<const float4 *>(brow + j);
            float4 o;
            o.x = acc[j + 0] + b.x;
            o.y = acc[j + 1] + b.y;
            o.z = acc[j + 2] + b.z;
            o.w = acc[j + 3] + b.w;
            o.x = fmaxf(o.x, 0.2f * o.x);
            o.y = fmaxf(o.y, 0.2f * o.y);
            o.z = fmaxf(o.z, 0.2f * o.z);
            o.w = fmaxf(o.w, 0.2f * o.w);
            if (row >= rows) continue;  // rows past the batch are never stored (the tile may extend past the workspace)
            if (C) *reinterpret_cast<float4 *>(C + coff + j) = o;
            if (C_hi) {
                float4 h, l;
                h.x = __uint_as_float(__float_as_uint(o.x) & 0xFFFFE000u); l.x = o.x - h.x;
                h.y = __uint_as_float(__float_as_uint(o.y) & 0xFFFFE000u); l.y = o.y - h.y;
                h.z = __uint_as_float(__float_as_uint(o.z) & 0xFFFFE000u); l.z = o.z - h.z;
                h.w = __uint_as_float(__float_as_uint(o.w) & 0xFFFFE000u); l.w = o.w - h.w;
                *reinterpret_cast<float4 *>(C_hi + coff + j) = h;
                *reinterpret_cast<float4 *>(C_lo + coff + j) = l;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// cta_group::2 variant: a CTA PAIR (cluster 2x1) computes a 256 x 256 tile.  Each CTA loads its own 128 rows of
// A and HALF of the B tile (128 of the 256 weight rows); tcgen05.mma.cta_group::2 reads both CTAs' shared memory,
// so the L2 -> SMEM operand traffic per MMA drops by a third (64 KB instead of 96 KB per CTA per k-block) -- the
// 1-CTA kernel is bound by exactly that traffic (8.15 GB per launch, ~9.7 TB/s).  3-stage ring of 64 KB stages.
// The leader CTA (cluster rank 0) owns the full barriers and issues the MMAs; tcgen05.commit multicasts the
// "stage free" / "accumulator ready" arrivals to both CTAs; both CTAs' drain warps arrive on the leader's
// "accumulator drained" barrier through the cluster shared window.
// ---------------------------------------------------------------------------------------------------
constexpr int T2_STAGES = 3;
constexpr int T2_BH_BYTES = 128 * TC_BK * 4;                       // half of the B tile: 16 KB
constexpr int T2_STAGE_BYTES = 2 * TC_A_BYTES + 2 * T2_BH_BYTES;    // 64 KB per CTA
constexpr int T2_SMEM_BYTES = T2_STAGES * T2_STAGE_BYTES + 1024 + 1024;
constexpr uint32_t T2_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t mapa_rank0(uint32_t addr) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {  // arrive on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int K, int CHUNK>
__global__ void __launch_bounds__(TC_THREADS, 1)
    k_fc_tc2(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
             const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
             const float *__restrict__ bias, float *__restrict__ C, float *__restrict__ C_hi, float *__restrict__ C_lo,
             const uint32_t *n_req, int max_rows) {
    constexpr int NKB = K / TC_BK;
    constexpr int NCHUNK = NKB / CHUNK;
    static_assert(NKB % CHUNK == 0, "chunking must tile K");
    extern __shared__ uint8_t smem_raw[];
    const int rows = (int)min(*n_req, (uint32_t)max_rows);
    // blockIdx.x = ((pair * 2 + n_tile) * 2 + rank): the CTA pair is a 2x1 cluster along x, and the two N tiles of one
    // row pair are neighbours in launch order so that the second reads its A rows from L2, not from DRAM
    const int pair = blockIdx.x >> 2;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int m0 = pair * 256 + (int)rank * TC_BM, n0 = (int)((blockIdx.x >> 1) & 1u) * TC_BN;
    if (pair * 256 >= rows) return;  // uniform for the whole cluster

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + T2_STAGES * T2_STAGE_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * T2_STAGES, tmem_full0 = bars + 16 * T2_STAGES, tmem_empty0 = tmem_full0 + 16,
                   tmem_slot = tmem_empty0 + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo));
        for (int s = 0; s < T2_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tmem_full0 + 8 * b, 1);
            mbar_init(tmem_empty0 + 8 * b, 16);  // 8 drain warps in each CTA of the pair
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers are initialised before anything crosses the pair
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer (both CTAs; transaction bytes land on the leader's barrier) =====
            for (int kb = 0; kb < NKB; ++kb) {
                const int s = kb % T2_STAGES;
                const uint32_t ph = (uint32_t)(kb / T2_STAGES) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                const uint32_t st = base + s * T2_STAGE_BYTES;
                const uint32_t lbar = mapa_rank0(full0 + 8 * s);
                if (leader) mbar_expect_tx(full0 + 8 * s, 2 * T2_STAGE_BYTES);
                tma_load_2d_2sm(st, &map_a_hi, lbar, kb * TC_BK, m0);
                tma_load_2d_2sm(st + TC_A_BYTES, &map_a_lo, lbar, kb * TC_BK, m0);
                tma_load_2d_2sm(st + 2 * TC_A_BYTES, &map_b_hi, lbar, kb * TC_BK, n0 + (int)rank * 128);
                tma_load_2d_2sm(st + 2 * TC_A_BYTES + T2_BH_BYTES, &map_b_lo, lbar, kb * TC_BK, n0 + (int)rank * 128);
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {  // ===== MMA issuer (leader CTA only) =====
            for (int ch = 0; ch < NCHUNK; ++ch) {
                const int buf = ch & 1;
                const uint32_t use = (uint32_t)(ch >> 1);
                mbar_wait(tmem_empty0 + 8 * buf, (use & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * TC_BN);
                for (int kc = 0; kc < CHUNK; ++kc) {
                    const int kb = ch * CHUNK + kc;
                    const int s = kb % T2_STAGES;
                    const uint32_t ph = (uint32_t)(kb / T2_STAGES) & 1u;
                    mbar_wait(full0 + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * T2_STAGE_BYTES;
                    const uint64_t a_hi = make_desc_sw128(st), a_lo = make_desc_sw128(st + TC_A_BYTES);
                    const uint64_t b_hi = make_desc_sw128(st + 2 * TC_A_BYTES), b_lo = make_desc_sw128(st + 2 * TC_A_BYTES + T2_BH_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
                        umma_tf32_2sm(tmem_acc, a_lo + adv, b_hi + adv, T2_IDESC, (kc | k) != 0 ? 1u : 0u);
                        umma_tf32_2sm(tmem_acc, a_hi + adv, b_lo + adv, T2_IDESC, 1u);
                        umma_tf32_2sm(tmem_acc, a_hi + adv, b_hi + adv, T2_IDESC, 1u);
                    }
                    umma_commit_2sm(empty0 + 8 * s);
                }
                umma_commit_2sm(tmem_full0 + 8 * buf);
            }
        }
    } else {  // ===== drain + epilogue (both CTAs, own 128 rows) =====
        const int q = warp & 3, half = (warp - 2) >> 2;
        float acc[128];
#pragma unroll
        for (int j = 0; j < 128; ++j) acc[j] = 0.0f;
        for (int ch = 0; ch < NCHUNK; ++ch) {
            const int buf = ch & 1;
            const uint32_t use = (uint32_t)(ch >> 1);
            mbar_wait(tmem_full0 + 8 * buf, use & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 128; c += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC_BN + half * 128 + c);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[c + j] += __uint_as_float(v[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const uint32_t rbar = mapa_rank0(tmem_empty0 + 8 * buf);
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
            }
        }
        const int row = m0 + q * 32 + lane;
        const size_t coff = (size_t)row * TC_N + n0 + half * 128;
        const float *brow = bias + n0 + half * 128;
#pragma unroll
        for (int j = 0; j < 128; j += 4) {
            const float4 b = *reinterpret_cast<const float4 *>(brow + j);
            float4 o;
            o.x = acc[j + 0] + b.x;
            o.y = acc[j + 1] + b.y;
            o.z = acc[j + 2] + b.z;
            o.w = acc[j + 3] + b.w;
            o.x = fmaxf(o.x, 0.2f * o.x);
            o.y = fmaxf(o.y, 0.2f * o.y);
            o.z = fmaxf(o.z, 0.2f * o.z);
            o.w = fmaxf(o.w, 0.2f * o.w);
            if (row >= rows) continue;  // rows past the batch are never stored (the tile may extend past the workspace)
            if (C) *reinterpret_cast<float4 *>(C + coff + j) = o;
            if (C_hi) {
                float4 h, l;
                h.x = __uint_as_float(__float_as_uint(o.x) & 0xFFFFE000u); l.x = o.x - h.x;
                h.y = __uint_as_float(__float_as_uint(o.y) & 0xFFFFE000u); l.y = o.y - h.y;
                h.z = __uint_as_float(__float_as_uint(o.z) & 0xFFFFE000u); l.z = o.z - h.z;
                h.w = __uint_as_float(__float_as_uint(o.w) & 0xFFFFE000u); l.w = o.w - h.w;
                *reinterpret_cast<float4 *>(C_hi + coff + j) = h;
                *reinterpret_cast<float4 *>(C_lo + coff + j) = l;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();  // the peer's shared memory and TMEM stay alive until every MMA and drain has finished
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

// W[10368][512] -> K-major transposed hi / lo parts Wt[512][10368]
__global__ void k_fc0_split_weights(const float *__restrict__ W, float *__restrict__ hi, float *__restrict__ lo, int K) {
    __shared__ float tile[32][33];
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) tile[r][tx] = W[(size_t)(k0 + r) * TC_N + n0 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const float x = tile[tx][r];  // W[k0+tx][n0+r]
        const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        hi[(size_t)(n0 + r) * K + k0 + tx] = h;
        lo[(size_t)(n0 + r) * K + k0 + tx] = x - h;
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool encode_map(CUtensorMap *map, float *ptr, uint64_t rows, uint32_t box_rows, int K = TC_K) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
        fn = (PFN_encodeTiled)p;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct Fc0TcState {
    CUtensorMap map_a_hi, map_a_lo, map_b_hi, map_b_lo;      // fc0
    CUtensorMap map1_a_hi, map1_a_lo, map1_b_hi, map1_b_lo;  // fc1
    CUtensorMap map2_b_hi, map2_b_lo;                        // fc0 weights with 128-row boxes (cta_group::2 kernel)
    float *a_hi_ptr = nullptr;
    int a_rows = 0;
    bool weights_ready = false;
};

static Fc0TcState *state_of(omk_ctx *c) {
    if (!c->fc0_tc_state) c->fc0_tc_state = new Fc0TcState();
    return reinterpret_cast<Fc0TcState *>(c->fc0_tc_state);
}

void fc0_tc_free(omk_ctx *c) {
    delete reinterpret_cast<Fc0TcState *>(c->fc0_tc_state);
    c->fc0_tc_state = nullptr;
}

// (re)build the split transposed weights of fc0 and fc1; call after the weights change
bool fc0_tc_prepare_weights(omk_ctx *c) {
    Fc0TcState *s = state_of(c);
    if (!c->net.fc0_wt_hi) {
        if (cudaMalloc(&c->net.fc0_wt_hi, sizeof(float) * (size_t)TC_K * TC_N) != cudaSuccess) return false;
        if (cudaMalloc(&c->net.fc0_wt_lo, sizeof(float) * (size_t)TC_K * TC_N) != cudaSuccess) return false;
        if (cudaMalloc(&c->net.fc1_wt_hi, sizeof(float) * (size_t)TC_K1 * TC_N) != cudaSuccess) return false;
        if (cudaMalloc(&c->net.fc1_wt_lo, sizeof(float) * (size_t)TC_K1 * TC_N) != cudaSuccess) return false;
    }
    k_fc0_split_weights<<<dim3(TC_K / 32, TC_N / 32), 256, 0, c->stream>>>(c->net.t[23], c->net.fc0_wt_hi, c->net.fc0_wt_lo, TC_K);
    k_fc0_split_weights<<<dim3(TC_K1 / 32, TC_N / 32), 256, 0, c->stream>>>(c->net.t[25], c->net.fc1_wt_hi, c->net.fc1_wt_lo, TC_K1);
    c->launches += 2;
    if (!encode_map(&s->map_b_hi, c->net.fc0_wt_hi, TC_N, TC_BN)) return false;
    if (!encode_map(&s->map_b_lo, c->net.fc0_wt_lo, TC_N, TC_BN)) return false;
    if (!encode_map(&s->map2_b_hi, c->net.fc0_wt_hi, TC_N, 128)) return false;
    if (!encode_map(&s->map2_b_lo, c->net.fc0_wt_lo, TC_N, 128)) return false;
    if (!encode_map(&s->map1_b_hi, c->net.fc1_wt_hi, TC_N, TC_BN, TC_K1)) return false;
    if (!encode_map(&s->map1_b_lo, c->net.fc1_wt_lo, TC_N, TC_BN, TC_K1)) return false;
    s->weights_ready = true;
    return true;
}

static bool refresh_activation_maps(omk_ctx *c, Fc0TcState *s) {
    if (s->a_hi_ptr == c->ws.act0_hi && s->a_rows == c->ws.max_rows) return true;
    if (!encode_map(&s->map_a_hi, c->ws.act0_hi, (uint64_t)c->ws.max_rows, TC_BM)) return false;
    if (!encode_map(&s->map_a_lo, c->ws.act0_lo, (uint64_t)c->ws.max_rows, TC_BM)) return false;
    if (!encode_map(&s->map1_a_hi, c->ws.act1_hi, (uint64_t)c->ws.max_rows, TC_BM, TC_K1)) return false;
    if (!encode_map(&s->map1_a_lo, c->ws.act1_lo, (uint64_t)c->ws.max_rows, TC_BM, TC_K1)) return false;
    s->a_hi_ptr = c->ws.act0_hi;
    s->a_rows = c->ws.max_rows;
    return true;
}

// fc0: act0_hi/lo -> act1 (fp32, for the CUDA-core fc1) and/or act1_hi/lo (for the tensor-core fc1)
bool launch_fc0_tc(omk_ctx *c, int rows_bound, bool split_out) {
    Fc0TcState *s = state_of(c);
    if (!s->weights_ready || !refresh_activation_maps(c, s)) return false;
    const int mt = (rows_bound + TC_BM - 1) / TC_BM;
    float *c_f32 = split_out ? nullptr : c->ws.act1, *c_hi = split_out ? c->ws.act1_hi : nullptr, *c_lo = split_out ? c->ws.act1_lo : nullptr;
    if (c->fc0_pair) {  // cta_group::2: CTA pairs along M
        auto kern = k_fc_tc2<TC_K, TC_CHUNK0>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES);
        const int pairs = (rows_bound + 255) / 256;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * (TC_N / TC_BN) * pairs, 1);
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = T2_SMEM_BYTES;
        cfg.stream = c->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const float *bias_p = c->net.t[24];
        const uint32_t *nreq_p = c->ws.n_req;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, s->map_a_hi, s->map_a_lo, s->map2_b_hi, s->map2_b_lo, bias_p, c_f32, c_hi,
                                                 c_lo, nreq_p, rows_bound);
        if (e != cudaSuccess) fprintf(stderr, "omok_b200: cudaLaunchKernelEx(k_fc_tc2): %s\n", cudaGetErrorString(e));
    } else {
        auto kern = k_fc_tc<TC_K, TC_CHUNK0>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
        kern<<<dim3(TC_N / TC_BN, mt), TC_THREADS, TC_SMEM_BYTES, c->stream>>>(
            s->map_a_hi, s->map_a_lo, s->map_b_hi, s->map_b_lo, c->net.t[24], c_f32, c_hi, c_lo, c->ws.n_req, rows_bound);
    }
    c->launches++;
    const cudaError_t le = cudaPeekAtLastError();
    if (le != cudaSuccess) {
        fprintf(stderr, "omok_b200: fc0 tensor-core launch failed: %s\n", cudaGetErrorString(le));
        return false;
    }
    return true;
}

// fc1: act1_hi/lo -> act2 (fp32)
bool launch_fc1_tc(omk_ctx *c, int rows_bound) {
    Fc0TcState *s = state_of(c);
    if (!s->weights_ready || !refresh_activation_maps(c, s)) return false;
    auto kern = k_fc_tc<TC_K1, TC_CHUNK1>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    const int mt = (rows_bound + TC_BM - 1) / TC_BM;
    kern<<<dim3(TC_N / TC_BN, mt), TC_THREADS, TC_SMEM_BYTES, c->stream>>>(
        s->map1_a_hi, s->map1_a_lo, s->map1_b_hi, s->map1_b_lo, c->net.t[26], c->ws.act2, nullptr, nullptr, c->ws.n_req, rows_bound);
    c->launches++;
    return true;
}

}  // namespace omk

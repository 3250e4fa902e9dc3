// env_kernels.cu -- K1: batched Gomoku environment on packed bitboards.
//
// Replaces environment::Environment::{new, place_stone, encode_board}
// (environment/src/lib.rs:73-193) and alpha_zero::encode_nn_input (encoder.rs:10-46).
//
// The standalone step is HBM-bound integer work: 78 algorithmic bytes per board-step
// (32 R + 32 W record, 1 R action, 1 W status, 12 W legal mask).  At that intensity a
// warp-per-board mapping cannot reach the roofline (<= 28 warp instructions per board-step
// at 50 % of HBM), so this kernel maps ONE LANE to one board: a warp moves 32 consecutive
// 32-byte records with 128-bit accesses and each lane runs the exact-five line scan on its
// own registers.  (The warp-cooperative ballot line scan lives in tree_kernels.cu, where a
// warp owns a whole tree.)
#include "omk_internal.h"

namespace omk {

__device__ __forceinline__ EnvRec load_env(const EnvRec *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    const uint4 a = q[0], b = q[1];
    EnvRec r;
    r.black[0] = a.x; r.black[1] = a.y; r.black[2] = a.z;
    r.white[0] = a.w; r.white[1] = b.x; r.white[2] = b.y;
    r.meta = b.z; r.pad = b.w;
    return r;
}
__device__ __forceinline__ void store_env(EnvRec *p, const EnvRec &r) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(r.black[0], r.black[1], r.black[2], r.white[0]);
    q[1] = make_uint4(r.white[1], r.white[2], r.meta, r.pad);
}
__device__ __forceinline__ EnvRec fresh_env() {  // Environment::new (:73-79)
    EnvRec r{};
    r.meta = 0u | ((uint32_t)kCells << 8);  // Black to move, 81 legal moves
    return r;
}

// place_stone (:104-166) on a record in registers.  Returns OMK_NONE without mutation when the
// cell is occupied; there is no terminal guard, like the reference.
__device__ __forceinline__ int place_stone(EnvRec &r, int idx) {
    const uint32_t bit = 1u << (idx & 31);
    const int word = idx >> 5;
    const uint32_t occ = sel3(r.black[0] | r.white[0], r.black[1] | r.white[1], r.black[2] | r.white[2], word);
    if (occ & bit) return OMK_NONE;
    const uint32_t turn = r.meta & 1u;
    uint32_t legal = (r.meta >> 8) & 0xFFu;
    legal -= 1;
    uint32_t *mine = turn == 0 ? r.black : r.white;
    if (word == 0) mine[0] |= bit; else if (word == 1) mine[1] |= bit; else mine[2] |= bit;
    const bool five = makes_five_scalar(mine, idx);
    r.meta = (turn ^ 1u) | (legal << 8);
    if (five) return turn == 0 ? kBlackWin : kWhiteWin;  // the win test precedes the draw test
    return legal == 0 ? kDraw : kInProgress;
}

__global__ void k_env_reset(EnvRec *envs, const int32_t *ids, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store_env(&envs[ids ? ids[i] : i], fresh_env());
}

__global__ void __launch_bounds__(256) k_env_step(EnvRec *envs, const int32_t *ids, const uint8_t *actions, int n,
                                                   int8_t *out_status, uint32_t *out_legal) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvRec *p = &envs[ids ? ids[i] : i];
    EnvRec r = load_env(p);
    // an index >= 81 panics in the reference (slice bounds); here it is reported as None, without mutation
    const int st = actions[i] < kCells ? place_stone(r, actions[i]) : OMK_NONE;
    if (st != OMK_NONE) store_env(p, r);
    if (out_status) out_status[i] = (int8_t)st;
    if (out_legal) {
        out_legal[3 * (size_t)i + 0] = ~(r.black[0] | r.white[0]);
        out_legal[3 * (size_t)i + 1] = ~(r.black[1] | r.white[1]);
        out_legal[3 * (size_t)i + 2] = ~(r.black[2] | r.white[2]) & 0x1FFFFu;
    }
}

__global__ void k_env_get(const EnvRec *envs, const int32_t *ids, int n, uint8_t *boards, uint8_t *turns,
                          uint16_t *legal) {
    const int i = blockIdx.x;
    if (i >= n) return;
    const EnvRec r = load_env(&envs[ids ? ids[i] : i]);
    for (int c = threadIdx.x; c < kCells; c += blockDim.x) {
        if (boards) boards[(size_t)i * kCells + c] = bit81(r.black, c) ? 1 : (bit81(r.white, c) ? 2 : 0);
    }
    if (threadIdx.x == 0) {
        if (turns) turns[i] = (uint8_t)(r.meta & 1u);
        if (legal) legal[i] = (uint16_t)((r.meta >> 8) & 0xFFu);
    }
}

__global__ void k_env_set(EnvRec *envs, const int32_t *ids, int n, const uint8_t *boards, const uint8_t *turns) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvRec r{};
    int stones = 0;
    for (int c = 0; c < kCells; ++c) {
        const uint8_t s = boards[(size_t)i * kCells + c];
        if (s == 1) { set81(r.black, c); ++stones; }
        if (s == 2) { set81(r.white, c); ++stones; }
    }
    r.meta = (uint32_t)(turns[i] & 1u) | ((uint32_t)(kCells - stones) << 8);
    store_env(&envs[ids ? ids[i] : i], r);
}

__global__ void k_env_encode(const EnvRec *envs, const int32_t *ids, int n, int mode, float *out) {
    const int i = blockIdx.x;
    if (i >= n) return;
    const EnvRec r = load_env(&envs[ids ? ids[i] : i]);
    for (int f = threadIdx.x; f < 243; f += blockDim.x)
        out[(size_t)i * 243 + f] = image_value(r.black, r.white, r.meta & 1u, (uint32_t)mode & 1u, f);
}

// cells + turns (host API of omk_net_eval) -> evaluator request rows
__global__ void k_pack_boards(const uint8_t *boards, const uint8_t *turns, int n, int mode, NNIn *nn_in,
                              uint32_t *n_req) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *n_req = (uint32_t)n;
    if (i >= n) return;
    NNIn in{};
    for (int c = 0; c < kCells; ++c) {
        const uint8_t s = boards[(size_t)i * kCells + c];
        if (s == 1) set81(in.black, c);
        if (s == 2) set81(in.white, c);
    }
    in.meta = (uint32_t)(turns[i] & 1u) | (((uint32_t)mode & 1u) << 1);
    nn_in[i] = in;
}

// BASELINE config 2: persistent random playout, boards in registers across plies.
// Stream = env id, one bounded draw per ply over the ascending list of empty cells.
__global__ void __launch_bounds__(256) k_env_playout(EnvRec *envs, int n, int plies, uint64_t seed,
                                                      uint8_t *out_actions, int8_t *out_status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvRec r = fresh_env();
    uint32_t ctr = 0;
    for (int ply = 0; ply < plies; ++ply) {
        uint32_t empty[3] = {~(r.black[0] | r.white[0]), ~(r.black[1] | r.white[1]),
                             ~(r.black[2] | r.white[2]) & 0x1FFFFu};
        const int k = popc81(empty);
        const int action = nth_set81(empty, (int)rng_below(seed, (uint32_t)i, ctr, (uint32_t)k));
        const int st = place_stone(r, action);
        if (out_actions) out_actions[(size_t)ply * n + i] = (uint8_t)action;
        if (out_status) out_status[(size_t)ply * n + i] = (int8_t)st;
        if (st != kInProgress) r = fresh_env();  // auto-reset after a terminal status
    }
    store_env(&envs[i], r);
}

// ---- launchers ----
void launch_env_reset(omk_ctx *c, const int32_t *ids_dev, int n) {
    if (n <= 0) return;
    k_env_reset<<<(n + 255) / 256, 256, 0, c->stream>>>(c->envs, ids_dev, n);
    c->launches++;
}
void launch_env_step(omk_ctx *c, const int32_t *ids_dev, const uint8_t *actions_dev, int n, int8_t *status_dev,
                     uint32_t *legal_dev) {
    if (n <= 0) return;
    k_env_step<<<(n + 255) / 256, 256, 0, c->stream>>>(c->envs, ids_dev, actions_dev, n, status_dev, legal_dev);
    c->launches++;
}
void launch_env_get(omk_ctx *c, const int32_t *ids_dev, int n, uint8_t *boards_dev, uint8_t *turns_dev,
                    uint16_t *legal_dev) {
    if (n <= 0) return;
    k_env_get<<<n, 96, 0, c->stream>>>(c->envs, ids_dev, n, boards_dev, turns_dev, legal_dev);
    c->launches++;
}
void launch_env_set(omk_ctx *c, const int32_t *ids_dev, int n, const uint8_t *boards_dev, const uint8_t *turns_dev) {
    if (n <= 0) return;
    k_env_set<<<(n + 127) / 128, 128, 0, c->stream>>>(c->envs, ids_dev, n, boards_dev, turns_dev);
    c->launches++;
}
void launch_env_encode(omk_ctx *c, const int32_t *ids_dev, int n, int mode, float *out_dev) {
    if (n <= 0) return;
    k_env_encode<<<n, 128, 0, c->stream>>>(c->envs, ids_dev, n, mode, out_dev);
    c->launches++;
}
void launch_env_playout(omk_ctx *c, int n, int plies, uint8_t *actions_dev, int8_t *status_dev) {
    if (n <= 0) return;
    k_env_playout<<<(n + 255) / 256, 256, 0, c->stream>>>(c->envs, n, plies, c->seed, actions_dev, status_dev);
    c->launches++;
}
void launch_pack_boards(omk_ctx *c, const uint8_t *boards_dev, const uint8_t *turns_dev, int n, int mode) {
    if (n <= 0) return;
    k_pack_boards<<<(n + 127) / 128, 128, 0, c->stream>>>(boards_dev, turns_dev, n, mode, c->ws.nn_in, c->ws.n_req);
    c->launches++;
}

}  // namespace omk

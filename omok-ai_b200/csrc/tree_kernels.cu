// tree_kernels.cu -- K2: the tree-pool kernel set (one warp owns one tree).
//
// Replaces, for thousands of concurrent trees in HBM:
//   mcts::Node::{select_leaf, expand, propagate}           mcts/src/node.rs:39-99
//   mcts::MCTS::transition                                  mcts/src/lib.rs:47-78
//   ParallelMCTSExecutor::execute's per-agent body         alpha-zero/src/parallel_mcts_executor.rs:44-265
//   Agent::{new, compute_policy, sample_action,
//           ensure_action_exists, play_action}              alpha-zero/src/agent.rs:16-232
// Semantics are the REFERENCE's, not textbook AlphaZero: a leaf is any node that is not
// fully expanded, expansion adds ONE uniformly random unexpanded child, there is no
// virtual loss, and argmax ties go to the last-created child.
//
// Compiled with -fmad=false; every f32 operation that must match the CPU bit for bit is
// additionally written with an explicit round-to-nearest intrinsic.
#include "omk_internal.h"

namespace omk {

constexpr int kWarpsPerBlock = 4;
constexpr unsigned kFull = 0xffffffffu;

struct TreeArgs {
    TreeHdr *hdrs;
    uint8_t *nodes;
    int cap_nodes;
    const int32_t *ids;
    int id_base;  // identity id lists (ids == nullptr) start here: a search lane owns a contiguous slice of the pool
    int n;
    uint64_t seed;
    uint32_t *dev_error;
    unsigned long long *dev_sims;
    // request buffers
    NNIn *nn_in;
    uint32_t *req_tree, *req_node;
    uint32_t *n_req;
    uint32_t *slot_base, *slot_count;
    const float *P, *V;
    int max_rows;
    int vloss;  // OPT-IN, NOT the reference's semantics: virtual loss inside a round (omk_search_set_virtual_loss)
};

__device__ __forceinline__ int tree_of(const TreeArgs &a, int slot) { return a.ids ? a.ids[slot] : slot + a.id_base; }
__device__ __forceinline__ uint8_t *tree_nodes_of(const TreeArgs &a, int tree) {
    return a.nodes + (size_t)tree * (size_t)a.cap_nodes * kNodeBytes;
}

// value of the implicit dummy policy of a node that has not been evaluated yet:
// uniform over its empty cells (parallel_mcts_executor.rs:137-156)
__device__ __forceinline__ float dummy_prior(const NodeHdr &h, int a) {
    const uint32_t occ = sel3(h.black[0] | h.white[0], h.black[1] | h.white[1], h.black[2] | h.white[2], a >> 5);
    if ((occ >> (a & 31)) & 1u) return 0.0f;
    if (h.legal == 0) return 1.0f;  // unreachable: an empty cell implies legal > 0
    return __fmul_rn(1.0f, __fdiv_rn(1.0f, (float)h.legal));
}

// sequential (index-order) f32 sum of 81 values spread as v[j] = value of cell lane+32j;
// every lane computes the same chain -- this IS the reference's `iter().sum::<f32>()`.
__device__ __forceinline__ float seq_sum81(const float v[3]) {
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int lim = j == 2 ? kCells - 64 : 32;
        for (int l = 0; l < lim; ++l) s = __fadd_rn(s, __shfl_sync(kFull, v[j], l));
    }
    return s;
}

// Node::propagate (mcts/src/node.rs:83-99) along a recorded path of (node, action) edges,
// deepest edge last.  Lane 0 only.  Root statistics live in registers.
__device__ __forceinline__ void backup_path(uint8_t *tn, const uint32_t *path, int depth, float v, uint32_t &root_n,
                                            float &root_w) {
    for (int d = depth - 1; d >= 0; --d) {
        uint8_t *pn = node_ptr(tn, path[d] >> 8);
        const int a = path[d] & 0xFF;
        node_edge_n(pn)[a] += 1u;
        node_edge_w(pn)[a] = __fadd_rn(node_edge_w(pn)[a], v);
        v = -v;
    }
    root_n += 1u;
    root_w = __fadd_rn(root_w, v);
}

// ---------------------------------------------------------------------------------------
// Agent::new (agent.rs:16-35): fresh tree with an empty-board root; the root's raw policy
// arrives through the evaluator (request row per tree, Player encoding).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    k_new_games(TreeArgs a, const uint32_t *streams, const int8_t *only_if_terminal, const float *root_policy) {
    const int slot = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (slot >= a.n) return;
    const int tree = tree_of(a, slot);
    uint8_t *tn = tree_nodes_of(a, tree);
    if (only_if_terminal) {
        const int st = only_if_terminal[slot];
        if (st != kDraw && st != kBlackWin && st != kWhiteWin) {
            if (lane == 0) a.slot_count[slot] = 0;
            return;
        }
    }
    if (root_policy) {
        float *pol = node_policy(tn);
        for (int c = lane; c < kCells; c += 32) pol[c] = root_policy[c];
    }
    if (lane == 0) {
        TreeHdr th{};
        th.n_nodes = 1;
        th.root_n = 0;
        th.root_w = 0.0f;
        th.root_p = 1.0f;  // mcts/src/lib.rs:27
        if (only_if_terminal) {  // restart inside the self-play driver: the tree's random stream keeps running
            const TreeHdr old = a.hdrs[tree];
            th.rng_counter = old.rng_counter;
            th.stream = old.stream;
            th.noise_epoch = old.noise_epoch;
        } else {
            th.rng_counter = 0;
            th.stream = streams ? streams[slot] : (uint32_t)tree;
            th.noise_epoch = 0;
        }
        a.hdrs[tree] = th;
        NodeHdr h{};
        h.parent = kNoNode;
        h.action = 0xFF;
        h.status = kInProgress;
        h.legal = kCells;
        h.turn = 0;
        h.has_policy = root_policy ? 1 : 0;
        store_hdr(tn, h);
        if (root_policy) {
            a.slot_count[slot] = 0;
        } else {
            const uint32_t row = atomicAdd(a.n_req, 1u);
            a.slot_base[slot] = row;
            a.slot_count[slot] = row < (uint32_t)a.max_rows ? 1 : 0;
            if (row < (uint32_t)a.max_rows) {
                NNIn in{};
                a.nn_in[row] = in;
                a.req_tree[row] = (uint32_t)tree;
                a.req_node[row] = 0;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Root Dirichlet noise (parallel_mcts_executor.rs:48-76): first round of every execute.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_root_noise(TreeArgs a, float epsilon, float alpha) {
    const int slot = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (slot >= a.n) return;
    const int tree = tree_of(a, slot);
    uint8_t *tn = tree_nodes_of(a, tree);
    TreeHdr *th = &a.hdrs[tree];
    NodeHdr h = load_hdr(tn);
    float *pol = node_policy(tn);

    float p[3], nz[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int c = lane + 32 * j;
        p[j] = 0.0f;
        nz[j] = 0.0f;
        if (c < kCells) p[j] = h.has_policy ? pol[c] : dummy_prior(h, c);
    }
    if (epsilon != 0.0f) {
        const uint32_t epoch = th->noise_epoch;
        const uint32_t stream = th->stream;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int c = lane + 32 * j;
            if (c < kCells) nz[j] = noise_gamma_cell(a.seed, stream, epoch, c, alpha);
        }
        __syncwarp();
        const float s = seq_sum81(nz);
        const float inv = __fdiv_rn(1.0f, s);
#pragma unroll
        for (int j = 0; j < 3; ++j) nz[j] = __fmul_rn(nz[j], inv);
        if (lane == 0) th->noise_epoch = epoch + 1;
    }
    const float keep = __fsub_rn(1.0f, epsilon);
#pragma unroll
    for (int j = 0; j < 3; ++j) p[j] = __fadd_rn(__fmul_rn(keep, p[j]), __fmul_rn(epsilon, nz[j]));
    const float s = seq_sum81(p);
    const float inv = __fdiv_rn(1.0f, s);  // no EPSILON guard here (:63-68)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int c = lane + 32 * j;
        if (c < kCells) pol[c] = __fmul_rn(p[j], inv);
    }
    if (lane == 0 && !h.has_policy) {
        h.has_policy = 1;
        store_hdr(tn, h);
    }
}

// ---------------------------------------------------------------------------------------
// One round of `batch` selections + expansions for every tree
// (parallel_mcts_executor.rs:78-189).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_select_expand(TreeArgs a, int batch) {
    __shared__ uint32_t s_path[kWarpsPerBlock][kCells + 3];
    __shared__ uint16_t s_req[kWarpsPerBlock][kMaxBatchPerTree];
    const int wib = threadIdx.x >> 5;
    const int slot = blockIdx.x * kWarpsPerBlock + wib;
    const int lane = threadIdx.x & 31;
    if (slot >= a.n) return;
    const int tree = tree_of(a, slot);
    uint8_t *tn = tree_nodes_of(a, tree);
    TreeHdr *th = &a.hdrs[tree];
    uint32_t *path = s_path[wib];
    uint16_t *req = s_req[wib];

    uint32_t n_nodes = th->n_nodes;
    uint32_t root_n = th->root_n;
    float root_w = th->root_w;
    uint32_t rng_counter = th->rng_counter;
    const uint32_t stream = th->stream;
    uint32_t err = th->error;
    uint32_t sims = 0;
    int nreq = 0;

    // The descent is REUSED inside a round.  The reference re-runs select_leaf for each of a round's `batch` simulations,
    // but it has no virtual loss and applies the network results only after the round (:222-265), so between two
    // simulations the statistics on the path change only through a terminal backup (:92-97, :177-181), and the leaf stops
    // being a leaf only when the expansion just made fills it (node.rs:45-51).  Otherwise the next select_leaf returns the
    // same leaf by the same path (measured: 49 of 50 rounds of a search return one leaf 16 times), so the state of the
    // walk (x, depth, cur_n, h, path) is kept: after a backup the walk restarts at the root, after a filling expansion
    // it continues downwards from the leaf.  Same results bit for bit, ~1/16 of the PUCT scans.
    bool restart = true;
    uint32_t x = 0;
    uint32_t cur_n = root_n;
    int depth = 0;
    NodeHdr h{};
    int it = 0;
    while (it < batch && !err) {
        // ---- select_leaf (node.rs:39-59) with the PUCT selector (:81-90, :277-286) ----
        if (restart) {
            x = 0;
            cur_n = root_n;
            depth = 0;
            h = load_hdr(tn);
            restart = false;
        }
        for (;;) {
            const int nch = popc81(h.cmask);
            if (nch != (int)h.legal || nch == 0) break;
            uint8_t *xn = node_ptr(tn, x);
            const uint32_t parent_n = cur_n > 1u ? cur_n : 1u;
            const float sq = __fsqrt_rn((float)parent_n);
            const uint32_t *en = node_edge_n(xn);
            const float *ew = node_edge_w(xn);
            const float *pol = node_policy(xn);
            const uint16_t *ch = node_child(xn);
            unsigned long long best = 0ull;
            uint32_t best_n = 0;
            int best_a = 0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int c = lane + 32 * j;
                if (c < kCells && bit81(h.cmask, c)) {
                    const uint32_t n = en[c];
                    const float w = ew[c];
                    const float p = h.has_policy ? pol[c] : dummy_prior(h, c);
                    const float q = __fdiv_rn(w, __fadd_rn((float)n, kF32Eps));
                    const float bias = __fdiv_rn(sq, (float)(1u + n));
                    const float u = __fadd_rn(q, __fmul_rn(__fmul_rn(1.0f, p), bias));  // C_PUCT = 1.0
                    const uint32_t okey = (uint32_t)total_order_key(u) ^ 0x80000000u;
                    const unsigned long long cand = ((unsigned long long)okey << 32) | (unsigned long long)(ch[c] + 1u);
                    if (cand > best) {  // larger key, or equal key and later-created child
                        best = cand;
                        best_n = n;
                        best_a = c;
                    }
                }
            }
            unsigned long long wbest = best;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const unsigned long long o = __shfl_xor_sync(kFull, wbest, off);
                wbest = o > wbest ? o : wbest;
            }
            const int src = __ffs(__ballot_sync(kFull, best == wbest && best != 0ull)) - 1;
            best_n = __shfl_sync(kFull, best_n, src);
            best_a = __shfl_sync(kFull, best_a, src);
            if (lane == 0) path[depth] = (x << 8) | (uint32_t)best_a;
            ++depth;
            x = (uint32_t)(wbest & 0xFFFFFFFFull) - 1u;
            cur_n = best_n;
            h = load_hdr(node_ptr(tn, x));
        }
        __syncwarp();

        if (h.status != kInProgress) {  // :92-97 terminal leaf: backup its z, no expansion
            ++it;
            ++sims;
            if (lane == 0) {
                const float z = (h.status == kBlackWin || h.status == kWhiteWin) ? 1.0f : 0.0f;
                backup_path(tn, path, depth, z, root_n, root_w);
            }
            root_n = __shfl_sync(kFull, root_n, 0);
            root_w = __shfl_sync(kFull, root_w, 0);
            __syncwarp();
            restart = true;  // statistics along the path changed
            continue;
        }

        // ---- :99-125 uniformly random unexpanded legal actions (ascending candidate list) ----
        uint32_t cand[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) cand[k] = ~(h.black[k] | h.white[k]) & ~h.cmask[k];
        cand[2] &= 0x1FFFFu;
        const int ncand = popc81(cand);
        if (ncand == 0) {  // (unreachable single-threaded, :119-125: the simulation is spent)
            ++it;
            ++sims;
            continue;
        }
        // A RUN of expansions at this leaf.  Until a terminal child is backed up nothing the selection reads changes (see
        // above), so the next simulations of the round come back to this leaf until it is full: their draws are made here
        // one after the other (register work: the stream position and the shrinking candidate set are the only state that
        // links them), and then LANE k places the stone of expansion k, tests five-in-a-row on its own registers and stores
        // child k -- the serial chain of a round is K draws instead of K x (draw, stone, warp-ballot line scan, header
        // store).  The run is cut behind the first terminal child: its reward is backed up at once (:177-181) and changes
        // the statistics, so the draws made for the simulations behind it are handed back (the stream position recorded
        // with expansion k is restored) and the walk restarts at the root -- the same children, ids, stream position and
        // statistics as one simulation at a time (every parity test).  Virtual loss changes the path after each expansion:
        // runs of one.
        const int K = a.vloss ? 1 : min(min(batch - it, ncand), 32);
        int my_action = 0;
        uint32_t my_ctr = rng_counter;
        {
            uint32_t cw[3] = {cand[0], cand[1], cand[2]};
            uint32_t ctr = rng_counter;
            for (int k = 0; k < K; ++k) {
                const int act = nth_set81(cw, (int)rng_below(a.seed, stream, ctr, (uint32_t)(ncand - k)));
                cw[act >> 5] &= ~(1u << (act & 31));  // (act >> 5 is warp-uniform)
                if (lane == k) {
                    my_action = act;
                    my_ctr = ctr;
                }
            }
        }
        // ---- :127-135 lane k: the stone of expansion k on a clone of the leaf's environment ----
        NodeHdr c = h;
        bool five = false;
        if (lane < K) {
            uint32_t *mine = h.turn == 0 ? c.black : c.white;
            set81(mine, my_action);
            five = makes_five_scalar(mine, my_action);
            c.legal = h.legal - 1;
            c.turn = h.turn ^ 1u;
            c.status = five ? (h.turn == 0 ? kBlackWin : kWhiteWin) : (c.legal == 0 ? kDraw : kInProgress);
            c.cmask[0] = c.cmask[1] = c.cmask[2] = 0;
            c.parent = x;
            c.action = (uint32_t)my_action;
            c.has_policy = 0;
        }
        const uint32_t term_mask = __ballot_sync(kFull, lane < K && c.status != kInProgress);
        const int t_first = term_mask ? __ffs(term_mask) - 1 : K;  // expansions before the first terminal child
        int accepted = t_first < K ? t_first + 1 : K;
        // ---- :158-175 expand: the child's prior is the leaf's policy entry (implicit) ----
        if (n_nodes + (uint32_t)accepted > (uint32_t)a.cap_nodes) {  // the tree is full: keep what fits, then fail loudly
            accepted = (int)((uint32_t)a.cap_nodes - n_nodes);
            err = 1;
        }
        const bool term_accepted = t_first < accepted;
        const uint32_t id = n_nodes + (uint32_t)lane;
        uint8_t *xn = node_ptr(tn, x);
        uint32_t add[3] = {0u, 0u, 0u};
        if (lane < accepted) {
            store_hdr(node_ptr(tn, id), c);
            node_edge_n(xn)[my_action] = 0u;
            node_edge_w(xn)[my_action] = 0.0f;
            node_child(xn)[my_action] = (uint16_t)id;
            add[my_action >> 5] = 1u << (my_action & 31);
            if (lane < t_first) req[nreq + lane] = (uint16_t)id;  // :182-187 queue for the evaluator
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            add[k] = __reduce_or_sync(kFull, add[k]);
            h.cmask[k] |= add[k];  // the kept copy of the leaf's header follows the one in memory
        }
        if (lane < 3 && sel3(add[0], add[1], add[2], lane)) reinterpret_cast<uint32_t *>(xn)[6 + lane] = sel3(h.cmask[0], h.cmask[1], h.cmask[2], lane);
        if (accepted > 0) rng_counter = __shfl_sync(kFull, my_ctr, accepted - 1);
        n_nodes += (uint32_t)accepted;
        nreq += min(t_first, accepted);
        it += accepted;
        sims += (uint32_t)accepted;
        __syncwarp();  // the lanes' stores above are ordered before lane 0's read-modify-writes below (and the next PUCT scan)
        if (term_accepted) {  // :177-181 terminal child: backup the reward now
            const int t_action = __shfl_sync(kFull, my_action, t_first);
            const bool t_five = __shfl_sync(kFull, (int)five, t_first) != 0;
            if (lane == 0) {
                path[depth] = (x << 8) | (uint32_t)t_action;
                backup_path(tn, path, depth + 1, t_five ? 1.0f : 0.0f, root_n, root_w);
            }
            restart = true;  // a terminal child was backed up along the path
        } else if (a.vloss && accepted == 1) {
            // VIRTUAL LOSS (opt-in; the reference has none, :80-90 reads the statistics as they are): the pending
            // evaluation counts as one visit with value -1 on every edge of its path, so the round's next
            // simulations spread over other lines.  k_apply replaces it with the real value (n + 0, w + v + 1).
            const int v_action = __shfl_sync(kFull, my_action, 0);
            if (lane == 0) {
                path[depth] = (x << 8) | (uint32_t)v_action;
                for (int d = depth; d >= 0; --d) {
                    uint8_t *pn = node_ptr(tn, path[d] >> 8);
                    const int pa = path[d] & 0xFF;
                    node_edge_n(pn)[pa] += 1u;
                    node_edge_w(pn)[pa] = __fsub_rn(node_edge_w(pn)[pa], 1.0f);
                }
                root_n += 1u;
            }
            restart = true;  // the virtual loss changed the path's statistics
        }
#ifdef OMK_NO_DESCENT_REUSE
        restart = true;  // A/B build: a full descent per run
#endif
        root_n = __shfl_sync(kFull, root_n, 0);
        root_w = __shfl_sync(kFull, root_w, 0);
        __syncwarp();
    }

    // ---- publish this tree's requests as one contiguous run of evaluator rows ----
    uint32_t base = 0;
    if (lane == 0) base = nreq ? atomicAdd(a.n_req, (uint32_t)nreq) : 0u;
    base = __shfl_sync(kFull, base, 0);
    if (base + (uint32_t)nreq > (uint32_t)a.max_rows) {  // cannot happen when max_rows >= n*batch
        err = 1;
        nreq = 0;
    }
    __syncwarp();
    for (int i = lane; i < nreq; i += 32) {
        const uint32_t id = req[i];
        const NodeHdr r = load_hdr(node_ptr(tn, id));
        NNIn in;
        in.black[0] = r.black[0]; in.black[1] = r.black[1]; in.black[2] = r.black[2];
        in.white[0] = r.white[0]; in.white[1] = r.white[1]; in.white[2] = r.white[2];
        in.meta = r.turn;  // EnvTurnMode::Player (:215-219)
        in.pad = 0;
        a.nn_in[base + i] = in;
        a.req_tree[base + i] = (uint32_t)tree;
        a.req_node[base + i] = id;
    }
    if (lane == 0) {
        a.slot_base[slot] = base;
        a.slot_count[slot] = (uint32_t)nreq;
        th->n_nodes = n_nodes;
        th->root_n = root_n;
        th->root_w = root_w;
        th->rng_counter = rng_counter;
        if (err) {
            th->error = err;
            atomicOr(a.dev_error, 1u);
        }
        atomicAdd(a.dev_sims, (unsigned long long)sims);
    }
}

// ---------------------------------------------------------------------------------------
// Evaluator results -> tree.  mode kApplySearch: parallel_mcts_executor.rs:222-265;
// kApplyNewGame: agent.rs:20-25 (raw, unmasked); kApplyEnsure: agent.rs:159-182.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_apply(TreeArgs a, int mode) {
    // Requests of a tree are handled in chunks of 16, in four warp-wide phases, so that the dependent chains of the 16
    // requests overlap instead of running back to back: (B) gather + mask the 16 network rows into shared memory,
    // (C) lane i runs request i's SEQUENTIAL index-order f32 sum (the reference's `iter().sum()`, :241) -- 16 chains side
    // by side instead of 16 x 81 dependent warp shuffles --, (D) scale + store the policies, (E) lane 0 applies the
    // backups in request order (the f32 adds along shared ancestors are order-dependent).
    constexpr int kChunk = 16;
    __shared__ float s_p[kWarpsPerBlock][kChunk][kCells];       // odd row stride: the per-lane chains are conflict-free
    __shared__ uint32_t s_meta[kWarpsPerBlock][kChunk][4];      // node id, header word 9 (parent | action), word 10, -V bits
    const int wib = threadIdx.x >> 5;
    const int slot = blockIdx.x * kWarpsPerBlock + wib;
    const int lane = threadIdx.x & 31;
    if (slot >= a.n) return;
    const int tree = tree_of(a, slot);
    uint8_t *tn = tree_nodes_of(a, tree);
    TreeHdr *th = &a.hdrs[tree];
    const uint32_t base = a.slot_base[slot];
    const int cnt = (int)a.slot_count[slot];
    if (cnt == 0) return;
    uint32_t root_n = th->root_n;
    float root_w = th->root_w;
    uint32_t root_occ[3] = {0u, 0u, 0u};
    if (mode == kApplyEnsure) {
        const NodeHdr root_h = load_hdr(tn);
#pragma unroll
        for (int k = 0; k < 3; ++k) root_occ[k] = root_h.black[k] | root_h.white[k];
    }
    float(*sp)[kCells] = s_p[wib];
    uint32_t(*meta)[4] = s_meta[wib];

    for (int c0 = 0; c0 < cnt; c0 += kChunk) {
        const int m = min(kChunk, cnt - c0);
        const uint32_t row0 = base + (uint32_t)c0;
        // ---- (A) lane i fetches request i's node id, value and node header: one round of loads for the chunk's headers
        //      (the first version broadcast-loaded one header per request inside the gather loop, whose four unrolled
        //      batches each waited for a header and then for its network row: 61 % of the kernel's stall samples) ----
        uint32_t my_id = 0, my_action = 0;
        float my_v = 0.0f;
        uint32_t my_occ[3] = {0u, 0u, 0u};
        if (lane < m) {
            my_id = a.req_node[row0 + lane];
            if (mode == kApplySearch) my_v = -a.V[row0 + lane];  // :229 value from the opponent's perspective
            const NodeHdr h = load_hdr(node_ptr(tn, my_id));
#pragma unroll
            for (int k = 0; k < 3; ++k) my_occ[k] = h.black[k] | h.white[k];
            my_action = h.action;
            meta[lane][0] = my_id;
            meta[lane][1] = h.parent | (h.action << 16);
            meta[lane][2] = h.legal | (h.turn << 8) | (1u << 16);  // header word 10 with has_policy set
            meta[lane][3] = __float_as_uint(my_v);
        }
        // ---- (B) gather + mask (:232-239): the rows' loads do not depend on the headers ----
#pragma unroll 8
        for (int i = 0; i < m; ++i) {
            const float *Prow = a.P + (size_t)(row0 + i) * kRow;
            float pv[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int c = lane + 32 * j;
                pv[j] = c < kCells ? Prow[c] : 0.0f;
            }
            const uint32_t o0 = __shfl_sync(kFull, my_occ[0], i), o1 = __shfl_sync(kFull, my_occ[1], i), o2 = __shfl_sync(kFull, my_occ[2], i);
            const int act_i = (int)__shfl_sync(kFull, my_action, i);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int c = lane + 32 * j;
                if (c < kCells) {
                    float v = pv[j];
                    if (mode == kApplySearch) {  // mask by the node's own board
                        const uint32_t occ = sel3(o0, o1, o2, j);
                        if ((occ >> (c & 31)) & 1u) v = 0.0f;
                    } else if (mode == kApplyEnsure) {  // agent.rs:165-171: the action, then the ROOT's occupancy
                        const uint32_t occ = sel3(root_occ[0], root_occ[1], root_occ[2], j);
                        if (c == act_i || ((occ >> (c & 31)) & 1u)) v = 0.0f;
                    }
                    sp[i][c] = v;
                }
            }
        }
        __syncwarp();
        // ---- (C) sequential sums, one request per lane (:241-249) ----
        float inv = 0.0f;
        bool scale = false;
        if (mode != kApplyNewGame && lane < m) {
            float sum = 0.0f;
#pragma unroll 9
            for (int c = 0; c < kCells; ++c) sum = __fadd_rn(sum, sp[lane][c]);
            scale = kF32Eps <= sum;
            if (scale) inv = __fdiv_rn(1.0f, sum);
        }
        // ---- (D) scale + store the policies (:243-252) ----
#pragma unroll 4
        for (int i = 0; i < m; ++i) {
            const float inv_i = __shfl_sync(kFull, inv, i);
            const bool scale_i = __shfl_sync(kFull, (int)scale, i) != 0;
            const uint32_t id = __shfl_sync(kFull, my_id, i);
            float *pol = node_policy(node_ptr(tn, id));
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int c = lane + 32 * j;
                if (c < kCells) {
                    const float v = sp[i][c];
                    pol[c] = scale_i ? __fmul_rn(v, inv_i) : v;
                }
            }
        }
        // ---- (E) has_policy flag + backups (:264, node.rs:83-99) ----
        // The f32 adds on an edge must happen in request order.  Two facts make most of them register work: (1) the edge
        // that leads to request i's own node -- (parent_i, action_i) -- is touched first by request i itself (later requests
        // can pass through the node, earlier ones cannot: it did not exist), and no two requests share it, so lane i
        // updates it on its own, all lanes at once; (2) a round's requests are mostly siblings (one leaf, filled child by
        // child), i.e. consecutive requests walk the SAME chain of ancestors above the leaf: lane 0 keeps that chain's
        // statistics in registers, applies the requests' values in order, and writes the chain back when the leaf changes.
        // (The first version walked every request's whole path through global memory: 16 x depth dependent
        // read-modify-writes, 18.8 us for one tree's round of 8.)
        if (lane < m) {
            uint8_t *nd = node_ptr(tn, meta[lane][0]);
            reinterpret_cast<uint32_t *>(nd)[10] = meta[lane][2];
            if (mode == kApplySearch) {
                const uint32_t par = meta[lane][1] & 0xFFFFu, act = (meta[lane][1] >> 16) & 0xFFu;
                if (par != kNoNode) {
                    uint8_t *pn = node_ptr(tn, par);
                    const float v = __uint_as_float(meta[lane][3]);
                    if (a.vloss) {  // the visit was counted at selection time with value -1: swap in the real value
                        node_edge_w(pn)[act] = __fadd_rn(node_edge_w(pn)[act], __fadd_rn(v, 1.0f));
                    } else {
                        node_edge_n(pn)[act] += 1u;
                        node_edge_w(pn)[act] = __fadd_rn(node_edge_w(pn)[act], v);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0 && mode == kApplySearch) {
            constexpr int kChain = 6;            // cached levels above the leaf; deeper levels go through memory
            uint32_t ch_leaf = 0xFFFFFFFFu;      // the leaf (parent of the requests' nodes) whose chain is cached
            int ch_len = 0;                      // cached levels
            uint32_t ch_next = kNoNode, ch_next_act = 0;  // where the walk continues above the cached levels
            uint32_t ch_node[kChain], ch_act[kChain], ch_n[kChain];
            float ch_w[kChain];
            auto flush = [&]() {
#pragma unroll
                for (int k = 0; k < kChain; ++k)
                    if (k < ch_len) {
                        uint8_t *pn = node_ptr(tn, ch_node[k]);
                        if (!a.vloss) node_edge_n(pn)[ch_act[k]] = ch_n[k];
                        node_edge_w(pn)[ch_act[k]] = ch_w[k];
                    }
                ch_len = 0;
                ch_leaf = 0xFFFFFFFFu;
            };
            for (int i = 0; i < m; ++i) {
                const uint32_t leaf = meta[i][1] & 0xFFFFu;
                float v = __uint_as_float(meta[i][3]);
                if (leaf != kNoNode) {
                    v = -v;  // the leaf-level edge was updated by lane i above
                    if (leaf != ch_leaf) {  // another chain: write the cached one back, load this one
                        flush();
                        ch_leaf = leaf;
                        uint32_t w9 = reinterpret_cast<const uint32_t *>(node_ptr(tn, leaf))[9];
                        uint32_t up = w9 & 0xFFFFu, up_act = (w9 >> 16) & 0xFFu;
#pragma unroll
                        for (int k = 0; k < kChain; ++k)
                            if (up != kNoNode && ch_len == k) {
                                uint8_t *pn = node_ptr(tn, up);
                                ch_node[k] = up;
                                ch_act[k] = up_act;
                                ch_n[k] = node_edge_n(pn)[up_act];
                                ch_w[k] = node_edge_w(pn)[up_act];
                                ch_len = k + 1;
                                w9 = reinterpret_cast<const uint32_t *>(pn)[9];
                                up = w9 & 0xFFFFu;
                                up_act = (w9 >> 16) & 0xFFu;
                            }
                        ch_next = up;
                        ch_next_act = up_act;
                    }
#pragma unroll
                    for (int k = 0; k < kChain; ++k)
                        if (k < ch_len) {
                            if (a.vloss) {
                                ch_w[k] = __fadd_rn(ch_w[k], __fadd_rn(v, 1.0f));
                            } else {
                                ch_n[k] += 1u;
                                ch_w[k] = __fadd_rn(ch_w[k], v);
                            }
                            v = -v;
                        }
                    // levels above the cached ones (deep paths): through memory, as every level used to
                    uint32_t cur_parent = ch_next, cur_action = ch_next_act;
                    while (cur_parent != kNoNode) {
                        uint8_t *pn = node_ptr(tn, cur_parent);
                        if (a.vloss) {
                            node_edge_w(pn)[cur_action] = __fadd_rn(node_edge_w(pn)[cur_action], __fadd_rn(v, 1.0f));
                        } else {
                            node_edge_n(pn)[cur_action] += 1u;
                            node_edge_w(pn)[cur_action] = __fadd_rn(node_edge_w(pn)[cur_action], v);
                        }
                        v = -v;
                        const uint32_t w9 = reinterpret_cast<const uint32_t *>(pn)[9];
                        cur_parent = w9 & 0xFFFFu;
                        cur_action = (w9 >> 16) & 0xFFu;
                    }
                }
                if (!a.vloss) root_n += 1u;
                root_w = __fadd_rn(root_w, v);
            }
            flush();
        }
        __syncwarp();
    }
    if (lane == 0 && mode == kApplySearch) {
        th->root_n = root_n;
        th->root_w = root_w;
    }
}

// ---------------------------------------------------------------------------------------
// Agent::compute_policy + sample_action (agent.rs:43-137)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    k_sample(TreeArgs a, const uint8_t *modes, const float *temps, int32_t *out_action, float *out_policy,
             uint8_t *out_valid) {
    const int slot = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (slot >= a.n) return;
    const int tree = tree_of(a, slot);
    uint8_t *tn = tree_nodes_of(a, tree);
    TreeHdr *th = &a.hdrs[tree];
    const NodeHdr h = load_hdr(tn);
    const uint32_t *en = node_edge_n(tn);

    uint32_t cnt[3];
    uint32_t total = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int c = lane + 32 * j;
        cnt[j] = (c < kCells && bit81(h.cmask, c)) ? en[c] : 0u;
        total += cnt[j];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(kFull, total, off);
    // all partial sums of `sum += n as f32` are integers < 2^24, hence exact in any order
    const float sum = (float)total;
    const bool none = popc81(h.cmask) == 0 || sum < kF32Eps;
    float pi[3];
    const float inv = none ? 0.0f : __fdiv_rn(1.0f, sum);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        pi[j] = __fmul_rn((float)cnt[j], inv);
        const int c = lane + 32 * j;
        if (out_policy && c < kCells) out_policy[(size_t)slot * kCells + c] = none ? 0.0f : pi[j];
    }
    if (out_valid && lane == 0) out_valid[slot] = none ? 0 : 1;
    if (!out_action) return;
    if (none) {
        if (lane == 0) out_action[slot] = OMK_NONE;
        return;
    }
    const int mode = modes ? modes[slot] : OMK_SAMPLE_BEST;
    if (mode == OMK_SAMPLE_BEST) {  // max_by(total_cmp): last maximum == highest index among maxima
        unsigned long long best = 0ull;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int c = lane + 32 * j;
            if (c < kCells) {
                const uint32_t okey = (uint32_t)total_order_key(pi[j]) ^ 0x80000000u;
                const unsigned long long cand = ((unsigned long long)okey << 32) | (unsigned long long)(c + 1);
                best = cand > best ? cand : best;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(kFull, best, off);
            best = o > best ? o : best;
        }
        if (lane == 0) out_action[slot] = (int32_t)(best & 0xFFFFFFFFull) - 1;
        return;
    }
    // Boltzmann(T): weights exp(pi/T) for pi >= EPSILON, normalised; WeightedIndex sampling
    const float tinv = __fdiv_rn(1.0f, temps[slot]);
    float heated[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int c = lane + 32 * j;
        heated[j] = 0.0f;
        if (c < kCells && !(pi[j] < kF32Eps)) heated[j] = (float)det_exp((double)__fmul_rn(pi[j], tinv));
    }
    __syncwarp();
    const float hs = seq_sum81(heated);
    const float hinv = __fdiv_rn(1.0f, hs);
#pragma unroll
    for (int j = 0; j < 3; ++j) heated[j] = __fmul_rn(heated[j], hinv);
    // cumulative weights (rand 0.8.5 WeightedIndex::new): cum[i] = sum of weights[0..=i], i < 80
    float cum[3] = {0.0f, 0.0f, 0.0f};
    float tot = 0.0f;
    for (int c = 0; c < kCells; ++c) {
        const float v = __shfl_sync(kFull, heated[c >> 5], c & 31);
        tot = c == 0 ? v : __fadd_rn(tot, v);
        if ((c & 31) == lane) {
            if ((c >> 5) == 0) cum[0] = tot; else if ((c >> 5) == 1) cum[1] = tot; else cum[2] = tot;
        }
    }
    float scale = tot;  // Uniform::new(0, total)
    const float max_rand = 0.99999988079071044921875f;
    while (__fadd_rn(__fmul_rn(scale, max_rand), 0.0f) >= tot) scale = __int_as_float(__float_as_int(scale) - 1);
    uint32_t ctr = th->rng_counter;
    const uint32_t r = rng_u32(a.seed, th->stream, ctr);
    const float value1_2 = __int_as_float((int)((r >> 9) | 0x3f800000u));
    const float chosen = __fadd_rn(__fmul_rn(__fsub_rn(value1_2, 1.0f), scale), 0.0f);
    int below = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int c = lane + 32 * j;
        const bool le = c < kCells - 1 && cum[j] <= chosen;  // partition_point over the first 80 sums
        below += __popc(__ballot_sync(kFull, le));
    }
    if (lane == 0) {
        out_action[slot] = below;
        th->rng_counter = ctr + 1;
    }
}

// ---------------------------------------------------------------------------------------
// Agent::ensure_action_exists, part 1 (agent.rs:144-157, 184-194): create the child and
// queue its evaluate_p (Opponent encoding).  The reference evaluates first and drops the
// result when the child already exists; skipping that evaluation is unobservable.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_ensure_prepare(TreeArgs a, const int32_t *actions) {
    const int slot = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (slot >= a.n) return;
    const int tree = tree_of(a, slot);
    uint8_t *tn = tree_nodes_of(a, tree);
    TreeHdr *th = &a.hdrs[tree];
    const int action = actions[slot];
    int nreq = 0;
    uint32_t id = 0;
    NodeHdr c{};
    if (action >= 0 && action < kCells) {
        const NodeHdr h = load_hdr(tn);
        if (!bit81(h.cmask, action)) {
            c = h;
            const uint32_t occ = sel3(h.black[0] | h.white[0], h.black[1] | h.white[1], h.black[2] | h.white[2], action >> 5);
            if (!((occ >> (action & 31)) & 1u)) {  // place_stone's result is ignored (:155)
                set81(h.turn == 0 ? c.black : c.white, action);
                c.legal = h.legal - 1;
                c.turn = h.turn ^ 1u;
            }
            c.status = kInProgress;  // :190 always InProgress
            c.cmask[0] = c.cmask[1] = c.cmask[2] = 0;
            c.parent = 0;
            c.action = (uint32_t)action;
            c.has_policy = 0;
            const uint32_t n_nodes = th->n_nodes;
            if (n_nodes >= (uint32_t)a.cap_nodes) {
                if (lane == 0) {
                    th->error = 1;
                    atomicOr(a.dev_error, 1u);
                }
            } else {
                id = n_nodes;
                nreq = 1;
                if (lane == 0) {
                    th->n_nodes = n_nodes + 1;
                    store_hdr(node_ptr(tn, id), c);
                    reinterpret_cast<uint32_t *>(tn)[6 + (action >> 5)] =
                        sel3(h.cmask[0], h.cmask[1], h.cmask[2], action >> 5) | (1u << (action & 31));
                    node_edge_n(tn)[action] = 0u;
                    node_edge_w(tn)[action] = 0.0f;
                    node_child(tn)[action] = (uint16_t)id;
                }
            }
        }
    }
    if (lane == 0) {
        uint32_t base = 0;
        if (nreq) {
            base = atomicAdd(a.n_req, 1u);
            if (base < (uint32_t)a.max_rows) {
                NNIn in;
                in.black[0] = c.black[0]; in.black[1] = c.black[1]; in.black[2] = c.black[2];
                in.white[0] = c.white[0]; in.white[1] = c.white[1]; in.white[2] = c.white[2];
                in.meta = c.turn | 2u;  // EnvTurnMode::Opponent (:157)
                in.pad = 0;
                a.nn_in[base] = in;
                a.req_tree[base] = (uint32_t)tree;
                a.req_node[base] = id;
            } else {
                nreq = 0;
            }
        }
        a.slot_base[slot] = base;
        a.slot_count[slot] = (uint32_t)nreq;
    }
}

// ---------------------------------------------------------------------------------------
// Agent::play_action (agent.rs:206-232) + MCTS::transition (mcts/src/lib.rs:47-78).
// The kept subtree is compacted in place to ids 0..kept-1 (creation order preserved), which
// is this design's equivalent of the reference's recursive free + free list.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    k_play(TreeArgs a, const int32_t *actions, int8_t *out_status, uint16_t *remap_all) {
    const int slot = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (slot >= a.n) return;
    const int tree = tree_of(a, slot);
    uint8_t *tn = tree_nodes_of(a, tree);
    TreeHdr *th = &a.hdrs[tree];
    uint16_t *remap = remap_all + (size_t)tree * (size_t)a.cap_nodes;
    const int action = actions[slot];
    const NodeHdr h0 = load_hdr(tn);

    int status = OMK_NONE;
    bool ok = h0.status == kInProgress && action >= 0 && action < kCells && bit81(h0.cmask, action);
    if (ok) {
        const uint32_t occ = sel3(h0.black[0] | h0.white[0], h0.black[1] | h0.white[1], h0.black[2] | h0.white[2], action >> 5);
        ok = !((occ >> (action & 31)) & 1u);  // self.env.place_stone(action)? (:224-228)
    }
    if (!ok) {
        if (lane == 0 && out_status) out_status[slot] = (int8_t)OMK_NONE;
        return;
    }
    {   // Agent.env.place_stone: the agent's env always equals the root node's env
        uint32_t mine[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) mine[k] = h0.turn == 0 ? h0.black[k] : h0.white[k];
        set81(mine, action);
        const bool five = makes_five_warp(mine, action, lane);
        status = five ? (h0.turn == 0 ? kBlackWin : kWhiteWin) : (h0.legal - 1 == 0 ? kDraw : kInProgress);
    }
    if (lane == 0 && out_status) out_status[slot] = (int8_t)status;

    // ---- transition: the chosen child becomes the root ----
    const uint32_t c = node_child(tn)[action];
    const NodeHdr hc = load_hdr(node_ptr(tn, c));
    uint32_t new_root_n = 0;  // :65-71 n := sum of its children's n
    {
        const uint32_t *en = node_edge_n(node_ptr(tn, c));
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int k = lane + 32 * j;
            if (k < kCells && bit81(hc.cmask, k)) new_root_n += en[k];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) new_root_n += __shfl_xor_sync(kFull, new_root_n, off);
    }
    const float new_root_w = node_edge_w(tn)[action];
    const float new_root_p = h0.has_policy ? node_policy(tn)[action] : dummy_prior(h0, action);
    const uint32_t n_nodes = th->n_nodes;
    __syncwarp();

    // pass 1: keep flags + new ids.  parent id < child id always, so one ascending sweep suffices;
    // parents inside the current 32-chunk are resolved by iterating to a fixed point.
    uint32_t kept = 0;
    for (uint32_t base = 0; base < n_nodes; base += 32) {
        const uint32_t id = base + lane;
        bool keep = false;
        uint32_t parent = kNoNode;
        if (id < n_nodes) {
            parent = reinterpret_cast<const uint32_t *>(node_ptr(tn, id))[9] & 0xFFFFu;
            keep = id == c;
            if (!keep && parent != kNoNode && parent < base && id != 0) keep = remap[parent] != kNoNode;
        }
        const bool in_chunk = id < n_nodes && parent != kNoNode && parent >= base && id != c;
        for (;;) {
            const uint32_t kmask = __ballot_sync(kFull, keep);
            bool nk = keep;
            if (in_chunk && ((kmask >> (parent - base)) & 1u)) nk = true;
            const bool changed = nk != keep;
            keep = nk;
            if (!__any_sync(kFull, changed)) break;
        }
        const uint32_t kmask = __ballot_sync(kFull, keep);
        if (id < n_nodes) remap[id] = keep ? (uint16_t)(kept + __popc(kmask & ((1u << lane) - 1u))) : (uint16_t)kNoNode;
        kept += __popc(kmask);
        __syncwarp();
    }

    // pass 2: move kept records downwards in ascending order (dest <= src, so in place is safe)
    for (uint32_t id = c; id < n_nodes; ++id) {
        const uint32_t nid = remap[id];
        if (nid == kNoNode) continue;
        uint8_t *src = node_ptr(tn, id);
        uint8_t *dst = node_ptr(tn, nid);
        const NodeHdr hs = load_hdr(src);
        const bool has_children = (hs.cmask[0] | hs.cmask[1] | hs.cmask[2]) != 0u;
        uint4 buf[3];
        bool use[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int q = lane + 32 * j;  // uint4 index within the 1408-byte record (88 of them)
            use[j] = q < 88 && (q < 4 || (q < 28 ? hs.has_policy != 0 : has_children));
            if (use[j]) buf[j] = reinterpret_cast<const uint4 *>(src)[q];
        }
        __syncwarp();
        if (nid != id) {
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (use[j]) reinterpret_cast<uint4 *>(dst)[lane + 32 * j] = buf[j];
        }
        __syncwarp();
        // fix-ups: parent id and child ids
        if (lane == 0) {
            const uint32_t np = id == c ? kNoNode : (uint32_t)remap[hs.parent];
            reinterpret_cast<uint32_t *>(dst)[9] = np | (hs.action << 16) | (hs.status << 24);
        }
        if (has_children) {
            uint16_t *ch = node_child(dst);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int k = lane + 32 * j;
                if (k < kCells && bit81(hs.cmask, k)) ch[k] = remap[ch[k]];
            }
        }
        __syncwarp();
    }
    if (lane == 0) {
        th->n_nodes = kept;
        th->root_n = new_root_n;
        th->root_w = new_root_w;
        th->root_p = new_root_p;
    }
}

// ---------------------------------------------------------------------------------------
// inspection: root children in creation order (== ascending child id) + root statistics
// ---------------------------------------------------------------------------------------
__global__ void k_root_children(TreeArgs a, int tree, int32_t *out_actions, unsigned long long *out_n, float *out_w,
                                float *out_p, int32_t *out_len, float *out_policy, uint32_t *out_misc) {
    uint8_t *tn = tree_nodes_of(a, tree);
    const TreeHdr *th = &a.hdrs[tree];
    const NodeHdr h = load_hdr(tn);
    if (threadIdx.x == 0) {
        int len = 0;
        long long last_id = -1;
        for (;;) {  // ascending child id == creation order
            int best_a = -1;
            long long best_id = 1ll << 40;
            for (int c = 0; c < kCells; ++c) {
                if (!bit81(h.cmask, c)) continue;
                const long long cid = node_child(tn)[c];
                if (cid > last_id && cid < best_id) {
                    best_id = cid;
                    best_a = c;
                }
            }
            if (best_a < 0) break;
            last_id = best_id;
            out_actions[len] = best_a;
            out_n[len] = node_edge_n(tn)[best_a];
            out_w[len] = node_edge_w(tn)[best_a];
            out_p[len] = h.has_policy ? node_policy(tn)[best_a] : dummy_prior(h, best_a);
            ++len;
        }
        *out_len = len;
        out_misc[0] = th->root_n;
        out_misc[1] = __float_as_uint(th->root_w);
        out_misc[2] = __float_as_uint(th->root_p);
        out_misc[3] = h.status;
        out_misc[4] = th->n_nodes;
        out_misc[5] = th->rng_counter;
        out_misc[6] = h.turn | (h.legal << 8);
        out_misc[7] = th->error;
        for (int k = 0; k < 3; ++k) {
            out_misc[8 + k] = h.black[k];
            out_misc[11 + k] = h.white[k];
        }
    }
    for (int c = threadIdx.x; c < kCells; c += blockDim.x)
        out_policy[c] = h.has_policy ? node_policy(tn)[c] : dummy_prior(h, c);
}

// Agent.env of many trees at once (the root node's board, turn, legal_move_count and status)
__global__ void k_pool_get_envs(TreeArgs a, uint8_t *boards, uint8_t *turns, uint16_t *legal, int8_t *status) {
    const int slot = blockIdx.x;
    if (slot >= a.n) return;
    const NodeHdr h = load_hdr(tree_nodes_of(a, tree_of(a, slot)));
    if (boards)
        for (int c = threadIdx.x; c < kCells; c += blockDim.x)
            boards[(size_t)slot * kCells + c] = bit81(h.black, c) ? 1 : (bit81(h.white, c) ? 2 : 0);
    if (threadIdx.x == 0) {
        if (turns) turns[slot] = (uint8_t)h.turn;
        if (legal) legal[slot] = (uint16_t)h.legal;
        if (status) status[slot] = (int8_t)h.status;
    }
}

// ---------------------------------------------------------------------------------------
// OMK_EVAL_HASH: exact fake evaluator over the pending request rows
// ---------------------------------------------------------------------------------------
__global__ void k_eval_hash(const NNIn *in, const uint32_t *n_req, float *P, float *V, int max_rows) {
    const uint32_t rows = min(*n_req, (uint32_t)max_rows);
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t row = idx / kRow, c = idx % kRow;
    if (row >= rows) return;
    const NNIn r = in[row];
    const uint64_t hsh = hash_board(r.black, r.white, r.meta & 1u, (r.meta >> 1) & 1u);
    if (c < (uint32_t)kCells) P[(size_t)row * kRow + c] = hash_policy(hsh, (int)c);
    if (c == (uint32_t)kCells) V[row] = hash_value(hsh);
}

// zero the request counter, folding the previous count into the evaluated-positions total
__global__ void k_reset_requests(uint32_t *n_req, unsigned long long *dev_sims) {
    atomicAdd(&dev_sims[1], (unsigned long long)*n_req);  // two search lanes fold their counters concurrently
    *n_req = 0;
}

// ---- self-play driver glue (src/trainer.rs:86-204 restated on the device) ----
__global__ void k_sp_prepare(int g0, int n, const int32_t *ply, int threshold, float temperature, int32_t *mover, int32_t *other,
                             uint8_t *modes, float *temps) {
    const int g = g0 + blockIdx.x * blockDim.x + threadIdx.x;  // games [g0, g0 + n): one search lane's slice
    if (g >= g0 + n) return;
    const int p = ply[g];
    mover[g] = 2 * g + (p & 1);       // black agent = tree 2g, white agent = tree 2g+1 (trainer.rs:86-94)
    other[g] = 2 * g + 1 - (p & 1);
    modes[g] = p < threshold ? OMK_SAMPLE_BOLTZMANN : OMK_SAMPLE_BEST;  // trainer.rs:138-146
    temps[g] = temperature;
}

__global__ void k_sp_record(TreeArgs a, const int32_t *actions, const float *policy_in, uint8_t *boards_out,
                            float *policy_out, int32_t *actions_out) {
    const int g = blockIdx.x;
    if (g >= a.n) return;
    const int tree = tree_of(a, g);
    const NodeHdr h = load_hdr(tree_nodes_of(a, tree));
    for (int c = threadIdx.x; c < kCells; c += blockDim.x) {
        boards_out[(size_t)g * kCells + c] = bit81(h.black, c) ? 1 : (bit81(h.white, c) ? 2 : 0);  // env before the action
        policy_out[(size_t)g * kCells + c] = policy_in[(size_t)g * kCells + c];
    }
    if (threadIdx.x == 0) actions_out[g] = actions[g];
}

__global__ void k_sp_advance(int g0, int n, int32_t *ply, const int8_t *status, unsigned long long *counters, uint32_t *dev_error) {
    const int g = g0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= g0 + n) return;
    const int st = status[g];
    if (st == kDraw || st == kBlackWin || st == kWhiteWin) {
        ply[g] = 0;
        atomicAdd(&counters[0], 1ull);
    } else if (st == kInProgress) {
        ply[g] += 1;
    } else {
        atomicOr(dev_error, 4u);  // Option::None: the sampled move was refused -- the game would sit at this ply for ever
    }
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
static TreeArgs make_args(omk_ctx *c, const int32_t *ids_dev, int n) {
    TreeArgs a{};
    a.hdrs = c->tree_hdrs;
    a.nodes = c->tree_nodes;
    a.cap_nodes = c->cap_nodes;
    a.ids = ids_dev;
    a.id_base = ids_dev ? 0 : c->id_base;
    a.n = n;
    a.seed = c->seed;
    a.dev_error = c->dev_error;
    a.dev_sims = c->dev_sims;
    a.nn_in = c->ws.nn_in;
    a.req_tree = c->ws.req_tree;
    a.req_node = c->ws.req_node;
    a.n_req = c->ws.n_req;
    a.slot_base = c->ws.slot_base;
    a.slot_count = c->ws.slot_count;
    a.P = c->ws.P;
    a.V = c->ws.V;
    a.max_rows = c->ws.max_rows;
    a.vloss = c->virtual_loss;
    return a;
}
static inline int warp_grid(int n) { return (n + kWarpsPerBlock - 1) / kWarpsPerBlock; }

void launch_reset_requests(omk_ctx *c) {
    k_reset_requests<<<1, 1, 0, c->stream>>>(c->ws.n_req, c->dev_sims);
    c->launches++;
}
void launch_new_games(omk_ctx *c, const int32_t *ids_dev, int n, const uint32_t *streams_dev,
                      const int8_t *only_if_terminal, const float *root_policy) {
    if (n <= 0) return;
    k_new_games<<<warp_grid(n), kWarpsPerBlock * 32, 0, c->stream>>>(make_args(c, ids_dev, n), streams_dev,
                                                                       only_if_terminal, root_policy);
    c->launches++;
}
void launch_sp_prepare(omk_ctx *c, int g0, int n, int32_t *mover, int32_t *other, uint8_t *modes, float *temps) {
    k_sp_prepare<<<(n + 127) / 128, 128, 0, c->stream>>>(g0, n, c->sp_ply, c->sp_cfg.temperature_threshold,
                                                          c->sp_cfg.temperature, mover, other, modes, temps);
    c->launches++;
}
void launch_sp_record(omk_ctx *c, int n, const int32_t *mover, const int32_t *actions, const float *policy_in,
                      uint8_t *boards_out, float *policy_out, int32_t *actions_out) {
    k_sp_record<<<n, 96, 0, c->stream>>>(make_args(c, mover, n), actions, policy_in, boards_out, policy_out, actions_out);
    c->launches++;
}
void launch_sp_advance(omk_ctx *c, int g0, int n, const int8_t *status, unsigned long long *counters) {
    k_sp_advance<<<(n + 127) / 128, 128, 0, c->stream>>>(g0, n, c->sp_ply, status, counters, c->dev_error);
    c->launches++;
}
void launch_root_noise(omk_ctx *c, const int32_t *ids_dev, int n, float epsilon, float alpha) {
    if (n <= 0) return;
    k_root_noise<<<warp_grid(n), kWarpsPerBlock * 32, 0, c->stream>>>(make_args(c, ids_dev, n), epsilon, alpha);
    c->launches++;
}
void launch_select_expand(omk_ctx *c, const int32_t *ids_dev, int n, int batch) {
    if (n <= 0) return;
    k_select_expand<<<warp_grid(n), kWarpsPerBlock * 32, 0, c->stream>>>(make_args(c, ids_dev, n), batch);
    c->launches++;
}
void launch_apply(omk_ctx *c, const int32_t *ids_dev, int n, int mode) {
    if (n <= 0) return;
    k_apply<<<warp_grid(n), kWarpsPerBlock * 32, 0, c->stream>>>(make_args(c, ids_dev, n), mode);
    c->launches++;
}
void launch_sample(omk_ctx *c, const int32_t *ids_dev, int n, const uint8_t *modes_dev, const float *temps_dev,
                   int32_t *actions_dev, float *policy_dev, uint8_t *valid_dev) {
    if (n <= 0) return;
    k_sample<<<warp_grid(n), kWarpsPerBlock * 32, 0, c->stream>>>(make_args(c, ids_dev, n), modes_dev, temps_dev,
                                                                    actions_dev, policy_dev, valid_dev);
    c->launches++;
}
void launch_ensure_prepare(omk_ctx *c, const int32_t *ids_dev, const int32_t *actions_dev, int n) {
    if (n <= 0) return;
    k_ensure_prepare<<<warp_grid(n), kWarpsPerBlock * 32, 0, c->stream>>>(make_args(c, ids_dev, n), actions_dev);
    c->launches++;
}
void launch_play(omk_ctx *c, const int32_t *ids_dev, const int32_t *actions_dev, int n, int8_t *status_dev) {
    if (n <= 0) return;
    k_play<<<warp_grid(n), kWarpsPerBlock * 32, 0, c->stream>>>(make_args(c, ids_dev, n), actions_dev, status_dev,
                                                                  c->remap);
    c->launches++;
}
void launch_root_children(omk_ctx *c, int tree, int32_t *actions_dev, unsigned long long *n_dev, float *w_dev,
                          float *p_dev, int32_t *len_dev, float *policy_dev, uint32_t *misc_dev) {
    k_root_children<<<1, 96, 0, c->stream>>>(make_args(c, nullptr, 1), tree, actions_dev, n_dev, w_dev, p_dev, len_dev,
                                              policy_dev, misc_dev);
    c->launches++;
}
void launch_pool_get_envs(omk_ctx *c, const int32_t *ids_dev, int n, uint8_t *boards_dev, uint8_t *turns_dev,
                          uint16_t *legal_dev, int8_t *status_dev) {
    if (n <= 0) return;
    k_pool_get_envs<<<n, 96, 0, c->stream>>>(make_args(c, ids_dev, n), boards_dev, turns_dev, legal_dev, status_dev);
    c->launches++;
}
void launch_eval_hash(omk_ctx *c, int rows_bound) {
    const int threads = 192;
    if (rows_bound > c->ws.max_rows) rows_bound = c->ws.max_rows;
    if (rows_bound <= 0) return;
    const long long total = (long long)rows_bound * kRow;
    k_eval_hash<<<(unsigned)((total + threads - 1) / threads), threads, 0, c->stream>>>(c->ws.nn_in, c->ws.n_req, c->ws.P,
                                                                                         c->ws.V, rows_bound);
    c->launches++;
}

}  // namespace omk

// omk_api.cu -- extern "C" boundary of libomok_b200.so (see include/omok_b200.h).
// Host-side orchestration only: pool allocation, argument staging, kernel sequencing.
// There is no CPU compute path: every operation is a CUDA kernel on the context's stream.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <utility>

#include "omk_internal.h"

using namespace omk;

static thread_local std::string g_err;
static int32_t fail(int32_t code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(OMK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));             \
    } while (0)
#define CKV(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) g_err = std::string(#call) + ": " + cudaGetErrorString(e__);            \
    } while (0)

static const long long kLens[kNetTensors] = {
    384, 128,
    4096, 32, 288, 1024, 32, 4096, 128,
    4096, 32, 288, 1024, 32, 4096, 128,
    4096, 32, 288, 1024, 32, 4096, 128,
    10368LL * 512, 512, 512 * 512, 512, 512, 1, 512 * 81, 81};


// every host <-> device copy of the API layer is counted in the context (omk_ctx_transfer_bytes; bench.py's e2e bytes)
static inline cudaError_t copy_h2d(omk_ctx *c, void *dst, const void *src, size_t bytes, cudaStream_t s) {
    c->h2d_bytes += (int64_t)bytes;
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
}
static inline cudaError_t copy_d2h(omk_ctx *c, void *dst, const void *src, size_t bytes, cudaStream_t s) {
    c->d2h_bytes += (int64_t)bytes;
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s);
}

extern "C" const char *omk_last_error(void) { return g_err.c_str(); }
extern "C" int32_t omk_version(void) { return 100; }

// ------------------------------------------------------------------ workspace
// grow-only device scratch slot i of the context (every call that uses it synchronises the stream before it returns)
template <typename T>
static int32_t scratch_dev(omk_ctx *c, int i, size_t count, T **out) {
    omk_ctx::Scratch &s = c->scratch[i];
    const size_t bytes = count * sizeof(T);
    if (bytes > s.cap) {
        CK(cudaStreamSynchronize(c->stream));
        cudaFree(s.p);
        s.p = nullptr;
        s.cap = 0;
        const size_t want = (bytes + 4095) & ~(size_t)4095;
        CK(cudaMalloc(&s.p, want));
        s.cap = want;
    }
    *out = reinterpret_cast<T *>(s.p);
    return OMK_OK;
}
#define SCRATCH(i, count, ptr)                          \
    do {                                                \
        int32_t rc__ = scratch_dev(c, (i), (count), &(ptr)); \
        if (rc__) return rc__;                          \
    } while (0)

static int32_t ensure_workspace(omk_ctx *c, int rows) {
    rows = (rows + 255) / 256 * 256;  // whole 256-row CTA-pair tiles
    if (rows <= c->ws.max_rows) return OMK_OK;
    CK(cudaStreamSynchronize(c->stream));
    Workspace &w = c->ws;
    cudaFree(w.nn_in); cudaFree(w.req_tree); cudaFree(w.req_node); cudaFree(w.P); cudaFree(w.V);
    w.nn_in = nullptr; w.req_tree = nullptr; w.req_node = nullptr; w.P = nullptr; w.V = nullptr;  // a failed cudaMalloc below must not leave dangling pointers for omk_ctx_destroy
    w.max_rows = 0;
    CK(cudaMalloc(&w.nn_in, sizeof(NNIn) * (size_t)rows));
    CK(cudaMalloc(&w.req_tree, sizeof(uint32_t) * (size_t)rows));
    CK(cudaMalloc(&w.req_node, sizeof(uint32_t) * (size_t)rows));
    CK(cudaMalloc(&w.P, sizeof(float) * (size_t)rows * kRow));
    CK(cudaMalloc(&w.V, sizeof(float) * (size_t)rows));
    CK(cudaMemsetAsync(w.nn_in, 0, sizeof(NNIn) * (size_t)rows, c->stream));
    w.max_rows = rows;
    return OMK_OK;
}

namespace omk {
// Activation buffers of the network, sized on first use (the hash evaluator never needs them): fp16 hi/lo operands of the
// tensor-core path always; the fp32 buffers of the CUDA-core A/B kernels only once such a mode has been selected.
bool ensure_activations(omk_ctx *c, int rows) {
    Workspace &w = c->ws;
    rows = (rows + 255) / 256 * 256;
    const bool want_fp32 = c->tower_mode == 0 || c->fc0_mode == 0;
    if (rows <= w.act_rows && (!want_fp32 || w.act_fp32)) return true;
    if (rows < w.act_rows) rows = w.act_rows;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return false;
    void **bufs[] = {(void **)&w.act0_h16, (void **)&w.act0_l16, (void **)&w.act1_h16, (void **)&w.act1_l16, (void **)&w.act2_h16,
                     (void **)&w.act2_l16, (void **)&w.act2, (void **)&w.logits, (void **)&w.act0, (void **)&w.act1};
    for (void **b : bufs) {
        cudaFree(*b);
        *b = nullptr;
    }
    w.act_rows = 0;
    w.act_fp32 = false;
    const size_t r = (size_t)rows;
    // (act0: two rows of slack -- k_tower16 stores whole position triples, the last one may reach past `rows`)
    const size_t sizes[] = {sizeof(__half) * (r + 2) * 10368, sizeof(__half) * (r + 2) * 10368, sizeof(__half) * r * 512, sizeof(__half) * r * 512,
                            sizeof(__half) * r * 512, sizeof(__half) * r * 512,
                            sizeof(float) * r * 512, sizeof(float) * r * 128, sizeof(float) * r * 10368, sizeof(float) * r * 512};
    for (int i = 0; i < (want_fp32 ? 10 : 8); ++i) {
        if (cudaMalloc(bufs[i], sizes[i]) != cudaSuccess) return false;
        if (cudaMemsetAsync(*bufs[i], 0, sizes[i], c->stream) != cudaSuccess) return false;  // padded tile rows read zeros, never NaNs
    }
    w.act_rows = rows;
    w.act_fp32 = want_fp32;
    return true;
}
}  // namespace omk

// ------------------------------------------------------------------ search lanes
// RAII: lane 1 = swap the context's stream and evaluator workspace with the second lane's for the launches in scope.
struct LaneScope {
    omk_ctx *c;
    bool swapped;
    LaneScope(omk_ctx *ctx, int lane, int id_base) : c(ctx), swapped(lane == 1) {
        if (swapped) {
            std::swap(c->stream, c->lane1_stream);
            std::swap(c->ws, c->lane1_ws);
        }
        c->id_base = id_base;
    }
    ~LaneScope() {
        if (swapped) {
            std::swap(c->stream, c->lane1_stream);
            std::swap(c->ws, c->lane1_ws);
        }
        c->id_base = 0;
    }
};
// how many lanes a search over n trees uses; creates lane 1 on first use
static int lanes_for(omk_ctx *c, int n) {
    if (c->lane_min_trees <= 0 || n < c->lane_min_trees || n < 2) return 1;
    if (!c->lane1_stream) {
        Workspace &w = c->lane1_ws;
        const size_t ct = (size_t)(c->cap_trees > 0 ? c->cap_trees : 1);
        if (cudaStreamCreateWithFlags(&c->lane1_stream, cudaStreamNonBlocking) != cudaSuccess) return 1;
        bool ok = cudaEventCreateWithFlags(&c->lane_fork, cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&c->lane_join, cudaEventDisableTiming) == cudaSuccess &&
                  cudaMalloc(&w.n_req, sizeof(uint32_t) * 4) == cudaSuccess &&
                  cudaMalloc(&w.slot_base, sizeof(uint32_t) * ct) == cudaSuccess &&
                  cudaMalloc(&w.slot_count, sizeof(uint32_t) * ct) == cudaSuccess &&
                  cudaMemsetAsync(w.n_req, 0, sizeof(uint32_t) * 4, c->lane1_stream) == cudaSuccess;
        if (!ok) {
            cudaGetLastError();
            c->lane_min_trees = 0;  // no second lane on this context
            return 1;
        }
    }
    return 2;
}
static void lanes_fork(omk_ctx *c) {  // lane 1 starts after everything already queued on the main stream
    cudaEventRecord(c->lane_fork, c->stream);
    cudaStreamWaitEvent(c->lane1_stream, c->lane_fork, 0);
}
static void lanes_join(omk_ctx *c) {  // the main stream continues after lane 1 has drained
    cudaEventRecord(c->lane_join, c->lane1_stream);
    cudaStreamWaitEvent(c->stream, c->lane_join, 0);
}

static int32_t check_device_error(omk_ctx *c) {
    uint32_t e = 0;
    CK(copy_d2h(c, &e, c->dev_error, sizeof e, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    if (e) {
        uint32_t z = 0;
        copy_h2d(c, c->dev_error, &z, sizeof z, c->stream);
    }
    if (e & 1u) return fail(OMK_ERR_CAPACITY, "a tree ran out of node slots (capacity_nodes=" + std::to_string(c->cap_nodes) + ")");
    if (e & 4u)
        return fail(OMK_ERR_STATE, "the self-play driver sampled a move its tree refused (play_action returned None): the transition stream of this call is not usable");
    if (e & 2u)
        return fail(OMK_ERR_NUMERIC, "the network produced a non-finite policy or value (activation beyond the fp16 operand range "
                                     "of the tensor-core path, |x| >= 65504, or non-finite weights)");
    return OMK_OK;
}

// stage a host id list (or the identity) on the device; returns nullptr for identity
static int32_t stage_ids(omk_ctx *c, const int32_t *ids, int n, const int32_t **out, int limit) {
    *out = nullptr;
    if (n < 0 || n > limit) return fail(OMK_ERR_INVALID, "n out of range for the pool capacity");
    if (!ids) return OMK_OK;
    // ids must be unique: the kernels give every listed slot its own lane / warp, and two of them on one record would
    // race (n_nodes bump, edge statistics, the in-place compaction of k_play)
    if ((int)c->id_seen.size() < limit) c->id_seen.assign((size_t)limit, 0);
    bool bad = false, dup = false;
    int filled = 0;
    for (; filled < n && !bad && !dup; ++filled) {
        const int32_t id = ids[filled];
        if (id < 0 || id >= limit) bad = true;
        else if (c->id_seen[(size_t)id]) dup = true;
        else c->id_seen[(size_t)id] = 1;
    }
    for (int i = 0; i < filled; ++i)
        if (ids[i] >= 0 && ids[i] < limit) c->id_seen[(size_t)ids[i]] = 0;
    if (bad) return fail(OMK_ERR_INVALID, "id out of range");
    if (dup) return fail(OMK_ERR_INVALID, "duplicate id in the id list (ids must be unique within one call)");
    CK(copy_h2d(c, c->ws.ids, ids, sizeof(int32_t) * (size_t)n, c->stream));
    *out = c->ws.ids;
    return OMK_OK;
}

namespace omk {
bool prof_begin(omk_ctx *c, int kind, int min_level) {
    if (c->prof_level < min_level) return false;
    omk_prof_span s;
    s.kind = kind;
    for (cudaEvent_t *e : {&s.a, &s.b}) {
        if (c->prof_pool.empty()) {
            cudaEventCreate(e);
        } else {
            *e = c->prof_pool.back();
            c->prof_pool.pop_back();
        }
    }
    cudaEventRecord(s.a, c->stream);
    c->prof_spans.push_back(s);
    return true;
}
void prof_end(omk_ctx *c, bool opened) {
    if (opened) cudaEventRecord(c->prof_spans.back().b, c->stream);
}
}  // namespace omk

// false: the evaluator did not run to the end of its launch sequence (ws.P / ws.V are stale): the caller returns
// OMK_ERR_CUDA WITHOUT launching k_apply, which would write those rows into tree policies and back up garbage values
static bool run_evaluator(omk_ctx *c, int evaluator, int rows_bound) {
    if (evaluator == OMK_EVAL_HASH) {
        const bool sp = prof_begin(c, OMK_K_HASH, 2);
        launch_eval_hash(c, rows_bound);
        prof_end(c, sp);
        return true;
    }
    return net_forward(c, nullptr, rows_bound);
}
#define RUN_EVAL(c, evaluator, rows)                                                                                   \
    do {                                                                                                               \
        if (!run_evaluator((c), (evaluator), (rows)))                                                                  \
            return fail(OMK_ERR_CUDA, "the evaluator failed to launch (see stderr); no result was applied to the trees"); \
    } while (0)

static int32_t weights_changed(omk_ctx *c);
// The trainer's 600 updates per iteration do not evaluate the network in between (the step's own forward passes read the fp32
// weights), so omk_train_apply only marks the tensor-core operand images stale; they are rebuilt here, on the main stream and
// before any search lane forks, by every entry point that is about to evaluate the network.
static int32_t ensure_packed(omk_ctx *c) {
    if (!c->net_pack_dirty) return OMK_OK;
    const int32_t rc = weights_changed(c);
    if (rc == OMK_OK) c->net_pack_dirty = false;  // a failed rebuild is retried (and reported again) by the next evaluation
    return rc;
}

static int32_t check_evaluator(omk_ctx *c, int evaluator) {
    if (evaluator != OMK_EVAL_NET && evaluator != OMK_EVAL_HASH) return fail(OMK_ERR_INVALID, "unknown evaluator");
    if (evaluator == OMK_EVAL_HASH && c->virtual_loss)
        return fail(OMK_ERR_STATE, "virtual loss is a non-reference search mode: it is refused with the parity evaluator (OMK_EVAL_HASH)");
    if (evaluator == OMK_EVAL_NET && !c->net.loaded) return fail(OMK_ERR_STATE, "network weights not loaded");
    if (evaluator == OMK_EVAL_NET) return ensure_packed(c);
    return OMK_OK;
}

// ------------------------------------------------------------------ context
extern "C" int32_t omk_ctx_create(int32_t device, int32_t capacity_envs, int32_t capacity_trees, int32_t capacity_nodes,
                                  uint64_t seed, omk_ctx **out) {
    if (!out) return fail(OMK_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (capacity_envs < 0 || capacity_trees < 0 || capacity_nodes < 2 || capacity_nodes > 65535)
        return fail(OMK_ERR_INVALID, "bad capacity (capacity_nodes must be in [2, 65535])");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(OMK_ERR_CUDA, "no CUDA device: libomok_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(OMK_ERR_INVALID, "device index out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(OMK_ERR_CUDA, "libomok_b200 is built for sm_100a (B200) only");

    omk_ctx *c = new omk_ctx();
    c->device = device;
    c->n_sms = prop.multiProcessorCount;
    c->cap_envs = capacity_envs;
    c->cap_trees = capacity_trees;
    c->cap_nodes = capacity_nodes;
    c->seed = seed;
    auto parse_mode = [](const char *m) { return strcmp(m, "simt") == 0 ? 0 : 1; };
    if (const char *m = getenv("OMK_FC0")) c->fc0_mode = parse_mode(m);
    if (const char *m = getenv("OMK_TOWER")) c->tower_mode = parse_mode(m);
    if (const char *m = getenv("OMK_LANE_MIN_TREES")) c->lane_min_trees = atoi(m);
    if (const char *m = getenv("OMK_FC0_CHUNK")) c->fc0_chunk = atoi(m) == 3 ? 3 : 9;
    if (const char *m = getenv("OMK_FC0_BALANCE")) c->fc0_balance = atoi(m) != 0;
    if (const char *m = getenv("OMK_SEARCH_VIRTUAL_LOSS")) c->virtual_loss = atoi(m) != 0;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    const int ce = capacity_envs > 0 ? capacity_envs : 1, ct = capacity_trees > 0 ? capacity_trees : 1;
    CK(cudaMalloc(&c->envs, sizeof(EnvRec) * (size_t)ce));
    CK(cudaMalloc(&c->tree_hdrs, sizeof(TreeHdr) * (size_t)ct));
    CK(cudaMalloc(&c->tree_nodes, (size_t)ct * (size_t)capacity_nodes * kNodeBytes));
    CK(cudaMalloc(&c->remap, sizeof(uint16_t) * (size_t)ct * (size_t)capacity_nodes));
    CK(cudaMalloc(&c->dev_error, sizeof(uint32_t)));
    CK(cudaMalloc(&c->dev_sims, sizeof(unsigned long long) * 2));
    CK(cudaMemsetAsync(c->dev_error, 0, sizeof(uint32_t), c->stream));
    CK(cudaMemsetAsync(c->dev_sims, 0, sizeof(unsigned long long) * 2, c->stream));
    CK(cudaMemsetAsync(c->tree_hdrs, 0, sizeof(TreeHdr) * (size_t)ct, c->stream));
    CK(cudaMemsetAsync(c->envs, 0, sizeof(EnvRec) * (size_t)ce, c->stream));
    Workspace &w = c->ws;
    const int cm = ce > ct ? ce : ct;
    CK(cudaMalloc(&w.n_req, sizeof(uint32_t) * 4));
    CK(cudaMemsetAsync(w.n_req, 0, sizeof(uint32_t) * 4, c->stream));
    CK(cudaMalloc(&w.slot_base, sizeof(uint32_t) * (size_t)ct));
    CK(cudaMalloc(&w.slot_count, sizeof(uint32_t) * (size_t)ct));
    CK(cudaMalloc(&w.ids, sizeof(int32_t) * (size_t)cm));
    CK(cudaMalloc(&w.actions, sizeof(int32_t) * (size_t)cm));
    CK(cudaMalloc(&w.modes, (size_t)cm));
    CK(cudaMalloc(&w.temps, sizeof(float) * (size_t)cm));
    CK(cudaMalloc(&w.status, (size_t)cm));
    CK(cudaMalloc(&w.policy_out, sizeof(float) * (size_t)ct * kCells));
    CK(cudaMalloc(&w.streams, sizeof(uint32_t) * (size_t)ct));
    for (int i = 0; i < kNetTensors; ++i) CK(cudaMalloc(&c->net.t[i], sizeof(float) * (size_t)kLens[i]));
    CK(cudaMalloc(&c->net.heads_w, sizeof(float) * 512 * 128));
    CK(cudaMalloc(&c->net.heads_b, sizeof(float) * 128));
    CK(cudaStreamSynchronize(c->stream));
    *out = c;
    return OMK_OK;
}

extern "C" int32_t omk_ctx_destroy(omk_ctx *c) {
    if (!c) return OMK_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    Workspace &w = c->ws;
    void *ptrs[] = {c->envs, c->tree_hdrs, c->tree_nodes, c->remap, c->dev_error, c->dev_sims, w.nn_in, w.req_tree,
                    w.req_node, w.P, w.V, w.act0, w.act1, w.act2, w.logits, w.n_req, w.slot_base, w.slot_count, w.ids,
                    w.actions, w.modes, w.temps, w.status, w.policy_out, w.streams, c->net.heads_w, c->net.heads_b,
                    c->sp_ply, c->sp_buf, w.act0_h16, w.act0_l16, w.act1_h16, w.act1_l16, w.act2_h16, w.act2_l16, c->net.heads_wt_h16, c->net.heads_wt_l16, c->net.fc0_wt_h16, c->net.fc0_wt_l16,
                    c->net.fc1_wt_h16, c->net.fc1_wt_l16, c->net.fc_inv_scale, c->net.fc_absmax, c->net.tower16_wimg,
                    c->net.tower16_pimg, c->net.tower16_absmax};
    fc16_free(c);
    train_free(c);
    for (omk_ctx::Scratch &sc : c->scratch) cudaFree(sc.p);
    free(c->tower16_params_host);
    if (c->lane1_stream) {
        cudaStreamSynchronize(c->lane1_stream);
        Workspace &l = c->lane1_ws;
        void *lp[] = {l.nn_in, l.req_tree, l.req_node, l.P, l.V, l.act0_h16, l.act0_l16, l.act1_h16, l.act1_l16, l.act2_h16, l.act2_l16,
                      l.act2, l.logits, l.act0, l.act1, l.n_req, l.slot_base, l.slot_count};
        for (void *q : lp) cudaFree(q);
        cudaEventDestroy(c->lane_fork);
        cudaEventDestroy(c->lane_join);
        cudaStreamDestroy(c->lane1_stream);
    }
    for (void *p : ptrs) cudaFree(p);
    for (int i = 0; i < kNetTensors; ++i) cudaFree(c->net.t[i]);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->sp_copy_stream) {
        cudaStreamSynchronize(c->sp_copy_stream);
        for (int i = 0; i < omk_ctx::kSpRing; ++i) {
            cudaEventDestroy(c->sp_ev_rec[i][0]);
            cudaEventDestroy(c->sp_ev_rec[i][1]);
            cudaEventDestroy(c->sp_ev_copy[i]);
        }
        cudaStreamDestroy(c->sp_copy_stream);
    }
    cudaFree(c->sp_ring_dev);
    if (c->sp_ring_host) cudaFreeHost(c->sp_ring_host);
    for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    delete c;
    return OMK_OK;
}

extern "C" int32_t omk_ctx_synchronize(omk_ctx *c) {
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return OMK_OK;
}
extern "C" void *omk_ctx_stream(omk_ctx *c) { return (void *)c->stream; }
extern "C" int64_t omk_ctx_launch_count(omk_ctx *c) { return c->launches; }
extern "C" int32_t omk_ctx_transfer_bytes(omk_ctx *c, int64_t *out_h2d, int64_t *out_d2h) {
    if (out_h2d) *out_h2d = c->h2d_bytes;
    if (out_d2h) *out_d2h = c->d2h_bytes;
    return OMK_OK;
}

// ------------------------------------------------------------------ network
static int32_t sp_refresh_root_policy(omk_ctx *c);  // self-play driver: re-evaluate the cached empty-board prior

extern "C" int32_t omk_net_load_params(omk_ctx *c, const float *const *tensors, const int64_t *lens) {
    CK(cudaSetDevice(c->device));
    for (int i = 0; i < kNetTensors; ++i)
        if (!tensors[i] || lens[i] != kLens[i])
            return fail(OMK_ERR_INVALID, "tensor " + std::to_string(i) + ": expected " + std::to_string(kLens[i]) + " elements");
    for (int i = 0; i < kNetTensors; ++i)
        CK(copy_h2d(c, c->net.t[i], tensors[i], sizeof(float) * (size_t)kLens[i], c->stream));
    c->net_pack_dirty = false;  // rebuilt right here
    net_pack_heads(c);
    if (!fc16_prepare_weights(c)) return fail(OMK_ERR_CUDA, "fp16-split fc weight preparation failed");
    if (!tower16_prepare_weights(c)) return fail(OMK_ERR_CUDA, "fp16-split tower weight preparation failed");
    CK(cudaStreamSynchronize(c->stream));
    c->net.loaded = true;
    return sp_refresh_root_policy(c);
}
extern "C" int32_t omk_net_get_params(omk_ctx *c, float *const *tensors, const int64_t *lens) {
    CK(cudaSetDevice(c->device));
    if (!c->net.loaded) return fail(OMK_ERR_STATE, "network weights not loaded");
    for (int i = 0; i < kNetTensors; ++i)
        if (!tensors[i] || lens[i] != kLens[i]) return fail(OMK_ERR_INVALID, "bad tensor length");
    for (int i = 0; i < kNetTensors; ++i)
        CK(copy_d2h(c, tensors[i], c->net.t[i], sizeof(float) * (size_t)kLens[i], c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OMK_OK;
}
extern "C" int32_t omk_net_init_random(omk_ctx *c, uint64_t seed) {
    CK(cudaSetDevice(c->device));
    launch_net_init_random(c, seed);
    c->net_pack_dirty = false;  // rebuilt right here
    net_pack_heads(c);
    if (!fc16_prepare_weights(c)) return fail(OMK_ERR_CUDA, "fp16-split fc weight preparation failed");
    if (!tower16_prepare_weights(c)) return fail(OMK_ERR_CUDA, "fp16-split tower weight preparation failed");
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    c->net.loaded = true;
    return sp_refresh_root_policy(c);
}

static int32_t net_eval_common(omk_ctx *c, int n, const float *images_dev, float *out_p, float *out_v) {
    int32_t rc_p = ensure_packed(c);
    if (rc_p) return rc_p;
    if (!net_forward(c, images_dev, n)) return fail(OMK_ERR_CUDA, "the network forward failed to launch (see stderr)");
    // P rows are padded to 96 floats on the device; compact on the way out
    CK(cudaMemcpy2DAsync(out_p, sizeof(float) * kCells, c->ws.P, sizeof(float) * kRow, sizeof(float) * kCells, (size_t)n,
                         cudaMemcpyDeviceToHost, c->stream));
    c->d2h_bytes += (int64_t)sizeof(float) * kCells * n;
    if (out_v) CK(copy_d2h(c, out_v, c->ws.V, sizeof(float) * (size_t)n, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return check_device_error(c);  // OMK_ERR_NUMERIC when the network produced a non-finite output
}

extern "C" int32_t omk_net_eval(omk_ctx *c, const uint8_t *boards, const uint8_t *turns, int32_t n, int32_t mode,
                                float *out_p, float *out_v) {
    CK(cudaSetDevice(c->device));
    if (!c->net.loaded) return fail(OMK_ERR_STATE, "network weights not loaded");
    if (n < 0 || !boards || !turns || !out_p) return fail(OMK_ERR_INVALID, "bad arguments");
    if (n == 0) return OMK_OK;
    int32_t rc = ensure_workspace(c, n);
    if (rc) return rc;
    uint8_t *d_boards = nullptr, *d_turns = nullptr;
    SCRATCH(0, (size_t)n * kCells, d_boards);
    SCRATCH(1, (size_t)n, d_turns);
    CK(copy_h2d(c, d_boards, boards, (size_t)n * kCells, c->stream));
    CK(copy_h2d(c, d_turns, turns, (size_t)n, c->stream));
    launch_pack_boards(c, d_boards, d_turns, n, mode);
    return net_eval_common(c, n, nullptr, out_p, out_v);
}

extern "C" int32_t omk_net_eval_images(omk_ctx *c, const float *images, int32_t n, float *out_p, float *out_v) {
    CK(cudaSetDevice(c->device));
    if (!c->net.loaded) return fail(OMK_ERR_STATE, "network weights not loaded");
    if (n < 0 || !images || !out_p) return fail(OMK_ERR_INVALID, "bad arguments");
    if (n == 0) return OMK_OK;
    int32_t rc = ensure_workspace(c, n);
    if (rc) return rc;
    float *d_img = nullptr;
    SCRATCH(0, (size_t)n * 243, d_img);
    CK(copy_h2d(c, d_img, images, sizeof(float) * (size_t)n * 243, c->stream));
    const uint32_t nn = (uint32_t)n;
    CK(copy_h2d(c, c->ws.n_req, &nn, sizeof nn, c->stream));
    return net_eval_common(c, n, d_img, out_p, out_v);
}

// ------------------------------------------------------------------ trainer step (AgentModel::train)
// after the optimizer changed the fp32 tensors: re-pack the heads, re-split the tensor-core operands, refresh the
// self-play driver's cached root prior -- exactly what omk_net_load_params does after its copies
static int32_t weights_changed(omk_ctx *c) {
    net_pack_heads(c);
    if (!fc16_prepare_weights(c)) return fail(OMK_ERR_CUDA, "fp16-split fc weight preparation failed");
    if (!tower16_prepare_weights(c)) return fail(OMK_ERR_CUDA, "fp16-split tower weight preparation failed");
    CK(cudaStreamSynchronize(c->stream));
    return sp_refresh_root_policy(c);
}

static int32_t train_backward_impl(omk_ctx *c, const float *images, const float *pi, const float *z, int32_t n,
                                   void **out_grads_device, int64_t *out_count, bool sync) {
    CK(cudaSetDevice(c->device));
    if (!c->net.loaded) return fail(OMK_ERR_STATE, "network weights not loaded");
    if (n < 1 || !images || !pi || !z) return fail(OMK_ERR_INVALID, "bad arguments");
    float *g = nullptr;
    c->train_n = 0;
    if (const char *e = train_backward_step(c, images, pi, z, n, &g)) return fail(OMK_ERR_CUDA, std::string("omk_train_backward: ") + e);
    if (sync) {  // the caller may all-reduce the gradient buffer on its own stream (the fused step stays on this one)
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaGetLastError());
    }
    c->train_n = n;
    if (out_grads_device) *out_grads_device = g;
    if (out_count) *out_count = 5643250;
    return OMK_OK;
}

extern "C" int32_t omk_train_backward(omk_ctx *c, const float *images, const float *pi, const float *z, int32_t n,
                                      void **out_grads_device, int64_t *out_count) {
    return train_backward_impl(c, images, pi, z, n, out_grads_device, out_count, true);
}

extern "C" int32_t omk_train_apply(omk_ctx *c, float *out_losses) {
    CK(cudaSetDevice(c->device));
    if (c->train_n < 1) return fail(OMK_ERR_STATE, "omk_train_backward has not produced a gradient");
    if (const char *e = train_apply_step(c)) return fail(OMK_ERR_CUDA, std::string("omk_train_apply: ") + e);
    c->net_pack_dirty = true;  // the operand images of the tensor-core path follow before the next evaluation (ensure_packed)
    float l[3] = {0, 0, 0};
    if (const char *e = train_report_losses(c, c->train_n, l)) return fail(OMK_ERR_CUDA, std::string("omk_train_apply: ") + e);
    c->train_n = 0;  // one gradient, one update
    if (!(l[2] == l[2]) || l[2] > 3.0e38f) return fail(OMK_ERR_NUMERIC, "the training loss is not finite");
    if (out_losses) memcpy(out_losses, l, sizeof l);
    return OMK_OK;
}

extern "C" int32_t omk_train_step(omk_ctx *c, const float *images, const float *pi, const float *z, int32_t n, float *out_losses) {
    int32_t rc = train_backward_impl(c, images, pi, z, n, nullptr, nullptr, false);
    if (rc) return rc;
    return omk_train_apply(c, out_losses);
}

extern "C" int32_t omk_train_get_grads(omk_ctx *c, float *const *tensors, const int64_t *lens) {
    CK(cudaSetDevice(c->device));
    const float *g = train_grad_buffer(c);
    if (!g) return fail(OMK_ERR_STATE, "omk_train_backward has not run");
    long long off = 0;
    for (int i = 0; i < kNetTensors; ++i) {
        if (!tensors[i] || lens[i] != kLens[i]) return fail(OMK_ERR_INVALID, "bad tensor length");
        CK(copy_d2h(c, tensors[i], g + off, sizeof(float) * (size_t)kLens[i], c->stream));
        off += kLens[i];
    }
    CK(cudaStreamSynchronize(c->stream));
    return OMK_OK;
}

extern "C" int32_t omk_train_reset_optimizer(omk_ctx *c) {
    CK(cudaSetDevice(c->device));
    train_reset_optimizer(c);
    CK(cudaStreamSynchronize(c->stream));
    return OMK_OK;
}

extern "C" int32_t omk_train_comm_unique_id(omk_ctx *c, uint8_t *out_id) {
    if (!out_id) return fail(OMK_ERR_INVALID, "out_id is NULL");
    if (const char *e = train_comm_unique_id(c, out_id)) return fail(OMK_ERR_STATE, std::string("NCCL: ") + e);
    return OMK_OK;
}
extern "C" int32_t omk_train_comm_init(omk_ctx *c, const uint8_t *id, int32_t nranks, int32_t rank) {
    CK(cudaSetDevice(c->device));
    if (!id || nranks < 1 || rank < 0 || rank >= nranks) return fail(OMK_ERR_INVALID, "bad communicator arguments");
    if (const char *e = train_comm_init(c, id, nranks, rank)) return fail(OMK_ERR_STATE, std::string("NCCL: ") + e);
    return OMK_OK;
}
extern "C" int32_t omk_train_comm_destroy(omk_ctx *c) {
    CK(cudaSetDevice(c->device));
    train_comm_destroy(c);
    return OMK_OK;
}

// ------------------------------------------------------------------ diagnostics
extern "C" int32_t omk_debug_set_fc0_mode(omk_ctx *c, int32_t mode) {
    if (mode != 0 && mode != 1) return fail(OMK_ERR_INVALID, "fc0 mode must be 1 (tcgen05 3xFP16) or 0 (fp32 CUDA cores, A/B check)");
    c->fc0_mode = mode;
    return OMK_OK;
}
extern "C" int32_t omk_debug_set_tower_mode(omk_ctx *c, int32_t mode) {
    if (mode != 0 && mode != 1) return fail(OMK_ERR_INVALID, "tower mode must be 1 (tcgen05 3xFP16) or 0 (fp32 CUDA cores, A/B check)");
    c->tower_mode = mode;
    return OMK_OK;
}
extern "C" int32_t omk_debug_set_fc0_balance(omk_ctx *c, int32_t on) {
    if (on != 0 && on != 1) return fail(OMK_ERR_INVALID, "fc0 balance must be 0 or 1");
    c->fc0_balance = on;
    return OMK_OK;
}
extern "C" int32_t omk_debug_set_fc0_chunk(omk_ctx *c, int32_t k_blocks) {
    if (k_blocks != 3 && k_blocks != 9) return fail(OMK_ERR_INVALID, "fc0 chunk must be 9 (default) or 3 (finer accumulation)");
    CK(cudaStreamSynchronize(c->stream));
    c->fc0_chunk = k_blocks;
    return OMK_OK;
}
extern "C" int32_t omk_debug_set_lane_min_trees(omk_ctx *c, int32_t min_trees) {
    if (min_trees < 0) return fail(OMK_ERR_INVALID, "min_trees < 0");
    c->lane_min_trees = min_trees;
    return OMK_OK;
}
extern "C" int32_t omk_debug_tower_timing(omk_ctx *c, int64_t *out64) {
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    tower16_read_timing(reinterpret_cast<long long *>(out64));
    return OMK_OK;
}
extern "C" int32_t omk_debug_get_buffer(omk_ctx *c, int32_t which, float *out, int64_t count) {
    CK(cudaSetDevice(c->device));
    if (which >= 8 && which <= 10) {  // fp16-split buffers, reconstructed as hi + lo: 8 = tower, 9 = fc0, 10 = fc1 output
        if (!out || count < 0) return fail(OMK_ERR_INVALID, "bad buffer");
        const __half *hi = which == 8 ? c->ws.act0_h16 : which == 9 ? c->ws.act1_h16 : c->ws.act2_h16;
        const __half *lo = which == 8 ? c->ws.act0_l16 : which == 9 ? c->ws.act1_l16 : c->ws.act2_l16;
        if (!hi || !lo) return fail(OMK_ERR_STATE, "the tensor-core activation buffers do not exist yet (no network call so far)");
        float *tmp = nullptr;
        SCRATCH(0, (size_t)(count > 0 ? count : 1), tmp);
        launch_split16_to_f32(c, hi, lo, tmp, count);
        CK(copy_d2h(c, out, tmp, sizeof(float) * (size_t)count, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return OMK_OK;
    }
    const float *src = which == 0 ? c->ws.act0 : which == 1 ? c->ws.act1 : which == 2 ? c->ws.act2 : which == 3 ? c->ws.logits : nullptr;
    if (!src || !out || count < 0) return fail(OMK_ERR_INVALID, "bad buffer id");
    CK(copy_d2h(c, out, src, sizeof(float) * (size_t)count, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OMK_OK;
}

// ------------------------------------------------------------------ environment
extern "C" int32_t omk_env_reset(omk_ctx *c, const int32_t *ids, int32_t n) {
    CK(cudaSetDevice(c->device));
    const int32_t *d_ids;
    int32_t rc = stage_ids(c, ids, n, &d_ids, c->cap_envs);
    if (rc) return rc;
    launch_env_reset(c, d_ids, n);
    CK(cudaStreamSynchronize(c->stream));
    return OMK_OK;
}

extern "C" int32_t omk_env_step(omk_ctx *c, const int32_t *ids, const uint8_t *actions, int32_t n, int8_t *out_status,
                                uint32_t *out_legal) {
    CK(cudaSetDevice(c->device));
    if (!actions) return fail(OMK_ERR_INVALID, "actions is NULL");
    const int32_t *d_ids;
    int32_t rc = stage_ids(c, ids, n, &d_ids, c->cap_envs);
    if (rc) return rc;
    if (n == 0) return OMK_OK;
    uint8_t *d_act = nullptr;
    int8_t *d_st = nullptr;
    uint32_t *d_legal = nullptr;
    SCRATCH(0, (size_t)n, d_act);
    SCRATCH(1, (size_t)n, d_st);
    if (out_legal) SCRATCH(2, 3 * (size_t)n, d_legal);
    CK(copy_h2d(c, d_act, actions, (size_t)n, c->stream));
    launch_env_step(c, d_ids, d_act, n, d_st, d_legal);
    if (out_status) CK(copy_d2h(c, out_status, d_st, (size_t)n, c->stream));
    if (out_legal) CK(copy_d2h(c, out_legal, d_legal, sizeof(uint32_t) * 3 * (size_t)n, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return OMK_OK;
}

extern "C" int32_t omk_env_step_device(omk_ctx *c, const uint8_t *actions_device, int32_t n, int8_t *out_status_device,
                                       uint32_t *out_legal_device) {
    if (n < 0 || n > c->cap_envs) return fail(OMK_ERR_INVALID, "n exceeds capacity_envs");
    launch_env_step(c, nullptr, actions_device, n, out_status_device, out_legal_device);
    return OMK_OK;
}

extern "C" int32_t omk_env_get(omk_ctx *c, const int32_t *ids, int32_t n, uint8_t *out_boards, uint8_t *out_turns,
                               uint16_t *out_legal_counts) {
    CK(cudaSetDevice(c->device));
    const int32_t *d_ids;
    int32_t rc = stage_ids(c, ids, n, &d_ids, c->cap_envs);
    if (rc) return rc;
    if (n == 0) return OMK_OK;
    uint8_t *d_b = nullptr, *d_t = nullptr;
    uint16_t *d_l = nullptr;
    SCRATCH(0, (size_t)n * kCells, d_b);
    SCRATCH(1, (size_t)n, d_t);
    SCRATCH(2, (size_t)n, d_l);
    launch_env_get(c, d_ids, n, d_b, d_t, d_l);
    if (out_boards) CK(copy_d2h(c, out_boards, d_b, (size_t)n * kCells, c->stream));
    if (out_turns) CK(copy_d2h(c, out_turns, d_t, (size_t)n, c->stream));
    if (out_legal_counts) CK(copy_d2h(c, out_legal_counts, d_l, sizeof(uint16_t) * (size_t)n, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OMK_OK;
}

extern "C" int32_t omk_env_set(omk_ctx *c, const int32_t *ids, int32_t n, const uint8_t *boards, const uint8_t *turns) {
    CK(cudaSetDevice(c->device));
    if (!boards || !turns) return fail(OMK_ERR_INVALID, "NULL argument");
    const int32_t *d_ids;
    int32_t rc = stage_ids(c, ids, n, &d_ids, c->cap_envs);
    if (rc) return rc;
    if (n == 0) return OMK_OK;
    uint8_t *d_b = nullptr, *d_t = nullptr;
    SCRATCH(0, (size_t)n * kCells, d_b);
    SCRATCH(1, (size_t)n, d_t);
    CK(copy_h2d(c, d_b, boards, (size_t)n * kCells, c->stream));
    CK(copy_h2d(c, d_t, turns, (size_t)n, c->stream));
    launch_env_set(c, d_ids, n, d_b, d_t);
    CK(cudaStreamSynchronize(c->stream));
    return OMK_OK;
}

extern "C" int32_t omk_env_encode(omk_ctx *c, const int32_t *ids, int32_t n, int32_t mode, float *out) {
    CK(cudaSetDevice(c->device));
    if (n == 0) return OMK_OK;
    if (!out) return fail(OMK_ERR_INVALID, "out is NULL");
    const int32_t *d_ids;
    int32_t rc = stage_ids(c, ids, n, &d_ids, c->cap_envs);
    if (rc) return rc;
    if (n == 0) return OMK_OK;
    float *d_o = nullptr;
    SCRATCH(0, 243 * (size_t)n, d_o);
    launch_env_encode(c, d_ids, n, mode, d_o);
    CK(copy_d2h(c, out, d_o, sizeof(float) * 243 * (size_t)n, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OMK_OK;
}

extern "C" int32_t omk_env_random_playout(omk_ctx *c, int32_t n, int32_t plies, uint8_t *out_actions, int8_t *out_status) {
    CK(cudaSetDevice(c->device));
    if (n < 0 || n > c->cap_envs || plies < 0) return fail(OMK_ERR_INVALID, "bad arguments");
    if (n == 0 || plies == 0) return OMK_OK;
    uint8_t *d_a = nullptr;
    int8_t *d_s = nullptr;
    const size_t tot = (size_t)n * (size_t)plies;
    if (out_actions) SCRATCH(0, tot, d_a);
    if (out_status) SCRATCH(1, tot, d_s);
    launch_env_playout(c, n, plies, d_a, d_s);
    if (out_actions) CK(copy_d2h(c, out_actions, d_a, tot, c->stream));
    if (out_status) CK(copy_d2h(c, out_status, d_s, tot, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return OMK_OK;
}

// ------------------------------------------------------------------ tree pool
extern "C" int32_t omk_pool_new_games(omk_ctx *c, const int32_t *ids, int32_t n, const uint32_t *streams, int32_t evaluator) {
    CK(cudaSetDevice(c->device));
    int32_t rc = check_evaluator(c, evaluator);
    if (rc) return rc;
    const int32_t *d_ids;
    rc = stage_ids(c, ids, n, &d_ids, c->cap_trees);
    if (rc) return rc;
    if (n == 0) return OMK_OK;
    rc = ensure_workspace(c, n);
    if (rc) return rc;
    const uint32_t *d_streams = nullptr;
    if (streams) {
        CK(copy_h2d(c, c->ws.streams, streams, sizeof(uint32_t) * (size_t)n, c->stream));
        d_streams = c->ws.streams;
    }
    launch_reset_requests(c);
    launch_new_games(c, d_ids, n, d_streams, nullptr, nullptr);
    RUN_EVAL(c, evaluator, n);
    launch_apply(c, d_ids, n, kApplyNewGame);
    return check_device_error(c);
}

extern "C" int32_t omk_pool_search(omk_ctx *c, const int32_t *ids, int32_t n, int32_t count, int32_t batch_size, float epsilon,
                                   float alpha, int32_t evaluator) {
    CK(cudaSetDevice(c->device));
    int32_t rc = check_evaluator(c, evaluator);
    if (rc) return rc;
    if (batch_size < 1 || batch_size > kMaxBatchPerTree) return fail(OMK_ERR_INVALID, "batch_size must be in [1, 64]");
    if (count < 0) return fail(OMK_ERR_INVALID, "count < 0");
    const int32_t *d_ids;
    rc = stage_ids(c, ids, n, &d_ids, c->cap_trees);
    if (rc) return rc;
    if (n == 0 || count == 0) return OMK_OK;
    const int rounds = (count + batch_size - 1) / batch_size;  // parallel_mcts_executor.rs:39-42,207
    // large searches run as two lanes (halves of the id list) on two streams: per-tree results do not depend on the split
    const int lanes = lanes_for(c, n);
    const int lane_n[2] = {lanes == 2 ? (n + 1) / 2 : n, lanes == 2 ? n / 2 : 0};
    const int lane_first[2] = {0, lane_n[0]};
    for (int l = 0; l < lanes; ++l) {
        LaneScope scope(c, l, 0);
        rc = ensure_workspace(c, lane_n[l] * batch_size);
        if (rc) return rc;
    }
    if (lanes == 2) lanes_fork(c);
    for (int r = 0; r < rounds; ++r) {
        for (int l = 0; l < lanes; ++l) {
            LaneScope scope(c, l, lane_first[l]);
            const int32_t *ids_l = d_ids ? d_ids + lane_first[l] : nullptr;
            if (r == 0) launch_root_noise(c, ids_l, lane_n[l], epsilon, alpha);
            launch_reset_requests(c);
            launch_select_expand(c, ids_l, lane_n[l], batch_size);
            RUN_EVAL(c, evaluator, lane_n[l] * batch_size);
            launch_apply(c, ids_l, lane_n[l], kApplySearch);
        }
    }
    if (lanes == 2) lanes_join(c);
    return check_device_error(c);
}

extern "C" int32_t omk_search_set_virtual_loss(omk_ctx *c, int32_t enabled) {
    CK(cudaStreamSynchronize(c->stream));
    c->virtual_loss = enabled != 0;
    return OMK_OK;
}

extern "C" int32_t omk_pool_sample(omk_ctx *c, const int32_t *ids, int32_t n, const uint8_t *modes, const float *temperatures,
                                   int32_t *out_actions, float *out_policy) {
    CK(cudaSetDevice(c->device));
    const int32_t *d_ids;
    int32_t rc = stage_ids(c, ids, n, &d_ids, c->cap_trees);
    if (rc) return rc;
    if (n == 0) return OMK_OK;
    const uint8_t *d_modes = nullptr;
    const float *d_temps = nullptr;
    if (modes) {
        bool any_boltz = false;
        for (int i = 0; i < n; ++i) any_boltz |= modes[i] == OMK_SAMPLE_BOLTZMANN;
        if (any_boltz && !temperatures) return fail(OMK_ERR_INVALID, "Boltzmann sampling needs temperatures");
        CK(copy_h2d(c, c->ws.modes, modes, (size_t)n, c->stream));
        d_modes = c->ws.modes;
        if (temperatures) {
            CK(copy_h2d(c, c->ws.temps, temperatures, sizeof(float) * (size_t)n, c->stream));
            d_temps = c->ws.temps;
        }
    }
    launch_sample(c, d_ids, n, d_modes, d_temps, c->ws.actions, c->ws.policy_out, nullptr);
    if (out_actions) CK(copy_d2h(c, out_actions, c->ws.actions, sizeof(int32_t) * (size_t)n, c->stream));
    if (out_policy) CK(copy_d2h(c, out_policy, c->ws.policy_out, sizeof(float) * kCells * (size_t)n, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return OMK_OK;
}

extern "C" int32_t omk_pool_policy(omk_ctx *c, const int32_t *ids, int32_t n, float *out_policy, uint8_t *out_valid) {
    CK(cudaSetDevice(c->device));
    const int32_t *d_ids;
    int32_t rc = stage_ids(c, ids, n, &d_ids, c->cap_trees);
    if (rc) return rc;
    if (n == 0) return OMK_OK;
    launch_sample(c, d_ids, n, nullptr, nullptr, nullptr, c->ws.policy_out, c->ws.modes);
    if (out_policy) CK(copy_d2h(c, out_policy, c->ws.policy_out, sizeof(float) * kCells * (size_t)n, c->stream));
    if (out_valid) CK(copy_d2h(c, out_valid, c->ws.modes, (size_t)n, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return OMK_OK;
}

extern "C" int32_t omk_pool_ensure_action(omk_ctx *c, const int32_t *ids, const int32_t *actions, int32_t n, int32_t evaluator) {
    CK(cudaSetDevice(c->device));
    int32_t rc = check_evaluator(c, evaluator);
    if (rc) return rc;
    if (!actions) return fail(OMK_ERR_INVALID, "actions is NULL");
    const int32_t *d_ids;
    rc = stage_ids(c, ids, n, &d_ids, c->cap_trees);
    if (rc) return rc;
    if (n == 0) return OMK_OK;
    rc = ensure_workspace(c, n);
    if (rc) return rc;
    CK(copy_h2d(c, c->ws.actions, actions, sizeof(int32_t) * (size_t)n, c->stream));
    launch_reset_requests(c);
    launch_ensure_prepare(c, d_ids, c->ws.actions, n);
    RUN_EVAL(c, evaluator, n);
    launch_apply(c, d_ids, n, kApplyEnsure);
    return check_device_error(c);
}

extern "C" int32_t omk_pool_play(omk_ctx *c, const int32_t *ids, const int32_t *actions, int32_t n, int8_t *out_status) {
    CK(cudaSetDevice(c->device));
    if (!actions) return fail(OMK_ERR_INVALID, "actions is NULL");
    const int32_t *d_ids;
    int32_t rc = stage_ids(c, ids, n, &d_ids, c->cap_trees);
    if (rc) return rc;
    if (n == 0) return OMK_OK;
    CK(copy_h2d(c, c->ws.actions, actions, sizeof(int32_t) * (size_t)n, c->stream));
    launch_play(c, d_ids, c->ws.actions, n, c->ws.status);
    if (out_status) CK(copy_d2h(c, out_status, c->ws.status, (size_t)n, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return OMK_OK;
}

struct RootDump {
    int32_t actions[kCells];
    unsigned long long n[kCells];
    float w[kCells], p[kCells];
    int32_t len;
    float policy[kCells];
    uint32_t misc[16];
};

static int32_t dump_root(omk_ctx *c, int32_t id, RootDump *host) {
    CK(cudaSetDevice(c->device));
    if (id < 0 || id >= c->cap_trees) return fail(OMK_ERR_INVALID, "tree id out of range");
    RootDump *d = nullptr;
    SCRATCH(0, 1, d);
    launch_root_children(c, id, d->actions, d->n, d->w, d->p, &d->len, d->policy, d->misc);
    CK(copy_d2h(c, host, d, sizeof(RootDump), c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return OMK_OK;
}

extern "C" int32_t omk_pool_root_children(omk_ctx *c, int32_t id, int32_t *out_actions, uint64_t *out_n, float *out_w,
                                          float *out_p, int32_t *out_len) {
    RootDump h;
    int32_t rc = dump_root(c, id, &h);
    if (rc) return rc;
    for (int i = 0; i < h.len; ++i) {
        if (out_actions) out_actions[i] = h.actions[i];
        if (out_n) out_n[i] = h.n[i];
        if (out_w) out_w[i] = h.w[i];
        if (out_p) out_p[i] = h.p[i];
    }
    if (out_len) *out_len = h.len;
    return OMK_OK;
}

extern "C" int32_t omk_pool_root_stats(omk_ctx *c, int32_t id, uint64_t *out_n, float *out_w, float *out_p, int32_t *out_status,
                                       float *out_policy) {
    RootDump h;
    int32_t rc = dump_root(c, id, &h);
    if (rc) return rc;
    if (out_n) *out_n = h.misc[0];
    if (out_w) memcpy(out_w, &h.misc[1], 4);
    if (out_p) memcpy(out_p, &h.misc[2], 4);
    if (out_status) *out_status = (int32_t)h.misc[3];
    if (out_policy) memcpy(out_policy, h.policy, sizeof h.policy);
    return OMK_OK;
}

extern "C" int32_t omk_pool_tree_info(omk_ctx *c, int32_t id, int32_t *out_nodes, uint32_t *out_rng_counter) {
    RootDump h;
    int32_t rc = dump_root(c, id, &h);
    if (rc) return rc;
    if (out_nodes) *out_nodes = (int32_t)h.misc[4];
    if (out_rng_counter) *out_rng_counter = h.misc[5];
    return OMK_OK;
}

extern "C" int32_t omk_pool_get_env(omk_ctx *c, int32_t id, uint8_t *out_board, uint8_t *out_turn, uint16_t *out_legal_count) {
    RootDump h;
    int32_t rc = dump_root(c, id, &h);
    if (rc) return rc;
    if (out_board)
        for (int i = 0; i < kCells; ++i) {
            const bool b = (h.misc[8 + (i >> 5)] >> (i & 31)) & 1u, w = (h.misc[11 + (i >> 5)] >> (i & 31)) & 1u;
            out_board[i] = b ? 1 : (w ? 2 : 0);
        }
    if (out_turn) *out_turn = (uint8_t)(h.misc[6] & 0xFF);
    if (out_legal_count) *out_legal_count = (uint16_t)((h.misc[6] >> 8) & 0xFF);
    return OMK_OK;
}

extern "C" int32_t omk_pool_get_envs(omk_ctx *c, const int32_t *ids, int32_t n, uint8_t *out_boards, uint8_t *out_turns,
                                     uint16_t *out_legal_counts, int8_t *out_status) {
    CK(cudaSetDevice(c->device));
    const int32_t *d_ids;
    int32_t rc = stage_ids(c, ids, n, &d_ids, c->cap_trees);
    if (rc) return rc;
    if (n == 0) return OMK_OK;
    // one scratch block: [boards n*81 | turns n | status n | legal n u16]
    uint8_t *blk = nullptr;
    const size_t nn = (size_t)n, off_t = nn * kCells, off_s = off_t + nn, off_l = (off_s + nn + 1) & ~(size_t)1;
    SCRATCH(0, off_l + 2 * nn, blk);
    launch_pool_get_envs(c, d_ids, n, blk, blk + off_t, reinterpret_cast<uint16_t *>(blk + off_l), reinterpret_cast<int8_t *>(blk + off_s));
    if (out_boards) CK(copy_d2h(c, out_boards, blk, nn * kCells, c->stream));
    if (out_turns) CK(copy_d2h(c, out_turns, blk + off_t, nn, c->stream));
    if (out_status) CK(copy_d2h(c, out_status, blk + off_s, nn, c->stream));
    if (out_legal_counts) CK(copy_d2h(c, out_legal_counts, blk + off_l, 2 * nn, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return OMK_OK;
}

// ------------------------------------------------------------------ self-play driver
struct SpLayout {
    int32_t *mover, *other, *actions;
    float *temps;
    uint8_t *modes;
    int8_t *status, *status2;
    float *root_policy;
    unsigned long long *counters;  // [0] games finished
    size_t bytes;
};
// one ply slot of the transition ring: [boards n*81 u8 | policy n*81 f32 | actions n i32 | status n i8], 256-byte aligned parts
struct SpSlot {
    size_t off_boards, off_policy, off_actions, off_status, bytes;
};
static SpSlot sp_slot_layout(int n) {
    SpSlot s;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    s.off_boards = 0;
    s.off_policy = up((size_t)n * kCells);
    s.off_actions = s.off_policy + up((size_t)n * kCells * sizeof(float));
    s.off_status = s.off_actions + up((size_t)n * sizeof(int32_t));
    s.bytes = s.off_status + up((size_t)n);
    return s;
}
static SpLayout sp_layout(omk_ctx *c, int n) {
    SpLayout L;
    uint8_t *base = reinterpret_cast<uint8_t *>(c->sp_buf);
    size_t off = 0;
    auto take = [&](size_t bytes) { uint8_t *p = base + off; off = (off + bytes + 255) & ~(size_t)255; return p; };
    L.mover = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * (size_t)n));
    L.other = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * (size_t)n));
    L.actions = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * (size_t)n));
    L.temps = reinterpret_cast<float *>(take(sizeof(float) * (size_t)n));
    L.modes = take((size_t)n);
    L.status = reinterpret_cast<int8_t *>(take((size_t)n));
    L.status2 = reinterpret_cast<int8_t *>(take((size_t)n));
    L.root_policy = reinterpret_cast<float *>(take(sizeof(float) * kRow));
    L.counters = reinterpret_cast<unsigned long long *>(take(64));
    L.bytes = off;
    return L;
}
// Restarted games install a cached raw prior of the empty board instead of evaluating it again (Agent::new, agent.rs:16-25,
// evaluates with the CURRENT weights).  The cache therefore follows every weight change: omk_net_load_params and
// omk_net_init_random call this while a self-play driver is active (trainer step -> sync weights -> next self-play phase).
static int32_t sp_refresh_root_policy(omk_ctx *c) {
    if (!c->sp_active || c->sp_cfg.evaluator != OMK_EVAL_NET) return OMK_OK;
    int32_t rc = ensure_workspace(c, 1);
    if (rc) return rc;
    const SpLayout L = sp_layout(c, c->sp_cfg.n_games);
    const uint32_t one = 1;
    CK(cudaMemsetAsync(c->ws.nn_in, 0, sizeof(NNIn), c->stream));  // empty board, Black to move, EnvTurnMode::Player (== k_new_games' request)
    CK(copy_h2d(c, c->ws.n_req, &one, sizeof one, c->stream));
    if (!net_forward(c, nullptr, 1)) return fail(OMK_ERR_CUDA, "re-evaluating the empty board after a weight change failed");
    CK(cudaMemcpyAsync(L.root_policy, c->ws.P, sizeof(float) * kCells, cudaMemcpyDeviceToDevice, c->stream));
    launch_reset_requests(c);
    return check_device_error(c);
}

extern "C" int32_t omk_selfplay_begin(omk_ctx *c, const omk_selfplay_config *cfg) {
    CK(cudaSetDevice(c->device));
    if (!cfg) return fail(OMK_ERR_INVALID, "cfg is NULL");
    if (cfg->n_games < 1 || 2 * cfg->n_games > c->cap_trees) return fail(OMK_ERR_INVALID, "need capacity_trees >= 2*n_games");
    if (cfg->batch_size < 1 || cfg->batch_size > kMaxBatchPerTree || cfg->count < 1) return fail(OMK_ERR_INVALID, "bad count/batch_size");
    int32_t rc = check_evaluator(c, cfg->evaluator);
    if (rc) return rc;
    const int n = cfg->n_games;
    const int rows = n * cfg->batch_size > 2 * n ? n * cfg->batch_size : 2 * n;
    rc = ensure_workspace(c, rows);
    if (rc) return rc;
    c->sp_cfg = *cfg;
    cudaFree(c->sp_ply);
    cudaFree(c->sp_buf);
    c->sp_ply = nullptr;
    c->sp_buf = nullptr;
    CK(cudaMalloc(&c->sp_ply, sizeof(int32_t) * (size_t)n));
    CK(cudaMemsetAsync(c->sp_ply, 0, sizeof(int32_t) * (size_t)n, c->stream));
    const SpLayout L = sp_layout(c, n);
    CK(cudaMalloc(&c->sp_buf, L.bytes));
    CK(cudaMemsetAsync(c->sp_buf, 0, L.bytes, c->stream));
    const SpLayout M = sp_layout(c, n);
    // transition ring: device slots + pinned host mirror + copy stream + events
    const SpSlot slot = sp_slot_layout(n);
    if (slot.bytes != c->sp_ring_slot_bytes) {
        cudaFree(c->sp_ring_dev);
        if (c->sp_ring_host) cudaFreeHost(c->sp_ring_host);
        c->sp_ring_dev = c->sp_ring_host = nullptr;
        c->sp_ring_slot_bytes = 0;
        CK(cudaMalloc(&c->sp_ring_dev, slot.bytes * omk_ctx::kSpRing));
        CK(cudaHostAlloc(&c->sp_ring_host, slot.bytes * omk_ctx::kSpRing, cudaHostAllocDefault));
        c->sp_ring_slot_bytes = slot.bytes;
    }
    if (!c->sp_copy_stream) {
        CK(cudaStreamCreateWithFlags(&c->sp_copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < omk_ctx::kSpRing; ++i) {
            CK(cudaEventCreateWithFlags(&c->sp_ev_rec[i][0], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->sp_ev_rec[i][1], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->sp_ev_copy[i], cudaEventDisableTiming));
        }
    }
    // Agent::new for all 2n trees (stream id = tree id); then cache the root's raw policy for restarts:
    // evaluating the empty board is deterministic, so reusing it is bit-identical to evaluating it again
    launch_reset_requests(c);
    launch_new_games(c, nullptr, 2 * n, nullptr, nullptr, nullptr);
    RUN_EVAL(c, cfg->evaluator, 2 * n);
    launch_apply(c, nullptr, 2 * n, kApplyNewGame);
    CK(cudaMemcpyAsync(M.root_policy, c->tree_nodes + kOffPolicy, sizeof(float) * kCells, cudaMemcpyDeviceToDevice, c->stream));
    c->sp_active = true;
    return check_device_error(c);
}

extern "C" int32_t omk_selfplay_run(omk_ctx *c, int32_t plies, int32_t profile, uint8_t *out_boards, float *out_policy,
                                    int8_t *out_status, int32_t *out_actions, omk_selfplay_stats *stats) {
    CK(cudaSetDevice(c->device));
    if (!c->sp_active) return fail(OMK_ERR_STATE, "omk_selfplay_begin has not been called");
    if (plies < 0) return fail(OMK_ERR_INVALID, "plies < 0");
    if (c->sp_cfg.evaluator == OMK_EVAL_NET) {  // weights updated by the trainer since the last evaluation: operand images + the cached root prior
        int32_t rc_p = ensure_packed(c);
        if (rc_p) return rc_p;
    }
    const omk_selfplay_config &cfg = c->sp_cfg;
    const int n = cfg.n_games;
    const int rounds = (cfg.count + cfg.batch_size - 1) / cfg.batch_size;

    const SpLayout L = sp_layout(c, n);
    int32_t *mover = L.mover, *other = L.other, *actions = L.actions;
    float *temps = L.temps;
    uint8_t *modes = L.modes;
    int8_t *status = L.status, *status2 = L.status2;
    float *root_policy = L.root_policy;
    unsigned long long *counters = L.counters;

    // Transitions stream through a ring of kSpRing ply slots: the ply's kernels record into device slot ply % kSpRing, the
    // slice is copied to the pinned host mirror on the copy stream while later plies search, and the host drains a slot
    // into the caller's [ply][game] arrays just before the device reuses it (the host thread is then at most kSpRing
    // plies ahead of the GPU).
    constexpr int kRing = omk_ctx::kSpRing;
    const size_t tot = (size_t)plies * (size_t)n;
    const SpSlot slot = sp_slot_layout(n);
    const bool want_out = out_boards || out_policy || out_status || out_actions;
    size_t d2h = 0;
    auto dev_slot = [&](int ply) { return c->sp_ring_dev + (size_t)(ply % kRing) * slot.bytes; };
    auto drain = [&](int ply) -> int32_t {  // pinned slot of `ply` -> the caller's arrays
        CK(cudaEventSynchronize(c->sp_ev_copy[ply % kRing]));
        const uint8_t *h = c->sp_ring_host + (size_t)(ply % kRing) * slot.bytes;
        const size_t row = (size_t)ply * n;
        if (out_boards) memcpy(out_boards + row * kCells, h + slot.off_boards, (size_t)n * kCells);
        if (out_policy) memcpy(out_policy + row * kCells, h + slot.off_policy, (size_t)n * kCells * sizeof(float));
        if (out_actions) memcpy(out_actions + row, h + slot.off_actions, (size_t)n * sizeof(int32_t));
        if (out_status) memcpy(out_status + row, h + slot.off_status, (size_t)n);
        return OMK_OK;
    };
    unsigned long long h_before[2] = {0, 0}, h_after[2] = {0, 0}, h_fin0 = 0, h_fin1 = 0;
    CK(copy_d2h(c, h_before, c->dev_sims, sizeof h_before, c->stream));
    CK(copy_d2h(c, &h_fin0, counters, sizeof h_fin0, c->stream));
    CK(cudaStreamSynchronize(c->stream));

    c->prof_level = profile;
    auto span_begin = [&](int kind) { return prof_begin(c, kind, 2); };
    auto span_end = [&](bool sp) { prof_end(c, sp); };

    // Two search lanes (halves of the games) when the pool is large enough: two independent kernel chains on two streams.
    const int lanes = lanes_for(c, n);
    const int lane_n[2] = {lanes == 2 ? (n + 1) / 2 : n, lanes == 2 ? n / 2 : 0};
    const int lane_g0[2] = {0, lane_n[0]};
    float *policy_out = c->ws.policy_out;
    for (int l = 0; l < lanes; ++l) {
        LaneScope scope(c, l, 0);
        int32_t rc_ws = ensure_workspace(c, lane_n[l] * cfg.batch_size);
        if (rc_ws) return rc_ws;
    }
    CK(cudaEventRecord(c->ev0, c->stream));
    if (lanes == 2) lanes_fork(c);
    for (int ply = 0; ply < plies; ++ply) {
        if (want_out && ply >= kRing) {  // the slot this ply records into still holds ply - kRing: hand it to the caller first
            int32_t rc_d = drain(ply - kRing);
            if (rc_d) return rc_d;
        }
        uint8_t *ds = dev_slot(ply);
        uint8_t *d_boards = ds + slot.off_boards;
        float *d_policy = reinterpret_cast<float *>(ds + slot.off_policy);
        int32_t *d_actions = reinterpret_cast<int32_t *>(ds + slot.off_actions);
        int8_t *d_status = reinterpret_cast<int8_t *>(ds + slot.off_status);
        for (int l = 0; l < lanes; ++l) {
            LaneScope scope(c, l, 0);
            const int g0 = lane_g0[l], nl = lane_n[l];
            bool sp = span_begin(OMK_K_MOVE);
            launch_sp_prepare(c, g0, nl, mover, other, modes, temps);
            launch_root_noise(c, mover + g0, nl, cfg.epsilon, cfg.alpha);
            span_end(sp);
        }
        for (int r = 0; r < rounds; ++r) {
            for (int l = 0; l < lanes; ++l) {
                LaneScope scope(c, l, 0);
                const int g0 = lane_g0[l], nl = lane_n[l];
                bool sp = span_begin(OMK_K_SELECT);
                launch_reset_requests(c);
                launch_select_expand(c, mover + g0, nl, cfg.batch_size);
                span_end(sp);
                RUN_EVAL(c, cfg.evaluator, nl * cfg.batch_size);
                sp = span_begin(OMK_K_APPLY);
                launch_apply(c, mover + g0, nl, kApplySearch);
                span_end(sp);
            }
        }
        for (int l = 0; l < lanes; ++l) {
            LaneScope scope(c, l, 0);
            const int g0 = lane_g0[l], nl = lane_n[l];
            const size_t row = (size_t)g0;  // within the ply's ring slot
            bool sp = span_begin(OMK_K_MOVE);
            launch_sample(c, mover + g0, nl, modes + g0, temps + g0, actions + g0, policy_out + (size_t)g0 * kCells, nullptr);
            launch_sp_record(c, nl, mover + g0, actions + g0, policy_out + (size_t)g0 * kCells, d_boards + row * kCells,
                             d_policy + row * kCells, d_actions + row);
            launch_play(c, mover + g0, actions + g0, nl, status + g0);
            launch_reset_requests(c);
            launch_ensure_prepare(c, other + g0, actions + g0, nl);
            span_end(sp);
            RUN_EVAL(c, cfg.evaluator, nl);
            sp = span_begin(OMK_K_MOVE);
            launch_apply(c, other + g0, nl, kApplyEnsure);
            launch_play(c, other + g0, actions + g0, nl, status2 + g0);
            CK(cudaMemcpyAsync(d_status + row, status + g0, (size_t)nl, cudaMemcpyDeviceToDevice, c->stream));
            // finished games restart with fresh trees for both colours (cached root policy) and ply 0
            launch_new_games(c, mover + g0, nl, nullptr, status + g0, root_policy);
            launch_new_games(c, other + g0, nl, nullptr, status + g0, root_policy);
            launch_sp_advance(c, g0, nl, status, counters);
            span_end(sp);
            if (want_out) CK(cudaEventRecord(c->sp_ev_rec[ply % kRing][l], c->stream));
        }
        if (want_out) {  // this ply's slice -> pinned host mirror, behind both lanes' recording kernels
            for (int l = 0; l < lanes; ++l) CK(cudaStreamWaitEvent(c->sp_copy_stream, c->sp_ev_rec[ply % kRing][l], 0));
            uint8_t *hs = c->sp_ring_host + (size_t)(ply % kRing) * slot.bytes;
            if (out_boards) { CK(copy_d2h(c, hs + slot.off_boards, d_boards, (size_t)n * kCells, c->sp_copy_stream)); d2h += (size_t)n * kCells; }
            if (out_policy) { CK(copy_d2h(c, hs + slot.off_policy, d_policy, (size_t)n * kCells * sizeof(float), c->sp_copy_stream)); d2h += (size_t)n * kCells * 4; }
            if (out_actions) { CK(copy_d2h(c, hs + slot.off_actions, d_actions, (size_t)n * 4, c->sp_copy_stream)); d2h += (size_t)n * 4; }
            if (out_status) { CK(copy_d2h(c, hs + slot.off_status, d_status, (size_t)n, c->sp_copy_stream)); d2h += (size_t)n; }
            CK(cudaEventRecord(c->sp_ev_copy[ply % kRing], c->sp_copy_stream));
            // the device slot is recorded into again kRing plies from now: by then the host has waited for this copy (drain)
        }
    }
    if (lanes == 2) {
        LaneScope scope(c, 1, 0);
        launch_reset_requests(c);
    }
    if (lanes == 2) lanes_join(c);
    launch_reset_requests(c);  // folds the last round's request count into the evaluator total
    CK(cudaEventRecord(c->ev1, c->stream));
    if (want_out)
        for (int ply = plies > kRing ? plies - kRing : 0; ply < plies; ++ply) {  // the slots still in flight
            int32_t rc_d = drain(ply);
            if (rc_d) return rc_d;
        }
    CK(copy_d2h(c, h_after, c->dev_sims, sizeof h_after, c->stream));
    CK(copy_d2h(c, &h_fin1, counters, sizeof h_fin1, c->stream));
    int32_t rc = check_device_error(c);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->simulations = (int64_t)(h_after[0] - h_before[0]);
        stats->nn_evals = (int64_t)(h_after[1] - h_before[1]);
        stats->positions = (int64_t)tot;
        stats->games_finished = (int64_t)(h_fin1 - h_fin0);
        stats->d2h_bytes = (int64_t)d2h;
        stats->h2d_bytes = 0;
        cudaEventElapsedTime(&stats->gpu_ms, c->ev0, c->ev1);
        for (auto &s : c->prof_spans) {
            float ms = 0;
            cudaEventElapsedTime(&ms, s.a, s.b);
            stats->kind_ms[s.kind] += ms;
            stats->kind_launches[s.kind] += 1;
        }
    }
    for (auto &s : c->prof_spans) {
        c->prof_pool.push_back(s.a);
        c->prof_pool.push_back(s.b);
    }
    c->prof_spans.clear();
    c->prof_level = 0;
    return rc;
}

/*
 * ORACLE -- test infrastructure, NOT product code.  See omok_oracle.h.
 *
 * Plain C99 restatement of the reference's self-play hot path.  The data
 * structures deliberately mirror the REFERENCE (array-of-structs nodes, per-node
 * children vectors in creation order, byte-per-cell boards), not the CUDA
 * product's layout, so the two are independent implementations of one spec.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: Rust never contracts
 * a*b+c into an fma, so neither may we).
 */
#include "omok_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define F32_EPSILON 1.1920929e-7f /* f32::EPSILON = 2^-23 */

/* ====================================================================== */
/* environment                                                            */
/* ====================================================================== */

/* environment/src/lib.rs:73-79 */
void orc_env_new(orc_env *env) {
    env->turn = ORC_TURN_BLACK;
    env->legal_move_count = ORC_CELLS;
    memset(env->board, ORC_EMPTY, sizeof env->board);
}

/* environment/src/lib.rs:168-193: walk up to 5 offsets, stop at edge / other stone */
static int count_serial_stones(const orc_env *env, int turn, int index, int dx, int dy) {
    const uint8_t stone = turn == ORC_TURN_BLACK ? ORC_BLACK : ORC_WHITE;
    const int x0 = index % ORC_SIDE, y0 = index / ORC_SIDE;
    int count = 0;
    for (int k = 1; k <= 5; ++k) {
        const int x = x0 + dx * k, y = y0 + dy * k;
        if (x < 0 || ORC_SIDE <= x || y < 0 || ORC_SIDE <= y) break;
        if (env->board[y * ORC_SIDE + x] != stone) break;
        ++count;
    }
    return count;
}

/* environment/src/lib.rs:104-166.  Returns ORC_NONE for an occupied cell (no
 * mutation, :105-107), else the GameStatus.  No terminal guard, like the reference. */
int orc_env_place_stone(orc_env *env, int index) {
    if (env->board[index] != ORC_EMPTY) return ORC_NONE;

    env->legal_move_count -= 1;
    env->board[index] = env->turn == ORC_TURN_BLACK ? ORC_BLACK : ORC_WHITE;

    const int t = env->turn;
    const int horizontal = 1 + count_serial_stones(env, t, index, -1, 0) + count_serial_stones(env, t, index, 1, 0);
    const int vertical = 1 + count_serial_stones(env, t, index, 0, -1) + count_serial_stones(env, t, index, 0, 1);
    const int lt_rb = 1 + count_serial_stones(env, t, index, -1, -1) + count_serial_stones(env, t, index, 1, 1);
    const int lb_rt = 1 + count_serial_stones(env, t, index, -1, 1) + count_serial_stones(env, t, index, 1, -1);

    env->turn = (uint8_t)(t ^ 1); /* :148 */

    /* :150-165 -- exactly five; the win test precedes the draw test */
    if (horizontal == 5 || vertical == 5 || lt_rb == 5 || lb_rt == 5)
        return t == ORC_TURN_BLACK ? ORC_BLACK_WIN : ORC_WHITE_WIN;
    if (env->legal_move_count == 0) return ORC_DRAW;
    return ORC_IN_PROGRESS;
}

/* environment/src/lib.rs:81-102: two interleaved planes, perspective side first */
void orc_env_encode_board(const orc_env *env, int turn, float *dst) {
    for (int i = 0; i < 2 * ORC_CELLS; ++i) dst[i] = 0.0f;
    const int black_offset = turn == ORC_TURN_BLACK ? 0 : 1;
    const int white_offset = turn == ORC_TURN_BLACK ? 1 : 0;
    for (int index = 0; index < ORC_CELLS; ++index) {
        const uint8_t stone = env->board[index];
        if (stone == ORC_EMPTY) continue;
        dst[index * 2 + (stone == ORC_BLACK ? black_offset : white_offset)] = 1.0f;
    }
}

/* alpha-zero/src/encoder.rs:10-46: per position a 243-float slot; floats [0,162)
 * are encode_board's (cell,2) pairs, floats [162,243) the turn plane.          */
void orc_encode_nn_input(const orc_env *envs, int n, int mode, float *out) {
    for (int i = 0; i < n; ++i) {
        const orc_env *env = &envs[i];
        float *slot = out + (size_t)i * 243;
        const int persp = mode == 0 ? env->turn : (env->turn ^ 1);
        orc_env_encode_board(env, persp, slot);
        const float value = env->turn == ORC_TURN_BLACK ? 1.0f : 0.0f;
        for (int k = 162; k < 243; ++k) slot[k] = value;
    }
}

/* ====================================================================== */
/* specified random stream + deterministic elementary functions           */
/* ====================================================================== */

#define GOLDEN 0x9E3779B97F4A7C15ull

static uint64_t mix64(uint64_t z) { /* splitmix64 finaliser */
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* SPEC: u32 = high half of mix64(mix64(seed + G*(stream+1)) + G*(counter+1)) */
uint32_t orc_rng_u32(uint64_t seed, uint32_t stream, uint32_t counter) {
    const uint64_t k = mix64(seed + GOLDEN * ((uint64_t)stream + 1));
    return (uint32_t)(mix64(k + GOLDEN * ((uint64_t)counter + 1)) >> 32);
}

/* Restates rand 0.8.5 UniformInt::<u32>::sample_single (the path taken by
 * SliceRandom::choose -> gen_index for len <= u32::MAX): widening multiply with
 * the conservative rejection zone `(range << lz) - 1`.  The crate is a
 * Cargo.lock dependency (rand 0.8.5, Cargo.lock:2655) not vendored under
 * /root/reference; restated from its published source.                       */
uint32_t orc_rng_below(uint64_t seed, uint32_t stream, uint32_t *counter, uint32_t bound) {
    const uint32_t range = bound;
    const uint32_t zone = (range << __builtin_clz(range)) - 1u;
    for (;;) {
        const uint32_t v = orc_rng_u32(seed, stream, (*counter)++);
        const uint64_t m = (uint64_t)v * (uint64_t)range;
        if ((uint32_t)m <= zone) return (uint32_t)(m >> 32);
    }
}

static double bits_to_double(uint64_t b) {
    double d;
    memcpy(&d, &b, 8);
    return d;
}
static uint64_t double_to_bits(double d) {
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
}

/* SPEC det_log: x normal positive.  x = m*2^e, m in [sqrt(.5), sqrt(2));
 * s=(m-1)/(m+1); log m = 2s*(1 + s2/3 + ... + s2^12/25) by Horner; + e*ln2.
 * Only IEEE +,-,*,/ in double, no fma: bit-reproducible on any IEEE machine. */
double orc_det_log(double x) {
    uint64_t b = double_to_bits(x);
    int e = (int)((b >> 52) & 0x7ff) - 1023;
    b = (b & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
    double m = bits_to_double(b); /* [1,2) */
    if (m > 1.4142135623730951) {
        m = m * 0.5;
        e += 1;
    }
    const double s = (m - 1.0) / (m + 1.0);
    const double s2 = s * s;
    double acc = 1.0 / 25.0;
    for (int k = 23; k >= 1; k -= 2) acc = acc * s2 + 1.0 / (double)k;
    const double logm = 2.0 * s * acc;
    return (double)e * 0.6931471805599453 + logm;
}

/* SPEC det_exp: k=floor(x*log2e+.5); r=(x-k*LN2_HI)-k*LN2_LO; degree-14 Taylor
 * by Horner; scaled by 2^(k/2) then 2^(k-k/2) (exact power-of-two factors).   */
double orc_det_exp(double x) {
    if (x < -745.0) return 0.0;
    if (x > 709.0) return bits_to_double(0x7ff0000000000000ull);
    const double kf = floor(x * 1.4426950408889634 + 0.5);
    const int k = (int)kf;
    const double r = (x - kf * 6.93147180369123816490e-01) - kf * 1.90821492927058770002e-10;
    double acc = 1.0 / 87178291200.0; /* 1/14! */
    static const double inv_fact[14] = {1.0,
                                        1.0,
                                        1.0 / 2.0,
                                        1.0 / 6.0,
                                        1.0 / 24.0,
                                        1.0 / 120.0,
                                        1.0 / 720.0,
                                        1.0 / 5040.0,
                                        1.0 / 40320.0,
                                        1.0 / 362880.0,
                                        1.0 / 3628800.0,
                                        1.0 / 39916800.0,
                                        1.0 / 479001600.0,
                                        1.0 / 6227020800.0};
    for (int i = 13; i >= 0; --i) acc = acc * r + inv_fact[i];
    const int k1 = k / 2, k2 = k - k1;
    const double f1 = bits_to_double((uint64_t)(k1 + 1023) << 52);
    const double f2 = bits_to_double((uint64_t)(k2 + 1023) << 52);
    return acc * f1 * f2;
}

/* ---- noise stream: its own domain, one sub-stream per (epoch, cell) ---- */
typedef struct {
    uint64_t seed;
    uint32_t stream, base, j;
} noise_rng;

static double noise_unif(noise_rng *g) { /* (r + 0.5) * 2^-32 in (0,1) */
    const uint32_t r = orc_rng_u32(g->seed, g->stream, g->base + (g->j++));
    return ((double)r + 0.5) * 2.3283064365386963e-10;
}

static double noise_normal(noise_rng *g) { /* Marsaglia polar, second variate dropped */
    for (;;) {
        const double u = 2.0 * noise_unif(g) - 1.0;
        const double v = 2.0 * noise_unif(g) - 1.0;
        const double s = u * u + v * v;
        if (s >= 1.0 || s == 0.0) continue;
        return u * sqrt(-2.0 * orc_det_log(s) / s);
    }
}

/* Marsaglia-Tsang, shape >= 1/3 (structure of rand_distr 0.4.3 GammaLargeShape) */
static double gamma_large(noise_rng *g, double shape) {
    const double d = shape - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    for (;;) {
        const double x = noise_normal(g);
        const double vc = 1.0 + c * x;
        if (vc <= 0.0) continue;
        const double v = vc * vc * vc;
        const double u = noise_unif(g);
        const double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2 || orc_det_log(u) < 0.5 * x2 + d * (1.0 - v + orc_det_log(v))) return d * v;
    }
}

static double gamma_sample(noise_rng *g, double shape) {
    if (shape < 1.0) { /* rand_distr GammaSmallShape: u first, then large(shape+1)*u^(1/shape) */
        const double u = noise_unif(g);
        return gamma_large(g, shape + 1.0) * orc_det_exp(orc_det_log(u) / shape);
    }
    return gamma_large(g, shape);
}

/* SPEC Dirichlet: cell i draws Gamma(alpha,1) from sub-stream
 * counter = (epoch<<16) + (i<<9) + j of the noise domain seed; gammas rounded to
 * f32; sequential f32 sum; multiply by 1/sum (rand_distr 0.4.3 Dirichlet::sample). */
void orc_dirichlet81(uint64_t seed, uint32_t stream, uint32_t noise_epoch, float alpha, float *out) {
    const uint64_t nseed = mix64(seed ^ 0xD1B54A32D192ED03ull);
    float sum = 0.0f;
    for (int i = 0; i < ORC_CELLS; ++i) {
        noise_rng g = {nseed, stream, (noise_epoch << 16) + ((uint32_t)i << 9), 0};
        out[i] = (float)gamma_sample(&g, (double)alpha);
        sum += out[i];
    }
    const float inv = 1.0f / sum;
    for (int i = 0; i < ORC_CELLS; ++i) out[i] *= inv;
}

/* BASELINE config 2 driver: n boards, each plays `plies` uniformly random legal moves
 * (stream = board index, one bounded draw per ply over the ascending list of empty
 * cells), auto-reset after a terminal status.  Traces are [ply][board].            */
void orc_random_playout(int n, int plies, uint64_t seed, uint8_t *out_actions, int8_t *out_status, orc_env *final_envs) {
    for (int i = 0; i < n; ++i) {
        orc_env env;
        orc_env_new(&env);
        uint32_t ctr = 0;
        for (int ply = 0; ply < plies; ++ply) {
            int avail[ORC_CELLS], k = 0;
            for (int c = 0; c < ORC_CELLS; ++c)
                if (env.board[c] == ORC_EMPTY) avail[k++] = c;
            const int action = avail[orc_rng_below(seed, (uint32_t)i, &ctr, (uint32_t)k)];
            const int st = orc_env_place_stone(&env, action);
            if (out_actions) out_actions[(size_t)ply * n + i] = (uint8_t)action;
            if (out_status) out_status[(size_t)ply * n + i] = (int8_t)st;
            if (st != ORC_IN_PROGRESS) orc_env_new(&env);
        }
        if (final_envs) final_envs[i] = env;
    }
}

/* ====================================================================== */
/* fake evaluator (exact): hash of the packed board -> P in (0,1], V in [-1,1) */
/* ====================================================================== */
void orc_hash_eval(const orc_env *env, int opponent_mode, float *p81, float *v) {
    uint32_t bw[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < ORC_CELLS; ++i) {
        if (env->board[i] == ORC_BLACK) bw[i >> 5] |= 1u << (i & 31);
        if (env->board[i] == ORC_WHITE) bw[3 + (i >> 5)] |= 1u << (i & 31);
    }
    uint64_t h = 0x243F6A8885A308D3ull;
    h = mix64(h ^ ((uint64_t)bw[0] | ((uint64_t)bw[1] << 32)));
    h = mix64(h ^ ((uint64_t)bw[2] | ((uint64_t)bw[3] << 32)));
    h = mix64(h ^ ((uint64_t)bw[4] | ((uint64_t)bw[5] << 32)));
    h = mix64(h ^ ((uint64_t)env->turn | ((uint64_t)(opponent_mode ? 1 : 0) << 1)));
    for (int a = 0; a < ORC_CELLS; ++a)
        p81[a] = (float)((mix64(h + GOLDEN * (uint64_t)(a + 1)) >> 40) + 1) * 5.9604644775390625e-08f; /* 2^-24 */
    if (v) *v = (float)(mix64(h + GOLDEN * 82ull) >> 40) * 1.1920928955078125e-07f - 1.0f; /* 2^-23 */
}

/* orc_eval_fn-shaped wrapper so searches can run the fake net without leaving C */
void orc_eval_hash_batch(void *user, const orc_env *envs, int n, int mode, float *out_p, float *out_v) {
    (void)user;
    for (int i = 0; i < n; ++i) orc_hash_eval(&envs[i], mode, out_p + (size_t)i * ORC_CELLS, out_v ? out_v + i : NULL);
}

/* ====================================================================== */
/* mcts crate                                                             */
/* ====================================================================== */

/* mcts/src/node.rs:10-23 with the BoardState payload (alpha-zero/src/mcts_node.rs:7-12) */
typedef struct orc_node {
    struct orc_node *parent;
    int action;
    struct orc_node **children; /* creation order */
    int n_children, cap_children;
    float p, w;
    uint64_t n;
    orc_env env;
    int status;
    float policy[ORC_CELLS];
    float z;
} orc_node;

struct orc_agent {
    orc_env env;
    orc_node *root;
    uint64_t seed;
    uint32_t stream, rng_counter, noise_epoch;
    int node_count;
    /* per-round scratch */
    orc_node **requests;
    int n_requests, cap_requests;
    int64_t sims;
};

/* mcts/src/node.rs:27-37 */
static orc_node *node_new(orc_agent *a, orc_node *parent, int action, float p) {
    orc_node *nd = (orc_node *)calloc(1, sizeof *nd);
    nd->parent = parent;
    nd->action = action;
    nd->p = p;
    nd->w = 0.0f;
    nd->n = 0;
    a->node_count++;
    return nd;
}

/* mcts/src/lib.rs:81-93 */
static void node_free_subtree(orc_agent *a, orc_node *nd) {
    for (int i = 0; i < nd->n_children; ++i) node_free_subtree(a, nd->children[i]);
    free(nd->children);
    free(nd);
    a->node_count--;
}

/* alpha-zero/src/parallel_mcts_executor.rs:277-286 (identical copy mcts_executor.rs:262-271) */
static float compute_ucb_1(uint64_t parent_n, const orc_node *node, float c) {
    const uint64_t n = node->n;
    const float q_s_a = node->w / ((float)n + F32_EPSILON);
    const float p_s_a = node->p;
    const float bias = sqrtf((float)parent_n) / (float)(1 + n);
    return q_s_a + c * p_s_a * bias;
}

/* f32::total_cmp key: sign-magnitude -> two's complement order */
static int32_t total_order_key(float f) {
    int32_t b;
    memcpy(&b, &f, 4);
    return b ^ (int32_t)((uint32_t)(b >> 31) >> 1);
}

/* mcts/src/node.rs:39-59 + selector closure parallel_mcts_executor.rs:81-90.
 * Iterator::max_by keeps the LAST maximum.                                   */
static orc_node *select_leaf(orc_node *root) {
    orc_node *node = root;
    for (;;) {
        if (node->n_children != (int)node->env.legal_move_count) return node;
        if (node->n_children == 0) return node;
        const uint64_t parent_n = node->n > 1 ? node->n : 1;
        int best = 0;
        int32_t best_key = 0;
        for (int i = 0; i < node->n_children; ++i) {
            const int32_t key = total_order_key(compute_ucb_1(parent_n, node->children[i], 1.0f));
            if (i == 0 || key >= best_key) {
                best = i;
                best_key = key;
            }
        }
        node = node->children[best];
    }
}

/* mcts/src/node.rs:83-99 */
static void propagate(orc_node *node, float w) {
    for (;;) {
        node->n += 1;
        node->w += w;
        w = -w;
        if (!node->parent) break;
        node = node->parent;
    }
}

/* mcts/src/node.rs:61-81: None if a child with that action exists */
static orc_node *expand(orc_agent *a, orc_node *node, int action, const orc_env *env, int status, const float *policy,
                        float z) {
    for (int i = 0; i < node->n_children; ++i)
        if (node->children[i]->action == action) return NULL;
    orc_node *child = node_new(a, node, action, node->policy[action]);
    child->env = *env;
    child->status = status;
    memcpy(child->policy, policy, sizeof child->policy);
    child->z = z;
    if (node->n_children == node->cap_children) {
        node->cap_children = node->cap_children ? node->cap_children * 2 : 32;
        node->children = (orc_node **)realloc(node->children, sizeof(orc_node *) * (size_t)node->cap_children);
    }
    node->children[node->n_children++] = child;
    return child;
}

/* ====================================================================== */
/* alpha-zero: Agent                                                      */
/* ====================================================================== */

/* alpha-zero/src/agent.rs:16-35: root policy = raw evaluate_p of the empty board */
orc_agent *orc_agent_new(orc_eval_fn eval, void *user, uint64_t seed, uint32_t stream) {
    orc_agent *a = (orc_agent *)calloc(1, sizeof *a);
    orc_env_new(&a->env);
    a->seed = seed;
    a->stream = stream;
    float p[ORC_CELLS], v;
    eval(user, &a->env, 1, 0, p, &v);
    a->root = node_new(a, NULL, -1, 1.0f); /* mcts/src/lib.rs:27: root p = 1 */
    a->root->env = a->env;
    a->root->status = ORC_IN_PROGRESS;
    memcpy(a->root->policy, p, sizeof p);
    a->root->z = 0.0f;
    return a;
}

void orc_agent_free(orc_agent *a) {
    if (!a) return;
    node_free_subtree(a, a->root);
    free(a->requests);
    free(a);
}

const orc_env *orc_agent_env(const orc_agent *a) { return &a->env; }
uint32_t orc_agent_rng_counter(const orc_agent *a) { return a->rng_counter; }
int orc_agent_node_count(const orc_agent *a) { return a->node_count; }

int orc_agent_root_children(const orc_agent *a, int32_t *actions, uint64_t *n, float *w, float *p) {
    const orc_node *r = a->root;
    for (int i = 0; i < r->n_children; ++i) {
        if (actions) actions[i] = r->children[i]->action;
        if (n) n[i] = r->children[i]->n;
        if (w) w[i] = r->children[i]->w;
        if (p) p[i] = r->children[i]->p;
    }
    return r->n_children;
}

void orc_agent_root_stats(const orc_agent *a, uint64_t *n, float *w, float *p, int *status, float *policy81) {
    const orc_node *r = a->root;
    if (n) *n = r->n;
    if (w) *w = r->w;
    if (p) *p = r->p;
    if (status) *status = r->status;
    if (policy81) memcpy(policy81, r->policy, sizeof r->policy);
}

/* alpha-zero/src/agent.rs:43-77 */
int orc_agent_compute_policy(const orc_agent *a, float *policy) {
    const orc_node *root = a->root;
    float sum = 0.0f;
    for (int i = 0; i < ORC_CELLS; ++i) policy[i] = 0.0f;
    if (root->n_children == 0) return 0;
    for (int i = 0; i < root->n_children; ++i) {
        const float n = (float)root->children[i]->n;
        sum += n;
        policy[root->children[i]->action] = n;
    }
    if (sum < F32_EPSILON) return 0;
    const float sum_inv = 1.0f / sum;
    for (int i = 0; i < ORC_CELLS; ++i) policy[i] *= sum_inv;
    return 1;
}

/* alpha-zero/src/agent.rs:83-137.  Best: max_by(total_cmp) -> last maximum.
 * Boltzmann: weights exp(pi/T) for pi >= EPSILON (softmax of pi/T, not pi^(1/T)),
 * normalised, then rand 0.8.5 WeightedIndex::<f32>: cumulative sums without the
 * last total, Uniform::<f32>::new(0,total) sampled as ((u32>>9) as [1,2) float
 * - 1) * scale, partition_point(cum <= chosen).  exp is the SPEC det_exp
 * (double) rounded to f32 in place of libm expf.                             */
int orc_agent_sample_action(orc_agent *a, int mode, float temperature, float *policy) {
    if (!orc_agent_compute_policy(a, policy)) return ORC_NONE;
    if (mode == 0) {
        int best = 0;
        int32_t best_key = total_order_key(policy[0]);
        for (int i = 1; i < ORC_CELLS; ++i) {
            const int32_t key = total_order_key(policy[i]);
            if (key >= best_key) {
                best = i;
                best_key = key;
            }
        }
        return best;
    }
    float sum = 0.0f;
    float heated[ORC_CELLS];
    const float temperature_inv = 1.0f / temperature;
    for (int i = 0; i < ORC_CELLS; ++i) {
        heated[i] = 0.0f;
        if (policy[i] < F32_EPSILON) continue;
        const float h = (float)orc_det_exp((double)(policy[i] * temperature_inv));
        sum += h;
        heated[i] = h;
    }
    const float sum_inv = 1.0f / sum;
    for (int i = 0; i < ORC_CELLS; ++i) heated[i] *= sum_inv;

    float cum[ORC_CELLS - 1];
    float total = heated[0];
    for (int i = 1; i < ORC_CELLS; ++i) {
        cum[i - 1] = total;
        total += heated[i];
    }
    float scale = total; /* Uniform::new(0, total): scale = high - low */
    const float max_rand = 0.99999988079071044921875f; /* 1 - 2^-23 */
    while (scale * max_rand + 0.0f >= total) {
        uint32_t sb;
        memcpy(&sb, &scale, 4);
        sb -= 1;
        memcpy(&scale, &sb, 4);
    }
    const uint32_t r = orc_rng_u32(a->seed, a->stream, a->rng_counter++);
    const uint32_t fb = (r >> 9) | 0x3f800000u;
    float value1_2;
    memcpy(&value1_2, &fb, 4);
    const float chosen = (value1_2 - 1.0f) * scale + 0.0f;
    int idx = 0;
    while (idx < ORC_CELLS - 1 && cum[idx] <= chosen) ++idx;
    return idx;
}

/* alpha-zero/src/agent.rs:144-197 */
void orc_agent_ensure_action_exists(orc_agent *a, int action, orc_eval_fn eval, void *user) {
    if (ORC_CELLS <= action || action < 0) return;
    orc_env env = a->env;
    orc_env_place_stone(&env, action);

    float policy[ORC_CELLS], v;
    eval(user, &env, 1, 1 /* EnvTurnMode::Opponent */, policy, &v);

    policy[action] = 0.0f;
    for (int i = 0; i < ORC_CELLS; ++i)
        if (a->root->env.board[i] != ORC_EMPTY) policy[i] = 0.0f;
    float sum = 0.0f;
    for (int i = 0; i < ORC_CELLS; ++i) sum += policy[i];
    if (F32_EPSILON <= sum) {
        const float sum_inv = 1.0f / sum;
        for (int i = 0; i < ORC_CELLS; ++i) policy[i] *= sum_inv;
    }
    expand(a, a->root, action, &env, ORC_IN_PROGRESS, policy, 0.0f);
}

/* alpha-zero/src/agent.rs:206-232 + mcts/src/lib.rs:47-78 */
int orc_agent_play_action(orc_agent *a, int action) {
    orc_node *root = a->root;
    if (root->status != ORC_IN_PROGRESS) return ORC_NONE;
    int idx = -1;
    for (int i = 0; i < root->n_children; ++i)
        if (root->children[i]->action == action) {
            idx = i;
            break;
        }
    if (idx < 0) return ORC_NONE;
    const int status = orc_env_place_stone(&a->env, action);
    if (status == ORC_NONE) return ORC_NONE;

    /* transition */
    for (int i = 0; i < root->n_children; ++i)
        if (i != idx) node_free_subtree(a, root->children[i]);
    orc_node *new_root = root->children[idx];
    new_root->parent = NULL;
    uint64_t new_n = 0;
    for (int i = 0; i < new_root->n_children; ++i) new_n += new_root->children[i]->n;
    new_root->n = new_n;
    free(root->children);
    free(root);
    a->node_count--;
    a->root = new_root;
    return status;
}

/* ====================================================================== */
/* alpha-zero: ParallelMCTSExecutor::execute                               */
/* ====================================================================== */

typedef struct {
    int processed_count, batch_size;
    float epsilon, alpha;
} round_args;

/* parallel_mcts_executor.rs:44-192, one agent, one round */
static void generate_requests(orc_agent *agent, const round_args *ra) {
    agent->n_requests = 0;
    if (agent->cap_requests < ra->batch_size) {
        agent->cap_requests = ra->batch_size;
        agent->requests = (orc_node **)realloc(agent->requests, sizeof(orc_node *) * (size_t)ra->batch_size);
    }

    /* :48-76 root noise, first round only; runs (and renormalises) even when epsilon == 0.
     * With epsilon == 0 the sample is multiplied by 0 and added to policy >= 0, which
     * leaves every term unchanged, so the SPEC skips drawing it in that case.   */
    if (ra->processed_count == 0) {
        float noise[ORC_CELLS];
        if (ra->epsilon != 0.0f) {
            orc_dirichlet81(agent->seed, agent->stream, agent->noise_epoch, ra->alpha, noise);
            agent->noise_epoch++;
        } else {
            for (int i = 0; i < ORC_CELLS; ++i) noise[i] = 0.0f;
        }
        float *policy = agent->root->policy;
        for (int i = 0; i < ORC_CELLS; ++i) policy[i] = (1.0f - ra->epsilon) * policy[i] + ra->epsilon * noise[i];
        float sum = 0.0f;
        for (int i = 0; i < ORC_CELLS; ++i) sum += policy[i];
        const float sum_inv = 1.0f / sum; /* no EPSILON guard here (:63-68) */
        for (int i = 0; i < ORC_CELLS; ++i) policy[i] *= sum_inv;
        for (int i = 0; i < agent->root->n_children; ++i) {
            orc_node *child = agent->root->children[i];
            child->p = policy[child->action];
        }
    }

    for (int it = 0; it < ra->batch_size; ++it) {
        agent->sims++;
        orc_node *node = select_leaf(agent->root);

        if (node->status != ORC_IN_PROGRESS) { /* :92-97 */
            propagate(node, node->z);
            continue;
        }

        /* :101-118 ascending list of empty cells without a child; uniform choice */
        int avail[ORC_CELLS], n_avail = 0;
        uint8_t has_child[ORC_CELLS] = {0};
        for (int i = 0; i < node->n_children; ++i) has_child[node->children[i]->action] = 1;
        for (int act = 0; act < ORC_CELLS; ++act)
            if (node->env.board[act] == ORC_EMPTY && !has_child[act]) avail[n_avail++] = act;
        if (n_avail == 0) continue; /* :119-125 */
        const int action = avail[orc_rng_below(agent->seed, agent->stream, &agent->rng_counter, (uint32_t)n_avail)];

        /* :127-135 */
        orc_env env = node->env;
        const int status = orc_env_place_stone(&env, action);
        const int has_reward = status != ORC_IN_PROGRESS;
        const float reward = (status == ORC_BLACK_WIN || status == ORC_WHITE_WIN) ? 1.0f : 0.0f;

        /* :137-156 dummy policy: uniform over the child's empty cells */
        float policy[ORC_CELLS];
        for (int i = 0; i < ORC_CELLS; ++i) policy[i] = env.board[i] != ORC_EMPTY ? 0.0f : 1.0f;
        float sum = 0.0f;
        for (int i = 0; i < ORC_CELLS; ++i) sum += policy[i];
        if (F32_EPSILON <= sum) {
            const float sum_inv = 1.0f / sum;
            for (int i = 0; i < ORC_CELLS; ++i) policy[i] *= sum_inv;
        }

        /* :158-175 */
        orc_node *child = expand(agent, node, action, &env, status, policy, has_reward ? reward : 0.0f);
        if (!child) continue;

        if (has_reward) /* :177-188 */
            propagate(child, reward);
        else
            agent->requests[agent->n_requests++] = child;
    }
}

typedef struct {
    orc_agent **agents;
    int n_agents, tid, n_threads;
    const round_args *ra;
} worker_args;

static void *worker_main(void *p) {
    worker_args *wa = (worker_args *)p;
    for (int i = wa->tid; i < wa->n_agents; i += wa->n_threads) generate_requests(wa->agents[i], wa->ra);
    return NULL;
}

/* parallel_mcts_executor.rs:26-270 */
int64_t orc_execute(orc_agent **agents, int n_agents, int count, int batch_size, float epsilon, float alpha,
                    orc_eval_fn eval, void *user, int n_threads) {
    int processed_count = 0;
    int64_t sims0 = 0;
    for (int i = 0; i < n_agents; ++i) sims0 += agents[i]->sims;
    const size_t cap = (size_t)n_agents * (size_t)batch_size;
    orc_node **requests = (orc_node **)malloc(sizeof(orc_node *) * (cap ? cap : 1));
    orc_env *envs = (orc_env *)malloc(sizeof(orc_env) * (cap ? cap : 1));
    float *pol = (float *)malloc(sizeof(float) * ORC_CELLS * (cap ? cap : 1));
    float *val = (float *)malloc(sizeof(float) * (cap ? cap : 1));
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_agents) n_threads = n_agents > 0 ? n_agents : 1;
    pthread_t *threads = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    worker_args *wargs = (worker_args *)malloc(sizeof(worker_args) * (size_t)n_threads);

    while (processed_count < count) { /* :39-42 */
        const round_args ra = {processed_count, batch_size, epsilon, alpha};
        /* :194-205: serial in the first round, par_iter afterwards; per-agent work is
         * independent, so thread count never changes results */
        if (n_threads == 1 || processed_count == 0) {
            for (int i = 0; i < n_agents; ++i) generate_requests(agents[i], &ra);
        } else {
            for (int t = 0; t < n_threads; ++t) {
                wargs[t] = (worker_args){agents, n_agents, t, n_threads, &ra};
                pthread_create(&threads[t], NULL, worker_main, &wargs[t]);
            }
            for (int t = 0; t < n_threads; ++t) pthread_join(threads[t], NULL);
        }
        /* flat_map().collect() keeps agent order */
        int n_req = 0;
        for (int i = 0; i < n_agents; ++i)
            for (int k = 0; k < agents[i]->n_requests; ++k) requests[n_req++] = agents[i]->requests[k];

        processed_count += batch_size; /* :207 */
        if (n_req == 0) continue;      /* :209-213 */

        for (int i = 0; i < n_req; ++i) envs[i] = requests[i]->env;
        eval(user, envs, n_req, 0 /* EnvTurnMode::Player */, pol, val); /* :215-220 */

        for (int b = 0; b < n_req; ++b) { /* :222-265 */
            orc_node *node = requests[b];
            const float value = -val[b];
            float policy[ORC_CELLS];
            memcpy(policy, pol + (size_t)b * ORC_CELLS, sizeof policy);
            for (int act = 0; act < ORC_CELLS; ++act)
                if (node->env.board[act] != ORC_EMPTY) policy[act] = 0.0f;
            float sum = 0.0f;
            for (int act = 0; act < ORC_CELLS; ++act) sum += policy[act];
            if (F32_EPSILON <= sum) {
                const float sum_inv = 1.0f / sum;
                for (int act = 0; act < ORC_CELLS; ++act) policy[act] *= sum_inv;
            }
            memcpy(node->policy, policy, sizeof policy);
            for (int i = 0; i < node->n_children; ++i) node->children[i]->p = policy[node->children[i]->action];
            propagate(node, value);
        }
    }

    free(threads);
    free(wargs);
    free(requests);
    free(envs);
    free(pol);
    free(val);
    int64_t sims1 = 0;
    for (int i = 0; i < n_agents; ++i) sims1 += agents[i]->sims;
    return sims1 - sims0;
}

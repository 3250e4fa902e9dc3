/*
 * ORACLE -- test infrastructure, NOT product code.
 *
 * CPU restatement (plain C99) of the self-play hot path of AcrylicShrimp/omok-ai:
 * board environment, the generic MCTS tree and the AlphaZero search driver.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may build, link or call anything in oracle/.
 * The product (omok-ai_b200/) never includes this header.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/).  The reference cannot be compiled here (no rustc/cargo, no
 * libtensorflow), so this restatement is pinned as follows:
 *   - environment: against every golden vector of the reference's own unit tests
 *     (environment/src/lib.rs:196-427) -- see tests/test_oracle_env.py;
 *   - mcts / alpha-zero search: the reference holds NO test for these crates
 *     ("parity unpinned"); the restatement is anchored on the code itself and on
 *     known-answer tests derived from it (tests/test_oracle_search.py);
 *   - random numbers: the reference draws from rand 0.8.5 `thread_rng()`
 *     (ChaCha12, OS-seeded: unreproducible even against itself).  It is replaced
 *     by a SPECIFIED counter-based stream (orc_rng_u32) consumed the same way by
 *     oracle and CUDA kernels; the integer-range algorithm on top restates
 *     rand 0.8.5 `UniformInt::<u32>::sample_single` (see orc_rng_below).
 */
#ifndef OMOK_ORACLE_H
#define OMOK_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_CELLS 81
#define ORC_SIDE 9

/* environment/src/lib.rs:4-9, :22-25, :46-51 (declaration order == discriminant) */
enum { ORC_EMPTY = 0, ORC_BLACK = 1, ORC_WHITE = 2 };
enum { ORC_TURN_BLACK = 0, ORC_TURN_WHITE = 1 };
enum { ORC_IN_PROGRESS = 0, ORC_DRAW = 1, ORC_BLACK_WIN = 2, ORC_WHITE_WIN = 3 };
#define ORC_NONE (-1) /* Option::None of place_stone / play_action */

/* environment/src/lib.rs:62-67 */
typedef struct {
    uint8_t turn;
    uint16_t legal_move_count;
    uint8_t board[ORC_CELLS];
} orc_env;

void orc_env_new(orc_env *env);
int orc_env_place_stone(orc_env *env, int index);
void orc_env_encode_board(const orc_env *env, int turn, float *dst162);
/* alpha-zero/src/encoder.rs:10-46; mode 0 = EnvTurnMode::Player, 1 = ::Opponent */
void orc_encode_nn_input(const orc_env *envs, int n, int mode, float *out243);

/* BASELINE config 2 (random-play boards); traces [ply][board]; final_envs[n] may be NULL */
void orc_random_playout(int n, int plies, uint64_t seed, uint8_t *out_actions, int8_t *out_status, orc_env *final_envs);

/* ---- specified random stream (replaces thread_rng) ---- */
uint32_t orc_rng_u32(uint64_t seed, uint32_t stream, uint32_t counter);
/* uniform integer in [0, bound) drawing from (seed, stream, *counter) */
uint32_t orc_rng_below(uint64_t seed, uint32_t stream, uint32_t *counter, uint32_t bound);
/* deterministic elementary functions shared by spec with the CUDA side */
double orc_det_log(double x);
double orc_det_exp(double x);
/* Dirichlet([alpha;81]) sample, f32 result (see omok_oracle.c for the spec).
 * noise_epoch = how many times noise was applied to this agent's roots so far. */
void orc_dirichlet81(uint64_t seed, uint32_t stream, uint32_t noise_epoch, float alpha, float *out81);

/* ---- fake evaluator: exact integer/f32 hash-net (bit-identical on device) ---- */
void orc_hash_eval(const orc_env *env, int opponent_mode, float *p81, float *v);

/* ---- evaluator injection ----
 * Called once per NN batch, from the thread that called orc_execute.
 * envs: n environments; mode: 0 Player / 1 Opponent encoding;
 * out_p: n*81, out_v: n (ignored by evaluate_p callers).                      */
typedef void (*orc_eval_fn)(void *user, const orc_env *envs, int n, int mode, float *out_p, float *out_v);

/* the fake net as an orc_eval_fn */
void orc_eval_hash_batch(void *user, const orc_env *envs, int n, int mode, float *out_p, float *out_v);

typedef struct orc_agent orc_agent;

/* alpha-zero/src/agent.rs:16-35 (evaluate_p of the empty board via `eval`) */
orc_agent *orc_agent_new(orc_eval_fn eval, void *user, uint64_t seed, uint32_t stream);
void orc_agent_free(orc_agent *a);
const orc_env *orc_agent_env(const orc_agent *a);
uint32_t orc_agent_rng_counter(const orc_agent *a);
int orc_agent_node_count(const orc_agent *a);

/* alpha-zero/src/parallel_mcts_executor.rs:26-270.  n_threads > 1 runs the
 * per-agent request generation on that many pthreads (rayon stand-in); results
 * do not depend on n_threads.  Returns total simulations executed.           */
int64_t orc_execute(orc_agent **agents, int n_agents, int count, int batch_size, float epsilon,
                    float alpha, orc_eval_fn eval, void *user, int n_threads);

/* root inspection (children in creation order); returns number of children */
int orc_agent_root_children(const orc_agent *a, int32_t *actions, uint64_t *n, float *w, float *p);
void orc_agent_root_stats(const orc_agent *a, uint64_t *n, float *w, float *p, int *status, float *policy81);

/* alpha-zero/src/agent.rs:43-77; returns 0 if None */
int orc_agent_compute_policy(const orc_agent *a, float *policy81);
/* alpha-zero/src/agent.rs:83-137; mode 0 Best, 1 Boltzmann(temperature).
 * returns action or ORC_NONE; policy81 = un-heated visit policy             */
int orc_agent_sample_action(orc_agent *a, int mode, float temperature, float *policy81);
/* alpha-zero/src/agent.rs:144-197 */
void orc_agent_ensure_action_exists(orc_agent *a, int action, orc_eval_fn eval, void *user);
/* alpha-zero/src/agent.rs:206-232; returns status or ORC_NONE */
int orc_agent_play_action(orc_agent *a, int action);

#ifdef __cplusplus
}
#endif
#endif

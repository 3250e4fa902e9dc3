// omok_b200.hpp -- header-only C++ mirror of the reference's public Rust API for the self-play path,
// over the C ABI of include/omok_b200.h.  Same names, argument meaning and None / Err behaviour as
//   environment::{Environment, Stone, Turn, GameStatus}             environment/src/lib.rs:4-193
//   alpha_zero::{Agent, ActionSamplingMode, AgentModel, EnvTurnMode,
//                encode_nn_input, MCTSExecutor, ParallelMCTSExecutor}   alpha-zero/src/*.rs
// Rust's Option<T> is std::optional<T>; Result<_, Status> is an omk::Error exception.
// (Rust is not available in the build image; the Rust shim crates in INTEGRATION.md bind the same ABI.)
#pragma once
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "omok_b200.h"

namespace omk {

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const char *msg) : std::runtime_error(msg), code(c) {}
};
inline void check(int32_t rc) {
    if (rc != OMK_OK) throw Error(rc, omk_last_error());
}

// One CUDA device + env pool + tree pool + network.  Stands in for the (Scope, Session) pair the reference
// threads through every call.
class Context {
  public:
    Context(int device, int capacity_envs, int capacity_trees, int capacity_nodes = 4096, uint64_t seed = 0)
        : cap_envs_(capacity_envs), cap_trees_(capacity_trees) {
        check(omk_ctx_create(device, capacity_envs, capacity_trees, capacity_nodes, seed, &ctx_));
        for (int i = capacity_envs - 1; i >= 0; --i) free_envs_.push_back(i);
        for (int i = capacity_trees - 1; i >= 0; --i) free_trees_.push_back(i);
    }
    ~Context() { omk_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    omk_ctx *raw() const { return ctx_; }
    int32_t take_env() { return take(free_envs_); }
    int32_t take_tree() { return take(free_trees_); }
    void give_env(int32_t i) { free_envs_.push_back(i); }
    void give_tree(int32_t i) { free_trees_.push_back(i); }

  private:
    static int32_t take(std::vector<int32_t> &v) {
        if (v.empty()) throw Error(OMK_ERR_CAPACITY, "pool exhausted");
        int32_t i = v.back();
        v.pop_back();
        return i;
    }
    omk_ctx *ctx_ = nullptr;
    int cap_envs_, cap_trees_;
    std::vector<int32_t> free_envs_, free_trees_;
};

}  // namespace omk

namespace environment {

enum class Stone : uint8_t { Empty = 0, Black = 1, White = 2 };          // lib.rs:4-9
enum class Turn : uint8_t { Black = 0, White = 1 };                      // lib.rs:21-25
inline Turn opponent(Turn t) { return t == Turn::Black ? Turn::White : Turn::Black; }
enum class GameStatus : int8_t { InProgress = 0, Draw = 1, BlackWin = 2, WhiteWin = 3 };  // lib.rs:45-51
inline bool is_terminal(GameStatus s) { return s != GameStatus::InProgress; }

// environment::Environment (lib.rs:62-193) on one slot of the device env pool
class Environment {
  public:
    static constexpr size_t BOARD_SIZE = 9;
    static constexpr size_t SERIAL_STONE_COUNT = 5;

    explicit Environment(omk::Context &ctx) : ctx_(&ctx), slot_(ctx.take_env()) { omk::check(omk_env_reset(ctx.raw(), &slot_, 1)); }
    ~Environment() { ctx_->give_env(slot_); }
    Environment(const Environment &o) : ctx_(o.ctx_), slot_(o.ctx_->take_env()) {  // Clone
        std::array<uint8_t, OMK_CELLS> b{};
        uint8_t t = 0;
        omk::check(omk_env_get(ctx_->raw(), &o.slot_, 1, b.data(), &t, nullptr));
        omk::check(omk_env_set(ctx_->raw(), &slot_, 1, b.data(), &t));
    }
    Turn turn() const {
        uint8_t t = 0;
        omk::check(omk_env_get(ctx_->raw(), &slot_, 1, nullptr, &t, nullptr));
        return Turn(t);
    }
    uint16_t legal_move_count() const {
        uint16_t l = 0;
        omk::check(omk_env_get(ctx_->raw(), &slot_, 1, nullptr, nullptr, &l));
        return l;
    }
    std::array<Stone, OMK_CELLS> board() const {
        std::array<uint8_t, OMK_CELLS> b{};
        omk::check(omk_env_get(ctx_->raw(), &slot_, 1, b.data(), nullptr, nullptr));
        std::array<Stone, OMK_CELLS> out{};
        for (size_t i = 0; i < OMK_CELLS; ++i) out[i] = Stone(b[i]);
        return out;
    }
    // place_stone (lib.rs:104-166): nullopt if the cell is occupied (no mutation)
    std::optional<GameStatus> place_stone(size_t index) {
        const uint8_t a = (uint8_t)index;
        int8_t st = 0;
        omk::check(omk_env_step(ctx_->raw(), &slot_, &a, 1, &st, nullptr));
        if (st == OMK_NONE) return std::nullopt;
        return GameStatus(st);
    }
    // encode_board (lib.rs:81-102): 162 floats, perspective `turn`
    void encode_board(Turn turn, float *dst162) const {
        std::array<float, 243> img{};
        const int mode = turn == this->turn() ? OMK_TURN_MODE_PLAYER : OMK_TURN_MODE_OPPONENT;
        omk::check(omk_env_encode(ctx_->raw(), &slot_, 1, mode, img.data()));
        for (int i = 0; i < 162; ++i) dst162[i] = img[i];
    }
    int32_t slot() const { return slot_; }

  private:
    omk::Context *ctx_;
    int32_t slot_;
};

}  // namespace environment

namespace alpha_zero {

using environment::GameStatus;
enum class EnvTurnMode : int32_t { Player = 0, Opponent = 1 };  // encoder.rs:4-8

struct ActionSamplingMode {  // agent.rs:236-241
    int32_t kind;
    float temperature;
    static ActionSamplingMode Best() { return {OMK_SAMPLE_BEST, 1.0f}; }
    static ActionSamplingMode Boltzmann(float t) { return {OMK_SAMPLE_BOLTZMANN, t}; }
};

// encode_nn_input (encoder.rs:10-46): n * 243 floats, read by the network as [n, 9, 9, 3]
inline std::vector<float> encode_nn_input(omk::Context &ctx, EnvTurnMode mode, const std::vector<const environment::Environment *> &envs) {
    std::vector<int32_t> ids;
    for (auto *e : envs) ids.push_back(e->slot());
    std::vector<float> out(ids.size() * 243);
    omk::check(omk_env_encode(ctx.raw(), ids.data(), (int32_t)ids.size(), (int32_t)mode, out.data()));
    return out;
}

// AgentModel forward side (agent_model.rs:105-134)
class AgentModel {
  public:
    explicit AgentModel(omk::Context &ctx, uint64_t seed = 0) : ctx_(&ctx) { omk::check(omk_net_init_random(ctx.raw(), seed)); }
    AgentModel(omk::Context &ctx, const float *const *tensors31, const int64_t *lens31) : ctx_(&ctx) {
        omk::check(omk_net_load_params(ctx.raw(), tensors31, lens31));
    }
    // input: n * 243 floats ([n,9,9,3]); returns p [n*81]
    std::vector<float> evaluate_p(const std::vector<float> &input) const {
        const int32_t n = (int32_t)(input.size() / 243);
        std::vector<float> p((size_t)n * OMK_CELLS);
        omk::check(omk_net_eval_images(ctx_->raw(), input.data(), n, p.data(), nullptr));
        return p;
    }
    std::pair<std::vector<float>, std::vector<float>> evaluate_pv(const std::vector<float> &input) const {
        const int32_t n = (int32_t)(input.size() / 243);
        std::vector<float> p((size_t)n * OMK_CELLS), v((size_t)n);
        omk::check(omk_net_eval_images(ctx_->raw(), input.data(), n, p.data(), v.data()));
        return {std::move(p), std::move(v)};
    }
    // AgentModel::train (agent_model.rs:136-168): one Adadelta step on (input [n*243], policy_target [n*81],
    // value_target [n]) and the reference's second forward: (p_loss, v_loss, loss) after the update
    struct Losses { float p_loss, v_loss, loss; };
    Losses train(const std::vector<float> &input, const std::vector<float> &policy_target, const std::vector<float> &value_target) const {
        const int32_t n = (int32_t)(input.size() / 243);
        if (policy_target.size() != (size_t)n * OMK_CELLS || value_target.size() != (size_t)n) throw omk::Error(OMK_ERR_INVALID, "train: target sizes");
        float l[3] = {0, 0, 0};
        omk::check(omk_train_step(ctx_->raw(), input.data(), policy_target.data(), value_target.data(), n, l));
        return {l[0], l[1], l[2]};
    }

  private:
    omk::Context *ctx_;
};

// alpha_zero::ModelIO::{save, load} (model_io.rs:59-120): `SavedData { variable_names: Vec<String>, parameters:
// Vec<Vec<f32>> }` in bincode 1.3.3's default encoding (little-endian, u64 length prefixes).  As in the reference,
// `load` matches parameters to variables by position and ignores the names; optimizer slots are not stored.
class ModelIO {
  public:
    static constexpr int kTensors = 31;
    static const int64_t *lens() {
        static const int64_t L[kTensors] = {384, 128, 4096, 32, 288, 1024, 32, 4096, 128, 4096, 32, 288, 1024, 32, 4096, 128,
                                           4096, 32, 288, 1024, 32, 4096, 128, 10368LL * 512, 512, 512 * 512, 512, 512, 1, 512 * 81, 81};
        return L;
    }
    static std::vector<std::string> variable_names() {  // TensorFlow op names of the reference's builders (see model_io.py)
        std::vector<std::string> n = {"conv_w", "conv_b"};
        for (int i = 0; i < 3; ++i) {
            const std::string r = "residual_" + std::to_string(i);
            for (const char *s : {"_conv0_w", "_conv0_b", "_conv1_w", "_conv1_w_1", "_conv1_b", "_conv2_w", "_conv2_b"}) n.push_back(r + s);
        }
        for (const char *s : {"fc0_w", "fc0_b", "fc1_w", "fc1_b", "v_fc0_w", "v_fc0_b", "p_fc0_w", "p_fc0_b"}) n.push_back(s);
        return n;
    }
    // read the network's variables from the context and write the checkpoint
    static void save(omk::Context &ctx, const std::string &path) {
        std::vector<std::vector<float>> params(kTensors);
        float *ptrs[kTensors];
        for (int i = 0; i < kTensors; ++i) {
            params[i].resize((size_t)lens()[i]);
            ptrs[i] = params[i].data();
        }
        omk::check(omk_net_get_params(ctx.raw(), ptrs, lens()));
        FILE *f = std::fopen(path.c_str(), "wb");
        if (!f) throw omk::Error(OMK_ERR_INVALID, "ModelIO::save: cannot open file");
        auto u64 = [&](uint64_t v) { std::fwrite(&v, 8, 1, f); };  // little-endian hosts only (x86-64 / aarch64)
        const auto names = variable_names();
        u64(names.size());
        for (const auto &n : names) {
            u64(n.size());
            std::fwrite(n.data(), 1, n.size(), f);
        }
        u64(kTensors);
        for (const auto &p : params) {
            u64(p.size());
            std::fwrite(p.data(), 4, p.size(), f);
        }
        std::fclose(f);
    }
    // read a checkpoint and install it as the context's network
    static void load(omk::Context &ctx, const std::string &path) {
        FILE *f = std::fopen(path.c_str(), "rb");
        if (!f) throw omk::Error(OMK_ERR_INVALID, "ModelIO::load: cannot open file");
        auto fail = [&](const char *m) {
            std::fclose(f);
            throw omk::Error(OMK_ERR_INVALID, m);
        };
        auto u64 = [&]() {
            uint64_t v = 0;
            if (std::fread(&v, 8, 1, f) != 1) fail("ModelIO::load: unexpected end of file");
            return v;
        };
        const uint64_t n_names = u64();
        for (uint64_t i = 0; i < n_names; ++i) {
            const uint64_t len = u64();
            if (len > 4096 || std::fseek(f, (long)len, SEEK_CUR) != 0) fail("ModelIO::load: bad variable name");
        }
        const uint64_t n_params = u64();
        if (n_params < (uint64_t)kTensors) fail("ModelIO::load: fewer parameters than network variables");
        std::vector<std::vector<float>> params(kTensors);
        const float *ptrs[kTensors];
        for (int i = 0; i < kTensors; ++i) {
            const uint64_t len = u64();
            if (len != (uint64_t)lens()[i]) fail("ModelIO::load: parameter length does not fit its variable");
            params[i].resize((size_t)len);
            if (std::fread(params[i].data(), 4, (size_t)len, f) != len) fail("ModelIO::load: unexpected end of file");
            ptrs[i] = params[i].data();
        }
        std::fclose(f);
        omk::check(omk_net_load_params(ctx.raw(), ptrs, lens()));
    }
};

// alpha_zero::Agent (agent.rs:10-232): environment + search tree on one tree slot
class Agent {
  public:
    explicit Agent(omk::Context &ctx, int32_t evaluator = OMK_EVAL_NET) : ctx_(&ctx), slot_(ctx.take_tree()), evaluator_(evaluator) {
        omk::check(omk_pool_new_games(ctx.raw(), &slot_, 1, nullptr, evaluator));
    }
    ~Agent() { ctx_->give_tree(slot_); }
    Agent(const Agent &) = delete;
    std::optional<std::array<float, OMK_CELLS>> compute_policy() const {
        std::array<float, OMK_CELLS> pol{};
        uint8_t valid = 0;
        omk::check(omk_pool_policy(ctx_->raw(), &slot_, 1, pol.data(), &valid));
        if (!valid) return std::nullopt;
        return pol;
    }
    std::optional<std::pair<size_t, std::array<float, OMK_CELLS>>> sample_action(ActionSamplingMode mode) {
        std::array<float, OMK_CELLS> pol{};
        int32_t action = OMK_NONE;
        const uint8_t m = (uint8_t)mode.kind;
        omk::check(omk_pool_sample(ctx_->raw(), &slot_, 1, &m, &mode.temperature, &action, pol.data()));
        if (action == OMK_NONE) return std::nullopt;
        return std::make_pair((size_t)action, pol);
    }
    void ensure_action_exists(size_t action) {
        const int32_t a = (int32_t)action;
        omk::check(omk_pool_ensure_action(ctx_->raw(), &slot_, &a, 1, evaluator_));
    }
    std::optional<GameStatus> play_action(size_t action) {
        const int32_t a = (int32_t)action;
        int8_t st = 0;
        omk::check(omk_pool_play(ctx_->raw(), &slot_, &a, 1, &st));
        if (st == OMK_NONE) return std::nullopt;
        return GameStatus(st);
    }
    int32_t slot() const { return slot_; }
    int32_t evaluator() const { return evaluator_; }
    omk::Context &context() const { return *ctx_; }

  private:
    omk::Context *ctx_;
    int32_t slot_;
    int32_t evaluator_;
};

// ParallelMCTSExecutor::execute (parallel_mcts_executor.rs:26-270)
class ParallelMCTSExecutor {
  public:
    static constexpr float C_PUCT = 1.0f;
    void execute(size_t count, size_t batch_size, float epsilon, float alpha, const std::vector<Agent *> &agents) const {
        if (agents.empty()) return;
        std::vector<int32_t> ids;
        for (auto *a : agents) ids.push_back(a->slot());
        omk::check(omk_pool_search(agents[0]->context().raw(), ids.data(), (int32_t)ids.size(), (int32_t)count, (int32_t)batch_size,
                                   epsilon, alpha, agents[0]->evaluator()));
    }
};

// MCTSExecutor::run (mcts_executor.rs:29-255): the reference races its rounds on one tree; the
// deterministic restatement is the same rounds in order == execute with one agent
class MCTSExecutor {
  public:
    static constexpr float C_PUCT = 1.0f;
    void run(size_t count, size_t batch_size, float epsilon, float alpha, Agent &agent) const {
        const int32_t id = agent.slot();
        omk::check(omk_pool_search(agent.context().raw(), &id, 1, (int32_t)count, (int32_t)batch_size, epsilon, alpha, agent.evaluator()));
    }
};

}  // namespace alpha_zero

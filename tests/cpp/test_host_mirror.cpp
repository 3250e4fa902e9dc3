// C++ host-mirror test: the reference's own environment unit tests (environment/src/lib.rs:196-427) and an
// arena-shaped search (benchmark/src/agent.rs:34-50) written against include/omok_b200.hpp.
// Built by tests/test_cpp_mirror.py; runs only where a GPU is present (`--run`), otherwise compile+link only.
#include <cstdio>
#include <cstring>

#include "omok_b200.hpp"

using environment::Environment;
using environment::GameStatus;
using environment::Stone;
using environment::Turn;

#define REQUIRE(cond)                                                   \
    do {                                                                \
        if (!(cond)) {                                                  \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                   \
        }                                                               \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 2 || std::strcmp(argv[1], "--run") != 0) {
        std::printf("linked ok (version %d)\n", omk_version());
        return 0;
    }
    omk::Context ctx(0, 64, 8, 2048, 7);
    {  // test_place_stone
        Environment env(ctx);
        REQUIRE(env.turn() == Turn::Black);
        for (size_t i = 0; i < 12; ++i) {
            REQUIRE(env.place_stone(i) == GameStatus::InProgress);
            REQUIRE(env.board()[i] == (i % 2 == 0 ? Stone::Black : Stone::White));
            REQUIRE(env.turn() == (i % 2 == 0 ? Turn::White : Turn::Black));
        }
        REQUIRE(!env.place_stone(3).has_value());  // occupied -> None
    }
    {  // test_game_ending_horizontal / vertical
        Environment h(ctx), v(ctx);
        const size_t hm[] = {0, 9, 1, 10, 2, 11, 3, 12}, vm[] = {0, 2, 9, 11, 18, 20, 27, 29};
        for (size_t m : hm) REQUIRE(h.place_stone(m) == GameStatus::InProgress);
        for (size_t m : vm) REQUIRE(v.place_stone(m) == GameStatus::InProgress);
        REQUIRE(h.place_stone(4) == GameStatus::BlackWin);
        REQUIRE(v.place_stone(36) == GameStatus::BlackWin);
    }
    {  // test_game_ending_lt_rb / lb_rt
        Environment a(ctx), b(ctx);
        for (size_t i = 0; i < 36; ++i) {
            a.place_stone(i);
            b.place_stone(i);
        }
        REQUIRE(a.place_stone(40) == GameStatus::BlackWin);
        REQUIRE(b.place_stone(36) == GameStatus::BlackWin);
    }
    {  // encoding_1 / encoding_2
        Environment env(ctx);
        for (size_t m : {0, 10, 2, 30}) env.place_stone(m);
        float e[162];
        env.encode_board(Turn::Black, e);
        for (int i = 0; i < 162; ++i) REQUIRE(e[i] == ((i == 0 || i == 21 || i == 4 || i == 61) ? 1.0f : 0.0f));
        env.encode_board(Turn::White, e);
        for (int i = 0; i < 162; ++i) REQUIRE(e[i] == ((i == 1 || i == 20 || i == 5 || i == 60) ? 1.0f : 0.0f));
        Environment copy(env);  // Clone
        REQUIRE(copy.legal_move_count() == 77 && copy.turn() == Turn::Black);
    }
    {  // arena-shaped move: MCTSExecutor::run(800, 8, 0, 1) + Best + play_action
        alpha_zero::AgentModel model(ctx, 0);
        alpha_zero::Agent agent(ctx);
        REQUIRE(!agent.compute_policy().has_value());
        alpha_zero::MCTSExecutor().run(800, 8, 0.0f, 1.0f, agent);
        auto s = agent.sample_action(alpha_zero::ActionSamplingMode::Best());
        REQUIRE(s.has_value());
        float sum = 0;
        for (float p : s->second) sum += p;
        REQUIRE(sum > 0.999f && sum < 1.001f);
        REQUIRE(agent.play_action(s->first) == GameStatus::InProgress);
        REQUIRE(!agent.play_action(s->first).has_value());
        alpha_zero::Agent other(ctx);
        other.ensure_action_exists(s->first);
        REQUIRE(other.play_action(s->first) == GameStatus::InProgress);
        auto in = alpha_zero::encode_nn_input(ctx, alpha_zero::EnvTurnMode::Player, {});
        REQUIRE(in.empty());
    }
    {  // ModelIO round trip: save -> random re-init -> load restores the network's outputs bit for bit
        alpha_zero::AgentModel model(ctx, 3);
        std::vector<float> img(243, 0.0f);
        img[0] = img[21] = 1.0f;
        for (int i = 162; i < 243; ++i) img[i] = 1.0f;
        const auto before = model.evaluate_pv(img);
        const char *path = argc > 2 ? argv[2] : "/tmp/omok_b200_cpp_ckpt";
        alpha_zero::ModelIO::save(ctx, path);
        alpha_zero::AgentModel other(ctx, 4);
        const auto changed = other.evaluate_pv(img);
        REQUIRE(changed.first != before.first);
        alpha_zero::ModelIO::load(ctx, path);
        const auto after = model.evaluate_pv(img);
        REQUIRE(after.first == before.first && after.second == before.second);
        REQUIRE(alpha_zero::ModelIO::variable_names().size() == 31);
    }
    {  // AgentModel::train through omk_train_step: the loss of a fixed minibatch falls, the new weights answer evaluate_pv
        alpha_zero::AgentModel model(ctx, 5);
        const int n = 8;
        std::vector<float> img((size_t)n * 243, 0.0f), pi((size_t)n * 81, 0.0f), z((size_t)n, 0.0f);
        for (int i = 0; i < n; ++i) {
            img[(size_t)i * 243 + 2 * i] = 1.0f;                       // one stone of the side to move
            for (int k = 162; k < 243; ++k) img[(size_t)i * 243 + k] = 1.0f;
            pi[(size_t)i * 81 + (i * 7 + 3) % 81] = 1.0f;
            z[i] = (i & 1) ? 1.0f : -1.0f;
        }
        const auto before = model.evaluate_pv(img);
        const auto first = model.train(img, pi, z);
        alpha_zero::AgentModel::Losses last = first;
        for (int it = 0; it < 40; ++it) last = model.train(img, pi, z);
        REQUIRE(first.loss == first.loss && last.loss < first.loss);
        REQUIRE(first.p_loss + first.v_loss > 0.999f * first.loss && first.p_loss + first.v_loss < 1.001f * first.loss);
        const auto after = model.evaluate_pv(img);
        REQUIRE(after.first != before.first);
    }
    std::printf("cpp host mirror ok\n");
    return 0;
}

// examples/q_learning.cpp — rsrl/examples/q_learning.rs written against include/rsrl_b200.hpp.
// Every env.transition / ql.handle / policy.sample below runs on the GPU through the C ABI.
//   usage: q_learning [episodes] [step_cap]      (the reference loop is uncapped: examples/q_learning.rs:40)
#include <cstdio>
#include <cstdlib>
#include <random>

#include "rsrl_b200.hpp"

using namespace rsrl;

int main(int argc, char** argv) {
    const int episodes = argc > 1 ? atoi(argv[1]) : 5;
    const long cap = argc > 2 ? atol(argv[2]) : 10000;
    try {
        MountainCar env;
        const std::size_t n_actions = env.action_space();

        std::mt19937_64 rng(0);  // StdRng::seed_from_u64(0): unused, the engine draws from its own Philox stream
        auto basis = Fourier::from_space(5, env.state_space()).with_bias();
        auto q_func = make_shared(LFA::vector(env, basis, SGD{0.001}, n_actions));
        Greedy policy(q_func);
        QLearning ql(q_func, 0.9);

        for (int e = 0; e < episodes; ++e) {
            // Episode loop:
            long j = 0;
            MountainCar env;
            std::size_t action = policy.sample(rng, env.emit().state());

            for (long i = 0; i < cap; ++i) {
                // Trajectory loop:
                j = i;
                Transition t = env.transition(action);

                ql.handle(t);
                action = policy.sample(rng, t.to.state());

                if (t.terminated()) break;
            }
            printf("Batch %d: %ld steps...\n", e + 1, j + 1);
        }
        auto dim = q_func->weights_dim();
        auto w = q_func->weights();
        double norm = 0;
        for (double x : w) norm += x * x;
        printf("weights %zux%zu |W|^2 = %.17g\n", dim.first, dim.second, norm);
    } catch (const Error& err) {
        fprintf(stderr, "%s\n", err.what());
        return 1;
    }
    return 0;
}

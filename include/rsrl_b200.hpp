// rsrl_b200.hpp — C++ host-side mirror of rsrl's trait surface over the C ABI (include/rsrl_b200.h).
//
// The reference is Rust and no Rust toolchain exists in the build image, so the compiled host side
// above the ABI is this header: same names, argument meaning and error behaviour as the reference
// types on the hot path, so that examples/q_learning.cpp reads like rsrl/examples/q_learning.rs.
//
//   rsrl_domains::{Domain, Observation, Transition}      rsrl_domains/src/lib.rs:52-62,129-142,417-446
//   MountainCar / CartPole / Acrobot                     rsrl_domains/src/{mountain_car/discrete,cart_pole,acrobot}.rs
//   fa::linear::{LFA, basis::Fourier, optim::SGD}        rsrl/src/fa/linear.rs:11-14 (lfa crate)
//   make_shared / Shared<T>                              rsrl/src/core.rs:13-44
//   control::td::{QLearning, SARSA, ExpectedSARSA}       rsrl/src/control/td/*.rs  (Handler<&Transition>::handle)
//   policies::{Greedy, EpsilonGreedy, Random}            rsrl/src/policies/*.rs    (Policy::sample / mode)
//   params::Parameterised                                rsrl/src/params/mod.rs:116-134
//
// One rsrl_engine_t embodies (LFA weights + agent + policy); the mirror objects share it the way the
// reference shares one LFA through Shared<T> = Rc<RefCell<T>> (examples/q_learning.rs:25-26).
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "rsrl_b200.h"

namespace rsrl {

// Handler::handle returns Result<Response, Error>: a failing ABI call throws (examples ignore it with .ok()).
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("rsrl_b200 error " + std::to_string(c) + ": " + m), code(c) {}
};
inline void check(int code) {
    if (code != RSRL_OK) throw Error(code, rsrl_last_error());
}

// ---- rsrl_domains ----
enum class ObservationKind { Full, Partial, Terminal };

struct Observation {  // Observation<Vec<f64>>
    ObservationKind kind;
    std::vector<double> s;
    const std::vector<double>& state() const { return s; }
    bool is_terminal() const { return kind == ObservationKind::Terminal; }
};

struct Transition {  // Transition<Vec<f64>, usize>
    Observation from;
    std::size_t action;
    double reward;
    Observation to;
    bool terminated() const { return to.is_terminal(); }
};

struct Interval { double lo, hi; };

class Domain {
  public:
    explicit Domain(int32_t kind) : kind_(kind) {
        int32_t d, a;
        double lo[RSRL_MAX_DIM], hi[RSRL_MAX_DIM], start[RSRL_MAX_DIM];
        check(rsrl_domain_info(kind, &d, &a, lo, hi, start));
        n_actions_ = (std::size_t)a;
        for (int i = 0; i < d; ++i) { space_.push_back({lo[i], hi[i]}); s_.push_back(start[i]); }
    }
    int32_t kind() const { return kind_; }
    const std::vector<Interval>& state_space() const { return space_; }  // ProductSpace<Interval>
    std::size_t action_space() const { return n_actions_; }              // Ordinal::card()
    Observation emit() const {
        uint8_t term = 0;
        check(rsrl_domain_is_terminal(kind_, 1, s_.data(), &term));
        return {term ? ObservationKind::Terminal : ObservationKind::Full, s_};
    }
    std::pair<Observation, double> step(std::size_t a) {
        const int32_t act = (int32_t)a;
        double r = 0;
        uint8_t term = 0;
        check(rsrl_domain_step(kind_, 1, s_.data(), &act, &r, &term));
        return {Observation{term ? ObservationKind::Terminal : ObservationKind::Full, s_}, r};
    }
    Transition transition(std::size_t a) {  // lib.rs:436-446
        Observation from = emit();
        auto nr = step(a);
        return Transition{std::move(from), a, nr.second, std::move(nr.first)};
    }

  protected:
    int32_t kind_;
    std::vector<double> s_;
    std::vector<Interval> space_;
    std::size_t n_actions_ = 0;
};
struct MountainCar : Domain { MountainCar() : Domain(RSRL_MOUNTAIN_CAR) {} };  // ::default()
struct CartPole : Domain { CartPole() : Domain(RSRL_CART_POLE) {} };
struct Acrobot : Domain { Acrobot() : Domain(RSRL_ACROBOT) {} };

// ---- fa::linear ----
struct Fourier {  // Fourier::from_space(order, space).with_bias()
    int order;
    static Fourier from_space(int order, const std::vector<Interval>&) { return Fourier{order}; }
    Fourier with_bias() const { return *this; }
};
struct SGD { double lr; };

// LFA::vector(basis, SGD(lr), n_actions) wrapped in Shared<>: configuration + lazily created engine
class LFA {
  public:
    static std::shared_ptr<LFA> vector(const Domain& env, Fourier basis, SGD opt, std::size_t /*n_actions*/, int32_t dtype = RSRL_F64) {
        auto q = std::shared_ptr<LFA>(new LFA());
        check(rsrl_config_default(&q->cfg_));
        q->cfg_.domain = env.kind();
        q->cfg_.basis = RSRL_FOURIER;
        q->cfg_.basis_order = basis.order;
        q->cfg_.lr = opt.lr;
        q->cfg_.dtype = dtype;
        q->cfg_.policy = RSRL_GREEDY;
        return q;
    }
    ~LFA() { if (e_) rsrl_engine_destroy(e_); }
    rsrl_config_t& config() {
        if (e_) throw Error(RSRL_EINVAL, "the shared LFA is already in use: configure agents and policies first");
        return cfg_;
    }
    rsrl_engine_t* engine() {
        if (!e_) check(rsrl_engine_create(&cfg_, &e_));
        return e_;
    }
    // Function<(S,)>::evaluate
    std::vector<double> evaluate(const std::vector<double>& s) {
        int32_t d, a; int64_t f;
        check(rsrl_config_dims(&cfg_, &d, &a, &f));
        std::vector<double> q((std::size_t)a);
        check(rsrl_engine_evaluate(engine(), 1, s.data(), q.data()));
        return q;
    }
    // Parameterised
    std::pair<std::size_t, std::size_t> weights_dim() const {
        int32_t d, a; int64_t f;
        check(rsrl_config_dims(&cfg_, &d, &a, &f));
        return {(std::size_t)f, (std::size_t)a};
    }
    std::vector<double> weights() {  // F x A row-major (Array2<f64>)
        auto dim = weights_dim();
        std::vector<double> w(dim.first * dim.second);
        check(rsrl_engine_get_weights(engine(), w.data()));
        return w;
    }
    // RNG draw index = number of transitions handled so far (the batched step index t of the fused engine):
    // Policy::sample before step t and Handler::handle of step t both use draw t, like rsrl_engine_step.
    uint64_t draw() const { return handled_; }
    void handled() { ++handled_; }

  private:
    LFA() = default;
    rsrl_config_t cfg_{};
    rsrl_engine_t* e_ = nullptr;
    uint64_t handled_ = 0;
};
template <class T> std::shared_ptr<T> make_shared(std::shared_ptr<T> v) { return v; }  // core.rs:42-44

// ---- policies ----
class Greedy {
  public:
    explicit Greedy(std::shared_ptr<LFA> q) : q_(std::move(q)) { q_->config().policy = RSRL_GREEDY; }
    // Policy::sample(rng, state): the engine draws from its counter-based stream
    template <class Rng> std::size_t sample(Rng&, const std::vector<double>& s) {
        int32_t a = 0;
        check(rsrl_engine_sample(q_->engine(), 1, s.data(), q_->draw(), &a));
        return (std::size_t)a;
    }
    std::size_t mode(const std::vector<double>& s) {
        int32_t a = 0;
        check(rsrl_engine_mode(q_->engine(), 1, s.data(), &a));
        return (std::size_t)a;
    }
  protected:
    Greedy(std::shared_ptr<LFA> q, int32_t policy, double eps) : q_(std::move(q)) { q_->config().policy = policy; q_->config().epsilon = eps; }
    std::shared_ptr<LFA> q_;
};
struct EpsilonGreedy : Greedy {
    EpsilonGreedy(std::shared_ptr<LFA> q, double epsilon) : Greedy(std::move(q), RSRL_EPSILON_GREEDY, epsilon) {}
};
// rsrl/src/policies/softmax.rs:52-69 (Gibbs = Softmax); the temperature travels in the config's epsilon field
struct Softmax : Greedy {
    Softmax(std::shared_ptr<LFA> q, double tau) : Greedy(std::move(q), RSRL_SOFTMAX, tau) {}
    static Softmax standard(std::shared_ptr<LFA> q) { return Softmax(std::move(q), 1.0); }
};
using Gibbs = Softmax;

// ---- control::td ----
struct Response { double error; };  // q_learning.rs:17-20

class TDAgent {
  public:
    Response handle(const Transition& t) {  // Handler<&Transition<S, usize>>::handle
        const int32_t a = (int32_t)t.action;
        const uint8_t term = t.terminated() ? 1 : 0;
        double td = 0;
        check(rsrl_engine_handle(q_func->engine(), 1, t.from.state().data(), &a, &t.reward, t.to.state().data(), &term,
                                 q_func->draw(), &td));
        q_func->handled();
        return Response{td};
    }
    std::shared_ptr<LFA> q_func;
    double gamma;
  protected:
    TDAgent(std::shared_ptr<LFA> q, double g, int32_t algo, double alpha) : q_func(std::move(q)), gamma(g) {
        q_func->config().algo = algo;
        q_func->config().gamma = g;
        q_func->config().alpha = alpha;
    }
};
struct QLearning : TDAgent { QLearning(std::shared_ptr<LFA> q, double gamma) : TDAgent(std::move(q), gamma, RSRL_QLEARNING, 1.0) {} };
struct SARSA : TDAgent { SARSA(std::shared_ptr<LFA> q, double gamma) : TDAgent(std::move(q), gamma, RSRL_SARSA, 1.0) {} };
struct ExpectedSARSA : TDAgent {
    ExpectedSARSA(std::shared_ptr<LFA> q, double alpha, double gamma) : TDAgent(std::move(q), gamma, RSRL_EXPECTED_SARSA, alpha) {}
};
// rsrl/src/control/td/pal.rs:18-24
struct PAL : TDAgent {
    PAL(std::shared_ptr<LFA> q, double alpha, double gamma) : TDAgent(std::move(q), gamma, RSRL_PAL, alpha) {}
};

}  // namespace rsrl

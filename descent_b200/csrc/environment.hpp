// Environment: parameters + the per-step executor (reference: src/environment.rs).
//
// `run` keeps the reference's contract (environment.rs:326-516) -- inputs are the parameters' current
// buffers, outputs replace parameter contents, `rand_seed` reaches every kernel -- but nothing is
// decided per step: on first sight of a graph the executor plans every intermediate into one arena
// (lifetimes are static), writes updated parameters in place when the reads allow it, JIT-compiles
// all clusters as one NVRTC module, and captures the whole step as a CUDA graph.  A step is then one
// seed store + one cudaGraphLaunch.
#pragma once
#include <functional>
#include <memory>
#include <random>

#include "../../include/descent_cuda.h"
#include "codegen.hpp"
#include "host_rng.hpp"

namespace descent {

// Host generator for Environment::reset_parameter.  The reference draws from rand_chacha's
// ChaCha20Rng (environment.rs:16-40), which is not vendored and is not pinned by any reference test
// (SURVEY.md §8c): the distributions are restated (Open01, Box-Muller), the bit stream is not.
// (Round 2: host_rng.hpp restates the generator itself -- ChaCha20Rng, seed_from_u64, Open01, gen_range, shuffle -- and
// reset_parameter has an overload for it; this splitmix generator stays for callers that only need reproducible values.)
class HostRng {
public:
    explicit HostRng(uint64_t seed) : state_(seed) {}
    uint64_t next_u64() {  // splitmix64
        uint64_t z = (state_ += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
    float open01() { return ((float)(next_u32() >> 9) + 0.5f) * (1.0f / 8388608.0f); }  // (0,1), 23 bits

private:
    uint64_t state_;
};

struct KernelTiming {
    std::string label;
    std::string entry;
    int cluster = -1;
    std::vector<int> covers;  // a fused launch: every cluster whose work it does
    double ms = 0;  // average per launch
    uint32_t grid[3] = {0, 0, 0}, block = 0, smem = 0;
    double algorithmic_bytes = 0;
    double flops = 0;
};

struct GraphStats {
    int kernel_launches = 0;   // our kernels per run (excludes memsets/copies/collectives)
    int total_nodes = 0;       // every node of the captured step
    int64_t arena_bytes = 0;
    double algorithmic_bytes = 0;
    double unfused_algorithmic_bytes = 0;  // the same with every fused launch counted as the kernels it replaces
    double flops = 0;
    double jit_ms = 0;
};

class Environment {
public:
    explicit Environment(int device = 0);
    ~Environment();
    Environment(const Environment&) = delete;
    Environment& operator=(const Environment&) = delete;

    Parameter static_parameter(const Shape& shape, const std::string& name);
    Parameter trainable_parameter(const Shape& shape, const std::string& name, Initializer reset_to);
    Parameter static_parameter_with_data(const Shape& shape, const std::string& name, const std::vector<float>& data);

    // ParameterWriter (environment.rs:42-61): write `count` floats from the start, zero-fill the rest
    void write_parameter(const Parameter& p, const float* data, size_t count, bool data_is_pinned = false);
    void zero_fill(const Parameter& p) { write_parameter(p, nullptr, 0); }
    // Asynchronous whole-parameter write from pinned host memory: the host -> device copy runs on a second
    // stream while earlier graphs execute, and lands in the parameter just before the next run() / read /
    // write.  `data` must stay untouched until that run has been issued.
    void prefetch_parameter(const Parameter& p, const float* pinned_data, size_t count);
    // ParameterReader
    void read_parameter(const Parameter& p, float* dst, size_t count);
    std::vector<float> read_parameter_to_vec(const Parameter& p);
    float read_parameter_scalar(const Parameter& p);
    void reset_parameter(const Parameter& p, HostRng& rng);
    void reset_parameter(const Parameter& p, ChaCha20Rng& rng);  // the reference examples' generator

    std::unique_ptr<Scope> scope() const { return std::make_unique<Scope>(parameters_, dp_); }
    std::unique_ptr<Graph> build_graph(const std::function<void(Scope&)>& f) const;

    void run(const Graph& graph, uint32_t rand_seed);
    void print_timings(const std::string& label);  // timestamp.rs:155-179
    void sync();

    // --- backend controls (no counterpart in the reference) ---
    void set_use_cuda_graph(bool on) { use_cuda_graph_ = on; }
    void set_profile_runs(bool on) { profile_runs_ = on; }  // time every kernel of every run with events
    // false (default): every GEMM is strict FP32 (SIMT).  true: plain dense MatMuls of graphs prepared afterwards run
    // on the tensor cores with TF32 operands and FP32 accumulation (BASELINE.json north_star (b)).
    void set_tf32(bool on) { use_tf32_ = on; }
    // Plan graphs as if the device had `n` SMs (0 = the real count).  Persistent kernels size their grids from it, so a
    // small value makes every CTA walk many tiles on a small problem: the regime of the large-batch benchmark, at sizes
    // the oracle finishes in seconds (tests).  A graph planned under other options is planned again at its next run.
    void set_sm_count_override(int n) { sm_count_override_ = n; }
    void init_data_parallel(int world, int rank, const void* nccl_unique_id128);
    void set_data_parallel_for_tracing(int world, int rank) {  // host-only environments: rank-specific graphs without NCCL
        DSC_CHECK(ctx_ == nullptr, "use init_data_parallel on a device environment");
        dp_.world = world;
        dp_.rank = rank;
    }
    const DataParallel& dp() const { return dp_; }
    std::vector<KernelTiming> profile(const Graph& graph, uint32_t rand_seed, int iterations);
    GraphStats stats(const Graph& graph);
    std::string kernel_source(const Graph& graph);  // generated CUDA C (also usable without a device)
    dsc_ctx* ctx() const { return ctx_; }
    uint64_t parameter_buffer(const Parameter& p) const;
    const SharedParameters& parameters() const { return parameters_; }

    struct GraphExec;

private:
    GraphExec& prepare(const Graph& graph);
    void require_device(const char* what) const;
    void launch_all(GraphExec& exec, std::vector<float>* per_launch_ms);
    void commit_prefetches(int only_param = -1);

    struct Prefetch { int param; uint64_t staging; size_t bytes; bool pending; };
    std::vector<Prefetch> prefetches_;  // one staging buffer per prefetched parameter, reused every step

    dsc_ctx* ctx_ = nullptr;
    int sm_count_ = 148;
    SharedParameters parameters_;
    DataParallel dp_;
    bool use_cuda_graph_ = true;
    bool profile_runs_ = false;
    bool use_tf32_ = false;
    int sm_count_override_ = 0;
    CodegenOptions codegen_options() const;
    std::vector<std::pair<std::string, double>> timing_totals_;  // label -> accumulated ms
    int timing_runs_ = 0;
    std::vector<std::shared_ptr<void>> live_execs_;
};

// Generated source + launch plan for a graph without touching a device (CPU tests, build check).
std::string generate_graph_source(const Graph& graph, const CodegenOptions& options, std::vector<ClusterCode>* per_cluster);

void check(int rc);  // throws std::runtime_error(dsc_last_error()) when rc != 0

}  // namespace descent

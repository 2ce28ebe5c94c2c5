// Executor: static memory plan, in-place parameter updates, one JIT module and one CUDA graph per
// descent graph.  Reference behaviour being replaced: Environment::run / run_kernel
// (src/environment.rs:241-516), BufferHeap (src/device/buffer_heap.rs), StagingWriter/Reader.
#include "environment.hpp"

#include <algorithm>
#include <cctype>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>

namespace descent {

void check(int rc) {
    if (rc != DSC_OK) fail(std::string("device layer: ") + dsc_last_error());
}

namespace {
// every occurrence of identifier prefix `from` (followed by a non-digit) -> `to`: kernel names are k<cluster>[_suffix]
std::string rename_kernels(const std::string& text, const std::string& from, const std::string& to) {
    std::string out;
    size_t pos = 0;
    for (;;) {
        const size_t hit = text.find(from, pos);
        if (hit == std::string::npos) break;
        const size_t end = hit + from.size();
        const bool boundary_before = hit == 0 || !(isalnum((unsigned char)text[hit - 1]) || text[hit - 1] == '_');
        const bool boundary_after = end >= text.size() || !isdigit((unsigned char)text[end]);
        out.append(text, pos, hit - pos);
        out += (boundary_before && boundary_after) ? to : from;
        pos = end;
    }
    out.append(text, pos, std::string::npos);
    return out;
}
}  // namespace

// One translation unit per graph.  Clusters whose generated code is identical up to the kernel name (the same GEMM
// or per-element program at every timestep of an unrolled LSTM, the per-level kernels of a hash grid) are compiled
// once: later clusters launch the first one's kernels with their own buffers.
std::string generate_graph_source(const Graph& graph, const CodegenOptions& options, std::vector<ClusterCode>* per_cluster) {
    std::string src = kernel_prelude();
    std::map<std::string, int> first_with_body;
    const int nc = (int)graph.clusters().size();

    // Operand prologues (graph.hpp OperandPrologue): a per-element producer is evaluated inside the operand loaders of
    // the GEMMs that consume it when EVERY consumer's kernel can do that under these options; then the producer has no
    // kernel and its output is never written.  Otherwise nothing changes for that producer.
    std::vector<PrologueRequest> accepted(nc);
    std::vector<ClusterCode> pregenerated(nc);
    std::vector<char> has_pregenerated(nc, 0), skipped(nc, 0), in_dense_chain(nc, 0);
    // Dense chains (graph.hpp DenseChain): one kernel at the chain's last cluster does the work of all of them when the
    // options and widths allow; the other clusters of the chain then run nothing and their intermediates never exist.
    for (const DenseChain& chain : graph.dense_chains()) {
        ClusterCode code;
        if (!generate_dense_chain_code(graph, chain, options, &code)) continue;
        const int host = chain.last_cluster();
        for (int ci : chain.all_clusters()) {
            in_dense_chain[ci] = 1;
            if (ci != host) skipped[ci] = 1;
        }
        pregenerated[host] = std::move(code);
        has_pregenerated[host] = 1;
    }
    // scatter_add groups: maximal runs of ScatterAdd clusters of one level (clusters are in level order)
    for (int ci = 0; ci < nc;) {
        int cj = ci;
        if (graph.clusters()[ci].kind == ClusterKind::ScatterAdd)
            while (cj + 1 < nc && graph.clusters()[cj + 1].kind == ClusterKind::ScatterAdd && graph.clusters()[cj + 1].level == graph.clusters()[ci].level) cj += 1;
        if (cj > ci) {
            std::vector<int> members;
            for (int k = ci; k <= cj; ++k) members.push_back(k);
            ClusterCode code;
            if (generate_scatter_group_code(graph, members, options, &code)) {
                for (int k = ci; k < cj; ++k) skipped[k] = 1;
                pregenerated[cj] = std::move(code);
                has_pregenerated[cj] = 1;
            }
        }
        ci = cj + 1;
    }
    for (const OperandPrologue& cand : graph.operand_prologues()) {
        bool feeds_dense_chain = false;
        for (const auto& use : cand.uses) feeds_dense_chain |= in_dense_chain[use.cluster] != 0;
        if (feeds_dense_chain || in_dense_chain[cand.producer]) continue;  // the fused kernel reads its operands from memory
        std::vector<std::pair<int, PrologueRequest>> requests;  // one per consumer cluster (A and B may both name this producer)
        for (const auto& use : cand.uses) {
            auto it = std::find_if(requests.begin(), requests.end(), [&](const auto& r) { return r.first == use.cluster; });
            if (it == requests.end()) {
                requests.push_back({use.cluster, accepted[use.cluster]});
                it = requests.end() - 1;
            }
            it->second.producer[use.operand] = &graph.clusters()[cand.producer];
        }
        std::vector<ClusterCode> codes;
        bool all_fused = true;
        for (auto& [ci, request] : requests) {
            codes.push_back(generate_cluster_code(graph, ci, options, &request));
            for (int k = 0; k < 2; ++k)
                if (request.producer[k] && !request.fused[k]) all_fused = false;
            if (!all_fused) break;
        }
        if (!all_fused) continue;
        for (size_t i = 0; i < requests.size(); ++i) {
            accepted[requests[i].first] = requests[i].second;
            pregenerated[requests[i].first] = std::move(codes[i]);
            has_pregenerated[requests[i].first] = 1;
        }
        skipped[cand.producer] = 1;
    }

    for (int ci = 0; ci < nc; ++ci) {
        {
            // generated kernels index with 32-bit integers: refuse anything they could not address instead of wrapping
            const Cluster& c = graph.clusters()[ci];
            constexpr int64_t kLimit = (int64_t)1 << 31;
            auto check_cluster = [&](const Cluster& k) {
                for (const auto& in : k.inputs)
                    DSC_CHECK(in.chain.input_count < kLimit && in.chain.output_count < kLimit && in.arg_shape.element_count() < kLimit,
                              "cluster '" << k.label << "' addresses " << std::max(in.chain.input_count, in.chain.output_count)
                                          << " elements through one operand; kernels index with 32 bits (limit 2^31): use a smaller per-GPU mini-batch");
                for (int out : k.outputs)
                    DSC_CHECK(graph.ops().nodes[out].shape.element_count() < kLimit,
                              "cluster '" << k.label << "' writes " << graph.ops().nodes[out].shape.element_count()
                                          << " elements; kernels index with 32 bits (limit 2^31): use a smaller per-GPU mini-batch");
            };
            check_cluster(c);
            for (const Cluster& sub : c.group) check_cluster(sub);
            for (const Cluster& sub : c.epilogue) check_cluster(sub);
            for (const Cluster& sub : c.column_sum) check_cluster(sub);
        }
        ClusterCode code;
        if (skipped[ci]) code.skipped = true;
        else if (has_pregenerated[ci]) code = std::move(pregenerated[ci]);
        else code = generate_cluster_code(graph, ci, options);
        const std::string name = "k" + std::to_string(ci);
        if (!code.source.empty()) {
            std::string body = rename_kernels(code.source, name, "k@");
            auto it = first_with_body.find(body);
            if (it == first_with_body.end()) {
                first_with_body.emplace(std::move(body), ci);
                src += code.source;
            } else {
                const std::string original = "k" + std::to_string(it->second);
                for (auto& l : code.launches)
                    if (!l.entry.empty()) l.entry = rename_kernels(l.entry, name, original);
                code.source.clear();
            }
        }
        if (per_cluster) per_cluster->push_back(std::move(code));
    }
    return src;
}

namespace {
constexpr int64_t kAlign = 256;
int64_t align_up(int64_t v, int64_t a = kAlign) { return (v + a - 1) / a * a; }

struct Storage {
    enum Kind { None, Param, Arena } kind = None;
    int param = -1;
    int64_t offset = 0;
};

// first-fit free list over one growing arena
struct ArenaAllocator {
    std::vector<std::pair<int64_t, int64_t>> free_list;  // (offset, size), sorted by offset
    int64_t top = 0;
    int64_t alloc(int64_t bytes) {
        bytes = align_up(std::max<int64_t>(bytes, 4));
        for (size_t i = 0; i < free_list.size(); ++i) {
            if (free_list[i].second >= bytes) {
                int64_t off = free_list[i].first;
                free_list[i].first += bytes;
                free_list[i].second -= bytes;
                if (free_list[i].second == 0) free_list.erase(free_list.begin() + i);
                return off;
            }
        }
        if (!free_list.empty() && free_list.back().first + free_list.back().second == top) {
            int64_t off = free_list.back().first;  // extend the trailing hole
            free_list.pop_back();
            top = off + bytes;
            return off;
        }
        int64_t off = top;
        top += bytes;
        return off;
    }
    void release(int64_t off, int64_t bytes) {
        bytes = align_up(std::max<int64_t>(bytes, 4));
        auto it = std::lower_bound(free_list.begin(), free_list.end(), std::make_pair(off, (int64_t)0));
        it = free_list.insert(it, {off, bytes});
        if (it + 1 != free_list.end() && it->first + it->second == (it + 1)->first) {
            it->second += (it + 1)->second;
            free_list.erase(it + 1);
        }
        if (it != free_list.begin() && (it - 1)->first + (it - 1)->second == it->first) {
            (it - 1)->second += it->second;
            free_list.erase(it);
        }
    }
};
}  // namespace

struct ResolvedLaunch {
    KernelLaunch::Kind kind = KernelLaunch::Kernel;
    dsc_kernel kernel = nullptr;
    uint32_t gx = 1, gy = 1, gz = 1, block = 256, smem = 0;
    std::vector<uint64_t> buffers;
    uint64_t ptr = 0;  // ZeroScratch / AllReduce / Copy target
    uint64_t src = 0;  // Copy source
    size_t bytes = 0;  // ZeroScratch / Copy bytes, AllReduce element count
    uint32_t fill_bits = 0;
    bool is_copy = false, is_fill = false;
    bool async_collective = false, join_collectives = false;  // AllReduce: on the side stream / wait for the side stream first
    std::string label, entry;
    int cluster = -1;
    int branch = -1, level = -1;  // parallel branch of its dependency level (-1: the main stream)
    std::vector<int> covers;
    double algorithmic_bytes = 0, flops = 0;
    int64_t gemm_m = 0, gemm_n = 0, gemm_k = 0;
    bool gemm_a_is_mk = true, gemm_b_is_kn = true;
    int gemm_splits = 1;
};

struct Environment::GraphExec {
    dsc_ctx* ctx = nullptr;
    dsc_module* module = nullptr;
    dsc_graph* cuda_graph = nullptr;
    uint64_t arena = 0;
    std::vector<ResolvedLaunch> launches;
    GraphStats stats;
    std::string source;
    std::vector<uint64_t> parameter_buffers_at_plan;  // the plan bakes device addresses in
    CodegenOptions options;  // what the plan was generated for: a change (set_tf32, set_sm_count_override) re-plans
    void release() {  // the environment is going away (or the graph is): give everything back while the context lives
        if (cuda_graph) dsc_graph_destroy(cuda_graph);
        if (module) dsc_module_destroy(module);
        if (arena && ctx) dsc_free(ctx, arena);
        cuda_graph = nullptr;
        module = nullptr;
        arena = 0;
        ctx = nullptr;
    }
    ~GraphExec() { release(); }
};

Environment::Environment(int device) : parameters_(std::make_shared<std::vector<ParameterStorage>>()) {
    if (device < 0) return;  // host-only: graphs and kernel source, never data
    check(dsc_ctx_create(device, &ctx_));
    check(dsc_ctx_sm_count(ctx_, &sm_count_));
}
void Environment::require_device(const char* what) const {
    DSC_CHECK(ctx_ != nullptr, what << " needs a CUDA device; this environment is host-only (there is no CPU execution path)");
}
Environment::~Environment() {
    if (ctx_) {
        dsc_sync(ctx_);
        for (auto& e : live_execs_) static_cast<GraphExec*>(e.get())->release();
        live_execs_.clear();
        for (auto& p : *parameters_)
            if (p.buffer) dsc_free(ctx_, p.buffer);
        for (auto& f : prefetches_) dsc_free(ctx_, f.staging);
        dsc_ctx_destroy(ctx_);
    }
}

Parameter Environment::static_parameter(const Shape& shape, const std::string& name) {
    ParameterStorage s;
    s.shape = shape;
    s.name = name;
    if (ctx_) check(dsc_alloc(ctx_, (size_t)shape.buffer_size(), &s.buffer));  // eager and never moved: plans bake the address in
    parameters_->push_back(s);
    return Parameter((int)parameters_->size() - 1, parameters_);
}
Parameter Environment::trainable_parameter(const Shape& shape, const std::string& name, Initializer reset_to) {
    Parameter p = static_parameter(shape, name);
    (*parameters_)[p.id()].reset_to = reset_to;
    return p;
}
Parameter Environment::static_parameter_with_data(const Shape& shape, const std::string& name, const std::vector<float>& data) {
    Parameter p = static_parameter(shape, name);
    write_parameter(p, data.data(), data.size());
    return p;
}
uint64_t Environment::parameter_buffer(const Parameter& p) const { return (*parameters_)[p.checked_id(parameters_)].buffer; }

void Environment::prefetch_parameter(const Parameter& p, const float* pinned_data, size_t count) {
    require_device("prefetch_parameter");
    const int id = p.checked_id(parameters_);
    const ParameterStorage& s = (*parameters_)[id];
    DSC_CHECK(count == (size_t)s.shape.element_count(), "prefetch_parameter writes whole parameters: '" << s.name << "' has "
                                                            << s.shape.element_count() << " elements, got " << count);
    auto it = std::find_if(prefetches_.begin(), prefetches_.end(), [&](const Prefetch& f) { return f.param == id; });
    if (it == prefetches_.end()) {
        Prefetch f{id, 0, count * 4, false};
        check(dsc_alloc(ctx_, f.bytes, &f.staging));
        check(dsc_sync(ctx_));  // the allocation is ordered on the compute stream; the copy stream is about to use it
        prefetches_.push_back(f);
        it = prefetches_.end() - 1;
    }
    check(dsc_prefetch(ctx_, it->staging, pinned_data, it->bytes));
    it->pending = true;
}
void Environment::commit_prefetches(int only_param) {
    for (Prefetch& f : prefetches_) {
        if (!f.pending || (only_param >= 0 && f.param != only_param)) continue;
        check(dsc_prefetch_commit(ctx_, (*parameters_)[f.param].buffer, f.staging, f.bytes));
        f.pending = false;
    }
}

void Environment::write_parameter(const Parameter& p, const float* data, size_t count, bool pinned) {
    if (!ctx_) return;  // host-only tracing: module/optimizer constructors may "write" initial state; nothing can read it back
    commit_prefetches(p.checked_id(parameters_));
    const ParameterStorage& s = (*parameters_)[p.checked_id(parameters_)];
    const size_t total = (size_t)s.shape.element_count();
    DSC_CHECK(count <= total, "writing " << count << " floats into parameter '" << s.name << "' of " << total);
    check(dsc_upload(ctx_, s.buffer, 0, data, count * 4, total * 4, pinned ? 1 : 0));
}
void Environment::read_parameter(const Parameter& p, float* dst, size_t count) {
    require_device("read_parameter");
    commit_prefetches(p.checked_id(parameters_));
    const ParameterStorage& s = (*parameters_)[p.checked_id(parameters_)];
    DSC_CHECK(count <= (size_t)s.shape.element_count(), "reading past the end of parameter '" << s.name << "'");
    check(dsc_download(ctx_, s.buffer, 0, dst, count * 4));
}
std::vector<float> Environment::read_parameter_to_vec(const Parameter& p) {
    std::vector<float> v((size_t)p.shape().element_count());
    read_parameter(p, v.data(), v.size());
    return v;
}
float Environment::read_parameter_scalar(const Parameter& p) {
    float v = 0.f;
    read_parameter(p, &v, 1);
    return v;
}
template <class Rng>
static std::vector<float> initial_values(const Initializer& init, size_t n, Rng& rng) {
    std::vector<float> data(n);
    const float pi = 3.14159265358979323846f;
    for (size_t i = 0; i < n; ++i) {
        if (init.kind == InitKind::RandNormal) {  // Box-Muller, environment.rs:16-27
            float u1 = rng.open01(), u2 = rng.open01();
            data[i] = init.scale * (std::sqrt(-2.0f * std::log(u1)) * std::cos(2.0f * pi * u2));
        } else {  // environment.rs:29-40
            data[i] = init.scale * (rng.open01() * 2.0f - 1.0f);
        }
    }
    return data;
}

void Environment::reset_parameter(const Parameter& p, HostRng& rng) {
    DSC_CHECK(p.is_trainable(), "reset_parameter on a static parameter");
    const Initializer init = *p.reset_to();
    const size_t n = (size_t)p.shape().element_count();
    if (init.kind == InitKind::Zero) return zero_fill(p);
    const auto data = initial_values(init, n, rng);
    write_parameter(p, data.data(), n);
}

// The reference's own generator (examples seed rand_chacha::ChaCha20Rng::seed_from_u64 and hand it to reset_parameter,
// environment.rs:190-202): same draws in the same order (host_rng.hpp).
void Environment::reset_parameter(const Parameter& p, ChaCha20Rng& rng) {
    DSC_CHECK(p.is_trainable(), "reset_parameter on a static parameter");
    const Initializer init = *p.reset_to();
    const size_t n = (size_t)p.shape().element_count();
    if (init.kind == InitKind::Zero) return zero_fill(p);
    const auto data = initial_values(init, n, rng);
    write_parameter(p, data.data(), n);
}

std::unique_ptr<Graph> Environment::build_graph(const std::function<void(Scope&)>& f) const {
    Scope s(parameters_, dp_);
    f(s);
    return std::unique_ptr<Graph>(s.build_graph());
}

void Environment::init_data_parallel(int world, int rank, const void* id) {
    if (ctx_) check(dsc_dp_init(ctx_, id, world, rank));
    dp_.world = world;
    dp_.rank = rank;
}
void Environment::sync() { if (ctx_) check(dsc_sync(ctx_)); }

// ---- planning -------------------------------------------------------------------------------------

Environment::GraphExec& Environment::prepare(const Graph& graph) {
    require_device("running a graph");
    if (graph.executor_state) {
        auto* exec = static_cast<GraphExec*>(graph.executor_state.get());
        const CodegenOptions now = codegen_options();
        if (exec->ctx == ctx_ && exec->options.use_tf32 == now.use_tf32 && exec->options.sm_count == now.sm_count) return *exec;
        if (exec->ctx == ctx_) {  // options changed since this graph was planned: drop the old plan
            check(dsc_sync(ctx_));
            live_execs_.erase(std::remove_if(live_execs_.begin(), live_execs_.end(), [&](const std::shared_ptr<void>& e) { return e.get() == exec; }),
                              live_execs_.end());
            graph.executor_state.reset();
        }
    }
    DSC_CHECK(graph.parameters() == parameters_, "graph was built for another environment");
    DSC_CHECK(graph.dp().world == dp_.world && graph.dp().rank == dp_.rank, "graph was built before data parallel was initialised");
    auto exec_ptr = std::make_shared<GraphExec>();
    GraphExec& exec = *exec_ptr;
    exec.ctx = ctx_;
    const OpGraph& ops = graph.ops();
    const auto& clusters = graph.clusters();
    const int n = (int)ops.nodes.size();
    const int nc = (int)clusters.size();
    auto cons = ops.consumers();

    const CodegenOptions opt = codegen_options();
    exec.options = opt;
    std::vector<ClusterCode> codes;
    exec.source = generate_graph_source(graph, opt, &codes);

    std::vector<Storage> storage(n);
    std::vector<int> alias(n, -1);  // AllReduce output -> its input node
    std::map<int, int> input_node_of_param;
    for (int id : graph.input_nodes()) {
        storage[id] = {Storage::Param, ops.nodes[id].op.parameter_id, 0};
        input_node_of_param[ops.nodes[id].op.parameter_id] = id;
    }
    auto cluster_of = [&](int node) { return ops.nodes[node].cluster_id; };

    struct EndOp { bool is_fill; int src_node; uint32_t bits; int param; };
    std::vector<EndOp> begin_ops, end_ops;
    for (int out_id : graph.output_nodes()) {
        const OpNode& out = ops.nodes[out_id];
        const int p = out.op.parameter_id;
        DSC_CHECK(out.in.size() == 1 && out.in[0].chain.is_identity(), "Output must read a plain array");
        const int x = out.in[0].src;
        const OpNode& xn = ops.nodes[x];
        if (xn.op.kind == OpKind::Input) {
            if (xn.op.parameter_id != p) {
                DSC_CHECK(!input_node_of_param.count(p), "copying one parameter into another that the same graph also reads is not supported");
                begin_ops.push_back({false, x, 0, p});
            }
            continue;
        }
        if (xn.op.kind == OpKind::Literal) { end_ops.push_back({true, -1, xn.op.literal_bits, p}); continue; }
        DSC_CHECK(xn.cluster_id >= 0, "Output of a value no kernel computes (" << xn.op.name() << ")");
        bool direct = storage[x].kind == Storage::None;
        auto in_it = input_node_of_param.find(p);
        if (direct && in_it != input_node_of_param.end()) {
            // the parameter is also read: write in place only if every read is earlier, or is the same
            // element in the same per-element kernel (environment.rs:372-383 swaps buffers instead)
            const int w = cluster_of(x);
            for (auto [dst, k] : cons[in_it->second]) {
                const OpNode& d = ops.nodes[dst];
                if (d.op.kind == OpKind::Output) continue;  // handled as a begin copy
                const int c = d.cluster_id;
                const OpEdge& e = d.in[k];
                bool same_element = c == w && e.chain.is_identity() &&
                                    ((clusters[w].kind == ClusterKind::PerElement && !d.op.is_gather_arg(e.arg)) ||
                                     (clusters[w].kind == ClusterKind::ScatterAdd && e.arg == 0));
                if (!(c < w || same_element)) direct = false;
            }
        }
        if (direct) storage[x] = {Storage::Param, p, 0};
        else end_ops.push_back({false, x, 0, p});
    }

    // gradient bucket: every AllReduce runs in place on its input, and the inputs are laid out
    // back to back so one collective covers them all
    ArenaAllocator arena;
    std::vector<int> bucket_nodes;
    int64_t bucket_bytes = 0;
    // clusters are in level order, and the graph pins AllReduce nodes to at most two levels (graph.cpp build_clusters):
    // each level is one contiguous range of the bucket = one collective
    struct BucketRange { int level; int64_t begin, end; };
    std::vector<BucketRange> bucket_ranges;
    for (int ci = 0; ci < nc; ++ci) {
        if (clusters[ci].kind != ClusterKind::AllReduce) continue;
        const int a = clusters[ci].outputs[0], x = clusters[ci].inputs[0].node_id;
        DSC_CHECK(storage[x].kind == Storage::None && cons[x].size() == 1, "all-reduce input must be a private intermediate");
        DSC_CHECK(storage[a].kind == Storage::None, "all-reduce output cannot be written straight into a parameter");
        storage[x] = {Storage::Arena, -1, bucket_bytes};
        alias[a] = x;
        bucket_nodes.push_back(x);
        if (bucket_ranges.empty() || bucket_ranges.back().level != clusters[ci].level) bucket_ranges.push_back({clusters[ci].level, bucket_bytes, bucket_bytes});
        bucket_bytes += align_up(ops.nodes[x].shape.buffer_size(), 16);
        bucket_ranges.back().end = bucket_bytes;
    }
    if (bucket_bytes) arena.top = align_up(bucket_bytes);

    // lifetimes of the remaining cluster outputs
    std::vector<int> death(n, -1);
    std::vector<char> lives_to_end(n, 0);
    for (const auto& eo : end_ops)
        if (!eo.is_fill) lives_to_end[eo.src_node] = 1;
    for (int id = 0; id < n; ++id) {
        if (!ops.nodes[id].alive) continue;
        for (auto [dst, k] : cons[id]) {
            (void)k;
            int c = ops.nodes[dst].cluster_id;
            if (c >= 0) death[id] = std::max(death[id], c);
        }
    }
    for (int ci = 0; ci < nc; ++ci)  // producers evaluated inside this cluster's operand loaders: their inputs live until here
        for (int id : codes[ci].extra_reads) death[id] = std::max(death[id], ci);
    // Parallel levels.  Clusters of one dependency level are independent by construction (graph.cpp build_clusters), so a
    // level with several kernels runs on parallel branches of the step's CUDA graph (dsc_branch_fork / select / join):
    // the hash-grid index and interpolation kernels, the column-sum / split-sum tails of weight gradients, the gates of an
    // unrolled LSTM step are small latency-bound launches that overlap instead of running back to back.  Not for a level
    // that holds an all-reduce (its own side stream) or a kernel that writes a parameter in place (another kernel of the
    // level may read the old value: the in-place rule above only orders EARLIER readers).  Memory follows: nothing that
    // dies inside a parallel level, and no scratch area of it, is reused before the level has ended.
    static const bool parallel_levels_enabled = [] { const char* e = std::getenv("DSC_PARALLEL_LEVELS"); return !e || std::atoi(e) != 0; }();
    static const double parallel_bytes_limit = [] { const char* e = std::getenv("DSC_PARALLEL_MB"); return (e ? std::atof(e) : 16.0) * 1.0e6; }();
    std::map<int, int> level_kernels;      // level -> clusters with launches
    std::map<int, bool> level_sequential;  // level -> must stay on the main stream
    for (int ci = 0; ci < nc; ++ci) {
        const int lv = clusters[ci].level;
        if (codes[ci].skipped || codes[ci].launches.empty()) continue;
        level_kernels[lv] += 1;
        if (clusters[ci].kind == ClusterKind::AllReduce) level_sequential[lv] = true;
        // kernels that fill the machine on their own gain nothing from a neighbour and lose cache to it (relu-pe m = 65536:
        // a 100 MB weight-gradient GEMM beside a 100 MB backward GEMM measured 3 % slower than back to back)
        for (const auto& l : codes[ci].launches)
            if (l.algorithmic_bytes > parallel_bytes_limit) level_sequential[lv] = true;
        auto writes_parameter = [&](int out) { return storage[out].kind == Storage::Param; };
        for (int out : clusters[ci].outputs) if (writes_parameter(out)) level_sequential[lv] = true;
        for (int out : codes[ci].extra_writes) if (writes_parameter(out)) level_sequential[lv] = true;
    }
    auto parallel_level = [&](int lv) { return parallel_levels_enabled && level_kernels[lv] >= 2 && !level_sequential[lv]; };
    std::vector<int> branch_of(nc, -1);
    {
        static const int branches = [] { const char* e = std::getenv("DSC_BRANCHES"); return e ? std::max(1, std::min(16, std::atoi(e))) : 16; }();  // measured on sentiment (m = 256): 7.58 ms sequential, 3.26 ms with 4 branches, 2.22 ms with 8, 1.83 ms with 16
        std::map<int, int> next_branch;
        for (int ci = 0; ci < nc; ++ci)
            if (!codes[ci].skipped && !codes[ci].launches.empty() && parallel_level(clusters[ci].level)) branch_of[ci] = next_branch[clusters[ci].level]++ % branches;
    }
    std::vector<int64_t> scratch_offset(nc, 0);
    std::vector<std::vector<int>> dying_at(nc + 1);
    std::vector<std::pair<int64_t, int64_t>> held_scratch;  // scratch areas of the current parallel level
    int released_upto = -1;
    for (int ci = 0; ci < nc; ++ci) {
        // release what nobody after the previous cluster needs -- at the end of its level when that level runs in parallel
        if (ci > 0 && (clusters[ci].level != clusters[ci - 1].level || !parallel_level(clusters[ci - 1].level))) {
            for (int c = released_upto + 1; c <= ci - 1; ++c)
                for (int id : dying_at[c]) arena.release(storage[id].offset, ops.nodes[id].shape.buffer_size());
            released_upto = ci - 1;
            for (auto [off, bytes] : held_scratch) arena.release(off, bytes);
            held_scratch.clear();
        }
        auto place = [&](int out) {
            if (alias[out] >= 0) {
                storage[out] = storage[alias[out]];
                return;
            }
            if (storage[out].kind != Storage::None) return;  // parameter or bucket
            storage[out] = {Storage::Arena, -1, arena.alloc(ops.nodes[out].shape.buffer_size())};
            int d = std::max(death[out], ci);
            if (!lives_to_end[out]) dying_at[d].push_back(out);
        };
        const bool fused_host = !codes[ci].extra_writes.empty();  // a dense chain's kernel: it writes exactly extra_writes
        if (!codes[ci].skipped && !fused_host)  // (skipped: computed on the fly by its consumers, never in memory)
            for (int out : clusters[ci].outputs) place(out);
        for (int out : codes[ci].extra_writes) place(out);
        if (codes[ci].scratch_bytes > 0) {
            scratch_offset[ci] = arena.alloc(codes[ci].scratch_bytes);
            if (parallel_level(clusters[ci].level)) held_scratch.push_back({scratch_offset[ci], codes[ci].scratch_bytes});  // ... until the level ends
            else arena.release(scratch_offset[ci], codes[ci].scratch_bytes);  // free again for the next cluster...
        }
        // ...but outputs of this cluster were allocated before the scratch, so they never overlap it
    }
    exec.stats.arena_bytes = arena.top;
    check(dsc_alloc(ctx_, (size_t)std::max<int64_t>(arena.top, 256), &exec.arena));
    check(dsc_fill_u32(ctx_, exec.arena, 0, 0, (size_t)std::max<int64_t>(arena.top, 256) / 4));

    // compile
    auto t0 = std::chrono::steady_clock::now();
    const char* options[] = {"-fmad=false"};
    bool any_kernel = false;
    for (const auto& code : codes)
        for (const auto& l : code.launches) any_kernel |= l.kind == KernelLaunch::Kernel;
    if (any_kernel) check(dsc_module_jit(ctx_, exec.source.c_str(), options, 1, &exec.module));
    exec.stats.jit_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

    auto device_address = [&](int node) -> uint64_t {
        const Storage& s = storage[node];
        DSC_CHECK(s.kind != Storage::None, "node " << node << " (" << ops.nodes[node].op.name() << ") has no storage");
        return s.kind == Storage::Param ? (*parameters_)[s.param].buffer : exec.arena + (uint64_t)s.offset;
    };
    auto add_copy_or_fill = [&](const EndOp& eo) {
        ResolvedLaunch r;
        const ParameterStorage& ps = (*parameters_)[eo.param];
        r.ptr = ps.buffer;
        r.bytes = (size_t)ps.shape.buffer_size();
        if (eo.is_fill) {
            r.is_fill = true;
            r.fill_bits = eo.bits;
            r.label = "Fill " + ps.shape.str();
        } else {
            r.is_copy = true;
            r.src = device_address(eo.src_node);
            r.label = "Copy " + ps.shape.str();
        }
        exec.launches.push_back(r);
    };
    for (const auto& eo : begin_ops) add_copy_or_fill(eo);
    size_t buckets_emitted = 0;
    int last_bucket_level = -1;
    for (int ci = 0; ci < nc; ++ci) {
        for (const KernelLaunch& l : codes[ci].launches) {
            ResolvedLaunch r;
            r.kind = l.kind;
            r.label = l.label;
            r.entry = l.entry;
            r.cluster = ci;
            r.branch = branch_of[ci];
            r.level = clusters[ci].level;
            r.covers = l.covers;
            r.algorithmic_bytes = l.algorithmic_bytes;
            r.flops = l.flops;
            if (l.kind == KernelLaunch::ZeroScratch) {
                r.ptr = exec.arena + (uint64_t)(scratch_offset[ci] + l.zero_offset);
                r.bytes = (size_t)l.zero_bytes;
            } else if (l.kind == KernelLaunch::AllReduce) {
                if (clusters[ci].level == last_bucket_level) continue;  // one collective per level: emitted at its first cluster
                last_bucket_level = clusters[ci].level;
                const BucketRange& range = bucket_ranges.at(buckets_emitted++);
                DSC_CHECK(range.level == clusters[ci].level, "gradient bucket ranges out of step with the cluster order");
                r.ptr = exec.arena + (uint64_t)range.begin;
                r.bytes = (size_t)((range.end - range.begin) / 4);
                // every bucket but the last runs on the side stream under the rest of the backward pass; the last one
                // first joins the side stream, so whatever follows it sees all gradients reduced
                r.async_collective = buckets_emitted < bucket_ranges.size();
                r.join_collectives = !r.async_collective;
                r.label = std::string(r.async_collective ? "AllReduce early bucket [" : "AllReduce bucket [") + std::to_string(r.bytes) + "]";
            } else if (l.kind == KernelLaunch::TensorGemm) {
                for (const auto& a : l.args)
                    r.buffers.push_back(a.kind == KernelArg::NodeBuffer ? device_address(a.node_id)
                                                                        : exec.arena + (uint64_t)(scratch_offset[ci] + a.scratch_offset));
                r.gemm_m = l.gemm_m; r.gemm_n = l.gemm_n; r.gemm_k = l.gemm_k;
                r.gemm_a_is_mk = l.gemm_a_is_mk; r.gemm_b_is_kn = l.gemm_b_is_kn;
                r.gemm_splits = l.gemm_splits;
                exec.stats.kernel_launches += 1;
                exec.stats.algorithmic_bytes += l.algorithmic_bytes;
                exec.stats.unfused_algorithmic_bytes += l.replaced_bytes > 0 ? l.replaced_bytes : l.algorithmic_bytes;
                exec.stats.flops += l.flops;
            } else {
                check(dsc_module_get_kernel(exec.module, l.entry.c_str(), &r.kernel));
                if (l.smem > 48 * 1024) check(dsc_kernel_set_max_dynamic_smem(r.kernel, (int)l.smem));
                r.gx = l.grid_x; r.gy = l.grid_y; r.gz = l.grid_z; r.block = l.block; r.smem = l.smem;
                for (const auto& a : l.args)
                    r.buffers.push_back(a.kind == KernelArg::NodeBuffer ? device_address(a.node_id)
                                                                        : exec.arena + (uint64_t)(scratch_offset[ci] + a.scratch_offset));
                exec.stats.kernel_launches += 1;
                exec.stats.algorithmic_bytes += l.algorithmic_bytes;
                exec.stats.unfused_algorithmic_bytes += l.replaced_bytes > 0 ? l.replaced_bytes : l.algorithmic_bytes;
                exec.stats.flops += l.flops;
            }
            exec.launches.push_back(r);
        }
    }
    for (const auto& eo : end_ops) add_copy_or_fill(eo);
    exec.stats.total_nodes = (int)exec.launches.size();
    for (const auto& p : *parameters_) exec.parameter_buffers_at_plan.push_back(p.buffer);

    graph.executor_state = exec_ptr;
    live_execs_.push_back(exec_ptr);
    return exec;
}

void Environment::launch_all(GraphExec& exec, std::vector<float>* per_launch_ms) {
    std::vector<void*> events;
    if (per_launch_ms) {
        events.resize(exec.launches.size() + 1);
        for (auto& e : events) check(dsc_event_create(&e));
        check(dsc_event_record(ctx_, events[0]));
    }
    bool forked = false;
    int forked_level = -1;
    for (size_t i = 0; i < exec.launches.size(); ++i) {
        const ResolvedLaunch& r = exec.launches[i];
        if (!per_launch_ms) {  // (per-launch timing keeps everything on one stream)
            if (forked && (r.branch < 0 || r.level != forked_level)) { check(dsc_branch_join(ctx_)); forked = false; }
            if (r.branch >= 0) {
                if (!forked) { check(dsc_branch_fork(ctx_)); forked = true; forked_level = r.level; }
                check(dsc_branch_select(ctx_, r.branch));
            }
        }
        if (r.is_copy) check(dsc_copy(ctx_, r.ptr, r.src, r.bytes));
        else if (r.is_fill) check(dsc_fill_u32(ctx_, r.ptr, 0, r.fill_bits, r.bytes / 4));
        else if (r.kind == KernelLaunch::ZeroScratch) check(dsc_fill_u32(ctx_, r.ptr, 0, 0, r.bytes / 4));
        else if (r.kind == KernelLaunch::AllReduce) {
            if (r.join_collectives) check(dsc_dp_allreduce_join(ctx_));
            if (r.async_collective) check(dsc_dp_allreduce_sum_f32_async(ctx_, r.ptr, r.bytes));
            else check(dsc_dp_allreduce_sum_f32(ctx_, r.ptr, r.bytes));
        }
        else if (r.kind == KernelLaunch::TensorGemm)
            check(dsc_gemm_tf32_split_k(ctx_, r.buffers[0], r.buffers[1], r.buffers[2], r.gemm_m, r.gemm_n, r.gemm_k, r.gemm_a_is_mk, r.gemm_b_is_kn, r.gemm_splits));
        else check(dsc_launch(ctx_, r.kernel, r.gx, r.gy, r.gz, r.block, r.smem, r.buffers.data(), (int)r.buffers.size()));
        if (per_launch_ms) check(dsc_event_record(ctx_, events[i + 1]));
    }
    if (forked) check(dsc_branch_join(ctx_));
    if (per_launch_ms) {
        per_launch_ms->resize(exec.launches.size());
        for (size_t i = 0; i < exec.launches.size(); ++i) check(dsc_event_elapsed_ms(events[i], events[i + 1], &(*per_launch_ms)[i]));
        for (auto& e : events) dsc_event_destroy(e);
    }
}

void Environment::run(const Graph& graph, uint32_t rand_seed) {
    GraphExec& exec = prepare(graph);
    commit_prefetches();
    if (profile_runs_) {
        check(dsc_set_rand_seed(ctx_, rand_seed));
        std::vector<float> ms;
        launch_all(exec, &ms);
        for (size_t i = 0; i < ms.size(); ++i) {
            auto it = std::find_if(timing_totals_.begin(), timing_totals_.end(), [&](auto& p) { return p.first == exec.launches[i].label; });
            if (it == timing_totals_.end()) timing_totals_.push_back({exec.launches[i].label, ms[i]});
            else it->second += ms[i];
        }
        timing_runs_ += 1;
        return;
    }
    if (!use_cuda_graph_) {
        check(dsc_set_rand_seed(ctx_, rand_seed));
        launch_all(exec, nullptr);
        return;
    }
    if (!exec.cuda_graph) {
        check(dsc_graph_begin_capture(ctx_));
        try {
            launch_all(exec, nullptr);
        } catch (...) {
            dsc_graph* dead = nullptr;
            dsc_graph_end_capture(ctx_, &dead);
            if (dead) dsc_graph_destroy(dead);
            throw;
        }
        check(dsc_graph_end_capture(ctx_, &exec.cuda_graph));
    }
    check(dsc_graph_launch(ctx_, exec.cuda_graph, rand_seed));
}

std::vector<KernelTiming> Environment::profile(const Graph& graph, uint32_t rand_seed, int iterations) {
    GraphExec& exec = prepare(graph);
    std::vector<KernelTiming> out(exec.launches.size());
    for (size_t i = 0; i < exec.launches.size(); ++i) {
        out[i].label = exec.launches[i].label;
        out[i].entry = exec.launches[i].entry;
        out[i].cluster = exec.launches[i].cluster;
        out[i].covers = exec.launches[i].covers;
        out[i].algorithmic_bytes = exec.launches[i].algorithmic_bytes;
        out[i].flops = exec.launches[i].flops;
        out[i].grid[0] = exec.launches[i].gx; out[i].grid[1] = exec.launches[i].gy; out[i].grid[2] = exec.launches[i].gz;
        out[i].block = exec.launches[i].block; out[i].smem = exec.launches[i].smem;
    }
    for (int it = 0; it < iterations; ++it) {
        check(dsc_set_rand_seed(ctx_, rand_seed + (uint32_t)it));
        std::vector<float> ms;
        launch_all(exec, &ms);
        for (size_t i = 0; i < ms.size(); ++i) out[i].ms += ms[i] / (double)iterations;
    }
    return out;
}

GraphStats Environment::stats(const Graph& graph) { return prepare(graph).stats; }
std::string Environment::kernel_source(const Graph& graph) { return generate_graph_source(graph, codegen_options(), nullptr); }
CodegenOptions Environment::codegen_options() const {
    CodegenOptions opt;
    opt.sm_count = sm_count_override_ > 0 ? sm_count_override_ : sm_count_;
    opt.dp_rank = dp_.rank;
    opt.use_tf32 = use_tf32_;
    return opt;
}

// average total + the five most expensive kernels by label (timestamp.rs:155-179)
void Environment::print_timings(const std::string& label) {
    if (timing_runs_ == 0) return;
    double total = 0;
    for (auto& p : timing_totals_) total += p.second;
    std::printf("%s: %.3f ms/run over %d runs\n", label.c_str(), total / timing_runs_, timing_runs_);
    auto sorted = timing_totals_;
    std::sort(sorted.begin(), sorted.end(), [](auto& a, auto& b) { return a.second > b.second; });
    for (size_t i = 0; i < sorted.size() && i < 5; ++i) std::printf("  %8.3f ms  %s\n", sorted[i].second / timing_runs_, sorted[i].first.c_str());
    timing_totals_.clear();
    timing_runs_ = 0;
}

}  // namespace descent

// Multi-threaded CPU restatement of descent's op-graph semantics (TEST INFRASTRUCTURE: the reported CPU baseline).
//
// Only bench.py's `cpu_baseline` / `--impl reference` legs and tests/ load this library; the product path
// (descent_b200) never does.  It executes the RAW op graph the frontend exports as JSON, node by node, the way
// the reference's kernels compute each op:
//   * one "invocation" per output element, 64 invocations per work item, work items handed to a pool of host
//     threads -- the way a CPU Vulkan implementation runs a compute dispatch (kernel.rs:239 local_size_x = 64);
//   * views: coord_in[a] = clamp(offset[a] + sum(step * coord_out[i])), replicate padding (kernel.rs:89-133);
//   * Reduce: one sequential f32 loop over K per output element (kernel.rs:559-642);
//   * MatMul: [r, b, m, n] / [r, m, b, n] with k chunks of ceil(ceil(K/16)/r)*16, f32 multiply-add in ascending k
//     (kernel.rs:385-557, kernel_matmul.glsl); operands are gathered through their views once, like the
//     reference's materialised copies (graph.rs:262-284);
//   * Unpad / WindowsToImage (with the window-range check) / Gather / ScatterAdd with float
//     compare-and-swap atomics (kernel.rs:644-874); pcg hash and Rand bit-exact (kernel_common.glsl:205-216).
// Every node is materialised (the reference fuses per-element chains into one kernel; this port does not), so it
// moves more bytes than the reference's own CPU execution would.  It is checked against the numpy oracle in
// tests/test_oracle_kat.py (1e-4 relative: sums here are sequential f32, there float64).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace {

// ---- a JSON reader just big enough for the graph export -------------------------------------------
struct Json {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    double number = 0;
    bool boolean = false;
    std::string string;
    std::vector<Json> array;
    std::vector<std::pair<std::string, Json>> object;
    const Json& at(const char* key) const {
        for (const auto& kv : object)
            if (kv.first == key) return kv.second;
        throw std::runtime_error(std::string("missing key ") + key);
    }
    bool has(const char* key) const {
        for (const auto& kv : object)
            if (kv.first == key) return true;
        return false;
    }
    int64_t i64() const { return (int64_t)number; }
};

struct JsonParser {
    const char* p;
    explicit JsonParser(const char* s) : p(s) {}
    void ws() { while (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r') ++p; }
    Json parse() {
        ws();
        Json j;
        if (*p == '{') {
            j.type = Json::Object;
            ++p; ws();
            if (*p == '}') { ++p; return j; }
            for (;;) {
                ws();
                Json key = parse();
                ws();
                if (*p != ':') throw std::runtime_error("json: expected ':'");
                ++p;
                j.object.emplace_back(key.string, parse());
                ws();
                if (*p == ',') { ++p; continue; }
                if (*p == '}') { ++p; return j; }
                throw std::runtime_error("json: expected ',' or '}'");
            }
        }
        if (*p == '[') {
            j.type = Json::Array;
            ++p; ws();
            if (*p == ']') { ++p; return j; }
            for (;;) {
                j.array.push_back(parse());
                ws();
                if (*p == ',') { ++p; continue; }
                if (*p == ']') { ++p; return j; }
                throw std::runtime_error("json: expected ',' or ']'");
            }
        }
        if (*p == '"') {
            j.type = Json::String;
            ++p;
            while (*p && *p != '"') {
                if (*p == '\\' && p[1]) ++p;
                j.string.push_back(*p++);
            }
            if (*p == '"') ++p;
            return j;
        }
        if (!strncmp(p, "null", 4)) { p += 4; return j; }
        if (!strncmp(p, "true", 4)) { p += 4; j.type = Json::Bool; j.boolean = true; return j; }
        if (!strncmp(p, "false", 5)) { p += 5; j.type = Json::Bool; return j; }
        char* end = nullptr;
        j.type = Json::Number;
        j.number = strtod(p, &end);
        if (end == p) throw std::runtime_error("json: unexpected character");
        p = end;
        return j;
    }
};

// ---- thread pool: parallel_for over work items of 64 invocations ----------------------------------
class Pool {
public:
    explicit Pool(int threads) : n_(std::max(1, threads)) {
        for (int i = 1; i < n_; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; ++generation_; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // body(begin, end) over [0, count) in blocks of `grain` elements
    void run(int64_t count, int64_t grain, const std::function<void(int64_t, int64_t)>& body) {
        if (count <= 0) return;
        if (n_ == 1 || count <= grain) { body(0, count); return; }
        body_ = &body; count_ = count; grain_ = grain; next_.store(0); pending_.store(n_ - 1);
        { std::lock_guard<std::mutex> l(m_); ++generation_; }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return pending_.load() == 0; });
    }
    int threads() const { return n_; }

private:
    void work() {
        for (;;) {
            const int64_t b = next_.fetch_add(grain_);
            if (b >= count_) break;
            (*body_)(b, std::min(count_, b + grain_));
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_) return;
            }
            work();
            if (pending_.fetch_sub(1) == 1) { std::lock_guard<std::mutex> l(m_); done_.notify_all(); }
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    uint64_t generation_ = 0;
    bool stop_ = false;
    const std::function<void(int64_t, int64_t)>* body_ = nullptr;
    int64_t count_ = 0, grain_ = 1;
    std::atomic<int64_t> next_{0};
    std::atomic<int> pending_{0};
};

// ---- graph ------------------------------------------------------------------------------------------
struct View {
    std::vector<int64_t> in_shape, in_offsets, out_shape, in_strides, out_strides;
    std::vector<int> map_axis;      // -1 = broadcast
    std::vector<int64_t> map_step;
};
struct Chain {
    int64_t input_count = 0, output_count = 0;
    std::vector<View> views;  // producer side first
    // consumer element -> producer element through views [0, upto) (shape.rs:354-360, kernel.rs:89-133)
    int64_t index(int64_t e, int upto = -1) const {
        for (int vi = (upto < 0 ? (int)views.size() : upto) - 1; vi >= 0; --vi) {
            const View& v = views[vi];
            int64_t coord[8];
            for (size_t a = 0; a < v.in_shape.size(); ++a) coord[a] = v.in_offsets[a];
            for (size_t i = 0; i < v.out_shape.size(); ++i) {
                if (v.map_axis[i] < 0) continue;
                coord[v.map_axis[i]] += v.map_step[i] * ((e / v.out_strides[i]) % v.out_shape[i]);
            }
            int64_t lin = 0;
            for (size_t a = 0; a < v.in_shape.size(); ++a) lin += std::min(std::max<int64_t>(coord[a], 0), v.in_shape[a] - 1) * v.in_strides[a];
            e = lin;
        }
        return e;
    }
    // f(i, producer index) for i in [b, e): the consumer-side view is walked like an odometer (no divisions per
    // element), any further views through index()
    template <class F>
    void for_each(int64_t b, int64_t e, F&& f) const {
        if (views.empty()) {
            for (int64_t i = b; i < e; ++i) f(i, i);
            return;
        }
        const int last = (int)views.size() - 1;
        const View& v = views[last];
        const int nd = (int)v.out_shape.size(), na = (int)v.in_shape.size();
        int64_t oc[8], raw[8];
        for (int a = 0; a < na; ++a) raw[a] = v.in_offsets[a];
        for (int i = 0; i < nd; ++i) {
            oc[i] = (b / v.out_strides[i]) % v.out_shape[i];
            if (v.map_axis[i] >= 0) raw[v.map_axis[i]] += v.map_step[i] * oc[i];
        }
        for (int64_t i = b; i < e; ++i) {
            int64_t lin = 0;
            for (int a = 0; a < na; ++a) lin += std::min(std::max<int64_t>(raw[a], 0), v.in_shape[a] - 1) * v.in_strides[a];
            f(i, last == 0 ? lin : index(lin, last));
            for (int d = nd - 1; d >= 0; --d) {  // increment the odometer
                if (++oc[d] < v.out_shape[d]) {
                    if (v.map_axis[d] >= 0) raw[v.map_axis[d]] += v.map_step[d];
                    break;
                }
                if (v.map_axis[d] >= 0) raw[v.map_axis[d]] -= v.map_step[d] * (v.out_shape[d] - 1);
                oc[d] = 0;
            }
        }
    }
};
struct Edge {
    int src = -1;
    std::vector<int64_t> arg_shape;
    Chain chain;
};
struct Node {
    int id = -1;
    std::string op, kind, mode;
    std::vector<int64_t> shape;
    int64_t count = 1;
    int axis = 0, parameter = -1, uid = 0;
    int64_t pad = 0, stride_w = 1, stride_h = 1;
    uint32_t bits = 0;
    std::vector<Edge> args;
    bool live = false;
    int uses = 0;  // live consumers
};

std::vector<int64_t> strides_of(const std::vector<int64_t>& shape) {
    std::vector<int64_t> s(shape.size(), 1);
    for (int i = (int)shape.size() - 2; i >= 0; --i) s[i] = s[i + 1] * shape[i + 1];
    return s;
}
std::vector<int64_t> ints(const Json& j) {
    std::vector<int64_t> v;
    for (const auto& e : j.array) v.push_back(e.i64());
    return v;
}

inline uint32_t pcg(uint32_t v) {  // kernel_common.glsl:205-211
    const uint32_t state = v * 747796405u + 2891336453u;
    const uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}
inline float rand_from_index(uint32_t uid, uint32_t index, uint32_t seed) {  // kernel_common.glsl:213-216
    return (float)pcg(pcg(index) + seed + uid) * 2.3283064365386963e-10f;
}
inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

struct Program {
    std::vector<Node> nodes;
    std::map<int, int> pos_of;  // node id -> index
    std::map<int, const float*> inputs;
    std::map<int, float*> outputs;
    std::map<int, int64_t> input_counts, output_counts;
    // checker mode (tests only; the timed baseline leaves both off): sums of Reduce / MatMul are accumulated in float64
    // and rounded once, exactly like oracle/interp.py (_reduce, _matmul), so large-batch steps can be checked at the
    // same tolerances as the numpy interpreter; MatMul nodes listed in tf32_nodes see their operands truncated to TF32
    // (10 explicit mantissa bits, interp.tf32_operand "trunc": what tcgen05 kind::tf32 does to FP32 operands).
    bool f64_accumulate = false;
    std::vector<int> tf32_nodes;
    // Near-tie resolution (checker mode).  Two correct FP32 evaluations of the same step differ in the last bits of
    // every sum, so a CompareAndSelect whose operands are closer than that noise (a pre-activation within 1e-6 of zero
    // in leaky_relu's `x > 0`, a runner-up within 1e-6 of the window maximum in max_pool2d's `a == max`) may legitimately
    // resolve either way; in a mini-batch of thousands of samples a few always do.  For the Select nodes listed in
    // tie_nodes, elements with a != b and |a - b| <= tie_margin * max(|a|_inf, |b|_inf) are resolved as
    // tie_resolution says (0: as computed, 1: condition true, 2: condition false) and counted, so a test can evaluate
    // the band of outputs that every resolution of those ties spans.
    std::vector<int> tie_nodes;
    float tie_margin = 0.f;
    int tie_resolution = 0;
    std::atomic<int64_t> ties_seen{0};
};

thread_local std::string g_error;

Program* build(const char* text) {
    JsonParser parser(text);
    Json root = parser.parse();
    auto prog = std::make_unique<Program>();
    for (const Json& jn : root.at("nodes").array) {
        Node n;
        n.id = (int)jn.at("id").i64();
        n.op = jn.at("op").string;
        n.shape = ints(jn.at("shape"));
        for (int64_t d : n.shape) n.count *= d;
        if (jn.has("kind")) n.kind = jn.at("kind").string;
        if (jn.has("mode")) n.mode = jn.at("mode").string;
        if (jn.has("axis")) n.axis = (int)jn.at("axis").i64();
        if (jn.has("pad")) n.pad = jn.at("pad").i64();
        if (jn.has("stride_w")) n.stride_w = jn.at("stride_w").i64();
        if (jn.has("stride_h")) n.stride_h = jn.at("stride_h").i64();
        if (jn.has("parameter")) n.parameter = (int)jn.at("parameter").i64();
        if (jn.has("uid")) n.uid = (int)jn.at("uid").i64();
        if (jn.has("bits")) n.bits = (uint32_t)jn.at("bits").number;
        for (const Json& ja : jn.at("args").array) {
            Edge e;
            e.src = (int)ja.at("src").i64();
            e.arg_shape = ints(ja.at("arg_shape"));
            const Json& jc = ja.at("chain");
            e.chain.input_count = jc.at("input_count").i64();
            e.chain.output_count = jc.at("output_count").i64();
            for (const Json& jv : jc.at("views").array) {
                View v;
                v.in_shape = ints(jv.at("input_shape"));
                v.in_offsets = ints(jv.at("input_offsets"));
                v.out_shape = ints(jv.at("output_shape"));
                v.in_strides = strides_of(v.in_shape);
                v.out_strides = strides_of(v.out_shape);
                for (const Json& jm : jv.at("mapping").array) {
                    if (jm.type == Json::Null) { v.map_axis.push_back(-1); v.map_step.push_back(0); }
                    else { v.map_axis.push_back((int)jm.array[0].i64()); v.map_step.push_back(jm.array[1].i64()); }
                }
                if (v.in_shape.size() > 8) throw std::runtime_error("view with more than 8 axes");
                e.chain.views.push_back(std::move(v));
            }
            n.args.push_back(std::move(e));
        }
        prog->pos_of[n.id] = (int)prog->nodes.size();
        prog->nodes.push_back(std::move(n));
    }
    // view folding (graph.rs:255-301): a Mov is index arithmetic on its consumers' edges, never a copy.  The reference
    // keeps a copy where a reshape cannot fold into one view; here views simply chain (reshapes preserve linear
    // indices), which is the cheaper of the two for a CPU as well.
    for (auto& n : prog->nodes) {
        for (Edge& e : n.args) {
            for (;;) {
                const Node& src = prog->nodes[prog->pos_of[e.src]];
                if (src.op != "Unary" || src.kind != "Mov" || src.args.empty()) break;
                const Edge& inner = src.args[0];  // already folded: nodes are in topological order
                Chain c;
                c.input_count = inner.chain.input_count;
                c.output_count = e.chain.output_count;
                c.views = inner.chain.views;
                c.views.insert(c.views.end(), e.chain.views.begin(), e.chain.views.end());
                e.chain = std::move(c);
                e.src = inner.src;
            }
        }
    }
    // liveness (graph.rs:150-165) and consumer counts for freeing intermediates
    std::vector<int> stack;
    for (auto& n : prog->nodes)
        if (n.op == "Output") stack.push_back(n.id);
    while (!stack.empty()) {
        Node& n = prog->nodes[prog->pos_of[stack.back()]];
        stack.pop_back();
        if (n.live) continue;
        n.live = true;
        for (const Edge& e : n.args) stack.push_back(e.src);
    }
    for (auto& n : prog->nodes)
        if (n.live)
            for (const Edge& e : n.args) prog->nodes[prog->pos_of[e.src]].uses += 1;
    return prog.release();
}

struct Runner {
    Program& prog;
    Pool& pool;
    uint32_t seed;
    std::vector<std::vector<float>> values;
    std::vector<int> remaining;

    Runner(Program& p, Pool& pl, uint32_t s) : prog(p), pool(pl), seed(s), values(p.nodes.size()), remaining(p.nodes.size()) {
        for (size_t i = 0; i < p.nodes.size(); ++i) remaining[i] = p.nodes[i].uses;
    }

    // an operand as a contiguous array: the producer's own storage when the edge is an identity, else a gathered copy
    struct Operand {
        std::vector<float> owned;
        const float* p = nullptr;
        const float& operator[](int64_t i) const { return p[i]; }
        const float* data() const { return p; }
    };
    Operand operand(const Node& node, int k) {
        const Edge& e = node.args[k];
        const int sp = prog.pos_of[e.src];
        const Node& src = prog.nodes[sp];
        Operand o;
        if (e.chain.views.empty() && src.op != "Coord" && src.op != "Rand" && (int64_t)values[sp].size() == e.chain.output_count) {
            o.p = values[sp].data();
            return o;
        }
        o.owned = arg(node, k);
        o.p = o.owned.data();
        return o;
    }
    // operand `k` of `node`, gathered through its view into a contiguous array (one invocation per element)
    std::vector<float> arg(const Node& node, int k) {
        const Edge& e = node.args[k];
        const int sp = prog.pos_of[e.src];
        const Node& src = prog.nodes[sp];
        std::vector<float> out((size_t)e.chain.output_count);
        float* o = out.data();
        if (src.op == "Coord") {  // kernel.rs:266-270
            pool.run(e.chain.output_count, 4096, [&](int64_t b, int64_t en) { e.chain.for_each(b, en, [&](int64_t i, int64_t j) { o[i] = (float)j; }); });
        } else if (src.op == "Rand") {  // kernel.rs:271-279
            const uint32_t uid = (uint32_t)src.uid, sd = seed;
            pool.run(e.chain.output_count, 4096, [&](int64_t b, int64_t en) { e.chain.for_each(b, en, [&](int64_t i, int64_t j) { o[i] = rand_from_index(uid, (uint32_t)j, sd); }); });
        } else {
            const std::vector<float>& v = values[sp];
            if (v.empty()) throw std::runtime_error("node " + std::to_string(src.id) + " (" + src.op + ") has no value");
            const float* s = v.data();
            if (e.chain.views.empty()) {
                if ((int64_t)v.size() != e.chain.output_count) throw std::runtime_error("identity edge with mismatched size");
                pool.run(e.chain.output_count, 1 << 16, [&](int64_t b, int64_t en) { memcpy(o + b, s + b, (size_t)(en - b) * 4); });
            } else {
                pool.run(e.chain.output_count, 4096, [&](int64_t b, int64_t en) { e.chain.for_each(b, en, [&](int64_t i, int64_t j) { o[i] = s[j]; }); });
            }
        }
        return out;
    }
    void release_args(const Node& node) {
        for (const Edge& e : node.args) {
            const int sp = prog.pos_of[e.src];
            if (--remaining[sp] == 0) std::vector<float>().swap(values[sp]);
        }
    }

    void run() {
        const bool profile = getenv("CPU_REF_PROFILE") != nullptr;
        std::map<std::string, double> by_op;
        for (size_t pos = 0; pos < prog.nodes.size(); ++pos) {
            const Node& n = prog.nodes[pos];
            if (!n.live) continue;
            const auto t0 = std::chrono::steady_clock::now();
            eval(n, values[pos]);
            release_args(n);
            if (profile) by_op[n.op + (n.op == "Unary" ? " " + n.kind : "")] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        if (profile)
            for (const auto& kv : by_op) fprintf(stderr, "cpu_ref %-24s %.3f s\n", kv.first.c_str(), kv.second);
    }

    void eval(const Node& n, std::vector<float>& out) {
        const std::string& op = n.op;
        if (op == "Input") {
            auto it = prog.inputs.find(n.parameter);
            if (it == prog.inputs.end() || prog.input_counts[n.parameter] != n.count) throw std::runtime_error("parameter " + std::to_string(n.parameter) + " not set or wrong size");
            out.assign(it->second, it->second + n.count);
            return;
        }
        if (op == "Literal") { out.assign(1, u2f(n.bits)); return; }
        if (op == "Coord" || op == "Rand") return;  // evaluated at the consumer
        if (op == "Output") {
            auto it = prog.outputs.find(n.parameter);
            if (it == prog.outputs.end() || prog.output_counts[n.parameter] != n.count) throw std::runtime_error("output " + std::to_string(n.parameter) + " not set or wrong size");
            std::vector<float> v = arg(n, 0);
            memcpy(it->second, v.data(), (size_t)n.count * 4);
            return;
        }
        if (op == "AllReduce") { out = arg(n, 0); return; }  // single rank
        if (op == "Unary") {
            if (n.args.empty()) throw std::runtime_error("live gradient accumulator was never written");
            out = arg(n, 0);
            float* o = out.data();
            const std::string& k = n.kind;
            if (k == "Mov") return;
            std::function<float(float)> f;
            if (k == "Neg") f = [](float a) { return -a; };
            else if (k == "Sqrt") f = [](float a) { return std::sqrt(a); };
            else if (k == "Exp") f = [](float a) { return std::exp(a); };
            else if (k == "Log") f = [](float a) { return std::log(a); };
            else if (k == "Sin") f = [](float a) { return (float)std::sin((double)a); };
            else if (k == "Cos") f = [](float a) { return (float)std::cos((double)a); };
            else if (k == "UintToFloat") f = [](float a) { return (float)f2u(a); };
            else if (k == "FloatToUint") f = [](float a) { return u2f(std::isnan(a) || a <= 0.f ? 0u : a >= 4294967296.f ? 0xffffffffu : (uint32_t)a); };
            else throw std::runtime_error("unary " + k);
            pool.run(n.count, 4096, [&](int64_t b, int64_t e) { for (int64_t i = b; i < e; ++i) o[i] = f(o[i]); });
            return;
        }
        if (op == "Binary") {
            out = arg(n, 0);
            Operand rhs = operand(n, 1);
            float* o = out.data();
            const float* r = rhs.data();
            const std::string& k = n.kind;
            std::function<float(float, float)> f;
            if (k == "Add") f = [](float a, float b) { return a + b; };
            else if (k == "Sub") f = [](float a, float b) { return a - b; };
            else if (k == "Mul") f = [](float a, float b) { return a * b; };
            else if (k == "Div") f = [](float a, float b) { return a / b; };
            else if (k == "Pow") f = [](float a, float b) { return (float)std::pow((double)a, (double)b); };
            else if (k == "UAdd") f = [](float a, float b) { return u2f(f2u(a) + f2u(b)); };
            else if (k == "UMul") f = [](float a, float b) { return u2f(f2u(a) * f2u(b)); };
            else if (k == "URem") f = [](float a, float b) { return u2f(f2u(a) % std::max(f2u(b), 1u)); };
            else if (k == "UBitXor") f = [](float a, float b) { return u2f(f2u(a) ^ f2u(b)); };
            else throw std::runtime_error("binary " + k);
            pool.run(n.count, 4096, [&](int64_t b, int64_t e) { for (int64_t i = b; i < e; ++i) o[i] = f(o[i], r[i]); });
            return;
        }
        if (op == "Select") {
            Operand a = operand(n, 0), b = operand(n, 1), p = operand(n, 2), q = operand(n, 3);
            out.resize((size_t)n.count);
            float* o = out.data();
            const bool eq = n.kind == "Eq";
            if (prog.tie_margin > 0.f && std::find(prog.tie_nodes.begin(), prog.tie_nodes.end(), n.id) != prog.tie_nodes.end()) {
                float scale = 0.f;
                for (int64_t i = 0; i < n.count; ++i) scale = std::max(scale, std::max(std::fabs(a[i]), std::fabs(b[i])));
                const float band = prog.tie_margin * scale;
                const int resolution = prog.tie_resolution;
                pool.run(n.count, 4096, [&](int64_t s, int64_t e) {
                    int64_t ties = 0;
                    for (int64_t i = s; i < e; ++i) {
                        bool cond = eq ? a[i] == b[i] : a[i] > b[i];
                        if (a[i] != b[i] && std::fabs(a[i] - b[i]) <= band) {
                            ++ties;
                            if (resolution == 1) cond = true;
                            if (resolution == 2) cond = false;
                        }
                        o[i] = cond ? p[i] : q[i];
                    }
                    prog.ties_seen += ties;
                });
                return;
            }
            pool.run(n.count, 4096, [&](int64_t s, int64_t e) { for (int64_t i = s; i < e; ++i) o[i] = (eq ? a[i] == b[i] : a[i] > b[i]) ? p[i] : q[i]; });
            return;
        }
        if (op == "Reduce") {  // kernel.rs:559-642: sequential K per output element
            Operand a = operand(n, 0);
            const auto& s = n.args[0].arg_shape;
            const int64_t K = s[n.axis];
            int64_t inner = 1;
            for (size_t d = n.axis + 1; d < s.size(); ++d) inner *= s[d];
            out.resize((size_t)n.count);
            float* o = out.data();
            const bool is_max = n.kind == "Max";
            pool.run(n.count, 64, [&](int64_t b, int64_t e) {
                for (int64_t i = b; i < e; ++i) {
                    const int64_t oo = i / inner, oi = i % inner;
                    const float* p = a.data() + oo * K * inner + oi;
                    if (prog.f64_accumulate && !is_max) {
                        double acc = 0.0;
                        for (int64_t k = 0; k < K; ++k) acc += (double)p[k * inner];
                        o[i] = (float)acc;
                        continue;
                    }
                    float acc = is_max ? -INFINITY : 0.f;
                    for (int64_t k = 0; k < K; ++k) acc = is_max ? std::max(acc, p[k * inner]) : acc + p[k * inner];
                    o[i] = acc;
                }
            });
            return;
        }
        if (op == "MatMul") {  // kernel.rs:385-557
            Operand a = operand(n, 0), bm = operand(n, 1);
            const int64_t BC = n.args[0].arg_shape[0], M = n.args[0].arg_shape[1], K = n.args[0].arg_shape[2], N = n.args[1].arg_shape[2];
            const int64_t R = n.shape[0], chunk = ((K + 15) / 16 + R - 1) / R * 16;
            const bool rows = n.mode == "Rows";
            out.assign((size_t)n.count, 0.f);
            float* o = out.data();
            const bool tf32 = std::find(prog.tf32_nodes.begin(), prog.tf32_nodes.end(), n.id) != prog.tf32_nodes.end();
            if (prog.f64_accumulate || tf32) {
                auto operand_value = [tf32](float v) { return tf32 ? u2f(f2u(v) & 0xFFFFE000u) : v; };
                pool.run(R * BC * M, 16, [&](int64_t s, int64_t e) {
                    std::vector<double> acc((size_t)N);
                    std::vector<float> row((size_t)N);
                    for (int64_t i = s; i < e; ++i) {
                        const int64_t m = i % M, b = (i / M) % BC, c = i / (M * BC);
                        const int64_t lo = c * chunk, hi = std::min(K, lo + chunk);
                        std::fill(acc.begin(), acc.end(), 0.0);
                        const float* arow = a.data() + (b * M + m) * K;
                        for (int64_t k = lo; k < hi; ++k) {
                            const double av = (double)operand_value(arow[k]);
                            const float* brow = bm.data() + (b * K + k) * N;
                            for (int64_t j = 0; j < N; ++j) acc[j] += av * (double)operand_value(brow[j]);
                        }
                        for (int64_t j = 0; j < N; ++j) row[j] = (float)acc[j];
                        float* dst = rows ? o + ((c * M + m) * BC + b) * N : o + ((c * BC + b) * M + m) * N;
                        memcpy(dst, row.data(), (size_t)N * 4);
                    }
                });
                return;
            }
            pool.run(R * BC * M, 16, [&](int64_t s, int64_t e) {
                std::vector<float> acc((size_t)N);
                for (int64_t i = s; i < e; ++i) {
                    const int64_t m = i % M, b = (i / M) % BC, c = i / (M * BC);
                    const int64_t lo = c * chunk, hi = std::min(K, lo + chunk);
                    std::fill(acc.begin(), acc.end(), 0.f);
                    const float* arow = a.data() + (b * M + m) * K;
                    for (int64_t k = lo; k < hi; ++k) {
                        const float av = arow[k];
                        const float* brow = bm.data() + (b * K + k) * N;
                        for (int64_t j = 0; j < N; ++j) acc[j] += av * brow[j];
                    }
                    float* dst = rows ? o + ((c * M + m) * BC + b) * N : o + ((c * BC + b) * M + m) * N;
                    memcpy(dst, acc.data(), (size_t)N * 4);
                }
            });
            return;
        }
        if (op == "Unpad") {  // kernel.rs:644-710
            Operand a = operand(n, 0);
            const int64_t len = n.shape[n.axis], pad = n.pad;
            int64_t inner = 1;
            for (size_t d = n.axis + 1; d < n.shape.size(); ++d) inner *= n.shape[d];
            out.resize((size_t)n.count);
            float* o = out.data();
            pool.run(n.count, 4096, [&](int64_t s, int64_t e) {
                for (int64_t i = s; i < e; ++i) {
                    const int64_t oi = i % inner, x = (i / inner) % len, oo = i / (inner * len);
                    const int64_t k0 = x + pad - (x == 0 ? pad : 0), k1 = x + pad + (x == len - 1 ? pad : 0);
                    float sum = 0.f;
                    for (int64_t k = k0; k <= k1; ++k) sum += a[(oo * (len + 2 * pad) + k) * inner + oi];
                    o[i] = sum;
                }
            });
            return;
        }
        if (op == "WindowsToImage") {  // kernel.rs:712-810, with the range check
            Operand a = operand(n, 0);
            const auto& ws = n.args[0].arg_shape;
            const size_t d = ws.size();
            const int64_t OH = ws[d - 6], OW = ws[d - 5], G = ws[d - 4], FH = ws[d - 3], FW = ws[d - 2], GC = ws[d - 1];
            const size_t dn = n.shape.size();
            const int64_t IH = n.shape[dn - 3], IW = n.shape[dn - 2], IC = n.shape[dn - 1], SW = n.stride_w, SH = n.stride_h;
            out.resize((size_t)n.count);
            float* o = out.data();
            pool.run(n.count, 1024, [&](int64_t s, int64_t e) {
                for (int64_t i = s; i < e; ++i) {
                    const int64_t c = i % IC, x = (i / IC) % IW, y = (i / (IC * IW)) % IH, batch = i / (IC * IW * IH);
                    const int64_t g = c / GC, gc = c % GC;
                    float sum = 0.f;
                    for (int64_t fy = y % SH; fy < FH; fy += SH) {
                        const int64_t oy = (y - fy) / SH;
                        if (y < fy || oy >= OH) continue;
                        for (int64_t fx = x % SW; fx < FW; fx += SW) {
                            const int64_t ox = (x - fx) / SW;
                            if (x < fx || ox >= OW) continue;
                            sum += a[(((((batch * OH + oy) * OW + ox) * G + g) * FH + fy) * FW + fx) * GC + gc];
                        }
                    }
                    o[i] = sum;
                }
            });
            return;
        }
        if (op == "Gather") {  // kernel.rs:336-351
            Operand v = operand(n, 0), idx = operand(n, 1);
            const auto& vs = n.args[0].arg_shape;
            const int64_t rows = vs[n.axis], len = n.shape[n.axis];
            int64_t inner = 1;
            for (size_t d = n.axis + 1; d < n.shape.size(); ++d) inner *= n.shape[d];
            out.resize((size_t)n.count);
            float* o = out.data();
            pool.run(n.count, 4096, [&](int64_t s, int64_t e) {
                for (int64_t i = s; i < e; ++i) {
                    const int64_t oi = i % inner, oo = i / (inner * len);
                    const int64_t row = (int64_t)(int32_t)f2u(idx[i]);
                    o[i] = v[(oo * rows + row) * inner + oi];
                }
            });
            return;
        }
        if (op == "ScatterAdd") {  // kernel.rs:812-874: float atomics, order unspecified
            out = arg(n, 0);
            Operand v = operand(n, 1), idx = operand(n, 2);
            const auto& vs = n.args[1].arg_shape;
            const int64_t rows = n.shape[n.axis], len = vs[n.axis];
            int64_t inner = 1, vcount = 1;
            for (size_t d = n.axis + 1; d < n.shape.size(); ++d) inner *= n.shape[d];
            for (int64_t d : vs) vcount *= d;
            float* o = out.data();
            pool.run(vcount, 4096, [&](int64_t s, int64_t e) {
                for (int64_t i = s; i < e; ++i) {
                    const int64_t oi = i % inner, p = (i / inner) % len, oo = i / (inner * len);
                    const int64_t row = (int64_t)(int32_t)f2u(idx[oo * len + p]);
                    if (row < 0 || row >= rows) continue;
                    auto* cell = reinterpret_cast<std::atomic<uint32_t>*>(o + (oo * rows + row) * inner + oi);
                    uint32_t old = cell->load(std::memory_order_relaxed);
                    while (!cell->compare_exchange_weak(old, f2u(u2f(old) + v[i]), std::memory_order_relaxed)) {}
                }
            });
            return;
        }
        throw std::runtime_error("unknown op " + op);
    }
};

}  // namespace

extern "C" {

const char* cpu_ref_error() { return g_error.c_str(); }

void* cpu_ref_create(const char* graph_json) {
    try {
        return build(graph_json);
    } catch (const std::exception& e) {
        g_error = e.what();
        return nullptr;
    }
}
void cpu_ref_destroy(void* h) { delete static_cast<Program*>(h); }

int cpu_ref_set_input(void* h, int parameter, const float* data, int64_t count) {
    auto* p = static_cast<Program*>(h);
    p->inputs[parameter] = data;
    p->input_counts[parameter] = count;
    return 0;
}
int cpu_ref_set_output(void* h, int parameter, float* data, int64_t count) {
    auto* p = static_cast<Program*>(h);
    p->outputs[parameter] = data;
    p->output_counts[parameter] = count;
    return 0;
}
int cpu_ref_set_checker_mode(void* h, int f64_accumulate, const int* tf32_nodes, int tf32_node_count) {
    auto* p = static_cast<Program*>(h);
    p->f64_accumulate = f64_accumulate != 0;
    p->tf32_nodes.assign(tf32_nodes, tf32_nodes + tf32_node_count);
    return 0;
}
int cpu_ref_set_tie_resolution(void* h, const int* select_nodes, int count, float margin, int resolution) {
    auto* p = static_cast<Program*>(h);
    p->tie_nodes.assign(select_nodes, select_nodes + count);
    p->tie_margin = margin;
    p->tie_resolution = resolution;
    p->ties_seen = 0;
    return 0;
}
long long cpu_ref_ties_seen(void* h) { return (long long)static_cast<Program*>(h)->ties_seen.load(); }
// One Environment::run of the graph on `threads` host threads; returns the wall time in seconds, or -1.
double cpu_ref_run(void* h, uint32_t rand_seed, int threads) {
    try {
        Pool pool(threads);
        Runner runner(*static_cast<Program*>(h), pool, rand_seed);
        const auto t0 = std::chrono::steady_clock::now();
        runner.run();
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1.0;
    }
}
int cpu_ref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"

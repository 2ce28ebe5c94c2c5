// Graph passes and clustering (reference: src/graph.rs).  See graph.hpp for the design notes.
#include "graph.hpp"
#include "codegen.hpp"

#include <climits>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <functional>
#include <map>
#include <numeric>
#include <queue>
#include <set>
#include <unordered_map>

namespace descent {

// ---- small helpers ----------------------------------------------------------------------------

std::string Op::name() const {
    static const char* un[] = {"Mov", "Neg", "Sqrt", "Exp", "Log", "Sin", "Cos", "FloatToUint", "UintToFloat"};
    static const char* bi[] = {"Add", "Sub", "Mul", "Div", "Pow", "UAdd", "UMul", "URem", "UBitXor"};
    std::ostringstream os;
    switch (kind) {
        case OpKind::Input: os << "Input(" << parameter_id << ")"; break;
        case OpKind::Output: os << "Output(" << parameter_id << ")"; break;
        case OpKind::Literal:
            if (literal_is_u32) os << "U32(" << literal_bits << ")";
            else os << "F32(" << literal_f32_value() << ")";
            break;
        case OpKind::BuiltIn: os << (built_in == BuiltInOp::Coord ? "Coord" : "Rand"); break;
        case OpKind::Unary: os << un[(int)unary]; break;
        case OpKind::Binary: os << bi[(int)binary]; break;
        case OpKind::CompareAndSelect: os << (compare == CompareMode::Eq ? "SelectEq" : "SelectGt"); break;
        case OpKind::MatMul: os << "MatMul"; break;
        case OpKind::Reduce: os << (reduce == ReduceOp::Max ? "ReduceMax(" : "ReduceSum(") << axis << ")"; break;
        case OpKind::Unpad: os << "Unpad" << pad << "(" << axis << ")"; break;
        case OpKind::WindowsToImage: os << "WindowsToImage"; break;
        case OpKind::Gather: os << "Gather(" << axis << ")"; break;
        case OpKind::ScatterAdd: os << "ScatterAdd(" << axis << ")"; break;
        case OpKind::AllReduce: os << "AllReduce"; break;
    }
    return os.str();
}

Initializer Initializer::for_relu(int64_t fan_in) { return rand_normal(std::sqrt(2.0f / (float)fan_in)); }
Initializer Initializer::for_siren(int64_t fan_in, bool first) {
    return rand_uniform(std::sqrt(6.0f / (float)fan_in) * (first ? 30.0f : 1.0f));
}

std::vector<int> OpGraph::topo_order() const {
    int n = (int)nodes.size();
    std::vector<int> indeg(n, 0);
    auto cons = consumers();
    for (int i = 0; i < n; ++i)
        if (nodes[i].alive) indeg[i] = (int)nodes[i].in.size();
    std::priority_queue<int, std::vector<int>, std::greater<int>> ready;
    for (int i = 0; i < n; ++i)
        if (nodes[i].alive && indeg[i] == 0) ready.push(i);
    std::vector<int> order;
    while (!ready.empty()) {
        int i = ready.top();
        ready.pop();
        order.push_back(i);
        for (auto [d, k] : cons[i]) {
            (void)k;
            if (--indeg[d] == 0) ready.push(d);
        }
    }
    int live = 0;
    for (const auto& nd : nodes) live += nd.alive;
    DSC_CHECK((int)order.size() == live, "op graph has a cycle");
    return order;
}

static void json_shape(std::ostringstream& os, const Shape& s) {
    os << "[";
    for (int i = 0; i < s.len(); ++i) os << (i ? "," : "") << s[i];
    os << "]";
}
static void json_chain(std::ostringstream& os, const ViewChain& c) {
    os << "{\"input_count\":" << c.input_count << ",\"output_count\":" << c.output_count << ",\"views\":[";
    for (size_t vi = 0; vi < c.views.size(); ++vi) {
        const View& v = c.views[vi];
        os << (vi ? "," : "") << "{\"input_shape\":";
        json_shape(os, v.input_shape);
        os << ",\"input_offsets\":[";
        for (size_t i = 0; i < v.input_offsets.size(); ++i) os << (i ? "," : "") << v.input_offsets[i];
        os << "],\"mapping\":[";
        for (size_t i = 0; i < v.output_mapping.size(); ++i) {
            const auto& m = v.output_mapping[i];
            os << (i ? "," : "");
            if (m.is_source) os << "[" << m.axis << "," << m.step << "]";
            else os << "null";
        }
        os << "],\"output_shape\":";
        json_shape(os, v.output_shape);
        os << "}";
    }
    os << "]}";
}

std::string export_ops_json(const OpGraph& ops, const std::vector<ParameterStorage>& parameters, const std::vector<Cluster>* clusters) {
    static const char* un[] = {"Mov", "Neg", "Sqrt", "Exp", "Log", "Sin", "Cos", "FloatToUint", "UintToFloat"};
    static const char* bi[] = {"Add", "Sub", "Mul", "Div", "Pow", "UAdd", "UMul", "URem", "UBitXor"};
    std::ostringstream os;
    os << "{\"parameters\":[";
    for (size_t i = 0; i < parameters.size(); ++i) {
        os << (i ? "," : "") << "{\"id\":" << i << ",\"name\":\"" << parameters[i].name << "\",\"shape\":";
        json_shape(os, parameters[i].shape);
        os << ",\"trainable\":" << (parameters[i].reset_to.has_value() ? "true" : "false") << "}";
    }
    os << "],\"nodes\":[";
    bool first = true;
    for (int id : ops.topo_order()) {
        const OpNode& n = ops.nodes[id];
        os << (first ? "" : ",") << "{\"id\":" << id << ",\"colour\":" << n.colour << ",\"shape\":";
        first = false;
        json_shape(os, n.shape);
        const Op& op = n.op;
        os << ",\"op\":\"";
        switch (op.kind) {
            case OpKind::Input: os << "Input\",\"parameter\":" << op.parameter_id; break;
            case OpKind::Output: os << "Output\",\"parameter\":" << op.parameter_id; break;
            case OpKind::Literal: os << "Literal\",\"is_u32\":" << (op.literal_is_u32 ? "true" : "false") << ",\"bits\":" << op.literal_bits; break;
            case OpKind::BuiltIn:
                if (op.built_in == BuiltInOp::Coord) os << "Coord\"";
                else os << "Rand\",\"uid\":" << op.rand_uid;
                break;
            case OpKind::Unary: os << "Unary\",\"kind\":\"" << un[(int)op.unary] << "\""; break;
            case OpKind::Binary: os << "Binary\",\"kind\":\"" << bi[(int)op.binary] << "\""; break;
            case OpKind::CompareAndSelect: os << "Select\",\"kind\":\"" << (op.compare == CompareMode::Eq ? "Eq" : "Gt") << "\""; break;
            case OpKind::MatMul: os << "MatMul\",\"mode\":\"" << (op.output_mode == MatMulOutputMode::Batches ? "Batches" : "Rows") << "\""; break;
            case OpKind::Reduce: os << "Reduce\",\"kind\":\"" << (op.reduce == ReduceOp::Max ? "Max" : "Sum") << "\",\"axis\":" << op.axis; break;
            case OpKind::Unpad: os << "Unpad\",\"axis\":" << op.axis << ",\"pad\":" << op.pad; break;
            case OpKind::WindowsToImage: os << "WindowsToImage\",\"stride_w\":" << op.stride_w << ",\"stride_h\":" << op.stride_h; break;
            case OpKind::Gather: os << "Gather\",\"axis\":" << op.axis; break;
            case OpKind::ScatterAdd: os << "ScatterAdd\",\"axis\":" << op.axis; break;
            case OpKind::AllReduce: os << "AllReduce\""; break;
        }
        os << ",\"cluster\":" << n.cluster_id << ",\"args\":[";
        for (int a = 0; a < n.arg_count(); ++a) {
            const OpEdge* e = n.arg_edge(a);
            DSC_CHECK(e != nullptr, "node " << id << " (" << op.name() << ") is missing argument " << a);
            os << (a ? "," : "") << "{\"src\":" << e->src << ",\"arg_shape\":";
            json_shape(os, e->arg_shape);
            os << ",\"chain\":";
            json_chain(os, e->chain);
            os << "}";
        }
        os << "]}";
    }
    os << "]";
    if (clusters) {
        os << ",\"clusters\":[";
        for (size_t i = 0; i < clusters->size(); ++i) {
            const Cluster& c = (*clusters)[i];
            os << (i ? "," : "") << "{\"kind\":" << (int)c.kind << ",\"level\":" << c.level << ",\"label\":\"" << c.label << "\",\"members\":[";
            for (size_t k = 0; k < c.members.size(); ++k) os << (k ? "," : "") << c.members[k];
            os << "],\"inputs\":[";
            for (size_t k = 0; k < c.inputs.size(); ++k) os << (k ? "," : "") << c.inputs[k].node_id;
            os << "],\"outputs\":[";
            for (size_t k = 0; k < c.outputs.size(); ++k) os << (k ? "," : "") << c.outputs[k];
            os << "]}";
        }
        os << "]";
    }
    os << "}";
    return os.str();
}

std::string Scope::export_json() const { return export_ops_json(ops_, *parameters_, nullptr); }
std::string Graph::export_json() const {
    std::string json = export_ops_json(ops_, *parameters_, &clusters_);
    // dense-chain candidates (graph.hpp DenseChain), appended as one more key of the top-level object
    std::ostringstream os;
    os << ",\"dense_chains\":[";
    for (size_t i = 0; i < dense_chains_.size(); ++i) {
        const DenseChain& ch = dense_chains_[i];
        auto list = [&](const char* key, const std::vector<int>& v) {
            os << "\"" << key << "\":[";
            for (size_t j = 0; j < v.size(); ++j) os << (j ? "," : "") << v[j];
            os << "]";
        };
        os << (i ? "," : "") << "{\"rows\":" << ch.rows << ",\"widths\":[";
        for (size_t j = 0; j < ch.widths.size(); ++j) os << (j ? "," : "") << ch.widths[j];
        os << "],";
        list("forward", ch.forward); os << ",";
        list("weight_gradient", ch.weight_gradient); os << ",";
        list("backward", ch.backward); os << ",";
        list("clusters", ch.all_clusters());
        os << ",\"loss\":" << ch.loss << "}";
    }
    os << "]";
    const size_t end = json.rfind('}');
    DSC_CHECK(end != std::string::npos, "malformed graph json");
    json.insert(end, os.str());
    return json;
}

// ---- Graph ------------------------------------------------------------------------------------

Graph::Graph(SharedParameters parameters, const OpGraph& ops, DataParallel dp)
    : parameters_(std::move(parameters)), ops_(ops), dp_(dp) {
    eliminate_dead_code();
    eliminate_moves();
    simplify_arithmetic();
    eliminate_common_subgraphs();
    reuse_activation_sign();
    eliminate_dead_code();
    hoist_all_reduce_views();
    sink_permutations_into_per_element();
    sink_views_into_selects();
    build_clusters();
}

std::vector<int> Graph::input_nodes() const {
    std::vector<int> v;
    for (int i = 0; i < (int)ops_.nodes.size(); ++i)
        if (ops_.nodes[i].alive && ops_.nodes[i].op.kind == OpKind::Input) v.push_back(i);
    return v;
}
std::vector<int> Graph::output_nodes() const {
    std::vector<int> v;
    for (int i = 0; i < (int)ops_.nodes.size(); ++i)
        if (ops_.nodes[i].alive && ops_.nodes[i].op.kind == OpKind::Output) v.push_back(i);
    return v;
}

// keep only what an Output depends on (graph.rs:150-165)
void Graph::eliminate_dead_code() {
    int n = (int)ops_.nodes.size();
    std::vector<char> live(n, 0);
    std::vector<int> stack;
    for (int i = 0; i < n; ++i)
        if (ops_.nodes[i].alive && ops_.nodes[i].op.kind == OpKind::Output) { live[i] = 1; stack.push_back(i); }
    while (!stack.empty()) {
        int i = stack.back();
        stack.pop_back();
        for (const auto& e : ops_.nodes[i].in)
            if (!live[e.src]) { live[e.src] = 1; stack.push_back(e.src); }
    }
    for (int i = 0; i < n; ++i)
        if (ops_.nodes[i].alive && !live[i]) ops_.remove_node(i);
}

// Fold every Mov into the chains of its consumers.  A Mov that feeds an Output stays, because the
// Output takes over its producer's buffer (graph.rs:276-279).
void Graph::eliminate_moves() {
    auto order = ops_.topo_order();
    auto cons = ops_.consumers();
    for (int id : order) {
        OpNode& node = ops_.nodes[id];
        if (!node.alive || !node.op.is_mov()) continue;
        DSC_CHECK(node.in.size() == 1, "a value or gradient is used but nothing was ever written to it (node " << id << ", shape "
                                           << node.shape.str() << ")");
        const OpEdge in_edge = node.in[0];
        bool feeds_output = false;
        for (auto [dst, k] : cons[id]) {
            OpNode& d = ops_.nodes[dst];
            if (!d.alive) continue;
            OpEdge& out_edge = d.in[k];
            if (d.op.kind == OpKind::Output) {
                // a pure reshape of a computed value needs no copy: the executor writes the producer
                // straight into the parameter's buffer
                const OpKind sk = ops_.nodes[in_edge.src].op.kind;
                const bool computed = sk != OpKind::Input && sk != OpKind::Literal && sk != OpKind::BuiltIn;
                if (!(in_edge.chain.is_identity() && computed)) { feeds_output = true; continue; }
            }
            DSC_CHECK(out_edge.src == id, "consumer table out of date");
            ViewChain chain = in_edge.chain;
            chain.append(out_edge.chain);
            out_edge.src = in_edge.src;
            out_edge.chain = chain;
            cons[in_edge.src].push_back({dst, k});
        }
        if (!feeds_output) ops_.remove_node(id);
    }
}

// x*1, x+0 (and the u32 forms) become moves, which then fold (graph.rs:221-253)
void Graph::simplify_arithmetic() {
    bool mov_added = false;
    for (int id : ops_.topo_order()) {
        OpNode& node = ops_.nodes[id];
        if (node.op.kind != OpKind::Binary) continue;
        auto is_skip = [&](const Op& lit) {
            switch (node.op.binary) {
                case BinaryOp::Mul: return lit.is_literal_f32(1.0f);
                case BinaryOp::Add: return lit.is_literal_f32(0.0f);
                case BinaryOp::UMul: return lit.is_literal_u32(1);
                case BinaryOp::UAdd: return lit.is_literal_u32(0);
                default: return false;
            }
        };
        int skip = -1;
        for (int a = 0; a < 2 && skip < 0; ++a) {
            const OpEdge* e = node.arg_edge(a);
            if (e && is_skip(ops_.nodes[e->src].op)) skip = a;
        }
        if (skip < 0) continue;
        OpEdge keep = *node.arg_edge(1 - skip);
        keep.arg = 0;
        node.in.clear();
        node.in.push_back(keep);
        node.op = Op::mov();
        mov_added = true;
    }
    if (mov_added) eliminate_moves();
}

// leaky_relu / relu keep the sign of their argument: with L = select(x > 0, x, c*x), c >= 0, the tests `x > 0` and
// `L > 0` agree for every float (c*x <= 0 or NaN whenever x is not > 0).  The backward pass of the activation
// (array.rs leaky_relu: select(x > 0, g, c*g)) therefore tests the activation it already needs for other reasons
// (pooling masks, the next layer) instead of the pre-activation, which then no longer has to be kept in memory
// between the forward and the backward pass.  Bit-exact.
void Graph::reuse_activation_sign() {
    auto is_zero = [&](const OpEdge* e) { return e && ops_.nodes[e->src].op.is_literal_f32(0.0f); };
    auto plain = [&](const OpEdge* e, int src) { return e && e->src == src && e->chain.is_identity(); };
    std::map<int, int> activation_of;  // x -> L
    for (int id : ops_.topo_order()) {
        const OpNode& n = ops_.nodes[id];
        if (n.op.kind != OpKind::CompareAndSelect || n.op.compare != CompareMode::Gt || !is_zero(n.arg_edge(1))) continue;
        const OpEdge* x = n.arg_edge(0);
        if (!x || !x->chain.is_identity() || !plain(n.arg_edge(2), x->src) || ops_.nodes[x->src].shape != n.shape) continue;
        const OpEdge* other = n.arg_edge(3);
        bool keeps_sign = is_zero(other);
        if (!keeps_sign && other) {
            const OpNode& m = ops_.nodes[other->src];
            if (other->chain.is_identity() && m.op.kind == OpKind::Binary && m.op.binary == BinaryOp::Mul) {
                for (int a = 0; a < 2; ++a) {
                    const Op& lit = ops_.nodes[m.arg_edge(1 - a)->src].op;
                    if (plain(m.arg_edge(a), x->src) && lit.kind == OpKind::Literal && !lit.literal_is_u32 && lit.literal_f32_value() >= 0.0f &&
                        lit.literal_f32_value() < 1e30f)
                        keeps_sign = true;
                }
            }
        }
        if (keeps_sign) activation_of.emplace(x->src, id);
    }
    if (activation_of.empty()) return;
    for (int id : ops_.topo_order()) {
        OpNode& n = ops_.nodes[id];
        if (n.op.kind != OpKind::CompareAndSelect || n.op.compare != CompareMode::Gt || !is_zero(n.arg_edge(1))) continue;
        for (OpEdge& e : n.in) {
            if (e.arg != 0 || !e.chain.is_identity()) continue;
            auto it = activation_of.find(e.src);
            if (it != activation_of.end() && it->second != id && it->second < id) e.src = it->second;
        }
    }
}

// merge structurally identical nodes (graph.rs:167-203)
void Graph::eliminate_common_subgraphs() {
    std::vector<int> remap(ops_.nodes.size());
    std::iota(remap.begin(), remap.end(), 0);
    std::unordered_map<std::string, std::vector<int>> by_key;
    for (int id : ops_.topo_order()) {
        OpNode& node = ops_.nodes[id];
        for (auto& e : node.in) e.src = remap[e.src];
        if (!node.op.can_merge()) continue;
        std::sort(node.in.begin(), node.in.end(), [](const OpEdge& a, const OpEdge& b) { return a.arg < b.arg; });
        std::ostringstream key;
        key << node.op.name() << "|" << (int)node.op.kind << "|" << node.op.literal_bits << "|" << node.op.rand_uid << "|"
            << (int)node.op.output_mode << "|" << node.op.stride_w << "," << node.op.stride_h << "|" << node.shape.str();
        for (const auto& e : node.in) key << "|" << e.src << ":" << e.arg << ":" << e.chain.views.size();
        auto& bucket = by_key[key.str()];
        int found = -1;
        for (int other : bucket) {
            const OpNode& o = ops_.nodes[other];
            if (o.op == node.op && o.shape == node.shape && o.in == node.in) { found = other; break; }
        }
        if (found >= 0) {
            remap[id] = found;
            ops_.remove_node(id);
        } else {
            bucket.push_back(id);
        }
    }
}

// A cross-rank sum is elementwise, so it commutes with any view: all-reduce the producer's buffer as it is
// (e.g. the raw output of a filter-gradient GEMM) and let the consumers read the result through the view.
void Graph::hoist_all_reduce_views() {
    auto cons = ops_.consumers();
    for (int id = 0; id < (int)ops_.nodes.size(); ++id) {
        OpNode& node = ops_.nodes[id];
        if (!node.alive || node.op.kind != OpKind::AllReduce || node.in[0].chain.is_identity()) continue;
        const ViewChain view = node.in[0].chain;
        const Shape src_shape = ops_.nodes[node.in[0].src].shape;
        node.shape = src_shape;
        node.in[0].chain = ViewChain::identity(src_shape.element_count());
        node.in[0].arg_shape = src_shape;
        for (auto [dst, k] : cons[id]) {
            OpEdge& e = ops_.nodes[dst].in[k];
            ViewChain chain = view;
            chain.append(e.chain);
            e.chain = chain;
        }
    }
}

// A per-element value that is only ever read through one axis permutation by other per-element ops (max-pool's
// backward pass computes in window order [image, oy, ox, fy, fx, c] and its consumer reads it back in image order)
// would be written to memory and gathered element by element.  Per-element ops commute with any bijective
// re-indexing, so the permutation is pushed onto the op's own operands instead: the op is evaluated directly
// in its consumers' order, the edge between them becomes an identity and the two kernels fuse.  Repeats until
// the permutations have sunk to loads of values that exist in memory anyway.
static bool chain_is_permutation(const ViewChain& chain) {
    if (chain.is_identity() || chain.input_count != chain.output_count) return false;
    for (const View& v : chain.views) {
        std::vector<int> hits(v.input_shape.len(), 0);
        for (int a = 0; a < v.input_shape.len(); ++a)
            if (v.input_offsets[a] != 0) return false;
        for (int i = 0; i < v.output_shape.len(); ++i) {
            if (v.output_shape[i] == 1) continue;
            const AxisMapping& m = v.output_mapping[i];
            if (!m.is_source || m.step != 1 || v.input_shape[m.axis] != v.output_shape[i]) return false;
            hits[m.axis] += 1;
        }
        for (int a = 0; a < v.input_shape.len(); ++a)
            if (v.input_shape[a] > 1 && hits[a] != 1) return false;
    }
    return true;
}

void Graph::sink_permutations_into_per_element() {
    auto movable = [&](const OpNode& n) {
        return n.alive && n.op.is_per_element() && !n.op.is_inline_source() && n.op.kind != OpKind::Gather;
    };
    for (bool changed = true; changed;) {
        changed = false;
        auto cons = ops_.consumers();
        // every consumer is a per-element op reading through the same permutation
        auto common_permutation = [&](int id, ViewChain* chain, Shape* shape) {
            if (cons[id].empty()) return false;
            for (size_t i = 0; i < cons[id].size(); ++i) {
                const OpNode& d = ops_.nodes[cons[id][i].first];
                const OpEdge& e = d.in[cons[id][i].second];
                if (!d.op.is_per_element() || d.op.is_gather_arg(e.arg) || d.shape.element_count() != e.chain.output_count) return false;
                if (i == 0) { *chain = e.chain; *shape = e.arg_shape; }
                else if (e.chain != *chain || e.arg_shape != *shape) return false;
            }
            return chain_is_permutation(*chain);
        };
        // the op and everything fused beneath it can follow: an operand that is computed in registers for this op
        // alone sinks next; one that other ops share would have to be spilled to memory, so then nothing moves
        std::function<bool(int)> can_sink = [&](int id) {
            for (const OpEdge& e : ops_.nodes[id].in) {
                const OpNode& src = ops_.nodes[e.src];
                if (!movable(src) || !e.chain.is_identity() || src.shape.element_count() != ops_.nodes[id].shape.element_count()) continue;
                if (cons[e.src].size() != 1 || !can_sink(e.src)) return false;
            }
            return true;
        };
        for (int id = (int)ops_.nodes.size() - 1; id >= 0; --id) {
            OpNode& node = ops_.nodes[id];
            if (!movable(node)) continue;
            ViewChain perm;
            Shape shape;
            if (!common_permutation(id, &perm, &shape) || !can_sink(id)) continue;
            for (OpEdge& e : node.in) {
                e.chain.append(perm);
                e.arg_shape = shape;
            }
            node.shape = shape;
            for (auto [dst, k] : cons[id]) ops_.nodes[dst].in[k].chain = ViewChain::identity(shape.element_count());
            changed = true;
            break;  // consumer lists are stale
        }
    }
}

static bool edge_is_fusable(const OpGraph& ops, int dst, const OpEdge& e);

// Concatenation chains.  `a.concat(b, axis)` (array.rs: two pads and a CompareAndSelect on the axis coordinate) reads both
// operands through clamping pad views, and a chain of k concats -- the ten per-level feature pairs of a hash grid joined
// into one [m, 20] MLP input, examples/image_fit/main.rs:201-215 -- re-reads and re-writes the growing prefix k times
// (nine kernels and 226 MB for [262144, 20]).  A select whose only reader is another select reading it through a view
// is evaluated in that reader's index space instead: the view is appended to the select's own operand chains (loads of
// arrays that exist in memory anyway, literals, coordinates), the edge becomes an identity and the two fuse; repeated
// along the chain, all k selects become ONE kernel that reads each piece once and writes the result once.  Values are
// selected, never recomputed differently: bit-exact.
void Graph::sink_views_into_selects() {
    for (bool changed = true; changed;) {
        changed = false;
        auto cons = ops_.consumers();
        for (int id = (int)ops_.nodes.size() - 1; id >= 0; --id) {
            OpNode& node = ops_.nodes[id];
            if (!node.alive || node.op.kind != OpKind::CompareAndSelect || cons[id].size() != 1) continue;
            const auto [dst, k] = cons[id][0];
            OpNode& reader = ops_.nodes[dst];
            OpEdge& edge = reader.in[k];
            if (!reader.alive || reader.op.kind != OpKind::CompareAndSelect || edge.chain.is_identity() ||
                reader.shape.element_count() != edge.chain.output_count || edge.chain.output_count > 4 * edge.chain.input_count)
                continue;
            // operands computed in registers for this select would have to be spilled to memory: then nothing moves
            bool ok = true;
            for (const OpEdge& e : node.in) {
                const OpNode& src = ops_.nodes[e.src];
                if (!src.op.is_inline_source() && src.op.kind != OpKind::Input && edge_is_fusable(ops_, id, e)) ok = false;
            }
            if (!ok) continue;
            const ViewChain view = edge.chain;
            const Shape shape = edge.arg_shape;
            for (OpEdge& e : node.in) {
                e.chain.append(view);
                e.arg_shape = shape;
            }
            node.shape = shape;
            edge.chain = ViewChain::identity(shape.element_count());
            changed = true;
            break;  // consumer lists are stale
        }
    }
}

// ---- clustering -------------------------------------------------------------------------------

static bool edge_is_fusable(const OpGraph& ops, int dst, const OpEdge& e) {
    const OpNode& s = ops.nodes[e.src];
    const OpNode& d = ops.nodes[dst];
    return s.op.is_per_element() && d.op.is_per_element() && !d.op.is_gather_arg(e.arg) && e.chain.is_identity() &&
           s.shape.element_count() == d.shape.element_count();
}

void Graph::build_per_element_program(Cluster& c) {
    std::map<int, int> member_op_index;
    auto find_input = [&](const ClusterInput& in) {
        for (size_t i = 0; i < c.inputs.size(); ++i)
            if (c.inputs[i] == in) return (int)i;
        c.inputs.push_back(in);
        return (int)c.inputs.size() - 1;
    };
    struct ArgKey { int src; ViewChain chain; Shape arg_shape; int op_index; };
    std::vector<ArgKey> loaded;
    auto cons = ops_.consumers();
    for (int id : c.members) {
        const OpNode& node = ops_.nodes[id];
        PerElementOp pe;
        pe.op = node.op;
        pe.shape = node.shape;
        for (int a = 0; a < node.arg_count(); ++a) {
            const OpEdge* e = node.arg_edge(a);
            DSC_CHECK(e != nullptr, "missing argument");
            auto m = member_op_index.find(e->src);
            if (m != member_op_index.end() && edge_is_fusable(ops_, id, *e)) { pe.args[a] = m->second; continue; }
            const OpNode& src = ops_.nodes[e->src];
            if (node.op.is_gather_arg(a)) {
                DSC_CHECK(!src.op.is_inline_source(), "gather from a literal/built-in is not supported");
                pe.input_index = find_input({e->src, e->chain, e->arg_shape});
                pe.arg_shape = e->arg_shape;
                continue;
            }
            int op_index = -1;
            for (const auto& l : loaded)
                if (l.src == e->src && l.chain == e->chain && l.arg_shape == e->arg_shape) op_index = l.op_index;
            if (op_index < 0) {
                PerElementOp ld;
                if (src.op.kind == OpKind::Literal) {
                    ld.kind = PerElementOp::Literal;
                    ld.op = src.op;
                } else if (src.op.kind == OpKind::BuiltIn) {
                    ld.kind = PerElementOp::BuiltIn;
                    ld.op = src.op;
                    ld.chain = e->chain;
                    ld.arg_shape = src.shape;
                } else {
                    ld.kind = PerElementOp::Load;
                    ld.input_index = find_input({e->src, e->chain, e->arg_shape});
                }
                op_index = (int)c.ops.size();
                c.ops.push_back(ld);
                loaded.push_back({e->src, e->chain, e->arg_shape, op_index});
            }
            pe.args[a] = op_index;
        }
        switch (node.op.kind) {
            case OpKind::Unary: pe.kind = PerElementOp::Unary; break;
            case OpKind::Binary: pe.kind = PerElementOp::Binary; break;
            case OpKind::CompareAndSelect: pe.kind = PerElementOp::Select; break;
            case OpKind::Gather: pe.kind = PerElementOp::Gather; break;
            default: fail("unexpected op in per-element cluster");
        }
        int op_index = (int)c.ops.size();
        c.ops.push_back(pe);
        member_op_index[id] = op_index;
        bool needs_store = false;
        for (auto [dst, k] : cons[id]) {
            const OpNode& d = ops_.nodes[dst];
            if (!d.alive) continue;
            if (d.cluster_id != node.cluster_id || !edge_is_fusable(ops_, dst, d.in[k])) needs_store = true;
        }
        if (needs_store) {
            c.outputs.push_back(id);
            c.output_ops.push_back(op_index);
        }
    }
    std::ostringstream label;
    label << "PerElement (" << c.ops.size() << " ops) [" << c.element_count << "]";
    c.label = label.str();
}


// conv2d / dense forward: MatMul -> (+ bias, activation, ...) as one kernel.  Runs after the per-element programs
// exist; the merged cluster takes the per-element cluster's level (all of its other inputs are ready there, and the
// MatMul's operands stay alive until then because the planner follows the final cluster order).
void Graph::absorb_per_element_epilogues(std::vector<Cluster>& clusters) {
    auto cons = ops_.consumers();
    for (size_t mi = 0; mi < clusters.size(); ++mi) {
        Cluster& mc = clusters[mi];
        if (mc.kind != ClusterKind::MatMul || mc.matmul_absorbs_reduce || mc.conv_backward_input.enabled || !mc.epilogue.empty()) continue;
        const int x = mc.outputs[0];
        const OpNode& mm = ops_.nodes[x];
        if (mm.op.kind != OpKind::MatMul || mm.shape[0] != 1 || mm.shape.at(-1) % 4 != 0 || cons[x].empty()) continue;
        int pi = -1;
        bool ok = true;
        for (auto [dst, k] : cons[x]) {
            const OpNode& d = ops_.nodes[dst];
            const OpEdge& e = d.in[k];
            if (d.cluster_id < 0 || clusters[d.cluster_id].kind != ClusterKind::PerElement || !e.chain.is_identity() || d.op.is_gather_arg(e.arg) ||
                (pi >= 0 && d.cluster_id != pi)) { ok = false; break; }
            pi = d.cluster_id;
        }
        if (!ok || pi < 0) continue;
        Cluster& pc = clusters[pi];
        if (pc.element_count != mm.shape.element_count() || pc.members.empty()) continue;
        int product_input = -1;
        for (size_t i = 0; i < pc.inputs.size() && ok; ++i) {
            if (pc.inputs[i].node_id != x) continue;
            if (!pc.inputs[i].chain.is_identity() || product_input >= 0) ok = false;
            product_input = (int)i;
        }
        // results that are stored straight into a parameter stay in their own kernel: the in-place update rules of
        // the planner are stated per per-element kernel
        for (int out : pc.outputs)
            for (auto [dst, k] : cons[out])
                if (ops_.nodes[dst].op.kind == OpKind::Output) ok = false;
        for (const auto& op : pc.ops)
            if (op.kind == PerElementOp::Gather && op.input_index == product_input) ok = false;
        if (!ok || product_input < 0) continue;
        mc.epilogue.push_back(pc);
        mc.epilogue_product_input = product_input;
        for (size_t i = 0; i < pc.inputs.size(); ++i)
            if ((int)i != product_input) mc.inputs.push_back(pc.inputs[i]);
        mc.outputs = pc.outputs;
        for (int id : pc.members) {
            mc.members.push_back(id);
            ops_.nodes[id].cluster_id = (int)mi;
        }
        mc.level = pc.level;
        mc.label += " + " + pc.label;
        pc.members.clear();  // dropped below
        pc.outputs.clear();
    }
}

// conv2d -> activation -> max_pool2d: the pooling Reduce joins the convolution's cluster (see Cluster::Pool).
void Graph::absorb_max_pools(std::vector<Cluster>& clusters) {
    auto cons = ops_.consumers();
    for (size_t mi = 0; mi < clusters.size(); ++mi) {
        Cluster& mc = clusters[mi];
        if (mc.kind != ClusterKind::MatMul || mc.epilogue.empty() || mc.outputs.size() != 1 || mc.pool.enabled) continue;
        const int x = mc.outputs[0];
        const OpNode& xn = ops_.nodes[x];
        if (xn.shape.len() != 4) continue;
        const int64_t B = xn.shape[0], H = xn.shape[1], W = xn.shape[2], C = xn.shape[3];
        for (auto [dst, k] : cons[x]) {
            const OpNode& r = ops_.nodes[dst];
            const OpEdge& e = r.in[k];
            if (r.op.kind != OpKind::Reduce || r.op.reduce != ReduceOp::Max || r.cluster_id < 0) continue;
            Cluster& rc = clusters[r.cluster_id];
            if (rc.kind != ClusterKind::Reduce || rc.members.size() != 1 || e.arg_shape.len() != 3 || r.op.axis != 1 || e.arg_shape[2] != C) continue;
            const int64_t O = e.arg_shape[0], K = e.arg_shape[1];
            bool found = false;
            int64_t ph = 0, pw = 0;
            for (ph = 1; ph <= K && !found; ++ph) {
                if (K % ph != 0) continue;
                pw = K / ph;
                if (H % ph != 0 || W % pw != 0 || O != B * (H / ph) * (W / pw)) continue;
                const int64_t PH = H / ph, PW = W / pw;
                auto expected = [&](int64_t o, int64_t kk, int64_t c) {
                    const int64_t image = o / (PH * PW), py = (o / PW) % PH, px = o % PW, wy = kk / pw, wx = kk % pw;
                    return ((image * H + py * ph + wy) * W + px * pw + wx) * C + c;
                };
                bool ok = true;
                for (int64_t o : {(int64_t)0, (int64_t)1, PW, PW + 1, PH * PW - 1, std::min<int64_t>(PH * PW + PW + 1, O - 1), O / 2, O - 1})
                    for (int64_t kk = 0; kk < K && ok; ++kk)
                        for (int64_t c : {(int64_t)0, C - 1})
                            if (eval_chain(e.chain, (o * K + kk) * C + c) != expected(o, kk, c)) ok = false;
                if (ok) { found = true; break; }
            }
            if (!found || ph * pw < 2) continue;
            mc.pool.enabled = true;
            mc.pool.images = B; mc.pool.height = H; mc.pool.width = W; mc.pool.channels = C;
            mc.pool.window_h = ph; mc.pool.window_w = pw;
            mc.pool.reduce.push_back(rc);
            mc.outputs.push_back(dst);
            mc.members.push_back(dst);
            mc.label += " +MaxPool";
            ops_.nodes[dst].cluster_id = (int)mi;
            rc.members.clear();
            rc.outputs.clear();
            break;
        }
    }
}

// dW = A^T x dY and db = column sums of dY read the same array: when the graph reduces the GEMM's B operand [K, C]
// over all of K (reduce_sum axis by axis, array.rs) and nothing needs the result before the GEMM's turn, the Reduce
// chain joins the MatMul's cluster (see Cluster::column_sum).
void Graph::absorb_column_sums(std::vector<Cluster>& clusters) {
    auto cons = ops_.consumers();
    for (size_t mi = 0; mi < clusters.size(); ++mi) {
        Cluster& mc = clusters[mi];
        if (mc.kind != ClusterKind::MatMul || mc.members.empty() || !mc.epilogue.empty() || mc.conv_backward_input.enabled || !mc.column_sum.empty()) continue;
        const ClusterInput& b = mc.inputs[1];
        const int64_t G = b.arg_shape[0], K = b.arg_shape[1], N = b.arg_shape[2], C = G * N;
        const OpNode& y = ops_.nodes[b.node_id];
        if (K < 4096 || y.shape.element_count() != K * C || y.shape.at(-1) != C) continue;
        // B[g, k, n] must be Y[k * C + g * N + n]
        if (b.chain.views.size() > 1 || (!b.chain.views.empty() && b.chain.views[0].any_clamp())) continue;
        if (eval_chain(b.chain, 0) != 0 || (G > 1 && eval_chain(b.chain, K * N) != N) || (K > 1 && eval_chain(b.chain, N) != C) ||
            (N > 1 && eval_chain(b.chain, 1) != 1))
            continue;
        // follow Reduce(sum) nodes from Y while each has a single consumer; stop when K has been summed away
        std::vector<int> chain;
        int cur = b.node_id;
        int64_t reduced = 1;
        while (reduced < K) {
            int next = -1;
            for (auto [dst, k] : cons[cur]) {
                const OpNode& d = ops_.nodes[dst];
                const OpEdge& e = d.in[k];
                if (d.op.kind == OpKind::Reduce && d.op.reduce == ReduceOp::Sum && e.chain.is_identity() && d.cluster_id >= 0 &&
                    clusters[d.cluster_id].kind == ClusterKind::Reduce && clusters[d.cluster_id].members.size() == 1 &&
                    d.op.axis < e.arg_shape.len() - 1 && e.arg_shape.at(-1) == C && (cur == b.node_id || cons[cur].size() == 1)) {
                    next = dst;
                    reduced *= e.arg_shape[d.op.axis];
                    break;
                }
            }
            if (next < 0) break;
            chain.push_back(next);
            cur = next;
        }
        if (chain.empty() || reduced != K || ops_.nodes[cur].shape.element_count() != C) continue;
        // nobody may need the sums before this cluster runs, and they must not be stored straight into a parameter
        bool ok = true;
        for (auto [dst, k] : cons[cur]) {
            const OpNode& d = ops_.nodes[dst];
            if (d.op.kind == OpKind::Output || (d.cluster_id >= 0 && clusters[d.cluster_id].level <= mc.level)) ok = false;
        }
        if (!ok) continue;
        for (int id : chain) {
            Cluster& rc = clusters[ops_.nodes[id].cluster_id];
            mc.column_sum.push_back(rc);
            rc.members.clear();
            rc.outputs.clear();
            mc.members.push_back(id);
            ops_.nodes[id].cluster_id = (int)mi;
        }
        mc.outputs.push_back(cur);
        mc.label += " +ColumnSum";
    }
}

// See Cluster::value_programs.
void Graph::absorb_scatter_values(std::vector<Cluster>& clusters) {
    auto cons = ops_.consumers();
    std::set<int> rebuilt;
    for (size_t si = 0; si < clusters.size(); ++si) {
        Cluster& sc = clusters[si];
        if (sc.kind != ClusterKind::ScatterAdd || sc.members.empty()) continue;
        const int nsrc = (int)sc.members.size();
        sc.value_programs.assign(nsrc, Cluster());
        sc.value_input_base.assign(nsrc, -1);
        for (int s = 0; s < nsrc; ++s) {
            const ClusterInput vin = sc.inputs[2 * s];
            const int v = vin.node_id;
            OpNode& vn = ops_.nodes[v];
            if (!vin.chain.is_identity() || !vn.op.is_per_element() || vn.op.is_inline_source() || vn.op.kind == OpKind::Gather || vn.cluster_id < 0) continue;
            if (cons[v].size() != 1 || ops_.nodes[cons[v][0].first].cluster_id != (int)si) continue;
            Cluster& pc = clusters[vn.cluster_id];
            if (pc.kind != ClusterKind::PerElement || !pc.group.empty() || pc.members.empty()) continue;
            // a leaf of its cluster: every operand is loaded (from memory, a literal, a coordinate), none computed beside it
            bool leaf = true;
            for (const OpEdge& e : vn.in) {
                const OpNode& src = ops_.nodes[e.src];
                if (src.op.kind == OpKind::BuiltIn && src.op.built_in != BuiltInOp::Coord) leaf = false;  // (rand: keep the kernel)
                if (!src.op.is_inline_source() && src.op.kind != OpKind::Input && src.cluster_id == vn.cluster_id && edge_is_fusable(ops_, v, e)) leaf = false;
            }
            if (!leaf) continue;
            const int pci = vn.cluster_id;
            pc.members.erase(std::remove(pc.members.begin(), pc.members.end(), v), pc.members.end());
            rebuilt.insert(pci);
            vn.cluster_id = (int)si;
            Cluster vp;
            vp.kind = ClusterKind::PerElement;
            vp.level = sc.level;
            vp.element_count = vn.shape.element_count();
            vp.members = {v};
            build_per_element_program(vp);
            DSC_CHECK(vp.outputs.size() == 1 && vp.outputs[0] == v, "scatter value program must produce exactly the values");
            sc.value_input_base[s] = (int)sc.inputs.size();
            for (const auto& in : vp.inputs) sc.inputs.push_back(in);
            sc.value_programs[s] = vp;
        }
    }
    for (int pci : rebuilt) {
        Cluster& pc = clusters[pci];
        pc.ops.clear();
        pc.inputs.clear();
        pc.outputs.clear();
        pc.output_ops.clear();
        if (!pc.members.empty()) build_per_element_program(pc);
    }
}

// Launch-bound tails (a few thousand elements per kernel, microseconds each) are merged per dependency level: the
// programs are independent by construction of the levels.  A program that reads a parameter is never grouped with
// one that writes the same parameter: the planner lets a kernel update a parameter in place when it reads the old
// value at the same element, which only holds inside one program.
// Row fusion (north_star (c): softmax cross-entropy as one kernel).  loss.rs:4-34 builds softmax cross-entropy, its
// accuracy and its gradient from per-element ops on [m, classes] / [m, 1] arrays joined by reductions along the class
// axis (max, sum of exponentials, the picked log-probability, argmax): each reduction is a level boundary, so the
// level schedule runs it as ten kernels of a few microseconds.  Here every maximal set of
//   * per-element clusters whose nodes all have shape [R, K] ("wide") or [R, 1] ("narrow"), K <= 32, and
//   * Reduce clusters over the last axis of a [R, K] array,
// connected by identity edges, by broadcasts of a narrow value along the row, or by a reduction's own input edge, becomes
// ONE Row cluster: one thread per row, the row's K values in registers, reductions as sequential loops in ascending k
// (the reference kernel's own order, kernel.rs:559-642).  External operands keep their view chains.  The set must be
// convex (no operand of the fused kernel may depend on one of its results); consumers are pushed to later levels.
void Graph::fuse_rows(std::vector<Cluster>& clusters) {
    auto cons = ops_.consumers();
    const int nc = (int)clusters.size();
    // candidate reductions define the (R, K) families
    std::set<std::pair<int64_t, int64_t>> families;
    for (const Cluster& c : clusters) {
        if (c.kind != ClusterKind::Reduce || c.members.size() != 1) continue;
        const OpNode& r = ops_.nodes[c.node_id];
        const OpEdge& e = r.in[0];
        const int64_t K = e.arg_shape.at(-1);
        if (r.op.axis != e.arg_shape.len() - 1 || K < 2 || K > 32 || !e.chain.is_identity()) continue;
        families.insert({e.arg_shape.element_count() / K, K});
    }
    for (auto [R, K] : families) {
        if (R < 2) continue;
        auto width_of = [&, R = R, K = K](const OpNode& n) {  // 2 wide, 1 narrow, 0 neither
            if (n.shape.element_count() == R * K && n.shape.at(-1) == K) return 2;
            if (n.shape.element_count() == R && n.shape.at(-1) == 1) return 1;
            return 0;
        };
        auto broadcast_along_row = [&, K = K](const OpEdge& e) {  // consumer element e reads producer element e / K
            if (e.chain.is_identity() || e.chain.output_count % K != 0 || e.chain.input_count != e.chain.output_count / K) return false;
            const int64_t count = e.chain.output_count;
            for (int64_t i = 0; i < count; i += std::max<int64_t>(1, count / 997))
                if (eval_chain(e.chain, i) != i / K) return false;
            for (int64_t i = std::max<int64_t>(0, count - 2 * K); i < count; ++i)
                if (eval_chain(e.chain, i) != i / K) return false;
            return true;
        };
        // eligible clusters
        std::vector<char> eligible(nc, 0);
        for (int ci = 0; ci < nc; ++ci) {
            const Cluster& c = clusters[ci];
            if (c.members.empty()) continue;
            if (c.kind == ClusterKind::Reduce && c.members.size() == 1) {
                const OpNode& r = ops_.nodes[c.node_id];
                const OpEdge& e = r.in[0];
                eligible[ci] = r.op.axis == e.arg_shape.len() - 1 && e.arg_shape.at(-1) == K && e.arg_shape.element_count() == R * K && width_of(r) == 1;
            } else if (c.kind == ClusterKind::PerElement && c.group.empty()) {
                bool ok = true;
                for (int id : c.members) {
                    const OpNode& n = ops_.nodes[id];
                    ok = ok && width_of(n) != 0 && n.op.kind != OpKind::Gather;
                }
                for (const auto& op : c.ops) ok = ok && op.kind != PerElementOp::Gather;
                eligible[ci] = ok;
            }
        }
        // union clusters across fusable internal edges
        std::vector<int> parent(nc);
        std::iota(parent.begin(), parent.end(), 0);
        std::function<int(int)> find = [&](int x) { return parent[x] == x ? x : parent[x] = find(parent[x]); };
        auto internal_edge_ok = [&](int dst, const OpEdge& e) {
            const OpNode& d = ops_.nodes[dst];
            const OpNode& src = ops_.nodes[e.src];
            const int ws = width_of(src), wd = width_of(d);
            if (d.op.kind == OpKind::Reduce) return ws == 2 && e.chain.is_identity();
            if (ws == wd && e.chain.is_identity()) return true;
            return ws == 1 && wd == 2 && broadcast_along_row(e);
        };
        for (int ci = 0; ci < nc; ++ci) {
            if (!eligible[ci]) continue;
            for (int id : clusters[ci].members)
                for (const auto& e : ops_.nodes[id].in) {
                    const int sc = ops_.nodes[e.src].cluster_id;
                    if (sc < 0 || sc == ci || !eligible[sc]) continue;
                    if (ops_.nodes[id].op.is_gather_arg(e.arg)) continue;
                    if (internal_edge_ok(id, e)) parent[find(ci)] = find(sc);
                }
        }
        std::map<int, std::vector<int>> components;
        for (int ci = 0; ci < nc; ++ci)
            if (eligible[ci]) components[find(ci)].push_back(ci);
        for (auto& [root, ids] : components) {
            (void)root;
            bool has_reduce = false;
            for (int ci : ids) has_reduce |= clusters[ci].kind == ClusterKind::Reduce;
            if (!has_reduce || ids.size() < 3) continue;
            std::set<int> member_clusters(ids.begin(), ids.end());
            std::vector<int> members;  // nodes, topological (node ids are created in topological order per cluster; merge by order below)
            for (int id : ops_.topo_order())
                if (ops_.nodes[id].alive && ops_.nodes[id].cluster_id >= 0 && member_clusters.count(ops_.nodes[id].cluster_id)) members.push_back(id);
            std::set<int> member_nodes(members.begin(), members.end());
            // every edge between two members must be one the kernel can keep in registers
            bool ok = true;
            for (int id : members)
                for (const auto& e : ops_.nodes[id].in)
                    if (member_nodes.count(e.src) && !internal_edge_ok(id, e)) ok = false;
            // convexity: no external operand may depend on a member
            int min_level = INT32_MAX, max_level = 0;
            for (int ci : ids) { min_level = std::min(min_level, clusters[ci].level); max_level = std::max(max_level, clusters[ci].level); }
            std::set<int> visited;
            std::function<bool(int)> reaches_member = [&](int id) {
                if (member_nodes.count(id)) return true;
                if (!visited.insert(id).second) return false;
                const OpNode& n = ops_.nodes[id];
                if (n.cluster_id >= 0 && clusters[n.cluster_id].level < min_level) return false;  // scheduled before any member
                for (const auto& e : n.in)
                    if (reaches_member(e.src)) return true;
                return false;
            };
            for (int id : members)
                for (const auto& e : ops_.nodes[id].in)
                    if (!member_nodes.count(e.src) && reaches_member(e.src)) ok = false;
            if (!ok) continue;

            // build the program
            Cluster row;
            row.kind = ClusterKind::Row;
            row.level = max_level;
            row.rows = R;
            row.row_length = K;
            row.members = members;
            std::map<int, int> op_of;  // member node -> op index
            struct Loaded { int src; ViewChain chain; Shape arg_shape; bool wide; int op_index; };
            std::vector<Loaded> loaded;
            auto find_input = [&](const ClusterInput& in) {
                for (size_t i = 0; i < row.inputs.size(); ++i)
                    if (row.inputs[i] == in) return (int)i;
                row.inputs.push_back(in);
                return (int)row.inputs.size() - 1;
            };
            for (int id : members) {
                const OpNode& node = ops_.nodes[id];
                PerElementOp pe;
                pe.op = node.op;
                pe.shape = node.shape;
                pe.wide = width_of(node) == 2;
                for (int a = 0; a < node.arg_count(); ++a) {
                    const OpEdge* e = node.arg_edge(a);
                    DSC_CHECK(e != nullptr, "missing argument");
                    auto m = op_of.find(e->src);
                    if (m != op_of.end()) { pe.args[a] = m->second; continue; }
                    const OpNode& src = ops_.nodes[e->src];
                    const bool arg_wide = node.op.kind == OpKind::Reduce ? true : pe.wide;
                    int op_index = -1;
                    for (const auto& l : loaded)
                        if (l.src == e->src && l.chain == e->chain && l.arg_shape == e->arg_shape && l.wide == arg_wide) op_index = l.op_index;
                    if (op_index < 0) {
                        PerElementOp ld;
                        ld.wide = arg_wide;
                        if (src.op.kind == OpKind::Literal) {
                            ld.kind = PerElementOp::Literal;
                            ld.op = src.op;
                            ld.wide = false;  // a constant
                        } else if (src.op.kind == OpKind::BuiltIn) {
                            ld.kind = PerElementOp::BuiltIn;
                            ld.op = src.op;
                            ld.chain = e->chain;
                            ld.arg_shape = src.shape;
                        } else {
                            ld.kind = PerElementOp::Load;
                            ld.input_index = find_input({e->src, e->chain, e->arg_shape});
                        }
                        op_index = (int)row.ops.size();
                        row.ops.push_back(ld);
                        loaded.push_back({e->src, e->chain, e->arg_shape, arg_wide, op_index});
                    }
                    pe.args[a] = op_index;
                }
                switch (node.op.kind) {
                    case OpKind::Unary: pe.kind = PerElementOp::Unary; break;
                    case OpKind::Binary: pe.kind = PerElementOp::Binary; break;
                    case OpKind::CompareAndSelect: pe.kind = PerElementOp::Select; break;
                    case OpKind::Reduce: pe.kind = PerElementOp::Reduce; pe.wide = false; break;
                    default: fail("unexpected op in row cluster");
                }
                const int op_index = (int)row.ops.size();
                row.ops.push_back(pe);
                op_of[id] = op_index;
                bool needs_store = false;
                for (auto [dst, k] : cons[id]) {
                    (void)k;
                    if (ops_.nodes[dst].alive && !member_nodes.count(dst)) needs_store = true;
                }
                if (needs_store) {
                    row.outputs.push_back(id);
                    row.output_ops.push_back(op_index);
                }
            }
            std::ostringstream label;
            label << "Row (" << row.ops.size() << " ops, " << ids.size() << " fused kernels) [" << R << ", " << K << "]";
            row.label = label.str();
            // the first member cluster becomes the row cluster, the others are emptied (dropped by the caller)
            const int keep = ids[0];
            for (int ci : ids) {
                clusters[ci].members.clear();
                clusters[ci].outputs.clear();
            }
            clusters[keep] = row;
            for (int id : members) ops_.nodes[id].cluster_id = keep;
        }
    }
    // consumers of a fused kernel's results must sit on later levels than the fused kernel itself
    bool changed = true;
    while (changed) {
        changed = false;
        int ar_level = -1;
        for (const auto& node : ops_.nodes) {
            if (!node.alive || node.cluster_id < 0) continue;
            Cluster& dc = clusters[node.cluster_id];
            for (const auto& e : node.in) {
                const OpNode& src = ops_.nodes[e.src];
                if (!src.alive || src.cluster_id < 0 || src.cluster_id == node.cluster_id) continue;
                const Cluster& sc = clusters[src.cluster_id];
                if (dc.level <= sc.level) { dc.level = sc.level + 1; changed = true; }
            }
        }
        (void)ar_level;
    }
}

// Multi-tensor optimiser step (north_star (c): SGD / Adam as one hand-scheduled kernel, optimizer.rs:62-112).  The
// reference's optimisers emit one per-element update per parameter tensor, which the level schedule places right
// behind that tensor's gradient: eight (conv-net) or sixteen (multi-hash) launches of a few microseconds each spread
// over the backward pass.  A per-element cluster whose every result goes straight into a parameter (theta, m, v, the
// running loss / accuracy sums) has no consumer inside the step, so it can run at the very end: all of them move to
// one final level, where group_small_per_element makes them ONE launch with a pointer table (block ranges select the
// tensor).  In-place updates stay valid: every other read of those parameters is now earlier.
void Graph::sink_parameter_updates(std::vector<Cluster>& clusters) {
    auto cons = ops_.consumers();
    int last_level = 0;
    for (const Cluster& c : clusters)
        if (!c.members.empty()) last_level = std::max(last_level, c.level);
    for (size_t ci = 0; ci < clusters.size(); ++ci) {
        Cluster& c = clusters[ci];
        if (c.kind != ClusterKind::PerElement || c.members.empty() || c.outputs.empty()) continue;
        bool only_parameter_writes = true;
        for (int out : c.outputs)
            for (auto [dst, k] : cons[out]) {
                const OpNode& d = ops_.nodes[dst];
                if (d.alive && d.op.kind != OpKind::Output && d.cluster_id != (int)ci) only_parameter_writes = false;  // members read each other in registers
            }
        if (only_parameter_writes) c.level = last_level + 1;
    }
}

void Graph::group_small_per_element(std::vector<Cluster>& clusters) {
    constexpr int64_t kSmall = 1 << 18;  // elements
    int final_level = 0;
    for (const Cluster& c : clusters)
        if (!c.members.empty()) final_level = std::max(final_level, c.level);
    auto cons = ops_.consumers();
    auto params_read = [&](const Cluster& c) {
        std::set<int> ids;
        for (const auto& in : c.inputs)
            if (ops_.nodes[in.node_id].op.kind == OpKind::Input) ids.insert(ops_.nodes[in.node_id].op.parameter_id);
        return ids;
    };
    auto params_written = [&](const Cluster& c) {
        std::set<int> ids;
        for (int out : c.outputs)
            for (auto [dst, k] : cons[out])
                if (ops_.nodes[dst].op.kind == OpKind::Output) ids.insert(ops_.nodes[dst].op.parameter_id);
        return ids;
    };
    std::map<int, std::vector<size_t>> by_level;
    for (size_t i = 0; i < clusters.size(); ++i) {
        const Cluster& c = clusters[i];
        if (c.kind == ClusterKind::PerElement && !c.members.empty() && c.group.empty() && c.element_count <= kSmall) by_level[c.level].push_back(i);
    }
    for (auto& [level, ids] : by_level) {
        // the final level holds the sunk parameter updates (~8 buffers per tensor): one launch for up to ~24 tensors
        const size_t kMaxBuffers = level == final_level ? 200 : 40;
        std::vector<size_t> open;  // indices of group heads at this level
        for (size_t i : ids) {
            const auto reads = params_read(clusters[i]), writes = params_written(clusters[i]);
            bool placed = false;
            for (size_t head : open) {
                Cluster& h = clusters[head];
                if (h.inputs.size() + h.outputs.size() + clusters[i].inputs.size() + clusters[i].outputs.size() > kMaxBuffers) continue;
                bool conflict = false;
                for (const Cluster& sub : h.group) {
                    const auto r2 = params_read(sub), w2 = params_written(sub);
                    for (int p : reads) conflict |= w2.count(p) > 0;
                    for (int p : writes) conflict |= r2.count(p) > 0 || w2.count(p) > 0;
                }
                if (conflict) continue;
                Cluster sub = clusters[i];
                h.group.push_back(sub);
                h.inputs.insert(h.inputs.end(), sub.inputs.begin(), sub.inputs.end());
                h.outputs.insert(h.outputs.end(), sub.outputs.begin(), sub.outputs.end());
                for (int id : sub.members) {
                    h.members.push_back(id);
                    ops_.nodes[id].cluster_id = (int)head;
                }
                h.element_count += sub.element_count;
                clusters[i].members.clear();
                clusters[i].outputs.clear();
                placed = true;
                break;
            }
            if (!placed) {
                Cluster& h = clusters[i];
                Cluster self = h;
                h.group.push_back(self);  // a group of one until someone joins
                open.push_back(i);
            }
        }
        for (size_t head : open) {
            Cluster& h = clusters[head];
            if (h.group.size() == 1) { h.group.clear(); continue; }  // stays an ordinary per-element cluster
            std::ostringstream label;
            label << "PerElementGroup (" << h.group.size() << " programs) [" << h.element_count << "]";
            h.label = label.str();
            h.ops.clear();
            h.output_ops.clear();
        }
    }
}

// The Unpad(s) that undo conv2d's replicate padding run in the epilogue of the fused backward-input kernel: the
// padded image gradient is never written either.
bool Graph::absorb_unpad(std::vector<Cluster>& clusters, const std::vector<std::vector<std::pair<int, int>>>& cons, int id) {
    OpNode& node = ops_.nodes[id];
    const OpEdge& e = node.in[0];
    const OpNode& src = ops_.nodes[e.src];
    if (!e.chain.is_identity() || src.cluster_id < 0 || cons[e.src].size() != 1 || node.shape.len() != 4) return false;
    Cluster& mc = clusters[src.cluster_id];
    auto& cbi = mc.conv_backward_input;
    if (mc.kind != ClusterKind::MatMul || !cbi.enabled || mc.outputs[0] != e.src || node.op.pad < 1) return false;
    if (node.op.axis == 1 && cbi.unpad_h == 0) cbi.unpad_h = node.op.pad;
    else if (node.op.axis == 2 && cbi.unpad_w == 0) cbi.unpad_w = node.op.pad;
    else return false;
    if (cbi.unpad_first_axis == 0) cbi.unpad_first_axis = node.op.axis;
    mc.members.push_back(id);
    mc.outputs[0] = id;
    mc.label += "+Unpad";
    node.cluster_id = src.cluster_id;
    return true;
}

// conv2d's backward-input pass is MatMul (windows gradient = dY x W^T) -> view -> WindowsToImage (col2im).  When
// the MatMul feeds nothing else, the pair is one implicit GEMM over the pixels of the (padded) image gradient:
// the window matrix (filter_h*filter_w times the size of dY) never touches memory.  Returns true after
// rewriting the MatMul's cluster in place; the caller skips the stand-alone WindowsToImage kernel.
bool Graph::absorb_windows_to_image(std::vector<Cluster>& clusters, const std::vector<std::vector<std::pair<int, int>>>& cons, int id) {
    OpNode& node = ops_.nodes[id];
    const bool strided = node.op.stride_w != 1 || node.op.stride_h != 1;
    const OpEdge& e = node.in[0];
    const OpNode& mm = ops_.nodes[e.src];
    if (mm.op.kind != OpKind::MatMul || mm.cluster_id < 0 || cons[e.src].size() != 1 || mm.shape[0] != 1) return false;
    Cluster& mc = clusters[mm.cluster_id];
    if (mc.kind != ClusterKind::MatMul || mc.matmul_absorbs_reduce || mc.conv_backward_input.enabled || mc.outputs[0] != e.src) return false;
    const Shape& ws = e.arg_shape;  // [image, out_h, out_w, group, filter_h, filter_w, channel in group]
    if (ws.len() != 7 || node.shape.len() != 4) return false;
    const int64_t B = ws[0], OH = ws[1], OW = ws[2], G = ws[3], FH = ws[4], FW = ws[5], GC = ws[6];
    const int64_t IH = node.shape[1], IW = node.shape[2];
    const ClusterInput a = mc.inputs[0], b = mc.inputs[1];
    const int64_t M = B * OH * OW, N = FH * FW * GC, K = a.arg_shape[2];
    if (a.arg_shape.len() != 3 || b.arg_shape.len() != 3 || a.arg_shape[0] != G || a.arg_shape[1] != M || b.arg_shape[2] != N) return false;
    if (node.shape[0] != B || node.shape[3] != G * GC) return false;
    // the view between the two must be an affine map (at most one view, no clamping) that sends window element
    // (image, oy, ox, group, fy, fx, c) to row (image, oy, ox) / column (fy, fx, c) of the group's product
    if (e.chain.views.size() > 1 || (!e.chain.views.empty() && e.chain.views[0].any_clamp())) return false;
    if (e.chain.input_count != G * M * N || e.chain.output_count != G * M * N) return false;
    const bool rows_mode = mm.op.output_mode == MatMulOutputMode::Rows;
    const int64_t stride_group = rows_mode ? N : M * N, stride_row = rows_mode ? G * N : N;
    const int64_t want[7] = {OH * OW * stride_row, OW * stride_row, stride_row, stride_group, FW * GC, GC, 1};
    auto ws_strides = ws.strides();
    if (eval_chain(e.chain, 0) != 0) return false;
    for (int axis = 0; axis < 7; ++axis)
        if (ws[axis] > 1 && eval_chain(e.chain, ws_strides[axis]) != want[axis]) return false;

    if (strided) {
        // see ConvBackwardInput::strided: only for tiny per-group products (depthwise-like), where a gather over the few
        // windows that contain a pixel beats a GEMM that writes the window matrix
        if (K * GC > 16 || FH * FW > 25) return false;
        mc.conv_backward_input = {true, IH, IW, OH, OW, FH, FW, K, {a, b}};
        mc.conv_backward_input.strided = true;
        mc.conv_backward_input.stride_h = node.op.stride_h;
        mc.conv_backward_input.stride_w = node.op.stride_w;
        mc.members.push_back(id);
        mc.outputs[0] = id;
        std::ostringstream label;
        label << "MatMul+WindowsToImage/" << node.op.stride_h << "x" << node.op.stride_w << " (k=" << K << ") " << node.shape.str();
        mc.label = label.str();
        node.cluster_id = mm.cluster_id;
        return true;
    }
    // A'[group, (image, y, x), (fy, fx, k)] = A[group, (image, y - fy, x - fx), k]; positions outside the window
    // grid are clamped here (so the address is always legal) and zeroed by the kernel's validity test
    View va;
    va.input_shape = Shape({G, B, OH, OW, K});
    va.input_offsets.assign(5, 0);
    va.output_shape = Shape({G, B, IH, IW, FH, FW, K});
    va.output_mapping = {AxisMapping::identity(0, G), AxisMapping::identity(1, B), AxisMapping::source(2, 1), AxisMapping::source(3, 1),
                         AxisMapping::source(2, -1), AxisMapping::source(3, -1), AxisMapping::identity(4, K)};
    // B'[group, (fy, fx, k), c] = B[group, k, (fy, fx, c)]
    View vb;
    vb.input_shape = Shape({G, K, FH, FW, GC});
    vb.input_offsets.assign(5, 0);
    vb.output_shape = Shape({G, FH, FW, K, GC});
    vb.output_mapping = {AxisMapping::identity(0, G), AxisMapping::identity(2, FH), AxisMapping::identity(3, FW), AxisMapping::identity(1, K),
                         AxisMapping::identity(4, GC)};
    mc.inputs[0].chain.views.push_back(va);  // a fresh boundary: never folded into the operand's own views
    mc.inputs[0].chain.output_count = va.output_shape.element_count();
    mc.inputs[0].arg_shape = Shape({G, B * IH * IW, FH * FW * K});
    mc.inputs[1].chain.push(vb);
    mc.inputs[1].arg_shape = Shape({G, FH * FW * K, GC});
    mc.conv_backward_input = {true, IH, IW, OH, OW, FH, FW, K, {a, b}};
    mc.members.push_back(id);
    mc.outputs[0] = id;
    std::ostringstream label;
    label << "MatMul+WindowsToImage (k=" << FH * FW * K << ") " << node.shape.str();
    mc.label = label.str();
    node.cluster_id = mm.cluster_id;
    return true;
}

void Graph::build_clusters() {
    auto order = ops_.topo_order();
    auto cons = ops_.consumers();
    int n = (int)ops_.nodes.size();

    // MatMul -> Reduce(Sum, axis 0) pairs become one GEMM (the reduce is the split-K sum, array.rs:515)
    std::vector<int> absorbed_by(n, -1);  // reduce node -> matmul node
    for (int id : order) {
        const OpNode& node = ops_.nodes[id];
        if (node.op.kind != OpKind::Reduce || node.op.reduce != ReduceOp::Sum || node.op.axis != 0) continue;
        const OpEdge& e = node.in[0];
        const OpNode& src = ops_.nodes[e.src];
        if (src.op.kind == OpKind::MatMul && e.chain.is_identity() && cons[e.src].size() == 1) absorbed_by[id] = e.src;
    }

    auto edge_cost = [&](int dst, const OpEdge& e) {
        const OpNode& s = ops_.nodes[e.src];
        if (s.op.kind == OpKind::Input || s.op.is_inline_source()) return 0;
        if (absorbed_by[dst] == e.src) return 0;
        if (ops_.nodes[dst].op.kind == OpKind::Output) return 0;
        return edge_is_fusable(ops_, dst, e) ? 0 : 1;
    };

    // as-soon-as-possible levels.  Gradient AllReduce nodes are pinned to at most two levels = two buckets (SURVEY.md
    // section 8e asks for one bucketed all-reduce; one bucket leaves the whole collective exposed between the last
    // weight gradient and the optimiser, 66 us per conv-net step at 2-8 GPUs in round 1): the EARLY bucket holds every
    // gradient that is ready no later than the largest gradient tensor (the dense layers' weights, computed first in the
    // backward pass) and is reduced on a side stream while the rest of the backward pass runs; the LATE bucket holds
    // the remaining (small) gradients.  Everything that reads an all-reduced gradient waits for both.
    std::vector<int> asap(n, 0), ar_target(n, 0);
    int ar_consumer_level = 0;
    auto forward = [&]() {
        for (int id : order) {
            int lv = 0;
            for (const auto& e : ops_.nodes[id].in) {
                lv = std::max(lv, asap[e.src] + edge_cost(id, e));
                if (ops_.nodes[e.src].op.kind == OpKind::AllReduce) lv = std::max(lv, ar_consumer_level);
            }
            if (ops_.nodes[id].op.kind == OpKind::AllReduce) lv = std::max(lv, ar_target[id]);
            asap[id] = lv;
        }
    };
    forward();
    {
        std::vector<int> ar_nodes;
        for (int id : order)
            if (ops_.nodes[id].op.kind == OpKind::AllReduce) ar_nodes.push_back(id);
        if (!ar_nodes.empty()) {
            // Two buckets.  The split is the level at which the LARGEST gradient tensor is ready (conv-net: the dense layers'
            // weights, computed first in the backward pass and 98 % of the bytes): everything ready by then is reduced on the
            // side stream under the rest of the backward pass.  When the largest tensor is itself among the last to be ready
            // (multi-hash: a hash table, behind the scatter_adds) that rule leaves one exposed bucket; then every gradient
            // that is ready before the last level goes early instead (the MLP's, reduced under the table scatter).
            int late_level = 0, largest = ar_nodes[0];
            for (int id : ar_nodes) {
                late_level = std::max(late_level, asap[id]);
                if (ops_.nodes[id].shape.element_count() > ops_.nodes[largest].shape.element_count()) largest = id;
            }
            int split = asap[largest];
            if (split == late_level) {
                split = -1;
                for (int id : ar_nodes)
                    if (asap[id] < late_level) split = std::max(split, asap[id]);
            }
            int early_level = 0;
            bool any_late = false;
            for (int id : ar_nodes) {
                if (asap[id] <= split) early_level = std::max(early_level, asap[id]);
                else any_late = true;
            }
            any_late = any_late && split >= 0;
            for (int id : ar_nodes) ar_target[id] = (any_late && asap[id] <= split) ? early_level : late_level;
            ar_consumer_level = late_level + 1;
            forward();
        }
    }

    // as-late-as-possible levels: producers move next to their first consumer
    std::vector<int> level(n, 0);
    for (auto it = order.rbegin(); it != order.rend(); ++it) {
        int id = *it;
        const OpNode& node = ops_.nodes[id];
        int lv = INT32_MAX;
        for (auto [dst, k] : cons[id])
            if (ops_.nodes[dst].alive) lv = std::min(lv, level[dst] - edge_cost(dst, ops_.nodes[dst].in[k]));
        if (lv == INT32_MAX || node.op.kind == OpKind::AllReduce || node.op.kind == OpKind::Input) lv = asap[id];
        // scalars (the step counter, Adam's bias-corrected step size) cost nothing to keep alive: computed as early as
        // possible they share one kernel instead of one per level
        if (node.op.is_per_element() && node.shape.element_count() == 1) lv = asap[id];
        DSC_CHECK(lv >= asap[id], "level inversion");
        level[id] = lv;
    }

    // per-element clusters: connected components of fusable edges within a level, then independent components of
    // one level with the same element count are fused horizontally (they cannot depend on each other: any
    // non-fusable edge climbs a level), which turns e.g. the ten hash-grid levels' identical index kernels into one
    // launch.  Both level assignments are valid schedules; the one that yields fewer kernels is used.
    std::vector<int> parent(n);
    std::function<int(int)> find = [&](int x) { return parent[x] == x ? x : parent[x] = find(parent[x]); };
    auto cluster_with = [&](const std::vector<int>& lv) {
        std::iota(parent.begin(), parent.end(), 0);
        for (int id : order)
            for (const auto& e : ops_.nodes[id].in)
                if (edge_is_fusable(ops_, id, e) && lv[id] == lv[e.src]) parent[find(id)] = find(e.src);
        // buffers a component binds: external producers + members someone outside reads
        std::map<int, std::set<int>> ext_in, ext_out;
        for (int id : order) {
            const OpNode& node = ops_.nodes[id];
            if (!node.op.is_per_element()) continue;
            const int root = find(id);
            for (const auto& e : node.in) {
                const OpNode& s = ops_.nodes[e.src];
                if (s.op.is_inline_source()) continue;
                if (!(s.op.is_per_element() && find(e.src) == root && edge_is_fusable(ops_, id, e))) ext_in[root].insert(e.src);
            }
            for (auto [dst, k] : cons[id]) {
                const OpNode& d = ops_.nodes[dst];
                if (!(d.op.is_per_element() && find(dst) == root && edge_is_fusable(ops_, dst, d.in[k]))) ext_out[root].insert(id);
            }
        }
        std::map<std::pair<int, int64_t>, std::pair<int, int>> open_group;  // (level, count) -> (root, buffers so far)
        int kernels = 0;
        for (int id : order) {
            const OpNode& node = ops_.nodes[id];
            if (!node.op.is_per_element() || find(id) != id) continue;
            const int buffers = (int)(ext_in[id].size() + ext_out[id].size());
            auto key = std::make_pair(lv[id], node.shape.element_count());
            auto it = open_group.find(key);
            if (it != open_group.end() && it->second.second + buffers <= 40) {
                parent[id] = it->second.first;
                it->second.second += buffers;
            } else {
                open_group[key] = {id, buffers};
                kernels += 1;
            }
        }
        return kernels;
    };
    {
        const int with_alap = cluster_with(level);
        const int with_asap = cluster_with(asap);
        if (with_asap < with_alap) level = asap;
        else cluster_with(level);
    }

    std::map<int, int> root_to_cluster;
    std::vector<Cluster> clusters;
    for (int id : order) {
        OpNode& node = ops_.nodes[id];
        if (node.op.is_per_element()) {
            int root = find(id);
            auto it = root_to_cluster.find(root);
            if (it == root_to_cluster.end()) {
                Cluster c;
                c.kind = ClusterKind::PerElement;
                c.level = level[id];
                c.element_count = node.shape.element_count();
                clusters.push_back(c);
                it = root_to_cluster.emplace(root, (int)clusters.size() - 1).first;
            }
            node.cluster_id = it->second;
            clusters[it->second].members.push_back(id);
            continue;
        }
        Cluster c;
        c.level = level[id];
        c.node_id = id;
        std::ostringstream label;
        auto add_input = [&](const OpEdge& e) { c.inputs.push_back({e.src, e.chain, e.arg_shape}); };
        switch (node.op.kind) {
            case OpKind::Reduce: {
                if (absorbed_by[id] >= 0) {
                    // joins the cluster of its MatMul, created earlier in topological order
                    int mm_cluster = ops_.nodes[absorbed_by[id]].cluster_id;
                    node.cluster_id = mm_cluster;
                    clusters[mm_cluster].members.push_back(id);
                    clusters[mm_cluster].outputs[0] = id;
                    clusters[mm_cluster].matmul_absorbs_reduce = true;
                    clusters[mm_cluster].level = level[id];
                    continue;
                }
                c.kind = ClusterKind::Reduce;
                add_input(node.in[0]);
                label << "Reduce (k=" << node.in[0].arg_shape[node.op.axis] << ") " << node.shape.str();
                break;
            }
            case OpKind::MatMul: {
                c.kind = ClusterKind::MatMul;
                add_input(*node.arg_edge(0));
                add_input(*node.arg_edge(1));
                label << "MatMul (k=" << node.arg_edge(0)->arg_shape.at(-1) << ") " << node.shape.str();
                break;
            }
            case OpKind::Unpad:
                if (absorb_unpad(clusters, cons, id)) {
                    clusters[node.cluster_id].level = level[id];
                    continue;
                }
                c.kind = ClusterKind::Unpad;
                add_input(node.in[0]);
                label << "Unpad " << node.shape.str();
                break;
            case OpKind::WindowsToImage:
                if (absorb_windows_to_image(clusters, cons, id)) {
                    clusters[node.cluster_id].level = level[id];
                    continue;
                }
                c.kind = ClusterKind::WindowsToImage;
                add_input(node.in[0]);
                label << "WindowsToImage " << node.shape.str();
                break;
            case OpKind::ScatterAdd: {
                c.kind = ClusterKind::ScatterAdd;
                const OpEdge* acc = node.arg_edge(0);
                DSC_CHECK(acc->chain.is_identity() || ops_.nodes[acc->src].op.kind == OpKind::Literal,
                          "scatter_add accumulator must be a plain array or a broadcast literal");
                {
                    // scatter_adds chained on one table (the four corners of a hash-grid cell,
                    // examples/image_fit/main.rs:190-199) become one kernel: its chunk list is the concatenation of the
                    // sources in chain order, so every row sees exactly the same addition order as the chained ops.
                    const OpNode& acc_node = ops_.nodes[acc->src];
                    if (acc_node.op.kind == OpKind::ScatterAdd && acc->chain.is_identity() && cons[acc->src].size() == 1 &&
                        acc_node.cluster_id >= 0 && acc_node.op.axis == node.op.axis && acc_node.shape == node.shape) {
                        Cluster& prev = clusters[acc_node.cluster_id];
                        const size_t at = 2 * prev.members.size();
                        const OpEdge* v = node.arg_edge(1);
                        const OpEdge* ix = node.arg_edge(2);
                        prev.inputs.insert(prev.inputs.begin() + at, ClusterInput{ix->src, ix->chain, ix->arg_shape});
                        prev.inputs.insert(prev.inputs.begin() + at, ClusterInput{v->src, v->chain, v->arg_shape});
                        prev.members.push_back(id);
                        prev.outputs[0] = id;
                        prev.level = level[id];
                        node.cluster_id = acc_node.cluster_id;
                        continue;
                    }
                }
                add_input(*node.arg_edge(1));
                add_input(*node.arg_edge(2));
                if (ops_.nodes[acc->src].op.kind != OpKind::Literal) add_input(*acc);
                c.copy_from = acc->src;
                label << "ScatterAdd " << node.arg_edge(1)->arg_shape.str();
                break;
            }
            case OpKind::AllReduce:
                c.kind = ClusterKind::AllReduce;
                DSC_CHECK(node.in[0].chain.is_identity(), "all-reduce input must be a plain array");
                add_input(node.in[0]);
                label << "AllReduce " << node.shape.str();
                break;
            default: continue;  // Input / Output / Literal / BuiltIn own no kernel
        }
        c.members.push_back(id);
        c.outputs.push_back(id);
        c.label = label.str();
        node.cluster_id = (int)clusters.size();
        clusters.push_back(c);
    }
    for (auto& c : clusters)
        if (c.kind == ClusterKind::PerElement) build_per_element_program(c);
    absorb_per_element_epilogues(clusters);
    absorb_column_sums(clusters);
    absorb_max_pools(clusters);
    absorb_scatter_values(clusters);
    fuse_rows(clusters);
    schedule_after_dense_chains(clusters);
    sink_parameter_updates(clusters);
    group_small_per_element(clusters);

    // levels are a topological order of clusters: fusable edges stay inside a cluster, all others climb
    std::vector<int> idx(clusters.size());
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return clusters[a].level < clusters[b].level; });
    std::vector<int> new_index(clusters.size());
    for (size_t i = 0; i < idx.size(); ++i) new_index[idx[i]] = (int)i;
    clusters_.clear();
    for (size_t i = 0, kept = 0; i < idx.size(); ++i) {  // clusters absorbed into an epilogue have no members left
        if (clusters[idx[i]].members.empty()) { new_index[idx[i]] = -1; continue; }
        new_index[idx[i]] = (int)kept++;
        clusters_.push_back(clusters[idx[i]]);
    }
    for (auto& node : ops_.nodes)
        if (node.alive && node.cluster_id >= 0) node.cluster_id = new_index[node.cluster_id];
    find_operand_prologues();
    // the final list: a candidate stands only if every reader of its results runs after its last cluster
    dense_chains_.clear();
    {
        auto cons = ops_.consumers();
        for (const auto& ch : detect_dense_chains(clusters_)) {
            const auto all = ch.all_clusters();
            std::set<int> members(all.begin(), all.end());
            bool ok = true;
            for (int c : all)
                for (int out : clusters_[c].outputs)
                    for (auto [dst, k] : cons[out]) {
                        (void)k;
                        const int dc = ops_.nodes[dst].cluster_id;
                        if (ops_.nodes[dst].alive && dc >= 0 && !members.count(dc) && dc <= all.back()) ok = false;
                    }
            if (ok) dense_chains_.push_back(ch);
        }
    }
}

std::vector<int> DenseChain::all_clusters() const {
    std::vector<int> all;
    for (int c : forward) all.push_back(c);
    for (int c : weight_gradient) all.push_back(c);
    for (int c : backward) if (c >= 0) all.push_back(c);
    if (!loss_in_epilogue) all.push_back(loss);
    for (const auto& s : sums) { all.push_back(s.row_reduce); all.push_back(s.batch_reduce); }
    std::sort(all.begin(), all.end());
    return all;
}

int DenseChain::last_cluster() const { return all_clusters().back(); }

// See DenseChain (graph.hpp).  A purely structural match on the clusters the other passes produced; anything that does
// not fit exactly is left alone.  `clusters` may be the working list of build_clusters (emptied clusters are skipped) or
// the final one: node.cluster_id indexes it either way.
std::vector<DenseChain> Graph::detect_dense_chains(const std::vector<Cluster>& clusters) const {
    std::vector<DenseChain> found;
    auto cons = ops_.consumers();
    const int nc = (int)clusters.size();
    auto cluster_of = [&](int node) { return ops_.nodes[node].cluster_id; };
    // element (i, j) of a [1, rows, cols] operand reads element (j, i) of a [cols, rows]-shaped source
    auto is_transposed = [&](const ClusterInput& in, int64_t rows, int64_t cols) {
        if (in.arg_shape.len() != 3 || in.arg_shape[0] != 1 || in.arg_shape[1] != rows || in.arg_shape[2] != cols) return false;
        if (in.chain.input_count != rows * cols || in.chain.output_count != rows * cols) return false;
        for (int64_t i : {(int64_t)0, (int64_t)1, rows / 2, rows - 1})
            for (int64_t j : {(int64_t)0, (int64_t)1, cols / 3, cols - 1})
                if (i < rows && j < cols && eval_chain(in.chain, i * cols + j) != j * rows + i) return false;
        return true;
    };
    auto is_plain = [&](const ClusterInput& in, int64_t rows, int64_t cols) {
        return in.chain.is_identity() && in.arg_shape.element_count() == rows * cols && in.arg_shape.at(-1) == cols &&
               ops_.nodes[in.node_id].shape.element_count() == rows * cols;
    };
    auto plain_matmul = [&](const Cluster& c) {
        return c.kind == ClusterKind::MatMul && !c.members.empty() && !c.conv_backward_input.enabled && !c.pool.enabled && c.inputs.size() >= 2;
    };
    auto is_parameter = [&](int node) { return ops_.nodes[node].op.kind == OpKind::Input; };
    // forward layer candidates: [M, K] activation (plain) x [K, N] parameter (plain)
    struct Forward { int64_t m, k, n; int a, w, out; bool terminal; };  // terminal: the loss program is this cluster's epilogue
    std::map<int, Forward> forward;  // cluster -> shape
    for (int ci = 0; ci < nc; ++ci) {
        const Cluster& c = clusters[ci];
        if (!plain_matmul(c) || c.matmul_absorbs_reduce || !c.column_sum.empty() || c.outputs.empty()) continue;
        if (c.outputs.size() != 1 && (c.epilogue.empty() || c.epilogue[0].outputs.size() != c.outputs.size())) continue;
        const OpNode& mm = ops_.nodes[c.node_id];
        if (mm.op.kind != OpKind::MatMul || mm.shape.len() != 4 || mm.shape[0] != 1 || mm.shape[1] != 1) continue;
        const int64_t M = mm.shape[2], N = mm.shape[3], K = c.inputs[0].arg_shape.at(-1);
        if (!is_plain(c.inputs[0], M, K) || !is_plain(c.inputs[1], K, N) || !is_parameter(c.inputs[1].node_id) || M < 1024) continue;
        bool ok = true;
        if (!c.epilogue.empty()) {
            const Cluster& p = c.epilogue[0];
            ok = p.outputs.size() >= 1 && p.element_count == M * N;
            for (size_t i = 0; i < p.inputs.size(); ++i)
                if ((int)i != c.epilogue_product_input) ok = ok && is_parameter(p.inputs[i].node_id);
            for (const auto& op : p.ops) ok = ok && op.kind != PerElementOp::Gather && op.kind != PerElementOp::BuiltIn;
        }
        const bool terminal = c.outputs.size() > 1;  // a two-output epilogue can only be the loss (gradient + summed values)
        if (ok) forward[ci] = {M, K, N, c.inputs[0].node_id, c.inputs[1].node_id, terminal ? -1 : c.outputs[0], terminal};
    }
    std::map<int, int> forward_reading;  // activation node -> forward cluster that takes it as A
    for (auto& [ci, f] : forward) {
        if (forward_reading.count(f.a)) forward_reading[f.a] = -1;  // two readers: not a chain
        else forward_reading[f.a] = ci;
    }
    std::set<int> produced_by_forward;
    for (auto& [ci, f] : forward) if (!f.terminal) produced_by_forward.insert(f.out);
    for (auto& [start, f0] : forward) {
        if (produced_by_forward.count(f0.a) && forward_reading.count(f0.a) && forward_reading[f0.a] >= 0) {
            // not the first layer of its chain
            bool has_pred = false;
            for (auto& [cj, fj] : forward) has_pred |= fj.out == f0.a;
            if (has_pred) continue;
        }
        DenseChain ch;
        ch.rows = f0.m;
        ch.widths.push_back(f0.k);
        for (int ci = start;;) {
            const Forward& f = forward[ci];
            if (f.m != ch.rows || f.k != ch.widths.back()) break;
            ch.forward.push_back(ci);
            ch.widths.push_back(f.n);
            if (f.terminal) break;
            auto it = forward_reading.find(f.out);
            if (it == forward_reading.end() || it->second < 0 || !forward.count(it->second)) break;
            ci = it->second;
        }
        const int L = (int)ch.forward.size();
        if (L < 2) continue;
        const int64_t M = ch.rows;
        bool ok = true;
        std::vector<int> act(L + 1, -1), dz(L, -1);  // act[l] = input of layer l (act[L] = last product), dz[l] = gradient at layer l's output
        act[0] = forward[ch.forward[0]].a;
        for (int l = 0; l < L; ++l) act[l + 1] = forward[ch.forward[l]].out;
        // the loss program: the epilogue of the last layer (when the graph's passes absorbed it there), else the per-element
        // cluster that is the only reader of the last product
        ch.loss_in_epilogue = forward[ch.forward[L - 1]].terminal;
        if (ch.loss_in_epilogue) {
            ch.loss = ch.forward[L - 1];
        } else {
            std::set<int> readers;
            for (auto [dst, k] : cons[act[L]]) {
                (void)k;
                if (!ops_.nodes[dst].alive) continue;
                if (ops_.nodes[dst].op.kind == OpKind::Output || cluster_of(dst) < 0) ok = false;
                else readers.insert(cluster_of(dst));
            }
            if (!ok || readers.size() != 1) continue;
            ch.loss = *readers.begin();
        }
        const Cluster& lc = ch.loss_in_epilogue ? clusters[ch.loss].epilogue[0] : clusters[ch.loss];
        if (lc.kind != ClusterKind::PerElement || !lc.group.empty() || lc.element_count != M * ch.widths[L]) continue;
        if (!ch.loss_in_epilogue) {
            int product_inputs = 0;
            for (const auto& in : lc.inputs) {
                if (in.node_id == act[L]) { product_inputs += 1; ok = ok && in.chain.is_identity(); }
                else ok = ok && is_parameter(in.node_id);
            }
            if (product_inputs != 1) ok = false;
        }
        for (const auto& op : lc.ops) ok = ok && op.kind != PerElementOp::Gather && op.kind != PerElementOp::BuiltIn;
        if (!ok) continue;
        // weight gradients and backward products, last layer first
        ch.weight_gradient.assign(L, -1);
        ch.backward.assign(L, -1);
        for (int l = L - 1; l >= 0 && ok; --l) {
            const int64_t K = ch.widths[l], N = ch.widths[l + 1];
            const int w = forward[ch.forward[l]].w;
            int g = -1;
            for (auto [dst, k] : cons[act[l]]) {
                (void)k;
                const int ci = cluster_of(dst);
                if (ci < 0 || !plain_matmul(clusters[ci]) || !clusters[ci].matmul_absorbs_reduce || !clusters[ci].epilogue.empty()) continue;
                if (clusters[ci].inputs[0].node_id == act[l] && is_transposed(clusters[ci].inputs[0], K, M) && is_plain(clusters[ci].inputs[1], M, N)) g = ci;
            }
            if (g < 0) { ok = false; break; }
            const Cluster& gc = clusters[g];
            ch.weight_gradient[l] = g;
            const int d = gc.inputs[1].node_id;
            if (l == L - 1) {
                dz[l] = d;
                ch.loss_gradient_output = -1;
                for (size_t i = 0; i < lc.outputs.size(); ++i)
                    if (lc.outputs[i] == d) ch.loss_gradient_output = (int)i;
                if (ch.loss_gradient_output < 0) { ok = false; break; }
            } else if (dz[l] != d) { ok = false; break; }
            if (ops_.nodes[gc.outputs[0]].shape.element_count() != K * N || gc.column_sum.size() > 1 || gc.outputs.size() != 1 + gc.column_sum.size()) { ok = false; break; }
            if (!gc.column_sum.empty()) {
                const Cluster& rc = gc.column_sum[0];
                ok = ok && rc.inputs.size() == 1 && rc.inputs[0].node_id == d && rc.inputs[0].chain.is_identity() &&
                     ops_.nodes[rc.node_id].op.reduce == ReduceOp::Sum && ops_.nodes[gc.outputs[1]].shape.element_count() == N;
            }
            // backward product dz_l W_l^T
            int b = -1;
            for (auto [dst, k] : cons[d]) {
                (void)k;
                const int ci = cluster_of(dst);
                if (ci < 0 || ci == g || !plain_matmul(clusters[ci]) || clusters[ci].matmul_absorbs_reduce || !clusters[ci].column_sum.empty()) continue;
                if (clusters[ci].inputs[0].node_id == d && is_plain(clusters[ci].inputs[0], M, N) && clusters[ci].inputs[1].node_id == w &&
                    is_transposed(clusters[ci].inputs[1], N, K) && clusters[ci].outputs.size() == 1)
                    b = ci;
            }
            ch.backward[l] = b;
            if (l > 0) {
                if (b < 0 || clusters[b].epilogue.empty()) { ok = false; break; }
                const Cluster& p = clusters[b].epilogue[0];
                ok = ok && p.outputs.size() == 1 && p.element_count == M * K;
                for (size_t i = 0; i < p.inputs.size(); ++i) {
                    if ((int)i == clusters[b].epilogue_product_input) continue;
                    if (p.inputs[i].node_id == act[l]) ok = ok && p.inputs[i].chain.is_identity();
                    else ok = ok && is_parameter(p.inputs[i].node_id);
                }
                for (const auto& op : p.ops) ok = ok && op.kind != PerElementOp::Gather && op.kind != PerElementOp::BuiltIn;
                dz[l - 1] = clusters[b].outputs[0];
            } else if (b >= 0 && !clusters[b].epilogue.empty()) {
                ok = false;
            }
        }
        if (!ok) continue;
        // sums of the loss cluster's other outputs: Reduce along the row, then over the batch
        for (size_t i = 0; i < lc.outputs.size() && ok; ++i) {
            if ((int)i == ch.loss_gradient_output) continue;
            auto single_reduce = [&](int node, int axis, int64_t count_out) {
                int r = -1, readers = 0;
                for (auto [dst, k] : cons[node]) {
                    if (!ops_.nodes[dst].alive) continue;
                    readers += 1;
                    const OpNode& d = ops_.nodes[dst];
                    const int ci = cluster_of(dst);
                    if (d.op.kind == OpKind::Reduce && d.op.reduce == ReduceOp::Sum && d.op.axis == axis && d.in[k].chain.is_identity() && ci >= 0 &&
                        clusters[ci].kind == ClusterKind::Reduce && clusters[ci].members.size() == 1 && d.shape.element_count() == count_out)
                        r = ci;
                }
                return readers == 1 ? r : -1;
            };
            const int rr = single_reduce(lc.outputs[i], 1, M);
            if (rr < 0 || ops_.nodes[lc.outputs[i]].shape.len() != 2) { ok = false; break; }
            const int br = single_reduce(clusters[rr].outputs[0], 0, 1);
            if (br < 0) { ok = false; break; }
            ch.sums.push_back({(int)i, rr, br});
        }
        if (!ok) continue;
        // every intermediate is read by the chain only, and no chain cluster belongs to two chains
        std::set<int> members;
        for (int c : ch.all_clusters()) members.insert(c);
        if ((int)members.size() != (int)ch.all_clusters().size()) continue;
        std::vector<int> internal;
        for (int l = 1; l <= L; ++l) if (act[l] >= 0) internal.push_back(act[l]);
        for (int l = 0; l < L; ++l) internal.push_back(dz[l]);
        for (const auto& s : ch.sums) { internal.push_back(lc.outputs[s.output]); internal.push_back(clusters[s.row_reduce].outputs[0]); }
        for (int node : internal)
            for (auto [dst, k] : cons[node]) {
                (void)k;
                if (!ops_.nodes[dst].alive) continue;
                if (ops_.nodes[dst].op.kind == OpKind::Output || !members.count(cluster_of(dst))) ok = false;
            }
        for (const auto& other : found)
            for (int c : other.all_clusters()) ok = ok && !members.count(c);
        if (ok) found.push_back(ch);
    }
    return found;
}

// Readers of a dense chain's results must come after the chain's last cluster, where the fused kernel runs.
void Graph::schedule_after_dense_chains(std::vector<Cluster>& clusters) {
    auto chains = detect_dense_chains(clusters);
    if (chains.empty()) return;
    auto cons = ops_.consumers();
    for (const auto& ch : chains) {
        std::set<int> members;
        int last_level = 0;
        for (int c : ch.all_clusters()) { members.insert(c); last_level = std::max(last_level, clusters[c].level); }
        for (int c : members)
            for (int out : clusters[c].outputs)
                for (auto [dst, k] : cons[out]) {
                    (void)k;
                    const int dc = ops_.nodes[dst].cluster_id;
                    if (ops_.nodes[dst].alive && dc >= 0 && !members.count(dc) && clusters[dc].level <= last_level) clusters[dc].level = last_level + 1;
                }
    }
    bool changed = true;
    while (changed) {
        changed = false;
        for (const auto& node : ops_.nodes) {
            if (!node.alive || node.cluster_id < 0) continue;
            Cluster& dc = clusters[node.cluster_id];
            for (const auto& e : node.in) {
                const OpNode& src = ops_.nodes[e.src];
                if (!src.alive || src.cluster_id < 0 || src.cluster_id == node.cluster_id) continue;
                if (dc.level <= clusters[src.cluster_id].level) { dc.level = clusters[src.cluster_id].level + 1; changed = true; }
            }
        }
    }
}

// See OperandPrologue (graph.hpp).  Candidates: ungrouped per-element clusters with one output X, a straight-line
// program without Gather / built-ins, where every reader of X is a MatMul cluster that takes X as its A or B operand
// (its absorbed bias column sums, which read X inside the same cluster, included).
void Graph::find_operand_prologues() {
    operand_prologues_.clear();
    auto cons = ops_.consumers();
    std::set<int> written_parameters;  // a parameter the step overwrites may change between the producer's slot and the GEMM's
    for (int id : output_nodes()) written_parameters.insert(ops_.nodes[id].op.parameter_id);
    for (int pi = 0; pi < (int)clusters_.size(); ++pi) {
        const Cluster& p = clusters_[pi];
        if (p.kind != ClusterKind::PerElement || !p.group.empty() || p.outputs.size() != 1 || p.element_count < (1 << 16)) continue;
        bool ok = true;
        for (const auto& op : p.ops)
            if (op.kind == PerElementOp::Gather || op.kind == PerElementOp::BuiltIn) ok = false;
        for (const auto& in : p.inputs) {
            const OpNode& src = ops_.nodes[in.node_id];
            if (src.op.kind == OpKind::Input && written_parameters.count(src.op.parameter_id)) ok = false;
        }
        const int x = p.outputs[0];
        OperandPrologue cand;
        cand.producer = pi;
        for (auto [dst, k] : cons[x]) {
            (void)k;
            const OpNode& d = ops_.nodes[dst];
            if (!d.alive) continue;
            if (d.op.kind == OpKind::Output || d.cluster_id < 0 || d.cluster_id == pi) { ok = false; break; }
            const Cluster& mc = clusters_[d.cluster_id];
            if (mc.kind == ClusterKind::WindowsToImage && mc.inputs[0].node_id == x) {  // the gather evaluates its windows itself (codegen gen_w2i)
                cand.uses.push_back({d.cluster_id, 0});
                continue;
            }
            if (mc.kind != ClusterKind::MatMul || !mc.epilogue.empty()) { ok = false; break; }
            bool as_operand = false;
            for (int operand = 0; operand < 2; ++operand) {
                if (mc.inputs[operand].node_id != x) continue;
                as_operand = true;
                bool seen = false;
                for (const auto& u : cand.uses) seen |= u.cluster == d.cluster_id && u.operand == operand;
                if (!seen) cand.uses.push_back({d.cluster_id, operand});
            }
            if (!as_operand) { ok = false; break; }  // read by the cluster in some other role only
        }
        // One consumer only.  With two (the second convolution's dY feeds both its weight-gradient and its backward-input
        // GEMM) the producer would be evaluated twice; that saves a third of the traffic on paper but measured slower on
        // B200 (conv-net m = 8192: 95 + 153 + 134 us as three kernels, 240 + 180 us fused): the tensor-core kernels stage
        // operands through registers and lose more to the three-fold loads in flight than the saved bytes return.
        std::set<int> consumer_clusters;
        for (const auto& u : cand.uses) consumer_clusters.insert(u.cluster);
        if (ok && consumer_clusters.size() == 1) operand_prologues_.push_back(cand);
    }
}

void Graph::write_dot_file(KernelDotOutput mode, const std::string& path) const {
    std::ofstream w(path);
    w << "digraph G {\n";
    auto emit_node = [&](int id) {
        const OpNode& n = ops_.nodes[id];
        w << "n" << id << " [shape=box,label=\"" << n.op.name() << "\\n" << n.shape.str() << "\"";
        if (mode == KernelDotOutput::Color) w << ",style=filled,fillcolor=\"/set312/" << (n.colour % 12) + 1 << "\"";
        w << "];\n";
    };
    std::vector<char> done(ops_.nodes.size(), 0);
    if (mode == KernelDotOutput::Cluster) {
        for (size_t ci = 0; ci < clusters_.size(); ++ci) {
            w << "subgraph cluster" << ci << " { style=filled; color=lightgrey; label=\"" << clusters_[ci].label << "\";\n";
            for (int id : clusters_[ci].members) { emit_node(id); done[id] = 1; }
            w << "}\n";
        }
    }
    for (int id = 0; id < (int)ops_.nodes.size(); ++id)
        if (ops_.nodes[id].alive && !done[id]) emit_node(id);
    for (int id = 0; id < (int)ops_.nodes.size(); ++id) {
        if (!ops_.nodes[id].alive) continue;
        for (const auto& e : ops_.nodes[id].in)
            w << "n" << e.src << " -> n" << id << " [label=\"" << (e.chain.is_identity() ? "" : "V") << "\"];\n";
    }
    w << "}\n";
}

}  // namespace descent

// Op IR of the descent expression graph.
//
// Restates the reference's `src/op.rs` (Op enum :83-97, UnaryOp/BinaryOp/CompareMode/ReduceOp/
// BuiltInOp/Literal :22-70, OpNode/OpEdge :162-174) and `src/parameter.rs`.  Two additions for the
// B200 backend: edges carry a `ViewChain` (shape.hpp) instead of a single View, and `AllReduce`
// marks where parameter gradients are summed across data-parallel ranks.
#pragma once
#include <cstring>
#include <memory>
#include <optional>
#include <string>
#include <vector>

#include "shape.hpp"

namespace descent {

enum class UnaryOp { Mov, Neg, Sqrt, Exp, Log, Sin, Cos, FloatToUint, UintToFloat };
enum class BinaryOp { Add, Sub, Mul, Div, Pow, UAdd, UMul, URem, UBitXor };
enum class CompareMode { Eq, Gt };
enum class ReduceOp { Max, Sum };
enum class BuiltInOp { Coord, Rand };
enum class MatMulOutputMode { Batches, Rows };

constexpr int MAX_OP_ARGS = 4;            // op.rs:72
constexpr int64_t MATMUL_MAX_K_SIZE = 1024;  // op.rs:74

enum class OpKind {
    Input, Output, Literal, BuiltIn, Unary, Binary, CompareAndSelect, MatMul, Reduce, Unpad, WindowsToImage, Gather,
    ScatterAdd, AllReduce
};

struct Op {
    OpKind kind = OpKind::Unary;
    int parameter_id = -1;           // Input / Output
    bool literal_is_u32 = false;     // Literal
    uint32_t literal_bits = 0;       // Literal: f32 bit pattern or the u32 value
    BuiltInOp built_in = BuiltInOp::Coord;
    int rand_uid = 0;                // BuiltIn::Rand
    UnaryOp unary = UnaryOp::Mov;
    BinaryOp binary = BinaryOp::Add;
    CompareMode compare = CompareMode::Eq;
    MatMulOutputMode output_mode = MatMulOutputMode::Batches;
    ReduceOp reduce = ReduceOp::Sum;
    int axis = 0;                    // Reduce / Unpad / Gather / ScatterAdd
    int64_t pad = 0;                 // Unpad
    int64_t stride_w = 1, stride_h = 1;  // WindowsToImage

    static Op input(int p) { Op o; o.kind = OpKind::Input; o.parameter_id = p; return o; }
    static Op output(int p) { Op o; o.kind = OpKind::Output; o.parameter_id = p; return o; }
    static Op literal_f32(float v) {
        Op o; o.kind = OpKind::Literal; std::memcpy(&o.literal_bits, &v, 4); return o;
    }
    static Op literal_u32(uint32_t v) { Op o; o.kind = OpKind::Literal; o.literal_is_u32 = true; o.literal_bits = v; return o; }
    static Op coord() { Op o; o.kind = OpKind::BuiltIn; o.built_in = BuiltInOp::Coord; return o; }
    static Op rand(int uid) { Op o; o.kind = OpKind::BuiltIn; o.built_in = BuiltInOp::Rand; o.rand_uid = uid; return o; }
    static Op un(UnaryOp u) { Op o; o.kind = OpKind::Unary; o.unary = u; return o; }
    static Op mov() { return un(UnaryOp::Mov); }
    static Op bin(BinaryOp b) { Op o; o.kind = OpKind::Binary; o.binary = b; return o; }
    static Op select(CompareMode m) { Op o; o.kind = OpKind::CompareAndSelect; o.compare = m; return o; }
    static Op matmul(MatMulOutputMode m) { Op o; o.kind = OpKind::MatMul; o.output_mode = m; return o; }
    static Op reduce_op(ReduceOp r, int axis) { Op o; o.kind = OpKind::Reduce; o.reduce = r; o.axis = axis; return o; }
    static Op unpad(int axis, int64_t pad) { Op o; o.kind = OpKind::Unpad; o.axis = axis; o.pad = pad; return o; }
    static Op windows_to_image(int64_t sw, int64_t sh) { Op o; o.kind = OpKind::WindowsToImage; o.stride_w = sw; o.stride_h = sh; return o; }
    static Op gather(int axis) { Op o; o.kind = OpKind::Gather; o.axis = axis; return o; }
    static Op scatter_add(int axis) { Op o; o.kind = OpKind::ScatterAdd; o.axis = axis; return o; }
    static Op all_reduce() { Op o; o.kind = OpKind::AllReduce; return o; }

    bool is_mov() const { return kind == OpKind::Unary && unary == UnaryOp::Mov; }
    bool is_literal_f32(float v) const {
        uint32_t b; std::memcpy(&b, &v, 4);
        return kind == OpKind::Literal && !literal_is_u32 && literal_bits == b;
    }
    bool is_literal_u32(uint32_t v) const { return kind == OpKind::Literal && literal_is_u32 && literal_bits == v; }
    float literal_f32_value() const { float f; std::memcpy(&f, &literal_bits, 4); return f; }
    // op.rs:114-134
    bool is_per_element() const {
        return kind == OpKind::Unary || kind == OpKind::Binary || kind == OpKind::CompareAndSelect || kind == OpKind::Gather;
    }
    bool is_gather_arg(int arg) const { return kind == OpKind::Gather && arg == 0; }
    bool is_inline_source() const { return kind == OpKind::Literal || kind == OpKind::BuiltIn; }
    bool can_merge() const { return kind != OpKind::Input && kind != OpKind::Output; }

    bool operator==(const Op& o) const {
        if (kind != o.kind) return false;
        switch (kind) {
            case OpKind::Input: case OpKind::Output: return parameter_id == o.parameter_id;
            case OpKind::Literal: return literal_is_u32 == o.literal_is_u32 && literal_bits == o.literal_bits;
            case OpKind::BuiltIn: return built_in == o.built_in && (built_in == BuiltInOp::Coord || rand_uid == o.rand_uid);
            case OpKind::Unary: return unary == o.unary;
            case OpKind::Binary: return binary == o.binary;
            case OpKind::CompareAndSelect: return compare == o.compare;
            case OpKind::MatMul: return output_mode == o.output_mode;
            case OpKind::Reduce: return reduce == o.reduce && axis == o.axis;
            case OpKind::Unpad: return axis == o.axis && pad == o.pad;
            case OpKind::WindowsToImage: return stride_w == o.stride_w && stride_h == o.stride_h;
            case OpKind::Gather: case OpKind::ScatterAdd: return axis == o.axis;
            case OpKind::AllReduce: return true;
        }
        return false;
    }
    std::string name() const;  // as the reference's Display impl (op.rs:136-160)
};

struct OpEdge {
    int src = -1;
    int arg = 0;
    ViewChain chain;  // producer elements -> the elements this argument reads
    Shape arg_shape;  // logical shape of the argument as the consumer sees it (the reference's view.output_shape)
    bool operator==(const OpEdge& o) const {
        return src == o.src && arg == o.arg && chain == o.chain && arg_shape == o.arg_shape;
    }
};

struct OpNode {
    int colour = 0;
    Shape shape;
    Op op;
    std::vector<OpEdge> in;  // at most one edge per arg
    bool alive = true;
    int cluster_id = -1;

    const OpEdge* arg_edge(int arg) const {
        for (const auto& e : in) if (e.arg == arg) return &e;
        return nullptr;
    }
    OpEdge* arg_edge(int arg) {
        for (auto& e : in) if (e.arg == arg) return &e;
        return nullptr;
    }
    int arg_count() const {
        int n = 0;
        for (const auto& e : in) n = std::max(n, e.arg + 1);
        return n;
    }
};

struct OpGraph {
    std::vector<OpNode> nodes;

    int new_node(int colour, const Shape& shape, const Op& op, const std::vector<int>& inputs) {
        OpNode n;
        n.colour = colour;
        n.shape = shape;
        n.op = op;
        for (size_t i = 0; i < inputs.size(); ++i) {
            OpEdge e;
            e.src = inputs[i];
            e.arg = (int)i;
            e.chain = ViewChain::identity(nodes[inputs[i]].shape.element_count());
            e.arg_shape = nodes[inputs[i]].shape;
            n.in.push_back(e);
        }
        nodes.push_back(std::move(n));
        return (int)nodes.size() - 1;
    }
    void add_edge(int src, int dst, int arg, const ViewChain& chain, const Shape& arg_shape) {
        DSC_CHECK(nodes[dst].arg_edge(arg) == nullptr, "argument " << arg << " already connected");
        DSC_CHECK(chain.input_count == nodes[src].shape.element_count() && chain.output_count == arg_shape.element_count(),
                  "edge chain does not match its endpoints");
        OpEdge e;
        e.src = src;
        e.arg = arg;
        e.chain = chain;
        e.arg_shape = arg_shape;
        nodes[dst].in.push_back(e);
    }
    void remove_node(int id) {
        nodes[id].alive = false;
        nodes[id].in.clear();
    }
    // consumers[n] = (dst node, index into dst.in)
    std::vector<std::vector<std::pair<int, int>>> consumers() const {
        std::vector<std::vector<std::pair<int, int>>> out(nodes.size());
        for (int d = 0; d < (int)nodes.size(); ++d) {
            if (!nodes[d].alive) continue;
            for (int k = 0; k < (int)nodes[d].in.size(); ++k) out[nodes[d].in[k].src].push_back({d, k});
        }
        return out;
    }
    // Kahn topological order over live nodes, ties broken by node id (deterministic)
    std::vector<int> topo_order() const;
};

// ---- parameters (parameter.rs) -------------------------------------------------------------

enum class InitKind { Zero, RandNormal, RandUniform };
struct Initializer {
    InitKind kind = InitKind::Zero;
    float scale = 0.f;
    static Initializer zero() { return {}; }
    static Initializer rand_normal(float s) { return {InitKind::RandNormal, s}; }
    static Initializer rand_uniform(float s) { return {InitKind::RandUniform, s}; }
    static Initializer for_relu(int64_t fan_in);                         // parameter.rs:17-20
    static Initializer for_siren(int64_t fan_in, bool is_first_layer);   // parameter.rs:22-25
};

struct ParameterStorage {
    Shape shape;
    std::string name;
    uint64_t buffer = 0;  // device buffer id (0 = none yet)
    std::optional<Initializer> reset_to;
};
using SharedParameters = std::shared_ptr<std::vector<ParameterStorage>>;

class Parameter {
public:
    Parameter() = default;
    Parameter(int id, SharedParameters owner) : id_(id), owner_(std::move(owner)) {}
    int checked_id(const SharedParameters& owner) const {
        DSC_CHECK(owner_ == owner, "parameter does not come from the same environment");
        return id_;
    }
    int id() const { return id_; }
    const Shape& shape() const { return (*owner_)[id_].shape; }
    const std::string& name() const { return (*owner_)[id_].name; }
    std::optional<Initializer> reset_to() const { return (*owner_)[id_].reset_to; }
    bool is_trainable() const { return (*owner_)[id_].reset_to.has_value(); }
    bool valid() const { return owner_ != nullptr; }

private:
    int id_ = -1;
    SharedParameters owner_;
};

}  // namespace descent

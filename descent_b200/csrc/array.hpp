// Array / UArray / DualArray / Scope: the graph-building API with hand-written reverse-mode
// gradients.  Names, argument meaning and lowering follow the reference's `src/array.rs`
// (Array ops :99-792, DualArray backward rules :794-1214, Scope :1216-1442) so that code written
// against the Rust API ports line by line.  Rust's `impl IntoArray` arguments become the small
// `ArrayArg` / `UArrayArg` / `DualArg` adaptor types.
#pragma once
#include <functional>
#include <map>
#include <set>
#include <utility>

#include "op.hpp"

namespace descent {

class Scope;
class Graph;
class Array;
class UArray;
class DualArray;

struct ArrayArg {
    enum Kind { kArray, kLiteral, kParameter } kind;
    int node_id = -1;
    const Scope* scope = nullptr;
    float value = 0.f;
    const Parameter* parameter = nullptr;
    ArrayArg(const Array& a);
    ArrayArg(float v) : kind(kLiteral), value(v) {}
    ArrayArg(double v) : kind(kLiteral), value((float)v) {}
    ArrayArg(int v) : kind(kLiteral), value((float)v) {}
    ArrayArg(const Parameter& p) : kind(kParameter), parameter(&p) {}
    Array into_array(Scope* scope) const;
};
struct UArrayArg {
    bool is_array;
    int node_id = -1;
    uint32_t value = 0;
    UArrayArg(const UArray& a);
    UArrayArg(uint32_t v) : is_array(false), value(v) {}
    UArrayArg(int v) : is_array(false), value((uint32_t)v) {}
    UArray into_array(Scope* scope) const;
};
struct DualArg {
    enum Kind { kDual, kLiteral, kParameter } kind;
    int value_node_id = -1, loss_grad_node_id = -1;
    float value = 0.f;
    const Parameter* parameter = nullptr;
    DualArg(const DualArray& a);
    DualArg(float v) : kind(kLiteral), value(v) {}
    DualArg(double v) : kind(kLiteral), value((float)v) {}
    DualArg(int v) : kind(kLiteral), value((float)v) {}
    DualArg(const Parameter& p) : kind(kParameter), parameter(&p) {}
    DualArray into_dual_array(Scope* scope) const;
};

// behaviour shared by Array and UArray (array.rs:99-225)
template <class Derived>
class ArrayCommon {
public:
    ArrayCommon() = default;
    ArrayCommon(int node_id, Scope* scope) : node_id_(node_id), scope_(scope) {}
    Scope* scope() const { return scope_; }
    int node_id() const { return node_id_; }
    Shape shape() const;
    Derived broadcast(const Shape& shape) const;
    Derived limit_axis(int axis, int64_t start, int64_t end) const;  // range [start, end)
    Derived lock_axis(int axis, int64_t coord, bool keep_axis) const;
    Derived reshape(const Shape& shape) const;
    Derived transpose() const;
    Derived view(const View& v) const;
    Derived unary_op(UnaryOp op) const;
    Derived keep_axis(int axis, bool keep) const;
    Derived remove_axis(int axis) const;

protected:
    Derived make(int node_id) const { return Derived(node_id, scope_); }
    int node_id_ = -1;
    Scope* scope_ = nullptr;
};

class UArray : public ArrayCommon<UArray> {
public:
    using ArrayCommon::ArrayCommon;
    Array to_f32_bits() const;
    Array into_f32() const;
    UArray binary_op(const UArrayArg& rhs, BinaryOp op) const;
};
UArray operator+(const UArray& a, const UArrayArg& b);
UArray operator*(const UArray& a, const UArrayArg& b);
UArray operator%(const UArray& a, const UArrayArg& b);
UArray operator^(const UArray& a, const UArrayArg& b);

class Array : public ArrayCommon<Array> {
public:
    using ArrayCommon::ArrayCommon;

    std::pair<Array, Array> with_empty_grad() const;
    Array concat(const ArrayArg& other, int axis) const;
    Array one_hot(int64_t count) const;
    Array reduce_max(int axis, bool keep_axis) const;
    Array reduce_sum(int axis, bool keep_axis) const;
    Array argmax(int axis, bool keep_axis) const;
    Array coord(int axis) const;
    Array gather(int axis, const UArrayArg& indices) const;
    Array scatter_add(const ArrayArg& values, int axis, const UArrayArg& indices) const;
    Array select_eq(const ArrayArg& rhs, const ArrayArg& pass, const ArrayArg& fail) const;
    Array select_gt(const ArrayArg& rhs, const ArrayArg& pass, const ArrayArg& fail) const;
    Array square() const;
    Array sqrt() const { return unary_op(UnaryOp::Sqrt); }
    Array exp() const { return unary_op(UnaryOp::Exp); }
    Array log() const { return unary_op(UnaryOp::Log); }
    Array sin() const { return unary_op(UnaryOp::Sin); }
    Array cos() const { return unary_op(UnaryOp::Cos); }
    UArray to_u32_bits() const { return UArray(node_id_, scope_); }
    UArray into_u32() const { return unary_op(UnaryOp::FloatToUint).to_u32_bits(); }
    Array sigmoid() const;
    Array tanh() const;
    Array pow(const ArrayArg& rhs) const { return binary_op(rhs, BinaryOp::Pow); }
    Array matmul(const ArrayArg& rhs) const;
    void accumulate(const ArrayArg& src) const;

    // crate-internal in the reference, used by DualArray and the modules
    Array binary_op(const ArrayArg& rhs, BinaryOp op) const;
    Array compare_and_select(CompareMode mode, const ArrayArg& rhs, const ArrayArg& pass, const ArrayArg& fail) const;
    Array reduce_op(ReduceOp op, int axis) const;
    Array unbroadcast(const Shape& shape) const;
    Array insert_axis(int axis) const;
    Array permute_axes(const std::vector<int>& perm) const;
    Array batched_matmul(const Array& rhs, MatMulOutputMode mode) const;
    Array pad(int axis, int64_t before, int64_t after) const;
    Array unpad(int axis, int64_t pad) const;
    Array pad_image(int64_t pad) const;
    Array unpad_image(int64_t pad) const;
    Array image_to_windows(int64_t filter_w, int64_t filter_h, int64_t stride_w, int64_t stride_h, int64_t groups) const;
    Array windows_to_image(int64_t stride_w, int64_t stride_h) const;
    void set_loss_grad_root() const;
    // data-parallel: sum this gradient accumulator over all ranks before anything else reads it
    void seal_with_all_reduce() const;
};
Array operator+(const Array& a, const ArrayArg& b);
Array operator-(const Array& a, const ArrayArg& b);
Array operator*(const Array& a, const ArrayArg& b);
Array operator/(const Array& a, const ArrayArg& b);
Array operator+(float a, const Array& b);
Array operator-(float a, const Array& b);
Array operator*(float a, const Array& b);
Array operator/(float a, const Array& b);
Array operator-(const Array& a);

class DualArray {
public:
    DualArray() = default;
    DualArray(const Array& value, const Array& loss_grad)
        : value_node_id_(value.node_id()), loss_grad_node_id_(loss_grad.node_id()), scope_(value.scope()) {}
    DualArray(const std::pair<Array, Array>& p) : DualArray(p.first, p.second) {}

    Array value() const { return Array(value_node_id_, scope_); }
    Array loss_grad() const { return Array(loss_grad_node_id_, scope_); }
    std::pair<Array, Array> into_inner() const { return {value(), loss_grad()}; }
    Shape shape() const { return value().shape(); }
    Scope* scope() const { return scope_; }

    DualArray square() const;
    DualArray sin() const;
    DualArray tanh() const;
    DualArray sigmoid() const;
    DualArray leaky_relu(float leakiness) const;
    DualArray matmul(const DualArg& rhs) const;
    DualArray transpose() const;
    DualArray pow(const DualArg& rhs) const;
    DualArray select_eq(const DualArg& rhs, const DualArg& pass, const DualArg& fail) const;
    DualArray lock_axis(int axis, int64_t coord, bool keep_axis) const;
    DualArray reshape(const Shape& shape) const;
    DualArray next_colour() const;
    DualArray map(const std::function<DualArray(DualArray)>& f) const { return f(*this); }
    DualArray conv2d(const DualArg& filter, int64_t pad, int64_t stride_w, int64_t stride_h) const;
    DualArray max_pool2d(int64_t filter_w, int64_t filter_h, int64_t stride_w, int64_t stride_h) const;
    DualArray reduce_sum(int axis, bool keep_axis) const;
    DualArray reduce_max(int axis, bool keep_axis) const;
    DualArray flatten() const;
    Array set_loss() const;
    DualArray concat(const DualArg& other, int axis) const;

    DualArray batched_matmul(const DualArray& rhs, MatMulOutputMode mode) const;
    DualArray pad_image(int64_t pad) const;
    DualArray image_to_windows(int64_t filter_w, int64_t filter_h, int64_t stride_w, int64_t stride_h, int64_t groups) const;
    DualArray permute_axes(const std::vector<int>& perm) const;
    DualArray insert_axis(int axis) const;
    DualArray remove_axis(int axis) const;
    DualArray keep_axis(int axis, bool keep) const;
    DualArray reduce_op(ReduceOp op, int axis) const;

private:
    int value_node_id_ = -1, loss_grad_node_id_ = -1;
    Scope* scope_ = nullptr;
};
DualArray operator+(const DualArray& a, const DualArg& b);
DualArray operator-(const DualArray& a, const DualArg& b);
DualArray operator*(const DualArray& a, const DualArg& b);

struct GraphInput {
    int value_node_id = -1;
    int grad_node_id = -1;  // -1 once the parameter has been overwritten in this scope
};

// data-parallel context a scope builds its graph for
struct DataParallel {
    int world = 1;
    int rank = 0;
};

class Scope {
public:
    Scope(SharedParameters parameters, DataParallel dp = {}) : parameters_(std::move(parameters)), dp_(dp) {}
    Scope(const Scope&) = delete;
    Scope& operator=(const Scope&) = delete;

    DualArray literal(float value);
    UArray literal_u32(uint32_t value);
    DualArray coord(int64_t len);
    DualArray rand(const Shape& shape);
    DualArray parameter(const Parameter& p);
    Array parameter_value(const Parameter& p);
    void write_parameter_value(const Parameter& p, const Array& rhs);
    Array update_parameter_value(const Parameter& p, const std::function<Array(Array)>& f);
    Array accumulator(const Shape& shape);
    void next_colour() { next_colour_ += 1; }
    std::vector<Parameter> trainable_parameters();
    Graph* build_graph();  // caller owns the result
    // Sum the loss gradients of `parameters` over the data-parallel ranks (one bucketed all-reduce);
    // idempotent per parameter and a no-op when world == 1.  Called by the optimisers and by
    // add_weight_decay_to_grad before they first read a gradient (optimizer.rs:8-12,36,91).
    void all_reduce_gradients(const std::vector<Parameter>& parameters);
    std::string export_json() const;  // raw (pre-pass) graph, for the oracle

    OpGraph& ops() { return ops_; }
    const OpGraph& ops() const { return ops_; }
    int colour() const { return next_colour_; }
    const DataParallel& dp() const { return dp_; }
    const SharedParameters& parameters() const { return parameters_; }

private:
    GraphInput input(const Parameter& p);
    OpGraph ops_;
    int next_colour_ = 0;
    int next_rand_uid_ = 0;
    SharedParameters parameters_;
    std::map<int, GraphInput> inputs_;
    std::map<int, int> outputs_;
    std::set<int> reduced_gradients_;
    DataParallel dp_;
};

}  // namespace descent

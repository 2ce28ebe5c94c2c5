// Graph-building ops and their reverse-mode gradient rules (reference: src/array.rs).
#include "array.hpp"

#include <cmath>
#include <cstdio>

#include "graph.hpp"

namespace descent {

// ---- argument adaptors ----------------------------------------------------------------------

ArrayArg::ArrayArg(const Array& a) : kind(kArray), node_id(a.node_id()), scope(a.scope()) {}
Array ArrayArg::into_array(Scope* s) const {
    switch (kind) {
        case kArray: return Array(node_id, s);
        case kLiteral: return s->literal(value).value();
        default: return s->parameter_value(*parameter);
    }
}
UArrayArg::UArrayArg(const UArray& a) : is_array(true), node_id(a.node_id()) {}
UArray UArrayArg::into_array(Scope* s) const { return is_array ? UArray(node_id, s) : s->literal_u32(value); }
DualArg::DualArg(const DualArray& a)
    : kind(kDual), value_node_id(a.value().node_id()), loss_grad_node_id(a.loss_grad().node_id()) {}
DualArray DualArg::into_dual_array(Scope* s) const {
    switch (kind) {
        case kDual: return DualArray(Array(value_node_id, s), Array(loss_grad_node_id, s));
        case kLiteral: return s->literal(value);
        default: return s->parameter(*parameter);
    }
}

// ---- ArrayCommon (array.rs:99-222) ------------------------------------------------------------

template <class D>
Shape ArrayCommon<D>::shape() const { return scope_->ops().nodes[node_id_].shape; }

template <class D>
D ArrayCommon<D>::view(const View& v) const {
    auto& ops = scope_->ops();
    DSC_CHECK(v.input_shape == ops.nodes[node_id_].shape, "view does not match array shape");
    int id = ops.new_node(scope_->colour(), v.output_shape, Op::mov(), {});
    ops.add_edge(node_id_, id, 0, ViewChain::of(v), v.output_shape);
    return make(id);
}
template <class D>
D ArrayCommon<D>::broadcast(const Shape& s) const { return view(View::broadcast(shape(), s)); }
template <class D>
D ArrayCommon<D>::unary_op(UnaryOp op) const {
    return make(scope_->ops().new_node(scope_->colour(), shape(), Op::un(op), {node_id_}));
}
template <class D>
D ArrayCommon<D>::remove_axis(int axis) const { return reshape(shape().remove_axis(axis)); }
template <class D>
D ArrayCommon<D>::keep_axis(int axis, bool keep) const { return keep ? make(node_id_) : remove_axis(shape().axis(axis)); }
template <class D>
D ArrayCommon<D>::limit_axis(int axis, int64_t start, int64_t end) const {
    Shape s = shape();
    return view(View::limited(s, s.axis(axis), start, end));
}
template <class D>
D ArrayCommon<D>::lock_axis(int axis, int64_t coord, bool keep) const {
    int a = shape().axis(axis);
    return limit_axis(a, coord, coord + 1).keep_axis(a, keep);
}
template <class D>
D ArrayCommon<D>::reshape(const Shape& s) const {
    DSC_CHECK(shape().element_count() == s.element_count(), "reshape " << shape().str() << " -> " << s.str());
    return make(scope_->ops().new_node(scope_->colour(), s, Op::mov(), {node_id_}));
}
template <class D>
D ArrayCommon<D>::transpose() const { return view(View::identity(shape()).transposed()); }

template class ArrayCommon<Array>;
template class ArrayCommon<UArray>;

// ---- UArray ----------------------------------------------------------------------------------

Array UArray::to_f32_bits() const { return Array(node_id_, scope_); }
Array UArray::into_f32() const { return unary_op(UnaryOp::UintToFloat).to_f32_bits(); }
UArray UArray::binary_op(const UArrayArg& rhs_arg, BinaryOp op) const {
    UArray rhs = rhs_arg.into_array(scope_);
    Shape op_shape = shape().broadcast_with(rhs.shape());
    int l = broadcast(op_shape).node_id(), r = rhs.broadcast(op_shape).node_id();
    return UArray(scope_->ops().new_node(scope_->colour(), op_shape, Op::bin(op), {l, r}), scope_);
}
UArray operator+(const UArray& a, const UArrayArg& b) { return a.binary_op(b, BinaryOp::UAdd); }
UArray operator*(const UArray& a, const UArrayArg& b) { return a.binary_op(b, BinaryOp::UMul); }
UArray operator%(const UArray& a, const UArrayArg& b) { return a.binary_op(b, BinaryOp::URem); }
UArray operator^(const UArray& a, const UArrayArg& b) { return a.binary_op(b, BinaryOp::UBitXor); }

// ---- Array -----------------------------------------------------------------------------------

std::pair<Array, Array> Array::with_empty_grad() const {
    int g = scope_->ops().new_node(scope_->colour(), shape(), Op::mov(), {});
    return {*this, Array(g, scope_)};
}

Array Array::binary_op(const ArrayArg& rhs_arg, BinaryOp op) const {
    Array rhs = rhs_arg.into_array(scope_);
    Shape op_shape = shape().broadcast_with(rhs.shape());
    int l = broadcast(op_shape).node_id(), r = rhs.broadcast(op_shape).node_id();
    return Array(scope_->ops().new_node(scope_->colour(), op_shape, Op::bin(op), {l, r}), scope_);
}
Array operator+(const Array& a, const ArrayArg& b) { return a.binary_op(b, BinaryOp::Add); }
Array operator-(const Array& a, const ArrayArg& b) { return a.binary_op(b, BinaryOp::Sub); }
Array operator*(const Array& a, const ArrayArg& b) { return a.binary_op(b, BinaryOp::Mul); }
Array operator/(const Array& a, const ArrayArg& b) { return a.binary_op(b, BinaryOp::Div); }
Array operator+(float a, const Array& b) { return b.scope()->literal(a).value().binary_op(b, BinaryOp::Add); }
Array operator-(float a, const Array& b) { return b.scope()->literal(a).value().binary_op(b, BinaryOp::Sub); }
Array operator*(float a, const Array& b) { return b.scope()->literal(a).value().binary_op(b, BinaryOp::Mul); }
Array operator/(float a, const Array& b) { return b.scope()->literal(a).value().binary_op(b, BinaryOp::Div); }
Array operator-(const Array& a) { return a.unary_op(UnaryOp::Neg); }

// sum away axes that were broadcast: leading axes first, then unit axes (array.rs:241-262)
Array Array::unbroadcast(const Shape& target) const {
    Array out = *this;
    while (out.shape().len() > target.len()) out = out.reduce_sum(0, false);
    DSC_CHECK(out.shape().len() == target.len(), "unbroadcast rank mismatch");
    for (int i = 0; i < target.len(); ++i) {
        if (out.shape()[i] != target[i]) {
            DSC_CHECK(target[i] == 1, "unbroadcast to non-unit axis");
            out = out.reduce_sum(i, true);
        }
    }
    return out;
}

Array Array::compare_and_select(CompareMode mode, const ArrayArg& rhs_arg, const ArrayArg& pass_arg, const ArrayArg& fail_arg) const {
    Array rhs = rhs_arg.into_array(scope_), pass = pass_arg.into_array(scope_), fail = fail_arg.into_array(scope_);
    Shape op_shape = shape().broadcast_with(rhs.shape()).broadcast_with(pass.shape()).broadcast_with(fail.shape());
    int a = broadcast(op_shape).node_id(), b = rhs.broadcast(op_shape).node_id();
    int p = pass.broadcast(op_shape).node_id(), f = fail.broadcast(op_shape).node_id();
    return Array(scope_->ops().new_node(scope_->colour(), op_shape, Op::select(mode), {a, b, p, f}), scope_);
}
Array Array::select_eq(const ArrayArg& r, const ArrayArg& p, const ArrayArg& f) const { return compare_and_select(CompareMode::Eq, r, p, f); }
Array Array::select_gt(const ArrayArg& r, const ArrayArg& p, const ArrayArg& f) const { return compare_and_select(CompareMode::Gt, r, p, f); }

// select between the two operands, each replicate-padded to the joined length (array.rs:299-324)
Array Array::concat(const ArrayArg& other_arg, int axis_in) const {
    Array other = other_arg.into_array(scope_);
    Shape s = shape(), os = other.shape();
    int axis = s.axis(axis_in);
    int64_t length = s[axis], other_length = os[axis], total = length + other_length;
    Shape out_shape = s.resize_axis(axis, total);
    DSC_CHECK(out_shape == os.resize_axis(axis, total), "concat shapes differ off-axis");
    Array out_coord = scope_->coord(total).value().reshape(out_shape.coord(axis));
    return out_coord.compare_and_select(CompareMode::Gt, (float)(length - 1), other.pad(axis, length, 0), pad(axis, 0, other_length));
}

Array Array::reduce_op(ReduceOp op, int axis_in) const {
    Shape s = shape();
    int axis = s.axis(axis_in);
    if (s[axis] == 1) return *this;  // array.rs:329-331
    return Array(scope_->ops().new_node(scope_->colour(), s.reduce(axis), Op::reduce_op(op, axis), {node_id_}), scope_);
}
Array Array::reduce_max(int axis_in, bool keep) const {
    int axis = shape().axis(axis_in);
    return reduce_op(ReduceOp::Max, axis).keep_axis(axis, keep);
}
Array Array::reduce_sum(int axis_in, bool keep) const {
    int axis = shape().axis(axis_in);
    return reduce_op(ReduceOp::Sum, axis).keep_axis(axis, keep);
}
Array Array::one_hot(int64_t count) const { return scope_->coord(count).value().select_eq(*this, 1.0f, 0.0f); }
Array Array::argmax(int axis_in, bool keep) const {  // largest index among ties (array.rs:362-367)
    int axis = shape().axis(axis_in);
    Array coord_or_zero = select_eq(reduce_max(axis, true), coord(axis), 0.0f);
    return coord_or_zero.reduce_max(axis, keep);
}
Array Array::coord(int axis_in) const {
    Shape s = shape();
    int axis = s.axis(axis_in);
    return scope_->coord(s[axis]).value().reshape(s.coord(axis));
}

Array Array::gather(int axis_in, const UArrayArg& indices_arg) const {
    UArray indices = indices_arg.into_array(scope_);
    DSC_CHECK(indices.shape().len() == 1, "gather indices must be 1-D");
    int64_t index_count = indices.shape()[0];
    Shape values_shape = shape();
    int axis = values_shape.axis(axis_in);
    Shape s = values_shape.resize_axis(axis, index_count);
    UArray index = indices.reshape(s.coord(axis)).broadcast(s);
    return Array(scope_->ops().new_node(scope_->colour(), s, Op::gather(axis), {node_id_, index.node_id()}), scope_);
}
Array Array::scatter_add(const ArrayArg& values_arg, int axis_in, const UArrayArg& indices_arg) const {
    Shape s = shape();
    Array values = values_arg.into_array(scope_);
    int axis = s.axis(axis_in);
    UArray indices = indices_arg.into_array(scope_);
    DSC_CHECK(indices.shape().len() == 1, "scatter_add indices must be 1-D");
    DSC_CHECK(s.resize_axis(axis, indices.shape()[0]) == values.shape(), "scatter_add values shape mismatch");
    return Array(scope_->ops().new_node(scope_->colour(), s, Op::scatter_add(axis), {node_id_, values.node_id(), indices.node_id()}), scope_);
}

Array Array::square() const { return *this * *this; }
Array Array::sigmoid() const { return exp() / (exp() + 1.0f); }
Array Array::tanh() const {
    Array a = exp(), b = (-*this).exp();
    return (a - b) / (a + b);
}
Array Array::insert_axis(int axis) const { return reshape(shape().insert_axis(axis, 1)); }
Array Array::permute_axes(const std::vector<int>& perm) const { return view(View::identity(shape()).permute_axes(perm)); }

Array Array::matmul(const ArrayArg& rhs_arg) const {
    Array lhs = insert_axis(0), rhs = rhs_arg.into_array(scope_).insert_axis(0);
    return lhs.batched_matmul(rhs, MatMulOutputMode::Batches).remove_axis(0);
}
// [b,m,k] x [b,k,n]; K is split into r = ceil(K/1024) chunks summed by a Reduce (array.rs:500-520)
Array Array::batched_matmul(const Array& rhs, MatMulOutputMode mode) const {
    Shape a = shape(), b = rhs.shape();
    DSC_CHECK(a.len() == 3 && b.len() == 3 && a[0] == b[0] && a[2] == b[1], "batched_matmul " << a.str() << " x " << b.str());
    int64_t r = div_round_up(a[2], MATMUL_MAX_K_SIZE);
    Shape s = mode == MatMulOutputMode::Batches ? Shape{r, a[0], a[1], b[2]} : Shape{r, a[1], a[0], b[2]};
    Array chunks(scope_->ops().new_node(scope_->colour(), s, Op::matmul(mode), {node_id_, rhs.node_id()}), scope_);
    Array output = chunks.reduce_sum(0, false);
    return mode == MatMulOutputMode::Batches ? output : output.permute_axes({1, 0, 2});
}

Array Array::pad(int axis_in, int64_t before, int64_t after) const {
    if (before + after == 0) return *this;
    Shape s = shape();
    return view(View::padded(s, s.axis(axis_in), before, after));
}
Array Array::unpad(int axis_in, int64_t pad) const {
    if (pad == 0) return *this;
    Shape s = shape();
    int axis = s.axis(axis_in);
    return Array(scope_->ops().new_node(scope_->colour(), s.unpad(axis, pad), Op::unpad(axis, pad), {node_id_}), scope_);
}
Array Array::pad_image(int64_t p) const { return pad(-3, p, p).pad(-2, p, p); }
Array Array::unpad_image(int64_t p) const { return unpad(-3, p).unpad(-2, p); }

// 7-D window view of an NHWC image: [.., oh, ow, g, fh, fw, c/g]  (array.rs:559-600)
Array Array::image_to_windows(int64_t filter_w, int64_t filter_h, int64_t stride_w, int64_t stride_h, int64_t groups) const {
    Shape in = shape();
    int y = in.axis(-3), x = in.axis(-2), c = in.axis(-1);
    View v = View::identity(in);
    v.output_shape = in.image_to_windows(filter_w, filter_h, stride_w, stride_h, groups);
    int64_t group_nc = v.output_shape.at(-1);
    v.output_mapping.resize(v.output_shape.len() - 6);
    v.output_mapping.push_back(AxisMapping::identity(y, in[y]).stepped(stride_h));
    v.output_mapping.push_back(AxisMapping::identity(x, in[x]).stepped(stride_w));
    v.output_mapping.push_back(AxisMapping::identity(c, in[c]).stepped(group_nc));
    v.output_mapping.push_back(AxisMapping::identity(y, in[y]));
    v.output_mapping.push_back(AxisMapping::identity(x, in[x]));
    v.output_mapping.push_back(AxisMapping::identity(c, in[c]));
    return view(v);
}
Array Array::windows_to_image(int64_t stride_w, int64_t stride_h) const {
    Shape s = shape().windows_to_image(stride_w, stride_h);
    {
        // Non-overlapping windows (stride == filter, e.g. the 2x2/2 max-pool backward): every image element
        // comes from exactly one window element, so col2im is a pure permutation and can be a view
        // ([.., oh, ow, g, fh, fw, c] -> [.., oh, fh, ow, fw, g, c] -> reshape) instead of a kernel.
        const Shape w = shape();
        const int n = w.len();
        if (w[n - 3] == stride_h && w[n - 2] == stride_w) {
            std::vector<int> perm;
            for (int i = 0; i < n - 6; ++i) perm.push_back(i);
            for (int i : {n - 6, n - 3, n - 5, n - 2, n - 4, n - 1}) perm.push_back(i);
            return permute_axes(perm).reshape(s);
        }
    }
    return Array(scope_->ops().new_node(scope_->colour(), s, Op::windows_to_image(stride_w, stride_h), {node_id_}), scope_);
}

// `self` is a gradient accumulator (an input-less or single-input Mov); add `src` to it (array.rs:617-650)
void Array::accumulate(const ArrayArg& src_arg) const {
    Array src = src_arg.into_array(scope_);
    auto& ops = scope_->ops();
    DSC_CHECK(ops.nodes[node_id_].op.is_mov(), "accumulate target must be a gradient accumulator");
    DSC_CHECK(ops.nodes[node_id_].shape == ops.nodes[src.node_id()].shape,
              "accumulate shape mismatch " << ops.nodes[node_id_].shape.str() << " += " << ops.nodes[src.node_id()].shape.str());
    int src_id = src.node_id();
    if (!ops.nodes[node_id_].in.empty()) {
        int prev = ops.nodes[node_id_].in[0].src;
        ops.nodes[node_id_].in.clear();
        src_id = ops.new_node(scope_->colour(), ops.nodes[src.node_id()].shape, Op::bin(BinaryOp::Add), {prev, src.node_id()});
    }
    const Shape& s = ops.nodes[src_id].shape;
    ops.add_edge(src_id, node_id_, 0, ViewChain::identity(s.element_count()), s);
}

// dL/dloss = 1/m over the (global) mini-batch (array.rs:652-672; SURVEY.md §8e condition 1)
void Array::set_loss_grad_root() const {
    Shape grad_shape = shape();
    int64_t mini_batch_size = grad_shape[0] * scope_->dp().world;
    Array scale = scope_->literal(1.0f / (float)mini_batch_size).value().broadcast(grad_shape);
    auto& ops = scope_->ops();
    DSC_CHECK(ops.nodes[node_id_].op.is_mov() && ops.nodes[node_id_].in.empty(), "loss gradient already has a source");
    ops.add_edge(scale.node_id(), node_id_, 0, ViewChain::identity(grad_shape.element_count()), grad_shape);
}

void Array::seal_with_all_reduce() const {
    auto& ops = scope_->ops();
    DSC_CHECK(ops.nodes[node_id_].op.is_mov() && !ops.nodes[node_id_].in.empty(), "all-reduce of a gradient nothing accumulated into");
    int prev = ops.nodes[node_id_].in[0].src;
    ops.nodes[node_id_].in.clear();
    const Shape s = ops.nodes[prev].shape;
    int ar = ops.new_node(scope_->colour(), s, Op::all_reduce(), {prev});
    ops.add_edge(ar, node_id_, 0, ViewChain::identity(s.element_count()), s);
}

// ---- DualArray (array.rs:794-1214) -------------------------------------------------------------

DualArray DualArray::square() const { return *this * *this; }

DualArray DualArray::sin() const {
    auto [a, da] = into_inner();
    auto [b, db] = a.sin().with_empty_grad();
    da.accumulate(db * a.cos());
    return {b, db};
}
DualArray DualArray::tanh() const {
    auto [a, da] = into_inner();
    auto [b, db] = a.tanh().with_empty_grad();
    da.accumulate(db * 4.0f / ((2.0f * a).exp() + 2.0f + (-2.0f * a).exp()));
    return {b, db};
}
DualArray DualArray::sigmoid() const {
    auto [a, da] = into_inner();
    auto [b, db] = a.sigmoid().with_empty_grad();
    da.accumulate(db * a.exp() / (a.exp() + 1.0f).square());
    return {b, db};
}
DualArray DualArray::leaky_relu(float leakiness) const {
    auto [a, da] = into_inner();
    auto [b, db] = a.select_gt(0.0f, a, a * leakiness).with_empty_grad();
    da.accumulate(a.select_gt(0.0f, db, db * leakiness));
    return {b, db};
}
DualArray DualArray::batched_matmul(const DualArray& rhs, MatMulOutputMode mode) const {
    auto [a, da] = into_inner();
    auto [b, db] = rhs.into_inner();
    auto [c, dc] = a.batched_matmul(b, mode).with_empty_grad();
    da.accumulate(dc.batched_matmul(b.transpose(), MatMulOutputMode::Batches));
    db.accumulate(a.transpose().batched_matmul(dc, MatMulOutputMode::Batches));
    return {c, dc};
}
DualArray DualArray::matmul(const DualArg& rhs_arg) const {
    DualArray lhs = insert_axis(0), rhs = rhs_arg.into_dual_array(scope_).insert_axis(0);
    return lhs.batched_matmul(rhs, MatMulOutputMode::Batches).remove_axis(0);
}
DualArray DualArray::transpose() const {
    auto [a, da] = into_inner();
    auto [b, db] = a.transpose().with_empty_grad();
    da.accumulate(db.transpose());
    return {b, db};
}
DualArray DualArray::pow(const DualArg& rhs_arg) const {
    auto [a, da] = into_inner();
    auto [b, db] = rhs_arg.into_dual_array(scope_).into_inner();
    auto [c, dc] = a.pow(b).with_empty_grad();
    da.accumulate((dc * b * a.pow(b - 1.0f)).unbroadcast(a.shape()));
    db.accumulate((dc * a.log() * c).unbroadcast(b.shape()));
    return {c, dc};
}
DualArray DualArray::select_eq(const DualArg& rhs_arg, const DualArg& pass_arg, const DualArg& fail_arg) const {
    Array a = value();
    Array b = rhs_arg.into_dual_array(scope_).value();
    auto [pass, dpass] = pass_arg.into_dual_array(scope_).into_inner();
    auto [fail, dfail] = fail_arg.into_dual_array(scope_).into_inner();
    auto [c, dc] = a.select_eq(b, pass, fail).with_empty_grad();
    dpass.accumulate(a.select_eq(b, dc, 0.0f).unbroadcast(pass.shape()));
    dfail.accumulate(a.select_eq(b, 0.0f, dc).unbroadcast(fail.shape()));
    return {c, dc};
}
DualArray DualArray::lock_axis(int axis_in, int64_t coord, bool keep) const {
    int axis = shape().axis(axis_in);
    auto [a, da] = into_inner();
    auto [b, db] = a.lock_axis(axis, coord, true).with_empty_grad();
    da.accumulate(a.coord(axis).select_eq((float)coord, db, 0.0f));
    return DualArray(b, db).keep_axis(axis, keep);
}
DualArray DualArray::reshape(const Shape& new_shape) const {
    Shape old_shape = shape();
    auto [a, da] = into_inner();
    auto [b, db] = a.reshape(new_shape).with_empty_grad();
    da.accumulate(db.reshape(old_shape));
    return {b, db};
}
DualArray DualArray::pad_image(int64_t pad) const {
    auto [a, da] = into_inner();
    auto [b, db] = a.pad_image(pad).with_empty_grad();
    da.accumulate(db.unpad_image(pad));
    return {b, db};
}
DualArray DualArray::image_to_windows(int64_t fw, int64_t fh, int64_t sw, int64_t sh, int64_t groups) const {
    auto [a, da] = into_inner();
    auto [b, db] = a.image_to_windows(fw, fh, sw, sh, groups).with_empty_grad();
    da.accumulate(db.windows_to_image(sw, sh));
    return {b, db};
}
DualArray DualArray::next_colour() const {
    scope_->next_colour();
    return *this;
}

// NHWC convolution as pad -> windows -> [M,g,K] -> batched matmul in Rows mode (array.rs:989-1031).
// filter: [g, oc/g, fh, fw, ic/g]; output channel = g_idx*(oc/g) + oc_idx.
DualArray DualArray::conv2d(const DualArg& filter_arg, int64_t pad, int64_t stride_w, int64_t stride_h) const {
    DualArray filter = filter_arg.into_dual_array(scope_);
    DualArray padded = pad_image(pad);
    Shape ps = padded.shape(), fs = filter.shape();
    DSC_CHECK(ps.len() == 4 && fs.len() == 5, "conv2d expects NHWC input and [g,oc,fh,fw,ic] filter");
    int64_t input_m = ps[0], input_nc = ps[3];
    int64_t filter_g = fs[0], filter_oc = fs[1], filter_h = fs[2], filter_w = fs[3], filter_ic = fs[4];
    DSC_CHECK(input_nc == filter_g * filter_ic, "conv2d channel mismatch");
    DualArray windows = padded.image_to_windows(filter_w, filter_h, stride_w, stride_h, filter_g);
    Shape ws = windows.shape();
    int64_t output_h = ws[1], output_w = ws[2];
    DualArray a = windows.reshape({input_m * output_h * output_w, filter_g, filter_h * filter_w * filter_ic}).permute_axes({1, 0, 2});
    DualArray b = filter.reshape({filter_g, filter_oc, filter_h * filter_w * filter_ic});
    DualArray c = a.batched_matmul(b.transpose(), MatMulOutputMode::Rows);
    return c.permute_axes({1, 0, 2}).reshape({input_m, output_h, output_w, filter_g * filter_oc});
}
DualArray DualArray::max_pool2d(int64_t filter_w, int64_t filter_h, int64_t stride_w, int64_t stride_h) const {
    DualArray windows = image_to_windows(filter_w, filter_h, stride_w, stride_h, 1);
    Shape ws = windows.shape();
    int64_t m = ws[0], oh = ws[1], ow = ws[2], groups = ws[3], fh = ws[4], fw = ws[5], gnc = ws[6];
    return windows.reshape({m * oh * ow * groups, fh * fw, gnc}).reduce_max(1, true).reshape({m, oh, ow, groups * gnc});
}
DualArray DualArray::reduce_op(ReduceOp op, int axis) const {
    auto [a, da] = into_inner();
    auto [b, db] = a.reduce_op(op, axis).with_empty_grad();
    if (op == ReduceOp::Max) da.accumulate(a.select_eq(b, db, 0.0f));  // every tied maximum gets the gradient
    else da.accumulate(db.broadcast(da.shape()));
    return {b, db};
}
DualArray DualArray::insert_axis(int axis) const {
    auto [a, da] = into_inner();
    auto [b, db] = a.insert_axis(axis).with_empty_grad();
    da.accumulate(db.remove_axis(axis));
    return {b, db};
}
DualArray DualArray::remove_axis(int axis) const {
    auto [a, da] = into_inner();
    auto [b, db] = a.remove_axis(axis).with_empty_grad();
    da.accumulate(db.insert_axis(axis));
    return {b, db};
}
DualArray DualArray::keep_axis(int axis, bool keep) const { return keep ? *this : remove_axis(axis); }
DualArray DualArray::reduce_sum(int axis_in, bool keep) const {
    int axis = shape().axis(axis_in);
    return reduce_op(ReduceOp::Sum, axis).keep_axis(axis, keep);
}
DualArray DualArray::reduce_max(int axis_in, bool keep) const {
    int axis = shape().axis(axis_in);
    return reduce_op(ReduceOp::Max, axis).keep_axis(axis, keep);
}
DualArray DualArray::flatten() const {
    Shape s = shape();
    return reshape({s[0], s.element_count() / s[0]});
}
Array DualArray::set_loss() const {
    loss_grad().set_loss_grad_root();
    return value();
}
DualArray DualArray::permute_axes(const std::vector<int>& perm) const {
    std::vector<int> inv(perm.size());
    for (size_t src = 0; src < perm.size(); ++src) inv[perm[src]] = (int)src;
    auto [a, da] = into_inner();
    auto [b, db] = a.permute_axes(perm).with_empty_grad();
    da.accumulate(db.permute_axes(inv));
    return {b, db};
}
DualArray DualArray::concat(const DualArg& other_arg, int axis_in) const {
    DualArray other = other_arg.into_dual_array(scope_);
    Shape s = shape();
    int axis = s.axis(axis_in);
    int64_t length = s[axis];
    auto [a, da] = into_inner();
    auto [b, db] = other.into_inner();
    auto [c, dc] = a.concat(b, axis).with_empty_grad();
    da.accumulate(dc.limit_axis(axis, 0, length));
    db.accumulate(dc.limit_axis(axis, length, dc.shape()[axis]));
    return {c, dc};
}
DualArray operator+(const DualArray& lhs, const DualArg& rhs_arg) {
    auto [a, da] = lhs.into_inner();
    auto [b, db] = rhs_arg.into_dual_array(lhs.scope()).into_inner();
    auto [c, dc] = (a + b).with_empty_grad();
    da.accumulate(dc.unbroadcast(a.shape()));
    db.accumulate(dc.unbroadcast(b.shape()));
    return {c, dc};
}
DualArray operator-(const DualArray& lhs, const DualArg& rhs_arg) {
    auto [a, da] = lhs.into_inner();
    auto [b, db] = rhs_arg.into_dual_array(lhs.scope()).into_inner();
    auto [c, dc] = (a - b).with_empty_grad();
    da.accumulate(dc.unbroadcast(a.shape()));
    db.accumulate(-dc.unbroadcast(b.shape()));
    return {c, dc};
}
DualArray operator*(const DualArray& lhs, const DualArg& rhs_arg) {
    auto [a, da] = lhs.into_inner();
    auto [b, db] = rhs_arg.into_dual_array(lhs.scope()).into_inner();
    auto [c, dc] = (a * b).with_empty_grad();
    da.accumulate((b * dc).unbroadcast(a.shape()));
    db.accumulate((a * dc).unbroadcast(b.shape()));
    return {c, dc};
}

// ---- Scope (array.rs:1231-1442) ----------------------------------------------------------------

DualArray Scope::literal(float value) {
    DSC_CHECK(!std::isnan(value), "literal must not be NaN");
    return Array(ops_.new_node(next_colour_, Shape{1}, Op::literal_f32(value), {}), this).with_empty_grad();
}
UArray Scope::literal_u32(uint32_t value) { return UArray(ops_.new_node(next_colour_, Shape{1}, Op::literal_u32(value), {}), this); }
DualArray Scope::coord(int64_t len) { return Array(ops_.new_node(next_colour_, Shape{len}, Op::coord(), {}), this).with_empty_grad(); }
DualArray Scope::rand(const Shape& shape) {
    int uid = next_rand_uid_++;
    return Array(ops_.new_node(next_colour_, shape, Op::rand(uid), {}), this).with_empty_grad();
}
GraphInput Scope::input(const Parameter& p) {
    int pid = p.checked_id(parameters_);
    auto it = inputs_.find(pid);
    if (it != inputs_.end()) return it->second;
    const Shape& shape = (*parameters_)[pid].shape;
    GraphInput gi;
    gi.value_node_id = ops_.new_node(next_colour_, shape, Op::input(pid), {});
    gi.grad_node_id = ops_.new_node(next_colour_, shape, Op::mov(), {});
    inputs_[pid] = gi;
    return gi;
}
DualArray Scope::parameter(const Parameter& p) {
    GraphInput gi = input(p);
    DSC_CHECK(gi.grad_node_id >= 0, "parameter '" << p.name() << "' was overwritten in this scope and has no gradient");
    return DualArray(Array(gi.value_node_id, this), Array(gi.grad_node_id, this));
}
Array Scope::parameter_value(const Parameter& p) { return Array(input(p).value_node_id, this); }
void Scope::write_parameter_value(const Parameter& p, const Array& rhs) {
    int pid = p.checked_id(parameters_);
    const Shape& shape = ops_.nodes[rhs.node_id()].shape;
    DSC_CHECK((*parameters_)[pid].shape == shape, "write_parameter_value shape mismatch for '" << p.name() << "'");
    int node_id = ops_.new_node(next_colour_, shape, Op::output(pid), {rhs.node_id()});
    auto it = outputs_.find(pid);
    if (it != outputs_.end()) ops_.remove_node(it->second);
    outputs_[pid] = node_id;
    // reading the parameter again in this scope sees the value just written (array.rs:1383-1391)
    inputs_[pid] = GraphInput{rhs.node_id(), -1};
}
Array Scope::update_parameter_value(const Parameter& p, const std::function<Array(Array)>& f) {
    Array result = f(parameter_value(p));
    write_parameter_value(p, result);
    return result;
}
Array Scope::accumulator(const Shape& shape) { return Array(ops_.new_node(next_colour_, shape, Op::mov(), {}), this); }
std::vector<Parameter> Scope::trainable_parameters() {
    std::vector<Parameter> v;
    for (const auto& n : ops_.nodes)
        if (n.alive && n.op.kind == OpKind::Input && (*parameters_)[n.op.parameter_id].reset_to.has_value())
            v.emplace_back(n.op.parameter_id, parameters_);
    return v;
}
void Scope::all_reduce_gradients(const std::vector<Parameter>& parameters) {
    if (dp_.world <= 1) return;
    for (const auto& p : parameters) {
        int pid = p.checked_id(parameters_);
        if (!reduced_gradients_.insert(pid).second) continue;
        parameter(p).loss_grad().seal_with_all_reduce();
    }
}
Graph* Scope::build_graph() { return new Graph(parameters_, ops_, dp_); }

}  // namespace descent

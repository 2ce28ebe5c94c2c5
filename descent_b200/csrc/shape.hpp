// Shapes, axes and views: the index algebra of the descent array API.
//
// Restates the semantics of the reference's `src/shape.rs` (Shape :41-237, AxisMapping :321-352,
// View :354-636) in C++, and extends it with `ViewChain`: a sequence of views joined by
// linear-index-preserving reshapes.  The reference materialises a copy whenever a reshape cannot be
// folded into a single View (e.g. the im2col reshape after `image_to_windows`, graph.rs:262-284);
// here every view/reshape folds into a chain that the CUDA kernels evaluate as index arithmetic,
// so no view ever costs a copy.
#pragma once
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace descent {

constexpr int MAX_DIM = 7;  // shape.rs:10

[[noreturn]] inline void fail(const std::string& msg) { throw std::runtime_error(msg); }
#define DSC_CHECK(cond, msg)                                                            \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            std::ostringstream os_;                                                     \
            os_ << "descent: " << msg << " [" #cond "] at " << __FILE__ << ":" << __LINE__; \
            ::descent::fail(os_.str());                                                 \
        }                                                                               \
    } while (0)

inline int64_t div_round_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

class Shape {
public:
    Shape() = default;
    Shape(std::initializer_list<int64_t> d) : d_(d) { check(); }
    explicit Shape(std::vector<int64_t> d) : d_(std::move(d)) { check(); }

    int len() const { return (int)d_.size(); }
    bool empty() const { return d_.empty(); }
    int64_t operator[](int i) const { return d_[i]; }
    int64_t& operator[](int i) { return d_[i]; }
    // negative indices address from the end (shape.rs:174-177, SignedIndex)
    int axis(int index) const {
        int a = index < 0 ? index + len() : index;
        DSC_CHECK(a >= 0 && a < len(), "axis " << index << " out of range for " << str());
        return a;
    }
    int64_t at(int index) const { return d_[axis(index)]; }
    const std::vector<int64_t>& dims() const { return d_; }

    int64_t element_count() const {
        int64_t n = 1;
        for (auto v : d_) n *= v;
        return n;
    }
    int64_t buffer_size() const { return element_count() * 4; }

    std::vector<int64_t> strides() const {
        std::vector<int64_t> s(d_.size());
        int64_t stride = 1;
        for (int i = len() - 1; i >= 0; --i) {
            s[i] = stride;
            stride *= d_[i];
        }
        return s;
    }

    Shape prefix_ones_to_len(int n) const {
        std::vector<int64_t> v;
        while ((int)v.size() + len() < n) v.push_back(1);
        v.insert(v.end(), d_.begin(), d_.end());
        return Shape(v);
    }
    // numpy-style broadcasting of two shapes (shape.rs:73-92)
    Shape broadcast_with(const Shape& rhs) const {
        int n = std::max(len(), rhs.len());
        Shape a = prefix_ones_to_len(n), b = rhs.prefix_ones_to_len(n);
        std::vector<int64_t> v(n);
        for (int i = 0; i < n; ++i) {
            if (a[i] == 1) v[i] = b[i];
            else if (b[i] == 1) v[i] = a[i];
            else {
                DSC_CHECK(a[i] == b[i], "cannot broadcast " << str() << " with " << rhs.str());
                v[i] = a[i];
            }
        }
        return Shape(v);
    }
    Shape reduce(int ax) const { Shape t = *this; t[ax] = 1; return t; }
    Shape resize_axis(int ax, int64_t n) const { Shape t = *this; t[ax] = n; return t; }
    Shape unpad(int ax, int64_t pad) const { Shape t = *this; t[ax] -= 2 * pad; t.check(); return t; }
    Shape pad(int ax, int64_t before, int64_t after) const { Shape t = *this; t[ax] += before + after; return t; }
    Shape insert_axis(int ax, int64_t n) const {
        Shape t = *this;
        t.d_.insert(t.d_.begin() + ax, n);
        t.check();
        return t;
    }
    Shape remove_axis(int ax) const {
        Shape t = *this;
        t.d_.erase(t.d_.begin() + ax);
        t.check();
        return t;
    }
    // [1,..,len(axis),..,1]  (shape.rs:194-198)
    Shape coord(int ax) const {
        std::vector<int64_t> v(d_.size(), 1);
        v[ax] = d_[ax];
        return Shape(v);
    }
    Shape concat(const Shape& rhs) const {  // `+` in the reference (shape.rs:312-319)
        std::vector<int64_t> v = d_;
        v.insert(v.end(), rhs.d_.begin(), rhs.d_.end());
        return Shape(v);
    }
    // [..., h, w, c] -> [..., out_h, out_w, groups, filter_h, filter_w, c/groups]  (shape.rs:120-141)
    Shape image_to_windows(int64_t filter_w, int64_t filter_h, int64_t stride_w, int64_t stride_h, int64_t groups) const {
        DSC_CHECK(len() >= 3, "image_to_windows needs [.., h, w, c]");
        int n = len();
        int64_t in_h = d_[n - 3], in_w = d_[n - 2], in_nc = d_[n - 1];
        DSC_CHECK(in_nc % groups == 0, "channels not divisible by groups");
        int64_t out_w = (in_w - filter_w) / stride_w + 1;
        int64_t out_h = (in_h - filter_h) / stride_h + 1;
        DSC_CHECK((out_w - 1) * stride_w == in_w - filter_w, "filter/stride does not tile width exactly");
        DSC_CHECK((out_h - 1) * stride_h == in_h - filter_h, "filter/stride does not tile height exactly");
        std::vector<int64_t> v(d_.begin(), d_.end() - 3);
        for (int64_t x : {out_h, out_w, groups, filter_h, filter_w, in_nc / groups}) v.push_back(x);
        return Shape(v);
    }
    Shape windows_to_image(int64_t stride_w, int64_t stride_h) const {  // shape.rs:143-156
        DSC_CHECK(len() >= 6, "windows_to_image needs 6 trailing axes");
        int n = len();
        int64_t out_h = d_[n - 6], out_w = d_[n - 5], groups = d_[n - 4], fh = d_[n - 3], fw = d_[n - 2], gnc = d_[n - 1];
        std::vector<int64_t> v(d_.begin(), d_.end() - 6);
        v.push_back((out_h - 1) * stride_h + fh);
        v.push_back((out_w - 1) * stride_w + fw);
        v.push_back(groups * gnc);
        return Shape(v);
    }

    bool operator==(const Shape& o) const { return d_ == o.d_; }
    bool operator!=(const Shape& o) const { return d_ != o.d_; }
    std::string str() const {  // "[a, b, c]" as the reference prints shapes (shape.rs:638-652)
        std::ostringstream os;
        os << "[";
        for (size_t i = 0; i < d_.size(); ++i) os << (i ? ", " : "") << d_[i];
        os << "]";
        return os.str();
    }

private:
    void check() const {
        DSC_CHECK(!d_.empty() && (int)d_.size() <= MAX_DIM, "shape must have 1.." << MAX_DIM << " axes");
        for (auto v : d_) DSC_CHECK(v > 0, "shape extents must be positive");
    }
    std::vector<int64_t> d_;
};

// One output axis of a view either walks an input axis with a step or is a broadcast (shape.rs:321-352).
struct AxisMapping {
    bool is_source = false;
    int axis = 0;
    int64_t step = 0;
    static AxisMapping broadcast() { return {}; }
    static AxisMapping source(int axis, int64_t step) { return {true, axis, step}; }
    // axes of length 1 never need a coordinate
    static AxisMapping identity(int axis, int64_t length) { return length > 1 ? source(axis, 1) : broadcast(); }
    AxisMapping stepped(int64_t m) const { return is_source ? source(axis, step * m) : broadcast(); }
    bool operator==(const AxisMapping& o) const {
        return is_source == o.is_source && (!is_source || (axis == o.axis && step == o.step));
    }
};

// input coordinate on axis a = clamp(offset[a] + sum_{i: mapping[i] -> a} step_i * out_coord[i])  
struct View {
    Shape input_shape;
    std::vector<int64_t> input_offsets;
    std::vector<AxisMapping> output_mapping;
    Shape output_shape;

    static View identity(const Shape& s) {
        View v;
        v.input_shape = s;
        v.input_offsets.assign(s.len(), 0);
        for (int i = 0; i < s.len(); ++i) v.output_mapping.push_back(AxisMapping::identity(i, s[i]));
        v.output_shape = s;
        return v;
    }
    // replicate ("clamp to edge") padding: offsets go negative and the kernel clamps (shape.rs:374-379)
    static View padded(const Shape& s, int axis, int64_t before, int64_t after) {
        View v = identity(s);
        v.input_offsets[axis] = -before;
        v.output_shape = v.output_shape.pad(axis, before, after);
        // a padded axis of input length 1 stays a broadcast, matching the reference
        return v;
    }
    static View limited(const Shape& s, int axis, int64_t start, int64_t end) {  // shape.rs:381-401
        DSC_CHECK(0 <= start && start < end && end <= s[axis], "bad limit range");
        View v = identity(s);
        v.input_offsets[axis] = start;
        v.output_mapping[axis] = AxisMapping::identity(axis, end - start);
        v.output_shape[axis] = end - start;
        return v;
    }
    static View broadcast(const Shape& in, const Shape& out) {  // shape.rs:578-602
        DSC_CHECK(in.len() <= out.len(), "broadcast to fewer axes");
        View v;
        v.input_shape = in;
        v.input_offsets.assign(in.len(), 0);
        int lead = out.len() - in.len();
        for (int i = 0; i < lead; ++i) v.output_mapping.push_back(AxisMapping::broadcast());
        for (int i = 0; i < in.len(); ++i) {
            if (in[i] == out[lead + i]) v.output_mapping.push_back(AxisMapping::identity(i, in[i]));
            else {
                DSC_CHECK(in[i] == 1, "cannot broadcast " << in.str() << " to " << out.str());
                v.output_mapping.push_back(AxisMapping::broadcast());
            }
        }
        v.output_shape = out;
        return v;
    }
    // A reshape that only inserts/removes unit axes is itself a view (shape.rs:421-455).
    static bool try_from_reshape(const Shape& in, const Shape& out, View* result) {
        if (in == out) { *result = identity(in); return true; }
        std::vector<AxisMapping> mapping;
        for (int ia = 0; ia < in.len(); ++ia) {
            if (in[ia] == 1) continue;
            for (;;) {
                if ((int)mapping.size() >= out.len()) return false;
                int64_t ol = out[(int)mapping.size()];
                if (ol == in[ia]) { mapping.push_back(AxisMapping::identity(ia, in[ia])); break; }
                if (ol != 1) return false;
                mapping.push_back(AxisMapping::broadcast());
            }
        }
        while ((int)mapping.size() < out.len()) {
            if (out[(int)mapping.size()] != 1) return false;
            mapping.push_back(AxisMapping::broadcast());
        }
        View v;
        v.input_shape = in;
        v.input_offsets.assign(in.len(), 0);
        v.output_mapping = mapping;
        v.output_shape = out;
        *result = v;
        return true;
    }

    bool is_identity() const { return *this == identity(output_shape) && input_shape == output_shape; }

    // linear output index == linear input index for every element (shape.rs:403-419)
    bool is_contiguous() const {
        if (input_shape.element_count() != output_shape.element_count()) return false;
        for (auto o : input_offsets) if (o != 0) return false;
        auto is = input_shape.strides(), os = output_shape.strides();
        for (int i = 0; i < output_shape.len(); ++i) {
            const auto& m = output_mapping[i];
            if (m.is_source && is[m.axis] * m.step != os[i]) return false;
        }
        return true;
    }

    // range of raw (unclamped) input coordinates reached on `input_axis`
    void input_span(int input_axis, int64_t* lo, int64_t* hi) const {
        int64_t mn = input_offsets[input_axis], mx = mn;
        for (int i = 0; i < output_shape.len(); ++i) {
            const auto& m = output_mapping[i];
            if (m.is_source && m.axis == input_axis) {
                int64_t off = (output_shape[i] - 1) * m.step;
                mn += std::min<int64_t>(off, 0);
                mx += std::max<int64_t>(off, 0);
            }
        }
        *lo = mn;
        *hi = mx;
    }
    bool input_needs_clamp(int input_axis) const {  // shape.rs:503-526
        int64_t lo, hi;
        input_span(input_axis, &lo, &hi);
        return lo < 0 || input_shape[input_axis] - 1 < hi;
    }
    bool any_clamp() const {
        for (int a = 0; a < input_shape.len(); ++a) if (input_needs_clamp(a)) return true;
        return false;
    }

    int input_axis_mapping_count(int input_axis) const {
        int n = 0;
        for (const auto& m : output_mapping) n += (m.is_source && m.axis == input_axis);
        return n;
    }
    // Clamping the coordinate of `output_axis` is the same as clamping the input axis it walks only
    // when it alone covers that whole input axis (shape.rs:467-488).
    bool can_pad_output(int output_axis) const {
        const auto& m = output_mapping[output_axis];
        if (!m.is_source) return true;
        if (input_axis_mapping_count(m.axis) != 1) return false;
        int64_t base = input_offsets[m.axis];
        int64_t off = (output_shape[output_axis] - 1) * m.step;
        int64_t span_min = base + std::min<int64_t>(off, 0), span_max = base + std::max<int64_t>(off, 0);
        return span_min <= 0 && input_shape[m.axis] - 1 <= span_max;
    }
    bool can_combine_with(const View& next) const {  // shape.rs:490-496
        if (output_shape != next.input_shape) return false;
        for (int a = 0; a < output_shape.len(); ++a)
            if (!can_pad_output(a) && next.input_needs_clamp(a)) return false;
        return true;
    }
    // this followed by `next` as a single view; requires can_combine_with(next)  (shape.rs:532-567)
    View through(const View& next) const {
        DSC_CHECK(can_combine_with(next), "views do not compose");
        View r;
        r.input_shape = input_shape;
        r.input_offsets = input_offsets;
        for (int i = 0; i < output_shape.len(); ++i) {
            const auto& m = output_mapping[i];
            if (m.is_source) r.input_offsets[m.axis] += m.step * next.input_offsets[i];
        }
        for (const auto& outer : next.output_mapping)
            r.output_mapping.push_back(outer.is_source ? output_mapping[outer.axis].stepped(outer.step) : AxisMapping::broadcast());
        r.output_shape = next.output_shape;
        return r;
    }
    View transposed() const {  // swap the last two output axes (shape.rs:569-576)
        int n = output_shape.len();
        DSC_CHECK(n >= 2, "transpose needs two axes");
        View t = *this;
        std::swap(t.output_mapping[n - 2], t.output_mapping[n - 1]);
        std::swap(t.output_shape[n - 2], t.output_shape[n - 1]);
        return t;
    }
    View permute_axes(const std::vector<int>& perm) const {  // shape.rs:627-635
        View t = *this;
        t.output_mapping.clear();
        std::vector<int64_t> s;
        for (int p : perm) {
            t.output_mapping.push_back(output_mapping[p]);
            s.push_back(output_shape[p]);
        }
        t.output_shape = Shape(s);
        return t;
    }

    bool operator==(const View& o) const {
        return input_shape == o.input_shape && input_offsets == o.input_offsets && output_mapping == o.output_mapping &&
               output_shape == o.output_shape;
    }
    bool operator!=(const View& o) const { return !(*this == o); }
};

// Views applied in order from the producer's buffer towards the consumer.  Between consecutive views
// (and between the last view and the consumer) the *linear* element index is preserved, which is
// exactly what a reshape means.  `views` may be empty: the identity on linear indices.
struct ViewChain {
    std::vector<View> views;
    int64_t input_count = 0;   // elements of the producer
    int64_t output_count = 0;  // elements the consumer addresses

    static ViewChain identity(int64_t count) {
        ViewChain c;
        c.input_count = c.output_count = count;
        return c;
    }
    static ViewChain of(const View& v) {
        ViewChain c = identity(v.input_shape.element_count());
        c.push(v);
        return c;
    }
    bool is_identity() const { return views.empty(); }

    // append a view on the consumer side, folding it into the previous one when the reference's
    // composition rules allow (shape.rs:490-567) and keeping a reshape boundary otherwise
    void push(const View& v) {
        DSC_CHECK(v.input_shape.element_count() == output_count, "view chain element count mismatch");
        output_count = v.output_shape.element_count();
        if (v.is_contiguous()) return;  // pure reshape: linear index unchanged
        if (!views.empty()) {
            View last = views.back();
            if (last.output_shape != v.input_shape) {
                View match;
                if (View::try_from_reshape(last.output_shape, v.input_shape, &match)) last = last.through(match);
            }
            if (last.can_combine_with(v)) {
                views.back() = last.through(v);
                if (views.back().is_contiguous()) views.pop_back();
                return;
            }
        }
        views.push_back(v);
    }
    void append(const ViewChain& next) {
        DSC_CHECK(next.input_count == output_count, "view chain element count mismatch");
        if (next.views.empty()) return;
        for (const auto& v : next.views) push(v);
        output_count = next.output_count;
    }
    // number of distinct producer elements that can be addressed (for traffic accounting)
    int64_t addressed_count() const { return std::min(input_count, output_count); }

    bool operator==(const ViewChain& o) const {
        return views == o.views && input_count == o.input_count && output_count == o.output_count;
    }
    bool operator!=(const ViewChain& o) const { return !(*this == o); }
};

}  // namespace descent

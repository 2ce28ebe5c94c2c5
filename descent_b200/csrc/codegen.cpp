// CUDA C emission for every cluster kind.  Semantics per kernel follow SURVEY.md Appendix A and the
// reference kernels cited at each generator; the code shape is B200-first: views are index
// arithmetic on compile-time constants, identity operands move as 128-bit vectors, reductions are
// cooperative and deterministic, GEMM operands (including the im2col view of conv2d) are gathered
// straight into shared-memory tiles, scatter_add is a sort + segmented sum instead of float atomics.
#include "codegen.hpp"

#include <cmath>
#include <functional>
#include <map>
#include <cstdlib>
#include <cstdio>

namespace descent {

namespace {

std::string replace_all(std::string s, const std::string& from, const std::string& to) {
    size_t pos = 0;
    while ((pos = s.find(from, pos)) != std::string::npos) {
        s.replace(pos, from.size(), to);
        pos += to.size();
    }
    return s;
}
// measurement hook: the halo conv kernels' MMA-issuing thread chosen by `tid == 0` (default) or by elect.sync under a
// warp-uniform condition (DSC_HALO_ELECT=1; see dsc_elect_one in the prelude)
bool halo_elect() {
    static const bool v = [] { const char* e = std::getenv("DSC_HALO_ELECT"); return e && std::atoi(e) != 0; }();
    return v;
}

std::string subst(std::string s, const std::vector<std::pair<std::string, std::string>>& kv) {
    for (const auto& p : kv) s = replace_all(s, "{{" + p.first + "}}", p.second);
    s = replace_all(s, "{{WARP_EXPR}}", halo_elect() ? "__shfl_sync(0xffffffffu, tid >> 5, 0)" : "tid >> 5");
    s = replace_all(s, "{{TMEM_EXPR}}", halo_elect() ? "__shfl_sync(0xffffffffu, *tmem_slot, 0)" : "*tmem_slot");
    s = replace_all(s, "{{ISSUE_COND}}", halo_elect() ? "warp == 0 && dsc_elect_one()" : "tid == 0");
    DSC_CHECK(s.find("{{") == std::string::npos, "unsubstituted placeholder in kernel template: " << s.substr(s.find("{{"), 24));
    return s;
}
std::string num(int64_t v) { return std::to_string(v); }
std::string unum(int64_t v) { return std::to_string(v) + "u"; }

int64_t pow2_ceil(int64_t v) { int64_t p = 1; while (p < v) p <<= 1; return p; }
int64_t pow2_floor(int64_t v) { int64_t p = 1; while (p * 2 <= v) p <<= 1; return p; }

// An index kept symbolically as a sum of non-overlapping mixed-radix fields: value = sum(expr_t * stride_t)
// with expr_t in [0, range_t) and stride_t * range_t <= stride of the next larger term.  Decoding such a sum
// into the coordinates of a shape is done field by field whenever the fields line up with the shape's digits,
// which keeps e.g. the row part (image, y, x) and the column part (fy, fx, channel) of conv2d's im2col index
// separate, so the compiler hoists each out of the loop over the other.
struct Term {
    std::string expr;
    int64_t stride;
    int64_t range;
};
std::vector<Term> linear_term(const std::string& e, int64_t count) { return {{e, 1, count}}; }

// Emits statements that map a consumer element (given as terms over the chain's output index space) to the
// producer buffer index, view by view from the consumer side (offset + sum(step*coord),
// clamped only where the view can leave the axis).  Returns the resulting index expression.
std::string emit_chain(std::ostringstream& os, const ViewChain& chain, std::vector<Term> terms, int& uniq, const char* indent = "    ") {
    auto materialize = [&](const std::vector<Term>& ts) {
        std::ostringstream sum;
        bool any = false;
        for (const auto& t : ts) {
            if (t.range <= 1 && t.expr == "0") continue;
            sum << (any ? " + " : "") << "(unsigned)(" << t.expr << ")";
            if (t.stride != 1) sum << " * " << unum(t.stride);
            any = true;
        }
        return any ? sum.str() : std::string("0u");
    };
    for (int vi = (int)chain.views.size() - 1; vi >= 0; --vi) {
        const View& v = chain.views[vi];
        const int id = uniq++;
        auto ostr = v.output_shape.strides();
        auto istr = v.input_shape.strides();
        // 1. coordinates of the needed output axes, term-wise when the fields align with the digits
        std::vector<std::string> coord(v.output_shape.len());
        bool aligned = true;
        for (int i = 0; i < v.output_shape.len() && aligned; ++i) {
            const auto& m = v.output_mapping[i];
            if (!(m.is_source && v.output_shape[i] > 1)) continue;
            const int64_t lo = ostr[i], hi = ostr[i] * v.output_shape[i];
            std::ostringstream parts;
            bool any = false;
            for (const auto& t : terms) {
                if (t.range <= 1 && t.expr == "0") continue;
                const int64_t flo = t.stride, fhi = t.stride * t.range;
                if (fhi <= lo) continue;  // entirely below this digit: no carry because fields do not overlap
                if (flo >= hi) {
                    if (flo % hi != 0) aligned = false;  // would leak into this digit
                    continue;
                }
                std::string e;
                if (flo >= lo && flo % lo == 0 && fhi <= hi) {
                    e = flo == lo ? t.expr : "(" + t.expr + ") * " + num(flo / lo);
                } else if (flo <= lo && lo % flo == 0 && (fhi <= hi || hi % flo == 0)) {
                    e = lo == flo ? "(" + t.expr + ")" : "(" + t.expr + ") / " + num(lo / flo);
                    if (fhi > hi) e = "(" + e + ") % " + num(v.output_shape[i]);
                } else {
                    aligned = false;
                    break;
                }
                parts << (any ? " + " : "") << e;
                any = true;
            }
            coord[i] = any ? parts.str() : "0";
        }
        if (!aligned) {
            os << indent << "const unsigned l" << id << " = " << materialize(terms) << ";\n";
            int64_t lead = 1;
            for (int i = 0; i < v.output_shape.len(); ++i) {
                const auto& m = v.output_mapping[i];
                if (m.is_source && v.output_shape[i] > 1) {
                    std::ostringstream e;
                    e << "l" << id;
                    if (ostr[i] != 1) e << " / " << unum(ostr[i]);
                    if (lead != 1) e << " % " << unum(v.output_shape[i]);
                    coord[i] = e.str();
                }
                lead *= v.output_shape[i];
            }
        }
        for (int i = 0; i < v.output_shape.len(); ++i) {
            const auto& m = v.output_mapping[i];
            if (m.is_source && v.output_shape[i] > 1) os << indent << "const int c" << id << "_" << i << " = (int)(" << coord[i] << ");\n";
        }
        // 2. one term per input axis
        std::vector<Term> next;
        for (int a = 0; a < v.input_shape.len(); ++a) {
            std::ostringstream t;
            bool has_terms = false;
            for (int i = 0; i < v.output_shape.len(); ++i) {
                const auto& m = v.output_mapping[i];
                if (m.is_source && m.axis == a && v.output_shape[i] > 1) {
                    t << (has_terms ? " + " : "");
                    if (m.step == 1) t << "c" << id << "_" << i;
                    else t << "(" << m.step << ")*c" << id << "_" << i;
                    has_terms = true;
                }
            }
            const int64_t off = v.input_offsets[a], len = v.input_shape[a];
            if (!has_terms) {
                const int64_t fixed = std::min<int64_t>(std::max<int64_t>(off, 0), len - 1);
                if (fixed != 0) next.push_back({num(fixed), istr[a], len});
                continue;
            }
            std::string expr = t.str();
            if (off != 0) expr = expr + " + (" + num(off) + ")";
            if (v.input_needs_clamp(a)) expr = "min(max(" + expr + ", 0), " + num(len - 1) + ")";
            os << indent << "const int a" << id << "_" << a << " = " << expr << ";\n";
            next.push_back({"a" + num(id) + "_" + num(a), istr[a], len});
        }
        std::sort(next.begin(), next.end(), [](const Term& x, const Term& y) { return x.stride > y.stride; });
        terms = next;
    }
    const int id = uniq++;
    os << indent << "const unsigned x" << id << " = " << materialize(terms) << ";\n";
    return "x" + num(id);
}
std::string emit_chain(std::ostringstream& os, const ViewChain& chain, const std::string& e, int& uniq, const char* indent = "    ") {
    if (chain.views.empty()) return e;
    return emit_chain(os, chain, linear_term(e, chain.output_count), uniq, indent);
}

// Largest power of two R <= 4 such that every R-aligned group of R consecutive consumer elements is read from
// R consecutive, R-aligned producer elements.  Per view: the innermost output axis must walk the innermost input
// axis with step 1 and without clamping, and every quantity that could break the alignment of a group (extents
// of those axes, the offset, steps of other output axes onto the same input axis) must be a multiple of R.
// Reshapes between views preserve linear indices, so the conditions compose.
int64_t chain_vector_run(const ViewChain& chain, int64_t innermost_extent) {
    auto pow2_divisor = [](int64_t r, int64_t v) {
        v = v < 0 ? -v : v;
        if (v == 0) return r;
        while (r > 1 && v % r != 0) r /= 2;
        return r;
    };
    int64_t run = pow2_divisor(4, innermost_extent);
    for (const View& v : chain.views) {
        int ao = v.output_shape.len() - 1, ai = v.input_shape.len() - 1;
        while (ao > 0 && v.output_shape[ao] == 1) --ao;
        while (ai > 0 && v.input_shape[ai] == 1) --ai;
        const AxisMapping& m = v.output_mapping[ao];
        if (!m.is_source || m.axis != ai || m.step != 1 || v.input_needs_clamp(ai)) return 1;
        run = pow2_divisor(run, v.output_shape[ao]);
        run = pow2_divisor(run, v.input_shape[ai]);
        run = pow2_divisor(run, v.input_offsets[ai]);
        for (int i = 0; i < v.output_shape.len(); ++i)
            if (i != ao && v.output_mapping[i].is_source && v.output_mapping[i].axis == ai && v.output_shape[i] > 1)
                run = pow2_divisor(run, v.output_mapping[i].step);
    }
    return run;
}

// The same question for runs along another axis of the consumer's logical shape (e.g. along m of a GEMM operand
// [b, m, k] whose memory is contiguous in m: the transposed operands of the weight-gradient GEMMs).  The
// consumer-side view must have exactly that logical shape and must map the axis, with step 1, onto its innermost
// input axis; from there on the innermost-axis conditions above apply.
int64_t chain_vector_run_axis(const ViewChain& chain, const Shape& arg_shape, int axis) {
    if (axis == arg_shape.len() - 1) return chain_vector_run(chain, arg_shape[axis]);
    if (chain.views.empty()) return 1;  // row-major buffer: another axis is never the contiguous one
    const View& v = chain.views.back();
    if (v.output_shape != arg_shape) return 1;
    auto pow2_divisor = [](int64_t r, int64_t x) {
        x = x < 0 ? -x : x;
        if (x == 0) return r;
        while (r > 1 && x % r != 0) r /= 2;
        return r;
    };
    int ai = v.input_shape.len() - 1;
    while (ai > 0 && v.input_shape[ai] == 1) --ai;
    const AxisMapping& m = v.output_mapping[axis];
    if (!m.is_source || m.axis != ai || m.step != 1 || v.input_needs_clamp(ai)) return 1;
    int64_t run = pow2_divisor(4, arg_shape[axis]);
    run = pow2_divisor(run, v.input_shape[ai]);
    run = pow2_divisor(run, v.input_offsets[ai]);
    for (int i = 0; i < v.output_shape.len(); ++i)
        if (i != axis && v.output_mapping[i].is_source && v.output_mapping[i].axis == ai && v.output_shape[i] > 1)
            run = pow2_divisor(run, v.output_mapping[i].step);
    ViewChain rest;
    rest.views.assign(chain.views.begin(), chain.views.end() - 1);
    return std::min(run, chain_vector_run(rest, v.input_shape[ai]));
}

double chain_bytes(const Graph& g, const ClusterInput& in) {
    (void)g;
    return 4.0 * (double)in.chain.addressed_count();
}

// ---- per-element (reference: PerElementKernel, kernel.rs:195-383) -------------------------------

// The value of a Unary / Binary / Select op as a CUDA expression; A(k) names argument k (kernel.rs:281-333 semantics:
// IEEE f32 per-element ops, u32 ops on the raw bits).
std::string op_expression(const PerElementOp& op, const std::function<std::string(int)>& A) {
    std::ostringstream os;
    switch (op.kind) {
        case PerElementOp::Unary:
            switch (op.op.unary) {
                case UnaryOp::Mov: os << A(0); break;
                case UnaryOp::Neg: os << "-" << A(0); break;
                case UnaryOp::Sqrt: os << "sqrtf(" << A(0) << ")"; break;
                case UnaryOp::Exp: os << "expf(" << A(0) << ")"; break;
                case UnaryOp::Log: os << "logf(" << A(0) << ")"; break;
                case UnaryOp::Sin: os << "sinf(" << A(0) << ")"; break;
                case UnaryOp::Cos: os << "cosf(" << A(0) << ")"; break;
                case UnaryOp::UintToFloat: os << "__uint2float_rn(__float_as_uint(" << A(0) << "))"; break;
                case UnaryOp::FloatToUint: os << "__uint_as_float(__float2uint_rz(" << A(0) << "))"; break;
            }
            break;
        case PerElementOp::Binary: {
            auto U = [&](const char* o) { os << "__uint_as_float(__float_as_uint(" << A(0) << ") " << o << " __float_as_uint(" << A(1) << "))"; };
            switch (op.op.binary) {
                case BinaryOp::Add: os << A(0) << " + " << A(1); break;
                case BinaryOp::Sub: os << A(0) << " - " << A(1); break;
                case BinaryOp::Mul: os << A(0) << " * " << A(1); break;
                case BinaryOp::Div: os << A(0) << " / " << A(1); break;
                case BinaryOp::Pow: os << "powf(" << A(0) << ", " << A(1) << ")"; break;
                case BinaryOp::UAdd: U("+"); break;
                case BinaryOp::UMul: U("*"); break;
                case BinaryOp::URem: U("%"); break;
                case BinaryOp::UBitXor: U("^"); break;
            }
            break;
        }
        case PerElementOp::Select:
            os << "(" << A(0) << (op.op.compare == CompareMode::Eq ? " == " : " > ") << A(1) << ") ? " << A(2) << " : " << A(3);
            break;
        default: fail("not an expression op");
    }
    return os.str();
}

// The straight-line program of a per-element cluster for element `e` (statements `const float t<i> = ...;`).
// Inputs flagged in `vector_load` were fetched as `vin<i>[v]`; input `register_input` (if >= 0) is not in memory
// at all: its value for this element is `acc[v]` (a GEMM epilogue evaluating the cluster on its accumulator).
// `g_load_override`, when set, may name the value of an input that is not in memory at all (dense chains: the accumulator
// element, the activation held in shared memory); an empty answer means "load it as usual".
std::function<std::string(int)>* g_load_override = nullptr;

void emit_per_element_ops(std::ostringstream& os, const Cluster& c, const CodegenOptions& opt, int& uniq, const std::vector<bool>& vector_load,
                          int register_input) {
    for (size_t oi = 0; oi < c.ops.size(); ++oi) {
        const PerElementOp& op = c.ops[oi];
        const std::string t = "t" + num(oi);
        std::function<std::string(int)> A = [&](int k) { return "t" + num(op.args[k]); };
        switch (op.kind) {
            case PerElementOp::Load: {
                const auto& in = c.inputs[op.input_index];
                const std::string named = g_load_override ? (*g_load_override)(op.input_index) : std::string();
                if (!named.empty()) {
                    os << "    const float " << t << " = " << named << ";\n";
                } else if (op.input_index == register_input) {
                    os << "    const float " << t << " = acc[v];\n";
                } else if (vector_load[op.input_index]) {
                    os << "    const float " << t << " = vin" << op.input_index << "[v];\n";
                } else {
                    std::string idx = emit_chain(os, in.chain, "e", uniq);
                    os << "    const float " << t << " = in" << op.input_index << "[" << idx << "];\n";
                }
                break;
            }
            case PerElementOp::Literal:
                os << "    const float " << t << " = __uint_as_float(" << op.op.literal_bits << "u);";
                if (!op.op.literal_is_u32) os << "  // " << op.op.literal_f32_value();
                os << "\n";
                break;
            case PerElementOp::BuiltIn: {
                std::string idx = emit_chain(os, op.chain, "e", uniq);
                if (op.op.built_in == BuiltInOp::Coord) {
                    os << "    const float " << t << " = (float)(int)(" << idx << ");\n";
                } else {
                    // flat element index of the *global* (unsharded) tensor: SURVEY.md §8e condition 3
                    const int64_t offset = (int64_t)opt.dp_rank * op.arg_shape.element_count();
                    os << "    const float " << t << " = dsc_rand(" << op.op.rand_uid << "u, " << idx << " + " << unum(offset) << ", dsc_seed);\n";
                }
                break;
            }
            case PerElementOp::Unary:
            case PerElementOp::Binary:
            case PerElementOp::Select:
                os << "    const float " << t << " = " << op_expression(op, A) << ";\n";
                break;
            case PerElementOp::Reduce: fail("reductions only appear in row clusters");
            case PerElementOp::Gather: {
                // out[.., i, ..] = values[.., F2I(index[i]), ..]  (kernel.rs:336-351)
                const int axis = op.op.axis;
                int64_t inner = 1;
                for (int d = axis + 1; d < op.shape.len(); ++d) inner *= op.shape[d];
                const int64_t len = op.shape[axis], rows = op.arg_shape[axis];
                const int id = uniq++;
                os << "    const unsigned g" << id << " = ((e / " << unum(len * inner) << ") * " << unum(rows) << " + (unsigned)__float_as_int(" << A(1)
                   << ")) * " << unum(inner) << " + (e % " << unum(inner) << ");\n";
                std::string idx = emit_chain(os, c.inputs[op.input_index].chain, "g" + num(id), uniq);
                os << "    const float " << t << " = in" << op.input_index << "[" << idx << "];\n";
                break;
            }
        }
    }
}

// Elements per thread.  Four (128-bit loads and stores on identity operands) when that still leaves a couple of waves of
// threads.  Programs over a few hundred thousand elements (the hash-grid index / interpolation / weighted-gradient kernels
// of image_fit: 262144-524288 elements, tens of dependent loads) are latency-bound, and at four elements per thread they
// fill less than half of the machine's thread slots: one element per thread there (multi-hash m = 262144: 0.419 -> 0.382
// ms/step).  Measurement hooks: DSC_PE_VEC_MIN (thread threshold, 0 = always four), DSC_PE_VEC_ALL=0 (only gathering programs).
int per_element_vector_width(const Cluster& c) {
    static const int64_t min_threads = [] { const char* e = std::getenv("DSC_PE_VEC_MIN"); return e ? std::atoll(e) : (int64_t)148 * 2048 * 2; }();
    static const bool all_small = [] { const char* e = std::getenv("DSC_PE_VEC_ALL"); return !e || std::atoi(e) != 0; }();
    if (c.element_count % 4 != 0) return 1;
    bool narrow = all_small;
    for (const auto& op : c.ops) narrow |= op.kind == PerElementOp::Gather;
    return (narrow && c.element_count / 4 < min_threads) ? 1 : 4;
}

// Body of a per-element kernel for the block `block` (an expression): loads, the straight-line program, stores.
// Uses the names in<i> / out<i> of the cluster's own inputs and outputs.
void emit_per_element_body(std::ostringstream& os, const Cluster& c, const CodegenOptions& opt, const std::string& block) {
    const int64_t n = c.element_count;
    const int vec = per_element_vector_width(c);
    os << "    const unsigned base = (" << block << " * 256u + threadIdx.x) * " << vec << "u;\n";
    os << "    if (base >= " << unum(n) << ") return;\n";
    std::vector<bool> vector_load(c.inputs.size(), false);
    std::vector<bool> is_loaded(c.inputs.size(), false);
    for (const auto& op : c.ops)
        if (op.kind == PerElementOp::Load) is_loaded[op.input_index] = true;
    if (vec == 4) {
        for (size_t i = 0; i < c.inputs.size(); ++i) {
            if (is_loaded[i] && c.inputs[i].chain.is_identity()) {
                vector_load[i] = true;
                os << "    const float4 q" << i << " = *reinterpret_cast<const float4*>(in" << i << " + base);\n";
                os << "    const float vin" << i << "[4] = {q" << i << ".x, q" << i << ".y, q" << i << ".z, q" << i << ".w};\n";
            }
        }
        for (size_t i = 0; i < c.outputs.size(); ++i) os << "    float vout" << i << "[4];\n";
        os << "    #pragma unroll\n    for (int v = 0; v < 4; ++v) {\n";
        os << "    const unsigned e = base + v;\n";
    } else {
        os << "    const unsigned e = base;\n    {\n";
    }
    int uniq = 0;
    emit_per_element_ops(os, c, opt, uniq, vector_load, -1);
    for (size_t i = 0; i < c.outputs.size(); ++i) {
        if (vec == 4) os << "    vout" << i << "[v] = t" << c.output_ops[i] << ";\n";
        else os << "    out" << i << "[e] = t" << c.output_ops[i] << ";\n";
    }
    os << "    }\n";
    if (vec == 4)
        for (size_t i = 0; i < c.outputs.size(); ++i)
            os << "    *reinterpret_cast<float4*>(out" << i << " + base) = make_float4(vout" << i << "[0], vout" << i << "[1], vout" << i
               << "[2], vout" << i << "[3]);\n";
}

int64_t per_element_blocks(const Cluster& c) { return div_round_up(div_round_up(c.element_count, per_element_vector_width(c)), 256); }

ClusterCode gen_per_element(const Graph& g, const Cluster& c, int ci, const CodegenOptions& opt, const std::string& name_suffix = "") {
    std::ostringstream os;
    const std::string name = "k" + num(ci) + name_suffix;
    os << "// " << c.label << "\n";
    os << "extern \"C\" __global__ void __launch_bounds__(256) " << name << "(";
    const bool grouped = !c.group.empty();
    for (size_t i = 0; i < c.inputs.size(); ++i) os << "const float* " << (grouped ? "gin" : "in") << i << ", ";
    for (size_t i = 0; i < c.outputs.size(); ++i) os << "float* " << (grouped ? "gout" : "out") << i << ", ";
    os << "const unsigned* dsc_step) {\n";
    os << "    const unsigned dsc_seed = dsc_step[0]; (void)dsc_seed;\n";
    int64_t blocks = 0;
    if (!grouped) {
        emit_per_element_body(os, c, opt, "blockIdx.x");
        blocks = per_element_blocks(c);
    } else {
        // block ranges select the program; each program sees its own buffers under the usual names
        size_t in_base = 0, out_base = 0;
        for (const Cluster& sub : c.group) {
            const int64_t b = per_element_blocks(sub);
            os << "    if (blockIdx.x < " << unum(blocks + b) << ") {  // " << sub.label << "\n";
            for (size_t i = 0; i < sub.inputs.size(); ++i) os << "    const float* in" << i << " = gin" << in_base + i << ";\n";
            for (size_t i = 0; i < sub.outputs.size(); ++i) os << "    float* out" << i << " = gout" << out_base + i << ";\n";
            emit_per_element_body(os, sub, opt, "(blockIdx.x - " + unum(blocks) + ")");
            os << "    return;\n    }\n";
            blocks += b;
            in_base += sub.inputs.size();
            out_base += sub.outputs.size();
        }
    }
    os << "}\n\n";

    ClusterCode code;
    code.source = os.str();
    KernelLaunch l;
    l.entry = name;
    l.grid_x = (uint32_t)blocks;
    l.label = c.label;
    l.cluster = ci;
    auto account = [&](const Cluster& sub) {
        for (const auto& in : sub.inputs) l.algorithmic_bytes += chain_bytes(g, in);
        l.algorithmic_bytes += 4.0 * (double)sub.element_count * (double)sub.outputs.size();
    };
    if (grouped) for (const Cluster& sub : c.group) account(sub);
    else account(c);
    for (const auto& in : c.inputs) l.args.push_back({KernelArg::NodeBuffer, in.node_id, 0});
    for (int out : c.outputs) l.args.push_back({KernelArg::NodeBuffer, out, 0});
    code.launches.push_back(l);
    return code;
}

// ---- row clusters (graph.cpp fuse_rows): softmax cross-entropy & co as one kernel ---------------------
// One thread per row.  A wide op is an unrolled loop over the row's K elements writing a register array, a narrow op a
// scalar, a reduction a sequential loop in ascending k (the order of the reference's ReduceKernel).  Operands from
// outside keep their view chains (element index r * K + k for wide values, r for narrow ones).
ClusterCode gen_row(const Graph& g, const Cluster& c, int ci, const CodegenOptions& opt) {
    const int64_t R = c.rows, K = c.row_length;
    std::ostringstream os;
    const std::string name = "k" + num(ci);
    os << "// " << c.label << "  [one thread per row, row in registers]\n";
    os << "extern \"C\" __global__ void __launch_bounds__(128) " << name << "(";
    for (size_t i = 0; i < c.inputs.size(); ++i) os << "const float* in" << i << ", ";
    for (size_t i = 0; i < c.outputs.size(); ++i) os << "float* out" << i << ", ";
    os << "const unsigned* dsc_step) {\n";
    os << "    const unsigned dsc_seed = dsc_step[0]; (void)dsc_seed;\n";
    os << "    constexpr unsigned K = " << K << "u;\n";
    os << "    const unsigned r = blockIdx.x * 128u + threadIdx.x;\n    if (r >= " << unum(R) << ") return;\n";
    int uniq = 0;
    for (size_t oi = 0; oi < c.ops.size(); ++oi) {
        const PerElementOp& op = c.ops[oi];
        const std::string t = "t" + num((int64_t)oi);
        // argument k as seen from an element of this op: wide values by [k], narrow ones (and constants) as scalars
        std::function<std::string(int)> A = [&](int a) {
            const PerElementOp& src = c.ops[op.args[a]];
            return "t" + num(op.args[a]) + ((src.wide && op.wide) ? "[k]" : "");
        };
        if (op.kind == PerElementOp::Literal) {
            os << "    const float " << t << " = __uint_as_float(" << op.op.literal_bits << "u);\n";
            continue;
        }
        if (op.kind == PerElementOp::Reduce) {
            const bool is_max = op.op.reduce == ReduceOp::Max;
            const std::string a = "t" + num(op.args[0]);
            os << "    float " << t << " = " << (is_max ? "__uint_as_float(0xff800000u)" : "0.f") << ";\n";
            os << "    #pragma unroll\n    for (unsigned k = 0; k < K; ++k) " << t << " = " << (is_max ? "fmaxf(" + t + ", " + a + "[k])" : t + " + " + a + "[k]") << ";\n";
            continue;
        }
        if (op.wide) os << "    float " << t << "[K];\n    #pragma unroll\n    for (unsigned k = 0; k < K; ++k) {\n    const unsigned e = r * K + k; (void)e;\n";
        else os << "    float " << t << ";\n    {\n    const unsigned e = r; (void)e;\n";
        const std::string lhs = op.wide ? t + "[k]" : t;
        switch (op.kind) {
            case PerElementOp::Load: {
                std::string idx = emit_chain(os, c.inputs[op.input_index].chain, "e", uniq);
                os << "    " << lhs << " = in" << op.input_index << "[" << idx << "];\n";
                break;
            }
            case PerElementOp::BuiltIn: {
                std::string idx = emit_chain(os, op.chain, "e", uniq);
                if (op.op.built_in == BuiltInOp::Coord) {
                    os << "    " << lhs << " = (float)(int)(" << idx << ");\n";
                } else {
                    const int64_t offset = (int64_t)opt.dp_rank * op.arg_shape.element_count();
                    os << "    " << lhs << " = dsc_rand(" << op.op.rand_uid << "u, " << idx << " + " << unum(offset) << ", dsc_seed);\n";
                }
                break;
            }
            case PerElementOp::Unary:
            case PerElementOp::Binary:
            case PerElementOp::Select:
                os << "    " << lhs << " = " << op_expression(op, A) << ";\n";
                break;
            default: fail("unexpected op in a row cluster");
        }
        os << "    }\n";
    }
    for (size_t i = 0; i < c.outputs.size(); ++i) {
        const PerElementOp& op = c.ops[c.output_ops[i]];
        const std::string t = "t" + num(c.output_ops[i]);
        if (op.wide) os << "    #pragma unroll\n    for (unsigned k = 0; k < K; ++k) out" << i << "[r * K + k] = " << t << "[k];\n";
        else os << "    out" << i << "[r] = " << t << ";\n";
    }
    os << "}\n\n";
    ClusterCode code;
    code.source = os.str();
    KernelLaunch l;
    l.entry = name;
    l.grid_x = (uint32_t)div_round_up(R, 128);
    l.block = 128;
    l.label = c.label;
    l.cluster = ci;
    for (const auto& in : c.inputs) {
        l.args.push_back({KernelArg::NodeBuffer, in.node_id, 0});
        l.algorithmic_bytes += chain_bytes(g, in);
    }
    for (size_t i = 0; i < c.outputs.size(); ++i) {
        l.args.push_back({KernelArg::NodeBuffer, c.outputs[i], 0});
        l.algorithmic_bytes += 4.0 * (double)g.ops().nodes[c.outputs[i]].shape.element_count();
    }
    code.launches.push_back(l);
    return code;
}

// ---- reduce (reference: ReduceKernel, kernel.rs:559-642) ----------------------------------------
// The reference walks K sequentially in one thread per output.  Here G threads share an output when K
// is long (G is a power of two chosen from the shape), each walks a strided slice in ascending k, and
// the G partials meet in a fixed shared-memory tree: deterministic, and coalesced in whichever of
// {k, output} direction is contiguous in memory.

const char* kReduceTemplate = R"(
// {{LABEL}}
extern "C" __global__ void __launch_bounds__(256) {{NAME}}(const float* in0, float* out0, const unsigned* dsc_step) {
    constexpr unsigned O = {{O}}u, K = {{K}}u, INNER = {{INNER}}u, G = {{G}}u, TO = 256u / G;
    constexpr unsigned S = {{S}}u, KS = (K + S - 1u) / S;  // S > 1: blockIdx.y owns a slice of k and writes a partial
    const unsigned k_lo = blockIdx.y * KS, k_hi = min(K, k_lo + KS);
    const unsigned tid = threadIdx.x;
    const unsigned g = {{G_OF_TID}};
    const unsigned ol = {{O_OF_TID}};
    const unsigned o = blockIdx.x * TO + ol;
    float acc = {{INIT}};
    if (o < O) {
        const unsigned oo = o / INNER, oi = o % INNER;
        {{UNROLL}}
        for (unsigned k = k_lo + g; k < k_hi; k += G) {
            const unsigned e = (oo * K + k) * INNER + oi;
{{CHAIN}}
            const float v = in0[{{IDX}}];
            acc = {{OP}};
        }
    }
    if (G > 1) {
        __shared__ float red[256];
        red[tid] = acc;
        __syncthreads();
        #pragma unroll
        for (unsigned s = G / 2; s > 0; s >>= 1) {
            if (g < s) {
                const float a = red[tid], v = red[tid + s * {{GSTRIDE}}];
                red[tid] = {{OP_AV}};
            }
            __syncthreads();
        }
        if (g == 0 && o < O) out0[blockIdx.y * O + o] = red[tid];
    } else if (o < O) {
        out0[blockIdx.y * O + o] = acc;
    }
}
)";

// The same reduction for four adjacent outputs per thread when they are contiguous in memory for every k (the
// reduced axis is not the innermost one and the operand's chain keeps aligned groups of four together): 128-bit
// loads, four independent accumulators, identical order of additions per output.
const char* kReduceVec4Template = R"(
// {{LABEL}}  [4 outputs per thread]
extern "C" __global__ void __launch_bounds__(256) {{NAME}}(const float* in0, float* out0, const unsigned* dsc_step) {
    constexpr unsigned O = {{O}}u, OV = O / 4u, K = {{K}}u, INNER = {{INNER}}u, G = {{G}}u, TO = 256u / G;
    constexpr unsigned S = {{S}}u, KS = (K + S - 1u) / S;  // S > 1: blockIdx.y owns a slice of k and writes a partial
    const unsigned k_lo = blockIdx.y * KS, k_hi = min(K, k_lo + KS);
    const unsigned tid = threadIdx.x;
    const unsigned g = tid / TO;
    const unsigned ol = tid % TO;
    const unsigned ov = blockIdx.x * TO + ol, o = ov * 4u;
    float4 acc = make_float4({{INIT}}, {{INIT}}, {{INIT}}, {{INIT}});
    if (ov < OV) {
        const unsigned oo = o / INNER, oi = o % INNER;
        {{UNROLL}}
        for (unsigned k = k_lo + g; k < k_hi; k += G) {
            const unsigned e = (oo * K + k) * INNER + oi;
{{CHAIN}}
            const float4 v = *reinterpret_cast<const float4*>(in0 + {{IDX}});
            acc.x = {{OPX}}; acc.y = {{OPY}}; acc.z = {{OPZ}}; acc.w = {{OPW}};
        }
    }
    if (G > 1) {
        __shared__ float4 red[256];
        red[tid] = acc;
        __syncthreads();
        #pragma unroll
        for (unsigned s = G / 2; s > 0; s >>= 1) {
            if (g < s) {
                float4 acc = red[tid];
                const float4 v = red[tid + s * TO];
                acc.x = {{OPX}}; acc.y = {{OPY}}; acc.z = {{OPZ}}; acc.w = {{OPW}};
                red[tid] = acc;
            }
            __syncthreads();
        }
        if (g == 0 && ov < OV) *reinterpret_cast<float4*>(out0 + blockIdx.y * O + o) = red[tid];
    } else if (ov < OV) {
        *reinterpret_cast<float4*>(out0 + blockIdx.y * O + o) = acc;
    }
}
)";

const char* kReduceSplitTemplate = R"(
// k-slice partials of {{LABEL}}, combined in ascending slice order
extern "C" __global__ void __launch_bounds__(256) {{NAME}}(const float* ws, float* out0, const unsigned* dsc_step) {
    constexpr unsigned COUNT = {{COUNT}}u, S = {{S}}u;
    const unsigned i = blockIdx.x * 256u + threadIdx.x;
    if (i >= COUNT) return;
    float acc = ws[i];
    #pragma unroll 8
    for (unsigned s = 1; s < S; ++s) {
        const float v = ws[s * COUNT + i];
        acc = {{OP}};
    }
    out0[i] = acc;
}
)";

ClusterCode gen_reduce(const Graph& g, const Cluster& c, int ci, const CodegenOptions& opt, const std::string& name_suffix = "") {
    const OpNode& node = g.ops().nodes[c.node_id];
    const ClusterInput& in = c.inputs[0];
    const int axis = node.op.axis;
    const int64_t K = in.arg_shape[axis];
    int64_t inner = 1;
    for (int d = axis + 1; d < in.arg_shape.len(); ++d) inner *= in.arg_shape[d];
    const int64_t O = node.shape.element_count();
    int64_t G = 1;
    if (K > 16) {
        const int64_t want = pow2_ceil(div_round_up((int64_t)opt.sm_count * 1024, O));
        G = std::max<int64_t>(1, std::min<int64_t>({want, (int64_t)256, pow2_floor(K / 4)}));
    }
    // is the reduced axis the contiguous one?  probe the chain around the middle of the tensor
    bool kfast = inner == 1;
    {
        const int64_t total = in.arg_shape.element_count();
        const int64_t mid = (total / 2 / (K * inner)) * K * inner;  // k = 0, oi = 0 of a middle row
        if (K > 1 && mid + inner < total && inner > 1) {
            int64_t dk = std::llabs(eval_chain(in.chain, mid + inner) - eval_chain(in.chain, mid));
            int64_t d_o = std::llabs(eval_chain(in.chain, mid + 1) - eval_chain(in.chain, mid));
            kfast = dk != 0 && (d_o == 0 || dk < d_o);
        }
    }
    const bool vec4 = !kfast && inner % 4 == 0 && O % 4 == 0 && chain_vector_run(in.chain, in.arg_shape.at(-1)) == 4;
    if (vec4 && K > 16) {
        const int64_t want = pow2_ceil(div_round_up((int64_t)opt.sm_count * 1024, O / 4));
        G = std::max<int64_t>(1, std::min<int64_t>({want, (int64_t)256, pow2_floor(K / 4)}));
    }
    // few outputs and a long axis: slice k across blockIdx.y as well, partials to scratch, combined in slice order
    int64_t S = 1;
    const int64_t blocks = div_round_up(vec4 ? O / 4 : O, 256 / G);
    if (blocks < 2 * opt.sm_count && K / G >= 64) S = std::max<int64_t>(1, std::min<int64_t>(div_round_up(4 * opt.sm_count, blocks), K / (G * 16)));
    std::ostringstream chain;
    int uniq = 0;
    std::string idx = emit_chain(chain, in.chain, "e", uniq, "            ");
    const bool is_max = node.op.reduce == ReduceOp::Max;
    const std::string name = "k" + num(ci) + name_suffix;
    ClusterCode code;
    if (vec4)
        code.source = subst(kReduceVec4Template,
                            {{"LABEL", c.label}, {"NAME", name}, {"O", num(O)}, {"K", num(K)}, {"INNER", num(inner)}, {"G", num(G)}, {"S", num(S)},
                             {"INIT", is_max ? "__uint_as_float(0xff800000u)" : "0.f"}, {"UNROLL", K <= 16 ? "#pragma unroll" : "#pragma unroll 4"},
                             {"CHAIN", chain.str()}, {"IDX", idx}, {"OPX", is_max ? "fmaxf(acc.x, v.x)" : "acc.x + v.x"},
                             {"OPY", is_max ? "fmaxf(acc.y, v.y)" : "acc.y + v.y"}, {"OPZ", is_max ? "fmaxf(acc.z, v.z)" : "acc.z + v.z"},
                             {"OPW", is_max ? "fmaxf(acc.w, v.w)" : "acc.w + v.w"}});
    else
    code.source = subst(kReduceTemplate, {{"LABEL", c.label}, {"NAME", name}, {"O", num(O)}, {"K", num(K)}, {"INNER", num(inner)}, {"G", num(G)}, {"S", num(S)},
                                          {"G_OF_TID", kfast ? "tid % G" : "tid / TO"}, {"O_OF_TID", kfast ? "tid / G" : "tid % TO"},
                                          {"GSTRIDE", kfast ? "1u" : "TO"}, {"INIT", is_max ? "__uint_as_float(0xff800000u)" : "0.f"},
                                          {"UNROLL", K <= 16 ? "#pragma unroll" : "#pragma unroll 4"}, {"CHAIN", chain.str()}, {"IDX", idx},
                                          {"OP", is_max ? "fmaxf(acc, v)" : "acc + v"}, {"OP_AV", is_max ? "fmaxf(a, v)" : "a + v"}});
    KernelLaunch l;
    l.entry = name;
    l.grid_x = (uint32_t)blocks;
    l.grid_y = (uint32_t)S;
    l.label = c.label;
    l.cluster = ci;
    l.args = {{KernelArg::NodeBuffer, in.node_id, 0}};
    if (S > 1) l.args.push_back({KernelArg::Scratch, -1, 0});
    else l.args.push_back({KernelArg::NodeBuffer, c.outputs[0], 0});
    l.algorithmic_bytes = chain_bytes(g, in) + 4.0 * (double)O;
    code.launches.push_back(l);
    if (S > 1) {
        code.scratch_bytes = S * O * 4;
        const std::string sname = name + "_slices";
        code.source += subst(kReduceSplitTemplate, {{"LABEL", c.label}, {"NAME", sname}, {"COUNT", num(O)}, {"S", num(S)},
                                                    {"OP", is_max ? "fmaxf(acc, v)" : "acc + v"}});
        KernelLaunch sl;
        sl.entry = sname;
        sl.grid_x = (uint32_t)div_round_up(O, 256);
        sl.label = "ReduceSlices " + c.label;
        sl.cluster = ci;
        sl.args = {{KernelArg::Scratch, -1, 0}, {KernelArg::NodeBuffer, c.outputs[0], 0}};
        code.launches.push_back(sl);
    }
    return code;
}

// ---- matmul (reference: MatMulKernel + kernel_matmul.glsl) ---------------------------------------
// Strict-FP32 SIMT GEMM, JIT-specialised per shape.  Operands are fetched through their chains into
// registers, then into double-buffered k-major shared tiles (zero fill outside M/N/K,
// kernel.rs:461-487), so the im2col "matrix" of conv2d and every transpose exist only as index
// arithmetic and the next tile's loads overlap the current tile's FMAs.  TMxTN register tile per
// thread, 128-bit shared loads, explicit fmaf.  Products are accumulated in ascending k inside a
// split and splits are summed in ascending order.

const char* kMatMulTemplate = R"(
// {{LABEL}}
{{PRO_FUNCS}}extern "C" __global__ void __launch_bounds__({{NT}}) {{NAME}}(const float* A, const float* B, float* C, {{PRO_PARAMS}}const unsigned* dsc_step) {
    constexpr int BM = {{BM}}, BN = {{BN}}, BK = {{BK}}, TM = {{TM}}, TN = {{TN}}, NT = {{NT}};
    constexpr int TX = BN / TN, TY = BM / TM, NC = TX * TY;
    constexpr int M = {{M}}, N = {{N}}, K = {{K}}, KC = {{KC}}, BC = {{BC}};
    constexpr int TILES_N = (N + BN - 1) / BN;
    constexpr int AS = BM + 4, BS = BN + 4;
    constexpr int LA = (BM * BK + NT - 1) / NT, LB = (BK * BN + NT - 1) / NT;
    __shared__ __align__(16) float As[2][BK * AS];
    __shared__ __align__(16) float Bs[2][BK * BS];
    const int tid = threadIdx.x;
    const int tile_m = blockIdx.x / TILES_N, tile_n = blockIdx.x % TILES_N;
    const int batch = blockIdx.y, split = blockIdx.z;
    const int m0 = tile_m * BM, n0 = tile_n * BN;
    const int k_begin = split * KC;
    const int k_end = min(K, k_begin + KC);
    const int tx = tid % TX, ty = tid / TX;
    float acc[TM][TN];
    #pragma unroll
    for (int i = 0; i < TM; ++i)
        #pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    float ra[LA], rb[LB];
    auto load_tile = [&](int k0) {
        #pragma unroll
        for (int j = 0; j < LA; ++j) {
            const int i = tid + j * NT;
            {{A_DECODE}}
            const int gm = m0 + lm, gk = k0 + lk;
            float v = 0.f;
            if ((BM * BK % NT == 0 || i < BM * BK) && gm < M && gk < k_end && ({{A_VALID}})) {
{{A_CHAIN}}
                v = {{A_LOAD1}};
            }
            ra[j] = v;
        }
        #pragma unroll
        for (int j = 0; j < LB; ++j) {
            const int i = tid + j * NT;
            {{B_DECODE}}
            const int gk = k0 + lk, gn = n0 + ln;
            float v = 0.f;
            if ((BK * BN % NT == 0 || i < BK * BN) && gn < N && gk < k_end) {
{{B_CHAIN}}
                v = {{B_LOAD1}};
            }
            rb[j] = v;
        }
    };
    auto store_tile = [&](int buf) {
        #pragma unroll
        for (int j = 0; j < LA; ++j) {
            const int i = tid + j * NT;
            {{A_DECODE}}
            if (BM * BK % NT == 0 || i < BM * BK) As[buf][lk * AS + lm] = ra[j];
        }
        #pragma unroll
        for (int j = 0; j < LB; ++j) {
            const int i = tid + j * NT;
            {{B_DECODE}}
            if (BK * BN % NT == 0 || i < BK * BN) Bs[buf][lk * BS + ln] = rb[j];
        }
    };
    load_tile(k_begin);
    store_tile(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        const bool more = k0 + BK < k_end;
        if (more) load_tile(k0 + BK);
        if (NC == NT || tid < NC) {
            const float* as = As[buf] + ty * TM;
            const float* bs = Bs[buf] + tx * TN;
            #pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float a[TM], b[TN];
                if (TM % 4 == 0) {
                    #pragma unroll
                    for (int i = 0; i < TM; i += 4) {
                        const float4 q = *reinterpret_cast<const float4*>(as + kk * AS + i);
                        a[i] = q.x; a[i + 1] = q.y; a[i + 2] = q.z; a[i + 3] = q.w;
                    }
                } else {
                    #pragma unroll
                    for (int i = 0; i < TM; ++i) a[i] = as[kk * AS + i];
                }
                if (TN % 4 == 0) {
                    #pragma unroll
                    for (int j = 0; j < TN; j += 4) {
                        const float4 q = *reinterpret_cast<const float4*>(bs + kk * BS + j);
                        b[j] = q.x; b[j + 1] = q.y; b[j + 2] = q.z; b[j + 3] = q.w;
                    }
                } else if (TN % 2 == 0) {
                    #pragma unroll
                    for (int j = 0; j < TN; j += 2) {
                        const float2 q = *reinterpret_cast<const float2*>(bs + kk * BS + j);
                        b[j] = q.x; b[j + 1] = q.y;
                    }
                } else {
                    #pragma unroll
                    for (int j = 0; j < TN; ++j) b[j] = bs[kk * BS + j];
                }
                #pragma unroll
                for (int i = 0; i < TM; ++i)
                    #pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
        if (more) store_tile(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    if (NC != NT && tid >= NC) return;
    #pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gm = m0 + ty * TM + i;
        if (gm >= M) continue;
        const int gn0 = n0 + tx * TN;
        float* crow = C + {{C_ROW}};
        if (TN % 4 == 0 && N % 4 == 0) {
            #pragma unroll
            for (int j = 0; j < TN; j += 4)
                if (gn0 + j < N) *reinterpret_cast<float4*>(crow + gn0 + j) = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
        } else if (TN % 2 == 0 && N % 2 == 0) {
            #pragma unroll
            for (int j = 0; j < TN; j += 2)
                if (gn0 + j < N) *reinterpret_cast<float2*>(crow + gn0 + j) = make_float2(acc[i][j], acc[i][j + 1]);
        } else {
            #pragma unroll
            for (int j = 0; j < TN; ++j)
                if (gn0 + j < N) crow[gn0 + j] = acc[i][j];
        }
    }
}
)";

#include "gemm_tc_template.inc"
#include "gemm_tc_async_template.inc"
#include "halo_conv_template.inc"
#include "halo_conv_pipelined_template.inc"
#include "thin_gemm_template.inc"
#include "halo_wgrad_template.inc"

const char* kSplitSumTemplate = R"(
// split-K partial sums of {{LABEL}}: 32 outputs x 8 split lanes per CTA.  Lane g adds splits g, g + 8, ... in ascending
// order, then the eight lane sums are added in lane order: a fixed order, independent of timing.
extern "C" __global__ void __launch_bounds__(256) {{NAME}}(const float* ws, float* out0, const unsigned* dsc_step) {
    constexpr unsigned COUNT = {{COUNT}}u, S = {{S}}u;
    __shared__ float red[8][32];
    const unsigned tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;
    const unsigned i = blockIdx.x * 32u + tx;
    float part = 0.f;
    if (i < COUNT) {
        #pragma unroll 4
        for (unsigned s = ty; s < S; s += 8u) part += ws[s * COUNT + i];
    }
    red[ty][tx] = part;
    __syncthreads();
    if (ty != 0 || i >= COUNT) return;
    float acc = red[0][tx];
    #pragma unroll
    for (unsigned g = 1; g < 8u; ++g) acc += red[g][tx];
    out0[i] = acc;
}
)";

// The same for TWO independent sets of partials in one launch (a weight gradient's partial products and the bias column
// sums its kernel produced on the side): the first blocks own set 0, the rest set 1; identical order of additions.
const char* kSplitSumPairTemplate = R"(
// split-K partial sums of {{LABEL}} and of its column sums, one launch
extern "C" __global__ void __launch_bounds__(256) {{NAME}}(const float* ws0, float* dst0, const float* ws1, float* dst1, const unsigned* dsc_step) {
    constexpr unsigned COUNT0 = {{COUNT0}}u, COUNT1 = {{COUNT1}}u, S = {{S}}u, BLOCKS0 = (COUNT0 + 31u) / 32u;
    __shared__ float red[8][32];
    const bool second = blockIdx.x >= BLOCKS0;
    const float* ws = second ? ws1 : ws0;
    float* out0 = second ? dst1 : dst0;
    const unsigned COUNT = second ? COUNT1 : COUNT0;
    const unsigned tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;
    const unsigned i = (second ? blockIdx.x - BLOCKS0 : blockIdx.x) * 32u + tx;
    float part = 0.f;
    if (i < COUNT) {
        #pragma unroll 4
        for (unsigned s = ty; s < S; s += 8u) part += ws[s * COUNT + i];
    }
    red[ty][tx] = part;
    __syncthreads();
    if (ty != 0 || i >= COUNT) return;
    float acc = red[0][tx];
    #pragma unroll
    for (unsigned g = 1; g < 8u; ++g) acc += red[g][tx];
    out0[i] = acc;
}
)";

// launches that add the S partials of a product (scratch offset 0 -> product_node) and, when the kernel also produced
// column-sum partials (scratch colsum_offset -> colsum_node), both in one kernel
void emit_split_sums(ClusterCode* out, const std::string& name, int ci, const std::string& label, int64_t S, int64_t out_count, int product_node,
                     bool colsum, int64_t colsum_count, int64_t colsum_offset, int colsum_node) {
    KernelLaunch s;
    s.cluster = ci;
    if (colsum) {
        s.entry = name + "_splitsums";
        out->source += subst(kSplitSumPairTemplate, {{"LABEL", label}, {"NAME", s.entry}, {"COUNT0", num(out_count)}, {"COUNT1", num(colsum_count)}, {"S", num(S)}});
        s.grid_x = (uint32_t)(div_round_up(out_count, 32) + div_round_up(colsum_count, 32));
        s.label = "SplitSum (+ column sums) " + label;
        s.args = {{KernelArg::Scratch, -1, 0}, {KernelArg::NodeBuffer, product_node, 0}, {KernelArg::Scratch, -1, colsum_offset}, {KernelArg::NodeBuffer, colsum_node, 0}};
    } else {
        s.entry = name + "_splitsum";
        out->source += subst(kSplitSumTemplate, {{"LABEL", label}, {"NAME", s.entry}, {"COUNT", num(out_count)}, {"S", num(S)}});
        s.grid_x = (uint32_t)div_round_up(out_count, 32);
        s.label = "SplitSum " + label;
        s.args = {{KernelArg::Scratch, -1, 0}, {KernelArg::NodeBuffer, product_node, 0}};
    }
    out->launches.push_back(s);
}

struct GemmTile { int bm, bn, bk, tm, tn, nt; };

GemmTile choose_gemm_tile(int64_t M, int64_t N) {
    GemmTile t;
    t.tn = N > 64 ? 8 : N > 16 ? 4 : N > 8 ? 2 : 1;
    t.bn = (int)std::min<int64_t>(div_round_up(N, t.tn) * t.tn, 128);
    const int tx = t.bn / t.tn;
    t.tm = (t.tn == 4 && t.bn <= 32) ? 4 : 8;
    int ty = std::min(256 / tx, 256 / t.tm);
    if (M < (int64_t)ty * t.tm) {  // short output: cover M with as few thread rows as possible
        ty = (int)div_round_up(M, 8);
        t.tm = (int)div_round_up(M, ty);
    }
    t.bm = ty * t.tm;
    t.nt = (int)div_round_up(ty * tx, 32) * 32;
    t.bk = (t.bm + t.bn + 8) * 32 * 8 <= 48 * 1024 ? 32 : 16;
    return t;
}


// ---- per-element epilogues of GEMM kernels ---------------------------------------------------------
// `<kernel>_store4(index, r, C, ..., seed)` receives four consecutive elements of the product starting at linear
// index `index`.  Without an absorbed cluster it stores them; with one it evaluates that cluster's program on them
// (the cluster's other operands are loaded at the same element indices, 128-bit where they are plain arrays) and
// stores the cluster's outputs instead.
struct EpilogueCode {
    std::string store4;   // the device function, emitted ahead of the kernel
    std::string params;   // extra kernel parameters ("const float* in1, float* out0, ")
    std::string args;     // the same names as call arguments ("in1, out0, ")
    std::vector<KernelArg> launch_args;
    double bytes = 0;     // algorithmic bytes of the extra operands and outputs
};

EpilogueCode gen_epilogue(const Graph& g, const Cluster& c, const std::string& kernel, const CodegenOptions& opt) {
    EpilogueCode code;
    std::ostringstream os;
    if (c.epilogue.empty()) {
        os << "__device__ __forceinline__ float4 " << kernel << "_store4(size_t index, float4 r, float* C, unsigned) { *reinterpret_cast<float4*>(C + index) = r; return r; }\n";
        code.store4 = os.str();
        return code;
    }
    const Cluster& p = c.epilogue[0];
    const int product = c.epilogue_product_input;
    std::ostringstream params, args;
    for (size_t i = 0; i < p.inputs.size(); ++i) {
        if ((int)i == product) continue;
        params << "const float* in" << i << ", ";
        args << "in" << i << ", ";
        code.launch_args.push_back({KernelArg::NodeBuffer, p.inputs[i].node_id, 0});
        code.bytes += chain_bytes(g, p.inputs[i]);
    }
    for (size_t i = 0; i < p.outputs.size(); ++i) {
        params << "float* out" << i << ", ";
        args << "out" << i << ", ";
        code.launch_args.push_back({KernelArg::NodeBuffer, p.outputs[i], 0});
        code.bytes += 4.0 * (double)p.element_count;
    }
    code.params = params.str();
    code.args = args.str();
    os << "// epilogue: " << p.label << "\n";
    os << "__device__ __forceinline__ float4 " << kernel << "_store4(size_t index, float4 r, float* C, " << code.params << "unsigned dsc_seed) {\n";
    os << "    (void)C; (void)dsc_seed;\n    const unsigned base = (unsigned)index;\n    const float acc[4] = {r.x, r.y, r.z, r.w};\n";
    std::vector<bool> vector_load(p.inputs.size(), false), is_loaded(p.inputs.size(), false);
    for (const auto& op : p.ops)
        if (op.kind == PerElementOp::Load) is_loaded[op.input_index] = true;
    int uniq = 0;
    for (size_t i = 0; i < p.inputs.size(); ++i) {
        if ((int)i == product || !is_loaded[i]) continue;
        if (p.inputs[i].chain.is_identity()) {
            os << "    const float4 q" << i << " = *reinterpret_cast<const float4*>(in" << i << " + base);\n";
        } else if (chain_vector_run(p.inputs[i].chain, p.inputs[i].arg_shape.at(-1)) == 4) {
            // e.g. the bias, broadcast along every axis but the channels: the four channels of this group are adjacent
            std::string idx = emit_chain(os, p.inputs[i].chain, "base", uniq);
            os << "    const float4 q" << i << " = *reinterpret_cast<const float4*>(in" << i << " + " << idx << ");\n";
        } else {
            continue;
        }
        vector_load[i] = true;
        os << "    const float vin" << i << "[4] = {q" << i << ".x, q" << i << ".y, q" << i << ".z, q" << i << ".w};\n";
    }
    for (size_t i = 0; i < p.outputs.size(); ++i) os << "    float vout" << i << "[4];\n";
    os << "    #pragma unroll\n    for (int v = 0; v < 4; ++v) {\n    const unsigned e = base + v;\n";
    emit_per_element_ops(os, p, opt, uniq, vector_load, product);
    for (size_t i = 0; i < p.outputs.size(); ++i) os << "    vout" << i << "[v] = t" << p.output_ops[i] << ";\n";
    os << "    }\n";
    for (size_t i = 0; i < p.outputs.size(); ++i)
        os << "    *reinterpret_cast<float4*>(out" << i << " + base) = make_float4(vout" << i << "[0], vout" << i << "[1], vout" << i << "[2], vout" << i
           << "[3]);\n";
    os << "    return make_float4(vout0[0], vout0[1], vout0[2], vout0[3]);  // the first output, for kernels that also pool it\n";
    os << "}\n";
    code.store4 = os.str();
    return code;
}

// ---- operand prologues of GEMM kernels ---------------------------------------------------------------
// How a GEMM kernel fetches operand k (0 = A, 1 = B) at element index `idx` of the operand's source array.  Plain
// operands are loaded from memory.  When a PrologueRequest names a per-element producer for the operand (graph.hpp
// OperandPrologue), the source array does not exist: `<kernel>_op<A|B>4(idx, ...)` / `...1(idx, ...)` evaluate the
// producer's program for the four consecutive elements starting at idx / for element idx, loading the producer's own
// inputs instead (128-bit where they are plain arrays or their chains keep aligned groups of four together).
PrologueRequest* g_prologue = nullptr;  // set by generate_cluster_code for the duration of one cluster's generation

bool prologue_requested(int operand) { return g_prologue && g_prologue->producer[operand]; }

struct OperandLoad {
    bool fused = false;
    std::string funcs;    // device functions, emitted ahead of the kernel
    std::string params;   // extra kernel parameters ("const float* pa0, const float* pa1, ")
    std::string call;     // the same names as call arguments (", pa0, pa1")
    std::string ptr;      // "A" / "B"
    std::string kernel;
    std::vector<KernelArg> launch_args;
    std::vector<int> reads;
    double bytes = 0;
    std::string load4(const std::string& idx) const {
        return fused ? kernel + "_op" + ptr + "4((unsigned)(" + idx + ")" + call + ")" : "*reinterpret_cast<const float4*>(" + ptr + " + " + idx + ")";
    }
    std::string load1(const std::string& idx) const {
        return fused ? kernel + "_op" + ptr + "1((unsigned)(" + idx + ")" + call + ")" : ptr + "[" + idx + "]";
    }
};

// Call only when the generator is certain to emit its kernel (after every `return false`): it marks the request fused.
OperandLoad operand_load(const Graph& g, int operand, const std::string& kernel, const ClusterInput& in, const CodegenOptions& opt) {
    OperandLoad o;
    o.ptr = operand == 0 ? "A" : "B";
    o.kernel = kernel;
    o.bytes = chain_bytes(g, in);
    if (!prologue_requested(operand)) return o;
    const Cluster& p = *g_prologue->producer[operand];
    DSC_CHECK(p.outputs.size() == 1 && p.outputs[0] == in.node_id, "operand prologue does not produce this operand");
    g_prologue->fused[operand] = true;
    o.fused = true;
    o.bytes = 0;
    const std::string prefix = operand == 0 ? "pa" : "pb";
    std::ostringstream params, call, fparams;
    std::vector<bool> is_loaded(p.inputs.size(), false);
    for (const auto& op : p.ops)
        if (op.kind == PerElementOp::Load) is_loaded[op.input_index] = true;
    for (size_t i = 0; i < p.inputs.size(); ++i) {
        params << "const float* " << prefix << i << ", ";
        call << ", " << prefix << i;
        fparams << ", const float* in" << i;
        o.launch_args.push_back({KernelArg::NodeBuffer, p.inputs[i].node_id, 0});
        o.reads.push_back(p.inputs[i].node_id);
        if (is_loaded[i]) o.bytes += chain_bytes(g, p.inputs[i]);
    }
    o.params = params.str();
    o.call = call.str();
    std::ostringstream os;
    os << "// operand " << o.ptr << " computed while loading: " << p.label << "\n";
    {   // four consecutive elements
        os << "__device__ __forceinline__ float4 " << kernel << "_op" << o.ptr << "4(unsigned base" << fparams.str() << ") {\n";
        std::vector<bool> vector_load(p.inputs.size(), false);
        int uniq = 0;
        for (size_t i = 0; i < p.inputs.size(); ++i) {
            if (!is_loaded[i]) continue;
            const ViewChain& chain = p.inputs[i].chain;
            if (chain.is_identity()) {
                os << "    const float4 q" << i << " = *reinterpret_cast<const float4*>(in" << i << " + base);\n";
            } else if (chain_vector_run(chain, p.inputs[i].arg_shape.at(-1)) == 4) {
                std::string idx = emit_chain(os, chain, "base", uniq);
                os << "    const float4 q" << i << " = *reinterpret_cast<const float4*>(in" << i << " + " << idx << ");\n";
            } else {
                continue;
            }
            vector_load[i] = true;
            os << "    const float vin" << i << "[4] = {q" << i << ".x, q" << i << ".y, q" << i << ".z, q" << i << ".w};\n";
        }
        os << "    float r[4];\n    #pragma unroll\n    for (int v = 0; v < 4; ++v) {\n    const unsigned e = base + v; (void)e;\n";
        emit_per_element_ops(os, p, opt, uniq, vector_load, -1);
        os << "    r[v] = t" << p.output_ops[0] << ";\n    }\n    return make_float4(r[0], r[1], r[2], r[3]);\n}\n";
    }
    {   // one element
        os << "__device__ __forceinline__ float " << kernel << "_op" << o.ptr << "1(unsigned e" << fparams.str() << ") {\n";
        std::vector<bool> vector_load(p.inputs.size(), false);
        int uniq = 0;
        emit_per_element_ops(os, p, opt, uniq, vector_load, -1);
        os << "    return t" << p.output_ops[0] << ";\n}\n";
    }
    o.funcs = os.str();
    return o;
}

// after a kernel's own arguments: bind the producers' inputs, and give the never-materialised operand a valid address
void bind_operand_loads(KernelLaunch& l, ClusterCode* code, const OperandLoad* a, const OperandLoad* b) {
    const OperandLoad* ops[2] = {a, b};
    for (int k = 0; k < 2; ++k) {
        if (!ops[k] || !ops[k]->fused) continue;
        l.args[k] = ops[k]->launch_args[0];  // the kernel never dereferences A / B itself
        l.args.insert(l.args.end(), ops[k]->launch_args.begin(), ops[k]->launch_args.end());
        code->extra_reads.insert(code->extra_reads.end(), ops[k]->reads.begin(), ops[k]->reads.end());
    }
}

// ---- stride-1 convolutions as halo-tiled implicit GEMMs (halo_conv_template.inc) ---------------------
// measurement hooks (environment variables, read once): CTAs per SM of the halo kernels, ring depth of the pipelined form
// (0 = use the one-tile-at-a-time form)
int halo_ctas_per_sm() {
    static const int v = [] { const char* e = std::getenv("DSC_HALO_CTAS"); return e ? std::atoi(e) : 5; }();
    return v;
}
int halo_pipeline_stages() {
    static const int v = [] { const char* e = std::getenv("DSC_HALO_STAGES"); return e ? std::atoi(e) : 2; }();
    return v;
}
struct HaloConv {
    bool backward_input = false;
    int64_t groups, images, out_h, out_w, filter_h, filter_w;
    int64_t k_per_group;  // reduction channels per tap and group
    int64_t n_per_group;  // output channels per group
    bool rows_mode = true;  // forward: product stored [pixel, group, channel] (Rows) or [group, pixel, channel]
    int64_t unpad_h = 0, unpad_w = 0;  // backward-input: Unpad amounts fused into the epilogue
    bool unpad_rows_first = true;
};

// `a` / `b` are the GEMM operands [group, pixel, k] and [group, k, n] behind their chains.  Returns false when the
// shape falls outside what the kernel covers (the caller then uses the gathered GEMM).
bool gen_halo_conv(const Graph& g, const Cluster& c, int ci, const CodegenOptions& opt, const HaloConv& h, const ClusterInput& a, const ClusterInput& b,
                   ClusterCode* out) {
    const int64_t G = h.groups, OH = h.out_h, OW = h.out_w, FH = h.filter_h, FW = h.filter_w, KG = h.k_per_group, NG = h.n_per_group;
    const int64_t W = OW + FW - 1, PH = OH + FH - 1;      // padded image
    const int64_t rows = h.backward_input ? PH : OH;      // output rows per image
    const int64_t M = a.arg_shape[1], K = a.arg_shape[2], N = b.arg_shape[2];
    if (W > 128 || 128 % W != 0 || KG % 8 != 0 || NG % 4 != 0 || NG > 256) return false;
    if (chain_vector_run_axis(a.chain, a.arg_shape, 2) != 4) return false;
    const int64_t BN = div_round_up(NG, 16) * 16;
    int64_t tmem_cols = 32;
    while (tmem_cols < G * BN) tmem_cols *= 2;
    if (tmem_cols > 512) return false;
    const int64_t TH = 128 / W, halo_rows = TH + FH - 1, Q = G * KG / 4, lead = h.backward_input ? FW - 1 : 0;
    // chunk stride in 16-byte units, mod 8: odd when 8 consecutive lanes store 8 chunks of one pixel, 8/Q when they
    // store Q < 8 chunks of 8/Q pixels
    const int64_t phase = (Q >= 8 || (Q & (Q - 1)) != 0) ? 1 : 8 / Q;
    const int64_t npix = div_round_up(lead + halo_rows * W + FW - 1 - phase, 8) * 8 + phase;
    const int64_t a_bytes = div_round_up(Q * npix * 16, 128) * 128, b_bytes = G * FH * FW * (KG / 4) * BN * 16;
    const bool unpad = h.unpad_h > 0 || h.unpad_w > 0;
    if (unpad) {
        // every padded row / column that folds into an output row / column must sit in the same tile as it
        if (h.unpad_h >= TH || (rows - h.unpad_h - 1) / TH != (rows - 1) / TH || rows - 2 * h.unpad_h < 1 || W - 2 * h.unpad_w < 1) return false;
        if (128 * (G * NG + 4) * 4 > 64 * 1024) return false;
    }
    const bool stage_out = !h.backward_input && h.rows_mode && (G * NG) % 4 == 0 && 128 * (G * NG + 4) * 4 <= 64 * 1024;
    const int64_t a_region = std::max<int64_t>(a_bytes, (unpad || stage_out) ? div_round_up(128 * (G * NG + 4) * 4, 128) * 128 : 0);
    const int64_t smem = a_region + b_bytes + 64 + 128;
    if (smem > 160 * 1024 || halo_rows * W * Q > 256 * 16) return false;  // operands must fit; at most 16 staged loads per thread
    const int64_t tiles = h.images * div_round_up(rows, TH);

    int uniq = 0;
    std::ostringstream ca, cb, a_coords, b_coords, tap_pixel, out_ok, out_index;
    std::string ia = emit_chain(ca, a.chain, {{"batch", M * K, G}, {"gm", K, M}, {"gk", 1, K}}, uniq, "                ");
    std::string ib = emit_chain(cb, b.chain, {{"batch", K * N, G}, {"gk", N, K}, {"gn", 1, N}}, uniq, "        ");
    if (h.backward_input) {
        // halo pixel (py, px) is window position (py - (FH-1), px); A is dY[group, (image, oy, ox), k]
        a_coords << "const int oy = py - " << FH - 1 << "; const bool ok = px < " << OW << " && (unsigned)oy < " << OH << "u; const int gk = kin, gm = (image * "
                 << OH << " + oy) * " << OW << " + px;";
        b_coords << "const int gk = kin, gn = tap * NG + n;";
        tap_pixel << "LEAD + (FH - 1 - fy) * W - fx";
        out_ok << "true";
        out_index << "((size_t)(image * ROWS + y) * W + x) * (G * NG) + g * NG";
    } else {
        // halo pixel (py, px) is padded-image position; any window element that reads it gives its address
        a_coords << "const bool ok = py < " << PH << "; const int fy = max(py - " << OH - 1 << ", 0), fx = max(px - " << OW - 1
                 << ", 0); const int gk = (fy * FW + fx) * KG + kin, gm = (image * " << OH << " + py - fy) * " << OW << " + px - fx;";
        b_coords << "const int gk = tap * KG + kin, gn = n;";
        tap_pixel << "fy * W + fx";
        out_ok << "x < " << OW;
        if (h.rows_mode) out_index << "((size_t)((image * ROWS + y) * " << OW << " + x) * G + g) * NG";
        else out_index << "((size_t)g * " << M << " + (image * ROWS + y) * " << OW << " + x) * NG";
    }
    const std::string name = "k" + num(ci);
    // CTAs per SM: shared memory, TMEM columns and registers all bound it.  The kernel needs ~50 registers; asking the
    // compiler for 5 CTAs (48 registers) keeps the persistent grid a single wave (7 were assumed from shared memory alone in
    // round 1 while the register file admitted 4: 1.75 waves, a quarter of the SM time idle in the tail).
    const int64_t resident = std::max<int64_t>(1, std::min<int64_t>({(int64_t)halo_ctas_per_sm(), (200 * 1024) / smem, 512 / tmem_cols}));
    if (!c.epilogue.empty() && (h.backward_input || !h.rows_mode || G * NG % 4 != 0)) return false;
    if (prologue_requested(1)) return false;  // the weights are staged once per CTA, element by element: no producer there
    const EpilogueCode epi = gen_epilogue(g, c, name, opt);
    const int64_t out_pixels = h.images * (rows - 2 * h.unpad_h) * ((h.backward_input ? W : OW) - 2 * h.unpad_w);

    // The software-pipelined form (halo_conv_pipelined_template.inc) addresses the operand through per-axis tables:
    // evaluate the chain for every (row, column, channel chunk) of the halo grid and check that the source index is
    // image * stride + ytab[py] + xtab[px] + ctab[q] -- true whenever padding, windows and group slices act per axis.
    // Measured on B200 (conv-net m = 8192, profiles/r2_halo_pipeline.md): forward 110 -> 103 us with a 2-deep ring at 4 CTAs per
    // SM.  Backward-input is paced by the tensor core's operand reads from shared memory (36 MMAs x 4 KB of A per 128 pixels:
    // the im2col expansion is read from shared memory whichever way the loads arrive), so it only loses the CTAs the ring's
    // shared memory costs (133 us at 5 CTAs with the one-tile form; 144 / 170 / 291 us at 3 / 2 / 1 CTAs with the ring):
    // it keeps the first form unless DSC_HALO_PIPELINE_BACKWARD=1.
    static const bool pipeline_backward = [] { const char* e = std::getenv("DSC_HALO_PIPELINE_BACKWARD"); return e && std::atoi(e) != 0; }();
    const int64_t stages = halo_pipeline_stages();
    if (stages >= 2 && !prologue_requested(0) && (!h.backward_input || pipeline_backward)) {
        const int64_t tiles_per_image = div_round_up(rows, TH), YT = tiles_per_image * TH + FH - 1;
        auto src_index = [&](int64_t image, int64_t py, int64_t px, int64_t q) -> int64_t {
            const int64_t batch = (q * 4) / KG, kin = (q * 4) % KG;
            int64_t gm, gk;
            if (h.backward_input) {
                const int64_t oy = py - (FH - 1);
                if (!(px < OW && oy >= 0 && oy < OH)) return -1;
                gk = kin;
                gm = (image * OH + oy) * OW + px;
            } else {
                if (py >= PH) return -1;
                const int64_t fy = std::max<int64_t>(py - (OH - 1), 0), fx = std::max<int64_t>(px - (OW - 1), 0);
                gk = (fy * FW + fx) * KG + kin;
                gm = (image * OH + py - fy) * OW + px - fx;
            }
            return eval_chain(a.chain, batch * M * K + gm * K + gk);
        };
        const int64_t py0 = h.backward_input ? FH - 1 : 0;
        const int64_t origin = src_index(0, py0, 0, 0);
        const int64_t image_stride = h.images > 1 ? src_index(1, py0, 0, 0) - origin : 0;
        std::vector<int64_t> ytab(YT), xtab(W), ctab(Q);
        for (int64_t py = 0; py < YT; ++py) ytab[py] = src_index(0, py, 0, 0);
        for (int64_t px = 0; px < W; ++px) { const int64_t v = src_index(0, py0, px, 0); xtab[px] = v < 0 ? -1 : v - origin; }
        for (int64_t q = 0; q < Q; ++q) ctab[q] = src_index(0, py0, 0, q) - origin;
        bool separable = origin >= 0;
        for (int64_t image : {(int64_t)0, std::min<int64_t>(1, h.images - 1), h.images - 1})
            for (int64_t py = 0; py < YT && separable; ++py)
                for (int64_t px = 0; px < W && separable; ++px)
                    for (int64_t q = 0; q < Q; ++q) {
                        const int64_t want = src_index(image, py, px, q);
                        const int64_t got = (ytab[py] < 0 || xtab[px] < 0) ? -1 : image * image_stride + ytab[py] + xtab[px] + ctab[q];
                        if (want != got || (want >= 0 && want % 4 != 0)) { separable = false; break; }
                    }
        int64_t acc_cols = 32;
        while (acc_cols < G * BN) acc_cols *= 2;
        const int64_t a_stage = div_round_up(Q * npix * 16, 128) * 128;
        const int64_t so_bytes = (unpad || stage_out) ? 128 * (G * NG + 4) * 4 : 0;
        const int64_t smem2 = stages * a_stage + b_bytes + so_bytes + 64 + YT * 4 + 128;
        const auto& pool = c.pool;
        // the tile must hold whole pooling windows of the activation it stages (forward, channels interleaved per pixel)
        const bool pooled = pool.enabled && !h.backward_input && stage_out && pool.images == h.images && pool.height == OH && pool.width == OW &&
                            pool.channels == G * NG && TH % pool.window_h == 0 && OH % pool.window_h == 0 && OW % pool.window_w == 0;
        if (separable && 2 * acc_cols <= 512 && smem2 <= 200 * 1024) {
            const int64_t resident2 = std::max<int64_t>(1, std::min<int64_t>({(int64_t)halo_ctas_per_sm(), (220 * 1024) / (smem2 + 1024), 512 / (2 * acc_cols)}));
            auto list = [](const std::vector<int64_t>& v) {
                std::string t;
                for (size_t i = 0; i < v.size(); ++i) t += (i ? ", " : "") + num(v[i]);
                return t;
            };
            out->pool_done = pooled;
            out->source = subst(kHaloConvPipelinedTemplate,
                                {{"LABEL", c.label}, {"POOL", pooled ? "true" : "false"}, {"POOL_H", num(pooled ? pool.window_h : 1)}, {"POOL_W", num(pooled ? pool.window_w : 1)},
                                 {"POOL_PARAMS", pooled ? "float* pool_out, " : ""}, {"POOL_DECL", pooled ? "" : "float* const pool_out = nullptr;"}, {"MIN_CTAS", num(resident2)}, {"NAME", name}, {"STORE4", epi.store4}, {"EPI_PARAMS", epi.params}, {"EPI_ARGS", epi.args},
                                 {"G", num(G)}, {"IMAGES", num(h.images)}, {"ROWS", num(rows)}, {"W", num(W)}, {"FH", num(FH)}, {"FW", num(FW)}, {"KG", num(KG)},
                                 {"NG", num(NG)}, {"BN", num(BN)}, {"LEAD", num(lead)}, {"TMEM_COLS", num(2 * acc_cols)}, {"NPIX", num(npix)}, {"STAGES", num(stages)},
                                 {"PY", num(h.unpad_h)}, {"PX", num(h.unpad_w)}, {"STAGE_OUT", stage_out ? "true" : "false"}, {"OUT_W", num(OW)},
                                 {"ROWS_FIRST", h.unpad_rows_first ? "true" : "false"}, {"YT", num(YT)}, {"YTAB", list(ytab)}, {"XTAB", list(xtab)}, {"CTAB", list(ctab)},
                                 {"IMG_STRIDE", num(image_stride)}, {"B_COORDS", b_coords.str()}, {"TAP_PIXEL", tap_pixel.str()}, {"OUT_OK", out_ok.str()},
                                 {"OUT_INDEX", out_index.str()}, {"B_CHAIN", cb.str()}, {"B_IDX", ib}});
            KernelLaunch l;
            l.entry = name;
            l.grid_x = (uint32_t)std::min<int64_t>(tiles, (int64_t)opt.sm_count * resident2);  // persistent over tiles: exactly one wave
            l.block = 256;
            l.smem = (uint32_t)smem2;
            l.label = "TensorCore" + c.label;
            l.cluster = ci;
            l.args = {{KernelArg::NodeBuffer, a.node_id, 0}, {KernelArg::NodeBuffer, b.node_id, 0}, {KernelArg::NodeBuffer, c.outputs[0], 0}};
            l.args.insert(l.args.end(), epi.launch_args.begin(), epi.launch_args.end());
            if (pooled) l.args.push_back({KernelArg::NodeBuffer, c.outputs.back(), 0});
            l.algorithmic_bytes = chain_bytes(g, a) + chain_bytes(g, b) + (c.epilogue.empty() ? 4.0 * (double)(out_pixels * G * NG) : epi.bytes) +
                                  (pooled ? 4.0 * (double)g.ops().nodes[c.outputs.back()].shape.element_count() : 0.0);
            l.flops = 2.0 * (double)(out_pixels * G * NG) * (double)(FH * FW * KG);
            out->launches.push_back(l);
            return true;
        }
    }
    const OperandLoad la = operand_load(g, 0, name, a, opt);
    out->source = subst(kHaloConvTemplate,
                        {{"LABEL", c.label + (la.fused ? "  [A = " + g_prologue->producer[0]->label + "]" : "")}, {"PRO_FUNCS", la.funcs}, {"PRO_PARAMS", la.params}, {"A_LOAD4", la.load4(ia)}, {"MIN_CTAS", num(resident)}, {"NAME", name}, {"STORE4", epi.store4}, {"EPI_PARAMS", epi.params}, {"EPI_ARGS", epi.args}, {"G", num(G)}, {"IMAGES", num(h.images)}, {"ROWS", num(rows)}, {"W", num(W)}, {"FH", num(FH)},
                         {"FW", num(FW)}, {"KG", num(KG)}, {"NG", num(NG)}, {"BN", num(BN)}, {"LEAD", num(lead)}, {"TMEM_COLS", num(tmem_cols)}, {"NPIX", num(npix)},
                         {"PY", num(h.unpad_h)}, {"PX", num(h.unpad_w)}, {"STAGE_OUT", stage_out ? "true" : "false"}, {"OUT_W", num(OW)},
                         // loading the next halo before the drain pays for the forward kernel (0.125 -> 0.117 ms) and costs the
                         // backward one, whose epilogue is the longer Unpad pass (0.136 -> 0.175 ms): measured, conv-net m=8192
                         {"PREFETCH", h.backward_input ? "false" : "true"}, {"ROWS_FIRST", h.unpad_rows_first ? "true" : "false"},
                         {"A_COORDS", a_coords.str()}, {"B_COORDS", b_coords.str()}, {"TAP_PIXEL", tap_pixel.str()}, {"OUT_OK", out_ok.str()},
                         {"OUT_INDEX", out_index.str()}, {"A_CHAIN", ca.str()}, {"B_CHAIN", cb.str()}, {"B_IDX", ib}});
    KernelLaunch l;
    l.entry = name;
    l.grid_x = (uint32_t)std::min<int64_t>(tiles, (int64_t)opt.sm_count * resident);  // persistent over tiles: exactly one wave
    l.block = 256;
    l.smem = (uint32_t)smem;
    l.label = "TensorCore" + c.label;
    l.cluster = ci;
    l.args = {{KernelArg::NodeBuffer, a.node_id, 0}, {KernelArg::NodeBuffer, b.node_id, 0}, {KernelArg::NodeBuffer, c.outputs[0], 0}};
    l.args.insert(l.args.end(), epi.launch_args.begin(), epi.launch_args.end());
    bind_operand_loads(l, out, &la, nullptr);
    l.algorithmic_bytes = la.bytes + chain_bytes(g, b) + (c.epilogue.empty() ? 4.0 * (double)(out_pixels * G * NG) : epi.bytes);
    l.flops = 2.0 * (double)(out_pixels * G * NG) * (double)(FH * FW * KG);
    out->launches.push_back(l);
    return true;
}

bool gen_conv_backward_input(const Graph& g, const Cluster& c, int ci, const CodegenOptions& opt, ClusterCode* out) {
    const auto& cbi = c.conv_backward_input;
    const ClusterInput& a = cbi.unfused[0];
    const ClusterInput& b = cbi.unfused[1];
    if (cbi.in_w != cbi.out_w + cbi.filter_w - 1 || cbi.in_h != cbi.out_h + cbi.filter_h - 1) return false;
    HaloConv h;
    h.backward_input = true;
    h.groups = a.arg_shape[0];
    h.out_h = cbi.out_h; h.out_w = cbi.out_w; h.filter_h = cbi.filter_h; h.filter_w = cbi.filter_w;
    h.images = a.arg_shape[1] / (cbi.out_h * cbi.out_w);
    h.k_per_group = a.arg_shape[2];
    h.n_per_group = b.arg_shape[2] / (cbi.filter_h * cbi.filter_w);
    h.unpad_h = cbi.unpad_h;
    h.unpad_w = cbi.unpad_w;
    h.unpad_rows_first = cbi.unpad_first_axis != 2;
    return gen_halo_conv(g, c, ci, opt, h, a, b, out);
}

// Forward conv2d: the A operand is image_to_windows of the (padded) image, i.e. its chain ends in a view with
// output [image, oy, ox, group, fy, fx, c] in which oy and fy walk one input axis with step 1, ox and fx
// another, and group / c the channel axis (array.rs image_to_windows; stride 1), optionally followed by the
// [pixel, group, k] -> [group, pixel, k] transposition.  Then A[g, (image, oy, ox), (fy, fx, c)] depends on
// (oy + fy, ox + fx) only, which is what the halo kernel needs.
bool gen_conv_forward(const Graph& g, const Cluster& c, int ci, const CodegenOptions& opt, ClusterCode* out) {
    const OpNode& mm = g.ops().nodes[c.node_id];
    const ClusterInput& a = c.inputs[0];
    const ClusterInput& b = c.inputs[1];
    const int64_t G = a.arg_shape[0], M = a.arg_shape[1], K = a.arg_shape[2], N = b.arg_shape[2];
    if (mm.shape[0] != 1 || c.matmul_absorbs_reduce || a.chain.views.empty()) return false;
    size_t wi = a.chain.views.size() - 1;
    if (G > 1) {  // the transposition [M, G, K] -> [G, M, K]
        const View& p = a.chain.views[wi];
        if (wi == 0 || p.input_shape != Shape({M, G, K}) || p.output_shape != Shape({G, M, K}) || p.any_clamp()) return false;
        const AxisMapping want[3] = {AxisMapping::identity(1, G), AxisMapping::identity(0, M), AxisMapping::identity(2, K)};
        for (int i = 0; i < 3; ++i)
            if ((p.output_shape[i] > 1 && !(p.output_mapping[i] == want[i])) || p.input_offsets[i] != 0) return false;
        --wi;
    }
    const View& w = a.chain.views[wi];
    if (w.output_shape.len() != 7 || w.input_shape.len() != 4) return false;
    const int64_t B = w.output_shape[0], OH = w.output_shape[1], OW = w.output_shape[2], FH = w.output_shape[4], FW = w.output_shape[5], C = w.output_shape[6];
    if (w.output_shape[3] != G || B * OH * OW != M || FH * FW * C != K) return false;
    auto maps = [&](int out_axis, int in_axis, int64_t step) {
        const AxisMapping& m = w.output_mapping[out_axis];
        if (w.output_shape[out_axis] == 1) return true;  // never needs a coordinate
        return m.is_source && m.axis == in_axis && m.step == step;
    };
    if (!maps(0, 0, 1) || !maps(1, 1, 1) || !maps(2, 2, 1) || !maps(3, 3, C) || !maps(4, 1, 1) || !maps(5, 2, 1) || !maps(6, 3, 1)) return false;
    HaloConv h;
    h.backward_input = false;
    h.groups = G; h.images = B; h.out_h = OH; h.out_w = OW; h.filter_h = FH; h.filter_w = FW;
    h.k_per_group = C;
    h.n_per_group = N;
    h.rows_mode = mm.op.output_mode == MatMulOutputMode::Rows;
    return gen_halo_conv(g, c, ci, opt, h, a, b, out);
}

// conv2d weight gradient: A = transposed window matrix [group, (fy, fx, c), pixel], B = dY [group, pixel, co], the
// Reduce over the k split absorbed (halo_wgrad_template.inc).  Same recognition of the window view as the forward.
bool gen_conv_weight_gradient(const Graph& g, const Cluster& c, int ci, const CodegenOptions& opt, ClusterCode* out) {
    const OpNode& mm = g.ops().nodes[c.node_id];
    const ClusterInput& a = c.inputs[0];
    const ClusterInput& b = c.inputs[1];
    const int64_t G = a.arg_shape[0], KW = a.arg_shape[1], MPIX = a.arg_shape[2], NCO = b.arg_shape[2];
    if (!(c.matmul_absorbs_reduce || mm.shape[0] == 1) || mm.op.output_mode == MatMulOutputMode::Rows || a.chain.views.size() < 2) return false;
    const View& p = a.chain.views.back();  // [pixel, group, (fy, fx, c)] -> [group, (fy, fx, c), pixel]
    if (p.output_shape != Shape({G, KW, MPIX}) || p.any_clamp()) return false;
    int src_axis[3];  // input axis that carries group / (fy, fx, c) / pixel
    if (p.input_shape == Shape({MPIX, G, KW})) { src_axis[0] = 1; src_axis[1] = 2; src_axis[2] = 0; }
    else if (G == 1 && p.input_shape == Shape({G, MPIX, KW})) { src_axis[0] = 0; src_axis[1] = 2; src_axis[2] = 1; }  // same order when there is one group
    else return false;
    for (int i = 0; i < 3; ++i) {
        const AxisMapping want = AxisMapping::source(src_axis[i], 1);
        if ((p.output_shape[i] > 1 && !(p.output_mapping[i] == want)) || p.input_offsets[i] != 0) return false;
    }
    const View& w = a.chain.views[a.chain.views.size() - 2];
    if (w.output_shape.len() != 7 || w.input_shape.len() != 4) return false;
    const int64_t B = w.output_shape[0], OH = w.output_shape[1], OW = w.output_shape[2], FH = w.output_shape[4], FW = w.output_shape[5], C = w.output_shape[6];
    if (w.output_shape[3] != G || B * OH * OW != MPIX || FH * FW * C != KW) return false;
    auto maps = [&](int out_axis, int in_axis, int64_t step) {
        const AxisMapping& m = w.output_mapping[out_axis];
        if (w.output_shape[out_axis] == 1) return true;
        return m.is_source && m.axis == in_axis && m.step == step;
    };
    if (!maps(0, 0, 1) || !maps(1, 1, 1) || !maps(2, 2, 1) || !maps(3, 3, C) || !maps(4, 1, 1) || !maps(5, 2, 1) || !maps(6, 3, 1)) return false;
    const int64_t W = OW + FW - 1;
    if (W > 128 || 128 % W != 0 || C % 4 != 0 || NCO % 4 != 0 || FW * G * C > 128 || (G * NCO) % 16 != 0 || G * NCO > 256) return false;
    if (chain_vector_run_axis(a.chain, a.arg_shape, 1) != 4 || chain_vector_run_axis(b.chain, b.arg_shape, 2) != 4) return false;
    int64_t tmem_cols = 32;
    while (tmem_cols < FH * G * NCO) tmem_cols *= 2;
    if (tmem_cols > 512) return false;
    const int64_t TH = 128 / W, halo_rows = TH + FH - 1;
    if (FW - 1 > 4 || W % 4 != 0) return false;
    const int64_t npix = div_round_up(4 + halo_rows * W, 4) * 4;  // mirrors NPIX / X_BLOCK / X_BYTES of the template
    const int64_t x_block = npix * 128, m_blocks = div_round_up(FW * G * C, 32);
    // M = 64 would fit 48 real rows, but measured slower on B200 (0.227 vs 0.200 ms, conv-net m=8192): the MMA is paced
    // as if M were 128 either way and the third resident CTA does not pay for itself
    const int64_t m_rows = 128;
    const bool colsum = !c.column_sum.empty() && 256 % (G * NCO / 4) == 0;  // dY is summed on the side (bias gradient), exact FP32
    // one wide MMA per k-step when all filter rows fit in N <= 256 and the FH dY copies fit in shared memory
    // (measured on conv-net m=8192: 0.201 -> 0.152 ms)
    bool wide = FH * G * NCO <= 256;
    auto smem_bytes = [&](bool w) {
        const int64_t y_block = (w ? halo_rows * W : 128) * 128, nb = div_round_up(w ? FH * G * NCO : G * NCO, 32);
        return std::max<int64_t>((m_rows / 32) * x_block, m_blocks * x_block + nb * y_block);
    };
    if (wide && (smem_bytes(true) + 64 + 1024 > 110 * 1024 || (halo_rows * W * 128) / 16 > 0x3fff)) wide = false;  // keep two CTAs per SM
    const int64_t x_bytes = smem_bytes(wide);
    if (x_block / 16 > 0x3fff || x_bytes + 64 + 1024 > 200 * 1024) return false;
    if (halo_rows * W * (G * C / 4) > 256 * 12 || 128 * (G * NCO / 4) > 256 * 12) return false;  // staged loads per thread
    const int64_t smem = x_bytes + 64 + 1024;
    const int64_t tiles = B * div_round_up(OH, TH);
    const int64_t resident = std::max<int64_t>(1, std::min<int64_t>({4, (224 * 1024) / (smem + 1024), 512 / tmem_cols}));  // 227 KB per SM, 1 KB reserved per CTA
    const int64_t S = std::min<int64_t>(tiles, (int64_t)opt.sm_count * resident);

    int uniq = 0;
    std::ostringstream ca, cb;
    std::string ia = emit_chain(ca, a.chain, {{"batch", KW * MPIX, G}, {"gm", MPIX, KW}, {"gk", 1, MPIX}}, uniq, "                ");
    std::string ib = emit_chain(cb, b.chain, {{"batch", MPIX * NCO, G}, {"gk", NCO, MPIX}, {"gn", 1, NCO}}, uniq, "                ");
    const std::string name = "k" + num(ci);
    const OperandLoad la = operand_load(g, 0, name, a, opt), lb = operand_load(g, 1, name, b, opt);
    out->source = subst(kHaloWgradTemplate, {{"LABEL", c.label + (lb.fused ? "  [B = " + g_prologue->producer[1]->label + "]" : "")}, {"NAME", name},
                                             {"PRO_FUNCS", la.funcs + lb.funcs}, {"PRO_PARAMS", la.params + lb.params}, {"A_LOAD4", la.load4(ia)}, {"B_LOAD4", lb.load4(ib)}, {"G", num(G)}, {"IMAGES", num(B)}, {"OH", num(OH)}, {"OW", num(OW)},
                                             {"FH", num(FH)}, {"FW", num(FW)}, {"CG", num(C)}, {"NCO", num(NCO)}, {"TMEM_COLS", num(tmem_cols)}, {"MROWS", num(m_rows)}, {"WIDE", wide ? "true" : "false"},
                                             {"COLSUM", colsum ? "true" : "false"},
                                             {"A_CHAIN", ca.str()}, {"B_CHAIN", cb.str()}});
    const int64_t out_count = G * KW * NCO;
    const int64_t colsum_offset = div_round_up(S * out_count * 4, 256) * 256;
    KernelLaunch l;
    l.entry = name;
    l.grid_x = (uint32_t)S;
    l.block = 256;
    l.smem = (uint32_t)smem;
    l.label = "TensorCore" + c.label;
    l.cluster = ci;
    l.args = {{KernelArg::NodeBuffer, a.node_id, 0}, {KernelArg::NodeBuffer, b.node_id, 0}, {KernelArg::Scratch, -1, 0},
              {KernelArg::Scratch, -1, colsum ? colsum_offset : 0}};
    bind_operand_loads(l, out, &la, &lb);
    l.algorithmic_bytes = la.bytes + lb.bytes + 4.0 * (double)out_count;
    l.flops = 2.0 * (double)G * (double)KW * (double)NCO * (double)MPIX;
    out->launches.push_back(l);
    out->scratch_bytes = S * out_count * 4;
    if (colsum) {
        out->column_sum_done = true;
        out->scratch_bytes = colsum_offset + S * G * NCO * 4;
    }
    emit_split_sums(out, name, ci, c.label, S, out_count, c.outputs[0], colsum, G * NCO, colsum_offset, colsum ? c.outputs[1] : -1);
    return true;
}

// GEMMs with one tiny extent stream their big operand once (thin_gemm_template.inc); strict FP32 either way
bool gen_thin_matmul(const Graph& g, const Cluster& c, int ci, const CodegenOptions& opt, ClusterCode* out) {
    const OpNode& mm = g.ops().nodes[c.node_id];
    const ClusterInput& a = c.inputs[0];
    const ClusterInput& b = c.inputs[1];
    const int64_t BC = a.arg_shape[0], M = a.arg_shape[1], K = a.arg_shape[2], N = b.arg_shape[2];
    if (c.conv_backward_input.enabled || !(c.matmul_absorbs_reduce || mm.shape[0] == 1)) return false;
    const bool rows_mode = mm.op.output_mode == MatMulOutputMode::Rows;
    const std::string c_row = rows_mode ? "(((size_t)split * M + gm) * BC + batch) * N" : "(((size_t)split * BC + batch) * M + gm) * N";
    const std::string name = "k" + num(ci);
    int uniq = 0;
    std::ostringstream ca, cb;
    KernelLaunch l;
    l.entry = name;
    l.label = c.label;
    l.cluster = ci;
    l.grid_y = (uint32_t)BC;
    l.algorithmic_bytes = chain_bytes(g, a) + chain_bytes(g, b) + 4.0 * (double)(BC * M * N);
    l.flops = 2.0 * (double)BC * (double)M * (double)N * (double)K;
    l.args = {{KernelArg::NodeBuffer, a.node_id, 0}, {KernelArg::NodeBuffer, b.node_id, 0}};
    if (K <= 32 && N <= 32 && M >= 65536 && !prologue_requested(0) && !prologue_requested(1)) {
        std::string ia = emit_chain(ca, a.chain, {{"batch", M * K, BC}, {"gm", K, M}, {"gk", 1, K}}, uniq, "            ");
        std::string ib = emit_chain(cb, b.chain, {{"batch", K * N, BC}, {"gk", N, K}, {"gn", 1, N}}, uniq, "            ");
        if (!c.epilogue.empty() && (N % 4 != 0 || !(rows_mode ? true : BC == 1))) return false;
        const EpilogueCode epi = gen_epilogue(g, c, name, opt);
        const auto& pool = c.pool;
        // rows are the pixels of an NHWC activation that is max-pooled right away: one pooling window x 4 channels per thread
        const bool pooled = pool.enabled && BC == 1 && N % 4 == 0 && pool.channels == N && pool.images * pool.height * pool.width == M &&
                            pool.window_h * pool.window_w <= 8 && pool.width % pool.window_w == 0 && pool.height % pool.window_h == 0 && !c.epilogue.empty();
        out->pool_done = pooled;
        l.algorithmic_bytes = chain_bytes(g, a) + chain_bytes(g, b) + (c.epilogue.empty() ? 4.0 * (double)(BC * M * N) : epi.bytes) +
                              (pooled ? 4.0 * (double)g.ops().nodes[c.outputs.back()].shape.element_count() : 0.0);
        // the pooled kernel's input patch: is A a stride-1 window view over a (padded) image with per-axis address tables?
        bool patch = false;
        int64_t fh = 1, fw = 1, cin = 1, img_stride = 0, c_stride = 0;
        std::vector<int64_t> ytab, xtab;
        if (pooled) {
            const int64_t H = pool.height, W = pool.width;
            for (cin = 1; cin <= K && !patch; ++cin) {
                if (K % cin != 0) continue;
                for (fh = 1; fh <= K / cin && !patch; ++fh) {
                    if ((K / cin) % fh != 0) continue;
                    fw = K / cin / fh;
                    if (fh * fw < 2 || (pool.window_h + fh - 1) * (pool.window_w + fw - 1) * cin > 64) continue;
                    auto index = [&](int64_t image, int64_t y, int64_t x, int64_t fy, int64_t fx, int64_t ch) {
                        return eval_chain(a.chain, ((image * H + y) * W + x) * K + (fy * fw + fx) * cin + ch);
                    };
                    const int64_t origin = index(0, 0, 0, 0, 0, 0);
                    img_stride = pool.images > 1 ? index(1, 0, 0, 0, 0, 0) - origin : 0;
                    c_stride = cin > 1 ? index(0, 0, 0, 0, 0, 1) - origin : 0;
                    ytab.assign(H + fh - 1, 0);
                    xtab.assign(W + fw - 1, 0);
                    for (int64_t t = 0; t < H + fh - 1; ++t) ytab[t] = index(0, std::min(t, H - 1), 0, t - std::min(t, H - 1), 0, 0);
                    for (int64_t t = 0; t < W + fw - 1; ++t) xtab[t] = index(0, 0, std::min(t, W - 1), 0, t - std::min(t, W - 1), 0) - origin;
                    bool ok = true;
                    for (int64_t image : {(int64_t)0, pool.images - 1})
                        for (int64_t y = 0; y < H && ok; ++y)
                            for (int64_t x = 0; x < W && ok; ++x)
                                for (int64_t gk = 0; gk < K; ++gk) {
                                    const int64_t fy = gk / (fw * cin), fx = (gk / cin) % fw, ch = gk % cin;
                                    if (index(image, y, x, fy, fx, ch) != image * img_stride + ytab[y + fy] + xtab[x + fx] + ch * c_stride) { ok = false; break; }
                                }
                    if (ok) { patch = true; break; }
                }
                if (patch) break;
            }
        }
        auto list = [](const std::vector<int64_t>& v) {
            std::string t;
            for (size_t i = 0; i < v.size(); ++i) t += (i ? ", " : "") + num(v[i]);
            return t.empty() ? std::string("0") : t;
        };
        if (pooled)
            out->source = subst(kThinRowsPoolTemplate, {{"LABEL", c.label}, {"NAME", name}, {"M", num(M)}, {"N", num(N)}, {"K", num(K)},
                                                       {"PATCH", patch ? "true" : "false"}, {"FH", num(patch ? fh : 1)}, {"FW", num(patch ? fw : 1)}, {"CIN", num(patch ? cin : 1)},
                                                       {"IMG_H", num(pool.height)}, {"YTAB", patch ? list(ytab) : list(std::vector<int64_t>(pool.height, 0))},
                                                       {"XTAB", patch ? list(xtab) : list(std::vector<int64_t>(pool.width, 0))}, {"IMG_STRIDE", num(img_stride)}, {"C_STRIDE", num(c_stride)},
                                                       {"POOL_H", num(pool.window_h)}, {"POOL_W", num(pool.window_w)}, {"IMG_W", num(pool.width)},
                                                       {"STORE4", epi.store4}, {"EPI_PARAMS", epi.params}, {"EPI_ARGS", epi.args},
                                                       {"A_CHAIN", ca.str()}, {"A_IDX", ia}, {"B_CHAIN", cb.str()}, {"B_IDX", ib}});
        else
        out->source = subst(kThinRowsTemplate, {{"LABEL", c.label}, {"NAME", name}, {"M", num(M)}, {"N", num(N)}, {"K", num(K)}, {"BC", num(BC)},
                                               {"STORE4", epi.store4}, {"EPI_PARAMS", epi.params}, {"EPI_ARGS", epi.args},
                                               {"A_CHAIN", ca.str()}, {"A_IDX", ia}, {"B_CHAIN", cb.str()}, {"B_IDX", ib}, {"C_ROW", c_row}});
        const int64_t tile = 256;
        if (pooled) {
            const int64_t windows = M / (pool.window_h * pool.window_w);
            l.grid_x = (uint32_t)div_round_up(windows * (N / 4), 256);
            l.args.push_back({KernelArg::NodeBuffer, c.outputs[0], 0});
            l.args.insert(l.args.end(), epi.launch_args.begin(), epi.launch_args.end());
            l.args.push_back({KernelArg::NodeBuffer, c.outputs.back(), 0});
            out->launches.push_back(l);
            return true;
        }
        l.grid_x = (uint32_t)div_round_up(M, tile);
        l.args.push_back({KernelArg::NodeBuffer, c.outputs[0], 0});
        l.args.insert(l.args.end(), epi.launch_args.begin(), epi.launch_args.end());
        out->launches.push_back(l);
        return true;
    }
    if (M <= 16 && M * N <= 288 && K >= 65536 && c.epilogue.empty() && !prologue_requested(0)) {
        int64_t nsplit = 1;
        // at most 80 accumulators per thread; splitting further re-reads A and measured slower (0.133 -> 0.234 ms at 36)
        while (M * (N / nsplit) > 80 && nsplit < 32 && (N / nsplit) % 2 == 0) nsplit *= 2;
        if (N % nsplit != 0 || M * (N / nsplit) > 96) return false;
        const int64_t nth = N / nsplit;
        const bool b_vec = nth % 4 == 0 && chain_vector_run_axis(b.chain, b.arg_shape, 2) == 4;
        std::string ia = emit_chain(ca, a.chain, {{"batch", M * K, BC}, {"gm", K, M}, {"gk", 1, K}}, uniq, "            ");
        std::string ib = emit_chain(cb, b.chain, {{"batch", K * N, BC}, {"gk", N, K}, {"gn", 1, N}}, uniq, "                ");
        const bool colsum = !c.column_sum.empty();
        const OperandLoad lb = operand_load(g, 1, name, b, opt);
        out->source = subst(kThinReduceTemplate, {{"LABEL", c.label + (lb.fused ? "  [B = " + g_prologue->producer[1]->label + "]" : "")}, {"NAME", name}, {"M", num(M)}, {"N", num(N)}, {"K", num(K)}, {"BC", num(BC)},
                                                 {"COLSUM", colsum ? "true" : "false"}, {"PRO_FUNCS", lb.funcs}, {"PRO_PARAMS", lb.params},
                                                 {"B_LOAD4", lb.load4(ib)}, {"B_LOAD1", lb.load1(ib)},
                                                 {"NSPLIT", num(nsplit)}, {"B_VEC", b_vec ? "true" : "false"}, {"A_CHAIN", ca.str()}, {"A_IDX", ia},
                                                 {"B_CHAIN", cb.str()}, {"C_ROW", c_row}});
        l.algorithmic_bytes = chain_bytes(g, a) + lb.bytes + 4.0 * (double)(BC * M * N);
        const int64_t rows_per_cta = 256 / nsplit;
        const int64_t S = std::max<int64_t>(1, std::min<int64_t>((int64_t)opt.sm_count * 2 / BC, K / (rows_per_cta * 8)));
        l.grid_x = (uint32_t)S;
        const int64_t out_count = BC * M * N;
        // column sums: [S, BC, N] partials behind the product's partials, or straight into outputs[1]
        const int64_t colsum_offset = div_round_up(S * out_count * 4, 256) * 256;
        const KernelArg colsum_arg = !colsum ? KernelArg{KernelArg::NodeBuffer, c.outputs[0], 0}  // unused by the kernel
                                     : S > 1 ? KernelArg{KernelArg::Scratch, -1, colsum_offset} : KernelArg{KernelArg::NodeBuffer, c.outputs[1], 0};
        out->column_sum_done = colsum;
        if (colsum) l.algorithmic_bytes += 4.0 * (double)(BC * N);
        if (S > 1) {
            l.args.push_back({KernelArg::Scratch, -1, 0});
            l.args.push_back(colsum_arg);
            bind_operand_loads(l, out, nullptr, &lb);
            out->launches.push_back(l);
            out->scratch_bytes = S * out_count * 4;
            if (colsum) out->scratch_bytes = colsum_offset + S * BC * N * 4;
            emit_split_sums(out, name, ci, c.label, S, out_count, c.outputs[0], colsum, BC * N, colsum_offset, colsum ? c.outputs[1] : -1);
        } else {
            l.args.push_back({KernelArg::NodeBuffer, c.outputs[0], 0});
            l.args.push_back(colsum_arg);
            bind_operand_loads(l, out, nullptr, &lb);
            out->launches.push_back(l);
        }
        return true;
    }
    return false;
}

extern const char* kUnpadTemplate;

// Strided conv2d backward-input as a gather (graph.hpp ConvBackwardInput::strided): one thread per element of the FINAL
// image gradient [image, y, x, channel] -- after the Unpads, the adjoint of replicate padding (kernel.rs:644-710: border
// pixels also collect the pad rows / columns beyond them) -- sums, over the padded positions that fold onto it and the
// filter taps (fy, fx) for which (position - tap) is a multiple of the stride inside the window grid (the col2im range
// check), the product dY[window, k] * W[k, (tap, channel)].  Replaces MatMul (k=1) + WindowsToImage + two
// Unpads of MaxBlurPool2D's blur convolution (conv-blur-net m = 8192: 877 + 741 + 366 + 350 us and 460 + 995 + 197 + 180 us).
bool gen_conv_backward_gather(const Graph& g, const Cluster& c, int ci, ClusterCode* out) {
    const auto& cbi = c.conv_backward_input;
    if (!cbi.enabled || !cbi.strided) return false;
    const ClusterInput& a = cbi.unfused[0];
    const ClusterInput& b = cbi.unfused[1];
    const int64_t G = a.arg_shape[0], M = a.arg_shape[1], K = cbi.matmul_k, N = b.arg_shape[2];
    const int64_t FH = cbi.filter_h, FW = cbi.filter_w, GC = N / (FH * FW), OH = cbi.out_h, OW = cbi.out_w, IH = cbi.in_h, IW = cbi.in_w;
    const int64_t images = M / (OH * OW), PY = cbi.unpad_h, PX = cbi.unpad_w, H = IH - 2 * PY, W = IW - 2 * PX, C = G * GC;
    const OpNode& mm = g.ops().nodes[c.node_id];
    DSC_CHECK(mm.shape[0] == 1 && H >= 1 && W >= 1 && GC * FH * FW == N, "strided conv backward: unexpected shapes");
    const bool rows_mode = mm.op.output_mode == MatMulOutputMode::Rows;
    (void)rows_mode;  // the operands are addressed through their own chains, whatever the product's layout was
    const int64_t out_count = images * H * W * C;
    const std::string name = "k" + num(ci);
    std::ostringstream os;
    os << "// " << c.label << "  [gather: one thread per image-gradient element]\n";
    os << "extern \"C\" __global__ void __launch_bounds__(256) " << name << "(const float* A, const float* B, float* C, const unsigned* dsc_step) {\n";
    os << "    (void)dsc_step;\n    const unsigned e = blockIdx.x * 256u + threadIdx.x;\n    if (e >= " << unum(out_count) << ") return;\n";
    os << "    const int ch = (int)(e % " << unum(C) << "), x = (int)((e / " << unum(C) << ") % " << unum(W) << "), y = (int)((e / " << unum(C * W) << ") % " << unum(H)
       << "), image = (int)(e / " << unum(C * W * H) << ");\n";
    os << "    const int batch = ch / " << GC << ", gc = ch % " << GC << ";\n";
    os << "    // padded positions that fold onto (y, x): the pixel itself, plus the pad rows / columns beyond a border pixel\n";
    os << "    const int y_lo = y == 0 ? 0 : y + " << PY << ", y_hi = y == " << H - 1 << " ? " << IH - 1 << " : y + " << PY << ";\n";
    os << "    const int x_lo = x == 0 ? 0 : x + " << PX << ", x_hi = x == " << W - 1 << " ? " << IW - 1 << " : x + " << PX << ";\n";
    os << "    float acc = 0.f;\n";
    os << "    for (int yp = y_lo; yp <= y_hi; ++yp)\n    for (int xp = x_lo; xp <= x_hi; ++xp) {\n";
    // only the taps congruent to the position modulo the stride can be window elements
    os << "        for (int fy = yp % " << cbi.stride_h << "; fy < " << FH << "; fy += " << cbi.stride_h << ") {\n";
    os << "            const int ty = yp - fy;\n            if (ty < 0 || ty / " << cbi.stride_h << " >= " << OH << ") continue;\n";
    os << "            for (int fx = xp % " << cbi.stride_w << "; fx < " << FW << "; fx += " << cbi.stride_w << ") {\n";
    os << "                const int tx = xp - fx;\n                if (tx < 0 || tx / " << cbi.stride_w << " >= " << OW << ") continue;\n";
    os << "                const int gm = (image * " << OH << " + ty / " << cbi.stride_h << ") * " << OW << " + tx / " << cbi.stride_w << ";\n";
    os << "                const int gn = (fy * " << FW << " + fx) * " << GC << " + gc;\n";
    os << "                #pragma unroll\n                for (int gk = 0; gk < " << K << "; ++gk) {\n";
    int uniq = 0;
    std::string ia = emit_chain(os, a.chain, {{"batch", M * K, G}, {"gm", K, M}, {"gk", 1, K}}, uniq, "                    ");
    std::string ib = emit_chain(os, b.chain, {{"batch", K * N, G}, {"gk", N, K}, {"gn", 1, N}}, uniq, "                    ");
    os << "                    acc = fmaf(A[" << ia << "], B[" << ib << "], acc);\n                }\n            }\n        }\n    }\n    C[e] = acc;\n}\n\n";
    out->source = os.str();
    KernelLaunch l;
    l.entry = name;
    l.grid_x = (uint32_t)div_round_up(out_count, 256);
    l.label = c.label;
    l.cluster = ci;
    l.args = {{KernelArg::NodeBuffer, a.node_id, 0}, {KernelArg::NodeBuffer, b.node_id, 0}, {KernelArg::NodeBuffer, c.outputs[0], 0}};
    l.algorithmic_bytes = chain_bytes(g, a) + chain_bytes(g, b) + 4.0 * (double)out_count;
    l.flops = 2.0 * (double)G * (double)M * (double)N * (double)K;
    out->launches.push_back(l);
    return true;
}

// Many tiny products: BC independent [M, K] x [K, N] with K * N a few dozen -- the depthwise 3x3 blur convolution of
// MaxBlurPool2D (module.rs:139-163, 221-245: groups = channels, one input and one output channel per group, K = 9, N = 1)
// and its backward outer product (K = 1, N = 9).  One thread per OUTPUT element in the output's own memory order, the whole
// reduction in registers (ascending k, explicit fmaf like every strict-FP32 GEMM here), operands through their view chains.
// In rows mode ([pixel, group, n], NHWC) consecutive threads are consecutive channels of one pixel, so every load of the
// window's nine taps is a contiguous run of channels; the thin-rows kernel this replaces ran one group per grid row and
// touched 4 bytes per 64-byte line (conv-blur-net m = 8192: 1126 + 977 us for the two forward blur convolutions).
bool gen_grouped_dot(const Graph& g, const Cluster& c, int ci, ClusterCode* out) {
    const OpNode& mm = g.ops().nodes[c.node_id];
    const ClusterInput& a = c.inputs[0];
    const ClusterInput& b = c.inputs[1];
    const int64_t BC = a.arg_shape[0], M = a.arg_shape[1], K = a.arg_shape[2], N = b.arg_shape[2];
    if (c.conv_backward_input.enabled || c.matmul_absorbs_reduce || mm.shape[0] != 1 || !c.column_sum.empty() || !c.epilogue.empty() || c.pool.enabled) return false;
    if (BC < 4 || K > 32 || N > 16 || K * N > 64 || BC * M * N < (1 << 16)) return false;
    const bool rows_mode = mm.op.output_mode == MatMulOutputMode::Rows;
    const int64_t out_count = BC * M * N;
    const std::string name = "k" + num(ci);
    std::ostringstream os;
    os << "// " << c.label << "  [one thread per output element, " << K << "-term dot product in registers]\n";
    os << "extern \"C\" __global__ void __launch_bounds__(256) " << name << "(const float* A, const float* B, float* C, const unsigned* dsc_step) {\n";
    os << "    (void)dsc_step;\n    const unsigned e = blockIdx.x * 256u + threadIdx.x;\n    if (e >= " << unum(out_count) << ") return;\n";
    os << "    const unsigned gn = e % " << unum(N) << ";\n";
    if (rows_mode) os << "    const unsigned batch = (e / " << unum(N) << ") % " << unum(BC) << ", gm = e / " << unum(N * BC) << ";\n";
    else os << "    const unsigned gm = (e / " << unum(N) << ") % " << unum(M) << ", batch = e / " << unum(N * M) << ";\n";
    os << "    float acc = 0.f;\n    #pragma unroll\n    for (unsigned gk = 0; gk < " << unum(K) << "; ++gk) {\n";
    int uniq = 0;
    std::string ia = emit_chain(os, a.chain, {{"batch", M * K, BC}, {"gm", K, M}, {"gk", 1, K}}, uniq, "        ");
    std::string ib = emit_chain(os, b.chain, {{"batch", K * N, BC}, {"gk", N, K}, {"gn", 1, N}}, uniq, "        ");
    os << "        acc = fmaf(A[" << ia << "], B[" << ib << "], acc);\n    }\n    C[e] = acc;\n}\n\n";
    out->source = os.str();
    KernelLaunch l;
    l.entry = name;
    l.grid_x = (uint32_t)div_round_up(out_count, 256);
    l.label = c.label;
    l.cluster = ci;
    l.args = {{KernelArg::NodeBuffer, a.node_id, 0}, {KernelArg::NodeBuffer, b.node_id, 0}, {KernelArg::NodeBuffer, c.outputs[0], 0}};
    l.algorithmic_bytes = chain_bytes(g, a) + chain_bytes(g, b) + 4.0 * (double)out_count;
    l.flops = 2.0 * (double)BC * (double)M * (double)N * (double)K;
    out->launches.push_back(l);
    return true;
}

// A dense layer's bias + activation (or activation backward) in the epilogue of the JIT tcgen05 GEMM: set while gen_matmul
// generates the product of a cluster whose absorbed epilogue the tensor-core kernel may evaluate on its accumulator; the
// kernel emitter sets `fused` when it did (SIREN / ReLU+PE layers: the TMA GEMM + a separate epilogue kernel moved every
// activation three times).
const Cluster* g_tc_epilogue = nullptr;
bool g_tc_epilogue_fused = false;

ClusterCode gen_matmul(const Graph& g, const Cluster& c, int ci, const CodegenOptions& opt) {
    const OpNode& mm = g.ops().nodes[c.node_id];
    const ClusterInput& a = c.inputs[0];
    const ClusterInput& b = c.inputs[1];
    const int64_t BC = a.arg_shape[0], M = a.arg_shape[1], K = a.arg_shape[2], N = b.arg_shape[2];
    const int64_t r_graph = mm.shape[0];
    const auto& cbi = c.conv_backward_input;
    if (cbi.enabled && cbi.strided) {
        ClusterCode code;
        DSC_CHECK(gen_conv_backward_gather(g, c, ci, &code), "strided conv backward-input cluster without a kernel");
        return code;
    }
    if (!c.epilogue.empty()) {
        // an absorbed per-element cluster runs in the epilogue of the kernels that support it ...
        ClusterCode code;
        if (opt.use_tf32 && gen_conv_forward(g, c, ci, opt, &code)) return code;
        code = ClusterCode();
        if (gen_thin_matmul(g, c, ci, opt, &code)) return code;
        // ... and as its own kernel behind any other GEMM: the product goes to scratch, never to a graph buffer
        Cluster plain = c;
        plain.epilogue.clear();
        plain.epilogue_product_input = -1;
        plain.inputs.resize(2);
        plain.outputs = {c.node_id};
        // Measured and rejected (B200, relu-pe / siren at m = 65536): 0.369 -> 0.450 ms and 0.388 -> 0.551 ms per step.  The
        // gathered GEMM runs one 128 x 256 tile per CTA with the epilogue serialised behind its MMAs (94 us for the widest
        // layer against 22 us TMA GEMM + 34 us epilogue kernel).  Kept behind DSC_TC_FUSE_EPILOGUE=1 as a measurement hook.
        static const bool fuse_in_tc = [] { const char* e = std::getenv("DSC_TC_FUSE_EPILOGUE"); return e && std::atoi(e) != 0; }();
        if (fuse_in_tc && opt.use_tf32 && !cbi.enabled && BC == 1 && r_graph == 1 && N % 4 == 0 && K <= 512 && !c.pool.enabled) {
            g_tc_epilogue = &c;
            g_tc_epilogue_fused = false;
            ClusterCode fused = gen_matmul(g, plain, ci, opt);
            g_tc_epilogue = nullptr;
            if (g_tc_epilogue_fused) return fused;
        }
        code = gen_matmul(g, plain, ci, opt);
        const int64_t product_offset = div_round_up(code.scratch_bytes, 256) * 256;
        code.scratch_bytes = product_offset + BC * M * N * 4;
        ClusterCode tail = gen_per_element(g, c.epilogue[0], ci, opt, "_epilogue");
        code.source += tail.source;
        for (auto& l : tail.launches) code.launches.push_back(l);
        for (auto& l : code.launches)
            for (auto& arg : l.args)
                if (arg.kind == KernelArg::NodeBuffer && arg.node_id == c.node_id) arg = KernelArg{KernelArg::Scratch, -1, product_offset};
        return code;
    }
    if (opt.use_tf32) {
        ClusterCode code;
        if (cbi.enabled ? gen_conv_backward_input(g, c, ci, opt, &code) : gen_conv_forward(g, c, ci, opt, &code)) return code;
        if (!cbi.enabled && gen_conv_weight_gradient(g, c, ci, opt, &code)) return code;
    }
    {
        ClusterCode code;
        if (gen_grouped_dot(g, c, ci, &code)) return code;
    }
    const bool rows_mode = mm.op.output_mode == MatMulOutputMode::Rows || cbi.enabled;  // fused output is [pixel, group, channel]
    const int64_t out_count = BC * M * N;
    // fused conv backward-input: an A element exists only where (y - fy, x - fx) is a window position
    std::string a_valid = "true";
    if (cbi.enabled) {
        std::ostringstream v;
        v << "(unsigned)((gm / " << cbi.in_w << ") % " << cbi.in_h << " - gk / " << cbi.filter_w * cbi.matmul_k << ") < " << cbi.out_h
          << "u && (unsigned)(gm % " << cbi.in_w << " - (gk / " << cbi.matmul_k << ") % " << cbi.filter_w << ") < " << cbi.out_w << "u";
        a_valid = v.str();
    }

    // Plain dense operands (row-major or transposed whole buffers) take the TMA + tcgen05 TF32 kernel when the
    // environment allows reduced operand precision; everything else stays on the strict-FP32 JIT path below.
    if (opt.use_tf32 && !g_tc_epilogue && BC == 1 && (c.matmul_absorbs_reduce || r_graph == 1) && M * N >= 128 * 128 && K >= 32) {
        auto plain = [&](const ClusterInput& in, int64_t rows, int64_t cols, bool* row_major) {
            if (in.chain.input_count != rows * cols || in.chain.views.size() > 1) return false;
            if (!in.chain.views.empty() && in.chain.views[0].any_clamp()) return false;
            if (eval_chain(in.chain, 0) != 0) return false;
            const int64_t sr = eval_chain(in.chain, cols), sc = eval_chain(in.chain, 1);
            if (sr == cols && sc == 1) { *row_major = true; return true; }
            if (sr == 1 && sc == rows) { *row_major = false; return true; }
            return false;
        };
        bool a_row_major = true, b_row_major = true;
        if (plain(a, M, K, &a_row_major) && plain(b, K, N, &b_row_major) && (a_row_major ? K : M) % 4 == 0 && (b_row_major ? N : K) % 4 == 0) {
            ClusterCode code;
            KernelLaunch l;
            l.kind = KernelLaunch::TensorGemm;
            l.entry = "dsc_gemm_tf32";
            l.label = "TensorCore" + c.label;
            l.cluster = ci;
            l.gemm_m = M; l.gemm_n = N; l.gemm_k = K;
            l.gemm_a_is_mk = a_row_major;
            l.gemm_b_is_kn = b_row_major;
            l.args = {{KernelArg::NodeBuffer, a.node_id, 0}, {KernelArg::NodeBuffer, b.node_id, 0}, {KernelArg::NodeBuffer, c.outputs[0], 0}};
            l.algorithmic_bytes = chain_bytes(g, a) + chain_bytes(g, b) + 4.0 * (double)out_count;
            l.flops = 2.0 * (double)M * (double)N * (double)K;
            // few output tiles and a long reduction (dense-layer weight gradients): slice k across the idle SMs
            const int64_t tiles = div_round_up(M, 128) * div_round_up(N, 128), k_blocks = div_round_up(K, 32);
            int64_t S = 1;
            if (tiles * 4 <= opt.sm_count && k_blocks >= 16) {  // below 4 slices the extra pass over [S, M, N] costs more than it saves
                S = std::min<int64_t>(opt.sm_count / tiles, k_blocks / 8);
                const int64_t per = div_round_up(k_blocks, S);
                S = div_round_up(k_blocks, per);
            }
            if (S > 1) {
                l.gemm_splits = (int)S;
                l.args[2] = {KernelArg::Scratch, -1, 0};
                code.launches.push_back(l);
                code.scratch_bytes = S * out_count * 4;
                const std::string sname = "k" + num(ci) + "_splitsum";
                code.source = subst(kSplitSumTemplate, {{"LABEL", c.label}, {"NAME", sname}, {"COUNT", num(out_count)}, {"S", num(S)}});
                KernelLaunch sum;
                sum.entry = sname;
                sum.grid_x = (uint32_t)div_round_up(out_count, 32);
                sum.label = "SplitSum " + c.label;
                sum.cluster = ci;
                sum.args = {{KernelArg::Scratch, -1, 0}, {KernelArg::NodeBuffer, c.outputs[0], 0}};
                code.launches.push_back(sum);
                return code;
            }
            code.launches.push_back(l);
            return code;
        }
    }

    {
        ClusterCode code;
        if (gen_thin_matmul(g, c, ci, opt, &code)) return code;
    }

    // Operands behind view chains (conv2d's im2col, grouped / transposed views) use the gathered tcgen05 kernel
    // when TF32 is allowed and the GEMM is big enough to matter; otherwise the strict-FP32 SIMT kernel.
    // shared-memory operand layout follows the direction that is contiguous in global memory
    const bool a_mn = !cbi.enabled && chain_vector_run_axis(a.chain, a.arg_shape, 2) < 4 && chain_vector_run_axis(a.chain, a.arg_shape, 1) == 4;
    const bool b_mn = chain_vector_run_axis(b.chain, b.arg_shape, 1) < 4 && chain_vector_run_axis(b.chain, b.arg_shape, 2) == 4;
    const bool tc = opt.use_tf32 && 2.0 * (double)BC * (double)M * (double)N * (double)K >= 5e7 && N >= 8 && (M >= 128 || (a_mn && K >= 1024));
    GemmTile t = choose_gemm_tile(M, N);
    if (tc) {
        t.bm = 128;
        t.bn = (int)std::min<int64_t>(div_round_up(N, 16) * 16, 256);
        t.nt = 256;
        t.tm = t.tn = 0;
        int64_t best_padded = INT64_MAX;
        for (int bk = 8; bk <= 96; bk += 8) {  // least zero padding of K, then the deepest tile
            if (bk > 64 && (128 + t.bn) * bk > 12 * 1024) break;  // staging registers: at most 12 float4 per thread beyond BK 64
            const int64_t padded = div_round_up(K, bk) * bk;
            if (padded <= best_padded) { best_padded = padded; t.bk = bk; }
        }
    }
    const int64_t tiles = div_round_up(M, t.bm) * div_round_up(N, t.bn) * BC;

    int64_t S, KC;
    if (c.matmul_absorbs_reduce || r_graph == 1) {
        // the backend owns the K split (the reference fixes it at ceil(K/1024), op.rs:74): enough CTAs to
        // fill the machine, at least 4 k-tiles per split, partial-sum traffic well under the operand traffic
        const int64_t target = (int64_t)opt.sm_count * (tc ? 4 : std::max(2, std::min(8, 1024 / t.nt)));
        const int64_t in_elems = a.chain.addressed_count() + b.chain.addressed_count();
        S = 1;
        if (tiles * 2 <= target && K >= std::max<int64_t>(8 * t.bk, 512)) {  // a split costs a second launch and a workspace round trip: not for short reductions
            S = std::min<int64_t>(div_round_up(target, tiles), K / (4 * t.bk));
            S = std::min<int64_t>(S, std::max<int64_t>(1, in_elems / (4 * out_count)));
        }
        S = std::max<int64_t>(S, 1);
        KC = div_round_up(div_round_up(K, S), t.bk) * t.bk;
        S = div_round_up(K, KC);
    } else {
        // graph-level chunks must be honoured because something other than the Reduce reads them (kernel.rs:436)
        S = r_graph;
        KC = div_round_up(div_round_up(K, 16), S) * 16;
    }
    const bool via_scratch = S > 1 && (c.matmul_absorbs_reduce || r_graph == 1);

    auto fast_axis = [&](const ClusterInput& in, int64_t rows, int64_t cols) {
        // true when walking the row index is the (more) contiguous direction in memory
        const int64_t mid = (rows / 2) * cols + cols / 2;
        int64_t d_col = cols > 1 ? std::llabs(eval_chain(in.chain, mid + (cols / 2 + 1 < cols ? 1 : -1)) - eval_chain(in.chain, mid)) : INT64_MAX;
        int64_t d_row = rows > 1 ? std::llabs(eval_chain(in.chain, mid + (rows / 2 + 1 < rows ? cols : -cols)) - eval_chain(in.chain, mid)) : INT64_MAX;
        if (d_col == 0) d_col = INT64_MAX;
        if (d_row == 0) d_row = INT64_MAX;
        return d_row < d_col;
    };
    const bool a_m_fast = fast_axis(a, M, K);   // A element (m,k): rows = m
    const bool b_k_fast = fast_axis(b, K, N);   // B element (k,n): rows = k
    // tile element i = tid + j*NT -> (fast coordinate, slow coordinate); written so that the fast coordinate is
    // visibly independent of j whenever the thread count allows, which lets its index math leave the loops
    auto decode = [&](const char* x, const char* bx, int bxv, const char* y, const char* by, int byv, bool x_fast) {
        const char *f = x_fast ? x : y, *bf = x_fast ? bx : by, *s = x_fast ? y : x;
        const int bfv = x_fast ? bxv : byv;
        std::ostringstream d;
        if (t.nt % bfv == 0) d << "const int " << f << " = tid % " << bf << ", " << s << " = tid / " << bf << " + j * (NT / " << bf << ");";
        else if (bfv % t.nt == 0) d << "const int " << f << " = tid + (j % (" << bf << " / NT)) * NT, " << s << " = j / (" << bf << " / NT);";
        else d << "const int " << f << " = i % " << bf << ", " << s << " = i / " << bf << ";";
        return d.str();
    };
    int uniq = 0;
    std::ostringstream ca, cb;
    std::string ia = emit_chain(ca, a.chain, {{"batch", M * K, BC}, {"gm", K, M}, {"gk", 1, K}}, uniq, "                ");
    std::string ib = emit_chain(cb, b.chain, {{"batch", K * N, BC}, {"gk", N, K}, {"gn", 1, N}}, uniq, "                ");
    const std::string name = "k" + num(ci);
    std::string c_row = rows_mode ? "(((size_t)split * M + gm) * BC + batch) * N" : "(((size_t)split * BC + batch) * M + gm) * N";
    ClusterCode code;
    int64_t tmem_cols = 32;
    while (tmem_cols < t.bn) tmem_cols *= 2;
    const bool a_vec = a_mn || (chain_vector_run_axis(a.chain, a.arg_shape, 2) == 4 && KC % 4 == 0);
    const bool b_vec = b_mn || (chain_vector_run_axis(b.chain, b.arg_shape, 1) == 4 && KC % 4 == 0);
    // shared-memory bytes of one pipeline stage (mirrors A_BYTES / B_BYTES of the templates)
    auto stage_bytes = [&](bool mn, int rows) {
        const int64_t bytes = mn ? (int64_t)div_round_up(rows, 32) * t.bk * 128 : (int64_t)(t.bk / 4) * ((rows / 8) * 128 + 16);
        return div_round_up(bytes, 1024) * 1024;
    };
    const int64_t one_stage = tc ? stage_bytes(a_mn, t.bm) + stage_bytes(b_mn, t.bn) : 0;
    const int64_t k_tiles = div_round_up(std::min<int64_t>(KC, K), t.bk);
    // every unit is one aligned 16-byte run and the k loop is long enough to fill a pipeline: cp.async stages;
    // short-K GEMMs (convolution forward / backward-input) use the persistent register-staged kernel instead
    const bool async_copy = tc && a_vec && b_vec && k_tiles >= 8 && !cbi.enabled;
    const int64_t stages = async_copy ? std::max<int64_t>(2, std::min<int64_t>({4, 196608 / std::max<int64_t>(one_stage, 1), k_tiles})) : 2;
    const bool fuse_epilogue = tc && g_tc_epilogue != nullptr && !via_scratch && N % 4 == 0;
    const EpilogueCode epi = tc ? gen_epilogue(g, fuse_epilogue ? *g_tc_epilogue : c, name, opt) : EpilogueCode();
    if (tc)
        code.source = subst(async_copy ? kMatMulTc3Template : kMatMulTc2Template,
                            {{"LABEL", fuse_epilogue ? g_tc_epilogue->label : c.label}, {"NAME", name}, {"STORE4", epi.store4}, {"EPI_PARAMS", epi.params}, {"EPI_ARGS", epi.args}, {"BN", num(t.bn)}, {"BK", num(t.bk)}, {"M", num(M)}, {"N", num(N)}, {"K", num(K)},
                             {"KC", num(KC)}, {"BC", num(BC)}, {"TMEM_COLS", num(tmem_cols)}, {"A_MN", a_mn ? "true" : "false"}, {"STAGES", num(stages)},
                             {"B_MN", b_mn ? "true" : "false"}, {"A_LAYOUT", a_mn ? "MN-major" : "K-major"}, {"B_LAYOUT", b_mn ? "MN-major" : "K-major"},
                             {"A_VEC", a_vec ? "true" : "false"}, {"B_VEC", b_vec ? "true" : "false"}, {"A_CHAIN", ca.str()}, {"A_IDX", ia},
                             {"B_CHAIN", cb.str()}, {"B_IDX", ib}, {"C_ROW", c_row}, {"A_VALID", a_valid}});
    OperandLoad la, lb;
    la.ptr = "A"; lb.ptr = "B";
    la.bytes = chain_bytes(g, a); lb.bytes = chain_bytes(g, b);
    if (!tc) {  // the strict-FP32 SIMT kernel evaluates operand producers element by element while gathering its tiles
        la = operand_load(g, 0, name, a, opt);
        lb = operand_load(g, 1, name, b, opt);
    }
    if (!tc)
    code.source = subst(kMatMulTemplate,
                        {{"LABEL", c.label + (la.fused ? "  [A = " + g_prologue->producer[0]->label + "]" : "") + (lb.fused ? "  [B = " + g_prologue->producer[1]->label + "]" : "")},
                         {"PRO_FUNCS", la.funcs + lb.funcs}, {"PRO_PARAMS", la.params + lb.params}, {"A_LOAD1", la.load1(ia)}, {"B_LOAD1", lb.load1(ib)},
                         {"NAME", name}, {"NT", num(t.nt)}, {"BM", num(t.bm)}, {"BN", num(t.bn)}, {"BK", num(t.bk)}, {"TM", num(t.tm)},
                         {"TN", num(t.tn)}, {"M", num(M)}, {"N", num(N)}, {"K", num(K)}, {"KC", num(KC)}, {"BC", num(BC)},
                         {"A_DECODE", decode("lm", "BM", t.bm, "lk", "BK", t.bk, a_m_fast)},
                         {"B_DECODE", decode("ln", "BN", t.bn, "lk", "BK", t.bk, !b_k_fast)},
                         {"A_CHAIN", ca.str()}, {"B_CHAIN", cb.str()}, {"C_ROW", c_row}, {"A_VALID", a_valid}});
    KernelLaunch l;
    l.entry = name;
    l.grid_x = (uint32_t)(div_round_up(M, t.bm) * div_round_up(N, t.bn));
    if (tc && !async_copy && t.bn <= 32) {
        // Persistent over output tiles when the accumulator is narrow: each CTA loops over tiles blockIdx.x,
        // blockIdx.x + gridDim.x, ... and gathers the next tile's operands while the previous accumulator drains.
        // Measured on B200 (conv-net, m=8192): conv1 forward 0.29 -> 0.19 ms, conv2 forward 0.30 -> 0.26 ms.  Wide
        // accumulators (conv2 backward-input, N=72) lose: the fence.proxy.async ahead of the next MMA then also
        // waits for the previous tile's many global stores (0.31 -> 0.41 ms), so those stay one tile per CTA.
        const int64_t smem = stages * one_stage + 64 + 1024;
        const int64_t resident = std::max<int64_t>(1, std::min<int64_t>({8, (200 * 1024) / smem, 512 / tmem_cols}));
        l.grid_x = (uint32_t)std::min<int64_t>(l.grid_x, (int64_t)opt.sm_count * resident);
    }
    l.grid_y = (uint32_t)BC;
    l.grid_z = (uint32_t)S;
    l.block = t.nt;
    l.label = tc ? "TensorCore" + c.label : c.label;
    if (tc) l.smem = (uint32_t)(stages * one_stage + 64 + 1024);  // stages, barriers, 1 KB alignment slack
    l.cluster = ci;
    l.args = {{KernelArg::NodeBuffer, a.node_id, 0}, {KernelArg::NodeBuffer, b.node_id, 0}};
    if (via_scratch) l.args.push_back({KernelArg::Scratch, -1, 0});
    else if (fuse_epilogue) l.args.push_back({KernelArg::NodeBuffer, g_tc_epilogue->outputs[0], 0});  // the raw product is never stored: any valid address
    else l.args.push_back({KernelArg::NodeBuffer, c.outputs[0], 0});
    if (fuse_epilogue) {
        l.args.insert(l.args.end(), epi.launch_args.begin(), epi.launch_args.end());
        l.label = "TensorCore" + g_tc_epilogue->label;
        g_tc_epilogue_fused = true;
    }
    bind_operand_loads(l, &code, &la, &lb);
    l.algorithmic_bytes = la.bytes + lb.bytes + (fuse_epilogue ? epi.bytes : 4.0 * (double)out_count);
    l.flops = 2.0 * (double)BC * (double)M * (double)N * (double)K;
    code.launches.push_back(l);
    if (via_scratch) {
        code.scratch_bytes = S * out_count * 4;
        const std::string sname = name + "_splitsum";
        code.source += subst(kSplitSumTemplate, {{"LABEL", c.label}, {"NAME", sname}, {"COUNT", num(out_count)}, {"S", num(S)}});
        KernelLaunch s;
        s.entry = sname;
        s.grid_x = (uint32_t)div_round_up(out_count, 32);
        s.label = "SplitSum " + c.label;
        s.cluster = ci;
        s.args = {{KernelArg::Scratch, -1, 0}, {KernelArg::NodeBuffer, c.outputs[0], 0}};
        code.launches.push_back(s);
    }
    if (cbi.enabled && (cbi.unpad_h > 0 || cbi.unpad_w > 0)) {
        // absorbed Unpads without the halo kernel's epilogue: the padded gradient goes to scratch and the plain
        // Unpad kernels run from there, in the graph's order
        const int64_t images = M / (cbi.in_h * cbi.in_w), channels = BC * N;
        int64_t offset = div_round_up(code.scratch_bytes, 256) * 256;
        KernelArg src{KernelArg::Scratch, -1, offset};
        code.launches.back().args.back() = src;
        int64_t h = cbi.in_h, w = cbi.in_w;
        offset += div_round_up(images * h * w * channels * 4, 256) * 256;
        const int order[2] = {cbi.unpad_first_axis == 2 ? 2 : 1, cbi.unpad_first_axis == 2 ? 1 : 2};
        int remaining = (cbi.unpad_h > 0) + (cbi.unpad_w > 0);
        for (int axis : order) {
            const int64_t pad = axis == 1 ? cbi.unpad_h : cbi.unpad_w;
            if (pad == 0) continue;
            (axis == 1 ? h : w) -= 2 * pad;
            const int64_t count = images * h * w * channels, inner = axis == 1 ? w * channels : channels;
            const std::string uname = name + "_unpad" + num(axis);
            code.source += subst(kUnpadTemplate, {{"LABEL", "Unpad of " + c.label}, {"NAME", uname}, {"COUNT", num(count)}, {"INNER", num(inner)},
                                                 {"LEN", num(axis == 1 ? h : w)}, {"PAD", num(pad)}, {"CHAIN", ""}, {"IDX", "e"}});
            KernelLaunch u;
            u.entry = uname;
            u.grid_x = (uint32_t)div_round_up(count, 256);
            u.label = "Unpad of " + c.label;
            u.cluster = ci;
            KernelArg dst = --remaining > 0 ? KernelArg{KernelArg::Scratch, -1, offset} : KernelArg{KernelArg::NodeBuffer, c.outputs[0], 0};
            u.args = {src, dst};
            code.launches.push_back(u);
            src = dst;
            offset += div_round_up(count * 4, 256) * 256;
        }
        code.scratch_bytes = offset;
    }
    return code;
}

// ---- unpad (reference: UnpadKernel, kernel.rs:644-710): adjoint of replicate padding ---------------

const char* kUnpadTemplate = R"(
// {{LABEL}}
extern "C" __global__ void __launch_bounds__(256) {{NAME}}(const float* in0, float* out0, const unsigned* dsc_step) {
    constexpr unsigned COUNT = {{COUNT}}u, INNER = {{INNER}}u;
    constexpr int LEN = {{LEN}}, PAD = {{PAD}};
    const unsigned o = blockIdx.x * 256u + threadIdx.x;
    if (o >= COUNT) return;
    const unsigned oi = o % INNER, oo = o / (INNER * LEN);
    const int a = (int)(o / INNER % LEN);
    const int k_min = a + PAD - (a == 0 ? PAD : 0);
    const int k_max = a + PAD + (a == LEN - 1 ? PAD : 0);
    float sum = 0.f;
    for (int k = k_min; k <= k_max; ++k) {
        const unsigned e = (oo * (LEN + 2 * PAD) + k) * INNER + oi;
{{CHAIN}}
        sum += in0[{{IDX}}];
    }
    out0[o] = sum;
}
)";

ClusterCode gen_unpad(const Graph& g, const Cluster& c, int ci) {
    const OpNode& node = g.ops().nodes[c.node_id];
    const ClusterInput& in = c.inputs[0];
    const int axis = node.op.axis;
    int64_t inner = 1;
    for (int d = axis + 1; d < node.shape.len(); ++d) inner *= node.shape[d];
    std::ostringstream chain;
    int uniq = 0;
    std::string idx = emit_chain(chain, in.chain, "e", uniq, "        ");
    const std::string name = "k" + num(ci);
    const int64_t count = node.shape.element_count();
    ClusterCode code;
    code.source = subst(kUnpadTemplate, {{"LABEL", c.label}, {"NAME", name}, {"COUNT", num(count)}, {"INNER", num(inner)},
                                         {"LEN", num(node.shape[axis])}, {"PAD", num(node.op.pad)}, {"CHAIN", chain.str()}, {"IDX", idx}});
    KernelLaunch l;
    l.entry = name;
    l.grid_x = (uint32_t)div_round_up(count, 256);
    l.label = c.label;
    l.cluster = ci;
    l.args = {{KernelArg::NodeBuffer, in.node_id, 0}, {KernelArg::NodeBuffer, c.outputs[0], 0}};
    l.algorithmic_bytes = chain_bytes(g, in) + 4.0 * (double)count;
    code.launches.push_back(l);
    return code;
}

// ---- windows_to_image (reference: WindowsToImageKernel, kernel.rs:712-810) -------------------------
// col2im in gather form.  Unlike the reference kernel this checks that the window position lies
// inside [0,out_h) x [0,out_w): the mathematically correct adjoint.

const char* kW2ITemplate = R"(
{{FUNCS}}// {{LABEL}}
extern "C" __global__ void __launch_bounds__(256) {{NAME}}(const float* A, float* out0, {{PARAMS}}const unsigned* dsc_step) {
    constexpr unsigned COUNT = {{COUNT}}u;
    constexpr int IN_H = {{IN_H}}, IN_W = {{IN_W}}, IN_C = {{IN_C}};
    constexpr int OUT_H = {{OUT_H}}, OUT_W = {{OUT_W}}, GROUPS = {{GROUPS}}, FH = {{FH}}, FW = {{FW}}, GNC = {{GNC}}, SW = {{SW}}, SH = {{SH}};
    const unsigned o = blockIdx.x * 256u + threadIdx.x;
    if (o >= COUNT) return;
    const int c = (int)(o % IN_C), x = (int)(o / IN_C % IN_W), y = (int)(o / (IN_C * IN_W) % IN_H);
    const unsigned batch = o / (IN_C * IN_W * IN_H);
    const int group = c / GNC, gc = c - group * GNC;
    float sum = 0.f;
    for (int fy = y % SH; fy < FH; fy += SH) {
        const int oy = (y - fy) / SH;
        if (y < fy || oy >= OUT_H) continue;
        for (int fx = x % SW; fx < FW; fx += SW) {
            const int ox = (x - fx) / SW;
            if (x < fx || ox >= OUT_W) continue;
            const unsigned e = (((((batch * OUT_H + oy) * OUT_W + ox) * GROUPS + group) * FH + fy) * FW + fx) * GNC + gc;
{{CHAIN}}
            sum += {{LOAD}};
        }
    }
    out0[o] = sum;
}
)";

// The window values may come from an operand prologue (graph.hpp OperandPrologue): max_pool2d's backward pass with
// overlapping windows (MaxBlurPool2D's stride-1 pooling, module.rs:221-245) selects the output gradient where the window
// element equals the window's maximum -- a per-element program over the 4x-expanded window array, whose only reader is this
// gather.  Evaluated here per gathered element, that array (1.5 GB at m = 8192 for the first pooling layer) never exists.
ClusterCode gen_w2i(const Graph& g, const Cluster& c, int ci, const CodegenOptions& opt) {
    const OpNode& node = g.ops().nodes[c.node_id];
    const ClusterInput& in = c.inputs[0];
    const Shape& ws = in.arg_shape;
    const int n = ws.len(), ni = node.shape.len();
    std::ostringstream chain;
    int uniq = 0;
    std::string idx = emit_chain(chain, in.chain, "e", uniq, "            ");
    const std::string name = "k" + num(ci);
    const int64_t count = node.shape.element_count();
    ClusterCode code;
    const OperandLoad la = operand_load(g, 0, name, in, opt);
    code.source = subst(kW2ITemplate,
                        {{"LABEL", c.label}, {"NAME", name}, {"COUNT", num(count)}, {"IN_H", num(node.shape[ni - 3])}, {"IN_W", num(node.shape[ni - 2])},
                         {"IN_C", num(node.shape[ni - 1])}, {"OUT_H", num(ws[n - 6])}, {"OUT_W", num(ws[n - 5])}, {"GROUPS", num(ws[n - 4])},
                         {"FH", num(ws[n - 3])}, {"FW", num(ws[n - 2])}, {"GNC", num(ws[n - 1])}, {"SW", num(node.op.stride_w)},
                         {"SH", num(node.op.stride_h)}, {"CHAIN", chain.str()}, {"LOAD", la.load1(idx)}, {"FUNCS", la.funcs}, {"PARAMS", la.params}});
    KernelLaunch l;
    l.entry = name;
    l.grid_x = (uint32_t)div_round_up(count, 256);
    l.label = c.label + (la.fused ? "  [windows = " + g_prologue->producer[0]->label + "]" : "");
    l.cluster = ci;
    l.args = {{KernelArg::NodeBuffer, in.node_id, 0}, {KernelArg::NodeBuffer, c.outputs[0], 0}};
    l.algorithmic_bytes = la.bytes + 4.0 * (double)count;
    bind_operand_loads(l, &code, &la, nullptr);
    code.launches.push_back(l);
    return code;
}

// ---- scatter_add (reference: ScatterAddKernel, kernel.rs:812-874) ----------------------------------
// The reference uses float atomics (order non-deterministic, README.md:21).  Here each CTA sorts the
// (row, position) pairs of one chunk of positions in shared memory, sums each run of equal rows with
// a fixed-shape segmented scan and writes one partial per touched row; a second kernel adds the
// chunk partials to the accumulator in ascending chunk order.  Bitwise reproducible run to run.

const char* kScatterTemplate = R"(
// {{LABEL}}: per-chunk sorted partial sums ({{NSRC}} chained source(s))
extern "C" __global__ void __launch_bounds__(256) {{NAME}}_part({{SRC_PARAMS}}float* partial, const unsigned* dsc_step) {
    constexpr unsigned CH = {{CH}}u, ROWS = {{ROWS}}u, INNER = {{INNER}}u, OUTER = {{OUTER}}u, PER = CH / 256u;
    constexpr unsigned NCHUNK = {{NCHUNK}}u, CPB = {{CPB}}u;  // chunks in total, chunks per CTA
    constexpr bool TABLE = {{TABLE}};  // the whole table fits in shared memory: a CTA adds its chunks there, in chunk order,
                                       // and writes one dense partial (no zero fill of the scratch buffer, CPB x fewer partials)
    constexpr unsigned NONE = 0xffffffffu;
    __shared__ unsigned keys[CH + 1];
    struct Run { unsigned row; float sum; unsigned open; };
    __shared__ Run warp_tail[8];
    __shared__ float table[TABLE ? ROWS * INNER : 1];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, outer = blockIdx.y;
    if (TABLE) {
        for (unsigned i = tid; i < ROWS * INNER; i += 256u) table[i] = 0.f;
        __syncthreads();
    }
    for (unsigned gchunk = blockIdx.x * CPB; gchunk < min(NCHUNK, (blockIdx.x + 1u) * CPB); ++gchunk) {
    // element i = tid + q * 256 lives in register q of thread tid while sorting
    unsigned sk[PER];
    #pragma unroll
    for (unsigned q = 0; q < PER; ++q) {
        const unsigned lp = tid + q * 256u;
        unsigned key = NONE;
{{LOAD_KEYS}}
        sk[q] = key;
    }
    // bitonic sort; keys are unique, so the order (row, then position) is fully determined.  Partners 256 or 512
    // apart sit in the same thread, partners closer than 32 in the same warp (shuffles); only the three distances
    // in between go through shared memory.
    #pragma unroll
    for (unsigned k = 2; k <= CH; k <<= 1) {
        #pragma unroll
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            if (j >= 256u) {
                #pragma unroll
                for (unsigned q = 0; q < PER; ++q) {
                    const unsigned qq = q ^ (j >> 8);
                    if (qq > q) {
                        const bool up = ((tid + q * 256u) & k) == 0;
                        const unsigned a = sk[q], b = sk[qq];
                        if ((a > b) == up) { sk[q] = b; sk[qq] = a; }
                    }
                }
            } else if (j < 32u) {
                #pragma unroll
                for (unsigned q = 0; q < PER; ++q) {
                    const unsigned i = tid + q * 256u;
                    const unsigned other = __shfl_xor_sync(0xffffffffu, sk[q], j);
                    const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
                    sk[q] = keep_min ? min(sk[q], other) : max(sk[q], other);
                }
            } else {
                #pragma unroll
                for (unsigned q = 0; q < PER; ++q) keys[tid + q * 256u] = sk[q];
                __syncthreads();
                #pragma unroll
                for (unsigned q = 0; q < PER; ++q) {
                    const unsigned i = tid + q * 256u;
                    const unsigned other = keys[i ^ j];
                    const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
                    sk[q] = keep_min ? min(sk[q], other) : max(sk[q], other);
                }
                __syncthreads();
            }
        }
    }
    #pragma unroll
    for (unsigned q = 0; q < PER; ++q) keys[tid + q * 256u] = sk[q];
    if (tid == 0) keys[CH] = NONE;
    __syncthreads();
    // sums of equal rows: thread t now owns the sorted elements t * PER .. t * PER + PER - 1
    unsigned row[PER + 1], pos[PER];
    #pragma unroll
    for (unsigned q = 0; q <= PER; ++q) {
        const unsigned key = keys[tid * PER + q];
        row[q] = key == NONE ? NONE : key / CH;
        if (q < PER) pos[q] = key % CH;
    }
    for (unsigned w = 0; w < INNER; ++w) {
        float val[PER];
        #pragma unroll
        for (unsigned q = 0; q < PER; ++q) {
            float v = 0.f;
            if (row[q] != NONE) {
                const unsigned lp = pos[q];
{{LOAD_VALUES}}
            }
            val[q] = v;
        }
        // inclusive segmented sums inside the thread, then across threads: a fixed order of additions
        #pragma unroll
        for (unsigned q = 1; q < PER; ++q)
            if (row[q] == row[q - 1]) val[q] += val[q - 1];
        Run run{row[PER - 1], val[PER - 1], row[0] == row[PER - 1] ? 1u : 0u};  // the run that ends at this thread's last element
        #pragma unroll
        for (unsigned d = 1; d < 32u; d <<= 1) {
            const unsigned r = __shfl_up_sync(0xffffffffu, run.row, d);
            const float s = __shfl_up_sync(0xffffffffu, run.sum, d);
            const unsigned o = __shfl_up_sync(0xffffffffu, run.open, d);
            if (lane >= d && run.open && r == run.row) { run.sum = s + run.sum; run.open = o; }
        }
        if (lane == 31u) warp_tail[warp] = run;
        __syncthreads();
        // carry into this thread = the run that ends at the previous thread's last element, over the whole block
        Run before{NONE, 0.f, 0u};
        for (unsigned pw = 0; pw < warp; ++pw) {  // earlier warps, in order
            const Run t = warp_tail[pw];
            if (t.open && before.row == t.row) before.sum = before.sum + t.sum;
            else before = t;
        }
        {
            Run prev;  // previous lane's inclusive run (within the warp)
            prev.row = __shfl_up_sync(0xffffffffu, run.row, 1);
            prev.sum = __shfl_up_sync(0xffffffffu, run.sum, 1);
            prev.open = __shfl_up_sync(0xffffffffu, run.open, 1);
            if (lane > 0) {
                if (prev.open && before.row == prev.row) before.sum = before.sum + prev.sum;
                else before = prev;
            }
        }
        #pragma unroll
        for (unsigned q = 0; q < PER; ++q) {
            if (row[q] == NONE) continue;
            const bool head_run = row[q] == row[0];
            const float total = (head_run && before.row == row[q]) ? before.sum + val[q] : val[q];
            if (row[q + 1] != row[q]) {
                if (TABLE) table[row[q] * INNER + w] += total;  // one thread per (row, w) and chunk; chunks are separated by barriers
                else partial[((gchunk * OUTER + outer) * ROWS + row[q]) * INNER + w] = total;
            }
        }
        __syncthreads();  // warp_tail is reused by the next w
    }
    }
    if (TABLE)
        for (unsigned i = tid; i < ROWS * INNER; i += 256u) partial[(blockIdx.x * OUTER + outer) * ROWS * INNER + i] = table[i];
}

// {{LABEL}}: accumulator + chunk partials in ascending chunk order
extern "C" __global__ void __launch_bounds__(1024) {{NAME}}_sum({{ACC_PARAM}}const float* partial, float* out0, const unsigned* dsc_step) {
    // 32 table elements x 32 chunk lanes per CTA: lane g adds chunks g, g+32, ... in ascending order, then the lane sums
    // are added to the accumulator in lane order -- a fixed order, independent of timing
    constexpr unsigned TOTAL = {{TOTAL}}u, NCHUNK = {{NBLOCKS}}u;
    __shared__ float red[32][33];
    const unsigned tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;
    const unsigned e = blockIdx.x * 32u + tx;
    float part = 0.f;
    if (e < TOTAL) {
        #pragma unroll 4
        for (unsigned c = ty; c < NCHUNK; c += 32u) part += partial[c * TOTAL + e];
    }
    red[ty][tx] = part;
    __syncthreads();
    if (ty != 0 || e >= TOTAL) return;
{{ACC_CHAIN}}
    float acc = {{ACC_VALUE}};
    #pragma unroll
    for (unsigned g = 0; g < 32u; ++g) acc += red[g][tx];
    out0[e] = acc;
}
)";

// Small tables (the hash-grid levels of image_fit: <= 4096 rows of 2 floats): the whole table lives in shared memory and
// the sort shrinks from 1024 keys per CTA (55 compare-exchange stages, a dozen block barriers) to 32 keys per WARP (15
// shuffle stages, no barrier).  Per round of 256 consecutive positions: every warp sorts its 32 (row, lane) keys, permutes
// the values along, runs a segmented inclusive scan over the sorted lanes (equal rows are adjacent; a fixed shuffle tree),
// and the last lane of each run holds that row's sum; the eight warps then add their sums to the table one warp after
// the other (eight barriers), so every row sees the additions in (round, warp, position) order: run-to-run reproducible.
// The next round's indices and values are loaded before the current one is processed.
const char* kScatterWarpTemplate = R"(
// {{LABEL}}: per-warp sorted partial sums into a shared-memory table ({{NSRC}} chained source(s))
extern "C" __global__ void __launch_bounds__(256) {{NAME}}_part({{SRC_PARAMS}}float* partial, const unsigned* dsc_step) {
    constexpr unsigned ROWS = {{ROWS}}u, INNER = {{INNER}}u, OUTER = {{OUTER}}u;
    constexpr unsigned NROUND = {{NROUND}}u, RPB = {{RPB}}u;  // rounds of 256 positions in total (every source padded to whole rounds), rounds per CTA
    constexpr unsigned RPI = 4u;  // rounds sorted per iteration: the eight ordered accumulation phases (block barriers) are paid once for all of them
    constexpr unsigned NONE = 0xffffffffu;
    __shared__ float table[ROWS * INNER];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, outer = blockIdx.y;
    for (unsigned i = tid; i < ROWS * INNER; i += 256u) table[i] = 0.f;
    __syncthreads();
    // index and values of this thread's position in `round` (both loads are issued together; a row outside the table drops the position)
    auto load_round = [&](unsigned round, unsigned& row, float (&val)[INNER]) {
        row = NONE;
        #pragma unroll
        for (unsigned w = 0; w < INNER; ++w) val[w] = 0.f;
        if (round >= NROUND) return;
{{LOAD}}
        if (row >= ROWS) row = NONE;
    };
    const unsigned first = blockIdx.x * RPB, last = min(NROUND, first + RPB);
    unsigned row[RPI], next_row[RPI];
    float val[RPI][INNER], next_val[RPI][INNER];
    #pragma unroll
    for (unsigned i = 0; i < RPI; ++i) load_round(first + i < last ? first + i : NROUND, row[i], val[i]);
    for (unsigned round = first; round < last; round += RPI) {
        #pragma unroll
        for (unsigned i = 0; i < RPI; ++i) load_round(round + RPI + i < last ? round + RPI + i : NROUND, next_row[i], next_val[i]);
        unsigned srow[RPI];
        float v[RPI][INNER];
        bool tail[RPI];
        #pragma unroll
        for (unsigned i = 0; i < RPI; ++i) {
            // sort the warp's (row, lane) keys: unique, so the order is fully determined (row, then position)
            unsigned key = row[i] == NONE ? NONE : row[i] * 32u + lane;
            #pragma unroll
            for (unsigned k = 2; k <= 32u; k <<= 1) {
                #pragma unroll
                for (unsigned j = k >> 1; j > 0; j >>= 1) {
                    const unsigned other = __shfl_xor_sync(0xffffffffu, key, j);
                    const bool keep_min = ((lane & j) == 0) == ((lane & k) == 0);
                    key = keep_min ? min(key, other) : max(key, other);
                }
            }
            srow[i] = key == NONE ? NONE : key >> 5;
            const unsigned src = key & 31u;
            #pragma unroll
            for (unsigned w = 0; w < INNER; ++w) v[i][w] = __shfl_sync(0xffffffffu, val[i][w], src);
            // segmented inclusive scan over the sorted lanes (equal rows are adjacent)
            #pragma unroll
            for (unsigned d = 1; d < 32u; d <<= 1) {
                const unsigned r_up = __shfl_up_sync(0xffffffffu, srow[i], d);
                #pragma unroll
                for (unsigned w = 0; w < INNER; ++w) {
                    const float v_up = __shfl_up_sync(0xffffffffu, v[i][w], d);
                    if (lane >= d && r_up == srow[i]) v[i][w] += v_up;
                }
            }
            const unsigned r_down = __shfl_down_sync(0xffffffffu, srow[i], 1);
            tail[i] = srow[i] != NONE && (lane == 31u || r_down != srow[i]);
        }
        // warps add their run sums one warp after the other, rounds in order inside a warp's turn: a fixed order per table row
        #pragma unroll
        for (unsigned phase = 0; phase < 8u; ++phase) {
            if (warp == phase) {
                #pragma unroll
                for (unsigned i = 0; i < RPI; ++i) {
                    if (tail[i]) {
                        #pragma unroll
                        for (unsigned w = 0; w < INNER; ++w) table[srow[i] * INNER + w] += v[i][w];
                    }
                    __syncwarp();  // two rounds of one warp may end runs of the same row
                }
            }
            __syncthreads();
        }
        #pragma unroll
        for (unsigned i = 0; i < RPI; ++i) {
            row[i] = next_row[i];
            #pragma unroll
            for (unsigned w = 0; w < INNER; ++w) val[i][w] = next_val[i][w];
        }
    }
    for (unsigned i = tid; i < ROWS * INNER; i += 256u) partial[(blockIdx.x * OUTER + outer) * ROWS * INNER + i] = table[i];
}
)";

int g_scatter_group_size = 1;  // members of the group being generated (generate_scatter_group_code), 1 = a kernel of its own

ClusterCode gen_scatter_add(const Graph& g, const Cluster& c, int ci, const CodegenOptions& opt) {
    const OpNode& node = g.ops().nodes[c.node_id];  // the first scatter of the chain: same table shape and axis as the rest
    const int nsrc = (int)c.members.size();
    const int axis = node.op.axis;
    const int64_t rows = node.shape[axis];
    int64_t inner = 1, outer = 1;
    for (int d = axis + 1; d < node.shape.len(); ++d) inner *= node.shape[d];
    for (int d = 0; d < axis; ++d) outer *= node.shape[d];
    int64_t max_count = 1;
    for (int s = 0; s < nsrc; ++s) max_count = std::max(max_count, c.inputs[2 * s].arg_shape[axis]);
    const int64_t ch = std::min<int64_t>(1024, std::max<int64_t>(256, pow2_ceil(max_count)));
    DSC_CHECK(rows * ch < (int64_t)0xffffffffLL, "scatter_add table too large for 32-bit sort keys");
    const int64_t total = node.shape.element_count();
    const OpNode& acc_node = g.ops().nodes[c.copy_from];
    const bool acc_literal = acc_node.op.kind == OpKind::Literal;

    int uniq = 0;
    std::ostringstream params, load_keys, load_values, value_funcs;
    int64_t chunk_base = 0;
    double bytes = 2.0 * 4.0 * (double)total;
    const std::string name = "k" + num(ci);
    // values of a source: an array element, or (Cluster::value_programs) a per-element program on loaded operands
    std::vector<std::string> value_params(nsrc), value_call(nsrc);
    auto has_program = [&](int s) { return s < (int)c.value_programs.size() && !c.value_programs[s].members.empty(); };
    for (int s = 0; s < nsrc; ++s) {
        if (!has_program(s)) {
            value_params[s] = "const float* values" + num(s) + ", ";
            continue;
        }
        const Cluster& vp = c.value_programs[s];
        std::ostringstream fparams, fargs;
        for (size_t i = 0; i < vp.inputs.size(); ++i) {
            value_params[s] += "const float* v" + num(s) + "_" + num((int64_t)i) + ", ";
            fparams << ", const float* in" << i;
            fargs << ", v" << s << "_" << i;
        }
        value_call[s] = fargs.str();
        value_funcs << "// values of source " << s << " computed while loading: " << vp.label << "\n";
        value_funcs << "__device__ __forceinline__ float " << name << "_value" << s << "(unsigned e" << fparams.str() << ") {\n";
        int vuniq = 0;
        std::vector<bool> no_vector(vp.inputs.size(), false);
        emit_per_element_ops(value_funcs, vp, opt, vuniq, no_vector, -1);
        value_funcs << "    return t" << vp.output_ops[0] << ";\n}\n";
    }
    auto value_at = [&](int s, const std::string& index) {
        return has_program(s) ? name + "_value" + num(s) + "((unsigned)(" + index + ")" + value_call[s] + ")" : "values" + num(s) + "[" + index + "]";
    };
    for (int s = 0; s < nsrc; ++s) {
        const ClusterInput& values = c.inputs[2 * s];
        const ClusterInput& indices = c.inputs[2 * s + 1];
        const int64_t count = values.arg_shape[axis];
        const int64_t nchunk = div_round_up(count, ch);
        params << value_params[s] << "const float* indices" << s << ", ";
        const std::string cond = "gchunk >= " + unum(chunk_base) + " && gchunk < " + unum(chunk_base + nchunk);
        load_keys << "        if (" << cond << ") {\n            const unsigned pos = (gchunk - " << unum(chunk_base) << ") * CH + lp;\n"
                  << "            if (pos < " << unum(count) << ") {\n                const unsigned e = pos;\n";
        std::string ii = emit_chain(load_keys, indices.chain, "e", uniq, "                ");
        load_keys << "                const unsigned row = (unsigned)__float_as_int(indices" << s << "[" << ii << "]);\n"
                  << "                if (row < ROWS) key = row * CH + lp;\n            }\n        }\n";
        load_values << "                if (" << cond << ") {\n                    const unsigned e = (outer * " << unum(count) << " + (gchunk - "
                    << unum(chunk_base) << ") * CH + lp) * INNER + w;\n";
        std::string vi = emit_chain(load_values, values.chain, "e", uniq, "                    ");
        load_values << "                    v = " << value_at(s, vi) << ";\n                }\n";
        chunk_base += nchunk;
        bytes += chain_bytes(g, indices);
        if (has_program(s)) for (const auto& in : c.value_programs[s].inputs) bytes += chain_bytes(g, in);
        else bytes += chain_bytes(g, values);
    }
    const int64_t nchunk_total = chunk_base;
    // warp-sorted form for tables that fit in shared memory (see kScatterWarpTemplate)
    const bool warp_form = rows * inner <= 8192 && inner <= 8 && rows * 32 < (int64_t)0xffffffffLL;
    std::ostringstream warp_load;
    int64_t nround_total = 0;
    if (warp_form) {
        int wuniq = 0;
        for (int s = 0; s < nsrc; ++s) {
            const ClusterInput& values = c.inputs[2 * s];
            const ClusterInput& indices = c.inputs[2 * s + 1];
            const int64_t count = values.arg_shape[axis];
            const int64_t nround = div_round_up(count, 256);
            warp_load << "        if (round >= " << unum(nround_total) << " && round < " << unum(nround_total + nround) << ") {\n"
                      << "            const unsigned pos = (round - " << unum(nround_total) << ") * 256u + tid;\n"
                      << "            if (pos < " << unum(count) << ") {\n                const unsigned e = pos;\n";
            std::string ii = emit_chain(warp_load, indices.chain, "e", wuniq, "                ");
            warp_load << "                row = (unsigned)__float_as_int(indices" << s << "[" << ii << "]);\n"
                      << "                #pragma unroll\n                for (unsigned w = 0; w < INNER; ++w) {\n"
                      << "                    const unsigned e = (outer * " << unum(count) << " + pos) * INNER + w;\n";
            std::string vi = emit_chain(warp_load, values.chain, "e", wuniq, "                    ");
            warp_load << "                    val[w] = " << value_at(s, vi) << ";\n                }\n            }\n        }\n";
            nround_total += nround;
        }
    }
    std::ostringstream ac;
    std::string acc_value;
    if (acc_literal) acc_value = "__uint_as_float(" + num(acc_node.op.literal_bits) + "u)";
    else acc_value = "acc_in[" + emit_chain(ac, c.inputs[2 * nsrc].chain, "e", uniq, "    ") + "]";
    // small tables are accumulated in shared memory, several chunks per CTA (two waves of CTAs at least)
    const bool table = rows * inner <= 8192;
    int64_t cpb = table ? std::max<int64_t>(1, std::min<int64_t>(8, nchunk_total / (2 * opt.sm_count))) : 1;
    int64_t nblocks = div_round_up(nchunk_total, cpb);
    ClusterCode code;
    code.source = value_funcs.str();
    if (warp_form) {
        // CTAs: enough to fill the machine, but each writes a whole table as its partial, so no more than the operands weigh
        const int64_t data_bytes = (int64_t)(bytes - 2.0 * 4.0 * (double)total);
        int64_t want = std::max<int64_t>(opt.sm_count, std::min<int64_t>(4 * (int64_t)opt.sm_count, data_bytes / std::max<int64_t>(1, total * 4)));
        // launched together with other tables (generate_scatter_group_code): the group as a whole fills the machine, so each
        // member takes its share of about two waves of CTAs -- fewer table-sized partials to write and to add up
        if (g_scatter_group_size > 1) want = std::min<int64_t>(want, std::max<int64_t>(32, 12 * (int64_t)opt.sm_count / g_scatter_group_size));
        want = std::min<int64_t>(want, nround_total);
        const int64_t rpb = div_round_up(nround_total, std::max<int64_t>(1, want));
        nblocks = div_round_up(nround_total, rpb);
        cpb = rpb;
        code.source += subst(kScatterWarpTemplate, {{"LABEL", c.label}, {"NAME", name}, {"NSRC", num(nsrc)}, {"SRC_PARAMS", params.str()}, {"ROWS", num(rows)},
                                                   {"INNER", num(inner)}, {"OUTER", num(outer)}, {"NROUND", num(nround_total)}, {"RPB", num(rpb)}, {"LOAD", warp_load.str()}});
        // the ordered sum of the partials is the second kernel of the other template: emit only that half
        std::string both = subst(kScatterTemplate,
                        {{"LABEL", c.label}, {"NAME", name}, {"NSRC", num(nsrc)}, {"SRC_PARAMS", params.str()}, {"CH", num(ch)}, {"ROWS", num(rows)},
                         {"INNER", num(inner)}, {"OUTER", num(outer)}, {"LOAD_KEYS", load_keys.str()}, {"LOAD_VALUES", load_values.str()},
                         {"ACC_PARAM", acc_literal ? "" : "const float* acc_in, "}, {"TOTAL", num(total)}, {"NCHUNK", num(nchunk_total)}, {"NBLOCKS", num(nblocks)},
                         {"CPB", num(cpb)}, {"TABLE", "true"}, {"ACC_CHAIN", ac.str()}, {"ACC_VALUE", acc_value}});
        const size_t at = both.find("// " + c.label + ": accumulator + chunk partials");
        DSC_CHECK(at != std::string::npos, "scatter template layout changed");
        code.source += both.substr(at);
    } else
    code.source += subst(kScatterTemplate,
                        {{"LABEL", c.label}, {"NAME", name}, {"NSRC", num(nsrc)}, {"SRC_PARAMS", params.str()}, {"CH", num(ch)}, {"ROWS", num(rows)},
                         {"INNER", num(inner)}, {"OUTER", num(outer)}, {"LOAD_KEYS", load_keys.str()}, {"LOAD_VALUES", load_values.str()},
                         {"ACC_PARAM", acc_literal ? "" : "const float* acc_in, "}, {"TOTAL", num(total)}, {"NCHUNK", num(nchunk_total)}, {"NBLOCKS", num(nblocks)},
                         {"CPB", num(cpb)}, {"TABLE", table ? "true" : "false"},
                         {"ACC_CHAIN", ac.str()}, {"ACC_VALUE", acc_value}});
    code.scratch_bytes = nblocks * total * 4;
    // one table-sized partial per block: fine for the hash grids this was built for, not for a large embedding
    // table with few indices -- refuse with a message instead of planning tens of GB of scratch
    DSC_CHECK(code.scratch_bytes <= ((int64_t)8 << 30),
              "scatter_add into a [" << rows << ", " << inner << "] table from " << max_count << " positions would need " << (code.scratch_bytes >> 20)
                                     << " MB of per-block partial tables; this backend's deterministic scatter keeps one table copy per block (tables up to ~1M floats)");
    if (!table) {
        KernelLaunch z;
        z.kind = KernelLaunch::ZeroScratch;
        z.zero_offset = 0;
        z.zero_bytes = code.scratch_bytes;
        z.label = "Fill(0) " + num(nblocks * total);
        z.cluster = ci;
        code.launches.push_back(z);
    }
    KernelLaunch p;
    p.entry = name + "_part";
    p.grid_x = (uint32_t)nblocks;
    p.grid_y = (uint32_t)outer;
    p.label = c.label;
    p.cluster = ci;
    for (int s = 0; s < nsrc; ++s) {
        if (has_program(s)) {
            for (size_t i = 0; i < c.value_programs[s].inputs.size(); ++i) {
                const int node = c.inputs[c.value_input_base[s] + i].node_id;
                p.args.push_back({KernelArg::NodeBuffer, node, 0});
                code.extra_reads.push_back(node);
            }
        } else {
            p.args.push_back({KernelArg::NodeBuffer, c.inputs[2 * s].node_id, 0});
        }
        p.args.push_back({KernelArg::NodeBuffer, c.inputs[2 * s + 1].node_id, 0});
    }
    p.args.push_back({KernelArg::Scratch, -1, 0});
    p.algorithmic_bytes = bytes;
    code.launches.push_back(p);
    KernelLaunch sm;
    sm.entry = name + "_sum";
    sm.grid_x = (uint32_t)div_round_up(total, 32);
    sm.block = 1024;
    sm.label = "ScatterSum " + node.shape.str();
    sm.cluster = ci;
    if (!acc_literal) sm.args.push_back({KernelArg::NodeBuffer, c.copy_from, 0});
    sm.args.push_back({KernelArg::Scratch, -1, 0});
    sm.args.push_back({KernelArg::NodeBuffer, c.outputs[0], 0});
    code.launches.push_back(sm);
    return code;
}

}  // namespace

// ---- scatter_add groups ---------------------------------------------------------------------------------------------------
// The ten hash-grid levels of image_fit (examples/image_fit/main.rs:190-199) each scatter into their own small table at
// the same point of the backward pass: ten partial kernels of a few hundred CTAs and ten ordered-sum kernels of 1-256
// CTAs, launched one after the other, each ending in a partial wave.  Independent scatter_adds of one dependency level
// whose tables live in shared memory are launched together: ONE partial kernel and ONE sum kernel whose block ranges
// select the member (its code is the member's own kernel, turned into a device function).  Every member keeps its own
// partials and its own fixed order of additions, so results are bit-identical to the ungrouped launches.
bool generate_scatter_group_code(const Graph& g, const std::vector<int>& members, const CodegenOptions& opt, ClusterCode* out) {
    if (members.size() < 2) return false;
    const int host = members.back();
    const std::string name = "k" + num(host);
    std::ostringstream src, part_params, sum_params, part_body, sum_body;
    KernelLaunch part, sum;
    int64_t scratch = 0, part_blocks = 0, sum_blocks = 0, max_table = 0;
    double bytes = 0;
    auto replace_once = [](std::string& text, const std::string& from, const std::string& to) {
        const size_t at = text.find(from);
        if (at == std::string::npos) return false;
        text.replace(at, from.size(), to);
        return true;
    };
    for (size_t m = 0; m < members.size(); ++m) {
        const int ci = members[m];
        const Cluster& c = g.clusters()[ci];
        if (c.kind != ClusterKind::ScatterAdd) return false;
        g_scatter_group_size = (int)members.size();
        ClusterCode code = gen_scatter_add(g, c, ci, opt);
        g_scatter_group_size = 1;
        if (code.launches.size() != 2 || code.launches[0].grid_y != 1 || code.source.find("per-warp sorted partial sums") == std::string::npos) return false;
        const std::string k = "k" + num(ci);
        std::string text = code.source;
        if (!replace_once(text, "extern \"C\" __global__ void __launch_bounds__(256) " + k + "_part(", "__device__ __forceinline__ void " + k + "_part_body(const unsigned dsc_bx, float* table, ") ||
            !replace_once(text, "    __shared__ float table[ROWS * INNER];\n", "") ||
            !replace_once(text, "extern \"C\" __global__ void __launch_bounds__(1024) " + k + "_sum(", "__device__ __forceinline__ void " + k + "_sum_body(const unsigned dsc_bx, float (*red)[33], ") ||
            !replace_once(text, "    __shared__ float red[32][33];\n", ""))
            return false;
        text = replace_all(replace_all(text, "blockIdx.x", "dsc_bx"), "blockIdx.y", "0u");
        src << text;
        const OpNode& node = g.ops().nodes[c.node_id];
        max_table = std::max<int64_t>(max_table, node.shape.element_count());
        const KernelLaunch& p = code.launches[0];
        const KernelLaunch& sm = code.launches[1];
        const int64_t base = scratch;
        scratch += div_round_up(code.scratch_bytes, 256) * 256;
        part_body << "    " << (m ? "else " : "") << "if (blockIdx.x < " << unum(part_blocks + p.grid_x) << ") " << k << "_part_body(blockIdx.x - " << unum(part_blocks) << ", table, ";
        for (size_t a = 0; a < p.args.size(); ++a) {
            const bool is_scratch = p.args[a].kind == KernelArg::Scratch;
            part_params << (is_scratch ? "float* " : "const float* ") << "a" << m << "_" << a << ", ";
            part_body << "a" << m << "_" << a << ", ";
            KernelArg arg = p.args[a];
            if (is_scratch) arg.scratch_offset += base;
            part.args.push_back(arg);
        }
        part_body << "dsc_step);\n";
        sum_body << "    " << (m ? "else " : "") << "if (blockIdx.x < " << unum(sum_blocks + sm.grid_x) << ") " << k << "_sum_body(blockIdx.x - " << unum(sum_blocks) << ", red, ";
        for (size_t a = 0; a < sm.args.size(); ++a) {
            const bool is_out = a + 1 == sm.args.size();
            sum_params << (is_out ? "float* " : "const float* ") << "b" << m << "_" << a << ", ";
            sum_body << "b" << m << "_" << a << ", ";
            KernelArg arg = sm.args[a];
            if (arg.kind == KernelArg::Scratch) arg.scratch_offset += base;
            sum.args.push_back(arg);
        }
        sum_body << "dsc_step);\n";
        part_blocks += p.grid_x;
        sum_blocks += sm.grid_x;
        bytes += p.algorithmic_bytes;
        for (const auto& in : c.inputs) out->extra_reads.push_back(in.node_id);
        out->extra_writes.push_back(c.outputs[0]);
    }
    src << "// " << members.size() << " scatter_add partial kernels of one level, one launch: block ranges select the table\n";
    src << "extern \"C\" __global__ void __launch_bounds__(256) " << name << "_parts(" << part_params.str() << "const unsigned* dsc_step) {\n";
    src << "    __shared__ float table[" << max_table << "];\n" << part_body.str() << "}\n";
    src << "// ... and their ordered sums\n";
    src << "extern \"C\" __global__ void __launch_bounds__(1024) " << name << "_sums(" << sum_params.str() << "const unsigned* dsc_step) {\n";
    src << "    __shared__ float red[32][33];\n" << sum_body.str() << "}\n\n";
    out->source = src.str();
    out->scratch_bytes = scratch;
    part.entry = name + "_parts";
    part.grid_x = (uint32_t)part_blocks;
    part.label = "ScatterAdd group (" + num((int64_t)members.size()) + " tables)";
    part.cluster = host;
    part.covers = members;
    part.algorithmic_bytes = bytes;
    sum.entry = name + "_sums";
    sum.grid_x = (uint32_t)sum_blocks;
    sum.block = 1024;
    sum.label = "ScatterSum group (" + num((int64_t)members.size()) + " tables)";
    sum.cluster = host;
    out->launches.push_back(part);
    out->launches.push_back(sum);
    return true;
}

namespace {

// ---- dense chains (graph.hpp DenseChain): a whole MLP training step per 128-row tile in ONE kernel -----------------
// SURVEY.md section 8f-1 (examples/image_fit/main.rs:50-118,259-275; module.rs:70-90).  One persistent CTA per SM walks
// 128-row tiles of the batch.  Per tile, in order: the forward layers (tcgen05 TF32 MMAs, accumulator in TMEM, the
// layer's bias + activation program evaluated on the accumulator, the activation written back to SHARED memory in the
// two layouts the next MMAs read: K-major for the next layer's A operand, MN-major for the weight gradient), the loss
// program, then for every layer from the last to the first the weight-gradient MMAs x_l^T dz_l (accumulated in TMEM
// across all tiles of the CTA) and the backward product dz_l W_l^T with the activation-backward program as its epilogue
// (dz_{l-1} overwrites x_l's MN-major copy in place, after reading it for the activation's sign).  Activations and their
// gradients never reach HBM: the kernel reads x_0 and the loss target, and writes the gradient with respect to x_0, one
// partial weight gradient / bias gradient / loss sum per CTA (added by a fixed-order split-sum launch).
// Weights live in shared memory for the whole kernel in both orientations (W_l^T as the forward B operand, W_l as the
// backward one), K-major without swizzle; MN-major operands use SWIZZLE_128B_BASE32B, the one MN-major layout tf32
// accepts (gemm_tc_template.inc).  All MMAs are M = 128: the weight gradients put the wider of (x_l, dz_l) on the
// accumulator lanes.  Bias gradients are exact FP32 column sums (warp transpose-reduction of the epilogue registers),
// not tensor-core products.
const char* kDenseChainHelpers = R"(
#ifndef DSC_DENSE_CHAIN_HELPERS
#define DSC_DENSE_CHAIN_HELPERS
// K-major operand without swizzle: 8-row x 16-byte core matrices, k-chunk (4 floats) stride lbo, 8-row group stride 128
__device__ __forceinline__ unsigned dc_km(unsigned lbo, int r, int c) {
    return (unsigned)(c >> 2) * lbo + (unsigned)(r >> 3) * 128u + (unsigned)(r & 7) * 16u + (unsigned)(c & 3) * 4u;
}
// MN-major operand, SWIZZLE_128B_BASE32B, 128 k (the tile's rows): 128-byte rows of 32 mn, 4 k per 512-byte atom,
// 32-byte chunks XOR-ed with (k mod 4), 32-mn block stride 16384
__device__ __forceinline__ unsigned dc_mn(int mn, int k) {
    return (unsigned)(mn >> 5) * 16384u + (unsigned)(k >> 2) * 512u + (unsigned)(k & 3) * 128u + (((((unsigned)mn & 31u) >> 3) ^ ((unsigned)k & 3u)) << 5) +
           (unsigned)(mn & 7) * 4u;
}
// shared memory through 32-bit shared-space addresses (the aligned base is computed from an integer, so the compiler would
// otherwise fall back to generic loads and stores)
__device__ __forceinline__ void dc_sts4(unsigned addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void dc_sts1(unsigned addr, float a) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory"); }
__device__ __forceinline__ float4 dc_lds4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long dc_desc(unsigned addr, unsigned lbo, unsigned sbo, unsigned mn_major) {
    return (unsigned long long)((addr >> 4) & 0x3fffu) | ((unsigned long long)((lbo >> 4) & 0x3fffu) << 16) | ((unsigned long long)((sbo >> 4) & 0x3fffu) << 32) |
           (1ull << 46) | ((unsigned long long)mn_major << 61);
}
__device__ __forceinline__ void dc_mma(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// one lane of a converged warp (the MMA-issuing thread): with elect.sync the compiler knows exactly one thread runs the
// branch and issues each tcgen05.mma once, instead of looping over the lanes it must assume active
__device__ __forceinline__ bool dc_elect() {
    unsigned pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void dc_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void dc_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "DC_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DC_DONE;\n"
        "bra DC_WAIT;\n"
        "DC_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void dc_ld16(unsigned taddr, float (&v)[16]) {
    unsigned u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]),
          "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(u[j]);
}
// Every lane holds 16 values (one row, 16 columns).  Returns in every lane the sum over the warp's 32 rows of column
// dc_colsum16_column(lane): a fixed tree of additions (lanes 16 apart first, then 8, 4, 2, 1).
__device__ __forceinline__ float dc_colsum16(const float (&o)[16], int lane) {
    float a[8], b[4], c[2];
    bool hi = (lane & 16) != 0;
    #pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (hi ? o[i + 8] : o[i]) + __shfl_xor_sync(0xffffffffu, hi ? o[i] : o[i + 8], 16);
    hi = (lane & 8) != 0;
    #pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = (hi ? a[i + 4] : a[i]) + __shfl_xor_sync(0xffffffffu, hi ? a[i] : a[i + 4], 8);
    hi = (lane & 4) != 0;
    #pragma unroll
    for (int i = 0; i < 2; ++i) c[i] = (hi ? b[i + 2] : b[i]) + __shfl_xor_sync(0xffffffffu, hi ? b[i] : b[i + 2], 4);
    hi = (lane & 2) != 0;
    float d = (hi ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, hi ? c[0] : c[1], 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}
__device__ __forceinline__ int dc_colsum16_column(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }
#endif
)";

}  // namespace

bool generate_dense_chain_code(const Graph& g, const DenseChain& ch, const CodegenOptions& opt, ClusterCode* out) {
    if (!opt.use_tf32) return false;
    const auto& clusters = g.clusters();
    const int L = (int)ch.forward.size();
    const int64_t M = ch.rows;
    auto up = [](int64_t v, int64_t a) { return (v + a - 1) / a * a; };
    std::vector<int64_t> W = ch.widths;  // W[l] = input width of layer l, W[l + 1] = its output width
    for (int l = 1; l < L; ++l)
        if (W[l] % 16 != 0 || W[l] > 128) return false;   // hidden activations are MMA K and N dimensions and 16-column epilogue chunks
    if (W[0] > 128 || W[L] > 128 || M >= ((int64_t)1 << 31) / 128) return false;
    const bool has_dx = ch.backward[0] >= 0;
    const int ci = ch.last_cluster();
    const std::string name = "k" + num(ci);

    // ---- shared memory plan -------------------------------------------------------------------------------------------
    // MN-major buffers first (an A operand narrower than 128 lets the MMA read up to 48 KB past its buffer: harmless
    // values on accumulator lanes nobody reads, but the addresses must exist), then the two K-major slots, the weights
    int64_t off = 0;
    std::vector<int64_t> mn_act(L), km_lbo_w_f(L), km_lbo_w_b(L), wf(L), wb(L, -1);
    for (int l = 0; l < L; ++l) { mn_act[l] = off; off += up(W[l], 32) / 32 * 16384; }
    const int64_t mn_dy = off;
    off += up(W[L], 32) / 32 * 16384;
    const int64_t mn_end = off;
    int64_t km_chunks = 0;
    for (int l = 0; l < L; ++l) km_chunks = std::max(km_chunks, up(W[l], 8) / 4);       // x_l as the forward A operand
    for (int l = 0; l < L; ++l) km_chunks = std::max(km_chunks, up(W[l + 1], 8) / 4);   // dz_l as the backward A operand
    constexpr int64_t kActLbo = 16 * 128 + 16;
    const int64_t km_slot_bytes = up(km_chunks * kActLbo, 1024);
    const int64_t km_slot[2] = {off, off + km_slot_bytes};
    off += 2 * km_slot_bytes;
    for (int l = 0; l < L; ++l) {
        const int64_t n16 = up(W[l + 1], 16), kp = up(W[l], 8);
        km_lbo_w_f[l] = n16 / 8 * 128 + 16;
        wf[l] = off;
        off += up(kp / 4 * km_lbo_w_f[l], 128);
        if (l > 0 || has_dx) {
            const int64_t k16 = up(W[l], 16), np = up(W[l + 1], 8);
            km_lbo_w_b[l] = k16 / 8 * 128 + 16;
            wb[l] = off;
            off += up(np / 4 * km_lbo_w_b[l], 128);
        }
    }
    const int64_t bar_off = off;
    off += 64;  // mbarrier, TMEM slot
    const int64_t red_off = off;
    off += 64;  // 16 warp partials of a loss sum
    // the widest read an M = 128 MN-major A operand can make starts at the last MN buffer and spans four 32-wide blocks:
    // small networks pad the allocation so that those addresses exist
    off = std::max(off, mn_dy + 4 * 16384);
    const int64_t smem_bytes = off + 1024;  // + alignment slack
    if (smem_bytes > 227 * 1024) return false;
    (void)mn_end;

    // ---- tensor memory plan: the working accumulator, then one weight-gradient accumulator per layer ---------------------
    int64_t work_cols = 16;
    for (int l = 0; l < L; ++l) work_cols = std::max(work_cols, up(W[l + 1], 16));
    for (int l = 0; l < L; ++l)
        if (l > 0 || has_dx) work_cols = std::max(work_cols, up(W[l], 16));
    std::vector<int64_t> dw_col(L), dw_cols(L);
    std::vector<char> dw_a_is_dz(L);
    int64_t cols = work_cols;
    for (int l = 0; l < L; ++l) {
        dw_a_is_dz[l] = W[l + 1] >= W[l];  // the wider operand goes on the lanes
        dw_cols[l] = up(dw_a_is_dz[l] ? W[l] : W[l + 1], 16);
        dw_col[l] = cols;
        cols += dw_cols[l];
    }
    int64_t tmem_cols = 32;
    while (tmem_cols < cols) tmem_cols *= 2;
    if (tmem_cols > 512) return false;

    // ---- kernel parameters ------------------------------------------------------------------------------------------------
    std::vector<KernelArg> args;
    std::ostringstream params;
    std::map<int, std::string> param_of_node;  // external node -> parameter name
    std::vector<int> reads;
    double bytes = 0;
    auto param_for = [&](int node, bool is_output) {
        auto it = param_of_node.find(node);
        if (it != param_of_node.end()) return it->second;
        const std::string p = "p" + num((int64_t)args.size());
        params << (is_output ? "float* __restrict__ " : "const float* __restrict__ ") << p << ", ";
        args.push_back({KernelArg::NodeBuffer, node, 0});
        param_of_node[node] = p;
        if (!is_output) reads.push_back(node);
        bytes += 4.0 * (double)g.ops().nodes[node].shape.element_count();
        return p;
    };
    const Cluster& f0 = clusters[ch.forward[0]];
    const std::string px = param_for(f0.inputs[0].node_id, false);
    std::vector<std::string> pw(L);
    for (int l = 0; l < L; ++l) pw[l] = param_for(clusters[ch.forward[l]].inputs[1].node_id, false);
    const std::string pdx = has_dx ? param_for(clusters[ch.backward[0]].outputs[0], true) : "";

    // one per-element program evaluated on 16 accumulator columns of this thread's row: `product` is the input fed from
    // the accumulator, `internal` (if >= 0) the input fed from hv[j] (the activation's MN-major copy)
    auto emit_program = [&](std::ostringstream& os, const Cluster& p, int product, int internal, int64_t N, const std::string& indent, bool guarded) {
        os << indent << "{\n";
        for (size_t i = 0; i < p.inputs.size(); ++i) {
            if ((int)i == product || (int)i == internal) continue;
            os << indent << "    const float* in" << i << " = " << param_for(p.inputs[i].node_id, false) << ";\n";
        }
        os << indent << "    #pragma unroll\n" << indent << "    for (int j = 0; j < 16; ++j) {\n";
        os << indent << "        const int c = c0 + j;\n";
        // rows past the batch and padding columns: only where a program loads per-row data or feeds a sum (the loss); the hidden
        // layers' widths are whole chunks, and what they compute on the zero rows of a ragged last tile is never used (the loss
        // program zeroes dz on those rows, and every later product of zeros is zero)
        os << indent << "        if (" << (guarded ? "c < " + num(N) + " && row_ok" : "true") << ") {\n";
        os << indent << "        const unsigned e = (unsigned)gm * " << unum(N) << " + (unsigned)c; (void)e;\n";
        std::function<std::string(int)> override_fn = [&](int input) -> std::string {
            if (input == product) return "acc[j]";
            if (input == internal) return "hv[j]";
            return "";
        };
        g_load_override = &override_fn;
        int uniq = 0;
        std::vector<bool> no_vector(p.inputs.size(), false);
        std::ostringstream body;
        emit_per_element_ops(body, p, opt, uniq, no_vector, -1);
        g_load_override = nullptr;
        os << body.str();
    };

    std::ostringstream os;
    os << kDenseChainHelpers;
    os << "// dense chain: " << L << " layers, widths";
    for (int64_t w : W) os << " " << w;
    os << ", " << M << " rows  [tcgen05 tf32, one persistent CTA per SM, activations in shared memory]\n";
    std::ostringstream body;  // the kernel body; the parameter list is complete only after every program was emitted
    const int64_t tiles = div_round_up(M, 128);
    const int64_t grid = std::min<int64_t>(tiles, opt.sm_count);
    // scratch layout (floats): per layer dW partials [grid][K*N], column-sum partials [grid*4][N]; loss sums [grid]
    std::vector<int64_t> ws_dw(L), ws_cs(L, -1);
    std::vector<int64_t> ws_sum(ch.sums.size());
    int64_t ws = 0;
    for (int l = 0; l < L; ++l) { ws_dw[l] = ws; ws += grid * W[l] * W[l + 1]; }
    for (int l = 0; l < L; ++l)
        if (!clusters[ch.weight_gradient[l]].column_sum.empty()) { ws_cs[l] = ws; ws += grid * 4 * W[l + 1]; }
    for (size_t s = 0; s < ch.sums.size(); ++s) { ws_sum[s] = ws; ws += grid; }

    body << "    constexpr int M = " << M << ", TILES = " << tiles << ";\n";
    body << "    extern __shared__ unsigned char dsc_smem_raw[];\n";
    body << "    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<unsigned long long>(dsc_smem_raw) + 1023ull) & ~1023ull);\n";
    body << "    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);\n";
    body << "    const unsigned bar = sbase + " << bar_off << "u;\n";
    body << "    unsigned* tmem_slot = reinterpret_cast<unsigned*>(smem + " << bar_off + 16 << ");\n";
    body << "    float* red = reinterpret_cast<float*>(smem + " << red_off << ");\n";
    body << "    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31, quad = warp & 3, grp = warp >> 2;  // 16 warps: TMEM lane quadrant x column group\n";
    body << "    const int r = quad * 32 + lane;  // this thread's row of the tile = its TMEM lane\n";
    body << "    const unsigned bar2 = bar + 8u;  // completion of the weight-gradient MMAs, which run under the backward epilogues\n";
    body << "    if (tid == 0) {\n        asm volatile(\"mbarrier.init.shared::cta.b64 [%0], 1;\" ::\"r\"(bar));\n        asm volatile(\"mbarrier.init.shared::cta.b64 [%0], 1;\" ::\"r\"(bar2));\n"
            "        asm volatile(\"fence.mbarrier_init.release.cluster;\" ::: \"memory\");\n    }\n";
    body << "    if (warp == 0) {\n        asm volatile(\"tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\" ::\"r\"((unsigned)__cvta_generic_to_shared(tmem_slot)), \"n\"("
         << tmem_cols << ") : \"memory\");\n        asm volatile(\"tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\" ::: \"memory\");\n    }\n";
    // weights
    for (int l = 0; l < L; ++l) {
        const int64_t K = W[l], N = W[l + 1], n16 = up(N, 16), kp = up(K, 8);
        body << "    for (int i = tid; i < " << n16 * kp << "; i += 512) {  // W_" << l << "^T: forward B operand [n][k]\n";
        body << "        const int n = i % " << n16 << ", k = i / " << n16 << ";\n";
        body << "        dc_sts1(sbase + " << wf[l] << "u + dc_km(" << km_lbo_w_f[l] << "u, n, k), (n < " << N << " && k < " << K << ") ? " << pw[l] << "[k * " << N
             << " + n] : 0.f);\n    }\n";
        if (wb[l] >= 0) {
            const int64_t k16 = up(K, 16), np = up(N, 8);
            body << "    for (int i = tid; i < " << k16 * np << "; i += 512) {  // W_" << l << ": backward B operand [k][n]\n";
            body << "        const int n = i % " << np << ", k = i / " << np << ";\n";
            body << "        dc_sts1(sbase + " << wb[l] << "u + dc_km(" << km_lbo_w_b[l] << "u, k, n), (n < " << N << " && k < " << K << ") ? " << pw[l] << "[k * " << N
                 << " + n] : 0.f);\n    }\n";
        }
    }
    body << "    asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n";
    body << "    asm volatile(\"tcgen05.fence::before_thread_sync;\" ::: \"memory\");\n    __syncthreads();\n    asm volatile(\"tcgen05.fence::after_thread_sync;\" ::: \"memory\");\n";
    // (broadcast through a shuffle so that the compiler knows the address is warp-uniform: tcgen05.mma takes uniform registers,
    // and a value it cannot prove uniform costs an elect / broadcast loop per MMA on the one issuing thread)
    body << "    const unsigned tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);\n";
    body << "    const unsigned tlane = tmem + ((unsigned)(quad * 32) << 16);\n";
    body << "    unsigned phase = 0, phase2 = 0;\n    bool first_tile = true;\n";
    // per-thread partial sums carried across tiles
    for (int l = 0; l < L; ++l)
        if (ws_cs[l] >= 0)
            for (int64_t i = 0; i < div_round_up(up(W[l + 1], 16) / 16, 4); ++i)
                body << "    float cs" << l << "_" << i << "[16];\n    #pragma unroll\n    for (int j = 0; j < 16; ++j) cs" << l << "_" << i << "[j] = 0.f;\n";
    for (size_t s = 0; s < ch.sums.size(); ++s) body << "    float lsum" << s << " = 0.f;\n";

    auto idesc = [&](bool a_mn, bool b_mn, int64_t n) {
        return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((unsigned)(n >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    };
    auto sync_then_issue = [&](std::ostringstream& o) {
        o << "        asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n";
        o << "        asm volatile(\"tcgen05.fence::before_thread_sync;\" ::: \"memory\");\n        __syncthreads();\n";
        o << "        if (warp == 0 && dc_elect()) {\n            asm volatile(\"tcgen05.fence::after_thread_sync;\" ::: \"memory\");\n";
    };
    auto commit_and_wait = [&](std::ostringstream& o) {
        o << "            dc_commit(bar);\n        }\n";
        o << "        dc_wait(bar, phase); phase ^= 1u;\n        asm volatile(\"tcgen05.fence::after_thread_sync;\" ::: \"memory\");\n";
    };
    auto store_km = [&](std::ostringstream& o, int64_t slot, int64_t kdim) {  // o[16] -> K-major A operand, columns below kdim
        o << "            #pragma unroll\n            for (int q = 0; q < 4; ++q)\n                if (c0 + 4 * q < " << kdim << ")\n";
        o << "                    dc_sts4(sbase + " << slot << "u + dc_km(" << kActLbo << "u, r, c0 + 4 * q), o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);\n";
    };
    auto store_mn = [&](std::ostringstream& o, int64_t buf) {
        o << "            #pragma unroll\n            for (int q = 0; q < 4; ++q)\n";
        o << "                dc_sts4(sbase + " << buf << "u + dc_mn(c0 + 4 * q, r), o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);\n";
    };
    auto column_sums = [&](std::ostringstream& o, int l) {  // of the o[16] just produced, into layer l's accumulators
        if (ws_cs[l] < 0) return;
        const int64_t per_warp = div_round_up(up(W[l + 1], 16) / 16, 4);
        // bias gradient: this thread's rows first (one register per column, across all tiles), the 128 rows of the tile slot
        // at the end of the kernel
        for (int64_t i = 0; i < per_warp; ++i) {
            o << "            if ((ch >> 2) == " << i << ") {\n                #pragma unroll\n                for (int j = 0; j < 16; ++j) cs" << l << "_" << i << "[j] += o[j];\n            }\n";
        }
    };

    int km_next = 0;  // K-major slots alternate: a tensor is written while the MMAs read the other slot
    body << "    for (int tile = blockIdx.x; tile < TILES; tile += gridDim.x) {\n";
    body << "        const int gm = tile * 128 + r;\n        const bool row_ok = gm < M;\n";
    {   // x_0
        const int64_t K = W[0], kp = up(K, 8);
        const int64_t slot = km_slot[km_next];
        if (K % 4 == 0) {
            body << "        for (int u = tid; u < " << 128 * (K / 4) << "; u += 512) {  // x_0 -> both operand layouts\n";
            body << "            const int xr = u / " << K / 4 << ", xc = (u % " << K / 4 << ") * 4;\n";
            body << "            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);\n";
            body << "            if (tile * 128 + xr < M) x = *reinterpret_cast<const float4*>(" << px << " + (size_t)(tile * 128 + xr) * " << K << " + xc);\n";
            body << "            dc_sts4(sbase + " << slot << "u + dc_km(" << kActLbo << "u, xr, xc), x.x, x.y, x.z, x.w);\n";
            body << "            dc_sts4(sbase + " << mn_act[0] << "u + dc_mn(xc, xr), x.x, x.y, x.z, x.w);\n        }\n";
        } else {
            body << "        for (int u = tid; u < " << 128 * K << "; u += 512) {  // x_0 -> both operand layouts\n";
            body << "            const int xr = u / " << K << ", xc = u % " << K << ";\n";
            body << "            const float x = tile * 128 + xr < M ? " << px << "[(size_t)(tile * 128 + xr) * " << K << " + xc] : 0.f;\n";
            body << "            dc_sts1(sbase + " << slot << "u + dc_km(" << kActLbo << "u, xr, xc), x);\n";
            body << "            dc_sts1(sbase + " << mn_act[0] << "u + dc_mn(xc, xr), x);\n        }\n";
        }
        if (kp != K) {
            body << "        for (int u = tid; u < " << 128 * (kp - K) << "; u += 512)  // zero the k padding of x_0\n";
            body << "            dc_sts1(sbase + " << slot << "u + dc_km(" << kActLbo << "u, u / " << (kp - K) << ", " << K << " + u % " << (kp - K) << "), 0.f);\n";
        }
    }
    // ---- forward layers ------------------------------------------------------------------------------------------------
    for (int l = 0; l < L; ++l) {
        const Cluster& fc = clusters[ch.forward[l]];
        const int64_t K = W[l], N = W[l + 1], kp = up(K, 8), n16 = up(N, 16);
        const int64_t a_slot = km_slot[km_next];
        km_next ^= 1;
        body << "        // forward layer " << l << ": [128, " << K << "] x W_" << l << " -> [128, " << N << "]\n";
        sync_then_issue(body);
        body << "            #pragma unroll\n            for (int kk = 0; kk < " << kp / 8 << "; ++kk)\n";
        body << "                dc_mma(tmem, dc_desc(sbase + " << a_slot << "u + kk * " << 2 * kActLbo << "u, " << kActLbo << "u, 128u, 0u), dc_desc(sbase + " << wf[l] << "u + kk * "
             << 2 * km_lbo_w_f[l] << "u, " << km_lbo_w_f[l] << "u, 128u, 0u), " << idesc(false, false, n16) << "u, kk != 0 ? 1u : 0u);\n";
        commit_and_wait(body);
        if (l < L - 1) {
            const int64_t out_slot = km_slot[km_next];
            body << "        for (int ch = grp; ch < " << n16 / 16 << "; ch += 4) {  // bias + activation on the accumulator -> x_" << l + 1 << " in shared memory\n";
            body << "            const int c0 = ch * 16;\n            float acc[16], o[16];\n            dc_ld16(tlane + (unsigned)c0, acc);\n";
            if (fc.epilogue.empty()) {
                body << "            #pragma unroll\n            for (int j = 0; j < 16; ++j) o[j] = acc[j];\n";
            } else {
                const Cluster& p = fc.epilogue[0];
                emit_program(body, p, fc.epilogue_product_input, -1, N, "            ", N % 16 != 0);
                body << "                o[j] = t" << p.output_ops[0] << ";\n";
                body << "                } else o[j] = 0.f;\n                }\n            }\n";
            }
            store_km(body, out_slot, up(N, 8));
            store_mn(body, mn_act[l + 1]);
            body << "        }\n";
        } else {
            // the loss program on the last product: dz_{L-1} and the values that are summed over the batch
            const Cluster& p = ch.loss_in_epilogue ? fc.epilogue[0] : clusters[ch.loss];
            int product = ch.loss_in_epilogue ? fc.epilogue_product_input : -1;
            for (size_t i = 0; i < p.inputs.size() && !ch.loss_in_epilogue; ++i)
                if (p.inputs[i].node_id == fc.outputs[0]) product = (int)i;
            const int64_t out_slot = km_slot[km_next];
            body << "        for (int ch = grp; ch < " << n16 / 16 << "; ch += 4) {  // loss program on the last product -> dz_" << l << "\n";
            body << "            const int c0 = ch * 16;\n            float acc[16], o[16];\n            dc_ld16(tlane + (unsigned)c0, acc);\n";
            if (!fc.epilogue.empty() && !ch.loss_in_epilogue) return false;  // (a bias absorbed into the last layer would need two programs here)
            emit_program(body, p, product, -1, N, "            ", true);
            body << "                o[j] = t" << p.output_ops[ch.loss_gradient_output] << ";\n";
            for (size_t s = 0; s < ch.sums.size(); ++s) body << "                lsum" << s << " += t" << p.output_ops[ch.sums[s].output] << ";\n";
            body << "                } else o[j] = 0.f;\n                }\n            }\n";
            store_km(body, out_slot, up(N, 8));
            store_mn(body, mn_dy);
            column_sums(body, l);
            body << "        }\n";
        }
    }
    // ---- backward: weight gradient and backward product per layer, last first ---------------------------------------------
    for (int l = L - 1; l >= 0; --l) {
        const int64_t K = W[l], N = W[l + 1], np = up(N, 8), k16 = up(K, 16);
        const int64_t dz_slot = km_slot[km_next];
        km_next ^= 1;
        const int64_t dz_mn = l == L - 1 ? mn_dy : mn_act[l + 1];  // dz_l lives where x_{l+1} was
        const bool has_b = l > 0 || has_dx;
        body << "        // layer " << l << ": dW_" << l << " += x_" << l << "^T dz_" << l << (has_b ? "; dz W^T -> activation backward\n" : "\n");
        // the backward product first (its epilogue is the critical path), then the weight-gradient MMAs, which the tensor core
        // works through while the epilogue warps read the accumulator and evaluate the activation backward; the epilogue's
        // stores (dz_{l-1} over x_l's MN-major copy, an operand of those MMAs) wait for their completion on bar2
        sync_then_issue(body);
        if (has_b) {
            body << "            #pragma unroll\n            for (int kk = 0; kk < " << np / 8 << "; ++kk)\n";
            body << "                dc_mma(tmem, dc_desc(sbase + " << dz_slot << "u + kk * " << 2 * kActLbo << "u, " << kActLbo << "u, 128u, 0u), dc_desc(sbase + " << wb[l] << "u + kk * "
                 << 2 * km_lbo_w_b[l] << "u, " << km_lbo_w_b[l] << "u, 128u, 0u), " << idesc(false, false, k16) << "u, kk != 0 ? 1u : 0u);\n";
            body << "            dc_commit(bar);\n";
        }
        {
            const int64_t a_buf = dw_a_is_dz[l] ? dz_mn : mn_act[l], b_buf = dw_a_is_dz[l] ? mn_act[l] : dz_mn;
            body << "            #pragma unroll\n            for (int kk = 0; kk < 16; ++kk)\n";
            body << "                dc_mma(tmem + " << dw_col[l] << "u, dc_desc(sbase + " << a_buf << "u + kk * 1024u, 16384u, 512u, 1u), dc_desc(sbase + " << b_buf
                 << "u + kk * 1024u, 16384u, 512u, 1u), " << idesc(true, true, dw_cols[l]) << "u, (!first_tile || kk != 0) ? 1u : 0u);\n";
            body << "            dc_commit(bar2);\n        }\n";
        }
        if (has_b) body << "        dc_wait(bar, phase); phase ^= 1u;\n        asm volatile(\"tcgen05.fence::after_thread_sync;\" ::: \"memory\");\n";
        if (l > 0) {
            const Cluster& bc = clusters[ch.backward[l]];
            const Cluster& p = bc.epilogue[0];
            int internal = -1;
            for (size_t i = 0; i < p.inputs.size(); ++i)
                if ((int)i != bc.epilogue_product_input && g.ops().nodes[p.inputs[i].node_id].op.kind != OpKind::Input) internal = (int)i;
            const int64_t out_slot = km_slot[km_next];
            body << "        for (int ch = grp; ch < " << k16 / 16 << "; ch += 4) {  // activation backward -> dz_" << l - 1 << " (over x_" << l << "'s MN-major copy)\n";
            body << "            const int c0 = ch * 16;\n            float acc[16], o[16], hv[16];\n            dc_ld16(tlane + (unsigned)c0, acc);\n";
            body << "            #pragma unroll\n            for (int q = 0; q < 4; ++q) {\n";
            body << "                const float4 h = dc_lds4(sbase + " << mn_act[l] << "u + dc_mn(c0 + 4 * q, r));\n";
            body << "                hv[4 * q] = h.x; hv[4 * q + 1] = h.y; hv[4 * q + 2] = h.z; hv[4 * q + 3] = h.w;\n            }\n";
            emit_program(body, p, bc.epilogue_product_input, internal, K, "            ", K % 16 != 0);
            body << "                o[j] = t" << p.output_ops[0] << ";\n";
            body << "                } else o[j] = 0.f;\n                }\n            }\n";
            body << "            dc_wait(bar2, phase2);  // x_" << l << "'s MN-major copy is free (a completed phase answers at once on later chunks)\n";
            store_km(body, out_slot, up(K, 8));
            store_mn(body, mn_act[l]);
            column_sums(body, l - 1);
            body << "        }\n        phase2 ^= 1u;\n";
        } else if (has_dx) {
            body << "        for (int ch = grp; ch < " << k16 / 16 << "; ch += 4) {  // gradient with respect to x_0 -> global memory\n";
            body << "            const int c0 = ch * 16;\n            float acc[16];\n            dc_ld16(tlane + (unsigned)c0, acc);\n";
            body << "            if (row_ok) {\n";
            if (K % 4 == 0) {
                body << "                #pragma unroll\n                for (int q = 0; q < 4; ++q)\n                    if (c0 + 4 * q < " << K << ")\n";
                body << "                        *reinterpret_cast<float4*>(" << pdx << " + (size_t)gm * " << K << " + c0 + 4 * q) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);\n";
            } else {
                body << "                #pragma unroll\n                for (int j = 0; j < 16; ++j)\n                    if (c0 + j < " << K << ") " << pdx << "[(size_t)gm * " << K << " + c0 + j] = acc[j];\n";
            }
            body << "            }\n        }\n";
        }
        if (l == 0) body << "        dc_wait(bar2, phase2); phase2 ^= 1u;  // the next tile's x_0 overwrites operands of the last weight-gradient MMAs\n";
    }
    body << "        first_tile = false;\n";
    body << "        asm volatile(\"tcgen05.fence::before_thread_sync;\" ::: \"memory\");\n        __syncthreads();  // the next tile's x_0 overwrites operands of this tile's last MMAs (waited for above)\n";
    body << "    }\n";
    // ---- per-CTA partial results -> scratch ---------------------------------------------------------------------------------
    body << "    asm volatile(\"tcgen05.fence::after_thread_sync;\" ::: \"memory\");\n";
    for (int l = 0; l < L; ++l) {
        const int64_t K = W[l], N = W[l + 1];
        body << "    for (int ch = grp; ch < " << dw_cols[l] / 16 << "; ch += 4) {  // this CTA's partial dW_" << l << "\n";
        body << "        const int c0 = ch * 16;\n        float acc[16];\n        dc_ld16(tlane + " << dw_col[l] << "u + (unsigned)c0, acc);\n";
        body << "        float* dst = ws + " << ws_dw[l] << " + (size_t)blockIdx.x * " << K * N << ";\n";
        body << "        #pragma unroll\n        for (int j = 0; j < 16; ++j) {\n";
        if (dw_a_is_dz[l]) body << "            if (r < " << N << " && c0 + j < " << K << ") dst[(c0 + j) * " << N << " + r] = acc[j];  // lanes = n, columns = k\n";
        else body << "            if (r < " << K << " && c0 + j < " << N << ") dst[r * " << N << " + c0 + j] = acc[j];  // lanes = k, columns = n\n";
        body << "        }\n    }\n";
        if (ws_cs[l] >= 0) {
            const int64_t per_warp = div_round_up(up(N, 16) / 16, 4);
            for (int64_t i = 0; i < per_warp; ++i) {
                body << "    {\n        const float s = dc_colsum16(cs" << l << "_" << i << ", lane);\n        const int c = (" << 4 * i << " + grp) * 16 + dc_colsum16_column(lane);\n";
                body << "        if ((lane & 1) == 0 && c < " << N << ") ws[" << ws_cs[l] << " + (size_t)(blockIdx.x * 4 + quad) * " << N << " + c] = s;\n    }\n";
            }
        }
    }
    for (size_t s = 0; s < ch.sums.size(); ++s) {
        body << "    {\n        float v = lsum" << s << ";\n        #pragma unroll\n        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);\n";
        body << "        __syncthreads();\n        if (lane == 0) red[warp] = v;\n        __syncthreads();\n";
        body << "        if (tid == 0) {\n            float t = red[0];\n            #pragma unroll\n            for (int w = 1; w < 16; ++w) t += red[w];\n";
        body << "            ws[" << ws_sum[s] << " + blockIdx.x] = t;\n        }\n    }\n";
    }
    body << "    asm volatile(\"tcgen05.fence::before_thread_sync;\" ::: \"memory\");\n    __syncthreads();\n";
    body << "    if (warp == 0) {\n        asm volatile(\"tcgen05.fence::after_thread_sync;\" ::: \"memory\");\n";
    body << "        asm volatile(\"tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\" ::\"r\"(tmem), \"n\"(" << tmem_cols << ") : \"memory\");\n    }\n";

    os << "extern \"C\" __global__ void __launch_bounds__(512, 1) " << name << "(" << params.str() << "float* __restrict__ ws, const unsigned* dsc_step) {\n";
    os << "    (void)dsc_step;\n" << body.str() << "}\n";

    // ---- split sums: every per-CTA partial set -> its node, one launch -------------------------------------------------------
    struct Set { int64_t offset, count, splits; int node; };
    std::vector<Set> sets;
    for (int l = 0; l < L; ++l) {
        const Cluster& gc = clusters[ch.weight_gradient[l]];
        sets.push_back({ws_dw[l], W[l] * W[l + 1], grid, gc.outputs[0]});
        if (ws_cs[l] >= 0) sets.push_back({ws_cs[l], W[l + 1], grid * 4, gc.outputs[1]});
    }
    for (size_t s = 0; s < ch.sums.size(); ++s) sets.push_back({ws_sum[s], 1, grid, clusters[ch.sums[s].batch_reduce].outputs[0]});
    const std::string sname = name + "_sums";
    os << "// per-CTA partial weight gradients, bias gradients and loss sums of the dense chain -> their arrays: 32 outputs x 8 split\n"
          "// lanes per CTA, lane g adds partials g, g + 8, ... in ascending order, then the eight lane sums in lane order\n";
    os << "extern \"C\" __global__ void __launch_bounds__(256) " << sname << "(const float* ws";
    for (size_t i = 0; i < sets.size(); ++i) os << ", float* out" << i;
    os << ", const unsigned* dsc_step) {\n    (void)dsc_step;\n    __shared__ float red[8][32];\n    const unsigned tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;\n";
    os << "    const float* src; float* dst; unsigned count, splits, block;\n";
    int64_t blocks = 0;
    for (size_t i = 0; i < sets.size(); ++i) {
        const int64_t b = div_round_up(sets[i].count, 32);
        os << "    " << (i ? "else " : "") << "if (blockIdx.x < " << blocks + b << "u) { src = ws + " << sets[i].offset << "; dst = out" << i << "; count = " << sets[i].count
           << "u; splits = " << sets[i].splits << "u; block = blockIdx.x - " << blocks << "u; }\n";
        blocks += b;
    }
    os << "    else return;\n    const unsigned i = block * 32u + tx;\n    float part = 0.f;\n    if (i < count) {\n        #pragma unroll 4\n"
          "        for (unsigned s = ty; s < splits; s += 8u) part += src[(size_t)s * count + i];\n    }\n    red[ty][tx] = part;\n    __syncthreads();\n"
          "    if (ty != 0 || i >= count) return;\n    float acc = red[0][tx];\n    #pragma unroll\n    for (unsigned g2 = 1; g2 < 8u; ++g2) acc += red[g2][tx];\n    dst[i] = acc;\n}\n\n";

    out->source = os.str();
    out->scratch_bytes = ws * 4;
    KernelLaunch l;
    l.entry = name;
    l.grid_x = (uint32_t)grid;
    l.block = 512;
    l.smem = (uint32_t)smem_bytes;
    std::ostringstream label;
    label << "DenseChain (" << L << " layers:";
    for (int64_t w : W) label << " " << w;
    label << ") [" << M << "] forward + loss + backward";
    l.label = "TensorCore" + label.str();
    l.cluster = ci;
    l.covers = ch.all_clusters();
    l.args = args;
    l.args.push_back({KernelArg::Scratch, -1, 0});
    l.algorithmic_bytes = bytes;
    for (int k = 0; k < L; ++k) l.flops += 6.0 * (double)M * (double)W[k] * (double)W[k + 1];
    // what the chain's clusters would move as kernels of their own (for the traffic-based roofline of the step: the fusion
    // removes bytes, it does not make the remaining ones faster)
    for (int member : ch.all_clusters()) {
        const ClusterCode unfused = generate_cluster_code(g, member, opt);
        for (const auto& ul : unfused.launches) l.replaced_bytes += ul.algorithmic_bytes;
    }
    out->launches.push_back(l);
    KernelLaunch s;
    s.entry = sname;
    s.grid_x = (uint32_t)blocks;
    s.label = "SplitSum dense chain (" + num((int64_t)sets.size()) + " arrays)";
    s.cluster = ci;
    s.args.push_back({KernelArg::Scratch, -1, 0});
    for (const auto& set : sets) s.args.push_back({KernelArg::NodeBuffer, set.node, 0});
    out->launches.push_back(s);
    out->extra_reads = reads;
    for (const auto& set : sets) out->extra_writes.push_back(set.node);
    if (has_dx) out->extra_writes.push_back(clusters[ch.backward[0]].outputs[0]);
    return true;
}

namespace {
}  // namespace

int64_t eval_chain(const ViewChain& chain, int64_t e) {
    for (int vi = (int)chain.views.size() - 1; vi >= 0; --vi) {
        const View& v = chain.views[vi];
        auto ostr = v.output_shape.strides();
        auto istr = v.input_shape.strides();
        std::vector<int64_t> in_coord(v.input_offsets.begin(), v.input_offsets.end());
        for (int i = 0; i < v.output_shape.len(); ++i) {
            const auto& m = v.output_mapping[i];
            if (!m.is_source) continue;
            in_coord[m.axis] += m.step * ((e / ostr[i]) % v.output_shape[i]);
        }
        int64_t idx = 0;
        for (int a = 0; a < v.input_shape.len(); ++a)
            idx += std::min<int64_t>(std::max<int64_t>(in_coord[a], 0), v.input_shape[a] - 1) * istr[a];
        e = idx;
    }
    return e;
}

std::string kernel_prelude() {
    // integer parts are bit-exact restatements of kernel_common.glsl:205-216
    return R"(// descent-b200 JIT kernels (generated)
__device__ __forceinline__ unsigned dsc_pcg(unsigned v) {
    const unsigned state = v * 747796405u + 2891336453u;
    const unsigned word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}
// float(hash)/float(0xffffffffu): the divisor rounds to 2^32, so this is an exact scale of the RNE conversion
// One lane of a converged warp.  The thread that issues the tensor-core MMAs is chosen with elect.sync under a warp-uniform
// condition: the compiler then knows exactly one thread runs the branch and emits each MMA once; under `tid == 0` it must
// assume any subset of lanes and wraps every MMA in an elect / broadcast loop (measured in SASS: BRA.U.ANY per MMA).
__device__ __forceinline__ bool dsc_elect_one() {
    unsigned pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float dsc_rand(unsigned uid, unsigned index, unsigned seed) {
    const unsigned hash = dsc_pcg(dsc_pcg(index) + seed + uid);
    return __uint2float_rn(hash) * 2.3283064365386962890625e-10f;
}

)";
}

ClusterCode generate_cluster_code(const Graph& graph, int ci, const CodegenOptions& opt, PrologueRequest* prologue) {
    const Cluster& c = graph.clusters()[ci];
    struct Scoped {  // the request is visible to the generators of this one cluster only
        explicit Scoped(PrologueRequest* r) { g_prologue = r; }
        ~Scoped() { g_prologue = nullptr; }
    } scoped(c.kind == ClusterKind::MatMul || c.kind == ClusterKind::WindowsToImage ? prologue : nullptr);
    if (prologue) prologue->fused[0] = prologue->fused[1] = false;
    switch (c.kind) {
        case ClusterKind::PerElement: return gen_per_element(graph, c, ci, opt);
        case ClusterKind::Reduce: return gen_reduce(graph, c, ci, opt);
        case ClusterKind::Row: return gen_row(graph, c, ci, opt);
        case ClusterKind::MatMul: {
            ClusterCode code = gen_matmul(graph, c, ci, opt);
            if (c.pool.enabled && !code.pool_done) {
                // the GEMM kernel chosen for this shape does not pool its output: run the absorbed max-pool Reduce after it
                ClusterCode r = gen_reduce(graph, c.pool.reduce[0], ci, opt, "_pool");
                const int64_t base = div_round_up(code.scratch_bytes, 256) * 256;
                code.scratch_bytes = base + div_round_up(r.scratch_bytes, 256) * 256;
                code.source += r.source;
                for (auto& l : r.launches) {
                    for (auto& arg : l.args)
                        if (arg.kind == KernelArg::Scratch) arg.scratch_offset += base;
                    code.launches.push_back(l);
                }
            }
            if (prologue && prologue->fused[1] && !c.column_sum.empty() && !code.column_sum_done) {
                prologue->fused[1] = false;  // the absorbed column sums would read B from memory, where it no longer exists
                return code;
            }
            if (!c.column_sum.empty() && !code.column_sum_done) {
                // the GEMM kernel chosen for this shape does not produce the column sums: run the absorbed Reduce chain
                // after it, intermediate results in scratch
                std::map<int, int64_t> scratch_of;  // intermediate reduce node -> scratch offset
                for (size_t i = 0; i < c.column_sum.size(); ++i) {
                    const Cluster& rc = c.column_sum[i];
                    ClusterCode r = gen_reduce(graph, rc, ci, opt, "_colsum" + num((int64_t)i));
                    const int64_t base = div_round_up(code.scratch_bytes, 256) * 256;
                    code.scratch_bytes = base + div_round_up(r.scratch_bytes, 256) * 256;
                    if (i + 1 < c.column_sum.size()) {
                        scratch_of[rc.outputs[0]] = code.scratch_bytes;
                        code.scratch_bytes += div_round_up(graph.ops().nodes[rc.outputs[0]].shape.element_count() * 4, 256) * 256;
                    }
                    code.source += r.source;
                    for (auto& l : r.launches) {
                        for (auto& arg : l.args) {
                            if (arg.kind == KernelArg::Scratch) arg.scratch_offset += base;
                            else if (scratch_of.count(arg.node_id)) arg = KernelArg{KernelArg::Scratch, -1, scratch_of[arg.node_id]};
                        }
                        if (l.kind == KernelLaunch::ZeroScratch) l.zero_offset += base;
                        code.launches.push_back(l);
                    }
                }
            }
            return code;
        }
        case ClusterKind::Unpad: return gen_unpad(graph, c, ci);
        case ClusterKind::WindowsToImage: return gen_w2i(graph, c, ci, opt);
        case ClusterKind::ScatterAdd: return gen_scatter_add(graph, c, ci, opt);
        case ClusterKind::AllReduce: {
            ClusterCode code;
            KernelLaunch l;
            l.kind = KernelLaunch::AllReduce;
            l.label = c.label;
            l.cluster = ci;
            l.args = {{KernelArg::NodeBuffer, c.outputs[0], 0}};
            code.launches.push_back(l);
            return code;
        }
    }
    fail("unknown cluster kind");
}

}  // namespace descent

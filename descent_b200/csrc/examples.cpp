// Example networks (reference: examples/fashion_mnist/main.rs:128-323, examples/image_fit/main.rs:50-350).
#include "examples.hpp"

#include <cmath>

namespace descent {

namespace {

// ---- fashion_mnist --------------------------------------------------------------------------------

struct Linear : Module {  // main.rs:128-144
    Dense fc;
    explicit Linear(Environment& env) : fc(Dense::builder(28 * 28, 10).build(env)) {}
    DualArray eval(DualArray input, const EvalContext& ctx) const override { return apply(input.flatten(), fc, ctx); }
};

struct SingleLayer : Module {  // main.rs:146-169 (+ optional dropout before fc1: BASELINE config 2)
    Dense fc1, fc2;
    bool dropout;
    SingleLayer(Environment& env, bool dropout)
        : fc1(Dense::builder(28 * 28, 300).build(env)), fc2(Dense::builder(300, 10).build(env)), dropout(dropout) {}
    DualArray eval(DualArray input, const EvalContext& ctx) const override {
        DualArray x = input.flatten();
        if (dropout) x = apply(x, Dropout(0.5f), ctx);
        return apply(apply(x, fc1, ctx).leaky_relu(0.01f), fc2, ctx);
    }
};

struct ConvNet : Module {  // main.rs:171-222
    Conv2D conv1;
    std::unique_ptr<Module> pool1;
    Conv2D conv2;
    std::unique_ptr<Module> pool2;
    Dense fc1, fc2;
    static std::unique_ptr<Module> make_pool(Environment& env, bool blur, int64_t channels) {
        if (blur) return std::make_unique<MaxBlurPool2D>(env, channels);
        return std::make_unique<MaxPool2D>();
    }
    ConvNet(Environment& env, bool use_blur_pool)
        : conv1(Conv2D::builder(1, 16, 3, 3).with_pad(1).build(env)), pool1(make_pool(env, use_blur_pool, 16)),
          conv2(Conv2D::builder(16, 32, 3, 3).with_pad(1).with_groups(2).build(env)), pool2(make_pool(env, use_blur_pool, 32)),
          fc1(Dense::builder(7 * 7 * 32, 128).build(env)), fc2(Dense::builder(128, 10).build(env)) {}
    DualArray eval(DualArray input, const EvalContext& ctx) const override {
        DualArray x = apply(input, conv1, ctx).leaky_relu(0.01f);
        x = apply(x, *pool1, ctx);
        x = apply(x, conv2, ctx).leaky_relu(0.01f);
        x = apply(x, *pool2, ctx);
        x = apply(x.flatten(), Dropout(0.5f), ctx);
        x = apply(x, fc1, ctx).leaky_relu(0.01f);
        return apply(x, fc2, ctx);
    }
};

// ---- image_fit ------------------------------------------------------------------------------------

DualArray positional_encoding(DualArray x, int64_t freq_count) {  // main.rs:259-275
    Scope* scope = x.scope();
    const float pi = 3.14159265358979323846f;
    DualArray freq = scope->literal(2.0f).pow(scope->coord(freq_count)) * pi;
    DualArray phase = scope->coord(2).reshape({2, 1}) * 0.5f * pi;
    Shape shape = x.shape();
    Shape calc_shape = shape.concat(Shape{1, 1});
    Shape output_shape = shape;
    output_shape[output_shape.len() - 1] *= 2 * freq_count;
    return (x.reshape(calc_shape) * freq + phase).sin().reshape(output_shape);
}

struct Relu : Module {  // main.rs:50-84
    int64_t freq_count;
    std::vector<Dense> hidden_layers;
    Dense final_layer;
    Relu(Environment& env, int64_t freq_count, const std::vector<int64_t>& hidden_units) : freq_count(freq_count) {
        int64_t prev = freq_count == 0 ? 2 : 4 * freq_count;
        for (int64_t h : hidden_units) {
            hidden_layers.push_back(Dense::builder(prev, h).build(env));
            prev = h;
        }
        final_layer = Dense::builder(prev, 3).build(env);
    }
    DualArray eval(DualArray input, const EvalContext& ctx) const override {
        DualArray x = input;
        if (freq_count != 0) x = positional_encoding(input, freq_count);
        for (const auto& layer : hidden_layers) x = apply(x, layer, ctx).leaky_relu(0.01f);
        return apply(x, final_layer, ctx);
    }
};

struct Siren : Module {  // main.rs:86-118
    std::vector<Dense> hidden_layers;
    Dense final_layer;
    Siren(Environment& env, const std::vector<int64_t>& hidden_units) {
        int64_t prev = 2;
        for (size_t i = 0; i < hidden_units.size(); ++i) {
            hidden_layers.push_back(Dense::builder(prev, hidden_units[i])
                                        .with_w_initializer(Initializer::for_siren(prev, i == 0))
                                        .with_b_initializer(Initializer::rand_uniform(1.0f))
                                        .build(env));
            prev = hidden_units[i];
        }
        final_layer = Dense::builder(prev, 3).build(env);
    }
    DualArray eval(DualArray input, const EvalContext& ctx) const override {
        DualArray x = input;
        for (const auto& layer : hidden_layers) x = apply(x, layer, ctx).sin();
        return apply(x, final_layer, ctx);
    }
};

struct HashGrid : Module {  // main.rs:120-204
    int64_t grid_size;
    int64_t stride;
    Parameter t;
    HashGrid(Environment& env, int64_t grid_size, int64_t entry_count, int64_t values_per_entry) : grid_size(grid_size) {
        const int64_t grid_point_count = grid_size + 1;
        const int64_t max_entry_count = grid_point_count * grid_point_count;
        entry_count = std::min(entry_count, max_entry_count);
        stride = entry_count == max_entry_count ? grid_point_count : 1526263;  // large prime
        t = env.trainable_parameter(Shape{entry_count, values_per_entry}, "t", Initializer::rand_uniform(1.0e-4f));
    }
    DualArray eval(DualArray input, const EvalContext&) const override {
        Scope* scope = input.scope();
        Array x = input.next_colour().value();  // the input gradient is discarded (main.rs:157)
        auto [tv, dt] = scope->parameter(t).into_inner();
        const uint32_t entry_count = (uint32_t)tv.shape()[0];
        const uint32_t s = (uint32_t)stride;

        Array cf = (x * 0.5f + 0.5f) * (float)grid_size;
        UArray c = cf.into_u32();
        Array f = cf - c.into_f32();

        UArray c0 = c.lock_axis(-1, 0, false), c1 = c.lock_axis(-1, 1, false);
        Array f0 = f.lock_axis(-1, 0, true), f1 = f.lock_axis(-1, 1, true);

        UArray ia = ((c0 + 0u) ^ (c1 * s + 0u)) % entry_count;
        UArray ib = ((c0 + 1u) ^ (c1 * s + 0u)) % entry_count;
        UArray ic = ((c0 + 0u) ^ (c1 * s + s)) % entry_count;
        UArray id = ((c0 + 1u) ^ (c1 * s + s)) % entry_count;

        Array ta = tv.gather(-2, ia), tb = tv.gather(-2, ib), tc = tv.gather(-2, ic), td = tv.gather(-2, id);
        Array g0 = 1.0f - f0, g1 = 1.0f - f1;
        Array wa = g0 * g1, wb = f0 * g1, wc = g0 * f1, wd = f0 * f1;

        auto [y, dy] = (ta * wa + tb * wb + tc * wc + td * wd).with_empty_grad();
        dt.accumulate(scope->literal(0.0f)
                          .value()
                          .broadcast(dt.shape())
                          .scatter_add(dy * wa, -2, ia)
                          .scatter_add(dy * wb, -2, ib)
                          .scatter_add(dy * wc, -2, ic)
                          .scatter_add(dy * wd, -2, id));
        return {y, dy};
    }
};

struct MultiHashGrid : Module {  // main.rs:206-257
    std::vector<HashGrid> grids;
    std::vector<Dense> hidden_layers;
    Dense final_layer;
    MultiHashGrid(Environment& env, int64_t min_grid, int64_t max_grid, int64_t level_count, int64_t entry_count,
                  const std::vector<int64_t>& hidden_units) {
        const int64_t values_per_entry = 2;
        const float b = std::exp((std::log((float)max_grid) - std::log((float)min_grid)) / (float)(level_count - 1));
        for (int64_t level = 0; level < level_count; ++level) {
            float p = 1.0f;  // b.powi(level): repeated f32 multiplication
            for (int64_t i = 0; i < level; ++i) p *= b;
            const int64_t grid_size = (int64_t)((float)min_grid * p);
            grids.emplace_back(env, grid_size, entry_count, values_per_entry);
        }
        int64_t prev = (int64_t)grids.size() * values_per_entry;
        for (int64_t h : hidden_units) {
            hidden_layers.push_back(Dense::builder(prev, h).build(env));
            prev = h;
        }
        final_layer = Dense::builder(prev, 3).build(env);
    }
    DualArray eval(DualArray input, const EvalContext& ctx) const override {
        DualArray x = grids[0].eval(input, ctx);
        for (size_t i = 1; i < grids.size(); ++i) x = x.concat(grids[i].eval(input, ctx), -1);
        for (const auto& layer : hidden_layers) x = layer.eval(x, ctx).leaky_relu(0.01f);
        return final_layer.eval(x, ctx);
    }
};

std::unique_ptr<Optimizer> make_optimizer(Environment& env, Scope& scope, const std::vector<Parameter>& parameters, const std::string& kind,
                                          const Array& lr_scale, float adam_lr, float beta2) {
    if (kind == "descent") return std::make_unique<StochasticGradientDescent>(env, scope, parameters, 0.1f * lr_scale, 0.9f);
    DSC_CHECK(kind == "adam", "unknown optimizer '" << kind << "'");
    return std::make_unique<Adam>(env, scope, parameters, adam_lr * lr_scale, 0.9f, beta2, 1.0e-8f);
}

}  // namespace

std::unique_ptr<Example> build_fashion_mnist(Environment& env, const ExampleConfig& cfg) {
    auto ex = std::make_unique<Example>();
    ex->family = "fashion_mnist";
    if (cfg.network == "linear") ex->module = std::make_unique<Linear>(env);
    else if (cfg.network == "single-layer") ex->module = std::make_unique<SingleLayer>(env, false);
    else if (cfg.network == "single-layer-dropout") ex->module = std::make_unique<SingleLayer>(env, true);
    else if (cfg.network == "conv-net") ex->module = std::make_unique<ConvNet>(env, false);
    else if (cfg.network == "conv-blur-net") ex->module = std::make_unique<ConvNet>(env, true);
    else fail("unknown fashion_mnist network '" + cfg.network + "'");

    const int64_t m = cfg.mini_batch_size;
    ex->x = env.static_parameter(Shape{m, 28, 28, 1}, "x");
    ex->y = env.static_parameter(Shape{m, 1}, "y");
    ex->learning_rate_scale = env.static_parameter(Shape{1}, "lr_scale");
    ex->loss_sum = env.static_parameter(Shape{1}, "loss");
    ex->accuracy_sum = env.static_parameter(Shape{1}, "accuracy");

    auto emit = [&](Scope& scope, bool training) {  // main.rs:247-261 / 295-309
        DualArray x = training ? ex->module->train(scope.parameter(ex->x)) : ex->module->test(scope.parameter(ex->x));
        Array loss = softmax_cross_entropy_loss(x, ex->y).set_loss();
        Array accuracy = softmax_cross_entropy_accuracy(x, ex->y);
        scope.update_parameter_value(ex->loss_sum, [&](Array s) { return s + loss.reduce_sum(0, false); });
        scope.update_parameter_value(ex->accuracy_sum, [&](Array s) { return s + accuracy.reduce_sum(0, false); });
    };
    {
        auto scope = env.scope();
        emit(*scope, true);
        Array lr_scale = scope->parameter_value(ex->learning_rate_scale);
        ex->parameters = scope->trainable_parameters();
        add_weight_decay_to_grad(*scope, ex->parameters, cfg.weight_decay);
        ex->optimizer = make_optimizer(env, *scope, ex->parameters, cfg.optimizer, lr_scale, 0.005f, 0.999f);
        ex->train_graph_json = scope->export_json();
        ex->train_graph.reset(scope->build_graph());
    }
    {
        auto scope = env.scope();
        emit(*scope, false);
        ex->test_graph_json = scope->export_json();
        ex->test_graph.reset(scope->build_graph());
    }
    return ex;
}

std::unique_ptr<Example> build_image_fit(Environment& env, const ExampleConfig& cfg) {
    auto ex = std::make_unique<Example>();
    ex->family = "image_fit";
    const std::vector<int64_t> hidden = {256, 128, 64, 32};
    if (cfg.network == "relu") ex->module = std::make_unique<Relu>(env, 0, hidden);
    else if (cfg.network == "relu-pe") ex->module = std::make_unique<Relu>(env, 8, hidden);
    else if (cfg.network == "siren") ex->module = std::make_unique<Siren>(env, hidden);
    else if (cfg.network == "multi-hash") ex->module = std::make_unique<MultiHashGrid>(env, 2, 512, 10, 4096, std::vector<int64_t>{64, 64});
    else fail("unknown image_fit network '" + cfg.network + "'");

    const int64_t m = cfg.mini_batch_size;
    ex->x = env.static_parameter(Shape{m, 2}, "x");
    ex->y = env.static_parameter(Shape{m, 3}, "y");
    ex->learning_rate_scale = env.static_parameter(Shape{1}, "lr_scale");
    ex->loss_sum = env.static_parameter(Shape{1}, "loss");
    {
        auto scope = env.scope();  // main.rs:308-330
        DualArray x = ex->module->train(scope->parameter(ex->x));
        Array loss = (x - ex->y).square().reduce_sum(-1, true).set_loss();
        scope->update_parameter_value(ex->loss_sum, [&](Array s) { return s + loss.reduce_sum(0, false); });
        Array lr_scale = scope->parameter_value(ex->learning_rate_scale);
        ex->parameters = scope->trainable_parameters();
        scope->all_reduce_gradients(ex->parameters);
        // main.rs:319 always uses Adam(0.02, 0.9, 0.99); "descent" is accepted so that the parity tests can hold the parameters
        // of these networks to 1e-5 after an SGD step as well (Adam's first step is sign-like in near-zero gradients)
        ex->optimizer = make_optimizer(env, *scope, ex->parameters, cfg.optimizer, lr_scale, 0.02f, 0.99f);
        ex->train_graph_json = scope->export_json();
        ex->train_graph.reset(scope->build_graph());
    }
    if (cfg.image_width > 0 && cfg.image_height > 0) {
        const int64_t width = cfg.image_width, height = cfg.image_height, pixel_count = width * height;
        ex->image = env.static_parameter(Shape{pixel_count, 3}, "image");
        auto scope = env.scope();  // main.rs:339-350
        DualArray u = (scope->coord(width) + 0.5f) * (2.0f / (float)width) - 1.0f;
        DualArray v = (scope->coord(height) + 0.5f) * (2.0f / (float)height) - 1.0f;
        DualArray x = scope->coord(2).select_eq(0.0f, u.reshape({1, width, 1}), v.reshape({height, 1, 1})).reshape({pixel_count, 2});
        x = ex->module->test(x);
        scope->write_parameter_value(ex->image, x.value());
        ex->test_graph_json = scope->export_json();
        ex->test_graph.reset(scope->build_graph());
    }
    return ex;
}

// examples/sentiment/main.rs:119-170: word indices -> one-hot -> embedding matmul -> LSTM over the sentence -> dense 3
// -> softmax cross-entropy, Adam(0.002).  The data pipeline (DynaSent jsonl, vocabulary) is out of scope; the
// vocabulary size and sentence length are configuration.
namespace {
class Sentiment : public Module {
public:
    Sentiment(Environment& env, int64_t vocab, int64_t words, int64_t embedding_size, int64_t lstm_size)
        : vocab_(vocab), words_(words), embedding_size_(embedding_size), lstm_(env, embedding_size, lstm_size),
          fc_(Dense::builder(lstm_size, 3).build(env)),
          embedding_(env.trainable_parameter(Shape{vocab, embedding_size}, "em", Initializer::rand_uniform(1.0f))) {}
    DualArray eval(DualArray input, const EvalContext& ctx) const override {
        const int64_t m = input.shape()[0];
        DualArray x = DualArray(input.value().one_hot(vocab_).with_empty_grad());
        x = x.reshape(Shape{m * words_, vocab_}).matmul(embedding_).reshape(Shape{m, words_, embedding_size_});
        return fc_.eval(lstm_.eval(x, ctx), ctx);
    }

private:
    int64_t vocab_, words_, embedding_size_;
    LSTMCell lstm_;
    Dense fc_;
    Parameter embedding_;
};
}  // namespace

std::unique_ptr<Example> build_sentiment(Environment& env, const ExampleConfig& cfg) {
    auto ex = std::make_unique<Example>();
    ex->family = "sentiment";
    const int64_t m = cfg.mini_batch_size, vocab = cfg.image_width > 0 ? cfg.image_width : 4096, words = cfg.image_height > 0 ? cfg.image_height : 32;
    ex->module = std::make_unique<Sentiment>(env, vocab, words, 128, 64);
    ex->x = env.static_parameter(Shape{m, words, 1}, "x");
    ex->y = env.static_parameter(Shape{m, 1}, "y");
    ex->learning_rate_scale = env.static_parameter(Shape{1}, "lr_scale");
    ex->loss_sum = env.static_parameter(Shape{1}, "loss");
    ex->accuracy_sum = env.static_parameter(Shape{1}, "accuracy");
    auto scope = env.scope();
    DualArray x = ex->module->train(scope->parameter(ex->x));
    Array loss = softmax_cross_entropy_loss(x, ex->y).set_loss();
    Array accuracy = softmax_cross_entropy_accuracy(x, ex->y);
    scope->update_parameter_value(ex->loss_sum, [&](Array s) { return s + loss.reduce_sum(0, false); });
    scope->update_parameter_value(ex->accuracy_sum, [&](Array s) { return s + accuracy.reduce_sum(0, false); });
    Array lr_scale = scope->parameter_value(ex->learning_rate_scale);
    ex->parameters = scope->trainable_parameters();
    scope->all_reduce_gradients(ex->parameters);
    ex->optimizer = make_optimizer(env, *scope, ex->parameters, cfg.optimizer, lr_scale, 0.002f, 0.999f);
    ex->train_graph_json = scope->export_json();
    ex->train_graph.reset(scope->build_graph());
    return ex;
}

std::unique_ptr<Example> build_example(Environment& env, const ExampleConfig& config) {
    for (const char* n : {"relu", "relu-pe", "siren", "multi-hash"})
        if (config.network == n) return build_image_fit(env, config);
    if (config.network == "sentiment") return build_sentiment(env, config);
    return build_fashion_mnist(env, config);
}

}  // namespace descent

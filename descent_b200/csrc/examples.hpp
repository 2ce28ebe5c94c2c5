// The reference's example networks and their training/test graphs, restated against the C++ API:
// examples/fashion_mnist/main.rs (linear, single-layer, conv-net, conv-blur-net) and
// examples/image_fit/main.rs (relu, relu-pe, siren, multi-hash).  Dataset/JPEG front ends are out of
// scope: callers feed x / y batches.
#pragma once
#include "module.hpp"

namespace descent {

struct ExampleConfig {
    std::string network;            // fashion_mnist: linear | single-layer | single-layer-dropout | conv-net | conv-blur-net
                                    // image_fit: relu | relu-pe | siren | multi-hash
                                    // sentiment: sentiment (image_width = vocabulary size, image_height = words per sentence)
    int64_t mini_batch_size = 1000; // per rank
    std::string optimizer = "adam"; // adam | descent
    float weight_decay = 1.0e-8f;   // fashion_mnist default (main.rs:103-104); image_fit uses none
    int64_t image_width = 0, image_height = 0;  // image_fit test graph (0 = no test graph)
};

struct Example {
    std::string family;  // "fashion_mnist" | "image_fit" | "sentiment"
    std::unique_ptr<Module> module;
    std::vector<std::unique_ptr<Module>> owned;  // sub-modules kept alive
    Parameter x, y, learning_rate_scale, loss_sum, accuracy_sum, image;
    std::vector<Parameter> parameters;  // trainable, in first-use order (array.rs:1419-1432)
    std::unique_ptr<Optimizer> optimizer;
    std::unique_ptr<Graph> train_graph, test_graph;
    std::string train_graph_json;  // raw op graph of the training step, for the oracle
    std::string test_graph_json;
};

std::unique_ptr<Example> build_fashion_mnist(Environment& env, const ExampleConfig& config);
std::unique_ptr<Example> build_image_fit(Environment& env, const ExampleConfig& config);
std::unique_ptr<Example> build_sentiment(Environment& env, const ExampleConfig& config);
std::unique_ptr<Example> build_example(Environment& env, const ExampleConfig& config);

}  // namespace descent

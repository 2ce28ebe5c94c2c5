/* descent_api.h -- C ABI mirror of descent's public host API (Environment / Scope / Array /
 * DualArray / Module / Optimizer; reference src/environment.rs, src/array.rs, src/module.rs,
 * src/loss.rs, src/optimizer.rs), for callers that cannot link the C++ classes in descent_b200/csrc
 * directly (the Python tests and bench.py bind it with ctypes).  The device boundary a Rust build of
 * the reference would bind is include/descent_cuda.h; this header sits above it.
 *
 * Conventions: every function returns 0 on success (DSC_ERR_* otherwise, message in
 * dsc_last_error()).  Arrays are graph-node handles (int) valid inside one scope; a DualArray is a
 * (value, loss_grad) pair of handles.  Shapes are int64 arrays of at most 7 extents (shape.rs:10).
 */
#ifndef DESCENT_API_H
#define DESCENT_API_H

#include "descent_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dsc_env dsc_env;        /* Environment        environment.rs:85-124 */
typedef struct dsc_scope dsc_scope;    /* Scope              array.rs:1231 */
typedef struct dsc_graphdef dsc_graphdef; /* Graph           graph.rs:102-139 */

/* initializer kinds for trainable parameters (parameter.rs:9-14) */
#define DSC_INIT_ZERO 0
#define DSC_INIT_RAND_NORMAL 1
#define DSC_INIT_RAND_UNIFORM 2

void dsc_string_free(char* s);

/* ---- Environment ---------------------------------------------------------------------------- */
/* device >= 0: a CUDA device.  device == -1: a host-only environment that can declare parameters, build
 * graphs and emit/compile kernel source (build check, CPU tests); writes are discarded and any read or run
 * call fails -- there is no CPU execution path. */
int dsc_env_create(int device, dsc_env** out);                                    /* Environment::new :104 */
int dsc_env_set_data_parallel_for_tracing(dsc_env* env, int world, int rank);     /* host-only envs: build rank-specific graphs */
int dsc_env_destroy(dsc_env* env);                                                /* Drop :523 */
int dsc_env_ctx(dsc_env* env, dsc_ctx** ctx);
int dsc_env_static_parameter(dsc_env* env, const int64_t* shape, int ndim, const char* name, int* param);               /* :141 */
int dsc_env_trainable_parameter(dsc_env* env, const int64_t* shape, int ndim, const char* name, int init_kind, float init_scale, int* param); /* :149 */
int dsc_env_parameter_count(dsc_env* env, int* count);
int dsc_env_parameter_info(dsc_env* env, int param, int64_t* shape7, int* ndim, char* name64, int* trainable);
int dsc_env_write_parameter(dsc_env* env, int param, const float* data, size_t count, int data_is_pinned);  /* writer :160, zero-fills the tail */
/* Asynchronous writer for the next mini-batch: `pinned_data` (from dsc_host_alloc, `count` = the whole parameter)
 * is copied on a second stream while earlier runs execute and becomes the parameter's contents at the next
 * dsc_env_run / read / write.  Plays the role of the reference's staging ring (staging.rs:100-140), which lets
 * the host fill batch i+1 while the GPU consumes batch i. */
int dsc_env_prefetch_parameter(dsc_env* env, int param, const float* pinned_data, size_t count);
int dsc_env_read_parameter(dsc_env* env, int param, float* dst, size_t count);                               /* reader :175 */
int dsc_env_reset_parameter(dsc_env* env, int param, uint64_t* rng_state);                                   /* reset_parameter :190 */
int dsc_env_scope(dsc_env* env, dsc_scope** out);                                                            /* scope :231 */
int dsc_scope_destroy(dsc_scope* scope);
int dsc_scope_build_graph(dsc_scope* scope, dsc_graphdef** out);                                             /* Scope::build_graph array.rs:1434 */
int dsc_graphdef_destroy(dsc_graphdef* graph);
int dsc_env_run(dsc_env* env, dsc_graphdef* graph, uint32_t rand_seed);                                      /* run :326 */
int dsc_env_sync(dsc_env* env);
int dsc_env_set_options(dsc_env* env, int use_cuda_graph, int profile_runs);
/* 0 (default): all GEMMs strict FP32.  1: MatMuls / convolutions run on tcgen05 with TF32 operands.  A graph that was
 * planned under the other setting is planned again at its next run. */
int dsc_env_set_tf32(dsc_env* env, int on);
/* Plan graphs as if the device had `sm_count` SMs (0 = the real count): persistent kernels then walk many tiles per CTA
 * on small problems, which is how the parity tests reach the regime of the large-batch benchmark. */
int dsc_env_set_sm_count(dsc_env* env, int sm_count);
int dsc_env_print_timings(dsc_env* env, const char* label);                                                  /* print_timings :518 */
int dsc_env_init_data_parallel(dsc_env* env, int world, int rank, const void* nccl_unique_id128);            /* new: SURVEY.md 8e */
/* JSON: [{"label","entry","cluster","ms","bytes","flops"}...] per launch of one run, averaged over `iterations` eager runs */
int dsc_env_profile(dsc_env* env, dsc_graphdef* graph, uint32_t rand_seed, int iterations, char** json_out);
/* JSON: {"kernel_launches","total_nodes","arena_bytes","algorithmic_bytes","flops","jit_ms"} */
int dsc_env_graph_stats(dsc_env* env, dsc_graphdef* graph, char** json_out);

/* Graph introspection that needs no device */
int dsc_scope_export_json(dsc_scope* scope, char** json_out);       /* raw op graph: what the oracle interprets */
int dsc_graphdef_export_json(dsc_graphdef* graph, char** json_out); /* optimised graph + clusters */
int dsc_graphdef_kernel_source(dsc_graphdef* graph, int sm_count, int dp_rank, char** cuda_source_out);
int dsc_graphdef_kernel_source_ex(dsc_graphdef* graph, int sm_count, int dp_rank, int use_tf32, char** cuda_source_out);
int dsc_graphdef_write_dot_file(dsc_graphdef* graph, int mode /*0 none,1 cluster,2 colour*/, const char* path); /* graph.rs:656 */

/* ---- Scope (array.rs:1257-1441) ------------------------------------------------------------- */
int dsc_scope_literal(dsc_scope* s, float value, int* value_node, int* grad_node);
int dsc_scope_literal_u32(dsc_scope* s, uint32_t value, int* node);
int dsc_scope_coord(dsc_scope* s, int64_t len, int* value_node, int* grad_node);
int dsc_scope_rand(dsc_scope* s, const int64_t* shape, int ndim, int* value_node, int* grad_node);
int dsc_scope_parameter(dsc_scope* s, int param, int* value_node, int* grad_node);
int dsc_scope_parameter_value(dsc_scope* s, int param, int* node);
int dsc_scope_write_parameter_value(dsc_scope* s, int param, int node);
int dsc_scope_accumulator(dsc_scope* s, const int64_t* shape, int ndim, int* node);
int dsc_scope_next_colour(dsc_scope* s);
int dsc_scope_trainable_parameters(dsc_scope* s, int* params, int capacity, int* count);
int dsc_scope_all_reduce_gradients(dsc_scope* s, const int* params, int count);
int dsc_array_shape(dsc_scope* s, int node, int64_t* shape7, int* ndim);

/* ---- Array / UArray / DualArray ops ---------------------------------------------------------
 * One dispatcher for the whole operator surface.  `op` is the reference method name; DualArray
 * methods are prefixed "dual.".  nodes: Array operands (DualArray operands take two slots: value,
 * grad).  iargs / fargs: the integer / float arguments in declaration order.  Outputs: one handle
 * (Array/UArray) or two (DualArray).  Table (array.rs line of the method):
 *   Array/UArray : broadcast[iargs=shape] :124  limit_axis[axis,start,end] :178  lock_axis[axis,coord,keep] :188
 *                  reshape[shape] :194  transpose :213
 *   Array        : add sub mul div :687-766 (2 nodes)  neg :768  concat[axis] :299  one_hot[count] :347
 *                  reduce_max[axis,keep] :351  reduce_sum[axis,keep] :356  argmax[axis,keep] :362  coord[axis] :369
 *                  gather[axis] (values, indices) :376  scatter_add[axis] (acc, values, indices) :398
 *                  select_eq select_gt (a, b, pass, fail) :427-442  square sqrt exp log sin cos :444-461
 *                  into_u32 :468  sigmoid :471  tanh :474  pow (a, b) :480  matmul (a, b) :492
 *                  accumulate (target, src; no output) :617  pad_image[pad] :551  unpad_image[pad] :555
 *   UArray       : uadd umul urem uxor (2 nodes) :734,775-792  into_f32 :682
 *   DualArray    : dual.add dual.sub dual.mul :1149-1214  dual.square dual.sin dual.tanh dual.sigmoid :829-858
 *                  dual.leaky_relu[fargs=leakiness] :860  dual.matmul :880  dual.transpose :888  dual.pow :897
 *                  dual.select_eq (a, b, pass, fail as duals) :909  dual.lock_axis[axis,coord,keep] :937
 *                  dual.reshape[shape] :942  dual.conv2d[pad,stride_w,stride_h] (x, filter) :989
 *                  dual.max_pool2d[filter_w,filter_h,stride_w,stride_h] :1033  dual.reduce_sum dual.reduce_max[axis,keep] :1087-1096
 *                  dual.flatten :1098  dual.set_loss (-> Array) :1106  dual.concat[axis] :1131
 */
int dsc_array_op(dsc_scope* s, const char* op, const int* nodes, int num_nodes, const int64_t* iargs, int num_iargs, const float* fargs,
                 int num_fargs, int* out_nodes, int* num_out);

/* ---- Modules (module.rs), loss (loss.rs), optimisers (optimizer.rs) -------------------------- */
int dsc_module_dense(dsc_env* env, int64_t input, int64_t output, int w_init_kind, float w_init_scale, int b_init_kind, float b_init_scale, int* module);
int dsc_module_conv2d(dsc_env* env, int64_t ic, int64_t oc, int64_t filter_w, int64_t filter_h, int64_t pad, int64_t stride_w, int64_t stride_h,
                      int64_t groups, int is_blur, int* module);
int dsc_module_max_pool2d(dsc_env* env, int* module);
int dsc_module_max_blur_pool2d(dsc_env* env, int64_t channels, int* module);
int dsc_module_dropout(dsc_env* env, float amount, int* module);
int dsc_module_lstm_cell(dsc_env* env, int64_t input, int64_t output, int* module);
int dsc_module_eval(dsc_env* env, dsc_scope* s, int module, int value_node, int grad_node, int is_training, int* out_value, int* out_grad);
int dsc_softmax_cross_entropy_loss(dsc_scope* s, int z_value, int z_grad, int y_node, int* loss_value, int* loss_grad);
int dsc_softmax_cross_entropy_accuracy(dsc_scope* s, int z_value, int z_grad, int y_node, int* node);
int dsc_add_weight_decay_to_grad(dsc_scope* s, const int* params, int count, float weight_decay);
int dsc_optimizer_sgd(dsc_env* env, dsc_scope* s, const int* params, int count, int learning_rate_node, float momentum, int* optimizer);
int dsc_optimizer_adam(dsc_env* env, dsc_scope* s, const int* params, int count, int learning_rate_node, float beta1, float beta2, float epsilon, int* optimizer);
int dsc_optimizer_reset_state(dsc_env* env, int optimizer);
int dsc_optimizer_state(dsc_env* env, int optimizer, int* params, int capacity, int* count);

/* ---- Example networks (examples/fashion_mnist, examples/image_fit) --------------------------- */
typedef struct dsc_example {
    int x, y, learning_rate_scale, loss_sum, accuracy_sum, image; /* parameter ids (-1 when absent) */
    int num_parameters;                                          /* trainable parameters, first-use order */
    int parameters[64];
    int num_optimizer_state;
    int optimizer_state[130];
    dsc_graphdef* train_graph;
    dsc_graphdef* test_graph; /* NULL when absent */
} dsc_example;
int dsc_example_create(dsc_env* env, const char* network, int64_t mini_batch_size, const char* optimizer, float weight_decay,
                       int64_t image_width, int64_t image_height, dsc_example* out);
int dsc_example_graph_json(dsc_env* env, int which /*0 train, 1 test*/, char** json_out); /* raw graph of the last example created */

/* ---- Host random numbers of the reference's examples -----------------------------------------
 * rand_chacha::ChaCha20Rng::seed_from_u64 and the rand 0.8 sampling rules the examples use (descent_b200/csrc/host_rng.hpp):
 * examples/fashion_mnist/main.rs:362-386, examples/image_fit/main.rs:352-395, examples/sentiment/main.rs:164. */
typedef struct dsc_rng dsc_rng;
int dsc_rng_create(uint64_t seed, dsc_rng** out);                       /* ChaCha20Rng::seed_from_u64 */
int dsc_rng_destroy(dsc_rng* rng);
int dsc_rng_next_u32(dsc_rng* rng, uint32_t* out);                      /* RngCore::next_u32: the per-step rand_seed */
int dsc_rng_next_u64(dsc_rng* rng, uint64_t* out);
int dsc_rng_open01_f32(dsc_rng* rng, float* out, size_t count);        /* rng.sample(Open01), environment.rs:22,35 */
int dsc_rng_gen_range(dsc_rng* rng, uint64_t low, uint64_t high, int as_u32, uint64_t* out);  /* gen_range(low..high): u32 or usize draws */
int dsc_rng_gen_range_pairs(dsc_rng* rng, uint64_t high0, uint64_t high1, uint64_t* out, size_t pairs); /* `pairs` x (gen_range(0..high0), gen_range(0..high1)) as usize draws: image_fit/main.rs:376-378 */
int dsc_rng_shuffle(dsc_rng* rng, uint64_t* indices, size_t count);    /* SliceRandom::shuffle, main.rs:382 */
int dsc_env_reset_parameter_rng(dsc_env* env, int param, dsc_rng* rng); /* Environment::reset_parameter(param, &mut rng), environment.rs:190-202 */

/* ---- Data front ends of the examples (descent_b200/csrc/host_io.cpp) --------------------------
 * Byte buffers returned through `bytes_out` are released with dsc_free_bytes. */
int dsc_load_gz_bytes(const char* path, uint8_t** bytes_out, size_t* size_out);                 /* fashion_mnist/main.rs:13-19 */
int dsc_gunzip(const uint8_t* data, size_t size, uint8_t** bytes_out, size_t* size_out);
int dsc_free_bytes(uint8_t* bytes);
int dsc_idx_images_info(const uint8_t* bytes, size_t size, uint32_t* images, uint32_t* rows, uint32_t* cols); /* read_images_info :26 */
int dsc_idx_labels_info(const uint8_t* bytes, size_t size, uint32_t* items);                                   /* read_labels_info :35 */
int dsc_idx_unpack_images(const uint8_t* bytes, size_t size, const uint64_t* indices, size_t count, float* out); /* unpack_images :42 */
int dsc_idx_unpack_labels(const uint8_t* bytes, size_t size, const uint64_t* indices, size_t count, float* out); /* unpack_labels :62 */
int dsc_jpeg_decode_rgb(const uint8_t* data, size_t size, int* width, int* height, uint8_t** rgb_out);          /* image_fit/main.rs:278-282 */
int dsc_write_ppm(const char* path, const float* rgb, int width, int height);                                  /* image_fit/main.rs:421-435 (as PPM) */

/* ---- Op-level entry points (the fused kernels without the graph builder; ops_api.cpp) ------------
 * A Rust kernel.rs that keeps the reference's graph passes can hand one cluster to these.  An op is planned once for its
 * shape and OWNS its operand / result buffers on the device: dsc_op_buffer returns their addresses (the caller's kernels
 * read and write them in place), dsc_op_parameter the parameter id (for dsc_env_write_parameter / dsc_env_read_parameter),
 * dsc_op_run launches the planned kernels on the environment's stream.  Precision follows dsc_env_set_tf32. */
typedef struct dsc_op dsc_op;
/* NHWC conv2d with replicate padding, stride and groups (Array::conv2d, array.rs:989-1031; kernel.rs:385-557,712-810).
 * backward = 0: buffers "x" [images, h, w, ic], "filter" [groups, oc / groups, fh, fw, ic / groups] -> "y".
 * backward = 1: buffers "x", "filter", "dy" -> "dx", "dfilter" (the backward-input and weight-gradient kernels). */
int dsc_op_conv2d(dsc_env* env, int64_t images, int64_t height, int64_t width, int64_t in_channels, int64_t out_channels, int64_t filter_h, int64_t filter_w,
                  int64_t pad, int64_t stride_w, int64_t stride_h, int64_t groups, int backward, dsc_op** out);
/* Deterministic sort-and-segmented-reduce scatter_add (kernel.rs:812-874 uses float atomics): "table" [rows, inner] +=
 * "values" [count, inner] at rows "indices" [count] (uint32 row numbers). */
int dsc_op_scatter_add(dsc_env* env, int64_t rows, int64_t inner, int64_t count, dsc_op** out);
/* loss.rs:4-34 as one row kernel: "z" [rows, classes], "y" [rows, 1] (labels as f32) -> "loss" [rows, 1], "accuracy" [rows, 1],
 * "dz" [rows, classes] = (softmax - onehot) / rows, the gradient of the batch-mean loss (DualArray::set_loss). */
int dsc_op_softmax_cross_entropy(dsc_env* env, int64_t rows, int64_t classes, dsc_op** out);
/* optimizer.rs:62-112 for `tensors` parameter tensors in ONE launch: buffers "theta<i>", "grad<i>" [counts[i]], and the
 * optimiser state "state<j>" (zeroed at creation: the step counter, then m and v per tensor). */
int dsc_op_adam_step(dsc_env* env, const int64_t* counts, int tensors, float learning_rate, float beta1, float beta2, float epsilon, dsc_op** out);
int dsc_op_buffer(dsc_op* op, const char* name, void** device_ptr, size_t* bytes);
int dsc_op_parameter(dsc_op* op, const char* name, int* param);
int dsc_op_run(dsc_op* op, uint32_t rand_seed);
int dsc_op_destroy(dsc_op* op);

#ifdef __cplusplus
}
#endif
#endif /* DESCENT_API_H */

// Graph compiler: optimisation passes over the op graph and partitioning into kernels ("clusters").
//
// The pass list follows the reference's `Graph::new` (src/graph.rs:111-139): dead-code elimination,
// move elimination, x*1 / x+0 simplification, common-subgraph elimination, clustering.  What differs
// is B200-first:
//   * every Mov folds into its consumers' ViewChain (no materialised im2col / window / permute
//     copies; the reference keeps a copy whenever a reshape cannot fold, graph.rs:262-284),
//   * a MatMul followed by its split-K Reduce (array.rs:515) is one GEMM cluster, the K split is
//     chosen by the backend rather than by MATMUL_MAX_K_SIZE,
//   * per-element clusters are connected components of fusable edges inside one dependency level,
//     which is cycle-free by construction (the reference re-checks reachability per candidate,
//     graph.rs:366-435),
//   * all gradient AllReduce nodes sit on one level so they form a single bucket.
// Results never depend on clustering.
#pragma once
#include <string>
#include <vector>

#include "array.hpp"

namespace descent {

enum class ClusterKind { PerElement, Reduce, MatMul, Unpad, WindowsToImage, ScatterAdd, AllReduce, Row };

// one buffer argument of a kernel: a producer node read through a chain
struct ClusterInput {
    int node_id = -1;
    ViewChain chain;
    Shape arg_shape;
    bool operator==(const ClusterInput& o) const { return node_id == o.node_id && chain == o.chain && arg_shape == o.arg_shape; }
};

// straight-line program of a per-element kernel (reference: PerElementKernelOp, kernel.rs:8-35)
struct PerElementOp {
    enum Kind { Load, Literal, BuiltIn, Unary, Binary, Select, Gather, Reduce } kind = Load;  // Reduce: Row clusters only (over the row, args[0])
    bool wide = true;            // Row clusters: one value per (row, k) -- false: one value per row
    int input_index = -1;        // Load / Gather: index into Cluster::inputs
    Op op;                       // Literal / BuiltIn / Unary / Binary / Select / Gather payload
    ViewChain chain;             // BuiltIn: chain from the built-in's own index space
    Shape arg_shape;             // BuiltIn / Gather
    int args[MAX_OP_ARGS] = {-1, -1, -1, -1};
    Shape shape;                 // Gather: shape of the gather node
};

struct Cluster {
    ClusterKind kind = ClusterKind::PerElement;
    int level = 0;
    std::vector<int> members;          // op nodes computed by this kernel, in topological order
    std::vector<ClusterInput> inputs;  // buffers read
    std::vector<int> outputs;          // op nodes whose values are written to memory
    // PerElement
    int64_t element_count = 0;
    std::vector<PerElementOp> ops;
    std::vector<int> output_ops;       // op index stored to outputs[i]
    // single-node kernels
    int node_id = -1;                  // the Reduce / MatMul / Unpad / WindowsToImage / ScatterAdd / AllReduce node
    bool matmul_absorbs_reduce = false;  // outputs[0] is the Reduce(axis 0) node that followed the MatMul
    // MatMul whose only consumer was a stride-1 WindowsToImage (conv2d's backward-input pass): one implicit GEMM
    // over output pixels, outputs[0] is the WindowsToImage node and the window matrix is never materialised.
    // Row r = (image, y, x), k = (fy, fx, matmul k), column = channel within the group; an A element is zero
    // unless (y - fy, x - fx) is a window position, i.e. lies in [0, out_h) x [0, out_w).
    struct ConvBackwardInput {
        bool enabled = false;
        int64_t in_h = 0, in_w = 0, out_h = 0, out_w = 0, filter_h = 0, filter_w = 0, matmul_k = 0;
        std::vector<ClusterInput> unfused;  // the MatMul's own operands [group, pixel, k] and [group, k, (fy, fx, channel)]
        // Unpad nodes (the adjoint of conv2d's replicate padding) that followed the WindowsToImage, also absorbed:
        // amounts along the image's h and w axes and which of the two the graph applies first (1 = h, 2 = w, 0 = none)
        int64_t unpad_h = 0, unpad_w = 0;
        int unpad_first_axis = 0;
        // Strided windows (MaxBlurPool2D's depthwise stride-2 blur, module.rs:139-163): a window position is a division by
        // the stride, which no view expresses, so the MatMul's operands keep their own chains (`unfused` = `inputs`) and the
        // cluster runs as ONE gather kernel: one thread per element of the final image gradient sums the few (window, tap)
        // pairs that contain it -- dY x W^T, col2im and both Unpads without the 9x window matrix or the padded gradient.
        bool strided = false;
        int64_t stride_h = 1, stride_w = 1;
    } conv_backward_input;
    int copy_from = -1;                // ScatterAdd: accumulator node taken in place (graph.rs:601-621)
    // ScatterAdd: the values of source s computed while they are loaded.  A hash grid's backward pass scatters
    // (gradient of the interpolated feature) x (corner weight) into the table (examples/image_fit/main.rs:190-199): forty
    // [m, 2] products per step that exist only to be read once by a scatter.  When a source's values are a per-element op
    // on loaded operands with no other reader, the op joins the scatter's cluster: `value_programs[s]` (members empty = the
    // source is an ordinary array) is its program, its operands are `inputs[value_input_base[s] + i]`, and the product
    // array is never written or read back (Graph::absorb_scatter_values).
    std::vector<Cluster> value_programs;
    std::vector<int> value_input_base;
    // MatMul whose product is consumed, element for element, by exactly one per-element cluster (conv2d's bias +
    // activation): that cluster is evaluated on the accumulator in the GEMM's epilogue and the raw product never
    // exists in memory.  `epilogue[0]` is the absorbed cluster; its input `epilogue_product_input` is the product;
    // its other inputs are appended to `inputs` (from index 2 on, in order) and its outputs replace `outputs`.
    std::vector<Cluster> epilogue;
    int epilogue_product_input = -1;
    // ... and whose single result, an NHWC activation [images, height, width, channels], is also max-pooled over
    // non-overlapping window_h x window_w windows (DualArray::max_pool2d, array.rs:1033-1047: windows view + reduce_max):
    // the GEMM kernels that hold a whole tile of the activation before storing it also store the pooled maxima, so the
    // pooling pass never re-reads the activation.  `pool.reduce[0]` is the absorbed Reduce cluster (other GEMM kernels
    // run it after the product); the pooled node is the LAST entry of `outputs`.
    struct Pool {
        bool enabled = false;
        int64_t images = 0, height = 0, width = 0, channels = 0, window_h = 0, window_w = 0;
        std::vector<Cluster> reduce;
    } pool;
    // MatMul whose B operand is a [K, C] array that the graph also sums over K with a chain of Reduce(sum) nodes (the
    // bias gradient next to a convolution's / dense layer's weight gradient: both read dY once per step).  The GEMM
    // kernels that stream B produce those column sums on the side; `column_sum[i]` are the absorbed Reduce clusters in
    // order (other GEMM kernels simply run them after the product) and outputs[1] is the last Reduce's node.
    std::vector<Cluster> column_sum;
    // PerElement: several small per-element programs of one dependency level (different element counts: the Adam
    // updates of all parameter tensors, the per-level bookkeeping of a hash grid) launched as ONE kernel; block ranges
    // select the program.  inputs / outputs are the concatenation of the sub-clusters' in order.
    std::vector<Cluster> group;
    // Row: per-element ops on [rows, row_length] / [rows, 1] arrays and the reductions along the row that connect them
    // (softmax cross-entropy with its accuracy and gradient, loss.rs:4-34: max, exp, sum, divide, log, pick, subtract)
    // as ONE kernel, one thread per row, the row in registers; `ops` is the program (PerElementOp::wide / Reduce).
    int64_t rows = 0, row_length = 0;
    std::string label;                 // as the reference's Kernel::label_name (kernel.rs)
};

// A per-element cluster whose single result is needed only as an operand of MatMul clusters (conv2d's backward pass:
// dY = max-pool backward o activation backward feeds the weight-gradient GEMM, its bias column sums and the
// backward-input GEMM and nothing else).  The code generator may evaluate that cluster's program inside the GEMMs'
// operand loaders instead of running it as a kernel: the array is then never written or read back.  The decision is
// the code generator's (it depends on which GEMM kernels the options select, environment.cpp generate_graph_source);
// the graph only records who could.
struct OperandPrologue {
    int producer = -1;  // index of the per-element cluster
    struct Use { int cluster; int operand; };  // MatMul cluster index, 0 = A / 1 = B
    std::vector<Use> uses;
};

// A multi-layer perceptron's whole training step over the batch rows (the image_fit heads,
// examples/image_fit/main.rs:50-118,259-275; module.rs:70-90 Dense): forward layers x_{l+1} = act(x_l W_l + b_l), a
// per-element loss on the last product, and for every layer the weight gradient x_l^T dz_l (+ bias column sums) and the
// backward product dz_l W_l^T (+ activation backward).  Every one of these is row-local except the sums over the batch,
// so a 128-row tile can go through all of them without its activations ever leaving the SM.  The graph only records the
// candidate (cluster indices in execution order); the code generator decides (tensor-core path, widths that fit shared
// memory): it then emits ONE kernel at the position of the last cluster and the other clusters run nothing.
struct DenseChain {
    int64_t rows = 0;                  // batch rows M
    std::vector<int64_t> widths;       // widths[0] = width of the input x_0, widths[l + 1] = output width of layer l
    std::vector<int> forward;          // per layer: MatMul cluster x_l W_l (+ per-element epilogue: bias, activation)
    std::vector<int> weight_gradient;  // per layer: MatMul cluster x_l^T dz_l (+ column sums of dz_l)
    std::vector<int> backward;         // per layer: MatMul cluster dz_l W_l^T (+ epilogue: activation backward of layer l-1); backward[0] = -1 if x_0 needs no gradient
    int loss = -1;                     // per-element cluster: last product (+ bias, target) -> dz_{L-1} and values that are summed
    bool loss_in_epilogue = false;     // ... absorbed as the (multi-output) epilogue of the last forward cluster: loss == forward.back()
    int loss_gradient_output = 0;      // which output of the loss cluster is dz_{L-1}
    struct Sum { int output; int row_reduce; int batch_reduce; };  // loss output -> Reduce along the row -> Reduce over the batch
    std::vector<Sum> sums;
    std::vector<int> all_clusters() const;
    int last_cluster() const;
};

class Graph {
public:
    Graph(SharedParameters parameters, const OpGraph& ops, DataParallel dp);

    const OpGraph& ops() const { return ops_; }
    const std::vector<Cluster>& clusters() const { return clusters_; }  // already in execution order
    const std::vector<OperandPrologue>& operand_prologues() const { return operand_prologues_; }
    const std::vector<DenseChain>& dense_chains() const { return dense_chains_; }
    const SharedParameters& parameters() const { return parameters_; }
    const DataParallel& dp() const { return dp_; }
    std::vector<int> input_nodes() const;
    std::vector<int> output_nodes() const;

    enum class KernelDotOutput { None, Cluster, Color };
    void write_dot_file(KernelDotOutput mode, const std::string& path) const;  // graph.rs:656
    std::string export_json() const;  // optimised graph + clusters (tests, debugging)

    // opaque per-graph state owned by the executor (compiled module, memory plan, CUDA graph)
    mutable std::shared_ptr<void> executor_state;

private:
    void eliminate_dead_code();
    void eliminate_moves();
    void simplify_arithmetic();
    void eliminate_common_subgraphs();
    void reuse_activation_sign();
    void hoist_all_reduce_views();
    void sink_permutations_into_per_element();
    void sink_views_into_selects();
    void absorb_per_element_epilogues(std::vector<Cluster>& clusters);
    void absorb_column_sums(std::vector<Cluster>& clusters);
    void absorb_max_pools(std::vector<Cluster>& clusters);
    void absorb_scatter_values(std::vector<Cluster>& clusters);
    void fuse_rows(std::vector<Cluster>& clusters);
    void sink_parameter_updates(std::vector<Cluster>& clusters);
    void group_small_per_element(std::vector<Cluster>& clusters);
    bool absorb_unpad(std::vector<Cluster>& clusters, const std::vector<std::vector<std::pair<int, int>>>& cons, int id);
    bool absorb_windows_to_image(std::vector<Cluster>& clusters, const std::vector<std::vector<std::pair<int, int>>>& cons, int id);
    void build_clusters();
    void build_per_element_program(Cluster& c);
    void find_operand_prologues();
    std::vector<DenseChain> detect_dense_chains(const std::vector<Cluster>& clusters) const;
    void schedule_after_dense_chains(std::vector<Cluster>& clusters);

    SharedParameters parameters_;
    OpGraph ops_;
    DataParallel dp_;
    std::vector<Cluster> clusters_;
    std::vector<OperandPrologue> operand_prologues_;
    std::vector<DenseChain> dense_chains_;
};

std::string export_ops_json(const OpGraph& ops, const std::vector<ParameterStorage>& parameters, const std::vector<Cluster>* clusters);

}  // namespace descent

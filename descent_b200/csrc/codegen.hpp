// CUDA C emission for clusters (replaces the GLSL emission of the reference's src/kernel.rs:151-874).
// One translation unit per graph is JIT-compiled with NVRTC for sm_100a (runtime.cu).
#pragma once
#include <string>
#include <vector>

#include "graph.hpp"

namespace descent {

struct KernelArg {
    enum Kind { NodeBuffer, Scratch } kind = NodeBuffer;
    int node_id = -1;            // NodeBuffer: the op node whose storage is bound
    int64_t scratch_offset = 0;  // Scratch: byte offset inside the cluster's scratch area
};

struct KernelLaunch {
    enum Kind { Kernel, ZeroScratch, AllReduce, TensorGemm } kind = Kernel;
    std::string entry;
    uint32_t grid_x = 1, grid_y = 1, grid_z = 1, block = 256, smem = 0;
    std::vector<KernelArg> args;
    int64_t zero_offset = 0, zero_bytes = 0;  // ZeroScratch
    int64_t gemm_m = 0, gemm_n = 0, gemm_k = 0;  // TensorGemm: args = {A, B, C}
    bool gemm_a_is_mk = true, gemm_b_is_kn = true;
    int gemm_splits = 1;  // TensorGemm: k slices, each writing a partial product [M, N] into args[2] (then a scratch workspace)
    std::string label;
    int cluster = -1;
    std::vector<int> covers;       // clusters whose work this launch does besides its own (a fused dense chain)
    double algorithmic_bytes = 0;  // 4*(sum of min(source, addressed) input elements + outputs)
    double flops = 0;              // 2*b*m*n*k for GEMMs
    double replaced_bytes = 0;     // a fused launch: the algorithmic bytes of the kernels it replaces (0 = its own)
};

struct ClusterCode {
    std::string source;
    std::vector<KernelLaunch> launches;
    int64_t scratch_bytes = 0;
    bool column_sum_done = false;  // the GEMM kernel also produced Cluster::column_sum's result (outputs[1])
    bool pool_done = false;        // the GEMM kernel also produced Cluster::pool's result (outputs.back())
    // operand prologues (graph.hpp OperandPrologue): nodes this cluster's kernels read although no graph edge says so
    // (the inputs of a producer evaluated inside the operand loader): the planner keeps them alive until this cluster
    std::vector<int> extra_reads;
    // results of OTHER clusters this cluster's kernels write (a fused dense chain runs at its last cluster's slot): the
    // planner gives them storage here
    std::vector<int> extra_writes;
    bool skipped = false;  // a producer every consumer evaluates on the fly: no kernel, its output is never materialised
};

// Request to evaluate per-element producers inside a MatMul cluster's operand loaders.  `fused[k]` is set by the kernel
// generator that honours it; a request that comes back unfused means the chosen kernel cannot (the caller then
// generates the cluster again without the request and the producer runs as its own kernel).
struct PrologueRequest {
    const Cluster* producer[2] = {nullptr, nullptr};  // operand A / B
    bool fused[2] = {false, false};
};

struct CodegenOptions {
    int sm_count = 148;
    int dp_rank = 0;
    bool use_tf32 = false;  // plain dense MatMuls go to the tcgen05 TF32 kernel (dsc_gemm_tf32)
};

std::string kernel_prelude();
ClusterCode generate_cluster_code(const Graph& graph, int cluster_index, const CodegenOptions& options, PrologueRequest* prologue = nullptr);

// One kernel for a whole dense chain (graph.hpp DenseChain), to run at the chain's last cluster; false if these options or
// widths rule it out (the clusters then run one by one as usual).
bool generate_dense_chain_code(const Graph& graph, const DenseChain& chain, const CodegenOptions& options, ClusterCode* out);

// One partial launch + one sum launch for several independent shared-memory-table scatter_adds of one level (cluster
// indices in execution order; the code runs at the last one's slot); false if a member does not qualify.
bool generate_scatter_group_code(const Graph& graph, const std::vector<int>& members, const CodegenOptions& options, ClusterCode* out);

// host-side evaluation of a chain (tests, layout heuristics): consumer element -> producer element
int64_t eval_chain(const ViewChain& chain, int64_t e);

}  // namespace descent

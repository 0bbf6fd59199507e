// canonicalvoting_b200/csrc/graph.cu -- CUDA-graph capture / instantiation / launch behind the C ABI.
//
// The inference engine records one scene (coordinate-map builder, convolution program, decode, vote) into a CUDA graph and
// replays it (canonicalvoting_b200/engine.py SceneGraph).  Capturing here instead of with torch.cuda.CUDAGraph gives control
// over the instantiation flags: the map-builder kernels are captured on a HIGH-PRIORITY stream branch and the graph is
// instantiated with cudaGraphInstantiateFlagUseNodePriority, so that -- with several scenes in flight -- the small map
// kernels of scene i+1 are dispatched next to the persistent convolution CTAs of scene i (they fit beside them: no shared
// memory, < 12 K registers) instead of queueing behind the pending CTAs of its programmatically launched convolutions.
#include "common.cuh"

using namespace cvb200;

extern "C" int cvb200_graph_begin(void *stream_) {
    CVB_CUDA(cudaStreamBeginCapture((cudaStream_t)stream_, cudaStreamCaptureModeThreadLocal));
    return 0;
}

extern "C" int cvb200_graph_end(void *stream_, int32_t use_node_priority, void **exec_out, int64_t *n_nodes) {
    CVB_REQUIRE(exec_out, CVB200_EINVAL, "graph_end: NULL argument");
    cudaGraph_t graph = nullptr;
    CVB_CUDA(cudaStreamEndCapture((cudaStream_t)stream_, &graph));
    CVB_REQUIRE(graph != nullptr, CVB200_EINVAL, "graph_end: the capture was invalidated");
    size_t nodes = 0;
    (void)cudaGraphGetNodes(graph, nullptr, &nodes);
    if (n_nodes) *n_nodes = (int64_t)nodes;
    cudaGraphExec_t exec = nullptr;
    const cudaError_t e = cudaGraphInstantiateWithFlags(&exec, graph, use_node_priority ? cudaGraphInstantiateFlagUseNodePriority : 0);
    (void)cudaGraphDestroy(graph);
    CVB_CUDA(e);
    *exec_out = (void *)exec;
    return 0;
}

extern "C" int cvb200_graph_abort(void *stream_) {
    cudaGraph_t graph = nullptr;
    (void)cudaStreamEndCapture((cudaStream_t)stream_, &graph);
    if (graph) (void)cudaGraphDestroy(graph);
    (void)cudaGetLastError();
    return 0;
}

extern "C" int cvb200_graph_launch(void *exec, void *stream_) {
    CVB_REQUIRE(exec, CVB200_EINVAL, "graph_launch: NULL graph");
    CVB_CUDA(cudaGraphLaunch((cudaGraphExec_t)exec, (cudaStream_t)stream_));
    return 0;
}

extern "C" int cvb200_graph_destroy(void *exec) {
    if (exec) CVB_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)exec));
    return 0;
}

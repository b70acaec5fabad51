// f2: the training loop itself on the device -- sweeps, validation statistics and the early-stopping rule run as ONE
// CUDA graph whose body is a WHILE conditional node, so a fit with a validation frame needs no host round trip per
// iteration (the reference decides on the host after every iteration: hpf_cavi.py:196-211, poisson_mf_cavi.py:200-217,
// gaussian_mf_cavi_bias.py:268-284; the tuning scripts always pass val_df, tune_all_models.py:76,124,177).
//
// Protocol (host side in _engine.DeviceLoop):
//   pmf_loop_begin(stream)      graph + conditional WHILE node; `stream` starts capturing INTO the node's body
//   ... the caller enqueues one iteration on `stream`: pass kernels, pmf_eval_stats ... (captured, not executed)
//   pmf_loop_decide(...)        captured last: counts the iteration, records the validation RMSE, applies the reference's
//                               stopping rule and sets the loop condition (cudaGraphSetConditional)
//   pmf_loop_end()              ends the capture and instantiates the graph
//   pmf_loop_run(stream)        launches it: the device iterates until the rule fires or max_iter is reached
// The iteration count and the RMSE history are read back once, after the loop.
#include "common.cuh"

namespace pmf {

// Same arithmetic as the host loops: rmse = sqrt(sum_sq / count) in float64, improvement = previous - current.
//   rule 0 (Poisson MF / HPF): stop when improvement < tol           (fires on negative improvement too)
//   rule 1 (Gaussian MF):      stop when 0 <= improvement < tol      (gaussian_mf_cavi_bias.py:279)
__global__ void loop_decide_kernel(cudaGraphConditionalHandle handle, const double* __restrict__ eval_out, int rule,
                                   double tol, int has_tol, int max_iter, int32_t* __restrict__ iter,
                                   double* __restrict__ history) {
    const int k = *iter + 1;   // iterations completed, this one included
    *iter = k;
    bool go = k < max_iter;
    if (eval_out != nullptr) {
        const double cnt = eval_out[0];
        const double rmse = cnt > 0.0 ? sqrt(eval_out[1] / cnt) : __longlong_as_double(0x7ff8000000000000ll);
        history[k - 1] = rmse;
        if (k >= 2 && has_tol) {
            const double imp = history[k - 2] - rmse;
            const bool stop = rule == 0 ? (imp < tol) : (imp >= 0.0 && imp < tol);
            if (stop) go = false;
        }
    }
    cudaGraphSetConditional(handle, go ? 1u : 0u);
}

}  // namespace pmf

using namespace pmf;

struct pmf_loop {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaGraphConditionalHandle handle = 0;
    cudaStream_t capture_stream = nullptr;
    bool capturing = false;
};

extern "C" {

int pmf_loop_begin(void* stream, pmf_loop** out) {
    PMF_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    PMF_REQUIRE(stream != nullptr, "the loop body is captured from a stream: the legacy default stream cannot capture");
    pmf_loop* L = new pmf_loop();
    auto fail = [&](int code) { pmf_loop_free(L); return code; };
    if (cudaGraphCreate(&L->graph, 0) != cudaSuccess) { set_error("cudaGraphCreate: %s", cudaGetErrorString(cudaGetLastError())); return fail(PMF_ECUDA); }
    cudaError_t e = cudaGraphConditionalHandleCreate(&L->handle, L->graph, 1, cudaGraphCondAssignDefault);
    if (e != cudaSuccess) { set_error("cudaGraphConditionalHandleCreate: %s", cudaGetErrorString(e)); cudaGetLastError(); return fail(PMF_EUNSUPPORTED); }
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = L->handle;
    p.conditional.type = cudaGraphCondTypeWhile;
    p.conditional.size = 1;
    cudaGraphNode_t node;
    e = cudaGraphAddNode(&node, L->graph, nullptr, 0, &p);
    if (e != cudaSuccess) { set_error("conditional graph node: %s", cudaGetErrorString(e)); cudaGetLastError(); return fail(PMF_EUNSUPPORTED); }
    L->capture_stream = (cudaStream_t)stream;
    e = cudaStreamBeginCaptureToGraph(L->capture_stream, p.conditional.phGraph_out[0], nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
    if (e != cudaSuccess) { set_error("cudaStreamBeginCaptureToGraph: %s", cudaGetErrorString(e)); cudaGetLastError(); return fail(PMF_ECUDA); }
    L->capturing = true;
    *out = L;
    return PMF_OK;
}

int pmf_loop_decide(pmf_loop* L, const double* d_eval_out, int32_t rule, double tol, int32_t has_tol, int32_t max_iter,
                    int32_t* d_iter, double* d_history, void* stream) {
    PMF_REQUIRE(L != nullptr && L->capturing, "pmf_loop_decide outside pmf_loop_begin / pmf_loop_end");
    PMF_REQUIRE((cudaStream_t)stream == L->capture_stream, "pmf_loop_decide must be enqueued on the capturing stream");
    PMF_REQUIRE(d_iter != nullptr && max_iter >= 1 && (rule == 0 || rule == 1), "bad argument");
    PMF_REQUIRE(d_eval_out == nullptr || d_history != nullptr, "history is NULL");
    loop_decide_kernel<<<1, 1, 0, L->capture_stream>>>(L->handle, d_eval_out, rule, tol, has_tol, max_iter, d_iter, d_history);
    PMF_LAUNCH_CHECK();
    return PMF_OK;
}

int pmf_loop_end(pmf_loop* L) {
    PMF_REQUIRE(L != nullptr && L->capturing, "no capture in progress");
    cudaGraph_t body = nullptr;
    L->capturing = false;
    PMF_CUDA(cudaStreamEndCapture(L->capture_stream, &body));
    PMF_CUDA(cudaGraphInstantiate(&L->exec, L->graph, 0));
    return PMF_OK;
}

int pmf_loop_run(pmf_loop* L, void* stream) {
    PMF_REQUIRE(L != nullptr && L->exec != nullptr, "loop is not instantiated");
    PMF_CUDA(cudaGraphLaunch(L->exec, (cudaStream_t)stream));
    return PMF_OK;
}

int pmf_loop_free(pmf_loop* L) {
    if (!L) return PMF_OK;
    if (L->capturing) {
        cudaGraph_t body = nullptr;
        cudaStreamEndCapture(L->capture_stream, &body);
        cudaGetLastError();
    }
    if (L->exec) cudaGraphExecDestroy(L->exec);
    if (L->graph) cudaGraphDestroy(L->graph);
    delete L;
    return PMF_OK;
}

}  // extern "C"

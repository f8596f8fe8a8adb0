// "The reference's cuDNN build on the same box": a call-for-call restatement of dg_nn::forward
// (src/libdg_nn/graph.rs:123-158 and layers/*.rs) against the cuDNN legacy API that the reference
// binds in src/libdg_cuda/cudnn/*.rs.  BASELINE / SECOND ORACLE ONLY -- never linked into the
// product library.  Built by baseline/Makefile against the image's cuDNN 9.10.2.
//
// Same descriptors as the reference: NHWC fp16 activations (common.rs:61-66), KRSC fp16 filters
// (conv2d.rs:111-117), CrossCorrelation, pad 1, TENSOR_OP math (convolution_descriptor.rs:165-179),
// compute type HALF for the tower and FLOAT for the heads / dense layers (conv2d.rs:54,
// policy_head.rs:51, value_head.rs:47, dense.rs:108-116), algorithm = top-1 of
// cudnnGetConvolutionForwardAlgorithm_v7 (convolution_fwd_algo_perf.rs:48-82), three streams and
// one event (graph.rs:145-147), blocking H2D / D2H (graph.rs:130-131,154-157).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cudnn.h>

#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

struct View { const char* name; const char* dtype; const void* data; unsigned long long nbytes; };   // == dg_tensor_view

namespace {

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { snprintf(g_err, sizeof g_err, "%s: %s", #x, cudaGetErrorString(e_)); return -1; } } while (0)
#define CN(x) do { cudnnStatus_t s_ = (x); if (s_ != CUDNN_STATUS_SUCCESS) { snprintf(g_err, sizeof g_err, "%s: %s", #x, cudnnGetErrorString(s_)); return -2; } } while (0)
char g_err[512];

struct Conv {
    cudnnTensorDescriptor_t x = nullptr, y = nullptr, b = nullptr;
    cudnnFilterDescriptor_t w = nullptr;
    cudnnConvolutionDescriptor_t c = nullptr;
    cudnnActivationDescriptor_t a = nullptr;
    cudnnConvolutionFwdAlgo_t algo;
    size_t ws = 0;
    float alpha[2] = {1.f, 0.f};
    __half *dw = nullptr, *db = nullptr;
    cudnnDataType_t compute;
};

struct Ref {
    int batch = 0, blocks = 0;
    cudnnHandle_t h = nullptr;
    cudaStream_t tower = nullptr, value = nullptr, policy = nullptr;
    cudaEvent_t done = nullptr, ev0 = nullptr, ev1 = nullptr, evv = nullptr, evp = nullptr;
    Conv up, pconv, vconv, pfc, vfc;
    std::vector<Conv> c1, c2;
    cudnnTensorDescriptor_t sm_desc = nullptr, v_desc = nullptr;
    cudnnActivationDescriptor_t tanh_desc = nullptr;
    __half *in = nullptr, *a = nullptr, *y = nullptr, *t = nullptr, *p1 = nullptr, *p2 = nullptr, *p3 = nullptr, *v1 = nullptr, *v2 = nullptr;
    void* workspace = nullptr;
    size_t ws_bytes = 0;
    std::string info;
};

int make_conv(Ref& r, Conv& c, int n, int cin, int cout, int wh, int ksize, cudnnDataType_t compute, bool relu, bool nchw_filter,
              float a1, float a2, const uint16_t* w_host, const uint16_t* b_host, const char* tag) {
    CN(cudnnCreateTensorDescriptor(&c.x));
    CN(cudnnCreateTensorDescriptor(&c.y));
    CN(cudnnCreateTensorDescriptor(&c.b));
    CN(cudnnCreateFilterDescriptor(&c.w));
    CN(cudnnCreateConvolutionDescriptor(&c.c));
    CN(cudnnCreateActivationDescriptor(&c.a));
    CN(cudnnSetTensor4dDescriptor(c.x, CUDNN_TENSOR_NHWC, CUDNN_DATA_HALF, n, cin, wh, wh));
    CN(cudnnSetTensor4dDescriptor(c.y, CUDNN_TENSOR_NHWC, CUDNN_DATA_HALF, n, cout, wh, wh));
    CN(cudnnSetTensor4dDescriptor(c.b, CUDNN_TENSOR_NHWC, CUDNN_DATA_HALF, 1, cout, 1, 1));
    CN(cudnnSetFilter4dDescriptor(c.w, CUDNN_DATA_HALF, nchw_filter ? CUDNN_TENSOR_NCHW : CUDNN_TENSOR_NHWC, cout, cin, ksize, ksize));
    CN(cudnnSetActivationDescriptor(c.a, relu ? CUDNN_ACTIVATION_RELU : CUDNN_ACTIVATION_IDENTITY, CUDNN_NOT_PROPAGATE_NAN, 0.0));
    c.alpha[0] = a1;
    c.alpha[1] = a2;
    c.compute = compute;
    for (int attempt = 0; attempt < 2; attempt++) {
        CN(cudnnSetConvolution2dDescriptor(c.c, ksize / 2, ksize / 2, 1, 1, 1, 1, CUDNN_CROSS_CORRELATION, c.compute));
        CN(cudnnSetConvolutionMathType(c.c, CUDNN_TENSOR_OP_MATH));
        int count = 0;
        cudnnConvolutionFwdAlgoPerf_t perf;
        cudnnStatus_t s = cudnnGetConvolutionForwardAlgorithm_v7(r.h, c.x, c.w, c.c, c.y, 1, &count, &perf);
        if (s == CUDNN_STATUS_SUCCESS && count == 1 && perf.status == CUDNN_STATUS_SUCCESS) {
            c.algo = perf.algo;
            c.ws = perf.memory;
            break;
        }
        if (attempt == 0 && c.compute == CUDNN_DATA_HALF) {      // fall back to FLOAT and record it
            c.compute = CUDNN_DATA_FLOAT;
            r.info += std::string(tag) + ": HALF compute rejected, using FLOAT; ";
            continue;
        }
        snprintf(g_err, sizeof g_err, "%s: no forward algorithm (%s)", tag, cudnnGetErrorString(s));
        return -2;
    }
    if (!relu && c.algo != CUDNN_CONVOLUTION_FWD_ALGO_IMPLICIT_PRECOMP_GEMM) {
        // cuDNN documents identity activation for IMPLICIT_PRECOMP_GEMM only
        c.algo = CUDNN_CONVOLUTION_FWD_ALGO_IMPLICIT_PRECOMP_GEMM;
        CN(cudnnGetConvolutionForwardWorkspaceSize(r.h, c.x, c.w, c.c, c.y, c.algo, &c.ws));
        r.info += std::string(tag) + ": identity activation -> PRECOMP_GEMM; ";
    }
    char buf[96];
    snprintf(buf, sizeof buf, "%s: algo %d compute %s ws %zu; ", tag, static_cast<int>(c.algo), c.compute == CUDNN_DATA_HALF ? "HALF" : "FLOAT", c.ws);
    r.info += buf;
    if (c.ws > r.ws_bytes) r.ws_bytes = c.ws;
    const size_t wn = static_cast<size_t>(cout) * cin * ksize * ksize;
    CK(cudaMalloc(&c.dw, wn * 2));
    CK(cudaMalloc(&c.db, (static_cast<size_t>(cout) + 16) * 2));
    CK(cudaMemcpy(c.dw, w_host, wn * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.db, b_host, static_cast<size_t>(cout) * 2, cudaMemcpyHostToDevice));
    return 0;
}

int run_conv(Ref& r, Conv& c, const __half* x, const __half* z, __half* y) {
    CN(cudnnConvolutionBiasActivationForward(r.h, &c.alpha[0], c.x, x, c.w, c.dw, c.c, c.algo, r.workspace, r.ws_bytes, &c.alpha[1], c.y,
                                             z ? z : y, c.b, c.db, c.a, c.y, y));
    return 0;
}

// the 25 cuDNN launches of one forward on the resident input
int enqueue(Ref& r) {
    CN(cudnnSetStream(r.h, r.tower));
    if (int rc = run_conv(r, r.up, r.in, nullptr, r.a)) return rc;
    __half *a = r.a, *t = r.t;
    for (int i = 0; i < r.blocks; i++) {
        if (int rc = run_conv(r, r.c1[i], a, nullptr, r.y)) return rc;       // residual_block.rs:76
        if (int rc = run_conv(r, r.c2[i], r.y, a, t)) return rc;             // residual_block.rs:77 (z = block input)
        std::swap(a, t);
    }
    CK(cudaEventRecord(r.done, r.tower));
    CK(cudaStreamWaitEvent(r.value, r.done, 0));
    CK(cudaStreamWaitEvent(r.policy, r.done, 0));
    const float one = 1.f, zero = 0.f;
    CN(cudnnSetStream(r.h, r.value));
    if (int rc = run_conv(r, r.vconv, a, nullptr, r.v1)) return rc;
    if (int rc = run_conv(r, r.vfc, r.v1, nullptr, r.v2)) return rc;
    CN(cudnnActivationForward(r.h, r.tanh_desc, &one, r.v_desc, r.v2, &zero, r.v_desc, r.v2));
    CN(cudnnSetStream(r.h, r.policy));
    if (int rc = run_conv(r, r.pconv, a, nullptr, r.p1)) return rc;
    if (int rc = run_conv(r, r.pfc, r.p1, nullptr, r.p2)) return rc;
    CN(cudnnSoftmaxForward(r.h, CUDNN_SOFTMAX_ACCURATE, CUDNN_SOFTMAX_MODE_INSTANCE, &one, r.sm_desc, r.p2, &zero, r.sm_desc, r.p3));
    return 0;
}

}  // namespace

extern "C" {

const char* dgref_last_error() { return g_err; }

int dgref_create(int device, int batch, const View* views, int count, float temperature, void** out) {
    std::map<std::string, const View*> t;
    for (int i = 0; i < count; i++) t[views[i].name] = &views[i];
    auto f16 = [&](const std::string& name) -> const uint16_t* {
        auto it = t.find(name);
        return it == t.end() ? nullptr : static_cast<const uint16_t*>(it->second->data);
    };
    CK(cudaSetDevice(device));
    Ref* r = new Ref();
    *out = r;
    r->batch = batch;
    CN(cudnnCreate(&r->h));
    CK(cudaStreamCreate(&r->tower));
    CK(cudaStreamCreate(&r->value));
    CK(cudaStreamCreate(&r->policy));
    CK(cudaEventCreateWithFlags(&r->done, cudaEventDisableTiming));
    CK(cudaEventCreate(&r->ev0));
    CK(cudaEventCreate(&r->ev1));
    CK(cudaEventCreateWithFlags(&r->evv, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&r->evp, cudaEventDisableTiming));
    const float tau = 1.0f / temperature;
    char name[96];
    if (!f16("01_upsample/conv_1:0")) { snprintf(g_err, sizeof g_err, "missing 01_upsample"); return -4; }
    if (int rc = make_conv(*r, r->up, batch, 32, 128, 19, 3, CUDNN_DATA_HALF, true, false, 1.f, 0.f, f16("01_upsample/conv_1:0"),
                           f16("01_upsample/conv_1/offset:0"), "up")) return rc;
    for (int i = 0;; i++) {
        snprintf(name, sizeof name, "%02d_residual/conv_1:0", i + 2);
        const uint16_t* w1 = f16(name);
        snprintf(name, sizeof name, "%02d_residual/conv_2:0", i + 2);
        const uint16_t* w2 = f16(name);
        if (!w1 || !w2) break;
        snprintf(name, sizeof name, "%02d_residual/conv_1/offset:0", i + 2);
        const uint16_t* b1 = f16(name);
        snprintf(name, sizeof name, "%02d_residual/conv_2/offset:0", i + 2);
        const uint16_t* b2 = f16(name);
        snprintf(name, sizeof name, "%02d_residual/alpha:0", i + 2);
        float g = 0.5f;
        if (t.count(name)) memcpy(&g, t[name]->data, 4);
        r->c1.emplace_back();
        r->c2.emplace_back();
        if (int rc = make_conv(*r, r->c1.back(), batch, 128, 128, 19, 3, CUDNN_DATA_HALF, true, false, 1.f, 0.f, w1, b1, i ? "" : "res conv_1")) return rc;
        if (int rc = make_conv(*r, r->c2.back(), batch, 128, 128, 19, 3, CUDNN_DATA_HALF, true, false, g, 1.0f - g, w2, b2, i ? "" : "res conv_2")) return rc;
        CN(cudnnScaleTensor(r->h, r->c2.back().b, r->c2.back().db, &g));       // residual_block.rs:72-74
        r->blocks++;
    }
    const int hidx = 2 + r->blocks;
    auto head = [&](const char* fmt) { snprintf(name, sizeof name, fmt, hidx); return f16(name); };
    const uint16_t *pw = head("%02dp_policy/conv_1:0"), *pb = head("%02dp_policy/conv_1/offset:0");
    const uint16_t *pl = head("%02dp_policy/linear_1:0"), *plb = head("%02dp_policy/linear_1/offset:0");
    const uint16_t *vw = head("%02dv_value/conv_1:0"), *vb = head("%02dv_value/conv_1/offset:0");
    const uint16_t *vl = head("%02dv_value/linear_2:0"), *vlb = head("%02dv_value/linear_2/offset:0");
    if (!pw || !pb || !pl || !plb || !vw || !vb || !vl || !vlb) { snprintf(g_err, sizeof g_err, "missing head tensors"); return -4; }
    if (int rc = make_conv(*r, r->pconv, batch, 128, 8, 19, 3, CUDNN_DATA_FLOAT, true, false, 1.f, 0.f, pw, pb, "policy conv")) return rc;
    if (int rc = make_conv(*r, r->vconv, batch, 128, 2, 19, 3, CUDNN_DATA_FLOAT, true, false, 1.f, 0.f, vw, vb, "value conv")) return rc;
    // dense.rs:79-96: the file stores [in][out]; cudnnTransformTensor makes it [out][in] on first use
    std::vector<uint16_t> plt(362 * 2888), vlt(722);
    for (int i = 0; i < 2888; i++)
        for (int o = 0; o < 362; o++) plt[static_cast<size_t>(o) * 2888 + i] = pl[static_cast<size_t>(i) * 362 + o];
    for (int i = 0; i < 722; i++) vlt[i] = vl[i];
    if (int rc = make_conv(*r, r->pfc, batch, 2888, 362, 1, 1, CUDNN_DATA_FLOAT, false, true, tau, 0.f, plt.data(), plb, "policy dense")) return rc;
    if (int rc = make_conv(*r, r->vfc, batch, 722, 1, 1, 1, CUDNN_DATA_FLOAT, false, true, 1.f, 0.f, vlt.data(), vlb, "value dense")) return rc;
    CN(cudnnScaleTensor(r->h, r->pfc.b, r->pfc.db, &tau));                      // policy_head.rs:87-89
    CN(cudnnCreateTensorDescriptor(&r->sm_desc));
    CN(cudnnSetTensor4dDescriptor(r->sm_desc, CUDNN_TENSOR_NHWC, CUDNN_DATA_HALF, batch, 362, 1, 1));
    CN(cudnnCreateTensorDescriptor(&r->v_desc));
    CN(cudnnSetTensor4dDescriptor(r->v_desc, CUDNN_TENSOR_NHWC, CUDNN_DATA_HALF, batch, 1, 1, 1));
    CN(cudnnCreateActivationDescriptor(&r->tanh_desc));
    CN(cudnnSetActivationDescriptor(r->tanh_desc, CUDNN_ACTIVATION_TANH, CUDNN_NOT_PROPAGATE_NAN, 0.0));
    const size_t act = static_cast<size_t>(batch) * 361 * 128 * 2;
    CK(cudaMalloc(&r->in, static_cast<size_t>(batch) * 361 * 32 * 2));
    CK(cudaMalloc(&r->a, act));
    CK(cudaMalloc(&r->y, act));
    CK(cudaMalloc(&r->t, act));
    CK(cudaMalloc(&r->p1, static_cast<size_t>(batch) * 2888 * 2));
    CK(cudaMalloc(&r->p2, static_cast<size_t>(batch) * 362 * 2));
    CK(cudaMalloc(&r->p3, static_cast<size_t>(batch) * 362 * 2));
    CK(cudaMalloc(&r->v1, static_cast<size_t>(batch) * 722 * 2));
    CK(cudaMalloc(&r->v2, (static_cast<size_t>(batch) + 16) * 2));
    CK(cudaMalloc(&r->workspace, r->ws_bytes + 256));
    CK(cudaDeviceSynchronize());
    return 0;
}

const char* dgref_info(void* h) { return static_cast<Ref*>(h)->info.c_str(); }
int dgref_num_blocks(void* h) { return static_cast<Ref*>(h)->blocks; }

// dg_nn::forward: blocking, host buffers (graph.rs:123-158)
int dgref_forward(void* h, const uint16_t* features, uint16_t* value, uint16_t* policy) {
    Ref& r = *static_cast<Ref*>(h);
    CK(cudaMemcpyAsync(r.in, features, static_cast<size_t>(r.batch) * 361 * 32 * 2, cudaMemcpyHostToDevice, r.tower));
    if (int rc = enqueue(r)) return rc;
    CK(cudaMemcpyAsync(value, r.v2, static_cast<size_t>(r.batch) * 2, cudaMemcpyDeviceToHost, r.value));
    CK(cudaStreamSynchronize(r.value));
    CK(cudaMemcpyAsync(policy, r.p3, static_cast<size_t>(r.batch) * 362 * 2, cudaMemcpyDeviceToHost, r.policy));
    CK(cudaStreamSynchronize(r.policy));
    return 0;
}

// device time of `iters` forwards on the resident input (events on the tower stream; the heads' streams are joined)
int dgref_time_resident(void* h, int iters, float* ms) {
    Ref& r = *static_cast<Ref*>(h);
    CK(cudaEventRecord(r.ev0, r.tower));
    for (int i = 0; i < iters; i++) {
        if (int rc = enqueue(r)) return rc;
        CK(cudaEventRecord(r.evv, r.value));
        CK(cudaEventRecord(r.evp, r.policy));
        CK(cudaStreamWaitEvent(r.tower, r.evv, 0));
        CK(cudaStreamWaitEvent(r.tower, r.evp, 0));
    }
    CK(cudaEventRecord(r.ev1, r.tower));
    CK(cudaStreamSynchronize(r.tower));
    CK(cudaEventElapsedTime(ms, r.ev0, r.ev1));
    return 0;
}

// final tower activation of the last forward, [batch][361][128] fp16
int dgref_read_tower(void* h, uint16_t* out) {
    Ref& r = *static_cast<Ref*>(h);
    CK(cudaDeviceSynchronize());
    const __half* a = (r.blocks % 2 == 0) ? r.a : r.t;
    CK(cudaMemcpy(out, a, static_cast<size_t>(r.batch) * 361 * 128 * 2, cudaMemcpyDeviceToHost));
    return 0;
}

// ---- the reference's predictor under the reference's batching rules ------------------------------------------------
// `NnPredictor::predict` (src/libdg_mcts/predictors/nn.rs:84-107) as the pool drives it: a batch holds at most
// `--batch-size` leaves (default 16, config.rs:137; batch.rs:113 splits off the last max_batch_size events), at most
// `max_num_threads() = 2 x devices` batches are alive at a time (nn.rs:64-67, batch.rs:98-105), the thread that cut a batch
// blocks in `batch.forward` (worker_thread.rs:88-99), the features are plain fp16 NHWC tensors in pageable host memory
// (23,104 bytes per leaf, batch.rs:87-91) and `nn::forward` copies them in and the outputs out with blocking copies.
// Signature == dg_predict_fn (include/dg_mcts.h), so the product's self-play loop can run on it: the loop is then the
// product's restatement of self_play.rs / tree.rs, the evaluation path is the reference's.
struct Packed { uint32_t planes[361]; uint16_t k_bits, reserved; };   // == dg_packed_position

struct RefPredictor {
    int batch_size = 16, lanes = 2, device = 0;
    std::vector<Ref*> refs;
    std::vector<std::vector<uint16_t>> feats, value, policy;
    std::vector<char> busy;
    std::mutex m;
    std::condition_variable cv;
    long long calls = 0, leaves = 0;
};

int dgref_predictor_create(int device, int batch_size, int lanes, const View* views, int count, float temperature, void** out) {
    if (batch_size < 8) batch_size = 8;             // the root evaluation is one call of 8 positions (lib.rs:97-111)
    RefPredictor* p = new RefPredictor();
    *out = p;
    p->batch_size = batch_size;
    p->lanes = lanes;
    p->device = device;
    for (int i = 0; i < lanes; i++) {
        void* r = nullptr;
        if (int rc = dgref_create(device, batch_size, views, count, temperature, &r)) return rc;
        p->refs.push_back(static_cast<Ref*>(r));
        p->feats.emplace_back(static_cast<size_t>(batch_size) * 361 * 32, 0);
        p->value.emplace_back(batch_size);
        p->policy.emplace_back(static_cast<size_t>(batch_size) * 362);
        p->busy.push_back(0);
        // first use loads cuDNN's kernels (seconds): not part of any timed sample
        if (int rc = dgref_forward(p->refs.back(), p->feats.back().data(), p->value.back().data(), p->policy.back().data())) return rc;
    }
    return 0;
}

int dgref_predict(void* ctx, const Packed* positions, int n, uint16_t* value, uint16_t* policy) {
    RefPredictor& p = *static_cast<RefPredictor*>(ctx);
    for (int at = 0; at < n; at += p.batch_size) {
        const int m = n - at < p.batch_size ? n - at : p.batch_size;
        int lane = -1;
        {
            std::unique_lock<std::mutex> lk(p.m);
            p.cv.wait(lk, [&] { for (int i = 0; i < p.lanes; i++) if (!p.busy[i]) return true; return false; });
            for (int i = 0; i < p.lanes; i++) if (!p.busy[i]) { lane = i; break; }
            p.busy[lane] = 1;
            p.calls++;
            p.leaves += m;
        }
        uint16_t* f = p.feats[lane].data();
        for (int i = 0; i < p.batch_size; i++) {        // a short batch is padded with its last position (cuDNN time is flat in the batch size here)
            const Packed& q = positions[at + (i < m ? i : m - 1)];
            uint16_t* o = f + static_cast<size_t>(i) * 361 * 32;
            for (int pt = 0; pt < 361; pt++) {
                const uint32_t mask = q.planes[pt];
                o[pt * 32 + 0] = (mask & 1u) ? q.k_bits : 0;
                o[pt * 32 + 1] = (mask & 2u) ? q.k_bits : 0;
                for (int c = 2; c < 32; c++) o[pt * 32 + c] = ((mask >> c) & 1u) ? 0x3c00 : 0;
            }
        }
        int rc = cudaSetDevice(p.device) == cudaSuccess ? 0 : -1;
        if (rc == 0) rc = dgref_forward(p.refs[lane], f, p.value[lane].data(), p.policy[lane].data());
        if (rc == 0) {
            memcpy(value + at, p.value[lane].data(), static_cast<size_t>(m) * 2);
            memcpy(policy + static_cast<size_t>(at) * 362, p.policy[lane].data(), static_cast<size_t>(m) * 362 * 2);
        }
        { std::lock_guard<std::mutex> lk(p.m); p.busy[lane] = 0; }
        p.cv.notify_one();
        if (rc) return rc;
    }
    return 0;
}

void dgref_predictor_stats(void* ctx, long long* calls, long long* leaves) {
    RefPredictor& p = *static_cast<RefPredictor*>(ctx);
    std::lock_guard<std::mutex> lk(p.m);
    *calls = p.calls;
    *leaves = p.leaves;
}

void dgref_destroy(void* h);
void dgref_predictor_destroy(void* ctx) {
    if (!ctx) return;
    RefPredictor* p = static_cast<RefPredictor*>(ctx);
    for (Ref* r : p->refs) dgref_destroy(r);
    delete p;
}

void dgref_destroy(void* h) {
    if (!h) return;
    Ref* r = static_cast<Ref*>(h);
    cudaDeviceSynchronize();
    if (r->h) cudnnDestroy(r->h);
    delete r;      // device buffers are released with the process; this is a short-lived measurement helper
}

}  // extern "C"

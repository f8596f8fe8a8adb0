// C-ABI engine: weight loading / re-layout, workspaces, forward orchestration, leaf queue.
// Host-side counterpart of src/libdg_nn/{network,graph,loader,tensor}.rs behind include/dg_engine.h.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <sys/prctl.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cstddef>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dg_engine.h"
#include "conv_tc.h"
#include "go_board.h"
#include "kernels.h"
#include "layout.h"
#include "weights_file.h"

namespace {

using dg::ConvTcParams;
using dg::ConvTcShape;

constexpr int kFeatBytes = DG_FEATURE_SIZE * 2;      // 23,104 bytes of fp16 features per position
constexpr int kChan = 128;
constexpr int kHeadChan = 16;

struct Workspace {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint8_t* d_in = nullptr;          // raw NHWC features or compact positions of the current batch
    __half *feat = nullptr, *x = nullptr, *y = nullptr;
    __half *h = nullptr;               // [row][16] head-conv output of the debug direct path only
    __half *pbuf = nullptr;            // [row][8] policy samples = A operand of the policy FC GEMM
    __half *vbuf = nullptr;            // [row][2] value samples
    float* part = nullptr;             // [K split][batch rounded to 128][384] fp32 partial sums
    CUtensorMap tm_pa;                 // [batch][3200] view of pbuf
    __half *d_policy = nullptr, *d_value = nullptr;
    uint8_t* h_legal = nullptr;       // pinned: legal masks of the raw-position path
    float* h_prior = nullptr;         // pinned: priors of the raw-position path
    bool want_plan = false;           // the next raw feature launch also derives candidates / orbit representatives
    uint8_t* h_in = nullptr;          // pinned staging
    __half *h_policy = nullptr, *h_value = nullptr;
    CUtensorMap tm_feat, tm_x, tm_y;           // 170-row load windows
    uint32_t* done = nullptr;          // per-tile progress flags of the persistent tower kernel
    uint32_t flag_gen = 1u << 20;      // eager launches count up from here; a graph launch zeroes the flags and uses kGraphGen
    int resident_batch = 0;
    int resident_kind = 0;            // 0 none, 1 raw features, 2 compact positions
    bool busy = false;
    // leaf batch (dg_leaf_batch): producers claim slots of h_in lock-free, one graph launch per submit, completion flag
    dg_engine* owner = nullptr;
    std::atomic<int32_t> fill{0};      // slots claimed so far
    std::atomic<int32_t> committed{0}; // slots whose 384 bytes are complete
    std::atomic<uint32_t> polls{0};    // dg_leaf_batch_ready calls (every 65536th asks the driver whether the stream failed)
    int32_t submitted = 0;             // leaves of the submit in flight / last completed
    uint32_t* h_flag = nullptr;        // pinned: 0 while a submit is in flight, 1 once its results are in the pinned outputs
    std::map<uint32_t, cudaGraphExec_t> graphs;   // (bucket << 1 | want_prior) -> instantiated forward
    bool capturing = false;
    Workspace() = default;
    Workspace(const Workspace&) = delete;
    Workspace& operator=(const Workspace&) = delete;
};
constexpr uint32_t kGraphGen = 64;     // flag base of a captured tower launch (the graph zeroes the flags first)
constexpr int kBucket = 16;            // captured forwards exist for batches rounded up to a multiple of this

struct ConvWeights {
    __half* w = nullptr;              // [9][ntot][k] fp16
    float* bias = nullptr;            // [ntot] fp32
    CUtensorMap tm;
};

struct DeviceNet {
    int num_blocks = 0;
    ConvWeights up;
    std::vector<ConvWeights> c1, c2;
    std::vector<float> gate;
    ConvWeights heads;
    CUtensorMap tm_heads_pair;        // heads.w as [8 out][64 in] boxes (one CTA's half) for the tower kernel's last layer
    __half* w_pfc = nullptr;          // [384 out][3200 = 400 board rows x 8 samples] (zero for halo rows / padding)
    CUtensorMap tm_pfc;
    float* b_pfc = nullptr;           // fp16(tau * b) as fp32
    __half* w_vfc = nullptr;          // [400 board rows][2 samples] (zero for halo rows)
    float b_vfc = 0.f;
    float tau = 1.f;
    std::vector<void*> allocations;
    bool loaded = false;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

struct dg_engine {
    dg_engine_config cfg;
    int num_sms = 0;
    EncodeTiledFn encode = nullptr;
    std::vector<std::unique_ptr<Workspace>> ws;
    std::mutex ws_mutex;
    std::condition_variable ws_cv;
    DeviceNet net;
    std::atomic<bool> kernels_configured{false};   // a raw forward has run once outside a capture (one-time kernel attributes)
    std::mutex graph_mutex;           // captures, and the first (eager) forward
    std::mutex err_mutex;
    std::string last_error;
    std::vector<void*> host_allocs;
    void* flush_buf = nullptr;
    long long* trace_buf = nullptr;   // set only inside dg_engine_debug_conv_trace
};

namespace {

int32_t fail(dg_engine* e, int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (e) {
        std::lock_guard<std::mutex> g(e->err_mutex);
        e->last_error = buf;
    }
    return code;
}

#define DG_CUDA(e, call)                                                                              \
    do {                                                                                              \
        cudaError_t err__ = (call);                                                                   \
        if (err__ != cudaSuccess)                                                                     \
            return fail((e), DG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
    } while (0)

bool make_tmap(dg_engine* e, CUtensorMap* m, void* base, uint64_t inner, uint64_t rows, uint32_t box_rows) {
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {inner * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = e->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// ------------------------------------------------------------------------------------ workspaces

int32_t create_workspace(dg_engine* e, Workspace& w) {
    const int mb = e->cfg.max_batch;
    const size_t rows = static_cast<size_t>(dg_alloc_rows(mb));
    DG_CUDA(e, cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
    DG_CUDA(e, cudaEventCreate(&w.ev0));
    DG_CUDA(e, cudaEventCreate(&w.ev1));
    DG_CUDA(e, cudaMalloc(&w.d_in, static_cast<size_t>(mb) * kFeatBytes));
    DG_CUDA(e, cudaMalloc(&w.feat, rows * 64 * 2));
    DG_CUDA(e, cudaMalloc(&w.x, rows * kChan * 2));
    DG_CUDA(e, cudaMalloc(&w.y, rows * kChan * 2));
    DG_CUDA(e, cudaMalloc(&w.h, rows * kHeadChan * 2));
    DG_CUDA(e, cudaMalloc(&w.d_policy, static_cast<size_t>(mb) * DG_POLICY_SIZE * 2));
    DG_CUDA(e, cudaMalloc(&w.d_value, static_cast<size_t>(mb) * 2));
    DG_CUDA(e, cudaMemset(w.feat, 0, rows * 64 * 2));
    DG_CUDA(e, cudaMemset(w.x, 0, rows * kChan * 2));
    DG_CUDA(e, cudaMemset(w.y, 0, rows * kChan * 2));
    DG_CUDA(e, cudaMemset(w.h, 0, rows * kHeadChan * 2));
    DG_CUDA(e, cudaMalloc(&w.pbuf, rows * 8 * 2));
    DG_CUDA(e, cudaMalloc(&w.vbuf, rows * 2 * 2));
    DG_CUDA(e, cudaMemset(w.pbuf, 0, rows * 8 * 2));
    DG_CUDA(e, cudaMemset(w.vbuf, 0, rows * 2 * 2));
    DG_CUDA(e, cudaMalloc(&w.part, static_cast<size_t>(dg::kPolicyFcSplit) * ((mb + 127) / 128 * 128) * dg::kPolicyFcN * 4));
    DG_CUDA(e, cudaMalloc(&w.done, (static_cast<size_t>(dg_num_tiles(mb)) + 2) * 4));
    DG_CUDA(e, cudaMemset(w.done, 0, (static_cast<size_t>(dg_num_tiles(mb)) + 2) * 4));
    DG_CUDA(e, cudaHostAlloc(&w.h_in, static_cast<size_t>(mb) * kFeatBytes, cudaHostAllocDefault));
    DG_CUDA(e, cudaHostAlloc(&w.h_policy, static_cast<size_t>(mb) * DG_POLICY_SIZE * 2, cudaHostAllocDefault));
    DG_CUDA(e, cudaHostAlloc(&w.h_value, static_cast<size_t>(mb) * 2, cudaHostAllocDefault));
    DG_CUDA(e, cudaHostAlloc(&w.h_legal, static_cast<size_t>(mb) * 361, cudaHostAllocDefault));
    DG_CUDA(e, cudaHostAlloc(&w.h_prior, static_cast<size_t>(mb) * 368 * 4, cudaHostAllocDefault));
    DG_CUDA(e, cudaHostAlloc(&w.h_flag, 64, cudaHostAllocDefault));
    *w.h_flag = 1;
    memset(w.h_in, 0, static_cast<size_t>(mb) * kFeatBytes);
    w.owner = e;
    if (!make_tmap(e, &w.tm_feat, w.feat, 64, rows, DG_WINDOW_ROWS) || !make_tmap(e, &w.tm_x, w.x, kChan, rows, DG_WINDOW_ROWS) ||
        !make_tmap(e, &w.tm_y, w.y, kChan, rows, DG_WINDOW_ROWS) ||
        !make_tmap(e, &w.tm_pa, w.pbuf + static_cast<size_t>(DG_GUARD_ROWS) * 8, dg::kPolicyFcK, mb, 128))
        return fail(e, DG_ERR_CUDA, "cuTensorMapEncodeTiled failed for an activation buffer");
    return DG_OK;
}

void destroy_workspace(Workspace& w) {
    if (w.stream) cudaStreamDestroy(w.stream);
    if (w.ev0) cudaEventDestroy(w.ev0);
    if (w.ev1) cudaEventDestroy(w.ev1);
    cudaFree(w.d_in); cudaFree(w.feat); cudaFree(w.x); cudaFree(w.y); cudaFree(w.h);
    cudaFree(w.d_policy); cudaFree(w.d_value); cudaFree(w.done); cudaFree(w.pbuf); cudaFree(w.vbuf); cudaFree(w.part);
    cudaFreeHost(w.h_in); cudaFreeHost(w.h_policy); cudaFreeHost(w.h_value); cudaFreeHost(w.h_legal); cudaFreeHost(w.h_prior);
    cudaFreeHost(w.h_flag);
    for (auto& g : w.graphs) cudaGraphExecDestroy(g.second);
}

Workspace* acquire(dg_engine* e) {
    std::unique_lock<std::mutex> lk(e->ws_mutex);
    for (;;) {
        for (auto& w : e->ws)
            if (!w->busy) { w->busy = true; return w.get(); }
        e->ws_cv.wait(lk);
    }
}
void release(dg_engine* e, Workspace* w) {
    { std::lock_guard<std::mutex> lk(e->ws_mutex); w->busy = false; }
    e->ws_cv.notify_one();
}
// The workspace that still holds `batch` resident inputs (measurement / debug hooks only).
Workspace* acquire_resident(dg_engine* e, int batch) {
    std::unique_lock<std::mutex> lk(e->ws_mutex);
    for (;;) {
        bool any = false;
        for (auto& w : e->ws) {
            if (w->resident_kind == 0 || w->resident_batch != batch) continue;
            any = true;
            if (!w->busy) { w->busy = true; return w.get(); }
        }
        if (!any) return nullptr;
        e->ws_cv.wait(lk);
    }
}
struct WsGuard {       // mirrors WorkspaceGuard's return-to-pool-on-drop (network.rs:73-79)
    dg_engine* e; Workspace* w;
    WsGuard(dg_engine* e_) : e(e_), w(acquire(e_)) {}
    WsGuard(dg_engine* e_, int resident_batch) : e(e_), w(acquire_resident(e_, resident_batch)) {}
    ~WsGuard() { if (w) release(e, w); }
};

// ------------------------------------------------------------------------------------ weights

void free_net(DeviceNet& n) {
    for (void* p : n.allocations) cudaFree(p);
    n = DeviceNet();
}

template <typename T>
int32_t upload(dg_engine* e, DeviceNet& n, const std::vector<T>& host, T** dev) {
    DG_CUDA(e, cudaMalloc(dev, host.size() * sizeof(T)));
    n.allocations.push_back(*dev);
    DG_CUDA(e, cudaMemcpy(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    return DG_OK;
}

const dg::HostTensor* find(const dg::TensorMap& t, const std::string& name, const char* dtype, size_t min_elems) {
    auto it = t.find(name);
    if (it == t.end()) return nullptr;
    const size_t esz = (it->second.dtype == "f2") ? 2 : (it->second.dtype == "i1") ? 1 : 4;
    if (it->second.dtype != dtype || it->second.bytes.size() / esz < min_elems) return nullptr;
    return &it->second;
}

float h2f(uint16_t bits) { __half_raw r; r.x = bits; return __half2float(__half(r)); }
uint16_t f2h(float f) { __half h = __float2half_rn(f); return __half_raw(h).x; }

// KRSC fp16 [cout][3][3][cin] -> [tap][ntot][kpad] at output-channel offset `n0` (zero elsewhere).
void relayout_conv(const uint16_t* krsc, int cout, int cin, int ntot, int kpad, int n0, std::vector<uint16_t>& dst) {
    for (int k = 0; k < cout; k++)
        for (int tap = 0; tap < 9; tap++)
            for (int c = 0; c < cin; c++)
                dst[(static_cast<size_t>(tap) * ntot + n0 + k) * kpad + c] = krsc[(static_cast<size_t>(k) * 9 + tap) * cin + c];
}

int32_t build_conv(dg_engine* e, DeviceNet& n, ConvWeights& cw, const std::vector<uint16_t>& w, const std::vector<float>& bias,
                   int ntot, int kpad, int box_rows) {
    uint16_t* dw = nullptr;
    int32_t rc = upload<uint16_t>(e, n, w, &dw);
    if (rc) return rc;
    cw.w = reinterpret_cast<__half*>(dw);
    rc = upload<float>(e, n, bias, &cw.bias);
    if (rc) return rc;
    if (!make_tmap(e, &cw.tm, cw.w, kpad, 9ull * ntot, box_rows)) return fail(e, DG_ERR_CUDA, "cuTensorMapEncodeTiled failed for a filter");
    return DG_OK;
}

// Restates Builder::get_workspace's tensor lookups (graph.rs:50-96) once, at load time.
int32_t load_net(dg_engine* e, const dg::TensorMap& t) {
    if (t.empty()) return fail(e, DG_ERR_MISSING_WEIGHTS, "no tensors");
    auto scalar_i4 = [&](const char* name, int dflt) {
        const dg::HostTensor* s = find(t, name, "i4", 1);
        int v = dflt;
        if (s) memcpy(&v, s->bytes.data(), 4);
        return v;
    };
    const int channels = scalar_i4("num_channels:0", 128);   // layers/common.rs:22-50
    const int samples = scalar_i4("num_samples:0", 8);
    if (channels != kChan || samples != 8)
        return fail(e, DG_ERR_KERNEL, "unsupported network shape: %d channels / %d samples (this engine is built for 128 / 8)", channels, samples);

    DeviceNet n;
    int32_t rc = DG_OK;
    char name[96];
    auto cleanup = [&](int32_t code) { free_net(n); return code; };

    // 01_upsample (up_block.rs:42-44)
    {
        const dg::HostTensor* w = find(t, "01_upsample/conv_1:0", "f2", 128 * 9 * 32);
        const dg::HostTensor* b = find(t, "01_upsample/conv_1/offset:0", "f2", 128);
        if (!w || !b) return cleanup(fail(e, DG_ERR_MISSING_WEIGHTS, "01_upsample tensors missing or of wrong type/size"));
        std::vector<uint16_t> wl(9 * 128 * 64, 0);
        relayout_conv(reinterpret_cast<const uint16_t*>(w->bytes.data()), 128, 32, 128, 64, 0, wl);
        std::vector<float> bl(128);
        for (int i = 0; i < 128; i++) bl[i] = h2f(reinterpret_cast<const uint16_t*>(b->bytes.data())[i]);
        if ((rc = build_conv(e, n, n.up, wl, bl, 128, 64, 64))) return cleanup(rc);
    }
    // NN_residual (residual_block.rs:40-78); discovery stops at the first missing block (graph.rs:76-96)
    for (int i = 0;; i++) {
        const dg::HostTensor *w[2], *b[2];
        bool ok = true;
        for (int j = 0; j < 2; j++) {
            snprintf(name, sizeof name, "%02d_residual/conv_%d:0", i + 2, j + 1);
            w[j] = find(t, name, "f2", 128 * 9 * 128);
            snprintf(name, sizeof name, "%02d_residual/conv_%d/offset:0", i + 2, j + 1);
            b[j] = find(t, name, "f2", 128);
            ok = ok && w[j] && b[j];
        }
        if (!ok) break;
        snprintf(name, sizeof name, "%02d_residual/alpha:0", i + 2);
        float g = 0.5f;                                             // residual_block.rs:43,50
        if (const dg::HostTensor* a = find(t, name, "f4", 1)) memcpy(&g, a->bytes.data(), 4);
        n.gate.push_back(g);
        n.c1.emplace_back();
        n.c2.emplace_back();
        for (int j = 0; j < 2; j++) {
            std::vector<uint16_t> wl(9 * 128 * 128, 0);
            relayout_conv(reinterpret_cast<const uint16_t*>(w[j]->bytes.data()), 128, 128, 128, 128, 0, wl);
            std::vector<float> bl(128);
            for (int k = 0; k < 128; k++) {
                const float bv = h2f(reinterpret_cast<const uint16_t*>(b[j]->bytes.data())[k]);
                // conv_2's offset is scaled by the gate in place, in fp16, once (residual_block.rs:72-74)
                bl[k] = (j == 0) ? bv : h2f(f2h(static_cast<float>(static_cast<double>(g) * static_cast<double>(bv))));
            }
            if ((rc = build_conv(e, n, j == 0 ? n.c1.back() : n.c2.back(), wl, bl, 128, 128, 64))) return cleanup(rc);
        }
        n.num_blocks++;
    }
    // heads (policy_head.rs:43-103, value_head.rs:40-84); layer index = 2 + num_blocks (graph.rs:55)
    const int hidx = 2 + n.num_blocks;
    {
        snprintf(name, sizeof name, "%02dp_policy/conv_1:0", hidx);
        const dg::HostTensor* pw = find(t, name, "f2", 8 * 9 * 128);
        snprintf(name, sizeof name, "%02dp_policy/conv_1/offset:0", hidx);
        const dg::HostTensor* pb = find(t, name, "f2", 8);
        snprintf(name, sizeof name, "%02dv_value/conv_1:0", hidx);
        const dg::HostTensor* vw = find(t, name, "f2", 2 * 9 * 128);
        snprintf(name, sizeof name, "%02dv_value/conv_1/offset:0", hidx);
        const dg::HostTensor* vb = find(t, name, "f2", 2);
        snprintf(name, sizeof name, "%02dp_policy/linear_1:0", hidx);
        const dg::HostTensor* pl = find(t, name, "f2", 2888 * 362);
        snprintf(name, sizeof name, "%02dp_policy/linear_1/offset:0", hidx);
        const dg::HostTensor* plb = find(t, name, "f2", 362);
        snprintf(name, sizeof name, "%02dv_value/linear_2:0", hidx);
        const dg::HostTensor* vl = find(t, name, "f2", 722);
        snprintf(name, sizeof name, "%02dv_value/linear_2/offset:0", hidx);
        const dg::HostTensor* vlb = find(t, name, "f2", 1);
        if (!pw || !pb || !vw || !vb || !pl || !plb || !vl || !vlb)
            return cleanup(fail(e, DG_ERR_MISSING_WEIGHTS, "head tensors (%02dp_policy / %02dv_value) missing or of wrong type/size", hidx, hidx));
        std::vector<uint16_t> wl(9 * kHeadChan * 128, 0);
        relayout_conv(reinterpret_cast<const uint16_t*>(pw->bytes.data()), 8, 128, kHeadChan, 128, 0, wl);
        relayout_conv(reinterpret_cast<const uint16_t*>(vw->bytes.data()), 2, 128, kHeadChan, 128, 8, wl);
        std::vector<float> bl(kChan, 0.f);       // (128 entries: the tower kernel stages a layer's bias as one 128-float vector)
        for (int i = 0; i < 8; i++) bl[i] = h2f(reinterpret_cast<const uint16_t*>(pb->bytes.data())[i]);
        for (int i = 0; i < 2; i++) bl[8 + i] = h2f(reinterpret_cast<const uint16_t*>(vb->bytes.data())[i]);
        if ((rc = build_conv(e, n, n.heads, wl, bl, kHeadChan, 128, kHeadChan))) return cleanup(rc);
        // the same filter bank as the last layer of the persistent tower kernel: each CTA of a pair holds 8 of the 16 channels
        if (!make_tmap(e, &n.tm_heads_pair, n.heads.w, 128, 9ull * kHeadChan, kHeadChan / 2))
            return cleanup(fail(e, DG_ERR_CUDA, "cuTensorMapEncodeTiled failed for the head filter"));

        const float temperature = e->cfg.softmax_temperature > 0.f ? e->cfg.softmax_temperature : 0.709888f;
        n.tau = 1.0f / temperature;                                  // policy_head.rs:46
        // policy FC weight, file order [2888 in = 8*point + sample][362 out] (dense.py:32-39), re-laid out K-major
        // over board ROWS: w'[out][8*row + sample], zero where the row is a halo row
        const uint16_t* plw = reinterpret_cast<const uint16_t*>(pl->bytes.data());
        std::vector<uint16_t> pfc(static_cast<size_t>(dg::kPolicyFcN) * dg::kPolicyFcK, 0);
        for (int y = 0; y < 19; y++)
            for (int x = 0; x < 19; x++)
                for (int sm = 0; sm < 8; sm++) {
                    const int in = 8 * (19 * y + x) + sm, kk = 8 * (DG_LINE_STRIDE * y + x) + sm;
                    for (int o = 0; o < 362; o++) pfc[static_cast<size_t>(o) * dg::kPolicyFcK + kk] = plw[static_cast<size_t>(in) * 362 + o];
                }
        uint16_t* d = nullptr;
        if ((rc = upload<uint16_t>(e, n, pfc, &d))) return cleanup(rc);
        n.w_pfc = reinterpret_cast<__half*>(d);
        if (!make_tmap(e, &n.tm_pfc, n.w_pfc, dg::kPolicyFcK, dg::kPolicyFcN, 128)) return cleanup(fail(e, DG_ERR_CUDA, "cuTensorMapEncodeTiled failed for the policy FC weight"));
        std::vector<float> pfb(362);
        for (int i = 0; i < 362; i++)   // offset scaled by tau in place, in fp16 (policy_head.rs:87-89)
            pfb[i] = h2f(f2h(static_cast<float>(static_cast<double>(n.tau) *
                                                 static_cast<double>(h2f(reinterpret_cast<const uint16_t*>(plb->bytes.data())[i])))));
        if ((rc = upload<float>(e, n, pfb, &n.b_pfc))) return cleanup(rc);
        const uint16_t* vlw = reinterpret_cast<const uint16_t*>(vl->bytes.data());
        std::vector<uint16_t> vfc(DG_POS_ROWS * 2, 0);
        for (int y = 0; y < 19; y++)
            for (int x = 0; x < 19; x++)
                for (int sm = 0; sm < 2; sm++) vfc[2 * (DG_LINE_STRIDE * y + x) + sm] = vlw[2 * (19 * y + x) + sm];
        if ((rc = upload<uint16_t>(e, n, vfc, &d))) return cleanup(rc);
        n.w_vfc = reinterpret_cast<__half*>(d);
        n.b_vfc = h2f(reinterpret_cast<const uint16_t*>(vlb->bytes.data())[0]);
    }
    n.loaded = true;
    free_net(e->net);
    e->net = std::move(n);
    return DG_OK;
}

// ------------------------------------------------------------------------------------ forward

int32_t run_conv(dg_engine* e, Workspace& w, ConvTcShape shape, const CUtensorMap& tm_in, const __half* in, int cin,
                 const ConvWeights& cw, int ntot, __half* out, int out_stride, const __half* skip, float alpha, float beta, int batch,
                 __half* out2 = nullptr) {
    if (e->cfg.flags & DG_FLAG_DEBUG_DIRECT_CONV) {
        DG_CUDA(e, dg::launch_conv_direct(in, cin, cw.w, ntot, cw.bias, alpha, beta, skip, kChan, out, out_stride, batch, w.stream));
        return DG_OK;
    }
    ConvTcParams p;
    p.ntiles = dg_num_tiles(batch);
    p.valid_rows = batch * DG_POS_ROWS;
    p.out = out;
    p.out_stride = out_stride;
    p.out2 = out2;
    p.skip = skip;
    p.skip_stride = kChan;
    p.bias = cw.bias;
    p.alpha = alpha;
    p.beta = beta;
    p.trace = e->trace_buf;
    const bool pdl = !(e->cfg.flags & DG_FLAG_NO_PDL);
    if (shape == ConvTcShape::kHeads)
        DG_CUDA(e, dg::launch_conv_tc(shape, tm_in, cw.tm, p, e->num_sms, w.stream, pdl));
    else   // tower and up-sampling layers run on CTA pairs
        DG_CUDA(e, dg::launch_conv_pair(shape == ConvTcShape::kUp ? 1 : 2, tm_in, cw.tm, p, e->num_sms, w.stream, pdl));
    return DG_OK;
}

// Waits for everything enqueued on the workspace's stream.  By default the calling thread spins in the driver
// (cudaStreamSynchronize: lowest latency, one core per waiter).  With DG_FLAG_BLOCKING_SYNC it polls the stream between
// short naps instead: self-play runs as many host threads as cores and a spinning waiter steals one (measured on a B200
// with 4 host cores: 418 k evaluations/s spinning, 471 k sleeping on a cudaEventBlockingSync event, 490 k with naps; with
// 16 cores naps equal spinning, the blocking event loses 10 %).
inline cudaError_t wait_stream(dg_engine* e, Workspace& w) {
    if (!(e->cfg.flags & DG_FLAG_BLOCKING_SYNC)) return cudaStreamSynchronize(w.stream);
    static thread_local bool slack_set = false;
    if (!slack_set) { prctl(PR_SET_TIMERSLACK, 1000UL, 0, 0, 0); slack_set = true; }   // naps of ~20 us, not 20 + 50 us
    for (;;) {
        cudaError_t q = cudaStreamQuery(w.stream);
        if (q != cudaErrorNotReady) return q;
        timespec nap{0, 20000};
        nanosleep(&nap, nullptr);
    }
}

// Raw-position path: d_in (max_batch x 23,104 B) holds [raw positions | compact planes | legal masks].
inline uint8_t* raw_planes(dg_engine* e, Workspace& w) { return w.d_in + static_cast<size_t>(e->cfg.max_batch) * 512; }
inline uint8_t* raw_legal(dg_engine* e, Workspace& w) { return w.d_in + static_cast<size_t>(e->cfg.max_batch) * 2048; }
inline uint8_t* raw_cand(dg_engine* e, Workspace& w) { return w.d_in + static_cast<size_t>(e->cfg.max_batch) * 2560; }
inline uint8_t* raw_rep(dg_engine* e, Workspace& w) { return w.d_in + static_cast<size_t>(e->cfg.max_batch) * 3072; }
inline float* raw_prior(dg_engine* e, Workspace& w) { return reinterpret_cast<float*>(w.d_in + static_cast<size_t>(e->cfg.max_batch) * 4096); }

// Enqueues pack + tower (+ heads) of the batch resident in w.d_in.  blocks < 0 = whole network.
// stage: 0 = everything, 1 = pack only, 2 = residual convolutions only (timing), 3 = everything but pack.
int32_t enqueue_network(dg_engine* e, Workspace& w, int batch, int blocks, int stage) {
    const DeviceNet& n = e->net;
    int32_t rc;
    if (stage == 0 || stage == 1) {
        if (w.resident_kind == 1) DG_CUDA(e, dg::launch_pack_features(w.d_in, w.feat, batch, w.stream));
        else if (w.resident_kind == 3) {
            // raw positions at d_in; compact planes and legal masks behind them in the same buffer
            // (the kernel also writes the tower's input rows: no separate pack launch)
            DG_CUDA(e, dg::launch_planes_from_stones(w.d_in, raw_planes(e, w), raw_legal(e, w), w.feat, w.want_plan ? raw_cand(e, w) : nullptr,
                                                     w.want_plan ? raw_rep(e, w) : nullptr, batch, w.stream));
        } else DG_CUDA(e, dg::launch_pack_compact(w.d_in, w.feat, batch, w.stream));
        if (stage == 1) return DG_OK;
    }
    const int nb = (blocks < 0 || blocks > n.num_blocks) ? n.num_blocks : blocks;
    const bool layerwise = (e->cfg.flags & (DG_FLAG_DEBUG_DIRECT_CONV | DG_FLAG_LAYERWISE)) != 0;
    const bool heads_in_tower = !layerwise && blocks < 0 && stage != 2 && !(e->cfg.flags & DG_FLAG_SEPARATE_HEAD_CONV);
    if (!layerwise) {
        // the whole tower in one persistent launch (tower_kernel)
        dg::TowerParams tp;
        tp.act[0] = w.tm_feat; tp.act[1] = w.tm_x; tp.act[2] = w.tm_y;
        int nl = 0;
        auto add = [&](int in_map, int nh, const ConvWeights& cw, __half* out, const __half* skip, float alpha, float beta) {
            tp.w[nl] = cw.tm;
            tp.layer[nl] = dg::TowerLayer{in_map, nh, skip != nullptr, alpha, beta, cw.bias, out, skip, kChan, nullptr};
            nl++;
        };
        if (stage != 2) add(0, 1, n.up, w.x, nullptr, 1.f, 0.f);
        for (int i = 0; i < nb; i++) {
            const float g = n.gate[i];
            add(1, 2, n.c1[i], w.y, nullptr, 1.f, 0.f);
            add(2, 2, n.c2[i], w.x, w.x, g, 1.0f - g);
        }
        if (heads_in_tower) {      // the head convolution (policy + value samples) as the last layer of the same launch
            tp.w[nl] = n.tm_heads_pair;
            tp.layer[nl] = dg::TowerLayer{1, 2, 0, 1.f, 0.f, n.heads.bias, w.pbuf, nullptr, kHeadChan, w.vbuf};
            nl++;
        }
        if (nl > 0) {
            tp.nlayers = nl;
            tp.ntiles = 3 * batch;                    // position-aligned tiles (conv_tc2.cu: tower_tile_base)
            tp.valid_rows = batch * DG_POS_ROWS;
            const int nunits = (tp.ntiles + 1) / 2;
            const int pairs = std::min(e->num_sms / 2, nunits);
            tp.rot = (e->cfg.flags & DG_FLAG_NO_ROTATE) ? 0 : nunits % pairs;
            if (w.capturing) {                        // a graph replays the same parameters: zero the flags, fixed base
                DG_CUDA(e, cudaMemsetAsync(w.done, 0, (static_cast<size_t>(dg_num_tiles(e->cfg.max_batch)) + 2) * 4, w.stream));
                tp.gen = kGraphGen;
            } else {
                w.flag_gen += 64;                     // > layers per launch; flags compare modulo 2^32
                if (w.flag_gen < (1u << 20)) {        // wrapped: stay above what a graph launch leaves in the flags
                    w.flag_gen = 1u << 20;
                    DG_CUDA(e, cudaMemsetAsync(w.done, 0, (static_cast<size_t>(dg_num_tiles(e->cfg.max_batch)) + 2) * 4, w.stream));
                }
                tp.gen = w.flag_gen;
            }
            tp.done = w.done;
            tp.trace = e->trace_buf;
            tp.late_a = (e->cfg.flags & DG_FLAG_TOWER_LATE_A) ? 1 : 0;
            DG_CUDA(e, dg::launch_tower(tp, e->num_sms, w.stream));
        }
    } else {
        if (stage != 2)
            if ((rc = run_conv(e, w, ConvTcShape::kUp, w.tm_feat, w.feat, 64, n.up, 128, w.x, kChan, nullptr, 1.f, 0.f, batch))) return rc;
        for (int i = 0; i < nb; i++) {
            const float g = n.gate[i];
            if ((rc = run_conv(e, w, ConvTcShape::kTower, w.tm_x, w.x, kChan, n.c1[i], 128, w.y, kChan, nullptr, 1.f, 0.f, batch))) return rc;
            if ((rc = run_conv(e, w, ConvTcShape::kTower, w.tm_y, w.y, kChan, n.c2[i], 128, w.x, kChan, w.x, g, 1.0f - g, batch))) return rc;
        }
    }
    if (blocks >= 0 || stage == 2) return DG_OK;
    if (e->cfg.flags & DG_FLAG_DEBUG_DIRECT_CONV) {
        if ((rc = run_conv(e, w, ConvTcShape::kHeads, w.tm_x, w.x, kChan, n.heads, kHeadChan, w.h, kHeadChan, nullptr, 1.f, 0.f, batch))) return rc;
        DG_CUDA(e, dg::launch_split_heads(w.h, w.pbuf, w.vbuf, batch, w.stream));
    } else if (!heads_in_tower) {
        if ((rc = run_conv(e, w, ConvTcShape::kHeads, w.tm_x, w.x, kChan, n.heads, kHeadChan, w.pbuf, 8, nullptr, 1.f, 0.f, batch, w.vbuf))) return rc;
    }
    DG_CUDA(e, dg::launch_policy_fc(w.tm_pa, n.tm_pfc, w.part, batch, w.stream));
    DG_CUDA(e, dg::launch_heads_finish(w.part, batch, n.b_pfc, n.tau, w.vbuf, n.w_vfc, n.b_vfc, w.d_policy, w.d_value, w.stream));
    return DG_OK;
}

bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int32_t forward_impl(dg_engine* e, const void* input, size_t bytes_per_pos, int kind, int batch, uint16_t* value_out, uint16_t* policy_out) {
    if (!e) return DG_ERR_INVALID_ARGUMENT;
    if (!input || !value_out || !policy_out) return fail(e, DG_ERR_INVALID_ARGUMENT, "null buffer");
    if (batch < 1 || batch > e->cfg.max_batch) return fail(e, DG_ERR_INVALID_ARGUMENT, "batch %d outside 1..%d", batch, e->cfg.max_batch);
    if (!e->net.loaded) return fail(e, DG_ERR_MISSING_WEIGHTS, "no weights loaded");
    DG_CUDA(e, cudaSetDevice(e->cfg.device));
    WsGuard guard(e);
    Workspace& w = *guard.w;
    const size_t in_bytes = bytes_per_pos * batch;
    const void* src = input;
    if (!is_pinned(input)) { memcpy(w.h_in, input, in_bytes); src = w.h_in; }
    DG_CUDA(e, cudaMemcpyAsync(w.d_in, src, in_bytes, cudaMemcpyHostToDevice, w.stream));
    w.resident_kind = kind;
    w.resident_batch = batch;
    int32_t rc = enqueue_network(e, w, batch, -1, 0);
    if (rc) { cudaStreamSynchronize(w.stream); return rc; }
    const bool pv = is_pinned(value_out), pp = is_pinned(policy_out);
    cudaError_t cerr = cudaMemcpyAsync(pv ? static_cast<void*>(value_out) : w.h_value, w.d_value, static_cast<size_t>(batch) * 2, cudaMemcpyDeviceToHost, w.stream);
    if (cerr == cudaSuccess)
        cerr = cudaMemcpyAsync(pp ? static_cast<void*>(policy_out) : w.h_policy, w.d_policy, static_cast<size_t>(batch) * DG_POLICY_SIZE * 2, cudaMemcpyDeviceToHost, w.stream);
    if (cerr != cudaSuccess) cudaStreamSynchronize(w.stream);     // the workspace goes back to the pool: nothing of this call may still run
    else cerr = wait_stream(e, w);
    if (cerr != cudaSuccess) return fail(e, DG_ERR_CUDA, "forward failed: %s", cudaGetErrorString(cerr));
    if (!pv) memcpy(value_out, w.h_value, static_cast<size_t>(batch) * 2);
    if (!pp) memcpy(policy_out, w.h_policy, static_cast<size_t>(batch) * DG_POLICY_SIZE * 2);
    return DG_OK;
}

}  // namespace

// ======================================================================================= C ABI

extern "C" {

int32_t dg_engine_abi_version(void) { return 1; }

int32_t dg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int usable = 0;
    for (; usable < n; usable++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, usable) != cudaSuccess || major != 10) break;
    }
    return usable;
}

int32_t dg_current_device(void) {
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); return -1; }
    return d;
}

int32_t dg_set_current_device(int32_t device) {
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return DG_ERR_CUDA; }
    return DG_OK;
}

int32_t dg_engine_create(const dg_engine_config* config, dg_engine** out) {
    if (!config || !out) return DG_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (config->max_batch < 1 || config->max_batch > 8192) return DG_ERR_INVALID_ARGUMENT;
    dg_engine* e = new dg_engine();
    e->cfg = *config;
    if (e->cfg.num_workspaces <= 0) e->cfg.num_workspaces = 2;
    *out = e;      // returned even on failure so the caller can read dg_engine_last_error, then destroy
    DG_CUDA(e, cudaSetDevice(config->device));
    cudaDeviceProp prop;
    DG_CUDA(e, cudaGetDeviceProperties(&prop, config->device));
    if (prop.major != 10)
        return fail(e, DG_ERR_CUDA, "device %d is sm_%d%d; this engine only runs on sm_100 (B200) -- there is no fallback path",
                    config->device, prop.major, prop.minor);
    e->num_sms = prop.multiProcessorCount;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    DG_CUDA(e, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(e, DG_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
    e->encode = reinterpret_cast<EncodeTiledFn>(fn);
    {   // constants of the device feature kernel: the engine's zobrist keys and the 8 symmetry maps (csrc/go_board.h)
        static_assert(sizeof(dg_raw_position) == 384, "dg_raw_position is 384 bytes");
        const dg::Tables& T = dg::tables();
        std::vector<unsigned long long> z(2 * 361);
        std::vector<uint16_t> sym(8 * 361);
        for (int c = 0; c < 2; c++)
            for (int p = 0; p < 361; p++) z[c * 361 + p] = T.zobrist[c + 1][p];
        for (int t = 0; t < 8; t++)
            for (int p = 0; p < 361; p++) sym[t * 361 + p] = T.sym[t][p];
        DG_CUDA(e, dg::upload_feature_tables(z.data(), sym.data()));
    }
    for (int i = 0; i < e->cfg.num_workspaces; i++) {
        e->ws.emplace_back(new Workspace());
        int32_t rc = create_workspace(e, *e->ws.back());
        if (rc) return rc;
    }
    return DG_OK;
}

void dg_engine_destroy(dg_engine* e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    cudaDeviceSynchronize();
    for (auto& w : e->ws) destroy_workspace(*w);
    free_net(e->net);
    for (void* p : e->host_allocs) cudaFreeHost(p);
    cudaFree(e->flush_buf);
    delete e;
}

const char* dg_engine_last_error(dg_engine* e) {
    if (!e) return "null engine";
    // a copy per calling thread: a concurrent failing call on the same engine may replace the engine's text at any time
    static thread_local std::string text;
    std::lock_guard<std::mutex> g(e->err_mutex);
    text = e->last_error;
    return text.c_str();
}

int32_t dg_engine_num_blocks(dg_engine* e) { return e ? e->net.num_blocks : 0; }
int32_t dg_engine_max_batch(dg_engine* e) { return e ? e->cfg.max_batch : 0; }
int32_t dg_engine_num_workspaces(dg_engine* e) { return e ? e->cfg.num_workspaces : 0; }

int32_t dg_engine_load_weights_json(dg_engine* e, const char* path) {
    if (!e || !path) return DG_ERR_INVALID_ARGUMENT;
    dg::TensorMap t;
    std::string why;
    int rc = dg::load_weights_file(path, t, why);
    if (rc == 1) return fail(e, DG_ERR_MISSING_WEIGHTS, "%s: %s", path, why.c_str());
    if (rc == 2) return fail(e, DG_ERR_MALFORMED_WEIGHTS, "%s: %s", path, why.c_str());
    DG_CUDA(e, cudaSetDevice(e->cfg.device));
    return load_net(e, t);
}

int32_t dg_weights_file_probe(const char* path, const char* name, int32_t* num_tensors, float* scale, uint64_t* nbytes) {
    if (!path) return DG_ERR_INVALID_ARGUMENT;
    dg::TensorMap t;
    std::string why;
    const int rc = dg::load_weights_file(path, t, why);
    if (num_tensors) *num_tensors = static_cast<int32_t>(t.size());
    if (scale) *scale = 0.f;
    if (nbytes) *nbytes = 0;
    if (rc == 1) return DG_ERR_MISSING_WEIGHTS;
    if (rc == 2) return DG_ERR_MALFORMED_WEIGHTS;
    if (name) {
        auto it = t.find(name);
        if (it != t.end()) {
            if (scale) *scale = it->second.scale;
            if (nbytes) *nbytes = it->second.bytes.size();
        }
    }
    return DG_OK;
}

int32_t dg_engine_load_weights_raw(dg_engine* e, const dg_tensor_view* tensors, int32_t count) {
    if (!e || (!tensors && count > 0)) return DG_ERR_INVALID_ARGUMENT;
    dg::TensorMap t;
    for (int i = 0; i < count; i++) {
        if (!tensors[i].name || !tensors[i].dtype || (!tensors[i].data && tensors[i].nbytes)) return fail(e, DG_ERR_INVALID_ARGUMENT, "tensor %d is incomplete", i);
        const std::string dt = tensors[i].dtype;
        if (dt != "f2" && dt != "f4" && dt != "i4" && dt != "i1") return fail(e, DG_ERR_MALFORMED_WEIGHTS, "tensor %s has unknown type %s", tensors[i].name, dt.c_str());
        dg::HostTensor ht;
        ht.dtype = dt;
        ht.bytes.assign(static_cast<const uint8_t*>(tensors[i].data), static_cast<const uint8_t*>(tensors[i].data) + tensors[i].nbytes);
        t[tensors[i].name] = std::move(ht);
    }
    DG_CUDA(e, cudaSetDevice(e->cfg.device));
    return load_net(e, t);
}

int32_t dg_engine_forward_f16(dg_engine* e, const uint16_t* features, int32_t batch, uint16_t* value_out, uint16_t* policy_out) {
    return forward_impl(e, features, kFeatBytes, 1, batch, value_out, policy_out);
}

int32_t dg_engine_forward_packed(dg_engine* e, const dg_packed_position* positions, int32_t batch, uint16_t* value_out, uint16_t* policy_out) {
    return forward_impl(e, positions, sizeof(dg_packed_position), 2, batch, value_out, policy_out);
}

// The raw-position forward of the `batch` positions staged in w.h_in, enqueued on the workspace's stream: H2D, feature
// kernel (+ network) (+ priors), D2H of every result into the workspace's pinned outputs.  No synchronisation.
static int32_t enqueue_raw(dg_engine* e, Workspace& w, int batch, bool network, bool prior, dg_packed_position* planes_out) {
    const size_t in_bytes = sizeof(dg_raw_position) * static_cast<size_t>(batch);
    DG_CUDA(e, cudaMemcpyAsync(w.d_in, w.h_in, in_bytes, cudaMemcpyHostToDevice, w.stream));
    w.resident_kind = 3;
    w.resident_batch = batch;
    w.want_plan = prior;
    int32_t rc = enqueue_network(e, w, batch, -1, network ? 0 : 1);
    w.want_plan = false;
    if (rc) return rc;
    if (prior) {
        DG_CUDA(e, dg::launch_prior_from_policy(w.d_in, w.d_policy, raw_cand(e, w), raw_rep(e, w), raw_prior(e, w), batch, w.stream));
        DG_CUDA(e, cudaMemcpyAsync(w.h_prior, raw_prior(e, w), static_cast<size_t>(batch) * 368 * 4, cudaMemcpyDeviceToHost, w.stream));
    }
    DG_CUDA(e, cudaMemcpyAsync(w.h_legal, raw_legal(e, w), static_cast<size_t>(batch) * 361, cudaMemcpyDeviceToHost, w.stream));
    if (network) {
        DG_CUDA(e, cudaMemcpyAsync(w.h_value, w.d_value, static_cast<size_t>(batch) * 2, cudaMemcpyDeviceToHost, w.stream));
        DG_CUDA(e, cudaMemcpyAsync(w.h_policy, w.d_policy, static_cast<size_t>(batch) * DG_POLICY_SIZE * 2, cudaMemcpyDeviceToHost, w.stream));
    }
    if (planes_out)
        DG_CUDA(e, cudaMemcpyAsync(planes_out, raw_planes(e, w), sizeof(dg_packed_position) * static_cast<size_t>(batch), cudaMemcpyDeviceToHost, w.stream));
    return DG_OK;
}

static int32_t raw_impl(dg_engine* e, const dg_raw_position* positions, int32_t batch, uint16_t* value_out, uint16_t* policy_out,
                        dg_packed_position* planes_out, uint8_t* legal_out, float* prior_out = nullptr) {
    if (!e) return DG_ERR_INVALID_ARGUMENT;
    if (!positions || !legal_out) return fail(e, DG_ERR_INVALID_ARGUMENT, "null buffer");
    if (batch < 1 || batch > e->cfg.max_batch) return fail(e, DG_ERR_INVALID_ARGUMENT, "batch %d outside 1..%d", batch, e->cfg.max_batch);
    const bool network = value_out && policy_out;
    if (network && !e->net.loaded) return fail(e, DG_ERR_MISSING_WEIGHTS, "no weights loaded");
    // the colour to move selects tables on the device: anything but 1 / 2 stops here (every other field is masked or only compared)
    for (int32_t i = 0; i < batch; i++)
        if (positions[i].to_move != 1 && positions[i].to_move != 2) return fail(e, DG_ERR_INVALID_ARGUMENT, "position %d: to_move %d is not a colour", i, positions[i].to_move);
    DG_CUDA(e, cudaSetDevice(e->cfg.device));
    WsGuard guard(e);
    Workspace& w = *guard.w;
    memcpy(w.h_in, positions, sizeof(dg_raw_position) * static_cast<size_t>(batch));
    int32_t rc = enqueue_raw(e, w, batch, network, prior_out != nullptr, planes_out);
    if (rc) { cudaStreamSynchronize(w.stream); return rc; }      // nothing of this call may still run when the workspace goes back
    cudaError_t werr = wait_stream(e, w);
    if (werr != cudaSuccess) return fail(e, DG_ERR_CUDA, "forward failed: %s", cudaGetErrorString(werr));
    memcpy(legal_out, w.h_legal, static_cast<size_t>(batch) * 361);
    if (prior_out) memcpy(prior_out, w.h_prior, static_cast<size_t>(batch) * 368 * 4);
    if (network) {
        memcpy(value_out, w.h_value, static_cast<size_t>(batch) * 2);
        memcpy(policy_out, w.h_policy, static_cast<size_t>(batch) * DG_POLICY_SIZE * 2);
    }
    return DG_OK;
}

int32_t dg_engine_forward_raw(dg_engine* e, const dg_raw_position* positions, int32_t batch, uint16_t* value_out, uint16_t* policy_out,
                              uint8_t* legal_out) {
    if (!value_out || !policy_out) return e ? fail(e, DG_ERR_INVALID_ARGUMENT, "null buffer") : DG_ERR_INVALID_ARGUMENT;
    return raw_impl(e, positions, batch, value_out, policy_out, nullptr, legal_out);
}

int32_t dg_engine_forward_raw_prior(dg_engine* e, const dg_raw_position* positions, int32_t batch, uint16_t* value_out,
                                    uint16_t* policy_out, uint8_t* legal_out, float* prior_out) {
    if (!value_out || !policy_out || !prior_out) return e ? fail(e, DG_ERR_INVALID_ARGUMENT, "null buffer") : DG_ERR_INVALID_ARGUMENT;
    return raw_impl(e, positions, batch, value_out, policy_out, nullptr, legal_out, prior_out);
}

int32_t dg_engine_features_raw(dg_engine* e, const dg_raw_position* positions, int32_t batch, dg_packed_position* planes_out,
                               uint8_t* legal_out) {
    if (!planes_out) return e ? fail(e, DG_ERR_INVALID_ARGUMENT, "null buffer") : DG_ERR_INVALID_ARGUMENT;
    return raw_impl(e, positions, batch, nullptr, nullptr, planes_out, legal_out);
}

int32_t dg_engine_synchronize(dg_engine* e) {
    if (!e) return DG_ERR_INVALID_ARGUMENT;
    DG_CUDA(e, cudaSetDevice(e->cfg.device));
    DG_CUDA(e, cudaDeviceSynchronize());
    return DG_OK;
}

void* dg_engine_alloc_host(dg_engine* e, uint64_t nbytes) {
    if (!e || !nbytes) return nullptr;
    void* p = nullptr;
    cudaSetDevice(e->cfg.device);
    if (cudaHostAlloc(&p, nbytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    std::lock_guard<std::mutex> g(e->err_mutex);
    e->host_allocs.push_back(p);
    return p;
}

void dg_engine_free_host(dg_engine* e, void* ptr) {
    if (!e || !ptr) return;
    {
        std::lock_guard<std::mutex> g(e->err_mutex);
        for (auto& p : e->host_allocs)
            if (p == ptr) { p = e->host_allocs.back(); e->host_allocs.pop_back(); break; }
    }
    cudaFreeHost(ptr);
}

// ---------------------------------------------------------------------------------- leaf-batch queue
//
// Replaces `pool::Batcher` (src/libdg_mcts/pool/batch.rs:61-124: a mutex, a Vec and a 23 KB memcpy per leaf, at most
// `max_batches` batches alive) and the blocking `batch.forward(predictor)` of pool/worker_thread.rs:88-99.  A leaf batch is
// one of the engine's workspaces: producers claim slots of its pinned input array lock-free (384 B per leaf), ONE graph
// launch evaluates them (H2D, feature kernel, tower, heads, priors, D2H -- captured once per batch bucket), and completion
// is a word in pinned host memory written by the last kernel of the launch, so waiting costs no driver call and any
// thread can notice it.  The self-play driver (search_api.cpp) runs on this.

namespace {

constexpr int32_t kSealed = INT32_MIN;

// the opaque dg_leaf_batch of the ABI is a workspace of its engine
inline dg_leaf_batch* LB(Workspace* w) { return reinterpret_cast<dg_leaf_batch*>(w); }
inline Workspace& WS(dg_leaf_batch* b) { return *reinterpret_cast<Workspace*>(b); }
inline const Workspace& WSC(const dg_leaf_batch* b) { return *reinterpret_cast<const Workspace*>(b); }

// The forward of `bucket` staged positions as an executable graph (captured on first use).
int32_t batch_graph(dg_engine* e, Workspace& w, int bucket, bool prior, cudaGraphExec_t* out) {
    const uint32_t key = (static_cast<uint32_t>(bucket) << 1) | (prior ? 1u : 0u);
    auto it = w.graphs.find(key);
    if (it != w.graphs.end()) { *out = it->second; return DG_OK; }
    std::lock_guard<std::mutex> g(e->graph_mutex);
    cudaGraph_t graph = nullptr;
    DG_CUDA(e, cudaStreamBeginCapture(w.stream, cudaStreamCaptureModeThreadLocal));
    w.capturing = true;
    int32_t rc = enqueue_raw(e, w, bucket, true, prior, nullptr);
    if (rc == DG_OK && dg::launch_signal_host(w.h_flag, 1u, w.stream) != cudaSuccess) rc = fail(e, DG_ERR_CUDA, "signal launch failed during capture");
    w.capturing = false;
    cudaError_t cerr = cudaStreamEndCapture(w.stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
    if (cerr != cudaSuccess || !graph) return fail(e, DG_ERR_CUDA, "stream capture failed: %s", cudaGetErrorString(cerr));
    cudaGraphExec_t exec = nullptr;
    cerr = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (cerr != cudaSuccess) return fail(e, DG_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(cerr));
    w.graphs[key] = exec;
    *out = exec;
    return DG_OK;
}

}  // namespace

int32_t dg_engine_batch_acquire(dg_engine* e, dg_leaf_batch** out) {
    if (!e || !out) return DG_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (!e->net.loaded) return fail(e, DG_ERR_MISSING_WEIGHTS, "no weights loaded");
    Workspace* w = acquire(e);
    w->fill.store(0);
    w->committed.store(0);
    w->submitted = 0;
    *w->h_flag = 1;
    *out = LB(w);
    return DG_OK;
}

void dg_engine_batch_release(dg_leaf_batch* b) {
    if (!b) return;
    Workspace& w = WS(b);
    if (*static_cast<volatile uint32_t*>(w.h_flag) == 0) { cudaSetDevice(w.owner->cfg.device); cudaStreamSynchronize(w.stream); *w.h_flag = 1; }
    release(w.owner, &w);
}

int32_t dg_leaf_batch_capacity(const dg_leaf_batch* b) { return b ? WSC(b).owner->cfg.max_batch : 0; }

int32_t dg_leaf_batch_push(dg_leaf_batch* b, const dg_raw_position* positions, int32_t n) {
    if (!b || !positions || n < 1) return DG_ERR_INVALID_ARGUMENT;
    for (int32_t i = 0; i < n; i++)                            // as dg_engine_forward_raw: the colour to move selects tables on the device
        if (positions[i].to_move != 1 && positions[i].to_move != 2) return DG_ERR_INVALID_ARGUMENT;
    Workspace& w = WS(b);
    const int32_t cap = w.owner->cfg.max_batch;
    int32_t at = w.fill.load(std::memory_order_relaxed);
    do {
        if (at < 0 || at + n > cap) return -1;             // sealed (a submit is in flight) or full: submit / wait, then retry
    } while (!w.fill.compare_exchange_weak(at, at + n, std::memory_order_acq_rel, std::memory_order_relaxed));
    memcpy(reinterpret_cast<dg_raw_position*>(w.h_in) + at, positions, sizeof(dg_raw_position) * static_cast<size_t>(n));
    w.committed.fetch_add(n, std::memory_order_release);
    return at;
}

int32_t dg_leaf_batch_submit(dg_leaf_batch* b, uint32_t outputs) {
    if (!b) return DG_ERR_INVALID_ARGUMENT;
    Workspace& w = WS(b);
    dg_engine* e = w.owner;
    if (*static_cast<volatile uint32_t*>(w.h_flag) == 0) return fail(e, DG_ERR_INVALID_ARGUMENT, "the previous submit of this leaf batch has not been waited for");
    const int32_t n = w.fill.exchange(kSealed, std::memory_order_acq_rel);      // later pushes fail until dg_leaf_batch_reset
    if (n <= 0) { w.fill.store(n < 0 ? kSealed : 0); return n == 0 ? fail(e, DG_ERR_INVALID_ARGUMENT, "empty leaf batch") : fail(e, DG_ERR_INVALID_ARGUMENT, "leaf batch already submitted"); }
    while (w.committed.load(std::memory_order_acquire) < n) {}               // producers that claimed a slot finish their 384-byte copy
    DG_CUDA(e, cudaSetDevice(e->cfg.device));
    const bool prior = (outputs & DG_LEAF_PRIOR) != 0;
    const int cap = e->cfg.max_batch;
    int bucket = (n + kBucket - 1) / kBucket * kBucket;
    if (bucket > cap) bucket = cap;
    dg_raw_position* slots = reinterpret_cast<dg_raw_position*>(w.h_in);
    for (int i = n; i < bucket; i++) slots[i] = slots[0];                       // the padding of a bucket is evaluated too: keep it a position
    w.submitted = n;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    *static_cast<volatile uint32_t*>(w.h_flag) = 0;
    int32_t rc;
    const bool eager = (e->cfg.flags & DG_FLAG_NO_GRAPH) != 0 || !e->kernels_configured.load(std::memory_order_acquire);
    if (eager) {
        // the first forward of an engine runs outside a capture: the kernels' one-time attribute calls are not capturable
        std::unique_lock<std::mutex> g(e->graph_mutex, std::defer_lock);
        if (!(e->cfg.flags & DG_FLAG_NO_GRAPH)) g.lock();
        rc = enqueue_raw(e, w, bucket, true, prior, nullptr);
        if (rc == DG_OK && dg::launch_signal_host(w.h_flag, 1u, w.stream) != cudaSuccess) rc = fail(e, DG_ERR_CUDA, "signal launch failed");
        if (rc == DG_OK && g.owns_lock()) {
            if (cudaStreamSynchronize(w.stream) != cudaSuccess) rc = fail(e, DG_ERR_CUDA, "first forward failed");
            else e->kernels_configured.store(true, std::memory_order_release);
        }
    } else {
        cudaGraphExec_t exec = nullptr;
        rc = batch_graph(e, w, bucket, prior, &exec);
        if (rc == DG_OK) {
            cudaError_t cerr = cudaGraphLaunch(exec, w.stream);
            if (cerr != cudaSuccess) rc = fail(e, DG_ERR_CUDA, "cudaGraphLaunch failed: %s", cudaGetErrorString(cerr));
        }
    }
    if (rc) { cudaStreamSynchronize(w.stream); *w.h_flag = 1; w.submitted = 0; }
    return rc;
}

int32_t dg_leaf_batch_ready(dg_leaf_batch* b) {
    if (!b) return DG_ERR_INVALID_ARGUMENT;
    Workspace& w = WS(b);
    const uint32_t f = *static_cast<volatile uint32_t*>(w.h_flag);
    std::atomic_thread_fence(std::memory_order_acquire);
    if (f != 0) return 1;
    // a launch that failed on the device never writes the flag: ask the driver once in a while
    if ((w.polls.fetch_add(1, std::memory_order_relaxed) & 0xffff) == 0xffff) {
        cudaSetDevice(w.owner->cfg.device);
        const cudaError_t q = cudaStreamQuery(w.stream);
        if (q != cudaSuccess && q != cudaErrorNotReady) { *w.h_flag = 1; return fail(w.owner, DG_ERR_CUDA, "leaf batch failed: %s", cudaGetErrorString(q)); }
    }
    return 0;
}

int32_t dg_leaf_batch_wait(dg_leaf_batch* b) {
    if (!b) return DG_ERR_INVALID_ARGUMENT;
    Workspace& w = WS(b);
    dg_engine* e = w.owner;
    const bool nap = (e->cfg.flags & DG_FLAG_BLOCKING_SYNC) != 0;
    for (uint32_t spins = 1;; spins++) {
        if (dg_leaf_batch_ready(b)) return DG_OK;
        if ((spins & 1023) == 0 || nap) {
            if (nap) { timespec ts{0, 20000}; nanosleep(&ts, nullptr); }
            if ((spins & (nap ? 63 : 0xfffff)) == 0) {                          // a failed launch never writes the flag
                cudaSetDevice(e->cfg.device);
                cudaError_t q = cudaStreamQuery(w.stream);
                if (q != cudaSuccess && q != cudaErrorNotReady) { *w.h_flag = 1; return fail(e, DG_ERR_CUDA, "leaf batch failed: %s", cudaGetErrorString(q)); }
            }
        } else {
            __builtin_ia32_pause();
        }
    }
}

int32_t dg_leaf_batch_size(const dg_leaf_batch* b) { return b ? WSC(b).submitted : 0; }
dg_raw_position* dg_leaf_batch_slots(dg_leaf_batch* b) { return b ? reinterpret_cast<dg_raw_position*>(WS(b).h_in) : nullptr; }
const uint16_t* dg_leaf_batch_value(const dg_leaf_batch* b) { return reinterpret_cast<const uint16_t*>(WSC(b).h_value); }
const uint16_t* dg_leaf_batch_policy(const dg_leaf_batch* b) { return reinterpret_cast<const uint16_t*>(WSC(b).h_policy); }
const uint8_t* dg_leaf_batch_legal(const dg_leaf_batch* b) { return WSC(b).h_legal; }
const float* dg_leaf_batch_prior(const dg_leaf_batch* b) { return WSC(b).h_prior; }

void dg_leaf_batch_reset(dg_leaf_batch* b) {
    if (!b) return;
    WS(b).committed.store(0, std::memory_order_relaxed);
    WS(b).fill.store(0, std::memory_order_release);
}

// ---------------------------------------------------------------------------------- measurement

// `callers` host threads each issue blocking dg_engine_forward_f16 calls from their own pinned buffers until `steps`
// calls have been made in total; seconds = wall time from the first call to the last return.
int32_t dg_engine_time_e2e(dg_engine* e, const uint16_t* features, int32_t batch, int32_t steps, int32_t callers, double* seconds) {
    if (!e || !features || !seconds || batch < 1 || steps < 1 || callers < 1 || callers > 16) return DG_ERR_INVALID_ARGUMENT;
    struct Bufs { uint16_t *in, *value, *policy; };
    std::vector<Bufs> bufs(callers);
    const size_t in_bytes = static_cast<size_t>(batch) * kFeatBytes;
    for (auto& b : bufs) {
        b.in = static_cast<uint16_t*>(dg_engine_alloc_host(e, in_bytes));
        b.value = static_cast<uint16_t*>(dg_engine_alloc_host(e, static_cast<size_t>(batch) * 2));
        b.policy = static_cast<uint16_t*>(dg_engine_alloc_host(e, static_cast<size_t>(batch) * DG_POLICY_SIZE * 2));
        if (!b.in || !b.value || !b.policy) return fail(e, DG_ERR_CUDA, "pinned allocation failed");
        memcpy(b.in, features, in_bytes);
    }
    for (auto& b : bufs) { int32_t rc = dg_engine_forward_f16(e, b.in, batch, b.value, b.policy); if (rc) return rc; }
    std::atomic<int32_t> next{0}, failed{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> threads;
    for (int c = 0; c < callers; c++)
        threads.emplace_back([&, c] {
            while (next.fetch_add(1) < steps) {
                int32_t rc = dg_engine_forward_f16(e, bufs[c].in, batch, bufs[c].value, bufs[c].policy);
                if (rc) { failed.store(rc); break; }
            }
        });
    for (auto& t : threads) t.join();
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (auto& b : bufs) { dg_engine_free_host(e, b.in); dg_engine_free_host(e, b.value); dg_engine_free_host(e, b.policy); }
    return failed.load();
}

int32_t dg_engine_time_resident(dg_engine* e, int32_t batch, int32_t iters, int32_t flush_l2, float* ms_total, float* tower_ms,
                                int32_t* launches) {
    if (!e || iters < 1) return DG_ERR_INVALID_ARGUMENT;
    if (!e->net.loaded) return fail(e, DG_ERR_MISSING_WEIGHTS, "no weights loaded");
    DG_CUDA(e, cudaSetDevice(e->cfg.device));
    WsGuard guard(e, batch);
    if (!guard.w) return fail(e, DG_ERR_INVALID_ARGUMENT, "no workspace holds a resident batch of %d positions (run a forward first)", batch);
    Workspace& w = *guard.w;
    int32_t rc;
    constexpr size_t kFlushBytes = 256u << 20;       // > 126 MB of L2
    if (flush_l2 && !e->flush_buf) DG_CUDA(e, cudaMalloc(&e->flush_buf, kFlushBytes));
    // one event pair per step so that the L2 flush between steps stays outside the timed region
    std::vector<cudaEvent_t> ev(2 * static_cast<size_t>(iters));
    for (auto& x : ev) DG_CUDA(e, cudaEventCreate(&x));
    auto timed_pass = [&](int stage, float* total) -> int32_t {
        for (int i = 0; i < iters; i++) {
            if (flush_l2) DG_CUDA(e, cudaMemsetAsync(e->flush_buf, i & 0xff, kFlushBytes, w.stream));
            DG_CUDA(e, cudaEventRecord(ev[2 * i], w.stream));
            if ((rc = enqueue_network(e, w, batch, -1, stage))) return rc;
            DG_CUDA(e, cudaEventRecord(ev[2 * i + 1], w.stream));
        }
        DG_CUDA(e, cudaStreamSynchronize(w.stream));
        double sum = 0.0;
        for (int i = 0; i < iters; i++) {
            float ms = 0.f;
            DG_CUDA(e, cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]));
            sum += ms;
        }
        *total = static_cast<float>(sum);
        return DG_OK;
    };
    float ms = 0.f;
    rc = timed_pass(0, &ms);
    if (rc == DG_OK && ms_total) *ms_total = ms;
    if (launches) *launches = (e->cfg.flags & (DG_FLAG_DEBUG_DIRECT_CONV | DG_FLAG_LAYERWISE)) ? 5 + 2 * e->net.num_blocks
                            : (e->cfg.flags & DG_FLAG_SEPARATE_HEAD_CONV) ? 5 : 4;   // pack, tower (+ head conv), policy FC, finish
    if (rc == DG_OK && tower_ms) {
        rc = timed_pass(2, tower_ms);
        // leave the workspace holding a complete forward again
        if (rc == DG_OK) rc = enqueue_network(e, w, batch, -1, 0);
        cudaStreamSynchronize(w.stream);
    }
    for (auto& x : ev) cudaEventDestroy(x);
    return rc;
}

int32_t dg_engine_debug_conv_trace(dg_engine* e, int32_t batch, int64_t* out, int32_t out_len) {
    if (!e || !out) return DG_ERR_INVALID_ARGUMENT;
    if (!e->net.loaded || e->net.num_blocks < 1) return fail(e, DG_ERR_MISSING_WEIGHTS, "no weights loaded");
    DG_CUDA(e, cudaSetDevice(e->cfg.device));
    WsGuard guard(e, batch);
    if (!guard.w) return fail(e, DG_ERR_INVALID_ARGUMENT, "no workspace holds a resident batch of %d positions", batch);
    Workspace& w = *guard.w;
    const size_t n = static_cast<size_t>(e->num_sms) * 3 * 64;
    long long* buf = nullptr;
    DG_CUDA(e, cudaMalloc(&buf, n * 8));
    DG_CUDA(e, cudaMemset(buf, 0, n * 8));
    e->trace_buf = buf;
    const float g = e->net.gate[0];
    int32_t rc = run_conv(e, w, ConvTcShape::kTower, w.tm_y, w.y, kChan, e->net.c2[0], 128, w.x, kChan, w.x, g, 1.0f - g, batch);
    e->trace_buf = nullptr;
    cudaStreamSynchronize(w.stream);
    if (rc == DG_OK) cudaMemcpy(out, buf, (n < static_cast<size_t>(out_len) ? n : static_cast<size_t>(out_len)) * 8, cudaMemcpyDeviceToHost);
    cudaFree(buf);
    return rc;
}

int32_t dg_engine_debug_tower_trace(dg_engine* e, int32_t batch, int64_t* out, int32_t out_len) {
    if (!e || !out) return DG_ERR_INVALID_ARGUMENT;
    if (!e->net.loaded) return fail(e, DG_ERR_MISSING_WEIGHTS, "no weights loaded");
    DG_CUDA(e, cudaSetDevice(e->cfg.device));
    WsGuard guard(e, batch);
    if (!guard.w) return fail(e, DG_ERR_INVALID_ARGUMENT, "no workspace holds a resident batch of %d positions", batch);
    Workspace& w = *guard.w;
    const size_t n = static_cast<size_t>(e->num_sms) * 3 * 1024;
    long long* buf = nullptr;
    DG_CUDA(e, cudaMalloc(&buf, n * 8));
    DG_CUDA(e, cudaMemset(buf, 0, n * 8));
    e->trace_buf = buf;
    int32_t rc = enqueue_network(e, w, batch, e->net.num_blocks, 3);
    e->trace_buf = nullptr;
    cudaStreamSynchronize(w.stream);
    if (rc == DG_OK) cudaMemcpy(out, buf, (n < static_cast<size_t>(out_len) ? n : static_cast<size_t>(out_len)) * 8, cudaMemcpyDeviceToHost);
    cudaFree(buf);
    return rc;
}

int32_t dg_engine_debug_read_tower(dg_engine* e, int32_t layer, int32_t batch, uint16_t* out) {
    if (!e || !out) return DG_ERR_INVALID_ARGUMENT;
    if (!e->net.loaded) return fail(e, DG_ERR_MISSING_WEIGHTS, "no weights loaded");
    DG_CUDA(e, cudaSetDevice(e->cfg.device));
    WsGuard guard(e, batch);
    if (!guard.w) return fail(e, DG_ERR_INVALID_ARGUMENT, "no workspace holds a resident batch of %d positions", batch);
    Workspace& w = *guard.w;
    const int blocks = (layer < 0) ? e->net.num_blocks : layer;
    int32_t rc = enqueue_network(e, w, batch, blocks, 0);
    if (rc) return rc;
    const size_t rows = static_cast<size_t>(batch) * DG_POS_ROWS;
    std::vector<uint16_t> host(rows * kChan);
    DG_CUDA(e, cudaMemcpyAsync(host.data(), w.x + static_cast<size_t>(DG_GUARD_ROWS) * kChan, rows * kChan * 2, cudaMemcpyDeviceToHost, w.stream));
    DG_CUDA(e, cudaStreamSynchronize(w.stream));
    for (int n = 0; n < batch; n++)
        for (int y = 0; y < 19; y++)
            memcpy(out + (static_cast<size_t>(n) * 361 + y * 19) * kChan,
                   host.data() + (static_cast<size_t>(n) * DG_POS_ROWS + y * DG_LINE_STRIDE) * kChan, 19 * kChan * 2);
    return DG_OK;
}

}  // extern "C"

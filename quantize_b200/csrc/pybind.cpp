// `quant_engine` — the Python-facing torch extension of the B200 engine.
//
// Exports the reference extension's 8 names with the same positional signatures
// (reference engine/kernels/pybind.cpp:9-16, tpack/tpack.h:17-32, functions/funcs.h:17-151) so that
// `from quant_engine import *` in the reference's engine/__init__.py:1-5 — and therefore modelzoo/modules
// (quantconv2d.py:16, operator/quantconv2dop.py:10-13) — work unchanged.  This file is a thin shim: argument
// checks with the reference's error messages, ATen allocation on the current stream, a cache of
// derived operands, and calls into the C-ABI of include/qb200.h.  No arithmetic lives here and there is no
// CPU implementation: without a CUDA device every op raises.
//
// Extension over the reference signature (keyword arguments, all optional):
//   quantconv2d_float_input(..., stride, padding, input_scale=None, input_zero=None, input_qmin=None,
//                           input_qmax=None, residual=None, fuse_relu=False)
// With the activation quantizer's parameters (Quantizer.scale/zero/qmin/qmax, modelzoo/modules/quantizer.py:119-123)
// the op runs the fused activation-quantize + int8 tensor-core path; without them it computes the reference's
// weight-only fp32 semantic.  residual / fuse_relu fuse `relu(out + residual)` into the epilogue (bit-identical to the
// separate torch ops) for callers that own the surrounding block.
//
// Threading: Python objects are converted while the GIL is held; everything after that — cache lookups,
// allocation, the C-ABI launches — runs with the GIL RELEASED.  g_mu guards only the derived-operand caches (short
// critical sections; no code holding it ever takes the GIL).  Launches go to the current stream of the input's device.
#include <pybind11/pybind11.h>
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <ATen/cuda/CUDAEvent.h>
#include <c10/cuda/CUDAGuard.h>

#include <atomic>
#include <list>
#include <memory>
#include <mutex>
#include <unordered_map>

#include "qb200.h"

namespace py = pybind11;

namespace {

#define CHECK_NBITS(b) TORCH_CHECK(b > 0 && b <= 8, #b " must be in the range (0, 8]")
#define CHECK_CUDA(x) TORCH_CHECK(x.device().is_cuda(), #x " must be a CUDA tensor")
#define CHECK_CONTIGUOUS(x) TORCH_CHECK(x.is_contiguous(), #x " must be contiguous")
#define CHECK_INPUT(x) CHECK_CUDA(x); CHECK_CONTIGUOUS(x)
#define CHECK_FLOAT(x) TORCH_CHECK(x.dtype() == torch::kFloat32, #x " must be a float tensor")

void check_rc(int rc, const char* what) {
    if (rc == 0) return;
    const char* msg = qb200_last_error();
    TORCH_CHECK(false, what, " failed (", rc, "): ", msg ? msg : "");
}

void require_cuda_runtime() {
    TORCH_CHECK(at::cuda::is_available(),
                "quant_engine (B200): no CUDA device is available and this engine has no CPU path");
}

int dtype_code(const at::Tensor& x) {
    switch (x.scalar_type()) {
        case at::kByte: return QB200_U8;
        case at::kChar: return QB200_I8;
        case at::kShort: return QB200_I16;
        case at::kInt: return QB200_I32;
        case at::kLong: return QB200_I64;
        case at::kHalf: return QB200_F16;
        case at::kBFloat16: return QB200_BF16;
        case at::kFloat: return QB200_F32;
        case at::kDouble: return QB200_F64;
        default: TORCH_CHECK(false, "tpack: unsupported dtype ", x.scalar_type());
    }
    return -1;
}

void* cur_stream() { return static_cast<void*>(at::cuda::getCurrentCUDAStream().stream()); }

// ------------------------------------------------------------------------------------------------
// tpack / tunpack
// ------------------------------------------------------------------------------------------------
std::vector<at::Tensor> tpack(at::Tensor x, int n_bits, bool sign) {
    CHECK_NBITS(n_bits);
    CHECK_CONTIGUOUS(x);
    require_cuda_runtime();
    const auto home = x.device();
    // The reference also takes CPU tensors (tpack.cu:140-190, one .item() per element).  This engine has no CPU
    // arithmetic: a CPU tensor is staged through the current CUDA device so checkpoint code keeps working.
    at::Tensor xd = home.is_cuda() ? x : x.to(at::Device(at::kCUDA, at::cuda::current_device()));
    c10::cuda::CUDAGuard guard(xd.device());
    const int64_t n = xd.numel();
    const int64_t n_out = (n * n_bits + 7) / 8;
    auto bytes = at::empty({n_out}, xd.options().dtype(at::kByte));
    auto flag = at::zeros({1}, xd.options().dtype(at::kInt));
    check_rc(qb200_tpack(xd.data_ptr(), dtype_code(xd), n, n_bits, sign ? 1 : 0, bytes.data_ptr<uint8_t>(),
                         flag.data_ptr<int32_t>(), cur_stream()),
             "tpack");
    TORCH_CHECK(flag.item<int>() == 0, "The input tensor is out of range.");  // tpack.cu:14, :211-215
    std::vector<int32_t> d;
    d.push_back(n_bits);
    d.push_back(sign ? 1 : 0);
    for (auto s : x.sizes()) d.push_back((int32_t)s);
    auto des = at::tensor(d, at::TensorOptions().dtype(at::kInt)).to(home);
    return {home.is_cuda() ? bytes : bytes.to(home), des};
}

at::Tensor tunpack(at::Tensor x, at::Tensor des) {
    TORCH_CHECK(des.dim() >= 1 && des.size(0) >= 3, "The description is too short, which should be at least 3.");
    auto dh = des.to(at::kCPU, at::kLong).contiguous();
    const int64_t* dp = dh.data_ptr<int64_t>();
    const int n_bits = (int)dp[0];
    CHECK_NBITS(n_bits);
    CHECK_CONTIGUOUS(x);
    TORCH_CHECK(x.dtype() == torch::kByte, "The input tensor must be torch.uint8.");
    require_cuda_runtime();
    const int sign = dp[1] != 0;
    std::vector<int64_t> shape;
    int64_t n = 1;
    for (int64_t i = 2; i < dh.numel(); ++i) {
        shape.push_back(dp[i]);
        n *= dp[i];
    }
    TORCH_CHECK(x.numel() >= (n * n_bits + 7) / 8, "The packed tensor is shorter than its description.");
    const auto home = x.device();
    at::Tensor xd = home.is_cuda() ? x : x.to(at::Device(at::kCUDA, at::cuda::current_device()));
    c10::cuda::CUDAGuard guard(xd.device());
    auto out = at::empty({n}, xd.options().dtype(sign ? at::kChar : at::kByte));
    check_rc(qb200_tunpack(xd.data_ptr<uint8_t>(), n, n_bits, sign, out.data_ptr(), cur_stream()), "tunpack");
    out = out.reshape(shape);
    return home.is_cuda() ? out : out.to(home);
}

// ------------------------------------------------------------------------------------------------
// derived-operand caches (host copies of descriptors, prepared weights, float copies of qmin/qmax)
// ------------------------------------------------------------------------------------------------
struct Key {
    const void* ptr;
    uint64_t version;
    const void* aux;
    uint64_t aux2;
    bool operator==(const Key& o) const { return ptr == o.ptr && version == o.version && aux == o.aux && aux2 == o.aux2; }
};
struct KeyHash {
    size_t operator()(const Key& k) const {
        return std::hash<const void*>()(k.ptr) ^ (std::hash<uint64_t>()(k.version) * 1000003u) ^
               (std::hash<const void*>()(k.aux) << 1) ^ (std::hash<uint64_t>()(k.aux2) * 0x9E3779B97F4A7C15ull);
    }
};

// An entry is valid only while the storage it was derived from is alive (a freed block can be handed out again
// at the same address), so every entry keeps a weak reference to that storage.  Values are shared_ptrs: a caller keeps
// its value alive after the lock is dropped even if another thread evicts the entry.
template <typename V>
class DerivedCache {
  public:
    explicit DerivedCache(size_t cap) : cap_(cap) {}
    std::shared_ptr<V> find(const Key& k, const at::Tensor& src) {
        auto it = map_.find(k);
        if (it == map_.end()) return nullptr;
        auto live = it->second->weak.lock();
        if (!live || live.get() != src.storage().unsafeGetStorageImpl()) {
            order_.erase(it->second);
            map_.erase(it);
            return nullptr;
        }
        order_.splice(order_.begin(), order_, it->second);
        return it->second->value;
    }
    std::shared_ptr<V> insert(const Key& k, const at::Tensor& src, V v) {
        auto old = map_.find(k);
        if (old != map_.end()) {
            order_.erase(old->second);
            map_.erase(old);
        }
        order_.push_front(Entry{k, c10::weak_intrusive_ptr<c10::StorageImpl>(src.storage().getWeakStorageImpl()),
                                std::make_shared<V>(std::move(v))});
        map_[k] = order_.begin();
        while (order_.size() > cap_) {
            map_.erase(order_.back().key);
            order_.pop_back();
        }
        return order_.front().value;
    }

  private:
    struct Entry {
        Key key;
        c10::weak_intrusive_ptr<c10::StorageImpl> weak;
        std::shared_ptr<V> value;
    };
    size_t cap_;
    std::list<Entry> order_;
    std::unordered_map<Key, typename std::list<Entry>::iterator, KeyHash> map_;
};

struct PreparedWeights {
    at::Tensor buffer;
    bool zero_is_zero;
    // the buffer is written on the stream of the first call; other streams wait on this event before reading it
    std::shared_ptr<at::cuda::CUDAEvent> ready;
    cudaStream_t stream;
    std::atomic<bool> settled{false};
    PreparedWeights() = default;
    PreparedWeights(PreparedWeights&& o) noexcept
        : buffer(std::move(o.buffer)), zero_is_zero(o.zero_is_zero), ready(std::move(o.ready)), stream(o.stream),
          settled(o.settled.load()) {}
};

// The caches hold CUDA tensors and events.  They are intentionally never destroyed: static destructors run after the CUDA
// runtime has begun to shut down, and destroying an event of a second device there crashed the interpreter at exit.
std::mutex& g_mu = *new std::mutex;   // guards the four caches below, nothing else
auto& g_des_cache = *new DerivedCache<std::vector<int64_t>>(4096);
auto& g_prep_cache = *new DerivedCache<PreparedWeights>(2048);
auto& g_float_cache = *new DerivedCache<at::Tensor>(4096);
auto& g_range_cache = *new DerivedCache<std::pair<float, float>>(4096);

// Tensor::_version() throws for inference-mode tensors (torch.inference_mode()): they cannot be modified in place
// outside inference mode, so the address alone identifies them.
uint64_t version_of(const at::Tensor& t) { return t.is_inference() ? 0ull : (uint64_t)t._version(); }
Key key_of(const at::Tensor& t, const void* aux = nullptr, uint64_t aux2 = 0) { return Key{t.data_ptr(), version_of(t), aux, aux2}; }

std::shared_ptr<std::vector<int64_t>> host_des(const at::Tensor& des) {
    const Key k = key_of(des);
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if (auto v = g_des_cache.find(k, des)) return v;
    }
    auto dh = des.to(at::kCPU, at::kLong).contiguous();  // one blocking copy per descriptor (the reference: 6 per call)
    std::vector<int64_t> v(dh.data_ptr<int64_t>(), dh.data_ptr<int64_t>() + dh.numel());
    std::lock_guard<std::mutex> lock(g_mu);
    return g_des_cache.insert(k, des, std::move(v));
}

// A quantizer parameter as Python hands it over (tensor with one element, or a number), converted under the GIL.
struct QParam {
    bool none = true, is_tensor = false;
    at::Tensor t;
    double val = 0.0;
};
QParam qparam(const py::object& o, const char* name) {
    QParam q;
    if (o.is_none()) return q;
    q.none = false;
    if (THPVariable_Check(o.ptr())) {
        q.is_tensor = true;
        q.t = THPVariable_Unpack(o.ptr());
        TORCH_CHECK(q.t.numel() == 1, name, " must have exactly one element (per-tensor activation quantization)");
    } else {
        q.val = o.cast<double>();
    }
    return q;
}

// scale / zero / qmin / qmax as a device float pointer without per-call work (no GIL needed)
const float* device_float(const QParam& q, const at::Device& dev, std::vector<at::Tensor>& keep) {
    if (q.is_tensor) {
        const at::Tensor& t = q.t;
        if (t.device() == dev && t.scalar_type() == at::kFloat && t.is_contiguous()) {
            keep.push_back(t);
            return t.data_ptr<float>();
        }
        const Key k = key_of(t, reinterpret_cast<const void*>((intptr_t)dev.index() + 1));
        {
            std::lock_guard<std::mutex> lock(g_mu);
            if (auto v = g_float_cache.find(k, t)) { keep.push_back(*v); return v->data_ptr<float>(); }
        }
        auto f = t.detach().to(dev, at::kFloat).contiguous();
        std::lock_guard<std::mutex> lock(g_mu);
        auto v = g_float_cache.insert(k, t, f);
        keep.push_back(*v);
        return v->data_ptr<float>();
    }
    // python number: one cached device scalar per (value, device)
    static auto& scalars = *new std::unordered_map<int64_t, at::Tensor>;   // (never destroyed, see the caches above)
    float fv = (float)q.val;
    int32_t bits;
    memcpy(&bits, &fv, 4);
    const int64_t sk = ((int64_t)dev.index() << 32) | (uint32_t)bits;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        auto it = scalars.find(sk);
        if (it != scalars.end()) return it->second.data_ptr<float>();
    }
    auto fresh = at::full({1}, q.val, at::TensorOptions().dtype(at::kFloat).device(dev));
    std::lock_guard<std::mutex> lock(g_mu);
    return scalars.emplace(sk, fresh).first->second.data_ptr<float>();
}

// The integer path stores activations as unsigned bytes and feeds the MMA an unsigned A operand: it needs
// 0 <= qmin <= qmax <= 255.  Other ranges (signed symmetric activation quantizers: qmin = -2^(n-1), minmax.py:124-127; an
// uncalibrated Quantizer's default buffers) take the fake-quantize + fp32 path.  Tensors are read once per (ptr, version).
bool activation_range_fits_u8(const QParam& qmin, const QParam& qmax) {
    auto value = [](const QParam& q) -> float {
        if (!q.is_tensor) return (float)q.val;
        return q.t.detach().to(at::kCPU, at::kFloat).item<float>();
    };
    float lo, hi;
    if (qmin.is_tensor || qmax.is_tensor) {
        const at::Tensor& anchor = qmin.is_tensor ? qmin.t : qmax.t;
        const Key k = key_of(anchor, qmax.is_tensor ? qmax.t.data_ptr() : nullptr, qmax.is_tensor ? version_of(qmax.t) : 0);
        std::shared_ptr<std::pair<float, float>> v;
        {
            std::lock_guard<std::mutex> lock(g_mu);
            v = g_range_cache.find(k, anchor);
        }
        if (!v) {
            std::pair<float, float> r(value(qmin), value(qmax));   // one sync per quantizer
            std::lock_guard<std::mutex> lock(g_mu);
            v = g_range_cache.insert(k, anchor, r);
        }
        lo = v->first;
        hi = v->second;
    } else {
        lo = (float)qmin.val;
        hi = (float)qmax.val;
    }
    return lo >= 0.f && hi <= 255.f && lo <= hi;
}

// prepared weights: derived once per (weight storage, version, descriptor, weight_zero storage + version, groups)
std::shared_ptr<PreparedWeights> prepared_weights(const qb200_conv_shape& s, const at::Tensor& weight, const at::Tensor& weight_des,
                                                  const at::Tensor& weight_zero, void* st) {
    const uint64_t aux2 = (uint64_t)reinterpret_cast<uintptr_t>(weight_zero.data_ptr()) * 31u + version_of(weight_zero) * 1000003u +
                          (uint64_t)(s.C / (s.Cg > 0 ? s.Cg : 1));
    const Key wk = key_of(weight, weight_des.data_ptr(), aux2);
    std::shared_ptr<PreparedWeights> pw;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        pw = g_prep_cache.find(wk, weight);
    }
    if (!pw) {
        PreparedWeights fresh;
        fresh.buffer = at::empty({(int64_t)qb200_conv_prepared_bytes(&s)}, weight.options());
        check_rc(qb200_conv_prepare_weights(&s, weight.data_ptr<uint8_t>(), fresh.buffer.data_ptr(), st), "prepare_weights");
        fresh.ready = std::make_shared<at::cuda::CUDAEvent>();
        fresh.ready->record(at::cuda::getCurrentCUDAStream());
        fresh.stream = static_cast<cudaStream_t>(st);
        fresh.zero_is_zero = weight_zero.abs().max().item<float>() == 0.f;  // one sync per weight tensor
        std::lock_guard<std::mutex> lock(g_mu);
        pw = g_prep_cache.insert(wk, weight, std::move(fresh));
    }
    // first use on another stream: order it after the preparation (once the event has completed no wait is needed any
    // more — this also keeps replays / stream captures free of cross-stream waits)
    if (pw->stream != static_cast<cudaStream_t>(st) && !pw->settled.load(std::memory_order_acquire)) {
        // cudaEventQuery / waits on outside events are illegal while a stream captures: a capture is preceded by a warm-up
        // and a device synchronisation (host.GraphedForward, torch.cuda.graph's own contract), so the weights are ready
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(static_cast<cudaStream_t>(st), &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) return pw;
        if (pw->ready->query()) pw->settled.store(true, std::memory_order_release);
        else pw->ready->block(at::cuda::getCurrentCUDAStream());
    }
    return pw;
}

// Quantizer.simulate on the device for the fp32 fall-backs (any range): one engine kernel
at::Tensor fake_quantize_dev(const at::Tensor& input, const qb200_act_quant& aq, void* st) {
    auto out = at::empty_like(input);
    check_rc(qb200_fake_quantize_f32(input.data_ptr<float>(), input.numel(), &aq, out.data_ptr<float>(), st), "fake_quantize");
    return out;
}

// ------------------------------------------------------------------------------------------------
// quantconv2d_float_input
// ------------------------------------------------------------------------------------------------
struct QuantArgs {
    QParam scale, zero, qmin, qmax;
};

// (GIL released, device guard set) d = [n_bits, sign, K, Cg, R, S]; input is 4-D
at::Tensor conv_core(const at::Tensor& input, const at::Tensor& weight, const at::Tensor& weight_des, const std::vector<int64_t>& d,
                     const at::Tensor& weight_scale, const at::Tensor& weight_zero, const c10::optional<at::Tensor>& bias,
                     const int stride, const int padding, const QuantArgs& qa, const c10::optional<at::Tensor>& residual,
                     const bool fuse_relu) {
    qb200_conv_shape s;
    s.N = (int32_t)input.size(0);
    s.C = (int32_t)input.size(1);
    s.H = (int32_t)input.size(2);
    s.W = (int32_t)input.size(3);
    s.w_bits = (int32_t)d[0];
    s.w_sign = d[1] != 0;
    s.K = (int32_t)d[2];
    s.Cg = (int32_t)d[3];
    s.R = (int32_t)d[4];
    s.S = (int32_t)d[5];
    s.stride = stride;
    s.pad = padding;
    int32_t P = 0, Q = 0;
    check_rc(qb200_conv_out_hw(&s, &P, &Q), "quantconv2d_float_input");
    TORCH_CHECK(weight.numel() >= qb200_packed_bytes((int64_t)s.K * s.Cg * s.R * s.S, s.w_bits),
                "weight is shorter than weight_des describes");
    const int64_t n_ws = weight_scale.numel();
    TORCH_CHECK(n_ws == 1 || n_ws == s.K, "weight_scale must have 1 or ", s.K, " elements");
    TORCH_CHECK(weight_zero.numel() == n_ws, "weight_zero must have as many elements as weight_scale");
    if (bias.has_value()) TORCH_CHECK(bias.value().numel() == s.K, "bias must have ", s.K, " elements");

    auto out = at::empty({s.N, s.K, P, Q}, input.options());
    const float* bias_p = bias.has_value() ? bias.value().data_ptr<float>() : nullptr;
    void* st = cur_stream();

    const bool fused = !qa.scale.none;
    TORCH_CHECK(fused || (!residual.has_value() && !fuse_relu),
                "residual / fuse_relu need the activation quantizer parameters (the fused path)");
    if (residual.has_value()) {
        const at::Tensor& r = residual.value();
        CHECK_INPUT(r);
        TORCH_CHECK(r.dtype() == torch::kFloat32 && r.sizes() == out.sizes(), "residual must be a float tensor shaped like the output");
    }
    if (!fused) {
        // the reference's semantic: no activation quantization, fp32 accumulate in the reference's order
        check_rc(qb200_quantconv2d_weightonly(&s, input.data_ptr<float>(), weight.data_ptr<uint8_t>(),
                                              weight_scale.data_ptr<float>(), weight_zero.data_ptr<float>(), (int32_t)n_ws,
                                              bias_p, out.data_ptr<float>(), st),
                 "quantconv2d_float_input");
        return out;
    }
    TORCH_CHECK(!qa.zero.none && !qa.qmin.none && !qa.qmax.none,
                "input_scale, input_zero, input_qmin and input_qmax must be given together");

    auto pw = prepared_weights(s, weight, weight_des, weight_zero, st);

    std::vector<at::Tensor> keep;
    qb200_act_quant aq;
    aq.scale = device_float(qa.scale, input.device(), keep);
    aq.zero = device_float(qa.zero, input.device(), keep);
    aq.qmin = device_float(qa.qmin, input.device(), keep);
    aq.qmax = device_float(qa.qmax, input.device(), keep);

    if (!pw->zero_is_zero || !activation_range_fits_u8(qa.qmin, qa.qmax)) {
        // asymmetric weights do not factor into an integer GEMM with a float zero point, and activation ranges outside
        // [0, 255] do not fit the unsigned byte operand: fake-quantize the activations on the device
        // (quantizer.py:215-218, one engine kernel) and run the fp32 weight-only kernel
        TORCH_CHECK(s.C == s.Cg, "asymmetric weights / signed activation ranges with groups > 1 are not supported");
        auto xdq = fake_quantize_dev(input, aq, st);
        check_rc(qb200_quantconv2d_weightonly(&s, xdq.data_ptr<float>(), weight.data_ptr<uint8_t>(),
                                              weight_scale.data_ptr<float>(), weight_zero.data_ptr<float>(), (int32_t)n_ws,
                                              bias_p, out.data_ptr<float>(), st),
                 "quantconv2d_float_input");
        // the optional tail, as the separate torch ops
        if (residual.has_value()) out.add_(residual.value());
        if (fuse_relu) out.relu_();
        return out;
    }

    auto ws = at::empty({(int64_t)qb200_conv_workspace_bytes(&s)}, weight.options());
    qb200_conv_tail tail;
    tail.residual = residual.has_value() ? residual.value().data_ptr<float>() : nullptr;
    tail.relu = fuse_relu ? 1 : 0;
    tail.next_shape = nullptr;
    tail.next_quant = nullptr;
    tail.next_workspace = nullptr;
    check_rc(qb200_quantconv2d_fused_ex(&s, input.data_ptr<float>(), pw->buffer.data_ptr(), weight_scale.data_ptr<float>(),
                                        (int32_t)n_ws, bias_p, &aq, &tail, ws.data_ptr(), out.data_ptr(), QB200_OUT_F32, st),
             "quantconv2d_float_input");
    return out;
}

at::Tensor quantconv2d_float_input(const at::Tensor& input, const at::Tensor& weight, const at::Tensor& weight_des,
                                   const at::Tensor& weight_scale, const at::Tensor& weight_zero,
                                   const c10::optional<at::Tensor>& bias, const int stride, const int padding,
                                   const py::object& input_scale, const py::object& input_zero,
                                   const py::object& input_qmin, const py::object& input_qmax,
                                   const c10::optional<at::Tensor>& residual, const bool fuse_relu) {
    // same checks, same messages as quantconv2d_float_input.cu:151-159
    CHECK_INPUT(input);
    CHECK_FLOAT(input);
    CHECK_INPUT(weight);
    CHECK_INPUT(weight_des);
    CHECK_INPUT(weight_scale);
    CHECK_INPUT(weight_zero);
    if (bias.has_value()) { CHECK_INPUT(bias.value()); }
    TORCH_CHECK(input.dim() == 4, "input must be a 4D tensor");
    TORCH_CHECK(weight.dtype() == torch::kByte, "weight must be a packed uint8 tensor");
    CHECK_FLOAT(weight_scale);
    CHECK_FLOAT(weight_zero);
    if (bias.has_value()) { TORCH_CHECK(bias.value().dtype() == torch::kFloat32, "bias must be a float tensor"); }
    QuantArgs qa{qparam(input_scale, "input_scale"), qparam(input_zero, "input_zero"), qparam(input_qmin, "input_qmin"),
                 qparam(input_qmax, "input_qmax")};

    py::gil_scoped_release nogil;
    c10::cuda::CUDAGuard guard(input.device());
    auto d = host_des(weight_des);
    TORCH_CHECK(d->size() >= 6, "weight_des must hold [n_bits, sign, K, C, R, S]");
    return conv_core(input, weight, weight_des, *d, weight_scale, weight_zero, bias, stride, padding, qa, residual, fuse_relu);
}

// ------------------------------------------------------------------------------------------------
// quantconv2d_chain (engine extension, SURVEY §8(f) next-1: int8 activations between layers)
//
// Runs consecutive fused convs  x -> L0 -> [relu] -> L1 -> [relu] -> ... -> L(n-1) -> (+ residual) -> [relu]  where each
// intermediate result is handed to the next layer already quantized with that layer's activation quantizer (written
// by the producer's epilogue into the consumer's workspace) instead of as an fp32 tensor.  The result is bit-identical
// to calling quantconv2d_float_input layer by layer: the same fp32 value goes through the same quantizer arithmetic.
// Each element of `layers` is a tuple
//   (weight, weight_des, weight_scale, weight_zero, bias|None, stride, padding,
//    input_scale, input_zero, input_qmin, input_qmax, relu_after)
// Pairs the engine cannot chain (qb200_conv_handoff_supported) fall back to an fp32 intermediate.
// emit_next = such a tuple for the layer that will consume this chain's RESULT: the last epilogue then writes the fp32
// result (the residual path needs it) and the consumer's quantized workspace; the call returns (out, workspace|None)
// and a later chain starting with that layer takes the workspace as input_handoff (its `input` is then only used for
// the shape).

// ------------------------------------------------------------------------------------------------
struct ChainLayer {
    at::Tensor weight, des, scale, zero;
    c10::optional<at::Tensor> bias;
    int stride = 1, pad = 0;
    QuantArgs qa;
    qb200_conv_shape s;
    int32_t P, Q;
    qb200_act_quant aq;
    std::shared_ptr<PreparedWeights> pw;
    bool relu;
};

py::object quantconv2d_chain(const at::Tensor& input, const py::list& layers, const c10::optional<at::Tensor>& residual,
                             const c10::optional<at::Tensor>& input_handoff, const py::object& emit_next, const bool emit_only) {
    CHECK_INPUT(input);
    CHECK_FLOAT(input);
    TORCH_CHECK(input.dim() == 4, "input must be a 4D tensor");
    const size_t n = layers.size();  // layers of the chain proper; an optional extra entry describes emit_next's consumer
    TORCH_CHECK(n >= 1, "quantconv2d_chain: no layers");
    const bool emit = !emit_next.is_none();
    std::vector<ChainLayer> L(n + (emit ? 1 : 0));
    // ---- Python objects -> C++ (GIL held) ----
    for (size_t i = 0; i < L.size(); ++i) {
        const py::tuple t = (i < n ? py::object(layers[i]) : emit_next).cast<py::tuple>();
        TORCH_CHECK(t.size() == 12, "quantconv2d_chain: each layer is a 12-tuple");
        ChainLayer& l = L[i];
        l.weight = t[0].cast<at::Tensor>();
        l.des = t[1].cast<at::Tensor>();
        l.scale = t[2].cast<at::Tensor>();
        l.zero = t[3].cast<at::Tensor>();
        if (!t[4].is_none()) l.bias = t[4].cast<at::Tensor>();
        l.stride = t[5].cast<int>();
        l.pad = t[6].cast<int>();
        l.qa = QuantArgs{qparam(t[7], "input_scale"), qparam(t[8], "input_zero"), qparam(t[9], "input_qmin"), qparam(t[10], "input_qmax")};
        TORCH_CHECK(!l.qa.scale.none && !l.qa.zero.none && !l.qa.qmin.none && !l.qa.qmax.none,
                    "quantconv2d_chain: every layer needs its activation quantizer parameters");
        l.relu = t[11].cast<bool>();
    }
    at::Tensor out, ws_emit;
    {
        py::gil_scoped_release nogil;
        c10::cuda::CUDAGuard guard(input.device());
        void* st = cur_stream();
        std::vector<at::Tensor> keep;
        int32_t N = (int32_t)input.size(0), C = (int32_t)input.size(1), H = (int32_t)input.size(2), W = (int32_t)input.size(3);
        for (size_t i = 0; i < L.size(); ++i) {
            ChainLayer& l = L[i];
            CHECK_INPUT(l.weight);
            CHECK_INPUT(l.des);
            CHECK_INPUT(l.scale);
            CHECK_INPUT(l.zero);
            TORCH_CHECK(l.weight.dtype() == torch::kByte, "weight must be a packed uint8 tensor");
            CHECK_FLOAT(l.scale);
            CHECK_FLOAT(l.zero);
            if (l.bias.has_value()) {
                CHECK_INPUT(l.bias.value());
                TORCH_CHECK(l.bias.value().dtype() == torch::kFloat32, "bias must be a float tensor");
            }
            auto dp = host_des(l.des);
            const std::vector<int64_t>& d = *dp;
            TORCH_CHECK(d.size() >= 6, "weight_des must hold [n_bits, sign, K, C, R, S]");
            qb200_conv_shape& s = l.s;
            s.N = N; s.C = C; s.H = H; s.W = W;
            s.w_bits = (int32_t)d[0];
            s.w_sign = d[1] != 0;
            s.K = (int32_t)d[2]; s.Cg = (int32_t)d[3]; s.R = (int32_t)d[4]; s.S = (int32_t)d[5];
            s.stride = l.stride;
            s.pad = l.pad;
            check_rc(qb200_conv_out_hw(&s, &l.P, &l.Q), "quantconv2d_chain");
            TORCH_CHECK(l.weight.numel() >= qb200_packed_bytes((int64_t)s.K * s.Cg * s.R * s.S, s.w_bits),
                        "weight is shorter than weight_des describes");
            const int64_t n_ws = l.scale.numel();
            TORCH_CHECK(n_ws == 1 || n_ws == s.K, "weight_scale must have 1 or ", s.K, " elements");
            TORCH_CHECK(l.zero.numel() == n_ws, "weight_zero must have as many elements as weight_scale");
            if (l.bias.has_value()) TORCH_CHECK(l.bias.value().numel() == s.K, "bias must have ", s.K, " elements");
            l.pw = prepared_weights(s, l.weight, l.des, l.zero, st);
            TORCH_CHECK(l.pw->zero_is_zero, "quantconv2d_chain needs symmetric weights (weight_zero == 0)");
            TORCH_CHECK(activation_range_fits_u8(l.qa.qmin, l.qa.qmax),
                        "quantconv2d_chain needs activation ranges inside [0, 255] (unsigned byte hand-off)");
            l.aq.scale = device_float(l.qa.scale, input.device(), keep);
            l.aq.zero = device_float(l.qa.zero, input.device(), keep);
            l.aq.qmin = device_float(l.qa.qmin, input.device(), keep);
            l.aq.qmax = device_float(l.qa.qmax, input.device(), keep);
            C = s.K; H = l.P; W = l.Q;
        }
        if (residual.has_value()) {
            const at::Tensor& r = residual.value();
            CHECK_INPUT(r);
            const ChainLayer& e = L[n - 1];
            TORCH_CHECK(r.dtype() == torch::kFloat32 && r.dim() == 4 && r.size(0) == N && r.size(1) == e.s.K && r.size(2) == e.P &&
                            r.size(3) == e.Q, "residual must be a float tensor shaped like the output");
        }
        at::Tensor x = input;          // fp32 input of the current layer (when not handed off)
        at::Tensor ws_in;              // quantized workspace of the current layer (when handed off)
        bool handed = false;
        if (input_handoff.has_value()) {
            // the first layer's input arrives already quantized (written by an earlier chain's emit_next for this very layer)
            const at::Tensor& h = input_handoff.value();
            CHECK_INPUT(h);
            TORCH_CHECK(h.dtype() == torch::kByte && (size_t)h.numel() >= qb200_conv_workspace_bytes(&L[0].s),
                        "input_handoff is not a workspace of the first layer");
            ws_in = h;
            handed = true;
        }
        for (size_t i = 0; i < n; ++i) {
            ChainLayer& l = L[i];
            const bool last = i + 1 == n;
            const float* bias_p = l.bias.has_value() ? l.bias.value().data_ptr<float>() : nullptr;
            qb200_conv_tail tail;
            tail.residual = (last && residual.has_value()) ? residual.value().data_ptr<float>() : nullptr;
            tail.relu = l.relu ? 1 : 0;
            tail.next_shape = nullptr;
            tail.next_quant = nullptr;
            tail.next_workspace = nullptr;
            at::Tensor ws_next;
            // the last layer hands over only on request (emit_next) and then writes BOTH the fp32 result and the bytes
            const bool hand = (!last || emit) && qb200_conv_handoff_supported(&l.s, &L[i + 1].s) != 0;
            if (hand) {
                ws_next = at::empty({(int64_t)qb200_conv_workspace_bytes(&L[i + 1].s)}, l.weight.options());
                tail.next_shape = &L[i + 1].s;
                tail.next_quant = &L[i + 1].aq;
                tail.next_workspace = ws_next.data_ptr();
            }
            if (!hand || last) out = at::empty({l.s.N, l.s.K, l.P, l.Q}, input.options());
            // emit_only: the caller needs the last layer's result ONLY as the consumer's bytes (the next block reads nothing
            // else): no fp32 store — `out` is then returned uninitialised, for its shape
            void* out_p = ((hand && !last) || (hand && last && emit_only)) ? nullptr : out.data_ptr();
            if (last && hand) ws_emit = ws_next;
            if (handed) {
                check_rc(qb200_conv_from_workspace_ex(&l.s, ws_in.data_ptr(), l.pw->buffer.data_ptr(), l.scale.data_ptr<float>(),
                                                      (int32_t)l.scale.numel(), bias_p, &l.aq, &tail, out_p, QB200_OUT_F32, st),
                         "quantconv2d_chain");
            } else {
                auto ws = at::empty({(int64_t)qb200_conv_workspace_bytes(&l.s)}, l.weight.options());
                check_rc(qb200_quantconv2d_fused_ex(&l.s, x.data_ptr<float>(), l.pw->buffer.data_ptr(), l.scale.data_ptr<float>(),
                                                    (int32_t)l.scale.numel(), bias_p, &l.aq, &tail, ws.data_ptr(), out_p,
                                                    QB200_OUT_F32, st),
                         "quantconv2d_chain");
            }
            handed = hand;
            if (hand) ws_in = ws_next;
            else x = out;
        }
    }
    if (!emit) return py::cast(out);
    return py::make_tuple(out, ws_emit.defined() ? py::cast(ws_emit) : py::none());
}

// quantconv2d_u8_nhwc (engine extension): a fused quantized conv whose input is ALREADY quantized — the NHWC(Cp) byte
// workspace another layer's epilogue wrote with this very activation quantizer (quantconv2d_chain's emit_next).  Used for
// the 1x1 / stride-2 shortcut conv of a down-sampling residual block: its quantizer sees the same tensor as the block's
// conv1 and, calibrated on the same data, has bit-identical parameters, so conv1's hand-off bytes serve both and the
// separate quantizer pass over the fp32 tensor disappears.  `layer` is a chain_args 12-tuple; in_shape = (N, C, H, W).
at::Tensor quantconv2d_u8_nhwc(const at::Tensor& q_nhwc, const std::vector<int64_t>& in_shape, const py::tuple& layer) {
    CHECK_INPUT(q_nhwc);
    TORCH_CHECK(q_nhwc.dtype() == torch::kByte, "q_nhwc must be a uint8 tensor");
    TORCH_CHECK(in_shape.size() == 4, "in_shape must be (N, C, H, W)");
    TORCH_CHECK(layer.size() == 12, "layer is a 12-tuple (chain_args)");
    ChainLayer l;
    l.weight = layer[0].cast<at::Tensor>();
    l.des = layer[1].cast<at::Tensor>();
    l.scale = layer[2].cast<at::Tensor>();
    l.zero = layer[3].cast<at::Tensor>();
    if (!layer[4].is_none()) l.bias = layer[4].cast<at::Tensor>();
    l.stride = layer[5].cast<int>();
    l.pad = layer[6].cast<int>();
    l.qa = QuantArgs{qparam(layer[7], "input_scale"), qparam(layer[8], "input_zero"), qparam(layer[9], "input_qmin"), qparam(layer[10], "input_qmax")};
    TORCH_CHECK(!l.qa.scale.none && !l.qa.zero.none && !l.qa.qmin.none && !l.qa.qmax.none,
                "quantconv2d_u8_nhwc: the activation quantizer parameters are needed (scale and zero point of the bytes)");
    TORCH_CHECK(!layer[11].cast<bool>(), "quantconv2d_u8_nhwc: no fused ReLU");
    CHECK_INPUT(l.weight);
    CHECK_INPUT(l.des);
    CHECK_INPUT(l.scale);
    CHECK_INPUT(l.zero);
    TORCH_CHECK(l.weight.dtype() == torch::kByte, "weight must be a packed uint8 tensor");
    CHECK_FLOAT(l.scale);
    CHECK_FLOAT(l.zero);
    if (l.bias.has_value()) {
        CHECK_INPUT(l.bias.value());
        TORCH_CHECK(l.bias.value().dtype() == torch::kFloat32, "bias must be a float tensor");
    }
    py::gil_scoped_release nogil;
    c10::cuda::CUDAGuard guard(q_nhwc.device());
    void* st = cur_stream();
    auto dp = host_des(l.des);
    const std::vector<int64_t>& d = *dp;
    TORCH_CHECK(d.size() >= 6, "weight_des must hold [n_bits, sign, K, C, R, S]");
    qb200_conv_shape& s = l.s;
    s.N = (int32_t)in_shape[0]; s.C = (int32_t)in_shape[1]; s.H = (int32_t)in_shape[2]; s.W = (int32_t)in_shape[3];
    s.w_bits = (int32_t)d[0];
    s.w_sign = d[1] != 0;
    s.K = (int32_t)d[2]; s.Cg = (int32_t)d[3]; s.R = (int32_t)d[4]; s.S = (int32_t)d[5];
    s.stride = l.stride;
    s.pad = l.pad;
    TORCH_CHECK(s.C == s.Cg, "quantconv2d_u8_nhwc: groups must be 1");
    check_rc(qb200_conv_out_hw(&s, &l.P, &l.Q), "quantconv2d_u8_nhwc");
    TORCH_CHECK(q_nhwc.numel() >= (int64_t)s.N * s.H * s.W * qb200_padded_channels(s.C),
                "q_nhwc is smaller than N*H*W*Cp bytes");
    TORCH_CHECK(l.weight.numel() >= qb200_packed_bytes((int64_t)s.K * s.Cg * s.R * s.S, s.w_bits),
                "weight is shorter than weight_des describes");
    const int64_t n_ws = l.scale.numel();
    TORCH_CHECK(n_ws == 1 || n_ws == s.K, "weight_scale must have 1 or ", s.K, " elements");
    TORCH_CHECK(l.zero.numel() == n_ws, "weight_zero must have as many elements as weight_scale");
    if (l.bias.has_value()) TORCH_CHECK(l.bias.value().numel() == s.K, "bias must have ", s.K, " elements");
    l.pw = prepared_weights(s, l.weight, l.des, l.zero, st);
    TORCH_CHECK(l.pw->zero_is_zero, "quantconv2d_u8_nhwc needs symmetric weights (weight_zero == 0)");
    std::vector<at::Tensor> keep;
    l.aq.scale = device_float(l.qa.scale, q_nhwc.device(), keep);
    l.aq.zero = device_float(l.qa.zero, q_nhwc.device(), keep);
    l.aq.qmin = device_float(l.qa.qmin, q_nhwc.device(), keep);
    l.aq.qmax = device_float(l.qa.qmax, q_nhwc.device(), keep);
    auto out = at::empty({s.N, s.K, l.P, l.Q}, l.scale.options());
    const float* bias_p = l.bias.has_value() ? l.bias.value().data_ptr<float>() : nullptr;
    check_rc(qb200_conv2d_q8_nhwc(&s, q_nhwc.data_ptr<uint8_t>(), l.pw->buffer.data_ptr(), l.scale.data_ptr<float>(), (int32_t)n_ws,
                                  bias_p, &l.aq, out.data_ptr(), QB200_OUT_F32, st),
             "quantconv2d_u8_nhwc");
    return out;
}


// fake_quantize: Quantizer.simulate of a per-tensor quantizer as one kernel
at::Tensor fake_quantize(const at::Tensor& input, const py::object& scale, const py::object& zero, const py::object& qmin,
                         const py::object& qmax) {
    CHECK_INPUT(input);
    CHECK_FLOAT(input);
    QuantArgs qa{qparam(scale, "scale"), qparam(zero, "zero"), qparam(qmin, "qmin"), qparam(qmax, "qmax")};
    TORCH_CHECK(!qa.scale.none && !qa.zero.none && !qa.qmin.none && !qa.qmax.none, "fake_quantize: scale, zero, qmin and qmax are required");
    py::gil_scoped_release nogil;
    c10::cuda::CUDAGuard guard(input.device());
    std::vector<at::Tensor> keep;
    qb200_act_quant aq;
    aq.scale = device_float(qa.scale, input.device(), keep);
    aq.zero = device_float(qa.zero, input.device(), keep);
    aq.qmin = device_float(qa.qmin, input.device(), keep);
    aq.qmax = device_float(qa.qmax, input.device(), keep);
    return fake_quantize_dev(input, aq, cur_stream());
}

// max_pool2d (engine helper for the packed ResNet forward; same result as torch.nn.functional.max_pool2d)
at::Tensor max_pool2d(const at::Tensor& input, int kernel, int stride, int padding, const c10::optional<at::Tensor>& out_opt) {
    CHECK_INPUT(input);
    CHECK_FLOAT(input);
    TORCH_CHECK(input.dim() == 4, "input must be a 4D tensor");
    py::gil_scoped_release nogil;
    c10::cuda::CUDAGuard guard(input.device());
    const int H = (int)input.size(2), W = (int)input.size(3);
    TORCH_CHECK(kernel >= 1 && stride >= 1 && padding >= 0 && 2 * padding <= kernel, "max_pool2d: bad geometry");
    const int P = (H + 2 * padding - kernel) / stride + 1, Q = (W + 2 * padding - kernel) / stride + 1;
    TORCH_CHECK(P >= 1 && Q >= 1, "max_pool2d: empty output");
    at::Tensor out;
    if (out_opt.has_value()) {   // e.g. a batch slice of a larger tensor (the forward in chunks of images)
        out = out_opt.value();
        CHECK_INPUT(out);
        TORCH_CHECK(out.dtype() == torch::kFloat32 && out.dim() == 4 && out.size(0) == input.size(0) && out.size(1) == input.size(1) &&
                        out.size(2) == P && out.size(3) == Q && out.device() == input.device(),
                    "max_pool2d: out must be a contiguous float tensor of the output's shape on the input's device");
    } else {
        out = at::empty({input.size(0), input.size(1), P, Q}, input.options());
    }
    check_rc(qb200_maxpool2d_f32(input.data_ptr<float>(), input.size(0) * input.size(1), H, W, kernel, stride, padding,
                                 out.data_ptr<float>(), cur_stream()),
             "max_pool2d");
    return out;
}

// avg_pool_global (engine helper: AdaptiveAvgPool2d((1, 1)) of the packed ResNet forward) -> [N, C, 1, 1]
at::Tensor avg_pool_global(const at::Tensor& input) {
    CHECK_INPUT(input);
    CHECK_FLOAT(input);
    TORCH_CHECK(input.dim() == 4, "input must be a 4D tensor");
    TORCH_CHECK(input.size(2) * input.size(3) >= 1, "avg_pool_global: empty planes");
    py::gil_scoped_release nogil;
    c10::cuda::CUDAGuard guard(input.device());
    auto out = at::empty({input.size(0), input.size(1), 1, 1}, input.options());
    check_rc(qb200_avgpool_global_f32(input.data_ptr<float>(), input.size(0) * input.size(1), (int)(input.size(2) * input.size(3)),
                                      out.data_ptr<float>(), cur_stream()),
             "avg_pool_global");
    return out;
}

// ------------------------------------------------------------------------------------------------
// Calibration reductions (reference range/minmax.py:62-108, :44-60, :184-203).
//   minmax(x, granularity, flag, symmetric, update_mode=0, momentum=0.0, run_min=None, run_max=None) -> (xmin, xmax)
//     granularity 0: per tensor (0-d results); 1: per channel — weights: dim 0; activations (flag 1): dim 1
//     update_mode 1 / 2: run_min / run_max (float tensors shaped like the result) are updated in place (MinMax / MAMinMax)
//   kthvalue(x, k, granularity, flag, use_abs) -> values      (torch.kthvalue(...)[0] of the estimator's flattened view)
// ------------------------------------------------------------------------------------------------
struct RedView {
    int64_t A, R, B;
    std::vector<int64_t> out_shape;
};
RedView reduction_view(const at::Tensor& x, int granularity, int flag) {
    RedView v;
    TORCH_CHECK(x.numel() > 0, "cannot reduce an empty tensor");
    if (granularity == 0) {
        v.A = 1; v.R = 1; v.B = x.numel();
    } else if (flag == 1) {   // activation: (N, C, ...) -> rows = C
        TORCH_CHECK(x.dim() >= 2, "per-channel activation ranges need at least 2 dimensions");
        v.A = x.size(0); v.R = x.size(1); v.B = x.numel() / (x.size(0) * x.size(1));
        v.out_shape = {x.size(1)};
    } else {                  // weight: (C, ...) -> rows = dim 0
        TORCH_CHECK(x.dim() >= 1, "per-channel weight ranges need at least 1 dimension");
        v.A = 1; v.R = x.size(0); v.B = x.numel() / x.size(0);
        v.out_shape = {x.size(0)};
    }
    return v;
}

std::vector<at::Tensor> minmax(const at::Tensor& input, int granularity, int flag, bool symmetric, int update_mode, double momentum,
                               const c10::optional<at::Tensor>& run_min, const c10::optional<at::Tensor>& run_max) {
    CHECK_INPUT(input);
    CHECK_FLOAT(input);
    py::gil_scoped_release nogil;
    c10::cuda::CUDAGuard guard(input.device());
    const RedView v = reduction_view(input, granularity, flag);
    auto lo = at::empty(v.out_shape, input.options()), hi = at::empty(v.out_shape, input.options());
    float* rmin = nullptr;
    float* rmax = nullptr;
    if (update_mode != 0) {
        TORCH_CHECK(run_min.has_value() && run_max.has_value(), "minmax: update_mode needs run_min / run_max");
        for (const at::Tensor* t : {&run_min.value(), &run_max.value()}) {
            CHECK_INPUT((*t));
            TORCH_CHECK(t->dtype() == torch::kFloat32 && t->numel() == v.R && t->device() == input.device(),
                        "minmax: run_min / run_max must be float tensors with one element per row on the input's device");
        }
        rmin = run_min.value().data_ptr<float>();
        rmax = run_max.value().data_ptr<float>();
    }
    auto ws = at::empty({(int64_t)qb200_minmax_workspace_bytes(v.R)}, input.options().dtype(at::kByte));
    // torch evaluates `momentum * x + (1 - momentum) * old` with the python scalars cast to fp32 (1 - momentum in double first)
    check_rc(qb200_minmax_f32(input.data_ptr<float>(), v.A, v.R, v.B, symmetric ? 1 : 0, lo.data_ptr<float>(), hi.data_ptr<float>(),
                              update_mode, (float)momentum, (float)(1.0 - momentum), rmin, rmax, ws.data_ptr(), cur_stream()),
             "minmax");
    return {lo, hi};
}

at::Tensor kthvalue(const at::Tensor& input, int64_t k, int granularity, int flag, bool use_abs) {
    CHECK_INPUT(input);
    CHECK_FLOAT(input);
    py::gil_scoped_release nogil;
    c10::cuda::CUDAGuard guard(input.device());
    const RedView v = reduction_view(input, granularity, flag);
    TORCH_CHECK(k >= 1 && k <= v.A * v.B, "kthvalue(): selected number k out of range for dimension");
    auto out = at::empty(v.out_shape, input.options());
    auto ws = at::empty({(int64_t)qb200_kthvalue_workspace_bytes(v.R)}, input.options().dtype(at::kByte));
    check_rc(qb200_kthvalue_f32(input.data_ptr<float>(), v.A, v.R, v.B, use_abs ? 1 : 0, k, out.data_ptr<float>(), ws.data_ptr(),
                                cur_stream()),
             "kthvalue");
    return out;
}

// ------------------------------------------------------------------------------------------------
// off-path ops: exported so that `from quant_engine import *` binds all 8 names; the reference's plain
// float `linear` / `conv2d` (linear.cu, conv2d.cu) are not quantized operators and stay outside this engine.
// ------------------------------------------------------------------------------------------------
[[noreturn]] void off_path(const char* name) {
    TORCH_CHECK(false, "quant_engine.", name,
                " is outside the quantized-operator path this engine implements (tpack, tunpack, quantconv2d, "
                "quantconv2d_float_input, quantlinear, quantlinear_float_input)");
    abort();
}
at::Tensor linear(const at::Tensor&, const at::Tensor&, const c10::optional<at::Tensor>&, int) { off_path("linear"); }
at::Tensor conv2d(const at::Tensor&, const at::Tensor&, const c10::optional<at::Tensor>&, int, int, int) { off_path("conv2d"); }

// quantlinear_float_input (reference quantlinear_float_input.cu:120-182, funcs.h).  The six positional
// arguments are the reference's: weight-only fp32 semantic in the reference kernel's accumulation order.  With the
// activation quantizer's parameters (input_scale / zero / qmin / qmax, as for the conv op) the layer runs as the 1x1 case
// of the fused integer convolution: M = batch (x tokens) rows through the tensor-core kernel.
at::Tensor quantlinear_float_input(const at::Tensor& input, const at::Tensor& weight, const at::Tensor& weight_des,
                                   const at::Tensor& weight_scale, const at::Tensor& weight_zero,
                                   const c10::optional<at::Tensor>& bias, const py::object& input_scale,
                                   const py::object& input_zero, const py::object& input_qmin, const py::object& input_qmax) {
    CHECK_INPUT(input);
    CHECK_FLOAT(input);
    CHECK_INPUT(weight);
    CHECK_INPUT(weight_des);
    CHECK_INPUT(weight_scale);
    CHECK_INPUT(weight_zero);
    if (bias.has_value()) { CHECK_INPUT(bias.value()); }
    TORCH_CHECK(input.dim() == 2, "input must be a 2D tensor (batch_size, input_size)");
    TORCH_CHECK(weight.dtype() == torch::kByte, "weight must be a packed uint8 tensor");
    CHECK_FLOAT(weight_scale);
    CHECK_FLOAT(weight_zero);
    if (bias.has_value()) { TORCH_CHECK(bias.value().dtype() == torch::kFloat32, "bias must be a float tensor"); }
    QuantArgs qa{qparam(input_scale, "input_scale"), qparam(input_zero, "input_zero"), qparam(input_qmin, "input_qmin"),
                 qparam(input_qmax, "input_qmax")};
    py::gil_scoped_release nogil;
    c10::cuda::CUDAGuard guard(input.device());
    auto dp = host_des(weight_des);
    const std::vector<int64_t>& d = *dp;
    TORCH_CHECK(d.size() >= 4, "weight_des must hold [n_bits, sign, out_features, in_features]");
    const int64_t B = input.size(0), in_f = input.size(1), out_f = d[2];
    TORCH_CHECK(d[3] == in_f, "input has ", in_f, " features, the weight expects ", d[3]);
    TORCH_CHECK(weight.numel() >= qb200_packed_bytes(out_f * in_f, (int)d[0]), "weight is shorter than weight_des describes");
    const int64_t n_ws = weight_scale.numel();
    TORCH_CHECK(n_ws == 1 || n_ws == out_f, "weight_scale must have 1 or ", out_f, " elements");
    TORCH_CHECK(weight_zero.numel() == n_ws, "weight_zero must have as many elements as weight_scale");
    if (bias.has_value()) TORCH_CHECK(bias.value().numel() == out_f, "bias must have ", out_f, " elements");
    if (qa.scale.none) {
        auto out = at::empty({B, out_f}, input.options());
        check_rc(qb200_quantlinear_weightonly(input.data_ptr<float>(), B, (int32_t)in_f, (int32_t)out_f, weight.data_ptr<uint8_t>(),
                                              (int32_t)d[0], d[1] != 0, weight_scale.data_ptr<float>(),
                                              weight_zero.data_ptr<float>(), (int32_t)n_ws,
                                              bias.has_value() ? bias.value().data_ptr<float>() : nullptr,
                                              out.data_ptr<float>(), cur_stream()),
                 "quantlinear_float_input");
        return out;
    }
    const std::vector<int64_t> d6 = {d[0], d[1], out_f, in_f, 1, 1};
    auto out4 = conv_core(input.view({B, in_f, 1, 1}), weight, weight_des, d6, weight_scale, weight_zero, bias, 1, 0, qa,
                          c10::nullopt, false);
    return out4.view({B, out_f});
}

// quantconv2d (reference quantconv2d.cu:164-264, funcs.h:113-124, dispatched by
// quantconv2dop.py:88-91 when input and weight are both uint8 tpack streams).  Eleven positional arguments, as the
// reference's.  input_des = [n_bits, sign, N, C, H, W]; input_scale / input_zero: one element (per tensor) or C elements
// (per input channel), dequantized as (q - zero) * scale like the weights.
at::Tensor quantconv2d(const at::Tensor& input, const at::Tensor& input_des, const at::Tensor& input_scale,
                       const at::Tensor& input_zero, const at::Tensor& weight, const at::Tensor& weight_des,
                       const at::Tensor& weight_scale, const at::Tensor& weight_zero, const c10::optional<at::Tensor>& bias,
                       const int stride, const int padding) {
    CHECK_INPUT(input);          // quantconv2d.cu:178-189
    CHECK_INPUT(input_des);
    CHECK_INPUT(input_scale);
    CHECK_INPUT(input_zero);
    CHECK_INPUT(weight);
    CHECK_INPUT(weight_des);
    CHECK_INPUT(weight_scale);
    CHECK_INPUT(weight_zero);
    if (bias.has_value()) { CHECK_INPUT(bias.value()); }
    TORCH_CHECK(input.dtype() == torch::kByte && weight.dtype() == torch::kByte, "input and weight must be packed uint8 tensors");
    CHECK_FLOAT(input_scale);
    CHECK_FLOAT(input_zero);
    CHECK_FLOAT(weight_scale);
    CHECK_FLOAT(weight_zero);
    if (bias.has_value()) { TORCH_CHECK(bias.value().dtype() == torch::kFloat32, "bias must be a float tensor"); }
    py::gil_scoped_release nogil;
    c10::cuda::CUDAGuard guard(input.device());
    auto idp = host_des(input_des);
    auto wdp = host_des(weight_des);
    const std::vector<int64_t>&id = *idp, &wd = *wdp;
    TORCH_CHECK(id.size() >= 6, "input_des must hold [n_bits, sign, N, C, H, W]");
    TORCH_CHECK(wd.size() >= 6, "weight_des must hold [n_bits, sign, K, C, R, S]");
    const int in_bits = (int)id[0], in_sign = id[1] != 0;
    CHECK_NBITS(in_bits);
    qb200_conv_shape s;
    s.N = (int32_t)id[2]; s.C = (int32_t)id[3]; s.H = (int32_t)id[4]; s.W = (int32_t)id[5];
    s.w_bits = (int32_t)wd[0];
    s.w_sign = wd[1] != 0;
    s.K = (int32_t)wd[2]; s.Cg = (int32_t)wd[3]; s.R = (int32_t)wd[4]; s.S = (int32_t)wd[5];
    s.stride = stride;
    s.pad = padding;
    TORCH_CHECK(s.Cg == s.C, "quantconv2d: the weight has ", s.Cg, " input channels, the input ", s.C,
                " (the reference op has no groups, quantconv2d.cu:100-124)");
    int32_t P = 0, Q = 0;
    check_rc(qb200_conv_out_hw(&s, &P, &Q), "quantconv2d");
    const int64_t n_in = (int64_t)s.N * s.C * s.H * s.W;
    TORCH_CHECK(input.numel() >= qb200_packed_bytes(n_in, in_bits), "input is shorter than input_des describes");
    TORCH_CHECK(weight.numel() >= qb200_packed_bytes((int64_t)s.K * s.Cg * s.R * s.S, s.w_bits),
                "weight is shorter than weight_des describes");
    const int64_t n_is = input_scale.numel(), n_ws = weight_scale.numel();
    TORCH_CHECK(n_is == 1 || n_is == s.C, "input_scale must have 1 or ", s.C, " elements");
    TORCH_CHECK(input_zero.numel() == n_is, "input_zero must have as many elements as input_scale");
    TORCH_CHECK(n_ws == 1 || n_ws == s.K, "weight_scale must have 1 or ", s.K, " elements");
    TORCH_CHECK(weight_zero.numel() == n_ws, "weight_zero must have as many elements as weight_scale");
    if (bias.has_value()) TORCH_CHECK(bias.value().numel() == s.K, "bias must have ", s.K, " elements");
    const float* bias_p = bias.has_value() ? bias.value().data_ptr<float>() : nullptr;
    void* st = cur_stream();
    auto out = at::empty({s.N, s.K, P, Q}, input_scale.options());
    if (n_in == 0) return out;
    auto pw = prepared_weights(s, weight, weight_des, weight_zero, st);
    if (n_is == 1 && pw->zero_is_zero) {
        // per-tensor input quantizer, symmetric weights: integer GEMM on the tensor cores
        auto ws = at::empty({(int64_t)qb200_quantconv2d_packed_workspace_bytes(&s)}, input.options());
        check_rc(qb200_quantconv2d_packed(&s, input.data_ptr<uint8_t>(), in_bits, in_sign, input_scale.data_ptr<float>(),
                                          input_zero.data_ptr<float>(), pw->buffer.data_ptr(), weight_scale.data_ptr<float>(),
                                          (int32_t)n_ws, bias_p, ws.data_ptr(), out.data_ptr(), QB200_OUT_F32, st),
                 "quantconv2d");
        return out;
    }
    // per-input-channel input scales / asymmetric weights: the reference's fp32 arithmetic, in its order
    auto xf = at::empty({s.N, s.C, s.H, s.W}, input_scale.options());
    check_rc(qb200_dequant_packed_f32(input.data_ptr<uint8_t>(), in_bits, in_sign, n_in, (int64_t)s.H * s.W, s.C,
                                      input_scale.data_ptr<float>(), input_zero.data_ptr<float>(), (int32_t)n_is, 0,
                                      xf.data_ptr<float>(), st),
             "quantconv2d");
    check_rc(qb200_quantconv2d_weightonly(&s, xf.data_ptr<float>(), weight.data_ptr<uint8_t>(), weight_scale.data_ptr<float>(),
                                          weight_zero.data_ptr<float>(), (int32_t)n_ws, bias_p, out.data_ptr<float>(), st),
             "quantconv2d");
    return out;
}

// quantlinear (reference quantlinear.cu:231-297, funcs.h:37-46, dispatched by quantlinearop.py:71-74).
// Nine positional arguments, as the reference's.  input_des = [n_bits, sign, batch, in_features]; input_scale / input_zero:
// 0-d (expanded over the batch, quantlinear.cu:275-281) or one element per row; weight_scale / weight_zero: 0-d or one per
// output feature.  (q + zero) convention on both operands.
at::Tensor quantlinear(const at::Tensor& input, const at::Tensor& input_des, const at::Tensor& input_scale,
                       const at::Tensor& input_zero, const at::Tensor& weight, const at::Tensor& weight_des,
                       const at::Tensor& weight_scale, const at::Tensor& weight_zero, const c10::optional<at::Tensor>& bias) {
    CHECK_INPUT(input);          // quantlinear.cu:243-250
    CHECK_INPUT(weight);
    CHECK_INPUT(input_des);
    CHECK_INPUT(input_scale);
    CHECK_INPUT(input_zero);
    CHECK_INPUT(weight_des);
    CHECK_INPUT(weight_scale);
    CHECK_INPUT(weight_zero);
    TORCH_CHECK(input.dtype() == torch::kByte && weight.dtype() == torch::kByte, "input and weight must be packed uint8 tensors");
    CHECK_FLOAT(input_scale);
    CHECK_FLOAT(input_zero);
    CHECK_FLOAT(weight_scale);
    CHECK_FLOAT(weight_zero);
    py::gil_scoped_release nogil;
    c10::cuda::CUDAGuard guard(input.device());
    auto idp = host_des(input_des);
    auto wdp = host_des(weight_des);
    const std::vector<int64_t>&id = *idp, &wd = *wdp;
    TORCH_CHECK(id.size() >= 4, "input_des must hold [n_bits, sign, batch, in_features]");
    TORCH_CHECK(wd.size() >= 4, "weight_des must hold [n_bits, sign, out_features, in_features]");
    TORCH_CHECK(id[3] == wd[3], "Input and weight do not match");                       // quantlinear.cu:259
    const int in_bits = (int)id[0], w_bits = (int)wd[0];
    CHECK_NBITS(in_bits);
    CHECK_NBITS(w_bits);
    const int64_t B = id[2], in_f = id[3], out_f = wd[2];
    at::Tensor bias_f;
    if (bias.has_value()) {
        CHECK_INPUT(bias.value());
        bias_f = bias.value().to(at::kFloat);                                              // :265
        TORCH_CHECK(bias_f.numel() == out_f, "Weight and bias do not match");             // :266
    }
    TORCH_CHECK(input.numel() >= qb200_packed_bytes(B * in_f, in_bits), "input is shorter than input_des describes");
    TORCH_CHECK(weight.numel() >= qb200_packed_bytes(out_f * in_f, w_bits), "weight is shorter than weight_des describes");
    auto expand = [](const at::Tensor& t, int64_t n, const char* name) {
        if (t.numel() == 1 && n != 1) return t.reshape({1}).expand({n}).contiguous();     // :275-289 (0-d -> one per row)
        TORCH_CHECK(t.numel() == n, name, " must have 1 or ", n, " elements");
        return t;
    };
    const at::Tensor is = expand(input_scale, B, "input_scale"), iz = expand(input_zero, B, "input_zero");
    const at::Tensor wsc = expand(weight_scale, out_f, "weight_scale"), wz = expand(weight_zero, out_f, "weight_zero");
    auto out = at::empty({B, out_f}, input_scale.options());
    check_rc(qb200_quantlinear_packed(input.data_ptr<uint8_t>(), in_bits, id[1] != 0, is.data_ptr<float>(), iz.data_ptr<float>(), B,
                                      (int32_t)in_f, (int32_t)out_f, weight.data_ptr<uint8_t>(), w_bits, wd[1] != 0,
                                      wsc.data_ptr<float>(), wz.data_ptr<float>(),
                                      bias_f.defined() ? bias_f.data_ptr<float>() : nullptr, out.data_ptr<float>(), cur_stream()),
             "quantlinear");
    return out;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "B200 (sm_100a) implementation of JingInAI/Quantize's quant_engine hot path";
    m.def("tpack", &tpack, "Packs the given tensor into a vector of tensors.", py::arg("x"), py::arg("n_bits"), py::arg("sign"),
          py::call_guard<py::gil_scoped_release>());
    m.def("tunpack", &tunpack, "Unpacks the given vector of tensors into a tensor.", py::arg("x"), py::arg("des"),
          py::call_guard<py::gil_scoped_release>());
    m.def("linear", &linear, "Linear function.", py::arg("input"), py::arg("weight"), py::arg("bias") = py::none(),
          py::arg("mode") = 0);
    m.def("quantlinear", &quantlinear, "Quantized linear function.", py::arg("input"), py::arg("input_des"), py::arg("input_scale"),
          py::arg("input_zero"), py::arg("weight"), py::arg("weight_des"), py::arg("weight_scale"), py::arg("weight_zero"),
          py::arg("bias") = py::none());
    m.def("quantlinear_float_input", &quantlinear_float_input, "Quantized linear function with float input.",
          py::arg("input"), py::arg("weight"), py::arg("weight_des"), py::arg("weight_scale"), py::arg("weight_zero"),
          py::arg("bias") = py::none(), py::arg("input_scale") = py::none(), py::arg("input_zero") = py::none(),
          py::arg("input_qmin") = py::none(), py::arg("input_qmax") = py::none());
    m.def("conv2d", &conv2d, "Conv2d function.", py::arg("input"), py::arg("weight"), py::arg("bias"), py::arg("stride"),
          py::arg("padding"), py::arg("mode") = 0);
    m.def("quantconv2d", &quantconv2d, "Quantized conv2d function.", py::arg("input"), py::arg("input_des"), py::arg("input_scale"),
          py::arg("input_zero"), py::arg("weight"), py::arg("weight_des"), py::arg("weight_scale"), py::arg("weight_zero"),
          py::arg("bias"), py::arg("stride"), py::arg("padding"));
    m.def("quantconv2d_float_input", &quantconv2d_float_input, "Quantized conv2d function with float input.",
          py::arg("input"), py::arg("weight"), py::arg("weight_des"), py::arg("weight_scale"), py::arg("weight_zero"),
          py::arg("bias"), py::arg("stride"), py::arg("padding"), py::arg("input_scale") = py::none(),
          py::arg("input_zero") = py::none(), py::arg("input_qmin") = py::none(), py::arg("input_qmax") = py::none(),
          py::arg("residual") = py::none(), py::arg("fuse_relu") = false);
    m.def("quantconv2d_chain", &quantconv2d_chain,
          "Consecutive fused quantized convs with int8 activations handed from one layer's epilogue to the next.",
          py::arg("input"), py::arg("layers"), py::arg("residual") = py::none(), py::arg("input_handoff") = py::none(),
          py::arg("emit_next") = py::none(), py::arg("emit_only") = false);
    m.def("quantconv2d_u8_nhwc", &quantconv2d_u8_nhwc,
          "Fused quantized conv on an already quantized NHWC(Cp) byte workspace (another layer's hand-off).",
          py::arg("q_nhwc"), py::arg("in_shape"), py::arg("layer"));
    m.def("fake_quantize", &fake_quantize,
          "(clamp(round(x / scale - zero), qmin, qmax) + zero) * scale for a per-tensor quantizer, one kernel.",
          py::arg("input"), py::arg("scale"), py::arg("zero"), py::arg("qmin"), py::arg("qmax"));
    m.def("max_pool2d", &max_pool2d, "fp32 NCHW max pooling (square kernel / stride, -inf padding, floor mode).",
          py::arg("input"), py::arg("kernel_size"), py::arg("stride"), py::arg("padding") = 0, py::arg("out") = py::none());
    m.def("avg_pool_global", &avg_pool_global, "fp32 NCHW global average pooling -> [N, C, 1, 1].", py::arg("input"));
    m.def("minmax", &minmax, "The range estimators' (xmin, xmax) in one pass, with the optional running / moving-average update.",
          py::arg("input"), py::arg("granularity"), py::arg("flag"), py::arg("symmetric"), py::arg("update_mode") = 0,
          py::arg("momentum") = 0.0, py::arg("run_min") = py::none(), py::arg("run_max") = py::none());
    m.def("kthvalue", &kthvalue, "k-th smallest value (of |x| with use_abs) of the estimator's flattened view: torch.kthvalue(...)[0].",
          py::arg("input"), py::arg("k"), py::arg("granularity"), py::arg("flag"), py::arg("use_abs") = false);
    // engine-level helpers (not part of the reference surface)
    m.def("_launch_count", []() { return (uint64_t)qb200_launch_count(); });
    m.def("_launch_count_reset", []() { qb200_launch_count_reset(); });
    m.def("_set_conv_algo", [](int a) { qb200_set_conv_algo(a); });
    m.def("_abi_version", []() { return qb200_version(); });
}

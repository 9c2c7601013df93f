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
#include <pybind11/pybind11.h>
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <list>
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
    bool operator==(const Key& o) const { return ptr == o.ptr && version == o.version && aux == o.aux; }
};
struct KeyHash {
    size_t operator()(const Key& k) const {
        return std::hash<const void*>()(k.ptr) ^ (std::hash<uint64_t>()(k.version) * 1000003u) ^
               (std::hash<const void*>()(k.aux) << 1);
    }
};

// An entry is valid only while the storage it was derived from is alive (a freed block can be handed out again
// at the same address), so every entry keeps a weak reference to that storage.
template <typename V>
class DerivedCache {
  public:
    explicit DerivedCache(size_t cap) : cap_(cap) {}
    V* find(const Key& k, const at::Tensor& src) {
        auto it = map_.find(k);
        if (it == map_.end()) return nullptr;
        auto live = it->second->weak.lock();
        if (!live || live.get() != src.storage().unsafeGetStorageImpl()) {
            order_.erase(it->second);
            map_.erase(it);
            return nullptr;
        }
        order_.splice(order_.begin(), order_, it->second);
        return &it->second->value;
    }
    V* insert(const Key& k, const at::Tensor& src, V v) {
        order_.push_front(Entry{k, c10::weak_intrusive_ptr<c10::StorageImpl>(src.storage().getWeakStorageImpl()), std::move(v)});
        map_[k] = order_.begin();
        while (order_.size() > cap_) {
            map_.erase(order_.back().key);
            order_.pop_back();
        }
        return &order_.front().value;
    }

  private:
    struct Entry {
        Key key;
        c10::weak_intrusive_ptr<c10::StorageImpl> weak;
        V value;
    };
    size_t cap_;
    std::list<Entry> order_;
    std::unordered_map<Key, typename std::list<Entry>::iterator, KeyHash> map_;
};

struct PreparedWeights {
    at::Tensor buffer;
    bool zero_is_zero;
};

std::mutex g_mu;
DerivedCache<std::vector<int64_t>> g_des_cache(4096);
DerivedCache<PreparedWeights> g_prep_cache(2048);
DerivedCache<at::Tensor> g_float_cache(4096);

Key key_of(const at::Tensor& t, const void* aux = nullptr) { return Key{t.data_ptr(), (uint64_t)t._version(), aux}; }

const std::vector<int64_t>& host_des(const at::Tensor& des) {
    const Key k = key_of(des);
    if (auto* v = g_des_cache.find(k, des)) return *v;
    auto dh = des.to(at::kCPU, at::kLong).contiguous();  // one blocking copy per descriptor (the reference: 6 per call)
    std::vector<int64_t> v(dh.data_ptr<int64_t>(), dh.data_ptr<int64_t>() + dh.numel());
    return *g_des_cache.insert(k, des, std::move(v));
}

// scale / zero / qmin / qmax as a device float pointer without per-call work
const float* device_float(const py::object& o, const at::Device& dev, std::vector<at::Tensor>& keep, const char* name) {
    at::Tensor t;
    if (THPVariable_Check(o.ptr())) {
        t = THPVariable_Unpack(o.ptr());
        TORCH_CHECK(t.numel() == 1, name, " must have exactly one element (per-tensor activation quantization)");
        if (t.device() == dev && t.scalar_type() == at::kFloat && t.is_contiguous()) {
            keep.push_back(t);
            return t.data_ptr<float>();
        }
        const Key k = key_of(t, reinterpret_cast<const void*>((intptr_t)dev.index() + 1));
        if (auto* v = g_float_cache.find(k, t)) return v->data_ptr<float>();
        auto f = t.detach().to(dev, at::kFloat).contiguous();
        return g_float_cache.insert(k, t, f)->data_ptr<float>();
    }
    // python number: one cached device scalar per (value, device)
    static std::unordered_map<int64_t, at::Tensor> scalars;
    const double val = o.cast<double>();
    float fv = (float)val;
    int32_t bits;
    memcpy(&bits, &fv, 4);
    const int64_t sk = ((int64_t)dev.index() << 32) | (uint32_t)bits;
    auto it = scalars.find(sk);
    if (it == scalars.end()) it = scalars.emplace(sk, at::full({1}, val, at::TensorOptions().dtype(at::kFloat).device(dev))).first;
    return it->second.data_ptr<float>();
}

// ------------------------------------------------------------------------------------------------
// quantconv2d_float_input
// ------------------------------------------------------------------------------------------------
at::Tensor conv_core(const at::Tensor& input, const at::Tensor& weight, const at::Tensor& weight_des, const std::vector<int64_t>& d,
                     const at::Tensor& weight_scale, const at::Tensor& weight_zero, const c10::optional<at::Tensor>& bias,
                     const int stride, const int padding, const py::object& input_scale, const py::object& input_zero,
                     const py::object& input_qmin, const py::object& input_qmax, const c10::optional<at::Tensor>& residual,
                     const bool fuse_relu);

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

    c10::cuda::CUDAGuard guard(input.device());
    std::lock_guard<std::mutex> lock(g_mu);

    const std::vector<int64_t>& d = host_des(weight_des);
    TORCH_CHECK(d.size() >= 6, "weight_des must hold [n_bits, sign, K, C, R, S]");
    return conv_core(input, weight, weight_des, d, weight_scale, weight_zero, bias, stride, padding, input_scale, input_zero,
                     input_qmin, input_qmax, residual, fuse_relu);
}

// (g_mu held, device guard set) d = [n_bits, sign, K, Cg, R, S]; input is 4-D
at::Tensor conv_core(const at::Tensor& input, const at::Tensor& weight, const at::Tensor& weight_des, const std::vector<int64_t>& d,
                     const at::Tensor& weight_scale, const at::Tensor& weight_zero, const c10::optional<at::Tensor>& bias,
                     const int stride, const int padding, const py::object& input_scale, const py::object& input_zero,
                     const py::object& input_qmin, const py::object& input_qmax, const c10::optional<at::Tensor>& residual,
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

    const bool fused = !input_scale.is_none();
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
    TORCH_CHECK(!input_zero.is_none() && !input_qmin.is_none() && !input_qmax.is_none(),
                "input_scale, input_zero, input_qmin and input_qmax must be given together");

    // prepared weights: derived once per (weight storage, version, descriptor)
    const Key wk = key_of(weight, weight_des.data_ptr());
    PreparedWeights* pw = g_prep_cache.find(wk, weight);
    if (!pw) {
        PreparedWeights fresh;
        fresh.buffer = at::empty({(int64_t)qb200_conv_prepared_bytes(&s)}, weight.options());
        check_rc(qb200_conv_prepare_weights(&s, weight.data_ptr<uint8_t>(), fresh.buffer.data_ptr(), st), "prepare_weights");
        fresh.zero_is_zero = weight_zero.abs().max().item<float>() == 0.f;  // one sync per weight tensor
        pw = g_prep_cache.insert(wk, weight, std::move(fresh));
    }

    std::vector<at::Tensor> keep;
    qb200_act_quant aq;
    aq.scale = device_float(input_scale, input.device(), keep, "input_scale");
    aq.zero = device_float(input_zero, input.device(), keep, "input_zero");
    aq.qmin = device_float(input_qmin, input.device(), keep, "input_qmin");
    aq.qmax = device_float(input_qmax, input.device(), keep, "input_qmax");

    if (!pw->zero_is_zero) {
        // asymmetric weights do not factor into an integer GEMM with a float zero point: fake-quantize the
        // activations on the device (quantizer.py:215-218) and run the fp32 weight-only kernel
        TORCH_CHECK(s.C == s.Cg, "asymmetric weights with groups > 1 are not supported");
        TORCH_CHECK(!residual.has_value() && !fuse_relu, "residual / fuse_relu are not supported with asymmetric weights");
        auto f = [&](const py::object& o) {
            return THPVariable_Check(o.ptr()) ? THPVariable_Unpack(o.ptr()).detach().to(input.device(), at::kFloat).reshape({1})
                                              : at::full({1}, o.cast<double>(), input.options());
        };
        auto sc = f(input_scale), ze = f(input_zero), lo = f(input_qmin), hi = f(input_qmax);
        auto xq = at::maximum(at::minimum(at::round(input / sc - ze), hi), lo);
        auto xdq = ((xq + ze) * sc).contiguous();
        check_rc(qb200_quantconv2d_weightonly(&s, xdq.data_ptr<float>(), weight.data_ptr<uint8_t>(),
                                              weight_scale.data_ptr<float>(), weight_zero.data_ptr<float>(), (int32_t)n_ws,
                                              bias_p, out.data_ptr<float>(), st),
                 "quantconv2d_float_input");
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
    qb200_conv_shape s;
    int32_t P, Q;
    qb200_act_quant aq;
    PreparedWeights* pw;
    bool relu;
};

py::object quantconv2d_chain(const at::Tensor& input, const py::list& layers, const c10::optional<at::Tensor>& residual,
                             const c10::optional<at::Tensor>& input_handoff, const py::object& emit_next) {
    CHECK_INPUT(input);
    CHECK_FLOAT(input);
    TORCH_CHECK(input.dim() == 4, "input must be a 4D tensor");
    const size_t n = layers.size();  // layers of the chain proper; an optional extra entry describes emit_next's consumer
    TORCH_CHECK(n >= 1, "quantconv2d_chain: no layers");
    const bool emit = !emit_next.is_none();
    c10::cuda::CUDAGuard guard(input.device());
    std::lock_guard<std::mutex> lock(g_mu);
    void* st = cur_stream();
    std::vector<at::Tensor> keep;
    std::vector<ChainLayer> L(n + (emit ? 1 : 0));
    int32_t N = (int32_t)input.size(0), C = (int32_t)input.size(1), H = (int32_t)input.size(2), W = (int32_t)input.size(3);
    for (size_t i = 0; i < L.size(); ++i) {
        const py::tuple t = (i < n ? py::object(layers[i]) : emit_next).cast<py::tuple>();
        TORCH_CHECK(t.size() == 12, "quantconv2d_chain: each layer is a 12-tuple");
        ChainLayer& l = L[i];
        l.weight = t[0].cast<at::Tensor>();
        l.des = t[1].cast<at::Tensor>();
        l.scale = t[2].cast<at::Tensor>();
        l.zero = t[3].cast<at::Tensor>();
        if (!t[4].is_none()) l.bias = t[4].cast<at::Tensor>();
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
        const std::vector<int64_t>& d = host_des(l.des);
        TORCH_CHECK(d.size() >= 6, "weight_des must hold [n_bits, sign, K, C, R, S]");
        qb200_conv_shape& s = l.s;
        s.N = N; s.C = C; s.H = H; s.W = W;
        s.w_bits = (int32_t)d[0];
        s.w_sign = d[1] != 0;
        s.K = (int32_t)d[2]; s.Cg = (int32_t)d[3]; s.R = (int32_t)d[4]; s.S = (int32_t)d[5];
        s.stride = t[5].cast<int>();
        s.pad = t[6].cast<int>();
        check_rc(qb200_conv_out_hw(&s, &l.P, &l.Q), "quantconv2d_chain");
        TORCH_CHECK(l.weight.numel() >= qb200_packed_bytes((int64_t)s.K * s.Cg * s.R * s.S, s.w_bits),
                    "weight is shorter than weight_des describes");
        const int64_t n_ws = l.scale.numel();
        TORCH_CHECK(n_ws == 1 || n_ws == s.K, "weight_scale must have 1 or ", s.K, " elements");
        TORCH_CHECK(l.zero.numel() == n_ws, "weight_zero must have as many elements as weight_scale");
        if (l.bias.has_value()) TORCH_CHECK(l.bias.value().numel() == s.K, "bias must have ", s.K, " elements");
        const Key wk = key_of(l.weight, l.des.data_ptr());
        l.pw = g_prep_cache.find(wk, l.weight);
        if (!l.pw) {
            PreparedWeights fresh;
            fresh.buffer = at::empty({(int64_t)qb200_conv_prepared_bytes(&s)}, l.weight.options());
            check_rc(qb200_conv_prepare_weights(&s, l.weight.data_ptr<uint8_t>(), fresh.buffer.data_ptr(), st), "prepare_weights");
            fresh.zero_is_zero = l.zero.abs().max().item<float>() == 0.f;
            l.pw = g_prep_cache.insert(wk, l.weight, std::move(fresh));
        }
        TORCH_CHECK(l.pw->zero_is_zero, "quantconv2d_chain needs symmetric weights (weight_zero == 0)");
        l.aq.scale = device_float(t[7], input.device(), keep, "input_scale");
        l.aq.zero = device_float(t[8], input.device(), keep, "input_zero");
        l.aq.qmin = device_float(t[9], input.device(), keep, "input_qmin");
        l.aq.qmax = device_float(t[10], input.device(), keep, "input_qmax");
        l.relu = t[11].cast<bool>();
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
    at::Tensor out, ws_emit;
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
        void* out_p = (hand && !last) ? nullptr : out.data_ptr();
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
    if (!emit) return py::cast(out);
    return py::make_tuple(out, ws_emit.defined() ? py::cast(ws_emit) : py::none());
}

// fake_quantize: Quantizer.simulate of a per-tensor quantizer as one kernel
at::Tensor fake_quantize(const at::Tensor& input, const py::object& scale, const py::object& zero, const py::object& qmin,
                         const py::object& qmax) {
    CHECK_INPUT(input);
    CHECK_FLOAT(input);
    c10::cuda::CUDAGuard guard(input.device());
    std::lock_guard<std::mutex> lock(g_mu);
    std::vector<at::Tensor> keep;
    qb200_act_quant aq;
    aq.scale = device_float(scale, input.device(), keep, "scale");
    aq.zero = device_float(zero, input.device(), keep, "zero");
    aq.qmin = device_float(qmin, input.device(), keep, "qmin");
    aq.qmax = device_float(qmax, input.device(), keep, "qmax");
    auto out = at::empty_like(input);
    check_rc(qb200_fake_quantize_f32(input.data_ptr<float>(), input.numel(), &aq, out.data_ptr<float>(), cur_stream()),
             "fake_quantize");
    return out;
}

// max_pool2d (engine helper for the packed ResNet forward; same result as torch.nn.functional.max_pool2d)
at::Tensor max_pool2d(const at::Tensor& input, int kernel, int stride, int padding) {
    CHECK_INPUT(input);
    CHECK_FLOAT(input);
    TORCH_CHECK(input.dim() == 4, "input must be a 4D tensor");
    c10::cuda::CUDAGuard guard(input.device());
    const int H = (int)input.size(2), W = (int)input.size(3);
    TORCH_CHECK(kernel >= 1 && stride >= 1 && padding >= 0 && 2 * padding <= kernel, "max_pool2d: bad geometry");
    const int P = (H + 2 * padding - kernel) / stride + 1, Q = (W + 2 * padding - kernel) / stride + 1;
    TORCH_CHECK(P >= 1 && Q >= 1, "max_pool2d: empty output");
    auto out = at::empty({input.size(0), input.size(1), P, Q}, input.options());
    check_rc(qb200_maxpool2d_f32(input.data_ptr<float>(), input.size(0) * input.size(1), H, W, kernel, stride, padding,
                                 out.data_ptr<float>(), cur_stream()),
             "max_pool2d");
    return out;
}

// ------------------------------------------------------------------------------------------------
// off-path ops: exported so that `from quant_engine import *` binds all 8 names; no module of
// the reference calls them today (quantconv2d.py:198-210 has the call commented out).
// ------------------------------------------------------------------------------------------------
[[noreturn]] void off_path(const char* name) {
    TORCH_CHECK(false, "quant_engine.", name,
                " is outside the hot path this engine implements (tpack, tunpack, quantconv2d_float_input)");
    abort();
}
at::Tensor linear(const at::Tensor&, const at::Tensor&, const c10::optional<at::Tensor>&, int) { off_path("linear"); }
at::Tensor quantlinear(const at::Tensor&, const at::Tensor&, const at::Tensor&, const at::Tensor&, const at::Tensor&,
                       const at::Tensor&, const at::Tensor&, const at::Tensor&, const c10::optional<at::Tensor>&) {
    off_path("quantlinear");
}
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
    c10::cuda::CUDAGuard guard(input.device());
    std::lock_guard<std::mutex> lock(g_mu);
    const std::vector<int64_t>& d = host_des(weight_des);
    TORCH_CHECK(d.size() >= 4, "weight_des must hold [n_bits, sign, out_features, in_features]");
    const int64_t B = input.size(0), in_f = input.size(1), out_f = d[2];
    TORCH_CHECK(d[3] == in_f, "input has ", in_f, " features, the weight expects ", d[3]);
    TORCH_CHECK(weight.numel() >= qb200_packed_bytes(out_f * in_f, (int)d[0]), "weight is shorter than weight_des describes");
    const int64_t n_ws = weight_scale.numel();
    TORCH_CHECK(n_ws == 1 || n_ws == out_f, "weight_scale must have 1 or ", out_f, " elements");
    TORCH_CHECK(weight_zero.numel() == n_ws, "weight_zero must have as many elements as weight_scale");
    if (bias.has_value()) TORCH_CHECK(bias.value().numel() == out_f, "bias must have ", out_f, " elements");
    if (input_scale.is_none()) {
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
    auto out4 = conv_core(input.view({B, in_f, 1, 1}), weight, weight_des, d6, weight_scale, weight_zero, bias, 1, 0, input_scale,
                          input_zero, input_qmin, input_qmax, c10::nullopt, false);
    return out4.view({B, out_f});
}
at::Tensor conv2d(const at::Tensor&, const at::Tensor&, const c10::optional<at::Tensor>&, int, int, int) { off_path("conv2d"); }
at::Tensor quantconv2d(const at::Tensor&, const at::Tensor&, const at::Tensor&, const at::Tensor&, const at::Tensor&,
                       const at::Tensor&, const at::Tensor&, const at::Tensor&, const c10::optional<at::Tensor>&, int, int) {
    off_path("quantconv2d");
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "B200 (sm_100a) implementation of JingInAI/Quantize's quant_engine hot path";
    m.def("tpack", &tpack, "Packs the given tensor into a vector of tensors.", py::arg("x"), py::arg("n_bits"), py::arg("sign"));
    m.def("tunpack", &tunpack, "Unpacks the given vector of tensors into a tensor.", py::arg("x"), py::arg("des"));
    m.def("linear", &linear, "Linear function.", py::arg("input"), py::arg("weight"), py::arg("bias") = py::none(),
          py::arg("mode") = 0);
    m.def("quantlinear", &quantlinear, "Quantized linear function.");
    m.def("quantlinear_float_input", &quantlinear_float_input, "Quantized linear function with float input.",
          py::arg("input"), py::arg("weight"), py::arg("weight_des"), py::arg("weight_scale"), py::arg("weight_zero"),
          py::arg("bias") = py::none(), py::arg("input_scale") = py::none(), py::arg("input_zero") = py::none(),
          py::arg("input_qmin") = py::none(), py::arg("input_qmax") = py::none());
    m.def("conv2d", &conv2d, "Conv2d function.", py::arg("input"), py::arg("weight"), py::arg("bias"), py::arg("stride"),
          py::arg("padding"), py::arg("mode") = 0);
    m.def("quantconv2d", &quantconv2d, "Quantized conv2d function.");
    m.def("quantconv2d_float_input", &quantconv2d_float_input, "Quantized conv2d function with float input.",
          py::arg("input"), py::arg("weight"), py::arg("weight_des"), py::arg("weight_scale"), py::arg("weight_zero"),
          py::arg("bias"), py::arg("stride"), py::arg("padding"), py::arg("input_scale") = py::none(),
          py::arg("input_zero") = py::none(), py::arg("input_qmin") = py::none(), py::arg("input_qmax") = py::none(),
          py::arg("residual") = py::none(), py::arg("fuse_relu") = false);
    m.def("quantconv2d_chain", &quantconv2d_chain,
          "Consecutive fused quantized convs with int8 activations handed from one layer's epilogue to the next.",
          py::arg("input"), py::arg("layers"), py::arg("residual") = py::none(), py::arg("input_handoff") = py::none(),
          py::arg("emit_next") = py::none());
    m.def("fake_quantize", &fake_quantize,
          "(clamp(round(x / scale - zero), qmin, qmax) + zero) * scale for a per-tensor quantizer, one kernel.",
          py::arg("input"), py::arg("scale"), py::arg("zero"), py::arg("qmin"), py::arg("qmax"));
    m.def("max_pool2d", &max_pool2d, "fp32 NCHW max pooling (square kernel / stride, -inf padding, floor mode).",
          py::arg("input"), py::arg("kernel_size"), py::arg("stride"), py::arg("padding") = 0);
    // engine-level helpers (not part of the reference surface)
    m.def("_launch_count", []() { return (uint64_t)qb200_launch_count(); });
    m.def("_launch_count_reset", []() { qb200_launch_count_reset(); });
    m.def("_set_conv_algo", [](int a) { qb200_set_conv_algo(a); });
    m.def("_abi_version", []() { return qb200_version(); });
}

// fewbit_b200 -- torch operator library `torch.ops.fewbit.*` (libfewbit.so).
//
// A thin layer: schema registration (verbatim from the reference, fewbit/fewbit.cc:5-39),
// argument checks, allocation of the packed state, autograd bookkeeping (mark_dirty +
// save_for_backward, as ContinousCudaFunction fewbit/cuda/activation.cc:337-382 and the
// eight *CudaFunction classes :23-330), and one call into the C ABI (include/fewbit_b200.h)
// per pass.  CUDA tensors only ever reach the CUDA kernels (no fallback); CPU tensors are served
// for the three operators the reference serves on CPU (gelu, quantize, quantize_backward).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/torch.h>

#include "fewbit_b200.h"

namespace {

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

int dtype_code(const Tensor &t, const char *what) {
    if (t.scalar_type() == torch::kFloat32) return FEWBIT_F32;
    if (t.scalar_type() == torch::kBFloat16) return FEWBIT_BF16;
    TORCH_CHECK(false, "fewbit: ", what, " must be float32 or bfloat16, got ", t.scalar_type());
}

void check_status(int status, const char *op) {
    TORCH_CHECK(status == FEWBIT_OK, "fewbit::", op, " failed: ", fewbit_error_string(status));
}

void check_activation(const Tensor &self, const char *op) {
    TORCH_CHECK(self.is_cuda(), "fewbit::", op,
                ": expected a CUDA tensor (this library has no CPU path), got ", self.device());
    TORCH_CHECK(self.is_contiguous(), "fewbit::", op, ": tensor must be contiguous");
    dtype_code(self, "activation");
}

// Tables arrive already cast by the Python layer (functional/activations.py:212 of the
// reference); direct callers (benchmark/bench-roberta.py:138-139) pass fp32 tables, so cast
// and make them contiguous on the activation's device here.
Tensor prepare_table(const Tensor &table, const Tensor &like, const char *what) {
    TORCH_CHECK(table.dim() == 1, "fewbit: `", what, "` must be one-dimensional");
    return table.to(like.device(), like.scalar_type()).contiguous();
}

Tensor new_state(const Tensor &like, int64_t n, int bits) {
    auto opts = torch::TensorOptions().device(like.device()).dtype(torch::kUInt8);
    return torch::empty({(int64_t)fewbit_state_bytes(n, bits)}, opts);
}

// ---------------------------------------------------------------- continuous family ----

class StepwiseFunction : public torch::autograd::Function<StepwiseFunction> {
public:
    static Tensor forward(AutogradContext *ctx, const Tensor &self, const Tensor &bounds,
                          const Tensor &levels, int64_t func, double p0, double p1) {
        check_activation(self, "stepwise_forward");
        TORCH_CHECK(levels.numel() >= 1 && levels.numel() <= 256,
                    "fewbit: maximal number of steps is limited to 256, got ", levels.numel());
        TORCH_CHECK(bounds.numel() + 1 == levels.numel(),
                    "fewbit: size of `bounds` should be lesser than size of `levels` by one, got ",
                    bounds.numel(), " and ", levels.numel());
        c10::cuda::CUDAGuard guard(self.device());
        auto stream = at::cuda::getCurrentCUDAStream();
        const int bits = fewbit_bits_for_levels((int)levels.numel());
        Tensor table = prepare_table(bounds, self, "bounds");
        Tensor values = prepare_table(levels, self, "levels");
        Tensor state = new_state(self, self.numel(), bits);
        ctx->mark_dirty({self});
        ctx->save_for_backward({state, values});
        ctx->saved_data["bits"] = (int64_t)bits;
        check_status(fewbit_stepwise_forward((int)func, dtype_code(self, "activation"),
                                             self.data_ptr(), self.data_ptr(),
                                             state.data_ptr<uint8_t>(), self.numel(), bits,
                                             table.data_ptr(), (int)table.numel(), p0, p1,
                                             stream.stream()),
                     "stepwise_forward");
        return self;
    }

    static variable_list backward(AutogradContext *ctx, variable_list grad_output) {
        auto saved = ctx->get_saved_variables();
        const Tensor &state = saved[0], &levels = saved[1];
        const int bits = (int)ctx->saved_data["bits"].toInt();
        Tensor gout = grad_output[0].contiguous();
        TORCH_CHECK(gout.scalar_type() == levels.scalar_type(),
                    "fewbit: gradient dtype ", gout.scalar_type(), " differs from activation dtype ",
                    levels.scalar_type());
        c10::cuda::CUDAGuard guard(gout.device());
        auto stream = at::cuda::getCurrentCUDAStream();
        Tensor gin = torch::empty_like(gout);
        check_status(fewbit_stepwise_backward(dtype_code(gout, "gradient"),
                                              state.data_ptr<uint8_t>(), gout.data_ptr(),
                                              gin.data_ptr(), gout.numel(), bits, levels.data_ptr(),
                                              (int)levels.numel(), stream.stream()),
                     "stepwise_backward");
        return {gin, Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
    }
};

#define FEWBIT_CONTINUOUS0(name, id)                                                     \
    Tensor name(Tensor &self, const Tensor &bounds, const Tensor &levels) {              \
        return StepwiseFunction::apply(self, bounds, levels, (int64_t)(id), 0.0, 0.0);   \
    }
#define FEWBIT_CONTINUOUS1(name, id)                                                     \
    Tensor name(Tensor &self, const Tensor &bounds, const Tensor &levels, double a) {    \
        return StepwiseFunction::apply(self, bounds, levels, (int64_t)(id), a, 0.0);     \
    }

FEWBIT_CONTINUOUS1(celu, FEWBIT_CELU)
FEWBIT_CONTINUOUS1(elu, FEWBIT_ELU)
FEWBIT_CONTINUOUS0(gelu, FEWBIT_GELU)
FEWBIT_CONTINUOUS0(hardswish, FEWBIT_HARDSWISH)
FEWBIT_CONTINUOUS0(logsigmoid, FEWBIT_LOGSIGMOID)
FEWBIT_CONTINUOUS0(mish, FEWBIT_MISH)
FEWBIT_CONTINUOUS0(selu, FEWBIT_SELU)
FEWBIT_CONTINUOUS0(sigmoid, FEWBIT_SIGMOID)
FEWBIT_CONTINUOUS0(silu, FEWBIT_SILU)
FEWBIT_CONTINUOUS0(softsign, FEWBIT_SOFTSIGN)
FEWBIT_CONTINUOUS0(tanh_, FEWBIT_TANH)
FEWBIT_CONTINUOUS0(tanhshrink, FEWBIT_TANHSHRINK)

Tensor softplus(Tensor &self, const Tensor &bounds, const Tensor &levels, double beta,
                double threshold) {
    return StepwiseFunction::apply(self, bounds, levels, (int64_t)FEWBIT_SOFTPLUS, beta, threshold);
}

// --------------------------------------------------------- custom table: `stepwise` ----
// Declared by the reference (fewbit/fewbit.cc:37) and wrapped by its `Stepwise` module
// (fewbit/modules/activations.py:97-134), but without any kernel there.  Here: the activation
// whose derivative the table describes -- continuous, piecewise linear, slopes `levels`, kinks at
// `bounds`, zero at the anchor -- in place, saving only the packed codes and the levels.
//
// `parity` / `shift` (README.md:111-112 of the reference: "the parity property allows ... to
// increase precision"): the table then describes only x > x0 and is mirrored about the point
// (x0, s0) = shift (default (0, 0)) -- even: s(x0 - t) = s(x0 + t); odd: s(x0 - t) = 2 s0 -
// s(x0 + t) -- which doubles the number of steps a table of a given size stands for.  The schema
// types `shift` as integers; the Python module expands non-integer shifts itself.
class CustomStepwiseFunction : public torch::autograd::Function<CustomStepwiseFunction> {
public:
    static Tensor forward(AutogradContext *ctx, const Tensor &self, const Tensor &bounds,
                          const Tensor &levels, double anchor) {
        check_activation(self, "stepwise");
        TORCH_CHECK(levels.numel() >= 1 && levels.numel() <= 256,
                    "fewbit: maximal number of steps is limited to 256, got ", levels.numel());
        TORCH_CHECK(bounds.numel() + 1 == levels.numel(),
                    "fewbit: size of `bounds` should be lesser than size of `levels` by one, got ",
                    bounds.numel(), " and ", levels.numel());
        c10::cuda::CUDAGuard guard(self.device());
        auto stream = at::cuda::getCurrentCUDAStream();
        const int bits = fewbit_bits_for_levels((int)levels.numel());
        Tensor table = prepare_table(bounds, self, "bounds");
        Tensor values = prepare_table(levels, self, "levels");
        Tensor state = new_state(self, self.numel(), bits);
        ctx->mark_dirty({self});
        ctx->save_for_backward({state, values});
        ctx->saved_data["bits"] = (int64_t)bits;
        check_status(fewbit_stepwise_custom_forward(dtype_code(self, "activation"), self.data_ptr(),
                                                    self.data_ptr(), state.data_ptr<uint8_t>(),
                                                    self.numel(), bits, table.data_ptr(),
                                                    (int)table.numel(), values.data_ptr(),
                                                    (int)values.numel(), anchor, stream.stream()),
                     "stepwise");
        return self;
    }

    static variable_list backward(AutogradContext *ctx, variable_list grad_output) {
        variable_list grads = StepwiseFunction::backward(ctx, std::move(grad_output));
        return {grads[0], Tensor(), Tensor(), Tensor()};
    }
};

// Not in the reference: the same operator with a real-valued anchor (the Python layer expands
// mirrored tables itself, with non-integer shifts, and calls this).
Tensor stepwise_anchored(Tensor &self, const Tensor &bounds, const Tensor &levels, double anchor) {
    return CustomStepwiseFunction::apply(self, bounds, levels, anchor);
}

Tensor stepwise(Tensor &self, const Tensor &bounds, const Tensor &levels, std::optional<bool> parity,
                at::OptionalIntArrayRef shift) {
    const double x0 = shift.has_value() ? (double)(*shift)[0] : 0.0;
    const double s0 = shift.has_value() ? (double)(*shift)[1] : 0.0;
    if (!parity.has_value()) return CustomStepwiseFunction::apply(self, bounds, levels, x0);
    TORCH_CHECK(bounds.dim() == 1 && levels.dim() == 1 && bounds.numel() + 1 == levels.numel(),
                "fewbit::stepwise: `bounds` (one-dimensional) must be one shorter than `levels`");
    TORCH_CHECK(2 * levels.numel() <= 256,
                "fewbit::stepwise: a mirrored table doubles its steps; at most 128 levels, got ", levels.numel());
    // even: the mirror image keeps its level; odd: it is reflected through s0
    Tensor mirrored = *parity ? levels.flip(0) : (2.0 * s0 - levels.flip(0));
    Tensor centre = torch::full({1}, x0, bounds.options());
    Tensor full_bounds = torch::cat({2.0 * x0 - bounds.flip(0), centre, bounds});
    Tensor full_levels = torch::cat({mirrored, levels});
    return CustomStepwiseFunction::apply(self, full_bounds, full_levels, x0);
}

// ------------------------------------------------------------------- 1-bit family ----

class PiecewiseFunction : public torch::autograd::Function<PiecewiseFunction> {
public:
    static Tensor forward(AutogradContext *ctx, const Tensor &self, int64_t func, double p0,
                          double p1) {
        check_activation(self, "piecewise_forward");
        c10::cuda::CUDAGuard guard(self.device());
        auto stream = at::cuda::getCurrentCUDAStream();
        Tensor state = new_state(self, self.numel(), 1);
        ctx->mark_dirty({self});
        ctx->save_for_backward({state});
        ctx->saved_data["func"] = func;
        ctx->saved_data["p0"] = p0;  // plain double: no device sync in backward (the reference
                                     // keeps the slope in a tensor and .item()s it)
        check_status(fewbit_piecewise_forward((int)func, dtype_code(self, "activation"),
                                              self.data_ptr(), self.data_ptr(),
                                              state.data_ptr<uint8_t>(), self.numel(), p0, p1,
                                              stream.stream()),
                     "piecewise_forward");
        return self;
    }

    static variable_list backward(AutogradContext *ctx, variable_list grad_output) {
        auto saved = ctx->get_saved_variables();
        Tensor gout = grad_output[0].contiguous();
        c10::cuda::CUDAGuard guard(gout.device());
        auto stream = at::cuda::getCurrentCUDAStream();
        Tensor gin = torch::empty_like(gout);
        check_status(fewbit_piecewise_backward((int)ctx->saved_data["func"].toInt(),
                                               dtype_code(gout, "gradient"),
                                               saved[0].data_ptr<uint8_t>(), gout.data_ptr(),
                                               gin.data_ptr(), gout.numel(),
                                               ctx->saved_data["p0"].toDouble(), stream.stream()),
                     "piecewise_backward");
        return {gin, Tensor(), Tensor(), Tensor()};
    }
};

Tensor hardshrink(Tensor &self, double lambd) {
    return PiecewiseFunction::apply(self, (int64_t)FEWBIT_HARDSHRINK, lambd, 0.0);
}
Tensor hardsigmoid(Tensor &self) {
    return PiecewiseFunction::apply(self, (int64_t)FEWBIT_HARDSIGMOID, 0.0, 0.0);
}
Tensor hardtanh(Tensor &self, double min_val, double max_val) {
    return PiecewiseFunction::apply(self, (int64_t)FEWBIT_HARDTANH, min_val, max_val);
}
Tensor leaky_relu(Tensor &self, double negative_slope) {
    return PiecewiseFunction::apply(self, (int64_t)FEWBIT_LEAKY_RELU, negative_slope, 0.0);
}
Tensor relu(Tensor &self) { return PiecewiseFunction::apply(self, (int64_t)FEWBIT_RELU, 0.0, 0.0); }
Tensor relu6(Tensor &self) {
    return PiecewiseFunction::apply(self, (int64_t)FEWBIT_RELU6, 0.0, 0.0);
}
Tensor softshrink(Tensor &self, double lambd) {
    return PiecewiseFunction::apply(self, (int64_t)FEWBIT_SOFTSHRINK, lambd, 0.0);
}
Tensor threshold(Tensor &self, double threshold, double value) {
    return PiecewiseFunction::apply(self, (int64_t)FEWBIT_THRESHOLD, threshold, value);
}

// ------------------------------------------- quantize / quantize_backward (GELU) ----
// Out-of-place pair of the reference (fewbit/cpu/gelu.cc:7-31, 33-45), here for CUDA tensors.

std::tuple<Tensor, Tensor> quantize(const Tensor &inputs, const Tensor &bounds) {
    check_activation(inputs, "quantize");
    TORCH_CHECK(bounds.numel() >= 1 && bounds.numel() <= 255, "fewbit: 1..255 bounds expected");
    c10::cuda::CUDAGuard guard(inputs.device());
    auto stream = at::cuda::getCurrentCUDAStream();
    const int bits = fewbit_bits_for_levels((int)bounds.numel() + 1);
    Tensor table = prepare_table(bounds, inputs, "bounds");
    Tensor outputs = torch::empty_like(inputs);
    Tensor state = new_state(inputs, inputs.numel(), bits);
    check_status(fewbit_stepwise_forward(FEWBIT_GELU, dtype_code(inputs, "inputs"),
                                         inputs.data_ptr(), outputs.data_ptr(),
                                         state.data_ptr<uint8_t>(), inputs.numel(), bits,
                                         table.data_ptr(), (int)table.numel(), 0.0, 0.0,
                                         stream.stream()),
                 "quantize");
    return {outputs, state};
}

Tensor quantize_backward(const Tensor &grads, const Tensor &buffer, const Tensor &levels) {
    check_activation(grads, "quantize_backward");
    TORCH_CHECK(buffer.is_cuda() && buffer.scalar_type() == torch::kUInt8 && buffer.is_contiguous(),
                "fewbit::quantize_backward: `buffer` must be a contiguous CUDA uint8 tensor");
    TORCH_CHECK(levels.numel() >= 1 && levels.numel() <= 256, "fewbit: 1..256 levels expected");
    const int bits = fewbit_bits_for_levels((int)levels.numel());
    TORCH_CHECK((size_t)buffer.numel() >= fewbit_state_bytes(grads.numel(), bits),
                "fewbit::quantize_backward: `buffer` holds ", buffer.numel(), " bytes, ",
                fewbit_state_bytes(grads.numel(), bits), " needed");
    c10::cuda::CUDAGuard guard(grads.device());
    auto stream = at::cuda::getCurrentCUDAStream();
    Tensor values = prepare_table(levels, grads, "levels");
    Tensor gin = torch::empty_like(grads);
    check_status(fewbit_stepwise_backward(dtype_code(grads, "grads"), buffer.data_ptr<uint8_t>(),
                                          grads.data_ptr(), gin.data_ptr(), grads.numel(), bits,
                                          values.data_ptr(), (int)values.numel(), stream.stream()),
                 "quantize_backward");
    return gin;
}

// ------------------------------------------------ RandomizedLinear projection ----
// out[rows, D] = scale * S[rows, N] x[N, D], S generated inside the tcgen05 kernel from
// (seed, offset); replaces randn + matmul of LinearGRPFunc (fewbit/functional/linear.py:133-137).

// sketch_to: the result in fp32 or bf16 (rounded once, in the kernel), and with `column_sums` one more
// row = scale * x.sum(0), the bias gradient of LinearGRPFunc.backward (fewbit/functional/linear.py:217).
Tensor sketch_to(const Tensor &x, int64_t rows, int64_t seed, int64_t offset, int64_t kind, double scale,
                 bool bf16_out, bool column_sums) {
    TORCH_CHECK(x.is_cuda() && x.dim() == 2 && x.is_contiguous(),
                "fewbit::sketch: expected a contiguous 2-D CUDA tensor [tokens, features]");
    TORCH_CHECK(x.scalar_type() == torch::kBFloat16, "fewbit::sketch: x must be bfloat16, got ", x.scalar_type());
    TORCH_CHECK(x.size(1) % 8 == 0, "fewbit::sketch: the feature count must be a multiple of 8, got ", x.size(1));
    TORCH_CHECK(rows > 0, "fewbit::sketch: rows must be positive");
    c10::cuda::CUDAGuard guard(x.device());
    auto stream = at::cuda::getCurrentCUDAStream();
    const int64_t out_rows = rows + (column_sums ? 1 : 0);
    Tensor out = torch::empty({out_rows, x.size(1)}, x.options().dtype(bf16_out ? torch::kBFloat16 : torch::kFloat32));
    const size_t nbytes = fewbit_sketch_workspace_bytes(x.size(0), (int)x.size(1), (int)out_rows);
    Tensor workspace = torch::empty({(int64_t)std::max<size_t>(nbytes, 16)}, x.options().dtype(torch::kUInt8));
    check_status(fewbit_sketch_project(x.data_ptr(), out.data_ptr(), bf16_out ? FEWBIT_BF16 : FEWBIT_F32,
                                       workspace.data_ptr(), x.size(0), (int)x.size(1), (int)rows, column_sums ? 1 : 0,
                                       (int)kind, (float)scale, (uint64_t)seed, (uint64_t)offset, stream.stream()),
                 "sketch");
    return out;
}

Tensor sketch(const Tensor &x, int64_t rows, int64_t seed, int64_t offset, int64_t kind, double scale) {
    return sketch_to(x, rows, seed, offset, kind, scale, false, false);
}

Tensor sketch_matrix(const Tensor &like, int64_t rows, int64_t cols, int64_t seed, int64_t offset, int64_t kind) {
    TORCH_CHECK(like.is_cuda(), "fewbit::sketch_matrix: `like` must be a CUDA tensor");
    c10::cuda::CUDAGuard guard(like.device());
    auto stream = at::cuda::getCurrentCUDAStream();
    Tensor s = torch::empty({rows, cols}, like.options().dtype(torch::kBFloat16));
    check_status(fewbit_sketch_matrix(s.data_ptr(), (int)rows, cols, (int)kind, (uint64_t)seed, (uint64_t)offset,
                                      stream.stream()),
                 "sketch_matrix");
    return s;
}

// --------------------------------------------------- CPU tensors (dispatch key CPU) ----
// The reference serves CPU tensors for gelu / quantize / quantize_backward
// (fewbit/cpu/gelu.cc:7-76).  Same semantics here -- ATen for the values and the bucket search,
// and the exact-`bits` LSB-first stream -- but packing and unpacking run in parallel over octets
// (the reference's Deflate / Inflate are single-threaded by construction).  This is the CPU side
// of the device switch, not a fallback: CUDA tensors never come here.

int bits_of(int64_t nlevels) { return fewbit_bits_for_levels((int)nlevels); }

Tensor pack_codes_cpu(const Tensor &codes, int bits) {   // codes: contiguous int32
    const int64_t n = codes.numel();
    Tensor state = torch::zeros({(int64_t)fewbit_state_bytes(n, bits)}, torch::kUInt8);
    const int32_t *src = codes.data_ptr<int32_t>();
    uint8_t *dst = state.data_ptr<uint8_t>();
    const int64_t noctets = (n + 7) / 8, nbytes = state.numel();
    at::parallel_for(0, noctets, 4096, [&](int64_t begin, int64_t end) {
        for (int64_t o = begin; o < end; ++o) {
            uint64_t octet = 0;
            for (int j = 0; j < 8 && 8 * o + j < n; ++j)
                octet |= (uint64_t)((uint32_t)src[8 * o + j] & ((1u << bits) - 1u)) << (bits * j);
            for (int k = 0; k < bits && o * bits + k < nbytes; ++k) dst[o * bits + k] = (uint8_t)(octet >> (8 * k));
        }
    });
    return state;
}

Tensor unpack_codes_cpu(const Tensor &state, int64_t n, int bits) {   // -> int64 codes
    Tensor codes = torch::empty({n}, torch::kInt64);
    const uint8_t *src = state.data_ptr<uint8_t>();
    int64_t *dst = codes.data_ptr<int64_t>();
    const int64_t noctets = (n + 7) / 8, nbytes = state.numel();
    at::parallel_for(0, noctets, 4096, [&](int64_t begin, int64_t end) {
        for (int64_t o = begin; o < end; ++o) {
            uint64_t octet = 0;
            for (int k = 0; k < bits && o * bits + k < nbytes; ++k) octet |= (uint64_t)src[o * bits + k] << (8 * k);
            for (int j = 0; j < 8 && 8 * o + j < n; ++j) dst[8 * o + j] = (int64_t)((octet >> (bits * j)) & ((1u << bits) - 1u));
        }
    });
    return codes;
}

std::tuple<Tensor, Tensor> quantize_cpu(const Tensor &inputs, const Tensor &bounds) {
    TORCH_CHECK(bounds.dim() == 1 && bounds.numel() >= 1 && bounds.numel() <= 255, "fewbit: 1..255 bounds expected");
    Tensor x = inputs.contiguous();
    Tensor outputs = torch::gelu(x);
    Tensor codes = torch::searchsorted(bounds.to(x.scalar_type()).contiguous(), x, /*out_int32=*/true);
    return {outputs, pack_codes_cpu(codes.reshape({-1}).contiguous(), bits_of(bounds.numel() + 1))};
}

Tensor quantize_backward_cpu(const Tensor &grads, const Tensor &buffer, const Tensor &levels) {
    TORCH_CHECK(buffer.scalar_type() == torch::kUInt8 && buffer.is_contiguous(),
                "fewbit::quantize_backward: `buffer` must be a contiguous uint8 tensor");
    const int bits = bits_of(levels.numel());
    TORCH_CHECK((size_t)buffer.numel() >= fewbit_state_bytes(grads.numel(), bits),
                "fewbit::quantize_backward: `buffer` is too short");
    Tensor codes = unpack_codes_cpu(buffer, grads.numel(), bits).view(grads.sizes());
    return levels.to(grads.scalar_type()).index({codes}) * grads;
}

class GeluCpuFunction : public torch::autograd::Function<GeluCpuFunction> {
public:
    static Tensor forward(AutogradContext *ctx, const Tensor &inputs, const Tensor &bounds, const Tensor &levels) {
        TORCH_CHECK(bounds.numel() + 1 == levels.numel(),
                    "fewbit: size of `bounds` should be lesser than size of `levels` by one, got ",
                    bounds.numel(), " and ", levels.numel());
        auto [outputs, buffer] = quantize_cpu(inputs, bounds);
        ctx->save_for_backward({buffer, levels});
        return outputs;   // out of place on CPU, exactly as the reference
    }
    static variable_list backward(AutogradContext *ctx, variable_list grad_output) {
        auto saved = ctx->get_saved_variables();
        return {quantize_backward_cpu(grad_output[0].contiguous(), saved[0], saved[1]), Tensor(), Tensor()};
    }
};

Tensor gelu_cpu(Tensor &self, const Tensor &bounds, const Tensor &levels) {
    return GeluCpuFunction::apply(self, bounds, levels);
}

}  // namespace

// Schemas: verbatim from the reference (fewbit/fewbit.cc:6-37) -- they are the public ABI.
TORCH_LIBRARY(fewbit, m) {
    m.def("quantize(Tensor inputs, Tensor bounds) -> (Tensor, Tensor)");
    m.def("quantize_backward(Tensor grads, Tensor buffer, Tensor levels) -> Tensor");

    m.def("hardshrink (Tensor(a!) self, float lambd = 0.5) -> Tensor(a!)");
    m.def("hardsigmoid(Tensor(a!) self) -> Tensor(a!)");
    m.def("hardtanh   (Tensor(a!) self, float min_val = -1.0, float max_val = 1.0) -> Tensor(a!)");
    m.def("leaky_relu (Tensor(a!) self, float negative_slope = 0.01) -> Tensor(a!)");
    m.def("relu       (Tensor(a!) self) -> Tensor(a!)");
    m.def("relu6      (Tensor(a!) self) -> Tensor(a!)");
    m.def("softshrink (Tensor(a!) self, float lambd = 0.5) -> Tensor(a!)");
    m.def("threshold  (Tensor(a!) self, float threshold, float value) -> Tensor(a!)");

    m.def("celu      (Tensor(a!) self, Tensor bounds, Tensor levels, float alpha = 1.0) -> Tensor(a!)");
    m.def("elu       (Tensor(a!) self, Tensor bounds, Tensor levels, float alpha = 1.0) -> Tensor(a!)");
    m.def("gelu      (Tensor(a!) self, Tensor bounds, Tensor levels) -> Tensor(a!)");
    m.def("hardswish (Tensor(a!) self, Tensor bounds, Tensor levels) -> Tensor(a!)");
    m.def("logsigmoid(Tensor(a!) self, Tensor bounds, Tensor levels) -> Tensor(a!)");
    m.def("mish      (Tensor(a!) self, Tensor bounds, Tensor levels) -> Tensor(a!)");
    m.def("selu      (Tensor(a!) self, Tensor bounds, Tensor levels) -> Tensor(a!)");
    m.def("sigmoid   (Tensor(a!) self, Tensor bounds, Tensor levels) -> Tensor(a!)");
    m.def("silu      (Tensor(a!) self, Tensor bounds, Tensor levels) -> Tensor(a!)");
    m.def("softplus  (Tensor(a!) self, Tensor bounds, Tensor levels, float beta = 1.0, float threshold = 20.0) -> Tensor(a!)");
    m.def("softsign  (Tensor(a!) self, Tensor bounds, Tensor levels) -> Tensor(a!)");
    m.def("tanh      (Tensor(a!) self, Tensor bounds, Tensor levels) -> Tensor(a!)");
    m.def("tanhshrink(Tensor(a!) self, Tensor bounds, Tensor levels) -> Tensor(a!)");

    // Not in the reference: the projection of RandomizedLinear as one operator.
    m.def("sketch(Tensor x, int rows, int seed, int offset, int kind, float scale) -> Tensor");
    m.def("sketch_to(Tensor x, int rows, int seed, int offset, int kind, float scale, bool bf16_out, bool column_sums) -> Tensor");
    m.def("sketch_matrix(Tensor like, int rows, int cols, int seed, int offset, int kind) -> Tensor");
    m.def("stepwise_anchored(Tensor(a!) self, Tensor bounds, Tensor levels, float anchor = 0.0) -> Tensor(a!)");

    // Declared by the reference without any kernel (fewbit/fewbit.cc:37); the schema is kept
    // verbatim, the kernel is this package's (CustomStepwiseFunction above).
    m.def("stepwise   (Tensor(a!) self, Tensor bounds, Tensor levels, bool? parity=None, int[2]? shift=None) -> Tensor(a!)");
}

// Same dispatch key as the reference (fewbit/cuda/activation.cc:445-470).
TORCH_LIBRARY_IMPL(fewbit, AutogradCUDA, m) {
    m.impl("hardshrink", hardshrink);
    m.impl("hardsigmoid", hardsigmoid);
    m.impl("hardtanh", hardtanh);
    m.impl("leaky_relu", leaky_relu);
    m.impl("relu", relu);
    m.impl("relu6", relu6);
    m.impl("softshrink", softshrink);
    m.impl("threshold", threshold);

    m.impl("celu", celu);
    m.impl("elu", elu);
    m.impl("gelu", gelu);
    m.impl("hardswish", hardswish);
    m.impl("logsigmoid", logsigmoid);
    m.impl("mish", mish);
    m.impl("selu", selu);
    m.impl("sigmoid", sigmoid);
    m.impl("silu", silu);
    m.impl("softplus", softplus);
    m.impl("softsign", softsign);
    m.impl("tanh", tanh_);
    m.impl("tanhshrink", tanhshrink);
    m.impl("stepwise", stepwise);
    m.impl("stepwise_anchored", stepwise_anchored);
}

// As the reference: only gelu and the quantize pair exist for CPU tensors (fewbit/cpu/gelu.cc:74-76).
TORCH_LIBRARY_IMPL(fewbit, AutogradCPU, m) { m.impl("gelu", gelu_cpu); }
TORCH_LIBRARY_IMPL(fewbit, CPU, m) {
    m.impl("quantize", quantize_cpu);
    m.impl("quantize_backward", quantize_backward_cpu);
}

TORCH_LIBRARY_IMPL(fewbit, CUDA, m) {
    m.impl("quantize", quantize);
    m.impl("quantize_backward", quantize_backward);
    m.impl("sketch", sketch);
    m.impl("sketch_to", sketch_to);
    m.impl("sketch_matrix", sketch_matrix);
}

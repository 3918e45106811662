// fewbit_b200 -- the 1-bit (mask) family: forward value + mask pack, backward mask * grad.
#include "launch.cuh"

namespace fewbit {

template <class Fn, typename T> static cudaError_t fwd(const ForwardArgs &a) {
    MaskOp<Fn> op{Fn(a.p0, a.p1)};
    return launch_forward<decltype(op), T>(static_cast<const T *>(a.x), static_cast<T *>(a.y),
                                           a.state, a.n, op, a.stream);
}

template <class Fn> static cudaError_t fwd_any(const ForwardArgs &a) {
    return a.dtype == 0 ? fwd<Fn, float>(a) : fwd<Fn, __nv_bfloat16>(a);
}

// `func` indexes fewbit_piecewise_t (include/fewbit_b200.h).
cudaError_t launch_piecewise_forward(int func, const ForwardArgs &a) {
    switch (func) {
        case 0: return fwd_any<HardshrinkFn>(a);
        case 1: return fwd_any<HardsigmoidFn>(a);
        case 2: return fwd_any<HardtanhFn>(a);
        case 3: return fwd_any<LeakyReluFn>(a);
        case 4: return fwd_any<ReluFn>(a);
        case 5: return fwd_any<Relu6Fn>(a);
        case 6: return fwd_any<SoftshrinkFn>(a);
        case 7: return fwd_any<ThresholdFn>(a);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_piecewise_backward(int func, const BackwardArgs &a) {
    MaskFactorOp op{1.0f, 0.0f};
    if (func == 1) op = MaskFactorOp{1.0f / 6.0f, 0.0f};  // hardsigmoid
    if (func == 3) op = MaskFactorOp{(float)a.p0, 1.0f};  // leaky_relu: mask = negative side
    if (a.dtype == 0)
        return launch_backward<MaskFactorOp, float>(a.state, static_cast<const float *>(a.gout),
                                                    static_cast<float *>(a.gin), a.n, op, a.stream);
    return launch_backward<MaskFactorOp, __nv_bfloat16>(
        a.state, static_cast<const __nv_bfloat16 *>(a.gout), static_cast<__nv_bfloat16 *>(a.gin),
        a.n, op, a.stream);
}

}  // namespace fewbit

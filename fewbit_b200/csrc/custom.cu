// fewbit_b200 -- forward of the custom-table operator `stepwise` (IntegralOp, ops.cuh).
#include "launch.cuh"

namespace fewbit {

template <typename T, int B> static cudaError_t run(const ForwardArgs &a, const void *levels, int nlevels) {
    IntegralOp<T, B> op{{static_cast<const T *>(a.table), a.ntable}, static_cast<const T *>(levels), nlevels,
                        (float)a.p0};
    return launch_forward<decltype(op), T>(static_cast<const T *>(a.x), static_cast<T *>(a.y), a.state, a.n, op,
                                           a.stream);
}

// a.table / a.ntable: the borders; a.p0: the anchor (F(anchor) = 0).
cudaError_t launch_custom_forward(const ForwardArgs &a, const void *levels, int nlevels) {
    cudaError_t err = cudaErrorInvalidValue;
    if (a.dtype == 0) {
        FEWBIT_DISPATCH_BITS(a.bits, err = (run<float, B>(a, levels, nlevels)));
    } else {
        FEWBIT_DISPATCH_BITS(a.bits, err = (run<__nv_bfloat16, B>(a, levels, nlevels)));
    }
    return err;
}

}  // namespace fewbit

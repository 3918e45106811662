// fewbit_b200 -- unpack + level lookup + multiply for every continuous activation.
#include "launch.cuh"

namespace fewbit {

template <typename T, int B> static cudaError_t run(const BackwardArgs &a) {
    LevelsOp<T, B> op{static_cast<const T *>(a.table), a.ntable, nullptr};
    return launch_backward<decltype(op), T>(a.state, static_cast<const T *>(a.gout),
                                            static_cast<T *>(a.gin), a.n, op, a.stream);
}

cudaError_t launch_levels_backward(const BackwardArgs &a) {
    cudaError_t err = cudaErrorInvalidValue;
    if (a.dtype == 0) {
        FEWBIT_DISPATCH_BITS(a.bits, err = (run<float, B>(a)));
    } else {
        FEWBIT_DISPATCH_BITS(a.bits, err = (run<__nv_bfloat16, B>(a)));
    }
    return err;
}

}  // namespace fewbit

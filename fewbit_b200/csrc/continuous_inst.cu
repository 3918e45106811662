// fewbit_b200 -- one translation unit per continuous activation (compiled 13 times with
// -DFEWBIT_FN=<functor> -DFEWBIT_ENTRY=<symbol>; see Makefile) so the 13 x 2 dtypes x 8 bit
// widths instantiate in parallel.
#include "launch.cuh"

#ifndef FEWBIT_FN
#error "compile with -DFEWBIT_FN=<functor from ops.cuh> -DFEWBIT_ENTRY=<launcher symbol>"
#endif

namespace fewbit {

template <typename T, int B> static cudaError_t run(const ForwardArgs &a) {
    QuantizeOp<FEWBIT_FN, T, B> op{FEWBIT_FN(a.p0, a.p1),
                                   {static_cast<const T *>(a.table), a.ntable}};
    return launch_forward<decltype(op), T>(static_cast<const T *>(a.x), static_cast<T *>(a.y),
                                           a.state, a.n, op, a.stream);
}

cudaError_t FEWBIT_ENTRY(const ForwardArgs &a) {
    cudaError_t err = cudaErrorInvalidValue;
    if (a.dtype == 0) {
        FEWBIT_DISPATCH_BITS(a.bits, err = (run<float, B>(a)));
    } else {
        FEWBIT_DISPATCH_BITS(a.bits, err = (run<__nv_bfloat16, B>(a)));
    }
    return err;
}

}  // namespace fewbit

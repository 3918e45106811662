// TEST INFRASTRUCTURE -- not product code.
//
// extern "C" shim around the reference's *header-only* CPU codec so that tests can
// call the real thing through ctypes.  The header is included from where it lies
// under the reference tree (-I$(REF)); no reference source is copied here.
//   fewbit::Deflate  -> /root/reference/fewbit/cpu/codec.h:33-57
//   fewbit::Inflate  -> /root/reference/fewbit/cpu/codec.h:59-83
#include <cstdint>

#include <fewbit/cpu/codec.h>

extern "C" {

void ref_deflate_u8(const int32_t *codes, int64_t n, int32_t bits, uint8_t *out) {
    if (n > 0) fewbit::Deflate<uint8_t>(codes, codes + n, out, bits);
}

void ref_inflate_u8(int32_t *codes, int64_t n, int32_t bits, const uint8_t *in) {
    if (n > 0) fewbit::Inflate<uint8_t>(codes, codes + n, in, bits);
}

}  // extern "C"

// Host build of csrc/ec.cuh (plain-C emulation bodies) for CPU unit tests.
#include "ec.cuh"
#include <cstring>
using namespace zk;
extern "C" {
// acc (XYZZ, 32 u32) += affine p (16 u32)
void h_xyzz_add_affine(uint32_t* acc, const uint32_t* p) {
    G1Xyzz a; G1Affine q; memcpy(&a, acc, 128); memcpy(&q, p, 64);
    a.add_affine(q); memcpy(acc, &a, 128);
}
void h_xyzz_add(uint32_t* acc, const uint32_t* o) {
    G1Xyzz a, b; memcpy(&a, acc, 128); memcpy(&b, o, 128);
    a.add(b); memcpy(acc, &a, 128);
}
void h_xyzz_dbl(uint32_t* acc) {
    G1Xyzz a; memcpy(&a, acc, 128); a = a.dbl(); memcpy(acc, &a, 128);
}
void h_xyzz_to_jac(const uint32_t* acc, uint32_t* out, int normalized) {
    G1Xyzz a; memcpy(&a, acc, 128);
    G1Jacobian j = normalized ? a.to_jacobian_normalized() : a.to_jacobian();
    memcpy(out, &j, 96);
}
}

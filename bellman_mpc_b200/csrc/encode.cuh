// ZCash BLS12-381 point encodings on the device (G1Affine::to_uncompressed /
// from_uncompressed_unchecked, used by Parameters / Proof I/O in src/groth16/mod.rs:42-48,
// 261-290): big-endian coordinates, flag bits in the first byte.
#pragma once
#include "curve.cuh"

namespace bmpc {

// Big-endian ZCash uncompressed encoding of an affine point (G1Affine::to_uncompressed).
__device__ __forceinline__ void put_fp_be(const Fp& m, uint8_t* out) {
    Fp c = m.from_mont();
    for (int j = 0; j < 12; j++) {
        uint32_t v = c.l[11 - j];
        out[4 * j] = (uint8_t)(v >> 24); out[4 * j + 1] = (uint8_t)(v >> 16);
        out[4 * j + 2] = (uint8_t)(v >> 8); out[4 * j + 3] = (uint8_t)v;
    }
}
__device__ __forceinline__ void put_coord_be(const Fp& m, uint8_t* out) { put_fp_be(m, out); }
__device__ __forceinline__ void put_coord_be(const Fp2& m, uint8_t* out) {
    put_fp_be(m.c1, out);       // c1 first (same order the reference prints at gt_bytes.rs:41-59)
    put_fp_be(m.c0, out + 48);
}
template <class F>
__device__ __forceinline__ void encode_uncompressed(const Affine<F>& p, uint8_t* out) {
    const int CB = sizeof(F);
    if (p.is_identity()) {
        for (int j = 0; j < 2 * CB; j++) out[j] = 0;
        out[0] = 0x40;
        return;
    }
    put_coord_be(p.x, out);
    put_coord_be(p.y, out + CB);
}

__device__ __forceinline__ bool get_fp_be(const uint8_t* in, uint8_t mask0, Fp& out) {
    Fp c;
    for (int j = 0; j < 12; j++) {
        uint32_t b0 = in[4 * j], b1 = in[4 * j + 1], b2 = in[4 * j + 2], b3 = in[4 * j + 3];
        if (j == 0) b0 &= mask0;
        c.l[11 - j] = (b0 << 24) | (b1 << 16) | (b2 << 8) | b3;
    }
    out = c.to_mont();
    return true;
}
__device__ __forceinline__ void get_coord_be(const uint8_t* in, uint8_t mask0, Fp& out) {
    get_fp_be(in, mask0, out);
}
__device__ __forceinline__ void get_coord_be(const uint8_t* in, uint8_t mask0, Fp2& out) {
    get_fp_be(in, mask0, out.c1);
    get_fp_be(in + 48, 0xff, out.c0);
}

template <class T>
__device__ __forceinline__ T load_struct(const T* p) {
    T r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4* w = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int j = 0; j < (int)(sizeof(T) / 16); j++) w[j] = q[j];
    return r;
}
template <class T>
__device__ __forceinline__ void store_struct(T* p, const T& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    const uint4* w = reinterpret_cast<const uint4*>(&v);
#pragma unroll
    for (int j = 0; j < (int)(sizeof(T) / 16); j++) q[j] = w[j];
}

}  // namespace bmpc

// fastdiv.cuh -- division of 64-bit unsigned values by a loop-invariant divisor (shape extents, broadcast factors,
// group widths) with one multiply-high instead of the ~60-instruction 64-bit division sequence.
// Granlund & Montgomery, "Division by invariant integers using multiplication" (N = 64, round-up method).
#pragma once
#include <stdint.h>

namespace b2s {

struct FastDiv {
    uint64_t d;      // divisor
    uint64_t m;      // magic multiplier (unused for powers of two)
    int shift;       // log2(d) for powers of two, else ceil(log2 d)
    int pow2;

    __host__ __device__ __forceinline__ uint64_t div(uint64_t n) const {
#ifdef __CUDA_ARCH__
        if (pow2) return n >> shift;
        const uint64_t t = __umul64hi(m, n);
        return (t + ((n - t) >> 1)) >> (shift - 1);
#else
        return n / d;
#endif
    }
    // quotient and remainder
    __host__ __device__ __forceinline__ void divmod(uint64_t n, uint64_t &q, uint64_t &r) const {
        q = div(n);
        r = n - q * d;
    }
};

inline FastDiv make_fastdiv(uint64_t d) {
    FastDiv f;
    if (d == 0) d = 1;
    f.d = d;
    f.m = 0;
    if ((d & (d - 1)) == 0) {
        f.pow2 = 1;
        f.shift = 0;
        while ((uint64_t(1) << f.shift) < d) ++f.shift;
        return f;
    }
    f.pow2 = 0;
    const int l = 64 - __builtin_clzll(d - 1);  // ceil(log2 d) for d >= 3 that is not a power of two
    f.shift = l;
    const unsigned __int128 one = 1;
    const unsigned __int128 num = (one << 64) * ((one << l) - d);
    f.m = (uint64_t)(num / d) + 1;
    return f;
}

}  // namespace b2s

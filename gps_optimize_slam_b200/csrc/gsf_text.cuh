// Scalar text <-> double conversions of the file formats (N4), host + device:
//   parse_double   strtod-quality decimal -> double (what np.loadtxt delivers, /root/reference/EKFGPSSLAM.py:110-125, :252-258):
//                  Clinger's exact fast path, else the Eisel-Lemire algorithm with a 128-bit table of powers of five
//                  (correctly rounded for up to 19 significant digits; longer inputs are truncated and re-checked with
//                  the next significand, and flagged when the two disagree);
//   format_fixed   printf("%.Df") of a double, exactly (the np.savetxt formats of :1087-1102): integer part and the
//                  D-digit fraction from the binary expansion with 128-bit integer arithmetic, round-half-even.
#pragma once
#include <stdint.h>
#include "gsf_common.cuh"

namespace gsf {

#ifdef __CUDACC__
__device__ __constant__
#endif
static const uint64_t POW5_128[2 * 651] = {
#include "gsf_pow5_table.inc"
};

GSF_HD inline void mul64(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
#ifdef __CUDA_ARCH__
    lo = a * b; hi = __umul64hi(a, b);
#else
    const unsigned __int128 p = (unsigned __int128)a * b;
    lo = (uint64_t)p; hi = (uint64_t)(p >> 64);
#endif
}
GSF_HD inline int clz64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}
GSF_HD inline double bits_to_double(uint64_t b) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double d; __builtin_memcpy(&d, &b, 8); return d;
#endif
}
GSF_HD inline uint64_t double_to_bits(double d) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b; __builtin_memcpy(&b, &d, 8); return b;
#endif
}

// w * 10^q -> nearest double (Eisel-Lemire); w != 0.
GSF_HD inline double eisel_lemire(uint64_t w, int q) {
    if (q < -342) return 0.0;
    if (q > 308) return bits_to_double(0x7ff0000000000000ull);
    const int lz = clz64(w);
    w <<= lz;
    const uint64_t* T = POW5_128 + 2 * (q + 342);
    uint64_t hi, lo;
    mul64(w, T[0], hi, lo);
    if ((hi & 0x1FFull) == 0x1FFull) {
        uint64_t h2, l2;
        mul64(w, T[1], h2, l2);
        lo += h2;
        if (h2 > lo) ++hi;
    }
    const int upperbit = (int)(hi >> 63);
    uint64_t mant = hi >> (upperbit + 64 - 52 - 3);
    int power2 = ((((152170 + 65536) * q) >> 16) + 63) + upperbit - lz + 1023;
    if (power2 <= 0) {                                      // subnormal
        if (-power2 + 1 >= 64) return 0.0;
        mant >>= -power2 + 1;
        mant += (mant & 1); mant >>= 1;
        power2 = (mant < (1ull << 52)) ? 0 : 1;
        return bits_to_double(((uint64_t)power2 << 52) | (mant & ~(1ull << 52)));
    }
    if (lo <= 1 && q >= -4 && q <= 23 && (mant & 3) == 1) {  // exactly half-way: round to even
        if ((mant << (upperbit + 64 - 52 - 3)) == hi) mant &= ~1ull;
    }
    mant += (mant & 1); mant >>= 1;
    if (mant >= (2ull << 52)) { mant = 1ull << 52; ++power2; }
    mant &= ~(1ull << 52);
    if (power2 >= 0x7FF) return bits_to_double(0x7ff0000000000000ull);
    return bits_to_double(((uint64_t)power2 << 52) | mant);
}

// Parses one number at [p, end): optional sign, digits with an optional point, optional exponent, or nan / inf / infinity.
// Returns the number of characters consumed (0: not a number).  *inexact = 1 when more than 19 significant digits made
// the result depend on digits that were dropped (two neighbouring significands round differently): the caller reports it.
GSF_HD inline int parse_double(const char* p, const char* end, double* out, int* inexact) {
    const char* s = p;
    bool neg = false;
    if (s < end && (*s == '-' || *s == '+')) { neg = *s == '-'; ++s; }
    if (s < end && ((*s | 32) == 'n' || (*s | 32) == 'i')) {
        const char* names[3] = {"nan", "infinity", "inf"};
        for (int k = 0; k < 3; ++k) {
            int len = 0; while (names[k][len]) ++len;
            if (end - s >= len) {
                bool ok = true;
                for (int j = 0; j < len; ++j) if ((s[j] | 32) != names[k][j]) { ok = false; break; }
                if (ok) {
                    const double v = k == 0 ? bits_to_double(0x7ff8000000000000ull) : bits_to_double(0x7ff0000000000000ull);
                    *out = neg ? -v : v;
                    return (int)(s + len - p);
                }
            }
        }
        return 0;
    }
    uint64_t w = 0; int nd = 0, dropped = 0, exp10 = 0; bool any = false, dropped_nonzero = false;
    for (; s < end && *s >= '0' && *s <= '9'; ++s) {
        any = true;
        if (nd < 19) { w = w * 10 + (uint64_t)(*s - '0'); if (w) ++nd; }
        else { ++dropped; if (*s != '0') dropped_nonzero = true; }
    }
    exp10 += dropped;
    if (s < end && *s == '.') {
        ++s;
        for (; s < end && *s >= '0' && *s <= '9'; ++s) {
            any = true;
            if (nd < 19) { w = w * 10 + (uint64_t)(*s - '0'); if (w) ++nd; --exp10; }
            else if (*s != '0') dropped_nonzero = true;
        }
    }
    if (!any) return 0;
    if (s < end && (*s | 32) == 'e') {
        const char* e = s + 1;
        bool eneg = false;
        if (e < end && (*e == '-' || *e == '+')) { eneg = *e == '-'; ++e; }
        if (e < end && *e >= '0' && *e <= '9') {
            int ev = 0;
            for (; e < end && *e >= '0' && *e <= '9'; ++e) if (ev < 100000) ev = ev * 10 + (*e - '0');
            exp10 += eneg ? -ev : ev;
            s = e;
        }
    }
    double v;
    if (w == 0) v = 0.0;
    else if (w < (1ull << 53) && exp10 >= -22 && exp10 <= 22 && !dropped_nonzero) {      // Clinger: both factors exact
        static const double P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16,
                                       1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
        v = exp10 < 0 ? (double)w / P10[-exp10] : (double)w * P10[exp10];
    } else {
        v = eisel_lemire(w, exp10);
        if (dropped_nonzero && eisel_lemire(w + 1, exp10) != v) *inexact = 1;
    }
    *out = neg ? -v : v;
    return (int)(s - p);
}

// printf("%.Df", x) into buf (no terminator); returns the length.  D <= 9.  Finite |x| < 2^63, else "nan" / "inf" / "-inf" or,
// beyond the range, 0 with *bad = 1.
GSF_HD inline int format_fixed(double x, int D, char* buf, int* bad) {
    const uint64_t b = double_to_bits(x);
    const bool neg = (b >> 63) != 0;
    const int ex = (int)((b >> 52) & 0x7FF);
    const uint64_t frac = b & ((1ull << 52) - 1);
    int len = 0;
    if (ex == 0x7FF) {
        if (frac) { buf[0] = 'n'; buf[1] = 'a'; buf[2] = 'n'; return 3; }
        if (neg) buf[len++] = '-';
        buf[len++] = 'i'; buf[len++] = 'n'; buf[len++] = 'f';
        return len;
    }
    const uint64_t m = ex ? (frac | (1ull << 52)) : frac;
    const int e2 = (ex ? ex : 1) - 1075;                     // |x| = m * 2^e2
    uint64_t T = 1;
    for (int k = 0; k < D; ++k) T *= 10;
    uint64_t I, Q;
    if (e2 >= 0) {
        if (e2 > 10) { *bad = 1; return 0; }                 // >= 2^63
        I = m << e2; Q = 0;
    } else {
        const int s = -e2;
        uint64_t f;
        if (s >= 64) { I = 0; f = m; } else { I = m >> s; f = m & ((1ull << s) - 1); }
        if (s > 64 + 30) Q = 0;                              // |fraction| * 10^D < 2^(53 + 30 - s) < 1/2: rounds to zero
        else {
            uint64_t hi, lo;
            mul64(f, T, hi, lo);                             // P = f * 10^D < 2^83
            uint64_t rem_hi, rem_lo, half_hi, half_lo;
            if (s < 64) { Q = (lo >> s) | (s ? (hi << (64 - s)) : 0); rem_hi = 0; rem_lo = lo & ((1ull << s) - 1); half_hi = 0; half_lo = 1ull << (s - 1); }
            else if (s == 64) { Q = hi; rem_hi = 0; rem_lo = lo; half_hi = 0; half_lo = 1ull << 63; }
            else { Q = hi >> (s - 64); rem_hi = hi & ((1ull << (s - 64)) - 1); rem_lo = lo; half_hi = 1ull << (s - 65); half_lo = 0; }
            const bool gt = rem_hi > half_hi || (rem_hi == half_hi && rem_lo > half_lo);
            const bool eq = rem_hi == half_hi && rem_lo == half_lo;
            if (gt || (eq && (Q & 1))) ++Q;
            if (Q >= T) { Q -= T; ++I; }
        }
    }
    if (neg) buf[len++] = '-';
    char tmp[20]; int nd = 0;
    do { tmp[nd++] = (char)('0' + I % 10); I /= 10; } while (I);
    while (nd) buf[len++] = tmp[--nd];
    if (D > 0) {
        buf[len++] = '.';
        for (int k = D - 1; k >= 0; --k) { buf[len + k] = (char)('0' + Q % 10); Q /= 10; }
        len += D;
    }
    return len;
}

}  // namespace gsf

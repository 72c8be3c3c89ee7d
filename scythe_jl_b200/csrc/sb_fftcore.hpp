// sb_fftcore.hpp -- register-level FFT building blocks shared by the ring-FFT kernels.
#pragma once
#include "sb_internal.hpp"

namespace sb {

#define C_PI8 0.92387953251128673848   // cos(pi/8)
#define S_PI8 0.38268343236508978178   // sin(pi/8)
#define C_PI4 0.70710678118654752440   // cos(pi/4)

__device__ __forceinline__ double2 operator+(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 operator-(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cm(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cmc(double2 a, double2 b) { return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
// multiply by a compile-time constant w (forward) or conj(w) (inverse)
template <bool INV>
__device__ __forceinline__ double2 cw(double2 a, double wr, double wi) {
  return INV ? make_double2(a.x * wr + a.y * wi, a.y * wr - a.x * wi) : make_double2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
}

template <bool INV>
__device__ __forceinline__ void bf4(double2& a0, double2& a1, double2& a2, double2& a3) {
  double2 t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3, d = a1 - a3;
  double2 t3 = INV ? make_double2(-d.y, d.x) : make_double2(d.y, -d.x);
  a0 = t0 + t2; a1 = t1 + t3; a2 = t0 - t2; a3 = t1 - t3;
}

// second half of the 16-point DFT: inner twiddles, 4 x radix-4 across, natural-order output
template <bool INV>
__device__ __forceinline__ void fft16_tail(double2 (&v)[16]) {
  v[1 + 4] = cw<INV>(v[1 + 4], C_PI8, -S_PI8);    // W16^1
  v[1 + 8] = cw<INV>(v[1 + 8], C_PI4, -C_PI4);    // W16^2
  v[1 + 12] = cw<INV>(v[1 + 12], S_PI8, -C_PI8);  // W16^3
  v[2 + 4] = cw<INV>(v[2 + 4], C_PI4, -C_PI4);    // W16^2
  v[2 + 8] = cw<INV>(v[2 + 8], 0.0, -1.0);        // W16^4
  v[2 + 12] = cw<INV>(v[2 + 12], -C_PI4, -C_PI4); // W16^6
  v[3 + 4] = cw<INV>(v[3 + 4], S_PI8, -C_PI8);    // W16^3
  v[3 + 8] = cw<INV>(v[3 + 8], -C_PI4, -C_PI4);   // W16^6
  v[3 + 12] = cw<INV>(v[3 + 12], -C_PI8, S_PI8);  // W16^9
#pragma unroll
  for (int q = 0; q < 4; ++q) bf4<INV>(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);  // v[4q+s] = X[q+4s]
  double2 o[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) o[k] = v[4 * (k & 3) + (k >> 2)];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = o[k];
}

// 16-point DFT in registers, natural order in and out.  forward: W = exp(-2 pi i/16)
template <bool INV>
__device__ __forceinline__ void fft16(double2 (&v)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) bf4<INV>(v[c], v[c + 4], v[c + 8], v[c + 12]);   // v[c+4q] = t[c][q]
  fft16_tail<INV>(v);
}

// forward 16-point DFT of a sequence whose upper half v[8..15] is zero (zero-padded Bluestein input)
__device__ __forceinline__ void fft16_fwd_lo8(double2 (&v)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double2 a0 = v[c], a1 = v[c + 4];
    const double2 t3 = make_double2(a1.y, -a1.x);
    v[c] = a0 + a1; v[c + 4] = a0 + t3; v[c + 8] = a0 - a1; v[c + 12] = a0 - t3;
  }
  fft16_tail<false>(v);
}

template <bool INV>
__device__ __forceinline__ void fft8(double2* v) {
  bf4<INV>(v[0], v[2], v[4], v[6]);   // t[0][q] at v[2q]
  bf4<INV>(v[1], v[3], v[5], v[7]);   // t[1][q] at v[1+2q]
  v[3] = cw<INV>(v[3], C_PI4, -C_PI4);
  v[5] = cw<INV>(v[5], 0.0, -1.0);
  v[7] = cw<INV>(v[7], -C_PI4, -C_PI4);
  double2 o[8];
#pragma unroll
  for (int q = 0; q < 4; ++q) { o[q] = v[2 * q] + v[2 * q + 1]; o[q + 4] = v[2 * q] - v[2 * q + 1]; }
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = o[k];
}

// register-local pass of radix rf on contiguous blocks (16/rf blocks per thread)
template <bool INV>
__device__ __forceinline__ void fft_final(double2 (&v)[16], int rf) {
  if (rf == 16) {
    fft16<INV>(v);
  } else if (rf == 8) {
    fft8<INV>(&v[0]);
    fft8<INV>(&v[8]);
  } else if (rf == 4) {
#pragma unroll
    for (int b = 0; b < 4; ++b) bf4<INV>(v[4 * b], v[4 * b + 1], v[4 * b + 2], v[4 * b + 3]);
  } else {
#pragma unroll
    for (int b = 0; b < 8; ++b) { double2 a = v[2 * b], c = v[2 * b + 1]; v[2 * b] = a + c; v[2 * b + 1] = a - c; }
  }
}

__device__ __forceinline__ int padi(int i) { return i + (i >> 4); }

}  // namespace sb

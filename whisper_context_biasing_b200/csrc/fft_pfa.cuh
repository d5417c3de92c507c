// 400-point real DFT by the prime-factor (Good-Thomas) algorithm, 400 = 16 x 25 (coprime), written
// on an abstract "vector" type V so the same straight-line code runs
//   - on the device with V = two packed float32 (one value for each of two audio frames; every op
//     is one FADD2 / FMUL2 / FFMA2 instruction and every constant is an immediate), and
//   - on the host with V = double, where tests/test_fft_host.py checks it against numpy.
//
// Index maps (n = sample inside the Hann-windowed frame, k = frequency bin):
//   input   n = (25*n1 + 16*n2) mod 400          n1 in [0,16), n2 in [0,25)
//   output  k = (225*k1 + 176*k2) mod 400        k1 in [0,16), k2 in [0,25)
// so that W400^(n k) = W16^(n1 k1) * W25^(n2 k2) and NO twiddles are needed between the stages:
//   stage 1 (per n1):  Y[n1][k2] = sum_n2 x[n(n1,n2)] W25^(n2 k2)     real input -> 13 unique k2
//   stage 2 (per k2):  X[k(k1,k2)] = sum_n1 Y[n1][k2] W16^(n1 k1)
// Real input => X[400-k] = conj X[k]; the 13 stage-2 "slots" below use one k2 of every conjugate
// pair {k2, 25-k2}, which yields every bin 0..200 exactly once after folding k -> min(k, 400-k).
#pragma once

#ifdef __CUDACC__
#define WLM_HD __host__ __device__ __forceinline__
#else
#define WLM_HD inline
#endif

namespace wlm {
namespace fft {

// ---- required operations on V (overloaded for the packed device type and for double) ----------
//   vadd(a,b) = a+b      vsub(a,b) = a-b      vmul(a,b) = a*b       vfma(a,b,c) = a*b+c
//   vmulc(a,s) = a*s     vfmac(a,s,c) = a*s+c      (s: compile-time float constant)

constexpr float kC5_1 = 0.30901699437494742f;   // cos(2 pi / 5)
constexpr float kC5_2 = -0.80901699437494742f;  // cos(4 pi / 5)
constexpr float kS5_1 = 0.95105651629515357f;   // sin(2 pi / 5)
constexpr float kS5_2 = 0.58778525229247313f;   // sin(4 pi / 5)

// W25^m = cos(2 pi m/25) - i sin(2 pi m/25)
constexpr float kC25[9] = {1.0f,
                           0.96858316112863108f, 0.87630668004386358f, 0.72896862742141155f,
                           0.53582679497899666f, 0.30901699437494742f, 0.062790519529313374f,
                           -0.18738131458572463f, -0.42577929156507272f};
constexpr float kS25[9] = {0.0f,
                           0.24868988716485479f, 0.48175367410171532f, 0.68454710592868873f,
                           0.84432792550201508f, 0.95105651629515357f, 0.99802672842827156f,
                           0.98228725072868872f, 0.90482705246601958f};
// W16^m = cos(2 pi m/16) - i sin(2 pi m/16), m = 0..9
constexpr float kC16[10] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f, 0.0f,
                            -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f, -1.0f,
                            -0.92387953251128674f};
constexpr float kS16[10] = {0.0f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f, 1.0f,
                            0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f, 0.0f,
                            -0.38268343236508977f};

// The 13 k2 values handled by stage 2, in the order stage 1 emits them ("slots").
// slot 0 is k2 = 0 (purely real Y), slots 1..12 are complex.
constexpr int kNumSlots = 13;
constexpr int kSlotK2[kNumSlots] = {0, 5, 10, 1, 6, 11, 16, 21, 2, 7, 12, 17, 22};
// component index inside a stage-1 output row of 25 floats: re at kSlotComp[s], im at +1 (slot 0: re only)
constexpr int kSlotComp[kNumSlots] = {0, 1, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23};

WLM_HD constexpr int input_index(int n1, int n2) { return (25 * n1 + 16 * n2) % 400; }
WLM_HD constexpr int output_bin(int k1, int k2) {
    const int k = (225 * k1 + 176 * k2) % 400;
    return k <= 200 ? k : 400 - k;
}

// ---- 5-point kernels ---------------------------------------------------------------------------
// real input, outputs X0 (real), X1, X2 (X3 = conj X2, X4 = conj X1)
template <class V>
WLM_HD void rdft5(V x0, V x1, V x2, V x3, V x4, V& X0, V& X1r, V& X1i, V& X2r, V& X2i) {
    const V a1 = vadd(x1, x4), a2 = vadd(x2, x3), b1 = vsub(x1, x4), b2 = vsub(x2, x3);
    X0 = vadd(x0, vadd(a1, a2));
    X1r = vfmac(a2, kC5_2, vfmac(a1, kC5_1, x0));
    X2r = vfmac(a2, kC5_1, vfmac(a1, kC5_2, x0));
    X1i = vfmac(b2, -kS5_2, vmulc(b1, -kS5_1));
    X2i = vfmac(b2, kS5_1, vmulc(b1, -kS5_2));
}

// complex input, all five outputs
template <class V>
WLM_HD void cdft5(const V (&zr)[5], const V (&zi)[5], V (&Xr)[5], V (&Xi)[5]) {
    const V a1r = vadd(zr[1], zr[4]), a1i = vadd(zi[1], zi[4]);
    const V a2r = vadd(zr[2], zr[3]), a2i = vadd(zi[2], zi[3]);
    const V b1r = vsub(zr[1], zr[4]), b1i = vsub(zi[1], zi[4]);
    const V b2r = vsub(zr[2], zr[3]), b2i = vsub(zi[2], zi[3]);
    Xr[0] = vadd(zr[0], vadd(a1r, a2r));
    Xi[0] = vadd(zi[0], vadd(a1i, a2i));
    const V t1r = vfmac(a2r, kC5_2, vfmac(a1r, kC5_1, zr[0]));
    const V t1i = vfmac(a2i, kC5_2, vfmac(a1i, kC5_1, zi[0]));
    const V t2r = vfmac(a2r, kC5_1, vfmac(a1r, kC5_2, zr[0]));
    const V t2i = vfmac(a2i, kC5_1, vfmac(a1i, kC5_2, zi[0]));
    const V u1r = vfmac(b2r, kS5_2, vmulc(b1r, kS5_1)), u1i = vfmac(b2i, kS5_2, vmulc(b1i, kS5_1));
    const V u2r = vfmac(b2r, -kS5_1, vmulc(b1r, kS5_2)), u2i = vfmac(b2i, -kS5_1, vmulc(b1i, kS5_2));
    // X1 = t1 - i u1, X4 = t1 + i u1, X2 = t2 - i u2, X3 = t2 + i u2   (-i (a+ib) = b - i a)
    Xr[1] = vadd(t1r, u1i); Xi[1] = vsub(t1i, u1r);
    Xr[4] = vsub(t1r, u1i); Xi[4] = vadd(t1i, u1r);
    Xr[2] = vadd(t2r, u2i); Xi[2] = vsub(t2i, u2r);
    Xr[3] = vsub(t2r, u2i); Xi[3] = vadd(t2i, u2r);
}

// (xr + i xi) * W25^m, m compile-time
template <int M, class V>
WLM_HD void twiddle25(V& xr, V& xi) {
    const V r = vfmac(xi, kS25[M], vmulc(xr, kC25[M]));    // xr c + xi s
    const V i = vfmac(xr, -kS25[M], vmulc(xi, kC25[M]));   // xi c - xr s
    xr = r; xi = i;
}

// ---- stage 1: 25-point DFT of real input, 13 unique outputs ---------------------------------------
// y[n2], n2 = 5p + q.  out[25]: out[0] = Y[0] (real); for slot s >= 1: out[kSlotComp[s]] = Re Y[k2],
// out[kSlotComp[s]+1] = Im Y[k2] with k2 = kSlotK2[s].
template <class V>
WLM_HD void rfft25(const V (&y)[25], V (&out)[25]) {
    V A0[5], A1r[5], A1i[5], A2r[5], A2i[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) rdft5(y[q], y[5 + q], y[10 + q], y[15 + q], y[20 + q], A0[q], A1r[q], A1i[q], A2r[q], A2i[q]);
    // twiddles W25^(q r), r = 1, 2
    twiddle25<1>(A1r[1], A1i[1]); twiddle25<2>(A1r[2], A1i[2]); twiddle25<3>(A1r[3], A1i[3]); twiddle25<4>(A1r[4], A1i[4]);
    twiddle25<2>(A2r[1], A2i[1]); twiddle25<4>(A2r[2], A2i[2]); twiddle25<6>(A2r[3], A2i[3]); twiddle25<8>(A2r[4], A2i[4]);
    // r = 0: real 5-point over q -> k2 = 0, 5, 10
    rdft5(A0[0], A0[1], A0[2], A0[3], A0[4], out[0], out[1], out[2], out[3], out[4]);
    // r = 1: k2 = 1, 6, 11, 16, 21 ; r = 2: k2 = 2, 7, 12, 17, 22
    V Xr[5], Xi[5];
    cdft5(A1r, A1i, Xr, Xi);
#pragma unroll
    for (int s = 0; s < 5; ++s) { out[5 + 2 * s] = Xr[s]; out[6 + 2 * s] = Xi[s]; }
    cdft5(A2r, A2i, Xr, Xi);
#pragma unroll
    for (int s = 0; s < 5; ++s) { out[15 + 2 * s] = Xr[s]; out[16 + 2 * s] = Xi[s]; }
}

// ---- stage 2: 16-point complex DFT (radix 4 x 4) --------------------------------------------------
template <class V>
WLM_HD void dft4(V& r0, V& i0, V& r1, V& i1, V& r2, V& i2, V& r3, V& i3) {
    const V sr = vadd(r0, r2), si = vadd(i0, i2), dr = vsub(r0, r2), di = vsub(i0, i2);
    const V tr = vadd(r1, r3), ti = vadd(i1, i3), er = vsub(r1, r3), ei = vsub(i1, i3);
    r0 = vadd(sr, tr); i0 = vadd(si, ti);
    r2 = vsub(sr, tr); i2 = vsub(si, ti);
    // X1 = d - i e = (dr + ei, di - er) ; X3 = d + i e = (dr - ei, di + er)
    r1 = vadd(dr, ei); i1 = vsub(di, er);
    r3 = vsub(dr, ei); i3 = vadd(di, er);
}

template <int M, class V>
WLM_HD void twiddle16(V& xr, V& xi) {
    if (M == 0) return;
    if (M == 4) {  // * (-i)
        const V t = xr;
        xr = xi;
        xi = vmulc(t, -1.0f);
        return;
    }
    const V r = vfmac(xi, kS16[M], vmulc(xr, kC16[M]));
    const V i = vfmac(xr, -kS16[M], vmulc(xi, kC16[M]));
    xr = r; xi = i;
}

// in/out: xr[n1], xi[n1] -> Xr[k1], Xi[k1]   (n1 = 4a + b, k1 = c + 4d)
template <class V>
WLM_HD void cfft16(V (&xr)[16], V (&xi)[16]) {
    // 4-point DFTs over a for each b: elements b, b+4, b+8, b+12 -> G[b][c] left in place at 4c + b
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4(xr[b], xi[b], xr[b + 4], xi[b + 4], xr[b + 8], xi[b + 8], xr[b + 12], xi[b + 12]);
    // twiddle G[b][c] *= W16^(b c)   (stored at index 4c + b)
    twiddle16<1>(xr[4 + 1], xi[4 + 1]); twiddle16<2>(xr[4 + 2], xi[4 + 2]); twiddle16<3>(xr[4 + 3], xi[4 + 3]);
    twiddle16<2>(xr[8 + 1], xi[8 + 1]); twiddle16<4>(xr[8 + 2], xi[8 + 2]); twiddle16<6>(xr[8 + 3], xi[8 + 3]);
    twiddle16<3>(xr[12 + 1], xi[12 + 1]); twiddle16<6>(xr[12 + 2], xi[12 + 2]); twiddle16<9>(xr[12 + 3], xi[12 + 3]);
    // 4-point DFTs over b for each c: elements 4c+0..4c+3 -> X[c + 4d] left at 4c + d
#pragma unroll
    for (int c = 0; c < 4; ++c) dft4(xr[4 * c], xi[4 * c], xr[4 * c + 1], xi[4 * c + 1], xr[4 * c + 2], xi[4 * c + 2], xr[4 * c + 3], xi[4 * c + 3]);
}
// after cfft16 the value for k1 = c + 4d sits at array index 4c + d:
WLM_HD constexpr int fft16_slot_of_k1(int k1) { return 4 * (k1 & 3) + (k1 >> 2); }

WLM_HD constexpr int fft16_k1_of_pos(int i) { return (i >> 2) + 4 * (i & 3); }

// Stage 2 leaves |X|^2 of slot s, array position i, in row 16 s + i of its output buffer; the mel stage
// walks FFT bins in order, so it needs the inverse: the row that holds bin k (slot 0 holds each of its
// bins twice, k1 and 16 - k1 being conjugates: the first one is taken).
struct RowOfBin {
    short r[208];
};
WLM_HD constexpr RowOfBin make_row_of_bin() {
    RowOfBin t{};
    for (int k = 0; k < 208; ++k) t.r[k] = -1;
    for (int s = 0; s < kNumSlots; ++s)
        for (int i = 0; i < 16; ++i) {
            const int k = output_bin(fft16_k1_of_pos(i), kSlotK2[s]);
            if (t.r[k] < 0) t.r[k] = static_cast<short>(16 * s + i);
        }
    return t;
}
#ifdef __CUDACC__
__device__ constexpr RowOfBin kRowOfBin = make_row_of_bin();
#endif
WLM_HD int row_of_bin(int k) {
#ifdef __CUDA_ARCH__
    return kRowOfBin.r[k];
#else
    return make_row_of_bin().r[k];
#endif
}

}  // namespace fft
}  // namespace wlm

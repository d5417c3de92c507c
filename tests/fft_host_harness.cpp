// Host instantiation (V = double) of the device FFT templates, for tests/test_fft_host.py.
#include <cmath>

// the operations must be visible before the templates are parsed (double has no ADL)
namespace wlm { namespace fft {
inline double vadd(double a, double b) { return a + b; }
inline double vsub(double a, double b) { return a - b; }
inline double vmul(double a, double b) { return a * b; }
inline double vfma(double a, double b, double c) { return a * b + c; }
inline double vmulc(double a, float s) { return a * (double)s; }
inline double vfmac(double a, float s, double c) { return a * (double)s + c; }
}}
#include "fft_pfa.cuh"
using namespace wlm::fft;

// x[400] real -> re[201], im[201] (folded: bins > 200 are conjugated back)
extern "C" void pfa_rfft400(const double* x, double* re, double* im) {
    double Y[16][25];
    for (int n1 = 0; n1 < 16; ++n1) {
        double y[25];
        for (int n2 = 0; n2 < 25; ++n2) y[n2] = x[input_index(n1, n2)];
        rfft25<double>(y, Y[n1]);
    }
    for (int s = 0; s < kNumSlots; ++s) {
        double xr[16], xi[16];
        for (int n1 = 0; n1 < 16; ++n1) {
            xr[n1] = Y[n1][kSlotComp[s]];
            xi[n1] = s == 0 ? 0.0 : Y[n1][kSlotComp[s] + 1];
        }
        cfft16<double>(xr, xi);
        for (int k1 = 0; k1 < 16; ++k1) {
            const int k = (225 * k1 + 176 * kSlotK2[s]) % 400;
            const int idx = fft16_slot_of_k1(k1);
            if (k <= 200) { re[k] = xr[idx]; im[k] = xi[idx]; }
            else if (s != 0) { re[400 - k] = xr[idx]; im[400 - k] = -xi[idx]; }
        }
    }
}
extern "C" int pfa_output_bin(int k1, int k2) { return output_bin(k1, k2); }

// the kernel's P layout: stage 2 leaves |X|^2 of (slot, FFT16 output position) in row 16 slot + position,
// the mel stage reads bin k from row_of_bin(k)
extern "C" int pfa_row_of_bin(int k) { return row_of_bin(k); }
extern "C" int pfa_bin_of_row(int row) { return output_bin(fft16_k1_of_pos(row & 15), kSlotK2[row >> 4]); }
extern "C" int pfa_fft16_pos_of_k1(int k1) { return fft16_slot_of_k1(k1); }

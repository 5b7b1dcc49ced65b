// TEST INFRASTRUCTURE ONLY (see Makefile).  A C ABI around the reference's own f_u, diffqc.cc:95-135, compiled from
// the reference source: fu_body.inc is lines 75-135 of /root/reference/diffqc.cc, cut out at build time.
// The includes and typedefs are the ones diffqc.cc:5-19 has in scope for those lines (minus Eigen / pybind11,
// which lines 75-135 do not use); the globals are diffqc.cc:21-25 with the Eigen members left out.
#include <stdio.h>
#include <vector>
#include <math.h>
#include <cmath>
#include <complex>
#include <iostream>

typedef double Scalar;
typedef std::complex<Scalar> Complex;

std::vector<std::vector<std::vector<Scalar>>> g_channels;
Scalar g_duration;
int g_func_type; // 0: legendre. 1: b_spline

#include "fu_body.inc"

extern "C" {

// set_H's share that f_u reads (diffqc.cc:69-72): channels[h][c] = 4 doubles, chan_counts[h] channels for term h
void ref_set_channels(int n_H, const int* chan_counts, const double* channels, double duration, int func_type) {
    g_channels.clear();
    int k = 0;
    for (int h = 0; h < n_H; ++h) {
        std::vector<std::vector<Scalar>> hc;
        for (int c = 0; c < chan_counts[h]; ++c, ++k) hc.push_back(std::vector<Scalar>(channels + 4 * k, channels + 4 * k + 4));
        g_channels.push_back(hc);
    }
    g_duration = duration;
    g_func_type = func_type;
}

// vv is [2][n_param][n_basis], as diffqc.trotter receives it (diffqc.cc:178)
double ref_f_u(int h, double t, const double* vv, int n_param, int n_basis) {
    std::vector<std::vector<std::vector<Scalar>>> v(2);
    for (int a = 0; a < 2; ++a)
        for (int p = 0; p < n_param; ++p)
            v[a].push_back(std::vector<Scalar>(vv + ((size_t)a * n_param + p) * n_basis, vv + ((size_t)a * n_param + p + 1) * n_basis));
    return f_u(h, t, v);
}

double ref_my_expit(double x) { return my_expit(x); }
double ref_bspline(int b, int n_basis, double t) { return bspline(b, n_basis, t); }

}

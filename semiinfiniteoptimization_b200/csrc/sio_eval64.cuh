// sio_eval64.cuh -- float64 device functions shared by the float64 translation units (sio_fp64.cu, sio_lockstep.cu).
// Every unit that includes this is compiled with --fmad=false: the arithmetic is the plain IEEE sequence of the oracle.
#pragma once
#include "sio_common.cuh"

namespace sio {

// Appendix A.2 in float64.  p = grid-local point.  gl (optional) = gradient w.r.t. p.
__device__ __forceinline__ double eval64(const GridDev& g, const double p[3], double* gl) {
    const int dim[3] = {g.nx, g.ny, g.nz};
    double o[3], f[3], u[3];
    int i0[3];
    double dout2 = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double c = fmin(fmax(p[a], g.bmin[a]), g.bmax[a]);
        o[a] = p[a] - c;
        dout2 += o[a] * o[a];
        double h = (g.bmax[a] - g.bmin[a]) / (double)dim[a];
        u[a] = (c - g.bmin[a]) / h - 0.5;
        double uc = fmin(fmax(u[a], 0.0), (double)(dim[a] - 1));
        double fl = floor(uc);
        i0[a] = (int)fl;
        f[a] = uc - fl;
    }
    const float* b0 = g.v + ((size_t)i0[0] * g.sx + (size_t)i0[1] * g.sy + i0[2]);
    const float* b1 = b0 + g.sy;
    const float* b2 = b0 + g.sx;
    const float* b3 = b2 + g.sy;
    double v000 = __ldg(b0), v001 = __ldg(b0 + 1), v010 = __ldg(b1), v011 = __ldg(b1 + 1);
    double v100 = __ldg(b2), v101 = __ldg(b2 + 1), v110 = __ldg(b3), v111 = __ldg(b3 + 1);
    double dz00 = v001 - v000, dz01 = v011 - v010, dz10 = v101 - v100, dz11 = v111 - v110;
    double a00 = v000 + f[2] * dz00, a01 = v010 + f[2] * dz01;
    double a10 = v100 + f[2] * dz10, a11 = v110 + f[2] * dz11;
    double dy0 = a01 - a00, dy1 = a11 - a10;
    double c0 = a00 + f[1] * dy0, c1 = a10 + f[1] * dy1;
    double dx = c1 - c0;
    double val = c0 + f[0] * dx;
    double dout = sqrt(dout2);
    if (gl) {
        double gz0 = dz00 + f[1] * (dz01 - dz00), gz1 = dz10 + f[1] * (dz11 - dz10);
        double gf[3] = {dx, dy0 + f[0] * (dy1 - dy0), gz0 + f[0] * (gz1 - gz0)};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double h = (g.bmax[a] - g.bmin[a]) / (double)dim[a];
            double ga = (u[a] > 0.0 && u[a] < (double)(dim[a] - 1)) ? gf[a] / h : 0.0;
            if (dout > 0.0) ga += o[a] / dout;
            gl[a] = ga;
        }
    }
    return val + dout;
}

__device__ __forceinline__ void xf64(const double* T, double x, double y, double z, double* p) {
    p[0] = T[0] * x + T[1] * y + T[2] * z + T[3];
    p[1] = T[4] * x + T[5] * y + T[6] * z + T[7];
    p[2] = T[8] * x + T[9] * y + T[10] * z + T[11];
}

// gradient w.r.t. the input point: R^T gl, optionally normalised
__device__ __forceinline__ void grad_to_input(const double* T, const double* gl, int normalize, double* g) {
    double gx = T[0] * gl[0] + T[4] * gl[1] + T[8] * gl[2];
    double gy = T[1] * gl[0] + T[5] * gl[1] + T[9] * gl[2];
    double gz = T[2] * gl[0] + T[6] * gl[1] + T[10] * gl[2];
    if (normalize) {
        double n = sqrt(gx * gx + gy * gy + gz * gz);
        if (n > 0.0) { gx /= n; gy /= n; gz /= n; }
    }
    g[0] = gx; g[1] = gy; g[2] = gz;
}

// ---- rigid transforms (row-major 3x4) ----------------------------------------------------------------------
__device__ __forceinline__ void invert34(const double* T, double* Ti) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) Ti[4 * r + c] = T[4 * c + r];
        Ti[4 * r + 3] = -(T[0 + r] * T[3] + T[4 + r] * T[7] + T[8 + r] * T[11]);
    }
}

__device__ __forceinline__ void mul34(const double* A, const double* B, double* C) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            C[4 * r + c] = A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c] + A[4 * r + 2] * B[8 + c];
        C[4 * r + 3] = A[4 * r] * B[3] + A[4 * r + 1] * B[7] + A[4 * r + 2] * B[11] + A[4 * r + 3];
    }
}

}  // namespace sio

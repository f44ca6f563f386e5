// pmp_tableau.cuh -- explicit Runge-Kutta tableaus as constexpr functions.
// After full unrolling every coefficient is a compile-time constant, so it becomes an FFMA
// immediate / constant-bank operand and never occupies a register.
#pragma once
#include "../../include/b200pmp.h"

namespace pmp {

template <int METHOD> struct Tab;

// classic RK4 (BASELINE.json config 0; not in diffrax)
template <> struct Tab<PMP_METHOD_RK4> {
    static constexpr int S = 4, ORDER = 4;
    static constexpr bool FSAL = false, EMBEDDED = false;
    __host__ __device__ static constexpr double a(int i, int j) {
        constexpr double A[4][4] = {{0, 0, 0, 0}, {0.5, 0, 0, 0}, {0, 0.5, 0, 0}, {0, 0, 1.0, 0}};
        return A[i][j];
    }
    __host__ __device__ static constexpr double b(int i) {
        constexpr double B[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
        return B[i];
    }
    __host__ __device__ static constexpr double berr(int) { return 0.0; }
};

// Tsitouras 5(4) -- diffrax.Tsit5, the reference's solver (pontryagin_utils.py:515,522)
template <> struct Tab<PMP_METHOD_TSIT5> {
    static constexpr int S = 7, ORDER = 5;
    static constexpr bool FSAL = true, EMBEDDED = true;
    __host__ __device__ static constexpr double a(int i, int j) {
        constexpr double A[7][7] = {
            {0, 0, 0, 0, 0, 0, 0},
            {0.161, 0, 0, 0, 0, 0, 0},
            {-0.008480655492356992, 0.3354806554923570, 0, 0, 0, 0, 0},
            {2.897153057105494, -6.359448489975075, 4.362295432869581, 0, 0, 0, 0},
            {5.325864828439259, -11.74888356406283, 7.495539342889836, -0.09249506636175525, 0, 0, 0},
            {5.861455442946420, -12.92096931784711, 8.159367898576159, -0.07158497328140100,
             -0.02826905039406838, 0, 0},
            {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,
             2.324710524099774, 0}};
        return A[i][j];
    }
    __host__ __device__ static constexpr double b(int i) { return a(6, i); }
    __host__ __device__ static constexpr double berr(int i) {  // b - bhat
        constexpr double BH[7] = {0.09468075576583945, 0.009183565540343254, 0.4877705284247616,
                                  1.234297566930479, -2.707712349983526, 1.866628418170587, 1.0 / 66.0};
        return (i < 6 ? a(6, i) : 0.0) - BH[i];
    }
};

// Dormand-Prince 5(4) (BASELINE.json config 1)
template <> struct Tab<PMP_METHOD_DOPRI5> {
    static constexpr int S = 7, ORDER = 5;
    static constexpr bool FSAL = true, EMBEDDED = true;
    __host__ __device__ static constexpr double a(int i, int j) {
        constexpr double A[7][7] = {
            {0, 0, 0, 0, 0, 0, 0},
            {1.0 / 5, 0, 0, 0, 0, 0, 0},
            {3.0 / 40, 9.0 / 40, 0, 0, 0, 0, 0},
            {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0, 0},
            {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0, 0},
            {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0, 0},
            {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84, 0}};
        return A[i][j];
    }
    __host__ __device__ static constexpr double b(int i) { return a(6, i); }
    __host__ __device__ static constexpr double berr(int i) {
        constexpr double BH[7] = {5179.0 / 57600, 0, 7571.0 / 16695, 393.0 / 640, -92097.0 / 339200,
                                  187.0 / 2100, 1.0 / 40};
        return (i < 6 ? a(6, i) : 0.0) - BH[i];
    }
};

}  // namespace pmp

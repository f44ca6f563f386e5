// opcount.hpp -- op-counting scalar for the ORACLE (test infrastructure, not product code).
//
// Instantiating the oracle's RHS / stepper templates on `Counted` yields the ALGORITHMIC work per
// RHS evaluation and per step attempt that bench.py divides by the kernel time:
//   add, sub, mul = 1 flop (so an FMA = 2); div, sqrt = 1; sin, cos, pow = "special", counted apart;
//   min / max / abs / negate / compare / select = 0.
#pragma once
#include <cmath>
#include <cstdint>

namespace pmpo {

struct OpCounts {
    int64_t flops = 0;    // add + sub + mul + div + sqrt
    int64_t special = 0;  // sin, cos, pow
};
inline OpCounts& op_counts() {
    static thread_local OpCounts c;
    return c;
}

struct Counted {
    double v;
    Counted() : v(0) {}
    Counted(double x) : v(x) {}
    explicit Counted(float x) : v(x) {}
    explicit Counted(int x) : v(x) {}
};
inline Counted operator+(Counted a, Counted b) { op_counts().flops++; return Counted(a.v + b.v); }
inline Counted operator-(Counted a, Counted b) { op_counts().flops++; return Counted(a.v - b.v); }
inline Counted operator*(Counted a, Counted b) { op_counts().flops++; return Counted(a.v * b.v); }
inline Counted operator/(Counted a, Counted b) { op_counts().flops++; return Counted(a.v / b.v); }
inline Counted operator-(Counted a) { return Counted(-a.v); }
inline Counted r_sin(Counted x) { op_counts().special++; return Counted(std::sin(x.v)); }
inline Counted r_cos(Counted x) { op_counts().special++; return Counted(std::cos(x.v)); }
inline Counted r_pow(Counted x, Counted y) { op_counts().special++; return Counted(std::pow(x.v, y.v)); }
inline Counted r_sqrt(Counted x) { op_counts().flops++; return Counted(std::sqrt(x.v)); }
inline Counted r_abs(Counted x) { return Counted(std::fabs(x.v)); }
inline bool r_isnan(Counted x) { return std::isnan(x.v); }
inline double r_val(Counted x) { return x.v; }

}  // namespace pmpo

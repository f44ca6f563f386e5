// dual.hpp -- forward-mode dual numbers for the ORACLE (test infrastructure, not product code).
//
// Instantiating the oracle's templates on `Dual` gives the derivative of a whole solve along one direction of its
// terminal state -- what jax.jacobian / jvp through diffrax.diffeqsolve computes for experiment_orbits.py:166-197
// (dsol_final = jax.jacobian(get_final_x)(theta)): every arithmetic operation carries its tangent, every DECISION
// (step-size control, clipping, argmin over candidates) looks at the value only, exactly like jax's jvp of the
// traced program (the controller's own sensitivity to the tangent is zero: it is a function of primal values).
#pragma once
#include <cmath>

namespace pmpo {

struct Dual {
    double v, d;
    Dual() : v(0), d(0) {}
    Dual(double x) : v(x), d(0) {}
    Dual(double x, double dx) : v(x), d(dx) {}
    explicit Dual(float x) : v(x), d(0) {}
    explicit Dual(int x) : v(x), d(0) {}
};
inline Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
inline Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
inline Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
inline Dual operator/(Dual a, Dual b) { const double q = a.v / b.v; return Dual(q, (a.d - q * b.d) / b.v); }
inline Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
inline Dual r_sin(Dual x) { return Dual(std::sin(x.v), std::cos(x.v) * x.d); }
inline Dual r_cos(Dual x) { return Dual(std::cos(x.v), -std::sin(x.v) * x.d); }
inline Dual r_sqrt(Dual x) { const double s = std::sqrt(x.v); return Dual(s, x.d / (2 * s)); }
inline Dual r_pow(Dual x, Dual y) { return Dual(std::pow(x.v, y.v), 0.0); }  // step-size factor only: tangent unused
inline Dual r_abs(Dual x) { return x.v < 0 ? -x : x; }
inline bool r_isnan(Dual x) { return std::isnan(x.v); }
inline double r_val(Dual x) { return x.v; }

}  // namespace pmpo

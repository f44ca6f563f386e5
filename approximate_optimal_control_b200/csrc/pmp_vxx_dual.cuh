// pmp_vxx_dual.cuh -- the value-Hessian branch (PMP_FLAG_VXX) for ANY dynamics functor, by forward-mode differentiation.
//
// pontryagin_utils.py:460-467 is generic: jax.jacobian(pmp_rhs, argnums=(0, 1)) differentiates whatever f, l the user
// supplied THROUGH u_star_2d and forms  vxx' = gx + glam vxx - vxx fx - vxx flam vxx.  The hand-written flatquad functor
// carries those Jacobians in closed form (pmp_systems.cuh); a generated functor (codegen.py) does not.  This adaptor gives
// every functor the interface the integrator's vxx phases expect (JacPrim / JacCoef / rhs_prim / jac_from_prim / vxx_dot)
// without ever forming the four Jacobians: column j of vxx' is a DIRECTIONAL derivative of pmp_rhs,
//     (d xdot, d lamdot) = D pmp_rhs(x, lam) [e_j, vxx e_j]      =>  d xdot   = (fx + flam vxx) e_j
//                                                                    d lamdot = (gx + glam vxx) e_j
//     vxx'[:, j] = d lamdot - vxx d xdot,
// so nx passes of the functor instantiated on dual numbers (pmp_dual.cuh, pmp_dual_math.cuh) replace 2 nx passes plus three
// nx^3 products.  Decisions inside u* (clipping, which edge candidate wins) are taken on values, as jax's autodiff does.
// The column loop is ROLLED (one copy of the dual right hand side in the instruction stream): the tangent seed vxx[:, j] is
// picked out of the register-resident stage matrix with a select chain.
#pragma once
#include "pmp_dual_math.cuh"
#include "pmp_systems.cuh"

namespace pmp {

template <template <class> class SysT, class R> struct DualVxx : SysT<R> {
    using Base = SysT<R>;
    using D = Dual<R>;
    using SysD = SysT<D>;
    static constexpr int NX = Base::NX, NU = Base::NU;
    static constexpr int NPRIM = 2 * NX;  // what a V_xx stage needs of its base stage: the stage state (x, lambda) itself
    struct Consts : Base::Consts {
        typename SysD::Consts dual;  // the same constants as dual numbers (zero tangents)
    };
    static __host__ Consts make(const pmp_problem_t& p) {
        Consts c;
        static_cast<typename Base::Consts&>(c) = Base::make(p);
        c.dual = SysD::make(p);
        return c;
    }
    struct JacPrim {
        R y[2 * NX];
        int act;
    };
    using JacCoef = JacPrim;
    static __device__ __forceinline__ void prim_store(const JacPrim& p, R* a, int& act) {
#pragma unroll
        for (int q = 0; q < 2 * NX; q++) a[q] = p.y[q];
        act = 0;
    }
    static __device__ __forceinline__ void prim_load(JacPrim& p, const R* a, int) {
#pragma unroll
        for (int q = 0; q < 2 * NX; q++) p.y[q] = a[q];
        p.act = 0;
    }
    // rhs<2> (KKT-form u*) that also keeps the point it was evaluated at
    static __device__ __forceinline__ void rhs_prim(const Consts& c, const R* x, const R* lam, R* xd, R& vd, R* ld, JacPrim& p) {
        Base::template rhs<2>(c, x, lam, xd, vd, ld);
#pragma unroll
        for (int q = 0; q < NX; q++) { p.y[q] = x[q]; p.y[NX + q] = lam[q]; }
        p.act = 0;
    }
    static __device__ __forceinline__ void jac_from_prim(const Consts&, const JacPrim& p, JacCoef& j) { j = p; }
    static __device__ __forceinline__ void rhs_jac(const Consts& c, const R* x, const R* lam, R* xd, R& vd, R* ld, JacCoef& j) {
        rhs_prim(c, x, lam, xd, vd, ld, j);
    }
    // vxx' for V stored as VxxMap<NX, VXX> says; every entry is handed to put(storage entry, value) once
    template <int VXX, class Put>
    static __device__ __forceinline__ void vxx_dot(const Consts& c, const JacCoef& j, const R* V, Put&& put) {
        using MP = VxxMap<NX, VXX>;
        using M = Math<R>;
#pragma unroll 1
        for (int col = 0; col < NX; col++) {
            D x[NX], lam[NX], xd[NX], ld[NX], vd;
#pragma unroll
            for (int q = 0; q < NX; q++) {
                R seed = V[MP::ent(q, 0)];
#pragma unroll
                for (int m = 1; m < NX; m++) seed = (col == m) ? V[MP::ent(q, m)] : seed;
                x[q] = D(j.y[q], col == q ? R(1) : R(0));
                lam[q] = D(j.y[NX + q], seed);
            }
            SysD::template rhs<2>(c.dual, x, lam, xd, vd, ld);
#pragma unroll
            for (int i = 0; i < NX; i++) {
                if (VXX == 2 && i > col) continue;  // symmetric storage: the upper triangle only
                R a = ld[i].d;
#pragma unroll
                for (int k = 0; k < NX; k++) a = M::fma(-V[MP::ent(i, k)], xd[k].d, a);
                put(MP::ent(i, col), a);
            }
        }
    }
};

}  // namespace pmp

// pmp_capi.cu -- extern "C" surface of libb200pmp.so (declared in include/b200pmp.h).
#include <cuda_runtime.h>

#include <cstdio>
#include <string>
#include <vector>

#include "../../include/b200pmp.h"
#include "pmp_hostq.h"

namespace pmp {
struct PushArgs {  // (pmp_solve.cuh)
    void* peer[7];
    int n_peers;
    void* mc;
};
#ifndef PMP_PLUGIN
int solve_flatquad_push(const pmp_problem_t&, const pmp_opts_t&, long long, const void*, void*, int32_t*, int32_t*, int32_t*, void*,
                        const PushArgs&, cudaStream_t);
int solve_orbits_push(const pmp_problem_t&, const pmp_opts_t&, long long, const void*, void*, int32_t*, int32_t*, int32_t*, void*,
                      const PushArgs&, cudaStream_t);
int solve_lqr_push(const pmp_problem_t&, const pmp_opts_t&, long long, const void*, void*, int32_t*, int32_t*, int32_t*, void*,
                   const PushArgs&, cudaStream_t);
#else
int solve_generated_push(const pmp_problem_t&, const pmp_opts_t&, long long, const void*, void*, int32_t*, int32_t*, int32_t*, void*,
                         const PushArgs&, cudaStream_t);
#endif
#ifndef PMP_PLUGIN
int solve_flatquad(const pmp_problem_t&, const pmp_opts_t&, long long, const void*, void*, const pmp_dense_t*,
                   int32_t*, int32_t*, int32_t*, void*, cudaStream_t);
int solve_flatquad_vxx(const pmp_problem_t&, const pmp_opts_t&, long long, const void*, void*, const pmp_dense_t*,
                       int32_t*, int32_t*, int32_t*, void*, cudaStream_t);
int solve_orbits(const pmp_problem_t&, const pmp_opts_t&, long long, const void*, void*, const pmp_dense_t*,
                 int32_t*, int32_t*, int32_t*, void*, cudaStream_t);
int solve_lqr(const pmp_problem_t&, const pmp_opts_t&, long long, const void*, void*, const pmp_dense_t*,
              int32_t*, int32_t*, int32_t*, void*, cudaStream_t);
#endif
#ifdef PMP_PLUGIN
int generated_nx();
int generated_nu();
int solve_generated(const pmp_problem_t&, const pmp_opts_t&, long long, const void*, void*, const pmp_dense_t*,
                    int32_t*, int32_t*, int32_t*, void*, cudaStream_t);
int solve_generated_vxx(const pmp_problem_t&, const pmp_opts_t&, long long, const void*, void*, const pmp_dense_t*,
                        int32_t*, int32_t*, int32_t*, void*, cudaStream_t);
#endif
int eval_batch(int what, const pmp_problem_t& p, int dtype, long long n, const void* a, const void* b, void* out,
               cudaStream_t st);
int forward_batch(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* x0, const double* P,
                  const double* x_eq, const pmp_nn_costate_t* nn, void* z_end, void* ts, void* xs, void* cost, int32_t* a,
                  int32_t* s, int32_t* r, cudaStream_t st);
int nn_value_batch(const pmp_problem_t& p, const pmp_nn_costate_t* nn, int dtype, long long n, const void* x, void* lam,
                   void* vm, void* vs, cudaStream_t st);
int solve_host(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* x_f, const void* t_f, const void* v_f,
               const void* vx_f, void* x_end, void* t_end, void* v_end, void* vx_end, int32_t* n_accepted, int32_t* n_steps,
               int32_t* result, int n_pieces, const double* term_P, const double* term_xeq);
int host_plan(long long n, int n_pieces, std::vector<long long>* bounds, long long* cs_max);
int host_release();
int comm_unique_id(void* id128, std::string* msg);
int comm_init(void** comm, int world, int rank, const void* id128, std::string* msg);
int comm_destroy(void* comm, std::string* msg);
int allgather(void* comm, const void* local, void* global, size_t count, int dtype, cudaStream_t st, std::string* msg);
int solve_jvp(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, const void* dy_f, void* y_end, void* dy_end,
              int32_t* a, int32_t* s, int32_t* r, cudaStream_t st);
int select_count(int dtype, long long n, int s1, const void* v, const void* vxx, int nvxx, const int32_t* n_accepted, double lo,
                 double hi, double vxx_max, int32_t* counts, cudaStream_t st);
int select_write(int dtype, long long n, int s1, int nx, const void* x, const void* t, const void* v, const void* vx, const void* vxx,
                 int nvxx, const int32_t* n_accepted, const long long* offsets, double lo, double hi, double vxx_max, void* ox, void* ot,
                 void* ov, void* ovx, void* ovxx, double* stats, int32_t* flags, cudaStream_t st);
int host_probe_copy(void* dst, const void* src, long long bytes, int blocks, cudaStream_t st);
int fma_peak(int dtype, int variant, int iters, void* scratch, double* flops_out, cudaStream_t st);
int interp_batch(const pmp_problem_t& p, const pmp_opts_t& o, long long nq, const void* y0, const void* t1,
                 const void* target, void* y_out, int mode, cudaStream_t st);
}  // namespace pmp

namespace {
thread_local std::string g_err = "";

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int cuda_fail(int e) {
    return fail(PMP_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString((cudaError_t)e));
}

#ifdef PMP_PLUGIN
int expected_nx(int) { return pmp::generated_nx(); }
int expected_nu(int) { return pmp::generated_nu(); }
#else
int expected_nx(int sys) { return sys == PMP_SYS_FLATQUAD ? 6 : 2; }
int expected_nu(int sys) { return sys == PMP_SYS_FLATQUAD ? 2 : 1; }
#endif

int check_problem(const pmp_problem_t* p) {
    if (!p) return fail(PMP_ERR_BAD_ARG, "problem is NULL");
#ifdef PMP_PLUGIN
    if (p->system_id != PMP_SYS_GENERATED)
        return fail(PMP_ERR_UNSUPPORTED, "this library is a codegen plug-in: it serves system_id PMP_SYS_GENERATED only");
#else
    if (p->system_id < 0 || p->system_id > PMP_SYS_LQR_STATEPENALTY)
        return fail(PMP_ERR_UNSUPPORTED, "unknown system_id " + std::to_string(p->system_id) +
                                             " (compiled systems: 0 flatquad, 1 orbits, 2 lqr_statepenalty; "
                                             "PMP_SYS_GENERATED lives in the plug-in library codegen.py builds)");
#endif
    if (p->nx != expected_nx(p->system_id) || p->nu != expected_nu(p->system_id))
        return fail(PMP_ERR_BAD_ARG, "nx/nu do not match system_id");
    for (int i = 0; i < p->nu; i++)
        if (!(p->u_lo[i] < p->u_hi[i])) return fail(PMP_ERR_BAD_ARG, "U_interval: need u_lo < u_hi");
    return 0;
}

int check_opts(const pmp_opts_t* o) {
    if (!o) return fail(PMP_ERR_BAD_ARG, "opts is NULL");
    if (o->method < PMP_METHOD_RK4 || o->method > PMP_METHOD_DOPRI5) return fail(PMP_ERR_UNSUPPORTED, "unknown method");
    if (o->dtype != PMP_F32 && o->dtype != PMP_F64) return fail(PMP_ERR_BAD_ARG, "dtype must be 0 (f32) or 1 (f64)");
    if (o->indep != PMP_INDEP_TIME && o->indep != PMP_INDEP_VALUE) return fail(PMP_ERR_BAD_ARG, "bad indep");
    if (o->max_steps < 1) return fail(PMP_ERR_BAD_ARG, "max_steps must be >= 1");
    if (o->event < PMP_EVENT_NONE || o->event > PMP_EVENT_NN_VALUE_UCB_LE) return fail(PMP_ERR_UNSUPPORTED, "unknown event");
    if (o->event == PMP_EVENT_VXX_NORM && !(o->flags & PMP_FLAG_VXX))
        return fail(PMP_ERR_BAD_ARG, "PMP_EVENT_VXX_NORM needs PMP_FLAG_VXX");
    if (o->adaptive && o->method == PMP_METHOD_RK4)
        return fail(PMP_ERR_UNSUPPORTED, "RK4 has no embedded error estimate: adaptive must be 0");
    if (!(o->dt0 != 0.0) || o->dt0 != o->dt0) return fail(PMP_ERR_BAD_ARG, "dt0 must be non-zero");
    if (o->indep == PMP_INDEP_TIME) {
        if (!(o->t0 != o->t1)) return fail(PMP_ERR_BAD_ARG, "t0 == t1");
        if ((o->t1 - o->t0) * o->dt0 < 0) return fail(PMP_ERR_BAD_ARG, "dt0 must point from t0 to t1");
    } else if (!(o->dt0 > 0)) {
        return fail(PMP_ERR_BAD_ARG, "value-parameterised solve needs dt0 > 0");
    }
    if ((o->flags & ~(PMP_FLAG_USTAR_LITERAL | PMP_FLAG_VXX | PMP_FLAG_VXX_SYMMETRIC)) != 0 || o->reserved != 0)
        return fail(PMP_ERR_BAD_ARG, "unknown flags / reserved != 0");
    if ((o->flags & PMP_FLAG_VXX_SYMMETRIC) && !(o->flags & PMP_FLAG_VXX))
        return fail(PMP_ERR_BAD_ARG, "PMP_FLAG_VXX_SYMMETRIC needs PMP_FLAG_VXX");
    if (o->flags & PMP_FLAG_VXX) {
        if (o->indep != PMP_INDEP_TIME) return fail(PMP_ERR_UNSUPPORTED, "PMP_FLAG_VXX: time-parameterised solves only");
        if (o->flags & PMP_FLAG_USTAR_LITERAL)
            return fail(PMP_ERR_UNSUPPORTED, "PMP_FLAG_VXX differentiates the KKT form of u*: clear PMP_FLAG_USTAR_LITERAL");
    }
    if (o->adaptive) {
        if (!(o->rtol >= 0) || !(o->atol >= 0) || !(o->rtol + o->atol > 0)) return fail(PMP_ERR_BAD_ARG, "bad rtol/atol");
        if (!(o->safety > 0) || !(o->factormin > 0) || !(o->factormax >= 1)) return fail(PMP_ERR_BAD_ARG, "bad controller factors");
    }
    return 0;
}
// PMP_FLAG_VXX: the hand-written flatquad functor carries closed-form Jacobians through u*; a generated functor is
// differentiated in forward mode (pmp_vxx_dual.cuh); V_xx and its stage derivatives must fit one CTA's shared memory.
// orbits / LQR run through u_star_new in the reference, which has no vxx branch.
int check_vxx_system(const pmp_problem_t* p, int32_t flags) {
    if (!(flags & PMP_FLAG_VXX)) return 0;
    if (p->system_id == PMP_SYS_GENERATED) {
        if (p->nx > 9) return fail(PMP_ERR_UNSUPPORTED, "PMP_FLAG_VXX on a generated system: nx <= 9 (V_xx and its stage derivatives live in shared memory)");
        return 0;
    }
    if (p->system_id != PMP_SYS_FLATQUAD)
        return fail(PMP_ERR_UNSUPPORTED, "PMP_FLAG_VXX: flatquad and generated systems only (the reference's vxx branch runs through u_star_2d)");
    return 0;
}
// options that need a particular system
int check_pair(const pmp_problem_t* p, const pmp_opts_t* o) {
    if (o->event == PMP_EVENT_LQR_VALUE_LE || o->event == PMP_EVENT_NN_VALUE_UCB_LE)
        return fail(PMP_ERR_BAD_ARG, "this event belongs to pmp_forward_batch_cuda / pmp_forward_nn_batch_cuda (the backward "
                                     "solve has no costate law)");
    if (int e = check_vxx_system(p, o->flags)) return e;
    return 0;
}
int64_t vxx_rows(const pmp_problem_t* p, int32_t flags) { return (flags & PMP_FLAG_VXX) ? (int64_t)p->nx * p->nx : 0; }
}  // namespace

namespace pmp {
#ifdef PMP_PLUGIN
int solve_generated_hostq(const pmp_problem_t&, const pmp_opts_t&, long long, const HostQueueRaw&, cudaStream_t);
#else
int solve_flatquad_hostq(const pmp_problem_t&, const pmp_opts_t&, long long, const HostQueueRaw&, cudaStream_t);
int solve_orbits_hostq(const pmp_problem_t&, const pmp_opts_t&, long long, const HostQueueRaw&, cudaStream_t);
int solve_lqr_hostq(const pmp_problem_t&, const pmp_opts_t&, long long, const HostQueueRaw&, cudaStream_t);
#endif
int solve_hostq_dispatch(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const HostQueueRaw& q, cudaStream_t st) {
    if (o.indep != PMP_INDEP_TIME || o.save_dense || o.flags != 0) return -1000;
#ifdef PMP_PLUGIN
    return solve_generated_hostq(p, o, n, q, st);
#else
    switch (p.system_id) {
        case PMP_SYS_FLATQUAD: return solve_flatquad_hostq(p, o, n, q, st);
        case PMP_SYS_ORBITS: return solve_orbits_hostq(p, o, n, q, st);
        case PMP_SYS_LQR_STATEPENALTY: return solve_lqr_hostq(p, o, n, q, st);
    }
    return -1000;
#endif
}
}  // namespace pmp

extern "C" {

int pmp_abi_version(void) { return PMP_ABI_VERSION; }
const char* pmp_last_error(void) { return g_err.c_str(); }

int pmp_state_rows(const pmp_problem_t* p) {
    if (int e = check_problem(p)) return e;
    return 2 * p->nx + 2;
}

int pmp_state_rows_flags(const pmp_problem_t* p, int32_t flags) {
    if (int e = check_problem(p)) return e;
    if (int e = check_vxx_system(p, flags)) return e;
    return 2 * p->nx + 2 + (int)vxx_rows(p, flags);
}

int pmp_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) return cuda_fail((int)e);
    return n;
}

int64_t pmp_solve_workspace_bytes(const pmp_problem_t* p, const pmp_opts_t* o, int64_t n) {
    if (int e = check_problem(p)) return e;
    if (int e = check_opts(o)) return e;
    if (int e = check_pair(p, o)) return e;
    if (n < 0) return fail(PMP_ERR_BAD_ARG, "n < 0");
    const int64_t nd = 2 * p->nx + 1 + vxx_rows(p, o->flags);
    return 256 + nd * n * (o->dtype == PMP_F32 ? 4 : 8);
}

int pmp_solve_launch_count(const pmp_problem_t* p, const pmp_opts_t* o, int64_t n) {
    if (int e = check_problem(p)) return e;
    if (int e = check_opts(o)) return e;
    if (int e = check_pair(p, o)) return e;
    return n > 0 ? 2 : 0;  // prep (k1 / counter reset) + integrator
}

int pmp_solve_batch_cuda(const pmp_problem_t* p, const pmp_opts_t* o, int64_t n, const void* y_f, void* y_end,
                         const pmp_dense_t* dense, int32_t* n_accepted, int32_t* n_steps, int32_t* result,
                         void* workspace, int64_t workspace_bytes, void* cuda_stream) {
    if (int e = check_problem(p)) return e;
    if (int e = check_opts(o)) return e;
    if (int e = check_pair(p, o)) return e;
    if (n < 0) return fail(PMP_ERR_BAD_ARG, "n < 0");
    if (n == 0) return PMP_OK;
    if (!y_f || !y_end) return fail(PMP_ERR_BAD_ARG, "y_f / y_end is NULL");
    if (o->save_dense && !dense) return fail(PMP_ERR_BAD_ARG, "save_dense set but dense is NULL");
    if (o->save_dense) {
        const void* ptrs[5] = {dense->ts, dense->x, dense->t, dense->v, dense->vx};
        for (const void* q : ptrs)
            if (!q || ((uintptr_t)q & 15u) != 0)
                return fail(PMP_ERR_BAD_ARG, "dense output: all five leaves are required and must be 16-byte aligned");
        if ((o->flags & PMP_FLAG_VXX) && (!dense->vxx || ((uintptr_t)dense->vxx & 15u) != 0))
            return fail(PMP_ERR_BAD_ARG, "dense output with PMP_FLAG_VXX: the vxx leaf is required and must be 16-byte aligned");
        // slot indices travel as 32-bit integers inside the staging buffer
        if ((double)n * ((double)o->max_steps + 1.0) >= 4294967295.0)
            return fail(PMP_ERR_UNSUPPORTED, "dense output: n * (max_steps + 1) must be < 2^32; split the batch");
    }
    if (!workspace || workspace_bytes < pmp_solve_workspace_bytes(p, o, n))
        return fail(PMP_ERR_BAD_ARG, "workspace is NULL or smaller than pmp_solve_workspace_bytes()");
    if (((uintptr_t)workspace & 255u) != 0) return fail(PMP_ERR_BAD_ARG, "workspace must be 256-byte aligned");
    int ndev = pmp_device_count();
    if (ndev <= 0) return fail(PMP_ERR_NO_DEVICE, "no CUDA device: libb200pmp has no CPU fallback");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc = -1000;
#ifdef PMP_PLUGIN
    if (o->flags & PMP_FLAG_VXX) rc = pmp::solve_generated_vxx(*p, *o, n, y_f, y_end, dense, n_accepted, n_steps, result, workspace, st);
    else rc = pmp::solve_generated(*p, *o, n, y_f, y_end, dense, n_accepted, n_steps, result, workspace, st);
#else
    switch (p->system_id) {
        case PMP_SYS_FLATQUAD:
            if (o->flags & PMP_FLAG_VXX) rc = pmp::solve_flatquad_vxx(*p, *o, n, y_f, y_end, dense, n_accepted, n_steps, result, workspace, st);
            else rc = pmp::solve_flatquad(*p, *o, n, y_f, y_end, dense, n_accepted, n_steps, result, workspace, st);
            break;
        case PMP_SYS_ORBITS: rc = pmp::solve_orbits(*p, *o, n, y_f, y_end, dense, n_accepted, n_steps, result, workspace, st); break;
        case PMP_SYS_LQR_STATEPENALTY: rc = pmp::solve_lqr(*p, *o, n, y_f, y_end, dense, n_accepted, n_steps, result, workspace, st); break;
    }
#endif
    if (rc == -1000) return fail(PMP_ERR_UNSUPPORTED, "no kernel for this (system, dtype, method) combination");
    if (rc != 0) return cuda_fail(rc);
    return PMP_OK;
}

int pmp_solve_batch_push_cuda(const pmp_problem_t* p, const pmp_opts_t* o, int64_t n, const void* y_f, void* y_end,
                              int32_t* n_accepted, int32_t* n_steps, int32_t* result, void* workspace, int64_t workspace_bytes,
                              void* const* peer_slots, int32_t n_peers, void* multicast_slot, void* cuda_stream) {
    if (int e = check_problem(p)) return e;
    if (int e = check_opts(o)) return e;
    if (int e = check_pair(p, o)) return e;
    if (o->save_dense || o->flags != 0 || o->indep != PMP_INDEP_TIME)
        return fail(PMP_ERR_UNSUPPORTED, "pmp_solve_batch_push_cuda: time-parameterised endpoint solves with the default u* selection");
    if (n < 0) return fail(PMP_ERR_BAD_ARG, "n < 0");
    if (n_peers < 0 || n_peers > 7 || (n_peers > 0 && !peer_slots)) return fail(PMP_ERR_BAD_ARG, "n_peers must be 0..7 with peer_slots set");
    if (n == 0) return PMP_OK;
    if (!y_f || !y_end) return fail(PMP_ERR_BAD_ARG, "y_f / y_end is NULL");
    if (!workspace || workspace_bytes < pmp_solve_workspace_bytes(p, o, n))
        return fail(PMP_ERR_BAD_ARG, "workspace is NULL or smaller than pmp_solve_workspace_bytes()");
    if (((uintptr_t)workspace & 255u) != 0) return fail(PMP_ERR_BAD_ARG, "workspace must be 256-byte aligned");
    if (pmp_device_count() <= 0) return fail(PMP_ERR_NO_DEVICE, "no CUDA device: libb200pmp has no CPU fallback");
    pmp::PushArgs push = {};
    for (int q = 0; q < n_peers; q++) {
        if (!peer_slots[q]) return fail(PMP_ERR_BAD_ARG, "peer_slots holds a NULL pointer");
        push.peer[q] = peer_slots[q];
    }
    push.n_peers = n_peers;
    push.mc = multicast_slot;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc = -1000;
#ifdef PMP_PLUGIN
    rc = pmp::solve_generated_push(*p, *o, n, y_f, y_end, n_accepted, n_steps, result, workspace, push, st);
#else
    switch (p->system_id) {
        case PMP_SYS_FLATQUAD: rc = pmp::solve_flatquad_push(*p, *o, n, y_f, y_end, n_accepted, n_steps, result, workspace, push, st); break;
        case PMP_SYS_ORBITS: rc = pmp::solve_orbits_push(*p, *o, n, y_f, y_end, n_accepted, n_steps, result, workspace, push, st); break;
        case PMP_SYS_LQR_STATEPENALTY: rc = pmp::solve_lqr_push(*p, *o, n, y_f, y_end, n_accepted, n_steps, result, workspace, push, st); break;
    }
#endif
    if (rc == -1000) return fail(PMP_ERR_UNSUPPORTED, "no kernel for this (system, dtype, method) combination");
    if (rc != 0) return cuda_fail(rc);
    return PMP_OK;
}

static int check_host(const pmp_problem_t* p, const pmp_opts_t* o, int64_t n, int32_t n_pieces) {
    if (int e = check_problem(p)) return e;
    if (int e = check_opts(o)) return e;
    if (int e = check_pair(p, o)) return e;
    if (o->save_dense || (o->flags & PMP_FLAG_VXX))
        return fail(PMP_ERR_UNSUPPORTED, "pmp_solve_host returns endpoints only: save_dense / PMP_FLAG_VXX go through "
                                         "pmp_solve_batch_cuda with device buffers");
    if (n < 0) return fail(PMP_ERR_BAD_ARG, "n < 0");
    if (n_pieces < 0 || n_pieces > 1024) return fail(PMP_ERR_BAD_ARG, "n_pieces must be in 0..1024");
    return 0;
}

int pmp_solve_jvp_cuda(const pmp_problem_t* p, const pmp_opts_t* o, int64_t n, const void* y_f, const void* dy_f, void* y_end,
                       void* dy_end, int32_t* n_accepted, int32_t* n_steps, int32_t* result, void* cuda_stream) {
    if (int e = check_problem(p)) return e;
    if (int e = check_opts(o)) return e;
    if (int e = check_pair(p, o)) return e;
    if (o->save_dense || o->flags != 0) return fail(PMP_ERR_UNSUPPORTED, "pmp_solve_jvp_cuda: endpoints only, default flags");
    if (n < 0) return fail(PMP_ERR_BAD_ARG, "n < 0");
    if (n == 0) return PMP_OK;
    if (!y_f || !dy_f || !y_end || !dy_end) return fail(PMP_ERR_BAD_ARG, "y_f / dy_f / y_end / dy_end is NULL");
    if (pmp_device_count() <= 0) return fail(PMP_ERR_NO_DEVICE, "no CUDA device: libb200pmp has no CPU fallback");
    int rc = pmp::solve_jvp(*p, *o, n, y_f, dy_f, y_end, dy_end, n_accepted, n_steps, result, (cudaStream_t)cuda_stream);
    if (rc == -1000) return fail(PMP_ERR_UNSUPPORTED, "pmp_solve_jvp_cuda: no instantiation for this system / method / dtype");
    return rc ? cuda_fail(rc) : PMP_OK;
}

int pmp_solve_host(const pmp_problem_t* p, const pmp_opts_t* o, int64_t n, const void* x_f, const void* t_f, const void* v_f,
                   const void* vx_f, void* x_end, void* t_end, void* v_end, void* vx_end, int32_t* n_accepted,
                   int32_t* n_steps, int32_t* result, int32_t n_pieces) {
    if (int e = check_host(p, o, n, n_pieces)) return e;
    if (n == 0) return PMP_OK;
    if (!x_f || !v_f || !vx_f) return fail(PMP_ERR_BAD_ARG, "x_f / v_f / vx_f is NULL");
    if (pmp_device_count() <= 0) return fail(PMP_ERR_NO_DEVICE, "no CUDA device: libb200pmp has no CPU fallback");
    int rc = pmp::solve_host(*p, *o, n, x_f, t_f, v_f, vx_f, x_end, t_end, v_end, vx_end, n_accepted, n_steps, result, n_pieces,
                             nullptr, nullptr);
    if (rc > 0) return cuda_fail(rc);
    return rc;  // < 0: recorded by pmp_solve_batch_cuda
}

int pmp_solve_host_lqr(const pmp_problem_t* p, const pmp_opts_t* o, int64_t n, const void* x_f, const void* t_f,
                       const double* P_lqr, const double* x_eq, void* x_end, void* t_end, void* v_end, void* vx_end,
                       int32_t* n_accepted, int32_t* n_steps, int32_t* result, int32_t n_pieces) {
    if (int e = check_host(p, o, n, n_pieces)) return e;
    if (n == 0) return PMP_OK;
    if (!x_f || !P_lqr) return fail(PMP_ERR_BAD_ARG, "x_f / P_lqr is NULL");
    if (p->nx > 16) return fail(PMP_ERR_UNSUPPORTED, "pmp_solve_host_lqr: nx <= 16");
    if (pmp_device_count() <= 0) return fail(PMP_ERR_NO_DEVICE, "no CUDA device: libb200pmp has no CPU fallback");
    int rc = pmp::solve_host(*p, *o, n, x_f, t_f, nullptr, nullptr, x_end, t_end, v_end, vx_end, n_accepted, n_steps, result, n_pieces,
                             P_lqr, x_eq);
    if (rc > 0) return cuda_fail(rc);
    return rc;
}

int pmp_solve_host_plan(const pmp_problem_t* p, const pmp_opts_t* o, int64_t n, int32_t n_pieces, int32_t* kernels,
                        int32_t* copies) {
    if (int e = check_host(p, o, n, n_pieces)) return e;
    const int pieces = n > 0 ? pmp::host_plan(n, n_pieces, nullptr, nullptr) : 0;
    // n >= 2^16 with default flags: ONE persistent integrator launch (no pack / prep / unpack kernels); else per piece
    const bool persistent = n >= (1 << 16) && o->indep == PMP_INDEP_TIME && o->flags == 0 && pieces <= 32;
    if (kernels) *kernels = persistent ? 1 : pieces * (2 + pmp_solve_launch_count(p, o, 1));  // pack + solve + unpack
    // leaves in; leaves + counters out (all outputs requested, t_f given; the counters travel as one packed word when
    // max_steps < 2^14)
    // (+ the 8-byte frontier copy behind every piece of a persistent launch)
    if (copies) *copies = pieces * (4 + (o->max_steps < (1 << 14) ? 5 : 7) + (persistent ? 1 : 0));
    return pieces;
}

int pmp_select_count_cuda(int32_t dtype, int64_t n, int32_t s1, const void* v, const void* vxx, int32_t nvxx, const int32_t* n_accepted,
                          double v_lower, double v_upper, double vxx_max_norm, int32_t* counts, void* cuda_stream) {
    if (dtype != PMP_F32 && dtype != PMP_F64) return fail(PMP_ERR_BAD_ARG, "dtype");
    if (n < 0 || s1 < 1 || (n > 0 && (!v || !counts)) || (vxx && nvxx < 1)) return fail(PMP_ERR_BAD_ARG, "select_count: n, s1, v, counts");
    int rc = pmp::select_count(dtype, n, s1, v, vxx, nvxx, n_accepted, v_lower, v_upper, vxx_max_norm, counts, (cudaStream_t)cuda_stream);
    return rc ? cuda_fail(rc) : PMP_OK;
}

int pmp_select_write_cuda(int32_t dtype, int64_t n, int32_t s1, int32_t nx, const void* x, const void* t, const void* v, const void* vx,
                          const void* vxx, int32_t nvxx, const int32_t* n_accepted, const int64_t* offsets, double v_lower, double v_upper,
                          double vxx_max_norm, void* x_out, void* t_out, void* v_out, void* vx_out, void* vxx_out, double* stats,
                          int32_t* flags, void* cuda_stream) {
    if (dtype != PMP_F32 && dtype != PMP_F64) return fail(PMP_ERR_BAD_ARG, "dtype");
    if (n < 0 || s1 < 1 || nx < 1 || nx > 16) return fail(PMP_ERR_BAD_ARG, "select_write: n, s1, nx (<= 16)");
    if (n > 0 && (!x || !t || !v || !vx || !offsets || !x_out || !t_out || !v_out || !vx_out || !stats || !flags))
        return fail(PMP_ERR_BAD_ARG, "select_write: NULL argument");
    if ((vxx != nullptr) != (vxx_out != nullptr)) return fail(PMP_ERR_BAD_ARG, "select_write: vxx and vxx_out go together");
    int rc = pmp::select_write(dtype, n, s1, nx, x, t, v, vx, vxx, nvxx, n_accepted, (const long long*)offsets, v_lower, v_upper,
                               vxx_max_norm, x_out, t_out, v_out, vx_out, vxx_out, stats, flags, (cudaStream_t)cuda_stream);
    return rc ? cuda_fail(rc) : PMP_OK;
}

int pmp_comm_unique_id(void* id128) {
    if (!id128) return fail(PMP_ERR_BAD_ARG, "id128 is NULL");
    std::string msg;
    if (pmp::comm_unique_id(id128, &msg)) return fail(PMP_ERR_UNSUPPORTED, msg.c_str());
    return PMP_OK;
}

int pmp_comm_init(void** comm, int32_t world, int32_t rank, const void* id128) {
    if (!comm || !id128 || world < 1 || rank < 0 || rank >= world) return fail(PMP_ERR_BAD_ARG, "comm / id128 / world / rank");
    std::string msg;
    if (pmp::comm_init(comm, world, rank, id128, &msg)) return fail(PMP_ERR_UNSUPPORTED, msg.c_str());
    return PMP_OK;
}

int pmp_comm_destroy(void* comm) {
    if (!comm) return PMP_OK;
    std::string msg;
    if (pmp::comm_destroy(comm, &msg)) return fail(PMP_ERR_UNSUPPORTED, msg.c_str());
    return PMP_OK;
}

int pmp_allgather_endpoints(void* comm, const void* local, void* global, int64_t n_local, int32_t rows, int32_t dtype, void* cuda_stream) {
    if (!comm || n_local < 0 || rows < 1 || (dtype != PMP_F32 && dtype != PMP_F64)) return fail(PMP_ERR_BAD_ARG, "comm / n_local / rows / dtype");
    if (n_local == 0) return PMP_OK;
    if (!local || !global) return fail(PMP_ERR_BAD_ARG, "local / global is NULL");
    std::string msg;
    if (pmp::allgather(comm, local, global, (size_t)n_local * (size_t)rows, dtype, (cudaStream_t)cuda_stream, &msg))
        return fail(PMP_ERR_UNSUPPORTED, msg.c_str());
    return PMP_OK;
}

int pmp_host_probe_copy(void* dst, const void* src, int64_t bytes, int32_t blocks, void* cuda_stream) {
    if (!dst || !src || bytes < 0 || (bytes & 15) || blocks == 0) return fail(PMP_ERR_BAD_ARG, "probe copy: 16-byte multiples, blocks != 0");
    int rc = pmp::host_probe_copy(dst, src, bytes, blocks, (cudaStream_t)cuda_stream);
    return rc ? cuda_fail(rc) : PMP_OK;
}

int pmp_host_release(void) {
    int rc = pmp::host_release();
    return rc ? cuda_fail(rc) : PMP_OK;
}

int pmp_rhs_batch_cuda(const pmp_problem_t* p, int32_t dtype, int64_t n, const void* y, void* dy, void* cuda_stream) {
    if (int e = check_problem(p)) return e;
    if (dtype != PMP_F32 && dtype != PMP_F64) return fail(PMP_ERR_BAD_ARG, "dtype must be 0 (f32) or 1 (f64)");
    if (n < 0 || (n > 0 && (!y || !dy))) return fail(PMP_ERR_BAD_ARG, "bad n or NULL buffers");
    if (n == 0) return PMP_OK;
    if (pmp_device_count() <= 0) return fail(PMP_ERR_NO_DEVICE, "no CUDA device: libb200pmp has no CPU fallback");
    int rc = pmp::eval_batch(0, *p, dtype, n, y, nullptr, dy, (cudaStream_t)cuda_stream);
    return rc ? cuda_fail(rc) : PMP_OK;
}

int pmp_rhs_batch_cuda_flags(const pmp_problem_t* p, int32_t dtype, int32_t flags, int64_t n, const void* y, void* dy,
                             void* cuda_stream) {
    if (!(flags & PMP_FLAG_VXX)) {
        if (flags != 0) return fail(PMP_ERR_BAD_ARG, "pmp_rhs_batch_cuda_flags: only PMP_FLAG_VXX is meaningful here");
        return pmp_rhs_batch_cuda(p, dtype, n, y, dy, cuda_stream);
    }
    if (int e = check_problem(p)) return e;
    if (flags != PMP_FLAG_VXX) return fail(PMP_ERR_BAD_ARG, "pmp_rhs_batch_cuda_flags: only PMP_FLAG_VXX is meaningful here");
    if (int e = check_vxx_system(p, flags)) return e;
    if (dtype != PMP_F32 && dtype != PMP_F64) return fail(PMP_ERR_BAD_ARG, "dtype must be 0 (f32) or 1 (f64)");
    if (n < 0 || (n > 0 && (!y || !dy))) return fail(PMP_ERR_BAD_ARG, "bad n or NULL buffers");
    if (n == 0) return PMP_OK;
    if (pmp_device_count() <= 0) return fail(PMP_ERR_NO_DEVICE, "no CUDA device: libb200pmp has no CPU fallback");
    int rc = pmp::eval_batch(3, *p, dtype, n, y, nullptr, dy, (cudaStream_t)cuda_stream);
    return rc ? cuda_fail(rc) : PMP_OK;
}

int pmp_ustar_batch_cuda(const pmp_problem_t* p, int32_t dtype, int64_t n, const void* x, const void* lam, void* u,
                         void* cuda_stream) {
    if (int e = check_problem(p)) return e;
    if (dtype != PMP_F32 && dtype != PMP_F64) return fail(PMP_ERR_BAD_ARG, "dtype must be 0 (f32) or 1 (f64)");
    if (n < 0 || (n > 0 && (!x || !lam || !u))) return fail(PMP_ERR_BAD_ARG, "bad n or NULL buffers");
    if (n == 0) return PMP_OK;
    if (pmp_device_count() <= 0) return fail(PMP_ERR_NO_DEVICE, "no CUDA device: libb200pmp has no CPU fallback");
    int rc = pmp::eval_batch(1, *p, dtype, n, x, lam, u, (cudaStream_t)cuda_stream);
    return rc ? cuda_fail(rc) : PMP_OK;
}

static int check_nn(const pmp_problem_t* p, const pmp_nn_costate_t* nn) {
    if (!nn) return fail(PMP_ERR_BAD_ARG, "nn is NULL");
    if (nn->width != 32 && nn->width != 64) return fail(PMP_ERR_UNSUPPORTED, "NN costate: width must be 32 or 64");
    if (nn->n_hidden_layers < 1 || nn->n_hidden_layers > 4) return fail(PMP_ERR_UNSUPPORTED, "NN costate: 1..4 hidden layers");
    if (nn->n_nets < 1 || nn->n_nets > 16) return fail(PMP_ERR_UNSUPPORTED, "NN costate: 1..16 ensemble members");
    if (!nn->weights || ((uintptr_t)nn->weights & 15u) != 0) return fail(PMP_ERR_BAD_ARG, "NN costate: weights NULL or not 16-byte aligned");
    if (nn->reserved != 0) return fail(PMP_ERR_BAD_ARG, "NN costate: reserved != 0");
    for (int q = 0; q < p->nx; q++)
        if (!(nn->x_std[q] > 0)) return fail(PMP_ERR_BAD_ARG, "NN costate: x_std must be > 0");
    return 0;
}

static int forward_common(const pmp_problem_t* p, const pmp_opts_t* o, int64_t n, const void* x0, const double* P,
                          const double* x_eq, const pmp_nn_costate_t* nn, void* z_end, void* dense_ts, void* dense_x,
                          void* dense_cost, int32_t* n_accepted, int32_t* n_steps, int32_t* result, void* cuda_stream) {
    if (int e = check_problem(p)) return e;
    if (int e = check_opts(o)) return e;
    if (o->indep != PMP_INDEP_TIME || o->flags != 0) return fail(PMP_ERR_BAD_ARG, "forward simulation: indep must be TIME, flags 0");
    const int own_event = nn ? PMP_EVENT_NN_VALUE_UCB_LE : PMP_EVENT_LQR_VALUE_LE;
    if (o->event != PMP_EVENT_NONE && o->event != own_event)
        return fail(PMP_ERR_BAD_ARG, nn ? "forward simulation (NN costate): event must be NONE or PMP_EVENT_NN_VALUE_UCB_LE"
                                        : "forward simulation: event must be NONE or PMP_EVENT_LQR_VALUE_LE");
    if (!(o->t0 < o->t1)) return fail(PMP_ERR_BAD_ARG, "forward simulation: need t0 < t1");
    if (n < 0) return fail(PMP_ERR_BAD_ARG, "n < 0");
    if (nn)
        if (int e = check_nn(p, nn)) return e;
    if (n == 0) return PMP_OK;
    if (!x0 || (!P && !nn) || !z_end) return fail(PMP_ERR_BAD_ARG, "x0 / P / z_end is NULL");
    const bool dense = dense_ts && dense_x && dense_cost;
    if (o->save_dense ? !dense : (dense_ts || dense_x || dense_cost))
        return fail(PMP_ERR_BAD_ARG, "dense_ts, dense_x, dense_cost: all three iff save_dense");
    if (pmp_device_count() <= 0) return fail(PMP_ERR_NO_DEVICE, "no CUDA device: libb200pmp has no CPU fallback");
    int rc = pmp::forward_batch(*p, *o, n, x0, P, x_eq, nn, z_end, dense_ts, dense_x, dense_cost, n_accepted, n_steps, result,
                                (cudaStream_t)cuda_stream);
    if (rc == -1000) return fail(PMP_ERR_UNSUPPORTED, "no kernel for this (system, dtype, method) combination");
    return rc ? cuda_fail(rc) : PMP_OK;
}

int pmp_forward_batch_cuda(const pmp_problem_t* p, const pmp_opts_t* o, int64_t n, const void* x0, const double* P,
                           const double* x_eq, void* z_end, void* dense_ts, void* dense_x, void* dense_cost,
                           int32_t* n_accepted, int32_t* n_steps, int32_t* result, void* cuda_stream) {
    return forward_common(p, o, n, x0, P, x_eq, nullptr, z_end, dense_ts, dense_x, dense_cost, n_accepted, n_steps, result,
                          cuda_stream);
}

int pmp_nn_value_batch_cuda(const pmp_problem_t* p, const pmp_nn_costate_t* nn, int32_t dtype, int64_t n, const void* x,
                            void* lam, void* v_mean, void* v_std, void* cuda_stream) {
    if (int e = check_problem(p)) return e;
    if (int e = check_nn(p, nn)) return e;
    if (dtype != PMP_F32 && dtype != PMP_F64) return fail(PMP_ERR_BAD_ARG, "dtype must be 0 (f32) or 1 (f64)");
    if (n < 0 || (n > 0 && (!x || !lam || !v_mean || !v_std))) return fail(PMP_ERR_BAD_ARG, "bad n or NULL buffers");
    if (n == 0) return PMP_OK;
    if (pmp_device_count() <= 0) return fail(PMP_ERR_NO_DEVICE, "no CUDA device: libb200pmp has no CPU fallback");
    int rc = pmp::nn_value_batch(*p, nn, dtype, n, x, lam, v_mean, v_std, (cudaStream_t)cuda_stream);
    if (rc == -1000) return fail(PMP_ERR_UNSUPPORTED, "no kernel for this (nx, width) combination");
    return rc ? cuda_fail(rc) : PMP_OK;
}

int pmp_forward_nn_batch_cuda(const pmp_problem_t* p, const pmp_opts_t* o, const pmp_nn_costate_t* nn, int64_t n,
                              const void* x0, void* z_end, void* dense_ts, void* dense_x, void* dense_cost,
                              int32_t* n_accepted, int32_t* n_steps, int32_t* result, void* cuda_stream) {
    if (!nn) return fail(PMP_ERR_BAD_ARG, "nn is NULL");
    return forward_common(p, o, n, x0, nullptr, nullptr, nn, z_end, dense_ts, dense_x, dense_cost, n_accepted, n_steps, result,
                          cuda_stream);
}

int pmp_interp_batch_cuda(const pmp_problem_t* p, const pmp_opts_t* o, int64_t nq, const void* y0, const void* t1,
                          const void* target, int32_t mode, void* y_out, void* cuda_stream) {
    if (int e = check_problem(p)) return e;
    if (int e = check_opts(o)) return e;
    if (mode != 0 && mode != 1) return fail(PMP_ERR_BAD_ARG, "mode must be 0 (evaluate) or 1 (value level)");
    if (mode == 1 && o->indep != PMP_INDEP_TIME) return fail(PMP_ERR_UNSUPPORTED, "mode 1 needs a time-parameterised solve");
    if (nq < 0 || (nq > 0 && (!y0 || !t1 || !target || !y_out))) return fail(PMP_ERR_BAD_ARG, "bad nq or NULL buffers");
    if (nq == 0) return PMP_OK;
    if (pmp_device_count() <= 0) return fail(PMP_ERR_NO_DEVICE, "no CUDA device: libb200pmp has no CPU fallback");
    int rc = pmp::interp_batch(*p, *o, nq, y0, t1, target, y_out, mode, (cudaStream_t)cuda_stream);
    if (rc == -1000) return fail(PMP_ERR_UNSUPPORTED, "no kernel for this (system, dtype, method) combination");
    return rc ? cuda_fail(rc) : PMP_OK;
}

int pmp_ustar_batch_cuda_flags(const pmp_problem_t* p, int32_t dtype, int32_t flags, int64_t n, const void* x,
                               const void* lam, void* u, void* cuda_stream) {
    if (int e = check_problem(p)) return e;
    if (dtype != PMP_F32 && dtype != PMP_F64) return fail(PMP_ERR_BAD_ARG, "dtype must be 0 (f32) or 1 (f64)");
    if (flags != PMP_FLAG_USTAR_SMOOTH && (flags & ~PMP_FLAG_USTAR_LITERAL) != 0) return fail(PMP_ERR_BAD_ARG, "unknown flags");
    if (flags == PMP_FLAG_USTAR_SMOOTH && p->nu != 2) return fail(PMP_ERR_UNSUPPORTED, "PMP_FLAG_USTAR_SMOOTH: u_star_2d is for nu == 2");
    if (n < 0 || (n > 0 && (!x || !lam || !u))) return fail(PMP_ERR_BAD_ARG, "bad n or NULL buffers");
    if (n == 0) return PMP_OK;
    if (pmp_device_count() <= 0) return fail(PMP_ERR_NO_DEVICE, "no CUDA device: libb200pmp has no CPU fallback");
    const int what = flags == PMP_FLAG_USTAR_SMOOTH ? 4 : ((flags & PMP_FLAG_USTAR_LITERAL) ? 1 : 2);
    int rc = pmp::eval_batch(what, *p, dtype, n, x, lam, u, (cudaStream_t)cuda_stream);
    return rc ? cuda_fail(rc) : PMP_OK;
}

int pmp_fma_peak_cuda(int32_t dtype, int32_t variant, int32_t iters, void* scratch, double* flops_out, void* cuda_stream) {
    if (dtype != PMP_F32 && dtype != PMP_F64) return fail(PMP_ERR_BAD_ARG, "dtype must be 0 (f32) or 1 (f64)");
    if (variant < 0 || variant > 2 || iters < 1 || !scratch) return fail(PMP_ERR_BAD_ARG, "bad variant / iters / scratch");
    if (pmp_device_count() <= 0) return fail(PMP_ERR_NO_DEVICE, "no CUDA device");
    int rc = pmp::fma_peak(dtype, variant, iters, scratch, flops_out, (cudaStream_t)cuda_stream);
    return rc ? cuda_fail(rc) : PMP_OK;
}

}  // extern "C"

// pmp_hostq.h -- untyped description of pmp_solve_host's persistent launch; crosses compile units (pmp_host.cu builds
// it, pmp_capi.cu routes it to the system's unit, pmp_solve.cuh turns it into HostQueue<R>).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200pmp.h"

namespace pmp {

constexpr int kHostMaxPieces = 32;

struct HostQueueRaw {
    const void *in_x, *in_vx, *in_t, *in_v;   // state-dict leaves of the whole batch in device memory ([n][nx], [n])
    const void* term;                         // [nx*nx + nx] P, x_eq in the kernel's precision, or NULL (see HostQueue)
    void *out_x, *out_vx, *out_t, *out_v;     // endpoints, same layouts; any may be NULL
    uint32_t* out_pk;                         // packed counters, or NULL
    int32_t *out_na, *out_ns, *out_res;       // plain counters (when they do not fit the packed word), or NULL
    unsigned long long* counter;              // work-queue counter (zeroed by the caller before the launch)
    const unsigned long long* ready;          // written by the host->device stream: trajectories [0, *ready) have landed
    unsigned int *done, *abort;               // done[piece]: retired trajectories; abort: watchdog flag
    int n_pieces;
    long long bounds[kHostMaxPieces + 1];     // piece c = [bounds[c], bounds[c+1])
};

// pmp_capi.cu: the (system, dtype, method) instantiation; -1000 if there is none
int solve_hostq_dispatch(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const HostQueueRaw& q, cudaStream_t st);

}  // namespace pmp

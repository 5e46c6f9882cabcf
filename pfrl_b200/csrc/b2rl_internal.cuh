// Internal declarations shared by the translation units of libb2rl.so.
// Not part of the ABI (see include/b2rl.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "b2rl.h"

// ---- error plumbing --------------------------------------------------------
void b2rl_set_error(const char *fmt, ...);

#define B2RL_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) {                                              \
            b2rl_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,       \
                           cudaGetErrorString(e__));                           \
            return B2RL_ERR_CUDA;                                              \
        }                                                                      \
    } while (0)

#define B2RL_REQUIRE(cond, status, ...)                                        \
    do {                                                                       \
        if (!(cond)) {                                                         \
            b2rl_set_error(__VA_ARGS__);                                       \
            return (status);                                                   \
        }                                                                      \
    } while (0)

// ---- device-side scalar state ----------------------------------------------
// Lives in HBM so that kernels never need host-supplied counters; the host
// keeps mirrors for the (deterministic) host-driven paths.
struct B2rlDevState {
    long long napp;      // absolute append counter
    long long npop;      // absolute pop counter
    double max_priority; // collections/prioritized.py:32,114
    double last_total;   // sum-tree root read before the last sample's draws
    double last_min;     // min-tree root at the last sample
    int last_n;          // size of the last sample
    int pad;
};

static constexpr int B2RL_U_RING = 8;

// ---- the handle --------------------------------------------------------------
struct b2rl_replay {
    b2rl_replay_config cfg;
    int64_t P;      // pow2 >= capacity
    int64_t nslots; // 2P leaf / record slots, slot = abs_index mod 2P
    int levels;     // log2(nslots): leaves live at heap level `levels`

    // HBM
    uint8_t *parts;        // [part_capacity][part_bytes]
    int32_t *state_parts;  // [nslots][stack]
    int32_t *next_parts;   // [nslots][stack]
    uint8_t *action;       // [nslots][action_bytes]
    double *rewards;       // [nslots][n_step]
    uint8_t *len;          // [nslots]
    uint8_t *terminal;     // [nslots]
    double *sum;           // heap [2*nslots], node 1 = top, leaves at [nslots, 2*nslots)
    double *mn;            // same layout, empty = +inf
    B2rlDevState *st;
    int32_t *last_slots;   // [max_batch] slots drawn by the last sample
    double *last_prio;     // [max_batch] their priorities (as found)
    double *new_prio;      // [max_batch] scratch for update
    double *u_dev;         // [max_batch]
    int32_t *winner;       // [nslots] scratch for duplicate resolution (-1 idle)
    double *gamma_pow_dev; // [n_step + 1]
    int64_t device_bytes;

    // host mirrors
    int64_t napp, npop;
    int64_t part_head; // next part sequence number
    bool wait_priority;
    int last_n;
    int last_mode;

    // fused replay step (step.cu): cross-CTA "draws ready" counter, launch sequence
    // number, ring of pinned slots for host-drawn uniforms, deferred write-back
    unsigned long long *ready_dev;
    unsigned long long *times_dev; // [8 + 256] %globaltimer stamps of the last fused launch
    int last_step_grid;
    uint64_t step_seq;
    double *u_ring_pin, *u_ring_dev;
    cudaEvent_t u_ev[B2RL_U_RING];
    uint64_t u_head;
    int sm_count;
    bool pending;          // TD errors of the last sample registered, trees not yet updated
    const void *pend_err;
    int pend_is_f64, pend_n;
    double pend_alpha, pend_eps, pend_emin, pend_emax;

    // staging (host pinned + device), guarded by an event
    uint8_t *pin;
    uint8_t *stage;
    size_t stage_bytes;
    cudaEvent_t stage_ev;
    bool stage_busy;
};

static inline int b2rl_ilog2(int64_t x)
{
    int l = 0;
    while ((int64_t(1) << l) < x) l++;
    return l;
}

// Make sure the staging buffers hold `bytes`; waits for the previous user.
int b2rl_stage_acquire(b2rl_replay *h, size_t bytes);
int b2rl_stage_release(b2rl_replay *h, cudaStream_t s);

// "Last CTA finishes the reduction" elections use one of B2RL_N_TICKETS device
// counters per LAUNCH (rotating ticket, self-resetting atomicInc), so launches
// that overlap on different streams never share a counter.
static constexpr unsigned B2RL_N_TICKETS = 256;
unsigned b2rl_next_ticket();

// exact sampler generation for a tree of this depth (sampler.cu; env B2RL_SAMPLER=v5)
bool b2rl_use_v6(int levels);
int b2rl_v6_slow_every();
double b2rl_v6_eps_scale();
int b2rl_v6_sleep_scale();

// Apply a write-back registered with b2rl_per_defer_errors (no-op if none).
int b2rl_flush_pending(b2rl_replay *h, cudaStream_t s);

// Ancestors of up to 4 ring-slot ranges (<= 512 leaves in total) recomputed by the
// multi-CTA path kernel, then napp += bump_n and eviction down to capacity.
int b2rl_launch_repair_multi(b2rl_replay *h, int nranges, const long long *first_slot,
                             const long long *count, long long bump_n, cudaStream_t s);

// Kernel launchers implemented in the other TUs.
// Recompute the ancestors of up to 4 contiguous leaf-node ranges [lo, hi]
// (inclusive heap indices at the leaf level; modified in place), then add
// bump_n to the device append counter and evict down to capacity.
int b2rl_launch_tree_fix(b2rl_replay *h, int nranges, long long *lo, long long *hi,
                         long long bump_n, cudaStream_t s);

// soccer_planner.cuh -- Bellman backups and whole planners over the transition dynamics, on the device.
//
// The reference's planners (/root/reference/gym_soccer/utils/planners.py = PL) walk env.P: for every
// observation s and action key a, the list of (prob, next_state, reward, done) built by SIM:167-293, and
// accumulate   Q[s][a] += prob * (reward + discount_factor * V[next_state] * (not done))   (PL:11, 26, 39).
// Here a thread owns one (s, key), enumerates that list on the fly with the rules path -- same order: the 9
// slip combinations of SIM:209-223, zero-probability ones skipped, then the 1 / 2 / 4 outcomes of SIM:315-360
// -- and accumulates with __dmul_rn / __dadd_rn in the reference's operation order, so Q, V, the number of
// sweeps and the greedy policy are BIT-IDENTICAL to the reference's (tests/test_gpu_planners.py compares with
// ==).  No table is materialised: 761 x 5 lists of <= 15 entries are recomputed every sweep (~10 us).
//
// k_plan runs a whole value iteration (PL:4-18) or policy evaluation (PL:20-31) in ONE cooperative launch:
// sweep, grid-wide barrier, max-norm of the change through an atomicMax on the bit pattern of a non-negative
// double, barrier, convergence test -- no host round trip per sweep (the reference needs 40-183 sweeps).
// SIM = /root/reference/gym_soccer/envs/soccer_simultaneous_env.py.
#pragma once
#include <cooperative_groups.h>

#include "soccer_rules.cuh"

namespace soccer {

// One entry of Q: the reference's inner loop over P[s][key].
__device__ __forceinline__ double bellman_q(const PitchDev& P, const uint8_t* __restrict__ lut, int32_t s_obs, int key,
                                            const int8_t* __restrict__ policy_a, const int8_t* __restrict__ policy_b,
                                            const double* V, double gamma)
{
    // V is read with ld.global.cg: inside k_plan it is rewritten between sweeps by other SMs, so it must not go
    // through the non-coherent read-only path
    // P[0][key]: self-loops of the goal states (SIM:182-183, 300-301): prob * (0 + gamma * V[0] * False) = 0
    if (s_obs == 0) return 0.0;
    const uint32_t st = obs_to_packed(P, s_obs);
    const uint32_t a = st & 0xFFu, b = (st >> 8) & 0xFFu, p = (st >> 24) & 1u;
    uint32_t aa, ab;
    if (!policy_a && !policy_b) { aa = (uint32_t)key / 5u; ab = (uint32_t)key % 5u; }
    else if (policy_b) { aa = (uint32_t)key; ab = (uint32_t)policy_b[s_obs]; }        // SIM:187-188
    else { aa = (uint32_t)policy_a[s_obs]; ab = (uint32_t)key; }
    const bool flip = policy_a != nullptr;                                            // SIM:243-244
    double q = 0.0;
    for (int c = 0; c < 9; ++c) {
        const double mp = P.mp[c];
        if (mp == 0.0) continue;                                                      // SIM:226-227
        const int ca = combo_a(c), cb = combo_b(c);
        const uint32_t ma = ca == 0 ? aa : slip_move(aa, ca - 1);
        const uint32_t mb = cb == 0 ? ab : slip_move(ab, cb - 1);
        const Resolved o0 = resolve(lut, a, b, p, ma, mb, aa == 0, ab == 0, 0u);
        const uint32_t n = 1u << o0.nlog2;
        const double pr = __dmul_rn(mp, n == 4 ? 0.25 : (n == 2 ? 0.5 : 1.0));        // SIM:241
        for (uint32_t k = 0; k < n; ++k) {
            const Resolved o = resolve(lut, a, b, p, ma, mb, aa == 0, ab == 0, n == 2 ? (k << 1) : k);
            const StepOut f = finish_step<false>(P, o, 0u, 0u, 0u, flip);
            const bool done = (f.flags & 1u) != 0;
            // prob * (reward + discount_factor * V[next_state] * (not done)), left to right as Python evaluates it
            double t = __dmul_rn(gamma, __ldcg(V + f.obs));
            t = __dmul_rn(t, done ? 0.0 : 1.0);
            t = __dadd_rn((double)f.reward, t);
            q = __dadd_rn(q, __dmul_rn(pr, t));
        }
    }
    return q;
}

// Q[nS][nkeys] of one backup (PL:8-11, 35-39)
__global__ void __launch_bounds__(kThreads)
k_bellman_q(const PitchDev P, int32_t nS, const int8_t* __restrict__ policy_a, const int8_t* __restrict__ policy_b,
            const double* __restrict__ V, double gamma, double* __restrict__ Q)
{
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    build_cand_lut(lut, P);
    const int nkeys = (!policy_a && !policy_b) ? 25 : 5;
    const int64_t total = (int64_t)nS * nkeys;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
        Q[i] = bellman_q(P, lut, (int32_t)(i / nkeys), (int)(i % nkeys), policy_a, policy_b, V, gamma);
}

struct PlanArgs {
    int32_t nS; const int8_t* policy_a; const int8_t* policy_b;
    const int32_t* pi_in;          // nullptr: value iteration; else policy evaluation of pi_in[nS]
    double theta, gamma; int32_t max_sweeps;
    double* v0; double* v1;        // ping-pong value vectors [nS], v0 zero-filled by the caller
    unsigned long long* delta;     // [3] zero-filled: max-norm of the change, as the bits of a non-negative double
    double* V_out; double* Q_out; int32_t* pi_out; int32_t* sweeps_out;
};

__device__ __forceinline__ void block_max_to_global(double d, unsigned long long* dst, double* red)
{
    // non-negative doubles order like their bit patterns
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d = fmax(d, __shfl_xor_sync(0xFFFFFFFFu, d, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
        if (m > 0.0) atomicMax(dst, (unsigned long long)__double_as_longlong(m));
    }
    __syncthreads();
}

// PL:4-18 (value iteration) or PL:20-31 (policy evaluation) in one cooperative launch.
__global__ void __launch_bounds__(kThreads)
k_plan(const PitchDev P, const PlanArgs a)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    __shared__ double red[kThreads / 32];
    build_cand_lut(lut, P);
    const bool multi = !a.policy_a && !a.policy_b;
    const int nkeys = multi ? 25 : 5;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool eval = a.pi_in != nullptr;
    double* V = a.v0;
    double* Vn = a.v1;
    int32_t sweep = 0;
    while (true) {
        ++sweep;
        // three slots in rotation: the one cleared here was last read before the previous sweep's barrier,
        // and is written again only after this sweep's barrier
        unsigned long long* dslot = a.delta + sweep % 3;
        if (gtid == 0) a.delta[(sweep + 1) % 3] = 0ull;
        if (eval) {
            // V[s] = sum over P[s][pi[s]] (PL:25-26)
            double d = 0.0;
            for (int64_t s = gtid; s < a.nS; s += stride) {
                const double v = bellman_q(P, lut, (int32_t)s, a.pi_in[s], a.policy_a, a.policy_b, V, a.gamma);
                Vn[s] = v;
                d = fmax(d, fabs(__ldcg(V + s) - v));                // PL:27
            }
            block_max_to_global(d, dslot, red);
        } else {
            const int64_t total = (int64_t)a.nS * nkeys;
            for (int64_t i = gtid; i < total; i += stride)
                a.Q_out[i] = bellman_q(P, lut, (int32_t)(i / nkeys), (int)(i % nkeys), a.policy_a, a.policy_b, V, a.gamma);
            grid.sync();
            double d = 0.0;
            for (int64_t s = gtid; s < a.nS; s += stride) {
                double m = __ldcg(a.Q_out + s * nkeys);
                for (int k = 1; k < nkeys; ++k) m = fmax(m, __ldcg(a.Q_out + s * nkeys + k));   // np.max(Q, axis=1)
                Vn[s] = m;
                d = fmax(d, fabs(__ldcg(V + s) - m));                // PL:14
            }
            block_max_to_global(d, dslot, red);
        }
        grid.sync();
        const double dmax = __longlong_as_double((long long)*(volatile unsigned long long*)dslot);
        const bool stop = dmax < a.theta || sweep >= a.max_sweeps;   // the same value in every thread
        if (eval) {
            // PL:27-30: the freshly computed V is what policy_evaluation returns
            double* t = V; V = Vn; Vn = t;
            if (stop) break;
        } else {
            if (stop) break;                                         // PL:14-15: V stays the PREVIOUS vector
            double* t = V; V = Vn; Vn = t;                           // PL:16
        }
    }
    for (int64_t s = gtid; s < a.nS; s += stride) {
        a.V_out[s] = __ldcg(V + s);
        if (!eval && a.pi_out) {
            int best = 0;
            double m = __ldcg(a.Q_out + s * nkeys);
            for (int k = 1; k < nkeys; ++k) {
                const double q = __ldcg(a.Q_out + s * nkeys + k);
                if (q > m) { m = q; best = k; }                      // np.argmax: first maximum
            }
            a.pi_out[s] = best;
        }
    }
    if (gtid == 0) *a.sweeps_out = sweep;
}

} // namespace soccer

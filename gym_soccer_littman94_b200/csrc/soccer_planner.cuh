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

// ---- the DENSE planners: policy_eval (PL:57-70) and modified_policy_iteration (PL:73-87) are written against
// Pmat[s][s'][a] / Rmat[s][a] in the reference (np.dot).  Pmat is >= 99.9 % zeros -- a row has at most 15 non-zeros -- so
// the contraction is done sparsely, over the same on-the-fly list as above:
//     Rmat[s][a]           = sum over the list of prob * reward                      (SIM:263, same order: bit-identical)
//     (Pmat[s][:][a] . v)  = sum over the list of prob * v[next_obs]                 (SIM:262 merges equal next_obs first and
//                                                                                     BLAS sums in its own order: equal to
//                                                                                     fp64 round-off, ~1e-16 relative)
// There is no (not done) factor in the dense form: a goal contributes prob * v[0], and row 0 carries the reference's
// quirk Pmat[0][0][a] = (number of goal states) x sum of the combination probabilities (SIM:182-183, 262).
struct DenseSA { double r, pv; };
__device__ __forceinline__ DenseSA dense_backup(const PitchDev& P, const uint8_t* __restrict__ lut, int32_t s_obs, int key,
                                                const int8_t* __restrict__ policy_a, const int8_t* __restrict__ policy_b,
                                                const double* v, int32_t n_goal_states)
{
    DenseSA o = { 0.0, 0.0 };
    if (s_obs == 0) {
        double acc = 0.0;
        for (int gsi = 0; gsi < n_goal_states; ++gsi)
            for (int c = 0; c < 9; ++c)
                if (P.mp[c] != 0.0) acc = __dadd_rn(acc, __dmul_rn(P.mp[c], 1.0));
        o.pv = __dmul_rn(acc, __ldcg(v));
        return o;
    }
    const uint32_t st = obs_to_packed(P, s_obs);
    const uint32_t a = st & 0xFFu, b = (st >> 8) & 0xFFu, p = (st >> 24) & 1u;
    uint32_t aa, ab;
    if (!policy_a && !policy_b) { aa = (uint32_t)key / 5u; ab = (uint32_t)key % 5u; }
    else if (policy_b) { aa = (uint32_t)key; ab = (uint32_t)policy_b[s_obs]; }
    else { aa = (uint32_t)policy_a[s_obs]; ab = (uint32_t)key; }
    const bool flip = policy_a != nullptr;
    for (int c = 0; c < 9; ++c) {
        const double mp = P.mp[c];
        if (mp == 0.0) continue;
        const int ca = combo_a(c), cb = combo_b(c);
        const uint32_t ma = ca == 0 ? aa : slip_move(aa, ca - 1);
        const uint32_t mb = cb == 0 ? ab : slip_move(ab, cb - 1);
        const Resolved o0 = resolve(lut, a, b, p, ma, mb, aa == 0, ab == 0, 0u);
        const uint32_t n = 1u << o0.nlog2;
        const double pr = __dmul_rn(mp, n == 4 ? 0.25 : (n == 2 ? 0.5 : 1.0));
        for (uint32_t k = 0; k < n; ++k) {
            const Resolved r = resolve(lut, a, b, p, ma, mb, aa == 0, ab == 0, n == 2 ? (k << 1) : k);
            const StepOut f = finish_step<false>(P, r, 0u, 0u, 0u, flip);
            o.r = __dadd_rn(o.r, __dmul_rn(pr, (double)f.reward));
            o.pv = __dadd_rn(o.pv, __dmul_rn(pr, __ldcg(v + f.obs)));
        }
    }
    return o;
}

// q[s][key] = Rmat[s][key] + gamma * (Pmat[s][:][key] . v)   (PL:78-79)
__global__ void __launch_bounds__(kThreads)
k_dense_q(const PitchDev P, int32_t nS, int32_t n_goal_states, const int8_t* __restrict__ policy_a,
          const int8_t* __restrict__ policy_b, const double* __restrict__ v, double gamma, double* __restrict__ q)
{
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    build_cand_lut(lut, P);
    const int nkeys = (!policy_a && !policy_b) ? 25 : 5;
    const int64_t total = (int64_t)nS * nkeys;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const DenseSA d = dense_backup(P, lut, (int32_t)(i / nkeys), (int)(i % nkeys), policy_a, policy_b, v, n_goal_states);
        q[i] = __dadd_rn(d.r, __dmul_rn(gamma, d.pv));
    }
}

struct PolicyEvalArgs {
    int32_t nS, n_goal_states; const int8_t* policy_a; const int8_t* policy_b;
    const double* policy;          // [nS][nkeys] stochastic policy (PL:57: policy[s, :])
    double theta, gamma; int32_t max_sweeps;
    double* v0; double* v1;        // ping-pong vectors [nS]; v0 holds the initial v (PL:58)
    double* part;                  // [nS][nkeys] per-(s, key) terms of a sweep
    unsigned long long* delta;     // [3] zero-filled
    double* V_out; int32_t* sweeps_out;
};

// policy_eval(env, policy, theta, gamma, k, init) (PL:57-70) in ONE cooperative launch: at most max_sweeps sweeps of
//     v[s] <- policy[s, :] . Rmat[s, :] + gamma * (Pmat[s, :, :]^T v) . policy[s, :]
// until the sup-norm change is below theta.  One thread per (s, key) for the sparse contraction, one per s for the
// policy-weighted sum (in key order, like np.dot over 5 or 25 elements).
__global__ void __launch_bounds__(kThreads)
k_policy_eval(const PitchDev P, const PolicyEvalArgs a)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    __shared__ double red[kThreads / 32];
    build_cand_lut(lut, P);
    const int nkeys = (!a.policy_a && !a.policy_b) ? 25 : 5;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t total = (int64_t)a.nS * nkeys;
    double* V = a.v0;
    double* Vn = a.v1;
    int32_t sweep = 0;
    while (sweep < a.max_sweeps) {
        ++sweep;
        unsigned long long* dslot = a.delta + sweep % 3;
        if (gtid == 0) a.delta[(sweep + 1) % 3] = 0ull;
        for (int64_t i = gtid; i < total; i += stride) {
            const DenseSA d = dense_backup(P, lut, (int32_t)(i / nkeys), (int)(i % nkeys), a.policy_a, a.policy_b, V, a.n_goal_states);
            const double w = a.policy[i];
            // r_pi and p_pi terms of this key: combined per state below
            a.part[2 * i] = __dmul_rn(w, d.r);
            a.part[2 * i + 1] = __dmul_rn(d.pv, w);
        }
        grid.sync();
        double dl = 0.0;
        for (int64_t s = gtid; s < a.nS; s += stride) {
            double r_pi = 0.0, p_pi = 0.0;
            for (int k = 0; k < nkeys; ++k) {
                r_pi = __dadd_rn(r_pi, __ldcg(a.part + 2 * (s * nkeys + k)));
                p_pi = __dadd_rn(p_pi, __ldcg(a.part + 2 * (s * nkeys + k) + 1));
            }
            const double val = __dadd_rn(r_pi, __dmul_rn(a.gamma, p_pi));      // PL:65
            Vn[s] = val;
            dl = fmax(dl, fabs(val - __ldcg(V + s)));                          // PL:66
        }
        block_max_to_global(dl, dslot, red);
        grid.sync();
        const double dmax = __longlong_as_double((long long)*(volatile unsigned long long*)dslot);
        double* t = V; V = Vn; Vn = t;                                         // PL:67
        if (dmax < a.theta) break;                                             // PL:69-70
    }
    for (int64_t s = gtid; s < a.nS; s += stride) a.V_out[s] = __ldcg(V + s);
    if (gtid == 0) *a.sweeps_out = sweep;
}

} // namespace soccer

// soccer_rollout.cuh -- K2, rules variant: K fused steps with the state register-resident.
//
// Each thread owns VEC envs for all K steps; state, timestep and the current Philox block stay in
// registers; only the obs / reward / flags streams ([K][n]) are written, as 128 / 128 / 32-bit
// stores when VEC == 4.  Uniform random policy: the four envs advance through the byte-parallel
// step4_noslip().  On-device TABLE policies (int8[nS], the reference's utils/policies.py dict
// format) need the observation index per env and take the scalar step.  Any pitch; slip_prob == 0.
// Statistics cost ~2 instructions per env-step (see k_rollout_table).
// SIM = /root/reference/gym_soccer/envs/soccer_simultaneous_env.py.
#pragma once
#include "soccer_rules4.cuh"

namespace soccer {

template <int VEC>
__global__ void __launch_bounds__(kThreads)
k_rollout(const PitchDev P, uint32_t* __restrict__ state, const int8_t* __restrict__ policy_a,
          const int8_t* __restrict__ policy_b, uint64_t seed, uint64_t step0, int32_t K,
          uint64_t env_id_base, int32_t* __restrict__ obs, float* __restrict__ reward,
          uint8_t* __restrict__ flags, unsigned long long* __restrict__ stats, int64_t n)
{
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    __shared__ unsigned int blk_stats[4];
    __shared__ int blk_net;
    build_cand_lut(lut, P);
    if (threadIdx.x < 4) blk_stats[threadIdx.x] = 0;
    if (threadIdx.x == 4) blk_net = 0;
    __syncthreads();
    const Isd4 I = make_isd4(P);
    const bool table_policy = policy_a || policy_b;

    uint32_t c_done = 0, c_trunc = 0, c_len = 0, c_steps = 0;
    int32_t c_net = 0;
    const int64_t n_groups = n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint64_t step_end = step0 + (uint64_t)K;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        const int64_t i0 = g * VEC;
        uint32_t s[4] = { 0, 0, 0, 0 };
        if (VEC == 4) {
            const uint4 v = reinterpret_cast<const uint4*>(state)[g];
            s[0] = v.x; s[1] = v.y; s[2] = v.z; s[3] = v.w;
        } else {
            s[0] = state[i0];
        }
        uint32_t t_in = 0;
#pragma unroll
        for (int e = 0; e < VEC; ++e) t_in += (s[e] >> 16) & 0xFFu;
        uint32_t acc_d = 0, acc_t = 0, since_flush = 0;
        int32_t* op = obs ? obs + i0 : nullptr;
        float* rp = reward ? reward + i0 : nullptr;
        uint8_t* fp = flags ? flags + i0 : nullptr;
        for (uint64_t blk = step0 >> 2; (blk << 2) < step_end; ++blk) {
            uint32_t w[VEC][4];
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const uint64_t env = env_id_base + (uint64_t)(i0 + e);
                philox4x32_10((uint32_t)env, (uint32_t)(env >> 32), (uint32_t)blk, (uint32_t)(blk >> 32),
                              (uint32_t)seed, (uint32_t)(seed >> 32), w[e]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint64_t step = (blk << 2) + (uint64_t)j;
                if (step < step0 || step >= step_end) continue;         // warp-uniform
                uint32_t oo[4], rr[4], fw = 0;
                if (VEC == 4 && !table_policy) {
                    uint32_t aa[4], ab[4], rg[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t word = w[e % VEC][j];
                        philox_actions(word, aa[e], ab[e]);
                        rg[e] = philox_rng8(word);
                    }
                    Step4 o;
                    step4_noslip<false>(P, I, lut, s, pack4(aa[0], aa[1], aa[2], aa[3]), pack4(ab[0], ab[1], ab[2], ab[3]),
                                        pack4(rg[0], rg[1], rg[2], rg[3]), o);
#pragma unroll
                    for (int e = 0; e < 4; ++e) { s[e] = o.s[e]; oo[e] = o.obs[e]; rr[e] = o.rew[e]; }
                    fw = o.flags4;
                    c_net += o.rew_sum;
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        const uint32_t word = w[e][j];
                        uint32_t aa, ab;
                        philox_actions(word, aa, ab);
                        if (table_policy) {     // SIM:187-188: a table policy acts on the current observation
                            const int32_t cur = obs_index(P, s[e] & 0xFFu, (s[e] >> 8) & 0xFFu, (s[e] >> 24) & 1u);
                            if (policy_a) aa = (uint32_t)policy_a[cur];
                            if (policy_b) ab = (uint32_t)policy_b[cur];
                        }
                        const StepOut o = step_noslip<true, false>(P, lut, s[e], aa, ab, philox_rng8(word), false);
                        s[e] = o.state; oo[e] = (uint32_t)o.obs; rr[e] = __float_as_uint(o.reward);
                        fw |= (o.flags & 3u) << (8 * e);
                        c_net += (o.reward > 0.0f) - (o.reward < 0.0f);
                    }
                }
                acc_d += fw & 0x01010101u;
                acc_t += (fw >> 1) & ~fw & 0x01010101u;                 // truncated WITHOUT a goal
                if (VEC == 4) {
                    if (op) { st_stream(reinterpret_cast<uint4*>(op), make_uint4(oo[0], oo[1], oo[2], oo[3])); op += n; }
                    if (rp) { st_stream(reinterpret_cast<uint4*>(rp), make_uint4(rr[0], rr[1], rr[2], rr[3])); rp += n; }
                    if (fp) { st_stream(reinterpret_cast<uint32_t*>(fp), fw); fp += n; }
                } else {
                    if (op) { *op = (int32_t)oo[0]; op += n; }
                    if (rp) { *rp = __uint_as_float(rr[0]); rp += n; }
                    if (fp) { *fp = (uint8_t)fw; fp += n; }
                }
            }
            since_flush += 4;
            if (since_flush >= 64) {                                    // bytes hold at most 64 + 3 counts
                c_done = __dp4a(acc_d, 0x01010101u, c_done); c_trunc = __dp4a(acc_t, 0x01010101u, c_trunc);
                acc_d = acc_t = since_flush = 0;
            }
        }
        c_done = __dp4a(acc_d, 0x01010101u, c_done); c_trunc = __dp4a(acc_t, 0x01010101u, c_trunc);
        uint32_t t_out = 0;
#pragma unroll
        for (int e = 0; e < VEC; ++e) t_out += (s[e] >> 16) & 0xFFu;
        c_len += t_in + (uint32_t)K * VEC - t_out;      // sum(t_in) + steps = sum(finished lengths) + sum(t_out)
        c_steps += (uint32_t)K * VEC;
        if (VEC == 4) reinterpret_cast<uint4*>(state)[g] = make_uint4(s[0], s[1], s[2], s[3]);
        else state[i0] = s[0];
    }
    if (stats) {
        uint32_t v[4] = { c_done, c_trunc, c_steps, c_len };
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t r = __reduce_add_sync(0xFFFFFFFFu, v[j]);
            if ((threadIdx.x & 31) == 0 && r) atomicAdd(&blk_stats[j], r);
        }
        const int32_t rn = __reduce_add_sync(0xFFFFFFFFu, c_net);
        if ((threadIdx.x & 31) == 0 && rn) atomicAdd(&blk_net, rn);
        __syncthreads();
        if (threadIdx.x == 0) {
            // episodes = goals + truncated-only (a goal on step 100 carries both flags and counts once)
            const long long d = blk_stats[0], tr_only = blk_stats[1], net = blk_net;
            atomicAdd(&stats[0], (unsigned long long)(d + tr_only));
            atomicAdd(&stats[1], (unsigned long long)((d + net) / 2));        // goals_A (reward +1)
            atomicAdd(&stats[2], (unsigned long long)((d - net) / 2));        // goals_B (reward -1)
            atomicAdd(&stats[3], (unsigned long long)tr_only);
            atomicAdd(&stats[4], (unsigned long long)blk_stats[2]);
            atomicAdd(&stats[5], (unsigned long long)blk_stats[3]);
        }
    }
}

} // namespace soccer

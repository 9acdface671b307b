// soccer_rollout.cuh -- K2: K fused steps with the state register-resident.
//
//   k_rollout_table        5x4-class pitches: transition table in shared memory (soccer_table.cuh), uniform or table
//                          policies; 148 x 512-thread CTAs
//   k_rollout_table_slipi  the same for slip_prob > 0: combination and slot of the 32-bit Philox draw from constant
//                          integer thresholds (slip index plane 1); the reference's cumulative walk only where flagged
//   k_rollout              any pitch: rules inline -- byte-parallel step4_noslip() for the uniform policy,
//                          the scalar step for on-device TABLE policies (int8[nS], the reference's
//                          utils/policies.py dict format); 256-thread CTAs
//
// Each thread owns VEC envs for all K steps; state and timestep stay in registers; only the obs / reward / flags
// streams ([K][n]) are written, as 128 / 128 / 32-bit stores when VEC == 4.  Randomness: Philox contract v2
// (soccer_rules.cuh) -- ONE Philox4x32-10 call per step yields the words of the thread's four envs (an aligned
// group of the global env ids).
// Episode statistics cost ~2 instructions per env-step: episodes / truncations by byte-parallel
// accumulation of the packed flags word (flushed with dp4a every 64 steps); goals_A - goals_B =
// sum of rewards; sum_episode_len from sum(t_in) + K*VEC = sum(finished lengths) + sum(t_out).
// SIM = /root/reference/gym_soccer/envs/soccer_simultaneous_env.py.
#pragma once
#include "soccer_rules4.cuh"
#include "soccer_table.cuh"

namespace soccer {

// ---- slip_prob > 0 for the RULES kernels (any pitch), byte-parallel: the integer-threshold decision of
// soccer_table.cuh is state-independent, so the four envs of a thread pick their slip combination and their slot inside
// a 2-way / 4-way collision from the constant tables, and step4_noslip resolves the slipped moves exactly as it resolves
// chosen ones (a slipped move is NOOP iff the action is, so the NOOP-keyed cases of SIM:330-344 agree).  The (in
// practice never taken) walk of a listed draw re-does that env with the scalar step_slip().
// aa / ab: action bytes (values < 5 or clamped here); r32: the 32-bit draws; RST: rng8-style bytes with the reset draw
// in bits 2..3.
__device__ __noinline__ StepOut step_slip_call(const PitchDev& P, const uint8_t* lut, uint32_t prt, uint32_t first_k, uint32_t s,
                                               uint32_t aa, uint32_t ab, uint32_t r32, uint32_t reset_sel)
{
    const SlipCtx sc = { prt, first_k };
    return step_slip<true>(P, lut, sc, s, aa, ab, u_from_rng32(r32), reset_sel, false);
}
// J4: the joint action of each env as one index byte -- WIDE: aa * 8 + ab (caller-supplied action bytes, 3-bit fields),
// else aa * 5 + ab (the Philox joint action); the move pair of (combination, joint action) is ONE byte ma | mb << 4.
// in / out: the four states as packed bytes (K2 keeps them that way between steps); o.s is not written.
template <bool RESET_OBS, bool WIDE>
__device__ __forceinline__ void step4_slip_int(const PitchDev& P, const Isd4& I, const uint8_t* __restrict__ lut,
                                               const SlipInt& f, const SlipCtx& sc, const Soa4& in,
                                               uint32_t J4, const uint32_t r32[4], uint32_t RST, Step4& o, Soa4& out)
{
    uint32_t mv[4], r4[4], r2[4], walk = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const uint32_t k = slip_int_k(f, r32[e]);
        walk |= (k >= 9u ? 1u : 0u) << e;
        const uint32_t kc = min(k, 8u);
        mv[e] = lds_u8_r((WIDE ? f.mvs() + kc * 64u : f.mvj() + kc * 32u) + byte_of(J4, e));
        const uint32_t t2 = lds_u32_r(f.sl() + kc * 32u);                      // 2-way row: one threshold
        const uint4 t4 = lds_v4_r(f.sl() + kc * 32u + 16u);                    // 4-way row: three
        r2[e] = r32[e] > t2 ? 2u : 0u;                                       // draw value 2 * slot
        r4[e] = (r32[e] > t4.x ? 1u : 0u) + (r32[e] > t4.y ? 1u : 0u) + (r32[e] > t4.z ? 1u : 0u);
    }
    const uint32_t MV = pack4(mv[0], mv[1], mv[2], mv[3]);
    step4_core<RESET_OBS>(P, I, lut, in, MV, MV >> 4, pack4(r4[0], r4[1], r4[2], r4[3]), o, pack4(r2[0], r2[1], r2[2], r2[3]), RST, out);
    if (walk) {                                                              // (in practice never)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if ((walk >> e) & 1u) {
                const uint32_t j = byte_of(J4, e);
                const StepOut w = step_slip_call(P, lut, sc.prt, sc.first_k, soa4_word(in, e), WIDE ? min(j >> 3, 4u) : min(j / 5u, 4u),
                                                 WIDE ? min(j & 7u, 4u) : j % 5u, r32[e], (byte_of(RST, e) >> 2) & 3u);
                const int32_t ri = (w.reward > 0.0f) - (w.reward < 0.0f);
                o.rew_sum += ri - (int32_t)(signed char)(o.rew4 >> (8 * e));
                soa4_set_word(out, e, w.state);
                o.obs[e] = (uint32_t)w.obs; o.rew[e] = __float_as_uint(w.reward);
                o.rew4 = (o.rew4 & ~(0xFFu << (8 * e))) | (((uint32_t)ri & 0xFFu) << (8 * e));
                o.flags4 = (o.flags4 & ~(0xFFu << (8 * e))) | ((w.flags & 3u) << (8 * e));
                if (RESET_OBS) o.robs[e] = (uint32_t)w.reset_obs;
            }
        }
    }
}

#ifndef SOCCER_K2_SLIPI_BLOCK
#define SOCCER_K2_SLIPI_BLOCK 4        // steps of the slip table rollout whose Philox calls are issued together (1, 2 or 4)
#endif
#ifndef SOCCER_K2_PHILOX_WIDE
#define SOCCER_K2_PHILOX_WIDE 1
#endif
struct RolloutArgs {
    uint32_t* state; uint64_t seed, step0; int32_t K; uint64_t env_id_base;
    int32_t* obs; float* reward; uint8_t* flags; unsigned long long* stats; int64_t n;
    PhiloxRoundKeys rk;   // philox_round_keys(seed)
    int32_t flip;    // 1: the streamed reward is player B's (= -A's): the env's return agent is player_b (SIM:243-244).
                     // The goals_A / goals_B statistics always count by the unflipped sign.
};

// ---- steppers: advance the VEC envs of a thread by one step given their Philox words.
// kCollective: step() contains warp-level exchanges, so every lane of the warp has to call it every step.
// A variant of the table stepper that moved the selects / shifts / compares onto the FMA pipe (mul.hi carries, IMAD
// packing, 2^23 magic-add int -> float; ALU pipe 81 % -> 71 % busy) measured 2 % SLOWER on B200 in three A/B
// runs: with 2^20+ envs the kernel sits at the HBM write ceiling (0.92-0.99 of the traffic probe), not on issue.
template <bool POLICY, bool SLIP>
struct TableStepper {
    static constexpr bool kCollective = false, kHasPolicy = POLICY;
    TblCtx c;
    uint32_t pol_a, pol_b;      // shared-window addresses of the int8[nS] table policies, 0 = uniform (POLICY only)
    SlipCtx sc;                 // slip-combination probability table (SLIP only)
    __device__ __forceinline__ uint32_t timestep(uint32_t s) const { return s >> 16; }
    template <int VEC> __device__ __forceinline__ void enter(uint32_t*) const {}
    template <int VEC> __device__ __forceinline__ void leave(uint32_t*) const {}
    template <int VEC>
    __device__ __forceinline__ void step(uint32_t* s, const uint32_t* word, uint32_t* oo, uint32_t* rr,
                                         uint32_t& fw, int32_t& net, bool flip) const
    {
        uint32_t ff[4] = { 0, 0, 0, 0 };
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            // jr = mulhi(w, 100) is the table column; w & 0xC the byte offset of the start observation (reset draw = bits 2..3)
            uint32_t jr = philox_jr(word[e]);
            uint32_t aa = 0, ab = 0;
            if (POLICY || SLIP) {
                philox_actions(word[e], aa, ab);
                // SIM:187-188: a table policy picks the player's action from the CURRENT observation;
                // the other player's action and the draws stay the Philox ones
                const uint32_t cur = min(s[e] & 0xFFFFu, c.maxobs);
                if (POLICY && pol_a) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(aa) : "r"(pol_a + cur));
                if (POLICY && pol_b) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(ab) : "r"(pol_b + cur));
                jr = aa * 20u + ab * 4u + (jr & 3u);
            }
            // slip_prob > 0: the in-place walk with the word's 32-bit step draw (6x4, where the slip index does not fit)
            const TblOut o = SLIP ? table_step_slip(c, sc, s[e], aa, ab, u_from_rng32(philox_r32(word[e])),
                                                    word[e] & 0xCu)
                                  : table_step(c, s[e], jr, word[e] & 0xCu);
            s[e] = o.state; oo[e] = o.obs; rr[e] = __float_as_uint((float)(flip ? -o.rew_i : o.rew_i)); ff[e] = o.flags;
            net += o.rew_i;
        }
        fw = VEC == 4 ? pack4(ff[0], ff[1], ff[2], ff[3]) : ff[0];
    }
};

// slip_prob > 0 with the slip index (5x4): every Philox draw is a 32-bit draw, so combination AND slot come from
// constant integer thresholds (table_step_slip_int, soccer_table.cuh); the reference's cumulative walk is only taken
// by the (in practice non-existent) picks the index flags.  4 envs per thread like the slip-0 stepper.
// (First version of this round: the fp64 constant-prefix fast path + a per-step warp queue through shared memory
// for the ~15 % of the env-steps it could not decide -- bit-identical, 150-164 G env-steps/s: one 170-instruction walk
// pass per warp and step whatever the number of queued envs, + ~25 instructions of queue bookkeeping per env-step;
// profiles/r02b_time_round2.log, r02b_ncu_k2slip_queue_by_line.txt.)
template <bool POLICY>
struct TableSlipIntStepper {
    static constexpr bool kCollective = false, kHasPolicy = POLICY;
    static constexpr int kBlock = SOCCER_K2_SLIPI_BLOCK;
    TblCtx c; SlipCtx sc; SlipInt sf;
    uint32_t pol_a, pol_b;
    __device__ __forceinline__ uint32_t timestep(uint32_t s) const { return s >> 16; }
    template <int VEC> __device__ __forceinline__ void enter(uint32_t*) const {}
    template <int VEC> __device__ __forceinline__ void leave(uint32_t*) const {}
    template <int VEC>
    __device__ __forceinline__ void step(uint32_t* s, const uint32_t* word, uint32_t* oo, uint32_t* rr,
                                         uint32_t& fw, int32_t& net, bool flip) const
    {
        uint32_t ff[4] = { 0, 0, 0, 0 }, so[4] = { 0, 0, 0, 0 }, ri[4] = { 0, 0, 0, 0 }, walks = 0;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const uint32_t r32 = philox_r32(word[e]), rsel4 = word[e] & 0xCu;
            bool walk;
            TblOut o;
            if (POLICY) {
                uint32_t aa, ab;
                philox_actions(word[e], aa, ab);
                const uint32_t cur = min(s[e] & 0xFFFFu, c.maxobs);
                if (pol_a) aa = lds_u8_r(pol_a + cur);
                if (pol_b) ab = lds_u8_r(pol_b + cur);
                o = table_step_slip_int<false>(c, sf, s[e], aa, ab, r32, rsel4, walk);
            } else {
                // the joint action mulhi(w, 25) indexes the (combination, joint action) -> move pair table directly
                o = table_step_slip_int_j<false>(c, sf, s[e], philox_ja(word[e]), r32, rsel4, walk);
            }
            walks |= walk ? 1u << e : 0u;
            so[e] = o.state; oo[e] = o.obs; ri[e] = (uint32_t)o.rew_i; ff[e] = o.flags;
        }
        if (walks) {                                                          // (in practice never): ONE test per step
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                if (!((walks >> e) & 1u)) continue;
                uint32_t aa, ab;
                philox_actions(word[e], aa, ab);
                const uint32_t cur = min(s[e] & 0xFFFFu, c.maxobs);
                if (POLICY && pol_a) aa = lds_u8_r(pol_a + cur);
                if (POLICY && pol_b) ab = lds_u8_r(pol_b + cur);
                const TblOut o = table_step_slip_walk(c, sc, s[e], aa, ab, philox_r32(word[e]), word[e] & 0xCu);
                so[e] = o.state; oo[e] = o.obs; ri[e] = (uint32_t)o.rew_i; ff[e] = o.flags;
            }
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            s[e] = so[e];
            rr[e] = __float_as_uint((float)(flip ? -(int32_t)ri[e] : (int32_t)ri[e]));
            net += (int32_t)ri[e];
        }
        fw = VEC == 4 ? pack4(ff[0], ff[1], ff[2], ff[3]) : ff[0];
    }
};

// SLIP is a template parameter so that the slip-0 instantiation carries none of the walk's registers (with a
// run-time branch the 4-env kernel needed 170 registers instead of 116 and lost its second resident CTA)
template <bool SLIP>
struct RulesStepper {
    static constexpr bool kCollective = false, kHasPolicy = true;
    const PitchDev& P; const uint8_t* lut; Isd4 I; const int8_t* policy_a; const int8_t* policy_b; SlipCtx sc;
    uint32_t jlut;      // shared-window address of the 100-byte decode table of mulhi(w, 100): aa | ab << 3 | r << 6
    __device__ __forceinline__ uint32_t timestep(uint32_t s) const { return (s >> 16) & 0xFFu; }
    // the byte-parallel path keeps the four states as packed bytes (Soa4 in s[0..3]) between enter() and leave()
    template <int VEC> __device__ __forceinline__ bool packed() const { return !SLIP && VEC == 4; }
    template <int VEC> __device__ __forceinline__ void enter(uint32_t* s) const
    {
        if (packed<VEC>()) { const Soa4 x = soa4_from_words(s); s[0] = x.A; s[1] = x.B; s[2] = x.T; s[3] = x.P; }
    }
    template <int VEC> __device__ __forceinline__ void leave(uint32_t* s) const
    {
        if (packed<VEC>()) { const Soa4 x = { s[0], s[1], s[2], s[3] }; soa4_to_words(x, s); }
    }
    template <int VEC>
    __device__ __forceinline__ void step(uint32_t* s, const uint32_t* word, uint32_t* oo, uint32_t* rr,
                                         uint32_t& fw, int32_t& net, bool flip) const
    {
        if (SLIP) {
            // slip_prob > 0 (SIM:203-227): the scalar 9-combination walk with the word's 32-bit step draw
            fw = 0;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                uint32_t aa, ab;
                philox_actions(word[e], aa, ab);
                if (policy_a || policy_b) {
                    const int32_t cur = obs_index(P, s[e] & 0xFFu, (s[e] >> 8) & 0xFFu, (s[e] >> 24) & 1u);
                    if (policy_a) aa = (uint32_t)policy_a[cur];
                    if (policy_b) ab = (uint32_t)policy_b[cur];
                }
                const StepOut o = step_slip<true>(P, lut, sc, s[e], aa, ab, u_from_rng32(philox_r32(word[e])),
                                                  (word[e] >> 2) & 3u, false);
                const int32_t ri = (o.reward > 0.0f) - (o.reward < 0.0f);
                s[e] = o.state; oo[e] = (uint32_t)o.obs; rr[e] = __float_as_uint((float)(flip ? -ri : ri));
                fw |= (o.flags & 3u) << (8 * e);
                net += ri;
            }
        } else if (VEC == 4) {
            // decode by table: byte = aa | ab << 3 | r << 6 at index mulhi(w, 100) (one IMAD.HI + one LDS per env on
            // the idle pipes instead of seven integer instructions on the ALU pipe that binds this kernel); the reset
            // draw is bits 2..3 of w itself
            uint32_t d[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) d[e] = lds_u8_r(jlut + philox_jr(word[e % VEC]));
            const uint32_t D = pack4(d[0], d[1], d[2], d[3]);
            const uint32_t W = pack4(word[0], word[1 % VEC], word[2 % VEC], word[3 % VEC]);
            const Soa4 in = { s[0], s[1], s[2], s[3] };
            uint32_t MA4 = D, MB4 = D >> 3;
            if (policy_a || policy_b) {     // SIM:187-188: a table policy acts on the current observation (gathered per env)
                uint32_t pa[4], pb[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t cur = min((uint32_t)obs_index(P, byte_of(in.A, e), byte_of(in.B, e), byte_of(in.P, e)), P.nSm1);
                    pa[e] = policy_a ? (uint32_t)(uint8_t)policy_a[cur] : 0u;
                    pb[e] = policy_b ? (uint32_t)(uint8_t)policy_b[cur] : 0u;
                }
                if (policy_a) MA4 = pack4(pa[0], pa[1], pa[2], pa[3]);
                if (policy_b) MB4 = pack4(pb[0], pb[1], pb[2], pb[3]);
            }
            Soa4 out;
            Step4 o;
            step4_core<false>(P, I, lut, in, MA4, MB4, D >> 6, o, D >> 6, W, out);
            s[0] = out.A; s[1] = out.B; s[2] = out.T; s[3] = out.P;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                oo[e] = o.obs[e];
                rr[e] = flip ? __float_as_uint((float)(-(int)(signed char)(o.rew4 >> (8 * e)))) : o.rew[e];
            }
            fw = o.flags4;
            net += o.rew_sum;
        } else {
            fw = 0;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                uint32_t aa, ab;
                philox_actions(word[e], aa, ab);
                if (policy_a || policy_b) {     // SIM:187-188: a table policy acts on the current observation
                    const int32_t cur = obs_index(P, s[e] & 0xFFu, (s[e] >> 8) & 0xFFu, (s[e] >> 24) & 1u);
                    if (policy_a) aa = (uint32_t)policy_a[cur];
                    if (policy_b) ab = (uint32_t)policy_b[cur];
                }
                const StepOut o = step_noslip<true, false>(P, lut, s[e], aa, ab, philox_rng8(word[e]), false);
                const int32_t ri = (o.reward > 0.0f) - (o.reward < 0.0f);
                s[e] = o.state; oo[e] = (uint32_t)o.obs; rr[e] = __float_as_uint((float)(flip ? -ri : ri));
                fw |= (o.flags & 3u) << (8 * e);
                net += ri;
            }
        }
    }
};

// slip_prob > 0, uniform policy, any pitch: four envs per thread, combination and slot by integer thresholds, resolution
// byte-parallel (step4_slip_int)
template <bool POLICY>
struct RulesSlipIntStepper {
    static constexpr bool kCollective = false, kHasPolicy = POLICY;
    const PitchDev& P; const uint8_t* lut; Isd4 I; SlipCtx sc; SlipInt fi; const int8_t* policy_a; const int8_t* policy_b;
    __device__ __forceinline__ uint32_t timestep(uint32_t s) const { return (s >> 16) & 0xFFu; }
    template <int VEC> __device__ __forceinline__ void enter(uint32_t* s) const
    {
        const Soa4 x = soa4_from_words(s); s[0] = x.A; s[1] = x.B; s[2] = x.T; s[3] = x.P;
    }
    template <int VEC> __device__ __forceinline__ void leave(uint32_t* s) const
    {
        const Soa4 x = { s[0], s[1], s[2], s[3] }; soa4_to_words(x, s);
    }
    template <int VEC>
    __device__ __forceinline__ void step(uint32_t* s, const uint32_t* word, uint32_t* oo, uint32_t* rr,
                                         uint32_t& fw, int32_t& net, bool flip) const
    {
        static_assert(VEC == 4, "four envs per thread");
        uint32_t ja[4], r32[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            ja[e] = philox_ja(word[e]);
            r32[e] = philox_r32(word[e]);
        }
        const Soa4 in = { s[0], s[1], s[2], s[3] };
        const uint32_t W = pack4(word[0], word[1], word[2], word[3]);
        Soa4 out;
        Step4 o;
        if (POLICY) {                   // SIM:187-188: a table policy acts on the current observation (gathered per env)
            uint32_t j8[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                uint32_t aa = (ja[e] * 52u) >> 8, ab = ja[e] - aa * 5u;
                const uint32_t cur = min((uint32_t)obs_index(P, byte_of(in.A, e), byte_of(in.B, e), byte_of(in.P, e)), P.nSm1);
                if (policy_a) aa = (uint32_t)(uint8_t)policy_a[cur] & 7u;
                if (policy_b) ab = (uint32_t)(uint8_t)policy_b[cur] & 7u;
                j8[e] = aa * 8u + ab;
            }
            step4_slip_int<false, true>(P, I, lut, fi, sc, in, pack4(j8[0], j8[1], j8[2], j8[3]), r32, W, o, out);
        } else {
            step4_slip_int<false, false>(P, I, lut, fi, sc, in, pack4(ja[0], ja[1], ja[2], ja[3]), r32, W, o, out);
        }
        s[0] = out.A; s[1] = out.B; s[2] = out.T; s[3] = out.P;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            oo[e] = o.obs[e];
            rr[e] = flip ? __float_as_uint((float)(-(int)(signed char)(o.rew4 >> (8 * e)))) : o.rew[e];
        }
        fw = o.flags4;
        net += o.rew_sum;
    }
};

// steps per Philox block: 4 unless the stepper says otherwise (kBlock)
template <class S, class = void> struct stepper_block { static constexpr int value = 4; };
template <class S> struct stepper_block<S, decltype((void)S::kBlock)> { static constexpr int value = S::kBlock; };

// 64-bit warp sum (statistics only; once per thread at the end of the kernel)
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// per-CTA statistics block in shared memory: done, trunc_only, steps, len (unsigned) and net (signed, two's complement)
struct BlkStats { unsigned long long v[5]; };

// ---- the K-step loop shared by all variants.  STREAMS: all three output streams present.
template <int VEC, bool STREAMS, class Stepper>
__device__ __forceinline__ void rollout_body(const Stepper& S, const RolloutArgs& a, BlkStats* blk)
{
    // 64-bit totals (a CTA of a 2^26-env batch steps 450 k envs: 32-bit counters would wrap at K ~ 9.5 k)
    unsigned long long c_done = 0, c_trunc = 0, c_len = 0, c_steps = 0;
    long long c_net = 0;
    const int64_t n_groups = a.n / VEC;
    const int32_t K = a.K;
    // only an env with a folded policy has a flipped return agent: the uniform-policy instantiations carry no flip code
    const bool flip = Stepper::kHasPolicy && a.flip != 0;
    // Slot order (a slot = 32 threads x VEC envs).  Full passes are CTA-major: the 16 warps of a CTA own 16
    // adjacent slots, so every step the SM writes 8 KB / 8 KB / 2 KB contiguous per stream - measured with the
    // traffic probe (profiles/probe_modes.py), 8 KB chunks written by ONE SM reach 5.9-6.4 TB/s where 512 B chunks
    // whose neighbours come from other SMs (arbitrary skew) reach 5.7-5.9.  The last, partial pass is dealt
    // warp by warp round-robin over the CTAs, so it leaves every SM the same number of busy warps instead of whole
    // SMs idle: 2^20 envs are 8192 slots = 55.35 per SM, which pure CTA-major order ran as 4 passes on 68 SMs, 3 on 80.
    const int64_t n_slots = (n_groups + 31) >> 5;
    const int32_t wpc = blockDim.x >> 5, wib = threadIdx.x >> 5;
    const int64_t per_pass = (int64_t)gridDim.x * wpc;
    const int64_t full = n_slots / per_pass;
    for (int64_t pass = 0; pass <= full; ++pass) {
        const int64_t slot = pass < full ? (pass * gridDim.x + blockIdx.x) * wpc + wib
                                         : full * per_pass + (int64_t)wib * gridDim.x + blockIdx.x;
        if (slot >= n_slots) break;                                      // warp-uniform
        int64_t g = slot * 32 + (threadIdx.x & 31);
        const bool valid = g < n_groups;
        if (!valid) {
            if (!Stepper::kCollective) break;
            g = n_groups - 1;                                            // keeps the warp whole: computes, stores nothing
        }
        const int64_t i0 = g * VEC;
        uint32_t s[4] = { 0, 0, 0, 0 };
        if (VEC == 4) {
            const uint4 v = reinterpret_cast<const uint4*>(a.state)[g];
            s[0] = v.x; s[1] = v.y; s[2] = v.z; s[3] = v.w;
        } else {
            s[0] = a.state[i0];
        }
        uint32_t t_in = 0;
#pragma unroll
        for (int e = 0; e < VEC; ++e) t_in += S.timestep(s[e]);
        S.template enter<VEC>(s);                                        // (rules steppers: words -> packed bytes)
        uint32_t acc_d = 0, acc_t = 0, p_done = 0, p_trunc = 0;
        int32_t p_net = 0;
        int32_t* op = (a.obs && valid) ? a.obs + i0 : nullptr;
        float* rp = (a.reward && valid) ? a.reward + i0 : nullptr;
        uint8_t* fp = (a.flags && valid) ? a.flags + i0 : nullptr;
        // contract v2: the thread's VEC == 4 envs are one aligned group of the global ids (the dispatcher sends
        // env_id_base % 4 != 0 to the VEC == 1 kernels, which pick their env's word of the group's call)
        const uint64_t env0 = a.env_id_base + (uint64_t)i0;
        const uint64_t grp = env0 >> 2;
        const uint32_t grp_lo = (uint32_t)grp, grp_hi = (uint32_t)(grp >> 32), widx = (uint32_t)env0 & 3u;
        // The words do not depend on the state: the four Philox calls of a block of four steps are issued together
        // (four independent 10-round chains in flight), then the four steps consume them -- one call per step in
        // program order put the chain's latency in front of every table look-up (measured: -8 % at 2^20 envs).
        uint64_t step = a.step0;
        constexpr int PB = stepper_block<Stepper>::value;             // steps whose Philox calls are issued together
        for (int32_t kb = 0; kb < K; kb += PB, step += PB) {
            uint32_t w[PB][4];
#pragma unroll
            for (int j = 0; j < PB; ++j) {
                const uint64_t sj = step + (uint64_t)j;
                // <true>: one IMAD.WIDE per product.  The round-1 build of this loop got that from the compiler's own fusion of
                // mul.hi + mul.lo; with the per-step counters it stopped fusing (85 IMAD.HI + 85 IMAD in the 4-step body
                // instead of 45 IMAD.WIDE: +5 instructions per env-step), so it is asked for explicitly
                philox4x32_10_rk<SOCCER_K2_PHILOX_WIDE != 0>(grp_lo, grp_hi, (uint32_t)sj, (uint32_t)(sj >> 32), a.rk, w[j]);
            }
#pragma unroll
            for (int j = 0; j < PB; ++j) {
                if (kb + j >= K) continue;                              // warp-uniform
                uint32_t word[4], oo[4], rr[4], fw;
                if (VEC == 4) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) word[e] = w[j][e];
                } else {
                    word[0] = widx == 0 ? w[j][0] : (widx == 1 ? w[j][1] : (widx == 2 ? w[j][2] : w[j][3]));
                }
                S.template step<VEC>(s, word, oo, rr, fw, p_net, flip);
                acc_d += fw & 0x01010101u;
                acc_t += (fw >> 1) & ~fw & 0x01010101u;                 // truncated WITHOUT a goal
                const bool with_streams = STREAMS && (!Stepper::kCollective || valid);
                if (VEC == 4) {
                    if (with_streams || op) { st_stream(reinterpret_cast<uint4*>(op), make_uint4(oo[0], oo[1], oo[2], oo[3])); op += a.n; }
                    if (with_streams || rp) { st_stream(reinterpret_cast<uint4*>(rp), make_uint4(rr[0], rr[1], rr[2], rr[3])); rp += a.n; }
                    if (with_streams || fp) { st_stream(reinterpret_cast<uint32_t*>(fp), fw); fp += a.n; }
                } else {
                    if (with_streams || op) { *op = (int32_t)oo[0]; op += a.n; }
                    if (with_streams || rp) { *rp = __uint_as_float(rr[0]); rp += a.n; }
                    if (with_streams || fp) { *fp = (uint8_t)fw; fp += a.n; }
                }
            }
            if ((kb & 63) == 64 - PB) {                                 // bytes hold at most 64 counts
                p_done = __dp4a(acc_d, 0x01010101u, p_done); p_trunc = __dp4a(acc_t, 0x01010101u, p_trunc);
                acc_d = acc_t = 0;
            }
        }
        p_done = __dp4a(acc_d, 0x01010101u, p_done); p_trunc = __dp4a(acc_t, 0x01010101u, p_trunc);
        S.template leave<VEC>(s);
        uint32_t t_out = 0;
#pragma unroll
        for (int e = 0; e < VEC; ++e) t_out += S.timestep(s[e]);
        if (valid) {
            c_done += p_done; c_trunc += p_trunc; c_net += p_net;
            c_len += (unsigned long long)t_in + (unsigned long long)K * VEC - t_out;
            c_steps += (unsigned long long)K * VEC;
            if (VEC == 4) reinterpret_cast<uint4*>(a.state)[g] = make_uint4(s[0], s[1], s[2], s[3]);
            else a.state[i0] = s[0];
        }
    }
    if (a.stats) {
        unsigned long long v[5] = { c_done, c_trunc, c_steps, c_len, (unsigned long long)c_net };
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const unsigned long long r = warp_sum_u64(v[j]);
            if ((threadIdx.x & 31) == 0 && r) atomicAdd(&blk->v[j], r);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            // episodes = goals + truncated-only (a goal on step 100 carries both flags and counts once)
            const long long d = (long long)blk->v[0], tr_only = (long long)blk->v[1], net = (long long)blk->v[4];
            atomicAdd(&a.stats[0], (unsigned long long)(d + tr_only));
            atomicAdd(&a.stats[1], (unsigned long long)((d + net) / 2));      // goals_A (reward +1)
            atomicAdd(&a.stats[2], (unsigned long long)((d - net) / 2));      // goals_B (reward -1)
            atomicAdd(&a.stats[3], (unsigned long long)tr_only);
            atomicAdd(&a.stats[4], blk->v[2]);
            atomicAdd(&a.stats[5], blk->v[3]);
        }
    }
}

// int8[nS] table policies: global -> shared (clamped so the column stays inside the row).  Called AFTER the grid
// dependency wait: a policy tensor may have been produced by the preceding kernel in the stream.
__device__ __forceinline__ void stage_policies(uint8_t* pa, uint8_t* pb, const int8_t* policy_a, const int8_t* policy_b, int nS)
{
    for (int i = threadIdx.x; i < nS; i += blockDim.x) {
        if (policy_a) pa[i] = (uint8_t)min(max((int)policy_a[i], 0), 4);
        if (policy_b) pb[i] = (uint8_t)min(max((int)policy_b[i], 0), 4);
    }
}

// Shared-memory image: the table, the 4 start observations (16 bytes), then - POLICY only - the two int8[nS]
// table policies (each padded to 16 bytes; an absent one is not copied and its address stays 0).
template <int VEC, bool STREAMS, bool POLICY, bool SLIP>
__global__ void __launch_bounds__(kRolloutThreads, 1)
k_rollout_table(const PitchDev P, const uint16_t* __restrict__ gtable, uint32_t table_bytes,
                const int8_t* __restrict__ policy_a, const int8_t* __restrict__ policy_b, const RolloutArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ BlkStats blk;
    // programmatic dependent launch: back-to-back rollouts stage their table (a constant input) while the previous
    // launch drains; policies, state, streams and statistics are only touched after pdl_wait()
    pdl_launch_dependents();
    if (threadIdx.x < 5) blk.v[threadIdx.x] = 0;
    TableStepper<POLICY, SLIP> S;
    S.pol_a = S.pol_b = 0;
    __shared__ __align__(16) double prt[kPrtDoubles];
    S.sc.prt = 0; S.sc.first_k = 0;
    if (SLIP) { slip_build_prt(prt, P); S.sc.prt = smem_u32(prt); S.sc.first_k = slip_first_k(P); }
    stage_table(smem_raw, gtable, table_bytes, &bar, P);      // ends with __syncthreads()
    S.c = make_ctx(smem_raw, table_bytes, P);
    wait_table(&bar);
    if (SLIP) { launder(S.c.tbl); launder(S.c.isd); launder(S.sc.prt); }   // table_step_slip's relaxed loads stay below the wait
    pdl_wait();
    if (POLICY) {
        const uint32_t pol_bytes = ((uint32_t)P.nS + 15u) & ~15u;
        uint8_t* pa = smem_raw + table_bytes + 16, *pb = pa + pol_bytes;
        stage_policies(pa, pb, policy_a, policy_b, P.nS);
        if (policy_a) S.pol_a = smem_u32(pa);
        if (policy_b) S.pol_b = smem_u32(pb);
        __syncthreads();
    }
    rollout_body<VEC, STREAMS>(S, a, &blk);
}

// slip_prob > 0.  Shared-memory image: [table][isd 16 B][policy a][policy b][look-up tables of the integer fast path].
#ifndef SOCCER_K2_SLIPI_THREADS
#define SOCCER_K2_SLIPI_THREADS SOCCER_ROLLOUT_THREADS
#endif
constexpr int kRolloutSlipThreads = SOCCER_K2_SLIPI_THREADS;
template <int VEC, bool STREAMS, bool POLICY>
__global__ void __launch_bounds__(kRolloutSlipThreads, 1)
k_rollout_table_slipi(const PitchDev P, const uint16_t* __restrict__ gtable, uint32_t table_bytes,
                      const SlipE E, const SlipDanger dg, const SlipBits lut_bits,
                      const int8_t* __restrict__ policy_a, const int8_t* __restrict__ policy_b, const RolloutArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ BlkStats blk;
    __shared__ __align__(16) double prt[kPrtDoubles];
    pdl_launch_dependents();
    if (threadIdx.x < 5) blk.v[threadIdx.x] = 0;
    const uint32_t pol_bytes = POLICY ? (((uint32_t)P.nS + 15u) & ~15u) : 0u;
    uint8_t* pa = smem_raw + table_bytes + 16, *pb = pa + pol_bytes;
    uint8_t* luts = pb + pol_bytes;
    slip_build_prt(prt, P);
    slip_int_build_luts(luts, E, P, dg, lut_bits.bits);
    stage_table(smem_raw, gtable, table_bytes, &bar, P);                               // ends with __syncthreads()
    TableSlipIntStepper<POLICY> S;
    S.c = make_ctx(smem_raw, table_bytes, P);
    S.sc.prt = smem_u32(prt); S.sc.first_k = slip_first_k(P);
    S.sf = slip_int_ctx(luts, lut_bits);
    S.pol_a = S.pol_b = 0;
    wait_table(&bar);
    launder(S.c.tbl); launder(S.c.isd); launder(S.sc.prt);
    launder(S.sf.base);
    pdl_wait();
    if (POLICY) {
        stage_policies(pa, pb, policy_a, policy_b, P.nS);
        if (policy_a) S.pol_a = smem_u32(pa);
        if (policy_b) S.pol_b = smem_u32(pb);
        __syncthreads();
        launder(S.pol_a); launder(S.pol_b);
    }
    rollout_body<VEC, STREAMS>(S, a, &blk);
}

// ---- K2 for mid-size pitches: the step table sharded over the shared memory of a thread-block CLUSTER.
// A 7x5 table is 476 KB, a 9x5 one 792 KB: more than one SM's 227 KB, but a cluster of 4 / 8 CTAs holds it once, each
// CTA one 128 KB slice (byte offset >> 17 = owner rank), and every CTA reads any entry through distributed shared
// memory (mapa + ld.shared::cluster).  Entries keep the 16-bit format, so nS <= 4096 (7x5, 8x5, 9x5, 7x6 ...).
constexpr uint32_t kClusterSliceLog2 = 17;
constexpr uint32_t kClusterSliceBytes = 1u << kClusterSliceLog2;
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
struct ClusterTableStepper {
    static constexpr bool kCollective = false, kHasPolicy = false;
    uint32_t slice;             // shared::cta address of this CTA's slice (the same offset in every CTA of the cluster)
    uint32_t isd, last;
    __device__ __forceinline__ uint32_t timestep(uint32_t s) const { return s >> 16; }
    template <int VEC> __device__ __forceinline__ void enter(uint32_t*) const {}
    template <int VEC> __device__ __forceinline__ void leave(uint32_t*) const {}
    template <int VEC>
    __device__ __forceinline__ void step(uint32_t* s, const uint32_t* word, uint32_t* oo, uint32_t* rr,
                                         uint32_t& fw, int32_t& net, bool) const
    {
        uint32_t ff[4] = { 0, 0, 0, 0 };
        int32_t e[4];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const uint32_t off = min((s[i] & 0xFFFFu) * 100u + philox_jr(word[i]), last) * 2u;
            uint32_t raddr;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(slice + (off & (kClusterSliceBytes - 1u))), "r"(off >> kClusterSliceLog2));
            asm volatile("ld.shared::cluster.s16 %0, [%1];" : "=r"(e[i]) : "r"(raddr));
        }
        const TblCtx c = { 0u, isd, last };
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const TblOut o = table_finish(c, s[i], e[i], word[i] & 0xCu);
            s[i] = o.state; oo[i] = o.obs; rr[i] = __float_as_uint((float)o.rew_i); ff[i] = o.flags;
            net += o.rew_i;
        }
        fw = VEC == 4 ? pack4(ff[0], ff[1], ff[2], ff[3]) : ff[0];
    }
};

template <int VEC, bool STREAMS>
__global__ void __launch_bounds__(kRolloutThreads, 1)
k_rollout_table_cluster(const PitchDev P, const uint16_t* __restrict__ gtable, uint32_t table_bytes, const RolloutArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];     // [slice 128 KB][isd 16 B]
    __shared__ __align__(8) uint64_t bar;
    __shared__ BlkStats blk;
    if (threadIdx.x < 5) blk.v[threadIdx.x] = 0;
    const uint32_t rank = cluster_ctarank();
    const uint32_t lo = rank * kClusterSliceBytes;
    const uint32_t mine = lo < table_bytes ? min(kClusterSliceBytes, table_bytes - lo) : 0u;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&bar, mine);
        const uint32_t chunk = 16384;
        for (uint32_t off = 0; off < mine; off += chunk)
            tma_load_1d(smem_raw + off, reinterpret_cast<const uint8_t*>(gtable) + lo + off, min(chunk, mine - off), &bar);
    }
    if (threadIdx.x < 4) reinterpret_cast<int32_t*>(smem_raw + kClusterSliceBytes)[threadIdx.x] = P.isd_obs[threadIdx.x];
    __syncthreads();
    wait_table(&bar);
    cluster_sync_all();                                      // every slice of the cluster is in place
    ClusterTableStepper S;
    S.slice = smem_u32(smem_raw); S.isd = S.slice + kClusterSliceBytes; S.last = (uint32_t)P.nS * 100u - 1u;
    rollout_body<VEC, STREAMS>(S, a, &blk);
    cluster_sync_all();                                      // no CTA leaves while a neighbour may still read its slice
}

// (forcing 3 resident CTAs - 80 registers - measured 283 vs 291 G env-steps/s: the kernel is issue-bound)
template <int VEC, bool STREAMS, bool SLIP>
__global__ void __launch_bounds__(kThreads)
k_rollout(const PitchDev P, const int8_t* __restrict__ policy_a, const int8_t* __restrict__ policy_b, const RolloutArgs a)
{
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    __shared__ __align__(16) double prt[kPrtDoubles];
    __shared__ BlkStats blk;
    __shared__ __align__(4) uint8_t jlut[100];
    if (SLIP) slip_build_prt(prt, P);
    build_cand_lut(lut, P);
    if (threadIdx.x < 5) blk.v[threadIdx.x] = 0;
    if (threadIdx.x < 100) {
        const uint32_t ja = threadIdx.x >> 2, aa = ja / 5u;
        jlut[threadIdx.x] = (uint8_t)(aa | ((ja - aa * 5u) << 3) | ((threadIdx.x & 3u) << 6));
    }
    __syncthreads();
    const SlipCtx sc = { (uint32_t)__cvta_generic_to_shared(prt), SLIP ? slip_first_k(P) : 0u };
    const RulesStepper<SLIP> S = { P, lut, make_isd4(P), policy_a, policy_b, sc, (uint32_t)__cvta_generic_to_shared(jlut) };
    rollout_body<VEC, STREAMS>(S, a, &blk);
}

// slip_prob > 0, uniform policy, 4 envs per thread (RulesSlipIntStepper).  The 10-bit bucket table + slot thresholds take
// 5.6 KB of shared memory next to the 4 KB candidate table.
constexpr int kRulesSlipLutBits = 10;
struct RulesSlipArgs { SlipE E; SlipDanger dg; int32_t use_int; };
template <bool STREAMS, bool POLICY>
__global__ void __launch_bounds__(kThreads, 2)
k_rollout_slipi(const PitchDev P, const RolloutArgs a, const RulesSlipArgs sa, const int8_t* __restrict__ policy_a,
                const int8_t* __restrict__ policy_b)
{
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    __shared__ __align__(16) double prt[kPrtDoubles];
    __shared__ __align__(16) uint8_t ilut[slip_int_lut_bytes(kRulesSlipLutBits)];
    __shared__ BlkStats blk;
    slip_build_prt(prt, P);
    slip_int_build_luts(ilut, sa.E, P, sa.dg, kRulesSlipLutBits, 1u, 16u);
    build_cand_lut(lut, P);
    if (threadIdx.x < 5) blk.v[threadIdx.x] = 0;
    __syncthreads();
    const SlipCtx sc = { (uint32_t)__cvta_generic_to_shared(prt), slip_first_k(P) };
    const RulesSlipIntStepper<POLICY> S = { P, lut, make_isd4(P), sc, slip_int_ctx(ilut, slip_bits(kRulesSlipLutBits)), policy_a, policy_b };
    rollout_body<4, STREAMS>(S, a, &blk);
}

// Measurement probe, not part of the game: K2's memory traffic (state read and written once, K x
// (16 + 16 + 4) bytes per 4-env group streamed out with the same stores, launch shape and slot order)
// with no Philox and no game logic.  Its throughput is the practical HBM ceiling for K2's write-only mix.
template <int MODE> __device__ __forceinline__ void st_probe(uint4* p, uint4 v)
{
    if (MODE == 1) *p = v;
    else if (MODE == 2) asm volatile("st.global.wt.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (MODE == 3) asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else __stcs(p, v);
}
template <int MODE> __device__ __forceinline__ void st_probe(uint32_t* p, uint32_t v)
{
    if (MODE == 1) *p = v;
    else if (MODE == 2) asm volatile("st.global.wt.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
    else if (MODE == 3) asm volatile("st.global.cg.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
    else __stcs(p, v);
}

// MODE 0: K2's stores (st.global.cs) and slot order, 1: default cache policy, 2: write-through, 3: .cg,
// 4: pure CTA-major slot order, 5: pure warp round-robin slot order
template <int MODE>
__global__ void __launch_bounds__(kRolloutThreads, 1)
k_rollout_probe(const RolloutArgs a)
{
    constexpr int G = 1;
    const int64_t n_groups = a.n / (4 * G);
    const int64_t n_slots = (n_groups + 31) >> 5;
    const int32_t wpc = blockDim.x >> 5, wib = threadIdx.x >> 5;
    const int64_t per_pass = (int64_t)gridDim.x * wpc;
    const int64_t full = MODE == 4 ? (n_slots + per_pass - 1) / per_pass : MODE == 5 ? 0 : n_slots / per_pass;
    const int64_t last = MODE == 5 ? (n_slots + per_pass - 1) / per_pass : full;
    for (int64_t pass = 0; pass <= last; ++pass) {
        const int64_t slot = pass < full ? (pass * gridDim.x + blockIdx.x) * wpc + wib
                                         : pass * per_pass + (int64_t)wib * gridDim.x + blockIdx.x;
        if (slot >= n_slots) break;
        const int64_t g = slot * 32 + (threadIdx.x & 31);
        if (g >= n_groups) break;
        uint4 v = reinterpret_cast<const uint4*>(a.state)[g * G];
        int32_t* op = a.obs + g * 4 * G;
        float* rp = a.reward + g * 4 * G;
        uint8_t* fp = a.flags + g * 4 * G;
        for (int32_t k = 0; k < a.K; ++k) {
            v.x += v.y; v.y ^= v.z; v.z += v.w; v.w += 0x9E3779B9u;
#pragma unroll
            for (int q = 0; q < G; ++q) {
                st_probe<MODE>(reinterpret_cast<uint4*>(op) + q, v);
                st_probe<MODE>(reinterpret_cast<uint4*>(rp) + q, make_uint4(v.y, v.x, v.w, v.z));
                st_probe<MODE>(reinterpret_cast<uint32_t*>(fp) + q, v.x ^ v.w);
            }
            op += a.n; rp += a.n; fp += a.n;
        }
#pragma unroll
        for (int q = 0; q < G; ++q) reinterpret_cast<uint4*>(a.state)[g * G + q] = v;
    }
}

} // namespace soccer

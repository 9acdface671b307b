// soccer_kernels.cu -- sm_100a kernels + the C ABI of include/soccer_b200.h.
//
//   K1  k_step_fast / k_step_generic   one lock-step step() of n envs, auto-reset fused
//   K2  k_rollout                      K steps with state in registers, Philox, on-device policy
//   K3  k_sweep                        exhaustive (state, joint action, slip combo, slot) table
//   K4  k_reset / k_set_state / k_get_obs
//       k_dense                        Pmat / Rmat in the reference's accumulation order
//
// All of them are HBM-streaming or issue-bound integer kernels: there is no dense contraction,
// so no tensor-core path.  SIM = /root/reference/gym_soccer/envs/soccer_simultaneous_env.py.
#include "../../include/soccer_b200.h"
#include "soccer_rules.cuh"
#include "soccer_rules4.cuh"
#include "soccer_rollout.cuh"
#include "soccer_replay.cuh"
#include "soccer_planner.cuh"

#include <cuda_runtime.h>
#include <sys/mman.h>
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include <unordered_set>

using namespace soccer;

namespace soccer {



// ------------------------------------------------------------------ host: pitch
bool pitch_ok(const soccer_pitch* p)
{
    return p && p->width >= 5 && p->height >= 4 && p->height <= 16 &&
           (int64_t)p->width * p->height <= kMaxField && p->slip_prob >= 0.0 && p->slip_prob <= 1.0;
}

void goal_rows_of(int H, int rows[3], int* n)
{
    if (H % 2 == 0) { *n = 2; rows[0] = (H - 1) / 2; rows[1] = H / 2; rows[2] = -1; }   // SIM:60
    else { *n = 3; rows[0] = H / 2 - 1; rows[1] = H / 2; rows[2] = H / 2 + 1; }
}

uint32_t field_code(int w, int row, int col_padded) { return (uint32_t)(row * w + col_padded - 1); }

int host_obs_index(int F, uint32_t a, uint32_t b, uint32_t p)
{
    return 1 + 2 * ((int)a * (F - 1) + (int)b - (b > a ? 1 : 0)) + (int)p;
}

int fill_info(const soccer_pitch* p, soccer_pitch_info* o)
{
    if (!p || !o) return SOCCER_EINVAL;
    if (!pitch_ok(p)) return SOCCER_EPITCH;
    const int w = p->width, H = p->height, F = w * H, W = w + 2;
    o->padded_width = W; o->height = H; o->n_field_cells = F;
    o->nS = 1 + 2 * F * (F - 1); o->nA = 5;
    goal_rows_of(H, o->goal_rows, &o->n_goal_rows);
    // SIM:146-165
    const int col_a = 2, col_b = W - 3;
    int tup[4][5]; int n = 0;
    if (o->n_goal_rows % 2 == 0) {
        const int mid = o->n_goal_rows / 2;
        const int opt[2] = { o->goal_rows[mid - 1], o->goal_rows[mid] };
        for (int i = 0; i < 2; ++i)
            for (int poss = 0; poss < 2; ++poss) {
                const int ra = opt[i], rb = (ra == opt[0]) ? opt[1] : opt[0];
                const int t[5] = { ra, col_a, rb, col_b, poss };
                for (int k = 0; k < 5; ++k) tup[n][k] = t[k];
                ++n;
            }
    } else {
        const int mid = o->goal_rows[o->n_goal_rows / 2];
        for (int poss = 0; poss < 2; ++poss) {
            const int t[5] = { mid, col_a, mid, col_b, poss };
            for (int k = 0; k < 5; ++k) tup[n][k] = t[k];
            ++n;
        }
    }
    o->n_isd = n;
    for (int i = 0; i < 4; ++i) {
        const int j = i < n ? i : n - 1;
        for (int k = 0; k < 5; ++k) o->isd_tuple[i][k] = tup[j][k];
        const uint32_t a = field_code(w, tup[j][0], tup[j][1]), b = field_code(w, tup[j][2], tup[j][3]);
        o->isd_state[i] = a | (b << 8) | ((uint32_t)tup[j][4] << 24);
        o->isd_obs[i] = host_obs_index(F, a, b, (uint32_t)tup[j][4]);
    }
    // SIM:209-223: identical fp64 expressions, evaluated in the same order
    const double s = p->slip_prob;
    o->slip_combo_prob[0] = (1 - s) * (1 - s);
    o->slip_combo_prob[1] = (1 - s) * s * 0.5;
    o->slip_combo_prob[2] = (1 - s) * s * 0.5;
    o->slip_combo_prob[3] = s * (1 - s) * 0.5;
    o->slip_combo_prob[4] = s * (1 - s) * 0.5;
    for (int c = 5; c < 9; ++c) o->slip_combo_prob[c] = s * s * 0.25;
    return SOCCER_OK;
}

int make_pitch_dev(const soccer_pitch* p, PitchDev* d)
{
    soccer_pitch_info info;
    const int rc = fill_info(p, &info);
    if (rc) return rc;
    d->w = p->width; d->H = p->height; d->F = info.n_field_cells; d->Fm1 = d->F - 1; d->nS = info.nS;
    d->nSm1 = (uint32_t)info.nS - 1u; d->tlast = (uint32_t)info.nS * 100u - 1u;
    d->goal_row_mask = 0;
    for (int i = 0; i < info.n_goal_rows; ++i) d->goal_row_mask |= 1u << info.goal_rows[i];
    // injected 2-bit draw r -> isd index floor(n_isd * (r+0.5)/4): r for 4 starts, r>>1 for 2
    for (int r = 0; r < 4; ++r) {
        const int idx = info.n_isd == 4 ? r : (r >> 1);
        d->isd_state[r] = info.isd_state[idx];
        d->isd_obs[r] = info.isd_obs[idx];
    }
    d->isd_d1 = d->isd_state[1] - d->isd_state[0];
    d->isd_d2 = d->isd_state[2] - d->isd_state[0];
    d->obs_d1 = d->isd_obs[1] - d->isd_obs[0];
    d->obs_d2 = d->isd_obs[2] - d->isd_obs[0];
    if (d->isd_state[3] != d->isd_state[0] + d->isd_d1 + d->isd_d2 ||
        d->isd_obs[3] != d->isd_obs[0] + d->obs_d1 + d->obs_d2 ||
        d->isd_state[3] != (d->isd_state[0] ^ d->isd_state[1] ^ d->isd_state[2]) ||     // bytewise XOR form (rules4)
        d->obs_d1 < 0 || d->obs_d2 < 0 || d->isd_obs[3] > 0xFFFF)
        return SOCCER_EPITCH;   // cannot happen for the reference's start distributions
    for (int c = 0; c < 9; ++c) d->mp[c] = info.slip_combo_prob[c];
    d->slip = p->slip_prob != 0.0;
    return SOCCER_OK;
}

int sm_count()
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
    return sms;
}

template <typename K>
int resident_blocks(K kernel)
{
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kThreads, 0) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}

inline int launch_status() { return (int)cudaGetLastError(); }

// Launch with programmatic dependent launch allowed: the kernel's prologue (shared-memory table
// fill by TMA, candidate-table build) may overlap the tail of the previous kernel in the stream;
// the kernel itself orders its first global access after griddepcontrol.wait.
template <typename... KArgs, typename... Args>
int launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }


// The single-agent modes (SIM:187-188, 243-244) on the byte-parallel rules kernel: the folded player's action byte of each
// of the four envs is its table policy (int8[nS] in global memory, L1-resident) at the env's CURRENT observation.
__device__ __forceinline__ uint32_t policy_actions4(const PitchDev& P, const int8_t* __restrict__ policy, const uint32_t sv[4])
{
    uint32_t act[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const uint32_t cur = (uint32_t)obs_index(P, sv[e] & 0xFFu, (sv[e] >> 8) & 0xFFu, (sv[e] >> 24) & 1u);
        act[e] = (uint32_t)(uint8_t)policy[min(cur, P.nSm1)];          // the clamp keeps a corrupt state word inside the table
    }
    return pack4(act[0], act[1], act[2], act[3]);
}

template <bool RESET_OBS, bool NARROW = false, bool POLICY = false>
__device__ __forceinline__ void step_group(const PitchDev& P, const Isd4& I, const uint8_t* lut, const Group4& x,
                                           int64_t g, uint4* st, uint4* obs, uint4* rew, uint32_t* flg, uint4* rob,
                                           bool want_stats, K1Stats& acc, const int8_t* __restrict__ policy_a = nullptr,
                                           const int8_t* __restrict__ policy_b = nullptr)
{
    const uint32_t sv[4] = { x.s.x, x.s.y, x.s.z, x.s.w };
    Step4 o;
    uint32_t A4 = x.a, B4 = x.b;
    if (POLICY && policy_a) A4 = policy_actions4(P, policy_a, sv);
    if (POLICY && policy_b) B4 = policy_actions4(P, policy_b, sv);
    step4_noslip<RESET_OBS>(P, I, lut, sv, A4, B4, x.r, o);       // byte-parallel over the 4 envs
    if (POLICY && policy_a) {                                     // the return agent is player_b: its reward is -A's (SIM:243-244)
#pragma unroll
        for (int e = 0; e < 4; ++e) o.rew[e] = __float_as_uint((float)(-(int)(signed char)(o.rew4 >> (8 * e))));
    }
    if (want_stats) {
        // timesteps live in byte 2 of the CELL-layout words
        const uint32_t ti = (sv[0] & 0xFF0000u) + (sv[1] & 0xFF0000u) + (sv[2] & 0xFF0000u) + (sv[3] & 0xFF0000u);
        const uint32_t to = (o.s[0] & 0xFF0000u) + (o.s[1] & 0xFF0000u) + (o.s[2] & 0xFF0000u) + (o.s[3] & 0xFF0000u);
        acc.add4(o.flags4, o.rew_sum, ti, to);
    }
    st_keep(st + g, make_uint4(o.s[0], o.s[1], o.s[2], o.s[3]));
    if (NARROW) {                                                 // uint16 obs, int8 reward (soccer_step_narrow)
        st_stream(reinterpret_cast<uint2*>(obs) + g, make_uint2(o.obs[0] | (o.obs[1] << 16), o.obs[2] | (o.obs[3] << 16)));
        st_stream(reinterpret_cast<uint32_t*>(rew) + g, o.rew4);
    } else {
        st_stream(obs + g, make_uint4(o.obs[0], o.obs[1], o.obs[2], o.obs[3]));
        st_stream(rew + g, make_uint4(o.rew[0], o.rew[1], o.rew[2], o.rew[3]));
    }
    st_stream(flg + g, o.flags4);
    if (RESET_OBS) st_stream(rob + g, make_uint4(o.robs[0], o.robs[1], o.robs[2], o.robs[3]));
}

template <bool RESET_OBS, bool PHILOX = false, bool NARROW = false, bool POLICY = false>
__global__ void __launch_bounds__(kThreads)
k_step_fast(const PitchDev P, uint32_t* __restrict__ state, const uint8_t* __restrict__ act_a,
            const uint8_t* __restrict__ act_b, const uint8_t* __restrict__ rng, int32_t* __restrict__ obs,
            float* __restrict__ reward, uint8_t* __restrict__ flags, int32_t* __restrict__ reset_obs,
            int64_t n_groups, const PhiloxKey key, unsigned long long* __restrict__ stats,
            const int8_t* __restrict__ policy_a = nullptr, const int8_t* __restrict__ policy_b = nullptr)
{
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    __shared__ K1StatsBlk sblk;
    pdl_launch_dependents();
    k1_stats_init(&sblk);
    build_cand_lut(lut, P);
    const Isd4 I = make_isd4(P);
    pdl_wait();                  // everything above overlaps the previous kernel's tail
    const bool want_stats = stats != nullptr;
    K1Stats acc = {};

    uint4* st4 = reinterpret_cast<uint4*>(state);
    // a folded player's action stream does not exist: alias the other one (loaded, never used)
    const uint32_t* a4 = reinterpret_cast<const uint32_t*>(POLICY && !act_a ? act_b : act_a);
    const uint32_t* b4 = reinterpret_cast<const uint32_t*>(POLICY && !act_b ? act_a : act_b);
    const uint32_t* r4 = reinterpret_cast<const uint32_t*>(rng);
    uint4* o4 = reinterpret_cast<uint4*>(obs);
    uint4* w4 = reinterpret_cast<uint4*>(reward);
    uint32_t* f4 = reinterpret_cast<uint32_t*>(flags);
    uint4* q4 = reinterpret_cast<uint4*>(reset_obs);

    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += 2 * stride) {
        const int64_t g2 = g + stride;
        const bool two = g2 < n_groups;
        // PHILOX: the draw stream is not read (r4 aliases the action stream, its value is replaced)
        Group4 x0 = load_group(st4, a4, b4, PHILOX ? a4 : r4, g);
        Group4 x1 = x0;
        if (two) x1 = load_group(st4, a4, b4, PHILOX ? a4 : r4, g2);
        if (PHILOX) { x0.r = philox_rng8x4(key, g); if (two) x1.r = philox_rng8x4(key, g2); }
        step_group<RESET_OBS, NARROW, POLICY>(P, I, lut, x0, g, st4, o4, w4, f4, q4, want_stats, acc, policy_a, policy_b);
        if (two) step_group<RESET_OBS, NARROW, POLICY>(P, I, lut, x1, g2, st4, o4, w4, f4, q4, want_stats, acc, policy_a, policy_b);
    }
    if (want_stats) k1_stats_flush(acc, &sblk, stats, (unsigned long long)n_groups * 4ull);
}

// K1 for slip_prob > 0 on ANY pitch (rules inline): 32-bit draws -- the injected rng32 stream or Philox -- decided by
// integer thresholds and resolved byte-parallel (step4_slip_int); 24 B / env-step (19 with Philox draws).
#ifndef SOCCER_FAST_SLIP_MINBLOCKS
#define SOCCER_FAST_SLIP_MINBLOCKS 2
#endif
template <bool RESET_OBS, bool PHILOX, bool POLICY = false>
__global__ void __launch_bounds__(kThreads, SOCCER_FAST_SLIP_MINBLOCKS)
k_step_fast_slip(const PitchDev P, const RulesSlipArgs sa, uint32_t* __restrict__ state, const uint8_t* __restrict__ act_a,
                 const uint8_t* __restrict__ act_b, const uint8_t* __restrict__ rng, const uint32_t* __restrict__ draw,
                 int32_t* __restrict__ obs, float* __restrict__ reward, uint8_t* __restrict__ flags,
                 int32_t* __restrict__ reset_obs, int64_t n_groups, const PhiloxKey key,
                 const int8_t* __restrict__ policy_a = nullptr, const int8_t* __restrict__ policy_b = nullptr)
{
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    __shared__ __align__(16) double prt[kPrtDoubles];
    __shared__ __align__(16) uint8_t ilut[slip_int_lut_bytes(kRulesSlipLutBits)];
    slip_build_prt(prt, P);
    slip_int_build_luts(ilut, sa.E, P, sa.dg, kRulesSlipLutBits, 1u, 16u);
    build_cand_lut(lut, P);                      // ends with __syncthreads()
    const Isd4 I = make_isd4(P);
    const SlipCtx sc = { (uint32_t)__cvta_generic_to_shared(prt), slip_first_k(P) };
    const SlipInt fi = slip_int_ctx(ilut, slip_bits(kRulesSlipLutBits));
    uint4* st4 = reinterpret_cast<uint4*>(state);
    // a folded player's action stream does not exist: alias the other one (loaded, never used)
    const uint32_t* a4 = reinterpret_cast<const uint32_t*>(POLICY && !act_a ? act_b : act_a);
    const uint32_t* b4 = reinterpret_cast<const uint32_t*>(POLICY && !act_b ? act_a : act_b);
    const uint32_t* r4 = reinterpret_cast<const uint32_t*>(rng);
    const uint4* d4 = reinterpret_cast<const uint4*>(draw);
    uint4* o4 = reinterpret_cast<uint4*>(obs);
    uint4* w4 = reinterpret_cast<uint4*>(reward);
    uint32_t* f4 = reinterpret_cast<uint32_t*>(flags);
    uint4* q4 = reinterpret_cast<uint4*>(reset_obs);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool one = g < n_groups;
    GroupS x = {};
    if (one) x = load_group_slip<PHILOX>(st4, a4, b4, r4, d4, g);
    while (one) {
        const int64_t gn = g + stride;
        const bool n_one = gn < n_groups;
        GroupS y = x;
        if (n_one) y = load_group_slip<PHILOX>(st4, a4, b4, r4, d4, gn);       // register prefetch of the next group
        if (PHILOX) {
            uint32_t w[4];
            philox_words4(key, g, w);
            x.d = make_uint4(philox_r32(w[0]), philox_r32(w[1]), philox_r32(w[2]), philox_r32(w[3]));
            x.r = (w[0] & 0xCu) | ((w[1] & 0xCu) << 8) | ((w[2] & 0xCu) << 16) | ((w[3] & 0xCu) << 24);
        }
        const uint32_t sv[4] = { x.s.x, x.s.y, x.s.z, x.s.w };
        const uint32_t r32[4] = { x.d.x, x.d.y, x.d.z, x.d.w };
        Step4 o;
        Soa4 so;
        uint32_t A4 = x.a, B4 = x.b;
        if (POLICY && policy_a) A4 = policy_actions4(P, policy_a, sv);      // SIM:187-188 (single-agent modes)
        if (POLICY && policy_b) B4 = policy_actions4(P, policy_b, sv);
        step4_slip_int<RESET_OBS, true>(P, I, lut, fi, sc, soa4_from_words(sv), ((A4 & 0x07070707u) << 3) | (B4 & 0x07070707u),
                                        r32, x.r, o, so);
        soa4_to_words(so, o.s);
        if (POLICY && policy_a) {                                           // the return agent is player_b (SIM:243-244)
#pragma unroll
            for (int e = 0; e < 4; ++e) o.rew[e] = __float_as_uint((float)(-(int)(signed char)(o.rew4 >> (8 * e))));
        }
        st_keep(st4 + g, make_uint4(o.s[0], o.s[1], o.s[2], o.s[3]));
        st_stream(o4 + g, make_uint4(o.obs[0], o.obs[1], o.obs[2], o.obs[3]));
        st_stream(w4 + g, make_uint4(o.rew[0], o.rew[1], o.rew[2], o.rew[3]));
        st_stream(f4 + g, o.flags4);
        if (RESET_OBS) st_stream(q4 + g, make_uint4(o.robs[0], o.robs[1], o.robs[2], o.robs[3]));
        x = y; g = gn; one = n_one;
    }
}

// ------------------------------------------------------------------ K1 generic path
// Every option of soccer_step_args, one env per thread (scalar but warp-coalesced accesses).
struct StepOpts {
    uint32_t* state; const uint8_t* act_a; const uint8_t* act_b; const uint8_t* rng8;
    const uint32_t* rng32; const double* rngf64; const int8_t* policy_a; const int8_t* policy_b;
    int32_t* obs; float* reward; uint8_t* flags; int32_t* reset_obs;
    int64_t n; int32_t auto_reset; int32_t use_philox; int32_t detail; uint64_t seed, step, env_id_base;
    int32_t narrow;      // obs points at uint16[n], reward at int8[n] (soccer_step_narrow)
    unsigned long long* stats;
};
__device__ __forceinline__ void put_obs(const StepOpts& o, int64_t i, int32_t v)
{
    if (o.narrow) reinterpret_cast<uint16_t*>(o.obs)[i] = (uint16_t)v; else o.obs[i] = v;
}
__device__ __forceinline__ void put_reward(const StepOpts& o, int64_t i, float v)
{
    if (o.narrow) reinterpret_cast<int8_t*>(o.reward)[i] = (int8_t)(int)v; else o.reward[i] = v;
}

__global__ void __launch_bounds__(kThreads) k_step_generic(const PitchDev P, const StepOpts o)
{
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    __shared__ __align__(16) double prt[kPrtDoubles];
    if (P.slip) slip_build_prt(prt, P);
    const SlipCtx sc = { (uint32_t)__cvta_generic_to_shared(prt), P.slip ? slip_first_k(P) : 0u };
    __shared__ K1StatsBlk sblk;
    k1_stats_init(&sblk);
    K1Stats acc = {};
    build_cand_lut(lut, P);                      // ends with __syncthreads(): prt visible as well
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < o.n; i += stride) {
        const uint32_t s = o.state[i];
        const uint32_t a = s & 0xFFu, b = (s >> 8) & 0xFFu, p = (s >> 24) & 1u;
        const bool terminal_in = ((a | b) & kGoalBit) != 0;
        if (s & kNeedsReset) {                       // the assert at SIM:376: leave the env alone
            if (o.obs) put_obs(o, i, terminal_in ? 0 : obs_index(P, a, b, p));
            if (o.reward) put_reward(o, i, 0.0f);
            if (o.flags) o.flags[i] = 0xFFu;
            if (o.reset_obs) o.reset_obs[i] = terminal_in ? 0 : obs_index(P, a, b, p);
            continue;
        }
        uint32_t rng = 0, pw = 0;
        if (o.use_philox) { pw = philox_word(o.seed, o.env_id_base + (uint64_t)i, o.step); rng = philox_rng8(pw); }
        else if (o.rng8) rng = o.rng8[i];
        const int32_t cur = terminal_in ? 0 : obs_index(P, a, b, p);
        // SIM:187-188: a folded player's action is its table policy at the current observation
        const uint32_t aa = o.policy_a ? (uint32_t)o.policy_a[cur] : (uint32_t)o.act_a[i];
        const uint32_t ab = o.policy_b ? (uint32_t)o.policy_b[cur] : (uint32_t)o.act_b[i];
        const bool flip = o.policy_a != nullptr;     // return agent is player_b (SIM:243-244)
        StepOut r;
        if (terminal_in) {
            // a goal tuple injected through `env.state = ...`: absorbing, done, reward 0 (SIM:235-236, 300-301)
            const uint32_t t1 = ((s >> 16) & 0xFFu) + 1u;
            const bool trunc = t1 >= (uint32_t)kMaxT;
            r.obs = 0; r.reward = flip ? -0.0f : 0.0f; r.flags = 1u | (trunc ? 2u : 0u);
            if (o.auto_reset) { r.state = P.isd_state[(rng >> 2) & 3u]; r.reset_obs = P.isd_obs[(rng >> 2) & 3u]; }
            else { r.state = (s & 0x0100FFFFu) | (t1 << 16) | kNeedsReset; r.reset_obs = 0; }
        } else if (P.slip) {
            double u;
            if (o.rngf64) u = o.rngf64[i];
            else if (o.rng32) u = ((double)o.rng32[i] + 0.5) * (1.0 / 4294967296.0);
            else u = u_from_rng32(philox_r32(pw));       // philox mode with slip: the word's 32-bit step draw
            r = o.auto_reset ? step_slip<true>(P, lut, sc, s, aa, ab, u, (rng >> 2) & 3u, flip)
                             : step_slip<false>(P, lut, sc, s, aa, ab, u, (rng >> 2) & 3u, flip);
        } else {
            r = o.auto_reset ? step_noslip<true>(P, lut, s, aa, ab, rng, flip)
                             : step_noslip<false>(P, lut, s, aa, ab, rng, flip);
        }
        o.state[i] = r.state;
        if (o.obs) put_obs(o, i, r.obs);
        if (o.reward) put_reward(o, i, r.reward);
        if (o.flags) o.flags[i] = (uint8_t)(o.detail ? r.flags : (r.flags & 3u));
        if (o.reset_obs) o.reset_obs[i] = r.reset_obs;
        // statistics by player A's reward sign (the streamed reward is flipped for a player_b env)
        const int32_t ri = (r.reward > 0.0f) - (r.reward < 0.0f);
        acc.add1(r.flags, flip ? -ri : ri, ((s >> 16) & 0xFFu) + 1u);
    }
    if (o.stats) k1_stats_flush(acc, &sblk, o.stats);
}

// ------------------------------------------------------------------ single-env speculation
// The single-env drop-in's step() is one launch + one wait per call, all latency.  Between two step() calls the env's
// state is known but the joint action and the draw are not -- so ONE launch, enqueued as soon as the state is known,
// steps the env for ALL 25 joint actions x 4 draw values (slip_prob == 0: the draw is 2 bits) and
// writes the 100 results into (pinned host) memory while the caller is still busy; step() then picks record
// (aa * 5 + ab) * 4 + floor(4u).  One thread per (joint action, draw); no shared memory, no look-up table: the two
// candidate cells are computed arithmetically.  Record = { next state word (no auto-reset: needs_reset is set like
// SIM:406), obs, reward bits, detail flags | seq << 8 }, written with ONE 128-bit store: the sequence number travels in
// the same PCIe write as the data it vouches for, so there is no fence, no barrier and no separate flag on the critical
// path -- the reader polls the last word of the record it wants.
// slip_prob > 0 (SLIP): the draw is a full fp64 number, so it cannot be enumerated -- the caller hands over the draw
// the env's generator WILL make (the host mirrors the generator one draw ahead, see the Python class) and the launch
// steps the 25 joint actions with exactly that u by the reference's cumulative walk (records r = 0 only).
template <bool SLIP>
__global__ void __launch_bounds__(128)
k_step_speculate(const PitchDev P, uint32_t s, const int8_t* __restrict__ policy_a, const int8_t* __restrict__ policy_b,
                 uint32_t* __restrict__ rec, uint32_t seq, double u)
{
    __shared__ __align__(16) uint8_t lut[SLIP ? kLutBytes : 16];
    __shared__ __align__(16) double prt[SLIP ? kPrtDoubles : 2];
    const uint32_t i = threadIdx.x;
    const uint32_t a = s & 0xFFu, b = (s >> 8) & 0xFFu, t = (s >> 16) & 0xFFu, p = (s >> 24) & 1u;
    if (SLIP) {
        // the walk reads the candidate table only in the rows of the two occupied cells: 32 entries
        if (i < 32u) {
            const uint32_t cell = i < 16u ? a : b, idx = i & 15u, move = idx & 7u;
            lut[cell * 16u + idx] = (uint8_t)next_cell_code(P, cell, idx >> 3, move > 4u ? 0u : move);
        }
        slip_build_prt(prt, P);
        __syncthreads();
    }
    if (i >= (SLIP ? 25u : 100u)) return;
    const uint32_t ja = SLIP ? i : i >> 2, r = SLIP ? 0u : i & 3u;
    uint32_t aa = ja / 5u, ab = ja - aa * 5u;
    const int32_t cur = obs_index(P, a, b, p);
    if (policy_a) aa = (uint32_t)policy_a[cur];        // SIM:187-188: the folded player's table policy
    if (policy_b) ab = (uint32_t)policy_b[cur];
    StepOut out;
    if (SLIP) {
        const SlipCtx sc = { (uint32_t)__cvta_generic_to_shared(prt), slip_first_k(P) };
        out = step_slip<false>(P, lut, sc, s, aa, ab, u, 0u, policy_a != nullptr);
    } else {
        const uint32_t ma = (aa & 7u) > 4u ? 0u : (aa & 7u), mb = (ab & 7u) > 4u ? 0u : (ab & 7u);   // as build_cand_lut
        const uint32_t na = next_cell_code(P, a, p ^ 1u, ma), nb = next_cell_code(P, b, p, mb);
        const Resolved o = resolve_cand(na, nb, a, b, p, aa == 0, ab == 0, r);
        out = finish_step<false, true>(P, o, t, 0u, 0u, policy_a != nullptr);
    }
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(rec + 4u * (ja * 4u + r)), "r"(out.state), "r"((uint32_t)out.obs),
                 "r"(__float_as_uint(out.reward)), "r"((out.flags & 0xFFu) | (seq << 8)) : "memory");
}

// ------------------------------------------------------------------ K3 sweep
__global__ void __launch_bounds__(kThreads)
k_sweep(const PitchDev P, int32_t n_combos, int32_t nS, uint8_t* __restrict__ n_out,
        uint32_t* __restrict__ next_state, int32_t* __restrict__ next_obs, int8_t* __restrict__ reward,
        uint8_t* __restrict__ done)
{
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    build_cand_lut(lut, P);
    const int64_t total = (int64_t)(nS - 1) * 25 * n_combos;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % n_combos);
        const int64_t sj = i / n_combos;
        const uint32_t ja = (uint32_t)(sj % 25), aa = ja / 5u, ab = ja % 5u;
        const int32_t s_obs = (int32_t)(sj / 25) + 1;
        const uint32_t st = obs_to_packed(P, s_obs);
        const uint32_t a = st & 0xFFu, b = (st >> 8) & 0xFFu, p = (st >> 24) & 1u;
        const int ca = combo_a(c), cb = combo_b(c);
        const uint32_t ma = ca == 0 ? aa : slip_move(aa, ca - 1);
        const uint32_t mb = cb == 0 ? ab : slip_move(ab, cb - 1);
        const Resolved o0 = resolve(lut, a, b, p, ma, mb, aa == 0, ab == 0, 0u);
        const uint32_t n = 1u << o0.nlog2;
        if (n_out) n_out[i] = (uint8_t)n;
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
            uint32_t ns = 0; int32_t no = 0; int8_t rw = 0; uint8_t dn = 0;
            if (k < n) {
                const Resolved o = resolve(lut, a, b, p, ma, mb, aa == 0, ab == 0, n == 2 ? (k << 1) : k);
                const StepOut f = finish_step<false>(P, o, 0u, 0u, 0u, false);
                ns = o.a | (o.b << 8) | (o.p << 24); no = f.obs; rw = (int8_t)f.reward; dn = (uint8_t)(f.flags & 1u);
            }
            const int64_t e = i * 4 + k;
            if (next_state) next_state[e] = ns;
            if (next_obs) next_obs[e] = no;
            if (reward) reward[e] = rw;
            if (done) done[e] = dn;
        }
    }
}

// ------------------------------------------------------------------ dense Pmat / Rmat
// One thread per (observation s, key): it alone owns column Pmat[s][:][key] and Rmat[s][key], so
// plain fp64 read-modify-writes in list order reproduce the reference's accumulation order
// (SIM:258-279) bit for bit.  Row s == 0 reproduces the reference quirk that every goal state
// adds its self-loop into P[0] / Pmat[0][0] (SIM:182-183, 262).
__global__ void __launch_bounds__(kThreads)
k_dense(const PitchDev P, int32_t nS, int32_t n_goal_states, const int8_t* __restrict__ policy_a,
        const int8_t* __restrict__ policy_b, double* __restrict__ Pmat, double* __restrict__ Rmat)
{
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    build_cand_lut(lut, P);
    const bool multi = !policy_a && !policy_b;
    const int nkeys = multi ? 25 : 5;
    const int64_t total = (int64_t)nS * nkeys;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int key = (int)(i % nkeys);
        const int32_t s_obs = (int32_t)(i / nkeys);
        double* col = Pmat + (int64_t)s_obs * nS * nkeys + key;     // Pmat[s][ns][key] = col[ns*nkeys]
        double racc = 0.0;                                          // SIM:260
        if (s_obs == 0) {
            double acc = 0.0;
            for (int gsi = 0; gsi < n_goal_states; ++gsi)
                for (int c = 0; c < 9; ++c)
                    if (P.mp[c] != 0.0) acc = __dadd_rn(acc, __dmul_rn(P.mp[c], 1.0));
            col[0] = acc;
            Rmat[i] = 0.0;
            continue;
        }
        const uint32_t st = obs_to_packed(P, s_obs);
        const uint32_t a = st & 0xFFu, b = (st >> 8) & 0xFFu, p = (st >> 24) & 1u;
        uint32_t aa, ab;
        if (multi) { aa = (uint32_t)key / 5u; ab = (uint32_t)key % 5u; }
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
                const Resolved o = resolve(lut, a, b, p, ma, mb, aa == 0, ab == 0, n == 2 ? (k << 1) : k);
                const StepOut f = finish_step<false>(P, o, 0u, 0u, 0u, flip);
                double* cell = col + (int64_t)f.obs * nkeys;
                *cell = __dadd_rn(*cell, pr);                                   // SIM:262
                racc = __dadd_rn(racc, __dmul_rn(pr, (double)f.reward));        // SIM:263
            }
        }
        Rmat[i] = racc;
    }
}

// ------------------------------------------------------------------ K4
__global__ void __launch_bounds__(kThreads)
k_reset(const PitchDev P, uint32_t* __restrict__ state, int32_t* __restrict__ obs_out,
        const uint8_t* __restrict__ rng8, const uint8_t* __restrict__ mask, int use_philox, uint64_t seed,
        uint64_t step, uint64_t env_id_base, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (mask && !mask[i]) continue;
        uint32_t r;
        if (use_philox) r = philox_rng8(philox_word(seed, env_id_base + (uint64_t)i, step));
        else r = rng8 ? rng8[i] : 0u;
        const uint32_t sel = (r >> 2) & 3u;
        state[i] = P.isd_state[sel];
        if (obs_out) obs_out[i] = P.isd_obs[sel];
    }
}

__global__ void __launch_bounds__(kThreads)
k_set_state(const PitchDev P, int32_t nS, uint32_t* __restrict__ state, const int32_t* __restrict__ obs_in,
            const int32_t* __restrict__ timestep_in, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int32_t o = obs_in[i];
        int32_t t = timestep_in ? timestep_in[i] : 0;
        t = t < 0 ? 0 : (t > kMaxT ? kMaxT : t);      // the byte-parallel kernels add 28 to t + 1 inside a byte
        if (o >= 1 && o < nS) state[i] = obs_to_packed(P, o) | ((uint32_t)t << 16);
        else state[i] = kNeedsReset | kGoalBit | ((uint32_t)t << 16);   // terminal / invalid: needs reset
    }
}

__global__ void __launch_bounds__(kThreads)
k_get_obs(const PitchDev P, const uint32_t* __restrict__ state, int32_t* __restrict__ obs_out, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t s = state[i];
        const uint32_t a = s & 0xFFu, b = (s >> 8) & 0xFFu, p = (s >> 24) & 1u;
        obs_out[i] = ((a | b) & kGoalBit) ? 0 : obs_index(P, a, b, p);
    }
}

int table_bytes_of(const PitchDev& P, int64_t* bytes)
{
    const int64_t nS = 1 + 2 * (int64_t)P.F * P.Fm1;
    if (nS - 1 > kMaxTableStates) return SOCCER_ETABLE;      // the table itself does not depend on slip_prob
    *bytes = (nS * 200 + 15) / 16 * 16;     // nS rows: row 0 is the absorbing terminal observation
    return SOCCER_OK;
}

// Measurement probe (bench / profiles only): K1's exact memory traffic -- same streams, same
// access pattern, same cache hints, same launch shape -- with the game logic replaced by a few
// XORs.  Its throughput is the practical HBM ceiling for K1's 7-bytes-read / 13-bytes-written mix,
// which a read:write = 1:1 copy benchmark does not measure.
// MODE 0: K1's order (groups g and g + grid * 1024 per iteration); 1: every CTA streams through ONE contiguous
// region of the batch; 2: pairs of adjacent 1024-group tiles dealt round-robin over the CTAs
template <int MODE>
__global__ void __launch_bounds__(kTableThreads, 1)
k_stream_mix_probe(uint32_t* __restrict__ state, const uint8_t* __restrict__ act_a, const uint8_t* __restrict__ act_b,
                   const uint8_t* __restrict__ rng, int32_t* __restrict__ obs, float* __restrict__ reward,
                   uint8_t* __restrict__ flags, int64_t n_groups)
{
    uint4* st4 = reinterpret_cast<uint4*>(state);
    const uint32_t* a4 = reinterpret_cast<const uint32_t*>(act_a);
    const uint32_t* b4 = reinterpret_cast<const uint32_t*>(act_b);
    const uint32_t* r4 = reinterpret_cast<const uint32_t*>(rng);
    uint4* o4 = reinterpret_cast<uint4*>(obs);
    uint4* w4 = reinterpret_cast<uint4*>(reward);
    uint32_t* f4 = reinterpret_cast<uint32_t*>(flags);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t g, gend, step2, second;
    if (MODE == 1) {
        const int64_t per = ((n_groups + gridDim.x - 1) / gridDim.x + 2 * blockDim.x - 1) / (2 * blockDim.x) * (2 * blockDim.x);
        g = (int64_t)blockIdx.x * per + threadIdx.x;
        gend = min(n_groups, (int64_t)(blockIdx.x + 1) * per);
        step2 = 2 * blockDim.x; second = blockDim.x;
    } else if (MODE == 2) {
        g = (int64_t)blockIdx.x * 2 * blockDim.x + threadIdx.x; gend = n_groups; step2 = 2 * stride; second = blockDim.x;
    } else {
        g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gend = n_groups; step2 = 2 * stride; second = stride;
    }
    for (; g < gend; g += step2) {
        const int64_t g2 = g + second;
        const bool two = g2 < gend;
        const Group4 x0 = load_group(st4, a4, b4, r4, g);
        Group4 x1 = x0;
        if (two) x1 = load_group(st4, a4, b4, r4, g2);
        const uint32_t m0 = x0.a ^ x0.b ^ x0.r, m1 = x1.a ^ x1.b ^ x1.r;
        st_keep(st4 + g, make_uint4(x0.s.x ^ m0, x0.s.y, x0.s.z, x0.s.w));
        st_stream(o4 + g, make_uint4(x0.s.y, m0, x0.s.z, x0.s.w));
        st_stream(w4 + g, make_uint4(m0, x0.s.x, x0.s.w, x0.s.z));
        st_stream(f4 + g, m0);
        if (two) {
            st_keep(st4 + g2, make_uint4(x1.s.x ^ m1, x1.s.y, x1.s.z, x1.s.w));
            st_stream(o4 + g2, make_uint4(x1.s.y, m1, x1.s.z, x1.s.w));
            st_stream(w4 + g2, make_uint4(m1, x1.s.x, x1.s.w, x1.s.z));
            st_stream(f4 + g2, m1);
        }
    }
}

// Episode statistics of ONE lock-step step from its flags (+ reward) streams: the K1 counterpart
// of K2's fused statistics.  16 envs per thread (one 128-bit flags load), warp redux, 4 atomics/CTA.
__global__ void __launch_bounds__(kThreads)
k_step_stats(const uint8_t* __restrict__ flags, const float* __restrict__ reward, int64_t n,
             unsigned long long* __restrict__ stats)
{
    __shared__ unsigned int blk[4];
    if (threadIdx.x < 4) blk[threadIdx.x] = 0;
    __syncthreads();
    uint32_t ended = 0, trunc_only = 0, ga = 0, gb = 0;
    const int64_t n16 = n / 16;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool vec = (reinterpret_cast<uintptr_t>(flags) & 15) == 0;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; vec && g < n16; g += stride) {
        const uint4 f = __ldcs(reinterpret_cast<const uint4*>(flags) + g);
        const uint32_t w[4] = { f.x, f.y, f.z, f.w };
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t d = w[k] & 0x01010101u, t = (w[k] >> 1) & ~w[k] & 0x01010101u;
            ended = __dp4a(d | t, 0x01010101u, ended);
            trunc_only = __dp4a(t, 0x01010101u, trunc_only);
            if (reward && d) {                              // goals are rare (~3 % of env-steps)
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if ((d >> (8 * e)) & 1u) {
                        const float r = reward[g * 16 + k * 4 + e];
                        ga += r > 0.0f; gb += r < 0.0f;
                    }
            }
        }
    }
    const int64_t tail0 = vec ? n16 * 16 : 0;
    for (int64_t i = tail0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t f = flags[i];
        ended += (f & 3u) != 0; trunc_only += (f & 3u) == 2u;
        if (reward && (f & 1u)) { const float r = reward[i]; ga += r > 0.0f; gb += r < 0.0f; }
    }
    uint32_t v[4] = { ended, ga, gb, trunc_only };
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t r = __reduce_add_sync(0xFFFFFFFFu, v[j]);
        if ((threadIdx.x & 31) == 0 && r) atomicAdd(&blk[j], r);
    }
    __syncthreads();
    if (threadIdx.x < 4 && blk[threadIdx.x]) atomicAdd(&stats[threadIdx.x], (unsigned long long)blk[threadIdx.x]);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&stats[4], (unsigned long long)n);
}

// ---- the one collective of the path, written over NVLink peer memory: sum all-reduce of the 6-entry statistics
// vector.  Every rank owns a symmetric buffer of 2 x 16 slots of 8 x uint64 (parity x source rank; 6 values, pad,
// flag) that all ranks can address (peer[] = the same buffer on every rank).  ONE warp: lane r stores this rank's
// vector into rank r's slot [parity][my rank], fences (system scope) and publishes the epoch in the slot's flag; lane r
// then waits for rank r's contribution to arrive in the LOCAL buffer, and a warp reduction leaves the sum in
// stats[0..5].  No host round trip, no second stream: the kernel is enqueued right behind the last step kernel.
// Two parities suffice: a rank can only enter call e + 2 after every rank has left call e (it needs their e + 1 data).
struct P2PStatsArgs { unsigned long long* peer[16]; int32_t rank, world; unsigned long long epoch; unsigned long long* stats; };
__global__ void __launch_bounds__(32)
k_stats_allreduce_p2p(const P2PStatsArgs a)
{
    const int lane = threadIdx.x;
    const unsigned long long par = a.epoch & 1ull;
    unsigned long long v[6] = { 0, 0, 0, 0, 0, 0 };
    if (lane < a.world) {
        unsigned long long* dst = a.peer[lane] + (par * 16ull + (unsigned long long)a.rank) * 8ull;
#pragma unroll
        for (int j = 0; j < 6; ++j) asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(dst + j), "l"(a.stats[j]) : "memory");
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(dst + 7), "l"(a.epoch) : "memory");
        const unsigned long long* src = a.peer[a.rank] + (par * 16ull + (unsigned long long)lane) * 8ull;
        const long long t0 = clock64();
        unsigned long long f;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(src + 7) : "memory");
            if (f != a.epoch && clock64() - t0 > (40ll << 30)) __trap();    // ~20 s: a peer died; do not hang the GPU
        } while (f != a.epoch);
#pragma unroll
        for (int j = 0; j < 6; ++j) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v[j]) : "l"(src + j) : "memory");
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xFFFFFFFFu, v[j], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 6; ++j) a.stats[j] = v[j];
    }
}

// E_k (sequential fp64 sums of the combination probabilities, SIM:241 with nsp = 1) and the draws the integer fast path
// must hand to the reference's walk (see soccer_table.cuh "32-bit draws"): every threshold value x = sum * 2^32 - 0.5 --
// end of a combination, slot inside a 2-way / 4-way combination -- that lies within kDangerWidth of an integer m makes
// the draw r == m uncertain; a slot threshold of 0 (sum below 2^-33: leading combinations of almost zero probability)
// cannot be written as a strict compare and makes r < 4 uncertain.  false: more than 12 such draws (no fast path).
// SOCCER_B200_SLIP_WALK=1 (tests, A/B): every slip env takes the reference's walk
bool soccer_force_slip_walk()
{
    const char* v = getenv("SOCCER_B200_SLIP_WALK");
    return v && v[0] == '1';
}
constexpr double kDangerWidth = 1e-4;       // > 4 x the bound 48 * 2^-53 * 2^32 = 2.3e-5 on |true - constant| thresholds
bool slip_consts_host(const PitchDev& P, SlipE* E, SlipDanger* dg)
{
    double acc = 0.0;
    for (int c = 0; c < 9; ++c) { acc += P.mp[c]; E->e[c] = acc; }
    dg->n = 0;
    bool ok = true, low = false;
    auto add = [&](uint32_t r) {
        for (uint32_t i = 0; i < dg->n; ++i) if (dg->r[i] == r) return;
        if (dg->n < 12) dg->r[dg->n++] = r; else ok = false;
    };
    auto check = [&](double sum) {
        const double x = sum * 4294967296.0 - 0.5, m = nearbyint(x);
        if (fabs(x - m) < kDangerWidth && m >= 0.0 && m <= 4294967295.0) add((uint32_t)m);
    };
    for (int k = 0; k < 9; ++k) {
        check(E->e[k]);
        if (P.mp[k] == 0.0) continue;
        for (int nl = 1; nl <= 2; ++nl) {
            const double pr = P.mp[k] * (nl == 2 ? 0.25 : 0.5);
            double cc = k == 0 ? 0.0 : E->e[k - 1];
            for (int j = 0; j + 1 < (1 << nl); ++j) {
                cc += pr;
                check(cc);
                if (slip_thr(cc) == 0ull) low = true;
            }
        }
    }
    if (low) for (uint32_t r = 0; r < 4; ++r) add(r);
    return ok;
}

constexpr int32_t kMaxRolloutK = 1 << 28;     // per-pass statistics are 32-bit: 4 envs x K flag counts
int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
// scratch layout of soccer_step_host: three input byte streams, flags, obs, reward, obs16, rew8
struct HostScratch { uint8_t *a, *b, *r, *f; int32_t* obs; float* rew; uint16_t* obs16; int8_t* rew8; int64_t bytes; };
HostScratch host_scratch(void* base, int64_t n)
{
    HostScratch s;
    uint8_t* p = reinterpret_cast<uint8_t*>(base);
    const int64_t nb = round_up(n, 256);
    s.a = p; p += nb; s.b = p; p += nb; s.r = p; p += nb; s.f = p; p += nb;
    s.obs = reinterpret_cast<int32_t*>(p); p += 4 * nb;
    s.rew = reinterpret_cast<float*>(p); p += 4 * nb;
    s.obs16 = reinterpret_cast<uint16_t*>(p); p += 2 * nb;
    s.rew8 = reinterpret_cast<int8_t*>(p); p += nb;
    s.bytes = p - reinterpret_cast<uint8_t*>(base);
    return s;
}

int grid_for(int64_t work_items, int blocks_per_sm)
{
    const int64_t need = (work_items + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * blocks_per_sm;
    int64_t g = need < cap ? need : cap;
    return (int)(g < 1 ? 1 : g);
}

int launch_generic(const PitchDev& P, const StepOpts& o, cudaStream_t stream)
{
    if (o.n == 0) return SOCCER_OK;
    static const int nb = resident_blocks(k_step_generic);
    k_step_generic<<<grid_for(o.n, nb), kThreads, 0, stream>>>(P, o);
    return launch_status();
}

int table_grid(int64_t n_groups, int threads = kTableThreads)
{
    const int64_t need = (n_groups + threads - 1) / threads;
    const int64_t cap = sm_count();
    return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

template <typename Kern>
int allow_big_smem(Kern k, int64_t bytes)
{
    return (int)cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

} // namespace soccer

// ====================================================================== C ABI
extern "C" {

int soccer_abi_version(void) { return SOCCER_ABI_VERSION; }

// ---- pinned host memory for the host-buffer path
// Page-locked memory whose physical pages are 2 MB huge pages (2 MB-aligned anonymous mapping + MADV_HUGEPAGE,
// touched, then cudaHostRegister): measured on the B200 boxes, DMA reads (host -> device copies and the kernels'
// zero-copy loads) from a small cudaHostAlloc'd buffer run anywhere between 21 and 55 GB/s depending on the
// physical pages it happened to get, from such a region always at 50-55 GB/s (profiles/r01g_probe_pcie2.log).
static const size_t kHuge = (size_t)2 << 20;
// allocations that fell back to cudaHostAlloc (mmap or cudaHostRegister refused, e.g. a low RLIMIT_MEMLOCK)
static std::mutex g_host_mu;
static std::unordered_set<void*> g_host_plain;
int soccer_host_alloc(size_t bytes, void** ptr)
{
    if (!ptr || bytes == 0) return SOCCER_EINVAL;
    const size_t len = (bytes + kHuge - 1) / kHuge * kHuge;
    uint8_t* raw = (uint8_t*)mmap(nullptr, len + kHuge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (raw != MAP_FAILED) {
        uint8_t* p = (uint8_t*)(((uintptr_t)raw + kHuge - 1) / kHuge * kHuge);
        if (p > raw) munmap(raw, (size_t)(p - raw));                       // trim the slack on both sides:
        if (p + len < raw + len + kHuge) munmap(p + len, (size_t)(raw + len + kHuge - (p + len)));   // [p, p + len) stays
        madvise(p, len, MADV_HUGEPAGE);                                    // best effort; plain pages otherwise
        memset(p, 0, len);                                                 // fault the pages in before pinning
        if (cudaHostRegister(p, len, cudaHostRegisterPortable | cudaHostRegisterMapped) == cudaSuccess) {
            *ptr = p;
            return SOCCER_OK;
        }
        (void)cudaGetLastError();
        munmap(p, len);
    }
    void* q = nullptr;
    const cudaError_t e = cudaHostAlloc(&q, bytes, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) return (int)e;
    { std::lock_guard<std::mutex> lk(g_host_mu); g_host_plain.insert(q); }
    *ptr = q;
    return SOCCER_OK;
}
int soccer_host_free(void* ptr, size_t bytes)
{
    if (!ptr || bytes == 0) return SOCCER_EINVAL;
    bool plain;
    { std::lock_guard<std::mutex> lk(g_host_mu); plain = g_host_plain.erase(ptr) != 0; }
    if (plain) return (int)cudaFreeHost(ptr);
    const size_t len = (bytes + kHuge - 1) / kHuge * kHuge;
    const cudaError_t e = cudaHostUnregister(ptr);
    munmap(ptr, len);
    return (int)e;
}

int soccer_pitch_info_host(const soccer_pitch* pitch, soccer_pitch_info* out) { return fill_info(pitch, out); }

int soccer_pack_state_host(const soccer_pitch* pitch, const int32_t tuple[5], int32_t timestep,
                           int32_t needs_reset, uint32_t* packed)
{
    if (!pitch || !tuple || !packed) return SOCCER_EINVAL;
    if (!pitch_ok(pitch)) return SOCCER_EPITCH;
    const int w = pitch->width, H = pitch->height, W = w + 2;
    int rows[3], ng; goal_rows_of(H, rows, &ng);
    const int xa = tuple[0], ya = tuple[1], xb = tuple[2], yb = tuple[3], p = tuple[4];
    if (xa < 0 || xa >= H || xb < 0 || xb >= H || ya < 0 || ya >= W || yb < 0 || yb >= W || p < 0 || p > 1 ||
        timestep < 0 || timestep > 255)
        return SOCCER_EINVAL;
    auto in_gr = [&](int x) { for (int i = 0; i < ng; ++i) if (rows[i] == x) return true; return false; };
    auto code = [&](int x, int y, bool has, uint32_t* c) {
        if (y == 0 || y == W - 1) {
            if (!in_gr(x) || !has) return false;            // SIM:74-83 unreachable
            *c = kGoalBit | (y != 0 ? kRightBit : 0u) | (uint32_t)x;
        } else *c = field_code(w, x, y);
        return true;
    };
    uint32_t a, b;
    if (!code(xa, ya, p == 0, &a) || !code(xb, yb, p == 1, &b)) return SOCCER_EINVAL;
    if (xa == xb && ya == yb) return SOCCER_EINVAL;         // SIM:86-88
    *packed = a | (b << 8) | ((uint32_t)timestep << 16) | ((uint32_t)p << 24) | (needs_reset ? kNeedsReset : 0u);
    return SOCCER_OK;
}

int soccer_unpack_state_host(const soccer_pitch* pitch, uint32_t packed, int32_t tuple[5],
                             int32_t* timestep, int32_t* needs_reset)
{
    if (!pitch || !tuple) return SOCCER_EINVAL;
    if (!pitch_ok(pitch)) return SOCCER_EPITCH;
    const int w = pitch->width;
    auto dec = [&](uint32_t c, int32_t* x, int32_t* y) {
        if (c & kGoalBit) { *x = (int32_t)(c & 0x3Fu); *y = (c & kRightBit) ? w + 1 : 0; }
        else { *x = (int32_t)c / w; *y = (int32_t)c % w + 1; }
    };
    dec(packed & 0xFFu, &tuple[0], &tuple[1]);
    dec((packed >> 8) & 0xFFu, &tuple[2], &tuple[3]);
    tuple[4] = (int32_t)((packed >> 24) & 1u);
    if (timestep) *timestep = (int32_t)((packed >> 16) & 0xFFu);
    if (needs_reset) *needs_reset = (packed & kNeedsReset) ? 1 : 0;
    return SOCCER_OK;
}

int soccer_state_to_obs_host(const soccer_pitch* pitch, uint32_t packed, int32_t* obs)
{
    if (!pitch || !obs) return SOCCER_EINVAL;
    if (!pitch_ok(pitch)) return SOCCER_EPITCH;
    const uint32_t a = packed & 0xFFu, b = (packed >> 8) & 0xFFu, p = (packed >> 24) & 1u;
    const int F = pitch->width * pitch->height;
    if ((a | b) & kGoalBit) { *obs = 0; return SOCCER_OK; }   // SIM:493
    if ((int)a >= F || (int)b >= F || a == b) return SOCCER_EINVAL;
    *obs = host_obs_index(F, a, b, p);
    return SOCCER_OK;
}

int soccer_obs_to_state_host(const soccer_pitch* pitch, int32_t obs, uint32_t* packed)
{
    if (!pitch || !packed) return SOCCER_EINVAL;
    if (!pitch_ok(pitch)) return SOCCER_EPITCH;
    const int F = pitch->width * pitch->height;
    if (obs < 1 || obs >= 1 + 2 * F * (F - 1)) return SOCCER_EINVAL;
    const uint32_t idx = (uint32_t)(obs - 1), p = idx & 1u, q = idx >> 1;
    const uint32_t a = q / (uint32_t)(F - 1), rb = q % (uint32_t)(F - 1), b = rb + (rb >= a ? 1u : 0u);
    *packed = a | (b << 8) | (p << 24);
    return SOCCER_OK;
}

int soccer_reset(const soccer_pitch* pitch, uint32_t* state, int32_t* obs_out, const uint8_t* rng8,
                 const uint8_t* mask, int64_t n, soccer_stream_t stream)
{
    if (!state || !rng8 || n < 0) return SOCCER_EINVAL;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    if (n == 0) return SOCCER_OK;
    k_reset<<<grid_for(n, 8), kThreads, 0, (cudaStream_t)stream>>>(P, state, obs_out, rng8, mask, 0, 0, 0, 0, n);
    return launch_status();
}

int soccer_reset_philox(const soccer_pitch* pitch, uint32_t* state, int32_t* obs_out, const uint8_t* mask,
                        uint64_t seed, uint64_t step, uint64_t env_id_base, int64_t n, soccer_stream_t stream)
{
    if (!state || n < 0) return SOCCER_EINVAL;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    if (n == 0) return SOCCER_OK;
    k_reset<<<grid_for(n, 8), kThreads, 0, (cudaStream_t)stream>>>(P, state, obs_out, nullptr, mask, 1, seed, step,
                                                                    env_id_base, n);
    return launch_status();
}

int soccer_set_state(const soccer_pitch* pitch, uint32_t* state, const int32_t* obs_in,
                     const int32_t* timestep_in, int64_t n, soccer_stream_t stream)
{
    if (!state || !obs_in || n < 0) return SOCCER_EINVAL;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    if (n == 0) return SOCCER_OK;
    const int nS = 1 + 2 * P.F * P.Fm1;
    k_set_state<<<grid_for(n, 8), kThreads, 0, (cudaStream_t)stream>>>(P, nS, state, obs_in, timestep_in, n);
    return launch_status();
}

int soccer_get_obs(const soccer_pitch* pitch, const uint32_t* state, int32_t* obs_out, int64_t n,
                   soccer_stream_t stream)
{
    if (!state || !obs_out || n < 0) return SOCCER_EINVAL;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    if (n == 0) return SOCCER_OK;
    k_get_obs<<<grid_for(n, 8), kThreads, 0, (cudaStream_t)stream>>>(P, state, obs_out, n);
    return launch_status();
}

} // extern "C"
namespace { int step_table_ex(const soccer_pitch* pitch, const soccer_step_args* a, soccer_stream_t stream); }
extern "C" {

int soccer_step_ex(const soccer_pitch* pitch, const soccer_step_args* a, soccer_stream_t stream)
{
    if (!a || !a->state || a->n < 0) return SOCCER_EINVAL;
    if (a->policy_a && a->policy_b) return SOCCER_EPOLICY;                 // SIM:38
    if (a->table) return step_table_ex(pitch, a, stream);                  // INDEX-layout states, shared-memory table kernels
    if (a->slip_index) return SOCCER_EINVAL;
    if ((!a->policy_a && !a->act_a) || (!a->policy_b && !a->act_b)) return SOCCER_EINVAL;
    if (!a->use_philox && !a->rng8) return SOCCER_EINVAL;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    if (P.slip && !a->use_philox && !a->rng32 && !a->rngf64) return SOCCER_ESLIP;
    if (a->n == 0) return SOCCER_OK;
    cudaStream_t st = (cudaStream_t)stream;

    if (a->narrow != 0 && a->narrow != 1) return SOCCER_EINVAL;
    const bool narrow = a->narrow == 1;
    const bool pol = a->policy_a || a->policy_b;     // single-agent modes: the folded player's action stream may be NULL
    const bool fast_ok = !P.slip && a->auto_reset && !a->detail && !(pol && narrow) &&
                         a->obs && a->reward && a->flags && a->n >= 4 &&
                         aligned(a->state, 16) && aligned(a->obs, narrow ? 8 : 16) && aligned(a->reward, narrow ? 4 : 16) &&
                         (!a->reset_obs || aligned(a->reset_obs, 16)) &&
                         (a->policy_a ? (!a->act_a || aligned(a->act_a, 4)) : (a->act_a && aligned(a->act_a, 4))) &&
                         (a->policy_b ? (!a->act_b || aligned(a->act_b, 4)) : (a->act_b && aligned(a->act_b, 4))) &&
                         (a->use_philox || aligned(a->rng8, 4)) && aligned(a->flags, 4) &&
                         !(a->use_philox && (a->env_id_base & 3u)) &&      // Philox contract v2: a thread's 4 envs = one aligned group
                         !(narrow && (a->use_philox || a->reset_obs));     // narrow fast path: the plain step only
    int64_t done_n = 0;
    // slip_prob > 0 with 32-bit draws (rng32 stream or Philox), the plain options: integer-threshold kernel
    RulesSlipArgs rsa = {};
    const bool slip_fast_ok = P.slip && a->auto_reset && !a->detail && !narrow && !a->stats &&
                              a->obs && a->reward && a->flags && a->n >= 4 && !a->rngf64 && (a->use_philox || a->rng32) &&
                              aligned(a->state, 16) && aligned(a->obs, 16) && aligned(a->reward, 16) &&
                              (!a->reset_obs || aligned(a->reset_obs, 16)) &&
                              (a->policy_a ? (!a->act_a || aligned(a->act_a, 4)) : (a->act_a && aligned(a->act_a, 4))) &&
                              (a->policy_b ? (!a->act_b || aligned(a->act_b, 4)) : (a->act_b && aligned(a->act_b, 4))) &&
                              aligned(a->flags, 4) &&
                              (a->use_philox ? (a->env_id_base & 3u) == 0 : (aligned(a->rng8, 4) && aligned(a->rng32, 16))) &&
                              slip_consts_host(P, &rsa.E, &rsa.dg) && !soccer_force_slip_walk();
    if (slip_fast_ok) {
        rsa.use_int = 1;
        const int64_t n_groups = a->n / 4;
        const PhiloxKey key = { a->seed, a->step, a->env_id_base };
#define SOCCER_LAUNCH_FAST_SLIP(RO, PH)                                                                           \
        do {                                                                                                      \
            static const int nb = resident_blocks(k_step_fast_slip<RO, PH>);                                      \
            k_step_fast_slip<RO, PH><<<grid_for(n_groups, nb), kThreads, 0, st>>>(P, rsa, a->state, a->act_a, a->act_b, \
                a->rng8, a->rng32, a->obs, a->reward, a->flags, RO ? a->reset_obs : (int32_t*)nullptr, n_groups, key); \
        } while (0)
#define SOCCER_LAUNCH_FAST_SLIP_POL(RO, PH)                                                                       \
        do {                                                                                                      \
            static const int nb = resident_blocks(k_step_fast_slip<RO, PH, true>);                                \
            k_step_fast_slip<RO, PH, true><<<grid_for(n_groups, nb), kThreads, 0, st>>>(P, rsa, a->state, a->act_a, a->act_b, \
                a->rng8, a->rng32, a->obs, a->reward, a->flags, RO ? a->reset_obs : (int32_t*)nullptr, n_groups, key, \
                a->policy_a, a->policy_b);                                                                        \
        } while (0)
        if (a->policy_a || a->policy_b) {
            if (a->use_philox) { if (a->reset_obs) SOCCER_LAUNCH_FAST_SLIP_POL(true, true); else SOCCER_LAUNCH_FAST_SLIP_POL(false, true); }
            else { if (a->reset_obs) SOCCER_LAUNCH_FAST_SLIP_POL(true, false); else SOCCER_LAUNCH_FAST_SLIP_POL(false, false); }
        }
        else if (a->use_philox) { if (a->reset_obs) SOCCER_LAUNCH_FAST_SLIP(true, true); else SOCCER_LAUNCH_FAST_SLIP(false, true); }
        else { if (a->reset_obs) SOCCER_LAUNCH_FAST_SLIP(true, false); else SOCCER_LAUNCH_FAST_SLIP(false, false); }
#undef SOCCER_LAUNCH_FAST_SLIP
#undef SOCCER_LAUNCH_FAST_SLIP_POL
        const int e = launch_status();
        if (e) return e;
        done_n = n_groups * 4;
        if (done_n == a->n) return SOCCER_OK;
    }
    if (fast_ok) {
        // byte-parallel rules kernel; with on-device Philox draws (soccer_step_philox) 19 B / env-step
        const int64_t n_groups = a->n / 4;
        const PhiloxKey key = { a->seed, a->step, a->env_id_base };
#define SOCCER_LAUNCH_FAST(RO, PH, NR)                                                                            \
        do {                                                                                                      \
            static const int nb = resident_blocks(k_step_fast<RO, PH, NR>);                                       \
            const int e1 = launch_pdl(k_step_fast<RO, PH, NR>, grid_for(n_groups, nb), kThreads, 0, st, P, a->state, \
                                      a->act_a, a->act_b, PH ? (const uint8_t*)nullptr : a->rng8, a->obs, a->reward, \
                                      a->flags, RO ? a->reset_obs : (int32_t*)nullptr, n_groups, key, a->stats,   \
                                      (const int8_t*)nullptr, (const int8_t*)nullptr);                            \
            if (e1) return e1;                                                                                    \
        } while (0)
#define SOCCER_LAUNCH_FAST_POL(RO, PH)                                                                            \
        do {                                                                                                      \
            static const int nb = resident_blocks(k_step_fast<RO, PH, false, true>);                              \
            const int e1 = launch_pdl(k_step_fast<RO, PH, false, true>, grid_for(n_groups, nb), kThreads, 0, st, P, a->state, \
                                      a->act_a, a->act_b, PH ? (const uint8_t*)nullptr : a->rng8, a->obs, a->reward, \
                                      a->flags, RO ? a->reset_obs : (int32_t*)nullptr, n_groups, key, a->stats,   \
                                      a->policy_a, a->policy_b);                                                  \
            if (e1) return e1;                                                                                    \
        } while (0)
        if (pol) {
            if (a->use_philox) { if (a->reset_obs) SOCCER_LAUNCH_FAST_POL(true, true); else SOCCER_LAUNCH_FAST_POL(false, true); }
            else { if (a->reset_obs) SOCCER_LAUNCH_FAST_POL(true, false); else SOCCER_LAUNCH_FAST_POL(false, false); }
        }
        else if (narrow) SOCCER_LAUNCH_FAST(false, false, true);
        else if (a->use_philox) { if (a->reset_obs) SOCCER_LAUNCH_FAST(true, true, false); else SOCCER_LAUNCH_FAST(false, true, false); }
        else { if (a->reset_obs) SOCCER_LAUNCH_FAST(true, false, false); else SOCCER_LAUNCH_FAST(false, false, false); }
#undef SOCCER_LAUNCH_FAST
#undef SOCCER_LAUNCH_FAST_POL
        const int e = launch_status();
        if (e) return e;
        done_n = n_groups * 4;
        if (done_n == a->n) return SOCCER_OK;
    }
    StepOpts o;
    const int64_t k = done_n;
    o.state = a->state + k;
    o.act_a = a->act_a ? a->act_a + k : nullptr;
    o.act_b = a->act_b ? a->act_b + k : nullptr;
    o.rng8 = a->rng8 ? a->rng8 + k : nullptr;
    o.rng32 = a->rng32 ? a->rng32 + k : nullptr;
    o.rngf64 = a->rngf64 ? a->rngf64 + k : nullptr;
    o.policy_a = a->policy_a; o.policy_b = a->policy_b;
    // narrow: obs / reward hold 2- / 1-byte elements
    o.obs = !a->obs ? nullptr : (narrow ? reinterpret_cast<int32_t*>(reinterpret_cast<uint16_t*>(a->obs) + k) : a->obs + k);
    o.reward = !a->reward ? nullptr : (narrow ? reinterpret_cast<float*>(reinterpret_cast<int8_t*>(a->reward) + k) : a->reward + k);
    o.flags = a->flags ? a->flags + k : nullptr;
    o.reset_obs = a->reset_obs ? a->reset_obs + k : nullptr;
    o.n = a->n - k; o.auto_reset = a->auto_reset; o.use_philox = a->use_philox; o.detail = a->detail;
    o.seed = a->seed; o.step = a->step; o.env_id_base = a->env_id_base + (uint64_t)k;
    o.narrow = a->narrow;
    o.stats = a->stats;
    return launch_generic(P, o, st);
}

int soccer_step(const soccer_pitch* pitch, uint32_t* state, const uint8_t* act_a, const uint8_t* act_b,
                const uint8_t* rng8, int32_t* obs, float* reward, uint8_t* flags, int32_t* reset_obs,
                int64_t n, soccer_stream_t stream)
{
    if (!obs || !reward || !flags) return SOCCER_EINVAL;
    if (pitch && pitch->slip_prob != 0.0) return SOCCER_ESLIP;
    soccer_step_args a = {};
    a.state = state; a.act_a = act_a; a.act_b = act_b; a.rng8 = rng8; a.obs = obs; a.reward = reward;
    a.flags = flags; a.reset_obs = reset_obs; a.n = n; a.auto_reset = 1;
    return soccer_step_ex(pitch, &a, stream);
}

int soccer_step_philox(const soccer_pitch* pitch, uint32_t* state, const uint8_t* act_a, const uint8_t* act_b,
                       uint64_t seed, uint64_t step, uint64_t env_id_base, int32_t* obs, float* reward,
                       uint8_t* flags, int32_t* reset_obs, int64_t n, soccer_stream_t stream)
{
    if (!obs || !reward || !flags) return SOCCER_EINVAL;
    soccer_step_args a = {};
    a.state = state; a.act_a = act_a; a.act_b = act_b; a.obs = obs; a.reward = reward; a.flags = flags;
    a.reset_obs = reset_obs; a.n = n; a.auto_reset = 1; a.use_philox = 1; a.seed = seed; a.step = step;
    a.env_id_base = env_id_base;
    return soccer_step_ex(pitch, &a, stream);
}

int soccer_rollout(const soccer_pitch* pitch, uint32_t* state, const int8_t* policy_a, const int8_t* policy_b,
                   uint64_t seed, uint64_t step0, int32_t K, uint64_t env_id_base, int32_t flip_reward, int32_t* obs,
                   float* reward, uint8_t* flags, unsigned long long* stats, int64_t n, soccer_stream_t stream)
{
    if (!state || n < 0 || K < 0 || K > kMaxRolloutK) return SOCCER_EINVAL;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    if (n == 0 || K == 0) return SOCCER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // Philox contract v2: the 4 envs of a thread must be one aligned group of the GLOBAL env ids
    const bool vec = (n % 4 == 0) && (env_id_base % 4 == 0) && aligned(state, 16) && (!obs || aligned(obs, 16)) &&
                     (!reward || aligned(reward, 16)) && (!flags || aligned(flags, 4));
    const RolloutArgs ra = { state, seed, step0, K, env_id_base, obs, reward, flags, stats, n, philox_round_keys(seed), flip_reward ? 1 : 0 };
    const bool streams = obs && reward && flags;
#define SOCCER_LAUNCH_ROLLOUT(VEC, STR, SLIP, ITEMS)                                                    \
    do {                                                                                                 \
        static const int nb = resident_blocks(k_rollout<VEC, STR, SLIP>);                                \
        k_rollout<VEC, STR, SLIP><<<grid_for(ITEMS, nb), kThreads, 0, st>>>(P, policy_a, policy_b, ra);   \
    } while (0)
#define SOCCER_PICK_ROLLOUT(SLIP)                                                                        \
    do {                                                                                                 \
        if (vec && streams) SOCCER_LAUNCH_ROLLOUT(4, true, SLIP, n / 4);                                 \
        else if (vec) SOCCER_LAUNCH_ROLLOUT(4, false, SLIP, n / 4);                                      \
        else if (streams) SOCCER_LAUNCH_ROLLOUT(1, true, SLIP, n);                                       \
        else SOCCER_LAUNCH_ROLLOUT(1, false, SLIP, n);                                                   \
    } while (0)
    RulesSlipArgs rsa = {};
    if (P.slip) {
        rsa.use_int = slip_consts_host(P, &rsa.E, &rsa.dg) && !soccer_force_slip_walk() ? 1 : 0;
        // uniform policy, 4 envs per thread: combination and slot of the 32-bit draw by integer thresholds, resolution
        // byte-parallel (step4_slip_int), table policies gathered per env.  When the fast path is off the scalar walk, one env per
        // thread: it needs ~170 registers with 4 envs per thread and measured 2.2x faster that way (44 vs 20 G env-steps/s)
        if (rsa.use_int && vec) {
#define SOCCER_LAUNCH_ROLLOUT_SLIPI(STR, POL)                                                            \
            do {                                                                                         \
                static const int nb = resident_blocks(k_rollout_slipi<STR, POL>);                        \
                k_rollout_slipi<STR, POL><<<grid_for(n / 4, nb), kThreads, 0, st>>>(P, ra, rsa, policy_a, policy_b); \
            } while (0)
            const bool polr = policy_a || policy_b;
            if (streams) { if (polr) SOCCER_LAUNCH_ROLLOUT_SLIPI(true, true); else SOCCER_LAUNCH_ROLLOUT_SLIPI(true, false); }
            else { if (polr) SOCCER_LAUNCH_ROLLOUT_SLIPI(false, true); else SOCCER_LAUNCH_ROLLOUT_SLIPI(false, false); }
#undef SOCCER_LAUNCH_ROLLOUT_SLIPI
        }
        else if (streams) SOCCER_LAUNCH_ROLLOUT(1, true, true, n);
        else SOCCER_LAUNCH_ROLLOUT(1, false, true, n);
    } else {
        SOCCER_PICK_ROLLOUT(false);
    }
#undef SOCCER_PICK_ROLLOUT
#undef SOCCER_LAUNCH_ROLLOUT
    return launch_status();
}

int soccer_sweep(const soccer_pitch* pitch, int32_t n_combos, uint8_t* n_out, uint32_t* next_state,
                 int32_t* next_obs, int8_t* reward, uint8_t* done, soccer_stream_t stream)
{
    if (n_combos != 1 && n_combos != 9) return SOCCER_EINVAL;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    const int nS = 1 + 2 * P.F * P.Fm1;
    const int64_t total = (int64_t)(nS - 1) * 25 * n_combos;
    k_sweep<<<grid_for(total, 4), kThreads, 0, (cudaStream_t)stream>>>(P, n_combos, nS, n_out, next_state,
                                                                        next_obs, reward, done);
    return launch_status();
}

int soccer_step_table_bytes_host(const soccer_pitch* pitch, int64_t* bytes)
{
    if (!bytes) return SOCCER_EINVAL;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    return table_bytes_of(P, bytes);
}

int soccer_build_step_table(const soccer_pitch* pitch, uint16_t* table, soccer_stream_t stream)
{
    if (!table) return SOCCER_EINVAL;
    PitchDev P; int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    int64_t bytes; rc = table_bytes_of(P, &bytes); if (rc) return rc;
    k_build_step_table<<<grid_for((int64_t)P.nS * 100, 4), kThreads, 0, (cudaStream_t)stream>>>(P, P.nS, table);
    return launch_status();
}

} // extern "C"
namespace {
using namespace soccer;

// slip_prob > 0 through the table (k_step_table_slip_q / k_step_table_slip / scalar tail); draws from rng32 / rngf64 or,
// when both are NULL, from Philox
int step_table_slip_impl(const PitchDev& P, int64_t bytes, const soccer_step_args* a, cudaStream_t st)
{
    const bool philox = a->use_philox != 0;
    const int64_t n = a->n;
    const uint8_t* act_a = a->act_a; const uint8_t* act_b = a->act_b;
    const void* draw = a->rngf64 ? (const void*)a->rngf64 : (const void*)a->rng32;
    const SlipExtra ex = { a->policy_a, a->policy_b, { a->seed, a->step, a->env_id_base } };
    const bool vec = n >= 4 && aligned(a->state, 16) && aligned(a->obs, 16) && aligned(a->reward, 16) &&
                     (!a->reset_obs || aligned(a->reset_obs, 16)) && (!act_a || aligned(act_a, 4)) &&
                     (!act_b || aligned(act_b, 4)) && aligned(a->flags, 4) &&
                     (philox ? (a->env_id_base & 3u) == 0 : (aligned(a->rng8, 4) && aligned(draw, 16)));
    const int64_t pol_bytes = (a->policy_a || a->policy_b) ? 2 * (((int64_t)P.nS + 15) & ~15) : 0;
    int64_t done_n = 0;
    if (vec) {
        const int64_t n_groups = n / 4;
        // 32-bit draws (rng32 / Philox) take the integer-threshold fast path (k_step_table_slip_i; needs no index: 5x4 with
        // a 12-bit, 6x4 with a 10-bit bucket table); fp64 draws, with a slip index and room for it and the deferral queue
        // next to the table (5x4), the constant-prefix fast path + queued walk (k_step_table_slip_q, index plane 0); else
        // the in-place walk.
        const int64_t fc_bytes = ((int64_t)P.nS * 25 + 15) / 16 * 16;
        const int64_t smem_q = bytes + 16 + fc_bytes + pol_bytes + (int64_t)kSlipQueueBytes;
        const int64_t smem_w = bytes + 16 + pol_bytes;
        if (smem_w > 227 * 1024 - 2048) return SOCCER_ETABLE;
        SlipE E; SlipDanger dg;
        const bool dg_ok = slip_consts_host(P, &E, &dg);
        const bool f64 = !philox && a->rngf64;
        int lut_bits = 12;
        if (smem_w + slip_int_lut_bytes(lut_bits) > 227 * 1024 - 1536) lut_bits = 10;
        const int64_t smem_i = smem_w + slip_int_lut_bytes(lut_bits);
        const bool queued = a->slip_index && f64 && smem_q <= 227 * 1024 - 1024;
        const bool integer = !f64 && dg_ok && smem_i <= 227 * 1024 - 1536 && !soccer_force_slip_walk();
        // the input ring (cp.async, 2 / 3 stages) where the table leaves room for it: with the 12-bit bucket table if that
        // fits too, else with the 10-bit one; the 6x4 table (221 KB) leaves none: register prefetch
        const int64_t smem_cap = 227 * 1024 - 1536;
        auto ring_bytes = [&](bool ph) { return (int64_t)slip_ring_stages(ph) * slip_i_threads<true>() * slip_ring_stage_bytes(ph); };
        static_assert(slip_i_threads<true>() == slip_i_threads<false>(), "one CTA size for both draw sources");
        int ring_lut_bits = 12;
        if (smem_w + ((slip_int_lut_bytes(12) + 15) & ~15) + ring_bytes(philox) > smem_cap) ring_lut_bits = 10;
        const int64_t smem_r = smem_w + ((slip_int_lut_bytes(ring_lut_bits) + 15) & ~15) + ring_bytes(philox);
        const bool ring = SOCCER_SLIP_I_RING && smem_r <= smem_cap;
#define SOCCER_LAUNCH_SLIP_I3(RO, PH, POL, RING, SMEM, BITS)                                              \
        do {                                                                                              \
            const int e0 = allow_big_smem(k_step_table_slip_i<RO, PH, POL, RING>, SMEM);                  \
            if (e0) return e0;                                                                            \
            k_step_table_slip_i<RO, PH, POL, RING><<<table_grid(n_groups, slip_i_threads<PH>()), slip_i_threads<PH>(), (size_t)(SMEM), st>>>(  \
                P, a->table, (uint32_t)bytes, E, dg, slip_bits(BITS), a->state, act_a, act_b,             \
                a->rng8, a->rng32, a->obs, a->reward, a->flags, a->reset_obs, n_groups, ex);              \
        } while (0)
#define SOCCER_LAUNCH_SLIP_I2(RO, PH, POL)                                                                \
        do {                                                                                              \
            if (ring) SOCCER_LAUNCH_SLIP_I3(RO, PH, POL, true, smem_r, ring_lut_bits);                    \
            else SOCCER_LAUNCH_SLIP_I3(RO, PH, POL, false, smem_i, lut_bits);                             \
        } while (0)
#define SOCCER_LAUNCH_SLIP_I(RO, PH)                                                                      \
        do { if (pol_bytes) SOCCER_LAUNCH_SLIP_I2(RO, PH, true); else SOCCER_LAUNCH_SLIP_I2(RO, PH, false); } while (0)
#define SOCCER_LAUNCH_SLIP_T(RO, DRAW)                                                                    \
        do {                                                                                              \
            if (queued) {                                                                                 \
                const int e0 = allow_big_smem(k_step_table_slip_q<RO, DRAW>, smem_q);                     \
                if (e0) return e0;                                                                        \
                k_step_table_slip_q<RO, DRAW><<<table_grid(n_groups, kTableThreads), kTableThreads, (size_t)smem_q, st>>>(  \
                    P, a->table, (uint32_t)bytes, a->slip_index, (uint32_t)fc_bytes, E, a->state, act_a, act_b, a->rng8, \
                    draw, a->obs, a->reward, a->flags, a->reset_obs, n_groups, ex);                       \
            } else {                                                                                      \
                const int e0 = allow_big_smem(k_step_table_slip<RO, DRAW>, smem_w);                       \
                if (e0) return e0;                                                                        \
                k_step_table_slip<RO, DRAW><<<table_grid(n_groups, kSlipThreads), kSlipThreads, (size_t)smem_w, st>>>(  \
                    P, a->table, (uint32_t)bytes, a->state, act_a, act_b, a->rng8, draw, a->obs, a->reward, a->flags, \
                    a->reset_obs, n_groups, ex);                                                          \
            }                                                                                             \
        } while (0)
#define SOCCER_PICK_SLIP_T(RO)                                                                            \
        do {                                                                                              \
            if (integer) { if (philox) SOCCER_LAUNCH_SLIP_I(RO, true); else SOCCER_LAUNCH_SLIP_I(RO, false); } \
            else if (philox) SOCCER_LAUNCH_SLIP_T(RO, kDrawPhilox);                                       \
            else if (a->rngf64) SOCCER_LAUNCH_SLIP_T(RO, kDrawF64);                                       \
            else SOCCER_LAUNCH_SLIP_T(RO, kDrawU32);                                                      \
        } while (0)
        if (a->reset_obs) SOCCER_PICK_SLIP_T(true); else SOCCER_PICK_SLIP_T(false);
#undef SOCCER_PICK_SLIP_T
#undef SOCCER_LAUNCH_SLIP_T
#undef SOCCER_LAUNCH_SLIP_I
#undef SOCCER_LAUNCH_SLIP_I2
#undef SOCCER_LAUNCH_SLIP_I3
        const int e = launch_status();
        if (e) return e;
        done_n = n_groups * 4;
        if (done_n == n) return SOCCER_OK;
    }
    const int64_t k = done_n, m = n - k;
    const SlipExtra exk = { a->policy_a, a->policy_b, { a->seed, a->step, a->env_id_base + (uint64_t)k } };
    k_step_table_slip_scalar<<<grid_for(m, 8), kThreads, 0, st>>>(
        P, a->table, a->state + k, act_a ? act_a + k : nullptr, act_b ? act_b + k : nullptr, philox ? nullptr : a->rng8 + k,
        (philox || a->rngf64) ? nullptr : a->rng32 + k, (!philox && a->rngf64) ? a->rngf64 + k : nullptr, a->obs + k,
        a->reward + k, a->flags + k, a->reset_obs ? a->reset_obs + k : nullptr, m, philox ? 1 : 0, exk);
    return launch_status();
}

// Every option the shared-memory-table step offers (soccer_step_ex with args->table != NULL): injected or Philox
// draws, folded table policies (single-agent modes), slip_prob > 0, narrow streams, fused statistics.
int step_table_ex(const soccer_pitch* pitch, const soccer_step_args* a, soccer_stream_t stream)
{
    const bool philox = a->use_philox != 0;
    const bool narrow = a->narrow == 1;
    if (a->narrow != 0 && a->narrow != 1) return SOCCER_EINVAL;
    if (!a->auto_reset || a->detail) return SOCCER_EINVAL;     // INDEX-layout states cannot hold needs_reset / detail flags
    if (!a->obs || !a->reward || !a->flags) return SOCCER_EINVAL;
    if ((!a->policy_a && !a->act_a) || (!a->policy_b && !a->act_b)) return SOCCER_EINVAL;
    if (!philox && !a->rng8) return SOCCER_EINVAL;
    const bool pol = a->policy_a || a->policy_b;
    if (narrow && (philox || a->reset_obs || pol)) return SOCCER_EINVAL;
    PitchDev P; int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    int64_t bytes; rc = table_bytes_of(P, &bytes); if (rc) return rc;
    if (!aligned(a->table, 16) || (a->slip_index && !aligned(a->slip_index, 16))) return SOCCER_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (P.slip) {
        if (!philox && !a->rng32 && !a->rngf64) return SOCCER_ESLIP;
        if (narrow || a->stats) return SOCCER_EINVAL;          // not offered by the slip kernels
        if (a->n == 0) return SOCCER_OK;
        return step_table_slip_impl(P, bytes, a, st);
    }
    if (a->n == 0) return SOCCER_OK;
    const int64_t n = a->n;
    const PhiloxKey key = { a->seed, a->step, a->env_id_base };
    const K1Extra ex = { a->policy_a, a->policy_b, a->stats };
    const int64_t smem = bytes + 16 + (pol ? 2 * (((int64_t)P.nS + 15) & ~15) : 0);
    if (smem > 227 * 1024 - 1024) return SOCCER_ETABLE;
    const bool vec = n >= 4 && aligned(a->state, 16) && aligned(a->obs, narrow ? 8 : 16) && aligned(a->reward, narrow ? 4 : 16) &&
                     (!a->reset_obs || aligned(a->reset_obs, 16)) && (!a->act_a || aligned(a->act_a, 4)) &&
                     (!a->act_b || aligned(a->act_b, 4)) && (philox ? (a->env_id_base & 3u) == 0 : aligned(a->rng8, 4)) &&
                     aligned(a->flags, 4);
    int64_t done_n = 0;
    if (vec) {
        const int64_t n_groups = n / 4;
#define SOCCER_LAUNCH_STEP_T4(RO, PH, NR, PL, ST)                                                                 \
        do {                                                                                                      \
            const int e0 = allow_big_smem(k_step_table<RO, PH, NR, PL, ST>, smem);                                \
            if (e0) return e0;                                                                                    \
            constexpr int thr = k1_threads<PH, PL, ST>();                                                         \
            const int e1 = launch_pdl(k_step_table<RO, PH, NR, PL, ST>, table_grid(n_groups, thr), thr, (size_t)smem, st, \
                                      P, a->table, (uint32_t)bytes, a->state, a->act_a, a->act_b, a->rng8, a->obs, a->reward, \
                                      a->flags, RO ? a->reset_obs : (int32_t*)nullptr, n_groups, key, ex);        \
            if (e1) return e1;                                                                                    \
        } while (0)
#define SOCCER_LAUNCH_STEP_T(RO, PH, NR, PL)                                                                      \
        do { if (a->stats) SOCCER_LAUNCH_STEP_T4(RO, PH, NR, PL, true); else SOCCER_LAUNCH_STEP_T4(RO, PH, NR, PL, false); } while (0)
#define SOCCER_PICK_STEP_T(RO, PH)                                                                                \
        do { if (pol) SOCCER_LAUNCH_STEP_T(RO, PH, false, true); else SOCCER_LAUNCH_STEP_T(RO, PH, false, false); } while (0)
        if (narrow) SOCCER_LAUNCH_STEP_T(false, false, true, false);
        else if (a->reset_obs) { if (philox) SOCCER_PICK_STEP_T(true, true); else SOCCER_PICK_STEP_T(true, false); }
        else { if (philox) SOCCER_PICK_STEP_T(false, true); else SOCCER_PICK_STEP_T(false, false); }
#undef SOCCER_PICK_STEP_T
#undef SOCCER_LAUNCH_STEP_T
#undef SOCCER_LAUNCH_STEP_T4
        const int e = launch_status();
        if (e) return e;
        done_n = n_groups * 4;
        if (done_n == n) return SOCCER_OK;
    }
    const int64_t k = done_n, m = n - k;
    const PhiloxKey tail_key = { key.seed, key.step, key.env_id_base + (uint64_t)k };
    int32_t* obs_k = narrow ? reinterpret_cast<int32_t*>(reinterpret_cast<uint16_t*>(a->obs) + k) : a->obs + k;
    float* rew_k = narrow ? reinterpret_cast<float*>(reinterpret_cast<int8_t*>(a->reward) + k) : a->reward + k;
    k_step_table_scalar<<<grid_for(m, 8), kThreads, 0, st>>>(P, a->table, a->state + k, a->act_a ? a->act_a + k : nullptr,
                                                             a->act_b ? a->act_b + k : nullptr,
                                                             philox ? nullptr : a->rng8 + k, obs_k, rew_k, a->flags + k,
                                                             a->reset_obs ? a->reset_obs + k : nullptr, m, philox ? 1 : 0, tail_key,
                                                             narrow ? 1 : 0, ex);
    return launch_status();
}

int packed_impl(const soccer_pitch* pitch, const uint16_t* table, uint32_t* state, const uint8_t* joint,
                const uint8_t* rng8, const PhiloxKey key, uint16_t* result, int64_t n, soccer_stream_t stream)
{
    const bool philox = rng8 == nullptr;
    if (!table || !state || !joint || !result || n < 0) return SOCCER_EINVAL;
    if (pitch && pitch->slip_prob != 0.0) return SOCCER_ESLIP;
    PitchDev P; int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    int64_t bytes; rc = table_bytes_of(P, &bytes); if (rc) return rc;
    if (!aligned(table, 16)) return SOCCER_EINVAL;
    if (n == 0) return SOCCER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t done_n = 0;
    if (n >= 4 && aligned(state, 16) && aligned(joint, 4) && aligned(result, 8) &&
        (philox ? (key.env_id_base & 3u) == 0 : aligned(rng8, 4))) {
        const int64_t n_groups = n / 4;
        if (philox) {
            const int e0 = allow_big_smem(k_step_table_packed<true>, bytes + 16);
            if (e0) return e0;
            const int e1 = launch_pdl(k_step_table_packed<true>, table_grid(n_groups), kTableThreads, (size_t)bytes + 16, st,
                                      P, table, (uint32_t)bytes, state, joint, rng8, result, n_groups, key);
            if (e1) return e1;
        } else {
            const int e0 = allow_big_smem(k_step_table_packed<false>, bytes + 16);
            if (e0) return e0;
            const int e1 = launch_pdl(k_step_table_packed<false>, table_grid(n_groups), kTableThreads, (size_t)bytes + 16, st,
                                      P, table, (uint32_t)bytes, state, joint, rng8, result, n_groups, key);
            if (e1) return e1;
        }
        const int e = launch_status();
        if (e) return e;
        done_n = n_groups * 4;
        if (done_n == n) return SOCCER_OK;
    }
    const int64_t k = done_n, m = n - k;
    const PhiloxKey tail_key = { key.seed, key.step, key.env_id_base + (uint64_t)k };
    k_step_table_packed_scalar<<<grid_for(m, 8), kThreads, 0, st>>>(P, table, state + k, joint + k, philox ? nullptr : rng8 + k,
                                                                    result + k, m, philox ? 1 : 0, tail_key);
    return launch_status();
}
} // namespace
extern "C" {

int soccer_step_table(const soccer_pitch* pitch, const uint16_t* table, uint32_t* state, const uint8_t* act_a,
                      const uint8_t* act_b, const uint8_t* rng8, int32_t* obs, float* reward, uint8_t* flags,
                      int32_t* reset_obs, int64_t n, soccer_stream_t stream)
{
    if (!table || !state || !rng8 || n < 0) return SOCCER_EINVAL;
    if (pitch && pitch->slip_prob != 0.0) return SOCCER_ESLIP;          // soccer_step_table_slip takes the step draw
    soccer_step_args a = {};
    a.table = table; a.state = state; a.act_a = act_a; a.act_b = act_b; a.rng8 = rng8; a.obs = obs; a.reward = reward;
    a.flags = flags; a.reset_obs = reset_obs; a.n = n; a.auto_reset = 1;
    return step_table_ex(pitch, &a, stream);
}

int soccer_step_table_philox(const soccer_pitch* pitch, const uint16_t* table, uint32_t* state, const uint8_t* act_a,
                             const uint8_t* act_b, uint64_t seed, uint64_t step, uint64_t env_id_base, int32_t* obs,
                             float* reward, uint8_t* flags, int32_t* reset_obs, int64_t n, soccer_stream_t stream)
{
    if (!table || !state || n < 0) return SOCCER_EINVAL;
    if (pitch && pitch->slip_prob != 0.0) return SOCCER_ESLIP;
    soccer_step_args a = {};
    a.table = table; a.state = state; a.act_a = act_a; a.act_b = act_b; a.obs = obs; a.reward = reward;
    a.flags = flags; a.reset_obs = reset_obs; a.n = n; a.auto_reset = 1; a.use_philox = 1;
    a.seed = seed; a.step = step; a.env_id_base = env_id_base;
    return step_table_ex(pitch, &a, stream);
}

int soccer_step_narrow(const soccer_pitch* pitch, const uint16_t* table, uint32_t* state, const uint8_t* act_a,
                       const uint8_t* act_b, const uint8_t* rng8, uint16_t* obs16, int8_t* reward8, uint8_t* flags,
                       int64_t n, soccer_stream_t stream)
{
    if (!state || !obs16 || !reward8 || !flags || !rng8 || n < 0) return SOCCER_EINVAL;
    if (pitch && pitch->slip_prob != 0.0) return SOCCER_ESLIP;
    soccer_step_args a = {};
    a.table = table; a.state = state; a.act_a = act_a; a.act_b = act_b; a.rng8 = rng8; a.obs = reinterpret_cast<int32_t*>(obs16);
    a.reward = reinterpret_cast<float*>(reward8); a.flags = flags; a.n = n; a.auto_reset = 1; a.narrow = 1;
    return soccer_step_ex(pitch, &a, stream);
}

int soccer_step_table_packed(const soccer_pitch* pitch, const uint16_t* table, uint32_t* state, const uint8_t* joint,
                             const uint8_t* rng8, uint16_t* result, int64_t n, soccer_stream_t stream)
{
    if (!rng8) return SOCCER_EINVAL;
    return packed_impl(pitch, table, state, joint, rng8, PhiloxKey(), result, n, stream);
}

int soccer_step_table_packed_philox(const soccer_pitch* pitch, const uint16_t* table, uint32_t* state, const uint8_t* joint,
                                    uint64_t seed, uint64_t step, uint64_t env_id_base, uint16_t* result, int64_t n,
                                    soccer_stream_t stream)
{
    const PhiloxKey key = { seed, step, env_id_base };
    return packed_impl(pitch, table, state, joint, nullptr, key, result, n, stream);
}

int soccer_slip_index_bytes_host(const soccer_pitch* pitch, int64_t* bytes)
{
    if (!bytes) return SOCCER_EINVAL;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    int64_t tb; const int rc2 = table_bytes_of(P, &tb); if (rc2) return rc2;
    *bytes = 3 * (((int64_t)P.nS * 25 + 15) / 16 * 16);      // plane 0 (fp64 draws): bytes; plane 1 (32-bit draws): uint16
    return SOCCER_OK;
}

int soccer_build_slip_index(const soccer_pitch* pitch, const uint16_t* table, uint8_t* slip_index, soccer_stream_t stream)
{
    if (!table || !slip_index) return SOCCER_EINVAL;
    PitchDev P; int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    int64_t tb; rc = table_bytes_of(P, &tb); if (rc) return rc;
    const int64_t plane = ((int64_t)P.nS * 25 + 15) / 16 * 16;
    const cudaError_t e = cudaMemsetAsync(slip_index, 0, (size_t)(3 * plane), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    k_build_slip_index<<<grid_for((int64_t)P.nS * 25, 4), kThreads, 0, (cudaStream_t)stream>>>(P, P.nS, table, slip_index, plane);
    return launch_status();
}

int soccer_step_table_slip(const soccer_pitch* pitch, const uint16_t* table, const uint8_t* slip_index, uint32_t* state,
                           const uint8_t* act_a, const uint8_t* act_b, const uint8_t* rng8, const uint32_t* rng32,
                           const double* rngf64, int32_t* obs, float* reward, uint8_t* flags, int32_t* reset_obs, int64_t n,
                           soccer_stream_t stream)
{
    if (!table || !state || !act_a || !act_b || !rng8 || !obs || !reward || !flags || n < 0) return SOCCER_EINVAL;
    if (!rng32 && !rngf64) return SOCCER_EINVAL;
    soccer_step_args a = {};
    a.table = table; a.slip_index = slip_index; a.state = state; a.act_a = act_a; a.act_b = act_b; a.rng8 = rng8;
    a.rng32 = rngf64 ? nullptr : rng32; a.rngf64 = rngf64; a.obs = obs; a.reward = reward; a.flags = flags;
    a.reset_obs = reset_obs; a.n = n; a.auto_reset = 1;
    PitchDev P; int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    int64_t bytes; rc = table_bytes_of(P, &bytes); if (rc) return rc;
    if (!aligned(table, 16) || (slip_index && !aligned(slip_index, 16))) return SOCCER_EINVAL;
    if (n == 0) return SOCCER_OK;
    return step_table_slip_impl(P, bytes, &a, (cudaStream_t)stream);
}

int soccer_rollout_table_policy(const soccer_pitch* pitch, const uint16_t* table, const uint8_t* slip_index, uint32_t* state,
                                const int8_t* policy_a, const int8_t* policy_b, uint64_t seed, uint64_t step0, int32_t K,
                                uint64_t env_id_base, int32_t flip_reward, int32_t* obs, float* reward, uint8_t* flags,
                                unsigned long long* stats, int64_t n, soccer_stream_t stream)
{
    if (!table || !state || n < 0 || K < 0 || K > kMaxRolloutK) return SOCCER_EINVAL;
    PitchDev P; int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    int64_t bytes; rc = table_bytes_of(P, &bytes); if (rc) return rc;
    if (!aligned(table, 16) || (slip_index && !aligned(slip_index, 16))) return SOCCER_EINVAL;
    if (n == 0 || K == 0) return SOCCER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // Philox contract v2: the 4 envs of a thread must be one aligned group of the GLOBAL env ids
    const bool vec = (n % 4 == 0) && (env_id_base % 4 == 0) && aligned(state, 16) && (!obs || aligned(obs, 16)) &&
                     (!reward || aligned(reward, 16)) && (!flags || aligned(flags, 4));
    const RolloutArgs ra = { state, seed, step0, K, env_id_base, obs, reward, flags, stats, n, philox_round_keys(seed), flip_reward ? 1 : 0 };
    const bool streams = obs && reward && flags;
    const bool pol = policy_a || policy_b;
    const int64_t pol_bytes = 2 * (((int64_t)P.nS + 15) & ~15);
    const int64_t smem = bytes + 16 + ((pol || P.slip) ? pol_bytes : 0);
    if (smem > 227 * 1024 - 1024) return SOCCER_ETABLE;
#define SOCCER_LAUNCH_ROLLOUT_T(VEC, STR, POL, SLIP, ITEMS)                                              \
    do {                                                                                                 \
        const int e0 = allow_big_smem(k_rollout_table<VEC, STR, POL, SLIP>, smem);                       \
        if (e0) return e0;                                                                               \
        const int e1 = launch_pdl(k_rollout_table<VEC, STR, POL, SLIP>, table_grid(ITEMS, kRolloutThreads),      \
                                  kRolloutThreads, (size_t)smem, st, P, table, (uint32_t)bytes, policy_a, policy_b, ra); \
        if (e1) return e1;                                                                               \
    } while (0)
#define SOCCER_PICK_ROLLOUT_T(POL, SLIP)                                                                 \
    do {                                                                                                 \
        if (vec && streams) SOCCER_LAUNCH_ROLLOUT_T(4, true, POL, SLIP, n / 4);                          \
        else if (vec) SOCCER_LAUNCH_ROLLOUT_T(4, false, POL, SLIP, n / 4);                               \
        else if (streams) SOCCER_LAUNCH_ROLLOUT_T(1, true, POL, SLIP, n);                                \
        else SOCCER_LAUNCH_ROLLOUT_T(1, false, POL, SLIP, n);                                            \
    } while (0)
    if (P.slip) {
        // integer-threshold fast path, 4 envs per thread (k_rollout_table_slipi; 5x4 with a 12-bit, 6x4 with a 10-bit
        // bucket table); else the in-place walk, one env per thread (the walk's registers)
        SlipE E; SlipDanger dg;
        const bool dg_ok = slip_consts_host(P, &E, &dg);
        const int64_t smem_b = bytes + 16 + (pol ? pol_bytes : 0);
        int lut_bits = 12;
        if (smem_b + slip_int_lut_bytes(lut_bits) > 227 * 1024 - 2048) lut_bits = 10;
        const int64_t smem_i = smem_b + slip_int_lut_bytes(lut_bits);
        if (dg_ok && smem_i <= 227 * 1024 - 2048 && !soccer_force_slip_walk()) {
#define SOCCER_LAUNCH_ROLLOUT_I(VEC, STR, POL, ITEMS)                                                    \
            do {                                                                                         \
                const int e0 = allow_big_smem(k_rollout_table_slipi<VEC, STR, POL>, smem_i);             \
                if (e0) return e0;                                                                       \
                const int e1 = launch_pdl(k_rollout_table_slipi<VEC, STR, POL>, table_grid(ITEMS, kRolloutSlipThreads), \
                                          kRolloutSlipThreads, (size_t)smem_i, st, P, table, (uint32_t)bytes, \
                                          E, dg, slip_bits(lut_bits), policy_a, policy_b, ra);                      \
                if (e1) return e1;                                                                       \
            } while (0)
#define SOCCER_PICK_ROLLOUT_I(POL)                                                                       \
            do {                                                                                         \
                if (vec && streams) SOCCER_LAUNCH_ROLLOUT_I(4, true, POL, n / 4);                        \
                else if (vec) SOCCER_LAUNCH_ROLLOUT_I(4, false, POL, n / 4);                             \
                else if (streams) SOCCER_LAUNCH_ROLLOUT_I(1, true, POL, n);                              \
                else SOCCER_LAUNCH_ROLLOUT_I(1, false, POL, n);                                          \
            } while (0)
            if (pol) SOCCER_PICK_ROLLOUT_I(true); else SOCCER_PICK_ROLLOUT_I(false);
#undef SOCCER_PICK_ROLLOUT_I
#undef SOCCER_LAUNCH_ROLLOUT_I
        }
        else if (streams) SOCCER_LAUNCH_ROLLOUT_T(1, true, true, true, n);
        else SOCCER_LAUNCH_ROLLOUT_T(1, false, true, true, n);
    }
    else if (pol) SOCCER_PICK_ROLLOUT_T(true, false);
    else SOCCER_PICK_ROLLOUT_T(false, false);
#undef SOCCER_PICK_ROLLOUT_T
#undef SOCCER_LAUNCH_ROLLOUT_T
    return launch_status();
}

int soccer_cluster_table_bytes_host(const soccer_pitch* pitch, int64_t* bytes, int32_t* cluster_size)
{
    if (!bytes) return SOCCER_EINVAL;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    if (P.nS > 4096) return SOCCER_ETABLE;                    // 12-bit next-observation field of the 16-bit entries
    const int64_t tb = ((int64_t)P.nS * 200 + 15) / 16 * 16;
    int cl = 1;
    while ((int64_t)cl * kClusterSliceBytes < tb) cl *= 2;
    if (cl > 8) return SOCCER_ETABLE;
    *bytes = tb;
    if (cluster_size) *cluster_size = cl;
    return SOCCER_OK;
}

int soccer_build_cluster_table(const soccer_pitch* pitch, uint16_t* table, soccer_stream_t stream)
{
    if (!table) return SOCCER_EINVAL;
    PitchDev P; int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    int64_t bytes; rc = soccer_cluster_table_bytes_host(pitch, &bytes, nullptr); if (rc) return rc;
    k_build_step_table<<<grid_for((int64_t)P.nS * 100, 4), kThreads, 0, (cudaStream_t)stream>>>(P, P.nS, table);
    return launch_status();
}

int soccer_rollout_table_cluster(const soccer_pitch* pitch, const uint16_t* table, uint32_t* state, uint64_t seed,
                                 uint64_t step0, int32_t K, uint64_t env_id_base, int32_t* obs, float* reward,
                                 uint8_t* flags, unsigned long long* stats, int64_t n, soccer_stream_t stream)
{
    if (!table || !state || n < 0 || K < 0 || K > kMaxRolloutK) return SOCCER_EINVAL;
    if (pitch && pitch->slip_prob != 0.0) return SOCCER_ESLIP;
    PitchDev P; int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    int64_t bytes; int32_t cl; rc = soccer_cluster_table_bytes_host(pitch, &bytes, &cl); if (rc) return rc;
    if (!aligned(table, 16)) return SOCCER_EINVAL;
    if (n == 0 || K == 0) return SOCCER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (n % 4 == 0) && (env_id_base % 4 == 0) && aligned(state, 16) && (!obs || aligned(obs, 16)) &&
                     (!reward || aligned(reward, 16)) && (!flags || aligned(flags, 4));
    const RolloutArgs ra = { state, seed, step0, K, env_id_base, obs, reward, flags, stats, n, philox_round_keys(seed), 0 };
    const bool streams = obs && reward && flags;
    const size_t smem = (size_t)kClusterSliceBytes + 16;
#define SOCCER_LAUNCH_ROLLOUT_C(VEC, STR, ITEMS)                                                          \
    do {                                                                                                  \
        auto kern = k_rollout_table_cluster<VEC, STR>;                                                    \
        const int e0 = allow_big_smem(kern, (int64_t)smem);                                               \
        if (e0) return e0;                                                                                \
        cudaLaunchConfig_t cfg = {};                                                                      \
        cfg.blockDim = dim3(kRolloutThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;               \
        cudaLaunchAttribute attr[1];                                                                      \
        attr[0].id = cudaLaunchAttributeClusterDimension;                                                 \
        attr[0].val.clusterDim.x = (unsigned)cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1; \
        cfg.attrs = attr; cfg.numAttrs = 1;                                                               \
        cfg.gridDim = dim3((unsigned)cl);                                                                 \
        int n_clusters = 0;                                                                               \
        if (cudaOccupancyMaxActiveClusters(&n_clusters, kern, &cfg) != cudaSuccess || n_clusters < 1) {   \
            (void)cudaGetLastError(); n_clusters = sm_count() / cl;                                       \
        }                                                                                                 \
        const int64_t need = ((ITEMS) + kRolloutThreads - 1) / kRolloutThreads;                           \
        int64_t ctas = (int64_t)n_clusters * cl;                                                          \
        if (need < ctas) ctas = (need + cl - 1) / cl * cl;                                                \
        cfg.gridDim = dim3((unsigned)ctas);                                                               \
        const cudaError_t e1 = cudaLaunchKernelEx(&cfg, kern, P, table, (uint32_t)bytes, ra);             \
        if (e1 != cudaSuccess) return (int)e1;                                                            \
    } while (0)
    if (vec && streams) SOCCER_LAUNCH_ROLLOUT_C(4, true, n / 4);
    else if (vec) SOCCER_LAUNCH_ROLLOUT_C(4, false, n / 4);
    else if (streams) SOCCER_LAUNCH_ROLLOUT_C(1, true, n);
    else SOCCER_LAUNCH_ROLLOUT_C(1, false, n);
#undef SOCCER_LAUNCH_ROLLOUT_C
    return launch_status();
}

int soccer_rollout_table(const soccer_pitch* pitch, const uint16_t* table, uint32_t* state, uint64_t seed,
                         uint64_t step0, int32_t K, uint64_t env_id_base, int32_t* obs, float* reward,
                         uint8_t* flags, unsigned long long* stats, int64_t n, soccer_stream_t stream)
{
    return soccer_rollout_table_policy(pitch, table, nullptr, state, nullptr, nullptr, seed, step0, K, env_id_base, 0, obs,
                                       reward, flags, stats, n, stream);
}

int soccer_convert_state(const soccer_pitch* pitch, const uint32_t* in, uint32_t* out, int32_t to_layout, int64_t n,
                         soccer_stream_t stream)
{
    if (!in || !out || n < 0 || (to_layout != SOCCER_LAYOUT_CELL && to_layout != SOCCER_LAYOUT_INDEX)) return SOCCER_EINVAL;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    if (n == 0) return SOCCER_OK;
    k_convert_state<<<grid_for(n, 8), kThreads, 0, (cudaStream_t)stream>>>(P, P.nS, in, out, to_layout, n);
    return launch_status();
}

int soccer_bench_stream_mix(uint32_t* state, const uint8_t* act_a, const uint8_t* act_b, const uint8_t* rng8,
                            int32_t* obs, float* reward, uint8_t* flags, int64_t n, int32_t mode, soccer_stream_t stream)
{
    if (!state || !act_a || !act_b || !rng8 || !obs || !reward || !flags || n < 4 || (n & 3)) return SOCCER_EINVAL;
    if (!aligned(state, 16) || !aligned(obs, 16) || !aligned(reward, 16) || !aligned(act_a, 4) || !aligned(act_b, 4) ||
        !aligned(rng8, 4) || !aligned(flags, 4))
        return SOCCER_EINVAL;
    const int grid = table_grid(n / 4);
    cudaStream_t st = (cudaStream_t)stream;
    switch (mode) {
    case 0: k_stream_mix_probe<0><<<grid, kTableThreads, 0, st>>>(state, act_a, act_b, rng8, obs, reward, flags, n / 4); break;
    case 1: k_stream_mix_probe<1><<<grid, kTableThreads, 0, st>>>(state, act_a, act_b, rng8, obs, reward, flags, n / 4); break;
    case 2: k_stream_mix_probe<2><<<grid, kTableThreads, 0, st>>>(state, act_a, act_b, rng8, obs, reward, flags, n / 4); break;
    default: return SOCCER_EINVAL;
    }
    return launch_status();
}

int soccer_bench_rollout_probe(uint32_t* state, int32_t K, int32_t* obs, float* reward, uint8_t* flags, int64_t n,
                               int32_t mode, soccer_stream_t stream)
{
    if (!state || !obs || !reward || !flags || K < 0 || n < 8 || (n & 7)) return SOCCER_EINVAL;
    if (!aligned(state, 16) || !aligned(obs, 16) || !aligned(reward, 16) || !aligned(flags, 4)) return SOCCER_EINVAL;
    const RolloutArgs ra = { state, 0, 0, K, 0, obs, reward, flags, nullptr, n, philox_round_keys(0), 0 };
    const int grid = table_grid(n / 4, kRolloutThreads);
    cudaStream_t st = (cudaStream_t)stream;
    switch (mode) {
    case 0: k_rollout_probe<0><<<grid, kRolloutThreads, 0, st>>>(ra); break;
    case 1: k_rollout_probe<1><<<grid, kRolloutThreads, 0, st>>>(ra); break;
    case 2: k_rollout_probe<2><<<grid, kRolloutThreads, 0, st>>>(ra); break;
    case 3: k_rollout_probe<3><<<grid, kRolloutThreads, 0, st>>>(ra); break;
    case 4: k_rollout_probe<4><<<grid, kRolloutThreads, 0, st>>>(ra); break;
    case 5: k_rollout_probe<5><<<grid, kRolloutThreads, 0, st>>>(ra); break;
    default: return SOCCER_EINVAL;
    }
    return launch_status();
}

int soccer_step_stats(const uint8_t* flags, const float* reward, int64_t n, unsigned long long* stats,
                      soccer_stream_t stream)
{
    if (!flags || !stats || n < 0) return SOCCER_EINVAL;
    if (n == 0) return SOCCER_OK;
    k_step_stats<<<grid_for((n + 15) / 16, 8), kThreads, 0, (cudaStream_t)stream>>>(flags, reward, n, stats);
    return launch_status();
}

// launch shape of the fused replay: spread few envs over many SMs (one warp per CTA for a 4096-env
// batch: the step chain is latency-bound, so every warp gets an SM sub-partition of its own), grow the
// CTAs up to max_threads once every SM has one
static void replay_shape(int64_t items, int max_threads, int cta_cap, int* grid, int* block)
{
    const int64_t slots = (items + 31) / 32;
    int64_t wpc = (slots + cta_cap - 1) / cta_cap;
    if (wpc < 1) wpc = 1;
    if (wpc > max_threads / 32) wpc = max_threads / 32;
    const int64_t need = (slots + wpc - 1) / wpc;
    *block = (int)wpc * 32;
    *grid = (int)(need < cta_cap ? (need < 1 ? 1 : need) : cta_cap);
}

int soccer_step_many(const soccer_pitch* pitch, const uint16_t* table, uint32_t* state, int32_t T,
                     const uint8_t* act_a, const uint8_t* act_b, const uint8_t* rng8, int32_t* obs, float* reward,
                     uint8_t* flags, int32_t* reset_obs, int64_t n, soccer_stream_t stream)
{
    if (T < 0 || n < 0 || !state || !act_a || !act_b || !rng8 || !obs || !reward || !flags) return SOCCER_EINVAL;
    if (pitch && pitch->slip_prob != 0.0) return SOCCER_ESLIP;
    PitchDev P; int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    int64_t bytes = 0;
    if (table) {
        rc = table_bytes_of(P, &bytes); if (rc) return rc;
        if (!aligned(table, 16)) return SOCCER_EINVAL;
    }
    if (T == 0 || n == 0) return SOCCER_OK;
    if (T == 1)
        return table ? soccer_step_table(pitch, table, state, act_a, act_b, rng8, obs, reward, flags, reset_obs, n, stream)
                     : soccer_step(pitch, state, act_a, act_b, rng8, obs, reward, flags, reset_obs, n, stream);
    cudaStream_t st = (cudaStream_t)stream;
    // rows t of the [T][n] arrays start at t * n: 128-bit / 32-bit accesses need n % 4 == 0 as well
    const bool vec = (n % 4 == 0) && aligned(state, 16) && aligned(obs, 16) && aligned(reward, 16) &&
                     (!reset_obs || aligned(reset_obs, 16)) && aligned(act_a, 4) && aligned(act_b, 4) &&
                     aligned(rng8, 4) && aligned(flags, 4);
    const ReplayArgs ra = { state, act_a, act_b, rng8, T, obs, reward, flags, reset_obs, n };
    int grid = 1, block = 32;
#ifndef SOCCER_REPLAY_VEC1_WARPS
#define SOCCER_REPLAY_VEC1_WARPS 4      // measured: one env per thread wins up to 4 warps per SM (profiles/r01d_ab_replay4.log)
#endif
#define SOCCER_LAUNCH_REPLAY_T(VEC, RO, U, ITEMS)                                                         \
    do {                                                                                                  \
        const int e0 = allow_big_smem(k_replay_table<VEC, RO, U>, bytes + 16);                            \
        if (e0) return e0;                                                                                \
        replay_shape(ITEMS, kRolloutThreads, sms, &grid, &block);                                         \
        k_replay_table<VEC, RO, U><<<grid, block, bytes + 16, st>>>(P, table, (uint32_t)bytes, ra);       \
    } while (0)
#define SOCCER_LAUNCH_REPLAY(VEC, RO, U, ITEMS)                                                           \
    do {                                                                                                  \
        static const int nb = resident_blocks(k_replay<VEC, RO, U>);                                      \
        replay_shape(ITEMS, kThreads, sms * nb, &grid, &block);                                           \
        k_replay<VEC, RO, U><<<grid, block, 0, st>>>(P, ra);                                              \
    } while (0)
#define SOCCER_REPLAY_PICK(LAUNCH)                                                                        \
    do {                                                                                                  \
        if (!vec || tiny) { if (reset_obs) LAUNCH(1, true, 8, n); else LAUNCH(1, false, 8, n); }          \
        else if (!fills) { if (reset_obs) LAUNCH(4, true, 8, n / 4); else LAUNCH(4, false, 8, n / 4); }   \
        else { if (reset_obs) LAUNCH(4, true, 4, n / 4); else LAUNCH(4, false, 4, n / 4); }               \
    } while (0)
    const int sms = sm_count();
    // tiny: one env per thread still leaves the SMs at <= 4 warps -> shortest dependent chain per warp wins;
    // fills: every SM gets a full CTA of 4-env threads -> HBM-bound regime, 4 rows of look-ahead suffice
    const bool tiny = (n + 31) / 32 <= (int64_t)sms * SOCCER_REPLAY_VEC1_WARPS;
    const bool fills = (n / 4 + 31) / 32 >= (int64_t)sms * (table ? kRolloutThreads / 32 : kThreads / 32);
    if (table) SOCCER_REPLAY_PICK(SOCCER_LAUNCH_REPLAY_T);
    else SOCCER_REPLAY_PICK(SOCCER_LAUNCH_REPLAY);
#undef SOCCER_REPLAY_PICK
#undef SOCCER_LAUNCH_REPLAY_T
#undef SOCCER_LAUNCH_REPLAY
    return launch_status();
}

int soccer_step_host_scratch_bytes_host(int64_t n, int64_t* bytes)
{
    if (!bytes || n < 0) return SOCCER_EINVAL;
    *bytes = host_scratch(nullptr, n).bytes;
    return SOCCER_OK;
}

int soccer_step_host(const soccer_pitch* pitch, const soccer_step_host_args* a)
{
    if (!a || !a->state || !a->scratch || !a->h_act_a || !a->h_rng8 || !a->h_obs || a->n < 0 || a->n_chunks < 1 ||
        a->narrow < 0 || a->narrow > SOCCER_HOST_PACKED)
        return SOCCER_EINVAL;
    const bool packed = a->narrow == SOCCER_HOST_PACKED;
    if (packed ? !a->table : (!a->h_act_b || !a->h_reward || !a->h_flags)) return SOCCER_EINVAL;
    if (a->d2h_zero_copy && !packed) return SOCCER_EINVAL;
    if (pitch && pitch->slip_prob != 0.0) return SOCCER_ESLIP;
    if (a->n == 0) return SOCCER_OK;
    cudaStream_t s_in = (cudaStream_t)a->s_in, s_k = (cudaStream_t)a->s_compute, s_out = (cudaStream_t)a->s_out;
    const HostScratch sc = host_scratch(a->scratch, a->n);
    // slices below 64 Ki envs only add launch / copy latency: small batches go through in one piece
    int64_t chunk = round_up((a->n + a->n_chunks - 1) / a->n_chunks, 256);
    if (chunk < 65536) chunk = 65536;
#define SOCCER_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = (int)e_; goto done; } } while (0)
    int rc = SOCCER_OK;
    cudaEvent_t ev_prev_k = nullptr, ev_prev_out = nullptr, ev_in = nullptr, ev_k = nullptr;
    // cross-call hazards on the scratch buffers: this call's uploads overwrite inputs the previous
    // call's kernels read, and its kernels overwrite outputs the previous call's downloads read
    SOCCER_CUDA(cudaEventCreateWithFlags(&ev_prev_k, cudaEventDisableTiming));
    SOCCER_CUDA(cudaEventCreateWithFlags(&ev_prev_out, cudaEventDisableTiming));
    SOCCER_CUDA(cudaEventRecord(ev_prev_k, s_k));
    SOCCER_CUDA(cudaStreamWaitEvent(s_in, ev_prev_k, 0));
    SOCCER_CUDA(cudaEventRecord(ev_prev_out, s_out));
    SOCCER_CUDA(cudaStreamWaitEvent(s_k, ev_prev_out, 0));
    for (int64_t lo = 0; lo < a->n; lo += chunk) {
        const int64_t m = a->n - lo < chunk ? a->n - lo : chunk;
        SOCCER_CUDA(cudaMemcpyAsync(sc.a + lo, a->h_act_a + lo, m, cudaMemcpyHostToDevice, s_in));   // packed: joint bytes
        if (!packed) SOCCER_CUDA(cudaMemcpyAsync(sc.b + lo, a->h_act_b + lo, m, cudaMemcpyHostToDevice, s_in));
        SOCCER_CUDA(cudaMemcpyAsync(sc.r + lo, a->h_rng8 + lo, m, cudaMemcpyHostToDevice, s_in));
        SOCCER_CUDA(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
        SOCCER_CUDA(cudaEventRecord(ev_in, s_in));
        SOCCER_CUDA(cudaStreamWaitEvent(s_k, ev_in, 0));
        SOCCER_CUDA(cudaEventDestroy(ev_in)); ev_in = nullptr;
        if (packed)
            rc = soccer_step_table_packed(pitch, a->table, a->state + lo, sc.a + lo, sc.r + lo,
                                          a->d2h_zero_copy ? (uint16_t*)a->h_obs + lo : sc.obs16 + lo, m, (soccer_stream_t)s_k);
        else if (a->narrow == SOCCER_HOST_NARROW)
            rc = soccer_step_narrow(pitch, a->table, a->state + lo, sc.a + lo, sc.b + lo, sc.r + lo, sc.obs16 + lo,
                                    sc.rew8 + lo, sc.f + lo, m, (soccer_stream_t)s_k);
        else if (a->table)
            rc = soccer_step_table(pitch, a->table, a->state + lo, sc.a + lo, sc.b + lo, sc.r + lo, sc.obs + lo,
                                   sc.rew + lo, sc.f + lo, nullptr, m, (soccer_stream_t)s_k);
        else
            rc = soccer_step(pitch, a->state + lo, sc.a + lo, sc.b + lo, sc.r + lo, sc.obs + lo, sc.rew + lo,
                             sc.f + lo, nullptr, m, (soccer_stream_t)s_k);
        if (rc) goto done;
        if (packed && a->d2h_zero_copy) continue;               // the kernel wrote the slice's results into h_obs itself
        SOCCER_CUDA(cudaEventCreateWithFlags(&ev_k, cudaEventDisableTiming));
        SOCCER_CUDA(cudaEventRecord(ev_k, s_k));
        SOCCER_CUDA(cudaStreamWaitEvent(s_out, ev_k, 0));
        SOCCER_CUDA(cudaEventDestroy(ev_k)); ev_k = nullptr;
        if (packed) {
            SOCCER_CUDA(cudaMemcpyAsync((uint16_t*)a->h_obs + lo, sc.obs16 + lo, 2 * m, cudaMemcpyDeviceToHost, s_out));
            continue;
        }
        if (a->narrow) {
            SOCCER_CUDA(cudaMemcpyAsync((uint16_t*)a->h_obs + lo, sc.obs16 + lo, 2 * m, cudaMemcpyDeviceToHost, s_out));
            SOCCER_CUDA(cudaMemcpyAsync((int8_t*)a->h_reward + lo, sc.rew8 + lo, m, cudaMemcpyDeviceToHost, s_out));
        } else {
            SOCCER_CUDA(cudaMemcpyAsync((int32_t*)a->h_obs + lo, sc.obs + lo, 4 * m, cudaMemcpyDeviceToHost, s_out));
            SOCCER_CUDA(cudaMemcpyAsync((float*)a->h_reward + lo, sc.rew + lo, 4 * m, cudaMemcpyDeviceToHost, s_out));
        }
        SOCCER_CUDA(cudaMemcpyAsync(a->h_flags + lo, sc.f + lo, m, cudaMemcpyDeviceToHost, s_out));
    }
done:
#undef SOCCER_CUDA
    if (ev_prev_k) cudaEventDestroy(ev_prev_k);
    if (ev_prev_out) cudaEventDestroy(ev_prev_out);
    if (ev_in) cudaEventDestroy(ev_in);
    if (ev_k) cudaEventDestroy(ev_k);
    return rc;
}

int soccer_dense(const soccer_pitch* pitch, const int8_t* policy_a, const int8_t* policy_b, double* Pmat,
                 double* Rmat, soccer_stream_t stream)
{
    if (!Pmat || !Rmat) return SOCCER_EINVAL;
    if (policy_a && policy_b) return SOCCER_EPOLICY;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    soccer_pitch_info info; fill_info(pitch, &info);
    const int nS = info.nS, nkeys = (policy_a || policy_b) ? 5 : 25;
    // goal states (SIM:91-103): holder in one of 2*n_goal_rows goal cells, the other player on any
    // field cell, either player holding
    const int n_goal_states = 2 * (2 * info.n_goal_rows) * info.n_field_cells;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(Pmat, 0, sizeof(double) * (size_t)nS * nS * nkeys, st);
    if (e != cudaSuccess) return (int)e;
    k_dense<<<grid_for((int64_t)nS * nkeys, 4), kThreads, 0, st>>>(P, nS, n_goal_states, policy_a, policy_b, Pmat, Rmat);
    return launch_status();
}

int soccer_bellman_q(const soccer_pitch* pitch, const int8_t* policy_a, const int8_t* policy_b, const double* V,
                     double gamma, double* Q, soccer_stream_t stream)
{
    if (!V || !Q) return SOCCER_EINVAL;
    if (policy_a && policy_b) return SOCCER_EPOLICY;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    const int nkeys = (policy_a || policy_b) ? 5 : 25;
    k_bellman_q<<<grid_for((int64_t)P.nS * nkeys, 8), kThreads, 0, (cudaStream_t)stream>>>(P, P.nS, policy_a, policy_b, V,
                                                                                           gamma, Q);
    return launch_status();
}

int soccer_plan_workspace_bytes_host(const soccer_pitch* pitch, int64_t* bytes)
{
    if (!bytes) return SOCCER_EINVAL;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    *bytes = 2 * (int64_t)P.nS * 8 + 32;
    return SOCCER_OK;
}

int soccer_plan(const soccer_pitch* pitch, const int8_t* policy_a, const int8_t* policy_b, const int32_t* pi_in,
                double theta, double gamma, int32_t max_sweeps, double* V_out, double* Q_out, int32_t* pi_out,
                int32_t* sweeps_out, void* workspace, soccer_stream_t stream)
{
    if (!V_out || !sweeps_out || !workspace || max_sweeps < 1 || !aligned(workspace, 8)) return SOCCER_EINVAL;
    if (!pi_in && !Q_out) return SOCCER_EINVAL;                     // value iteration returns Q
    if (policy_a && policy_b) return SOCCER_EPOLICY;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ws = 2 * (int64_t)P.nS * 8 + 32;
    cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)ws, st);   // V = 0 (PL:6, 22), delta slots = 0
    if (e != cudaSuccess) return (int)e;
    PlanArgs a;
    a.nS = P.nS; a.policy_a = policy_a; a.policy_b = policy_b; a.pi_in = pi_in;
    a.theta = theta; a.gamma = gamma; a.max_sweeps = max_sweeps;
    a.v0 = reinterpret_cast<double*>(workspace); a.v1 = a.v0 + P.nS;
    a.delta = reinterpret_cast<unsigned long long*>(a.v1 + P.nS);
    a.V_out = V_out; a.Q_out = Q_out; a.pi_out = pi_out; a.sweeps_out = sweeps_out;
    // cooperative launch: every CTA must be resident for the grid-wide barriers
    static const int nb = resident_blocks(k_plan);
    const int nkeys = (policy_a || policy_b) ? 5 : 25;
    const int grid = grid_for((int64_t)P.nS * (pi_in ? 1 : nkeys), nb);
    void* args[] = { (void*)&P, (void*)&a };
    e = cudaLaunchCooperativeKernel((const void*)k_plan, dim3((unsigned)grid), dim3(kThreads), args, 0, st);
    return (int)e;
}

int soccer_dense_q(const soccer_pitch* pitch, const int8_t* policy_a, const int8_t* policy_b, const double* v, double gamma,
                   double* q, soccer_stream_t stream)
{
    if (!v || !q) return SOCCER_EINVAL;
    if (policy_a && policy_b) return SOCCER_EPOLICY;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    soccer_pitch_info info; fill_info(pitch, &info);
    const int nkeys = (policy_a || policy_b) ? 5 : 25;
    const int n_goal_states = 2 * (2 * info.n_goal_rows) * info.n_field_cells;
    k_dense_q<<<grid_for((int64_t)P.nS * nkeys, 8), kThreads, 0, (cudaStream_t)stream>>>(P, P.nS, n_goal_states, policy_a, policy_b,
                                                                                         v, gamma, q);
    return launch_status();
}

int soccer_policy_eval_workspace_bytes_host(const soccer_pitch* pitch, int32_t nkeys, int64_t* bytes)
{
    if (!bytes || (nkeys != 5 && nkeys != 25)) return SOCCER_EINVAL;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    *bytes = 2 * (int64_t)P.nS * 8 + 2 * (int64_t)P.nS * nkeys * 8 + 32;
    return SOCCER_OK;
}

int soccer_policy_eval(const soccer_pitch* pitch, const int8_t* policy_a, const int8_t* policy_b, const double* policy,
                       const double* v_init, double theta, double gamma, int32_t max_sweeps, double* V_out,
                       int32_t* sweeps_out, void* workspace, soccer_stream_t stream)
{
    if (!policy || !V_out || !sweeps_out || !workspace || max_sweeps < 0 || !aligned(workspace, 8)) return SOCCER_EINVAL;
    if (policy_a && policy_b) return SOCCER_EPOLICY;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    soccer_pitch_info info; fill_info(pitch, &info);
    cudaStream_t st = (cudaStream_t)stream;
    const int nkeys = (policy_a || policy_b) ? 5 : 25;
    PolicyEvalArgs a;
    a.nS = P.nS; a.n_goal_states = 2 * (2 * info.n_goal_rows) * info.n_field_cells;
    a.policy_a = policy_a; a.policy_b = policy_b; a.policy = policy;
    a.theta = theta; a.gamma = gamma; a.max_sweeps = max_sweeps;
    a.v0 = reinterpret_cast<double*>(workspace); a.v1 = a.v0 + P.nS;
    a.part = a.v1 + P.nS;
    a.delta = reinterpret_cast<unsigned long long*>(a.part + 2 * (int64_t)P.nS * nkeys);
    a.V_out = V_out; a.sweeps_out = sweeps_out;
    cudaError_t e = cudaMemsetAsync(a.delta, 0, 32, st);
    if (e != cudaSuccess) return (int)e;
    e = v_init ? cudaMemcpyAsync(a.v0, v_init, sizeof(double) * (size_t)P.nS, cudaMemcpyDeviceToDevice, st)
               : cudaMemsetAsync(a.v0, 0, sizeof(double) * (size_t)P.nS, st);         // PL:58
    if (e != cudaSuccess) return (int)e;
    static const int nb = resident_blocks(k_policy_eval);
    const int grid = grid_for((int64_t)P.nS * nkeys, nb);
    void* args[] = { (void*)&P, (void*)&a };
    e = cudaLaunchCooperativeKernel((const void*)k_policy_eval, dim3((unsigned)grid), dim3(kThreads), args, 0, st);
    return (int)e;
}

int soccer_stats_allreduce_p2p_bytes_host(int64_t* bytes)
{
    if (!bytes) return SOCCER_EINVAL;
    *bytes = 2 * 16 * 8 * 8;
    return SOCCER_OK;
}

int soccer_stats_allreduce_p2p(const uint64_t* peer_ptrs, int32_t rank, int32_t world, uint64_t epoch,
                               unsigned long long* stats, soccer_stream_t stream)
{
    if (!peer_ptrs || !stats || world < 1 || world > 16 || rank < 0 || rank >= world || epoch == 0) return SOCCER_EINVAL;
    P2PStatsArgs a = {};
    for (int r = 0; r < world; ++r) {
        if (!peer_ptrs[r] || (peer_ptrs[r] & 7u)) return SOCCER_EINVAL;
        a.peer[r] = reinterpret_cast<unsigned long long*>(peer_ptrs[r]);
    }
    a.rank = rank; a.world = world; a.epoch = epoch; a.stats = stats;
    k_stats_allreduce_p2p<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    return launch_status();
}

int soccer_step_speculate(const soccer_pitch* pitch, uint32_t state_word, const int8_t* policy_a, const int8_t* policy_b,
                          const double* u, uint32_t* records, uint32_t seq, soccer_stream_t stream)
{
    if (!records || !aligned(records, 16) || seq >= (1u << 24)) return SOCCER_EINVAL;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    if (P.slip && !u) return SOCCER_ESLIP;                         // the draw is not 2 bits: it has to be handed over
    if (u && !(*u >= 0.0 && *u < 1.0)) return SOCCER_EINVAL;
    const uint32_t a = state_word & 0xFFu, b = (state_word >> 8) & 0xFFu;
    if (((a | b) & kGoalBit) || (state_word & kNeedsReset) || a >= (uint32_t)P.F || b >= (uint32_t)P.F || a == b)
        return SOCCER_EINVAL;                                      // field-cell states of a running episode only
    if (P.slip) k_step_speculate<true><<<1, 128, 0, (cudaStream_t)stream>>>(P, state_word, policy_a, policy_b, records, seq, *u);
    else k_step_speculate<false><<<1, 128, 0, (cudaStream_t)stream>>>(P, state_word, policy_a, policy_b, records, seq, 0.0);
    return launch_status();
}

int soccer_slip_danger_host(const soccer_pitch* pitch, uint32_t draws[12], int32_t* n)
{
    if (!draws || !n) return SOCCER_EINVAL;
    PitchDev P; const int rc = make_pitch_dev(pitch, &P); if (rc) return rc;
    SlipE E; SlipDanger dg;
    const bool ok = slip_consts_host(P, &E, &dg);
    for (uint32_t i = 0; i < dg.n && i < 12; ++i) draws[i] = dg.r[i];
    *n = ok ? (int32_t)dg.n : -1;
    return SOCCER_OK;
}

} // extern "C"

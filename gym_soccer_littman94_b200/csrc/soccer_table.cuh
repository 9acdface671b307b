// soccer_table.cuh -- K1 / K2 with the whole transition table resident in shared memory.
//
// The table is the reference's P (SIM:167-293) for slip_prob == 0, produced ON THE DEVICE by the
// rules path (k_build_step_table -> resolve()/finish_step()), compacted to one int16 per
// (state, joint action, 2-bit draw):
//     table[obs*100 + (aa*5+ab)*4 + r] = next_obs | nlog2 << 12 | (reward & 3) << 14
// so that a SIGN-EXTENDING 16-bit shared-memory load yields next_obs = e & 0xFFF and
// reward = e >> 14 (arithmetic: -1, 0, +1) with no further decoding.  761 * 100 * 2 = 152,200 bytes
// for the 5x4 pitch, 1105 * 200 = 221,000 for 6x4 (the largest that fits): the 227 KB of one SM hold it once, hence ONE persistent 1024-thread CTA per
// SM; the copy global -> shared is a TMA bulk copy (cp.async.bulk + mbarrier) that overlaps the
// first HBM loads.  States use SOCCER_LAYOUT_INDEX: obs | timestep << 16.
// SIM = /root/reference/gym_soccer/envs/soccer_simultaneous_env.py.
#pragma once
#include "soccer_rules.cuh"

namespace soccer {

constexpr int kTableThreads = 1024;          // one CTA per SM owns the whole shared memory
#ifndef SOCCER_ROLLOUT_THREADS
#define SOCCER_ROLLOUT_THREADS 512           // K2 keeps 4 envs x 4 Philox words in registers: 128 registers per thread
#endif
constexpr int kRolloutThreads = SOCCER_ROLLOUT_THREADS;
constexpr int kMaxTableStates = 1130;        // rows * 200 B + start observations + table policies <= 227 KB (next_obs has 12 bits)
constexpr uint32_t kTblObsMask = 0xFFFu;
constexpr uint32_t kTruncWord = (uint32_t)kMaxT << 16;   // (obs | t<<16) >= this  <=>  t >= 100

__global__ void __launch_bounds__(kThreads)
k_build_step_table(const PitchDev P, int32_t nS, uint16_t* __restrict__ table)
{
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    build_cand_lut(lut, P);
    const int64_t total = (int64_t)nS * 100;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        if (i < 100) { table[i] = 0; continue; }   // row 0 = the terminal observation: absorbing, reward 0 (SIM:300-301)
        const uint32_t r = (uint32_t)(i & 3), ja = (uint32_t)((i >> 2) % 25), aa = ja / 5u, ab = ja % 5u;
        const uint32_t st = obs_to_packed(P, (int32_t)(i / 100));
        const uint32_t a = st & 0xFFu, b = (st >> 8) & 0xFFu, p = (st >> 24) & 1u;
        const Resolved o = resolve(lut, a, b, p, aa, ab, aa == 0, ab == 0, r);
        const StepOut f = finish_step<false>(P, o, 0u, 0u, 0u, false);
        const uint32_t rew2 = f.reward > 0.0f ? 1u : (f.reward < 0.0f ? 3u : 0u);
        table[i] = (uint16_t)((uint32_t)f.obs | (o.nlog2 << 12) | (rew2 << 14));
    }
}

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// Shared-memory image: [0, table_bytes) the table, then 16 bytes = the 4 start observations
// (SIM:146-165) indexed by the 2-bit reset draw.
__device__ __forceinline__ void stage_table(uint8_t* smem, const uint16_t* gtable, uint32_t bytes, uint64_t* bar,
                                            const PitchDev& P)
{
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, bytes);
        const uint32_t chunk = 16384;
        for (uint32_t off = 0; off < bytes; off += chunk) {
            const uint32_t nb = bytes - off < chunk ? bytes - off : chunk;
            tma_load_1d(smem + off, reinterpret_cast<const uint8_t*>(gtable) + off, nb, bar);
        }
    }
    if (threadIdx.x < 4) reinterpret_cast<int32_t*>(smem + bytes)[threadIdx.x] = P.isd_obs[threadIdx.x];
    __syncthreads();    // barrier init + isd words visible to all before anyone polls / reads
}
// bounded spin -> trap, so a fault cannot hang the GPU
__device__ __forceinline__ void wait_table(uint64_t* bar)
{
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, 0)) {
        if (++spins > (1u << 22)) __trap();
    }
}

// One env-step through the table.  `jr` = (aa*5+ab)*4 + r (0..99), `rsel4` = 4 * reset draw.
struct TblCtx { uint32_t tbl, isd, last, maxobs; };    // shared-window addresses of the table / the isd words; last entry index; nS - 1
struct TblOut { uint32_t state, obs, flags; int32_t rew_i; uint32_t reset_obs; };
// everything after the table look-up: observation, reward, done / truncated, fused reset (SIM:399-424, 493)
__device__ __forceinline__ TblOut table_finish(const TblCtx& c, uint32_t s, int32_t e, uint32_t rsel4)
{
    uint32_t ro;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ro) : "r"(c.isd + rsel4));            // SIM:414-415
    TblOut o;
    o.obs = (uint32_t)e & kTblObsMask;
    o.rew_i = e >> 14;                                     // -1 / 0 / +1 (SIM:235-240)
    const bool done = o.obs == 0;                          // SIM:493: goal -> observation 0
    const uint32_t s1 = s + 0x10000u;                      // timestep + 1 in place (SIM:399)
    const bool trunc = s1 >= kTruncWord;                   // SIM:404
    const bool reset = done | trunc;                       // SIM:406
    o.flags = (done ? 1u : 0u) + (trunc ? 2u : 0u);
    o.state = reset ? ro : ((s1 & 0xFFFF0000u) | o.obs);
    o.reset_obs = reset ? ro : o.obs;
    return o;
}

__device__ __forceinline__ TblOut table_step(const TblCtx& c, uint32_t s, uint32_t jr, uint32_t rsel4)
{
    // the clamp keeps a corrupt state word / action byte inside the table
    const uint32_t idx = min((s & 0xFFFFu) * 100u + jr, c.last);
    int32_t e;
    // volatile: ordered after the (volatile) mbarrier wait that publishes the table
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(e) : "r"(c.tbl + idx * 2u));          // sign-extending LDS.S16
    return table_finish(c, s, e, rsel4);
}

// ---- slip_prob > 0 through the SAME table (SIM:203-256).
// The reference lists, for the 9 slip combinations in the order of SIM:209-223, the outcomes of the slipped move
// pair (ma, mb) with probability mp * nsp and draws with categorical_sample: the first entry whose running fp64
// sum exceeds u (gym 0.26.2: argmax(cumsum > u), all-False -> 0).  The outcomes of a move pair are exactly the
// slip-0 transitions of (state, ma, mb) -- a slipped move is NOOP iff the action is (slip_move), so the NOOP-keyed
// cases 2/3 of SIM:330-344 agree -- i.e. the table row of the state: entry (ma*5+mb)*4 + r, whose bits 12..13
// hold log2(#outcomes).  Per env-step: 9 look-ups for the outcome counts, the sequential sum in the reference's
// order with __dadd_rn / __dmul_rn (no FMA contraction), one look-up for the chosen entry.  No rules evaluation,
// no per-combination branch.
// LD(byte offset in the row) returns the 16-bit entry zero-extended.
template <class LD>
__device__ __forceinline__ uint32_t slip_pick(const PitchDev& P, const LD& ld, uint32_t aa, uint32_t ab, double u)
{
    aa = min(aa, 4u); ab = min(ab, 4u);
    // byte offsets inside the row of the r = 0 entry: (ma * 5 + mb) * 4 entries * 2 bytes
    const uint32_t oa[3] = { aa * 40u, slip_move(aa, 0) * 40u, slip_move(aa, 1) * 40u };
    const uint32_t ob[3] = { ab * 8u, slip_move(ab, 0) * 8u, slip_move(ab, 1) * 8u };
    double cs = 0.0;
    bool found = false, have_first = false;
    uint32_t pick = 0, first = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const double mp = P.mp[k];
        if (mp == 0.0) continue;                                            // SIM:226-227 (same for every env)
        const uint32_t off = oa[combo_a(k)] + ob[combo_b(k)];
        const uint32_t nl = (ld(off) >> 12) & 3u;                           // log2(#outcomes) of this move pair
        if (!have_first) { first = off; have_first = true; }
        const double pr = __dmul_rn(mp, nl == 2u ? 0.25 : (nl == 1u ? 0.5 : 1.0));   // SIM:241 (exact: power of two)
        cs = __dadd_rn(cs, pr);
        if (!found && cs > u) { found = true; pick = off; }                 // outcome slot 0 (draw value 0)
        if (nl) {                                                           // rare: a collision with 2 or 4 outcomes
            cs = __dadd_rn(cs, pr);
            if (!found && cs > u) { found = true; pick = off + (nl == 1u ? 4u : 2u); }   // slot 1: r = 2 of 2, r = 1 of 4
            if (nl == 2u) {
                cs = __dadd_rn(cs, pr);
                if (!found && cs > u) { found = true; pick = off + 4u; }
                cs = __dadd_rn(cs, pr);
                if (!found && cs > u) { found = true; pick = off + 6u; }
            }
        }
    }
    return found ? pick : first;                                            // all-False -> index 0
}

struct LdShared {
    uint32_t row;
    __device__ __forceinline__ uint32_t operator()(uint32_t off) const
    {
        uint32_t e;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(row + off));
        return e;
    }
};
struct LdGlobal {
    const uint8_t* row;
    __device__ __forceinline__ uint32_t operator()(uint32_t off) const
    {
        return (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(row + off));
    }
};

// Shared-memory fast form of slip_pick (same result, ~4x fewer instructions).  The end-of-combination sums
// E_k are non-decreasing, so "first entry whose running sum exceeds u" needs no found flag: the candidate moves
// on to combination k + 1 exactly while E_k <= u.  Per combination: one byte load of the entry's high byte
// (outcome count in bits 4..5), the probability mp_k * nsp from a 9 x 3 fp64 shared-memory table (no select,
// no multiply), one DADD, one DSETP, one select.  A combination with 2 or 4 outcomes (< 3 % of them) takes a
// short branch that adds its partial sums and moves the candidate to the slot inside it.  A zero-probability
// combination (SIM:226-227) adds 0 and so can never become the pick.
__device__ __forceinline__ TblOut table_step_slip(const TblCtx& c, const SlipCtx& sc, uint32_t s, uint32_t aa, uint32_t ab,
                                                  double u, uint32_t rsel4)
{
    const uint32_t row = c.tbl + min(s & 0xFFFFu, c.maxobs) * 200u;
    aa = min(aa, 4u); ab = min(ab, 4u);
    // shared addresses of the r = 0 entry of a move pair: row + (ma * 5 + mb) * 8
    const uint32_t ra[3] = { row + aa * 40u, row + slip_move(aa, 0) * 40u, row + slip_move(aa, 1) * 40u };
    const uint32_t ob[3] = { ab * 8u, slip_move(ab, 0) * 8u, slip_move(ab, 1) * 8u };
    double E = 0.0;
    bool le = true;                     // E_{k-1} <= u: the pick is not before combination k
    uint32_t pick = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        uint32_t ent = ra[combo_a(k)] + ob[combo_b(k)];
        const uint32_t hi = lds_u8_r(ent + 1u);                                   // bits 4..5 = log2(#outcomes)
        const uint32_t nl4 = hi & 0x30u;                                          // 16 * log2(#outcomes) = prt byte offset
        const double pr = lds_f64_r(sc.prt + k * 48 + nl4);
        E = __dadd_rn(E, pr);
        if (nl4) {                                                                // rare: 2 or 4 outcomes
            uint32_t slot = E <= u ? 1u : 0u;
            E = __dadd_rn(E, pr);
            if (nl4 == 0x20u) {
                slot += E <= u ? 1u : 0u; E = __dadd_rn(E, pr);
                slot += E <= u ? 1u : 0u; E = __dadd_rn(E, pr);
                ent += slot * 2u;                                                 // 4-way: draw value r = slot
            } else {
                ent += slot * 4u;                                                 // 2-way: r = 2 * slot
            }
        }
        pick = le ? ent : pick;
        le = E <= u;
    }
    if (le) {                                                                     // all-False -> index 0 of the list
        const uint32_t ca = (uint32_t)combo_a((int)sc.first_k), cb = (uint32_t)combo_b((int)sc.first_k);
        pick = (ca == 0 ? ra[0] : (ca == 1 ? ra[1] : ra[2])) + (cb == 0 ? ob[0] : (cb == 1 ? ob[1] : ob[2]));
    }
    const int32_t e = lds_s16_r(pick);
    return table_finish(c, s, e, rsel4);
}

// the injected draw of a slip step as the reference's u in [0, 1): a raw fp64 value, or a 32-bit integer r
// standing for (r + 0.5) / 2^32 (as in k_step_generic)
__device__ __forceinline__ double u_from_rng32(uint32_t r) { return ((double)r + 0.5) * (1.0 / 4294967296.0); }

__device__ __forceinline__ TblCtx make_ctx(const uint8_t* smem, uint32_t table_bytes, const PitchDev& P)
{
    TblCtx c;
    c.tbl = smem_u32(smem);
    c.isd = c.tbl + table_bytes;
    c.last = P.tlast;
    c.maxobs = P.nSm1;
    return c;
}

// ---- episode statistics fused into K1 (optional `stats` argument of the step entry points): what K2 accumulates,
// for ONE lock-step step -- stats[0] += episodes ended, [1] += goals_A, [2] += goals_B (by player A's reward sign,
// whatever the return agent), [3] += truncations without a goal, [4] += env-steps, [5] += summed length of the
// episodes that ended.  ~2.5 instructions per env: byte-parallel flag counts (dp4a), sum of the rewards, and
// sum(len) from sum(t_in + 1) - sum(t_out) (an env that continues contributes 0, one that ended t_in + 1).
struct K1Stats {
    // dt: episodes that ended with a goal (low 16 bits) | truncated-only (high 16 bits): a thread steps < 2^14 groups
    // for any batch that fits the GPU's memory.  `steps` is only used by the one-env-per-thread kernels (the 4-env
    // kernels step every env of the batch: the flush adds n).
    uint32_t dt, len, steps; int32_t net;
    // s_in_sum / s_out_sum: the sums of the four INDEX-layout state words (obs | t << 16; four observations cannot
    // carry into bit 16), or of the four timesteps << 16
    __device__ __forceinline__ void add4(uint32_t fw, int32_t rew_sum, uint32_t s_in_sum, uint32_t s_out_sum)
    {
        dt = __dp4a(fw & 0x01010101u, 0x01010101u, dt);
        dt += __dp4a((fw >> 1) & ~fw & 0x01010101u, 0x01010101u, 0u) << 16;
        net += rew_sum; len += (s_in_sum >> 16) + 4u - (s_out_sum >> 16);
    }
    __device__ __forceinline__ void add1(uint32_t f, int32_t rew, uint32_t ep_len)
    {
        dt += (f & 1u) | (((f >> 1) & ~f & 1u) << 16); net += rew; steps += 1u; len += (f & 3u) ? ep_len : 0u;
    }
};
struct K1StatsBlk { unsigned long long v[4]; long long net; };
__device__ __forceinline__ void k1_stats_init(K1StatsBlk* blk)
{
    if (threadIdx.x < 4) blk->v[threadIdx.x] = 0;
    if (threadIdx.x == 4) blk->net = 0;
}
// every thread of the CTA calls this once at the end of the kernel (contains a CTA barrier); all_steps: env-steps the
// whole launch made when the per-thread `steps` counters are not used (added once, by CTA 0)
__device__ __forceinline__ void k1_stats_flush(const K1Stats& a, K1StatsBlk* blk, unsigned long long* stats,
                                               unsigned long long all_steps = 0)
{
    const uint32_t v[4] = { a.dt & 0xFFFFu, a.dt >> 16, a.steps, a.len };
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t r = __reduce_add_sync(0xFFFFFFFFu, v[j]);
        if ((threadIdx.x & 31) == 0 && r) atomicAdd(&blk->v[j], (unsigned long long)r);
    }
    const int32_t rn = __reduce_add_sync(0xFFFFFFFFu, a.net);
    if ((threadIdx.x & 31) == 0 && rn) atomicAdd(reinterpret_cast<unsigned long long*>(&blk->net), (unsigned long long)(long long)rn);
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long d = (long long)blk->v[0], tr = (long long)blk->v[1], net = blk->net;
        const unsigned long long steps = blk->v[2] + (blockIdx.x == 0 ? all_steps : 0ull);
        if (d + tr) atomicAdd(&stats[0], (unsigned long long)(d + tr));
        if (d + net) atomicAdd(&stats[1], (unsigned long long)((d + net) / 2));
        if (d - net) atomicAdd(&stats[2], (unsigned long long)((d - net) / 2));
        if (tr) atomicAdd(&stats[3], (unsigned long long)tr);
        if (steps) atomicAdd(&stats[4], steps);
        if (blk->v[3]) atomicAdd(&stats[5], blk->v[3]);
    }
}

// Folded table policies of the single-agent modes (SIM:187-188: the folded player's action is its int8[nS] policy at
// the CURRENT observation; SIM:243-244: the reward is flipped when the return agent is player_b, i.e. when player A
// is the folded one).  Shared-window addresses, 0 = that player is not folded.
struct K1Policy { uint32_t pol_a, pol_b; };

// NARROW (soccer_step_narrow): obs as uint16, reward as int8 -- `obs` then points at uint16[n] (one uint2 per group),
// `rew` at int8[n] (one word per group); same values, 8 instead of 13 bytes written per env
template <bool RESET_OBS, bool NARROW = false, bool POLICY = false, bool STATS = false>
__device__ __forceinline__ void table_step_group(const TblCtx& c, const Group4& x, int64_t g, uint4* st, uint4* obs,
                                                 uint4* rew, uint32_t* flg, uint4* rob, const K1Policy& pol, K1Stats& acc)
{
    // byte-parallel: jr = aa*20 + ab*4 + (rng & 3) for all four envs in three instructions.  Action bytes are masked
    // to 3 bits so that an out-of-range byte (>= 13 would carry into the neighbour's column) stays inside its own env;
    // the index clamp in table_step() then keeps it inside the table.
    const uint32_t jr4 = (x.a & 0x07070707u) * 20u + (x.b & 0x07070707u) * 4u + (x.r & 0x03030303u);
    const uint32_t rs4 = x.r & 0x0C0C0C0Cu;
    const uint32_t sv[4] = { x.s.x, x.s.y, x.s.z, x.s.w };
    const bool flip = POLICY && pol.pol_a != 0u;
    uint32_t so[4], oo[4], ro[4], rr[4], ff[4];
    int32_t rsum = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        uint32_t jr = __byte_perm(jr4, 0, 0x4440 + e);
        if (POLICY) {
            const uint32_t cur = min(sv[e] & 0xFFFFu, c.maxobs);
            uint32_t aa = __byte_perm(x.a, 0, 0x4440 + e) & 7u, ab = __byte_perm(x.b, 0, 0x4440 + e) & 7u;
            if (pol.pol_a) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(aa) : "r"(pol.pol_a + cur));
            if (pol.pol_b) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(ab) : "r"(pol.pol_b + cur));
            jr = aa * 20u + ab * 4u + (__byte_perm(x.r, 0, 0x4440 + e) & 3u);
        }
        const TblOut o = table_step(c, sv[e], jr, __byte_perm(rs4, 0, 0x4440 + e));
        so[e] = o.state; oo[e] = o.obs; ro[e] = o.reset_obs; ff[e] = o.flags;
        const int32_t rv = flip ? -o.rew_i : o.rew_i;
        rr[e] = NARROW ? (uint32_t)rv : __float_as_uint((float)rv);
        rsum += o.rew_i;
    }
    const uint32_t fw = __byte_perm(__byte_perm(ff[0], ff[1], 0x0040), __byte_perm(ff[2], ff[3], 0x0040), 0x5410);
    if (STATS)
        acc.add4(fw, rsum, sv[0] + sv[1] + sv[2] + sv[3], so[0] + so[1] + so[2] + so[3]);
    st_keep(st + g, make_uint4(so[0], so[1], so[2], so[3]));
    if (NARROW) {
        st_stream(reinterpret_cast<uint2*>(obs) + g, make_uint2(oo[0] | (oo[1] << 16), oo[2] | (oo[3] << 16)));
        st_stream(reinterpret_cast<uint32_t*>(rew) + g,
                  __byte_perm(__byte_perm(rr[0], rr[1], 0x0040), __byte_perm(rr[2], rr[3], 0x0040), 0x5410));
    } else {
        st_stream(obs + g, make_uint4(oo[0], oo[1], oo[2], oo[3]));
        st_stream(rew + g, make_uint4(rr[0], rr[1], rr[2], rr[3]));
    }
    st_stream(flg + g, fw);
    if (RESET_OBS) st_stream(rob + g, make_uint4(ro[0], ro[1], ro[2], ro[3]));
}

// int8[nS] table policies global -> shared (after the grid-dependency wait: a policy tensor may come from the
// preceding kernel), clamped to 0..4; returns the shared-window addresses
__device__ __forceinline__ K1Policy stage_k1_policies(uint8_t* dst, const int8_t* policy_a, const int8_t* policy_b, int nS)
{
    const uint32_t pol_bytes = ((uint32_t)nS + 15u) & ~15u;
    uint8_t* pa = dst, *pb = dst + pol_bytes;
    for (int i = threadIdx.x; i < nS; i += blockDim.x) {
        if (policy_a) pa[i] = (uint8_t)min(max((int)policy_a[i], 0), 4);
        if (policy_b) pb[i] = (uint8_t)min(max((int)policy_b[i], 0), 4);
    }
    __syncthreads();
    K1Policy k = { policy_a ? smem_u32(pa) : 0u, policy_b ? smem_u32(pb) : 0u };
    return k;
}

// K1, table variant: persistent, one 1024-thread CTA per SM.  PHILOX (soccer_step_table_philox): the draw stream is
// not read (19 B / env-step); the draws of a group -- ONE Philox call, contract v2 -- are computed while the next
// pair's loads are in flight.  POLICY: a folded player's action stream is not read either (its pointer aliases the
// other player's); shared-memory image [table][isd 16 B][policy a][policy b].
struct K1Extra { const int8_t* policy_a; const int8_t* policy_b; unsigned long long* stats; };
// Variants that compute Philox words AND carry policies or statistics need more than the 64 registers a 1024-thread
// CTA leaves per thread (ptxas spilled 40-150 bytes into the loop: 337 -> 251 G env-steps/s); they run 768 threads.
#ifndef SOCCER_K1_HEAVY_THREADS
#define SOCCER_K1_HEAVY_THREADS 768
#endif
template <bool PHILOX, bool POLICY, bool STATS>
constexpr int k1_threads() { return (PHILOX && (POLICY || STATS)) ? SOCCER_K1_HEAVY_THREADS : kTableThreads; }
template <bool RESET_OBS, bool PHILOX = false, bool NARROW = false, bool POLICY = false, bool STATS = false>
__global__ void __launch_bounds__((k1_threads<PHILOX, POLICY, STATS>()), 1)
k_step_table(const PitchDev P, const uint16_t* __restrict__ gtable, uint32_t table_bytes,
             uint32_t* __restrict__ state, const uint8_t* __restrict__ act_a, const uint8_t* __restrict__ act_b,
             const uint8_t* __restrict__ rng, int32_t* __restrict__ obs, float* __restrict__ reward,
             uint8_t* __restrict__ flags, int32_t* __restrict__ reset_obs, int64_t n_groups,
             const PhiloxKey key, const K1Extra ex)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ K1StatsBlk sblk;
    pdl_launch_dependents();
    if (STATS) k1_stats_init(&sblk);
    stage_table(smem_raw, gtable, table_bytes, &bar, P);     // the table is constant: safe before pdl_wait
    const TblCtx c = make_ctx(smem_raw, table_bytes, P);
    pdl_wait();                  // the table fill overlaps the previous kernel's tail
    K1Policy pol = { 0u, 0u };
    if (POLICY) pol = stage_k1_policies(smem_raw + table_bytes + 16, ex.policy_a, ex.policy_b, P.nS);
    K1Stats acc = {};

    uint4* st4 = reinterpret_cast<uint4*>(state);
    // a folded player's action stream does not exist: alias the other one (loaded, never used)
    const uint32_t* a4 = reinterpret_cast<const uint32_t*>(act_a ? act_a : act_b);
    const uint32_t* b4 = reinterpret_cast<const uint32_t*>(act_b ? act_b : act_a);
    const uint32_t* r4 = PHILOX ? a4 : reinterpret_cast<const uint32_t*>(rng);   // PHILOX: value replaced below
    uint4* o4 = reinterpret_cast<uint4*>(obs);
    uint4* w4 = reinterpret_cast<uint4*>(reward);
    uint32_t* f4 = reinterpret_cast<uint32_t*>(flags);
    uint4* q4 = reinterpret_cast<uint4*>(reset_obs);

    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // first pair of groups: issue the HBM loads, THEN wait for the table
    bool one = g < n_groups, two = g + stride < n_groups;
    Group4 x0 = {}, x1 = {};
    if (one) x0 = load_group(st4, a4, b4, r4, g);
    if (two) x1 = load_group(st4, a4, b4, r4, g + stride);
    wait_table(&bar);
    while (one) {
        const int64_t gn = g + 2 * stride;
        const bool n_one = gn < n_groups, n_two = gn + stride < n_groups;
        Group4 y0 = x0, y1 = x1;
        if (n_one) y0 = load_group(st4, a4, b4, r4, gn);          // prefetch the next pair
        if (n_two) y1 = load_group(st4, a4, b4, r4, gn + stride);
        if (SOCCER_K1_L2_PREFETCH) {
            const int64_t gp = gn + (int64_t)SOCCER_K1_L2_PREFETCH * 2 * stride;
            if (gp < n_groups) prefetch_group(st4, a4, b4, r4, gp, !PHILOX);
            if (gp + stride < n_groups) prefetch_group(st4, a4, b4, r4, gp + stride, !PHILOX);
        }
        if (PHILOX) { x0.r = philox_rng8x4(key, g); if (two) x1.r = philox_rng8x4(key, g + stride); }
        table_step_group<RESET_OBS, NARROW, POLICY, STATS>(c, x0, g, st4, o4, w4, f4, q4, pol, acc);
        if (two) table_step_group<RESET_OBS, NARROW, POLICY, STATS>(c, x1, g + stride, st4, o4, w4, f4, q4, pol, acc);
        x0 = y0; x1 = y1; g = gn; one = n_one; two = n_two;
    }
    if (STATS) k1_stats_flush(acc, &sblk, ex.stats, (unsigned long long)n_groups * 4ull);
}

// scalar tail / misaligned fallback of the table path (global-memory table, one env per thread)
__global__ void __launch_bounds__(kThreads)
k_step_table_scalar(const PitchDev P, const uint16_t* __restrict__ gtable, uint32_t* __restrict__ state,
                    const uint8_t* __restrict__ act_a, const uint8_t* __restrict__ act_b,
                    const uint8_t* __restrict__ rng, int32_t* __restrict__ obs, float* __restrict__ reward,
                    uint8_t* __restrict__ flags, int32_t* __restrict__ reset_obs, int64_t n,
                    int use_philox, const PhiloxKey key, int narrow, const K1Extra ex)
{
    __shared__ K1StatsBlk sblk;
    k1_stats_init(&sblk);
    __syncthreads();
    K1Stats acc = {};
    const int16_t* tbl = reinterpret_cast<const int16_t*>(gtable);
    const uint32_t last = (uint32_t)P.nS * 100u - 1u;
    const bool flip = ex.policy_a != nullptr;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t s = state[i];
        const uint32_t rg = use_philox ? philox_rng8(philox_word(key.seed, key.env_id_base + (uint64_t)i, key.step)) : rng[i];
        const uint32_t cur = min(s & 0xFFFFu, (uint32_t)P.nS - 1u);
        const uint32_t aa = ex.policy_a ? (uint32_t)min(max((int)ex.policy_a[cur], 0), 4) : ((uint32_t)act_a[i] & 7u);
        const uint32_t ab = ex.policy_b ? (uint32_t)min(max((int)ex.policy_b[cur], 0), 4) : ((uint32_t)act_b[i] & 7u);
        const uint32_t idx = min((s & 0xFFFFu) * 100u + aa * 20u + ab * 4u + (rg & 3u), last);
        const int32_t e = tbl[idx];
        const uint32_t nobs = (uint32_t)e & kTblObsMask;
        const bool done = nobs == 0;
        const uint32_t s1 = s + 0x10000u;
        const bool trunc = s1 >= kTruncWord, reset = done | trunc;
        const uint32_t ro = (uint32_t)P.isd_obs[(rg >> 2) & 3u];
        state[i] = reset ? ro : ((s1 & 0xFFFF0000u) | nobs);
        const int32_t rv = flip ? -(e >> 14) : (e >> 14);
        if (narrow) {
            reinterpret_cast<uint16_t*>(obs)[i] = (uint16_t)nobs;
            reinterpret_cast<int8_t*>(reward)[i] = (int8_t)rv;
        } else {
            obs[i] = (int32_t)nobs;
            reward[i] = (float)rv;
        }
        const uint32_t f = (done ? 1u : 0u) + (trunc ? 2u : 0u);
        flags[i] = (uint8_t)f;
        if (reset_obs) reset_obs[i] = (int32_t)(reset ? ro : nobs);
        acc.add1(f, e >> 14, s1 >> 16);
    }
    if (ex.stats) k1_stats_flush(acc, &sblk, ex.stats);
}

// K1, table variant with PACKED host-facing streams (soccer_step_table_packed): the joint action of an env in one
// byte (aa | ab << 4), the draws in one byte, and ONE 16-bit result word per env
//     result = next_obs | terminated << 12 | truncated << 13 | (reward & 3) << 14
// (as int16: reward = w >> 14, obs = w & 0xFFF) -- 2 bytes in and 2 bytes out per env-step instead of 3 + 9, for
// the PCIe-bound host path; 12 algorithmic bytes per env-step on the device.  The pointers may be device memory
// or pinned host memory (zero copy).
__device__ __forceinline__ uint32_t packed_result(int32_t e, uint32_t flags)
{
    return ((uint32_t)e & 0xCFFFu) | (flags << 12);
}
struct GroupP { uint4 s; uint32_t j, r; };
template <bool PHILOX>
__device__ __forceinline__ GroupP load_group_packed(const uint4* st, const uint32_t* j4, const uint32_t* r4, int64_t g)
{
    GroupP x;
    x.s = ld_keep(st + g);
    x.j = ld_stream(j4 + g);
    x.r = PHILOX ? 0u : ld_stream(r4 + g);
    return x;
}
__device__ __forceinline__ void table_step_group_packed(const TblCtx& c, const GroupP& x, int64_t g, uint4* st, uint2* res)
{
    // nibbles masked to 3 bits: an out-of-range action stays inside its own env (see table_step_group)
    const uint32_t jr4 = (x.j & 0x07070707u) * 20u + ((x.j >> 4) & 0x07070707u) * 4u + (x.r & 0x03030303u);
    const uint32_t rs4 = x.r & 0x0C0C0C0Cu;
    const uint32_t sv[4] = { x.s.x, x.s.y, x.s.z, x.s.w };
    uint32_t so[4], w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const TblOut o = table_step(c, sv[e], __byte_perm(jr4, 0, 0x4440 + e), __byte_perm(rs4, 0, 0x4440 + e));
        so[e] = o.state;
        w[e] = o.obs | (o.flags << 12) | (((uint32_t)o.rew_i & 3u) << 14);
    }
    st_keep(st + g, make_uint4(so[0], so[1], so[2], so[3]));
    st_stream(res + g, make_uint2(w[0] | (w[1] << 16), w[2] | (w[3] << 16)));
}

// PHILOX: no draw stream (1 byte in, 2 bytes out per env-step): the draws come from the env's Philox word
template <bool PHILOX>
__global__ void __launch_bounds__(kTableThreads, 1)
k_step_table_packed(const PitchDev P, const uint16_t* __restrict__ gtable, uint32_t table_bytes,
                    uint32_t* __restrict__ state, const uint8_t* __restrict__ joint, const uint8_t* __restrict__ rng,
                    uint16_t* __restrict__ result, int64_t n_groups, const PhiloxKey key)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    pdl_launch_dependents();
    stage_table(smem_raw, gtable, table_bytes, &bar, P);
    const TblCtx c = make_ctx(smem_raw, table_bytes, P);
    pdl_wait();

    uint4* st4 = reinterpret_cast<uint4*>(state);
    const uint32_t* j4 = reinterpret_cast<const uint32_t*>(joint);
    const uint32_t* r4 = reinterpret_cast<const uint32_t*>(rng);
    uint2* w2 = reinterpret_cast<uint2*>(result);

    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool one = g < n_groups, two = g + stride < n_groups;
    GroupP x0 = {}, x1 = {};
    if (one) x0 = load_group_packed<PHILOX>(st4, j4, r4, g);
    if (two) x1 = load_group_packed<PHILOX>(st4, j4, r4, g + stride);
    wait_table(&bar);
    while (one) {
        const int64_t gn = g + 2 * stride;
        const bool n_one = gn < n_groups, n_two = gn + stride < n_groups;
        GroupP y0 = x0, y1 = x1;
        if (n_one) y0 = load_group_packed<PHILOX>(st4, j4, r4, gn);
        if (n_two) y1 = load_group_packed<PHILOX>(st4, j4, r4, gn + stride);
        if (PHILOX) { x0.r = philox_rng8x4(key, g); if (two) x1.r = philox_rng8x4(key, g + stride); }
        table_step_group_packed(c, x0, g, st4, w2);
        if (two) table_step_group_packed(c, x1, g + stride, st4, w2);
        x0 = y0; x1 = y1; g = gn; one = n_one; two = n_two;
    }
}

// scalar tail / misaligned fallback of the packed table step (global-memory table, one env per thread)
__global__ void __launch_bounds__(kThreads)
k_step_table_packed_scalar(const PitchDev P, const uint16_t* __restrict__ gtable, uint32_t* __restrict__ state,
                           const uint8_t* __restrict__ joint, const uint8_t* __restrict__ rng,
                           uint16_t* __restrict__ result, int64_t n, int use_philox, const PhiloxKey key)
{
    const int16_t* tbl = reinterpret_cast<const int16_t*>(gtable);
    const uint32_t last = (uint32_t)P.nS * 100u - 1u;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t s = state[i], j = joint[i];
        const uint32_t rg = use_philox ? philox_rng8(philox_word(key.seed, key.env_id_base + (uint64_t)i, key.step)) : rng[i];
        const uint32_t idx = min((s & 0xFFFFu) * 100u + (j & 7u) * 20u + ((j >> 4) & 7u) * 4u + (rg & 3u), last);
        const int32_t e = tbl[idx];
        const uint32_t nobs = (uint32_t)e & kTblObsMask;
        const bool done = nobs == 0;
        const uint32_t s1 = s + 0x10000u;
        const bool trunc = s1 >= kTruncWord, reset = done | trunc;
        const uint32_t ro = (uint32_t)P.isd_obs[(rg >> 2) & 3u];
        state[i] = reset ? ro : ((s1 & 0xFFFF0000u) | nobs);
        result[i] = (uint16_t)packed_result(e, (done ? 1u : 0u) + (trunc ? 2u : 0u));
    }
}

// K1, table variant for slip_prob > 0: same launch shape as k_step_table, one more input stream (the step draw:
// uint32 or fp64 per env; rng8 still carries the reset draw in bits 2..3).  24 or 28 algorithmic bytes per env-step.
// DRAW: where the step draw comes from -- kDrawU32 (rng32 stream, u = (r + 0.5) / 2^32), kDrawF64 (raw fp64 stream),
// kDrawPhilox (no draw / rng8 streams: r32 = lo32(25 w) and the reset draw from the env's Philox word; 19 B).
#ifndef SOCCER_SLIP_THREADS
#define SOCCER_SLIP_THREADS 1024     // latency-bound (dependent DADD chain): 1024 threads x 64 registers beat 512 x 83 (98 vs 83 G env-steps/s)
#endif
constexpr int kSlipThreads = SOCCER_SLIP_THREADS;       // the slip walk keeps 4 envs x (sum, candidate, 6 addresses) live
constexpr int kDrawU32 = 0, kDrawF64 = 1, kDrawPhilox = 2;

// the 4 envs' step draws + rng8 bytes of one group
struct SlipIn { Group4 x; uint4 d0, d1; };
template <int DRAW, bool PLAIN>
__device__ __forceinline__ SlipIn slip_load_in(const uint4* st4, const uint32_t* a4, const uint32_t* b4, const uint32_t* r4,
                                               const void* draw, int64_t g)
{
    SlipIn in;
    if (PLAIN) {
        // evict-normal loads: the queued walk re-reads the deferred envs' inputs a few microseconds later
        in.x.s = ld_keep(st4 + g); in.x.a = a4[g]; in.x.b = b4[g];
        in.x.r = DRAW == kDrawPhilox ? 0u : r4[g];
        if (DRAW == kDrawF64) {
            in.d0 = reinterpret_cast<const uint4*>(draw)[2 * g];
            in.d1 = reinterpret_cast<const uint4*>(draw)[2 * g + 1];
        } else if (DRAW == kDrawU32) {
            in.d0 = reinterpret_cast<const uint4*>(draw)[g];
            in.d1 = in.d0;
        } else {
            in.d0 = in.d1 = make_uint4(0, 0, 0, 0);
        }
    } else {
        in.x.s = ld_keep(st4 + g); in.x.a = ld_stream(a4 + g); in.x.b = ld_stream(b4 + g);
        in.x.r = DRAW == kDrawPhilox ? 0u : ld_stream(r4 + g);
        if (DRAW == kDrawF64) {
            in.d0 = __ldcs(reinterpret_cast<const uint4*>(draw) + 2 * g);
            in.d1 = __ldcs(reinterpret_cast<const uint4*>(draw) + 2 * g + 1);
        } else if (DRAW == kDrawU32) {
            in.d0 = __ldcs(reinterpret_cast<const uint4*>(draw) + g);
            in.d1 = in.d0;
        } else {
            in.d0 = in.d1 = make_uint4(0, 0, 0, 0);
        }
    }
    return in;
}
// Philox mode: fill the draws of group g (one call, contract v2)
__device__ __forceinline__ void slip_philox_fill(SlipIn& in, const PhiloxKey& key, int64_t g)
{
    uint32_t w[4];
    philox_words4(key, g, w);
    in.d0 = make_uint4(philox_r32(w[0]), philox_r32(w[1]), philox_r32(w[2]), philox_r32(w[3]));
    in.d1 = in.d0;
    in.x.r = (w[0] & 0xCu) | ((w[1] & 0xCu) << 8) | ((w[2] & 0xCu) << 16) | ((w[3] & 0xCu) << 24);
}

struct SlipExtra { const int8_t* policy_a; const int8_t* policy_b; PhiloxKey key; };

template <bool RESET_OBS, int DRAW>
__global__ void __launch_bounds__(kSlipThreads, 1)
k_step_table_slip(const PitchDev P, const uint16_t* __restrict__ gtable, uint32_t table_bytes,
                  uint32_t* __restrict__ state, const uint8_t* __restrict__ act_a, const uint8_t* __restrict__ act_b,
                  const uint8_t* __restrict__ rng, const void* __restrict__ draw, int32_t* __restrict__ obs,
                  float* __restrict__ reward, uint8_t* __restrict__ flags, int32_t* __restrict__ reset_obs, int64_t n_groups,
                  const SlipExtra ex)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];     // [table][isd 16 B][policy a][policy b]
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(16) double prt[kPrtDoubles];
    slip_build_prt(prt, P);
    stage_table(smem_raw, gtable, table_bytes, &bar, P);     // ends with __syncthreads(): prt visible
    TblCtx c = make_ctx(smem_raw, table_bytes, P);
    SlipCtx sc = { smem_u32(prt), slip_first_k(P) };
    K1Policy pol = { 0u, 0u };
    if (ex.policy_a || ex.policy_b) pol = stage_k1_policies(smem_raw + table_bytes + 16, ex.policy_a, ex.policy_b, P.nS);
    const bool flip = pol.pol_a != 0u;
    uint4* st4 = reinterpret_cast<uint4*>(state);
    const uint32_t* a4 = reinterpret_cast<const uint32_t*>(act_a ? act_a : act_b);
    const uint32_t* b4 = reinterpret_cast<const uint32_t*>(act_b ? act_b : act_a);
    const uint32_t* r4 = reinterpret_cast<const uint32_t*>(rng);
    uint4* o4 = reinterpret_cast<uint4*>(obs);
    uint4* w4 = reinterpret_cast<uint4*>(reward);
    uint32_t* f4 = reinterpret_cast<uint32_t*>(flags);
    uint4* q4 = reinterpret_cast<uint4*>(reset_obs);
    wait_table(&bar);
    launder(c.tbl); launder(c.isd); launder(sc.prt);        // the relaxed loads below stay below the wait
    launder(pol.pol_a); launder(pol.pol_b);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        SlipIn in = slip_load_in<DRAW, false>(st4, a4, b4, r4, draw, g);
        if (DRAW == kDrawPhilox) slip_philox_fill(in, ex.key, g);
        const Group4 x = in.x;
        double u[4];
        if (DRAW == kDrawF64) {
            u[0] = __hiloint2double((int)in.d0.y, (int)in.d0.x); u[1] = __hiloint2double((int)in.d0.w, (int)in.d0.z);
            u[2] = __hiloint2double((int)in.d1.y, (int)in.d1.x); u[3] = __hiloint2double((int)in.d1.w, (int)in.d1.z);
        } else {
            u[0] = u_from_rng32(in.d0.x); u[1] = u_from_rng32(in.d0.y); u[2] = u_from_rng32(in.d0.z); u[3] = u_from_rng32(in.d0.w);
        }
        const uint32_t rs4 = x.r & 0x0C0C0C0Cu;
        const uint32_t sv[4] = { x.s.x, x.s.y, x.s.z, x.s.w };
        uint32_t so[4], oo[4], ro[4], rr[4], ff[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            uint32_t aa = __byte_perm(x.a, 0, 0x4440 + e), ab = __byte_perm(x.b, 0, 0x4440 + e);
            const uint32_t cur = min(sv[e] & 0xFFFFu, c.maxobs);
            if (pol.pol_a) aa = lds_u8_r(pol.pol_a + cur);
            if (pol.pol_b) ab = lds_u8_r(pol.pol_b + cur);
            const TblOut o = table_step_slip(c, sc, sv[e], aa, ab, u[e], __byte_perm(rs4, 0, 0x4440 + e));
            so[e] = o.state; oo[e] = o.obs; rr[e] = __float_as_uint((float)(flip ? -o.rew_i : o.rew_i));
            ro[e] = o.reset_obs; ff[e] = o.flags;
        }
        st_keep(st4 + g, make_uint4(so[0], so[1], so[2], so[3]));
        st_stream(o4 + g, make_uint4(oo[0], oo[1], oo[2], oo[3]));
        st_stream(w4 + g, make_uint4(rr[0], rr[1], rr[2], rr[3]));
        st_stream(f4 + g, __byte_perm(__byte_perm(ff[0], ff[1], 0x0040), __byte_perm(ff[2], ff[3], 0x0040), 0x5410));
        if (RESET_OBS) st_stream(q4 + g, make_uint4(ro[0], ro[1], ro[2], ro[3]));
    }
}

// ---- K1 slip, fast path + deferred queue (soccer_step_table_slip with a slip index).
// 97 % of all (state, move pair)s have ONE outcome, and the walk ends at the first combination in 64 % of the draws
// (slip_prob 0.2), so for most envs the reference's cumulative sums are the CONSTANT prefix sums E_k of the nine
// combination probabilities mp_k (accumulated like the reference: sequential __dadd_rn), and the pick is
// k = #{E_j <= u} -- exact as long as combination k has one outcome and the true end-of-combination sums up to k
// equal the constant ones bit for bit.  The slip index -- one byte per (obs, aa*5+ab): bit k set <=> pick k must be
// walked, built once on the device from the step table and the env's slip_prob (k_build_slip_index) -- decides that
// with one byte load.  Envs whose pick has its bit set (or is 8 or 9) -- a few percent -- are not walked
// in place (one such lane would hold its whole warp in the 150-instruction walk): their ids go to the warp's
// shared-memory queue (ballot + popc), and after the fast pass over its 256 envs the warp takes the queued envs
// through table_step_slip 32 at a time.  Results are bit-identical to k_step_table_slip (and to the reference): same sums, same compares.
// first 32-bit draw r whose u = (r + 0.5) / 2^32 satisfies E <= u, i.e. ceil(E * 2^32 - 0.5) (exact in fp64: a power-of-two
// scaling and a subtraction of 0.5 below 2^52), clamped to [0, 2^32]: 2^32 = no draw reaches it
__host__ __device__ inline uint64_t slip_thr(double E)
{
    const double x = ceil(E * 4294967296.0 - 0.5);
    return x <= 0.0 ? 0ull : (x >= 4294967296.0 ? 4294967296ull : (uint64_t)x);
}

// The slip index has TWO planes, one entry per (obs, aa * 5 + ab); plane_bytes = nS * 25 rounded up to 16:
//   plane 0, bytes [0, plane_bytes), one BYTE per entry (fp64 draws, k_step_table_slip_q): bit k <=> a draw whose
//     constant-prefix pick is combination k must be walked: combination k itself has 2 or 4 outcomes (the pick falls on a
//     slot inside it), or some end-of-combination sum up to k differs from the constant E_j in ANY bit (a 2- or
//     4-outcome combination adds mp/2 twice or mp/4 four times instead of mp once; the roundings almost always agree,
//     this checks it).  Picks 8 and 9 are always walked.
//   plane 1, bytes [plane_bytes, 3 * plane_bytes), one UINT16 per entry (32-bit draws -- injected rng32 and every Philox
//     draw; table_step_slip_int): with u = (r + 0.5) / 2^32 every comparison "running sum <= u" of the reference's walk is
//     "r >= slip_thr(running sum)", an INTEGER threshold, and a last-bit difference between the true running sum and the
//     constant one moves that integer only if an integer lies between them (probability ~2^-21 per threshold).  So the
//     fast path decides combination AND slot from constant integer thresholds, and bit k (k = 0 .. 8) marks the picks
//     for which some threshold of combination k or k - 1 -- its end or a slot inside it -- is NOT the constant one, or
//     is 0 and cannot be written as a strict 32-bit compare.  Bit 9 is always set: pick 9 = "no running sum exceeds u".
//     The kernels do NOT read this plane any more: the integer fast path proves the same thing state-independently
//     (SlipDanger, below); the plane is the constructive cross-check of that argument -- whenever the danger list of a
//     slip_prob is empty, bits 0 .. 8 must be clear for every (obs, joint action) (tests/test_gpu_round2.py).
__global__ void __launch_bounds__(kThreads)
k_build_slip_index(const PitchDev P, int32_t nS, const uint16_t* __restrict__ table, uint8_t* __restrict__ fc, int64_t plane_bytes)
{
    uint16_t* fc16 = reinterpret_cast<uint16_t*>(fc + plane_bytes);
    const int64_t total = (int64_t)nS * 25;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t obs = (uint32_t)(i / 25), ja = (uint32_t)(i % 25), aa = ja / 5u, ab = ja % 5u;
        uint32_t mask = 0, imask = 1u << 9;
        double ec = 0.0, et = 0.0;
        bool off = false;                           // the true sums have left the constant ones
        for (int k = 0; k < 9; ++k) {
            const int ca = combo_a(k), cb = combo_b(k);
            const uint32_t ma = ca == 0 ? aa : slip_move(aa, ca - 1), mb = cb == 0 ? ab : slip_move(ab, cb - 1);
            const uint32_t nl = (table[obs * 100u + (ma * 5u + mb) * 4u] >> 12) & 3u;
            const double mp = P.mp[k];
            const double ec_prev = ec;
            ec = __dadd_rn(ec, __dmul_rn(mp, 1.0));
            const double pr = __dmul_rn(mp, nl == 2u ? 0.25 : (nl == 1u ? 0.5 : 1.0));       // SIM:241
            bool ibad = false;
            double cc = ec_prev;                    // what the fast path adds up inside combination k: constant prefix + slots
            for (uint32_t j = 0; j < (1u << nl); ++j) {                                      // the reference's running sum
                et = __dadd_rn(et, pr);
                cc = __dadd_rn(cc, pr);
                // slot thresholds inside the combination come from (constant prefix + j * pr), its end from E_k
                const bool last = j + 1u == (1u << nl);
                const uint64_t t_fast = slip_thr(last ? ec : cc);
                ibad |= slip_thr(et) != t_fast;
                // inner thresholds are used as strict 32-bit compares r > t - 1 (2^32 -> 0xFFFFFFFF = never: correct);
                // t == 0 ("always") cannot be written that way
                if (!last && mp != 0.0) ibad |= t_fast == 0ull;
            }
            off |= __double_as_longlong(et) != __double_as_longlong(ec);
            if (k < 8 && ((nl != 0u && mp != 0.0) || off)) mask |= 1u << k;
            if (ibad) imask |= (1u << k) | (1u << min(k + 1, 8));
        }
        fc[i] = (uint8_t)mask;
        fc16[i] = (uint16_t)imask;
    }
}

#ifndef SOCCER_SLIP_PLAIN_LOADS
#define SOCCER_SLIP_PLAIN_LOADS 1
#endif
#ifndef SOCCER_SLIP_PLAIN_STORES
#define SOCCER_SLIP_PLAIN_STORES 1
#endif
#if SOCCER_SLIP_PLAIN_STORES
#define SOCCER_SLIP_ST(p, v) (*(p) = (v))
#else
#define SOCCER_SLIP_ST(p, v) st_stream((p), (v))
#endif
#ifndef SOCCER_SLIP_CARRY
#define SOCCER_SLIP_CARRY 1
#endif
constexpr int kSlipQueueWarp = 288;                         // per warp: all 2 x 32 x 4 envs of an iteration + < 32 carried over
constexpr int kSlipQueueBytes = kSlipQueueWarp * (kTableThreads / 32);   // one byte per entry: every KB of shared memory saved is L1 for the walk's re-reads
struct SlipFast { uint32_t fc, cacb, mv3, klut; };          // shared-window addresses
struct SlipE { double e[9]; };                              // E_k, a kernel parameter: DSETP reads it from the constant bank
// fast path of one env: returns the finished step and whether it has to be redone by the walk.  The draw is the raw
// fp64 u (F64: nine compares against the constant bank), or the uint32 r standing for (r + 0.5) / 2^32: k(r) is a step
// function of r with nine steps, so a 4096-entry shared-memory table indexed by the top 12 bits of r yields k with
// ONE load for every bucket that contains no step; the <= 9 buckets that do are marked 0xFF and walked.
constexpr int kSlipLutBits = 12;
template <bool F64>
__device__ __forceinline__ TblOut table_step_slip_fast(const TblCtx& c, const SlipFast& f, const SlipE& E, uint32_t s,
                                                       uint32_t aa, uint32_t ab, double u, uint32_t r32, uint32_t rsel4,
                                                       bool& defer)
{
    const uint32_t obsi = min(s & 0xFFFFu, c.maxobs);
    aa = min(aa, 4u); ab = min(ab, 4u);
    const uint32_t dm = lds_u8_r(f.fc + obsi * 25u + aa * 5u + ab);                  // bit k: pick k must be walked
    uint32_t k = 0;
    if (F64) {
#pragma unroll
        for (int j = 0; j < 9; ++j) k += E.e[j] <= u ? 1u : 0u;                   // E_j non-decreasing: k = first E_k > u
    } else {
        k = lds_u8_r(f.klut + (r32 >> (32 - kSlipLutBits)));                         // 0xFF: a step of k(r) inside the bucket
    }
    defer = k >= 8u || ((dm >> k) & 1u) != 0u;                                     // incl. k == 9 (all-False) and 0xFF: walk
    const uint32_t cc = lds_u8_r(f.cacb + min(k, 8u));
    const uint32_t ma = lds_u8_r(f.mv3 + aa * 3u + (cc & 3u)), mb = lds_u8_r(f.mv3 + ab * 3u + (cc >> 4));
    const int32_t e = lds_s16_r(c.tbl + obsi * 200u + (ma * 5u + mb) * 8u);
    return table_finish(c, s, e, rsel4);
}

// the small look-up tables of the fast path, built by every CTA: cacb[k] = move selectors of combination k,
// mv3[a * 3 + sel] = the (slipped) move, klut[top 12 bits of r] = k(r) or 0xFF (uint32 draws only)
template <bool F64>
__device__ __forceinline__ void slip_fast_build_luts(uint8_t* cacb, uint8_t* mv3, uint8_t* klut, const SlipE& E)
{
    if (threadIdx.x < 9) cacb[threadIdx.x] = (uint8_t)(combo_a((int)threadIdx.x) | (combo_b((int)threadIdx.x) << 4));
    if (threadIdx.x < 15) {
        const uint32_t a = threadIdx.x / 3u, cmb = threadIdx.x % 3u;
        mv3[threadIdx.x] = (uint8_t)(cmb == 0 ? a : slip_move(a, (int)cmb - 1));
    }
    if (!F64) {
        // k(r) per bucket of 2^20 consecutive draws: the same count at both ends <=> no step inside (k is monotone in r)
        for (uint32_t b = threadIdx.x; b < (1u << kSlipLutBits); b += blockDim.x) {
            const uint32_t lo = b << (32 - kSlipLutBits), hi = lo | ((1u << (32 - kSlipLutBits)) - 1u);
            const double ulo = u_from_rng32(lo), uhi = u_from_rng32(hi);
            uint32_t klo = 0, khi = 0;
#pragma unroll
            for (int j = 0; j < 9; ++j) { klo += E.e[j] <= ulo ? 1u : 0u; khi += E.e[j] <= uhi ? 1u : 0u; }
            klut[b] = (uint8_t)(klo == khi ? klo : 0xFFu);
        }
    }
}
// stage table + slip index with ONE mbarrier; ends with __syncthreads()
__device__ __forceinline__ void stage_table_and_index(uint8_t* smem, const uint16_t* gtable, uint32_t table_bytes,
                                                      const uint8_t* gfc, uint32_t fc_bytes, uint64_t* bar, const PitchDev& P)
{
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, table_bytes + fc_bytes);
        const uint32_t chunk = 16384;
        for (uint32_t off = 0; off < table_bytes; off += chunk)
            tma_load_1d(smem + off, reinterpret_cast<const uint8_t*>(gtable) + off, min(chunk, table_bytes - off), bar);
        for (uint32_t off = 0; off < fc_bytes; off += chunk)
            tma_load_1d(smem + table_bytes + 16 + off, gfc + off, min(chunk, fc_bytes - off), bar);
    }
    if (threadIdx.x < 4) reinterpret_cast<int32_t*>(smem + table_bytes)[threadIdx.x] = P.isd_obs[threadIdx.x];
    __syncthreads();
}

// ---- 32-bit draws: combination and slot from constant INTEGER thresholds, no queue, no per-state index.
// With u = (r + 0.5) / 2^32 every comparison "running sum <= u" of the reference's walk is "r >= slip_thr(running sum)".
// The TRUE running sum (which depends on the outcome counts of the earlier combinations) and the CONSTANT one the fast
// path uses (E_{k-1} + j * pr) are fp64 sums of the same <= 36 non-negative terms <= 1 in different groupings (the
// products mp * 0.5, mp * 0.25 are exact), so they differ by < 48 * 2^-53 = 5.3e-15, i.e. by < 2.3e-5 in threshold units
// (x 2^32): the two integer thresholds can only differ if sum * 2^32 - 0.5 lies within 2.3e-5 of an integer m -- and then
// only the draw r == m is affected.  The host lists those integers for the env's slip_prob, with a 4x margin (SlipDanger:
// empty for 0.2, 0.5, 0.9, 0.123456789, ...); a draw on the list, the never-reached "no sum exceeds u" pick and a bucket
// of k(r) with two steps take the reference's walk, everything else is decided by integer compares.  State-independent:
// the same code serves the rules kernels of any pitch.  (k_build_slip_index's plane 1 is the constructive check of this
// argument for the table pitches: tests assert it flags nothing the list does not cover.)
// Shared look-up tables, built by every CTA from E_k and the combination probabilities (B = 2^bits buckets, S = 32 - bits):
//   kt[b]   (B uint32)     (klo << S) + (2^S - 1 - low S bits of t - 1), klo = #{j : E_j <= u(low end of bucket b)} and t the
//                          ONE step of k(r) inside the bucket (none: low part 0), so that
//                              k = (kt[b] + (r & (2^S - 1))) >> S          = klo + (r > t - 1)
//                          -- the compare is the carry out of the low S bits: one load, three integer instructions.
//                          klo = 9 = undecided (two or more steps inside the bucket: tiny slip_prob) -> walk
//   sl[k][nl - 1]          (20 rows of uint4) strict thresholds t - 1 of slots 1, 2, 3 inside a 4-way combination k,
//                          {t, t, never} for the one threshold of a 2-way one; unused = 0xFFFFFFFF
//   mva[k][a], mvb[k][a]   (10 x 8 bytes each) byte offsets inside the table row of the move player A / B makes in
//                          combination k with action a: (slipped move) * 40, * 8 (kernels with a folded policy)
//   mvj[k][aa * 5 + ab]    (10 x 32 bytes) both players' moves of combination k in ONE byte, indexed by the Philox joint
//                          action ja = mulhi(w, 25): ma * 40 + mb * 8 = the byte offset of the move pair inside the table
//                          row (table kernels), ma | mb << 4 (rules kernels)
//   mvs[k][aa * 8 + ab]    (10 x 64 bytes) the same indexed by two 3-bit action fields (caller-supplied action bytes,
//                          combined byte-parallel for the four envs of a thread); actions > 4 act as 4
// ONE shared-window base address (one register; every look-up is [base + index + immediate]): the small tables first, the
// bucket table behind them.  shift = 32 - bits, mask = 2^shift - 1.
struct SlipInt {
    uint32_t base, shift, mask;
    __device__ __forceinline__ uint32_t sl() const { return base; }
    __device__ __forceinline__ uint32_t mva() const { return base + 320u; }
    __device__ __forceinline__ uint32_t mvb() const { return base + 400u; }
    __device__ __forceinline__ uint32_t mvj() const { return base + 480u; }
    __device__ __forceinline__ uint32_t mvs() const { return base + 800u; }
    __device__ __forceinline__ uint32_t kt() const { return base + 1440u; }
};
struct SlipDanger { uint32_t n; uint32_t r[12]; };             // draws that must take the walk (kernel parameter)
// bucket shape as a kernel parameter (host-computed, so that shift and mask are constant-bank operands, not registers)
struct SlipBits { int32_t bits; uint32_t shift, mask; };
__host__ __device__ constexpr SlipBits slip_bits(int bits) { return SlipBits{ bits, (uint32_t)(32 - bits), (1u << (32 - bits)) - 1u }; }
__host__ __device__ constexpr int slip_int_lut_bytes(int bits) { return (1 << bits) * 4 + 20 * 16 + 80 + 80 + 320 + 640; }
// scale_a / scale_b: 40 / 8 = byte offsets inside a table row (table kernels), 1 / 16 = the move ids (rules kernels)
// dg: the listed draws do not cost the hot path anything -- their BUCKET is marked undecided (klo = 9), so every draw of
// it takes the walk (<= 12 of 2^bits buckets, and none at all for ordinary slip_prob values).
__device__ __forceinline__ void slip_int_build_luts(uint8_t* base, const SlipE& E, const PitchDev& P, const SlipDanger& dg, int bits,
                                                    uint32_t scale_a = 40u, uint32_t scale_b = 8u)
{
    const uint32_t nb = 1u << bits;
    uint32_t* sl = reinterpret_cast<uint32_t*>(base);
    uint8_t* mva = base + 320;
    uint8_t* mvb = mva + 80; uint8_t* mvj = mvb + 80; uint8_t* mvs = mvj + 320;
    uint32_t* kt = reinterpret_cast<uint32_t*>(base + 1440);
    auto move_of = [](int k, uint32_t a, bool second) {
        const int cmb = second ? combo_b(k) : combo_a(k);
        return cmb == 0 ? a : slip_move(a, cmb - 1);
    };
    if (threadIdx.x < 80) {
        const int k = min((int)threadIdx.x >> 3, 8), a = min((int)threadIdx.x & 7, 4);
        mva[threadIdx.x] = (uint8_t)(move_of(k, (uint32_t)a, false) * scale_a);
        mvb[threadIdx.x] = (uint8_t)(move_of(k, (uint32_t)a, true) * scale_b);
    }
    for (uint32_t i = threadIdx.x; i < 320u + 640u; i += blockDim.x) {
        const bool wide = i >= 320u;
        const uint32_t j = wide ? i - 320u : i;
        const int k = min((int)(wide ? j >> 6 : j >> 5), 8);
        const uint32_t aa = wide ? min((j >> 3) & 7u, 4u) : min((j & 31u) / 5u, 4u);
        const uint32_t ab = wide ? min(j & 7u, 4u) : (j & 31u) % 5u;
        (wide ? mvs : mvj)[j] = (uint8_t)(move_of(k, aa, false) * scale_a + move_of(k, ab, true) * scale_b);
    }
    if (threadIdx.x >= 96 && threadIdx.x < 96 + 20) {
        const int row = threadIdx.x - 96, k = row >> 1, nl = (row & 1) + 1;
        uint32_t t[4] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu };
        if (k < 9) {
            const double pr = __dmul_rn(P.mp[k], nl == 2 ? 0.25 : 0.5);
            double cc = k == 0 ? 0.0 : E.e[k - 1];
            for (int j = 0; j + 1 < (1 << nl); ++j) {
                cc = __dadd_rn(cc, pr);
                const uint64_t th = slip_thr(cc);      // 0 here: the danger list sends the draws that could use it to the walk
                t[j] = (th == 0ull || th > 0xFFFFFFFFull) ? 0xFFFFFFFFu : (uint32_t)(th - 1ull);
            }
        }
        if (nl == 1) t[1] = t[0];              // 2-way row {t, t, never}: #(r > t_j) = 0 / 2, the 4-way rows count 0 .. 3, so that
                                               // the entry of the chosen slot is ALWAYS 2 bytes * count behind slot 0
        sl[row * 4 + 0] = t[0]; sl[row * 4 + 1] = t[1]; sl[row * 4 + 2] = t[2]; sl[row * 4 + 3] = t[3];
    }
    const uint32_t low = (1u << (32 - bits)) - 1u;
    for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) {
        const uint32_t lo = b << (32 - bits), hi = lo | low;
        const double ulo = u_from_rng32(lo), uhi = u_from_rng32(hi);
        uint32_t kl = 0, kh = 0;
#pragma unroll
        for (int j = 0; j < 9; ++j) { kl += E.e[j] <= ulo ? 1u : 0u; kh += E.e[j] <= uhi ? 1u : 0u; }
        uint32_t thr = 0xFFFFFFFFu;
        if (kh == kl + 1u) thr = (uint32_t)(slip_thr(E.e[kl]) - 1ull);     // the step lies inside (lo, hi]: thr in [lo, hi)
        else if (kh != kl) kl = 9u;
#pragma unroll
        for (uint32_t i = 0; i < 12u; ++i)                                   // constant indices: dg stays in the constant bank
            if (i < dg.n && (dg.r[i] >> (32 - bits)) == b) kl = 9u;
        if (kl == 9u) thr = 0xFFFFFFFFu;
        kt[b] = (kl << (32 - bits)) + (low - (thr & low));
    }
}
__device__ __forceinline__ SlipInt slip_int_ctx(const uint8_t* base, const SlipBits& lb)
{
    SlipInt f = { smem_u32(base), lb.shift, lb.mask };
    return f;
}
// combination index of a 32-bit draw: 0 .. 8, or 9 = walk (undecided bucket -- incl. the bucket of a listed draw -- or
// "no sum exceeds u")
__device__ __forceinline__ uint32_t slip_int_k(const SlipInt& f, uint32_t r32)
{
    uint32_t e;
    asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(f.kt() + (r32 >> f.shift) * 4u));
    return (e + (r32 & f.mask)) >> f.shift;
}
__device__ __forceinline__ uint32_t lds_u32_r(uint32_t addr)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16_r(uint32_t addr)
{
    uint32_t v;
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_v4_r(uint32_t addr)
{
    uint4 v;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// One env-step with a 32-bit draw whose combination k (slip_int_k, < 9) and move pair (mv = byte offset of the pair
// inside the table row, from mva + mvb, mvj or mvs) the caller has looked up.
__device__ __forceinline__ TblOut table_step_slip_int_at(const TblCtx& c, const SlipInt& f, uint32_t s, uint32_t k, uint32_t mv,
                                                         uint32_t r32, uint32_t rsel4)
{
    const uint32_t obsi = min(s & 0xFFFFu, c.maxobs);
    uint32_t ent = c.tbl + obsi * 200u + mv;
    int32_t e = lds_s16_r(ent);
    if ((uint32_t)e & 0x3000u) {                               // 3 % of the (state, move pair)s: 2 or 4 outcomes
        // row (k, nl) of the slot thresholds: 16 bytes at sl + k * 32 + (nl - 1) * 16, and nl * 16 = (e >> 8) & 0x30
        const uint4 t = lds_v4_r(f.sl() - 16u + k * 32u + (((uint32_t)e >> 8) & 0x30u));
        if (r32 > t.x) ent += 2u;                              // 2-way: draw value 2 * slot (row {t, t, never}); 4-way: slot
        if (r32 > t.y) ent += 2u;
        if (r32 > t.z) ent += 2u;
        e = lds_s16_r(ent);
    }
    return table_finish(c, s, e, rsel4);
}
// Separate action values (kernels with a folded policy).  walk == true: the result is void, the caller takes the
// reference's walk.  CLAMP_ACT: the action values come from a caller's stream (Philox-decoded and policy-table actions
// are < 5 already).
template <bool CLAMP_ACT>
__device__ __forceinline__ TblOut table_step_slip_int(const TblCtx& c, const SlipInt& f, uint32_t s,
                                                      uint32_t aa, uint32_t ab, uint32_t r32, uint32_t rsel4, bool& walk)
{
    if (CLAMP_ACT) { aa = min(aa, 4u); ab = min(ab, 4u); }
    const uint32_t k = slip_int_k(f, r32);
    walk = k >= 9u;
    return table_step_slip_int_at(c, f, s, k, lds_u8_r(f.mva() + k * 8u + aa) + lds_u8_r(f.mvb() + k * 8u + ab), r32, rsel4);
}
// The joint action as ONE index: WIDE = false: ja = aa * 5 + ab < 25 (the Philox joint action mulhi(w, 25));
// WIDE = true: aa * 8 + ab < 64 (two 3-bit fields of caller-supplied action bytes).
template <bool WIDE>
__device__ __forceinline__ TblOut table_step_slip_int_j(const TblCtx& c, const SlipInt& f, uint32_t s,
                                                        uint32_t j, uint32_t r32, uint32_t rsel4, bool& walk)
{
    const uint32_t k = slip_int_k(f, r32);
    walk = k >= 9u;
    return table_step_slip_int_at(c, f, s, k, lds_u8_r((WIDE ? f.mvs() + k * 64u : f.mvj() + k * 32u) + j), r32, rsel4);
}

// The reference's walk as an out-of-line call for the integer fast path's (in practice never taken) fallback: inlined
// four times it cost the K1 kernel 150-190 bytes of spills at 64 registers.
__device__ __noinline__ uint2 table_step_slip_call(uint32_t tbl, uint32_t isd, uint32_t last, uint32_t prt, uint32_t first_k,
                                                   uint32_t s, uint32_t aa, uint32_t ab, uint32_t r32, uint32_t rsel4)
{
    const TblCtx c = { tbl, isd, last, last / 100u };
    const SlipCtx sc = { prt, first_k };
    const TblOut o = table_step_slip(c, sc, s, aa, ab, u_from_rng32(r32), rsel4);
    // state | obs, flags, reward and the post-reset observation packed into the second word
    return make_uint2(o.state, o.obs | (o.flags << 12) | (((uint32_t)o.rew_i & 3u) << 14) | (o.reset_obs << 16));
}
__device__ __forceinline__ TblOut table_step_slip_walk(const TblCtx& c, const SlipCtx& sc, uint32_t s, uint32_t aa, uint32_t ab,
                                                       uint32_t r32, uint32_t rsel4)
{
    const uint2 w = table_step_slip_call(c.tbl, c.isd, c.last, sc.prt, sc.first_k, s, aa, ab, r32, rsel4);
    TblOut o;
    o.state = w.x; o.obs = w.y & kTblObsMask; o.flags = (w.y >> 12) & 3u;
    o.rew_i = (int32_t)(int16_t)(w.y & 0xFFFFu) >> 14; o.reset_obs = w.y >> 16;
    return o;
}

// K1 for slip_prob > 0 and 32-bit draws (injected rng32: 24 B / env-step; Philox: 19 B): k_step_table's structure --
// persistent 1024-thread CTAs, two groups in flight + register prefetch of the next pair -- around table_step_slip_int.
// Shared-memory image: [table][isd 16 B][policy a][policy b][look-up tables of the fast path].
struct GroupS { uint4 s; uint32_t a, b, r; uint4 d; };
template <bool PHILOX>
__device__ __forceinline__ GroupS load_group_slip(const uint4* st, const uint32_t* aa, const uint32_t* ab, const uint32_t* rg,
                                                  const uint4* dr, int64_t g)
{
    GroupS x;
    x.s = ld_keep(st + g);
    x.a = ld_stream(aa + g);
    x.b = ld_stream(ab + g);
    if (PHILOX) { x.r = 0u; x.d = make_uint4(0, 0, 0, 0); }
    else { x.r = ld_stream(rg + g); x.d = __ldcs(dr + g); }
    return x;
}
// Launch shape, measured on B200 at 2^24 envs (profiles/r02g_time_k1slip.log; groups in flight x threads):
//   injected rng32  2 x 768: 151 G   1 x 1024: 183 G   1 x 768: 186 G   2 x 512: 187 G   1 x 512: 203 G env-steps/s
//   Philox          2 x 768: 187 G   1 x 1024: 215 G   1 x 768: 239 G   2 x 512: 226 G   1 x 512: 232 G
// (a group + its prefetched successor are 22 data registers; 1024 threads = 64 registers spill, 512 leave 16 warps)
#ifndef SOCCER_SLIP_I_GROUPS
#define SOCCER_SLIP_I_GROUPS 1         // groups of 4 envs a thread has in flight (1 or 2)
#endif
#ifndef SOCCER_SLIP_I_L2_PREFETCH
#define SOCCER_SLIP_I_L2_PREFETCH 0    // iterations ahead of the register prefetch to pull into L2 (0 = off)
#endif
#ifndef SOCCER_SLIP_I_RING
#define SOCCER_SLIP_I_RING 1           // cp.async input ring where shared memory allows (0: register prefetch everywhere)
#endif
#ifndef SOCCER_K1_SLIP_WALK_MERGED
#define SOCCER_K1_SLIP_WALK_MERGED 1   // one walk test per group of 4 envs (else one per env)
#endif
#ifndef SOCCER_SLIP_I_THREADS
#define SOCCER_SLIP_I_THREADS 0        // 0: 512 (128 registers; at 768 threads = 80 registers the Philox variant loses 10 %)
#endif
template <bool PHILOX>
constexpr int slip_i_threads() { return SOCCER_SLIP_I_THREADS ? SOCCER_SLIP_I_THREADS : 512; }
// ---- asynchronous input ring (cp.async): global -> shared without passing through registers
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds_v4_v(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32_v(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// bytes of one ring stage per thread: state (16) + action bytes (4 + 4), + draw bytes (4) + 32-bit draws (16) when injected
__host__ __device__ constexpr int slip_ring_stage_bytes(bool philox) { return philox ? 24 : 44; }
// groups of a thread in flight behind the one being stepped; measured at 2^24 envs (profiles/r02z_time.log):
//   injected rng32 (44 B per group and thread): register prefetch 209 G, 2 stages 228 G, 3 stages (10-bit bucket table) 219 G
//   Philox         (24 B):                      register prefetch 246 G, 2 stages 256 G, 3 stages 270 G, 4 stages 266 G
// (injected, only state + draws in a 3-stage ring and the action / draw bytes through L2 prefetch + registers: 176 G)
#ifndef SOCCER_SLIP_I_RING_STAGES
#define SOCCER_SLIP_I_RING_STAGES 0    // 0: 2 for injected draws, 3 for Philox
#endif
__host__ __device__ constexpr int slip_ring_stages(bool philox)
{
    return SOCCER_SLIP_I_RING_STAGES ? SOCCER_SLIP_I_RING_STAGES : (philox ? 3 : 2);
}

// POLICY: a folded player (its table policy in shared memory, separate move look-ups); else both action bytes -> one index.
// RING: the input streams of the next TWO groups of a thread are in flight as cp.async copies into a per-thread
// shared-memory ring (the shared memory the table leaves free) while the current group is stepped out of registers --
// twice the bytes in flight of the register prefetch, with 11 registers fewer.  ncu of the register-prefetch version:
// 16 warps per SM, long-scoreboard stall 7.3 per issue, DRAM 55 % busy = the kernel waits for its HBM loads.
template <bool RESET_OBS, bool PHILOX, bool POLICY, bool RING>
__global__ void __launch_bounds__((slip_i_threads<PHILOX>()), 1)
k_step_table_slip_i(const PitchDev P, const uint16_t* __restrict__ gtable, uint32_t table_bytes,
                    const SlipE E, const SlipDanger dg, const SlipBits lut_bits,
                    uint32_t* __restrict__ state, const uint8_t* __restrict__ act_a, const uint8_t* __restrict__ act_b,
                    const uint8_t* __restrict__ rng, const uint32_t* __restrict__ draw, int32_t* __restrict__ obs,
                    float* __restrict__ reward, uint8_t* __restrict__ flags, int32_t* __restrict__ reset_obs, int64_t n_groups,
                    const SlipExtra ex)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(16) double prt[kPrtDoubles];
    const bool has_pol = POLICY && (ex.policy_a || ex.policy_b);
    const uint32_t pol_total = has_pol ? 2u * (((uint32_t)P.nS + 15u) & ~15u) : 0u;
    uint8_t* luts = smem_raw + table_bytes + 16 + pol_total;
    slip_build_prt(prt, P);
    slip_int_build_luts(luts, E, P, dg, lut_bits.bits);
    stage_table(smem_raw, gtable, table_bytes, &bar, P);                                   // ends with __syncthreads()
    TblCtx c = make_ctx(smem_raw, table_bytes, P);
    SlipCtx sc = { smem_u32(prt), slip_first_k(P) };
    SlipInt sf = slip_int_ctx(luts, lut_bits);
    K1Policy pol = { 0u, 0u };
    if (has_pol) pol = stage_k1_policies(smem_raw + table_bytes + 16, ex.policy_a, ex.policy_b, P.nS);
    const bool flip = POLICY && pol.pol_a != 0u;
    uint4* st4 = reinterpret_cast<uint4*>(state);
    const uint32_t* a4 = reinterpret_cast<const uint32_t*>(act_a ? act_a : act_b);
    const uint32_t* b4 = reinterpret_cast<const uint32_t*>(act_b ? act_b : act_a);
    const uint32_t* r4 = reinterpret_cast<const uint32_t*>(rng);
    const uint4* d4 = reinterpret_cast<const uint4*>(draw);
    uint4* o4 = reinterpret_cast<uint4*>(obs);
    uint4* w4 = reinterpret_cast<uint4*>(reward);
    uint32_t* f4 = reinterpret_cast<uint32_t*>(flags);
    uint4* q4 = reinterpret_cast<uint4*>(reset_obs);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool one = g < n_groups, two = SOCCER_SLIP_I_GROUPS == 2 && g + stride < n_groups;
    GroupS x0 = {}, x1 = {};
    if (RING) {
        // the first two groups of the thread go out before the table wait (same ring layout as in the loop below)
        const uint32_t T = blockDim.x, tid = threadIdx.x;
        const uint32_t ring0 = smem_u32(luts) + ((uint32_t)slip_int_lut_bytes(lut_bits.bits) + 15u & ~15u);
        const uint32_t stage_bytes = T * (uint32_t)slip_ring_stage_bytes(PHILOX);
        const uint32_t o_d = T * 16u, o_a = PHILOX ? T * 16u : T * 32u, o_b = o_a + T * 4u, o_r = o_b + T * 4u;
#pragma unroll
        for (uint32_t k = 0; k < (uint32_t)slip_ring_stages(PHILOX); ++k) {
            const int64_t gg = g + (int64_t)k * stride;
            if (gg < n_groups) {
                const uint32_t base = ring0 + k * stage_bytes;
                cp_async16(base + tid * 16u, st4 + gg);
                if (!PHILOX) cp_async16(base + o_d + tid * 16u, d4 + gg);
                cp_async4(base + o_a + tid * 4u, a4 + gg);
                cp_async4(base + o_b + tid * 4u, b4 + gg);
                if (!PHILOX) cp_async4(base + o_r + tid * 4u, r4 + gg);
            }
            cp_async_commit();
        }
    } else {
        if (one) x0 = load_group_slip<PHILOX>(st4, a4, b4, r4, d4, g);
        if (two) x1 = load_group_slip<PHILOX>(st4, a4, b4, r4, d4, g + stride);
    }
    wait_table(&bar);
    launder(c.tbl); launder(c.isd); launder(sc.prt);
    launder(sf.base);
    launder(pol.pol_a); launder(pol.pol_b);
    auto do_group = [&](GroupS& x, int64_t gg) {
        if (PHILOX) {
            uint32_t w[4];
            philox_words4(ex.key, gg, w);
            x.d = make_uint4(philox_r32(w[0]), philox_r32(w[1]), philox_r32(w[2]), philox_r32(w[3]));
            x.r = (w[0] & 0xCu) | ((w[1] & 0xCu) << 8) | ((w[2] & 0xCu) << 16) | ((w[3] & 0xCu) << 24);
        }
        const uint32_t rs4 = x.r & 0x0C0C0C0Cu;
        const uint32_t sv[4] = { x.s.x, x.s.y, x.s.z, x.s.w };
        const uint32_t r32[4] = { x.d.x, x.d.y, x.d.z, x.d.w };
        uint32_t so[4], oo[4], ro[4], rr[4], ff[4];
        // both players' action bytes as one index per env, byte-parallel: aa * 8 + ab (3-bit fields; > 4 acts as 4)
        const uint32_t j4 = ((x.a & 0x07070707u) << 3) | (x.b & 0x07070707u);
        uint32_t walks = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t rsel4 = __byte_perm(rs4, 0, 0x4440 + e);
            bool walk;
            TblOut o;
            if (POLICY) {
                uint32_t aa = __byte_perm(x.a, 0, 0x4440 + e), ab = __byte_perm(x.b, 0, 0x4440 + e);
                const uint32_t cs = min(sv[e] & 0xFFFFu, c.maxobs);
                if (pol.pol_a) aa = lds_u8_r(pol.pol_a + cs);
                if (pol.pol_b) ab = lds_u8_r(pol.pol_b + cs);
                o = table_step_slip_int<true>(c, sf, sv[e], aa, ab, r32[e], rsel4, walk);
            } else {
                o = table_step_slip_int_j<true>(c, sf, sv[e], __byte_perm(j4, 0, 0x4440 + e), r32[e], rsel4, walk);
            }
#if SOCCER_K1_SLIP_WALK_MERGED
            walks |= walk ? 1u << e : 0u;
#else
            if (walk) {
                uint32_t aa = __byte_perm(x.a, 0, 0x4440 + e), ab = __byte_perm(x.b, 0, 0x4440 + e);
                const uint32_t cs = min(sv[e] & 0xFFFFu, c.maxobs);
                if (POLICY && pol.pol_a) aa = lds_u8_r(pol.pol_a + cs);
                if (POLICY && pol.pol_b) ab = lds_u8_r(pol.pol_b + cs);
                o = table_step_slip_walk(c, sc, sv[e], aa, ab, r32[e], rsel4);
            }
#endif
            so[e] = o.state; oo[e] = o.obs; rr[e] = __float_as_uint((float)(flip ? -o.rew_i : o.rew_i));
            ro[e] = o.reset_obs; ff[e] = o.flags;
        }
        if (walks) {                                         // (in practice never): redo those envs by the reference's walk
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (!((walks >> e) & 1u)) continue;
                uint32_t aa = __byte_perm(x.a, 0, 0x4440 + e), ab = __byte_perm(x.b, 0, 0x4440 + e);
                const uint32_t cs = min(sv[e] & 0xFFFFu, c.maxobs);
                if (POLICY && pol.pol_a) aa = lds_u8_r(pol.pol_a + cs);
                if (POLICY && pol.pol_b) ab = lds_u8_r(pol.pol_b + cs);
                const TblOut o = table_step_slip_walk(c, sc, sv[e], aa, ab, r32[e], __byte_perm(rs4, 0, 0x4440 + e));
                so[e] = o.state; oo[e] = o.obs; rr[e] = __float_as_uint((float)(flip ? -o.rew_i : o.rew_i));
                ro[e] = o.reset_obs; ff[e] = o.flags;
            }
        }
        st_keep(st4 + gg, make_uint4(so[0], so[1], so[2], so[3]));
        st_stream(o4 + gg, make_uint4(oo[0], oo[1], oo[2], oo[3]));
        st_stream(w4 + gg, make_uint4(rr[0], rr[1], rr[2], rr[3]));
        st_stream(f4 + gg, __byte_perm(__byte_perm(ff[0], ff[1], 0x0040), __byte_perm(ff[2], ff[3], 0x0040), 0x5410));
        if (RESET_OBS) st_stream(q4 + gg, make_uint4(ro[0], ro[1], ro[2], ro[3]));
    };
    if (RING) {
        // stage k of the ring: [state 16 B x T][draws 16 B x T (injected)][act_a 4 B x T][act_b 4 B x T][rng8 4 B x T (injected)]
        const uint32_t T = blockDim.x, tid = threadIdx.x;
        const uint32_t ring0 = smem_u32(luts) + ((uint32_t)slip_int_lut_bytes(lut_bits.bits) + 15u & ~15u);
        const uint32_t stage_bytes = T * (uint32_t)slip_ring_stage_bytes(PHILOX);
        const uint32_t o_d = T * 16u, o_a = PHILOX ? T * 16u : T * 32u, o_b = o_a + T * 4u, o_r = o_b + T * 4u;
        auto issue = [&](uint32_t k, int64_t gg) {
            if (gg < n_groups) {
                const uint32_t base = ring0 + k * stage_bytes;
                cp_async16(base + tid * 16u, st4 + gg);
                if (!PHILOX) cp_async16(base + o_d + tid * 16u, d4 + gg);
                cp_async4(base + o_a + tid * 4u, a4 + gg);
                cp_async4(base + o_b + tid * 4u, b4 + gg);
                if (!PHILOX) cp_async4(base + o_r + tid * 4u, r4 + gg);
            }
            cp_async_commit();                                       // (an empty group keeps the count uniform)
        };
        auto fetch = [&](uint32_t k) {
            const uint32_t base = ring0 + k * stage_bytes;
            GroupS x;
            x.s = lds_v4_v(base + tid * 16u);
            x.a = lds_u32_v(base + o_a + tid * 4u);
            x.b = lds_u32_v(base + o_b + tid * 4u);
            if (PHILOX) { x.r = 0u; x.d = make_uint4(0, 0, 0, 0); }
            else { x.r = lds_u32_v(base + o_r + tid * 4u); x.d = lds_v4_v(base + o_d + tid * 16u); }
            return x;
        };
        uint32_t k = 0;
        constexpr int S = slip_ring_stages(PHILOX);
        cp_async_wait<S - 1>();                                      // the first group (issued before the table wait)
        if (one) x0 = fetch(0);
        while (one) {
            issue(k, g + S * stride);                                // stage k is free: its group sits in x0
            do_group(x0, g);
            cp_async_wait<S - 1>();                                  // group g + stride has landed; the later ones may be in flight
            k = k + 1u == (uint32_t)S ? 0u : k + 1u;
            g += stride;
            one = g < n_groups;
            if (one) x0 = fetch(k);
        }
        cp_async_wait<0>();
        return;
    }
    if (SOCCER_SLIP_I_GROUPS == 1) {
        // one group in flight + register prefetch of the next one (half the data registers: 1024 threads fit 64 registers)
        if (two) { /* x1 was loaded above only in the two-group variant */ }
        while (one) {
            const int64_t gn = g + stride;
            const bool n_one = gn < n_groups;
            GroupS y0 = x0;
            if (n_one) y0 = load_group_slip<PHILOX>(st4, a4, b4, r4, d4, gn);
            if (SOCCER_SLIP_I_L2_PREFETCH) {
                // deepen the load pipeline without registers: the lines of the group SOCCER_SLIP_I_L2_PREFETCH iterations
                // ahead are pulled into L2 now (this kernel has 16-24 warps per SM and one group per thread in flight)
                const int64_t gp = gn + (int64_t)SOCCER_SLIP_I_L2_PREFETCH * stride;
                if (gp < n_groups) {
                    prefetch_l2(st4 + gp);
                    if (!PHILOX) prefetch_l2(d4 + gp);
                    if ((threadIdx.x & 31) == 0) {
                        prefetch_l2(a4 + gp); prefetch_l2(b4 + gp);
                        if (!PHILOX) prefetch_l2(r4 + gp);
                    }
                }
            }
            do_group(x0, g);
            x0 = y0; g = gn; one = n_one;
        }
        return;
    }
    while (one) {
        const int64_t gn = g + 2 * stride;
        const bool n_one = gn < n_groups, n_two = gn + stride < n_groups;
        GroupS y0 = x0, y1 = x1;
        if (n_one) y0 = load_group_slip<PHILOX>(st4, a4, b4, r4, d4, gn);          // prefetch the next pair
        if (n_two) y1 = load_group_slip<PHILOX>(st4, a4, b4, r4, d4, gn + stride);
        do_group(x0, g);
        if (two) do_group(x1, g + stride);
        x0 = y0; x1 = y1; g = gn; one = n_one; two = n_two;
    }
}

// shared-memory image: [table][isd 16 B][slip index][policy a][policy b][queues]
template <bool RESET_OBS, int DRAW>
__global__ void __launch_bounds__(kTableThreads, 1)
k_step_table_slip_q(const PitchDev P, const uint16_t* __restrict__ gtable, uint32_t table_bytes,
                    const uint8_t* __restrict__ gfc, uint32_t fc_bytes, const SlipE E,
                    uint32_t* __restrict__ state, const uint8_t* __restrict__ act_a, const uint8_t* __restrict__ act_b,
                    const uint8_t* __restrict__ rng, const void* __restrict__ draw, int32_t* __restrict__ obs,
                    float* __restrict__ reward, uint8_t* __restrict__ flags, int32_t* __restrict__ reset_obs, int64_t n_groups,
                    const SlipExtra ex)
{
    constexpr bool F64 = DRAW == kDrawF64;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(16) double prt[kPrtDoubles];
    __shared__ uint8_t cacb[16], mv3[16];
    __shared__ __align__(16) uint8_t klut[F64 ? 16 : (1 << kSlipLutBits)];
    slip_build_prt(prt, P);
    slip_fast_build_luts<F64>(cacb, mv3, klut, E);
    stage_table_and_index(smem_raw, gtable, table_bytes, gfc, fc_bytes, &bar, P);
    TblCtx c = make_ctx(smem_raw, table_bytes, P);
    SlipCtx sc = { smem_u32(prt), slip_first_k(P) };
    SlipFast sf = { c.isd + 16u, smem_u32(cacb), smem_u32(mv3), smem_u32(klut) };
    const bool has_pol = ex.policy_a || ex.policy_b;
    const uint32_t pol_total = has_pol ? 2u * (((uint32_t)P.nS + 15u) & ~15u) : 0u;
    K1Policy pol = { 0u, 0u };
    if (has_pol) pol = stage_k1_policies(smem_raw + table_bytes + 16 + fc_bytes, ex.policy_a, ex.policy_b, P.nS);
    const bool flip = pol.pol_a != 0u;
    uint8_t* queue = smem_raw + table_bytes + 16 + fc_bytes + pol_total;
    uint4* st4 = reinterpret_cast<uint4*>(state);
    const uint8_t* act_a1 = act_a ? act_a : act_b;           // a folded player's stream does not exist: alias the other
    const uint8_t* act_b1 = act_b ? act_b : act_a;
    const uint32_t* a4 = reinterpret_cast<const uint32_t*>(act_a1);
    const uint32_t* b4 = reinterpret_cast<const uint32_t*>(act_b1);
    const uint32_t* r4 = reinterpret_cast<const uint32_t*>(rng);
    uint4* o4 = reinterpret_cast<uint4*>(obs);
    uint4* w4 = reinterpret_cast<uint4*>(reward);
    uint32_t* f4 = reinterpret_cast<uint32_t*>(flags);
    uint4* q4 = reinterpret_cast<uint4*>(reset_obs);
    wait_table(&bar);
    launder(c.tbl); launder(c.isd); launder(sc.prt);        // the relaxed loads below stay below the wait
    launder(sf.fc); launder(sf.cacb); launder(sf.mv3); launder(sf.klut);
    launder(pol.pol_a); launder(pol.pol_b);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    uint8_t* wq = queue + (threadIdx.x >> 5) * kSlipQueueWarp;   // this warp's queue: (half << 7 | lane << 2 | env) of the iteration
    // warp-level deferral: no CTA barrier, so the warps of the CTA drift apart and their loads overlap the others' walks
    uint32_t cnt = 0;                                        // warp-uniform: queued envs of this warp
    // the walk of up to 32 queued envs of the iteration at `wbase`, one per lane (exact: table_step_slip)
    uint32_t carry = 0;                                      // warp-uniform: entries [0, carry) come from the previous iteration
    int64_t wbase_prev = 0;
    auto walk = [&](int64_t wbase, uint32_t first, uint32_t count) {
        if (lane < count) {
            const uint32_t idx = first + lane, id = wq[idx];
            const int64_t wb = idx < carry ? wbase_prev : wbase;
            const int64_t env = (wb + (int64_t)(id >> 7) * stride + ((id >> 2) & 31u)) * 4 + (id & 3u);
            const uint32_t s = __ldcg(state + env);
            uint32_t rg; double ud;
            if (DRAW == kDrawPhilox) {
                const uint32_t w = philox_word(ex.key.seed, ex.key.env_id_base + (uint64_t)env, ex.key.step);
                rg = w & 0xCu; ud = u_from_rng32(philox_r32(w));
            } else {
                rg = rng[env];
                ud = F64 ? reinterpret_cast<const double*>(draw)[env] : u_from_rng32(reinterpret_cast<const uint32_t*>(draw)[env]);
            }
            uint32_t aa = act_a1[env], ab = act_b1[env];
            const uint32_t cur = min(s & 0xFFFFu, c.maxobs);
            if (pol.pol_a) aa = lds_u8_r(pol.pol_a + cur);
            if (pol.pol_b) ab = lds_u8_r(pol.pol_b + cur);
            const TblOut o = table_step_slip(c, sc, s, aa, ab, ud, rg & 0xCu);
            state[env] = o.state; obs[env] = (int32_t)o.obs; reward[env] = (float)(flip ? -o.rew_i : o.rew_i);
            flags[env] = (uint8_t)o.flags;
            if (RESET_OBS) reset_obs[env] = (int32_t)o.reset_obs;
        }
    };
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = g < n_groups;
    SlipIn cur = {};
    if (valid) cur = slip_load_in<DRAW, SOCCER_SLIP_PLAIN_LOADS != 0>(st4, a4, b4, r4, draw, g);
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n_groups; base += 2 * stride) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            // register prefetch of the thread's next group (the loop has no other way to overlap HBM latency with the
            // shared-memory look-up chains of the current one)
            const int64_t gn = g + stride;
            const bool validn = gn < n_groups;
            SlipIn nxt = cur;
            if (validn) nxt = slip_load_in<DRAW, SOCCER_SLIP_PLAIN_LOADS != 0>(st4, a4, b4, r4, draw, gn);
            if (DRAW == kDrawPhilox && valid) slip_philox_fill(cur, ex.key, g);
            const Group4 x = cur.x;
            double u[4] = { 0.0, 0.0, 0.0, 0.0 };
            uint32_t r32[4] = { cur.d0.x, cur.d0.y, cur.d0.z, cur.d0.w };
            if (F64) {
                u[0] = __hiloint2double((int)cur.d0.y, (int)cur.d0.x); u[1] = __hiloint2double((int)cur.d0.w, (int)cur.d0.z);
                u[2] = __hiloint2double((int)cur.d1.y, (int)cur.d1.x); u[3] = __hiloint2double((int)cur.d1.w, (int)cur.d1.z);
            }
            const uint32_t rs4 = x.r & 0x0C0C0C0Cu;
            const uint32_t sv[4] = { x.s.x, x.s.y, x.s.z, x.s.w };
            uint32_t so[4], oo[4], ro[4], rr[4], ff[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                bool defer;
                uint32_t aa = __byte_perm(x.a, 0, 0x4440 + e), ab = __byte_perm(x.b, 0, 0x4440 + e);
                if (pol.pol_a | pol.pol_b) {                 // warp-uniform
                    const uint32_t cs = min(sv[e] & 0xFFFFu, c.maxobs);
                    if (pol.pol_a) aa = lds_u8_r(pol.pol_a + cs);
                    if (pol.pol_b) ab = lds_u8_r(pol.pol_b + cs);
                }
                const TblOut o = table_step_slip_fast<F64>(c, sf, E, sv[e], aa, ab, u[e], r32[e],
                                                           __byte_perm(rs4, 0, 0x4440 + e), defer);
                defer &= valid;
                so[e] = defer ? sv[e] : o.state;             // deferred: the ORIGINAL state stays for the walk to read
                oo[e] = o.obs; rr[e] = __float_as_uint((float)(flip ? -o.rew_i : o.rew_i)); ro[e] = o.reset_obs; ff[e] = o.flags;
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, defer);
                if (defer) wq[cnt + __popc(m & lt_mask)] = (uint8_t)((h << 7) | (lane << 2) | e);
                cnt += __popc(m);
            }
            if (valid) {
                st_keep(st4 + g, make_uint4(so[0], so[1], so[2], so[3]));
                // plain (evict-normal) stores, not st.global.cs: the deferred envs' 1- and 4-byte fix-ups follow within
                // microseconds and must find these sectors in L2 -- a partial write to a sector already evicted to
                // (ECC) HBM costs a read-modify-write there
                SOCCER_SLIP_ST(o4 + g, make_uint4(oo[0], oo[1], oo[2], oo[3]));
                SOCCER_SLIP_ST(w4 + g, make_uint4(rr[0], rr[1], rr[2], rr[3]));
                SOCCER_SLIP_ST(f4 + g, __byte_perm(__byte_perm(ff[0], ff[1], 0x0040), __byte_perm(ff[2], ff[3], 0x0040), 0x5410));
                if (RESET_OBS) SOCCER_SLIP_ST(q4 + g, make_uint4(ro[0], ro[1], ro[2], ro[3]));
            }
            cur = nxt; g = gn; valid = validn;
        }
        __syncwarp();                                        // queue entries and the placeholder stores ordered before the walks
#if SOCCER_SLIP_CARRY
        // Walk FULL warps of queued envs; fewer than 32 left over wait ONE iteration for company (they sit at the front of
        // the queue, so the next iteration's first pass takes them), unless this was the last iteration or they already
        // waited.  With evict-normal loads and stores their lines are still in L2 when they are patched.
        const int64_t wbase = base + (threadIdx.x & ~31u);
        const bool last = base + 2 * stride >= n_groups;
        uint32_t first = 0;
        while (cnt - first >= 32u) { walk(wbase, first, 32u); first += 32u; }
        uint32_t rem = cnt - first;
        if (rem && (last || first < carry)) { walk(wbase, first, rem); first = cnt; rem = 0; }
        __syncwarp();
        if (rem) {                                           // compact the leftovers to the front
            const uint8_t v = lane < rem ? wq[first + lane] : (uint8_t)0;
            __syncwarp();
            if (lane < rem) wq[lane] = v;
        }
        carry = rem; cnt = rem; wbase_prev = wbase;
#else
        // Walk them right away, while the lines the fast pass has just written are still in L2.
        for (uint32_t first = 0; first < cnt; first += 32u) walk(base + (threadIdx.x & ~31u), first, min(32u, cnt - first));
        cnt = 0;
#endif
        __syncwarp();                                        // queue settled before the warp refills it
    }
}

// scalar tail / misaligned fallback of the slip table path (global-memory table, one env per thread)
__global__ void __launch_bounds__(kThreads)
k_step_table_slip_scalar(const PitchDev P, const uint16_t* __restrict__ gtable, uint32_t* __restrict__ state,
                         const uint8_t* __restrict__ act_a, const uint8_t* __restrict__ act_b,
                         const uint8_t* __restrict__ rng, const uint32_t* __restrict__ rng32,
                         const double* __restrict__ rngf64, int32_t* __restrict__ obs, float* __restrict__ reward,
                         uint8_t* __restrict__ flags, int32_t* __restrict__ reset_obs, int64_t n, int use_philox,
                         const SlipExtra ex)
{
    const uint32_t last_row = (uint32_t)P.nS - 1u;
    const bool flip = ex.policy_a != nullptr;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t s = state[i];
        uint32_t rg; double u;
        if (use_philox) {
            const uint32_t w = philox_word(ex.key.seed, ex.key.env_id_base + (uint64_t)i, ex.key.step);
            rg = w & 0xCu; u = u_from_rng32(philox_r32(w));
        } else {
            rg = rng[i]; u = rngf64 ? rngf64[i] : u_from_rng32(rng32[i]);
        }
        const uint32_t cur = min(s & 0xFFFFu, last_row);
        const LdGlobal ld = { reinterpret_cast<const uint8_t*>(gtable) + (size_t)cur * 200u };
        const uint32_t aa = ex.policy_a ? (uint32_t)min(max((int)ex.policy_a[cur], 0), 4) : (uint32_t)act_a[i];
        const uint32_t ab = ex.policy_b ? (uint32_t)min(max((int)ex.policy_b[cur], 0), 4) : (uint32_t)act_b[i];
        const uint32_t pick = slip_pick(P, ld, aa, ab, u);
        const int32_t e = (int32_t)(int16_t)ld(pick);
        const uint32_t nobs = (uint32_t)e & kTblObsMask;
        const bool done = nobs == 0;
        const uint32_t s1 = s + 0x10000u;
        const bool trunc = s1 >= kTruncWord, reset = done | trunc;
        const uint32_t ro = (uint32_t)P.isd_obs[(rg >> 2) & 3u];
        state[i] = reset ? ro : ((s1 & 0xFFFF0000u) | nobs);
        obs[i] = (int32_t)nobs;
        reward[i] = (float)(flip ? -(e >> 14) : (e >> 14));
        flags[i] = (uint8_t)((done ? 1u : 0u) + (trunc ? 2u : 0u));
        if (reset_obs) reset_obs[i] = (int32_t)(reset ? ro : nobs);
    }
}

__global__ void __launch_bounds__(kThreads)
k_convert_state(const PitchDev P, int32_t nS, const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                int32_t to_layout, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t s = in[i];
        if (to_layout == 1) {   // SOCCER_LAYOUT_INDEX
            const uint32_t a = s & 0xFFu, b = (s >> 8) & 0xFFu, p = (s >> 24) & 1u, t = (s >> 16) & 0xFFu;
            const bool bad = ((a | b) & kGoalBit) || (s & kNeedsReset);
            out[i] = (bad ? 0u : (uint32_t)obs_index(P, a, b, p)) | (t << 16);
        } else {
            const int32_t o = (int32_t)(s & 0xFFFFu);
            const uint32_t t = (s >> 16) & 0xFFu;
            out[i] = (o >= 1 && o < nS) ? (obs_to_packed(P, o) | (t << 16)) : (kNeedsReset | kGoalBit | (t << 16));
        }
    }
}

} // namespace soccer

// soccer_rules.cuh -- device-side game rules shared by every kernel (sm_100a).
//
// Formulation (NOT the reference's): players live on cell codes (one byte), the wall clamp and
// goal-mouth rule of SIM:364-373 are folded into a per-CTA shared-memory candidate table
// cand[cell][has_ball][move] that every CTA computes arithmetically at start-up (<= 2 KB), and
// the four collision cases of SIM:315-356 collapse into five byte compares and a handful of
// predicate ops.  SIM = /root/reference/gym_soccer/envs/soccer_simultaneous_env.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace soccer {

constexpr int kMaxField = 126;          // F = width*height must fit a 7-bit cell code
constexpr int kLutBytes = 256 * 16;     // cand[cell][has][8 moves]; sized so ANY byte-valued cell code stays in bounds
constexpr uint32_t kGoalBit = 0x80u;    // cell code of a goal cell: 0x80 | right<<6 | row
constexpr uint32_t kRightBit = 0x40u;
constexpr uint32_t kNeedsReset = 1u << 25;
constexpr int kMaxT = 100;              // SIM:404

// Kernel-parameter copy of the pitch (constructor products, SIM:48-65, 146-165).
struct PitchDev {
    int32_t  w;              // unpadded width
    int32_t  H;
    int32_t  F;              // w*H
    int32_t  Fm1;
    int32_t  nS;             // 1 + 2F(F-1)
    uint32_t nSm1, tlast;    // nS - 1 and nS * 100 - 1 (last entry of the step table): clamps read from the constant bank
    uint32_t goal_row_mask;  // bit r set iff r in goal_rows (SIM:60)
    uint32_t isd_state[4];   // packed start states (2-start pitches: s0,s0,s1,s1 so idx = r)
    int32_t  isd_obs[4];
    uint32_t isd_d1, isd_d2; // isd_state[r] == isd_state[0] + (r&1)*d1 + (r>>1)*d2  (mod 2^32)
    int32_t  obs_d1, obs_d2; // same for isd_obs
    double   mp[9];          // slip-combination probabilities (SIM:209-223)
    int32_t  slip;           // 1 iff slip_prob != 0
};

// slipped move ids in the order of SIM:205-206: mas[0] = (-dr, dc), mas[1] = (dr, -dc)
//   NOOP->NOOP, N->(E,W), S->(W,E), E->(S,N), W->(N,S)
__device__ __forceinline__ uint32_t slip_move(uint32_t act, int which)
{
    // nibble tables indexed by action: which==0: [0,3,4,2,1], which==1: [0,4,3,1,2]
    const uint32_t t = which == 0 ? 0x12430u : 0x21340u;
    return (t >> (act * 4)) & 7u;
}
// move used by player A / B in slip combination c (SIM:209-223): 0 intended, 1 mas[0], 2 mas[1]
__device__ __forceinline__ int combo_a(int c) { return (int)((0x221121000ull >> (c * 4)) & 3ull); } // [0,0,0,1,2,1,1,2,2]
__device__ __forceinline__ int combo_b(int c) { return (int)((0x212100210ull >> (c * 4)) & 3ull); } // [0,1,2,0,0,1,2,1,2]

// SIM:364-373 for one (cell, has_ball, move) -> candidate cell code.
__device__ __forceinline__ uint32_t next_cell_code(const PitchDev& P, uint32_t cell, uint32_t has, uint32_t move)
{
    int row = (int)cell / P.w;
    int col = (int)cell - row * P.w + 1;                 // padded column 1..w
    int dr = (move == 2) - (move == 1);                  // SIM:26-29 (d_col, d_row)
    int dc = (move == 3) - (move == 4);
    int nx = min(max(row + dr, 0), P.H - 1);             // SIM:365
    int ny = col + dc;                                   // SIM:366
    bool xoob = (ny == 0) || (ny == P.w + 1);            // SIM:369
    bool goal = xoob && ((P.goal_row_mask >> nx) & 1u) && has; // SIM:370
    if (xoob && !goal) ny = col;                         // SIM:371-372
    if (ny == 0 || ny == P.w + 1)
        return kGoalBit | (ny != 0 ? kRightBit : 0u) | (uint32_t)nx;
    return (uint32_t)(nx * P.w + ny - 1);
}

// Every CTA fills its own candidate table; index = cell*16 + has*8 + move.
__device__ __forceinline__ void build_cand_lut(uint8_t* lut, const PitchDev& P)
{
    for (int i = threadIdx.x; i < kLutBytes; i += blockDim.x) {
        uint32_t move = i & 7, has = (i >> 3) & 1, cell = i >> 4;
        // cells >= F never occur in a valid state; map them to themselves so garbage stays inert
        lut[i] = cell < (uint32_t)P.F ? (uint8_t)next_cell_code(P, cell, has, move > 4 ? 0u : move) : (uint8_t)cell;
    }
    __syncthreads();
}

// Simultaneous-move resolution, SIM:296-362, for field-cell states.
//   a, b    current cell codes; p possession; ma, mb the moves actually attempted (slipped or
//   not); a_noop / b_noop whether the ORIGINAL actions were NOOP (SIM:330-331, 338-339);
//   r       2-bit outcome index: 4-way outcome r (SIM:352-356 order), 2-way outcome r >> 1.
// Returns final cells / possession and log2(number of outcomes).
//
// With cell codes the four cases reduce to (proof in DESIGN.md "Collision algebra"):
//   aib = A's candidate is B's cell, bia = B's candidate is A's cell,
//   ast / bst = candidate equals own cell (NOOP or bounce)
//   stay  = (aib & (bia | bst)) | (bia & ast)          cases 1, 2, 3: nobody moves
//   c2    = (aib & b_noop) | (bia & a_noop)            case 2: possession flips
//   four  = (na == nb) & !stay                         case 4
struct Resolved { uint32_t a, b, p, nlog2; };

// na / nb: the candidate cells of the two players (SIM:308-309)
__device__ __forceinline__ Resolved resolve_cand(uint32_t na, uint32_t nb, uint32_t a, uint32_t b, uint32_t p,
                                                 bool a_noop, bool b_noop, uint32_t r)
{
    // bitwise (not short-circuit) boolean algebra keeps this branch-free
    const bool aib = na == b, bia = nb == a, ast = na == a, bst = nb == b;
    const bool stay = (aib & (bia | bst)) | (bia & ast);
    const bool c2 = (aib & b_noop) | (bia & a_noop);
    const bool four = (na == nb) & !stay;
    const bool rhi = (r & 2u) != 0;
    const bool move_a = !stay & (!four | rhi);          // SIM:352-356: slots 0,1 B moves; 2,3 A moves
    const bool move_b = !stay & (!four | !rhi);
    Resolved o;
    o.a = move_a ? na : a;
    o.b = move_b ? nb : b;
    const uint32_t rsel = four ? (r & 1u) : (r >> 1);
    const uint32_t pc = c2 ? (p ^ 1u) : rsel;
    o.p = (stay | four) ? pc : p;
    o.nlog2 = four ? 2u : ((stay & !c2) ? 1u : 0u);
    return o;
}
__device__ __forceinline__ Resolved resolve(const uint8_t* __restrict__ lut, uint32_t a, uint32_t b,
                                            uint32_t p, uint32_t ma, uint32_t mb, bool a_noop,
                                            bool b_noop, uint32_t r)
{
    const uint32_t p8 = p << 3;
    const uint32_t na = lut[(a << 4) + 8u - p8 + (ma & 7u)];   // A has the ball iff p == 0 (SIM:308)
    const uint32_t nb = lut[(b << 4) + p8 + (mb & 7u)];        // SIM:309
    return resolve_cand(na, nb, a, b, p, a_noop, b_noop, r);
}

// SIM:487-494 closed form (field-cell states only).
__device__ __forceinline__ int32_t obs_index(const PitchDev& P, uint32_t a, uint32_t b, uint32_t p)
{
    return 1 + 2 * ((int)a * P.Fm1 + (int)b - (b > a ? 1 : 0)) + (int)p;
}

// inverse of obs_index for 1 <= obs < nS
__device__ __forceinline__ uint32_t obs_to_packed(const PitchDev& P, int32_t obs)
{
    uint32_t idx = (uint32_t)(obs - 1);
    uint32_t p = idx & 1u, q = idx >> 1;
    uint32_t a = q / (uint32_t)P.Fm1;
    uint32_t rb = q - a * (uint32_t)P.Fm1;
    uint32_t b = rb + (rb >= a ? 1u : 0u);
    return a | (b << 8) | (p << 24);
}

// ---- Philox4x32-10 (Salmon et al. SC'11).  Contract v2: ONE call = the words of the 4 envs of an ALIGNED GROUP
// (global env id >> 2) at ONE step: counter = (group_lo, group_hi, step_lo, step_hi), key = seed, word index =
// env id & 3.  Every kernel here owns 4 consecutive envs per thread, so K1 and K2 alike pay one call per four
// env-steps (v1 keyed the counter by env and gave K1 one word of four: 40 multiplies per env-step).  Still a pure
// function of (seed, global env id, step): independent of GPU count, K and launch boundaries.
template <bool WIDE = false>
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t h0, l0, h1, l1;
        if (WIDE) {
            // one 32x32 -> 64 multiply per product (IMAD.WIDE.U32): fewer instructions, but measured SLOWER on B200
            // (profiles/r01g_time_k1_philox.log) -- A/B only
            const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
            h0 = (uint32_t)(p0 >> 32); l0 = (uint32_t)p0; h1 = (uint32_t)(p1 >> 32); l1 = (uint32_t)p1;
        } else {
            // mul.hi + mul.lo (both on the FMA pipe) rather than one wide multiply: keeps the ALU pipe,
            // which bounds the rollout kernels, for the 3-input XORs only
            h0 = __umulhi(0xD2511F53u, c0); l0 = 0xD2511F53u * c0;
            h1 = __umulhi(0xCD9E8D57u, c2); l1 = 0xCD9E8D57u * c2;
        }
        const uint32_t n0 = h1 ^ c1 ^ k0;
        const uint32_t n2 = h0 ^ c3 ^ k1;
        c1 = l1; c3 = l0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// The same with the ten round keys precomputed (kernel parameters -> constant bank operands of the XORs): the rollout
// kernels call it once per step, and 20 key additions per call are a third of the multiplies
struct PhiloxRoundKeys { uint32_t k0[10], k1[10]; };
inline PhiloxRoundKeys philox_round_keys(uint64_t seed)
{
    PhiloxRoundKeys r;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; ++i) { r.k0[i] = a; r.k1[i] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
    return r;
}
template <bool WIDE = false>
__device__ __forceinline__ void philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 const PhiloxRoundKeys& rk, uint32_t out[4])
{
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t h0, l0, h1, l1;
        if (WIDE) {
            const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
            h0 = (uint32_t)(p0 >> 32); l0 = (uint32_t)p0; h1 = (uint32_t)(p1 >> 32); l1 = (uint32_t)p1;
        } else {
            h0 = __umulhi(0xD2511F53u, c0); l0 = 0xD2511F53u * c0;
            h1 = __umulhi(0xCD9E8D57u, c2); l1 = 0xCD9E8D57u * c2;
        }
        const uint32_t n0 = h1 ^ c1 ^ rk.k0[i];
        const uint32_t n2 = h0 ^ c3 ^ rk.k1[i];
        c1 = l1; c3 = l0; c0 = n0; c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#ifndef SOCCER_K1_PHILOX_WIDE
#define SOCCER_K1_PHILOX_WIDE 0
#endif
// the four words of group `group` at `step`
template <bool WIDE = false>
__device__ __forceinline__ void philox_group(uint64_t seed, uint64_t group, uint64_t step, uint32_t out[4])
{
    philox4x32_10<WIDE>((uint32_t)group, (uint32_t)(group >> 32), (uint32_t)step, (uint32_t)(step >> 32),
                        (uint32_t)seed, (uint32_t)(seed >> 32), out);
}
// one env's word (the one-env-per-thread paths: generic kernel, scalar tails, unaligned env_id_base)
template <bool WIDE = false>
__device__ __forceinline__ uint32_t philox_word(uint64_t seed, uint64_t env_id, uint64_t step)
{
    uint32_t o[4];
    philox_group<WIDE>(seed, env_id >> 2, step, o);
    const uint32_t k = (uint32_t)env_id & 3u;
    return k == 0 ? o[0] : (k == 1 ? o[1] : (k == 2 ? o[2] : o[3]));
}

// Decode of one word w, read as the fixed-point number x = w / 2^32 in [0, 1):
//   joint action ja = floor(25 x) = mulhi(w, 25), aa = ja / 5, ab = ja % 5;
//   step draw    r32 = frac(25 x) * 2^32 = lo32(25 w): u = (r32 + 0.5) / 2^32, the rng32 format (25 is odd, so
//                w -> r32 is a bijection: exactly uniform).  slip_prob == 0 needs its top two bits only, and
//                jr = mulhi(w, 100) = ja * 4 + (r32 >> 30) is at once the column index of the step table;
//   reset draw   (w >> 2) & 3, i.e. w & 0xC is at once the byte offset into the start-observation words and the
//                reset field of an rng8 byte.
__device__ __forceinline__ uint32_t philox_jr(uint32_t w) { return __umulhi(w, 100u); }
__device__ __forceinline__ uint32_t philox_r32(uint32_t w) { return w * 25u; }
__device__ __forceinline__ uint32_t philox_ja(uint32_t w) { return __umulhi(w, 25u); }   // with philox_r32: one IMAD.WIDE
__device__ __forceinline__ void philox_actions(uint32_t w, uint32_t& aa, uint32_t& ab)
{
    const uint32_t ja = philox_jr(w) >> 2;
    aa = (ja * 52u) >> 8;      // ja / 5 for ja < 25
    ab = ja - aa * 5u;
}
// rng8-compatible nibble: bits 0..1 step draw, bits 2..3 reset draw
__device__ __forceinline__ uint32_t philox_rng8(uint32_t w) { return (philox_jr(w) & 3u) | (w & 0xCu); }

// Philox draws of the 4 envs of a group as rng8-compatible bytes (soccer_step_philox / soccer_step_table_philox:
// K1 with on-device draws).  env_id_base is a multiple of 4 here (the dispatcher sends other bases to the
// one-env-per-thread kernel), so local group g is global group (env_id_base >> 2) + g.
struct PhiloxKey { uint64_t seed, step, env_id_base; };
__device__ __forceinline__ void philox_words4(const PhiloxKey& k, int64_t g, uint32_t w[4])
{
    philox_group<SOCCER_K1_PHILOX_WIDE != 0>(k.seed, (k.env_id_base >> 2) + (uint64_t)g, k.step, w);
}
__device__ __forceinline__ uint32_t philox_rng8x4(const PhiloxKey& k, int64_t g)
{
    uint32_t w[4];
    philox_words4(k, g, w);
    return philox_rng8(w[0]) | (philox_rng8(w[1]) << 8) | (philox_rng8(w[2]) << 16) | (philox_rng8(w[3]) << 24);
}

// ---- programmatic dependent launch (no-ops unless the launch carries the PDL attribute) ----
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- streaming accessors and the 4-env group shared by the K1 kernels ----
constexpr int kThreads = 256;
__device__ __forceinline__ uint4 ld_stream(const uint4* p) { return __ldcs(p); }
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(uint4* p, uint4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(uint32_t* p, uint32_t v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(uint2* p, uint2 v) { __stcs(p, v); }
// The state word is the one stream that is re-read (by the next step).  Marking its lines L2::evict_last
// (-DSOCCER_STATE_EVICT_LAST=1) was meant to keep the state tensor resident in the 126 MB L2; measured on
// B200 it LOSES to the default replacement policy -- 2^24 envs (67 MB of state): 308 vs 329 G env-steps/s for
// the table kernel, 296 vs 310 G for the rules kernel; equal at 2^22 and 2^26 (profiles/r01g_ab_evict_last.log)
// -- so plain accesses are the default.
#ifndef SOCCER_STATE_EVICT_LAST
#define SOCCER_STATE_EVICT_LAST 0
#endif
__device__ __forceinline__ uint64_t keep_policy()
{
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));   // hoisted / CSE'd by the compiler
    return pol;
}
__device__ __forceinline__ uint4 ld_keep(const uint4* p)
{
#if SOCCER_STATE_EVICT_LAST
    uint4 v;
    asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(keep_policy()));
    return v;
#else
    return *p;
#endif
}
__device__ __forceinline__ void st_keep(uint4* p, uint4 v)
{
#if SOCCER_STATE_EVICT_LAST
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(keep_policy()) : "memory");
#else
    *p = v;
#endif
}

// One thread owns groups of 4 consecutive envs: 128-bit accesses on the 32-bit streams (state,
// obs, reward), 32-bit accesses on the byte streams (actions, draws, flags); a warp therefore
// touches 512 B / 128 B contiguous per instruction.
// A/B switch: prefetch.global.L2 of the groups SOCCER_K1_L2_PREFETCH iterations ahead of the register prefetch
// (deepens the load pipeline without registers); 0 = off.  Measured on B200: 1 and 2 both LOSE (2^24 envs 325 -> 290 G
// env-steps/s, 2^26 280 -> 276 / 273 G; profiles/r01g_ab_k1_l2prefetch.log), so it stays off.
#ifndef SOCCER_K1_L2_PREFETCH
#define SOCCER_K1_L2_PREFETCH 0
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_group(const uint4* st, const uint32_t* aa, const uint32_t* ab, const uint32_t* rg,
                                               int64_t g, bool with_rng)
{
    prefetch_l2(st + g);
    if ((threadIdx.x & 31) == 0) {          // one 128-byte line per warp and byte stream
        prefetch_l2(aa + g); prefetch_l2(ab + g);
        if (with_rng) prefetch_l2(rg + g);
    }
}
struct Group4 { uint4 s; uint32_t a, b, r; };
__device__ __forceinline__ Group4 load_group(const uint4* st, const uint32_t* aa, const uint32_t* ab,
                                             const uint32_t* rg, int64_t g)
{
    Group4 x;
    x.s = ld_keep(st + g);    // state is re-read by the next step: keep it in L2
    x.a = ld_stream(aa + g);
    x.b = ld_stream(ab + g);
    x.r = ld_stream(rg + g);
    return x;
}

// ---- one env-step given the chosen outcome: terminal detection, reward, obs, bookkeeping ----
struct StepOut {
    uint32_t state;     // next packed state (post-reset when auto-reset fired)
    int32_t  obs;       // what the reference's step() returned
    float    reward;
    uint32_t flags;
    int32_t  reset_obs;
};

template <bool AUTO_RESET, bool DETAIL = true>
__device__ __forceinline__ StepOut finish_step(const PitchDev& P, Resolved o, uint32_t t, uint32_t combo,
                                               uint32_t reset_sel, bool flip_reward)
{
    StepOut out;
    const uint32_t hc = o.p ? o.b : o.a;                    // cell of the ball holder
    const bool done = (hc & kGoalBit) != 0;                 // SIM:237-238 (goal states, SIM:91-103)
    // +1 in the right-hand goal column (A scores / B own goal), -1 in the left (SIM:94-102);
    // flipped for a single-agent player_b env (SIM:243-244; compared by value, -0.0 == 0.0)
    const bool plus = ((hc & kRightBit) != 0) != flip_reward;
    const float rew = done ? (plus ? 1.0f : -1.0f) : 0.0f;
    const int32_t q = (int)o.a * P.Fm1 + (int)o.b;          // SIM:487-494 closed form
    const int32_t obs_live = 2 * q + (int)o.p + (o.b > o.a ? -1 : 1);
    const int32_t obs = done ? 0 : obs_live;                // SIM:493
    const uint32_t t1 = t + 1;                              // SIM:399
    const bool trunc = t1 >= (uint32_t)kMaxT;               // SIM:404
    const bool reset = done | trunc;                        // SIM:406
    const uint32_t st = __byte_perm(__byte_perm(o.a, o.b, 0x4440), t1 | (o.p << 8), 0x5410);
    out.obs = obs;
    out.reward = rew;
    out.flags = (done ? 1u : 0u) | (trunc ? 2u : 0u);
    if (DETAIL) out.flags |= (o.nlog2 << 2) | (combo << 4);
    if (AUTO_RESET) {
        // SIM:410-424 fused: isd index = floor(n_isd * u) for the injected 2-bit draw
        const uint32_t m1 = (reset_sel & 1u) ? 0xFFFFFFFFu : 0u, m2 = (reset_sel & 2u) ? 0xFFFFFFFFu : 0u;
        const uint32_t rs = P.isd_state[0] + (m1 & P.isd_d1) + (m2 & P.isd_d2);
        const int32_t ro = P.isd_obs[0] + (int32_t)(m1 & (uint32_t)P.obs_d1) + (int32_t)(m2 & (uint32_t)P.obs_d2);
        out.state = reset ? rs : st;
        out.reset_obs = reset ? ro : obs;
    } else {
        out.state = st | (reset ? kNeedsReset : 0u);
        out.reset_obs = obs;
    }
    return out;
}

// slip_prob == 0 step of one env (the hot path)
template <bool AUTO_RESET, bool DETAIL = true>
__device__ __forceinline__ StepOut step_noslip(const PitchDev& P, const uint8_t* __restrict__ lut,
                                               uint32_t s, uint32_t aa, uint32_t ab, uint32_t rng,
                                               bool flip_reward)
{
    const uint32_t a = s & 0xFFu, b = (s >> 8) & 0xFFu, t = (s >> 16) & 0xFFu, p = (s >> 24) & 1u;
    const Resolved o = resolve(lut, a, b, p, aa, ab, aa == 0, ab == 0, rng & 3u);
    return finish_step<AUTO_RESET, DETAIL>(P, o, t, 0u, (rng >> 2) & 3u, flip_reward);
}

// ---- slip_prob > 0 (SIM:203-256)
// The reference lists, for the 9 slip combinations in the order of SIM:209-223, the outcomes of the slipped move
// pair with probability mp * nsp and draws with gym's categorical_sample: the first entry whose running fp64 sum
// exceeds u (argmax(cumsum > u), all-False -> 0).  Shared by the rules kernels here and the table kernels
// (soccer_table.cuh): a 9 x 3 table of the products mp_k * {1, 0.5, 0.25} (16-byte stride, built per CTA with
// __dmul_rn like SIM:241) and the first combination with a non-zero probability (SIM:226-227 skips the others;
// adding their 0.0 leaves the sums unchanged, so they can never become the pick).
constexpr int kPrtDoubles = 9 * 3 * 2;
struct SlipCtx { uint32_t prt; uint32_t first_k; };     // shared-window address of the table; first combination with mp != 0
__device__ __forceinline__ void slip_build_prt(double* prt, const PitchDev& P)
{
    if (threadIdx.x < 27) {
        const int k = threadIdx.x / 3, j = threadIdx.x % 3;
        prt[2 * threadIdx.x] = __dmul_rn(P.mp[k], j == 2 ? 0.25 : (j == 1 ? 0.5 : 1.0));     // SIM:241
        prt[2 * threadIdx.x + 1] = 0.0;
    }
}
__device__ __forceinline__ uint32_t slip_first_k(const PitchDev& P)
{
    uint32_t f = 0;
    for (int k = 8; k >= 0; --k) if (P.mp[k] != 0.0) f = (uint32_t)k;
    return f;
}
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
// Relaxed (non-volatile) shared-memory loads for the look-up chains of the table slip steppers.  `asm volatile`
// statements keep their program order, which serialises the dependent chains of the 4 envs of a thread (env 1's first
// load cannot issue before env 0's last one, which waits for env 0's earlier loads): ILP 1.  Without `volatile` the
// compiler interleaves them -- and could also hoist them above the barrier / mbarrier wait that publishes the data,
// so every base address they use is passed through launder() AFTER that point: a value the compiler cannot see
// through, hence no load that depends on it can move above it.
#ifndef SOCCER_LDS_RELAXED
#define SOCCER_LDS_RELAXED 1
#endif
__device__ __forceinline__ void launder(uint32_t& x) { asm volatile("" : "+r"(x) :: "memory"); }
__device__ __forceinline__ double lds_f64_r(uint32_t addr)
{
    double v;
#if SOCCER_LDS_RELAXED
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
#else
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
#endif
    return v;
}
__device__ __forceinline__ uint32_t lds_u8_r(uint32_t addr)
{
    uint32_t v;
#if SOCCER_LDS_RELAXED
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
#else
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
#endif
    return v;
}
__device__ __forceinline__ int32_t lds_s16_r(uint32_t addr)
{
    int32_t v;
#if SOCCER_LDS_RELAXED
    asm("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr));
#else
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr));
#endif
    return v;
}

// slip step of one env by the rules.  The sums at the end of each combination are non-decreasing, so "first
// entry whose running sum exceeds u" needs no found flag: the candidate moves on to combination k + 1 exactly
// while E_k <= u.  The outcome COUNT of a combination needs only the collision predicates (6 candidate look-ups
// and 12 compares shared by the 9 combinations, ~7 logic ops each); the full resolution runs once, for the pick.
// Sums use __dadd_rn in the reference's order (no FMA contraction): bit-identical to numpy's cumsum.
template <bool AUTO_RESET>
__device__ __forceinline__ StepOut step_slip(const PitchDev& P, const uint8_t* __restrict__ lut, const SlipCtx& sc,
                                             uint32_t s, uint32_t aa, uint32_t ab, double u,
                                             uint32_t reset_sel, bool flip_reward)
{
    const uint32_t a = s & 0xFFu, b = (s >> 8) & 0xFFu, t = (s >> 16) & 0xFFu, p = (s >> 24) & 1u;
    const uint32_t ma[3] = { aa, slip_move(aa, 0), slip_move(aa, 1) };
    const uint32_t mb[3] = { ab, slip_move(ab, 0), slip_move(ab, 1) };
    const bool a_noop = aa == 0, b_noop = ab == 0;
    const uint32_t p8 = p << 3;
    uint32_t na[3], nb[3];
    bool aib[3], ast[3], bia[3], bst[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        na[i] = lut[(a << 4) + 8u - p8 + (ma[i] & 7u)];       // SIM:308
        nb[i] = lut[(b << 4) + p8 + (mb[i] & 7u)];            // SIM:309
        aib[i] = na[i] == b; ast[i] = na[i] == a; bia[i] = nb[i] == a; bst[i] = nb[i] == b;
    }
    double E = 0.0;
    bool le = true;                     // E_{k-1} <= u: the pick is not before combination k
    uint32_t pick = 0;                  // combination | draw value << 4
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int i = combo_a(k), j = combo_b(k);
        const bool stay = (aib[i] & (bia[j] | bst[j])) | (bia[j] & ast[i]);
        const bool c2 = (aib[i] & b_noop) | (bia[j] & a_noop);
        const bool four = (na[i] == nb[j]) & !stay;
        const uint32_t nl16 = four ? 32u : ((stay & !c2) ? 16u : 0u);           // 16 * log2(#outcomes)
        const double pr = lds_f64(sc.prt + k * 48 + nl16);
        E = __dadd_rn(E, pr);
        uint32_t cand = (uint32_t)k;
        if (nl16) {                                                             // rare: 2 or 4 outcomes
            uint32_t slot = E <= u ? 1u : 0u;
            E = __dadd_rn(E, pr);
            if (nl16 == 32u) {
                slot += E <= u ? 1u : 0u; E = __dadd_rn(E, pr);
                slot += E <= u ? 1u : 0u; E = __dadd_rn(E, pr);
                cand |= slot << 4;                                              // 4-way: draw value r = slot
            } else {
                cand |= slot << 5;                                              // 2-way: r = 2 * slot
            }
        }
        pick = le ? cand : pick;
        le = E <= u;
    }
    if (le) pick = sc.first_k;                                                  // all-False -> index 0 of the list
    const uint32_t pk = pick & 15u;
    const uint32_t ca = (uint32_t)(0x221121000ull >> (pk * 4)) & 3u, cb = (uint32_t)(0x212100210ull >> (pk * 4)) & 3u;
    const uint32_t mas = ca == 0 ? ma[0] : (ca == 1 ? ma[1] : ma[2]);
    const uint32_t mbs = cb == 0 ? mb[0] : (cb == 1 ? mb[1] : mb[2]);
    const Resolved o = resolve(lut, a, b, p, mas, mbs, a_noop, b_noop, pick >> 4);
    return finish_step<AUTO_RESET>(P, o, t, pk, reset_sel, flip_reward);
}

} // namespace soccer

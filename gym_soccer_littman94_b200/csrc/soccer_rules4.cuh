// soccer_rules4.cuh -- the slip_prob == 0 auto-reset step of FOUR envs at once, byte-parallel.
//
// Same rules as resolve()/finish_step() in soccer_rules.cuh (SIM:296-373, 235-240, 399-424), but
// the four envs a thread owns are kept as packed bytes -- A4 = (a0,a1,a2,a3), B4, T4, P4, the
// candidates NA4 / NB4 -- so that every cell comparison, the collision algebra, possession,
// terminal detection, truncation, the fused reset and the flags byte cost one integer instruction
// per FOUR envs instead of per env.  Booleans live in bit 7 of each byte ("flags", mask 0x80808080);
// mask4() widens them to 0xFF byte masks for the selects.  Only the two candidate-table lookups,
// the observation index (16-bit lanes) and the float reward are per env.
// Measured effect (B200, 2^24 envs, 5x4): see DESIGN.md section 7.
// SIM = /root/reference/gym_soccer/envs/soccer_simultaneous_env.py.
#pragma once
#include "soccer_rules.cuh"

namespace soccer {

constexpr uint32_t kH = 0x80808080u, kL = 0x01010101u;

// per byte: x == y  ->  0x80 (any 8-bit values)
__device__ __forceinline__ uint32_t eq4(uint32_t x, uint32_t y)
{
    const uint32_t z = x ^ y;
    return ~(((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z) & kH;
}
// 0x80 flags -> 0xFF byte masks
__device__ __forceinline__ uint32_t mask4(uint32_t f)
{
    uint32_t m;                                 // PRMT with the sign-replicate bit of every selector nibble set: ONE instruction
    asm("prmt.b32 %0, %1, %2, 0xBA98;" : "=r"(m) : "r"(f), "r"(0u));
    return m;
}
// per byte: m ? x : y
__device__ __forceinline__ uint32_t sel4(uint32_t m, uint32_t x, uint32_t y) { return (x & m) | (y & ~m); }
__device__ __forceinline__ uint32_t pack4(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3)
{
    return __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
}
__device__ __forceinline__ uint32_t byte_of(uint32_t x, int e) { return __byte_perm(x, 0, 0x4440 + e); }

// Start-state constants broadcast to every byte / 16-bit lane.  isd[r] = isd[0] ^ (r&1 ? d1 : 0) ^
// (r&2 ? d2 : 0) bytewise (checked on the host), isd_obs[r] = o0 + (r&1)*od1 + (r>>1)*od2.
struct Isd4 {
    uint32_t a0, da1, da2, b0, db1, db2, p0, dp1, dp2;   // bytes x 4
    uint32_t o0, od1, od2;                                // 16-bit lanes x 2
};
__device__ __forceinline__ Isd4 make_isd4(const PitchDev& P)
{
    Isd4 I;
    const uint32_t s0 = P.isd_state[0], x1 = s0 ^ P.isd_state[1], x2 = s0 ^ P.isd_state[2];
    I.a0 = (s0 & 0xFFu) * kL;          I.da1 = (x1 & 0xFFu) * kL;          I.da2 = (x2 & 0xFFu) * kL;
    I.b0 = ((s0 >> 8) & 0xFFu) * kL;   I.db1 = ((x1 >> 8) & 0xFFu) * kL;   I.db2 = ((x2 >> 8) & 0xFFu) * kL;
    I.p0 = ((s0 >> 24) & 1u) * kL;     I.dp1 = ((x1 >> 24) & 1u) * kL;     I.dp2 = ((x2 >> 24) & 1u) * kL;
    I.o0 = (uint32_t)P.isd_obs[0] * 0x00010001u;
    I.od1 = (uint32_t)P.obs_d1 & 0xFFFFu;     // multiplied into 16-bit lanes below (mod 2^16 per lane)
    I.od2 = (uint32_t)P.obs_d2 & 0xFFFFu;
    return I;
}

struct Step4 {
    uint32_t s[4];      // next packed CELL-layout states (post-reset where a reset fired)
    uint32_t obs[4];    // what step() returned (0 on a goal)
    uint32_t rew[4];    // float bits
    uint32_t rew4;      // the four rewards as int8 bytes
    uint32_t flags4;    // four flags bytes
    uint32_t robs[4];   // observation the next step starts from (only when RESET_OBS)
    int32_t  rew_sum;   // sum of the four rewards (+1 / -1 / 0), for K2's statistics
};

// The four envs of a thread as packed bytes: cells of A and B, timestep, possession (0 / 1).  K2 and the fused replay keep
// the state in this form for all their steps (the 4x4 byte transposes are paid once per launch, not once per step).
struct Soa4 { uint32_t A, B, T, P; };
__device__ __forceinline__ Soa4 soa4_from_words(const uint32_t sv[4])
{
    // 4x4 byte transpose: words (a,b,t,p) per env -> A4, B4, T4, P4
    const uint32_t u0 = __byte_perm(sv[0], sv[1], 0x5140), u1 = __byte_perm(sv[2], sv[3], 0x5140);
    const uint32_t v0 = __byte_perm(sv[0], sv[1], 0x7362), v1 = __byte_perm(sv[2], sv[3], 0x7362);
    Soa4 x;
    x.A = __byte_perm(u0, u1, 0x5410); x.B = __byte_perm(u0, u1, 0x7632);
    x.T = __byte_perm(v0, v1, 0x5410); x.P = __byte_perm(v0, v1, 0x7632) & kL;
    return x;
}
__device__ __forceinline__ void soa4_to_words(const Soa4& x, uint32_t s[4])
{
    const uint32_t w0 = __byte_perm(x.A, x.B, 0x5140), w1 = __byte_perm(x.A, x.B, 0x7362);
    const uint32_t x0 = __byte_perm(x.T, x.P, 0x5140), x1 = __byte_perm(x.T, x.P, 0x7362);
    s[0] = __byte_perm(w0, x0, 0x5410); s[1] = __byte_perm(w0, x0, 0x7632);
    s[2] = __byte_perm(w1, x1, 0x5410); s[3] = __byte_perm(w1, x1, 0x7632);
}
// one env's word out of / into the packed form (the slip steppers' never-taken walk)
__device__ __forceinline__ uint32_t soa4_word(const Soa4& x, int e)
{
    return byte_of(x.A, e) | (byte_of(x.B, e) << 8) | (byte_of(x.T, e) << 16) | (byte_of(x.P, e) << 24);
}
__device__ __forceinline__ void soa4_set_word(Soa4& x, int e, uint32_t w)
{
    const uint32_t m = 0xFFu << (8 * e);
    x.A = (x.A & ~m) | ((w & 0xFFu) << (8 * e));          x.B = (x.B & ~m) | (((w >> 8) & 0xFFu) << (8 * e));
    x.T = (x.T & ~m) | (((w >> 16) & 0xFFu) << (8 * e));  x.P = (x.P & ~m) | (((w >> 24) & 1u) << (8 * e));
}

// in / out: the four states; MA4 / MB4: action bytes (only bits 0..2 of each byte are read); R4: step draw in bits 0..1 of
// each byte, RST4: reset draw in bits 2..3 (the other bits of R4 / RST4 are ignored: an rng8 word serves as both).
// R2: the draw of a 2-outcome collision in bit 1 of each byte; the same as R4 when one 2-bit draw serves both kinds
// (slip_prob == 0: outcome r of 4, r >> 1 of 2), a separate word for the slip steppers, whose slot inside a 2-way and
// inside a 4-way combination comes from different thresholds.  o.s is NOT written here (step4_noslip does).
template <bool RESET_OBS>
__device__ __forceinline__ void step4_core(const PitchDev& P, const Isd4& I, const uint8_t* __restrict__ lut,
                                           const Soa4& in, uint32_t MA4, uint32_t MB4, uint32_t R4, Step4& o,
                                           uint32_t R2, uint32_t RST4, Soa4& out)
{
    const uint32_t A4 = in.A, B4 = in.B, T4 = in.T, P4 = in.P;

    // ---- candidates (SIM:308-309, 364-373): cand[cell*16 + has_ball*8 + move], two lookups per env
    const uint32_t MA = MA4 & 0x07070707u, MB = MB4 & 0x07070707u;
    const uint32_t HMA = ((P4 ^ kL) << 3) | MA;         // A has the ball iff p == 0
    const uint32_t HMB = (P4 << 3) | MB;
    uint32_t na[4], nb[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        na[e] = lut[byte_of(A4, e) * 16u + byte_of(HMA, e)];
        nb[e] = lut[byte_of(B4, e) * 16u + byte_of(HMB, e)];
    }
    const uint32_t NA = pack4(na[0], na[1], na[2], na[3]), NB = pack4(nb[0], nb[1], nb[2], nb[3]);

    // ---- collision algebra (SIM:315-360; derivation in DESIGN.md section 3), flags in bit 7
    const uint32_t aib = eq4(NA, B4), bia = eq4(NB, A4), ast = eq4(NA, A4), bst = eq4(NB, B4), same = eq4(NA, NB);
    const uint32_t anoop = eq4(MA, 0u), bnoop = eq4(MB, 0u);
    const uint32_t stay = (aib & (bia | bst)) | (bia & ast);
    const uint32_t c2 = (aib & bnoop) | (bia & anoop);
    const uint32_t four = same & ~stay;
    const uint32_t rhi = (R4 << 6) & kH, rlo = (R4 << 7) & kH;      // draw bit 1 / bit 0
    const uint32_t move_a = ~stay & (~four | rhi) & kH;             // SIM:352-356: slots 0,1 B moves; 2,3 A moves
    const uint32_t move_b = ~stay & (~four | ~rhi) & kH;
    const uint32_t FA = sel4(mask4(move_a), NA, A4), FB = sel4(mask4(move_b), NB, B4);
    const uint32_t pf = P4 << 7;
    const uint32_t rsel = (four & rlo) | (~four & ((R2 << 6) & kH));
    const uint32_t pc = (c2 & ~pf) | (~c2 & rsel);
    const uint32_t sf = stay | four;
    const uint32_t fpf = ((sf & pc) | (~sf & pf)) & kH;             // final possession flag
    const uint32_t FP = fpf >> 7;

    // ---- terminal / reward (SIM:235-240, 91-103), truncation (SIM:399-404)
    const uint32_t HC = sel4(mask4(fpf), FB, FA);                   // the ball holder's cell
    const uint32_t done = HC & kH;                                  // goal bit
    const uint32_t plus = (HC << 1) & kH;                           // right-hand goal -> +1
    const uint32_t mD = mask4(done);
    const uint32_t done1 = done >> 7;
    const uint32_t RW = (mD & ~mask4(plus)) | done1;                // signed bytes: +1 / -1 (0xFF) / 0
    const uint32_t T1 = T4 + kL;
    const uint32_t trunc = (T1 + 28u * kL) & kH;                    // t + 1 >= 100  (t <= 99 on entry)
    const uint32_t reset = done | trunc;                            // SIM:406
    o.flags4 = done1 | (trunc >> 6);

    // ---- fused reset (SIM:410-424): start state from the 2-bit reset draw
    const uint32_t m1 = mask4((RST4 << 5) & kH), m2 = mask4((RST4 << 4) & kH), mR = mask4(reset);
    const uint32_t An = I.a0 ^ (m1 & I.da1) ^ (m2 & I.da2);
    const uint32_t Bn = I.b0 ^ (m1 & I.db1) ^ (m2 & I.db2);
    const uint32_t Pn = I.p0 ^ (m1 & I.dp1) ^ (m2 & I.dp2);
    out.A = sel4(mR, An, FA); out.B = sel4(mR, Bn, FB); out.P = sel4(mR, Pn, FP); out.T = T1 & ~mR;

    // ---- observation index (SIM:487-494) in 16-bit lanes: envs (0,2) and (1,3)
    //      obs = 1 + 2*(a*(F-1) + b) + p - 2*(b > a); no lane ever borrows or overflows
    const uint32_t FA02 = FA & 0x00FF00FFu, FA13 = (FA >> 8) & 0x00FF00FFu;
    const uint32_t FB02 = FB & 0x00FF00FFu, FB13 = (FB >> 8) & 0x00FF00FFu;
    const uint32_t q02 = FA02 * (uint32_t)P.Fm1 + FB02, q13 = FA13 * (uint32_t)P.Fm1 + FB13;
    const uint32_t gt = ~((FA | kH) - (FB & 0x7F7F7F7Fu)) & kH;     // fb > fa
    const uint32_t lo02 = q02 * 2u + (FP & 0x00010001u) + 0x00010001u - ((gt >> 6) & 0x00020002u);
    const uint32_t lo13 = q13 * 2u + ((FP >> 8) & 0x00010001u) + 0x00010001u - ((gt >> 14) & 0x00020002u);
    const uint32_t ob02 = lo02 & ~__byte_perm(mD, 0, 0x2200), ob13 = lo13 & ~__byte_perm(mD, 0, 0x3311);   // SIM:493
    o.obs[0] = ob02 & 0xFFFFu; o.obs[2] = ob02 >> 16; o.obs[1] = ob13 & 0xFFFFu; o.obs[3] = ob13 >> 16;

    // ---- float reward per env, and its sum for the statistics
#pragma unroll
    for (int e = 0; e < 4; ++e)
        o.rew[e] = __float_as_uint((float)(int)(signed char)(RW >> (8 * e)));
    o.rew4 = RW;
    o.rew_sum = (int)__dp4a((int)RW, (int)kL, 0);

    if (RESET_OBS) {
        const uint32_t s1 = (RST4 >> 2) & kL, s2 = (RST4 >> 3) & kL;   // reset-draw bits as bytes 0/1
        const uint32_t n02 = I.o0 + (s1 & 0x00010001u) * I.od1 + (s2 & 0x00010001u) * I.od2;
        const uint32_t n13 = I.o0 + ((s1 >> 8) & 0x00010001u) * I.od1 + ((s2 >> 8) & 0x00010001u) * I.od2;
        const uint32_t r02 = __byte_perm(mR, 0, 0x2200), r13 = __byte_perm(mR, 0, 0x3311);
        const uint32_t z02 = sel4(r02, n02, ob02), z13 = sel4(r13, n13, ob13);
        o.robs[0] = z02 & 0xFFFFu; o.robs[2] = z02 >> 16; o.robs[1] = z13 & 0xFFFFu; o.robs[3] = z13 >> 16;
    }
}

// sv: the four state words (CELL layout); R4: rng8 bytes (bits 0..1 step draw, 2..3 reset draw)
template <bool RESET_OBS>
__device__ __forceinline__ void step4_noslip(const PitchDev& P, const Isd4& I, const uint8_t* __restrict__ lut,
                                             const uint32_t sv[4], uint32_t MA4, uint32_t MB4, uint32_t R4, Step4& o,
                                             uint32_t R2)
{
    Soa4 out;
    step4_core<RESET_OBS>(P, I, lut, soa4_from_words(sv), MA4, MB4, R4, o, R2, R4, out);
    soa4_to_words(out, o.s);
}
template <bool RESET_OBS>
__device__ __forceinline__ void step4_noslip(const PitchDev& P, const Isd4& I, const uint8_t* __restrict__ lut,
                                             const uint32_t sv[4], uint32_t MA4, uint32_t MB4, uint32_t R4, Step4& o)
{
    step4_noslip<RESET_OBS>(P, I, lut, sv, MA4, MB4, R4, o, R4);
}

} // namespace soccer

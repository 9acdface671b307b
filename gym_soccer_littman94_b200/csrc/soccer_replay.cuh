// soccer_replay.cuh -- T lock-step step() calls of n envs in ONE launch (soccer_step_many).
//
// The reference's replay loop is `for t in range(T): env.step(actions[t])` (SIM:375-408) with a reset()
// (SIM:410-424) wherever an episode ended.  When the joint actions and the injected draws of all T
// steps are known up front (parity replays such as BASELINE config 2, open-loop evaluation of a
// recorded action sequence), the state word never has to leave the registers: per env-step the kernel
// reads 3 bytes (act_a, act_b, rng8 of row t) and writes 9 (obs, reward, flags), i.e. 12 + 8/T bytes
// instead of K1's 20, and a 4096-env batch no longer pays one launch per step (1.7 us even inside a
// CUDA graph) but one dependent table look-up (~0.1 us).
//
// The inputs do not depend on the state, so they are fetched ahead of the step that consumes them
// through a register double buffer of U rows: U = 4 when the batch fills the machine (HBM-bound: 2^22
// envs x 64 steps run at 0.87 of the measured HBM peak), U = 8 when it does not (a step is then a
// dependent ~130-cycle chain per warp and the loads must be 8 rows ahead to hide the memory latency).
// An additional prefetch.global.L2 32 rows ahead was measured and did not pay (profiles/r01d_ab_replay*.log).
// Results are bit-identical to T calls of K1 (tests/test_gpu_parity.py).
// SIM = /root/reference/gym_soccer/envs/soccer_simultaneous_env.py.
#pragma once
#include "soccer_rules4.cuh"
#include "soccer_table.cuh"

namespace soccer {

struct ReplayArgs {
    uint32_t* state; const uint8_t* act_a; const uint8_t* act_b; const uint8_t* rng8; int32_t T;
    int32_t* obs; float* reward; uint8_t* flags; int32_t* reset_obs; int64_t n;
};

// inputs of one row for the VEC envs of a thread: 4 action / draw bytes in one word, or one byte
template <int VEC> __device__ __forceinline__ uint32_t ld_row(const uint8_t* p)
{
    if (VEC == 4) return __ldcs(reinterpret_cast<const uint32_t*>(p));
    return (uint32_t)__ldcs(p);
}

// ---- steppers: one step of the VEC envs of a thread given their action / draw bytes
struct ReplayTable {
    TblCtx c;
    template <int VEC> __device__ __forceinline__ void enter(uint32_t*) const {}
    template <int VEC> __device__ __forceinline__ void leave(uint32_t*) const {}
    template <int VEC, bool RO>
    __device__ __forceinline__ void step(uint32_t* s, uint32_t a4, uint32_t b4, uint32_t r4, uint32_t* oo, uint32_t* rr,
                                         uint32_t& fw, uint32_t* ro) const
    {
        // byte-parallel column index jr = aa*20 + ab*4 + (draw & 3) of all envs at once (as in K1)
        const uint32_t jr4 = a4 * 20u + b4 * 4u + (r4 & 0x03030303u);
        const uint32_t rs4 = r4 & 0x0C0C0C0Cu;
        uint32_t ff[4] = { 0, 0, 0, 0 };
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const TblOut o = table_step(c, s[e], byte_of(jr4, e), byte_of(rs4, e));
            s[e] = o.state; oo[e] = o.obs; rr[e] = __float_as_uint((float)o.rew_i); ff[e] = o.flags;
            if (RO) ro[e] = o.reset_obs;
        }
        fw = VEC == 4 ? pack4(ff[0], ff[1], ff[2], ff[3]) : ff[0];
    }
};

struct ReplayRules {
    const PitchDev& P; const uint8_t* lut; Isd4 I;
    // four envs per thread: the states stay packed bytes (Soa4 in s[0..3]) for all T steps
    template <int VEC> __device__ __forceinline__ void enter(uint32_t* s) const
    {
        if (VEC == 4) { const Soa4 x = soa4_from_words(s); s[0] = x.A; s[1] = x.B; s[2] = x.T; s[3] = x.P; }
    }
    template <int VEC> __device__ __forceinline__ void leave(uint32_t* s) const
    {
        if (VEC == 4) { const Soa4 x = { s[0], s[1], s[2], s[3] }; soa4_to_words(x, s); }
    }
    template <int VEC, bool RO>
    __device__ __forceinline__ void step(uint32_t* s, uint32_t a4, uint32_t b4, uint32_t r4, uint32_t* oo, uint32_t* rr,
                                         uint32_t& fw, uint32_t* ro) const
    {
        if (VEC == 4) {
            const Soa4 in = { s[0], s[1], s[2], s[3] };
            Soa4 out;
            Step4 o;
            step4_core<RO>(P, I, lut, in, a4, b4, r4, o, r4, r4, out);
            s[0] = out.A; s[1] = out.B; s[2] = out.T; s[3] = out.P;
#pragma unroll
            for (int e = 0; e < 4; ++e) { oo[e] = o.obs[e]; rr[e] = o.rew[e]; if (RO) ro[e] = o.robs[e]; }
            fw = o.flags4;
        } else {
            const StepOut o = step_noslip<true, false>(P, lut, s[0], a4 & 0xFFu, b4 & 0xFFu, r4 & 0xFFu, false);
            s[0] = o.state; oo[0] = (uint32_t)o.obs; rr[0] = __float_as_uint(o.reward); fw = o.flags & 3u;
            if (RO) ro[0] = (uint32_t)o.reset_obs;
        }
    }
};

template <int VEC, bool RO, int kU, class Stepper>
__device__ __forceinline__ void replay_body(const Stepper& S, const ReplayArgs& a)
{
    const int64_t n_groups = a.n / VEC;
    const int32_t T = a.T;
    const int64_t n = a.n;
    // slot order as in K2 (soccer_rollout.cuh): CTA-major full passes, warp round-robin remainder
    const int64_t n_slots = (n_groups + 31) >> 5;
    const int32_t wpc = blockDim.x >> 5, wib = threadIdx.x >> 5;
    const int64_t per_pass = (int64_t)gridDim.x * wpc;
    const int64_t full = n_slots / per_pass;
    for (int64_t pass = 0; pass <= full; ++pass) {
        const int64_t slot = pass < full ? (pass * gridDim.x + blockIdx.x) * wpc + wib
                                         : full * per_pass + (int64_t)wib * gridDim.x + blockIdx.x;
        if (slot >= n_slots) break;
        const int64_t g = slot * 32 + (threadIdx.x & 31);
        if (g >= n_groups) break;
        const int64_t i0 = g * VEC;
        uint32_t s[4] = { 0, 0, 0, 0 };
        if (VEC == 4) {
            const uint4 v = reinterpret_cast<const uint4*>(a.state)[g];
            s[0] = v.x; s[1] = v.y; s[2] = v.z; s[3] = v.w;
        } else {
            s[0] = a.state[i0];
        }
        S.template enter<VEC>(s);
        const uint8_t *pa = a.act_a + i0, *pb = a.act_b + i0, *pr = a.rng8 + i0;
        int32_t* op = a.obs + i0;
        float* rp = a.reward + i0;
        uint8_t* fp = a.flags + i0;
        int32_t* qp = RO ? a.reset_obs + i0 : nullptr;
        uint32_t ia[kU], ib[kU], ir[kU];
#pragma unroll
        for (int j = 0; j < kU; ++j) {
            const bool ok = j < T;
            ia[j] = ok ? ld_row<VEC>(pa + j * n) : 0u;
            ib[j] = ok ? ld_row<VEC>(pb + j * n) : 0u;
            ir[j] = ok ? ld_row<VEC>(pr + j * n) : 0u;
        }
        for (int32_t t0 = 0; t0 < T; t0 += kU) {
            pa += kU * n; pb += kU * n; pr += kU * n;              // rows t0 + kU ...
            uint32_t na[kU], nb[kU], nr[kU];
#pragma unroll
            for (int j = 0; j < kU; ++j) {
                const bool ok = t0 + kU + j < T;                    // warp-uniform
                na[j] = ok ? ld_row<VEC>(pa + j * n) : 0u;
                nb[j] = ok ? ld_row<VEC>(pb + j * n) : 0u;
                nr[j] = ok ? ld_row<VEC>(pr + j * n) : 0u;
            }
#pragma unroll
            for (int j = 0; j < kU; ++j) {
                if (t0 + j >= T) continue;                          // warp-uniform
                uint32_t oo[4], rr[4], ro[4], fw;
                S.template step<VEC, RO>(s, ia[j], ib[j], ir[j], oo, rr, fw, ro);
                if (VEC == 4) {
                    st_stream(reinterpret_cast<uint4*>(op), make_uint4(oo[0], oo[1], oo[2], oo[3]));
                    st_stream(reinterpret_cast<uint4*>(rp), make_uint4(rr[0], rr[1], rr[2], rr[3]));
                    st_stream(reinterpret_cast<uint32_t*>(fp), fw);
                    if (RO) st_stream(reinterpret_cast<uint4*>(qp), make_uint4(ro[0], ro[1], ro[2], ro[3]));
                } else {
                    *op = (int32_t)oo[0]; *rp = __uint_as_float(rr[0]); *fp = (uint8_t)fw;
                    if (RO) *qp = (int32_t)ro[0];
                }
                op += n; rp += n; fp += n;
                if (RO) qp += n;
            }
#pragma unroll
            for (int j = 0; j < kU; ++j) { ia[j] = na[j]; ib[j] = nb[j]; ir[j] = nr[j]; }
        }
        S.template leave<VEC>(s);
        if (VEC == 4) reinterpret_cast<uint4*>(a.state)[g] = make_uint4(s[0], s[1], s[2], s[3]);
        else a.state[i0] = s[0];
    }
}

template <int VEC, bool RO, int U>
__global__ void __launch_bounds__(kRolloutThreads, 1)
k_replay_table(const PitchDev P, const uint16_t* __restrict__ gtable, uint32_t table_bytes, const ReplayArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    stage_table(smem_raw, gtable, table_bytes, &bar, P);
    ReplayTable S;
    S.c = make_ctx(smem_raw, table_bytes, P);
    wait_table(&bar);
    replay_body<VEC, RO, U>(S, a);
}

template <int VEC, bool RO, int U>
__global__ void __launch_bounds__(kThreads)
k_replay(const PitchDev P, const ReplayArgs a)
{
    __shared__ __align__(16) uint8_t lut[kLutBytes];
    build_cand_lut(lut, P);
    __syncthreads();
    const ReplayRules S = { P, lut, make_isd4(P) };
    replay_body<VEC, RO, U>(S, a);
}

} // namespace soccer

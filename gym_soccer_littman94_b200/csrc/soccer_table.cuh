// soccer_table.cuh -- K1 / K2 with the whole transition table resident in shared memory.
//
// The table is the reference's P (SIM:167-293) for slip_prob == 0, produced ON THE DEVICE by the
// rules path (k_build_step_table -> resolve()/finish_step()), compacted to one int16 per
// (state, joint action, 2-bit draw):
//     table[obs*100 + (aa*5+ab)*4 + r] = next_obs | nlog2 << 10 | (reward & 3) << 14
// so that a SIGN-EXTENDING 16-bit shared-memory load yields next_obs = e & 0x3FF and
// reward = e >> 14 (arithmetic: -1, 0, +1) with no further decoding.  761 * 100 * 2 = 152,200 bytes
// for the 5x4 pitch: it fits the 227 KB of one SM once, hence ONE persistent 1024-thread CTA per
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
constexpr int kMaxTableStates = 1023;        // next_obs must fit 10 bits
constexpr uint32_t kTblObsMask = 0x3FFu;
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
        table[i] = (uint16_t)((uint32_t)f.obs | (o.nlog2 << 10) | (rew2 << 14));
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
struct TblCtx { uint32_t tbl, isd, last; };    // shared-window addresses of the table / the isd words; last entry index
struct TblOut { uint32_t state, obs, flags; int32_t rew_i; uint32_t reset_obs; };
__device__ __forceinline__ TblOut table_step(const TblCtx& c, uint32_t s, uint32_t jr, uint32_t rsel4)
{
    // the clamp keeps a corrupt state word / action byte inside the table
    const uint32_t idx = min((s & 0xFFFFu) * 100u + jr, c.last);
    int32_t e;
    uint32_t ro;
    // volatile: ordered after the (volatile) mbarrier wait that publishes the table
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(e) : "r"(c.tbl + idx * 2u));          // sign-extending LDS.S16
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

__device__ __forceinline__ TblCtx make_ctx(const uint8_t* smem, uint32_t table_bytes, const PitchDev& P)
{
    TblCtx c;
    c.tbl = smem_u32(smem);
    c.isd = c.tbl + table_bytes;
    c.last = (uint32_t)P.nS * 100u - 1u;
    return c;
}

template <bool RESET_OBS>
__device__ __forceinline__ void table_step_group(const TblCtx& c, const Group4& x, int64_t g, uint4* st, uint4* obs,
                                                 uint4* rew, uint32_t* flg, uint4* rob)
{
    // byte-parallel: jr = aa*20 + ab*4 + (rng & 3) for all four envs in three instructions
    const uint32_t jr4 = x.a * 20u + x.b * 4u + (x.r & 0x03030303u);
    const uint32_t rs4 = x.r & 0x0C0C0C0Cu;
    const uint32_t sv[4] = { x.s.x, x.s.y, x.s.z, x.s.w };
    uint32_t so[4], oo[4], ro[4], rr[4], ff[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const TblOut o = table_step(c, sv[e], __byte_perm(jr4, 0, 0x4440 + e), __byte_perm(rs4, 0, 0x4440 + e));
        so[e] = o.state; oo[e] = o.obs; rr[e] = __float_as_uint((float)o.rew_i); ro[e] = o.reset_obs; ff[e] = o.flags;
    }
    st_keep(st + g, make_uint4(so[0], so[1], so[2], so[3]));
    st_stream(obs + g, make_uint4(oo[0], oo[1], oo[2], oo[3]));
    st_stream(rew + g, make_uint4(rr[0], rr[1], rr[2], rr[3]));
    st_stream(flg + g, __byte_perm(__byte_perm(ff[0], ff[1], 0x0040), __byte_perm(ff[2], ff[3], 0x0040), 0x5410));
    if (RESET_OBS) st_stream(rob + g, make_uint4(ro[0], ro[1], ro[2], ro[3]));
}

// K1, table variant: persistent, one 1024-thread CTA per SM.
template <bool RESET_OBS>
__global__ void __launch_bounds__(kTableThreads, 1)
k_step_table(const PitchDev P, const uint16_t* __restrict__ gtable, uint32_t table_bytes,
             uint32_t* __restrict__ state, const uint8_t* __restrict__ act_a, const uint8_t* __restrict__ act_b,
             const uint8_t* __restrict__ rng, int32_t* __restrict__ obs, float* __restrict__ reward,
             uint8_t* __restrict__ flags, int32_t* __restrict__ reset_obs, int64_t n_groups)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    pdl_launch_dependents();
    stage_table(smem_raw, gtable, table_bytes, &bar, P);     // the table is constant: safe before pdl_wait
    const TblCtx c = make_ctx(smem_raw, table_bytes, P);
    pdl_wait();                  // the table fill overlaps the previous kernel's tail

    uint4* st4 = reinterpret_cast<uint4*>(state);
    const uint32_t* a4 = reinterpret_cast<const uint32_t*>(act_a);
    const uint32_t* b4 = reinterpret_cast<const uint32_t*>(act_b);
    const uint32_t* r4 = reinterpret_cast<const uint32_t*>(rng);
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
        table_step_group<RESET_OBS>(c, x0, g, st4, o4, w4, f4, q4);
        if (two) table_step_group<RESET_OBS>(c, x1, g + stride, st4, o4, w4, f4, q4);
        x0 = y0; x1 = y1; g = gn; one = n_one; two = n_two;
    }
}

// scalar tail / misaligned fallback of the table path (global-memory table, one env per thread)
__global__ void __launch_bounds__(kThreads)
k_step_table_scalar(const PitchDev P, const uint16_t* __restrict__ gtable, uint32_t* __restrict__ state,
                    const uint8_t* __restrict__ act_a, const uint8_t* __restrict__ act_b,
                    const uint8_t* __restrict__ rng, int32_t* __restrict__ obs, float* __restrict__ reward,
                    uint8_t* __restrict__ flags, int32_t* __restrict__ reset_obs, int64_t n)
{
    const int16_t* tbl = reinterpret_cast<const int16_t*>(gtable);
    const uint32_t last = (uint32_t)P.nS * 100u - 1u;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t s = state[i], rg = rng[i];
        const uint32_t idx = min((s & 0xFFFFu) * 100u + (uint32_t)act_a[i] * 20u + (uint32_t)act_b[i] * 4u + (rg & 3u), last);
        const int32_t e = tbl[idx];
        const uint32_t nobs = (uint32_t)e & kTblObsMask;
        const bool done = nobs == 0;
        const uint32_t s1 = s + 0x10000u;
        const bool trunc = s1 >= kTruncWord, reset = done | trunc;
        const uint32_t ro = (uint32_t)P.isd_obs[(rg >> 2) & 3u];
        state[i] = reset ? ro : ((s1 & 0xFFFF0000u) | nobs);
        obs[i] = (int32_t)nobs;
        reward[i] = (float)(e >> 14);
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

/*
 * soccer_b200.h -- C ABI of libsoccer_b200.so: the Littman'94 soccer step/reset hot path
 * as hand-written sm_100a CUDA kernels.
 *
 * The reference (mimoralea/gym-soccer-littman94) is pure Python and has no FFI; the
 * boundary this library sits under is the Python class surface of
 *   SIM = /root/reference/gym_soccer/envs/soccer_simultaneous_env.py
 * Each entry point cites the reference code it replaces.  INTEGRATION.md shows the ctypes
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *   - Every pointer named in a kernel call is a DEVICE pointer into caller-owned memory
 *     (e.g. a torch tensor's data_ptr()).  Kernel entry points never allocate, keep no global
 *     state, and are re-entrant across streams.  The one exception is the pinned-host-memory pair
 *     soccer_host_alloc / soccer_host_free: it allocates (mmap + cudaHostRegister, or cudaHostAlloc)
 *     and keeps a mutex-protected set of the allocations that took the cudaHostAlloc fallback.
 *     Entry points ending in _host() are pure host helpers and need no GPU.
 *   - Every kernel call is asynchronous on `stream` (a cudaStream_t / CUstream).
 *   - Return value: 0 = OK, negative = argument error (SOCCER_E*), positive = cudaError_t.
 *   - There is no CPU fallback: without a CUDA device kernel calls return a cudaError_t.
 *
 * Packed environment state (one uint32 per env, structure-of-arrays):
 *   bits  0..7   cell code of player A      bits 16..23  timestep (0..100, SIM:399)
 *   bits  8..15  cell code of player B      bit  24      possession p (0 = A has the ball)
 *                                           bit  25      needs_reset (SIM:406; only ever set
 *                                                        when auto_reset == 0)
 *   cell code: field cell = row * width + (col - 1), col being the reference's padded column
 *   1..width; goal cell = 0x80 | (right_goal << 6) | row (only the ball holder can be there,
 *   SIM:369-372, and such a state is terminal, SIM:91-103).
 *
 * Observation index (SIM:63-109, 487-497), closed form checked against the reference's
 * enumeration for every state of 5x4, 6x4, 7x5, 9x6 and 11x7:
 *   obs = 0 for a goal state, else 1 + 2*(a*(F-1) + b - (b > a)) + p,  F = width*height.
 *
 * Injected randomness (bit-exact replay of the reference):
 *   rng8  : bits 0..1 = step draw r  -> u = (r + 0.5)/4      (slip_prob == 0: outcome r of 4,
 *                                                            r>>1 of 2, 0 of 1)
 *           bits 2..3 = reset draw r -> u = (r + 0.5)/4      (start state r of 4, r>>1 of 2)
 *   rng32 : step draw for slip_prob > 0, u = (r + 0.5)/2^32
 *   rngf64: step draw for slip_prob > 0 as the raw fp64 uniform the reference's
 *           np_random.random() returned (SIM:395)
 * Counter-based randomness (contract v2): Philox4x32-10, key = seed, counter = (group, step) with
 *   group = global env id >> 2; the env's word w is output word (env id & 3) -- one call serves the 4
 *   envs of an aligned group at one step, a pure function of (seed, global env id, step).  Decode,
 *   reading w as the fixed-point number x = w / 2^32:
 *     joint action ja = floor(25 x) = mulhi(w, 25), aa = ja / 5, ab = ja % 5;
 *     step draw    r32 = frac(25 x) * 2^32 = lo32(25 w): u = (r32 + 0.5) / 2^32, exactly the rng32
 *                  format (slip_prob == 0 uses its top two bits: mulhi(w, 100) = ja * 4 + (r32 >> 30));
 *     reset draw   (w >> 2) & 3.
 *   Kernels that own 4 envs per thread need env_id_base % 4 == 0 to take their vector path (other
 *   bases fall to the one-env-per-thread kernels; same results).
 *
 * flags byte: bit 0 terminated (SIM:403), bit 1 truncated (SIM:404).  Only when
 *   soccer_step_args.detail != 0: bits 2..3 = log2(number of outcomes the chosen slip
 *   combination had), bits 4..7 = index (0..8) of the chosen slip combination in the order of
 *   SIM:209-223.  0xFF = env skipped because its needs_reset bit was set (the assert at SIM:376).
 *
 * Two state layouts exist (a state tensor is in exactly one; soccer_convert_state translates):
 *   SOCCER_LAYOUT_CELL   the packed word above; used by the rules kernels (any pitch, any option)
 *   SOCCER_LAYOUT_INDEX  bits 0..15 observation index of the current state (1..nS-1), bits
 *                        16..23 timestep; used by the *_table kernels, which keep the whole
 *                        (state, joint action, draw) -> (next state, reward, done) table -- the
 *                        reference's P (SIM:167-293), produced on the device by the rules path --
 *                        resident in shared memory (5x4 pitch: 152 KB of the 227 KB per SM).
 */
#ifndef SOCCER_B200_H
#define SOCCER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOCCER_ABI_VERSION 2   /* 2: Philox contract v2, soccer_step_args.{stats,table,slip_index}, rollout flip_reward / slip_index */

#define SOCCER_OK        0
#define SOCCER_EINVAL   (-1)  /* NULL where a pointer is required, n < 0, bad option */
#define SOCCER_EPITCH   (-2)  /* width < 5, height < 4 (SIM:45-46); or width*height > 126 / height > 16: the 7-bit cell code
                                 and 16-bit observation lanes of this library (the reference has no upper limit) */
#define SOCCER_ESLIP    (-3)  /* slip_prob > 0 without rng32 / rngf64, or unsupported here */
#define SOCCER_EPOLICY  (-4)  /* both policies given (SIM:38) */

#define SOCCER_MAX_EPISODE_STEPS 100  /* SIM:404 */

typedef struct CUstream_st *soccer_stream_t; /* == cudaStream_t */

/* Constructor arguments of the reference class (SIM:35). */
typedef struct soccer_pitch {
    int32_t width;      /* unpadded, >= 5 */
    int32_t height;     /* >= 4 */
    double  slip_prob;  /* SIM:50 */
} soccer_pitch;

/* Constructor products (SIM:48-65, 146-165), computed on the host in closed form. */
typedef struct soccer_pitch_info {
    int32_t  padded_width;   /* SIM:48 */
    int32_t  height;
    int32_t  n_field_cells;  /* F */
    int32_t  nS;             /* 1 + 2F(F-1), SIM:105-108 */
    int32_t  nA;             /* 5, SIM:112 */
    int32_t  n_goal_rows;    /* SIM:60 */
    int32_t  goal_rows[3];
    int32_t  n_isd;          /* 4 or 2, SIM:151-163 */
    int32_t  isd_obs[4];
    uint32_t isd_state[4];   /* packed */
    int32_t  isd_tuple[4][5];
    double   slip_combo_prob[9]; /* mp of SIM:209-223, same fp64 expressions */
} soccer_pitch_info;

/* ---- host helpers (no GPU) ---- */
int soccer_abi_version(void);
/* SIM:45-65, 146-165 */
int soccer_pitch_info_host(const soccer_pitch *pitch, soccer_pitch_info *out);
/* tuple (xa, ya, xb, yb, p) <-> packed state; SIM:487-497 for the observation index.
 * pack returns SOCCER_EINVAL for tuples the reference classifies unreachable (SIM:74-88). */
int soccer_pack_state_host(const soccer_pitch *pitch, const int32_t tuple[5], int32_t timestep,
                           int32_t needs_reset, uint32_t *packed);
int soccer_unpack_state_host(const soccer_pitch *pitch, uint32_t packed, int32_t tuple[5],
                             int32_t *timestep, int32_t *needs_reset);
int soccer_state_to_obs_host(const soccer_pitch *pitch, uint32_t packed, int32_t *obs);
int soccer_obs_to_state_host(const soccer_pitch *pitch, int32_t obs, uint32_t *packed);

/* ---- K4: batched reset / state injection ---- */
/* reset() for n envs (SIM:410-424): start state from rng8 bits 2..3; mask (optional, [n])
 * selects the envs to reset; obs_out optional. */
int soccer_reset(const soccer_pitch *pitch, uint32_t *state, int32_t *obs_out,
                 const uint8_t *rng8, const uint8_t *mask, int64_t n, soccer_stream_t stream);
/* same, reset draw = Philox word (seed, env_id_base + i, step) bits 0..1 */
int soccer_reset_philox(const soccer_pitch *pitch, uint32_t *state, int32_t *obs_out,
                        const uint8_t *mask, uint64_t seed, uint64_t step, uint64_t env_id_base,
                        int64_t n, soccer_stream_t stream);
/* `env.state = ...` of the reference's tests, from observation indices 1..nS-1 (obs 0 or out of
 * range -> the env is left with needs_reset set); timestep_in optional (default 0). */
int soccer_set_state(const soccer_pitch *pitch, uint32_t *state, const int32_t *obs_in,
                     const int32_t *timestep_in, int64_t n, soccer_stream_t stream);
/* _state_to_observation (SIM:487-494) of the packed states */
int soccer_get_obs(const soccer_pitch *pitch, const uint32_t *state, int32_t *obs_out, int64_t n,
                   soccer_stream_t stream);

/* ---- K1: one lock-step step() of n envs (SIM:375-408) with fused auto-reset (SIM:410-424) ---- */
/* slip_prob == 0, joint actions and draws supplied by the caller; reset_obs optional */
int soccer_step(const soccer_pitch *pitch, uint32_t *state, const uint8_t *act_a,
                const uint8_t *act_b, const uint8_t *rng8, int32_t *obs, float *reward,
                uint8_t *flags, int32_t *reset_obs, int64_t n, soccer_stream_t stream);
/* same with Philox draws keyed (seed, env_id_base + i, step) */
int soccer_step_philox(const soccer_pitch *pitch, uint32_t *state, const uint8_t *act_a,
                       const uint8_t *act_b, uint64_t seed, uint64_t step, uint64_t env_id_base,
                       int32_t *obs, float *reward, uint8_t *flags, int32_t *reset_obs,
                       int64_t n, soccer_stream_t stream);

/* every option of the step path */
typedef struct soccer_step_args {
    uint32_t       *state;      /* [n] in/out */
    const uint8_t  *act_a;      /* [n]; NULL iff policy_a folds player A's action */
    const uint8_t  *act_b;      /* [n]; NULL iff policy_b folds player B's action */
    const uint8_t  *rng8;       /* [n] injected draws; NULL iff use_philox */
    const uint32_t *rng32;      /* [n] injected step draw, slip_prob > 0 */
    const double   *rngf64;     /* [n] injected step draw as raw fp64, slip_prob > 0 */
    const int8_t   *policy_a;   /* [nS] obs -> action, player A folded (SIM:187), reward sign
                                   flipped because the return agent is player_b (SIM:243-244) */
    const int8_t   *policy_b;   /* [nS] obs -> action, player B folded (SIM:188) */
    int32_t        *obs;        /* [n] out: what step() returned (0 on a goal) */
    float          *reward;     /* [n] out: player_a's reward (the single agent's when folded) */
    uint8_t        *flags;      /* [n] out */
    int32_t        *reset_obs;  /* [n] out, optional: observation the next step starts from */
    int64_t         n;
    int32_t         auto_reset; /* 1: reset fused (vector env); 0: set needs_reset like SIM:406 */
    int32_t         use_philox;
    int32_t         detail;     /* 1: fill flags bits 2..7 (needed for info["p"], SIM:405) */
    int32_t         narrow;     /* 1: obs points at uint16[n], reward at int8[n] (same values); else 0 */
    uint64_t        seed, step, env_id_base;
    /* optional, NULL = off: episode statistics of this step fused into the kernel (no second pass
     * over the streams): stats[0] += episodes ended, [1] += goals_A, [2] += goals_B (by player A's
     * reward sign whatever the return agent), [3] += truncations without a goal, [4] += env-steps,
     * [5] += summed length of the episodes that ended -- the vector soccer_rollout accumulates and
     * the multi-GPU path all-reduces.  Not offered together with slip_prob > 0 on the table path. */
    unsigned long long *stats;
    /* optional, NULL = rules kernels on SOCCER_LAYOUT_CELL states.  Non-NULL: the step table of
     * soccer_build_step_table; `state` is then in SOCCER_LAYOUT_INDEX and the shared-memory-table
     * kernels run (see below): injected or Philox draws, folded table policies (the single-agent
     * modes, SIM:187-188, 243-244), slip_prob > 0 (rng32 / rngf64 / Philox), narrow streams, stats.
     * Needs auto_reset == 1 and detail == 0. */
    const uint16_t *table;
    const uint8_t  *slip_index; /* optional accelerator of the table path for slip_prob > 0 (soccer_build_slip_index) */
} soccer_step_args;
int soccer_step_ex(const soccer_pitch *pitch, const soccer_step_args *args, soccer_stream_t stream);

/* Speculative step of ONE env (the single-env drop-in's latency path).
 * Replaces: SoccerSimultaneousEnv.step, soccer_simultaneous_env.py:375-408, called once per Python-loop iteration.
 * state_word: CELL-layout state of a running episode (field cells, timestep in bits 16..23, needs_reset clear).
 * slip_prob == 0 (u == NULL): one launch steps that state for all 25 joint actions x 4 values of the 2-bit step draw
 * (a folded player's action is its table policy's, SIM:187-188) WITHOUT auto-reset and writes 100 records of four
 * 32-bit words
 *     { next state word (needs_reset set when the episode ended, SIM:406), obs, reward (float bits),
 *       detail flags in bits 0..7 | seq << 8 }                                   (seq < 2^24)
 * to records[(aa * 5 + ab) * 4 + r], each with ONE 128-bit store, so the sequence number arrives together with the data
 * it vouches for.  slip_prob > 0: the fp64 draw cannot be enumerated; *u (a HOST pointer, read during the call) is the
 * draw the env's generator is going to make -- the caller mirrors its generator one draw ahead -- and the launch
 * steps the 25 joint actions with exactly that u by the reference's cumulative walk into records[(aa * 5 + ab) * 4]
 * (SOCCER_ESLIP when u == NULL).  `records` (1,600 bytes, 16-byte aligned) may be pinned host memory: the caller
 * enqueues this as soon as it knows the state, keeps working, and when the action (and the draw) arrive polls the last
 * word of the record it wants for `seq`, then reads the record. */
int soccer_step_speculate(const soccer_pitch *pitch, uint32_t state_word, const int8_t *policy_a, const int8_t *policy_b,
                          const double *u, uint32_t *records, uint32_t seq, soccer_stream_t stream);

/* Episode statistics of one lock-step step as a SEPARATE pass over its flags (and, optionally,
 * reward) streams: stats[0] += episodes ended, [1] += goals_A (reward > 0), [2] += goals_B,
 * [3] += truncations without a goal, [4] += n (steps); [5] (sum_episode_len) is left alone.
 * soccer_step_args.stats fuses the same counts (and [5]) into the step kernel itself. */
int soccer_step_stats(const uint8_t *flags, const float *reward, int64_t n,
                      unsigned long long *stats, soccer_stream_t stream);

/* Measurement probe, not part of the game: K1's memory traffic (7 bytes read, 13 written per env,
 * same access pattern and cache hints) with no game logic.  bench.py times it next to K1 to show
 * the practical HBM ceiling for K1's read:write mix.  n % 4 == 0, aligned pointers; it overwrites
 * state / obs / reward / flags with meaningless values.  mode 0 = K1's group order; 1, 2 = order
 * experiments (see k_stream_mix_probe). */
int soccer_bench_stream_mix(uint32_t *state, const uint8_t *act_a, const uint8_t *act_b,
                            const uint8_t *rng8, int32_t *obs, float *reward, uint8_t *flags,
                            int64_t n, int32_t mode, soccer_stream_t stream);

/* Measurement probe for K2: the fused rollout's memory traffic (state in/out once, K x 9 bytes per
 * env streamed to the [K][n] obs / reward / flags arrays with K2's stores, launch shape and slot
 * order) with no Philox and no game logic = practical HBM ceiling for a write-only stream mix.
 * mode 0 = K2's stores; 1-5 = store-flavour / slot-order experiments (see k_rollout_probe). */
int soccer_bench_rollout_probe(uint32_t *state, int32_t K, int32_t *obs, float *reward,
                               uint8_t *flags, int64_t n, int32_t mode, soccer_stream_t stream);

/* ---- K2: fused K-step rollout, state register-resident, on-device policy ---- */
/* policy_* == NULL -> uniform random joint action from the Philox word; else int8[nS] table.
 * obs/reward/flags are [K][n] streams (each optional).  stats[6] (optional, uint64, accumulated
 * with atomics): episodes, goals_A, goals_B, truncations, steps, sum_episode_len.
 * slip_prob >= 0 (slip_prob > 0: the word's 32-bit step draw through the reference's cumulative walk).
 * flip_reward != 0: the streamed reward is player B's (the negated value) -- what step() returns for an
 * env whose return agent is player_b (player A folded, SIM:243-244); the goals_A / goals_B statistics
 * keep counting by the unflipped sign.  K <= 2^28. */
int soccer_rollout(const soccer_pitch *pitch, uint32_t *state, const int8_t *policy_a,
                   const int8_t *policy_b, uint64_t seed, uint64_t step0, int32_t K,
                   uint64_t env_id_base, int32_t flip_reward, int32_t *obs, float *reward,
                   uint8_t *flags, unsigned long long *stats, int64_t n, soccer_stream_t stream);

/* ---- K3: exhaustive sweep = the transition table builder (SIM:167-293) ---- */
/* For every observation s in 1..nS-1, joint action ja = aa*5+ab, slip combination c
 * (n_combos = 1: only the no-slip moves; 9: all of SIM:209-223) and outcome slot k < 4:
 *   n_out[s-1][ja][c]            1, 2 or 4 (SIM:296-362)
 *   next_state[s-1][ja][c][k]    packed next state (goal cells kept, like P_readable)
 *   next_obs / reward / done     SIM:235-250; slots k >= n_out are zero-filled
 * Any output pointer may be NULL. */
int soccer_sweep(const soccer_pitch *pitch, int32_t n_combos, uint8_t *n_out, uint32_t *next_state,
                 int32_t *next_obs, int8_t *reward, uint8_t *done, soccer_stream_t stream);

/* ---- shared-memory-table variants of K1 / K2 (pitches whose table fits the 227 KB of an SM:
 *      nS <= 1131, i.e. 5x4 and 6x4) ---- */
#define SOCCER_LAYOUT_CELL  0
#define SOCCER_LAYOUT_INDEX 1
#define SOCCER_ETABLE (-5)   /* pitch too large for the shared-memory step table / slip_prob != 0 */
/* bytes of the step table for this pitch: nS * 100 * sizeof(int16_t), rounded up to 16 */
int soccer_step_table_bytes_host(const soccer_pitch *pitch, int64_t *bytes);
/* Fill table[obs*100 + aa*20 + ab*4 + r] for every state, joint action and 2-bit draw by
 * running the rules path (SIM:296-373, 235-240) on the device.  Entry (int16): bits 0..11 next
 * observation (0 = goal), bits 12..13 log2(#outcomes), bits 14..15 the reward as a signed 2-bit
 * field, so reward = entry >> 14 (arithmetic).  Row 0 is the absorbing terminal observation. */
int soccer_build_step_table(const soccer_pitch *pitch, uint16_t *table, soccer_stream_t stream);
/* soccer_step on SOCCER_LAYOUT_INDEX states with the table staged into shared memory by TMA */
int soccer_step_table(const soccer_pitch *pitch, const uint16_t *table, uint32_t *state,
                      const uint8_t *act_a, const uint8_t *act_b, const uint8_t *rng8, int32_t *obs,
                      float *reward, uint8_t *flags, int32_t *reset_obs, int64_t n,
                      soccer_stream_t stream);
/* soccer_step_philox (caller-supplied actions, draws from Philox4x32-10 keyed (seed, env_id_base + i,
 * step) exactly as there) through the shared-memory table: 19 algorithmic bytes per env-step.
 * SOCCER_LAYOUT_INDEX states; results identical to soccer_step_philox on the converted states. */
int soccer_step_table_philox(const soccer_pitch *pitch, const uint16_t *table, uint32_t *state,
                             const uint8_t *act_a, const uint8_t *act_b, uint64_t seed, uint64_t step,
                             uint64_t env_id_base, int32_t *obs, float *reward, uint8_t *flags,
                             int32_t *reset_obs, int64_t n, soccer_stream_t stream);
/* soccer_step / soccer_step_table (table == NULL: rules kernel on CELL-layout states, else the table kernel on
 * INDEX-layout states) with NARROW result streams: obs as uint16, reward as int8 -- the same values in 8 instead
 * of 13 written bytes per env-step.  The streams may be device memory or pinned host memory (zero copy). */
int soccer_step_narrow(const soccer_pitch *pitch, const uint16_t *table, uint32_t *state,
                       const uint8_t *act_a, const uint8_t *act_b, const uint8_t *rng8, uint16_t *obs16,
                       int8_t *reward8, uint8_t *flags, int64_t n, soccer_stream_t stream);
/* soccer_step_table with PACKED host-facing streams, for the PCIe-bound host path: joint[i] = aa | ab << 4 (the
 * joint action (aa, ab) that keys the reference's P[s], SIM:181-185, in one byte), rng8 as in soccer_step, and
 * ONE 16-bit result word per env:
 *     result[i] = obs | terminated << 12 | truncated << 13 | (reward & 3) << 14
 * i.e. read as int16: reward = w >> 14 (-1 / 0 / +1), obs = w & 0xFFF -- the same values soccer_step_table
 * returns, 2 bytes in and 2 bytes out per env-step.  joint / rng8 / result may be device memory or pinned host
 * memory (the kernel then moves them over PCIe itself).  SOCCER_LAYOUT_INDEX states. */
int soccer_step_table_packed(const soccer_pitch *pitch, const uint16_t *table, uint32_t *state,
                             const uint8_t *joint, const uint8_t *rng8, uint16_t *result, int64_t n,
                             soccer_stream_t stream);
/* the same with Philox draws keyed (seed, env_id_base + i, step): no draw stream, 1 byte in and 2 bytes
 * out per env-step */
int soccer_step_table_packed_philox(const soccer_pitch *pitch, const uint16_t *table, uint32_t *state,
                                    const uint8_t *joint, uint64_t seed, uint64_t step,
                                    uint64_t env_id_base, uint16_t *result, int64_t n,
                                    soccer_stream_t stream);
/* step() with slip_prob > 0 (SIM:203-256) through the same shared-memory table: the outcome counts of
 * the 9 slipped move pairs are read from the state's table row, the categorical draw walks them in
 * the reference's order with sequential fp64 sums (bit-exact), one more look-up yields the chosen
 * outcome.  The step draw u comes from rngf64[n] (raw) or rng32[n] ((r + 0.5) / 2^32), exactly one of
 * them; rng8 bits 2..3 carry the reset draw as in soccer_step.  SOCCER_LAYOUT_INDEX states. */
int soccer_step_table_slip(const soccer_pitch *pitch, const uint16_t *table, const uint8_t *slip_index,
                           uint32_t *state, const uint8_t *act_a, const uint8_t *act_b, const uint8_t *rng8,
                           const uint32_t *rng32, const double *rngf64, int32_t *obs, float *reward,
                           uint8_t *flags, int32_t *reset_obs, int64_t n, soccer_stream_t stream);
/* Optional accelerator of the slip_prob > 0 table kernels, built on the device from the step table and the pitch's
 * slip_prob; two planes with one entry per (obs, aa*5 + ab), P = nS * 25 rounded up to 16:
 *   plane 0, bytes [0, P), one byte per entry, fp64 draws (rngf64): an 8-bit mask over the first 8 of the 9 slip combinations (SIM:209-223 order): bit k
 *     set <=> a draw that the CONSTANT prefix sums of the combination probabilities assign to combination k has to take
 *     the reference's walk (combination k has 2 or 4 outcomes, or an earlier one changed a running sum in some bit).
 *     Most envs take the constant-prefix fast path, the others (and picks 8, 9) go through a warp-level queue.
 *   plane 1, bytes [P, 3P), one uint16 per entry, 32-bit draws (rng32 and every Philox draw): with u = (r + 0.5) / 2^32 each comparison "running sum <= u" of
 *     the reference's walk is "r >= ceil(sum * 2^32 - 0.5)", an integer threshold, and a last-bit difference between the
 *     true running sum and the constant one moves that integer only if an integer lies between the two.  The fast path
 *     therefore decides combination AND slot from constant integer thresholds; bit k (0 .. 8) marks the picks for which
 *     some threshold is not the constant one, bit 9 is always set.  The kernels do not read this plane: they use the
 *     state-independent list of soccer_slip_danger_host (same argument, proved by an error bound); the plane is its
 *     constructive cross-check (tests).
 * Results are identical with or without the index.  NULL = the fp64 path walks every env (32-bit draws need no index).
 * bytes: 3 * P. */
int soccer_slip_index_bytes_host(const soccer_pitch *pitch, int64_t *bytes);
/* The 32-bit draws that the integer-threshold slip path of this slip_prob hands to the reference's walk (a threshold
 * within 1e-4 of an integer, where the last bits of the running sums could matter): *n of them
 * (0 for ordinary values), or *n = -1 when there are more than 12 and the fast path is off.  Host only. */
int soccer_slip_danger_host(const soccer_pitch *pitch, uint32_t draws[12], int32_t *n);
int soccer_build_slip_index(const soccer_pitch *pitch, const uint16_t *table, uint8_t *slip_index,
                            soccer_stream_t stream);
/* soccer_rollout (uniform random policy; slip_prob >= 0) on SOCCER_LAYOUT_INDEX states */
int soccer_rollout_table(const soccer_pitch *pitch, const uint16_t *table, uint32_t *state,
                         uint64_t seed, uint64_t step0, int32_t K, uint64_t env_id_base,
                         int32_t *obs, float *reward, uint8_t *flags, unsigned long long *stats,
                         int64_t n, soccer_stream_t stream);
/* the same with on-device table policies (int8[nS] each, NULL = uniform random from the Philox
 * word, SIM:187-188 semantics as in soccer_rollout); the policies ride in shared memory next to
 * the table.  slip_index (optional, soccer_build_slip_index): with slip_prob > 0 and room for it next to
 * the table (5x4), 4 envs per thread take the constant-prefix fast path and the envs whose draw needs the
 * reference's cumulative walk go through a per-step warp queue; NULL = in-place walk.  flip_reward as in
 * soccer_rollout. */
int soccer_rollout_table_policy(const soccer_pitch *pitch, const uint16_t *table,
                                const uint8_t *slip_index, uint32_t *state, const int8_t *policy_a,
                                const int8_t *policy_b, uint64_t seed, uint64_t step0, int32_t K,
                                uint64_t env_id_base, int32_t flip_reward, int32_t *obs, float *reward,
                                uint8_t *flags, unsigned long long *stats, int64_t n,
                                soccer_stream_t stream);
/* ---- K2 for mid-size pitches (7x5, 8x5, 9x5, 7x6: nS <= 4096): the step table sharded over the shared memory of a
 * thread-block cluster of 2 / 4 / 8 CTAs, 128 KB per CTA, read through distributed shared memory.  Same table format
 * (soccer_build_cluster_table fills nS * 100 entries), SOCCER_LAYOUT_INDEX states, uniform policy, slip_prob == 0;
 * results identical to soccer_rollout on the converted states. */
int soccer_cluster_table_bytes_host(const soccer_pitch *pitch, int64_t *bytes, int32_t *cluster_size);
int soccer_build_cluster_table(const soccer_pitch *pitch, uint16_t *table, soccer_stream_t stream);
int soccer_rollout_table_cluster(const soccer_pitch *pitch, const uint16_t *table, uint32_t *state,
                                 uint64_t seed, uint64_t step0, int32_t K, uint64_t env_id_base,
                                 int32_t *obs, float *reward, uint8_t *flags, unsigned long long *stats,
                                 int64_t n, soccer_stream_t stream);
/* translate a state tensor between layouts (in place allowed); goal / needs_reset states map
 * to observation 0 in the INDEX layout and cannot be converted back */
int soccer_convert_state(const soccer_pitch *pitch, const uint32_t *in, uint32_t *out,
                         int32_t to_layout, int64_t n, soccer_stream_t stream);

/* T consecutive lock-steps -- the reference's `for t in range(T): env.step(actions[t])` replay loop
 * (SIM:375-408, with reset() SIM:410-424 fused wherever an episode ended) -- from [T][n] action /
 * draw arrays into [T][n] output arrays in ONE kernel launch: the state word stays in registers for
 * the T steps (k_replay / k_replay_table), so an env-step moves 3 bytes in and 9 out (12 + 8/T)
 * instead of K1's 20, and a small batch pays one dependent table look-up per step instead of one
 * launch.  Bit-identical to T calls of soccer_step / soccer_step_table.  table == NULL -> rules
 * kernel on CELL-layout states, else the table kernel on INDEX-layout states.  slip_prob must be 0.
 * reset_obs optional.  T == 1 forwards to the single-step entry points. */
int soccer_step_many(const soccer_pitch *pitch, const uint16_t *table, uint32_t *state, int32_t T,
                     const uint8_t *act_a, const uint8_t *act_b, const uint8_t *rng8, int32_t *obs,
                     float *reward, uint8_t *flags, int32_t *reset_obs, int64_t n,
                     soccer_stream_t stream);

/* ---- pinned host memory for the host-buffer path ----
 * bytes of page-locked, device-mapped host memory backed by 2 MB huge pages where the kernel grants them
 * (2 MB-aligned anonymous mapping, MADV_HUGEPAGE, cudaHostRegister).  DMA reads from such a region run at the
 * PCIe rate (50-55 GB/s here); from small cudaHostAlloc'd buffers they were measured at 21-55 GB/s depending on
 * the physical pages.  Falls back to cudaHostAlloc when the mapping cannot be registered (e.g. a low
 * RLIMIT_MEMLOCK).  Free with the same byte count.  Returns 0, SOCCER_EINVAL or a CUDA error code. */
int soccer_host_alloc(size_t bytes, void **ptr);
int soccer_host_free(void *ptr, size_t bytes);

/* ---- step() with HOST buffers (the end-to-end path) ----
 * Uploads the joint actions / draws, steps, downloads the results, software-pipelined over
 * n_chunks slices of the batch on three caller-supplied streams (upload of slice c+1, kernel of
 * slice c and download of slice c-1 overlap; PCIe is full duplex).  Host pointers should be
 * pinned (page-locked); pageable memory works but serialises.  The call only ENQUEUES: results
 * are valid once s_out has been synchronised.  slip_prob must be 0 here.
 *   narrow = SOCCER_HOST_WIDE (0):   h_obs int32[n], h_reward float32[n]   (3 bytes up, 9 down per env)
 *   narrow = SOCCER_HOST_NARROW (1): h_obs uint16[n], h_reward int8[n]     (3 up, 4 down; same values)
 *   narrow = SOCCER_HOST_PACKED (2): the streams of soccer_step_table_packed: h_act_a = joint bytes
 *            aa | ab << 4, h_act_b unused, h_obs = uint16 result words, h_reward / h_flags unused
 *            (2 up, 2 down); needs table != NULL
 * scratch: device memory, soccer_step_host_scratch_bytes_host(n) bytes, 256-byte aligned. */
#define SOCCER_HOST_WIDE   0
#define SOCCER_HOST_NARROW 1
#define SOCCER_HOST_PACKED 2
typedef struct soccer_step_host_args {
    uint32_t       *state;      /* device [n]; INDEX layout iff table != NULL, else CELL layout */
    const uint16_t *table;      /* device step table (soccer_build_step_table) or NULL: rules kernel */
    void           *scratch;    /* device */
    const uint8_t  *h_act_a, *h_act_b, *h_rng8;   /* host [n] */
    void           *h_obs;      /* host [n] */
    void           *h_reward;   /* host [n] */
    uint8_t        *h_flags;    /* host [n] */
    int64_t         n;
    int32_t         narrow;     /* SOCCER_HOST_WIDE / _NARROW / _PACKED */
    int32_t         n_chunks;   /* >= 1 */
    soccer_stream_t s_in, s_compute, s_out;
    int32_t         d2h_zero_copy; /* 1 (SOCCER_HOST_PACKED only; h_obs must be pinned + device-mapped, e.g. soccer_host_alloc):
                                      uploads go through the copy engine slice by slice as above, but each slice's kernel
                                      WRITES its result words straight into h_obs over PCIe (posted writes run at the link
                                      rate, unlike the kernel's PCIe reads) -- no download copies; results are valid once
                                      s_compute has been synchronised */
} soccer_step_host_args;
int soccer_step_host_scratch_bytes_host(int64_t n, int64_t *bytes);
int soccer_step_host(const soccer_pitch *pitch, const soccer_step_host_args *args);

/* Dense Pmat / Rmat (SIM:170-171, 258-279), fp64, accumulated in the reference's order.
 * Multi-agent: Pmat[nS][nS][5][5], Rmat[nS][5][5]; with a folded policy: Pmat[nS][nS][5],
 * Rmat[nS][5].  The call zero-fills both. */
int soccer_dense(const soccer_pitch *pitch, const int8_t *policy_a, const int8_t *policy_b,
                 double *Pmat, double *Rmat, soccer_stream_t stream);

/* ---- planners over the transition dynamics (/root/reference/gym_soccer/utils/planners.py = PL) ---- */
/* One Bellman backup, PL:8-11 / 35-39: Q[s][key] = sum over the reference's P[s][key] list, in list
 * order, of prob * (reward + gamma * V[next_state] * (not done)), with the reference's fp64 operation
 * order (bit-identical Q).  key = own action (5) for a single-agent env (one table policy given,
 * SIM:266-279; rewards from that player's side), the joint action aa*5+ab (25) otherwise.
 * V fp64[nS], Q fp64[nS][nkeys], device pointers.  The lists are enumerated on the fly by the
 * rules path; no table is read. */
int soccer_bellman_q(const soccer_pitch *pitch, const int8_t *policy_a, const int8_t *policy_b,
                     const double *V, double gamma, double *Q, soccer_stream_t stream);
/* A whole planner in ONE cooperative launch (grid-wide barriers, no host round trip per sweep):
 *   pi_in == NULL: value_iteration(env, theta, gamma), PL:4-18 -> V_out (the vector BEFORE the last
 *                  sweep, as there), Q_out[nS][nkeys], pi_out[nS] = argmax (first maximum), *sweeps_out;
 *   pi_in != NULL: policy_evaluation(pi_in, env, theta, gamma), PL:20-31 -> V_out, *sweeps_out.
 * Bit-identical to the reference's results, sweep count included.  workspace: device memory of
 * soccer_plan_workspace_bytes_host() bytes.  max_sweeps bounds the loop (the reference has none). */
int soccer_plan_workspace_bytes_host(const soccer_pitch *pitch, int64_t *bytes);
int soccer_plan(const soccer_pitch *pitch, const int8_t *policy_a, const int8_t *policy_b,
                const int32_t *pi_in, double theta, double gamma, int32_t max_sweeps, double *V_out,
                double *Q_out, int32_t *pi_out, int32_t *sweeps_out, void *workspace,
                soccer_stream_t stream);

/* The reference's DENSE planners (PL:57-87) are written against Pmat[s][s'][a] / Rmat[s][a] with np.dot.  Pmat has at
 * most 15 non-zeros per (s, a), so the contraction runs sparsely over the same on-the-fly transition lists as
 * soccer_bellman_q (no dense matrix is read, no library GEMM): Rmat[s][a] in the reference's accumulation order
 * (bit-identical), Pmat[s][:][a] . v as the list-order sum of prob * v[next_obs] (equal to the reference's BLAS sum to
 * fp64 round-off), including the reference's Pmat[0][0] quirk and no (not done) factor.
 *   soccer_dense_q:      q[s][key] = Rmat[s][key] + gamma * (Pmat[s][:][key] . v)                     (PL:78-79)
 *   soccer_policy_eval:  policy_eval(env, policy, theta, gamma, k = max_sweeps, init = v_init) (PL:57-70) in ONE
 *                        cooperative launch: v <- r_pi + gamma * P_pi v for the stochastic policy[nS][nkeys] until the
 *                        sup-norm change < theta or max_sweeps sweeps; v_init NULL = zeros; V_out[nS], *sweeps_out.
 * nkeys = 5 with one table policy folded (the single-agent env the reference's planners are written for), else 25.
 * workspace: device memory of soccer_policy_eval_workspace_bytes_host(pitch, nkeys) bytes. */
int soccer_dense_q(const soccer_pitch *pitch, const int8_t *policy_a, const int8_t *policy_b,
                   const double *v, double gamma, double *q, soccer_stream_t stream);
int soccer_policy_eval_workspace_bytes_host(const soccer_pitch *pitch, int32_t nkeys, int64_t *bytes);
int soccer_policy_eval(const soccer_pitch *pitch, const int8_t *policy_a, const int8_t *policy_b,
                       const double *policy, const double *v_init, double theta, double gamma,
                       int32_t max_sweeps, double *V_out, int32_t *sweeps_out, void *workspace,
                       soccer_stream_t stream);

/* ---- the one collective of the path over NVLink peer memory (multi-GPU, one process per GPU) ----
 * Sum all-reduce of the 6-entry statistics vector written as a kernel: every rank stores its vector into every
 * rank's SYMMETRIC buffer (peer stores through NVLink / NVSwitch), publishes an epoch flag with release semantics at
 * system scope, waits for the other ranks' flags in its own buffer and sums.  One 32-thread launch on the caller's
 * stream, directly behind the last step kernel: no host round trip, no helper stream (NCCL's 48-byte all-reduce
 * measures 20-25 us + the hop onto its own stream; this one ~5 us).
 *   peer_ptrs[world]: HOST array with the device address of every rank's buffer as mapped into THIS process (e.g.
 *     torch.distributed._symmetric_memory rendezvous -> buffer_ptrs), each of soccer_stats_allreduce_p2p_bytes_host()
 *     bytes, zero-filled once before the first call (with a barrier after the fill);
 *   epoch: 1, 2, 3 ... incremented by every rank on every call (all ranks make the same sequence of calls);
 *   stats[6]: in = this rank's vector, out = the sum over all ranks.  world <= 16.
 * A peer that never arrives traps the waiting kernel after ~20 s instead of hanging the GPU. */
int soccer_stats_allreduce_p2p_bytes_host(int64_t *bytes);
int soccer_stats_allreduce_p2p(const uint64_t *peer_ptrs, int32_t rank, int32_t world, uint64_t epoch,
                               unsigned long long *stats, soccer_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SOCCER_B200_H */

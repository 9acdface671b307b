/*
 * soccer_oracle.h -- CPU ORACLE for the Littman'94 soccer step/reset hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product path (gym_soccer_littman94_b200) never links, imports or calls it.
 *
 * It is a plain-C restatement of the reference's algorithm, following the
 * reference's own structure (tuple states, nested enumeration, a transition
 * table built at construction, step = table lookup + one categorical draw) --
 * deliberately NOT the cell-id / LUT formulation the CUDA kernels use, so that
 * agreement between the two is evidence and not tautology.
 *
 * Citations: SIM = /root/reference/gym_soccer/envs/soccer_simultaneous_env.py.
 *
 * Parity pin: the oracle is checked against (a) the known-answer transitions in
 * the reference's own tests (tests/test_oracle_kat.py), and (b) golden vectors
 * produced by running the UNMODIFIED reference in the build container behind
 * oracle/ref_shim (oracle/make_golden.py -> tests/golden/).
 */
#ifndef SOCCER_ORACLE_H
#define SOCCER_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* SIM:8-12 */
enum { ORC_NOOP = 0, ORC_NORTH = 1, ORC_SOUTH = 2, ORC_EAST = 3, ORC_WEST = 4 };

/* A state tuple (xa, ya, xb, yb, p): x = row, y = column (padded), p = 0 -> A has the ball. */
typedef struct { int xa, ya, xb, yb, p; } orc_state;

/* One entry of a P / P_readable transition list (SIM:245-256). */
typedef struct {
    double    prob;   /* mp * nsp */
    orc_state ns;     /* next tuple (goal tuples kept as tuples, like P_readable) */
    int       obs;    /* _state_to_observation(ns): 0 for goal tuples */
    double    reward; /* +1 / -1 / 0 (sign-flipped for a single-agent player_b env) */
    int       done;
} orc_trans;

typedef struct orc_model orc_model;

/* SIM:35-144.  width/height are the UNPADDED constructor arguments.
 * policy_a / policy_b: NULL, or int[nS] tables obs -> action (at most one non-NULL, SIM:38).
 * Returns NULL on an argument the reference would assert on. */
orc_model *orc_model_new(int width, int height, double slip_prob,
                         const int *policy_a, const int *policy_b);
void orc_model_free(orc_model *m);

int    orc_nS(const orc_model *m);
int    orc_nA(const orc_model *m);
int    orc_width(const orc_model *m);   /* padded, SIM:48 */
int    orc_height(const orc_model *m);
int    orc_multiagent(const orc_model *m);
int    orc_n_goal_rows(const orc_model *m);
int    orc_goal_row(const orc_model *m, int i);
int    orc_n_unreachable(const orc_model *m);
int    orc_n_goal_states(const orc_model *m);

/* SIM:146-165 */
int    orc_isd_len(const orc_model *m);
double orc_isd_prob(const orc_model *m, int i);
orc_state orc_isd_state(const orc_model *m, int i);

/* SIM:487-497.  state_to_obs: -1 when the tuple is unreachable / out of range. */
int       orc_state_to_obs(const orc_model *m, orc_state s);
orc_state orc_obs_to_state(const orc_model *m, int obs); /* obs 0 -> (-1,-1,-1,-1,-1) */
/* goal-state reward for player A (SIM:102); 0.0 if the tuple is not a goal state */
double    orc_goal_reward(const orc_model *m, orc_state s);
int       orc_is_goal_state(const orc_model *m, orc_state s);

/* SIM:364-373 */
void orc_next_cell(const orc_model *m, int x, int y, int d_col, int d_row, int has_ball,
                   int *nx, int *ny);
/* SIM:296-362.  out must hold 4 entries; returns the list length (1, 2 or 4). */
int  orc_get_next_state(const orc_model *m, orc_state st, int aa, int ab,
                        int maa_dc, int maa_dr, int mab_dc, int mab_dr,
                        double *probs, orc_state *next);

/* P_readable[st][key] (SIM:394).  key = aa*5+ab when multiagent, else the free
 * player's action.  Returns the list length and a pointer to the list, or -1. */
int orc_transitions(const orc_model *m, orc_state st, int key, const orc_trans **list);

/* Bulk dump of the whole table, observation-major; see the .c file. */
int orc_dump_table(const orc_model *m, int L, uint8_t *count, double *prob, int32_t *next_obs,
                   int8_t *reward, uint8_t *done, int8_t *next_tuple);

/* gym 0.26.2 categorical_sample restated (call sites SIM:395, SIM:414). */
int orc_categorical_sample(const double *probs, int n, double u);

/* Dense Pmat / Rmat exactly as SIM:170-171, 258-279 accumulate them (fp64, same order).
 * multiagent: Pmat[nS][nS][nA][nA], Rmat[nS][nA][nA]; single: Pmat[nS][nS][nA], Rmat[nS][nA].
 * Caller allocates and the function zero-fills. */
void orc_fill_pmat_rmat(const orc_model *m, double *Pmat, double *Rmat);

/* ---- one mutable environment: SIM:140-144, 375-424 ---- */
typedef struct {
    const orc_model *m;
    orc_state state;
    int timestep;
    int needs_reset;
} orc_env;

void orc_env_init(orc_env *e, const orc_model *m);
/* reset with the uniform draw u (the one np_random.random() call at SIM:414).
 * Returns the observation; *prob gets isd prob. */
int  orc_env_reset(orc_env *e, double u, double *prob);
/* step with the uniform draw u (SIM:395).  key as in orc_transitions.
 * Returns 0, or -1 if needs_reset is set (the assert at SIM:376). */
int  orc_env_step(orc_env *e, int key, double u,
                  int *obs, double *reward, int *done, int *truncated, double *prob);

/* ---- lock-step batch with the auto-reset contract: reference step, then
 * reference reset iff done or truncated.  Injected randomness:
 *   step draw  u = rng32 ? (rng32[i]+0.5)/2^32 : ((rng8[i]&3)+0.5)/4
 *   reset draw u = (((rng8[i]>>2)&3)+0.5)/4
 * Arrays are [T][N] row-major.  state_* arrays are [N] (in/out): tuples + timestep.
 * Outputs: obs = what reference step returned (0 on a goal), reward = player_a's (or the
 * single agent's) reward, flags bit0 terminated bit1 truncated, reset_obs = observation
 * the next step starts from (post-reset obs when a reset happened, else obs).
 * act_b may be NULL in single-agent mode (act_a then carries the free player's action).
 * n_threads <= 1 -> serial. */
void orc_rollout_injected(const orc_model *m, int64_t T, int64_t N,
                          orc_state *state, int32_t *timestep,
                          const uint8_t *act_a, const uint8_t *act_b,
                          const uint8_t *rng8, const uint32_t *rng32,
                          int32_t *obs, float *reward, uint8_t *flags, int32_t *reset_obs,
                          int n_threads);

/* ---- counter-based randomness (DESIGN.md "Philox contract") ---- */
/* Philox4x32-10 (Salmon et al., SC'11; Random123 reference constants). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* The 32-bit word that drives env `env_id` at absolute step `step` (contract v2: one Philox call =
 * the 4 envs of the aligned group env_id >> 2 at one step; word index env_id & 3). */
uint32_t orc_philox_word(uint64_t seed, uint64_t env_id, uint64_t step);
/* decode a word: joint action ja = mulhi(w, 25); step draw r32 = lo32(25 w) (u = (r32 + 0.5) / 2^32;
 * *r_step = r32 >> 30, the 2-bit draw of slip_prob == 0); reset draw = (w >> 2) & 3 */
void orc_philox_decode(uint32_t w, int *aa, int *ab, int *r_step, int *r_reset);
uint32_t orc_philox_r32(uint32_t w);

/* K-step rollout, uniform (policy == NULL) or table policies (int8 obs->action), Philox
 * draws keyed (seed, env_id_base + i, step0 + k).  stats[6] += {episodes, goals_A, goals_B,
 * truncations, steps, sum_episode_len}.  obs/reward/flags may be NULL.  slip_prob > 0: the step draw
 * is u = (r32 + 0.5) / 2^32. */
void orc_rollout_philox(const orc_model *m, int64_t K, int64_t N,
                        orc_state *state, int32_t *timestep,
                        const int8_t *policy_a, const int8_t *policy_b,
                        uint64_t seed, uint64_t step0, uint64_t env_id_base,
                        int32_t *obs, float *reward, uint8_t *flags,
                        int64_t *stats, int n_threads);

#ifdef __cplusplus
}
#endif
#endif

/*
 * soccer_oracle.c -- CPU ORACLE (test infrastructure, not product; see soccer_oracle.h).
 *
 * Plain-C restatement of /root/reference/gym_soccer/envs/soccer_simultaneous_env.py
 * ("SIM").  Every function cites the SIM lines it follows.  Structure mirrors the
 * reference on purpose: tuple states, nested-loop enumeration, transition table built
 * at construction, step = lookup + one categorical draw over fp64 cumulative sums.
 */
#include "soccer_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* SIM:24-30  ACTION_INT_TO_MOVE: action -> (d_col, d_row) */
static const int MOVE_DC[5] = { 0, 0, 0, 1, -1 };
static const int MOVE_DR[5] = { 0, -1, 1, 0, 0 };

enum { CLS_UNREACHABLE = -2, CLS_GOAL = -1 }; /* >= 1: observation index */

struct orc_model {
    int W, H;            /* padded width (SIM:48), height */
    double slip;
    int multiagent;      /* SIM:54 */
    int flip_reward;     /* SIM:243: single-agent env whose return agent is player_b */
    int *policy_a, *policy_b;
    int goal_rows[3], n_goal_rows; /* SIM:60 */
    int nS, nA;
    int n_tuples;
    int *cls;            /* per flattened tuple: CLS_* or obs index (SIM:66-106) */
    double *goal_r;      /* per flattened tuple: goal_states[tuple] (SIM:102) */
    orc_state *reverse;  /* obs -> tuple (SIM:109) */
    int n_unreachable, n_goal_states;
    int n_isd; double isd_p[4]; orc_state isd_s[4]; /* SIM:146-165 */
    int nkeys;           /* 25 multiagent, 5 single */
    int64_t *tr_off;     /* [n_tuples*nkeys + 1] CSR offsets, -1 where the tuple has no entry */
    int *tr_cnt;
    orc_trans *tr; int64_t tr_len, tr_cap;
};

static int flat(const orc_model *m, orc_state s)
{
    if (s.xa < 0 || s.xa >= m->H || s.xb < 0 || s.xb >= m->H || s.ya < 0 || s.ya >= m->W ||
        s.yb < 0 || s.yb >= m->W || s.p < 0 || s.p > 1)
        return -1;
    return (((s.xa * m->W + s.ya) * m->H + s.xb) * m->W + s.yb) * 2 + s.p;
}

static int in_goal_rows(const orc_model *m, int x)
{
    for (int i = 0; i < m->n_goal_rows; ++i)
        if (m->goal_rows[i] == x) return 1;
    return 0;
}

static int in_goal_cols(const orc_model *m, int y) { return y == 0 || y == m->W - 1; } /* SIM:61 */

static int state_eq(orc_state a, orc_state b)
{
    return a.xa == b.xa && a.ya == b.ya && a.xb == b.xb && a.yb == b.yb && a.p == b.p;
}

/* SIM:364-373 */
void orc_next_cell(const orc_model *m, int x, int y, int d_col, int d_row, int has_ball,
                   int *nx_out, int *ny_out)
{
    int nx = x + d_row;
    if (nx > m->H - 1) nx = m->H - 1;
    if (nx < 0) nx = 0;
    int ny = y + d_col;
    int xoob = (ny == 0 || ny == m->W - 1);
    int goal = xoob && in_goal_rows(m, nx) && has_ball;
    if (xoob && !goal) ny = y;
    *nx_out = nx;
    *ny_out = ny;
}

int orc_is_goal_state(const orc_model *m, orc_state s)
{
    int f = flat(m, s);
    return f >= 0 && m->cls[f] == CLS_GOAL;
}

double orc_goal_reward(const orc_model *m, orc_state s)
{
    int f = flat(m, s);
    return (f >= 0 && m->cls[f] == CLS_GOAL) ? m->goal_r[f] : 0.0;
}

/* SIM:296-362 */
int orc_get_next_state(const orc_model *m, orc_state st, int aa, int ab,
                       int maa_dc, int maa_dr, int mab_dc, int mab_dr,
                       double *probs, orc_state *next)
{
    int xa = st.xa, ya = st.ya, xb = st.xb, yb = st.yb, p = st.p;

    if (orc_is_goal_state(m, st)) { /* SIM:300-301 */
        probs[0] = 1.0; next[0] = st;
        return 1;
    }

    int nxa, nya, nxb, nyb;
    orc_next_cell(m, xa, ya, maa_dc, maa_dr, p == 0, &nxa, &nya); /* SIM:308 */
    orc_next_cell(m, xb, yb, mab_dc, mab_dr, p == 1, &nxb, &nyb); /* SIM:309 */

    int n = 0;
    int dy = ya - yb; if (dy < 0) dy = -dy;
    int dx = xa - xb; if (dx < 0) dx = -dx;

    if ((xa == xb && dy == 1 && nya == yb && nyb == ya) ||
        (ya == yb && dx == 1 && nxa == xb && nxb == xa)) {
        /* SIM:315-327  case 1: moving through each other */
        probs[n] = 0.5; next[n++] = (orc_state){ xa, ya, xb, yb, 0 };
        probs[n] = 0.5; next[n++] = (orc_state){ xa, ya, xb, yb, 1 };
    } else if ((nxa == xb && nya == yb && ab == ORC_NOOP) ||
               (nxb == xa && nyb == ya && aa == ORC_NOOP)) {
        /* SIM:330-335  case 2: into a standing opponent */
        probs[n] = 1.0; next[n++] = (orc_state){ xa, ya, xb, yb, 1 - p };
    } else if ((xa == nxa && ya == nya && aa != ORC_NOOP && nxb == xa && nyb == ya) ||
               (xb == nxb && yb == nyb && ab != ORC_NOOP && nxa == xb && nya == yb)) {
        /* SIM:338-344  case 3: into an opponent who bounced */
        probs[n] = 0.5; next[n++] = (orc_state){ xa, ya, xb, yb, 0 };
        probs[n] = 0.5; next[n++] = (orc_state){ xa, ya, xb, yb, 1 };
    } else if (nxa == nxb && nya == nyb) {
        /* SIM:347-356  case 4: same empty target */
        probs[n] = 0.25; next[n++] = (orc_state){ xa, ya, nxb, nyb, 0 };
        probs[n] = 0.25; next[n++] = (orc_state){ xa, ya, nxb, nyb, 1 };
        probs[n] = 0.25; next[n++] = (orc_state){ nxa, nya, xb, yb, 0 };
        probs[n] = 0.25; next[n++] = (orc_state){ nxa, nya, xb, yb, 1 };
    } else {
        /* SIM:357-360 */
        probs[n] = 1.0; next[n++] = (orc_state){ nxa, nya, nxb, nyb, p };
    }
    return n;
}

int orc_state_to_obs(const orc_model *m, orc_state s)
{
    int f = flat(m, s);
    if (f < 0) return -1;
    if (m->cls[f] == CLS_GOAL) return 0; /* SIM:493 */
    if (m->cls[f] == CLS_UNREACHABLE) return -1;
    return m->cls[f];
}

orc_state orc_obs_to_state(const orc_model *m, int obs)
{
    if (obs <= 0 || obs >= m->nS) return (orc_state){ -1, -1, -1, -1, -1 };
    return m->reverse[obs];
}

static void tr_push(orc_model *m, orc_trans t)
{
    if (m->tr_len == m->tr_cap) {
        m->tr_cap = m->tr_cap ? m->tr_cap * 2 : 1 << 16;
        m->tr = (orc_trans *)realloc(m->tr, (size_t)m->tr_cap * sizeof(orc_trans));
    }
    m->tr[m->tr_len++] = t;
}

/* SIM:66-106 */
static void enumerate_states(orc_model *m)
{
    int W = m->W, H = m->H;
    m->n_tuples = H * W * H * W * 2;
    m->cls = (int *)malloc(sizeof(int) * m->n_tuples);
    m->goal_r = (double *)calloc(m->n_tuples, sizeof(double));
    m->reverse = (orc_state *)malloc(sizeof(orc_state) * (m->n_tuples + 1));
    m->nS = 1; /* index 0 = TERMINAL_STATE, SIM:64-65 */
    m->reverse[0] = (orc_state){ -1, -1, -1, -1, -1 };
    for (int xa = 0; xa < H; ++xa)
    for (int ya = 0; ya < W; ++ya)
    for (int xb = 0; xb < H; ++xb)
    for (int yb = 0; yb < W; ++yb)
    for (int p = 0; p < 2; ++p) {
        orc_state s = { xa, ya, xb, yb, p };
        int f = flat(m, s);
        int a_gr = in_goal_rows(m, xa), b_gr = in_goal_rows(m, xb);
        int a_gc = in_goal_cols(m, ya), b_gc = in_goal_cols(m, yb);
        /* SIM:74-77 corners */
        if ((a_gc && !a_gr) || (b_gc && !b_gr)) { m->cls[f] = CLS_UNREACHABLE; m->n_unreachable++; continue; }
        /* SIM:80-83 in a goal without the ball */
        if ((a_gr && a_gc && p != 0) || (b_gr && b_gc && p != 1)) { m->cls[f] = CLS_UNREACHABLE; m->n_unreachable++; continue; }
        /* SIM:86-88 same cell */
        if (xa == xb && ya == yb) { m->cls[f] = CLS_UNREACHABLE; m->n_unreachable++; continue; }
        /* SIM:91-103 goal states */
        if ((a_gr && a_gc && p == 0) || (b_gr && b_gc && p == 1)) {
            int ga = (p == 0 && a_gr && ya == W - 1) || (p == 1 && b_gr && yb == W - 1);
            int gb = (p == 1 && b_gr && yb == 0) || (p == 0 && a_gr && ya == 0);
            m->cls[f] = CLS_GOAL;
            m->goal_r[f] = ga ? 1.0 : (gb ? -1.0 : 0.0);
            m->n_goal_states++;
            continue;
        }
        m->cls[f] = m->nS; /* SIM:105-106 */
        m->reverse[m->nS] = s;
        m->nS++;
    }
}

/* SIM:146-165 */
static void generate_isd(orc_model *m)
{
    int col_a = 2, col_b = m->W - 3;
    m->n_isd = 0;
    if (m->n_goal_rows % 2 == 0) {
        int mid = m->n_goal_rows / 2;
        int opts[2] = { m->goal_rows[mid - 1], m->goal_rows[mid] };
        for (int i = 0; i < 2; ++i) {
            int row_a = opts[i];
            int row_b = (row_a == opts[0]) ? opts[1] : opts[0];
            for (int poss = 0; poss < 2; ++poss) {
                m->isd_p[m->n_isd] = 0.25;
                m->isd_s[m->n_isd++] = (orc_state){ row_a, col_a, row_b, col_b, poss };
            }
        }
    } else {
        int mid = m->goal_rows[m->n_goal_rows / 2];
        for (int poss = 0; poss < 2; ++poss) {
            m->isd_p[m->n_isd] = 0.5;
            m->isd_s[m->n_isd++] = (orc_state){ mid, col_a, mid, col_b, poss };
        }
    }
}

/* SIM:167-293 (the P / P_readable part; Pmat/Rmat are in orc_fill_pmat_rmat) */
static void build_table(orc_model *m)
{
    int W = m->W, H = m->H;
    double sp = m->slip;
    m->nkeys = m->multiagent ? 25 : 5;
    int64_t nslots = (int64_t)m->n_tuples * m->nkeys;
    m->tr_off = (int64_t *)malloc(sizeof(int64_t) * nslots);
    m->tr_cnt = (int *)calloc(nslots, sizeof(int));
    for (int64_t i = 0; i < nslots; ++i) m->tr_off[i] = -1;

    for (int xa = 0; xa < H; ++xa)
    for (int ya = 0; ya < W; ++ya)
    for (int xb = 0; xb < H; ++xb)
    for (int yb = 0; yb < W; ++yb)
    for (int p = 0; p < 2; ++p) {
        orc_state st = { xa, ya, xb, yb, p };
        int f = flat(m, st);
        if (m->cls[f] == CLS_UNREACHABLE) continue; /* SIM:179 */
        int s = orc_state_to_obs(m, st);            /* SIM:182 */
        int st_goal = (m->cls[f] == CLS_GOAL);

        int aa_lo = 0, aa_hi = 5, ab_lo = 0, ab_hi = 5;  /* SIM:187-188 */
        if (m->policy_a) { aa_lo = m->policy_a[s]; aa_hi = aa_lo + 1; }
        if (m->policy_b) { ab_lo = m->policy_b[s]; ab_hi = ab_lo + 1; }

        for (int aa = aa_lo; aa < aa_hi; ++aa)
        for (int ab = ab_lo; ab < ab_hi; ++ab) {
            /* SIM:203-206 intended moves (d_col, d_row) and their two orthogonal slips */
            int ma[2] = { MOVE_DC[aa], MOVE_DR[aa] };
            int mb[2] = { MOVE_DC[ab], MOVE_DR[ab] };
            int mas[2][2] = { { -ma[1], ma[0] }, { ma[1], -ma[0] } };
            int mbs[2][2] = { { -mb[1], mb[0] }, { mb[1], -mb[0] } };
            /* SIM:209-223, same order, same fp64 expressions */
            const int *cma[9] = { ma, ma, ma, mas[0], mas[1], mas[0], mas[0], mas[1], mas[1] };
            const int *cmb[9] = { mb, mbs[0], mbs[1], mb, mb, mbs[0], mbs[1], mbs[0], mbs[1] };
            double cmp[9] = {
                (1 - sp) * (1 - sp),
                (1 - sp) * sp * 0.5, (1 - sp) * sp * 0.5,
                sp * (1 - sp) * 0.5, sp * (1 - sp) * 0.5,
                sp * sp * 0.25, sp * sp * 0.25, sp * sp * 0.25, sp * sp * 0.25,
            };
            int key = m->multiagent ? aa * 5 + ab : (m->policy_a ? ab : aa); /* SIM:259,267,274 */
            int64_t slot = (int64_t)f * m->nkeys + key;
            m->tr_off[slot] = m->tr_len;
            for (int c = 0; c < 9; ++c) {
                if (cmp[c] == 0) continue; /* SIM:226-227 */
                double nsp[4]; orc_state nss[4];
                int n = orc_get_next_state(m, st, aa, ab, cma[c][0], cma[c][1],
                                           cmb[c][0], cmb[c][1], nsp, nss); /* SIM:233 */
                for (int k = 0; k < n; ++k) {
                    orc_state ns = nss[k];
                    int d; double r;
                    int ns_goal = orc_is_goal_state(m, ns);
                    if (state_eq(st, ns) && st_goal) { d = 1; r = 0.0; }              /* SIM:235-236 */
                    else if (!state_eq(st, ns) && ns_goal) { d = 1; r = orc_goal_reward(m, ns); } /* SIM:237-238 */
                    else { d = 0; r = 0.0; }                                            /* SIM:239-240 */
                    double pr = cmp[c] * nsp[k];                                        /* SIM:241 */
                    if (m->flip_reward) r = -1 * r;                                     /* SIM:243-244 */
                    orc_trans t = { pr, ns, orc_state_to_obs(m, ns), r, d };
                    tr_push(m, t);
                }
            }
            m->tr_cnt[slot] = (int)(m->tr_len - m->tr_off[slot]);
        }
    }
}

orc_model *orc_model_new(int width, int height, double slip_prob,
                         const int *policy_a, const int *policy_b)
{
    if (policy_a && policy_b) return NULL;     /* SIM:38 */
    if (width < 5 || height < 4) return NULL;  /* SIM:45-46 */
    orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
    m->W = width + 2; /* SIM:48 */
    m->H = height;
    m->slip = slip_prob;
    m->multiagent = !policy_a && !policy_b;    /* SIM:54 */
    m->flip_reward = (policy_a != NULL);       /* return_agent == ['player_b'], SIM:55-56, 243 */
    m->nA = 5;
    if (height % 2 == 0) {                     /* SIM:60 */
        m->n_goal_rows = 2;
        m->goal_rows[0] = (height - 1) / 2; m->goal_rows[1] = height / 2;
    } else {
        m->n_goal_rows = 3;
        m->goal_rows[0] = height / 2 - 1; m->goal_rows[1] = height / 2; m->goal_rows[2] = height / 2 + 1;
    }
    enumerate_states(m);
    if (policy_a) { m->policy_a = (int *)malloc(sizeof(int) * m->nS); memcpy(m->policy_a, policy_a, sizeof(int) * m->nS); }
    if (policy_b) { m->policy_b = (int *)malloc(sizeof(int) * m->nS); memcpy(m->policy_b, policy_b, sizeof(int) * m->nS); }
    generate_isd(m);
    build_table(m);
    return m;
}

void orc_model_free(orc_model *m)
{
    if (!m) return;
    free(m->cls); free(m->goal_r); free(m->reverse); free(m->tr_off); free(m->tr_cnt); free(m->tr);
    free(m->policy_a); free(m->policy_b);
    free(m);
}

int orc_nS(const orc_model *m) { return m->nS; }
int orc_nA(const orc_model *m) { return m->nA; }
int orc_width(const orc_model *m) { return m->W; }
int orc_height(const orc_model *m) { return m->H; }
int orc_multiagent(const orc_model *m) { return m->multiagent; }
int orc_n_goal_rows(const orc_model *m) { return m->n_goal_rows; }
int orc_goal_row(const orc_model *m, int i) { return m->goal_rows[i]; }
int orc_n_unreachable(const orc_model *m) { return m->n_unreachable; }
int orc_n_goal_states(const orc_model *m) { return m->n_goal_states; }
int orc_isd_len(const orc_model *m) { return m->n_isd; }
double orc_isd_prob(const orc_model *m, int i) { return m->isd_p[i]; }
orc_state orc_isd_state(const orc_model *m, int i) { return m->isd_s[i]; }

int orc_transitions(const orc_model *m, orc_state st, int key, const orc_trans **list)
{
    int f = flat(m, st);
    if (f < 0 || key < 0 || key >= m->nkeys) return -1;
    int64_t slot = (int64_t)f * m->nkeys + key;
    if (m->tr_off[slot] < 0) return -1;
    *list = m->tr + m->tr_off[slot];
    return m->tr_cnt[slot];
}

/* Bulk dump of P / P_readable in observation-major order (obs 1..nS-1; row 0 = the entry
 * the last-enumerated goal state leaves in P[0], SIM:182-183).  Arrays are
 * [nS][nkeys][L] (next_tuple [nS][nkeys][L][5]); returns the longest list, or -1 if L is
 * too small.  Pass L = 0 to query the longest list only. */
int orc_dump_table(const orc_model *m, int L, uint8_t *count, double *prob, int32_t *next_obs,
                   int8_t *reward, uint8_t *done, int8_t *next_tuple)
{
    int maxlen = 0;
    for (int f = 0; f < m->n_tuples; ++f) {
        if (m->cls[f] == CLS_UNREACHABLE) continue;
        int s = (m->cls[f] == CLS_GOAL) ? 0 : m->cls[f];
        for (int key = 0; key < m->nkeys; ++key) {
            int64_t slot = (int64_t)f * m->nkeys + key;
            if (m->tr_off[slot] < 0) continue;
            int n = m->tr_cnt[slot];
            if (n > maxlen) maxlen = n;
            if (L == 0) continue;
            if (n > L) return -1;
            const orc_trans *l = m->tr + m->tr_off[slot];
            size_t base = ((size_t)s * m->nkeys + key);
            count[base] = (uint8_t)n; /* later goal states overwrite row 0, like P[0] */
            for (int k = 0; k < n; ++k) {
                size_t e = base * L + k;
                prob[e] = l[k].prob; next_obs[e] = l[k].obs;
                reward[e] = (int8_t)l[k].reward; done[e] = (uint8_t)l[k].done;
                next_tuple[e * 5 + 0] = (int8_t)l[k].ns.xa; next_tuple[e * 5 + 1] = (int8_t)l[k].ns.ya;
                next_tuple[e * 5 + 2] = (int8_t)l[k].ns.xb; next_tuple[e * 5 + 3] = (int8_t)l[k].ns.yb;
                next_tuple[e * 5 + 4] = (int8_t)l[k].ns.p;
            }
        }
    }
    return maxlen;
}

/* gym 0.26.2 categorical_sample: np.argmax(np.cumsum(p) > u); all-False -> 0 */
int orc_categorical_sample(const double *probs, int n, double u)
{
    double cs = 0.0;
    for (int i = 0; i < n; ++i) {
        cs = cs + probs[i]; /* np.cumsum: sequential fp64 adds */
        if (cs > u) return i;
    }
    return 0;
}

/* SIM:170-171, 258-279 */
void orc_fill_pmat_rmat(const orc_model *m, double *Pmat, double *Rmat)
{
    int nS = m->nS, nA = m->nA, W = m->W, H = m->H;
    size_t inner = m->multiagent ? (size_t)nA * nA : (size_t)nA;
    memset(Pmat, 0, sizeof(double) * (size_t)nS * nS * inner);
    memset(Rmat, 0, sizeof(double) * (size_t)nS * inner);
    for (int xa = 0; xa < H; ++xa)
    for (int ya = 0; ya < W; ++ya)
    for (int xb = 0; xb < H; ++xb)
    for (int yb = 0; yb < W; ++yb)
    for (int p = 0; p < 2; ++p) {
        orc_state st = { xa, ya, xb, yb, p };
        int f = flat(m, st);
        if (m->cls[f] == CLS_UNREACHABLE) continue;
        int s = orc_state_to_obs(m, st);
        for (int key = 0; key < m->nkeys; ++key) {
            const orc_trans *l; int n = orc_transitions(m, st, key, &l);
            if (n < 0) continue;
            double *R = Rmat + (size_t)s * inner + key;
            *R = 0; /* SIM:260,268,275 */
            for (int k = 0; k < n; ++k) {
                Pmat[((size_t)s * nS + l[k].obs) * inner + key] += l[k].prob; /* SIM:262 */
                *R += l[k].prob * l[k].reward;                                 /* SIM:263 */
            }
        }
    }
}

/* ---- one environment ---- */
void orc_env_init(orc_env *e, const orc_model *m)
{
    e->m = m;
    e->state = (orc_state){ -1, -1, -1, -1, -1 };
    e->timestep = 0;
    e->needs_reset = 1; /* SIM:140 */
}

/* SIM:410-424 */
int orc_env_reset(orc_env *e, double u, double *prob)
{
    int i = orc_categorical_sample(e->m->isd_p, e->m->n_isd, u); /* SIM:414 */
    e->state = e->m->isd_s[i];                                    /* SIM:415 */
    if (prob) *prob = e->m->isd_p[i];
    e->needs_reset = 0;
    e->timestep = 0;
    return orc_state_to_obs(e->m, e->state);
}

/* SIM:375-408 */
int orc_env_step(orc_env *e, int key, double u,
                 int *obs, double *reward, int *done, int *truncated, double *prob)
{
    if (e->needs_reset) return -1; /* SIM:376 */
    const orc_trans *l;
    int n = orc_transitions(e->m, e->state, key, &l); /* SIM:394 */
    if (n <= 0) return -2;
    double pr[64];
    if (n > 64) return -3;
    for (int i = 0; i < n; ++i) pr[i] = l[i].prob;
    int i = orc_categorical_sample(pr, n, u);         /* SIM:395 */
    e->state = l[i].ns;                               /* SIM:396 */
    e->timestep += 1;                                 /* SIM:399 */
    *obs = l[i].obs;                                  /* SIM:397 */
    *reward = l[i].reward;
    *done = l[i].done;
    *truncated = e->timestep >= 100;                  /* SIM:404 */
    if (prob) *prob = l[i].prob;
    e->needs_reset = *done || *truncated;             /* SIM:406 */
    return 0;
}

/* ---- lock-step batch, auto-reset = reference step then reference reset ---- */
void orc_rollout_injected(const orc_model *m, int64_t T, int64_t N,
                          orc_state *state, int32_t *timestep,
                          const uint8_t *act_a, const uint8_t *act_b,
                          const uint8_t *rng8, const uint32_t *rng32,
                          int32_t *obs, float *reward, uint8_t *flags, int32_t *reset_obs,
                          int n_threads)
{
#ifdef _OPENMP
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for num_threads(n_threads) schedule(static)
#endif
    for (int64_t i = 0; i < N; ++i) {
        orc_env e; orc_env_init(&e, m);
        e.state = state[i]; e.timestep = timestep[i]; e.needs_reset = 0;
        for (int64_t t = 0; t < T; ++t) {
            int64_t j = t * N + i;
            int key = m->multiagent ? act_a[j] * 5 + act_b[j] : act_a[j];
            double u = rng32 ? ((double)rng32[j] + 0.5) / 4294967296.0
                             : ((double)(rng8[j] & 3) + 0.5) / 4.0;
            int o, d, tr; double r;
            orc_env_step(&e, key, u, &o, &r, &d, &tr, NULL);
            int ro = o;
            if (d || tr) {
                double ur = ((double)((rng8[j] >> 2) & 3) + 0.5) / 4.0;
                ro = orc_env_reset(&e, ur, NULL);
            }
            if (obs) obs[j] = o;
            if (reward) reward[j] = (float)r;
            if (flags) flags[j] = (uint8_t)((d ? 1 : 0) | (tr ? 2 : 0));
            if (reset_obs) reset_obs[j] = ro;
        }
        state[i] = e.state; timestep[i] = e.timestep;
    }
}

/* ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as
 * 1, 2, 3", SC'11).  Multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key increments
 * 0x9E3779B9 / 0xBB67AE85, 10 rounds.  Checked against the Random123 known-answer
 * vectors in tests/test_philox.py. ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Philox contract v2.  One Philox call serves the 4 envs of an ALIGNED GROUP at one step:
 * counter = (group_lo, group_hi, step_lo, step_hi) with group = global env id >> 2, key = (seed_lo,
 * seed_hi); the env's word is output word (env id & 3).  A pure function of (seed, global env id,
 * step), hence independent of GPU count, launch boundaries and K. */
uint32_t orc_philox_word(uint64_t seed, uint64_t env_id, uint64_t step)
{
    uint64_t grp = env_id >> 2;
    uint32_t ctr[4] = { (uint32_t)grp, (uint32_t)(grp >> 32), (uint32_t)step, (uint32_t)(step >> 32) };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t out[4];
    orc_philox4x32_10(ctr, key, out);
    return out[env_id & 3];
}

/* Decode of one word w (fixed point: x = w / 2^32 uniform on [0, 1)):
 *   joint action  ja = floor(25 x) = mulhi(w, 25)            aa = ja / 5, ab = ja % 5
 *   step draw     r32 = frac(25 x) * 2^32 = lo32(25 w)       u = (r32 + 0.5) / 2^32, the rng32 format;
 *                 slip_prob == 0 uses its top two bits (r32 >> 30), which select the same outcome of
 *                 1 / 2 / 4 equiprobable ones as u does
 *   reset draw    (w >> 2) & 3
 * (25 is odd, so w -> r32 is a bijection of the 32-bit words: r32 is exactly uniform.) */
void orc_philox_decode(uint32_t w, int *aa, int *ab, int *r_step, int *r_reset)
{
    uint32_t ja = (uint32_t)(((uint64_t)w * 25u) >> 32);
    uint32_t r32 = (uint32_t)((uint64_t)w * 25u);
    *aa = (int)(ja / 5u);
    *ab = (int)(ja % 5u);
    *r_step = (int)(r32 >> 30);
    *r_reset = (int)((w >> 2) & 3u);
}
uint32_t orc_philox_r32(uint32_t w) { return (uint32_t)((uint64_t)w * 25u); }

void orc_rollout_philox(const orc_model *m, int64_t K, int64_t N,
                        orc_state *state, int32_t *timestep,
                        const int8_t *policy_a, const int8_t *policy_b,
                        uint64_t seed, uint64_t step0, uint64_t env_id_base,
                        int32_t *obs, float *reward, uint8_t *flags,
                        int64_t *stats, int n_threads)
{
    int64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0;
#ifdef _OPENMP
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for num_threads(n_threads) schedule(static) reduction(+:s0,s1,s2,s3,s4,s5)
#endif
    for (int64_t i = 0; i < N; ++i) {
        orc_env e; orc_env_init(&e, m);
        e.state = state[i]; e.timestep = timestep[i]; e.needs_reset = 0;
        for (int64_t k = 0; k < K; ++k) {
            int64_t j = k * N + i;
            uint32_t w = orc_philox_word(seed, env_id_base + (uint64_t)i, step0 + (uint64_t)k);
            int aa, ab, rs, rr;
            orc_philox_decode(w, &aa, &ab, &rs, &rr);
            int cur = orc_state_to_obs(m, e.state);
            if (policy_a) aa = policy_a[cur];
            if (policy_b) ab = policy_b[cur];
            int o, d, tr; double r;
            /* slip_prob > 0: the 32-bit step draw, u = (r32 + 0.5) / 2^32 like the injected rng32 format */
            double u = m->slip != 0.0 ? ((double)orc_philox_r32(w) + 0.5) / 4294967296.0
                                      : ((double)rs + 0.5) / 4.0;
            orc_env_step(&e, aa * 5 + ab, u, &o, &r, &d, &tr, NULL);
            s4 += 1;
            if (d || tr) {
                s0 += 1;
                s5 += e.timestep;
                if (d && r > 0) s1 += 1;
                if (d && r < 0) s2 += 1;
                if (!d) s3 += 1;
                orc_env_reset(&e, ((double)rr + 0.5) / 4.0, NULL);
            }
            if (obs) obs[j] = o;
            if (reward) reward[j] = (float)r;
            if (flags) flags[j] = (uint8_t)((d ? 1 : 0) | (tr ? 2 : 0));
        }
        state[i] = e.state; timestep[i] = e.timestep;
    }
    if (stats) { stats[0] += s0; stats[1] += s1; stats[2] += s2; stats[3] += s3; stats[4] += s4; stats[5] += s5; }
}

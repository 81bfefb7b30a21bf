/*
 * oracle/rvo2_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-precision CPU restatement of the ORCA agent step that DR-MPC
 * runs through Python-RVO2 (third-party, un-vendored, un-pinned:
 * github.com/sybrenstuvel/Python-RVO2, bundling RVO2 Library 2.0.x; installed
 * by `git clone && python setup.py install`, reference README.md:23-36).
 * The library is ABSENT from /root/reference, so this file restates its
 * published algorithm (RVO2 Agent::computeNeighbors / computeNewVelocity /
 * linearProgram1-3 / KdTree agent tree / RVOSimulator::doStep) and anchors on
 * the reference's own call sites:
 *     scripts/policy/orca.py:79-114                     (parameters, agent order, pref velocity)
 *     environment/human_avoidance/subENVs/crowd_sim.py:264-284 (observation order: other humans, robot last)
 *     environment/human_avoidance/utils/agent.py:96-97,164-168 (robot seen with vx=vy=0; fp64 Euler update)
 * PARITY UNPINNED: the reference has no tests / golden vectors for this path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file's shared object.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no fast-math, so every
 * fp32 operation rounds exactly once like the x86-64 SSE build of RVO2).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define RVO_EPSILON 0.00001f
#define RVO_MAX_LEAF 10

typedef struct { float x, y; } v2;
typedef struct { v2 point, dir; } line_t;

static inline v2 V(float x, float y) { v2 r = {x, y}; return r; }
static inline v2 vadd(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
static inline v2 vsub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
static inline v2 vmul(float s, v2 a) { return V(s * a.x, s * a.y); }
static inline float vdot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline float vdet(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }
static inline float vabssq(v2 a) { return vdot(a, a); }
static inline float vabs(v2 a) { return sqrtf(vdot(a, a)); }
/* RVO2 Vector2::operator/(float): multiply by the reciprocal. */
static inline v2 vdiv(v2 a, float s) { const float inv = 1.0f / s; return V(a.x * inv, a.y * inv); }
static inline v2 vnormalize(v2 a) { return vdiv(a, vabs(a)); }
static inline float sqrf(float a) { return a * a; }

/* ---- statistics a solve can report (used for "LP-branch divergence" counting) ---- */
typedef struct {
    int32_t n_lines;      /* neighbours that produced a half-plane                */
    int32_t lp2_fail;     /* line index where LP2 failed, == n_lines when feasible */
    int32_t n_lp1;        /* number of LP1 invocations in the top-level LP2        */
    int32_t lp1_work;     /* sum of pivot indices i over those LP1 calls           */
    uint32_t branch_mask; /* bit0 cut-off circle, bit1 left leg, bit2 right leg, bit3 collision seen */
    int32_t n_lp3_proj;   /* LP3 outer iterations that re-solved                   */
} orca_stats;

/* RVO2 linearProgram1 (Agent.cpp). */
static int lp1(const line_t *L, int lineNo, float radius, v2 opt, int dirOpt, v2 *res)
{
    const float dotProduct = vdot(L[lineNo].point, L[lineNo].dir);
    const float disc = sqrf(dotProduct) + sqrf(radius) - vabssq(L[lineNo].point);
    if (disc < 0.0f) return 0;
    const float sq = sqrtf(disc);
    float tLeft = -dotProduct - sq;
    float tRight = -dotProduct + sq;
    for (int i = 0; i < lineNo; ++i) {
        const float den = vdet(L[lineNo].dir, L[i].dir);
        const float num = vdet(L[i].dir, vsub(L[lineNo].point, L[i].point));
        if (fabsf(den) <= RVO_EPSILON) {
            if (num < 0.0f) return 0;
            continue;
        }
        const float t = num / den;
        if (den >= 0.0f) tRight = fminf(tRight, t);
        else tLeft = fmaxf(tLeft, t);
        if (tLeft > tRight) return 0;
    }
    if (dirOpt) {
        if (vdot(opt, L[lineNo].dir) > 0.0f) *res = vadd(L[lineNo].point, vmul(tRight, L[lineNo].dir));
        else *res = vadd(L[lineNo].point, vmul(tLeft, L[lineNo].dir));
    } else {
        const float t = vdot(L[lineNo].dir, vsub(opt, L[lineNo].point));
        if (t < tLeft) *res = vadd(L[lineNo].point, vmul(tLeft, L[lineNo].dir));
        else if (t > tRight) *res = vadd(L[lineNo].point, vmul(tRight, L[lineNo].dir));
        else *res = vadd(L[lineNo].point, vmul(t, L[lineNo].dir));
    }
    return 1;
}

/* RVO2 linearProgram2. */
static int lp2(const line_t *L, int n, float radius, v2 opt, int dirOpt, v2 *res, orca_stats *st)
{
    if (dirOpt) *res = vmul(radius, opt);
    else if (vabssq(opt) > sqrf(radius)) *res = vmul(radius, vnormalize(opt));
    else *res = opt;
    for (int i = 0; i < n; ++i) {
        if (vdet(L[i].dir, vsub(L[i].point, *res)) > 0.0f) {
            const v2 tmp = *res;
            if (st) { st->n_lp1++; st->lp1_work += i; }
            if (!lp1(L, i, radius, opt, dirOpt, res)) { *res = tmp; return i; }
        }
    }
    return n;
}

/* RVO2 linearProgram3 (numObstLines == 0: DR-MPC never adds obstacles). */
static void lp3(const line_t *L, int n, int beginLine, float radius, v2 *res, line_t *proj, orca_stats *st)
{
    float distance = 0.0f;
    for (int i = beginLine; i < n; ++i) {
        if (vdet(L[i].dir, vsub(L[i].point, *res)) > distance) {
            int m = 0;
            for (int j = 0; j < i; ++j) {
                line_t ln;
                const float determinant = vdet(L[i].dir, L[j].dir);
                if (fabsf(determinant) <= RVO_EPSILON) {
                    if (vdot(L[i].dir, L[j].dir) > 0.0f) continue;
                    ln.point = vmul(0.5f, vadd(L[i].point, L[j].point));
                } else {
                    ln.point = vadd(L[i].point,
                                    vmul(vdet(L[j].dir, vsub(L[i].point, L[j].point)) / determinant, L[i].dir));
                }
                ln.dir = vnormalize(vsub(L[j].dir, L[i].dir));
                proj[m++] = ln;
            }
            const v2 tmp = *res;
            if (st) st->n_lp3_proj++;
            if (lp2(proj, m, radius, V(-L[i].dir.y, L[i].dir.x), 1, res, NULL) < m) *res = tmp;
            distance = vdet(L[i].dir, vsub(L[i].point, *res));
        }
    }
}

/* One ORCA half-plane (RVO2 Agent::computeNewVelocity, agent-agent part). */
static line_t orca_line(v2 pos, v2 vel, float radius, v2 opos, v2 ovel, float oradius,
                        float invTimeHorizon, float invTimeStep, uint32_t *mask)
{
    const v2 rp = vsub(opos, pos);
    const v2 rv = vsub(vel, ovel);
    const float distSq = vabssq(rp);
    const float cr = radius + oradius;
    const float crSq = sqrf(cr);
    line_t line;
    v2 u;
    if (distSq > crSq) {
        const v2 w = vsub(rv, vmul(invTimeHorizon, rp));
        const float wLenSq = vabssq(w);
        const float dp1 = vdot(w, rp);
        if (dp1 < 0.0f && sqrf(dp1) > crSq * wLenSq) {
            const float wLen = sqrtf(wLenSq);
            const v2 unitW = vdiv(w, wLen);
            line.dir = V(unitW.y, -unitW.x);
            u = vmul(cr * invTimeHorizon - wLen, unitW);
            *mask |= 1u;
        } else {
            const float leg = sqrtf(distSq - crSq);
            if (vdet(rp, w) > 0.0f) {
                line.dir = vdiv(V(rp.x * leg - rp.y * cr, rp.x * cr + rp.y * leg), distSq);
                *mask |= 2u;
            } else {
                const v2 t = vdiv(V(rp.x * leg + rp.y * cr, -rp.x * cr + rp.y * leg), distSq);
                line.dir = V(-t.x, -t.y);
                *mask |= 4u;
            }
            const float dp2 = vdot(rv, line.dir);
            u = vsub(vmul(dp2, line.dir), rv);
        }
    } else {
        const v2 w = vsub(rv, vmul(invTimeStep, rp));
        const float wLen = vabs(w);
        const v2 unitW = vdiv(w, wLen);
        line.dir = V(unitW.y, -unitW.x);
        u = vmul(cr * invTimeStep - wLen, unitW);
        *mask |= 8u;
    }
    line.point = vadd(vel, vmul(0.5f, u));
    return line;
}

typedef struct { float distSq; int idx; } nb_t;

/* RVO2 Agent::insertAgentNeighbor. `nb` has capacity maxNeighbors. */
static void insert_neighbor(nb_t *nb, int *count, int maxNeighbors, float distSq, int idx, float *rangeSq)
{
    if (distSq < *rangeSq) {
        if (*count < maxNeighbors) { nb[*count].distSq = distSq; nb[*count].idx = idx; (*count)++; }
        int i = *count - 1;
        while (i != 0 && distSq < nb[i - 1].distSq) { nb[i] = nb[i - 1]; --i; }
        nb[i].distSq = distSq; nb[i].idx = idx;
        if (*count == maxNeighbors) *rangeSq = nb[*count - 1].distSq;
    }
}

/*
 * Single ORCA solve: agent `self` against `n` other agents visited in index order
 * (the visit order of a single-leaf kd-tree, i.e. n+1 <= 10; for larger crowds the
 * kd-tree visit order only matters on exact fp32 distSq ties).
 *   self8   = px,py,vx,vy,prefvx,prefvy,radius,maxSpeed      (fp32, already cast)
 *   others5 = n x (px,py,vx,vy,radius)
 *   out2    = new velocity
 */
void orca_ref_solve(int n, const float *self8, const float *others5,
                    float neighborDist, int maxNeighbors, float timeHorizon, float timeStep,
                    float *out2, orca_stats *st_out)
{
    orca_stats st; memset(&st, 0, sizeof st);
    const v2 pos = V(self8[0], self8[1]), vel = V(self8[2], self8[3]), pref = V(self8[4], self8[5]);
    const float radius = self8[6], maxSpeed = self8[7];
    nb_t *nb = (nb_t *)malloc(sizeof(nb_t) * (size_t)(maxNeighbors > 0 ? maxNeighbors : 1));
    line_t *L = (line_t *)malloc(sizeof(line_t) * (size_t)(n > 0 ? n : 1) * 2);
    line_t *proj = L + (n > 0 ? n : 1);
    int cnt = 0;
    if (maxNeighbors > 0) {
        float rangeSq = sqrf(neighborDist);
        for (int j = 0; j < n; ++j) {
            const v2 op = V(others5[5 * j], others5[5 * j + 1]);
            insert_neighbor(nb, &cnt, maxNeighbors, vabssq(vsub(pos, op)), j, &rangeSq);
        }
    }
    const float invTH = 1.0f / timeHorizon, invDt = 1.0f / timeStep;
    for (int k = 0; k < cnt; ++k) {
        const float *o = others5 + 5 * nb[k].idx;
        L[k] = orca_line(pos, vel, radius, V(o[0], o[1]), V(o[2], o[3]), o[4], invTH, invDt, &st.branch_mask);
    }
    v2 res;
    st.n_lines = cnt;
    st.lp2_fail = lp2(L, cnt, maxSpeed, pref, 0, &res, &st);
    if (st.lp2_fail < cnt) lp3(L, cnt, st.lp2_fail, maxSpeed, &res, proj, &st);
    out2[0] = res.x; out2[1] = res.y;
    if (st_out) *st_out = st;
    free(nb); free(L);
}

/* ------------------------------------------------------------------------------------------
 * Batched crowd pass with the reference's per-human semantics (crowd_sim.py:264-284 +
 * orca.py:64-117): for human i, agent 0 = human i (pref = goal-pos, normalised iff norm > 1,
 * computed in fp64 then cast), agents 1.. = other humans in list order, robot last with zero
 * velocity; every agent has ORCA radius `radius`; self maxSpeed = vmax_self.
 *   hpos,goal: [E,Nh,2] fp64 ; hvel: [E,Nh,2] fp32 ; rpos: [E,2] fp64 ; is_static: [E,Nh] or NULL
 *   out vnew: [E,Nh,2] fp32 ; stats: [E,Nh] or NULL
 * ------------------------------------------------------------------------------------------ */
void orca_ref_crowd_pass(int E, int Nh, const double *hpos, const float *hvel, const double *goal,
                         const double *rpos, const uint8_t *is_static, int robot_visible,
                         float neighborDist, float timeHorizon, float timeStep, float radius, float vmax_self,
                         float *vnew, orca_stats *stats, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
#endif
    {
        float *others = (float *)malloc(sizeof(float) * 5 * (size_t)(Nh + 1));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int e = 0; e < E; ++e) {
            const double *P = hpos + (size_t)e * Nh * 2, *G = goal + (size_t)e * Nh * 2;
            const float *Vv = hvel + (size_t)e * Nh * 2;
            for (int i = 0; i < Nh; ++i) {
                float *o2 = vnew + ((size_t)e * Nh + i) * 2;
                if (is_static && is_static[(size_t)e * Nh + i]) {
                    o2[0] = 0.f; o2[1] = 0.f;
                    if (stats) memset(&stats[(size_t)e * Nh + i], 0, sizeof(orca_stats));
                    continue;
                }
                int n = 0;
                for (int j = 0; j < Nh; ++j) {
                    if (j == i) continue;
                    others[5 * n + 0] = (float)P[2 * j]; others[5 * n + 1] = (float)P[2 * j + 1];
                    others[5 * n + 2] = Vv[2 * j]; others[5 * n + 3] = Vv[2 * j + 1];
                    others[5 * n + 4] = radius; ++n;
                }
                if (robot_visible) {
                    others[5 * n + 0] = (float)rpos[2 * e]; others[5 * n + 1] = (float)rpos[2 * e + 1];
                    others[5 * n + 2] = 0.f; others[5 * n + 3] = 0.f; others[5 * n + 4] = radius; ++n;
                }
                /* orca.py:97-100 in fp64 */
                const double dx = G[2 * i] - P[2 * i], dy = G[2 * i + 1] - P[2 * i + 1];
                const double speed = sqrt(dx * dx + dy * dy);
                double pvx = dx, pvy = dy;
                if (speed > 1.0) { pvx = dx / speed; pvy = dy / speed; }
                float self8[8] = {(float)P[2 * i], (float)P[2 * i + 1], Vv[2 * i], Vv[2 * i + 1],
                                  (float)pvx, (float)pvy, radius, vmax_self};
                orca_ref_solve(n, self8, others, neighborDist, n, timeHorizon, timeStep, o2,
                               stats ? &stats[(size_t)e * Nh + i] : NULL);
            }
        }
        free(others);
    }
}

/* ------------------------------------------------------------------------------------------
 * Full simulator with RVO2's kd-tree: the surface scripts/policy/orca.py:80-114 uses
 * (PyRVOSimulator / addAgent / setAgent* / doStep / getAgentVelocity). doStep solves ALL
 * agents like RVO2 does (the reference discards all but agent 0).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    v2 pos, vel, pref, newvel;
    float neighborDist, timeHorizon, timeHorizonObst, radius, maxSpeed;
    int maxNeighbors;
} sim_agent;

typedef struct { int begin, end, left, right; float maxX, maxY, minX, minY; } kd_node;

typedef struct {
    float timeStep, globalTime;
    sim_agent defaults;
    sim_agent *agents; int n, cap;
    int *kd_agents; int kd_n;      /* persistent permutation, like KdTree::agents_ */
    kd_node *tree;
} rvo_sim;

rvo_sim *rvo_sim_create(float timeStep, float neighborDist, int maxNeighbors, float timeHorizon,
                        float timeHorizonObst, float radius, float maxSpeed, float vx, float vy)
{
    rvo_sim *s = (rvo_sim *)calloc(1, sizeof(rvo_sim));
    s->timeStep = timeStep;
    s->defaults.neighborDist = neighborDist; s->defaults.maxNeighbors = maxNeighbors;
    s->defaults.timeHorizon = timeHorizon; s->defaults.timeHorizonObst = timeHorizonObst;
    s->defaults.radius = radius; s->defaults.maxSpeed = maxSpeed; s->defaults.vel = V(vx, vy);
    return s;
}
void rvo_sim_destroy(rvo_sim *s) { if (!s) return; free(s->agents); free(s->kd_agents); free(s->tree); free(s); }

int rvo_sim_add_agent(rvo_sim *s, float px, float py, float neighborDist, int maxNeighbors, float timeHorizon,
                      float timeHorizonObst, float radius, float maxSpeed, float vx, float vy)
{
    if (s->n == s->cap) { s->cap = s->cap ? 2 * s->cap : 16; s->agents = (sim_agent *)realloc(s->agents, sizeof(sim_agent) * (size_t)s->cap); }
    sim_agent *a = &s->agents[s->n];
    memset(a, 0, sizeof *a);
    a->pos = V(px, py); a->vel = V(vx, vy);
    a->neighborDist = neighborDist; a->maxNeighbors = maxNeighbors; a->timeHorizon = timeHorizon;
    a->timeHorizonObst = timeHorizonObst; a->radius = radius; a->maxSpeed = maxSpeed;
    return s->n++;
}
int rvo_sim_num_agents(const rvo_sim *s) { return s->n; }
void rvo_sim_set_position(rvo_sim *s, int i, float x, float y) { s->agents[i].pos = V(x, y); }
void rvo_sim_set_velocity(rvo_sim *s, int i, float x, float y) { s->agents[i].vel = V(x, y); }
void rvo_sim_set_pref_velocity(rvo_sim *s, int i, float x, float y) { s->agents[i].pref = V(x, y); }
void rvo_sim_get_velocity(const rvo_sim *s, int i, float *out2) { out2[0] = s->agents[i].vel.x; out2[1] = s->agents[i].vel.y; }
void rvo_sim_get_position(const rvo_sim *s, int i, float *out2) { out2[0] = s->agents[i].pos.x; out2[1] = s->agents[i].pos.y; }

static void kd_build_rec(rvo_sim *s, int begin, int end, int node)
{
    kd_node *T = s->tree; int *A = s->kd_agents;
    T[node].begin = begin; T[node].end = end;
    T[node].minX = T[node].maxX = s->agents[A[begin]].pos.x;
    T[node].minY = T[node].maxY = s->agents[A[begin]].pos.y;
    for (int i = begin + 1; i < end; ++i) {
        T[node].maxX = fmaxf(T[node].maxX, s->agents[A[i]].pos.x);
        T[node].minX = fminf(T[node].minX, s->agents[A[i]].pos.x);
        T[node].maxY = fmaxf(T[node].maxY, s->agents[A[i]].pos.y);
        T[node].minY = fminf(T[node].minY, s->agents[A[i]].pos.y);
    }
    if (end - begin > RVO_MAX_LEAF) {
        const int isVertical = (T[node].maxX - T[node].minX > T[node].maxY - T[node].minY);
        const float split = isVertical ? 0.5f * (T[node].maxX + T[node].minX) : 0.5f * (T[node].maxY + T[node].minY);
        int left = begin, right = end;
        while (left < right) {
            while (left < right && (isVertical ? s->agents[A[left]].pos.x : s->agents[A[left]].pos.y) < split) ++left;
            while (right > left && (isVertical ? s->agents[A[right - 1]].pos.x : s->agents[A[right - 1]].pos.y) >= split) --right;
            if (left < right) { int t = A[left]; A[left] = A[right - 1]; A[right - 1] = t; ++left; --right; }
        }
        if (left == begin) { ++left; ++right; }
        T[node].left = node + 1;
        T[node].right = node + 2 * (left - begin);
        kd_build_rec(s, begin, left, T[node].left);
        kd_build_rec(s, left, end, T[node].right);
    }
}

static void kd_query_rec(const rvo_sim *s, int self, nb_t *nb, int *cnt, float *rangeSq, int node)
{
    const kd_node *T = s->tree; const sim_agent *me = &s->agents[self];
    if (T[node].end - T[node].begin <= RVO_MAX_LEAF) {
        for (int i = T[node].begin; i < T[node].end; ++i) {
            const int j = s->kd_agents[i];
            if (j != self) insert_neighbor(nb, cnt, me->maxNeighbors, vabssq(vsub(me->pos, s->agents[j].pos)), j, rangeSq);
        }
    } else {
        const kd_node *l = &T[T[node].left], *r = &T[T[node].right];
        const float dl = sqrf(fmaxf(0.0f, l->minX - me->pos.x)) + sqrf(fmaxf(0.0f, me->pos.x - l->maxX)) +
                         sqrf(fmaxf(0.0f, l->minY - me->pos.y)) + sqrf(fmaxf(0.0f, me->pos.y - l->maxY));
        const float dr = sqrf(fmaxf(0.0f, r->minX - me->pos.x)) + sqrf(fmaxf(0.0f, me->pos.x - r->maxX)) +
                         sqrf(fmaxf(0.0f, r->minY - me->pos.y)) + sqrf(fmaxf(0.0f, me->pos.y - r->maxY));
        if (dl < dr) {
            if (dl < *rangeSq) {
                kd_query_rec(s, self, nb, cnt, rangeSq, T[node].left);
                if (dr < *rangeSq) kd_query_rec(s, self, nb, cnt, rangeSq, T[node].right);
            }
        } else {
            if (dr < *rangeSq) {
                kd_query_rec(s, self, nb, cnt, rangeSq, T[node].right);
                if (dl < *rangeSq) kd_query_rec(s, self, nb, cnt, rangeSq, T[node].left);
            }
        }
    }
}

void rvo_sim_do_step(rvo_sim *s)
{
    const int n = s->n;
    if (n == 0) return;
    if (s->kd_n < n) {
        s->kd_agents = (int *)realloc(s->kd_agents, sizeof(int) * (size_t)n);
        for (int i = s->kd_n; i < n; ++i) s->kd_agents[i] = i;
        s->kd_n = n;
        s->tree = (kd_node *)realloc(s->tree, sizeof(kd_node) * (size_t)(2 * n - 1));
    }
    kd_build_rec(s, 0, n, 0);
    nb_t *nb = (nb_t *)malloc(sizeof(nb_t) * (size_t)n);
    line_t *L = (line_t *)malloc(sizeof(line_t) * (size_t)n * 2);
    for (int a = 0; a < n; ++a) {
        sim_agent *me = &s->agents[a];
        int cnt = 0;
        if (me->maxNeighbors > 0) {
            float rangeSq = sqrf(me->neighborDist);
            kd_query_rec(s, a, nb, &cnt, &rangeSq, 0);
        }
        const float invTH = 1.0f / me->timeHorizon, invDt = 1.0f / s->timeStep;
        uint32_t mask = 0;
        for (int k = 0; k < cnt; ++k) {
            const sim_agent *o = &s->agents[nb[k].idx];
            L[k] = orca_line(me->pos, me->vel, me->radius, o->pos, o->vel, o->radius, invTH, invDt, &mask);
        }
        v2 res;
        const int fail = lp2(L, cnt, me->maxSpeed, me->pref, 0, &res, NULL);
        if (fail < cnt) lp3(L, cnt, fail, me->maxSpeed, &res, L + n, NULL);
        me->newvel = res;
    }
    for (int a = 0; a < n; ++a) {
        sim_agent *me = &s->agents[a];
        me->vel = me->newvel;
        me->pos = vadd(me->pos, vmul(s->timeStep, me->vel));
    }
    s->globalTime += s->timeStep;
    free(nb); free(L);
}

/* ORACLE — test infrastructure only.  Whole-episode CPU port of the hot path in C: Agent.evaluate (base/core/agent.py:63-138)
 * = CitationEnv.reset/step (envs/phlabenv.py:401-482) + fault shims (envs/{be,jr,sa,se}/citation.py:71-79) +
 * Actor.select_action (base/core/genetic_agent.py:104-109, LayerNorm base/core/mod_utils.py:47-50) + the plant restatement
 * (plant_oracle.c).  It exists so that large parity sweeps and an optimised multi-core CPU baseline do not pay the
 * Python / torch per-step dispatch of the reference's own execution model (that one is oracle/phlab.py + oracle/actor.py,
 * which stays the pinned oracle).  The forward pass sums in index order in float32 — yet another legal summation order,
 * so agreement with the torch oracle is to fp32 round-off, not bitwise. */
#include <math.h>
#include <string.h>

void plant_get_ic(int v, double* X);
void plant_step(int v, double* X, const double* U);

static float act_fn(int act, float x)
{
    if (act == 0) return tanhf(x);
    if (act == 1) return x > 0.f ? x : expm1f(x);
    return x > 0.f ? x : 0.01f * x;
}

static void actor_forward(const float* p, int S, int A, int H, int L, int act, const float* obs, float* action)
{
    float a[512], b[512];
    for (int j = 0; j < H; ++j) {
        float acc = 0.f;
        for (int i = 0; i < S; ++i) acc += p[j * S + i] * obs[i];
        a[j] = act_fn(act, acc + p[H * S + j]);
    }
    p += H * S + H;
    for (int l = 0; l < L; ++l) {
        const float *W = p, *bias = p + H * H, *gamma = bias + H, *beta = gamma + H;
        float sum = 0.f;
        for (int j = 0; j < H; ++j) {
            float acc = 0.f;
            for (int i = 0; i < H; ++i) acc += W[j * H + i] * a[i];
            b[j] = acc + bias[j];
            sum += b[j];
        }
        const float mean = sum / (float)H;
        float ss = 0.f;
        for (int j = 0; j < H; ++j) ss += (b[j] - mean) * (b[j] - mean);
        const float den = sqrtf(ss / (float)(H - 1)) + 1e-6f;
        for (int j = 0; j < H; ++j) a[j] = act_fn(act, gamma[j] * (b[j] - mean) / den + beta[j]);
        p += H * H + 3 * H;
    }
    for (int j = 0; j < A; ++j) {
        float acc = 0.f;
        for (int i = 0; i < H; ++i) acc += p[j * H + i] * a[i];
        action[j] = tanhf(acc + p[A * H + j]);
    }
}

/* width-list generalisation (BASELINE config 5): Linear(S,w0) act {Linear(w_i,w_i+1) LayerNorm act} Linear(w_n,A) tanh */
static void actor_forward_wide(const float* p, int S, int A, const int* widths, int nw, int act, const float* obs, float* action)
{
    float a[1024], b[1024];
    int H = widths[0];
    for (int j = 0; j < H; ++j) {
        float acc = 0.f;
        for (int i = 0; i < S; ++i) acc += p[j * S + i] * obs[i];
        a[j] = act_fn(act, acc + p[H * S + j]);
    }
    p += H * S + H;
    for (int l = 1; l < nw; ++l) {
        const int Hi = widths[l - 1], Ho = widths[l];
        const float *W = p, *bias = p + Ho * Hi, *gamma = bias + Ho, *beta = gamma + Ho;
        float sum = 0.f;
        for (int j = 0; j < Ho; ++j) {
            float acc = 0.f;
            for (int i = 0; i < Hi; ++i) acc += W[j * Hi + i] * a[i];
            b[j] = acc + bias[j];
            sum += b[j];
        }
        const float mean = sum / (float)Ho;
        float ss = 0.f;
        for (int j = 0; j < Ho; ++j) ss += (b[j] - mean) * (b[j] - mean);
        const float den = sqrtf(ss / (float)(Ho - 1)) + 1e-6f;
        for (int j = 0; j < Ho; ++j) a[j] = act_fn(act, gamma[j] * (b[j] - mean) / den + beta[j]);
        p += Ho * Hi + 3 * Ho;
        H = Ho;
    }
    for (int j = 0; j < A; ++j) {
        float acc = 0.f;
        for (int i = 0; i < H; ++i) acc += p[j * H + i] * a[i];
        action[j] = tanhf(acc + p[A * H + j]);
    }
}

static const int* g_widths = 0;      /* set by oracle_population_wide for the duration of a call */
static int g_nw = 0;

static double ref_deg(const double* lv, const double* st, double t, double offset, double smooth_w)
{
    int k = 0;
    for (int j = 1; j < 6; ++j)
        if (t >= st[j]) k = j;
    if (k == 0) return offset + lv[0];
    const double x = (t - st[k]) / smooth_w;
    if (x >= 1.0) return offset + lv[k];
    return offset + (lv[k - 1] + (lv[k] - lv[k - 1]) * (0.5 * (1.0 - cos(3.141592653589793 * x))));
}

static void apply_fault(int fault, const double* u, double* c)
{
    const double DEG2RAD = 0.017453292519943295;
    c[0] = u[0]; c[1] = u[1]; c[2] = u[2];
    if (fault == 1) c[0] = u[0] * 0.3;
    else if (fault == 2) c[2] = 15 * 3.14159 / 180;
    else if (fault == 3) { const double b = 1.0 * DEG2RAD; c[1] = fmin(fmax(u[1], -b), b); }
    else if (fault == 4) { const double b = 2.5 * DEG2RAD; c[0] = fmin(fmax(u[0], -b), b); }
}

void ko_actor_forward(const float* p, int S, int A, int H, int L, int act, const float* obs, float* action);   /* actor_kernel_order.c */

/* one episode; mode = variant | fault << 8; levels / starts: [2][6]; returns the episodic return, *steps = executed steps.
 * actor_order: 0 = index-order float32 sums + libm tanh (a stand-in for the reference's torch forward), 1 = the device
 * kernel's summation order and activation arithmetic (actor_kernel_order.c): the GPU must reproduce mode 1 exactly. */
double oracle_episode(const float* genome, int S, int A, int H, int L, int act, int mode,
                      const double* levels, const double* starts, double t_max, double smooth_w, int horizon, int* steps,
                      int actor_order)
{
    const double DEG2RAD = 0.017453292519943295, RAD2DEG = 57.29577951308232;
    const double bound = 10.0 * DEG2RAD, max_theta = 60.0 * DEG2RAD, max_phi = 75.0 * DEG2RAD;
    const double k_err = 6.0 / 3.141592653589793, k_err4 = k_err * 4.0;
    const int variant = mode & 0xff, fault = (mode >> 8) & 0xff;
    double X[19], U[3] = {0, 0, 0}, cmd[3], xo[12];
    plant_get_ic(variant, X);
    float obs[7] = {0.f, 0.f, 0.f, (float)X[0], (float)X[1], (float)X[2], (float)X[4]};
    const double theta_trim = X[7] * RAD2DEG;
    apply_fault(fault, U, cmd);
    plant_step(variant, X, cmd);
    double t = 0.0, ret = 0.0;
    int k = 0;
    for (; k < horizon;) {
        float a[3];
        if (actor_order == 2) actor_forward_wide(genome, S, A, g_widths, g_nw, act, obs, a);
        else if (actor_order == 1) ko_actor_forward(genome, S, A, H, L, act, obs, a);
        else actor_forward(genome, S, A, H, L, act, obs, a);
        for (int i = 0; i < 3; ++i) {
            const float t1 = a[i] + 1.0f;
            const float t2 = 0.5f * t1;
            U[i] = -bound + (double)t2 * (bound - (-bound));
        }
        apply_fault(fault, U, cmd);
        memcpy(xo, X, sizeof(xo));
        plant_step(variant, X, cmd);
        const double e0 = ref_deg(levels, starts, t, t <= t_max ? theta_trim : 0.0, smooth_w) * DEG2RAD - xo[7];   /* Const(0, t_max, trim) */
        const double e1 = ref_deg(levels + 6, starts + 6, t, 0.0, smooth_w) * DEG2RAD - xo[6];
        const double e2 = 0.0 - xo[5];
        const double c0 = fabs(fmin(fmax(k_err * e0, -1.0), 1.0)), c1 = fabs(fmin(fmax(k_err * e1, -1.0), 1.0));
        const double c2 = fabs(fmin(fmax(k_err4 * e2, -1.0), 1.0));
        double reward = -((c0 + c1) + c2) / 3.0;
        const int done = (t >= t_max) || (fabs(xo[7]) > max_theta) || (fabs(xo[6]) > max_phi) || (xo[9] < 50.0);
        if (done) reward += (-1.0 / 0.01) * (t_max - t) * 2.0;
        ret += reward;
        obs[0] = (float)e0; obs[1] = (float)e1; obs[2] = (float)e2;
        obs[3] = (float)xo[0]; obs[4] = (float)xo[1]; obs[5] = (float)xo[2]; obs[6] = (float)xo[4];
        t += 0.01;
        ++k;
        if (done) break;
    }
    *steps = k;
    return ret;
}

/* pop x n_envs episodes, OpenMP over trajectories; returns total executed steps */
long oracle_population(const float* genomes, int pop, int P, int S, int A, int H, int L, int act, const int* modes,
                       const double* levels, const double* starts, int n_envs, double t_max, double smooth_w, int horizon,
                       double* returns, int* steps, int actor_order)
{
    long total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int j = 0; j < pop * n_envs; ++j) {
        const int a = j / n_envs, e = j % n_envs;
        returns[j] = oracle_episode(genomes + (long)a * P, S, A, H, L, act, modes[e], levels + (long)e * 12, starts + (long)e * 12,
                                    t_max, smooth_w, horizon, steps + j, actor_order);
        total += steps[j];
    }
    return total;
}

/* the same for width-list actors (actor_order 2: index-order float32 sums + libm tanh) */
long oracle_population_wide(const float* genomes, int pop, int P, const int* widths, int nw, int act, const int* modes,
                            const double* levels, const double* starts, int n_envs, double t_max, double smooth_w, int horizon,
                            double* returns, int* steps)
{
    g_widths = widths; g_nw = nw;
    return oracle_population(genomes, pop, P, 7, 3, widths[0], 0, act, modes, levels, starts, n_envs, t_max, smooth_w, horizon, returns, steps, 2);
}

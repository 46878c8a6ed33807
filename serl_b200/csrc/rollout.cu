// K1 — fused population rollout for sm_100a (+ K6 smoothness metric, batched plant step).
//
// One lane = one (actor, env) trajectory; one warp = 32 envs of one actor; one CTA = 1-2 actors x 128 envs with their
// fp32 genomes (and the plant tables) staged once in shared memory.  Per step a warp runs, entirely on chip:
//   Actor.select_action  (base/core/genetic_agent.py:104-109; LayerNorm base/core/mod_utils.py:47-50)        fp32
//   CitationEnv.step     (envs/phlabenv.py:430-482: action scaling :62-73, fault shims envs/{be,jr,sa,se}/citation.py,
//                         reward :362-367, termination + penalty :391-399)                                fp64
//   native plant step    (envs/<variant>/_citation*.so step @0x6030: 6-stage Dormand-Prince ode5, h = 0.01, RHS
//                         generated from the binary by tools/lift -> csrc/gen/plant_rhs_{common,ice}.h)      fp64
// and accumulates the episodic return (base/core/agent.py:129).  HBM is touched only at episode start
// (genome, reference-signal parameters) and end (return, step count) unless a trace / action history is requested.
//
// Two actor implementations:
//   rollout_kernel_warp<H>  warp-autonomous: a warp never synchronises with other warps.  The MLP is a register-tiled
//                           GEMM inside the warp (lane = 1/4 of the output neurons x 4 envs, packed FFMA2, activations
//                           exchanged with warp shuffles, weights broadcast from shared memory).  (h in {32,64,72,96,128})
//   rollout_kernel_simple   every thread runs the whole MLP for its env (any h that fits); cross-check / fallback shape.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/serl_b200.h"
#include "common.cuh"

#ifdef PLANT_F32
typedef float real;      // experimental build: single-precision right-hand side, double-precision integrator state
#else
typedef double real;
#endif
// lookup tables: one blob (gen/plant_tables_blob.h) that the kernels stage into shared memory; the generated
// right-hand sides address it through the `plant_tab` pointer they are handed.
#define PLANT_TAB(name) (plant_tab + PT_OFF_##name)
#define PLANT_XARGS , const real* __restrict__ plant_tab
// ---- fast fp64 math for the device plant (<= ~1 ulp; the oracle keeps the reference's exact operations) -------
// division: 20-bit hardware reciprocal seed + two Newton steps + one residual correction (9 instructions instead of ~33)
__device__ __forceinline__ double plant_div_fast(double a, double b)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}
// sqrt: 20-bit rsqrt seed + two Newton steps + residual correction; zero / negative / non-finite go to the library
__device__ __forceinline__ double plant_sqrt_fast(double x)
{
    if (!(x > 1e-300 && x < 1e300)) return sqrt(x);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = y * fma(-hx * y, y, 1.5);
    y = y * fma(-hx * y, y, 1.5);
    const double s = x * y;
    return fma(fma(-s, s, x), 0.5 * y, s);
}
// sincos for |x| <= 8 (the plant's angles are bounded by the termination rule): two-term Cody-Waite reduction by pi/2
// and the fdlibm kernel polynomials; larger arguments fall back to the library.
__device__ __forceinline__ void plant_sincos_fast(double x, double* sp, double* cp)
{
    if (!(fabs(x) <= 8.0)) { sincos(x, sp, cp); return; }
    const double q = rint(x * 0.6366197723675814);
    double r = fma(-q, 1.5707963267948966, x);
    r = fma(-q, 6.123233995736766e-17, r);
    const double z = r * r;
    double ps = 1.58969099521155010221e-10;
    ps = fma(ps, z, -2.50507602534068634195e-08);
    ps = fma(ps, z, 2.75573137070700676789e-06);
    ps = fma(ps, z, -1.98412698298579493134e-04);
    ps = fma(ps, z, 8.33333333332248946124e-03);
    ps = fma(ps, z, -1.66666666666666324348e-01);
    const double s = fma(r * z, ps, r);
    double pc = -1.13596475577881948265e-11;
    pc = fma(pc, z, 2.08757232129817482790e-09);
    pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05);
    pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
    const double c = fma(z * z, pc, fma(-0.5, z, 1.0));
    const int n = (int)q & 3;
    const double s1 = (n & 1) ? c : s, c1 = (n & 1) ? s : c;
    *sp = (n & 2) ? -s1 : s1;
    *cp = ((n + 1) & 2) ? -c1 : c1;
}
__device__ __forceinline__ double plant_sin_fast(double x) { double s, c; plant_sincos_fast(x, &s, &c); return s; }
__device__ __forceinline__ double plant_cos_fast(double x) { double s, c; plant_sincos_fast(x, &s, &c); return c; }
#if defined(PLANT_F32)
// experimental build (`python -m serl_b200.build --f32`): fp32 right-hand side (BASELINE north_star: "fp32 ODE integration")
#define PLANT_GEN(f) PLANT_STR(gen_f32/f)
#define PLANT_DIV(a, b) ((a) / (b))
#define PLANT_SQRT sqrtf
#define PLANT_FABS fabsf
#define PLANT_SIN sinf
#define PLANT_COS cosf
#define PLANT_SINCOS sincosf
#define PLANT_TAN tanf
#define PLANT_EXP expf
#define PLANT_LOG10 log10f
#define PLANT_POW powf
#elif defined(PLANT_EXACT)
// validation build (`python -m serl_b200.build --exact`): reference operation order, library math, no FMA contraction
#define PLANT_GEN(f) PLANT_STR(gen_exact/f)
#define PLANT_DIV(a, b) ((a) / (b))
#define PLANT_SQRT sqrt
#define PLANT_FABS fabs
#define PLANT_SIN sin
#define PLANT_COS cos
#define PLANT_SINCOS sincos
#else
#define PLANT_GEN(f) PLANT_STR(gen/f)
#define PLANT_DIV(a, b) plant_div_fast((a), (b))
#define PLANT_SQRT plant_sqrt_fast
#define PLANT_FABS fabs
#define PLANT_SIN plant_sin_fast
#define PLANT_COS plant_cos_fast
#define PLANT_SINCOS plant_sincos_fast
#endif
#ifndef PLANT_F32
#define PLANT_TAN tan
#define PLANT_EXP exp
#define PLANT_LOG10 log10
#define PLANT_POW pow
#endif
#define PLANT_STR(x) #x
#define PLANT_FN static __device__ __forceinline__
#include "plant_support.h"
#undef PLANT_FN
#define PLANT_FN static __device__ __noinline__
#include PLANT_GEN(plant_tables_blob.h)
#define PLANT_CONSTS(n) static __constant__ real plant_k[n]
#define PLANT_K(i) plant_k[i]
#include PLANT_GEN(plant_consts.h)
#define PLANT_IC(v) static __device__ const double plant_ic_unused_##v[19]
#define PLANT_IC_TABLE static __device__ const double plant_ic_table[SERL_PLANT_COUNT][19]
#define PLANT_PV_TABLE static __device__ const real plant_pv[SERL_PLANT_COUNT][PLANT_NPV]
#define PLANT_PV(k) plant_pvrow[k]
#include PLANT_GEN(plant_ic.h)
#include PLANT_GEN(plant_rhs_common.h)     // h2000_v90, cg, cg_for, h2000_v150, h10000_v90: one function + parameter rows
#include PLANT_GEN(plant_rhs_ice.h)        // structurally different build
#include PLANT_GEN(plant_rhs_nav.h)

#define ROLLOUT_THREADS 128
#define NX 19

// live continuous states of the plant: p q r V alpha beta phi theta | h | washout | N1 N1 N2 N2
// (psi, x_e, y_e never feed back and are integrated only for traces; Parameter_CSTATE(_g) are folded constants).
__device__ __forceinline__ void plant_rhs(int variant, const real* X, const real* U, real* xdot, const real* tab)
{
    if (variant == SERL_PLANT_ICE) plant_rhs_ice(X, U, xdot, tab);
    else plant_rhs_common(X, U, xdot, tab, plant_pv[variant]);
}

__device__ __forceinline__ const double* plant_ic(int variant) { return plant_ic_table[variant]; }

// Simulink fixed-step ode5 exactly as inlined in the reference's step(): stage states are
// y + (f0*hB0 + f1*hB1 + ...) with hB = h*B[s][j], summed left to right (zero coefficients included).
#define ODE5_B_INIT { \
        {1.0 / 5.0, 0, 0, 0, 0, 0}, \
        {3.0 / 40.0, 9.0 / 40.0, 0, 0, 0, 0}, \
        {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 0, 0, 0}, \
        {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, 0, 0}, \
        {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0, 0}, \
        {35.0 / 384.0, 0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}}
#define ODE5_LIVE_INIT {0, 1, 2, 3, 4, 5, 6, 7, 9, 12, 15, 16, 17, 18}
static __constant__ double c_ode5_B[6][6] = ODE5_B_INIT;      // dynamically indexed copy (trace path)
static __constant__ int c_ode5_live[14] = ODE5_LIVE_INIT;

// trace mode only: psi, x_e, y_e (rtX 8, 10, 11).  Their derivatives depend on the live states alone, so they are
// integrated after the fact with the same stage states, rebuilt from the stored stage derivatives f[6][NX].
__device__ __noinline__ void plant_step_nav(double* Xnav, const double* X0, const real (*f)[NX], const real* U, const real* tab)
{
    const double h = 0.01;
    const int NAV[3] = {8, 10, 11};
    double g[6][3], xs[NX];
    real xr[NX], xd[NX];
#pragma unroll 1
    for (int s = 0; s < 6; ++s) {
        for (int i = 0; i < NX; ++i) xs[i] = X0[i];
        if (s > 0) {
            for (int li = 0; li < 14; ++li) {
                const int i = c_ode5_live[li];
                double acc = (double)f[0][i] * (h * c_ode5_B[s - 1][0]);
                for (int j = 1; j < s; ++j) acc += (double)f[j][i] * (h * c_ode5_B[s - 1][j]);
                xs[i] = X0[i] + acc;
            }
            for (int q = 0; q < 3; ++q) {
                double acc = g[0][q] * (h * c_ode5_B[s - 1][0]);
                for (int j = 1; j < s; ++j) acc += g[j][q] * (h * c_ode5_B[s - 1][j]);
                xs[NAV[q]] = X0[NAV[q]] + acc;
            }
        }
        for (int i = 0; i < NX; ++i) xr[i] = (real)xs[i];
        plant_rhs_nav(xr, U, xd, tab);
        g[s][0] = (double)xd[8]; g[s][1] = (double)xd[10]; g[s][2] = (double)xd[11];
    }
    for (int q = 0; q < 3; ++q) {
        double acc = g[0][q] * (h * c_ode5_B[5][0]);
        for (int j = 1; j < 6; ++j) acc += g[j][q] * (h * c_ode5_B[5][j]);
        Xnav[q] = X0[NAV[q]] + acc;
    }
}

// Stage loop fully unrolled (every f[j][i] load of a stage is independent, h*B folds to constants); the right-hand
// sides are __noinline__ calls, so x and the stage derivatives f live in local memory (L1/L2-resident scratch: it is
// the source of the kernel's DRAM write-back traffic, see profiles/).  A rolled loop with the RHS inlined cuts that
// traffic 20x but runs 22 % slower (measured), so this form is kept.  The integrator state and the stage
// combinations are double in every build; `real` (the type of the right-hand side) is double unless PLANT_F32.
__device__ void plant_step(int variant, double* X, const double* U, const real* tab, bool nav = false)
{
    constexpr double h = 0.01;
    constexpr double B[6][6] = ODE5_B_INIT;
    constexpr int LIVE[14] = ODE5_LIVE_INIT;
    real f[6][NX], x[NX], u[3];
    u[0] = (real)U[0]; u[1] = (real)U[1]; u[2] = (real)U[2];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = (real)X[i];
    double xl[14];
#pragma unroll
    for (int s = 0; s < 6; ++s) {
        plant_rhs(variant, x, u, f[s], tab);
#pragma unroll
        for (int li = 0; li < 14; ++li) {
            const int i = LIVE[li];
            double acc = (double)f[0][i] * (h * B[s][0]);
#pragma unroll
            for (int j = 1; j <= s; ++j) acc += (double)f[j][i] * (h * B[s][j]);
            xl[li] = X[i] + acc;
            x[i] = (real)xl[li];
        }
    }
    if (nav) {
        double xn[3];
        plant_step_nav(xn, X, f, u, tab);
        X[8] = xn[0]; X[10] = xn[1]; X[11] = xn[2];
    }
#pragma unroll
    for (int li = 0; li < 14; ++li) X[LIVE[li]] = xl[li];
}

__device__ __forceinline__ float act_fn(int act, float x)
{
    if (act == SERL_ACT_TANH) return tanhf(x);
    if (act == SERL_ACT_ELU) return x > 0.f ? x : expm1f(x);
    return x > 0.f ? x : 0.01f * x;
}

// reference-signal value in degrees (serl_b200/refsig.py; recovered shape of signals.RandomizedCosineStepSequence)
__device__ __forceinline__ double ref_deg(const double* __restrict__ lv, const double* __restrict__ st, double t, double offset, double smooth_w)
{
    int k = 0;
#pragma unroll
    for (int j = 1; j < SERL_REF_BLOCKS; ++j)
        if (t >= st[j]) k = j;
    if (k == 0) return offset + lv[0];
    const double x = (t - st[k]) / smooth_w;
    if (x >= 1.0) return offset + lv[k];
    return offset + (lv[k - 1] + (lv[k] - lv[k - 1]) * (0.5 * (1.0 - cos(3.141592653589793 * x))));
}

// ---- per-trajectory environment (CitationEnv restated for one thread) --------------------------------------
struct RolloutArgs {
    const float* weights; int P; serl_actor_shape sh;
    const double* ref_levels; const double* ref_starts; const int* env_mode; int n_envs; int horizon;
    const float* action_noise;      // optional [pop, n_envs, horizon, 3]: clipped exploration noise (agent.py:90-93)
    double* returns; int* steps; double* trace;   // trace optional [pop, n_envs, horizon, SERL_TRACE_COLS]
    float* actions;                 // optional [pop, n_envs, horizon, 3] fp32: commanded deflection last_u (smoothness metric)
    int pop;
    double t_max;                   // episode length [s] (envs/phlabenv.py:181; 80 in evaluation mode :295-301)
    double smooth_w;                // width of the raised-cosine reference transitions [s] (t_max // 6)
};

struct Env {
    double X[NX];
    const real* tab;         // plant tables (shared or global memory)
    const double* ref_lv;    // this env's reference-signal levels / starts [2][SERL_REF_BLOCKS] (global, read per step)
    const double* ref_st;
    double t, ret, theta_trim;
    int variant, fault, k;
    bool done;
};

#define DEG2RAD 0.017453292519943295   // numpy deg2rad multiplier (pi/180)
#define RAD2DEG 57.29577951308232      // numpy rad2deg multiplier (180/pi)

__device__ __forceinline__ void apply_fault(int fault, const double* u, double* c)
{
    c[0] = u[0]; c[1] = u[1]; c[2] = u[2];
    if (fault == SERL_FAULT_BE) c[0] = u[0] * 0.3;                                       // envs/be/citation.py:71-75
    else if (fault == SERL_FAULT_JR) c[2] = 15 * 3.14159 / 180;                          // envs/jr/citation.py:71-75
    else if (fault == SERL_FAULT_SA) { const double b = 1.0 * DEG2RAD; c[1] = fmin(fmax(u[1], -b), b); }   // envs/sa :73-79
    else if (fault == SERL_FAULT_SE) { const double b = 2.5 * DEG2RAD; c[0] = fmin(fmax(u[0], -b), b); }   // envs/se :73-79
}

// reset(): initialize(), one zero-command step returns the initial state (phlabenv.py:401-428). obs = [0,0,0,p,q,r,alpha]
__device__ void env_reset(Env& e, const RolloutArgs& a, int env, float* obs)
{
    const int mode = a.env_mode[env];
    e.variant = mode & 0xff;
    e.fault = (mode >> 8) & 0xff;
    e.ref_lv = a.ref_levels + (size_t)env * 2 * SERL_REF_BLOCKS;
    e.ref_st = a.ref_starts + (size_t)env * 2 * SERL_REF_BLOCKS;
    const double* ic = plant_ic(e.variant);
#pragma unroll
    for (int i = 0; i < NX; ++i) e.X[i] = ic[i];
    obs[0] = obs[1] = obs[2] = 0.f;
    obs[3] = (float)e.X[0]; obs[4] = (float)e.X[1]; obs[5] = (float)e.X[2]; obs[6] = (float)e.X[4];
    e.theta_trim = e.X[7] * RAD2DEG;
    double U[3] = {0.0, 0.0, 0.0}, cmd[3];
    apply_fault(e.fault, U, cmd);
    plant_step(e.variant, e.X, cmd, e.tab, a.trace != nullptr);
    e.t = 0.0; e.ret = 0.0; e.k = 0; e.done = false;
}

// one CitationEnv.step (phlabenv.py:430-482) + the bookkeeping of Agent.evaluate (agent.py:85-118)
__device__ void env_step(Env& e, const RolloutArgs& ar, size_t traj, const float* a, float* obs)
{
    const double bound = 10.0 * DEG2RAD;                       // phlabenv.py:208
    const double max_theta = 60.0 * DEG2RAD, max_phi = 75.0 * DEG2RAD;
    const double k_err = 6.0 / 3.141592653589793;              // phlabenv.py:226-231
    const double k_err4 = k_err * 4.0;
    double U[3], cmd[3], act_d[3];
    if (ar.action_noise) {
        // action = clip(action + clipped_noise, -1, 1) in float64, then scale_action in float64 (agent.py:90-96)
        const float* nz = ar.action_noise + (traj * ar.horizon + e.k) * 3;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            act_d[i] = fmin(fmax((double)a[i] + (double)nz[i], -1.0), 1.0);
            U[i] = -bound + 0.5 * (act_d[i] + 1.0) * (bound - (-bound));
        }
    } else {
        // scale_action: low + 0.5*(a + 1.0)*(high - low) with a float32: (a + 1.0) and the halving round in fp32 (:72-73)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            act_d[i] = (double)a[i];
            const float t1 = a[i] + 1.0f;
            const float t2 = 0.5f * t1;
            U[i] = -bound + (double)t2 * (bound - (-bound));
        }
    }
    apply_fault(e.fault, U, cmd);
    double xo[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) xo[i] = e.X[i];
    plant_step(e.variant, e.X, cmd, e.tab, ar.trace != nullptr);

    const double t = e.t;
    const double r_th = ref_deg(e.ref_lv, e.ref_st, t, e.theta_trim, ar.smooth_w) * DEG2RAD;
    const double r_ph = ref_deg(e.ref_lv + SERL_REF_BLOCKS, e.ref_st + SERL_REF_BLOCKS, t, 0.0, ar.smooth_w) * DEG2RAD;
    const double e0 = r_th - xo[7], e1 = r_ph - xo[6], e2 = 0.0 - xo[5];
    const double c0 = fabs(fmin(fmax(k_err * e0, -1.0), 1.0));
    const double c1 = fabs(fmin(fmax(k_err * e1, -1.0), 1.0));
    const double c2 = fabs(fmin(fmax(k_err4 * e2, -1.0), 1.0));
    double reward = -((c0 + c1) + c2) / 3.0;
    const bool done = (t >= ar.t_max) || (fabs(xo[7]) > max_theta) || (fabs(xo[6]) > max_phi) || (xo[9] < 50.0);
    if (done) reward += (-1.0 / 0.01) * (ar.t_max - t) * 2.0;  // check_bounds penalty (:391-399)
    e.ret += reward;
    if (ar.actions) {
        float* au = ar.actions + (traj * ar.horizon + e.k) * 3;
        au[0] = (float)U[0]; au[1] = (float)U[1]; au[2] = (float)U[2];
    }
    if (ar.trace) {
        double* tr = ar.trace + (traj * ar.horizon + e.k) * SERL_TRACE_COLS;
#pragma unroll
        for (int i = 0; i < 12; ++i) tr[i] = xo[i];
        tr[12] = U[0]; tr[13] = U[1]; tr[14] = U[2];
        tr[15] = reward;
        tr[16] = act_d[0]; tr[17] = act_d[1]; tr[18] = act_d[2];
        tr[19] = e0; tr[20] = e1; tr[21] = e2;
    }
    obs[0] = (float)e0; obs[1] = (float)e1; obs[2] = (float)e2;
    obs[3] = (float)xo[0]; obs[4] = (float)xo[1]; obs[5] = (float)xo[2]; obs[6] = (float)xo[4];
    e.t = t + 0.01;
    e.k += 1;
    e.done = done || (e.k >= ar.horizon);
}

// ---- simple actor: every thread evaluates the whole MLP for its own observation -------------------------
__device__ void actor_forward_simple(const float* __restrict__ w, const serl_actor_shape sh, float* bufA, float* bufB,
                                     int tid, int nthr, const float* obs, float* action)
{
    const int S = sh.state_dim, A = sh.action_dim, H = sh.hidden, L = sh.num_layers;
    const float* p = w;
    for (int j = 0; j < H; ++j) {
        float acc = 0.f;
        for (int i = 0; i < S; ++i) acc = fmaf(p[j * S + i], obs[i], acc);
        acc += p[H * S + j];
        bufA[j * nthr + tid] = act_fn(sh.activation, acc);
    }
    p += H * S + H;
    float* in = bufA;
    float* out = bufB;
    for (int l = 0; l < L; ++l) {
        const float* W = p;
        const float* b = p + H * H;
        const float* gamma = b + H;
        const float* beta = gamma + H;
        float sum = 0.f;
        for (int j = 0; j < H; ++j) {
            float acc = 0.f;
            for (int i = 0; i < H; ++i) acc = fmaf(W[j * H + i], in[i * nthr + tid], acc);
            acc += b[j];
            out[j * nthr + tid] = acc;
            sum += acc;
        }
        const float mean = sum / (float)H;
        float ss = 0.f;
        for (int j = 0; j < H; ++j) {
            const float d = out[j * nthr + tid] - mean;
            ss = fmaf(d, d, ss);
        }
        const float den = sqrtf(ss / (float)(H - 1)) + 1e-6f;
        for (int j = 0; j < H; ++j) {
            const float d = out[j * nthr + tid] - mean;
            out[j * nthr + tid] = act_fn(sh.activation, gamma[j] * d / den + beta[j]);
        }
        p += H * H + 3 * H;
        float* t = in; in = out; out = t;
    }
    for (int j = 0; j < A; ++j) {
        float acc = 0.f;
        for (int i = 0; i < H; ++i) acc = fmaf(p[j * H + i], in[i * nthr + tid], acc);
        acc += p[A * H + j];
        action[j] = tanhf(acc);
    }
}

__global__ void __launch_bounds__(ROLLOUT_THREADS)
rollout_kernel_simple(RolloutArgs ar)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* w = reinterpret_cast<float*>(smem_raw);
    const int P4 = (ar.P + 3) & ~3;
    float* bufA = w + P4;
    float* bufB = bufA + ar.sh.hidden * ROLLOUT_THREADS;
    const int actor = blockIdx.y, tid = threadIdx.x;
    const int env = blockIdx.x * ROLLOUT_THREADS + tid;
    const float* gw = ar.weights + (size_t)actor * ar.P;
    for (int i = tid; i < ar.P; i += ROLLOUT_THREADS) w[i] = gw[i];
    __syncthreads();
    if (env >= ar.n_envs) return;
    Env e;
    e.tab = plant_tables_blob;
    float obs[7], a[3];
    env_reset(e, ar, env, obs);
    const size_t traj = (size_t)actor * ar.n_envs + env;
    while (!e.done) {
        actor_forward_simple(w, ar.sh, bufA, bufB, tid, ROLLOUT_THREADS, obs, a);
        env_step(e, ar, traj, a, obs);
    }
    ar.returns[traj] = e.ret;
    ar.steps[traj] = e.k;
}

// ---- warp-autonomous actor + env ---------------------------------------------------------------------------
// smem: plant tables [PT_TOTAL] f64 | per actor of the CTA: transposed weights
//       Wt0[S][H] b0[H] | L x { Wt[H][H] b[H] gamma[H] beta[H] } | Wo[A][H] bo[A]
// lane = (g = lane>>2, og = lane&3): output neurons og*TM .. og*TM+TM-1 of the envs 4g .. 4g+3 of this warp.
// The activation of neuron k for env 4g+c lives in lane (g, og = k/TM), register in[k%TM][c]; the next layer
// fetches it with one shuffle per (k, c).  Reductions over neurons are xor-butterflies over the two og bits, so
// the four lanes of a group hold bit-identical means / deviations.
// All MLP arithmetic uses the packed FP32 FMA of sm_100 (FFMA2: two IEEE-rn fmas per issued instruction); a float2
// register pair holds two consecutive output neurons (m, m+1) of one env, exactly what one LDS.64 of the transposed
// weight row delivers, and the activation of the source neuron is broadcast into both halves.
template <int H>
__device__ __forceinline__ void warp_layer(const float* __restrict__ Wt, const float2 (&in)[H / 8][4], float2 (&acc)[H / 8][4],
                                           int og, int lane)
{
    constexpr int TM = H / 4, TM2 = H / 8;
#pragma unroll
    for (int m = 0; m < TM2; ++m)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[m][c] = make_float2(0.f, 0.f);
    const int gbase = lane & ~3;
#pragma unroll 1
    for (int so = 0; so < 4; ++so) {
        const float* wrow = Wt + (size_t)(so * TM) * H + og * TM;
#pragma unroll
        for (int m = 0; m < TM; ++m) {
            // activation of source neuron k = so*TM + m for the four envs of this group
            const float src0 = (m & 1) ? in[m >> 1][0].y : in[m >> 1][0].x;
            const float src1 = (m & 1) ? in[m >> 1][1].y : in[m >> 1][1].x;
            const float src2 = (m & 1) ? in[m >> 1][2].y : in[m >> 1][2].x;
            const float src3 = (m & 1) ? in[m >> 1][3].y : in[m >> 1][3].x;
            const float a0 = __shfl_sync(0xffffffffu, src0, gbase + so);
            const float a1 = __shfl_sync(0xffffffffu, src1, gbase + so);
            const float a2 = __shfl_sync(0xffffffffu, src2, gbase + so);
            const float a3 = __shfl_sync(0xffffffffu, src3, gbase + so);
            const float2 b0 = make_float2(a0, a0), b1 = make_float2(a1, a1), b2 = make_float2(a2, a2), b3 = make_float2(a3, a3);
            const float2* wp = reinterpret_cast<const float2*>(wrow + m * H);
#pragma unroll
            for (int m2 = 0; m2 < TM2; ++m2) {
                const float2 w2 = wp[m2];
                acc[m2][0] = __ffma2_rn(w2, b0, acc[m2][0]);
                acc[m2][1] = __ffma2_rn(w2, b1, acc[m2][1]);
                acc[m2][2] = __ffma2_rn(w2, b2, acc[m2][2]);
                acc[m2][3] = __ffma2_rn(w2, b3, acc[m2][3]);
            }
        }
    }
}

__device__ __forceinline__ float group_sum(float v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

template <int H, int ACT>
__device__ __noinline__ void actor_forward_warp(const float* __restrict__ w, int L, int lane, const float* obs, float* action)
{
    constexpr int actfn = ACT;
    constexpr int TM = H / 4, TM2 = H / 8;
    constexpr int S = 7, A = 3;
    const int og = lane & 3, gbase = lane & ~3;
    const float* Wt0 = w;
    const float* b0 = Wt0 + S * H;
    const float* hid = b0 + H;
    const float* Wo = hid + (size_t)L * (H * H + 3 * H);
    const float* bo = Wo + A * H;
    float2 in[TM2][4], acc[TM2][4];
    // input layer: observation of env 4g+c lives in lane gbase+c
#pragma unroll
    for (int m = 0; m < TM2; ++m)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[m][c] = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < S; ++k) {
        const float a0 = __shfl_sync(0xffffffffu, obs[k], gbase + 0);
        const float a1 = __shfl_sync(0xffffffffu, obs[k], gbase + 1);
        const float a2 = __shfl_sync(0xffffffffu, obs[k], gbase + 2);
        const float a3 = __shfl_sync(0xffffffffu, obs[k], gbase + 3);
        const float2 b0v = make_float2(a0, a0), b1v = make_float2(a1, a1), b2v = make_float2(a2, a2), b3v = make_float2(a3, a3);
        const float2* wp = reinterpret_cast<const float2*>(Wt0 + k * H + og * TM);
#pragma unroll
        for (int m2 = 0; m2 < TM2; ++m2) {
            const float2 w2 = wp[m2];
            acc[m2][0] = __ffma2_rn(w2, b0v, acc[m2][0]); acc[m2][1] = __ffma2_rn(w2, b1v, acc[m2][1]);
            acc[m2][2] = __ffma2_rn(w2, b2v, acc[m2][2]); acc[m2][3] = __ffma2_rn(w2, b3v, acc[m2][3]);
        }
    }
#pragma unroll
    for (int m2 = 0; m2 < TM2; ++m2) {
        const float2 b = *reinterpret_cast<const float2*>(b0 + og * TM + 2 * m2);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            in[m2][c].x = act_fn(actfn, acc[m2][c].x + b.x);
            in[m2][c].y = act_fn(actfn, acc[m2][c].y + b.y);
        }
    }
    // hidden layers: Linear -> LayerNorm (unbiased std, eps on std; mod_utils.py:47-50) -> activation
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
        const float* Wt = hid + (size_t)l * (H * H + 3 * H);
        const float* bb = Wt + H * H;
        const float* gamma = bb + H;
        const float* beta = gamma + H;
        warp_layer<H>(Wt, in, acc, og, lane);
        float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int m2 = 0; m2 < TM2; ++m2) {
            const float2 b = *reinterpret_cast<const float2*>(bb + og * TM + 2 * m2);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                acc[m2][c].x += b.x; s[c] += acc[m2][c].x;
                acc[m2][c].y += b.y; s[c] += acc[m2][c].y;
            }
        }
        float mean[4], den[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) mean[c] = group_sum(s[c]) / (float)H;
        float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int m2 = 0; m2 < TM2; ++m2)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                acc[m2][c].x -= mean[c]; q[c] = fmaf(acc[m2][c].x, acc[m2][c].x, q[c]);
                acc[m2][c].y -= mean[c]; q[c] = fmaf(acc[m2][c].y, acc[m2][c].y, q[c]);
            }
        // gamma * (x - mean) / (std + eps) + beta with one reciprocal per env instead of H/4 divisions per lane
        // (<= 1 ulp from the reference's division, the same order as the summation-order differences of the GEMM)
#pragma unroll
        for (int c = 0; c < 4; ++c) den[c] = 1.0f / (sqrtf(group_sum(q[c]) / (float)(H - 1)) + 1e-6f);
#pragma unroll
        for (int m2 = 0; m2 < TM2; ++m2) {
            const float2 g = *reinterpret_cast<const float2*>(gamma + og * TM + 2 * m2);
            const float2 be = *reinterpret_cast<const float2*>(beta + og * TM + 2 * m2);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                in[m2][c].x = act_fn(actfn, g.x * acc[m2][c].x * den[c] + be.x);
                in[m2][c].y = act_fn(actfn, g.y * acc[m2][c].y * den[c] + be.y);
            }
        }
    }
    // output layer: partial dot products over this lane's neurons, reduced over the group; lane og keeps env 4g+og
#pragma unroll
    for (int j = 0; j < A; ++j) {
        float p[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int m2 = 0; m2 < TM2; ++m2) {
            const float2 wv = *reinterpret_cast<const float2*>(Wo + j * H + og * TM + 2 * m2);
#pragma unroll
            for (int c = 0; c < 4; ++c) { p[c] = fmaf(wv.x, in[m2][c].x, p[c]); p[c] = fmaf(wv.y, in[m2][c].y, p[c]); }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) p[c] = group_sum(p[c]);
        const float mine = og == 0 ? p[0] : (og == 1 ? p[1] : (og == 2 ? p[2] : p[3]));
        action[j] = tanhf(mine + bo[j]);
    }
}

// TABS: plant tables staged in shared memory (true) or read from global memory through L1 (false: h = 128, whose
// 207 KB genome leaves no room for them).
template <int H, int APC, bool TABS>
__global__ void __launch_bounds__(ROLLOUT_THREADS * APC, 1)
rollout_kernel_warp(RolloutArgs ar)
{
    constexpr int S = 7;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real* tab_s = reinterpret_cast<real*>(smem_raw);
    float* wbase = reinterpret_cast<float*>(tab_s + (TABS ? PT_TOTAL : 0));
    const real* tab = TABS ? tab_s : plant_tables_blob;
    const int L = ar.sh.num_layers;
    const int P4 = (ar.P + 3) & ~3;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int al = warp >> 2;                               // actor slot of this warp inside the CTA
    const int actor = blockIdx.y * APC + al;
    const int env = blockIdx.x * ROLLOUT_THREADS + (warp & 3) * 32 + lane;
    if (TABS)
        for (int i = tid; i < PT_TOTAL; i += ROLLOUT_THREADS * APC) tab_s[i] = plant_tables_blob[i];
    // stage the genomes: parameters() order in HBM (row-major [out][in]) -> transposed [in][out] in smem
    for (int slot = 0; slot < APC; ++slot) {
        const int ga = blockIdx.y * APC + slot;
        if (ga >= ar.pop) break;
        const float* gw = ar.weights + (size_t)ga * ar.P;
        float* w = wbase + (size_t)slot * P4;
        float* Wt0 = w; float* b0 = Wt0 + S * H; float* hid = b0 + H;
        float* Wo = hid + (size_t)L * (H * H + 3 * H);
        for (int i = tid; i < ar.P; i += ROLLOUT_THREADS * APC) {
            const float v = gw[i];
            int r = i;
            if (r < S * H) { const int j = r / S, k = r % S; Wt0[k * H + j] = v; continue; }
            r -= S * H;
            if (r < H) { b0[r] = v; continue; }
            r -= H;
            const int per = H * H + 3 * H;
            if (r < L * per) {
                const int l = r / per, q = r % per;
                float* base = hid + (size_t)l * per;
                if (q < H * H) { const int j = q / H, k = q % H; base[k * H + j] = v; }
                else base[q] = v;
                continue;
            }
            r -= L * per;
            Wo[r] = v;      // Wo [A][H] then bo[A], contiguous
        }
    }
    __syncthreads();
    if (actor >= ar.pop) return;
    const float* w = wbase + (size_t)al * P4;
    Env e;
    e.tab = tab;
    float obs[7], a[3];
    const bool valid = env < ar.n_envs && actor < ar.pop;
    if (valid) env_reset(e, ar, env, obs);
    else { e.done = true; e.k = 0; e.ret = 0.0;
#pragma unroll
        for (int i = 0; i < 7; ++i) obs[i] = 0.f; }
    const size_t traj = valid ? (size_t)actor * ar.n_envs + env : 0;
    const int actfn = ar.sh.activation;
    while (__any_sync(0xffffffffu, !e.done)) {
        // one instantiation per activation: the choice is compiled into the 4 x h/4 activation calls of every layer
        if (actfn == SERL_ACT_TANH) actor_forward_warp<H, SERL_ACT_TANH>(w, L, lane, obs, a);
        else if (actfn == SERL_ACT_ELU) actor_forward_warp<H, SERL_ACT_ELU>(w, L, lane, obs, a);
        else actor_forward_warp<H, SERL_ACT_LEAKY_RELU>(w, L, lane, obs, a);
        if (!e.done) env_step(e, ar, traj, a, obs);
    }
    if (valid) {
        ar.returns[traj] = e.ret;
        ar.steps[traj] = e.k;
    }
}

// fitness[a] = mean over envs of returns[a, :]  (base/core/agent.py:245, np.mean over the evaluation axis)
__global__ void fitness_mean_kernel(const double* __restrict__ returns, int pop, int n_envs, double* __restrict__ fitness)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= pop) return;
    double s = 0.0;
    for (int e = 0; e < n_envs; ++e) s += returns[(size_t)a * n_envs + e];
    fitness[a] = s / (double)n_envs;
}

// batched native-plant step: X[n,19] advanced in place by one major step with command cmd[n,3] (inputs 3..9 are 0)
__global__ void plant_step_kernel(double* __restrict__ X, const double* __restrict__ cmd, const int* __restrict__ variant, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x[NX], u[3];
#pragma unroll
    for (int k = 0; k < NX; ++k) x[k] = X[(size_t)i * NX + k];
    u[0] = cmd[3 * i]; u[1] = cmd[3 * i + 1]; u[2] = cmd[3 * i + 2];
    plant_step(variant[i] & 0xff, x, u, plant_tables_blob);
#pragma unroll
    for (int k = 0; k < NX; ++k) X[(size_t)i * NX + k] = x[k];
}

__global__ void plant_ic_kernel(double* __restrict__ X, const int* __restrict__ variant, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* ic = plant_ic(variant[i] & 0xff);
    for (int k = 0; k < NX; ++k) X[(size_t)i * NX + k] = ic[k];
}

extern "C" int serl_plant_init(double* d_X, const int32_t* d_variant, int32_t n, void* stream)
{
    if (!d_X || !d_variant || n <= 0) return serl_fail(SERL_ERR_ARG, "serl_plant_init: bad argument");
    plant_ic_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_X, d_variant, n);
    serl_count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SERL_OK : serl_fail_cuda(e, "plant_ic_kernel");
}

extern "C" int serl_plant_step(double* d_X, const double* d_cmd, const int32_t* d_variant, int32_t n, void* stream)
{
    if (!d_X || !d_cmd || !d_variant || n <= 0) return serl_fail(SERL_ERR_ARG, "serl_plant_step: bad argument");
    plant_step_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(d_X, d_cmd, d_variant, n);
    serl_count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SERL_OK : serl_fail_cuda(e, "plant_step_kernel");
}

// ---- K6: action-smoothness metric (base/core/utils.py:82-120) --------------------------------------------
// One CTA per trajectory: direct DFT of the three actuator signals over the executed steps N (N = 2001 is 3*23*29,
// no radix-2 structure; 12 M fp32 MACs per trajectory), frequency-weighted power summed in double:
//   S = sum_i sum_{k=1}^{N/2-1} f_k |Y_i[k]|^2 dt * 2/N,  f = linspace(dt, 1/(2dt), N/2-1),  result = -sqrt(S)*100*(80/(N dt)).
__global__ void __launch_bounds__(256)
smoothness_kernel(const float* __restrict__ actions, const int* __restrict__ steps, int horizon, double dt, double* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int traj = blockIdx.x;
    const int N = steps[traj];
    const int M = N / 2 - 1;
    if (M <= 0) { if (threadIdx.x == 0) out[traj] = -0.0; return; }
    float2* tw = reinterpret_cast<float2*>(sm_raw);            // [N] (cos, sin)(2 pi j / N)
    float* y0 = reinterpret_cast<float*>(tw + horizon);       // [3][N]
    float* y1 = y0 + horizon;
    float* y2 = y1 + horizon;
    const float* a = actions + (size_t)traj * horizon * 3;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float sv, cv;
        sincospif(2.0f * (float)n / (float)N, &sv, &cv);
        tw[n] = make_float2(cv, sv);
        y0[n] = a[3 * n]; y1[n] = a[3 * n + 1]; y2[n] = a[3 * n + 2];
    }
    __syncthreads();
    const double fstep = M > 1 ? (1.0 / (2.0 * dt) - dt) / (double)(M - 1) : 0.0;
    double acc = 0.0;
    // four frequencies per thread and pass: every y[n] broadcast load feeds 8 fmas per signal
    for (int kb = 1 + 4 * threadIdx.x; kb <= M; kb += 4 * blockDim.x) {
        float re[4][3], im[4][3];
        int idx[4], kk[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            kk[q] = (kb + q <= M) ? kb + q : 0;        // k = 0 is a harmless dummy (weight 0 below)
            idx[q] = 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) { re[q][c] = 0.f; im[q][c] = 0.f; }
        }
        for (int n = 0; n < N; ++n) {
            const float v0 = y0[n], v1 = y1[n], v2 = y2[n];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 w = tw[idx[q]];
                re[q][0] = fmaf(v0, w.x, re[q][0]); im[q][0] = fmaf(v0, w.y, im[q][0]);
                re[q][1] = fmaf(v1, w.x, re[q][1]); im[q][1] = fmaf(v1, w.y, im[q][1]);
                re[q][2] = fmaf(v2, w.x, re[q][2]); im[q][2] = fmaf(v2, w.y, im[q][2]);
                idx[q] += kk[q];
                if (idx[q] >= N) idx[q] -= N;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (kk[q] == 0) continue;
            double p = 0.0;
#pragma unroll
            for (int c = 0; c < 3; ++c) p += (double)re[q][c] * re[q][c] + (double)im[q][c] * im[q][c];
            acc += (dt + (double)(kk[q] - 1) * fstep) * p;
        }
    }
    __shared__ double red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double S = red[0] * dt * 2.0 / (double)N;
        out[traj] = -(sqrt(S) * 100.0 * (80.0 / ((double)N * dt)));
    }
}

extern "C" int serl_smoothness(const float* d_actions, const int32_t* d_steps, int32_t n_traj, int32_t horizon, double dt,
                               double* d_out, void* stream)
{
    if (!d_actions || !d_steps || !d_out || n_traj <= 0 || horizon <= 0) return serl_fail(SERL_ERR_ARG, "serl_smoothness: bad argument");
    const size_t smem = (size_t)horizon * (8 + 12);
    if (smem > 200 * 1024) return serl_fail(SERL_ERR_UNSUPPORTED, "serl_smoothness: horizon too long for the shared-memory DFT");
    cudaError_t e = cudaFuncSetAttribute(smoothness_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return serl_fail_cuda(e, "cudaFuncSetAttribute(smoothness)");
    smoothness_kernel<<<n_traj, 256, smem, (cudaStream_t)stream>>>(d_actions, d_steps, horizon, dt, d_out);
    serl_count_launch();
    e = cudaGetLastError();
    return e == cudaSuccess ? SERL_OK : serl_fail_cuda(e, "smoothness_kernel");
}

extern "C" int64_t serl_actor_num_params(const serl_actor_shape* s)
{
    if (!s) return -1;
    const int64_t S = s->state_dim, A = s->action_dim, H = s->hidden, L = s->num_layers;
    return S * H + H + L * (H * H + 3 * H) + H * A + A;
}

static int g_force_simple = -1;

template <int H, int APC, bool TABS>
static cudaError_t launch_warp(const RolloutArgs& ar, cudaStream_t s)
{
    const int P4 = (ar.P + 3) & ~3;
    const size_t smem = (TABS ? (size_t)PT_TOTAL * sizeof(real) : 0) + (size_t)APC * P4 * 4;
    cudaError_t e = cudaFuncSetAttribute(rollout_kernel_warp<H, APC, TABS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((ar.n_envs + ROLLOUT_THREADS - 1) / ROLLOUT_THREADS, (ar.pop + APC - 1) / APC);
    rollout_kernel_warp<H, APC, TABS><<<grid, ROLLOUT_THREADS * APC, smem, s>>>(ar);
    return cudaGetLastError();
}

static int rollout_impl(const float* d_weights, int32_t pop, const serl_actor_shape* shape,
                        const double* d_ref_levels, const double* d_ref_starts, const int32_t* d_env_mode,
                        int32_t n_envs, int32_t horizon, const float* d_action_noise,
                        double* d_returns, int32_t* d_steps, double* d_fitness, double* d_trace, float* d_actions,
                        double t_max, double smooth_w, void* stream);

extern "C" int serl_rollout(const float* d_weights, int32_t pop, const serl_actor_shape* shape,
                            const double* d_ref_levels, const double* d_ref_starts, const int32_t* d_env_mode,
                            int32_t n_envs, int32_t horizon, const float* d_action_noise,
                            double* d_returns, int32_t* d_steps, double* d_fitness, double* d_trace, float* d_actions, void* stream)
{
    return rollout_impl(d_weights, pop, shape, d_ref_levels, d_ref_starts, d_env_mode, n_envs, horizon, d_action_noise,
                        d_returns, d_steps, d_fitness, d_trace, d_actions, 20.0, 3.0, stream);
}

extern "C" int serl_rollout_eval(const float* d_weights, int32_t pop, const serl_actor_shape* shape,
                                 const double* d_ref_levels, const double* d_ref_starts, const int32_t* d_env_mode,
                                 int32_t n_envs, int32_t horizon, const float* d_action_noise,
                                 double* d_returns, int32_t* d_steps, double* d_fitness, double* d_trace, float* d_actions,
                                 double t_max, double smooth_width, void* stream)
{
    if (!(t_max > 0.0) || !(smooth_width > 0.0)) return serl_fail(SERL_ERR_ARG, "serl_rollout_eval: t_max and smooth_width must be > 0");
    return rollout_impl(d_weights, pop, shape, d_ref_levels, d_ref_starts, d_env_mode, n_envs, horizon, d_action_noise,
                        d_returns, d_steps, d_fitness, d_trace, d_actions, t_max, smooth_width, stream);
}

static int rollout_impl(const float* d_weights, int32_t pop, const serl_actor_shape* shape,
                        const double* d_ref_levels, const double* d_ref_starts, const int32_t* d_env_mode,
                        int32_t n_envs, int32_t horizon, const float* d_action_noise,
                        double* d_returns, int32_t* d_steps, double* d_fitness, double* d_trace, float* d_actions,
                        double t_max, double smooth_w, void* stream)
{
    if (!d_weights || !shape || !d_ref_levels || !d_ref_starts || !d_env_mode || !d_returns || !d_steps)
        return serl_fail(SERL_ERR_ARG, "serl_rollout: null pointer argument");
    if (pop <= 0 || n_envs <= 0 || horizon <= 0) return serl_fail(SERL_ERR_ARG, "serl_rollout: pop, n_envs, horizon must be > 0");
    if (pop > 65535) return serl_fail(SERL_ERR_ARG, "serl_rollout: pop must be <= 65535 per call");
    if (shape->state_dim != 7 || shape->action_dim != 3)
        return serl_fail(SERL_ERR_ARG, "serl_rollout: PH-LAB attitude task needs state_dim=7, action_dim=3");
    if (shape->hidden < 2 || shape->hidden > 256 || shape->num_layers < 0 || shape->activation < 0 || shape->activation > 2)
        return serl_fail(SERL_ERR_ARG, "serl_rollout: unsupported actor shape");
    if (g_force_simple < 0) {
        const char* v = getenv("SERL_ROLLOUT_IMPL");
        g_force_simple = (v && strcmp(v, "simple") == 0) ? 1 : 0;
    }
    cudaStream_t s = (cudaStream_t)stream;
    RolloutArgs ar;
    ar.weights = d_weights; ar.P = (int)serl_actor_num_params(shape); ar.sh = *shape;
    ar.ref_levels = d_ref_levels; ar.ref_starts = d_ref_starts; ar.env_mode = d_env_mode; ar.n_envs = n_envs; ar.horizon = horizon;
    ar.action_noise = d_action_noise; ar.returns = d_returns; ar.steps = d_steps; ar.trace = d_trace; ar.actions = d_actions; ar.pop = pop; ar.t_max = t_max; ar.smooth_w = smooth_w;
    const int P4 = (ar.P + 3) & ~3;
    const int H = shape->hidden;
    dim3 grid((n_envs + ROLLOUT_THREADS - 1) / ROLLOUT_THREADS, pop);
    cudaError_t e;
    const bool warp_ok = !g_force_simple && (H == 32 || H == 64 || H == 72 || H == 96 || H == 128) && (size_t)P4 * 4 <= 227 * 1024;
    if (warp_ok) {
        // two actors per CTA (8 autonomous warps) when two genomes + the plant tables fit in shared memory; one actor
        // with the tables in shared memory when that fits; else (h = 128) one actor and the tables through L1
        // ... and when that still leaves at least one CTA per SM: a small population spreads over more SMs with one
        // actor per CTA (4 warps each) instead of filling half as many SMs with 8 warps
        static int num_sms = 0;
        if (num_sms == 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
        }
        const long ctas2 = (long)((pop + 1) / 2) * ((n_envs + ROLLOUT_THREADS - 1) / ROLLOUT_THREADS);
        const bool two = (size_t)PT_TOTAL * sizeof(real) + 2ull * P4 * 4 <= 227 * 1024 && pop > 1 && ctas2 >= num_sms;
        const bool tabs = (size_t)PT_TOTAL * sizeof(real) + (size_t)P4 * 4 <= 227 * 1024;
        static int apc_exp = -1;          // experiment knob (SERL_ROLLOUT_APC=3|4, h = 32 only): more resident warps per SM
        if (apc_exp < 0) { const char* v = getenv("SERL_ROLLOUT_APC"); apc_exp = v ? atoi(v) : 0; }
        if (H == 32 && apc_exp == 3 && pop > 2) e = launch_warp<32, 3, true>(ar, s);
        else if (H == 32 && apc_exp == 4 && pop > 3) e = launch_warp<32, 4, true>(ar, s);
        else if (H == 32) e = two ? launch_warp<32, 2, true>(ar, s) : launch_warp<32, 1, true>(ar, s);
        else if (H == 64) e = two ? launch_warp<64, 2, true>(ar, s) : launch_warp<64, 1, true>(ar, s);
#ifdef PLANT_F32
        else if (H == 72 && apc_exp == 3 && pop > 2) e = launch_warp<72, 3, true>(ar, s);   // float tables: three genomes fit
#endif
        else if (H == 72) e = two ? launch_warp<72, 2, true>(ar, s) : launch_warp<72, 1, true>(ar, s);
        else if (H == 96) e = launch_warp<96, 1, true>(ar, s);
        else e = tabs ? launch_warp<128, 1, true>(ar, s) : launch_warp<128, 1, false>(ar, s);
    } else {
        const size_t smem = (size_t)P4 * 4 + 2ull * H * ROLLOUT_THREADS * 4;
        if (smem > 227 * 1024) return serl_fail(SERL_ERR_UNSUPPORTED, "serl_rollout: genome + activations exceed 227 KB of shared memory");
        e = cudaFuncSetAttribute(rollout_kernel_simple, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return serl_fail_cuda(e, "cudaFuncSetAttribute(rollout)");
        rollout_kernel_simple<<<grid, ROLLOUT_THREADS, smem, s>>>(ar);
        e = cudaGetLastError();
    }
    serl_count_launch();
    if (e != cudaSuccess) return serl_fail_cuda(e, "rollout_kernel launch");
    if (d_fitness) {
        fitness_mean_kernel<<<(pop + 127) / 128, 128, 0, s>>>(d_returns, pop, n_envs, d_fitness);
        serl_count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return serl_fail_cuda(e, "fitness_mean_kernel launch");
    }
    return SERL_OK;
}

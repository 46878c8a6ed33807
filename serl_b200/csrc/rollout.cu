// K1 — fused population rollout for sm_100a.
//
// One thread = one (actor, env) trajectory; one CTA = one actor x up to ROLLOUT_THREADS envs, with that
// actor's fp32 genome staged once in shared memory.  Per step the thread runs, entirely on chip:
//   Actor.select_action  (base/core/genetic_agent.py:104-109; LayerNorm base/core/mod_utils.py:47-50)  fp32
//   CitationEnv.step     (envs/phlabenv.py:430-482: action scaling :62-73, fault shims envs/{be,jr,sa,se}/citation.py,
//                         reward :362-367, termination + penalty :391-399)                                fp64
//   native plant step    (envs/<variant>/_citation*.so step @0x6030: 6-stage Dormand-Prince ode5, h = 0.01, RHS
//                         generated from the binary by tools/lift -> csrc/gen/plant_rhs_<variant>.h)          fp64
// and accumulates the episodic return (base/core/agent.py:129).  HBM is touched only at episode start
// (genome, reference-signal parameters) and end (return, step count) unless traces are requested.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/serl_b200.h"
#include "common.cuh"

typedef double real;
#define PLANT_TABLE(name, n) static __device__ const double name[n]
#define PLANT_TAB(name) name
#define PLANT_IC(v) static __device__ const double plant_ic_##v[19]
#define PLANT_SQRT sqrt
#define PLANT_FABS fabs
#define PLANT_SIN sin
#define PLANT_COS cos
#define PLANT_TAN tan
#define PLANT_EXP exp
#define PLANT_LOG10 log10
#define PLANT_POW pow
#define PLANT_FN static __device__ __forceinline__
#include "plant_support.h"
#undef PLANT_FN
#define PLANT_FN static __device__ __noinline__
#include "gen/plant_tables.h"
#include "gen/plant_rhs_h2000_v90.h"
#include "gen/plant_rhs_ice.h"
#include "gen/plant_rhs_cg.h"
#include "gen/plant_rhs_cg_for.h"
#include "gen/plant_rhs_h2000_v150.h"
#include "gen/plant_rhs_h10000_v90.h"

#define ROLLOUT_THREADS 128
#define NX 19

// live continuous states of the plant: p q r V alpha beta phi theta | h | washout | N1 N1 N2 N2.
// psi, x_e, y_e never feed back; Parameter_CSTATE(_g) have zero derivative and are folded into the RHS.
__device__ __forceinline__ void plant_rhs(int variant, const double* X, const double* U, double* xdot)
{
    switch (variant) {
    case SERL_PLANT_ICE: plant_rhs_ice(X, U, xdot); break;
    case SERL_PLANT_CG: plant_rhs_cg(X, U, xdot); break;
    case SERL_PLANT_CG_FOR: plant_rhs_cg_for(X, U, xdot); break;
    case SERL_PLANT_H2000_V150: plant_rhs_h2000_v150(X, U, xdot); break;
    case SERL_PLANT_H10000_V90: plant_rhs_h10000_v90(X, U, xdot); break;
    default: plant_rhs_h2000_v90(X, U, xdot); break;
    }
}

__device__ __forceinline__ const double* plant_ic(int variant)
{
    switch (variant) {
    case SERL_PLANT_ICE: return plant_ic_ice;
    case SERL_PLANT_CG: return plant_ic_cg;
    case SERL_PLANT_CG_FOR: return plant_ic_cg_for;
    case SERL_PLANT_H2000_V150: return plant_ic_h2000_v150;
    case SERL_PLANT_H10000_V90: return plant_ic_h10000_v90;
    default: return plant_ic_h2000_v90;
    }
}

// Simulink fixed-step ode5 exactly as inlined in the reference's step(): stage states are
// y + (f0*hB0 + f1*hB1 + ...) with hB = h*B[s][j], summed left to right.
__device__ void plant_step(int variant, double* X, const double* U)
{
    const double h = 0.01;
    const double B[6][6] = {
        {1.0 / 5.0, 0, 0, 0, 0, 0},
        {3.0 / 40.0, 9.0 / 40.0, 0, 0, 0, 0},
        {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 0, 0, 0},
        {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, 0, 0},
        {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0, 0},
        {35.0 / 384.0, 0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}};
    double f[6][NX], x[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) { x[i] = X[i]; }
#pragma unroll 1
    for (int s = 0; s < 6; ++s) {
#pragma unroll
        for (int i = 0; i < NX; ++i) f[s][i] = 0.0;
        plant_rhs(variant, x, U, f[s]);
        for (int i = 0; i < NX; ++i) {
            double acc = f[0][i] * (h * B[s][0]);
            for (int j = 1; j <= s; ++j) acc += f[j][i] * (h * B[s][j]);
            x[i] = X[i] + acc;
        }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) X[i] = x[i];
}

__device__ __forceinline__ float act_fn(int act, float x)
{
    if (act == SERL_ACT_TANH) return tanhf(x);
    if (act == SERL_ACT_ELU) return x > 0.f ? x : expm1f(x);
    return x > 0.f ? x : 0.01f * x;
}

// reference-signal value in degrees (oracle/refsig.py: ref_value_deg)
__device__ __forceinline__ double ref_deg(const double* lv, const double* st, double t, double offset)
{
    int k = 0;
#pragma unroll
    for (int j = 1; j < SERL_REF_BLOCKS; ++j)
        if (t >= st[j]) k = j;
    if (k == 0) return offset + lv[0];
    const double x = (t - st[k]) / 3.0;
    if (x >= 1.0) return offset + lv[k];
    return offset + (lv[k - 1] + (lv[k] - lv[k - 1]) * (0.5 * (1.0 - cos(3.141592653589793 * x))));
}

// v1 actor forward: every thread evaluates the whole MLP for its own observation; weights are broadcast
// reads from shared memory, activations live in a per-thread column of shared memory (conflict-free).
__device__ void actor_forward(const float* __restrict__ w, const serl_actor_shape sh, float* bufA, float* bufB,
                              int tid, int nthr, const float* obs, float* action)
{
    const int S = sh.state_dim, A = sh.action_dim, H = sh.hidden, L = sh.num_layers;
    const float* p = w;
    // input layer
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
        const float stdv = sqrtf(ss / (float)(H - 1));
        const float den = stdv + 1e-6f;
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
rollout_kernel_v1(const float* __restrict__ weights, int P, serl_actor_shape sh,
                  const double* __restrict__ ref_levels, const double* __restrict__ ref_starts,
                  const int* __restrict__ env_mode, int n_envs, int horizon,
                  double* __restrict__ returns, int* __restrict__ steps,
                  double* __restrict__ trace_x, double* __restrict__ trace_u, double* __restrict__ trace_r)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* w = reinterpret_cast<float*>(smem_raw);
    const int P4 = (P + 3) & ~3;
    float* bufA = w + P4;
    float* bufB = bufA + sh.hidden * ROLLOUT_THREADS;

    const int actor = blockIdx.y;
    const int tid = threadIdx.x;
    const int env = blockIdx.x * ROLLOUT_THREADS + tid;
    const float* gw = weights + (size_t)actor * P;
    for (int i = tid; i < P; i += ROLLOUT_THREADS) w[i] = gw[i];
    __syncthreads();
    if (env >= n_envs) return;

    const int mode = env_mode[env];
    const int variant = mode & 0xff;
    const int fault = (mode >> 8) & 0xff;
    double lv[2][SERL_REF_BLOCKS], st[2][SERL_REF_BLOCKS];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < SERL_REF_BLOCKS; ++j) {
            lv[c][j] = ref_levels[((size_t)env * 2 + c) * SERL_REF_BLOCKS + j];
            st[c][j] = ref_starts[((size_t)env * 2 + c) * SERL_REF_BLOCKS + j];
        }

    const double DEG2RAD = 0.017453292519943295;   // numpy deg2rad multiplier (pi/180)
    const double RAD2DEG = 57.29577951308232;      // numpy rad2deg multiplier (180/pi)
    const double bound = 10.0 * DEG2RAD;           // envs/phlabenv.py:208
    const double max_theta = 60.0 * DEG2RAD, max_phi = 75.0 * DEG2RAD;
    const double k_err = 6.0 / 3.141592653589793;   // envs/phlabenv.py:226-231
    const double k_err4 = k_err * 4.0;

    double X[NX];
    const double* ic = plant_ic(variant);
#pragma unroll
    for (int i = 0; i < NX; ++i) X[i] = ic[i];

    // reset(): one zero-command step returns the initial state (phlabenv.py:409-413)
    double xo[12];
    double U[3] = {0.0, 0.0, 0.0};
    double cmd[3];
    auto apply_fault = [&](const double* u, double* c) {
        c[0] = u[0]; c[1] = u[1]; c[2] = u[2];
        if (fault == SERL_FAULT_BE) c[0] = u[0] * 0.3;
        else if (fault == SERL_FAULT_JR) c[2] = 15 * 3.14159 / 180;
        else if (fault == SERL_FAULT_SA) { const double b = 1.0 * DEG2RAD; c[1] = fmin(fmax(u[1], -b), b); }
        else if (fault == SERL_FAULT_SE) { const double b = 2.5 * DEG2RAD; c[0] = fmin(fmax(u[0], -b), b); }
    };
#pragma unroll
    for (int i = 0; i < 12; ++i) xo[i] = X[i];
    apply_fault(U, cmd);
    plant_step(variant, X, cmd);
    const double theta_trim = xo[7] * RAD2DEG;

    float obs[7] = {0.f, 0.f, 0.f, (float)xo[0], (float)xo[1], (float)xo[2], (float)xo[4]};
    double t = 0.0, ret = 0.0;
    int k = 0;
    const size_t traj = (size_t)actor * n_envs + env;
    for (; k < horizon; ++k) {
        float a[3];
        actor_forward(w, sh, bufA, bufB, tid, ROLLOUT_THREADS, obs, a);
        // scale_action: low + 0.5*(a + 1.0)*(high - low), (a + 1.0) and the halving in float32 (phlabenv.py:72-73)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float t1 = a[i] + 1.0f;
            const float t2 = 0.5f * t1;
            U[i] = -bound + (double)t2 * (bound - (-bound));
        }
        apply_fault(U, cmd);
#pragma unroll
        for (int i = 0; i < 12; ++i) xo[i] = X[i];
        plant_step(variant, X, cmd);

        const double r_th = ref_deg(lv[0], st[0], t, theta_trim) * DEG2RAD;
        const double r_ph = ref_deg(lv[1], st[1], t, 0.0) * DEG2RAD;
        const double e0 = r_th - xo[7], e1 = r_ph - xo[6], e2 = 0.0 - xo[5];
        const double c0 = fabs(fmin(fmax(k_err * e0, -1.0), 1.0));
        const double c1 = fabs(fmin(fmax(k_err * e1, -1.0), 1.0));
        const double c2 = fabs(fmin(fmax(k_err4 * e2, -1.0), 1.0));
        double reward = -((c0 + c1) + c2) / 3.0;
        const bool done = (t >= 20.0) || (fabs(xo[7]) > max_theta) || (fabs(xo[6]) > max_phi) || (xo[9] < 50.0);
        if (done) reward += (-1.0 / 0.01) * (20.0 - t) * 2.0;
        ret += reward;
        if (trace_x) {
            double* tx = trace_x + (traj * horizon + k) * 12;
#pragma unroll
            for (int i = 0; i < 12; ++i) tx[i] = xo[i];
        }
        if (trace_u) {
            double* tu = trace_u + (traj * horizon + k) * 3;
            tu[0] = U[0]; tu[1] = U[1]; tu[2] = U[2];
        }
        if (trace_r) trace_r[traj * horizon + k] = reward;
        obs[0] = (float)e0; obs[1] = (float)e1; obs[2] = (float)e2;
        obs[3] = (float)xo[0]; obs[4] = (float)xo[1]; obs[5] = (float)xo[2]; obs[6] = (float)xo[4];
        t += 0.01;
        if (done) { ++k; break; }
    }
    returns[traj] = ret;
    steps[traj] = k;
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

extern "C" int64_t serl_actor_num_params(const serl_actor_shape* s)
{
    if (!s) return -1;
    const int64_t S = s->state_dim, A = s->action_dim, H = s->hidden, L = s->num_layers;
    return S * H + H + L * (H * H + 3 * H) + H * A + A;
}

extern "C" int serl_rollout(const float* d_weights, int32_t pop, const serl_actor_shape* shape,
                            const double* d_ref_levels, const double* d_ref_starts, const int32_t* d_env_mode,
                            int32_t n_envs, int32_t horizon,
                            double* d_returns, int32_t* d_steps, double* d_fitness,
                            double* d_trace_x, double* d_trace_u, double* d_trace_r, void* stream)
{
    if (!d_weights || !shape || !d_ref_levels || !d_ref_starts || !d_env_mode || !d_returns || !d_steps)
        return serl_fail(SERL_ERR_ARG, "serl_rollout: null pointer argument");
    if (pop <= 0 || n_envs <= 0 || horizon <= 0) return serl_fail(SERL_ERR_ARG, "serl_rollout: pop, n_envs, horizon must be > 0");
    if (shape->state_dim != 7 || shape->action_dim != 3)
        return serl_fail(SERL_ERR_ARG, "serl_rollout: PH-LAB attitude task needs state_dim=7, action_dim=3");
    if (shape->hidden < 2 || shape->hidden > 256 || shape->num_layers < 0 || shape->activation < 0 || shape->activation > 2)
        return serl_fail(SERL_ERR_ARG, "serl_rollout: unsupported actor shape");
    cudaStream_t s = (cudaStream_t)stream;
    const int P = (int)serl_actor_num_params(shape);
    const int P4 = (P + 3) & ~3;
    const size_t smem = (size_t)P4 * 4 + 2ull * shape->hidden * ROLLOUT_THREADS * 4;
    if (smem > 227 * 1024) return serl_fail(SERL_ERR_UNSUPPORTED, "serl_rollout: genome + activations exceed 227 KB of shared memory");
    cudaError_t e = cudaFuncSetAttribute(rollout_kernel_v1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return serl_fail_cuda(e, "cudaFuncSetAttribute(rollout)");
    dim3 grid((n_envs + ROLLOUT_THREADS - 1) / ROLLOUT_THREADS, pop);
    rollout_kernel_v1<<<grid, ROLLOUT_THREADS, smem, s>>>(d_weights, P, *shape, d_ref_levels, d_ref_starts, d_env_mode,
                                                          n_envs, horizon, d_returns, d_steps, d_trace_x, d_trace_u, d_trace_r);
    serl_count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return serl_fail_cuda(e, "rollout_kernel launch");
    if (d_fitness) {
        fitness_mean_kernel<<<(pop + 127) / 128, 128, 0, s>>>(d_returns, pop, n_envs, d_fitness);
        serl_count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return serl_fail_cuda(e, "fitness_mean_kernel launch");
    }
    return SERL_OK;
}

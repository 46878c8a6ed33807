/* ORACLE — test infrastructure only.  The reference actor (base/core/genetic_agent.py:78-109, LayerNorm
 * base/core/mod_utils.py:47-50, activations :14-18) restated with the SUMMATION ORDER AND ACTIVATION ARITHMETIC OF THE
 * DEVICE KERNEL (serl_b200/csrc/rollout.cu actor_forward_warp, serl_b200/csrc/actor_math.cuh), so that a forward pass on
 * the CPU reproduces the GPU's actions bit for bit.  The reference (torch CPU) sums its GEMV in a library-defined order
 * and calls a vendor tanh; neither is specified by the reference, so a float32 forward pass is only defined up to its
 * last bits.  This file pins ONE legal choice — the product's:
 *
 *   dot products     acc = 0; for k = 0..n_in-1: acc = fmaf(W[j][k], in[k], acc); then x = acc + bias[j]
 *   LayerNorm sums   four partial sums over the contiguous quarter blocks of h/4 neurons (sequential inside a block),
 *                    combined as (q0 + q1) + (q2 + q3); mean = sum / h; ss likewise with fmaf(d, d, ss);
 *                    inv = 1 / (sqrtf(ss / (h-1)) + 1e-6f); y = fmaf(gamma * d, inv, beta)
 *   output layer     four quarter-block partial dot products (fmaf, sequential), combined the same way, + bias, tanh
 *   tanh / expm1     IEEE-only sequences (fma, add, mul, one correctly rounded division), <= ~2.5 ulp from the true value
 *
 * Every operation is a correctly rounded IEEE-754 single operation, so gcc (-ffp-contract=off, fmaf) and the GPU agree.
 * The torch-order actor (oracle/actor.py; episode.c actor_forward) stays the reference-order oracle; the difference
 * between the two is the reference's own float32 self-sensitivity and is reported, not used as a tolerance. */
/* the x86 FMA instruction when the host has it (function multiversioning), libm's correctly rounded fmaf() otherwise */
#if defined(__x86_64__) && defined(__GNUC__)
#define KO_FN __attribute__((target_clones("fma", "default")))
#else
#define KO_FN
#endif
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* tanh(x) = em1 / (em1 + 2), em1 = expm1(2|x|) = 2^n * expm1(2r) + (2^n - 1), |x| = n ln2/2 + r */
KO_FN float ko_tanhf(float x)
{
    const float a = x != x ? x : fminf(fabsf(x), 10.0f);            /* NaN stays NaN */
    const float m = fmaf(a, 0x1.715476p+1f, 12582912.0f);          /* 1.5*2^23 + rint(a * 2 log2 e) */
    const float n = m - 12582912.0f;
    float r = fmaf(n, -0x1.62ep-2f, a);
    r = fmaf(n, -0x1.0bfbe8p-16f, r);
    const float z = r * r;
    float p = 0x1.a12fbp-7f;
    p = fmaf(p, r, 0x1.6d4f3cp-5f);
    p = fmaf(p, r, 0x1.1110dcp-3f);
    p = fmaf(p, r, 0x1.5554ep-2f);
    p = fmaf(p, r, 0x1.555556p-1f);
    p = fmaf(p, r, 1.0f);
    const float h = fmaf(z, p, r);                                  /* expm1(2r) / 2 */
    const float s2 = bits2f(0x40000000u + (f2bits(m) << 23));       /* 2^(n+1) */
    const float sm1 = fmaf(s2, 0.5f, -1.0f);
    const float em1 = fmaf(s2, h, sm1);
    const float d = em1 + 2.0f;
    const float y = em1 / d;
    return copysignf(y, x);
}

/* expm1(x) for x <= 0 (ELU): 2^n * expm1(r) + (2^n - 1), x = n ln2 + r */
KO_FN float ko_expm1f_neg(float x)
{
    const float a = x != x ? x : fmaxf(x, -18.0f);
    const float m = fmaf(a, 0x1.715476p+0f, 12582912.0f);
    const float n = m - 12582912.0f;
    float r = fmaf(n, -0x1.62ep-1f, a);
    r = fmaf(n, -0x1.0bfbe8p-15f, r);
    const float z = r * r;
    float p = 0x1.a12fbp-13f;
    p = fmaf(p, r, 0x1.6d4f3cp-10f);
    p = fmaf(p, r, 0x1.1110dcp-7f);
    p = fmaf(p, r, 0x1.5554ep-5f);
    p = fmaf(p, r, 0x1.555556p-3f);
    p = fmaf(p, r, 0.5f);
    const float h = fmaf(z, p, r);
    const float s = bits2f(0x3f800000u + (f2bits(m) << 23));        /* 2^n, n <= 0 */
    const float sm1 = s - 1.0f;
    return fmaf(s, h, sm1);
}

static inline float ko_act(int act, float x)
{
    if (act == 0) return ko_tanhf(x);
    if (act == 1) return x > 0.f ? x : ko_expm1f_neg(x);
    return x > 0.f ? x : 0.01f * x;
}

static inline float quarter_combine(const float* q) { return (q[0] + q[1]) + (q[2] + q[3]); }

/* genome layout = nn.Module.parameters() order (oracle/actor.py) */
KO_FN void ko_actor_forward(const float* p, int S, int A, int H, int L, int act, const float* obs, float* action)
{
    float a[1024], b[1024];
    const int TM = (H + 3) / 4;          /* quarter blocks [g*TM, min((g+1)*TM, H)) */
    for (int j = 0; j < H; ++j) {
        float acc = 0.f;
        for (int i = 0; i < S; ++i) acc = fmaf(p[j * S + i], obs[i], acc);
        a[j] = ko_act(act, acc + p[H * S + j]);
    }
    p += H * S + H;
    for (int l = 0; l < L; ++l) {
        const float *W = p, *bias = p + H * H, *gamma = bias + H, *beta = gamma + H;
        float q[4];
        for (int g = 0; g < 4; ++g) {
            float s = 0.f;
            for (int j = g * TM; j < (g + 1) * TM && j < H; ++j) {
                float acc = 0.f;
                for (int i = 0; i < H; ++i) acc = fmaf(W[j * H + i], a[i], acc);
                b[j] = acc + bias[j];
                s += b[j];
            }
            q[g] = s;
        }
        const float mean = quarter_combine(q) / (float)H;
        for (int g = 0; g < 4; ++g) {
            float s = 0.f;
            for (int j = g * TM; j < (g + 1) * TM && j < H; ++j) {
                b[j] = b[j] - mean;
                s = fmaf(b[j], b[j], s);
            }
            q[g] = s;
        }
        const float inv = 1.0f / (sqrtf(quarter_combine(q) / (float)(H - 1)) + 1e-6f);
        for (int j = 0; j < H; ++j) a[j] = ko_act(act, fmaf(gamma[j] * b[j], inv, beta[j]));
        p += H * H + 3 * H;
    }
    for (int j = 0; j < A; ++j) {
        float q[4];
        for (int g = 0; g < 4; ++g) {
            float s = 0.f;
            for (int i = g * TM; i < (g + 1) * TM && i < H; ++i) s = fmaf(p[j * H + i], a[i], s);
            q[g] = s;
        }
        action[j] = ko_tanhf(quarter_combine(q) + p[A * H + j]);
    }
}

/* batched helper for tests: n observations [n,S] -> actions [n,A] */
void ko_actor_forward_batch(const float* p, int S, int A, int H, int L, int act, const float* obs, int n, float* actions)
{
    for (int i = 0; i < n; ++i) ko_actor_forward(p, S, A, H, L, act, obs + (long)i * S, actions + (long)i * A);
}

void ko_tanhf_batch(const float* x, int n, float* y) { for (int i = 0; i < n; ++i) y[i] = ko_tanhf(x[i]); }
void ko_expm1f_neg_batch(const float* x, int n, float* y) { for (int i = 0; i < n; ++i) y[i] = ko_expm1f_neg(x[i]); }

// K1-TC — population rollout for WIDE two-hidden-layer actors (BASELINE config 5: hidden = [400,300] / [128,128]) on the
// 5th-generation tensor cores of sm_100a.
//
//   actor  Linear(7,w1) -> act -> Linear(w1,w2) -> LayerNorm(w2) -> act -> Linear(w2,3) -> tanh
//          (the two-hidden-layer generalisation of base/core/genetic_agent.py:78-101; LayerNorm base/core/mod_utils.py:47-50)
//
// One CTA per SM = 256 threads = two GROUPS of 128 threads; a group = 128 envs of one actor (thread = env = row of the layer
// GEMM = TMEM lane).  The two groups run independent task loops, share the plant tables in shared memory, the weight ring,
// the TMEM accumulator and the tensor core, and take turns on them (a shared-memory lock): while one group streams its W1
// slabs through the tensor core the other integrates its plant step.  Per step of a group:
//   layer 1   on CUDA cores, 8 neurons at a time: every thread computes its env's activations and writes them, split
//             into TF32 hi + lo parts, as one K-slab of the A operand in shared memory (UMMA canonical K-major layout,
//             no swizzle);
//   layer 2   D[128 x w2] += A[128 x 8] . W1^T[8 x w2]  by tcgen05.mma (kind::tf32, M = 128, N <= 256 per instruction,
//             fp32 accumulator in TENSOR MEMORY), three products per slab (hi.hi + hi.lo + lo.hi = "3xTF32", fp32-level
//             accuracy); the W1 K-slabs (pre-split and pre-tiled once per launch) are STREAMED from L2 into a 3-stage
//             shared-memory ring by bulk TMA copies (cp.async.bulk + mbarrier complete_tx) — a [400,300] genome is 500 KB
//             and never fits on chip; tcgen05.commit hands each ring slot back when its MMAs have read it;
//   epilogue  tcgen05.ld brings the thread's accumulator row out of TMEM 16 columns at a time: bias, LayerNorm (the whole
//             row lives in one thread: no shuffles), activation, and the 3-row output layer folded into the same pass;
//   plant     CitationEnv.step + the ode5 plant step on CUDA cores (plant_env.cuh), exactly as in K1.
// (Round-2 history: the first version ran two 128-thread CTAs per SM with the plant tables read through L1; at [400,300]
// the two weight rings left ~50 KB of L1 and the plant's table / local-memory traffic thrashed it — 26 % long-scoreboard
// stalls, 1.1e8 env-steps/s.  Sharing one ring, one accumulator and shared-memory tables between two groups fixed that.)
//
// Numerics: tensor-core accumulation order is not reproducible on a CPU, so this path is checked against the torch fp32
// oracle with a tolerance (tests/test_wide_actor_gpu.py), not bit for bit like K1.
#include "plant_env.cuh"

#define TC_THREADS 128
#ifndef TC_STAGES
#define TC_STAGES 3
#endif
#ifndef TC_KSLAB
#define TC_KSLAB 8                 // K values per pipeline stage (a multiple of 8 = the tcgen05.mma K step for TF32)
#endif
#define TC_CH (TC_KSLAB / 4)        // 16-byte K chunks (4 TF32 values) per stage

struct TcArgs {
    RolloutArgs r;                 // env / output part (weights, wt, P4, apc ... unused)
    int w1, w2, n2pad;             // layer widths (w1 already padded to a multiple of TC_KSLAB with zero neurons); w2 padded to 16
    int w1_real;
    int tmem_cols;                 // TMEM columns to allocate (power of two >= 32)
    int small_floats;              // per-actor small parameter block (floats, multiple of 4)
    int stage_floats;              // per-stage W1 slab: hi[2][n2pad][4] + lo[2][n2pad][4]
    const float* small;            // [pop][small_floats]
    const float* tiles;            // [pop][w1/8][stage_floats]
    long long n_tasks; int n_chunks;
    // forward-only mode (serl_actor_forward_wide)
    const float* obs_in; float* act_out; int n_obs;
};

__device__ __forceinline__ float rn_tf32(float x)       // round to nearest TF32 (10-bit mantissa), ties away
{
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// small block layout (floats): W0p[w1][8] (7 weights + bias) | b1[n2pad] | gamma[n2pad] | beta[n2pad] | Wo[3][n2pad] | bo[4]
__host__ __device__ inline int tc_small_floats(int w1, int n2pad) { return w1 * 8 + 6 * n2pad + 4; }

// K0-TC: genome (parameters() order: W0[w1,7] b0[w1] W1[w2,w1] b1 gamma beta Wo[3,w2] bo[3]) -> small block + W1 slabs
__global__ void tc_layout_kernel(const float* __restrict__ w, int pop, int P, int w1, int w1r, int w2, int n2pad, int small_floats,
                                 int stage_floats, float* __restrict__ small, float* __restrict__ tiles)
{
    const int n_stages = w1 / TC_KSLAB;
    const long long per = (long long)small_floats + (long long)n_stages * stage_floats;
    const long long total = (long long)pop * per;
    // genome offsets use the REAL layer-1 width w1r; rows / columns w1r..w1 of the kernel layout are zero neurons
    const int oW1 = 8 * w1r, ob1 = oW1 + w2 * w1r, og = ob1 + w2, obe = og + w2, oWo = obe + w2, obo = oWo + 3 * w2;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(g / per);
        long long i = g - (long long)a * per;
        const float* ga = w + (size_t)a * P;
        if (i < small_floats) {
            float v = 0.f;
            int r = (int)i;
            if (r < w1 * 8) { const int k = r >> 3, c = r & 7; v = k >= w1r ? 0.f : (c < 7 ? ga[k * 7 + c] : ga[7 * w1r + k]); }
            else {
                r -= w1 * 8;
                if (r < n2pad) v = r < w2 ? ga[ob1 + r] : 0.f;
                else if (r < 2 * n2pad) { r -= n2pad; v = r < w2 ? ga[og + r] : 0.f; }
                else if (r < 3 * n2pad) { r -= 2 * n2pad; v = r < w2 ? ga[obe + r] : 0.f; }
                else if (r < 6 * n2pad) { r -= 3 * n2pad; const int j = r / n2pad, n = r % n2pad; v = n < w2 ? ga[oWo + j * w2 + n] : 0.f; }
                else { r -= 6 * n2pad; v = r < 3 ? ga[obo + r] : 0.f; }
            }
            small[(size_t)a * small_floats + i] = v;
        } else {
            i -= small_floats;
            const int s = (int)(i / stage_floats);
            int r = (int)(i - (long long)s * stage_floats);
            const int half = TC_CH * n2pad * 4;
            const int part = r / half;
            r -= part * half;
            const int j = r / (n2pad * 4), n = (r >> 2) % n2pad, kk = r & 3;
            const int k = s * TC_KSLAB + j * 4 + kk;
            const float v = (n < w2 && k < w1r) ? ga[oW1 + n * w1r + k] : 0.f;
            const float hi = rn_tf32(v);
            tiles[(size_t)a * n_stages * stage_floats + (size_t)s * stage_floats + (i - (long long)s * stage_floats)] =
                part == 0 ? hi : rn_tf32(v - hi);
        }
    }
}

// ---- tcgen05 primitives ---------------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes (contiguous 128 B);
// SBO = byte distance between 8-row groups, LBO = byte distance between the two 16-byte K chunks of one MMA K step
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);          // descriptor version 1 (sm_100), layout type 0 = no swizzle
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v)
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct TcCtx {
    const float* small;            // smem: small parameter block of the current actor
    float* a_ring;                 // smem: TC_STAGES x { hi[2][128][4], lo[2][128][4] }
    float* b_ring;                 // smem: TC_STAGES x stage_floats
    uint64_t* full_b;              // [TC_STAGES] TMA landed
    uint64_t* free_s;              // [TC_STAGES] MMAs of the slot retired
    uint64_t* acc_bar;             // accumulator complete
    uint32_t* tmem_slot;           // smem word tcgen05.alloc writes
    uint32_t g;                    // stages issued so far (valid while the group holds the lock)
    uint32_t steps;                // accumulators completed so far (same)
    uint32_t tmem;                 // TMEM base address
    uint32_t* shared_state;        // smem: {lock, g, steps} shared by the two groups
    int grp, gtid;                 // group of this thread (0/1), thread index inside the group
};

__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(TC_THREADS) : "memory"); }
__device__ __forceinline__ bool group_any(int grp, bool pred)
{
    uint32_t r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, %2, %3, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(r) : "r"((uint32_t)pred), "r"(3 + grp), "r"(TC_THREADS) : "memory");
    return r != 0;
}
// the tensor core, its accumulator and the A / W1 rings belong to one group at a time
__device__ __forceinline__ void tc_acquire(TcCtx& c)
{
    if (c.gtid == 0) {
        while (atomicCAS(&c.shared_state[0], 0u, 1u) != 0u) __nanosleep(100);
        __threadfence_block();
    }
    group_sync(c.grp);
    c.g = *reinterpret_cast<volatile uint32_t*>(&c.shared_state[1]);
    c.steps = *reinterpret_cast<volatile uint32_t*>(&c.shared_state[2]);
}
__device__ __forceinline__ void tc_release(TcCtx& c)
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    group_sync(c.grp);                       // every warp of the group has read its TMEM lanes
    if (c.gtid == 0) {
        c.shared_state[1] = c.g;
        c.shared_state[2] = c.steps;
        __threadfence_block();
        atomicExch(&c.shared_state[0], 0u);
    }
}

// one actor forward for the 128 envs of the CTA; every thread passes its own observation and receives its own action
template <int ACT>
__device__ __forceinline__ void tc_actor_forward(TcCtx& c, const TcArgs& ar, const float* tiles_actor, const float* obs, float* action)
{
    const int tid = c.gtid, warp = tid >> 5;          // group-local: warp = TMEM lane quadrant of this thread
    const int w1 = ar.w1, w2 = ar.w2, n2pad = ar.n2pad;
    const int n_stages = w1 / TC_KSLAB;
    const uint32_t stage_bytes = (uint32_t)ar.stage_floats * 4u;
    constexpr int A_STAGE_FLOATS = 2 * TC_CH * TC_THREADS * 4;             // hi + lo, TC_CH chunks x 128 rows x 4 floats
    const float* W0p = c.small;
    tc_acquire(c);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t g0 = c.g;
    // W1 slabs of the first stages of this step (their slots were released by the previous step's MMAs)
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES - 1 && s < n_stages; ++s) {
            const uint32_t g = g0 + s, slot = g % TC_STAGES;
            if (g >= TC_STAGES) mbar_wait(&c.free_s[slot], ((g / TC_STAGES) + 1) & 1);
            mbar_expect_tx(&c.full_b[slot], stage_bytes);
            tma_bulk_g2s(c.b_ring + (size_t)slot * ar.stage_floats, tiles_actor + (size_t)s * ar.stage_floats, stage_bytes, &c.full_b[slot]);
        }
    }
    const uint32_t idesc0 = umma_idesc_tf32(n2pad <= 256 ? n2pad : 256);
    const uint32_t idesc1 = umma_idesc_tf32(n2pad <= 256 ? 16 : n2pad - 256);
    for (int s = 0; s < n_stages; ++s) {
        const uint32_t g = g0 + s, slot = g % TC_STAGES;
        // the slot's previous MMAs must have read A (and B) before it is overwritten
        if (g >= TC_STAGES) mbar_wait(&c.free_s[slot], ((g / TC_STAGES) + 1) & 1);
        // layer 1: this env's 8 activations of the slab, split into TF32 hi / lo, as A rows
        float* a_hi = c.a_ring + (size_t)slot * A_STAGE_FLOATS;
        float* a_lo = a_hi + TC_CH * TC_THREADS * 4;
#pragma unroll
        for (int j = 0; j < TC_CH; ++j) {
            float h[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 wa = *reinterpret_cast<const float4*>(W0p + (size_t)(s * TC_KSLAB + j * 4 + kk) * 8);
                const float4 wb = *reinterpret_cast<const float4*>(W0p + (size_t)(s * TC_KSLAB + j * 4 + kk) * 8 + 4);
                float acc = wb.w;                                    // bias
                acc = __fmaf_rn(wa.x, obs[0], acc); acc = __fmaf_rn(wa.y, obs[1], acc); acc = __fmaf_rn(wa.z, obs[2], acc);
                acc = __fmaf_rn(wa.w, obs[3], acc); acc = __fmaf_rn(wb.x, obs[4], acc); acc = __fmaf_rn(wb.y, obs[5], acc);
                acc = __fmaf_rn(wb.z, obs[6], acc);
                h[kk] = acc;
            }
            const float2 p0 = am_act2<ACT>(make_float2(h[0], h[1])), p1 = am_act2<ACT>(make_float2(h[2], h[3]));
            const float4 hi = make_float4(rn_tf32(p0.x), rn_tf32(p0.y), rn_tf32(p1.x), rn_tf32(p1.y));
            const float4 lo = make_float4(rn_tf32(p0.x - hi.x), rn_tf32(p0.y - hi.y), rn_tf32(p1.x - hi.z), rn_tf32(p1.y - hi.w));
            *reinterpret_cast<float4*>(a_hi + ((size_t)j * TC_THREADS + tid) * 4) = hi;
            *reinterpret_cast<float4*>(a_lo + ((size_t)j * TC_THREADS + tid) * 4) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes of A -> tensor-core (async proxy) reads
        group_sync(c.grp);
        if (tid == 0) {
            mbar_wait(&c.full_b[slot], (g / TC_STAGES) & 1);               // W1 slab landed
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi_addr = smem_u32(a_hi), a_lo_addr = smem_u32(a_lo);
            const uint32_t b_hi_addr = smem_u32(c.b_ring + (size_t)slot * ar.stage_floats);
            const uint32_t b_lo_addr = b_hi_addr + (uint32_t)TC_CH * (uint32_t)n2pad * 16u;
            const uint32_t lbo_a = TC_THREADS * 16, lbo_b = (uint32_t)n2pad * 16;
#pragma unroll
            for (int ks = 0; ks < TC_KSLAB / 8; ++ks) {                         // one K step = two 16-byte chunks
                const uint32_t ao = (uint32_t)(2 * ks) * lbo_a, bo2 = (uint32_t)(2 * ks) * lbo_b;
                const uint64_t dah = umma_desc(a_hi_addr + ao, lbo_a, 128), dal = umma_desc(a_lo_addr + ao, lbo_a, 128);
                const uint32_t acc0 = (s > 0 || ks > 0) ? 1u : 0u;
                // N part 0 (columns 0 .. min(n2pad,256))
                umma_tf32(c.tmem, dah, umma_desc(b_hi_addr + bo2, lbo_b, 128), idesc0, acc0);
                umma_tf32(c.tmem, dah, umma_desc(b_lo_addr + bo2, lbo_b, 128), idesc0, 1u);
                umma_tf32(c.tmem, dal, umma_desc(b_hi_addr + bo2, lbo_b, 128), idesc0, 1u);
                if (n2pad > 256) {                                             // N part 1 (columns 256 .. n2pad)
                    umma_tf32(c.tmem + 256, dah, umma_desc(b_hi_addr + bo2 + 256 * 16, lbo_b, 128), idesc1, acc0);
                    umma_tf32(c.tmem + 256, dah, umma_desc(b_lo_addr + bo2 + 256 * 16, lbo_b, 128), idesc1, 1u);
                    umma_tf32(c.tmem + 256, dal, umma_desc(b_hi_addr + bo2 + 256 * 16, lbo_b, 128), idesc1, 1u);
                }
            }
            umma_commit(&c.free_s[slot]);                                      // slot reusable when these MMAs retire
            if (s == n_stages - 1) umma_commit(c.acc_bar);
            // prefetch the slab TC_STAGES-1 ahead into the slot stage g-1 used
            const int sn = s + TC_STAGES - 1;
            if (sn < n_stages) {
                const uint32_t gn = g + TC_STAGES - 1, sl = gn % TC_STAGES;
                if (gn >= TC_STAGES) mbar_wait(&c.free_s[sl], ((gn / TC_STAGES) + 1) & 1);
                mbar_expect_tx(&c.full_b[sl], stage_bytes);
                tma_bulk_g2s(c.b_ring + (size_t)sl * ar.stage_floats, tiles_actor + (size_t)sn * ar.stage_floats, stage_bytes, &c.full_b[sl]);
            }
        }
    }
    c.g = g0 + n_stages;
    // ---- epilogue: this thread's accumulator row (TMEM lane = thread) ----
    mbar_wait(c.acc_bar, c.steps & 1);
    c.steps += 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const float* b1 = c.small + w1 * 8;
    const float* gamma = b1 + n2pad;
    const float* beta = gamma + n2pad;
    const float* Wo = beta + n2pad;
    const float* bo = Wo + 3 * n2pad;
    const uint32_t trow = c.tmem + ((uint32_t)(warp * 32) << 16);
    const int n_chunks = n2pad / 16;
    float v[16];
#ifdef TC_EPI2
    // mean and variance in ONE pass over the accumulator row, shifted by its first element (no cancellation problem:
    // |mean - x0| is of the order of the row's own spread)
    float x0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int ch = 0; ch < n_chunks; ++ch) {
        tmem_ld16(trow + ch * 16, v);
        if (ch == 0) x0 = __fadd_rn(v[0], b1[0]);
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (ch * 16 + i < w2) { const float d = __fadd_rn(__fadd_rn(v[i], b1[ch * 16 + i]), -x0); s1 = __fadd_rn(s1, d); s2 = __fmaf_rn(d, d, s2); }
    }
    const float mean = __fadd_rn(x0, __fdiv_rn(s1, (float)w2));
    const float ss = fmaxf(__fmaf_rn(-s1, __fdiv_rn(s1, (float)w2), s2), 0.f);
#else
    float sum = 0.f;
    for (int ch = 0; ch < n_chunks; ++ch) {
        tmem_ld16(trow + ch * 16, v);
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (ch * 16 + i < w2) sum = __fadd_rn(sum, __fadd_rn(v[i], b1[ch * 16 + i]));
    }
    const float mean = __fdiv_rn(sum, (float)w2);
    float ss = 0.f;
    for (int ch = 0; ch < n_chunks; ++ch) {
        tmem_ld16(trow + ch * 16, v);
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (ch * 16 + i < w2) { const float d = __fadd_rn(__fadd_rn(v[i], b1[ch * 16 + i]), -mean); ss = __fmaf_rn(d, d, ss); }
    }
#endif
    const float inv = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__fdiv_rn(ss, (float)(w2 - 1))), 1e-6f));
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
    for (int ch = 0; ch < n_chunks; ++ch) {
        tmem_ld16(trow + ch * 16, v);
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            const int j = ch * 16 + i;
            const float d0 = __fadd_rn(__fadd_rn(v[i], b1[j]), -mean), d1 = __fadd_rn(__fadd_rn(v[i + 1], b1[j + 1]), -mean);
            const float2 y = am_act2<ACT>(make_float2(__fmaf_rn(__fmul_rn(gamma[j], d0), inv, beta[j]),
                                                      __fmaf_rn(__fmul_rn(gamma[j + 1], d1), inv, beta[j + 1])));
            // padded columns carry zero weights in Wo, so they add nothing
            o0 = __fmaf_rn(Wo[j], y.x, o0); o0 = __fmaf_rn(Wo[j + 1], y.y, o0);
            o1 = __fmaf_rn(Wo[n2pad + j], y.x, o1); o1 = __fmaf_rn(Wo[n2pad + j + 1], y.y, o1);
            o2 = __fmaf_rn(Wo[2 * n2pad + j], y.x, o2); o2 = __fmaf_rn(Wo[2 * n2pad + j + 1], y.y, o2);
        }
    }
    action[0] = am_tanh1(__fadd_rn(o0, bo[0]));
    action[1] = am_tanh1(__fadd_rn(o1, bo[1]));
    action[2] = am_tanh1(__fadd_rn(o2, bo[2]));
    // every warp has read its TMEM lanes: hand the tensor core, the accumulator and the rings to the other group
    tc_release(c);
}

constexpr int TC_TABN = PT_TOTAL + SERL_PLANT_COUNT * PLANT_NPV;      // plant tables + per-variant parameter rows
constexpr int TC_TABN2 = (TC_TABN + 15) & ~15;                          // keeps the float regions 128-byte aligned

__device__ __forceinline__ void tc_setup(TcCtx& c, const TcArgs& ar, unsigned char* smem_raw, uint64_t* bars, uint32_t* tmem_slot,
                                         uint32_t* shared_state)
{
    constexpr int A_STAGE_FLOATS = 2 * TC_CH * TC_THREADS * 4;
    real* tab_s = reinterpret_cast<real*>(smem_raw);
    for (int i = threadIdx.x; i < PT_TOTAL; i += blockDim.x) tab_s[i] = plant_tables_blob[i];
    for (int i = threadIdx.x; i < SERL_PLANT_COUNT * PLANT_NPV; i += blockDim.x) tab_s[PT_TOTAL + i] = (&plant_pv[0][0])[i];
    float* f = reinterpret_cast<float*>(tab_s + TC_TABN2);
    const int small_pad = (ar.small_floats + 31) & ~31;
    c.grp = threadIdx.x >> 7;
    c.gtid = threadIdx.x & (TC_THREADS - 1);
    c.small = f + (size_t)c.grp * small_pad;
    c.a_ring = f + 2 * (size_t)small_pad;
    c.b_ring = c.a_ring + TC_STAGES * A_STAGE_FLOATS;
    c.full_b = bars; c.free_s = bars + TC_STAGES; c.acc_bar = bars + 2 * TC_STAGES;
    c.tmem_slot = tmem_slot;
    c.shared_state = shared_state;
    c.g = 0; c.steps = 0; c.tmem = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * TC_STAGES + 1; ++i) mbar_init(&bars[i], 1);
        shared_state[0] = shared_state[1] = shared_state[2] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ar.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    c.tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
}

__device__ __forceinline__ void tc_teardown(TcCtx& c, const TcArgs& ar)
{
    __syncthreads();
    if ((threadIdx.x >> 5) == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "r"(ar.tmem_cols) : "memory");
}

__device__ __forceinline__ void tc_load_small(TcCtx& c, const TcArgs& ar, int actor)
{
    group_sync(c.grp);                // the group's previous actor's parameters are no longer read
    float* dst = const_cast<float*>(c.small);
    const float4* src = reinterpret_cast<const float4*>(ar.small + (size_t)actor * ar.small_floats);
    for (int i = c.gtid; i < ar.small_floats / 4; i += TC_THREADS) reinterpret_cast<float4*>(dst)[i] = src[i];
    group_sync(c.grp);
}

template <int ACT>
__global__ void __launch_bounds__(2 * TC_THREADS, 1)
rollout_kernel_tc(TcArgs ar)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[2 * TC_STAGES + 1];
    __shared__ uint32_t tmem_slot;
    __shared__ uint32_t shared_state[4];
    TcCtx c;
    plant_tab_check(smem_raw);
    tc_setup(c, ar, smem_raw, bars, &tmem_slot, shared_state);
    const RolloutArgs& r = ar.r;
    const real* tab = reinterpret_cast<const real*>(smem_raw);
    const real* pv_base = tab + PT_TOTAL;
    int gust_mine = 0;
    for (int i = threadIdx.x; i < r.n_envs; i += blockDim.x) gust_mine |= r.env_mode[i] & SERL_MODE_GUST;
    const bool any_gust = __syncthreads_or(gust_mine) != 0;      // does any env of the launch fly the gust build?
    const int tid = c.gtid;
    const int n_stages = ar.w1 / TC_KSLAB;
    // the two groups of a CTA run independent task loops
    for (long long task = (long long)blockIdx.x * 2 + c.grp; task < ar.n_tasks; task += 2LL * gridDim.x) {
        const int actor = (int)(task / ar.n_chunks), chunk = (int)(task - (long long)actor * ar.n_chunks);
        tc_load_small(c, ar, actor);
        const float* tiles_actor = ar.tiles + (size_t)actor * n_stages * ar.stage_floats;
        const int eslot = chunk * TC_THREADS + tid;
        const bool valid = eslot < r.n_envs;
        const int env = valid ? (r.env_order ? r.env_order[eslot] : eslot) : 0;
        Env e;
        e.tab = tab;
        float obs[7], a[3];
        if (valid) {
            env_bind(e, r, env, pv_base, (size_t)actor * r.n_envs + env);
            env_reset<true>(e, r, env, obs, (size_t)actor * r.n_envs + env);
        } else {
            e.done = true; e.k = 0; e.ret = 0.0; e.t = 0.0; e.fault = 0; e.gust = 0; e.pv = pv_base; e.pv_post = nullptr; e.theta_trim = 0.0;
            e.ref_lv = r.ref_levels; e.ref_st = r.ref_starts;
#pragma unroll
            for (int i = 0; i < NX; ++i) e.X[i] = 0.0;
#pragma unroll
            for (int i = 0; i < 7; ++i) obs[i] = 0.f;
        }
        const size_t traj = (size_t)actor * r.n_envs + env;
        const bool replay = valid && r.replay != nullptr && env == r.replay_env;
        while (group_any(c.grp, !e.done)) {
            tc_actor_forward<ACT>(c, ar, tiles_actor, obs, a);
            if (!e.done) {
                if (any_gust) env_step<true, false, true>(e, r, traj, actor, replay, a, obs);
                else env_step<true>(e, r, traj, actor, replay, a, obs);
            }
        }
        if (valid) {
            r.returns[traj] = e.ret;
            r.steps[traj] = e.k;
            if (r.status && !isfinite(e.ret + e.X[3] + e.X[7] + e.X[9])) atomicOr(r.status, SERL_STATUS_NONFINITE);
        }
    }
    tc_teardown(c, ar);
}

// Actor.forward for a batch through the same tensor-core device code (parity tests of the GEMM path)
template <int ACT>
__global__ void __launch_bounds__(2 * TC_THREADS, 1)
actor_forward_tc_kernel(TcArgs ar)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[2 * TC_STAGES + 1];
    __shared__ uint32_t tmem_slot;
    __shared__ uint32_t shared_state[4];
    TcCtx c;
    tc_setup(c, ar, smem_raw, bars, &tmem_slot, shared_state);
    tc_load_small(c, ar, 0);
    const int tid = c.gtid;
    for (int base = (blockIdx.x * 2 + c.grp) * TC_THREADS; base < ar.n_obs; base += 2 * gridDim.x * TC_THREADS) {
        const int i = base + tid;
        float obs[7], a[3];
#pragma unroll
        for (int k = 0; k < 7; ++k) obs[k] = i < ar.n_obs ? ar.obs_in[(size_t)i * 7 + k] : 0.f;
        tc_actor_forward<ACT>(c, ar, ar.tiles, obs, a);
        if (i < ar.n_obs) { ar.act_out[(size_t)i * 3] = a[0]; ar.act_out[(size_t)i * 3 + 1] = a[1]; ar.act_out[(size_t)i * 3 + 2] = a[2]; }
    }
    tc_teardown(c, ar);
}

// ---- host side ------------------------------------------------------------------------------------------------
#include <map>
#include <mutex>
static std::mutex g_tc_mu;
static std::map<std::pair<int, cudaStream_t>, std::pair<void*, size_t>> g_tc_scratch;
static cudaError_t tc_scratch_get(cudaStream_t s, size_t bytes, void** out)
{
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_tc_mu);
    auto& b = g_tc_scratch[std::make_pair(dev, s)];
    if (b.second < bytes) {
        if (b.first) { cudaStreamSynchronize(s); cudaFree(b.first); b.first = nullptr; b.second = 0; }
        cudaError_t e = cudaMalloc(&b.first, bytes + bytes / 8);
        if (e != cudaSuccess) return e;
        b.second = bytes + bytes / 8;
    }
    *out = b.first;
    return cudaSuccess;
}

extern "C" int64_t serl_actor_num_params_wide(const int32_t* widths, int32_t n_widths)
{
    if (!widths || n_widths < 1) return -1;
    int64_t P = 7 * (int64_t)widths[0] + widths[0];
    for (int i = 1; i < n_widths; ++i) P += (int64_t)widths[i] * widths[i - 1] + 3 * (int64_t)widths[i];
    return P + 3 * (int64_t)widths[n_widths - 1] + 3;
}

static int tc_prepare(TcArgs& ar, const float* d_weights, int pop, const int32_t* widths, int n_widths, cudaStream_t s, size_t* smem_out)
{
    if (n_widths != 2) return serl_fail(SERL_ERR_UNSUPPORTED, "wide actors: the tensor-core path implements two hidden layers [w1, w2]");
    const int w1 = widths[0], w2 = widths[1];
    if (w1 < 8 || w1 % 8 != 0 || w1 > 1024 || w2 < 8 || w2 > 320)
        return serl_fail(SERL_ERR_UNSUPPORTED, "wide actors: need w1 % 8 == 0, 8 <= w1 <= 1024, 8 <= w2 <= 320");
    ar.w1_real = w1;
    ar.w1 = (w1 + TC_KSLAB - 1) / TC_KSLAB * TC_KSLAB;
    ar.w2 = w2; ar.n2pad = (w2 + 15) & ~15;
    int cols = 32;
    while (cols < ar.n2pad) cols <<= 1;
    ar.tmem_cols = cols;
    ar.small_floats = (tc_small_floats(ar.w1, ar.n2pad) + 3) & ~3;
    ar.stage_floats = 2 * TC_CH * ar.n2pad * 4;
    const int P = (int)serl_actor_num_params_wide(widths, n_widths);
    const int n_stages = ar.w1 / TC_KSLAB;
    const size_t small_bytes = (size_t)pop * ar.small_floats * 4;
    const size_t tile_bytes = (size_t)pop * n_stages * ar.stage_floats * 4;
    void* scratch = nullptr;
    cudaError_t e = tc_scratch_get(s, ((small_bytes + 255) & ~(size_t)255) + tile_bytes + 256, &scratch);
    if (e != cudaSuccess) return serl_fail_cuda(e, "wide actors: scratch allocation");
    float* small = (float*)scratch;
    float* tiles = (float*)((unsigned char*)scratch + ((small_bytes + 255) & ~(size_t)255));
    ar.small = small; ar.tiles = tiles;
    const long long total = (long long)pop * ((long long)ar.small_floats + (long long)n_stages * ar.stage_floats);
    const int grid = (int)((total + 255) / 256 < 8192 ? (total + 255) / 256 : 8192);
    tc_layout_kernel<<<grid, 256, 0, s>>>(d_weights, pop, P, ar.w1, w1, w2, ar.n2pad, ar.small_floats, ar.stage_floats, small, tiles);
    serl_count_launch();
    constexpr int A_STAGE_FLOATS = 2 * TC_CH * TC_THREADS * 4;
    *smem_out = (size_t)TC_TABN2 * sizeof(real) +
                (size_t)(2 * ((ar.small_floats + 31) & ~31) + TC_STAGES * A_STAGE_FLOATS + TC_STAGES * ar.stage_floats) * 4;
    if (*smem_out > 227 * 1024 - 256) return serl_fail(SERL_ERR_UNSUPPORTED, "wide actors: tables + parameters + rings exceed the shared memory of an SM");
    return SERL_OK;
}

int rollout_tc_impl(const serl_rollout_desc& d, const int32_t* widths, int n_widths, void* stream)
{
    if (!d.d_weights || !d.d_ref_levels || !d.d_ref_starts || !d.d_env_mode || !d.d_returns || !d.d_steps)
        return serl_fail(SERL_ERR_ARG, "serl_rollout: null pointer argument");
    if (d.pop <= 0 || d.n_envs <= 0 || d.horizon <= 0) return serl_fail(SERL_ERR_ARG, "serl_rollout: pop, n_envs, horizon must be > 0");
    if (d.shape.activation < 0 || d.shape.activation > 2) return serl_fail(SERL_ERR_ARG, "serl_rollout: unsupported activation");
    if (d.d_trace) return serl_fail(SERL_ERR_UNSUPPORTED, "wide actors: per-step traces are not produced by the tensor-core kernel");
    cudaStream_t s = (cudaStream_t)stream;
    TcArgs ar;
    memset(&ar, 0, sizeof(ar));
    RolloutArgs& r = ar.r;
    r.ref_levels = d.d_ref_levels; r.ref_starts = d.d_ref_starts; r.env_mode = d.d_env_mode; r.n_envs = d.n_envs; r.horizon = d.horizon;
    r.action_noise = d.d_action_noise; r.returns = d.d_returns; r.steps = d.d_steps; r.actions = d.d_actions; r.pop = d.pop;
    r.t_max = d.t_max > 0.0 ? d.t_max : 20.0;
    r.smooth_w = d.t_max > 0.0 ? d.smooth_width : 3.0;
    r.env_order = d.d_env_order; r.replay = d.d_replay; r.replay_env = d.replay_env; r.status = d.d_status; r.sensor_noise = d.d_sensor_noise;
    size_t smem = 0;
    int rc = tc_prepare(ar, d.d_weights, d.pop, widths, n_widths, s, &smem);
    if (rc != SERL_OK) return rc;
    ar.n_chunks = (d.n_envs + TC_THREADS - 1) / TC_THREADS;
    ar.n_tasks = (long long)d.pop * ar.n_chunks;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (d.sm_limit > 0 && d.sm_limit < sms) sms = d.sm_limit;
    const long long grid = (ar.n_tasks + 1) / 2 < sms ? (ar.n_tasks + 1) / 2 : sms;
    cudaError_t e;
#define TC_LAUNCH(A) do { e = cudaFuncSetAttribute(rollout_kernel_tc<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e == cudaSuccess) { rollout_kernel_tc<A><<<(unsigned)grid, 2 * TC_THREADS, smem, s>>>(ar); e = cudaGetLastError(); } } while (0)
    if (d.shape.activation == SERL_ACT_TANH) TC_LAUNCH(SERL_ACT_TANH);
    else if (d.shape.activation == SERL_ACT_ELU) TC_LAUNCH(SERL_ACT_ELU);
    else TC_LAUNCH(SERL_ACT_LEAKY_RELU);
#undef TC_LAUNCH
    serl_count_launch();
    if (e != cudaSuccess) return serl_fail_cuda(e, "rollout_kernel_tc launch");
    return SERL_OK;
}

extern "C" int serl_actor_forward_wide(const float* d_genome, const int32_t* widths, int32_t n_widths, int32_t activation,
                                       const float* d_obs, int32_t n, float* d_actions, void* stream)
{
    if (!d_genome || !widths || !d_obs || !d_actions || n <= 0 || activation < 0 || activation > 2)
        return serl_fail(SERL_ERR_ARG, "serl_actor_forward_wide: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    TcArgs ar;
    memset(&ar, 0, sizeof(ar));
    size_t smem = 0;
    int rc = tc_prepare(ar, d_genome, 1, widths, n_widths, s, &smem);
    if (rc != SERL_OK) return rc;
    ar.obs_in = d_obs; ar.act_out = d_actions; ar.n_obs = n;
    const int blocks = (n + 2 * TC_THREADS - 1) / (2 * TC_THREADS);
    const int grid = blocks < 148 ? blocks : 148;
    cudaError_t e;
#define TC_LAUNCH(A) do { e = cudaFuncSetAttribute(actor_forward_tc_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e == cudaSuccess) { actor_forward_tc_kernel<A><<<grid, 2 * TC_THREADS, smem, s>>>(ar); e = cudaGetLastError(); } } while (0)
    if (activation == SERL_ACT_TANH) TC_LAUNCH(SERL_ACT_TANH);
    else if (activation == SERL_ACT_ELU) TC_LAUNCH(SERL_ACT_ELU);
    else TC_LAUNCH(SERL_ACT_LEAKY_RELU);
#undef TC_LAUNCH
    serl_count_launch();
    return e == cudaSuccess ? SERL_OK : serl_fail_cuda(e, "actor_forward_tc_kernel");
}

// K1 — fused population rollout for sm_100a (+ K6 smoothness metric, batched plant step).
//
// One lane = one (actor, env) trajectory; one warp = 32 envs of one actor; one CTA = 1-2 actors x 128 envs with their
// fp32 genomes (and the plant tables) staged once in shared memory.  Per step a warp runs, entirely on chip:
//   Actor.select_action  (base/core/genetic_agent.py:104-109; LayerNorm base/core/mod_utils.py:47-50)        fp32
//   CitationEnv.step     (envs/phlabenv.py:430-482: action scaling :62-73, fault shims envs/{be,jr,sa,se}/citation.py,
//                         reward :362-367, termination + penalty :391-399)                                fp64
//   native plant step    (envs/<variant>/_citation*.so step @0x6030: 6-stage Dormand-Prince ode5, h = 0.01, RHS
//                         generated from the binaries by tools/lift -> csrc/gen/plant_rhs_common.h)          fp64
// and accumulates the episodic return (base/core/agent.py:129).  HBM is touched only at episode start
// (genome, reference-signal parameters) and end (return, step count) unless a trace / action history is requested.
//
// Two actor implementations:
//   rollout_kernel_persist<H, TABS, GUST>  persistent grid, one CTA per SM.  The MLP is a register-tiled GEMM inside a warp
//                           (lane = 1/4 of the output neurons x 4 envs, packed FFMA2, activations exchanged with warp
//                           shuffles, weights broadcast from shared memory; h in {32,64,72,96,128}).  The CTA's warps take
//                           their steps in lockstep (one barrier per step: shared instruction fetch), genomes arrive by bulk
//                           TMA copies, the ode5 stage derivatives live in tensor memory.
//   rollout_kernel_simple   every thread runs the whole MLP for its env (any h that fits); cross-check / fallback shape.
#include "plant_env.cuh"

// ---- simple actor: every thread evaluates the whole MLP for its own observation -------------------------
__device__ void actor_forward_simple(const float* __restrict__ w, const serl_actor_shape sh, float* bufA, float* bufB,
                                     int tid, int nthr, const float* obs, float* action)
{
    // same arithmetic specification as actor_forward_warp (dot products: sequential fma from 0; LayerNorm / output-layer
    // sums: four contiguous quarter blocks, combined (q0+q1)+(q2+q3)), so both kernels produce identical bits
    const int S = sh.state_dim, A = sh.action_dim, H = sh.hidden, L = sh.num_layers;
    const int TM = (H + 3) / 4;
    const float* p = w;
    for (int j = 0; j < H; ++j) {
        float acc = 0.f;
        for (int i = 0; i < S; ++i) acc = __fmaf_rn(p[j * S + i], obs[i], acc);
        acc = __fadd_rn(acc, p[H * S + j]);
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
        float q[4];
        for (int g = 0; g < 4; ++g) {
            float sum = 0.f;
            for (int j = g * TM; j < min((g + 1) * TM, H); ++j) {
                float acc = 0.f;
                for (int i = 0; i < H; ++i) acc = __fmaf_rn(W[j * H + i], in[i * nthr + tid], acc);
                acc = __fadd_rn(acc, b[j]);
                out[j * nthr + tid] = acc;
                sum = __fadd_rn(sum, acc);
            }
            q[g] = sum;
        }
        const float mean = __fdiv_rn(__fadd_rn(__fadd_rn(q[0], q[1]), __fadd_rn(q[2], q[3])), (float)H);
        for (int g = 0; g < 4; ++g) {
            float ss = 0.f;
            for (int j = g * TM; j < min((g + 1) * TM, H); ++j) {
                const float d = __fadd_rn(out[j * nthr + tid], -mean);
                out[j * nthr + tid] = d;
                ss = __fmaf_rn(d, d, ss);
            }
            q[g] = ss;
        }
        const float var = __fdiv_rn(__fadd_rn(__fadd_rn(q[0], q[1]), __fadd_rn(q[2], q[3])), (float)(H - 1));
        const float inv = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(var), 1e-6f));
        for (int j = 0; j < H; ++j)
            out[j * nthr + tid] = act_fn(sh.activation, __fmaf_rn(__fmul_rn(gamma[j], out[j * nthr + tid]), inv, beta[j]));
        p += H * H + 3 * H;
        float* t = in; in = out; out = t;
    }
    for (int j = 0; j < A; ++j) {
        float q[4];
        for (int g = 0; g < 4; ++g) {
            float acc = 0.f;
            for (int i = g * TM; i < min((g + 1) * TM, H); ++i) acc = __fmaf_rn(p[j * H + i], in[i * nthr + tid], acc);
            q[g] = acc;
        }
        action[j] = am_tanh1(__fadd_rn(__fadd_rn(__fadd_rn(q[0], q[1]), __fadd_rn(q[2], q[3])), p[A * H + j]));
    }
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
    v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return v;
}

template <int H, int ACT>
__device__ __noinline__ void actor_forward_warp(const float* __restrict__ w, int L, int lane, const float* obs, float* action)
{
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
        for (int c = 0; c < 4; ++c) in[m2][c] = am_act2<ACT>(am_fma2(acc[m2][c], am_splat(1.0f), b));
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
                acc[m2][c] = am_fma2(acc[m2][c], am_splat(1.0f), b);
                s[c] = __fadd_rn(s[c], acc[m2][c].x);
                s[c] = __fadd_rn(s[c], acc[m2][c].y);
            }
        }
        float mean[4], den[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) mean[c] = __fdiv_rn(group_sum(s[c]), (float)H);
        float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int m2 = 0; m2 < TM2; ++m2)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                acc[m2][c] = am_fma2(acc[m2][c], am_splat(1.0f), am_splat(-mean[c]));
                q[c] = __fmaf_rn(acc[m2][c].x, acc[m2][c].x, q[c]);
                q[c] = __fmaf_rn(acc[m2][c].y, acc[m2][c].y, q[c]);
            }
        // gamma * (x - mean) / (std + eps) + beta as fma(gamma * d, 1 / (std + eps), beta): one reciprocal per env instead
        // of H/4 divisions per lane (<= 1 ulp from the reference's expression, the order of its summation-order freedom)
#pragma unroll
        for (int c = 0; c < 4; ++c)
            den[c] = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__fdiv_rn(group_sum(q[c]), (float)(H - 1))), 1e-6f));
#pragma unroll
        for (int m2 = 0; m2 < TM2; ++m2) {
            const float2 g = *reinterpret_cast<const float2*>(gamma + og * TM + 2 * m2);
            const float2 be = *reinterpret_cast<const float2*>(beta + og * TM + 2 * m2);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                in[m2][c] = am_act2<ACT>(am_fma2(am_fma2(g, acc[m2][c], am_splat(-0.0f)), am_splat(den[c]), be));
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
            for (int c = 0; c < 4; ++c) { p[c] = __fmaf_rn(wv.x, in[m2][c].x, p[c]); p[c] = __fmaf_rn(wv.y, in[m2][c].y, p[c]); }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) p[c] = group_sum(p[c]);
        const float mine = og == 0 ? p[0] : (og == 1 ? p[1] : (og == 2 ? p[2] : p[3]));
        action[j] = am_tanh1(__fadd_rn(mine, bo[j]));
    }
}


// ---- genome layout in shared memory -------------------------------------------------------------------------
// parameters() order in HBM (row-major [out][in]) -> the kernel's layout (matrices transposed to [in][out]):
//   Wt0[S][H] b0[H] | L x { Wt[H][H] b[H] gamma[H] beta[H] } | Wo[A][H] bo[A]
__device__ __forceinline__ int genome_layout_index(int i, int S, int H, int L)
{
    int r = i;
    if (r < S * H) { const int j = r / S, k = r % S; return k * H + j; }
    r -= S * H;
    if (r < H) return S * H + r;
    r -= H;
    const int per = H * H + 3 * H;
    if (r < L * per) {
        const int l = r / per, q = r % per;
        const int base = S * H + H + l * per;
        if (q < H * H) { const int j = q / H, k = q % H; return base + k * H + j; }
        return base + q;
    }
    return i;      // Wo [A][H] then bo[A]: unchanged
}

// K0: all genomes of a launch into the shared-memory layout, rows padded to 16 bytes, so that the rollout kernel can
// bring a genome into shared memory with ONE bulk TMA copy (cp.async.bulk) instead of a scattered staging loop.
__global__ void genome_layout_kernel(const float* __restrict__ w, float* __restrict__ wt, int pop, int P, int P4, int S, int H, int L)
{
    const long long n = (long long)pop * P;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(g / P), i = (int)(g % P);
        wt[(size_t)a * P4 + genome_layout_index(i, S, H, L)] = w[g];
    }
}

// ---- K1: persistent rollout ------------------------------------------------------------------------------------
// grid = one CTA per SM (or fewer when there is less work); a CTA has `apc` genome slots of `wps` warps each.  A task is
// (actor, chunk of wps*32 envs) x horizon steps.  Tasks are dealt to the slots as EQUAL SHARES OF STEPS: slot s owns
// the range [s W/NS, (s+1) W/NS) of the linearised (task, step) space, W = n_tasks * horizon — i.e. possibly the tail of
// one task, some whole tasks, and the head of another.  A slot flies the head segment FIRST and publishes the
// trajectories' state in HBM (Handoff), then its whole tasks, and LAST the tail segment, whose first part the previous
// slot published long before: no slot ever waits in practice, and all SMs finish together (512 actors x 128 envs on
// 148 SMs x 2 slots = 1.73 tasks per slot: two full rounds without the split).  When there are fewer tasks than
// slots every task is flown whole by one slot.  The genome of a slot is swapped by ONE elected thread with a bulk TMA copy;
// the slot's warps meet at its named barrier for that and for the slot-uniform decisions of the lockstep loop (see there).
// TABS: plant tables staged in shared memory (true) or read from global memory through L1 (false: h = 128, whose
// 207 KB genome leaves no room for them).
// GUST: the launch contains envs of the gust build (serl_rollout_desc.flags & SERL_ROLLOUT_GUST); the training instantiation
// carries no trace of the feature (a gust env in it raises SERL_STATUS_GUST_FLAG)
template <int H, bool TABS, bool GUST>
__global__ void __launch_bounds__(MAX_CTA_THREADS, 1)
rollout_kernel_persist(RolloutArgs ar)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];     // same alignment as plant_smem_tab (plant_env.cuh)
    __shared__ uint64_t gbar[4];                       // one mbarrier per genome slot
    __shared__ uint32_t tmem_slot;                     // base address of the CTA's tensor memory (stage derivatives)
    if (TABS) plant_tab_check(smem_raw);
    real* tab_s = reinterpret_cast<real*>(smem_raw);
    constexpr int TABN = PT_TOTAL + SERL_PLANT_COUNT * PLANT_NPV;      // tables + per-variant parameter rows
    constexpr int TABN2 = (TABN + 1) & ~1;
    float* wbase = reinterpret_cast<float*>(tab_s + (TABS ? TABN2 : 0));
    const real* tab = TABS ? tab_s : plant_tables_blob;
    const real* pv_base = TABS ? tab_s + PT_TOTAL : &plant_pv[0][0];
    const int L = ar.sh.num_layers;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wps = ar.wps;
    const int slot_l = warp / wps, wslot = warp - slot_l * wps;
    const long long slot = (long long)blockIdx.x * ar.apc + slot_l;
    if (TABS) {
        for (int i = tid; i < PT_TOTAL; i += blockDim.x) tab_s[i] = plant_tables_blob[i];
        for (int i = tid; i < SERL_PLANT_COUNT * PLANT_NPV; i += blockDim.x) tab_s[PT_TOTAL + i] = (&plant_pv[0][0])[i];
    }
    if (tid == 0) {
        for (int i = 0; i < ar.apc; ++i) mbar_init(&gbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // Tensor memory for the ode5 stage derivatives (plant_env.cuh): the whole 512 columns, allocated by warp 0.  Traced
    // launches (single episodes; the navigation integrator wants all six stages) and the float build keep local memory.
    const bool use_tmem = sizeof(real) == 8 && ar.trace == nullptr;
    if (use_tmem && warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = use_tmem ? *reinterpret_cast<volatile uint32_t*>(&tmem_slot) : 0u;
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * PLANT_TMEM_COLS_PER_WARP);
    float* w = wbase + (size_t)slot_l * ar.P4;
    const int slot_threads = wps * 32;
    const int actfn = ar.sh.activation;
    const int horizon = ar.horizon;
    int cur_actor = -1;
    uint32_t gphase = 0;

    // this slot's share of the (task, step) space
    const long long NT = ar.n_tasks, NS = ar.n_slots;
    long long t_first = 0, t_last = 0;
    int k0 = 0, k1 = 0;
    int stage = 0;                                     // 3 = no (more) segments
    if (NT <= NS) {
        if (slot >= NT) stage = 3;                     // idle slot: still takes part in the CTA barriers below
        t_first = slot; t_last = slot + 1;
    } else {
        const long long W = NT * horizon;
        const long long lo = slot * W / NS, hi = (slot + 1) * W / NS;
        t_first = lo / horizon; t_last = hi / horizon;
        k0 = (int)(lo - t_first * horizon); k1 = (int)(hi - t_last * horizon);
    }
    // segments in the order: head of the last task (published) -> whole tasks -> tail of the first task (continued)
    long long t_cur = t_first + (k0 > 0 ? 1 : 0);
    // LOCKSTEP: all warps of the CTA take their steps together (one CTA barrier per step).  The step body is ~180 KB of
    // straight-line code (actor 100 KB, right-hand side 41 KB x 6 calls, integrator, environment); eight warps drifting
    // through it independently each stream it through the instruction caches on their own, and instruction fetch was the
    // top stall (no_instruction 27 % of the warp samples).  In lockstep a fetched line serves every warp of the SM.
    Env e;
    e.tab = tab;
    e.done = true; e.k = 0;
    float obs[7], a[3];
    bool in_seg = false, pending = false, to_h = false, valid = false, replay = false;
    int ke = 0, actor = 0;
    size_t traj = 0;
    for (;;) {
        // is this slot's segment still flying?  (slot-uniform: OR over the slot's warps at its named barrier)
        bool slot_alive = false;
        if (in_seg) {
            int any;
            asm volatile("{ .reg .pred p, q; setp.ne.s32 p, %1, 0; barrier.cta.red.or.pred q, %2, %3, p; selp.s32 %0, 1, 0, q; }"
                         : "=r"(any) : "r"((int)(!e.done && e.k < ke)), "r"(1 + slot_l), "r"(slot_threads) : "memory");
            slot_alive = any != 0;
        }
        if (!slot_alive) {
            if (in_seg) {                              // close the finished segment
                if (to_h) {
                    const long long hx = slot * slot_threads + wslot * 32 + lane;
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < NX; ++i) __stcg(ar.ho.X + (size_t)i * ar.ho.n + hx, e.X[i]);
                        __stcg(ar.ho.t + hx, e.t); __stcg(ar.ho.ret + hx, e.ret);
#pragma unroll
                        for (int i = 0; i < 7; ++i) __stcg(ar.ho.obs + (size_t)i * ar.ho.n + hx, obs[i]);
                        __stcg(ar.ho.k + hx, e.k | ((e.done ? 1 : 0) << 30));
                    }
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) atomicExch(ar.ho.flag + slot * wps + wslot, 1);
                } else if (valid) {
                    ar.returns[traj] = e.ret;
                    ar.steps[traj] = e.k;
                    if (ar.status && !isfinite(e.ret + e.X[3] + e.X[7] + e.X[9])) atomicOr(ar.status, SERL_STATUS_NONFINITE);   // NaN actions poison the state at once
                }
                in_seg = false;
            }
            // open the next one, if any.  The tail segment continues trajectories the PREVIOUS slot publishes at the end of
            // its head segment: normally long done, but that slot may sit in this very CTA and advance only with this
            // one's steps (lockstep), so the slot never blocks on the record — it stays `pending` and asks again at the next
            // step.  All decisions are slot-uniform (AND over the slot's warps at its named barrier).
            long long task = 0;
            bool have = false;
            if (!pending) {
                to_h = false;
                while (!have && stage < 3) {
                    if (stage == 0) {
                        stage = 1;
                        if (k1 != 0) { task = t_last; ke = k1; to_h = true; have = true; }
                    } else if (stage == 1) {
                        if (t_cur >= t_last) stage = 2;
                        else { task = t_cur++; ke = horizon; have = true; }
                    } else {
                        stage = 3;
                        if (k0 != 0) { ke = horizon; pending = true; }
                    }
                }
            }
            const bool from_h = pending;
            if (pending) {
                int ok = 0;
                if (lane == 0) ok = *(const volatile int*)(ar.ho.flag + (slot - 1) * wps + wslot) != 0;
                ok = __shfl_sync(0xffffffffu, ok, 0);
                int all;
                asm volatile("{ .reg .pred p, q; setp.ne.s32 p, %1, 0; barrier.cta.red.and.pred q, %2, %3, p; selp.s32 %0, 1, 0, q; }"
                             : "=r"(all) : "r"(ok), "r"(1 + slot_l), "r"(slot_threads) : "memory");
                if (all) { pending = false; task = t_first; have = true; __threadfence(); }
                else __nanosleep(500);
            }
            if (have) {
                actor = (int)(task / ar.n_chunks);
                const int chunk = (int)(task - (long long)actor * ar.n_chunks);
                if (actor != cur_actor) {
                    // swap the genome of this slot: everyone has left the previous segment -> one thread launches the bulk copy
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + slot_l), "r"(slot_threads) : "memory");
                    if (wslot == 0 && lane == 0) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of w before the async write
                        const uint32_t bytes = (uint32_t)ar.P4 * 4u;
                        mbar_expect_tx(&gbar[slot_l], bytes);
                        tma_bulk_g2s(w, ar.wt + (size_t)actor * ar.P4, bytes, &gbar[slot_l]);
                    }
                    mbar_wait(&gbar[slot_l], gphase);
                    gphase ^= 1;
                    cur_actor = actor;
                }
                const int eslot = chunk * slot_threads + wslot * 32 + lane;
                valid = eslot < ar.n_envs;
                const int env = valid ? (ar.env_order ? ar.env_order[eslot] : eslot) : 0;
                traj = (size_t)actor * ar.n_envs + env;
                if (valid) {
                    env_bind(e, ar, env, pv_base, traj);
                    if (!GUST && e.gust && ar.status) atomicOr(ar.status, SERL_STATUS_GUST_FLAG);
                    if (from_h) {
                        const long long hx = (slot - 1) * slot_threads + wslot * 32 + lane;
#pragma unroll
                        for (int i = 0; i < NX; ++i) e.X[i] = __ldcg(ar.ho.X + (size_t)i * ar.ho.n + hx);
                        e.t = __ldcg(ar.ho.t + hx); e.ret = __ldcg(ar.ho.ret + hx);
#pragma unroll
                        for (int i = 0; i < 7; ++i) obs[i] = __ldcg(ar.ho.obs + (size_t)i * ar.ho.n + hx);
                        const int kk = __ldcg(ar.ho.k + hx);
                        e.k = kk & 0x3fffffff; e.done = ((kk >> 30) & 1) != 0;
                    } else {
                        env_reset<TABS>(e, ar, env, obs, traj);
                    }
                } else {
                    e.done = true; e.k = 0; e.ret = 0.0; e.t = 0.0; e.fault = 0; e.gust = 0; e.pv = pv_base; e.pv_post = nullptr; e.theta_trim = 0.0;
                    e.ref_lv = ar.ref_levels; e.ref_st = ar.ref_starts;
#pragma unroll
                    for (int i = 0; i < NX; ++i) e.X[i] = 0.0;
#pragma unroll
                    for (int i = 0; i < 7; ++i) obs[i] = 0.f;
                }
                replay = valid && ar.replay != nullptr && env == ar.replay_env;
                in_seg = true;
            }
        }
        // the CTA's warps meet here once per step; the launch ends when no slot has a segment left
        if (!__syncthreads_or(in_seg || pending)) break;
        const bool mine = in_seg && !e.done && e.k < ke;
        if (__any_sync(0xffffffffu, mine)) {
            // one instantiation per activation: the choice is compiled into the 4 x h/4 activation calls of every layer
            if (actfn == SERL_ACT_TANH) actor_forward_warp<H, SERL_ACT_TANH>(w, L, lane, obs, a);
            else if (actfn == SERL_ACT_ELU) actor_forward_warp<H, SERL_ACT_ELU>(w, L, lane, obs, a);
            else actor_forward_warp<H, SERL_ACT_LEAKY_RELU>(w, L, lane, obs, a);
            // with the stage derivatives in tensor memory the plant's transfers are warp-collective: every lane steps,
            // lanes whose trajectory is over change nothing
            if (use_tmem) env_step<TABS, true, GUST>(e, ar, traj, actor, replay, a, obs, mine, taddr);
            else if (mine) env_step<TABS, false, GUST>(e, ar, traj, actor, replay, a, obs);
        }
    }
    if (use_tmem) {                                    // every warp has left the loop (it ends at a CTA barrier)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- cross-check kernel: every thread evaluates the whole MLP for its own env (any hidden size that fits) ------
__global__ void __launch_bounds__(128)
rollout_kernel_simple(RolloutArgs ar)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* w = reinterpret_cast<float*>(smem_raw);
    const int P4 = (ar.P + 3) & ~3;
    float* bufA = w + P4;
    float* bufB = bufA + ar.sh.hidden * 128;
    const int actor = blockIdx.y, tid = threadIdx.x;
    const int eslot = blockIdx.x * 128 + tid;
    const float* gw = ar.weights + (size_t)actor * ar.P;
    for (int i = tid; i < ar.P; i += 128) w[i] = gw[i];
    __syncthreads();
    if (eslot >= ar.n_envs) return;
    const int env = ar.env_order ? ar.env_order[eslot] : eslot;
    Env e;
    e.tab = plant_tables_blob;
    float obs[7], a[3];
    env_bind(e, ar, env, &plant_pv[0][0], (size_t)actor * ar.n_envs + env);
    env_reset(e, ar, env, obs, (size_t)actor * ar.n_envs + env);
    const size_t traj = (size_t)actor * ar.n_envs + env;
    const bool replay = ar.replay != nullptr && env == ar.replay_env;
    while (!e.done) {
        actor_forward_simple(w, ar.sh, bufA, bufB, tid, 128, obs, a);
        env_step<false, false, true>(e, ar, traj, actor, replay, a, obs);       // (the gust schedule costs nothing that matters here)
    }
    ar.returns[traj] = e.ret;
    ar.steps[traj] = e.k;
    if (ar.status && !isfinite(e.ret + e.X[3] + e.X[7] + e.X[9])) atomicOr(ar.status, SERL_STATUS_NONFINITE);   // NaN actions poison the state at once
}

// ---- Actor.forward for a batch of observations (same device functions as the rollout) --------------------------
template <int H>
__global__ void __launch_bounds__(128)
actor_forward_kernel(const float* __restrict__ genome, int P, serl_actor_shape sh, const float* __restrict__ obs_in, int n,
                     float* __restrict__ act_out)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* w = reinterpret_cast<float*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < P; i += 128) w[genome_layout_index(i, sh.state_dim, H, sh.num_layers)] = genome[i];
    __syncthreads();
    const int base = (blockIdx.x * 128 + (tid & ~31));
    if (base >= n) return;
    const int i = base + lane;
    float obs[7], a[3];
#pragma unroll
    for (int k = 0; k < 7; ++k) obs[k] = i < n ? obs_in[(size_t)i * 7 + k] : 0.f;
    if (sh.activation == SERL_ACT_TANH) actor_forward_warp<H, SERL_ACT_TANH>(w, sh.num_layers, lane, obs, a);
    else if (sh.activation == SERL_ACT_ELU) actor_forward_warp<H, SERL_ACT_ELU>(w, sh.num_layers, lane, obs, a);
    else actor_forward_warp<H, SERL_ACT_LEAKY_RELU>(w, sh.num_layers, lane, obs, a);
    if (i < n) { act_out[(size_t)i * 3] = a[0]; act_out[(size_t)i * 3 + 1] = a[1]; act_out[(size_t)i * 3 + 2] = a[2]; }
}

__global__ void __launch_bounds__(128)
actor_forward_kernel_simple(const float* __restrict__ genome, int P, serl_actor_shape sh, const float* __restrict__ obs_in, int n,
                            float* __restrict__ act_out)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* w = reinterpret_cast<float*>(smem_raw);
    float* bufA = w + ((P + 3) & ~3);
    float* bufB = bufA + sh.hidden * 128;
    const int tid = threadIdx.x;
    for (int i = tid; i < P; i += 128) w[i] = genome[i];
    __syncthreads();
    const int i = blockIdx.x * 128 + tid;
    if (i >= n) return;
    float obs[7], a[3];
    for (int k = 0; k < 7; ++k) obs[k] = obs_in[(size_t)i * 7 + k];
    actor_forward_simple(w, sh, bufA, bufB, tid, 128, obs, a);
    act_out[(size_t)i * 3] = a[0]; act_out[(size_t)i * 3 + 1] = a[1]; act_out[(size_t)i * 3 + 2] = a[2];
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
__global__ void plant_step_kernel(double* __restrict__ X, const double* __restrict__ cmd, const int* __restrict__ variant,
                                  const int* __restrict__ call, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x[NX], u[3];
#pragma unroll
    for (int k = 0; k < NX; ++k) x[k] = X[(size_t)i * NX + k];
    u[0] = cmd[3 * i]; u[1] = cmd[3 * i + 1]; u[2] = cmd[3 * i + 2];
    const int post = (variant[i] >> 16) & 0xff;
    plant_step<false, false, true>(plant_pv[variant[i] & 0xff], x, u, plant_tables_blob, false, post ? plant_pv[post] : nullptr,
                                   (call ? call[i] : 0) | ((variant[i] & SERL_MODE_GUST) ? PLANT_CALL_GUST : 0) |
                                       ((variant[i] & SERL_MODE_GUST_UP) ? PLANT_CALL_GUST_UP : 0));
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
    plant_step_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(d_X, d_cmd, d_variant, nullptr, n);
    serl_count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SERL_OK : serl_fail_cuda(e, "plant_step_kernel");
}

extern "C" int serl_plant_step_timed(double* d_X, const double* d_cmd, const int32_t* d_variant, const int32_t* d_call, int32_t n, void* stream)
{
    if (!d_X || !d_cmd || !d_variant || !d_call || n <= 0) return serl_fail(SERL_ERR_ARG, "serl_plant_step_timed: bad argument");
    plant_step_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(d_X, d_cmd, d_variant, d_call, n);
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

// ---- K6 (fast path): the same metric through a Bluestein (chirp-z) FFT --------------------------------------------
// N (the episode length) is arbitrary (2001 = 3*23*29 for a full episode, anything for an early termination), so the
// length-N DFT is written as a circular convolution of size FM = 4096 >= 2N-1 with the chirp b[m] = exp(i pi m^2 / N):
//   Y[k] = conj(b[k]) * sum_n (y[n] conj(b[n])) b[k-n]
// = three FFTs of size 4096 = 16^3 in shared memory per transform (three radix-16 passes each; forward DIF: natural -> digit-reversed order; the
// pointwise product with the chirp spectrum in digit-reversed order; inverse DIT: digit-reversed -> natural), O(N log N)
// instead of the O(N^2) of the direct form.  Two real channels share one complex transform (their spectra are separated by
// conjugate symmetry), the channel means are removed first (bin 0 is not part of the metric), phases are reduced exactly
// in integer arithmetic (m^2 mod 2N).  One CTA per trajectory.
#define FM 4096
#define FLOG 12
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 chirp(int m, int N)      // exp(+i pi m^2 / N)
{
    const int r = (int)(((long long)m * m) % (2 * N));
    float s, c;
    sincospif((float)r / (float)N, &s, &c);
    return make_float2(c, s);
}
// tw[j] = exp(-2 pi i j / FM) for j < FM/2; the upper half of the circle is the negated lower half
__device__ __forceinline__ float2 twiddle(const float2* tw, int j)
{
    const float2 t = tw[j & (FM / 2 - 1)];
    return (j & (FM / 2)) ? make_float2(-t.x, -t.y) : t;
}
// The work arrays are padded by one element per 16 (index i lives at ZI(i)): in the last pass a thread owns 16 CONSECUTIVE
// elements, and without the padding all 32 lanes of a warp would hit the same banks.
#define ZI(i) ((i) + ((i) >> 4))
#define ZN (FM + FM / 16)
__device__ __forceinline__ void dft4(float2 a, float2 b, float2 c, float2 d, float2& x0, float2& x1, float2& x2, float2& x3)
{
    const float2 t0 = make_float2(a.x + c.x, a.y + c.y), t1 = make_float2(a.x - c.x, a.y - c.y);
    const float2 t2 = make_float2(b.x + d.x, b.y + d.y), t3 = make_float2(b.x - d.x, b.y - d.y);
    x0 = make_float2(t0.x + t2.x, t0.y + t2.y); x2 = make_float2(t0.x - t2.x, t0.y - t2.y);     // X1 = t1 - i t3, X3 = t1 + i t3
    x1 = make_float2(t1.x + t3.y, t1.y - t3.x); x3 = make_float2(t1.x - t3.y, t1.y + t3.x);
}
__device__ __forceinline__ void idft4(float2 a, float2 b, float2 c, float2 d, float2& x0, float2& x1, float2& x2, float2& x3)
{
    const float2 t0 = make_float2(a.x + c.x, a.y + c.y), t1 = make_float2(a.x - c.x, a.y - c.y);
    const float2 t2 = make_float2(b.x + d.x, b.y + d.y), t3 = make_float2(b.x - d.x, b.y - d.y);
    x0 = make_float2(t0.x + t2.x, t0.y + t2.y); x2 = make_float2(t0.x - t2.x, t0.y - t2.y);     // x1 = t1 + i t3, x3 = t1 - i t3
    x1 = make_float2(t1.x - t3.y, t1.y + t3.x); x3 = make_float2(t1.x + t3.y, t1.y - t3.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) { return cmul(a, make_float2(b.x, -b.y)); }      // a * conj(b)
// forward FFT, decimation in frequency, natural order in, base-4 digit-reversed order out.  FM = 16^3: THREE passes over
// shared memory, each thread (256 of them) transforms 16 elements in registers per pass = two radix-4 stages back to back
// (stage A: span 16q, butterflies over a; stage B: span 4q, butterflies over r; element (a, r) at base + (4a + r) q).
__device__ void fft_dif(float2* z, const float2* tw, int tid)
{
#pragma unroll 1
    for (int lq = FLOG - 4; lq >= 0; lq -= 4) {
        const int q = 1 << lq, tA = FM >> (lq + 4), tB = FM >> (lq + 2);
        const int pos = tid & (q - 1), base = ((tid >> lq) << (lq + 4)) + pos;
        float2 v[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float2 x0, x1, x2, x3;
            dft4(z[ZI(base + r * q)], z[ZI(base + (4 + r) * q)], z[ZI(base + (8 + r) * q)], z[ZI(base + (12 + r) * q)], x0, x1, x2, x3);
            const int w = (pos + r * q) * tA;
            v[0][r] = x0; v[1][r] = cmul(x1, twiddle(tw, w)); v[2][r] = cmul(x2, twiddle(tw, 2 * w)); v[3][r] = cmul(x3, twiddle(tw, 3 * w));
        }
        const int wb = pos * tB;
        const float2 b1 = twiddle(tw, wb), b2 = twiddle(tw, 2 * wb), b3 = twiddle(tw, 3 * wb);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float2 x0, x1, x2, x3;
            dft4(v[a][0], v[a][1], v[a][2], v[a][3], x0, x1, x2, x3);
            z[ZI(base + (4 * a) * q)] = x0;
            z[ZI(base + (4 * a + 1) * q)] = cmul(x1, b1);
            z[ZI(base + (4 * a + 2) * q)] = cmul(x2, b2);
            z[ZI(base + (4 * a + 3) * q)] = cmul(x3, b3);
        }
        __syncthreads();
    }
}
// inverse FFT (unnormalised), decimation in time: digit-reversed order in, natural order out; the mirror image
__device__ void ifft_dit(float2* z, const float2* tw, int tid)
{
#pragma unroll 1
    for (int lq = 0; lq <= FLOG - 4; lq += 4) {
        const int q = 1 << lq, tA = FM >> (lq + 4), tB = FM >> (lq + 2);
        const int pos = tid & (q - 1), base = ((tid >> lq) << (lq + 4)) + pos;
        const int wb = pos * tB;
        const float2 b1 = twiddle(tw, wb), b2 = twiddle(tw, 2 * wb), b3 = twiddle(tw, 3 * wb);
        float2 v[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
            idft4(z[ZI(base + (4 * a) * q)], cmulc(z[ZI(base + (4 * a + 1) * q)], b1), cmulc(z[ZI(base + (4 * a + 2) * q)], b2),
                  cmulc(z[ZI(base + (4 * a + 3) * q)], b3), v[a][0], v[a][1], v[a][2], v[a][3]);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int w = (pos + r * q) * tA;
            float2 x0, x1, x2, x3;
            idft4(v[0][r], cmulc(v[1][r], twiddle(tw, w)), cmulc(v[2][r], twiddle(tw, 2 * w)), cmulc(v[3][r], twiddle(tw, 3 * w)), x0, x1, x2, x3);
            z[ZI(base + r * q)] = x0; z[ZI(base + (4 + r) * q)] = x1; z[ZI(base + (8 + r) * q)] = x2; z[ZI(base + (12 + r) * q)] = x3;
        }
        __syncthreads();
    }
}

// twiddles + the chirp-filter spectrum of a FULL episode (N = horizon), once per launch: most trajectories of a trained
// population run the whole horizon, and for them the filter transform is a fifth of the work
__global__ void __launch_bounds__(256)
smoothness_prep_kernel(int horizon, float2* __restrict__ tw_g, float2* __restrict__ hf_g, float2* __restrict__ cb_g)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    float2* hf = reinterpret_cast<float2*>(sm_raw);          // [ZN]
    float2* tw = hf + ZN;                                     // [FM/2]
    const int tid = threadIdx.x;
    for (int j = tid; j < FM / 2; j += 256) {
        float s, c;
        sincospif(-2.0f * (float)j / (float)FM, &s, &c);
        tw[j] = make_float2(c, s);
    }
    for (int m = tid; m < FM; m += 256) {
        const int d = m < horizon ? m : (FM - m < horizon ? FM - m : -1);
        hf[ZI(m)] = d >= 0 ? chirp(d, horizon) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    fft_dif(hf, tw, tid);
    for (int j = tid; j < FM / 2; j += 256) tw_g[j] = tw[j];
    for (int m = tid; m < FM; m += 256) hf_g[m] = hf[ZI(m)];
    for (int m = tid; m <= horizon; m += 256) cb_g[m] = chirp(m, horizon);       // b[m], m = 0 .. N
}

__global__ void __launch_bounds__(256)
smoothness_fft_kernel(const float* __restrict__ actions, const int* __restrict__ steps, int horizon, double dt, double* __restrict__ out,
                      const float2* __restrict__ tw_g, const float2* __restrict__ hf_g, const float2* __restrict__ cb_g)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    float2* z = reinterpret_cast<float2*>(sm_raw);            // [ZN] work buffer (padded index ZI)
    float2* hf = z + ZN;                                      // [ZN] spectrum of the chirp filter (digit-reversed order)
    float2* tw = hf + ZN;                                     // [FM/2] twiddles
    __shared__ double red[256];
    __shared__ float mean_s[3];
    const int traj = blockIdx.x, tid = threadIdx.x;
    const int N = steps[traj];
    const int Mb = N / 2 - 1;
    if (Mb <= 0) { if (tid == 0) out[traj] = -0.0; return; }
    const float* a = actions + (size_t)traj * horizon * 3;
    for (int j = tid; j < FM / 2; j += 256) tw[j] = tw_g[j];
    // channel means (bin 0 is excluded from the metric; removing it keeps the float32 transform accurate)
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int n = tid; n < N; n += 256) { s0 += a[3 * n]; s1 += a[3 * n + 1]; s2 += a[3 * n + 2]; }
    for (int c = 0; c < 3; ++c) {
        red[tid] = c == 0 ? s0 : (c == 1 ? s1 : s2);
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) { if (tid < s) red[tid] += red[tid + s]; __syncthreads(); }
        if (tid == 0) mean_s[c] = (float)(red[0] / (double)N);
        __syncthreads();
    }
    // chirp filter h[m] = b[|m|] for |m| < N (circular), its forward transform stays in hf (full episodes: precomputed)
    if (N == horizon) {
        for (int m = tid; m < FM; m += 256) hf[ZI(m)] = hf_g[m];
        __syncthreads();
    } else {
        for (int m = tid; m < FM; m += 256) {
            const int d = m < N ? m : (FM - m < N ? FM - m : -1);
            hf[ZI(m)] = d >= 0 ? chirp(d, N) : make_float2(0.f, 0.f);
        }
        __syncthreads();
        fft_dif(hf, tw, tid);
    }
    const double fstep = Mb > 1 ? (1.0 / (2.0 * dt) - dt) / (double)(Mb - 1) : 0.0;
    const float inv_m = 1.0f / (float)FM;
    const bool full = N == horizon;              // the chirp b[m] of a full episode comes from the per-launch table
#define K6_CHIRP(m) (full ? cb_g[m] : chirp((m), N))
    double acc = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        for (int n = tid; n < FM; n += 256) {
            float2 v = make_float2(0.f, 0.f);
            if (n < N) {
                const float re = pass == 0 ? a[3 * n] - mean_s[0] : a[3 * n + 2] - mean_s[2];
                const float im = pass == 0 ? a[3 * n + 1] - mean_s[1] : 0.f;
                const float2 b = K6_CHIRP(n);
                v = cmul(make_float2(re, im), make_float2(b.x, -b.y));
            }
            z[ZI(n)] = v;
        }
        __syncthreads();
        fft_dif(z, tw, tid);
        for (int m = tid; m < FM; m += 256) z[ZI(m)] = cmul(z[ZI(m)], hf[ZI(m)]);
        __syncthreads();
        ifft_dit(z, tw, tid);
        for (int k = 1 + tid; k <= Mb; k += 256) {
            const double f = dt + (double)(k - 1) * fstep;
            if (pass == 0) {
                // T[k] = conj(b[k]) c[k] = Y0[k] + i Y1[k];  Y0 = (T[k] + conj(T[N-k])) / 2,  Y1 = (T[k] - conj(T[N-k])) / (2i)
                const float2 bk = K6_CHIRP(k), bn = K6_CHIRP(N - k);
                float2 tk = cmul(z[ZI(k)], make_float2(bk.x, -bk.y)), tn = cmul(z[ZI(N - k)], make_float2(bn.x, -bn.y));
                tk.x *= inv_m; tk.y *= inv_m; tn.x *= inv_m; tn.y *= inv_m;
                const float y0r = 0.5f * (tk.x + tn.x), y0i = 0.5f * (tk.y - tn.y);
                const float y1r = 0.5f * (tk.y + tn.y), y1i = 0.5f * (tn.x - tk.x);
                acc += f * ((double)y0r * y0r + (double)y0i * y0i + (double)y1r * y1r + (double)y1i * y1i);
            } else {
                const float cr = z[ZI(k)].x * inv_m, ci = z[ZI(k)].y * inv_m;
                acc += f * ((double)cr * cr + (double)ci * ci);
            }
        }
        __syncthreads();
    }
    red[tid] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if (tid < s) red[tid] += red[tid + s]; __syncthreads(); }
    if (tid == 0) {
        const double S = red[0] * dt * 2.0 / (double)N;
        out[traj] = -(sqrt(S) * 100.0 * (80.0 / ((double)N * dt)));
    }
}

static cudaError_t scratch_get(cudaStream_t s, size_t bytes, void** out, int which);
extern "C" int serl_smoothness(const float* d_actions, const int32_t* d_steps, int32_t n_traj, int32_t horizon, double dt,
                               double* d_out, void* stream)
{
    if (!d_actions || !d_steps || !d_out || n_traj <= 0 || horizon <= 0) return serl_fail(SERL_ERR_ARG, "serl_smoothness: bad argument");
    static int force_direct = -1;
    if (force_direct < 0) { const char* v = getenv("SERL_SMOOTHNESS_IMPL"); force_direct = (v && strcmp(v, "direct") == 0) ? 1 : 0; }
    cudaError_t e;
    if (!force_direct && 2 * horizon - 1 <= FM) {
        // episodes of up to 2048 steps (training: 2001): Bluestein FFT, O(N log N)
        const size_t smem = (size_t)(2 * ZN + FM / 2) * sizeof(float2), smem_prep = (size_t)(ZN + FM / 2) * sizeof(float2);
        e = cudaFuncSetAttribute(smoothness_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep);
        if (e != cudaSuccess) return serl_fail_cuda(e, "cudaFuncSetAttribute(smoothness_prep)");
        e = cudaFuncSetAttribute(smoothness_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return serl_fail_cuda(e, "cudaFuncSetAttribute(smoothness_fft)");
        void* tabs = nullptr;
        e = scratch_get((cudaStream_t)stream, (size_t)(FM + FM / 2 + FM) * sizeof(float2), &tabs, 1);
        if (e != cudaSuccess) return serl_fail_cuda(e, "smoothness scratch");
        float2* tw_g = (float2*)tabs;
        float2* hf_g = tw_g + FM / 2;
        float2* cb_g = hf_g + FM;
        smoothness_prep_kernel<<<1, 256, smem_prep, (cudaStream_t)stream>>>(horizon, tw_g, hf_g, cb_g);
        serl_count_launch();
        smoothness_fft_kernel<<<n_traj, 256, smem, (cudaStream_t)stream>>>(d_actions, d_steps, horizon, dt, d_out, tw_g, hf_g, cb_g);
        serl_count_launch();
        e = cudaGetLastError();
        return e == cudaSuccess ? SERL_OK : serl_fail_cuda(e, "smoothness_fft_kernel");
    }
    // longer episodes (80 s evaluation mode: 8001 steps): direct DFT
    const size_t smem = (size_t)horizon * (8 + 12);
    if (smem > 200 * 1024) return serl_fail(SERL_ERR_UNSUPPORTED, "serl_smoothness: horizon too long for the shared-memory DFT");
    e = cudaFuncSetAttribute(smoothness_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
static int env_int(const char* name)
{
    const char* v = getenv(name);
    return v ? atoi(v) : 0;
}

static int device_sms()
{
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    }
    return num_sms;
}

// CTA shape of the persistent kernel: `apc` genome slots x `wps` warps.  A slot's warps fly wps*32 envs of one actor;
// the estimate below is (rounds of work per slot) x (time of one step with apc*wps resident warps per SM), the latter a
// linear fit of measurements at 4 and 8 warps (profiles/): a lone warp steps 1.4x faster than one of eight.
static void choose_shape(int pop, int n_envs, int apc_max, int sms, int* apc_out, int* wps_out)
{
    double best = 1e300;
    *apc_out = 1; *wps_out = 1;
    const int max_wps = (n_envs + 31) / 32;
    for (int wps = 4; wps >= 1; wps >>= 1) {
        if (wps > max_wps && wps > 1) continue;
        const long long chunks = (n_envs + wps * 32 - 1) / (wps * 32);
        const long long nt = (long long)pop * chunks;
        for (int apc = apc_max; apc >= 1; --apc) {
            if (apc * wps * 32 > MAX_CTA_THREADS) continue;
            const long long grid = nt / apc < sms ? (nt + apc - 1) / apc : sms;
            const long long ns = grid * apc;
            const double rounds = nt <= ns ? 1.0 : (double)nt / (double)ns;
            const double lanes = (double)chunks * wps * 32 / n_envs;        // idle-lane overhead of a ragged last chunk
            const double est = rounds * (13.2 + 0.825 * apc * wps) * (lanes > 1.0 ? 1.0 + 0.2 * (lanes - 1.0) : 1.0);
            if (est < best - 1e-9) { best = est; *apc_out = apc; *wps_out = wps; }
        }
    }
}

// Scratch of a launch (genomes in the shared-memory layout + hand-over records): one grow-only buffer per (device,
// stream), kept for the life of the process.  Launches on one stream are ordered, so they can share it; launches on
// different streams (the Agent's side-stream episodes next to the population rollout) get their own.  No stream-ordered
// allocator here: growing its pool maps memory, which waits for kernels in flight on OTHER streams.
#include <map>
#include <mutex>
struct ScratchBuf { void* p; size_t bytes; };
static std::mutex g_scratch_mu;
static std::map<std::pair<std::pair<int, int>, cudaStream_t>, ScratchBuf> g_scratch;
static cudaError_t scratch_get(cudaStream_t s, size_t bytes, void** out, int which)      // which: 0 = K1, 1 = K6 tables
{
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    ScratchBuf& b = g_scratch[std::make_pair(std::make_pair(dev, which), s)];
    if (b.bytes < bytes) {
        if (b.p) { cudaStreamSynchronize(s); cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
        const size_t want = bytes + bytes / 4;
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) return e;
        b.bytes = want;
    }
    *out = b.p;
    return cudaSuccess;
}

template <int H, bool TABS, bool GUST = false>
static cudaError_t launch_persist(RolloutArgs& ar, int apc_max, cudaStream_t s, void** scratch)
{
    constexpr int TABN2 = (PT_TOTAL + SERL_PLANT_COUNT * PLANT_NPV + 1) & ~1;
    const int sms = ar.sm_limit > 0 && ar.sm_limit < device_sms() ? ar.sm_limit : device_sms();
    int apc, wps;
    choose_shape(ar.pop, ar.n_envs, apc_max, sms, &apc, &wps);
    static int f_apc = -1, f_wps = -1;       // experiment knobs
    if (f_apc < 0) { f_apc = env_int("SERL_ROLLOUT_APC"); f_wps = env_int("SERL_ROLLOUT_WPS"); }
    if (f_apc > 0 && f_apc <= apc_max) apc = f_apc;
    if (f_wps == 1 || f_wps == 2 || f_wps == 4) wps = f_wps;
    while (apc * wps * 32 > MAX_CTA_THREADS) --apc;
    ar.apc = apc; ar.wps = wps;
    ar.n_chunks = (ar.n_envs + wps * 32 - 1) / (wps * 32);
    ar.n_tasks = (long long)ar.pop * ar.n_chunks;
    const long long grid = ar.n_tasks / apc < sms ? (ar.n_tasks + apc - 1) / apc : sms;
    ar.n_slots = grid * apc;
    // scratch: genomes in the shared-memory layout + hand-over records of the time-split schedule (stream-ordered)
    const size_t wt_bytes = (size_t)ar.pop * ar.P4 * 4;
    const long long hn = ar.n_tasks > ar.n_slots ? ar.n_slots * wps * 32 : 0;
    const size_t ho_bytes = (size_t)hn * (NX * 8 + 8 + 8 + 7 * 4 + 4) + (size_t)(hn / 32) * 4;
    cudaError_t e = scratch_get(s, wt_bytes + ho_bytes + 512, scratch, 0);
    if (e != cudaSuccess) return e;
    unsigned char* base = (unsigned char*)*scratch;
    float* wt = (float*)base;
    ar.wt = wt;
    ar.ho.n = hn;
    if (hn) {
        unsigned char* p = base + ((wt_bytes + 255) & ~(size_t)255);
        ar.ho.X = (double*)p; p += (size_t)hn * NX * 8;
        ar.ho.t = (double*)p; p += (size_t)hn * 8;
        ar.ho.ret = (double*)p; p += (size_t)hn * 8;
        ar.ho.obs = (float*)p; p += (size_t)hn * 7 * 4;
        ar.ho.k = (int*)p; p += (size_t)hn * 4;
        ar.ho.flag = (int*)p;
        e = cudaMemsetAsync(ar.ho.flag, 0, (size_t)(hn / 32) * 4, s);
        if (e != cudaSuccess) return e;
    }
    const long long n = (long long)ar.pop * ar.P;
    const int lay_grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
    genome_layout_kernel<<<lay_grid, 256, 0, s>>>(ar.weights, wt, ar.pop, ar.P, ar.P4, ar.sh.state_dim, H, ar.sh.num_layers);
    serl_count_launch();
    const size_t smem = (TABS ? (size_t)TABN2 * sizeof(real) : 0) + (size_t)apc * ar.P4 * 4;
    e = cudaFuncSetAttribute(rollout_kernel_persist<H, TABS, GUST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    rollout_kernel_persist<H, TABS, GUST><<<(unsigned)grid, apc * wps * 32, smem, s>>>(ar);
    serl_count_launch();
    return cudaGetLastError();
}

static int rollout_impl(const serl_rollout_desc& d, void* stream)
{
    const serl_actor_shape* shape = &d.shape;
    if (!d.d_weights || !d.d_ref_levels || !d.d_ref_starts || !d.d_env_mode || !d.d_returns || !d.d_steps)
        return serl_fail(SERL_ERR_ARG, "serl_rollout: null pointer argument");
    if (d.pop <= 0 || d.n_envs <= 0 || d.horizon <= 0) return serl_fail(SERL_ERR_ARG, "serl_rollout: pop, n_envs, horizon must be > 0");
    if (d.pop > 65535) return serl_fail(SERL_ERR_ARG, "serl_rollout: pop must be <= 65535 per call");
    if (shape->state_dim != 7 || shape->action_dim != 3)
        return serl_fail(SERL_ERR_ARG, "serl_rollout: PH-LAB attitude task needs state_dim=7, action_dim=3");
    if (shape->hidden < 2 || shape->hidden > 256 || shape->num_layers < 0 || shape->activation < 0 || shape->activation > 2)
        return serl_fail(SERL_ERR_ARG, "serl_rollout: unsupported actor shape");
    if (d.d_replay && (d.replay_env < 0 || d.replay_env >= d.n_envs)) return serl_fail(SERL_ERR_ARG, "serl_rollout: replay_env out of range");
    if (d.horizon >= (1 << 30)) return serl_fail(SERL_ERR_ARG, "serl_rollout: horizon too long");
    if (g_force_simple < 0) {
        const char* v = getenv("SERL_ROLLOUT_IMPL");
        g_force_simple = (v && strcmp(v, "simple") == 0) ? 1 : 0;
    }
    cudaStream_t s = (cudaStream_t)stream;
    RolloutArgs ar;
    memset(&ar, 0, sizeof(ar));
    ar.weights = d.d_weights; ar.P = (int)serl_actor_num_params(shape); ar.sh = *shape;
    ar.ref_levels = d.d_ref_levels; ar.ref_starts = d.d_ref_starts; ar.env_mode = d.d_env_mode; ar.n_envs = d.n_envs; ar.horizon = d.horizon;
    ar.action_noise = d.d_action_noise; ar.returns = d.d_returns; ar.steps = d.d_steps; ar.trace = d.d_trace; ar.actions = d.d_actions;
    ar.pop = d.pop;
    ar.t_max = d.t_max > 0.0 ? d.t_max : 20.0;
    ar.smooth_w = d.t_max > 0.0 ? d.smooth_width : 3.0;
    ar.env_order = d.d_env_order; ar.replay = d.d_replay; ar.replay_env = d.replay_env; ar.status = d.d_status; ar.sm_limit = d.sm_limit; ar.sensor_noise = d.d_sensor_noise;
    ar.P4 = (ar.P + 3) & ~3;
    const int H = shape->hidden;
    cudaError_t e;
    const size_t tab_bytes = (size_t)((PT_TOTAL + SERL_PLANT_COUNT * PLANT_NPV + 1) & ~1) * sizeof(real);
    const bool warp_ok = !g_force_simple && (H == 32 || H == 64 || H == 72 || H == 96 || H == 128) && (size_t)ar.P4 * 4 <= 227 * 1024;
    if (warp_ok) {
        // as many genome slots per CTA as shared memory holds next to the plant tables (h <= 72: two; h = 96: one);
        // h = 128 (207 KB genome) reads the tables through L1 instead
        const size_t budget = 227 * 1024 - 256;       // static shared memory + alignment of the dynamic part
        const bool tabs = tab_bytes + (size_t)ar.P4 * 4 <= budget;
        int apc_max = (int)(((tabs ? budget - tab_bytes : budget)) / ((size_t)ar.P4 * 4));
        if (apc_max > 4) apc_max = 4;
        if (apc_max > 2 && H > 32) apc_max = 2;
        void* scratch = nullptr;
        const bool gust = (d.flags & SERL_ROLLOUT_GUST) != 0;
#define K1_LAUNCH(HH, TT) (gust ? launch_persist<HH, TT, true>(ar, apc_max, s, &scratch) : launch_persist<HH, TT, false>(ar, apc_max, s, &scratch))
        if (H == 32) e = K1_LAUNCH(32, true);
        else if (H == 64) e = K1_LAUNCH(64, true);
        else if (H == 72) e = K1_LAUNCH(72, true);
        else if (H == 96) e = K1_LAUNCH(96, true);
        else e = tabs ? K1_LAUNCH(128, true) : K1_LAUNCH(128, false);
#undef K1_LAUNCH
    } else {
        const size_t smem = (size_t)ar.P4 * 4 + 2ull * H * 128 * 4;
        if (smem > 227 * 1024) return serl_fail(SERL_ERR_UNSUPPORTED, "serl_rollout: genome + activations exceed 227 KB of shared memory");
        e = cudaFuncSetAttribute(rollout_kernel_simple, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return serl_fail_cuda(e, "cudaFuncSetAttribute(rollout)");
        dim3 grid((d.n_envs + 127) / 128, d.pop);
        rollout_kernel_simple<<<grid, 128, smem, s>>>(ar);
        serl_count_launch();
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) return serl_fail_cuda(e, "rollout_kernel launch");
    if (d.d_fitness) {
        fitness_mean_kernel<<<(d.pop + 127) / 128, 128, 0, s>>>(d.d_returns, d.pop, d.n_envs, d.d_fitness);
        serl_count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return serl_fail_cuda(e, "fitness_mean_kernel launch");
    }
    return SERL_OK;
}

int rollout_tc_impl(const serl_rollout_desc& d, const int32_t* widths, int n_widths, void* stream);     // rollout_tc.cu

extern "C" int serl_rollout_run(const serl_rollout_desc* desc, void* stream)
{
    if (!desc) return serl_fail(SERL_ERR_ARG, "serl_rollout_run: null descriptor");
    if (desc->t_max > 0.0 && !(desc->smooth_width > 0.0)) return serl_fail(SERL_ERR_ARG, "serl_rollout_run: smooth_width must be > 0");
    if (desc->n_widths > 0) {
        if (!desc->widths) return serl_fail(SERL_ERR_ARG, "serl_rollout_run: n_widths > 0 but widths is null");
        const int rc = rollout_tc_impl(*desc, desc->widths, desc->n_widths, stream);
        if (rc != SERL_OK) return rc;
        if (desc->d_fitness) {
            fitness_mean_kernel<<<(desc->pop + 127) / 128, 128, 0, (cudaStream_t)stream>>>(desc->d_returns, desc->pop, desc->n_envs, desc->d_fitness);
            serl_count_launch();
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return serl_fail_cuda(e, "fitness_mean_kernel launch");
        }
        return SERL_OK;
    }
    return rollout_impl(*desc, stream);
}

static serl_rollout_desc make_desc(const float* d_weights, int32_t pop, const serl_actor_shape* shape,
                                   const double* d_ref_levels, const double* d_ref_starts, const int32_t* d_env_mode,
                                   int32_t n_envs, int32_t horizon, const float* d_action_noise,
                                   double* d_returns, int32_t* d_steps, double* d_fitness, double* d_trace, float* d_actions)
{
    serl_rollout_desc d;
    memset(&d, 0, sizeof(d));
    d.d_weights = d_weights; d.pop = pop; if (shape) d.shape = *shape;
    d.d_ref_levels = d_ref_levels; d.d_ref_starts = d_ref_starts; d.d_env_mode = d_env_mode; d.n_envs = n_envs; d.horizon = horizon;
    d.d_action_noise = d_action_noise; d.d_returns = d_returns; d.d_steps = d_steps; d.d_fitness = d_fitness; d.d_trace = d_trace;
    d.d_actions = d_actions;
    return d;
}

extern "C" int serl_rollout(const float* d_weights, int32_t pop, const serl_actor_shape* shape,
                            const double* d_ref_levels, const double* d_ref_starts, const int32_t* d_env_mode,
                            int32_t n_envs, int32_t horizon, const float* d_action_noise,
                            double* d_returns, int32_t* d_steps, double* d_fitness, double* d_trace, float* d_actions, void* stream)
{
    if (!shape) return serl_fail(SERL_ERR_ARG, "serl_rollout: null pointer argument");
    return rollout_impl(make_desc(d_weights, pop, shape, d_ref_levels, d_ref_starts, d_env_mode, n_envs, horizon, d_action_noise,
                                  d_returns, d_steps, d_fitness, d_trace, d_actions), stream);
}

extern "C" int serl_rollout_eval(const float* d_weights, int32_t pop, const serl_actor_shape* shape,
                                 const double* d_ref_levels, const double* d_ref_starts, const int32_t* d_env_mode,
                                 int32_t n_envs, int32_t horizon, const float* d_action_noise,
                                 double* d_returns, int32_t* d_steps, double* d_fitness, double* d_trace, float* d_actions,
                                 double t_max, double smooth_width, void* stream)
{
    if (!shape) return serl_fail(SERL_ERR_ARG, "serl_rollout_eval: null pointer argument");
    if (!(t_max > 0.0) || !(smooth_width > 0.0)) return serl_fail(SERL_ERR_ARG, "serl_rollout_eval: t_max and smooth_width must be > 0");
    serl_rollout_desc d = make_desc(d_weights, pop, shape, d_ref_levels, d_ref_starts, d_env_mode, n_envs, horizon, d_action_noise,
                                    d_returns, d_steps, d_fitness, d_trace, d_actions);
    d.t_max = t_max; d.smooth_width = smooth_width;
    return rollout_impl(d, stream);
}

extern "C" int serl_actor_forward(const float* d_genome, const serl_actor_shape* shape, const float* d_obs, int32_t n,
                                  float* d_actions, void* stream)
{
    if (!d_genome || !shape || !d_obs || !d_actions || n <= 0) return serl_fail(SERL_ERR_ARG, "serl_actor_forward: bad argument");
    if (shape->state_dim != 7 || shape->action_dim != 3 || shape->hidden < 2 || shape->hidden > 256 || shape->num_layers < 0 ||
        shape->activation < 0 || shape->activation > 2)
        return serl_fail(SERL_ERR_ARG, "serl_actor_forward: unsupported actor shape");
    cudaStream_t s = (cudaStream_t)stream;
    const int P = (int)serl_actor_num_params(shape), H = shape->hidden;
    const int grid = (n + 127) / 128;
    cudaError_t e = cudaSuccess;
    const size_t smem = (size_t)((P + 3) & ~3) * 4;
#define AF_LAUNCH(HH) do { e = cudaFuncSetAttribute(actor_forward_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e == cudaSuccess) { actor_forward_kernel<HH><<<grid, 128, smem, s>>>(d_genome, P, *shape, d_obs, n, d_actions); e = cudaGetLastError(); } } while (0)
    if (g_force_simple < 0) {
        const char* v = getenv("SERL_ROLLOUT_IMPL");
        g_force_simple = (v && strcmp(v, "simple") == 0) ? 1 : 0;
    }
    if (!g_force_simple && H == 32) AF_LAUNCH(32);
    else if (!g_force_simple && H == 64) AF_LAUNCH(64);
    else if (!g_force_simple && H == 72) AF_LAUNCH(72);
    else if (!g_force_simple && H == 96) AF_LAUNCH(96);
    else if (!g_force_simple && H == 128) AF_LAUNCH(128);
    else {
        const size_t sm2 = smem + 2ull * H * 128 * 4;
        if (sm2 > 227 * 1024) return serl_fail(SERL_ERR_UNSUPPORTED, "serl_actor_forward: genome + activations exceed shared memory");
        e = cudaFuncSetAttribute(actor_forward_kernel_simple, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
        if (e == cudaSuccess) { actor_forward_kernel_simple<<<grid, 128, sm2, s>>>(d_genome, P, *shape, d_obs, n, d_actions); e = cudaGetLastError(); }
    }
#undef AF_LAUNCH
    serl_count_launch();
    return e == cudaSuccess ? SERL_OK : serl_fail_cuda(e, "actor_forward_kernel");
}

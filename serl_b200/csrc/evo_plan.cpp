// Native host planner for the neuro-evolution step: the draw-for-draw equivalent of the two heavy loops of
// serl_b200/evo.py:plan_epoch (classic crossover, base/core/mod_neuro_evo.py:516-523 + :61-93, and mutation, :537-539 +
// :329-369), consuming the SAME random streams as the Python / reference code:
//   * CPython's `random` module: MT19937 + random() (53-bit), _randbelow_with_getrandbits, randint/randrange/choice,
//     gauss() with its cached second variate (Lib/random.py), libm cos/sin/log/sqrt (the same glibc Python links);
//   * NumPy's legacy global RandomState: MT19937 + rk_double for np.random.uniform(0, 1, n).
// The generator states are imported from random.getstate() / np.random.get_state() and handed back afterwards, so the
// rest of the program continues on exactly the stream position the reference would be at.
// No CUDA here; it lives in the same library so that the engine stays one .so.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/serl_b200.h"
#include "common.cuh"

namespace {

struct MT {
    uint32_t mt[624];
    int idx;
    uint32_t next()
    {
        if (idx >= 624) {
            static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
            int kk;
            uint32_t y;
            for (kk = 0; kk < 624 - 397; kk++) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u];
            }
            for (; kk < 623; kk++) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u];
            }
            y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 1u];
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    double random53()   // random.random() and numpy's rk_double
    {
        const uint32_t a = next() >> 5, b = next() >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
};

struct PyRandom {
    MT g;
    bool has_gauss;
    double gauss_next;
    uint32_t randbelow(uint32_t n)   // Random._randbelow_with_getrandbits, n >= 1 (n == 0 is rejected by serl_plan_create)
    {
        if (n == 0) return 0;
        int k = 0;
        for (uint32_t v = n; v; v >>= 1) ++k;      // n.bit_length()
        uint32_t r = g.next() >> (32 - k);
        while (r >= n) r = g.next() >> (32 - k);
        return r;
    }
    double random() { return g.random53(); }
    int randint(int a, int b) { return a + (int)randbelow((uint32_t)(b - a + 1)); }
    int randrange(int n) { return (int)randbelow((uint32_t)n); }
    double gauss01()                 // random.gauss(0, 1)
    {
        double z;
        if (has_gauss) {
            z = gauss_next;
            has_gauss = false;
        } else {
            const double x2pi = random() * 6.283185307179586;          // TWOPI = 2.0 * pi
            const double g2rad = sqrt(-2.0 * log(1.0 - random()));
            z = cos(x2pi) * g2rad;
            gauss_next = sin(x2pi) * g2rad;
            has_gauss = true;
        }
        return 0.0 + z * 1.0;
    }
};

struct Plan {
    std::vector<int32_t> pairs;      // [n_pairs, 6]
    std::vector<int32_t> ops;        // [n_ops, 3]
    std::vector<int32_t> seg;        // [n_seg, 3]
    std::vector<int32_t> m_off, m_kind;
    std::vector<float> m_z;
};

}  // namespace

extern "C" {

// Plan handle API ------------------------------------------------------------------------------------------
// py_state: 624 MT words + index (random.getstate()[1]); py_gauss: [has, value]; np_state: 624 words + pos
// table: [n_params, 3] = (offset, rows, cols) in parameters() order (cols == 0: 1-D)
// unselects: even-length list after padding; new_elitists / offsprings: choice pools; mut_order = index_rank[num_elitists:]
void* serl_plan_create(uint32_t* py_state, double* py_gauss, uint32_t* np_state,
                       const int32_t* table, int32_t n_params,
                       const int32_t* unselects, int32_t n_unselects,
                       const int32_t* new_elitists, int32_t n_new_elitists,
                       const int32_t* offsprings, int32_t n_offsprings,
                       const int32_t* mut_order, int32_t n_mut, double mutation_prob)
{
    // random.choice([]) raises IndexError in the reference (mod_neuro_evo.py:519-520); an empty choice pool next to a
    // non-empty crossover list is reported as a failed plan (NULL) instead of spinning in randbelow(0)
    if (n_unselects >= 2 && (n_new_elitists <= 0 || n_offsprings <= 0)) return nullptr;
    PyRandom R;
    memcpy(R.g.mt, py_state, 624 * sizeof(uint32_t));
    R.g.idx = (int)py_state[624];
    R.has_gauss = py_gauss[0] != 0.0;
    R.gauss_next = py_gauss[1];
    MT N;
    memcpy(N.mt, np_state, 624 * sizeof(uint32_t));
    N.idx = (int)np_state[624];

    Plan* P = new Plan();
    // classic crossover (mod_neuro_evo.py:518-523, crossover_inplace :61-93)
    for (int q = 0; q + 1 < n_unselects; q += 2) {
        const int i = unselects[q], j = unselects[q + 1];
        const int off_i = new_elitists[R.randrange(n_new_elitists)];    // random.choice
        const int off_j = offsprings[R.randrange(n_offsprings)];
        const int begin = (int)(P->ops.size() / 3);
        for (int k = 0; k < n_params; ++k) {
            const int off = table[3 * k], rows = table[3 * k + 1], cols = table[3 * k + 2];
            if (cols > 0) {
                const int n = R.randint(0, rows * 2);
                for (int t = 0; t < n; ++t) {
                    const int d = R.random() < 0.5 ? 0 : 1;
                    const int r = R.randrange(rows);
                    P->ops.push_back(off + r * cols); P->ops.push_back(cols); P->ops.push_back(d);
                }
            } else {
                const int n = R.randint(0, rows);
                for (int t = 0; t < n; ++t) {
                    const int d = R.random() < 0.5 ? 0 : 1;
                    const int r = R.randrange(rows);
                    P->ops.push_back(off + r); P->ops.push_back(1); P->ops.push_back(d);
                }
            }
        }
        const int cnt = (int)(P->ops.size() / 3) - begin;
        const int32_t row[6] = {i, j, off_i, off_j, begin, cnt};
        P->pairs.insert(P->pairs.end(), row, row + 6);
    }
    // mutation (:537-539, mutate_inplace :329-369)
    std::vector<double> probs(n_params);
    for (int a = 0; a < n_mut; ++a) {
        if (!(R.random() < mutation_prob)) continue;
        for (int k = 0; k < n_params; ++k) probs[k] = (0.0 + 1.0 * N.random53()) * 2;   // np.random.uniform(0, 1, n) * 2
        for (int k = 0; k < n_params; ++k) {
            const int off = table[3 * k], rows = table[3 * k + 1], cols = table[3 * k + 2];
            if (cols == 0) continue;
            if (R.random() < probs[k]) {
                const int n = R.randint(0, (int)ceil(0.1 * (double)(rows * cols)));
                const int begin = (int)P->m_off.size();
                for (int t = 0; t < n; ++t) {
                    const int d1 = R.randrange(rows);
                    const int d2 = R.randrange(cols);
                    const double r = R.random();
                    P->m_off.push_back(off + d1 * cols + d2);
                    P->m_kind.push_back(r < 0.05 ? 1 : (r < 0.1 ? 2 : 0));
                    P->m_z.push_back((float)R.gauss01());
                }
                if (n) {
                    P->seg.push_back(mut_order[a]); P->seg.push_back(begin); P->seg.push_back(n);
                }
            }
        }
    }
    memcpy(py_state, R.g.mt, 624 * sizeof(uint32_t));
    py_state[624] = (uint32_t)R.g.idx;
    py_gauss[0] = R.has_gauss ? 1.0 : 0.0;
    py_gauss[1] = R.gauss_next;
    memcpy(np_state, N.mt, 624 * sizeof(uint32_t));
    np_state[624] = (uint32_t)N.idx;
    return P;
}

void serl_plan_sizes(void* h, int64_t* out5)
{
    Plan* P = (Plan*)h;
    out5[0] = (int64_t)(P->pairs.size() / 6);
    out5[1] = (int64_t)(P->ops.size() / 3);
    out5[2] = (int64_t)(P->seg.size() / 3);
    out5[3] = (int64_t)P->m_off.size();
    out5[4] = 0;
}

void serl_plan_copy(void* h, int32_t* pairs, int32_t* ops, int32_t* seg, int32_t* m_off, int32_t* m_kind, float* m_z)
{
    Plan* P = (Plan*)h;
    if (!P->pairs.empty()) memcpy(pairs, P->pairs.data(), P->pairs.size() * 4);
    if (!P->ops.empty()) memcpy(ops, P->ops.data(), P->ops.size() * 4);
    if (!P->seg.empty()) memcpy(seg, P->seg.data(), P->seg.size() * 4);
    if (!P->m_off.empty()) {
        memcpy(m_off, P->m_off.data(), P->m_off.size() * 4);
        memcpy(m_kind, P->m_kind.data(), P->m_kind.size() * 4);
        memcpy(m_z, P->m_z.data(), P->m_z.size() * 4);
    }
}

void serl_plan_destroy(void* h) { delete (Plan*)h; }

}  // extern "C"

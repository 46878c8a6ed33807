/* ORACLE — test infrastructure only (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).
 *
 * CPU restatement (fp64, reference operation order, no FMA contraction) of the reference's native plant
 * /root/reference/envs/<variant>/_citation.cpython-38-x86_64-linux-gnu.so:
 *   plant_rhs   <- step() Outputs pass (@0x6030) + citation_to_python_derivatives (@0x5f60)
 *   plant_step  <- the inlined rt_ertODEUpdateContinuousStates in step(): Simulink fixed-step ode5
 *                  (Dormand-Prince, h = 0.01, 6 stages; SURVEY.md 2.3 "Solver")
 * and of the gym wrapper /root/reference/envs/phlabenv.py (episode loop: oracle_episode, below).
 * The generated headers under gen/ come from tools/lift (symbolic trace of the binary).
 * Pinned: bit-identical to the reference binary on 2000 random RHS evaluations and on the 15 logged
 * episodes under /root/reference/logs/wandb (tests/test_oracle_plant.py; golden vectors tests/golden). */
#include <math.h>
#include <stdbool.h>
#include <string.h>

typedef double real;
#define PLANT_FN static
#define PLANT_TABLE(name, n) static const double name[n]
#define PLANT_TAB(name) name
#define PLANT_IC(v) static const double plant_ic_##v[19]
#define PLANT_SQRT sqrt
#define PLANT_FABS fabs
#define PLANT_SIN sin
#define PLANT_COS cos
#define PLANT_TAN tan
#define PLANT_EXP exp
#define PLANT_LOG10 log10
#define PLANT_POW pow
#define PLANT_XARGS
#define PLANT_CONSTS(n) static const double plant_k[n]
#define PLANT_K(i) plant_k[i]
#include "plant_support.h"
#include "gen/plant_tables.h"
#include "gen/plant_consts.h"
#include "gen/plant_rhs_h2000_v90.h"
#include "gen/plant_rhs_ice.h"
#include "gen/plant_rhs_cg.h"
#include "gen/plant_rhs_cg_for.h"
#include "gen/plant_rhs_h2000_v150.h"
#include "gen/plant_rhs_h10000_v90.h"

#define NVARIANT 6
static const char* const names[NVARIANT] = {"h2000_v90", "ice", "cg", "cg_for", "h2000_v150", "h10000_v90"};
typedef void (*rhs_fn)(const real*, const real*, real*);
static const rhs_fn rhs_tab[NVARIANT] = {plant_rhs_h2000_v90, plant_rhs_ice, plant_rhs_cg, plant_rhs_cg_for,
                                         plant_rhs_h2000_v150, plant_rhs_h10000_v90};
static const double* const ic_tab[NVARIANT] = {plant_ic_h2000_v90, plant_ic_ice, plant_ic_cg, plant_ic_cg_for,
                                               plant_ic_h2000_v150, plant_ic_h10000_v90};

int plant_num_variants(void) { return NVARIANT; }
const char* plant_variant_name(int v) { return (v >= 0 && v < NVARIANT) ? names[v] : 0; }
void plant_get_ic(int v, double* X) { memcpy(X, ic_tab[v], 19 * sizeof(double)); }

void plant_rhs(int v, const double* X, const double* U, double* xdot)
{
    for (int i = 0; i < 19; ++i) xdot[i] = 0.0;
    rhs_tab[v](X, U, xdot);
}

/* one major step of the reference integrator; X is advanced in place */
void plant_step(int v, double* X, const double* U)
{
    static const double B[6][6] = {
        {1.0 / 5.0, 0, 0, 0, 0, 0},
        {3.0 / 40.0, 9.0 / 40.0, 0, 0, 0, 0},
        {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 0, 0, 0},
        {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, 0, 0},
        {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0, 0},
        {35.0 / 384.0, 0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}};
    const double h = 0.01;
    double f[6][19], y[19], x[19], hB[6];
    memcpy(y, X, sizeof(y));
    plant_rhs(v, y, U, f[0]);
    for (int s = 0; s < 6; ++s) {
        for (int j = 0; j <= s; ++j) hB[j] = h * B[s][j];
        for (int i = 0; i < 19; ++i) {
            double acc = f[0][i] * hB[0];
            for (int j = 1; j <= s; ++j) acc += f[j][i] * hB[j];
            x[i] = y[i] + acc;
        }
        if (s < 5) plant_rhs(v, x, U, f[s + 1]);
    }
    memcpy(X, x, sizeof(x));
}

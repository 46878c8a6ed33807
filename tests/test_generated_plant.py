"""The GENERATED device sources (serl_b200/csrc/gen: fast mode, merged variants, pooled constants, table blob) are
checked on the CPU: compiled with gcc behind trivial macro definitions and compared with the oracle's exact restatement on
the reference-recorded right-hand-side vectors.  Also: the committed generated files are reproducible from the reference
binaries (container only)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAT = np.load(os.path.join(ROOT, 'tests', 'golden', 'plant_rhs_kat.npz'))
VARIANTS = ['h2000_v90', 'ice', 'cg', 'cg_for', 'h2000_v150', 'h10000_v90']
LIVE = [0, 1, 2, 3, 4, 5, 6, 7, 9, 12, 15, 16, 17, 18]

HARNESS = r'''
#include <math.h>
#include <stdbool.h>
typedef %(real)s real;
#define __restrict__ restrict
#define __device__
#define PLANT_FN static
#define PLANT_XARGS , const real* restrict plant_tab
#define PLANT_TAB(name) (plant_tab + PT_OFF_##name)
#define PLANT_CONSTS(n) static const real plant_k[n]
#define PLANT_K(i) plant_k[i]
#define PLANT_IC_TABLE static const double plant_ic_table[8][19]
#define PLANT_IC(v) static const double plant_ic_unused_##v[19]
#define PLANT_PV_TABLE static const real plant_pv[8][PLANT_NPV]
#define PLANT_PV(k) plant_pvrow[k]
#define PLANT_XI(i) (i)
#define PLANT_DIV(a, b) ((a) / (b))
#define PLANT_SQRT sqrt%(sfx)s
#define PLANT_FABS fabs%(sfx)s
#define PLANT_SIN sin%(sfx)s
#define PLANT_COS cos%(sfx)s
#define PLANT_SINCOS sincos%(sfx)s
#define PLANT_TAN tan%(sfx)s
#define PLANT_EXP exp%(sfx)s
#define PLANT_LOG10 log10%(sfx)s
#define PLANT_POW pow%(sfx)s
#define _GNU_SOURCE
#include "%(support)s"
#include "%(gen)s/plant_tables_blob.h"
#include "%(gen)s/plant_consts.h"
#include "%(gen)s/plant_ic.h"
#include "%(gen)s/plant_rhs_common.h"
#include "%(gen)s/plant_rhs_nav.h"
void dev_rhs(int variant, const double* Xd, const double* Ud, double* out) {
    real X[19], U[3], xdot[19], nav[19];
    for (int i = 0; i < 19; ++i) { X[i] = (real)Xd[i]; xdot[i] = 0; }
    for (int i = 0; i < 3; ++i) U[i] = (real)Ud[i];
    plant_rhs_common(X, U, xdot, plant_tables_blob, plant_pv[variant]);      /* one function for every variant */
    plant_rhs_nav(X, U, nav, plant_tables_blob);
    xdot[8] = nav[8]; xdot[10] = nav[10]; xdot[11] = nav[11];
    for (int i = 0; i < 19; ++i) out[i] = (double)xdot[i];
}
void dev_ic(int variant, double* X) { for (int i = 0; i < 19; ++i) X[i] = plant_ic_table[variant][i]; }
'''


@pytest.fixture(scope='module', params=['gen', 'gen_exact', 'gen_f32'])
def devlib(request, tmp_path_factory):
    d = tmp_path_factory.mktemp('devplant_' + request.param)
    src = d / 'h.c'
    src.write_text(HARNESS % {'support': os.path.join(ROOT, 'serl_b200', 'csrc', 'plant_support.h'),
                              'gen': os.path.join(ROOT, 'serl_b200', 'csrc', request.param),
                              'real': 'float' if request.param == 'gen_f32' else 'double', 'sfx': 'f' if request.param == 'gen_f32' else ''})
    so = d / 'h.so'
    subprocess.check_call(['gcc', '-O1', '-D_GNU_SOURCE', '-ffp-contract=off', '-fPIC', '-shared', '-o', str(so), str(src), '-lm'])
    lib = ctypes.CDLL(str(so))
    return request.param, lib


@pytest.mark.parametrize('variant', VARIANTS)
def test_generated_device_rhs_matches_reference_binary_vectors(devlib, variant):
    which, lib = devlib
    D = ctypes.c_double
    v = VARIANTS.index(variant)
    ic = (D * 19)()
    lib.dev_ic(v, ic)
    assert np.array_equal(np.array(ic[:]), KAT[variant + '_ic'])
    worst = 0.0
    for x, u, f in zip(KAT[variant + '_X'], KAT[variant + '_U'], KAT[variant + '_F']):
        xd = (D * 19)()
        lib.dev_rhs(v, (D * 19)(*x), (D * 3)(*u), xd)
        got = np.array(xd[:])
        idx = LIVE + [8, 10, 11]
        if which == 'gen_exact':
            assert np.array_equal(got[idx], f[idx])          # reference operation order: bit-exact
        err = np.abs(got[idx] - f[idx]) / np.maximum(np.abs(f[idx]), 1e-3 if which != 'gen_f32' else 1.0)
        worst = max(worst, err.max())
    # fast mode (reciprocal tables / constants, merged rows): 1e-11; single-precision right-hand side: float round-off
    assert worst < (5e-4 if which == 'gen_f32' else 1e-11), worst


@pytest.mark.skipif(not os.path.isdir('/root/reference/envs'), reason='needs the reference tree (build container only)')
def test_committed_generated_sources_are_reproducible(tmp_path):
    """tools/lift regenerates byte-identical device sources from the reference binaries."""
    code = ("import sys, os; sys.path.insert(0, %r); import gen_all as G; G.emit_set(%r, live=True)" %
            (os.path.join(ROOT, 'tools', 'lift'), str(tmp_path / 'gen')))
    subprocess.check_call([sys.executable, '-c', code], stdout=subprocess.DEVNULL)
    for f in sorted(os.listdir(tmp_path / 'gen')):
        a = open(tmp_path / 'gen' / f).read()
        b = open(os.path.join(ROOT, 'serl_b200', 'csrc', 'gen', f)).read()
        assert a == b, f

"""Exact (bit-for-bit, IEEE double) evaluator for a traced expression graph; used to check the trace
against the live reference binary before any code is generated."""
import ctypes, math
import symtrace as S

_libm = ctypes.CDLL('libm.so.6')
_libm.pow.restype = ctypes.c_double
_libm.pow.argtypes = [ctypes.c_double, ctypes.c_double]


def interval_t3(xs, u):
    """table3 index rule: i with xs[i] < u <= xs[i+1], clamped to [0, n-2]."""
    n = len(xs)
    i = 0
    while i < n and xs[i] < u:
        i += 1
    i -= 1
    return min(max(i, 0), n - 2)


def lerp_t3(v0, v1, u, xlo, xhi):
    if u == xhi:
        return v1
    return (v1 - v0) * (u - xlo) / (xhi - xlo) + v0


def table3(P1, P2, P3, P4, u0, u1, u2):
    i1, i2, i3 = interval_t3(P1, u0), interval_t3(P2, u1), interval_t3(P3, u2)
    n1, n2 = len(P1), len(P2)

    def t2(k):
        w = []
        for j in (i2, i2 + 1):
            v0 = P4[(k * n1 + i1) * n2 + j]
            v1 = P4[(k * n1 + i1 + 1) * n2 + j]
            w.append(lerp_t3(v0, v1, u0, P1[i1], P1[i1 + 1]))
        return lerp_t3(w[0], w[1], u1, P2[i2], P2[i2 + 1])
    a = t2(i3)
    b = t2(i3 + 1)
    return lerp_t3(a, b, u2, P3[i3], P3[i3 + 1])


class Evaluator:
    def __init__(self, tracer, powsnf=None):
        self.tr = tracer
        self.powsnf = powsnf

    def run(self, outputs, X, U):
        tabs = self.tr.tables
        vals = {}
        li = S.Tracer.lookup_index

        def ev(n):
            if not S.is_sym(n):
                return S.fval(n)
            stack = [n]
            while stack:
                m = stack[-1]
                if m.id in vals:
                    stack.pop()
                    continue
                pend = [a for a in m.args if S.is_sym(a) and a.id not in vals]
                if pend:
                    stack.extend(pend)
                    continue
                stack.pop()
                a = [vals[x.id] if S.is_sym(x) else x for x in m.args]
                op = m.op
                if op == 'const':
                    v = S.fval(a[0])
                elif op == 'X':
                    v = X[a[0]]
                elif op == 'U':
                    v = U[a[0]]
                elif op == 'add':
                    v = a[0] + a[1]
                elif op == 'sub':
                    v = a[0] - a[1]
                elif op == 'mul':
                    v = a[0] * a[1]
                elif op == 'div':
                    v = a[0] / a[1] if a[1] != 0 else math.copysign(math.inf, a[0]) * math.copysign(1, a[1])
                elif op == 'sqrt':
                    v = math.sqrt(a[0]) if a[0] >= 0 else math.nan
                elif op == 'max':
                    v = a[0] if a[0] > a[1] else a[1]
                elif op == 'min':
                    v = a[0] if a[0] < a[1] else a[1]
                elif op == 'neg':
                    v = -a[0]
                elif op == 'abs':
                    v = abs(a[0])
                elif op in ('cmp', 'cmpmask'):
                    p, x, y = a
                    v = {'gt': x > y, 'ge': x >= y, 'lt': x < y, 'le': x <= y, 'eq': x == y, 'ne': x != y}[p]
                elif op == 'select':
                    v = a[1] if a[0] else a[2]
                elif op == 'mand':
                    v = a[1] if a[0] else 0.0
                elif op == 'mandn':
                    v = 0.0 if a[0] else a[1]
                elif op in ('sin', 'cos', 'tan', 'exp', 'log10', 'atan', 'asin', 'acos'):
                    v = getattr(math, op)(a[0])
                elif op == 'pow':
                    v = _libm.pow(a[0], a[1])
                elif op == 'powsnf':
                    v = self.powsnf(a[0], a[1])
                elif op == 'lookup1':
                    xs, ys = tabs[a[0]], tabs[a[1]]
                    i = li(xs, a[2])
                    v = (ys[i + 1] - ys[i]) / (xs[i + 1] - xs[i]) * (a[2] - xs[i]) + ys[i]
                elif op == 'lookup2':
                    xs, ys, zs = tabs[a[0]], tabs[a[1]], tabs[a[2]]
                    nx = len(xs)
                    ix, iy = li(xs, a[3]), li(ys, a[4])
                    dx = xs[ix + 1] - xs[ix]
                    ux = a[3] - xs[ix]
                    lo = (zs[ix + 1 + nx * iy] - zs[ix + nx * iy]) / dx * ux + zs[ix + nx * iy]
                    hi = (zs[ix + 1 + nx * (iy + 1)] - zs[ix + nx * (iy + 1)]) / dx * ux + zs[ix + nx * (iy + 1)]
                    v = (hi - lo) / (ys[iy + 1] - ys[iy]) * (a[4] - ys[iy]) + lo
                elif op == 'table3':
                    v = table3(tabs[a[0]], tabs[a[1]], tabs[a[2]], tabs[a[3]], a[4], a[5], a[6])
                else:
                    raise NotImplementedError(op)
                vals[m.id] = v
            return vals[n.id]
        return [ev(o) for o in outputs]

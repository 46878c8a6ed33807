"""Trace xdot = f(X, cmd) out of a plant binary. See symtrace.py."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(__file__))
sys.setrecursionlimit(100000)
import symtrace as S

SIMTIMESTEP_OFF = 0xba48     # rtM_.Timing.simTimeStep (step @0x6060 reads rtM_+0xba48)
DERIVS_OFF = 46416           # rtM_.derivs


# model time inside rtM_ (found by stepping the cg_timed build and scanning rtM_: two doubles that read k*0.01 after k steps,
# two uint32 tick counters); poked to trace the right-hand side AFTER a time-triggered switch (cg_timed: t >= 20 s)
TIME_DOUBLES, TIME_TICKS = (64, 47808), (47640, 47656)


def trace(so, nsym_u=3, verbose=False, assume=None, concrete=(), t_poke=None):
    img = S.Image(so, '/tmp/lift')
    img.lib.initialize()
    if t_poke is not None:
        base = img.addr('rtM_')
        for off in TIME_DOUBLES:
            ctypes.c_double.from_address(base + off).value = float(t_poke)
        for off in TIME_TICKS:
            ctypes.c_uint32.from_address(base + off).value = int(round(t_poke / 0.01))
    st = S.State(img)
    stack = ctypes.create_string_buffer(1 << 20)
    st.g[4] = (ctypes.addressof(stack) + (1 << 20) - 65536) & ~0xf
    cmd = ctypes.create_string_buffer(80)
    out = ctypes.create_string_buffer(96)
    ca, oa = ctypes.addressof(cmd), ctypes.addressof(out)
    rtX = img.addr('rtX')
    for i in range(19):
        if i not in concrete:
            st.wr64(rtX + 8 * i, S.G.mk('X', i))
    for i in range(nsym_u):
        st.wr64(ca + 8 * i, S.G.mk('U', i))
    rtM = img.addr('rtM_')
    st.wr_n(rtM + SIMTIMESTEP_OFF, 0, 4)
    tr = S.Tracer(img, assume=assume, verbose=verbose)
    tr._keep = (stack, cmd, out)
    st.g[7], st.g[6] = ca, oa
    dummy = S.Insn(); dummy.addr = 0
    tr.call_function(st, img.addr('step'), dummy)
    st.g[7], st.g[6] = ca, oa
    tr.call_function(st, img.addr('citation_to_python_derivatives'), dummy)
    derivs = st.rd64(rtM + DERIVS_OFF)
    xdot = [st.rd64(derivs + 8 * i) for i in range(19)]
    return img, tr, st, xdot


if __name__ == '__main__':
    so = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/envs/h2000_v90/_citation.cpython-38-x86_64-linux-gnu.so'
    img, tr, st, xdot = trace(so, verbose=True)
    print('executed', tr.nexec, 'instructions; graph nodes', len(S.G.nodes))
    for i, x in enumerate(xdot):
        print(i, x if S.is_sym(x) else S.fval(x))

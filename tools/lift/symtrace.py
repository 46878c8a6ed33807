"""Symbolic tracer for the PH-LAB plant binary (test/oracle tooling, never shipped on the product path).

The reference ships the aircraft dynamics only as machine code
(/root/reference/envs/<variant>/_citation.cpython-38-x86_64-linux-gnu.so).  This tool
loads that shared object in-process with ctypes, then *symbolically executes* the x86-64 code of
`step` (minor time step = one Outputs pass) followed by `citation_to_python_derivatives`:

  * integer registers, pointers and every value derived from model constants stay concrete (they are
    read straight out of the live process image, so SimStruct pointer chasing, loop counters, table
    addresses, S-function parameters ... all resolve by themselves and loops unroll);
  * doubles that depend on the continuous state X[19] or the command cmd[10] are symbolic nodes;
  * data-dependent branches (saturations, switches, sqrt domain checks) are if-converted by forking
    at the branch and merging at the immediate post-dominator with select() nodes;
  * table lookups (rt_Lookup, rt_Lookup2D_Normal), rt_powd_snf and libm calls become intrinsic nodes.

The result is a straight-line SSA expression graph  xdot = f(X, cmd)  that `codegen.py` optimises and
prints as C / CUDA.  Nothing here is imported by the product; it needs /root/reference and runs only
in the build container.
"""
import ctypes
import math
import os
import re
import shutil
import struct
import subprocess

M64 = (1 << 64) - 1
SIGN = 1 << 63


def fval(bits):
    return struct.unpack('<d', struct.pack('<Q', bits & M64))[0]


def fbits(x):
    return struct.unpack('<Q', struct.pack('<d', x))[0]


# ----------------------------------------------------------------------------------------------
# expression graph
# ----------------------------------------------------------------------------------------------
class Node:
    __slots__ = ('op', 'args', 'id')

    def __repr__(self):
        return 'n%d:%s' % (self.id, self.op)


class Graph:
    def __init__(self):
        self.nodes = []
        self.memo = {}

    def mk(self, op, *args):
        key = (op,) + tuple(a.id if isinstance(a, Node) else ('c', a) for a in args)
        n = self.memo.get(key)
        if n is None:
            n = Node()
            n.op = op
            n.args = args
            n.id = len(self.nodes)
            self.nodes.append(n)
            self.memo[key] = n
        return n

    def const(self, bits):
        return self.mk('const', bits & M64)

    def lift(self, v):
        return v if isinstance(v, Node) else self.const(v)


G = Graph()


def is_sym(v):
    return isinstance(v, Node)


# exact concrete evaluation of the fp ops (python floats are IEEE doubles; libm is glibc's, the same
# one the reference binary links against)
def _maxsd(a, b):   # maxsd dst=a, src=b : a > b ? a : b
    return a if a > b else b


def _minsd(a, b):
    return a if a < b else b


CONC2 = {
    'add': lambda a, b: a + b, 'sub': lambda a, b: a - b, 'mul': lambda a, b: a * b,
    'div': lambda a, b: (a / b) if b != 0.0 else (math.copysign(math.inf, a) * math.copysign(1.0, b) if a != 0 and a == a else math.nan),
    'max': _maxsd, 'min': _minsd,
}


SIMPLIFY_ZERO = True    # 0*x -> 0, x+0 -> x, x-0 -> x  (exact up to the sign of a zero result; checked numerically)


def _is_zero(v):
    return (not is_sym(v)) and (v & (SIGN - 1)) == 0


def fop2(op, a, b):
    if is_sym(a) or is_sym(b):
        if SIMPLIFY_ZERO:
            if op == 'mul' and (_is_zero(a) or _is_zero(b)):
                return 0
            if op == 'add':
                if _is_zero(a):
                    return b
                if _is_zero(b):
                    return a
            if op == 'sub' and _is_zero(b):
                return a
            if op == 'div' and _is_zero(a):
                return 0
        return G.mk(op, G.lift(a), G.lift(b))
    return fbits(CONC2[op](fval(a), fval(b)))


def fsqrt(a):
    if is_sym(a):
        return G.mk('sqrt', a)
    x = fval(a)
    return fbits(math.sqrt(x)) if x >= 0 else fbits(math.nan)


def bitop(op, a, b):
    """andpd / andnpd / orpd / xorpd on one 64-bit lane."""
    if not is_sym(a) and not is_sym(b):
        if op == 'and':
            return a & b
        if op == 'andn':
            return (~a & M64) & b
        if op == 'or':
            return a | b
        if op == 'xor':
            return a ^ b
    # symbolic idioms
    if op == 'xor':
        if a is b:
            return 0
        for x, y in ((a, b), (b, a)):
            if not is_sym(y) and y == SIGN:
                return G.mk('neg', x)
            if not is_sym(y) and y == 0:
                return x
    if op == 'and':
        for x, y in ((a, b), (b, a)):
            if not is_sym(y) and y == (SIGN - 1):
                return G.mk('abs', x)
            if not is_sym(y) and y == M64:
                return x
            if not is_sym(y) and y == 0:
                return 0
            if is_sym(y) and y.op == 'cmpmask':
                return G.mk('mand', y, G.lift(x))       # mask ? x : 0
    if op == 'andn':   # ~a & b
        if is_sym(a) and a.op == 'cmpmask':
            return G.mk('mandn', a, G.lift(b))          # mask ? 0 : b
        if not is_sym(a) and a == SIGN:
            return G.mk('abs', b)
        if not is_sym(a) and a == 0:
            return b
    if op == 'or':
        for x, y in ((a, b), (b, a)):
            if is_sym(x) and is_sym(y) and x.op == 'mand' and y.op == 'mandn' and x.args[0] is y.args[0]:
                return select(x.args[0], x.args[1], y.args[1])
            if not is_sym(y) and y == 0:
                return x
        # sign transfer idiom: (abs-part) | (sign-part) is not expected here
    raise NotImplementedError('bitop %s %r %r' % (op, a, b))


def select(c, a, b):
    a = G.lift(a)
    b = G.lift(b)
    if a is b:
        return a
    return G.mk('select', c, a, b)


def cond_node(pred, a, b):
    """boolean node for an fp comparison a <pred> b (ordered)."""
    return G.mk('cmp', pred, G.lift(a), G.lift(b))


NEG = {'gt': 'le', 'le': 'gt', 'ge': 'lt', 'lt': 'ge', 'eq': 'ne', 'ne': 'eq'}


# ----------------------------------------------------------------------------------------------
# binary image
# ----------------------------------------------------------------------------------------------
class Insn:
    __slots__ = ('addr', 'op', 'ops', 'next', 'text', 'target', 'tname')


def split_ops(s):
    out, depth, cur = [], 0, ''
    for ch in s:
        if ch == '(':
            depth += 1
        elif ch == ')':
            depth -= 1
        if ch == ',' and depth == 0:
            out.append(cur.strip())
            cur = ''
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


class Image:
    def __init__(self, so_path, workdir):
        os.makedirs(workdir, exist_ok=True)
        Image._n = getattr(Image, '_n', 0) + 1      # a fresh copy per load: never overwrite a mapped file
        self.path = os.path.join(workdir, 'plant_%d_%d_' % (os.getpid(), Image._n) + re.sub(r'\W', '_', os.path.dirname(so_path)[-20:]) + '.so')
        shutil.copy(so_path, self.path)
        os.chmod(self.path, 0o755)
        self.lib = ctypes.CDLL(self.path)
        self.base = None
        for line in open('/proc/self/maps'):
            if self.path in line:
                lo = int(line.split('-')[0], 16)
                self.base = lo if self.base is None else min(self.base, lo)
        dis = subprocess.run(['objdump', '-d', '--no-show-raw-insn', self.path], capture_output=True, text=True).stdout
        self.insns = {}
        self.funcs = {}       # start -> (name, [addrs])
        cur = None
        prev = None
        for l in dis.splitlines():
            m = re.match(r'^([0-9a-f]+) <(.+)>:', l)
            if m:
                cur = int(m.group(1), 16)
                self.funcs[cur] = (m.group(2), [])
                prev = None
                continue
            m = re.match(r'^\s+([0-9a-f]+):\t(.*)$', l)
            if not m or cur is None:
                continue
            a = int(m.group(1), 16)
            text = m.group(2).split('#')[0].rstrip()
            parts = text.split(None, 1)
            op = parts[0]
            rest = parts[1] if len(parts) > 1 else ''
            while op in ('cs', 'data16', 'notrack', 'bnd', 'rep', 'repz') and rest:
                parts = rest.split(None, 1)
                op = parts[0]
                rest = parts[1] if len(parts) > 1 else ''
            ins = Insn()
            ins.addr = a
            ins.op = op
            ins.text = text
            ins.target = None
            ins.tname = None
            ins.next = None
            if op.startswith('j') or op == 'call':
                mm = re.match(r'^([0-9a-f]+) <(.+)>$', rest.strip())
                if mm:
                    ins.target = int(mm.group(1), 16)
                    ins.tname = mm.group(2)
                    ins.ops = []
                else:
                    ins.ops = [rest.strip()]
            else:
                ins.ops = split_ops(rest)
            if prev is not None:
                prev.next = a
            prev = ins
            self.insns[a] = ins
            self.funcs[cur][1].append(a)
        syms = subprocess.run(['nm', self.path], capture_output=True, text=True).stdout
        self.sym = {}
        for l in syms.splitlines():
            p = l.split()
            if len(p) == 3:
                self.sym.setdefault(p[2], int(p[0], 16))
        self.func_of = {}
        for s, (name, addrs) in self.funcs.items():
            for a in addrs:
                self.func_of[a] = s
        self._ipdom = {}

    def addr(self, name):
        return self.base + self.sym[name]

    def read(self, a, n):
        return ctypes.string_at(a, n)

    # ---- static CFG / post-dominators (per function) ----
    def ipdom_of(self, branch_addr):
        f = self.func_of[branch_addr]
        if f not in self._ipdom:
            self._ipdom[f] = self._build_ipdom(f)
        blk_of, ipd = self._ipdom[f]
        return ipd[blk_of[branch_addr]]

    def _build_ipdom(self, f):
        name, addrs = self.funcs[f]
        aset = set(addrs)
        leaders = {addrs[0]}
        for a in addrs:
            i = self.insns[a]
            if i.op.startswith('j'):
                if i.target is not None and i.target in aset:
                    leaders.add(i.target)
                if i.next is not None:
                    leaders.add(i.next)
            elif i.op in ('ret',):
                if i.next is not None:
                    leaders.add(i.next)
        blocks = {}
        blk_of = {}
        cur = None
        for a in addrs:
            if a in leaders:
                cur = a
                blocks[cur] = []
            blocks[cur].append(a)
            blk_of[a] = cur
        EXIT = -1
        succ = {}
        for b, ins_list in blocks.items():
            last = self.insns[ins_list[-1]]
            s = []
            if last.op == 'ret':
                s = [EXIT]
            elif last.op == 'jmp':
                s = [last.target if (last.target in aset) else EXIT]
            elif last.op.startswith('j'):
                s = [last.target if last.target in aset else EXIT]
                if last.next in aset:
                    s.append(last.next)
            elif last.op == 'call' and last.tname and 'stack_chk_fail' in last.tname:
                s = [EXIT]
            else:
                s = [last.next] if (last.next in aset) else [EXIT]
            succ[b] = s
        nodes = list(blocks.keys()) + [EXIT]
        full = set(nodes)
        pdom = {n: set(full) for n in nodes}
        pdom[EXIT] = {EXIT}
        changed = True
        order = list(reversed(list(blocks.keys())))
        while changed:
            changed = False
            for n in order:
                new = None
                for s in succ[n]:
                    new = set(pdom[s]) if new is None else (new & pdom[s])
                new = (new or set()) | {n}
                if new != pdom[n]:
                    pdom[n] = new
                    changed = True
        ipd = {}
        for n in blocks:
            cands = pdom[n] - {n}
            best = None
            for c in cands:
                if len(pdom[c]) == len(cands):
                    best = c
            ipd[n] = best if best is not None else EXIT
        return blk_of, ipd


# ----------------------------------------------------------------------------------------------
# machine state
# ----------------------------------------------------------------------------------------------
R64 = ['rax', 'rcx', 'rdx', 'rbx', 'rsp', 'rbp', 'rsi', 'rdi', 'r8', 'r9', 'r10', 'r11', 'r12', 'r13', 'r14', 'r15']
REGMAP = {}
for i, r in enumerate(R64):
    REGMAP[r] = (i, 8)
for i, r in enumerate(['eax', 'ecx', 'edx', 'ebx', 'esp', 'ebp', 'esi', 'edi']):
    REGMAP[r] = (i, 4)
for i, r in enumerate(['ax', 'cx', 'dx', 'bx', 'sp', 'bp', 'si', 'di']):
    REGMAP[r] = (i, 2)
for i, r in enumerate(['al', 'cl', 'dl', 'bl', 'spl', 'bpl', 'sil', 'dil']):
    REGMAP[r] = (i, 1)
for i in range(8, 16):
    REGMAP['r%dd' % i] = (i, 4)
    REGMAP['r%dw' % i] = (i, 2)
    REGMAP['r%db' % i] = (i, 1)


class Mem:
    def __init__(self, parent=None):
        self.d = {}
        self.parent = parent

    def get(self, a):
        m = self
        while m is not None:
            v = m.d.get(a)
            if v is not None:
                return v
            m = m.parent
        return None


class Poison:
    def __repr__(self):
        return 'POISON'


POISON = Poison()


class State:
    def __init__(self, img):
        self.img = img
        self.g = [0] * 16
        self.x = [[0, 0] for _ in range(16)]
        self.mem = Mem()
        self.flags = None

    def fork(self):
        s = State.__new__(State)
        s.img = self.img
        s.g = list(self.g)
        s.x = [list(v) for v in self.x]
        s.mem = Mem(self.mem)
        s.flags = self.flags
        return s

    # memory, 8-byte slots
    def rd64(self, a):
        if a & 7:
            lo = self.rd64(a & ~7)
            hi = self.rd64((a & ~7) + 8)
            if is_sym(lo) or is_sym(hi):
                raise NotImplementedError('unaligned symbolic read %x' % a)
            sh = (a & 7) * 8
            return ((lo >> sh) | (hi << (64 - sh))) & M64
        v = self.mem.get(a)
        if v is None:
            v = int.from_bytes(self.img.read(a, 8), 'little')
        if v is POISON:
            raise RuntimeError('read of poisoned memory %x' % a)
        return v

    def wr64(self, a, v):
        if a & 7:
            if is_sym(v):
                raise NotImplementedError('unaligned symbolic write')
            for k in range(8):
                self.wr_n(a + k, (v >> (8 * k)) & 0xff, 1)
            return
        self.mem.d[a] = v

    def rd_n(self, a, n):
        if n == 8:
            return self.rd64(a)
        base = a & ~7
        sh = (a - base) * 8
        if sh + n * 8 > 64:
            raise NotImplementedError('straddling read')
        v = self.rd64(base)
        if is_sym(v):
            raise NotImplementedError('partial read of symbolic slot %x' % a)
        return (v >> sh) & ((1 << (8 * n)) - 1)

    def wr_n(self, a, v, n):
        if n == 8:
            return self.wr64(a, v)
        base = a & ~7
        sh = (a - base) * 8
        if sh + n * 8 > 64:
            raise NotImplementedError('straddling write')
        old = self.rd64(base)
        if is_sym(old):
            old = 0
        mask = ((1 << (8 * n)) - 1) << sh
        self.mem.d[base] = (old & ~mask & M64) | ((v << sh) & mask)


class TraceError(Exception):
    pass


# ----------------------------------------------------------------------------------------------
# interpreter
# ----------------------------------------------------------------------------------------------
class Tracer:
    def __init__(self, img, assume=None, verbose=False):
        self.img = img
        self.verbose = verbose
        self.nexec = 0
        self.tables = {}       # (addr, n) -> name ; concrete arrays referenced by intrinsic nodes
        self.assume = assume or (lambda tracer, ins, cond: None)
        self.branch_log = []
        self.callstack = []

    # ---- operand helpers ----
    def ea(self, st, s):
        m = re.match(r'^(%fs:)?(-?0x[0-9a-f]+|-?\d+)?(?:\((%\w+)?(?:,(%\w+)(?:,(\d))?)?\))?$', s)
        if not m:
            raise TraceError('bad mem operand ' + s)
        if m.group(1):
            return ('fs', int(m.group(2), 16))
        disp = int(m.group(2), 16) if m.group(2) else 0
        a = disp
        if m.group(3):
            base = m.group(3)[1:]
            if base == 'rip':
                raise TraceError('rip handled by caller')
            a += self.greg(st, base)
        if m.group(4):
            a += self.greg(st, m.group(4)[1:]) * (int(m.group(5)) if m.group(5) else 1)
        return a & M64

    def mem_addr(self, st, ins, s):
        if '(%rip)' in s:
            disp = int(s.split('(')[0], 16)
            return (self.img.base + ins.next + disp) & M64
        a = self.ea(st, s)
        return a

    def greg(self, st, name):
        i, n = REGMAP[name]
        v = st.g[i]
        if v is POISON:
            raise TraceError('use of poisoned register ' + name)
        if is_sym(v):
            if n == 8:
                return v
            raise TraceError('partial use of symbolic gpr ' + name)
        return v & ((1 << (8 * n)) - 1)

    def sreg(self, st, name, v):
        i, n = REGMAP[name]
        if is_sym(v):
            if n != 8:
                raise TraceError('symbolic into partial gpr')
            st.g[i] = v
            return
        if n == 8:
            st.g[i] = v & M64
        elif n == 4:
            st.g[i] = v & 0xffffffff
        else:
            old = st.g[i]
            if is_sym(old) or old is POISON:
                old = 0
            mask = (1 << (8 * n)) - 1
            st.g[i] = (old & ~mask) | (v & mask)

    @staticmethod
    def is_xmm(s):
        return s.startswith('%xmm')

    @staticmethod
    def is_reg(s):
        return s.startswith('%') and not s.startswith('%fs')

    def opsize(self, ins, *ops):
        op = ins.op
        for o in ops:
            if self.is_reg(o) and not self.is_xmm(o):
                return REGMAP[o[1:]][1]
        if op.endswith('q'):
            return 8
        if op.endswith('l'):
            return 4
        if op.endswith('w'):
            return 2
        if op.endswith('b'):
            return 1
        raise TraceError('cannot size ' + ins.text)

    def rd_int(self, st, ins, s, n):
        if s.startswith('$'):
            return int(s[1:], 16) & ((1 << (8 * n)) - 1)
        if self.is_reg(s):
            return self.greg(st, s[1:])
        a = self.mem_addr(st, ins, s)
        if isinstance(a, tuple):
            return 0     # %fs:0x28 stack canary
        return st.rd_n(a, n)

    def wr_int(self, st, ins, s, v, n):
        if self.is_reg(s):
            self.sreg(st, s[1:], v)
        else:
            a = self.mem_addr(st, ins, s)
            st.wr_n(a, v, n)

    def rd_f64(self, st, ins, s):
        if self.is_xmm(s):
            return st.x[int(s[4:])][0]
        return st.rd64(self.mem_addr(st, ins, s))

    # ---- flags ----
    def cond(self, st, cc, ins):
        """returns True/False for a concrete condition, or a cmp Node for a symbolic one."""
        f = st.flags
        if f is None:
            raise TraceError('flags undefined at %x' % ins.addr)
        if f[0] == 'fp':
            _, a, b = f
            if is_sym(a) or is_sym(b):
                pred = {'a': 'gt', 'ae': 'ge', 'b': 'lt', 'be': 'le', 'e': 'eq', 'ne': 'ne', 'nb': 'ge', 'na': 'le', 'nbe': 'gt', 'nae': 'lt'}.get(cc)
                if cc == 'p':
                    return False       # assumption: no NaNs in the state
                if cc == 'np':
                    return True
                if pred is None:
                    raise TraceError('fp cond ' + cc)
                return cond_node(pred, a, b)
            x, y = fval(a), fval(b)
            unordered = (x != x) or (y != y)
            CF = unordered or x < y
            ZF = unordered or x == y
            PF = unordered
            return {'a': (not CF and not ZF), 'ae': not CF, 'b': CF, 'be': CF or ZF, 'e': ZF, 'ne': not ZF, 'p': PF, 'np': not PF}[cc]
        _, kind, res, a, b, n = f
        bits = 8 * n
        mask = (1 << bits) - 1
        sign = 1 << (bits - 1)
        ZF = (res & mask) == 0
        SF = bool(res & sign)
        if kind == 'sub':
            CF = (a & mask) < (b & mask)
            sa = (a & mask) - ((a & sign) << 1)
            sb = (b & mask) - ((b & sign) << 1)
            r = sa - sb
            OF = not (-(sign) <= r < sign)
        elif kind == 'add':
            CF = ((a & mask) + (b & mask)) > mask
            sa = (a & mask) - ((a & sign) << 1)
            sb = (b & mask) - ((b & sign) << 1)
            r = sa + sb
            OF = not (-(sign) <= r < sign)
        else:
            CF = False
            OF = False
        tbl = {'e': ZF, 'ne': not ZF, 'a': (not CF and not ZF), 'ae': not CF, 'b': CF, 'be': CF or ZF,
               's': SF, 'ns': not SF, 'l': SF != OF, 'ge': SF == OF, 'le': ZF or (SF != OF), 'g': (not ZF) and SF == OF}
        return tbl[cc]

    # ---- intrinsics ----
    def table_ref(self, addr, n):
        key = (addr, n)
        if key not in self.tables:
            vals = struct.unpack('<%dd' % n, self.img.read(addr, 8 * n))
            self.tables[key] = vals
        return key

    @staticmethod
    def lookup_index(xs, u):
        """rt_GetLookupIndex @0xf470 semantics (exact, incl. tie handling)."""
        n = len(xs)
        if xs[0] >= u:
            return 0
        if not (u < xs[n - 1]):
            return n - 2
        bottom, top = 0, n - 1
        while True:
            s = bottom + top
            idx = (s + (1 if s < 0 else 0)) >> 1
            if u >= 0.0:
                if xs[idx] < u:
                    bottom = idx + 1
                    if u > xs[bottom]:
                        continue
                    return idx
                top = idx - 1
            else:
                if xs[idx] <= u:
                    bottom = idx + 1
                    if u < xs[bottom]:
                        return idx
                    continue
                top = idx - 1

    def intrinsic(self, st, name, ins):
        x0 = st.x[0][0]
        x1 = st.x[1][0]
        if name in ('sin', 'cos', 'tan', 'exp', 'log10', 'log', 'sqrt', 'atan', 'asin', 'acos'):
            if is_sym(x0):
                st.x[0] = [G.mk(name, x0), 0]
            else:
                st.x[0] = [fbits(getattr(math, name)(fval(x0))), 0]
            return True
        if name == 'sincos':
            ps, pc = st.g[7], st.g[6]
            if is_sym(x0):
                st.wr64(ps, G.mk('sin', x0))
                st.wr64(pc, G.mk('cos', x0))
            else:
                st.wr64(ps, fbits(math.sin(fval(x0))))
                st.wr64(pc, fbits(math.cos(fval(x0))))
            return True
        if name in ('pow', 'rt_powd_snf', 'atan2'):
            nm = {'pow': 'pow', 'rt_powd_snf': 'powsnf', 'atan2': 'atan2'}[name]
            if is_sym(x0) or is_sym(x1):
                st.x[0] = [G.mk(nm, G.lift(x0), G.lift(x1)), 0]
            else:
                if nm == 'atan2':
                    st.x[0] = [fbits(math.atan2(fval(x0), fval(x1))), 0]
                else:
                    # concrete pow: evaluate through the real libm / real rt_powd_snf
                    fn = getattr(self.img.lib, name) if name == 'rt_powd_snf' else ctypes.CDLL('libm.so.6').pow
                    fn.restype = ctypes.c_double
                    fn.argtypes = [ctypes.c_double, ctypes.c_double]
                    st.x[0] = [fbits(fn(fval(x0), fval(x1))), 0]
            return True
        if name == 'rt_Lookup':       # (x*, n, u, y*)
            xa, n, ya = st.g[7], st.g[6] & 0xffffffff, st.g[2]
            kx = self.table_ref(xa, n)
            ky = self.table_ref(ya, n)
            st.x[0] = [self.lookup1(kx, ky, x0), 0]
            return True
        if name == 'rt_Lookup2D_Normal':   # (xVals, numX, yVals, numY, z, x, y)
            xa, nx, ya, ny, za = st.g[7], st.g[6] & 0xffffffff, st.g[2], st.g[1] & 0xffffffff, st.g[8]
            kx = self.table_ref(xa, nx)
            ky = self.table_ref(ya, ny)
            kz = self.table_ref(za, nx * ny)
            st.x[0] = [self.lookup2(kx, ky, kz, x0, x1), 0]
            return True
        if name == 'rt_GetLookupIndex':
            xa, n = st.g[7], st.g[6] & 0xffffffff
            if is_sym(x0):
                raise TraceError('symbolic rt_GetLookupIndex outside lookup')
            xs = self.tables[self.table_ref(xa, n)]
            st.g[0] = self.lookup_index(xs, fval(x0))
            return True
        return False

    def table3(self, st):
        """table3 S-function mdlOutputs (3-D table, incremental index search with cached indices; the
        search is stateless up to exact ties): y = P4 interpolated at (u0,u1,u2) over breakpoints P1,P2,P3."""
        S = st.g[7]
        uptrs = st.rd64(st.rd64(S + 0xe8) + 0x10)
        u = [st.rd64(st.rd64(uptrs + 8 * k)) for k in range(3)]
        prm = st.rd64(S + 0x100)
        keys = []
        for k in range(4):
            p = st.rd64(prm + 8 * k)
            cnt = int(fval(st.rd64(p)) * fval(st.rd64(p + 8)))
            keys.append(self.table_ref(p + 16, cnt))
        y = st.rd64(st.rd64(S + 0xf0) + 0x10)
        st.wr64(y, G.mk('table3', keys[0], keys[1], keys[2], keys[3], G.lift(u[0]), G.lift(u[1]), G.lift(u[2])))

    def lookup1(self, kx, ky, u):
        if is_sym(u):
            return G.mk('lookup1', kx, ky, u)
        xs, ys = self.tables[kx], self.tables[ky]
        i = self.lookup_index(xs, fval(u))
        return fbits((ys[i + 1] - ys[i]) / (xs[i + 1] - xs[i]) * (fval(u) - xs[i]) + ys[i])

    def lookup2(self, kx, ky, kz, x, y):
        if is_sym(x) or is_sym(y):
            return G.mk('lookup2', kx, ky, kz, G.lift(x), G.lift(y))
        xs, ys, zs = self.tables[kx], self.tables[ky], self.tables[kz]
        nx = len(xs)
        xv, yv = fval(x), fval(y)
        ix = self.lookup_index(xs, xv)
        iy = self.lookup_index(ys, yv)
        dx = xs[ix + 1] - xs[ix]
        ux = xv - xs[ix]
        a = (zs[ix + 1 + nx * iy] - zs[ix + nx * iy]) / dx * ux + zs[ix + nx * iy]
        b = (zs[ix + 1 + nx * (iy + 1)] - zs[ix + nx * (iy + 1)]) / dx * ux + zs[ix + nx * (iy + 1)]
        return fbits((b - a) / (ys[iy + 1] - ys[iy]) * (yv - ys[iy]) + a)

    # ---- execution ----
    def call_function(self, st, target_abs, ins):
        rel = target_abs - self.img.base
        if rel not in self.img.funcs:
            raise TraceError('call to unknown address %x from %x' % (rel, ins.addr))
        name = self.img.funcs[rel][0]
        base_name = name.split('@')[0]
        if name == 'mdlOutputs' and any(self.img.insns[a].tname == 'Table2' for a in self.img.funcs[rel][1]):
            return self.table3(st)
        if self.intrinsic(st, base_name, ins):
            return
        if name.endswith('@plt'):
            raise TraceError('unhandled PLT call ' + name)
        # real call: push return address, run to ret
        st.g[4] = (st.g[4] - 8) & M64
        st.wr64(st.g[4], 0xdeadbeef)
        self.callstack.append(name)
        r = self.run(st, rel, None)
        self.callstack.pop()
        if r != 'ret':
            raise TraceError('function did not return')

    def run(self, st, pc, stop):
        img = self.img
        while True:
            if pc == stop:
                return 'stop'
            ins = img.insns.get(pc)
            if ins is None:
                raise TraceError('no instruction at %x' % pc)
            self.nexec += 1
            if self.nexec > 5_000_000:
                raise TraceError('runaway')
            op = ins.op
            nxt = ins.next
            # ---------------- control flow ----------------
            if op == 'ret':
                st.g[4] = (st.g[4] + 8) & M64
                return 'ret'
            if op == 'call':
                if ins.target is not None:
                    if ins.tname and 'stack_chk_fail' in ins.tname:
                        raise TraceError('stack_chk_fail reached')
                    self.call_function(st, img.base + ins.target, ins)
                else:
                    o = ins.ops[0]
                    assert o.startswith('*')
                    tgt = self.rd_int(st, ins, o[1:], 8)
                    self.call_function(st, tgt, ins)
                pc = nxt
                continue
            if op == 'jmp':
                if ins.target is None:
                    raise TraceError('indirect jmp at %x' % pc)
                pc = ins.target
                continue
            if op.startswith('j'):
                c = self.cond(st, op[1:], ins)
                if c is True or c is False:
                    pc = ins.target if c else nxt
                    continue
                forced = self.assume(self, ins, c)
                if forced is not None:
                    self.branch_log.append((pc, c, forced))
                    pc = ins.target if forced else nxt
                    continue
                join = img.ipdom_of(pc)
                s1 = st.fork()
                s2 = st.fork()
                jstop = None if join == -1 else join
                if self.verbose:
                    print('  ' * len(self.callstack) + 'fork @%x join %s' % (pc, hex(join) if join != -1 else 'EXIT'))
                r1 = self.run(s1, ins.target, jstop)
                r2 = self.run(s2, nxt, jstop)
                if r1 != r2:
                    raise TraceError('fork paths ended differently at %x: %s %s' % (pc, r1, r2))
                self.merge(st, c, s1, s2)
                if r1 == 'ret':
                    return 'ret'
                pc = join
                continue
            # ---------------- everything else ----------------
            self.step_insn(st, ins)
            pc = nxt

    def merge(self, st, c, s1, s2):
        for i in range(16):
            a, b = s1.g[i], s2.g[i]
            if a is b or (not is_sym(a) and not is_sym(b) and a is not POISON and b is not POISON and a == b):
                st.g[i] = a
            elif (is_sym(a) or is_sym(b)) and a is not POISON and b is not POISON:
                st.g[i] = select(c, a, b)
            else:
                st.g[i] = POISON
            for k in range(2):
                a, b = s1.x[i][k], s2.x[i][k]
                if a is b or (not is_sym(a) and not is_sym(b) and a == b):
                    st.x[i][k] = a
                else:
                    st.x[i][k] = select(c, a, b)
        # deterministic order, relative to the image / stack bases (absolute addresses change from load to load)
        keys = sorted(set(s1.mem.d) | set(s2.mem.d), key=lambda a: (a - self.img.base) & M64)
        for a in keys:
            v1 = s1.rd64(a) if s1.mem.get(a) is not POISON else POISON
            v2 = s2.rd64(a) if s2.mem.get(a) is not POISON else POISON
            if v1 is v2 or (not is_sym(v1) and not is_sym(v2) and v1 is not POISON and v2 is not POISON and v1 == v2):
                st.mem.d[a] = v1
            elif v1 is POISON or v2 is POISON:
                st.mem.d[a] = POISON
            else:
                st.mem.d[a] = select(c, v1, v2)
        st.flags = None

    def step_insn(self, st, ins):
        op = ins.op
        o = ins.ops
        if op in ('nop', 'nopw', 'nopl', 'endbr64', 'xchg') and (op != 'xchg' or o[0] == o[1]):
            return
        # ---- SSE scalar double ----
        if op == 'movsd':
            s, d = o
            if self.is_xmm(d):
                if self.is_xmm(s):
                    st.x[int(d[4:])][0] = st.x[int(s[4:])][0]
                else:
                    st.x[int(d[4:])] = [st.rd64(self.mem_addr(st, ins, s)), 0]
            else:
                st.wr64(self.mem_addr(st, ins, d), st.x[int(s[4:])][0])
            return
        if op in ('movapd', 'movaps', 'movdqa', 'movups', 'movdqu', 'movupd'):
            s, d = o
            if self.is_xmm(s):
                v = list(st.x[int(s[4:])])
            else:
                a = self.mem_addr(st, ins, s)
                v = [st.rd64(a), st.rd64(a + 8)]
            if self.is_xmm(d):
                st.x[int(d[4:])] = v
            else:
                a = self.mem_addr(st, ins, d)
                st.wr64(a, v[0])
                st.wr64(a + 8, v[1])
            return
        if op in ('addsd', 'subsd', 'mulsd', 'divsd', 'maxsd', 'minsd'):
            s, d = o
            b = self.rd_f64(st, ins, s)
            di = int(d[4:])
            st.x[di][0] = fop2(op[:3], st.x[di][0], b)
            return
        if op == 'sqrtsd':
            s, d = o
            st.x[int(d[4:])][0] = fsqrt(self.rd_f64(st, ins, s))
            return
        if op in ('comisd', 'ucomisd'):
            s, d = o
            st.flags = ('fp', st.x[int(d[4:])][0], self.rd_f64(st, ins, s))
            return
        if op in ('andpd', 'andnpd', 'orpd', 'xorpd', 'pxor', 'pand', 'por', 'xorps', 'andps'):
            s, d = o
            di = int(d[4:])
            if self.is_xmm(s):
                sv = st.x[int(s[4:])]
            else:
                a = self.mem_addr(st, ins, s)
                sv = [st.rd64(a), st.rd64(a + 8)]
            kind = {'andpd': 'and', 'andnpd': 'andn', 'orpd': 'or', 'xorpd': 'xor', 'pxor': 'xor', 'pand': 'and', 'por': 'or', 'xorps': 'xor', 'andps': 'and'}[op]
            dv = st.x[di]
            res = []
            for k in range(2):
                if kind == 'andn':
                    res.append(bitop('andn', dv[k], sv[k]))
                else:
                    try:
                        res.append(bitop(kind, dv[k], sv[k]))
                    except NotImplementedError:
                        if k == 1:
                            res.append(0)    # high lane garbage is never consumed
                        else:
                            raise
            st.x[di] = res
            return
        if op in ('cmplesd', 'cmpnlesd', 'cmpltsd', 'cmpnltsd', 'cmpeqsd', 'cmpneqsd', 'cmpunordsd', 'cmpordsd'):
            s, d = o
            di = int(d[4:])
            a = st.x[di][0]
            b = self.rd_f64(st, ins, s)
            pred = {'cmplesd': 'le', 'cmpnlesd': 'gt', 'cmpltsd': 'lt', 'cmpnltsd': 'ge', 'cmpeqsd': 'eq', 'cmpneqsd': 'ne'}.get(op)
            if pred is None:
                raise TraceError(op)
            if is_sym(a) or is_sym(b):
                st.x[di][0] = G.mk('cmpmask', pred, G.lift(a), G.lift(b))
            else:
                x, y = fval(a), fval(b)
                r = {'le': x <= y, 'gt': not (x <= y), 'lt': x < y, 'ge': not (x < y), 'eq': x == y, 'ne': x != y}[pred]
                st.x[di][0] = M64 if r else 0
            return
        if op == 'unpcklpd':
            s, d = o
            di = int(d[4:])
            sv = st.x[int(s[4:])][0] if self.is_xmm(s) else st.rd64(self.mem_addr(st, ins, s))
            st.x[di] = [st.x[di][0], sv]
            return
        if op == 'movhpd':
            s, d = o
            if self.is_xmm(d):
                st.x[int(d[4:])][1] = st.rd64(self.mem_addr(st, ins, s))
            else:
                st.wr64(self.mem_addr(st, ins, d), st.x[int(s[4:])][1])
            return
        if op == 'movq' and (self.is_xmm(o[0]) or self.is_xmm(o[1])):
            s, d = o
            if self.is_xmm(d):
                if self.is_xmm(s):
                    v = st.x[int(s[4:])][0]
                elif self.is_reg(s):
                    v = st.g[REGMAP[s[1:]][0]]
                else:
                    v = st.rd64(self.mem_addr(st, ins, s))
                st.x[int(d[4:])] = [v, 0]
            else:
                v = st.x[int(s[4:])][0]
                if self.is_reg(d):
                    st.g[REGMAP[d[1:]][0]] = v
                else:
                    st.wr64(self.mem_addr(st, ins, d), v)
            return
        if op == 'cvtsi2sd' or op == 'cvtsi2sdl' or op == 'cvtsi2sdq':
            s, d = o
            n = 8 if op.endswith('q') else (REGMAP[s[1:]][1] if self.is_reg(s) else 4)
            v = self.rd_int(st, ins, s, n)
            if is_sym(v):
                raise TraceError('cvtsi2sd symbolic')
            if v >> (8 * n - 1):
                v -= 1 << (8 * n)
            st.x[int(d[4:])][0] = fbits(float(v))
            return
        if op == 'cvttsd2si':
            s, d = o
            v = self.rd_f64(st, ins, s)
            if is_sym(v):
                raise TraceError('cvttsd2si of symbolic value at %x' % ins.addr)
            self.sreg(st, d[1:], int(fval(v)) & M64)
            return
        # ---- integer ----
        if op in ('mov', 'movl', 'movq', 'movb', 'movw', 'movabs'):
            s, d = o
            if self.is_reg(s) and self.is_reg(d) and REGMAP[s[1:]][1] == 8:
                st.g[REGMAP[d[1:]][0]] = st.g[REGMAP[s[1:]][0]]
                return
            n = self.opsize(ins, s, d)
            if n == 8:
                # may move symbolic doubles through GPRs / memory
                if s.startswith('$'):
                    v = int(s[1:], 16) & M64
                elif self.is_reg(s):
                    v = st.g[REGMAP[s[1:]][0]]
                else:
                    a = self.mem_addr(st, ins, s)
                    v = 0 if isinstance(a, tuple) else st.rd64(a)
                if self.is_reg(d):
                    st.g[REGMAP[d[1:]][0]] = v
                else:
                    st.wr64(self.mem_addr(st, ins, d), v)
                return
            v = self.rd_int(st, ins, s, n)
            self.wr_int(st, ins, d, v, n)
            return
        if op in ('movslq', 'movsbl', 'movswl', 'movsbq', 'movswq', 'cltq', 'movzbl', 'movzwl', 'movzbw'):
            if op == 'cltq':
                v = st.g[0] & 0xffffffff
                if v >> 31:
                    v |= 0xffffffff00000000
                st.g[0] = v
                return
            s, d = o
            sn = {'movslq': 4, 'movsbl': 1, 'movswl': 2, 'movsbq': 1, 'movswq': 2, 'movzbl': 1, 'movzwl': 2, 'movzbw': 1}[op]
            v = self.rd_int(st, ins, s, sn)
            if op.startswith('movs') and (v >> (8 * sn - 1)):
                v |= (M64 << (8 * sn)) & M64
            self.sreg(st, d[1:], v)
            return
        if op == 'lea':
            s, d = o
            a = self.mem_addr(st, ins, s)
            self.sreg(st, d[1:], a)
            return
        if op in ('add', 'sub', 'and', 'or', 'xor', 'cmp', 'test', 'addl', 'subl', 'cmpl', 'cmpq', 'addq', 'subq', 'andl', 'orl', 'orb', 'andb', 'testb', 'cmpb', 'imul', 'testl'):
            s, d = o[0], o[1]
            base = op.rstrip('lqb') if op not in ('imul', 'sub', 'subl') else op
            base = {'addl': 'add', 'addq': 'add', 'subl': 'sub', 'subq': 'sub', 'cmpl': 'cmp', 'cmpq': 'cmp', 'cmpb': 'cmp', 'andl': 'and',
                    'andb': 'and', 'orl': 'or', 'orb': 'or', 'testb': 'test', 'testl': 'test'}.get(op, op)
            n = self.opsize(ins, s, d)
            if base == 'xor' and s == d:
                self.sreg(st, d[1:], 0)
                st.flags = ('int', 'logic', 0, 0, 0, n)
                return
            a = self.rd_int(st, ins, d, n)
            b = self.rd_int(st, ins, s, n)
            if is_sym(a) or is_sym(b):
                raise TraceError('integer op on symbolic value at %x: %s' % (ins.addr, ins.text))
            mask = (1 << (8 * n)) - 1
            if s.startswith('$'):
                # sign-extended imm32
                iv = int(s[1:], 16)
                if n == 8 and iv >> 31 and iv < (1 << 32):
                    iv |= 0xffffffff00000000
                b = iv & mask
            if base in ('add',):
                r = (a + b) & mask
                st.flags = ('int', 'add', r, a, b, n)
                self.wr_int(st, ins, d, r, n)
            elif base in ('sub', 'cmp'):
                r = (a - b) & mask
                st.flags = ('int', 'sub', r, a, b, n)
                if base == 'sub':
                    self.wr_int(st, ins, d, r, n)
            elif base in ('and', 'test'):
                r = a & b
                st.flags = ('int', 'logic', r, a, b, n)
                if base == 'and':
                    self.wr_int(st, ins, d, r, n)
            elif base == 'or':
                r = a | b
                st.flags = ('int', 'logic', r, a, b, n)
                self.wr_int(st, ins, d, r, n)
            elif base == 'xor':
                r = a ^ b
                st.flags = ('int', 'logic', r, a, b, n)
                self.wr_int(st, ins, d, r, n)
            elif base == 'imul':
                sa = a - ((a >> (8 * n - 1)) << (8 * n))
                sb = b - ((b >> (8 * n - 1)) << (8 * n))
                r = (sa * sb) & mask
                self.wr_int(st, ins, d, r, n)
                st.flags = None
            else:
                raise TraceError('int op ' + op)
            return
        if op in ('shr', 'sar', 'shl', 'sal'):
            if len(o) == 1:
                s, d = '$0x1', o[0]
            else:
                s, d = o
            n = self.opsize(ins, d)
            a = self.rd_int(st, ins, d, n)
            c = self.rd_int(st, ins, s, 1) & (63 if n == 8 else 31)
            mask = (1 << (8 * n)) - 1
            if op == 'shr':
                r = (a & mask) >> c
            elif op == 'sar':
                sa = a - ((a >> (8 * n - 1)) << (8 * n))
                r = (sa >> c) & mask
            else:
                r = (a << c) & mask
            self.wr_int(st, ins, d, r, n)
            st.flags = ('int', 'logic', r, a, c, n)
            return
        if op == 'btc':
            s, d = o
            bit = int(s[1:], 16)
            i = REGMAP[d[1:]][0]
            v = st.g[i]
            if is_sym(v):
                if bit != 63:
                    raise TraceError('btc on symbolic')
                st.g[i] = G.mk('neg', v)
            else:
                st.g[i] = v ^ (1 << bit)
            st.flags = None
            return
        if op in ('push', 'pop'):
            r = o[0]
            i = REGMAP[r[1:]][0]
            if op == 'push':
                st.g[4] = (st.g[4] - 8) & M64
                st.wr64(st.g[4], st.g[i] if st.g[i] is not POISON else 0)
            else:
                st.g[i] = st.rd64(st.g[4])
                st.g[4] = (st.g[4] + 8) & M64
            return
        if op.startswith('set'):
            c = self.cond(st, op[3:], ins)
            if c is not True and c is not False:
                raise TraceError('setcc on symbolic flags at %x' % ins.addr)
            self.sreg(st, o[0][1:], 1 if c else 0)
            return
        if op.startswith('cmov'):
            c = self.cond(st, op[4:], ins)
            if c is not True and c is not False:
                raise TraceError('cmov on symbolic flags at %x' % ins.addr)
            s, d = o
            n = self.opsize(ins, s, d)
            if c:
                self.wr_int(st, ins, d, self.rd_int(st, ins, s, n), n)
            elif n == 4:
                self.sreg(st, d[1:], self.greg(st, d[1:]))
            return
        if op == 'stos' and 'rep' in ins.text and '%rax' in ins.text:
            cnt = st.g[1]
            for k in range(cnt):
                st.wr64(st.g[7] + 8 * k, st.g[0])
            st.g[7] = (st.g[7] + 8 * cnt) & M64
            st.g[1] = 0
            return
        raise TraceError('unhandled instruction %x: %s' % (ins.addr, ins.text))

"""Flat expression DAG ("tape") used by the SymPy -> CUDA-C emitter.

This replaces the SymPy-level machinery the reference uses between the
discretised equations of motion and the generated C code:

- ``sm.cse`` + ``ccode`` inside ``ufuncify_matrix`` (opty/utils.py:745-757)
- the CSE-aware forward-mode symbolic Jacobian ``_forward_jacobian``
  (opty/utils.py:82-228)

Instead of differentiating SymPy trees, expressions are lowered once to a
hash-consed DAG of binary/unary double-precision operations and all further
work (forward-mode differentiation, dead-code elimination, node-invariance
analysis, partitioning into output groups) happens on integer ids.  That keeps
the set-up cost linear in the DAG size (seconds where the reference needs
minutes for large models) and gives the emitter what it needs to place every
value: *literal*, *node-invariant* (depends only on parameters / the time
interval: computed once per call into constant memory) or *node-varying*
(computed per collocation node).

All arithmetic is IEEE float64; this module does no numerical work itself
besides folding literal constants.
"""

import math

# ---------------------------------------------------------------------------
# op codes
# ---------------------------------------------------------------------------
CONST = 0      # val
VIN = 1        # node-varying input, a = slot
UIN = 2        # uniform (node-invariant) input, a = slot
NEG = 3
ADD = 4
SUB = 5
MUL = 6
DIV = 7
SQRT = 8
POW = 9        # general a**b
EXP = 10
LOG = 11
SIN = 12
COS = 13
TAN = 14
ASIN = 15
ACOS = 16
ATAN = 17
SINH = 18
COSH = 19
TANH = 20
ABS = 21
SIGN = 22
ATAN2 = 23
MIN = 24
MAX = 25
SEL = 26       # a ? b : c  (a is a condition node)
LT = 27
LE = 28
EQ = 29
NE = 30
AND = 31
OR = 32
NOT = 33
ASINH = 34
ACOSH = 35
ATANH = 36
FLOOR = 37
CEIL = 38
CBRT = 39
ERF = 40

OP_NAMES = {
    CONST: 'const', VIN: 'vin', UIN: 'uin', NEG: 'neg', ADD: 'add', SUB: 'sub',
    MUL: 'mul', DIV: 'div', SQRT: 'sqrt', POW: 'pow', EXP: 'exp', LOG: 'log',
    SIN: 'sin', COS: 'cos', TAN: 'tan', ASIN: 'asin', ACOS: 'acos',
    ATAN: 'atan', SINH: 'sinh', COSH: 'cosh', TANH: 'tanh', ABS: 'fabs',
    SIGN: 'sign', ATAN2: 'atan2', MIN: 'fmin', MAX: 'fmax', SEL: 'sel',
    LT: 'lt', LE: 'le', EQ: 'eq', NE: 'ne', AND: 'and', OR: 'or', NOT: 'not',
    ASINH: 'asinh', ACOSH: 'acosh', ATANH: 'atanh', FLOOR: 'floor',
    CEIL: 'ceil', CBRT: 'cbrt', ERF: 'erf',
}

UNARY_MATH = {SQRT, EXP, LOG, SIN, COS, TAN, ASIN, ACOS, ATAN, SINH, COSH,
              TANH, ABS, ASINH, ACOSH, ATANH, FLOOR, CEIL, CBRT, ERF}
BOOL_OPS = {LT, LE, EQ, NE, AND, OR, NOT}

# rough issue-slot cost of one op in FP64 instructions on sm_100a; only used to
# balance work between output groups, not for any reported number
OP_COST = {
    CONST: 0, VIN: 0, UIN: 0, NEG: 0.5, ADD: 1, SUB: 1, MUL: 1, DIV: 12,
    SQRT: 14, POW: 80, EXP: 25, LOG: 30, SIN: 40, COS: 40, TAN: 60, ASIN: 40,
    ACOS: 40, ATAN: 40, SINH: 40, COSH: 40, TANH: 40, ABS: 0.5, SIGN: 3,
    ATAN2: 60, MIN: 1, MAX: 1, SEL: 1, LT: 1, LE: 1, EQ: 1, NE: 1, AND: 1,
    OR: 1, NOT: 1, ASINH: 50, ACOSH: 50, ATANH: 50, FLOOR: 1, CEIL: 1,
    CBRT: 30, ERF: 40,
}

_FOLD_UNARY = {
    SQRT: math.sqrt, EXP: math.exp, LOG: math.log, SIN: math.sin,
    COS: math.cos, TAN: math.tan, ASIN: math.asin, ACOS: math.acos,
    ATAN: math.atan, SINH: math.sinh, COSH: math.cosh, TANH: math.tanh,
    ABS: abs, ASINH: math.asinh, ACOSH: math.acosh, ATANH: math.atanh,
    FLOOR: lambda v: float(math.floor(v)), CEIL: lambda v: float(math.ceil(v)),
    CBRT: lambda v: math.copysign(abs(v) ** (1.0 / 3.0), v), ERF: math.erf,
}


class Tape(object):
    """Hash-consed DAG of float64 operations.

    Node ids are dense integers in creation order, so operands always have
    smaller ids than the node that uses them (the tape is topologically
    sorted by construction).
    """

    def __init__(self):
        self.op = []
        self.a = []
        self.b = []
        self.c = []
        self.val = []
        self.varying = []   # True if the node depends on a node-varying input
        self._memo = {}
        self._const_memo = {}
        self.zero = self.const(0.0)
        self.one = self.const(1.0)

    def __len__(self):
        return len(self.op)

    # -- construction ------------------------------------------------------
    def _new(self, op, a=-1, b=-1, c=-1, val=0.0, varying=False):
        self.op.append(op)
        self.a.append(a)
        self.b.append(b)
        self.c.append(c)
        self.val.append(val)
        self.varying.append(varying)
        return len(self.op) - 1

    def const(self, v):
        v = float(v)
        if v == 0.0:
            v = 0.0  # merge -0.0 with 0.0, SymPy has no signed zero either
        key = v.hex() if v == v else 'nan'
        i = self._const_memo.get(key)
        if i is None:
            i = self._new(CONST, val=v)
            self._const_memo[key] = i
        return i

    def is_const(self, i):
        return self.op[i] == CONST

    def cval(self, i):
        return self.val[i]

    def vin(self, slot):
        key = (VIN, slot)
        i = self._memo.get(key)
        if i is None:
            i = self._new(VIN, a=slot, varying=True)
            self._memo[key] = i
        return i

    def uin(self, slot):
        key = (UIN, slot)
        i = self._memo.get(key)
        if i is None:
            i = self._new(UIN, a=slot, varying=False)
            self._memo[key] = i
        return i

    def _node(self, op, a, b=-1, c=-1):
        key = (op, a, b, c)
        i = self._memo.get(key)
        if i is None:
            var = self.varying[a]
            if b >= 0:
                var = var or self.varying[b]
            if c >= 0:
                var = var or self.varying[c]
            i = self._new(op, a, b, c, varying=var)
            self._memo[key] = i
        return i

    # -- arithmetic with local simplification -------------------------------
    def neg(self, x):
        op = self.op
        if op[x] == CONST:
            return self.const(-self.val[x])
        if op[x] == NEG:
            return self.a[x]
        if op[x] == SUB:
            return self.sub(self.b[x], self.a[x])
        return self._node(NEG, x)

    def add(self, x, y):
        op = self.op
        if op[x] == CONST and op[y] == CONST:
            return self.const(self.val[x] + self.val[y])
        if op[x] == CONST and self.val[x] == 0.0:
            return y
        if op[y] == CONST and self.val[y] == 0.0:
            return x
        if op[y] == NEG:
            return self.sub(x, self.a[y])
        if op[x] == NEG:
            return self.sub(y, self.a[x])
        if x > y:
            x, y = y, x
        return self._node(ADD, x, y)

    def sub(self, x, y):
        op = self.op
        if x == y:
            return self.zero
        if op[x] == CONST and op[y] == CONST:
            return self.const(self.val[x] - self.val[y])
        if op[y] == CONST and self.val[y] == 0.0:
            return x
        if op[x] == CONST and self.val[x] == 0.0:
            return self.neg(y)
        if op[y] == NEG:
            return self.add(x, self.a[y])
        return self._node(SUB, x, y)

    def mul(self, x, y):
        op = self.op
        if op[x] == CONST and op[y] == CONST:
            return self.const(self.val[x] * self.val[y])
        if op[y] == CONST:
            x, y = y, x
        if op[x] == CONST:
            v = self.val[x]
            if v == 0.0:
                return self.zero
            if v == 1.0:
                return y
            if v == -1.0:
                return self.neg(y)
            if op[y] == NEG:
                return self.mul(self.const(-v), self.a[y])
        if op[x] == NEG and op[y] == NEG:
            return self.mul(self.a[x], self.a[y])
        if op[x] == NEG:
            return self.neg(self.mul(self.a[x], y))
        if op[y] == NEG:
            return self.neg(self.mul(x, self.a[y]))
        if x > y:
            x, y = y, x
        return self._node(MUL, x, y)

    def div(self, x, y):
        op = self.op
        if op[y] == CONST:
            v = self.val[y]
            if v == 1.0:
                return x
            if v == -1.0:
                return self.neg(x)
            if op[x] == CONST and v != 0.0:
                return self.const(self.val[x] / v)
            # division by a literal power of two is an exact scaling
            if v != 0.0 and math.frexp(v)[0] in (0.5, -0.5):
                return self.mul(self.const(1.0 / v), x)
        if op[x] == CONST and self.val[x] == 0.0:
            return self.zero
        if op[x] == NEG and op[y] == NEG:
            return self.div(self.a[x], self.a[y])
        if op[x] == NEG:
            return self.neg(self.div(self.a[x], y))
        if op[y] == NEG:
            return self.neg(self.div(x, self.a[y]))
        return self._node(DIV, x, y)

    def recip(self, x):
        return self.div(self.one, x)

    def powi(self, x, n):
        """x**n for a Python integer n, expanded to multiplications (the
        reference prints ``pow(x, n)`` through C99CodePrinter; glibc's pow is
        correctly rounded for these cases in practice, so ``x*x`` agrees)."""
        if n == 0:
            return self.one
        if n < 0:
            return self.recip(self.powi(x, -n))
        result = None
        base = x
        while n:
            if n & 1:
                result = base if result is None else self.mul(result, base)
            n >>= 1
            if n:
                base = self.mul(base, base)
        return result

    def unary(self, opcode, x):
        if self.op[x] == CONST and opcode in _FOLD_UNARY:
            try:
                return self.const(_FOLD_UNARY[opcode](self.val[x]))
            except (ValueError, OverflowError):
                pass
        if opcode == ABS and self.op[x] == NEG:
            return self.unary(ABS, self.a[x])
        return self._node(opcode, x)

    def sign(self, x):
        if self.op[x] == CONST:
            v = self.val[x]
            return self.const((v > 0) - (v < 0))
        return self._node(SIGN, x)

    def pow(self, x, y):
        if self.op[y] == CONST:
            v = self.val[y]
            if v == int(v) and abs(v) <= 64:
                return self.powi(x, int(v))
            if v == 0.5:
                return self.unary(SQRT, x)
            if v == -0.5:
                return self.recip(self.unary(SQRT, x))
            if v == 1.5:
                return self.mul(x, self.unary(SQRT, x))
            if v == -1.5:
                return self.recip(self.mul(x, self.unary(SQRT, x)))
            if self.op[x] == CONST:
                try:
                    return self.const(math.pow(self.val[x], v))
                except (ValueError, OverflowError):
                    pass
        return self._node(POW, x, y)

    def binary(self, opcode, x, y):
        if opcode in (MIN, MAX) and x > y:
            x, y = y, x
        return self._node(opcode, x, y)

    def cmp(self, opcode, x, y):
        return self._node(opcode, x, y)

    def logic_not(self, x):
        return self._node(NOT, x)

    def sel(self, cond, x, y):
        if x == y:
            return x
        return self._node(SEL, cond, x, y)

    # -- analysis ----------------------------------------------------------
    def operands(self, i):
        a, b, c = self.a[i], self.b[i], self.c[i]
        o = self.op[i]
        if o in (CONST, VIN, UIN):
            return ()
        if c >= 0:
            return (a, b, c)
        if b >= 0:
            return (a, b)
        return (a,)

    def reachable(self, roots):
        """Returns the sorted list of node ids reachable from ``roots``."""
        seen = bytearray(len(self.op))
        stack = [r for r in roots]
        a_, b_, c_, op_ = self.a, self.b, self.c, self.op
        while stack:
            i = stack.pop()
            if seen[i]:
                continue
            seen[i] = 1
            o = op_[i]
            if o <= UIN:
                continue
            stack.append(a_[i])
            if b_[i] >= 0:
                stack.append(b_[i])
                if c_[i] >= 0:
                    stack.append(c_[i])
        return [i for i in range(len(seen)) if seen[i]]

    def flops(self, ids):
        """Number of arithmetic/math operations among ``ids``."""
        op_ = self.op
        return sum(1 for i in ids if op_[i] > UIN)

    def cost(self, ids):
        op_ = self.op
        return sum(OP_COST[op_[i]] for i in ids)

    # -- evaluation (set-up time checks and tests; scalar, slow) ------------
    def evaluate(self, roots, vin_vals, uin_vals):
        """Evaluates ``roots`` numerically with Python floats.

        Used only for unit tests of the lowering/differentiation; the product
        path evaluates tapes exclusively through the emitted CUDA code.
        """
        ids = self.reachable(roots)
        vals = {}
        for i in ids:
            vals[i] = self._eval_one(i, vals, vin_vals, uin_vals)
        return [vals[r] for r in roots]

    def _eval_one(self, i, vals, vin_vals, uin_vals):
        o = self.op[i]
        a, b, c = self.a[i], self.b[i], self.c[i]
        if o == CONST:
            return self.val[i]
        if o == VIN:
            return float(vin_vals[a])
        if o == UIN:
            return float(uin_vals[a])
        x = vals[a]
        if o == NEG:
            return -x
        if o == ADD:
            return x + vals[b]
        if o == SUB:
            return x - vals[b]
        if o == MUL:
            return x * vals[b]
        if o == DIV:
            return x / vals[b]
        if o == POW:
            return math.pow(x, vals[b])
        if o in _FOLD_UNARY:
            return _FOLD_UNARY[o](x)
        if o == SIGN:
            return float((x > 0) - (x < 0))
        if o == ATAN2:
            return math.atan2(x, vals[b])
        if o == MIN:
            return min(x, vals[b])
        if o == MAX:
            return max(x, vals[b])
        if o == SEL:
            return vals[b] if x else vals[c]
        if o == LT:
            return x < vals[b]
        if o == LE:
            return x <= vals[b]
        if o == EQ:
            return x == vals[b]
        if o == NE:
            return x != vals[b]
        if o == AND:
            return bool(x) and bool(vals[b])
        if o == OR:
            return bool(x) or bool(vals[b])
        if o == NOT:
            return not x
        raise NotImplementedError(OP_NAMES[o])


# ---------------------------------------------------------------------------
# forward-mode differentiation on the tape
# ---------------------------------------------------------------------------

def forward_jacobian(tape, outputs, wrt_nodes, chain=None):
    """Sparse vector forward-mode derivative of ``outputs`` with respect to
    the input nodes ``wrt_nodes``.

    Plays the role of ``opty.utils._forward_jacobian`` (opty/utils.py:82-228):
    every intermediate carries a sparse row ``{k: d(node)/d(wrt_k)}`` that is
    propagated in definition order by the chain rule (opty/utils.py:167-185),
    and only what the requested outputs need survives (opty/utils.py:187-209
    does this with an explicit pruning pass; here unreachable derivative nodes
    are simply never emitted because the emitter walks from the outputs).

    ``chain`` maps an input node that is itself a function of another input
    (an implicit known trajectory ``r(x)``) to ``[(x node, dr/dx node)]``.

    Returns
    -------
    rows : list of dict
        ``rows[j][k]`` is the tape id of ``d outputs[j] / d wrt_nodes[k]``;
        missing keys are structural zeros.
    """
    chain = chain or {}
    T = tape
    seeds = {}
    for k, w in enumerate(wrt_nodes):
        seeds.setdefault(w, []).append(k)

    ids = T.reachable(outputs)   # sorted => topological
    d = {}
    op_, a_, b_, c_ = T.op, T.a, T.b, T.c
    empty = {}

    def axpy(acc, coeff, row):
        """acc += coeff * row (coeff a tape id or None for 1)."""
        for k, g in row.items():
            term = g if coeff is None else T.mul(coeff, g)
            prev = acc.get(k)
            acc[k] = term if prev is None else T.add(prev, term)

    def axmy(acc, coeff, row):
        """acc -= coeff * row."""
        for k, g in row.items():
            term = g if coeff is None else T.mul(coeff, g)
            prev = acc.get(k)
            acc[k] = T.neg(term) if prev is None else T.sub(prev, term)

    for i in ids:
        o = op_[i]
        if o == CONST:
            continue
        if o == VIN or o == UIN:
            ks = seeds.get(i)
            row = {k: T.one for k in ks} if ks else {}
            for var, deriv in chain.get(i, ()):
                for k in seeds.get(var, ()):
                    row[k] = deriv if k not in row else T.add(row[k], deriv)
            if row:
                d[i] = row
            continue
        a = a_[i]
        b = b_[i]
        da = d.get(a, empty)
        db = d.get(b, empty) if b >= 0 else empty
        if o in BOOL_OPS:
            continue
        if o == SEL:
            dbb = d.get(b, empty)
            dcc = d.get(c_[i], empty)
            if not dbb and not dcc:
                continue
            row = {}
            for k in set(dbb) | set(dcc):
                row[k] = T.sel(a, dbb.get(k, T.zero), dcc.get(k, T.zero))
            d[i] = row
            continue
        if not da and not db:
            continue
        row = {}
        if o == NEG:
            row = {k: T.neg(g) for k, g in da.items()}
        elif o == ADD:
            row = dict(da)
            axpy(row, None, db)
        elif o == SUB:
            row = dict(da)
            axmy(row, None, db)
        elif o == MUL:
            if da:
                axpy(row, b, da)
            if db:
                axpy(row, a, db)
        elif o == DIV:
            # d(a/b) = da/b - (a/b) db / b
            if da:
                if op_[b] == CONST or not T.varying[b]:
                    inv = T.recip(b)
                    axpy(row, inv, da)
                else:
                    for k, g in da.items():
                        row[k] = T.div(g, b)
            if db:
                q = T.div(i, b)   # (a/b)/b
                axmy(row, q, db)
        elif o == SQRT:
            coeff = T.div(T.const(0.5), i)
            axpy(row, coeff, da)
        elif o == POW:
            # d(a**b) = b a**(b-1) da + a**b log(a) db
            if da:
                coeff = T.mul(b, T.pow(a, T.sub(b, T.one)))
                axpy(row, coeff, da)
            if db:
                coeff = T.mul(i, T.unary(LOG, a))
                axpy(row, coeff, db)
        elif o == EXP:
            axpy(row, i, da)
        elif o == LOG:
            axpy(row, T.recip(a), da)
        elif o == SIN:
            axpy(row, T.unary(COS, a), da)
        elif o == COS:
            axmy(row, T.unary(SIN, a), da)
        elif o == TAN:
            axpy(row, T.add(T.one, T.mul(i, i)), da)
        elif o == ASIN:
            axpy(row, T.recip(T.unary(SQRT, T.sub(T.one, T.mul(a, a)))), da)
        elif o == ACOS:
            axmy(row, T.recip(T.unary(SQRT, T.sub(T.one, T.mul(a, a)))), da)
        elif o == ATAN:
            axpy(row, T.recip(T.add(T.one, T.mul(a, a))), da)
        elif o == SINH:
            axpy(row, T.unary(COSH, a), da)
        elif o == COSH:
            axpy(row, T.unary(SINH, a), da)
        elif o == TANH:
            axpy(row, T.sub(T.one, T.mul(i, i)), da)
        elif o == ASINH:
            axpy(row, T.recip(T.unary(SQRT, T.add(T.mul(a, a), T.one))), da)
        elif o == ACOSH:
            axpy(row, T.recip(T.unary(SQRT, T.sub(T.mul(a, a), T.one))), da)
        elif o == ATANH:
            axpy(row, T.recip(T.sub(T.one, T.mul(a, a))), da)
        elif o == CBRT:
            axpy(row, T.recip(T.mul(T.const(3.0), T.mul(i, i))), da)
        elif o == ERF:
            coeff = T.mul(T.const(2.0 / math.sqrt(math.pi)),
                          T.unary(EXP, T.neg(T.mul(a, a))))
            axpy(row, coeff, da)
        elif o == ABS:
            axpy(row, T.sign(a), da)
        elif o in (SIGN, FLOOR, CEIL):
            row = {}
        elif o == ATAN2:
            # d atan2(a, b) = (b da - a db) / (a^2 + b^2)
            den = T.add(T.mul(a, a), T.mul(b, b))
            if da:
                axpy(row, T.div(b, den), da)
            if db:
                axmy(row, T.div(a, den), db)
        elif o in (MIN, MAX):
            cond = T.cmp(LE, a, b) if o == MIN else T.cmp(LE, b, a)
            for k in set(da) | set(db):
                row[k] = T.sel(cond, da.get(k, T.zero), db.get(k, T.zero))
        else:
            raise NotImplementedError(
                'No derivative rule for op {}'.format(OP_NAMES[o]))
        # drop entries that simplified to literal zero
        row = {k: g for k, g in row.items()
               if not (op_[g] == CONST and T.val[g] == 0.0)}
        if row:
            d[i] = row

    return [d.get(o, {}) for o in outputs]

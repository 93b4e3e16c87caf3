"""SymPy expression -> :class:`opty_b200.ir.Tape` lowering.

The reference turns SymPy expressions into C source text with
``OptyC99CodePrinter`` (opty/utils.py:61-79) after ``sm.cse`` (opty/utils.py:
748-749).  Here they are lowered to the tape instead; the CUDA-C text is
produced from the tape by :mod:`opty_b200.codegen`.
"""

import sympy as sm
from sympy.core.function import AppliedUndef

from . import ir

_UNARY = {
    sm.sin: ir.SIN, sm.cos: ir.COS, sm.tan: ir.TAN, sm.asin: ir.ASIN,
    sm.acos: ir.ACOS, sm.atan: ir.ATAN, sm.sinh: ir.SINH, sm.cosh: ir.COSH,
    sm.tanh: ir.TANH, sm.exp: ir.EXP, sm.log: ir.LOG, sm.Abs: ir.ABS,
    sm.asinh: ir.ASINH, sm.acosh: ir.ACOSH, sm.atanh: ir.ATANH,
    sm.floor: ir.FLOOR, sm.ceiling: ir.CEIL, sm.erf: ir.ERF,
}


class Lowerer(object):
    """Lowers SymPy expressions onto one tape.

    Parameters
    ----------
    tape : ir.Tape
    leaf_map : dict
        Maps SymPy leaves (Symbols or applied functions) to tape ids.
    """

    def __init__(self, tape, leaf_map):
        self.tape = tape
        self.memo = dict(leaf_map)

    def bind(self, symbol, node):
        self.memo[symbol] = node

    def lower(self, expr):
        memo = self.memo
        hit = memo.get(expr)
        if hit is not None:
            return hit
        # iterative post-order walk: SymPy trees of large multibody models are
        # deep enough to overflow Python's recursion limit
        stack = [(expr, False)]
        while stack:
            node, expanded = stack.pop()
            if node in memo:
                continue
            if expanded or not node.args or isinstance(node, sm.Number):
                memo[node] = self._emit(node)
                continue
            stack.append((node, True))
            for arg in self._children(node):
                if arg not in memo:
                    stack.append((arg, False))
        return memo[expr]

    @staticmethod
    def _children(node):
        if isinstance(node, sm.Piecewise):
            out = []
            for e, c in node.args:
                out.append(e)
                if c is not sm.true and c is not sm.false:
                    out.append(c)
            return out
        if isinstance(node, sm.Pow):
            base, exp = node.args
            if exp.is_Integer or (exp.is_Rational and exp.q == 2):
                return [base]
            return [base, exp]
        return node.args

    def _emit(self, e):
        T = self.tape
        memo = self.memo
        if isinstance(e, sm.Number) or isinstance(e, sm.NumberSymbol):
            if e.is_Rational and not e.is_Integer:
                # mirrors the C printer's ``p.0/q.0`` literal
                return T.const(float(e.p) / float(e.q))
            return T.const(float(e))
        if e is sm.true:
            return T.one
        if e is sm.false:
            return T.zero
        if isinstance(e, sm.Symbol) or isinstance(e, AppliedUndef):
            raise ValueError(
                '{} appears in the expressions but is not an argument of the '
                'generated function.'.format(e))
        if isinstance(e, sm.Add):
            # same left-to-right summation order as the C text the reference
            # prints (CodePrinter._print_Add orders with as_ordered_terms)
            ids = [memo[a] if a in memo else self.lower(a)
                   for a in e.as_ordered_terms()]
            acc = ids[0]
            for i in ids[1:]:
                acc = T.add(acc, i)
            return acc
        if isinstance(e, sm.Mul):
            # numerator factors left to right in printer order, then one
            # division by the product of the denominator factors
            # (CodePrinter._print_Mul)
            coeff = None
            num = []
            den = []
            for a in e.as_ordered_factors():
                if isinstance(a, sm.Number):
                    # the ordered factors may split the sign off the
                    # coefficient; fold all numeric factors into one literal
                    ca = memo[a] if a in memo else self._emit(a)
                    coeff = ca if coeff is None else T.mul(coeff, ca)
                elif (isinstance(a, sm.Pow) and a.exp.is_Rational and
                      a.exp.is_negative):
                    den.append(self._lower_pow(a.base, -a.exp))
                else:
                    num.append(memo[a] if a in memo else self.lower(a))
            acc = None
            if coeff is not None:
                acc = coeff
            for i in num:
                acc = i if acc is None else T.mul(acc, i)
            if acc is None:
                acc = T.one
            if den:
                dacc = den[0]
                for i in den[1:]:
                    dacc = T.mul(dacc, i)
                acc = T.div(acc, dacc)
            return acc
        if isinstance(e, sm.Pow):
            return self._lower_pow(e.base, e.exp)
        f = e.func
        if f in _UNARY:
            return T.unary(_UNARY[f], memo[e.args[0]])
        if f is sm.sign:
            return T.sign(memo[e.args[0]])
        if f is sm.atan2:
            return T.binary(ir.ATAN2, memo[e.args[0]], memo[e.args[1]])
        if f is sm.Max or f is sm.Min:
            opc = ir.MAX if f is sm.Max else ir.MIN
            ids = [memo[a] for a in e.args]
            acc = ids[0]
            for i in ids[1:]:
                acc = T.binary(opc, acc, i)
            return acc
        if f is sm.Heaviside:
            x = memo[e.args[0]]
            h0 = memo[e.args[1]] if len(e.args) > 1 else T.const(0.5)
            return T.sel(T.cmp(ir.LT, T.zero, x), T.one,
                         T.sel(T.cmp(ir.LT, x, T.zero), T.zero, h0))
        if isinstance(e, sm.Piecewise):
            # evaluated like the C printer's nested ternaries: first true
            # condition wins
            acc = None
            for expr, cond in reversed(e.args):
                val = memo[expr]
                if cond is sm.true:
                    acc = val
                else:
                    if acc is None:
                        # no default branch: C printer would error; use NaN
                        acc = T.const(float('nan'))
                    acc = T.sel(memo[cond], val, acc)
            return acc
        if isinstance(e, sm.StrictLessThan):
            return T.cmp(ir.LT, memo[e.args[0]], memo[e.args[1]])
        if isinstance(e, sm.LessThan):
            return T.cmp(ir.LE, memo[e.args[0]], memo[e.args[1]])
        if isinstance(e, sm.StrictGreaterThan):
            return T.cmp(ir.LT, memo[e.args[1]], memo[e.args[0]])
        if isinstance(e, sm.GreaterThan):
            return T.cmp(ir.LE, memo[e.args[1]], memo[e.args[0]])
        if isinstance(e, sm.Equality):
            return T.cmp(ir.EQ, memo[e.args[0]], memo[e.args[1]])
        if isinstance(e, sm.Unequality):
            return T.cmp(ir.NE, memo[e.args[0]], memo[e.args[1]])
        if isinstance(e, sm.And):
            ids = [memo[a] for a in e.args]
            acc = ids[0]
            for i in ids[1:]:
                acc = T.cmp(ir.AND, acc, i)
            return acc
        if isinstance(e, sm.Or):
            ids = [memo[a] for a in e.args]
            acc = ids[0]
            for i in ids[1:]:
                acc = T.cmp(ir.OR, acc, i)
            return acc
        if isinstance(e, sm.Not):
            return T.logic_not(memo[e.args[0]])
        if f is sm.cbrt:
            return T.unary(ir.CBRT, memo[e.args[0]])
        raise NotImplementedError(
            'Cannot lower SymPy node of type {} to CUDA: {}'.format(
                type(e).__name__, e))

    def _lower_pow(self, base, exp):
        T = self.tape
        b = self.memo[base] if base in self.memo else self.lower(base)
        if exp.is_Integer:
            return T.powi(b, int(exp))
        if exp.is_Rational and exp.q == 2:
            # x**(k/2) = sqrt(x)**k  (k odd)
            r = T.unary(ir.SQRT, b)
            k = int(exp.p)
            if k == 1:
                return r
            if k == -1:
                return T.recip(r)
            if k > 0:
                return T.mul(T.powi(b, k // 2), r)
            return T.recip(T.mul(T.powi(b, (-k) // 2), r))
        ex = self.memo[exp] if exp in self.memo else self.lower(exp)
        return T.pow(b, ex)


def lower_matrix(tape, leaf_map, exprs, use_sympy_cse=True):
    """Lowers a list of SymPy expressions, returns their tape ids.

    With ``use_sympy_cse`` the expressions first go through ``sm.cse(...,
    order='none')`` exactly like the reference does before printing C
    (opty/utils.py:748-749); SymPy's matching of common sub-sums/-products is
    stronger than the tape's structural hash-consing alone.
    """
    low = Lowerer(tape, leaf_map)
    exprs = list(exprs)
    if use_sympy_cse:
        repl, reduced = sm.cse(exprs, sm.numbered_symbols('z_opty_cse_'),
                               order='none')
        for sym, sub in repl:
            low.bind(sym, low.lower(sub))
        exprs = reduced
    return [low.lower(sm.sympify(e)) for e in exprs]

"""Builds the per-node evaluation program of a collocation problem.

Input: the discretised equations of motion (a SymPy column matrix in the
``...i / ...p / ...n`` discrete symbols, opty/direct_collocation.py:2120-2156)
together with the layout of the device-side data.  Output: a
:class:`CollocationProgram` holding one tape with

- the ``M`` constraint residuals of one node
  (what opty/direct_collocation.py:2375 hands to ``ufuncify_matrix``), and
- the ``M x P`` partial derivatives of one node
  (opty/direct_collocation.py:2755, 2800), every entry classified as
  literal / node-invariant / node-varying,

plus the partition of the EOM rows into *output groups*.  A group is the unit
of work of one warp in the CUDA kernel (32 lanes = 32 consecutive nodes, all
executing the same group's straight-line code).
"""

import hashlib

from . import ir
from .lowering import lower_matrix

# Cost of one stored Jacobian column in units of one float64 operation, for
# balancing the output groups.  (12 was tried after the 50-link measurements
# and rejected: at the 10-link pendulum it pairs the dynamic equations and
# splits the cheap kinematic group, 34-35 us instead of 30.8,
# profiles/r02l_*.  Long chains of store-only phases are cut by
# MAX_PHASES_PER_GROUP in direct_collocation.prepare_program_module instead.)
STORE_COST = 2.0


class CollocationProgram(object):
    """
    Parameters
    ----------
    discrete_eom : sequence of SymPy expressions, length M
    traj_symbols : list of (sym_at_col_i, sym_at_col_i_plus_1)
        One pair per row of the device trajectory matrix.  For midpoint the
        pair is (current, next); for backward Euler it is (previous, current).
        Entries may be None if the symbol does not occur (e.g. inputs under
        backward Euler have no "previous" symbol).
    uniform_symbols : list of Symbol
        Node-invariant arguments in the order of the device ``uni`` array:
        known parameters, unknown parameters, time interval.
    wrt : list of Symbol
        Differentiation variables in reference column order
        (opty/direct_collocation.py:2719-2721, 2734-2737).
    """

    @classmethod
    def from_matrix(cls, args, expr, const=(), use_sympy_cse=True,
                    next_args=None):
        """Program that evaluates a matrix of expressions element-wise over
        arrays: the generic operator of ``ufuncify_matrix`` (opty/utils.py:
        639-670).  Non-``const`` args become rows of the trajectory matrix,
        ``const`` args uniform inputs; the ``rows x cols`` matrix entries take
        the place of the Jacobian block and there are no residual outputs.

        ``expr`` is a SymPy matrix or the ``(replacements, [matrix])`` pair
        that ``sm.cse`` returns.  ``next_args`` optionally maps an array
        argument to the symbol that stands for its value at the NEXT point
        (midpoint-rule integrands read both, :mod:`opty_b200.objective`)."""
        from .lowering import Lowerer
        self = cls.__new__(cls)
        const = tuple(const)
        next_args = next_args or {}
        T = ir.Tape()
        leaf = {}
        row = 0
        uni = 0
        for a in args:
            if a in const:
                leaf[a] = T.uin(uni)
                uni += 1
            else:
                leaf[a] = T.vin(2 * row)
                if a in next_args:
                    leaf[next_args[a]] = T.vin(2 * row + 1)
                row += 1
        self.tape = T
        self.R = row
        self.num_uniform = uni
        if isinstance(expr, tuple) and len(expr) == 2:
            repl, (mat,) = expr
            low = Lowerer(T, leaf)
            for sym, sub in repl:
                low.bind(sym, low.lower(sub))
            ids = [low.lower(e) for e in mat]
        else:
            mat = expr
            ids = lower_matrix(T, leaf, list(mat), use_sympy_cse=use_sympy_cse)
        rows, cols = mat.shape
        self.M, self.P = rows, cols
        self.K = rows * cols
        self.con = []
        self.jac = [ids[j * cols:(j + 1) * cols] for j in range(rows)]
        self._classify()
        return self

    def __init__(self, discrete_eom, traj_symbols, uniform_symbols, wrt,
                 use_sympy_cse=True, chain_rules=()):
        self.M = len(discrete_eom)
        self.P = len(wrt)
        self.K = self.M * self.P
        self.R = len(traj_symbols)
        self.num_uniform = len(uniform_symbols)

        T = ir.Tape()
        leaf = {}
        for r, (s0, s1) in enumerate(traj_symbols):
            if s0 is not None:
                leaf[s0] = T.vin(2 * r)
            if s1 is not None:
                leaf[s1] = T.vin(2 * r + 1)
        for u, s in enumerate(uniform_symbols):
            leaf[s] = T.uin(u)
        self.tape = T
        self.con = lower_matrix(T, leaf, list(discrete_eom),
                                use_sympy_cse=use_sympy_cse)
        wrt_nodes = [leaf[w] for w in wrt]
        # implicit known trajectories r(x): d r / d x is another input row
        chain = {}
        for func, var, deriv in chain_rules:
            if func in leaf and var in leaf:
                chain.setdefault(leaf[func], []).append(
                    (leaf[var], leaf[deriv]))
        rows = ir.forward_jacobian(T, self.con, wrt_nodes, chain=chain)
        zero = T.zero
        # dense M x P table of tape ids, structural zeros point at literal 0
        self.jac = [[row.get(k, zero) for k in range(self.P)] for row in rows]
        self._classify()

    # ------------------------------------------------------------------
    def _classify(self):
        T = self.tape
        outs = list(self.con) + [e for row in self.jac for e in row]
        live = T.reachable(outs)
        self.live = live
        varying = T.varying
        op = T.op
        # invariant nodes consumed by varying nodes or written as outputs
        need_inv = set()
        for i in live:
            if varying[i]:
                for o in T.operands(i):
                    if not varying[o] and op[o] != ir.CONST:
                        need_inv.add(o)
        for o in outs:
            if not varying[o] and op[o] != ir.CONST:
                need_inv.add(o)
        self.inv_nodes = sorted(need_inv)
        self.inv_index = {nid: k for k, nid in enumerate(self.inv_nodes)}
        self.inv_closure = [i for i in T.reachable(self.inv_nodes)
                            if op[i] != ir.CONST]

        n_lit = n_inv = n_var = 0
        for row in self.jac:
            for e in row:
                if op[e] == ir.CONST:
                    n_lit += 1
                elif varying[e]:
                    n_var += 1
                else:
                    n_inv += 1
        self.num_literal_entries = n_lit
        self.num_invariant_entries = n_inv
        self.num_varying_entries = n_var

    def entry_kind(self):
        """Returns an ``M*P`` list: 0 literal, 1 node-invariant, 2 varying."""
        T = self.tape
        out = []
        for row in self.jac:
            for e in row:
                if T.op[e] == ir.CONST:
                    out.append(0)
                elif T.varying[e]:
                    out.append(2)
                else:
                    out.append(1)
        return out

    # ------------------------------------------------------------------
    def group_nodes(self, rows, stop=None):
        """Node-varying tape ids (topologically sorted) that the outputs of
        EOM ``rows`` need.  Nodes in ``stop`` (derived rows computed by the
        pre-pass) are treated as inputs."""
        T = self.tape
        roots = []
        for j in rows:
            if self.con:
                roots.append(self.con[j])
            roots.extend(self.jac[j])
        ids = self._reachable_stop(roots, stop) if stop else T.reachable(roots)
        return [i for i in ids if T.varying[i] and T.op[i] != ir.VIN and
                not (stop and i in stop)]

    def range_roots(self, c0, c1):
        """Outputs owned by the column range ``[c0, c1)`` of the flattened
        ``M*P`` node block: its Jacobian entries and the residuals of the
        rows whose first column lies in the range."""
        P = self.P
        roots = []
        for col in range(c0, c1):
            j, k = divmod(col, P)
            if k == 0 and self.con:
                roots.append(self.con[j])
            roots.append(self.jac[j][k])
        return roots

    def range_nodes(self, c0, c1, stop=None):
        """Like :meth:`group_nodes` for a column range."""
        T = self.tape
        roots = self.range_roots(c0, c1)
        ids = self._reachable_stop(roots, stop) if stop else T.reachable(roots)
        return [i for i in ids if T.varying[i] and T.op[i] != ir.VIN and
                not (stop and i in stop)]

    def range_cost(self, c0, c1, stop=None):
        carved = getattr(self, 'carved', None)
        stored = (c1 - c0) if not carved else \
            (c1 - c0) - sum(carved[c0:c1])
        return (self.tape.cost(self.range_nodes(c0, c1, stop)) +
                STORE_COST * stored)

    def split_heavy(self, cparts, max_cost, col_align=2, min_cols=16,
                    stop=None):
        """Splits column ranges whose body would cost more than ``max_cost``
        into column blocks of about that cost.  Large bodies (one equation of
        the 50-link chain is 20 k operations) exhaust the register file --
        ptxas spills kilobytes per thread to local memory -- and the
        instruction caches; the partials of one equation share little beyond
        the inputs, so column blocks recompute little."""
        out = []
        for c0, c1 in cparts:
            cost = self.range_cost(c0, c1, stop)
            pieces = int(-(-cost // max_cost))
            width = c1 - c0
            if pieces <= 1 or width < 2 * min_cols:
                out.append((c0, c1))
                continue
            step = -(-width // pieces)
            step = max(min_cols, -(-step // col_align) * col_align)
            c = c0
            while c < c1:
                e = min(c1, c + step)
                if c1 - e < min_cols // 2:
                    e = c1
                out.append((c, e))
                c = e
        return out

    def _reachable_stop(self, roots, stop):
        T = self.tape
        seen = set()
        stack = list(roots)
        while stack:
            i = stack.pop()
            if i in seen:
                continue
            seen.add(i)
            if i in stop:
                continue
            stack.extend(T.operands(i))
        return sorted(seen)

    def select_derived(self, parts, min_cost=12.0, max_rows=128):
        """Node-varying sub-expressions worth computing once per node in the
        pre-pass instead of once per output group: expensive operations
        (transcendentals, divisions, roots: ``OP_COST >= min_cost``) that at
        least two groups need.  Returns tape ids, most valuable first."""
        T = self.tape
        use = {}
        for c0, c1 in parts:
            for i in self.range_nodes(c0, c1):
                if ir.OP_COST[T.op[i]] >= min_cost:
                    use[i] = use.get(i, 0) + 1
        cand = [i for i, c in use.items() if c >= 2]
        cand.sort(key=lambda i: (-(use[i] - 1) * ir.OP_COST[T.op[i]], i))
        return sorted(cand[:max_rows])

    def row_costs(self):
        T = self.tape
        return [T.cost(self.group_nodes([j])) +
                STORE_COST * self.row_store_cols(j)
                for j in range(self.M)]

    def row_store_cols(self, j):
        """Jacobian columns of EOM row ``j`` that the group bodies store
        (all ``P`` unless some were carved out as constant runs)."""
        carved = getattr(self, 'carved', None)
        if not carved:
            return self.P
        P = self.P
        return P - sum(carved[j * P:(j + 1) * P])

    def constant_runs(self, min_len=16, max_runs=32, max_doubles=24576):
        """Maximal runs of consecutive Jacobian columns (of the flattened
        ``M*P`` node block) whose entries are literals or node-invariant:
        identical for every node, so one shared-memory image can be
        replicated into all node rows by bulk copies instead of being
        recomputed and staged per node.  Runs start at an even column and
        have even length (16-byte alignment of TMA bulk copies).

        Returns a list of ``(col0, length)`` sorted by column."""
        kinds = self.entry_kind()
        K = self.K
        runs = []
        c = 0
        while c < K:
            if kinds[c] == 2:
                c += 1
                continue
            e = c
            while e < K and kinds[e] != 2:
                e += 1
            a = c + (c & 1)
            b = e - ((e - a) & 1)
            if b - a >= min_len:
                runs.append((a, b - a))
            c = e
        runs.sort(key=lambda r: -r[1])
        keep = []
        total = 0
        for a, ln in runs:
            if len(keep) >= max_runs or total + ln > max_doubles:
                continue
            keep.append((a, ln))
            total += ln
        return sorted(keep)

    def set_carved(self, runs):
        carved = [False] * self.K
        for a, ln in runs:
            for c in range(a, a + ln):
                carved[c] = True
        self.carved = carved if runs else None

    def partition_rows(self, num_groups, col_align=2, stop=None):
        """Splits the EOM rows into at most ``num_groups`` contiguous ranges
        of balanced cost.  Each group's first Jacobian column ``r0*P`` is kept
        a multiple of ``col_align`` (TMA needs 16-byte aligned tile origins).

        Returns a list of ``(r0, r1)``.
        """
        M, P = self.M, self.P
        num_groups = max(1, min(num_groups, M))
        cuts_ok = [r for r in range(1, M) if (r * P) % col_align == 0]
        if num_groups == 1 or not cuts_ok:
            return [(0, M)]
        # cost of a contiguous range counts shared work once per group: the
        # union of the rows' node sets.  Row sets are computed once; a range
        # is grown row by row from the previous one with the same first row
        # (the greedy scan below asks for exactly that sequence), so a query
        # costs the nodes it adds, not a traversal of the tape.  Operation
        # costs are dyadic numbers: the sums do not depend on the order.
        op_cost = [ir.OP_COST[o] for o in self.tape.op]
        row_sets = [frozenset(self.group_nodes([j], stop)) for j in range(M)]
        row_store = [STORE_COST * self.row_store_cols(j) for j in range(M)]
        cost_cache = {}
        chain = {}      # first row -> (end row, node set, node cost, stores)

        def rng_cost(r0, r1):
            key = (r0, r1)
            c = cost_cache.get(key)
            if c is not None:
                return c
            state = chain.get(r0)
            if state is None or state[0] > r1:
                state = (r0, set(), 0.0, 0.0)
            end, nodes, ncost, stores = state
            if end == r0 and not nodes:
                nodes = set()
            while end < r1:
                new = row_sets[end] - nodes
                ncost += sum(op_cost[i] for i in new)
                nodes |= new
                stores += row_store[end]
                end += 1
                cost_cache[(r0, end)] = ncost + stores
            chain[r0] = (end, nodes, ncost, stores)
            return cost_cache[key]

        # minimise the maximum group cost: binary search on the bound with a
        # greedy feasibility check
        total = rng_cost(0, M)
        lo, hi = total / num_groups * 0.5, total
        best = [(0, M)]
        for _ in range(24):
            mid = 0.5 * (lo + hi)
            parts = []
            r0 = 0
            ok = True
            while r0 < M:
                r1 = r0 + 1
                # extend while within the bound and the cut stays legal
                last_legal = None
                while r1 <= M:
                    if r1 == M or r1 in cuts_ok:
                        if rng_cost(r0, r1) <= mid:
                            last_legal = r1
                        else:
                            break
                    r1 += 1
                if last_legal is None:
                    # a single (legal) block already exceeds the bound
                    nxt = next((r for r in cuts_ok if r > r0), M)
                    if rng_cost(r0, nxt) > mid:
                        ok = False
                        break
                    last_legal = nxt
                parts.append((r0, last_legal))
                r0 = last_legal
            if ok and len(parts) <= num_groups:
                best = parts
                hi = mid
            else:
                lo = mid
        return best

    def stats(self):
        T = self.tape
        var_nodes = [i for i in self.live if T.varying[i] and
                     T.op[i] != ir.VIN]
        return {
            'M': self.M, 'P': self.P, 'R': self.R,
            'tape_nodes': len(T),
            'live_nodes': len(self.live),
            'varying_ops': len(var_nodes),
            'varying_cost': T.cost(var_nodes),
            'invariant_table': len(self.inv_nodes),
            'invariant_ops': len(self.inv_closure),
            'jac_literal': self.num_literal_entries,
            'jac_invariant': self.num_invariant_entries,
            'jac_varying': self.num_varying_entries,
        }


def source_hash(text):
    return hashlib.sha256(text.encode()).hexdigest()

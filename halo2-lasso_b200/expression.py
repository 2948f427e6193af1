"""Host mirror of the reference's expression layer for the generic sum-check path:

    Expression / Query / Rotation / CommonPolynomial   pb/util/expression.rs:13-182, 488-560
    BooleanHypercube                                   pb/util/arithmetic/bh.rs:76-153
    vanilla_plonk_expression / compose (incl. LogUp)   pb/backend/hyperplonk/util.rs:30-98,
                                                       pb/backend/hyperplonk/preprocessor.rs:25-60, 111-170
    compile()  (the role of ExpressionRegistry,        pb/util/expression/evaluator.rs:22-228)

`compile` folds challenges into constants, shares common sub-expressions and emits a straight-line
program over "slots" for the bytecode-interpreting round kernel (csrc/generic.cu). Every leaf becomes a dense
table on the device: polynomial queries (rotated ones are gathered through the LFSR map), eq_xy tables, the
identity polynomial and one-hot Lagrange tables — the round polynomial values, and hence the transcript, are
the same field elements the reference's registry-based evaluator produces.
"""
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001

PRIMITIVES = [1, 3, 7, 11, 19, 37, 67, 131, 285, 529, 1033, 2053, 4179, 8219, 16427, 32771, 65581, 131081, 262183,
              524327, 1048585, 2097157, 4194307, 8388641, 16777243, 33554441, 67108935, 134217767, 268435465,
              536870917, 1073741907, 2147483657]


class BooleanHypercube:
    """Rows in LFSR order: 0, 1, x, x^2, ... in GF(2^n) (bh.rs)."""

    def __init__(self, num_vars):
        assert num_vars < 32
        self.num_vars, self.primitive = num_vars, PRIMITIVES[num_vars]
        self.x_inv = self.primitive >> 1

    def next(self, b):
        b <<= 1
        return b ^ ((b >> self.num_vars) * self.primitive)

    def prev(self, b):
        return (b >> 1) ^ ((b & 1) * self.x_inv)

    def rotate(self, b, rotation):
        for _ in range(rotation):
            b = self.next(b)
        for _ in range(-rotation):
            b = self.prev(b)
        return b

    def iter(self):
        out, b = [0], 1
        while len(out) < 1 << self.num_vars:
            out.append(b)
            b = self.next(b)
        return out

    def nth(self, i):
        """i-th row in LFSR order without walking the whole cycle; negative i counts from the end
        (`i.rem_euclid(1 << num_vars)`, classic.rs:50-53)."""
        size = 1 << self.num_vars
        i %= size
        if i == 0:
            return 0
        b = 1
        if i - 1 <= size - 1 - i:
            for _ in range(i - 1):
                b = self.next(b)
        else:  # position size-j is prev^j(1): the non-zero rows form a cycle of length size-1
            for _ in range(size - i):
                b = self.prev(b)
        return b

    def nth_map(self):
        m = [0] * (1 << self.num_vars)
        for nth, b in enumerate(self.iter()):
            m[b] = nth
        return m


class Expression:
    """Immutable AST node; `node` is a nested tuple so that == is the reference's structural equality."""

    __slots__ = ("node",)

    def __init__(self, node):
        self.node = node

    # constructors (expression.rs:67-106)
    @staticmethod
    def constant(v):
        return Expression(("const", v % R_MOD))

    @staticmethod
    def zero():
        return Expression.constant(0)

    @staticmethod
    def one():
        return Expression.constant(1)

    @staticmethod
    def identity():
        return Expression(("identity",))

    @staticmethod
    def lagrange(i):
        return Expression(("lagrange", i))

    @staticmethod
    def eq_xy(idx):
        return Expression(("eq", idx))

    @staticmethod
    def polynomial(poly, rotation=0):
        return Expression(("poly", poly, rotation))

    @staticmethod
    def challenge(idx):
        return Expression(("chal", idx))

    @staticmethod
    def distribute_powers(exprs, base):
        exprs = list(exprs)
        assert exprs
        if len(exprs) == 1:
            return exprs[0]
        return Expression(("dpow", tuple(e.node for e in exprs), base.node))

    # operators (expression.rs:488-560)
    def __neg__(self):
        return Expression(("neg", self.node))

    def __add__(self, o):
        return Expression(("sum", self.node, _lift(o).node))

    def __sub__(self, o):
        return Expression(("sum", self.node, ("neg", _lift(o).node)))

    def __mul__(self, o):
        if isinstance(o, int):
            return Expression(("scaled", self.node, o % R_MOD))
        return Expression(("prod", self.node, o.node))

    def __eq__(self, o):
        return isinstance(o, Expression) and self.node == o.node

    def __hash__(self):
        return hash(self.node)

    def degree(self):  # expression.rs:171-182
        def deg(n):
            k = n[0]
            if k in ("const", "chal"):
                return 0
            if k in ("identity", "lagrange", "eq", "poly"):
                return 1
            if k in ("neg", "scaled"):
                return deg(n[1])
            if k == "sum":
                return max(deg(n[1]), deg(n[2]))
            if k == "prod":
                return deg(n[1]) + deg(n[2])
            return max(deg(c) for c in n[1]) + deg(n[2])

        return deg(self.node)

    def leaves(self):
        """Ordered unique leaves that become device tables."""
        out = []

        def walk(n):
            k = n[0]
            if k in ("identity", "lagrange", "eq", "poly"):
                if n not in out:
                    out.append(n)
            elif k in ("neg", "scaled"):
                walk(n[1])
            elif k in ("sum", "prod"):
                walk(n[1])
                walk(n[2])
            elif k == "dpow":
                for c in n[1]:
                    walk(c)
                walk(n[2])

        walk(self.node)
        return out


def serialize_expression(expr, tokens, consts):
    """Append the prefix-token form of `expr` (include/b200_lasso.h, b200_sumcheck_prove_expression) to `tokens`;
    constants are appended to `consts` as canonical ints."""

    def const_idx(v):
        consts.append(v % R_MOD)
        return len(consts) - 1

    def walk(n):
        k = n[0]
        if k == "const":
            tokens.extend([0, const_idx(n[1])])
        elif k == "identity":
            tokens.append(1)
        elif k == "lagrange":
            tokens.extend([2, n[1]])
        elif k == "eq":
            tokens.extend([3, n[1]])
        elif k == "poly":
            tokens.extend([4, n[1], n[2]])
        elif k == "chal":
            tokens.extend([5, n[1]])
        elif k == "neg":
            tokens.append(6)
            walk(n[1])
        elif k in ("sum", "prod"):
            tokens.append(7 if k == "sum" else 8)
            walk(n[1])
            walk(n[2])
        elif k == "scaled":
            tokens.extend([9, const_idx(n[2])])
            walk(n[1])
        else:  # dpow
            tokens.extend([10, len(n[1])])
            for c in n[1]:
                walk(c)
            walk(n[2])

    walk(expr.node)
    return tokens, consts


def product(exprs):
    exprs = list(exprs)
    acc = exprs[0]
    for e in exprs[1:]:
        acc = acc * e
    return acc


def _lift(o):
    return o if isinstance(o, Expression) else Expression.constant(o)


OP_ADD, OP_SUB, OP_MUL, OP_NEG = 0, 1, 2, 3


def compile_expression(expr, challenges):
    """-> (leaves, consts, ops): slots [0, K) = leaf tables, [K, K+C) = constants, then one slot per op.
    ops are (opcode, dst, a, b); the value of the expression is the dst of the last op (or a leaf/constant)."""
    leaves = expr.leaves()
    consts, ops, memo = [], [], {}
    K = len(leaves)

    def const_slot(v):
        v %= R_MOD
        key = ("c", v)
        if key not in memo:
            consts.append(v)
            memo[key] = ("const", len(consts) - 1)
        return memo[key]

    def emit(op, a, b=None):
        if op in (OP_ADD, OP_MUL) and b is not None and b < a:  # canonical operand order (evaluator.rs:185-189)
            a, b = b, a
        key = (op, a, b)
        if key not in memo:
            ops.append([op, None, a, b if b is not None else a])
            memo[key] = ("op", len(ops) - 1)
        return memo[key]

    # two passes: symbolic refs first (constants are numbered while walking), resolved to slots afterwards
    def walk(n):
        k = n[0]
        if k == "const":
            return const_slot(n[1])
        if k == "chal":
            return const_slot(challenges[n[1]])
        if k in ("identity", "lagrange", "eq", "poly"):
            return ("leaf", leaves.index(n))
        if k == "neg":
            return emit(OP_NEG, walk(n[1]))
        if k == "sum":
            a, b = n[1], n[2]
            if b[0] == "neg":
                return emit(OP_SUB, walk(a), walk(b[1]))
            return emit(OP_ADD, walk(a), walk(b))
        if k == "prod":
            return emit(OP_MUL, walk(n[1]), walk(n[2]))
        if k == "scaled":
            return emit(OP_MUL, walk(n[1]), const_slot(n[2]))
        if k == "dpow":
            base = walk(n[2])
            acc, pw = walk(n[1][0]), base
            for c in n[1][1:]:
                acc = emit(OP_ADD, acc, emit(OP_MUL, pw, walk(c)))
                pw = emit(OP_MUL, pw, base)
            return acc
        raise ValueError(k)

    root = walk(expr.node)
    C = len(consts)

    def slot(ref):
        kind, i = ref
        return i if kind == "leaf" else (K + i if kind == "const" else K + C + i)

    prog = []
    for i, (op, _, a, b) in enumerate(ops):
        prog.append((op, K + C + i, slot(a), slot(b)))
    # drop ops that do not feed the root (e.g. the unused last power of a dpow base)
    live, need = set(), [slot(root)]
    by_dst = {p[1]: p for p in prog}
    while need:
        s = need.pop()
        if s in by_dst and s not in live:
            live.add(s)
            need += [by_dst[s][2], by_dst[s][3]]
    prog = [p for p in prog if p[1] in live]
    if not prog:  # the expression is a single leaf / constant: copy it through an addition with zero
        z = const_slot(0)
        C = len(consts)
        prog = [(OP_ADD, K + C, slot(root) if root[0] != "const" else K + root[1], K + z[1])]
    # liveness-based reuse of temporary slots (the kernel keeps the slot file in shared memory)
    last_use = {}
    for i, (_, d, a, b) in enumerate(prog):
        last_use[a] = i
        last_use[b] = i
    free, mapping, next_tmp, out = [], {}, K + C, []
    for i, (op, d, a, b) in enumerate(prog):
        ra, rb = mapping.get(a, a), mapping.get(b, b)
        for s in {a, b}:
            if s >= K + C and last_use[s] == i:
                free.append(mapping[s])
        if free:
            rd = free.pop()
        else:
            rd, next_tmp = next_tmp, next_tmp + 1
        mapping[d] = rd
        out.append((op, rd, ra, rb))
    return leaves, consts, out


# ---- vanilla plonk (pb/backend/hyperplonk/util.rs:30-62 + preprocessor.rs) -----------------------
def permutation_constraints(num_vars, num_poly, permutation_polys, max_degree, beta, gamma, num_builtin_witness_polys=0):
    """preprocessor.rs:111-170"""
    chunk = max_degree - 1
    nchunks = -(-len(permutation_polys) // chunk)
    perm_off = num_poly
    z_off = perm_off + len(permutation_polys) + num_builtin_witness_polys
    polys = [Expression.polynomial(i) for i in permutation_polys]
    ids = [Expression.constant(i << num_vars) + Expression.identity() for i in range(len(polys))]
    perms = [Expression.polynomial(perm_off + i) for i in range(len(polys))]
    zs = [Expression.polynomial(z_off + i) for i in range(nchunks)]
    z0_next = Expression.polynomial(z_off, 1)
    one = Expression.one()
    cons = [Expression.lagrange(1) * (zs[0] - one)] if zs else []
    for c in range(nchunks):
        sl = slice(c * chunk, (c + 1) * chunk)
        z_l, z_r = zs[c], (zs[c + 1] if c + 1 < nchunks else z0_next)
        lhs = z_l * product(p + beta * i + gamma for p, i in zip(polys[sl], ids[sl]))
        rhs = z_r * product(p + beta * s + gamma for p, s in zip(polys[sl], perms[sl]))
        cons.append(lhs - rhs)
    return nchunks, cons


def lookup_constraints(lookups, num_poly, num_permutation_polys, beta, gamma):
    """preprocessor.rs:78-109 (LogUp): per lookup  h (input+γ)(table+γ) - (table+γ) + m (input+γ)  on every row,
    plus the plain sum check Σ_b h(b) = 0. Polynomial order: ... | permutation | m polys | h polys | z polys."""
    m_off = num_poly + num_permutation_polys
    h_off = m_off + len(lookups)
    cons = []
    for i, lookup in enumerate(lookups):
        m, h = Expression.polynomial(m_off + i), Expression.polynomial(h_off + i)
        inp = Expression.distribute_powers([a for a, _ in lookup], beta)
        tab = Expression.distribute_powers([b for _, b in lookup], beta)
        cons.append(h * (inp + gamma) * (tab + gamma) - (tab + gamma) + m * (inp + gamma))
    return cons, [Expression.polynomial(h_off + i) for i in range(len(lookups))]


def compose(num_vars, constraints, num_poly, permutation_polys, num_challenges=0, max_degree=4, lookups=()):
    """preprocessor.rs:25-60: (num_permutation_z_polys, zero-check expression)."""
    beta, gamma, alpha = (Expression.challenge(num_challenges + i) for i in range(3))
    lookup_cons, lookup_sums = lookup_constraints(lookups, num_poly, len(permutation_polys), beta, gamma)
    md = max([c.degree() for c in constraints] + [c.degree() for c in lookup_cons] + [max_degree, 2])
    nz, perm = permutation_constraints(num_vars, num_poly, permutation_polys, md, beta, gamma, 2 * len(lookups))
    on_every_row = Expression.distribute_powers(list(constraints) + lookup_cons + perm, alpha) * Expression.eq_xy(0)
    return nz, Expression.distribute_powers(lookup_sums + [on_every_row], alpha)


def vanilla_plonk_expression(num_vars):
    """util.rs:51-62: polys 0 pi, 1-5 q_l q_r q_m q_o q_c, 6-8 w_l w_r w_o, 9-11 sigma, 12 z; challenges beta gamma alpha."""
    pi, q_l, q_r, q_m, q_o, q_c, w_l, w_r, w_o = (Expression.polynomial(i) for i in range(9))
    gate = q_l * w_l + q_r * w_r + q_m * w_l * w_r + q_o * w_o + q_c + pi
    nz, expr = compose(num_vars, [gate], 9, [6, 7, 8])
    assert nz == 1
    return expr

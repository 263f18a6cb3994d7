"""Small PLONKish test circuits (k = 5..10) that exercise every feature of the prover engine the Taiga circuits
use: several advice/fixed/instance columns, gates with rotations -1/0/+1, copy constraints across more columns than
one permutation chunk holds, single- and multi-column lookups, constants.  Used by the parity tests at sizes the
CPU oracle finishes in milliseconds."""
import random

from .circuit import Assignment, CircuitKeyData, ConstraintSystem, P


def standard_plonk(k=6, seed=0, n_lookups=2, wide=False):
    """a*b*q_m + a*q_l + b*q_r + c*q_o + q_c + PI = 0 style gate + add-with-rotation gate + range lookups."""
    rnd = random.Random(seed)
    cs = ConstraintSystem()
    a, b, c = cs.advice_column(), cs.advice_column(), cs.advice_column()
    extra = [cs.advice_column() for _ in range(3 if wide else 0)]
    inst = cs.instance_column()
    q_m, q_l, q_r, q_o, q_c = [cs.fixed_column() for _ in range(5)]
    s_rot = cs.selector()
    s_lk = cs.selector()
    table = cs.fixed_column()
    table2 = cs.fixed_column()
    consts = cs.fixed_column()
    for col in [a, b, c, inst] + extra:
        cs.enable_equality(col)
    cs.enable_constant(consts)

    A, B, C = cs.query(a), cs.query(b), cs.query(c)
    cs.create_gate("arith", [A * B * cs.query(q_m) + A * cs.query(q_l) + B * cs.query(q_r) + C * cs.query(q_o) + cs.query(q_c)])
    # rotation gate: c(next) = a(cur) + b(cur) * c(prev); second poly: boolean check on b
    S = cs.query(s_rot)
    cs.create_gate("rot", [S * (cs.query(c, 1) - (A + B * cs.query(c, -1))), S * (B * (1 - B)) * 3])
    if wide:
        E = [cs.query(e) for e in extra]
        cs.create_gate("wide", [cs.query(q_m) * (E[0] * E[1] * E[2] * A - cs.query(extra[0], 1))])
    # lookup 1: s_lk * a in table (table has 0..15; row 0 of table is 0 so disabled rows look up 0)
    SL = cs.query(s_lk)
    if n_lookups >= 1:
        cs.lookup([(SL * A, cs.query(table))])
    # lookup 2 (two columns compressed with theta): (s_lk*a, s_lk*b) in (table, table2) where table2 = table^2
    if n_lookups >= 2:
        cs.lookup([(SL * A, cs.query(table)), (SL * B, cs.query(table2))])

    def synthesize(asg, wseed):
        r = random.Random(wseed)
        row = 0
        pub = []
        # arithmetic rows: x*y = z, then z + 5 = w (constant via q_c), then expose w as public input through copy
        for _ in range(3):
            x, y = r.randrange(P), r.randrange(P)
            ca = asg.assign(a, row, x); asg.assign(b, row, y); cz = asg.assign(c, row, x * y)
            asg.assign(q_m, row, 1); asg.assign(q_o, row, P - 1)
            row += 1
            asg.assign(a, row, x * y); cw = asg.assign(c, row, x * y + 5)
            ca2 = (a, row)
            asg.assign(q_l, row, 1); asg.assign(q_c, row, 5); asg.assign(q_o, row, P - 1)
            asg.copy(cz, ca2)
            pub.append((x * y + 5) % P)
            asg.copy(cw, (inst, len(pub) - 1))
            row += 1
        # constant cell: a == 7 via the constants column
        cc = asg.assign(a, row, 7)
        asg.copy(cc, asg.constant_cell(7))
        row += 1
        # rotation gate region: rows row-? need c(prev): place at row+1 with prev = row
        cprev, av, bv = r.randrange(P), r.randrange(P), 1
        asg.assign(c, row, cprev)
        asg.assign(a, row + 1, av); asg.assign(b, row + 1, bv); asg.assign(c, row + 1, r.randrange(P))
        asg.assign(c, row + 2, av + bv * cprev)
        asg.enable(s_rot, row + 1)
        row += 3
        if wide:
            vals = [r.randrange(P) for _ in range(3)]
            av2 = r.randrange(P)
            for e, v in zip(extra, vals):
                asg.assign(e, row, v)
            asg.assign(a, row, av2)
            asg.assign(q_m, row, 1); asg.assign(q_o, row, 0)
            # arith gate on this row: a*b*q_m must vanish: b = 0, c free
            asg.assign(b, row, 0)
            out = asg.assign(extra[0], row + 1, vals[0] * vals[1] * vals[2] * av2)
            asg.copy(out, asg.assign(extra[2], row + 3, vals[0] * vals[1] * vals[2] * av2))
            row += 4
        # lookup rows (several, with repeats so the permuted column has runs)
        for i in range(10):
            t = r.randrange(16) if i % 3 else 3
            asg.assign(a, row, t); asg.assign(b, row, t * t); asg.enable(s_lk, row)
            # arith gate is off on these rows (all q_* are zero)
            row += 1
        asg.set_instance(inst, pub)
        return asg

    def fixed_only(asg):
        for i in range(16):
            asg.assign(table, i, i)
            asg.assign(table2, i, i * i)
        return asg

    def make(wseed=1):
        asg = Assignment(cs, k)
        fixed_only(asg)
        synthesize(asg, wseed)
        return asg

    # keygen: fixed columns + copy constraints come from a synthesis run (selectors/copies do not depend on witness values)
    kd = CircuitKeyData(cs, k, make(1), name="standard_plonk_k%d" % k)
    return kd, make

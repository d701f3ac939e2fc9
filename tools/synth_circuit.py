"""Synthetic satisfiable R1CS instances for the LegoGroth16 replays (BASELINE config 3: "2^18-constraint synthetic
R1CS").  Test / measurement infrastructure: builds crypto_b200.groth16.ConstraintMatrices plus a full assignment.

The circuit is a multiplication chain with non-trivial linear combinations (the shape of circom's nconstraints.circom
with coefficients added): constraint i is
    (c1 * x_p + c2 * x_q [+ c3]) * (c4 * x_u) = x_new
over earlier variables, so it is satisfied by construction; a fraction of the witnesses is forced to 0 / 1 / 16-bit
values the way real witnesses look (booleans, small ranges)."""
import random

from crypto_b200.groth16 import ConstraintMatrices
from crypto_b200.group import R_MODULUS


def synthetic_r1cs(num_constraints, num_public=2, seed=1, small_fraction=0.3):
    """-> (ConstraintMatrices, full_assignment).  Variables: 0 = one, 1..num_public = instance, then witnesses."""
    rng = random.Random(seed)
    ni = 1 + num_public
    assign = [1] + [rng.randrange(1, R_MODULUS) for _ in range(num_public)]
    # free witnesses the constraints draw from (some boolean / small like real circuits)
    nfree = 8
    for k in range(nfree):
        assign.append(rng.choice([0, 1, rng.randrange(1 << 16), rng.randrange(R_MODULUS)]))
    cs = ConstraintMatrices(num_instance_variables=ni, num_witness_variables=0)
    for i in range(num_constraints):
        nv = len(assign)
        p, q, u = rng.randrange(nv), rng.randrange(nv), rng.randrange(1, nv)
        if rng.random() < small_fraction:
            # boolean-style constraint: b * (1 - b) = 0 on a fresh 0/1 witness  ->  (b) * (one - b) = zero_var
            bit = rng.randrange(2)
            assign.append(bit)
            bi = len(assign) - 1
            assign.append(0)
            zi = len(assign) - 1
            cs.a.append([(1, bi)])
            cs.b.append([(1, 0), (R_MODULUS - 1, bi)])
            cs.c.append([(1, zi)])
            continue
        c1, c2, c4 = rng.randrange(1, 1 << 20), rng.choice([1, 1, rng.randrange(1, R_MODULUS)]), rng.choice([1, rng.randrange(1, 1 << 10)])
        row_a = [(c1, p), (c2, q)] if p != q else [((c1 + c2) % R_MODULUS, p)]
        if rng.random() < 0.25:
            row_a.append((rng.randrange(1, 1 << 8), 0))
        row_b = [(c4, u)]
        va = sum(c * assign[j] for c, j in row_a) % R_MODULUS
        vb = sum(c * assign[j] for c, j in row_b) % R_MODULUS
        assign.append(va * vb % R_MODULUS)
        cs.a.append(row_a)
        cs.b.append(row_b)
        cs.c.append([(1, len(assign) - 1)])
    cs.num_witness_variables = len(assign) - ni
    return cs, assign

"""CPU: the Circom .r1cs reader (crypto_b200/r1cs.py, row f4) - round trips through its own writer, the acceptance
rules and error messages of legogroth16/src/circom/r1cs_reader.rs, and (when the reference checkout is present,
i.e. in the build container only) the reference's own BLS12-381 fixture files."""
import glob
import os
import struct

import pytest

from crypto_b200 import r1cs
from crypto_b200.msm import R_MODULUS as R


def squaring_chain(n):
    """x_{i+1} = x_i * x_i: wires [1, out, x0, x1, ..]; the shape of the reference's nconstraints.circom."""
    cons = [([(2 + i, 1)], [(2 + i, 1)], [(3 + i if i + 1 < n else 1, 1)]) for i in range(n)]
    return r1cs.write_r1cs(1, 0, 1, n + 2, cons), cons


def witness_for_chain(n, x0):
    xs = [x0]
    for _ in range(n):
        xs.append(xs[-1] * xs[-1] % R)
    return [1, xs[-1]] + xs[:-1]


def test_round_trip_and_evaluation():
    data, cons = squaring_chain(6)
    f = r1cs.R1CSFile.new(data)
    assert f.version == 1 and f.header.field_size == 32 and f.header.curve == 'Bls12_381'
    assert (f.header.n_wires, f.header.n_pub_out, f.header.n_pub_in, f.header.n_prv_in, f.header.n_constraints) == (8, 1, 0, 1, 6)
    assert f.constraints == cons and f.wire_mapping == list(range(8))
    assert f.num_public == 2 and f.num_private == 6
    w = witness_for_chain(6, 7)
    assert f.is_satisfied(w)
    w[3] = (w[3] + 1) % R
    assert not f.is_satisfied(w)
    (rp, col, val), _, (rpc, colc, valc) = f.matrices()
    assert list(rp) == list(range(7)) and list(col) == [2, 3, 4, 5, 6, 7] and val.shape == (6, 32) and val[0, 0] == 1
    assert list(colc) == [3, 4, 5, 6, 7, 1]


def test_sections_in_any_order():
    cons = [([(1, 3), (2, R - 1)], [(0, 1)], [(3, 5)])]
    a = r1cs.R1CSFile.new(r1cs.write_r1cs(0, 1, 2, 4, cons, section_order=(2, 3, 1)))
    b = r1cs.R1CSFile.new(r1cs.write_r1cs(0, 1, 2, 4, cons))
    assert a.constraints == b.constraints == cons and a.header.n_pub_in == 1


@pytest.mark.parametrize('mutate,msg', [
    (lambda d: b'r1cx' + d[4:], 'Invalid magic number'),
    (lambda d: d[:4] + struct.pack('<I', 2) + d[8:], 'Unsupported version'),
    (lambda d: d[:24] + struct.pack('<I', 31) + d[28:], 'This parser only supports 32-byte fields'),
    (lambda d: d[:16] + struct.pack('<Q', 65) + d[24:88] + b'\x00' + d[88:], 'Invalid header section size'),
    (lambda d: d[:-5], 'failed to fill whole buffer'),
    (lambda d: d[:len(d) - 8 * 5] + struct.pack('<Q', 1) + d[len(d) - 8 * 5 + 8:], 'Wire 0 should always be mapped to 0'),
    (lambda d: d[:16] + struct.pack('<Q', 1 << 63) + d[24:], 'Invalid section size'),
])
def test_rejections_match_the_reference_reader(mutate, msg):
    data, _ = squaring_chain(3)
    with pytest.raises(r1cs.CircomError) as ei:
        r1cs.R1CSFile.new(mutate(data))
    assert msg in str(ei.value)


def test_other_curves_and_missing_sections_are_refused():
    bn254_r = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    with pytest.raises(r1cs.CircomError) as ei:
        r1cs.R1CSFile.new(r1cs.write_r1cs(1, 0, 1, 3, [([(1, 1)], [(2, 1)], [(0, 1)])], prime=bn254_r))
    assert ei.value.kind == 'IncompatibleWithCurve'
    with pytest.raises(r1cs.CircomError) as ei:
        r1cs.R1CSFile.new(r1cs.write_r1cs(1, 0, 1, 3, [], section_order=(1, 2)))
    assert 'wire2label' in str(ei.value)
    with pytest.raises(r1cs.CircomError) as ei:
        r1cs.R1CSFile.new_from_file('/nonexistent/x.r1cs')
    assert ei.value.kind == 'UnableToOpenR1CSFile'


REF_DIR = '/root/reference/proof_system/tests/r1cs/circom/bls12-381'


@pytest.mark.skipif(not os.path.isdir(REF_DIR), reason='reference checkout not present (GPU box)')
def test_reference_fixture_files_parse():
    files = sorted(glob.glob(os.path.join(REF_DIR, '*.r1cs')))
    assert files
    for path in files:
        f = r1cs.R1CSFile.new_from_file(path)
        assert f.header.n_constraints == len(f.constraints) and len(f.wire_mapping) == f.header.n_wires
        assert all(w < f.header.n_wires for con in f.constraints for lc in con for w, _ in lc)
        rp = f.matrices()[0][0]
        assert len(rp) == f.header.n_constraints + 1
    m2 = r1cs.R1CSFile.new_from_file(os.path.join(REF_DIR, 'multiply2.r1cs'))      # c = a * b, wires [1, c, a, b]
    assert (m2.header.n_wires, m2.header.n_constraints) == (4, 1)
    assert m2.is_satisfied([1, 33, 3, 11]) and not m2.is_satisfied([1, 34, 3, 11])

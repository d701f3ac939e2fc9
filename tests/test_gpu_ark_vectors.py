"""GPU: the backend against arkworks' own outputs (tests/golden/ark_vectors.json) through the C ABI -- skipped until
that file exists (tools/gen_ark_vectors needs cargo)."""
import numpy as np
import pytest

from tests import ark_vectors as av

pytestmark = pytest.mark.gpu


def test_gpu_msm_matches_arkworks(dg):
    inp, vec = av.load()
    for c in inp['msm_g1']:
        res = dg.normalize_batch(dg.msm(av.g1_records(c['bases']), av.scalar_records(c['scalars'])))
        assert av.g1_hex(res) == av.case(vec['msm_g1'], c['name'])['result'], c['name']
    for c in inp['msm_g2']:
        res = dg.normalize_batch(dg.msm(av.g2_records(c['bases']), av.scalar_records(c['scalars']), g2=True), g2=True)
        assert av.g2_hex(res) == av.case(vec['msm_g2'], c['name'])['result'], c['name']


def test_gpu_fixed_base_and_mul_bigint_match_arkworks(dg):
    inp, vec = av.load()
    for c in inp['fixed_base_g1']:
        exp = av.case(vec['fixed_base_g1'], c['name'])
        tbl = dg.FixedBaseTable(av.g1_records(c['point']), c['hint'])
        assert (tbl.window, tbl.num_windows) == (exp['window'], exp['num_windows'])
        aff = bytes(tbl.mul_many_normalized(av.scalar_records(c['scalars'])))
        assert [av.g1_hex(aff[96 * i:96 * i + 96]) for i in range(len(c['scalars']))] == exp['results']
        row1 = bytes(tbl.download())[96 * (1 << tbl.window):]
        assert [av.g1_hex(row1[96 * i:96 * i + 96]) for i in range(8)] == exp['table_row1_first8']
        tbl.free()
    for c in inp['mul_bigint_g1']:
        out = bytes(dg.normalize_batch(dg.batch_mul(av.g1_records(c['points']), av.scalar_records(c['scalars']))))
        assert [av.g1_hex(out[96 * i:96 * i + 96]) for i in range(len(c['scalars']))] == av.case(vec['mul_bigint_g1'], c['name'])['results']


def test_gpu_pairing_matches_arkworks(dg):
    inp, vec = av.load()
    for c in inp['pairing']:
        exp = av.case(vec['pairing'], c['name'])
        ml = dg.multi_miller_loop(av.g1_records(c['g1']), av.g2_records(c['g2']))
        assert av.fp12_hex(ml) == exp['miller_loop'], c['name']
        assert av.fp12_hex(dg.final_exponentiation(ml)) == exp['final_exponentiation'], c['name']
        assert av.fp12_hex(dg.multi_pairing(av.g1_records(c['g1']), av.g2_records(c['g2']))) == exp['final_exponentiation']


def test_gpu_serialization_matches_arkworks_inputs(dg):
    """The inputs themselves are ark-encoded: the device (de)serializers must round-trip them (this part needs no Rust)."""
    import json
    inp = json.load(open(av.INPUTS))
    hexes = av.case(inp['msm_g1'], 'n1000')['bases']
    out, st, bad = dg.deserialize_points(b''.join(bytes.fromhex(x) for x in hexes), compressed=True, validate=True)
    assert bad == 0 and bytes(out) == bytes(av.g1_records(hexes))
    assert bytes(dg.serialize_points(out, compressed=True)).hex() == ''.join(hexes)

"""CPU: the oracle against arkworks' own outputs (tests/golden/ark_vectors.json, tools/gen_ark_vectors) -- skipped until
that file exists.  This is what upgrades the oracle from "pinned by mathematics" to "pinned by the reference"."""
import numpy as np

from tests import ark_vectors as av


def test_oracle_msm_matches_arkworks(cref):
    inp, vec = av.load()
    for c in inp['msm_g1']:
        res = cref.msm_g1(av.g1_records(c['bases']), av.scalar_records(c['scalars']))
        assert av.g1_hex(cref.normalize_batch_g1(res)) == av.case(vec['msm_g1'], c['name'])['result'], c['name']
    for c in inp['msm_g2']:
        res = cref.msm_g2(av.g2_records(c['bases']), av.scalar_records(c['scalars']))
        assert av.g2_hex(cref.normalize_batch_g2(res)) == av.case(vec['msm_g2'], c['name'])['result'], c['name']


def test_oracle_fixed_base_and_mul_bigint_match_arkworks(cref):
    inp, vec = av.load()
    for c in inp['fixed_base_g1']:
        exp = av.case(vec['fixed_base_g1'], c['name'])
        out, w, nw = cref.fixed_base_mul_many_g1(av.g1_records(c['point']), c['hint'], av.scalar_records(c['scalars']))
        assert (w, nw) == (exp['window'], exp['num_windows'])
        aff = bytes(cref.normalize_batch_g1(out))
        assert [av.g1_hex(aff[96 * i:96 * i + 96]) for i in range(len(c['scalars']))] == exp['results']
        tbl = bytes(cref.fixed_base_table_g1(av.g1_records(c['point']), w))
        row1 = tbl[96 * (1 << w):]
        assert [av.g1_hex(row1[96 * i:96 * i + 96]) for i in range(8)] == exp['table_row1_first8']
    for c in inp['mul_bigint_g1']:
        out = bytes(cref.normalize_batch_g1(cref.batch_mul_g1(av.g1_records(c['points']), av.scalar_records(c['scalars']))))
        assert [av.g1_hex(out[96 * i:96 * i + 96]) for i in range(len(c['scalars']))] == av.case(vec['mul_bigint_g1'], c['name'])['results']


def test_oracle_pairing_matches_arkworks(cref):
    """Pins the Miller-loop VALUE (pre final exponentiation) and the GT convention to arkworks."""
    inp, vec = av.load()
    for c in inp['pairing']:
        exp = av.case(vec['pairing'], c['name'])
        ml = cref.multi_miller_loop(av.g1_records(c['g1']), av.g2_records(c['g2']))
        assert av.fp12_hex(ml) == exp['miller_loop'], c['name']
        assert av.fp12_hex(cref.final_exp(ml)) == exp['final_exponentiation'], c['name']

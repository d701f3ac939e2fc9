"""GPU: BASELINE.json configs 3 and 4 replayed at reduced size through the C ABI (the same
drivers tools/replay_workloads.py times at full size): LegoGroth16-prover-shaped G1/G2 MSMs over
witness-like scalars (30 % zeros / ones / 16-bit values), BBS+-shaped MSM + 2-pair verification
check, lazy-checker-sized multi-pairing, accumulator batch witness update."""
import pytest

from tools import replay_workloads as rw

pytestmark = pytest.mark.gpu


def test_config3_legogroth16_prover_shape(dg):
    res = rw.config3(11, with_cpu=False)
    assert len(res['msms']) == 5 and all(m['ok'] for m in res['msms'])
    assert [m['group'] for m in res['msms']].count('G2') == 1


def test_config4_bbs_plus_and_accumulator_shape(dg):
    res = rw.config4(500, with_cpu=False)
    assert res['sign_msm']['ok'] and res['verify_2pair_check']['ok']
    assert res['multi_pairing_256']['ok'] and res['witness_update']['ok']


def test_next_f2_crs_generator_shape(dg):
    res = rw.config_generator(9, with_cpu=False)
    assert len(res['tables']) == 6 and all(t['ok'] for t in res['tables'])


def test_next_f3_snarkpack_shape(dg):
    res = rw.config_snarkpack(16, with_cpu=True)
    assert all(r['ok'] for r in res['rounds']) and res['z_c_msm']['ok']
    assert res['pairs_total'] == 10 * 15

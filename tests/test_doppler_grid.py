"""a1 (dem_base:130-165): the Doppler grid anchors SURVEY.md records from the reference's own statements
(section 7 step 1 probe: bench_GMSK, N = 2^15 -> s_d[0..2] = 5858, 5932, 6006, s_d[63] = 10526; section 8(d) C1:
CC11xx.json -> bins 7698 ... 12782, step ~ 80.7).  The oracle's grid is checked here on the CPU; the GPU test below
checks that the product class builds the very same table."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import RADIO, conf_variant

ANCHORS = [
    # config, blockSize override, D, first three shifts, last shift, mean step
    ("benchmark/bench_GMSK.json", 15, 64, (5858, 5932, 6006), 10526, 74.1),
    ("CC11xx.json", None, 64, (7698, 7779, 7859), 12782, 80.7),
    ("c2_base_2p18_256bins.json", None, 256, (46867, 47013, 47159), 84205, 146.4),
    ("c4_sband_2p20_4096bins.json", None, 4096, (52422, 52525, 52627), 471866, 102.4),
]


@pytest.mark.parametrize("cfg,bs,D,head,last,step", ANCHORS)
def test_oracle_grid_anchors(cfg, bs, D, head, last, step):
    conf = conf_variant(cfg, blockSize=bs)
    N = 2 ** conf["GPU"]["UHF"]["blockSize"]
    g = O.doppler_grid(conf, RADIO, N)
    s = g["shifts"]
    assert s.dtype == np.int32 and len(s) == D and g["element_offset"] == 0
    assert tuple(int(v) for v in s[:3]) == head and int(s[-1]) == last
    assert abs(float(np.diff(s.astype(np.int64)).mean()) - step) < 0.05
    # doppHzLUT = doppIdxNorm * fs (dem_base:163): the table the Hz interpolation reads
    cr = conf["Radios"]["Rx"][RADIO]
    fs = cr["baud"] * cr["samplesPerSym"]
    np.testing.assert_allclose(g["doppHzLUT"] / fs * N, s, atol=0.5 + 1e-9)


def test_noise_row_is_prepended():
    conf = conf_variant("benchmark/bench_GMSK.json", blockSize=15, noise_measure_offset_Hz=30000.0)
    g = O.doppler_grid(conf, RADIO, 2 ** 15)
    assert g["element_offset"] == 1 and len(g["shifts"]) == 65
    assert int(g["shifts"][0]) == int(np.round(30000.0 / 153600.0 * 2 ** 15))
    assert tuple(int(v) for v in g["shifts"][1:4]) == (5858, 5932, 6006)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,bs,D,head,last,step", ANCHORS[:3])
def test_product_class_builds_the_same_grid(cfg, bs, D, head, last, step):
    from pycusdr_b200.demodulator import UHF
    from tests.helpers import protocol_for
    conf = conf_variant(cfg, blockSize=bs)
    dem = UHF.Demodulator(conf, protocol_for(conf), RADIO)
    g = O.doppler_grid(conf, RADIO, dem.Nfft)
    np.testing.assert_array_equal(dem.doppCyperSymNorm, g["shifts"])
    np.testing.assert_array_equal(dem.doppHzLUT, g["doppHzLUT"])
    assert tuple(int(v) for v in dem.doppCyperSymNorm[:3]) == head and int(dem.doppCyperSymNorm[-1]) == last

"""CPU: the C-ABI library loads and exports every symbol include/pvae_b200.h declares (no compute
calls - there is no GPU here), and the host-side mirror builds the reference's initial state."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from poissonvae_b200 import _lib
    return _lib.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pvae_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pvae_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pvae_b200.h but not exported"


def test_binding_covers_header(lib_path):
    from poissonvae_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()
    lib = _lib.load()
    assert lib.pvae_version() >= 100
    assert lib.pvae_num_partials(0) == 0 and lib.pvae_num_partials(4096) >= 512


def test_step_descriptor_layout_matches_the_header(lib_path):
    import ctypes as C

    from poissonvae_b200 import _lib
    assert _lib.load().pvae_lin_step_desc_size() == C.sizeof(_lib.LinStepDesc)


def test_argument_validation_needs_no_gpu(lib_path):
    from poissonvae_b200 import _lib
    lib = _lib.load()
    # rejected before any CUDA call: negative temperature, unknown indicator, K not a multiple of 4
    assert lib.pvae_rsample_fwd(None, None, None, None, 4, 8, 0, 4, -1.0, 0, -1.0, 0.0, 1, 0, None, None) == -1
    assert b"non-neg" in lib.pvae_last_error()
    assert lib.pvae_rsample_fwd(None, None, None, None, 4, 8, 0, 4, 1.0, 7, -1.0, 0.0, 1, 0, None, None) == -1
    assert lib.pvae_rsample_fwd(None, None, None, None, 4, 6, 0, 4, 1.0, 0, -1.0, 0.0, 1, 0, None, None) == -2
    with pytest.raises(ValueError):
        _lib.check(-1)
    with pytest.raises(NotImplementedError):
        _lib.check(-2)


def test_mirror_initial_state_is_the_reference_s():
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "vae_step_*.npz"))):
        g = np.load(path)
        K = g["p0_log_rate"].shape[1]
        model = PoissonVAE(ConfigPoisVAE(n_latents=K, prior_clamp=-4, exc_only=bool(g["exc_only"]),
                                         seed=int(g["cfg_seed"])))
        sd = model.state_dict()
        assert list(sd) == ["log_rate", "temp", "n_exp", "fc_enc.weight", "fc_dec.weight"]
        for k in ("log_rate", "fc_enc.weight", "fc_dec.weight"):
            assert np.array_equal(sd[k].numpy(), g["init_" + k]), k
        assert int(sd["n_exp"]) == 263 and float(sd["temp"]) == 1.0


def test_no_cpu_fallback():
    from poissonvae_b200 import Poisson
    with pytest.raises(RuntimeError):
        Poisson(torch.zeros(4, 8), temp=1.0).rsample()


def test_host_only_shape_helpers(lib_path):
    """Workspace sizes and shape predicates are host functions: callable without a GPU."""
    from poissonvae_b200 import _lib
    lib = _lib.load()
    # fused forward kernel: K a multiple of 256 up to 1024, D in {128, 256}, at least 12 rows per SM (148 SMs assumed without a device)
    assert lib.pvae_fused_fwd_supported(4096, 512, 256) == 1
    assert lib.pvae_fused_fwd_supported(8192, 1024, 256) == 1
    assert lib.pvae_fused_fwd_supported(2000, 256, 128) == 1
    assert lib.pvae_fused_fwd_supported(1024, 512, 256) == 0  # 7 rows per SM: the pools would be too shallow
    assert lib.pvae_fused_fwd_supported(4096, 520, 256) == 0
    assert lib.pvae_fused_fwd_supported(4096, 2048, 256) == 0
    assert lib.pvae_fused_fwd_supported(4096, 512, 64) == 0
    assert lib.pvae_kl_diag_ws_floats(4096, 512) == 128 * 512 and lib.pvae_kl_diag_ws_floats(0, 512) == 0
    assert lib.pvae_gemm_residual_parts(256) == 8  # two column halves per 64-wide tile
    assert lib.pvae_lin_step_ws_floats(4096, 512, 256) > 3 * 4096 * 512
    # knobs validate their argument
    assert lib.pvae_set_pool(7) != 0 and b"pvae_set_pool" in lib.pvae_last_error()
    assert lib.pvae_set_pool(3) == 0
    assert lib.pvae_set_wave_factor(0.0) != 0 and lib.pvae_set_wave_factor(1.0) == 0


def test_average_meter_and_log_freq_config():
    """Host pieces of the statistics path (base/utils_model.py AverageMeter; cfg.log_freq, main/config_vae.py:311)."""
    from poissonvae_b200.trainer import ConfigTrainVAE, _AverageMeter
    m = _AverageMeter()
    for v in (1.0, 2.0, 6.0):
        m.update(v)
    assert m.avg == 3.0 and m.cnt == 3
    m.reset()
    m.update(5.0, n=2)
    assert m.avg == 5.0 and m.sum == 10.0
    assert ConfigTrainVAE().log_freq == 10 and ConfigTrainVAE(log_freq=3).log_freq == 3

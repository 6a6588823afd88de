"""CPU: the C-ABI shared library builds, loads, and exports every symbol include/*.h declares
(no compute calls here -- there is no GPU in the build container)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    inc = os.path.join(ROOT, "include")
    for f in sorted(os.listdir(inc)):
        if not f.endswith(".h"):
            continue
        src = open(os.path.join(inc, f)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(sa_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_symbols_are_exported_and_bound():
    from synthanatomy_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_status_calls_work_without_a_gpu():
    from synthanatomy_b200 import _lib
    lib = _lib.load()
    assert lib.sa_version() == 100
    lib.sa_launch_count_reset()
    assert lib.sa_launch_count() == 0
    assert lib.sa_last_path() in (0, 1, 2)


def test_cpu_tensors_fail_loudly():
    import torch
    from synthanatomy_b200.networks.vqvae import B200VQVAE
    net = B200VQVAE(n_levels=1, downsample_parameters=((4, 2, 1, 1),), upsample_parameters=((4, 2, 1, 0, 1),),
                    n_embed=16, embed_dim=8, n_channels=8, n_res_channels=8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.rand(1, 1, 8, 8, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "synthanatomy_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} mentions the oracle"


def test_state_dict_matches_reference_keys():
    import torch
    from synthanatomy_b200.networks.vqvae import B200VQVAE, get_vqvae_network
    from tests import golden_util as gu
    for name in ("vqvae_cfg1", "vqvae_l2"):
        cfg, sd, _ = gu.vqvae_case(name)
        net = B200VQVAE(**cfg)
        assert set(net.state_dict()) == set(sd)
        for k, v in net.state_dict().items():
            assert tuple(v.shape) == tuple(sd[k].shape), k
    cfg = dict(network="baseline_vqvae", no_levels=4, downsample_parameters=((4, 2, 1, 1),) * 4,
               upsample_parameters=((4, 2, 1, 0, 1),) * 4, num_embeddings=(2048,), embedding_dim=(32,),
               commitment_cost=(0.25,), no_channels=256, no_res_layers=3, dropout=0.0, decay=(0.5,),
               use_subpixel_conv=False)
    with torch.device("meta"):
        net = get_vqvae_network(cfg)
    assert sum(p.numel() for p in net.parameters() if p.requires_grad) == 28123937
    with pytest.raises(ValueError):
        get_vqvae_network(dict(cfg, network="nope"))

"""Network registry: mirror of /root/reference/src/networks/vqvae/configure.py:14-39.  The reference's single
registry value ``baseline_vqvae`` resolves to the B200 implementation, constructed from the same config keys."""
from __future__ import annotations

from enum import Enum

from .b200 import B200VQVAE
from .vqvae import VQVAEBase


class VQVAENetworks(Enum):
    BASELINE_VQVAE = "baseline_vqvae"
    B200_VQVAE = "b200_vqvae"  # explicit alias


def get_vqvae_network(config: dict) -> VQVAEBase:
    if config["network"] in (VQVAENetworks.BASELINE_VQVAE.value, VQVAENetworks.B200_VQVAE.value):
        network = B200VQVAE(
            n_levels=config["no_levels"],
            downsample_parameters=config["downsample_parameters"],
            upsample_parameters=config["upsample_parameters"],
            n_embed=config["num_embeddings"][0],
            embed_dim=config["embedding_dim"][0],
            commitment_cost=config["commitment_cost"][0],
            n_channels=config["no_channels"],
            n_res_channels=config["no_channels"],
            n_res_layers=config["no_res_layers"],
            p_dropout=config["dropout"],
            vq_decay=config["decay"][0],
            use_subpixel_conv=config["use_subpixel_conv"],
        )
    else:
        raise ValueError(
            f"VQVAE unknown. Was given {config['network']} but choices are {[v.value for v in VQVAENetworks]}."
        )
    return network

from .vqvae import VQVAEBase  # noqa: F401
from .b200 import B200VQVAE  # noqa: F401
from .configure import VQVAENetworks, get_vqvae_network  # noqa: F401

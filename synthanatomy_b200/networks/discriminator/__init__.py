from .b200 import B200Discriminator, weights_init  # noqa: F401

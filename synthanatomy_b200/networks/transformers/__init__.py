from .img2seq_ordering import Ordering, OrderingTransformations, OrderingType  # noqa: F401
from .performer import Performer  # noqa: F401
from .transformer import TransformerBase  # noqa: F401

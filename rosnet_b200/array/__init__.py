from .block import BlockArray  # noqa: F401
from .maybe import MaybeArray  # noqa: F401

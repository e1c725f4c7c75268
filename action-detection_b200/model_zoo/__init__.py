# Same import surface as the reference's model_zoo/__init__.py:1-3, restricted to the backbone this
# hot path covers (the other backbones are out of scope).
from .bninception import BNInception  # noqa: F401
